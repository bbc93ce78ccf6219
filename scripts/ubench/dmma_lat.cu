// Micro-benchmark (development aid): latency / per-warp throughput of mma.sync.m8n8k4.f64 (DMMA) and DFMA on one SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k_dmma(long long* out, int iters) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { acc[i][0] = 0; acc[i][1] = 0; }
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (s == 12345.678) out[0] = 0;
}
template <int NACC>
__global__ void k_dfma(long long* out, int iters) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
  double a = 1.0 - 1e-9, b = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (s == 12345.678) out[0] = 0;
}
template <class K> void run(const char* name, K kern, int nacc, int threads, int iters, long long* d) {
  kern<<<1, threads>>>(d, iters); cudaDeviceSynchronize();
  kern<<<1, threads>>>(d, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-6s acc=%d warps=%d: %.1f cycles per op per warp (%.1f per round)\n", name, nacc, threads / 32, (double)h / iters / nacc, (double)h / iters);
}
int main() {
  long long* d; cudaMalloc(&d, 1024);
  const int it = 2000;
  for (int th : {32, 128, 256, 512}) {
    run("dmma", k_dmma<1>, 1, th, it, d); run("dmma", k_dmma<2>, 2, th, it, d); run("dmma", k_dmma<4>, 4, th, it, d);
    run("dmma", k_dmma<6>, 6, th, it, d); run("dmma", k_dmma<8>, 8, th, it, d);
  }
  for (int th : {32, 128, 256}) {
    run("dfma", k_dfma<1>, 1, th, it, d); run("dfma", k_dfma<4>, 4, th, it, d); run("dfma", k_dfma<8>, 8, th, it, d);
  }
  return 0;
}
