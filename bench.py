#!/usr/bin/env python
"""Benchmark of the hot path: batched solves of BoundMPC's per-step OCP.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4], the configuration the metric's target is quoted on): the
65,536 mixed experiment1 / experiment2 instances, sharded contiguously, 8,192 instances per
GPU (weak scaling: N GPUs solve 8,192 N instances; N = 8 is the full config).  A "step" is one
pass over the shard: one launch of the solver kernel on inputs already resident in HBM, plus —
for N > 1 — the NCCL all-gather of the solutions and statistics.  `value` = instances of all
ranks / device time (CUDA events on the launching stream, max over ranks).  `e2e` is the same
step through the host-pointer C-ABI entry (`bmpc_solve_batch_host`: pinned host buffers,
host-to-device and device-to-host copies inside the timed region).

`--impl reference` times the CPU restatement of the reference path (oracle/, all host threads)
on a bounded sample of the same workload: CasADi / Ipopt / MUMPS cannot be installed offline
(DESIGN.md "Oracle"), so the reference arm is the oracle port.  That arm never touches the CUDA
library or a CUDA context: its inputs come from the same generator driven by the oracle.

`--config NAME` selects another BASELINE.json workload for the B200 arm (one JSON line each, kept
under profiles/): exp1_1024, exp2_8192, exp1_N20_tight_8192, and spec_mixed_65536 = the default
mixed workload drawn by the literal SURVEY 8d generator (boundmpc_b200/batches.py).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "batched OCP solves/sec"
UNIT = "solves/s"
PER_GPU = 8192
# Termination tolerance of both arms: the reference's own Ipopt setting (`'tol': 10e-6`, BoundMPC.py:121).  The B200 arm also
# reports the tight setting the parity tests run at (1e-9: the converged KKT point, 1e-6 relative in the joint trajectory)
# as `tight_tol` in the same line; `--tol` selects another value for both arms.
REF_TOL = 1e-5
TIGHT_TOL = 1e-9
# algorithmic flops of one interior-point iteration of one instance:
# N (266,517 factorisation + 18,432 solve + 12,000 evaluation)
F_ITER = {10: 2.97e6, 20: 5.94e6}
# algorithmic HBM bytes per instance: x0, p in; x, g, lam_g, lam_x, f, kkt, iters, status out
IO_BYTES = {10: 21504, 20: 38944}


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), power_w_max=float(max(power)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def workload(name, per_gpu):
    """Generator arguments, horizon and instances per GPU of a named config (batches.CONFIGS)."""
    from boundmpc_b200 import batches
    c = dict(batches.CONFIGS[name])
    count = c.pop("count")
    n = c.pop("n")
    return c, n, min(per_gpu, count), count


def load_inputs(solver, name, rank, per_gpu, workers, count=None):
    from boundmpc_b200 import batches
    gen, n, per, _ = workload(name, per_gpu)
    t = time.perf_counter()
    x0, p, scale = batches.make_batch(solver, first=rank * per, count=count or per, n=n, workers=workers, return_scales=True, **gen)
    return x0, p, scale, time.perf_counter() - t


def cpu_solves_per_s(x0, p, cores, budget_s, tol, N=10):
    """Oracle port on `cores` host threads (ctypes releases the GIL), time-boxed: every thread pulls
    the next instance until the sample or the budget is spent.  Returns (solves/s, solved, iteration mean)."""
    from oracle import oracle as O
    O.lib()
    lock = threading.Lock()
    state = {"next": 0, "done": 0, "iters": 0, "fail": 0}
    t0 = time.perf_counter()

    def work():
        while True:
            with lock:
                i = state["next"]
                if i >= len(x0) or time.perf_counter() - t0 > budget_s:
                    return
                state["next"] += 1
            r = O.solve(x0[i], p[i], N=N, tol=tol)
            with lock:
                state["done"] += 1
                state["iters"] += r["iters"]
                state["fail"] += int(r["status"] != 0)

    th = [threading.Thread(target=work) for _ in range(cores)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    el = time.perf_counter() - t0
    return state["done"] / el, state["done"], state["iters"] / max(1, state["done"]), state["fail"], el


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-gpu", type=int, default=PER_GPU)
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of wall time for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tol", type=float, default=REF_TOL, help="termination tolerance of both arms (default: the reference's ipopt.tol)")
    ap.add_argument("--config", default="mixed_65536",
                    choices=["mixed_65536", "exp1_1024", "exp2_8192", "exp1_N20_tight_8192", "spec_mixed_65536"])
    args = ap.parse_args()
    TOL = args.tol
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference" and rank != 0:
        return 0

    cores = os.cpu_count() or 1
    gen_args, NH, per_gpu, cfg_total = workload(args.config, args.per_gpu)
    n, m, npar = 44 * NH, 43 * NH, 141 + 91 * 4
    names = {"mixed_65536": "BASELINE configs[4]", "exp1_1024": "BASELINE configs[1]", "exp2_8192": "BASELINE configs[2]",
             "exp1_N20_tight_8192": "BASELINE configs[3]", "spec_mixed_65536": "BASELINE configs[4], literal SURVEY 8d generator"}
    config = {"workload": f"{args.config} ({names[args.config]}): {per_gpu} OCP instances per GPU"
                          + (" (N=8 GPUs = the full 65,536)" if cfg_total > args.per_gpu else ""),
              "instances_per_gpu": per_gpu, "total_instances": per_gpu * max(1, world), "horizon_N": NH, "nr_segs": 4,
              "n_var": n, "n_con": m, "tol": TOL,
              "tol_source": "the reference's ipopt.tol (BoundMPC.py:121)" if TOL == REF_TOL else "--tol",
              "generator": "literal SURVEY 8d (sigma_q 0.02, odd instances cold-started, widths x U(0.75, 1.25), nothing repaired)"
                           if gen_args.get("spec") else "repaired (boundmpc_b200/batches.py: sigma_q 5e-3, widths x U(1, 1.25), perturbations "
                           "that leave the error bounds halved)",
              "l2": f"inputs+outputs {per_gpu * IO_BYTES[NH] / 1e6:.0f} MB per step"
                    + (" (> 126 MB L2), no explicit flush" if per_gpu * IO_BYTES[NH] > 126e6 else ", L2 flushed between timed steps (256 MB write)"),
              "parallelism": f"independent instances, contiguous shards over {max(1, world)} GPU(s)"}

    # ------------------------------------------------------------------ reference arm (CPU restatement)
    if args.impl == "reference":
        # no CUDA library, no CUDA context: the inputs are generated by the same generator with the oracle as its solver
        from oracle import oracle as O
        O.lib()
        sample = min(per_gpu, max(32 * cores, 256))
        x0, p, _, t_gen = load_inputs(O.OracleSolver(NH, 4, 0.1, TIGHT_TOL), args.config, 0, args.per_gpu, workers=max(1, min(16, cores)), count=sample)
        for _ in range(args.warmup):
            cpu_solves_per_s(x0[:4 * cores], p[:4 * cores], cores, 1e9, TOL, N=NH)
        t0 = time.perf_counter()
        solved = iters = fails = 0
        for k in range(args.steps):
            _, d, itm, fl, _ = cpu_solves_per_s(x0, p, cores, 1e9, TOL, N=NH)
            solved += d; iters += itm * d; fails += fl
        el = time.perf_counter() - t0
        v = solved / el
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"first {sample} instances of the workload ({sample // cores} per thread) per step x {args.steps} steps, "
                                           f"oracle/ interior-point port (-O3 -march=native) at tol {TOL:g} (CasADi/Ipopt not installable "
                                           f"offline), mean {iters / max(1, solved):.1f} iterations, {fails} failures; inputs generated "
                                           f"on the host in {t_gen:.0f} s (untimed)"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the solver has no CPU path)")
    torch.cuda.set_device(local)
    from boundmpc_b200 import build as bld
    bld.build()
    from boundmpc_b200.ocp import default_solver
    solver = default_solver(N=NH, nr_segs=4, dt=0.1, solver_opts={"b200": {"tol": TOL}}, device=local)
    # (the workload does not depend on --tol: the nominal closed loops the instances are drawn from are always run tight)
    gen_solver = solver if TOL == TIGHT_TOL else default_solver(N=NH, nr_segs=4, dt=0.1, solver_opts={"b200": {"tol": TIGHT_TOL}}, device=local)
    x0, p, scale, t_gen = load_inputs(gen_solver, args.config, rank, args.per_gpu, workers=max(1, min(16, cores // max(1, world))))
    del gen_solver
    assert (solver.n, solver.m, solver.np) == (n, m, npar)
    config["launch_shape"] = solver.launch_shape()
    default_cfg = args.config == "mixed_65536"

    # ------------------------------------------------------------------ B200 arm
    import torch.distributed as dist
    json_fd = None
    if world > 1:
        # one JSON line on stdout and nothing else: NCCL prints its version banner on file descriptor 1 whatever
        # NCCL_DEBUG_FILE says, so fd 1 is pointed at stderr for the run and the line goes to a saved copy of it
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    peak_dfma = solver.fp64_peak(0)
    peak_dmma = solver.fp64_peak(1)
    xd, pd = torch.from_numpy(x0).to(dev), torch.from_numpy(p).to(dev)
    out = solver.solve_batch(xd, pd)
    torch.cuda.synchronize()
    from boundmpc_b200 import sharding
    total = per_gpu * world
    gbuf = sharding.gather_buffers(total, world, n, dev) if world > 1 else None
    gathered = {}

    ev_k, ev_g = [], []
    need_flush = per_gpu * IO_BYTES[NH] <= 126e6          # working set inside the 126 MB L2: flush it between timed steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if need_flush else None

    def step(timed):
        if flush is not None:
            flush.zero_()
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        solver.solve_batch(xd, pd, out)
        if timed:
            b.record()
            ev_k.append((a, b))
        if world > 1:   # NCCL gather of solutions and statistics, timed separately
            gathered.update(sharding.gather_results(out, total, rank, world, gbuf))
            if timed:
                c = torch.cuda.Event(enable_timing=True)
                c.record()
                ev_g.append((b, c))

    for _ in range(args.warmup):
        step(False)
    launches0 = solver.launch_count()
    clocks = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(True)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ck = clocks.stop()
    launches = solver.launch_count() - launches0
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_k]))
    g_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_g])) if ev_g else 0.0
    # (with the L2 flush between steps the timed quantity is the sum of the per-step event pairs, flush excluded)
    ms = torch.tensor([e0.elapsed_time(e1) if flush is None else (k_ms + g_ms) * args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    iters = out["iters"].cpu().numpy()
    status = out["status"].cpu().numpy()
    kkt = out["kkt"].cpu().numpy()
    ok = int((status == 0).sum())
    cnt = torch.tensor([ok, int(iters.sum()), per_gpu], dtype=torch.float64, device=dev)
    rank_stats = torch.tensor([k_ms, float(iters.max()), float(iters.mean()), float(ok), g_ms], dtype=torch.float64, device=dev)
    all_stats = [rank_stats.clone() for _ in range(max(1, world))]
    if world > 1:
        dist.all_reduce(cnt)
        dist.all_gather(all_stats, rank_stats)
        assert int((gathered["status"] == 0).sum().item()) == int(cnt[0].item())
    per_rank = [{"kernel_ms": float(t_[0]), "iters_max": int(t_[1]), "iters_mean": float(t_[2]), "success": int(t_[3]), "gather_ms": float(t_[4])}
                for t_ in all_stats]

    # ---- end to end through the host-pointer C-ABI entry, pinned host buffers
    def pinned(shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
    hx0, hp = pinned((per_gpu, n)), pinned((per_gpu, npar))
    hx0[:], hp[:] = x0, p
    hout = {"x": pinned((per_gpu, n)), "g": pinned((per_gpu, m)), "lam_g": pinned((per_gpu, m)), "lam_x": pinned((per_gpu, n)),
            "f": pinned((per_gpu,)), "kkt": pinned((per_gpu,)), "iters": pinned((per_gpu,), torch.int32),
            "status": pinned((per_gpu,), torch.int32)}
    for _ in range(2):
        solver.solve_batch(hx0, hp, hout)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.solve_batch(hx0, hp, hout)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_same = bool(np.array_equal(hout["x"], out["x"].cpu().numpy()))
    # same call with ordinary (pageable) numpy buffers: explicit copies around the launch instead of the kernel reading /
    # writing the page-locked buffers itself
    px0, pp = x0.copy(), p.copy()
    pout = {k: np.empty_like(v) for k, v in hout.items()}
    solver.solve_batch(px0, pp, pout)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.solve_batch(px0, pp, pout)
    e2e_pg = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_pg, op=dist.ReduceOp.MAX)
    h2d = per_gpu * (n + npar) * 8
    d2h = per_gpu * ((2 * n + 2 * m) * 8 + 8 + 8 + 4 + 4)

    # ---- the same workload converged to the tight tolerance of the parity tests (device-resident and through the host entry)
    tight = None
    if world == 1 and TOL != TIGHT_TOL:
        solver_t = default_solver(N=NH, nr_segs=4, dt=0.1, solver_opts={"b200": {"tol": TIGHT_TOL}}, device=local)
        out_t = solver_t.solve_batch(xd, pd)
        for _ in range(2):
            solver_t.solve_batch(xd, pd, out_t)
        ev_t = []
        for _ in range(args.steps):
            if flush is not None:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            solver_t.solve_batch(xd, pd, out_t)
            b.record()
            ev_t.append((a, b))
        torch.cuda.synchronize()
        t_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_t]))
        solver_t.solve_batch(hx0, hp, hout)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            solver_t.solve_batch(hx0, hp, hout)
        t_e2e = (time.perf_counter() - t0) / args.steps
        it_t, st_t = out_t["iters"].cpu().numpy(), out_t["status"].cpu().numpy()
        x_t, x_r = out_t["x"].cpu().numpy(), out["x"].cpu().numpy()
        both = (st_t == 0) & (status == 0)
        qcols = (np.arange(n) % 44 >= 8) & (np.arange(n) % 44 < 15)
        dq_rel = np.abs(x_t[both][:, qcols] - x_r[both][:, qcols]).max(axis=1) / np.maximum(1.0, np.abs(x_t[both][:, qcols]).max(axis=1))
        tight = {"tol": TIGHT_TOL, "value": per_gpu / (t_ms * 1e-3), "unit": UNIT, "kernel_ms": t_ms, "e2e": per_gpu / t_e2e,
                 "success": int((st_t == 0).sum()), "iters_mean": float(it_t.mean()), "iters_max": int(it_t.max()),
                 "roofline_frac": float(it_t.sum()) * F_ITER[NH] / (t_ms * 1e-3) / peak_dfma,
                 "q_rel_diff_to_headline": {"p50": float(np.percentile(dq_rel, 50)), "p99": float(np.percentile(dq_rel, 99)), "max": float(dq_rel.max())},
                 "what": "same inputs, same kernel, handle created with b200.tol = 1e-9 (the setting of the parity tests and of the "
                         "round-1 / early round-2 bench lines); q_rel_diff = joint trajectory of the headline solve against this one"}
        del solver_t

    # ---- single-instance latency (B = 1 through the same host entry), p50 over 64 instances
    lat = []
    for i in range(min(64, per_gpu)):
        t = time.perf_counter()
        solver.solve_batch(x0[i:i + 1], p[i:i + 1])
        lat.append((time.perf_counter() - t) * 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    def extra_sections():
        # ---- p50 latency of one MPC step on BASELINE configs[0]: the headless experiment1 closed loop (bound_mpc_node.py:
        # 292-372 restated in batches.nominal_sequence), timed like the reference times its solver call (BoundMPC.py:445-455:
        # wall clock around `self.solver(...)` + the conversion of sol['x']), one instance per call through the host entry
        from boundmpc_b200 import batches, scenarios
        _, seq_stats, _ = batches.nominal_sequence(scenarios.experiment1(n=10), solver, device_step=False)
        t_step = np.array([s_[1] for s_ in seq_stats]) * 1e3
        t_wall = np.array([s_[3] for s_ in seq_stats]) * 1e3
        it_step = np.array([s_[0] for s_ in seq_stats])
        mpc_step = {"p50": float(np.percentile(t_step, 50)), "p90": float(np.percentile(t_step, 90)), "max": float(t_step.max()),
                    "steps": int(len(t_step)), "iters_mean": float(it_step.mean()), "all_converged": bool(all(s_[2] for s_ in seq_stats)),
                    "what": "experiment1 closed loop (BASELINE configs[0]), wall clock around solver(x0, p) incl. H2D/D2H, B = 1; pre- and "
                            "post-processing by the numpy mirror",
                    "python_pre_post_ms_per_step": float((t_wall - t_step).mean())}
        # the drop-in step: BoundMPC.step() as the reference's node calls it, whole step on the device (one library call)
        for log_on in (False, True):
            _, ds, _ = batches.nominal_sequence(scenarios.experiment1(n=10), solver, device_step=True, real_time=not log_on)
            tw = np.array([s_[3] for s_ in ds]) * 1e3
            mpc_step["dropin_log" if log_on else "dropin"] = {
                "p50": float(np.percentile(tw, 50)), "p90": float(np.percentile(tw, 90)), "max": float(tw.max()), "steps": int(len(tw)),
                "iters_mean": float(np.mean([s_[0] for s_ in ds])), "all_converged": bool(all(s_[2] for s_ in ds)),
                "what": "wall clock of BoundMPC.step() (BoundMPC.py:306-506 incl. compute_return_data"
                        + (" and its logging branch" if log_on else "") + "): k_prepare -> k_solve -> k_finish in one "
                        "bmpc_mpc_step_batch_host call, B = 1"}
        # ---- batched parameter builder (the pre-solve half of BoundMPC.step), rank 0
        nb = 512
        t0 = time.perf_counter()
        D = batches.make_builder_batch(solver, ("exp1", "exp2"), 0, nb, bound_scale=True)
        t_mirror = time.perf_counter() - t0          # host mirror: controller restore + prepare() per instance (Python)
        rep = (per_gpu + nb - 1) // nb
        tile = lambda a: np.ascontiguousarray(np.concatenate([a] * rep)[:per_gpu])
        tb = {k: torch.from_numpy(tile(D[k])).to(dev) for k in ("path_id", "sector", "state", "prev")}
        tb["tables"] = torch.from_numpy(D["tables"]).to(dev)
        sec0 = tb["sector"].clone()
        bo = solver.prepare_batch(tb["tables"], tb["path_id"], tb["sector"], tb["state"], tb["prev"])
        torch.cuda.synchronize()
        perr = float((np.abs(bo["p"][:nb].cpu().numpy() - D["p"]) / np.maximum(1.0, np.abs(D["p"]))).max())
        x0_same = bool(np.array_equal(bo["x0"][:nb].cpu().numpy(), D["x0"]))
        b_ms = []
        for k in range(args.warmup + args.steps):
            tb["sector"].copy_(sec0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            solver.prepare_batch(tb["tables"], tb["path_id"], tb["sector"], tb["state"], tb["prev"], bo)
            b.record()
            torch.cuda.synchronize()
            if k >= args.warmup:
                b_ms.append(a.elapsed_time(b))
        b_ms = float(np.mean(b_ms))
        bytes_inst = 76 * 8 + 8 + n * 8 + n * 8 + npar * 8          # state, path id + sector, prev_x in; x0, p out
        hs = {k: tile(D[k]) for k in ("path_id", "sector", "state", "prev")}
        solver.prepare_batch(D["tables"], hs["path_id"], hs["sector"], hs["state"], hs["prev"])
        t0 = time.perf_counter()
        solver.prepare_batch(D["tables"], hs["path_id"], hs["sector"], hs["state"], hs["prev"])
        b_e2e = time.perf_counter() - t0
        builder = {"kernel": "k_prepare", "instances": per_gpu, "ms": b_ms, "instances_per_s": per_gpu / (b_ms * 1e-3),
                   "bytes_per_instance": bytes_inst,
                   "hbm": {"achieved": per_gpu * bytes_inst / (b_ms * 1e-3) / 1e9, "unit": "GB/s"},
                   "e2e_instances_per_s": per_gpu / b_e2e,
                   "parity": {"p_rel_err_vs_host_mirror": perr, "x0_bitwise_equal": x0_same, "checked": nb},
                   "cpu_mirror": {"instances_per_s": nb / t_mirror, "what": "boundmpc_b200.bound_mpc.BoundMPC.prepare (numpy mirror of "
                                  "BoundMPC.py:310-443) incl. controller-state restore, 1 thread"}}
        # ---- batched post-processing (compute_return_data) and the whole batched MPC step on the device:
        # k_prepare -> k_solve -> k_post on one stream, inputs (controller states, previous solutions) resident in HBM
        tb["sector"].copy_(sec0)
        so = solver.solve_batch(bo["x0"], bo["p"])
        po = solver.post_batch(tb["tables"], tb["path_id"], tb["sector"], tb["state"], so["x"])
        torch.cuda.synchronize()
        p_ms, s_ms = [], []
        for k in range(args.warmup + args.steps):
            tb["sector"].copy_(sec0)
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            solver.prepare_batch(tb["tables"], tb["path_id"], tb["sector"], tb["state"], tb["prev"], bo)
            solver.solve_batch(bo["x0"], bo["p"], so)
            b.record()
            solver.post_batch(tb["tables"], tb["path_id"], tb["sector"], tb["state"], so["x"], None, po)
            c.record()
            torch.cuda.synchronize()
            if k >= args.warmup:
                p_ms.append(b.elapsed_time(c)); s_ms.append(a.elapsed_time(c))
        p_ms, s_ms = float(np.mean(p_ms)), float(np.mean(s_ms))
        # ---- on-device closed loop: per_gpu robots from the start of both experiments, per-robot bound widths
        from boundmpc_b200.rollout import initial_state, rollout
        r_st, r_sec = [], []
        for nm in ("exp1", "exp2"):
            scn_ = scenarios.experiment1(n=10) if nm == "exp1" else scenarios.experiment2(n=10)
            m_ = batches.make_mpc(scn_, batches._BoundsOnly(solver.bounds()))
            st_, sec_, _ = initial_state(m_, scn_['q0'])
            r_st.append(st_); r_sec.append(sec_)
        r_pid = (np.arange(per_gpu) % 2).astype(np.int32)
        r_state = np.stack([r_st[k] for k in r_pid])
        r_state[:, 53:57] = np.random.default_rng(20261017).uniform(1.0, 1.25, (per_gpu, 4))
        r_args = (solver, tb["tables"], torch.from_numpy(r_pid).to(dev), torch.from_numpy(r_state).to(dev),
                  torch.from_numpy(np.array([r_sec[k] for k in r_pid], np.int32)).to(dev))
        r_steps = 16
        rollout(*r_args, 2, record=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ro = rollout(*r_args, r_steps, record=True)
        torch.cuda.synchronize()
        r_el = time.perf_counter() - t0
        roll = {"robots": per_gpu, "steps": r_steps, "wall_ms": r_el * 1e3, "mpc_steps_per_s": per_gpu * r_steps / r_el,
                "converged_frac": float((ro["status"] == 0).float().mean().item()), "iters_mean": float(ro["iters"].float().mean().item()),
                "kernel_launches_per_step": 3,
                "what": "closed loop of bound_mpc_node.py:292-372 for a batch on the device: k_prepare -> k_solve -> k_finish per step, "
                        "state resident in HBM, host only enqueues; robots start at the initial state of experiment1 / experiment2 with "
                        "bound widths x U(1, 1.25)"}
        post_bytes = 76 * 8 + 12 + n * 8 + 10 * 42 * 8 + 76 * 8       # state, ids, w in; traj, state out
        post = {"kernel": "k_post", "instances": per_gpu, "ms": p_ms, "instances_per_s": per_gpu / (p_ms * 1e-3),
                "bytes_per_instance": post_bytes, "hbm": {"achieved": per_gpu * post_bytes / (p_ms * 1e-3) / 1e9, "unit": "GB/s"},
                "mpc_step_on_device": {"ms": s_ms, "steps_per_s": per_gpu / (s_ms * 1e-3), "converged": int((so["status"] == 0).sum().item()),
                                       "what": "k_prepare + k_solve + k_post back to back on one stream for the batch (controller states and "
                                               "previous solutions resident in HBM); instances = 512 perturbed controller states tiled"}}
        return mpc_step, builder, post, roll

    mpc_step = builder = post = roll = None
    if default_cfg:
        mpc_step, builder, post, roll = extra_sections()

    sum_iters = float(cnt[1].item())
    ach = float(iters.sum()) * F_ITER[NH] / (k_ms * 1e-3)          # rank 0's kernel: flop / s
    hbm = per_gpu * IO_BYTES[NH] / (k_ms * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    if default_cfg:
        builder["hbm"].update(peak=hbm_peak, frac=builder["hbm"]["achieved"] / hbm_peak)
        post["hbm"].update(peak=hbm_peak, frac=post["hbm"]["achieved"] / hbm_peak)
    traffic = None
    for tf in ("r2b_traffic.json", "r2_traffic.json"):          # committed ncu capture of the default workload (bytes per launch)
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", tf))).get("k_solve_dram_bytes_per_launch")
            break
        except (OSError, ValueError):
            pass
    if not default_cfg or TOL != REF_TOL:
        traffic = None                                           # (the capture is of the default workload at the default tolerance)
    line = {"metric": METRIC, "value": total * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": max(1, world),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA m8n8k4 on the tensor sub-pipe, same issue rate as DFMA)", "achieved": ach / 1e12, "peak": peak_dfma / 1e12, "unit": "TFLOP/s",
                         "frac": ach / peak_dfma, "traffic": traffic,
                         "note": f"k_solve, algorithmic flops = sum of interior-point iterations x {F_ITER[NH] / 1e6:.2f} MFLOP (SURVEY 8d) / mean "
                                 "CUDA-event launch duration; peak = DFMA loop measured in this run (no FP64 entry in "
                                 f"MEASURED_PEAKS.json); DMMA m8n8k4 loop measured {peak_dmma / 1e12:.1f} TFLOP/s",
                         "kernel_ms": k_ms,
                         "hbm": {"achieved": hbm, "peak": hbm_peak, "unit": "GB/s", "frac": hbm / hbm_peak,
                                 "peak_source": "measured" if peaks else "fallback"}},
            "e2e": {"value": total * args.steps / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "bitwise_equal_to_device_path": e2e_same,
                    "buffers": "page-locked host buffers (torch pin_memory): the kernel fetches x0 / p and stores the results over PCIe "
                               "itself, instance by instance, so the transfers run under the launch",
                    "pageable": {"value": total * args.steps / float(e2e_pg.item()), "unit": UNIT,
                                 "what": "same call with pageable numpy buffers: cudaMemcpy before and after the launch"}},
            "gpu_launches": launches,
            "tight_tol": tight,
            "clocks": ck,
            "gather_ms": g_ms,
            "solver": {"success": int(cnt[0].item()), "instances": total, "iters_mean": sum_iters / total,
                       "fail_rate": 1.0 - float(cnt[0].item()) / total,
                       "iters_max_rank0": int(iters.max()), "kkt_max_rank0": float(kkt[status == 0].max()) if ok else None,
                       "iters_hist_rank0": np.bincount(np.minimum(iters, 60), minlength=61).tolist(),
                       "status_hist_rank0": {str(int(k)): int(v) for k, v in zip(*np.unique(status, return_counts=True))},
                       "per_rank": per_rank,
                       "perturbation_scale_hist_rank0": {str(v): int((scale == v).sum()) for v in np.unique(scale)},
                       "input_generation_s": t_gen},
            "latency_b1_ms": {"p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90)), "max": float(max(lat)),
                              "what": "one instance through bmpc_solve_batch_host incl. H2D/D2H, wall clock"}}
    if default_cfg:
        line.update(builder=builder, post=post, rollout=roll, latency_mpc_step_ms=mpc_step,
                    latency_step_dropin_ms=mpc_step["dropin"])
    if world == 1 and not args.no_cpu_baseline:
        v, d, itm, fl, el = cpu_solves_per_s(x0, p, cores, args.cpu_budget, TOL, N=NH)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {d} instances of the workload in {el:.1f} s, oracle/ interior-point port at tol "
                                          f"{TOL:g} on {cores} threads (CasADi/Ipopt not installable offline), mean {itm:.1f} "
                                          f"iterations, {fl} failures"}
    if json_fd is None:
        print(json.dumps(line))
    else:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
        os.close(json_fd)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
