"""Multi-GPU data path: contiguous shards of independent OCP instances, one process per GPU, and
one gather of the solutions and statistics.  The solve itself has no exchange step,
so this is the only collective of the path (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Instances [lo, hi) of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_stats(out):
    """[B, 4] fp64 record (f, kkt_err, iters, status) of a solve_batch result."""
    return torch.stack((out["f"], out["kkt"], out["iters"].to(torch.float64), out["status"].to(torch.float64)), dim=1)


def gather_results(out, total, rank, world, buffers=None):
    """All-gather of x and the packed statistics over the ranks; every rank receives the
    results of all `total` instances in instance order.  `buffers` (from `gather_buffers`)
    avoids per-call allocation inside timed loops."""
    x, st = out["x"], pack_stats(out)
    if world == 1:
        gx, gs = x, st
    else:
        base, rem = divmod(int(total), int(world))
        cap = base + (1 if rem else 0)
        if buffers is None:
            buffers = gather_buffers(total, world, x.shape[1], x.device)
        gx_p, gs_p = buffers
        if x.shape[0] < cap:   # ragged last shards: pad to the common capacity
            x = torch.cat((x, x.new_zeros((cap - x.shape[0], x.shape[1]))))
            st = torch.cat((st, st.new_zeros((cap - st.shape[0], 4))))
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(gx_p, x.contiguous())
            dist.all_gather_into_tensor(gs_p, st.contiguous())
        else:
            dist.all_gather(list(gx_p.view(world, cap, -1).unbind(0)), x.contiguous())
            dist.all_gather(list(gs_p.view(world, cap, 4).unbind(0)), st.contiguous())
        if rem:
            keep = torch.cat([torch.arange(r * cap, r * cap + shard_range(total, r, world)[1] - shard_range(total, r, world)[0])
                              for r in range(world)]).to(gx_p.device)
            gx, gs = gx_p[keep], gs_p[keep]
        else:
            gx, gs = gx_p, gs_p
    return {"x": gx, "f": gs[:, 0], "kkt": gs[:, 1], "iters": gs[:, 2].to(torch.int32), "status": gs[:, 3].to(torch.int32)}


def gather_buffers(total, world, n, device):
    base, rem = divmod(int(total), int(world))
    cap = base + (1 if rem else 0)
    return (torch.empty((world * cap, n), dtype=torch.float64, device=device),
            torch.empty((world * cap, 4), dtype=torch.float64, device=device))
