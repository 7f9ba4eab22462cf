"""Host-side plumbing for the one path that shards across GPUs: a batch of independent filters
(BASELINE config 3) split by filter index, one process per GPU.  There is no data-path collective:
torch.distributed carries only the barrier, the max-over-ranks timing and — for reporting — an
all-gather of the 14-entry camera states.  Works on NCCL (GPU box) and gloo (CPU tests)."""
import numpy as np


def ensemble_slice(rank: int, world: int, filters_per_rank: int):
    """Global filter indices owned by `rank` (weak scaling: every rank owns filters_per_rank)."""
    if not (0 <= rank < world) or filters_per_rank < 1:
        raise ValueError("bad rank / world / filters_per_rank")
    return range(rank * filters_per_rank, (rank + 1) * filters_per_rank)


def ensemble_camera_states(mu14, global_indices, seed=4242, sigma=2e-4):
    """Per-hypothesis perturbed camera states (n, 14) for the given GLOBAL filter indices.  Filter g
    always receives the same draw regardless of how the ensemble is sharded."""
    mu14 = np.asarray(mu14, dtype=np.float64)
    out = np.tile(mu14, (len(global_indices), 1))
    for row, g in enumerate(global_indices):
        rng = np.random.default_rng([seed, int(g)])
        d = rng.normal(scale=sigma, size=13)
        out[row, 0:13] += d
        out[row, 3:7] /= np.linalg.norm(out[row, 3:7])
    return out


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Timing rule of bench.py: a multi-GPU number is the max over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_camera_states(mu14_local, dist=None, device=None):
    """(B_local, 14) per rank -> (world * B_local, 14) on every rank, in global filter order."""
    import torch
    a = torch.as_tensor(np.ascontiguousarray(mu14_local, dtype=np.float64))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return a.numpy()
    a = a.to(device or "cpu")
    parts = [torch.empty_like(a) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, a)
    return torch.cat(parts, dim=0).cpu().numpy()


def attach_row_partition(filt, dist, device):
    """Attaches a VSlamFilter replica to a fresh NCCL communicator spanning the torch.distributed
    world: rank 0 creates the id, torch.distributed broadcasts its 128 bytes, every rank joins.
    Afterwards the filter's stacked update is partitioned by covariance row blocks (BASELINE config 4)."""
    import torch
    from . import filter as _f
    rank, world = dist.get_rank(), dist.get_world_size()
    _f.load_nccl()
    raw = _f.nccl_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    filt.dist_attach(bytes(t.cpu().tolist()), rank, world)
    return rank, world


def row_block(rank: int, world: int, n: int):
    """Rows [r0, r1) of an n x n covariance owned by `rank` (same rule as ekf_api.cu::stacked_update)."""
    rpr = ((n + world - 1) // world + 31) & ~31
    r0 = min(n, rank * rpr)
    return r0, min(n, r0 + rpr), rpr
