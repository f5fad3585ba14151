"""Multi-GPU sharding of independent MPC instances: contiguous shards per rank, no data-path
collective, one all-gather of the packed results at the end.  Backend-agnostic (`nccl` on GPUs, `gloo` in CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int):
    """Instances [lo, hi) of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(x: torch.Tensor, u: torch.Tensor, status: torch.Tensor = None) -> torch.Tensor:
    """(B,nx,N), (B,nu,N-1)[, (B,S)] -> (B, nx*N + nu*(N-1) [+ S]) float64, one row per instance."""
    B = x.shape[0]
    parts = [x.reshape(B, -1), u.reshape(B, -1)]
    if status is not None:
        parts.append(status.to(x.dtype).reshape(B, -1))
    return torch.cat(parts, dim=1).contiguous()


def unpack_results(packed: torch.Tensor, nx: int, nu: int, N: int, nstatus: int = 0):
    B = packed.shape[0]
    a, b = nx * N, nx * N + nu * (N - 1)
    x = packed[:, :a].reshape(B, nx, N)
    u = packed[:, a:b].reshape(B, nu, N - 1)
    st = packed[:, b:b + nstatus].round().to(torch.int64) if nstatus else None
    return x, u, st


def all_gather_results(packed: torch.Tensor, total: int) -> torch.Tensor:
    """All ranks receive the `total` result rows in instance order.  Shards may differ in size by one (padded)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return packed
    rank = dist.get_rank()
    sizes = [shard_range(total, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    buf = packed
    if packed.shape[0] < mx:
        buf = torch.cat([packed, packed.new_zeros((mx - packed.shape[0], packed.shape[1]))], dim=0)
    out = packed.new_empty((world * mx, packed.shape[1]))
    dist.all_gather_into_tensor(out, buf.contiguous())
    rows = [out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(rows, dim=0)
