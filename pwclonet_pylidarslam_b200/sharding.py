"""Frame-pair sharding across GPUs.  Frame pairs are independent in inference (BatchNorm in eval
mode, every softmax / pool is per sample), so the path shards with NO data-path collective: rank r
of W runs the fused forward on a contiguous slice of the pairs.  The only collectives are
bookkeeping: gathering the [b,4,7] poses and the max-over-ranks of the measured time."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """contiguous, balanced slice [lo, hi) of `total` frame pairs for `rank`"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_poses(local_pose, total):
    """local_pose [b,4,7] on every rank -> [total,4,7] on every rank (ranks ordered by slice)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_pose
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    bufs = [torch.empty((hi - lo,) + tuple(local_pose.shape[1:]), dtype=local_pose.dtype, device=local_pose.device)
            for lo, hi in sizes]
    dist.all_gather(bufs, local_pose.contiguous())
    return torch.cat(bufs, dim=0)


def max_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def forward_sharded(net, xyz_f1, xyz_f2):
    """Run `net` (PWCLONet in eval mode on this rank's GPU) on this rank's slice of a global batch of
    host tensors [T,3,N]; returns the gathered [T,4,7] poses."""
    total = xyz_f1.shape[0]
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(total, rank, world)
    dev = next(net.parameters()).device
    with torch.no_grad():
        pose, _ = net(xyz_f1[lo:hi].to(dev, non_blocking=True), None, xyz_f2[lo:hi].to(dev, non_blocking=True), None)
    return gather_poses(pose, total)
