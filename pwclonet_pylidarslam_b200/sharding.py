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


class PosePipeline:
    """Streaming inference for host-resident batches: the host->device copy of batch i+1 (pinned memory, own copy
    stream, second set of device buffers) overlaps the forward of batch i; the [b,4,7] result returns through a
    pinned buffer.  `run(batches)` yields one host pose tensor per batch, in order; a yielded tensor is valid until
    `depth` further batches have been submitted.  Frame pairs stay independent: this is per-rank plumbing, the
    sharding across GPUs is unchanged."""

    def __init__(self, net, pairs, n_points, depth=2):
        self.net, self.depth = net, depth
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PosePipeline needs the network on a CUDA device (there is no CPU path)")
        self.device = dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.buf = [(torch.empty(pairs, 3, n_points, device=dev), torch.empty(pairs, 3, n_points, device=dev))
                    for _ in range(depth)]
        self.pose_host = [torch.empty(pairs, 4, 7).pin_memory() for _ in range(depth)]
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.freed = [torch.cuda.Event() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]

    def run(self, batches):
        """batches: iterable of (xyz_f1, xyz_f2) host tensors [pairs,3,N] (pinned for an asynchronous copy)"""
        main = torch.cuda.current_stream(self.device)
        pending = []
        for i, (h1, h2) in enumerate(batches):
            s = i % self.depth
            with torch.cuda.stream(self.copy_stream):
                if i >= self.depth:
                    self.copy_stream.wait_event(self.freed[s])      # the forward that read these buffers has finished
                self.buf[s][0].copy_(h1, non_blocking=True)
                self.buf[s][1].copy_(h2, non_blocking=True)
                self.copied[s].record(self.copy_stream)
            main.wait_event(self.copied[s])
            with torch.no_grad():
                pose, _ = self.net(self.buf[s][0], None, self.buf[s][1], None)
            self.freed[s].record(main)
            self.pose_host[s].copy_(pose, non_blocking=True)
            self.done[s].record(main)
            pending.append(s)
            if len(pending) == self.depth:          # keep depth-1 batches queued behind the one we wait for
                s0 = pending.pop(0)
                self.done[s0].synchronize()
                yield self.pose_host[s0]
        for s0 in pending:
            self.done[s0].synchronize()
            yield self.pose_host[s0]
