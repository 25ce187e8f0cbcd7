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


class ForwardStreams:
    """Several inference forwards in flight: forward i runs on compute stream i % n with its own captured graph and
    static buffers (FusedPWCLONet.forward_graphed(slot=...)).  The sampling chain of a forward is ~1.3 ms of
    latency-bound work on one SM per cloud; with few clouds per GPU (a sharded batch) it leaves most of the machine idle,
    and the layer kernels of the previous batch fill it.  Frame pairs stay independent; results are identical."""

    def __init__(self, net, n=2):
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("ForwardStreams needs the network on a CUDA device (there is no CPU path)")
        self.net, self.device, self.n, self.count = net, dev, max(1, int(n)), 0
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.n)]

    def submit(self, xyz_f1, xyz_f2, after=None, then=None):
        """enqueue one forward on the next stream (ordered after the caller's current stream and after event `after`);
        `then(pose)` runs under that stream right behind it (e.g. the copy of the result to the host).
        Returns (pose, event recorded on the stream once the forward and `then` are complete)."""
        s = self.count % self.n
        self.count += 1
        st = self.streams[s]
        st.wait_stream(torch.cuda.current_stream(self.device))
        if after is not None:
            st.wait_event(after)
        with torch.cuda.stream(st), torch.no_grad():
            self.net.fused_slot = s
            try:
                pose, _ = self.net(xyz_f1, None, xyz_f2, None)
            finally:
                self.net.fused_slot = 0
            if then is not None:
                then(pose)
            done = torch.cuda.Event()
            done.record(st)
        return pose, done

    def join(self):
        """make the caller's current stream wait for everything submitted so far"""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)


def auto_compute_streams(pairs):
    """forwards in flight: the sampling chain of a forward (one SM per cloud, ~1.3 ms whatever the batch) overlaps the layer
    kernels of its neighbours.  Measured on one B200 (profiles/round2_in_flight_sweep.txt), pairs/s with 2 / 3 / 4 / 6 / 8
    in flight: 64 pairs 9 811 / 10 013 / 10 128 / 10 129 / 10 149; 8 pairs 5 175 / 6 198 / 6 664 / 7 495 / 7 507;
    1 pair 1 030 / 1 491 / 1 809 / 2 563 / 2 568.  Six everywhere (each in-flight forward owns a captured graph and its
    static buffers: ~2 GB at 64 pairs).  PWCLO_STREAMS overrides."""
    import os
    if os.environ.get("PWCLO_STREAMS"):
        return max(1, int(os.environ["PWCLO_STREAMS"]))
    return 6


class PosePipeline:
    """Streaming inference for host-resident batches: the host->device copy of batch i+1 (pinned memory, own copy
    stream, second set of device buffers) overlaps the forward of batch i; the [b,4,7] result returns through a
    pinned buffer.  `run(batches)` yields one host pose tensor per batch, in order; a yielded tensor is valid until
    `depth` further batches have been submitted.  Several forwards (auto_compute_streams: six) are in flight on their own
    compute streams (ForwardStreams).  Frame pairs stay independent: this is per-rank plumbing, the sharding across GPUs is
    unchanged."""

    def __init__(self, net, pairs, n_points, depth=None, compute_streams=None):
        """compute_streams: forwards in flight (default: auto_compute_streams(pairs)); depth: host / device buffer
        sets (default: compute_streams + 1)"""
        compute_streams = auto_compute_streams(pairs) if compute_streams is None else int(compute_streams)
        depth = compute_streams + 1 if depth is None else max(int(depth), compute_streams + 1)
        self.net, self.depth = net, depth
        self.fwd = ForwardStreams(net, compute_streams)
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
            def finish(pose, s=s):       # under the forward's compute stream: inputs consumed, result to the host
                self.freed[s].record(torch.cuda.current_stream(self.device))
                self.pose_host[s].copy_(pose, non_blocking=True)

            _, self.done[s] = self.fwd.submit(self.buf[s][0], self.buf[s][1], after=self.copied[s], then=finish)
            pending.append(s)
            if len(pending) == self.depth:          # keep depth-1 batches queued behind the one we wait for
                s0 = pending.pop(0)
                self.done[s0].synchronize()
                yield self.pose_host[s0]
        for s0 in pending:
            self.done[s0].synchronize()
            yield self.pose_host[s0]
