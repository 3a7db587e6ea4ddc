"""Multi-GPU plumbing for the head: images are independent (no cross-image term anywhere in
nets_factory.py:242-352, and no batch-norm in the head: normalizer_fn=None :268,:303), so the
batch is scattered contiguously over ranks — keeping each video's F frames/crops on one rank
so the frame mean (nets_factory.py:354-374) stays local — and the only collective is ONE
all-gather of the [B/G, K] logits.  The reference's analogue is in-graph towers
(models/slim/deployment/model_deploy.py:145 create_clones); it has no collective library.

One process per GPU, `torch.distributed` (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations


def shard_bounds(num_videos, rank, world):
    """Contiguous, balanced [lo, hi) of videos for `rank` (first `num_videos % world` ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(int(num_videos), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_features(X, frames, rank, world):
    """X: [B*F, ...] -> this rank's rows (whole videos only)."""
    if X.shape[0] % frames:
        raise ValueError("batch %d is not a multiple of frames %d" % (X.shape[0], frames))
    lo, hi = shard_bounds(X.shape[0] // frames, rank, world)
    return X[lo * frames:hi * frames]


def all_gather_logits(local, out=None, group=None):
    """One all-gather of equally sized [b, K] blocks into [world*b, K] (rank-major == batch order)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def all_gather_logits_uneven(local, num_videos, group=None):
    """All-gather when num_videos % world != 0: pad every block to the largest shard, gather once, trim."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(num_videos, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if local.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d videos, expected %d" % (rank, local.shape[0], hi - lo))
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:hi - lo] = local
    full = all_gather_logits(pad, group=group).reshape((world, mx) + tuple(local.shape[1:]))
    return torch.cat([full[r, :sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


class ShardedHead(object):
    """network_fn-like callable for a batch that is sharded over the ranks of `group`.

    `head(X_local) -> logits_local` is any per-rank head (the CUDA AttnPoolOp on the GPU box).
    """

    def __init__(self, head, num_videos, group=None):
        self.head, self.num_videos, self.group = head, int(num_videos), group

    def __call__(self, X_local):
        import torch.distributed as dist
        local = self.head(X_local)
        world = dist.get_world_size(self.group)
        if self.num_videos % world == 0:
            return all_gather_logits(local, group=self.group)
        return all_gather_logits_uneven(local, self.num_videos, group=self.group)


class PeerGather(object):
    """The all-gather of the logits fused into the kernel: a symmetric ring of `depth` gathered buffers
    [depth, world * b, K] on every rank; each rank's kernel stores its block of call `step` into buffer
    step % depth on every peer over NVLink and signals flags (include/apa.h: ApaPeerGather).  Built on
    torch.distributed symmetric memory for the peer mappings only.  `world == 1` (no process group) degenerates
    to a local ring, which is how single-GPU tests cover it.

        step = op.forward_gather(X, pg)      # with ApaTuning.gather_wait_lag = L the buffer of step - L is
        pg.buffer(step - L)                  #   complete in stream order after the call
        pg.wait(0); pg.buffer(step)          # or wait explicitly for the newest one
    The step is counted on the device (CUDA-graph replays keep counting); `pg.step` mirrors it on the host for
    eager calls, `pg.sync_step()` re-reads it after graph replays.
    """

    def __init__(self, videos_per_rank, num_classes, device, group=None, depth=1):
        import torch
        from .custom_ops import attn_pool_factory as ops
        self.rank, self.world = 0, 1
        self.depth = int(depth)
        if not (1 <= self.depth <= 4):
            raise ValueError("depth must be in 1..4")
        self.step = 0
        self.rows = int(videos_per_rank)
        self._keep = []
        use_dist = False
        try:
            import torch.distributed as dist
            use_dist = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        except Exception:
            use_dist = False
        g = ops.ApaPeerGather()
        if use_dist:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            self.rank, self.world = dist.get_rank(grp), dist.get_world_size(grp)
            if self.world > ops.APA_MAX_PEERS:
                raise ValueError("at most %d ranks" % ops.APA_MAX_PEERS)
            self.ring = symm_mem.empty((self.depth, self.world * videos_per_rank, num_classes), dtype=torch.float32,
                                       device=device)
            self.flags = symm_mem.empty((64,), dtype=torch.int64, device=device)
            self.ring.zero_()
            self.flags.zero_()
            torch.cuda.synchronize(device)
            h_out = symm_mem.rendezvous(self.ring, grp)
            h_flags = symm_mem.rendezvous(self.flags, grp)
            self._keep += [h_out, h_flags]
            for q in range(self.world):
                g.peer_logits[q] = int(h_out.buffer_ptrs[q])
                g.peer_flags[q] = int(h_flags.buffer_ptrs[q])
            dist.barrier(grp)
        else:
            self.ring = torch.zeros((self.depth, videos_per_rank, num_classes), dtype=torch.float32, device=device)
            self.flags = torch.zeros((64,), dtype=torch.int64, device=device)
            g.peer_logits[0] = self.ring.data_ptr()
            g.peer_flags[0] = self.flags.data_ptr()
        self.done = torch.zeros((8,), dtype=torch.int64, device=device)
        g.world, g.rank, g.depth = self.world, self.rank, self.depth
        g.counters = self.done.data_ptr()
        self.struct = g
        self.device = device

    def next_step(self):
        self.step += 1
        return self.step

    def sync_step(self):
        """Re-read the device-side count of completed calls (after CUDA-graph replays); synchronises."""
        self.step = int(self.done[1].item())
        return self.step

    def buffer(self, step):
        """The gathered [world * b, K] logits of call `step` (valid once that step is complete, see class doc)."""
        return self.ring[int(step) % self.depth]

    @property
    def out(self):
        """The gathered logits of the most recent call."""
        return self.buffer(self.step)

    def wait(self, lag=0, stream=None):
        """Enqueue a wait until the gathered buffer of call (latest - lag) is complete on this rank."""
        import torch
        from .custom_ops import attn_pool_factory as ops
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        ops.peer_gather_wait(self.struct, lag, stream)
