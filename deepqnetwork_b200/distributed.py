"""Host-side plumbing for more than one B200 (one process per GPU, torch.distributed; NCCL on GPUs, gloo in tests).

Two partitionings (the reference itself is single-process, rl/deep_q_agent.py has no multi-device code):

* replicas  -- C4: N independent learners, own replay / sum tree / weights / seed per rank, NO data-path
               collective.  Only the measurement is collective: a barrier, the max-over-ranks time and the sum of
               steps (`aggregate_throughput`).
* data-parallel -- C5 (new capability): every rank samples B/N transitions from its own replay shard and runs the
               forward/backward (`dqn_train_step_phase(0)`), the flat fp32 gradient slab (P floats, one contiguous
               buffer: `dqn_device_ptr("grads")`) is summed over the ranks (each rank scaled its loss by 1/N, so the sum
               is the gradient of the global-batch mean), then every rank applies the same optimizer step
               (`dqn_train_step_phase(1)`), so the weights stay identical on all ranks.  The one exchange step of the
               path is that all-reduce; it is issued in two buckets on a communication stream: the two dense hidden
               kernels (98.6 % of the bytes, complete first in the backward pass: `dqn_phase_wait("fc_grads")`) go
               while the conv backward chain is still running, the remaining 0.4 MB follow when the slab is complete.
"""
import numpy as np


class _DeviceArray:
    """Zero-copy view of a device buffer of the learner through __cuda_array_interface__ (consumed by torch)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = dict(shape=(int(n_floats),), typestr='<f4', data=(int(ptr), False), version=2)


def grads_tensor(learner):
    """The learner's gradient slab as a torch CUDA tensor (no copy).  The slab is padded: every parameter tensor
    starts 16-byte aligned; padding elements are zero and stay zero under sum/mean."""
    import torch
    ptr, nbytes = learner.device_ptr('grads')
    return torch.as_tensor(_DeviceArray(ptr, nbytes // 4), device='cuda')


def allreduce_mean_(t, group=None):
    """In-place mean over the ranks of `group` (sum all-reduce, then 1/world): the DP exchange step."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.mul_(1.0 / world)
    return t


def aggregate_throughput(steps_local, ms_local, device=None, group=None):
    """Whole-job throughput of independent replicas: (sum of steps over ranks) / (max over ranks of the device time).
    Returns (steps_per_s, ms_max, steps_total) on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return steps_local / (ms_local / 1000.0), ms_local, steps_local
    tmax = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    tsum = torch.tensor([float(steps_local)], dtype=torch.float64, device=device)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM, group=group)
    ms, steps = float(tmax.item()), float(tsum.item())
    return steps / (ms / 1000.0), ms, int(round(steps))


def shard_capacity(total_transitions, rank, world):
    """Replay rows owned by `rank` when a replay of `total_transitions` is split over `world` shards (remainder to the
    lowest ranks); the shards are disjoint and cover the whole range."""
    base, rem = divmod(int(total_transitions), int(world))
    n = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, n


def gradient_buckets(tensors):
    """Split the gradient slab into the exchange buckets.  tensors: [(name, count, offset)] (Learner.tensors()).
    Returns (big, small): big = [(offset, count)] of the dense hidden kernels ('dense/kernel', 'dense_2/kernel'), each one
    contiguous range; small = [(offset, count)] of everything else, coalesced into maximal contiguous ranges (tensor
    starts are 16-byte aligned, so the ranges include the zero padding between tensors)."""
    big = [(off, cnt) for name, cnt, off in tensors if name in ('dense/kernel', 'dense_2/kernel')]
    end = max(off + (cnt + 3) // 4 * 4 for _, cnt, off in tensors)
    small, cur = [], 0
    for off, cnt in sorted(big):
        if off > cur:
            small.append((cur, off - cur))
        cur = off + (cnt + 3) // 4 * 4
    if end > cur:
        small.append((cur, end - cur))
    return big, small


class DataParallelLearner:
    """Large-batch data-parallel step over `world` learners (global batch = world x per-rank batch).

    `learner` needs train_step_phase(phase, grad_scale) and a gradient tensor (default: the device slab of
    deepqnetwork_b200.learner.Learner); tests drive it with a CPU stand-in over gloo.

    overlap=True (device learners): the learner must run on a torch stream (Learner.set_stream(torch stream)); the
    collectives are issued on a second torch stream that waits for events recorded INSIDE phase 0, so the 31.5 MB
    dense-kernel bucket is on the wire while the conv backward chain still runs, and phase 1 waits for both buckets on
    the device (no host synchronisation anywhere in the step)."""

    def __init__(self, learner, group=None, grads=None, stream_sync=None, overlap=False, learner_stream=None):
        import torch.distributed as dist
        self.learner = learner
        self.group = group
        self._grads = grads if grads is not None else grads_tensor(learner)
        self._sync = stream_sync if stream_sync is not None else getattr(learner, 'sync', lambda: None)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.overlap = bool(overlap) and self.world > 1
        self.steps = 0
        self._per = bool(getattr(learner, 'use_per', False)) and hasattr(learner, 'set_per_normaliser')
        if self._per:
            learner.set_option('external_normaliser', 1)
        if self.overlap:
            import torch
            self._stream = learner_stream
            assert learner_stream is not None, 'overlap needs the torch stream the learner runs on'
            self._comm = torch.cuda.Stream()
            big, small = gradient_buckets(learner.tensors())
            self._big = [self._grads[o:o + c] for o, c in big]
            self._small = [self._grads[o:o + c] for o, c in small]
            self._small_flat = torch.empty(sum(c for _, c in small), dtype=torch.float32, device=self._grads.device)

    def sync_per_normaliser(self):
        """One scale for the importance weights of all shards (deep_q_agent.py:506-510 evaluated per shard, then the
        maximum over the ranks: the shard that holds the globally smallest sampling probability sets the scale).  The
        learner's own refresh is switched off (option external_normaliser); this runs at the reference's cadence."""
        import torch
        import torch.distributed as dist
        from .deep_q_agent import max_weight
        L = self.learner
        mn, mx, avg, total = L.tree_stats()
        per_b = L.cfg.per_b_start
        if L.cfg.use_per_annealing:
            per_b = min(L.cfg.per_b_end, per_b + self.steps * ((L.cfg.per_b_end - L.cfg.per_b_start) / L.cfg.per_anneal_steps))
        mw = float(max_weight(float(mn), float(total), L.batch_size, per_b))
        if self.world > 1:
            t = torch.tensor([mw], dtype=torch.float64, device=self._grads.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            mw = float(t.item())
        L.set_per_normaliser(mw)
        return mw

    def train_step(self):
        import torch
        import torch.distributed as dist
        scale = 1.0 / self.world
        if self._per and self.steps % max(int(self.learner.cfg.per_calculate_steps), 1) == 0:
            self.sync_per_normaliser()
        if not self.overlap:
            # phase 0: sample the local shard, forward, backward -> local gradient (already scaled by 1/world) in the slab
            self.learner.train_step_phase(0, scale)
            if self.world > 1:
                self._sync()  # the learner runs on its own stream; the collective runs on torch's
                dist.all_reduce(self._grads, op=dist.ReduceOp.SUM, group=self.group)
                if self._grads.is_cuda:
                    torch.cuda.current_stream().synchronize()
            self.learner.train_step_phase(1, scale)  # identical optimizer step on every rank
            self.steps += 1
            return
        L = self.learner
        L.train_step_phase(0, scale)
        works = []
        with torch.cuda.stream(self._comm):
            L.phase_wait('fc_grads', self._comm.cuda_stream)      # the two dense kernels: ready early in the backward pass
            for t in self._big:
                works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            L.phase_wait('grads', self._comm.cuda_stream)          # conv / heads / biases: ready when phase 0 ends
            torch.cat(self._small, out=self._small_flat)
            works.append(dist.all_reduce(self._small_flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        with torch.cuda.stream(self._stream):
            for w in works:
                w.wait()                                           # the learner's stream waits on the device
            off = 0
            for t in self._small:
                t.copy_(self._small_flat[off:off + t.numel()]); off += t.numel()
            L.train_step_phase(1, scale)
        self.steps += 1
