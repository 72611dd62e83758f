"""Host-side plumbing for more than one B200 (one process per GPU, torch.distributed; NCCL on GPUs, gloo in tests).

Two partitionings (the reference itself is single-process, rl/deep_q_agent.py has no multi-device code):

* replicas  -- C4: N independent learners, own replay / sum tree / weights / seed per rank, NO data-path
               collective.  Only the measurement is collective: a barrier, the max-over-ranks time and the sum of
               steps (`aggregate_throughput`).
* data-parallel -- C5 (new capability): every rank samples B/N transitions from its own replay shard and runs the
               forward/backward (`dqn_train_step_phase(0)`), the flat fp32 gradient slab (P floats, one contiguous
               buffer: `dqn_device_ptr("grads")`) is all-reduced and averaged, then every rank applies the same
               optimizer step (`dqn_train_step_phase(1)`), so the weights stay identical on all ranks.  The one
               exchange step of the path is that all-reduce.
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


class DataParallelLearner:
    """Large-batch data-parallel step over `world` learners (global batch = world x per-rank batch).

    `learner` needs train_step_phase(phase, grad_scale) and a gradient tensor (default: the device slab of
    deepqnetwork_b200.learner.Learner); tests drive it with a CPU stand-in over gloo."""

    def __init__(self, learner, group=None, grads=None, stream_sync=None):
        self.learner = learner
        self.group = group
        self._grads = grads if grads is not None else grads_tensor(learner)
        self._sync = stream_sync if stream_sync is not None else getattr(learner, 'sync', lambda: None)

    def train_step(self):
        # phase 0: sample the local shard, forward, backward -> local mean gradient in the slab
        self.learner.train_step_phase(0, 1.0)
        self._sync()  # the learner runs on its own stream; the collective runs on torch's
        allreduce_mean_(self._grads, self.group)  # mean over ranks == gradient of the global-batch mean loss
        import torch
        if self._grads.is_cuda:
            torch.cuda.current_stream().synchronize()
        # phase 1: identical optimizer step on every rank
        self.learner.train_step_phase(1, 1.0)
