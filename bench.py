#!/usr/bin/env python
"""Throughput benchmark: DQN train steps/sec (batch 32, PER, reference Q-network) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c1|c2|c3|nature|c5] [--capacity R]

A *step* is one DeepQAgent._train_run_train equivalent (rl/deep_q_agent.py:326-349): PER sum-tree sample ->
replay gather -> target/online forward -> TD loss -> backward -> Adam -> priority write-back, with the
target sync every copy_network_steps and the PER stats refresh every per_calculate_steps included.

N=1 workload = BASELINE.json configs[1] (C2): Breakout double+dueling+PER, 89x80x4 uint8, batch 32, 1M-transition
replay resident in HBM, synthetic data.  N>1 (torchrun, one rank per GPU) = configs[3] (C4):
N independent learners, no data-path collective, value = total steps of all ranks / max-over-ranks time.
`--config c5` under torchrun = configs[4] (C5): ONE learner with a global batch of 2048 split over the ranks, replay and
sum tree sharded, the 31.9 MB gradient all-reduced with NCCL in two buckets overlapped with the backward pass
(deepqnetwork_b200/distributed.py); value = global-batch steps/s, "scaling": "strong".

Printed JSON (one line, rank 0): the driver contract keys plus
  roofline     -- the dominant kernel (largest CUDA-event time in a per-kernel profile pass of the same step),
                  algorithmic bytes per launch (DESIGN.md "kernels" table) / its average launch time, vs the measured
                  HBM peak of MEASURED_PEAKS.json;
  cpu_baseline -- the oracle's CPU restatement of the reference step (NumPy/Python replay+PER exactly as the
                  reference structures it, torch-CPU fp32 network) timed on this box's host cores;
  e2e          -- the step as the drop-in DeepQAgent runs it (deepqnetwork_b200/deep_q_agent.py:_train_run_train): one
                  dqn_train_step call per step fed the step's host inputs -- the B random.random() draws, pinned -- and
                  returning step / per-sample losses / loss to the host; the replay stays where the design keeps it, in
                  HBM (it is appended as frames arrive, not per train step).  `e2e.host_batch` is the other integration:
                  seams 3 + 2 + 3 with the sampled rows gathered on the HOST from a host replay and uploaded every step.
`--impl reference` times the CPU restatement of the reference alone (TensorFlow cannot be installed here; see
DESIGN.md) and prints the same line shape with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    'c1': dict(name='C1 Breakout vanilla DQN', h=89, w=80, a=4, dueling=False, double=False, per=False, variant='reference'),
    'c2': dict(name='C2 Breakout double+dueling+PER', h=89, w=80, a=4, dueling=True, double=True, per=True, variant='reference'),
    'c3': dict(name='C3 MsPacman double+dueling+PER', h=86, w=80, a=9, dueling=True, double=True, per=True, variant='reference'),
    'nature': dict(name='canonical Nature-84 vanilla (labelled extra)', h=84, w=84, a=4, dueling=False, double=False, per=False, variant='nature'),
    'c5': dict(name='C5 large-batch data-parallel DQN (global batch 2048), Breakout double+dueling+PER', h=89, w=80, a=4, dueling=True,
               double=True, per=True, variant='reference', global_batch=2048),
}
METRIC = 'DQN train steps/sec (batch 32, PER, Nature CNN) per B200 and at 1/2/4/8 GPUs'
UNIT = 'train steps/s'
BATCH = 32


# ---------------------------------------------------------------------------------------------------------
# algorithmic work: P params, F fwd MACs/sample, per-kernel bytes
def net_numbers(cfg):
    from deepqnetwork_b200.deep_q_model import conv_geometry, tower_param_shapes
    geo = conv_geometry(cfg['h'], cfg['w'], 4, cfg['variant'])
    shapes = tower_param_shapes(cfg['h'], cfg['w'], 4, cfg['a'], 512, cfg['dueling'], cfg['variant'])
    P = sum(int(np.prod(s)) for s in shapes.values())
    conv_macs = [g['oh'] * g['ow'] * g['cout'] * g['k'] * g['k'] * g['cin'] for g in geo]
    flat = geo[-1]['oh'] * geo[-1]['ow'] * geo[-1]['cout']
    streams = 2 if cfg['dueling'] else 1
    fc_macs = flat * 512 * streams + 512 * cfg['a'] + (512 if cfg['dueling'] else 0)
    F = sum(conv_macs) + fc_macs
    n_fwd = 3 if cfg['double'] else 2
    flops = 2 * BATCH * (n_fwd * F + F + (F - conv_macs[0]))
    S = cfg['h'] * cfg['w'] * 4
    step_bytes = 8 * 4 * P + 2 * BATCH * S + 15 * BATCH
    return dict(P=P, F=F, F1=conv_macs[0], flops=flops, step_bytes=step_bytes, S=S, flat=flat, geo=geo, streams=streams, n_fwd=n_fwd)


def kernel_bytes(name, cfg, nn):
    """Algorithmic HBM bytes of one launch of the named profile segment (unique reads + writes, fp32)."""
    geo, flat, st, B = nn['geo'], nn['flat'], nn['streams'], BATCH
    rows_on = 2 * B if cfg['double'] else B
    base = name[3:] if name[:3] in ('on_', 'tg_') else name
    rows = rows_on if name.startswith('on_') else B
    if base == name and base.startswith('conv') and base.endswith('_fwd'):
        rows = rows_on + B  # both towers in one launch
    acts = [g['oh'] * g['ow'] * g['cout'] for g in geo]
    wts = [g['k'] * g['k'] * g['cin'] * g['cout'] for g in geo]
    ins = [nn['S'] // 4 * 1, acts[0], acts[1]]  # elements per sample feeding conv l (conv1 input counted as bytes below)
    if base == 'conv_tower_fwd':   # fused conv tower, both towers: frames in (uint8), fp32 weights of both towers once, the
        # training rows' activations out (the backward pass reads them) + conv3's output of every row (the dense layer's operand)
        return (rows_on + B) * nn['S'] + 2 * sum(wts) * 4 + B * (acts[0] + acts[1]) * 4 + (rows_on + B) * acts[2] * 4
    if base.startswith('conv') and base.endswith('_fwd'):
        l = int(base[4]) - 1
        in_b = rows * nn['S'] if l == 0 else rows * acts[l - 1] * 4
        return in_b + wts[l] * 4 + rows * acts[l] * 4
    if base == 'fc_fwd':
        return rows * flat * 4 + flat * 512 * st * 4 + rows * 512 * st * 4
    if base.endswith('_dgrad_reduce') and base.startswith('conv'):
        l = int(base[4]) - 1
        return 4 * B * acts[l - 1] * 4
    if base == 'fc_dgrad_reduce':
        return 2 * B * flat * 4 * 2 + B * flat * 4
    if base == 'fc_reduce_heads':
        return rows * 512 * st * 4 * 2
    per = nn.get('split_fc')       # dense wgrad / optimizer launched once per hidden stream ('...1' = the value stream)
    if base in ('fc_wgrad', 'fc_wgrad1'):
        n_st = 1 if per else st
        return B * flat * 4 + B * 512 * n_st * 4 + flat * 512 * n_st * 4
    if base == 'fc_dgrad':
        return B * 512 * st * 4 + flat * 512 * st * 4 + 2 * B * flat * 4
    if base.endswith('_wgrad'):
        l = int(base[4]) - 1
        in_b = B * nn['S'] if l == 0 else B * acts[l - 1] * 4
        return in_b + B * acts[l] * 4 + wts[l] * 4
    if base.endswith('_wgrad_reduce'):
        l = int(base[4]) - 1
        return 2 * wts[l] * 4
    if base.endswith('_dgrad'):
        l = int(base[4]) - 1
        return B * acts[l] * 4 + wts[l] * 4 + 2 * B * acts[l - 1] * 4
    if base == 'fc_wgrad_adam':    # read w,m,v + write w,m,v of the dense kernels; activations / dH tiles come from L2
        return 6 * 4 * flat * 512 * st + B * flat * 4 + B * 512 * st * 4
    if base in ('optimizer_fc', 'optimizer_fc1'):  # dense hidden kernels only (98.6% of P), issued inside the backward pass
        return 7 * 4 * flat * 512 * (1 if per else st)
    if base == 'optimizer':        # read g,p,m,v ; write p,m,v over whatever optimizer_fc has not already updated
        return 7 * 4 * (nn['P'] - (flat * 512 * st if nn.get('early_fc') else 0))
    if base == 'gather':
        return 2 * 2 * B * nn['S']
    return None


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML every few milliseconds
    (a 100 ms nvidia-smi poll cannot see a region of a few hundred milliseconds)."""

    def __init__(self, device_index, period_s=0.004):
        import threading
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[device_index]) if vis and all(x.strip().isdigit() for x in vis.split(',')) else device_index
            self._nv, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None
            return
        self._period = period_s
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv, h = self._nv, self._h
        names = (('hw_slowdown', nv.nvmlClocksEventReasonHwSlowdown), ('hw_thermal_slowdown', nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ('sw_thermal_slowdown', nv.nvmlClocksEventReasonSwThermalSlowdown), ('sw_power_cap', nv.nvmlClocksEventReasonSwPowerCap),
                 ('hw_power_brake', nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for name, bit in names:
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(self._period)

    def start(self):
        if self._thread:
            self._thread.start()
        return self

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=[])
        if self._thread:
            self._stop.set(); self._thread.join(timeout=2)
        if self.samples:
            out['sm_mhz'] = statistics.median(self.samples); out['sm_min_mhz'] = min(self.samples); out['samples'] = len(self.samples)
            out['power_w_max'] = max(self.power) if self.power else None
        out['reasons'] = sorted(self.reasons)
        out['how'] = 'in-process NVML, %.0f ms period, over the timed region' % (1000 * getattr(self, '_period', 0))
        return out


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind a profile name, from the ncu --set full
# capture committed as profiles/r01c_top_kernels_raw.csv (adam_kernel: 62.9 MB read + 1.0-2.4 MB written per launch;
# the stores are still dirty in L2 when the kernel ends)
NCU_TRAFFIC = {'optimizer_fc': 62.92e6 + 0.92e6, 'optimizer_fc1': 62.92e6 + 2.35e6,
               # fcwg_adam_rows_kernel (profiles/r02_top_kernels_raw.csv): 95.55 MB read + 37.71 MB written per launch (C2 shapes)
               'fc_wgrad_adam': 95.55e6 + 37.71e6,
               # cvtower_kernel, 96 images (same capture): 3.54 MB read, stores still in L2
               'conv_tower_fwd': 3.54e6}
# dominant kernels that are not bandwidth-bound: what bounds them instead
KERNEL_NOTES = {'conv_tower_fwd': 'tensor-issue / shared-memory-bandwidth / latency bound (one CTA per image through three conv layers): '
                                  'its few MB of algorithmic bytes make the HBM fraction meaningless; in-kernel timeline in profiles/README.md'}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, steps, warmup, replay_rows, time_budget_s, threads=None, BATCH=BATCH):
    """The reference's CPU path, restated (oracle/): NumPy/Python replay ring + sum tree structured exactly like
    rl/data/*.py, torch-CPU fp32 network + TF-formula Adam like rl/deep_q_model.py, driven like
    DeepQAgent._train_run_train.  Returns (steps_per_s, steps_timed, threads)."""
    import random
    import torch
    from oracle import nn_ref
    from oracle.sumtree_ref import ReplayRingRef, SumTreeRef, per_sample, per_update, uniform_rows
    from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    spec = nn_ref.NetSpec(cfg['h'], cfg['w'], 4, cfg['a'], use_dueling=cfg['dueling'], use_double=cfg['double'],
                          use_per=cfg['per'], variant=cfg['variant'])
    shapes = tower_param_shapes(cfg['h'], cfg['w'], 4, cfg['a'], 512, cfg['dueling'], cfg['variant'])
    o = nn_ref.to_torch(init_tower_params(shapes, 42), torch.float32)
    t = nn_ref.to_torch(init_tower_params(shapes, 43), torch.float32)
    opt = nn_ref.OptState(o, spec, torch.float32)
    rng = np.random.default_rng(1234)
    ring = ReplayRingRef(cfg['h'], cfg['w'], 4, replay_rows)
    ring.states[:] = rng.integers(0, 256, ring.states.shape, dtype=np.uint8)
    ring.next_states[:] = rng.integers(0, 256, ring.states.shape, dtype=np.uint8)
    ring.actions[:] = rng.integers(0, cfg['a'], replay_rows); ring.rewards[:] = rng.choice([0, 1, 4, 7], replay_rows, p=[.9, .06, .03, .01])
    ring.continues[:] = rng.random(replay_rows) < 0.99
    ring.cur_idx, ring.is_full = 0, True
    tree = SumTreeRef(replay_rows)
    pr = nn_ref.make_priority(np.abs(rng.normal(0, 0.3, replay_rows)).astype(np.float32))
    if cfg['per']:
        # first fill: leaves + bottom-up sums (timing only; the bit-exact add path is exercised in tests)
        tree.tree[replay_rows - 1:] = pr
        for i in range(replay_rows - 2, -1, -1):
            tree.tree[i] = tree.tree[2 * i + 1] + tree.tree[2 * i + 2]
        tree.data[:] = np.arange(replay_rows); tree.size = replay_rows
    random.seed(0)
    per_b, max_w = 0.4, None

    def one_step():
        nonlocal max_w
        if cfg['per']:
            if max_w is None:
                max_w = nn_ref.max_weight(tree.get_min(), tree.total, BATCH, per_b)
            u = [random.random() for _ in range(BATCH)]
            ti, di, p, mi = per_sample(tree, BATCH, u)
            isw = nn_ref.is_weights(p, tree.total, BATCH, per_b, max_w)
        else:
            mi = uniform_rows(replay_rows, BATCH, random); isw = None
        batch = ring.gather(mi)
        step, losses, loss, _ = nn_ref.train_step(o, t, opt, batch, isw, spec, torch.float32)
        if cfg['per']:
            per_update(tree, ti, nn_ref.make_priority(losses), ring.losses)
        return loss

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        one_step()
        done += 1
        if time.perf_counter() - t0 > time_budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, threads


# ---------------------------------------------------------------------------------------------------------
# Only the JSON line may reach stdout: fd 1 is pointed at stderr for the whole run (NCCL prints its version banner to
# stdout, subprocesses inherit fd 1) and the line is written to the saved descriptor.
_JSON_FD = None


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=list(CONFIGS))
    ap.add_argument('--capacity', type=int, default=None, help='replay transitions resident in HBM (default 1M; c3: 10M)')
    ap.add_argument('--layout', default=None, choices=['stacked', 'frames'],
                    help="replay layout: the reference's stacked s/s' rows, or frame-de-duplicated (default for c3: 10M transitions = 69 GB)")
    ap.add_argument('--math', default='tc3x', choices=['fp32', 'tc3x', 'tc1x'],
                    help='fp32: FFMA; tc3x: tcgen05 TF32 3-term split (fp32-grade, default); tc1x: single-pass TF32')
    ap.add_argument('--no-fuse', action='store_true', help='ablation: optimizer as a separate pass over the gradient slab')
    ap.add_argument('--fuse', action='store_true', help='ablation: Adam for the dense kernels in the dense wgrad epilogue')
    ap.add_argument('--ffma-adam', action='store_true', help='ablation: dense wgrad + optimizer as one fp32 FFMA pass')
    ap.add_argument('--no-pdl', action='store_true', help='ablation: no programmatic dependent launch between consecutive kernels')
    ap.add_argument('--no-overlap', action='store_true', help='ablation: single stream, no forked branches in the step graph')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cfg = CONFIGS[args.config]
    nn = net_numbers(cfg)
    if args.capacity is None:
        args.capacity = 10000000 if args.config == 'c3' else 1000000
    if args.layout is None:
        args.layout = 'frames' if args.config == 'c3' else 'stacked'
    replay_gb = args.capacity * ((nn['S'] // 4 + 32) if args.layout == 'frames' else 2 * nn['S']) / 1e9
    workload = '%s, %dx%dx4 uint8, batch %d, %d-transition replay in HBM (%s layout, %.1f GB)' % (
        cfg['name'], cfg['h'], cfg['w'], BATCH, args.capacity, "frame-de-duplicated" if args.layout == 'frames' else "stacked s/s'", replay_gb)
    if world > 1:
        workload = 'C4 %d independent learners (one per GPU, no collectives), each: ' % world + workload
    config = dict(workload=workload, batch=BATCH, replay_transitions=args.capacity, parallelism='replicas x%d' % world,
                  params=nn['P'], gflop_per_step=round(nn['flops'] / 1e9, 3), algorithmic_mb_per_step=round(nn['step_bytes'] / 1e6, 1),
                  l2='no flush: steady-state step; per-step working set (5 parameter-sized slabs = %.0f MB + random replay rows '
                     'out of %.1f GB) exceeds the 126 MB L2' % (5 * 4 * nn['P'] / 1e6, replay_gb),
                  math={'fp32': 'fp32 FFMA (DQN_MATH_FP32)', 'tc3x': 'tcgen05 kind::tf32, 3-term split, fp32 accumulate in TMEM (DQN_MATH_TC3X; parity-tested at 1e-4 like fp32)',
                        'tc1x': 'tcgen05 kind::tf32 single pass (DQN_MATH_TC1X; NOT fp32-grade, ~1e-3)'}[args.math])

    # ---- reference arm: the CPU restatement of the reference, rank 0 only -------------------------------------
    if args.impl == 'reference':
        if rank != 0:
            return
        gb = cfg.get('global_batch', BATCH)
        v, done, threads = cpu_reference_run(cfg, args.steps, min(args.warmup, 5) if gb == BATCH else 1, 20000, 150.0, BATCH=gb)
        if gb != BATCH:
            config.update(batch=gb, workload='%s, %dx%dx4 uint8, global batch %d on the host cores, 20000-row replay' % (cfg['name'], cfg['h'], cfg['w'], gb))
        line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=done, warmup=min(args.warmup, 5),
                    ms_per_step=1000.0 / v, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                    data='synthetic', impl='reference', config=config, gpu_launches=0,
                    cpu_baseline=dict(value=v, unit=UNIT, cores=threads, kind='port',
                                      sample='%d consecutive train steps, 20000-row host replay (sampler cost is O(log N))' % done),
                    e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        emit(line)
        return

    # ---- B200 arm ---------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: deepqnetwork_b200 has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if rank == 0:
        from deepqnetwork_b200 import build as lib_build  # the CUDA library only: this arm never touches oracle/
        lib_build.build()
    if world > 1:
        dist.barrier()
    from deepqnetwork_b200.learner import Learner
    from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes

    dp = dp_mode = args.config == 'c5'
    batch = cfg['global_batch'] // world if dp else BATCH
    capacity = args.capacity // world if dp else args.capacity
    L = Learner(cfg['h'], cfg['w'], 4, num_actions=cfg['a'], use_dueling=cfg['dueling'], use_double=cfg['double'],
                use_per=cfg['per'], batch_size=batch, replay_capacity=capacity, conv_variant=cfg['variant'],
                math_mode=args.math, seed=7 + rank, device=local_rank, max_rows=max(batch, 256), frame_dedup=args.layout == 'frames')
    # a real (non-default) stream: stream capture is illegal on the legacy stream; high priority, because the learner's
    # own side streams (weight gradients, optimizer, look-ahead sampling) only carry work with slack
    # Replicas are timed on the learner's OWN stream (the configuration a user of the handle gets), wrapped for torch so that
    # torch.cuda.Event brackets exactly the learner's work.  Measured at the end of round 2: handing the learner a torch pool
    # stream (torch.cuda.Stream(priority=-1) + dqn_set_stream) costs 5 % of the step rate (6630 vs 6960 steps/s, same build).
    stream = None
    if not dp_mode:
        try:
            stream = torch.cuda.ExternalStream(int(L.device_ptr('stream')[0]), device=torch.device('cuda', local_rank))
        except Exception as e:  # (library without the "stream" query)
            sys.stderr.write('bench: own-stream timing unavailable (%s); using a torch stream\n' % (e,))
            stream = None
    config['stream'] = "the learner's own stream (torch.cuda.ExternalStream around dqn_device_ptr('stream'))"
    if stream is None:
        stream = torch.cuda.Stream(priority=-1)
        L.set_stream(stream.cuda_stream)
        config['stream'] = 'torch.cuda.Stream(priority=-1) handed to the learner (dqn_set_stream)'
    torch.cuda.set_stream(stream)
    shapes = tower_param_shapes(cfg['h'], cfg['w'], 4, cfg['a'], 512, cfg['dueling'], cfg['variant'])
    if args.no_fuse:
        L.set_option('fuse_fc_adam', 0)
    if args.fuse:
        L.set_option('fuse_fc_adam', 1)
    if args.no_overlap:
        L.set_option('overlap', 0)
    if args.no_pdl:
        L.set_option('pdl', 0)
    if args.ffma_adam:
        L.set_option('fc_ffma_adam', 1)
    # replicas: every learner has its own weights; data-parallel: all ranks start from the same weights
    wseed = 42 if dp else 42 + 2 * rank
    L.set_tower(init_tower_params(shapes, wseed), 0)
    L.set_tower(init_tower_params(shapes, wseed + 1), 1)
    L.replay_fill_synthetic(capacity, seed=1234 + rank)
    L.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if dp:
        from deepqnetwork_b200.distributed import DataParallelLearner, grads_tensor
        D = DataParallelLearner(L, overlap=not args.no_overlap, learner_stream=stream)
        run_steps = lambda n: [D.train_step() for _ in range(n)]
    else:
        run_steps = L.train_steps

    # warm-up (also captures the CUDA graph of the fused step)
    step_before = L.get_step()
    run_steps(args.warmup)
    barrier()
    clk = ClockSampler(local_rank).start() if rank == 0 else None
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    run_steps(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - l0
    clocks = clk.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
        lsum = torch.tensor([launches], device='cuda', dtype=torch.int64)
        dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
        launches = int(lsum.item())
    value = (1 if dp else world) * args.steps / (ms / 1000.0)
    step_now = L.get_step()
    assert step_now == step_before + args.warmup + args.steps, 'step counter does not match the steps requested'

    if dp:
        # every rank must hold the same weights after the same number of exchanged steps
        chk = torch.tensor([float(np.float64(L.get_param('dense/kernel').astype(np.float64).sum()))], device='cuda', dtype=torch.float64)
        lo, hi = chk.clone(), chk.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(lo.item()) == float(hi.item()), 'data-parallel ranks diverged'
        # the exchange step alone: the same two buckets, back to back, nothing else on the GPU
        coll = None
        if world > 1:
            g = grads_tensor(L)
            from deepqnetwork_b200.distributed import gradient_buckets
            big, small = gradient_buckets(L.tensors())
            tb = [g[o:o + c] for o, c in big]
            nbytes = sum(c for _, c in big) * 4
            for _ in range(5):
                for t in tb:
                    dist.all_reduce(t)
            torch.cuda.synchronize(); dist.barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            iters = 50
            for _ in range(iters):
                for t in tb:
                    dist.all_reduce(t)
            c1.record(stream); torch.cuda.synchronize()
            us = c0.elapsed_time(c1) * 1000.0 / iters
            tm = torch.tensor([us], device='cuda', dtype=torch.float64); dist.all_reduce(tm, op=dist.ReduceOp.MAX); us = float(tm.item())
            coll = dict(what='NCCL all-reduce (sum, fp32) of the two dense-kernel gradient buckets', bytes=nbytes, us=us,
                        bus_gb_s=2.0 * (world - 1) / world * nbytes / (us * 1e-6) / 1e9, ranks=world,
                        exposed='overlapped with the conv backward chain (dqn_phase_wait "fc_grads"); the 0.4 MB bucket follows the backward pass')
        config.update(parallelism='data-parallel x%d (global batch %d = %d per GPU, replay and sum tree sharded, gradient all-reduce)' % (
            world, cfg['global_batch'], batch), batch=cfg['global_batch'], per_gpu_batch=batch, samples_per_s=value * cfg['global_batch'],
            collective=coll, replay_transitions=args.capacity)
        gf = 2 * batch * (nn['n_fwd'] * nn['F'] + nn['F'] + (nn['F'] - nn['F1']))  # SURVEY.md 8d, per GPU
        config['gflop_per_step'] = round(gf * world / 1e9, 2)
        config['workload'] = '%s, %dx%dx4 uint8, global batch %d over %d GPU(s), %d-transition replay sharded in HBM' % (
            cfg['name'], cfg['h'], cfg['w'], cfg['global_batch'], world, args.capacity)

    # ---- per-kernel profile pass (CUDA events around every launch, un-graphed) -> roofline of the dominant kernel
    roofline = None
    if rank == 0 and not dp:
        prof_steps = max(20, min(200, args.steps // 10))
        L.set_profile(True)
        L.train_steps(prof_steps)
        L.sync()
        prof, n = L.get_profile()
        L.set_profile(False)
        per_kernel = {k: v / n * 1000.0 for k, v in prof.items()}  # us per step
        nn['early_fc'] = 'optimizer_fc' in per_kernel or 'fc_wgrad_adam' in per_kernel
        nn['split_fc'] = 'optimizer_fc1' in per_kernel
        total_us = sum(per_kernel.values())
        dom = max(per_kernel, key=per_kernel.get)
        kb = kernel_bytes(dom, cfg, nn)
        peak, how = measured_peak_hbm()
        achieved = kb / (per_kernel[dom] * 1e-6) / 1e9 if kb else None
        roofline = dict(bound='hbm', kernel=dom, achieved=achieved, peak=peak, unit='GB/s',
                        frac=(achieved / peak) if achieved else None, traffic=NCU_TRAFFIC.get(dom), peak_source=how,
                        algorithmic_bytes_per_launch=kb, us_per_launch=per_kernel[dom],
                        kernel_share_of_step=per_kernel[dom] / total_us, note=KERNEL_NOTES.get(dom),
                        step_frac_of_hbm_roofline=(nn['step_bytes'] / (ms / args.steps * 1e-3) / 1e9) / peak,
                        per_kernel_us={k: round(v, 2) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])})
    elif rank == 0:
        # C5 at per-GPU batch >= 256 is tensor-bound: whole-step FLOP/s against the measured bf16 peak
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peak_tf, how = float(json.load(f)['bf16_tflops_sustained']), 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
        except Exception:
            peak_tf, how = 1400.0, 'fallback (B200_PROFILING.md sustained bf16)'
        ach = config['gflop_per_step'] / world / (ms / args.steps * 1e-3) / 1e3
        roofline = dict(bound='tensor', kernel='whole data-parallel step (per GPU)', achieved=ach, peak=peak_tf, unit='TFLOP/s', frac=ach / peak_tf,
                        traffic=None, peak_source=how, note='fp32-grade arithmetic = 3 tf32 MMAs per product; FLOPs counted once')

    # ---- e2e ------------------------------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e and not dp:
        import random as pyrandom
        from deepqnetwork_b200.deep_q_agent import make_priority as host_make_priority
        S = nn['S']
        e2e_steps = max(50, min(args.steps, 500))
        pyrandom.seed(5)
        u_pin = torch.empty(BATCH, dtype=torch.float64).pin_memory()  # the step's host input: B random.random() draws
        u_np = u_pin.numpy()

        def draw_rows():
            size = capacity / BATCH
            return np.array([pyrandom.randint(int(i * size), int((i + 1) * size - 1)) for i in range(BATCH)], np.int64)

        def agent_step(i):
            # DeepQAgent._train_run_train (deepqnetwork_b200/deep_q_agent.py): host draws in, step / losses / loss out
            if cfg['per']:
                for k in range(BATCH):
                    u_np[k] = pyrandom.random()
                return L.train_step(uniforms=u_np)
            return L.train_step(rows=draw_rows())

        def agent_step_async(i):
            # the same step through dqn_train_step_async / dqn_train_step_result (DeepQAgent with pipeline_train): step i is queued,
            # then step i-1's (step, losses[B], loss) are read -- every step's draws go in and every step's results come out
            if cfg['per']:
                for k in range(BATCH):
                    u_np[k] = pyrandom.random()
                L.train_step_async(uniforms=u_np)
            else:
                L.train_step_async(rows=draw_rows())
            return L.train_step_result() if i > 0 else None

        def timed(fn, n, drain=None):
            for i in range(10):
                fn(i)
            if drain:
                drain()
            barrier()
            t0 = time.perf_counter()
            for i in range(n):
                fn(i)
            if drain:
                drain()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                tm = torch.tensor([dt], device='cuda', dtype=torch.float64)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                dt = float(tm.item())
            return world * n / dt

        v_sync = timed(agent_step, e2e_steps)
        v_async = timed(agent_step_async, e2e_steps, drain=lambda: L.train_step_result())
        e2e = dict(value=v_async, unit=UNIT, h2d_bytes_per_step=BATCH * 8, d2h_bytes_per_step=BATCH * 4 + 8 + 4,
                   steps=e2e_steps, path='DeepQAgent._train_run_train as the drop-in agent runs it with pipeline_train: '
                   'dqn_train_step_async(h_u = B host draws from pinned memory) queues step i, dqn_train_step_result returns '
                   '(step, losses[B], loss) of step i-1 to the host -- every step, in order; replay, sum tree and batch stay in HBM',
                   synchronous=dict(value=v_sync, unit=UNIT, path='dqn_train_step(h_u) -> (step, losses[B], loss): the host waits for '
                                    'the step it has just launched'),
                   excluded='replay append (actor side, once per environment step, not per train step)')

        # the other integration: seams 3 + 2 + 3 with a HOST replay -- the sampled rows are gathered on the host from a pinned
        # host ring and uploaded every step (ReplayMemory.copy x B, rl/data/replay_memory.py:113-122, paid on the host)
        R_host = 20000
        rng = np.random.default_rng(0)
        ring_s = torch.from_numpy(rng.integers(0, 256, (R_host, cfg['h'], cfg['w'], 4), dtype=np.uint8)).pin_memory().numpy()
        ring_n = torch.from_numpy(rng.integers(0, 256, (R_host, cfg['h'], cfg['w'], 4), dtype=np.uint8)).pin_memory().numpy()
        ring_a = rng.integers(0, cfg['a'], R_host).astype(np.uint8); ring_r = rng.choice([0, 1, 4, 7], R_host).astype(np.uint8)
        ring_c = (rng.random(R_host) < 0.99)
        bs = torch.empty((2, BATCH, cfg['h'], cfg['w'], 4), dtype=torch.uint8).pin_memory().numpy()  # s and s' in one pinned block
        rs = np.random.default_rng(5)

        def seam_step(i):
            if cfg['per']:
                t, p, w = L.per_sample(rs.random(BATCH), i % 5000 == 0, 0.4)        # seam 3: H2D u, D2H idx/prio/is_w
                rows = (t - (capacity - 1)) % R_host                                # the host ring stands in for the 1M-row one
            else:
                rows = draw_rows() % R_host; w = None
            np.take(ring_s, rows, axis=0, out=bs[0]); np.take(ring_n, rows, axis=0, out=bs[1])   # host gather of the 2B rows
            step, losses, loss = L.train_batch(bs[0], ring_a[rows], ring_r[rows], ring_c[rows], bs[1], w)  # seam 2: H2D batch, D2H losses
            if cfg['per']:
                L.tree_update(t, host_make_priority(losses), store_losses=True)     # seam 3: H2D idx/prio
            return loss

        hb_steps = max(50, min(args.steps, 300))
        for i in range(10):
            seam_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(hb_steps):
            seam_step(i)
        torch.cuda.synchronize()
        dth = time.perf_counter() - t0
        e2e['host_batch'] = dict(value=hb_steps / dth, unit=UNIT, steps=hb_steps,
                                 h2d_bytes_per_step=2 * BATCH * S + 3 * BATCH + (BATCH * (8 + 4 + 8 + 4) if cfg['per'] else 0),
                                 d2h_bytes_per_step=BATCH * 4 + 12 + (BATCH * 32 if cfg['per'] else 0),
                                 path='seam 3 dqn_per_sample -> host gather of the sampled rows from a pinned %d-row host ring -> seam 2 '
                                      'dqn_train_batch (H2D of the batch) -> seam 3 dqn_tree_update' % R_host)
    elif dp and not args.no_e2e:
        # the data-parallel step is driven from the host every step already (phase 0, collectives, phase 1): its wall-clock
        # rate, with the scalar loss read back every step, is the end-to-end number
        e2e_steps = max(20, min(args.steps, 200))
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            D.train_step()
            L.get_stats()  # device -> host read of the step's scalars (loss, PER statistics)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tm = torch.tensor([dt], device='cuda', dtype=torch.float64); dist.all_reduce(tm, op=dist.ReduceOp.MAX); dt = float(tm.item())
        e2e = dict(value=e2e_steps / dt, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=64, steps=e2e_steps,
                   path='DataParallelLearner.train_step (device Philox draws, replay shards in HBM) + scalar read-back, wall clock')

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not dp:
        v, done, threads = cpu_reference_run(cfg, 10 ** 9, 3, 20000, args.cpu_seconds)
        v1, done1, _ = cpu_reference_run(cfg, 10 ** 9, 1, 20000, min(8.0, args.cpu_seconds), threads=1)
        cpu = dict(value=v, unit=UNIT, cores=threads, kind='port', value_1_thread=v1,
                   sample='%d consecutive train steps (%.0f s) on %d threads, %d steps on 1 thread; 20000-row host replay; NumPy/Python '
                          'replay+PER as in rl/data/*.py, torch-CPU fp32 network (stand-in for TF 1.x CPU)' % (done, args.cpu_seconds, threads, done1))

    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())  # the learner's own stream dies with the handle: nothing of torch's may refer to it
    L.close()
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='strong' if dp else 'weak', vs_baseline=None,
                    dtype={'fp32': 'f32', 'tc3x': 'f32 (tf32x3 on tcgen05)', 'tc1x': 'tf32'}[args.math],
                    data='synthetic', config=config, clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roofline,
                    cpu_baseline=cpu)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
