# multi-GPU measurement set (run under gpurun --gpus 8): C5 data-parallel at 2/4/8, C4 replicas at 8
R=${1:-r02}
P=29600
for n in 2 4 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --config c5 --steps 50 --warmup 5 > gpurun_out/${R}_bench_c5_${n}gpu.json 2> gpurun_out/${R}_bench_c5_${n}gpu.err || tail -5 gpurun_out/${R}_bench_c5_${n}gpu.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 500 --warmup 50 > gpurun_out/${R}_bench_c4_8gpu.json 2> gpurun_out/${R}_bench_c4_8gpu.err || tail -5 gpurun_out/${R}_bench_c4_8gpu.err
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --config c5 --steps 5 --warmup 3 --no-e2e 2>&1 | grep -i "NVLS\|Connected\|nRanks\|Using network\|comm 0x" | head -12 > gpurun_out/${R}_nccl_info_8gpu.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_c[45]_*gpu.json')):
    try: d = json.load(open(f))
    except Exception as e: print(f, 'unreadable', e); continue
    c = d['config'].get('collective') or {}
    print('%-36s n=%d %9.1f steps/s %8.1f us  samples/s %s  e2e %8.1f  allreduce %s us bus %s GB/s' % (f.split('/')[-1], d['n_gpus'], d['value'], d['ms_per_step'] * 1000,
          d['config'].get('samples_per_s'), (d.get('e2e') or {}).get('value', 0), c.get('us'), c.get('bus_gb_s')))
PY
cat gpurun_out/${R}_nccl_info_8gpu.txt | head -6
