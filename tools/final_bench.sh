# round-end measurement set on one GPU (run under gpurun from the repo root); results -> gpurun_out/r02_*
R=${1:-r02}
python bench.py > gpurun_out/${R}_bench_c2.json 2> gpurun_out/${R}_bench_c2.err || tail -5 gpurun_out/${R}_bench_c2.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_c2_driver_args.json 2> /dev/null
for c in c1 c3 nature; do python bench.py --config $c > gpurun_out/${R}_bench_$c.json 2> gpurun_out/${R}_bench_$c.err || tail -5 gpurun_out/${R}_bench_$c.err; done
python bench.py --config c5 --steps 50 --warmup 5 > gpurun_out/${R}_bench_c5_1gpu.json 2> gpurun_out/${R}_bench_c5_1gpu.err || tail -5 gpurun_out/${R}_bench_c5_1gpu.err
python bench.py --impl reference --steps 200 --warmup 3 > gpurun_out/${R}_bench_reference_arm.json 2> /dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    e = d.get('e2e') or {}
    c = d.get('cpu_baseline') or {}
    r = d.get('roofline') or {}
    print('%-44s %9.1f steps/s %8.1f us  e2e %8.1f  host_batch %7.1f  cpu %6.1f / %5.1f  %s frac %.3f step_frac %s clocks %s' % (
        f.split('/')[-1], d['value'], d['ms_per_step'] * 1000, e.get('value', 0), (e.get('host_batch') or {}).get('value', 0),
        c.get('value', 0) or 0, c.get('value_1_thread', 0) or 0, r.get('kernel'), r.get('frac') or 0, r.get('step_frac_of_hbm_roofline'), (d.get('clocks') or {}).get('sm_mhz')))
PY
