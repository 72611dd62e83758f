CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --capacity 50000"
$CMD > gpurun_out/ncu_plain_r7.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r7.csv $CMD > gpurun_out/ncu_launch_r7.log 2>&1
$CMD > gpurun_out/ncu_plain_r7b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"cvconv|cvwgrad|adam_kernel|fcwg_adam|fcfwd|tsgemm" -s 50 -c 22 -o gpurun_out/prof_r7 $CMD > gpurun_out/ncu_full_r7.log 2>&1
tail -n 2 gpurun_out/ncu_full_r7.log
python tools/trace_step.py 1000000 > gpurun_out/trace_r7.txt 2>&1; cp gpurun_out/trace.csv gpurun_out/trace_r7.csv
python bench.py > gpurun_out/bench_r7.json 2> gpurun_out/bench_r7.err; tail -2 gpurun_out/bench_r7.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r7_ref.json 2> gpurun_out/bench_r7_ref.err; tail -2 gpurun_out/bench_r7_ref.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ls -la gpurun_out/ | tail -5
