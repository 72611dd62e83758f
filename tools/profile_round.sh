# ncu evidence for a round (run under gpurun from the repo root): launch list + --set full pages of the top kernels.
# Each ncu pass runs only after the same command exited 0 without ncu.
R=${1:-r02}
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --capacity 50000"
$CMD > gpurun_out/ncu_plain_$R.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/ncu_launch_$R.log 2>&1
$CMD > gpurun_out/ncu_plain_${R}b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"cvtower|cvconv|cvwgrad|fcwg_adam|fcfwd|tsgemm|gather_rows|tree_sample|tree_update|optimizer_tail" -s 60 -c 26 -o gpurun_out/${R}_prof $CMD > gpurun_out/ncu_full_$R.log 2>&1
tail -n 2 gpurun_out/ncu_full_$R.log
ncu -i gpurun_out/${R}_prof.ncu-rep --page raw --csv > gpurun_out/${R}_top_kernels_raw.csv 2>/dev/null
python tools/trace_step.py 1000000 > gpurun_out/${R}_cupti_step_trace.txt 2>&1
ls -la gpurun_out/ | tail -8
