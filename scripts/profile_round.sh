# Round profile of the DEFAULT configuration (run under gpurun; one GPU).
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --spp 50 --no-cpu-baseline"
$B > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_default.csv $B > gpurun_out/ncu_launches.log 2>&1
$B > gpurun_out/prof_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:path_trace -s 1 -c 1 -o gpurun_out/prof_default $B > gpurun_out/ncu_full.log 2>&1
python scripts/diag.py 100 > gpurun_out/diag_default.log 2>&1
echo done
