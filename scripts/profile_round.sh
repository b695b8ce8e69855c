# Round profile of the DEFAULT configuration (run under gpurun; one GPU).  Each ncu pass runs only
# after the same command exited 0 without the profiler.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
B="timeout 300 python bench.py --steps 2 --warmup 1 --spp ${SPP:-500} --no-cpu-baseline --no-other-configs"
$B > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
$B > gpurun_out/prof_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:path_trace -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_default $B > gpurun_out/ncu_full.log 2>&1
timeout 120 python scripts/diag.py 100 > gpurun_out/${TAG}_diag_default.log 2>&1
echo done
