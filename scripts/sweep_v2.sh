mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
B="python bench.py --steps 2 --warmup 1 --spp 100 --no-cpu-baseline"
run() { echo "### $1" >> gpurun_out/sweep.log; env $1 $B >> gpurun_out/sweep.log 2>> gpurun_out/sweep.err; }
rm -f gpurun_out/sweep.log gpurun_out/sweep.err
run "B200RT_KERNEL=1"
for T in 1 8 12 16 20 24 28; do run "B200RT_KERNEL=2 B200RT_TRAV_THRESHOLD=$T"; done
run "B200RT_KERNEL=2 B200RT_TRAV_THRESHOLD=16 B200RT_FAST_SLAB=0"
for BL in 512 768 1024; do run "B200RT_KERNEL=2 B200RT_TRAV_THRESHOLD=16 B200RT_BLOCK=$BL"; done
echo done
