mkdir -p gpurun_out
B200RT_KERNEL=3 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu_v3.log
rm -f gpurun_out/diag_v3.log
run() { env $1 timeout 300 python scripts/diag.py 100 >> gpurun_out/diag_v3.log 2>&1; }
run "B200RT_KERNEL=2 B200RT_TRAV_THRESHOLD=8 B200RT_BLOCK=768"
run "B200RT_KERNEL=3"
for I in 8 16 26; do run "B200RT_KERNEL=3 B200RT_WF_INNER=$I"; done
for F in 4 16 24; do run "B200RT_KERNEL=3 B200RT_WF_FETCH=$F"; done
for PK in 16; do run "B200RT_KERNEL=3 B200RT_WF_PARK=$PK"; done
for PL in 64 128; do run "B200RT_KERNEL=3 B200RT_WF_POOL=$PL"; done
run "B200RT_KERNEL=3 B200RT_FAST_SLAB=0"
echo done
