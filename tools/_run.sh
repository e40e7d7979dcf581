set -x
timeout 300 python tools/verify_timing.py 128 20 10 2>&1 | tail -4
timeout 300 python tools/verify_timing.py 128 17 10 2>&1 | tail -2
LCB_VERIFY_FUSED=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_verify -s 3 -c 1 -o gpurun_out/prof_r2_verify_fused_a python tools/verify_timing.py 128 18 2 > gpurun_out/ncu_fused.log 2>&1
tail -3 gpurun_out/ncu_fused.log
