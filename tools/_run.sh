set -x
LCB_AGG_LANES=2 python -m pytest tests/test_gpu_parity.py -q -x -k "agg_coefs or bklm" 2>&1 | tail -5
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x -k "agg_coefs or bklm" 2>&1 | tail -5
timeout 600 python tools/agg_coefs_timing.py 16 8192 57344 2>&1 | tail -8
timeout 600 python tools/agg_coefs_timing.py 16 16384 0 2>&1 | tail -8
timeout 900 python tools/agg_coefs_timing.py 16 65536 0 2>&1 | tail -8
