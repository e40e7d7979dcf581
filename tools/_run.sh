set -x
python -m pytest tests/test_gpu_parity.py -q -x -k "agg_coefs" 2>&1 | tail -5
LCB_AGG_EARLY=1 python -m pytest tests/test_gpu_parity.py -q -x -k "agg_coefs" 2>&1 | tail -5
timeout 600 python tools/agg_coefs_timing.py 16 8192 57344 2>&1 | grep -v subsample
timeout 900 python tools/agg_coefs_timing.py 16 65536 0 2>&1 | grep -v subsample
timeout 900 python -m pytest tests/test_gpu_bklm_full.py -q -x 2>&1 | tail -15
