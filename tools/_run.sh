set -x
timeout 300 python tools/verify_timing.py 128 20 10 2>&1 | tail -2
