set -x
python -m pytest tests/test_gpu_checked.py -q -x 2>&1 | tail -15
python tools/sign_timing.py 128 20 2>&1 | tail -3
python tools/sign_timing.py 256 18 2>&1 | tail -3
