set -x
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -4
python tools/sign_timing.py 128 20 2>&1 | tail -2
python tools/sign_timing.py 256 18 2>&1 | tail -1
