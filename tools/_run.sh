set -x
python -m pytest tests/test_gpu_wide.py -q -x 2>&1 | tail -30
