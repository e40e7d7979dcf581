set -x
python -m pytest tests/test_gpu_mctx.py tests/test_gpu_seeds.py -q -x 2>&1 | tail -15
