set -x
python -m pytest tests -q -x -m gpu 2>&1 | tail -12
