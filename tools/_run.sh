set -x
python -m pytest tests -q -x -m gpu 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --log2n 16 --bklm-log2n 12 --keygen-log2n 14 --adaptor-log2n 12 --cpu-per-core 2 --cpu-bklm-log2n 4 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo rc=$?; tail -5 gpurun_out/bench_small.err
python bench.py > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo rc=$?; tail -5 gpurun_out/bench_r2_a.err
