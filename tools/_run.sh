set -x
nvidia-smi -L
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_multi.py tests/test_gpu_mctx.py -q -x 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_r2_2gpu_a.json 2> gpurun_out/bench_r2_2gpu_a.err; echo rc=$?; tail -3 gpurun_out/bench_r2_2gpu_a.err
