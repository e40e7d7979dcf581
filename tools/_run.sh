set -x
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py tests/test_gpu_mctx.py -q -x 2>&1 | tail -4
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err; echo rc=$?; tail -2 gpurun_out/bench_r2_${N}gpu.err
done
