#!/bin/bash
# Runs ON the GPU box (gpurun -- bash tools/refresh_profiles.sh): plain bench lines first, then the ncu
# launch list and one --set full capture per kernel, all into gpurun_out/ (digests are made afterwards
# with tools/digest_profiles.sh and copied into profiles/).
set -u
R=${ROUND:-r2}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -1 | tee $O/pytest_gpu_${R}.txt
python bench.py > $O/bench_${R}_final.json 2> $O/bench_${R}_final.err || { echo "bench failed"; tail -5 $O/bench_${R}_final.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${R}_reference.json 2> $O/bench_${R}_reference.err || echo "reference arm failed"
SMALL="--no-cpu-baseline --e2e-steps 1 --keygen-log2n 14 --adaptor-log2n 12"
SHORT="python bench.py --steps 2 --warmup 3 $SMALL --bklm-log2n 12"
$SHORT > $O/plain_${R}.log 2>&1 || { echo "short bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_${R}.csv $SHORT > $O/ncu_l.log 2>&1
echo "launch list rc=$?"
# captures: 2^18 verifies per launch (digest_profiles.sh records units=262144), 8,192 aggregation streams of a 2^13 aggregate
CAP="python bench.py --steps 1 --warmup 3 $SMALL --log2n 18 --bklm-log2n 13"
$CAP > $O/plain2_${R}.log 2>&1 || { echo "capture bench failed"; exit 1; }
cap() {  # name, kernel regex, launches to skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o $O/prof_${R}_$1 $CAP > $O/ncu_$1.log 2>&1
  echo "capture $1 rc=$?"
}
cap verify k_verify 2
cap sampler_challenge k_sampler 5
cap sign k_sign 0
cap agg_coefs_il k_agg_coefs_il 0
cap matvec k_matvec 0
# the BKLM algebra kernels at the configured aggregate size (2^16 signatures per launch)
CAP="python bench.py --steps 1 --warmup 3 $SMALL --log2n 16 --bklm-log2n 16"
cap agg_partial k_agg_partial 1
cap aggv_partial k_aggv_partial 1
ls -la $O | grep prof_${R}
python tools/sign_timing.py 128 20 2>&1 | tail -1
python tools/sign_timing.py 256 18 2>&1 | tail -1
python tools/verify_timing.py 128 20 10 2>&1 | tail -1
python tools/verify_timing.py 256 18 10 2>&1 | tail -1
