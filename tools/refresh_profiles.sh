#!/bin/bash
# Runs ON the GPU box (gpurun -- bash tools/refresh_profiles.sh): plain bench lines first, then the ncu
# launch list and one --set full capture per kernel, all into gpurun_out/ (digests are made afterwards
# with profiles/ncu_summary.py and copied into profiles/).
set -u
R=${ROUND:-r2}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -1 | tee $O/pytest_gpu_${R}.txt
python bench.py > $O/bench_${R}_final.json 2> $O/bench_${R}_final.err || { echo "bench failed"; tail -5 $O/bench_${R}_final.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${R}_reference.json 2> $O/bench_${R}_reference.err || echo "reference arm failed"
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --bklm-log2n 12"
$SHORT > $O/plain_${R}.log 2>&1 || { echo "short bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_${R}.csv $SHORT > $O/ncu_l.log 2>&1
echo "launch list rc=$?"
CAP="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --log2n 18 --bklm-log2n 13"
$CAP > $O/plain2_${R}.log 2>&1 || { echo "capture bench failed"; exit 1; }
cap() {  # name, kernel regex, launches to skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o $O/prof_${R}_$1 $CAP > $O/ncu_$1.log 2>&1
  echo "capture $1 rc=$?"
}
cap verify k_verify 2
cap sampler_sk k_sampler 1
cap sampler_challenge k_sampler 5
cap matvec k_matvec 0
cap sign k_sign 0
cap agg_coefs k_agg_coefs 0
cap unpack k_unpack 1
ls -la $O | grep prof_${R}
