#!/bin/bash
# Runs HERE (no GPU needed) after `gpurun -- bash tools/refresh_profiles.sh`: turns gpurun_out/ into the tracked
# records under profiles/ (ncu digests, launch summary, bench lines).  ROUND=r2 tools/digest_profiles.sh for round 2.
set -eu
R=${ROUND:-r2}
cd "$(dirname "$0")/.."
for rep in gpurun_out/prof_${R}_*.ncu-rep; do
  k=$(basename "$rep" .ncu-rep)
  python profiles/ncu_summary.py "$rep" > "profiles/$k.summary.txt" 2>&1 || { echo "digest of $k failed"; rm -f "profiles/$k.summary.txt"; continue; }
  # what the captured launch processed (tools/refresh_profiles.sh captures `bench.py --log2n 18`) and the commit the
  # binary was built from: bench.py only quotes digests that carry this line
  case "$k" in
    *aggv_partial*|*agg_partial*) units=65536 ;;
    *verify*|*sign*) units=262144 ;;
    *) units= ;;
  esac
  [ -n "$units" ] && echo "# capture: units=$units commit=$(cat gpurun_out/commit_${R}.txt 2>/dev/null || git rev-parse --short HEAD)" >> "profiles/$k.summary.txt"
done
for f in bench_${R}_final.json bench_${R}_reference.json bench_${R}_2gpu.json bench_${R}_4gpu.json bench_${R}_8gpu.json launches_${R}.csv; do
  [ -s "gpurun_out/$f" ] && cp "gpurun_out/$f" profiles/
done
[ -s "profiles/launches_${R}.csv" ] && python profiles/launch_summary.py "profiles/launches_${R}.csv" > "profiles/launches_${R}_summary.csv"
ls -la profiles | grep "_${R}" | awk '{print $5, $9}'
