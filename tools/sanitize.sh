#!/bin/bash
# compute-sanitizer passes over the small-batch GPU parity tests; run on a GPU box:
#   gpurun --timeout 2400 -- 'bash tools/sanitize.sh'
# Writes gpurun_out/sanitize_<tool>.log (full) and gpurun_out/sanitize_<tool>.summary.txt (what profiles/ keeps).
set -u
mkdir -p gpurun_out
SEL_MEM='golden or boundaries or field_widths or packed or agg_coefs or generic or multi_device or csprng'
SEL_RACE='golden or packed or generic'
run() {   # tool, selector, per-run limit in seconds, extra flags
    local tool=$1 sel=$2 lim=$3; shift 3
    local log=gpurun_out/sanitize_${tool}.log
    timeout "$lim" compute-sanitizer --tool "$tool" --target-processes all --error-exitcode 86 "$@" \
        python -m pytest tests -q -m gpu -x -k "$sel" -p no:cacheprovider > "$log" 2>&1
    local rc=$?
    {
        echo "tool: $tool   selector: $sel   exit: $rc   (86 = sanitizer reported errors, 124 = time limit)"
        echo "commit: $(cat gpurun_out/.commit 2>/dev/null)"
        grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error' "$log" | tail -20
        grep -E -c 'Invalid __|Race reported|Uninitialized __' "$log" | sed 's/^/hazard or invalid-access records: /'
    } > gpurun_out/sanitize_${tool}.summary.txt
    cat gpurun_out/sanitize_${tool}.summary.txt
}
run memcheck "$SEL_MEM" 1500
run racecheck "$SEL_RACE" 900 --racecheck-report all
