#!/bin/bash
# profiles/capture.sh <tag> — the measurement set behind the numbers in DESIGN.md, run on one B200 under gpurun:
#   1. python bench.py (default flags) -> bench line; nothing below runs unless it exits 0
#   2. ncu launch list of a short bench command (per-launch gpu__time_duration, kernel shares of the step)
#   3. ncu --set full of one layer's three SpMM kernels (second layer launched: warm L2 / TLB)
#   4. ncu --set full of one launch of the f16 scorer
# Raw output goes to gpurun_out/<tag>/; summaries are made afterwards with profiles/ncu_summary.py.
set -u
tag=${1:-cap}
out=gpurun_out/$tag
mkdir -p $out
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err || { echo "bench failed"; tail -5 $out/bench_n1.err; exit 1; }
short="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-small-configs --eval-users 151552"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv $short > $out/ncu_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:spmm_ --launch-skip 3 --launch-count 3 -f -o $out/prof_spmm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-small-configs --no-eval --no-e2e --no-train > $out/ncu_spmm.log 2>&1
ncu -i $out/prof_spmm.ncu-rep --page raw --csv > $out/raw_spmm.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:score_topk_f16 --launch-skip 1 --launch-count 1 -f -o $out/prof_f16 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-small-configs --no-e2e --no-train --eval-users 151552 > $out/ncu_f16.log 2>&1
ncu -i $out/prof_f16.ncu-rep --page raw --csv > $out/raw_f16.csv 2>/dev/null
ls -la $out
