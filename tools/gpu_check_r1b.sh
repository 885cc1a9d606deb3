#!/bin/bash
# One-shot GPU validation of the PR-accumulation variants (TA_PR_IMPL=0..3, csrc/ta_pr.cu).
# Most important evidence first: the call may be cut short by the remaining GPU budget.
#   gpurun --timeout 280 -- 'bash tools/gpu_check_r1b.sh'
mkdir -p gpurun_out
S=gpurun_out/status_r1c.txt
: > $S
date +%s > gpurun_out/t0
el() { echo $(( $(date +%s) - $(cat gpurun_out/t0) )); }

# 1. parity of every variant (goldens of the unmodified reference, random multi-chunk cases)
timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "pr_accumulate_both or each_pr_implementation" > gpurun_out/t_new_r1c.log 2>&1
echo "new_tests rc=$? t=$(el)" >> $S
tail -1 gpurun_out/t_new_r1c.log >> $S

# 2. A/B at the bench workload + bit-identity of every output tensor with variant 0
timeout 120 python tools/ab_variants.py --out gpurun_out/ab_variants_r1c.json > gpurun_out/ab_r1c.log 2>&1
echo "ab rc=$? t=$(el)" >> $S
tail -4 gpurun_out/ab_r1c.log >> $S
BEST=$(python - <<'PY'
import json
try:
    v = json.load(open("gpurun_out/ab_variants_r1c.json"))["variants"]
    ok = {k: r for k, r in v.items() if r.get("all_identical") and "ms_per_step" in r}
    print(min(ok, key=lambda k: ok[k]["ms_per_step"])[2:] if ok else 0)
except Exception:
    print(0)
PY
)
echo "fastest verified variant: TA_PR_IMPL=$BEST" >> $S
export TA_PR_IMPL=$BEST

# 3. the bench line with that variant
timeout 150 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
echo "bench rc=$? t=$(el)" >> $S

# 4. ncu launch list of the same command (shares only: cold caches, serialised)
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv \
    --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_list_r1c.log 2>&1
echo "ncu_list rc=$? t=$(el)" >> $S

# 5. the whole GPU suite with that variant as the process-wide choice (= the shipped default)
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/t_all_r1c.log 2>&1
echo "all_tests rc=$? t=$(el)" >> $S
tail -1 gpurun_out/t_all_r1c.log >> $S
cat $S
