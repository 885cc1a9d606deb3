#!/bin/bash
# One-shot GPU validation of the runtime-selectable variants added late in round 1
# (TA_PR_IMPL=1 bit-plane PR kernels, TA_FF_NODIV=1 flat frame kernel, segm path).
# Most important evidence first: the call may be cut short by the remaining GPU budget.
#   gpurun --timeout 700 -- 'bash tools/gpu_check_r1b.sh'
mkdir -p gpurun_out
S=gpurun_out/status_r1b.txt
: > $S
date +%s > gpurun_out/t0
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $S 2>&1

# 1. parity of the new code paths (goldens of the unmodified reference, random differential)
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_segm.py tests/test_mask_codec.py \
    -m gpu -q -k "pr_accumulate_both or each_pr_implementation or each_candidate_variant or segm or rle_iou" \
    > gpurun_out/t_new_r1b.log 2>&1
echo "new_tests rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" >> $S
tail -3 gpurun_out/t_new_r1b.log >> $S

# 2. A/B at the bench workload + bit-identity of every output tensor with the pinned baseline
timeout 150 python tools/ab_variants.py --out gpurun_out/ab_variants.json > gpurun_out/ab_r1b.log 2>&1
echo "ab rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" >> $S
cat gpurun_out/ab_r1b.log | tail -5 >> $S

# 3. the bench line with the new variants
TA_PR_IMPL=1 TA_FF_NODIV=1 timeout 180 python bench.py > gpurun_out/bench_r1b_new.json 2> gpurun_out/bench_r1b_new.err
echo "bench_new rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" >> $S

# 4. ncu launch list of the same command (shares only: cold caches, serialised)
TA_PR_IMPL=1 TA_FF_NODIV=1 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv \
    --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_list_r1b.log 2>&1
echo "ncu_list rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" >> $S

# 5. the whole GPU suite with the new variants as the process-wide choice (= flipped defaults;
#    the tests that pin a variant explicitly still run both)
TA_PR_IMPL=1 TA_FF_NODIV=1 timeout 400 python -m pytest tests -m gpu -q > gpurun_out/t_all_r1b.log 2>&1
echo "all_tests rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" >> $S
tail -3 gpurun_out/t_all_r1b.log >> $S

# 6. one full-set capture of the new kernels
TA_PR_IMPL=1 TA_FF_NODIV=1 timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'k_pr_bits|k_pr_envelope_bits|k_pr_scan_live|k_pr_finalize_2d|k_frame_flat' -c 10 \
    -o gpurun_out/prof_r1b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_r1b.log 2>&1
echo "ncu_full rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))" >> $S
cat $S
