#!/bin/bash
# GPU check on one B200, most important evidence first (a call may be cut short by the budget):
#   gpurun --timeout 600 -- 'bash tools/gpu_check.sh [tag]'
# 1 parity tests of the kernels   2 A/B + bit-identity of the PR kernel sets at the bench size
# 3 bench line   4 ncu launch list of the bench command   5 whole GPU suite   6 ncu --set full
TAG=${1:-check}
mkdir -p gpurun_out
S=gpurun_out/status_$TAG.txt
: > $S
date +%s > gpurun_out/t0
el() { echo $(( $(date +%s) - $(cat gpurun_out/t0) )); }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $S 2>&1

timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_segm.py tests/test_mask_codec.py \
    -m gpu -q > gpurun_out/t_kernels_$TAG.log 2>&1
echo "kernel_tests rc=$? t=$(el)" >> $S
tail -1 gpurun_out/t_kernels_$TAG.log >> $S

timeout 150 python tools/ab_variants.py --out gpurun_out/ab_variants_$TAG.json > gpurun_out/ab_$TAG.log 2>&1
echo "ab rc=$? t=$(el)" >> $S
tail -3 gpurun_out/ab_$TAG.log >> $S

timeout 180 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$? t=$(el)" >> $S

timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_list_$TAG.log 2>&1
echo "ncu_list rc=$? t=$(el)" >> $S

timeout 400 python -m pytest tests -m gpu -q > gpurun_out/t_all_$TAG.log 2>&1
echo "all_tests rc=$? t=$(el)" >> $S
tail -1 gpurun_out/t_all_$TAG.log >> $S

timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'k_frame_flat|k_track_iou_tiled|k_pr_bits|k_pr_envelope_bits|k_pr_finalize_tile' -c 10 \
    -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu_full rc=$? t=$(el)" >> $S
cat $S
