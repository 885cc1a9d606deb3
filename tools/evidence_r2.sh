#!/bin/bash
# Round-2 evidence on one B200 (gpurun --timeout 2400 -- 'bash tools/evidence_r2.sh'):
#   1 whole GPU test suite   2 bench line (driver's arguments) + reference arm
#   3 ncu launch list of the bench command   4 ncu --set full of the top kernels
#   5 compute-sanitizer memcheck / racecheck / synccheck on the small parity cases
mkdir -p gpurun_out
S=gpurun_out/r2_status.txt
: > $S
date +%s > gpurun_out/t0
el() { echo $(( $(date +%s) - $(cat gpurun_out/t0) )); }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $S 2>&1

timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_tests.log 2>&1
echo "all_tests rc=$? t=$(el)" >> $S
tail -1 gpurun_out/r2_tests.log >> $S

timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$? t=$(el)" >> $S

timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
echo "reference_arm rc=$? t=$(el)" >> $S

timeout 200 python tools/trace_host_call.py > /dev/null 2> gpurun_out/r2_host_call_trace.txt
echo "host_call_trace rc=$? t=$(el)" >> $S

for c in cfg2 cfg5; do
  timeout 300 python tools/bench_cli.py --config $c --cold --out gpurun_out/r2_cli_cold_$c.json > gpurun_out/r2_cli_cold_$c.log 2>&1
  echo "cli_cold_$c rc=$? t=$(el)" >> $S
done

timeout 200 python tools/ab_variants.py --out gpurun_out/r2_ab_variants.json > gpurun_out/r2_ab.log 2>&1
echo "ab rc=$? t=$(el)" >> $S
tail -3 gpurun_out/r2_ab.log >> $S

timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu_list rc=$? t=$(el)" >> $S

timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'k_frame_flat|k_track_iou_tiled|k_pr_bits|k_pr_envelope_bits|k_pr_finalize_tile|k_pr_scan_live|k_frame_eval|k_match_greedy' \
    -s 8 -c 14 -o gpurun_out/r2_prof python tools/ab_variants.py --variants 1 --steps 1 \
    --out gpurun_out/r2_ab_ncu.json > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu_full rc=$? t=$(el)" >> $S

for tool in memcheck racecheck synccheck; do
  sel="tiny or edge_mix or oversize or random_small"
  [ $tool != memcheck ] && sel="tiny or oversize"
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py \
      -m gpu -x -q -k "$sel" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "sanitizer_$tool rc=$? t=$(el)" >> $S
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_$tool.log | tail -3 >> $S
done
cat $S
