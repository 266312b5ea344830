set -x
mkdir -p gpurun_out
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
cat gpurun_out/r11_bench.json
for v in v1 v2; do POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_$v.so python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r11_variants.jsonl 2>> gpurun_out/r11_variants.err; done
for v in "" _v1 _v2; do POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200$v.so ncu --metrics $M --clock-control none -k regex:poa_b200 -c 1 --csv --log-file gpurun_out/r11_metrics$v.csv python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r11_metrics$v.log 2>&1; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poa_b200 -c 1 -o gpurun_out/r11_full python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r11_ncu_full.log 2>&1
