set -x
mkdir -p gpurun_out
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err
cat gpurun_out/r15_bench.json
POA_B200_LIB=$PWD/smoothxg_b200/lib/libpoa_b200_mb13.so python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r15_mb13.json 2>> gpurun_out/r15_bench.err
ncu --metrics $M --clock-control none -k regex:poa_b200 -c 1 --csv --log-file gpurun_out/r15_metrics.csv python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r15_metrics.log 2>&1
