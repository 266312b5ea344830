set -x
mkdir -p gpurun_out
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err
cat gpurun_out/r14_bench.json
ncu --metrics $M --clock-control none -k regex:poa_b200 -c 1 --csv --log-file gpurun_out/r14_metrics.csv python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r14_metrics.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r14_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r14_pytest_gpu.log
tail -3 gpurun_out/r14_pytest_gpu.log
