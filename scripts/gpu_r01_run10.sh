set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r10_pytest_gpu.log
tail -5 gpurun_out/r10_pytest_gpu.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err
cat gpurun_out/r10_bench.json
python bench.py --workload 1000x16x1kb --warps 1 --steps 2 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r10_variants.jsonl 2>> gpurun_out/r10_variants.err
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio --clock-control none -k regex:poa_b200 -c 1 --csv --log-file gpurun_out/r10_metrics.csv python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r10_metrics.log 2>&1
