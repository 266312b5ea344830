set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r08_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r08_pytest_gpu.log
tail -5 gpurun_out/r08_pytest_gpu.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r08_bench.json 2> gpurun_out/r08_bench.err
cat gpurun_out/r08_bench.json
python bench.py --workload 1000x16x1kb --warps 1 --steps 2 --warmup 1 --no-cpu --no-e2e >> gpurun_out/r08_variants.jsonl 2>> gpurun_out/r08_variants.err
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active --clock-control none -k regex:poa_b200 -c 1 --csv --log-file gpurun_out/r08_metrics.csv python bench.py --blocks 1776 --warps 1 --ctas-per-sm 12 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r08_metrics.log 2>&1
