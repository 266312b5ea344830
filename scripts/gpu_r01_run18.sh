set -x
mkdir -p gpurun_out
python bench.py --no-cpu > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err
cat gpurun_out/r18_bench.json
compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r18_memcheck.log 2>&1; tail -5 gpurun_out/r18_memcheck.log
compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r18_racecheck.log 2>&1; tail -5 gpurun_out/r18_racecheck.log
