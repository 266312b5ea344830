# round 2, call K (one GPU): final-candidate build: parity, headline, deep blocks + a sectioned ncu capture of the deep-block kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log; tail -3 gpurun_out/r02k_pytest.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02k_config2.json 2> gpurun_out/r02k_config2.err
python -c "import json; d=json.load(open('gpurun_out/r02k_config2.json')); print('CONFIG2', round(d['value'],1), round(d['ms_per_step'],1))"
python bench.py --workload 100x256x8kb --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02k_config3.json 2> gpurun_out/r02k_config3.err
python -c "import json; d=json.load(open('gpurun_out/r02k_config3.json')); print('CONFIG3', round(d['value'],1), round(d['ms_per_step']), d['engine']['warps_per_block'])"
timeout 900 ncu --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section InstructionStats --clock-control none -k regex:poa_b200_block -c 1 --csv --log-file gpurun_out/r02k_ncu_config3_sections.csv python bench.py --workload 100x256x8kb --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r02k_ncu_config3.log 2>&1
tail -2 gpurun_out/r02k_ncu_config3.log | cut -c1-300
