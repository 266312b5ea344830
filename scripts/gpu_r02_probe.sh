# sectioned ncu capture of the 8-warp packed fill on a short deep-block probe (8 blocks x 64 x 4 kb)
set -x
mkdir -p gpurun_out
python bench.py --workload 8x64x4kb_probe --warps 8 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02q_probe.json 2> gpurun_out/r02q_probe.err
python -c "import json; d=json.load(open('gpurun_out/r02q_probe.json')); print('PROBE', round(d['value'],1), round(d['ms_per_step'],1), d['engine']['phase_cycles'])"
timeout 500 ncu --section WarpStateStats --section SchedulerStats --section InstructionStats --section LaunchStats --section Occupancy --section SourceCounters --import-source on --clock-control none -k regex:poa_b200_block -c 1 -o gpurun_out/r02q_probe python bench.py --workload 8x64x4kb_probe --warps 8 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r02q_ncu.log 2>&1
tail -2 gpurun_out/r02q_ncu.log | cut -c1-200; ls -la gpurun_out/r02q_probe.ncu-rep
