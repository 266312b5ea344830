# round 2, call M (one GPU): sectioned ncu capture (source counters + warp states) of the 8-warp packed fill on 4 deep blocks
set -x
mkdir -p gpurun_out
timeout 800 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section InstructionStats --section LaunchStats --section Occupancy --import-source on --clock-control none -k regex:poa_b200_block -c 1 -o gpurun_out/r02m_deep python bench.py --workload 100x256x8kb --blocks 4 --warps 8 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r02m_ncu.log 2>&1
tail -3 gpurun_out/r02m_ncu.log | cut -c1-300; ls -la gpurun_out/r02m_deep.ncu-rep
