#!/bin/bash
# Round-2 evidence run (one GPU): bench line, launch list of the same command, per-kernel rooflines,
# ncu --set full of the fused 2-D kernel, the RQ-spline apply kernel and the N-D tensor-core kernels.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 300 gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
python scratch/kernel_roofline.py > gpurun_out/r02_kernel_roofline.log 2>&1
cp profiles/r02_kernel_roofline.json gpurun_out/ 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:fused2d_tc -s 4 -c 1 -o gpurun_out/r02_fused2d_tc \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
B4=256 B5=32 ncu --set full --clock-control none --import-source on -k regex:"nd_layer1|convnd" -s 9 -c 3 -o gpurun_out/r02_nd3d \
    python scratch/nd_time.py > /dev/null 2>&1
B4=256 B5=32 ncu --set full --clock-control none --import-source on -k regex:"nd_layer1|convnd" -s 27 -c 3 -o gpurun_out/r02_nd4d \
    python scratch/nd_time.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
