#!/bin/bash
# Round-2 evidence run (one GPU): bench line, launch list of the same command, per-kernel rooflines, step timings of
# the N-D / wide-conditioner paths, GPU test log with the printed tolerances, and ncu --set full captures of the fused
# 2-D kernel, the RQ-spline apply kernel, the N-D tensor-core kernels and the tiled N-D weight gradient.
set -x
python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 300 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
python scratch/kernel_roofline.py > gpurun_out/r02_kernel_roofline.log 2>&1
(python scratch/nd_time.py; python scratch/wide_time.py) > gpurun_out/r02_nd_step_times.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_nd_launches_raw.csv python scratch/nd_time.py > /dev/null 2>&1
B4=256 B5=32 ncu --set full --clock-control none --import-source on -k regex:"nd_layer1|convnd" -s 9 -c 3 -o gpurun_out/r02_nd3d \
    python scratch/nd_time.py > /dev/null 2>&1
B4=256 B5=32 ncu --set full --clock-control none --import-source on -k regex:"nd_layer1|convnd" -s 27 -c 3 -o gpurun_out/r02_nd4d \
    python scratch/nd_time.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"PairSites<nfk::RqsOp" -c 1 -o gpurun_out/r02_rqs_fwd \
    python scratch/kernel_roofline.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"prior_normal_fast" -c 1 -o gpurun_out/r02_prior \
    python scratch/kernel_roofline.py > /dev/null 2>&1
# tensor-core backward: weight gradient (MN-major operands) and data gradient of the 8 -> 28 layer at config-4 geometry
ncu --set full --clock-control none --import-source on -k regex:"convnd_wgrad_tc|convnd_tc_kernel" -s 2 -c 2 -o gpurun_out/r02_bwd3d \
    python scratch/bwd_ops.py 32x32x32 64 28 > /dev/null 2>&1
(python scratch/nd_train_time.py 4 64; python scratch/nd_train_time.py 5 16; NFK_DGRAD_TC=0 NFK_WGRAD_ND_TC=0 python scratch/nd_train_time.py 4 64; \
 NFK_DGRAD_TC=0 NFK_WGRAD_ND_TC=0 python scratch/nd_train_time.py 5 16) > gpurun_out/r02_nd_train_times.log 2>&1
# summaries stay, the reports (80 MB) do not: gpurun copies at most 64 MiB back
for n in nd3d nd4d rqs_fwd prior bwd3d; do
    python scratch/ncu_brief.py gpurun_out/r02_$n.ncu-rep > gpurun_out/r02_${n}_ncu_full.txt 2>/dev/null
    rm -f gpurun_out/r02_$n.ncu-rep
done
ls -la gpurun_out/
