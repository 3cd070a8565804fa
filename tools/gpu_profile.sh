#!/bin/bash
# ncu evidence for profiles/: the launch list of a short bench run and one full capture of the block1 conv GEMM launches.
# Each ncu run follows a plain run of the same command (B200_PROFILING.md).
mkdir -p gpurun_out
SMALL="python bench.py --channels 1 --minutes 10 --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0"
KERNELS='regex:gemm_taps|stem_kernel|head_kernel|fbank_kernel|pcm_sum|segment_'
$SMALL > gpurun_out/plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
$SMALL > gpurun_out/plain_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_taps -s 38 -c 4 -o gpurun_out/prof_gemm $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
cat gpurun_out/plain_small.log | tail -2
ls -la gpurun_out
