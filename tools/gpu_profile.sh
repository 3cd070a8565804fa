#!/bin/bash
# ncu evidence for profiles/ (one ncu pass per gpurun call; each follows a plain run of the same command, B200_PROFILING.md):
#   tools/gpu_profile.sh launches   launch list (gpu__time_duration) of a short bench run
#   tools/gpu_profile.sh full       --set full capture of the four block1 conv GEMM launches of the timed step
mkdir -p gpurun_out
SMALL="python bench.py --channels 1 --minutes 10 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --train-steps 0"
KERNELS='regex:gemm_taps|stem_kernel|head_kernel|fbank_kernel|pcm_sum|segment_'
$SMALL > gpurun_out/plain_small.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_small.log; exit 1; }
if [ "$1" = "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:gemm_taps -s 38 -c 4 -o gpurun_out/prof_gemm $SMALL > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
else
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
fi
tail -2 gpurun_out/plain_small.log | cut -c1-300
ls -la gpurun_out | tail -8
