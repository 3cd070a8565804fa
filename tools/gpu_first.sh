#!/bin/bash
# First GPU pass of a change: parity tests, a short bench, the ncu launch list and one full capture of the conv GEMM.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
SMALL="python bench.py --channels 1 --minutes 10 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
$SMALL > gpurun_out/plain_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_taps -s 4 -c 3 -o gpurun_out/prof_gemm $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
