#!/bin/bash
# DRAM traffic of the layer-pipelined block1 launch (LD_GEMM_PIPE=1) against the four separate launches: usage  bash tools/gpu_pipe_dram.sh [env ...]
mkdir -p gpurun_out
SMALL="python bench.py --channels 1 --minutes 10 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --train-steps 0"
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
env LD_GEMM_PIPE=1 "$@" $SMALL > gpurun_out/plain_pipe.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_pipe.log; exit 1; }
env LD_GEMM_PIPE=1 "$@" ncu --metrics $M --clock-control none -k regex:gemm_taps_pipe -s 2 -c 2 --csv --log-file gpurun_out/pipe_dram.csv $SMALL > gpurun_out/ncu_pipe.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/pipe_dram.csv")) if len(r) > 10]
h = rows[0]; ik, im, iu, iv, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
for r in rows[1:]:
    print(r[ii], r[ik][:40], r[im], r[iv], r[iu])
PY
