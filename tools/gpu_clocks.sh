#!/bin/bash
# Samples SM clock / power every 50 ms while the default bench runs; prints the distribution.
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown --format=csv,noheader -lms 50 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_clk.json 2>gpurun_out/bench_clk.err
kill $SMI
python - <<'PY'
import collections
rows=[l.strip().split(', ') for l in open('gpurun_out/clocks.csv') if l.strip()]
c=collections.Counter((r[0],r[4]) for r in rows)
print('samples',len(rows))
for k,v in sorted(c.items()): print(k,v)
pw=[float(r[2].split()[0]) for r in rows]
print('power max',max(pw),'mean',sum(pw)/len(pw))
PY
python tools/show_bench.py < gpurun_out/bench_clk.json
