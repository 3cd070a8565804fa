for cr in 0 8192 4096 2048; do
python bench.py --chunk-rows $cr --steps 2 --warmup 2 --no-cpu-baseline --no-parity --train-steps 0 > gpurun_out/ab_c$cr.json 2> gpurun_out/ab_c$cr.err
python - $cr <<'PY'
import json, sys
t=open(f"gpurun_out/ab_c{sys.argv[1]}.json").read().strip()
if not t: print(sys.argv[1], "FAILED", open(f"gpurun_out/ab_c{sys.argv[1]}.err").read()[-400:])
else:
    l=json.loads(t.splitlines()[-1]); print("chunk_rows", sys.argv[1], "value %.3f e2e %.3f" % (l["value"], l["e2e"]["value"]), l["roofline"]["class_ms_per_step"])
PY
done
