# A/B of conv-kernel knobs on the default workload (short runs): usage  bash tools/gpu_ab.sh "LD_X=1" "LD_X=0 LD_Y=2" ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  env $cfg python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-parity --train-steps 0 > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - "$cfg" gpurun_out/ab_$i.json <<'PY'
import json, sys
t = open(sys.argv[2]).read().strip()
if not t:
    print(sys.argv[1], "FAILED", open(sys.argv[2].replace(".json", ".err")).read()[-600:])
else:
    l = json.loads(t.splitlines()[-1])
    r = l["roofline"]
    print(sys.argv[1], "| value %.3f e2e %.3f |" % (l["value"], l["e2e"]["value"]), r["class_ms_per_step"], {k: v for k, v in r["per_conv_ms_per_step"].items()})
PY
  i=$((i+1))
done
