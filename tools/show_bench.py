import json, sys
for l in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.json'):
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value', round(d['value'], 3), 'e2e', round(d['e2e']['value'], 3), 'ms/step', round(d['ms_per_step'], 1), 'gemm ms',
              round(r['kernel_ms_per_step'], 1), 'hbm GB/s', round(r['achieved']), 'frac', round(r['frac'], 3), 'tensor exec frac',
              round(r['tensor']['executed_frac'], 3), 'fbank ms', round(r['fbank']['ms_per_step'], 2))
        if 'train' in d:
            print('train', round(d['train']['value']), 'samples/s', round(d['train']['ms_per_step'], 2), 'ms/step')
        if 'cpu_baseline' in d:
            print('cpu', d['cpu_baseline']['value'])
