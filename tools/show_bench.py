import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value',round(d['value'],3),'e2e',round(d['e2e']['value'],3),'ms/step',round(d['ms_per_step'],1),'gemm ms',round(r['kernel_ms_per_step'],1),'exec frac',round(r['executed_frac'],3), 'fbank ms', round(r['fbank']['ms_per_step'],2))
        print(r['per_conv_ms_per_step'])
        print(r.get('class_ms_per_step'))
        print('train', d.get('train'))
        print('cpu', d.get('cpu_baseline'))
    else: print(l.strip()[:400])
