import sys, json
# usage: python tools/show_bench.py < bench.json   (or pass the file name)
src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for l in src:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('value', round(d['value'], 1), d['unit'], '| e2e', round(d['e2e']['value'], 1), '| ms/step', round(d['ms_per_step'], 2),
              '| roofline', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d['roofline'].items() if k in ('achieved', 'frac', 'tensor_pipe_frac', 'gemm_share_of_unet_time')})
        print('clocks', d.get('clocks'))
        if d.get('train'):
            print('train', json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d['train'].items() if k != 'what'}))
    elif l:
        print(l[:300])
