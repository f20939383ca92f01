import glob, json, sys
for f in sorted(glob.glob('/root/repo/gpurun_out/diag_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ph = {k: v['ms'] for k, v in d['phases_rank0'].items()}
        print('%-28s %10.0f views/s  %.3f ms/step  phases %.3f  cull %.3f  e2e %.0f' % (f.split('diag_')[1][:-5], d['value'], d['ms_per_step'], sum(ph.values()), ph['cull'], d['e2e']['value']))
    except Exception as e:
        print(f, 'ERR', e)
