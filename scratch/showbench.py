import json, sys
for f in sys.argv[1:]:
    try:
        t=open(f).read().strip().splitlines()[-1]
        d=json.loads(t)
    except Exception as e:
        print(f, 'ERR', e); continue
    r=d.get('roofline') or {}
    e=d.get('e2e') or {}
    c=d.get('cpu_baseline') or {}
    print(f.split('/')[-1], 'n_gpus', d.get('n_gpus'), 'value %.1fM'%(d['value']/1e6), 'ms %.3f'%d['ms_per_step'], 'e2e %s'%('%.1fM'%(e['value']/1e6) if e.get('value') else None), 'e2e_ms', e.get('ms_per_step'), 'launches', d.get('gpu_launches'), 'graphs/s %.0f'%(d.get('training_graphs_per_s') or 0))
    print('   roof:', r.get('kernel'), 'frac %.3f'%(r.get('frac') or 0), 'traffic', r.get('traffic'), 'alg', r.get('algorithmic_bytes_per_launch'), 'iter_frac', (r.get('fixed_point_iteration') or {}).get('frac_of_hbm_peak'))
    print('   cpu:', c.get('value'), c.get('cores'), (c.get('sample') or '')[:80], '| clocks', d.get('clocks'))
