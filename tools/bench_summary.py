"""Scratch: prints the interesting numbers of bench.py JSON lines (gpurun_out/*.log)."""
import json, sys
for f in sys.argv[1:]:
    txt = open(f).read().strip().split('\n')
    try:
        d = json.loads(txt[-1])
    except Exception:
        print(f, 'FAILED'); print('\n'.join(txt[-15:])); continue
    print('==', f, 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'ms/step %.1f' % d['ms_per_step'],
          'qp/step/inst %.2f' % d.get('qp_per_mpc_step_per_instance', 0), 'status', d.get('bnb_status_counts_rank0'), 'clocks', d.get('clocks', {}).get('sm_mhz'))
    if 'roofline' in d:
        r = d['roofline']
        print('  roofline frac %.4f achieved %.3f TF hbm frac %.5f lanes %s' % (r['frac'], r['achieved'], r['hbm']['frac'], r['solver_lanes_per_sm']))
    if 'from_fresh_states' in d:
        print('  fresh', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d['from_fresh_states'].items() if k not in ('note', 'one_launch')}, 'one launch: %.0f' % d['from_fresh_states'].get('one_launch', {}).get('value', 0))
    for k in ('roofline_k2k4', 'cold_start', 'per_mpc_step_api', 'single_instance_nominal', 'cpu_baseline'):
        if d.get(k):
            print('  ', k, {kk: (round(v, 4) if isinstance(v, float) else v) for kk, v in d[k].items() if kk not in ('note', 'sample')})
    for k, v in d.get('published_protocol', {}).items():
        print('   pub', k, v)
