"""Print the judged keys of a bench.py JSON line (development aid). usage: print_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 5), 'fpl', d['config']['frames_per_launch'])
e = d['e2e']
print('e2e', round(e['value'], 1), 'via rank0', (e.get('via_rank0_gpu') or {}).get('value'), e.get('consumer_sees_every_ranks_frames'), 'single_blocking', e.get('single_frame_blocking'), 'd2h_floor', {k: v for k, v in (e.get('d2h_floor') or {}).items() if k != 'note'})
r = d['roofline']
print('roofline frac', r['frac'], 'achieved', r['achieved'], 'peak', r['peak'], 'dense', (d.get('roofline_dense') or {}).get('frac'))
print('timeline', {k: v for k, v in d['rank0_timeline_ms_per_step'].items() if k != 'note'}, 'timeouts', d.get('sortfirst_wait_timeouts'))
print('clocks', d.get('clocks'))
for k, v in (d.get('configs') or {}).items():
    print(k, {kk: v.get(kk) for kk in ('frames_per_s', 'ms_per_frame', 'busiest_rank_kernel_ms_per_frame', 'checks', 'phase_ms_mean_max_over_ranks', 'exchange_ms',
                                      'exchange_share_of_frame', 'error', 'not_run') if v.get(kk) is not None})
    if v.get('roofline'):
        print('   roofline', {kk: v['roofline'].get(kk) for kk in ('achieved', 'peak', 'frac')})
pc = d.get('parity_checks') or {}
print('parity', {k: (v.get('ok') if isinstance(v, dict) else v) for k, v in pc.items()})
for k in ('bonsai_standin', 'm0_reference_exact', 'cpu_baseline'):
    if d.get(k):
        print(k, json.dumps(d[k])[:400])
