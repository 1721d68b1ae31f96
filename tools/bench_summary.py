"""Print the parts of a bench.py JSON line that matter when iterating on kernels.  python tools/bench_summary.py FILE"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches', 'n_gpus')})
print('e2e', {k: v for k, v in (d.get('e2e') or {}).items() if k != 'call'})
print('clocks', d.get('clocks'), '| cpu', d.get('cpu_baseline'))
print('gather', d.get('final_gather'))
r = d.get('roofline') or {}
print('roofline', {k: r.get(k) for k in ('kernel', 'achieved', 'peak', 'frac', 'traffic', 'share_of_step', 'avg_launch_us')})


def kl(lst, ind=''):
    for k in lst or []:
        extra = f" frac={k['frac']:.3f}" if 'frac' in k else ''
        if 'frac_incl_rhat_and_handover' in k:
            extra += f" (incl rhat+handover {k['frac_incl_rhat_and_handover']:.3f})"
        print(f"{ind}{k['share']:.3f} {k['avg_us']:8.1f} us x {k['launches_per_rollout']:7.1f}  {k['kernel']}{extra}")


kl(d.get('roofline_kernels'))
for n in ('motion_only', 'map_encoder', 'configs2', 'configs4_one_gpu'):
    e = d.get(n)
    if not e:
        continue
    print(n, {k: v for k, v in e.items() if k not in ('roofline_kernels', 'workload')})
    kl(e.get('roofline_kernels'), '    ')
