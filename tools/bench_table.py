"""profiles/r02_bench/*.json -> the markdown table of profiles/r02_bench_configs.md (one row per bench line)."""
import glob, json, os, sys
d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'profiles', 'r02_bench')
print('| file | workload | GPUs × B | N | ms per step (p50 / p99) | RTI iterations/s | e2e (host buffers) | IPM it./solve | kernels launched per step | dominant QP kernel: fraction of the HBM peak |')
print('|---|---|---|---|---|---|---|---|---|---|')
for f in sorted(glob.glob(os.path.join(d, '*.json'))):
    try:
        j = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])     # (a torchrun line may be preceded by NCCL's banner)
    except Exception as e:
        print(f'| {os.path.basename(f)} | unreadable: {e} |'); continue
    name = os.path.basename(f)
    if j.get('impl') == 'reference':
        cb = j.get('cpu_baseline', {})
        print(f"| {name} | CPU arm ({cb.get('kind')}, {cb.get('cores')} host threads): {cb.get('sample')} | | | | {j['value']:.0f} | | | | |")
        continue
    c = j['config']
    w = c['workload']
    ctrl = w.split('controller=')[1].split(',')[0]
    N = w.split('N=')[1].split(',')[0]
    steps = j['steps']
    r = j.get('roofline') or {}
    print(f"| {name} | controller={ctrl}, {j['dtype']}, nn={c.get('nn_precision')} | {j['n_gpus']} × {c['batch_per_gpu']} | {N} | {j['ms_per_step']:.2f} ({j['p50_step_ms']:.1f} / {j['p99_step_ms']:.1f}) | "
          f"{j['value']:.0f} | {j['e2e']['value']:.0f} | {j['ipm_iterations_per_solve']:.1f} | {j['gpu_launches'] / steps / j['n_gpus']:.0f} | "
          f"{(r.get('kernel') or '').split(' /')[0]}: {r.get('frac', 0):.2f} |" if 'e2e' in j else
          f"| {name} | controller={ctrl}, {j['dtype']} | {j['n_gpus']} × {c['batch_per_gpu']} | {N} | {j['ms_per_step']:.2f} ({j['p50_step_ms']:.1f} / {j['p99_step_ms']:.1f}) | {j['value']:.0f} | | {j['ipm_iterations_per_solve']:.1f} | {j['gpu_launches'] / steps / j['n_gpus']:.0f} | |")
