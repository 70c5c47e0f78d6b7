#!/bin/bash
# CTA size of the step kernels at the same 8 warps per SM: 2 / 4 / 8 warps per CTA
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c46; mkdir -p $O
run() { tag=$1; shift; for i in 1 2; do env "$@" timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --no-e2e --probe-steps 2 > $O/${tag}_$i.json 2> $O/${tag}_$i.err; done; }
run spw4 X=1
run spw2 SMPC_LIB=$PWD/build/variants/libspw2.so
run spw8 SMPC_LIB=$PWD/build/variants/libspw8.so
python - <<'PY'
import json,glob,collections
r=collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/r2c46/*.json')):
    try: d=json.load(open(f))
    except Exception: continue
    k=d['qp_solve']['kernel_ms']
    r[f.split('/')[-1].rsplit('_',1)[0]].append((d['ms_per_step'], k['qs_step0'], k['qs_step1']))
for k,v in r.items(): print(k, ' '.join('%.2f[%s %s]'%t for t in v))
PY
