#!/bin/bash
# scheduling / compaction knobs of the QP solver on the headline workload (3 runs each, 30 timed steps)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2tune2; mkdir -p $O
run() { tag=$1; shift; for i in 1 2 3; do env "$@" timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --no-e2e $EXTRA > $O/${tag}_$i.json 2> $O/${tag}_$i.err; done; }
run g1d0 SMPC_QP_GROUPS=1
run g1d1 SMPC_QP_GROUPS=1 SMPC_QP_DEPTH=1
run g1d2 SMPC_QP_GROUPS=1 SMPC_QP_DEPTH=2
run g1d4 SMPC_QP_GROUPS=1 SMPC_QP_DEPTH=4
run g2d2 SMPC_QP_GROUPS=2 SMPC_QP_DEPTH=2
EXTRA="--config cfg2 --controller htwa" run htwa_g1 SMPC_QP_GROUPS=1
EXTRA="--config cfg2 --controller htwa" run htwa_g3 SMPC_QP_GROUPS=3
EXTRA="--config cfg2 --controller htwa" run htwa_g1d2 SMPC_QP_GROUPS=1 SMPC_QP_DEPTH=2
python - <<'PY'
import json,glob,collections
r=collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/r2tune2/*.json')):
    try: d=json.load(open(f))
    except Exception: continue
    r[f.split('/')[-1].rsplit('_',1)[0]].append((d['ms_per_step'], d['p50_step_ms'], d['p99_step_ms']))
for k,v in r.items(): print(k, ' '.join('%.2f/%.2f/%.1f'%t for t in v))
PY
