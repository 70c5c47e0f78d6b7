#!/bin/bash
# run-time knobs re-checked on the final build (cfg[1], 2 runs x 30 steps each)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c45; mkdir -p $O
run() { tag=$1; shift; for i in 1 2; do env "$@" timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --no-e2e --probe-steps 2 > $O/${tag}_$i.json 2> $O/${tag}_$i.err; done; }
run default X=1
run redo_off SMPC_QP_REDO_LIST=0
run redo_quarter SMPC_QP_REDO_LIST=2500
run compact70 SMPC_QP_COMPACT_AT=0.70
run compact92 SMPC_QP_COMPACT_AT=0.92
run tail768 SMPC_QP_TAIL=768
run tail192 SMPC_QP_TAIL=192
python - <<'PY'
import json,glob,collections
r=collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/r2c45/*.json')):
    try: d=json.load(open(f))
    except Exception: continue
    r[f.split('/')[-1].rsplit('_',1)[0]].append((d['ms_per_step'], d['p50_step_ms'], d['p99_step_ms']))
for k,v in r.items(): print(k, ' '.join('%.2f/%.2f/%.1f'%t for t in v))
PY
