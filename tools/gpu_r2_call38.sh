#!/bin/bash
# final build on 8 B200 (torchrun, one rank per GPU): configs[2] at 100 000 problems (HTWA, receding) and the headline configuration per rank
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c38; mkdir -p $O
for c in htwa receding; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 \
     --config cfg2 --controller $c --no-mlp > $O/cfg2_${c}_8gpu.out 2> $O/cfg2_${c}_8gpu.err
  echo "$c rc=$?"; grep '^{' $O/cfg2_${c}_8gpu.out | tail -1 > $O/cfg2_${c}_8gpu.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 --no-mlp > $O/cfg1_8gpu.out 2> $O/cfg1_8gpu.err
echo "cfg1 rc=$?"; grep '^{' $O/cfg1_8gpu.out | tail -1 > $O/cfg1_8gpu.json
python - <<'PY'
import json
for c in ('cfg2_htwa','cfg2_receding','cfg1'):
    try: d=json.load(open(f'gpurun_out/r2c38/{c}_8gpu.json'))
    except Exception as e: print(c,'failed',e); continue
    print(c, 'n_gpus', d['n_gpus'], 'global batch', d['config']['global_batch'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],2), 'p99', round(d['p99_step_ms'],1))
PY
