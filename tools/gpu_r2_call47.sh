#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c47; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py tests/test_gpu_precision_f32.py tests/test_golden.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py > $O/bench_cfg1.json 2> $O/bench_cfg1.err
tail -2 $O/tests.log; python -c "
import json; d=json.load(open('gpurun_out/r2c47/bench_cfg1.json')); print(round(d['value']), round(d['ms_per_step'],2), round(d['p50_step_ms'],2), round(d['p99_step_ms'],1), 'e2e', round(d['e2e']['value']), round(d['roofline']['frac'],3), {k:round(v.get('hbm_frac',0),2) for k,v in d['roofline_kernels'].items()}, 'cpu', round(d['cpu_baseline']['value']))"
