#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c7; mkdir -p $O
timeout 600 python tools/f32_forms_probe.py > $O/f32_forms.log 2>&1
timeout 900 python -m pytest tests/test_gpu_precision_f32.py tests/test_gpu_kernel_variants.py -q -s -m gpu > $O/test_f32.log 2>&1; echo "f32+variants rc=$?" >> $O/summary.txt
cat $O/f32_forms.log | tail -8; grep -E "passed|failed|FAILED|rel err|identical" $O/test_f32.log | tail -20; cat $O/summary.txt
