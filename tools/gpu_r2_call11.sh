#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c11; mkdir -p $O
for v in lin3 lin4; do SMPC_LIB=$PWD/build/variants/lib$v.so timeout 300 python tools/lin_probe.py 10000 2>&1 | tail -3 | sed "s/^/$v /" >> $O/lin_variants.log; done
cat $O/lin_variants.log
