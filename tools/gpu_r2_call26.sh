#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c26; mkdir -p $O
timeout 1500 python tools/cfg0_more_probe.py > $O/probe.log 2>&1
grep -v "^  File\|^    " $O/probe.log | tail -60
