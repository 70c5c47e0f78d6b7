#!/bin/bash
# round 2, GPU call 1: import probe of the reference stack, GPU test suite (new BASELINE-config tests first), bench cfg1, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt
for m in acados_template casadi adam l4casadi pinocchio urdf_parser_py; do python -c "import $m; print('$m', 'OK', getattr($m,'__version__',''))" 2>&1 | tail -1; done > $O/import_probe.txt
timeout 1500 python -m pytest tests/test_gpu_baseline_configs.py -x -q -s -m gpu > $O/test_baseline.log 2>&1; echo "baseline tests rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_baseline_configs.py > $O/test_rest.log 2>&1; echo "rest tests rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?" >> $O/summary.txt
# executed FP64 instruction tally of the linearisation kernel (exact flop count of the implementation)
timeout 600 ncu --kernel-name linearize_kernel -c 2 --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size --clock-control none --csv --log-file $O/linearize_flops.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-mlp > $O/ncu1.log 2>&1
# full capture of one all-active qs_step<0> launch and of the linearisation kernel
timeout 900 ncu --kernel-name regex:"qs_step_kernel|linearize_kernel" --launch-skip 6 -c 3 --set full --import-source on --clock-control none -o $O/step_lin_full python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-mlp > $O/ncu2.log 2>&1
ncu -i $O/step_lin_full.ncu-rep --page raw --csv > $O/step_lin_full_raw.csv 2>/dev/null
tail -5 $O/test_baseline.log; tail -3 $O/test_rest.log; cat $O/summary.txt; cat $O/import_probe.txt; head -c 600 $O/bench_cfg1.json
