#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c49; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rk4_sens.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"
tail -3 $O/tests.log
timeout 300 python tools/rk4_bench.py > $O/rk4_bench.json 2> $O/rk4_bench.err; cat $O/rk4_bench.json; tail -3 $O/rk4_bench.err
timeout 300 ncu --kernel-name regex:rk4_sens_kernel -c 2 --clock-control none --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum --csv --log-file $O/rk4_flops.csv python tools/rk4_bench.py 460000 > /dev/null 2>&1
cut -c1-400 $O/rk4_flops.csv | tail -30
