cd /root/repo; export PYTHONPATH=$PWD S4_BENCH_B=24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 6 -c 8 -f -o gpurun_out/prof_outproj python tools/bench_ops.py gemm 1 > gpurun_out/ncu_outproj.log 2>&1
tail -1 gpurun_out/ncu_outproj.log | cut -c1-100
