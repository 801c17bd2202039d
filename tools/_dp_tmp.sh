timeout 300 python -m pytest tests/test_tc_gpu.py -x -q -k "attention" 2>&1 | tail -3
S4_BENCH_B=16 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"delta_peel|dq_convert" -c 12 python tools/bench_ops.py attn_bwd 4 2>&1 | grep -E "attn_bwd |gpu__time_duration|delta_peel|dq_convert" | awk '/kernel/{n=$1} /gpu__time/{print n, $NF}' | sort | uniq -c | sort -rn | head -6
S4_BENCH_B=16 timeout 100 python tools/bench_ops.py attn_bwd 8
