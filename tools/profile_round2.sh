#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE GPU).  Outputs land in gpurun_out/; summaries are made locally by tools/ncu_summary.py.
set -x
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_hotpath.py -m gpu -q -k "copyto or fill or finalizer" 2>&1 | tail -3
timeout 300 python tools/perf_linalg.py > $O/r2_perf_linalg.txt 2>&1
# launch list of the bench step (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > $O/r2_bench_under_ncu.log 2>&1
# dominant kernel: DRAM traffic per launch of THIS build
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ew1_kernel -s 3 -c 1 -o $O/r2_ew1 python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-parity > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reduce_kernel -s 3 -c 1 -o $O/r2_reduce python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --no-parity > /dev/null 2>&1
# K12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 2 -c 1 -o $O/r2_gemm python tools/perf_gemm.py > /dev/null 2>&1
# 8 GiB chunk (north star size): bench line at log2n 31
timeout 600 python bench.py --steps 10 --warmup 3 --log2n 31 --no-extras > $O/r2_bench_log2n31.json 2> $O/r2_bench_log2n31.err
tail -c 1500 $O/r2_bench_log2n31.json
ls -la $O | tail -12
