#!/bin/bash
# One-shot validation used during development (run on the GPU box through gpurun): the full -m gpu suite, then the
# opt-in GEMM cluster mode (suite + micro-benchmarks + denoise loop), then one ncu capture of the attention kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest (default)"; timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== denoise (cluster off)"; UNET_GRAPH=1 timeout 200 python tools/ncu_unet.py 50 2>&1 | tail -1
echo "== gemm bench (cluster off)"; timeout 120 python tools/gemm_bench.py > gpurun_out/gemm_bench_cl0.txt 2>&1; tail -3 gpurun_out/gemm_bench_cl0.txt
echo "== ncu attn"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 1 -o gpurun_out/attn_tc_4096 -f python tools/ncu_attn.py 2 4096 10 64 > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
echo "== pytest gemm/conv (cluster on)"; EMU_GEMM_CLUSTER=1 timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "gemm or conv" 2>&1 | tail -4
echo "== gemm bench (cluster on)"; EMU_GEMM_CLUSTER=1 timeout 120 python tools/gemm_bench.py > gpurun_out/gemm_bench_cl1.txt 2>&1; tail -3 gpurun_out/gemm_bench_cl1.txt
echo "== pytest all (cluster on)"; EMU_GEMM_CLUSTER=1 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== denoise (cluster on)"; EMU_GEMM_CLUSTER=1 UNET_GRAPH=1 timeout 200 python tools/ncu_unet.py 50 2>&1 | tail -1
