cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "skinny or beam_topk or beam_step" > gpurun_out/r02c_skinny_tests.txt 2>&1; echo "skinny tests rc=$?"
tail -15 gpurun_out/r02c_skinny_tests.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_gpu_tests.txt 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02c_gpu_tests.txt
timeout 400 python tools/wide_decode_bench.py 20 1024 12 > gpurun_out/r02c_wide_decode.txt 2>&1; echo rc=$?
EMU_WIDE_SKINNY=0 timeout 400 python tools/wide_decode_bench.py 20 1024 12 >> gpurun_out/r02c_wide_decode.txt 2>&1; echo rc=$?
grep -v Warning gpurun_out/r02c_wide_decode.txt | tail -4
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'gemm_tc|gemm_skinny|attn_decode|rope_kv|rmsnorm|kv_indir|logits|embed' -c 5000 --csv --log-file gpurun_out/r02c_wide_launches.csv python tools/wide_decode_bench.py 20 1024 1 > gpurun_out/r02c_wide_ncu.log 2>&1; echo rc=$?
