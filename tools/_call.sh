cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02d_gpu_tests.txt 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02d_gpu_tests.txt
timeout 400 python tools/wide_decode_bench.py 20 2048 10 > gpurun_out/r02d_wide_decode.txt 2>&1; echo rc=$?
grep "wide decode" gpurun_out/r02d_wide_decode.txt
timeout 400 python tools/wide_decode_bench.py 20 1024 10 >> gpurun_out/r02d_wide_decode.txt 2>&1; echo rc=$?
grep "wide decode" gpurun_out/r02d_wide_decode.txt | tail -1
timeout 300 python tools/skinny_bench.py 20 1,2,8 > gpurun_out/r02d_skinny_bench.txt 2>&1; echo rc=$?
grep -v Warn gpurun_out/r02d_skinny_bench.txt
