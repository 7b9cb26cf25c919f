# what a verification call on the GPU box runs (gpurun -- 'bash tools/_call.sh'); outputs land in gpurun_out/
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.txt 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/gpu_tests.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
