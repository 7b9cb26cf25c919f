cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 95 python -m pytest tests/test_generation_gpu.py -x -q > gpurun_out/r02i_generation_gpu.txt 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r02i_generation_gpu.txt
