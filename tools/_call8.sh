cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 240 $T 29511 tools/tp_check.py > gpurun_out/r02_tp_check_n$N.txt 2>&1; echo "tp_check rc=$?"
grep -v "Warn\|warn" gpurun_out/r02_tp_check_n$N.txt | tail -4
timeout 420 $T 29512 bench.py --gpus $N --steps 2 --warmup 3 --no-beam > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
head -c 1500 gpurun_out/r02_bench_n$N.json; echo
timeout 420 $T 29513 bench.py --gpus $N --config c4 --steps 1 --warmup 3 > gpurun_out/r02_bench_c4_n$N.json 2> gpurun_out/r02_bench_c4_n$N.err; echo "c4 rc=$?"
head -c 3000 gpurun_out/r02_bench_c4_n$N.json; echo
timeout 300 $T 29514 bench.py --gpus $N --config c5 --steps 2 --warmup 3 > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err; echo "c5 rc=$?"
head -c 1500 gpurun_out/r02_bench_c5_n$N.json; echo
timeout 200 $T 29515 tools/tp_decode_bench.py 96 > gpurun_out/r02_tp_decode_n$N.txt 2>&1; echo "tp_decode rc=$?"
grep "decode step" gpurun_out/r02_tp_decode_n$N.txt
timeout 240 $T 29516 tools/wide_decode_bench.py 20 4100 10 > gpurun_out/r02_wide_decode_n$N.txt 2>&1; echo "wide rc=$?"
grep "wide decode" gpurun_out/r02_wide_decode_n$N.txt
tail -3 gpurun_out/r02_bench_n$N.err gpurun_out/r02_bench_c4_n$N.err | tail -12
