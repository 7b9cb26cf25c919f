cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 240 $T 29516 tools/wide_decode_bench.py 20 4100 10 > gpurun_out/r02e_wide_decode_n$N.txt 2>&1; echo "wide rc=$?"
grep "wide decode" gpurun_out/r02e_wide_decode_n$N.txt
timeout 420 $T 29513 bench.py --gpus $N --config c4 --steps 1 --warmup 3 > gpurun_out/r02e_bench_c4_n$N.json 2> gpurun_out/r02e_bench_c4_n$N.err; echo "c4 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02e_bench_c4_n$N.json'))
print({k:d[k] for k in ('value','vit_ms','prefill_ms','decode_step_ms','tokens_sha1')}, d['roofline']['frac'], d['prefill_roofline']['frac'])
PY
timeout 200 $T 29511 tools/tp_check.py > gpurun_out/r02e_tp_check_n$N.txt 2>&1; echo "tp_check rc=$?"
grep "TP$N\|TP_CHECK" gpurun_out/r02e_tp_check_n$N.txt
