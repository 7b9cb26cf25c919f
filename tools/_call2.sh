cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 240 $T 29511 tools/tp_check.py > gpurun_out/r02_tp_check_n$N.txt 2>&1; echo "tp_check rc=$?"
grep -v "Warn\|warn\|\*\*\*" gpurun_out/r02_tp_check_n$N.txt | tail -6
timeout 300 $T 29516 tools/wide_decode_bench.py 20 4100 10 > gpurun_out/r02_wide_decode_n$N.txt 2>&1; echo "wide rc=$?"
grep "wide decode" gpurun_out/r02_wide_decode_n$N.txt
EMU_TP_P2P=0 timeout 300 $T 29517 tools/wide_decode_bench.py 20 4100 10 > gpurun_out/r02_wide_decode_n${N}_nccl.txt 2>&1; echo "wide nccl rc=$?"
grep "wide decode" gpurun_out/r02_wide_decode_n${N}_nccl.txt
timeout 420 $T 29513 bench.py --gpus $N --config c4 --steps 1 --warmup 3 > gpurun_out/r02_bench_c4_n$N.json 2> gpurun_out/r02_bench_c4_n$N.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c4_n2.json'))
print({k:d[k] for k in ('value','vit_ms','prefill_ms','decode_step_ms','tokens_sha1')}, d['roofline']['frac'], d['prefill_roofline']['frac'])
PY
