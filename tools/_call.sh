cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu > gpurun_out/r02h_gpu_tests.txt 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02h_gpu_tests.txt
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench_n1.json'))
print(d['value'], d['e2e']['value'], d['roofline']['decode_step_ms'], d['roofline']['frac'], d['vit_ms'], d['prefill_ms'])
print(d['beam5']['value'], d['denoise']['value'], d['denoise']['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
