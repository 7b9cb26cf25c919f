"""bench.py's reference arm runs on the CPU: check the one-JSON-line contract (keys the driver reads) without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_emits_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines            # native-library banners must not reach stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "emu2_img2text_decode_tok_per_s" and d["unit"] == "tok/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # "reference" = the LLaMA part (prefill + decode, most of the CPU time) ran through transformers' own LlamaForCausalLM, the
    # class the reference instantiates; "port" = the oracle restatement (fallback when that API is not usable)
    assert (cb["kind"] == "reference") == ("LLaMA via transformers" in cb["sample"] and "failed" not in cb["sample"])
    assert d["e2e"] == {"value": d["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
