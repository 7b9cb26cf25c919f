"""Per-phase timing of the persistent decode-step kernel (CTA 0 globaltimer stamps).  Run on the GPU box."""
import collections
import ctypes
import os
import sys

os.environ["EMU_MEGA_PROF"] = "1"
os.environ["EMU_MEGA"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from emu_b200 import _lib  # noqa: E402
from emu_b200.emu2 import synthetic  # noqa: E402
from emu_b200.emu2.conf import TextDecoderCfg  # noqa: E402
from emu_b200.emu2.emu import EmuModel  # noqa: E402


def main():
    vc, lc = bench.emu2_cfgs(False)
    vc.layers = 1  # the ViT is irrelevant here; keep start-up short
    model = EmuModel(vc, TextDecoderCfg(), tokenizer=synthetic.SyntheticTokenizer(), llama_config=lc, max_batch=1, max_seq=256)
    synthetic.load_random_weights(model, vc, lc, synthetic.VOCAB_EMU2)
    ids, mask = synthetic.image_prompt_ids(n_query=64, n_text=8)
    emb = model.engine.llm_embed(ids.cuda())
    model.engine.llm_reset()
    _, lg = model.engine.llm_prefill(emb, mask.cuda(), want_logits=True)
    ping = [lg.argmax(-1).to(torch.int32), torch.empty(1, dtype=torch.int32, device="cuda")]
    for s in range(1, 20):
        model.engine.llm_decode(token_ids=ping[(s - 1) & 1], next_ids=ping[s & 1], ban_id=2, B=1)
    torch.cuda.synchronize()
    n = 4096
    stamps = (ctypes.c_ulonglong * (2 * n))()
    types = (ctypes.c_int * n)()
    lib = _lib.load()
    k = lib.emu_debug_mega_profile(model.engine.h, stamps, types, n)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    names = {0: "gemv:o/down/lm_head", 2: "gemv:gate_up", 16: "gemv:qkv", 100: "attention", 200: "embed", 300: "norm_out", 400: "final"}
    prev_end = None
    for i in range(k):
        t0, t1 = stamps[2 * i], stamps[2 * i + 1]
        a = agg[names.get(types[i], str(types[i]))]
        a[0] += 1
        a[1] += (t1 - t0) / 1000.0
        if prev_end is not None:
            a[2] += (t0 - prev_end) / 1000.0  # barrier wait before the phase
        prev_end = t1
    tot = (stamps[2 * k - 1] - stamps[0]) / 1000.0
    print("phases %d, step total %.1f us" % (k, tot))
    for name, (cnt, busy, wait) in agg.items():
        print("  %-22s n=%4d  in-phase %8.1f us (avg %6.2f)   barrier-before %8.1f us (avg %5.2f)" % (name, cnt, busy, busy / cnt, wait, wait / cnt))


if __name__ == "__main__":
    main()
