"""Wide (more than 8 cache rows) decode-step timing of LLaMA-33B: the C4 shape (4 prompts x 5 beams = 20 rows) on 1..N GPUs.
usage: [torchrun ...] tools/wide_decode_bench.py [rows=20] [ctx=1024] [steps=12]
Prints the median CUDA-graphed step with and without a per-step beam re-parent, and the HBM rate the step implies
(weights once + the KV of rows x ctx tokens).  EMU_KV_COPY=1 moves the cache on each re-parent (HF behaviour) for comparison."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from emu_b200 import _lib  # noqa: E402
from emu_b200.emu2 import synthetic  # noqa: E402
from emu_b200.emu2.conf import CLIPVisionCfg, TextDecoderCfg  # noqa: E402
from emu_b200.emu2.emu import EmuModel  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = ctypes.create_string_buffer(128)
            _lib.check(_lib.load().emu_nccl_unique_id(raw))
            buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    nb = 5 if rows % 5 == 0 else 1
    _, lc = bench.emu2_cfgs(False)
    vc = CLIPVisionCfg(image_size=56, width=128, layers=1, head_width=32, mlp_ratio=4.0, n_query=4)  # the ViT is not timed here
    vocab = synthetic.VOCAB_EMU2
    model = EmuModel(vc, TextDecoderCfg(), tokenizer=synthetic.SyntheticTokenizer(vocab), llama_config=lc, max_batch=rows,
                     max_seq=ctx + 2 * steps + 8, tp_rank=rank, tp_size=world, nccl_uid=uid)
    synthetic.load_random_weights(model, vc, lc, vocab, seed=0)
    eng = model.engine
    g = torch.Generator().manual_seed(3)
    Bp = rows // nb
    emb = (torch.randn(Bp, ctx, lc["hidden_size"], generator=g) * 0.02).to(torch.bfloat16).cuda()
    mask = torch.ones(Bp, ctx, dtype=torch.int32, device="cuda")
    eng.llm_reset()
    eng.llm_prefill(emb, mask, hf_positions=True, want_logits=True)
    eng.llm_expand(torch.arange(rows, dtype=torch.int32, device="cuda") // nb, rows)
    tok = torch.randint(100, 30000, (rows,), generator=g).to(torch.int32).cuda()
    logits = torch.empty(rows, vocab, dtype=torch.float32, device="cuda")
    src = ((torch.arange(rows) // nb) * nb + torch.randint(0, nb, (rows,), generator=g)).to(torch.int32).cuda()
    out = {}
    for name, bs in (("plain", None), ("reparent", src)):
        for _ in range(2):
            eng.llm_decode(token_ids=tok, logits=logits, B=rows, beam_src=bs)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs[0].record()
        for s in range(steps):
            eng.llm_decode(token_ids=tok, logits=logits, B=rows, beam_src=bs)
            evs[s + 1].record()
        torch.cuda.synchronize()
        ms = sorted(evs[s].elapsed_time(evs[s + 1]) for s in range(steps))
        t = torch.tensor([ms[len(ms) // 2]], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = float(t.item())
    alg = (bench.llm_bytes_per_token(lc, vocab) + bench.kv_bytes_per_ctx_token(lc) * rows * (ctx + steps * 2)) / world
    if rank == 0:
        print("wide decode TP%d rows=%d ctx=%d: step %.3f ms plain, %.3f ms with a beam re-parent | %.0f GB/s per GPU (plain) | "
              "kv_copy=%s" % (world, rows, ctx, out["plain"], out["reparent"], alg / (out["plain"] / 1e3) / 1e9,
                              os.environ.get("EMU_KV_COPY", "0")), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
