"""Decode-step timing of the tensor-parallel LLaMA-33B (no ViT work): run under torchrun with N GPUs.
Prints the median CUDA-graphed decode step (ms), the implied HBM rate per GPU and the greedy token hash.  A/B switches:
EMU_TP_FOLD=0 (separate poll+reduce launch), EMU_TP_LL=0 (no fused push), EMU_TP_P2P=0 (NCCL)."""
import ctypes
import hashlib
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
    _, lc = bench.emu2_cfgs(False)
    vc = CLIPVisionCfg(image_size=56, width=128, layers=1, head_width=32, mlp_ratio=4.0, n_query=4)  # the ViT is not timed here
    vocab = synthetic.VOCAB_EMU2
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    model = EmuModel(vc, TextDecoderCfg(), tokenizer=synthetic.SyntheticTokenizer(vocab), llama_config=lc, max_batch=1,
                     max_seq=75 + steps + 8, tp_rank=rank, tp_size=world, nccl_uid=uid)
    synthetic.load_random_weights(model, vc, lc, vocab, seed=0)
    eng = model.engine
    g = torch.Generator().manual_seed(3)
    emb = (torch.randn(1, 75, lc["hidden_size"], generator=g) * 0.02).to(torch.bfloat16).cuda()
    mask = torch.ones(1, 75, dtype=torch.int32, device="cuda")
    for rep in range(2):
        eng.llm_reset()
        _, logits = eng.llm_prefill(emb, mask, hf_positions=True, want_logits=True)
        ping = [logits.argmax(-1).to(torch.int32), torch.empty(1, dtype=torch.int32, device="cuda")]
        toks = [ping[0].clone()]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs[0].record()
        for s in range(1, steps):
            eng.llm_decode(token_ids=ping[(s - 1) & 1], next_ids=ping[s & 1], ban_id=2, B=1)
            toks.append(ping[s & 1].clone())
            evs[s].record()
        torch.cuda.synchronize()
    ms = sorted(evs[s - 1].elapsed_time(evs[s]) for s in range(1, steps))
    med = ms[len(ms) // 2]
    ids = torch.cat(toks).cpu()
    alg = (bench.llm_bytes_per_token(lc, vocab) + bench.kv_bytes_per_ctx_token(lc) * (75 + steps / 2)) / world
    t = torch.tensor([med], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("TP%d decode step: median %.3f ms (max over ranks %.3f), p10 %.3f, p90 %.3f | %.0f GB/s per GPU | tok/s %.1f | "
              "ids sha1 %s | fold=%s ll=%s p2p=%s"
              % (world, med, float(t.item()), ms[len(ms) // 10], ms[len(ms) * 9 // 10], alg / (float(t.item()) / 1e3) / 1e9,
                 1000.0 / float(t.item()), hashlib.sha1(ids.numpy().tobytes()).hexdigest()[:12],
                 os.environ.get("EMU_TP_FOLD", "1"), os.environ.get("EMU_TP_LL", "1"), os.environ.get("EMU_TP_P2P", "1")), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
