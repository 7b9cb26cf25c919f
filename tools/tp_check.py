"""Tensor-parallel parity check: run under torchrun with N GPUs.

Every rank builds the tiny Emu2 LLM with tp_size = WORLD_SIZE (NCCL communicator created inside libemu_b200.so from a
unique id broadcast over torch.distributed), runs prefill + 4 decode steps, and rank 0 compares logits with the fp32
CPU oracle.  The head count never divides the rank count (3 heads on 2 ranks = 2/1, 5 on 4 = 2/1/1/1, 13 on 8 =
2/2/2/2/2/1/1/1 — the same uneven split as Emu2's 52 heads on 8 GPUs = 7/7/7/7/6/6/6/6) to exercise the zero-weight head
slot, and the vocabulary is the Emu2-Chat one (32274: not divisible by 4 or 8) to exercise the padded lm_head shard.
"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import StubTokenizer, make_emu2_state_dict  # noqa: E402
from emu_b200 import _lib  # noqa: E402
from oracle import emu_oracle as O  # noqa: E402

VIS = dict(image_size=56, patch_size=14, width=128, layers=1, head_width=32, mlp_ratio=4.0, n_query=4, v_query=4)
HEADS = {1: 3, 2: 3, 4: 5, 8: 13}
VOCAB_CHAT = 32274   # 32000 + [PAD] + 271 specials + [USER] + [ASSISTANT]  (Emu2/emu/lm.py:63-65, instruct=True)


def llama_cfg(world):
    h = HEADS.get(world, 3)
    return dict(hidden_size=128 * h, num_hidden_layers=2, num_attention_heads=h, intermediate_size=1024, rms_norm_eps=1e-6,
                max_position_embeddings=256, vocab_size=32000, rope_theta=10000.0)


class ChatTokenizer(StubTokenizer):
    def __len__(self):
        return VOCAB_CHAT


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().emu_nccl_unique_id(raw))
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())

    from emu_b200.emu2.conf import CLIPVisionCfg, TextDecoderCfg
    from emu_b200.emu2.emu import EmuModel
    LLAMA = llama_cfg(world)
    sd = make_emu2_state_dict(vision=VIS, llama=LLAMA, vocab=VOCAB_CHAT)
    m = EmuModel(CLIPVisionCfg(**VIS), TextDecoderCfg(), tokenizer=ChatTokenizer(), llama_config=LLAMA, max_batch=2,
                 max_seq=64, tp_rank=rank, tp_size=world, nccl_uid=uid)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(100, 30000, (2, 9), generator=g)
    mask = torch.ones_like(ids)
    mask[1, :3] = 0
    emb = torch.nn.functional.embedding(ids, sd["decoder.lm.model.embed_tokens.weight"])
    toks, logit_list = O.generate_greedy(sd, emb, mask, layers=2, heads=LLAMA["num_attention_heads"], max_new_tokens=5, min_len=5, return_logits=True)
    e_emb = m.engine.llm_embed(ids.cuda())
    m.engine.llm_reset()
    _, lg = m.engine.llm_prefill(e_emb, mask.cuda(), hf_positions=True, want_logits=True)
    errs = [O.rel_err(lg.cpu(), logit_list[0])]
    out = torch.empty_like(lg)
    for s in range(1, len(logit_list)):
        m.engine.llm_decode(token_ids=toks[:, s - 1].to(torch.int32).cuda().contiguous(), logits=out, B=2)
        errs.append(O.rel_err(out.cpu(), logit_list[s]))
        # every rank must hold bitwise identical logits (fixed-order reduction), or greedy tokens could diverge
        ref0 = out.clone()
        dist.broadcast(ref0, 0)
        if not torch.equal(ref0, out):
            errs.append(1e9)
    worst = torch.tensor([max(errs)], device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("TP%d logits rel err per step: %s" % (world, ["%.2e" % e for e in errs]), flush=True)
    ok = float(worst) < 3e-2
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)
    if rank == 0:
        print("TP_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
