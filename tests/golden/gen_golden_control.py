"""Generate tests/golden/emu2_tiny_control.pt: token ids the UNMODIFIED reference decoder (`model.decoder.lm.generate`, driven
exactly as Emu2/emu/emu.py:213-229 drives it) produces for the generation-control knobs beyond plain greedy / beam search —
logits processors under greedy and beam search, num_return_sequences (Emu1's num_captions), prefix_allowed_tokens_fn
(Emu1/mm_eval/models/emu.py:97-109) and beam-sample (do_sample with the default num_beams > 1, what the chat demo runs).

Run in the authoring container only:  python tests/golden/gen_golden_control.py
Same tiny model, weights, image and prompts as gen_golden.py (emu2_tiny.pt); sampling cases record the torch seed.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import TINY_LLAMA, TINY_VISION, make_emu2_state_dict  # noqa: E402
from oracle import ref_shim  # noqa: E402

ALLOWED = list(range(100, 140)) + [2]      # the token set of the prefix-constrained cases


def prefix_fn(batch_id, ids):
    """a VizWiz-style constraint: after 3 tokens only EOS may follow"""
    return [2] if len(ids) >= 3 else ALLOWED


SMALL = [6554, 18722, 15312, 29412, 3895, 1741]   # a 6-token vocabulary: repeats are certain, so the repetition processors bite


def small_fn(batch_id, ids):
    return SMALL + ([2] if len(ids) >= 8 else [])


CASES = {
    "greedy_small": dict(B=2, kw=dict(num_beams=1, prefix_allowed_tokens_fn=small_fn)),
    "greedy_small_rep": dict(B=2, kw=dict(num_beams=1, prefix_allowed_tokens_fn=small_fn, repetition_penalty=1.5)),
    "greedy_small_ngram2": dict(B=2, kw=dict(num_beams=1, prefix_allowed_tokens_fn=small_fn, no_repeat_ngram_size=2)),
    "beam3_small": dict(B=2, kw=dict(num_beams=3, prefix_allowed_tokens_fn=small_fn)),
    "beam3_small_rep_ngram": dict(B=2, kw=dict(num_beams=3, prefix_allowed_tokens_fn=small_fn, repetition_penalty=1.4,
                                               no_repeat_ngram_size=2)),
    "greedy_rep": dict(B=2, kw=dict(num_beams=1, repetition_penalty=1.5)),
    "greedy_ngram2": dict(B=2, kw=dict(num_beams=1, no_repeat_ngram_size=2)),
    "greedy_rep_ngram": dict(B=2, kw=dict(num_beams=1, repetition_penalty=1.3, no_repeat_ngram_size=3)),
    "greedy_prefix": dict(B=2, kw=dict(num_beams=1, prefix_allowed_tokens_fn=prefix_fn)),
    "beam3_rep_ngram": dict(B=2, kw=dict(num_beams=3, length_penalty=1.0, repetition_penalty=1.4, no_repeat_ngram_size=2)),
    "beam4_ret3": dict(B=2, kw=dict(num_beams=4, length_penalty=-1, num_return_sequences=3)),
    "beam3_prefix": dict(B=2, kw=dict(num_beams=3, length_penalty=0.0, prefix_allowed_tokens_fn=prefix_fn)),
    "beamsample3": dict(B=2, seed=77, kw=dict(num_beams=3, do_sample=True, temperature=0.8, top_p=0.9, length_penalty=1.0)),
    "beamsample2_topk": dict(B=1, seed=5, kw=dict(num_beams=2, do_sample=True, top_k=8, length_penalty=-1)),
    # Emu2/demo/frontend/libs/chat_frontend.py:178-184 defaults with "Do Sample" ticked
    "beamsample5_demo": dict(B=1, seed=9, kw=dict(num_beams=5, do_sample=True, top_k=3, top_p=0.9, temperature=0.7,
                                                  length_penalty=1.0)),
}


def main():
    torch.manual_seed(0)
    d = ref_shim.make_llama_config_dir(TINY_LLAMA["hidden_size"], TINY_LLAMA["num_hidden_layers"],
                                       TINY_LLAMA["num_attention_heads"], TINY_LLAMA["intermediate_size"],
                                       max_pos=TINY_LLAMA["max_position_embeddings"])
    vk = dict(TINY_VISION)
    vk.pop("patch_size")
    model = ref_shim.build_emu2_model(dict(vk, patch_size=14), d)
    model.load_state_dict(make_emu2_state_dict(), strict=True)
    base = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu2_tiny.pt"))
    ids, mask, image = base["gen_input_ids"], base["gen_attention_mask"], base["image"]
    out = {}
    with torch.no_grad():
        emb = model.decoder.lm.model.embed_tokens(ids)
        pe = model.project_up(model.encode_image(image).view(-1, TINY_VISION["width"]))
        emb[ids == 32003] = pe
        from emu.emu import GENERATION_CONFIG
        for name, case in CASES.items():
            B = case["B"]
            if "seed" in case:
                torch.manual_seed(case["seed"])
            # every knob EmuModel.generate forwards (Emu2/emu/emu.py:213-229), at the reference's defaults: top_k / top_p /
            # temperature / penalty_alpha are passed as None, which switches the GenerationConfig defaults (top_k = 50) OFF
            kw = dict(length_penalty=1.0, repetition_penalty=1.0, penalty_alpha=None, top_k=None, top_p=None, temperature=None)
            kw.update(case["kw"])
            seq = model.decoder.lm.generate(generation_config=GENERATION_CONFIG, inputs_embeds=emb[:B],
                                            attention_mask=mask[:B], do_sample=kw.pop("do_sample", False),
                                            max_new_tokens=12, min_length=1, **kw)
            out[name] = seq
            print(name, tuple(seq.shape), seq.tolist())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu2_tiny_control.pt")
    torch.save(out, path)
    print("wrote", path)


if __name__ == "__main__":
    main()
