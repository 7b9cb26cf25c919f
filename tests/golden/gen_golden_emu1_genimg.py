"""Generates tests/golden/emu1_genimg_tiny.pt — `Emu.generate_image` of the UNMODIFIED reference
(Emu1/models/modeling_emu.py:187-249: n_causal full re-forwards, stu_regress_head fed back) on CPU, fp32, for the seeded
tiny weights of tests/helpers.emu1_state_dict: a text-only prompt and a prompt with one image.  Stores the token ids the
reference's own tokenizer produced for iteration 0 (prompt + "[IMG]") and the regressed embeddings.

Run in the authoring container (needs /root/reference):  python tests/golden/gen_golden_emu1_genimg.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import EMU1_LLAMA, EMU1_VIS, emu1_state_dict  # noqa: E402
from oracle import ref_shim  # noqa: E402

T5_TINY = dict(d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2)


def main():
    model = ref_shim.build_emu1_model(EMU1_VIS, EMU1_LLAMA, n_causal=8, t5_overrides=T5_TINY).float()
    # the constructor casts the LLaMA to bf16 (modeling_llama.py:173), which under transformers 5.x also rounds the rotary
    # `inv_freq` BUFFER; .float() does not bring those bits back.  Restore the exact fp32 frequencies for the fp32 golden.
    n_fix = 0
    for mod in model.modules():
        if hasattr(mod, "inv_freq") and torch.is_tensor(mod.inv_freq):
            dim = mod.inv_freq.numel() * 2
            inv = 1.0 / (EMU1_LLAMA["rope_theta"] ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
            mod.inv_freq.copy_(inv)
            if hasattr(mod, "original_inv_freq") and torch.is_tensor(mod.original_inv_freq):
                mod.original_inv_freq = inv.clone()
            n_fix += 1
    assert n_fix >= 1
    sd = emu1_state_dict(EMU1_VIS)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    ok_missing = ("embed_tokens", "visual.head", "visual.norm", "visual.fc_norm", "rotary_emb")
    assert not unexpected, unexpected
    assert all(any(t in k for t in ok_missing) and "lm.model.embed_tokens" not in k for k in missing), missing
    tok = model.decoder.tokenizer
    out = {}
    with torch.no_grad():
        text = ["An image of a dog."]
        ids = tok([text[0] + "[IMG]"], padding="longest", return_tensors="pt")
        out["text_only"] = {"input_ids": ids.input_ids, "attention_mask": ids.attention_mask,
                            "embeds": model.generate_image(text=text)}
        img = torch.randn(1, 3, 56, 56, generator=torch.Generator().manual_seed(4))
        text = ["[<IMG_PLH>]A similar picture."]
        full = text[0].replace("[<IMG_PLH>]", model.image_placeholder) + "[IMG]"
        ids = tok([full], padding="longest", return_tensors="pt")
        out["with_image"] = {"input_ids": ids.input_ids, "attention_mask": ids.attention_mask, "image": img,
                             "embeds": model.generate_image(text=text, image=img)}
    path = os.path.join(HERE, "emu1_genimg_tiny.pt")
    torch.save(out, path)
    print("wrote", path, {k: {kk: tuple(vv.shape) for kk, vv in v.items()} for k, v in out.items()})
    print(out["text_only"]["input_ids"], out["with_image"]["input_ids"])


if __name__ == "__main__":
    main()
