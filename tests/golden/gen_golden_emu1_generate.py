"""Generates tests/golden/emu1_generate_tiny.pt — BASELINE configs[0] in miniature: the UNMODIFIED reference `Emu.generate`
(Emu1/models/modeling_emu.py:100-185, the call Emu1/inference.py:66-80 makes: image -> EVA ViT -> ln_visual -> Causal-Former ->
spliced into the left-padded prompt -> lm.generate) on CPU for the seeded tiny weights of tests/helpers.emu1_state_dict, in the
dtype the reference itself runs (the whole model in bf16 under bf16 autocast: `generate` casts the image to bf16,
modeling_emu.py:124,150).  Stores the images, the reference tokenizer's ids for the prompts and the new token ids that
`lm.generate` returned (captured at the tokenizer's batch_decode), greedy and 3-beam, two prompts of different length.

Run in the authoring container (needs /root/reference):  python tests/golden/gen_golden_emu1_generate.py
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
N_CAUSAL = 8
IMG = "[IMG]" + "<image>" * N_CAUSAL + "[/IMG]"                      # Emu1/inference.py:9 with this model's n_causal
PROMPTS = [IMG + "a picture of", "Look: " + IMG + "What is shown in this image? Answer:"]


def main():
    model = ref_shim.build_emu1_model(EMU1_VIS, EMU1_LLAMA, n_causal=N_CAUSAL, t5_overrides=T5_TINY)
    sd = emu1_state_dict(EMU1_VIS)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    model = model.to(torch.bfloat16).eval()                              # as inference.py runs it
    model.args.device = torch.device("cpu")
    tok = model.decoder.tokenizer
    captured = []
    decode = tok.batch_decode
    tok.batch_decode = lambda ids, **kw: (captured.append(ids.clone()), decode(ids, **kw))[1]
    image = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(21))
    tok.padding_side = "left"
    enc = tok(PROMPTS, padding="longest", return_tensors="pt", add_special_tokens=True)
    tok.padding_side = "right"
    out = {"image": image, "input_ids": enc.input_ids, "attention_mask": enc.attention_mask, "prompts": PROMPTS}
    with torch.no_grad():
        for name, kw in (("greedy", dict(num_beams=1)), ("beam3", dict(num_beams=3, length_penalty=0.0)),
                         ("beam3_lp1_ret2", dict(num_beams=3, length_penalty=1.0, num_captions=2))):
            text = model.generate({"image": image, "prompt": PROMPTS}, max_new_tokens=12, **kw)
            out["ids_" + name], out["text_" + name] = captured[-1], text
            print(name, captured[-1].tolist(), text)
    path = os.path.join(HERE, "emu1_generate_tiny.pt")
    torch.save(out, path)
    print("wrote", path, tuple(enc.input_ids.shape))


if __name__ == "__main__":
    main()
