"""Generates tests/golden/emu1_cformer_tiny.pt — outputs of the UNMODIFIED reference Causal-Former
(Emu1/models/causal_former.py over the vendored modeling_t5.py, imported through oracle/ref_shim.py) for a shrunk T5
stack (2 layers, d_model 128, 2 heads, d_ff 256) and seeded weights.  The weights are NOT stored: tests regenerate them
with the same seed through oracle.diffusion_oracle.random_state_dict(oracle.t5_oracle.param_shapes(...)).

Run in the authoring container (needs /root/reference):  python tests/golden/gen_golden_cformer.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import diffusion_oracle as D, ref_shim, t5_oracle as T  # noqa: E402

CFG = dict(T.T5_BASE, layers=2, d_model=128, heads=2, d_ff=256)
ENC_W, OUT_DIM, N_CAUSAL, SEED = 96, 256, 8, 11


def seeded_state_dict():
    sd = D.random_state_dict(T.param_shapes(CFG, ENC_W, OUT_DIM, n_causal=N_CAUSAL), seed=SEED)
    for k in sd:  # T5 attention is unscaled: keep q small so the softmax is not saturated (Mesh-TF init does the same)
        if k.endswith("Attention.q.weight"):
            sd[k] = sd[k] * 0.125
    return sd


def main():
    CF = ref_shim.import_emu1_causal_former(dict(d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2))
    sd = seeded_state_dict()
    out = {"cfg": CFG, "enc_w": ENC_W, "out_dim": OUT_DIM, "n_causal": N_CAUSAL, "seed": SEED}
    x = torch.randn(2, 10, ENC_W, generator=torch.Generator().manual_seed(SEED + 1))
    out["img_embeds"] = x
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        # the constructor creates the T5 blocks in bf16 (causal_former.py:33-34): cast FIRST so that loading the seeded
        # fp32 weights is not silently rounded in the fp32 run
        m = CF(None, n_causal=N_CAUSAL, vision_width=ENC_W, output_dim=OUT_DIM).eval().to(dt)
        missing, unexpected = m.load_state_dict({k[len("cformer."):]: v for k, v in sd.items()}, strict=False)
        assert not unexpected and all("embed_tokens" in k for k in missing), (missing, unexpected)
        with torch.no_grad():
            out["out_" + name] = m(x.to(dt)).float()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu1_cformer_tiny.pt")
    torch.save(out, path)
    print("wrote", path, {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
