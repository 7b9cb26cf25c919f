"""Generates tests/golden/emu1_tiny.pt — outputs of the UNMODIFIED reference Emu1 modules (EVAVisionTransformer
.forward_features -> ln_visual -> CausalFormer, Emu1/models/modeling_emu.py:125-126) for the seeded tiny weights of
tests/helpers.emu1_state_dict.  Weights are regenerated from the seed by the tests, only inputs/outputs are stored.

Run in the authoring container (needs /root/reference):  python tests/golden/gen_golden_emu1.py
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import EMU1_VIS, EMU1_VIS88, emu1_state_dict  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    CF = ref_shim.import_emu1_causal_former(dict(d_model=128, d_kv=64, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2))
    out = {}
    for vis in (EMU1_VIS, EMU1_VIS88):
        sd = emu1_state_dict(vis)
        img = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(3))
        vit = ref_shim.build_emu1_vit(vis).float()
        missing, unexpected = vit.load_state_dict({k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")},
                                                  strict=False)
        assert not unexpected and all(k.startswith(("head.", "norm.", "fc_norm")) for k in missing), (missing, unexpected)
        cf = CF(None, n_causal=8, vision_width=vis["width"], output_dim=256).eval().float()
        m2, u2 = cf.load_state_dict({k[len("cformer."):]: v for k, v in sd.items() if k.startswith("cformer.")}, strict=False)
        assert not u2 and all("embed_tokens" in k for k in m2), (m2, u2)
        with torch.no_grad():
            feats = vit.forward_features(img)
            feats = F.layer_norm(feats, (vis["width"],), sd["ln_visual.weight"], sd["ln_visual.bias"], 1e-6)
            out["w%d" % vis["width"]] = {"image": img, "ln_visual_features": feats, "cformer_out": cf(feats)}
    path = os.path.join(HERE, "emu1_tiny.pt")
    torch.save(out, path)
    print("wrote", path, {k: {kk: tuple(vv.shape) for kk, vv in v.items()} for k, v in out.items()})


if __name__ == "__main__":
    main()
