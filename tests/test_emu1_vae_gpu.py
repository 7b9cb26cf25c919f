"""GPU parity for the Emu1-specific blocks (pre-norm EVA ViT + ln_visual, Causal-Former, stu_regress_head path) and the
VAE decoder, against the CPU oracles (oracle/emu_oracle.py, oracle/t5_oracle.py, oracle/diffusion_oracle.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import StubTokenizer, make_emu2_state_dict
from oracle import diffusion_oracle as D
from oracle import emu_oracle as O
from oracle import t5_oracle as T

pytestmark = pytest.mark.gpu

VIS = dict(image_size=56, patch_size=14, width=128, layers=2, head_width=32, mlp_ratio=4.0)  # head_dim 32
VIS88 = dict(image_size=56, patch_size=14, width=176, layers=2, head_width=88, mlp_ratio=4.0)  # head_dim 88 like EVA-g
LLAMA = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, rms_norm_eps=1e-6,
             max_position_embeddings=256, vocab_size=32000, rope_theta=10000.0)
T5 = dict(T.T5_BASE, layers=2, d_model=128, heads=2, d_ff=256)


def emu1_state_dict(vis, seed=0):
    sd = make_emu2_state_dict(vision=dict(vis, n_query=4, v_query=4), llama=LLAMA, vocab=32004, seed=seed)
    sd.pop("project_up.weight"), sd.pop("project_down.weight")
    g = torch.Generator().manual_seed(seed + 100)
    W = vis["width"]
    sd["ln_visual.weight"] = 1 + 0.1 * torch.randn(W, generator=g)
    sd["ln_visual.bias"] = 0.05 * torch.randn(W, generator=g)
    sd["decoder.lm.stu_regress_head.weight"] = torch.randn(256, 256, generator=g) / 16
    cf = D.random_state_dict(T.param_shapes(T5, W, 256, n_causal=8), seed=seed + 7)
    for k in cf:  # T5 attention is unscaled: the Mesh-TF init keeps q small so that the softmax is not saturated
        if k.endswith("Attention.q.weight"):
            cf[k] = cf[k] * 0.125
    sd.update(cf)
    return sd


class Tok(StubTokenizer):
    def __len__(self):
        return 32004


def build(vis, sd):
    from emu_b200.emu1.modeling_emu import Emu
    m = Emu(vision_cfg=vis, vladapter_cfg={"n_causal": 8}, tokenizer=Tok(), llama_config=LLAMA,
            cformer_cfg=dict(layers=2, d_model=128, heads=2, d_ff=256), max_batch=4, max_seq=64)
    m.load_state_dict(sd)
    return m


@pytest.mark.parametrize("vis", [VIS, VIS88])
def test_emu1_vit_lnvisual_cformer(cuda, vis):
    sd = emu1_state_dict(vis)
    m = build(vis, sd)
    img = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(3))
    heads = vis["width"] // vis["head_width"]
    feats = O.vit_forward_features(sd, img, patch=14, num_heads=heads, layers=2, postnorm=False)
    feats = F.layer_norm(feats, (vis["width"],), sd["ln_visual.weight"], sd["ln_visual.bias"], 1e-6)
    got = m.engine.vit_forward(img.cuda(), 0, pool=False).float().cpu()
    assert O.rel_err(got, feats) < 3e-2
    ref = T.causal_former(sd, feats, T5)
    out = m.encode_image(img.cuda()).float().cpu()
    assert O.rel_err(out, ref) < 3e-2


def test_emu1_generate_image(cuda):
    sd = emu1_state_dict(VIS)
    m = build(VIS, sd)
    ids = torch.tensor([[1, 500, 600, 700, 32001]])
    mask = torch.ones_like(ids)
    out = m.generate_image_from_ids(ids, mask).float().cpu()
    # oracle: cache-less literal loop of Emu1/models/modeling_emu.py:205-243 (regressed embeds fed back directly)
    emb = F.embedding(ids, sd["decoder.lm.model.embed_tokens.weight"])
    outs = []
    for k in range(8):
        h = O.llama_forward(sd, emb, torch.ones(1, emb.shape[1], dtype=torch.long), layers=2, heads=2)
        reg = F.linear(h[:, -1], sd["decoder.lm.stu_regress_head.weight"])
        outs.append(reg)
        emb = torch.cat((emb, reg[:, None]), dim=1)
    ref = torch.stack(outs, dim=1)
    assert O.rel_err(out, ref) < 3e-2


def test_vae_decode(cuda):
    from emu_b200 import _lib
    cfg = dict(latent_channels=4, out_channels=3, block_out_channels=(32, 64, 64), layers_per_block=1, norm_num_groups=32)
    sd = D.random_state_dict(D.vae_decoder_param_shapes(cfg), seed=5)
    eng = _lib.Engine(_lib.EmuConfig())
    v = _lib.EmuVAEConfig()
    v.latent_channels, v.out_channels, v.n_blocks = 4, 3, 3
    for i, c in enumerate(cfg["block_out_channels"]):
        v.block_out_channels[i] = c
    v.layers_per_block, v.norm_groups = 1, 32
    eng.vae_configure(v)
    eng.load_state_dict({"vae." + k: t for k, t in sd.items()})
    z = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(6))
    ref = (D.vae_decode(sd, cfg, z.bfloat16().float()) / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1)
    out = eng.vae_decode(z.cuda()).cpu()
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) < 3e-2   # image values live in [0, 1]
