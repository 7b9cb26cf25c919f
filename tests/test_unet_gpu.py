"""Parity of the CUDA UNet / denoise step (through the C ABI) against oracle/diffusion_oracle.py on a tiny
SDXL-topology configuration with seeded random weights.  The diffusion oracle restates diffusers 0.24.0
("parity unpinned": diffusers is neither vendored nor installed)."""
import pytest
import torch

from oracle import diffusion_oracle as D
from oracle import emu_oracle as O

pytestmark = pytest.mark.gpu

TINY_UNET = dict(in_channels=4, out_channels=4, block_out_channels=(64, 128, 256), layers_per_block=2,
                 transformer_layers_per_block=(0, 1, 2), attention_head_dim=64, cross_attention_dim=128,
                 addition_time_embed_dim=32, projection_class_embeddings_input_dim=128 + 6 * 32, norm_num_groups=32,
                 norm_eps=1e-5)


def make_engine(cfg, sd):
    from emu_b200 import _lib
    eng = _lib.Engine(_lib.EmuConfig())
    u = _lib.EmuUNetConfig()
    u.in_channels, u.out_channels = cfg["in_channels"], cfg["out_channels"]
    u.n_blocks = len(cfg["block_out_channels"])
    for i, c in enumerate(cfg["block_out_channels"]):
        u.block_out_channels[i] = c
        u.transformer_layers[i] = cfg["transformer_layers_per_block"][i]
    u.layers_per_block = cfg["layers_per_block"]
    u.head_dim, u.cross_attention_dim = cfg.get("attention_head_dim", 0), cfg["cross_attention_dim"]
    u.num_heads = cfg.get("num_heads", 0)
    u.mid_transformer_layers = cfg.get("mid_block_layers", 0)
    u.use_linear_projection = 1 if cfg.get("use_linear_projection", True) else 0
    u.addition_time_embed_dim = cfg.get("addition_time_embed_dim") or 0
    u.projection_class_embeddings_input_dim = cfg.get("projection_class_embeddings_input_dim") or 0
    u.norm_groups, u.norm_eps = cfg["norm_num_groups"], cfg["norm_eps"]
    eng.unet_configure(u)
    eng.load_state_dict({"unet." + k: v for k, v in sd.items()})
    return eng


@pytest.fixture(scope="module")
def setup(cuda):
    sd = D.random_state_dict(D.unet_param_shapes(TINY_UNET), seed=3)
    eng = make_engine(TINY_UNET, sd)
    g = torch.Generator().manual_seed(11)
    ctx = torch.randn(2, 8, 128, generator=g)
    tid = torch.tensor([[1024, 1024, 0, 0, 1024, 1024]] * 2)
    return sd, eng, ctx, tid


@pytest.mark.parametrize("hw", [(32, 32), (16, 64)])
def test_unet_forward(setup, hw):
    sd, eng, ctx, tid = setup
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 4, *hw, generator=g)
    te = ctx.mean(1)
    bsd = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    ref32 = D.unet_forward(sd, TINY_UNET, x.bfloat16().float(), 981.0, ctx.bfloat16().float(), te.bfloat16().float(), tid)
    ref16 = D.unet_forward(bsd, TINY_UNET, x.bfloat16(), 981.0, ctx.bfloat16(), te.bfloat16(), tid).float()
    out = eng.unet_forward(x.cuda(), 981.0, ctx.cuda(), te.cuda(), tid.cuda()).float().cpu()
    from helpers import assert_bf16_parity
    assert_bf16_parity("tiny unet forward %dx%d" % hw, out, ref32, ref16)


def test_denoise_loop(setup):
    """5 CFG + Euler steps (graph-captured from the 2nd step on) vs the oracle loop in fp32."""
    sd, eng, ctx, tid = setup
    steps, guidance = 5, 3.0
    ts, sig, init_sigma = D.euler_tables(steps)
    g = torch.Generator().manual_seed(13)
    lat0 = torch.randn(1, 4, 32, 32, generator=g) * init_sigma
    te = ctx.mean(1)
    ref = D.denoise_loop(lambda x, t, c, e, i: D.unet_forward(sd, TINY_UNET, x, t, c, e, i), lat0.clone(), ctx, te, tid,
                         steps, guidance)
    bsd = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    ref16 = D.denoise_loop(lambda x, t, c, e, i: D.unet_forward(bsd, TINY_UNET, x.to(torch.bfloat16), t, c.to(torch.bfloat16),
                                                                 e.to(torch.bfloat16), i).float(),
                           lat0.clone(), ctx, te, tid, steps, guidance)
    lat = lat0.clone().cuda().contiguous()
    ctx_d, te_d, tid_d = ctx.to(torch.bfloat16).cuda(), te.to(torch.bfloat16).cuda(), tid.to(torch.int32).cuda()
    for i in range(steps):
        eng.denoise_step(lat, float(sig[i]), float(sig[i + 1]), float(ts[i]), guidance, ctx_d, te_d, tid_d)
    from helpers import assert_bf16_parity
    assert_bf16_parity("tiny denoise loop (5 Euler steps)", lat.cpu(), ref, ref16)


def test_denoise_graph_matches_eager(setup):
    import os
    sd, eng, ctx, tid = setup
    ts, sig, init_sigma = D.euler_tables(4)
    lat0 = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(14)) * init_sigma
    te = ctx.mean(1)
    ctx_d, te_d, tid_d = ctx.to(torch.bfloat16).cuda(), te.to(torch.bfloat16).cuda(), tid.to(torch.int32).cuda()
    outs = []
    lat = lat0.clone().cuda().contiguous()
    for mode in ("0", "1"):
        os.environ["EMU_NO_GRAPH"] = mode
        lat.copy_(lat0)
        for i in range(4):
            eng.denoise_step(lat, float(sig[i]), float(sig[i + 1]), float(ts[i]), 3.0, ctx_d, te_d, tid_d)
        outs.append(lat.clone())
    os.environ.pop("EMU_NO_GRAPH", None)
    assert torch.equal(outs[0], outs[1])


# ------------------------------------------------------------------------------------------------------------------
# Emu1 visual decoder: Stable-Diffusion-1.5 topology at its real widths (320 / 640 / 1280 / 1280, 8 heads per level = head
# widths 40 / 80 / 160, 1x1-conv projections, mid-block attention after a plain last down block, 32 context tokens of width
# 5120, no added conditioning), one resnet per block to keep the CPU oracle quick; PNDM / PLMS loop
# ------------------------------------------------------------------------------------------------------------------
SD15 = dict(D.EMU1_UNET, layers_per_block=1, mid_block_layers=1)


@pytest.fixture(scope="module")
def sd15(cuda):
    sd = D.random_state_dict(D.unet_param_shapes(SD15), seed=5)
    eng = make_engine(SD15, sd)
    g = torch.Generator().manual_seed(15)
    ctx = torch.randn(2, 32, 5120, generator=g).to(torch.bfloat16)
    return sd, eng, ctx


def test_sd15_unet_forward(sd15):
    sd, eng, ctx = sd15
    g = torch.Generator().manual_seed(16)
    x = torch.randn(2, 4, 32, 32, generator=g).to(torch.bfloat16)
    bsd = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    with torch.no_grad():
        ref32 = D.unet_forward(sd, SD15, x.float(), 961.0, ctx.float())
        ref16 = D.unet_forward(bsd, SD15, x, 961.0, ctx).float()
    out = eng.unet_forward(x.cuda(), 961.0, ctx.cuda()).float().cpu()
    e_eng, e_bf = O.rel_err(out, ref32), O.rel_err(ref16, ref32)
    print("\n[sd15] unet forward: engine-vs-fp32 %.3e | bf16-oracle-vs-fp32 %.3e | ratio %.2f" % (e_eng, e_bf, e_eng / e_bf))
    assert e_eng <= 1.5 * e_bf, (e_eng, e_bf)


def test_sd15_pndm_loop(sd15):
    """4 PLMS steps (5 UNet evaluations, graph replays from the 3rd on) under CFG 7.5 vs the literal oracle loop"""
    from emu_b200.emu1.scheduler import PNDMScheduler
    sd, eng, ctx = sd15
    steps, guidance = 4, 7.5
    g = torch.Generator().manual_seed(17)
    lat0 = torch.randn(1, 4, 16, 16, generator=g).to(torch.bfloat16).float()
    bsd = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    with torch.no_grad():
        ref32 = D.pndm_denoise_loop(lambda x, t, c: D.unet_forward(sd, SD15, x, t, c), lat0.clone(), ctx.float(), steps, guidance)
        ref16 = D.pndm_denoise_loop(lambda x, t, c: D.unet_forward(bsd, SD15, x.to(torch.bfloat16), t, c).float(), lat0.clone(),
                                    ctx, steps, guidance)
    sch = PNDMScheduler()
    sch.set_timesteps(steps)
    lat = lat0.clone().cuda().contiguous()
    state = torch.zeros(4, *lat.shape, dtype=torch.float32, device="cuda")
    ctx_d = ctx.cuda()
    for i, t in enumerate(sch.timesteps.tolist()):
        eng.denoise_step_multistep(lat, state, sch.step_coefficients(i), float(t), guidance, ctx_d)
    e_eng, e_bf = O.rel_err(lat.cpu(), ref32), O.rel_err(ref16, ref32)
    print("\n[sd15] PNDM loop: engine-vs-fp32 %.3e | bf16-oracle-vs-fp32 %.3e | ratio %.2f" % (e_eng, e_bf, e_eng / e_bf))
    assert e_eng <= max(1.5 * e_bf, 2e-3), (e_eng, e_bf)
