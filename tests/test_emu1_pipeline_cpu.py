"""CPU: host logic of the Emu1 generation pipeline (emu_b200/emu1/scheduler.py): the PNDM / PLMS coefficient form the CUDA step
consumes reproduces the literal list-of-tensors scheduler of the oracle on random noise-prediction streams, and the UNet config
reader maps the SD-1.5 / SDXL head conventions correctly."""
import pytest
import torch

from oracle import diffusion_oracle as D


@pytest.mark.parametrize("steps", [3, 5, 20, 50])
def test_pndm_coefficients_match_literal_scheduler(steps):
    from emu_b200.emu1.scheduler import PNDMScheduler
    s = PNDMScheduler()
    s.set_timesteps(steps)
    ref = D.PNDMOracle()
    ref.set_timesteps(steps)
    assert torch.equal(s.timesteps, ref.timesteps) and len(s.timesteps) == steps + 1
    g = torch.Generator().manual_seed(steps)
    x_ref = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    x = x_ref.clone()
    hist = [torch.zeros_like(x) for _ in range(3)]
    saved = torch.zeros_like(x)
    ref.alphas_cumprod = ref.alphas_cumprod.double()
    ref.final_alpha_cumprod = ref.final_alpha_cumprod.double()
    for i, t in enumerate(s.timesteps.tolist()):
        e = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
        x_ref = ref.step(e, t, x_ref)
        a, b, wc, w0, w1, w2, _, flags = s.step_coefficients(i)   # what cfg_multistep_kernel evaluates
        flags = int(flags)
        m = wc * e + w0 * hist[0] + w1 * hist[1] + w2 * hist[2]
        src = saved if flags & 2 else x
        if flags & 4:
            saved = x.clone()
        if flags & 1:
            hist = [e, hist[0], hist[1]]
        x = a * src + b * m
        assert torch.allclose(x, x_ref, rtol=1e-6, atol=1e-6), (i, float((x - x_ref).abs().max()))


def test_unet_config_reader_head_conventions():
    from emu_b200.emu2.diffusion import unet_config_from_json
    sdxl = dict(in_channels=4, out_channels=4, block_out_channels=[320, 640, 1280], layers_per_block=2,
                transformer_layers_per_block=[1, 2, 10], attention_head_dim=[5, 10, 20], cross_attention_dim=1792,
                down_block_types=["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"], use_linear_projection=True,
                addition_time_embed_dim=256, projection_class_embeddings_input_dim=3328, norm_num_groups=32, norm_eps=1e-5)
    u = unet_config_from_json(sdxl)
    assert (u.head_dim, u.num_heads, u.mid_transformer_layers) == (64, 0, 10)
    assert list(u.transformer_layers)[:3] == [0, 2, 10]
    sd15 = dict(in_channels=4, out_channels=4, block_out_channels=[320, 640, 1280, 1280], layers_per_block=2,
                attention_head_dim=8, cross_attention_dim=5120, use_linear_projection=False,
                down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"], norm_num_groups=32, norm_eps=1e-5)
    u = unet_config_from_json(sd15)
    assert (u.head_dim, u.num_heads, u.mid_transformer_layers, u.use_linear_projection) == (0, 8, 1, 0)
    assert list(u.transformer_layers) == [1, 1, 1, 0] and u.addition_time_embed_dim == 0


def test_sd15_param_count():
    """the SD-1.5 topology with the stock 768-wide text cross-attention has 859.5 M parameters (published figure): pins the
    module tree of the oracle (and through the GPU tests, of emu_unet_configure) for the Emu1 decoder"""
    import math
    cfg = dict(D.EMU1_UNET, cross_attention_dim=768, mid_block_layers=1)
    n = sum(math.prod(s) for s in D.unet_param_shapes(cfg).values())
    assert n == 859_520_964, n          # the exact published parameter count of the SD-1.5 UNet


def test_sdxl_topology_param_count():
    """External anchor for the Emu2-Gen UNet restatement: Emu2's UNet is the SDXL-base topology with a 1792-wide context and a
    3328-wide text_time input (Emu2/emu/conf/diffusion_config/unet/config.json).  With SDXL-base's own two widths (2048 / 2816)
    the oracle's module tree must add up to SDXL-base's published 2,567,463,684 parameters exactly; with Emu2's it gives the
    2.526 B the survey computed from the reference's JSON."""
    import math
    sdxl = dict(D.EMU2_UNET, cross_attention_dim=2048, projection_class_embeddings_input_dim=2816)
    assert sum(math.prod(s) for s in D.unet_param_shapes(sdxl).values()) == 2_567_463_684
    assert sum(math.prod(s) for s in D.unet_param_shapes(D.EMU2_UNET).values()) == 2_525_520_644


def test_vae_decoder_param_count():
    """External anchor for the unpinned VAE restatement (diffusers is not installable here): the Stable-Diffusion
    AutoencoderKL of the reference's vae/config.json (block_out_channels 128/256/512/512, 2 layers per block, 4 latent
    channels) has a 49,490,179-parameter decoder — the published size of the SD / SDXL VAE decoder — plus the 20-parameter
    1x1 post_quant_conv.  The oracle's module tree (which the engine's weight loader mirrors key for key) must add up to it."""
    import torch
    from oracle import diffusion_oracle as D
    shapes = D.vae_decoder_param_shapes(D.EMU2_VAE)
    total = sum(torch.Size(s).numel() for s in shapes.values())
    pqc = sum(torch.Size(s).numel() for k, s in shapes.items() if k.startswith("post_quant_conv"))
    assert pqc == 20 and total - pqc == 49_490_179
