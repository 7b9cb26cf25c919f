"""CPU: host-side logic that mirrors the reference without touching the GPU — prompt assembly of the chat
pipeline, the Euler scheduler tables against the oracle restatement, config translation."""
import json
import os

import pytest
import torch

from oracle import diffusion_oracle as D


def test_euler_tables_match_oracle():
    from emu_b200.emu2.scheduler import EulerDiscreteScheduler
    s = EulerDiscreteScheduler()
    for n in (50, 20, 7):
        s.set_timesteps(n)
        ts, sig, init = D.euler_tables(n)
        assert torch.equal(s.timesteps, ts)
        assert torch.allclose(s.sigmas, sig, rtol=0, atol=0)
        assert abs(s.init_noise_sigma - init) < 1e-6
    s.set_timesteps(50)
    assert s.timesteps[0] == 981 and s.timesteps[-1] == 1 and s.sigmas[-1] == 0
    # external anchor (diffusers is not installable here): the scaled-linear 0.00085 -> 0.012 / 1000-step noise
    # schedule of Emu2/emu/conf/diffusion_config/scheduler/scheduler_config.json has the published Stable-Diffusion sigma
    # range sigma_min = 0.0292, sigma_max = 14.6146 (k-diffusion / SD v1 sampling configs quote exactly these)
    assert abs(float(s._sigmas_all[-1]) - 14.6146) < 1e-3
    assert abs(float(s._sigmas_all[0]) - 0.0292) < 1e-4


class _FakeModel:
    def __init__(self):
        self.calls = []

    def device(self):
        return torch.device("cpu")

    def dtype(self):
        return torch.float32

    def generate(self, **kw):
        self.calls.append(kw)
        return ["ok"]


def test_chat_prompt_assembly():
    from PIL import Image
    from emu_b200.emu2.chat import EmuChatGeneration
    from emu_b200.emu2 import constants as K
    fm = _FakeModel()
    pipe = EmuChatGeneration(fm)
    img = Image.new("RGB", (64, 48), (10, 200, 30))
    assert pipe.forward([img, "describe"], max_new_tokens=3) == "ok"
    kw = fm.calls[-1]
    assert kw["text"] == [K.DEFAULT_IMG_PLACEHOLDER + "describe"] and kw["image"].shape == (1, 3, 448, 448)
    assert kw["num_beams"] == 5 and kw["length_penalty"] == -1          # reference defaults
    pipe.forward([[img, "what is this?"], ["a cat"], ["and this?", img]], is_grounding=True)
    t = fm.calls[-1]["text"][0]
    assert t.startswith(K.GROUND_SYSTEM_MESSAGE + " [USER]: " + K.DEFAULT_IMG_PLACEHOLDER + "what is this?")
    assert " [ASSISTANT]: a cat</s>[USER]: and this?" in t and t.endswith(" [ASSISTANT]:" + K.GRD_SYMBOL)
    assert fm.calls[-1]["image"].shape[0] == 2


def test_image_transform_matches_reference_formula():
    """Resize(448, bicubic) -> ToTensor -> Normalize(OPENAI mean/std)  (Emu2/emu/chat.py:35-39)."""
    from PIL import Image
    import numpy as np
    from emu_b200.emu2.diffusion import image_transform
    from emu_b200.emu2.constants import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD
    rng = np.random.RandomState(0)
    img = Image.fromarray(rng.randint(0, 255, (40, 60, 3), dtype=np.uint8))
    x = image_transform(img)
    ref = np.asarray(img.resize((448, 448), resample=Image.BICUBIC), dtype=np.float32) / 255.0
    ref = (ref - np.array(OPENAI_DATASET_MEAN, dtype=np.float32)) / np.array(OPENAI_DATASET_STD, dtype=np.float32)
    assert x.shape == (3, 448, 448)
    assert np.allclose(x.permute(1, 2, 0).numpy(), ref, atol=1e-6)


def test_unet_config_translation_and_param_count():
    from emu_b200.emu2.diffusion import unet_config_from_json
    p = "/root/reference/Emu2/emu/conf/diffusion_config/unet/config.json"
    if os.path.exists(p):
        cfg = json.load(open(p))
    else:
        import bench
        cfg = bench.emu2_unet_json()
    u = unet_config_from_json(cfg)
    assert list(u.block_out_channels)[:3] == [320, 640, 1280] and list(u.transformer_layers)[:3] == [0, 2, 10]
    assert u.head_dim == 64 and u.cross_attention_dim == 1792 and u.projection_class_embeddings_input_dim == 3328
    n = sum(torch.Size(s).numel() for s in D.unet_param_shapes(D.EMU2_UNET).values())
    assert abs(n / 1e9 - 2.526) < 2e-3   # 2.526 B parameters from the reference's config
    import bench
    assert sum(torch.Size(s).numel() for _, s in bench.unet_param_shapes(bench.emu2_unet_json())) == n
