"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (baaivision/Emu, /root/reference) on CPU so the oracle
restatement (oracle/emu_oracle.py) can be pinned against the reference's own forward, and golden fixtures can be
generated (tests/golden/gen_golden.py).  /root/reference only exists in the authoring container; nothing that runs
on the GPU box imports this module.

Shims: the reference imports `timm.models.layers.{drop_path,to_2tuple}` (Emu2/emu/eva_vit.py:13-16)
and, for Emu1, `trunc_normal_`; timm is not installed, so an in-memory module provides the three helpers.
"""
import collections.abc
import itertools
import json
import os
import shutil
import sys
import tempfile
import types

import torch

REFERENCE_ROOT = os.environ.get("EMU_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Emu2", "emu"))


def _install_timm_shim():
    import transformers  # noqa: F401  (must be imported BEFORE the shim: transformers probes timm.__spec__)
    import transformers.models.llama.modeling_llama  # noqa: F401
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_emu_shim", False):
        return

    def to_2tuple(x):
        if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
            return tuple(x)
        return tuple(itertools.repeat(x, 2))

    def drop_path(x, drop_prob: float = 0.0, training: bool = False, scale_by_keep: bool = True):
        if drop_prob == 0.0 or not training:
            return x
        keep_prob = 1 - drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        random_tensor = x.new_empty(shape).bernoulli_(keep_prob)
        if keep_prob > 0.0 and scale_by_keep:
            random_tensor.div_(keep_prob)
        return x * random_tensor

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    import importlib.machinery
    timm = types.ModuleType("timm")
    timm._emu_shim = True
    timm.__spec__ = importlib.machinery.ModuleSpec("timm", None)
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers2 = types.ModuleType("timm.layers")
    for m in (layers, layers2):
        m.to_2tuple = to_2tuple
        m.drop_path = drop_path
        m.trunc_normal_ = trunc_normal_
    timm.models = models
    models.layers = layers
    timm.layers = layers2
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers, "timm.layers": layers2})


def make_llama_config_dir(hidden, layers, heads, ffn, max_pos=2048, rms_eps=1e-6, dst=None):
    """A shrunk copy of the reference's llama_config dir (its tokenizer files + a small config.json)."""
    src = os.path.join(REFERENCE_ROOT, "Emu2", "emu", "conf", "llama_config")
    dst = dst or tempfile.mkdtemp(prefix="emu_llama_cfg_")
    for f in ("tokenizer.model", "tokenizer_config.json", "special_tokens_map.json", "tokenizer.json",
              "generation_config.json"):
        if os.path.exists(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    cfg = json.load(open(os.path.join(src, "config.json")))
    cfg.update(hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads, intermediate_size=ffn,
               max_position_embeddings=max_pos, rms_norm_eps=rms_eps, torch_dtype="float32",
               attn_implementation="eager", _attn_implementation="eager")
    cfg.pop("num_key_value_heads", None)
    json.dump(cfg, open(os.path.join(dst, "config.json"), "w"))
    return dst


def import_emu2():
    """Return the reference package `emu` (Emu2/emu) imported from /root/reference."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    _install_timm_shim()
    p = os.path.join(REFERENCE_ROOT, "Emu2")
    if p not in sys.path:
        sys.path.insert(0, p)
    import emu.emu  # noqa: F401
    import emu.conf.emu_conf  # noqa: F401
    return sys.modules["emu"]


def build_emu2_model(vision_kwargs, llama_dir, instruct=False, seed=0, dtype=torch.float32):
    """Instantiate the reference EmuModel with random-init weights (seeded), eval mode."""
    pkg = import_emu2()
    from emu.conf.emu_conf import CLIPVisionCfg, TextDecoderCfg
    from emu.emu import EmuModel
    torch.manual_seed(seed)
    model = EmuModel(vision_cfg=CLIPVisionCfg(**vision_kwargs),
                     text_decoder_cfg=TextDecoderCfg(llama_config_path=llama_dir, instruct=instruct))
    # the reference leaves cls_token / pos_embed / q_bias / v_bias / norms at zeros/ones; randomise so tests bite
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1 or "cls_token" in n or "pos_embed" in n:
                if "norm" in n and n.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
    model.eval()
    try:
        model.decoder.lm.config._attn_implementation = "eager"
    except Exception:
        pass
    return model.to(dtype)
