"""Config dataclasses with the reference's field names and defaults (Emu2/emu/conf/emu_conf.py:7-40)."""
import json
import os.path as osp
from dataclasses import dataclass, field, make_dataclass
from typing import Optional


# Field names and defaults are the reference's (Emu2/emu/conf/emu_conf.py:7-40) so that existing call sites and JSON configs keep
# working; only the fields marked "engine" influence this implementation, the others are accepted and must keep their
# published values (checked in EmuModel.__init__).
_VISION_FIELDS = [
    # (name, type, default)                                        # role
    ("eva_model_name", str, "eva-clip-4b-14-x"),                   # label only
    ("image_size", int, 448), ("patch_size", int, 14),             # engine: 32 x 32 patches + CLS
    ("width", int, 1792), ("layers", int, 64), ("head_width", int, 112),   # engine: 16 heads of 112
    ("mlp_ratio", float, 8.571428571428571),                       # engine: int(1792 * ratio) = 15360
    ("qkv_bias", bool, True),                                      # engine: q/v bias, k bias is zero
    ("drop_path_rate", float, 0.), ("init_value", Optional[float], None), ("patch_dropout", float, 0.),   # training only
    ("rope", bool, False), ("global_average_pool", bool, False), ("xattn", bool, False),
    ("postnorm", bool, True),                                      # engine: x + LN(f(x)) blocks, no final norm
    ("pt_hw_seq_len", int, 16), ("intp_freq", bool, False), ("naiveswiglu", bool, False), ("subln", bool, False),
    ("n_query", int, 64), ("v_query", int, 64),                    # engine: pooled tokens per image / video frame
]
CLIPVisionCfg = make_dataclass("CLIPVisionCfg", [(n, t, field(default=d)) for n, t, d in _VISION_FIELDS])
CLIPVisionCfg.__doc__ = "EVA-CLIP tower hyper-parameters (reference field names and defaults)."


@dataclass
class TextDecoderCfg:
    # directory holding config.json (+ tokenizer files when a tokenizer is not injected)
    llama_config_path: str = osp.join(osp.dirname(__file__), "llama_config")
    instruct: bool = False


# the published Emu2 decoder (Emu2/emu/conf/llama_config/config.json) — used when no config.json is given
EMU2_LLAMA_33B = dict(hidden_size=6656, num_hidden_layers=60, num_attention_heads=52, intermediate_size=17920,
                      rms_norm_eps=1e-6, max_position_embeddings=2048, vocab_size=32000, rope_theta=10000.0)


def load_llama_config(path_or_dict):
    if isinstance(path_or_dict, dict):
        cfg = dict(path_or_dict)
    else:
        p = path_or_dict
        if osp.isdir(p):
            p = osp.join(p, "config.json")
        cfg = json.load(open(p)) if osp.exists(p) else dict(EMU2_LLAMA_33B)
    cfg.setdefault("rope_theta", 10000.0)
    cfg.setdefault("rms_norm_eps", 1e-6)
    if cfg.get("num_key_value_heads", cfg["num_attention_heads"]) != cfg["num_attention_heads"]:
        raise ValueError("grouped-query attention is not part of the Emu decoder")
    return cfg
