"""Config dataclasses with the reference's field names and defaults (Emu2/emu/conf/emu_conf.py:7-40)."""
import json
import os.path as osp
from dataclasses import dataclass
from typing import Optional


@dataclass
class CLIPVisionCfg:
    """EVA-CLIP tower hyper-parameters: the reference's field names and defaults (Emu2/emu/conf/emu_conf.py:7-33) so that
    existing call sites and JSON configs keep working.  The engine reads image_size / patch_size / width / layers / head_width
    / mlp_ratio / postnorm / n_query / v_query; the training-only and variant switches are accepted and must keep their
    published values (checked in EmuModel.__init__)."""
    eva_model_name: str = "eva-clip-4b-14-x"
    image_size: int = 448
    patch_size: int = 14
    width: int = 1792
    layers: int = 64
    head_width: int = 112
    mlp_ratio: float = 8.571428571428571   # int(1792 * ratio) = 15360
    qkv_bias: bool = True                   # q / v bias, the k bias is zero
    drop_path_rate: float = 0.
    init_value: Optional[float] = None
    patch_dropout: float = 0.
    rope: bool = False
    global_average_pool: bool = False
    xattn: bool = False
    postnorm: bool = True                   # x + LN(f(x)) blocks, no final norm
    pt_hw_seq_len: int = 16
    intp_freq: bool = False
    naiveswiglu: bool = False
    subln: bool = False
    n_query: int = 64                       # pooled tokens per image
    v_query: int = 64                       # ... per video frame


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
