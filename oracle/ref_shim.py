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


# ------------------------------------------------------------------------------------------------
# Emu1 Causal-Former (Emu1/models/causal_former.py + the vendored Emu1/models/modeling_t5.py)
# ------------------------------------------------------------------------------------------------
def import_emu1_causal_former(t5_overrides=None):
    """Import the UNMODIFIED reference CausalFormer under the installed transformers 5.x.  The vendored T5 was written
    for transformers 4.31; what 5.x removed is put back as inert shims — none of them touches arithmetic:
      * transformers.pytorch_utils.find_pruneable_heads_and_indices / prune_linear_layer (head pruning, never called)
      * transformers.utils.model_parallel_utils (device-map helpers, never called)
      * docstring decorators / DUMMY_* constants if missing
      * PreTrainedModel.get_head_mask (returns [None] * n_layers for head_mask=None, as 4.31 did)
      * PretrainedConfig.add_cross_attention = False (the 4.31 default the vendored T5Block reads)
      * T5Config.from_pretrained("t5-base") needs the network: replaced by the published t5-base config values.
    The module files are loaded by path into a synthetic package so Emu1/models/__init__.py (which pulls decord, peft,
    xformers-configured ViT …) is not executed.  `t5_overrides` (dict of T5Config fields) shrinks the stack for golden
    fixtures; None gives the real t5-base stack.  Returns the `CausalFormer` class."""
    import importlib.util

    import transformers  # noqa: F401
    import transformers.pytorch_utils as pu
    import transformers.utils as tu
    from transformers.modeling_utils import PreTrainedModel
    from transformers.models.t5.configuration_t5 import T5Config

    def _never(*a, **k):
        raise NotImplementedError("pruning / model-parallel helpers are not on the generate path")
    for name in ("find_pruneable_heads_and_indices", "prune_linear_layer"):
        if not hasattr(pu, name):
            setattr(pu, name, _never)
    for name in ("DUMMY_INPUTS", "DUMMY_MASK"):
        if not hasattr(tu, name):
            setattr(tu, name, [[0]])
    for name in ("add_start_docstrings", "add_start_docstrings_to_model_forward", "replace_return_docstrings"):
        if not hasattr(tu, name):
            setattr(tu, name, lambda *a, **k: (lambda f: f))
    if not hasattr(tu, "is_torch_fx_proxy"):
        tu.is_torch_fx_proxy = lambda x: False
    if "transformers.utils.model_parallel_utils" not in sys.modules:
        m = types.ModuleType("transformers.utils.model_parallel_utils")
        m.assert_device_map = lambda *a, **k: None
        m.get_device_map = lambda *a, **k: None
        sys.modules["transformers.utils.model_parallel_utils"] = m
    if not hasattr(PreTrainedModel, "get_head_mask"):
        def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
            if head_mask is not None:
                raise NotImplementedError
            return [None] * num_hidden_layers
        PreTrainedModel.get_head_mask = get_head_mask

    def t5_base_config(name, **kw):
        fields = dict(vocab_size=32128, d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_decoder_layers=12, num_heads=12,
                      relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.1,
                      layer_norm_epsilon=1e-6, feed_forward_proj="relu", is_encoder_decoder=True)
        fields.update(t5_overrides or {})
        c = T5Config(**fields)
        c.add_cross_attention = False
        return c
    T5Config.from_pretrained = staticmethod(t5_base_config)

    pkg_name = "_emu1_ref_models"
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "Emu1", "models")]
        sys.modules[pkg_name] = pkg
    mods = {}
    for name in ("modeling_t5", "causal_former"):
        full = pkg_name + "." + name
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(REFERENCE_ROOT, "Emu1", "models", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
        mods[name] = sys.modules[full]
    return mods["causal_former"].CausalFormer


def _load_emu1_module(name):
    """Load Emu1/models/<name>.py by path into the synthetic package (Emu1/models/__init__.py is not executed)."""
    import importlib.util
    pkg_name = "_emu1_ref_models"
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "Emu1", "models")]
        sys.modules[pkg_name] = pkg
    full = pkg_name + "." + name
    if full not in sys.modules:
        spec = importlib.util.spec_from_file_location(full, os.path.join(REFERENCE_ROOT, "Emu1", "models", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
    return sys.modules[full]


def import_emu1_vit():
    """The UNMODIFIED Emu1 `EVAVisionTransformer` (Emu1/models/eva_vit_model.py; pre-norm blocks, external ln_visual).
    Needs only the timm shim; `xattn` must stay False (xformers is not installed — same math)."""
    _install_timm_shim()
    _load_emu1_module("transformer")
    _load_emu1_module("rope")
    return _load_emu1_module("eva_vit_model").EVAVisionTransformer


def build_emu1_vit(vis, embed_dim=64):
    """Construct it the way Emu1/models/model.py:_build_vision_tower does for Emu-14B.json (with xattn off)."""
    from functools import partial
    EVA = import_emu1_vit()
    return EVA(img_size=vis["image_size"], patch_size=vis["patch_size"], num_classes=embed_dim, use_mean_pooling=False,
               init_values=None, patch_dropout=0., embed_dim=vis["width"], depth=vis["layers"],
               num_heads=vis["width"] // vis["head_width"], mlp_ratio=vis["mlp_ratio"], qkv_bias=True, drop_path_rate=0.,
               norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), xattn=False, rope=False, postnorm=False, pt_hw_seq_len=16,
               intp_freq=False, naiveswiglu=False, subln=False).eval()


def build_emu1_model(vis, llama, n_causal=8, t5_overrides=None):
    """Instantiate the UNMODIFIED reference `Emu` (Emu1/models/modeling_emu.py) on CPU with shrunk sub-models:
    `vis` -> vision_cfg (xattn off), `llama` -> a temporary ./models/llama_config (the class reads that cwd-relative
    path, Emu1/models/modeling_llama.py:5) holding the reference's own tokenizer files and a small config.json,
    `t5_overrides` -> the Causal-Former's T5 stack.  decord / peft are stubbed (only imported, never used here)."""
    import importlib
    _install_timm_shim()
    import_emu1_causal_former(t5_overrides)  # installs the transformers shims + the local T5 config
    for stub in ("decord", "peft"):
        if stub not in sys.modules:
            sys.modules[stub] = types.ModuleType(stub)
    if not hasattr(sys.modules["peft"], "PeftModel"):  # names prediction_mixin.py imports (LoRA eval helpers, unused here)
        for name in ("PeftModel", "LoraConfig", "TaskType", "get_peft_model"):
            setattr(sys.modules["peft"], name, type(name, (), {}))
    root = os.path.join(REFERENCE_ROOT, "Emu1")
    if root not in sys.path:
        sys.path.insert(0, root)
    work = tempfile.mkdtemp(prefix="emu1_ref_")
    dst = os.path.join(work, "models", "llama_config")
    os.makedirs(dst)
    src = os.path.join(root, "models", "llama_config")
    for f in os.listdir(src):
        if f != "config.json":
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    cfg = json.load(open(os.path.join(src, "config.json")))
    cfg.update(hidden_size=llama["hidden_size"], num_hidden_layers=llama["num_hidden_layers"],
               num_attention_heads=llama["num_attention_heads"], intermediate_size=llama["intermediate_size"],
               max_position_embeddings=llama.get("max_position_embeddings", 256), rms_norm_eps=llama.get("rms_norm_eps", 1e-6),
               torch_dtype="float32", attn_implementation="eager", _attn_implementation="eager")
    cfg.pop("num_key_value_heads", None)
    json.dump(cfg, open(os.path.join(dst, "config.json"), "w"))
    cwd = os.getcwd()
    os.chdir(work)
    try:
        me = importlib.import_module("models.modeling_emu")
        args = types.SimpleNamespace(instruct=False)
        vision_cfg = dict(image_size=vis["image_size"], layers=vis["layers"], width=vis["width"], head_width=vis["head_width"],
                          mlp_ratio=vis["mlp_ratio"], patch_size=vis["patch_size"], eva_model_name="eva-clip-g-14-x",
                          drop_path_rate=0, xattn=False, freeze=False)
        model = me.Emu(embed_dim=64, multimodal_cfg=dict(name="llama-13B", xattn=False, n_causal=n_causal, freeze=False),
                       vision_cfg=vision_cfg, vladapter_cfg=dict(name="cformer", n_causal=n_causal), args=args)
    finally:
        os.chdir(cwd)
    return model.eval()
