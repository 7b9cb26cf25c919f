"""Streaming checkpoint ingestion: the reference formats go tensor by tensor from disk into the engine's
packed / tensor-parallel device buffers — the whole state dict is never materialised in host memory.

Formats (the files the reference's loaders read):
  * Emu2 single file  `Emu2-*_pytorch_model.bf16.safetensors` / `.bin`      Emu2/emu/chat.py:129-149 (`from_pretrained`)
  * HF sharded dir    `pytorch_model.bin.index.json` / `model.safetensors.index.json` + shards
                                                                            Emu2/emu/conf/llama_config/pytorch_model.bin.index.json
  * Emu2-Gen dir      `multimodal_encoder/`, `unet/diffusion_pytorch_model.safetensors`, `vae/…`   Emu2/emu/diffusion.py:251-318
  * Emu1              `ckpt['module']` (DeepSpeed-style wrapper) with an optional LoRA adapter merged on the fly
                                                                            Emu1/inference.py:40-57

`iter_checkpoint(path)` yields (key, tensor) lazily; `load_into(sink, path, …)` feeds any object with a
`load_tensor(key, tensor)` method (the C-ABI `emu_engine_load_tensor` behind `_lib.Engine.load_tensor`), which does the
repacking (fused QKV, interleaved gate/up, conv layouts) and keeps only this rank's tensor-parallel shard on the device.
"""
import json
import os
import os.path as osp
from typing import Callable, Dict, Iterator, Optional, Tuple

import torch


def _iter_safetensors(path: str) -> Iterator[Tuple[str, torch.Tensor]]:
    from safetensors import safe_open
    with safe_open(path, framework="pt", device="cpu") as f:
        for k in f.keys():
            yield k, f.get_tensor(k)  # one tensor resident at a time (the file is memory-mapped)


def _load_torch_bin(path: str, allow_pickle: bool = False):
    """torch.load restricted to tensors (weights_only=True).  Legacy (non-zip) pickles cannot be memory-mapped, so the
    retry drops mmap — it never widens the unpickler.  A checkpoint that needs arbitrary pickled objects loads only
    when the caller opts in with allow_pickle=True (the reference's plain torch.load, Emu1/inference.py:44, trusts
    the file; a drop-in must not do that silently)."""
    try:
        sd = torch.load(path, map_location="cpu", mmap=True, weights_only=True)
    except Exception:
        try:
            sd = torch.load(path, map_location="cpu", mmap=False, weights_only=True)
        except Exception:
            if not allow_pickle:
                raise
            sd = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(sd, dict) and "module" in sd and isinstance(sd["module"], dict):
        sd = sd["module"]  # Emu1: torch.load(ckpt)['module']  (Emu1/inference.py:44-45)
    return sd


def _iter_torch_bin(path: str, allow_pickle: bool = False) -> Iterator[Tuple[str, torch.Tensor]]:
    sd = _load_torch_bin(path, allow_pickle)
    for k in list(sd.keys()):
        yield k, sd.pop(k)


def _files_of(path: str):
    if not osp.isdir(path):
        return [path]
    for index in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
        ip = osp.join(path, index)
        if osp.exists(ip):
            shards = sorted(set(json.load(open(ip))["weight_map"].values()))
            missing = [s for s in shards if not osp.exists(osp.join(path, s))]
            if missing:  # the reference loads with strict=True (Emu2/emu/chat.py:212): an absent shard is an error
                raise FileNotFoundError("checkpoint shards listed in %s are missing: %s" % (index, missing[:3]))
            return [osp.join(path, s) for s in shards]
    files = sorted(f for f in os.listdir(path) if f.endswith((".safetensors", ".bin", ".pt", ".pth")))
    if not files:
        raise FileNotFoundError("no checkpoint files under %s" % path)
    return [osp.join(path, f) for f in files]


def iter_checkpoint(path: str, allow_pickle: bool = False) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (key, cpu tensor) pairs of a file, a sharded directory (index.json) or a directory of weight files."""
    for f in _files_of(path):
        if f.endswith(".safetensors"):
            yield from _iter_safetensors(f)
        else:
            yield from _iter_torch_bin(f, allow_pickle)


def iter_keys(path: str, allow_pickle: bool = False) -> Iterator[str]:
    """Key names only (safetensors: header read; torch files: memory-mapped load, no tensor data touched)."""
    for f in _files_of(path):
        if f.endswith(".safetensors"):
            from safetensors import safe_open
            with safe_open(f, framework="pt", device="cpu") as h:
                yield from h.keys()
        else:
            yield from _load_torch_bin(f, allow_pickle).keys()


def lora_stems(keys) -> set:
    """Module paths that carry an adapter: `<stem>.lora_A.<name>.weight` / `<stem>.lora_B.<name>.weight`."""
    return {k.split(".lora_")[0] for k in keys if ".lora_A." in k or ".lora_B." in k}


def merge_lora(pairs: Iterator[Tuple[str, torch.Tensor]], scaling: Optional[float] = None, lora_alpha: float = 16.0,
               stems: Optional[set] = None) -> Iterator[Tuple[str, torch.Tensor]]:
    """peft adapters are folded into the base weight: W + (alpha / r) * B @ A.  Both key layouts are handled:
      * peft >= 0.7   `<stem>.base_layer.weight` + `<stem>.lora_A.default.weight` / `lora_B.default.weight`
      * 2023-era peft (the Emu1 instruct checkpoint, Emu1/inference.py:40-57)   `<stem>.weight` next to the adapters.
    `stems` = the module paths that carry an adapter (lora_stems(iter_keys(path))); with it only those base weights are
    held back until their A / B arrive and everything else streams through.  Without it (single pass over an
    in-memory dict) every 2-D `.weight` is held until the end.  Adapter tensors left without a base weight raise."""
    A: Dict[str, torch.Tensor] = {}
    B: Dict[str, torch.Tensor] = {}
    held: Dict[str, torch.Tensor] = {}

    def clean(k):
        return k.replace("base_model.model.", "").replace(".base_layer", "")

    def emit(stem):
        w = held.pop(stem)
        a, b = A.pop(stem), B.pop(stem)
        sc = scaling if scaling is not None else lora_alpha / a.shape[0]
        return clean(stem) + ".weight", (w.float() + sc * (b.float() @ a.float())).to(w.dtype)

    for k, t in pairs:
        if ".lora_A." in k or ".lora_B." in k:
            stem = k.split(".lora_")[0]
            (A if ".lora_A." in k else B)[stem] = t
        elif k.endswith(".base_layer.weight"):
            held[k[: -len(".base_layer.weight")]] = t
        elif k.endswith(".weight") and t.dim() == 2 and (stems is None or k[: -len(".weight")] in stems):
            held[k[: -len(".weight")]] = t  # legacy layout: the adapter of this module may still arrive
        else:
            yield clean(k), t
            continue
        for stem in [s for s in list(held) if s in A and s in B]:
            yield emit(stem)
    for stem in list(held):  # base layers without an adapter
        if stem in A or stem in B:
            raise KeyError("incomplete LoRA adapter for %s (lora_A and lora_B are both required)" % stem)
        yield clean(stem) + ".weight", held.pop(stem)
    left = sorted(set(A) | set(B))
    if left:
        raise KeyError("LoRA adapters without a base weight: %s" % left[:3])


def load_into(sink, path: str, prefix: str = "", rename: Optional[Callable[[str], Optional[str]]] = None,
              lora: bool = False, strict_keys: Optional[set] = None, allow_pickle: bool = False) -> int:
    """Stream every tensor under `path` into sink.load_tensor(prefix + key, tensor).  `rename(key)` may map or drop
    (return None) keys.  `strict_keys`: names that must all arrive (the reference loads with strict=True).  Returns the
    number of tensors loaded."""
    pairs = iter_checkpoint(path, allow_pickle)
    if lora:
        pairs = merge_lora(pairs, stems=lora_stems(iter_keys(path, allow_pickle)))
    n = 0
    seen = set()
    for k, t in pairs:
        if k.endswith("rotary_emb.inv_freq"):
            continue
        if rename is not None:
            k = rename(k)
            if k is None:
                continue
        sink.load_tensor(prefix + k, t)
        seen.add(k)
        n += 1
    if strict_keys is not None:
        missing = strict_keys - seen
        if missing:
            raise KeyError("checkpoint is missing %d tensors, e.g. %s" % (len(missing), sorted(missing)[:3]))
    return n
