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


def _iter_torch_bin(path: str) -> Iterator[Tuple[str, torch.Tensor]]:
    try:
        sd = torch.load(path, map_location="cpu", mmap=True, weights_only=True)
    except Exception:  # legacy (non-zip) pickles cannot be memory-mapped
        sd = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(sd, dict) and "module" in sd and isinstance(sd["module"], dict):
        sd = sd["module"]  # Emu1: torch.load(ckpt)['module']  (Emu1/inference.py:44-45)
    for k in list(sd.keys()):
        yield k, sd.pop(k)


def iter_checkpoint(path: str) -> Iterator[Tuple[str, torch.Tensor]]:
    """Yield (key, cpu tensor) pairs of a file, a sharded directory (index.json) or a directory of weight files."""
    if osp.isdir(path):
        for index in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
            ip = osp.join(path, index)
            if osp.exists(ip):
                shards = sorted(set(json.load(open(ip))["weight_map"].values()))
                for s in shards:
                    yield from iter_checkpoint(osp.join(path, s))
                return
        files = sorted(f for f in os.listdir(path) if f.endswith((".safetensors", ".bin", ".pt", ".pth")))
        if not files:
            raise FileNotFoundError("no checkpoint files under %s" % path)
        for f in files:
            yield from iter_checkpoint(osp.join(path, f))
        return
    if path.endswith(".safetensors"):
        yield from _iter_safetensors(path)
    else:
        yield from _iter_torch_bin(path)


def merge_lora(pairs: Iterator[Tuple[str, torch.Tensor]], scaling: Optional[float] = None, lora_alpha: float = 16.0
               ) -> Iterator[Tuple[str, torch.Tensor]]:
    """peft-style adapters (`…base_layer.weight` / `…lora_A.default.weight` / `…lora_B.default.weight`, as produced for the
    Emu1 instruct checkpoint, Emu1/inference.py:47-57) are folded into the base weight: W + (alpha / r) * B @ A.  Keys
    are renamed to the plain module path.  Adapter tensors are small, so they are buffered; base weights stream through."""
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
        else:
            yield clean(k), t
            continue
        for stem in [s for s in list(held) if s in A and s in B]:
            yield emit(stem)
    for stem in list(held):  # base layers without an adapter
        yield clean(stem) + ".weight", held.pop(stem)


def load_into(sink, path: str, prefix: str = "", rename: Optional[Callable[[str], Optional[str]]] = None,
              lora: bool = False, strict_keys: Optional[set] = None) -> int:
    """Stream every tensor under `path` into sink.load_tensor(prefix + key, tensor).  `rename(key)` may map or drop
    (return None) keys.  Returns the number of tensors loaded."""
    pairs = iter_checkpoint(path)
    if lora:
        pairs = merge_lora(pairs)
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
