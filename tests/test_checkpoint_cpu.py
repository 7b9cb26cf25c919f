"""Streaming checkpoint ingestion (emu_b200/checkpoint.py): every reference format yields the same (key, tensor) stream."""
import json
import os

import pytest
import torch

from emu_b200 import checkpoint as ck


class Sink:
    def __init__(self):
        self.got = {}

    def load_tensor(self, k, t):
        self.got[k] = t.clone()


def _sd():
    g = torch.Generator().manual_seed(0)
    return {"visual.cls_token": torch.randn(1, 1, 8, generator=g).to(torch.bfloat16),
            "decoder.lm.model.layers.0.self_attn.q_proj.weight": torch.randn(8, 8, generator=g).to(torch.bfloat16),
            "decoder.lm.model.layers.0.self_attn.rotary_emb.inv_freq": torch.randn(4, generator=g),
            "project_up.weight": torch.randn(8, 4, generator=g)}


def _check(sink, sd, prefix=""):
    want = {prefix + k: v for k, v in sd.items() if not k.endswith("inv_freq")}
    assert set(sink.got) == set(want)
    for k in want:
        assert torch.equal(sink.got[k], want[k])


def test_single_files(tmp_path):
    from safetensors.torch import save_file
    sd = _sd()
    save_file(sd, str(tmp_path / "m.safetensors"))
    torch.save(sd, str(tmp_path / "m.bin"))
    torch.save({"module": sd}, str(tmp_path / "emu1.pt"))
    for f in ("m.safetensors", "m.bin", "emu1.pt"):
        s = Sink()
        assert ck.load_into(s, str(tmp_path / f)) == 3
        _check(s, sd)


def test_sharded_index_and_prefix(tmp_path):
    from safetensors.torch import save_file
    sd = _sd()
    keys = sorted(sd)
    shards = {"model-00001-of-00002.safetensors": keys[:2], "model-00002-of-00002.safetensors": keys[2:]}
    wm = {}
    for name, ks in shards.items():
        save_file({k: sd[k] for k in ks}, str(tmp_path / name))
        wm.update({k: name for k in ks})
    json.dump({"weight_map": wm}, open(tmp_path / "model.safetensors.index.json", "w"))
    s = Sink()
    ck.load_into(s, str(tmp_path), prefix="unet.")
    _check(s, sd, "unet.")
    with pytest.raises(KeyError):
        ck.load_into(Sink(), str(tmp_path), strict_keys={"not.there"})


def test_lora_merge(tmp_path):
    g = torch.Generator().manual_seed(1)
    W, A, B = torch.randn(8, 8, generator=g), torch.randn(2, 8, generator=g), torch.randn(8, 2, generator=g)
    sd = {"base_model.model.decoder.q_proj.base_layer.weight": W, "base_model.model.decoder.q_proj.lora_A.default.weight": A,
          "base_model.model.decoder.q_proj.lora_B.default.weight": B, "base_model.model.decoder.norm.weight": torch.ones(8)}
    torch.save(sd, str(tmp_path / "lora.bin"))
    s = Sink()
    ck.load_into(s, str(tmp_path / "lora.bin"), lora=True)
    assert set(s.got) == {"decoder.q_proj.weight", "decoder.norm.weight"}
    assert torch.allclose(s.got["decoder.q_proj.weight"], W + (16.0 / 2) * (B @ A), atol=1e-5)
