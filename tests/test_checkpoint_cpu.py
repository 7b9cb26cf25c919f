"""Streaming checkpoint ingestion (emu_b200/checkpoint.py): every reference format yields the same (key, tensor) stream."""
import json
import os

import pytest
import torch

from emu_b200 import checkpoint as ck


class _Payload:  # a global the weights_only unpickler rejects
    def __init__(self):
        self.t = torch.ones(2)


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


def test_lora_merge_legacy_key_layout(tmp_path):
    """2023-era peft (the Emu1 instruct checkpoint, Emu1/inference.py:40-57): the base weight is `<stem>.weight` right next to
    `<stem>.lora_A.default.weight` / `lora_B.default.weight`, in either order; it must be merged, never passed through."""
    g = torch.Generator().manual_seed(2)
    W, A, B = torch.randn(8, 8, generator=g), torch.randn(4, 8, generator=g), torch.randn(8, 4, generator=g)
    W2 = torch.randn(8, 8, generator=g)
    for order in (0, 1):
        items = [("base_model.model.decoder.q_proj.weight", W), ("base_model.model.decoder.q_proj.lora_A.default.weight", A),
                 ("base_model.model.decoder.q_proj.lora_B.default.weight", B), ("base_model.model.decoder.k_proj.weight", W2),
                 ("base_model.model.decoder.norm.weight", torch.ones(8))]
        if order:
            items = items[1:3] + items[:1] + items[3:]
        f = str(tmp_path / ("legacy%d.bin" % order))
        torch.save(dict(items), f)
        s = Sink()
        ck.load_into(s, f, lora=True)
        assert set(s.got) == {"decoder.q_proj.weight", "decoder.k_proj.weight", "decoder.norm.weight"}
        assert torch.allclose(s.got["decoder.q_proj.weight"], W + (16.0 / 4) * (B @ A), atol=1e-5)
        assert torch.equal(s.got["decoder.k_proj.weight"], W2)     # no adapter: untouched
    # an adapter whose base weight never arrives is an error, not a silent drop
    torch.save({"x.lora_A.default.weight": A, "x.lora_B.default.weight": B}, str(tmp_path / "orphan.bin"))
    with pytest.raises(KeyError):
        ck.load_into(Sink(), str(tmp_path / "orphan.bin"), lora=True)


def test_unsafe_pickle_needs_opt_in(tmp_path):
    """A .bin that is not a plain tensor dict must NOT be re-loaded with the unrestricted unpickler behind the caller's back."""
    torch.save({"w": torch.ones(2), "meta": _Payload()}, str(tmp_path / "odd.bin"))
    with pytest.raises(Exception):
        list(ck.iter_checkpoint(str(tmp_path / "odd.bin")))
    got = dict(ck.iter_checkpoint(str(tmp_path / "odd.bin"), allow_pickle=True))   # explicit opt-in still works
    assert set(got) == {"w", "meta"}


def test_missing_shard_is_an_error(tmp_path):
    from safetensors.torch import save_file
    sd = _sd()
    save_file({k: sd[k] for k in list(sd)[:2]}, str(tmp_path / "a.safetensors"))
    json.dump({"weight_map": {"x": "a.safetensors", "y": "b.safetensors"}}, open(tmp_path / "model.safetensors.index.json", "w"))
    with pytest.raises(FileNotFoundError):
        ck.load_into(Sink(), str(tmp_path))
