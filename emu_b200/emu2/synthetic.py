"""Synthetic prompts / weights of the named shapes for benchmarks and smoke tests (no checkpoint, no tokenizer
file: the reference's tokenizer.model and weights are not redistributed and there is no network)."""
import math

import torch

VOCAB_EMU2 = 32272
IDS = {"bos": 1, "eos": 2, "pad": 32000, "[IMG]": 32001, "[/IMG]": 32002, "<image>": 32003, "[gIMG]": 32004}


class SyntheticTokenizer:
    """Token-id level stand-in with the HF attribute surface EmuModel touches."""
    pad_token_id, bos_token_id, eos_token_id = IDS["pad"], IDS["bos"], IDS["eos"]
    padding_side = truncation_side = "left"

    def __init__(self, vocab=VOCAB_EMU2):
        self.vocab = vocab

    def __len__(self):
        return self.vocab

    def convert_tokens_to_ids(self, toks):
        return [IDS[t] for t in toks]

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(str(int(i)) for i in row) for row in ids]


def image_prompt_ids(n_query=64, n_text=8, batch=1, seed=0):
    """ids of "<s>[IMG]<image>*n_query[/IMG]" + n_text text pieces — the shape of
    "[<IMG_PLH>]Describe the image in details:" after Emu2/emu/emu.py:184-189."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(100, 31000, (batch, n_text), generator=g)
    head = torch.tensor([IDS["bos"], IDS["[IMG]"]] + [IDS["<image>"]] * n_query + [IDS["[/IMG]"]])
    ids = torch.cat((head[None].expand(batch, -1), text), dim=1)
    return ids, torch.ones_like(ids)


def load_random_weights(model, vision_cfg, llama_cfg, vocab, seed=0, std=0.02, device="cuda"):
    """Random-init weights of the real architecture, generated on the device tensor by tensor and handed to the
    engine under the reference's state-dict keys (N(0, std) matrices, unit norm scales, small biases)."""
    g = torch.Generator(device=device).manual_seed(seed)
    eng = model.engine

    def put(key, shape, kind="w"):
        if kind == "w":
            t = torch.randn(shape, generator=g, device=device, dtype=torch.float32).mul_(std).to(torch.bfloat16)
        elif kind == "one":
            t = torch.ones(shape, device=device, dtype=torch.bfloat16)
        else:
            t = torch.randn(shape, generator=g, device=device, dtype=torch.float32).mul_(0.01).to(torch.bfloat16)
        eng.load_tensor(key, t)
        del t

    W, L, P = vision_cfg.width, vision_cfg.layers, vision_cfg.patch_size
    G = vision_cfg.image_size // P
    mlp = int(W * vision_cfg.mlp_ratio)
    put("visual.cls_token", (1, 1, W))
    put("visual.pos_embed", (1, G * G + 1, W))
    put("visual.patch_embed.proj.weight", (W, 3, P, P))
    put("visual.patch_embed.proj.bias", (W,), "b")
    for l in range(L):
        p = f"visual.blocks.{l}."
        for n in ("norm1", "norm2"):
            put(p + n + ".weight", (W,), "one")
            put(p + n + ".bias", (W,), "b")
        put(p + "attn.q_bias", (W,), "b")
        put(p + "attn.v_bias", (W,), "b")
        put(p + "attn.qkv.weight", (3 * W, W))
        put(p + "attn.proj.weight", (W, W))
        put(p + "attn.proj.bias", (W,), "b")
        put(p + "mlp.fc1.weight", (mlp, W))
        put(p + "mlp.fc1.bias", (mlp,), "b")
        put(p + "mlp.fc2.weight", (W, mlp))
        put(p + "mlp.fc2.bias", (W,), "b")
    H, F, NL = llama_cfg["hidden_size"], llama_cfg["intermediate_size"], llama_cfg["num_hidden_layers"]
    put("decoder.lm.model.embed_tokens.weight", (vocab, H))
    for l in range(NL):
        p = f"decoder.lm.model.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            put(p + f"self_attn.{n}.weight", (H, H))
        put(p + "mlp.gate_proj.weight", (F, H))
        put(p + "mlp.up_proj.weight", (F, H))
        put(p + "mlp.down_proj.weight", (H, F))
        put(p + "input_layernorm.weight", (H,), "one")
        put(p + "post_attention_layernorm.weight", (H,), "one")
    put("decoder.lm.model.norm.weight", (H,), "one")
    put("decoder.lm.lm_head.weight", (vocab, H))
    put("project_up.weight", (H, W))
    put("project_down.weight", (W, H))
    torch.cuda.synchronize()
