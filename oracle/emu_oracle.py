"""TEST INFRASTRUCTURE — CPU oracle for the Emu generate path (never imported by the product package).

A functional restatement, in plain PyTorch on CPU, of the arithmetic the reference executes on the hot path.
Every function works on a reference-format ``state_dict`` and runs in whatever dtype the tensors have:
fp32 gives the numerical oracle; bf16 reproduces the reference's own rounding points (the reference scripts run
the model in bf16: Emu2/emu/chat.py:202, Emu1/inference.py:176).

Pinned against the UNMODIFIED reference modules imported from /root/reference (tests/test_oracle_vs_reference.py,
tests/golden/gen_golden.py): EVAVisionTransformer, EmuModel.encode_image / generate / generate_image.
The LLaMA block arithmetic lives in third-party `transformers` (pinned 4.31.0 in Emu2/requirements.txt:2, 5.5.0
installed here); it is restated from the published HF algorithm and pinned through the reference's call sites.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# EVA-CLIP ViT — Emu2/emu/eva_vit.py (post-norm) and Emu1/models/eva_vit_model.py (pre-norm)
# ------------------------------------------------------------------------------------------------
def vit_attention(x, sd, pre, num_heads):
    """Attention.forward, math path — Emu2/emu/eva_vit.py:182-252 (xattn=False)."""
    B, N, C = x.shape
    q_bias, v_bias = sd[pre + "attn.q_bias"], sd[pre + "attn.v_bias"]
    qkv_bias = torch.cat((q_bias, torch.zeros_like(v_bias), v_bias))  # K has no bias (:194-196)
    qkv = F.linear(x, sd[pre + "attn.qkv.weight"], qkv_bias)
    qkv = qkv.reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scale = (C // num_heads) ** -0.5
    q = q * scale
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    return F.linear(x, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])


def vit_mlp(x, sd, pre):
    """Mlp.forward — Emu2/emu/eva_vit.py:105-114 (exact-erf GELU, no ffn_ln)."""
    x = F.linear(x, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
    x = F.gelu(x)
    return F.linear(x, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def vit_forward_features(sd, image, *, patch, num_heads, layers, postnorm=True, eps=1e-6, prefix="visual."):
    """EVAVisionTransformer.forward_features — Emu2/emu/eva_vit.py:402-431; Emu1 eva_vit_model.py:636-665."""
    w = sd[prefix + "patch_embed.proj.weight"]
    x = F.conv2d(image, w, sd[prefix + "patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    B = x.shape[0]
    cls = sd[prefix + "cls_token"].expand(B, -1, -1)
    x = torch.cat((cls, x), dim=1) + sd[prefix + "pos_embed"]
    C = x.shape[-1]
    for l in range(layers):
        pre = f"{prefix}blocks.{l}."
        n1 = lambda t: F.layer_norm(t, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps)
        n2 = lambda t: F.layer_norm(t, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps)
        if postnorm:  # Emu2: x + LN(f(x))  (eva_vit.py:298-300)
            x = x + n1(vit_attention(x, sd, pre, num_heads))
            x = x + n2(vit_mlp(x, sd, pre))
        else:  # Emu1: x + f(LN(x))  (eva_vit_model.py:415-416)
            x = x + vit_attention(n1(x), sd, pre, num_heads)
            x = x + vit_mlp(n2(x), sd, pre)
    return x


def encode_image(sd, image, *, patch, num_heads, layers, n_query, prefix="visual."):
    """EmuModel.encode_image — Emu2/emu/emu.py:77-90."""
    x = vit_forward_features(sd, image, patch=patch, num_heads=num_heads, layers=layers, postnorm=True, prefix=prefix)
    x = x[:, 1:, :]
    b, n, c = x.shape
    s = int(n ** 0.5)
    x = x.permute(0, 2, 1).reshape(b, c, s, s)
    stride = int(s // (n_query ** 0.5))
    x = F.avg_pool2d(x, kernel_size=(stride, stride), stride=stride)
    return x.reshape(b, c, -1).permute(0, 2, 1).contiguous()


# ------------------------------------------------------------------------------------------------
# LLaMA decoder (HF LlamaModel / LlamaDecoderLayer, eager attention)
# ------------------------------------------------------------------------------------------------
def rms_norm(x, w, eps):
    """HF LlamaRMSNorm: fp32 normalise, cast back, THEN multiply by the weight."""
    dt = x.dtype
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(dt)


def rotate_half(x):
    d = x.shape[-1] // 2
    return torch.cat((-x[..., d:], x[..., :d]), dim=-1)


def rope_cos_sin(position_ids, head_dim, theta, dtype):
    """HF LlamaRotaryEmbedding.forward: fp32 freqs, cos/sin cast to the activation dtype."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = position_ids[..., None].float() * inv_freq
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


class KVCache:
    def __init__(self, layers):
        self.k = [None] * layers
        self.v = [None] * layers

    def append(self, l, k, v):
        self.k[l] = k if self.k[l] is None else torch.cat((self.k[l], k), dim=2)
        self.v[l] = v if self.v[l] is None else torch.cat((self.v[l], v), dim=2)
        return self.k[l], self.v[l]

    def reorder(self, idx):
        self.k = [t.index_select(0, idx) for t in self.k]
        self.v = [t.index_select(0, idx) for t in self.v]

    def length(self):
        return 0 if self.k[0] is None else self.k[0].shape[2]


def llama_forward(sd, embeds, attention_mask, *, layers, heads, eps=1e-6, theta=10000.0, position_ids=None,
                  cache=None, prefix="decoder.lm.model.", final_norm=True):
    """LlamaModel.forward on inputs_embeds [B,N,H]; attention_mask [B, past+N] (1 = keep).
    position_ids default: arange(past, past+N) — what lm.model(inputs_embeds, attention_mask) uses
    (Emu2/emu/emu.py:133-138).  lm.generate() passes cumsum(mask)-1 instead (HF prepare_inputs_for_generation)."""
    B, N, H = embeds.shape
    D = H // heads
    past = cache.length() if cache is not None else 0
    if position_ids is None:
        position_ids = torch.arange(past, past + N).unsqueeze(0).expand(B, -1)
    cos, sin = rope_cos_sin(position_ids, D, theta, embeds.dtype)
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    T = past + N
    neg = torch.finfo(embeds.dtype).min
    causal = torch.zeros(N, T, dtype=embeds.dtype)
    idx_q = torch.arange(past, past + N)[:, None]
    idx_k = torch.arange(T)[None, :]
    causal = causal.masked_fill(idx_k > idx_q, neg)
    mask = causal[None, None].expand(B, 1, N, T).clone()
    if attention_mask is not None:
        mask = mask.masked_fill(attention_mask[:, None, None, :T] == 0, neg)
    h = embeds
    for l in range(layers):
        p = f"{prefix}layers.{l}."
        x = rms_norm(h, sd[p + "input_layernorm.weight"], eps)
        q = F.linear(x, sd[p + "self_attn.q_proj.weight"]).view(B, N, heads, D).transpose(1, 2)
        k = F.linear(x, sd[p + "self_attn.k_proj.weight"]).view(B, N, heads, D).transpose(1, 2)
        v = F.linear(x, sd[p + "self_attn.v_proj.weight"]).view(B, N, heads, D).transpose(1, 2)
        q = (q * cos) + (rotate_half(q) * sin)
        k = (k * cos) + (rotate_half(k) * sin)
        if cache is not None:
            k, v = cache.append(l, k, v)
        w = torch.matmul(q, k.transpose(2, 3)) * (D ** -0.5) + mask
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        a = torch.matmul(w, v).transpose(1, 2).reshape(B, N, H)
        h = h + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        x = rms_norm(h, sd[p + "post_attention_layernorm.weight"], eps)
        g = F.linear(x, sd[p + "mlp.gate_proj.weight"])
        u = F.linear(x, sd[p + "mlp.up_proj.weight"])
        h = h + F.linear(F.silu(g) * u, sd[p + "mlp.down_proj.weight"])
    return rms_norm(h, sd[prefix + "norm.weight"], eps) if final_norm else h


def lm_logits(sd, hidden, prefix="decoder.lm."):
    return F.linear(hidden, sd[prefix + "lm_head.weight"])


def hf_position_ids(attention_mask):
    """HF generate: position_ids = cumsum(mask) - 1, pads set to 1."""
    pos = attention_mask.long().cumsum(-1) - 1
    return pos.masked_fill(attention_mask == 0, 1)


# ------------------------------------------------------------------------------------------------
# EmuModel.generate / generate_image on token ids (tokenisation itself stays with the HF tokenizer)
# ------------------------------------------------------------------------------------------------
def splice_embeds(sd, input_ids, image_embeds, image_token_id, prefix="decoder.lm.model."):
    """embed_tokens + masked scatter of projected image embeddings — Emu2/emu/emu.py:193-203."""
    text_embeds = F.embedding(input_ids, sd[prefix + "embed_tokens.weight"])
    if image_embeds is not None:
        text_embeds = text_embeds.clone()
        text_embeds[input_ids == image_token_id] = image_embeds.to(text_embeds.dtype)
    return text_embeds


def generate_greedy(sd, inputs_embeds, attention_mask, *, layers, heads, max_new_tokens, eos_id=2, min_len=0,
                    eps=1e-6, theta=10000.0, return_logits=False):
    """Greedy lm.generate(inputs_embeds=..., num_beams=1, do_sample=False) — Emu2/emu/emu.py:213-229.
    Returns only the new tokens (HF semantics when driven by inputs_embeds)."""
    B = inputs_embeds.shape[0]
    cache = KVCache(layers)
    mask = attention_mask.clone()
    pos = hf_position_ids(mask)
    h = llama_forward(sd, inputs_embeds, mask, layers=layers, heads=heads, eps=eps, theta=theta, position_ids=pos,
                      cache=cache)
    out, all_logits = [], []
    finished = torch.zeros(B, dtype=torch.bool)
    for step in range(max_new_tokens):
        logits = lm_logits(sd, h[:, -1, :]).float()
        if return_logits:
            all_logits.append(logits.clone())
        if step < min_len:
            logits[:, eos_id] = -float("inf")
        nxt = logits.argmax(-1)
        nxt = torch.where(finished, torch.full_like(nxt, 32000), nxt)
        out.append(nxt)
        finished |= nxt == eos_id
        if bool(finished.all()) or step == max_new_tokens - 1:
            break
        mask = torch.cat((mask, torch.ones(B, 1, dtype=mask.dtype)), dim=1)
        pos_new = (mask.long().sum(-1, keepdim=True) - 1)
        emb = F.embedding(nxt, sd["decoder.lm.model.embed_tokens.weight"]).unsqueeze(1)
        h = llama_forward(sd, emb, mask, layers=layers, heads=heads, eps=eps, theta=theta, position_ids=pos_new,
                          cache=cache)
    toks = torch.stack(out, dim=1)
    return (toks, all_logits) if return_logits else toks


def generate_image_regress(sd, input_ids_fn, n_query, *, layers, heads, image_token_id, boi_token_id,
                           prompt_image_embeds=None, eps=1e-6, theta=10000.0):
    """EmuModel.generate_image, literal (no cache) — Emu2/emu/emu.py:92-153.
    ``input_ids_fn(k)`` returns (input_ids [B,N_k], attention_mask) of iteration k, i.e. the tokenisation of
    text + "[IMG]" + "<image>" * k (the reference re-tokenises the growing string each iteration, :109-115)."""
    target = None
    for k in range(n_query):
        input_ids, attention_mask = input_ids_fn(k)
        text_embeds = F.embedding(input_ids, sd["decoder.lm.model.embed_tokens.weight"]).clone()
        image_idx = input_ids == image_token_id
        cumsum_idx = torch.flip(torch.cumsum(torch.flip(image_idx, dims=[1]), dim=1), dims=[1])
        if prompt_image_embeds is not None:
            prompt_idx = torch.logical_and(image_idx, cumsum_idx > k)
            text_embeds[prompt_idx] = prompt_image_embeds.to(text_embeds.dtype)
        if target is not None:
            target_idx = torch.logical_and(image_idx, torch.logical_and(cumsum_idx > 0, cumsum_idx <= k))
            text_embeds[target_idx] = F.linear(target, sd["project_up.weight"])
        hidden = llama_forward(sd, text_embeds, attention_mask, layers=layers, heads=heads, eps=eps, theta=theta)
        image_idx = (input_ids == image_token_id) + (input_ids == boi_token_id)
        cumsum_idx = torch.flip(torch.cumsum(torch.flip(image_idx, dims=[1]), dim=1), dims=[1])
        target_idx = torch.logical_and(image_idx, torch.logical_and(cumsum_idx > 0, cumsum_idx <= k + 1))
        target = hidden[target_idx]
        target = F.linear(target.view(-1, target.shape[-1]), sd["project_down.weight"])
    B = hidden.shape[0]
    return target.view(B, -1, target.shape[-1])


def generate_image_cached(sd, prompt_embeds, attention_mask, n_query, *, layers, heads, eps=1e-6, theta=10000.0):
    """Cache-equivalent form (SURVEY.md §8a' item 2): prefill the prompt ending in [IMG] once, then n_query-1
    single-position steps feeding project_up(project_down(h_last)).  Positions are arange (no position_ids)."""
    B = prompt_embeds.shape[0]
    cache = KVCache(layers)
    mask = attention_mask.clone()
    h = llama_forward(sd, prompt_embeds, mask, layers=layers, heads=heads, eps=eps, theta=theta, cache=cache)
    outs = []
    last = h[:, -1, :]
    for k in range(n_query):
        down = F.linear(last, sd["project_down.weight"])
        outs.append(down)
        if k == n_query - 1:
            break
        emb = F.linear(down, sd["project_up.weight"]).unsqueeze(1)
        mask = torch.cat((mask, torch.ones(B, 1, dtype=mask.dtype)), dim=1)
        h = llama_forward(sd, emb, mask, layers=layers, heads=heads, eps=eps, theta=theta, cache=cache)
        last = h[:, -1, :]
    return torch.stack(outs, dim=1)


# ------------------------------------------------------------------------------------------------
# op-level oracles used by the kernel parity tests (fp32 math on the given inputs)
# ------------------------------------------------------------------------------------------------
def op_linear(x, w, bias=None):
    return F.linear(x.float(), w.float(), None if bias is None else bias.float())


def op_attention(q, k, v, scale, causal=False, kv_start=None, bias=None):
    """q [B,Nq,H,D], k/v [B,Nk,H,D] -> [B,Nq,H,D]; causal aligned to the END of the key axis."""
    q, k, v = q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)
    B, H, Nq, D = q.shape
    Nk = k.shape[2]
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if bias is not None:
        s = s + bias.float()[None]
    neg = float("-inf")
    if causal:
        iq = torch.arange(Nq)[:, None] + (Nk - Nq)
        ik = torch.arange(Nk)[None, :]
        s = s.masked_fill((ik > iq)[None, None], neg)
    if kv_start is not None:
        ik = torch.arange(Nk)[None, :]
        s = s.masked_fill((ik < kv_start[:, None])[:, None, None, :], neg)
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)
    return torch.matmul(p, v).transpose(1, 2)


def rel_err(a, b):
    """max-abs error relative to the reference's max-abs (the parity metric of BASELINE.json)."""
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
