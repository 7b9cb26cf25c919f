"""TEST INFRASTRUCTURE — CPU oracle for the Emu1 Causal-Former (never imported by the product package).

Restates Emu1/models/causal_former.py:43-62 over the vendored T5 decoder stack (Emu1/models/modeling_t5.py):
T5LayerNorm :318-331, T5Attention.forward :537-689 (no score scaling, relative position bias from block 0 shared by
all blocks :455-536, fp32 softmax :666, cross-attention K/V from encoder_width :422-424), T5DenseActDense :352-365
(ReLU), T5Block :787-905, T5Stack.forward :1100-1366 (decoder => causal self-attention mask, final_layer_norm).

PINNED (bit-exact): the vendored modeling_t5.py was written for transformers 4.31; oracle/ref_shim.py
`import_emu1_causal_former()` puts back the inert helpers transformers 5.x removed (pruning / device-map utilities,
`get_head_mask`, the `add_cross_attention` config default, a local t5-base config instead of the hub download) and
imports the UNMODIFIED reference module.  tests/test_oracle_cpu.py::test_t5_oracle_vs_live_reference_t5_base checks this
file bit for bit against it at the real t5-base dimensions (fp32 and bf16), and
tests/golden/emu1_cformer_tiny.pt (tests/golden/gen_golden_cformer.py) carries reference outputs of a shrunk stack to the
GPU box for `test_cformer_vs_reference_golden`.  t5-base constants: d_model 768, d_kv 64, 12 heads, d_ff 3072, 12 layers,
32 buckets, max distance 128, eps 1e-6.
"""
import math

import torch
import torch.nn.functional as F

T5_BASE = dict(d_model=768, heads=12, d_ff=3072, layers=12, buckets=32, max_distance=128, eps=1e-6)


def t5_layer_norm(x, w, eps):
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    h = x * torch.rsqrt(var + eps)
    if w.dtype in (torch.float16, torch.bfloat16):
        h = h.to(w.dtype)
    return w * h


def relative_position_bucket(relative_position, num_buckets=32, max_distance=128):
    """bidirectional=False branch of T5Attention._relative_position_bucket."""
    rp = -torch.min(relative_position, torch.zeros_like(relative_position))
    max_exact = num_buckets // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return torch.where(is_small, rp, large)


def compute_bias(table, q_len, k_len, cfg):
    ctx = torch.arange(q_len)[:, None]
    mem = torch.arange(k_len)[None, :]
    b = relative_position_bucket(mem - ctx, cfg["buckets"], cfg["max_distance"])
    return F.embedding(b, table).permute(2, 0, 1).unsqueeze(0)  # [1, H, q, k]


def t5_attention(x, kv_src, sd, p, heads, bias):
    B, N, _ = x.shape
    inner = sd[p + "q.weight"].shape[0]
    D = inner // heads
    shape = lambda t: t.view(B, -1, heads, D).transpose(1, 2)
    q = shape(F.linear(x, sd[p + "q.weight"]))
    k = shape(F.linear(kv_src, sd[p + "k.weight"]))
    v = shape(F.linear(kv_src, sd[p + "v.weight"]))
    scores = torch.matmul(q, k.transpose(3, 2))
    scores = scores + bias.to(scores.dtype)
    w = F.softmax(scores.float(), dim=-1).type_as(scores)
    o = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, -1, inner)
    return F.linear(o, sd[p + "o.weight"])


def causal_former(sd, img_embeds, cfg=T5_BASE, prefix="cformer."):
    """CausalFormer.forward: img_embeds [B, Nv, enc_width] -> [B, n_causal, out_dim]."""
    B = img_embeds.shape[0]
    tokens = sd[prefix + "causal_tokens"]
    h = tokens.expand(B, -1, -1).to(img_embeds.dtype)
    Q = h.shape[1]
    heads, eps = cfg["heads"], cfg["eps"]
    P = prefix + "cformer."
    neg = torch.finfo(h.dtype).min
    causal = torch.zeros(Q, Q, dtype=h.dtype).masked_fill(torch.arange(Q)[None, :] > torch.arange(Q)[:, None], neg)
    bias_self = compute_bias(sd[P + "block.0.layer.0.SelfAttention.relative_attention_bias.weight"], Q, Q, cfg)
    bias_self = bias_self.to(h.dtype) + causal[None, None]
    bias_cross = torch.zeros(1, heads, Q, img_embeds.shape[1], dtype=h.dtype)
    for i in range(cfg["layers"]):
        b = f"{P}block.{i}.layer."
        n = t5_layer_norm(h, sd[b + "0.layer_norm.weight"], eps)
        h = h + t5_attention(n, n, sd, b + "0.SelfAttention.", heads, bias_self)
        n = t5_layer_norm(h, sd[b + "1.layer_norm.weight"], eps)
        h = h + t5_attention(n, img_embeds, sd, b + "1.EncDecAttention.", heads, bias_cross)
        n = t5_layer_norm(h, sd[b + "2.layer_norm.weight"], eps)
        h = h + F.linear(F.relu(F.linear(n, sd[b + "2.DenseReluDense.wi.weight"])), sd[b + "2.DenseReluDense.wo.weight"])
    h = t5_layer_norm(h, sd[P + "final_layer_norm.weight"], eps)
    return F.linear(h, sd[prefix + "projection.weight"], sd[prefix + "projection.bias"])


def param_shapes(cfg, enc_width, out_dim, n_causal=32, prefix="cformer."):
    d, ff, H = cfg["d_model"], cfg["d_ff"], cfg["heads"]
    s = {prefix + "causal_tokens": (1, n_causal, d), prefix + "projection.weight": (out_dim, d),
         prefix + "projection.bias": (out_dim,), prefix + "cformer.final_layer_norm.weight": (d,)}
    for i in range(cfg["layers"]):
        b = f"{prefix}cformer.block.{i}.layer."
        for n in "qkvo":
            s[b + f"0.SelfAttention.{n}.weight"] = (d, d)
        if i == 0:
            s[b + "0.SelfAttention.relative_attention_bias.weight"] = (cfg["buckets"], H)
        s[b + "0.layer_norm.weight"] = (d,)
        s[b + "1.EncDecAttention.q.weight"] = (d, d)
        s[b + "1.EncDecAttention.k.weight"] = (d, enc_width)
        s[b + "1.EncDecAttention.v.weight"] = (d, enc_width)
        s[b + "1.EncDecAttention.o.weight"] = (d, d)
        s[b + "1.layer_norm.weight"] = (d,)
        s[b + "2.DenseReluDense.wi.weight"] = (ff, d)
        s[b + "2.DenseReluDense.wo.weight"] = (d, ff)
        s[b + "2.layer_norm.weight"] = (d,)
    return s
