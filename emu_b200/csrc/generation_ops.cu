// emu_b200 — device-side generation control for beam search: the vocabulary-wide part of one HF
// `_beam_search` step (transformers GenerationMixin, driven from Emu2/emu/emu.py:213-229 with num_beams=5,
// length_penalty=-1) — log_softmax, repetition penalty, min-length EOS ban, "+ running beam score" and the top-2·beams
// selection over beams x vocab — as three small kernels instead of full-vocabulary torch ops.  Only the [batch, 2·beams]
// bookkeeping that follows stays on the host side (emu_b200/generation.py).
#include <cuda_runtime.h>

#include "common.cuh"
#include "engine.h"
#include "ops.h"

namespace emu {

// x[r, :] = log_softmax(x[r, :]) + add[r]     (one 1024-thread CTA per row, fp32, three passes over an L2-resident row)
__global__ void __launch_bounds__(1024) logsoftmax_add_kernel(float* __restrict__ x, const float* __restrict__ add, int V) {
  __shared__ float red[33];
  float* row = x + (long)blockIdx.x * V;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) m = fmaxf(m, row[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = red[threadIdx.x];
    v = warp_max(v);
    if (threadIdx.x == 0) red[32] = v;
  }
  __syncthreads();
  m = red[32];
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(row[i] - m);
  s = block_sum(s, red);
  const float shift = m + logf(s) - (add ? add[blockIdx.x] : 0.f);
  for (int i = threadIdx.x; i < V; i += blockDim.x) row[i] -= shift;
}

// HF RepetitionPenaltyLogitsProcessor on scores that already carry `add[r]`: s < 0 ? s * p : s / p at every previously
// generated token (each distinct token once, like gather -> where -> scatter)
__global__ void rep_penalty_kernel(float* x, const float* __restrict__ add, const long long* __restrict__ prev, int prev_len,
                                   int V, float penalty) {
  const int r = blockIdx.x;
  const long long* p = prev + (long)r * prev_len;
  for (int j = threadIdx.x; j < prev_len; j += blockDim.x) {
    const long long t = p[j];
    if (t < 0 || t >= V) continue;
    bool first = true;
    for (int k = 0; k < j; ++k) first = first && (p[k] != t);
    if (!first) continue;
    const float a = add ? add[r] : 0.f;
    float s = x[(long)r * V + t] - a;
    s = s < 0.f ? s * penalty : s / penalty;
    x[(long)r * V + t] = s + a;
  }
}

__global__ void ban_token_kernel(float* x, int V, int ban) {
  if (ban >= 0 && ban < V) x[(long)blockIdx.x * V + ban] = -INFINITY;
}

// top-`keep` of each group of n contiguous values (n = beams * V), largest first, ties to the lower index.
// DESTRUCTIVE: every selected entry is overwritten with -inf (the buffer is the decode step's scratch logits).
__global__ void __launch_bounds__(1024) topk_group_kernel(float* x, long n, int keep, float* out_val, int* out_idx) {
  __shared__ float sv[32];
  __shared__ long si[32];
  float* g = x + (long)blockIdx.x * n;
  for (int k = 0; k < keep; ++k) {
    float best = -INFINITY;
    long bi = 0x7fffffffffffffffLL;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = __ldcg(g + i);
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
      best = sv[threadIdx.x];
      bi = si[threadIdx.x];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (threadIdx.x == 0) {
        const bool any = bi != 0x7fffffffffffffffLL;  // all -inf / NaN rows: report index 0
        out_val[(long)blockIdx.x * keep + k] = best;
        out_idx[(long)blockIdx.x * keep + k] = any ? (int)bi : 0;
        if (any) g[bi] = -INFINITY;
        __threadfence_block();
      }
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------
// sampling: HF warper order temperature -> top-k -> top-p, then one multinomial draw per row (Emu2/emu/chat.py:46-57 passes
// do_sample / top_k / top_p / temperature through to GenerationMixin).  Sort-free: both filters are thresholds found by a
// 32-step bitwise search over the order-preserving integer image of the scores, each step one pass over the L2-resident
// row; the draw is an inverse-CDF lookup in index order.  One 1024-thread CTA per row, logits are not modified.
//   top-k : keep x >= (k-th largest x)                         (ties with the k-th value are kept, as `scores < kth` does)
//   top-p : keep x_i iff mass{x > x_i} < top_p                  (== sorted-ascending cumsum <= 1 - top_p removed)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t order_key(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(1024) sample_kernel(const float* __restrict__ logits, int V, float inv_temp, int top_k,
                                                      float top_p, int ban_id, unsigned long long seed,
                                                      unsigned long long offset, int* out_ids) {
  __shared__ float red[33];
  __shared__ float scan[32];
  __shared__ int s_pick;
  const float* row = logits + (long)blockIdx.x * V;
  const int tid = threadIdx.x;
  auto val = [&](int i) { return i == ban_id ? -INFINITY : row[i] * inv_temp; };
  // row maximum (softmax shift)
  float m = -INFINITY;
  for (int i = tid; i < V; i += 1024) m = fmaxf(m, val(i));
  m = warp_max(m);
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  if (tid < 32) {
    float v = warp_max(red[tid]);
    if (tid == 0) red[32] = v;
  }
  __syncthreads();
  m = red[32];
  __syncthreads();
  // ---- top-k threshold: largest t with count{key >= t} >= k ----
  uint32_t tk = 0;
  if (top_k > 0 && top_k < V) {
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = tk | (1u << bit);
      float cnt = 0.f;
      for (int i = tid; i < V; i += 1024) cnt += order_key(val(i)) >= cand ? 1.f : 0.f;
      cnt = block_sum(cnt, red);
      if (cnt >= (float)top_k) tk = cand;
    }
  }
  // ---- top-p threshold on the top-k-filtered distribution: smallest t with mass{key > t} < top_p * Z ----
  float z = 0.f;
  for (int i = tid; i < V; i += 1024) {
    const float x = val(i);
    z += order_key(x) >= tk ? __expf(x - m) : 0.f;
  }
  z = block_sum(z, red);
  uint32_t tp = 0;
  if (top_p < 1.0f) {
    const float target = top_p * z;
    tp = 0xFFFFFFFFu;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = tp & ~(1u << bit);
      float mass = 0.f;
      for (int i = tid; i < V; i += 1024) {
        const float x = val(i);
        const uint32_t k = order_key(x);
        mass += (k > cand && k >= tk) ? __expf(x - m) : 0.f;
      }
      mass = block_sum(mass, red);
      if (mass < target) tp = cand;
    }
  }
  const uint32_t thr = tk > tp ? tk : tp;
  // ---- multinomial draw over the kept set, inverse CDF in index order (thread t owns a contiguous index chunk) ----
  const int chunk = (V + 1023) / 1024;
  const int i0 = tid * chunk, i1 = min(V, i0 + chunk);
  float local = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float x = val(i);
    local += (order_key(x) >= thr && x > -INFINITY) ? __expf(x - m) : 0.f;
  }
  // block exclusive scan of `local`
  float incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += n;
  }
  if ((tid & 31) == 31) scan[tid >> 5] = incl;
  if (tid == 0) s_pick = -1;
  __syncthreads();
  if (tid < 32) {
    float w = scan[tid], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, wi, o);
      if (tid >= o) wi += n;
    }
    scan[tid] = wi - w;  // exclusive prefix of the warp totals
    if (tid == 31) red[32] = wi;
  }
  __syncthreads();
  const float total = red[32];
  const float before = scan[tid >> 5] + incl - local;
  const uint64_t h = mix64(seed ^ mix64(offset * 0x100000001B3ull + blockIdx.x));
  const float u = (float)(h >> 40) * (1.0f / 16777216.0f) * total;  // 24 random bits -> [0, total)
  if (local > 0.f && u >= before && u < before + local) {
    float acc = before;
    int pick = -1;
    for (int i = i0; i < i1; ++i) {
      const float x = val(i);
      if (order_key(x) >= thr && x > -INFINITY) {
        acc += __expf(x - m);
        pick = i;
        if (u < acc) break;
      }
    }
    s_pick = pick;
  }
  __syncthreads();
  if (tid == 0) {
    int pick = s_pick;
    if (pick < 0) {  // rounding at the very end of the CDF: fall back to the arg-max (always kept)
      pick = 0;
    }
    out_ids[blockIdx.x] = pick;
  }
  // arg-max fallback needs the whole block: recompute only in the (rare) miss case
  if (s_pick < 0) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 1024) {
      const float x = val(i);
      if (x > best || (x == best && i < bi)) { best = x; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __shared__ float sv[32];
    __shared__ int si[32];
    if ((tid & 31) == 0) { sv[tid >> 5] = best; si[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 32; ++w)
        if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
      out_ids[blockIdx.x] = bi;
    }
  }
}

}  // namespace emu

using namespace emu;

extern "C" int emu_sample_tokens(const float* logits, int rows, int vocab, float temperature, int top_k, float top_p,
                                 int ban_id, uint64_t seed, uint64_t offset, int32_t* out_ids, emu_stream_t stream) {
  if (!logits || !out_ids || rows < 1 || vocab < 1 || !(temperature > 0.f)) return EMU_ERR_INVALID;
  if (!(top_p > 0.f)) return EMU_ERR_INVALID;
  sample_kernel<<<rows, 1024, 0, (cudaStream_t)stream>>>(logits, vocab, 1.0f / temperature, top_k, top_p, ban_id,
                                                         (unsigned long long)seed, (unsigned long long)offset, out_ids);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? EMU_OK : EMU_ERR_CUDA;
}


extern "C" int emu_beam_topk(float* logits, const float* running_scores, int batch, int beams, int vocab, int keep, int ban_id,
                             const long long* prev_tokens, int prev_len, float repetition_penalty, float* out_lp,
                             int* out_idx, emu_stream_t stream) {
  if (!logits || !out_lp || !out_idx || batch < 1 || beams < 1 || vocab < 1 || keep < 1 || (long)keep > (long)beams * vocab)
    return EMU_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = batch * beams;
  logsoftmax_add_kernel<<<rows, 1024, 0, st>>>(logits, running_scores, vocab);
  if (prev_tokens && prev_len > 0 && repetition_penalty != 1.0f)
    rep_penalty_kernel<<<rows, 128, 0, st>>>(logits, running_scores, prev_tokens, prev_len, vocab, repetition_penalty);
  if (ban_id >= 0) ban_token_kernel<<<rows, 1, 0, st>>>(logits, vocab, ban_id);
  topk_group_kernel<<<batch, 1024, 0, st>>>(logits, (long)beams * vocab, keep, out_lp, out_idx);
  count_launch(2 + (ban_id >= 0 ? 1 : 0));
  return cudaGetLastError() == cudaSuccess ? EMU_OK : EMU_ERR_CUDA;
}
