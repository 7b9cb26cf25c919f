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

}  // namespace emu

using namespace emu;

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
