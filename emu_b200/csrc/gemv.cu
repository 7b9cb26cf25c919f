// emu_b200 — weight-streaming skinny GEMM for the autoregressive decode loop.
//
// y[b, n] = epilogue( sum_k x[b, k] * W[n, k] ),  1 <= b <= 8 (batch x beams), W bf16 row-major [N, K].
//
// This is the dominant kernel of the headline metric: per decoded token the LLaMA-33B decoder streams
// 64.6 GB of bf16 weights (HF LlamaDecoderLayer q/k/v/o/gate/up/down + lm_head, driven from
// Emu2/emu/emu.py:213-229 and :133-138), so the kernel is judged on HBM GB/s, not FLOPs.
//
// Mapping: a CTA owns 16*RT consecutive weight rows; its 8 warps split K in interleaved 32-element blocks,
// every lane streams 16-byte pieces of two rows per block with ld.global.nc.L1::no_allocate (each request is a
// full 64 B per row, 8 rows per instruction) and keeps up to 16 such loads in flight (register double buffer).
// The tiny x operand (<= 8 rows) is staged once per CTA in shared memory — optionally through a fused
// RMSNorm prologue (HF LlamaRMSNorm rounding) — and fed as the 8-wide N operand of mma.sync.m16n8k16, so
// batch 1..8 (greedy .. 5-beam search) all run at the same, bandwidth-bound, speed. fp32 partial sums are
// reduced across the 8 warps through shared memory; the epilogue fuses bias / residual / SwiGLU / RoPE +
// KV-cache append.  Launched with programmatic dependent launch: the first weight tiles are requested before
// griddepcontrol.wait, so HBM keeps streaming across kernel boundaries.
#include "common.cuh"
#include "ops.h"

namespace emu {

constexpr int kGemvWarps = 8;
constexpr int kGemvThreads = kGemvWarps * 32;

struct GemvParams {
  GemvArgs a;
  int ldxs;      // smem row stride of staged x (elements)
  int red_off;   // byte offset of the reduction buffer
};

template <int RT>
__global__ void __launch_bounds__(kGemvThreads) gemv_kernel(const GemvParams p) {
  constexpr int U = 4 / RT;  // k-blocks per register chunk
  const GemvArgs& a = p.a;
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* xs = reinterpret_cast<bf16*>(smem);
  float* red = reinterpret_cast<float*>(smem + p.red_off);  // [8 warps][RT][16][8]
  __shared__ float s_ss[kGemvWarps][8];
  __shared__ float s_rstd[8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int K = a.K, N = a.N, B = a.B;
  const int row_base = blockIdx.x * (16 * RT);

  const bf16* wrow[RT][2];
#pragma unroll
  for (int rt = 0; rt < RT; ++rt) {
    int r0 = row_base + rt * 16 + g, r1 = r0 + 8;
    r0 = r0 < N ? r0 : N - 1;
    r1 = r1 < N ? r1 : N - 1;
    wrow[rt][0] = a.W + (long)r0 * K + t * 8;
    wrow[rt][1] = a.W + (long)r1 * K + t * 8;
  }
  const int KB = K >> 5;
  const int iters = warp < KB ? (KB - warp + kGemvWarps - 1) / kGemvWarps : 0;
  const int nchunks = (iters + U - 1) / U;

  uint4 wq[2][U][RT][2];
#define GEMV_LOAD(BUF, CHUNK)                                                        \
  {                                                                                  \
    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                  \
      const int i_ = (CHUNK)*U + u;                                                  \
      const bool ok_ = i_ < iters;                                                   \
      const int kb_ = warp + kGemvWarps * i_;                                        \
      _Pragma("unroll") for (int rt = 0; rt < RT; ++rt) {                            \
        _Pragma("unroll") for (int h = 0; h < 2; ++h) {                              \
          wq[BUF][u][rt][h] = ok_ ? ldg_stream(wrow[rt][h] + (long)kb_ * 32) : make_uint4(0, 0, 0, 0); \
        }                                                                            \
      }                                                                              \
    }                                                                                \
  }

  // weights do not depend on the previous kernel: request the first chunk before the grid dependency resolves
  GEMV_LOAD(0, 0);
  if (a.pdl) {
    pdl_launch_dependents();
    pdl_wait();
  }

  // ---- stage x (optionally RMS-normalised) into shared memory ----
  const int vec_per_row = K >> 3;
  if (a.norm_w != nullptr) {
    float ss[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) ss[b] = 0.f;
    for (int b = 0; b < B; ++b) {
      const uint4* src = reinterpret_cast<const uint4*>(a.x + (long)b * a.ldx);
      float s = 0.f;
      for (int i = threadIdx.x; i < vec_per_row; i += kGemvThreads) {
        const uint4 v = src[i];
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lo = bf16_lo(w4[j]), hi = bf16_hi(w4[j]);
          s += lo * lo + hi * hi;
        }
      }
      ss[b] = warp_sum(s);
    }
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < 8; ++b) s_ss[warp][b] = ss[b];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kGemvWarps; ++w) tot += s_ss[w][threadIdx.x];
      s_rstd[threadIdx.x] = rsqrtf(tot / (float)K + a.norm_eps);
    }
    __syncthreads();
    for (int b = 0; b < B; ++b) {
      const float rstd = s_rstd[b];
      const uint4* src = reinterpret_cast<const uint4*>(a.x + (long)b * a.ldx);
      const uint4* wsrc = reinterpret_cast<const uint4*>(a.norm_w);
      uint4* dst = reinterpret_cast<uint4*>(xs + (long)b * p.ldxs);
      for (int i = threadIdx.x; i < vec_per_row; i += kGemvThreads) {
        const uint4 v = src[i], w = wsrc[i];
        const uint32_t v4[4] = {v.x, v.y, v.z, v.w}, w4[4] = {w.x, w.y, w.z, w.w};
        uint32_t o4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // HF: weight * (x.float() * rsqrt(var + eps)).to(bf16)
          const float lo = round_bf16(bf16_lo(v4[j]) * rstd) * bf16_lo(w4[j]);
          const float hi = round_bf16(bf16_hi(v4[j]) * rstd) * bf16_hi(w4[j]);
          o4[j] = pack_bf16(lo, hi);
        }
        dst[i] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  } else {
    for (int b = 0; b < B; ++b) {
      const uint4* src = reinterpret_cast<const uint4*>(a.x + (long)b * a.ldx);
      uint4* dst = reinterpret_cast<uint4*>(xs + (long)b * p.ldxs);
      for (int i = threadIdx.x; i < vec_per_row; i += kGemvThreads) dst[i] = src[i];
    }
  }
  __syncthreads();

  float acc[RT][4];
#pragma unroll
  for (int rt = 0; rt < RT; ++rt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[rt][j] = 0.f;

  const bf16* xrow = xs + (long)g * p.ldxs + t * 8;
  const bool has_x = g < B;

#define GEMV_COMPUTE(BUF, CHUNK)                                                     \
  {                                                                                  \
    _Pragma("unroll") for (int u = 0; u < U; ++u) {                                  \
      const int i_ = (CHUNK)*U + u;                                                  \
      if (i_ < iters) {                                                              \
        const int kb_ = warp + kGemvWarps * i_;                                      \
        uint4 xb = make_uint4(0, 0, 0, 0);                                           \
        if (has_x) xb = *reinterpret_cast<const uint4*>(xrow + kb_ * 32);            \
        _Pragma("unroll") for (int rt = 0; rt < RT; ++rt) {                          \
          const uint4 wa = wq[BUF][u][rt][0], wb = wq[BUF][u][rt][1];                \
          const uint32_t a1[4] = {wa.x, wb.x, wa.y, wb.y};                           \
          const uint32_t b1[2] = {xb.x, xb.y};                                       \
          mma_bf16_16816(acc[rt], a1, b1);                                           \
          const uint32_t a2[4] = {wa.z, wb.z, wa.w, wb.w};                           \
          const uint32_t b2[2] = {xb.z, xb.w};                                       \
          mma_bf16_16816(acc[rt], a2, b2);                                           \
        }                                                                            \
      }                                                                              \
    }                                                                                \
  }

  for (int c = 0; c < nchunks; c += 2) {
    if (c + 1 < nchunks) GEMV_LOAD(1, c + 1);
    GEMV_COMPUTE(0, c);
    if (c + 1 < nchunks) {
      if (c + 2 < nchunks) GEMV_LOAD(0, c + 2);
      GEMV_COMPUTE(1, c + 1);
    }
  }
#undef GEMV_LOAD
#undef GEMV_COMPUTE

  // ---- cross-warp reduction ----
#pragma unroll
  for (int rt = 0; rt < RT; ++rt) {
    float* r = red + ((warp * RT + rt) * 16) * 8;
    r[g * 8 + 2 * t] = acc[rt][0];
    r[g * 8 + 2 * t + 1] = acc[rt][1];
    r[(g + 8) * 8 + 2 * t] = acc[rt][2];
    r[(g + 8) * 8 + 2 * t + 1] = acc[rt][3];
  }
  __syncthreads();

  auto reduced = [&](int rt, int r, int b) -> float {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kGemvWarps; ++w) s += red[((w * RT + rt) * 16 + r) * 8 + b];
    return s;
  };

  for (int idx = threadIdx.x; idx < RT * 128; idx += kGemvThreads) {
    const int r = idx & 15, b = (idx >> 4) & 7, rt = idx >> 7;
    const int n = row_base + rt * 16 + r;
    if (b >= B || n >= N) continue;
    if (a.mode == EPI_NONE) {
      float v = reduced(rt, r, b);
      if (a.bias) v += __bfloat162float(a.bias[n]);
      if (a.residual) v = round_bf16(v) + __bfloat162float(a.residual[(long)b * a.ldr + n]);
      if (a.out_fp32) reinterpret_cast<float*>(a.y)[(long)b * a.ldy + n] = v;
      else reinterpret_cast<bf16*>(a.y)[(long)b * a.ldy + n] = __float2bfloat16_rn(v);
    } else if (a.mode == EPI_SWIGLU) {
      if (r & 1) continue;
      const float gate = round_bf16(reduced(rt, r, b)), up = round_bf16(reduced(rt, r + 1, b));
      const float v = round_bf16(silu(gate)) * up;
      reinterpret_cast<bf16*>(a.y)[(long)b * a.ldy + (n >> 1)] = __float2bfloat16_rn(v);
    } else {  // GEMV_ROPE_QKV
      const int D = a.head_dim, H = a.n_heads;
      const int hh = n / D, i = n - hh * D;
      const int slot = a.pos[b];
      if (hh < 2 * H) {
        if (r & 1) continue;
        const float x1 = round_bf16(reduced(rt, r, b)), x2 = round_bf16(reduced(rt, r + 1, b));
        const int rp = slot - (a.pos_off ? a.pos_off[b] : 0);
        const float c = __bfloat162float(a.rope_cos[(long)rp * (D / 2) + (i >> 1)]);
        const float s = __bfloat162float(a.rope_sin[(long)rp * (D / 2) + (i >> 1)]);
        // HF apply_rotary_pos_emb in bf16: (q*cos) + (rotate_half(q)*sin), each op rounded
        const float o1 = round_bf16(x1 * c) + round_bf16(-x2 * s);
        const float o2 = round_bf16(x2 * c) + round_bf16(x1 * s);
        bf16* dst;
        if (hh < H) dst = reinterpret_cast<bf16*>(a.y) + (long)b * a.ldy + n;
        else dst = a.k_cache + (((long)b * H + (hh - H)) * a.t_max + slot) * D + i;
        *reinterpret_cast<uint32_t*>(dst) = pack_bf16(o1, o2);
      } else {
        const float v = reduced(rt, r, b);
        a.v_cache[(((long)b * H + (hh - 2 * H)) * a.t_max + slot) * D + i] = __float2bfloat16_rn(v);
      }
    }
  }
}

template <int RT>
static int launch_gemv(const GemvParams& p, int grid, size_t smem, cudaStream_t st) {
  static size_t cur_max = 0;
  if (smem > cur_max) {
    if (cudaFuncSetAttribute(gemv_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return EMU_ERR_CUDA;
    cur_max = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.a.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemv_kernel<RT>, p) == cudaSuccess ? EMU_OK : EMU_ERR_CUDA;
}

int gemv_bf16(const GemvArgs& a, cudaStream_t st) {
  if (a.B < 1 || a.B > 8 || a.N < 1 || a.K < 32 || (a.K % 32) || (a.ldx % 8)) return EMU_ERR_INVALID;
  if ((a.mode == EPI_SWIGLU || a.mode == GEMV_ROPE_QKV) && (a.N % 16)) return EMU_ERR_INVALID;
  GemvParams p;
  p.a = a;
  p.ldxs = a.K + 32;  // row stride = 64 B (mod 128 B): conflict-free B-fragment reads
  const size_t xs_bytes = (size_t)a.B * p.ldxs * sizeof(bf16);
  const int tiles = (a.N + 15) / 16;
  // widest row group that still gives >= ~3 CTAs per SM
  int rt = 1;
  if (tiles >= 4 * 3 * kNumSMs) rt = 4;
  else if (tiles >= 2 * 3 * kNumSMs) rt = 2;
  p.red_off = (int)((xs_bytes + 15) & ~size_t(15));
  const size_t smem = p.red_off + (size_t)kGemvWarps * rt * 16 * 8 * sizeof(float);
  if (smem > 220 * 1024) return EMU_ERR_UNSUPPORTED;
  const int grid = (tiles + rt - 1) / rt;
  switch (rt) {
    case 4: return launch_gemv<4>(p, grid, smem, st);
    case 2: return launch_gemv<2>(p, grid, smem, st);
    default: return launch_gemv<1>(p, grid, smem, st);
  }
}

}  // namespace emu
