// emu_b200 — TMA-fed variant of the decode-loop skinny GEMM (see gemv.cu for the problem statement and the
// stream-K decomposition; this file is the production path whenever the shape allows, gemv.cu is the fallback).
//
// Why a second kernel: gemv.cu keeps its bytes-in-flight in registers (16 x 16 B per lane).  That needs two 256-thread
// CTAs per SM to cover HBM latency, which fills the register file — so the NEXT kernel of the decode step cannot
// become resident until this one exits and programmatic dependent launch has nothing to overlap: ~4 us of HBM idle per
// launch x 300 launches per token (measured: 12.9 ms/token vs 11.5 ms of summed kernel time).  Here the bytes in flight
// live in shared memory instead:
//   * one producer thread issues cp.async.bulk.tensor (TMA) loads of [32 rows x 64 cols] 128B-swizzled tiles into a
//     5-stage x 16 KB ring, signalled by mbarriers — 80 KB in flight per CTA with zero registers and zero address math;
//   * 8 consumer warps pull A fragments with ldmatrix (conflict-free through the swizzle) and run mma.sync with the
//     staged x rows as the 8-wide N operand, then release the stage;
//   * ONE CTA per SM (grid = 148, ~100 KB smem, 64 regs): the next kernel's CTA fits beside it, so with PDL its ring is
//     already full when this kernel drains — HBM never idles across kernel boundaries.
// Ragged N and K tails cost nothing (TMA zero-fills out-of-bounds rows/columns).
#include <stdlib.h>

#include "common.cuh"
#include "ops.h"

namespace emu {

int make_tmap_2d(CUtensorMap* out, const void* base, long rows, long cols, long ld, int box_rows);  // gemm_tc.cu

constexpr int kTW = 8;                  // consumer warps
constexpr int kTThreads = (kTW + 1) * 32;  // + 1 producer warp
constexpr int kTRT = 2;                 // 16-row tiles per chunk
constexpr int kTRows = 16 * kTRT;       // 32 rows
constexpr int kTCols = 256;             // columns per chunk (4 TMA tiles of 64)
constexpr int kTStageBytes = kTRows * kTCols * 2;  // 16 KB
constexpr int kTStages = 10;  // ring slots carved; p.nstages of them are used
constexpr int kTMaxParts = 8;
constexpr int kTWsGroups = 8192;

struct GemvTmaParams {
  GemvArgs a;
  int ldxs;      // smem row stride of staged x (elements), = Kpad + 8
  int kpad;      // K rounded up to 256
  int cpt;       // chunks per row group
  int nstages;   // ring depth actually used (<= kTStages; fewer when x is wide so two kernels still co-reside)
  long total;    // total chunks
  float* ws;     // [groups][kTMaxParts][kTRT*128]
  int* counters;
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// finish k-segment [kc_lo, kc_hi] of row group grp (consumer threads only; `red` already holds per-warp partials)
__device__ __noinline__ void gemv_tma_flush(const GemvTmaParams* sp, float* red, float* fin, int* s_last_p, int grp,
                                            int kc_lo, int kc_hi) {
  const GemvTmaParams& p = *sp;
  const GemvArgs& a = p.a;
  const int N = a.N, B = a.B, CPT = p.cpt;
  constexpr int nval = kTRT * 128;
  const long G = gridDim.x;
  const int tid = threadIdx.x;  // < 256
  consumer_bar();
  const bool whole = (kc_lo == 0 && kc_hi == CPT - 1);
  bool do_epilogue = whole;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < kTW; ++w) v += red[w * nval + tid];
  if (whole) {
    fin[tid] = v;
  } else {
    const long first_chunk = (long)grp * CPT;
    const int first_owner = (int)(((first_chunk + 1) * G - 1) / p.total);
    const int last_owner = (int)(((first_chunk + CPT) * G - 1) / p.total);
    const int nparts = last_owner - first_owner + 1;
    const int my = (int)blockIdx.x - first_owner;
    float* wt = p.ws + ((long)grp * kTMaxParts) * nval;
    wt[my * nval + tid] = v;
    __threadfence();
    consumer_bar();
    if (tid == 0) {
      const int prev = atomicAdd(&p.counters[grp], 1);
      *s_last_p = (prev == nparts - 1);
      if (prev == nparts - 1) p.counters[grp] = 0;
    }
    consumer_bar();
    do_epilogue = *s_last_p != 0;
    if (do_epilogue) {
      __threadfence();
      float s = 0.f;
      for (int q = 0; q < nparts; ++q) s += __ldcg(&wt[q * nval + tid]);  // fixed order: deterministic
      fin[tid] = s;
    }
  }
  if (do_epilogue) {
    consumer_bar();
    const int rt = tid >> 7, r = tid & 15, b = (tid >> 4) & 7;
    const float* f = fin + rt * 128;
    const int nrow = (grp * kTRT + rt) * 16 + r;
    if (b < B && nrow < N) {
      if (a.mode == EPI_NONE) {
        float o = f[r * 8 + b];
        if (a.bias) o += __bfloat162float(a.bias[nrow]);
        if (a.residual) o = round_bf16(o) + __bfloat162float(a.residual[(long)b * a.ldr + nrow]);
        if (a.out_fp32) reinterpret_cast<float*>(a.y)[(long)b * a.ldy + nrow] = o;
        else reinterpret_cast<bf16*>(a.y)[(long)b * a.ldy + nrow] = __float2bfloat16_rn(o);
      } else if (a.mode == EPI_SWIGLU) {
        if (!(r & 1)) {
          const float gate = round_bf16(f[r * 8 + b]), up = round_bf16(f[(r + 1) * 8 + b]);
          reinterpret_cast<bf16*>(a.y)[(long)b * a.ldy + (nrow >> 1)] = __float2bfloat16_rn(round_bf16(silu(gate)) * up);
        }
      } else {  // GEMV_ROPE_QKV
        const int D = a.head_dim, H = a.n_heads;
        const int hh = nrow / D, i = nrow - hh * D;
        const int slot = a.pos[b];
        if (hh < 2 * H) {
          if (!(r & 1)) {
            const float x1 = round_bf16(f[r * 8 + b]), x2 = round_bf16(f[(r + 1) * 8 + b]);
            const int rp = slot - (a.pos_off ? a.pos_off[b] : 0);
            const float c = __bfloat162float(a.rope_cos[(long)rp * (D / 2) + (i >> 1)]);
            const float s = __bfloat162float(a.rope_sin[(long)rp * (D / 2) + (i >> 1)]);
            const float o1 = round_bf16(x1 * c) + round_bf16(-x2 * s);
            const float o2 = round_bf16(x2 * c) + round_bf16(x1 * s);
            bf16* dst;
            if (hh < H) dst = reinterpret_cast<bf16*>(a.y) + (long)b * a.ldy + nrow;
            else dst = a.k_cache + (((long)b * H + (hh - H)) * a.t_max + slot) * D + i;
            *reinterpret_cast<uint32_t*>(dst) = pack_bf16(o1, o2);
          }
        } else {
          a.v_cache[(((long)b * H + (hh - 2 * H)) * a.t_max + slot) * D + i] = __float2bfloat16_rn(f[r * 8 + b]);
        }
      }
    }
  }
  consumer_bar();
}

__global__ void __launch_bounds__(kTThreads, 2) gemv_tma_kernel(const __grid_constant__ CUtensorMap tmW,
                                                                const GemvTmaParams p) {
  const GemvArgs& a = p.a;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                                                   // kTStages x 16 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + p.nstages * kTStageBytes);
  uint64_t* empty_bar = full_bar + kTStages;
  float* red = reinterpret_cast<float*>(empty_bar + kTStages);            // [8 warps][kTRT*128]
  float* fin = red + kTW * kTRT * 128;                                    // [kTRT*128]
  bf16* xs = reinterpret_cast<bf16*>(fin + kTRT * 128);                   // [B][ldxs]
  __shared__ float s_ss[kTW][8];
  __shared__ float s_rstd[8];
  __shared__ int s_last;
  __shared__ GemvTmaParams s_params;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K, B = a.B, CPT = p.cpt;
  const long G = gridDim.x;
  const long c0 = (long)blockIdx.x * p.total / G, c1 = ((long)blockIdx.x + 1) * p.total / G;
  const int n = (int)(c1 - c0);

  if (threadIdx.x == 0) {
    s_params = p;
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < kTStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kTW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (a.pdl) pdl_launch_dependents();
  if (warp == kTW) {
    // ===================== producer: weights do not depend on the previous kernel =====================
    if (lane == 0) {
      int grp = (int)(c0 / CPT), kc = (int)(c0 % CPT);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* dst = ring + stage * kTStageBytes;
        mbar_expect_tx(&full_bar[stage], kTStageBytes);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tma_load_2d(dst + j * (kTRows * 128), &tmW, &full_bar[stage], kc * kTCols + j * 64, grp * kTRows);
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        if (++kc == CPT) { kc = 0; ++grp; }
      }
    }
    return;
  }

  // ===================== consumers (threads 0..255) =====================
  if (a.pdl) pdl_wait();
  const int g = lane >> 2, t = lane & 3;
  // ---- stage x (optionally RMS-normalised), zero-padded to kpad columns ----
  const int vec_per_row = K >> 3, vec_pad = p.kpad >> 3;
  if (a.norm_w != nullptr) {
    float ss[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) ss[b] = 0.f;
    for (int b = 0; b < B; ++b) {
      const uint4* src = reinterpret_cast<const uint4*>(a.x + (long)b * a.ldx);
      float s = 0.f;
      for (int i = threadIdx.x; i < vec_per_row; i += 256) {
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
    consumer_bar();
    if (threadIdx.x < 8) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kTW; ++w) tot += s_ss[w][threadIdx.x];
      s_rstd[threadIdx.x] = rsqrtf(tot / (float)K + a.norm_eps);
    }
    consumer_bar();
  }
  for (int b = 0; b < B; ++b) {
    const float rstd = a.norm_w ? s_rstd[b] : 1.f;
    const uint4* src = reinterpret_cast<const uint4*>(a.x + (long)b * a.ldx);
    const uint4* wsrc = reinterpret_cast<const uint4*>(a.norm_w);
    uint4* dst = reinterpret_cast<uint4*>(xs + (long)b * p.ldxs);
    for (int i = threadIdx.x; i < vec_pad; i += 256) {
      uint4 o = make_uint4(0, 0, 0, 0);
      if (i < vec_per_row) {
        const uint4 v = src[i];
        if (a.norm_w) {
          const uint4 w = wsrc[i];
          const uint32_t v4[4] = {v.x, v.y, v.z, v.w}, w4[4] = {w.x, w.y, w.z, w.w};
          uint32_t o4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)  // HF: weight * (x.float() * rsqrt(var + eps)).to(bf16)
            o4[j] = pack_bf16(round_bf16(bf16_lo(v4[j]) * rstd) * bf16_lo(w4[j]),
                              round_bf16(bf16_hi(v4[j]) * rstd) * bf16_hi(w4[j]));
          o = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        } else {
          o = v;
        }
      }
      dst[i] = o;
    }
  }
  consumer_bar();

  float acc[kTRT][4];
#pragma unroll
  for (int rt = 0; rt < kTRT; ++rt)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[rt][q] = 0.f;

  // fragment addressing inside a 128B-swizzled [32 rows x 64 cols] TMA tile
  const int tile_j = warp >> 1;                  // which 64-column tile of the chunk this warp reads
  const int cbase = (warp & 1) * 4;              // first logical 16-byte chunk of this warp's 32 columns
  const int lrow = lane & 15, lhalf = lane >> 4;  // ldmatrix.x4 lane -> (row, k-half)
  const bool has_x = g < B;
  const bf16* xrow = xs + (long)g * p.ldxs + warp * 32 + 2 * t;

  int cp_grp = (int)(c0 / CPT), cp_kc = (int)(c0 % CPT);
  int seg_lo = cp_kc;
  int stage = 0;
  uint32_t phase = 0;
  for (int i = 0; i < n; ++i) {
    mbar_wait(&full_bar[stage], phase);
    const uint32_t tbase = smem_u32(ring + stage * kTStageBytes + tile_j * (kTRows * 128));
    const bf16* xk = xrow + cp_kc * kTCols;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t bfr[2] = {0u, 0u};
      if (has_x) {
        bfr[0] = *reinterpret_cast<const uint32_t*>(xk + ks * 16);
        bfr[1] = *reinterpret_cast<const uint32_t*>(xk + ks * 16 + 8);
      }
#pragma unroll
      for (int rt = 0; rt < kTRT; ++rt) {
        const int r = rt * 16 + lrow;
        const int c = cbase + ks * 2 + lhalf;
        uint32_t af[4];
        ldmatrix_x4(af, tbase + r * 128 + ((c ^ (r & 7)) << 4));
        mma_bf16_16816(acc[rt], af, bfr);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (++stage == p.nstages) { stage = 0; phase ^= 1; }
    const bool grp_done = (cp_kc == CPT - 1) || (i == n - 1);
    if (grp_done) {
#pragma unroll
      for (int rt = 0; rt < kTRT; ++rt) {
        float* r = red + (warp * kTRT + rt) * 128;
        r[g * 8 + 2 * t] = acc[rt][0];
        r[g * 8 + 2 * t + 1] = acc[rt][1];
        r[(g + 8) * 8 + 2 * t] = acc[rt][2];
        r[(g + 8) * 8 + 2 * t + 1] = acc[rt][3];
        acc[rt][0] = acc[rt][1] = acc[rt][2] = acc[rt][3] = 0.f;
      }
      gemv_tma_flush(&s_params, red, fin, &s_last, cp_grp, seg_lo, cp_kc);
    }
    if (++cp_kc == CPT) { cp_kc = 0; ++cp_grp; }
    if (grp_done) seg_lo = cp_kc;
  }
}

static float* g_tws = nullptr;
static int* g_tcounters = nullptr;

int gemv_tma_init() {
  if (g_tws) return EMU_OK;
  if (cudaMalloc((void**)&g_tws, (size_t)kTWsGroups * kTMaxParts * kTRT * 128 * sizeof(float)) != cudaSuccess) return EMU_ERR_NOMEM;
  if (cudaMalloc((void**)&g_tcounters, (size_t)kTWsGroups * sizeof(int)) != cudaSuccess) return EMU_ERR_NOMEM;
  if (cudaMemset(g_tcounters, 0, (size_t)kTWsGroups * sizeof(int)) != cudaSuccess) return EMU_ERR_CUDA;
  if (cudaFuncSetAttribute(gemv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
    return EMU_ERR_CUDA;
  // ask for the largest shared-memory carve-out so that this kernel's CTA and its PDL successor's CTA (2 x ~105 KB)
  // can be resident on one SM at the same time
  if (cudaFuncSetAttribute(gemv_tma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) !=
      cudaSuccess)
    return EMU_ERR_CUDA;
  return EMU_OK;
}

// returns EMU_ERR_UNSUPPORTED when the shape does not fit this kernel (caller falls back to gemv.cu)
int gemv_tma_bf16(const GemvArgs& a, cudaStream_t st) {
  if (a.K % 8 || (reinterpret_cast<uintptr_t>(a.W) & 15)) return EMU_ERR_UNSUPPORTED;
  const int groups = (a.N + kTRows - 1) / kTRows;
  if (groups > kTWsGroups) return EMU_ERR_UNSUPPORTED;
  if (!g_tws) {  // first use must happen outside stream capture (the engine calls gemv_init() at create)
    int rc = gemv_tma_init();
    if (rc) return rc;
  }
  GemvTmaParams p;
  p.a = a;
  p.kpad = (a.K + kTCols - 1) / kTCols * kTCols;
  p.ldxs = p.kpad + 8;  // row stride = 16 B (mod 128 B): the 8 batch rows hit distinct bank groups
  p.cpt = p.kpad / kTCols;
  p.total = (long)groups * p.cpt;
  p.ws = g_tws;
  p.counters = g_tcounters;
  const size_t xs_bytes = (size_t)a.B * p.ldxs * 2;
  // the ring buffers are carved for kTStages; a CTA with a wide x uses fewer so that this kernel (<= ~110 KB) and its
  // PDL successor still fit one SM together
  static int env_stages = -1;
  if (env_stages < 0) {
    const char* v = getenv("EMU_GEMV_STAGES");
    env_stages = v ? atoi(v) : 0;
  }
  // default 8 stages = 128 KB in flight per SM.  Measured (profiles/r01_kernel_bench_tma_stages.txt): 8 stages reach
  // 6.3 TB/s on the large projections and 182 us per decoder layer chained, 5 stages (which would let this CTA and its
  // PDL successor co-reside) only 200 us: depth of prefetch beats co-residency
  p.nstages = env_stages > 0 ? env_stages : 8;
  if (p.nstages > kTStages) p.nstages = kTStages;
  size_t smem;
  for (;;) {  // shrink the ring until the CTA fits (wide x at batch > 1)
    smem = 1024 + (size_t)p.nstages * kTStageBytes + 2 * kTStages * 8 + (size_t)(kTW + 1) * kTRT * 128 * 4 + xs_bytes + 64;
    if (smem <= 200 * 1024 || p.nstages <= 3) break;
    --p.nstages;
  }
  if (smem > 200 * 1024) return EMU_ERR_UNSUPPORTED;
  long grid = kNumSMs;
  if (grid > p.total) grid = p.total;
  const long max_grid = (long)groups * (kTMaxParts - 3);
  if (grid > max_grid) grid = max_grid;
  CUtensorMap tm;
  if (make_tmap_2d(&tm, a.W, a.N, a.K, a.K, kTRows) != EMU_OK) return EMU_ERR_UNSUPPORTED;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kTThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemv_tma_kernel, tm, p) == cudaSuccess ? EMU_OK : EMU_ERR_CUDA;
}

}  // namespace emu
