// emu_b200 — tcgen05 GEMM:  C[M,N] = epilogue( A[M,K] · W[N,K]^T )
//
// The dense-contraction workhorse of the generate path: ViT QKV/proj/MLP (Emu2/emu/eva_vit.py:194-200,
// 105-114), LLaMA prefill q/k/v/o/gate/up/down (HF LlamaDecoderLayer, called from Emu2/emu/emu.py:133-138,
// 213-229), project_up/project_down (emu.py:53-55), UNet linears and — through the 4-D TMA "conv" A-loader —
// the UNet/VAE 3x3 convolutions (diffusers UNet2DConditionModel, called from Emu2/emu/diffusion.py:136-141).
//
// Design (one CTA per SM, persistent over output tiles):
//   warp 0      : TMA producer — cp.async.bulk.tensor loads of 128x64 (A) and BNx64 (W) bf16 tiles into a
//                 kStages-deep ring of 128B-swizzled shared-memory stages, completion on mbarriers
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, M=128, N=BN, K=16),
//                 fp32 accumulators double-buffered in TMEM so the epilogue of tile i overlaps tile i+1
//   warps 2..9  : epilogue — tcgen05.ld accumulator rows to registers (two warps per TMEM lane quarter, alternating
//                 32-column chunks), fused bias / GELU / residual / SwiGLU / GEGLU with 16-byte operand loads
//                 prefetched ahead of the accumulator, bf16 or fp32 stores
// A and W are both K-major, so neither operand needs a transpose anywhere in the model.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ops.h"

namespace emu {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps

template <int BN>
struct GemmSmem {
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "BN: multiple of 32 (epilogue chunks) within the tcgen05 N range");
  static constexpr int kStageBytes = (BM + BN) * BK * 2;
  static constexpr int kFit = (227 * 1024 - 1024 - 256) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;  // 8 / 8 / 7 / 6 / 5 / 5 / 4 for BN = 64 .. 256
  static constexpr int kBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  // two accumulator stages; tcgen05.alloc wants a power of two
  static constexpr uint32_t kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
};

struct GemmParams {
  int M, N, K;
  // epilogue
  void* C;         // bf16 or fp32
  int ldc;         // elements
  const bf16* bias;      // [N] or null
  const bf16* residual;  // [M, ldr] or null (added AFTER rounding the linear output to bf16, like `x + lin(x)`)
  int ldr;
  const bf16* bias2;     // [M / bias2_rows, N] or null: per-row-group bias (ResnetBlock2D time embedding), added after
  int bias2_rows;        //   rounding like `conv(x) + temb[:, :, None, None]`
  int epi;         // EpiMode
  int out_fp32;
  // conv A-loader (mode 1): A is an NHWC tensor [NB, H, W, Cin]; M = NB*H*W output pixels (stride 1, pad 1),
  // K index = tap * Cin + c.  tile rows = th x tw spatial patch (th*tw == 128)
  int conv;        // 0 = plain 2-D A, 1 = 3x3 conv (pad 1), 2 = 3x3 conv stride 2 is NOT handled here
  int H, W, Cin, tw, th;
  int pdl;  // launched with programmatic dependent launch: griddepcontrol.wait before touching activations
  // diagnostics (emu_debug_gemm_phases): when non-null, every CTA writes 8 x u64 = {globaltimer at entry, clock64 at entry,
  // after set-up, first TMA issued, first stage landed (MMA side), last MMA committed, epilogue released by the MMAs,
  // epilogue done} for its FIRST tile
  unsigned long long* dbg;
};
__device__ __forceinline__ unsigned long long clk64() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// CL = thread-block cluster size along M (1 or 2).  With CL == 2 the two CTAs of a cluster work on vertically adjacent
// output tiles (same weight columns): each loads HALF of the W tile and TMA-multicasts it into both CTAs' shared memory,
// so W crosses the L2->SM fabric once per pair — the fabric, not the tensor pipe, bounds the BN <= 160 tiles.
template <int BN, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int kStages = GemmSmem<BN>::kStages;
  constexpr int kStageBytes = GemmSmem<BN>::kStageBytes;
  constexpr uint32_t kTmemCols = GemmSmem<BN>::kTmemCols;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (p.M + BM - 1) / BM;
  const int tiles_n = (p.N + BN - 1) / BN;
  // work units: one tile (CL == 1) or one vertical tile pair (CL == 2; a ragged last pair computes a dummy tile whose
  // loads are out of bounds = zeros and whose rows fail the epilogue's row check, keeping the pair in lock step)
  const int units_m = (tiles_m + CL - 1) / CL;
  const int num_tiles = units_m * tiles_n;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / CL, unit_step = gridDim.x / CL;
  const int kblocks_per_tap = p.conv ? (p.Cin + BK - 1) / BK : (p.K + BK - 1) / BK;
  const int num_kb = p.conv ? 9 * kblocks_per_tap : kblocks_per_tap;

  unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 8 : nullptr;
  if (dbg && threadIdx.x == 0) { dbg[0] = gtimer(); dbg[1] = clk64(); }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CL);  // one tcgen05.commit arrival from every CTA that reads the multicast stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    mbar_fence_init();
  }
  if (p.pdl) pdl_launch_dependents();  // the next kernel of the chain may start its own prologue now
  if (warp == 1) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // the peer's barriers must be initialised before anything is multicast at them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (dbg && threadIdx.x == 0) dbg[2] = clk64();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.pdl) pdl_wait();  // activations (A) come from the predecessor; everything above overlapped its tail
      if (dbg) dbg[3] = clk64();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        const int tm = (tile % units_m) * CL + crank, tn = tile / units_m;
        int img = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          const int tiles_w = p.W / p.tw, tiles_h = p.H / p.th;
          w0 = (tm % tiles_w) * p.tw;
          h0 = ((tm / tiles_w) % tiles_h) * p.th;
          img = tm / (tiles_w * tiles_h);
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + BM * BK * 2;
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          int kcol;
          if (p.conv) {
            const int tap = kb / kblocks_per_tap, cb = kb % kblocks_per_tap;
            const int r = tap / 3, s = tap % 3;
            tma_load_4d(sa, &tmA, &full_bar[stage], cb * BK, w0 + s - 1, h0 + r - 1, img);
            kcol = tap * p.Cin + cb * BK;
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, tm * BM);
            kcol = kb * BK;
          }
          if (CL > 1) {
            // my half of the W tile, written to the same offset in both CTAs; the peer supplies the other half
            constexpr int kHalf = BN / CL;
            tma_load_2d_mc(sb + crank * kHalf * BK * 2, &tmB, &full_bar[stage], kcol, tn * BN + crank * kHalf,
                           (uint16_t)((1u << CL) - 1));
          } else {
            tma_load_2d(sb, &tmB, &full_bar[stage], kcol, tn * BN);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (dbg && kb == 0 && tile == unit0) dbg[4] = clk64();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + BM * BK * 2;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (addr>>4) field
            umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          // frees the smem stage when these MMAs retire — in every CTA of the cluster (the peer's producer overwrites
          // half of OUR stage, so it must see our consumption too)
          if (CL > 1) umma_commit_mc(&empty_bar[stage], (uint16_t)((1u << CL) - 1));
          else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        if (dbg && tile == unit0) dbg[5] = clk64();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    // warp w owns TMEM lane quarter (w & 3) — the hardware restriction — and every second 32-column chunk
    // (chunk parity = (w - 2) >> 2), so two warps per SM sub-partition share a tile's epilogue.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    if (p.pdl) pdl_wait();  // residual / bias2 are predecessor outputs and C may alias a buffer it still reads
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_step) {
      const int tm = (tile % units_m) * CL + crank, tn = tile / units_m;
      // output row owned by this thread
      long row;
      const int r_in_tile = q * 32 + lane;
      if (p.conv) {
        const int tiles_w = p.W / p.tw, tiles_h = p.H / p.th;
        const int w0 = (tm % tiles_w) * p.tw;
        const int h0 = ((tm / tiles_w) % tiles_h) * p.th;
        const int img = tm / (tiles_w * tiles_h);
        const int hh = h0 + r_in_tile / p.tw, ww = w0 + r_in_tile % p.tw;
        row = ((long)img * p.H + hh) * p.W + ww;
      } else {
        row = (long)tm * BM + r_in_tile;
      }
      const bool row_ok = row < p.M && tm < tiles_m;
      const bool pair = (p.epi == EPI_SWIGLU || p.epi == EPI_GEGLU);
      // 16-byte paths need aligned rows; everything in the models is, ragged shapes take the scalar path
      const bool vec_res = p.residual != nullptr && (p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
      const bool vec_b2 = p.bias2 != nullptr && (p.N % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.bias2) & 15) == 0);
      const bool vec_bias = p.bias != nullptr && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
      // operands that do not depend on the accumulator are fetched BEFORE waiting for the MMAs of this tile
      uint4 rsd_v[4] = {}, b2_v[4] = {};
      auto prefetch = [&](int c) {
        const int col0 = tn * BN + c * 32;
        if (!row_ok || col0 + 32 > p.N) return;
        if (vec_res) {
          const uint4* src = reinterpret_cast<const uint4*>(p.residual + row * p.ldr + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i) rsd_v[i] = __ldg(src + i);
        }
        if (vec_b2) {
          const uint4* src = reinterpret_cast<const uint4*>(p.bias2 + (row / p.bias2_rows) * p.N + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i) b2_v[i] = __ldg(src + i);
        }
      };
      prefetch(half);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (dbg && tile == unit0 && threadIdx.x == 64) dbg[6] = clk64();
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        const int col0 = tn * BN + c * 32;
        uint4 rsd_c[4], b2_c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { rsd_c[i] = rsd_v[i]; b2_c[i] = b2_v[i]; }
        if (c + 2 < BN / 32) prefetch(c + 2);  // next chunk's operands fly while this one is processed
        if (row_ok && col0 < p.N) {
          const bool full = col0 + 32 <= p.N;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
            if (full && vec_bias) {
              const uint4* bsrc = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 b = __ldg(bsrc + i);
                const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { f[8 * i + 2 * j] += bf16_lo(bw[j]); f[8 * i + 2 * j + 1] += bf16_hi(bw[j]); }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) f[i] += __bfloat162float(p.bias[col0 + i]);
            }
          }
          if (p.bias2 != nullptr) {
            if (full && vec_b2) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t bw[4] = {b2_c[i].x, b2_c[i].y, b2_c[i].z, b2_c[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  f[8 * i + 2 * j] = round_bf16(f[8 * i + 2 * j]) + bf16_lo(bw[j]);
                  f[8 * i + 2 * j + 1] = round_bf16(f[8 * i + 2 * j + 1]) + bf16_hi(bw[j]);
                }
              }
            } else {
              const bf16* b2 = p.bias2 + (row / p.bias2_rows) * p.N + col0;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) f[i] = round_bf16(f[i]) + __bfloat162float(b2[i]);
            }
          }
          if (pair) {
            // interleaved (a_j, b_j) column pairs -> one output column j
            // SwiGLU: silu(gate)*up with HF's bf16 rounding points; GEGLU: hidden * gelu(gate)
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = round_bf16(f[2 * i]), b = round_bf16(f[2 * i + 1]);
              if (p.epi == EPI_SWIGLU) o[i] = round_bf16(silu(a)) * b;
              else o[i] = a * round_bf16(gelu_erf_fast(b));
            }
            const int oc0 = col0 >> 1;
            bf16* dst = reinterpret_cast<bf16*>(p.C) + row * p.ldc + oc0;
            if (oc0 + 16 <= (p.N >> 1) && (p.ldc % 8 == 0)) {
              uint4 w0, w1;
              w0.x = pack_bf16(o[0], o[1]); w0.y = pack_bf16(o[2], o[3]); w0.z = pack_bf16(o[4], o[5]); w0.w = pack_bf16(o[6], o[7]);
              w1.x = pack_bf16(o[8], o[9]); w1.y = pack_bf16(o[10], o[11]); w1.z = pack_bf16(o[12], o[13]); w1.w = pack_bf16(o[14], o[15]);
              reinterpret_cast<uint4*>(dst)[0] = w0;
              reinterpret_cast<uint4*>(dst)[1] = w1;
            } else {
              for (int i = 0; i < 16; ++i)
                if (oc0 + i < (p.N >> 1)) dst[i] = __float2bfloat16_rn(o[i]);
            }
          } else {
            if (p.epi == EPI_GELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = gelu_erf_fast(round_bf16(f[i]));
            } else if (p.epi == EPI_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (p.residual != nullptr) {
              if (full && vec_res) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint32_t rw[4] = {rsd_c[i].x, rsd_c[i].y, rsd_c[i].z, rsd_c[i].w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    f[8 * i + 2 * j] = round_bf16(f[8 * i + 2 * j]) + bf16_lo(rw[j]);
                    f[8 * i + 2 * j + 1] = round_bf16(f[8 * i + 2 * j + 1]) + bf16_hi(rw[j]);
                  }
                }
              } else {
                const bf16* rsd = p.residual + row * p.ldr + col0;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) f[i] = round_bf16(f[i]) + __bfloat162float(rsd[i]);
              }
            }
            if (p.out_fp32) {
              float* dst = reinterpret_cast<float*>(p.C) + row * p.ldc + col0;
              if (full && (p.ldc % 4 == 0)) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  reinterpret_cast<float4*>(dst)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
              } else {
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) dst[i] = f[i];
              }
            } else {
              bf16* dst = reinterpret_cast<bf16*>(p.C) + row * p.ldc + col0;
              if (full && (p.ldc % 8 == 0)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  uint4 w;
                  w.x = pack_bf16(f[8 * i], f[8 * i + 1]);
                  w.y = pack_bf16(f[8 * i + 2], f[8 * i + 3]);
                  w.z = pack_bf16(f[8 * i + 4], f[8 * i + 5]);
                  w.w = pack_bf16(f[8 * i + 6], f[8 * i + 7]);
                  reinterpret_cast<uint4*>(dst)[i] = w;
                }
              } else {
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) dst[i] = __float2bfloat16_rn(f[i]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (dbg && tile == unit0 && threadIdx.x == 64) dbg[7] = clk64();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();  // nobody leaves while the peer can still multicast into / arrive on its shared memory
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

// 2-D K-major bf16 matrix [rows, cols] with row stride ld (elements); box = box_rows x 64 cols, 128B swizzle
int make_tmap_2d(CUtensorMap* out, const void* base, long rows, long cols, long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return EMU_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EMU_OK : EMU_ERR_CUDA;
}

// 4-D NHWC bf16 activation [NB, H, W, C]; box = {64 ch, tw, th, 1}; out-of-bounds (the conv halo) reads as zero
int make_tmap_nhwc(CUtensorMap* out, const void* base, int NB, int H, int W, int C, int tw, int th) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return EMU_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)tw, (cuuint32_t)th, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EMU_OK : EMU_ERR_CUDA;
}

// strided [B, N, H, D] bf16 view (attention operands) as a 4-D map; box = {64 d, box_rows tokens, 1 head, 1 batch}.
// Dimensions are ordered by stride (head-before-token when heads are interleaved inside a token row, as in fused QKV
// outputs); *head_first tells the kernel which coordinate order to use.  Head-dim padding and rows past N read as zero.
int make_tmap_bnhd(CUtensorMap* out, const void* base, int D, long N, int H, int B, long ts, long hs, long bs, int box_rows,
                   int* head_first) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return EMU_ERR_CUDA;
  if ((ts % 8) || (hs % 8) || (bs % 8) || (reinterpret_cast<uintptr_t>(base) & 15)) return EMU_ERR_INVALID;
  const bool hf = hs < ts;
  *head_first = hf ? 1 : 0;
  cuuint64_t bstride = bs > 0 ? (cuuint64_t)bs * 2 : 16;
  cuuint64_t tstride = ts > 0 ? (cuuint64_t)ts * 2 : 16;
  cuuint64_t hstride = hs > 0 ? (cuuint64_t)hs * 2 : 16;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4];
  dims[0] = (cuuint64_t)D;
  box[0] = 64;
  if (hf) {
    dims[1] = (cuuint64_t)H; dims[2] = (cuuint64_t)N; strides[0] = hstride; strides[1] = tstride;
    box[1] = 1; box[2] = (cuuint32_t)box_rows;
  } else {
    dims[1] = (cuuint64_t)N; dims[2] = (cuuint64_t)H; strides[0] = tstride; strides[1] = hstride;
    box[1] = (cuuint32_t)box_rows; box[2] = 1;
  }
  dims[3] = (cuuint64_t)B;
  strides[2] = bstride;
  box[3] = 1;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EMU_OK : EMU_ERR_CUDA;
}

template <int BN, int CL>
static int launch_gemm_cl(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  static bool attr_set = false;
  static int max_ctas = kNumSMs;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::kBytes) !=
        cudaSuccess)
      return EMU_ERR_CUDA;
    if (CL > 1) {
      // how many clusters can be resident at once (GPC boundaries may leave an SM without a partner)
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(kNumSMs / CL * CL);
      q.blockDim = dim3(kGemmThreads);
      q.dynamicSmemBytes = GemmSmem<BN>::kBytes;
      cudaLaunchAttribute a[1];
      a[0].id = cudaLaunchAttributeClusterDimension;
      a[0].val.clusterDim.x = CL; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      q.attrs = a;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN, CL>, &q) == cudaSuccess && n > 0) max_ctas = n * CL;
      else cudaGetLastError();
      if (max_ctas > kNumSMs) max_ctas = kNumSMs / CL * CL;
    }
    attr_set = true;
  }
  const int units = (((p.M + BM - 1) / BM + CL - 1) / CL) * ((p.N + BN - 1) / BN);
  int grid = units * CL < max_ctas ? units * CL : max_ctas;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = GemmSmem<BN>::kBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (p.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CL>, tmA, tmB, p) == cudaSuccess ? EMU_OK : EMU_ERR_CUDA;
}

// Cluster mode (W tile TMA-multicast over a 2-CTA cluster).  Measured on B200 (profiles/r01_gemm_bench_cluster.txt):
// +8..15 % on the K >= 5760 implicit-GEMM convolutions, -2..8 % on the one-wave linear layers (cluster start-up and
// lock step cost more than the halved W traffic saves).  Default: convolutions only.  EMU_GEMM_CLUSTER = 0 (never),
// 1 (always), unset (convs); force bits: 1024 = off, 2048 = on (tests / A-B runs).
static bool use_cluster(int M, int force, bool is_conv) {
  static int env = -2;
  if (env == -2) {
    const char* v = getenv("EMU_GEMM_CLUSTER");
    env = v ? atoi(v) : -1;
  }
  if (force & 1024) return false;
  const int tiles_m = (M + BM - 1) / BM;
  if (force & 2048) return tiles_m >= 1;
  if (env == 0 || (env < 0 && !is_conv)) return false;
  return tiles_m >= 2 && (tiles_m % 2 == 0 || tiles_m >= 7);
}

// Tile width: minimise  waves(BN) x time-per-tile(BN)  over the instantiated widths.  Per K=16 step a tile costs
// max(tensor pipe: 128*BN/256 cycles, shared-memory operand reads: (128 + BN) * 32 B at 128 B/cycle); odd widths such
// as 160 exist because e.g. M=2048, N=1280 is 160 tiles at BN=128 (two waves on 148 SMs, the second 8 % full) but 128
// tiles at BN=160 (one wave).
static int pick_bn(int M, int N) {
  static const int cand[] = {256, 224, 192, 160, 128, 96, 64};
  const long tm = (M + BM - 1) / BM;
  int best = 128;
  double best_cost = 1e30;
  for (int bn : cand) {
    const long tiles = tm * ((N + bn - 1) / bn);
    const long waves = (tiles + kNumSMs - 1) / kNumSMs;
    const double mma = 128.0 * bn / 256.0, smem = (128.0 + bn) * 32.0 / 128.0;
    const double cost = (double)waves * (mma > smem ? mma : smem) + 4.0 * waves;  // + per-tile fixed overhead
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

template <typename F>
static int dispatch_bn(int bn, F&& f) {
  switch (bn) {
    case 256: return f(std::integral_constant<int, 256>());
    case 224: return f(std::integral_constant<int, 224>());
    case 192: return f(std::integral_constant<int, 192>());
    case 160: return f(std::integral_constant<int, 160>());
    case 128: return f(std::integral_constant<int, 128>());
    case 96: return f(std::integral_constant<int, 96>());
    case 64: return f(std::integral_constant<int, 64>());
  }
  return EMU_ERR_INVALID;
}

int gemm_bf16(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const GemmEpilogue& e,
              cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return EMU_ERR_INVALID;
  if ((lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return EMU_ERR_INVALID;
  const bool pair = e.mode == EPI_SWIGLU || e.mode == EPI_GEGLU;
  if (pair && (N & 1)) return EMU_ERR_INVALID;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.C = e.C; p.ldc = e.ldc; p.bias = e.bias; p.residual = e.residual; p.ldr = e.ldr;
  p.bias2 = e.bias2; p.bias2_rows = e.bias2_rows > 0 ? e.bias2_rows : 1;
  p.epi = e.mode; p.out_fp32 = e.out_fp32; p.conv = 0; p.pdl = g_pdl_chain; p.dbg = e.dbg;
  const int bn = (e.force_bn & 1023) ? (e.force_bn & 1023) : pick_bn(M, N);
  const bool cl = use_cluster(M, e.force_bn, false);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, A, M, K, lda, BM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, W, N, K, ldw, cl ? bn / 2 : bn);
  if (rc) return rc;
  return dispatch_bn(bn, [&](auto w) {
    constexpr int kBN = decltype(w)::value;
    return cl ? launch_gemm_cl<kBN, 2>(tmA, tmB, p, st) : launch_gemm_cl<kBN, 1>(tmA, tmB, p, st);
  });
}

// 3x3 stride-1 pad-1 convolution on NHWC bf16 as an implicit GEMM. Wk is [Cout, 9*Cin] with k = (r*3+s)*Cin + c.
int conv3x3_bf16(const bf16* X, int NB, int H, int W, int Cin, const bf16* Wk, int Cout, const GemmEpilogue& e,
                 cudaStream_t st) {
  if (Cin % 8) return EMU_ERR_INVALID;
  int tw = W >= 128 ? 128 : W;  // tile = th x tw pixels, th*tw = 128
  if (128 % tw) return EMU_ERR_UNSUPPORTED;
  int th = 128 / tw;
  if (H % th || W % tw) return EMU_ERR_UNSUPPORTED;
  GemmParams p{};
  p.M = NB * H * W; p.N = Cout; p.K = 9 * Cin;
  p.C = e.C; p.ldc = e.ldc; p.bias = e.bias; p.residual = e.residual; p.ldr = e.ldr;
  p.bias2 = e.bias2; p.bias2_rows = e.bias2_rows > 0 ? e.bias2_rows : 1;
  p.epi = e.mode; p.out_fp32 = e.out_fp32; p.pdl = g_pdl_chain; p.dbg = e.dbg;
  p.conv = 1; p.H = H; p.W = W; p.Cin = Cin; p.tw = tw; p.th = th;
  const int bn = (e.force_bn & 1023) ? (e.force_bn & 1023) : pick_bn(p.M, Cout);
  const bool cl = use_cluster(p.M, e.force_bn, true);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_nhwc(&tmA, X, NB, H, W, Cin, tw, th);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, Wk, Cout, 9L * Cin, 9L * Cin, cl ? bn / 2 : bn);
  if (rc) return rc;
  return dispatch_bn(bn, [&](auto w) {
    constexpr int kBN = decltype(w)::value;
    return cl ? launch_gemm_cl<kBN, 2>(tmA, tmB, p, st) : launch_gemm_cl<kBN, 1>(tmA, tmB, p, st);
  });
}

}  // namespace emu
