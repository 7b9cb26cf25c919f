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
//                 32-column slabs), fused bias / GELU / residual / SwiGLU / GEGLU.  bf16 outputs are STAGED: the tile is
//                 assembled in shared memory as 64B-swizzled [128 rows x 32 columns] slabs (one conflict-free 16-byte
//                 st.shared per 8 outputs) and leaves with TMA stores; a residual tile arrives the same way (TMA load issued
//                 while the MMAs still run).  Round 1 stored / loaded rows straight from registers — one row per lane, so
//                 every 16-byte access of a warp touched 32 different cache lines: 3.3 k cycles (plain) to 7.6 k (residual)
//                 to 13.8 k (GEGLU) per 128 x 160..224 tile against a 9.7 k-cycle main loop (profiles/r02_gemm_phases_*.txt).
//                 fp32 outputs, unaligned or very wide (> 192 columns) tiles keep the direct path.
// A and W are both K-major, so neither operand needs a transpose anywhere in the model.
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "ops.h"

namespace emu {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_u32(smem)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}

template <int BN>
struct GemmSmem {
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "BN: multiple of 32 (epilogue chunks) within the tcgen05 N range");
  static constexpr int kStageBytes = (BM + BN) * BK * 2;
  // staging for the TMA-store epilogue: [slabs of 32 output columns][128 rows][64 B].  Tiles wider than 192 columns are
  // only staged by the pair epilogues (SwiGLU / GEGLU), which emit BN / 2 columns.
  static constexpr int kStagingBytes = BM * (BN <= 192 ? BN : BN / 2) * 2;
  static constexpr int kFit = (227 * 1024 - 1024 - 512 - kStagingBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;  // 8 / 7 / 6 / 5 / 4 / 4 / 4 for BN = 64 .. 256
  static constexpr int kBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(kStages >= 3, "pipeline too shallow");
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
  int staged;    // bf16 output through shared memory + TMA store (tmC)
  int res_smem;  // residual tile through TMA load into the staging buffer (tmR); else direct global loads
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
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  constexpr int kStages = GemmSmem<BN>::kStages;
  constexpr int kStageBytes = GemmSmem<BN>::kStageBytes;
  constexpr uint32_t kTmemCols = GemmSmem<BN>::kTmemCols;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * kStageBytes;  // 1024-aligned: kStageBytes is a multiple of 4096
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + GemmSmem<BN>::kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint64_t* res_full = tmem_empty + 2;         // [2] residual slabs of epilogue group 0 / 1 have landed
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_full + 2);

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
    if (p.staged) tma_prefetch_desc(&tmC);
    if (p.res_smem) tma_prefetch_desc(&tmR);
  }
  if (threadIdx.x < kStages) {  // barrier inits spread over threads (one thread walking ~22 of them costs ~0.3 us per launch)
    mbar_init(&full_bar[threadIdx.x], 1);
    mbar_init(&empty_bar[threadIdx.x], CL);  // one tcgen05.commit arrival from every CTA that reads the multicast stage
    mbar_fence_init();
  } else if (threadIdx.x >= 32 && threadIdx.x < 34) {
    const int i = threadIdx.x - 32;
    mbar_init(&tmem_full[i], 1);
    mbar_init(&tmem_empty[i], 8);
    mbar_init(&res_full[i], 1);
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
    if (p.staged) {
      // ---------- staged epilogue: registers -> 64B-swizzled shared-memory slabs -> TMA store ----------
      // The two warp groups (half = 0 / 1, four warps = the four TMEM lane quarters each) own the even / odd 32-column
      // output slabs of the tile and run independently: own named barrier, own elected issuer thread (TMA stores, the
      // residual loads of the next tile), own residual mbarrier.
      const bool pair = (p.epi == EPI_SWIGLU || p.epi == EPI_GEGLU);
      const int n_out = pair ? (p.N >> 1) : p.N;            // output columns of the whole matrix
      const int bn_out = pair ? BN / 2 : BN;                // ... of one tile
      const int n_units = bn_out / 32;
      const int r_in_tile = q * 32 + lane;
      const uint32_t stg = smem_u32(staging);
      const uint32_t my_row = (uint32_t)r_in_tile * 64u;
      const uint32_t swz = (uint32_t)((r_in_tile >> 1) & 3);
      const bool issuer = (q == 0 && lane == 0);
      const int bar_id = 1 + half;
      const bool vec_bias = p.bias != nullptr && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
      const bool vec_b2 = p.bias2 != nullptr && (p.N % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.bias2) & 15) == 0);
      const bool vec_res = p.residual != nullptr && (p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
      auto tile_row0 = [&](int tm) -> long {
        if (!p.conv) return (long)tm * BM;
        const int tiles_w = p.W / p.tw, tiles_h = p.H / p.th;
        const int w0 = (tm % tiles_w) * p.tw, h0 = ((tm / tiles_w) % tiles_h) * p.th, img = tm / (tiles_w * tiles_h);
        return ((long)img * p.H + h0) * p.W + w0;  // a tile is th full-width rows or one 128-pixel run: 128 consecutive rows
      };
      auto load_residual = [&](int tile) {  // issuer only: this group's residual slabs of `tile` -> staging
        const int tm = (tile % units_m) * CL + crank, tn = tile / units_m;
        int cnt = 0;
        for (int u = half; u < n_units; u += 2)
          if (tn * bn_out + u * 32 < n_out) ++cnt;
        if (tm >= tiles_m) cnt = 0;
        mbar_expect_tx(&res_full[half], (uint32_t)cnt * 8192u);  // arrive + expect: with 0 bytes the phase completes at once
        if (cnt == 0) return;
        const int row0 = (int)tile_row0(tm);
        for (int u = half; u < n_units; u += 2)
          if (tn * bn_out + u * 32 < n_out) tma_load_2d(staging + u * 8192, &tmR, &res_full[half], tn * bn_out + u * 32, row0);
      };
      if (p.res_smem && issuer && unit0 < num_tiles) load_residual(unit0);
      uint32_t res_phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        const int tm = (tile % units_m) * CL + crank, tn = tile / units_m;
        const long row0 = tile_row0(tm);
        const long row = row0 + r_in_tile;
        const bool row_ok = row < p.M && tm < tiles_m;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (dbg && tile == unit0 && threadIdx.x == 64) dbg[6] = clk64();
        if (p.res_smem) mbar_wait(&res_full[half], res_phase);
#pragma unroll 1
        for (int u = half; u < n_units; u += 2) {
          const int oc0 = tn * bn_out + u * 32;  // first output column of this slab
          if (oc0 >= n_out) break;
          const uint32_t slab = stg + (uint32_t)u * 8192u + my_row;
          if (!pair) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + u * 32), v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            const bool full = oc0 + 32 <= p.N;
            if (p.bias != nullptr) {
              if (full && vec_bias) {
                const uint4* bsrc = reinterpret_cast<const uint4*>(p.bias + oc0);
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
                  if (oc0 + i < p.N) f[i] += __bfloat162float(p.bias[oc0 + i]);
              }
            }
            if (p.bias2 != nullptr && row_ok) {
              const bf16* b2 = p.bias2 + (row / p.bias2_rows) * p.N + oc0;
              if (full && vec_b2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint4 b = __ldg(reinterpret_cast<const uint4*>(b2) + i);
                  const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    round_bf16x2(f[8 * i + 2 * j], f[8 * i + 2 * j + 1]);
                    f[8 * i + 2 * j] += bf16_lo(bw[j]);
                    f[8 * i + 2 * j + 1] += bf16_hi(bw[j]);
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (oc0 + i < p.N) f[i] = round_bf16(f[i]) + __bfloat162float(b2[i]);
              }
            }
            if (p.epi == EPI_GELU) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                round_bf16x2(f[i], f[i + 1]);
                f[i] = gelu_erf_fast(f[i]);
                f[i + 1] = gelu_erf_fast(f[i + 1]);
              }
            } else if (p.epi == EPI_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (p.res_smem) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 r = lds128(slab + (((uint32_t)i ^ swz) << 4));
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  round_bf16x2(f[8 * i + 2 * j], f[8 * i + 2 * j + 1]);
                  f[8 * i + 2 * j] += bf16_lo(rw[j]);
                  f[8 * i + 2 * j + 1] += bf16_hi(rw[j]);
                }
              }
            } else if (p.residual != nullptr && row_ok) {
              const bf16* rsd = p.residual + row * p.ldr + oc0;
              if (full && vec_res) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint4 r = __ldg(reinterpret_cast<const uint4*>(rsd) + i);
                  const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    round_bf16x2(f[8 * i + 2 * j], f[8 * i + 2 * j + 1]);
                    f[8 * i + 2 * j] += bf16_lo(rw[j]);
                    f[8 * i + 2 * j + 1] += bf16_hi(rw[j]);
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (oc0 + i < p.N) f[i] = round_bf16(f[i]) + __bfloat162float(rsd[i]);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 w;
              w.x = pack_bf16(f[8 * i], f[8 * i + 1]);
              w.y = pack_bf16(f[8 * i + 2], f[8 * i + 3]);
              w.z = pack_bf16(f[8 * i + 4], f[8 * i + 5]);
              w.w = pack_bf16(f[8 * i + 6], f[8 * i + 7]);
              sts128(slab + (((uint32_t)i ^ swz) << 4), w);
            }
          } else {
            // interleaved (a_j, b_j) accumulator column pairs -> one output column: slab u <- accumulator chunks 2u, 2u+1
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
              const int ac0 = tn * BN + (2 * u + hc) * 32;  // first accumulator column of this chunk
              uint32_t v[32];
              tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + (2 * u + hc) * 32), v);
              tmem_ld_wait();
              float f[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
              if (p.bias != nullptr) {
                if (ac0 + 32 <= p.N && vec_bias) {
                  const uint4* bsrc = reinterpret_cast<const uint4*>(p.bias + ac0);
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
                    if (ac0 + i < p.N) f[i] += __bfloat162float(p.bias[ac0 + i]);
                }
              }
              // rounding points of the reference (Linear -> bf16, activation -> bf16, product -> bf16), two values per
              // conversion: (a_i, b_i) together, then the activations of two neighbouring outputs together
              float o[16];
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                round_bf16x2(f[2 * i], f[2 * i + 1]);
                round_bf16x2(f[2 * i + 2], f[2 * i + 3]);
                float g0, g1;
                if (p.epi == EPI_SWIGLU) { g0 = silu(f[2 * i]); g1 = silu(f[2 * i + 2]); }
                else { g0 = gelu_erf_fast(f[2 * i + 1]); g1 = gelu_erf_fast(f[2 * i + 3]); }
                round_bf16x2(g0, g1);
                if (p.epi == EPI_SWIGLU) { o[i] = g0 * f[2 * i + 1]; o[i + 1] = g1 * f[2 * i + 3]; }
                else { o[i] = f[2 * i] * g0; o[i + 1] = f[2 * i + 2] * g1; }
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                uint4 w;
                w.x = pack_bf16(o[8 * i], o[8 * i + 1]);
                w.y = pack_bf16(o[8 * i + 2], o[8 * i + 3]);
                w.z = pack_bf16(o[8 * i + 4], o[8 * i + 5]);
                w.w = pack_bf16(o[8 * i + 6], o[8 * i + 7]);
                sts128(slab + (((uint32_t)(hc * 2 + i) ^ swz) << 4), w);
              }
            }
          }
        }
        // the accumulator is drained: hand the TMEM stage back before the stores leave
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        fence_async_smem();        // generic-proxy slab writes -> visible to the TMA engine
        named_bar(bar_id, 128);
        if (issuer) {
          if (tm < tiles_m) {
            for (int u = half; u < n_units; u += 2) {
              const int oc0 = tn * bn_out + u * 32;
              if (oc0 < n_out) tma_store_2d(&tmC, staging + u * 8192, oc0, (int)row0);
            }
          }
          bulk_commit();
          bulk_wait_read0();       // the slabs may be overwritten once the stores have READ them
          if (p.res_smem && tile + unit_step < num_tiles) load_residual(tile + unit_step);
        }
        named_bar(bar_id, 128);    // nobody of the group touches the slabs before the issuer got here
        if (dbg && tile == unit0 && threadIdx.x == 64) dbg[7] = clk64();
        res_phase ^= 1;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (issuer) bulk_wait0();    // all stores complete (global writes performed) before the CTA retires
    } else
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

// bf16 [rows, cols] output / residual matrix for the staged epilogue: box = 128 rows x 32 columns (64 B), 64B swizzle.
// TMA clips the rows / columns of a box that fall outside the matrix on a store and zero-fills them on a load.
int make_tmap_out(CUtensorMap* out, const void* base, long rows, long cols, long ld) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return EMU_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
static int launch_gemm_cl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                          const GemmParams& p, cudaStream_t st) {
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
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CL>, tmA, tmB, tmC, tmR, p) == cudaSuccess ? EMU_OK : EMU_ERR_CUDA;
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

// Can this problem use the staged (shared memory + TMA store) epilogue at tile width bn?
static bool can_stage(const GemmEpilogue& e, int N, int bn) {
  if (e.out_fp32 || e.C == nullptr) return false;
  const bool pair = e.mode == EPI_SWIGLU || e.mode == EPI_GEGLU;
  const int n_out = pair ? N / 2 : N;
  if ((reinterpret_cast<uintptr_t>(e.C) & 15) || (e.ldc % 8) || (n_out % 8)) return false;
  if (pair) return bn % 64 == 0;  // a 32-column output slab = 64 accumulator columns
  return bn <= 192;               // the staging buffer holds at most 192 columns
}
static bool env_no_stage() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("EMU_GEMM_DIRECT");
    v = (s && atoi(s) == 1) ? 1 : 0;
  }
  return v == 1;
}

// Tile width: minimise the modelled time of one CTA's tile stream over the instantiated widths.  Per tile the main loop
// costs kblocks x max(tensor pipe: 2 x bn cycles per 64-deep k block, L2 -> shared-memory operand traffic: (128 + bn) x 128 B at
// ~75 B/cycle/SM) and the epilogue, which overlaps the NEXT tile's main loop (double-buffered TMEM), costs per output
// column ~6 cycles staged and 21 (plain) / 47 (residual) / 60 (GELU, GEGLU) direct — all measured with
// emu_debug_gemm_phases (profiles/r02_gemm_phases_*.txt).  Odd widths such as 160 exist because e.g. M=2048, N=1280 is 160
// tiles at BN=128 (two waves on 148 SMs, the second 8 % full) but 128 tiles at BN=160 (one wave).
static int pick_bn(int M, int N, int K, const GemmEpilogue& e) {
  static const int cand[] = {256, 224, 192, 160, 128, 96, 64};
  const long tm = (M + BM - 1) / BM;
  const long kb = (K + BK - 1) / BK;
  const bool act = e.mode == EPI_GELU || e.mode == EPI_SWIGLU || e.mode == EPI_GEGLU;
  int best = 128;
  double best_cost = 1e30;
  for (int bn : cand) {
    const long tiles = tm * ((N + bn - 1) / bn);
    const long waves = (tiles + kNumSMs - 1) / kNumSMs;
    const double a_rows = M < 128 ? (double)M : 128.0;  // rows past M are zero-filled by TMA, not fetched
    const double mma = 2.0 * bn, l2 = (a_rows + bn) * 128.0 / 75.0;
    double per_kb = mma > l2 ? mma : l2;
    if (tm == 1) {  // one row of tiles: every weight byte comes from HBM exactly once, shared by the busy SMs (~3370 B/clk)
      const double active = tiles < kNumSMs ? (double)tiles : (double)kNumSMs;
      const double hbm = bn * 128.0 * active / 3370.0;
      if (hbm > per_kb) per_kb = hbm;
    }
    const double ml = (double)kb * per_kb + 1500.0;  // + pipeline fill
    const bool staged = !env_no_stage() && can_stage(e, N, bn);
    double per_col = staged ? (act ? 20.0 : 6.0) : (act ? 60.0 : (e.residual ? 47.0 : 21.0));
    const double epi = per_col * bn + 400.0;
    const double cost = (double)(waves - 1) * (ml > epi ? ml : epi) + ml + epi;
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

// decide the epilogue style of this launch and build the tensor maps it needs (dummies otherwise: the kernel never touches
// a map whose flag is off)
static int setup_staging(GemmParams& p, const GemmEpilogue& e, int M, int N, int bn, CUtensorMap* tmC, CUtensorMap* tmR) {
  memset(tmC, 0, sizeof(*tmC));
  memset(tmR, 0, sizeof(*tmR));
  p.staged = 0;
  p.res_smem = 0;
  const bool forced_direct = (e.force_bn & 4096) != 0;
  if (forced_direct || env_no_stage() || !can_stage(e, N, bn)) return EMU_OK;
  const bool pair = e.mode == EPI_SWIGLU || e.mode == EPI_GEGLU;
  const int n_out = pair ? N / 2 : N;
  if (make_tmap_out(tmC, e.C, M, n_out, e.ldc) != EMU_OK) return EMU_OK;  // odd geometry: keep the direct path
  p.staged = 1;
  if (e.residual != nullptr && !pair && (e.ldr % 8 == 0) && !(reinterpret_cast<uintptr_t>(e.residual) & 15) &&
      make_tmap_out(tmR, e.residual, M, n_out, e.ldr) == EMU_OK)
    p.res_smem = 1;
  return EMU_OK;
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
  const int bn = (e.force_bn & 1023) ? (e.force_bn & 1023) : pick_bn(M, N, K, e);
  const bool cl = use_cluster(M, e.force_bn, false);
  CUtensorMap tmA, tmB, tmC, tmR;
  int rc = make_tmap_2d(&tmA, A, M, K, lda, BM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, W, N, K, ldw, cl ? bn / 2 : bn);
  if (rc) return rc;
  rc = setup_staging(p, e, M, N, bn, &tmC, &tmR);
  if (rc) return rc;
  return dispatch_bn(bn, [&](auto w) {
    constexpr int kBN = decltype(w)::value;
    return cl ? launch_gemm_cl<kBN, 2>(tmA, tmB, tmC, tmR, p, st) : launch_gemm_cl<kBN, 1>(tmA, tmB, tmC, tmR, p, st);
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
  const int bn = (e.force_bn & 1023) ? (e.force_bn & 1023) : pick_bn(p.M, Cout, 9 * Cin, e);
  const bool cl = use_cluster(p.M, e.force_bn, true);
  CUtensorMap tmA, tmB, tmC, tmR;
  int rc = make_tmap_nhwc(&tmA, X, NB, H, W, Cin, tw, th);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, Wk, Cout, 9L * Cin, 9L * Cin, cl ? bn / 2 : bn);
  if (rc) return rc;
  rc = setup_staging(p, e, p.M, Cout, bn, &tmC, &tmR);
  if (rc) return rc;
  return dispatch_bn(bn, [&](auto w) {
    constexpr int kBN = decltype(w)::value;
    return cl ? launch_gemm_cl<kBN, 2>(tmA, tmB, tmC, tmR, p, st) : launch_gemm_cl<kBN, 1>(tmA, tmB, tmC, tmR, p, st);
  });
}

}  // namespace emu
