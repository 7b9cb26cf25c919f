// emu_b200 — persistent decode-step kernel: ONE cooperative launch runs a whole autoregressive step of the LLaMA decoder
// (embed -> 60 x [RMSNorm+QKV+RoPE+KV-append, attention, o_proj+residual, RMSNorm+gate/up+SwiGLU, down+residual] ->
// final norm -> lm_head -> argmax), i.e. what HF LlamaModel.forward + lm_head do for one token inside
// `lm.generate` (Emu2/emu/emu.py:213-229) and inside the cache-equivalent generate_image steps (emu.py:109-147).
//
// Why: with one kernel per projection the dependency "x = f(all outputs of the previous kernel)" forces HBM to idle
// for the drain + x-staging of every launch (~4 us x 300 launches per token, measured 12.8 ms/token against 10.9 ms of
// pure streaming).  Here each SM runs one persistent CTA:
//   * the producer thread walks the step's weight matrices in order and keeps a 128 KB TMA ring full — it never
//     stops at phase boundaries, so the next projection's weights stream in while consumers synchronise;
//   * the 8 consumer warps execute the phases, separated by a grid-wide barrier (atomic counter + fence), stage the
//     freshly produced activations, drain the ring with ldmatrix + mma.sync and finish row groups stream-K style
//     (device code shared with the one-GEMV kernel: gemv_tma.cuh);
//   * decode attention runs as a phase on the same CTAs (split-KV items, last-arriver combine).
// Activations produced by other CTAs inside the kernel are always read with ld.global.cg (L2), never through the
// non-coherent path.
#include <string.h>

#include <map>
#include <tuple>
#include <vector>

#include "engine.h"
#include "gemv_tma.cuh"

namespace emu {

enum { MP_GEMV = 0, MP_ATTN = 1, MP_EMBED = 2, MP_NORM_OUT = 3, MP_FINAL = 4 };

struct MegaPhase {
  int type;
  int tmap;         // MP_GEMV: index of the weight's tensor map
  GemvTmaParams g;  // MP_GEMV: a.W unused (TMA), everything else as in gemv_tma.cu
  const bf16 *q, *kc, *vc;  // MP_ATTN (this layer's q buffer and cache slabs)
  bf16* out;
};

struct MegaCtl {
  const CUtensorMap* maps;
  const MegaPhase* phases;
  int nphases;
  unsigned* bar;  // grid barrier counter, zeroed before every launch
  int B, H, D, t_max, nsplit;
  const int* pos;
  const int* start;
  float scale;
  float* attn_ws;  // [B*H*nsplit][D] partial outputs, then [B*H*nsplit][2] (m, l)
  int* attn_counters;
  const bf16* embed_table;
  const int* token_ids;
  const bf16* embeds_in;
  bf16* h;
  int hidden;
  const bf16* final_norm;
  float eps;
  bf16* hidden_out;
  float* logits;
  int vocab;
  int* next_ids;
  int ban_id;
  int* pos_rw;
  int nstages;
  unsigned long long* prof;  // optional [nphases][2] globaltimer stamps of CTA 0 (EMU_MEGA_PROF=1)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  return __uint_as_float((uint32_t)__ldcg(reinterpret_cast<const unsigned short*>(p)) << 16);
}

// all consumer threads of all CTAs: nothing after the call is executed before every CTA has finished what precedes it.
// bar.sync orders the CTA's writes before thread 0's release-add; the acquire-load + bar.sync order everybody's
// later reads after it (PTX memory model cumulativity) — no separate membar needed on the critical path.
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned target) {
  consumer_bar();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    while (ld_acquire_u32(bar) < target) {
    }
  }
  consumer_bar();
}

__device__ __forceinline__ float cta_sum(float v, float* red) {  // 256 consumer threads; red >= 8 floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  consumer_bar();
  if (lane == 0) red[warp] = v;
  consumer_bar();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kTW; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ float cta_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  consumer_bar();
  if (lane == 0) red[warp] = v;
  consumer_bar();
  float t = red[0];
#pragma unroll
  for (int w = 1; w < kTW; ++w) t = fmaxf(t, red[w]);
  return t;
}

// one (sequence, head, kv-split) item of single-query attention over the bf16 cache; 256 consumer threads
template <int D>
__device__ void mega_attn_item(const MegaCtl& c, const MegaPhase* P, int item, float* scratch, int* s_flag) {
  constexpr int EPL = D / 8, VPL = EPL / 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsplit = c.nsplit, H = c.H;
  const int sp = item % nsplit, bh = item / nsplit;
  const int h = bh % H, b = bh / H;
  float* qs = scratch;         // [D]
  float* red = qs + D;         // [16]
  float* opart = red + 16;     // [32][D]
  float* sc = opart + 32 * D;  // [per]
  const int t_end = __ldcg(&c.pos[b]) + 1;
  const int t_begin = c.start ? __ldcg(&c.start[b]) : 0;
  const int n = t_end - t_begin;
  const int per = (n + nsplit - 1) / nsplit;
  const int t0 = t_begin + sp * per;
  const int t1 = min(t0 + per, t_end);
  consumer_bar();  // scratch may still be in use by the previous item
  for (int i = tid; i < D; i += 256) qs[i] = ldcg_bf16(P->q + ((long)b * H + h) * D + i) * c.scale;
  consumer_bar();
  const bf16* kb = P->kc + ((long)b * H + h) * c.t_max * D;
  const bf16* vb = P->vc + ((long)b * H + h) * c.t_max * D;
  // ---- scores ----
  const int part = lane & 7, tig = lane >> 3;  // 8 lanes per token, 4 tokens per warp pass, 32 per CTA pass
  float qreg[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) qreg[i] = qs[part * EPL + i];
  float lmax = -INFINITY;
  for (int tb = t0 + warp * 4; tb < t1; tb += 32) {
    const int t = tb + tig;
    float s = 0.f;
    const bool ok = t < t1;
    if (ok) {
      const uint4* kr = reinterpret_cast<const uint4*>(kb + (long)t * D + part * EPL);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const uint4 kv = __ldcg(kr + v);
        const uint32_t k4[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) s += bf16_lo(k4[j]) * qreg[v * 8 + 2 * j] + bf16_hi(k4[j]) * qreg[v * 8 + 2 * j + 1];
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (ok && part == 0) sc[t - t0] = s;
    if (ok) lmax = fmaxf(lmax, s);
  }
  const float m = cta_max(lmax, red);
  consumer_bar();
  float lsum = 0.f;
  for (int i = tid; i < t1 - t0; i += 256) {
    const float p = __expf(sc[i] - m);
    sc[i] = p;
    lsum += p;
  }
  const float l = cta_sum(lsum, red + 8);
  // ---- P V ----
  const int dpart = tid & 7, tl = tid >> 3;  // 32 token lanes x 8 d-slices
  float acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;
  for (int t = t0 + tl; t < t1; t += 32) {
    const float p = sc[t - t0];
    const uint4* vr = reinterpret_cast<const uint4*>(vb + (long)t * D + dpart * EPL);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const uint4 vv = __ldcg(vr + v);
      const uint32_t v4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[v * 8 + 2 * j] += p * bf16_lo(v4[j]);
        acc[v * 8 + 2 * j + 1] += p * bf16_hi(v4[j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < EPL; ++i) opart[tl * D + dpart * EPL + i] = acc[i];
  consumer_bar();
  float o = 0.f;
  if (tid < D) {
#pragma unroll 8
    for (int j = 0; j < 32; ++j) o += opart[j * D + tid];
  }
  if (nsplit == 1) {
    if (tid < D) P->out[((long)b * H + h) * D + tid] = __float2bfloat16_rn(l > 0.f ? o / l : 0.f);
    return;
  }
  float* ws_o = c.attn_ws;
  float* ws_ml = c.attn_ws + (size_t)c.B * H * nsplit * D;
  if (tid < D) ws_o[((long)bh * nsplit + sp) * D + tid] = o;
  if (tid == 0) {
    ws_ml[((long)bh * nsplit + sp) * 2] = m;
    ws_ml[((long)bh * nsplit + sp) * 2 + 1] = l;
  }
  __threadfence();
  consumer_bar();
  if (tid == 0) {
    const int prev = atomicAdd(&c.attn_counters[bh], 1);
    *s_flag = (prev == nsplit - 1);
    if (prev == nsplit - 1) c.attn_counters[bh] = 0;
  }
  consumer_bar();
  if (!*s_flag) return;
  __threadfence();
  if (tid < D) {
    float M = -INFINITY;
    for (int i = 0; i < nsplit; ++i) M = fmaxf(M, __ldcg(&ws_ml[((long)bh * nsplit + i) * 2]));
    float L = 0.f, O = 0.f;
    for (int i = 0; i < nsplit; ++i) {
      const float mi = __ldcg(&ws_ml[((long)bh * nsplit + i) * 2]), li = __ldcg(&ws_ml[((long)bh * nsplit + i) * 2 + 1]);
      const float w = (li > 0.f) ? __expf(mi - M) : 0.f;
      L += li * w;
      O += __ldcg(&ws_o[((long)bh * nsplit + i) * D + tid]) * w;
    }
    P->out[(long)bh * D + tid] = __float2bfloat16_rn(L > 0.f ? O / L : 0.f);
  }
}

__global__ void __launch_bounds__(kTThreads, 1) decode_mega_kernel(const MegaCtl c) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + c.nstages * kTStageBytes);
  uint64_t* empty_bar = full_bar + kTStages;
  float* red = reinterpret_cast<float*>(empty_bar + kTStages);  // [8 warps][kTRT*128]
  float* fin = red + kTW * kTRT * 128;
  bf16* xs = reinterpret_cast<bf16*>(fin + kTRT * 128);         // x staging / attention scratch
  __shared__ float s_ss[kTW][8];
  __shared__ float s_rstd[8];
  __shared__ int s_last;
  __shared__ GemvTmaParams s_params;
  __shared__ float s_bv[kTW];
  __shared__ int s_bi[kTW];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long G = gridDim.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kTW);
    }
    mbar_fence_init();
  }
  __syncthreads();
  int stage = 0;
  uint32_t phase = 0;

  if (warp == kTW) {
    // ===== producer: stream every projection of the step, in order, never waiting for phase boundaries =====
    if (lane == 0) {
      for (int ph = 0; ph < c.nphases; ++ph) {
        const MegaPhase* P = &c.phases[ph];
        if (P->type != MP_GEMV) continue;
        const long total = P->g.total, ge = P->g.geff;
        if ((long)blockIdx.x >= ge) continue;  // this matrix is shared by the first `geff` CTAs only
        const long c0 = (long)blockIdx.x * total / ge, c1 = ((long)blockIdx.x + 1) * total / ge;
        tma_produce(&c.maps[P->tmap], P->g.cpt, c0, c1, ring, full_bar, empty_bar, c.nstages, stage, phase);
      }
    }
    return;
  }

  // ===== consumers =====
  const int tid = threadIdx.x;
  unsigned epoch = 0;
  for (int ph = 0; ph < c.nphases; ++ph) {
    const MegaPhase* P = &c.phases[ph];
    if (ph > 0) grid_sync(c.bar, (unsigned)(++epoch * G));
    if (c.prof && blockIdx.x == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      c.prof[2 * ph] = t;
    }
    switch (P->type) {
      case MP_EMBED: {
        for (int b = blockIdx.x; b < c.B; b += (int)G) {
          const bf16* src = c.token_ids ? c.embed_table + (long)__ldcg(&c.token_ids[b]) * c.hidden
                                        : c.embeds_in + (long)b * c.hidden;
          const uint4* s4 = reinterpret_cast<const uint4*>(src);
          uint4* d4 = reinterpret_cast<uint4*>(c.h + (long)b * c.hidden);
          for (int i = tid; i < (c.hidden >> 3); i += 256) d4[i] = __ldcg(s4 + i);
        }
      } break;
      case MP_GEMV: {
        if (tid == 0) s_params = P->g;
        consumer_bar();
        const long total = s_params.total, ge = s_params.geff;
        if ((long)blockIdx.x < ge) {
          const long c0 = (long)blockIdx.x * total / ge, c1 = ((long)blockIdx.x + 1) * total / ge;
          tma_stage_x(s_params, xs, s_ss, s_rstd);
          tma_consume(&s_params, c0, c1, ring, full_bar, empty_bar, red, fin, xs, &s_last, stage, phase);
        }
      } break;
      case MP_ATTN: {
        const int items = c.B * c.H * c.nsplit;
        for (int item = blockIdx.x; item < items; item += (int)G) {
          if (c.D == 128) mega_attn_item<128>(c, P, item, reinterpret_cast<float*>(xs), &s_last);
          else mega_attn_item<64>(c, P, item, reinterpret_cast<float*>(xs), &s_last);
        }
      } break;
      case MP_NORM_OUT: {
        for (int b = blockIdx.x; b < c.B; b += (int)G) {
          float s = 0.f;
          for (int i = tid; i < c.hidden; i += 256) {
            const float v = ldcg_bf16(c.h + (long)b * c.hidden + i);
            s += v * v;
          }
          const float tot = cta_sum(s, red);
          const float rstd = rsqrtf(tot / (float)c.hidden + c.eps);
          for (int i = tid; i < c.hidden; i += 256) {
            const float v = ldcg_bf16(c.h + (long)b * c.hidden + i);
            c.hidden_out[(long)b * c.hidden + i] =
                __float2bfloat16_rn(round_bf16(v * rstd) * __bfloat162float(c.final_norm[i]));
          }
          consumer_bar();
        }
      } break;
      case MP_FINAL: {
        if (c.next_ids != nullptr) {
          for (int b = blockIdx.x; b < c.B; b += (int)G) {  // first-max argmax with the banned id excluded
            const float* row = c.logits + (long)b * c.vocab;
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int i = tid; i < c.vocab; i += 256) {
              const float v = (i == c.ban_id) ? -INFINITY : __ldcg(row + i);
              if (v > best || (v == best && i < bi)) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, best, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            consumer_bar();
            if (lane == 0) { s_bv[warp] = best; s_bi[warp] = bi; }
            consumer_bar();
            if (tid == 0) {
              for (int w = 1; w < kTW; ++w)
                if (s_bv[w] > best || (s_bv[w] == best && s_bi[w] < bi)) { best = s_bv[w]; bi = s_bi[w]; }
              c.next_ids[b] = bi;
            }
          }
        }
        if (blockIdx.x == 0 && tid < 8) c.pos_rw[tid] += 1;  // slot of the next token
      } break;
    }
    if (c.prof && blockIdx.x == 0 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      c.prof[2 * ph + 1] = t;
    }
  }
}

// ================================================================================================
// host side: plan (tensor maps + phase list) cached per argument set, cooperative launch
// ================================================================================================
struct MegaPlan {
  CUtensorMap* d_maps = nullptr;
  MegaPhase* d_phases = nullptr;
  int nphases = 0;
  size_t smem = 0;
  int nstages = 0;
  int nsplit = 1;
};

struct MegaState {
  std::map<std::tuple<int, const void*, const void*, const void*, const void*, const void*, int>, MegaPlan> plans;
  unsigned* d_bar = nullptr;
  float* attn_ws = nullptr;
  int* attn_counters = nullptr;
  float* gemv_ws = nullptr;
  int* gemv_counters = nullptr;
  bool attr_set = false;
  int disabled = -1;
  unsigned long long* d_prof = nullptr;
};

static MegaState* mega_state(EmuEngine* e) {
  if (!e->mega) e->mega = new MegaState();
  return (MegaState*)e->mega;
}
void mega_destroy(void* p) { delete (MegaState*)p; }

// debugging aid: copy CTA 0's per-phase timestamps of the last step to the host (ns); returns the phase count
extern "C" int emu_debug_mega_profile(EmuEngine* e, unsigned long long* out, int* types, int max_phases) {
  if (!e || !e->mega) return 0;
  MegaState* ms = (MegaState*)e->mega;
  if (!ms->d_prof || ms->plans.empty()) return 0;
  const MegaPlan& pl = ms->plans.begin()->second;
  const int n = pl.nphases < max_phases ? pl.nphases : max_phases;
  cudaDeviceSynchronize();
  cudaMemcpy(out, ms->d_prof, (size_t)n * 16, cudaMemcpyDeviceToHost);
  std::vector<MegaPhase> ph(n);
  cudaMemcpy(ph.data(), pl.d_phases, (size_t)n * sizeof(MegaPhase), cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; ++i) types[i] = ph[i].type * 100 + (ph[i].type == 0 ? ph[i].g.a.mode : 0);
  return n;
}

static bf16* kv_slab(EmuEngine* e, int layer, int kv) {
  const EmuConfig& c = e->cfg;
  const size_t per = (size_t)c.llm_max_batch * e->Hl * c.llm_max_seq * c.llm_head_dim;
  return e->kv + ((size_t)layer * 2 + kv) * per;
}

static void fill_gemv(GemvTmaParams& g, int N, int K, const bf16* x, int ldx, int B, const bf16* norm_w, float eps,
                      int mode, const bf16* residual, int ldr, void* y, int ldy, int out_fp32, int nstages, float* ws,
                      int* counters) {
  memset(&g, 0, sizeof(g));
  g.a.N = N; g.a.K = K; g.a.x = x; g.a.ldx = ldx; g.a.B = B; g.a.norm_w = norm_w; g.a.norm_eps = eps; g.a.mode = mode;
  g.a.residual = residual; g.a.ldr = ldr; g.a.y = y; g.a.ldy = ldy; g.a.out_fp32 = out_fp32;
  g.kpad = (K + kTCols - 1) / kTCols * kTCols;
  g.ldxs = g.kpad + 8;
  g.cpt = g.kpad / kTCols;
  g.nstages = nstages;
  const int groups = (N + kTRows - 1) / kTRows;
  g.total = (long)groups * g.cpt;
  long geff = kNumSMs;
  if (geff > g.total) geff = g.total;
  if (geff > (long)groups * (kTMaxParts - 3)) geff = (long)groups * (kTMaxParts - 3);
  g.geff = (int)geff;
  g.ws = ws;
  g.counters = counters;
}

// returns EMU_ERR_UNSUPPORTED when the step does not fit the persistent kernel (caller uses the multi-kernel graph)
int decode_mega_step(EmuEngine* e, const int32_t* token_ids, const void* embeds, int B, float* logits, void* hidden,
                     int32_t* next_ids, int ban_id, cudaStream_t st) {
  MegaState* ms = mega_state(e);
  {
    // Opt-in (EMU_MEGA=1): measured on B200 the grid barriers (~4.8 us x 302 per step) cost as much as the launch
    // boundaries they replace (12.2 ms vs 11.6 ms per token for the CUDA-graphed multi-kernel step), see DESIGN.md.
    // EMU_NO_MEGA=1 always wins.  Read every call: tests flip both to cover both paths.
    const char* on = getenv("EMU_MEGA");
    const char* off = getenv("EMU_NO_MEGA");
    ms->disabled = (on && on[0] == '1' && !(off && off[0] == '1')) ? 0 : 1;
  }
  const EmuConfig& c = e->cfg;
  const int Hd = c.llm_hidden, D = c.llm_head_dim, Hl = e->Hl, Fl = e->Fl;
  if (ms->disabled || e->tp_size != 1 || (Hd % 8) || (Fl % 8)) return EMU_ERR_UNSUPPORTED;
  const int kmax = Hd > Fl ? Hd : Fl;
  int nsplit = kNumSMs / (B * Hl);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 16) nsplit = 16;
  const size_t per = (size_t)(c.llm_max_seq + nsplit - 1) / nsplit + 8;
  // one shared-memory region serves as x staging (GEMV phases) and as attention scratch (D + 16 + 32*D + per floats)
  size_t xs_bytes = (size_t)B * (((size_t)kmax + 255) / 256 * 256 + 8) * 2;
  const size_t attn_bytes = ((size_t)D + 16 + 32 * (size_t)D + per) * 4;
  if (attn_bytes > xs_bytes) xs_bytes = attn_bytes;
  xs_bytes = (xs_bytes + 15) & ~size_t(15);
  int nstages = 8;
  size_t smem;
  for (;;) {
    smem = 1024 + (size_t)nstages * kTStageBytes + 2 * kTStages * 8 + (size_t)(kTW + 1) * kTRT * 128 * 4 + xs_bytes + 64;
    if (smem <= 200 * 1024 || nstages <= 4) break;
    --nstages;
  }
  if (smem > 200 * 1024) return EMU_ERR_UNSUPPORTED;
  {  // a row group may be finished by at most kTMaxParts CTAs
    const int shapes[5][2] = {{3 * Hl * D, Hd}, {Hd, Hl * D}, {2 * Fl, Hd}, {Hd, Fl}, {e->Vl, Hd}};
    for (auto& s : shapes) {
      const long cpt = (s[1] + kTCols - 1) / kTCols, groups = (s[0] + kTRows - 1) / kTRows;
      long share = groups * cpt / kNumSMs;
      if (share < 1) share = 1;
      if (cpt > (kTMaxParts - 2) * share) return EMU_ERR_UNSUPPORTED;
    }
  }
  const int groups_max = ((3 * Hl * D > 2 * Fl ? 3 * Hl * D : 2 * Fl) + kTRows - 1) / kTRows;
  const int groups_head = (e->Vl + kTRows - 1) / kTRows;
  if (groups_max > kTWsGroups || groups_head > kTWsGroups) return EMU_ERR_UNSUPPORTED;

  if (!ms->d_bar) {
    if (cudaMalloc((void**)&ms->d_bar, 64) != cudaSuccess) return e->fail(EMU_ERR_NOMEM, "mega alloc");
    if (cudaMalloc((void**)&ms->attn_ws, (size_t)8 * Hl * 16 * (D + 2) * sizeof(float)) != cudaSuccess) return e->fail(EMU_ERR_NOMEM, "mega alloc");
    if (cudaMalloc((void**)&ms->attn_counters, (size_t)8 * Hl * sizeof(int)) != cudaSuccess) return e->fail(EMU_ERR_NOMEM, "mega alloc");
    if (cudaMalloc((void**)&ms->gemv_ws, (size_t)kTWsGroups * kTMaxParts * kTRT * 128 * sizeof(float)) != cudaSuccess) return e->fail(EMU_ERR_NOMEM, "mega alloc");
    if (cudaMalloc((void**)&ms->gemv_counters, (size_t)kTWsGroups * sizeof(int)) != cudaSuccess) return e->fail(EMU_ERR_NOMEM, "mega alloc");
    cudaMemset(ms->attn_counters, 0, (size_t)8 * Hl * sizeof(int));
    cudaMemset(ms->gemv_counters, 0, (size_t)kTWsGroups * sizeof(int));
  }
  if (!ms->attr_set) {
    if (cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return e->fail(EMU_ERR_CUDA, "mega smem attribute");
    ms->attr_set = true;
  }

  // the plan only bakes in the logits buffer and which outputs exist; token/embeds/next-id pointers travel in MegaCtl
  const float* lg_key = logits ? logits : e->dec_logits_local;
  auto key = std::make_tuple(B, (const void*)lg_key, (const void*)nullptr, (const void*)nullptr,
                             (const void*)(hidden ? e : nullptr), (const void*)((logits || next_ids) ? e : nullptr), 0);
  auto it = ms->plans.find(key);
  if (it == ms->plans.end()) {
    MegaPlan pl;
    pl.nstages = nstages;
    pl.smem = smem;
    pl.nsplit = nsplit;
    std::vector<CUtensorMap> maps;
    std::vector<MegaPhase> phases;
    auto add_map = [&](const bf16* W, long N, long K) -> int {
      CUtensorMap tm;
      if (make_tmap_2d(&tm, W, N, K, K, kTRows) != EMU_OK) return -1;
      maps.push_back(tm);
      return (int)maps.size() - 1;
    };
    MegaPhase ph;
    memset(&ph, 0, sizeof(ph));
    ph.type = MP_EMBED;
    phases.push_back(ph);
    for (int l = 0; l < c.llm_layers; ++l) {
      const LlmLayer& L = e->layers[l];
      bf16* kc = kv_slab(e, l, 0);
      bf16* vc = kv_slab(e, l, 1);
      // 1. RMSNorm + QKV + RoPE + KV append
      memset(&ph, 0, sizeof(ph));
      ph.type = MP_GEMV;
      ph.tmap = add_map(L.wqkv, 3L * Hl * D, Hd);
      fill_gemv(ph.g, 3 * Hl * D, Hd, e->dec_h, Hd, B, L.ln1, c.llm_rms_eps, GEMV_ROPE_QKV, nullptr, 0, e->dec_q, Hl * D,
                0, nstages, ms->gemv_ws, ms->gemv_counters);
      ph.g.a.n_heads = Hl; ph.g.a.head_dim = D; ph.g.a.rope_cos = e->rope_cos; ph.g.a.rope_sin = e->rope_sin;
      ph.g.a.pos = e->d_pos; ph.g.a.pos_off = e->d_posoff; ph.g.a.k_cache = kc; ph.g.a.v_cache = vc;
      ph.g.a.t_max = c.llm_max_seq;
      phases.push_back(ph);
      // 2. attention
      memset(&ph, 0, sizeof(ph));
      ph.type = MP_ATTN;
      ph.q = e->dec_q; ph.kc = kc; ph.vc = vc; ph.out = e->dec_attn;
      phases.push_back(ph);
      // 3. o_proj + residual
      memset(&ph, 0, sizeof(ph));
      ph.type = MP_GEMV;
      ph.tmap = add_map(L.wo, Hd, (long)Hl * D);
      fill_gemv(ph.g, Hd, Hl * D, e->dec_attn, Hl * D, B, nullptr, 0.f, EPI_NONE, e->dec_h, Hd, e->dec_h, Hd, 0, nstages,
                ms->gemv_ws, ms->gemv_counters);
      phases.push_back(ph);
      // 4. RMSNorm + gate/up + SwiGLU
      memset(&ph, 0, sizeof(ph));
      ph.type = MP_GEMV;
      ph.tmap = add_map(L.wgu, 2L * Fl, Hd);
      fill_gemv(ph.g, 2 * Fl, Hd, e->dec_h, Hd, B, L.ln2, c.llm_rms_eps, EPI_SWIGLU, nullptr, 0, e->dec_act, Fl, 0, nstages,
                ms->gemv_ws, ms->gemv_counters);
      phases.push_back(ph);
      // 5. down + residual
      memset(&ph, 0, sizeof(ph));
      ph.type = MP_GEMV;
      ph.tmap = add_map(L.wdown, Hd, Fl);
      fill_gemv(ph.g, Hd, Fl, e->dec_act, Fl, B, nullptr, 0.f, EPI_NONE, e->dec_h, Hd, e->dec_h, Hd, 0, nstages,
                ms->gemv_ws, ms->gemv_counters);
      phases.push_back(ph);
    }
    if (hidden) {
      memset(&ph, 0, sizeof(ph));
      ph.type = MP_NORM_OUT;
      phases.push_back(ph);
    }
    float* lg = logits ? logits : e->dec_logits_local;
    if (logits || next_ids) {
      memset(&ph, 0, sizeof(ph));
      ph.type = MP_GEMV;
      ph.tmap = add_map(e->lm_head, e->Vl, Hd);
      fill_gemv(ph.g, e->Vl, Hd, e->dec_h, Hd, B, e->final_norm, c.llm_rms_eps, EPI_NONE, nullptr, 0, lg, c.llm_vocab, 1,
                nstages, ms->gemv_ws, ms->gemv_counters);
      phases.push_back(ph);
    }
    memset(&ph, 0, sizeof(ph));
    ph.type = MP_FINAL;
    phases.push_back(ph);
    for (auto& p : phases)
      if (p.type == MP_GEMV && p.tmap < 0) return e->fail(EMU_ERR_CUDA, "mega tensor map encode failed");
    pl.nphases = (int)phases.size();
    if (cudaMalloc((void**)&pl.d_maps, maps.size() * sizeof(CUtensorMap)) != cudaSuccess ||
        cudaMalloc((void**)&pl.d_phases, phases.size() * sizeof(MegaPhase)) != cudaSuccess)
      return e->fail(EMU_ERR_NOMEM, "mega plan alloc");
    e->owned.push_back(pl.d_maps);
    e->owned.push_back(pl.d_phases);
    if (cudaMemcpy(pl.d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(pl.d_phases, phases.data(), phases.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice) != cudaSuccess)
      return e->fail(EMU_ERR_CUDA, "mega plan upload");
    ms->plans[key] = pl;
    it = ms->plans.find(key);
  }
  const MegaPlan& pl = it->second;
  MegaCtl ctl;
  memset(&ctl, 0, sizeof(ctl));
  ctl.maps = pl.d_maps; ctl.phases = pl.d_phases; ctl.nphases = pl.nphases; ctl.bar = ms->d_bar;
  ctl.B = B; ctl.H = Hl; ctl.D = D; ctl.t_max = c.llm_max_seq; ctl.nsplit = pl.nsplit;
  ctl.pos = e->d_pos; ctl.start = e->d_start; ctl.scale = 1.0f / sqrtf((float)D);
  ctl.attn_ws = ms->attn_ws; ctl.attn_counters = ms->attn_counters;
  ctl.embed_table = e->embed; ctl.token_ids = token_ids; ctl.embeds_in = (const bf16*)embeds; ctl.h = e->dec_h;
  ctl.hidden = Hd; ctl.final_norm = e->final_norm; ctl.eps = c.llm_rms_eps; ctl.hidden_out = (bf16*)hidden;
  ctl.logits = logits ? logits : e->dec_logits_local; ctl.vocab = c.llm_vocab; ctl.next_ids = next_ids;
  ctl.ban_id = ban_id; ctl.pos_rw = e->d_pos; ctl.nstages = pl.nstages;
  if (getenv("EMU_MEGA_PROF")) {
    if (!ms->d_prof && cudaMalloc((void**)&ms->d_prof, 4096 * 16) != cudaSuccess) return e->fail(EMU_ERR_NOMEM, "prof alloc");
    ctl.prof = pl.nphases <= 4096 ? ms->d_prof : nullptr;
  }
  if (cudaMemsetAsync(ms->d_bar, 0, 4, st) != cudaSuccess) return e->fail(EMU_ERR_CUDA, "mega barrier reset");
  void* args[] = {(void*)&ctl};
  cudaError_t ce = cudaLaunchCooperativeKernel((void*)decode_mega_kernel, dim3(kNumSMs), dim3(kTThreads), args, pl.smem, st);
  if (ce != cudaSuccess) return e->fail(EMU_ERR_CUDA, std::string("mega launch failed: ") + cudaGetErrorString(ce));
  count_launch(1);
  return EMU_OK;
}

}  // namespace emu
