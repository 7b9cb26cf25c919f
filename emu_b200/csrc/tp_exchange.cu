// emu_b200 — tensor-parallel exchange over NVLink peer memory for the decode loop.
//
// A decode step at tensor-parallel degree N has 2 row-parallel GEMVs per decoder layer (o_proj, down_proj) whose
// partial outputs ([batch <= 8, hidden] — 13..213 KB) must be summed over ranks and added to the residual stream, plus
// one gather of the vocab-sharded logits.  At that size a collective is pure latency: NCCL costs ~15-20 us per call
// x 121 calls per token, more than the weight streaming itself at N >= 4.  Here every rank owns one cudaMalloc'd
// exchange buffer that all peers map through CUDA IPC; one small kernel per exchange
//   1. PUSHES its fp32 partial straight into every peer's receive slot with 16-byte stores over NVLink,
//   2. publishes a monotonically increasing sequence flag per (source rank, CTA) with a system-scope release,
//   3. spins (acquire, system scope) on the flags the peers wrote into ITS OWN memory — local polling only,
//   4. reduces the N slots in fixed rank order (bitwise identical result on every rank) + residual -> bf16.
// CTAs are independent (each owns a column slice and its own flags/sequence counter), so there is no grid barrier, the
// kernel is CUDA-graph capturable (all state lives on the device) and sits in the PDL chain between the GEMVs.
// Two receive slots (sequence parity) suffice: a rank can only be one exchange ahead of its slowest peer.
//
// Replaces (reference side): the reference has no tensor parallelism — SURVEY.md §8(e) defines this split; the math
// is o_proj / down_proj of transformers' LlamaDecoderLayer (Emu2/emu/lm.py:37-41 instantiates it).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "engine.h"

namespace emu {

constexpr int kTpMaxCtas = 32;
constexpr int kTpThreads = 256;

struct TpLayout {
  size_t red_slot;     // floats per reduce slot  (Bmax * hidden)
  size_t gat_slot;     // floats per gather slot  (Bmax * Vl)
  size_t red_off, gat_off, flag_red_off, flag_gat_off, seq_off, ll_off, step_off, total;  // in floats / 4-byte words
};
static TpLayout tp_layout(int n, int bmax, int hidden, int vl) {
  TpLayout L;
  L.red_slot = ((size_t)bmax * hidden + 3) / 4 * 4;
  L.gat_slot = ((size_t)bmax * vl + 3) / 4 * 4;
  L.red_off = 0;
  L.gat_off = L.red_off + 2 * (size_t)n * L.red_slot;
  L.flag_red_off = L.gat_off + 2 * (size_t)n * L.gat_slot;
  L.flag_gat_off = L.flag_red_off + (size_t)8 * kTpMaxCtas;
  L.seq_off = L.flag_gat_off + (size_t)8 * kTpMaxCtas;
  L.ll_off = (L.seq_off + 2 * kTpMaxCtas + 3) / 4 * 4;  // {value, flag} words: 2 floats per element
  L.step_off = L.ll_off + 2 * 2 * (size_t)n * L.red_slot;
  L.total = L.step_off + 4;
  return L;
}

struct TpPeers {
  float* base[8];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// common push / publish / wait part.  `src` = this rank's contribution (n_elem floats, 4-float aligned count),
// slot(parity, r) = base + (parity * n + r) * slot_elems.
__device__ __forceinline__ unsigned tp_push_and_wait(const TpPeers& peers, int rank, int n, const float* __restrict__ src,
                                                     long n_elem, size_t data_off, size_t slot_elems, size_t flag_off,
                                                     size_t seq_off, int seq_idx) {
  __shared__ unsigned s_seq;
  float* local = peers.base[rank];
  unsigned* seq_ptr = reinterpret_cast<unsigned*>(local + seq_off) + seq_idx * kTpMaxCtas + blockIdx.x;
  if (threadIdx.x == 0) s_seq = *seq_ptr;
  __syncthreads();
  const unsigned seq = s_seq;
  const unsigned parity = seq & 1u;
  const long n4 = n_elem >> 2;
  const size_t my_slot = data_off + ((size_t)parity * n + rank) * slot_elems;
  for (long i = (long)blockIdx.x * kTpThreads + threadIdx.x; i < n4; i += (long)gridDim.x * kTpThreads) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src) + i);
    for (int r = 0; r < n; ++r) {
      if (r == rank) continue;
      reinterpret_cast<float4*>(peers.base[r] + my_slot)[i] = v;  // NVLink peer store
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < n && threadIdx.x != rank) {
    unsigned* f = reinterpret_cast<unsigned*>(peers.base[threadIdx.x] + flag_off) + rank * kTpMaxCtas + blockIdx.x;
    st_release_sys(f, seq + 1);
    const unsigned* mine = reinterpret_cast<const unsigned*>(local + flag_off) + threadIdx.x * kTpMaxCtas + blockIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(mine) - (seq + 1)) < 0) {
      if (globaltimer_ns() - t0 > 20000000000ull) {  // 20 s: a peer died — fail loudly instead of hanging the GPU
        printf("emu_b200: tensor-parallel exchange timed out (rank %d waiting for rank %d)\n", rank, (int)threadIdx.x);
        __trap();
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *seq_ptr = seq + 1;
  return parity;
}

// h[b, :] += sum_r partial_r[b, :]   (fixed rank order; own partial read from `part`)
__global__ void __launch_bounds__(kTpThreads) tp_reduce_add_kernel(TpPeers peers, int rank, int n,
                                                                   const float* __restrict__ part, bf16* h, long n_elem,
                                                                   size_t data_off, size_t slot_elems, size_t flag_off,
                                                                   size_t seq_off, int pdl) {
  if (pdl) {
    pdl_launch_dependents();
    pdl_wait();
  }
  const unsigned parity = tp_push_and_wait(peers, rank, n, part, n_elem, data_off, slot_elems, flag_off, seq_off, 0);
  const float* local = peers.base[rank];
  const long n4 = n_elem >> 2;
  for (long i = (long)blockIdx.x * kTpThreads + threadIdx.x; i < n4; i += (long)gridDim.x * kTpThreads) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < n; ++r) {
      const float4 v = (r == rank) ? __ldcg(reinterpret_cast<const float4*>(part) + i)
                                   : __ldcg(reinterpret_cast<const float4*>(local + data_off +
                                                                             ((size_t)parity * n + r) * slot_elems) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    uint2 hv = reinterpret_cast<uint2*>(h)[i];
    // the row-parallel projection output is rounded to bf16 before the residual add, as in the unsharded model
    hv.x = pack_bf16(bf16_lo(hv.x) + round_bf16(acc.x), bf16_hi(hv.x) + round_bf16(acc.y));
    hv.y = pack_bf16(bf16_lo(hv.y) + round_bf16(acc.z), bf16_hi(hv.y) + round_bf16(acc.w));
    reinterpret_cast<uint2*>(h)[i] = hv;
  }
}

// logits[b, r*Vl + v] = shard_r[b, v]   (Vl = ceil(V / n))
__global__ void __launch_bounds__(kTpThreads) tp_gather_logits_kernel(TpPeers peers, int rank, int n,
                                                                      const float* __restrict__ shard, float* logits,
                                                                      int B, int Vl, int V, long n_pad, size_t data_off,
                                                                      size_t slot_elems, size_t flag_off, size_t seq_off,
                                                                      int pdl) {
  if (pdl) {
    pdl_launch_dependents();
    pdl_wait();
  }
  const unsigned parity = tp_push_and_wait(peers, rank, n, shard, n_pad, data_off, slot_elems, flag_off, seq_off, 1);
  const float* local = peers.base[rank];
  const long per = (long)B * Vl;
  for (long i = (long)blockIdx.x * kTpThreads + threadIdx.x; i < per * n; i += (long)gridDim.x * kTpThreads) {
    const int r = (int)(i / per);
    const long j = i - (long)r * per;
    const int b = (int)(j / Vl), v = (int)(j - (long)b * Vl);
    const float x = (r == rank) ? __ldcg(shard + j) : __ldcg(local + data_off + ((size_t)parity * n + r) * slot_elems + j);
    const long col = (long)r * Vl + v;
    if (col < V) logits[(long)b * V + col] = x;  // rows past V are the zero padding of the last shard
  }
}

// consumer of the fused push (GemvArgs::ll_*): poll the {value, flag} words of all ranks, reduce in rank order, add to h
__global__ void __launch_bounds__(kTpThreads) tp_ll_reduce_kernel(const uint4* __restrict__ ll, const unsigned* step, int idx,
                                                                  int rank, int n, size_t slot_elems, bf16* h, long n_elem,
                                                                  int pdl) {
  if (pdl) {
    pdl_launch_dependents();
    pdl_wait();  // the step counter and h are only stable once every earlier kernel of the stream has retired
  }
  const unsigned flag = __ldcg(step) * 256u + (unsigned)idx + 1u;
  const uint4* base = ll + ((size_t)(idx & 1) * n * slot_elems) / 2;
  const long n2 = n_elem >> 1;
  for (long i = (long)blockIdx.x * kTpThreads + threadIdx.x; i < n2; i += (long)gridDim.x * kTpThreads) {
    float a0 = 0.f, a1 = 0.f;
    for (int r = 0; r < n; ++r) {
      const uint4* p = base + ((size_t)r * slot_elems) / 2 + i;
      uint4 v;
      unsigned long long t0 = 0;
      for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        if (v.y == flag && v.w == flag) break;
        if (t0 == 0) t0 = globaltimer_ns();
        else if (globaltimer_ns() - t0 > 20000000000ull) {
          printf("emu_b200: tensor-parallel exchange timed out (rank %d waiting for rank %d, exchange %d)\n", rank, r, idx);
          __trap();
        }
      }
      a0 += __uint_as_float(v.x);
      a1 += __uint_as_float(v.z);
    }
    uint32_t hv = reinterpret_cast<uint32_t*>(h)[i];
    hv = pack_bf16(bf16_lo(hv) + round_bf16(a0), bf16_hi(hv) + round_bf16(a1));
    reinterpret_cast<uint32_t*>(h)[i] = hv;
  }
}

static int tp_grid(long n_elem) {
  long g = (n_elem / 4 + kTpThreads - 1) / kTpThreads;
  if (g < 1) g = 1;
  return (int)(g > kTpMaxCtas ? kTpMaxCtas : g);
}

static int launch_pdl(const void* fn, dim3 grid, void** args, int pdl, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kTpThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelExC(&cfg, fn, args) == cudaSuccess ? EMU_OK : EMU_ERR_CUDA;
}

int tp_reduce_add(EmuEngine* e, const float* part, bf16* h, long n_elem, int pdl, cudaStream_t st) {
  if (!e->tp_p2p || (n_elem & 3)) return EMU_ERR_STATE;
  const TpLayout L = tp_layout(e->tp_size, e->cfg.llm_max_batch, e->cfg.llm_hidden, e->Vl);
  if ((size_t)n_elem > L.red_slot) return EMU_ERR_INVALID;
  TpPeers peers;
  for (int r = 0; r < 8; ++r) peers.base[r] = e->tp_peer[r];
  int rank = e->tp_rank, n = e->tp_size;
  size_t data_off = L.red_off, slot = L.red_slot, flag_off = L.flag_red_off, seq_off = L.seq_off;
  void* args[] = {&peers, &rank, &n, &part, &h, &n_elem, &data_off, &slot, &flag_off, &seq_off, &pdl};
  return launch_pdl((const void*)tp_reduce_add_kernel, dim3(tp_grid(n_elem)), args, pdl, st);
}

// fill the fused-push fields of a row-parallel GEMV (exchange index idx = 2*layer + {0: o_proj, 1: down_proj})
int tp_ll_prepare(EmuEngine* e, GemvArgs& g, int idx) {
  if (!e->tp_p2p || !e->tp_ll) return EMU_ERR_UNSUPPORTED;
  const TpLayout L = tp_layout(e->tp_size, e->cfg.llm_max_batch, e->cfg.llm_hidden, e->Vl);
  for (int r = 0; r < e->tp_size; ++r) g.ll_peer[r] = e->tp_peer[r] + L.ll_off;
  g.ll_n = e->tp_size; g.ll_rank = e->tp_rank; g.ll_idx = idx;
  g.ll_slot_elems = (long)L.red_slot;
  g.ll_step = reinterpret_cast<const unsigned*>(e->tp_peer[e->tp_rank] + L.step_off);
  return EMU_OK;
}
unsigned* tp_step_counter(EmuEngine* e) {
  if (!e->tp_p2p) return nullptr;
  const TpLayout L = tp_layout(e->tp_size, e->cfg.llm_max_batch, e->cfg.llm_hidden, e->Vl);
  return reinterpret_cast<unsigned*>(e->tp_peer[e->tp_rank] + L.step_off);
}
int tp_ll_reduce(EmuEngine* e, bf16* h, long n_elem, int idx, int pdl, cudaStream_t st) {
  if (!e->tp_p2p || (n_elem & 1)) return EMU_ERR_STATE;
  const TpLayout L = tp_layout(e->tp_size, e->cfg.llm_max_batch, e->cfg.llm_hidden, e->Vl);
  const uint4* ll = reinterpret_cast<const uint4*>(e->tp_peer[e->tp_rank] + L.ll_off);
  const unsigned* step = reinterpret_cast<const unsigned*>(e->tp_peer[e->tp_rank] + L.step_off);
  int rank = e->tp_rank, n = e->tp_size;
  size_t slot = L.red_slot;
  void* args[] = {&ll, &step, &idx, &rank, &n, &slot, &h, &n_elem, &pdl};
  long g = (n_elem / 2 + kTpThreads - 1) / kTpThreads;
  if (g > kTpMaxCtas) g = kTpMaxCtas;
  return launch_pdl((const void*)tp_ll_reduce_kernel, dim3((unsigned)g), args, pdl, st);
}

int tp_gather_logits(EmuEngine* e, const float* shard, float* logits, int B, int pdl, cudaStream_t st) {
  if (!e->tp_p2p) return EMU_ERR_STATE;
  const TpLayout L = tp_layout(e->tp_size, e->cfg.llm_max_batch, e->cfg.llm_hidden, e->Vl);
  long n_pad = ((long)B * e->Vl + 3) / 4 * 4;  // the shard buffer is allocated with this padding
  TpPeers peers;
  for (int r = 0; r < 8; ++r) peers.base[r] = e->tp_peer[r];
  int rank = e->tp_rank, n = e->tp_size, Vl = e->Vl, V = e->cfg.llm_vocab;
  size_t data_off = L.gat_off, slot = L.gat_slot, flag_off = L.flag_gat_off, seq_off = L.seq_off;
  void* args[] = {&peers, &rank, &n, &shard, &logits, &B, &Vl, &V, &n_pad, &data_off, &slot, &flag_off, &seq_off, &pdl};
  return launch_pdl((const void*)tp_gather_logits_kernel, dim3(tp_grid(n_pad)), args, pdl, st);
}

// Allocate the exchange buffer, swap CUDA IPC handles with the peers (through the NCCL communicator that already
// exists) and map theirs.  Leaves e->tp_p2p false (NCCL path) if any rank cannot map any peer.
int tp_exchange_setup(EmuEngine* e, int (*allgather_bytes)(EmuEngine*, const void*, void*, size_t),
                      int (*allreduce_min_int)(EmuEngine*, int*)) {
  const char* env = getenv("EMU_TP_P2P");
  int want = !(env && atoi(env) == 0);
  const int n = e->tp_size;
  if (n > 8) want = 0;
  const TpLayout L = tp_layout(n, e->cfg.llm_max_batch, e->cfg.llm_hidden, e->Vl);
  int ok = want;
  float* buf = nullptr;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    // plain cudaMalloc (not a pool / VMM allocation): required for cudaIpcGetMemHandle
    if (cudaMalloc((void**)&buf, L.total * sizeof(float)) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(buf, 0, L.total * sizeof(float)) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, buf) != cudaSuccess) ok = 0;
    cudaGetLastError();
  }
  // every rank takes part in the collectives below whether or not its own setup worked
  char all[8 * sizeof(cudaIpcMemHandle_t)];
  if (allgather_bytes(e, &mine, all, sizeof(mine)) != EMU_OK) return EMU_ERR_NCCL;
  if (ok) {
    for (int r = 0; r < n && ok; ++r) {
      if (r == e->tp_rank) {
        e->tp_peer[r] = buf;
        continue;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, all + (size_t)r * sizeof(h), sizeof(h));
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      } else {
        e->tp_peer[r] = (float*)p;
      }
    }
  }
  int all_ok = ok;
  if (allreduce_min_int(e, &all_ok) != EMU_OK) return EMU_ERR_NCCL;
  if (!all_ok) {
    if (getenv("EMU_TP_DEBUG")) fprintf(stderr, "emu_b200: rank %d peer exchange unavailable, using NCCL\n", e->tp_rank);
    tp_exchange_teardown(e);
    if (buf) cudaFree(buf);
    return EMU_OK;
  }
  e->tp_comm = buf;
  e->tp_p2p = true;
  {
    const char* ll = getenv("EMU_TP_LL");
    e->tp_ll = !(ll && atoi(ll) == 0);
    const char* fold = getenv("EMU_TP_FOLD");
    e->tp_fold = e->tp_ll && !(fold && atoi(fold) == 0);
  }
  if (getenv("EMU_TP_DEBUG")) fprintf(stderr, "emu_b200: rank %d/%d NVLink peer exchange enabled\n", e->tp_rank, n);
  return EMU_OK;
}

void tp_exchange_teardown(EmuEngine* e) {
  for (int r = 0; r < 8; ++r) {
    if (e->tp_peer[r] && r != e->tp_rank) cudaIpcCloseMemHandle(e->tp_peer[r]);
    e->tp_peer[r] = nullptr;
  }
  if (e->tp_comm) cudaFree(e->tp_comm);
  e->tp_comm = nullptr;
  e->tp_p2p = false;
  cudaGetLastError();
}

}  // namespace emu
