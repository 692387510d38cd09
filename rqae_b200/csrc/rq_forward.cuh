// Fused RQAE forward / encode kernel for sm_100a (B200).
//
// Replaces the per-layer op sequence of rqae/model.py:199-230 (reference, harish-kamath/rqae):
// F.linear(D->4), L2-normalise, [T,4]x[4,K] cos-sim matmul, argmax, gather, straight-through
// add/sub, F.linear(4->D), residual subtract, reconstruction add -- ~13 launches per layer,
// 1024 layers -- by ONE persistent kernel in which the residual never leaves registers.
//
// CTA = 384 threads, one CTA per SM (the register file is the capacity that matters):
//   warps 0-7   compute   : 256 threads; thread t owns elements d = j*256 + t (j < E) of the D axis for ALL
//                           2*TG tokens of the CTA's current unit, residual r[2*TG][E] in registers.
//                           The tokens form two *phases* A and B of TG tokens each.  Every warp runs
//                               pass(A, l)  pass(B, l)  pass(A, l+1)  pass(B, l+1) ...
//                           so the layer-l argmax of phase A is computed (by a quantizer warp) while the
//                           same compute warps are busy with pass(B, l): the serial dependency of the
//                           residual recurrence is hidden by construction, not by timing luck.
//   warp  8, 9  quantizer : warp 8 serves phase A, warp 9 phase B: cross-warp sum of the in-projection
//                           partials, cos-sim argmax (lowest original index wins ties, NaN -> index 0 as
//                           torch.argmax), straight-through value, code output
//   warp  10    producer  : one lane streaming weight chunks L2 -> shared memory with bulk TMA
//                           (cp.async.bulk + mbarrier complete_tx) through an NSLOT-deep ring; a chunk is
//                           read by pass(A, l) and again by pass(B, l), then released
// Registers are re-split with setmaxnreg: the compute warpgroups grow, the helper warpgroup shrinks.
//
// One *pass* over stage s (see rq_layout.h) does, per owned element d and per token pair, with packed
// FFMA2 (two tokens per instruction, each half IEEE fp32 round-to-nearest):
//     o   = fma(w_out[d][3], c'3, fma(w_out[d][2], c'2, fma(w_out[d][1], c'1, fma(w_out[d][0], c'0, b_out[d]))))
//     r_d = r_d - o                                  (model.py:221-223)
//     acc[k] = fma(w_in[k][d], r_d, acc[k]), k<4     (model.py:211, partial over the thread's elements)
// then reduces acc over the warp with a shuffle butterfly and hands 8 per-warp partials per (token, k)
// to the quantizer warp through shared memory.  Summation order is fixed (thread-sequential over j,
// lane tree with strides 1,2,16,8,4, the 8 warps as a pairwise tree, then + b_in), so results do not depend on
// the tile a token lands in, on the grid size or on timing.  The reconstruction is emitted as
// q = x - r_final (one extra read of x) instead of a second register-resident accumulator.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "rq_common.cuh"
#include "rq_layout.h"

namespace rq {

struct FwdParams {
  const unsigned char* packed;  // packed buffer (rq_layout.h)
  unsigned long long off_bin, off_cbt, off_map, off_tp, off_map3, off_stage;
  const float* codebook;        // original codebook, device: [nq][K][4] or [1][K][4]
  int cb_shared;                // 1: one table for all layers (fsq / round_fsq)
  int K;
  int nq_run;                   // layers to run = min(max_layers, nq)
  int D;
  const float* x;               // [n_tokens][D]
  long long n_tokens;
  void* codes;                  // [n_tokens][code_stride] of code_dtype (nullable)
  int code_dtype;               // 0: int16, 1: int32, 2: int64
  long long code_stride;        // elements between consecutive tokens
  float* q_out;                 // [n_tokens][D] (nullable -> encode only)
  const int* teacher;           // nullable: [n_tokens][nq_run] codes that drive the recurrence
  float* z_out;                 // nullable debug: [n_tokens][nq_run][4] in-projection values
  float l2_hot;                 // fraction of the weight stream pinned in L2 with evict_last (rest evict_first)
  unsigned int* sync_ctr;       // nullable: grid lock-step counter of this launch (zeroed before the launch)
  // ---- hook mode (HOOK instantiation; rqae/model.py:276-289 with the Gemma-2 norm / denorm of rqae/llm.py:65-73) ----
  const void* hs;               // hidden states [n_tokens][D] of hs_dtype; the kernel normalises them itself (x unused)
  void* hs_out;                 // nullable: denormalised reconstruction in hs_dtype (may alias hs: same thread, same element)
  const float* rms_w;           // [D] RMSNorm weight w; the norm multiplies by (1 + w)
  float rms_eps;                // eps of the norm (the reference's denorm uses 1e-6, llm.py:71)
  int hs_dtype;                 // 0: fp32, 1: fp16, 2: bf16
  int seq_len, skip_bos;        // token t is left untouched when skip_bos && t % seq_len == 0 (model.py:285-286)
};

__device__ __forceinline__ float hook_load(const void* base, long long i, int dt) {
  if (dt == 1) return __half2float(reinterpret_cast<const __half*>(base)[i]);
  if (dt == 2) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
  return reinterpret_cast<const float*>(base)[i];
}
__device__ __forceinline__ void hook_store(void* base, long long i, int dt, float v) {
  if (dt == 1) reinterpret_cast<__half*>(base)[i] = __float2half_rn(v);
  else if (dt == 2) reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(base)[i] = v;
}

#ifndef RQ_SKEW_CLK
#define RQ_SKEW_CLK 350
#endif
#ifndef RQ_REGC
#define RQ_REGC 224
#define RQ_REGH 56
#endif
constexpr int kComputeWarps = 8;
constexpr int kThreads = 384;
constexpr int kCodeBuf = 16;  // layers buffered per token before a 128-byte code store

// CS > 1: the D-split cluster variant.  A cluster of CS CTAs works on one unit; E, CH then describe ONE CTA's slice of
// the D axis (elements d = (rank * E + j) * 256 + t) and of every stage (chunks rank * CH .. rank * CH + CH - 1).
// Each compute warp hands its in-projection partials to the quantizer warps of EVERY CTA of the cluster (own shared
// memory + st.async into the peers', whose completion counts bytes on the peer's own barrier); every CTA then sums the same CS * 8
// partials in the same order, so all of them derive the same z, the same code and the same c' without exchanging
// anything else.  The partial buffers alternate with the layer's parity: a fast CTA may already deliver layer l + 1
// while a slow one still reads layer l (it cannot get two layers ahead: its layer l + 1 needs everybody's l).
template <int E, int EC, int CH, int NSLOT, int TG, int CS = 1>
struct FwdCfg {
  static constexpr int NP = TG / 2;              // token pairs per phase
  static constexpr int JC = E / CH;              // elements per thread per chunk
  static constexpr int NB = JC / EC;             // register blocks per chunk
  static constexpr int CHUNK_BYTES = JC * RQ_GROUP_THREADS * RQ_BYTES_PER_ELEM;
  static constexpr int OFF_WIN = JC * RQ_GROUP_THREADS * 16;
  static constexpr int OFF_BO = JC * RQ_GROUP_THREADS * 32;
  static_assert(E % CH == 0 && JC % EC == 0 && TG % 2 == 0 && TG <= 8 && CH < NSLOT, "bad shape");
  // shared memory carve-up (bytes)
  static constexpr int SM_RING = 0;
  static constexpr int SM_CBT = NSLOT * CHUNK_BYTES;                        // float4[RQ_SMEM_ROWS] de-duplicated table
  static constexpr int SM_MAP = SM_CBT + RQ_SMEM_ROWS * 16;                 // uint16[RQ_SMEM_ROWS]
  static constexpr int SM_TP = SM_MAP + RQ_SMEM_ROWS * 2;                   // float4[24][RQ_CAN_MAX] canonical rows per order
  static constexpr int SM_MAP3 = SM_TP + RQ_NPERM * RQ_CAN_MAX * 16;        // uint16[16][24][RQ_CAN_MAX]
  static constexpr int PART_WARPS = CS * kComputeWarps;                       // partial rows per (phase, parity)
  static constexpr int PART_SETS = CS > 1 ? 4 : 2;                            // [phase] or [phase][layer parity]
  static constexpr int SM_PART = SM_MAP3 + RQ_NSIGN * RQ_NPERM * RQ_CAN_MAX * 2;  // float[PART_SETS][PART_WARPS][32]
  static constexpr int SM_CPR = SM_PART + PART_SETS * PART_WARPS * 32 * 4;    // u64[2][4][4] (pairs padded to 4)
  static constexpr int SM_CODES = SM_CPR + 2 * 4 * 4 * 8;                   // uint16[2][8][kCodeBuf]
  static constexpr int SM_THR = SM_CODES + 2 * 8 * kCodeBuf * 2;            // float[4] search thresholds
  static constexpr int SM_HOOK = SM_THR + 16;                               // float[8][16] warp partials + float[2][16] scales (hook mode)
  static constexpr int SM_BAR = SM_HOOK + (kComputeWarps * 16 + 32) * 4;    // mbarriers
  static constexpr int N_BAR = 2 * NSLOT + 4;
  static constexpr int SM_TOTAL = SM_BAR + N_BAR * 8;
  static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
};

// Warp reduction of 32 values per lane: lane L ends with the sum over the 32 lanes of *logical* value L
// (token = L >> 2, k = L & 3).  Register v[4*t + c] of lane L holds the partial of token t and logical
// k = c ^ (L & 3): the in-projection weights are stored with their four k components XOR-permuted by the
// owning thread's low lane bits (pack_stages_kernel), so in the two widest exchange steps (lane bits 0 and 1
// against the k bits) every lane keeps the registers whose index bit is 0 and sends those whose bit is 1 --
// no per-lane selects.  The token bits follow with lane bits 4, 3, 2.  Summation tree per logical value:
// lanes paired by strides 1, 2, 16, 8, 4 (oracle: tree order 1).
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) v[i] = __fadd_rn(v[i], __shfl_xor_sync(0xffffffffu, v[i + 1], 1));
#pragma unroll
  for (int i = 0; i < 32; i += 4) v[i] = __fadd_rn(v[i], __shfl_xor_sync(0xffffffffu, v[i + 2], 2));
#pragma unroll
  for (int s = 4; s >= 1; s >>= 1) {      // token stride s <-> lane stride 4*s
    const bool up = (lane & (4 * s)) != 0;
#pragma unroll
    for (int t = 0; t < s; t++) {
      const float keep = up ? v[4 * (t + s)] : v[4 * t];
      const float send = up ? v[4 * t] : v[4 * (t + s)];
      v[4 * t] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 4 * s));
    }
  }
  return v[0];
}

// One pass of phase PH over the CH chunks of one stage (ring slots slot0, slot0+1, ...), as two sweeps over
// the thread's E elements:
//   sweep 1  out-projection of the previous layer's codeword + residual update.  The phase's c' values (NP
//            pairs x 4 packed values) stay in registers for the whole sweep -- the in-projection accumulators
//            are not live yet, so there is room -- and are read from shared memory once per pass.
//   sweep 2  in-projection partials of this layer from the updated residual.
// Phase B releases a ring slot after its sweep-2 reads of it.
template <int PH, int E, int EC, int CH, int NSLOT, int TG>
__device__ __forceinline__ void fwd_pass(u64 (&r2)[TG][E], u64 (&acc)[TG / 2][4], const uint32_t ring,
                                         const uint32_t cpr_ph, uint64_t* full, uint64_t* empty, const uint32_t slot0,
                                         const uint32_t par0, const int ct, const int lane) {
  using C = FwdCfg<E, EC, CH, NSLOT, TG>;
  const u64 neg1 = pack2(-1.0f, -1.0f);
  {
    u64 cp[C::NP][4];
#pragma unroll
    for (int pi = 0; pi < C::NP; pi++) {
      lds128_u64(cpr_ph + pi * 32, cp[pi][0], cp[pi][1]);
      lds128_u64(cpr_ph + pi * 32 + 16, cp[pi][2], cp[pi][3]);
    }
#pragma unroll
    for (int c = 0; c < CH; c++) {
      uint32_t s = slot0 + c, par = par0;
      if (s >= (uint32_t)NSLOT) { s -= NSLOT; par ^= 1; }
      if (PH == 0) mbar_wait(&full[s], par);   // phase B re-reads a chunk this warp has already seen arrive
      const uint32_t sb = ring + s * C::CHUNK_BYTES + ct * 16;
#pragma unroll
      for (int nb = 0; nb < C::NB; nb++) {
        float4 wo[EC];
        float bo[EC];
#pragma unroll
        for (int e = 0; e < EC; e++) {
          const int jj = nb * EC + e;
          wo[e] = lds128(sb + jj * (RQ_GROUP_THREADS * 16));
          bo[e] = lds32(sb + C::OFF_BO - ct * 12 + jj * (RQ_GROUP_THREADS * 4));
        }
#pragma unroll
        for (int pi = 0; pi < C::NP; pi++) {
#pragma unroll
          for (int e = 0; e < EC; e++) {
            const int j = c * C::JC + nb * EC + e;
            u64 o = fma2(pack2(wo[e].x, wo[e].x), cp[pi][0], pack2(bo[e], bo[e]));
            o = fma2(pack2(wo[e].y, wo[e].y), cp[pi][1], o);
            o = fma2(pack2(wo[e].z, wo[e].z), cp[pi][2], o);
            o = fma2(pack2(wo[e].w, wo[e].w), cp[pi][3], o);
            r2[PH * C::NP + pi][j] = fma2(o, neg1, r2[PH * C::NP + pi][j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int pi = 0; pi < C::NP; pi++)
#pragma unroll
    for (int k = 0; k < 4; k++) acc[pi][k] = 0ull;
#pragma unroll
  for (int c = 0; c < CH; c++) {
    uint32_t s = slot0 + c;
    if (s >= (uint32_t)NSLOT) s -= NSLOT;
    const uint32_t sb = ring + s * C::CHUNK_BYTES + ct * 16;
#pragma unroll
    for (int nb = 0; nb < C::NB; nb++) {
      float4 wi[EC];
#pragma unroll
      for (int e = 0; e < EC; e++) wi[e] = lds128(sb + C::OFF_WIN + (nb * EC + e) * (RQ_GROUP_THREADS * 16));
#pragma unroll
      for (int pi = 0; pi < C::NP; pi++) {
#pragma unroll
        for (int e = 0; e < EC; e++) {
          const int j = c * C::JC + nb * EC + e;
          const u64 r = r2[PH * C::NP + pi][j];
          acc[pi][0] = fma2(pack2(wi[e].x, wi[e].x), r, acc[pi][0]);
          acc[pi][1] = fma2(pack2(wi[e].y, wi[e].y), r, acc[pi][1]);
          acc[pi][2] = fma2(pack2(wi[e].z, wi[e].z), r, acc[pi][2]);
          acc[pi][3] = fma2(pack2(wi[e].w, wi[e].w), r, acc[pi][3]);
        }
      }
    }
    if (PH == 1) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

// Register budget: ptxas compiles the kernel for 384 threads/CTA at 168 registers, so the CTA owns a pool
// of 384*168 = 64512 registers; setmaxnreg can only re-split THAT pool (an .inc beyond it spins forever).
template <int E, int EC, int CH, int NSLOT, int TG, bool DBG = false, bool HOOK = false, int CS = 1, int REG_COMPUTE = RQ_REGC,
          int REG_HELPER = RQ_REGH>
__global__ void __launch_bounds__(kThreads, 1) rq_forward_kernel(const FwdParams p) {
  static_assert(2 * 128 * REG_COMPUTE + 128 * REG_HELPER <= kThreads * 168, "register pool budget");
  static_assert(CS >= 1 && CS <= 4 && !(HOOK && CS > 1), "cluster variant: plain forward only");
  using C = FwdCfg<E, EC, CH, NSLOT, TG, CS>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::SM_BAR);
  uint64_t* full = bars;                  // [NSLOT] producer -> compute
  uint64_t* empty = bars + NSLOT;         // [NSLOT] compute -> producer (8 warp arrivals, after phase B)
  uint64_t* part_full = bars + 2 * NSLOT; // [2] compute -> quantizer of the phase (8 warp arrivals)
  uint64_t* c_ready = part_full + 2;      // [2] quantizer -> compute (1 warp arrival)

  // work split: a unit is 2*TG consecutive tokens (phase A = first TG, phase B = next TG); CTA b handles
  // units b, b+grid, ...
  const long long n_units = (p.n_tokens + 2 * TG - 1) / (2 * TG);
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0u;      // this CTA's slice of the D axis
  const long long cl_id = (long long)blockIdx.x / CS;          // unit stream of this CTA (cluster)
  const long long n_cl = (long long)gridDim.x / CS;
  const long long my_iters = (n_units > cl_id) ? (n_units - 1 - cl_id) / n_cl + 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], kComputeWarps); }
    // cluster variant: 8 local warp arrivals + the quantizer's own expect_tx arrival; the peers' partials come as bytes
    for (int g = 0; g < 2; g++) { mbar_init(&part_full[g], CS > 1 ? kComputeWarps + 1 : kComputeWarps); mbar_init(&c_ready[g], 1); }
    mbar_fence_init();
  }
  // search tables -> shared memory (shared-codebook mode): the de-duplicated table if it fits, and the
  // canonical-row tables when the codebook is sign/permutation symmetric (rqae_capi.cu, build_search_tables)
  const RqHeader* hdr = reinterpret_cast<const RqHeader*>(p.packed);
  const int kd_pad = p.cb_shared ? hdr->kd_pad : 0;
  const int can_rows = p.cb_shared ? hdr->can_rows : 0;
  const bool cb_in_smem = p.cb_shared && kd_pad <= RQ_SMEM_ROWS;
  if (cb_in_smem) {
    const float4* src = reinterpret_cast<const float4*>(p.packed + p.off_cbt);
    const unsigned short* msrc = reinterpret_cast<const unsigned short*>(p.packed + p.off_map);
    float4* dst = reinterpret_cast<float4*>(smem + C::SM_CBT);
    unsigned short* mdst = reinterpret_cast<unsigned short*>(smem + C::SM_MAP);
    for (int i = threadIdx.x; i < kd_pad; i += kThreads) { dst[i] = src[i]; mdst[i] = msrc[i]; }
  }
  if (threadIdx.x == 32)
    *reinterpret_cast<float4*>(smem + C::SM_THR) = make_float4(hdr->thr_tiny, hdr->thr_gap, hdr->thr_sep, 0.f);
  if (can_rows > 0) {
    const float4* src = reinterpret_cast<const float4*>(p.packed + p.off_tp);
    const uint32_t* msrc = reinterpret_cast<const uint32_t*>(p.packed + p.off_map3);
    float4* dst = reinterpret_cast<float4*>(smem + C::SM_TP);
    uint32_t* mdst = reinterpret_cast<uint32_t*>(smem + C::SM_MAP3);
    for (int i = threadIdx.x; i < RQ_NPERM * RQ_CAN_MAX; i += kThreads) dst[i] = src[i];
    for (int i = threadIdx.x; i < RQ_NSIGN * RQ_NPERM * RQ_CAN_MAX / 2; i += kThreads) mdst[i] = msrc[i];
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();   // the peers' barriers exist before anybody arrives on them

  if (warp < kComputeWarps) {
    // =============================== compute warps ===============================
    reg_inc<REG_COMPUTE>();
    const int ct = threadIdx.x;              // 0..255
    const uint32_t ring = smem_u32(smem + C::SM_RING);
    const uint32_t cpr = smem_u32(smem + C::SM_CPR);
    const uint32_t part = smem_u32(smem + C::SM_PART) + (crank * kComputeWarps + warp) * 32 * 4 + lane * 4;   // + set * PART_WARPS * 128
    uint32_t part_r[CS > 1 ? CS : 1], pfull_r[CS > 1 ? CS : 1];   // the same slot / the phase barriers in every CTA of the cluster
    if (CS > 1) {
#pragma unroll
      for (int r = 0; r < CS; r++) { part_r[r] = mapa_u32(part, r); pfull_r[r] = mapa_u32(smem_u32(&part_full[0]), r); }
    }
    uint32_t slot = 0, full_par = 0, cr_par = 0;

    for (long long it = 0; it < my_iters; ++it) {
      const long long tok0 = (cl_id + it * n_cl) * (2 * TG);
      // ---- load the unit's activations into registers (zeros outside [0,n_tokens) x [0,D)) ----
      u64 r2[TG][E];   // [phase * NP + pair][element]
      if constexpr (!HOOK) {
#pragma unroll
        for (int pi = 0; pi < TG; pi++) {
          const long long ta = tok0 + 2 * pi, tb = ta + 1;
          const float* xa = p.x + ta * (long long)p.D;
          const float* xb = p.x + tb * (long long)p.D;
#pragma unroll
          for (int j = 0; j < E; j++) {
            const int d = ((int)crank * E + j) * RQ_GROUP_THREADS + ct;
            const float a = (ta < p.n_tokens && d < p.D) ? __ldcs(xa + d) : 0.0f;
            const float b = (tb < p.n_tokens && d < p.D) ? __ldcs(xb + d) : 0.0f;
            r2[pi][j] = pack2(a, b);
          }
        }
      } else {
        // Hook mode: the unit's hidden states are RMS-normalised on the way in (rqae/llm.py:65-66 = Gemma2RMSNorm:
        // x * rsqrt(mean(x^2) + eps) * (1 + w), all in fp32 as the hook's .float() makes it).  Sum of squares: per
        // thread over its E elements, xor-shuffle tree over the warp, the 8 warps as a pairwise tree.
        float ss[2 * TG];
#pragma unroll
        for (int i = 0; i < 2 * TG; i++) ss[i] = 0.0f;
#pragma unroll
        for (int pi = 0; pi < TG; pi++) {
          const long long ta = tok0 + 2 * pi, tb = ta + 1;
#pragma unroll
          for (int j = 0; j < E; j++) {
            const int d = ((int)crank * E + j) * RQ_GROUP_THREADS + ct;
            const float a = (ta < p.n_tokens && d < p.D) ? hook_load(p.hs, ta * (long long)p.D + d, p.hs_dtype) : 0.0f;
            const float b = (tb < p.n_tokens && d < p.D) ? hook_load(p.hs, tb * (long long)p.D + d, p.hs_dtype) : 0.0f;
            r2[pi][j] = pack2(a, b);
            ss[2 * pi] = __fmaf_rn(a, a, ss[2 * pi]);
            ss[2 * pi + 1] = __fmaf_rn(b, b, ss[2 * pi + 1]);
          }
        }
        float* hk = reinterpret_cast<float*>(smem + C::SM_HOOK);
#pragma unroll
        for (int i = 0; i < 2 * TG; i++) {
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) ss[i] = __fadd_rn(ss[i], __shfl_xor_sync(0xffffffffu, ss[i], o));
          if (lane == 0) hk[warp * 16 + i] = ss[i];
        }
        named_bar_sync(1, RQ_GROUP_THREADS);
        if (ct < 2 * TG) {
          const float t = __fadd_rn(__fadd_rn(__fadd_rn(hk[ct], hk[16 + ct]), __fadd_rn(hk[32 + ct], hk[48 + ct])),
                                    __fadd_rn(__fadd_rn(hk[64 + ct], hk[80 + ct]), __fadd_rn(hk[96 + ct], hk[112 + ct])));
          const float ms = __fdiv_rn(t, (float)p.D);
          hk[kComputeWarps * 16 + ct] = rsqrtf(__fadd_rn(ms, p.rms_eps));        // norm
          hk[kComputeWarps * 16 + 16 + ct] = rsqrtf(__fadd_rn(ms, 1e-6f));       // denorm (llm.py:71)
        }
        named_bar_sync(1, RQ_GROUP_THREADS);
#pragma unroll
        for (int j = 0; j < E; j++) {
          const int d = ((int)crank * E + j) * RQ_GROUP_THREADS + ct;
          const float w1 = d < p.D ? __fadd_rn(1.0f, __ldg(p.rms_w + d)) : 0.0f;
#pragma unroll
          for (int pi = 0; pi < TG; pi++) {
            float a, b;
            unpack2(r2[pi][j], a, b);
            a = __fmul_rn(__fmul_rn(a, hk[kComputeWarps * 16 + 2 * pi]), w1);
            b = __fmul_rn(__fmul_rn(b, hk[kComputeWarps * 16 + 2 * pi + 1]), w1);
            r2[pi][j] = pack2(a, b);
          }
        }
      }

      // Pass 0 has no code to project out yet: stage 0 carries W_out = b_out = 0 and the c' slots are zeroed,
      // so o = fma(0, 0, 0) = 0 and r is unchanged.  First barrier: every warp has finished reading the
      // previous unit's last c'; second: the zeros are visible.  (The quantizer warps cannot touch the slots
      // before this unit's pass-0 partials arrive.)
      named_bar_sync(1, RQ_GROUP_THREADS);
      if (ct < 64) sts32(cpr + ct * 4, 0.0f);
      named_bar_sync(1, RQ_GROUP_THREADS);
#if RQ_SKEW_CLK > 0
      // The two compute warps of a scheduler (w and w+4) leave this barrier together and would reach the
      // shuffle-bound reduction at the end of every pass together, leaving the FMA pipe idle twice per
      // pass pair.  Starting warps 4-7 a fraction of a pass late puts one warp's reduction under the other
      // warp's FMA stream; nothing re-aligns them until the next unit (the hand-overs have a pass of slack).
      if (warp >= 4) {
        const long long t0 = clock64();
        while (clock64() - t0 < RQ_SKEW_CLK) {
        }
      }
#endif

      for (int l = 0; l <= p.nq_run; ++l) {
#pragma unroll
        for (int ph = 0; ph < 2; ph++) {
          u64 acc[C::NP][4];
          if (l > 0) mbar_wait(&c_ready[ph], cr_par);
          if (ph == 0)
            fwd_pass<0, E, EC, CH, NSLOT, TG>(r2, acc, ring, cpr, full, empty, slot, full_par, ct, lane);
          else
            fwd_pass<1, E, EC, CH, NSLOT, TG>(r2, acc, ring, cpr + 128, full, empty, slot, full_par, ct, lane);
          if (l < p.nq_run) {
            // logical value index = token * 4 + k, token = 2*pair + half
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = 0.0f;
#pragma unroll
            for (int pi = 0; pi < C::NP; pi++)
#pragma unroll
              for (int k = 0; k < 4; k++) unpack2(acc[pi][k], v[(2 * pi) * 4 + k], v[(2 * pi + 1) * 4 + k]);
            const float s = butterfly32(v, lane);
            if (CS == 1) {
              sts32(part + ph * (kComputeWarps * 32 * 4), s);
              __syncwarp();
              if (lane == 0) mbar_arrive(&part_full[ph]);
            } else {
              const uint32_t set = (uint32_t)(ph * 2 + (l & 1)) * (C::PART_WARPS * 32 * 4);
#pragma unroll
              for (int r = 0; r < CS; r++) {
                if ((uint32_t)r != crank) st_async_cluster_f32(part_r[r] + set, s, pfull_r[r] + ph * 8);
              }
              sts32(part + set, s);
              __syncwarp();
              if (lane == 0) mbar_arrive(&part_full[ph]);
            }
          }
        }
        if (l > 0) cr_par ^= 1;
        slot += CH;
        if (slot >= (uint32_t)NSLOT) { slot -= NSLOT; full_par ^= 1; }
      }

      // ---- reconstruction q = x - r_final ----
      if constexpr (HOOK) {
        // q is denormalised (llm.py:68-73: / (1 + w), then / rsqrt(mean(hs^2) + 1e-6)), the BOS position keeps its
        // hidden state (model.py:285-286) and the result replaces the hidden states in their own dtype (model.py:289).
        // x is recomputed from hs with the operations of the way in, so it is the same value bit for bit.
        if (p.hs_out != nullptr) {
          const float* hk = reinterpret_cast<const float*>(smem + C::SM_HOOK) + kComputeWarps * 16;
#pragma unroll
          for (int j = 0; j < E; j++) {
            const int d = ((int)crank * E + j) * RQ_GROUP_THREADS + ct;
            const float w1 = d < p.D ? __fadd_rn(1.0f, __ldg(p.rms_w + d)) : 1.0f;
#pragma unroll
            for (int pi = 0; pi < TG; pi++) {
              const long long ta = tok0 + 2 * pi, tb = ta + 1;
              float ra, rb;
              unpack2(r2[pi][j], ra, rb);
              if (d < p.D) {
                if (ta < p.n_tokens && !(p.skip_bos && ta % p.seq_len == 0)) {
                  const float x = __fmul_rn(__fmul_rn(hook_load(p.hs, ta * (long long)p.D + d, p.hs_dtype), hk[2 * pi]), w1);
                  hook_store(p.hs_out, ta * (long long)p.D + d, p.hs_dtype,
                             __fdiv_rn(__fdiv_rn(__fsub_rn(x, ra), w1), hk[16 + 2 * pi]));
                }
                if (tb < p.n_tokens && !(p.skip_bos && tb % p.seq_len == 0)) {
                  const float x = __fmul_rn(__fmul_rn(hook_load(p.hs, tb * (long long)p.D + d, p.hs_dtype), hk[2 * pi + 1]), w1);
                  hook_store(p.hs_out, tb * (long long)p.D + d, p.hs_dtype,
                             __fdiv_rn(__fdiv_rn(__fsub_rn(x, rb), w1), hk[16 + 2 * pi + 1]));
                }
              }
            }
          }
        }
      } else if (p.q_out != nullptr) {
#pragma unroll
        for (int pi = 0; pi < TG; pi++) {
          const long long ta = tok0 + 2 * pi, tb = ta + 1;
#pragma unroll
          for (int j = 0; j < E; j++) {
            const int d = ((int)crank * E + j) * RQ_GROUP_THREADS + ct;
            float ra, rb;
            unpack2(r2[pi][j], ra, rb);
            if (d < p.D) {
              if (ta < p.n_tokens) __stcs(p.q_out + ta * (long long)p.D + d, __fsub_rn(__ldcs(p.x + ta * (long long)p.D + d), ra));
              if (tb < p.n_tokens) __stcs(p.q_out + tb * (long long)p.D + d, __fsub_rn(__ldcs(p.x + tb * (long long)p.D + d), rb));
            }
          }
        }
      }
    }
  } else {
    reg_dec<REG_HELPER>();
    const int hw = warp - kComputeWarps;       // 0..3
    if (hw == 2) {
      // =============================== weight producer (warp 10, one lane) ===============================
      // Streams stage chunks L2 -> shared memory with bulk TMA through the NSLOT-deep ring; blocks on the
      // slot's `empty` barrier (hardware-suspended try_wait, no polling).
      if (lane == 0) {
        const uint32_t ring = smem_u32(smem + C::SM_RING);
        const unsigned char* stages = p.packed + p.off_stage + (size_t)crank * CH * (size_t)C::CHUNK_BYTES;
        constexpr size_t kStageBytes = (size_t)CS * CH * (size_t)C::CHUNK_BYTES;
        const uint64_t pol = p.l2_hot >= 1.0f ? l2_policy_evict_last() : l2_policy_hot_fraction(p.l2_hot);
        uint32_t slot = 0, par = 1;  // a fresh barrier passes a wait on parity 1
        unsigned int target = 0;
        for (long long it = 0; it < my_iters; ++it) {
          // Grid lock-step: before streaming the weights of its next unit, every CTA waits until all CTAs
          // that have a unit `it` have finished issuing unit it-1.  Without it the CTAs drift apart over the
          // hundreds of units of a launch, the set of stages in flight grows to the whole 85 MB of weights,
          // which does not stay L2-resident next to the activation streams, and most weight fetches go to
          // HBM.  In step, the stages in flight span a few layers and each stage is read from HBM about
          // once per unit.  Only this lane waits; the compute warps are throttled by the ring.
          if (it > 0 && p.sync_ctr != nullptr) {
            const long long left = n_units - it * n_cl;
            target += (unsigned int)(left < n_cl ? left : n_cl) * CS;
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.sync_ctr) : "memory");
            unsigned int seen;
            do {
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.sync_ctr) : "memory");
              if (seen < target) __nanosleep(200);
            } while (seen < target);
          }
          for (int l = 0; l <= p.nq_run; ++l) {
#pragma unroll 1
            for (int c = 0; c < CH; ++c) {
              mbar_wait(&empty[slot], par);
              mbar_arrive_expect_tx(&full[slot], C::CHUNK_BYTES);
              if (p.l2_hot > 0.f)
                tma_bulk_g2s_hint(ring + slot * C::CHUNK_BYTES, stages + (size_t)l * kStageBytes + (size_t)c * C::CHUNK_BYTES,
                                  C::CHUNK_BYTES, &full[slot], pol);
              else
                tma_bulk_g2s(ring + slot * C::CHUNK_BYTES, stages + (size_t)l * kStageBytes + (size_t)c * C::CHUNK_BYTES,
                             C::CHUNK_BYTES, &full[slot]);
              if (++slot == NSLOT) { slot = 0; par ^= 1; }
            }
          }
        }
      }
    } else if (hw < 2) {
      // =============================== quantizer warps ===============================
      // Warp 8 serves phase A, warp 9 phase B: TG <= 8 tokens per warp, a team of 4 lanes per token; lane
      // `sub` of a team owns coordinate k = sub of z and of c'.
      //
      // Common case (sign/permutation-symmetric codebook; no division or square root on the way to the code):
      // the argmax of cos(n, c_k), n = z/|z|, is the argmax of s_k = z . c_k, and by symmetry it is attained by
      // a canonical row (c0 >= c1 >= c2 >= c3 >= 0) matched to the magnitudes of z in descending order and
      // signed like z.  The team sorts |z| (min/max network), each lane scores 4 of the <= 16 canonical rows
      // and keeps its two largest, and two xor-shuffle steps combine them.  If the best score leads the
      // runner-up by more than thr_gap*|z|, no |z_i| is below thr_tiny*|z| and no two |z_i| are closer than
      // thr_sep*|z| -- margins that the rounding of the reference's normalise-then-dot sequence cannot
      // overturn ("Search"; |z| is bounded by 2*max|z_i| here) -- the leader IS the reference's
      // first maximum.  Otherwise the whole warp runs the reference's exact sequence (IEEE sqrt / divide,
      // fp32 fma chain, first maximum over the whole table).
      //
      // The hand-over to the compute warps (c' in shared memory, c_ready) happens as early as possible; the
      // code index itself is looked up after it.
      const int ph = hw;
      const int sub = lane & 3;                  // lane within the token's team = owned coordinate
      const int tok = lane >> 2;                 // token within the phase
      const bool tok_live = tok < TG;
      const uint32_t cb_smem = smem_u32(smem + C::SM_CBT);
      const uint32_t can_smem = smem_u32(smem + C::SM_TP);    // order 0 = canonical rows as stored
      const uint32_t map3_smem = smem_u32(smem + C::SM_MAP3);
      const uint32_t codes_s = smem_u32(smem + C::SM_CODES) + (ph * 8 + tok) * kCodeBuf * 2;
      const uint32_t pa0 = smem_u32(smem + C::SM_PART) + lane * 4;   // + set * PART_WARPS * 128
      const uint32_t cpr_t = smem_u32(smem + C::SM_CPR) + ph * 128 + (tok >> 1) * 32 + (tok & 1) * 4 + sub * 8;
      uint32_t pf_par = 0;

      for (long long it = 0; it < my_iters; ++it) {
#pragma unroll 1
        for (int l = 0; l < p.nq_run; ++l) {
          // requested before the wait so that the latencies are hidden: this layer's in-projection bias
          // (coordinate `sub`) and the lane's four canonical rows
          const float bsub = __ldg(reinterpret_cast<const float*>(p.packed + p.off_bin) + l * 4 + sub);
          float4 c0, c1, c2, c3;   // canonical rows sub, sub+4, sub+8, sub+12 (zero rows beyond can_rows)
          lds128x4<64>(can_smem + sub * 16, c0, c1, c2, c3);
          if (CS > 1 && lane == 0) mbar_arrive_expect_tx(&part_full[ph], (CS - 1) * kComputeWarps * 32 * 4);   // the peers' partials
          mbar_wait(&part_full[ph], pf_par);
          pf_par ^= 1;
          // z_sub = ((P0 + P1) + (P2 + P3)) + ((P4 + P5) + (P6 + P7)) + b_in   (model.py:211); cluster variant: the
          // same tree per CTA slice, the slices added in rank order, then the bias
          float zm;
          {
            const uint32_t pa = pa0 + (uint32_t)(CS == 1 ? ph : ph * 2 + (l & 1)) * (C::PART_WARPS * 32 * 4);
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < CS; r++) {
              const uint32_t q = pa + r * (kComputeWarps * 32 * 4);
              const float p0 = lds32(q), p1 = lds32(q + 128), p2 = lds32(q + 256), p3 = lds32(q + 384);
              const float p4 = lds32(q + 512), p5 = lds32(q + 640), p6 = lds32(q + 768), p7 = lds32(q + 896);
              const float t = __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3)),
                                        __fadd_rn(__fadd_rn(p4, p5), __fadd_rn(p6, p7)));
              acc = r == 0 ? t : __fadd_rn(acc, t);
            }
            zm = __fadd_rn(acc, bsub);
          }
          const int base = lane & ~3;
          const float z0 = __shfl_sync(0xffffffffu, zm, base), z1 = __shfl_sync(0xffffffffu, zm, base + 1);
          const float z2 = __shfl_sync(0xffffffffu, zm, base + 2), z3 = __shfl_sync(0xffffffffu, zm, base + 3);
          const float a0 = fabsf(z0), a1 = fabsf(z1), a2 = fabsf(z2), a3 = fabsf(z3);
          int code = -1;       // -1: fast path, looked up after the hand-over
          float cwm = 0.f;     // coordinate `sub` of the chosen codeword
          bool fast = false;
          int kw = 0;
          if (can_rows > 0) {
            // |z| in descending order: 5-comparator network (min/max run on the ALU pipe)
            const float h01 = fmaxf(a0, a1), l01 = fminf(a0, a1), h23 = fmaxf(a2, a3), l23 = fminf(a2, a3);
            const float s1 = fmaxf(h01, h23), mx = fminf(h01, h23), my = fmaxf(l01, l23), s4 = fminf(l01, l23);
            const float s2 = fmaxf(mx, my), s3 = fminf(mx, my);
            const float sa = __fmaf_rn(s4, c0.w, __fmaf_rn(s3, c0.z, __fmaf_rn(s2, c0.y, __fmul_rn(s1, c0.x))));
            const float sb = __fmaf_rn(s4, c1.w, __fmaf_rn(s3, c1.z, __fmaf_rn(s2, c1.y, __fmul_rn(s1, c1.x))));
            const float sc = __fmaf_rn(s4, c2.w, __fmaf_rn(s3, c2.z, __fmaf_rn(s2, c2.y, __fmul_rn(s1, c2.x))));
            const float sd = __fmaf_rn(s4, c3.w, __fmaf_rn(s3, c3.z, __fmaf_rn(s2, c3.y, __fmul_rn(s1, c3.x))));
            // largest two of the lane's four scores and the row of the largest (all scores are >= +0)
            const float h1 = fmaxf(sa, sb), l1 = fminf(sa, sb), h2 = fmaxf(sc, sd), l2 = fminf(sc, sd);
            const int p1 = sb > sa ? 4 : 0, p2 = sd > sc ? 12 : 8;
            float m1 = fmaxf(h1, h2);
            float m2 = fmaxf(fminf(h1, h2), fmaxf(l1, l2));
            kw = (h2 > h1 ? p2 : p1) + sub;
            // team combine (lanes 4*tok .. 4*tok+3): best, runner-up and row of the best.  An exact tie for the
            // lead makes runner-up == best, the lead 0, and the token takes the exhaustive path below.
#pragma unroll
            for (int s = 1; s <= 2; s <<= 1) {
              const float om1 = __shfl_xor_sync(0xffffffffu, m1, s);
              const float om2 = __shfl_xor_sync(0xffffffffu, m2, s);
              const int ok = __shfl_xor_sync(0xffffffffu, kw, s);
              m2 = fmaxf(fminf(m1, om1), fmaxf(m2, om2));
              kw = (om1 > m1 || (om1 == m1 && ok < kw)) ? ok : kw;
              m1 = fmaxf(m1, om1);
            }
            // coordinate `sub` of the winner: the canonical value at the rank of |z_sub|, signed like z_sub
            const float am = fabsf(zm);
            const int rank = (a0 > am) + (a1 > am) + (a2 > am) + (a3 > am);
            cwm = copysignf(lds32(can_smem + kw * 16 + rank * 4), zm);
            // validity of the shortcut (NaN / inf / zero input fail these comparisons); |z| <= 2 * s1
            const float4 thr = lds128(smem_u32(smem + C::SM_THR));   // (tiny, gap, sep, -)
            const float nz = s1 + s1;
            const float sep = fminf(fminf(s1 - s2, s2 - s3), s3 - s4);
            const float lead = __fsub_rn(m1, m2);
            fast = lead > thr.y * nz && s4 >= thr.x * nz && sep >= thr.z * nz && s1 >= 1.0e-15f && s1 <= 1.0e15f;
          }
          if (can_rows == 0 || __any_sync(0xffffffffu, !fast)) {
            // ---- the reference's own sequence, whole warp (teams with `fast` keep their result) ----
            // x / x.norm()  (model.py:188): sqrt(((z0^2 + z1^2) + z2^2) + z3^2), IEEE divide
            const float nrm = __fsqrt_rn(__fadd_rn(
                __fadd_rn(__fadd_rn(__fmul_rn(z0, z0), __fmul_rn(z1, z1)), __fmul_rn(z2, z2)), __fmul_rn(z3, z3)));
            const float n0 = __fdiv_rn(z0, nrm), n1 = __fdiv_rn(z1, nrm), n2 = __fdiv_rn(z2, nrm), n3 = __fdiv_rn(z3, nrm);
            // cos = fma(n3,c3, fma(n2,c2, fma(n1,c1, n0*c0)))  (model.py:190); first maximum (model.py:182)
            const int n_rows = p.cb_shared ? kd_pad : p.K;
            const float4* cb_l = p.cb_shared ? reinterpret_cast<const float4*>(p.packed + p.off_cbt)
                                             : reinterpret_cast<const float4*>(p.codebook) + (size_t)l * p.K;
            float va = -INFINITY;
            int ka = 0x7fffffff;
#pragma unroll 2
            for (int k = sub; k < n_rows; k += 4) {
              const float4 c = cb_in_smem ? lds128(cb_smem + k * 16) : __ldg(cb_l + k);
              const float x = __fmaf_rn(n3, c.w, __fmaf_rn(n2, c.z, __fmaf_rn(n1, c.y, __fmul_rn(n0, c.x))));
              if (k == sub || x > va) { va = x; ka = k; }
            }
#pragma unroll
            for (int s = 2; s >= 1; s >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, va, s);
              const int ok = __shfl_xor_sync(0xffffffffu, ka, s);
              if (ov > va || (ov == va && ok < ka)) { va = ov; ka = ok; }
            }
            // a NaN row (z == 0, inf or NaN input) compares false everywhere: the team's lane 0 still holds
            // row 0, which is what torch.argmax returns for an all-NaN row
            ka = __shfl_sync(0xffffffffu, ka, lane & ~3);
            if (!fast) {
              const unsigned short* map_full = cb_in_smem ? reinterpret_cast<const unsigned short*>(smem + C::SM_MAP)
                                                          : reinterpret_cast<const unsigned short*>(p.packed + p.off_map);
              code = p.cb_shared ? (int)map_full[ka] : ka;
              cwm = cb_in_smem ? lds32(cb_smem + ka * 16 + sub * 4) : __ldg(reinterpret_cast<const float*>(cb_l + ka) + sub);
            }
          }
          if (DBG) {  // parity-test instantiation only: let given codes drive the recurrence
            if (p.teacher != nullptr) {
              const long long token = (cl_id + it * n_cl) * (2 * TG) + ph * TG + tok;
              const int tc = (tok_live && token < p.n_tokens) ? p.teacher[token * p.nq_run + l] : 0;
              cwm = p.codebook[((p.cb_shared ? 0 : (size_t)l * p.K) + tc) * 4 + sub];
            }
          }
          // straight-through value c' = z + (c - z)  (model.py:218-220)
          if (tok_live) sts32(cpr_t, __fadd_rn(zm, __fsub_rn(cwm, zm)));
          __syncwarp();
          if (lane == 0) mbar_arrive(&c_ready[ph]);

          // ---- after the hand-over: the code index of the fast path ----
          if (code < 0) {
            // magnitude order as a Lehmer code (must match lehmer_order() in rqae_capi.cu) and sign pattern
            const int ord = 6 * ((a1 > a0) + (a2 > a0) + (a3 > a0)) + 2 * ((a2 > a1) + (a3 > a1)) + (a3 > a2);
            const int sidx = (z0 < 0.f ? 1 : 0) | (z1 < 0.f ? 2 : 0) | (z2 < 0.f ? 4 : 0) | (z3 < 0.f ? 8 : 0);
            unsigned short cs;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(cs) : "r"(map3_smem + ((sidx * RQ_NPERM + ord) * RQ_CAN_MAX + kw) * 2));
            code = (int)cs;
          }
          if (DBG) {
            const long long token = (cl_id + it * n_cl) * (2 * TG) + ph * TG + tok;
            if (p.z_out != nullptr && crank == 0 && sub == 0 && tok_live && token < p.n_tokens)
              reinterpret_cast<float4*>(p.z_out)[token * p.nq_run + l] = make_float4(z0, z1, z2, z3);
          }
          if (sub == 0 && tok_live)
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(codes_s + (l & (kCodeBuf - 1)) * 2), "h"((unsigned short)code) : "memory");
          // ---- flush buffered codes: 16 consecutive layers of one token = one 128-byte store ----
          if (p.codes != nullptr && crank == 0 && ((l & (kCodeBuf - 1)) == kCodeBuf - 1 || l == p.nq_run - 1)) {
            __syncwarp();
            const int l0 = l & ~(kCodeBuf - 1);
            const long long token = (cl_id + it * n_cl) * (2 * TG) + ph * TG + tok;
            const bool tok_valid = tok_live && token < p.n_tokens;
#pragma unroll
            for (int r = 0; r < 4; r++) {
              const int li = sub * 4 + r;
              if (li <= l - l0 && tok_valid) {
                unsigned short cval;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(cval) : "r"(codes_s + li * 2));
                const long long off = token * p.code_stride + l0 + li;
                if (p.code_dtype == 2) reinterpret_cast<long long*>(p.codes)[off] = (long long)cval;
                else if (p.code_dtype == 1) reinterpret_cast<int*>(p.codes)[off] = (int)cval;
                else reinterpret_cast<short*>(p.codes)[off] = (short)cval;
              }
            }
            __syncwarp();
          }
        }
      }
    }
  }
  if (CS > 1) {   // nobody leaves while a peer may still store into its shared memory or arrive on its barriers
    __syncwarp();
    cluster_sync_all();
  }
}

}  // namespace rq
