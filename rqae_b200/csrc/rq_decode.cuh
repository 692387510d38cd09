// RQAE decode kernel for sm_100a: q[t] = sum over selected layers l (ascending) of
// W_out[l] * c[t][l] + b_out[l], c[t][l] = codebook[0][codes[t][l]].
//
// Replaces RQAE.indices_to_codebook_values + decode_from_codebook_values (rqae/model.py:232-252):
// a [B,S,nq,4] fp32 gather materialised in memory followed by nq rank-4 F.linear calls and nq
// in-place adds.  Here the accumulator q[TGD tokens][D] stays in registers for all layers, the
// W_out / b_out halves of the packed stages (rq_layout.h) are streamed through shared memory with
// bulk TMA, and the per-element arithmetic is the reference's own fp32 sequence
//     o = fma(c3,w3, fma(c2,w2, fma(c1,w1, c0*w0))) + b ;  q = q + o
// (what torch's CPU F.linear computes for K=4, probed -- see DESIGN.md), two tokens per packed
// FFMA2/FADD2 instruction, so the result is bit-identical to the reference up to the sign of zero.
//
// CTA = 384 threads:
//   warps 0-7   compute : 256 threads, thread t owns d = j*256 + t (j < E) for all TGD tokens of the unit
//   warp  8     producer: one lane, bulk-TMA of the W_out / b_out parts of each selected stage into the ring
//   warps 9-11  stagers : gather the codewords of the next kDecLB layers for the unit's tokens from the code
//                         tensor (or copy them from `cv`) into a double-buffered shared table, so the compute
//                         warps never touch global memory or a CTA-wide barrier inside the layer loop
#pragma once
#include "rq_common.cuh"
#include "rq_layout.h"

namespace rq {

struct DecParams {
  const unsigned char* packed;
  unsigned long long off_stage, stage_bytes;
  const float* codebook0;  // [K][4]
  int K, nq_codes, D, E, CH;
  const void* codes;
  int code_dtype;
  long long code_stride;
  const float* cv;             // nullable [n][nq_codes][4]
  const unsigned char* layer_mask;  // nullable [nq_codes...]
  long long n_tokens;
  float* q_out;
};

constexpr int kDecThreads = 384;
constexpr int kDecLB = 16;      // layers of codewords per staged block
constexpr int kDecStagers = 3;  // warps 9..11

template <int E, int EC, int CH, int NSLOT, int TGD>
struct DecCfg {
  static constexpr int NP = TGD / 2;
  static constexpr int JC = E / CH;
  static constexpr int NB = JC / EC;
  static constexpr int W_BYTES = JC * RQ_GROUP_THREADS * 16;
  static constexpr int B_BYTES = JC * RQ_GROUP_THREADS * 4;
  static constexpr int SLOT_BYTES = W_BYTES + B_BYTES;
  static constexpr int CST_BYTES = kDecLB * NP * 4 * 8;                   // u64[kDecLB][NP][4]
  static constexpr int SM_RING = 0;
  static constexpr int SM_CST = NSLOT * SLOT_BYTES;                       // two CST buffers
  static constexpr int SM_BAR = SM_CST + 2 * CST_BYTES;
  static constexpr int SM_TOTAL = SM_BAR + (2 * NSLOT + 4) * 8;
  static_assert(E % CH == 0 && JC % EC == 0 && TGD % 2 == 0, "bad shape");
  static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
};

template <int E, int EC, int CH, int NSLOT, int TGD>
__global__ void __launch_bounds__(kDecThreads, 1) rq_decode_kernel(const DecParams p) {
  using C = DecCfg<E, EC, CH, NSLOT, TGD>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::SM_BAR);
  uint64_t* empty = full + NSLOT;
  uint64_t* cst_full = empty + NSLOT;   // [2] stagers -> compute (kDecStagers warp arrivals)
  uint64_t* cst_empty = cst_full + 2;   // [2] compute -> stagers (8 warp arrivals)
  const long long n_units = (p.n_tokens + TGD - 1) / TGD;
  const long long my_iters = (n_units > (long long)blockIdx.x) ? (n_units - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
    for (int i = 0; i < 2; i++) { mbar_init(&cst_full[i], kDecStagers); mbar_init(&cst_empty[i], 8); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= 8) {
    reg_dec<40>();
    if (warp == 8) {
      // ---- weight producer ----
      if (lane == 0) {
        const uint32_t ring = smem_u32(smem + C::SM_RING);
        const uint64_t pol = l2_policy_evict_last();
        uint32_t slot = 0, par = 1;
        for (long long it = 0; it < my_iters; ++it) {
          for (int l = 0; l < p.nq_codes; ++l) {
            if (p.layer_mask != nullptr && p.layer_mask[l] == 0) continue;
            // W_out[l], b_out[l] live in stage l+1
            const unsigned char* st = p.packed + p.off_stage + (size_t)(l + 1) * p.stage_bytes;
            for (int c = 0; c < CH; c++) {
              const unsigned char* chunk = st + (size_t)c * (p.stage_bytes / CH);
              mbar_wait(&empty[slot], par);
              mbar_arrive_expect_tx(&full[slot], C::SLOT_BYTES);
              const uint32_t dst = ring + slot * C::SLOT_BYTES;
              tma_bulk_g2s_hint(dst, chunk, C::W_BYTES, &full[slot], pol);
              tma_bulk_g2s_hint(dst + C::W_BYTES, chunk + 2 * (size_t)C::W_BYTES, C::B_BYTES, &full[slot], pol);
              if (++slot == NSLOT) { slot = 0; par ^= 1; }
            }
          }
        }
      }
    } else {
      // ---- codeword stagers ----
      const int st_tid = (warp - 9) * 32 + lane;
      uint32_t nblk = 0;
      for (long long it = 0; it < my_iters; ++it) {
        const long long tok_cta = ((long long)blockIdx.x + it * gridDim.x) * TGD;
        for (int l0 = 0; l0 < p.nq_codes; l0 += kDecLB, ++nblk) {
          const uint32_t buf = nblk & 1;
          mbar_wait(&cst_empty[buf], ((nblk >> 1) & 1) ^ 1);   // a fresh barrier passes a wait on parity 1
          float* cst = reinterpret_cast<float*>(smem + C::SM_CST + buf * C::CST_BYTES);
          for (int i = st_tid; i < kDecLB * TGD; i += kDecStagers * 32) {
            const int tk = i / kDecLB, li = i % kDecLB;  // consecutive threads -> consecutive layers of a token
            const long long token = tok_cta + tk;
            const int l = l0 + li;
            float4 cw = make_float4(0.f, 0.f, 0.f, 0.f);
            if (token < p.n_tokens && l < p.nq_codes) {
              if (p.cv != nullptr) {
                cw = __ldg(reinterpret_cast<const float4*>(p.cv) + token * p.nq_codes + l);
              } else {
                long long idx;
                const long long off = token * p.code_stride + l;
                if (p.code_dtype == 2) idx = __ldcs(reinterpret_cast<const long long*>(p.codes) + off);
                else if (p.code_dtype == 1) idx = __ldcs(reinterpret_cast<const int*>(p.codes) + off);
                else idx = __ldcs(reinterpret_cast<const short*>(p.codes) + off);
                if (idx < 0) idx += p.K;  // python-style negative index, as torch indexing
                if (idx >= 0 && idx < p.K) cw = __ldg(reinterpret_cast<const float4*>(p.codebook0) + idx);
              }
            }
            float* dst = cst + (((li * C::NP + (tk >> 1)) * 4) * 2 + (tk & 1));
            dst[0] = cw.x; dst[2] = cw.y; dst[4] = cw.z; dst[6] = cw.w;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&cst_full[buf]);
        }
      }
    }
    return;
  }

  reg_inc<232>();
  const int ct = threadIdx.x;
  const uint32_t ring = smem_u32(smem + C::SM_RING);
  uint32_t slot = 0, full_par = 0, nblk = 0;

  for (long long it = 0; it < my_iters; ++it) {
    const long long tok0 = ((long long)blockIdx.x + it * gridDim.x) * TGD;
    u64 q2[C::NP][E];
#pragma unroll
    for (int pi = 0; pi < C::NP; pi++)
#pragma unroll
      for (int j = 0; j < E; j++) q2[pi][j] = 0ull;

    for (int l0 = 0; l0 < p.nq_codes; l0 += kDecLB, ++nblk) {
      const uint32_t buf = nblk & 1;
      mbar_wait(&cst_full[buf], (nblk >> 1) & 1);
      const uint32_t cst = smem_u32(smem + C::SM_CST) + buf * C::CST_BYTES;
      const int l1 = (l0 + kDecLB < p.nq_codes) ? l0 + kDecLB : p.nq_codes;
      for (int l = l0; l < l1; ++l) {
        if (p.layer_mask != nullptr && p.layer_mask[l] == 0) continue;
        const uint32_t cl = cst + (l - l0) * (C::NP * 32);
#pragma unroll
        for (int c = 0; c < CH; c++) {
          mbar_wait(&full[slot], full_par);
          const uint32_t sb = ring + slot * C::SLOT_BYTES + ct * 16;
#pragma unroll
          for (int nb = 0; nb < C::NB; nb++) {
            float4 wo[EC];
            float bo[EC];
#pragma unroll
            for (int e = 0; e < EC; e++) {
              const int jj = nb * EC + e;
              wo[e] = lds128(sb + jj * (RQ_GROUP_THREADS * 16));
              bo[e] = lds32(sb + C::W_BYTES - ct * 12 + jj * (RQ_GROUP_THREADS * 4));
            }
#pragma unroll
            for (int pi = 0; pi < C::NP; pi++) {
              u64 c0, c1, c2, c3;
              lds128_u64(cl + pi * 32, c0, c1);
              lds128_u64(cl + pi * 32 + 16, c2, c3);
#pragma unroll
              for (int e = 0; e < EC; e++) {
                const int j = c * C::JC + nb * EC + e;
                u64 o = mul2(c0, pack2(wo[e].x, wo[e].x));
                o = fma2(c1, pack2(wo[e].y, wo[e].y), o);
                o = fma2(c2, pack2(wo[e].z, wo[e].z), o);
                o = fma2(c3, pack2(wo[e].w, wo[e].w), o);
                o = add2(o, pack2(bo[e], bo[e]));
                q2[pi][j] = add2(q2[pi][j], o);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[slot]);
          if (++slot == NSLOT) { slot = 0; full_par ^= 1; }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&cst_empty[buf]);
    }

#pragma unroll
    for (int pi = 0; pi < C::NP; pi++) {
      const long long ta = tok0 + 2 * pi, tb = ta + 1;
#pragma unroll
      for (int j = 0; j < E; j++) {
        const int d = j * RQ_GROUP_THREADS + ct;
        float a, b;
        unpack2(q2[pi][j], a, b);
        if (d < p.D) {
          if (ta < p.n_tokens) __stcs(p.q_out + ta * (long long)p.D + d, a);
          if (tb < p.n_tokens) __stcs(p.q_out + tb * (long long)p.D + d, b);
        }
      }
    }
  }
}

template <int E, int EC, int CH, int NSLOT, int TGD>
inline int launch_decode_t(const DecParams& prm, int sms, cudaStream_t st) {
  using C = DecCfg<E, EC, CH, NSLOT, TGD>;
  auto kern = rq_decode_kernel<E, EC, CH, NSLOT, TGD>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SM_TOTAL);
  if (e != cudaSuccess) return 3;
  const long long n_units = (prm.n_tokens + TGD - 1) / TGD;
  const int grid = (int)(n_units < sms ? n_units : sms);
  kern<<<grid, kDecThreads, C::SM_TOTAL, st>>>(prm);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

inline int launch_decode(const DecParams& prm, int sms, cudaStream_t st) {
  // CH must match the packed layout (rq_pick_shape); E selects the instantiation.
  // TGD tokens per unit: the accumulator q2[TGD/2][E] must fit the 232-register budget (TGD * E <= 168)
  switch (prm.E) {
    case 1: return launch_decode_t<1, 1, 1, 8, 16>(prm, sms, st);
    case 3: return launch_decode_t<3, 3, 1, 6, 16>(prm, sms, st);
    case 6: return launch_decode_t<6, 3, 2, 6, 16>(prm, sms, st);
    case 9: return launch_decode_t<9, 3, RQ_E9_CH, RQ_E9_DEC_NSLOT, 16>(prm, sms, st);
    case 14: return launch_decode_t<14, 2, 7, 12, 12>(prm, sms, st);
    default: return 2;
  }
}

}  // namespace rq
