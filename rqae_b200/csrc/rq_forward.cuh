// Fused RQAE forward / encode kernel for sm_100a (B200).
//
// Replaces the per-layer op sequence of rqae/model.py:199-230 (reference, harish-kamath/rqae):
// F.linear(D->4), L2-normalise, [T,4]x[4,K] cos-sim matmul, argmax, gather, straight-through
// add/sub, F.linear(4->D), residual subtract, reconstruction add -- ~13 launches per layer,
// 1024 layers -- by ONE persistent kernel in which the residual never leaves registers.
//
// CTA = 384 threads, one CTA per SM (the register file is the capacity that matters):
//   warps 0-3   compute group A : TG tokens, residual r[TG][D] in registers (D split over 128 threads)
//   warps 4-7   compute group B : TG other tokens; runs half a layer out of phase with A so that
//                                 one group's argmax latency is hidden behind the other's FMA work
//   warps 8-10  quantizer       : cross-warp sum of the in-projection partials, cos-sim argmax over
//                                 the de-duplicated codebook (lowest original index wins ties, NaN
//                                 -> index 0 as torch.argmax), straight-through value, code output
//   warp  11    producer        : one lane streaming weight chunks L2 -> shared memory with bulk TMA
//                                 (cp.async.bulk + mbarrier complete_tx) through an NSLOT-deep ring
// Registers are re-split with setmaxnreg: compute warpgroups grow, the helper warpgroup shrinks.
//
// One *pass* of a compute group over stage s (see rq_layout.h) does, per owned element d and per
// token pair, with packed FFMA2 (two tokens per instruction, each half IEEE fp32 round-to-nearest):
//     o   = fma(w_out[d][3], c'3, fma(w_out[d][2], c'2, fma(w_out[d][1], c'1, fma(w_out[d][0], c'0, b_out[d]))))
//     r_d = r_d - o                                  (model.py:221-223)
//     acc[k] = fma(w_in[k][d], r_d, acc[k]), k<4     (model.py:211, partial over the thread's elements)
// then reduces acc over the warp with a shuffle butterfly and hands 4 per-warp partials per (token, k)
// to the quantizer warps through shared memory.  Summation order is fixed (thread-sequential over j,
// lane tree with strides 16,8,4,2,1, warps 0..3 sequentially, then + b_in), so results do not depend on
// the tile a token lands in, on the grid size or on timing.  The reconstruction is emitted as
// q = x - r_final (one extra read of x) instead of a second register-resident accumulator.
#pragma once
#include "rq_common.cuh"
#include "rq_layout.h"

namespace rq {

struct FwdParams {
  const unsigned char* packed;  // packed buffer (rq_layout.h)
  unsigned long long off_bin, off_cbt, off_map, off_ort, off_ortmap, off_stage;
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
};

#ifndef RQ_REGC
#define RQ_REGC 224
#define RQ_REGH 56
#endif
constexpr int kComputeWarps = 8;
constexpr int kQuantWarps = 4;     // two per compute group
constexpr int kThreads = 384;
constexpr int kCodeBuf = 16;  // layers buffered per token before a 128-byte code store

template <int E, int EC, int CH, int NSLOT, int TG>
struct FwdCfg {
  static constexpr int NP = TG / 2;              // token pairs per group
  static constexpr int JC = E / CH;              // elements per thread per chunk
  static constexpr int NB = JC / EC;             // register blocks per chunk
  static constexpr int CHUNK_BYTES = JC * RQ_GROUP_THREADS * RQ_BYTES_PER_ELEM;
  static constexpr int OFF_WIN = JC * RQ_GROUP_THREADS * 16;
  static constexpr int OFF_BO = JC * RQ_GROUP_THREADS * 32;
  static_assert(E % CH == 0 && JC % EC == 0 && TG % 2 == 0 && TG <= 8, "bad shape");
  // shared memory carve-up (bytes)
  static constexpr int SM_RING = 0;
  static constexpr int SM_CBT = NSLOT * CHUNK_BYTES;                        // float4[RQ_SMEM_ROWS]: orthant lists or full table
  static constexpr int SM_MAP = SM_CBT + RQ_SMEM_ROWS * 16;                 // uint16[RQ_SMEM_ROWS]
  static constexpr int SM_PART = SM_MAP + RQ_SMEM_ROWS * 2;                 // float[2][4][32]
  static constexpr int SM_CPR = SM_PART + 2 * 4 * 32 * 4;                   // u64[2][NP][4]
  static constexpr int SM_CODES = SM_CPR + 2 * 4 * 4 * 8;                   // uint16[2][8][kCodeBuf]
  static constexpr int SM_BAR = SM_CODES + 2 * 8 * kCodeBuf * 2;            // mbarriers
  static constexpr int N_BAR = 2 * NSLOT + 4;
  static constexpr int SM_TOTAL = SM_BAR + N_BAR * 8;
  static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");
};

// lane L ends with the sum over the 32 lanes of logical value L (see file header for the order)
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; i++) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, s));
    }
  }
  return v[0];
}

// Register budget: ptxas compiles the kernel for 384 threads/CTA at 168 registers, so the CTA owns a pool
// of 384*168 = 64512 registers; setmaxnreg can only re-split THAT pool (an .inc beyond it spins forever).
template <int E, int EC, int CH, int NSLOT, int TG, bool DBG = false, int REG_COMPUTE = RQ_REGC, int REG_HELPER = RQ_REGH>
__global__ void __launch_bounds__(kThreads, 1) rq_forward_kernel(const FwdParams p) {
  static_assert(2 * 128 * REG_COMPUTE + 128 * REG_HELPER <= kThreads * 168, "register pool budget");
  using C = FwdCfg<E, EC, CH, NSLOT, TG>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::SM_BAR);
  uint64_t* full = bars;                  // [NSLOT] producer -> compute
  uint64_t* empty = bars + NSLOT;         // [NSLOT] compute -> producer (8 warp arrivals)
  uint64_t* part_full = bars + 2 * NSLOT; // [2] compute group -> quantizer (4 warp arrivals)
  uint64_t* c_ready = part_full + 2;      // [2] quantizer -> compute group (2 warp arrivals)

  // work split: a unit is TG consecutive tokens; CTA b handles unit pairs b, b+grid, ...
  const long long n_units = (p.n_tokens + TG - 1) / TG;
  const long long n_pairs = (n_units + 1) / 2;
  const long long my_iters = (n_pairs > (long long)blockIdx.x) ? (n_pairs - 1 - blockIdx.x) / gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], kComputeWarps); }
    for (int g = 0; g < 2; g++) { mbar_init(&part_full[g], 4); mbar_init(&c_ready[g], 2); }
    mbar_fence_init();
  }
  // search tables -> shared memory (shared-codebook mode): the 16 sign-orthant lists when the table is
  // sign-symmetric (the full table then stays in global memory / L2 for the rare full scan), else the whole
  // de-duplicated table if it fits
  const RqHeader* hdr = reinterpret_cast<const RqHeader*>(p.packed);
  const int kd_pad = p.cb_shared ? hdr->kd_pad : 0;
  const int ort_rows = p.cb_shared ? hdr->ort_rows : 0;
  const float ort_thr = hdr->ort_thr;
  const bool cb_in_smem = p.cb_shared && ort_rows == 0 && kd_pad <= RQ_SMEM_ROWS;
  if (ort_rows > 0 || cb_in_smem) {
    const float4* src = reinterpret_cast<const float4*>(p.packed + (ort_rows > 0 ? p.off_ort : p.off_cbt));
    const unsigned short* msrc = reinterpret_cast<const unsigned short*>(p.packed + (ort_rows > 0 ? p.off_ortmap : p.off_map));
    float4* dst = reinterpret_cast<float4*>(smem + C::SM_CBT);
    unsigned short* mdst = reinterpret_cast<unsigned short*>(smem + C::SM_MAP);
    const int nrow = ort_rows > 0 ? RQ_SMEM_ROWS : kd_pad;
    for (int i = threadIdx.x; i < nrow; i += kThreads) { dst[i] = src[i]; mdst[i] = msrc[i]; }
  }
  __syncthreads();

  if (warp < kComputeWarps) {
    // =============================== compute groups ===============================
    reg_inc<REG_COMPUTE>();
    const int g = warp >> 2;
    const int wg = warp & 3;                 // warp within group
    const int tg = threadIdx.x & 127;        // thread within group
    const uint32_t ring = smem_u32(smem + C::SM_RING);
    const uint32_t cpr = smem_u32(smem + C::SM_CPR) + g * (4 * 4 * 8);
    const uint32_t part = smem_u32(smem + C::SM_PART) + (g * 4 + wg) * 32 * 4 + lane * 4;
    const u64 neg1 = pack2(-1.0f, -1.0f);
    uint32_t slot = 0, full_par = 0, cr_par = 0;

    for (long long it = 0; it < my_iters; ++it) {
      const long long unit = 2 * ((long long)blockIdx.x + it * gridDim.x) + g;
      const long long tok0 = unit * TG;
      // ---- load the unit's activations into registers (zeros outside [0,n_tokens) x [0,D)) ----
      u64 r2[C::NP][E];
#pragma unroll
      for (int pi = 0; pi < C::NP; pi++) {
        const long long ta = tok0 + 2 * pi, tb = ta + 1;
        const float* xa = p.x + ta * (long long)p.D;
        const float* xb = p.x + tb * (long long)p.D;
#pragma unroll
        for (int j = 0; j < E; j++) {
          const int d = j * RQ_GROUP_THREADS + tg;
          const float a = (ta < p.n_tokens && d < p.D) ? __ldcs(xa + d) : 0.0f;
          const float b = (tb < p.n_tokens && d < p.D) ? __ldcs(xb + d) : 0.0f;
          r2[pi][j] = pack2(a, b);
        }
      }

      // Pass 0 has no code to project out yet: stage 0 carries W_out = b_out = 0 and the group zeroes
      // its c' slots, so o = fma(0, 0, 0) = 0 and r is unchanged.  The quantizer warps cannot touch the
      // slots again before this group's pass-0 partials arrive, hence a group-local barrier suffices.
      if (tg < 32) sts32(cpr + tg * 4, 0.0f);
      named_bar_sync(1 + g, RQ_GROUP_THREADS);

      for (int l = 0; l <= p.nq_run; ++l) {
        u64 acc[C::NP][4];
#pragma unroll
        for (int pi = 0; pi < C::NP; pi++)
#pragma unroll
          for (int k = 0; k < 4; k++) acc[pi][k] = 0ull;
        if (l > 0) { mbar_wait(&c_ready[g], cr_par); cr_par ^= 1; }

#pragma unroll
        for (int c = 0; c < CH; c++) {
          mbar_wait(&full[slot], full_par);
          const uint32_t sb = ring + slot * C::CHUNK_BYTES + tg * 16;
#pragma unroll
          for (int nb = 0; nb < C::NB; nb++) {
            float4 wo[EC], wi[EC];
            float bo[EC];
#pragma unroll
            for (int e = 0; e < EC; e++) {
              const int jj = nb * EC + e;
              wo[e] = lds128(sb + jj * (RQ_GROUP_THREADS * 16));
              wi[e] = lds128(sb + C::OFF_WIN + jj * (RQ_GROUP_THREADS * 16));
              bo[e] = lds32(sb + C::OFF_BO - tg * 12 + jj * (RQ_GROUP_THREADS * 4));
            }
#pragma unroll
            for (int pi = 0; pi < C::NP; pi++) {
              u64 c0, c1, c2, c3;
              lds128_u64(cpr + pi * 32, c0, c1);
              lds128_u64(cpr + pi * 32 + 16, c2, c3);
#pragma unroll
              for (int e = 0; e < EC; e++) {
                const int j = c * C::JC + nb * EC + e;
                u64 o = fma2(pack2(wo[e].x, wo[e].x), c0, pack2(bo[e], bo[e]));
                o = fma2(pack2(wo[e].y, wo[e].y), c1, o);
                o = fma2(pack2(wo[e].z, wo[e].z), c2, o);
                o = fma2(pack2(wo[e].w, wo[e].w), c3, o);
                const u64 r = fma2(o, neg1, r2[pi][j]);
                r2[pi][j] = r;
                acc[pi][0] = fma2(pack2(wi[e].x, wi[e].x), r, acc[pi][0]);
                acc[pi][1] = fma2(pack2(wi[e].y, wi[e].y), r, acc[pi][1]);
                acc[pi][2] = fma2(pack2(wi[e].z, wi[e].z), r, acc[pi][2]);
                acc[pi][3] = fma2(pack2(wi[e].w, wi[e].w), r, acc[pi][3]);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[slot]);
          if (++slot == NSLOT) { slot = 0; full_par ^= 1; }
        }

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
          sts32(part, s);
          __syncwarp();
          if (lane == 0) mbar_arrive(&part_full[g]);
        }
      }

      // ---- reconstruction q = x - r_final ----
      if (p.q_out != nullptr) {
#pragma unroll
        for (int pi = 0; pi < C::NP; pi++) {
          const long long ta = tok0 + 2 * pi, tb = ta + 1;
#pragma unroll
          for (int j = 0; j < E; j++) {
            const int d = j * RQ_GROUP_THREADS + tg;
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
    // =============================== quantizer warps (+ weight producer) ===============================
    // Warps 8,9 serve group A, warps 10,11 serve group B.  A warp handles 4 tokens at once, 8 lanes per
    // token: each lane scans every 8th row of the search table with two independent running maxima, then
    // the 8 lanes combine with 3 shuffle steps ((value desc, row asc) ordering == first maximum).
    // Lane 0 of warp 11 doubles as the weight producer: whenever it is about to wait it first tops up the
    // bulk-TMA ring (non-blocking mbarrier.test_wait on the slot's `empty` barrier).
    const int hw = warp - kComputeWarps;       // 0..3
    const int g = hw >> 1;
    const int sub = lane & 7;                  // lane within the token's 8-lane team
    const int tok = (hw & 1) * 4 + (lane >> 3);  // token within the group handled by this team
    const bool tok_live = tok < TG;
    const bool is_producer = (hw == 3 && lane == 0);
    const uint32_t ring = smem_u32(smem + C::SM_RING);
    const unsigned char* stages = p.packed + p.off_stage;
    const uint64_t pol = l2_policy_evict_last();
    // producer state: next chunk to fetch (it_f, l_f, c_f) -> ring slot slot_f; `empty` parity par_f
    long long it_f = 0;
    int l_f = 0, c_f = 0;
    uint32_t slot_f = 0, par_f = 1;  // a fresh barrier passes a wait on parity 1
    auto top_up = [&]() {
      while (it_f < my_iters) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(&empty[slot_f])), "r"(par_f) : "memory");
        if (!ok) break;
        mbar_arrive_expect_tx(&full[slot_f], C::CHUNK_BYTES);
        tma_bulk_g2s_hint(ring + slot_f * C::CHUNK_BYTES,
                          stages + ((size_t)l_f * CH + c_f) * (size_t)C::CHUNK_BYTES, C::CHUNK_BYTES, &full[slot_f], pol);
        if (++slot_f == NSLOT) { slot_f = 0; par_f ^= 1; }
        if (++c_f == CH) { c_f = 0; if (++l_f > p.nq_run) { l_f = 0; ++it_f; } }
      }
    };

    const uint32_t cb_smem = smem_u32(smem + C::SM_CBT);
    const float4* cb_glob = reinterpret_cast<const float4*>(p.packed + p.off_cbt);
    const unsigned short* map_s = reinterpret_cast<const unsigned short*>(smem + C::SM_MAP);   // orthant lists / full table in smem
    const unsigned short* map_full = cb_in_smem ? map_s : reinterpret_cast<const unsigned short*>(p.packed + p.off_map);
    const uint32_t codes_s = smem_u32(smem + C::SM_CODES) + (g * 8 + tok) * kCodeBuf * 2;
    const uint32_t pa = smem_u32(smem + C::SM_PART) + (g * 4 * 32 + tok * 4) * 4;
    const uint32_t cpr_t = smem_u32(smem + C::SM_CPR) + g * 128 + (tok >> 1) * 32 + (tok & 1) * 4;
    const int n_rows = p.cb_shared ? kd_pad : p.K;
    const unsigned team_lane0 = lane & ~7;
    uint32_t pf_par = 0;

    for (long long it = 0; it < my_iters; ++it) {
      const long long token = (2 * ((long long)blockIdx.x + it * gridDim.x) + g) * TG + tok;
      const bool tok_valid = tok_live && token < p.n_tokens;
#pragma unroll 1
      for (int l = 0; l < p.nq_run; ++l) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.packed + p.off_bin) + l);
        const float4* cb_l = p.cb_shared ? cb_glob : reinterpret_cast<const float4*>(p.codebook) + (size_t)l * p.K;
        if (is_producer) top_up();
        while (!mbar_try_wait(&part_full[g], pf_par)) {
          if (is_producer) top_up();
        }
        pf_par ^= 1;
        // z = (((P0 + P1) + P2) + P3) + b_in   (model.py:211)
        const float4 s0 = lds128(pa), s1 = lds128(pa + 128), s2 = lds128(pa + 256), s3 = lds128(pa + 384);
        const float z0 = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s0.x, s1.x), s2.x), s3.x), b4.x);
        const float z1 = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s0.y, s1.y), s2.y), s3.y), b4.y);
        const float z2 = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s0.z, s1.z), s2.z), s3.z), b4.z);
        const float z3 = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s0.w, s1.w), s2.w), s3.w), b4.w);
        // x / x.norm()  (model.py:188): sqrt(((z0^2 + z1^2) + z2^2) + z3^2), IEEE divide
        const float nrm = __fsqrt_rn(__fadd_rn(
            __fadd_rn(__fadd_rn(__fmul_rn(z0, z0), __fmul_rn(z1, z1)), __fmul_rn(z2, z2)), __fmul_rn(z3, z3)));
        const float n0 = __fdiv_rn(z0, nrm), n1 = __fdiv_rn(z1, nrm), n2 = __fdiv_rn(z2, nrm), n3 = __fdiv_rn(z3, nrm);
        // cos = fma(n3,c3, fma(n2,c2, fma(n1,c1, n0*c0)))  (model.py:190); first maximum (model.py:182)
        int code = 0;
        float4 cw = make_float4(0.f, 0.f, 0.f, 0.f);
        bool fast = false;
        if (ort_rows > 0) {
          // ---- sign-orthant search (rq_layout.h / pack_codebook_kernel): list picked by the signs of n ----
          fast = (fabsf(n0) >= ort_thr) && (fabsf(n1) >= ort_thr) && (fabsf(n2) >= ort_thr) && (fabsf(n3) >= ort_thr);
          const int sidx = (n0 < 0.f ? 1 : 0) | (n1 < 0.f ? 2 : 0) | (n2 < 0.f ? 4 : 0) | (n3 < 0.f ? 8 : 0);
          const uint32_t lb = cb_smem + sidx * (RQ_ORT_MAX * 16);
          float va = -INFINITY, vb = -INFINITY;
          int ka = 0x7fffffff, kb = 0x7fffffff;
#pragma unroll 2
          for (int k = sub; k < ort_rows; k += 16) {
            const float4 ca = lds128(lb + k * 16), cc = lds128(lb + (k + 8) * 16);
            const float xa = __fmaf_rn(n3, ca.w, __fmaf_rn(n2, ca.z, __fmaf_rn(n1, ca.y, __fmul_rn(n0, ca.x))));
            const float xb = __fmaf_rn(n3, cc.w, __fmaf_rn(n2, cc.z, __fmaf_rn(n1, cc.y, __fmul_rn(n0, cc.x))));
            if (k == sub || xa > va) { va = xa; ka = k; }
            if (k == sub || xb > vb) { vb = xb; kb = k + 8; }
          }
          if (vb > va || (vb == va && kb < ka)) { va = vb; ka = kb; }
#pragma unroll
          for (int s = 4; s >= 1; s >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, va, s);
            const int ok = __shfl_xor_sync(0xffffffffu, ka, s);
            if (ov > va || (ov == va && ok < ka)) { va = ov; ka = ok; }
          }
          ka = __shfl_sync(0xffffffffu, ka, team_lane0);
          code = (int)map_s[sidx * RQ_ORT_MAX + ka];
          cw = lds128(lb + ka * 16);
        }
        if (ort_rows == 0 || __any_sync(0xffffffffu, !fast)) {
          // ---- full scan: generic tables, and tokens with a tiny / zero / NaN coordinate ----
          float va = -INFINITY, vb = -INFINITY;
          int ka = 0x7fffffff, kb = 0x7fffffff;
          if (cb_in_smem) {
#pragma unroll 2
            for (int k = sub; k < n_rows; k += 16) {   // n_rows is a multiple of 32 in shared mode
              const float4 ca = lds128(cb_smem + k * 16), cc = lds128(cb_smem + (k + 8) * 16);
              const float xa = __fmaf_rn(n3, ca.w, __fmaf_rn(n2, ca.z, __fmaf_rn(n1, ca.y, __fmul_rn(n0, ca.x))));
              const float xb = __fmaf_rn(n3, cc.w, __fmaf_rn(n2, cc.z, __fmaf_rn(n1, cc.y, __fmul_rn(n0, cc.x))));
              if (k == sub || xa > va) { va = xa; ka = k; }
              if (k == sub || xb > vb) { vb = xb; kb = k + 8; }
            }
          } else {
            for (int k = sub; k < n_rows; k += 8) {
              const float4 ca = __ldg(cb_l + k);
              const float xa = __fmaf_rn(n3, ca.w, __fmaf_rn(n2, ca.z, __fmaf_rn(n1, ca.y, __fmul_rn(n0, ca.x))));
              if (k == sub || xa > va) { va = xa; ka = k; }
            }
          }
          if (vb > va || (vb == va && kb < ka)) { va = vb; ka = kb; }
#pragma unroll
          for (int s = 4; s >= 1; s >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, va, s);
            const int ok = __shfl_xor_sync(0xffffffffu, ka, s);
            if (ov > va || (ov == va && ok < ka)) { va = ov; ka = ok; }
          }
          // a NaN row (z == 0, inf or NaN input) compares false everywhere: the team's lane 0 still holds
          // row 0, which is what torch.argmax returns for an all-NaN row
          ka = __shfl_sync(0xffffffffu, ka, team_lane0);
          if (!fast) {
            code = p.cb_shared ? (int)map_full[ka] : ka;
            cw = cb_in_smem ? lds128(cb_smem + ka * 16) : __ldg(cb_l + ka);
          }
        }
        if (DBG) {  // parity-test instantiation only: export z, let given codes drive the recurrence
          if (p.z_out != nullptr && sub == 0 && tok_valid)
            reinterpret_cast<float4*>(p.z_out)[token * p.nq_run + l] = make_float4(z0, z1, z2, z3);
          if (p.teacher != nullptr) {
            const int tc = tok_valid ? p.teacher[token * p.nq_run + l] : 0;
            cw = reinterpret_cast<const float4*>(p.codebook)[(p.cb_shared ? 0 : (size_t)l * p.K) + tc];
          }
        }
        // straight-through value c' = z + (c - z)  (model.py:218-220)
        if (sub < 4 && tok_live) {
          const float zc = sub == 0 ? z0 : sub == 1 ? z1 : sub == 2 ? z2 : z3;
          const float cc = sub == 0 ? cw.x : sub == 1 ? cw.y : sub == 2 ? cw.z : cw.w;
          sts32(cpr_t + sub * 8, __fadd_rn(zc, __fsub_rn(cc, zc)));
        }
        if (sub == 0 && tok_live)
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(codes_s + (l & (kCodeBuf - 1)) * 2), "h"((unsigned short)code) : "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&c_ready[g]);
        // ---- flush buffered codes: 16 consecutive layers of one token = one 128-byte store ----
        if (p.codes != nullptr && ((l & (kCodeBuf - 1)) == kCodeBuf - 1 || l == p.nq_run - 1)) {
          const int l0 = l & ~(kCodeBuf - 1);
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const int li = sub + 8 * r;
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
    // drain: the last stages of the last unit are fetched while the quantizer has nothing left to wait for
    if (is_producer) {
      while (it_f < my_iters) top_up();
    }
  }
}

}  // namespace rq
