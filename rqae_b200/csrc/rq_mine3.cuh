// rq_mine3_kernel: top / middle / bottom-k selection (scripts/3_make_rqae_features.py:116-128, reference
// harish-kamath/rqae) with ONE streaming pass over a row instead of three histogram passes.
//
// rq_mine2_kernel is exact but instruction-bound (49 thread-instructions per value over its three passes; an
// HBM-bound pass affords about 11).  Here the four window boundaries are BRACKETED from a 1/16 sample of the row
// before the row is streamed:
//   A  sample   one 32-byte sector out of every 16 (the sector inside a group is picked by a hash, so that a
//               period in the token order cannot alias with the sample), histogram of the high 11 key bits, and from
//               its prefix sums four thresholds at bin edges: T_top (at least c_top sample values lie above it), T_bot,
//               and [M_lo, M_hi] around the median ranks (4.5 sigma of the sample's rank error on either side)
//   B  stream   every value is compared against the thresholds with packed fp16 compares (two values per
//               instruction): values above M_hi are COUNTED (G), values >= T_top, <= T_bot or inside [M_lo, M_hi]
//               (about 6 % of a row) are appended, bit pattern and index, to a shared-memory candidate buffer
//   C  select   exact selection among the candidates: top-k / bottom-k by a rank sort of their class, the middle
//               window by a two-level radix select over the candidates of the bracket (ranks m0 - G .. m1 - 1 - G)
// The brackets are VERIFIED, not trusted: G <= m0, G + |bracket| >= m1, at least k values in either tail class, no
// NaN in the row, no buffer overflow.  A row that fails any of it is appended to a fallback list and finished by
// rq_mine2_kernel in list mode (same stream, no host round trip), so the result is the exact selection in the same total
// order (value descending, index ascending; +0 before -0) whatever the data looks like -- heavily tied rows, constant
// rows and rows with NaN simply take the old path.  fp compares treat +0 and -0 as equal; both then fall into the
// same class and are ordered by their exact keys in step C.
#pragma once
#include <cuda_fp16.h>

#include "rq_mine.cuh"

namespace rq {

constexpr int M3_THREADS = 512;
constexpr int M3_WARPS = M3_THREADS / 32;
constexpr int M3_CAP = 12288;      // candidates per row
constexpr int M3_LIST = 1024;      // members of one class that enter a rank sort
constexpr int M3_BINS = 2048;

struct Mine3Smem {
  uint32_t hist[M3_BINS];
  uint32_t candi[M3_CAP];
  unsigned short candk[M3_CAP];
  uint32_t lk[M3_LIST];
  uint32_t li[M3_LIST];
  uint32_t h2[2][32];
  uint32_t scan[M3_WARPS];
  uint32_t ccount, G, nan, lcount;
  uint32_t pt, pb, ps, pe;           // bin positions (descending-value order) of the four thresholds
  uint32_t binA, exA, binB, exB;     // level-1 bins of the two middle-window ranks and their exclusive prefixes
  uint32_t klo, khi, less_lo;
};

__device__ __forceinline__ uint32_t m3_hash(uint32_t g, uint32_t row) {
  uint32_t h = g * 0x9E3779B1u + row * 0x85EBCA6Bu + 0x165667B1u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}

// slot for every lane with pred set: one atomic per warp
__device__ __forceinline__ uint32_t m3_warp_slot(bool pred, uint32_t* counter, int lane) {
  const uint32_t mask = __ballot_sync(0xffffffffu, pred);
  if (mask == 0) return 0;
  uint32_t base = 0;
  const int leader = __ffs(mask) - 1;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ uint32_t m3_block_exclusive(uint32_t v, uint32_t* scan, int warp, int lane, uint32_t* total) {
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) scan[warp] = x;
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < M3_WARPS; w++) {
    const uint32_t s = scan[w];
    before += w < warp ? s : 0u;
    all += s;
  }
  __syncthreads();
  if (total) *total = all;
  return before + x - v;
}

struct M3Thr { __half stop, sbot, smhi, smlo; };
__device__ __forceinline__ bool m3_in_class(int cls, __half h, const M3Thr& t) {
  return cls == 0 ? __hge(h, t.stop) : (cls == 2 ? __hle(h, t.sbot) : (__hge(h, t.smlo) && !__hgt(h, t.smhi)));
}

// One window out of one class of the candidate buffer (0: values >= T_top, window = its first k; 1: the bracket of the
// median, window = ranks m0 - G .. m1 - 1 - G; 2: values <= T_bot, window = its last k), by a two-level radix select
// over the class's key range [kbase, kend] and a rank sort of the keys that remain.  Whole block; returns false
// (uniformly) when the class does not hold the window.
__device__ bool m3_select(Mine3Smem& sm, const int cls, const M3Thr thr, const uint32_t ncand, const uint32_t kbase,
                          const uint32_t kend, const uint32_t G, const long long m0, const long long m1, const int k,
                          const long long row, const MineParams& p) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int shift = 0;
  while (((kend - kbase) >> shift) >= (uint32_t)M3_BINS) shift++;
  const uint32_t lowmask = (1u << shift) - 1u;
  for (int i = tid; i < M3_BINS; i += M3_THREADS) sm.hist[i] = 0;
  if (tid < 64) (&sm.h2[0][0])[tid] = 0;
  if (tid == 0) sm.lcount = 0;
  __syncthreads();
  for (uint32_t i = tid; i < ncand; i += M3_THREADS) {
    const uint32_t bits = sm.candk[i];
    if (m3_in_class(cls, __ushort_as_half((unsigned short)bits), thr)) atomicAdd(&sm.hist[(mn_dkey(bits) - kbase) >> shift], 1u);
  }
  __syncthreads();
  uint32_t c[4], tot = 0, total = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) { c[i] = sm.hist[tid * 4 + i]; tot += c[i]; }
  uint32_t ex = m3_block_exclusive(tot, sm.scan, warp, lane, &total);
  uint32_t r_lo, r_hi;
  if (cls == 1) {
    if ((long long)G > m0 || (long long)G + (long long)total < m1) return false;
    r_lo = (uint32_t)(m0 - (long long)G); r_hi = (uint32_t)(m1 - 1 - (long long)G);
  } else {
    if (total < (uint32_t)k) return false;
    r_lo = cls == 0 ? 0u : total - (uint32_t)k; r_hi = r_lo + (uint32_t)k - 1u;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t in = ex + c[i];
    if (ex <= r_lo && r_lo < in) { sm.binA = (uint32_t)tid * 4 + i; sm.exA = ex; }
    if (ex <= r_hi && r_hi < in) { sm.binB = (uint32_t)tid * 4 + i; sm.exB = ex; }
    ex = in;
  }
  __syncthreads();
  const uint32_t binA = sm.binA, binB = sm.binB;
  if (shift > 0) {   // level 2: the low key bits inside the two boundary bins
    for (uint32_t i = tid; i < ncand; i += M3_THREADS) {
      const uint32_t bits = sm.candk[i];
      if (m3_in_class(cls, __ushort_as_half((unsigned short)bits), thr)) {
        const uint32_t d = mn_dkey(bits) - kbase;
        if ((d >> shift) == binA) atomicAdd(&sm.h2[0][d & lowmask], 1u);
        if ((d >> shift) == binB) atomicAdd(&sm.h2[1][d & lowmask], 1u);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    uint32_t run = sm.exA, klo = kbase + (binA << shift);
    for (uint32_t j = 0; shift > 0 && j <= lowmask; j++) {
      if (r_lo < run + sm.h2[0][j]) { klo += j; break; }
      run += sm.h2[0][j];
    }
    sm.klo = klo; sm.less_lo = run;
    uint32_t run2 = sm.exB, khi = kbase + (binB << shift);
    for (uint32_t j = 0; shift > 0 && j <= lowmask; j++) {
      if (r_hi < run2 + sm.h2[1][j]) { khi += j; break; }
      run2 += sm.h2[1][j];
    }
    sm.khi = khi;
  }
  __syncthreads();
  const uint32_t klo = sm.klo, khi = sm.khi;
  for (uint32_t i0 = 0; i0 < ncand; i0 += M3_THREADS) {
    const uint32_t i = i0 + tid;
    bool take = false;
    uint32_t d = 0;
    if (i < ncand) {
      const uint32_t bits = sm.candk[i];
      d = mn_dkey(bits);
      take = m3_in_class(cls, __ushort_as_half((unsigned short)bits), thr) && d >= klo && d <= khi;
    }
    const uint32_t s = m3_warp_slot(take, &sm.lcount, lane);
    if (take && s < M3_LIST) { sm.lk[s] = d; sm.li[s] = sm.candi[i]; }
  }
  __syncthreads();
  const int cnt = (int)sm.lcount;
  if (cnt > M3_LIST) return false;
  // rank sort: (key, index) ascending = value descending, index ascending
  const int first = (int)(r_lo - sm.less_lo), size = (int)(r_hi - r_lo + 1u);
  int* io = p.idx_out + (row * 3 + cls) * (long long)k;
  __half* vo = p.val_out ? p.val_out + (row * 3 + cls) * (long long)k : nullptr;
  for (int t = tid; t < cnt; t += M3_THREADS) {
    const uint32_t dk = sm.lk[t], di = sm.li[t];
    int r = 0;
    for (int o = 0; o < cnt; o++) {
      const uint32_t ok = sm.lk[o], oi = sm.li[o];
      r += (ok < dk) || (ok == dk && oi < di);
    }
    r -= first;
    if (r >= 0 && r < size) {
      io[r] = (int)di;
      if (vo) vo[r] = __ushort_as_half((unsigned short)mn_bits(dk));
    }
  }
  if (tid >= size && tid < k) { io[tid] = -1; if (vo) vo[tid] = __ushort_as_half((unsigned short)0); }
  __syncthreads();
  return true;
}

__global__ void __launch_bounds__(M3_THREADS, 2) rq_mine3_kernel(const MineParams p) {
  extern __shared__ __align__(16) unsigned char m3_raw[];
  Mine3Smem& sm = *reinterpret_cast<Mine3Smem*>(m3_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n = p.n;
  const int k = p.k, kh = k / 2;
  const long long m0 = n / 2 - kh, m1 = n / 2 + kh;
  const long long nv = n / 8;                                // whole 16-byte vectors
  const int slog = p.sample_log2;
  const long long ngroups = (n / 16) >> slog;                // sample: one sector (16 values) per group
  const uint32_t ms = (uint32_t)ngroups * 16u;

  for (long long row = blockIdx.x; row < p.rows; row += gridDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(p.vals + row * p.row_stride);
    const unsigned short* src16 = reinterpret_cast<const unsigned short*>(p.vals + row * p.row_stride);
    for (int i = tid; i < M3_BINS; i += M3_THREADS) sm.hist[i] = 0;
    if (tid == 0) {
      sm.ccount = 0; sm.G = 0; sm.nan = 0;
      sm.pt = 0; sm.pb = M3_BINS - 1; sm.ps = 0; sm.pe = M3_BINS - 1;
    }
    __syncthreads();

    // ---- A: histogram of the sample, thresholds at bin edges ----
    for (long long g = tid; g < ngroups; g += M3_THREADS) {
      const long long sector = (g << slog) + (m3_hash((uint32_t)g, (uint32_t)row) >> (32 - slog));
      const uint4 a = __ldg(src + sector * 2), b = __ldg(src + sector * 2 + 1);
      const uint32_t w8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const uint32_t w = w8[e];
        const uint32_t d2 = w ^ ((((w >> 15) & 0x00010001u) * 0x7FFFu) ^ 0x7FFF7FFFu);   // packed descending keys
        atomicAdd(&sm.hist[(d2 & 0xFFFFu) >> 5], 1u);
        atomicAdd(&sm.hist[d2 >> 21], 1u);
      }
    }
    __syncthreads();
    {
      uint32_t c[4], tot = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { c[i] = sm.hist[tid * 4 + i]; tot += c[i]; }
      uint32_t ex = m3_block_exclusive(tot, sm.scan, warp, lane, nullptr);
      const uint32_t ct = (uint32_t)p.c_top;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t in = ex + c[i], pos = (uint32_t)tid * 4 + i;
        if (c[i] != 0) {
          if (ex < ct && ct <= in) sm.pt = pos;
          if (ms - in < ct && ct <= ms - ex) sm.pb = pos;
          if (p.c_hi >= 0 && ex <= (uint32_t)p.c_hi && (uint32_t)p.c_hi < in) sm.ps = pos;
          if ((uint32_t)p.c_lo <= ms && ex < (uint32_t)p.c_lo && (uint32_t)p.c_lo <= in) sm.pe = pos;
        }
        ex = in;
      }
    }
    __syncthreads();
    // bit patterns: lowest value of the top bin, highest of the bottom bin, highest of the first / lowest of the last
    // bracket bin
    const uint32_t Ttop = mn_bits((sm.pt << 5) | 31u), Tbot = mn_bits(sm.pb << 5);
    const uint32_t Mhi = mn_bits(sm.ps << 5), Mlo = mn_bits((sm.pe << 5) | 31u);
    const uint32_t ttop2 = Ttop * 0x10001u, tbot2 = Tbot * 0x10001u, mhi2 = Mhi * 0x10001u, mlo2 = Mlo * 0x10001u;
    const __half2 htop = *reinterpret_cast<const __half2*>(&ttop2), hbot = *reinterpret_cast<const __half2*>(&tbot2);
    const __half2 hmhi = *reinterpret_cast<const __half2*>(&mhi2), hmlo = *reinterpret_cast<const __half2*>(&mlo2);
    const __half stop = __ushort_as_half((unsigned short)Ttop), sbot = __ushort_as_half((unsigned short)Tbot);
    const __half smhi = __ushort_as_half((unsigned short)Mhi), smlo = __ushort_as_half((unsigned short)Mlo);

    // ---- B: one pass over the row ----
    {
      uint32_t gcount = 0, nanacc = 0;
      for (long long v0 = 0; v0 < nv; v0 += 4LL * M3_THREADS) {      // warp-uniform trip count
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const long long v = v0 + (long long)u * M3_THREADS + tid;
          q[u] = make_uint4(0u, 0u, 0u, 0u);
          if (v < nv) q[u] = __ldg(src + v);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const long long v = v0 + (long long)u * M3_THREADS + tid;
          const uint32_t w4[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
          uint32_t cm[4], gm[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const __half2 h = *reinterpret_cast<const __half2*>(&w4[e]);
            gm[e] = __hgt2_mask(h, hmhi);
            cm[e] = __hge2_mask(h, htop) | __hle2_mask(h, hbot) | (__hge2_mask(h, hmlo) & ~gm[e]);
            nanacc |= __hneu2_mask(h, h);
          }
          uint32_t m8 = 0;
          if (v < nv) {
            gcount += __popc((gm[0] & 0x00010001u) | (gm[1] & 0x00020002u) | (gm[2] & 0x00040004u) | (gm[3] & 0x00080008u));
            m8 = (cm[0] & 0x00020001u) | (cm[1] & 0x00080004u) | (cm[2] & 0x00200010u) | (cm[3] & 0x00800040u);
            m8 = (m8 | (m8 >> 16)) & 0xFFu;
          }   // lanes beyond the row hold zeros: no NaN, no candidate
          const uint32_t cnt = __popc(m8);
          uint32_t x = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
          }
          const uint32_t total = __shfl_sync(0xffffffffu, x, 31);
          if (total == 0) continue;
          uint32_t base = 0;
          if (lane == 31) base = atomicAdd(&sm.ccount, total);
          uint32_t pos = __shfl_sync(0xffffffffu, base, 31) + x - cnt;
          while (m8) {
            const int e = __ffs(m8) - 1;
            m8 &= m8 - 1;
            const uint32_t lohi = e < 4 ? (e < 2 ? w4[0] : w4[1]) : (e < 6 ? w4[2] : w4[3]);
            if (pos < M3_CAP) {
              sm.candk[pos] = (unsigned short)((e & 1) ? (lohi >> 16) : (lohi & 0xFFFFu));
              sm.candi[pos] = (uint32_t)(v * 8 + e);
            }
            pos++;
          }
        }
      }
      // the last n % 8 values
      if (tid == 0) {
        for (long long i = nv * 8; i < n; i++) {
          const unsigned short bits = src16[i];
          const __half h = __ushort_as_half(bits);
          if (__hisnan(h)) nanacc = 1;
          if (__hgt(h, smhi)) gcount++;
          if (__hge(h, stop) || __hle(h, sbot) || (__hge(h, smlo) && !__hgt(h, smhi))) {
            const uint32_t pos = atomicAdd(&sm.ccount, 1u);
            if (pos < M3_CAP) { sm.candk[pos] = bits; sm.candi[pos] = (uint32_t)i; }
          }
        }
      }
      // block totals
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gcount += __shfl_xor_sync(0xffffffffu, gcount, o);
        nanacc |= __shfl_xor_sync(0xffffffffu, nanacc, o);
      }
      if (lane == 0) {
        atomicAdd(&sm.G, gcount);
        if (nanacc) atomicOr(&sm.nan, 1u);
      }
    }
    __syncthreads();

    // ---- C: exact selection among the candidates ----
    const uint32_t ncand = sm.ccount;
    bool ok = sm.nan == 0 && ncand <= (uint32_t)M3_CAP;
    M3Thr thr;
    thr.stop = stop; thr.sbot = sbot; thr.smhi = smhi; thr.smlo = smlo;
    // key ranges of the classes (+0 / -0 compare equal: a range that ends at one of them is widened over both)
    const uint32_t top_end = Ttop == 0x0000u ? 0x8000u : mn_dkey(Ttop);
    const uint32_t bot_base = Tbot == 0x8000u ? 0x7FFFu : mn_dkey(Tbot);
    const uint32_t mid_base = Mhi == 0x8000u ? 0x7FFFu : mn_dkey(Mhi);
    const uint32_t mid_end = Mlo == 0x0000u ? 0x8000u : mn_dkey(Mlo);
    ok = ok && m3_select(sm, 0, thr, ncand, 0u, top_end, sm.G, m0, m1, k, row, p);
    ok = ok && (kh == 0 || m3_select(sm, 1, thr, ncand, mid_base, mid_end, sm.G, m0, m1, k, row, p));
    ok = ok && m3_select(sm, 2, thr, ncand, bot_base, 0xFFFFu, sm.G, m0, m1, k, row, p);
    if (ok && kh == 0 && tid < k) {
      p.idx_out[(row * 3 + 1) * (long long)k + tid] = -1;
      if (p.val_out) p.val_out[(row * 3 + 1) * (long long)k + tid] = __ushort_as_half((unsigned short)0);
    }
    if (!ok && tid == 0) {
      const int slot = atomicAdd(p.fb_count, 1);
      p.fb_list[slot] = (int)row;
    }
    __syncthreads();
  }
}

}  // namespace rq
