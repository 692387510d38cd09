// Shared device helpers for the sm_100a RQAE kernels: packed-f32x2 arithmetic (FFMA2/FADD2,
// new on Blackwell), mbarrier / bulk-TMA wrappers, register re-allocation, shared-memory
// loads by 32-bit address.  Everything here is inline PTX for sm_100a; there is no
// fallback path for other architectures.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rq {

typedef unsigned long long u64;

// ---- packed fp32 pairs ---------------------------------------------------------------
// A u64 carries two IEEE fp32 values {lo, hi}; fma.rn.f32x2 / add.rn.f32x2 round each half
// exactly like the scalar fma.rn.f32 / add.rn.f32, so results are bit-identical to scalar
// code while using one issue slot for two lanes of work (SASS: FFMA2 / FADD2).  ptxas folds
// pack2(w, w) into the instruction's scalar-broadcast operand form (Rn.F32).
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- shared memory by 32-bit address ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// four 128-bit loads at addr, addr+STRIDE, ... issued back to back (one asm block: the loads cannot be
// interleaved with their consumers, so their latencies overlap)
template <int STRIDE>
__device__ __forceinline__ void lds128x4(uint32_t addr, float4& a, float4& b, float4& c, float4& d) {
  asm volatile(
      "ld.shared.v4.f32 {%0, %1, %2, %3}, [%16];\n\t"
      "ld.shared.v4.f32 {%4, %5, %6, %7}, [%16+%17];\n\t"
      "ld.shared.v4.f32 {%8, %9, %10, %11}, [%16+%18];\n\t"
      "ld.shared.v4.f32 {%12, %13, %14, %15}, [%16+%19];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w), "=f"(c.x), "=f"(c.y),
        "=f"(c.z), "=f"(c.w), "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w)
      : "r"(addr), "n"(STRIDE), "n"(2 * STRIDE), "n"(3 * STRIDE));
}
__device__ __forceinline__ void lds128_u64(uint32_t addr, u64& a, u64& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// ---- mbarrier --------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) expires, so a
// waiting warp does not burn issue slots of the compute warps that share its scheduler.
#ifndef RQ_WAIT_HINT_NS
#define RQ_WAIT_HINT_NS 100000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(RQ_WAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- thread-block clusters: distributed shared memory and cluster-scope barriers -----------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA's window) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t caddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
// Asynchronous 4-byte store into another CTA's shared memory that, on completion, counts 4 bytes on a barrier of THAT
// CTA (complete_tx): the receiver learns of the data through its own barrier, as with a bulk copy, and no fence is
// needed on either side.  (A generic st.shared::cluster followed by a release.cluster arrive costs the sending warp a
// cluster-scope memory barrier per hand-over: measured 6x slower on the forward kernel's critical path.)
__device__ __forceinline__ void st_async_cluster_f32(uint32_t caddr, float v, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(caddr),
               "r"(__float_as_uint(v)), "r"(cbar)
               : "memory");
}
// arrive on a barrier of another CTA of the cluster; release at cluster scope publishes the stores before it
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
// wait that also acquires what other CTAs of the cluster stored before their (cluster-scope) arrives
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(RQ_WAIT_HINT_NS)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- bulk TMA (1-D): global -> shared, completion counted in bytes on an mbarrier --------
// SASS: UBLKCP.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          dst_smem),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// `fraction` of the addressed lines (chosen by an address hash) are kept with evict_last priority, the rest are
// streamed with evict_first: pins a hot subset of a working set that does not fit the L2 as a whole.
__device__ __forceinline__ uint64_t l2_policy_hot_fraction(float fraction) {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(p) : "f"(fraction));
  return p;
}

// ---- register re-allocation between warp roles (warpgroup-aligned) -----------------------
template <int N>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace rq
