// Thin inline-PTX wrappers for the sm_100a features the MLP kernel uses:
// mbarrier, 1-D bulk async copy (TMA engine, SASS UBLKCP), tcgen05 MMA / TMEM.
// Descriptor bit layouts follow the PTX ISA tcgen05 "shared memory descriptor"
// and "instruction descriptor" tables.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace gcb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
// Blocking wait on the phase with the given parity.  With GCB_BOUNDED_WAIT a
// wait that exceeds ~1-2 s of SM clocks traps: a protocol bug then surfaces as
// a launch failure instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr(bar);
  uint32_t done = 0;
#ifdef GCB_BOUNDED_WAIT
  const long long t0 = clock64();
#endif
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x2710;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
#ifdef GCB_BOUNDED_WAIT
    if (clock64() - t0 > 3000000000ll) asm volatile("trap;");
#endif
  }
}

// Busy-polling wait (mbarrier.test_wait, no suspend): lowest wake-up latency; for the
// single issuing lanes of the TMA / MMA warps, which have nothing else to do.
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr(bar);
  uint32_t done = 0;
#ifdef GCB_BOUNDED_WAIT
  const long long t0 = clock64();
#endif
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
#ifdef GCB_BOUNDED_WAIT
    if (clock64() - t0 > 3000000000ll) asm volatile("trap;");
#endif
  }
}

// ---- proxies / fences ---------------------------------------------------------
// Make generic-proxy shared-memory writes (st.shared) visible to the async
// proxy (tcgen05.mma operand reads, bulk copies).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Same for global memory: generic-proxy st.global made visible to later bulk copies
// (cp.async.bulk reads through the async proxy) that are ordered after this thread.
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map) ------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

// ---- L2 cache policies -----------------------------------------------------------------
// evict_last: lines of the chain kernel's scratch ring.  They are rewritten in place every few
// units; with the default policy the GBs streaming through the L2 in between evict them and every
// scratch write ends up in DRAM (ncu: 10.7 GB written by a launch whose only HBM output is 3.3 GB).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_global_v4_hint(void* ptr, const uint4& v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
               : "memory");
}
// bulk copy global -> shared, multicast, with an L2 cache policy for the source lines
__device__ __forceinline__ void bulk_g2s_multicast_hint(void* smem_dst, const void* gmem_src,
                                                        uint32_t bytes, uint64_t* bar,
                                                        uint16_t cta_mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4, %5;"
      ::"r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "h"(cta_mask), "l"(policy)
      : "memory");
}

// Hint: bring [gmem_src, +bytes) into L2 (no shared-memory destination, no completion).
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// Same, multicast to every CTA of the cluster selected by cta_mask: the bytes land at
// the same CTA-relative offset in each destination CTA and complete_tx is signalled on
// the mbarrier at the same CTA-relative offset there.
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src,
                                                   uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1], %2, [%3], %4;"
      ::"r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "h"(cta_mask)
      : "memory");
}

// ---- thread-block cluster -----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Distributed shared memory: address of the same CTA-relative location in CTA `rank`.
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t raddr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(raddr), "f"(a), "f"(b) : "memory");
}
// Asynchronous 8-byte store into another CTA's shared memory that signals that CTA's
// mbarrier (complete_tx, 8 bytes) when the data is visible there.  No release fence in the
// sender: a per-thread `mbarrier.arrive.release.cluster` drains the thread's outstanding
// global stores first (ERRBAR + MEMBAR: 11 % of the LayerNorm epilogue's time in ncu).
__device__ __forceinline__ void st_async_f32x2(uint32_t raddr, float a, float b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "f"(a), "f"(b), "r"(rbar) : "memory");
}
// Arrive with release at cluster scope on an mbarrier of THIS CTA (pairs with a
// mbar_wait_cluster by a thread that then acts on behalf of the whole cluster).
__device__ __forceinline__ void mbar_arrive_release_cluster(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// Arrive (release at cluster scope) on an mbarrier of another CTA of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// Wait with acquire at cluster scope (pairs with mbar_arrive_remote).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr(bar);
  uint32_t done = 0;
#ifdef GCB_BOUNDED_WAIT
  const long long t0 = clock64();
#endif
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, 0x2710;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
#ifdef GCB_BOUNDED_WAIT
    if (clock64() - t0 > 3000000000ll) asm volatile("trap;");
#endif
  }
}

// ---- TMEM ---------------------------------------------------------------------
// Whole-warp, .sync.aligned.  Writes the allocated base address to smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_addr(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets row
// (lane base + i), registers j = columns (col base + j).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- tcgen05.mma ------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, no swizzle ("interleave"):
// the operand is a grid of core matrices, each 8 rows x 16 bytes stored as 128
// contiguous bytes;  SBO = byte distance between core matrices adjacent along
// M/N (next 8 rows),  LBO = byte distance between core matrices adjacent along
// K (next 16 bytes of K).  Fields are in units of 16 bytes.
//   [0,14)  start address >> 4      [16,30) LBO >> 4      [32,46) SBO >> 4
//   [46,48) version = 1 (sm_100)    [61,64) layout type = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> f32, A and B K-major.
//   [4,6) D format: 1 = f32     [7,10) A format: 1 = bf16   [10,13) B format: 1 = bf16
//   [15] A major: 0 = K         [16] B major: 0 = K
//   [17,23) N >> 3              [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier when all previously issued MMAs of this thread are done
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                   "r"(smem_addr(bar))
               : "memory");
}

// Same, arriving on the barrier at the same CTA-relative offset in every CTA of cta_mask.
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_addr(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---- warpgroup register re-allocation -----------------------------------------------
// All four warps of a warpgroup (warps 4k..4k+3) must execute the same setmaxnreg.
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

// ---- misc ---------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// Split four floats into packed bf16 "hi" and "lo" parts: x ~= hi + lo with
// |x - hi - lo| <= 2^-17 |x|.
__device__ __forceinline__ void split_bf16x4(const float4& x, uint2& hi, uint2& lo) {
  __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y);
  __nv_bfloat162 h23 = __floats2bfloat162_rn(x.z, x.w);
  float2 f01 = __bfloat1622float2(h01);
  float2 f23 = __bfloat1622float2(h23);
  __nv_bfloat162 l01 = __floats2bfloat162_rn(x.x - f01.x, x.y - f01.y);
  __nv_bfloat162 l23 = __floats2bfloat162_rn(x.z - f23.x, x.w - f23.y);
  hi.x = *reinterpret_cast<uint32_t*>(&h01);
  hi.y = *reinterpret_cast<uint32_t*>(&h23);
  lo.x = *reinterpret_cast<uint32_t*>(&l01);
  lo.y = *reinterpret_cast<uint32_t*>(&l23);
}

}  // namespace ptx
}  // namespace gcb
