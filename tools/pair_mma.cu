// ROUND-2 EXPERIMENT (compiled here, NOT yet run on hardware): tcgen05.mma.cta_group::2.
//
// the operand ring of the fused layer is latency-bound; the structural
// fix is a CTA pair (cta_group::2): M = 256 rows per MMA (128 per CTA), each CTA stages only
// HALF of the weight block (128 of the 256 columns), which cuts the bytes per stage from 24.8 KB
// to 16.6 KB (9 stages in the same shared memory) and halves the weight traffic per SM.
//
// This program answers the two open questions before the kernel is rewritten:
//   1. semantics - with the no-swizzle K-major layout, is "A: each CTA its own 128 rows at the
//      same shared-memory offset; B: each CTA 128 of the 256 columns, LBO = 128*16" the right
//      operand placement?  -> a small exact integer GEMM (bf16 operands, values in [-3, 3]) is
//      checked against the host, separately for the rows held by each CTA.
//   2. speed - cycles per K-step of the TMA -> MMA ring for 6 and 9 stages (3 MMAs per K-step
//      as in the bf16x3 mode), to compare with tools/pipe_rate.cu (cta_group::1: 411 / 389).
//
// Protocol: cluster of 2.  Both CTAs run a TMA warp that fills their own stage (A rows + B half)
// and signals their own `full` barrier; the peer's relay lane forwards its `full` to the leader
// (remote arrive, release.cluster); the leader's elected lane waits for both, issues the MMAs and
// commits with multicast to both CTAs' `empty` barriers.  Waits are bounded (-DGCB_BOUNDED_WAIT):
// a wrong protocol traps instead of hanging the GPU.
//
//   nvcc -std=c++17 -O3 -DGCB_BOUNDED_WAIT -gencode arch=compute_100a,code=sm_100a \
//        -o tools/pair_mma tools/pair_mma.cu && timeout 30 ./tools/pair_mma
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../graphcast_b200/csrc/ptx.cuh"

using namespace gcb;

constexpr int kRowsPerCta = 128;
constexpr int kN = 256;                 // MMA N (columns of the pair's accumulator)
constexpr int kNHalf = kN / 2;          // columns of B staged by each CTA
constexpr int kABytes = kRowsPerCta * 16 * 2 + 0;        // one K-step (16 bf16) of 128 rows: 4096
constexpr int kBBytes = kNHalf * 16 * 2;                  // 128 columns x 16 k: 4096
constexpr int kALbo = kRowsPerCta * 16;                   // K chunk (8 elements) stride of A
constexpr int kBLbo = kNHalf * 16;                        // K chunk stride of the B half

// ---- cta_group::2 flavours of the tcgen05 helpers ----------------------------------------------
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   ptx::smem_addr(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(ptx::smem_addr(bar)), "h"(mask) : "memory");
}

// Operand images in global memory, per K-step: [chunk c (2)][row r][8 bf16]  (row = A row or B column)
__device__ __forceinline__ size_t img_off(int kstep, int rows, int row, int k) {
  return (static_cast<size_t>(kstep) * 2 + ((k >> 3) & 1)) * rows * 8 + static_cast<size_t>(row) * 8 + (k & 7);
}

template <int kStages, int kMmasPerStep>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256)
pair_kernel(const __nv_bfloat16* a_img, const __nv_bfloat16* b_img, int ksteps, int repeats,
            float* d_out, long long* cycles) {
  // a_img: [pair][cta][kstep][2][128][8]; b_img: [kstep][half][2][128][8]
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[kStages], peer_full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_slot;
  constexpr int kStage = kABytes + kBBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&peer_full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);          // one multicast commit from the leader
    }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_mbar_init();
  }
  // Open question: does one warp of EACH CTA issue the cta_group::2 allocation (default here),
  // or only a warp of the leader (-DALLOC_LEADER_ONLY)?  Run under `timeout`: a wrong choice can
  // block inside tcgen05.alloc, which the bounded mbarrier waits do not cover.
#ifdef ALLOC_LEADER_ONLY
  if (warp == 2 && leader) { tmem_alloc2(&tmem_slot, 256); tmem_relinquish2(); }
  if (!leader && threadIdx.x == 0) tmem_slot = 0;          // full-TMEM allocation starts at column 0
#else
  if (warp == 2) { tmem_alloc2(&tmem_slot, 256); tmem_relinquish2(); }
#endif
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int total = ksteps * repeats;
  const size_t pair = blockIdx.x / 2;
  const __nv_bfloat16* my_a = a_img + (pair * 2 + rank) * static_cast<size_t>(ksteps) * 2 * kRowsPerCta * 8;

  if (warp == 0) {                                   // TMA warp (both CTAs)
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < total; ++it) {
      const int ks = it % ksteps;
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* dst = smem + stage * kStage;
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&full_bar[stage], kStage);
        ptx::bulk_g2s(dst, my_a + static_cast<size_t>(ks) * 2 * kRowsPerCta * 8, kABytes, &full_bar[stage]);
        ptx::bulk_g2s(dst + kABytes, b_img + (static_cast<size_t>(ks) * 2 + rank) * 2 * kNHalf * 8, kBBytes,
                      &full_bar[stage]);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && !leader) {                 // relay: my stage is full -> tell the leader
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < total; ++it) {
      ptx::mbar_wait(&full_bar[stage], phase);
      if (ptx::elect_one())
        ptx::mbar_arrive_remote(ptx::mapa(ptx::smem_addr(&peer_full_bar[stage]), 0));
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && leader) {                  // MMA warp of the pair
    const uint32_t idesc = ptx::make_idesc_bf16(2 * kRowsPerCta, kN);
    uint32_t stage = 0, phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < total; ++it) {
      ptx::mbar_wait(&full_bar[stage], phase);
      ptx::mbar_wait_cluster(&peer_full_bar[stage], phase);
      ptx::tc_fence_after_sync();
      const uint32_t sa = ptx::smem_addr(smem + stage * kStage), sb = sa + kABytes;
      const uint64_t a_desc = ptx::make_smem_desc(sa, kALbo, 128);
      const uint64_t b_desc = ptx::make_smem_desc(sb, kBLbo, 128);
      if (ptx::elect_one()) {
        for (int m = 0; m < kMmasPerStep; ++m)
          mma2_bf16_ss(tmem, a_desc, b_desc, idesc, (it % ksteps > 0 || m > 0) ? 1u : 0u);
        mma2_commit_multicast(&empty_bar[stage], 0b11);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (ptx::elect_one()) mma2_commit_multicast(&done_bar, 0b11);
    __syncwarp();
    ptx::mbar_wait(&done_bar, 0);
    if (lane == 0) cycles[pair] = clock64() - t0;
  }
  if (warp >= 4) {                                   // read the accumulator back (both CTAs)
    ptx::mbar_wait(&done_bar, 0);
    ptx::tc_fence_after_sync();
    const int ew = warp - 4;
    if (d_out != nullptr && pair == 0) {
      for (int c0 = 0; c0 < kN; c0 += 32) {
        float v[32];
        ptx::tmem_ld32(tmem + (static_cast<uint32_t>(ew * 32) << 16) + c0, v);
        const int row = static_cast<int>(rank) * kRowsPerCta + ew * 32 + lane;
        for (int j = 0; j < 32; ++j) d_out[static_cast<size_t>(row) * kN + c0 + j] = v[j];
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();
#ifdef ALLOC_LEADER_ONLY
  if (warp == 2 && leader) { ptx::tc_fence_after_sync(); tmem_dealloc2(tmem, 256); }
#else
  if (warp == 2) { ptx::tc_fence_after_sync(); tmem_dealloc2(tmem, 256); }
#endif
}

template <int kStages, int kMmas>
double run(const __nv_bfloat16* a, const __nv_bfloat16* b, int ksteps, int repeats, float* d_out,
           long long* d_cycles, int pairs) {
  const int smem = kStages * (kABytes + kBBytes);
  cudaFuncSetAttribute(pair_kernel<kStages, kMmas>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pair_kernel<kStages, kMmas><<<2 * pairs, 256, smem>>>(a, b, ksteps, repeats, d_out, d_cycles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(pairs);
  cudaMemcpy(h.data(), d_cycles, pairs * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long c : h) mx = c > mx ? c : mx;
  return double(mx) / (double(ksteps) * repeats);
}

int main() {
  const int pairs = 74, ksteps = 32;
  // exact integer operands: A [pairs][256 rows][512], B [256 cols][512]
  std::vector<float> A(static_cast<size_t>(256) * 512), B(static_cast<size_t>(256) * 512);
  srand(7);
  for (auto& x : A) x = float(rand() % 7 - 3);
  for (auto& x : B) x = float(rand() % 7 - 3);
  std::vector<__nv_bfloat16> a_img(static_cast<size_t>(pairs) * 2 * ksteps * 2 * 128 * 8),
      b_img(static_cast<size_t>(ksteps) * 2 * 2 * 128 * 8);
  for (int p = 0; p < pairs; ++p)
    for (int cta = 0; cta < 2; ++cta)
      for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 512; ++k) {
          const size_t base = (static_cast<size_t>(p) * 2 + cta) * ksteps * 2 * 128 * 8;
          a_img[base + (static_cast<size_t>(k >> 4) * 2 + ((k >> 3) & 1)) * 128 * 8 + r * 8 + (k & 7)] =
              __float2bfloat16(A[static_cast<size_t>(cta * 128 + r) * 512 + k]);
        }
  for (int n = 0; n < 256; ++n)
    for (int k = 0; k < 512; ++k) {
      const int half = n >> 7, r = n & 127;
      b_img[((static_cast<size_t>(k >> 4) * 2 + half) * 2 + ((k >> 3) & 1)) * 128 * 8 + r * 8 + (k & 7)] =
          __float2bfloat16(B[static_cast<size_t>(n) * 512 + k]);
    }
  __nv_bfloat16 *d_a, *d_b;
  float* d_out;
  long long* d_cycles;
  cudaMalloc(&d_a, a_img.size() * 2); cudaMalloc(&d_b, b_img.size() * 2);
  cudaMalloc(&d_out, 256 * 256 * 4); cudaMalloc(&d_cycles, pairs * 8);
  cudaMemcpy(d_a, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice);

  // 1. semantics: one pass over K = 512 with one MMA per K-step, checked exactly
  run<6, 1>(d_a, d_b, ksteps, 1, d_out, d_cycles, pairs);
  std::vector<float> D(256 * 256);
  cudaMemcpy(D.data(), d_out, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad[2] = {0, 0};
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < 256; ++n) {
      float ref = 0.f;
      for (int k = 0; k < 512; ++k) ref += A[static_cast<size_t>(m) * 512 + k] * B[static_cast<size_t>(n) * 512 + k];
      if (D[static_cast<size_t>(m) * 256 + n] != ref) ++bad[m >> 7];
    }
  printf("cta_group::2 GEMM 256x256x512: mismatches in the leader's rows %d, in the peer's rows %d (of 32768 each)\n",
         bad[0], bad[1]);
  // 2. speed: 3 MMAs per K-step, 64 passes
  printf("cycles per K-step, 3 MMAs (M=256,N=256,K=16) per step: 6 stages %.1f, 9 stages %.1f"
         "  (cta_group::1 ring, tools/pipe_rate.cu: 411 / 389; ideal 384)\n",
         run<6, 3>(d_a, d_b, ksteps, 64, nullptr, d_cycles, pairs),
         run<9, 3>(d_a, d_b, ksteps, 64, nullptr, d_cycles, pairs));
  return bad[0] + bad[1] ? 2 : 0;
}
