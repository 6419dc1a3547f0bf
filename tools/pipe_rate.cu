// Microbenchmark of the TMA -> tcgen05.mma stage pipeline protocol (no epilogue): one TMA
// warp and one MMA warp per CTA, 6 stages of {A 8448 B, B 16384 B}, bf16x3 (3 MMAs per
// K-step), commit -> empty barrier.  Compares ways of writing the two issuing warps:
//   style 0  `if (lane == 0) { whole loop }`                       (divergent, "waterfall" SASS)
//   style 1  converged warp, every lane try_wait, elect_one issues
//   style 2  converged warp, elected lane test_wait-spins and issues, __syncwarp per K-step
//   style 3  converged warp, every lane test_wait-spins, elect_one issues
//   style 4  `if (lane == 0)` loop with try_wait but no suspend-time hint
// nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_rate tools/pipe_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../graphcast_b200/csrc/ptx.cuh"

using namespace gcb;

constexpr int kABytes = 8448, kBBytes = 16384, kStage = kABytes + kBBytes;

__device__ __forceinline__ void wait_try_nohint(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = ptx::smem_addr(bar);
  uint32_t done = 0;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
}

template <int kStyle>
__device__ __forceinline__ void wait_all(uint64_t* bar, uint32_t parity) {
  if (kStyle == 1) ptx::mbar_wait(bar, parity);
  else ptx::mbar_spin(bar, parity);
}

template <int kTmaStyle, int kMmaStyle, int kStages>
__global__ void __launch_bounds__(256) pipe_kernel(int ksteps_total, const uint8_t* ga, const uint8_t* gb,
                                                   long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    ptx::mbar_init(&done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint8_t* my_a = ga + static_cast<size_t>(blockIdx.x) * ksteps_total * kABytes;

  if (warp == 0) {
    if (kTmaStyle == 0 || kTmaStyle == 4) {
      if (lane == 0) {
        for (int it = 0; it < ksteps_total; ++it) {
          const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
          if (kTmaStyle == 0) ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          else wait_try_nohint(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], kStage);
          uint8_t* dst = base + stage * kStage;
          ptx::bulk_g2s(dst, my_a + static_cast<size_t>(it) * kABytes, kABytes, &full_bar[stage]);
          ptx::bulk_g2s(dst + kABytes, gb + static_cast<size_t>(it & 31) * kBBytes, kBBytes, &full_bar[stage]);
        }
      }
    } else {
      for (int it = 0; it < ksteps_total; ++it) {
        const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
        if (kTmaStyle != 2) wait_all<kTmaStyle>(&empty_bar[stage], phase ^ 1);
        uint8_t* dst = base + stage * kStage;
        if (ptx::elect_one()) {
          if (kTmaStyle == 2) ptx::mbar_spin(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], kStage);
          ptx::bulk_g2s(dst, my_a + static_cast<size_t>(it) * kABytes, kABytes, &full_bar[stage]);
          ptx::bulk_g2s(dst + kABytes, gb + static_cast<size_t>(it & 31) * kBBytes, kBBytes, &full_bar[stage]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, 256);
    const long long t0 = clock64();
    if (kMmaStyle == 0 || kMmaStyle == 4) {
      if (lane == 0) {
        for (int it = 0; it < ksteps_total; ++it) {
          const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
          if (kMmaStyle == 0) ptx::mbar_wait(&full_bar[stage], phase);
          else wait_try_nohint(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sa = ptx::smem_addr(base + stage * kStage), sb = sa + kABytes;
          const uint64_t a_hi = ptx::make_smem_desc(sa, 2112, 128), b_hi = ptx::make_smem_desc(sb, 4096, 128);
          const uint64_t a_lo = ptx::make_smem_desc(sa + 4224, 2112, 128), b_lo = ptx::make_smem_desc(sb + 8192, 4096, 128);
          ptx::mma_bf16_ss(tmem, a_hi, b_hi, idesc, (it & 31) > 0 ? 1u : 0u);
          ptx::mma_bf16_ss(tmem, a_hi, b_lo, idesc, 1u);
          ptx::mma_bf16_ss(tmem, a_lo, b_hi, idesc, 1u);
          ptx::mma_commit(&empty_bar[stage]);
        }
        ptx::mma_commit(&done_bar);
        ptx::mbar_wait(&done_bar, 0);
      }
    } else {
      for (int it = 0; it < ksteps_total; ++it) {
        const uint32_t stage = it % kStages, phase = (it / kStages) & 1;
        if (kMmaStyle != 2) wait_all<kMmaStyle>(&full_bar[stage], phase);
        const uint32_t sa = ptx::smem_addr(base + stage * kStage), sb = sa + kABytes;
        const uint64_t a_hi = ptx::make_smem_desc(sa, 2112, 128), b_hi = ptx::make_smem_desc(sb, 4096, 128);
        const uint64_t a_lo = a_hi + (4224 >> 4), b_lo = b_hi + (8192 >> 4);
        if (ptx::elect_one()) {
          if (kMmaStyle == 2) ptx::mbar_spin(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          ptx::mma_bf16_ss(tmem, a_hi, b_hi, idesc, (it & 31) > 0 ? 1u : 0u);
          ptx::mma_bf16_ss(tmem, a_hi, b_lo, idesc, 1u);
          ptx::mma_bf16_ss(tmem, a_lo, b_hi, idesc, 1u);
          ptx::mma_commit(&empty_bar[stage]);
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::mma_commit(&done_bar);
      __syncwarp();
      ptx::mbar_wait(&done_bar, 0);
    }
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 512); }
}

template <int T, int M, int kStages = 6>
void run(const char* name, int ksteps, const uint8_t* ga, const uint8_t* gb, long long* d_out) {
  const int smem_bytes = kStages * kStage + 2048;
  cudaFuncSetAttribute(pipe_kernel<T, M, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int rep = 0; rep < 2; ++rep) {
    pipe_kernel<T, M, kStages><<<148, 256, smem_bytes>>>(ksteps, ga, gb, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, sum = 0;
  for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
  printf("stages %d TMA style %d, MMA style %d  %-28s  max %7.1f  mean %7.1f cycles/K-step\n", kStages, T, M, name,
         double(mx) / ksteps, double(sum) / 148 / ksteps);
}

int main() {
  const int ksteps = 1024;
  uint8_t *ga, *gb;
  long long* d_out;
  cudaMalloc(&ga, static_cast<size_t>(148) * ksteps * kABytes);
  cudaMalloc(&gb, 32 * kBBytes);
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaMemset(ga, 0, static_cast<size_t>(148) * ksteps * kABytes);
  cudaMemset(gb, 0, 32 * kBBytes);
  run<0, 0, 3>("lane0 / lane0", ksteps, ga, gb, d_out);
  run<0, 0, 4>("lane0 / lane0", ksteps, ga, gb, d_out);
  run<0, 0, 6>("lane0 / lane0", ksteps, ga, gb, d_out);
  run<0, 0, 8>("lane0 / lane0", ksteps, ga, gb, d_out);
  run<1, 1, 3>("elect all-try_wait", ksteps, ga, gb, d_out);
  run<1, 1, 4>("elect all-try_wait", ksteps, ga, gb, d_out);
  run<1, 1, 6>("elect all-try_wait", ksteps, ga, gb, d_out);
  run<1, 1, 8>("elect all-try_wait", ksteps, ga, gb, d_out);
  run<0, 1, 8>("lane0 TMA / elect MMA", ksteps, ga, gb, d_out);
  return 0;
}
