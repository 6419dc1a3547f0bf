// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128) as a function of
// the shared-memory operand layout.  One elected thread issues `reps` MMAs back to back on
// fixed shared-memory operands (contents irrelevant), commits, and waits; clock64 brackets it.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../graphcast_b200/csrc/ptx.cuh"

using namespace gcb;

struct Variant {
  const char* name;
  uint32_t n;          // MMA N
  uint32_t layout;     // descriptor bits [61,64): 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_alt, b_alt;   // byte offset of the alternate ("lo") operand; 0 = always the same
  uint32_t nstage, stage_bytes;   // rotate operands through nstage buffers
  uint32_t mode;   // 1: commit per k-step | 2: + (already complete) mbarrier wait + fence | 4: epilogue warps hammer tcgen05.ld | 8: TMA lane streams 24.8 KB stages from global | 16: epilogue warps also write/read smem transposes
};

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__global__ void __launch_bounds__(256) rate_kernel(Variant v, int reps, long long* out, const uint8_t* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, kbar[8], donebar, tbar[4];
  __shared__ volatile int done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 190 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1); ptx::mbar_init(&donebar, 1);
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&kbar[i], 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&tbar[i], 1);
    done = 0;
    ptx::fence_mbar_init();
  }
  if (warp == 0) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && (v.mode & 64)) {
    // Warp-converged issue: every lane runs the loop on uniform values; one elected lane
    // issues the tcgen05 instructions.
    const uint32_t idesc = ptx::make_idesc_bf16(128, v.n);
    const uint32_t s0 = ptx::smem_addr(base);
    const uint32_t mode = v.mode, a_alt = v.a_alt, b_alt = v.b_alt;
    const uint64_t a_hi = desc(s0, v.a_lbo, v.a_sbo, v.layout);
    const uint64_t b_hi = desc(s0 + 16384, v.b_lbo, v.b_sbo, v.layout);
    const uint64_t a_lo = a_hi + (a_alt >> 4), b_lo = b_hi + (b_alt >> 4);
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (mode & 2) {
          if (!(mode & 128) || (r & 1) == 0) {
            ptx::mbar_wait(&donebar, 1);
            if (!(mode & 32)) ptx::tc_fence_after_sync();
          }
        }
        if (ptx::elect_one()) {
          ptx::mma_bf16_ss(tmem, a_hi, b_hi, idesc, r > 0);
          if (a_alt) {
            ptx::mma_bf16_ss(tmem, a_hi, b_lo, idesc, 1u);
            ptx::mma_bf16_ss(tmem, a_lo, b_hi, idesc, 1u);
          }
          if ((mode & 1) && (!(mode & 128) || (r & 1) == 1)) ptx::mma_commit(&kbar[r & 7]);
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::mma_commit(&bar);
      __syncwarp();
      ptx::mbar_wait(&bar, pass & 1);
      const long long t1 = clock64();
      if (pass == 1 && lane == 0) out[blockIdx.x] = t1 - t0;
    }
    done = 1;
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, v.n);
    const uint32_t s0 = ptx::smem_addr(base);
    const uint32_t mode = v.mode, a_alt = v.a_alt, b_alt = v.b_alt;
    const uint64_t a_hi = desc(s0, v.a_lbo, v.a_sbo, v.layout);
    const uint64_t b_hi = desc(s0 + 16384, v.b_lbo, v.b_sbo, v.layout);
    const uint64_t a_lo = a_hi + (a_alt >> 4), b_lo = b_hi + (b_alt >> 4);
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (mode & 2) {
          if (!(mode & 128) || (r & 1) == 0) {
            ptx::mbar_wait(&donebar, 1);
            if (!(mode & 32)) ptx::tc_fence_after_sync();
          }
        }
        ptx::mma_bf16_ss(tmem, a_hi, b_hi, idesc, r > 0);
        if (a_alt) {
          ptx::mma_bf16_ss(tmem, a_hi, b_lo, idesc, 1u);
          ptx::mma_bf16_ss(tmem, a_lo, b_hi, idesc, 1u);
        }
        if ((mode & 1) && (!(mode & 128) || (r & 1) == 1)) ptx::mma_commit(&kbar[r & 7]);
      }
      ptx::mma_commit(&bar);
      ptx::mbar_wait(&bar, pass & 1);
      const long long t1 = clock64();
      if (pass == 1) out[blockIdx.x] = t1 - t0;
    }
    done = 1;
  } else if (warp >= 4 && (v.mode & 4)) {
    const uint32_t lane_base = static_cast<uint32_t>((warp - 4) * 32) << 16;
    float acc = 0.f;
    float* tile = reinterpret_cast<float*>(base + 150 * 1024) + (warp - 4) * 32 * 36;
    while (!done) {
      for (int c0 = 0; c0 < 256; c0 += 32) {
        float x[32];
        ptx::tmem_ld32(tmem + lane_base + 256 + c0, x);
        if (v.mode & 16) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(tile + lane * 36 + q * 4) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 y = *reinterpret_cast<const float4*>(tile + ((lane >> 3) + 4 * i) * 36 + (lane & 7) * 4);
            acc += y.x + y.y + y.z + y.w;
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += x[j];
        }
      }
    }
    if (acc == 123.456f) out[0] = 1;
  } else if (warp == 2 && lane == 0 && (v.mode & 8)) {
    uint32_t it = 0;
    while (!done) {
      const uint32_t st = it & 3;
      if (it >= 4) ptx::mbar_wait(&tbar[st], ((it >> 2) - 1) & 1);
      ptx::mbar_arrive_expect_tx(&tbar[st], 24832);
      ptx::bulk_g2s(base + 49152 + st * 24832, gsrc + (size_t)((blockIdx.x * 64 + (it & 63)) * 24832), 24832, &tbar[st]);
      ++it;
    }
    for (uint32_t k = (it > 4 ? it - 4 : 0); k < it; ++k) ptx::mbar_wait(&tbar[k & 3], (k >> 2) & 1);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  const int smem_bytes = 200 * 1024;
  uint8_t* d_src;
  cudaMalloc(&d_src, (size_t)148 * 64 * 24832);
  cudaMemset(d_src, 1, (size_t)148 * 64 * 24832);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  // stage layout: A at +0 (<= 16 KB), B at +16384 (<= 16 KB)  => 32 KB stage
#define X3(name, mode) {name, 256, 0, 2112, 128, 4096, 128, 4224, 8192, 1, 32768, mode}
  const Variant vs[] = {
      X3("lane0: x3 only", 0),
      X3("lane0: + commit", 1),
      X3("lane0: + wait+fence", 2),
      X3("lane0: + wait (no fence)", 2 | 32),
      X3("lane0: + commit + wait + fence", 3),
      X3("lane0: + commit + wait (no fence)", 3 | 32),
      X3("lane0: commit+wait(no fence) every 2nd k-step", 3 | 32 | 128),
      X3("elect: x3 only", 64),
      X3("elect: + commit", 64 | 1),
      X3("elect: + commit + wait + fence", 64 | 3),
      X3("elect: + commit + wait (no fence)", 64 | 3 | 32),
      X3("elect: commit+wait(no fence) every 2nd k-step", 64 | 3 | 32 | 128),
      X3("elect: all contention, commit+wait nofence", 64 | 3 | 32 | 4 | 8 | 16),
      X3("elect: all contention, every 2nd", 64 | 3 | 32 | 128 | 4 | 8 | 16),
      {"elect x1 bf16: commit + wait (no fence)", 256, 0, 2112, 128, 4096, 128, 0, 0, 1, 32768, 64 | 3 | 32},
      {"elect x1 bf16: commit+wait every 2nd", 256, 0, 2112, 128, 4096, 128, 0, 0, 1, 32768, 64 | 3 | 32 | 128},
  };
  const int reps = 512;
  for (const Variant& v : vs) {
    for (int grid : {148}) {
      rate_kernel<<<grid, 256, smem_bytes>>>(v, reps, d_out, d_src);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const int per = v.a_alt ? 3 : 1;
      printf("%-50s grid=%3d  %7.1f cycles/MMA\n", v.name, grid, double(mx) / (reps * per));
    }
  }
  return 0;
}
