// Fused linear layer on tcgen05 tensor cores (sm_100a).
//
//   out[r] = residual[r] + LN( act( concat_s gather_s(r) @ W + bias ) )
//
// One persistent CTA per SM processes 128-row tiles.  Warp roles:
//   warp 0        weight producer: one lane streams the pre-packed bf16 weight
//                 image of each K-step into shared memory with cp.async.bulk
//                 (TMA engine), completing on the stage's mbarrier.
//   warp 1        MMA issuer: one lane issues tcgen05.mma (M=128, N=256 x n/256,
//                 K=16) per stage -- three products per stage in BF16X3 mode
//                 (hi*hi, hi*lo, lo*hi) -- accumulating fp32 in TMEM; commits
//                 free the stage and finally publish the accumulator.
//   warp 2        TMEM allocator (512 columns = 128 x 512 fp32 accumulator).
//   warps 4-7     epilogue: tcgen05.ld the accumulator (thread = row), bias,
//                 swish | LayerNorm (+ residual), vectorised global stores.
//   warps 8-15    activation producers (two groups of four warps, alternating
//                 K-steps): gather the fp32 rows of every K-segment through the
//                 sender / receiver index (ld.global.v4), split to bf16 hi/lo and
//                 store them into the stage in the UMMA K-major core-matrix
//                 layout; the concatenated [E,1536] edge input is never
//                 materialised.
//
// Shared-memory operand layout (no swizzle, K-major): a [R x 16] bf16 operand of
// one K-step is two "K chunks" of 8 elements; chunk c, row r lives at byte
// c * (R*16) + r*16.  Eight consecutive rows form one 128-byte core matrix,
// so SBO = 128 and LBO = R*16 (see ptx.cuh make_smem_desc).
#pragma once
#include "../../include/graphcast_b200.h"
#include "ptx.cuh"

namespace gcb {

constexpr int kTileM = 128;
constexpr int kKStep = 16;
constexpr int kThreads = 512;
constexpr int kAPartBytes = kTileM * kKStep * 2;  // 4096: one of {hi, lo}
constexpr int kMaxN = 512;
constexpr int kMaxKSteps = 128;                   // K <= 2048
constexpr int kTmemCols = 512;

template <bool kSplit>
struct TcConfig {
  static constexpr int kStages = kSplit ? 5 : 8;
  static constexpr int kAStageBytes = kSplit ? 2 * kAPartBytes : kAPartBytes;
  static constexpr int kBStageBytes = kSplit ? kMaxN * kKStep * 4 : kMaxN * kKStep * 2;
  static constexpr int kStageBytes = kAStageBytes + kBStageBytes;
  static constexpr int kParamBytes = 3 * kMaxN * 4;  // bias, ln scale, ln offset
  static constexpr int kSmemBytes = kStages * kStageBytes + kParamBytes + 1024;
};

__device__ __forceinline__ float swish_f(float x) {
  // x * sigmoid(x); __expf / __frcp_rn keep the error ~1e-7 relative.
  return x * __frcp_rn(1.0f + __expf(-x));
}

struct KStepInfo {
  uint8_t seg;
  uint16_t koff;  // element offset of this K-step inside its segment
};

template <bool kSplit>
__global__ void __launch_bounds__(kThreads, 1)
mlp_layer_tc_kernel(const __grid_constant__ gcb_layer_desc d) {
  using Cfg = TcConfig<kSplit>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage_base = smem;
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  float* s_scale = s_bias + kMaxN;
  float* s_offset = s_scale + kMaxN;
  uint8_t* tail = reinterpret_cast<uint8_t*>(s_offset + kMaxN);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);          // [kStages]
  uint64_t* empty_bar = full_bar + Cfg::kStages;                   // [kStages]
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;              // [1]
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;                    // [1]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);
  KStepInfo* ks_info = reinterpret_cast<KStepInfo*>(tmem_base_slot + 2);  // [kMaxKSteps]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n = d.n;
  const int num_tiles = (d.rows + kTileM - 1) / kTileM;
  int ksteps = 0;
  for (int s = 0; s < d.nseg; ++s) ksteps += d.seg[s].k / kKStep;
  const bool has_ln = d.ln_scale != nullptr;

  // ---- one-time setup ---------------------------------------------------------
  for (int i = threadIdx.x; i < n; i += kThreads) {
    s_bias[i] = d.bias[i];
    s_scale[i] = has_ln ? d.ln_scale[i] : 1.0f;
    s_offset[i] = has_ln ? d.ln_offset[i] : 0.0f;
  }
  if (threadIdx.x == 0) {
    int ks = 0;
    for (int s = 0; s < d.nseg; ++s)
      for (int k = 0; k < d.seg[s].k; k += kKStep) {
        ks_info[ks].seg = static_cast<uint8_t>(s);
        ks_info[ks].koff = static_cast<uint16_t>(k);
        ++ks;
      }
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 5);   // 1 weight producer + 4 activation warps
      ptx::mbar_init(&empty_bar[s], 1);  // tcgen05.commit
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::mbar_init(tmem_empty_bar, 4);   // 4 epilogue warps
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_slot;

  // ---- roles ------------------------------------------------------------------
  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      const uint32_t b_bytes = static_cast<uint32_t>(n) * kKStep * (kSplit ? 4 : 2);
      const size_t b_stride = static_cast<size_t>(n) * kKStep * 4;  // image always holds hi|lo
      const uint8_t* wimg = static_cast<const uint8_t*>(d.w_packed);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const uint32_t stage = it % Cfg::kStages;
          const uint32_t phase = (it / Cfg::kStages) & 1;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], b_bytes);
          ptx::bulk_g2s(stage_base + stage * Cfg::kStageBytes + Cfg::kAStageBytes,
                        wimg + ks * b_stride, b_bytes, &full_bar[stage]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(kTileM, 256);
      const uint32_t n_halves = n / 256;
      const uint32_t b_lbo = static_cast<uint32_t>(n) * 16;
      const uint32_t b_part = static_cast<uint32_t>(n) * kKStep * 2;
      uint32_t it = 0, tile_iter = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
        ptx::mbar_wait(tmem_empty_bar, (tile_iter & 1) ^ 1);
        ptx::tc_fence_after_sync();
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const uint32_t stage = it % Cfg::kStages;
          const uint32_t phase = (it / Cfg::kStages) & 1;
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sa = ptx::smem_addr(stage_base + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kAStageBytes;
          const uint64_t a_hi = ptx::make_smem_desc(sa, kTileM * 16, 128);
          const uint64_t a_lo = ptx::make_smem_desc(sa + kAPartBytes, kTileM * 16, 128);
          for (uint32_t h = 0; h < n_halves; ++h) {
            const uint32_t boff = h * 256 * 16;
            const uint64_t b_hi = ptx::make_smem_desc(sb + boff, b_lbo, 128);
            const uint32_t dcol = tmem_base + h * 256;
            ptx::mma_bf16_ss(dcol, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
            if (kSplit) {
              const uint64_t b_lo = ptx::make_smem_desc(sb + b_part + boff, b_lbo, 128);
              ptx::mma_bf16_ss(dcol, a_hi, b_lo, idesc, 1u);
              ptx::mma_bf16_ss(dcol, a_lo, b_hi, idesc, 1u);
            }
          }
          ptx::mma_commit(&empty_bar[stage]);   // stage reusable once these MMAs retire
        }
        ptx::mma_commit(tmem_full_bar);          // accumulator complete
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue =====
    const int ew = warp - 4;                     // == warp % 4: TMEM lane quarter
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const int n_valid = d.n_valid;
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
      ptx::mbar_wait(tmem_full_bar, tile_iter & 1);
      ptx::tc_fence_after_sync();
      const long long grow = static_cast<long long>(tile) * kTileM + ew * 32 + lane;
      const bool valid = grow < d.rows;
      const uint32_t taddr = tmem_base + lane_base;
      float mean = 0.f, rstd = 1.f;
      if (has_ln) {
        // Pass 1: shifted sums for mean / biased variance over the n_valid columns.
        float shift = 0.f, s1 = 0.f, s2 = 0.f;
        for (int c0 = 0; c0 < n_valid; c0 += 32) {
          float v[32];
          ptx::tmem_ld32(taddr + c0, v);
          if (c0 == 0) shift = v[0] + s_bias[0];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < n_valid) {
              const float x = v[j] + s_bias[c0 + j] - shift;
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
          }
        }
        const float inv_n = 1.0f / static_cast<float>(n_valid);
        const float m1 = s1 * inv_n;
        mean = shift + m1;
        const float var = fmaxf(s2 * inv_n - m1 * m1, 0.f);
        rstd = rsqrtf(var + 1e-5f);
      }
      // Pass 2 (or the only pass): finish and store.
      for (int c0 = 0; c0 < n_valid; c0 += 32) {
        float v[32];
        ptx::tmem_ld32(taddr + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = v[j] + s_bias[c0 + j];
          if (d.act == GCB_ACT_SWISH) x = swish_f(x);
          if (has_ln) x = (x - mean) * rstd * s_scale[c0 + j] + s_offset[c0 + j];
          v[j] = x;
        }
        if (valid) {
          const bool full_chunk = (c0 + 32 <= n_valid);
          if (d.out_y != nullptr) {
            float* p = d.out_y + grow * d.ld_out_y + c0;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              for (int j = 0; j < 32 && c0 + j < n_valid; ++j) p[j] = v[j];
            }
          }
          if (d.out != nullptr) {
            float* p = d.out + grow * d.ld_out + c0;
            const float* rp = d.residual ? d.residual + grow * d.ld_res + c0 : nullptr;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (rp) {
                  const float4 r = *reinterpret_cast<const float4*>(rp + j);
                  o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4*>(p + j) = o;
              }
            } else {
              for (int j = 0; j < 32 && c0 + j < n_valid; ++j) p[j] = v[j] + (rp ? rp[j] : 0.f);
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);
    }
  } else if (warp >= 8) {
    // ===== activation producers =====
    const int group = (warp - 8) >> 2;            // 0 or 1: alternating K-steps
    const int tid_g = threadIdx.x - 256 - group * 128;
    const int sub = tid_g & 3;                    // which float4 of the 16-wide K-step
    const int rg = tid_g >> 2;                    // 0..31; rows rg + 32*i
    const uint32_t sts_off = (sub >> 1) * (kTileM * 16) + (sub & 1) * 8;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      // Source row of each of my 4 tile rows, per segment (-1 = out of range).
      long long src[3][4];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          src[s][i] = -1;
          if (s < d.nseg) {
            const long long grow = static_cast<long long>(tile) * kTileM + rg + 32 * i;
            if (grow < d.rows)
              src[s][i] = d.seg[s].idx ? static_cast<long long>(d.seg[s].idx[grow]) : grow;
          }
        }
      }
      float4 cur[4];
      bool have_cur = false;
      uint32_t cur_it = 0;
      // Software pipeline over the K-steps this group owns: the loads of the next
      // owned K-step are in flight while the current one is converted and stored.
      for (int ks = 0; ks <= ksteps; ++ks) {
        const uint32_t this_it = it + ks;
        const bool mine = (ks < ksteps) && ((this_it & 1u) == static_cast<uint32_t>(group));
        float4 nxt[4];
        if (mine) {
          const int s = ks_info[ks].seg;
          const int koff = ks_info[ks].koff + sub * 4;
          const gcb_segment& sg = d.seg[s];
          const bool kvalid = koff < sg.k_valid;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const long long sr = (s == 0) ? src[0][i] : (s == 1 ? src[1][i] : src[2][i]);
            if (kvalid && sr >= 0) {
              const float* p = sg.table + sr * sg.fan * sg.ld + koff;
              acc = __ldg(reinterpret_cast<const float4*>(p));
              for (int j = 1; j < sg.fan; ++j) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p + static_cast<long long>(j) * sg.ld));
                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
              }
            }
            nxt[i] = acc;
          }
        }
        if (have_cur && (mine || ks == ksteps)) {
          const uint32_t stage = cur_it % Cfg::kStages;
          const uint32_t phase = (cur_it / Cfg::kStages) & 1;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_hi = stage_base + stage * Cfg::kStageBytes;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint2 hi, lo;
            ptx::split_bf16x4(cur[i], hi, lo);
            const uint32_t off = sts_off + (rg + 32 * i) * 16;
            *reinterpret_cast<uint2*>(a_hi + off) = hi;
            if (kSplit) *reinterpret_cast<uint2*>(a_hi + kAPartBytes + off) = lo;
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);
          have_cur = false;
        }
        if (mine) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
          cur_it = this_it;
          have_cur = true;
        }
      }
      it += ksteps;
    }
  }

  // ---- teardown ---------------------------------------------------------------
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace gcb
