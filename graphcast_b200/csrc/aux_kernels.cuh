// HBM-bound helper kernels of the GraphCast step: receiver-sorted segment sum and
// the channel pack / unpack transposes (with the normalisation affine fused in).
#pragma once
#include <stdint.h>

namespace gcb {

// out[node, :] = sum over in-edges (CSR row_ptr, edges receiver-sorted) of msg[e, :].
// One warp per receiver node; lane l owns float4 columns l, l+32, ... so every
// edge row is read with fully coalesced 512-byte warp loads and each output row
// is written once -- no atomics, deterministic order.
template <int kVecPerLane>
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ msg, int ld_msg, const int* __restrict__ row_ptr,
                   int num_nodes, float* __restrict__ out, int ld_out) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long gwarp = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * warps_per_block;
  for (long long node = gwarp; node < num_nodes; node += nwarps) {
    const int beg = row_ptr[node], end = row_ptr[node + 1];
    float4 acc[kVecPerLane];
#pragma unroll
    for (int j = 0; j < kVecPerLane; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    int e = beg;
    for (; e + 1 < end; e += 2) {
      float4 a[kVecPerLane], b[kVecPerLane];
      const float4* pa = reinterpret_cast<const float4*>(msg + static_cast<long long>(e) * ld_msg);
      const float4* pb = reinterpret_cast<const float4*>(msg + static_cast<long long>(e + 1) * ld_msg);
#pragma unroll
      for (int j = 0; j < kVecPerLane; ++j) { a[j] = __ldg(pa + lane + 32 * j); b[j] = __ldg(pb + lane + 32 * j); }
#pragma unroll
      for (int j = 0; j < kVecPerLane; ++j) {
        acc[j].x += a[j].x; acc[j].y += a[j].y; acc[j].z += a[j].z; acc[j].w += a[j].w;
        acc[j].x += b[j].x; acc[j].y += b[j].y; acc[j].z += b[j].z; acc[j].w += b[j].w;
      }
    }
    if (e < end) {
      const float4* pa = reinterpret_cast<const float4*>(msg + static_cast<long long>(e) * ld_msg);
#pragma unroll
      for (int j = 0; j < kVecPerLane; ++j) {
        const float4 a = __ldg(pa + lane + 32 * j);
        acc[j].x += a.x; acc[j].y += a.y; acc[j].z += a.z; acc[j].w += a.w;
      }
    }
    float4* po = reinterpret_cast<float4*>(out + node * ld_out);
#pragma unroll
    for (int j = 0; j < kVecPerLane; ++j) po[lane + 32 * j] = acc[j];
  }
}

// planes [n_ch, n_nodes] (+ node_static [n_nodes, n_static]) -> feats [n_nodes, ld],
// feats[i, c] = (planes[c, i] - mean[c]) / scale[c]; 32x32 smem-tiled transpose so
// both the plane reads (along nodes) and the feature writes (along channels) are
// coalesced.
__global__ void __launch_bounds__(256)
pack_grid_features_kernel(const float* __restrict__ planes, int n_ch, long long n_nodes,
                          const float* __restrict__ mean, const float* __restrict__ scale,
                          const float* __restrict__ node_static, int n_static,
                          float* __restrict__ feats, int ld) {
  __shared__ float tile[32][33];
  const long long node0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const long long node = node0 + tx;
    float v = 0.f;
    if (node < n_nodes) {
      if (c < n_ch) {
        v = planes[static_cast<long long>(c) * n_nodes + node];
        if (mean) v -= mean[c];
        if (scale) v /= scale[c];
      } else if (c < n_ch + n_static) {
        v = node_static[node * n_static + (c - n_ch)];
      }
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long long node = node0 + r;
    const int c = c0 + tx;
    if (node < n_nodes && c < ld) feats[node * ld + c] = tile[tx][r];
  }
}

// y [n_nodes, ld_y] -> planes_out [n_out, n_nodes] with per-channel affine and an
// optional additive plane (the last input frame for residual targets).
__global__ void __launch_bounds__(256)
unpack_grid_outputs_kernel(const float* __restrict__ y, int ld_y, int n_out, long long n_nodes,
                           const float* __restrict__ scale, const float* __restrict__ offset,
                           const float* __restrict__ add_planes,
                           const int* __restrict__ add_plane_index,
                           float* __restrict__ planes_out) {
  __shared__ float tile[32][33];
  const long long node0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long node = node0 + r;
    const int c = c0 + tx;
    tile[r][tx] = (node < n_nodes && c < n_out) ? y[node * ld_y + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const long long node = node0 + tx;
    if (c < n_out && node < n_nodes) {
      float v = tile[tx][r];
      if (scale) v *= scale[c];
      if (offset) v += offset[c];
      if (add_plane_index) {
        const int ap = add_plane_index[c];
        if (ap >= 0) v += add_planes[static_cast<long long>(ap) * n_nodes + node];
      }
      planes_out[static_cast<long long>(c) * n_nodes + node] = v;
    }
  }
}

}  // namespace gcb
