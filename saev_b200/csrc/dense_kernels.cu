// Elementwise / layout kernels of the DENSE (ReLU) SAE path.  saev's ReluActivation (src/saev/nn/modeling.py:150-156)
// keeps every positive pre-activation, so encode, decode and all three weight-gradient contractions are dense
// [B, S] x [S, D] products.  They run on the tcgen05 kernel of encode_gemm.cu as error-compensated bf16 split
// products (hi.hi + hi.lo + lo.hi, ~2^-17 relative), which needs every operand as a bf16 (hi, lo) pair in K-major
// layout; the kernels here produce those operands (splits and transposed splits) and the small fp32 pieces in
// between (residual, MSE partials, gradient projection).
#include "common.cuh"
#include "kernels.h"

namespace sb {

__device__ __forceinline__ void split1(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ __nv_bfloat16 split_third(float v, __nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __float2bfloat16_rn((v - __bfloat162float(hi)) - __bfloat162float(lo));
}

// dst_hi/lo[c, r] = split(scale * src[r, c]) for r < R, c < C (dst row pitch ldr).  When ones_row != 0, row C of dst
// is set to 1.0 (hi) / 0 (lo) over r < R and rows C+1 .. C_pad-1 to zero: the extra "ones" operand row that makes
// a column sum fall out of the following contraction.
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ src, int R, int C, float scale,
                                                              __nv_bfloat16* __restrict__ dst_hi,
                                                              __nv_bfloat16* __restrict__ dst_lo, long long ldr,
                                                              int ones_row, int C_pad,
                                                              __nv_bfloat16* __restrict__ dst_lo2,
                                                              const int* __restrict__ gate) {
  __shared__ float tile[32][33];
  if (gate != nullptr && *gate == 0) return;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (c0 < C) {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      tile[i][tx] = (r < R && c < C) ? __ldg(src + static_cast<long long>(r) * C + c) * scale : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;
      if (c < C && r < R) {
        __nv_bfloat16 h, l;
        split1(tile[tx][i], h, l);
        dst_hi[static_cast<long long>(c) * ldr + r] = h;
        dst_lo[static_cast<long long>(c) * ldr + r] = l;
        if (dst_lo2 != nullptr) dst_lo2[static_cast<long long>(c) * ldr + r] = split_third(tile[tx][i], h, l);
      }
    }
  }
  if (ones_row && blockIdx.y == gridDim.y - 1) {
    for (int c = C + ty; c < C_pad; c += 8) {
      const int r = r0 + tx;
      if (r < R) {
        dst_hi[static_cast<long long>(c) * ldr + r] = __float2bfloat16_rn(c == C ? 1.f : 0.f);
        dst_lo[static_cast<long long>(c) * ldr + r] = __float2bfloat16_rn(0.f);
        if (dst_lo2 != nullptr) dst_lo2[static_cast<long long>(c) * ldr + r] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

int launch_transpose_split(const float* src, int R, int C, float scale, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo,
                           long long ldr, int ones_row, int C_pad, cudaStream_t s, __nv_bfloat16* dst_lo2,
                           const int* gate) {
  if (R <= 0 || C <= 0) return 0;
  dim3 grid((R + 31) / 32, (C + 31) / 32 + (ones_row ? 1 : 0));
  transpose_split_kernel<<<grid, 256, 0, s>>>(src, R, C, scale, dst_hi, dst_lo, ldr, ones_row, C_pad, dst_lo2, gate);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// In place on xhat[B, D]:  r = xhat - x  (saev objectives.py:133-138);  row_sse[b] = sum r^2;
// when g_hi != null also G = grad_scale * r as a bf16 (hi, lo) pair (operand of the backward contractions).
__global__ void __launch_bounds__(256) dense_resid_kernel(float* __restrict__ xhat, const float* __restrict__ x, int B,
                                                          int D, float grad_scale, float* __restrict__ row_sse,
                                                          __nv_bfloat16* __restrict__ g_hi,
                                                          __nv_bfloat16* __restrict__ g_lo,
                                                          __nv_bfloat16* __restrict__ g_lo2) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const long long o = static_cast<long long>(b) * D;
  float sse = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float r = xhat[o + d] - __ldg(x + o + d);
    xhat[o + d] = r;
    sse = fmaf(r, r, sse);
    if (g_hi != nullptr) {
      __nv_bfloat16 h, l;
      split1(r * grad_scale, h, l);
      g_hi[o + d] = h;
      g_lo[o + d] = l;
      if (g_lo2 != nullptr) g_lo2[o + d] = split_third(r * grad_scale, h, l);
    }
  }
  sse = warp_sum(sse);
  if (lane == 0) row_sse[b] = sse;
}

int launch_dense_resid(float* xhat, const float* x, int B, int D, float grad_scale, float* row_sse, __nv_bfloat16* g_hi,
                       __nv_bfloat16* g_lo, cudaStream_t s, __nv_bfloat16* g_lo2) {
  dense_resid_kernel<<<(B + 7) / 8, 256, 0, s>>>(xhat, x, B, D, grad_scale, row_sse, g_hi, g_lo, g_lo2);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// g[j, :] -= (<g_j, w_j> / ||w_j||^2) w_j for every row (saev modeling.py:419-445; rows with ||w||^2 == 0 untouched)
__global__ void __launch_bounds__(256) project_rows_kernel(float* __restrict__ g, const float* __restrict__ w, int rows,
                                                           int D) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= rows) return;
  float* gr = g + static_cast<long long>(j) * D;
  const float* wr = w + static_cast<long long>(j) * D;
  float dot = 0.f, nsq = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float wv = __ldg(wr + d);
    dot = fmaf(gr[d], wv, dot);
    nsq = fmaf(wv, wv, nsq);
  }
  dot = warp_sum(dot);
  nsq = warp_sum(nsq);
  if (!(nsq > 0.f)) return;
  const float sc = dot / nsq;
  for (int d = lane; d < D; d += 32) gr[d] = fmaf(-sc, __ldg(wr + d), gr[d]);
}

int launch_project_rows(float* g, const float* w, int rows, int D, cudaStream_t s) {
  project_rows_kernel<<<(rows + 7) / 8, 256, 0, s>>>(g, w, rows, D);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// f[b, s] = hi + lo   (lazy dense f_x for saev's logging block / evaluate; exact to ~2^-17 relative)
__global__ void join_bf16_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                 const __nv_bfloat16* __restrict__ lo2, long long n, float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __bfloat162float(hi[i]) + (__bfloat162float(lo[i]) + (lo2 ? __bfloat162float(lo2[i]) : 0.f));
}
int launch_join_bf16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long n, float* out, cudaStream_t s,
                     const __nv_bfloat16* lo2) {
  join_bf16_kernel<<<148 * 8, 256, 0, s>>>(hi, lo, lo2, n, out);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

}  // namespace sb
