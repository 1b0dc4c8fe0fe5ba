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
                                                              const int* __restrict__ gate, long long src_ld) {
  __shared__ float tile[32][33];
  if (gate != nullptr && *gate == 0) return;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (c0 < C) {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      tile[i][tx] = (r < R && c < C) ? __ldg(src + static_cast<long long>(r) * src_ld + c) * scale : 0.f;
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
                           const int* gate, long long src_ld) {
  if (R <= 0 || C <= 0) return 0;
  dim3 grid((R + 31) / 32, (C + 31) / 32 + (ones_row ? 1 : 0));
  transpose_split_kernel<<<grid, 256, 0, s>>>(src, R, C, scale, dst_hi, dst_lo, ldr, ones_row, C_pad, dst_lo2, gate,
                                              src_ld > 0 ? src_ld : C);
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

// Matryoshka prefixes on the dense path (saev modeling.py:364-406, objectives.py:124-138).  The decoder contraction of
// prefix block c = dictionary columns [cut_{c-1}, cut_c) runs on the tensor cores from the first column that is a
// multiple of 8 (TMA reads 16-byte aligned spans along the contraction dimension) into y[c][b][:]; the <= 7 columns in
// front of it are added here in fp32 from the bf16 pieces of f and the fp32 dictionary rows, as is b_dec.  So
// x_hat_i = b_dec + sum_{c <= i} (head_c + y_c) and r_i = x_hat_i - x.  One warp per row writes resid = r_{P-1} (AuxK /
// logging use the full prefix), row_sse[b] = sum_i ||r_i||^2, the suffix sums sfx[b][c][:] = sum_{i >= c} r_i (what the
// columns of block c see in the backward pass; same layout as the sparse path's, so gb_dec and x_hats come from the
// same kernels) and, when training, G_c = grad_scale * sfx_c as bf16 pieces g[c][b][:] (operand of the per-block dh
// contraction).  `tensor_mask` bit c: y_c was written (the block reaches past its aligned start).
__global__ void __launch_bounds__(256) dense_prefix_resid_kernel(const float* __restrict__ y, const float* __restrict__ x, int B,
                                                                 int D, PrefixCuts pf, unsigned int tensor_mask,
                                                                 const __nv_bfloat16* __restrict__ f_hi,
                                                                 const __nv_bfloat16* __restrict__ f_lo,
                                                                 const __nv_bfloat16* __restrict__ f_lo2, long long ldf,
                                                                 const float* __restrict__ W_dec,
                                                                 const float* __restrict__ b_dec, float grad_scale,
                                                                 float* __restrict__ resid, float* __restrict__ sfx,
                                                                 float* __restrict__ row_sse,
                                                                 __nv_bfloat16* __restrict__ g_hi,
                                                                 __nv_bfloat16* __restrict__ g_lo,
                                                                 __nv_bfloat16* __restrict__ g_lo2) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const int P = pf.n;
  const long long BD = static_cast<long long>(B) * D, o = static_cast<long long>(b) * D;
  float sse = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float xv = __ldg(x + o + d);
    float r[MAX_PREFIXES];
    float run = __ldg(b_dec + d);
#pragma unroll 1
    for (int c = 0; c < P; ++c) {
      const int k0 = c ? pf.cut[c - 1] : 0, k1 = pf.cut[c];
      const int ka = min(k1, (k0 + 7) & ~7);
      for (int k = k0; k < ka; ++k) {
        const long long fi = static_cast<long long>(b) * ldf + k;
        float fv = __bfloat162float(f_hi[fi]) + __bfloat162float(f_lo[fi]);
        if (f_lo2 != nullptr) fv += __bfloat162float(f_lo2[fi]);
        if (fv != 0.f) run = fmaf(fv, __ldg(W_dec + static_cast<long long>(k) * D + d), run);
      }
      if ((tensor_mask >> c) & 1u) run += __ldg(y + c * BD + o + d);
      r[c] = run - xv;
      sse = fmaf(r[c], r[c], sse);
    }
    resid[o + d] = r[P - 1];
    float suf = 0.f;
#pragma unroll 1
    for (int c = P - 1; c >= 0; --c) {
      suf += r[c];
      sfx[(static_cast<long long>(b) * P + c) * D + d] = suf;
      if (g_hi != nullptr) {
        __nv_bfloat16 h, l;
        split1(suf * grad_scale, h, l);
        g_hi[c * BD + o + d] = h;
        g_lo[c * BD + o + d] = l;
        if (g_lo2 != nullptr) g_lo2[c * BD + o + d] = split_third(suf * grad_scale, h, l);
      }
    }
  }
  sse = warp_sum(sse);
  if (lane == 0) row_sse[b] = sse;
}

int launch_dense_prefix_resid(const float* y, const float* x, int B, int D, const PrefixCuts& pf, unsigned int tensor_mask,
                              const __nv_bfloat16* f_hi, const __nv_bfloat16* f_lo, const __nv_bfloat16* f_lo2,
                              long long ldf, const float* W_dec, const float* b_dec, float grad_scale, float* resid,
                              float* sfx, float* row_sse, __nv_bfloat16* g_hi, __nv_bfloat16* g_lo, __nv_bfloat16* g_lo2,
                              cudaStream_t s) {
  if (pf.n < 1 || pf.n > MAX_PREFIXES) return 24;
  dense_prefix_resid_kernel<<<(B + 7) / 8, 256, 0, s>>>(y, x, B, D, pf, tensor_mask, f_hi, f_lo, f_lo2, ldf, W_dec, b_dec,
                                                        grad_scale, resid, sfx, row_sse, g_hi, g_lo, g_lo2);
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

// ------------------------------------------------------------------------------------------------
// Dictionary coherence  max_{i<j} |<w_i, w_j>| / (||w_i|| ||w_j||)   (saev train.py:415-421, the log block).
// The reference forms the whole [S, S] Gram matrix in fp32 (17 GB at S = 65536, + abs / triu copies); here the
// upper triangle is screened on the tensor cores as a two-piece bf16 split product (epilogue 5 of encode_gemm.cu:
// per row the largest |.| right of the diagonal and its column), and the few pairs within `slack` of the screen
// maximum are recomputed exactly below.
// ------------------------------------------------------------------------------------------------
// hi/lo[j, :] = split(W[j, :] / ||W[j, :]||)     (one warp per row; any D)
__global__ void __launch_bounds__(256) unit_rows_split_kernel(const float* __restrict__ W, int rows, int D,
                                                              __nv_bfloat16* __restrict__ hi,
                                                              __nv_bfloat16* __restrict__ lo) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= rows) return;
  const float* w = W + static_cast<long long>(j) * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) ss = fmaf(__ldg(w + d), __ldg(w + d), ss);
  const float nrm = sqrtf(warp_sum(ss));
  for (int d = lane; d < D; d += 32) {
    __nv_bfloat16 h, l;
    split1(__ldg(w + d) / nrm, h, l);
    hi[static_cast<long long>(j) * D + d] = h;
    lo[static_cast<long long>(j) * D + d] = l;
  }
}

// One block.  out[0] = coherence (exact fp32 inputs, fp64 accumulation over the candidate pairs), out[1] = screen
// maximum, out[2], out[3] = the pair (as floats).  row_best / row_col: [n_entries] screen results (col < 0: none).
__global__ void __launch_bounds__(1024) coherence_finish_kernel(const float* __restrict__ W, int D,
                                                                const float* __restrict__ row_best,
                                                                const int* __restrict__ row_col, int n_entries,
                                                                int nsplit, float slack, float* __restrict__ out) {
  __shared__ float s_max[32];
  __shared__ double s_val[32];
  __shared__ int s_i[32], s_j[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -1.f;
  for (int e = threadIdx.x; e < n_entries; e += blockDim.x)
    if (row_col[e] >= 0) m = fmaxf(m, row_best[e]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if (lane == 0) s_max[warp] = m;
  __syncthreads();
  float gmax = s_max[0];
  for (int w = 1; w < 32; ++w) gmax = fmaxf(gmax, s_max[w]);
  double best = -1.0;
  int bi = -1, bj = -1;
  for (int e = warp; e < n_entries; e += 32) {  // warp-uniform
    const int j = row_col[e];
    if (j < 0 || row_best[e] < gmax - slack) continue;
    const int i = e / nsplit;
    const float* wi = W + static_cast<long long>(i) * D;
    const float* wj = W + static_cast<long long>(j) * D;
    double dot = 0.0, ni = 0.0, nj = 0.0;
    for (int d = lane; d < D; d += 32) {
      const double a = wi[d], b = wj[d];
      dot += a * b;
      ni += a * a;
      nj += b * b;
    }
    for (int o = 16; o; o >>= 1) {
      dot += __shfl_xor_sync(FULL, dot, o);
      ni += __shfl_xor_sync(FULL, ni, o);
      nj += __shfl_xor_sync(FULL, nj, o);
    }
    const double c = fabs(dot) / (sqrt(ni) * sqrt(nj));
    if (c > best) {
      best = c;
      bi = i;
      bj = j;
    }
  }
  if (lane == 0) {
    s_val[warp] = best;
    s_i[warp] = bi;
    s_j[warp] = bj;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w)
      if (s_val[w] > best) {
        best = s_val[w];
        bi = s_i[w];
        bj = s_j[w];
      }
    out[0] = bi >= 0 ? static_cast<float>(best) : 0.f;  // fewer than two rows: triu(1) is empty
    out[1] = gmax;
    out[2] = static_cast<float>(bi);
    out[3] = static_cast<float>(bj);
  }
}

// ------------------------------------------------------------------------------------------------
// The other metrics of the log block (train.py:380-423) in one pass over x / residual and one over W_dec.
// acc (double): [0] sum x^2  [1] sum r^2  [2] sum r  [3] sum x  [4] sum_j ||w_j||  [5] atoms that did not fire
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) log_sums_kernel(const float* __restrict__ x, const float* __restrict__ r, int B,
                                                       int D, double* __restrict__ acc, double* __restrict__ colsum) {
  __shared__ double red[4][4];
  const int col = blockIdx.x * 128 + threadIdx.x;
  double cs = 0.0, sx2 = 0.0, sr2 = 0.0, sr = 0.0;
  if (col < D) {
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
      const double xv = __ldg(x + static_cast<long long>(b) * D + col);
      const double rv = __ldg(r + static_cast<long long>(b) * D + col);
      cs += xv;
      sx2 += xv * xv;
      sr2 += rv * rv;
      sr += rv;
    }
    atomicAdd(colsum + col, cs);
  }
  double v[4] = {sx2, sr2, sr, cs};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], o);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) atomicAdd(acc + threadIdx.x, red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3]);
}

__global__ void __launch_bounds__(256) log_rows_kernel(const float* __restrict__ W, int S, int D,
                                                       const int* __restrict__ fired, double* __restrict__ acc) {
  __shared__ double red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double nsum = 0.0, idle = 0.0;
  for (int j = blockIdx.x * 8 + warp; j < S; j += gridDim.x * 8) {
    const float* w = W + static_cast<long long>(j) * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) ss = fmaf(__ldg(w + d), __ldg(w + d), ss);
    ss = warp_sum(ss);
    nsum += sqrtf(ss);
    if (fired != nullptr && fired[j] == 0) idle += 1.0;
  }
  if (lane == 0) {
    red[0][warp] = nsum;
    red[1][warp] = idle;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(acc + 4 + threadIdx.x, t);
  }
}

// out (double[8]): explained_variance, dead_unit_pct, dictionary_coherence, avg_decoder_row_norm, sse_sae,
// sse_baseline, normalized_mse, coherence screen maximum
__global__ void log_finish_kernel(const double* __restrict__ acc, const double* __restrict__ colsum, int B, int D, int S,
                                  const float* __restrict__ coh, double* __restrict__ out) {
  __shared__ double red[32];
  double t = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) t += colsum[d] * colsum[d];
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double cc = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) cc += red[w];
    const double n = static_cast<double>(B) * D;
    const double sse_base = acc[0] - cc / B;                       // train.py:384-391
    const double var_r = (acc[1] - acc[2] * acc[2] / n) / (n - 1);  // torch .var(): unbiased, over all elements
    const double var_x = (acc[0] - acc[3] * acc[3] / n) / (n - 1);
    out[0] = 1.0 - var_r / var_x;
    out[1] = acc[5] / S;
    out[2] = coh[0];
    out[3] = acc[4] / S;
    out[4] = acc[1];
    out[5] = sse_base;
    out[6] = acc[1] / sse_base;
    out[7] = coh[1];
  }
}

int launch_log_metrics(const float* x, const float* r, int B, int D, const float* W, int S, const int* fired,
                       double* acc /* [8 + D], zeroed here */, const float* coh, double* out, cudaStream_t s) {
  if (cudaMemsetAsync(acc, 0, (8 + static_cast<size_t>(D)) * sizeof(double), s) != cudaSuccess) return 22;
  dim3 grid((D + 127) / 128, 148);
  log_sums_kernel<<<grid, 128, 0, s>>>(x, r, B, D, acc, acc + 8);
  log_rows_kernel<<<148 * 4, 256, 0, s>>>(W, S, D, fired, acc);
  log_finish_kernel<<<1, 256, 0, s>>>(acc, acc + 8, B, D, S, coh, out);
  g_launch_count += 3;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// evaluate() accumulators (saev train.py:546-566): per-atom firing counts and value sums from the sparse forward
// state, the batch sums of the normalised-MSE baseline, and the batch-weighted loss scalars.
// ------------------------------------------------------------------------------------------------
// n_fired[j] += #(f[b, j] > 0), values[j] += sum_b f[b, j] over the B*K (index, value) slots (index < 0: empty)
__global__ void __launch_bounds__(256) feature_stats_topk_kernel(const int* __restrict__ idx, const float* __restrict__ val,
                                                                 long long n, float* __restrict__ n_fired,
                                                                 float* __restrict__ values) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int j = __ldg(idx + i);
    if (j < 0) continue;
    const float v = __ldg(val + i);
    if (v > 0.f) atomicAdd(n_fired + j, 1.f);
    if (v != 0.f) atomicAdd(values + j, v);
  }
}

// same from the dense ReLU activations kept as bf16 pieces [B, ld]
__global__ void __launch_bounds__(128) feature_stats_dense_kernel(const __nv_bfloat16* __restrict__ hi,
                                                                  const __nv_bfloat16* __restrict__ lo,
                                                                  const __nv_bfloat16* __restrict__ lo2, int B, int S,
                                                                  long long ld, float* __restrict__ n_fired,
                                                                  float* __restrict__ values) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= S) return;
  float cnt = 0.f, sum = 0.f;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const long long o = static_cast<long long>(b) * ld + j;
    const float f = __bfloat162float(hi[o]) + (__bfloat162float(lo[o]) + (lo2 ? __bfloat162float(lo2[o]) : 0.f));
    cnt += (f > 0.f) ? 1.f : 0.f;
    sum += f;
  }
  if (cnt > 0.f) atomicAdd(n_fired + j, cnt);
  if (sum != 0.f) atomicAdd(values + j, sum);
}

// acc[4..7] += {l0, l1, mse} * B, B        (train.py:564-566: batch-weighted means, fp64)
__global__ void eval_scalars_kernel(const float* __restrict__ losses, int B, double* __restrict__ acc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    acc[4] += static_cast<double>(losses[3]) * B;
    acc[5] += static_cast<double>(losses[4]) * B;
    acc[6] += static_cast<double>(losses[0]) * B;
    acc[7] += B;
  }
}

int launch_eval_accumulate(const float* x, const float* r, int B, int D, const float* losses, double* acc,
                           cudaStream_t s) {
  dim3 grid((D + 127) / 128, 148);
  log_sums_kernel<<<grid, 128, 0, s>>>(x, r, B, D, acc, acc + 8);
  eval_scalars_kernel<<<1, 32, 0, s>>>(losses, B, acc);
  g_launch_count += 2;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

int launch_feature_stats_topk(const int* idx, const float* val, long long n, float* n_fired, float* values,
                              cudaStream_t s) {
  feature_stats_topk_kernel<<<148 * 4, 256, 0, s>>>(idx, val, n, n_fired, values);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

int launch_feature_stats_dense(const __nv_bfloat16* hi, const __nv_bfloat16* lo, const __nv_bfloat16* lo2, int B, int S,
                               long long ld, float* n_fired, float* values, cudaStream_t s) {
  dim3 grid((S + 127) / 128, 8);
  feature_stats_dense_kernel<<<grid, 128, 0, s>>>(hi, lo, lo2, B, S, ld, n_fired, values);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

int launch_unit_rows_split(const float* W, int rows, int D, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s) {
  unit_rows_split_kernel<<<(rows + 7) / 8, 256, 0, s>>>(W, rows, D, hi, lo);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

int launch_coherence_finish(const float* W, int D, const float* row_best, const int* row_col, int n_entries, int nsplit,
                            float slack, float* out, cudaStream_t s) {
  coherence_finish_kernel<<<1, 1024, 0, s>>>(W, D, row_best, row_col, n_entries, nsplit, slack, out);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

}  // namespace sb
