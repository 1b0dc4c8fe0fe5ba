// HBM/L2-bound kernels of the SAE training step (everything except the tensor-core contraction).
//
// Layout conventions (all fp32, row-major):
//   x[B,D]  W_enc_t[S,D] (= saev's W_enc[D,S] transposed: one dictionary atom per row)  b_enc[S]
//   W_dec[S,D]  b_dec[D]  topk_idx/topk_val/dh[B,K]  resid[B,D]
// Keeping both weight matrices atom-major makes every sparse access a contiguous D-float row:
// decode gathers K rows of W_dec per sample, the weight gradients are per-atom segmented sums of
// rows of resid / x, Adam + decoder renorm is a per-row pass.
//
// Row-per-warp kernels keep a whole D-vector in registers: lane l owns float4 vectors l, l+32, ...,
// VPL of them (VPL = ceil(D/128), D <= 2048).
#include "common.cuh"
#include "kernels.h"

namespace sb {

constexpr float NORM_UP = 1.00001f;  // stored norms scale upper bounds: round them up past the fp32 summation error

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void fma4(float4& acc, float s, float4 v) {
  acc.x = fmaf(s, v.x, acc.x);
  acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z);
  acc.w = fmaf(s, v.w, acc.w);
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 (hi) [+ bf16 residual (lo)]   (operand prep for the tensor-core screen)
// ------------------------------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, __nv_bfloat16* __restrict__ lo2_out, long long n4,
                                  const int* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldg4(src + 4 * i);
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(hi + 4 * i);
    ho[0] = __nv_bfloat162(h0, h1);
    ho[1] = __nv_bfloat162(h2, h3);
    if (lo != nullptr) {
      __nv_bfloat162* lo2 = reinterpret_cast<__nv_bfloat162*>(lo + 4 * i);
      lo2[0] = __nv_bfloat162(__float2bfloat16_rn(v.x - __bfloat162float(h0)),
                              __float2bfloat16_rn(v.y - __bfloat162float(h1)));
      lo2[1] = __nv_bfloat162(__float2bfloat16_rn(v.z - __bfloat162float(h2)),
                              __float2bfloat16_rn(v.w - __bfloat162float(h3)));
      if (lo2_out != nullptr) {  // third piece: what hi + lo still miss
        const float r[4] = {v.x - __bfloat162float(h0), v.y - __bfloat162float(h1), v.z - __bfloat162float(h2),
                            v.w - __bfloat162float(h3)};
        const __nv_bfloat162 l01 = lo2[0], l23 = lo2[1];
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(lo2_out + 4 * i);
        o[0] = __nv_bfloat162(__float2bfloat16_rn(r[0] - __bfloat162float(l01.x)),
                              __float2bfloat16_rn(r[1] - __bfloat162float(l01.y)));
        o[1] = __nv_bfloat162(__float2bfloat16_rn(r[2] - __bfloat162float(l23.x)),
                              __float2bfloat16_rn(r[3] - __bfloat162float(l23.y)));
      }
    }
  }
}

int launch_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n, cudaStream_t s,
                      __nv_bfloat16* lo2, const int* gate) {
  if (n <= 0) return 0;
  if (n % 4) return 21;
  const long long n4 = n / 4;
  const int threads = 256;
  long long want = (n4 + threads - 1) / threads;
  const int blocks = static_cast<int>(want < 148LL * 16 ? want : 148LL * 16);
  split_bf16_kernel<<<blocks, threads, 0, s>>>(src, hi, lo, lo2, n4, gate);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// x[B,D] -> fp16 operand of the top-k screen, one warp per row.  The row is divided by 2^e_b, the power of two just
// above ||x_b||_inf (exact: only the exponent changes), so that every row uses the fp16 range [2^-14, 1) whatever its
// magnitude; row_scale[b] = 2^e_b goes back on in the GEMM epilogue, row_norm[b] = ||x_b||_2 feeds the error bound
// (kernels.h, screen_bound).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_x_kernel(const float* __restrict__ x, int B, int D,
                                                     __half* __restrict__ x16, float* __restrict__ row_norm,
                                                     float* __restrict__ row_dx, float* __restrict__ row_scale,
                                                     const float* __restrict__ scalars,
                                                     unsigned int* __restrict__ tau_keys, float* __restrict__ guess_L,
                                                     int rows_padded) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31, D4 = D >> 2;
  if (b >= B) {
    if (b < rows_padded && lane == 0) tau_keys[b] = 0u;  // rows past the batch: no threshold (they admit nothing anyway)
    return;
  }
  const float* row = x + static_cast<long long>(b) * D;
  __half* orow = x16 + static_cast<long long>(b) * D;
  float mx = 0.f, ss = 0.f;
  for (int v = lane; v < D4; v += 32) {
    const float4 q = ldg4(row + 4 * v);
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w))));
    ss += dot4(q, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
  ss = warp_sum(ss);
  int e = 0;
  if (mx > 0.f && mx <= 3.0e38f) frexpf(mx, &e);  // mx = m 2^e, m in [0.5, 1)
  e = max(-100, min(e, 126));
  const float down = ldexpf(1.f, -e);
  float sd = 0.f;  // ||x16 - x 2^-e||^2, in the scaled units
  for (int v = lane; v < D4; v += 32) {
    const float4 q = ldg4(row + 4 * v);  // L1 hit
    const float4 qs = make_float4(q.x * down, q.y * down, q.z * down, q.w * down);
    const __half2 h01 = __floats2half2_rn(qs.x, qs.y), h23 = __floats2half2_rn(qs.z, qs.w);
    __half2* o = reinterpret_cast<__half2*>(orow + 4 * v);
    o[0] = h01;
    o[1] = h23;
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const float4 dq = make_float4(f01.x - qs.x, f01.y - qs.y, f23.x - qs.z, f23.y - qs.w);
    sd += dot4(dq, dq);
  }
  sd = warp_sum(sd);
  if (lane == 0) {
    const float up = ldexpf(1.f, e);
    const float xn = sqrtf(ss) * NORM_UP;  // (rounded up: they scale an upper bound)
    const float dxn = sqrtf(sd) * up * NORM_UP;
    row_norm[b] = xn;
    row_dx[b] = dxn;
    row_scale[b] = up;
    // threshold guess (kernels.h): L_guess = rho ||x|| max||w|| - max|b|; the kernel admits t_j > tau with tau = L - Q_b
    const float rho = scalars[SC_RHO_GUESS];
    float Lg = -INFINITY;
    unsigned int key = 0u;
    if (rho > -1.5f && rho < 1.5f) {
      Lg = rho * (sqrtf(ss) * sqrtf(scalars[SC_WNORM_SQ_MAX])) - scalars[SC_BIAS_ABS_MAX];
      const ScreenBound sbd = screen_bound(D, scalars[SC_RHO], scalars[SC_BIAS_ABS_MAX]);
      key = fkey(Lg - screen_Q(sbd));
    }
    guess_L[b] = Lg;
    tau_keys[b] = key;
  }
}
int launch_prep_x(const float* x, int B, int D, __half* x16, float* row_norm, float* row_dx, float* row_scale,
                  const float* scalars, unsigned int* tau_keys, float* guess_L, int rows_padded, cudaStream_t s) {
  if (D % 4) return 21;
  prep_x_kernel<<<(rows_padded + 7) / 8, 256, 0, s>>>(x, B, D, x16, row_norm, row_dx, row_scale, scalars, tau_keys,
                                                      guess_L, rows_padded);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// rho for the next screen = the `quantile` of the ratios the previous forward recorded, shrunk by `safety` (towards
// "admit more"); -inf when fewer than `min_rows` rows were recorded.  Clears the histogram.
__global__ void __launch_bounds__(1024) screen_guess_kernel(int* __restrict__ hist, float quantile, float safety,
                                                            int min_rows, float* __restrict__ scalars) {
  constexpr int PER = GUESS_BINS / 1024;  // consecutive bins per thread
  __shared__ int wsum[32];
  __shared__ int s_total;
  __shared__ float s_q, s_med;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int c[PER], mine = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    c[i] = hist[PER * t + i];
    hist[PER * t + i] = 0;
    mine += c[i];
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, w, o);
      if (lane >= o) w += v;
    }
    wsum[lane] = w;
    if (lane == 31) s_total = w;
  }
  __syncthreads();
  const int total = s_total;
  if (t == 0 && total < min_rows) scalars[SC_RHO_GUESS] = -INFINITY;
  if (total < min_rows) return;
  const int want = max(1, static_cast<int>(quantile * total));  // the `want`-th smallest ratio ...
  const int half = max(1, total / 2);                            // ... and the median
  int before = (warp > 0 ? wsum[warp - 1] : 0) + incl - mine;    // rows in lower bins
#pragma unroll
  for (int i = 0; i < PER; ++i) {  // the bins holding them: before < rank <= before + c[i]; lower edges
    if (before < want && want <= before + c[i]) s_q = -1.f + 2.f * (PER * t + i) / GUESS_BINS;
    if (before < half && half <= before + c[i]) s_med = -1.f + 2.f * (PER * t + i) / GUESS_BINS;
    before += c[i];
  }
  __syncthreads();
  if (t == 0) {
    // shrink towards "admit more": relative safety, one bin of resolution, and a share of the distribution's own
    // width (rows of a heavy-tailed batch undercut the previous batch's quantile far more often than Gaussian rows)
    const float q = s_q, width = fmaxf(0.f, s_med - s_q);
    scalars[SC_RHO_GUESS] = q - fmaxf(safety * fabsf(q), 0.25f * width) - 2.f / GUESS_BINS;
  }
}
int launch_screen_guess(int* hist, float quantile, float safety, int min_rows, float* scalars, cudaStream_t s) {
  static_assert(GUESS_BINS % 1024 == 0, "screen_guess_kernel gives every thread GUESS_BINS / 1024 bins");
  screen_guess_kernel<<<1, 1024, 0, s>>>(hist, quantile, safety, min_rows, scalars);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

__global__ void to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n4) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldg4(src + 4 * i);
    __half2* o = reinterpret_cast<__half2*>(dst + 4 * i);
    o[0] = __floats2half2_rn(v.x, v.y);
    o[1] = __floats2half2_rn(v.z, v.w);
  }
}
int launch_to_half(const float* src, __half* dst, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  if (n % 4) return 21;
  const long long n4 = n / 4;
  const long long want = (n4 + 255) / 256;
  to_half_kernel<<<static_cast<int>(want < 148LL * 16 ? want : 148LL * 16), 256, 0, s>>>(src, dst, n4);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

__global__ void __launch_bounds__(256) abs_max_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(v[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
int launch_abs_max(const float* v, int n, float* out, cudaStream_t s) {
  if (cudaMemsetAsync(out, 0, 4, s) != cudaSuccess) return 23;
  abs_max_kernel<<<64, 256, 0, s>>>(v, n, out);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// Per-row inputs of the screen's error bound from three sums of squares of a row and its fp16 rounding:
//   c = max(||w||, ||fp16(w)||) rounded up,   ratio = ||w - fp16(w)|| / c   (0 for a zero row)
__device__ __forceinline__ void screen_col_stats(float ssq, float ssq16, float ssd, float& c, float& ratio) {
  c = sqrtf(fmaxf(ssq, ssq16)) * NORM_UP;
  ratio = c > 0.f ? sqrtf(ssd) * NORM_UP / c : 0.f;
}
__device__ __forceinline__ void accum_fp16_stats(float4 p, float& ssq, float& ssq16, float& ssd) {
  const float2 a = __half22float2(__floats2half2_rn(p.x, p.y)), b = __half22float2(__floats2half2_rn(p.z, p.w));
  const float4 h = make_float4(a.x, a.y, b.x, b.y);
  const float4 d = make_float4(h.x - p.x, h.y - p.y, h.z - p.z, h.w - p.w);
  ssq += dot4(p, p);
  ssq16 += dot4(h, h);
  ssd += dot4(d, d);
}

// *out = max_j ||W[j,:]||^2   (non-negative floats order like their bit patterns: atomicMax on int)
__global__ void __launch_bounds__(256) row_sumsq_max_kernel(const float* __restrict__ W, int rows, int D,
                                                            float* __restrict__ out, float* __restrict__ col_norm,
                                                            float* __restrict__ rho) {
  const int lane = threadIdx.x & 31;
  float best = 0.f, best_ratio = 0.f;
  for (int j = blockIdx.x * 8 + (threadIdx.x >> 5); j < rows; j += gridDim.x * 8) {
    const float* row = W + static_cast<long long>(j) * D;
    float ss = 0.f, ss16 = 0.f, sd = 0.f;
    for (int v = lane; v < (D >> 2); v += 32) accum_fp16_stats(ldg4(row + 4 * v), ss, ss16, sd);
    ss = warp_sum(ss);
    ss16 = warp_sum(ss16);
    sd = warp_sum(sd);
    float c, ratio;
    screen_col_stats(ss, ss16, sd, c, ratio);
    if (col_norm != nullptr && lane == 0) col_norm[j] = c;
    best = fmaxf(best, ss);
    best_ratio = fmaxf(best_ratio, ratio);
  }
  if (lane == 0) {
    atomicMax(reinterpret_cast<int*>(out), __float_as_int(best));
    if (rho != nullptr) atomicMax(reinterpret_cast<int*>(rho), __float_as_int(best_ratio));
  }
}
int launch_row_sumsq_max(const float* W, int rows, int cols, float* out, cudaStream_t s, float* col_norm, float* rho) {
  if (cols % 4) return 21;
  if (cudaMemsetAsync(out, 0, 4, s) != cudaSuccess) return 23;
  if (rho != nullptr && cudaMemsetAsync(rho, 0, 4, s) != cudaSuccess) return 23;
  row_sumsq_max_kernel<<<148 * 4, 256, 0, s>>>(W, rows, cols, out, col_norm, rho);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}
// ------------------------------------------------------------------------------------------------
// W[j,:] /= ||W[j,:]||_2      (saev modeling.py:411-417 normalize_w_dec)
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256) normalize_rows_kernel(float* __restrict__ W, int rows, int D) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= rows) return;
  const int lane = threadIdx.x & 31, D4 = D >> 2;
  float* row = W + static_cast<long long>(j) * D;
  float4 w[VPL];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    w[i] = (v < D4) ? *reinterpret_cast<const float4*>(row + 4 * v) : make_float4(0, 0, 0, 0);
    ss += dot4(w[i], w[i]);
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < D4) {
      float4 o = w[i];
      o.x /= nrm; o.y /= nrm; o.z /= nrm; o.w /= nrm;
      *reinterpret_cast<float4*>(row + 4 * v) = o;
    }
  }
}

int launch_normalize_rows(float* W, int rows, int cols, cudaStream_t s) {
  if (cols % 4) return 21;
  SB_DISPATCH_VPL(cols, (normalize_rows_kernel<VPL><<<(rows + 7) / 8, 256, 0, s>>>(W, rows, cols)));
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// Datapoint initialisation of the dictionary (saev train.py:141-185), one warp per atom j:
//   enc_j   = blend * (acts[src_row[j]] - mean) + (1 - blend) * noise[j]         (:168-171)
//   W_dec[j] = enc_j (when the transpose is tied, :177-178), normalised (:179, cfg.normalize_w_dec)
//   W_enc[:, j] = W_dec[j]                                                        (:181)
// Both matrices are atom-major here, so the last two lines are one row written twice.
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256) datapoint_init_kernel(const float* __restrict__ acts,
                                                             const long long* __restrict__ src_row,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ noise,
                                                             const long long* __restrict__ noise_row, float blend, int tie,
                                                             int normalize, int S, int D, float* __restrict__ W_enc_t,
                                                             float* __restrict__ W_dec) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= S) return;
  const int lane = threadIdx.x & 31, D4 = D >> 2;
  const float* arow = acts + src_row[j] * D;
  const float* nrow = noise + (noise_row != nullptr ? noise_row[j] : static_cast<long long>(j)) * D;
  float* drow = W_dec + static_cast<long long>(j) * D;
  float4 w[VPL];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    w[i] = make_float4(0, 0, 0, 0);
    if (v < D4) {
      if (tie) {
        const float4 a = ldg4(arow + 4 * v), m = ldg4(mean + 4 * v), n = ldg4(nrow + 4 * v);
        // same operation order as the reference: blend * (a - m) + (1 - blend) * n
        w[i].x = blend * (a.x - m.x) + (1.f - blend) * n.x;
        w[i].y = blend * (a.y - m.y) + (1.f - blend) * n.y;
        w[i].z = blend * (a.z - m.z) + (1.f - blend) * n.z;
        w[i].w = blend * (a.w - m.w) + (1.f - blend) * n.w;
      } else {
        w[i] = *reinterpret_cast<const float4*>(drow + 4 * v);  // W_dec keeps its rows; only W_enc follows it
      }
      ss += dot4(w[i], w[i]);
    }
  }
  if (normalize) {
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      w[i].x /= nrm; w[i].y /= nrm; w[i].z /= nrm; w[i].w /= nrm;
    }
  }
  float* erow = W_enc_t + static_cast<long long>(j) * D;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < D4) {
      *reinterpret_cast<float4*>(drow + 4 * v) = w[i];
      *reinterpret_cast<float4*>(erow + 4 * v) = w[i];
    }
  }
}
int launch_datapoint_init(const float* acts, const long long* src_row, const float* mean, const float* noise,
                          const long long* noise_row, float blend, int tie, int normalize, int S, int D, float* W_enc_t,
                          float* W_dec, cudaStream_t s) {
  if (D % 4) return 21;
  SB_DISPATCH_VPL(D, (datapoint_init_kernel<VPL><<<(S + 7) / 8, 256, 0, s>>>(acts, src_row, mean, noise, noise_row, blend, tie,
                                                                             normalize, S, D, W_enc_t, W_dec)));
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// Merge the per-split candidate buffers of a row, recompute the exact fp32 pre-activation of each
// candidate (h = <x_b, W_enc_t[j]> + b_enc[j], saev modeling.py:344-347) and select the top-k of
// those (TopKActivation.forward, modeling.py:169-179: no ReLU, exactly k kept).  One warp per row.
// ------------------------------------------------------------------------------------------------
constexpr int RESCORE_WARPS = 8;
constexpr int RESCORE_CAP = 192;      // most candidates of one row that survive the merged threshold (top_k <= 64)
constexpr int RESCORE_CAP_WIDE = 384; // ... for 64 < top_k <= 128 (the row capacity of the BatchTopK path)

// One radix pass of the warp-wide k-th-largest search over a row's candidate lists: histogram of the 8-bit digit
// at `shift` of every screen key whose higher digits equal `prefix`, then the bin holding the `need`-th largest.
__device__ __forceinline__ void rescore_radix_pass(const int2* cbuf, const int* cnts, int nlists, int stride, int shift,
                                                   unsigned int& prefix, int& need, int* hist, int lane) {
  int4* h4 = reinterpret_cast<int4*>(hist);
  h4[2 * lane] = make_int4(0, 0, 0, 0);
  h4[2 * lane + 1] = make_int4(0, 0, 0, 0);
  __syncwarp();
  for (int l = 0; l < nlists; ++l) {
    const int c = abs(cnts[l]);
    const int2* lb = cbuf + static_cast<long long>(l) * stride;
    for (int e = lane; e < c; e += 32) {
      const unsigned int key = fkey(__int_as_float(__ldg(&lb[e].x)));
      if (shift == 24 || (key >> (shift + 8)) == prefix) atomicAdd(hist + ((key >> shift) & 255u), 1);
    }
  }
  __syncwarp();
  const int4 a = h4[2 * lane], b = h4[2 * lane + 1];
  const int c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const int mine = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  int suf = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(FULL, suf, o);
    if (lane + o < 32) suf += t;
  }
  const unsigned int bal = __ballot_sync(FULL, suf >= need);
  const int L = 31 - __clz(bal);
  int cum = suf - mine, j = 7;
#pragma unroll
  for (int jj = 7; jj > 0; --jj) {
    if (j == jj && cum + c[jj] < need) {
      cum += c[jj];
      j = jj - 1;
    }
  }
  const int bin = __shfl_sync(FULL, 8 * lane + j, L);
  need = __shfl_sync(FULL, need - cum, L);
  prefix = (prefix << 8) | static_cast<unsigned int>(bin);
  __syncwarp();
}

// Same pass over a row's candidates staged in shared memory (value bits in .x): the common case, see the kernel.
__device__ __forceinline__ void rescore_radix_pass_staged(const int2* st, int n, int shift, unsigned int& prefix, int& need,
                                                          int* hist, int lane) {
  int4* h4 = reinterpret_cast<int4*>(hist);
  h4[2 * lane] = make_int4(0, 0, 0, 0);
  h4[2 * lane + 1] = make_int4(0, 0, 0, 0);
  __syncwarp();
  for (int e = lane; e < n; e += 32) {
    const unsigned int key = fkey(__int_as_float(st[e].x));
    if (shift == 24 || (key >> (shift + 8)) == prefix) atomicAdd(hist + ((key >> shift) & 255u), 1);
  }
  __syncwarp();
  const int4 a = h4[2 * lane], b = h4[2 * lane + 1];
  const int c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const int mine = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  int suf = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(FULL, suf, o);
    if (lane + o < 32) suf += t;
  }
  const unsigned int bal = __ballot_sync(FULL, suf >= need);
  const int L = 31 - __clz(bal);
  int cum = suf - mine, j = 7;
#pragma unroll
  for (int jj = 7; jj > 0; --jj) {
    if (j == jj && cum + c[jj] < need) {
      cum += c[jj];
      j = jj - 1;
    }
  }
  const int bin = __shfl_sync(FULL, 8 * lane + j, L);
  need = __shfl_sync(FULL, need - cum, L);
  prefix = (prefix << 8) | static_cast<unsigned int>(bin);
  __syncwarp();
}

constexpr int RESCORE_STAGE = 256;  // candidates of one row staged in shared memory (8 per lane, loaded in one go)

// WPB rows (warps) per block; the rows of a block hold their SM slot until the slowest one (longest candidate list)
// is done, so small blocks keep more warps busy (same 24 resident warps per SM either way)
template <int VPL, int WPB, int CAP = RESCORE_CAP>
__global__ void __launch_bounds__(32 * WPB, 24 / WPB) rescore_topk_kernel(RescoreArgs a) {
  __shared__ int hist_s[WPB][256];
  __shared__ float sv_s[WPB][CAP];
  __shared__ int si_s[WPB][CAP];
  __shared__ float se_s[WPB][CAP];
  __shared__ int2 stage_s[WPB][RESCORE_STAGE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * WPB + warp;
  if (b >= a.B) return;
  const int D4 = a.D >> 2;
  float* sv = sv_s[warp];  // screen values
  int* si = si_s[warp];    // columns
  float* se = se_s[warp];  // exact values
  int* hist = hist_s[warp];

  const int2* cbuf = reinterpret_cast<const int2*>(a.cand) + static_cast<long long>(b) * a.nsplit * a.cand_stride;
  const int* cnts = a.cand_cnt + static_cast<long long>(b) * a.nsplit;
  int n_total = 0;
  bool overflow = false;
  for (int l = 0; l < a.nsplit; ++l) {
    const int c = cnts[l];
    overflow |= c < 0;
    n_total += abs(c);
  }
  // The common case (a row leaves ~100 list entries): pull all of them into shared memory with ONE round of loads
  // (8 independent loads per lane) -- the radix passes and the collection below then never wait on global memory.
  // Entries keep their (list, position) order, so the survivors come out exactly as in the list-walking path.
  int2* stage = stage_s[warp];
  const bool staged = n_total <= RESCORE_STAGE && a.nsplit <= 32;
  if (staged) {
    const int my_cnt = (lane < a.nsplit) ? abs(cnts[lane]) : 0;
    int my_off = my_cnt;  // inclusive scan over the lists -> exclusive offset of list `lane`
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, my_off, o);
      if (lane >= o) my_off += t;
    }
    my_off -= my_cnt;
    int2 t[RESCORE_STAGE / 32];
#pragma unroll
    for (int i = 0; i < RESCORE_STAGE / 32; ++i) {
      const int g = lane + 32 * i;
      t[i] = make_int2(0, -1);
      // list holding global entry g: the last list whose offset is <= g (lists are few: a ballot per step)
      int l = 0, off_l = 0;
      for (int q = 0; q < a.nsplit; ++q) {
        const int oq = __shfl_sync(FULL, my_off, q), cq = __shfl_sync(FULL, my_cnt, q);
        if (g >= oq && g < oq + cq) {
          l = q;
          off_l = oq;
        }
      }
      if (g < n_total) t[i] = __ldg(cbuf + static_cast<long long>(l) * a.cand_stride + (g - off_l));
    }
#pragma unroll
    for (int i = 0; i < RESCORE_STAGE / 32; ++i) {
      const int g = lane + 32 * i;
      if (g < n_total) stage[g] = t[i];
    }
    __syncwarp();
  }
  const float wn = sqrtf(a.scalars[SC_WNORM_SQ_MAX]);
  const ScreenBound sbd = screen_bound(a.D, a.scalars[SC_RHO], a.scalars[SC_BIAS_ABS_MAX]);
  const float Pb = screen_P(sbd, a.row_norm[b], a.row_dx[b]);  // E_bj = c_j Pb + Qb
  const float Qb = screen_Q(sbd);
  // an encoder row outside the fp16 range makes the whole screen meaningless (inf / nan operands)
  overflow |= !(wn < FP16_MAX) || a.force_unsafe != 0;

  // The lists hold LOWER bounds l_j = h~_j - E_bj of the exact pre-activations.  L = k-th largest lower bound of the
  // row (to 16 bits, rounded down): a column can only be in the exact top-k if its UPPER bound l_j + 2 E_bj reaches L.
  float Lk = -INFINITY;
  const float Lg = a.guess_L != nullptr ? a.guess_L[b] : -INFINITY;
  bool guess_failed = false;
  if (n_total > a.K) {
    unsigned int prefix = 0u;
    int need = a.K;
    auto pass = [&](int shift) {
      if (staged) rescore_radix_pass_staged(stage, n_total, shift, prefix, need, hist, lane);
      else rescore_radix_pass(cbuf, cnts, a.nsplit, a.cand_stride, shift, prefix, need, hist, lane);
    };
    pass(24);
    pass(16);
    Lk = funkey(prefix << 16);
    if (Lg > Lk && Lg <= funkey((prefix << 16) | 0xffffu)) {  // too close to call at 16 bits: the exact k-th largest
      pass(8);
      pass(0);
      Lk = funkey(prefix);
    }
  }
  // The row was screened from a GUESSED threshold: everything whose upper bound reached L_guess was admitted.  If k of
  // the admitted columns have lower bounds >= L_guess, the guess was a true lower bound of the exact k-th largest
  // value and no top-k column can be missing; otherwise the row has to be redone without the guess.
  if (Lg > -INFINITY && !(Lk >= Lg)) guess_failed = true;
  if (a.guess_hist != nullptr && lane == 0 && !overflow) {
    // this row's ratio for the next forward's guess (failed rows land in the lowest bin: they pull the quantile down)
    const float den = a.row_norm[b] * wn;
    float ratio = -1.f;
    if (!guess_failed && n_total > a.K && den > 0.f) ratio = (Lk + a.scalars[SC_BIAS_ABS_MAX]) / den;
    const int bin = min(GUESS_BINS - 1, max(0, static_cast<int>(floorf((ratio + 1.f) * (GUESS_BINS / 2)))));
    if (n_total > a.K || guess_failed) atomicAdd(a.guess_hist + bin, 1);
  }
  if (guess_failed) {
    overflow = true;
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned int*>(a.scalars) + SC_GUESS_FAILED, 1u);
  }
  // ---- collect the survivors (sv = the column's error bound E_bj, used again below) ----
  int n = 0;
  if (staged) {
    for (int e0 = 0; e0 < n_total; e0 += 32) {
      const int e = e0 + lane;
      int2 t = make_int2(0, -1);
      float E = 0.f;
      bool take = false;
      if (e < n_total) {
        t = stage[e];
        E = fmaf(__ldg(a.col_norm + t.y), Pb, Qb);
        take = fmaf(2.f, E, __int_as_float(t.x)) >= Lk;
      }
      const unsigned bal = __ballot_sync(FULL, take);
      const int o = n + __popc(bal & ((1u << lane) - 1u));
      if (take && o < CAP) {
        sv[o] = __int_as_float(t.x) + E;  // the screen value h~_j
        si[o] = t.y;
      }
      n += __popc(bal);
    }
  } else {
    for (int l = 0; l < a.nsplit; ++l) {
      const int c = abs(cnts[l]);
      const int2* lb = cbuf + static_cast<long long>(l) * a.cand_stride;
      for (int e0 = 0; e0 < c; e0 += 32) {
        const int e = e0 + lane;
        int2 t = make_int2(0, -1);
        float E = 0.f;
        bool take = false;
        if (e < c) {
          t = __ldg(lb + e);
          E = fmaf(__ldg(a.col_norm + t.y), Pb, Qb);
          take = fmaf(2.f, E, __int_as_float(t.x)) >= Lk;
        }
        const unsigned bal = __ballot_sync(FULL, take);
        const int o = n + __popc(bal & ((1u << lane) - 1u));
        if (take && o < CAP) {
          sv[o] = __int_as_float(t.x) + E;  // the screen value h~_j
          si[o] = t.y;
        }
        n += __popc(bal);
      }
    }
  }
  if (n > CAP) {  // more near-ties than one warp re-scores: the exact path takes the row
    overflow = true;
    n = CAP;
  }
  if (n < min(a.K, a.S)) overflow = true;  // (cannot happen with a healthy screen: every list keeps >= k entries)
  __syncwarp();

  // ---- exact re-score, RB candidates in flight ----
  float4 xr[VPL];
  const float* xrow = a.x + static_cast<long long>(b) * a.D;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    xr[i] = (v < D4) ? ldg4(xrow + 4 * v) : make_float4(0, 0, 0, 0);
  }
  // RB candidates per round; all RB x VPL row loads are issued before the first multiply so that a warp keeps
  // RB x D x 4 bytes in flight (the loop is bound by the latency of these gathers, not by their volume).
  constexpr int RB = 2;
  if (!overflow) {
    for (int c0 = 0; c0 < n; c0 += RB) {
      int jj[RB];
      float bj[RB];
      float4 w[RB][VPL];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        jj[u] = si[min(c0 + u, n - 1)];
        bj[u] = __ldg(a.b_enc + jj[u]);  // issued with the row loads (a dependent load after the dot costs a round trip)
        const float* r = a.W_enc_t + static_cast<long long>(jj[u]) * a.D;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          w[u][i] = (v < D4) ? ldg4(r + 4 * v) : make_float4(0, 0, 0, 0);
        }
      }
      float acc[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) acc[u] = row_dot<VPL>(xr, w[u]);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < RB; ++u)
          if (c0 + u < n) se[c0 + u] = acc[u] + bj[u];
      }
    }
    __syncwarp();
    // The candidate set covers the exact top-k when every screen error is <= E_bj.  The bound is deterministic; the
    // check below guards its assumptions (accumulation model of the tensor core) on the columns we can see.
    bool bad = false;
    for (int c = lane; c < n; c += 32) bad |= !(fabsf(sv[c] - se[c]) <= fmaf(__ldg(a.col_norm + si[c]), Pb, Qb));  // (NaN -> bad)
    overflow = __any_sync(FULL, bad);
    if (overflow && lane == 0) atomicAdd(reinterpret_cast<unsigned int*>(a.scalars) + SC_UNSAFE_ERR, 1u);
  }
  if (lane == 0) {
    unsigned int* cnt_u = reinterpret_cast<unsigned int*>(a.scalars);
    atomicAdd(cnt_u + SC_RESCORED, static_cast<unsigned int>(n));
    atomicAdd(cnt_u + SC_MERGED, static_cast<unsigned int>(n_total));
  }
  if (overflow) {  // warp-uniform: leave the row to launch_repair_topk (nothing of it has been written)
    if (lane == 0) {
      atomicAdd(reinterpret_cast<unsigned int*>(a.scalars) + SC_UNSAFE_TOTAL, 1u);
      a.unsafe_list[atomicAdd(reinterpret_cast<int*>(a.scalars) + SC_N_UNSAFE, 1)] = b;
    }
    return;
  }

  // ---- final selection on exact values; order: value desc, then column asc ----
  const int k_eff = min(a.K, n);
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int c = c0 + lane;
    if (c < n) {
      const float v = se[c];
      const int id = si[c];
      int rank = 0;
      for (int t = 0; t < n; ++t) {
        const float vt = se[t];
        rank += (vt > v) || (vt == v && si[t] < id);
      }
      if (rank < a.K) {
        const long long o = static_cast<long long>(b) * a.K + rank;
        a.topk_idx[o] = id;
        a.topk_val[o] = v;
        if (a.feat_count) atomicAdd(a.feat_count + id, 1);
        if (a.active && v != 0.f) a.active[id] = 1;
      }
    }
  }
  for (int r = k_eff + lane; r < a.K; r += 32) {  // d_sae < k is rejected at create; defensive
    a.topk_idx[static_cast<long long>(b) * a.K + r] = -1;
    a.topk_val[static_cast<long long>(b) * a.K + r] = 0.f;
  }
}

int launch_rescore_topk(const RescoreArgs& a, cudaStream_t s) {
  if (a.D % 4 || a.K > 128) return 21;
  if (cudaMemsetAsync(reinterpret_cast<int*>(a.scalars) + SC_N_UNSAFE, 0, 4, s) != cudaSuccess) return 23;
  ++g_launch_count;
  const int need_ = (a.D + 127) / 128;
  static const int wpb = [] { const char* v = getenv("SAEV_B200_RESCORE_WPB"); return v ? atoi(v) : 1; }();
#define SB_RESCORE(V)                                                                        \
  {                                                                                          \
    if (a.K > 64) rescore_topk_kernel<V, 1, RESCORE_CAP_WIDE><<<a.B, 32, 0, s>>>(a);         \
    else if (wpb == 1) rescore_topk_kernel<V, 1><<<a.B, 32, 0, s>>>(a);                      \
    else if (wpb == 2) rescore_topk_kernel<V, 2><<<(a.B + 1) / 2, 64, 0, s>>>(a);            \
    else rescore_topk_kernel<V, RESCORE_WARPS><<<(a.B + RESCORE_WARPS - 1) / RESCORE_WARPS, 32 * RESCORE_WARPS, 0, s>>>(a); \
  }
  if (need_ <= 1) SB_RESCORE(1)
  else if (need_ <= 2) SB_RESCORE(2)
  else if (need_ <= 4) SB_RESCORE(4)
  else if (need_ <= 6) SB_RESCORE(6)
  else if (need_ <= 8) SB_RESCORE(8)
  else if (need_ <= 12) SB_RESCORE(12)
  else if (need_ <= 16) SB_RESCORE(16)
  else return 20;
#undef SB_RESCORE
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// Sparse decode + residual + MSE partials (+ d loss / d h on the active set).  One warp per row.
//   x_hat = sum_k f_k W_dec[j_k] + b_dec          (saev modeling.py:386-406, single prefix)
//   r = x_hat - x ; row_sse = sum r^2             (objectives.py:133-138, 224-237)
//   dh_k = grad_scale * <r, W_dec[j_k]> (+ l1/B * sign(f_k))   (autograd of the above)
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256, 2) decode_kernel(DecodeArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= a.B) return;
  const int D4 = a.D >> 2, K = a.K;
  float4 acc[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    acc[i] = (v < D4) ? ldg4(a.b_dec + 4 * v) : make_float4(0, 0, 0, 0);
  }
  const long long kb = static_cast<long long>(b) * K;
  float l1 = 0.f, l0 = 0.f;
  // Both passes gather dictionary rows whose addresses are known up front; every round issues the 2 x VPL loads of two
  // rows before the first FMA so that a warp keeps 2 x D x 4 bytes in flight (the kernel is bound by gather latency).
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int kk = k0 + lane;
    const int mj = (kk < K) ? a.topk_idx[kb + kk] : -1;
    const float mf = (kk < K) ? a.topk_val[kb + kk] : 0.f;
    if (kk < K && mj >= 0) {
      l1 += fabsf(mf);
      l0 += (mf != 0.f) ? 1.f : 0.f;
    }
    const int cnt = min(32, K - k0);
    for (int t = 0; t < cnt; t += 2) {
      const int j0 = __shfl_sync(FULL, mj, t), j1 = (t + 1 < cnt) ? __shfl_sync(FULL, mj, t + 1) : -1;
      const float f0 = __shfl_sync(FULL, mf, t), f1 = __shfl_sync(FULL, mf, (t + 1) & 31);
      const float* r0 = a.W_dec + static_cast<long long>(max(j0, 0)) * a.D;
      const float* r1 = a.W_dec + static_cast<long long>(max(j1, 0)) * a.D;
      float4 w0[VPL], w1[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        w0[i] = (v < D4 && j0 >= 0) ? ldg4(r0 + 4 * v) : make_float4(0, 0, 0, 0);
        w1[i] = (v < D4 && j1 >= 0) ? ldg4(r1 + 4 * v) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        fma4(acc[i], f0, w0[i]);
        fma4(acc[i], f1, w1[i]);
      }
    }
  }
  const float* xrow = a.x + static_cast<long long>(b) * a.D;
  float* rrow = a.resid + static_cast<long long>(b) * a.D;
  float sse = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < D4) {
      const float4 xv = ldg4(xrow + 4 * v);
      float4 r = acc[i];
      r.x -= xv.x; r.y -= xv.y; r.z -= xv.z; r.w -= xv.w;
      acc[i] = r;
      *reinterpret_cast<float4*>(rrow + 4 * v) = r;
      sse += dot4(r, r);
    } else {
      acc[i] = make_float4(0, 0, 0, 0);
    }
  }
  sse = warp_sum(sse);
  l1 = warp_sum(l1);
  l0 = warp_sum(l0);
  if (lane == 0) {
    a.row_sse[b] = sse;
    a.row_l1[b] = l1;
    a.row_l0[b] = l0;
  }
  if (a.dh == nullptr) return;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int kk = k0 + lane;
    const int mj = (kk < K) ? a.topk_idx[kb + kk] : -1;
    const float mf = (kk < K) ? a.topk_val[kb + kk] : 0.f;
    float mine = 0.f;
    const int cnt = min(32, K - k0);
    for (int t = 0; t < cnt; t += 2) {
      const int j0 = __shfl_sync(FULL, mj, t), j1 = (t + 1 < cnt) ? __shfl_sync(FULL, mj, t + 1) : -1;
      const float* r0 = a.W_dec + static_cast<long long>(max(j0, 0)) * a.D;
      const float* r1 = a.W_dec + static_cast<long long>(max(j1, 0)) * a.D;
      float4 w0[VPL], w1[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        w0[i] = (v < D4 && j0 >= 0) ? ldg4(r0 + 4 * v) : make_float4(0, 0, 0, 0);
        w1[i] = (v < D4 && j1 >= 0) ? ldg4(r1 + 4 * v) : make_float4(0, 0, 0, 0);
      }
      float p0 = 0.f, p1 = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        p0 += dot4(acc[i], w0[i]);
        p1 += dot4(acc[i], w1[i]);
      }
      p0 = warp_sum(p0);
      p1 = warp_sum(p1);
      if (lane == t) mine = p0;
      if (lane == t + 1) mine = p1;
    }
    if (kk < K) {
      float d = a.grad_scale * mine;
      if (a.l1_over_b != 0.f) d += a.l1_over_b * ((mf > 0.f) ? 1.f : ((mf < 0.f) ? -1.f : 0.f));
      a.dh[kb + kk] = (mj >= 0) ? d : 0.f;
    }
  }
}

int launch_decode(const DecodeArgs& a, cudaStream_t s) {
  if (a.D % 4) return 21;
  static const int wpb = [] { const char* v = getenv("SAEV_B200_DECODE_WPB"); const int w = v ? atoi(v) : 8; return (w >= 1 && w <= 8) ? w : 8; }();
  SB_DISPATCH_VPL(a.D, (decode_kernel<VPL><<<(a.B + wpb - 1) / wpb, 32 * wpb, 0, s>>>(a)));
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// Matryoshka decode (saev modeling.py:364-406, objectives.py:124-138): the active columns of a row are walked in
// ascending column order; whenever the walk crosses a prefix cut the running reconstruction is that prefix's x_hat.
// One warp per row; K <= 128 (two or four top-k slots per lane).
// ------------------------------------------------------------------------------------------------
// NS = top-k slots per lane: 2 (K <= 64) or 4 (K <= 128, the row capacity of the BatchTopK path)
template <int VPL, int NS>
__global__ void __launch_bounds__(256, 2) decode_prefix_kernel(DecodeArgs a, PrefixCuts pf, float* __restrict__ sfx) {
  __shared__ int order_s[8][32 * NS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= a.B) return;
  const int D4 = a.D >> 2, K = a.K, P = pf.n;
  const long long kb = static_cast<long long>(b) * K;
  // this lane's (up to) NS slots
  int mj[NS];
  float mf[NS];
  float l1 = 0.f, l0 = 0.f;
#pragma unroll
  for (int u = 0; u < NS; ++u) {
    const int kk = lane + 32 * u;
    mj[u] = (kk < K) ? a.topk_idx[kb + kk] : -1;
    mf[u] = (kk < K) ? a.topk_val[kb + kk] : 0.f;
    if (mj[u] >= 0) {
      l1 += fabsf(mf[u]);
      l0 += (mf[u] != 0.f) ? 1.f : 0.f;
    }
  }
  // rank of every slot by column (empty slots last, ties among them by slot)
  int* order = order_s[warp];
  {
    int rank[NS];
    unsigned int key[NS];
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      rank[u] = 0;
      key[u] = mj[u] < 0 ? 0x7fffffffu : static_cast<unsigned int>(mj[u]);
    }
    for (int t = 0; t < K; ++t) {
      const int src = t & 31, su = t >> 5;
      unsigned int kt = 0u;
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        const unsigned int ku = __shfl_sync(FULL, key[u], src);
        if (u == su) kt = ku;
      }
#pragma unroll
      for (int u = 0; u < NS; ++u) rank[u] += (kt < key[u]) || (kt == key[u] && t < lane + 32 * u);
    }
#pragma unroll
    for (int u = 0; u < NS; ++u)
      if (lane + 32 * u < K) order[rank[u]] = lane + 32 * u;
  }
  __syncwarp();

  float4 acc[VPL];
  const float* xrow = a.x + static_cast<long long>(b) * a.D;  // re-read (L1) at every cut instead of held in registers
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    acc[i] = (v < D4) ? ldg4(a.b_dec + 4 * v) : make_float4(0, 0, 0, 0);
  }
  float* srow = sfx + static_cast<long long>(b) * P * a.D;
  float sse = 0.f;
  auto emit = [&](int c) {  // r_c = running x_hat - x
    float* o = srow + static_cast<long long>(c) * a.D;
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < D4) {
        const float4 xv = ldg4(xrow + 4 * v);
        float4 r = acc[i];
        r.x -= xv.x; r.y -= xv.y; r.z -= xv.z; r.w -= xv.w;
        *reinterpret_cast<float4*>(o + 4 * v) = r;
        part += dot4(r, r);
      }
    }
    sse += part;
  };
  int cur = 0;
  // slot -> (column, value) of the t-th active in ascending column order (empty slots sort last)
  auto entry = [&](int t, int& j, float& f) {
    const int k = order[min(t, K - 1)];
    int jj = -1;
    float ff = 0.f;
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      const int ju = __shfl_sync(FULL, mj[u], k & 31);
      const float fu = __shfl_sync(FULL, mf[u], k & 31);
      if ((k >> 5) == u) {
        jj = ju;
        ff = fu;
      }
    }
    j = (t < K) ? jj : -1;
    f = ff;
  };
  // two dictionary rows per round, all 2 x VPL loads issued before the first FMA (gather latency, see decode_kernel)
  for (int t = 0; t < K; t += 2) {
    int ja, jb;
    float fa, fb;
    entry(t, ja, fa);
    entry(t + 1, jb, fb);
    if (ja < 0) break;
    const float* r0 = a.W_dec + static_cast<long long>(ja) * a.D;
    const float* r1 = a.W_dec + static_cast<long long>(max(jb, 0)) * a.D;
    float4 w0[VPL], w1[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      w0[i] = (v < D4) ? ldg4(r0 + 4 * v) : make_float4(0, 0, 0, 0);
      w1[i] = (v < D4 && jb >= 0) ? ldg4(r1 + 4 * v) : make_float4(0, 0, 0, 0);
    }
    while (cur < P - 1 && ja >= pf.cut[cur]) {
      emit(cur);
      ++cur;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) fma4(acc[i], fa, w0[i]);
    if (jb >= 0) {
      while (cur < P - 1 && jb >= pf.cut[cur]) {
        emit(cur);
        ++cur;
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) fma4(acc[i], fb, w1[i]);
    }
  }
  while (cur < P) {
    emit(cur);
    ++cur;
  }
  // last prefix = the full reconstruction: residual for AuxK / logging
  {
    float* rrow = a.resid + static_cast<long long>(b) * a.D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < D4) {
        const float4 xv = ldg4(xrow + 4 * v);
        float4 r = acc[i];
        r.x -= xv.x; r.y -= xv.y; r.z -= xv.z; r.w -= xv.w;
        *reinterpret_cast<float4*>(rrow + 4 * v) = r;
      }
    }
  }
  sse = warp_sum(sse);
  l1 = warp_sum(l1);
  l0 = warp_sum(l0);
  if (lane == 0) {
    a.row_sse[b] = sse;
    a.row_l1[b] = l1;
    a.row_l0[b] = l0;
  }
  // suffix sums in place: sfx[c] = sum_{i >= c} r_i   (each lane re-reads only what it wrote itself)
#pragma unroll
  for (int i = 0; i < VPL; ++i) acc[i] = make_float4(0, 0, 0, 0);
  for (int c = P - 1; c >= 0; --c) {
    float* o = srow + static_cast<long long>(c) * a.D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < D4) {
        const float4 r = *reinterpret_cast<const float4*>(o + 4 * v);
        acc[i].x += r.x; acc[i].y += r.y; acc[i].z += r.z; acc[i].w += r.w;
        *reinterpret_cast<float4*>(o + 4 * v) = acc[i];
      }
    }
  }
  if (a.dh == nullptr) return;
  // dh_k = grad_scale * <sfx[c(j_k)], W_dec[j_k]>  (acc holds sfx[0] now; reload when the block changes)
  cur = 0;
  for (int t = 0; t < K; t += 2) {
    int jj[2];
    float ff[2];
    entry(t, jj[0], ff[0]);
    entry(t + 1, jj[1], ff[1]);
    float4 w0[VPL], w1[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      w0[i] = (v < D4 && jj[0] >= 0) ? ldg4(a.W_dec + static_cast<long long>(jj[0]) * a.D + 4 * v) : make_float4(0, 0, 0, 0);
      w1[i] = (v < D4 && jj[1] >= 0) ? ldg4(a.W_dec + static_cast<long long>(jj[1]) * a.D + 4 * v) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (t + u >= K) break;
      const int j = jj[u];
      float d = 0.f;
      if (j >= 0) {
        int c = cur;
        while (c < P - 1 && j >= pf.cut[c]) ++c;
        if (c != cur) {  // the walk entered a new prefix block: its suffix sum replaces the one in registers
          cur = c;
          const float* o = srow + static_cast<long long>(c) * a.D;
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < D4) acc[i] = *reinterpret_cast<const float4*>(o + 4 * v);
          }
        }
        float pdot = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) pdot += dot4(acc[i], u == 0 ? w0[i] : w1[i]);
        pdot = warp_sum(pdot);
        d = a.grad_scale * pdot;
        const float f = ff[u];
        if (a.l1_over_b != 0.f) d += a.l1_over_b * ((f > 0.f) ? 1.f : ((f < 0.f) ? -1.f : 0.f));
      }
      const int k = order[t + u];
      if (lane == 0) a.dh[kb + k] = d;
    }
  }
}

int launch_decode_prefix(const DecodeArgs& a, const PrefixCuts& pf, float* sfx, cudaStream_t s) {
  if (a.D % 4) return 21;
  if (a.K > 128 || pf.n < 1 || pf.n > MAX_PREFIXES) return 24;
  if (a.K <= 64) { SB_DISPATCH_VPL(a.D, (decode_prefix_kernel<VPL, 2><<<(a.B + 7) / 8, 256, 0, s>>>(a, pf, sfx))); }
  else { SB_DISPATCH_VPL(a.D, (decode_prefix_kernel<VPL, 4><<<(a.B + 7) / 8, 256, 0, s>>>(a, pf, sfx))); }
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

__global__ void x_hats_prefix_kernel(const float* __restrict__ sfx, const float* __restrict__ x, int B, int D, int P,
                                     float* __restrict__ out) {
  const long long n = static_cast<long long>(B) * P * D;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int d = static_cast<int>(i % D);
    const long long bp = i / D;
    const int c = static_cast<int>(bp % P);
    const long long b = bp / P;
    const float nxt = (c + 1 < P) ? sfx[i + D] : 0.f;
    out[i] = x[b * D + d] + sfx[i] - nxt;
  }
}
int launch_x_hats_prefix(const float* sfx, const float* x, int B, int D, int P, float* out, cudaStream_t s) {
  x_hats_prefix_kernel<<<148 * 8, 256, 0, s>>>(sfx, x, B, D, P, out);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// CSC of the active set: for every dictionary atom j the list of (b, k) slots where it fired.
// ------------------------------------------------------------------------------------------------
// Two-level exclusive scan of the per-atom counts (1024 atoms per block): block totals first, then every block
// adds the totals of the blocks before it to its local scan.
__global__ void __launch_bounds__(1024) block_totals_kernel(const int* __restrict__ cnt, int S, int* __restrict__ totals,
                                                            int* __restrict__ n_heavy) {
  __shared__ int ws[32];
  if (n_heavy != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *n_heavy = 0;  // appended to by scan_counts_kernel
  const int i = blockIdx.x * 1024 + threadIdx.x;
  int c = (i < S) ? cnt[i] : 0;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = ws[threadIdx.x];
    t = warp_sum(t);
    if (threadIdx.x == 0) totals[blockIdx.x] = t;
  }
}

// prefix = sum of totals[0 .. blockIdx.x), computed by the whole block
__device__ __forceinline__ int block_prefix(const int* __restrict__ totals, int* ws) {
  int t = 0;
  for (int b = threadIdx.x; b < static_cast<int>(blockIdx.x); b += 1024) t += totals[b];
  t = warp_sum(t);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = t;
  __syncthreads();
  int r = 0;
  if (threadIdx.x < 32) r = warp_sum(ws[threadIdx.x]);
  __syncthreads();
  if (threadIdx.x == 0) ws[0] = r;
  __syncthreads();
  r = ws[0];
  __syncthreads();
  return r;
}

// inclusive scan of `v` over the 1024 threads of the block; returns this thread's inclusive value, block total in *tot
__device__ __forceinline__ int block_scan_incl(int v, int* ws, int* tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = ws[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, w, o);
      if (lane >= o) w += t;
    }
    ws[lane] = w;
  }
  __syncthreads();
  const int base = warp > 0 ? ws[warp - 1] : 0;
  *tot = ws[31];
  __syncthreads();
  return base + incl;
}

__global__ void __launch_bounds__(1024) scan_counts_kernel(const int* __restrict__ cnt, const int* __restrict__ totals,
                                                           int* __restrict__ off, int* __restrict__ cursor, int S,
                                                           int heavy_thr, int* __restrict__ heavy_list,
                                                           int* __restrict__ n_heavy) {
  __shared__ int ws[32];
  const int prefix = block_prefix(totals, ws);
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int c = (i < S) ? cnt[i] : 0;
  // atoms with very long lists ("dense" features) get a whole block in wgrad_heavy_kernel instead of one warp
  if (heavy_list != nullptr && c > heavy_thr) heavy_list[atomicAdd(n_heavy, 1)] = i;
  int tot;
  const int incl = block_scan_incl(c, ws, &tot);
  if (i < S) {
    off[i] = prefix + incl - c;
    cursor[i] = 0;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) off[S] = prefix + tot;
}

__global__ void csc_fill_kernel(const int* __restrict__ idx, long long n, const int* __restrict__ off,
                                int* __restrict__ cursor, int* __restrict__ entries) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int j = idx[p];
  if (j < 0) return;
  const int pos = off[j] + atomicAdd(cursor + j, 1);
  entries[pos] = static_cast<int>(p);
}

int launch_csc_build(const int* topk_idx, int B, int K, int S, const int* feat_count, int* feat_off, int* cursor,
                     int* entries, int* block_totals, cudaStream_t s, int* heavy_list, int* n_heavy) {
  const int nb = (S + 1023) / 1024;
  block_totals_kernel<<<nb, 1024, 0, s>>>(feat_count, S, block_totals, n_heavy);
  ++g_launch_count;
  scan_counts_kernel<<<nb, 1024, 0, s>>>(feat_count, block_totals, feat_off, cursor, S, WGRAD_HEAVY_ENTRIES, heavy_list,
                                         n_heavy);
  ++g_launch_count;
  const long long n = static_cast<long long>(B) * K;
  csc_fill_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(topk_idx, n, feat_off, cursor, entries);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// Weight gradients of the sparse path, one warp per dictionary atom j (autograd of decode/encode):
//   gW_dec[j]   = grad_scale * sum_{(b,k): idx=j} f_bk * r_b      then minus its component along W_dec[j]
//                 (saev modeling.py:419-445 remove_parallel_grads)
//   gW_enc_t[j] = sum dh_bk * x_b ;  gb_enc[j] = sum dh_bk
// Atoms that did not fire get zero rows (the dense .grad tensors saev's loop expects).
// ------------------------------------------------------------------------------------------------
// FUSE_DH: d loss / d h of every active entry is computed here, dh_bk = grad_scale * <r_b, W_dec[j]> (+ l1 / B * sign f),
// from the residual row the entry gathers anyway and this atom's dictionary row (staged in shared memory) -- which
// removes the second gather pass of the decode kernel (K dictionary rows per sample).
template <int VPL, bool FUSE_DH, int FUSE_WPB = 8>
__global__ void __launch_bounds__(256, 2) wgrad_kernel(WgradArgs a) {
  __shared__ float4 wsm[FUSE_DH ? FUSE_WPB * VPL * 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warps of a block hold their SM slot until the slowest one is done and the list lengths vary (Poisson around
  // B K / S), so small blocks (a.warps_per_block, default 1) keep more warps busy
  const int j = a.row_begin + blockIdx.x * (blockDim.x >> 5) + warp;
  if (j >= a.row_end) return;
  const int D4 = a.D >> 2;
  const int beg = a.feat_off[j], end = a.feat_off[j + 1];
  // staged backward: the AuxK kernels have already written the rows of the dead atoms (which never fire)
  if (a.skip_toks != nullptr && beg == end && a.skip_toks[j] >= a.skip_threshold) return;
  if (a.heavy_list != nullptr && end - beg > WGRAD_HEAVY_ENTRIES) {  // wgrad_heavy_kernel's: hand it zeroed rows
    float* gdrow0 = a.gW_dec + static_cast<long long>(j) * a.D;
    float* gerow0 = a.gW_enc_t + static_cast<long long>(j) * a.D;
    for (int v = lane; v < D4; v += 32) {
      *reinterpret_cast<float4*>(gdrow0 + 4 * v) = make_float4(0, 0, 0, 0);
      *reinterpret_cast<float4*>(gerow0 + 4 * v) = make_float4(0, 0, 0, 0);
    }
    if (lane == 0) {
      a.gb_enc[j] = 0.f;
      a.heavy_ticket[j] = 0;
    }
    return;
  }
  float4 gd[VPL], ge[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) gd[i] = ge[i] = make_float4(0, 0, 0, 0);
  float sdh = 0.f;
  float4* wmine = wsm + (FUSE_DH ? warp * VPL * 32 : 0);
  if (FUSE_DH && beg != end) {
    const float* wrow0 = a.W_dec + static_cast<long long>(j) * a.D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      wmine[v] = (v < D4) ? __ldcs(reinterpret_cast<const float4*>(wrow0 + 4 * v)) : make_float4(0, 0, 0, 0);
    }
    __syncwarp();
  }
  // Matryoshka: column j sits in prefix block c(j) and sees the suffix sum of the residuals of prefixes >= c(j)
  const float* rbase = a.resid;
  long long rstride = a.D;
  if (a.sfx != nullptr) {
    int c = 0;
    while (c < a.pf.n - 1 && j >= a.pf.cut[c]) ++c;
    rbase = a.sfx + static_cast<long long>(c) * a.D;
    rstride = static_cast<long long>(a.pf.n) * a.D;
  }
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    int mb = 0;
    float mf = 0.f, md = 0.f;
    if (e < end) {
      const int p = a.entries[e];
      mb = p / a.K;
      mf = a.topk_val[p];
      if (!FUSE_DH) {
        md = a.dh[p];
        sdh += md;
      }
    }
    const int cnt = min(32, end - e0);
    if (FUSE_DH) {
      // (measured at c3 with one atom per block: decode 0.47 -> 0.27 ms, this kernel 0.62 -> 0.74 ms; a two-loop variant
      //  that separates the residual and the x gathers was slower, as were two warps per atom and L2 evict_last hints)
      float sd = 0.f;
#pragma unroll 2
      for (int t = 0; t < cnt; ++t) {
        const int bb = __shfl_sync(FULL, mb, t);
        const float f = __shfl_sync(FULL, mf, t);
        const float* rrow = rbase + static_cast<long long>(bb) * rstride;
        const float* xrow = a.x + static_cast<long long>(bb) * a.D;
        float4 xv[VPL];
        float pd = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          xv[i] = (v < D4) ? ldg4(xrow + 4 * v) : make_float4(0, 0, 0, 0);  // issued before the reduction below
          if (v < D4) {
            const float4 rv = ldg4(rrow + 4 * v);
            fma4(gd[i], f, rv);
            pd += dot4(rv, wmine[v]);
          }
        }
        pd = warp_sum(pd);
        float d = a.grad_scale * pd;
        if (a.l1_over_b != 0.f) d += a.l1_over_b * ((f > 0.f) ? 1.f : ((f < 0.f) ? -1.f : 0.f));
        sd += d;  // warp-uniform
#pragma unroll
        for (int i = 0; i < VPL; ++i) fma4(ge[i], d, xv[i]);
      }
      if (lane == 0) sdh += sd;
    } else {
#pragma unroll 2
      for (int t = 0; t < cnt; ++t) {
        const int bb = __shfl_sync(FULL, mb, t);
        const float f = __shfl_sync(FULL, mf, t);
        const float d = __shfl_sync(FULL, md, t);
        const float* rrow = rbase + static_cast<long long>(bb) * rstride;
        const float* xrow = a.x + static_cast<long long>(bb) * a.D;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (v < D4) {
            fma4(gd[i], f, ldg4(rrow + 4 * v));
            fma4(ge[i], d, ldg4(xrow + 4 * v));
          }
        }
      }
    }
  }
  sdh = warp_sum(sdh);
  float* gdrow = a.gW_dec + static_cast<long long>(j) * a.D;
  float* gerow = a.gW_enc_t + static_cast<long long>(j) * a.D;
  if (beg == end) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < D4) {
        __stcs(reinterpret_cast<float4*>(gdrow + 4 * v), make_float4(0, 0, 0, 0));
        __stcs(reinterpret_cast<float4*>(gerow + 4 * v), make_float4(0, 0, 0, 0));
      }
    }
    if (lane == 0) {
      a.gb_enc[j] = 0.f;
      if (a.row_gsq) a.row_gsq[j] = 0.f;
    }
    return;
  }
  const float* wrow = a.W_dec + static_cast<long long>(j) * a.D;
  float4 w[VPL];
  float dot = 0.f, nsq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    gd[i].x *= a.grad_scale; gd[i].y *= a.grad_scale; gd[i].z *= a.grad_scale; gd[i].w *= a.grad_scale;
    if (FUSE_DH) w[i] = wmine[v];
    else w[i] = (v < D4) ? __ldcs(reinterpret_cast<const float4*>(wrow + 4 * v)) : make_float4(0, 0, 0, 0);
    dot += dot4(gd[i], w[i]);
    nsq += dot4(w[i], w[i]);
  }
  if (a.remove_parallel) {
    dot = warp_sum(dot);
    nsq = warp_sum(nsq);
    const float sc = (nsq > 0.f) ? dot / nsq : 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) fma4(gd[i], -sc, w[i]);
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < D4) {  // streaming stores: the 537 MB of gradients must not evict x / resid (the gather set) from L2
      __stcs(reinterpret_cast<float4*>(gdrow + 4 * v), gd[i]);
      __stcs(reinterpret_cast<float4*>(gerow + 4 * v), ge[i]);
    }
  }
  if (a.row_gsq != nullptr) {  // this atom's share of ||g||^2 (saves the separate pass of clip_grad_norm_)
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) ss += dot4(gd[i], gd[i]) + dot4(ge[i], ge[i]);
    ss = warp_sum(ss);
    if (lane == 0) a.row_gsq[j] = ss + sdh * sdh;
  }
  if (lane == 0) a.gb_enc[j] = sdh;
}

// Atoms whose list is longer than WGRAD_HEAVY_ENTRIES ("dense" features that fire on a large share of the batch).  One
// warp per atom would stream the whole list alone (a feature firing on every row of a 16 k batch = 134 MB of gathers;
// measured 14 ms for 8 such atoms at c3), so the list is cut into slices of WGRAD_SLICE entries and every
// (atom, slice) is one block: its 8 warps take interleaved 32-entry chunks, fold their partial rows in a fixed order
// through shared memory and warp 0 adds the slice's row into the gradient rows with vector RED (wgrad_kernel zeroed
// them and the atom's ticket when it skipped the atom).  The block that draws the last ticket finishes the row
// (scale, projection, ||g||^2) as wgrad_kernel does.
constexpr int WGRAD_SLICE = 1024;
__device__ __forceinline__ void red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int VPL, bool FUSE_DH>
__global__ void __launch_bounds__(256, 2) wgrad_heavy_kernel(WgradArgs a) {
  __shared__ float4 sgd[VPL * 32], sge[VPL * 32];
  __shared__ float4 swr[FUSE_DH ? VPL * 32 : 1];  // this atom's dictionary row (dh computed here, see wgrad_kernel)
  __shared__ float ssd[8];
  __shared__ int s_ticket;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D4 = a.D >> 2;
  const int n_heavy = *a.n_heavy;
  int slice_base = 0;  // slices of the atoms walked so far: slice g of the whole list goes to block g % gridDim.x
  for (int hi = 0; hi < n_heavy; ++hi) {
    const int j = a.heavy_list[hi];
    if (j < a.row_begin || j >= a.row_end) continue;  // block-uniform
    const int beg = a.feat_off[j], end = a.feat_off[j + 1];
    const int n_slices = (end - beg + WGRAD_SLICE - 1) / WGRAD_SLICE;
    const int first = static_cast<int>((blockIdx.x + gridDim.x - slice_base % gridDim.x) % gridDim.x);
    slice_base += n_slices;
    const float* rbase = a.resid;
    long long rstride = a.D;
    if (a.sfx != nullptr) {
      int c = 0;
      while (c < a.pf.n - 1 && j >= a.pf.cut[c]) ++c;
      rbase = a.sfx + static_cast<long long>(c) * a.D;
      rstride = static_cast<long long>(a.pf.n) * a.D;
    }
    float* gdrow = a.gW_dec + static_cast<long long>(j) * a.D;
    float* gerow = a.gW_enc_t + static_cast<long long>(j) * a.D;
    if (FUSE_DH && first < n_slices) {
      const float* wrow0 = a.W_dec + static_cast<long long>(j) * a.D;
      for (int v = threadIdx.x; v < VPL * 32; v += 256)
        swr[v] = (v < D4) ? __ldg(reinterpret_cast<const float4*>(wrow0 + 4 * v)) : make_float4(0, 0, 0, 0);
      __syncthreads();
    }
    for (int sl = first; sl < n_slices; sl += gridDim.x) {
      const int sbeg = beg + sl * WGRAD_SLICE, send = min(end, sbeg + WGRAD_SLICE);
      float4 gd[VPL], ge[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) gd[i] = ge[i] = make_float4(0, 0, 0, 0);
      float sdh = 0.f;
      for (int e0 = sbeg + warp * 32; e0 < send; e0 += 8 * 32) {
        const int e = e0 + lane;
        int mb = 0;
        float mf = 0.f, md = 0.f;
        if (e < send) {
          const int p = a.entries[e];
          mb = p / a.K;
          mf = a.topk_val[p];
          if (!FUSE_DH) {
            md = a.dh[p];
            sdh += md;
          }
        }
        const int cnt = min(32, send - e0);
        float sd = 0.f;
#pragma unroll 2
        for (int t = 0; t < cnt; ++t) {
          const int bb = __shfl_sync(FULL, mb, t);
          const float f = __shfl_sync(FULL, mf, t);
          const float* rrow = rbase + static_cast<long long>(bb) * rstride;
          const float* xrow = a.x + static_cast<long long>(bb) * a.D;
          if (FUSE_DH) {
            float4 xv[VPL];
            float pd = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
              const int v = lane + 32 * i;
              xv[i] = (v < D4) ? ldg4(xrow + 4 * v) : make_float4(0, 0, 0, 0);
              if (v < D4) {
                const float4 rv = ldg4(rrow + 4 * v);
                fma4(gd[i], f, rv);
                pd += dot4(rv, swr[v]);
              }
            }
            pd = warp_sum(pd);
            float d = a.grad_scale * pd;
            if (a.l1_over_b != 0.f) d += a.l1_over_b * ((f > 0.f) ? 1.f : ((f < 0.f) ? -1.f : 0.f));
            sd += d;  // warp-uniform
#pragma unroll
            for (int i = 0; i < VPL; ++i) fma4(ge[i], d, xv[i]);
          } else {
            const float d = __shfl_sync(FULL, md, t);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
              const int v = lane + 32 * i;
              if (v < D4) {
                fma4(gd[i], f, ldg4(rrow + 4 * v));
                fma4(ge[i], d, ldg4(xrow + 4 * v));
              }
            }
          }
        }
        if (FUSE_DH && lane == 0) sdh += sd;
      }
      sdh = warp_sum(sdh);
      if (lane == 0) ssd[warp] = sdh;
      for (int w = 7; w >= 0; --w) {  // warp 7 stores, 6 .. 1 add, warp 0 ends up with the slice total in registers
        if (warp == w) {
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (w < 7) {
              const float4 pd = sgd[v], pe = sge[v];
              gd[i].x += pd.x; gd[i].y += pd.y; gd[i].z += pd.z; gd[i].w += pd.w;
              ge[i].x += pe.x; ge[i].y += pe.y; ge[i].z += pe.z; ge[i].w += pe.w;
            }
            if (w > 0) {
              sgd[v] = gd[i];
              sge[v] = ge[i];
            }
          }
        }
        __syncthreads();
      }
      if (warp == 0) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (v < D4) {
            red_add4(gdrow + 4 * v, gd[i]);
            red_add4(gerow + 4 * v, ge[i]);
          }
        }
        if (lane == 0) atomicAdd(a.gb_enc + j, ((ssd[0] + ssd[1]) + (ssd[2] + ssd[3])) + ((ssd[4] + ssd[5]) + (ssd[6] + ssd[7])));
        __threadfence();
        __syncwarp();
        if (lane == 0) s_ticket = atomicAdd(a.heavy_ticket + j, 1);
      }
      __syncthreads();
      const bool last = s_ticket == n_slices - 1;
      __syncthreads();  // s_ticket / ssd / sgd are reused by the next slice of this block
      if (last && warp == 0) {
        __threadfence();
        const float* wrow = a.W_dec + static_cast<long long>(j) * a.D;
        float4 w[VPL];
        float dot = 0.f, nsq = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          gd[i] = (v < D4) ? __ldcg(reinterpret_cast<const float4*>(gdrow + 4 * v)) : make_float4(0, 0, 0, 0);
          ge[i] = (v < D4) ? __ldcg(reinterpret_cast<const float4*>(gerow + 4 * v)) : make_float4(0, 0, 0, 0);
          gd[i].x *= a.grad_scale; gd[i].y *= a.grad_scale; gd[i].z *= a.grad_scale; gd[i].w *= a.grad_scale;
          w[i] = (v < D4) ? __ldcs(reinterpret_cast<const float4*>(wrow + 4 * v)) : make_float4(0, 0, 0, 0);
          dot += dot4(gd[i], w[i]);
          nsq += dot4(w[i], w[i]);
        }
        if (a.remove_parallel) {
          dot = warp_sum(dot);
          nsq = warp_sum(nsq);
          const float sc = (nsq > 0.f) ? dot / nsq : 0.f;
#pragma unroll
          for (int i = 0; i < VPL; ++i) fma4(gd[i], -sc, w[i]);
        }
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          if (v < D4) {
            __stcs(reinterpret_cast<float4*>(gdrow + 4 * v), gd[i]);
            ss += dot4(gd[i], gd[i]) + dot4(ge[i], ge[i]);
          }
        }
        ss = warp_sum(ss);
        if (lane == 0 && a.row_gsq != nullptr) {
          const float sd = __ldcg(a.gb_enc + j);
          a.row_gsq[j] = ss + sd * sd;
        }
      }
    }
  }
}

static int launch_wgrad_light(const WgradArgs& a, cudaStream_t s) {
  if (a.D % 4) return 21;
  const int rows = a.row_end - a.row_begin;
  if (rows <= 0) return 0;
  const int wpb = (a.warps_per_block >= 1 && a.warps_per_block <= 8) ? a.warps_per_block : 1;
  const int nthr = 32 * wpb, nblk = (rows + wpb - 1) / wpb;
  if (a.dh == nullptr) {  // fused dh: dictionary row staged in (static) shared memory, d_model <= 1024
    const int need = (a.D + 127) / 128;
    const int nthr = 256, nblk = (rows + 7) / 8;  // (the shared-memory stage is sized for 8 warps, or for 1)
    const int fw = wpb == 1 ? 1 : 8;
    ++g_launch_count;
    if (need <= 1) { if (fw == 1) wgrad_kernel<1, true, 1><<<rows, 32, 0, s>>>(a); else wgrad_kernel<1, true, 8><<<nblk, nthr, 0, s>>>(a); }
    else if (need <= 2) { if (fw == 1) wgrad_kernel<2, true, 1><<<rows, 32, 0, s>>>(a); else wgrad_kernel<2, true, 8><<<nblk, nthr, 0, s>>>(a); }
    else if (need <= 4) { if (fw == 1) wgrad_kernel<4, true, 1><<<rows, 32, 0, s>>>(a); else wgrad_kernel<4, true, 8><<<nblk, nthr, 0, s>>>(a); }
    else if (need <= 6) { if (fw == 1) wgrad_kernel<6, true, 1><<<rows, 32, 0, s>>>(a); else wgrad_kernel<6, true, 8><<<nblk, nthr, 0, s>>>(a); }
    else if (need <= 8) { if (fw == 1) wgrad_kernel<8, true, 1><<<rows, 32, 0, s>>>(a); else wgrad_kernel<8, true, 8><<<nblk, nthr, 0, s>>>(a); }
    else return 20;
    return cudaGetLastError() == cudaSuccess ? 0 : 22;
  }
  SB_DISPATCH_VPL(a.D, (wgrad_kernel<VPL, false><<<nblk, nthr, 0, s>>>(a)));
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}


// ------------------------------------------------------------------------------------------------
// Two-pass form of the weight gradients (the default; SAEV_B200_WGRAD_SPLIT=0 selects the one-pass kernel above, which
// gathers a residual row
// AND an x row per active entry: a 134 MB gather set at c3 that the L2 serves only half of (ncu: 2.15 GB of the 4.3 GB
// of gathers come from DRAM).  Here the decoder side runs first over ALL atoms -- gW_dec, gb_enc, and dh_bk =
// grad_scale <r_b, w_j> written to the [B, K] scratch, gathering residual rows only (67 MB) -- then the encoder side
// gathers x rows only.  Each pass works on a gather set that fits the L2.  Heavy atoms (lists > WGRAD_HEAVY_ENTRIES) stay
// with wgrad_heavy_kernel, which both passes prepare exactly as wgrad_kernel does.
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(32, 16) wgrad_dec_kernel(WgradArgs a, float* __restrict__ dh_out) {
  __shared__ float4 w[VPL * 32];  // this atom's dictionary row (lane-strided: no bank conflicts)
  const int lane = threadIdx.x;
  const int j = a.row_begin + blockIdx.x;
  if (j >= a.row_end) return;
  const int D4 = a.D >> 2;
  const int beg = a.feat_off[j], end = a.feat_off[j + 1];
  if (a.skip_toks != nullptr && beg == end && a.skip_toks[j] >= a.skip_threshold) return;
  float* gdrow = a.gW_dec + static_cast<long long>(j) * a.D;
  if ((a.heavy_list != nullptr && end - beg > WGRAD_HEAVY_ENTRIES) || beg == end) {
    // heavy: wgrad_heavy_kernel adds its slices into zeroed rows; empty: the dense .grad rows saev's loop expects
    for (int v = lane; v < D4; v += 32) __stcs(reinterpret_cast<float4*>(gdrow + 4 * v), make_float4(0, 0, 0, 0));
    if (lane == 0) {
      a.gb_enc[j] = 0.f;
      if (beg != end) a.heavy_ticket[j] = 0;
      else if (a.row_gsq) a.row_gsq[j] = 0.f;
    }
    return;
  }
  float4 gd[VPL];
  const float* wrow = a.W_dec + static_cast<long long>(j) * a.D;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    gd[i] = make_float4(0, 0, 0, 0);
    w[v] = (v < D4) ? __ldcs(reinterpret_cast<const float4*>(wrow + 4 * v)) : make_float4(0, 0, 0, 0);
  }
  __syncwarp();
  const float* rbase = a.resid;
  long long rstride = a.D;
  if (a.sfx != nullptr) {
    int c = 0;
    while (c < a.pf.n - 1 && j >= a.pf.cut[c]) ++c;
    rbase = a.sfx + static_cast<long long>(c) * a.D;
    rstride = static_cast<long long>(a.pf.n) * a.D;
  }
  float sdh = 0.f;
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    int mp = 0;
    float mf = 0.f, mine = 0.f;
    if (e < end) {
      mp = a.entries[e];
      mf = a.topk_val[mp];
    }
    const int cnt = min(32, end - e0);
    for (int t = 0; t < cnt; t += 2) {  // two residual rows in flight
      const int p0 = __shfl_sync(FULL, mp, t), p1 = __shfl_sync(FULL, mp, min(t + 1, cnt - 1));
      const float f0 = __shfl_sync(FULL, mf, t), f1 = (t + 1 < cnt) ? __shfl_sync(FULL, mf, t + 1) : 0.f;
      const float* r0 = rbase + static_cast<long long>(p0 / a.K) * rstride;
      const float* r1 = rbase + static_cast<long long>(p1 / a.K) * rstride;
      float4 v0[VPL], v1[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        v0[i] = (v < D4) ? ldg4(r0 + 4 * v) : make_float4(0, 0, 0, 0);
        v1[i] = (v < D4) ? ldg4(r1 + 4 * v) : make_float4(0, 0, 0, 0);
      }
      float pd0 = 0.f, pd1 = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        fma4(gd[i], f0, v0[i]);
        fma4(gd[i], f1, v1[i]);  // (f1 == 0 for the padding round)
        const float4 wv = w[lane + 32 * i];
        pd0 += dot4(v0[i], wv);
        pd1 += dot4(v1[i], wv);
      }
      pd0 = warp_sum(pd0);
      pd1 = warp_sum(pd1);
      float d0 = a.grad_scale * pd0, d1 = a.grad_scale * pd1;
      if (a.l1_over_b != 0.f) {
        d0 += a.l1_over_b * ((f0 > 0.f) ? 1.f : ((f0 < 0.f) ? -1.f : 0.f));
        d1 += a.l1_over_b * ((f1 > 0.f) ? 1.f : ((f1 < 0.f) ? -1.f : 0.f));
      }
      if (lane == t) mine = d0;
      if (lane == t + 1 && t + 1 < cnt) mine = d1;
    }
    if (e < end) {
      dh_out[mp] = mine;
      sdh += mine;
    }
  }
  sdh = warp_sum(sdh);
  float dot = 0.f, nsq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    gd[i].x *= a.grad_scale; gd[i].y *= a.grad_scale; gd[i].z *= a.grad_scale; gd[i].w *= a.grad_scale;
    dot += dot4(gd[i], w[lane + 32 * i]);
    nsq += dot4(w[lane + 32 * i], w[lane + 32 * i]);
  }
  if (a.remove_parallel) {
    dot = warp_sum(dot);
    nsq = warp_sum(nsq);
    const float sc = (nsq > 0.f) ? dot / nsq : 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) fma4(gd[i], -sc, w[lane + 32 * i]);
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < D4) __stcs(reinterpret_cast<float4*>(gdrow + 4 * v), gd[i]);
    ss += dot4(gd[i], gd[i]);
  }
  if (a.row_gsq != nullptr) {
    ss = warp_sum(ss);
    if (lane == 0) a.row_gsq[j] = ss + sdh * sdh;
  }
  if (lane == 0) a.gb_enc[j] = sdh;
}

template <int VPL>
__global__ void __launch_bounds__(32, 20) wgrad_enc_kernel(WgradArgs a, const float* __restrict__ dh_in) {
  const int lane = threadIdx.x;
  const int j = a.row_begin + blockIdx.x;
  if (j >= a.row_end) return;
  const int D4 = a.D >> 2;
  const int beg = a.feat_off[j], end = a.feat_off[j + 1];
  if (a.skip_toks != nullptr && beg == end && a.skip_toks[j] >= a.skip_threshold) return;
  float* gerow = a.gW_enc_t + static_cast<long long>(j) * a.D;
  if ((a.heavy_list != nullptr && end - beg > WGRAD_HEAVY_ENTRIES) || beg == end) {
    for (int v = lane; v < D4; v += 32) __stcs(reinterpret_cast<float4*>(gerow + 4 * v), make_float4(0, 0, 0, 0));
    return;
  }
  float4 ge[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) ge[i] = make_float4(0, 0, 0, 0);
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    int mb = 0;
    float md = 0.f;
    if (e < end) {
      const int p = a.entries[e];
      mb = p / a.K;
      md = dh_in[p];
    }
    const int cnt = min(32, end - e0);
    for (int t = 0; t < cnt; t += 2) {  // two x rows in flight
      const int b0 = __shfl_sync(FULL, mb, t), b1 = __shfl_sync(FULL, mb, min(t + 1, cnt - 1));
      const float d0 = __shfl_sync(FULL, md, t), d1 = (t + 1 < cnt) ? __shfl_sync(FULL, md, t + 1) : 0.f;
      const float* x0 = a.x + static_cast<long long>(b0) * a.D;
      const float* x1 = a.x + static_cast<long long>(b1) * a.D;
      float4 v0[VPL], v1[VPL];
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        v0[i] = (v < D4) ? ldg4(x0 + 4 * v) : make_float4(0, 0, 0, 0);
        v1[i] = (v < D4) ? ldg4(x1 + 4 * v) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        fma4(ge[i], d0, v0[i]);
        fma4(ge[i], d1, v1[i]);
      }
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (v < D4) __stcs(reinterpret_cast<float4*>(gerow + 4 * v), ge[i]);
    ss += dot4(ge[i], ge[i]);
  }
  if (a.row_gsq != nullptr) {
    ss = warp_sum(ss);
    if (lane == 0) a.row_gsq[j] += ss;  // (wgrad_dec_kernel wrote this atom's decoder share)
  }
}

static int launch_wgrad_split(const WgradArgs& a, cudaStream_t s) {
  if (a.D % 4 || a.dh_scratch == nullptr) return 21;
  const int rows = a.row_end - a.row_begin;
  if (rows <= 0) return 0;
  SB_DISPATCH_VPL(a.D, (wgrad_dec_kernel<VPL><<<rows, 32, 0, s>>>(a, a.dh_scratch)));
  SB_DISPATCH_VPL(a.D, (wgrad_enc_kernel<VPL><<<rows, 32, 0, s>>>(a, a.dh_scratch)));
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------

int launch_wgrad(const WgradArgs& a, cudaStream_t s) {
  // the two-pass form applies where dh is computed in the weight-gradient kernel (a.dh == nullptr)
  if (int rc = (a.split && a.dh == nullptr && a.dh_scratch != nullptr) ? launch_wgrad_split(a, s) : launch_wgrad_light(a, s))
    return rc;
  if (a.heavy_list == nullptr || a.row_end <= a.row_begin) return 0;
  const int grid = 148 * 2;  // every block walks the (device-side) heavy list and takes its slices; exits at once if empty
  if (a.dh == nullptr) { SB_DISPATCH_VPL(a.D, (wgrad_heavy_kernel<VPL, true><<<grid, 256, 0, s>>>(a))); }
  else { SB_DISPATCH_VPL(a.D, (wgrad_heavy_kernel<VPL, false><<<grid, 256, 0, s>>>(a))); }
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// out[d] (+)= scale * sum_b src[b,d]     (gb_dec; two deterministic stages)
// ------------------------------------------------------------------------------------------------
constexpr int COLSUM_ROWS_PER_BLOCK = 16;
int colsum_partial_rows(int B) { return (B + COLSUM_ROWS_PER_BLOCK - 1) / COLSUM_ROWS_PER_BLOCK; }

// `gate` (optional, accumulate mode only): device flag, 0 = nothing to add (AuxK with no dead latents)
__global__ void __launch_bounds__(256) colsum_stage1(const float* __restrict__ src, int B, int D, long long row_stride,
                                                     float* __restrict__ partial, const int* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int r0 = blockIdx.x * COLSUM_ROWS_PER_BLOCK;
  const int r1 = min(B, r0 + COLSUM_ROWS_PER_BLOCK);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += __ldg(src + static_cast<long long>(r) * row_stride + d);
    partial[static_cast<long long>(blockIdx.x) * D + d] = s;
  }
}
// 32 columns per block; warp w sums the partial rows w, w + 8, ... in fp64, the 8 warp sums are added in a fixed order
__global__ void __launch_bounds__(256) colsum_stage2(const float* __restrict__ partial, int P, int D, float scale,
                                                     int accumulate, float* __restrict__ out,
                                                     const int* __restrict__ gate) {
  __shared__ double ws[8][32];
  if (gate != nullptr && *gate == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (d < D)
    for (int p = warp; p < P; p += 8) s += partial[static_cast<long long>(p) * D + d];
  ws[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && d < D) {
    double tot = ws[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) tot += ws[w][lane];
    const float r = static_cast<float>(tot) * scale;
    out[d] = accumulate ? out[d] + r : r;
  }
}

int launch_colsum(const float* src, int B, int D, float scale, int accumulate, float* partial, float* out,
                  cudaStream_t s, long long row_stride, const int* gate) {
  const int P = colsum_partial_rows(B);
  if (gate != nullptr && !accumulate) return 23;  // a gated launch may leave `out` untouched: only valid when adding
  colsum_stage1<<<P, 256, 0, s>>>(src, B, D, row_stride > 0 ? row_stride : D, partial, gate);
  ++g_launch_count;
  colsum_stage2<<<(D + 31) / 32, 256, 0, s>>>(partial, P, D, scale, accumulate, out, gate);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// sum of squares of a flat fp32 buffer (global gradient norm, saev train.py:358-360), deterministic
// ------------------------------------------------------------------------------------------------
constexpr int SUMSQ_BLOCKS = 592;  // 4 x 148
__global__ void __launch_bounds__(256) sumsq_stage1(const float* __restrict__ g, long long n, double* partial) {
  __shared__ double ws[8];
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  double dacc = 0.0;
  int since = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldg4(g + 4 * i);
    acc += dot4(v, v);
    if (++since == 64) {  // bound fp32 accumulation length
      dacc += acc;
      acc = 0.f;
      since = 0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) acc += g[i] * g[i];
  dacc += acc;
  dacc = warp_sum(dacc);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void sumsq_stage2(const double* partial, int n, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += partial[i];
    *out = static_cast<float>(t);
  }
}
// ||g||^2 from the per-atom partials the weight-gradient kernels left behind, plus the b_dec gradient
__global__ void __launch_bounds__(1024) sumsq_fused_kernel(const float* __restrict__ row_gsq, int S,
                                                           const float* __restrict__ gb_dec, int D, float* __restrict__ out) {
  __shared__ double ws[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < S; i += 1024) acc += row_gsq[i];
  for (int d = threadIdx.x; d < D; d += 1024) acc += static_cast<double>(gb_dec[d]) * gb_dec[d];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += ws[w];
    *out = static_cast<float>(t);
  }
}
int launch_sumsq_fused(const float* row_gsq, int S, const float* gb_dec, int D, float* out_sumsq, cudaStream_t s) {
  sumsq_fused_kernel<<<1, 1024, 0, s>>>(row_gsq, S, gb_dec, D, out_sumsq);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

int launch_sumsq_ranges(const float* g, int n_ranges, const long long* begins, const long long* ends, double* partial,
                        float* out_sumsq, cudaStream_t s) {
  if (n_ranges < 1 || n_ranges > SUMSQ_MAX_RANGES) return 21;
  for (int r = 0; r < n_ranges; ++r) {
    sumsq_stage1<<<SUMSQ_BLOCKS, 256, 0, s>>>(g + begins[r], ends[r] - begins[r], partial + r * SUMSQ_BLOCKS);
    ++g_launch_count;
  }
  sumsq_stage2<<<1, 32, 0, s>>>(partial, n_ranges * SUMSQ_BLOCKS, out_sumsq);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

int launch_sumsq(const float* g, long long n, double* partial, float* out_sumsq, cudaStream_t s) {
  sumsq_stage1<<<SUMSQ_BLOCKS, 256, 0, s>>>(g, n, partial);
  ++g_launch_count;
  sumsq_stage2<<<1, 32, 0, s>>>(partial, SUMSQ_BLOCKS, out_sumsq);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// clip-scale + Adam + decoder row renorm + fp16 screen copy of W_enc_t, one warp per dictionary atom.
//   c = min(1, max_norm / (grad_scale*||g|| + 1e-6))                 (saev train.py:358-360)
//   m += (g-m)(1-b1); v = b2 v + (1-b2) g^2; p -= (lr/bc1) m / (sqrt(v)/sqrt(bc2) + eps)
//                                                                    (torch Adam(fused=True), train.py:294,444-446)
//   W_dec[j] /= ||W_dec[j]||  when renorm_w_dec                      (modeling.py:411-417, hoisted from the
//                                                                     start of the next step: nothing reads
//                                                                     W_dec in between)
// ------------------------------------------------------------------------------------------------
struct AdamScalars {
  float lr_over_bc1, beta1, beta2, eps, inv_bc2_sqrt, gmul;
};
__device__ __forceinline__ AdamScalars adam_scalars(const AdamArgs& a) {
  AdamScalars s;
  const float gn = a.grad_scale * sqrtf(*a.gnorm_sq);
  const float c = fminf(1.f, a.max_norm / (gn + 1e-6f));
  s.gmul = (a.max_norm > 0.f ? c : 1.f) * a.grad_scale;
  s.lr_over_bc1 = a.lr / a.bc1;
  s.beta1 = a.beta1;
  s.beta2 = a.beta2;
  s.eps = a.eps;
  s.inv_bc2_sqrt = 1.f / a.bc2_sqrt;
  return s;
}
__device__ __forceinline__ float adam_1(float p, float g, float& m, float& v, const AdamScalars& s) {
  g *= s.gmul;
  m = m + (g - m) * (1.f - s.beta1);
  v = v * s.beta2 + (1.f - s.beta2) * g * g;
  const float denom = sqrtf(v) * s.inv_bc2_sqrt + s.eps;
  return p - s.lr_over_bc1 * (m / denom);
}
__device__ __forceinline__ float4 adam_4(float4 p, float4 g, float4& m, float4& v, const AdamScalars& s) {
  float4 o;
  o.x = adam_1(p.x, g.x, m.x, v.x, s);
  o.y = adam_1(p.y, g.y, m.y, v.y, s);
  o.z = adam_1(p.z, g.z, m.z, v.z, s);
  o.w = adam_1(p.w, g.w, m.w, v.w, s);
  return o;
}

template <int VPL>
__global__ void __launch_bounds__(256) adam_rows_kernel(AdamArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = a.row_begin + blockIdx.x * (blockDim.x >> 5) + warp;
  if (j >= a.row_end) return;
  const AdamScalars sc = adam_scalars(a);
  const int D4 = a.D >> 2;
  const long long SD = static_cast<long long>(a.S) * a.D;
  const long long ro = static_cast<long long>(j) * a.D;
  // ---- W_enc_t row (+ fp16 screen copy, + max row norm for the screen's error bound) ----
  if (a.parts & 1) {
    float* mrow = a.m + ro;
    float* vrow = a.v + ro;
    float ssq = 0.f, ssq16 = 0.f, ssd = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v4 = lane + 32 * i;
      if (v4 < D4) {
        const float4 g = ldg4(a.gW_enc_t + ro + 4 * v4);
        float4 m = *reinterpret_cast<const float4*>(mrow + 4 * v4);
        float4 v = *reinterpret_cast<const float4*>(vrow + 4 * v4);
        const float4 p = adam_4(*reinterpret_cast<const float4*>(a.W_enc_t + ro + 4 * v4), g, m, v, sc);
        *reinterpret_cast<float4*>(a.W_enc_t + ro + 4 * v4) = p;
        *reinterpret_cast<float4*>(mrow + 4 * v4) = m;
        *reinterpret_cast<float4*>(vrow + 4 * v4) = v;
        accum_fp16_stats(p, ssq, ssq16, ssd);
        if (a.shadow16 != nullptr) {
          __half2* so = reinterpret_cast<__half2*>(a.shadow16 + ro + 4 * v4);
          so[0] = __floats2half2_rn(p.x, p.y);
          so[1] = __floats2half2_rn(p.z, p.w);
        }
      }
    }
    if (a.wnorm_sq_max != nullptr) {
      ssq = warp_sum(ssq);
      ssq16 = warp_sum(ssq16);
      ssd = warp_sum(ssd);
      if (lane == 0) {
        float c, ratio;
        screen_col_stats(ssq, ssq16, ssd, c, ratio);
        atomicMax(reinterpret_cast<int*>(a.wnorm_sq_max), __float_as_int(ssq));
        if (a.wnorm_rows != nullptr) a.wnorm_rows[j] = c;
        if (a.rho != nullptr) atomicMax(reinterpret_cast<int*>(a.rho), __float_as_int(ratio));
      }
    }
  }
  // ---- b_enc[j] (a sharded optimizer updates the whole bias vector on every rank instead) ----
  if (lane == 0 && !a.b_enc_separately && (a.parts & 1)) {
    float m = a.m[SD + j], v = a.v[SD + j];
    const float nb = adam_1(a.b_enc[j], a.gb_enc[j], m, v, sc);
    a.b_enc[j] = nb;
    a.m[SD + j] = m;
    a.v[SD + j] = v;
    if (a.bias_abs_max != nullptr) atomicMax(reinterpret_cast<int*>(a.bias_abs_max), __float_as_int(fabsf(nb)));
  }
  // ---- W_dec row (+ renorm) ----
  if (a.parts & 2) {
    float* mrow = a.m + SD + a.S + ro;
    float* vrow = a.v + SD + a.S + ro;
    float4 p[VPL];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v4 = lane + 32 * i;
      p[i] = make_float4(0, 0, 0, 0);
      if (v4 < D4) {
        const float4 g = ldg4(a.gW_dec + ro + 4 * v4);
        float4 m = *reinterpret_cast<const float4*>(mrow + 4 * v4);
        float4 v = *reinterpret_cast<const float4*>(vrow + 4 * v4);
        p[i] = adam_4(*reinterpret_cast<const float4*>(a.W_dec + ro + 4 * v4), g, m, v, sc);
        *reinterpret_cast<float4*>(mrow + 4 * v4) = m;
        *reinterpret_cast<float4*>(vrow + 4 * v4) = v;
        ss += dot4(p[i], p[i]);
      }
    }
    if (a.renorm_w_dec) {
      ss = warp_sum(ss);
      const float nrm = sqrtf(ss);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        p[i].x /= nrm; p[i].y /= nrm; p[i].z /= nrm; p[i].w /= nrm;
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v4 = lane + 32 * i;
      if (v4 < D4) *reinterpret_cast<float4*>(a.W_dec + ro + 4 * v4) = p[i];
    }
  }
}

// plain vector Adam over b_enc[S] (sharded optimizer: the bias gradients are all-reduced, not scattered)
__global__ void adam_benc_kernel(AdamArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.S) return;
  const AdamScalars sc = adam_scalars(a);
  const long long o = static_cast<long long>(a.S) * a.D + j;
  float m = a.m[o], v = a.v[o];
  const float nb = adam_1(a.b_enc[j], a.gb_enc[j], m, v, sc);
  a.b_enc[j] = nb;
  a.m[o] = m;
  a.v[o] = v;
  if (a.bias_abs_max != nullptr) atomicMax(reinterpret_cast<int*>(a.bias_abs_max), __float_as_int(fabsf(nb)));
}

__global__ void adam_bdec_kernel(AdamArgs a) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const AdamScalars sc = adam_scalars(a);
  if (d == 0 && a.gnorm_out != nullptr) *a.gnorm_out = a.grad_scale * sqrtf(*a.gnorm_sq);
  if (d >= a.D) return;
  const long long o = 2LL * a.S * a.D + a.S + d;
  float m = a.m[o], v = a.v[o];
  a.b_dec[d] = adam_1(a.b_dec[d], a.gb_dec[d], m, v, sc);
  a.m[o] = m;
  a.v[o] = v;
}

int launch_adam(const AdamArgs& a, cudaStream_t s) {
  if (a.D % 4 || a.S % 4) return 21;
  if ((a.parts & 1) && !(a.parts & 8)) {
    if (a.wnorm_sq_max != nullptr && cudaMemsetAsync(a.wnorm_sq_max, 0, 4, s) != cudaSuccess) return 23;
    if (a.bias_abs_max != nullptr && cudaMemsetAsync(a.bias_abs_max, 0, 4, s) != cudaSuccess) return 23;
    if (a.rho != nullptr && cudaMemsetAsync(a.rho, 0, 4, s) != cudaSuccess) return 23;
  }
  const int rows = a.row_end - a.row_begin;
  const int wpb = a.small_blocks ? 2 : 8;
  if (rows > 0) SB_DISPATCH_VPL(a.D, (adam_rows_kernel<VPL><<<(rows + wpb - 1) / wpb, 32 * wpb, 0, s>>>(a)));
  if (a.b_enc_separately && (a.parts & 1) && !(a.parts & 4)) {
    adam_benc_kernel<<<(a.S + 255) / 256, 256, 0, s>>>(a);
    ++g_launch_count;
  }
  if ((a.parts & 2) && !(a.parts & 4)) {
    adam_bdec_kernel<<<(a.D + 255) / 256, 256, 0, s>>>(a);
    ++g_launch_count;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// loss scalars (saev objectives.py:133-156): mse, aux, sparsity, l0, l1, n_dead, loss
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) finalize_kernel(FinalizeArgs a) {
  __shared__ double s0[32], s1[32], s2[32];
  double sse = 0.0, l1 = 0.0, l0 = 0.0;
  for (int b = threadIdx.x; b < a.B; b += blockDim.x) {
    sse += a.row_sse[b];
    l1 += a.row_l1[b];
    l0 += a.row_l0[b];
  }
  sse = warp_sum(sse);
  l1 = warp_sum(l1);
  l0 = warp_sum(l0);
  if ((threadIdx.x & 31) == 0) {
    s0[threadIdx.x >> 5] = sse;
    s1[threadIdx.x >> 5] = l1;
    s2[threadIdx.x >> 5] = l0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sse = l1 = l0 = 0.0;
    for (int w = 0; w < 32; ++w) {
      sse += s0[w];
      l1 += s1[w];
      l0 += s2[w];
    }
    const float mse = static_cast<float>(sse * a.inv_bd);
    const float l1m = static_cast<float>(l1 * a.inv_b);
    const float aux = a.aux_loss ? *a.aux_loss : 0.f;
    const float sp = a.l1_coeff * l1m;
    a.losses[0] = mse;
    a.losses[1] = aux;
    a.losses[2] = sp;
    a.losses[3] = static_cast<float>(l0 * a.inv_b);
    a.losses[4] = l1m;
    a.losses[5] = a.n_dead ? static_cast<float>(*a.n_dead) : 0.f;
    a.losses[6] = mse + sp + aux;
    a.losses[7] = 0.f;
  }
}
int launch_finalize(const FinalizeArgs& a, cudaStream_t s) {
  finalize_kernel<<<1, 1024, 0, s>>>(a);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// dead-latent tracker (saev objectives.py:107-120): toks += tokens; toks[active] = 0;
// dead = toks >= threshold; emits the ascending list of dead atoms and its length; clears `active`.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) dead_update_kernel(long long* __restrict__ toks, int* __restrict__ active, int S,
                                                           long long batch_tokens, long long threshold,
                                                           int* __restrict__ totals) {
  __shared__ int ws[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  int dead = 0;
  if (i < S) {
    const long long t = active[i] ? 0 : toks[i] + batch_tokens;
    toks[i] = t;  // (the flags stay valid until the next forward re-zeroes them: the log block reads them)
    dead = t >= threshold;
  }
  dead = warp_sum(dead);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = dead;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = warp_sum(ws[threadIdx.x]);
    if (threadIdx.x == 0) totals[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) dead_list_kernel(const long long* __restrict__ toks, int S, long long threshold,
                                                         const int* __restrict__ totals, int* __restrict__ dead_list,
                                                         int* __restrict__ n_dead) {
  __shared__ int ws[32];
  const int prefix = block_prefix(totals, ws);
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int dead = (i < S) && toks[i] >= threshold;
  int tot;
  const int incl = block_scan_incl(dead, ws, &tot);
  if (dead) dead_list[prefix + incl - 1] = i;  // ascending atom order
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_dead = prefix + tot;
}
int launch_dead_update(long long* toks, int* active, int S, long long batch_tokens, long long threshold,
                       int* dead_list, int* n_dead, int* block_totals, cudaStream_t s) {
  const int nb = (S + 1023) / 1024;
  dead_update_kernel<<<nb, 1024, 0, s>>>(toks, active, S, batch_tokens, threshold, block_totals);
  ++g_launch_count;
  dead_list_kernel<<<nb, 1024, 0, s>>>(toks, S, threshold, block_totals, dead_list, n_dead);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

// ------------------------------------------------------------------------------------------------
// lazy materialisation of the dense tensors saev's logging block reads (train.py:365-442)
// ------------------------------------------------------------------------------------------------
__global__ void densify_kernel(const int* __restrict__ idx, const float* __restrict__ val, long long n, int K, int S,
                               float* __restrict__ out) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int j = idx[p];
  if (j >= 0) out[(p / K) * S + j] = val[p];
}
int launch_densify(const int* idx, const float* val, int B, int K, int S, float* out, cudaStream_t s) {
  if (cudaMemsetAsync(out, 0, static_cast<size_t>(B) * S * 4, s) != cudaSuccess) return 23;
  const long long n = static_cast<long long>(B) * K;
  densify_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, s>>>(idx, val, n, K, S, out);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                           float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = a[i] + b[i];
}
int launch_add_rows(const float* a, const float* b, long long n, float* out, cudaStream_t s) {
  add_kernel<<<148 * 8, 256, 0, s>>>(a, b, n, out);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

}  // namespace sb
