// AuxK dead-latent loss (saev src/saev/nn/modeling.py:75-103) and its gradients, restricted to the
// compact list L of dead dictionary atoms the dead tracker produced (device-side length n_dead; the host
// never reads it, so every kernel here takes its problem size from device memory and runs a fixed grid).
//
//   h_L   = x . W_enc_t[L]^T + b_enc[L]                  pre-activations of the dead atoms       [B, n_dead]
//   f_aux = top-k_use of h_L per row, k_use = min(k_aux, n_dead)   (raw pre-acts, may be negative)
//   x_aux = f_aux . W_dec[L] + b_dec ;  r_aux = x_aux - e,  e = x - x_hat = -resid
//   aux   = alpha * mean(r_aux^2)
//   G_a = 2 alpha r_aux / (B D);  gW_dec[L] = f_aux^T G_a (minus parallel part);  gb_dec += sum_b G_a
//   dh_a = mask_a * (G_a . W_dec[L]^T);  gW_enc_t[L] = dh_a^T x;  gb_enc[L] = sum_b dh_a
// Dead atoms never fired in this batch, so their rows of the main-path gradients are zero and are simply
// overwritten.  Arithmetic is plain fp32 on CUDA cores (exact w.r.t. the reference's fp32): this path only
// runs once latents have died and its contractions are n_dead wide, not d_sae wide.
#include "common.cuh"
#include "kernels.h"

namespace sb {

constexpr int TM = 64, TN = 64, TK = 16, TPAD = 4;
constexpr int SGEMM_GRID = 148 * 3;

struct SgemmArgs {
  const float* A; long long lda;
  const float* B; long long ldb;
  float* C; long long ldc;
  int M, N, K;
  const int* dyn; int dyn_which;   // 0 none, 1: M = min(M,*dyn), 2: N, 3: K
  const int* gatherM;              // row index of C
  const int* gatherB;              // row index of B (N index when B_T, K index otherwise)
  float alpha;
};

// C[gm(m), n] = alpha * sum_k A(m,k) B(k,n)
//   A(m,k) = A_T ? A[k*lda + m] : A[m*lda + k]
//   B(k,n) = B_T ? B[gb(n)*ldb + k] : B[gb(k)*ldb + n]
template <bool A_T, bool B_T>
__global__ void __launch_bounds__(256) sgemm_kernel(SgemmArgs a) {
  __shared__ __align__(16) float As[TK][TM + TPAD];
  __shared__ __align__(16) float Bs[TK][TN + TPAD];
  int M = a.M, N = a.N, K = a.K;
  if (a.dyn_which) {
    const int d = *a.dyn;
    if (a.dyn_which == 1) M = min(M, d);
    if (a.dyn_which == 2) N = min(N, d);
    if (a.dyn_which == 3) K = min(K, d);
    // empty contraction (no dead latents): the only such call decodes into r_aux, which nobody reads in that case
    if (a.dyn_which == 3 && K == 0) return;
  }
  const int tiles_m = (M + TM - 1) / TM, tiles_n = (N + TN - 1) / TN;
  const long long tiles = static_cast<long long>(tiles_m) * tiles_n;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const bool veca = (a.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0;
  const bool vecb = (a.ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int m0 = static_cast<int>(tile / tiles_n) * TM, n0 = static_cast<int>(tile % tiles_n) * TN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += TK) {
      // ---- A tile -> As[k][m] ----
      if (!A_T) {
        const int m = t >> 2, kq = (t & 3) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m0 + m < M) {
          const float* p = a.A + static_cast<long long>(m0 + m) * a.lda + k0 + kq;
          if (veca && k0 + kq + 3 < K) {
            const float4 q = *reinterpret_cast<const float4*>(p);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (k0 + kq + i < K) v[i] = p[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) As[kq + i][m] = v[i];
      } else {
        const int k = t >> 4, mq = (t & 15) * 4;
        float4 q = make_float4(0, 0, 0, 0);
        if (k0 + k < K) {
          const float* p = a.A + static_cast<long long>(k0 + k) * a.lda + m0 + mq;
          if (veca && m0 + mq + 3 < M) {
            q = *reinterpret_cast<const float4*>(p);
          } else {
            if (m0 + mq + 0 < M) q.x = p[0];
            if (m0 + mq + 1 < M) q.y = p[1];
            if (m0 + mq + 2 < M) q.z = p[2];
            if (m0 + mq + 3 < M) q.w = p[3];
          }
        }
        *reinterpret_cast<float4*>(&As[k][mq]) = q;
      }
      // ---- B tile -> Bs[k][n] ----
      if (B_T) {
        const int n = t >> 2, kq = (t & 3) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (n0 + n < N) {
          const long long r = a.gatherB ? a.gatherB[n0 + n] : (n0 + n);
          const float* p = a.B + r * a.ldb + k0 + kq;
          if (vecb && k0 + kq + 3 < K) {
            const float4 q = *reinterpret_cast<const float4*>(p);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (k0 + kq + i < K) v[i] = p[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) Bs[kq + i][n] = v[i];
      } else {
        const int k = t >> 4, nq = (t & 15) * 4;
        float4 q = make_float4(0, 0, 0, 0);
        if (k0 + k < K) {
          const long long r = a.gatherB ? a.gatherB[k0 + k] : (k0 + k);
          const float* p = a.B + r * a.ldb + n0 + nq;
          if (vecb && n0 + nq + 3 < N) {
            q = *reinterpret_cast<const float4*>(p);
          } else {
            if (n0 + nq + 0 < N) q.x = p[0];
            if (n0 + nq + 1 < N) q.y = p[1];
            if (n0 + nq + 2 < N) q.z = p[2];
            if (n0 + nq + 3 < N) q.w = p[3];
          }
        }
        *reinterpret_cast<float4*>(&Bs[k][nq]) = q;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float am[4] = {av.x, av.y, av.z, av.w};
        const float bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m >= M) continue;
      const long long cm = a.gatherM ? a.gatherM[m] : m;
      float* crow = a.C + cm * a.ldc;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < N) crow[n] = a.alpha * acc[i][j];
      }
    }
  }
}

template <bool A_T, bool B_T>
static void launch_sgemm(const SgemmArgs& a, cudaStream_t s) {
  sgemm_kernel<A_T, B_T><<<SGEMM_GRID, 256, 0, s>>>(a);
  ++g_launch_count;
}

// ------------------------------------------------------------------------------------------------
// per-row top-k_use among the n_dead pre-activations: radix select on order-preserving keys
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) aux_select_kernel(float* __restrict__ h_aux, unsigned char* __restrict__ mask,
                                                         long long ld, const float* __restrict__ b_enc,
                                                         const int* __restrict__ dead_list,
                                                         const int* __restrict__ n_dead_p, int B, int k_aux) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sh_prefix, sh_need;
  __shared__ int wtot[8];
  __shared__ int sh_carry;
  const int n = *n_dead_p;
  if (n <= 0) return;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    float* row = h_aux + static_cast<long long>(b) * ld;
    unsigned char* mrow = mask + static_cast<long long>(b) * ld;
    // add the bias (h = x.W + b_enc, modeling.py:344-347)
    for (int i = t; i < n; i += 256) row[i] += __ldg(b_enc + dead_list[i]);
    __syncthreads();
    if (n <= k_aux) {
      for (int i = t; i < n; i += 256) mrow[i] = 1;
      __syncthreads();
      continue;
    }
    // radix select of the k_aux-th largest key, 8 bits per pass from the top
    if (t == 0) {
      sh_prefix = 0;
      sh_need = static_cast<unsigned int>(k_aux);
    }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      hist[t] = 0;
      __syncthreads();
      const unsigned int prefix = sh_prefix;
      const unsigned int himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = t; i < n; i += 256) {
        const unsigned int key = fkey(row[i]);
        if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (t == 0) {
        unsigned int need = sh_need, cum = 0;
        int bin = 255;
        for (; bin > 0; --bin) {
          if (cum + hist[bin] >= need) break;
          cum += hist[bin];
        }
        sh_need = need - cum;  // how many to take inside the chosen bin
        sh_prefix = prefix | (static_cast<unsigned int>(bin) << shift);
      }
      __syncthreads();
    }
    const unsigned int T = sh_prefix;    // key of the k-th largest value
    const int need_eq = static_cast<int>(sh_need);  // ties at T to keep, lowest column first
    if (t == 0) sh_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
      const int i = base + t;
      unsigned int key = 0;
      float v = 0.f;
      if (i < n) {
        v = row[i];
        key = fkey(v);
      }
      const int eq = (i < n) && key == T;
      const unsigned bal = __ballot_sync(FULL, eq);
      const int incl = __popc(bal & (0xffffffffu >> (31 - lane)));
      if (lane == 31) wtot[warp] = incl;
      __syncthreads();
      int before = sh_carry;
      for (int w = 0; w < warp; ++w) before += wtot[w];
      const int rank_eq = before + incl - 1;  // 0-based rank among ties, in column order
      if (i < n) {
        const bool sel = key > T || (eq && rank_eq < need_eq);
        mrow[i] = sel ? 1 : 0;
        if (!sel) row[i] = 0.f;
      }
      __syncthreads();
      if (t == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += wtot[w];
        sh_carry += tot;
      }
      __syncthreads();
    }
  }
}

// r_aux[b,:] += b_dec + resid[b,:]  (= x_aux - e);  row_sse_aux[b] = sum r_aux^2.   n_dead == 0 -> zeros.
__global__ void __launch_bounds__(256) aux_resid_kernel(float* __restrict__ r_aux, const float* __restrict__ resid,
                                                        const float* __restrict__ b_dec, int B, int D,
                                                        const int* __restrict__ n_dead_p,
                                                        float* __restrict__ row_sse_aux) {
  const int n = *n_dead_p;
  const int lane = threadIdx.x & 31;
  if (n == 0) {  // aux = 0 (modeling.py:92-94); r_aux is not read by the backward either (gated on n_dead)
    for (int b = blockIdx.x * 256 + threadIdx.x; b < B; b += gridDim.x * 256) row_sse_aux[b] = 0.f;
    return;
  }
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < B; b += gridDim.x * 8) {
    float* rr = r_aux + static_cast<long long>(b) * D;
    const float* r0 = resid + static_cast<long long>(b) * D;
    float sse = 0.f;
    for (int d = lane; d < D; d += 32) {
      float v = 0.f;
      if (n > 0) v = rr[d] + b_dec[d] + r0[d];
      rr[d] = v;
      sse += v * v;
    }
    sse = warp_sum(sse);
    if (lane == 0) row_sse_aux[b] = sse;
  }
}

__global__ void __launch_bounds__(1024) aux_loss_kernel(const float* __restrict__ row_sse_aux, int B, float alpha,
                                                        float inv_bd, float* __restrict__ aux_loss) {
  __shared__ double ws[32];
  double s = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += row_sse_aux[b];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 32; ++w) tot += ws[w];
    *aux_loss = static_cast<float>(static_cast<double>(alpha) * tot * static_cast<double>(inv_bd));
  }
}

// remove the component of gW_dec[L[i]] parallel to W_dec[L[i]]  (modeling.py:419-445)
__global__ void __launch_bounds__(256) aux_project_kernel(float* __restrict__ gW_dec, const float* __restrict__ W_dec,
                                                          const int* __restrict__ dead_list,
                                                          const int* __restrict__ n_dead_p, int D) {
  const int n = *n_dead_p;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const long long j = dead_list[i];
    float* g = gW_dec + j * D;
    const float* w = W_dec + j * D;
    float dot = 0.f, nsq = 0.f;
    for (int d = lane; d < D; d += 32) {
      dot = fmaf(g[d], w[d], dot);
      nsq = fmaf(w[d], w[d], nsq);
    }
    dot = warp_sum(dot);
    nsq = warp_sum(nsq);
    const float sc = nsq > 0.f ? dot / nsq : 0.f;
    for (int d = lane; d < D; d += 32) g[d] = fmaf(-sc, w[d], g[d]);
  }
}

// dh_aux *= mask ; partial column sums over a slab of rows
constexpr int AUX_SLABS = 32;
__global__ void __launch_bounds__(256) aux_mask_colsum_kernel(float* __restrict__ dh, const unsigned char* __restrict__ mask,
                                                              long long ld, int B, const int* __restrict__ n_dead_p,
                                                              float* __restrict__ partial /* [AUX_SLABS, ld] */) {
  const int n = *n_dead_p;
  const int slab = blockIdx.y;
  const int rows = (B + AUX_SLABS - 1) / AUX_SLABS;
  const int b0 = slab * rows, b1 = min(B, b0 + rows);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = b0; b < b1; ++b) {
      const long long o = static_cast<long long>(b) * ld + i;
      const float v = mask[o] ? dh[o] : 0.f;
      dh[o] = v;
      s += v;
    }
    partial[static_cast<long long>(slab) * ld + i] = s;
  }
}
__global__ void __launch_bounds__(256) aux_colsum_final_kernel(const float* __restrict__ partial, long long ld,
                                                               const int* __restrict__ dead_list,
                                                               const int* __restrict__ n_dead_p,
                                                               float* __restrict__ gb_enc) {
  const int n = *n_dead_p;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < AUX_SLABS; ++p) s += partial[static_cast<long long>(p) * ld + i];
    gb_enc[dead_list[i]] = s;
  }
}

// row_gsq[L[i]] = ||gW_dec[L[i]]||^2 + ||gW_enc_t[L[i]]||^2 + gb_enc[L[i]]^2 for the dead atoms (their rows were rewritten)
__global__ void __launch_bounds__(256) aux_rows_gsq_kernel(const float* __restrict__ gW_dec, const float* __restrict__ gW_enc_t,
                                                           const float* __restrict__ gb_enc, const int* __restrict__ dead_list,
                                                           const int* __restrict__ n_dead_p, int D, float* __restrict__ row_gsq) {
  const int n = *n_dead_p;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const long long j = dead_list[i];
    const float* a = gW_dec + j * D;
    const float* b = gW_enc_t + j * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) ss += a[d] * a[d] + b[d] * b[d];
    ss = warp_sum(ss);
    if (lane == 0) row_gsq[j] = ss + gb_enc[j] * gb_enc[j];
  }
}

// ------------------------------------------------------------------------------------------------
// Operand preparation for the tensor-core AuxK path: bf16 (hi, lo, lo2) pieces of everything the five n_dead-wide
// contractions read, in K-major layout.  Sizes come from device memory (n_dead), grids are fixed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split3(float v, __nv_bfloat16& p0, __nv_bfloat16& p1, __nv_bfloat16& p2) {
  p0 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(p0);
  p1 = __float2bfloat16_rn(r1);
  p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
}

// rows of W[S, D] listed in dead_list -> pieces [cap, D] (row i = atom dead_list[i])
__global__ void __launch_bounds__(256) aux_gather_rows_kernel(const float* __restrict__ W, const int* __restrict__ dead_list,
                                                              const int* __restrict__ n_dead_p, int D,
                                                              __nv_bfloat16* __restrict__ p0, __nv_bfloat16* __restrict__ p1,
                                                              __nv_bfloat16* __restrict__ p2) {
  const int n = *n_dead_p;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const float* src = W + static_cast<long long>(dead_list[i]) * D;
    const long long o = static_cast<long long>(i) * D;
    for (int d = lane; d < D; d += 32) {
      __nv_bfloat16 a, b, c;
      split3(__ldg(src + d), a, b, c);
      p0[o + d] = a;
      p1[o + d] = b;
      p2[o + d] = c;
    }
  }
}

// W[dead_list[i], d] -> pieces T[d, i] ([D, ldc]); columns n .. roundup64(n) are zero-filled (they sit inside the last
// k-block of a contraction over the dead latents)
__global__ void __launch_bounds__(256) aux_gather_rows_T_kernel(const float* __restrict__ W, const int* __restrict__ dead_list,
                                                                const int* __restrict__ n_dead_p, int D, long long ldc,
                                                                __nv_bfloat16* __restrict__ p0, __nv_bfloat16* __restrict__ p1,
                                                                __nv_bfloat16* __restrict__ p2) {
  __shared__ float tile[32][33];
  const int n = *n_dead_p;
  const int n_pad = (n + 63) / 64 * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int d_tiles = (D + 31) / 32;
  for (long long blk = blockIdx.x; blk < static_cast<long long>((n_pad + 31) / 32) * d_tiles; blk += gridDim.x) {
    const int i0 = static_cast<int>(blk / d_tiles) * 32, d0 = static_cast<int>(blk % d_tiles) * 32;
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + r, d = d0 + tx;
      tile[r][tx] = (i < n && d < D) ? __ldg(W + static_cast<long long>(dead_list[i]) * D + d) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int d = d0 + r, i = i0 + tx;
      if (d < D && i < n_pad) {
        __nv_bfloat16 a, b, c;
        split3(tile[tx][r], a, b, c);
        const long long o = static_cast<long long>(d) * ldc + i;
        p0[o] = a;
        p1[o] = b;
        p2[o] = c;
      }
    }
    __syncthreads();
  }
}

// src[B, ld] (fp32, first n columns valid) -> row-major pieces [B, ldc] (optional; columns n .. roundup64(n) zero)
// and transposed pieces T[i, b] ([cap, ldb], rows i < n)
__global__ void __launch_bounds__(256) aux_split_cols_kernel(const float* __restrict__ src, long long ld, int B,
                                                             const int* __restrict__ n_dead_p, long long ldc, long long ldb,
                                                             __nv_bfloat16* __restrict__ r0, __nv_bfloat16* __restrict__ r1,
                                                             __nv_bfloat16* __restrict__ r2, __nv_bfloat16* __restrict__ t0,
                                                             __nv_bfloat16* __restrict__ t1, __nv_bfloat16* __restrict__ t2) {
  __shared__ float tile[32][33];
  const int n = *n_dead_p;
  const int n_pad = (n + 63) / 64 * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c_tiles = (n_pad + 31) / 32, b_tiles = (B + 31) / 32;
  for (long long blk = blockIdx.x; blk < static_cast<long long>(c_tiles) * b_tiles; blk += gridDim.x) {
    const int b0 = static_cast<int>(blk / c_tiles) * 32, c0 = static_cast<int>(blk % c_tiles) * 32;
    for (int r = ty; r < 32; r += 8) {
      const int b = b0 + r, c = c0 + tx;
      const float val = (b < B && c < n) ? src[static_cast<long long>(b) * ld + c] : 0.f;
      tile[r][tx] = val;
      if (r0 != nullptr && b < B && c < n_pad) {
        __nv_bfloat16 p, q, s;
        split3(val, p, q, s);
        const long long o = static_cast<long long>(b) * ldc + c;
        r0[o] = p;
        r1[o] = q;
        r2[o] = s;
      }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int c = c0 + r, b = b0 + tx;
      if (c < n && b < B) {
        __nv_bfloat16 p, q, s;
        split3(tile[tx][r], p, q, s);
        const long long o = static_cast<long long>(c) * ldb + b;
        t0[o] = p;
        t1[o] = q;
        t2[o] = s;
      }
    }
    __syncthreads();
  }
}

// zero the gradient rows of the dead atoms (the K-split weight-gradient contractions add into them)
__global__ void __launch_bounds__(256) aux_zero_rows_kernel(float* __restrict__ g0, float* __restrict__ g1,
                                                            const int* __restrict__ dead_list,
                                                            const int* __restrict__ n_dead_p, int D) {
  const int n = *n_dead_p;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const long long o = static_cast<long long>(dead_list[i]) * D;
    for (int d = lane; d < D; d += 32) {
      g0[o + d] = 0.f;
      g1[o + d] = 0.f;
    }
  }
}

size_t aux_colpart_bytes(int cap) { return static_cast<size_t>(AUX_SLABS) * cap * 4; }

// ------------------------------------------------------------------------------------------------
// Tensor-core path: the five n_dead-wide contractions as bf16 split products on the tcgen05 kernel of encode_gemm.cu
// (dynamic row / column / contraction limits read from device memory).  Measured need: with the fp32 CUDA-core
// tiles 2 k dead latents already cost 16 ms per step at c3, 32 k cost 210 ms.
// ------------------------------------------------------------------------------------------------
static EncodeGemmArgs aux_gemm(const AuxArgs& a, __nv_bfloat16* const A[3], long long lda, __nv_bfloat16* const Bp[3],
                               long long ldb, int M, int N, int K, int epilogue) {
  EncodeGemmArgs g;
  g.A_hi = A[0];
  g.A_lo = A[1];
  g.A_lo2 = A[2];
  g.B_hi = Bp[0];
  g.B_lo = Bp[1];
  g.B_lo2 = Bp[2];
  g.lda = lda;
  g.ldb = ldb;
  g.nterms = a.nterms;
  g.k_chunk_blocks = 8;
  g.M = M;
  g.N = N;
  g.K = K;
  g.epilogue = epilogue;
  g.nsplit = 0;
  g.num_sms = a.num_sms;
  return g;
}

static int launch_aux_forward_tc(const AuxArgs& a, cudaStream_t s) {
  const long long ld = a.S;
  const int cap = a.S;
  aux_gather_rows_kernel<<<148 * 4, 256, 0, s>>>(a.W_enc_t, a.dead_list, a.n_dead, a.D, a.tc_we[0], a.tc_we[1], a.tc_we[2]);
  aux_gather_rows_kernel<<<148 * 4, 256, 0, s>>>(a.W_dec, a.dead_list, a.n_dead, a.D, a.tc_wd[0], a.tc_wd[1], a.tc_wd[2]);
  aux_gather_rows_T_kernel<<<148 * 8, 256, 0, s>>>(a.W_dec, a.dead_list, a.n_dead, a.D, a.ldc, a.tc_wdT[0], a.tc_wdT[1],
                                                   a.tc_wdT[2]);
  g_launch_count += 3;
  if (launch_split_bf16(a.x, a.tc_x[0], a.tc_x[1], static_cast<long long>(a.B) * a.D, s, a.tc_x[2], a.n_dead)) return 22;
  // h_L = x . W_enc_t[L]^T   (the bias is added by the selection kernel)
  EncodeGemmArgs g = aux_gemm(a, a.tc_x, a.D, a.tc_we, a.D, a.B, cap, a.D, 1);
  g.n_limit_dev = a.n_dead;
  g.out = a.h_aux;
  g.ldo = ld;
  if (launch_encode_gemm(g, s)) return 22;
  aux_select_kernel<<<min(a.B, 148 * 8), 256, 0, s>>>(a.h_aux, a.mask_aux, ld, a.b_enc, a.dead_list, a.n_dead, a.B,
                                                       a.k_aux);
  ++g_launch_count;
  // f_aux as operand pieces: row-major for the decode, transposed for gW_dec
  aux_split_cols_kernel<<<148 * 8, 256, 0, s>>>(a.h_aux, ld, a.B, a.n_dead, a.ldc, a.ldb, a.tc_f[0], a.tc_f[1], a.tc_f[2],
                                                a.tc_fT[0], a.tc_fT[1], a.tc_fT[2]);
  ++g_launch_count;
  // x_aux (without bias) = f_aux . W_dec[L]   (contraction over the dead latents: dynamic K)
  EncodeGemmArgs d = aux_gemm(a, a.tc_f, a.ldc, a.tc_wdT, a.ldc, a.B, a.D, static_cast<int>(a.ldc), 1);
  d.k_limit_dev = a.n_dead;
  d.out = a.r_aux;
  d.ldo = a.D;
  if (launch_encode_gemm(d, s)) return 22;
  aux_resid_kernel<<<min((a.B + 7) / 8, 148 * 8), 256, 0, s>>>(a.r_aux, a.resid, a.b_dec, a.B, a.D, a.n_dead,
                                                                a.row_sse_aux);
  ++g_launch_count;
  aux_loss_kernel<<<1, 1024, 0, s>>>(a.row_sse_aux, a.B, a.alpha, a.inv_bd, a.aux_loss);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

static int launch_aux_backward_tc(const AuxArgs& a, cudaStream_t s) {
  const long long ld = a.S;
  const int cap = a.S;
  const float gscale = 2.f * a.alpha * a.inv_bd;
  if (launch_colsum(a.r_aux, a.B, a.D, gscale, 1, a.colsum_partial, a.gb_dec, s, 0, a.n_dead)) return 22;
  if (launch_split_bf16(a.r_aux, a.tc_r[0], a.tc_r[1], static_cast<long long>(a.B) * a.D, s, a.tc_r[2], a.n_dead) ||
      launch_transpose_split(a.r_aux, a.B, a.D, 1.f, a.tc_rT[0], a.tc_rT[1], a.ldb, 0, a.D, s, a.tc_rT[2], a.n_dead) ||
      launch_transpose_split(a.x, a.B, a.D, 1.f, a.tc_xT[0], a.tc_xT[1], a.ldb, 0, a.D, s, a.tc_xT[2], a.n_dead))
    return 22;
  // The two weight-gradient contractions run over K = B with as few as one 128-row block of output when only a
  // handful of latents is dead: their K chunks are spread over 8 CTAs per tile, adding into zeroed rows.
  aux_zero_rows_kernel<<<148 * 2, 256, 0, s>>>(a.gW_dec, a.gW_enc_t, a.dead_list, a.n_dead, a.D);
  ++g_launch_count;
  const int n_tiles_d = (a.D + 255) / 256;
  // gW_dec[L] = gscale * f_aux^T r_aux, scattered to the rows of the dead atoms
  EncodeGemmArgs g = aux_gemm(a, a.tc_fT, a.ldb, a.tc_rT, a.ldb, cap, a.D, a.B, 4);
  g.ksplit = 8;
  g.nsplit = n_tiles_d;
  g.m_limit_dev = a.n_dead;
  g.row_map = a.dead_list;
  g.out = a.gW_dec;
  g.ldo = a.D;
  g.n_main = a.D;
  g.alpha = gscale;
  if (launch_encode_gemm(g, s)) return 22;
  if (a.remove_parallel) {
    aux_project_kernel<<<148 * 4, 256, 0, s>>>(a.gW_dec, a.W_dec, a.dead_list, a.n_dead, a.D);
    ++g_launch_count;
  }
  // dh_a = mask_a * gscale * (r_aux . W_dec[L]^T)   (overwrites f_aux)
  EncodeGemmArgs e = aux_gemm(a, a.tc_r, a.D, a.tc_wd, a.D, a.B, cap, a.D, 1);
  e.n_limit_dev = a.n_dead;
  e.out = a.h_aux;
  e.ldo = ld;
  e.alpha = gscale;
  if (launch_encode_gemm(e, s)) return 22;
  aux_mask_colsum_kernel<<<dim3(148, AUX_SLABS), 256, 0, s>>>(a.h_aux, a.mask_aux, ld, a.B, a.n_dead, a.aux_colpart);
  ++g_launch_count;
  aux_colsum_final_kernel<<<148, 256, 0, s>>>(a.aux_colpart, ld, a.dead_list, a.n_dead, a.gb_enc);
  ++g_launch_count;
  aux_split_cols_kernel<<<148 * 8, 256, 0, s>>>(a.h_aux, ld, a.B, a.n_dead, a.ldc, a.ldb, nullptr, nullptr, nullptr,
                                                a.tc_fT[0], a.tc_fT[1], a.tc_fT[2]);
  ++g_launch_count;
  // gW_enc_t[L] = dh_a^T x
  EncodeGemmArgs w = aux_gemm(a, a.tc_fT, a.ldb, a.tc_xT, a.ldb, cap, a.D, a.B, 4);
  w.ksplit = 8;
  w.nsplit = n_tiles_d;
  w.m_limit_dev = a.n_dead;
  w.row_map = a.dead_list;
  w.out = a.gW_enc_t;
  w.ldo = a.D;
  w.n_main = a.D;
  if (launch_encode_gemm(w, s)) return 22;
  if (a.row_gsq != nullptr) {
    aux_rows_gsq_kernel<<<148 * 2, 256, 0, s>>>(a.gW_dec, a.gW_enc_t, a.gb_enc, a.dead_list, a.n_dead, a.D, a.row_gsq);
    ++g_launch_count;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

int launch_aux_forward(const AuxArgs& a, cudaStream_t s) {
  if (a.tc_we[0] != nullptr) return launch_aux_forward_tc(a, s);
  const long long ld = a.S;  // leading dimension of the [B, cap] scratch matrices
  SgemmArgs g{};
  // h_L = x . W_enc_t[L]^T
  g.A = a.x; g.lda = a.D;
  g.B = a.W_enc_t; g.ldb = a.D; g.gatherB = a.dead_list;
  g.C = a.h_aux; g.ldc = ld;
  g.M = a.B; g.N = a.S; g.K = a.D;
  g.dyn = a.n_dead; g.dyn_which = 2;
  g.gatherM = nullptr; g.alpha = 1.f;
  launch_sgemm<false, true>(g, s);
  aux_select_kernel<<<min(a.B, 148 * 8), 256, 0, s>>>(a.h_aux, a.mask_aux, ld, a.b_enc, a.dead_list, a.n_dead, a.B,
                                                       a.k_aux);
                                                       ++g_launch_count;
  // x_aux (without bias) = f_aux . W_dec[L]
  SgemmArgs d{};
  d.A = a.h_aux; d.lda = ld;
  d.B = a.W_dec; d.ldb = a.D; d.gatherB = a.dead_list;
  d.C = a.r_aux; d.ldc = a.D;
  d.M = a.B; d.N = a.D; d.K = a.S;
  d.dyn = a.n_dead; d.dyn_which = 3;
  d.gatherM = nullptr; d.alpha = 1.f;
  launch_sgemm<false, false>(d, s);
  aux_resid_kernel<<<min((a.B + 7) / 8, 148 * 8), 256, 0, s>>>(a.r_aux, a.resid, a.b_dec, a.B, a.D, a.n_dead,
                                                                a.row_sse_aux);
                                                                ++g_launch_count;
  aux_loss_kernel<<<1, 1024, 0, s>>>(a.row_sse_aux, a.B, a.alpha, a.inv_bd, a.aux_loss);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

int launch_aux_backward(const AuxArgs& a, cudaStream_t s) {
  if (a.tc_we[0] != nullptr) return launch_aux_backward_tc(a, s);
  const long long ld = a.S;
  const float gscale = 2.f * a.alpha * a.inv_bd;
  // gb_dec += sum_b G_a      (r_aux is all zeros when nothing is dead)
  if (launch_colsum(a.r_aux, a.B, a.D, gscale, 1, a.colsum_partial, a.gb_dec, s, 0, a.n_dead)) return 22;
  // gW_dec[L] = f_aux^T G_a
  SgemmArgs g{};
  g.A = a.h_aux; g.lda = ld;             // A(m,k) = f_aux[k, m]
  g.B = a.r_aux; g.ldb = a.D; g.gatherB = nullptr;
  g.C = a.gW_dec; g.ldc = a.D; g.gatherM = a.dead_list;
  g.M = a.S; g.N = a.D; g.K = a.B;
  g.dyn = a.n_dead; g.dyn_which = 1;
  g.alpha = gscale;
  launch_sgemm<true, false>(g, s);
  if (a.remove_parallel)
    aux_project_kernel<<<148 * 4, 256, 0, s>>>(a.gW_dec, a.W_dec, a.dead_list, a.n_dead, a.D);
    ++g_launch_count;
  // dh_a = mask_a * (G_a . W_dec[L]^T)   (overwrites f_aux, which is no longer needed)
  SgemmArgs e{};
  e.A = a.r_aux; e.lda = a.D;
  e.B = a.W_dec; e.ldb = a.D; e.gatherB = a.dead_list;
  e.C = a.h_aux; e.ldc = ld; e.gatherM = nullptr;
  e.M = a.B; e.N = a.S; e.K = a.D;
  e.dyn = a.n_dead; e.dyn_which = 2;
  e.alpha = gscale;
  launch_sgemm<false, true>(e, s);
  aux_mask_colsum_kernel<<<dim3(148, AUX_SLABS), 256, 0, s>>>(a.h_aux, a.mask_aux, ld, a.B, a.n_dead, a.aux_colpart);
  ++g_launch_count;
  aux_colsum_final_kernel<<<148, 256, 0, s>>>(a.aux_colpart, ld, a.dead_list, a.n_dead, a.gb_enc);
  ++g_launch_count;
  // gW_enc_t[L] = dh_a^T x
  SgemmArgs w{};
  w.A = a.h_aux; w.lda = ld;             // A(m,k) = dh_a[k, m]
  w.B = a.x; w.ldb = a.D; w.gatherB = nullptr;
  w.C = a.gW_enc_t; w.ldc = a.D; w.gatherM = a.dead_list;
  w.M = a.S; w.N = a.D; w.K = a.B;
  w.dyn = a.n_dead; w.dyn_which = 1;
  w.alpha = 1.f;
  launch_sgemm<true, false>(w, s);
  if (a.row_gsq != nullptr) {
    aux_rows_gsq_kernel<<<148 * 2, 256, 0, s>>>(a.gW_dec, a.gW_enc_t, a.gb_enc, a.dead_list, a.n_dead, a.D, a.row_gsq);
    ++g_launch_count;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

}  // namespace sb
