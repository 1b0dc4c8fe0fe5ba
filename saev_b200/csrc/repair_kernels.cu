// Exact top-k for the rows the tensor-core screen could not certify (saev modeling.py:169-179 on the fp32
// pre-activations of modeling.py:344-347, no screen involved).
//
// rescore_topk_kernel appends a row to `unsafe_list` when one of its candidate lists overflowed (more near-ties than a
// list holds), when an observed screen error exceeds the deterministic bound, or when the encoder leaves the fp16
// range.  For those rows every one of the d_sae pre-activations is recomputed in fp32 (same FMA chain + shuffle tree
// as the re-score, common.cuh row_dot) and the k largest are selected by a block-wide radix select with the
// reference's tie rule (value descending, then column ascending).
//
//   repair_scan_kernel    grid (n_slices, G): block (s, g) takes groups of RP_ROWS unsafe rows (g, g + G, ...) and
//                         the s-th slice of the dictionary; it walks the slice in tiles of RP_TILE atoms (8 warps, one
//                         atom per warp at a time, two in flight, each multiplied with all rows of the group), keeps
//                         the best k so far per row in shared memory and leaves them, in column order, in the row's
//                         own candidate-list storage (consumed by then): slot [s][0..k).
//   repair_select_kernel  grid R: merges the n_slices x k survivors of a row and writes topk_idx / topk_val
//                         (rank order), the per-atom counts and the activity flags, exactly what the re-score
//                         kernel writes for a certified row.
//
// Both run fixed grids and read the number of unsafe rows from the device; with none they exit at once.
#include "common.cuh"
#include "kernels.h"

namespace sb {
namespace {

constexpr int RP_THREADS = 256;
constexpr int RP_WARPS = RP_THREADS / 32;
constexpr int RP_TILE = 1024;      // atoms per selection round of the scan
constexpr int RP_ROWS = 8;         // unsafe rows multiplied with one load of a dictionary row (4 when d_model > 1024)
constexpr int RP_KMAX = 128;       // top_k <= 128 (saev_b200_create)
constexpr int RP_MAX_SLICES = 64;
constexpr int RP_MERGE_CAP = 4096; // survivors repair_select_kernel merges per row: n_slices * top_k (64 slices up to top_k = 64)
constexpr int RP_ROW_SLOTS = 8;    // gridDim.y of the scan (row groups in flight)

__device__ __forceinline__ float4 ldg4r(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Block-wide selection of the k largest of v[0 .. n) (n <= RP_TILE + RP_KMAX, 256 threads), ties towards the lower
// index.  The winners are written in INDEX order to out_v / out_c (out_c[i] = col_of(index)); returns their number
// (min(k, n)).  `hist` = 256 ints, `wtot` = 2 * RP_WARPS ints, `bc` = 4 ints of shared scratch.
template <typename ColFn>
__device__ int block_select_topk(const float* v, int n, int k, ColFn col_of, float* out_v, int* out_c, int* hist,
                                 int* wtot, int* bc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (n <= k) {
    for (int i = tid; i < n; i += RP_THREADS) {
      out_v[i] = v[i];
      out_c[i] = col_of(i);
    }
    __syncthreads();
    return n;
  }
  // ---- exact k-th largest key: four 8-bit radix passes, most significant digit first ----
  unsigned int prefix = 0u;
  int need = k;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += RP_THREADS) {
      const unsigned int key = fkey(v[i]);
      if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(hist + ((key >> shift) & 255u), 1);
    }
    __syncthreads();
    if (warp == 0) {
      int c[8], mine = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c[j] = hist[8 * lane + j];
        mine += c[j];
      }
      int suf = mine;  // keys in this lane's bins and above
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(FULL, suf, o);
        if (lane + o < 32) suf += t;
      }
      const unsigned int bal = __ballot_sync(FULL, suf >= need);
      const int L = 31 - __clz(bal);  // highest lane whose suffix count still reaches `need`
      int cum = suf - mine, j = 7;
#pragma unroll
      for (int jj = 7; jj > 0; --jj) {
        if (j == jj && cum + c[jj] < need) {
          cum += c[jj];
          j = jj - 1;
        }
      }
      if (lane == L) {
        bc[0] = 8 * lane + j;   // the digit
        bc[1] = need - cum;     // rank of the k-th largest inside that bin
      }
    }
    __syncthreads();
    prefix = (prefix << 8) | static_cast<unsigned int>(bc[0]);
    need = bc[1];
    __syncthreads();
  }
  const unsigned int tkey = prefix;  // key of the k-th largest value; `need` of the entries equal to it are taken
  // ---- ordered compaction: everything above the threshold, and the first `need` ties ----
  int taken_before = 0, eq_before = 0;
  for (int base = 0; base < n; base += RP_THREADS) {
    const int i = base + tid;
    const unsigned int key = (i < n) ? fkey(v[i]) : 0u;
    const bool gt = (i < n) && key > tkey;
    const bool eq = (i < n) && key == tkey;
    const unsigned int bal_eq = __ballot_sync(FULL, eq);
    if (lane == 0) wtot[warp] = __popc(bal_eq);
    __syncthreads();
    int eq_rank = eq_before + __popc(bal_eq & ((1u << lane) - 1u));
    int eq_chunk = 0;
    for (int w = 0; w < RP_WARPS; ++w) {
      if (w < warp) eq_rank += wtot[w];
      eq_chunk += wtot[w];
    }
    const bool take = gt || (eq && eq_rank < need);
    const unsigned int bal = __ballot_sync(FULL, take);
    if (lane == 0) wtot[RP_WARPS + warp] = __popc(bal);
    __syncthreads();
    int pos = taken_before + __popc(bal & ((1u << lane) - 1u));
    int chunk = 0;
    for (int w = 0; w < RP_WARPS; ++w) {
      if (w < warp) pos += wtot[RP_WARPS + w];
      chunk += wtot[RP_WARPS + w];
    }
    if (take) {
      out_v[pos] = v[i];
      out_c[pos] = col_of(i);
    }
    taken_before += chunk;
    eq_before += eq_chunk;
    __syncthreads();
  }
  return taken_before;  // == k
}

// same FMA chain + shuffle tree as row_dot (common.cuh) with the x row read from shared memory: bit-identical values
template <int VPL>
__device__ __forceinline__ float row_dot_smem(const float4* __restrict__ xs4, const float4 (&w)[VPL], int lane, int D4) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    const float4 x = (v < D4) ? xs4[v] : make_float4(0, 0, 0, 0);
    acc = fmaf(x.x, w[i].x, acc);
    acc = fmaf(x.y, w[i].y, acc);
    acc = fmaf(x.z, w[i].z, acc);
    acc = fmaf(x.w, w[i].w, acc);
  }
  return warp_sum(acc);
}

// R unsafe rows per pass over the block's slice of the dictionary: every W_enc_t row is loaded once and multiplied
// with all R batch rows (staged in shared memory), so repairing a handful of rows costs one sweep of the dictionary
// (~270 MB at c3: ~50 us over 64 slices) instead of one sweep per row.
template <int VPL, int R>
__global__ void __launch_bounds__(RP_THREADS) repair_scan_kernel(RescoreArgs a, int n_slices, int slice_len) {
  extern __shared__ float4 rp_smem4[];
  __shared__ float best_v[R][RP_KMAX], tmp_v[RP_KMAX];
  __shared__ int best_c[R][RP_KMAX], tmp_c[RP_KMAX];
  __shared__ int hist[256], wtot[2 * RP_WARPS], bc[4];
  const int n_unsafe = min(reinterpret_cast<const int*>(a.scalars)[SC_N_UNSAFE], a.B);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D4 = a.D >> 2, K = a.K;
  float4* xs4 = rp_smem4;                                                   // [R][D4]
  float* hv = reinterpret_cast<float*>(rp_smem4 + static_cast<size_t>(R) * D4);  // [R][RP_TILE + RP_KMAX]
  constexpr int HV = RP_TILE + RP_KMAX;
  const int j_begin = blockIdx.x * slice_len, j_end = min(a.S, j_begin + slice_len);
  for (int g0 = blockIdx.y * R; g0 < n_unsafe; g0 += gridDim.y * R) {
    const int nr = min(R, n_unsafe - g0);
    __syncthreads();
    for (int i = tid; i < R * D4; i += RP_THREADS) {
      const int r = i / D4, v = i - r * D4;
      xs4[i] = r < nr ? ldg4r(a.x + static_cast<long long>(a.unsafe_list[g0 + r]) * a.D + 4 * v) : make_float4(0, 0, 0, 0);
    }
    int n_best = 0;  // (the same for every row of the group: min(K, atoms seen so far))
    __syncthreads();
    for (int t0 = j_begin; t0 < j_end; t0 += RP_TILE) {
      const int nt = min(RP_TILE, j_end - t0);
      // the carried winners come from lower columns: they go in front, so that index order stays column order
      for (int i = tid; i < R * n_best; i += RP_THREADS) {
        const int r = i / n_best, e = i - r * n_best;
        hv[r * HV + e] = best_v[r][e];
      }
      for (int a0 = warp; a0 < nt; a0 += 2 * RP_WARPS) {
        const int a1 = a0 + RP_WARPS;
        const int j0 = t0 + a0, j1 = t0 + min(a1, nt - 1);
        float4 w0[VPL], w1[VPL];
        const float* r0 = a.W_enc_t + static_cast<long long>(j0) * a.D;
        const float* r1 = a.W_enc_t + static_cast<long long>(j1) * a.D;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int v = lane + 32 * i;
          w0[i] = (v < D4) ? ldg4r(r0 + 4 * v) : make_float4(0, 0, 0, 0);
          w1[i] = (v < D4) ? ldg4r(r1 + 4 * v) : make_float4(0, 0, 0, 0);
        }
        const float b0 = __ldg(a.b_enc + j0), b1 = __ldg(a.b_enc + j1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float h0 = row_dot_smem<VPL>(xs4 + r * D4, w0, lane, D4);
          const float h1 = row_dot_smem<VPL>(xs4 + r * D4, w1, lane, D4);
          if (lane == 0) {
            hv[r * HV + n_best + a0] = h0 + b0;
            if (a1 < nt) hv[r * HV + n_best + a1] = h1 + b1;
          }
        }
      }
      __syncthreads();
      const int nb = n_best;
      int got = 0;
      for (int r = 0; r < nr; ++r) {
        const int* bcr = best_c[r];
        got = block_select_topk(hv + r * HV, nb + nt, K, [&](int i) { return i < nb ? bcr[i] : t0 + (i - nb); }, tmp_v,
                                tmp_c, hist, wtot, bc);
        __syncthreads();
        for (int i = tid; i < got; i += RP_THREADS) {
          best_v[r][i] = tmp_v[i];
          best_c[r][i] = tmp_c[i];
        }
        __syncthreads();
      }
      n_best = got;
    }
    // slot [slice][0 .. K) of each row's candidate storage; unused entries carry column -1
    for (int i = tid; i < nr * K; i += RP_THREADS) {
      const int r = i / K, e = i - r * K;
      int2* out = reinterpret_cast<int2*>(a.cand) + static_cast<long long>(a.unsafe_list[g0 + r]) * a.nsplit * a.cand_stride +
                  static_cast<long long>(blockIdx.x) * K;
      out[e] = e < n_best ? make_int2(__float_as_int(best_v[r][e]), best_c[r][e]) : make_int2(0, -1);
    }
  }
}

__global__ void __launch_bounds__(RP_THREADS) repair_select_kernel(RescoreArgs a, int n_slices) {
  __shared__ float hv[RP_MERGE_CAP];
  __shared__ int hc[RP_MERGE_CAP];
  __shared__ float win_v[RP_KMAX];
  __shared__ int win_c[RP_KMAX];
  __shared__ int hist[256], wtot[2 * RP_WARPS], bc[4];
  __shared__ int s_n;
  const int n_unsafe = min(reinterpret_cast<const int*>(a.scalars)[SC_N_UNSAFE], a.B);
  const int tid = threadIdx.x, K = a.K;
  for (int u = blockIdx.x; u < n_unsafe; u += gridDim.x) {
    const int b = a.unsafe_list[u];
    const int2* in = reinterpret_cast<const int2*>(a.cand) + static_cast<long long>(b) * a.nsplit * a.cand_stride;
    // gather the survivors of all slices, keeping (slice, position) = column order; one thread compacts the holes
    if (tid == 0) s_n = 0;
    __syncthreads();
    if (tid < 32) {  // warp 0: ordered compaction, 32 entries per step
      int n = 0;
      for (int e0 = 0; e0 < n_slices * K; e0 += 32) {
        const int e = e0 + tid;
        int2 t = make_int2(0, -1);
        if (e < n_slices * K) t = __ldcg(in + e);
        const unsigned int bal = __ballot_sync(FULL, t.y >= 0);
        if (t.y >= 0) {
          const int o = n + __popc(bal & ((1u << tid) - 1u));
          hv[o] = __int_as_float(t.x);
          hc[o] = t.y;
        }
        n += __popc(bal);
      }
      if (tid == 0) s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    const int got = block_select_topk(hv, n, K, [&](int i) { return hc[i]; }, win_v, win_c, hist, wtot, bc);
    __syncthreads();
    // rank order (value desc, column asc), as rescore_topk_kernel writes it
    if (tid < got) {
      const float v = win_v[tid];
      const int id = win_c[tid];
      int rank = 0;
      for (int t = 0; t < got; ++t) rank += (win_v[t] > v) || (win_v[t] == v && win_c[t] < id);
      const long long o = static_cast<long long>(b) * K + rank;
      a.topk_idx[o] = id;
      a.topk_val[o] = v;
      if (a.feat_count) atomicAdd(a.feat_count + id, 1);
      if (a.active && v != 0.f) a.active[id] = 1;
    }
    for (int r = got + tid; r < K; r += RP_THREADS) {
      a.topk_idx[static_cast<long long>(b) * K + r] = -1;
      a.topk_val[static_cast<long long>(b) * K + r] = 0.f;
    }
    if (tid == 0) atomicAdd(reinterpret_cast<unsigned int*>(a.scalars) + SC_REPAIRED, 1u);
    __syncthreads();
  }
}

}  // namespace

int launch_repair_topk(const RescoreArgs& a, cudaStream_t s) {
  if (a.D % 4 || a.K > RP_KMAX || a.K < 1) return 21;
  // slices: as many as the row's candidate storage can hold k survivors for (at most RP_MAX_SLICES)
  const long long cap = static_cast<long long>(a.nsplit) * a.cand_stride;
  int n_slices = static_cast<int>(cap / a.K);
  if (n_slices > RP_MAX_SLICES) n_slices = RP_MAX_SLICES;
  if (n_slices > RP_MERGE_CAP / a.K) n_slices = RP_MERGE_CAP / a.K;
  const int by_len = (a.S + 255) / 256;  // no slice shorter than 256 atoms
  if (n_slices > by_len) n_slices = by_len;
  if (n_slices < 1) return 24;
  const int slice_len = (a.S + n_slices - 1) / n_slices;
  n_slices = (a.S + slice_len - 1) / slice_len;
  const dim3 grid(n_slices, RP_ROW_SLOTS);
  const int need = (a.D + 127) / 128;
  ++g_launch_count;
#define SB_REPAIR(V, R)                                                                                              \
  {                                                                                                                  \
    const size_t smem = static_cast<size_t>(R) * (a.D / 4) * 16 + static_cast<size_t>(R) * (RP_TILE + RP_KMAX) * 4;  \
    static bool attr_set = false;                                                                                    \
    if (!attr_set) {                                                                                                 \
      if (cudaFuncSetAttribute(repair_scan_kernel<V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                               static_cast<int>(RP_ROWS * 2048 * 4 + RP_ROWS * (RP_TILE + RP_KMAX) * 4)) != cudaSuccess) \
        return 3;                                                                                                    \
      attr_set = true;                                                                                               \
    }                                                                                                                \
    repair_scan_kernel<V, R><<<grid, RP_THREADS, smem, s>>>(a, n_slices, slice_len);                                 \
  }
  if (need <= 1) SB_REPAIR(1, RP_ROWS)
  else if (need <= 2) SB_REPAIR(2, RP_ROWS)
  else if (need <= 4) SB_REPAIR(4, RP_ROWS)
  else if (need <= 6) SB_REPAIR(6, RP_ROWS)
  else if (need <= 8) SB_REPAIR(8, RP_ROWS)
  else if (need <= 12) SB_REPAIR(12, RP_ROWS / 2)
  else if (need <= 16) SB_REPAIR(16, RP_ROWS / 2)
  else return 20;
#undef SB_REPAIR
  repair_select_kernel<<<64, RP_THREADS, 0, s>>>(a, n_slices);
  ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

}  // namespace sb
