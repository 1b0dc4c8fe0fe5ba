// BatchTopK on the sparse forward state (saev src/saev/nn/modeling.py:183-244).
//
// The reference flattens the [B, d_sae] pre-activations, keeps the B * k largest entries of the whole batch
// (`torch.topk(x_flat, k * bsz)` -> scatter mask, :227-235) and folds the smallest positive survivor into an EMA
// threshold (:237-242); in eval mode it is an element-wise JumpReLU with that threshold (:220-224).
//
// Here the screen + exact re-score (encode_gemm2.cu, sparse_kernels.cu) have already left, per row, its `cap`
// largest exact fp32 pre-activations in rank order (topk_idx / topk_val [B, cap], cap = the handle's top_k).  As long
// as no row owns more than `cap` of the batch-wide winners, the B * k largest entries of the batch are the B * k
// largest of these B * cap values:
//
//   btk_hist_kernel / btk_pick_kernel   4 x (8-bit digit histogram over all valid entries, then the bin holding the
//                                       n-th largest key): the exact key of the n-th largest value, and how many
//                                       entries EQUAL to it are kept (ties).
//   btk_row_ties_kernel / btk_tie_scan_kernel   ties at the cut value are granted in flat (row, rank) order: per-row
//                                       tie counts, then their exclusive scan (one block).
//   btk_apply_kernel                    one warp per row: empties the slots that lost (idx = -1, value 0), rebuilds the
//                                       per-atom counts / activity flags from the survivors, tracks the smallest
//                                       positive survivor and counts the rows whose capacity may have truncated the
//                                       selection (all `cap` slots kept while d_sae > cap).
//   btk_finish_kernel                   threshold <- (1 - momentum) threshold + momentum min_pos ; stats.
//
// Eval: only the last two, with the cut `value > max(threshold, 0)`.
#include "common.cuh"
#include "kernels.h"

namespace sb {
namespace {

// state block (ints) inside the scratch region: [0] prefix (key bits found so far), [1] need (rank inside the current
// bin), [2] kept entries, [3] truncated rows, [4] entries tied at the cut, [5] min positive survivor (float bits),
// [6] valid entries, [8 .. 8 + 256) histogram, then tie_off[B]
constexpr int BT_PREFIX = 0, BT_NEED = 1, BT_KEPT = 2, BT_TRUNC = 3, BT_TIES = 4, BT_MINPOS = 5, BT_VALID = 6, BT_HIST = 8,
              BT_TIEOFF = BT_HIST + 256;

__global__ void btk_init_kernel(int* st, long long n_keep) {
  const int t = threadIdx.x;
  if (t < 256) st[BT_HIST + t] = 0;
  if (t == 0) {
    st[BT_PREFIX] = 0;
    st[BT_NEED] = static_cast<int>(n_keep);
    st[BT_KEPT] = 0;
    st[BT_TRUNC] = 0;
    st[BT_TIES] = 0;
    st[BT_MINPOS] = 0x7f800000;  // +inf
    st[BT_VALID] = 0;
  }
}

// histogram of the 8-bit digit at `shift` over the keys whose higher digits equal the prefix found so far
__global__ void __launch_bounds__(256) btk_hist_kernel(const int* __restrict__ idx, const float* __restrict__ val, long long n,
                                                       int shift, int* st) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const unsigned int prefix = static_cast<unsigned int>(st[BT_PREFIX]);
  int valid = 0;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    if (idx[i] < 0) continue;
    ++valid;
    const unsigned int key = fkey(val[i]);
    if (shift == 24 || (key >> (shift + 8)) == prefix) atomicAdd(h + ((key >> shift) & 255u), 1);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(st + BT_HIST + threadIdx.x, h[threadIdx.x]);
  if (shift == 24) {
    valid = __reduce_add_sync(FULL, valid);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(st + BT_VALID, valid);
  }
}

// the bin that holds the `need`-th largest key of this pass; one warp
__global__ void btk_pick_kernel(int* st, int shift) {
  const int lane = threadIdx.x;
  int need = st[BT_NEED];
  if (shift == 24 && st[BT_VALID] <= need) {
    // fewer valid entries than the budget: everything is kept (test_batchtopk_k_exceeds_total_elements); key 0 is below
    // every real key, and the remaining passes keep it there
    need = 0;
  }
  int c[8], mine = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    c[j] = st[BT_HIST + 8 * lane + j];
    mine += c[j];
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) st[BT_HIST + 8 * lane + j] = 0;
  if (need <= 0) {
    if (lane == 0) {
      st[BT_PREFIX] = 0;
      st[BT_NEED] = 0;
    }
    return;
  }
  int suf = mine;  // keys in this lane's bins and above
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(FULL, suf, o);
    if (lane + o < 32) suf += t;
  }
  const unsigned int bal = __ballot_sync(FULL, suf >= need);
  const int L = 31 - __clz(bal);  // highest lane whose suffix count still reaches `need` (bal != 0: need <= valid)
  int cum = suf - mine, j = 7;
#pragma unroll
  for (int jj = 7; jj > 0; --jj) {
    if (j == jj && cum + c[jj] < need) {
      cum += c[jj];
      j = jj - 1;
    }
  }
  if (lane == L) {
    st[BT_PREFIX] = static_cast<int>((static_cast<unsigned int>(st[BT_PREFIX]) << 8) | static_cast<unsigned int>(8 * lane + j));
    st[BT_NEED] = need - cum;  // rank of the n-th largest inside that bin
  }
}

// tie_off[b] <- number of entries of row b equal to the cut value; one warp per row
__global__ void __launch_bounds__(256) btk_row_ties_kernel(const int* __restrict__ idx, const float* __restrict__ val, int B,
                                                           int cap, int* st) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  const unsigned int tkey = static_cast<unsigned int>(st[BT_PREFIX]);
  int mine = 0;
  if (tkey != 0u)
    for (int r = lane; r < cap; r += 32) {
      const long long i = static_cast<long long>(b) * cap + r;
      mine += (idx[i] >= 0 && fkey(val[i]) == tkey) ? 1 : 0;
    }
  mine = __reduce_add_sync(FULL, mine);
  if (lane == 0) st[BT_TIEOFF + b] = mine;
}

// in place: tie_off[b] <- number of entries equal to the cut value in rows < b (ties are granted in flat order); one block
__global__ void __launch_bounds__(1024) btk_tie_scan_kernel(int B, int* st) {
  __shared__ int part[1024];
  int* tie_off = st + BT_TIEOFF;
  const int t = threadIdx.x;
  const int per = (B + 1023) / 1024;
  const int b0 = min(B, t * per), b1 = min(B, b0 + per);
  int mine = 0;
  for (int b = b0; b < b1; ++b) mine += tie_off[b];
  part[t] = mine;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan of the 1024 partials (Hillis-Steele)
    const int v = (t >= o) ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = part[t] - mine;
  for (int b = b0; b < b1; ++b) {
    const int c = tie_off[b];
    tie_off[b] = run;
    run += c;
  }
  if (t == 1023) st[BT_TIES] = part[1023];
}

// one warp per row.  mode 0 (training): keep key > tkey, plus the first (need - tie_off[b]) entries equal to it;
// mode 1 (eval): keep value > max(*threshold, 0)
__global__ void __launch_bounds__(256) btk_apply_kernel(int* __restrict__ idx, float* __restrict__ val, int B, int cap, int S,
                                                        int mode, const float* __restrict__ threshold, int* st,
                                                        int* __restrict__ feat_count, int* __restrict__ active) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  const unsigned int tkey = static_cast<unsigned int>(st[BT_PREFIX]);
  const int need_eq = st[BT_NEED];
  const float theta = (mode == 1 && threshold != nullptr) ? fmaxf(*threshold, 0.f) : 0.f;
  int allowed = (mode == 0) ? max(0, need_eq - st[BT_TIEOFF + b]) : 0;
  int kept = 0, slots = 0;
  unsigned int minpos = 0x7f800000u;
  for (int r0 = 0; r0 < cap; r0 += 32) {
    const int r = r0 + lane;
    const long long i = static_cast<long long>(b) * cap + r;
    const int j = (r < cap) ? idx[i] : -1;
    const float v = (r < cap) ? val[i] : 0.f;
    bool keep = false, eq = false;
    if (j >= 0) {
      if (mode == 0) {
        const unsigned int key = fkey(v);
        keep = key > tkey;
        eq = key == tkey && tkey != 0u;
      } else {
        keep = v > theta;
      }
    }
    const unsigned int bal_eq = __ballot_sync(FULL, eq);
    if (eq && __popc(bal_eq & ((1u << lane) - 1u)) < allowed) keep = true;
    allowed = max(0, allowed - __popc(bal_eq));
    if (r < cap) {
      if (keep) {
        if (feat_count != nullptr) atomicAdd(feat_count + j, 1);
        if (active != nullptr && v != 0.f) active[j] = 1;
        if (v > 0.f) minpos = min(minpos, __float_as_uint(v));
      } else if (j >= 0) {
        idx[i] = -1;
        val[i] = 0.f;
      }
    }
    kept += __popc(__ballot_sync(FULL, keep));
    slots += __popc(__ballot_sync(FULL, j >= 0));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) minpos = min(minpos, __shfl_xor_sync(FULL, minpos, o));
  if (lane == 0) {
    if (kept) atomicAdd(st + BT_KEPT, kept);
    // every slot the row had was kept and the dictionary is wider than the row's capacity: the reference may have kept
    // further entries of this row (and, in training, a different cut) -- not certifiable from the lists
    if (kept == slots && slots == cap && cap < S) atomicAdd(st + BT_TRUNC, 1);
    if (minpos != 0x7f800000u) atomicMin(reinterpret_cast<unsigned int*>(st) + BT_MINPOS, minpos);
  }
}

__global__ void btk_finish_kernel(int* st, int training, float* threshold, float momentum, int* stats) {
  if (threadIdx.x != 0) return;
  if (training && threshold != nullptr) {
    const unsigned int mp = static_cast<unsigned int>(st[BT_MINPOS]);
    // modeling.py:237-242 (the reference's pos.min() raises on an empty selection; nothing to fold in then)
    if (mp != 0x7f800000u) *threshold = (1.f - momentum) * *threshold + momentum * __uint_as_float(mp);
  }
  if (stats != nullptr) {
    stats[0] = st[BT_KEPT];
    stats[1] = st[BT_TRUNC];
    stats[2] = training ? st[BT_TIES] : 0;
    stats[3] = st[BT_PREFIX];  // key of the cut value (training)
  }
}

}  // namespace

size_t batch_topk_scratch_bytes(int max_batch) { return (static_cast<size_t>(BT_TIEOFF) + max_batch + 8) * 4; }

int launch_batch_topk(int* topk_idx, float* topk_val, int B, int cap, int S, long long n_keep, int training,
                      float* threshold, float momentum, int* feat_count, int* active, int* scratch, int* stats,
                      cudaStream_t s) {
  if (B <= 0 || cap <= 0) return 21;
  const long long n = static_cast<long long>(B) * cap;
  if (n_keep > n) n_keep = n;
  if (n_keep >= (1ll << 31)) return 24;
  if (active != nullptr && cudaMemsetAsync(active, 0, static_cast<size_t>(S) * 4, s) != cudaSuccess) return 23;
  if (feat_count != nullptr && cudaMemsetAsync(feat_count, 0, static_cast<size_t>(S) * 4, s) != cudaSuccess) return 23;
  btk_init_kernel<<<1, 256, 0, s>>>(scratch, n_keep);
  ++g_launch_count;
  if (training) {
    const long long want = (n + 2047) / 2048;
    const int grid = static_cast<int>(want < 592 ? want : 592);
    for (int shift = 24; shift >= 0; shift -= 8) {
      btk_hist_kernel<<<grid, 256, 0, s>>>(topk_idx, topk_val, n, shift, scratch);
      btk_pick_kernel<<<1, 32, 0, s>>>(scratch, shift);
      g_launch_count += 2;
    }
    btk_row_ties_kernel<<<(B + 7) / 8, 256, 0, s>>>(topk_idx, topk_val, B, cap, scratch);
    btk_tie_scan_kernel<<<1, 1024, 0, s>>>(B, scratch);
    g_launch_count += 2;
  }
  btk_apply_kernel<<<(B + 7) / 8, 256, 0, s>>>(topk_idx, topk_val, B, cap, S, training ? 0 : 1, threshold, scratch,
                                               feat_count, active);
  btk_finish_kernel<<<1, 32, 0, s>>>(scratch, training, threshold, momentum, stats);
  g_launch_count += 2;
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

}  // namespace sb
