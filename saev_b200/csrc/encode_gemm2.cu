// Encoder contraction + streaming top-k screen (tcgen05.mma.cta_group::2 on CTA pairs).
//
// Replaces the dense `einsum(x, W_enc) + b_enc` of saev (src/saev/nn/modeling.py:343-347) and the
// `topk -> scatter -> mul` of TopKActivation.forward (modeling.py:169-179): the [B, S] pre-activation matrix is never
// written to HBM.  The tensor cores compute a SCREEN value  h~ = 2^e_b <fp16(x_b 2^-e_b), fp16(w_j)> + b_j  (fp16
// operands: 11 significant bits at the bf16 rate; the power-of-two row scale keeps every row inside the fp16 range
// without rounding) and every row of the batch leaves candidate lists that PROVABLY cover its exact top-k:
//
//   admission rule: with E_bj = c_j P_b + Q_b the deterministic bound on |h~_bj - h_bj| of kernels.h
//   (screen_bound: Cauchy-Schwarz on the fp16 roundings, plus the accumulation terms), every column has the interval
//   [l_j, u_j] = [h~_j - E_bj, h~_j + E_bj] around its exact value.  L = the k-th largest LOWER bound seen so far in
//   the row is a lower bound of the exact k-th largest value, so a column whose UPPER bound is below L can never be
//   in the exact top-k; everything else is kept:  admit iff  h~_j + c_j P_b > L - Q_b.  The lists store
//   (l_j, column).  `rescore_topk_kernel` (sparse_kernels.cu) recomputes the candidates in fp32 from the fp32 master
//   weights, picks the final top-k, and hands rows whose lists overflowed (or whose observed error contradicts the
//   bound) to the exact repair path.
//
// Organisation (profiles/r01_encode_gemm_full.md has the measurements that led here):
//
//   * operand feed: a CTA pair (two SMs of one TPC) computes a 256 x 256 tile; each CTA stages only ITS 128 rows of
//     x and ITS 128 of the 256 dictionary columns (32 KB per k-block instead of 48 KB), the tensor cores of both SMs
//     read both halves.  That is 2/3 of the L2->SM bytes per FLOP and leaves room for a 6-deep TMA ring.
//   * work distribution: the (row block, column tile) space is linearised and cut into equal contiguous ranges, one
//     per resident CTA pair (all 148 SMs busy; a row block may be covered by up to `nsplit` ranges, each leaving its
//     own candidate list).
//   * epilogue issue rate: eight epilogue warps per CTA (two per SM sub-partition, so the dependent
//     FADD -> FMNMX -> compare chains of one warp are hidden by the other); the two warps that share a TMEM lane
//     quadrant take the low / high 128 columns of every tile and exchange their admission thresholds through shared
//     memory, so the threshold of a row tightens as fast as with a single stream.
//
// Roles (384 threads): warp 0 = TMA producer (both CTAs), warp 1 = MMA issuer (leader CTA only), warp 2 = TMEM
// allocator, warps 4-11 = epilogue.  Barriers: full/empty per smem stage (full lives in the leader, both CTAs'
// TMA transactions complete on it; empty is multicast to both CTAs by tcgen05.commit), tfull (multicast commit)
// / tempty (leader, 16 arrivals: 8 warps x 2 CTAs) per TMEM accumulator stage.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "pair_common.cuh"

namespace sb {
namespace g2 {
using namespace pairx;

constexpr int BM = 128;        // rows per CTA (pair: 256 = UMMA M)
#ifndef SB_G2_BN
#define SB_G2_BN 256
#endif
constexpr int BN = SB_G2_BN;   // dictionary columns per tile (UMMA N): 256 (2 accumulator stages) or 128 (4 stages)
constexpr int ACC = 512 / BN;  // TMEM accumulator stages (512 columns of TMEM per SM)
constexpr int BK = 64;         // bf16 per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 16;
constexpr int STAGES = BN == 256 ? 6 : 8;
constexpr int A_BYTES = BM * BK * 2;         // 16 KB: this CTA's rows
constexpr int B_BYTES = (BN / 2) * BK * 2;   // 16 KB: this CTA's half of the tile's columns
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int CHUNK = 16;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
constexpr int HALF = BN / 2;   // columns per epilogue warp per tile
constexpr int CAPG = ENCODE2_CAPG;             // entries per candidate list
constexpr int TRIGGER_MAX = CAPG - HALF;        // a list above this could not absorb a whole further tile

constexpr size_t OFF_BIAS = static_cast<size_t>(STAGES) * STAGE_BYTES;            // [8 warps][2 acc stages][2][128] f32:
                                                                                  // bias slice, then column-norm slice
constexpr size_t OFF_TAU = OFF_BIAS + static_cast<size_t>(EPI_WARPS) * ACC * 2 * HALF * 4;  // [2 halves][128 rows] f32
constexpr size_t OFF_HIST = OFF_TAU + 2 * BM * 4;                                    // [8 warps][256] i32
constexpr size_t OFF_BARS = OFF_HIST + static_cast<size_t>(EPI_WARPS) * 256 * 4;
constexpr size_t SMEM_TOTAL = OFF_BARS + (2 * STAGES + 2 * ACC) * 8 + 16 + 1024;

template <int PER>
__device__ __forceinline__ unsigned int warp_kth_largest(const unsigned int (&key)[PER], int k) {
  unsigned int T = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const unsigned int cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) c += __popc(__ballot_sync(FULL, key[e] >= cand));
    if (c >= k) T = cand;
  }
  return T;
}

// Lower bound (exact in its top 16 bits, low 16 bits zero) of the k-th largest of the warp's 8 x 32 register-resident
// keys: two 8-bit radix passes over a warp-private 256-bin histogram in shared memory.  Needs at least k keys > 0.
__device__ __forceinline__ unsigned int warp_kth_coarse(const unsigned int (&key)[CAPG / 32], int k, int* hist, int lane) {
  unsigned int prefix = 0u;
  int need = k;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int shift = 24 - 8 * pass;
    int4* h4 = reinterpret_cast<int4*>(hist);
    h4[2 * lane] = make_int4(0, 0, 0, 0);
    h4[2 * lane + 1] = make_int4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < CAPG / 32; ++e)
      if (pass == 0 || (key[e] >> 24) == prefix) atomicAdd(hist + ((key[e] >> shift) & 255u), 1);
    __syncwarp();
    const int4 a = h4[2 * lane], b = h4[2 * lane + 1];  // bins 8*lane .. 8*lane+7
    const int c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const int mine = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
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
    const int bin = __shfl_sync(FULL, 8 * lane + j, L);
    need = __shfl_sync(FULL, need - cum, L);
    prefix = (prefix << 8) | static_cast<unsigned int>(bin);
    __syncwarp();
  }
  return prefix << 16;
}

// Warp-cooperative compaction of one row's candidate list (n >= k entries of {value bits, column}): keep every entry
// whose value exceeds max(k-th largest - margin, floor), packed to the front; if more than CAPG/2 qualify keep the
// CAPG/2 largest and report overflow.  Returns the threshold that was applied.
__device__ __forceinline__ float compact_list(int2* buf, int n, int k, float margin, float floor_thr, int lane, int* hist,
                                              int& n_out, bool& ovf) {
  constexpr int PER = CAPG / 32;
  unsigned int key[PER];
  int col[PER];
  __syncwarp();
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int sl = lane + 32 * e;
    key[e] = 0u;
    col[e] = -1;
    if (sl < n) {
      const int2 t = __ldcg(buf + sl);
      key[e] = fkey(__int_as_float(t.x));
      col[e] = t.y;
    }
  }
  ovf = false;
  float thr = fmaxf(funkey(warp_kth_coarse(key, k, hist, lane)) - margin, floor_thr);
  unsigned int tkey = fkey(thr);
  int n_ge = 0, n_gt = 0;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    n_ge += __popc(__ballot_sync(FULL, key[e] >= tkey));
    n_gt += __popc(__ballot_sync(FULL, key[e] > tkey));
  }
  int need_eq = n_ge - n_gt;
  if (n_ge > CAPG / 2) {
    ovf = true;
    tkey = warp_kth_largest<PER>(key, CAPG / 2);
    thr = funkey(tkey);
    n_gt = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) n_gt += __popc(__ballot_sync(FULL, key[e] > tkey));
    need_eq = CAPG / 2 - n_gt;
  }
  const unsigned int lt_mask = (1u << lane) - 1u;
  int base = 0, eq_seen = 0;
  __syncwarp();
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const bool gt = key[e] > tkey;
    const bool eq = key[e] == tkey;
    const unsigned int bal_eq = __ballot_sync(FULL, eq);
    const bool take = gt || (eq && (eq_seen + __popc(bal_eq & lt_mask)) < need_eq);
    const unsigned int bal = __ballot_sync(FULL, take);
    if (take) buf[base + __popc(bal & lt_mask)] = make_int2(__float_as_int(funkey(key[e])), col[e]);
    base += __popc(bal);
    eq_seen += __popc(bal_eq);
  }
  __syncwarp();
  n_out = base;
  return thr;
}

// The pair's contiguous range [t_begin, t_end) of the linearised (row block, tile) space, cut at row-block
// boundaries into units.  Units are PROCESSED head-first (the unit that starts at tile 0 of its row block before the
// unit that ends at the last tile of the previous one): every pair then sweeps the dictionary tiles in roughly the
// same phase, so the bf16 dictionary is streamed from HBM once per launch instead of once per pair.
struct Unit {
  int m_pair, nb, ne;
};
__device__ __forceinline__ int range_units(long long t_begin, long long t_end, int n_tiles) {
  return static_cast<int>((t_end - 1) / n_tiles - t_begin / n_tiles) + 1;
}
__device__ __forceinline__ Unit range_unit(long long t_begin, long long t_end, int n_tiles, int n_units, int j) {
  const int idx = n_units > 1 ? (j + 1) % n_units : 0;  // processing order -> linear order
  const int m0 = static_cast<int>(t_begin / n_tiles);
  Unit u;
  u.m_pair = m0 + idx;
  if (idx == 0) {
    u.nb = static_cast<int>(t_begin - static_cast<long long>(m0) * n_tiles);
    u.ne = static_cast<int>(min(static_cast<long long>(n_tiles), u.nb + (t_end - t_begin)));
  } else {
    u.nb = 0;
    u.ne = static_cast<int>(min(static_cast<long long>(n_tiles), t_end - static_cast<long long>(u.m_pair) * n_tiles));
  }
  return u;
}

struct Params {
  int kblocks;          // ceil(K / 64)
  int M, N;
  int n_tiles;          // ceil(N / 256)
  long long total;      // m_pairs * n_tiles
  int q;                // tile-steps per CTA pair
  int nlists;           // candidate lists per row = 2 * nsplit
  int top_k;
  int trigger;          // compact a list once it holds more than this many entries (<= TRIGGER_MAX)
  const float* bias;
  const float* row_norm;   // [M] ||x_b||_2
  const float* row_dx;     // [M] ||2^e_b x16_b - x_b||_2
  const float* row_scale;  // [M] 2^e_b (the screen value is acc * 2^e_b + bias)
  const float* scalars;    // workspace scalar block: SC_WNORM_SQ_MAX, SC_BIAS_ABS_MAX
  const float* col_norm;   // [N] ||w_j||_2 (rounded up)
  int D;                   // contraction length (for the error bound)
  int debug;               // SAEV_B200_SCREEN_DEBUG (timing experiments only; results are garbage): 1 = admit nothing,
                           // 2 = the epilogue only hands the accumulator back
  int2* cand;
  int* cand_cnt;
  unsigned int* tau_g;  // [rows padded to 256] order-preserving key of the best admission threshold known per row
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
encode_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_u32 + pad;

  float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* tau_s = reinterpret_cast<float*>(smem + OFF_TAU);
  const uint32_t bars = smem_base + static_cast<uint32_t>(OFF_BARS);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + ACC + s); };
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + OFF_BARS + (2 * STAGES + 2 * ACC) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;

  const long long t_begin = static_cast<long long>(pair) * p.q;
  const long long t_end = min(p.total, t_begin + p.q);
  const int n_units = range_units(t_begin, t_end, p.n_tiles);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < ACC; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(tmem_ptr_s));
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int uj = 0; uj < n_units; ++uj) {
        const Unit u = range_unit(t_begin, t_end, p.n_tiles, n_units, uj);
        const int row0 = u.m_pair * (2 * BM) + static_cast<int>(rank) * BM;
        for (int n = u.nb; n < u.ne; ++n) {
          const int col0 = n * BN + static_cast<int>(rank) * (BN / 2);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            tma_load_2d_pair(sa, &tmA, full_bar(stage), kb * BK, row0);
            tma_load_2d_pair(sa + A_BYTES, &tmB, full_bar(stage), kb * BK, col0);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();  // reconverge before the (warp-aligned) cluster barrier at the end
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16_f32(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      const long long n_steps = t_end - t_begin;
      for (long long tc = 0; tc < n_steps; ++tc) {
        const int as = static_cast<int>(tc % ACC);
        const uint32_t aphase = static_cast<uint32_t>(tc / ACC) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = umma_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_f16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
          umma_commit_pair(empty_bar(stage));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_pair(tfull_bar(as));
      }
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (both CTAs) =====================
    const int w = warp - EPI_WARP0;
    const int q = warp & 3;   // TMEM lane quadrant this warp may read
    const int half = w >> 2;  // low / high 128 columns of every tile
    const int row_local = q * 32 + lane;
    float* my_tau_s = tau_s + half * BM + row_local;
    const float* other_tau_s = tau_s + (half ^ 1) * BM + row_local;
    const ScreenBound sbd = screen_bound(p.D, p.scalars[SC_RHO], p.scalars[SC_BIAS_ABS_MAX]);
    const float wn_max = sqrtf(p.scalars[SC_WNORM_SQ_MAX]) * 1.00001f;
    int* hist_w = reinterpret_cast<int*>(smem + OFF_HIST) + w * 256;
    long long tc = 0;
    for (int uj = 0; uj < n_units; ++uj) {
      const Unit u = range_unit(t_begin, t_end, p.n_tiles, n_units, uj);
      const int m_pair = u.m_pair, nb = u.nb, ne = u.ne;
      const int split = pair - static_cast<int>((static_cast<long long>(m_pair) * p.n_tiles) / p.q);
      const int list = split * 2 + half;
      const int row = m_pair * (2 * BM) + static_cast<int>(rank) * BM + row_local;
      int2* warp_buf = p.cand + (static_cast<long long>(row - lane) * p.nlists + list) * CAPG;
      const long long lane_stride = static_cast<long long>(p.nlists) * CAPG;
      int2* my_buf = warp_buf + lane * lane_stride;
      const bool live = row < p.M;
      float tau = live ? -INFINITY : INFINITY;
      const float rs = live ? p.row_scale[row] : 1.f;
      const float Pb = live ? screen_P(sbd, p.row_norm[row], p.row_dx[row]) : 0.f;  // E_bj = c_j Pb + Qb
      const float Qb = screen_Q(sbd);
      const float neg2P = -2.f * Pb;
      // list maintenance works in l-space (stored lower bounds) with the widest band any column can have; the
      // admission threshold `tau` lives in t-space (t_j = h~_j + ||w_j|| Pb):  tau = L - Qb = thr_l + shift
      const float margin = 2.f * (wn_max * Pb + Qb);
      const float shift = margin - Qb;
      int2* wp = my_buf;  // append cursor (a running pointer keeps the per-hit address arithmetic to one add)
      bool overflowed = false;
      *my_tau_s = -INFINITY;
      named_bar_sync(2 + q, 64);  // both column halves of this quadrant start the row block together

      unsigned int* my_tau_g = p.tau_g + (row - lane) + lane;  // == p.tau_g + row (rows are padded to 256)
      for (int n = nb; n < ne; ++n, ++tc) {
        const int as = static_cast<int>(tc % ACC);
        const uint32_t aphase = static_cast<uint32_t>(tc / ACC) & 1u;
        const int n0 = n * BN + half * HALF;
        float* bs = bias_s + (w * ACC + as) * 2 * HALF;
        float* ws = bs + HALF;
        {  // warp-private bias / column-norm slices; columns past the end get -inf so that they can never be admitted
          if (lane * 4 < HALF) {
            float4 bv, wv;
            const int c = n0 + lane * 4;
            bv.x = (c + 0 < p.N) ? __ldg(p.bias + c + 0) : -INFINITY;
            bv.y = (c + 1 < p.N) ? __ldg(p.bias + c + 1) : -INFINITY;
            bv.z = (c + 2 < p.N) ? __ldg(p.bias + c + 2) : -INFINITY;
            bv.w = (c + 3 < p.N) ? __ldg(p.bias + c + 3) : -INFINITY;
            wv.x = (c + 0 < p.N) ? __ldg(p.col_norm + c + 0) : 0.f;
            wv.y = (c + 1 < p.N) ? __ldg(p.col_norm + c + 1) : 0.f;
            wv.z = (c + 2 < p.N) ? __ldg(p.col_norm + c + 2) : 0.f;
            wv.w = (c + 3 < p.N) ? __ldg(p.col_norm + c + 3) : 0.f;
            *reinterpret_cast<float4*>(bs + lane * 4) = bv;
            *reinterpret_cast<float4*>(ws + lane * 4) = wv;
          }
        }
        // thresholds other warps / other CTA pairs have already proven for this row (its other column ranges)
        const unsigned int gkey = __ldcg(my_tau_g);
        __syncwarp();
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        tau = fmaxf(tau, *other_tau_s);
        if (live && gkey != 0u) tau = fmaxf(tau, funkey(gkey));
        if (p.debug >= 1) tau = INFINITY;
        if (p.debug >= 2) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_bar(as), 0);
          continue;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * HALF;

        auto process = [&](uint32_t (&a)[CHUNK], int c) {
          const int col0 = n0 + c * CHUNK;
          float v[CHUNK];
#pragma unroll
          for (int i = 0; i < CHUNK; i += 4) {  // v = t_j = h~_j + ||w_j|| Pb
            const float4 b4 = *reinterpret_cast<const float4*>(bs + c * CHUNK + i);
            const float4 w4 = *reinterpret_cast<const float4*>(ws + c * CHUNK + i);
            v[i] = fmaf(w4.x, Pb, fmaf(__uint_as_float(a[i]), rs, b4.x));
            v[i + 1] = fmaf(w4.y, Pb, fmaf(__uint_as_float(a[i + 1]), rs, b4.y));
            v[i + 2] = fmaf(w4.z, Pb, fmaf(__uint_as_float(a[i + 2]), rs, b4.z));
            v[i + 3] = fmaf(w4.w, Pb, fmaf(__uint_as_float(a[i + 3]), rs, b4.w));
          }
          float gm[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) gm[g] = fmaxf(fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), v[4 * g + 2]), v[4 * g + 3]);
          const float mx = fmaxf(fmaxf(fmaxf(gm[0], gm[1]), gm[2]), gm[3]);
          if (__any_sync(FULL, mx > tau)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (__any_sync(FULL, gm[g] > tau)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const bool hit = v[4 * g + i] > tau;
                  if (__any_sync(FULL, hit)) {  // warp-uniform: usually a single lane of a single column
                    if (hit) {  // store the lower bound l_j = t_j - 2 ||w_j|| Pb - Qb
                      const float lj = fmaf(neg2P, ws[c * CHUNK + 4 * g + i], v[4 * g + i]) - Qb;
                      *wp = make_int2(__float_as_int(lj), col0 + 4 * g + i);
                      ++wp;
                    }
                  }
                }
              }
            }
          }
        };
        uint32_t acc0[CHUNK], acc1[CHUNK];
        tmem_ld_32x32b_x16(taddr, acc0);
#pragma unroll 1
        for (int c = 0; c < HALF / CHUNK; c += 2) {
          tmem_ld_wait_dep(acc0);
          tmem_ld_32x32b_x16(taddr + (c + 1) * CHUNK, acc1);
          process(acc0, c);
          tmem_ld_wait_dep(acc1);
          if (c + 2 < HALF / CHUNK) tmem_ld_32x32b_x16(taddr + (c + 2) * CHUNK, acc0);
          process(acc1, c + 1);
        }
        // the accumulator stage is drained: hand it back BEFORE any list maintenance, so that a slow compaction
        // in one of the 16 epilogue warps of the pair does not hold up the next MMA
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(as), 0);  // the leader's MMA warp owns this barrier

        // A list may grow by at most HALF entries per tile (one per column), so compacting every list that is
        // above CAPG - HALF here guarantees that appends never overflow.
        const int cnt = static_cast<int>(wp - my_buf);
        unsigned need = __ballot_sync(FULL, cnt > p.trigger);
        while (need) {
          const int l = __ffs(need) - 1;
          need &= need - 1;
          const int nn = __shfl_sync(FULL, cnt, l);
          const float mg = __shfl_sync(FULL, margin, l);
          const float fl = __shfl_sync(FULL, tau - shift, l);
          int n_out;
          bool ovf;
          const float thr = compact_list(warp_buf + l * lane_stride, nn, p.top_k, mg, fl, lane, hist_w, n_out, ovf);
          if (lane == l) {
            wp = my_buf + n_out;
            tau = fmaxf(tau, thr + shift);
            overflowed |= ovf;
            *my_tau_s = tau;
            atomicMax(my_tau_g, fkey(tau));
          }
        }
      }

      // publish this (row block, range, half): trim each list to the margin band of its k-th largest
      const int cnt_end = static_cast<int>(wp - my_buf);
      for (int l = 0; l < 32; ++l) {
        const int nn = __shfl_sync(FULL, cnt_end, l);
        const float mg = __shfl_sync(FULL, margin, l);
        const float fl = __shfl_sync(FULL, tau - shift, l);
        const int grow = row - lane + l;
        if (grow >= p.M) continue;  // warp-uniform
        int n_out = nn;
        bool ovf = false;
        if (nn > p.top_k) compact_list(warp_buf + l * lane_stride, nn, p.top_k, mg, fl, lane, hist_w, n_out, ovf);
        if (lane == l) {
          overflowed |= ovf;
          p.cand_cnt[static_cast<long long>(grow) * p.nlists + list] = overflowed ? -n_out : n_out;
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}
static int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return 1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

}  // namespace g2

// the lists keep the k largest plus the margin band within CAPG / 2 = 192 entries, and 2 k <= TRIGGER_MAX
int encode2_max_top_k() { return 128; }

// Number of CTA pairs that can be co-resident (one per TPC with both SMs free); 0 on failure.
int encode2_max_pairs() {
  static int cached = -1;
  if (cached >= 0) return cached;
  if (cudaFuncSetAttribute(g2::encode_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(g2::SMEM_TOTAL)) != cudaSuccess) {
    cudaGetLastError();
    return cached = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 148);
  cfg.blockDim = dim3(g2::NUM_THREADS);
  cfg.dynamicSmemBytes = g2::SMEM_TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, g2::encode_gemm2_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return cached = n;
}

Encode2Plan encode2_plan(int M, int N, int max_pairs) {
  Encode2Plan pl;
  pl.m_pairs = (M + 2 * g2::BM - 1) / (2 * g2::BM);
  pl.n_tiles = (N + g2::BN - 1) / g2::BN;
  const long long total = static_cast<long long>(pl.m_pairs) * pl.n_tiles;
  if (max_pairs < 1) max_pairs = 1;
  long long q = (total + max_pairs - 1) / max_pairs;
  // at most ENCODE_MAX_NSPLIT ranges may touch one row block
  const long long q_min = (pl.n_tiles + ENCODE_MAX_NSPLIT - 2) / (ENCODE_MAX_NSPLIT - 1);
  if (q < q_min) q = q_min;
  if (q < 1) q = 1;
  pl.q = static_cast<int>(q);
  pl.n_pairs = static_cast<int>((total + q - 1) / q);
  int nsplit = 1;
  for (int m = 0; m < pl.m_pairs; ++m) {
    const long long first = (static_cast<long long>(m) * pl.n_tiles) / q;
    const long long last = (static_cast<long long>(m + 1) * pl.n_tiles - 1) / q;
    nsplit = max(nsplit, static_cast<int>(last - first + 1));
  }
  pl.nsplit = nsplit;
  pl.nlists = 2 * nsplit;
  return pl;
}

int launch_encode_gemm2(const EncodeGemmArgs& a, const Encode2Plan& pl, cudaStream_t stream) {
  using namespace g2;
  if (a.M <= 0 || a.N <= 0) return 0;
  if ((a.K % 8) != 0) return 10;
  if (a.top_k <= 0 || a.top_k > encode2_max_top_k() || !a.cand || !a.cand_cnt || !a.row_norm || !a.row_dx || !a.row_scale || !a.scalars ||
      !a.col_norm)
    return 12;
  CUtensorMap maps[2];
  if (make_tmap(&maps[0], a.A_hi, a.M, a.K, BM)) return 11;
  if (make_tmap(&maps[1], a.B_hi, a.N, a.K, BN / 2)) return 11;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(encode_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(SMEM_TOTAL)) != cudaSuccess)
      return 3;
    attr_set = true;
  }
  Params p;
  p.kblocks = (a.K + BK - 1) / BK;
  p.M = a.M;
  p.N = a.N;
  p.n_tiles = pl.n_tiles;
  p.total = static_cast<long long>(pl.m_pairs) * pl.n_tiles;
  p.q = pl.q;
  p.nlists = pl.nlists;
  p.top_k = a.top_k;
  {
    // Lists are compacted (and the row's admission threshold tightened) after every tile that leaves them above
    // `trigger`.  Appends made under a stale threshold cost more than the compaction, so the trigger sits well
    // below the capacity bound; it must leave room for k entries plus the margin band.
    const char* e = getenv("SAEV_B200_TRIGGER");
    const int trig = e ? atoi(e) : 192;
    p.trigger = max(2 * a.top_k, min(trig, TRIGGER_MAX));
  }
  p.bias = a.bias;
  p.row_norm = a.row_norm;
  p.row_dx = a.row_dx;
  p.row_scale = a.row_scale;
  p.scalars = a.scalars;
  p.col_norm = a.col_norm;
  p.D = a.K;
  p.cand = reinterpret_cast<int2*>(a.cand);
  p.cand_cnt = a.cand_cnt;
  p.tau_g = a.tau_keys;
  if (!a.tau_keys) return 12;
  {
    const char* e = getenv("SAEV_B200_SCREEN_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }

  // (debug 3, timing experiment: keep the thresholds of the previous launch = a perfect warm start when the same batch
  //  is screened again)
  if (p.debug != 3 && !a.tau_preset &&
      cudaMemsetAsync(a.tau_keys, 0, static_cast<size_t>(pl.m_pairs) * 2 * BM * 4, stream) != cudaSuccess)
    return 23;
  // lists that no range covers for a given row block must read as empty
  if (cudaMemsetAsync(a.cand_cnt, 0, static_cast<size_t>(pl.m_pairs) * 2 * BM * pl.nlists * 4, stream) != cudaSuccess)
    return 23;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pl.n_pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, encode_gemm2_kernel, maps[0], maps[1], p) != cudaSuccess) return 4;
  ++g_launch_count;
  return 0;
}

}  // namespace sb
