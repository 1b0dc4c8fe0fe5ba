// Dense split-product contractions on CTA pairs:  out = A . B^T  with A, B held as 2 or 3 bf16 pieces (hi, lo[, lo2]),
// tcgen05.mma.cta_group::2 on 256 x 256 tiles, fp32 accumulation in TMEM, the epilogues of dense_epilogue.cuh.
//
// The dense (ReLU) path of the SAE step (saev src/saev/nn/modeling.py:150-156, 343-409 and their autograd) is five such
// contractions.  On the single-CTA kernel (encode_gemm.cu) every TERM of the error-compensated product walks its own
// pair of operand pieces through the smem ring: 48 KB of TMA loads per k-block per term, 113 GB of L2->SM reads per
// contraction at the c5 shape -- the L2->SM path (about 43 GB/s per SM) is the bound, every contraction sits at
// 0.35-0.39 of the tensor peak.  Two changes here raise the FLOPs per loaded byte 2.25x (3 terms) / 3x (6 terms):
//
//   * CTA pairs (as the top-k screen, encode_gemm2.cu): a pair computes a 256 x 256 tile, each CTA stages only ITS 128
//     rows of A and ITS 128 of the 256 columns of B; the tensor cores of both SMs read both halves.
//   * pieces staged ONCE per k-block: a stage holds this CTA's slices of ALL pieces of A and B (2 x NP x 16 KB) and the
//     MMA warp issues every term of the split product from them -- hi.hi, hi.lo, lo.hi (NP = 2), plus hi.lo2, lo2.hi,
//     lo.lo (NP = 3) -- instead of re-loading the hi pieces for every term.
//
// Roles (384 threads): warp 0 = TMA producer (both CTAs), warp 1 = MMA issuer (leader CTA, one lane), warp 2 = TMEM
// allocator, warps 4-11 = epilogue (TMEM lane quadrant = warp % 4, column half = (warp - 4) / 4).  Barriers as in the
// screen kernel: full/empty per smem stage (full lives in the leader, both CTAs' TMA bytes complete on it; empty is
// multicast to both CTAs by tcgen05.commit), tfull (multicast commit) / tempty (leader, 16 arrivals) per accumulator
// stage.  The K chunking of the dense-store / weight-gradient epilogues (bounded truncation bias of the tensor-core
// accumulator, see encode_gemm.cu) is kept: one accumulator stage per chunk, folded into the output with vector REDs.
//
// Handles epilogues 1-4 with static problem sizes and optional (m, n, k) windows; the device-sized AuxK contractions,
// K splits and the coherence screen stay on the single-CTA kernel.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "dense_epilogue.cuh"
#include "pair_common.cuh"

namespace sb {
namespace dg2 {
using namespace pairx;

constexpr int BM = 128;        // rows per CTA (pair: 256 = UMMA M)
constexpr int BN = 256;        // columns per tile (UMMA N)
constexpr int ACC = 2;         // TMEM accumulator stages (2 x 256 columns)
constexpr int BK = 64;         // bf16 per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 16;
constexpr int PIECE_BYTES = BM * BK * 2;  // 16 KB: 128 rows (of A) or 128 columns (of B) of one piece
constexpr int CHUNK = EPI_CHUNK;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
constexpr int HALF = BN / 2;

template <int NP>
struct Cfg {
  static constexpr int STAGE_BYTES = 2 * NP * PIECE_BYTES;          // 64 KB (NP = 2) / 96 KB (NP = 3)
  static constexpr int STAGES = NP == 2 ? 3 : 2;
  static constexpr size_t OFF_BIAS = static_cast<size_t>(STAGES) * STAGE_BYTES;  // [8 warps][ACC][128] f32
  static constexpr size_t OFF_BARS = OFF_BIAS + static_cast<size_t>(EPI_WARPS) * ACC * HALF * 4;
  static constexpr size_t SMEM_TOTAL = OFF_BARS + (2 * STAGES + 2 * ACC) * 8 + 16 + 1024;
};

struct Params {
  int kblocks;        // k-blocks of the contraction window
  int kchunk;         // k-blocks per accumulation chunk
  int M, N;           // absolute extents (rows / columns at or beyond them are not stored)
  int n_tiles;        // column tiles of the window
  long long total;    // m_pairs * n_tiles
  int q;              // (row pair, tile) units per CTA pair
  const float* bias;
  float* out;
  long long ldo;
  EpiExtra ex;
};

template <int EPI, int NP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
dense_gemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB0,
                   const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const Params p) {
  using C = Cfg<NP>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_u32 + pad;

  float* bias_s = reinterpret_cast<float*>(smem + C::OFF_BIAS);
  const uint32_t bars = smem_base + static_cast<uint32_t>(C::OFF_BARS);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + ACC + s); };
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + C::OFF_BARS + (2 * STAGES + 2 * ACC) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;

  const long long u_begin = static_cast<long long>(pair) * p.q;
  const long long u_end = min(p.total, u_begin + p.q);
  const int n_chunks = (p.kblocks + p.kchunk - 1) / p.kchunk;
  const EpiExtra& ex = p.ex;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB0);
    tma_prefetch_desc(&tmB1);
    if (NP == 3) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
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
      for (long long u = u_begin; u < u_end; ++u) {
        const int m_pair = static_cast<int>(u / p.n_tiles), n = static_cast<int>(u - static_cast<long long>(m_pair) * p.n_tiles);
        const int row0 = ex.m_begin + m_pair * (2 * BM) + static_cast<int>(rank) * BM;
        const int col0 = ex.n_begin + n * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          const int k0 = ex.k_begin + kb * BK;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          tma_load_2d_pair(sa, &tmA0, full_bar(stage), k0, row0);
          tma_load_2d_pair(sa + PIECE_BYTES, &tmA1, full_bar(stage), k0, row0);
          if (NP == 3) tma_load_2d_pair(sa + 2 * PIECE_BYTES, &tmA2, full_bar(stage), k0, row0);
          const uint32_t sb_ = sa + NP * PIECE_BYTES;
          tma_load_2d_pair(sb_, &tmB0, full_bar(stage), k0, col0);
          tma_load_2d_pair(sb_ + PIECE_BYTES, &tmB1, full_bar(stage), k0, col0);
          if (NP == 3) tma_load_2d_pair(sb_ + 2 * PIECE_BYTES, &tmB2, full_bar(stage), k0, col0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();  // reconverge before the (warp-aligned) cluster barrier at the end
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(2 * BM, BN);
      // (A piece, B piece) per term: hi.hi, hi.lo, lo.hi [, hi.lo2, lo2.hi, lo.lo]
      constexpr int NT = NP == 2 ? 3 : 6;
      constexpr int TA[6] = {0, 0, 1, 0, 2, 1};
      constexpr int TB[6] = {0, 1, 0, 2, 0, 1};
      int stage = 0;
      uint32_t phase = 0;
      long long tc = 0;
      for (long long u = u_begin; u < u_end; ++u) {
        for (int kc = 0; kc < n_chunks; ++kc, ++tc) {
          const int as = static_cast<int>(tc % ACC);
          const uint32_t aphase = static_cast<uint32_t>(tc / ACC) & 1u;
          const int clen = min(p.kchunk, p.kblocks - kc * p.kchunk);
          mbar_wait(tempty_bar(as), aphase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kb = 0; kb < clen; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
            uint64_t adesc[NP], bdesc[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              adesc[i] = umma_desc_kmajor_sw128(sa + i * PIECE_BYTES);
              bdesc[i] = umma_desc_kmajor_sw128(sa + (NP + i) * PIECE_BYTES);
            }
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
#pragma unroll
              for (int t = 0; t < NT; ++t)
                umma_f16_pair(d_tmem, adesc[TA[t]] + 2u * k, bdesc[TB[t]] + 2u * k, idesc, (kb | k | t) != 0);
            }
            umma_commit_pair(empty_bar(stage));
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit_pair(tfull_bar(as));
        }
      }
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (both CTAs) =====================
    const int w = warp - EPI_WARP0;
    const int q = warp & 3;   // TMEM lane quadrant this warp may read
    const int half = w >> 2;  // low / high 128 columns of every tile
    const int row_local = q * 32 + lane;
    const int n_cols = p.N;
    float acc_l1 = 0.f, acc_l0 = 0.f;  // EPI 2: per-row partials over the tiles of one row pair
    int acc_row = -1;
    float best = -1.f;                 // (EPI 5 is not instantiated here)
    int best_col = -1;
    long long tc = 0;
    for (long long u = u_begin; u < u_end; ++u) {
      const int m_pair = static_cast<int>(u / p.n_tiles), n = static_cast<int>(u - static_cast<long long>(m_pair) * p.n_tiles);
      const int row = ex.m_begin + m_pair * (2 * BM) + static_cast<int>(rank) * BM + row_local;
      const int n0 = ex.n_begin + n * BN + half * HALF;
      if (EPI == 2 && row != acc_row) {
        if (acc_row >= 0 && acc_row < p.M) {
          atomicAdd(ex.row_l1 + acc_row, acc_l1);
          atomicAdd(ex.row_l0 + acc_row, acc_l0);
        }
        acc_l1 = acc_l0 = 0.f;
        acc_row = row;
      }
      for (int kc = 0; kc < n_chunks; ++kc, ++tc) {
        const bool accum = kc > 0;
        const int as = static_cast<int>(tc % ACC);
        const uint32_t aphase = static_cast<uint32_t>(tc / ACC) & 1u;
        float* bs = bias_s + (w * ACC + as) * HALF;
        if (lane * 4 < HALF) {  // warp-private bias slice of this tile's column half (zero for the later K chunks)
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          const int c = n0 + lane * 4;
          if (p.bias != nullptr && !accum) {
            bv.x = (c + 0 < n_cols) ? __ldg(p.bias + c + 0) : 0.f;
            bv.y = (c + 1 < n_cols) ? __ldg(p.bias + c + 1) : 0.f;
            bv.z = (c + 2 < n_cols) ? __ldg(p.bias + c + 2) : 0.f;
            bv.w = (c + 3 < n_cols) ? __ldg(p.bias + c + 3) : 0.f;
          }
          *reinterpret_cast<float4*>(bs + lane * 4) = bv;
        }
        __syncwarp();
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * HALF;
        auto process = [&](uint32_t (&a)[CHUNK], int c) {
          dense_epi_chunk<EPI>(ex, a, bs + c * CHUNK, row, n0 + c * CHUNK, p.M, n_cols, accum, p.out, p.ldo, lane, acc_l1,
                               acc_l0, best, best_col);
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(as), 0);  // the leader's MMA warp owns this barrier
      }
    }
    if (EPI == 2 && acc_row >= 0 && acc_row < p.M) {
      atomicAdd(ex.row_l1 + acc_row, acc_l1);
      atomicAdd(ex.row_l0 + acc_row, acc_l0);
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
// bf16 row-major [rows, cols] matrix with row pitch ld, box = [128 rows, 64 cols], 128-byte swizzle, OOB -> zeros
static int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return 1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

template <int EPI, int NP>
static int launch_variant(const EncodeGemmArgs& a, const CUtensorMap* mA, const CUtensorMap* mB, cudaStream_t stream) {
  using C = Cfg<NP>;
  auto kern = dense_gemm2_kernel<EPI, NP>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(C::SMEM_TOTAL)) != cudaSuccess)
      return 3;
    attr_set = true;
  }
  Params p;
  p.kblocks = (a.K - a.k_begin + BK - 1) / BK;
  p.kchunk = ((EPI == 1 || EPI == 4) && a.k_chunk_blocks > 0) ? a.k_chunk_blocks : p.kblocks;
  p.M = a.M;
  p.N = a.N;
  const int m_pairs = (a.M - a.m_begin + 2 * BM - 1) / (2 * BM);
  p.n_tiles = (a.N - a.n_begin + BN - 1) / BN;
  p.total = static_cast<long long>(m_pairs) * p.n_tiles;
  int max_pairs = a.num_sms / 2;
  if (max_pairs < 1) max_pairs = 1;
  long long q = (p.total + max_pairs - 1) / max_pairs;
  if (q < 1) q = 1;
  p.q = static_cast<int>(q);
  const int n_pairs = static_cast<int>((p.total + q - 1) / q);
  p.bias = a.bias;
  p.out = a.out;
  p.ldo = a.ldo;
  EpiExtra& ex = p.ex;
  ex.f_hi = a.f_hi; ex.f_lo = a.f_lo; ex.t_hi = a.t_hi; ex.t_lo = a.t_lo; ex.ldf = a.ldf; ex.ldt = a.ldt;
  ex.f_lo2 = a.f_lo2; ex.t_lo2 = a.t_lo2;
  ex.row_l1 = a.row_l1; ex.row_l0 = a.row_l0; ex.active = a.active; ex.l1_over_b = a.l1_over_b;
  ex.n_main = a.n_main; ex.extra = a.extra;
  ex.m_limit_dev = nullptr; ex.k_limit_dev = nullptr; ex.row_map = a.row_map; ex.alpha = a.alpha;
  ex.ksplit = 1;
  ex.m_begin = a.m_begin; ex.n_begin = a.n_begin; ex.k_begin = a.k_begin;

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * n_pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, mA[0], mA[1], mA[2], mB[0], mB[1], mB[2], p) != cudaSuccess) return 4;
  ++g_launch_count;
  return 0;
}

}  // namespace dg2

// Whether launch_dense_gemm2 takes this contraction (otherwise the single-CTA kernel of encode_gemm.cu runs it).
bool dense_gemm2_eligible(const EncodeGemmArgs& a) {
  static const bool on = [] { const char* v = getenv("SAEV_B200_DENSE_PAIR"); return !(v && v[0] == '0'); }();
  if (!on) return false;
  if (a.nterms != 3 && a.nterms != 6) return false;
  if (a.epilogue < 1 || a.epilogue > 4) return false;
  if (a.m_limit_dev || a.n_limit_dev || a.k_limit_dev || a.ksplit > 1) return false;
  if (!a.A_lo || !a.B_lo || (a.nterms == 6 && (!a.A_lo2 || !a.B_lo2))) return false;
  return true;
}

int launch_dense_gemm2(const EncodeGemmArgs& a, cudaStream_t stream) {
  using namespace dg2;
  if (a.M <= a.m_begin || a.N <= a.n_begin || a.K <= a.k_begin) return 0;
  const long long lda = a.lda > 0 ? a.lda : a.K, ldb = a.ldb > 0 ? a.ldb : a.K;
  if ((lda % 8) != 0 || (ldb % 8) != 0 || (a.k_begin % 8) != 0) return 10;
  const int np = a.nterms == 6 ? 3 : 2;
  CUtensorMap mA[3], mB[3];
  const void* pa[3] = {a.A_hi, a.A_lo, a.A_lo2};
  const void* pb[3] = {a.B_hi, a.B_lo, a.B_lo2};
  for (int i = 0; i < 3; ++i) {
    const int s = i < np ? i : 0;  // (unused third maps alias the first)
    if (make_tmap(&mA[i], pa[s], a.M, a.K, lda)) return 11;
    if (make_tmap(&mB[i], pb[s], a.N, a.K, ldb)) return 11;
  }
#define SB_DG2(E)                                                                    \
  return np == 2 ? launch_variant<E, 2>(a, mA, mB, stream) : launch_variant<E, 3>(a, mA, mB, stream);
  switch (a.epilogue) {
    case 1: SB_DG2(1)
    case 2: SB_DG2(2)
    case 3: SB_DG2(3)
    case 4: SB_DG2(4)
    default: return 12;
  }
#undef SB_DG2
}

}  // namespace sb
