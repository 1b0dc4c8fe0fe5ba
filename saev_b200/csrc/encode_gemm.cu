// Dense contractions on tcgen05 tensor cores, single-CTA version:  out = A . B^T (+ bias), bf16 operand pieces, fp32
// accumulate in TMEM.
//
// One kernel, five fused epilogues (template parameter EPI, see dense_epilogue.cuh).  launch_encode_gemm is the entry
// for every dense contraction of the library; it forwards epilogues 1-4 with static problem sizes -- the dense (ReLU)
// path of the SAE step (saev src/saev/nn/modeling.py:150-156, 343-409 and their autograd) and the saev_b200_gemm_nt test
// hook -- to the CTA-pair kernel of dense_gemm2.cu (SAEV_B200_DENSE_PAIR=0 keeps them here) and runs the rest itself: the
// AuxK contractions over the dead latents, whose sizes live on the device (modeling.py:75-103), K splits, and the
// dictionary-coherence screen of the log block (train.py:411-417).  The TopK screen of the encoder contraction is the
// CTA-pair kernel in encode_gemm2.cu.
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator,
// warps 4-7 = epilogue (TMEM lane quadrant = warp_idx % 4).  Pipelines: STAGES-deep smem ring (TMA <-> MMA),
// 2-deep TMEM accumulator ring (MMA <-> epilogue), so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operands are K-major bf16 with 128-byte swizzle.  `nterms == 3` runs the error-compensated split product
// (x_hi.W_hi + x_hi.W_lo + x_lo.W_hi, ~2^-17 relative) by walking three (A,B) tensor-map pairs along K; `nterms == 6`
// adds a third piece per operand (hi + lo + lo2 = 24 bits: fp32-class accuracy, terms hi.hi, hi.lo, lo.hi, hi.lo2,
// lo2.hi, lo.lo).  Optional (m, n, k) windows (Matryoshka prefix blocks): see EpiExtra.
#include <stdio.h>

#include "common.cuh"
#include "kernels.h"
#include "dense_epilogue.cuh"

namespace sb {

constexpr int BM = 128;       // rows of the batch per CTA (UMMA M)
constexpr int BN = 256;       // dictionary columns per tile (UMMA N)
constexpr int BK = 64;        // bf16 elements per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int CHUNK = 16;     // accumulator columns per tcgen05.ld
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;

struct EncodeSmemLayout {
  int stages;
  size_t off_bias, off_bars, total;
};

__host__ __device__ inline EncodeSmemLayout encode_smem_layout(int stages) {
  EncodeSmemLayout L;
  L.stages = stages;
  size_t o = static_cast<size_t>(stages) * STAGE_BYTES;
  L.off_bias = o;
  o += 2 * BN * 4;
  L.off_bars = (o + 7) & ~size_t(7);
  o = L.off_bars + (2 * stages + 4) * 8 + 16;
  L.total = o + 1024;  // slack for the manual 1024-byte alignment of the dynamic smem base
  return L;
}

// EPI: 1 = dense fp32 store of (acc + bias)
//      2 = ReLU forward (dense SAE path): f = relu(acc + bias) written as a bf16 hi/lo pair both row-major
//          [M, ldf] (operand of the decoder contraction) and transposed [N, ldt] (operand of the weight-gradient
//          contraction); per-row sum f / count f>0, per-column "fired" flags
//      3 = ReLU backward: dh = (f > 0) ? acc + l1_over_b : 0, written transposed as a bf16 hi/lo pair [N, ldt]
//      4 = weight gradient: out[row, col] = acc for col < n_main, extra[row] = acc for col == n_main
//      5 = dictionary coherence screen (A == B == unit rows of W_dec): per (row, split) the largest |acc| over the
//          columns col > row and its column, into extra / active [M, nsplit]; tiles wholly below the diagonal are
//          skipped by all three roles
template <int EPI, int CAPG, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
encode_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   const __grid_constant__ CUtensorMap tmA_lo2, const __grid_constant__ CUtensorMap tmB_lo2,
                   int nterms, int kblocks_per_term, int kchunk, const float* __restrict__ bias, int M, int N, int m_blocks,
                   int tiles_per_split, int nsplit, const int* __restrict__ n_limit_dev, float* __restrict__ out,
                   long long ldo, const EpiExtra ex) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_u32 + pad;

  const EncodeSmemLayout L = encode_smem_layout(STAGES);
  float* bias_s = reinterpret_cast<float*>(smem + L.off_bias);
  const uint32_t bars = smem_base + static_cast<uint32_t>(L.off_bars);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + L.off_bars + (2 * STAGES + 4) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blk = blockIdx.x % m_blocks;
  const int split = (blockIdx.x / m_blocks) % nsplit;
  const int ks = blockIdx.x / (m_blocks * nsplit);  // K split (dense epilogues 1 / 4 with RED accumulation)
  // The column count may live on the device (AuxK dead-latent list whose length the host never reads).
  const int n_cols = n_limit_dev ? min(N, *n_limit_dev) : N;
  if (ex.m_limit_dev != nullptr) M = min(M, *ex.m_limit_dev);
  if (ex.k_limit_dev != nullptr) kblocks_per_term = min(kblocks_per_term, (*ex.k_limit_dev + BK - 1) / BK);
  // nothing to do (no dead latents / rows past the dynamic row count): leave before any barrier or TMEM is touched
  if (n_cols <= ex.n_begin || kblocks_per_term <= 0 || ex.m_begin + (blockIdx.x % m_blocks) * BM >= M) return;
  const int n_tiles_total = (n_cols - ex.n_begin + BN - 1) / BN;
  int tile_begin = split * tiles_per_split;
  const int tile_end = min(n_tiles_total, tile_begin + tiles_per_split);
  if (EPI == 5) tile_begin = max(tile_begin, (m_blk * BM + 1) / BN);  // first tile holding a column > row
  const int num_tiles = max(0, tile_end - tile_begin);
  // K chunking (dense epilogues 1 / 4): the contraction of one output tile is cut into chunks of `kchunk` k-blocks per
  // term, each accumulated in its own TMEM stage and added to the fp32 output by the epilogue.  Tensor-core
  // accumulation truncates (rounds toward zero) at every MMA, a bias that grows linearly with the number of
  // accumulator updates; short chunks + round-to-nearest fp32 adds keep it at the 1e-5 level for any K.
  const int n_chunks_all = (kblocks_per_term + kchunk - 1) / kchunk;
  // with ex.ksplit > 1 the chunks of one output tile are spread over ksplit CTAs, all of which ADD into a
  // pre-zeroed output (the contraction is long and the tile count small: AuxK weight gradients of a few dead atoms)
  const int chunks_per_ks = (n_chunks_all + ex.ksplit - 1) / ex.ksplit;
  const int chunk_begin = ks * chunks_per_ks;
  const int n_chunks = max(0, min(n_chunks_all, chunk_begin + chunks_per_ks) - chunk_begin);
  const int num_vtiles = num_tiles * n_chunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (nterms > 3) {
      tma_prefetch_desc(&tmA_lo2);
      tma_prefetch_desc(&tmB_lo2);
    }
    if (nterms > 1) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_ptr_s));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int vt = 0; vt < num_vtiles; ++vt) {
        const int t = vt / n_chunks, kc = chunk_begin + vt - t * n_chunks;
        const int n0 = ex.n_begin + (tile_begin + t) * BN;
        const int clen = min(kchunk, kblocks_per_term - kc * kchunk);
        for (int kb = 0; kb < nterms * clen; ++kb) {
          const int term = kb / clen;
          const int k0 = ex.k_begin + (kc * kchunk + kb - term * clen) * BK;
          // (A piece, B piece) per term: 0:(hi,hi) 1:(hi,lo) 2:(lo,hi) 3:(hi,lo2) 4:(lo2,hi) 5:(lo,lo)
          const CUtensorMap* ma = (term == 2 || term == 5) ? &tmA_lo : (term == 4 ? &tmA_lo2 : &tmA_hi);
          const CUtensorMap* mb = (term == 1 || term == 5) ? &tmB_lo : (term == 3 ? &tmB_lo2 : &tmB_hi);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          tma_load_2d(sa, ma, full_bar(stage), k0, ex.m_begin + m_blk * BM);
          tma_load_2d(sa + A_BYTES, mb, full_bar(stage), k0, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int vt = 0; vt < num_vtiles; ++vt) {
        const int as = vt & 1;
        const uint32_t aphase = (vt >> 1) & 1u;
        const int kc = chunk_begin + vt % n_chunks;
        const int kblocks = nterms * min(kchunk, kblocks_per_term - kc * kchunk);
        mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = umma_desc_kmajor_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing K inside the 128-byte swizzle span = +32 bytes on the start address (>>4 => +2)
            umma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row_local = q * 32 + lane;    // accumulator row == TMEM lane
    const int row = ex.m_begin + m_blk * BM + row_local;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127
    float acc_l1 = 0.f, acc_l0 = 0.f;  // EPI 2
    float best = -1.f;                 // EPI 5
    int best_col = -1;

    for (int vt = 0; vt < num_vtiles; ++vt) {
      const int t = vt / n_chunks;
      // later K chunks of the same output tile are added to it; with a K split every chunk is added (pre-zeroed output)
      const bool accum = (vt - t * n_chunks) > 0 || ex.ksplit > 1;
      const int as = vt & 1;
      const uint32_t aphase = (vt >> 1) & 1u;
      const int n0 = ex.n_begin + (tile_begin + t) * BN;
      float* bs = bias_s + as * BN;
      // Stage the bias slice of this tile (double buffered by accumulator stage; see barrier note below).
      for (int c = et; c < BN; c += 128) {
        float bv = 0.f;  // (columns past the end are never stored)
        if (n0 + c < n_cols) bv = (bias != nullptr && !accum && ks == 0) ? bias[n0 + c] : 0.f;
        bs[c] = bv;
      }
      named_bar_sync(1, 128);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;

      // Two register buffers, unrolled by hand so the accumulator stays in registers: while chunk c is
      // being filtered the tcgen05.ld of chunk c+1 is in flight.
      auto process = [&](uint32_t (&a)[CHUNK], int c) {
        dense_epi_chunk<EPI>(ex, a, bs + c * CHUNK, row, n0 + c * CHUNK, M, n_cols, accum, out, ldo, lane, acc_l1, acc_l0, best,
                             best_col);
      };
      uint32_t acc0[CHUNK], acc1[CHUNK];
      tmem_ld_32x32b_x16(taddr, acc0);
#pragma unroll 1
      for (int c = 0; c < BN / CHUNK; c += 2) {
        tmem_ld_wait_dep(acc0);
        tmem_ld_32x32b_x16(taddr + (c + 1) * CHUNK, acc1);
        process(acc0, c);
        tmem_ld_wait_dep(acc1);
        if (c + 2 < BN / CHUNK) tmem_ld_32x32b_x16(taddr + (c + 2) * CHUNK, acc0);
        process(acc1, c + 1);
      }
      // accumulator stage drained -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      // Bias double-buffering: a warp can only reach the fill of tile t+2 (same buffer) after the
      // named barrier of tile t+1, which every warp reaches only after finishing tile t.
    }

    if (EPI == 5 && row < M) {
      ex.extra[static_cast<long long>(row) * nsplit + split] = best;
      ex.active[static_cast<long long>(row) * nsplit + split] = best_col;
    }
    if (EPI == 2 && row < M && num_tiles > 0) {
      atomicAdd(ex.row_l1 + row, acc_l1);
      atomicAdd(ex.row_l0 + row, acc_l0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// bf16 row-major [rows, cols] matrix, box = [box_rows, 64 cols], 128-byte swizzle, OOB -> zeros.
static int make_tmap_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld_elems,
                          int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return 1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fprintf(stderr, "saev_b200: cuTensorMapEncodeTiled failed (CUresult %d): ptr=%p rows=%lld cols=%lld ld=%lld box_rows=%d\n",
            static_cast<int>(r), ptr, rows, cols, ld_elems, box_rows);
  return r == CUDA_SUCCESS ? 0 : 2;
}

template <int EPI, int CAPG, int STAGES>
static int launch_variant(const EncodeGemmArgs& a, const CUtensorMap* maps, int m_blocks, int tiles_per_split,
                          int nsplit, cudaStream_t stream) {
  auto kern = encode_gemm_kernel<EPI, CAPG, STAGES>;
  const EncodeSmemLayout L = encode_smem_layout(STAGES);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total)) !=
        cudaSuccess)
      return 3;
    attr_set = true;
  }
  const int kblocks_per_term = (a.K - a.k_begin + BK - 1) / BK;
  EpiExtra ex;
  ex.m_begin = a.m_begin; ex.n_begin = a.n_begin; ex.k_begin = a.k_begin;
  ex.f_hi = a.f_hi; ex.f_lo = a.f_lo; ex.t_hi = a.t_hi; ex.t_lo = a.t_lo; ex.ldf = a.ldf; ex.ldt = a.ldt;
  ex.f_lo2 = a.f_lo2; ex.t_lo2 = a.t_lo2;
  ex.row_l1 = a.row_l1; ex.row_l0 = a.row_l0; ex.active = a.active; ex.l1_over_b = a.l1_over_b;
  ex.n_main = a.n_main; ex.extra = a.extra;
  ex.m_limit_dev = a.m_limit_dev; ex.k_limit_dev = a.k_limit_dev; ex.row_map = a.row_map; ex.alpha = a.alpha;
  ex.ksplit = ((EPI == 1 || EPI == 4) && a.ksplit > 1 && a.k_chunk_blocks > 0) ? a.ksplit : 1;
  // K chunking only where the epilogue can add partial results (dense store / weight gradient); 8 k-blocks = 512
  // bf16 per term per chunk
  const int kchunk = ((EPI == 1 || EPI == 4) && a.k_chunk_blocks > 0) ? a.k_chunk_blocks : kblocks_per_term;
  kern<<<m_blocks * nsplit * ex.ksplit, NUM_THREADS, L.total, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], a.nterms,
                                                           kblocks_per_term, kchunk, a.bias, a.M, a.N, m_blocks,
                                                           tiles_per_split, nsplit, a.n_limit_dev, a.out, a.ldo, ex);
                                                           ++g_launch_count;
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}

int encode_gemm_nsplit(int M, int N, int num_sms) {
  const int m_blocks = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  int nsplit = num_sms / m_blocks;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > n_tiles) nsplit = n_tiles;
  if (nsplit > ENCODE_MAX_NSPLIT) nsplit = ENCODE_MAX_NSPLIT;
  // re-balance so that no split is empty
  const int tps = (n_tiles + nsplit - 1) / nsplit;
  return (n_tiles + tps - 1) / tps;
}

int launch_encode_gemm(const EncodeGemmArgs& a, cudaStream_t stream) {
  if (dense_gemm2_eligible(a)) return launch_dense_gemm2(a, stream);
  if (a.M <= a.m_begin || a.N <= a.n_begin || a.K <= a.k_begin) return 0;
  if (a.m_begin < 0 || a.n_begin < 0 || a.k_begin < 0) return 12;
  if ((a.m_begin || a.n_begin || a.k_begin) && (a.m_limit_dev || a.n_limit_dev || a.k_limit_dev || a.epilogue == 5)) return 12;
  if (a.lda <= 0 && a.ldb <= 0 && (a.K % 8) != 0) return 10;  // TMA needs 16-byte aligned row pitch
  CUtensorMap maps[6];
  const long long lda = a.lda > 0 ? a.lda : a.K, ldb = a.ldb > 0 ? a.ldb : a.K;
  if ((lda % 8) != 0 || (ldb % 8) != 0) return 10;
  if (make_tmap_bf16(&maps[0], a.A_hi, a.M, a.K, lda, BM)) return 11;
  if (make_tmap_bf16(&maps[2], a.B_hi, a.N, a.K, ldb, BN)) return 11;
  if (a.nterms == 6) {
    if (!a.A_lo2 || !a.B_lo2) return 12;
    if (make_tmap_bf16(&maps[4], a.A_lo2, a.M, a.K, lda, BM)) return 11;
    if (make_tmap_bf16(&maps[5], a.B_lo2, a.N, a.K, ldb, BN)) return 11;
  } else {
    maps[4] = maps[0];
    maps[5] = maps[2];
  }
  if (a.nterms >= 3) {
    if (make_tmap_bf16(&maps[1], a.A_lo, a.M, a.K, lda, BM)) return 11;
    if (make_tmap_bf16(&maps[3], a.B_lo, a.N, a.K, ldb, BN)) return 11;
  } else {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  const int m_blocks = (a.M - a.m_begin + BM - 1) / BM;
  const int n_tiles = (a.N - a.n_begin + BN - 1) / BN;
  if (a.epilogue >= 1) {
    // dense epilogues: any column split works; use enough CTAs to fill the GPU
    int nsplit = a.nsplit > 0 ? a.nsplit : encode_gemm_nsplit(a.M - a.m_begin, a.N - a.n_begin, a.num_sms);
    const int tps = (n_tiles + nsplit - 1) / nsplit;
    nsplit = (n_tiles + tps - 1) / tps;
    switch (a.epilogue) {
      case 1: return launch_variant<1, ENCODE_CAPG, 4>(a, maps, m_blocks, tps, nsplit, stream);
      case 2: return launch_variant<2, ENCODE_CAPG, 4>(a, maps, m_blocks, tps, nsplit, stream);
      case 3: return launch_variant<3, ENCODE_CAPG, 4>(a, maps, m_blocks, tps, nsplit, stream);
      case 4: return launch_variant<4, ENCODE_CAPG, 4>(a, maps, m_blocks, tps, nsplit, stream);
      case 5: return launch_variant<5, ENCODE_CAPG, 4>(a, maps, m_blocks, tps, nsplit, stream);
      default: return 12;
    }
  }
  return 12;  // the top-k screen lives in encode_gemm2.cu
}

}  // namespace sb
