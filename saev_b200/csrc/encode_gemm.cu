// Encoder contraction on tcgen05 tensor cores:  h~ = x_bf16 . W_enc_bf16^T (+ b_enc), fp32 accumulate in TMEM.
//
// Replaces the dense `einsum(x, W_enc) + b_enc` of saev (src/saev/nn/modeling.py:343-347) and, in the
// TopK epilogue, the `topk -> scatter -> mul` of TopKActivation.forward (modeling.py:169-179): the [B,S]
// pre-activation matrix is never written to HBM.  Each CTA owns one 128-row block of the batch and sweeps a
// contiguous range of 256-column tiles of the dictionary; the epilogue warps read the accumulator out of
// TMEM (one thread = one batch row) and keep a running list of the KP largest pre-activations of their row
// in shared memory (threshold filter + warp-cooperative compaction).  The lists are *candidates*: the bf16
// products carry ~2^-9 relative error, so `rescore_topk_kernel` (sparse_kernels.cu) recomputes the exact fp32
// pre-activation of every candidate from the fp32 master weights and picks the final top-k from those.
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator,
// warps 4-7 = epilogue (TMEM lane quadrant = warp_idx % 4).  Pipelines: STAGES-deep smem ring (TMA <-> MMA),
// 2-deep TMEM accumulator ring (MMA <-> epilogue), so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operands are K-major bf16 with 128-byte swizzle.  `nterms == 3` runs the error-compensated split product
// (x_hi.W_hi + x_hi.W_lo + x_lo.W_hi, ~2^-17 relative) by walking three (A,B) tensor-map pairs along K;
// that mode + the dense-store epilogue are used for the dense (ReLU / AuxK) paths and for testing the
// contraction itself.
#include "common.cuh"
#include "kernels.h"

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
  int stages, cap, list_stride;
  size_t off_lists_val, off_lists_idx, off_bias, off_bars, total;
};

__host__ __device__ inline EncodeSmemLayout encode_smem_layout(int stages, int cap) {
  EncodeSmemLayout L;
  L.stages = stages;
  L.cap = cap;
  L.list_stride = cap + 1;  // odd stride: lanes (=rows) appending at equal counts hit distinct banks
  size_t o = static_cast<size_t>(stages) * STAGE_BYTES;
  L.off_lists_val = o;
  o += static_cast<size_t>(BM) * L.list_stride * 4;
  L.off_lists_idx = o;
  o += static_cast<size_t>(BM) * L.list_stride * 4;
  L.off_bias = o;
  o += 2 * BN * 4;
  L.off_bars = (o + 7) & ~size_t(7);
  o = L.off_bars + (2 * stages + 4) * 8 + 16;
  L.total = o + 1024;  // slack for the manual 1024-byte alignment of the dynamic smem base
  return L;
}

// Warp-cooperative compaction of one row's candidate list: keep the `KP` largest of `n` entries
// (n <= CAP <= 64+32 handled with up to 3 entries per lane), sorted descending into slots [0, KP).
// Returns the KP-th largest value (the new admission threshold) or -inf when n < KP.
template <int KP, int CAP>
__device__ __forceinline__ float compact_row(float* vals, int* idxs, int n, int lane) {
  constexpr int PER = (CAP + 31) / 32;
  float v[PER];
  int id[PER];
  int rank[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int s = lane + 32 * e;
    v[e] = (s < n) ? vals[s] : -INFINITY;
    id[e] = (s < n) ? idxs[s] : -1;
    rank[e] = 0;
  }
  for (int s = 0; s < n; ++s) {
    const float vs = vals[s];  // broadcast read
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int mine = lane + 32 * e;
      rank[e] += (vs > v[e]) || (vs == v[e] && s < mine);
    }
  }
  __syncwarp();
  float kth = -INFINITY;
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    const int s = lane + 32 * e;
    const bool keep = (s < n) && (rank[e] < KP);
    if (keep) {
      vals[rank[e]] = v[e];
      idxs[rank[e]] = id[e];
    }
    const unsigned hit = __ballot_sync(FULL, (s < n) && (rank[e] == KP - 1));
    if (hit) kth = __shfl_sync(FULL, v[e], __ffs(hit) - 1);
  }
  __syncwarp();
  return kth;
}

// EPI: 0 = running top-KP candidate lists, 1 = dense fp32 store of (acc + bias).
template <int EPI, int KP, int CAP, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
encode_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   int nterms, int kblocks_per_term, const float* __restrict__ bias, int M, int N, int m_blocks,
                   int tiles_per_split, int nsplit, const int* __restrict__ n_limit_dev,
                   float* __restrict__ cand_val, int* __restrict__ cand_idx, float* __restrict__ out, long long ldo) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_u32 + pad;

  const EncodeSmemLayout L = encode_smem_layout(STAGES, CAP);
  float* list_val = reinterpret_cast<float*>(smem + L.off_lists_val);
  int* list_idx = reinterpret_cast<int*>(smem + L.off_lists_idx);
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
  const int split = blockIdx.x / m_blocks;
  // The column count may live on the device (AuxK dead-latent list whose length the host never reads).
  const int n_cols = n_limit_dev ? min(N, *n_limit_dev) : N;
  const int n_tiles_total = (n_cols + BN - 1) / BN;
  const int tile_begin = split * tiles_per_split;
  const int tile_end = min(n_tiles_total, tile_begin + tiles_per_split);
  const int num_tiles = max(0, tile_end - tile_begin);
  const int kblocks = nterms * kblocks_per_term;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
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
      for (int t = 0; t < num_tiles; ++t) {
        const int n0 = (tile_begin + t) * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          const int term = kb / kblocks_per_term;
          const int k0 = (kb - term * kblocks_per_term) * BK;
          const CUtensorMap* ma = (term == 2) ? &tmA_lo : &tmA_hi;
          const CUtensorMap* mb = (term == 1) ? &tmB_lo : &tmB_hi;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          tma_load_2d(sa, ma, full_bar(stage), k0, m_blk * BM);
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
      for (int t = 0; t < num_tiles; ++t) {
        const int as = t & 1;
        const uint32_t aphase = (t >> 1) & 1u;
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
    const int row = m_blk * BM + row_local;
    float* my_val = list_val + row_local * L.list_stride;
    int* my_idx = list_idx + row_local * L.list_stride;
    float tau = -INFINITY;
    int cnt = 0;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127

    for (int t = 0; t < num_tiles; ++t) {
      const int as = t & 1;
      const uint32_t aphase = (t >> 1) & 1u;
      const int n0 = (tile_begin + t) * BN;
      float* bs = bias_s + as * BN;
      // stage the bias slice of this tile (double buffered by accumulator stage; see barrier note below)
      for (int c = et; c < BN; c += 128) bs[c] = (bias != nullptr && n0 + c < n_cols) ? bias[n0 + c] : 0.f;
      named_bar_sync(1, 128);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;

      // Two register buffers, unrolled by hand so the accumulator stays in registers: while chunk c is
      // being filtered the tcgen05.ld of chunk c+1 is in flight.
      auto process = [&](uint32_t (&a)[CHUNK], int c) {
        const int col0 = n0 + c * CHUNK;
        if (EPI == 0) {
          // make room: every lane must be able to take CHUNK appends
          unsigned need = __ballot_sync(FULL, cnt > CAP - CHUNK);
          while (need) {
            const int l = __ffs(need) - 1;
            need &= need - 1;
            const int n = __shfl_sync(FULL, cnt, l);
            const float kth = compact_row<KP, CAP>(list_val + (q * 32 + l) * L.list_stride,
                                                   list_idx + (q * 32 + l) * L.list_stride, n, lane);
            if (lane == l) {
              cnt = KP;
              tau = kth;
            }
          }
#pragma unroll
          for (int i = 0; i < CHUNK; ++i) {
            const float v = __uint_as_float(a[i]) + bs[c * CHUNK + i];
            if (v > tau && col0 + i < n_cols) {
              my_val[cnt] = v;
              my_idx[cnt] = col0 + i;
              ++cnt;
            }
          }
          __syncwarp();
        } else {
          if (row < M) {
            float* o = out + static_cast<long long>(row) * ldo + col0;
            if (col0 + CHUNK <= n_cols && (ldo & 3) == 0) {
#pragma unroll
              for (int i = 0; i < CHUNK; i += 4) {
                float4 w;
                w.x = __uint_as_float(a[i]) + bs[c * CHUNK + i];
                w.y = __uint_as_float(a[i + 1]) + bs[c * CHUNK + i + 1];
                w.z = __uint_as_float(a[i + 2]) + bs[c * CHUNK + i + 2];
                w.w = __uint_as_float(a[i + 3]) + bs[c * CHUNK + i + 3];
                *reinterpret_cast<float4*>(o + i) = w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < CHUNK; ++i)
                if (col0 + i < n_cols) o[i] = __uint_as_float(a[i]) + bs[c * CHUNK + i];
            }
          }
        }
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

    if (EPI == 0) {
      // final compaction of all 32 rows of this warp, then a coalesced write of the KP-entry lists
      for (int l = 0; l < 32; ++l) {
        const int n = __shfl_sync(FULL, cnt, l);
        float* rv = list_val + (q * 32 + l) * L.list_stride;
        int* ri = list_idx + (q * 32 + l) * L.list_stride;
        compact_row<KP, CAP>(rv, ri, n, lane);
        const int keep = min(n, KP);
        const int grow = m_blk * BM + q * 32 + l;
        if (grow < M) {
          const long long base = (static_cast<long long>(grow) * nsplit + split) * KP;
          for (int s = lane; s < KP; s += 32) {
            cand_val[base + s] = (s < keep) ? rv[s] : -INFINITY;
            cand_idx[base + s] = (s < keep) ? ri[s] : -1;
          }
        }
      }
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
  return r == CUDA_SUCCESS ? 0 : 2;
}

template <int EPI, int KP, int CAP, int STAGES>
static int launch_variant(const EncodeGemmArgs& a, const CUtensorMap* maps, int m_blocks, int tiles_per_split,
                          int nsplit, cudaStream_t stream) {
  auto kern = encode_gemm_kernel<EPI, KP, CAP, STAGES>;
  const EncodeSmemLayout L = encode_smem_layout(STAGES, CAP);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total)) !=
        cudaSuccess)
      return 3;
    attr_set = true;
  }
  const int kblocks_per_term = (a.K + BK - 1) / BK;
  kern<<<m_blocks * nsplit, NUM_THREADS, L.total, stream>>>(maps[0], maps[1], maps[2], maps[3], a.nterms,
                                                           kblocks_per_term, a.bias, a.M, a.N, m_blocks,
                                                           tiles_per_split, nsplit, a.n_limit_dev, a.cand_val,
                                                           a.cand_idx, a.out, a.ldo);
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

int encode_gemm_kp(int top_k) {
  if (top_k <= 32) return 40;
  if (top_k <= 64) return 72;
  return -1;
}

int launch_encode_gemm(const EncodeGemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return 0;
  if ((a.K % 8) != 0) return 10;  // TMA needs 16-byte aligned row pitch
  CUtensorMap maps[4];
  const long long ldk = a.K;
  if (make_tmap_bf16(&maps[0], a.A_hi, a.M, a.K, ldk, BM)) return 11;
  if (make_tmap_bf16(&maps[2], a.B_hi, a.N, a.K, ldk, BN)) return 11;
  if (a.nterms == 3) {
    if (make_tmap_bf16(&maps[1], a.A_lo, a.M, a.K, ldk, BM)) return 11;
    if (make_tmap_bf16(&maps[3], a.B_lo, a.N, a.K, ldk, BN)) return 11;
  } else {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  const int m_blocks = (a.M + BM - 1) / BM;
  const int n_tiles = (a.N + BN - 1) / BN;
  if (a.epilogue == 1) {
    // dense store: any split works; use enough CTAs to fill the GPU
    int nsplit = a.nsplit > 0 ? a.nsplit : encode_gemm_nsplit(a.M, a.N, a.num_sms);
    const int tps = (n_tiles + nsplit - 1) / nsplit;
    nsplit = (n_tiles + tps - 1) / tps;
    return launch_variant<1, 8, 8, 4>(a, maps, m_blocks, tps, nsplit, stream);
  }
  const int nsplit = a.nsplit;
  const int tps = (n_tiles + nsplit - 1) / nsplit;
  if (a.kp == 40) return launch_variant<0, 40, 64, 3>(a, maps, m_blocks, tps, nsplit, stream);
  if (a.kp == 72) return launch_variant<0, 72, 96, 2>(a, maps, m_blocks, tps, nsplit, stream);
  return 12;
}

}  // namespace sb
