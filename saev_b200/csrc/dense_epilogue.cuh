// Epilogues of the dense tcgen05 contractions, shared by the single-CTA kernel (encode_gemm.cu) and the CTA-pair kernel
// (dense_gemm2.cu).  One call handles 16 consecutive accumulator columns of one row (thread = TMEM lane = row).
#pragma once

#include "common.cuh"

namespace sb {

constexpr int EPI_CHUNK = 16;  // accumulator columns per tcgen05.ld

struct EpiExtra {
  __nv_bfloat16* f_hi;   // EPI 2: out, EPI 3: in   [M, ldf]
  __nv_bfloat16* f_lo;   // EPI 2: out
  __nv_bfloat16* t_hi;   // EPI 2/3: transposed out  [N, ldt]
  __nv_bfloat16* t_lo;
  __nv_bfloat16* f_lo2;  // optional third pieces (6-term split: value = hi + lo + lo2 to ~2^-24)
  __nv_bfloat16* t_lo2;
  long long ldf, ldt;
  float* row_l1;         // EPI 2: += sum_cols f
  float* row_l0;         // EPI 2: += count_cols (f > 0)
  int* active;           // EPI 2: [N] = 1 where some row fired
  float l1_over_b;       // EPI 3
  int n_main;            // EPI 4
  float* extra;          // EPI 4: [M]
  const int* m_limit_dev;  // optional device-side row count (<= M) and contraction length (<= K): the AuxK path
  const int* k_limit_dev;  // works on the dead latents, whose number the host never reads
  const int* row_map;      // EPI 4: output row index of accumulator row r (scatter into the full gradient)
  float alpha;             // EPI 1 / 4: scale of the accumulator
  int ksplit;              // EPI 1 / 4: CTAs sharing the K chunks of one output tile (>= 1)
  // sub-range of the problem (Matryoshka prefix blocks of the dense path): rows [m_begin, M), columns [n_begin, N) and
  // contraction elements [k_begin, K) of the operands the tensor maps describe; TMA coordinates and every epilogue
  // address stay ABSOLUTE, so a block of dictionary columns is just a window on the full operands
  int m_begin, n_begin, k_begin;
};

// fire-and-forget fp32 vector add to global memory (performed by the L2 with round-to-nearest): how the epilogue
// folds the K chunks of one output tile together without a read-modify-write round trip
__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// EPI: 1 = dense fp32 store of alpha * (acc + bias)   2 = ReLU forward   3 = ReLU backward   4 = weight gradient
//      5 = dictionary coherence screen (see encode_gemm.cu for the full description of each)
// a: the 16 accumulator values (raw bits); bs: the 16 bias values of these columns (shared memory); row / col0: ABSOLUTE
// output coordinates; accum: add into the output (later K chunks of a tile, or a K split) instead of storing.
template <int EPI>
__device__ __forceinline__ void dense_epi_chunk(const EpiExtra& ex, const uint32_t (&a)[EPI_CHUNK], const float* bs, int row,
                                                int col0, int M, int n_cols, bool accum, float* __restrict__ out,
                                                long long ldo, int lane, float& acc_l1, float& acc_l0, float& best,
                                                int& best_col) {
  constexpr int CHUNK = EPI_CHUNK;
  float v[CHUNK];
#pragma unroll
  for (int i = 0; i < CHUNK; i += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bs + i);  // broadcast read
    v[i] = __uint_as_float(a[i]) + b4.x;
    v[i + 1] = __uint_as_float(a[i + 1]) + b4.y;
    v[i + 2] = __uint_as_float(a[i + 2]) + b4.z;
    v[i + 3] = __uint_as_float(a[i + 3]) + b4.w;
  }
  if (EPI == 1 || EPI == 4) {
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) v[i] *= ex.alpha;
  }
  if (EPI == 1) {
    if (row < M) {
      float* o = out + static_cast<long long>(row) * ldo + col0;
      if (col0 + CHUNK <= n_cols && ((ldo | col0) & 3) == 0) {
#pragma unroll
        for (int i = 0; i < CHUNK; i += 4) {
          if (accum) red_add_f32x4(o + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
          else *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CHUNK; ++i)
          if (col0 + i < n_cols) {
            if (accum) atomicAdd(o + i, v[i]);
            else o[i] = v[i];
          }
      }
    }
  } else if (EPI == 2) {
    const bool rv = row < M;
    __nv_bfloat16 hi[CHUNK], lo[CHUNK], lo2[CHUNK];
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) {
      const float f = (rv && col0 + i < n_cols) ? fmaxf(v[i], 0.f) : 0.f;
      hi[i] = __float2bfloat16_rn(f);
      const float r1 = f - __bfloat162float(hi[i]);
      lo[i] = __float2bfloat16_rn(r1);
      lo2[i] = __float2bfloat16_rn(r1 - __bfloat162float(lo[i]));
      acc_l1 += f;
      acc_l0 += (f > 0.f) ? 1.f : 0.f;
      const bool fired = __any_sync(FULL, f > 0.f);
      if (fired && lane == 0 && ex.active != nullptr) ex.active[col0 + i] = 1;
    }
    if (rv) {
      __nv_bfloat16* fh = ex.f_hi + static_cast<long long>(row) * ex.ldf + col0;
      __nv_bfloat16* fl = ex.f_lo + static_cast<long long>(row) * ex.ldf + col0;
      if (col0 + CHUNK <= n_cols && ((ex.ldf | col0) & 7) == 0) {
        *reinterpret_cast<uint4*>(fh) = *reinterpret_cast<const uint4*>(&hi[0]);
        *reinterpret_cast<uint4*>(fh + 8) = *reinterpret_cast<const uint4*>(&hi[8]);
        *reinterpret_cast<uint4*>(fl) = *reinterpret_cast<const uint4*>(&lo[0]);
        *reinterpret_cast<uint4*>(fl + 8) = *reinterpret_cast<const uint4*>(&lo[8]);
        if (ex.f_lo2 != nullptr) {
          __nv_bfloat16* f2 = ex.f_lo2 + static_cast<long long>(row) * ex.ldf + col0;
          *reinterpret_cast<uint4*>(f2) = *reinterpret_cast<const uint4*>(&lo2[0]);
          *reinterpret_cast<uint4*>(f2 + 8) = *reinterpret_cast<const uint4*>(&lo2[8]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CHUNK; ++i)
          if (col0 + i < n_cols) {
            fh[i] = hi[i];
            fl[i] = lo[i];
            if (ex.f_lo2 != nullptr) ex.f_lo2[static_cast<long long>(row) * ex.ldf + col0 + i] = lo2[i];
          }
      }
      if (ex.t_hi != nullptr) {
#pragma unroll
        for (int i = 0; i < CHUNK; ++i)
          if (col0 + i < n_cols) {
            ex.t_hi[static_cast<long long>(col0 + i) * ex.ldt + row] = hi[i];
            ex.t_lo[static_cast<long long>(col0 + i) * ex.ldt + row] = lo[i];
            if (ex.t_lo2 != nullptr) ex.t_lo2[static_cast<long long>(col0 + i) * ex.ldt + row] = lo2[i];
          }
      }
    }
  } else if (EPI == 3) {
    if (row < M) {
      __align__(16) __nv_bfloat16 fh[CHUNK];
      const __nv_bfloat16* fp = ex.f_hi + static_cast<long long>(row) * ex.ldf + col0;
      if (col0 + CHUNK <= n_cols && ((ex.ldf | col0) & 7) == 0) {
        *reinterpret_cast<uint4*>(&fh[0]) = __ldg(reinterpret_cast<const uint4*>(fp));
        *reinterpret_cast<uint4*>(&fh[8]) = __ldg(reinterpret_cast<const uint4*>(fp + 8));
      } else {
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) fh[i] = (col0 + i < n_cols) ? fp[i] : __float2bfloat16_rn(0.f);
      }
#pragma unroll
      for (int i = 0; i < CHUNK; ++i) {
        if (col0 + i < n_cols) {
          const float d = (__bfloat162float(fh[i]) > 0.f) ? v[i] + ex.l1_over_b : 0.f;
          const __nv_bfloat16 h = __float2bfloat16_rn(d);
          const float r1 = d - __bfloat162float(h);
          const __nv_bfloat16 l = __float2bfloat16_rn(r1);
          ex.t_hi[static_cast<long long>(col0 + i) * ex.ldt + row] = h;
          ex.t_lo[static_cast<long long>(col0 + i) * ex.ldt + row] = l;
          if (ex.t_lo2 != nullptr)
            ex.t_lo2[static_cast<long long>(col0 + i) * ex.ldt + row] = __float2bfloat16_rn(r1 - __bfloat162float(l));
        }
      }
    }
  } else if (EPI == 5) {
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) {
      const float a_abs = fabsf(v[i]);
      if (col0 + i > row && col0 + i < n_cols && a_abs > best) {
        best = a_abs;
        best_col = col0 + i;
      }
    }
  } else {  // EPI == 4
    if (row < M) {
      const long long orow = ex.row_map != nullptr ? ex.row_map[row] : row;
      float* o = out + orow * ldo + col0;
      if (col0 + CHUNK <= ex.n_main && ((ldo | col0) & 3) == 0) {
#pragma unroll
        for (int i = 0; i < CHUNK; i += 4) {
          if (accum) red_add_f32x4(o + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
          else *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) {
          if (col0 + i < ex.n_main) {
            if (accum) atomicAdd(o + i, v[i]);
            else o[i] = v[i];
          } else if (col0 + i == ex.n_main && ex.extra != nullptr) {
            if (accum) atomicAdd(ex.extra + orow, v[i]);
            else ex.extra[orow] = v[i];
          }
        }
      }
    }
  }
}

}  // namespace sb
