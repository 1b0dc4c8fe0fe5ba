// Internal launcher declarations shared by the .cu files of libsaev_b200.so (not part of the C ABI).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace sb {

constexpr int ENCODE_MAX_NSPLIT = 8;
constexpr int ENCODE_CAPG = 256;   // (template parameter of the dense tcgen05 kernel; no candidate lists there any more)
constexpr int ENCODE2_CAPG = 384;  // entries per candidate list of the CTA-pair screen (encode_gemm2.cu)

// ---- slots of the 128-byte scalar block in the workspace (Workspace::scalars), 4 bytes each ----------------------
enum ScalarSlot {
  SC_N_DEAD = 0,        // int    dead latents after the tracker update of the last training forward
  SC_UNSAFE_TOTAL = 1,  // uint   rows the screen could not certify (cumulative since sync_weights); all were repaired
  SC_AUX_LOSS = 2,      // float
  SC_RESCORED = 3,      // uint   candidates re-scored in fp32 (cumulative)
  // three floats a sharded optimizer MAX-all-reduces as one vector (saev_b200_wnorm_scalar):
  SC_WNORM_SQ_MAX = 4,  // float  max_j ||W_enc_t[j]||^2
  SC_BIAS_ABS_MAX = 5,  // float  max_j |b_enc[j]|
  SC_RHO = 6,           // float  max_j ||w_j - fp16(w_j)|| / max(||w_j||, ||fp16(w_j)||)
  SC_N_UNSAFE = 7,      // int    rows of the CURRENT forward handed to the exact repair path
  SC_REPAIRED = 8,      // uint   rows re-done by the exact path (cumulative)
  SC_UNSAFE_ERR = 9,    // uint   of SC_UNSAFE_TOTAL: rows whose OBSERVED screen error exceeded the bound (must stay 0:
                        //        it would mean the error model is wrong; list overflows are the expected cause)
  SC_MERGED = 10,       // uint   candidate-list entries merged (cumulative)
  SC_RHO_GUESS = 11,    // float  threshold guess of the current forward, as (L + max|b|) / (||x|| max||w||); -inf: none
  SC_GUESS_FAILED = 12, // uint   of SC_UNSAFE_TOTAL: rows whose threshold guess turned out too high (cumulative)
  SC_SLOTS = 32
};

// ---- deterministic error bound of the fp16 top-k screen -----------------------------------------------------------
// The screen computes  h~ = 2^e * sum_d x16_d w16_d + b  on the tensor cores (x16 = fp16(x 2^-e), w16 = fp16(w), fp32
// accumulation), the re-score kernel  he = fl32(sum_d x_d w_d) + b.  In exact arithmetic
//     h~ - h = 2^e <x16 - x 2^-e, w16> + <x, w16 - w>,
// so by Cauchy-Schwarz on the ACTUAL rounding residuals (their norms are computed where the fp16 copies are made:
// prep_x_kernel per batch row, the Adam / sync kernels per dictionary row)
//     |h~ - he| <= ||dx_b|| ||w16_j|| + ||x_b|| ||dw_j|| + (g + r) ||x_b|| ||w_j|| + r |b_j|
// where g bounds both accumulation errors (tensor core: one truncating fp32 update per 16 products, doubled for
// safety; re-score: D/32 sequential FMAs per lane + 5 shuffle levels) and r = 2^-21 the final roundings (bias add, the
// epilogue's own FMAs).  Residual norms instead of the unit roundoff (2^-11 per operand) make the bound 2.3x tighter
// for ordinary dense rows -- fp16 rounds to nearest, the residual of a dense vector is ~0.43 * 2^-11 of its norm --
// and cover subnormals / exact zeros without special terms.  With c_j = max(||w_j||, ||w16_j||) (stored per column,
// rounded up) and rho = max_j ||dw_j|| / c_j (one scalar):
//     E_bj = c_j P_b + Q_b,      P_b = ||dx_b|| + ||x_b|| (rho + g + r),      Q_b = r max_j |b_j|.
// The per-column factor matters once a few atoms carry much larger encoder rows than the rest: a bound through
// max_j ||w_j|| would widen the admission band of every column.
struct ScreenBound {
  float c, q;  // P_b = 1.01 (dxn_b + xn_b c),  Q_b = 1.01 q   (the bound is evaluated in fp32 itself: 1 % head room)
};
__host__ __device__ inline ScreenBound screen_bound(int D, float rho, float bias_abs_max) {
  const float g = (2.f * (D / 16 + 16) + (D / 32 + 8)) * 1.1920929e-7f;  // * 2^-23
  const float r = 4.7683716e-7f;                                           // 2^-21
  ScreenBound b;
  b.c = rho + g + r;
  b.q = r * bias_abs_max;
  return b;
}
__host__ __device__ inline float screen_P(const ScreenBound& s, float xn, float dxn) { return 1.01f * (dxn + xn * s.c); }
__host__ __device__ inline float screen_Q(const ScreenBound& s) { return 1.01f * s.q; }
// ---- threshold guess (warm start of the screen) --------------------------------------------------------------------
// A cold sweep admits ~k (1 + ln(S / k)) columns per row before its threshold has converged; with the final threshold
// known up front only the ~70 columns of the final band are admitted and the kernel runs at the tensor roofline
// (measured at c3: 1.33 ms instead of 2.3 ms).  So every row starts from a GUESS of its k-th largest lower bound,
//     L_guess_b = rho * ||x_b|| max_j||w_j|| - max_j|b_j|,
// rho = a low quantile (GUESS_QUANTILE, shrunk by GUESS_SAFETY) of the same ratio over the rows of the PREVIOUS
// forward (an 8192-bin device histogram filled by the re-score kernel).  The guess is verified, not trusted: the
// re-score accepts a row only if it holds k candidates whose lower bounds reach L_guess_b (then L_guess_b really was a
// lower bound of the exact k-th largest value and nothing was missed); any other row goes to the exact path.
constexpr int GUESS_BINS = 8192;          // ratio in [-1, 1), bin width 2 / GUESS_BINS
constexpr float GUESS_QUANTILE = 5e-4f;
constexpr float GUESS_SAFETY = 0.015f;

constexpr float FP16_MAX = 65504.f;  // encoder rows with a larger norm cannot be screened in fp16 (all rows repaired)

// number of kernels this library has launched (all handles); read through saev_b200_launch_count()
extern unsigned long long g_launch_count;

// Row-per-warp kernels keep a whole D-vector in registers: lane l owns float4 vectors l, l + 32, ..., VPL of them.
#define SB_DISPATCH_VPL(D, CALL)                                         \
  do {                                                                   \
    const int need_ = ((D) + 127) / 128;                                 \
    ++g_launch_count;                                                    \
    if (need_ <= 1) { constexpr int VPL = 1; CALL; }                     \
    else if (need_ <= 2) { constexpr int VPL = 2; CALL; }                \
    else if (need_ <= 4) { constexpr int VPL = 4; CALL; }                \
    else if (need_ <= 6) { constexpr int VPL = 6; CALL; }                \
    else if (need_ <= 8) { constexpr int VPL = 8; CALL; }                \
    else if (need_ <= 12) { constexpr int VPL = 12; CALL; }              \
    else if (need_ <= 16) { constexpr int VPL = 16; CALL; }              \
    else return 20;                                                      \
  } while (0)

// ---- encode_gemm.cu -------------------------------------------------------------------------------------
struct EncodeGemmArgs {
  const __nv_bfloat16* A_hi = nullptr;  // [M, K] row-major (K contiguous)
  const __nv_bfloat16* A_lo = nullptr;  // residual part, nterms == 3 only
  const __nv_bfloat16* B_hi = nullptr;  // [N, K] row-major (K contiguous)
  const __nv_bfloat16* B_lo = nullptr;
  const __nv_bfloat16* A_lo2 = nullptr; // third pieces, nterms == 6 only
  const __nv_bfloat16* B_lo2 = nullptr;
  int nterms = 1;                       // 1: A_hi.B_hi    3: + A_hi.B_lo + A_lo.B_hi    6: + A_hi.B_lo2 + A_lo2.B_hi + A_lo.B_lo
  int m_begin = 0, n_begin = 0, k_begin = 0;  // dense epilogues: work on rows [m_begin, M), columns [n_begin, N),
                                        // contraction [k_begin, K) of the operands (absolute indices everywhere)
  int M = 0, N = 0, K = 0;
  long long lda = 0, ldb = 0;           // row pitch of A / B in elements (0 = K); must be multiples of 8
  int k_chunk_blocks = 0;               // epilogues 1 / 4: accumulate K in chunks of this many 64-wide k-blocks per term
                                        // (partial results added in fp32 by the epilogue); 0 = one chunk
  const float* bias = nullptr;          // [N] or null
  const int* n_limit_dev = nullptr;     // optional device-side column count (<= N)
  const int* m_limit_dev = nullptr;     // optional device-side row count (<= M)          (dense epilogues)
  const int* k_limit_dev = nullptr;     // optional device-side contraction length (<= K; operands zero beyond it)
  const int* row_map = nullptr;         // epilogue 4: output row of accumulator row r
  float alpha = 1.f;                    // epilogues 1 / 4: out = alpha * (acc + bias)
  int ksplit = 1;                       // epilogues 1 / 4 with k_chunk_blocks > 0: spread the K chunks of one output tile
                                        // over this many CTAs, all adding into a PRE-ZEROED output (bias must be null)
  int epilogue = 1;                     // 1: dense fp32 store, 2: ReLU forward, 3: ReLU backward, 4: weight gradient,
                                        // 5: coherence screen (see encode_gemm.cu); the top-k screen is encode_gemm2.cu
  __nv_bfloat16* f_hi = nullptr;        // epilogue 2 (out) / 3 (in): relu(h) as bf16 hi [M, ldf]
  __nv_bfloat16* f_lo = nullptr;        // epilogue 2: bf16 residual
  __nv_bfloat16* t_hi = nullptr;        // epilogue 2/3: transposed bf16 hi/lo outputs [N, ldt]
  __nv_bfloat16* t_lo = nullptr;
  __nv_bfloat16* f_lo2 = nullptr;       // optional third pieces (written when non-null)
  __nv_bfloat16* t_lo2 = nullptr;
  long long ldf = 0, ldt = 0;
  float* row_l1 = nullptr;              // epilogue 2: [M] += sum f   (caller zeroes)
  float* row_l0 = nullptr;              // epilogue 2: [M] += count f > 0
  int* active = nullptr;                // epilogue 2: [N] = 1 where a row fired
  float l1_over_b = 0.f;                // epilogue 3
  int n_main = 0;                       // epilogue 4: columns < n_main go to out, column n_main to extra[row]
  float* extra = nullptr;
  int top_k = 32;                       // screen (pair kernel): k of the final selection
  const float* row_norm = nullptr;      // screen: [M] ||x_b||_2
  const float* row_dx = nullptr;        // screen: [M] ||2^e_b x16_b - x_b||_2, the row's fp16 rounding residual
  const float* row_scale = nullptr;     // screen: [M] 2^e_b, the power of two the fp16 operand row was divided by
  const float* scalars = nullptr;       // screen: the workspace scalar block (SC_WNORM_SQ_MAX, SC_BIAS_ABS_MAX)
  const float* col_norm = nullptr;      // screen: [N] c_j = max(||w_j||, ||fp16(w_j)||) (rounded up)
  int nsplit = 1;                       // column splits of the dense epilogues
  int num_sms = 148;
  int* cand_cnt = nullptr;              // screen: [rows, nlists] entries kept (negative: overflowed)
  void* cand = nullptr;                 // screen: [rows padded to 256, nlists, ENCODE2_CAPG] x {value bits, column}
  unsigned int* tau_keys = nullptr;     // screen: [rows padded to 256] shared admission thresholds (scratch)
  int tau_preset = 0;                   // screen: tau_keys already holds the rows' threshold guesses (do not clear)
  float* out = nullptr;                 // epilogue 1: [M, ldo]
  long long ldo = 0;
};
int launch_encode_gemm(const EncodeGemmArgs& a, cudaStream_t stream);
// CTA-pair version of the dense split-product contraction (dense_gemm2.cu): epilogues 1-4, 3 or 6 terms, static sizes;
// launch_encode_gemm forwards to it when dense_gemm2_eligible(a) (SAEV_B200_DENSE_PAIR=0 keeps the single-CTA kernel).
bool dense_gemm2_eligible(const EncodeGemmArgs& a);
int launch_dense_gemm2(const EncodeGemmArgs& a, cudaStream_t stream);

// ---- encode_gemm2.cu: CTA-pair (cta_group::2) top-k screen ------------------------------------------------
struct Encode2Plan {
  int m_pairs = 0;   // 256-row blocks of the batch
  int n_tiles = 0;   // 256-column tiles of the dictionary
  int q = 1;         // tile-steps per CTA pair (contiguous range of the linearised (row block, tile) space)
  int n_pairs = 1;   // CTA pairs launched
  int nsplit = 1;    // most ranges touching one row block
  int nlists = 2;    // candidate lists per row = 2 * nsplit (two column halves per range)
};
int encode2_max_pairs();                                  // co-resident CTA pairs on this device (0: unavailable)
Encode2Plan encode2_plan(int M, int N, int max_pairs);
// uses A_hi / B_hi (FP16 operands here: x 2^-e and W_enc_t), M, N, K, bias, top_k, row_norm, row_scale, scalars, cand
// ([rows padded to 256][nlists][ENCODE2_CAPG]), cand_cnt ([rows padded to 256][nlists]), tau_keys of `a`
int launch_encode_gemm2(const EncodeGemmArgs& a, const Encode2Plan& pl, cudaStream_t stream);
int encode_gemm_nsplit(int M, int N, int num_sms);
int encode2_max_top_k();

// ---- sparse_kernels.cu ----------------------------------------------------------------------------------
// `gate` (optional, device): the kernel does nothing when *gate == 0 (AuxK operand prep with no dead latents)
int launch_split_bf16(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n, cudaStream_t s,
                      __nv_bfloat16* lo2 = nullptr, const int* gate = nullptr);
// x[B,D] -> fp16 operand x16[b] = fp16(x[b] * 2^-e_b) (2^e_b: power of two just above ||x_b||_inf, exact scaling),
// row_norm[b] = ||x_b||_2, row_dx[b] = ||2^e_b x16_b - x_b||_2 (both rounded up), row_scale[b] = 2^e_b
// (operand and error-bound inputs of the top-k screen, encode_gemm2.cu)
// Also seeds the screen's per-row thresholds from the guess in scalars[SC_RHO_GUESS]: tau_keys[b] (rows up to
// `rows_padded`, 0 = none) in the kernel's threshold space, guess_L[b] = L_guess_b (-inf = none).
int launch_prep_x(const float* x, int B, int D, __half* x16, float* row_norm, float* row_dx, float* row_scale,
                  const float* scalars, unsigned int* tau_keys, float* guess_L, int rows_padded, cudaStream_t s);
// scalars[SC_RHO_GUESS] from the histogram the previous forward filled (then clears it for re-use)
int launch_screen_guess(int* hist, float quantile, float safety, int min_rows, float* scalars, cudaStream_t s);
// fp32 -> fp16 (round to nearest): the screen's copy of W_enc_t
int launch_to_half(const float* src, __half* dst, long long n, cudaStream_t s);
// *out_max = max_j ||W[j,:]||^2; optionally the per-row inputs of the screen's error bound: col_norm[j] =
// max(||W[j]||, ||fp16(W[j])||) (rounded up) and *rho = max_j ||W[j] - fp16(W[j])|| / col_norm[j]
int launch_row_sumsq_max(const float* W, int rows, int cols, float* out_max, cudaStream_t s, float* col_norm = nullptr,
                         float* rho = nullptr);
// *out = max_i |v[i]|
int launch_abs_max(const float* v, int n, float* out, cudaStream_t s);
int launch_normalize_rows(float* W, int rows, int cols, cudaStream_t s);
// datapoint initialisation (saev train.py:141-185); see the kernel
int launch_datapoint_init(const float* acts, const long long* src_row, const float* mean, const float* noise,
                          const long long* noise_row, float blend, int tie, int normalize, int S, int D, float* W_enc_t,
                          float* W_dec, cudaStream_t s);
int launch_log_metrics(const float* x, const float* r, int B, int D, const float* W, int S, const int* fired,
                       double* acc, const float* coh, double* out, cudaStream_t s);
int launch_eval_accumulate(const float* x, const float* r, int B, int D, const float* losses, double* acc,
                           cudaStream_t s);
int launch_feature_stats_topk(const int* idx, const float* val, long long n, float* n_fired, float* values,
                              cudaStream_t s);
int launch_feature_stats_dense(const __nv_bfloat16* hi, const __nv_bfloat16* lo, const __nv_bfloat16* lo2, int B, int S,
                               long long ld, float* n_fired, float* values, cudaStream_t s);
int launch_unit_rows_split(const float* W, int rows, int D, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s);
int launch_coherence_finish(const float* W, int D, const float* row_best, const int* row_col, int n_entries, int nsplit,
                            float slack, float* out, cudaStream_t s);

struct RescoreArgs {
  void* cand; const int* cand_cnt; int cand_stride; int nsplit;  // (row, list) candidate buffers
  const float* row_norm; const float* row_dx; const float* row_scale;
  const float* col_norm;  // [S] c_j = max(||w_j||, ||fp16(w_j)||) (rounded up)
  float* scalars;       // workspace scalar block (ScalarSlot): bounds in, counters out
  const float* x; const float* W_enc_t; const float* b_enc;
  int B, D, S, K;
  int* topk_idx; float* topk_val;
  int* feat_count;      // [S] += 1 per selected (b, j)   (may be null: eval)
  int* active;          // [S] = 1 where a non-zero activation was selected (may be null)
  const float* guess_L; // [B] the threshold guess each row was screened with (-inf: none), verified here
  int* guess_hist;      // [GUESS_BINS] histogram of this forward's threshold / norm ratios (input of the next guess)
  int* unsafe_list;     // [B] rows handed to the exact path (count in scalars[SC_N_UNSAFE], zeroed by the launcher)
  int force_unsafe;     // test switch: treat every row as uncertified
};
// Exact fp32 re-score of the screen's candidates and the final top-k; rows that cannot be certified (list overflow,
// observed screen error above the bound, encoder norm outside the fp16 range) are appended to unsafe_list instead.
int launch_rescore_topk(const RescoreArgs& a, cudaStream_t s);
// Exact path for the rows in unsafe_list: every pre-activation of the row in fp32 (same arithmetic as the re-score),
// block-wide radix selection; writes what launch_rescore_topk would have.  Fixed grids that read the row count from
// the device (no host sync); uses the rows' own (already consumed) candidate lists as scratch.
int launch_repair_topk(const RescoreArgs& a, cudaStream_t s);

// BatchTopK (saev modeling.py:183-244) on the per-row top-`cap` lists the re-score left (batch_topk_kernels.cu): training
// keeps the n_keep largest entries of the whole batch (ties at the cut in flat order) and folds the smallest positive
// survivor into *threshold; eval keeps value > max(*threshold, 0).  Losers become empty slots (idx -1, value 0); the
// per-atom counts (training) and activity flags are rebuilt from the survivors.  stats (device int[4]): kept entries,
// rows whose capacity may have truncated the selection, entries tied at the cut, key of the cut value.
size_t batch_topk_scratch_bytes(int max_batch);
int launch_batch_topk(int* topk_idx, float* topk_val, int B, int cap, int S, long long n_keep, int training,
                      float* threshold, float momentum, int* feat_count, int* active, int* scratch, int* stats,
                      cudaStream_t s);

struct DecodeArgs {
  const float* x; const int* topk_idx; const float* topk_val;
  const float* W_dec; const float* b_dec;
  int B, D, K;
  float grad_scale;     // 2 / (B_global * D)
  float l1_over_b;      // l1_coeff / B_global (0 => NoSparsity)
  float* resid;         // [B, D]  x_hat - x
  float* dh;            // [B, K]  d loss / d h on the active set (null: eval, skip)
  float* row_sse; float* row_l1; float* row_l0;  // [B]
};
int launch_decode(const DecodeArgs& a, cudaStream_t s);

// Matryoshka prefix cuts of one step (saev objectives.py:158-201; modeling.py:364-406): sorted, cut[n - 1] == d_sae.
constexpr int MAX_PREFIXES = 32;
struct PrefixCuts {
  int n;
  int cut[MAX_PREFIXES];
};
// Matryoshka prefixes on the dense path: y[P][B][D] tensor-core partial decodes of the prefix blocks (from each block's
// first 8-aligned column; bit c of tensor_mask: written) + the unaligned head columns and b_dec in fp32 -> resid (last
// prefix), sfx[B][P][D] suffix sums of the per-prefix residuals, row_sse, (training) g[P][B][D] = grad_scale * sfx_c
int launch_dense_prefix_resid(const float* y, const float* x, int B, int D, const PrefixCuts& pf, unsigned int tensor_mask,
                              const __nv_bfloat16* f_hi, const __nv_bfloat16* f_lo, const __nv_bfloat16* f_lo2,
                              long long ldf, const float* W_dec, const float* b_dec, float grad_scale, float* resid,
                              float* sfx, float* row_sse, __nv_bfloat16* g_hi, __nv_bfloat16* g_lo, __nv_bfloat16* g_lo2,
                              cudaStream_t s);
// Decode with prefixes: x_hat_i = b_dec + sum over the active columns below cut[i]; r_i = x_hat_i - x.
// Writes resid = r_{n-1}, sfx[b, c, :] = sum_{i >= c} r_i (what column block c sees in the backward pass),
// row_sse[b] = sum_i ||r_i||^2, and dh (scaled by a.grad_scale = 2 / (B_global * n * D)).
int launch_decode_prefix(const DecodeArgs& a, const PrefixCuts& pf, float* sfx, cudaStream_t s);
// x_hats[b, i, :] = x[b, :] + r_i = x + sfx[b, i] - sfx[b, i + 1]
int launch_x_hats_prefix(const float* sfx, const float* x, int B, int D, int P, float* out, cudaStream_t s);

constexpr int WGRAD_HEAVY_ENTRIES = 512;
int launch_csc_build(const int* topk_idx, int B, int K, int S, const int* feat_count, int* feat_off, int* cursor,
                     int* entries, int* block_totals /* [ceil(S/1024)] scratch */, cudaStream_t s,
                     int* heavy_list = nullptr /* [B K / WGRAD_HEAVY_ENTRIES + 1] */, int* n_heavy = nullptr);

struct WgradArgs {
  const int* feat_off; const int* entries; const float* topk_val;
  const float* dh;                     // null: computed in the kernel from resid / W_dec (single prefix, d_model <= 1024)
  float l1_over_b = 0.f;               // only read when dh == null
  const int* heavy_list = nullptr;     // atoms with more than WGRAD_HEAVY_ENTRIES entries (built with the CSC), handled by
  int* heavy_ticket = nullptr;         // [S] ints (the CSC fill cursor, free by then): slices of a heavy atom finished
  const int* n_heavy = nullptr;        // a block each; null: every atom by one warp
  int warps_per_block = 1;             // atoms (warps) per block of wgrad_kernel, 1..8
  const float* resid; const float* x; const float* W_dec;
  int B, D, S, K;
  float grad_scale; int remove_parallel;
  float* gW_enc_t; float* gb_enc; float* gW_dec;
  float* row_gsq;   // optional [S]: this atom's ||gW_enc_t[j]||^2 + ||gW_dec[j]||^2 + gb_enc[j]^2
  const float* sfx;                    // Matryoshka: [B, pf.n, D] suffix sums of residuals (null: single prefix)
  PrefixCuts pf;
  int row_begin, row_end;              // atoms handled by this launch
  const long long* skip_toks;          // optional: leave atoms with no entries and toks >= threshold untouched
  long long skip_threshold;
  int split = 0;                       // two-pass form: decoder side (+ dh into dh_scratch) over all atoms, then encoder side
  float* dh_scratch = nullptr;         // [B, K]
};
int launch_wgrad(const WgradArgs& a, cudaStream_t s);

// gb_dec[d] (+)= scale * sum_b src[b, d]
int launch_colsum(const float* src, int B, int D, float scale, int accumulate, float* partial, float* out,
                  cudaStream_t s, long long row_stride = 0 /* elements between rows of src; 0 = D */,
                  const int* gate = nullptr /* accumulate mode: device flag, 0 = skip */);
int colsum_partial_rows(int B);

int launch_sumsq(const float* g, long long n, double* partial, float* out_sumsq, cudaStream_t s);
constexpr int SUMSQ_MAX_RANGES = 18;
// sum of squares over up to SUMSQ_MAX_RANGES sub-ranges [begin, end) (element offsets, begins multiples of 4) of g;
// partial must hold SUMSQ_MAX_RANGES * 592 doubles
int launch_sumsq_ranges(const float* g, int n_ranges, const long long* begins, const long long* ends, double* partial,
                        float* out_sumsq, cudaStream_t s);
int launch_sumsq_fused(const float* row_gsq, int S, const float* gb_dec, int D, float* out_sumsq, cudaStream_t s);

struct AdamArgs {
  float* W_enc_t; float* b_enc; float* W_dec; float* b_dec;
  const float* gW_enc_t; const float* gb_enc; const float* gW_dec; const float* gb_dec;
  float* m; float* v;          // flat, same order/offsets as the gradient bucket
  __half* shadow16;            // fp16 copy of W_enc_t for the tensor-core screen (may be null)
  float* wnorm_sq_max;         // device scalar: max_j ||W_enc_t[j]||^2 after the update (zeroed by the launcher)
  float* wnorm_rows;           // [S] c_j = max(||w_j||, ||fp16(w_j)||) after the update, rounded up (may be null)
  float* rho;                  // device scalar: max_j ||w_j - fp16(w_j)|| / c_j (zeroed by the launcher)
  float* bias_abs_max;         // device scalar: max_j |b_enc[j]| after the update (zeroed by the launcher)
  int D, S;
  float lr, beta1, beta2, eps, bc1, bc2_sqrt;
  float max_norm; float grad_scale; const float* gnorm_sq;
  int renorm_w_dec;
  float* gnorm_out;            // optional: clipped-from norm (what clip_grad_norm_ returns)
  int row_begin, row_end;      // dictionary rows this call updates (sharded optimizer; default all)
  int b_enc_separately;        // 1: update the whole b_enc vector with a separate kernel (not only [row_begin, row_end))
  int parts;                   // 1: W_enc_t rows + b_enc (+ the screen's fp16 copy and norms), 2: W_dec rows + b_dec, 3: both;
                               // + 4: rows only (no whole-vector bias kernels), + 8: keep (do not zero) the screen's
                               // dictionary-wide maxima -- a sharded optimizer that owns several row ranges calls once
                               // per range: first call 3, the others 3 + 4 + 8
  int small_blocks;            // 64-thread blocks (fit beside a resident screen CTA: a parts == 2 launch on a side stream)
};
int launch_adam(const AdamArgs& a, cudaStream_t s);

struct FinalizeArgs {
  const float* row_sse; const float* row_l1; const float* row_l0; int B; int D;
  double inv_bd; double inv_b; float l1_coeff;
  const float* aux_loss;   // device scalar or null
  const int* n_dead;       // device scalar or null
  float* losses;           // [8]: mse, aux, sparsity, l0, l1, n_dead, loss, reserved
};
int launch_finalize(const FinalizeArgs& a, cudaStream_t s);

int launch_dead_update(long long* toks, int* active, int S, long long batch_tokens, long long threshold,
                       int* dead_list, int* n_dead, int* block_totals /* [ceil(S/1024)] scratch */, cudaStream_t s);

int launch_densify(const int* idx, const float* val, int B, int K, int S, float* out, cudaStream_t s);
int launch_add_rows(const float* a, const float* b, long long n, float* out, cudaStream_t s);  // out = a + b

// ---- dense_kernels.cu (ReLU / dense path) ----------------------------------------------------------------
int launch_transpose_split(const float* src, int R, int C, float scale, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo,
                           long long ldr, int ones_row, int C_pad, cudaStream_t s, __nv_bfloat16* dst_lo2 = nullptr,
                           const int* gate = nullptr, long long src_ld = 0 /* row pitch of src, 0 = C */);
int launch_dense_resid(float* xhat, const float* x, int B, int D, float grad_scale, float* row_sse, __nv_bfloat16* g_hi,
                       __nv_bfloat16* g_lo, cudaStream_t s, __nv_bfloat16* g_lo2 = nullptr);

int launch_project_rows(float* g, const float* w, int rows, int D, cudaStream_t s);
int launch_join_bf16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long n, float* out, cudaStream_t s,
                     const __nv_bfloat16* lo2 = nullptr);

// ---- aux_kernels.cu -------------------------------------------------------------------------------------
struct AuxArgs {
  const float* x; const float* resid; const float* W_enc_t; const float* b_enc; const float* W_dec;
  const float* b_dec;
  const int* dead_list; const int* n_dead;
  int B, D, S, k_aux; float alpha; float inv_bd;   // inv_bd = 1 / (B_global * D)
  int remove_parallel;
  float* h_aux;            // [B, S] scratch: pre-acts of dead latents -> f_aux -> dh_aux
  unsigned char* mask_aux; // [B, S] scratch
  float* r_aux;            // [B, D] scratch: x_hat_aux - e, then G_aux
  float* row_sse_aux;      // [B]
  float* aux_loss;         // device scalar out
  float* gW_enc_t; float* gb_enc; float* gW_dec;   // rows of dead latents are (over)written
  float* colsum_partial; float* gb_dec;            // gb_dec += sum_b G_aux
  float* aux_colpart;      // [32, S] scratch for the gb_enc column sums
  float* row_gsq;          // optional [d_sae]: refreshed for the dead atoms whose gradient rows are overwritten
  // tensor-core path (null tc_we[0] => the fp32 CUDA-core tiles): bf16 piece buffers, 3 pieces each
  int nterms;              // 3 or 6
  int num_sms;
  long long ldb;           // row pitch of the batch-major transposed operands (>= B, multiple of 8)
  long long ldc;           // row pitch of the [B, cap] piece matrices (cap rounded up to 64)
  __nv_bfloat16* tc_we[3];   // W_enc_t[L]      [cap, D]
  __nv_bfloat16* tc_wd[3];   // W_dec[L]        [cap, D]
  __nv_bfloat16* tc_wdT[3];  // W_dec[L]^T      [D, ldc]
  __nv_bfloat16* tc_x[3];    // x               [B, D]
  __nv_bfloat16* tc_xT[3];   // x^T             [D, ldb]
  __nv_bfloat16* tc_f[3];    // f_aux           [B, ldc]
  __nv_bfloat16* tc_fT[3];   // f_aux^T, later dh_aux^T   [cap, ldb]
  __nv_bfloat16* tc_r[3];    // r_aux           [B, D]
  __nv_bfloat16* tc_rT[3];   // r_aux^T         [D, ldb]
};
int launch_aux_forward(const AuxArgs& a, cudaStream_t s);
int launch_aux_backward(const AuxArgs& a, cudaStream_t s);
size_t aux_colpart_bytes(int cap);

}  // namespace sb
