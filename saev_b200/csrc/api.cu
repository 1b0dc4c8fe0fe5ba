// extern "C" surface of libsaev_b200.so (see include/saev_b200.h for the contract).
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#include <new>

#include "../../include/saev_b200.h"
#include "kernels.h"

using namespace sb;

namespace sb {
unsigned long long g_launch_count = 0;
}

namespace {

constexpr size_t ALIGN = 256;
inline size_t align_up(size_t x) { return (x + ALIGN - 1) & ~(ALIGN - 1); }

struct Workspace {
  size_t shadow_hi, x_hi, cand, tau_keys, cand_cnt, row_norm, row_dx, row_scale, unsafe_list, wnorm_rows, guess_hist, guess_L, dh, row_sse, row_l1, row_l0, row_sse_aux, feat_count, feat_off, cursor,
      entries, heavy_list, active, dead_list, scalars, block_totals, row_gsq, colsum_partial, sumsq_partial, h_aux, mask_aux, r_aux, aux_colpart, total;
  size_t btk;  // BatchTopK selection scratch (batch_topk_scratch_bytes)
  size_t yblk; // dense path with Matryoshka prefixes: [P][B][D] fp32 partial decodes of the prefix blocks
  size_t sfx;  // Matryoshka: [B, max_prefixes, D] suffix sums of the per-prefix residuals
  // tensor-core AuxK path: bf16 piece buffers (3 pieces each, see AuxArgs)
  size_t tc_we[3], tc_wd[3], tc_wdT[3], tc_x[3], tc_xT[3], tc_f[3], tc_fT[3], tc_r[3], tc_rT[3];
  long long ldc = 0;
  bool aux_tc = false;
  // dense (ReLU) path: bf16 (hi, lo) operand pairs
  size_t w_enc_lo, w_dec_hi, w_dec_lo, w_decT_hi, w_decT_lo, x_lo, xT_hi, xT_lo, g_hi, g_lo, gT_hi, gT_lo, f_hi, f_lo,
      fT_hi, fT_lo, dhT_hi, dhT_lo;
  // third pieces of the same operands (6-term split, fp32-class accuracy)
  size_t w_enc_l2, w_dec_l2, w_decT_l2, x_l2, xT_l2, g_l2, gT_l2, f_l2, fT_l2, dhT_l2;
  long long ldb = 0;  // row pitch of the batch-major transposed operands (max_batch rounded up to 8)
};

}  // namespace

struct saev_b200_handle {
  saev_b200_cfg cfg;
  int device = 0;
  int num_sms = 148;
  int aux_cap = 0;
  int max_pairs = 0;       // co-resident CTA pairs for the cta_group::2 screen
  int reserved_pairs = 0;  // SM pairs the screen leaves idle (saev_b200_set_reserved_sms)
  // AuxK path selection: the dead-latent count of an EARLIER step, read back asynchronously (never waited for).  Both
  // AuxK implementations are correct for any count; the lagged value only picks the cheaper one (the tensor-core
  // path launches worst-case grids, ~0.2 ms per step of empty CTAs when nothing is dead).
  int* nd_host = nullptr;        // pinned
  cudaEvent_t nd_event = nullptr;
  bool nd_pending = false;
  int n_dead_lagged = 0;
  bool aux_tc_step = false;      // path chosen by the last forward (backward must match)
  bool fuse_dh = true;           // d loss / d h computed inside the weight-gradient kernel instead of a second gather pass
                                 // of the decode kernel (c3: decode 0.47 -> 0.27 ms, wgrad 0.62 -> 0.74 ms);
                                 // SAEV_B200_FUSE_DH=0 restores the two-pass decode
  bool dh_fused_fwd = false;     // the last training forward left dh to the backward
  bool aux_tc_always = false;    // SAEV_B200_AUX=tc: no selection (tests pin each path)
  bool force_repair = false;     // SAEV_B200_FORCE_REPAIR=1: every row is re-done by the exact top-k path (tests)
  float guess_quantile = GUESS_QUANTILE;  // threshold guess of the screen (kernels.h); SAEV_B200_GUESS=0 turns it off,
  float guess_safety = GUESS_SAFETY;      // SAEV_B200_GUESS_Q / SAEV_B200_GUESS_S override the two parameters
  int guess_parity = 0;          // which half of the ratio histogram the current forward fills
  int dense_terms = 6;     // bf16 split of the dense (ReLU) contractions: 6 = three pieces per operand (fp32-class
                           // accuracy), 3 = two pieces (~2^-16 of sum |a b|, half the tensor work); SAEV_B200_DENSE_TERMS
  Workspace ws;
  bool last_forward_training = false;
  bool last_forward_tracked = false;
  bool row_gsq_valid = false;
  int shard_begin = 0, shard_end = 0;  // optimizer shard (rows of the dictionary); end == 0 => all rows
  int max_prefixes = 1;                // Matryoshka: most prefix cuts a step may use (cfg.max_prefixes)
  PrefixCuts pf;                       // cuts of the current step (pf.n == 1: plain single-prefix objective)
  PrefixCuts pf_fwd;                   // cuts the last forward used (what backward must use)  // the last backward left per-atom ||g||^2 partials in the workspace
  mutable char err[512];
  // optional per-stage CUDA-event timing (saev_b200_profile_*)
  bool prof_on = false;
  int prof_n = 0;
  static constexpr int PROF_CAP = 8192;
  cudaEvent_t* prof_beg = nullptr;
  cudaEvent_t* prof_end = nullptr;
  unsigned char* prof_stage = nullptr;
};

namespace {

thread_local char g_err[512] = "";

int fail(const saev_b200_handle* h, int code, const char* fmt, const char* detail = "") {
  char* dst = h ? h->err : g_err;
  const int n = snprintf(dst, 400, fmt, detail);
  // a failed launch usually has a CUDA error behind it: say which (and clear it, it has been reported)
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess && n > 0 && n < 400) snprintf(dst + n, 512 - n, " [CUDA: %s]", cudaGetErrorString(e));
  return code;
}

int check_cuda(const saev_b200_handle* h, const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  char buf[400];
  snprintf(buf, sizeof(buf), "%s: %s", where, cudaGetErrorString(e));
  return fail(h, 100, "CUDA error at %s", buf);
}

Workspace plan_workspace(const saev_b200_cfg& c, int aux_cap, int max_pairs, int dense_terms) {
  Workspace w;
  const size_t S = c.d_sae, D = c.d_model, B = c.max_batch, K = c.top_k;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += align_up(bytes);
    return at;
  };
  w.shadow_hi = take(S * D * 2);
  w.x_hi = take(B * D * 2);
  const bool relu = c.act_kind == SAEV_B200_ACT_RELU;
  w.ldb = static_cast<long long>((B + 7) / 8 * 8);
  if (relu) {
    const size_t LB = static_cast<size_t>(w.ldb);
    w.w_enc_lo = take(S * D * 2);
    w.w_dec_hi = take(S * D * 2);
    w.w_dec_lo = take(S * D * 2);
    w.w_decT_hi = take(D * S * 2);
    w.w_decT_lo = take(D * S * 2);
    w.x_lo = take(B * D * 2);
    w.xT_hi = take((D + 16) * LB * 2);
    w.xT_lo = take((D + 16) * LB * 2);
    // (with Matryoshka prefixes: one G = grad_scale * suffix-sum residual per prefix block, stacked)
    const size_t PF = c.max_prefixes > 1 ? static_cast<size_t>(c.max_prefixes) : 1;
    w.g_hi = take(PF * B * D * 2);
    w.g_lo = take(PF * B * D * 2);
    w.gT_hi = take(PF * D * LB * 2);
    w.gT_lo = take(PF * D * LB * 2);
    w.yblk = take(PF > 1 ? PF * B * D * 4 : 0);
    w.f_hi = take(B * S * 2);
    w.f_lo = take(B * S * 2);
    w.fT_hi = take(S * LB * 2);
    w.fT_lo = take(S * LB * 2);
    w.dhT_hi = take(S * LB * 2);
    w.dhT_lo = take(S * LB * 2);
    const bool third = dense_terms == 6;
    w.w_enc_l2 = take(third ? S * D * 2 : 0);
    w.w_dec_l2 = take(third ? S * D * 2 : 0);
    w.w_decT_l2 = take(third ? D * S * 2 : 0);
    w.x_l2 = take(third ? B * D * 2 : 0);
    w.xT_l2 = take(third ? (D + 16) * LB * 2 : 0);
    w.g_l2 = take(third ? PF * B * D * 2 : 0);
    w.gT_l2 = take(third ? PF * D * LB * 2 : 0);
    w.f_l2 = take(third ? B * S * 2 : 0);
    w.fT_l2 = take(third ? S * LB * 2 : 0);
    w.dhT_l2 = take(third ? S * LB * 2 : 0);
  }
  if (relu) {
    w.cand = w.tau_keys = w.cand_cnt = w.row_norm = w.row_dx = w.row_scale = w.unsafe_list = w.wnorm_rows = w.guess_hist = w.guess_L = o;
  } else {
    // candidate lists of the top-k screen: rows padded to 256, `nlists` lists per row; the list count depends on the
    // batch size (how many CTA-pair ranges touch one row block)
    size_t row_lists = 256, cand_bytes = 0;
    const int mp_max = static_cast<int>((B + 255) / 256);
    for (int mp = 1; mp <= mp_max; ++mp) {
      const Encode2Plan pl = encode2_plan(mp * 256, static_cast<int>(S), max_pairs > 0 ? max_pairs : 1);
      const size_t need = static_cast<size_t>(mp) * 256 * pl.nlists;
      if (need > row_lists) row_lists = need;
      if (need * ENCODE2_CAPG * 8 > cand_bytes) cand_bytes = need * ENCODE2_CAPG * 8;
    }
    w.cand = take(cand_bytes);
    w.tau_keys = take((B + 255) / 256 * 256 * 4);
    w.cand_cnt = take(row_lists * 4);
    w.row_norm = take((B + 255) / 256 * 256 * 4);
    w.row_dx = take((B + 255) / 256 * 256 * 4);
    w.row_scale = take((B + 255) / 256 * 256 * 4);
    w.unsafe_list = take(B * 4);
    w.wnorm_rows = take(S * 4);
    w.guess_hist = take(2 * GUESS_BINS * 4);
    w.guess_L = take((B + 255) / 256 * 256 * 4);
  }
  w.dh = take(B * K * 4);
  w.row_sse = take(B * 4);
  w.row_l1 = take(B * 4);
  w.row_l0 = take(B * 4);
  w.row_sse_aux = take(B * 4);
  w.feat_count = take(S * 4);
  w.feat_off = take((S + 1) * 4);
  w.cursor = take(S * 4);
  w.entries = take(B * K * 4);
  w.active = take(S * 4);
  w.dead_list = take(S * 4);
  w.scalars = take(SC_SLOTS * 4);  // see ScalarSlot (kernels.h)
  w.btk = take(relu ? 0 : batch_topk_scratch_bytes(c.max_batch));
  w.sfx = take(c.max_prefixes > 1 ? B * static_cast<size_t>(c.max_prefixes) * D * 4 : 0);
  w.block_totals = take(((S + 1023) / 1024) * 4);
  w.heavy_list = take((B * K / WGRAD_HEAVY_ENTRIES + 2) * 4);  // [0] = count, then the atoms
  w.row_gsq = take(S * 4);
  w.colsum_partial = take(static_cast<size_t>(colsum_partial_rows(static_cast<int>(B))) * D * 4);
  w.sumsq_partial = take(SUMSQ_MAX_RANGES * 1024 * 8);
  if (c.aux_kind == SAEV_B200_AUX_AUXK) {
    w.h_aux = take(B * static_cast<size_t>(aux_cap) * 4);
    w.mask_aux = take(B * static_cast<size_t>(aux_cap));
    w.r_aux = take(B * D * 4);
    w.aux_colpart = take(aux_colpart_bytes(aux_cap));
    const char* e = getenv("SAEV_B200_AUX");  // "sgemm": only the fp32 CUDA-core tiles (no tensor-core scratch)
    w.aux_tc = !(e && e[0] == 's');
    w.ldc = (static_cast<long long>(aux_cap) + 63) / 64 * 64;
    if (w.aux_tc) {
      const size_t CAP = static_cast<size_t>(aux_cap), LC = static_cast<size_t>(w.ldc), LB = static_cast<size_t>(w.ldb);
      for (int i = 0; i < 3; ++i) {
        w.tc_we[i] = take(CAP * D * 2);
        w.tc_wd[i] = take(CAP * D * 2);
        w.tc_wdT[i] = take(D * LC * 2);
        w.tc_x[i] = take(B * D * 2);
        w.tc_xT[i] = take(D * LB * 2);
        w.tc_f[i] = take(B * LC * 2);
        w.tc_fT[i] = take(CAP * LB * 2);
        w.tc_r[i] = take(B * D * 2);
        w.tc_rT[i] = take(D * LB * 2);
      }
    }
  } else {
    w.h_aux = w.mask_aux = w.r_aux = w.aux_colpart = o;
  }
  w.total = o;
  return w;
}

// RAII stage timer: records an event pair around a group of launches when profiling is enabled.
struct StageTimer {
  saev_b200_handle* h;
  cudaStream_t s;
  int slot = -1;
  StageTimer(saev_b200_handle* h_, int stage, cudaStream_t s_) : h(h_), s(s_) {
    if (!h->prof_on || h->prof_n >= saev_b200_handle::PROF_CAP) return;
    slot = h->prof_n++;
    h->prof_stage[slot] = static_cast<unsigned char>(stage);
    if (!h->prof_beg[slot]) {
      cudaEventCreate(&h->prof_beg[slot]);
      cudaEventCreate(&h->prof_end[slot]);
    }
    cudaEventRecord(h->prof_beg[slot], s);
  }
  ~StageTimer() {
    if (slot >= 0) cudaEventRecord(h->prof_end[slot], s);
  }
};

// PyTorch runs backward passes on autograd worker threads where no CUDA context may be current yet; the driver-API
// tensor-map encoder (unlike runtime launches) needs one.  cudaSetDevice binds the primary context to this thread.
inline void bind_context(const saev_b200_handle* h) { cudaSetDevice(h->device); }

template <typename T>
inline T* at(void* ws, size_t off) {
  return reinterpret_cast<T*>(static_cast<char*>(ws) + off);
}

// piece buffers / settings of the tensor-core AuxK path
void fill_aux_tc(const saev_b200_handle* h, void* workspace, AuxArgs& a) {
  const Workspace& w = h->ws;
  a.nterms = h->dense_terms;
  a.num_sms = h->num_sms;
  a.ldb = w.ldb;
  a.ldc = w.ldc;
  for (int i = 0; i < 3; ++i) {
    auto p = [&](const size_t (&off)[3]) { return h->aux_tc_step ? at<__nv_bfloat16>(workspace, off[i]) : nullptr; };
    a.tc_we[i] = p(w.tc_we);
    a.tc_wd[i] = p(w.tc_wd);
    a.tc_wdT[i] = p(w.tc_wdT);
    a.tc_x[i] = p(w.tc_x);
    a.tc_xT[i] = p(w.tc_xT);
    a.tc_f[i] = p(w.tc_f);
    a.tc_fT[i] = p(w.tc_fT);
    a.tc_r[i] = p(w.tc_r);
    a.tc_rT[i] = p(w.tc_rT);
  }
}

// Error-compensated bf16 split product on the tcgen05 kernel: out-of-line helper for the dense (ReLU) path.
EncodeGemmArgs dense_gemm(const saev_b200_handle* h, const __nv_bfloat16* A_hi, const __nv_bfloat16* A_lo,
                          const __nv_bfloat16* A_l2, long long lda, const __nv_bfloat16* B_hi, const __nv_bfloat16* B_lo,
                          const __nv_bfloat16* B_l2, long long ldb, int M, int N, int K, int epilogue) {
  EncodeGemmArgs g;
  g.A_hi = A_hi;
  g.A_lo = A_lo;
  g.A_lo2 = A_l2;
  g.B_hi = B_hi;
  g.B_lo = B_lo;
  g.B_lo2 = B_l2;
  g.lda = lda;
  g.ldb = ldb;
  g.nterms = h->dense_terms;
  g.k_chunk_blocks = 8;  // 512-wide K chunks: bounds the truncation bias of the tensor-core accumulator (encode_gemm.cu)
  g.M = M;
  g.N = N;
  g.K = K;
  g.epilogue = epilogue;
  g.nsplit = 0;
  g.num_sms = h->num_sms;
  return g;
}

// Phase A of the objective forward for the ReLU activation (saev modeling.py:150-156, 343-409; objectives.py:133-151)
int forward_relu_phase_a(saev_b200_handle* h, const float* x, int B, long long tokens_global, const float* W_enc_t,
                         const float* b_enc, const float* W_dec, const float* b_dec, int training, float* resid,
                         void* workspace, cudaStream_t s) {
  const saev_b200_cfg& c = h->cfg;
  const Workspace& w = h->ws;
  const int D = c.d_model, S = c.d_sae;
  auto bf = [&](size_t off) { return at<__nv_bfloat16>(workspace, off); };
  const bool third = h->dense_terms == 6;
  auto b3 = [&](size_t off) { return third ? at<__nv_bfloat16>(workspace, off) : nullptr; };
  const long long SD = static_cast<long long>(S) * D;
  {
    StageTimer tm(h, SAEV_B200_STAGE_PREP, s);
    // operands of this step's contractions from the fp32 master weights (W_dec is already normalised)
    if (launch_split_bf16(W_enc_t, bf(w.shadow_hi), bf(w.w_enc_lo), SD, s, b3(w.w_enc_l2)) ||
        launch_split_bf16(W_dec, bf(w.w_dec_hi), bf(w.w_dec_lo), SD, s, b3(w.w_dec_l2)) ||
        launch_transpose_split(W_dec, S, D, 1.f, bf(w.w_decT_hi), bf(w.w_decT_lo), S, 0, D, s, b3(w.w_decT_l2)) ||
        launch_split_bf16(x, bf(w.x_hi), bf(w.x_lo), static_cast<long long>(B) * D, s, b3(w.x_l2)))
      return fail(h, 41, "forward(relu): operand split launch failed%s");
    cudaMemsetAsync(at<float>(workspace, w.row_l1), 0, static_cast<size_t>(B) * 4, s);
    cudaMemsetAsync(at<float>(workspace, w.row_l0), 0, static_cast<size_t>(B) * 4, s);
  }
  {
    StageTimer tm(h, SAEV_B200_STAGE_ENCODE_GEMM, s);
    // f = relu(x W_enc + b_enc)
    EncodeGemmArgs g = dense_gemm(h, bf(w.x_hi), bf(w.x_lo), b3(w.x_l2), D, bf(w.shadow_hi), bf(w.w_enc_lo),
                                  b3(w.w_enc_l2), D, B, S, D, 2);
    g.bias = b_enc;
    g.f_hi = bf(w.f_hi);
    g.f_lo = bf(w.f_lo);
    g.f_lo2 = b3(w.f_l2);
    g.ldf = S;
    g.t_hi = training ? bf(w.fT_hi) : nullptr;
    g.t_lo = training ? bf(w.fT_lo) : nullptr;
    g.t_lo2 = training ? b3(w.fT_l2) : nullptr;
    g.ldt = w.ldb;
    g.row_l1 = at<float>(workspace, w.row_l1);
    g.row_l0 = at<float>(workspace, w.row_l0);
    g.active = at<int>(workspace, w.active);
    if (int rc = launch_encode_gemm(g, s)) {
      char buf[64];
      snprintf(buf, sizeof(buf), "%d", rc);
      return fail(h, 42, "forward(relu): encoder contraction launch failed (code %s)", buf);
    }
  }
  {
    StageTimer tm(h, SAEV_B200_STAGE_DECODE, s);
    const int P = h->pf.n;
    h->pf_fwd = h->pf;
    const float gs = static_cast<float>(2.0 / (static_cast<double>(tokens_global) * P * D));
    const long long LB = w.ldb, BD = static_cast<long long>(B) * D;
    if (P <= 1) {
      // x_hat = f W_dec + b_dec, then r = x_hat - x, SSE partials and G = 2 r / (B D)
      EncodeGemmArgs g = dense_gemm(h, bf(w.f_hi), bf(w.f_lo), b3(w.f_l2), S, bf(w.w_decT_hi), bf(w.w_decT_lo),
                                    b3(w.w_decT_l2), S, B, D, S, 1);
      g.bias = b_dec;
      g.out = resid;
      g.ldo = D;
      if (int rc = launch_encode_gemm(g, s)) {
        char buf[64];
        snprintf(buf, sizeof(buf), "%d", rc);
        return fail(h, 44, "forward(relu): decoder contraction launch failed (code %s)", buf);
      }
      if (launch_dense_resid(resid, x, B, D, gs, at<float>(workspace, w.row_sse), training ? bf(w.g_hi) : nullptr,
                             training ? bf(w.g_lo) : nullptr, s, training ? b3(w.g_l2) : nullptr))
        return fail(h, 44, "forward(relu): residual launch failed%s");
      if (training && launch_transpose_split(resid, B, D, gs, bf(w.gT_hi), bf(w.gT_lo), w.ldb, 0, D, s, b3(w.gT_l2)))
        return fail(h, 44, "forward(relu): G^T launch failed%s");
    } else {
      // Matryoshka (modeling.py:364-406): one decoder contraction per prefix block, over the dictionary columns
      // [cut_{c-1}, cut_c) only -- a window on the same K-major operands: the tensor maps end at cut_c and the TMA
      // coordinates start at the block's first 8-aligned column (16-byte aligned spans along the contraction); the <= 7
      // columns in front of it and b_dec are added in fp32 by the residual kernel
      float* y = at<float>(workspace, w.yblk);
      unsigned int tensor_mask = 0u;
      for (int cb = 0; cb < P; ++cb) {
        const int k0 = cb > 0 ? h->pf.cut[cb - 1] : 0, k1 = h->pf.cut[cb];
        const int ka = (k0 + 7) & ~7;
        if (ka >= k1) continue;
        tensor_mask |= 1u << cb;
        EncodeGemmArgs g = dense_gemm(h, bf(w.f_hi), bf(w.f_lo), b3(w.f_l2), S, bf(w.w_decT_hi), bf(w.w_decT_lo),
                                      b3(w.w_decT_l2), S, B, D, k1, 1);
        g.k_begin = ka;
        g.out = y + cb * BD;
        g.ldo = D;
        if (int rc = launch_encode_gemm(g, s)) {
          char buf[64];
          snprintf(buf, sizeof(buf), "%d", rc);
          return fail(h, 44, "forward(relu): prefix-block decoder contraction launch failed (code %s)", buf);
        }
      }
      float* sfx = at<float>(workspace, w.sfx);
      if (launch_dense_prefix_resid(y, x, B, D, h->pf, tensor_mask, bf(w.f_hi), bf(w.f_lo), b3(w.f_l2), S, W_dec, b_dec, gs,
                                    resid, sfx, at<float>(workspace, w.row_sse), training ? bf(w.g_hi) : nullptr,
                                    training ? bf(w.g_lo) : nullptr, training ? b3(w.g_l2) : nullptr, s))
        return fail(h, 44, "forward(relu): prefix residual launch failed%s");
      if (training)
        for (int cb = 0; cb < P; ++cb)  // G_c^T [D, ldb], the operand of the per-block W_dec gradient
          if (launch_transpose_split(sfx + static_cast<long long>(cb) * D, B, D, gs, bf(w.gT_hi) + cb * D * LB,
                                     bf(w.gT_lo) + cb * D * LB, w.ldb, 0, D, s, third ? bf(w.gT_l2) + cb * D * LB : nullptr,
                                     nullptr, static_cast<long long>(P) * D))
            return fail(h, 44, "forward(relu): G^T launch failed%s");
    }
  }
  return 0;
}

// Dense gradients of the ReLU path (autograd of the above; explicit formulas in SURVEY.md appendix A step 9)
int backward_relu(saev_b200_handle* h, const float* x, int B, long long tokens_global, const float* W_dec,
                  float* gW_enc_t, float* gb_enc, float* gW_dec, void* workspace, cudaStream_t s) {
  const saev_b200_cfg& c = h->cfg;
  const Workspace& w = h->ws;
  const int D = c.d_model, S = c.d_sae;
  auto bf = [&](size_t off) { return at<__nv_bfloat16>(workspace, off); };
  const bool third = h->dense_terms == 6;
  auto b3 = [&](size_t off) { return third ? at<__nv_bfloat16>(workspace, off) : nullptr; };
  StageTimer tm(h, SAEV_B200_STAGE_WGRAD, s);
  // dh = (f > 0) * (G W_dec^T + l1 / B), stored transposed as the operand of the W_enc gradient.  With Matryoshka
  // prefixes the columns of block c see G_c = the suffix sum of the per-prefix residual gradients: one contraction per
  // block over the dictionary columns [cut_{c-1}, cut_c)
  const int P = h->pf_fwd.n;
  const long long LB = w.ldb, BD = static_cast<long long>(B) * D;
  for (int cb = 0; cb < P; ++cb) {
    EncodeGemmArgs g3 = dense_gemm(h, bf(w.g_hi) + cb * BD, bf(w.g_lo) + cb * BD, third ? bf(w.g_l2) + cb * BD : nullptr, D,
                                   bf(w.w_dec_hi), bf(w.w_dec_lo), b3(w.w_dec_l2), D, B, P > 1 ? h->pf_fwd.cut[cb] : S, D, 3);
    g3.n_begin = cb > 0 ? h->pf_fwd.cut[cb - 1] : 0;
    g3.f_hi = bf(w.f_hi);
    g3.ldf = S;
    g3.t_hi = bf(w.dhT_hi);
    g3.t_lo = bf(w.dhT_lo);
    g3.t_lo2 = b3(w.dhT_l2);
    g3.ldt = w.ldb;
    g3.l1_over_b = c.l1_coeff != 0.f ? static_cast<float>(c.l1_coeff / static_cast<double>(tokens_global)) : 0.f;
    if (int rc = launch_encode_gemm(g3, s)) {
      char buf[32];
      snprintf(buf, sizeof(buf), "%d", rc);
      return fail(h, 52, "backward(relu): dh contraction launch failed (code %s)", buf);
    }
  }
  // x^T with an extra row of ones: column D of the next product is sum_b dh = gb_enc
  if (launch_transpose_split(x, B, D, 1.f, bf(w.xT_hi), bf(w.xT_lo), w.ldb, 1, D + 16, s, b3(w.xT_l2)))
    return fail(h, 52, "backward(relu): x^T launch failed%s");
  // gW_dec = f^T G  (per prefix block: the rows [cut_{c-1}, cut_c) of f^T against G_c^T)
  for (int cb = 0; cb < P; ++cb) {
    EncodeGemmArgs g4 = dense_gemm(h, bf(w.fT_hi), bf(w.fT_lo), b3(w.fT_l2), w.ldb, bf(w.gT_hi) + cb * D * LB,
                                   bf(w.gT_lo) + cb * D * LB, third ? bf(w.gT_l2) + cb * D * LB : nullptr, w.ldb,
                                   P > 1 ? h->pf_fwd.cut[cb] : S, D, B, 4);
    g4.m_begin = cb > 0 ? h->pf_fwd.cut[cb - 1] : 0;
    g4.out = gW_dec;
    g4.ldo = D;
    g4.n_main = D;
    if (launch_encode_gemm(g4, s)) return fail(h, 52, "backward(relu): gW_dec contraction launch failed%s");
  }
  if (c.remove_parallel_grads && launch_project_rows(gW_dec, W_dec, S, D, s))
    return fail(h, 52, "backward(relu): projection launch failed%s");
  // gW_enc_t = dh^T x ; gb_enc = dh^T 1
  EncodeGemmArgs g5 = dense_gemm(h, bf(w.dhT_hi), bf(w.dhT_lo), b3(w.dhT_l2), w.ldb, bf(w.xT_hi), bf(w.xT_lo),
                                 b3(w.xT_l2), w.ldb, S, D + 1, B, 4);
  g5.out = gW_enc_t;
  g5.ldo = D;
  g5.n_main = D;
  g5.extra = gb_enc;
  if (launch_encode_gemm(g5, s)) return fail(h, 52, "backward(relu): gW_enc contraction launch failed%s");
  return 0;
}

}  // namespace

extern "C" {

int saev_b200_abi_version(void) { return SAEV_B200_ABI_VERSION; }

uint64_t saev_b200_launch_count(void) { return sb::g_launch_count; }

const char* saev_b200_last_error(const saev_b200_handle* h) { return h ? h->err : g_err; }

int saev_b200_create(const saev_b200_cfg* cfg, saev_b200_handle** out) {
  if (!cfg || !out) return fail(nullptr, 1, "saev_b200_create: null argument%s");
  *out = nullptr;
  if (cfg->d_model <= 0 || cfg->d_sae <= 0 || cfg->max_batch <= 0)
    return fail(nullptr, 2, "saev_b200_create: d_model, d_sae and max_batch must be positive%s");
  if (cfg->d_model % 8 != 0 || cfg->d_model > 2048)
    return fail(nullptr, 2, "saev_b200_create: d_model must be a multiple of 8 and <= 2048%s");
  if (cfg->d_sae % 4 != 0) return fail(nullptr, 2, "saev_b200_create: d_sae must be a multiple of 4%s");
  if (cfg->act_kind != SAEV_B200_ACT_TOPK && cfg->act_kind != SAEV_B200_ACT_RELU)
    return fail(nullptr, 3, "saev_b200_create: unknown activation (TopK and ReLU have CUDA paths)%s");
  if (cfg->act_kind == SAEV_B200_ACT_TOPK) {
    if (cfg->top_k <= 0 || cfg->top_k > cfg->d_sae)
      return fail(nullptr, 2, "saev_b200_create: need 0 < top_k <= d_sae%s");
    if (cfg->top_k > encode2_max_top_k())
      return fail(nullptr, 3, "saev_b200_create: top_k > 128 is not supported by the screening kernel%s");
  } else if (cfg->d_sae % 8 != 0) {
    return fail(nullptr, 2, "saev_b200_create: the dense (ReLU) path needs d_sae to be a multiple of 8%s");
  }
  int dev = 0, cc_major = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(nullptr, 4, "saev_b200_create: no CUDA device (this library has no CPU path)%s");
  }
  cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cc_major != 10)
    return fail(nullptr, 4, "saev_b200_create: kernels are built for sm_100a (Blackwell B200) only%s");
  saev_b200_handle* h = new (std::nothrow) saev_b200_handle();
  if (!h) return fail(nullptr, 5, "saev_b200_create: out of host memory%s");
  h->cfg = *cfg;
  if (h->cfg.max_prefixes < 1) h->cfg.max_prefixes = 1;
  if (h->cfg.max_prefixes > MAX_PREFIXES) {
    delete h;
    return fail(nullptr, 2, "saev_b200_create: max_prefixes must be <= 32%s");
  }
  if (h->cfg.max_prefixes > 1 && cfg->act_kind == SAEV_B200_ACT_TOPK && cfg->top_k > 128) {
    delete h;
    return fail(nullptr, 3, "saev_b200_create: Matryoshka prefixes on the TopK path need top_k <= 128%s");
  }
  h->max_prefixes = h->cfg.max_prefixes;
  h->pf.n = 1;
  h->pf.cut[0] = cfg->d_sae;
  h->pf_fwd = h->pf;
  h->device = dev;
  h->num_sms = sms;
  h->aux_cap = (cfg->aux_cols_cap > 0 && cfg->aux_cols_cap < cfg->d_sae) ? cfg->aux_cols_cap : cfg->d_sae;
  h->max_pairs = encode2_max_pairs();
  if (cfg->act_kind == SAEV_B200_ACT_TOPK && h->max_pairs <= 0) {
    delete h;
    return fail(nullptr, 4, "saev_b200_create: the cta_group::2 screen kernel cannot be made resident on this device%s");
  }
  {
    const char* v = getenv("SAEV_B200_DENSE_TERMS");
    h->dense_terms = (v && v[0] == '3') ? 3 : 6;
  }
  h->ws = plan_workspace(h->cfg, h->aux_cap, h->max_pairs, h->dense_terms);
  {
    const char* v = getenv("SAEV_B200_FUSE_DH");
    h->fuse_dh = h->cfg.d_model <= 1024 && !(v && v[0] == '0');
  }
  {
    const char* v = getenv("SAEV_B200_AUX");  // "sgemm" / "tc": pin one AuxK implementation; default: pick per step
    h->aux_tc_always = v && v[0] == 't';
  }
  {
    const char* v = getenv("SAEV_B200_GUESS");
    if (v && v[0] == '0') h->guess_quantile = -1.f;
    if (const char* q = getenv("SAEV_B200_GUESS_Q")) h->guess_quantile = static_cast<float>(atof(q));
    if (const char* sfy = getenv("SAEV_B200_GUESS_S")) h->guess_safety = static_cast<float>(atof(sfy));
  }
  {
    const char* v = getenv("SAEV_B200_FORCE_REPAIR");  // test switch: every row takes the exact repair path
    h->force_repair = v && v[0] == '1';
  }
  if (cudaHostAlloc(reinterpret_cast<void**>(&h->nd_host), sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->nd_event, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    h->nd_host = nullptr;  // selection falls back to "always tensor-core"
  } else {
    *h->nd_host = 0;
  }
  h->err[0] = 0;
  *out = h;
  return 0;
}

int saev_b200_destroy(saev_b200_handle* h) {
  if (h && h->nd_host) cudaFreeHost(h->nd_host);
  if (h && h->nd_event) cudaEventDestroy(h->nd_event);
  if (h && h->prof_beg) {
    for (int i = 0; i < saev_b200_handle::PROF_CAP; ++i) {
      if (h->prof_beg[i]) cudaEventDestroy(h->prof_beg[i]);
      if (h->prof_end[i]) cudaEventDestroy(h->prof_end[i]);
    }
    delete[] h->prof_beg;
    delete[] h->prof_end;
    delete[] h->prof_stage;
  }
  delete h;
  return 0;
}

int saev_b200_profile_enable(saev_b200_handle* h, int32_t on) {
  if (on && !h->prof_beg) {
    h->prof_beg = new cudaEvent_t[saev_b200_handle::PROF_CAP]();
    h->prof_end = new cudaEvent_t[saev_b200_handle::PROF_CAP]();
    h->prof_stage = new unsigned char[saev_b200_handle::PROF_CAP]();
  }
  h->prof_on = on != 0;
  h->prof_n = 0;
  return 0;
}

int saev_b200_profile_read(saev_b200_handle* h, float* host_ms_sum, int32_t* host_count) {
  for (int i = 0; i < SAEV_B200_N_STAGES; ++i) {
    host_ms_sum[i] = 0.f;
    host_count[i] = 0;
  }
  for (int i = 0; i < h->prof_n; ++i) {
    if (cudaEventSynchronize(h->prof_end[i]) != cudaSuccess) return fail(h, 110, "profile_read: event sync failed%s");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof_beg[i], h->prof_end[i]) != cudaSuccess)
      return fail(h, 110, "profile_read: elapsed time failed%s");
    host_ms_sum[h->prof_stage[i]] += ms;
    host_count[h->prof_stage[i]] += 1;
  }
  h->prof_n = 0;
  return 0;
}

size_t saev_b200_workspace_bytes(const saev_b200_handle* h) { return h ? h->ws.total : 0; }

int32_t* saev_b200_active_flags(const saev_b200_handle* h, void* workspace) {
  return at<int32_t>(workspace, h->ws.active);
}
uint32_t* saev_b200_unsafe_rows(const saev_b200_handle* h, void* workspace) {
  return at<uint32_t>(workspace, h->ws.scalars) + SC_UNSAFE_TOTAL;
}

int saev_b200_sync_weights(saev_b200_handle* h, const float* W_enc_t, const float* b_enc, void* workspace,
                           void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n = static_cast<long long>(h->cfg.d_sae) * h->cfg.d_model;
  cudaMemsetAsync(at<char>(workspace, h->ws.scalars), 0, SC_SLOTS * 4, s);
  if (h->cfg.act_kind == SAEV_B200_ACT_TOPK) {  // (the dense path re-splits its operands every forward)
    // new weights: the threshold / norm ratios recorded so far say nothing about them (the next forward runs cold)
    cudaMemsetAsync(at<char>(workspace, h->ws.guess_hist), 0, 2 * GUESS_BINS * 4, s);
    if (launch_to_half(W_enc_t, at<__half>(workspace, h->ws.shadow_hi), n, s))
      return fail(h, 30, "sync_weights: fp16 copy launch failed%s");
    if (launch_row_sumsq_max(W_enc_t, h->cfg.d_sae, h->cfg.d_model, at<float>(workspace, h->ws.scalars) + SC_WNORM_SQ_MAX, s,
                             at<float>(workspace, h->ws.wnorm_rows), at<float>(workspace, h->ws.scalars) + SC_RHO))
      return fail(h, 30, "sync_weights: row-norm launch failed%s");
    if (launch_abs_max(b_enc, h->cfg.d_sae, at<float>(workspace, h->ws.scalars) + SC_BIAS_ABS_MAX, s))
      return fail(h, 30, "sync_weights: bias-max launch failed%s");
  }
  return check_cuda(h, "sync_weights");
}

int saev_b200_datapoint_init(saev_b200_handle* h, const float* acts, int64_t n_rows, const int64_t* src_row,
                             const float* noise, const int64_t* noise_row, float blend, int32_t tie_transpose, int32_t normalize,
                             float* mean_out, float* W_enc_t, const float* b_enc, float* W_dec, void* workspace,
                             void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!h || !acts || !src_row || !noise || !mean_out || !W_enc_t || !W_dec || !workspace)
    return fail(h, 32, "datapoint_init: null argument%s");
  if (n_rows <= 0 || !(blend >= 0.f && blend <= 1.f)) return fail(h, 32, "datapoint_init: need n_rows > 0 and 0 <= blend <= 1%s");
  const int D = h->cfg.d_model, S = h->cfg.d_sae;
  // mean over the sample rows (train.py:163), in chunks the column-sum scratch was sized for
  const int chunk = h->cfg.max_batch;
  for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
    const int rows = static_cast<int>(n_rows - r0 < chunk ? n_rows - r0 : chunk);
    if (launch_colsum(acts + r0 * D, rows, D, static_cast<float>(1.0 / static_cast<double>(n_rows)), r0 > 0 ? 1 : 0,
                      at<float>(workspace, h->ws.colsum_partial), mean_out, s))
      return fail(h, 33, "datapoint_init: mean launch failed%s");
  }
  if (launch_datapoint_init(acts, reinterpret_cast<const long long*>(src_row), mean_out, noise,
                            reinterpret_cast<const long long*>(noise_row), blend, tie_transpose,
                            normalize, S, D, W_enc_t, W_dec, s))
    return fail(h, 33, "datapoint_init: launch failed%s");
  return saev_b200_sync_weights(h, W_enc_t, b_enc, workspace, stream);
}

int saev_b200_normalize_w_dec(saev_b200_handle* h, float* W_dec, void* stream) {
  if (launch_normalize_rows(W_dec, h->cfg.d_sae, h->cfg.d_model, static_cast<cudaStream_t>(stream)))
    return fail(h, 31, "normalize_w_dec: launch failed%s");
  return check_cuda(h, "normalize_w_dec");
}

int saev_b200_forward(saev_b200_handle* h, int phase, const float* x, int32_t B, int64_t tokens_global,
                      const float* W_enc_t, const float* b_enc, const float* W_dec, const float* b_dec,
                      int64_t* toks_since_active, int32_t training, int32_t* topk_idx, float* topk_val,
                      float* resid, float* losses, void* workspace, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bind_context(h);
  const saev_b200_cfg& c = h->cfg;
  const Workspace& w = h->ws;
  if (B <= 0 || B > c.max_batch) return fail(h, 40, "forward: B out of range (0 < B <= cfg.max_batch)%s");
  if (tokens_global <= 0) tokens_global = B;
  h->row_gsq_valid = false;
  const int D = c.d_model, S = c.d_sae, K = c.top_k;
  int* scal_i = at<int>(workspace, w.scalars);
  float* aux_loss = at<float>(workspace, w.scalars) + SC_AUX_LOSS;
  const bool tracked = training && toks_since_active != nullptr;

  const bool do_screen = (phase & (SAEV_B200_PHASE_A | SAEV_B200_PHASE_A_SCREEN)) != 0;
  const bool do_rescore = (phase & (SAEV_B200_PHASE_A | SAEV_B200_PHASE_A_REST | SAEV_B200_PHASE_A_RESCORE)) != 0;
  const bool do_decode = (phase & (SAEV_B200_PHASE_A | SAEV_B200_PHASE_A_REST | SAEV_B200_PHASE_A_DECODE)) != 0;
  const bool do_rest = do_rescore || do_decode;
  if ((phase & (SAEV_B200_PHASE_A_SCREEN | SAEV_B200_PHASE_A_REST | SAEV_B200_PHASE_A_RESCORE | SAEV_B200_PHASE_A_DECODE)) &&
      c.act_kind == SAEV_B200_ACT_RELU)
    return fail(h, 40, "forward: the split phase A (screen / rest) exists for the TopK path only%s");
  if ((phase & SAEV_B200_PHASE_A) && c.act_kind == SAEV_B200_ACT_RELU) {
    cudaMemsetAsync(at<int>(workspace, w.active), 0, static_cast<size_t>(S) * 4, s);
    if (int rc = forward_relu_phase_a(h, x, B, tokens_global, W_enc_t, b_enc, W_dec, b_dec, training, resid, workspace, s))
      return rc;
    h->last_forward_training = training != 0;
    h->last_forward_tracked = false;
  } else if (do_screen || do_rest) {
    __half* x16 = at<__half>(workspace, w.x_hi);
    float* scal_f = at<float>(workspace, w.scalars);
    // `reserved_pairs` SM pairs are left idle (a data-parallel caller runs NCCL all-gathers beside this kernel)
    const Encode2Plan pl = encode2_plan(B, S, h->max_pairs > h->reserved_pairs ? h->max_pairs - h->reserved_pairs : 1);
    if (do_screen) {
      StageTimer tm(h, SAEV_B200_STAGE_PREP, s);
      // the threshold guess of this forward, from the ratios the previous one recorded (then that half of the
      // histogram is clear and becomes the one this forward fills)
      h->guess_parity ^= 1;
      if (launch_screen_guess(at<int>(workspace, w.guess_hist) + (h->guess_parity ^ 1) * GUESS_BINS, h->guess_quantile,
                              h->guess_safety, h->guess_quantile > 0.f ? 64 : 0x7fffffff, scal_f, s))
        return fail(h, 41, "forward: threshold-guess launch failed%s");
      if (launch_prep_x(x, B, D, x16, at<float>(workspace, w.row_norm), at<float>(workspace, w.row_dx),
                        at<float>(workspace, w.row_scale), scal_f, at<unsigned int>(workspace, w.tau_keys),
                        at<float>(workspace, w.guess_L), pl.m_pairs * 256, s))
        return fail(h, 41, "forward: prep_x launch failed%s");
    }

    EncodeGemmArgs g;
    g.A_hi = reinterpret_cast<const __nv_bfloat16*>(x16);  // (16-bit operands; the pair kernel reads them as fp16)
    g.B_hi = at<__nv_bfloat16>(workspace, w.shadow_hi);
    g.nterms = 1;
    g.M = B;
    g.N = S;
    g.K = D;
    g.bias = b_enc;
    g.top_k = K;
    g.row_norm = at<float>(workspace, w.row_norm);
    g.row_dx = at<float>(workspace, w.row_dx);
    g.row_scale = at<float>(workspace, w.row_scale);
    g.scalars = scal_f;
    g.col_norm = at<float>(workspace, w.wnorm_rows);
    g.cand_cnt = at<int>(workspace, w.cand_cnt);
    g.num_sms = h->num_sms;
    g.cand = at<char>(workspace, w.cand);
    g.tau_keys = at<unsigned int>(workspace, w.tau_keys);
    g.tau_preset = 1;  // prep_x_kernel wrote every row's starting threshold
    g.nsplit = pl.nlists;  // what the re-score kernel merges per row
    if (do_screen) {
      StageTimer tm(h, SAEV_B200_STAGE_ENCODE_GEMM, s);
      if (int rc = launch_encode_gemm2(g, pl, s)) {
        char buf[64];
        snprintf(buf, sizeof(buf), "%d", rc);
        return fail(h, 42, "forward: encode GEMM launch failed (code %s)", buf);
      }
    }

    if (do_rescore) {
    cudaMemsetAsync(at<int>(workspace, w.active), 0, static_cast<size_t>(S) * 4, s);
    if (training) cudaMemsetAsync(at<int>(workspace, w.feat_count), 0, static_cast<size_t>(S) * 4, s);
    RescoreArgs r;
    r.cand = g.cand;
    r.cand_cnt = g.cand_cnt;
    r.cand_stride = ENCODE2_CAPG;
    r.nsplit = g.nsplit;
    r.row_norm = g.row_norm;
    r.row_dx = g.row_dx;
    r.row_scale = g.row_scale;
    r.scalars = scal_f;
    r.col_norm = g.col_norm;
    r.guess_L = at<float>(workspace, w.guess_L);
    r.guess_hist = at<int>(workspace, w.guess_hist) + h->guess_parity * GUESS_BINS;
    r.unsafe_list = at<int>(workspace, w.unsafe_list);
    r.force_unsafe = h->force_repair ? 1 : 0;
    r.x = x;
    r.W_enc_t = W_enc_t;
    r.b_enc = b_enc;
    r.B = B;
    r.D = D;
    r.S = S;
    r.K = K;
    r.topk_idx = topk_idx;
    r.topk_val = topk_val;
    r.feat_count = training ? at<int>(workspace, w.feat_count) : nullptr;
    r.active = at<int>(workspace, w.active);
    {
      StageTimer tm(h, SAEV_B200_STAGE_RESCORE, s);
      if (launch_rescore_topk(r, s)) return fail(h, 43, "forward: rescore launch failed%s");
      if (launch_repair_topk(r, s)) return fail(h, 43, "forward: exact repair launch failed%s");
    }
    }  // do_rescore
    if (do_decode) {
    DecodeArgs d;
    d.x = x;
    d.topk_idx = topk_idx;
    d.topk_val = topk_val;
    d.W_dec = W_dec;
    d.b_dec = b_dec;
    d.B = B;
    d.D = D;
    d.K = K;
    const int P = h->pf.n;
    h->pf_fwd = h->pf;
    d.grad_scale = static_cast<float>(2.0 / (static_cast<double>(tokens_global) * P * D));
    d.l1_over_b = c.l1_coeff != 0.f ? static_cast<float>(c.l1_coeff / static_cast<double>(tokens_global)) : 0.f;
    d.resid = resid;
    h->dh_fused_fwd = training && h->fuse_dh && P == 1;
    d.dh = (training && !h->dh_fused_fwd) ? at<float>(workspace, w.dh) : nullptr;
    d.row_sse = at<float>(workspace, w.row_sse);
    d.row_l1 = at<float>(workspace, w.row_l1);
    d.row_l0 = at<float>(workspace, w.row_l0);
    {
      StageTimer tm(h, SAEV_B200_STAGE_DECODE, s);
      if (P > 1 ? launch_decode_prefix(d, h->pf, at<float>(workspace, w.sfx), s) : launch_decode(d, s))
        return fail(h, 44, "forward: decode launch failed%s");
    }
    h->last_forward_training = training != 0;
    h->last_forward_tracked = false;
    }  // do_decode
  }

  if (phase & SAEV_B200_PHASE_B) {
    StageTimer tm(h, SAEV_B200_STAGE_LOSS, s);
    bool aux_live = false;
    if (tracked) {
      if (launch_dead_update(reinterpret_cast<long long*>(toks_since_active), at<int>(workspace, w.active), S,
                             tokens_global, c.dead_threshold_tokens, at<int>(workspace, w.dead_list), scal_i,
                             at<int>(workspace, w.block_totals), s))
        return fail(h, 45, "forward: dead tracker launch failed%s");
      h->last_forward_tracked = true;
      if (h->nd_host != nullptr) {
        if (h->nd_pending && cudaEventQuery(h->nd_event) == cudaSuccess) {
          h->n_dead_lagged = *h->nd_host;
          h->nd_pending = false;
        } else {
          cudaGetLastError();  // cudaErrorNotReady is not an error
        }
        if (!h->nd_pending) {
          cudaMemcpyAsync(h->nd_host, scal_i, sizeof(int), cudaMemcpyDeviceToHost, s);
          cudaEventRecord(h->nd_event, s);
          h->nd_pending = true;
        }
      }
      h->aux_tc_step = w.aux_tc && (h->aux_tc_always || h->nd_host == nullptr || h->n_dead_lagged > 32);
      if (c.aux_kind == SAEV_B200_AUX_AUXK) {
        AuxArgs a;
        a.x = x;
        a.resid = resid;
        a.W_enc_t = W_enc_t;
        a.b_enc = b_enc;
        a.W_dec = W_dec;
        a.b_dec = b_dec;
        a.dead_list = at<int>(workspace, w.dead_list);
        a.n_dead = scal_i;
        a.B = B;
        a.D = D;
        a.S = h->aux_cap;
        a.k_aux = c.k_aux;
        a.alpha = c.aux_alpha;
        a.inv_bd = static_cast<float>(1.0 / (static_cast<double>(tokens_global) * D));
        a.remove_parallel = c.remove_parallel_grads;
        a.h_aux = at<float>(workspace, w.h_aux);
        a.mask_aux = at<unsigned char>(workspace, w.mask_aux);
        a.r_aux = at<float>(workspace, w.r_aux);
        a.row_sse_aux = at<float>(workspace, w.row_sse_aux);
        a.aux_loss = aux_loss;
        a.gW_enc_t = a.gb_enc = a.gW_dec = nullptr;
        a.colsum_partial = at<float>(workspace, w.colsum_partial);
        a.gb_dec = nullptr;
        a.aux_colpart = at<float>(workspace, w.aux_colpart);
        a.row_gsq = nullptr;
        fill_aux_tc(h, workspace, a);
        if (launch_aux_forward(a, s)) return fail(h, 46, "forward: AuxK launch failed%s");
        aux_live = true;
      }
    } else {
      cudaMemsetAsync(scal_i, 0, 4, s);  // n_dead = 0 (eval: dead_mask is None, objectives.py:121-122)
    }
    FinalizeArgs f;
    f.row_sse = at<float>(workspace, w.row_sse);
    f.row_l1 = at<float>(workspace, w.row_l1);
    f.row_l0 = at<float>(workspace, w.row_l0);
    f.B = B;
    f.D = D;
    // per-rank partial means compose: each rank divides by the GLOBAL batch
    f.inv_bd = 1.0 / (static_cast<double>(tokens_global) * h->pf_fwd.n * D);  // mean over B * P * D (objectives.py:133-138)
    f.inv_b = 1.0 / static_cast<double>(tokens_global);
    f.l1_coeff = c.l1_coeff;
    f.aux_loss = aux_live ? aux_loss : nullptr;
    f.n_dead = scal_i;
    f.losses = losses;
    if (launch_finalize(f, s)) return fail(h, 47, "forward: finalize launch failed%s");
  }
  return check_cuda(h, "forward");
}


int saev_b200_batch_topk(saev_b200_handle* h, int32_t B, int32_t k_per_sample, int32_t training, float* threshold,
                         float momentum, int32_t* topk_idx, float* topk_val, int32_t* stats, void* workspace,
                         void* stream) {
  bind_context(h);
  const saev_b200_cfg& c = h->cfg;
  const Workspace& w = h->ws;
  if (c.act_kind != SAEV_B200_ACT_TOPK) return fail(h, 48, "batch_topk: the handle must be created with act_kind TOPK%s");
  if (B <= 0 || B > c.max_batch) return fail(h, 40, "batch_topk: B out of range (0 < B <= cfg.max_batch)%s");
  if (k_per_sample <= 0) return fail(h, 48, "batch_topk: k_per_sample must be positive%s");
  if (!training && threshold == nullptr) return fail(h, 48, "batch_topk: eval mode needs the threshold buffer%s");
  StageTimer tm(h, SAEV_B200_STAGE_RESCORE, static_cast<cudaStream_t>(stream));
  if (launch_batch_topk(topk_idx, topk_val, B, c.top_k, c.d_sae, static_cast<long long>(k_per_sample) * B, training,
                        threshold, momentum, at<int>(workspace, w.feat_count), at<int>(workspace, w.active),
                        at<int>(workspace, w.btk), stats, static_cast<cudaStream_t>(stream)))
    return fail(h, 48, "batch_topk: launch failed%s");
  return check_cuda(h, "batch_topk");
}

}  // extern "C" (re-opened below)

namespace {

struct BwdCtx {
  saev_b200_handle* h;
  const float* x; int B; long long tokens_global;
  const float* W_enc_t; const float* b_enc; const float* W_dec; const float* b_dec;
  const int32_t* topk_idx; const float* topk_val; const float* resid;
  float* gW_enc_t; float* gb_enc; float* gW_dec; float* gb_dec;
  void* workspace; cudaStream_t s;
};

int bwd_csc(const BwdCtx& c) {
  saev_b200_handle* h = c.h;
  const Workspace& w = h->ws;
  StageTimer tm(h, SAEV_B200_STAGE_CSC, c.s);
  if (launch_csc_build(c.topk_idx, c.B, h->cfg.top_k, h->cfg.d_sae, at<int>(c.workspace, w.feat_count),
                       at<int>(c.workspace, w.feat_off), at<int>(c.workspace, w.cursor), at<int>(c.workspace, w.entries),
                       at<int>(c.workspace, w.block_totals), c.s, at<int>(c.workspace, w.heavy_list) + 1,
                       at<int>(c.workspace, w.heavy_list)))
    return fail(h, 51, "backward: CSC build launch failed%s");
  return 0;
}

int bwd_wgrad(const BwdCtx& c, int row_begin, int row_end, const long long* skip_toks) {
  saev_b200_handle* h = c.h;
  const Workspace& w = h->ws;
  const int D = h->cfg.d_model;
  WgradArgs g;
  g.feat_off = at<int>(c.workspace, w.feat_off);
  g.entries = at<int>(c.workspace, w.entries);
  g.topk_val = c.topk_val;
  g.dh = h->dh_fused_fwd ? nullptr : at<float>(c.workspace, w.dh);
  g.l1_over_b = h->cfg.l1_coeff != 0.f ? static_cast<float>(h->cfg.l1_coeff / static_cast<double>(c.tokens_global)) : 0.f;
  {
    static const int wpb = [] { const char* v = getenv("SAEV_B200_WGRAD_WPB"); return v ? atoi(v) : 1; }();
    g.warps_per_block = wpb;
    static const bool heavy = [] { const char* v = getenv("SAEV_B200_WGRAD_HEAVY"); return !(v && v[0] == '0'); }();
    g.heavy_list = heavy ? at<int>(c.workspace, w.heavy_list) + 1 : nullptr;
    g.n_heavy = at<int>(c.workspace, w.heavy_list);
    g.heavy_ticket = at<int>(c.workspace, w.cursor);
    // two-pass weight gradients (decoder side + dh, then encoder side): each pass gathers ONE 67 MB row set that the L2
    // can hold, instead of x and the residual together (c3: 0.78 -> 0.41 ms); SAEV_B200_WGRAD_SPLIT=0 = single pass
    static const bool split = [] { const char* v = getenv("SAEV_B200_WGRAD_SPLIT"); return !(v && v[0] == '0'); }();
    g.split = split ? 1 : 0;
    g.dh_scratch = at<float>(c.workspace, w.dh);
  }
  g.resid = c.resid;
  g.x = c.x;
  g.W_dec = c.W_dec;
  g.B = c.B;
  g.D = D;
  g.S = h->cfg.d_sae;
  g.K = h->cfg.top_k;
  g.grad_scale = static_cast<float>(2.0 / (static_cast<double>(c.tokens_global) * h->pf_fwd.n * D));
  g.sfx = h->pf_fwd.n > 1 ? at<float>(c.workspace, w.sfx) : nullptr;
  g.pf = h->pf_fwd;
  g.remove_parallel = h->cfg.remove_parallel_grads;
  g.gW_enc_t = c.gW_enc_t;
  g.gb_enc = c.gb_enc;
  g.gW_dec = c.gW_dec;
  g.row_gsq = at<float>(c.workspace, w.row_gsq);
  g.row_begin = row_begin;
  g.row_end = row_end;
  g.skip_toks = skip_toks;
  g.skip_threshold = h->cfg.dead_threshold_tokens;
  StageTimer tm(h, SAEV_B200_STAGE_WGRAD, c.s);
  if (launch_wgrad(g, c.s)) return fail(h, 52, "backward: weight-gradient launch failed%s");
  return 0;
}

// gb_dec = grad_scale * sum_b resid, then the AuxK gradients (rows of the dead atoms, gb_dec += ...)
int bwd_bias_aux(const BwdCtx& c) {
  saev_b200_handle* h = c.h;
  const saev_b200_cfg& cf = h->cfg;
  const Workspace& w = h->ws;
  const int D = cf.d_model;
  const int P = h->pf_fwd.n;
  const float grad_scale = static_cast<float>(2.0 / (static_cast<double>(c.tokens_global) * P * D));
  StageTimer tm_tail(h, SAEV_B200_STAGE_BIAS_AUX, c.s);
  // gb_dec = sum_b sum_i G_i: with prefixes that is the column sum of the block-0 suffix sums
  if (launch_colsum(P > 1 ? at<float>(c.workspace, w.sfx) : c.resid, c.B, D, grad_scale, 0,
                    at<float>(c.workspace, w.colsum_partial), c.gb_dec, c.s, P > 1 ? static_cast<long long>(P) * D : 0))
    return fail(h, 53, "backward: bias-gradient launch failed%s");
  if (cf.aux_kind == SAEV_B200_AUX_AUXK && h->last_forward_tracked) {
    AuxArgs a;
    a.x = c.x;
    a.resid = c.resid;
    a.W_enc_t = c.W_enc_t;
    a.b_enc = c.b_enc;
    a.W_dec = c.W_dec;
    a.b_dec = c.b_dec;
    a.dead_list = at<int>(c.workspace, w.dead_list);
    a.n_dead = at<int>(c.workspace, w.scalars);
    a.B = c.B;
    a.D = D;
    a.S = h->aux_cap;
    a.k_aux = cf.k_aux;
    a.alpha = cf.aux_alpha;
    a.inv_bd = static_cast<float>(1.0 / (static_cast<double>(c.tokens_global) * D));
    a.remove_parallel = cf.remove_parallel_grads;
    a.h_aux = at<float>(c.workspace, w.h_aux);
    a.mask_aux = at<unsigned char>(c.workspace, w.mask_aux);
    a.r_aux = at<float>(c.workspace, w.r_aux);
    a.row_sse_aux = at<float>(c.workspace, w.row_sse_aux);
    a.aux_loss = at<float>(c.workspace, w.scalars) + SC_AUX_LOSS;
    a.gW_enc_t = c.gW_enc_t;
    a.gb_enc = c.gb_enc;
    a.gW_dec = c.gW_dec;
    a.colsum_partial = at<float>(c.workspace, w.colsum_partial);
    a.gb_dec = c.gb_dec;
    a.aux_colpart = at<float>(c.workspace, w.aux_colpart);
    a.row_gsq = cf.act_kind == SAEV_B200_ACT_TOPK ? at<float>(c.workspace, w.row_gsq) : nullptr;
    fill_aux_tc(h, c.workspace, a);
    if (launch_aux_backward(a, c.s)) return fail(h, 54, "backward: AuxK launch failed%s");
  }
  return 0;
}

}  // namespace

extern "C" {

int saev_b200_backward(saev_b200_handle* h, const float* x, int32_t B, int64_t tokens_global,
                       const float* W_enc_t, const float* b_enc, const float* W_dec, const float* b_dec,
                       const int32_t* topk_idx, const float* topk_val, const float* resid, float* gW_enc_t,
                       float* gb_enc, float* gW_dec, float* gb_dec, void* workspace, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bind_context(h);
  const saev_b200_cfg& c = h->cfg;
  if (!h->last_forward_training) return fail(h, 50, "backward: the last forward was not a training forward%s");
  if (B <= 0 || B > c.max_batch) return fail(h, 40, "backward: B out of range%s");
  if (tokens_global <= 0) tokens_global = B;
  const BwdCtx ctx{h, x, B, tokens_global, W_enc_t, b_enc, W_dec, b_dec, topk_idx, topk_val, resid,
                   gW_enc_t, gb_enc, gW_dec, gb_dec, workspace, s};
  if (c.act_kind == SAEV_B200_ACT_RELU) {
    if (int rc = backward_relu(h, x, B, tokens_global, W_dec, gW_enc_t, gb_enc, gW_dec, workspace, s)) return rc;
  } else {
    if (int rc = bwd_csc(ctx)) return rc;
    if (int rc = bwd_wgrad(ctx, 0, c.d_sae, nullptr)) return rc;
  }
  if (int rc = bwd_bias_aux(ctx)) return rc;
  h->row_gsq_valid = c.act_kind == SAEV_B200_ACT_TOPK;
  return check_cuda(h, "backward");
}

/* Staged form of saev_b200_backward for overlapping the gradient exchange with the weight-gradient kernel (TopK
 * path).  stage 0: per-atom lists + gb_dec + AuxK gradients (rows of the dead atoms); stage 1: the weight-gradient
 * rows [row_begin, row_end) of every atom that is not dead -- so rows [r0, r1) of gW_enc_t / gW_dec are final once
 * the stage-1 call covering them has run, and may be all-reduced while later rows are still being computed. */
int saev_b200_backward_stage(saev_b200_handle* h, int32_t stage, int32_t row_begin, int32_t row_end, const float* x,
                             int32_t B, int64_t tokens_global, const float* W_enc_t, const float* b_enc,
                             const float* W_dec, const float* b_dec, const int64_t* toks_since_active,
                             const int32_t* topk_idx, const float* topk_val, const float* resid, float* gW_enc_t,
                             float* gb_enc, float* gW_dec, float* gb_dec, void* workspace, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bind_context(h);
  bind_context(h);
  const saev_b200_cfg& c = h->cfg;
  if (c.act_kind != SAEV_B200_ACT_TOPK) return fail(h, 55, "backward_stage: TopK path only%s");
  if (!h->last_forward_training) return fail(h, 50, "backward_stage: the last forward was not a training forward%s");
  if (B <= 0 || B > c.max_batch) return fail(h, 40, "backward_stage: B out of range%s");
  if (tokens_global <= 0) tokens_global = B;
  const BwdCtx ctx{h, x, B, tokens_global, W_enc_t, b_enc, W_dec, b_dec, topk_idx, topk_val, resid,
                   gW_enc_t, gb_enc, gW_dec, gb_dec, workspace, s};
  if (stage == 0) {
    if (int rc = bwd_csc(ctx)) return rc;
    if (int rc = bwd_bias_aux(ctx)) return rc;
    h->row_gsq_valid = false;
  } else if (stage == 1) {
    if (row_begin < 0 || row_end > c.d_sae || row_begin > row_end) return fail(h, 55, "backward_stage: bad row range%s");
    const bool aux_live = c.aux_kind == SAEV_B200_AUX_AUXK && h->last_forward_tracked;
    if (aux_live && toks_since_active == nullptr)
      return fail(h, 55, "backward_stage: toks_since_active is needed to leave the AuxK rows alone%s");
    if (int rc = bwd_wgrad(ctx, row_begin, row_end, aux_live ? reinterpret_cast<const long long*>(toks_since_active) : nullptr))
      return rc;
  } else {
    return fail(h, 55, "backward_stage: stage must be 0 or 1%s");
  }
  return check_cuda(h, "backward_stage");
}

int saev_b200_set_reserved_sms(saev_b200_handle* h, int32_t n_sms) {
  if (n_sms < 0 || n_sms > h->num_sms / 2) return fail(h, 66, "set_reserved_sms: 0 <= n_sms <= #SMs / 2%s");
  h->reserved_pairs = (n_sms + 1) / 2;
  return 0;
}

int saev_b200_set_optimizer_shard(saev_b200_handle* h, int32_t row_begin, int32_t row_end) {
  if (row_begin < 0 || row_end > h->cfg.d_sae || (row_end != 0 && row_begin >= row_end))
    return fail(h, 64, "set_optimizer_shard: need 0 <= row_begin < row_end <= d_sae (or 0, 0 to reset)%s");
  h->shard_begin = row_begin;
  h->shard_end = row_end;
  return 0;
}

int saev_b200_grad_sumsq_ranges(saev_b200_handle* h, const float* grads_flat, int32_t n_ranges, const int64_t* begins,
                                const int64_t* ends, float* sumsq_out, void* workspace, void* stream) {
  StageTimer tm(h, SAEV_B200_STAGE_SUMSQ, static_cast<cudaStream_t>(stream));
  long long b[SUMSQ_MAX_RANGES], e[SUMSQ_MAX_RANGES];
  if (n_ranges < 1 || n_ranges > SUMSQ_MAX_RANGES) return fail(h, 65, "grad_sumsq_ranges: 1..18 ranges%s");
  for (int r = 0; r < n_ranges; ++r) {
    if (begins[r] % 4 != 0 || ends[r] < begins[r]) return fail(h, 65, "grad_sumsq_ranges: begins must be multiples of 4%s");
    b[r] = begins[r];
    e[r] = ends[r];
  }
  if (launch_sumsq_ranges(grads_flat, n_ranges, b, e, at<double>(workspace, h->ws.sumsq_partial), sumsq_out,
                          static_cast<cudaStream_t>(stream)))
    return fail(h, 60, "grad_sumsq_ranges: launch failed%s");
  return check_cuda(h, "grad_sumsq_ranges");
}

void* saev_b200_shadow_weights(const saev_b200_handle* h, void* workspace) {
  return at<char>(workspace, h->ws.shadow_hi);
}
float* saev_b200_wnorm_rows(const saev_b200_handle* h, void* workspace) {
  return at<float>(workspace, h->ws.wnorm_rows);
}
float* saev_b200_wnorm_scalar(const saev_b200_handle* h, void* workspace) {
  return at<float>(workspace, h->ws.scalars) + SC_WNORM_SQ_MAX;
}


int saev_b200_grad_sumsq_local(saev_b200_handle* h, const float* gb_dec, float* sumsq_out, void* workspace,
                               void* stream) {
  if (!h->row_gsq_valid) return fail(h, 63, "grad_sumsq_local: no per-atom partials (call right after backward; TopK only)%s");
  StageTimer tm(h, SAEV_B200_STAGE_SUMSQ, static_cast<cudaStream_t>(stream));
  if (launch_sumsq_fused(at<float>(workspace, h->ws.row_gsq), h->cfg.d_sae, gb_dec, h->cfg.d_model, sumsq_out,
                         static_cast<cudaStream_t>(stream)))
    return fail(h, 60, "grad_sumsq_local: launch failed%s");
  return check_cuda(h, "grad_sumsq_local");
}

int saev_b200_grad_sumsq(saev_b200_handle* h, const float* grads_flat, int64_t n, float* sumsq_out,
                         void* workspace, void* stream) {
  StageTimer tm(h, SAEV_B200_STAGE_SUMSQ, static_cast<cudaStream_t>(stream));
  if (launch_sumsq(grads_flat, n, at<double>(workspace, h->ws.sumsq_partial), sumsq_out,
                   static_cast<cudaStream_t>(stream)))
    return fail(h, 60, "grad_sumsq: launch failed%s");
  return check_cuda(h, "grad_sumsq");
}

int saev_b200_adam_step(saev_b200_handle* h, float* W_enc_t, float* b_enc, float* W_dec, float* b_dec,
                        const float* grads_flat, float* m_flat, float* v_flat, float lr, float beta1,
                        float beta2, float eps, int64_t step, float max_norm, float grad_scale,
                        const float* sumsq, int32_t renorm_w_dec, float* gnorm_out, int32_t parts, void* workspace,
                        void* stream) {
  const long long S = h->cfg.d_sae, D = h->cfg.d_model;
  if (step < 1) return fail(h, 61, "adam_step: step must be >= 1%s");
  if ((parts & 3) == 0 || parts > 15) return fail(h, 61, "adam_step: parts must select the encoder (1) and / or decoder (2) half%s");
  AdamArgs a;
  a.W_enc_t = W_enc_t;
  a.b_enc = b_enc;
  a.W_dec = W_dec;
  a.b_dec = b_dec;
  a.gW_enc_t = grads_flat;
  a.gb_enc = grads_flat + S * D;
  a.gW_dec = grads_flat + S * D + S;
  a.gb_dec = grads_flat + 2 * S * D + S;
  a.m = m_flat;
  a.v = v_flat;
  const bool screen = workspace != nullptr && h->cfg.act_kind == SAEV_B200_ACT_TOPK;
  a.shadow16 = screen ? at<__half>(workspace, h->ws.shadow_hi) : nullptr;
  a.wnorm_sq_max = screen ? at<float>(workspace, h->ws.scalars) + SC_WNORM_SQ_MAX : nullptr;
  a.bias_abs_max = screen ? at<float>(workspace, h->ws.scalars) + SC_BIAS_ABS_MAX : nullptr;
  a.wnorm_rows = screen ? at<float>(workspace, h->ws.wnorm_rows) : nullptr;
  a.rho = screen ? at<float>(workspace, h->ws.scalars) + SC_RHO : nullptr;
  a.D = static_cast<int>(D);
  a.S = static_cast<int>(S);
  a.lr = lr;
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = eps;
  // bias corrections in double on the host (torch computes 1 - beta^step the same way, train.py:294)
  double b1p = 1.0, b2p = 1.0;
  {
    double x1 = beta1, x2 = beta2;
    long long e = step;
    while (e > 0) {
      if (e & 1) {
        b1p *= x1;
        b2p *= x2;
      }
      x1 *= x1;
      x2 *= x2;
      e >>= 1;
    }
  }
  a.bc1 = static_cast<float>(1.0 - b1p);
  a.bc2_sqrt = static_cast<float>(sqrt(1.0 - b2p));
  a.max_norm = max_norm;
  a.grad_scale = grad_scale;
  a.gnorm_sq = sumsq;
  a.renorm_w_dec = renorm_w_dec;
  a.gnorm_out = gnorm_out;
  const bool sharded = h->shard_end > 0;
  a.row_begin = sharded ? h->shard_begin : 0;
  a.row_end = sharded ? h->shard_end : static_cast<int>(S);
  a.b_enc_separately = sharded ? 1 : 0;
  a.parts = parts;
  a.small_blocks = (parts & 3) == 2 ? 1 : 0;  // a decoder-only update is meant to run beside the next step's screen kernel
  StageTimer tm(h, SAEV_B200_STAGE_ADAM, static_cast<cudaStream_t>(stream));
  if (launch_adam(a, static_cast<cudaStream_t>(stream))) return fail(h, 62, "adam_step: launch failed%s");
  return check_cuda(h, "adam_step");
}

int saev_b200_densify(saev_b200_handle* h, const int32_t* topk_idx, const float* topk_val, int32_t B,
                      float* f_x_out, void* stream) {
  if (launch_densify(topk_idx, topk_val, B, h->cfg.top_k, h->cfg.d_sae, f_x_out, static_cast<cudaStream_t>(stream)))
    return fail(h, 70, "densify: launch failed%s");
  return check_cuda(h, "densify");
}

int saev_b200_dense_f(saev_b200_handle* h, const int32_t* topk_idx, const float* topk_val, int32_t B, float* f_x_out,
                      void* workspace, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->cfg.act_kind == SAEV_B200_ACT_RELU) {
    if (!workspace) return fail(h, 70, "dense_f: the ReLU path needs the workspace%s");
    if (launch_join_bf16(at<__nv_bfloat16>(workspace, h->ws.f_hi), at<__nv_bfloat16>(workspace, h->ws.f_lo),
                         static_cast<long long>(B) * h->cfg.d_sae, f_x_out, s,
                         h->dense_terms == 6 ? at<__nv_bfloat16>(workspace, h->ws.f_l2) : nullptr))
      return fail(h, 70, "dense_f: launch failed%s");
    return check_cuda(h, "dense_f");
  }
  return saev_b200_densify(h, topk_idx, topk_val, B, f_x_out, stream);
}

int saev_b200_set_prefixes(saev_b200_handle* h, const int32_t* host_prefixes, int32_t n) {
  if (n < 1 || n > h->max_prefixes) return fail(h, 72, "set_prefixes: 1 <= n <= cfg.max_prefixes%s");
  if (host_prefixes == nullptr) {
    if (n != 1) return fail(h, 72, "set_prefixes: null prefixes%s");
    h->pf.n = 1;
    h->pf.cut[0] = h->cfg.d_sae;
    return 0;
  }
  for (int i = 0; i < n; ++i) {
    if (host_prefixes[i] < 1 || (i > 0 && host_prefixes[i] <= host_prefixes[i - 1]))
      return fail(h, 72, "set_prefixes: cuts must be strictly increasing and >= 1 (modeling.py:372-373)%s");
    h->pf.cut[i] = host_prefixes[i];
  }
  if (host_prefixes[n - 1] != h->cfg.d_sae) return fail(h, 72, "set_prefixes: the last cut must be d_sae%s");
  h->pf.n = n;
  return 0;
}

int saev_b200_x_hats(saev_b200_handle* h, const float* resid, const float* x, int32_t B, float* x_hats_out,
                     void* workspace, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int P = h->pf_fwd.n;
  if (P <= 1) return saev_b200_x_hat(h, resid, x, B, x_hats_out, stream);
  if (!workspace) return fail(h, 71, "x_hats: workspace needed%s");
  if (launch_x_hats_prefix(at<float>(workspace, h->ws.sfx), x, B, h->cfg.d_model, P, x_hats_out, s))
    return fail(h, 71, "x_hats: launch failed%s");
  return check_cuda(h, "x_hats");
}

int saev_b200_x_hat(saev_b200_handle* h, const float* resid, const float* x, int32_t B, float* x_hat_out,
                    void* stream) {
  if (launch_add_rows(resid, x, static_cast<long long>(B) * h->cfg.d_model, x_hat_out,
                      static_cast<cudaStream_t>(stream)))
    return fail(h, 71, "x_hat: launch failed%s");
  return check_cuda(h, "x_hat");
}

size_t saev_b200_coherence_scratch_bytes(const saev_b200_handle* h) {
  if (!h) return 0;
  const size_t S = static_cast<size_t>(h->cfg.d_sae), D = static_cast<size_t>(h->cfg.d_model);
  return 2 * align_up(S * D * 2) + 2 * align_up(S * ENCODE_MAX_NSPLIT * 4);
}

int saev_b200_dictionary_coherence(saev_b200_handle* h, const float* W_dec, void* scratch, size_t scratch_bytes,
                                   float* out, void* stream) {
  if (!h || !W_dec || !scratch || !out) return fail(h, 85, "dictionary_coherence: null argument%s");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bind_context(h);
  const int S = h->cfg.d_sae, D = h->cfg.d_model;
  if (D % 8) return fail(h, 85, "dictionary_coherence: d_model must be a multiple of 8%s");
  if (scratch_bytes < saev_b200_coherence_scratch_bytes(h)) return fail(h, 85, "dictionary_coherence: scratch too small%s");
  const size_t piece = align_up(static_cast<size_t>(S) * D * 2), col = align_up(static_cast<size_t>(S) * ENCODE_MAX_NSPLIT * 4);
  uint8_t* base = static_cast<uint8_t*>(scratch);
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(base + piece);
  float* row_best = reinterpret_cast<float*>(base + 2 * piece);
  int* row_col = reinterpret_cast<int*>(base + 2 * piece + col);
  if (launch_unit_rows_split(W_dec, S, D, hi, lo, s)) return fail(h, 86, "dictionary_coherence: split launch failed%s");
  EncodeGemmArgs g;
  g.A_hi = hi;
  g.A_lo = lo;
  g.B_hi = hi;
  g.B_lo = lo;
  g.nterms = 3;
  g.M = S;
  g.N = S;
  g.K = D;
  g.epilogue = 5;
  g.num_sms = h->num_sms;
  g.nsplit = encode_gemm_nsplit(S, S, h->num_sms);
  g.extra = row_best;
  g.active = row_col;
  if (int rc = launch_encode_gemm(g, s)) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d", rc);
    return fail(h, 86, "dictionary_coherence: screen launch failed (code %s)", buf);
  }
  if (launch_coherence_finish(W_dec, D, row_best, row_col, S * g.nsplit, g.nsplit, 1e-3f, out, s))
    return fail(h, 86, "dictionary_coherence: finish launch failed%s");
  return check_cuda(h, "dictionary_coherence");
}

size_t saev_b200_log_scratch_bytes(const saev_b200_handle* h) {
  if (!h) return 0;
  return saev_b200_coherence_scratch_bytes(h) + align_up((8 + static_cast<size_t>(h->cfg.d_model)) * 8) + align_up(16);
}

int saev_b200_log_metrics(saev_b200_handle* h, const float* x, const float* resid, int32_t B, const float* W_dec,
                          void* workspace, void* scratch, size_t scratch_bytes, double* out, void* stream) {
  if (!h || !x || !resid || !W_dec || !workspace || !scratch || !out) return fail(h, 87, "log_metrics: null argument%s");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, 87, "log_metrics: bad batch size%s");
  if (scratch_bytes < saev_b200_log_scratch_bytes(h)) return fail(h, 87, "log_metrics: scratch too small%s");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t coh_bytes = saev_b200_coherence_scratch_bytes(h);
  uint8_t* base = static_cast<uint8_t*>(scratch);
  double* acc = reinterpret_cast<double*>(base + coh_bytes);
  float* coh = reinterpret_cast<float*>(base + coh_bytes + align_up((8 + static_cast<size_t>(h->cfg.d_model)) * 8));
  if (int rc = saev_b200_dictionary_coherence(h, W_dec, scratch, coh_bytes, coh, stream)) return rc;
  if (launch_log_metrics(x, resid, B, h->cfg.d_model, W_dec, h->cfg.d_sae, saev_b200_active_flags(h, workspace), acc, coh,
                         out, s))
    return fail(h, 88, "log_metrics: launch failed%s");
  return check_cuda(h, "log_metrics");
}

int saev_b200_eval_accumulate(saev_b200_handle* h, const float* x, const float* resid, int32_t B,
                              const int32_t* topk_idx, const float* topk_val, const float* losses, double* acc,
                              float* n_fired, float* values, void* workspace, void* stream) {
  if (!h || !x || !resid || !losses || !acc || !n_fired || !values) return fail(h, 89, "eval_accumulate: null argument%s");
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, 89, "eval_accumulate: bad batch size%s");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bind_context(h);
  const int S = h->cfg.d_sae, D = h->cfg.d_model;
  if (launch_eval_accumulate(x, resid, B, D, losses, acc, s)) return fail(h, 89, "eval_accumulate: launch failed%s");
  if (h->cfg.act_kind == SAEV_B200_ACT_RELU) {
    if (!workspace) return fail(h, 89, "eval_accumulate: the ReLU path needs the workspace%s");
    if (launch_feature_stats_dense(at<__nv_bfloat16>(workspace, h->ws.f_hi), at<__nv_bfloat16>(workspace, h->ws.f_lo),
                                   h->dense_terms == 6 ? at<__nv_bfloat16>(workspace, h->ws.f_l2) : nullptr, B, S, S,
                                   n_fired, values, s))
      return fail(h, 89, "eval_accumulate: launch failed%s");
  } else {
    if (!topk_idx || !topk_val) return fail(h, 89, "eval_accumulate: null top-k buffers%s");
    if (launch_feature_stats_topk(topk_idx, topk_val, static_cast<long long>(B) * h->cfg.top_k, n_fired, values, s))
      return fail(h, 89, "eval_accumulate: launch failed%s");
  }
  return check_cuda(h, "eval_accumulate");
}

int saev_b200_aux_selection(const saev_b200_handle* h, void* workspace, const uint8_t** mask, int64_t* ld,
                            const int32_t** dead_list, const int32_t** n_dead) {
  if (!h || !workspace || !mask || !ld || !dead_list || !n_dead) return fail(h, 83, "aux_selection: null argument%s");
  if (h->cfg.aux_kind != SAEV_B200_AUX_AUXK) return fail(h, 83, "aux_selection: the handle has no AuxK%s");
  *mask = at<uint8_t>(workspace, h->ws.mask_aux);
  *ld = h->aux_cap;
  *dead_list = at<int32_t>(workspace, h->ws.dead_list);
  *n_dead = at<int32_t>(workspace, h->ws.scalars) + SC_N_DEAD;
  return 0;
}

int saev_b200_gemm_nt(saev_b200_handle* h, const float* A, const float* Bt, const float* bias, int32_t M,
                      int32_t N, int32_t K, int32_t nterms, float* out, void* scratch, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bind_context(h);
  if (K % 8) return fail(h, 80, "gemm_nt: K must be a multiple of 8%s");
  if (nterms != 1 && nterms != 3 && nterms != 6) return fail(h, 80, "gemm_nt: nterms must be 1, 3 or 6%s");
  __nv_bfloat16* a_hi = static_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* a_lo = a_hi + static_cast<size_t>(M) * K;
  __nv_bfloat16* a_l2 = a_lo + static_cast<size_t>(M) * K;
  __nv_bfloat16* b_hi = a_l2 + static_cast<size_t>(M) * K;
  __nv_bfloat16* b_lo = b_hi + static_cast<size_t>(N) * K;
  __nv_bfloat16* b_l2 = b_lo + static_cast<size_t>(N) * K;
  if (launch_split_bf16(A, a_hi, a_lo, static_cast<long long>(M) * K, s, nterms == 6 ? a_l2 : nullptr) ||
      launch_split_bf16(Bt, b_hi, b_lo, static_cast<long long>(N) * K, s, nterms == 6 ? b_l2 : nullptr))
    return fail(h, 81, "gemm_nt: split_bf16 launch failed%s");
  EncodeGemmArgs g;
  g.A_hi = a_hi;
  g.A_lo = a_lo;
  g.A_lo2 = a_l2;
  g.B_hi = b_hi;
  g.B_lo = b_lo;
  g.B_lo2 = b_l2;
  g.nterms = nterms;
  g.k_chunk_blocks = nterms == 6 ? 8 : 0;
  g.M = M;
  g.N = N;
  g.K = K;
  g.bias = bias;
  g.epilogue = 1;
  g.nsplit = 0;
  g.num_sms = h->num_sms;
  g.out = out;
  g.ldo = N;
  if (int rc = launch_encode_gemm(g, s)) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d", rc);
    return fail(h, 82, "gemm_nt: launch failed (code %s)", buf);
  }
  return check_cuda(h, "gemm_nt");
}

// ---------------------------------------------------------------------------------------------
// pinned staging ring
// ---------------------------------------------------------------------------------------------
struct saev_b200_ring {
  int n_slots = 0;
  size_t slot_bytes = 0;
  void** host = nullptr;
  cudaEvent_t* done = nullptr;
  cudaStream_t copy_stream = nullptr;
};

int saev_b200_ring_create(int32_t n_slots, size_t slot_bytes, saev_b200_ring** out) {
  if (!out || n_slots <= 0 || slot_bytes == 0) return fail(nullptr, 90, "ring_create: bad arguments%s");
  saev_b200_ring* r = new (std::nothrow) saev_b200_ring();
  if (!r) return fail(nullptr, 91, "ring_create: out of host memory%s");
  r->n_slots = n_slots;
  r->slot_bytes = slot_bytes;
  r->host = new void*[n_slots]();
  r->done = new cudaEvent_t[n_slots]();
  if (cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    saev_b200_ring_destroy(r);
    return fail(nullptr, 92, "ring_create: cudaStreamCreate failed%s");
  }
  for (int i = 0; i < n_slots; ++i) {
    if (cudaHostAlloc(&r->host[i], slot_bytes, cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->done[i], cudaEventDisableTiming) != cudaSuccess) {
      saev_b200_ring_destroy(r);
      return fail(nullptr, 93, "ring_create: pinned allocation failed%s");
    }
  }
  *out = r;
  return 0;
}

int saev_b200_ring_destroy(saev_b200_ring* r) {
  if (!r) return 0;
  if (r->copy_stream) cudaStreamSynchronize(r->copy_stream);
  for (int i = 0; i < r->n_slots; ++i) {
    if (r->host && r->host[i]) cudaFreeHost(r->host[i]);
    if (r->done && r->done[i]) cudaEventDestroy(r->done[i]);
  }
  if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
  delete[] r->host;
  delete[] r->done;
  delete r;
  return 0;
}

void* saev_b200_ring_host_ptr(saev_b200_ring* r, int32_t slot) {
  return (r && slot >= 0 && slot < r->n_slots) ? r->host[slot] : nullptr;
}

int saev_b200_ring_submit(saev_b200_ring* r, int32_t slot, void* dst_device, size_t bytes) {
  if (!r || slot < 0 || slot >= r->n_slots || bytes > r->slot_bytes)
    return fail(nullptr, 94, "ring_submit: bad arguments%s");
  if (cudaMemcpyAsync(dst_device, r->host[slot], bytes, cudaMemcpyHostToDevice, r->copy_stream) != cudaSuccess ||
      cudaEventRecord(r->done[slot], r->copy_stream) != cudaSuccess)
    return fail(nullptr, 95, "ring_submit: cudaMemcpyAsync failed%s");
  return 0;
}

int saev_b200_ring_wait(saev_b200_ring* r, int32_t slot, void* consumer_stream) {
  if (!r || slot < 0 || slot >= r->n_slots) return fail(nullptr, 94, "ring_wait: bad arguments%s");
  if (cudaStreamWaitEvent(static_cast<cudaStream_t>(consumer_stream), r->done[slot], 0) != cudaSuccess)
    return fail(nullptr, 96, "ring_wait: cudaStreamWaitEvent failed%s");
  return 0;
}

int saev_b200_ring_host_sync(saev_b200_ring* r, int32_t slot) {
  if (!r || slot < 0 || slot >= r->n_slots) return fail(nullptr, 94, "ring_host_sync: bad arguments%s");
  if (cudaEventSynchronize(r->done[slot]) != cudaSuccess)
    return fail(nullptr, 97, "ring_host_sync: cudaEventSynchronize failed%s");
  return 0;
}

}  // extern "C"
