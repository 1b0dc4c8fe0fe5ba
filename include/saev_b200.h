/* saev_b200 — C ABI of the B200-native SAE training step (libsaev_b200.so).
 *
 * Drop-in boundary for the hot path of OSU-NLP-Group/saev (citations are file:line under the
 * reference checkout):
 *   src/saev/nn/modeling.py:343-349    SparseAutoencoder.encode          -> saev_b200_forward (phase A)
 *   src/saev/nn/modeling.py:169-179    TopKActivation.forward            -> saev_b200_forward (phase A)
 *   src/saev/nn/modeling.py:150-156    ReluActivation.forward (dense)    -> saev_b200_forward (phase A, act_kind RELU)
 *   src/saev/nn/modeling.py:183-244    BatchTopKActivation.forward       -> saev_b200_batch_topk (between A_RESCORE and A_DECODE)
 *   src/saev/nn/modeling.py:351-409    SparseAutoencoder.decode          -> saev_b200_forward (phase A)
 *   src/saev/nn/objectives.py:101-156  MatryoshkaObjective.forward       -> saev_b200_forward (A + B)
 *   src/saev/nn/objectives.py:107-122  dead-latent tracker               -> saev_b200_forward (phase B)
 *   src/saev/nn/modeling.py:75-103     AuxK.loss                         -> saev_b200_forward (phase B)
 *   src/saev/framework/train.py:348    loss.backward() (autograd)        -> saev_b200_backward
 *   src/saev/nn/modeling.py:419-445    remove_parallel_grads             -> saev_b200_backward (fused)
 *   src/saev/framework/train.py:358    clip_grad_norm_                   -> saev_b200_grad_sumsq + saev_b200_adam_step
 *   src/saev/framework/train.py:294,444-446  torch.optim.Adam(fused=True).step -> saev_b200_adam_step
 *   src/saev/nn/modeling.py:411-417    normalize_w_dec                   -> saev_b200_normalize_w_dec / adam_step(renorm)
 *   src/saev/framework/train.py:333    batch["act"].to(device)           -> saev_b200_ring_* (pinned staging ring)
 *   src/saev/data/shuffled.py:133-363,380-699 ShuffledDataLoader (manager, I/O workers, __iter__)
 *   src/saev/data/buffers.py:91-231    ReservoirBuffer (shuffle pool)    -> saev_b200_loader_* (pool in HBM)
 *
 * Conventions
 *   - Plain C types only.  Every `float*` / `int*` below is a DEVICE pointer owned by the caller
 *     (PyTorch tensors in practice) unless the name says `host_`.  The library never frees or
 *     resizes caller memory and allocates no device memory of its own: scratch comes from the
 *     caller-provided workspace blob (saev_b200_workspace_bytes).
 *   - All tensors are fp32, row-major, contiguous.  The encoder weight is passed ATOM-MAJOR:
 *         W_enc_t[d_sae, d_model]  ==  saev's W_enc[d_model, d_sae] transposed
 *     (the Python module exposes saev's [d_model, d_sae] parameter as a transposed view of it).
 *   - Gradients and Adam moments use one flat order: [W_enc_t (S*D), b_enc (S), W_dec (S*D), b_dec (D)].
 *   - Every entry enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising the host.  Return value 0 = ok, otherwise an error code; the message is
 *     available from saev_b200_last_error().  One handle per (process, device); calls on one handle
 *     must not race.
 */
#ifndef SAEV_B200_H_
#define SAEV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAEV_B200_ABI_VERSION 11

enum { SAEV_B200_ACT_TOPK = 0, SAEV_B200_ACT_RELU = 1 };
enum { SAEV_B200_AUX_NONE = 0, SAEV_B200_AUX_AUXK = 1 };
enum {
  SAEV_B200_PHASE_A = 1,        /* all of phase A */
  SAEV_B200_PHASE_B = 2,
  SAEV_B200_PHASE_ALL = 3,
  SAEV_B200_PHASE_A_SCREEN = 4, /* first half of A: operand prep + tcgen05 screen (reads only the fp16 operand copy) */
  SAEV_B200_PHASE_A_REST = 8,   /* second half of A: exact re-score + decode (reads the fp32 W_enc_t / W_dec) */
  SAEV_B200_PHASE_A_RESCORE = 16, /* ... or in two calls: the re-score alone (the activity flags are final after it: a
                                     data-parallel caller starts their all-reduce here, beside the decode) */
  SAEV_B200_PHASE_A_DECODE = 32   /* ... and the decode */
};

typedef struct saev_b200_cfg {
  int32_t d_model;               /* SparseAutoencoderConfig.d_model      modeling.py:265 */
  int32_t d_sae;                 /* SparseAutoencoderConfig.d_sae        modeling.py:267 */
  int32_t act_kind;              /* TopK | Relu                          modeling.py:111-126 */
  int32_t top_k;                 /* TopK.top_k                           modeling.py:123 */
  int32_t aux_kind;              /* NoAux | AuxK                         modeling.py:50-106 */
  int32_t k_aux;                 /* AuxK.k_aux                           modeling.py:72 */
  float aux_alpha;               /* AuxK.alpha                           modeling.py:73 */
  float l1_coeff;                /* L1Sparsity.coeff, 0 = NoSparsity     modeling.py:25-42 */
  int64_t dead_threshold_tokens; /* Matryoshka.dead_threshold_tokens     objectives.py:24 */
  int32_t remove_parallel_grads; /* SparseAutoencoderConfig              modeling.py:281 */
  int32_t max_batch;             /* largest B any call will pass */
  int32_t aux_cols_cap;          /* max dead latents AuxK scratch is sized for; 0 = d_sae */
  int32_t max_prefixes;          /* Matryoshka.n_prefixes the workspace is sized for (0/1 = single prefix)  objectives.py:22 */
} saev_b200_cfg;

typedef struct saev_b200_handle saev_b200_handle;

int saev_b200_abi_version(void);
/* Number of CUDA kernels this library has launched in this process (all handles). */
uint64_t saev_b200_launch_count(void);
const char* saev_b200_last_error(const saev_b200_handle* h);

int saev_b200_create(const saev_b200_cfg* cfg, saev_b200_handle** out);
int saev_b200_destroy(saev_b200_handle* h);

/* Bytes of scratch the caller must provide (256-byte aligned) for batches up to cfg.max_batch. */
size_t saev_b200_workspace_bytes(const saev_b200_handle* h);

/* (Re)build what the top-k screen keeps of the encoder in the workspace: the fp16 operand copy of W_enc_t, the largest
 * row norm and the largest |b_enc| (inputs of its error bound), and reset the screen counters.  Call after the
 * weights were written by anything other than saev_b200_adam_step (init, load_state_dict, datapoint init). */
int saev_b200_sync_weights(saev_b200_handle* h, const float* W_enc_t, const float* b_enc, void* workspace,
                           void* stream);

/* Datapoint initialisation of the dictionary on the device (src/saev/framework/train.py:141-185; saev runs it on the
 * host, from batches it pulled through the loader).  acts[n_rows, d_model]: the sample rows (already shuffled as
 * train.py:161 does); src_row[d_sae] (int64, device): for atom j the row of `acts` it is seeded from (train.py:169
 * `idx`); noise[*, d_model]: the kaiming rows (train.py:165-166), atom j uses row noise_row[j] (int64, device; NULL: j).
 *   enc_j = blend * (acts[src_row[j]] - mean(acts)) + (1 - blend) * noise[noise_row[j]]
 *   tie_transpose (cfg.reinit_enc_dec_tranpose): W_dec[j] = enc_j;  normalize (cfg.normalize_w_dec): W_dec[j] /= ||.||
 *   W_enc[:, j] = W_dec[j]                                              (train.py:181)
 * mean_out[d_model] receives the column mean.  Ends with saev_b200_sync_weights. */
int saev_b200_datapoint_init(saev_b200_handle* h, const float* acts, int64_t n_rows, const int64_t* src_row,
                             const float* noise, const int64_t* noise_row, float blend, int32_t tie_transpose,
                             int32_t normalize,
                             float* mean_out, float* W_enc_t, const float* b_enc, float* W_dec, void* workspace,
                             void* stream);

/* W_dec[j,:] /= ||W_dec[j,:]||_2    (modeling.py:411-417) */
int saev_b200_normalize_w_dec(saev_b200_handle* h, float* W_dec, void* stream);

/* Objective forward on one batch x[B, d_model]  (objectives.py:101-156, Matryoshka n_prefixes = 1).
 *   phase A: encode + TopK + decode + residual + per-row loss partials (+ d loss/d h when training);
 *            marks the atoms that fired in `active` (workspace).
 *   phase B: dead tracker update on toks_since_active[S] (int64; pass NULL in eval), AuxK forward,
 *            loss scalars.  Data-parallel callers all-reduce saev_b200_active_flags() (MAX) between
 *            A and B and pass the global batch size as `tokens_global`.
 * Outputs: topk_idx[B,K] (int32, -1 = empty slot), topk_val[B,K], resid[B,D] = x_hat - x,
 *          losses[8] = {mse, aux, sparsity, l0, l1, n_dead, loss, 0} (device).
 * `training` != 0 additionally prepares what saev_b200_backward consumes (kept in the workspace). */
int saev_b200_forward(saev_b200_handle* h, int phase, const float* x, int32_t B, int64_t tokens_global,
                      const float* W_enc_t, const float* b_enc, const float* W_dec, const float* b_dec,
                      int64_t* toks_since_active, int32_t training, int32_t* topk_idx, float* topk_val,
                      float* resid, float* losses, void* workspace, void* stream);

/* BatchTopKActivation.forward (modeling.py:214-244) on the sparse forward state.  The handle is created with act_kind
 * TOPK and cfg.top_k = the per-row CAPACITY `cap` (<= 128; not BatchTopK.top_k).  Call it between
 * saev_b200_forward(phase A_SCREEN | A_RESCORE) -- which leaves each row's `cap` largest exact pre-activations in
 * topk_idx / topk_val -- and saev_b200_forward(phase A_DECODE | B):
 *   training != 0 (:226-242): keeps the min(k_per_sample * B, all) largest entries of the WHOLE batch (`torch.topk` on the
 *     flattened matrix; entries tied at the cut value are granted in (row, rank) order), and folds the smallest positive
 *     survivor into the EMA buffer: *threshold = (1 - momentum) * *threshold + momentum * min_pos   (:237-242).
 *   training == 0 (:220-224): JumpReLU, keeps value > max(*threshold, 0).
 * Losing slots become empty (idx -1, value 0); the per-atom counts and activity flags the backward / tracker use are
 * rebuilt from the survivors.  stats (device int32[4], optional): [0] entries kept, [1] rows that kept all `cap` slots
 * while d_sae > cap -- the reference may have kept more of such a row, so [1] != 0 means the result is NOT certified
 * equal to the reference's (raise cap, or treat as an error), [2] entries tied at the cut, [3] key of the cut value.
 * The global selection does not shard: single rank only. */
int saev_b200_batch_topk(saev_b200_handle* h, int32_t B, int32_t k_per_sample, int32_t training, float* threshold,
                         float momentum, int32_t* topk_idx, float* topk_val, int32_t* stats, void* workspace,
                         void* stream);

/* int32[d_sae] activity flags written by phase A (device pointer inside the workspace). */
int32_t* saev_b200_active_flags(const saev_b200_handle* h, void* workspace);
/* Screen diagnostics, cumulative since the last saev_b200_sync_weights (device pointer to uint32 counters):
 *   [0] rows the tensor-core screen could not certify (candidate list overflow, observed error above the bound);
 *       every one of them was re-done by the exact fp32 path inside the same forward      [2] candidates re-scored
 *   [7] rows re-done by the exact path ( == [0] once the forward has run)      [9] candidate-list entries merged
 *   [8] of [0]: rows whose OBSERVED screen error exceeded the deterministic bound (stays 0 unless the error model of
 *       the tensor-core accumulation is violated)
 *   [11] of [0]: rows whose threshold guess (the screen's warm start, predicted from the previous forward) was too high;
 *       a few per 10^4 rows by construction.  The other expected cause of [0] is a candidate-list overflow. */
uint32_t* saev_b200_unsafe_rows(const saev_b200_handle* h, void* workspace);

/* Gradients of loss = mse + sparsity + aux for the batch of the last training forward
 * (what loss.backward() + remove_parallel_grads() leave in .grad; train.py:348,352).
 * Overwrites gW_enc_t[S,D], gb_enc[S], gW_dec[S,D], gb_dec[D]. */
int saev_b200_backward(saev_b200_handle* h, const float* x, int32_t B, int64_t tokens_global,
                       const float* W_enc_t, const float* b_enc, const float* W_dec, const float* b_dec,
                       const int32_t* topk_idx, const float* topk_val, const float* resid, float* gW_enc_t,
                       float* gb_enc, float* gW_dec, float* gb_dec, void* workspace, void* stream);

/* Staged form of saev_b200_backward (TopK path), for overlapping the data-parallel gradient exchange with the
 * weight-gradient kernel.  stage 0: per-atom lists, gb_dec and the AuxK gradients (the rows of the dead atoms);
 * stage 1: the weight-gradient rows [row_begin, row_end) of every atom that is not dead.  Rows [r0, r1) of
 * gW_enc_t / gW_dec are final once stage 0 and the stage-1 call covering them have been enqueued; gb_enc / gb_dec are
 * final after the last stage-1 call.  `toks_since_active` is the tracker state the forward updated. */
int saev_b200_backward_stage(saev_b200_handle* h, int32_t stage, int32_t row_begin, int32_t row_end, const float* x,
                             int32_t B, int64_t tokens_global, const float* W_enc_t, const float* b_enc,
                             const float* W_dec, const float* b_dec, const int64_t* toks_since_active,
                             const int32_t* topk_idx, const float* topk_val, const float* resid, float* gW_enc_t,
                             float* gb_enc, float* gW_dec, float* gb_dec, void* workspace, void* stream);

/* sumsq_out[0] = sum of squares over the flat gradient bucket of `n` floats (after any all-reduce). */
int saev_b200_grad_sumsq(saev_b200_handle* h, const float* grads_flat, int64_t n, float* sumsq_out,
                         void* workspace, void* stream);

/* Same value for the gradient saev_b200_backward has just written (TopK path, single rank: nothing may have
 * modified the bucket in between), from the per-atom partials the backward kernels left in the workspace -- saves
 * the extra pass over the bucket.  gb_dec = the b_dec slice of the bucket. */
int saev_b200_grad_sumsq_local(saev_b200_handle* h, const float* gb_dec, float* sumsq_out, void* workspace,
                               void* stream);

/* ---- sharded optimizer for data parallelism (no counterpart in saev, which is single-GPU) ----
 * After saev_b200_set_optimizer_shard(h, r0, r1) every saev_b200_adam_step updates only dictionary rows [r0, r1) of
 * W_enc_t / W_dec (and their Adam moments, fp16 operand rows and row-norm maximum), plus both bias vectors in full.
 * The caller reduce-scatters the two weight-gradient regions, all-reduces the bias gradients, takes the norm with
 * saev_b200_grad_sumsq_ranges over what it owns (+ an all-reduce of that scalar), and all-gathers the updated rows,
 * the fp16 operand (saev_b200_shadow_weights, [d_sae, d_model] fp16), the row norms (saev_b200_wnorm_rows) and the
 * row-norm maximum (MAX). */
int saev_b200_set_optimizer_shard(saev_b200_handle* h, int32_t row_begin, int32_t row_end);
/* Leave `n_sms` SMs (rounded up to pairs) out of the top-k screen's persistent grid, so that NCCL kernels (the
 * all-gather of the fp32 rows a sharded optimizer step leaves behind) can run BESIDE the screen of the next step;
 * the split phase A (SCREEN, then REST after the gathers have landed) is the other half of that overlap. */
int saev_b200_set_reserved_sms(saev_b200_handle* h, int32_t n_sms);
int saev_b200_grad_sumsq_ranges(saev_b200_handle* h, const float* grads_flat, int32_t n_ranges,
                                const int64_t* host_begins, const int64_t* host_ends, float* sumsq_out,
                                void* workspace, void* stream);
void* saev_b200_shadow_weights(const saev_b200_handle* h, void* workspace); /* fp16 [d_sae, d_model] */
/* float[3], the dictionary-wide inputs of the screen's error bound: max_j ||W_enc_t[j]||^2, max_j |b_enc[j]|,
 * max_j ||w_j - fp16(w_j)|| / ||w_j||.  A sharded optimizer step computes them over its rows: MAX-all-reduce. */
float* saev_b200_wnorm_scalar(const saev_b200_handle* h, void* workspace);
/* float[d_sae]: max(||w_j||, ||fp16(w_j)||) (rounded up), the per-column input of the screen's error bound; a sharded optimizer
 * step refreshes rows [row_begin, row_end) only -- all-gather it with the fp16 operand. */
float* saev_b200_wnorm_rows(const saev_b200_handle* h, void* workspace);

/* clip_grad_norm_(max_norm) + Adam(fused) step + optional decoder row renorm + fp16 operand refresh.
 *   g_eff = grads * grad_scale;  coef = min(1, max_norm / (||g_eff|| + 1e-6))  (max_norm <= 0: no clip)
 *   `step` is the 1-based Adam step count AFTER this update (bias corrections use it).
 *   gnorm_out (optional, device) receives ||g_eff||, the value clip_grad_norm_ returns.
 *   `parts`: SAEV_B200_ADAM_ENCODER (W_enc_t, b_enc and what the screen keeps of them), SAEV_B200_ADAM_DECODER (W_dec with
 *   the renorm, b_dec, gnorm_out) or both.  The next forward's screen reads only the encoder side, so a caller may
 *   run the decoder half on a second stream beside it (same step, lr, sumsq; the HBM-bound update hides behind the
 *   tensor-bound screen); the decoder-only launch uses small blocks that fit beside a resident screen CTA.
 *   A sharded optimizer that owns SEVERAL row ranges (saev_b200_set_optimizer_shard before each call) passes
 *   SAEV_B200_ADAM_ALL for the first range and ALL | ROWS_ONLY | KEEP_MAXIMA for the others: the bias vectors are updated
 *   once, and the dictionary-wide maxima the screen's error bound uses accumulate over the ranges. */
enum { SAEV_B200_ADAM_ENCODER = 1, SAEV_B200_ADAM_DECODER = 2, SAEV_B200_ADAM_ALL = 3, SAEV_B200_ADAM_ROWS_ONLY = 4,
       SAEV_B200_ADAM_KEEP_MAXIMA = 8 };
int saev_b200_adam_step(saev_b200_handle* h, float* W_enc_t, float* b_enc, float* W_dec, float* b_dec,
                        const float* grads_flat, float* m_flat, float* v_flat, float lr, float beta1,
                        float beta2, float eps, int64_t step, float max_norm, float grad_scale,
                        const float* sumsq, int32_t renorm_w_dec, float* gnorm_out, int32_t parts, void* workspace,
                        void* stream);

/* Lazy dense views for saev's logging block (train.py:365-442). */
int saev_b200_densify(saev_b200_handle* h, const int32_t* topk_idx, const float* topk_val, int32_t B,
                      float* f_x_out /* [B, d_sae] */, void* stream);
/* Matryoshka prefix cuts for the following forward/backward calls (saev samples them per step on the host,
 * objectives.py:125,158-201): n strictly increasing column counts, the last one == d_sae.  n == 1 (or NULL) selects
 * the plain single-prefix objective.  x_hat_i = b_dec + sum of the active columns below cut i; the MSE is the mean
 * over batch x prefixes x d_model; AuxK and `resid` refer to the last (full) prefix.  TopK: one sparse decode that emits
 * every prefix (top_k <= 128).  ReLU: the decoder, dh and W_dec-gradient contractions run once per prefix block. */
int saev_b200_set_prefixes(saev_b200_handle* h, const int32_t* host_prefixes, int32_t n);
/* x_hats[B, n_prefixes, d_model] of the last forward (modeling.py:406); n_prefixes == 1 reduces to saev_b200_x_hat. */
int saev_b200_x_hats(saev_b200_handle* h, const float* resid, const float* x, int32_t B, float* x_hats_out,
                     void* workspace, void* stream);

/* Same for either activation: TopK scatters (topk_idx, topk_val); ReLU joins the bf16 (hi, lo) pair the dense path
 * keeps in the workspace (f to ~2^-17 relative). */
int saev_b200_dense_f(saev_b200_handle* h, const int32_t* topk_idx, const float* topk_val, int32_t B,
                      float* f_x_out /* [B, d_sae] */, void* workspace, void* stream);
int saev_b200_x_hat(saev_b200_handle* h, const float* resid, const float* x, int32_t B,
                    float* x_hat_out /* [B, d_model] */, void* stream);

/* Dictionary coherence of the log block, saev train.py:415-421:
 *     W_norm = W / W.norm(dim=1);  coherence = (W_norm @ W_norm.T).abs().triu(1).max()
 * without the [d_sae, d_sae] Gram matrix: the upper triangle is screened on the tensor cores (two-piece bf16 split,
 * ~2^-16 absolute on unit rows) and the pairs within 1e-3 of the screen maximum are recomputed from the fp32 rows
 * with fp64 accumulation.  out[0] = coherence, out[1] = screen maximum, out[2], out[3] = the pair (i < j, as floats).
 * scratch: saev_b200_coherence_scratch_bytes(h) bytes of device memory.  d_model must be a multiple of 8. */
size_t saev_b200_coherence_scratch_bytes(const saev_b200_handle* h);
int saev_b200_dictionary_coherence(saev_b200_handle* h, const float* W_dec, void* scratch, size_t scratch_bytes,
                                   float* out /* [4] */, void* stream);

/* The per-SAE metrics of saev's log block (train.py:380-423) for the batch of the LAST forward on this handle:
 *   out[0] explained_variance   1 - var(x - x_hat) / var(x)            (:407)
 *   out[1] dead_unit_pct        fraction of atoms that did not fire in the batch (:410; from the activity flags of
 *                               phase A, i.e. |f| > 0 where the reference tests |f| > 1e-12)
 *   out[2] dictionary_coherence (:415-418, as saev_b200_dictionary_coherence)
 *   out[3] avg_decoder_row_norm (:420)
 *   out[4] sse_sae   out[5] sse_baseline = sum x^2 - |sum_b x|^2 / B   out[6] normalized_mse   (:380-406, fp64)
 *   out[7] coherence screen maximum (diagnostic)
 * x[B, d_model] and resid[B, d_model] (= x_hat - x, as written by saev_b200_forward) are read once; nothing of size
 * [B, d_sae] or [d_sae, d_sae] is formed.  out: device double[8].  scratch: saev_b200_log_scratch_bytes(h). */
size_t saev_b200_log_scratch_bytes(const saev_b200_handle* h);
int saev_b200_log_metrics(saev_b200_handle* h, const float* x, const float* resid, int32_t B, const float* W_dec,
                          void* workspace, void* scratch, size_t scratch_bytes, double* out /* [8] */, void* stream);

/* One batch of saev's evaluate() loop (train.py:546-566), from the state of the LAST forward on this handle, without
 * the dense fwd.f_x[B, d_sae]:
 *   n_fired[j] += #(f[b, j] > 0)          (:560-562)      values[j] += sum_b f[b, j]      (:563)
 *   acc[0] += sum x^2   acc[8 + d] += sum_b x[b, d]       (:548-549, fp64)      acc[1] += sum (x - x_hat)^2   (:558-559)
 *   acc[4] += l0 * B    acc[5] += l1 * B    acc[6] += mse * B    acc[7] += B     (:564-566; acc[2], acc[3]: sums of
 *   the residual and of x, used by the log block only)
 * acc: device double[8 + d_model], n_fired / values: device float[d_sae]; all three are ACCUMULATED into (the caller
 * zeroes them before the first batch).  topk_idx / topk_val: as written by saev_b200_forward (ignored for ReLU, whose
 * dense activations live in the workspace).  losses: the device float[8] of the same forward. */
int saev_b200_eval_accumulate(saev_b200_handle* h, const float* x, const float* resid, int32_t B,
                              const int32_t* topk_idx, const float* topk_val, const float* losses, double* acc,
                              float* n_fired, float* values, void* workspace, void* stream);

/* Test hook: AuxK's selection of the last training forward (valid until the next forward).  mask[b * ld + i] != 0 iff
 * the i-th entry of dead_list (ascending atom ids, *n_dead of them) is among row b's top-k_aux dead latents
 * (modeling.py:93-97).  All four outputs are device pointers into the workspace. */
int saev_b200_aux_selection(const saev_b200_handle* h, void* workspace, const uint8_t** mask, int64_t* ld,
                            const int32_t** dead_list, const int32_t** n_dead);

/* Test hook for the tensor-core contraction alone: out[M, N] = A[M, K] . Bt[N, K]^T + bias[N], computed
 * from bf16 copies of the operands (nterms = 1), the 3-term two-piece split (nterms = 3, ~2^-16 of sum |a b|) or the
 * 6-term three-piece split (nterms = 6, fp32-class).  scratch must hold 3 * (M + N) * K bf16. */
int saev_b200_gemm_nt(saev_b200_handle* h, const float* A, const float* Bt, const float* bias, int32_t M,
                      int32_t N, int32_t K, int32_t nterms, float* out, void* scratch, void* stream);

/* Optional per-stage device timing (CUDA events recorded on the caller's stream around each group of
 * launches).  saev_b200_profile_read synchronises on the recorded events and returns, per stage, the summed
 * milliseconds and the number of recorded intervals since the last read (host arrays of
 * SAEV_B200_N_STAGES entries). */
enum {
  SAEV_B200_STAGE_PREP = 0,        /* x -> fp16 operand (row-scaled)           */
  SAEV_B200_STAGE_ENCODE_GEMM = 1, /* tcgen05 encoder contraction + top-k screen */
  SAEV_B200_STAGE_RESCORE = 2,     /* exact fp32 re-score + final top-k        */
  SAEV_B200_STAGE_DECODE = 3,      /* sparse decode, residual, d loss / d h    */
  SAEV_B200_STAGE_LOSS = 4,        /* dead tracker, AuxK forward, loss scalars */
  SAEV_B200_STAGE_CSC = 5,         /* per-atom lists of the active set         */
  SAEV_B200_STAGE_WGRAD = 6,       /* weight gradients + parallel-grad removal */
  SAEV_B200_STAGE_BIAS_AUX = 7,    /* b_dec gradient, AuxK backward            */
  SAEV_B200_STAGE_SUMSQ = 8,       /* global gradient norm                     */
  SAEV_B200_STAGE_ADAM = 9,        /* clip + Adam + renorm + fp16 refresh      */
  SAEV_B200_N_STAGES = 10
};
int saev_b200_profile_enable(saev_b200_handle* h, int32_t on);
int saev_b200_profile_read(saev_b200_handle* h, float* host_ms_sum, int32_t* host_count);

/* ---- pinned staging ring for activation batches (replaces the pageable-memory H2D copy of
 *      train.py:333 / buffers.py:199) ----
 * `n_slots` pinned host buffers of `slot_bytes`; saev_b200_ring_submit enqueues cudaMemcpyAsync of a
 * filled slot to `dst_device` on the ring's own copy stream and records an event;
 * saev_b200_ring_wait makes `consumer_stream` wait for that event (no host sync). */
typedef struct saev_b200_ring saev_b200_ring;
int saev_b200_ring_create(int32_t n_slots, size_t slot_bytes, saev_b200_ring** out);
int saev_b200_ring_destroy(saev_b200_ring* r);
void* saev_b200_ring_host_ptr(saev_b200_ring* r, int32_t slot);
int saev_b200_ring_submit(saev_b200_ring* r, int32_t slot, void* dst_device, size_t bytes);
int saev_b200_ring_wait(saev_b200_ring* r, int32_t slot, void* consumer_stream);
int saev_b200_ring_host_sync(saev_b200_ring* r, int32_t slot); /* block host until the slot's copy is done */

/* ---- shuffled activation loader with the shuffle pool in HBM (replaces saev's ShuffledDataLoader +
 *      ReservoirBuffer: src/saev/data/shuffled.py:133-363, 380-699; src/saev/data/buffers.py:91-231) ----
 * Shards are saev's `acts%06d.bin` files, fp32 [examples_per_shard, n_layers, tokens_per_example, d_model]
 * (src/saev/data/shards.py:168-180).  I/O threads pread whole examples into pinned staging chunks (or, for shards on
 * tmpfs, map + register the files so that the copies DMA out of the page cache: cfg.reserved), a feeder
 * thread appends them to a device-resident pool and prepares shuffled batches ahead of the consumer with a
 * gather kernel; saev_b200_loader_next returns DEVICE pointers, valid until the call after the next one.
 * Each (example, content token) row of the listed shards is delivered exactly once per epoch. */
typedef struct saev_b200_loader_cfg {
  const char* shards_dir;         /* directory holding acts%06d.bin                      shuffled.py:189-196 */
  int32_t examples_per_shard;     /* Metadata.examples_per_shard                         shards.py:158-166 */
  int32_t n_layers;               /* len(Metadata.layers) */
  int32_t tokens_per_example;     /* content tokens + [CLS]                              shards.py:136-144 */
  int32_t d_model;
  int32_t layer_index;            /* Metadata.layers.index(cfg.layer)                    shuffled.py:160 */
  int32_t cls_token;              /* 1: token 0 is [CLS], content starts at token 1      shuffled.py:204 */
  int32_t content_tokens;         /* Metadata.content_tokens_per_example */
  const int32_t* shard_order;     /* [n_order] shard ids this rank visits, in order      shuffled.py:326-328 */
  const int32_t* shard_examples;  /* [n_order] valid examples of each (shards.json)      shuffled.py:200-202 */
  int32_t n_order;
  int32_t batch_size;             /* Config.batch_size */
  int32_t pool_batches;           /* Config.buffer_size: pool capacity in batches        shuffled.py:459-468 */
  int32_t n_threads;              /* Config.n_threads */
  int32_t n_out_slots;            /* device batch buffers handed out round-robin (>= 2; 0 = 3) */
  int32_t chunk_examples;         /* examples per I/O chunk, 0 = about 8 MB */
  float min_buffer_fill;          /* Config.min_buffer_fill                              shuffled.py:578-633 */
  int32_t reserved;               /* zero-copy mode: 0 = automatic (on when the shards live on tmpfs), 1 = off, 2 = on */
  int64_t n_rows_limit;           /* unused by the library (the epoch length is passed to start_epoch) */
  uint64_t seed;                  /* Config.seed */
  const uint8_t* labels;          /* optional labels.bin [n_examples, content_tokens] (host pointer) or NULL */
  const uint8_t* ignore_lut;      /* optional [256]: 1 = drop rows carrying this label   shuffled.py:206-236 */
} saev_b200_loader_cfg;
typedef struct saev_b200_loader saev_b200_loader;

int saev_b200_loader_create(const saev_b200_loader_cfg* cfg, saev_b200_loader** out);
int saev_b200_loader_destroy(saev_b200_loader* l);
/* Starts the I/O + feeder threads for one pass over the configured shards; the epoch ends after `rows_expected`
 * rows were handed out.  Buffers still in use on `consumer_stream` from the previous epoch are protected. */
int saev_b200_loader_start_epoch(saev_b200_loader* l, int64_t rows_expected, uint64_t seed, void* consumer_stream);
/* Next shuffled batch (device pointers; *n_rows = 0 once the epoch is exhausted).  Work later enqueued on
 * `consumer_stream` sees the batch (event wait, no host sync).  Returns 131 if nothing arrived within
 * `timeout_s` seconds (the call can simply be repeated: nothing is consumed), 130 if a loader thread failed. */
int saev_b200_loader_next(saev_b200_loader* l, void* consumer_stream, double timeout_s, float** act,
                          int32_t** example_idx, int32_t** token_idx, int32_t* n_rows);
int saev_b200_loader_stats(saev_b200_loader* l, int64_t* pool_rows, int64_t* pool_capacity, int64_t* rows_delivered,
                           int64_t* bytes_read);
/* 1 if the I/O threads mmap + cudaHostRegister the shard files and the H2D copies DMA straight out of the page cache
 * (no pread -> pinned staging memcpy on the host), 0 if they pread into pinned staging chunks. */
int saev_b200_loader_zero_copy(const saev_b200_loader* l);
int saev_b200_loader_stop(saev_b200_loader* l);
const char* saev_b200_loader_last_error(const saev_b200_loader* l);
/* Host-only pieces, exported for tests: the chunk reader (returns rows kept, < 0 on error) and the draw /
 * compaction planner (`need` distinct positions of [0, fill) + the moves that re-pack the pool). */
int64_t saev_b200_loader_read_chunk(const saev_b200_loader_cfg* cfg, int32_t shard, int32_t ex_begin, int32_t n_ex,
                                    float* host_act, int32_t* host_meta);
int saev_b200_loader_plan_draw(uint64_t seed, int64_t fill, int32_t need, int32_t* sel, int32_t* mv_src,
                               int32_t* mv_dst, int32_t* n_moves);

#ifdef __cplusplus
}
#endif
#endif /* SAEV_B200_H_ */
