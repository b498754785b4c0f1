/* carel_head.h — C ABI of libcarel_head.so: the B200-native (sm_100a) CAREL-VAE disentanglement head.
 *
 * The reference (tk1363704/CAREL-VAE) has no FFI layer: its boundary for this path is the set
 * of Python callables the drl_classifier_ec_{mmd,hsic,vi,gan}*.py trainers define and call by
 * name.  Each entry point below is what the drop-in Python module (carel_vae_b200/) binds with
 * ctypes in place of the ATen call sequence of the cited reference function.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; fp32 row-major unless noted;
 *     `ld*` are row pitches in elements;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises;
 *   - the library allocates nothing: scratch comes from the caller (`*_workspace_bytes`);
 *     a workspace filled by a *_fwd call is the context its *_bwd call consumes;
 *   - return 0 on success; otherwise carel_last_error() (thread-local) describes the failure;
 *   - scalar results (losses) are written to device memory, so no call forces a host sync.
 */
#ifndef CAREL_HEAD_H_
#define CAREL_HEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAREL_VERSION 100 /* 0.1.0 */

/* flags */
#define CAREL_MMD_FULL_GRAM 1 /* compute all (2N)^2 entries like the reference instead of j > i */
#define CAREL_MMD_NO_PAIRS 2 /* long-K Gram on single CTAs (gemm_tc) instead of cta_group::2 pairs (debug / A-B) */
#define CAREL_BF16 0
#define CAREL_FP32 1

int carel_version(void);
const char* carel_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t carel_launch_count(void);
/* Per-kernel timing for eager profiling passes (never under CUDA-graph capture): when enabled every
 * tensor-core kernel launch is bracketed by CUDA events on its stream; carel_timing_report writes
 * "tag total_ms count" lines into buf (device-synchronising) and clears the log. */
void carel_timing_enable(int on);
/* Data-parallel overlap (no reference counterpart: the reference has no multi-GPU path): while n > 0
 * every kernel launched by this library sizes its grid for (SM count - n) SMs, so that a collective running on
 * another stream (the NCCL all-reduce of the decoder gradient under the rest of the backward pass) finds free SMs
 * instead of queueing behind the one-CTA-per-SM persistent kernels.  Process-wide, read at launch time (a captured
 * CUDA graph keeps the grid it was captured with). */
void carel_set_reserved_sms(int n);
int carel_get_reserved_sms(void);
/* Side branches (no reference counterpart): work inside ONE entry point that does not depend on the rest of it is
 * enqueued on a library-owned second stream per device, forked from and joined back into the caller's stream with
 * event record / wait pairs; under CUDA-graph capture it becomes a parallel branch of the graph.  Bit mask:
 *   1: carel_bow_fwd — what depends on the BoW target only (dense -> ELL scan or the copy of a caller's ELL triple,
 *      then the CSC view the backward gathers through) beside the operand packing and the tail of the dense pass;
 *   2: carel_bow_bwd — the two O(nnz) gathers (grad_z | grad_w, grad_b: disjoint outputs) side by side.
 * Results are identical (the same kernels on the same data).  Process-wide; the initial value is the environment
 * variable CAREL_SIDE_BRANCHES when it is set. */
void carel_set_side_branches(int mask);
int carel_get_side_branches(void);
int carel_timing_report(char* buf, size_t n);

/* ------------------------------------------------------------------------------------------
 * Debug / building block: C[M,N] (+)= A * B^T on the tcgen05 core (bf16 operands, fp32 accum).
 *   a_mn = 0: A is [M, K] row-major (K-major);  a_mn = 1: A is [K, M] row-major (MN-major)
 *   b_mn = 0: B is [N, K] row-major;            b_mn = 1: B is [K, N] row-major
 *   a, b are bf16; ld in elements (multiples of 8).  ksplit > 1 accumulates atomically (C must
 *   be zeroed by the caller).
 * ------------------------------------------------------------------------------------------ */
int carel_gemm_bf16(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn,
                    int64_t M, int64_t N, int64_t K, float* c, int64_t ldc, const float* bias,
                    int ksplit, void* stream);
/* The same product on CTA pairs (cluster of 2, tcgen05 cta_group::2; 256 x 256 tiles). */
int carel_gemm_bf16_2sm(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, int64_t M,
                        int64_t N, int64_t K, float* c, int64_t ldc, const float* bias, int ksplit, void* stream);

/* ------------------------------------------------------------------------------------------
 * MMD  — replaces MMDStatistic.__call__ + pdist
 *        (drl_classifier_ec_mmd_final_mul.py:537-569, :580-596; call site :231-233).
 *   mmd = 2 a01 sum K_xy + a00 (sum K_xx - tr K_xx) + a11 (sum K_yy - tr K_yy),
 *   K = sum_alpha exp(-alpha (|d2| + eps)),  d2 = |x_i|^2 + |y_j|^2 - 2 x_i.y_j.
 *   The (n1+n2)^2 Gram matrix never leaves the SM.
 * ------------------------------------------------------------------------------------------ */
/* precision: CAREL_BF16 (0) = bf16 tensor-core operands; CAREL_FP32 (1) = every operand split into two bf16
 * terms (x1 x1' + x2 x1' + x1 x2' in one K-concatenated bf16 GEMM, fp32 accumulate): ~fp32-accurate results
 * (1e-5) at 3x the tensor work.  The fused backward supports dim <= 250 (bf16) / 120 (fp32). */
size_t carel_mmd_workspace_bytes(int64_t n1, int64_t n2, int64_t dim, int precision);
/* Forward.  Packs cat(x, y) to bf16 in `ws`, accumulates the three block sums of Gram row tiles
 * [row_tile_begin, row_tile_end) (128-row tiles; pass 0, -1 for all) into partial_sums[3]
 * (double: xx off-diagonal, yy off-diagonal, xy + yx) and, if out_mmd != NULL, finalises
 * out_mmd[0].  Multi-GPU callers all-reduce partial_sums and call carel_mmd_finalize. */
int carel_mmd_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t n1,
                  int64_t n2, int64_t dim, const float* alphas_host, int n_alpha, float eps,
                  int flags, int64_t row_tile_begin, int64_t row_tile_end, double* partial_sums,
                  float* out_mmd, void* ws, size_t ws_bytes, int precision, void* stream);
int carel_mmd_finalize(const double* partial_sums, int64_t n1, int64_t n2, float* out_mmd,
                       void* stream);
/* Backward: grad_x[n1,dim], grad_y[n2,dim] (+)= d mmd / d(x,y) * grad_out[0]; rows of the Gram
 * [row_tile_begin, row_tile_end) only (a rank's rows see every column, so its gradient rows are
 * complete without a reduction).  `ws` must be the workspace of the matching carel_mmd_fwd. */
int carel_mmd_bwd(const float* grad_out, int64_t n1, int64_t n2, int64_t dim,
                  const float* alphas_host, int n_alpha, float eps, int64_t row_tile_begin,
                  int64_t row_tile_end, float* grad_x, int64_t ldgx, float* grad_y, int64_t ldgy,
                  void* ws, size_t ws_bytes, int precision, void* stream);

/* Dense pairwise matrices for API completeness (reference pdist :580-596 with norm=2, the
 * ret_matrix=True kernel matrix :566-567, hsic.py pairwise_distances :529-532 and
 * GaussianKernelMatrix :535-537).  mode: 0 kernel sum, 1 pdist, 2 squared distances, 3 exp(-d2/sigma)
 * with alphas_host[0] = 1/sigma.  out is [n1, n2] fp32. */
size_t carel_pairwise_workspace_bytes(int64_t n1, int64_t n2, int64_t dim);
int carel_pairwise(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t n1,
                   int64_t n2, int64_t dim, int mode, const float* alphas_host, int n_alpha,
                   float eps, float* out, int64_t ldo, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Latent layer — replaces get_emotion_emb / get_cause_emb (drl_classifier_ec_mmd_final_mul.py
 * :312-336), sample_prior (:345-351) and get_kl_loss (:525-534).
 *   w4[4], b4[4]: emotion_mu, emotion_log_var, cause_mu, cause_log_var ([L, D] / [L], nn.Linear layout)
 *   eps_e, eps_c: the [L] noise vectors the caller drew with torch.randn (shared by all rows)
 *   pre [N, 4L] = [mu_e | lv_e | mu_c | lv_c];  z [N, 2L] = [z_e | z_c];  kl[2] = (kl_e, kl_c)
 * ------------------------------------------------------------------------------------------ */
size_t carel_latent_workspace_bytes(int64_t N, int64_t D, int64_t L);
int carel_latent_fwd(const float* x, int64_t ldx, int64_t N, int64_t D, int64_t L,
                     const float* const* w4, const float* const* b4, const float* eps_e,
                     const float* eps_c, float* pre, float* z, float* kl, void* ws, size_t ws_bytes,
                     void* stream);
/* grad_z [N, 2L] (may be NULL), grad_kl[2] device (may be NULL); grad_x [N, D] (may be NULL);
 * grad_w4[i] / grad_b4[i] may be NULL individually. */
int carel_latent_bwd(const float* grad_z, int64_t ldgz, const float* grad_kl, const float* pre,
                     const float* eps_e, const float* eps_c, int64_t N, int64_t D, int64_t L,
                     float* grad_x, int64_t ldgx, float* const* grad_w4, float* const* grad_b4,
                     void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * BoW reconstruction — replaces nn.Softmax(dim=1)(self.decoder(generative_emb)) + get_reconstruct_loss
 * (drl_classifier_ec_mmd_final_mul.py:253-254, :381-387): BCELoss(softmax(Z W^T + b), bow (1-ls) + ls/V).
 *   z [N, K] (K = 2 ec_dim, multiple of 16, <= 256), dec_w [V, K], dec_b [V]
 *   target: dense bow [N, V] (reference layout) and/or an ELL list (cols/vals [N, cap], nnz [N]);
 *   rows whose nnz exceeds ell_cap are re-scanned from the dense pointer, which must then stay valid
 *   for the backward call.  loss_out: device scalar.
 * ------------------------------------------------------------------------------------------ */
size_t carel_bow_workspace_bytes(int64_t N, int64_t K, int64_t V, int64_t ell_cap);
/* Dense target -> ELL list on its own (what ECPEDataset._get_bow_representations + the H2D copy hand over,
 * :100-119, :830): cols / vals [N, ell_cap], nnz [N] (nnz may exceed ell_cap: such rows are re-scanned from the dense
 * rows by carel_bow_fwd / carel_bow_bwd).  HBM-bound: N V 4 bytes read once (6.6 TB/s measured at N = 8192,
 * V = 30000). */
int carel_bow_pack_ell(const float* bow_dense, int64_t ld_bow, int64_t N, int64_t V, int64_t ell_cap,
                       int32_t* ell_cols, float* ell_vals, int32_t* ell_nnz, void* stream);
/* The same scan on the HOST (plain C++, std::thread; no CUDA call): for callers whose dense rows live in host memory,
 * as the reference's DataLoader yields them (:100-119) before the `.to(device)` of :830.  Same output as
 * carel_bow_pack_ell as a SET per row (v != 0.f, nnz counts every non-zero, the first ell_cap are stored); the host list
 * is in ascending column order, the GPU scan interleaves the four columns of a lane inside a 128-column group.
 * n_threads <= 0: one per hardware thread.  The list (27 MB at N = 8192, cap = 32) replaces a 983 MB host-to-device copy. */
int carel_bow_pack_ell_host(const float* bow_dense, int64_t ld_bow, int64_t N, int64_t V, int64_t ell_cap,
                            int32_t* ell_cols, float* ell_vals, int32_t* ell_nnz, int n_threads);
int carel_bow_fwd(const float* z, int64_t ldz, const float* dec_w, const float* dec_b,
                  const float* bow_dense, int64_t ld_bow, const int32_t* ell_cols,
                  const float* ell_vals, const int32_t* ell_nnz, int64_t N, int64_t K, int64_t V,
                  int64_t ell_cap, float label_smoothing, float* loss_out, void* ws, size_t ws_bytes,
                  void* stream);
int carel_bow_bwd(const float* grad_out, const float* bow_dense, int64_t ld_bow, int64_t N, int64_t K,
                  int64_t V, int64_t ell_cap, float label_smoothing, float* grad_z, int64_t ldgz,
                  float* grad_w, float* grad_b, void* ws, size_t ws_bytes, void* stream);
/* carel_bow_bwd in two parts (data-parallel overlap; the reference has no multi-GPU path): parts = 1 runs everything
 * that completes grad_w / grad_b (and initialises grad_z), parts = 2 the dense grad_z kernel, 3 = both = carel_bow_bwd.
 * Call 1 then 2 with the same arguments; between them grad_w may be handed to a collective on another stream. */
int carel_bow_bwd_parts(const float* grad_out, const float* bow_dense, int64_t ld_bow, int64_t N, int64_t K,
                  int64_t V, int64_t ell_cap, float label_smoothing, float* grad_z, int64_t ldgz,
                  float* grad_w, float* grad_b, void* ws, size_t ws_bytes, int parts, void* stream);

/* ------------------------------------------------------------------------------------------
 * HSIC — replaces HSIC / GaussianKernelMatrix / pairwise_distances (drl_classifier_ec_hsic.py:529-547):
 *   trace(L H K H) / (m-1)^2 with K = exp(-d2_x / s_x), L = exp(-d2_y / s_y), evaluated in O(m^2)
 *   (no centring matmuls).  out: device scalar.  The workspace of the forward call carries the row
 *   sums the backward call needs.
 * ------------------------------------------------------------------------------------------ */
size_t carel_hsic_workspace_bytes(int64_t m, int64_t dx, int64_t dy);
/* Forward of Gram row tiles [row_tile_begin, row_tile_end) (128-row tiles; 0, -1 = all).  With out != NULL
 * the value is finalised.  Multi-GPU: every rank passes its row block with out = NULL, all-reduces the
 * three regions of `ws` that carel_hsic_ws_layout names (row sums rK, rL: m doubles each; acc: 4 doubles;
 * offsets in bytes) and calls carel_hsic_finalize.  Backward of a row block needs no collective. */
int carel_hsic_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t m, int64_t dx,
                   int64_t dy, float s_x, float s_y, int64_t row_tile_begin, int64_t row_tile_end,
                   float* out, void* ws, size_t ws_bytes, void* stream);
int carel_hsic_ws_layout(int64_t m, int64_t dx, int64_t dy, int64_t* offsets3);
int carel_hsic_finalize(int64_t m, int64_t dx, int64_t dy, float* out, void* ws, size_t ws_bytes,
                        void* stream);
int carel_hsic_bwd(const float* grad_out, int64_t m, int64_t dx, int64_t dy, float s_x, float s_y,
                   int64_t row_tile_begin, int64_t row_tile_end, float* grad_x, int64_t ldgx,
                   float* grad_y, int64_t ldgy, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Classifier heads — replaces get_emotion_mul_loss / get_cause_mul_loss / get_pair_mul_loss
 * (drl_classifier_ec_mmd_final_mul.py:461-513; drl_classifier_ec_hsic.py:459-505 for emo_mode 1) and their weighted
 * sum in the loss assembly (:256-261):
 *   emotion: W_e (z_e (.) m_e) + b_e -> CrossEntropy over C_e classes (emo_mode 0, int64 labels [N]) or
 *            sigmoid + BCELoss against y (1-ls) + ls / n_e (emo_mode 1, float labels [N]);
 *   cause:   w_c . (z_c (.) m_c) + b_c -> sigmoid + BCELoss against y (1-ls) + ls / n_c;
 *   pair:    w_p . (z (.) m_p) + b_p  -> BCEWithLogits(., y (1-ls) + ls / n_p, pos_weight = (N - sum y) / sum y),
 *            an infinite pair loss is replaced by 0 if guard_inf (reference :510-511).
 * z [N, 2L] = [z_e | z_c]; masks3 = {m_e [N,L], m_c [N,L], m_p [N,2L]}: dropout masks already scaled by 1/(1-p)
 * (host array of three device pointers; NULL or NULL entries = no dropout); weights6 = {W_e [C_e,L], b_e [C_e],
 * W_c [L], b_c [1], W_p [2L], b_p [1]} (host array of device pointers).  losses4 (device) = emotion, cause, pair and
 * loss_weights3_host . (emotion, cause, pair).  The backward takes d/d(losses4[3]) and writes grad_z [N, 2L]
 * (all three heads) and the six weight gradients (grad_weights6 entries may be NULL).
 * ------------------------------------------------------------------------------------------ */
size_t carel_heads_workspace_bytes(int64_t N, int64_t C_e);
int carel_heads_fwd(const float* z, int64_t ldz, int64_t N, int64_t L, int64_t C_e, int emo_mode,
                    const float* const* masks3, const float* const* weights6, const void* emo_labels,
                    const float* cau_labels, const float* pair_labels, float label_smoothing, float n_e,
                    float n_c, float n_p, int guard_inf, const float* loss_weights3_host, float* losses4,
                    void* ws, size_t ws_bytes, void* stream);
int carel_heads_bwd(const float* grad_out, const float* z, int64_t ldz, int64_t N, int64_t L, int64_t C_e,
                    int emo_mode, const float* const* masks3, const float* const* weights6,
                    const void* emo_labels, const float* cau_labels, const float* pair_labels,
                    float label_smoothing, float n_e, float n_c, float n_p, int guard_inf,
                    const float* loss_weights3_host, float* grad_z, int64_t ldgz,
                    float* const* grad_weights6, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Loss assembly — replaces the chain of scalar tensor ops that forms the training loss
 * (drl_classifier_ec_mmd_final_mul.py:256-261: 30 (-mmd) + heads + w KL_e + w KL_c + recon; drl_classifier_ec_hsic.py:
 * 253-258; drl_classifier_ec_vi_final.py:261-266; drl_classifier_ec_gan.py:275-281).
 *   out = sum_i weights_host[i] * *terms_host[i]      (fp32, left to right, products and sums rounded separately:
 *                                                      bit-identical to the torch expression it replaces)
 * terms_host = host array of n device pointers to fp32 scalars, n <= 16.  The backward writes
 * grads[i] = weights_host[i] * *grad_out.
 * ------------------------------------------------------------------------------------------ */
int carel_wsum_fwd(const float* const* terms_host, const float* weights_host, int n, float* out, void* stream);
int carel_wsum_bwd(const float* grad_out, const float* weights_host, int n, float* grads, void* stream);

/* ------------------------------------------------------------------------------------------
 * CLUB terms of the VI variant — replace the elementwise tails of get_ec_aprx_loss / get_ec_upper_loss
 * (drl_classifier_ec_vi_final.py:421-426, :428-439); the two small MLPs ec_mu / ec_log_var stay with the caller.
 *   mu_hat, lv_hat [N, L] contiguous; e = e_embedding [N, L] with row pitch lde (a column view of z is fine);
 *   perm == NULL:  out = mean_i sum_d [ (mu - e)^2 / exp(lv) + lv ]                               (aprx loss)
 *   perm != NULL:  out = mean_i sum_d [ (mu - e[perm[i]])^2 - (mu - e)^2 ] / exp(lv) / 2          (upper bound; perm =
 *                  the caller's torch.randperm(N) as int64 on the device, a permutation of 0..N-1)
 *   acc: 8 bytes of device scratch.  The backward writes grad_mu / grad_lv / grad_e [N, L] contiguous (each may be
 *   NULL) for d(out) = *grad_out.
 * ------------------------------------------------------------------------------------------ */
int carel_club_fwd(const float* mu_hat, const float* lv_hat, const float* e, int64_t lde, const int64_t* perm,
                   int64_t N, int64_t L, float* out, double* acc, void* stream);
int carel_club_bwd(const float* grad_out, const float* mu_hat, const float* lv_hat, const float* e, int64_t lde,
                   const int64_t* perm, int64_t N, int64_t L, float* grad_mu, float* grad_lv, float* grad_e,
                   void* stream);

/* ------------------------------------------------------------------------------------------
 * CLUB approximation network q(e | c) of the VI trainer — replaces get_ec_emb (drl_classifier_ec_vi_final.py:342-347) over
 * ec_mu = Linear-ReLU-Linear and ec_log_var = Linear-ReLU-Linear-Tanh (:155-162): both MLPs and both layers in one
 * kernel per direction (ec_dim <= 128, multiple of 4).
 *   w8[8] = {ec_mu.0.weight, ec_mu.0.bias, ec_mu.2.weight, ec_mu.2.bias, ec_log_var.0.weight, .0.bias, .2.weight, .2.bias}
 *   mu, lv [N, L] contiguous.  Backward: dmu / dlv (may be NULL), dc (NULL for a detached input), g8[8] (each may be NULL;
 *   accumulated: the caller passes them zeroed).
 * ------------------------------------------------------------------------------------------ */
size_t carel_club_mlp_workspace_bytes(int64_t N, int64_t L);
int carel_club_mlp_fwd(const float* c, int64_t ldc, int64_t N, int64_t L, const float* const* w8, float* mu, float* lv,
                       void* ws, size_t ws_bytes, void* stream);
int carel_club_mlp_bwd(const float* dmu, const float* dlv, const float* c, int64_t ldc, const float* lv, int64_t N, int64_t L,
                       const float* const* w8, float* dc, int64_t lddc, float* const* g8, void* ws, size_t ws_bytes,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * GAN discriminators — replace get_ec_disc_preds / get_ce_disc_preds / get_ec_disc_loss / get_entropy_loss
 * (drl_classifier_ec_gan.py:430-477; call sites in forward :222-233), ec_num_class = 1:
 *   p_ec = sigmoid(ec_disc(dropout(z_c.detach()))), p_ce = sigmoid(ce_disc(dropout(z_e.detach())))
 *   out[3] = { BCELoss(p_ec, smooth(y_emo)), BCELoss(p_ce, smooth(y_cau)), mean p_ec log(p_ec + eps) + mean p_ce log(p_ce + eps) }
 * masks[2] = {m_ec, m_ce}: scaled dropout masks [N, L] or NULL; params[4] = {ec_disc.weight [L], ec_disc.bias [1],
 * ce_disc.weight [L], ce_disc.bias [1]}.  Backward: g[3] device scalars (NULL = 0) -> grads[4] of the four parameters.
 * ------------------------------------------------------------------------------------------ */
size_t carel_disc_workspace_bytes(int64_t N);
int carel_disc_fwd(const float* z, int64_t ldz, int64_t N, int64_t L, const float* const* masks, const float* const* params,
                   const float* y_emo, const float* y_cau, float label_smoothing, float n_class, float eps, float* out,
                   void* ws, size_t ws_bytes, void* stream);
int carel_disc_bwd(const float* const* g, const float* z, int64_t ldz, int64_t N, int64_t L, const float* const* masks,
                   const float* const* params, const float* y_emo, const float* y_cau, float label_smoothing, float n_class,
                   float eps, float* const* grads, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Inference — replaces get_pair_preds (drl_classifier_ec_mmd_final_mul.py:265-282): out[N] = sigmoid(pair_classifier(
 * [sample_prior(emotion) | sample_prior(cause)])) in 4 launches (pack X, pack W4, projection GEMM, one fused
 * reparameterisation + dot product + sigmoid pass); no KL sums, no z.  pair_w [2L], pair_b [1]; precision 0 or 2.
 * ------------------------------------------------------------------------------------------ */
size_t carel_pair_preds_workspace_bytes(int64_t N, int64_t D, int64_t L);
int carel_pair_preds(const float* x, int64_t ldx, int64_t N, int64_t D, int64_t L, const float* const* w4,
                     const float* const* b4, const float* eps_e, const float* eps_c, const float* pair_w,
                     const float* pair_b, float* out, void* ws, size_t ws_bytes, int precision, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32 precision tier (north_star: "1e-5 in fp32"): the same operators with fp32 FFMA arithmetic on the CUDA cores
 * (csrc/sgemm.cuh) on materialised intermediates, as the reference's fp32 ATen path computes them.  Selected by
 * carel_vae_b200.set_precision("fp32") / precision="fp32"; the bf16 tier above is the product path.
 * carel_mmd_f32_bwd is also the gradient of the bf16 tier for feature widths its fused backward cannot hold
 * (dim > 250, e.g. the d = 768 Gram): MMDStatistic is differentiable at any width, like the reference's
 * (drl_classifier_ec_mmd_final_mul.py:547-596).
 * ------------------------------------------------------------------------------------------ */
/* C[M,N] (op)= alpha opA(A) opB(B) (+ bias[N]); ta: A is [K,M]; tb: B is [K,N] (else [N,K], nn.Linear);
 * mode 0 store, 1 add, 2 atomic add (split-K, ksplit <= 0: automatic). */
int carel_f32_gemm(const float* a, int64_t lda, int ta, const float* b, int64_t ldb, int tb, int64_t M, int64_t N,
                   int64_t K, float* c, int64_t ldc, const float* bias, float alpha, int mode, int ksplit, void* stream);
/* latent layer (same contract as carel_latent_fwd / carel_latent_bwd; the backward takes x and the weights again
 * instead of packed copies in the workspace) */
size_t carel_latent_f32_workspace_bytes(int64_t N, int64_t D, int64_t L);
int carel_latent_f32_fwd(const float* x, int64_t ldx, int64_t N, int64_t D, int64_t L, const float* const* w4,
                         const float* const* b4, const float* eps_e, const float* eps_c, float* pre, float* z, float* kl,
                         void* ws, size_t ws_bytes, void* stream);
int carel_latent_f32_bwd(const float* grad_z, int64_t ldgz, const float* grad_kl, const float* x, int64_t ldx,
                         const float* const* w4, const float* pre, const float* eps_e, const float* eps_c, int64_t N,
                         int64_t D, int64_t L, float* grad_x, int64_t ldgx, float* const* grad_w4, float* const* grad_b4,
                         void* ws, size_t ws_bytes, void* stream);
/* MMD on materialised row blocks (<= 256 MB each) of the Gram, any feature width; row-tile ranges as in carel_mmd_fwd */
size_t carel_mmd_f32_workspace_bytes(int64_t n1, int64_t n2, int64_t dim);
int carel_mmd_f32_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t n1, int64_t n2, int64_t dim,
                      const float* alphas_host, int n_alpha, float eps, int64_t row_tile_begin, int64_t row_tile_end,
                      double* partial_sums, float* out_mmd, void* ws, size_t ws_bytes, void* stream);
int carel_mmd_f32_bwd(const float* grad_out, const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t n1,
                      int64_t n2, int64_t dim, const float* alphas_host, int n_alpha, float eps, int64_t row_tile_begin,
                      int64_t row_tile_end, float* grad_x, int64_t ldgx, float* grad_y, int64_t ldgy, void* ws,
                      size_t ws_bytes, void* stream);
/* HSIC (drl_classifier_ec_hsic.py:529-547) with both m x m kernel matrices materialised (m <= 46340); the backward
 * must follow the forward on the same workspace */
size_t carel_hsic_f32_workspace_bytes(int64_t m, int64_t dx, int64_t dy);
int carel_hsic_f32_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t m, int64_t dx, int64_t dy,
                       float s_x, float s_y, float* out, void* ws, size_t ws_bytes, void* stream);
int carel_hsic_f32_bwd(const float* grad_out, const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t m,
                       int64_t dx, int64_t dy, float s_x, float s_y, float* grad_x, int64_t ldgx, float* grad_y,
                       int64_t ldgy, void* ws, size_t ws_bytes, void* stream);
/* BoW decoder + loss with the logits [N, V] materialised (any K); the backward must follow the forward on the same
 * workspace.  need_dense_target = 1 when only an ELL triple is given (it is expanded into the workspace). */
size_t carel_bow_f32_workspace_bytes(int64_t N, int64_t K, int64_t V, int need_dense_target);
int carel_bow_f32_fwd(const float* z, int64_t ldz, const float* dec_w, const float* dec_b, const float* bow_dense,
                      int64_t ld_bow, const int32_t* ell_cols, const float* ell_vals, const int32_t* ell_nnz, int64_t N,
                      int64_t K, int64_t V, int64_t ell_cap, float label_smoothing, float* loss_out, void* ws,
                      size_t ws_bytes, void* stream);
int carel_bow_f32_bwd(const float* grad_out, const float* z, int64_t ldz, const float* dec_w, const float* bow_dense,
                      int64_t ld_bow, int64_t N, int64_t K, int64_t V, float label_smoothing, float* grad_z, int64_t ldgz,
                      float* grad_w, float* grad_b, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAREL_HEAD_H_ */
