/* mggan_b200.h -- C ABI of the B200 (sm_100a) MG-GAN training-step kernels.
 *
 * The reference (selflein/MG-GAN) has no FFI: its hot path is Python nn.Modules calling
 * PyTorch library kernels.  This header is the boundary a maintainer binds
 * instead: each entry point replaces the reference call sites cited above it.  Conventions:
 *   - plain pointers to DEVICE memory (fp32 unless noted), explicit sizes, a cudaStream_t;
 *   - no allocation, no ownership transfer, no stream synchronisation inside;
 *   - every function returns 0 on success, non-zero on error (1 invalid argument / unsupported
 *     shape, 2 CUDA launch error, 3 wrong device); mggan_last_error() gives the message
 *     (thread-local);
 *   - "accumulated" outputs are added to atomically: the caller zero-fills them first.
 * Row-major everywhere; weights use the PyTorch (out_features, in_features) layout.
 */
#ifndef MGGAN_B200_H
#define MGGAN_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* mggan_last_error(void);
int mggan_version(void);
int mggan_device_check(void);

/* ---- trajectory-encoder LSTM: TrajectoryEncoder.forward, mggan/model/modules/common_modules.py:48-66
 * x (T,N,2); Wx (4H,2) = W_ih W_emb; b (4H) = W_ih b_emb + b_ih + b_hh; Whh (4H,H); H in {32,64}.
 * hT (N,H).  acts (T,N,6,H) saves (i,f,g,o,c,tanh c) for the backward, or NULL. */
int mggan_lstm_seq_fwd(const float* x, int T, int N, int H, const float* Wx, const float* b, const float* Whh,
                       float* hT, float* acts, cudaStream_t stream);
/* dWx (4H,2), db (4H), dWhh (4H,H) accumulated. */
int mggan_lstm_seq_bwd(const float* x, int T, int N, int H, const float* Whh, const float* acts, const float* dhT,
                       float* dWx, float* db, float* dWhh, cudaStream_t stream);

/* ---- dense layer: nn.Linear (+ activation) call sites standard.py:91-105, discriminators.py:46-56,76-108
 * act: 0 none, 1 ReLU, 2 LeakyReLU(slope), 3 sigmoid squashed to (1e-7, 1-1e-7) (discriminators.py:203-204). */
int mggan_linear_fwd(const float* X, int M, int K, const float* W, const float* bias, int O, int act, float slope,
                     float* Y, cudaStream_t stream);
/* dX (M,K) overwritten or NULL; dW (O,K), db (O) accumulated or NULL.  Y is the forward output. */
int mggan_linear_bwd(const float* X, int M, int K, const float* W, int O, int act, float slope, const float* Y,
                     const float* dY, float* dX, float* dW, float* db, cudaStream_t stream);

/* GEMM kernel behind mggan_linear_*: 1 = FP32 64 x 64 tile (default), 2 = FP32 128 x 64 tile with register prefetch,
 * 3 = tcgen05 tensor cores, 3 x TF32 with fp32-level accuracy (2 and 3 are opt-in until measured).  Process-wide switch;
 * returns the previous variant, -1 for an unknown one. */
int mggan_set_gemm_variant(int variant);

/* ---- discriminator heads over k samples per agent, per-agent part hoisted: MultiDiscriminatorTrajectory.forward
 * mggan/model/modules/discriminators.py:178-219 (discs[0] :76-85,198-204; gen_id_reconstructor :103-108,209-217).
 * The classifier input of row (sample s, agent i) is [soc | in_enc | pred_enc | scene]: only pred_enc depends on s and
 * soc is non-zero for s == 0 only.  First layers of both heads stacked (NZ = 2 HH rows, HH without
 * the generator-id head): z[s,i] = base[i] + (s == 0) soc0[i] + W1p pe[s,i];  pe (k*n, 32) row = s*n + i;
 * base, soc0 (n, NZ) per-agent products computed by the caller (bias in base); W1p (NZ, 32); HH in {64, 96}.
 * p (k*n) = sigmoid(Wd2 . LReLU_0.2(z[:HH]) + bd2) (1 - 2e-7) + 1e-7;  branch (k*n, G) = Wg2 LReLU_0.2(z[HH:]) + bg2
 * (G == 0 and branch == NULL for gan_type "gan"). */
int mggan_disc_heads_fwd(const float* pe, int n, int k, int HH, const float* base, const float* soc0, const float* W1p,
                         const float* Wd2, const float* bd2, const float* Wg2, const float* bg2, int G, float* p,
                         float* branch, cudaStream_t stream);
/* Input gradients only (generator step: the discriminator is frozen).  dp (k*n) = dL/dp or NULL, dbranch (k*n, G) or
 * NULL.  d_pe (k*n, 32) and d_soc0 (n, NZ) overwritten; d_base (n, NZ) accumulated (caller zero-fills) or NULL. */
int mggan_disc_heads_bwd(const float* pe, int n, int k, int HH, const float* base, const float* soc0, const float* W1p,
                         const float* Wd2, const float* Wg2, int G, const float* p, const float* dp,
                         const float* dbranch, float* d_pe, float* d_soc0, float* d_base, cudaStream_t stream);

/* ---- PM-Net sampling + generator selection: standard.py:217-225, utils.py:234-248, standard.py:190-214 */
/* dyn_offset: NULL, or a DEVICE counter added to `offset` (CUDA-graph replay: a new Philox offset per replay). */
int mggan_gumbel_sample(const float* logits, int n, int k, int G, unsigned long long seed, unsigned long long offset,
                        const unsigned long long* dyn_offset, long long* idx, cudaStream_t stream);
int mggan_selection_tiles(int n_seq, int G); /* host helper: tile-table length for n_seq sequences */
/* idx (n,k) int64 -> work list for the decoder.  scratch: cnt (n*G int32), rank (n*k bytes),
 * base_row (G+1 int32), err (1 int32, zero-filled by the caller; set to 1 on an out-of-range index).
 * outputs: totals (G) per-generator draw counts, tile_gen (n_tiles), seq_* (n_tiles*64). */
int mggan_selection_build(const long long* idx, int n, int k, int G, int n_tiles, int* cnt, unsigned char* rank,
                          int* base_row, int* err, int* totals, int* tile_gen, int* seq_agent, int* seq_noise,
                          int* seq_out, cudaStream_t stream);
/* all generators x k samples (forward_all, standard.py:227-265): G * 2*ceil(n*k/128) tiles.  Both builders pad every
 * generator's rows to a multiple of 128 (two 64-row tiles), so tiles 2s and 2s+1 always belong to one generator. */
int mggan_selection_all(int n, int k, int G, int* tile_gen, int* seq_agent, int* seq_noise, int* seq_out,
                        cudaStream_t stream);

/* ---- multi-generator decoder: RelativeDecoder.forward common_modules.py:97-131 x forward_all standard.py:227-265
 * A (n_agents,32) = enc_h_to_dec_h on the per-agent encoding (bias included), social (n_agents,32),
 * last_xy/last_dxdy (n_agents,2), noise (rows,Z); Wz (32,Z) noise block of enc_h_to_dec_h;
 * per generator: Wx (G,128,2), b (G,128), Whh (G,128,32), W1h/W1s (G,16,32) = halves of hidden2pos.0,
 * b1 (G,16), W2 (G,2,16), b2 (G,2).  out_abs/out_rel (pred_len, n_cols, 2).
 * Saved for the backward (all three or all NULL), n_tiles even, rows grouped in 128-row tiles with the row index second-fastest
 * (a warp's rows are contiguous): acts (pred_len, n_tiles/2, 16 unit-pairs, 128 rows, 4) = the recurrent state (h, c | h, c) of the
 * pair's two units after the step -- the backward recomputes the gates from h_{t-1} (one more tensor-pipe product per step),
 * u1save (pred_len, n_tiles/2, 4, 128 rows, 4), h0save (n_tiles/2, 8, 128 rows, 4)  -- csrc/common.cuh dec_*_off. */
int mggan_decoder_fwd(int n_tiles, const int* tile_gen, const int* seq_agent, const int* seq_noise,
                      const int* seq_out, const float* A, const float* social, const float* last_xy,
                      const float* last_dxdy, const float* noise, int Z, const float* Wz, const float* Wx,
                      const float* b, const float* Whh, const float* W1h, const float* W1s, const float* b1,
                      const float* W2, const float* b2, int pred_len, int n_cols, float* out_abs, float* out_rel,
                      float* acts, float* u1save, float* h0save, cudaStream_t stream);
/* Same operator, arguments and outputs as mggan_decoder_fwd, with the per-step gate contraction W_hh h_{t-1}
 * ((128 rows x 32) . (32 x 128)) on the tcgen05 tensor cores: kind::tf32 with a hi/lo operand split (3 products,
 * fp32-level accuracy), accumulator in TMEM, one 128-row tile (tiles 2s, 2s+1) per CTA.  n_tiles must be even. */
int mggan_decoder_fwd_tc(int n_tiles, const int* tile_gen, const int* seq_agent, const int* seq_noise,
                         const int* seq_out, const float* A, const float* social, const float* last_xy,
                         const float* last_dxdy, const float* noise, int Z, const float* Wz, const float* Wx,
                         const float* b, const float* Whh, const float* W1h, const float* W1s, const float* b1,
                         const float* W2, const float* b2, int pred_len, int n_cols, float* out_abs, float* out_rel,
                         float* acts, float* u1save, float* h0save, cudaStream_t stream);
/* Test hook for the tensor-core plumbing: D (128,128) = A (128,32) . B (128,32)^T by the same 3 x TF32 path. */
int mggan_tc_selftest(const float* A, const float* B, float* D, cudaStream_t stream);
/* d_abs / d_rel (pred_len, n_cols, 2), either may be NULL.  All gradient outputs accumulated. */
int mggan_decoder_bwd(int n_tiles, const int* tile_gen, const int* seq_agent, const int* seq_noise,
                      const int* seq_out, const float* social, const float* last_dxdy, const float* noise, int Z,
                      const float* Wz, const float* Wx, const float* b, const float* Whh, const float* W1h,
                      const float* W1s, const float* b1, const float* W2, const float* b2, int pred_len, int n_cols,
                      const float* out_rel, const float* acts, const float* u1save, const float* h0save,
                      const float* d_abs, const float* d_rel, float* dWz, float* dWx, float* db, float* dWhh,
                      float* dW1h, float* dW1s, float* db1, float* dW2, float* db2, float* dA, float* dsocial,
                      cudaStream_t stream);

/* ---- social attention: SocialAttention / AttentionPooling, mggan/model/modules/social.py:7-123
 * x4 (N,4) = (xy_T, dxdy_T); h (N,HD), HD in {32,64}; Us (N,65) = per-agent folded (u_j, s_j);
 * scene_off (S+1) agent ranges, pair_off (S+1) prefix of n_s^2; W1 (32,3), b1, W2 (64,32), b2.
 * S (N,HD); att (sum n_s^2) attention weights (saved for the backward). */
int mggan_social_attn_fwd(const float* x4, const float* h, int HD, const float* Us, const int* scene_off,
                          const int* pair_off, int n_scenes, const float* W1, const float* b1, const float* W2,
                          const float* b2, float* S, float* att, cudaStream_t stream);
/* dsig (sum n_s^2) scratch; dh (N,HD), dUs (N,65) zero-filled by the caller; dW1, db1, dW2, db2 accumulated. */
int mggan_social_attn_bwd(const float* x4, const float* h, int HD, const float* Us, const int* scene_off,
                          const int* pair_off, int n_scenes, const float* W1, const float* b1, const float* W2,
                          const float* b2, const float* att, const float* dS, float* dsig, float* dh, float* dUs,
                          float* dW1, float* db1, float* dW2, float* db2, cudaStream_t stream);

/* ---- physical attention: AttentionGlobal / CNN / Conv_Blocks, mggan/model/modules/cnn.py:101-116,119-282
 * img (.,4,33,33); rows (N) optional int32 gather of image rows (mask compaction) or NULL; C in {8,16}.
 * conv1 is linear in the crop, so train-mode BatchNorm-1 statistics and the dense half of conv1's weight
 * gradient come from data-only patch statistics R (36x36 doubles, symmetric), P (36 doubles), accumulated. */
int mggan_scene_patch_stats(const float* img, const int* rows, int N, double* R, double* P, cudaStream_t stream);
/* count = conv1 outputs per channel (N_total * 33 * 33).  training != 0: batch statistics + running-stat
 * update (momentum, unbiased variance) + num_batches_tracked += 1; else running statistics.
 * ab (2C) = BN as y = a x + b; mean_istd (2C). */
int mggan_scene_bn1_from_patches(const double* R, const double* P, double count, int C, const float* W,
                                 const float* bias, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, long long* num_batches_tracked, float momentum, float eps,
                                 int training, float* ab, float* mean_istd, cudaStream_t stream);
/* conv1 -> BN1 -> ReLU -> pool -> conv2.  x2 (N,C,16,16) pre-BatchNorm-2 output; stats2 (2C doubles: sum, sum of
 * squares) accumulated or NULL (eval); e1 (N,C,256) pre-BN value at the pool arg + idx1 (N,C,256 bytes: arg |
 * active<<2) saved for the backward, or both NULL. */
int mggan_scene_fused12_fwd(const float* img, const int* rows, int N, int C, const float* W1, const float* b1,
                            const float* ab1, const float* W2, const float* b2, float* x2, double* stats2, float* e1,
                            unsigned char* idx1, cudaStream_t stream);
/* BatchNorm-2 finalize (same contract as above, from per-channel sums). */
int mggan_scene_bn_finalize(const double* stats, double count, int C, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                            float eps, int training, float* ab, float* mean_istd, cudaStream_t stream);
/* Wa1 (32,C), ba1 (32), Wa2 (C,32), ba2 (C) = cnn_attention; out (N,64). */
int mggan_scene_attn_fwd(const float* x2, int N, int C, const float* ab2, const float* Wa1, const float* ba1,
                         const float* Wa2, const float* ba2, float* out, cudaStream_t stream);
/* dy2 (N,C,64) + idx2 (N,C,64 bytes): sparse gradient at BatchNorm-2 output; sums2 (2C doubles) accumulated. */
int mggan_scene_attn_bwd(const float* x2, int N, int C, const float* ab2, const float* mean_istd2, const float* Wa1,
                         const float* ba1, const float* Wa2, const float* ba2, const float* dout, float* dWa1,
                         float* dba1, float* dWa2, float* dba2, float* dy2, unsigned char* idx2, double* sums2,
                         cudaStream_t stream);
/* sums (2C doubles) -> m12 (2C) means for the BatchNorm backward; dgamma, dbeta (C) accumulated. */
int mggan_scene_bn_bwd_finalize(const double* sums, double count, int C, float* m12, float* dgamma, float* dbeta,
                                cudaStream_t stream);
/* dW2 (C,C,3,3), dbias2 (C), S1 (C,36) = sum dy1 * patch at the pool-arg positions, sums1 (2C doubles): accumulated. */
int mggan_scene_fused12_bwd(const float* img, const int* rows, int N, int C, const float* x2, const float* e1,
                            const unsigned char* idx1, const float* ab1, const float* mean_istd1, const float* ab2,
                            const float* mean_istd2, const float* m12_2, const float* W2, const float* dy2,
                            const unsigned char* idx2, float* dW2, float* dbias2, float* S1, double* sums1,
                            cudaStream_t stream);
/* closed-form BatchNorm-1 backward: dW1 (C,4,3,3) overwritten, dgamma/dbeta (C) overwritten with the local sums.
 * sums_global: all-reduced sums1 when data-parallel (same pointer as sums_local otherwise); R, P: local shard. */
int mggan_scene_bn1_bwd_finalize(const double* sums_global, const double* sums_local, double count, int C,
                                 const float* S1, const double* R, const double* P, const float* W, const float* bias,
                                 const float* ab1, const float* mean_istd1, float* dW, float* dgamma, float* dbeta,
                                 cudaStream_t stream);

/* ---- losses: mggan/model/train.py:55-113 (G step), :148-200 (D step), :626-639 (PM step) */
/* abs (T,k,n,2), gt (T,n,2); loss += sum_scenes min_s sum_{i in scene} sum_t |abs-gt| * inv_norm  (squared != 0:
 * |abs-gt|^2, l2_loss_type "mse", train.py:62-63); best (S) argmin sample; d_abs (T,k,n,2) zero-filled by the caller, or NULL. */
int mggan_l2_scene_min(const float* abs_, const float* gt, int T, int k, int n, const int* scene_off, int n_scenes,
                       float inv_norm, int squared, float* loss, int* best, float* d_abs, cudaStream_t stream);
/* loss += inv_denom * sum_i w_i BCE(p_i, label), w_i = 1/counts[gen_idx[i]] or 1; dp (n) or NULL.
 * (gan_obj NS; gan_obj MM's generator term -BCE(d_fake, l_fake) is the same call with a negative inv_denom.) */
int mggan_bce_scalar_label(const float* p, int n, float label, const long long* gen_idx, const int* counts,
                           float inv_denom, float* loss, float* dp, cudaStream_t stream);
/* gan_obj LS (abstract_train.py:72-75): loss += inv_denom * sum_i w_i (p_i - label)^2; dp (n) or NULL. */
int mggan_mse_scalar_label(const float* p, int n, float label, const long long* gen_idx, const int* counts,
                           float inv_denom, float* loss, float* dp, cudaStream_t stream);
int mggan_ce_generators(const float* logits, int n, int G, const long long* target, const int* counts,
                        float inv_denom, float* loss, float* dlogits, cudaStream_t stream);
int mggan_pm_ml_loss(const float* abs_all, const float* gt, int T, int ks, int G, int n, const float* logits,
                     float sigma, float weight, float inv_n, float* loss, float* dlogits, float* target_out,
                     cudaStream_t stream);

/* ---- crop features cut on the device from scene images resident in HBM: BaseDataset.ImageFeatures_small,
 * mggan/data_utils/BaseTrajectories.py:254-288 (called per agent from trajectories_scene.py:343-351).
 * atlas: the u8 RGB pixels (H, W, 3) of every scene's `small_image`, back to back; img_off (n_images) int64 byte offsets;
 * img_wh (n_images, 2) int32 (width, height); img_scale (n_images) = float32(1 / scaling_small) (1 for format "pixel");
 * agent_img (N) int32 image of each agent (out of range: no image, RGB channels = -1); last_xy (N, 2) = in_xy[-1].
 * features (N, 4, 33, 33): channels 0-2 = -1 + u8 * 2 / 256 of the box [c - 16, c + 17) around the centre pixel
 * c = int(last_xy * scale) (zeros outside the image, like PIL's crop), channel 3 one-hot at [16, 16].  Bit-exact. */
int mggan_scene_crop(const unsigned char* atlas, const long long* img_off, const int* img_wh, const float* img_scale,
                     int n_images, const int* agent_img, const float* last_xy, int N, float* features,
                     cudaStream_t stream);

/* ---- evaluation metrics (scripts/evaluate.py sweep): mggan/manifold.py:9-18,70-77 via evaluation.py:101-156, and
 * mggan/metrics.py:6-20,99-141 via evaluation.py:43-78. */
/* traj (P, T, 2) pool of trajectories; radius (T) DOUBLES = linspace(r / T, r, T); desc (n_tests, 3) int32 =
 * (test trajectory, first entry of its manifold in man_list, number of entries); man_list: int32 trajectory indices.
 * inside (n_tests) bytes: 1 when at every step the test is closer than radius[t] to some manifold sample (float32
 * distances exactly as numpy computes them; an empty manifold gives 0). */
int mggan_tube_inside(const float* traj, int T, const double* radius, const int* desc, int n_tests,
                      const int* man_list, unsigned char* inside, cudaStream_t stream);
/* preds (T, K, n, 2), gt (T, n, 2) without masked agents, scene_off (S+1) int32, scene_scale (S) or NULL (pixel datasets
 * scale coordinates by 1 / ratio).  Row sc of ade / fde (S, K) doubles: entry kk-1 = min over the first kk samples of the
 * scene's summed displacement error (all steps / final step); mode (S, K) int32: agents whose best final displacement
 * among the first kk samples is < mode_thresh.  K <= 64. */
int mggan_min_ade_fde(const float* preds, const float* gt, int T, int K, int n, const int* scene_off, int n_scenes,
                      const float* scene_scale, float mode_thresh, double* ade, double* fde, int* mode,
                      cudaStream_t stream);

/* ---- optimiser: clip_grad_norm_ + AdamW, train.py:131-135,209-213,656-658; abstract_train.py:45-57 */
#define MGGAN_TABLE_MAX 64
typedef struct MgganTensorTable {
    float* p[MGGAN_TABLE_MAX];
    const float* g[MGGAN_TABLE_MAX];
    float* m[MGGAN_TABLE_MAX];
    float* v[MGGAN_TABLE_MAX];
    int n[MGGAN_TABLE_MAX];
    float bc1[MGGAN_TABLE_MAX];      /* 1 - beta1^step */
    float bc2_sqrt[MGGAN_TABLE_MAX]; /* sqrt(1 - beta2^step) */
    const float* dyn;                /* NULL, or DEVICE array [lr, bc1[64], bc2_sqrt[64]] overriding lr / bc1 / bc2_sqrt
                                        (CUDA-graph replay: values refreshed by the host between replays) */
} MgganTensorTable;
/* table is a HOST pointer (copied into kernel-parameter space). */
int mggan_grad_sqnorm(const MgganTensorTable* table, int count, double* sqnorm_accum, cudaStream_t stream);
int mggan_clip_adamw(const MgganTensorTable* table, int count, const double* sqnorm, float max_norm,
                     float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                     cudaStream_t stream);
int mggan_multi_copy(const MgganTensorTable* table, int count, cudaStream_t stream); /* p[t] <- g[t] */

/* ---- fused two-layer perceptron Y = act2(W2 act1(W1 x + b1) + b2): the dense chains of the path in one launch each
 * (discriminators.py:46-56 in_encoder_fc / pred_encoder, :76-108 the two heads; standard.py:99-105 PM-Network layers 1-2).
 * X (M,K) row-major 16-byte aligned, K % 4 == 0, K <= 192; W1 (H,K), H % 4 == 0, H <= 96; W2 (O,H), O <= 32; biases may be
 * NULL; act codes as mggan_linear_* (0 none, 1 ReLU, 2 LeakyReLU(slope), 3 sigmoid * (1 - 2e-7) + 1e-7).  The hidden layer
 * stays in shared memory (forward) / is recomputed from X (backward).  Backward: dX (M,K) overwritten or NULL; dW1, db1, dW2,
 * db2 ACCUMULATED into (caller zero-fills), all set or all NULL (frozen weights -> input gradient only). */
int mggan_mlp2_fwd(const float* X, long long M, int K, const float* W1, const float* b1, int H, int act1, float slope1,
                   const float* W2, const float* b2, int O, int act2, float slope2, float* Y, cudaStream_t stream);
int mggan_mlp2_bwd(const float* X, long long M, int K, const float* W1, const float* b1, int H, int act1, float slope1,
                   const float* W2, int O, int act2, float slope2, const float* Y, const float* dY, float* dX,
                   float* dW1, float* db1, float* dW2, float* db2, cudaStream_t stream);

/* ---- data-parallel exchange (new; the reference is single-device): one-shot all-reduce over
 * NVLink peer memory, fused with the squared gradient norm that mggan_clip_adamw clips with.  `region[r]` is THIS call's
 * region of rank r's symmetric arena and `flags[r]` rank r's flag array ([64][16] uint32, zero-initialised once), both as
 * peer-mapped device pointers (torch.distributed._symmetric_memory buffer_ptrs + offsets).  The kernel copies `in` (n
 * elements; NULL when a packing kernel already wrote the operand into region[rank]) into its own region, meets the same
 * block of every rank at a flag barrier, and writes out[i] = sum_r region[r][i] in fixed rank order (bit-identical on every
 * rank); for float32 it also adds sum_i out[i]^2 to *sqnorm when sqnorm != NULL.  dtype: 0 float32, 1 float64, 2 int32.
 * Every rank must issue the same sequence of calls; a region may be reused only after another call of the sequence. */
#define MGGAN_PEER_MAX 16
typedef struct MgganPeerTable {
    void* region[MGGAN_PEER_MAX];
    unsigned int* flags[MGGAN_PEER_MAX];
    int rank, world;
} MgganPeerTable;
/* table is a HOST pointer (copied into kernel-parameter space). */
int mggan_peer_allreduce(const MgganPeerTable* table, int dtype, const void* in, long long n, void* out, double* sqnorm,
                         cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MGGAN_B200_H */
