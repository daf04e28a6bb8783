/* deco_b200 -- C ABI of the B200 (sm_100a) kernels behind the DeCo denoise / sample / DCT-loss hot path.
 *
 * The reference (hhhhzp/DeCo) is pure Python/PyTorch: it has no FFI or operator registry, its "plugin boundary" is
 * class substitution in the YAML configs.  This header is therefore the boundary the *new* Python
 * modules (deco_b200/*.py, which mirror the reference classes) bind with ctypes; each entry point cites the
 * reference call site(s) it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory unless stated; no allocation inside the library
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); all calls are asynchronous
 *   - return 0 on success, a cudaError_t (>0) or a negative DECO_ERR_* code; deco_last_error() gives the message
 *   - bf16 tensors are passed as void*; "ld*" / "*_stride" are in ELEMENTS
 *   - thread-compatible: no global mutable state besides lazily initialised, idempotent function attributes
 */
#ifndef DECO_B200_H
#define DECO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DECO_ERR_ARG (-1)
#define DECO_ERR_UNSUPPORTED (-2)
#define DECO_ERR_DRIVER (-3)

const char* deco_last_error(void);
int deco_abi_version(void);

/* GEMM epilogues */
#define DECO_EPI_BIAS 0           /* out = A.W^T + bias                                  (nn.Linear)              */
#define DECO_EPI_BIAS_SILU 1      /* out = silu(A.W^T + bias)                            (t_embedder.mlp[0:2])    */
#define DECO_EPI_GATE_RESIDUAL 2  /* out = resid + gate[row / rows_per_gate] * (A.W^T + bias); resid, out FP32      */
#define DECO_EPI_SWIGLU 3         /* W rows interleaved [16 x w1 | 16 x w3]: out[:, j] = silu(a_j) * b_j, N/2 cols */
#define DECO_EPI_BIAS_F32 4       /* out = A.W^T + bias, FP32 output (head of the fp32 residual stream)           */
/* training step: the SwiGLU passes of mlp (dit_c2i_DeCo.py:113) inside the GEMMs next to them; `resid` / `ldr` carry the
 * bf16 pre-activation matrix y13 = [16 x w1 | 16 x w3]-interleaved instead of the fp32 residual */
#define DECO_EPI_SWIGLU_DUAL 5    /* DECO_EPI_SWIGLU, and y13 = bf16(A.W^T) [M, N] is WRITTEN to resid (saved for backward) */
#define DECO_EPI_SWIGLU_BWD 6     /* du = A.W^T [M, N]; y13 [M, 2N] READ from resid; out = dy13 bf16 [M, 2N]             */

/* tcgen05 / TMEM / TMA bf16 GEMM: out[M, N(or N/2)] = epilogue(A[M,K] . W[N,K]^T), fp32 accumulate.
 * Replaces nn.Linear at src/models/transformer/dit_c2i_DeCo.py:496 (s_embedder), :55-57 (t_embedder.mlp),
 * :207 (adaLN_modulation, all blocks in one call), :176 (attn.qkv), :188+:208 (attn.proj + gated residual),
 * :113 (mlp.w1/w3/w2, SwiGLU), :209 (gated residual), :404 (dec_net.cond_embed).
 * K, N, lda, ldw, ldo multiples of 8; pointers 16-byte aligned; tile_n in {0 (auto), 128, 192, 256}. */
int deco_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                   int M, int N, int K, int epilogue, const float* bias,
                   const void* resid, long long ldr, const void* gate, long long gate_stride, int rows_per_gate,
                   int tile_n, void* stream);

/* Host-side view of the launch deco_gemm_bf16 would make (no GPU needed): the tile width it picks for M x N x K on `ctas`
 * CTAs (0 = this device's SM count) and, walking the kernels' own static tile schedule, the heaviest CTA group's work in
 * 1/256ths of a full-width tile plus the number of tiles not visited exactly once (0 for a correct schedule). */
int deco_gemm_tile_plan(int M, int N, int K, int ctas, int* tile_n_out, int* max_load_256ths, int* bad_tiles);

/* ---- GEMMs whose epilogues absorb the memory-bound neighbours of a DiT block (csrc/gemm_fused.cu) ----
 * RMSNorm commutes with the following Linear because its row factor is a scalar:
 *   Linear(modulate(RMSNorm(x))) = rstd[row] * ((x * w_norm * (1 + scale)) . W^T) + (shift . W^T)[image]
 * so the producer of x emits xg = bf16(x * w_norm * (1 + scale)) and per-row partial sums of squares, the consumer GEMM
 * applies rstd and adds the (tiny, separately computed) shift product.  Replaces, per FlattenDiTBlock.forward
 * (dit_c2i_DeCo.py:206-210), the stand-alone RMSNorm/modulate passes (:94-99, :11-12) and q/k-norm + RoPE (:178-180). */

/* Residual-stream update (dit_c2i_DeCo.py:208/:209, also :496 with resid = gate = NULL):
 *   out = [resid + gate[row / rows_per_image] *] (A.W^T + bias)          fp32 [M, N], may alias resid
 *   ssq_out[p][row] = sum over the p-th column tile of out[row]^2        fp32 [deco_gemm_stream_parts(N, K)][M] or NULL
 *   xg_out = bf16(out * next_norm_w * (1 + next_scale[row / rows_per_image]))   or nothing when next_norm_w = NULL
 * N % 32 == 0; gate / next_scale are bf16 [M / rows_per_image, stride]. */
int deco_gemm_stream_parts(int N, int K);
int deco_gemm_stream(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                     const float* bias, const float* resid, long long ldr, float* out, long long ldo,
                     const void* gate, long long gate_stride, int rows_per_image,
                     const float* next_norm_w, const void* next_scale, long long next_scale_stride,
                     void* xg_out, long long ldx, float* ssq_out, void* stream);

/* QKV projection of a normalised + modulated stream with q_norm / k_norm / RoPE fused (dit_c2i_DeCo.py:176-180;
 * dit_t2i_pixnerd.py:43-50, :170-173):  y = rstd[row] * (A.W^T) + shw[row / rows_per_image], rstd from ssq_in (NULL: 1,
 * shw NULL: 0).  N = 1..3 segments of heads*head_dim columns; segment i gets a per-head RMSNorm with weight w_seg<i>
 * (NULL: passed through) and RoPE when bit i of rope_mask is set (table fp32 [rows_per_image, head_dim/2, 2]).
 * rope_tokens_per_row > 0 declares the table AXIAL as built by precompute_freqs_cis_2d / _ex2d (dit_c2i_DeCo.py:116-131,
 * layers/rope.py:22-37): even pairs depend on tok % tokens_per_row only, odd pairs on tok / tokens_per_row only, so the
 * kernel keeps just those rows in shared memory; 0 = arbitrary table, read from global memory. bf16 out. */
int deco_gemm_norm_qkv(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                       int M, int N, int K, int rows_per_image,
                       const float* ssq_in, int ssq_parts, int norm_hidden, float norm_eps,
                       const float* shw, long long shw_stride,
                       int heads, int head_dim, const float* w_seg0, const float* w_seg1, const float* w_seg2,
                       int rope_mask, const float* rope_cos_sin, int rope_tokens_per_row, float head_eps,
                       int out_head_pitch /* output columns per head: 0 = head_dim (dense); 80 for head_dim 72 (zero padded) */,
                       void* stream);

/* SwiGLU up-projection of a normalised + modulated stream (dit_c2i_DeCo.py:113 on :209's modulate input): W rows
 * interleaved [16 x w1 | 16 x w3] as for DECO_EPI_SWIGLU; out bf16 [M, N/2]. */
int deco_gemm_norm_swiglu(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                          int M, int N, int K, int rows_per_image,
                          const float* ssq_in, int ssq_parts, int norm_hidden, float norm_eps,
                          const float* shw, long long shw_stride, void* stream);

/* Tuning knob (process-wide, for A/B measurements): cta_group -1 = auto (2-CTA pairs when M > 128), 1, 2;
 * staged_epilogue -1 = auto (staged with 2-CTA), 0 = direct row-per-thread stores, 1 = shared-memory staged. */
int deco_gemm_set_tuning(int cta_group, int staged_epilogue);
/* Process-wide: the persistent GEMM kernels size their grids for (SM count - n) SMs (n even, 0..64; default 0), leaving
 * room for kernels on other streams -- NCCL's, while gradients are averaged during the backward. */
int deco_gemm_reserve_sms(int n);

/* F.unfold + transpose (dit_c2i_DeCo.py:491): fp32 [B,C,H,W] -> bf16 [B*L, C*p*p], feature order c*p*p + ky*p + kx */
int deco_patchify(const float* x, void* out_bf16, int B, int C, int H, int W, int p, void* stream);

/* TimestepEmbedder.timestep_embedding (dit_c2i_DeCo.py:43-53): [B] -> bf16 [B, dim] = [cos || sin] */
int deco_timestep_freq(const float* t, void* out_bf16, int B, int dim, float max_period, void* stream);

/* c = silu(t_emb + embedding_table[y]) (dit_c2i_DeCo.py:493-494); labels int64, table fp32 [num_rows, hidden] */
int deco_cond_combine(const void* temb_bf16, const float* table, const long long* labels, void* c_bf16,
                      int B, int hidden, int num_rows, void* stream);

/* RMSNorm (dit_c2i_DeCo.py:94-99) + modulate (:11-12): out = w * rms(x) * (1 + scale) + shift.
 * shift/scale point into the batched adaLN output; row m uses modulation row m / rows_per_mod.
 * x is the residual stream: fp32 (x_is_f32, the product path) or bf16 (the reference's rounding points). */
int deco_rmsnorm_modulate(const void* x, int x_is_f32, const float* weight, const void* shift_bf16,
                          const void* scale_bf16, long long mod_row_stride, int rows_per_mod, void* out_bf16,
                          long long M, int hidden, float eps, void* stream);

/* q_norm / k_norm + apply_rotary_emb (dit_c2i_DeCo.py:178-180, :134-145), in place on the QKV GEMM output
 * [M, 3*heads*head_dim]; rope_cos_sin: fp32 [L, head_dim/2, 2]; token position = row % L. head_dim in {64, 72}. */
int deco_qknorm_rope(void* qkv_bf16, const float* q_weight, const float* k_weight, const float* rope_cos_sin,
                     long long M, int heads, int head_dim, int L, float eps, void* stream);

/* General form of the above for the t2i layouts (dit_t2i_pixnerd.py:43-50, :170-173): buf [M, row_stride]; segment 0 =
 * heads*head_dim columns from col0 normalised with w0, optional segment 1 from col1 with w1 (nseg = 1 or 2);
 * rope_cos_sin may be NULL (norm only: the text keys of kv_y and the text-refine q/k carry no RoPE). */
int deco_headnorm_rope(void* buf_bf16, long long row_stride, int nseg, int col0, int col1,
                       const float* w0, const float* w1, const float* rope_cos_sin,
                       long long M, int heads, int head_dim, int L, float eps, void* stream);

/* Out-of-place form of deco_headnorm_rope (training forward: the raw GEMM output is kept for the backward of the norm) */
int deco_headnorm_rope_to(const void* src_bf16, void* dst_bf16, long long row_stride, int nseg, int col0, int col1,
                          const float* w0, const float* w1, const float* rope_cos_sin,
                          long long M, int heads, int head_dim, int L, float eps, void* stream);

/* Text embedder tail (dit_t2i_pixnerd.py:280 with layers/patch_embed.py:19-22, layers/rmsnorm.py:15-20):
 * out[m,:] = weight * rms(x[m,:]) + pos[m % T,:]; x = y_embedder.proj output, pos = y_pos_embedding; all fp32. */
int deco_rmsnorm_addpos(const float* x, const float* weight, const float* pos, int T, float* out,
                        long long M, int hidden, float eps, void* stream);

/* fp32 -> bf16 copy (the refined text stream as the kv_y GEMM operand, dit_t2i_pixnerd.py:47); n % 8 == 0 */
int deco_cast_f32_bf16(const float* x, void* out_bf16, long long n, void* stream);

/* scaled_dot_product_attention, non-causal, no mask (dit_c2i_DeCo.py:185; layers/attention_op.py:4).
 * q/k/v/out are strided views ([B*L, stride] rows, head h at column h*head_dim), so Q/K/V are read in place from the
 * QKV GEMM output.  A second key/value segment (k1, v1, Lk1) implements the t2i [image || text] keys
 * (dit_t2i_pixnerd.py:52-59); pass Lk1 = 0 for plain self-attention.  tcgen05/TMEM kernel fed by TMA
 * (csrc/attention_tc.cu); q_norm / k_norm / RoPE are applied beforehand (deco_qknorm_rope). */
int deco_attention_fwd(const void* q, long long q_stride,
                       const void* k0, const void* v0, long long kv0_stride, int Lk0,
                       const void* k1, const void* v1, long long kv1_stride, int Lk1,
                       void* out, long long out_stride,
                       int B, int heads, int Lq, int head_dim, float scale, void* stream);

/* Pixel decoder on tcgen05 / TMEM, optionally with the CFG-batched sampler update fused into its epilogue
 * (csrc/decoder_tc.cu; reference: dit_c2i_DeCo.py:212-248, :313-332, :395-415, :501-509; sampler fusion:
 * base/guidance.py:3-6, flow_matching/sampling.py:89-104, adam_sampling.py:104-117, autoencoder/base.py:32-34).
 * ysilu = silu(cond_embed(s)) bf16 [rows * L, p*p*32] (deco_gemm_bf16 with the bias + SiLU epilogue); blob = weight image
 * of deco_decoder_tc_blob_bytes(R) bytes (deco_b200/denoiser.py::pack_decoder_tc).
 * pair == 0: out [rows,3,H,W] (bf16 / fp32) = decoder(x [rows,3,H,W] fp32).
 * pair != 0: rows = 2B stacked [uncond || cond] over ONE image state x [B,3,H,W] fp32:
 *   pred = u + g (c - u); v = c0 pred + c1 p1; x_out = x_base + dt v (x_base NULL = x; x_out may be x / x_base); optional
 *   pred_out (may be p1), u8_out.  x_base != x is the Heun corrector: the net sees x_hat, the update starts from x.
 *   {g, dt, c0, c1} come from dev_scalars (device memory, CUDA-graph replays) when it is non-NULL. */
int deco_decoder_tc_blob_bytes(int num_res_blocks);
int deco_pixel_decoder_tc(const float* x, const void* ysilu_bf16, const void* blob, void* out, int out_is_bf16,
                          int rows, int H, int W, int patch, int hidden_x, int num_res_blocks,
                          int pair, const float* dev_scalars, float g, float dt, float c0, float c1,
                          const float* x_base, const float* p1, float* x_out, float* pred_out, void* u8_out, void* stream);

/* Heun predictor / corrector with the SDE step functions (flow_matching/sampling.py:17-24, :266-293): the score
 * s = (kd v - x) / sden at (x, t_cur) and s_hat = (kdh v_hat - x_hat) / sdenh at (x_hat, t_next) are averaged like the
 * velocities.  corrector == 0: x_out = x + dt v + a_s s + a_n z.  corrector != 0: v_hat = cfg(net_out, g) and
 * x_out = x + dt (v + v_hat)/2 + a_s (s + s_hat)/2 + a_n z; v_hat / s_hat (fp32) are stored for the next predictor.
 * s_in NULL = s computed from (x, v, kd, sden).  All tensors fp32 [n] except net_out ([2n] rows [uncond || cond]). */
int deco_heun_sde_step(const float* x, const float* v, const float* s_in, const void* net_out, int net_is_bf16,
                       const float* x_hat, const float* noise, float g, float dt, float kd, float sden,
                       float kdh, float sdenh, float a_s, float a_n, int corrector,
                       float* x_out, float* v_hat_out, float* s_hat_out, float* v_avg_out, unsigned char* u8_out,
                       long long n, void* stream);

/* Hyper-network pixel decoder of the PixNerd baseline (csrc/nerf_decoder.cu; src/models/transformer/dit_c2i_pixnerd.py:
 * 212-283, :376-380).  params[j] = output of NerfBlock j's param_generator1 (bf16 [B*L, 2*64*128]: fc1 [64 x 128] | fc2
 * [128 x 64] per token, un-normalised); the kernel applies F.normalize(dim=-2) as reciprocal column norms on the fp32
 * accumulators.  blob: deco_nerf_decoder_blob_bytes(R) bytes packed by deco_b200/denoiser_pixnerd.py.  Built for patch 16,
 * hidden_size_x 64, nerf_mlpratio 2. */
int deco_nerf_decoder_blob_bytes(int num_nerf_blocks);
int deco_nerf_decoder(const float* x, const void* const* params, int num_nerf_blocks, const void* blob,
                      void* out, int out_is_bf16, int B, int H, int W, int patch, int hidden_x, int mlp_ratio, void* stream);

/* deco_attention_bwd on tcgen05 / TMEM (csrc/attention_bwd_tc.cu): same two deterministic passes and outputs; needs the
 * forward's statistics lse2 [B*heads*Lq] (deco_attention_fwd_lse); delta_ws [B*heads*Lq] fp32 is scratch filled by the first
 * pass.  All row strides multiples of 8 elements, pointers 16-byte aligned. */
int deco_attention_bwd_tc(const void* q, long long q_stride, const void* k, const void* v, long long kv_stride,
                          const void* o, long long o_stride, const void* dout, long long do_stride,
                          void* dq, long long dq_stride, void* dk, void* dv, long long dkv_stride,
                          const float* lse2, float* delta_ws, int B, int heads, int Lq, int Lk, int head_dim,
                          float scale, void* stream);

/* Training-step inputs (src/diffusion/flow_matching/training_repa_DeCo.py:222-237, src/diffusion/base/training.py:14-20).
 * The random draws stay the caller's (torch's CUDA generator, reference order); these fuse what follows them.
 * deco_train_timesteps: t = time_shift(where(u_select <= 0.9, sigmoid(nt), u_uniform)) (fp32 [B]); with
 *   linear_scheduler != 0 also coef[B][4] = (alpha, sigma, dalpha, dsigma) = (t, 1-t, 1, -1) (scheduling.py:6-14).
 * deco_flow_pair: x_t = alpha x + sigma eps, v_t = dalpha x + dsigma eps with per-image coef[B][4]; fp32 [B, per_image].
 * deco_label_dropout: out[i] = u[i] < p ? uncond[i] : cond[i] (int64 labels). */
int deco_train_timesteps(const float* nt, const float* u_uniform, const float* u_select, float timeshift,
                         int linear_scheduler, float* t_out, float* coef_out, int B, void* stream);
int deco_flow_pair(const float* x, const float* eps, const float* coef, float* x_t, float* v_t,
                   int B, long long per_image, void* stream);
int deco_label_dropout(const long long* cond, const long long* uncond, const float* u, float p,
                       long long* out, int B, void* stream);

/* deco_attention_fwd with explicit head pitches: heads of a q / k / v row sit *_head_pitch elements apart (0 = head_dim,
 * the dense [head][d] row).  The fused QKV GEMM (deco_gemm_norm_qkv with out_head_pitch = 80) writes head_dim-72 heads at a
 * pitch of 80 so that the 16-column TMA boxes of the attention operands are aligned 32-byte sectors. */
int deco_attention_fwd_pitched(const void* q, long long q_stride, int q_head_pitch,
                               const void* k0, const void* v0, long long kv0_stride, int kv0_head_pitch, int Lk0,
                               const void* k1, const void* v1, long long kv1_stride, int kv1_head_pitch, int Lk1,
                               void* out, long long out_stride,
                               int B, int heads, int Lq, int head_dim, float scale, void* stream);

/* s = silu(t + s) (dit_c2i_DeCo.py:499): out[m,:] = silu(x[m,:] + row[m / rows_per,:]); out may alias x */
int deco_silu_add_rows(const void* x, int x_is_f32, const void* row_bf16, void* out_bf16, long long M, int hidden,
                       int rows_per, void* stream);

/* NerfEmbedder + SimpleMLPAdaLN + fold (dit_c2i_DeCo.py:212-248, :288-415, :501-509).
 * x fp32 [B,3,H,W]; ycond = cond_embed output bf16 [B*L, p*p*32]; blob = packed weights
 * (deco_decoder_blob_bytes bytes, layout in csrc/decoder.cu, packed by deco_b200/denoiser.py); postab fp32 [p*p, 32].
 * out [B,3,H,W] bf16 or fp32.  Built for patch 16, hidden_size_x 32. */
int deco_decoder_blob_bytes(int num_res_blocks);
int deco_pixel_decoder(const float* x, const void* ycond_bf16, const void* blob, const float* postab,
                       void* out, int out_is_bf16, int B, int H, int W, int patch, int hidden_x,
                       int num_res_blocks, void* stream);

/* CFG combine + sampler state update (base/guidance.py:3-6; flow_matching/sampling.py:14-15,:89-104,:283-291;
 * flow_matching/adam_sampling.py:109-117; autoencoder/base.py:32-34):
 *   pred = u + g (c - u);  v = c0 pred + c1 p1 + c2 p2 + c3 p3;  x_out = x + dt v
 * net_out [2n] rows [uncond || cond] (bf16 or fp32); optional outputs pred_out, v_out (fp32), u8_out = fp2uint8(x_out).
 * n = elements per CFG half, multiple of 4. */
int deco_cfg_step(const float* x, const void* net_out, int net_is_bf16,
                  const float* p1, const float* p2, const float* p3,
                  float g, float dt, float c0, float c1, float c2, float c3,
                  float* x_out, float* pred_out, float* v_out, uint8_t* u8_out, long long n, void* stream);
int deco_fp2uint8(const float* x, uint8_t* out, long long n, void* stream);

/* The same update for CUDA-graph replays of the sampling loop (sampling.py:89-104 without per-step host work):
 * deco_sampler_advance copies row (*counter % rows) of the host-precomputed schedule table {g, dt, c0, c1, c2, c3, t, -}
 * (fp32 [rows, 8], device) into cur_params[8], broadcasts its t into t_out[nt] (the denoiser's timestep vector) and
 * increments the counter; deco_cfg_step_dev is deco_cfg_step with {g, dt, c0..c3} read from dev_params (= cur_params).
 * x_out may alias x, pred_out may alias p1. */
int deco_sampler_advance(const float* table, int rows, int* counter, float* cur_params, float* t_out, int nt, void* stream);
int deco_cfg_step_dev(const float* x, const void* net_out, int net_is_bf16,
                      const float* p1, const float* p2, const float* p3, const float* dev_params,
                      float* x_out, float* pred_out, float* v_out, uint8_t* u8_out, long long n, void* stream);

/* The same pass for the other samplers / step functions that share it (flow_matching/sampling.py):
 *   xpred_den > 0 (EulerSamplerJiT, :170): the net predicts x; u' = (u - x) / xpred_den, c' = (c - x) / xpred_den with
 *     xpred_den = clamp_min(1 - t, 0.05), then pred = u' + g (c' - u')
 *   SDE step functions (:17-24, score at :98): s = (kd v - x) / sden, kd = 1 / dalpha_over_alpha(t),
 *     sden = sigma(t)^2 - kd dsigma_mul_sigma(t);  x_out = x + dt v + a_s s + a_n noise with
 *     (a_s, a_n) = (w dt, 0) sde_mean | (w dt, sqrt(2 w dt)) sde | (w dt / 2, sqrt(w dt)) sde_preserve; noise fp32 [n]
 *     (caller-generated N(0,1)), read only when a_n != 0
 * dev_params: NULL, or the device vector {g, dt, c0, c1, c2, c3, t, xpred_den} written by deco_sampler_advance (then the
 * scalar g..c3 and xpred_den arguments are ignored).  x_out may alias x, pred_out may alias p1. */
int deco_cfg_step_ex(const float* x, const void* net_out, int net_is_bf16,
                     const float* p1, const float* p2, const float* p3, const float* dev_params,
                     float g, float dt, float c0, float c1, float c2, float c3,
                     float xpred_den, float kd, float sden, float a_s, float a_n, const float* noise,
                     float* x_out, float* pred_out, float* v_out, uint8_t* u8_out, long long n, void* stream);

/* Output head of the patch-linear baseline denoiser (models/transformer/dit_c2i_baseline.py):
 * deco_layernorm_modulate = FinalLayer's LayerNorm(no affine) + modulate (:76-82) on the fp32 stream x [M, hidden]
 *   -> bf16 [M, hidden]; shift / scale bf16 rows (one per rows_per_mod stream rows) sharing mod_row_stride;
 * deco_unpatchify = F.fold(kernel = stride = p) (:378): tokens bf16 [B*L, C*p*p] -> image bf16 [B, C, H, W]. */
int deco_layernorm_modulate(const float* x, const void* shift_bf16, const void* scale_bf16, long long mod_row_stride,
                            int rows_per_mod, void* out_bf16, long long M, int hidden, float eps, void* stream);
int deco_unpatchify(const void* tok_bf16, void* out_bf16, int B, int C, int H, int W, int p, void* stream);
/* out = x - rowmean(x), fp32 [M, hidden], in place allowed: LayerNorm(no affine) = RMSNorm of the centred row, which is how the
 * TRAINING path of the final layer (:76-82) reuses deco_rmsnorm_modulate and its backward (centre, norm; norm', centre). */
int deco_center_rows(const float* x, float* out, long long M, int hidden, void* stream);

/* Frequency-aware FM loss, forward and/or backward in one pass
 * (flow_matching/training_repa_DeCo.py:106-136 _rgb2ycbcr/_dct, :138-195 weights, :273-285 loss):
 *   losses[0] = mean((out - v_t)^2), losses[1] = mean(freq_w * dct(ycbcr(out - v_t))^2), losses[2] = [0] + flw * [1]
 *   grad = upstream * d losses[2] / d out   (same dtype as out; fp32 required for ragged H/W)
 * out [B,3,H,W] fp32 or bf16; v_t fp32; freq_w fp32 [3,8,8]; accum = deco_dct_scratch_doubles() doubles of scratch whose
 * first three must be ZERO on entry and are zero again on exit (every CTA leaves its partial sums in its own slot, the last
 * one adds them in a fixed order -- the losses are bit-reproducible -- publishes and clears the ticket: forward + backward
 * is one launch); losses and/or grad may be NULL; upstream = device scalar or NULL (1.0). */
int deco_dct_scratch_doubles(void);
int deco_dct_fm_loss(const void* out, int out_is_bf16, const float* v_t, const float* freq_w,
                     int B, int H, int W, float freq_loss_weight,
                     float* losses, void* grad, const float* upstream, double* accum, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Backward of the denoiser (training step, BASELINE configs[3]; reference: PyTorch autograd over dit_c2i_DeCo.py).
 * Dense dgrad / wgrad contractions reuse deco_gemm_bf16 (dX = dY.W with W^T as the "weight" operand; dW = dY^T.X with
 * both operands transposed by deco_transpose_cast); the entry points below are the memory-bound glue, the attention
 * backward and the pixel-decoder backward.  Arguments named *_accum are ACCUMULATED into (atomics): zero them first.
 * ------------------------------------------------------------------------------------------------------------------ */

/* dst[c][r] = bf16(src[r][c]) for r < R, zero for R <= r < Rp: [R, C] (fp32 or bf16, row stride lds) -> [C, ldd] */
int deco_transpose_cast(const void* src, int src_is_f32, long long lds, void* dst_bf16, long long ldd,
                        int R, int C, int Rp, void* stream);

/* out_accum[c] += sum_r x[r][c]  (bias gradients of nn.Linear) */
int deco_colsum(const void* x, int x_is_f32, long long ldx, float* out_accum, long long M, int N, void* stream);

/* gate_residual followed by the next block-level RMSNorm + modulation in one pass (training forward,
 * dit_c2i_DeCo.py:236-244): s_out = s + gate[row / L] * a (fp32) and h_out = rms(s_out) * weight * (1 + scale) + shift (bf16).
 * Bit-identical to deco_gate_residual + deco_rmsnorm_modulate. */
int deco_gate_residual_norm(const float* s, const void* a_bf16, const void* gate_bf16, long long gate_stride,
                            float* s_out, const float* weight, const void* shift_bf16, const void* scale_bf16,
                            long long mod_row_stride, int rows_per_image, void* h_out_bf16, long long M, int hidden,
                            float eps, void* stream);

/* out[m,:] = s[m,:] + gate[m / rows_per_image,:] * a[m,:]  (dit_c2i_DeCo.py:208-209 with the branch output a kept for
 * backward); out may alias s */
int deco_gate_residual(const float* s, const void* a_bf16, const void* gate_bf16, long long gate_stride,
                       float* out, int rows_per_image, long long M, int hidden, void* stream);

/* backward of the above: da = gate * ds (bf16); dgate[b,:] += sum_rows ds * a; dbias += sum_rows da (may be NULL; else
 * img_ws = ZEROED fp32 [B, hidden] scratch for the per-image partial sums) */
int deco_gate_bwd(const float* ds, const void* a_bf16, const void* gate_bf16, long long gate_stride,
                  void* da_bf16, float* dgate_accum, long long dgate_stride, float* dbias_accum, float* img_ws,
                  int rows_per_image, long long M, int hidden, void* stream);

/* backward of out = silu(x + row[b]) (dit_c2i_DeCo.py:499): dx = dout * silu'(x + row) (fp32, written);
 * drow[b,:] += sum_rows dx */
int deco_silu_add_rows_bwd(const void* dout_bf16, const float* x, const void* row_bf16, float* dx,
                           float* drow_accum, int rows_per_image, long long M, int hidden, void* stream);

/* SwiGLU on the interleaved [16 x w1 | 16 x w3] columns of y13 [M, 2*ffn_pad] (dit_c2i_DeCo.py:112-114):
 * forward u = silu(a) * b [M, ffn_pad]; backward dy13 from du (same interleaved layout as y13) */
int deco_swiglu_fwd(const void* y13_bf16, void* u_bf16, long long M, int ffn_pad, void* stream);
int deco_swiglu_bwd(const void* y13_bf16, const void* du_bf16, void* dy13_bf16, long long M, int ffn_pad, void* stream);

/* backward of deco_rmsnorm_modulate on the fp32 stream (dit_c2i_DeCo.py:94-99, :11-12): ds_accum[m,:] += d x;
 * dweight_accum[hidden], dshift_accum / dscale_accum [B, dmod_row_stride]; row_ws = 2*M floats of scratch (8-byte aligned)
 * for the per-row scalars of the first pass; img_ws = ZEROED fp32 [B, hidden] scratch for per-image partial sums */
int deco_rmsnorm_modulate_bwd(const void* dh_bf16, const float* x, const float* weight, const void* scale_bf16,
                              long long mod_row_stride, float* ds_accum, float* dweight_accum,
                              float* dshift_accum, float* dscale_accum, long long dmod_row_stride,
                              float* row_ws, float* img_ws, int rows_per_image, long long M, int hidden, float eps,
                              void* stream);

/* deco_rmsnorm_modulate_bwd followed by deco_gate_bwd on the updated stream gradient in ONE pass over it (the backward of
 * "s_mid = s + gate * a; h = norm(s_mid)", dit_c2i_DeCo.py:236-244 reversed): da = gate * ds_new, dgate += sum_rows ds_new * a,
 * dbias += sum_rows da (optional; gate_img_ws = zeroed fp32 [B, hidden]). */
int deco_rmsnorm_modulate_bwd_gate(const void* dh_bf16, const float* x, const float* weight, const void* scale_bf16,
                                   long long mod_row_stride, float* ds_accum, float* dweight_accum,
                                   float* dshift_accum, float* dscale_accum, long long dmod_row_stride,
                                   float* row_ws, float* img_ws, int rows_per_image, long long M, int hidden, float eps,
                                   const void* a_bf16, const void* gate_bf16, long long gate_stride, void* da_bf16,
                                   float* dgate_accum, long long dgate_stride, float* dbias_accum, float* gate_img_ws,
                                   void* stream);

/* backward of per-head RMSNorm (+ RoPE when rope_cos_sin != NULL) for one segment (dit_c2i_DeCo.py:178-180, :134-145):
 * g [M, g_stride] holds d(out) at columns [col, col + heads*head_dim) on entry and d(raw) on exit; raw = the QKV GEMM
 * output the forward normalised; dweight_accum [head_dim] */
int deco_headnorm_rope_bwd(void* g_bf16, long long g_stride, const void* raw_bf16, long long raw_stride,
                           int col, const float* weight, const float* rope_cos_sin, float* dweight_accum,
                           long long M, int heads, int head_dim, int L, float eps, void* stream);

/* backward of deco_cond_combine (dit_c2i_DeCo.py:493-494): dpre = dc * silu'(temb + table[y]);
 * dtemb_accum += dpre; dtable_accum[y] += dpre */
int deco_cond_combine_bwd(const float* dc, const void* temb_bf16, const float* table, const long long* labels,
                          float* dtemb_accum, float* dtable_accum, int B, int hidden, int num_rows, void* stream);

/* dz = dy * silu'(z)  (t_embedder.mlp[1], dit_c2i_DeCo.py:55-57) */
int deco_silu_bwd(const void* z_bf16, const void* dy_bf16, void* dz_bf16, long long n, void* stream);

/* deco_attention_fwd (one key segment) that also writes lse2_out [B*heads, Lq] = log2 sum_k exp2(scale log2(e) q.k) */
int deco_attention_fwd_lse(const void* q, long long q_stride, const void* k, const void* v, long long kv_stride,
                           int Lk, void* out, long long out_stride, float* lse2_out,
                           int B, int heads, int Lq, int head_dim, float scale, void* stream);

/* backward of deco_attention_fwd (single key segment): dq / dk / dv are strided views like q / k / v;
 * lse2_ws and delta_ws are fp32 buffers of B*heads*Lq elements (csrc/attention_bwd.cu); have_lse = 1: lse2_ws holds
 * deco_attention_fwd_lse's output (the dQ kernel then skips rebuilding the softmax statistics), 0: it is scratch */
int deco_attention_bwd(const void* q, long long q_stride, const void* k, const void* v, long long kv_stride,
                       const void* o, long long o_stride, const void* dout, long long do_stride,
                       void* dq, long long dq_stride, void* dk, void* dv, long long dkv_stride,
                       float* lse2_ws, float* delta_ws, int have_lse, int B, int heads, int Lq, int Lk, int head_dim,
                       float scale, void* stream);

/* backward of deco_pixel_decoder: dycond bf16 [B*L, p*p*32]; grad_accum = deco_decoder_train_blob_floats(R) floats in
 * the layout of blob_f32 (csrc/decoder_bwd.cu) followed by d postab [p*p, 32]; dout fp32 [B,3,H,W] */
int deco_decoder_train_blob_floats(int num_res_blocks);
int deco_pixel_decoder_bwd(const float* x, const void* ycond_bf16, const float* dout, const float* blob_f32,
                           const float* postab, void* dycond_bf16, float* grad_accum, int B, int H, int W,
                           int patch, int hidden_x, int num_res_blocks, void* stream);

/* Same contract on warp-level tensor-core MMAs (csrc/decoder_bwd_mma.cu, the product path): fwd_blob = the forward's packed
 * weights (deco_decoder_blob_bytes), bwd_blob = transposed weights in fragment order + fp32 final layer
 * (deco_decoder_bwd_blob_bytes; packed by deco_b200/autograd.py::pack_decoder_bwd); grad_accum as above */
int deco_decoder_bwd_blob_bytes(int num_res_blocks);
int deco_pixel_decoder_bwd_tc(const float* x, const void* ycond_bf16, const float* dout, const void* fwd_blob,
                              const void* bwd_blob, const float* postab, void* dycond_bf16, float* grad_accum,
                              int B, int H, int W, int patch, int hidden_x, int num_res_blocks, void* stream);

/* wgrad contraction read straight from row-major activations: out[M, N] fp32 = At^T . Wt with At [K, M], Wt [K, N] bf16
 * (dW = dY^T . X with K = tokens).  Operands are staged MN-major (csrc/gemm_tcgen05.cu, TN mode); M, N, lda, ldw multiples
 * of 8; tile_n in {0 (auto), 128, 256}. */
int deco_gemm_bf16_tn(const void* At, long long lda, const void* Wt, long long ldw, float* out, long long ldo,
                      int M, int N, int K, int tile_n, int split_k, void* stream);
/* The same contraction for the SwiGLU weight gradient: the M rows of dY^T are the [16 x w1 | 16 x w3]-interleaved rows of
 * [w1 ; w3] (dit_c2i_DeCo.py:101-113 fused into one GEMM); they are stored de-interleaved, out = [2][M / 2][ldo] = (dW1 ; dW3).
 * M a multiple of 32. */
int deco_gemm_bf16_tn_deint16(const void* At, long long lda, const void* Wt, long long ldw, float* out, long long ldo,
                              int M, int N, int K, int tile_n, int split_k, void* stream);

/* Split-K: when the M x N tiles alone cannot fill the GPU (small weight matrices with a long token reduction; d c = d mod .
 * Wada with M = batch), the K loop is cut into split_k slices (0 = automatic, 1 = off) that run as separate tiles and are
 * reduced with fp32 atomics into the (zeroed) output.  deco_gemm_bf16_f32_splitk is deco_gemm_bf16 with
 * DECO_EPI_BIAS_F32, no bias. */
int deco_gemm_bf16_f32_splitk(const void* A, long long lda, const void* W, long long ldw, float* out, long long ldo,
                              int M, int N, int K, int split_k, void* stream);

/* Fused multi-tensor AdamW + EMA update (csrc/optimizer.cu): torch.optim.AdamW semantics (configs_c2i/DeCo_XL.yaml:89-93)
 * followed by ema = decay * ema + (1 - decay) * p (src/callbacks/simple_ema.py:27-39), one launch for all tensors.
 * tensor_table: device array of {float* p, const float* g, float* m, float* v, float* ema (or NULL), int64 n};
 * chunk_table: device array of int32 pairs (tensor index, chunk index), one per CTA, chunks of deco_opt_chunk_elems()
 * elements; bias_correction{1,2} = 1 - beta^t. */
int deco_opt_chunk_elems(void);
int deco_adamw_ema_step(const void* tensor_table, const void* chunk_table, int num_chunks,
                        float lr, float beta1, float beta2, float eps, float weight_decay,
                        float bias_correction1, float bias_correction2, float ema_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DECO_B200_H */
