/* libs2i -- C ABI of the B200-native sketch-guided Stable-Diffusion sampling path.
 *
 * The reference (Mikubill/sketch2img) is pure Python: the boundary a maintainer binds is this
 * plain-C surface (ctypes), called from drop-in replacements of
 *   modules/pipeline.py:13-174          AntiGradientPipeline  (loop body :83-115, guidance :141-161)
 *   modules/latent_predictor.py:9-81    LatentEdgePredictor.forward, hook_unet
 *   modules/sketch_guided_attn.py:8-161 SatMixin / AttnModule
 * All pointers are raw device pointers unless marked host; no torch types cross this boundary.
 * Every function returns 0 on success or a negative code; s2i_last_error() has the message.
 */
#ifndef S2I_H_
#define S2I_H_

#ifdef __cplusplus
extern "C" {
#endif

#define S2I_OK 0
#define S2I_ERR_ARG (-1)
#define S2I_ERR_CUDA (-2)
#define S2I_ERR_STATE (-3)
#define S2I_ERR_OOM (-4)

const char* s2i_last_error(void);
/* kernels launched by this library since load (bench.py's gpu_launches) */
long long s2i_launch_count(void);
/* Opt-in per-launch device timing (bench.py's roofline numbers; no reference counterpart -- the reference ships no
 * profiling).  Between begin and end every kernel this library launches on `cuda_stream` is
 * bracketed by CUDA events; end synchronises the stream and writes one text line per kernel class:
 * "<class> <launches> <ms> <flops> <bytes>".  Returns the number of characters written, or a negative code. */
int s2i_profile_begin(void* cuda_stream);
int s2i_profile_end(char* report, int capacity);
/* Roofline denominators (dense fp16/bf16 TFLOP/s, HBM GB/s): with them set, each line of the report carries a sixth column,
 * the sum over the class's launches of  max(FLOPs / peak, algorithmic bytes / bandwidth)  in ms -- the time the launches
 * would take at their own roofline, whichever resource bounds each of them. */
int s2i_profile_set_peaks(double tflops, double hbm_gbs);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 + TMA implicit GEMM:  C[z] = alpha * A[z] * B[z]^T (+ bias, per-sample vector, ReLU, residual).
 * Replaces every dense contraction diffusers/torch run under modules/pipeline.py:96 (conv3x3, conv1x1,
 * Linear, attention QK^T / PV) and latent_predictor.py:45 (LGP MLP), forward and input-gradient.
 *   A, K-major : 4-D (C, W, H, B), C contiguous; GEMM row = pixel (b,y,x); K = taps x C (3x3 taps, pad 1).
 *   A, MN-major: same tensor read as [K rows = W][M = C]; batch on B.
 *   B          : 3-D (I, R, Z), I contiguous. K-major: R = N rows, I = K.  MN-major: R = K rows, I = N.
 *   z in [0,Z): zb = z / zh, zhd = z % zh; operand inner offset = c0 + zhd*hoff and batch coordinate zb
 *   (zmode 0) or inner offset c0 and batch coordinate z (zmode 1).  Output offset zb*c_sb + zhd*c_sh.
 * --------------------------------------------------------------------------------------------- */
typedef struct s2i_gemm_desc {
    const void* A; int a_mn; int aC, aW, aH, aB; long long a_sw, a_sh, a_sb; int taps; int a_c0, a_hoff, a_zmode;
    const void* B; int b_mn; int bI, bR, bZ; long long b_sr, b_sz; int b_c0, b_hoff, b_zmode;
    int N, Kc, Z, zh; int bf16; int BN; int splits;   /* BN / splits: 0 = chosen by the library; splits < 0 = never split K */
    float alpha; const float* bias; const float* rowvec; int rowvec_ld;
    const float* residual; long long res_ld;
    float* out32; long long ld32; void* out16; long long ld16; int out16_bf16;
    long long c_sb, c_sh; int relu;
    float qscale;   /* != 0: round the result as fp16(v / qscale) * qscale (mimics unscaled fp16 autograd rounding) */
    /* Gated-GELU epilogue (diffusers GEGLU, the feed-forward of every BasicTransformerBlock inside modules/pipeline.py:96):
     * the weight rows are interleaved in blocks of 32 -- output columns [64 b, 64 b + 32) = value features 32 b .., columns
     * [64 b + 32, 64 b + 64) = their gates -- and out_glu [rows][N / 2] (fp16, pixel stride ld_glu) receives
     * value * gelu(gate).  out16 (optional) still receives the interleaved projection itself.  Needs N % 64 == 0 and the
     * TMA-epilogue kernel (K-major operands, Z = 1, no residual / out32 / split-K). */
    void* out_glu; long long ld_glu;
    /* Optional fp32 scratch [rows][N] (dense) owned by the caller: an fp16-only output whose K is split reduce-adds its
     * partial tiles there before the conversion, instead of in the library's shared scratch -- a per-call buffer can be
     * zeroed ahead of time together with the other split-K outputs of a captured step. */
    float* scratch32;
    /* != 0: B is a tensor no kernel of the surrounding stream writes (packed weights).  gemm_tma_kernel then loads its first B
     * tiles BEFORE waiting for the preceding kernel (programmatic dependent launch), so the weight stream of launch n + 1
     * overlaps the epilogue of launch n.  Leave 0 when B is an activation. */
    int b_static;
    /* Optional per-channel statistics of the fp32 result, for a GroupNorm that follows (norm1 / norm2 of diffusers' ResnetBlock2D,
     * the norm of Transformer2DModel): colstat [B][colstat_cap][2][colstat_ld] floats receives, for every block of result rows a
     * CTA (or one CTA of a split-K cluster) owns, the column sums and the column sums of squares; *colstat_bps (host) is set to the
     * number of blocks per sample used, or to 0 when this launch could not provide them (the caller then computes the statistics
     * itself).  Needs the TMA-epilogue kernel, out32, and an exactly tiled pixel grid. */
    float* colstat; long long colstat_ld; int colstat_cap; int* colstat_bps;
} s2i_gemm_desc;

int s2i_gemm(const s2i_gemm_desc* d, void* cuda_stream);
/* Bisecting / A-B switch (no reference counterpart): 1 (default) = eligible GEMMs run gemm_tma_kernel (residual tile in,
 * result tiles out through cp.async.bulk.tensor), 0 = every GEMM runs gemm_tc_kernel (per-thread epilogue). */
int s2i_gemm_set_tma_epilogue(int on);
/* Tools / tests: 2 forces gemm_tma_kernel's two-sub-tile form (256 x BN per CTA: two A tiles share each B tile, two TMEM
 * accumulators) wherever it is legal, 1 forbids it, 0 (default) lets the cost model choose. */
int s2i_gemm_force_msub(int msub);
/* Tools / tests: CTA pairs in gemm_tma_kernel (two adjacent 128-row tiles run one tcgen05.mma.cta_group::2 stream and load half
 * of each B tile each).  1 = wherever legal, 0 = never, -1 (default) = the cost model's choice (env S2I_GEMM_PAIR overrides). */
int s2i_gemm_set_pair(int mode);
/* Debugging: device buffer of [ctas][16] uint64 that gemm_tma_kernel fills with %globaltimer stamps of its phases
 * (entry, setup done, loads issued, MMAs issued, epilogue start, accumulator ready, residual ready, chunks done, stores
 * read, exit); NULL switches it off. */
int s2i_gemm_set_trace(void* device_buf);


/* ---------------------------------------------------------------------------------------------
 * Fused attention forward (tcgen05 + TMA; scores stay on chip): replaces the xformers / CrossAttention call inside
 * the UNet (app.py:43, modules/pipeline.py:96) wherever the probabilities are not needed by a later backward.
 *   out[b, i, h*dp .. h*dp+dp) = softmax_j(scale * <Q[b,i,h], K[b,j,h]>) V[b,j,h]
 *   q : fp16 device [B][Nq][ldq], head h of Q at columns q_c0 + h*dp
 *   kv: fp16 device [B][Nk][ldkv], head h of K at k_c0 + h*dp, of V at v_c0 + h*dp
 *   dp = head dim padded to a multiple of 16 (padding columns zero); Nq >= 128, Nk >= 64
 *   lse (optional): fp32 device [B*heads][Nq] = log-sum-exp of the scaled scores
 * --------------------------------------------------------------------------------------------- */
int s2i_attention(const void* q, long long ldq, int q_c0, const void* kv, long long ldkv, int k_c0, int v_c0, int B,
                  int heads, int Nq, int Nk, int dp, int d_true, float scale, void* out, long long ldo, float* lse,
                  void* cuda_stream);

/* Fused attention backward (tcgen05 + TMA; scores / probabilities recomputed on chip from the forward's log-sum-exp):
 * the attention part of `torch.autograd.grad(loss, latents_prev)` (modules/pipeline.py:159).
 *   d_out, out: fp16 device [B][Nq][ldo] (gradient of / value of the forward output), head h at columns h*dp
 *   lse: the forward's log-sum-exp; delta_scratch: fp32 device [B*heads][Nq]
 *   dq : fp16 device [B][Nq][lddq], head h at dq_c0 + h*dp
 *   dkv: fp16 device [B][Nk][lddkv] or NULL (cross-attention to a constant context: only dQ); dK heads at dk_c0 + h*dp,
 *        dV heads at dv_c0 + h*dp.   Nq >= 128, Nk >= 64, dp a multiple of 16 and <= 192. */
int s2i_attention_backward(const void* q, long long ldq, int q_c0, const void* kv, long long ldkv, int k_c0, int v_c0,
                           const void* d_out, const void* out, long long ldo, const float* lse, float* delta_scratch, int B,
                           int heads, int Nq, int Nk, int dp, int d_true, float scale, void* dq, long long lddq, int dq_c0,
                           void* dkv, long long lddkv, int dk_c0, int dv_c0, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * GroupNorm(32 groups) [+ SiLU] of an NHWC fp32 tensor, statistics and apply in ONE launch (thread-block clusters own a
 * few groups of a sample, exchange their partial sums through distributed shared memory, and apply from the slab they
 * staged in shared memory): the `norm1/norm2 -> nonlinearity` of every diffusers ResnetBlock2D and the `norm` of every
 * Transformer2DModel inside `self.unet(...)` (modules/pipeline.py:96), and their part of the autograd backward (:159).
 *   x, dy, add, dx32: fp32 device [B][HW][ld*]; out16 / raw16 / dx16: fp16 device; gamma, beta: fp32 [C]
 *   stats: device scratch of B * 512 bytes, zeroed before the forward; the backward reads the forward's and needs its own
 *   forward : out16 = fp16(act(GN(x))), raw16 (optional) = fp16(x);  act = SiLU when silu != 0
 *   backward: dx = dGN/dx applied to (dy * act'(GN(x))) (+ add), written as fp32 (dx32) and / or fp16 (dx16)
 * C must be a multiple of 32 and of 4, every ld a multiple of 4.
 * --------------------------------------------------------------------------------------------- */
int s2i_groupnorm_forward(const float* x, long long ldx, int B, int HW, int C, const float* gamma, const float* beta,
                          float eps, int silu, void* out16, long long ld16, void* raw16, long long ldraw, void* stats,
                          void* cuda_stream);
/* The same forward with the statistics taken from the GEMM that produced x (s2i_gemm_desc.colstat: per-block column sums and sums
 * of squares [B][colstat_cap][2][colstat_ld], colstat_bps blocks per sample in use): ONE streaming pass over x, no reduction pass.
 * `stats` receives the per-group mean / rstd like s2i_groupnorm_forward (the backward reads them). */
int s2i_groupnorm_forward_colstat(const float* x, long long ldx, int B, int HW, int C, const float* colstat,
                                  long long colstat_ld, int colstat_cap, int colstat_bps, const float* gamma,
                                  const float* beta, float eps, int silu, void* out16, long long ld16, void* raw16,
                                  long long ldraw, void* stats, void* cuda_stream);
int s2i_groupnorm_backward(const float* dy, long long ldd, const float* x, long long ldx, int B, int HW, int C,
                           const float* gamma, const float* beta, float eps, int silu, const void* fwd_stats,
                           void* bwd_stats, const float* add, long long ldadd, float* dx32, long long ld32, void* dx16,
                           long long ld16, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * UNet2DCondition engine: replaces `self.unet(x, t, encoder_hidden_states=...)` (modules/pipeline.py:96),
 * the 9 forward hooks of hook_unet (modules/latent_predictor.py:47-81) and the UNet part of
 * `torch.autograd.grad(loss, latents_prev)` (modules/pipeline.py:159).
 * Weights are passed once as host fp32 tensors under their diffusers state-dict names.
 * --------------------------------------------------------------------------------------------- */
typedef struct s2i_unet_config {
    int in_channels, out_channels;
    int block_out_channels[4];
    int num_heads[4];              /* diffusers `attention_head_dim` (= number of heads) per level */
    int layers_per_block;
    int cross_attention_dim;
    int sample_size;
    int ctx_len;                   /* text tokens (77) */
} s2i_unet_config;

typedef struct s2i_unet s2i_unet;

int s2i_unet_create(const s2i_unet_config* cfg, s2i_unet** out);
void s2i_unet_destroy(s2i_unet* u);
/* shapes: n x 4 (unused dims = 1); host pointers must stay valid until the call returns */
int s2i_unet_load(s2i_unet* u, int n, const char* const* names, const float* const* host_ptrs, const int* ndims,
                  const long long* shapes);
/* x, eps: NCHW fp32 device [B,C,H,W]; ctx: fp32 device [B,ctx_len,cross_attention_dim] */
int s2i_unet_forward(s2i_unet* u, const float* x, int B, int H, int W, float t, const float* ctx, float* eps,
                     int save_for_backward, void* cuda_stream);
/* tap k of the last forward: NHWC fp32 device view owned by the engine (valid until the next forward) */
int s2i_unet_tap(s2i_unet* u, int k, float** ptr, int* B, int* H, int* W, int* C);
/* element stride between consecutive pixels of tap k (>= C: some taps are slices of the up path's concat buffers) */
int s2i_unet_tap_stride(s2i_unet* u, int k, long long* pixel_stride);
/* tap_grads[k]: NHWC fp32 device, shaped like tap k (NULL = no gradient); dx: NCHW fp32 device [B,C,H,W] */
int s2i_unet_backward(s2i_unet* u, float* const* tap_grads, float* dx, void* cuda_stream);
/* The same walk over the samples [b0, b0 + nb) of the forward's batch only: tap_grads[k] and dx hold nb samples.
 * Samples are independent computations, and modules/pipeline.py:159 keeps only the cond half of the gradient. */
int s2i_unet_backward_samples(s2i_unet* u, float* const* tap_grads, float* dx, int b0, int nb, void* cuda_stream);
/* Injected sketch attention (modules/sketch_guided_attn.py:8-161, SatMixin / AttnModule; forward only).
 * load_sat: host fp32 tensors under the reference's parameter names
 *   "sketch_attn_<block path, '.' -> '_'>_transformer_blocks_0.{sketch_norm.{weight,bias}, sketch_attn.to_{q,k,v}.weight,
 *    sketch_attn.to_out.0.{weight,bias}, sketch_conv.{weight,bias}}"  for each of the 16 transformer blocks.
 * set_sat_feature: the sketch-encoder feature of one block (AttnModule.set_res_sample), NCHW fp32 device [B,C,H,W] with B =
 *   the forward batch and (C, H*W) = the block's width and token count; block_path e.g. "down_blocks.0.attentions.1";
 *   NULL removes it (the block then runs unmodified, sketch_guided_attn.py:120).  K / V are projected once here.
 * set_sat_scale: AttnModule.set_scale for every block. */
int s2i_unet_load_sat(s2i_unet* u, int n, const char* const* names, const float* const* host_ptrs, const int* ndims,
                      const long long* shapes);
int s2i_unet_set_sat_feature(s2i_unet* u, const char* block_path, const float* feature_nchw, int B, int C, int H, int W,
                             void* cuda_stream);
int s2i_unet_set_sat_scale(s2i_unet* u, float scale, void* cuda_stream);
/* Sketch feature encoder (modules/sketch_encoder.py:11-98, SketchEncoder(UNet2DConditionModel)): conv_in, the time embedding
 * and the down path of a UNet whose four down blocks are attention-free (the reference's forward calls the down blocks
 * without encoder_hidden_states, :93-95, which only executes for DownBlock2D); the forward stops after the down blocks and
 * keeps every block's res_samples -- the input of SatMixin.set_res_samples (modules/sketch_guided_attn.py:29-40).
 * Weights: host fp32 tensors under their diffusers names (conv_in.*, time_embedding.*, down_blocks.*), loaded with
 * s2i_unet_load; destroyed with s2i_unet_destroy.  res_sample k: NHWC fp32 view owned by the engine (valid until the next
 * forward), in block order: per down block its layers_per_block resnet outputs, then the downsampled map (not the last block). */
int s2i_sketch_encoder_create(const s2i_unet_config* cfg, s2i_unet** out);
int s2i_sketch_encoder_forward(s2i_unet* u, const float* x, int B, int H, int W, float t, void* cuda_stream);
int s2i_sketch_encoder_num_res_samples(s2i_unet* u);
int s2i_sketch_encoder_res_sample(s2i_unet* u, int k, float** ptr, long long* pixel_stride, int* B, int* H, int* W, int* C);
/* bisecting aid: keep named block outputs of the next forwards ("conv_in", "down0".., "mid", "up0"..) */
int s2i_unet_debug(s2i_unet* u, int enable);
int s2i_unet_debug_get(s2i_unet* u, const char* name, float** ptr, long long* ld, int* B, int* H, int* W, int* C);
long long s2i_unet_arena_bytes(s2i_unet* u);

/* ---------------------------------------------------------------------------------------------
 * AutoencoderKL (the SD VAE), the step either side of the sampling loop: `vae.encode(sketch).latent_dist` of
 * app.py:107-109 (the sketch target of modules/pipeline.py:141-161) and `vae.decode(latents / 0.18215).sample` of
 * modules/pipeline.py:118 / :163-174 (decode_latents, decode_latents_L).  Weights: host fp32 tensors under their diffusers
 * names (encoder.*, decoder.*, quant_conv.*, post_quant_conv.*; mid-block attention as group_norm / query / key / value /
 * proj_attn).  All tensors NCHW fp32 device.
 *   encode: image [B, in_channels, H, W] (H, W multiples of 8) -> moments [B, 2 latent_channels, H/8, W/8] = (mean | logvar);
 *           the caller forms DiagonalGaussianDistribution (sample = mean + exp(0.5 clamp(logvar, -30, 20)) * noise).
 *   decode: latents [B, latent_channels, h, w] -> image [B, out_channels, 8h, 8w].
 * --------------------------------------------------------------------------------------------- */
typedef struct s2i_vae_config {
    int in_channels, out_channels, latent_channels;
    int block_out_channels[4];
    int layers_per_block;
} s2i_vae_config;
typedef struct s2i_vae s2i_vae;
int s2i_vae_create(const s2i_vae_config* cfg, s2i_vae** out);
void s2i_vae_destroy(s2i_vae* v);
int s2i_vae_load(s2i_vae* v, int n, const char* const* names, const float* const* host_ptrs, const int* ndims,
                 const long long* shapes);
int s2i_vae_encode(s2i_vae* v, const float* image, int B, int H, int W, float* moments, void* cuda_stream);
int s2i_vae_decode(s2i_vae* v, const float* latents, int B, int h, int w, float* image, void* cuda_stream);
long long s2i_vae_arena_bytes(s2i_vae* v);

/* ---------------------------------------------------------------------------------------------
 * Latent Guidance Predictor: replaces LatentEdgePredictor.forward (modules/latent_predictor.py:37-45), the
 * resize + concat of the taps (modules/pipeline.py:145-151), the edge loss (:155-157) and the LGP part of
 * autograd.grad (:159).  State-dict keys: layers.{0,3,6,9,12}.{weight,bias},
 * layers.{2,5,8,11}.{weight,bias,running_mean,running_var}.  Batches hold (uncond, cond) pairs; BatchNorm
 * statistics are per pair (train) or the running ones (eval).
 * --------------------------------------------------------------------------------------------- */
typedef struct s2i_lgp s2i_lgp;

int s2i_lgp_create(int input_dim, int output_dim, int num_pos_layers, s2i_lgp** out);
void s2i_lgp_destroy(s2i_lgp* l);
int s2i_lgp_load(s2i_lgp* l, int n, const char* const* names, const float* const* host_ptrs, const int* ndims,
                 const long long* shapes);
/* taps[k]: NHWC fp32 device [B][sizes[k]][sizes[k]][channels[k]]; noise: NCHW fp32 [B/2][4][L][L];
 * the noise-level input is sigma * noise for both halves of a pair (modules/pipeline.py:152-153) */
int s2i_lgp_forward_taps(s2i_lgp* l, const float* const* taps, const int* sizes, const int* channels, int B, int L,
                         const float* noise, float sigma, int train, void* cuda_stream);
/* emulate != 0 (default): round the back-propagated gradients like the reference's unscaled fp16 autograd
 * (latent_predictor.py:43 casts to fp16, so torch's backward runs in fp16); 0: loss-scaled, no extra rounding */
int s2i_lgp_set_grad_rounding(s2i_lgp* l, int emulate);
/* x: NCHW fp32 [B][input_dim-4-4P][L][L] (already resized + concatenated), t: NCHW fp32 [B][4][L][L] */
int s2i_lgp_forward_nchw(s2i_lgp* l, const float* x, const float* t, int B, int L, int train, void* cuda_stream);
/* out_rows: fp32 device [(b w h)][output_dim] -- the reference's row order (latent_predictor.py:43) */
int s2i_lgp_output(s2i_lgp* l, float* out_rows, void* cuda_stream);
/* target: NCHW fp32 [B/2][output_dim][L][L]; tap_grads[k]: NHWC fp32 like tap k (scaled by *grad_scale);
 * loss: device float [B/2] (MSE on the cond half, modules/pipeline.py:157) */
int s2i_lgp_loss_backward(s2i_lgp* l, const float* target, float* const* tap_grads, float* loss, float* grad_scale,
                          void* cuda_stream);
/* Same loss; tap_grads[k] holds only the B/2 cond samples' gradients (sample s = batch entry 2s+1), which is all
 * modules/pipeline.py:159 keeps.  The BatchNorm backward still covers both halves. */
int s2i_lgp_loss_backward_cond(s2i_lgp* l, const float* target, float* const* tap_grads, float* loss, float* grad_scale,
                               void* cuda_stream);

/* LGP training step (reference trainer.py:228-252: UNet forward on noised latents, taps resized and
 * concatenated, LatentEdgePredictor forward, MSE against the sketch latents, backward, optimizer step).
 *   forward_taps_batch: LatentEdgePredictor.forward as the trainer calls it -- B latents (no CFG pairs), BatchNorm batch
 *     statistics over all B * L * L rows, noise_level NCHW fp32 [B][4][L][L] = sqrt(1 - alpha_bar_t[b]) * noise[b] (trainer.py:197-204).
 *   train_step: loss = mean((LGP - target)^2) over [B][output_dim][L][L] (device float[1]); gradients of every Linear weight /
 *     bias and BatchNorm weight / bias (tcgen05 wgrad GEMMs, fixed-order sums); one AdamW update (torch.optim.AdamW's rule;
 *     the 8-bit state quantisation of the reference's bitsandbytes AdamW8bit is not reproduced) of the fp32 masters, after which
 *     every forward uses the new weights.  step counts from 1.
 *   get_param: a parameter's fp32 master by its state-dict name ("layers.0.weight", "layers.2.bias", ...) to host memory. */
int s2i_lgp_forward_taps_batch(s2i_lgp* l, const float* const* taps, const int* sizes, const int* channels, int B, int L,
                               const float* noise_level, void* cuda_stream);
int s2i_lgp_train_step(s2i_lgp* l, const float* target, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                       float* loss, void* cuda_stream);
int s2i_lgp_get_param(s2i_lgp* l, const char* name, float* host, long long n);

/* ---------------------------------------------------------------------------------------------
 * CFG combine + DDIM step (modules/pipeline.py:100-104; diffusers DDIMScheduler.step, eta = 0) and the
 * norm-ratio guidance update (modules/pipeline.py:160-161).  eps / dx: [2S][n] ordered (uncond_s, cond_s).
 * prediction: 0 = epsilon, 1 = v_prediction.
 * --------------------------------------------------------------------------------------------- */
int s2i_cfg_ddim_step(const float* latents, const float* eps, int S, int n, float guidance_scale, float sqrt_one_minus_a_t,
                      float sqrt_a_t, float sqrt_a_prev, float sqrt_one_minus_a_prev, int prediction, float* out,
                      void* cuda_stream);
/* CFG combine + one DPM-Solver++(2M, midpoint) update: diffusers DPMSolverMultistepScheduler.step as the reference's demo
 * configures it (app.py:14-25; called at modules/pipeline.py:104).  The host passes the step's fp32 scalars
 *   alpha_t, sigma_t (current timestep), c_x = sigma_prev / sigma_t, c_m0 = alpha_prev (exp(-h) - 1), c_d1 = 0.5 c_m0,
 *   inv_r0 = 1 / r0,  h = lambda_prev - lambda_t, r0 = (lambda_t - lambda_before) / h, lambda = log alpha - log sigma;
 *   m0 = x0 prediction; order 1: out = c_x x - c_m0 m0;  order 2: out = (c_x x - c_m0 m0) - c_d1 (inv_r0 (m0 - m1)).
 * x0_history [S][n]: the previous step's x0 prediction on entry (read when order == 2), this step's on return. */
int s2i_cfg_dpmpp_step(const float* latents, const float* eps, float* x0_history, int S, int n, float guidance_scale,
                       float alpha_t, float sigma_t, float c_x, float c_m0, float c_d1, float inv_r0, int order, int prediction,
                       float* out, void* cuda_stream);
/* x_new += beta * ||[x_old,x_old] - x_new||_F / ||g||_F * g,  g = -dx[cond]; scratch: device double [S][2] */
int s2i_guidance_update(const float* x_old, float* x_new, const float* dx, int S, int n, float beta, double* scratch,
                        void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * One whole denoising step on the device = the loop body of AntiGradientPipeline.__call__
 * (modules/pipeline.py:83-115) for S independent samples (SURVEY Q1 batch semantics).
 * latents [S][4][L][L] in/out; noise = the initial latents (:75); ctx [2S][ctx_len][D] ordered (uncond, cond);
 * target [S][4][L][L] or NULL (guidance skipped, :142-143); sigma = sqrt(1 - alpha_bar_t) (:133).
 * --------------------------------------------------------------------------------------------- */
typedef struct s2i_sampler s2i_sampler;
int s2i_sampler_create(s2i_unet* u, s2i_lgp* l, s2i_sampler** out);
void s2i_sampler_destroy(s2i_sampler* s);
/* The sampler reuses the text context's cross-attention K/V projections from step to step.  Call this whenever the VALUES
 * behind `ctx` change (a new prompt / image); a new ctx pointer or sample count is noticed by itself. */
int s2i_sampler_context_changed(s2i_sampler* s);
int s2i_sampler_step(s2i_sampler* s, float* latents, const float* noise, const float* ctx, const float* target, int S,
                     int L, float t, float guidance_scale, float sqrt_a_t, float sqrt_one_minus_a_t, float sqrt_a_prev,
                     float sqrt_one_minus_a_prev, int prediction, int guided, float sigma, float beta, int lgp_train,
                     float* loss_out, void* cuda_stream);
/* The same loop body when `self.scheduler` is the demo's DPM-Solver++(2M) scheduler (app.py:14-25): the scheduler.step of
 * modules/pipeline.py:104 becomes s2i_cfg_dpmpp_step (scalars as documented there; the LGP noise level of :133 is sigma_t);
 * the guidance update of :109 edits the latents AFTER the solver update, exactly like the reference (the multistep history
 * keeps the unedited x0 predictions).  x0_history [S][4][L][L]: caller-owned, carried from step to step of an image. */
int s2i_sampler_step_dpmpp(s2i_sampler* s, float* latents, const float* noise, const float* ctx, const float* target,
                           float* x0_history, int S, int L, float t, float guidance_scale, float alpha_t, float sigma_t,
                           float c_x, float c_m0, float c_d1, float inv_r0, int order, int prediction, int guided, float beta,
                           int lgp_train, float* loss_out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* S2I_H_ */
