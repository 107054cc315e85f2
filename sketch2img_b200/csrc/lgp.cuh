// Latent Guidance Predictor: per-pixel MLP over the concatenated, bilinearly resized UNet taps
// (reference: modules/latent_predictor.py:9-45 forward; modules/pipeline.py:145-159 feature build, edge loss and
// autograd backward to the taps).  BatchNorm1d runs with per-sample batch statistics (train mode)
// or running statistics (eval mode).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "unet.cuh"

namespace s2i {

struct LgpTap {
    const float* p;   // NHWC fp32 [B][S][S][C], pixel stride ld (0 = C: dense)
    int S, C;
    long ld = 0;
};

class LGP {
  public:
    LGP(int input_dim, int output_dim, int num_pos_layers);
    ~LGP();
    int load(const std::map<std::string, HostParam>& params);

    // Forward on B = 2*samples latents of side L.  noise: NCHW fp32 [samples,4,L,L]; lvl = sigma * noise is the
    // "noise level" input shared by both CFG halves (pipeline.py:152-153).  Rows are ordered (b, h, w).
    // out16: fp16 [B*L*L][8] (first output_dim columns valid).
    // dsigma (optional, device): overrides `sigma` at run time (graph-replayed steps).
    // taps_sample_major: the taps' batch is ordered [uncond_0.., cond_0..] (the sampler's internal order) instead of
    // (uncond_s, cond_s) pairs; the LGP's own rows stay in pair order either way.
    // groups: 0 = BatchNorm statistics per (uncond, cond) pair, noise shared by a pair (the sampling path); 1 = statistics over
    // the whole batch and one noise map per latent (LatentEdgePredictor.forward as the trainer calls it, trainer.py:245)
    int forward(const LgpTap taps[9], int B, int L, const float* noise, float sigma, bool train, cudaStream_t st,
                const float* dsigma = nullptr, bool taps_sample_major = false, int groups = 0);
    // One optimisation step of the LGP on the batch of the last forward(groups = 1, train) (reference: trainer.py:245-252):
    // loss = mse(LGP(features), target) over every latent, weight / bias / BatchNorm-affine gradients (tcgen05 wgrad GEMMs,
    // fixed-order column sums), AdamW update of the fp32 master parameters, fp16 operand copies re-packed.
    // target: NCHW fp32 [B][output_dim][L][L]; loss: device float[1]; step: 1-based count for the bias correction.
    int train_step(const float* target, float lr, float beta1, float beta2, float eps, float weight_decay, int step, float* loss,
                   cudaStream_t st);
    // copy a parameter's fp32 master (state-dict name, e.g. "layers.3.weight") to the host; n = element count.
    // "grad.layers.N.weight": the weight gradient of the last train_step, multiplied by grad_scale() (tests)
    int get_param(const std::string& name, float* host, size_t n);
    // Edge loss on the cond half + backward to the taps.  target: NCHW fp32 [samples,4,L,L].
    // tap_grads[k]: NHWC fp32 like tap k, multiplied by grad_scale(); loss: device float [samples].
    // cond_only: only the cond samples' tap gradients are produced (tap_grads[k] then holds `samples` maps, sample s =
    // gradient of batch entry 2s+1) -- the sampler discards the uncond half of the latent gradient (pipeline.py:159) and
    // the UNet is per-sample, so the uncond taps' gradients are never needed.  The BatchNorm backward still runs over
    // both halves (train-mode statistics couple them).
    int loss_backward(const float* target, float* const tap_grads[9], float* loss, cudaStream_t st, bool cond_only = false);
    float grad_scale() const { return gscale_; }
    const __half* output() const { return out16_; }
    // out_rows: fp32 [(b w h)][output_dim] in the reference's row order (latent_predictor.py:43)
    int export_output(float* out_rows, cudaStream_t st);
    // features from an already concatenated NCHW fp32 tensor x [B][input_dim-4-4*P][L][L] and t [B][4][L][L]
    int forward_nchw(const float* x, const float* t, int B, int L, bool train, cudaStream_t st);

    int input_dim() const { return D_; }
    // true (default): gradients are rounded like the reference's unscaled fp16 autograd (latent_predictor.py:43 casts
    // the activations to fp16, so torch back-propagates in fp16); false: loss-scaled gradients, no extra rounding.
    bool emulate_fp16_grad = true;

  private:
    int D_, O_, P_;
    long ldX_ = 0;
    int widths_[6];
    Lin lin_[5];
    Norm bn_[4];
    float* bn_rm_[4] = {nullptr, nullptr, nullptr, nullptr};
    float* bn_rv_[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<void*> owned_;
    bool loaded_ = false;

    // per-call state
    char* buf_ = nullptr;
    size_t buf_cap_ = 0;
    int B_ = 0, L_ = 0;
    bool train_ = true;
    LgpTap taps_[9];
    __half* X_ = nullptr;        // [rows][ldX]
    __half* h_[4] = {};          // ReLU outputs (BN inputs)   [rows][w]
    __half* a_[4] = {};          // BN outputs (next GEMM operand)
    double* bsum_[4] = {};       // per (sample, column) sum / sumsq
    __half* out16_ = nullptr;    // [rows][8]
    __half* dout_ = nullptr;     // [rows][8] scaled loss gradient
    __half *dA_ = nullptr, *dB_ = nullptr;   // backward ping-pong [rows][512]
    double* bbsum_[4] = {};      // backward per (sample, column) sums
    float* mean_[4] = {};
    float* rstd_[4] = {};
    // training (train_step): fp32 masters of the Linear weights, their gradients and the AdamW moments of every parameter
    float* master_w_[5] = {};
    float* grad_w_[5] = {};
    struct Moment { float *m = nullptr, *v = nullptr; };
    Moment mom_w_[5], mom_b_[5], mom_g_[4], mom_beta_[4];
    float* red_part_ = nullptr;          // per-chunk partial sums of the BatchNorm reductions (summed in chunk order)
    float* loss_part_ = nullptr;         // per-block partial sums of the edge loss
    unsigned int* red_counter_ = nullptr;    // arrival counters: [S] BatchNorm reductions, then [S] loss
    bool have_fwd_ = false;
    int groups_ = 1;             // BatchNorm statistic groups (pairs on the sampling path, 1 for forward())
    float gscale_ = 1.f;
    float* interp_tmp_ = nullptr;    // [B][L][S][C] intermediate of the separable resize adjoint
    size_t interp_cap_ = 0;

    int ensure(size_t bytes);
    int mlp(cudaStream_t st);
};

// ---- scheduler / guidance update (modules/pipeline.py:100-104, :160-161; DDIM step per SURVEY A.6) -----------
// eps: [2*S][n] ordered (uncond_s, cond_s); latents: [S][n].  prediction: 0 = epsilon, 1 = v_prediction.
// dparams (optional, device float[4] = sb_t, sa_t, sa_p, sb_p): overrides the by-value coefficients (graph-replayed steps).
// sample_major: eps is ordered [uncond_0.., cond_0..] (the sampler's internal batch order) instead of pairs.
int cfg_ddim_step(const float* latents, const float* eps, int S, int n, float guidance, float sb_t, float sa_t,
                  float sa_p, float sb_p, int prediction, float* out, cudaStream_t st, const float* dparams = nullptr,
                  bool sample_major = false);
// CFG combine + DPM-Solver++(2M, midpoint) update, the scheduler of the reference's demo (app.py:14-25).  The host supplies
// the step's fp32 scalars: sigma_t / alpha_t of the current timestep, A = sigma_prev / sigma_t, Bc = alpha_prev (exp(-h) - 1),
// Cc = 0.5 Bc, R = 1 / r0.  x0_hist [S][n]: previous x0 prediction in (second order), this step's out.  order: 1 | 2.
// dparams (optional, device float[8] = sigma_t, alpha_t, A, Bc, -, Cc, R): overrides the by-value scalars.
int cfg_dpmpp_step(const float* latents, const float* eps, float* x0_hist, int S, int n, float guidance, float sigma_t,
                   float alpha_t, float A, float Bc, float Cc, float R, int prediction, int order, float* out, cudaStream_t st,
                   const float* dparams = nullptr, bool sample_major = false);
// x_new += beta * ||x_in - x_new||_F / ||g||_F * g with g = -dx[cond half]; norms per sample; x_in = [x_old, x_old].
// dx: [2*S][n], or [S][n] holding only the cond halves (dx_cond_only).  scratch: double [S][2] device.
int guidance_update(const float* x_old, float* x_new, const float* dx, int S, int n, float beta, double* scratch,
                    cudaStream_t st, bool dx_cond_only = false);

}  // namespace s2i
