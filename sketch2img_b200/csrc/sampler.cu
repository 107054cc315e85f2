// One denoising step of the sketch-guided sampler, entirely on the device (reference loop body:
// modules/pipeline.py:83-115): CFG-doubled UNet forward -> CFG combine + DDIM step -> [guided steps] LGP forward on the
// 9 taps, edge loss, backward through LGP and UNet to x_in, norm-ratio gradient update.
#include "../../include/s2i.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "lgp.cuh"
#include "unet.cuh"

struct s2i_unet {
    s2i::UNet* impl;
};
struct s2i_lgp {
    s2i::LGP* impl;
};

namespace s2i {

class Sampler {
  public:
    UNet* unet;
    LGP* lgp;
    Sampler(UNet* u, LGP* l) : unet(u), lgp(l) {}
    ~Sampler() {
        if (buf_) cudaFree(buf_);
    }

    int step(float* latents, const float* noise, const float* ctx, const float* target, int S, int L, float t,
             float guidance, float sa_t, float sb_t, float sa_p, float sb_p, int prediction, int guided, float sigma,
             float beta, int lgp_train, float* loss_out, cudaStream_t st) {
        const int C = unet->cfg.in_ch;
        const int n = C * L * L;
        const int B = 2 * S;
        // scratch: x_in [B][n] | eps [B][n] | dx [B][n] | x_new [S][n] | loss [S] | norms double [S][2] | tap grads
        const int* boc = unet->cfg.boc;
        const int tapS[9] = {L / 2, L / 4, L / 8, L / 8, L / 8, L / 8, L / 4, L / 2, L};
        const int tapC[9] = {boc[0], boc[1], boc[2], boc[3], boc[3], boc[3], boc[3], boc[2], boc[1]};
        size_t need = (size_t)(3 * B + S) * n * sizeof(float) + (size_t)S * 20 + 16 * 256;
        for (int k = 0; k < 9; ++k) need += (size_t)B * tapS[k] * tapS[k] * tapC[k] * 4 + 256;
        S2I_TRY(ensure(need));
        char* p = buf_;
        auto take = [&](size_t bytes) {
            char* q = p;
            p += (bytes + 255) & ~size_t(255);
            return q;
        };
        float* x_in = (float*)take((size_t)B * n * 4);
        float* eps = (float*)take((size_t)B * n * 4);
        float* dx = (float*)take((size_t)B * n * 4);
        float* x_new = (float*)take((size_t)S * n * 4);
        float* loss = (float*)take((size_t)S * 4);
        double* norms = (double*)take((size_t)S * 16);

        // x_in = cat([latents] * 2) per sample, ordered (uncond_s, cond_s)   (pipeline.py:85)
        for (int s = 0; s < S; ++s) {
            S2I_CUDA(cudaMemcpyAsync(x_in + (size_t)(2 * s) * n, latents + (size_t)s * n, n * 4, cudaMemcpyDeviceToDevice, st));
            S2I_CUDA(cudaMemcpyAsync(x_in + (size_t)(2 * s + 1) * n, latents + (size_t)s * n, n * 4, cudaMemcpyDeviceToDevice, st));
        }
        const bool do_guide = guided && target != nullptr && lgp != nullptr;
        S2I_TRY(unet->forward(x_in, B, L, L, t, ctx, eps, do_guide, st));                          // :96
        S2I_TRY(cfg_ddim_step(latents, eps, S, n, guidance, sb_t, sa_t, sa_p, sb_p, prediction, x_new, st));   // :100-104
        if (do_guide) {
            // taps -> LGP -> edge loss -> tap gradients   (:145-159, LGP part)
            LgpTap taps[9];
            for (int k = 0; k < 9; ++k) {
                const F32& tp = unet->taps[k];
                if (tp.H != tp.W) return set_error(S2I_ERR_ARG, "guided sampling needs square latents (pipeline.py:147)");
                taps[k] = LgpTap{tp.p, tp.H, tp.C};
                if (tp.H != tapS[k] || tp.C != tapC[k]) return set_error(S2I_ERR_STATE, "sampler: unexpected tap %d geometry", k);
            }
            float* tg[9];
            for (int k = 0; k < 9; ++k) tg[k] = (float*)take((size_t)unet->taps[k].rows() * unet->taps[k].C * 4);
            S2I_TRY(lgp->forward(taps, B, L, noise, sigma, lgp_train != 0, st));
            S2I_TRY(lgp->loss_backward(target, tg, loss, st));
            S2I_TRY(unet->backward(tg, dx, st));                                                     // :159 (UNet part)
            S2I_TRY(guidance_update(latents, x_new, dx, S, n, beta, norms, st));                     // :160-161
            if (loss_out) S2I_CUDA(cudaMemcpyAsync(loss_out, loss, S * 4, cudaMemcpyDeviceToDevice, st));
        }
        S2I_CUDA(cudaMemcpyAsync(latents, x_new, (size_t)S * n * 4, cudaMemcpyDeviceToDevice, st));
        return 0;
    }

  private:
    char* buf_ = nullptr;
    size_t cap_ = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap_) return 0;
        if (buf_) {
            cudaDeviceSynchronize();
            cudaFree(buf_);
        }
        buf_ = nullptr;
        cap_ = 0;
        void* q = nullptr;
        if (cudaMalloc(&q, bytes) != cudaSuccess) {
            cudaGetLastError();
            return set_error(S2I_ERR_OOM, "sampler: cannot allocate %.2f GB scratch", bytes / 1e9);
        }
        buf_ = static_cast<char*>(q);
        cap_ = bytes;
        return 0;
    }
};

}  // namespace s2i

struct s2i_sampler {
    s2i::Sampler* impl;
};

static std::map<std::string, s2i::HostParam> collect(int n, const char* const* names, const float* const* host_ptrs,
                                                     const int* ndims, const long long* shapes) {
    std::map<std::string, s2i::HostParam> params;
    for (int i = 0; i < n; ++i) {
        s2i::HostParam hp;
        hp.data = host_ptrs[i];
        for (int k = 0; k < ndims[i]; ++k) hp.shape.push_back((long)shapes[i * 4 + k]);
        params[names[i]] = hp;
    }
    return params;
}

extern "C" {

int s2i_lgp_create(int input_dim, int output_dim, int num_pos_layers, s2i_lgp** out) {
    if (!out) return s2i::set_error(S2I_ERR_ARG, "s2i_lgp_create: null out");
    if (input_dim <= 4 + 4 * num_pos_layers) return s2i::set_error(S2I_ERR_ARG, "s2i_lgp_create: input_dim too small");
    *out = new s2i_lgp{new s2i::LGP(input_dim, output_dim, num_pos_layers)};
    return 0;
}

void s2i_lgp_destroy(s2i_lgp* l) {
    if (!l) return;
    delete l->impl;
    delete l;
}

int s2i_lgp_load(s2i_lgp* l, int n, const char* const* names, const float* const* host_ptrs, const int* ndims,
                 const long long* shapes) {
    if (!l) return s2i::set_error(S2I_ERR_ARG, "s2i_lgp_load: null handle");
    return l->impl->load(collect(n, names, host_ptrs, ndims, shapes));
}

int s2i_lgp_set_grad_rounding(s2i_lgp* l, int emulate) {
    if (!l) return s2i::set_error(S2I_ERR_ARG, "s2i_lgp_set_grad_rounding: null handle");
    l->impl->emulate_fp16_grad = emulate != 0;
    return 0;
}

int s2i_lgp_forward_taps(s2i_lgp* l, const float* const* taps, const int* sizes, const int* channels, int B, int L,
                         const float* noise, float sigma, int train, void* cuda_stream) {
    if (!l || !taps || !sizes || !channels || !noise) return s2i::set_error(S2I_ERR_ARG, "s2i_lgp_forward_taps: null argument");
    s2i::LgpTap tp[9];
    for (int k = 0; k < 9; ++k) tp[k] = s2i::LgpTap{taps[k], sizes[k], channels[k]};
    return l->impl->forward(tp, B, L, noise, sigma, train != 0, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_lgp_forward_nchw(s2i_lgp* l, const float* x, const float* t, int B, int L, int train, void* cuda_stream) {
    if (!l || !x || !t) return s2i::set_error(S2I_ERR_ARG, "s2i_lgp_forward_nchw: null argument");
    return l->impl->forward_nchw(x, t, B, L, train != 0, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_lgp_output(s2i_lgp* l, float* out_rows, void* cuda_stream) {
    if (!l || !out_rows) return s2i::set_error(S2I_ERR_ARG, "s2i_lgp_output: null argument");
    return l->impl->export_output(out_rows, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_lgp_loss_backward(s2i_lgp* l, const float* target, float* const* tap_grads, float* loss, float* grad_scale,
                          void* cuda_stream) {
    if (!l || !target || !tap_grads || !loss) return s2i::set_error(S2I_ERR_ARG, "s2i_lgp_loss_backward: null argument");
    int rc = l->impl->loss_backward(target, tap_grads, loss, static_cast<cudaStream_t>(cuda_stream));
    if (rc == 0 && grad_scale) *grad_scale = l->impl->grad_scale();
    return rc;
}

int s2i_cfg_ddim_step(const float* latents, const float* eps, int S, int n, float guidance_scale, float sqrt_one_minus_a_t,
                      float sqrt_a_t, float sqrt_a_prev, float sqrt_one_minus_a_prev, int prediction, float* out,
                      void* cuda_stream) {
    if (!latents || !eps || !out) return s2i::set_error(S2I_ERR_ARG, "s2i_cfg_ddim_step: null argument");
    return s2i::cfg_ddim_step(latents, eps, S, n, guidance_scale, sqrt_one_minus_a_t, sqrt_a_t, sqrt_a_prev,
                              sqrt_one_minus_a_prev, prediction, out, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_guidance_update(const float* x_old, float* x_new, const float* dx, int S, int n, float beta, double* scratch,
                        void* cuda_stream) {
    if (!x_old || !x_new || !dx || !scratch) return s2i::set_error(S2I_ERR_ARG, "s2i_guidance_update: null argument");
    return s2i::guidance_update(x_old, x_new, dx, S, n, beta, scratch, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_sampler_create(s2i_unet* u, s2i_lgp* l, s2i_sampler** out) {
    if (!u || !out) return s2i::set_error(S2I_ERR_ARG, "s2i_sampler_create: null argument");
    *out = new s2i_sampler{new s2i::Sampler(u->impl, l ? l->impl : nullptr)};
    return 0;
}

void s2i_sampler_destroy(s2i_sampler* s) {
    if (!s) return;
    delete s->impl;
    delete s;
}

int s2i_sampler_step(s2i_sampler* s, float* latents, const float* noise, const float* ctx, const float* target, int S,
                     int L, float t, float guidance_scale, float sqrt_a_t, float sqrt_one_minus_a_t, float sqrt_a_prev,
                     float sqrt_one_minus_a_prev, int prediction, int guided, float sigma, float beta, int lgp_train,
                     float* loss_out, void* cuda_stream) {
    if (!s || !latents || !ctx) return s2i::set_error(S2I_ERR_ARG, "s2i_sampler_step: null argument");
    if (guided && target && !noise) return s2i::set_error(S2I_ERR_ARG, "s2i_sampler_step: guided step needs the initial noise");
    return s->impl->step(latents, noise, ctx, target, S, L, t, guidance_scale, sqrt_a_t, sqrt_one_minus_a_t, sqrt_a_prev,
                         sqrt_one_minus_a_prev, prediction, guided, sigma, beta, lgp_train, loss_out,
                         static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
