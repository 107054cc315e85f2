// extern "C" surface of libs2i (declarations: include/s2i.h).
#include "../../include/s2i.h"
#include "common.cuh"
#include "attn.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "unet.cuh"

#include <cstdlib>

struct s2i_unet {
    s2i::UNet* impl;
};

extern "C" {

const char* s2i_last_error(void) { return s2i::last_error(); }
long long s2i_launch_count(void) { return s2i::g_launches; }
int s2i_profile_begin(void* cuda_stream) { return s2i::prof_begin(static_cast<cudaStream_t>(cuda_stream)); }
int s2i_profile_end(char* report, int capacity) {
    if (!report || capacity <= 0) return s2i::set_error(S2I_ERR_ARG, "s2i_profile_end: no report buffer");
    return s2i::prof_end(report, capacity);
}

int s2i_profile_set_peaks(double tflops, double hbm_gbs) {
    if (tflops < 0 || hbm_gbs < 0) return s2i::set_error(S2I_ERR_ARG, "s2i_profile_set_peaks: negative peak");
    s2i::prof_set_peaks(tflops, hbm_gbs);
    return 0;
}

int s2i_gemm_set_tma_epilogue(int on) {
    s2i::gemm_set_tma_epilogue(on);
    return 0;
}

int s2i_gemm_force_msub(int msub) {
    s2i::gemm_force_msub(msub);
    return 0;
}

int s2i_gemm_set_pair(int mode) {
    s2i::gemm_set_pair(mode);
    return 0;
}

int s2i_gemm_set_trace(void* device_buf) {
    s2i::gemm_set_trace(static_cast<unsigned long long*>(device_buf));
    return 0;
}

int s2i_gemm(const s2i_gemm_desc* d, void* cuda_stream) {
    if (!d) return s2i::set_error(S2I_ERR_ARG, "s2i_gemm: null descriptor");
    return s2i::gemm_launch(s2i::GemmDesc(*d), static_cast<cudaStream_t>(cuda_stream));
}


int s2i_groupnorm_forward(const float* x, long long ldx, int B, int HW, int C, const float* gamma, const float* beta,
                          float eps, int silu, void* out16, long long ld16, void* raw16, long long ldraw, void* stats,
                          void* cuda_stream) {
    if (!x || !gamma || !beta || !out16 || !stats) return s2i::set_error(S2I_ERR_ARG, "s2i_groupnorm_forward: null argument");
    if (B < 1 || HW < 1 || C < 32) return s2i::set_error(S2I_ERR_ARG, "s2i_groupnorm_forward: empty tensor");
    return s2i::gn_forward(x, ldx, B, HW, C, static_cast<double*>(stats), gamma, beta, eps, silu, out16, ld16, raw16, ldraw,
                           static_cast<cudaStream_t>(cuda_stream));
}

int s2i_groupnorm_forward_colstat(const float* x, long long ldx, int B, int HW, int C, const float* colstat,
                                  long long colstat_ld, int colstat_cap, int colstat_bps, const float* gamma,
                                  const float* beta, float eps, int silu, void* out16, long long ld16, void* raw16,
                                  long long ldraw, void* stats, void* cuda_stream) {
    if (!x || !colstat || !gamma || !beta || !out16 || !stats)
        return s2i::set_error(S2I_ERR_ARG, "s2i_groupnorm_forward_colstat: null argument");
    if (B < 1 || HW < 1 || C < 32 || colstat_bps < 1 || colstat_bps > colstat_cap || colstat_ld < C)
        return s2i::set_error(S2I_ERR_ARG, "s2i_groupnorm_forward_colstat: bad shape");
    if (!s2i::gn_norm_supported(C)) return s2i::set_error(S2I_ERR_ARG, "s2i_groupnorm_forward_colstat: unsupported channel count %d", C);
    s2i::GnStatSrc src;
    src.p = colstat; src.cap = colstat_cap; src.bps = colstat_bps; src.ld = colstat_ld; src.c0 = 0; src.c1 = C;
    return s2i::gn_norm(x, ldx, B, HW, C, &src, 1, static_cast<double*>(stats), gamma, beta, eps, silu, out16, ld16, raw16, ldraw,
                        static_cast<cudaStream_t>(cuda_stream));
}

int s2i_groupnorm_backward(const float* dy, long long ldd, const float* x, long long ldx, int B, int HW, int C,
                           const float* gamma, const float* beta, float eps, int silu, const void* fwd_stats,
                           void* bwd_stats, const float* add, long long ldadd, float* dx32, long long ld32, void* dx16,
                           long long ld16, void* cuda_stream) {
    if (!dy || !x || !gamma || !beta || !fwd_stats || !bwd_stats || (!dx32 && !dx16))
        return s2i::set_error(S2I_ERR_ARG, "s2i_groupnorm_backward: null argument");
    if (B < 1 || HW < 1 || C < 32) return s2i::set_error(S2I_ERR_ARG, "s2i_groupnorm_backward: empty tensor");
    return s2i::gn_backward(dy, ldd, x, ldx, B, HW, C, static_cast<const double*>(fwd_stats), static_cast<double*>(bwd_stats),
                            gamma, beta, eps, silu, add, ldadd, dx32, ld32, dx16, ld16, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_attention(const void* q, long long ldq, int q_c0, const void* kv, long long ldkv, int k_c0, int v_c0, int B,
                  int heads, int Nq, int Nk, int dp, int d_true, float scale, void* out, long long ldo, float* lse,
                  void* cuda_stream) {
    if (!q || !kv || !out) return s2i::set_error(S2I_ERR_ARG, "s2i_attention: null argument");
    s2i::AttnDesc a;
    a.q = static_cast<const __half*>(q); a.ldq = ldq; a.q_c0 = q_c0;
    a.kv = static_cast<const __half*>(kv); a.ldkv = ldkv; a.k_c0 = k_c0; a.v_c0 = v_c0;
    a.B = B; a.heads = heads; a.Nq = Nq; a.Nk = Nk; a.dp = dp; a.d_true = d_true;
    a.scale = scale;
    a.out = static_cast<__half*>(out); a.ldo = ldo; a.lse = lse;
    return s2i::attn_fwd_launch(a, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_attention_backward(const void* q, long long ldq, int q_c0, const void* kv, long long ldkv, int k_c0, int v_c0,
                           const void* d_out, const void* out, long long ldo, const float* lse, float* delta_scratch, int B,
                           int heads, int Nq, int Nk, int dp, int d_true, float scale, void* dq, long long lddq, int dq_c0,
                           void* dkv, long long lddkv, int dk_c0, int dv_c0, void* cuda_stream) {
    if (!q || !kv || !d_out || !out || !lse || !delta_scratch || !dq)
        return s2i::set_error(S2I_ERR_ARG, "s2i_attention_backward: null argument");
    s2i::AttnBwdDesc a;
    a.q = static_cast<const __half*>(q); a.ldq = ldq; a.q_c0 = q_c0;
    a.kv = static_cast<const __half*>(kv); a.ldkv = ldkv; a.k_c0 = k_c0; a.v_c0 = v_c0;
    a.dO = static_cast<const __half*>(d_out); a.lddo = ldo;
    a.o = static_cast<const __half*>(out); a.ldo = ldo;
    a.lse = lse; a.delta = delta_scratch;
    a.B = B; a.heads = heads; a.Nq = Nq; a.Nk = Nk; a.dp = dp; a.d_true = d_true; a.scale = scale;
    a.dq = static_cast<__half*>(dq); a.lddq = lddq; a.dq_c0 = dq_c0;
    a.dk = a.dv = static_cast<__half*>(dkv); a.lddkv = lddkv; a.dk_c0 = dk_c0; a.dv_c0 = dv_c0;
    return s2i::attn_bwd_launch(a, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_unet_create(const s2i_unet_config* c, s2i_unet** out) {
    if (!c || !out) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_create: null argument");
    s2i::UNetConfig cfg;
    cfg.in_ch = c->in_channels;
    cfg.out_ch = c->out_channels;
    for (int i = 0; i < 4; ++i) {
        cfg.boc[i] = c->block_out_channels[i];
        cfg.heads[i] = c->num_heads[i];
        if (cfg.boc[i] % 64 != 0) return s2i::set_error(S2I_ERR_ARG, "block_out_channels must be multiples of 64");
        if (cfg.heads[i] <= 0 || cfg.boc[i] % cfg.heads[i] != 0) return s2i::set_error(S2I_ERR_ARG, "bad head count");
    }
    cfg.layers = c->layers_per_block;
    cfg.cross_dim = c->cross_attention_dim;
    cfg.sample_size = c->sample_size;
    cfg.ctx_len = c->ctx_len;
    if (cfg.cross_dim % 8 != 0) return s2i::set_error(S2I_ERR_ARG, "cross_attention_dim must be a multiple of 8");
    *out = new s2i_unet{new s2i::UNet(cfg)};
    if (const char* e = getenv("S2I_NO_FLASH")) (*out)->impl->use_flash_ = !(e[0] == '1');
    if (const char* e = getenv("S2I_GLU_FUSION")) (*out)->impl->fuse_glu_ = atoi(e);
    return 0;
}

int s2i_sketch_encoder_create(const s2i_unet_config* c, s2i_unet** out) {
    int rc = s2i_unet_create(c, out);
    if (rc != 0) return rc;
    (*out)->impl->cfg.encoder_only = true;
    return 0;
}

int s2i_sketch_encoder_forward(s2i_unet* u, const float* x, int B, int H, int W, float t, void* cuda_stream) {
    if (!u || !x) return s2i::set_error(S2I_ERR_ARG, "s2i_sketch_encoder_forward: null argument");
    if (!u->impl->cfg.encoder_only) return s2i::set_error(S2I_ERR_STATE, "s2i_sketch_encoder_forward: not a sketch encoder");
    return u->impl->forward(x, B, H, W, t, nullptr, nullptr, false, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_sketch_encoder_num_res_samples(s2i_unet* u) {
    if (!u) return s2i::set_error(S2I_ERR_ARG, "s2i_sketch_encoder_num_res_samples: null engine");
    return (int)u->impl->res_samples.size();
}

int s2i_sketch_encoder_res_sample(s2i_unet* u, int k, float** ptr, long long* pixel_stride, int* B, int* H, int* W, int* C) {
    if (!u || !ptr || !pixel_stride || !B || !H || !W || !C) return s2i::set_error(S2I_ERR_ARG, "s2i_sketch_encoder_res_sample: null argument");
    if (k < 0 || k >= (int)u->impl->res_samples.size())
        return s2i::set_error(S2I_ERR_ARG, "s2i_sketch_encoder_res_sample: index %d outside the %d maps of the last forward", k,
                              (int)u->impl->res_samples.size());
    const s2i::F32& t = u->impl->res_samples[k];
    *ptr = t.p; *pixel_stride = t.ld; *B = t.B; *H = t.H; *W = t.W; *C = t.C;
    return 0;
}

void s2i_unet_destroy(s2i_unet* u) {
    if (!u) return;
    delete u->impl;
    delete u;
}

int s2i_unet_load(s2i_unet* u, int n, const char* const* names, const float* const* host_ptrs, const int* ndims,
                  const long long* shapes) {
    if (!u) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_load: null engine");
    std::map<std::string, s2i::HostParam> params;
    for (int i = 0; i < n; ++i) {
        s2i::HostParam hp;
        hp.data = host_ptrs[i];
        for (int k = 0; k < ndims[i]; ++k) hp.shape.push_back((long)shapes[i * 4 + k]);
        params[names[i]] = hp;
    }
    return u->impl->load(params);
}

int s2i_unet_forward(s2i_unet* u, const float* x, int B, int H, int W, float t, const float* ctx, float* eps,
                     int save_for_backward, void* cuda_stream) {
    if (!u || !x || !ctx || !eps) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_forward: null argument");
    return u->impl->forward(x, B, H, W, t, ctx, eps, save_for_backward != 0, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_unet_tap(s2i_unet* u, int k, float** ptr, int* B, int* H, int* W, int* C) {
    if (!u || k < 0 || k >= 9) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_tap: bad tap index");
    const s2i::F32& t = u->impl->taps[k];
    if (!t.p) return s2i::set_error(S2I_ERR_STATE, "s2i_unet_tap: no forward yet");
    *ptr = t.p; *B = t.B; *H = t.H; *W = t.W; *C = t.C;
    return 0;
}

int s2i_unet_tap_stride(s2i_unet* u, int k, long long* pixel_stride) {
    if (!u || k < 0 || k >= 9 || !pixel_stride) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_tap_stride: bad argument");
    const s2i::F32& t = u->impl->taps[k];
    if (!t.p) return s2i::set_error(S2I_ERR_STATE, "s2i_unet_tap_stride: no forward yet");
    *pixel_stride = t.ld;
    return 0;
}

int s2i_unet_backward(s2i_unet* u, float* const* tap_grads, float* dx, void* cuda_stream) {
    if (!u || !tap_grads || !dx) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_backward: null argument");
    return u->impl->backward(tap_grads, dx, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_unet_backward_samples(s2i_unet* u, float* const* tap_grads, float* dx, int b0, int nb, void* cuda_stream) {
    if (!u || !tap_grads || !dx) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_backward_samples: null argument");
    if (nb < 1) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_backward_samples: nb must be positive");
    return u->impl->backward(tap_grads, dx, static_cast<cudaStream_t>(cuda_stream), b0, nb);
}

int s2i_unet_load_sat(s2i_unet* u, int n, const char* const* names, const float* const* host_ptrs, const int* ndims,
                      const long long* shapes) {
    if (!u || !names || !host_ptrs || !ndims || !shapes) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_load_sat: null argument");
    std::map<std::string, s2i::HostParam> params;
    for (int i = 0; i < n; ++i) {
        s2i::HostParam hp;
        hp.data = host_ptrs[i];
        for (int k = 0; k < ndims[i]; ++k) hp.shape.push_back((long)shapes[i * 4 + k]);
        params[names[i]] = hp;
    }
    return u->impl->load_sat(params);
}

int s2i_unet_set_sat_feature(s2i_unet* u, const char* block_path, const float* feature_nchw, int B, int C, int H, int W,
                             void* cuda_stream) {
    if (!u || !block_path) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_set_sat_feature: null argument");
    return u->impl->set_sat_feature(block_path, feature_nchw, B, C, H, W, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_unet_set_sat_scale(s2i_unet* u, float scale, void* cuda_stream) {
    if (!u) return s2i::set_error(S2I_ERR_ARG, "s2i_unet_set_sat_scale: null engine");
    return u->impl->set_sat_scale(scale, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_unet_debug(s2i_unet* u, int enable) {
    if (!u) return s2i::set_error(S2I_ERR_ARG, "null engine");
    u->impl->keep_debug = enable != 0;
    return 0;
}

int s2i_unet_debug_get(s2i_unet* u, const char* name, float** ptr, long long* ld, int* B, int* H, int* W, int* C) {
    if (!u) return s2i::set_error(S2I_ERR_ARG, "null engine");
    auto it = u->impl->debug.find(name);
    if (it == u->impl->debug.end()) return s2i::set_error(S2I_ERR_ARG, "no debug tensor named %s", name);
    *ptr = it->second.p; *ld = it->second.ld; *B = it->second.B; *H = it->second.H; *W = it->second.W; *C = it->second.C;
    return 0;
}

long long s2i_unet_arena_bytes(s2i_unet* u) { return u ? (long long)u->impl->arena_bytes() : 0; }

}  // extern "C"
