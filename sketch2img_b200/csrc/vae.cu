// AutoencoderKL (the SD VAE) encode / decode on the engine's kernels: SURVEY 8f row f-1 -- the step either side of the sampling
// loop.  Reference call sites: /root/reference/app.py:107-109 (vae.encode(sketch).latent_dist.sample() * 0.18215 -> the sketch
// target of modules/pipeline.py:141-161) and modules/pipeline.py:118, :163-174 (vae.decode(latents / 0.18215).sample).
// diffusers' topology (un-vendored dependency, restated in oracle/diffusers_shim/diffusers/models/vae.py):
//   encoder: conv_in 3x3 -> 4 x [2 ResnetBlock2D (GroupNorm32 eps 1e-6, SiLU, no time embedding) + 3x3 stride-2 conv padded on
//            the bottom / right] -> mid (resnet, single-head attention, resnet) -> GroupNorm + SiLU -> conv_out 3x3 -> quant_conv 1x1
//   decoder: post_quant_conv 1x1 -> conv_in 3x3 -> mid -> 4 x [3 ResnetBlock2D + nearest 2x + 3x3 conv] -> GroupNorm + SiLU -> conv_out
// Every contraction runs on the tcgen05 + TMA implicit GEMM (gemm_tc.cu), GroupNorm on the cluster kernel, the mid-block
// attention (one head of 512 channels over (H/8)(W/8) tokens) on the GEMM + softmax path of UNet::attention.
#include "vae.cuh"

#include <cmath>

#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "loader.cuh"

namespace s2i {

namespace {

// moments[b][co][p] = bias[co] + sum_ci W[co][ci] * y[(b, p)][ci]   (quant_conv after the encoder: NHWC fp32 in, NCHW fp32 out)
__global__ void __launch_bounds__(256) pointwise_to_nchw_kernel(const float* __restrict__ y, long ldy, long B, long HW, int Ci, int Co,
                                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                                float* __restrict__ out) {
    pdl_wait();
    pdl_launch();
    const long total = B * Co * HW;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long p = idx % HW;
        const int co = (int)((idx / HW) % Co);
        const long b = idx / (HW * Co);
        const float* row = y + (b * HW + p) * ldy;
        float acc = bias[co];
        for (int ci = 0; ci < Ci; ++ci) acc = fmaf(w[co * Ci + ci], row[ci], acc);
        out[idx] = acc;
    }
}

// y[(b, p)][co] = bias[co] + sum_ci W[co][ci] * z[b][ci][p]   (post_quant_conv before the decoder: NCHW fp32 in, NHWC fp32 out)
__global__ void __launch_bounds__(256) pointwise_from_nchw_kernel(const float* __restrict__ z, long B, long HW, int Ci, int Co,
                                                                  const float* __restrict__ w, const float* __restrict__ bias,
                                                                  float* __restrict__ y, long ldy) {
    pdl_wait();
    pdl_launch();
    const long total = B * HW * ldy;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long pix = idx / ldy;
        const int co = (int)(idx - pix * ldy);
        const long b = pix / HW, p = pix - b * HW;
        float acc = 0.f;
        if (co < Co) {
            acc = bias[co];
            for (int ci = 0; ci < Ci; ++ci) acc = fmaf(w[co * Ci + ci], z[(b * Ci + ci) * HW + p], acc);
        }
        y[idx] = acc;
    }
}

inline unsigned grid1(long work) {
    long g = (work + 255) / 256;
    if (g < 1) g = 1;
    if (g > 148L * 32) g = 148L * 32;
    return (unsigned)g;
}

}  // namespace

VAE::VAE(const VaeConfig& c) : UNet(UNetConfig()), vcfg(c) {}

// ================================================================================================== loading
int VAE::load_vae(const std::map<std::string, HostParam>& params) {
    Loader L{params, owned_};
    auto fail = [&](const std::string& what) {
        if (L.staging) cudaFree(L.staging);
        return set_error(S2I_ERR_ARG, "vae load (%s): %s", what.c_str(), L.err.c_str());
    };
    const int* boc = vcfg.boc;
    for (int i = 0; i < 4; ++i)
        if (boc[i] % 64 != 0) return set_error(S2I_ERR_ARG, "vae: block_out_channels must be multiples of 64");
    if (vcfg.in_ch > 7 || vcfg.latent > 7) return set_error(S2I_ERR_ARG, "vae: at most 7 image / latent channels (im2col width 64)");

    auto load_res = [&](const std::string& pre, int Cin, int Cout) -> int {
        ResBlock r;
        r.Cin = Cin;
        r.Cout = Cout;
        r.temb_off = -1;                                   // no time embedding in the VAE's ResnetBlock2D
        if (!L.norm(pre + ".norm1", Cin, 1e-6f, r.n1) || !L.conv3(pre + ".conv1", Cout, Cin, r.c1, false) ||
            !L.norm(pre + ".norm2", Cout, 1e-6f, r.n2) || !L.conv3(pre + ".conv2", Cout, Cout, r.c2, false))
            return -1;
        r.has_sc = Cin != Cout;
        if (r.has_sc && !L.linear(pre + ".conv_shortcut", Cout, Cin, true, r.sc, false)) return -1;
        res_.push_back(r);
        return (int)res_.size() - 1;
    };
    auto load_attn = [&](const std::string& pre, int C, VAttn& a) -> bool {
        a.C = C;
        if (!L.norm(pre + ".group_norm", C, 1e-6f, a.gn)) return false;
        a.qkv.N = 3 * C;
        a.qkv.K = C;
        a.qkv.w = L.dmalloc<__half>((size_t)3 * C * C);
        a.qkv.b = L.dmalloc<float>((size_t)3 * C);
        if (!a.qkv.w || !a.qkv.b) return false;
        const char* names[3] = {".query", ".key", ".value"};
        for (int s = 0; s < 3; ++s) {
            if (!L.linear_into(pre + names[s] + ".weight", C, C, 0, 0, false, false, a.qkv.w, C, (long)s * C, nullptr, 0, 0, false))
                return false;
            const HostParam* hb = L.find(pre + names[s] + ".bias", C);
            if (!hb) return false;
            cudaMemcpy(a.qkv.b + (size_t)s * C, hb->data, C * sizeof(float), cudaMemcpyHostToDevice);
        }
        return L.linear(pre + ".proj_attn", C, C, true, a.proj, false);
    };
    auto conv_in = [&](const std::string& pre, int Co, int Ci, Lin& l) -> bool {      // as an im2col GEMM, K = 9 Ci padded to 64
        const HostParam* hp = L.find(pre + ".weight", (size_t)Co * Ci * 9);
        if (!hp) return false;
        const float* s = L.stage(hp, (size_t)Co * Ci * 9);
        l.N = Co;
        l.K = 64;
        l.w = L.dmalloc<__half>((size_t)Co * 64, true);
        if (!s || !l.w) return false;
        pack_conv_kernel<<<64, 256>>>(s, Co, Ci, 0, l.w, 64);
        if (cudaDeviceSynchronize() != cudaSuccess) return false;
        l.b = L.vec(pre + ".bias", Co);
        return l.b != nullptr;
    };

    // ---- encoder
    if (!conv_in("encoder.conv_in", boc[0], vcfg.in_ch, enc_in_)) return fail("encoder.conv_in");
    int ch = boc[0];
    for (int i = 0; i < 4; ++i) {
        const std::string pre = "encoder.down_blocks." + std::to_string(i);
        for (int j = 0; j < vcfg.layers; ++j) {
            const int r = load_res(pre + ".resnets." + std::to_string(j), j == 0 ? ch : boc[i], boc[i]);
            if (r < 0) return fail(pre);
            enc_res_.push_back(r);
        }
        ch = boc[i];
        if (i < 3) {
            Conv3 c;
            if (!L.conv3(pre + ".downsamplers.0.conv", ch, ch, c, false)) return fail(pre + ".downsamplers");
            enc_down_.push_back(c);
        }
    }
    for (int j = 0; j < 2; ++j) {
        const int r = load_res("encoder.mid_block.resnets." + std::to_string(j), boc[3], boc[3]);
        if (r < 0) return fail("encoder.mid_block");
        enc_mid_[j] = r;
    }
    if (!load_attn("encoder.mid_block.attentions.0", boc[3], enc_attn_)) return fail("encoder.mid_block.attentions.0");
    if (!L.norm("encoder.conv_norm_out", boc[3], 1e-6f, enc_norm_)) return fail("encoder.conv_norm_out");
    if (!L.conv3("encoder.conv_out", 2 * vcfg.latent, boc[3], enc_out_, false)) return fail("encoder.conv_out");
    quant_w_ = L.vec("quant_conv.weight", (size_t)4 * vcfg.latent * vcfg.latent);
    quant_b_ = L.vec("quant_conv.bias", (size_t)2 * vcfg.latent);
    pq_w_ = L.vec("post_quant_conv.weight", (size_t)vcfg.latent * vcfg.latent);
    pq_b_ = L.vec("post_quant_conv.bias", vcfg.latent);
    if (!quant_w_ || !quant_b_ || !pq_w_ || !pq_b_) return fail("quant_conv / post_quant_conv");

    // ---- decoder
    if (!conv_in("decoder.conv_in", boc[3], vcfg.latent, dec_in_)) return fail("decoder.conv_in");
    for (int j = 0; j < 2; ++j) {
        const int r = load_res("decoder.mid_block.resnets." + std::to_string(j), boc[3], boc[3]);
        if (r < 0) return fail("decoder.mid_block");
        dec_mid_[j] = r;
    }
    if (!load_attn("decoder.mid_block.attentions.0", boc[3], dec_attn_)) return fail("decoder.mid_block.attentions.0");
    ch = boc[3];
    for (int i = 0; i < 4; ++i) {
        const std::string pre = "decoder.up_blocks." + std::to_string(i);
        const int co = boc[3 - i];
        for (int j = 0; j < vcfg.layers + 1; ++j) {
            const int r = load_res(pre + ".resnets." + std::to_string(j), j == 0 ? ch : co, co);
            if (r < 0) return fail(pre);
            dec_res_.push_back(r);
        }
        ch = co;
        if (i < 3) {
            Conv3 c;
            if (!L.conv3(pre + ".upsamplers.0.conv", ch, ch, c, false)) return fail(pre + ".upsamplers");
            dec_up_.push_back(c);
        }
    }
    if (!L.norm("decoder.conv_norm_out", boc[0], 1e-6f, dec_norm_)) return fail("decoder.conv_norm_out");
    if (!L.conv3("decoder.conv_out", vcfg.out_ch, boc[0], dec_out_, false)) return fail("decoder.conv_out");
    if (L.staging) cudaFree(L.staging);
    if (cudaDeviceSynchronize() != cudaSuccess) return set_error(S2I_ERR_CUDA, "vae load: %s", cudaGetErrorString(cudaGetLastError()));
    rsave_.resize(res_.size());
    vae_loaded_ = true;
    return 0;
}

// ================================================================================================== blocks
#define VRUN(call)                                                                                                  \
    do {                                                                                                            \
        if (arena_.overflow) return set_error(S2I_ERR_STATE, "vae: activation arena overflow (%zu of %zu bytes)", arena_.off, arena_.cap); \
        if (!dry_) S2I_TRY(call);                                                                                   \
    } while (0)

// AttentionBlock (diffusers <= 0.14, one head): x + proj_attn(softmax(q k^T / sqrt(C)) v), q / k / v = Linear(GroupNorm(x))
int VAE::vattn(const VAttn& A, const F32& x, F32& out) {
    const int B = x.B, H = x.H, W = x.W, C = A.C;
    double* s = new_stats();
    H16 n16 = new16(B, H, W, C);
    VRUN(gn_forward(x.p, x.ld, B, H * W, C, s, A.gn.g, A.gn.b, A.gn.eps, 0, n16.p, n16.ld, nullptr, 0, st_));
    H16 qkv = new16(B, H, W, 3 * C);
    S2I_TRY(gemm(n16, false, 1, A.qkv.w, C, 3 * C, C, A.qkv.b, nullptr, nullptr, nullptr, &qkv));
    Transformer T;                      // shape carrier for UNet::attention: one head as wide as the block
    T.C = C; T.heads = 1; T.d = C; T.dp = C; T.HP = C;
    H16 P, o;
    float* lse = nullptr;
    S2I_TRY(attention(T, qkv, 0, qkv, C, 2L * C, H * W, P, o, false, &lse));
    out = new32(B, H, W, C);
    S2I_TRY(gemm(o, false, 1, A.proj.w, C, C, C, A.proj.b, nullptr, &x, &out, nullptr));
    return 0;
}

int VAE::begin_pass(int B) {
    arena_.reset();
    dest_set_ = false;
    B_ = B;
    save_ = false;
    stats_off_ = 0;
    stats_cap_ = (size_t)64 * B * kGroups * 2;
    stats_ = dalloc<double>(stats_cap_);
    if (!dry_) S2I_MEMOP(cudaMemsetAsync(stats_, 0, stats_cap_ * sizeof(double), st_));
    return 0;
}

int VAE::run_encode(const float* x_nchw, int B, int H, int W, float* moments) {
    S2I_TRY(begin_pass(B));
    const int* boc = vcfg.boc;
    F32 x = new32(B, H, W, 4);
    VRUN(nchw_to_nhwc(x_nchw, B, vcfg.in_ch, H, W, x.p, x.ld, st_));
    H16 col = new16(B, H, W, 64);
    VRUN(im2col3x3(x.p, x.ld, B, H, W, vcfg.in_ch, 1, col.p, col.ld, st_));
    F32 h = new32(B, H, W, boc[0]);
    S2I_TRY(gemm(col, false, 1, enc_in_.w, 64, boc[0], 64, enc_in_.b, nullptr, nullptr, &h, nullptr));
    size_t ri = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < vcfg.layers; ++j) {
            F32 o;
            S2I_TRY(resblock(enc_res_[ri++], h, o));
            h = o;
        }
        if (i < 3) {
            const int Ho = h.H / 2, Wo = h.W / 2, C = h.C;
            H16 c2 = new16(B, Ho, Wo, 9 * C);
            VRUN(im2col3x3(h.p, h.ld, B, h.H, h.W, C, 2, c2.p, c2.ld, st_, /*pad=*/0));
            F32 o = new32(B, Ho, Wo, C);
            S2I_TRY(gemm(c2, false, 1, enc_down_[i].w, 9L * C, C, 9 * C, enc_down_[i].b, nullptr, nullptr, &o, nullptr));
            h = o;
        }
    }
    {
        F32 o;
        S2I_TRY(resblock(enc_mid_[0], h, o));
        h = o;
        S2I_TRY(vattn(enc_attn_, h, o));
        h = o;
        S2I_TRY(resblock(enc_mid_[1], h, o));
        h = o;
    }
    double* so = new_stats();
    H16 a = new16(B, h.H, h.W, h.C);
    VRUN(gn_forward(h.p, h.ld, B, h.H * h.W, h.C, so, enc_norm_.g, enc_norm_.b, enc_norm_.eps, 1, a.p, a.ld, nullptr, 0, st_));
    const int Cm = 2 * vcfg.latent;
    F32 y = new32(B, h.H, h.W, Cm);
    S2I_TRY(gemm(a, true, 9, enc_out_.w, 9L * h.C, Cm, h.C, enc_out_.b, nullptr, nullptr, &y, nullptr));
    if (!dry_) {
        const long HW = (long)h.H * h.W;
        S2I_LAUNCH((pointwise_to_nchw_kernel), grid1((long)B * Cm * HW), 256, 0, st_, y.p, y.ld, (long)B, HW, Cm, Cm, quant_w_, quant_b_, moments);
        S2I_LAUNCH_CHECK();
    }
    return 0;
}

int VAE::run_decode(const float* z_nchw, int B, int h_, int w_, float* image) {
    S2I_TRY(begin_pass(B));
    const int* boc = vcfg.boc;
    F32 z = new32(B, h_, w_, 4);
    if (!dry_) {
        const long HW = (long)h_ * w_;
        S2I_LAUNCH((pointwise_from_nchw_kernel), grid1((long)B * HW * 4), 256, 0, st_, z_nchw, (long)B, HW, vcfg.latent, vcfg.latent, pq_w_, pq_b_,
                   z.p, z.ld);
        S2I_LAUNCH_CHECK();
    }
    H16 col = new16(B, h_, w_, 64);
    VRUN(im2col3x3(z.p, z.ld, B, h_, w_, vcfg.latent, 1, col.p, col.ld, st_));
    F32 h = new32(B, h_, w_, boc[3]);
    S2I_TRY(gemm(col, false, 1, dec_in_.w, 64, boc[3], 64, dec_in_.b, nullptr, nullptr, &h, nullptr));
    {
        F32 o;
        S2I_TRY(resblock(dec_mid_[0], h, o));
        h = o;
        S2I_TRY(vattn(dec_attn_, h, o));
        h = o;
        S2I_TRY(resblock(dec_mid_[1], h, o));
        h = o;
    }
    size_t ri = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < vcfg.layers + 1; ++j) {
            F32 o;
            S2I_TRY(resblock(dec_res_[ri++], h, o));
            h = o;
        }
        if (i < 3) {
            H16 u = new16(B, 2 * h.H, 2 * h.W, h.C);
            VRUN(upsample2x(h.p, h.ld, B, h.H, h.W, h.C, u.p, u.ld, st_));
            F32 o = new32(B, 2 * h.H, 2 * h.W, h.C);
            S2I_TRY(gemm(u, true, 9, dec_up_[i].w, 9L * h.C, h.C, h.C, dec_up_[i].b, nullptr, nullptr, &o, nullptr));
            h = o;
        }
    }
    double* so = new_stats();
    H16 a = new16(B, h.H, h.W, h.C);
    VRUN(gn_forward(h.p, h.ld, B, h.H * h.W, h.C, so, dec_norm_.g, dec_norm_.b, dec_norm_.eps, 1, a.p, a.ld, nullptr, 0, st_));
    F32 img = new32(B, h.H, h.W, 4);
    img.C = vcfg.out_ch;
    S2I_TRY(gemm(a, true, 9, dec_out_.w, 9L * h.C, vcfg.out_ch, h.C, dec_out_.b, nullptr, nullptr, &img, nullptr));
    VRUN(nhwc_to_nchw(img.p, img.ld, B, vcfg.out_ch, h.H, h.W, image, st_));
    return 0;
}

// Arena sizing like UNet::forward: a dry run of the pass measures its footprint for this (pass, B, H, W).
template <class Body>
int VAE::sized(long key, Body body) {
    if (key != arena_key_) {
        Arena real = arena_;
        arena_ = Arena();
        dry_ = true;
        const int rc = body();
        dry_ = false;
        const size_t need = arena_.peak + (64u << 20);
        arena_ = real;
        if (rc != 0) return rc;
        if (need > arena_.cap) {
            if (arena_.base) cudaFree(arena_.base);
            arena_.base = nullptr;
            arena_.cap = 0;
            void* p = nullptr;
            if (cudaMalloc(&p, need) != cudaSuccess) {
                cudaGetLastError();
                return set_error(S2I_ERR_OOM, "vae: cannot allocate %.1f GB activation arena", need / 1e9);
            }
            ++g_alloc_gen;
            arena_.base = static_cast<char*>(p);
            arena_.cap = need;
        }
        arena_key_ = key;
    }
    int rc = body();
    if (rc == 0 && arena_.overflow) rc = set_error(S2I_ERR_STATE, "vae: activation arena overflow (%zu of %zu bytes)", arena_.peak, arena_.cap);
    return rc;
}

int VAE::encode(const float* x_nchw, int B, int H, int W, float* moments, cudaStream_t st) {
    if (!vae_loaded_) return set_error(S2I_ERR_STATE, "vae: weights not loaded");
    if (B < 1 || H % 8 || W % 8 || H < 8 || W < 8) return set_error(S2I_ERR_ARG, "vae encode: image H, W must be multiples of 8 (got %d x %d)", H, W);
    st_ = st;
    const long key = (1L << 62) ^ ((long)B << 40) ^ ((long)H << 20) ^ (long)W;
    return sized(key, [&] { return run_encode(x_nchw, B, H, W, moments); });
}

int VAE::decode(const float* z_nchw, int B, int h, int w, float* image, cudaStream_t st) {
    if (!vae_loaded_) return set_error(S2I_ERR_STATE, "vae: weights not loaded");
    if (B < 1 || h < 1 || w < 1) return set_error(S2I_ERR_ARG, "vae decode: bad latent shape");
    st_ = st;
    const long key = (1L << 61) ^ ((long)B << 40) ^ ((long)h << 20) ^ (long)w;
    return sized(key, [&] { return run_decode(z_nchw, B, h, w, image); });
}

}  // namespace s2i

// ================================================================================================== C ABI
struct s2i_vae {
    s2i::VAE* impl;
};

extern "C" {

int s2i_vae_create(const s2i_vae_config* c, s2i_vae** out) {
    if (!c || !out) return s2i::set_error(S2I_ERR_ARG, "s2i_vae_create: null argument");
    s2i::VaeConfig cfg;
    cfg.in_ch = c->in_channels;
    cfg.out_ch = c->out_channels;
    cfg.latent = c->latent_channels;
    cfg.layers = c->layers_per_block;
    for (int i = 0; i < 4; ++i) cfg.boc[i] = c->block_out_channels[i];
    if (cfg.in_ch < 1 || cfg.out_ch < 1 || cfg.latent < 1 || cfg.layers < 1) return s2i::set_error(S2I_ERR_ARG, "s2i_vae_create: bad configuration");
    *out = new s2i_vae{new s2i::VAE(cfg)};
    return 0;
}

void s2i_vae_destroy(s2i_vae* v) {
    if (!v) return;
    delete v->impl;
    delete v;
}

int s2i_vae_load(s2i_vae* v, int n, const char* const* names, const float* const* host_ptrs, const int* ndims, const long long* shapes) {
    if (!v) return s2i::set_error(S2I_ERR_ARG, "s2i_vae_load: null engine");
    std::map<std::string, s2i::HostParam> params;
    for (int i = 0; i < n; ++i) {
        s2i::HostParam hp;
        hp.data = host_ptrs[i];
        for (int k = 0; k < ndims[i]; ++k) hp.shape.push_back((long)shapes[i * 4 + k]);
        params[names[i]] = hp;
    }
    return v->impl->load_vae(params);
}

int s2i_vae_encode(s2i_vae* v, const float* image, int B, int H, int W, float* moments, void* cuda_stream) {
    if (!v || !image || !moments) return s2i::set_error(S2I_ERR_ARG, "s2i_vae_encode: null argument");
    return v->impl->encode(image, B, H, W, moments, static_cast<cudaStream_t>(cuda_stream));
}

int s2i_vae_decode(s2i_vae* v, const float* latents, int B, int h, int w, float* image, void* cuda_stream) {
    if (!v || !latents || !image) return s2i::set_error(S2I_ERR_ARG, "s2i_vae_decode: null argument");
    return v->impl->decode(latents, B, h, w, image, static_cast<cudaStream_t>(cuda_stream));
}

long long s2i_vae_arena_bytes(s2i_vae* v) { return v ? (long long)v->impl->arena_bytes() : 0; }

}  // extern "C"
