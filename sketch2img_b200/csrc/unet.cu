// UNet2DCondition forward (+9 taps) and tap->input backward.  See unet.cuh.
#include "unet.cuh"

#include <cmath>
#include <cstring>

#include "attn.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "loader.cuh"

namespace s2i {


// Standalone helper (LGP): stage a host fp32 [N][K] matrix and pack fp16 forward [N][K] / transposed [K][N] copies.
int pack_linear_host(const float* host, int N, int K, __half* w, long w_ld, __half* wd, long wd_ld) {
    float* stg = nullptr;
    S2I_CUDA(cudaMalloc(&stg, (size_t)N * K * sizeof(float)));
    cudaMemcpy(stg, host, (size_t)N * K * sizeof(float), cudaMemcpyHostToDevice);
    pack2d_kernel<<<1024, 256>>>(stg, K, 0, N, K, 0, 0, 0, 0, w, w_ld);
    if (wd) pack2d_kernel<<<1024, 256>>>(stg, K, 1, K, N, 0, 0, 0, 0, wd, wd_ld);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(stg);
    if (e != cudaSuccess) return set_error(S2I_ERR_CUDA, "pack_linear_host: %s", cudaGetErrorString(e));
    return 0;
}

// Same from a device fp32 matrix, enqueued on `st` (LGP training: the fp16 operand copies follow the updated masters).
int pack_linear_device(const float* dev, int N, int K, __half* w, long w_ld, __half* wd, long wd_ld, cudaStream_t st) {
    pack2d_kernel<<<1024, 256, 0, st>>>(dev, K, 0, N, K, 0, 0, 0, 0, w, w_ld);
    if (wd) pack2d_kernel<<<1024, 256, 0, st>>>(dev, K, 1, K, N, 0, 0, 0, 0, wd, wd_ld);
    g_prev_kernel = false;
    S2I_CUDA(cudaGetLastError());
    return 0;
}

// ================================================================================================== loading
UNet::~UNet() {
    for (auto& T : tfm_) {
        if (T.kv2_cache) cudaFree(T.kv2_cache);
        if (T.sat.kv16) cudaFree(T.sat.kv16);
    }
    for (void* p : owned_) cudaFree(p);
    if (arena_.base) cudaFree(arena_.base);
}

int UNet::load(const std::map<std::string, HostParam>& params) {
    Loader L{params, owned_};
    const int* boc = cfg.boc;
    const int temb_dim = boc[0] * 4;
    int temb_total = 0;
    struct TembSrc { std::string name; int C; };
    std::vector<TembSrc> temb_srcs;

    auto fail = [&](const char* what) {
        if (L.staging) cudaFree(L.staging);
        return set_error(S2I_ERR_ARG, "unet load (%s): %s", what, L.err.c_str());
    };

    auto load_res = [&](const std::string& pre, int Cin, int Cout) -> bool {
        ResBlock r;
        r.Cin = Cin;
        r.Cout = Cout;
        if (!L.norm(pre + ".norm1", Cin, 1e-5f, r.n1)) return false;
        if (!L.conv3(pre + ".conv1", Cout, Cin, r.c1)) return false;
        if (!L.norm(pre + ".norm2", Cout, 1e-5f, r.n2)) return false;
        if (!L.conv3(pre + ".conv2", Cout, Cout, r.c2)) return false;
        r.has_sc = Cin != Cout;
        if (r.has_sc && !L.linear(pre + ".conv_shortcut", Cout, Cin, true, r.sc)) return false;
        r.temb_off = temb_total;
        temb_total += Cout;
        temb_srcs.push_back({pre + ".time_emb_proj", Cout});
        res_.push_back(r);
        return true;
    };
    auto load_tfm = [&](const std::string& pre, int C, int heads) -> bool {
        Transformer T;
        T.path = pre;
        T.C = C;
        T.heads = heads;
        T.d = C / heads;
        T.dp = (int)rup(T.d, 16);
        T.HP = heads * T.dp;
        const int D = cfg.cross_dim;
        const std::string tb = pre + ".transformer_blocks.0";
        if (!L.norm(pre + ".norm", C, 1e-6f, T.gn)) return false;
        if (!L.linear(pre + ".proj_in", C, C, true, T.proj_in)) return false;
        if (!L.linear(pre + ".proj_out", C, C, true, T.proj_out)) return false;
        if (!L.norm(tb + ".norm1", C, 1e-5f, T.ln1) || !L.norm(tb + ".norm2", C, 1e-5f, T.ln2) ||
            !L.norm(tb + ".norm3", C, 1e-5f, T.ln3))
            return false;
        // fused, head-padded self-attention projection
        T.qkv.N = 3 * T.HP;
        T.qkv.K = C;
        T.qkv.w = L.dmalloc<__half>((size_t)3 * T.HP * C, true);
        T.qkv.wd = L.dmalloc<__half>((size_t)3 * T.HP * C, true);
        if (!T.qkv.w || !T.qkv.wd) return false;
        const char* names[3] = {".attn1.to_q", ".attn1.to_k", ".attn1.to_v"};
        for (int s = 0; s < 3; ++s)
            if (!L.linear_into(tb + names[s] + ".weight", C, C, T.d, T.dp, true, false, T.qkv.w, C, (long)s * T.HP,
                               T.qkv.wd, 3L * T.HP, (long)s * T.HP, true))
                return false;
        auto out_proj = [&](const std::string& p, Lin& l) -> bool {
            l.N = C;
            l.K = T.HP;
            l.w = L.dmalloc<__half>((size_t)C * T.HP, true);
            l.wd = L.dmalloc<__half>((size_t)C * T.HP, true);
            if (!l.w || !l.wd) return false;
            if (!L.linear_into(p + ".weight", C, C, T.d, T.dp, false, true, l.w, T.HP, 0, l.wd, C, 0, true)) return false;
            l.b = L.vec(p + ".bias", C);
            return l.b != nullptr;
        };
        if (!out_proj(tb + ".attn1.to_out.0", T.o1)) return false;
        T.q2.N = T.HP;
        T.q2.K = C;
        T.q2.w = L.dmalloc<__half>((size_t)T.HP * C, true);
        T.q2.wd = L.dmalloc<__half>((size_t)T.HP * C, true);
        if (!T.q2.w || !T.q2.wd) return false;
        if (!L.linear_into(tb + ".attn2.to_q.weight", C, C, T.d, T.dp, true, false, T.q2.w, C, 0, T.q2.wd, T.HP, 0, true))
            return false;
        T.kv2.N = 2 * T.HP;
        T.kv2.K = D;
        T.kv2.w = L.dmalloc<__half>((size_t)2 * T.HP * D, true);
        if (!T.kv2.w) return false;
        if (!L.linear_into(tb + ".attn2.to_k.weight", C, D, T.d, T.dp, true, false, T.kv2.w, D, 0, nullptr, 0, 0, false))
            return false;
        if (!L.linear_into(tb + ".attn2.to_v.weight", C, D, T.d, T.dp, true, false, T.kv2.w, D, T.HP, nullptr, 0, 0, false))
            return false;
        if (!out_proj(tb + ".attn2.to_out.0", T.o2)) return false;
        if (!L.linear_glu(tb + ".ff.net.0.proj", C, T.ff1)) return false;
        if (!L.linear(tb + ".ff.net.2", C, 4 * C, true, T.ff2)) return false;
        tfm_.push_back(T);
        return true;
    };

    // conv_in: forward as an im2col GEMM (K = 9*in_ch padded to 64), backward as a 3x3 dgrad conv
    {
        const int Ci = cfg.in_ch, Co = boc[0];
        const HostParam* hp = L.find("conv_in.weight", (size_t)Co * Ci * 9);
        if (!hp) return fail("conv_in");
        const float* s = L.stage(hp, (size_t)Co * Ci * 9);
        conv_in_.N = Co;
        conv_in_.K = 64;
        conv_in_.w = L.dmalloc<__half>((size_t)Co * 64, true);
        if (!s || !conv_in_.w) return fail("conv_in");
        pack_conv_kernel<<<64, 256>>>(s, Co, Ci, 0, conv_in_.w, 64);
        conv_in_d_.Cin = Ci;
        conv_in_d_.Cout = Co;
        conv_in_d_.wd = L.dmalloc<__half>((size_t)Co * Ci * 9);
        if (!conv_in_d_.wd) return fail("conv_in");
        pack_conv_kernel<<<64, 256>>>(s, Co, Ci, 1, conv_in_d_.wd, 9L * Co);
        cudaDeviceSynchronize();
        conv_in_.b = L.vec("conv_in.bias", Co);
        if (!conv_in_.b) return fail("conv_in");
    }
    if (!L.linear("time_embedding.linear_1", temb_dim, boc[0], true, time1_, false)) return fail("time_embedding");
    if (!L.linear("time_embedding.linear_2", temb_dim, temb_dim, true, time2_, false)) return fail("time_embedding");

    // down path
    int ch = boc[0];
    for (int i = 0; i < 4; ++i) {
        const std::string pre = "down_blocks." + std::to_string(i);
        for (int j = 0; j < cfg.layers; ++j) {
            if (!load_res(pre + ".resnets." + std::to_string(j), j == 0 ? ch : boc[i], boc[i])) return fail("down resnet");
            if (i < 3 && !cfg.encoder_only && !load_tfm(pre + ".attentions." + std::to_string(j), boc[i], cfg.heads[i])) return fail("down attn");
        }
        ch = boc[i];
        if (i < 3) {
            Conv3 c;
            if (!L.conv3(pre + ".downsamplers.0.conv", ch, ch, c)) return fail("downsample");
            down_.push_back(c);
        }
    }
    // mid
    if (!cfg.encoder_only) {
    if (!load_res("mid_block.resnets.0", boc[3], boc[3])) return fail("mid");
    if (!load_tfm("mid_block.attentions.0", boc[3], cfg.heads[3])) return fail("mid");
    if (!load_res("mid_block.resnets.1", boc[3], boc[3])) return fail("mid");
    }
    // up path
    if (!cfg.encoder_only) {
        int rev[4] = {boc[3], boc[2], boc[1], boc[0]};
        int out_c = rev[0];
        for (int i = 0; i < 4; ++i) {
            const std::string pre = "up_blocks." + std::to_string(i);
            const int prev = out_c;
            out_c = rev[i];
            const int in_c = rev[i + 1 < 4 ? i + 1 : 3];
            for (int j = 0; j < cfg.layers + 1; ++j) {
                const int skip = (j == cfg.layers) ? in_c : out_c;
                const int first = (j == 0) ? prev : out_c;
                if (!load_res(pre + ".resnets." + std::to_string(j), first + skip, out_c)) return fail("up resnet");
                if (i > 0 && !load_tfm(pre + ".attentions." + std::to_string(j), out_c, cfg.heads[3 - i]))
                    return fail("up attn");
            }
            if (i < 3) {
                Conv3 c;
                if (!L.conv3(pre + ".upsamplers.0.conv", out_c, out_c, c)) return fail("upsample");
                up_.push_back(c);
            }
        }
    }
    if (!cfg.encoder_only) {
        if (!L.norm("conv_norm_out", boc[0], 1e-5f, norm_out_)) return fail("conv_norm_out");
        if (!L.conv3("conv_out", cfg.out_ch, boc[0], conv_out_, false, true)) return fail("conv_out");
    }

    // fused time_emb_proj of every resnet: [sum Cout][temb_dim]
    temb_all_.N = temb_total;
    temb_all_.K = temb_dim;
    temb_all_.w = L.dmalloc<__half>((size_t)temb_total * temb_dim);
    temb_all_.b = L.dmalloc<float>(temb_total);
    if (!temb_all_.w || !temb_all_.b) return fail("temb");
    {
        long row = 0;
        for (auto& ts : temb_srcs) {
            if (!L.linear_into(ts.name + ".weight", ts.C, temb_dim, 0, 0, false, false, temb_all_.w, temb_dim, row, nullptr,
                               0, 0, false))
                return fail("time_emb_proj");
            const HostParam* hb = L.find(ts.name + ".bias", ts.C);
            if (!hb) return fail("time_emb_proj");
            cudaMemcpy(temb_all_.b + row, hb->data, ts.C * sizeof(float), cudaMemcpyHostToDevice);
            row += ts.C;
        }
    }
    if (L.staging) cudaFree(L.staging);
    if (cudaDeviceSynchronize() != cudaSuccess) return set_error(S2I_ERR_CUDA, "unet load: %s", cudaGetErrorString(cudaGetLastError()));
    {
        const int temb_dim = boc[0] * 4;
        Loader L2{params, owned_};
        temb_ = L2.dmalloc<float>(temb_all_.N);
        te0_ = L2.dmalloc<float>(boc[0]);
        te1_ = L2.dmalloc<float>(temb_dim);
        te2_ = L2.dmalloc<float>(temb_dim);
        if (!temb_ || !te0_ || !te1_ || !te2_) return set_error(S2I_ERR_OOM, "unet load: time-embedding buffers");
    }
    rsave_.resize(res_.size());
    tsave_.resize(tfm_.size());
    loaded_ = true;
    return 0;
}

// ================================================================================================== injected sketch attention
int UNet::load_sat(const std::map<std::string, HostParam>& params) {
    if (!loaded_) return set_error(S2I_ERR_STATE, "unet: load the UNet weights before the sketch-attention weights");
    Loader L{params, owned_};
    auto fail = [&](const std::string& what) {
        if (L.staging) cudaFree(L.staging);
        return set_error(S2I_ERR_ARG, "sketch attention load (%s): %s", what.c_str(), L.err.c_str());
    };
    // Re-uploads (SatMixin.sync after load_state_dict / an optimizer step) reuse the device buffers of the first upload: no
    // leak, and step graphs captured earlier keep reading valid addresses -- they simply see the new weights.
    auto vec_into = [&](const std::string& nm, size_t n, float*& dst) -> bool {
        const HostParam* hp = L.find(nm, n);
        if (!hp) return false;
        if (!dst) dst = L.dmalloc<float>(n);
        if (!dst) return false;
        return cudaMemcpy(dst, hp->data, n * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    for (auto& T : tfm_) {
        std::string name = "sketch_attn_" + T.path + ".transformer_blocks.0";
        for (char& ch : name)
            if (ch == '.') ch = '_';
        const int C = T.C;
        SatBlock& S = T.sat;
        S.ln.C = C;
        S.ln.eps = 1e-5f;
        if (!vec_into(name + ".sketch_norm.weight", C, S.ln.g) || !vec_into(name + ".sketch_norm.bias", C, S.ln.b)) return fail(name);
        S.q.N = T.HP; S.q.K = C;
        if (!S.q.w) S.q.w = L.dmalloc<__half>((size_t)T.HP * C, true);
        S.kv.N = 2 * T.HP; S.kv.K = C;
        if (!S.kv.w) S.kv.w = L.dmalloc<__half>((size_t)2 * T.HP * C, true);
        S.o.N = C; S.o.K = T.HP;
        if (!S.o.w) S.o.w = L.dmalloc<__half>((size_t)C * T.HP, true);
        if (!S.q.w || !S.kv.w || !S.o.w) return fail(name);
        if (!L.linear_into(name + ".sketch_attn.to_q.weight", C, C, T.d, T.dp, true, false, S.q.w, C, 0, nullptr, 0, 0, false) ||
            !L.linear_into(name + ".sketch_attn.to_k.weight", C, C, T.d, T.dp, true, false, S.kv.w, C, 0, nullptr, 0, 0, false) ||
            !L.linear_into(name + ".sketch_attn.to_v.weight", C, C, T.d, T.dp, true, false, S.kv.w, C, T.HP, nullptr, 0, 0, false) ||
            !L.linear_into(name + ".sketch_attn.to_out.0.weight", C, C, T.d, T.dp, false, true, S.o.w, T.HP, 0, nullptr, 0, 0, false))
            return fail(name);
        if (!vec_into(name + ".sketch_attn.to_out.0.bias", C, S.o.b) ||
            !vec_into(name + ".sketch_conv.weight", (size_t)C * C, S.conv_w32) ||     // [C][C][1]
            !vec_into(name + ".sketch_conv.bias", C, S.conv_b32))
            return fail(name);
        S.conv.N = C; S.conv.K = C;
        if (!S.conv.w) S.conv.w = L.dmalloc<__half>((size_t)C * C);
        if (!S.conv.b) S.conv.b = L.dmalloc<float>(C);
        if (!S.conv.w || !S.conv.b) return fail(name);
        S.loaded = true;
    }
    if (L.staging) cudaFree(L.staging);
    return set_sat_scale(sat_scale_, nullptr);
}

int UNet::set_sat_scale(float scale, cudaStream_t st) {
    sat_scale_ = scale;
    for (auto& T : tfm_) {
        SatBlock& S = T.sat;
        if (!S.loaded) continue;
        scale_pack_kernel<<<256, 256, 0, st>>>(S.conv_w32, (long)T.C * T.C, scale, S.conv.w, S.conv_b32, T.C, S.conv.b);
    }
    S2I_CUDA(cudaGetLastError());
    g_prev_kernel = false;
    return 0;
}

int UNet::set_sat_feature(const char* block_path, const float* nchw, int B, int C, int H, int W, cudaStream_t st) {
    Transformer* T = nullptr;
    for (auto& t : tfm_)
        if (t.path == block_path) T = &t;
    if (!T) return set_error(S2I_ERR_ARG, "sketch attention: no transformer block at '%s'", block_path);
    SatBlock& S = T->sat;
    if (!nchw) {
        S.fB = S.fN = 0;
        return 0;
    }
    if (!S.loaded) return set_error(S2I_ERR_STATE, "sketch attention: weights not loaded");
    if (C != T->C) return set_error(S2I_ERR_ARG, "sketch attention: block %s has %d channels, feature has %d", block_path, T->C, C);
    const int N = H * W;
    const size_t need = (size_t)B * N * 2 * T->HP * sizeof(__half);
    if (need > S.kv_cap) {
        if (S.kv16) cudaFree(S.kv16);
        S.kv16 = nullptr;
        S.kv_cap = 0;
        void* q = nullptr;
        if (cudaMalloc(&q, need) != cudaSuccess) {
            cudaGetLastError();
            return set_error(S2I_ERR_OOM, "sketch attention: cannot allocate the K/V cache");
        }
        ++g_alloc_gen;
        S.kv16 = static_cast<__half*>(q);
        S.kv_cap = need;
    }
    // tokens "b c h w -> b (h w) c" (sketch_guided_attn.py:82) in fp16, then K | V = tokens [to_k | to_v]^T once per feature
    __half* tok = nullptr;
    S2I_CUDA(cudaMalloc(reinterpret_cast<void**>(&tok), (size_t)B * N * C * sizeof(__half)));
    nchw_to_tokens16_kernel<<<1024, 256, 0, st>>>(nchw, B, C, N, tok);
    g_prev_kernel = false;
    GemmDesc g;
    g.tag = "gemm_linear";
    g.A = tok; g.aC = C; g.aW = B * N; g.a_sw = C;
    g.B = S.kv.w; g.bI = C; g.bR = 2 * T->HP; g.b_sr = C;
    g.N = 2 * T->HP; g.Kc = C;
    g.out16 = S.kv16; g.ld16 = 2 * T->HP;
    int rc = gemm_launch(g, st);
    cudaStreamSynchronize(st);
    cudaFree(tok);
    if (rc != 0) return rc;
    S.fB = B;
    S.fN = N;
    return 0;
}

// ================================================================================================== helpers
#define ARENA_CHECK()                                                                                                   \
    do {                                                                                                                \
        if (arena_.overflow)                                                                                            \
            return set_error(S2I_ERR_STATE, "unet: activation arena overflow (%zu of %zu bytes): the forward's topology " \
                             "differs from the one the arena was sized for", arena_.off, arena_.cap);                   \
    } while (0)
#define RUN(call)                  \
    do {                           \
        ARENA_CHECK();             \
        if (!dry_) S2I_TRY(call);  \
    } while (0)

unsigned UNet::sat_signature() const {
    unsigned m = 0;
    for (size_t i = 0; i < tfm_.size() && i < 32; ++i)
        if (tfm_[i].sat.loaded && tfm_[i].sat.fN > 0) m |= 1u << i;
    return m;
}

F32 UNet::new32(int B, int H, int W, int C) {
    F32 t;
    t.B = B; t.H = H; t.W = W; t.C = C;
    t.ld = C;
    t.p = static_cast<float*>(arena_.alloc((size_t)B * H * W * C * sizeof(float)));
    return t;
}
// The next layer output: a fresh arena tensor, or -- when the caller reserved a destination (the slice of an up-path
// concat buffer this output is one half of) -- that slice, so torch.cat([h, skip], dim=1) never copies.
F32 UNet::out32(int B, int H, int W, int C) {
    if (!dest_set_) return new32(B, H, W, C);
    dest_set_ = false;
    return dest_;       // shape checked by reserve()
}
void UNet::reserve(const F32& cat, int c0, int C) {
    dest_ = cat;
    dest_.p = cat.p + c0;
    dest_.C = C;
    dest_set_ = true;
}
H16 UNet::new16(int B, int H, int W, int C) {
    H16 t;
    t.B = B; t.H = H; t.W = W; t.C = C;
    t.ld = C;
    t.p = static_cast<__half*>(arena_.alloc((size_t)B * H * W * C * sizeof(__half)));
    return t;
}
double* UNet::new_stats() {
    const size_t n = (size_t)B_ * kGroups * 2;
    double* p = stats_ + stats_off_;
    stats_off_ += n;
    return p;
}

// GroupNorm forward: statistics from the producer's column sums when the input carries them (F32::st), else the two-phase kernel.
int UNet::group_norm(const F32& x, int C, double* slot, const float* gamma, const float* beta, float eps, int silu, H16& out,
                     H16* raw) {
    const int B = x.B, HW = x.H * x.W;
    // (every CTA of gn_norm adds up its channels' partials: beyond ~128 blocks per sample that costs more than a second read of x)
    bool fused = x.nst > 0 && gn_norm_supported(C);
    for (int k = 0; k < x.nst; ++k) fused = fused && x.st[k].bps <= 128;
    if (fused)
        RUN(gn_norm(x.p, x.ld, B, HW, C, x.st, x.nst, slot, gamma, beta, eps, silu, out.p, out.ld, raw ? raw->p : nullptr,
                    raw ? raw->ld : 0, st_));
    else
        RUN(gn_forward(x.p, x.ld, B, HW, C, slot, gamma, beta, eps, silu, out.p, out.ld, raw ? raw->p : nullptr, raw ? raw->ld : 0,
                       st_));
    return 0;
}

int UNet::gemm(const H16& a, bool spatial, int taps, const __half* w, long w_ld, int N, int Kc, const float* bias,
               const float* rowvec, const F32* residual, F32* out32, H16* out16, bool stats) {
    // An fp16-only output with a long contraction may be computed split-K through an fp32 scratch: give it its own arena
    // buffer (same allocation in the sizing pass) so its zero-fill joins the step's zero plan instead of a per-call launch.
    float* scratch = nullptr;
    if (gemm_split_add_mode() && out16 && !out32 && (long)Kc * taps >= 1024 && a.rows() * (long)N * 4 <= (32L << 20))
        scratch = dalloc<float>((size_t)a.rows() * N);
    // column statistics for the GroupNorm that reads this output: [B][cap][2][N] partial sums (same allocation in the sizing pass)
    float* cst = nullptr;
    int cst_cap = 0;
    // (only while a sample has at most 128 row blocks: beyond that group_norm() keeps the two-phase kernel, e.g. the VAE at 512 x 512)
    if (stats && out32 && spatial_stats_ && gn_norm_supported(N) && a.H * a.W <= 128 * 128) {
        const int HW = a.H * a.W;
        cst_cap = HW / 64 > 32 ? HW / 64 : 32;
        cst = dalloc<float>((size_t)a.B * cst_cap * 2 * N);
    }
    if (out32) out32->nst = 0;
    if (dry_) return 0;
    ARENA_CHECK();
    GemmDesc d;
    d.A = a.p;
    d.aC = Kc;
    if (spatial) {
        d.aW = a.W; d.aH = a.H; d.aB = a.B;
        d.a_sw = a.ld; d.a_sh = a.ld * a.W; d.a_sb = a.ld * a.W * a.H;
    } else {
        d.aW = (int)a.rows(); d.aH = 1; d.aB = 1;
        d.a_sw = a.ld;
    }
    d.taps = taps;
    d.tag = taps == 9 ? "gemm_conv3x3" : "gemm_linear";
    d.B = w;
    d.b_static = 1;         // packed weights: nothing in the step writes them
    d.bI = taps * Kc;
    d.bR = N;
    d.b_sr = w_ld;
    d.N = N;
    d.Kc = Kc;
    d.bias = bias;
    d.rowvec = rowvec;
    d.rowvec_ld = 0;
    if (residual) {
        d.residual = residual->p;
        d.res_ld = residual->ld;
    }
    if (out32) {
        d.out32 = out32->p;
        d.ld32 = out32->ld;
    }
    if (out16) {
        d.out16 = out16->p;
        d.ld16 = out16->ld;
    }
    d.scratch32 = scratch;
    int bps = 0;
    if (cst) {
        // the pixel geometry of the OUTPUT: a non-spatial A (im2col rows) still produces B * H * W result rows in sample order
        if (!spatial) {
            d.aW = a.W; d.aH = a.H; d.aB = a.B;
            d.a_sw = a.ld; d.a_sh = a.ld * a.W; d.a_sb = a.ld * a.W * a.H;
        }
        d.colstat = cst;
        d.colstat_ld = N;
        d.colstat_cap = cst_cap;
        d.colstat_bps = &bps;
    }
    S2I_TRY(gemm_launch(d, st_));
    if (cst && bps > 0) {
        out32->nst = 1;
        out32->st[0].p = cst;
        out32->st[0].cap = cst_cap;
        out32->st[0].bps = bps;
        out32->st[0].ld = N;
        out32->st[0].c0 = 0;
        out32->st[0].c1 = N;
    }
    return 0;
}

int UNet::accumulate(F32& acc, const F32& g) {
    if (!g.p) return 0;
    if (!acc.p) {
        acc = g;
        return 0;
    }
    F32 out = new32(acc.B, acc.H, acc.W, acc.C);
    H16 oh = new16(acc.B, acc.H, acc.W, acc.C);
    RUN(add2d(acc.p, acc.ld, g.p, g.ld, acc.rows(), acc.C, out.p, out.ld, oh.p, oh.ld, st_));
    out.h = oh.p;
    out.hld = oh.ld;
    acc = out;
    return 0;
}

H16 UNet::half_of(const F32& t) {
    H16 h;
    h.B = t.B; h.H = t.H; h.W = t.W; h.C = t.C;
    if (t.h) {
        h.p = t.h;
        h.ld = t.hld;
        return h;
    }
    h = new16(t.B, t.H, t.W, t.C);
    if (!dry_) {
        if (cast2d(t.p, t.ld, t.rows(), t.C, 1.f, h.p, h.ld, st_) != 0) h.p = nullptr;
    }
    return h;
}

// ================================================================================================== ResnetBlock2D
int UNet::resblock(int idx, const F32& x, F32& out) {
    const ResBlock& R = res_[idx];
    const int B = x.B, H = x.H, W = x.W, HW = H * W;
    double* s1 = new_stats();
    H16 a1 = new16(B, H, W, R.Cin);
    H16 x16;
    if (R.has_sc) x16 = new16(B, H, W, R.Cin);
    S2I_TRY(group_norm(x, R.Cin, s1, R.n1.g, R.n1.b, R.n1.eps, 1, a1, R.has_sc ? &x16 : nullptr));
    F32 h1 = new32(B, H, W, R.Cout);
    S2I_TRY(gemm(a1, true, 9, R.c1.w, 9L * R.Cin, R.Cout, R.Cin, R.c1.b, R.temb_off >= 0 ? temb_ + R.temb_off : nullptr, nullptr, &h1,
                 nullptr, true));
    double* s2 = new_stats();
    H16 a2 = new16(B, H, W, R.Cout);
    S2I_TRY(group_norm(h1, R.Cout, s2, R.n2.g, R.n2.b, R.n2.eps, 1, a2, nullptr));
    F32 res = x;
    if (R.has_sc) {
        F32 sc = new32(B, H, W, R.Cout);
        S2I_TRY(gemm(x16, false, 1, R.sc.w, R.Cin, R.Cout, R.Cin, R.sc.b, nullptr, nullptr, &sc, nullptr));
        res = sc;
    }
    out = out32(B, H, W, R.Cout);
    S2I_TRY(gemm(a2, true, 9, R.c2.w, 9L * R.Cout, R.Cout, R.Cout, R.c2.b, nullptr, &res, &out, nullptr, true));
    if (keep_debug) {
        debug["r" + std::to_string(idx) + ".h1"] = h1;
        debug["r" + std::to_string(idx) + ".out"] = out;
    }
    if (save_) {
        rsave_[idx].x = x;
        rsave_[idx].h1 = h1;
        rsave_[idx].s1 = s1;
        rsave_[idx].s2 = s2;
    }
    return 0;
}

// Views of what the forward saved, restricted to the samples [bb0_, bb0_ + bnb_) the backward runs on.
F32 UNet::sub(const F32& t) const {
    F32 v = t;
    if (t.p) v.p = t.p + (size_t)bb0_ * t.H * t.W * t.ld;
    v.B = bnb_;
    return v;
}
H16 UNet::sub(const H16& t) const {
    H16 v = t;
    if (t.p) v.p = t.p + (size_t)bb0_ * t.H * t.W * t.ld;
    v.B = bnb_;
    return v;
}
ResSave UNet::sub(const ResSave& s) const {
    ResSave v;
    v.x = sub(s.x);
    v.h1 = sub(s.h1);
    v.s1 = s.s1 + (size_t)bb0_ * kGroups * 2;
    v.s2 = s.s2 + (size_t)bb0_ * kGroups * 2;
    return v;
}
TfmSave UNet::sub(const TfmSave& s, const Transformer& T) const {
    TfmSave v;
    v.x = sub(s.x); v.t0 = sub(s.t0); v.t1 = sub(s.t1); v.t2 = sub(s.t2);
    v.ff = sub(s.ff); v.qkv = sub(s.qkv); v.q2 = sub(s.q2); v.kv2 = sub(s.kv2); v.o1 = sub(s.o1); v.o2 = sub(s.o2);
    const size_t tok = (size_t)bb0_ * s.x.H * s.x.W;
    v.gs = s.gs + (size_t)bb0_ * kGroups * 2;
    v.l1 = s.l1 + tok * 2; v.l2 = s.l2 + tok * 2; v.l3 = s.l3 + tok * 2;
    v.lse1 = s.lse1 ? s.lse1 + tok * T.heads : nullptr;
    v.lse2 = s.lse2 ? s.lse2 + tok * T.heads : nullptr;
    // unfused attention keeps P as [B * heads][Nq][ldP]
    auto subP = [&](const H16& P) {
        H16 w = P;
        if (P.p) w.p = P.p + (size_t)bb0_ * T.heads * P.H * P.W * P.ld;
        w.B = bnb_ * T.heads;
        return w;
    };
    v.P1 = subP(s.P1);
    v.P2 = subP(s.P2);
    return v;
}

int UNet::resblock_bwd(int idx, const F32& dout, F32& dx) {
    const ResBlock& R = res_[idx];
    const ResSave S = sub(rsave_[idx]);
    const int B = dout.B, H = dout.H, W = dout.W, HW = H * W;
    H16 d16 = half_of(dout);
    if (!d16.p) return S2I_ERR_CUDA;
    F32 da2 = new32(B, H, W, R.Cout);
    S2I_TRY(gemm(d16, true, 9, R.c2.wd, 9L * R.Cout, R.Cout, R.Cout, nullptr, nullptr, nullptr, &da2, nullptr));
    double* bs2 = new_stats();
    H16 dh1 = new16(B, H, W, R.Cout);
    RUN(gn_backward(da2.p, da2.ld, S.h1.p, S.h1.ld, B, HW, R.Cout, S.s2, bs2, R.n2.g, R.n2.b, R.n2.eps, 1, nullptr, 0,
                    nullptr, 0, dh1.p, dh1.ld, st_));
    F32 da1 = new32(B, H, W, R.Cin);
    S2I_TRY(gemm(dh1, true, 9, R.c1.wd, 9L * R.Cout, R.Cin, R.Cout, nullptr, nullptr, nullptr, &da1, nullptr));
    double* bs1 = new_stats();
    dx = new32(B, H, W, R.Cin);
    H16 dxh = new16(B, H, W, R.Cin);        // fp16 copy for the next layer's GEMMs, written by the same kernel
    if (R.has_sc) {
        F32 tmp = new32(B, H, W, R.Cin);
        RUN(gn_backward(da1.p, da1.ld, S.x.p, S.x.ld, B, HW, R.Cin, S.s1, bs1, R.n1.g, R.n1.b, R.n1.eps, 1, nullptr, 0,
                        tmp.p, tmp.ld, nullptr, 0, st_));
        S2I_TRY(gemm(d16, false, 1, R.sc.wd, R.Cout, R.Cin, R.Cout, nullptr, nullptr, &tmp, &dx, &dxh));
    } else {
        RUN(gn_backward(da1.p, da1.ld, S.x.p, S.x.ld, B, HW, R.Cin, S.s1, bs1, R.n1.g, R.n1.b, R.n1.eps, 1, dout.p,
                        dout.ld, dx.p, dx.ld, dxh.p, dxh.ld, st_));
    }
    dx.h = dxh.p;
    dx.hld = dxh.ld;
    return 0;
}

// ================================================================================================== attention
// q: [B*Nq][q.ld] with this attention's Q heads at column q_c0; kv: [B*Nk][kv.ld] with K heads at k_c0, V at v_c0.
int UNet::attention(const Transformer& T, const H16& q, long q_c0, const H16& kv, long k_c0, long v_c0, int Nk, H16& P,
                    H16& o, bool need_bwd, float** lse) {
    const int B = q.B, Nq = q.H * q.W, Z = B * T.heads;
    *lse = nullptr;
    const bool fused_bwd = need_bwd && attn_bwd_supported(Nq, Nk, T.dp);
    if ((!need_bwd || fused_bwd) && use_flash_ && attn_fwd_supported(Nq, Nk, T.dp)) {
        // fused path: the scores stay in TMEM / shared memory; a later backward recomputes them from the log-sum-exp
        P = H16();
        o = new16(q.B, q.H, q.W, T.HP);
        if (fused_bwd) *lse = dalloc<float>((size_t)Z * Nq);
        if (dry_) return 0;
        ARENA_CHECK();
        AttnDesc a;
        a.q = q.p; a.ldq = q.ld; a.q_c0 = (int)q_c0;
        a.kv = kv.p; a.ldkv = kv.ld; a.k_c0 = (int)k_c0; a.v_c0 = (int)v_c0;
        a.B = B; a.heads = T.heads; a.Nq = Nq; a.Nk = Nk; a.dp = T.dp; a.d_true = T.d;
        a.scale = 1.f / sqrtf((float)T.d);
        a.out = o.p; a.ldo = o.ld;
        a.lse = *lse;
        return attn_fwd_launch(a, st_);
    }
    const long ldS = rup(Nk, 4), ldP = rup(Nk, 8);
    float* S = dalloc<float>((size_t)Z * Nq * ldS);
    P = H16();
    P.B = Z; P.H = 1; P.W = Nq; P.C = Nk; P.ld = ldP;
    P.p = dalloc<__half>((size_t)Z * Nq * ldP);
    o = new16(q.B, q.H, q.W, T.HP);
    if (dry_) return 0;
    ARENA_CHECK();
    GemmDesc d;
    d.tag = "gemm_attn";
    d.A = q.p; d.aC = (int)q.ld; d.aW = Nq; d.aB = B; d.a_sw = q.ld; d.a_sb = (long)Nq * q.ld;
    d.a_c0 = (int)q_c0; d.a_hoff = T.dp;
    d.B = kv.p; d.bI = (int)kv.ld; d.bR = Nk; d.bZ = B; d.b_sr = kv.ld; d.b_sz = (long)Nk * kv.ld;
    d.b_c0 = (int)k_c0; d.b_hoff = T.dp;
    d.N = Nk; d.Kc = T.dp; d.Z = Z; d.zh = T.heads;
    d.alpha = 1.f / sqrtf((float)T.d);
    d.out32 = S; d.ld32 = ldS; d.c_sb = (long)T.heads * Nq * ldS; d.c_sh = (long)Nq * ldS;
    S2I_TRY(gemm_launch(d, st_));
    S2I_TRY(softmax_fwd(S, ldS, (long)Z * Nq, Nk, P.p, ldP, st_));
    GemmDesc e;
    e.tag = "gemm_attn";
    e.A = P.p; e.aC = Nk; e.aW = Nq; e.aB = Z; e.a_sw = ldP; e.a_sb = (long)Nq * ldP; e.a_zmode = 1;
    e.B = kv.p; e.b_mn = 1; e.bI = (int)kv.ld; e.bR = Nk; e.bZ = B; e.b_sr = kv.ld; e.b_sz = (long)Nk * kv.ld;
    e.b_c0 = (int)v_c0; e.b_hoff = T.dp;
    e.N = T.dp; e.BN = T.dp <= 256 ? T.dp : 128; e.Kc = Nk; e.Z = Z; e.zh = T.heads;      // heads wider than a tile (VAE: 512) take several
    e.out16 = o.p; e.ld16 = o.ld; e.c_sb = (long)Nq * o.ld; e.c_sh = T.dp;
    S2I_TRY(gemm_launch(e, st_));
    return 0;
}

int UNet::attention_bwd(const Transformer& T, const H16& dO, const H16& q, long q_c0, const H16& kv, long k_c0,
                        long v_c0, int Nk, const H16& P, const H16& o, const float* lse, H16& dq, long dq_c0, H16* dkv,
                        long dk_c0, long dv_c0) {
    const int B = q.B, Nq = q.H * q.W, Z = B * T.heads;
    if (lse) {
        // fused: dQ (and dK, dV) with S / P / dP / dS recomputed on chip
        float* delta = dalloc<float>((size_t)Z * Nq);
        if (dry_) return 0;
        ARENA_CHECK();
        AttnBwdDesc a;
        a.q = q.p; a.ldq = q.ld; a.q_c0 = (int)q_c0;
        a.kv = kv.p; a.ldkv = kv.ld; a.k_c0 = (int)k_c0; a.v_c0 = (int)v_c0;
        a.dO = dO.p; a.lddo = dO.ld;
        a.o = o.p; a.ldo = o.ld;
        a.lse = lse; a.delta = delta;
        a.B = B; a.heads = T.heads; a.Nq = Nq; a.Nk = Nk; a.dp = T.dp; a.d_true = T.d;
        a.scale = 1.f / sqrtf((float)T.d);
        a.dq = dq.p; a.lddq = dq.ld; a.dq_c0 = (int)dq_c0;
        if (dkv) {
            a.dk = a.dv = dkv->p; a.lddkv = dkv->ld; a.dk_c0 = (int)dk_c0; a.dv_c0 = (int)dv_c0;
        }
        return attn_bwd_launch(a, st_);
    }
    const long ldS = rup(Nk, 4), ldP = P.ld;
    float* dP = dalloc<float>((size_t)Z * Nq * ldS);
    __half* dS = dalloc<__half>((size_t)Z * Nq * ldP);
    if (dry_) return 0;
    ARENA_CHECK();
    const float scale = 1.f / sqrtf((float)T.d);
    {   // dP = dO V^T
        GemmDesc d;
            d.tag = "gemm_attn_bwd";
        d.A = dO.p; d.aC = (int)dO.ld; d.aW = Nq; d.aB = B; d.a_sw = dO.ld; d.a_sb = (long)Nq * dO.ld; d.a_hoff = T.dp;
        d.B = kv.p; d.bI = (int)kv.ld; d.bR = Nk; d.bZ = B; d.b_sr = kv.ld; d.b_sz = (long)Nk * kv.ld;
        d.b_c0 = (int)v_c0; d.b_hoff = T.dp;
        d.N = Nk; d.Kc = T.dp; d.Z = Z; d.zh = T.heads;
        d.out32 = dP; d.ld32 = ldS; d.c_sb = (long)T.heads * Nq * ldS; d.c_sh = (long)Nq * ldS;
        S2I_TRY(gemm_launch(d, st_));
    }
    S2I_TRY(softmax_bwd(P.p, ldP, dP, ldS, (long)Z * Nq, Nk, scale, dS, ldP, st_));
    {   // dQ = dS K
        GemmDesc d;
            d.tag = "gemm_attn_bwd";
        d.A = dS; d.aC = Nk; d.aW = Nq; d.aB = Z; d.a_sw = ldP; d.a_sb = (long)Nq * ldP; d.a_zmode = 1;
        d.B = kv.p; d.b_mn = 1; d.bI = (int)kv.ld; d.bR = Nk; d.bZ = B; d.b_sr = kv.ld; d.b_sz = (long)Nk * kv.ld;
        d.b_c0 = (int)k_c0; d.b_hoff = T.dp;
        d.N = T.dp; d.BN = T.dp; d.Kc = Nk; d.Z = Z; d.zh = T.heads;
        d.out16 = dq.p + dq_c0; d.ld16 = dq.ld; d.c_sb = (long)Nq * dq.ld; d.c_sh = T.dp;
        S2I_TRY(gemm_launch(d, st_));
    }
    if (dkv) {
        {   // dV = P^T dO
            GemmDesc d;
            d.tag = "gemm_attn_bwd";
            d.A = P.p; d.a_mn = 1; d.aC = Nk; d.aW = Nq; d.aB = Z; d.a_sw = ldP; d.a_sb = (long)Nq * ldP; d.a_zmode = 1;
            d.B = dO.p; d.b_mn = 1; d.bI = (int)dO.ld; d.bR = Nq; d.bZ = B; d.b_sr = dO.ld; d.b_sz = (long)Nq * dO.ld;
            d.b_hoff = T.dp;
            d.N = T.dp; d.BN = T.dp; d.Kc = Nq; d.Z = Z; d.zh = T.heads;
            d.out16 = dkv->p + dv_c0; d.ld16 = dkv->ld; d.c_sb = (long)Nk * dkv->ld; d.c_sh = T.dp;
            S2I_TRY(gemm_launch(d, st_));
        }
        {   // dK = dS^T Q
            GemmDesc d;
            d.tag = "gemm_attn_bwd";
            d.A = dS; d.a_mn = 1; d.aC = Nk; d.aW = Nq; d.aB = Z; d.a_sw = ldP; d.a_sb = (long)Nq * ldP; d.a_zmode = 1;
            d.B = q.p; d.b_mn = 1; d.bI = (int)q.ld; d.bR = Nq; d.bZ = B; d.b_sr = q.ld; d.b_sz = (long)Nq * q.ld;
            d.b_c0 = (int)q_c0; d.b_hoff = T.dp;
            d.N = T.dp; d.BN = T.dp; d.Kc = Nq; d.Z = Z; d.zh = T.heads;
            d.out16 = dkv->p + dk_c0; d.ld16 = dkv->ld; d.c_sb = (long)Nk * dkv->ld; d.c_sh = T.dp;
            S2I_TRY(gemm_launch(d, st_));
        }
    }
    return 0;
}

// ================================================================================================== Transformer2D
int UNet::transformer(int idx, const F32& x, F32& out) {
    const Transformer& T = tfm_[idx];
    const int B = x.B, H = x.H, W = x.W, HW = H * W, C = T.C;
    const long rows = x.rows();
    TfmSave sv;
    sv.x = x;
    sv.gs = new_stats();
    H16 n16 = new16(B, H, W, C);
    S2I_TRY(group_norm(x, C, sv.gs, T.gn.g, T.gn.b, T.gn.eps, 0, n16, nullptr));
    sv.t0 = new32(B, H, W, C);
    S2I_TRY(gemm(n16, false, 1, T.proj_in.w, C, C, C, T.proj_in.b, nullptr, nullptr, &sv.t0, nullptr));
    // --- self attention
    H16 l16 = new16(B, H, W, C);
    sv.l1 = dalloc<float>(rows * 2);
    RUN(ln_fwd(sv.t0.p, sv.t0.ld, rows, C, T.ln1.g, T.ln1.b, T.ln1.eps, l16.p, l16.ld, sv.l1, st_));
    sv.qkv = new16(B, H, W, 3 * T.HP);
    S2I_TRY(gemm(l16, false, 1, T.qkv.w, C, 3 * T.HP, C, nullptr, nullptr, nullptr, nullptr, &sv.qkv));
    // P is only kept for the input-gradient pass, and up_blocks[3] (the last layers+1 transformers) is not on it
    const bool need_bwd = save_ && idx < (int)tfm_.size() - (cfg.layers + 1);
    S2I_TRY(attention(T, sv.qkv, 0, sv.qkv, T.HP, 2L * T.HP, HW, sv.P1, sv.o1, need_bwd, &sv.lse1));
    sv.t1 = new32(B, H, W, C);
    S2I_TRY(gemm(sv.o1, false, 1, T.o1.w, T.HP, C, T.HP, T.o1.b, nullptr, &sv.t0, &sv.t1, nullptr));
    // --- injected sketch attention (sketch_guided_attn.py:120-132), when a feature is set for this block
    if (T.sat.loaded && T.sat.fN > 0) {
        const SatBlock& S = T.sat;
        if (save_) return set_error(S2I_ERR_STATE, "sketch attention has no backward (the reference runs it under no_grad)");
        if (S.fB != B || S.fN != HW)
            return set_error(S2I_ERR_ARG, "sketch attention: block %s expects a [%d, %d, %d x %d tokens] feature, has [%d, %d tokens]",
                             T.path.c_str(), B, C, H, W, S.fB, S.fN);
        H16 ls = new16(B, H, W, C);
        float* lst = dalloc<float>(rows * 2);
        RUN(ln_fwd(sv.t1.p, sv.t1.ld, rows, C, S.ln.g, S.ln.b, S.ln.eps, ls.p, ls.ld, lst, st_));
        H16 qs = new16(B, H, W, T.HP);
        S2I_TRY(gemm(ls, false, 1, S.q.w, C, T.HP, C, nullptr, nullptr, nullptr, nullptr, &qs));
        H16 kvs;
        kvs.p = S.kv16; kvs.B = B; kvs.H = 1; kvs.W = HW; kvs.C = 2 * T.HP; kvs.ld = 2 * T.HP;
        H16 Ps, os;
        float* lse_s = nullptr;
        S2I_TRY(attention(T, qs, 0, kvs, 0, T.HP, HW, Ps, os, false, &lse_s));
        H16 u = new16(B, H, W, C);
        S2I_TRY(gemm(os, false, 1, S.o.w, T.HP, C, T.HP, S.o.b, nullptr, nullptr, nullptr, &u));
        F32 t1s = new32(B, H, W, C);
        S2I_TRY(gemm(u, false, 1, S.conv.w, C, C, C, S.conv.b, nullptr, &sv.t1, &t1s, nullptr));
        sv.t1 = t1s;
    }
    // --- cross attention (K/V from the text context)
    H16 l16b = new16(B, H, W, C);
    sv.l2 = dalloc<float>(rows * 2);
    RUN(ln_fwd(sv.t1.p, sv.t1.ld, rows, C, T.ln2.g, T.ln2.b, T.ln2.eps, l16b.p, l16b.ld, sv.l2, st_));
    sv.q2 = new16(B, H, W, T.HP);
    S2I_TRY(gemm(l16b, false, 1, T.q2.w, C, T.HP, C, nullptr, nullptr, nullptr, nullptr, &sv.q2));
    {
        // persistent (not arena) so it survives to the next steps of the image
        Transformer& Tm = tfm_[idx];
        const size_t need = (size_t)B * cfg.ctx_len * 2 * T.HP * sizeof(__half);
        if (!dry_ && need > Tm.kv2_cap) {
            if (Tm.kv2_cache) cudaFree(Tm.kv2_cache);
            Tm.kv2_cache = nullptr;
            Tm.kv2_cap = 0;
            void* q = nullptr;
            if (cudaMalloc(&q, need) != cudaSuccess) {
                cudaGetLastError();
                return set_error(S2I_ERR_OOM, "unet: cannot allocate the context K/V cache");
            }
            ++g_alloc_gen;
            Tm.kv2_cache = static_cast<__half*>(q);
            Tm.kv2_cap = need;
        }
        sv.kv2 = H16();
        sv.kv2.B = B; sv.kv2.H = 1; sv.kv2.W = cfg.ctx_len; sv.kv2.C = 2 * T.HP; sv.kv2.ld = 2 * T.HP;
        sv.kv2.p = Tm.kv2_cache;
        if (!reuse_kv_)
            S2I_TRY(gemm(ctx16_, false, 1, T.kv2.w, cfg.cross_dim, 2 * T.HP, cfg.cross_dim, nullptr, nullptr, nullptr, nullptr,
                         &sv.kv2));
    }
    S2I_TRY(attention(T, sv.q2, 0, sv.kv2, 0, T.HP, cfg.ctx_len, sv.P2, sv.o2, need_bwd, &sv.lse2));
    sv.t2 = new32(B, H, W, C);
    S2I_TRY(gemm(sv.o2, false, 1, T.o2.w, T.HP, C, T.HP, T.o2.b, nullptr, &sv.t1, &sv.t2, nullptr));
    // --- GEGLU feed-forward
    H16 l16c = new16(B, H, W, C);
    sv.l3 = dalloc<float>(rows * 2);
    RUN(ln_fwd(sv.t2.p, sv.t2.ld, rows, C, T.ln3.g, T.ln3.b, T.ln3.eps, l16c.p, l16c.ld, sv.l3, st_));
    // GEGLU: the projection's epilogue applies value * gelu(gate) (interleaved weight rows, see Loader::linear_glu); the
    // projection itself is only written when the backward needs it
    H16 g16 = new16(B, H, W, 4 * C);
    // S2I_GLU_FUSION=2: fuse only in forwards that save nothing for a backward (there the epilogue writes the gated output alone)
    if (fuse_glu_ == 1 || (fuse_glu_ == 2 && !save_)) {
        if (save_) sv.ff = new16(B, H, W, 8 * C);
        if (!dry_) {
            ARENA_CHECK();
            GemmDesc d;
            d.tag = "gemm_linear";
            d.A = l16c.p; d.aC = C; d.aW = (int)rows; d.aH = 1; d.aB = 1; d.a_sw = l16c.ld;
            d.B = T.ff1.w; d.bI = C; d.bR = 8 * C; d.b_sr = C;
            d.b_static = 1;
            d.N = 8 * C; d.Kc = C;
            d.bias = T.ff1.b;
            if (save_) {
                d.out16 = sv.ff.p;
                d.ld16 = sv.ff.ld;
            }
            d.out_glu = g16.p;
            d.ld_glu = g16.ld;
            S2I_TRY(gemm_launch(d, st_));
        }
    } else {
        sv.ff = new16(B, H, W, 8 * C);
        S2I_TRY(gemm(l16c, false, 1, T.ff1.w, C, 8 * C, C, T.ff1.b, nullptr, nullptr, nullptr, &sv.ff));
        RUN(geglu_fwd(sv.ff.p, sv.ff.ld, rows, 4 * C, g16.p, g16.ld, st_));
    }
    H16 t3 = new16(B, H, W, C);
    S2I_TRY(gemm(g16, false, 1, T.ff2.w, 4 * C, C, 4 * C, T.ff2.b, nullptr, &sv.t2, nullptr, &t3));
    out = out32(B, H, W, C);
    S2I_TRY(gemm(t3, false, 1, T.proj_out.w, C, C, C, T.proj_out.b, nullptr, &x, &out, nullptr, true));
    if (keep_debug) {
        const std::string pre = "t" + std::to_string(idx);
        debug[pre + ".t0"] = sv.t0;
        debug[pre + ".t1"] = sv.t1;
        debug[pre + ".t2"] = sv.t2;
        debug[pre + ".out"] = out;
    }
    if (save_) tsave_[idx] = sv;
    return 0;
}

int UNet::transformer_bwd(int idx, const F32& dout, F32& dx) {
    const Transformer& T = tfm_[idx];
    const TfmSave S = sub(tsave_[idx], T);
    const int B = dout.B, H = dout.H, W = dout.W, HW = H * W, C = T.C;
    const long rows = dout.rows();
    H16 d16 = half_of(dout);
    if (!d16.p) return S2I_ERR_CUDA;
    F32 dt3 = new32(B, H, W, C);
    H16 dt3h = new16(B, H, W, C);
    S2I_TRY(gemm(d16, false, 1, T.proj_out.wd, C, C, C, nullptr, nullptr, nullptr, &dt3, &dt3h));
    // feed-forward
    F32 dg = new32(B, H, W, 4 * C);
    S2I_TRY(gemm(dt3h, false, 1, T.ff2.wd, C, 4 * C, C, nullptr, nullptr, nullptr, &dg, nullptr));
    H16 dff = new16(B, H, W, 8 * C);
    RUN(geglu_bwd(dg.p, dg.ld, S.ff.p, S.ff.ld, rows, 4 * C, dff.p, dff.ld, st_));
    F32 dl3 = new32(B, H, W, C);
    S2I_TRY(gemm(dff, false, 1, T.ff1.wd, 8 * C, C, 8 * C, nullptr, nullptr, nullptr, &dl3, nullptr));
    F32 dt2 = new32(B, H, W, C);
    H16 dt2h = new16(B, H, W, C);
    RUN(ln_bwd(dl3.p, dl3.ld, S.t2.p, S.t2.ld, rows, C, T.ln3.g, S.l3, dt3.p, dt3.ld, dt2.p, dt2.ld, dt2h.p, dt2h.ld, st_));
    // cross attention (only dQ: the text context needs no gradient)
    H16 dO2 = new16(B, H, W, T.HP);
    S2I_TRY(gemm(dt2h, false, 1, T.o2.wd, C, T.HP, C, nullptr, nullptr, nullptr, nullptr, &dO2));
    H16 dq2 = new16(B, H, W, T.HP);
    S2I_TRY(attention_bwd(T, dO2, S.q2, 0, S.kv2, 0, T.HP, cfg.ctx_len, S.P2, S.o2, S.lse2, dq2, 0, nullptr, 0, 0));
    F32 dl2 = new32(B, H, W, C);
    S2I_TRY(gemm(dq2, false, 1, T.q2.wd, T.HP, C, T.HP, nullptr, nullptr, nullptr, &dl2, nullptr));
    F32 dt1 = new32(B, H, W, C);
    H16 dt1h = new16(B, H, W, C);
    RUN(ln_bwd(dl2.p, dl2.ld, S.t1.p, S.t1.ld, rows, C, T.ln2.g, S.l2, dt2.p, dt2.ld, dt1.p, dt1.ld, dt1h.p, dt1h.ld, st_));
    // self attention
    H16 dO1 = new16(B, H, W, T.HP);
    S2I_TRY(gemm(dt1h, false, 1, T.o1.wd, C, T.HP, C, nullptr, nullptr, nullptr, nullptr, &dO1));
    H16 dqkv = new16(B, H, W, 3 * T.HP);
    S2I_TRY(attention_bwd(T, dO1, S.qkv, 0, S.qkv, T.HP, 2L * T.HP, HW, S.P1, S.o1, S.lse1, dqkv, 0, &dqkv, T.HP, 2L * T.HP));
    F32 dl1 = new32(B, H, W, C);
    S2I_TRY(gemm(dqkv, false, 1, T.qkv.wd, 3 * T.HP, C, 3 * T.HP, nullptr, nullptr, nullptr, &dl1, nullptr));
    H16 dt0h = new16(B, H, W, C);
    RUN(ln_bwd(dl1.p, dl1.ld, S.t0.p, S.t0.ld, rows, C, T.ln1.g, S.l1, dt1.p, dt1.ld, nullptr, 0, dt0h.p, dt0h.ld, st_));
    F32 dn = new32(B, H, W, C);
    S2I_TRY(gemm(dt0h, false, 1, T.proj_in.wd, C, C, C, nullptr, nullptr, nullptr, &dn, nullptr));
    double* bs = new_stats();
    dx = new32(B, H, W, C);
    H16 dxh = new16(B, H, W, C);
    RUN(gn_backward(dn.p, dn.ld, S.x.p, S.x.ld, B, HW, C, S.gs, bs, T.gn.g, T.gn.b, T.gn.eps, 0, dout.p, dout.ld, dx.p,
                     dx.ld, dxh.p, dxh.ld, st_));
    dx.h = dxh.p;
    dx.hld = dxh.ld;
    return 0;
}

// ================================================================================================== whole network
static const int kStatsSlots = 192;

int UNet::run_forward(const float* x_nchw, float t, float* eps_nchw) {
    const int B = B_, H = H_, W = W_;
    const int* boc = cfg.boc;
    arena_.reset();
    dest_set_ = false;
    stats_off_ = 0;
    stats_cap_ = (size_t)kStatsSlots * B * kGroups * 2;
    stats_ = dalloc<double>(stats_cap_);
    if (!dry_) S2I_MEMOP(cudaMemsetAsync(stats_, 0, stats_cap_ * sizeof(double), st_));
    debug.clear();

    // time embedding -> fused per-resnet projections (persistent buffer; see prepare_time)
    if (!dry_ && !time_ready_) S2I_TRY(prepare_time(t, st_));

    // text context as fp16 GEMM operand
    const bool enc = cfg.encoder_only;
    if (!enc) {
        ctx16_ = new16(B, 1, cfg.ctx_len, cfg.cross_dim);
        if (!reuse_kv_) RUN(cast2d(ctx_, cfg.cross_dim, (long)B * cfg.ctx_len, cfg.cross_dim, 1.f, ctx16_.p, ctx16_.ld, st_));
    }

    // Up-path concat buffers, planned before anything runs: the k-th up resnet reads cat([h, skip]) (diffusers
    // UpBlock2D / CrossAttnUpBlock2D), skip = the (K-1-k)-th tensor the down path pushed.  Both producers write straight
    // into their half of the buffer (row stride = the concatenated width), so the concat costs no launch.
    const int K = 4 * (cfg.layers + 1);
    std::vector<F32> cats(K);
    std::vector<int> cat_h(K);
    if (!enc) {
        std::vector<int> sC, sH, sW;                  // skip channels / spatial size, push order
        int hh = H, ww = W;
        sC.push_back(boc[0]); sH.push_back(hh); sW.push_back(ww);
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < cfg.layers; ++j) { sC.push_back(boc[i]); sH.push_back(hh); sW.push_back(ww); }
            if (i < 3) { hh /= 2; ww /= 2; sC.push_back(boc[i]); sH.push_back(hh); sW.push_back(ww); }
        }
        if ((int)sC.size() != K) return set_error(S2I_ERR_STATE, "unet: skip plan out of sync (%d vs %d)", (int)sC.size(), K);
        int hC = boc[3], sp0 = K;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < cfg.layers + 1; ++j) {
                const int k = i * (cfg.layers + 1) + j;
                --sp0;
                cat_h[k] = hC;
                cats[k] = new32(B, sH[sp0], sW[sp0], hC + sC[sp0]);
                hC = boc[3 - i];
            }
    }
    // skip n (push order) is the second half of cats[K - 1 - n]
    auto reserve_skip = [&](int n) {
        if (!enc) reserve(cats[K - 1 - n], cat_h[K - 1 - n], cats[K - 1 - n].C - cat_h[K - 1 - n]);
    };

    // conv_in (im2col GEMM, K = 9*in_ch padded to 64)
    F32 x = new32(B, H, W, cfg.in_ch);
    RUN(nchw_to_nhwc(x_nchw, B, cfg.in_ch, H, W, x.p, x.ld, st_));
    H16 col = new16(B, H, W, 64);
    RUN(im2col3x3(x.p, x.ld, B, H, W, cfg.in_ch, 1, col.p, col.ld, st_));
    skips_.clear();
    reserve_skip(0);
    F32 h = out32(B, H, W, boc[0]);
    S2I_TRY(gemm(col, false, 1, conv_in_.w, 64, boc[0], 64, conv_in_.b, nullptr, nullptr, &h, nullptr, true));
    if (keep_debug) debug["conv_in"] = h;

    skips_.push_back(h);
    int ri = 0, ti = 0;
    // ---- down
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < cfg.layers; ++j) {
            F32 o;
            if (i == 3) reserve_skip((int)skips_.size());        // no transformer follows: the resnet output is the skip
            S2I_TRY(resblock(ri++, h, o));
            h = o;
            if (i < 3 && !enc) {
                reserve_skip((int)skips_.size());
                S2I_TRY(transformer(ti++, h, o));
                h = o;
            }
            skips_.push_back(h);
        }
        if (i < 3) {
            const int Ho = h.H / 2, Wo = h.W / 2, C = h.C;
            H16 c2 = new16(B, Ho, Wo, 9 * C);
            RUN(im2col3x3(h.p, h.ld, B, h.H, h.W, C, 2, c2.p, c2.ld, st_));
            reserve_skip((int)skips_.size());
            F32 o = out32(B, Ho, Wo, C);
            S2I_TRY(gemm(c2, false, 1, down_[i].w, 9L * C, C, 9 * C, down_[i].b, nullptr, nullptr, &o, nullptr, true));
            h = o;
            skips_.push_back(h);
            taps[i] = h;
        }
        if (keep_debug) debug["down" + std::to_string(i)] = h;
    }
    if (enc) {
        // sketch_encoder.py:93-98: the forward ends here; res_samples = everything the down blocks pushed
        res_samples.assign(skips_.begin() + 1, skips_.end());
        if (stats_off_ > stats_cap_) return set_error(S2I_ERR_STATE, "GroupNorm statistics arena overflow");
        return 0;
    }
    // ---- mid
    {
        F32 o;
        S2I_TRY(resblock(ri++, h, o));
        h = o;
        taps[4] = h;
        S2I_TRY(transformer(ti++, h, o));
        h = o;
        taps[3] = h;
        reserve(cats[0], 0, cat_h[0]);                           // the mid block's output is the first half of cats[0]
        S2I_TRY(resblock(ri++, h, o));
        h = o;
        taps[5] = h;
        if (keep_debug) debug["mid"] = h;
    }
    // ---- up
    int sp = (int)skips_.size();
    up_cat_.clear();
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < cfg.layers + 1; ++j) {
            const int k = i * (cfg.layers + 1) + j;
            const F32& sk = skips_[--sp];
            F32 cat = cats[k];
            if (cat.C != h.C + sk.C || cat.H != h.H || h.p != cat.p || sk.p != cat.p + h.C)
                return set_error(S2I_ERR_STATE, "unet: concat plan out of sync at up resnet %d", k);
            up_cat_.push_back({h.C, sp});
            // the concat's statistics: the first half from whatever produced h, the second from the skip's producer
            cat.nst = 0;
            if (h.nst == 1 && sk.nst == 1) {
                cat.nst = 2;
                cat.st[0] = h.st[0];
                cat.st[1] = sk.st[0];
                cat.st[1].c0 = h.C;
                cat.st[1].c1 = h.C + sk.C;
            }
            // whichever call ends this iteration produces the first half of the next concat buffer
            const bool ups = j == cfg.layers && i < 3, tfm = i > 0, more = k + 1 < K;
            F32 o;
            if (more && !tfm && !ups) reserve(cats[k + 1], 0, cat_h[k + 1]);
            S2I_TRY(resblock(ri++, cat, o));
            h = o;
            if (tfm) {
                if (more && !ups) reserve(cats[k + 1], 0, cat_h[k + 1]);
                S2I_TRY(transformer(ti++, h, o));
                h = o;
            }
        }
        if (i < 3) {
            H16 u = new16(B, 2 * h.H, 2 * h.W, h.C);
            RUN(upsample2x(h.p, h.ld, B, h.H, h.W, h.C, u.p, u.ld, st_));
            reserve(cats[(i + 1) * (cfg.layers + 1)], 0, cat_h[(i + 1) * (cfg.layers + 1)]);
            F32 o = out32(B, 2 * h.H, 2 * h.W, h.C);
            S2I_TRY(gemm(u, true, 9, up_[i].w, 9L * h.C, h.C, h.C, up_[i].b, nullptr, nullptr, &o, nullptr, true));
            h = o;
            taps[6 + i] = h;
        }
        if (keep_debug) debug["up" + std::to_string(i)] = h;
    }
    // ---- out
    double* so = new_stats();
    H16 a = new16(B, H, W, boc[0]);
    S2I_TRY(group_norm(h, boc[0], so, norm_out_.g, norm_out_.b, norm_out_.eps, 1, a, nullptr));
    F32 eps = new32(B, H, W, cfg.out_ch);
    S2I_TRY(gemm(a, true, 9, conv_out_.w, 9L * boc[0], cfg.out_ch, boc[0], conv_out_.b, nullptr, nullptr, &eps, nullptr));
    RUN(nhwc_to_nchw(eps.p, eps.ld, B, cfg.out_ch, H, W, eps_nchw, st_));
    if (stats_off_ > stats_cap_) return set_error(S2I_ERR_STATE, "GroupNorm statistics arena overflow");
    return 0;
}

int UNet::run_backward(float* const tap_grads[9], float* dx_nchw) {
    const int B = bnb_;
    const size_t bstat_cap = (size_t)kStatsSlots * B_ * kGroups * 2;   // new_stats() hands out B_-sample slots
    double* save_stats = stats_;
    size_t save_off = stats_off_;
    stats_ = dalloc<double>(bstat_cap);
    stats_off_ = 0;
    if (!dry_) S2I_MEMOP(cudaMemsetAsync(stats_, 0, bstat_cap * sizeof(double), st_));

    auto tapg = [&](int k) {
        F32 g = taps[k];
        g.B = B;
        g.p = tap_grads ? tap_grads[k] : nullptr;
        if (dry_) g.p = reinterpret_cast<float*>(256);
        g.ld = g.C;
        return g;
    };
    std::vector<F32> dskip(skips_.size());
    F32 d;
    const int nres = (int)res_.size(), ntf = (int)tfm_.size();
    // forward order: res: down (8), mid (2), up (12); tfm: down (6), mid (1), up (9)
    int ri = nres - (cfg.layers + 1);   // first resnet of up block 3 (off the backward path)
    int ti = ntf - (cfg.layers + 1);
    int ci = (int)up_cat_.size() - (cfg.layers + 1);
    for (int i = 2; i >= 0; --i) {
        S2I_TRY(accumulate(d, tapg(6 + i)));
        {   // upsampler backward: conv dgrad then 2x2 sum-pool
            H16 d16 = half_of(d);
            if (!d16.p) return S2I_ERR_CUDA;
            F32 du = new32(B, d.H, d.W, d.C);
            S2I_TRY(gemm(d16, true, 9, up_[i].wd, 9L * d.C, d.C, d.C, nullptr, nullptr, nullptr, &du, nullptr));
            F32 dn = new32(B, d.H / 2, d.W / 2, d.C);
            RUN(sumpool2x(du.p, du.ld, B, dn.H, dn.W, dn.C, dn.p, dn.ld, st_));
            d = dn;
        }
        for (int j = cfg.layers; j >= 0; --j) {
            F32 o;
            if (i > 0) {
                S2I_TRY(transformer_bwd(--ti, d, o));
                d = o;
            }
            S2I_TRY(resblock_bwd(--ri, d, o));
            const UpCat& uc = up_cat_[--ci];
            F32 dsk = o;
            dsk.p = o.p + uc.ch;
            if (o.h) dsk.h = o.h + uc.ch;
            dsk.C = o.C - uc.ch;
            dskip[uc.skip] = dsk;
            d = o;
            d.C = uc.ch;
        }
    }
    // mid
    S2I_TRY(accumulate(d, tapg(5)));
    {
        F32 o;
        S2I_TRY(resblock_bwd(--ri, d, o));
        d = o;
        S2I_TRY(accumulate(d, tapg(3)));
        S2I_TRY(transformer_bwd(--ti, d, o));
        d = o;
        S2I_TRY(accumulate(d, tapg(4)));
        S2I_TRY(resblock_bwd(--ri, d, o));
        d = o;
    }
    // down
    int sk = (int)skips_.size() - 1;
    for (int i = 3; i >= 0; --i) {
        if (i < 3) {
            S2I_TRY(accumulate(d, dskip[sk--]));
            S2I_TRY(accumulate(d, tapg(i)));
            H16 z = new16(B, 2 * d.H, 2 * d.W, d.C);
            RUN(zero_insert2x(d.p, d.ld, B, d.H, d.W, d.C, z.p, z.ld, st_));
            F32 o = new32(B, 2 * d.H, 2 * d.W, d.C);
            S2I_TRY(gemm(z, true, 9, down_[i].wd, 9L * d.C, d.C, d.C, nullptr, nullptr, nullptr, &o, nullptr));
            d = o;
        }
        for (int j = cfg.layers - 1; j >= 0; --j) {
            S2I_TRY(accumulate(d, dskip[sk--]));
            F32 o;
            if (i < 3) {
                S2I_TRY(transformer_bwd(--ti, d, o));
                d = o;
            }
            S2I_TRY(resblock_bwd(--ri, d, o));
            d = o;
        }
    }
    S2I_TRY(accumulate(d, dskip[0]));
    // conv_in backward
    H16 d16 = half_of(d);
    if (!d16.p) return S2I_ERR_CUDA;
    F32 dx = new32(B, d.H, d.W, cfg.in_ch);
    S2I_TRY(gemm(d16, true, 9, conv_in_d_.wd, 9L * d.C, cfg.in_ch, d.C, nullptr, nullptr, nullptr, &dx, nullptr));
    RUN(nhwc_to_nchw(dx.p, dx.ld, B, cfg.in_ch, d.H, d.W, dx_nchw, st_));
    if (ri != 0 || ti != 0) return set_error(S2I_ERR_STATE, "backward walk out of sync (ri=%d ti=%d)", ri, ti);
    if (stats_off_ > bstat_cap) return set_error(S2I_ERR_STATE, "GroupNorm backward statistics arena overflow");
    stats_ = save_stats;
    stats_off_ = save_off;
    return 0;
}

int UNet::prepare_time(float t, cudaStream_t st) {
    if (!loaded_) return set_error(S2I_ERR_STATE, "unet: weights not loaded");
    const int* boc = cfg.boc;
    const int temb_dim = boc[0] * 4;
    const size_t bytes = (size_t)temb_all_.N * sizeof(float);
    const bool integral = t == (float)(int)t && t >= 0.f && t < 100000.f;
    if (integral) {
        auto it = temb_cache_.find((int)t);
        if (it != temb_cache_.end()) {
            S2I_MEMOP(cudaMemcpyAsync(temb_, it->second, bytes, cudaMemcpyDeviceToDevice, st));
            return 0;
        }
    }
    S2I_TRY(timestep_embedding(t, boc[0], te0_, st));
    S2I_TRY(gemv(te0_, boc[0], time1_.w, time1_.b, temb_dim, 0, te1_, st));
    S2I_TRY(gemv(te1_, temb_dim, time2_.w, time2_.b, temb_dim, 1, te2_, st));
    S2I_TRY(gemv(te2_, temb_dim, temb_all_.w, temb_all_.b, temb_all_.N, 1, temb_, st));
    if (integral && temb_cache_.size() < 1024) {
        void* c = nullptr;
        if (cudaMalloc(&c, bytes) == cudaSuccess) {
            owned_.push_back(c);
            temb_cache_[(int)t] = static_cast<float*>(c);
            S2I_MEMOP(cudaMemcpyAsync(c, temb_, bytes, cudaMemcpyDeviceToDevice, st));
        } else {
            cudaGetLastError();
        }
    }
    return 0;
}

int UNet::forward(const float* x_nchw, int B, int H, int W, float t, const float* ctx, float* eps_nchw,
                  bool save_for_backward, cudaStream_t st, bool time_ready, bool reuse_ctx_kv) {
    if (!loaded_) return set_error(S2I_ERR_STATE, "unet: weights not loaded");
    time_ready_ = time_ready;
    reuse_kv_ = reuse_ctx_kv && kv_cache_B_ == B;      // a cache filled for another batch size is not reusable
    kv_cache_B_ = B;
    if (!reuse_kv_) ++kv_gen_;                         // this forward rewrites the cached context K/V projections
    if (H % 8 || W % 8) return set_error(S2I_ERR_ARG, "unet: latent H, W must be multiples of 8 (got %d x %d)", H, W);
    if (cfg.encoder_only && save_for_backward) return set_error(S2I_ERR_ARG, "sketch encoder: forward only");
    st_ = st;
    B_ = B; H_ = H; W_ = W;
    ctx_ = ctx;
    save_ = save_for_backward;
    have_saved_ = false;
    const long key = ((long)B << 40) ^ ((long)H << 24) ^ ((long)W << 8) ^ (save_ ? 1 : 0);
    const unsigned sat_sig = sat_signature();
    if (key != arena_key_ || sat_sig != arena_sat_) {
        // measure the footprint of this configuration with a dry run, then (re)allocate
        Arena real = arena_;
        arena_ = Arena();
        dry_ = true;
        int rc = run_forward(nullptr, t, nullptr);
        bb0_ = 0;
        bnb_ = B;      // sized for a backward over every sample (a sub-range needs less)
        if (rc == 0 && save_) rc = run_backward(nullptr, nullptr);
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
                return set_error(S2I_ERR_OOM, "unet: cannot allocate %.1f GB activation arena", need / 1e9);
            }
            ++g_alloc_gen;
            arena_.base = static_cast<char*>(p);
            arena_.cap = need;
        }
        arena_key_ = key;
        arena_sat_ = sat_sig;
    }
    int rc = run_forward(x_nchw, t, eps_nchw);
    if (rc == 0 && arena_.overflow)
        rc = set_error(S2I_ERR_STATE, "unet: activation arena overflow (%zu of %zu bytes)", arena_.peak, arena_.cap);
    if (rc == 0 && save_) have_saved_ = true;
    return rc;
}

int UNet::backward(float* const tap_grads[9], float* dx_nchw, cudaStream_t st, int b0, int nb) {
    if (!have_saved_) return set_error(S2I_ERR_STATE, "unet backward: no forward with save_for_backward precedes it");
    if (nb < 0) nb = B_ - b0;
    if (b0 < 0 || nb < 1 || b0 + nb > B_)
        return set_error(S2I_ERR_ARG, "unet backward: samples [%d, %d) are not inside the forward's batch of %d", b0, b0 + nb, B_);
    st_ = st;
    have_saved_ = false;
    bb0_ = b0;
    bnb_ = nb;
    int rc = run_backward(tap_grads, dx_nchw);
    if (rc == 0 && arena_.overflow)
        rc = set_error(S2I_ERR_STATE, "unet: activation arena overflow in the backward (%zu of %zu bytes)", arena_.peak, arena_.cap);
    return rc;
}

}  // namespace s2i
