// SD-family UNet2DCondition forward with the 9 LGP feature taps and the input-gradient backward from those
// taps (reference call sites: modules/pipeline.py:96 forward, :159 autograd.grad; tap set
// modules/latent_predictor.py:63-80).  Activations are NHWC / token-major: fp32 residual stream, fp16 GEMM
// operands; every contraction goes through gemm_tc.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace s2i {

struct UNetConfig {
    int in_ch = 4, out_ch = 4;
    int boc[4] = {320, 640, 1280, 1280};
    int heads[4] = {8, 8, 8, 8};
    int layers = 2;
    int cross_dim = 768;
    int sample_size = 64;
    int ctx_len = 77;
    // SketchEncoder (modules/sketch_encoder.py): conv_in + time embedding + four attention-free down blocks, forward stops
    // after the down path and hands out every block's res_samples.  The reference's forward calls the down blocks without
    // encoder_hidden_states (:93-95), which only executes for blocks without cross-attention (DownBlock2D).
    bool encoder_only = false;
};

struct HostParam {
    const float* data;
    std::vector<long> shape;
};

struct F32 {
    float* p = nullptr;
    long ld = 0;
    int B = 0, H = 0, W = 0, C = 0;
    // optional fp16 copy of the same values written by the producing kernel (backward pass: the next layer's GEMM operand,
    // so no separate cast launch); same shape, pixel stride hld
    __half* h = nullptr;
    long hld = 0;
    // per-channel partial sums left by the GEMM(s) that produced this tensor (kernels.cuh gn_norm): a following GroupNorm takes
    // its statistics from them instead of reading the tensor twice.  Two sources for an up-path concat buffer.
    GnStatSrc st[2];
    int nst = 0;
    long rows() const { return (long)B * H * W; }
};
struct H16 {
    __half* p = nullptr;
    long ld = 0;
    int B = 0, H = 0, W = 0, C = 0;
    long rows() const { return (long)B * H * W; }
};

struct Lin {     // y = x W^T + b;  w: [N][K] (forward), wd: [K][N] (input gradient)
    __half* w = nullptr;
    __half* wd = nullptr;
    float* b = nullptr;
    int N = 0, K = 0;
};
struct Conv3 {   // w: [Cout][9*Cin] (tap-major), wd: [Cin][9*Cout] (taps flipped)
    __half* w = nullptr;
    __half* wd = nullptr;
    float* b = nullptr;
    int Cin = 0, Cout = 0;
};
struct Norm {
    float* g = nullptr;
    float* b = nullptr;
    int C = 0;
    float eps = 1e-5f;
};
struct ResBlock {
    Norm n1, n2;
    Conv3 c1, c2;
    Lin sc;            // 1x1 shortcut when Cin != Cout
    bool has_sc = false;
    int temb_off = 0;  // offset of this block's time_emb_proj slice in the fused projection (< 0: no time embedding, VAE)
    int Cin = 0, Cout = 0;
};
// Injected sketch attention of one transformer block (reference: AttnModule, modules/sketch_guided_attn.py:46-132):
// after the self-attention residual,  h += scale * Conv1d_1x1( to_out( softmax(Q K^T / sqrt(d)) V ) ),
// Q = to_q(LayerNorm(h)), K / V = to_k / to_v of the sketch-encoder feature tokens of this block.
struct SatBlock {
    bool loaded = false;
    Norm ln;                   // sketch_norm
    Lin q, kv, o;              // to_q [HP][C]; to_k | to_v fused [2*HP][C]; to_out.0 [C][HP] + bias
    Lin conv;                  // sketch_conv as a Linear [C][C] + bias, multiplied by `scale` (fp16 weight, fp32 bias)
    float* conv_w32 = nullptr; // unscaled masters
    float* conv_b32 = nullptr;
    __half* kv16 = nullptr;    // cached K | V projections of the current feature tokens [B][N][2*HP]
    size_t kv_cap = 0;
    int fB = 0, fN = 0;        // batch / tokens of the cached feature (0 = no feature set: block runs unmodified)
};

struct Transformer {
    std::string path;          // diffusers module path, e.g. "down_blocks.0.attentions.1"
    SatBlock sat;
    Norm gn, ln1, ln2, ln3;
    Lin proj_in, proj_out;
    Lin qkv;   // fused self-attention q|k|v, head-padded: [3*HP][C]
    Lin o1;    // [C][HP]
    Lin q2;    // [HP][C]
    Lin kv2;   // [2*HP][Dctx]
    Lin o2;    // [C][HP]
    Lin ff1;   // [8C][C]
    Lin ff2;   // [C][4C]
    int C = 0, heads = 0, d = 0, dp = 0, HP = 0;
    // K | V projections of the text context [B][ctx_len][2*HP]: the context is the same at every denoising step of an
    // image, so they are computed on the first step and reused (UNet::forward's reuse_ctx_kv)
    __half* kv2_cache = nullptr;
    size_t kv2_cap = 0;
};

struct ResSave {
    F32 x, h1;
    double *s1 = nullptr, *s2 = nullptr;
};
struct TfmSave {
    F32 x, t0, t1, t2;
    H16 ff;      // GEGLU projection [rows][8C] (fp16: only ever consumed as a * gelu(gate) and its derivative)
    double* gs = nullptr;
    float *l1 = nullptr, *l2 = nullptr, *l3 = nullptr;
    H16 qkv, P1, q2, kv2, P2;
    // fused attention backward (attn_bwd.cu): the forward keeps its output and log-sum-exp instead of P
    H16 o1, o2;
    float *lse1 = nullptr, *lse2 = nullptr;
};

class Arena {
  public:
    char* base = nullptr;
    size_t cap = 0, off = 0, peak = 0;
    bool overflow = false;      // a real allocation ran past `cap` (the sizing pass saw another topology): callers must not launch
    void* alloc(size_t bytes) {
        off = (off + 255) & ~size_t(255);
        void* p = base ? base + off : reinterpret_cast<void*>(off + 256);   // fake pointers while measuring
        off += bytes;
        if (off > peak) peak = off;
        if (base && off > cap) {
            overflow = true;
            p = base;           // never hand out an address past the end
        }
        return p;
    }
    void reset() { off = 0; overflow = false; }
};

class UNet {
  public:
    UNetConfig cfg;
    explicit UNet(const UNetConfig& c) : cfg(c) {}
    ~UNet();

    // Weights: host fp32 tensors under their diffusers names (valid until load() returns).
    int load(const std::map<std::string, HostParam>& params);

    // eps[B,4,H,W] (NCHW fp32, device) = unet(x[B,4,H,W], t, ctx[B,ctx_len,cross_dim]).  With save_for_backward the
    // activations needed by backward() stay resident until the next forward().
    // reuse_ctx_kv: `ctx` holds the same values as in the previous forward of the same batch size -- skip the 16
    // cross-attention K/V projections and reuse the cached ones.
    int forward(const float* x_nchw, int B, int H, int W, float t, const float* ctx, float* eps_nchw,
                bool save_for_backward, cudaStream_t st, bool time_ready = false, bool reuse_ctx_kv = false);
    // The timestep-dependent part of the forward (sinusoidal embedding -> 2 Linear -> every ResBlock's time_emb_proj),
    // into a persistent buffer; results are cached per timestep.  forward(..., time_ready = true) then skips it, so the
    // rest of the step does not depend on t (the sampler replays it from a CUDA graph).
    int prepare_time(float t, cudaStream_t st);
    // dx[B,4,H,W] (NCHW fp32) = sum_k J_k^T tap_grad[k]; tap_grad[k] is NHWC fp32 shaped like tap(k).
    // (b0, nb): walk only the samples [b0, b0 + nb) of the forward's batch (nb < 0: through the last one) -- samples are
    // independent computations (GroupNorm per sample, LayerNorm per token, attention per sample), and the guided sampler
    // only keeps the cond half of the gradient (pipeline.py:159).  tap_grads / dx then hold nb samples.
    int backward(float* const tap_grads[9], float* dx_nchw, cudaStream_t st, int b0 = 0, int nb = -1);

    // Injected sketch attention (SatMixin): weights under the reference's names
    // "sketch_attn_<block path with '.' -> '_'>_transformer_blocks_0.{sketch_norm,sketch_attn.to_q,...,sketch_conv}.*",
    // one feature map (NCHW fp32 [B,C,H,W], B = the forward's batch) per transformer block, and the residual scale.
    int load_sat(const std::map<std::string, HostParam>& params);
    int set_sat_feature(const char* block_path, const float* nchw, int B, int C, int H, int W, cudaStream_t st);
    int set_sat_scale(float scale, cudaStream_t st);

    // The 9 taps of the last forward (NHWC fp32), hook order of latent_predictor.py:63-80.
    F32 taps[9];
    // encoder_only: the res_samples of the last forward in block order (sketch_encoder.py:93-96):
    // per down block its `layers` resnet outputs, then the downsampled map (all but the last block)
    std::vector<F32> res_samples;
    // Named intermediates of the last forward (debugging / parity bisecting).
    std::map<std::string, F32> debug;
    bool keep_debug = false;
    bool use_flash_ = true;   // fused attention forward wherever P is not needed afterwards (S2I_NO_FLASH=1 disables)
    // GEGLU inside the projection's epilogue (gemm_tma_kernel<2, true>) instead of a separate geglu_fwd launch: parity-green
    // but measured SLOWER on B200 (448 vs 442 ms/image) -- 10 M erf evaluations per level-0 projection land on the 8 epilogue
    // warps of each CTA instead of a full-occupancy elementwise kernel -- so it is off unless S2I_GLU_FUSION=1
    int fuse_glu_ = 1;               // gated-GELU in the ff1 epilogue: 0 never (separate geglu kernel), 1 always, 2 only when nothing is saved for a backward
    bool spatial_stats_ = true;      // GroupNorm statistics from the producing GEMM's epilogue (S2I_GN_FUSED_STATS=0: off)

    size_t arena_bytes() const { return arena_.cap; }
    // Which transformer blocks currently have a sketch feature (bit i = block i in load order): part of the activation
    // arena's sizing key and of the sampler's graph key -- a block with a feature runs six more tensors and four more launches.
    unsigned sat_signature() const;
    // Bumped by every forward that (re)computes the cached context K/V projections: the sampler compares it with the value it
    // saw after its own last forward before it reuses the cache (any other forward in between overwrote it).
    long kv_generation() const { return kv_gen_; }

  protected:      // the VAE engine (vae.cu) builds on the same arena / GEMM / GroupNorm / ResBlock / attention plumbing
    // parameters
    Lin conv_in_;       // as im2col GEMM: w [C0][64]
    Conv3 conv_in_d_;   // dgrad form
    Lin time1_, time2_, temb_all_;
    std::vector<ResBlock> res_;
    std::vector<Transformer> tfm_;
    std::vector<Conv3> down_, up_;
    Norm norm_out_;
    Conv3 conv_out_;
    std::vector<void*> owned_;
    bool loaded_ = false;

    // execution state
    Arena arena_;
    bool dry_ = false;
    bool save_ = false;
    cudaStream_t st_ = nullptr;
    int B_ = 0, H_ = 0, W_ = 0;
    int bb0_ = 0, bnb_ = 0;      // sample range of the current backward
    const float* ctx_ = nullptr;
    float* temb_ = nullptr;      // fused time_emb_proj output [sum Cout] (persistent)
    float *te0_ = nullptr, *te1_ = nullptr, *te2_ = nullptr;
    std::map<int, float*> temb_cache_;   // timestep -> cached temb_ contents
    double* stats_ = nullptr;    // GroupNorm sums arena (zeroed once per pass)
    size_t stats_off_ = 0, stats_cap_ = 0;
    std::vector<ResSave> rsave_;
    std::vector<TfmSave> tsave_;
    std::vector<F32> skips_;
    struct UpCat { int ch; int skip; };   // per up-path resnet: width of the `h` half of its concat input, skip index
    std::vector<UpCat> up_cat_;
    H16 ctx16_;
    long arena_key_ = -1;
    unsigned arena_sat_ = 0;
    long kv_gen_ = 0;
    bool have_saved_ = false;
    bool time_ready_ = false;
    bool reuse_kv_ = false;
    int kv_cache_B_ = 0;
    float sat_scale_ = 1.f;

    int run_forward(const float* x_nchw, float t, float* eps_nchw);
    int run_backward(float* const tap_grads[9], float* dx_nchw);

    F32 new32(int B, int H, int W, int C);
    F32 out32(int B, int H, int W, int C);
    void reserve(const F32& cat, int c0, int C);
    F32 dest_;                   // reserved destination of the next layer output (see out32)
    bool dest_set_ = false;
    H16 new16(int B, int H, int W, int C);
    double* new_stats();
    template <class T> T* dalloc(size_t n) { return static_cast<T*>(arena_.alloc(n * sizeof(T))); }

    int k(int rc) { return rc; }
    int gemm(const H16& a, bool spatial, int taps, const __half* w, long w_ld, int N, int Kc, const float* bias,
             const float* rowvec, const F32* residual, F32* out32, H16* out16, bool stats = false);
    int group_norm(const F32& x, int C, double* slot, const float* gamma, const float* beta, float eps, int silu, H16& out,
                   H16* raw);
    int resblock(int idx, const F32& x, F32& out);
    int resblock_bwd(int idx, const F32& dout, F32& dx);
    int transformer(int idx, const F32& x, F32& out);
    int transformer_bwd(int idx, const F32& dout, F32& dx);
    // need_bwd: a backward through this attention follows -- keep the log-sum-exp (*lse, fused backward) or P (unfused)
    int attention(const Transformer& T, const H16& q, long q_c0, const H16& kv, long k_c0, long v_c0, int Nk, H16& P,
                  H16& o, bool need_bwd, float** lse);
    int attention_bwd(const Transformer& T, const H16& dO, const H16& q, long q_c0, const H16& kv, long k_c0, long v_c0,
                      int Nk, const H16& P, const H16& o, const float* lse, H16& dq, long dq_c0, H16* dkv, long dk_c0,
                      long dv_c0);
    int accumulate(F32& acc, const F32& g);
    H16 half_of(const F32& t);      // the fp16 copy of a backward tensor: its companion if the producer wrote one, else a cast
    F32 sub(const F32& t) const;
    H16 sub(const H16& t) const;
    ResSave sub(const ResSave& s) const;
    TfmSave sub(const TfmSave& s, const Transformer& T) const;
};

}  // namespace s2i
