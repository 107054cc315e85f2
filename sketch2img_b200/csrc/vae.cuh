// AutoencoderKL (SD VAE) encode / decode on the UNet engine's kernels -- see vae.cu.
#pragma once
#include "unet.cuh"

namespace s2i {

struct VaeConfig {
    int in_ch = 3, out_ch = 3, latent = 4;
    int boc[4] = {128, 256, 512, 512};
    int layers = 2;
};

class VAE : public UNet {
  public:
    VaeConfig vcfg;
    explicit VAE(const VaeConfig& c);
    // diffusers-named host fp32 tensors (encoder.*, decoder.*, quant_conv.*, post_quant_conv.*)
    int load_vae(const std::map<std::string, HostParam>& params);
    // moments [B, 2 latent, H/8, W/8] (NCHW fp32: mean | logvar) = quant_conv(encoder(x));  x [B, in_ch, H, W] NCHW fp32
    int encode(const float* x_nchw, int B, int H, int W, float* moments, cudaStream_t st);
    // image [B, out_ch, 8h, 8w] = decoder(post_quant_conv(z));  z [B, latent, h, w] NCHW fp32
    int decode(const float* z_nchw, int B, int h, int w, float* image, cudaStream_t st);

  private:
    struct VAttn {
        Norm gn;
        Lin qkv;      // query | key | value fused: [3C][C] + bias [3C]
        Lin proj;     // proj_attn
        int C = 0;
    };
    Lin enc_in_, dec_in_;                       // conv_in as im2col GEMMs (K = 9 Cin padded to 64)
    std::vector<int> enc_res_, dec_res_;        // indices into res_ (ResBlock without time embedding)
    int enc_mid_[2] = {-1, -1}, dec_mid_[2] = {-1, -1};
    std::vector<Conv3> enc_down_, dec_up_;
    VAttn enc_attn_, dec_attn_;
    Norm enc_norm_, dec_norm_;
    Conv3 enc_out_, dec_out_;
    float *quant_w_ = nullptr, *quant_b_ = nullptr, *pq_w_ = nullptr, *pq_b_ = nullptr;
    bool vae_loaded_ = false;

    int begin_pass(int B);
    int vattn(const VAttn& A, const F32& x, F32& out);
    int run_encode(const float* x_nchw, int B, int H, int W, float* moments);
    int run_decode(const float* z_nchw, int B, int h, int w, float* image);
    template <class Body> int sized(long key, Body body);
};

}  // namespace s2i
