// HBM-bound kernels of the sketch-guided sampling path (NHWC / token-major layouts, fp32 residual stream,
// fp16 GEMM operands).  Every launcher enqueues on `st` and returns 0 or a negative s2i error code.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace s2i {

constexpr int kGroups = 32;   // GroupNorm groups (norm_num_groups of every SD UNet)

// ---- GroupNorm (per sample, per group over (C/32) x HW), optionally fused with SiLU ------------------------
// sums: the call's statistics slot, 64 doubles (512 bytes) per sample, zero-initialised once by the caller:
// gn_stats leaves float2[32] = (mean, rstd) per sample in it (the rest is the reduction's self-resetting arrival counter).
int gn_stats(const float* x, long ldx, int B, int HW, int C, float eps, double* sums, cudaStream_t st);
// out16 = fp16( act( (x-mean)*rstd*gamma + beta ) ); raw16 (optional) = fp16(x)
int gn_apply(const float* x, long ldx, int B, int HW, int C, const double* sums, const float* gamma, const float* beta,
             float eps, int silu, void* out16, long ld16, void* raw16, long ldraw, cudaStream_t st);
// Backward.  bsums: a second slot of the same shape; receives float2[32] = (mean dxhat, mean dxhat*xhat) per sample.
int gn_bwd_stats(const float* dy, long ldd, const float* x, long ldx, int B, int HW, int C, const double* sums,
                 const float* gamma, const float* beta, float eps, int silu, double* bsums, cudaStream_t st);
// dx = rstd*(dxhat - mean(dxhat) - xhat*mean(dxhat*xhat)) (+ add);  written as fp32 (dx32) and/or fp16 (dx16)
int gn_bwd_apply(const float* dy, long ldd, const float* x, long ldx, int B, int HW, int C, const double* sums,
                 const double* bsums, const float* gamma, const float* beta, float eps, int silu, const float* add,
                 long ldadd, float* dx32, long ld32, void* dx16, long ld16, cudaStream_t st);

// Fused forms (one launch: reduction, grid-wide arrival, apply); same slot convention as above.
int gn_forward(const float* x, long ldx, int B, int HW, int C, double* slot, const float* gamma, const float* beta, float eps,
               int silu, void* out16, long ld16, void* raw16, long ldraw, cudaStream_t st);
// GroupNorm forward with the statistics taken from the producing GEMM (GemmDesc::colstat): per-channel partial sums
// [B][cap][2][ld] floats (sum, sum of squares), `bps` row blocks per sample, covering the tensor's channels [c0, c1).  One source
// per producer: two for a concatenated input.  Same outputs as gn_forward (slot included, for the backward).
struct GnStatSrc {
    const float* p = nullptr;
    int cap = 0, bps = 0;
    long ld = 0;
    int c0 = 0, c1 = 0;
};
bool gn_norm_supported(int C);
int gn_norm(const float* x, long ldx, int B, int HW, int C, const GnStatSrc* src, int nsrc, double* slot, const float* gamma,
            const float* beta, float eps, int silu, void* out16, long ld16, void* raw16, long ldraw, cudaStream_t st);
int gn_backward(const float* dy, long ldd, const float* x, long ldx, int B, int HW, int C, const double* fslot, double* bslot,
                const float* gamma, const float* beta, float eps, int silu, const float* add, long ldadd, float* dx32,
                long ld32, void* dx16, long ld16, cudaStream_t st);

// ---- LayerNorm over the last dim (one warp per row) --------------------------------------------------------
int ln_fwd(const float* x, long ldx, long rows, int C, const float* gamma, const float* beta, float eps, void* out16,
           long ld16, float* stats /*[rows][2] mean,rstd*/, cudaStream_t st);
int ln_bwd(const float* dy, long ldd, const float* x, long ldx, long rows, int C, const float* gamma,
           const float* stats, const float* add, long ldadd, float* dx32, long ld32, void* dx16, long ld16,
           cudaStream_t st);

// ---- softmax over the last dim (one warp per row) ----------------------------------------------------------
int softmax_fwd(const float* s, long lds, long rows, int n, void* p16, long ldp, cudaStream_t st);
// ds16 = scale * p * (dp - sum_j dp_j p_j)
int softmax_bwd(const void* p16, long ldp, const float* dp, long lddp, long rows, int n, float scale, void* ds16,
                long ldds, cudaStream_t st);

// ---- GEGLU: out = a * gelu(gate); ff is the fp16 output of the C -> 8C Linear with its columns interleaved in blocks of 32
// (value features 32 b .. at columns [64 b, 64 b + 32), their gates at [64 b + 32, 64 b + 64); UNet Loader::linear_glu) ----
int geglu_fwd(const void* ff16, long ldf, long rows, int F, void* out16, long ld16, cudaStream_t st);
int geglu_bwd(const float* dg, long ldg, const void* ff16, long ldf, long rows, int F, void* dff16, long ld16,
              cudaStream_t st);

// ---- layout / movement --------------------------------------------------------------------------------------
// dst32 = a (+ b);  optional fp16 copy.  2-D strided, cols % 4 == 0.
int add2d(const float* a, long lda, const float* b, long ldb, long rows, int cols, float* dst32, long ld32, void* dst16,
          long ld16, cudaStream_t st);
// fp32 -> fp16 with a scalar multiplier
int cast2d(const float* a, long lda, long rows, int cols, float mul, void* dst16, long ld16, cudaStream_t st);
// nearest 2x upsample: x fp32 [B,H,W,C] -> out16 [B,2H,2W,C]
int upsample2x(const float* x, long ldx, int B, int H, int W, int C, void* out16, long ld16, cudaStream_t st);
// its adjoint: d fp32 [B,2H,2W,C] -> dx fp32 [B,H,W,C] (2x2 sums)
int sumpool2x(const float* d, long ldd, int B, int H, int W, int C, float* dx, long ldx, cudaStream_t st);
// zero-insertion (adjoint geometry of a stride-2 conv): d fp32 [B,Ho,Wo,C] -> out16 [B,2Ho,2Wo,C], d at even sites
int zero_insert2x(const float* d, long ldd, int B, int Ho, int Wo, int C, void* out16, long ld16, cudaStream_t st);
// 3x3 patches (pad 1, given stride): x fp32 [B,H,W,C] -> col16 [B*Ho*Wo][ldcol], column = tap*C + c, zero padded
// pad = 0 (stride 2 only): padding on the bottom / right side only, the VAE encoder's downsampler
int im2col3x3(const float* x, long ldx, int B, int H, int W, int C, int stride, void* col16, long ldcol, cudaStream_t st,
              int pad = 1);
// NCHW fp32 <-> NHWC fp32 (latents in/out of the engine; 4 channels padded to ldn)
int nchw_to_nhwc(const float* src, int B, int C, int H, int W, float* dst, long ldn, cudaStream_t st);
int nhwc_to_nchw(const float* src, long ldn, int B, int C, int H, int W, float* dst, cudaStream_t st);

// ---- time embedding ------------------------------------------------------------------------------------------
// out[n] = bias[n] + sum_k W[n][k] * act(x[k]);  W fp16 [N][K]; act = SiLU if silu_in.  (one warp per n)
int gemv(const float* x, int K, const void* w16, const float* bias, int N, int silu_in, float* out, cudaStream_t st);
// [cos | sin] sinusoidal embedding of one timestep (flip_sin_to_cos=True, freq_shift=0), dim values
int timestep_embedding(float t, int dim, float* out, cudaStream_t st);

}  // namespace s2i
