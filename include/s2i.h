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
    int N, Kc, Z, zh; int bf16; int BN;
    float alpha; const float* bias; const float* rowvec; int rowvec_ld;
    const float* residual; long long res_ld;
    float* out32; long long ld32; void* out16; long long ld16; int out16_bf16;
    long long c_sb, c_sh; int relu;
} s2i_gemm_desc;

int s2i_gemm(const s2i_gemm_desc* d, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* S2I_H_ */
