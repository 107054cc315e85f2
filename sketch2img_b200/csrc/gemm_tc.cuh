// tcgen05 + TMA implicit-GEMM used by every dense contraction on the sketch-guided sampling path:
// conv3x3 / conv1x1 / Linear (forward and input-gradient), attention QK^T / PV and their backward
// products, and the LGP MLP.  fp16 or bf16 operands, fp32 accumulation in TMEM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace s2i {

// Host-side description of one (possibly batched) GEMM:  C[z] = alpha * A[z] * B[z]^T (+ epilogue).
//
// A operand, K-major ("pixel rows"):  a 4-D tensor (C, W, H, B) with C contiguous.  Row m of the GEMM is
//   pixel (b, y, x); K runs over `taps` shifted windows (3x3, pad 1) times C.  Linear layers are the
//   degenerate case W = rows, H = B = 1, taps = 1.
// A operand, MN-major: the same 4-D tensor read as [K rows = W][M = C contiguous] (H = 1), batch on B.
// B operand: 3-D tensor (I, R, Z) with I contiguous.  K-major: R = N rows, I = K.  MN-major: R = K rows, I = N.
struct GemmDesc {
    // A
    const void* A = nullptr;
    int a_mn = 0;
    int aC = 0, aW = 0, aH = 1, aB = 1;   // extents
    long a_sw = 0, a_sh = 0, a_sb = 0;    // element strides of W, H, B
    int taps = 1;                         // 1 or 9 (K-major A only)
    int a_c0 = 0, a_hoff = 0, a_zmode = 0;
    // B
    const void* B = nullptr;
    int b_mn = 0;
    int bI = 0, bR = 0, bZ = 1;
    long b_sr = 0, b_sz = 0;
    int b_c0 = 0, b_hoff = 0, b_zmode = 0;
    // problem
    int N = 0;        // logical output columns
    int Kc = 0;       // K per tap
    int Z = 1, zh = 1;
    int bf16 = 0;     // operand element type: 0 = fp16, 1 = bf16
    int BN = 0;       // N tile (0 = choose)
    // epilogue:  v = alpha*acc + bias[n] + rowvec[sample][n];  v = relu ? max(v,0) : v;  v += residual[row][n]
    float alpha = 1.f;
    const float* bias = nullptr;
    const float* rowvec = nullptr;
    int rowvec_ld = 0;
    const float* residual = nullptr;
    long res_ld = 0;
    float* out32 = nullptr;
    long ld32 = 0;
    void* out16 = nullptr;   // fp16 (or bf16 if out16_bf16) copy of the result
    long ld16 = 0;
    int out16_bf16 = 0;
    long c_sb = 0, c_sh = 0;  // output element offsets per z-batch / z-head (applied to out32, out16, residual)
    int relu = 0;
};

// Enqueue on `stream`.  Returns 0 or a negative s2i error code (message via s2i_last_error()).
int gemm_launch(const GemmDesc& d, cudaStream_t stream);

// Number of kernel launches issued through gemm_launch since process start (bench.py's gpu_launches).
long gemm_launch_count();

}  // namespace s2i
