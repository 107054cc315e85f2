// tcgen05 + TMA implicit-GEMM used by every dense contraction on the sketch-guided sampling path:
// conv3x3 / conv1x1 / Linear (forward and input-gradient), attention QK^T / PV and their backward
// products, and the LGP MLP.  fp16 or bf16 operands, fp32 accumulation in TMEM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstring>

#include "../../include/s2i.h"

namespace s2i {

// Host-side description of one (possibly batched) GEMM:  C[z] = alpha * A[z] * B[z]^T (+ epilogue).
//
// A operand, K-major ("pixel rows"):  a 4-D tensor (C, W, H, B) with C contiguous.  Row m of the GEMM is
//   pixel (b, y, x); K runs over `taps` shifted windows (3x3, pad 1) times C.  Linear layers are the
//   degenerate case W = rows, H = B = 1, taps = 1.
// A operand, MN-major: the same 4-D tensor read as [K rows = W][M = C contiguous] (H = 1), batch on B.
// B operand: 3-D tensor (I, R, Z) with I contiguous.  K-major: R = N rows, I = K.  MN-major: R = K rows, I = N.
struct GemmDesc : s2i_gemm_desc {
    const char* tag = "gemm";   // profiler class of this launch (gemm_conv3x3 / gemm_linear / gemm_attn / gemm_lgp)
    GemmDesc() {
        memset(static_cast<s2i_gemm_desc*>(this), 0, sizeof(s2i_gemm_desc));
        taps = 1; aH = 1; aB = 1; bZ = 1; Z = 1; zh = 1; alpha = 1.f;
    }
    explicit GemmDesc(const s2i_gemm_desc& d) : s2i_gemm_desc(d) {}
};

// Enqueue on `stream`.  Returns 0 or a negative s2i error code (message via s2i_last_error()).
int gemm_launch(const GemmDesc& d, cudaStream_t stream);

// Bisecting switch: 0 routes every GEMM through gemm_tc_kernel (per-thread epilogue), 1 (default) lets eligible shapes
// use gemm_tma_kernel (epilogue through TMA).  Also settable with S2I_GEMM_TMA_EPI=0 in the environment.
void gemm_set_tma_epilogue(int on);

// True when S2I_SPLITK_ADD=1 selected the reduce-add split-K form (nondeterministic sum order) instead of the default cluster
// split-K (partial tiles reduced through distributed shared memory in rank order).
bool gemm_split_add_mode();

// Debugging: when set, gemm_tma_kernel stamps %globaltimer at its phase boundaries into buf[cta][16] (tools/gemm_trace.py).
void gemm_set_trace(unsigned long long* buf);

// Tools / tests: force one (1) or two (2) 128-row M sub-tiles per CTA in gemm_tma_kernel; 0 = the cost model decides.
void gemm_force_msub(int msub);
void gemm_set_pair(int mode);

// Split-K zero-fill plan.  A split-K GEMM reduce-adds its partial tiles into a zeroed output; zeroing each output with its
// own launch costs ~150 launches per denoising step.  The sampler therefore records the outputs while a step runs eagerly
// (RECORD: each GEMM still zeroes its own output and appends it to the plan) and, when the step is captured into a CUDA
// graph, zeroes all of them with ONE kernel at the top of the graph (APPLY: a GEMM whose output is in the plan skips its
// own zero-fill; anything else -- e.g. the shared split-K scratch, reused within a step -- keeps it).
struct ZeroRange {
    float* p;
    long ld, rows;
    int n4;
};
enum class ZeroMode { kOff, kRecord, kApply };
void gemm_zero_plan(ZeroMode mode, std::vector<ZeroRange>* plan);
// One launch that zeroes every range of a plan (ranges: device copy of the plan, n entries).
int gemm_zero_ranges(const ZeroRange* ranges, int n, cudaStream_t stream);

// Number of kernel launches issued through gemm_launch since process start (bench.py's gpu_launches).
long gemm_launch_count();

}  // namespace s2i
