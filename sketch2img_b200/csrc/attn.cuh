// Fused attention forward on tcgen05 + TMA (see attn.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace s2i {

// O[b, i, h*dp : h*dp+dp] = softmax_j(scale * <Q[b,i,h], K[b,j,h]>) V[b,j,h]
//   q : fp16 [B][Nq][ldq],  head h of Q at columns q_c0 + h*dp
//   kv: fp16 [B][Nk][ldkv], head h of K at k_c0 + h*dp and of V at v_c0 + h*dp
//   out: fp16 [B][Nq][ldo], head h at columns h*dp.  dp = head dim padded to a multiple of 16 (padding columns are
//   zero in Q/K/V); d_true = the real head dim (profiling only).
struct AttnDesc {
    const __half* q = nullptr;
    long ldq = 0;
    int q_c0 = 0;
    const __half* kv = nullptr;
    long ldkv = 0;
    int k_c0 = 0, v_c0 = 0;
    int B = 1, heads = 1, Nq = 0, Nk = 0, dp = 0, d_true = 0;
    float scale = 1.f;
    __half* out = nullptr;
    long ldo = 0;
    float* lse = nullptr;   // optional [B*heads][Nq]
};

bool attn_fwd_supported(int Nq, int Nk, int dp);
int attn_fwd_launch(const AttnDesc& d, cudaStream_t stream);

}  // namespace s2i
