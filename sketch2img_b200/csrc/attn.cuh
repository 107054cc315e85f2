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

// Fused backward (attn_bwd.cu): dQ (and dK, dV when `dk` is set) of the attention above from Q, K, V, the forward's
// output O and log-sum-exp, and dO -- scores and probabilities are recomputed on chip.
//   dO, o: fp16 [B][Nq][ld], head h at columns h*dp;  lse: fp32 [B*heads][Nq] (forward);  delta: fp32 scratch [B*heads][Nq]
//   dq: fp16 [B][Nq][lddq], head h at dq_c0 + h*dp;  dk / dv: fp16 [B][Nk][lddkv], heads at dk_c0 / dv_c0 + h*dp
struct AttnBwdDesc {
    const __half* q = nullptr;
    long ldq = 0;
    int q_c0 = 0;
    const __half* kv = nullptr;
    long ldkv = 0;
    int k_c0 = 0, v_c0 = 0;
    const __half* dO = nullptr;
    long lddo = 0;
    const __half* o = nullptr;
    long ldo = 0;
    const float* lse = nullptr;
    float* delta = nullptr;
    int B = 1, heads = 1, Nq = 0, Nk = 0, dp = 0, d_true = 0;
    float scale = 1.f;
    __half* dq = nullptr;
    long lddq = 0;
    int dq_c0 = 0;
    __half* dk = nullptr;   // null: only dQ (cross-attention to a constant context)
    __half* dv = nullptr;
    long lddkv = 0;
    int dk_c0 = 0, dv_c0 = 0;
};

bool attn_bwd_supported(int Nq, int Nk, int dp);
int attn_bwd_launch(const AttnBwdDesc& d, cudaStream_t stream);

}  // namespace s2i
