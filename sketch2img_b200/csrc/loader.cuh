// Weight packing kernels and the host-side loader shared by the engines that take diffusers-named fp32 state dicts
// (unet.cu: UNet2DConditionModel / SketchEncoder; vae.cu: AutoencoderKL).  Header-only: every including .cu gets its own copy.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "unet.cuh"

namespace s2i {

namespace {

// ------------------------------------------------------------------------------------------ weight packing
__device__ __forceinline__ long padmap(long x, int d, int dp) {
    if (d == dp || dp == 0) return x;
    const long h = x / dp, j = x - h * dp;
    return j < d ? h * d + j : -1;
}

// out(r,c) = S(pm_r(r), pm_c(c))  (or S(pm_c(c), pm_r(r)) when transpose); S fp32 row-major with src_ld.
__global__ void pack2d_kernel(const float* __restrict__ src, long src_ld, int transpose, long R, long Cc, int d_r,
                              int dp_r, int d_c, int dp_c, __half* __restrict__ out, long out_ld) {
    const long total = R * Cc;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / Cc, c = idx - r * Cc;
        const long i = padmap(r, d_r, dp_r), j = padmap(c, d_c, dp_c);
        float v = 0.f;
        if (i >= 0 && j >= 0) v = transpose ? src[j * src_ld + i] : src[i * src_ld + j];
        out[r * out_ld + c] = __float2half_rn(v);
    }
}

// src [Co][Ci][3][3] -> fwd out[co*out_ld + tap*Ci + ci]   |   dgrad out[ci*out_ld + tap*Co + co] with flipped taps
__global__ void pack_conv_kernel(const float* __restrict__ src, int Co, int Ci, int dgrad, __half* __restrict__ out,
                                 long out_ld) {
    const long total = (long)Co * Ci * 9;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % 9);
        const long rest = idx / 9;
        const int ci = (int)(rest % Ci);
        const int co = (int)(rest / Ci);
        const float v = src[idx];
        if (!dgrad) {
            out[(long)co * out_ld + (long)kk * Ci + ci] = __float2half_rn(v);
        } else {
            out[(long)ci * out_ld + (long)(8 - kk) * Co + co] = __float2half_rn(v);
        }
    }
}

inline long rup(long a, long b) { return (a + b - 1) / b * b; }

// dst16 = fp16(scale * src32) (n elements); dstb = scale * srcb (nb elements)
__global__ void scale_pack_kernel(const float* __restrict__ src, long n, float scale, __half* __restrict__ dst,
                                  const float* __restrict__ srcb, int nb, float* __restrict__ dstb) {
    const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long i = i0; i < n; i += (long)gridDim.x * blockDim.x) dst[i] = __float2half_rn(scale * src[i]);
    for (long i = i0; i < nb; i += (long)gridDim.x * blockDim.x) dstb[i] = scale * srcb[i];
}

// NCHW fp32 -> token-major fp16 [B][H*W][C]
__global__ void nchw_to_tokens16_kernel(const float* __restrict__ src, int B, int C, int HW, __half* __restrict__ dst) {
    const long total = (long)B * HW * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const long t = idx / C;
        const int p = (int)(t % HW);
        const int b = (int)(t / HW);
        dst[idx] = __float2half_rn(src[((long)b * C + c) * HW + p]);
    }
}

}  // namespace

struct Loader {
    const std::map<std::string, HostParam>& params;
    std::vector<void*>& owned;
    float* staging = nullptr;
    size_t staging_elems = 0;
    std::string err;

    const HostParam* find(const std::string& name, size_t expect_elems) {
        auto it = params.find(name);
        if (it == params.end()) {
            err = "missing parameter " + name;
            return nullptr;
        }
        size_t n = 1;
        for (long s : it->second.shape) n *= (size_t)s;
        if (n != expect_elems) {
            err = "parameter " + name + " has " + std::to_string(n) + " elements, expected " + std::to_string(expect_elems);
            return nullptr;
        }
        return &it->second;
    }
    template <class T>
    T* dmalloc(size_t n, bool zero = false) {
        void* p = nullptr;
        if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) {
            err = "cudaMalloc failed for weights";
            return nullptr;
        }
        if (zero) cudaMemset(p, 0, n * sizeof(T));
        owned.push_back(p);
        return static_cast<T*>(p);
    }
    const float* stage(const HostParam* hp, size_t n) {
        if (n > staging_elems) {
            if (staging) cudaFree(staging);
            staging_elems = n;
            if (cudaMalloc(&staging, n * sizeof(float)) != cudaSuccess) {
                err = "cudaMalloc failed for staging";
                return nullptr;
            }
        }
        if (cudaMemcpy(staging, hp->data, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            err = "H2D copy failed";
            return nullptr;
        }
        return staging;
    }
    float* vec(const std::string& name, size_t n) {
        const HostParam* hp = find(name, n);
        if (!hp) return nullptr;
        float* d = dmalloc<float>(n);
        if (!d) return nullptr;
        cudaMemcpy(d, hp->data, n * sizeof(float), cudaMemcpyHostToDevice);
        return d;
    }
    bool norm(const std::string& pre, int C, float eps, Norm& n) {
        n.C = C;
        n.eps = eps;
        n.g = vec(pre + ".weight", C);
        n.b = vec(pre + ".bias", C);
        return n.g && n.b;
    }
    // Plain / head-padded linear.  pad_n: output features are (heads x d) -> (heads x dp); pad_k likewise for inputs.
    bool linear_into(const std::string& wname, int N, int K, int d, int dp, bool pad_n, bool pad_k, __half* w, long w_ld,
                     long w_row0, __half* wd, long wd_ld, long wd_col0, bool want_dgrad) {
        const HostParam* hp = find(wname, (size_t)N * K);
        if (!hp) return false;
        const float* s = stage(hp, (size_t)N * K);
        if (!s) return false;
        const long Np = pad_n ? (long)N / d * dp : N;
        const long Kp = pad_k ? (long)K / d * dp : K;
        pack2d_kernel<<<1024, 256>>>(s, K, 0, Np, Kp, pad_n ? d : 0, pad_n ? dp : 0, pad_k ? d : 0, pad_k ? dp : 0,
                                     w + w_row0 * w_ld, w_ld);
        if (want_dgrad)
            pack2d_kernel<<<1024, 256>>>(s, K, 1, Kp, Np, pad_k ? d : 0, pad_k ? dp : 0, pad_n ? d : 0, pad_n ? dp : 0,
                                         wd + wd_col0, wd_ld);
        return cudaDeviceSynchronize() == cudaSuccess;
    }
    bool linear(const std::string& pre, int N, int K, bool bias, Lin& l, bool want_dgrad = true) {
        l.N = N;
        l.K = K;
        l.w = dmalloc<__half>((size_t)N * K);
        if (want_dgrad) l.wd = dmalloc<__half>((size_t)N * K);
        if (!l.w || (want_dgrad && !l.wd)) return false;
        if (!linear_into(pre + ".weight", N, K, 0, 0, false, false, l.w, K, 0, l.wd, N, 0, want_dgrad)) return false;
        if (bias) {
            l.b = vec(pre + ".bias", N);
            if (!l.b) return false;
        }
        return true;
    }
    // GEGLU projection [8C][C]: rows interleaved in blocks of 32 -- packed rows [64 b, 64 b + 32) = value features 32 b ..,
    // [64 b + 32, 64 b + 64) = their gates (rows 4C + 32 b .. of the checkpoint) -- so a 64-column slice of the GEMM's
    // output tile holds matching value / gate pairs (gated-GELU epilogue of gemm_tma_kernel); wd permuted alike along K.
    bool linear_glu(const std::string& pre, int C, Lin& l) {
        const int N = 8 * C, K = C, F = 4 * C;
        if (F % 32) {
            err = "GEGLU width must be a multiple of 32";
            return false;
        }
        const HostParam* hw = find(pre + ".weight", (size_t)N * K);
        const HostParam* hb = hw ? find(pre + ".bias", (size_t)N) : nullptr;
        if (!hw || !hb) return false;
        std::vector<float> w((size_t)N * K), b(N);
        for (int r = 0; r < N; ++r) {
            const int blk = r / 64, j = r % 64;
            const int src = j < 32 ? blk * 32 + j : F + blk * 32 + (j - 32);
            memcpy(&w[(size_t)r * K], hw->data + (size_t)src * K, (size_t)K * sizeof(float));
            b[r] = hb->data[src];
        }
        l.N = N;
        l.K = K;
        l.w = dmalloc<__half>((size_t)N * K);
        l.wd = dmalloc<__half>((size_t)N * K);
        l.b = dmalloc<float>(N);
        if (!l.w || !l.wd || !l.b) return false;
        HostParam tmp{w.data(), {N, K}};
        const float* s = stage(&tmp, (size_t)N * K);
        if (!s) return false;
        pack2d_kernel<<<1024, 256>>>(s, K, 0, N, K, 0, 0, 0, 0, l.w, K);
        pack2d_kernel<<<1024, 256>>>(s, K, 1, K, N, 0, 0, 0, 0, l.wd, N);
        cudaMemcpy(l.b, b.data(), (size_t)N * sizeof(float), cudaMemcpyHostToDevice);
        return cudaDeviceSynchronize() == cudaSuccess;
    }
    bool conv3(const std::string& pre, int Co, int Ci, Conv3& c, bool want_dgrad = true, bool want_fwd = true) {
        c.Cin = Ci;
        c.Cout = Co;
        const size_t n = (size_t)Co * Ci * 9;
        const HostParam* hp = find(pre + ".weight", n);
        if (!hp) return false;
        const float* s = stage(hp, n);
        if (!s) return false;
        if (want_fwd) {
            c.w = dmalloc<__half>(n);
            if (!c.w) return false;
            pack_conv_kernel<<<2048, 256>>>(s, Co, Ci, 0, c.w, 9L * Ci);
        }
        if (want_dgrad) {
            c.wd = dmalloc<__half>(n);
            if (!c.wd) return false;
            pack_conv_kernel<<<2048, 256>>>(s, Co, Ci, 1, c.wd, 9L * Co);
        }
        if (cudaDeviceSynchronize() != cudaSuccess) {
            err = "pack_conv failed";
            return false;
        }
        c.b = vec(pre + ".bias", Co);
        return c.b != nullptr;
    }
};


}  // namespace s2i
