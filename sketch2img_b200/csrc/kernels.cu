// HBM-bound kernels (see kernels.cuh): 128-bit vectorised loads/stores, warp-shuffle reductions,
// shared-memory staging of per-group statistics.  Layouts: activations [rows = (b,y,x)][C] with a row stride.
#include "kernels.cuh"
#include "common.cuh"

#include <cuda_fp16.h>
#include <math.h>

namespace s2i {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_(float y) { return 1.f / (1.f + __expf(-y)); }

__device__ __forceinline__ uint2 pack_half4(float a, float b, float c, float d) {
    __half2 lo = __floats2half2_rn(a, b);
    __half2 hi = __floats2half2_rn(c, d);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    return r;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

inline int grid_for(long work, int block, int cap = 148 * 16) {
    long g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// ------------------------------------------------------------------------------------------------ GroupNorm
struct GnStat {
    float mean[kGroups];
    float rstd[kGroups];
};

__device__ __forceinline__ void load_gn_stats(GnStat& s, const double* sums, int b, double n, float eps) {
    if (threadIdx.x < kGroups) {
        const double s1 = sums[((long)b * kGroups + threadIdx.x) * 2 + 0];
        const double s2 = sums[((long)b * kGroups + threadIdx.x) * 2 + 1];
        const double m = s1 / n;
        double var = s2 / n - m * m;
        if (var < 0.0) var = 0.0;
        s.mean[threadIdx.x] = (float)m;
        s.rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// MODE 0: accumulate (x, x^2).  MODE 1: accumulate (dxhat, dxhat*xhat) for the backward pass.
template <int MODE>
__global__ void __launch_bounds__(256) gn_reduce_kernel(const float* __restrict__ x, long ldx,
                                                        const float* __restrict__ dy, long ldd, int HW, int C, int P,
                                                        const double* __restrict__ fsums,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, int silu, double* __restrict__ out) {
    __shared__ float s1[kGroups], s2[kGroups];
    __shared__ GnStat st;
    const int b = blockIdx.y;
    const int Cg = C / kGroups;
    const int t = threadIdx.x;
    if (t < kGroups) {
        s1[t] = 0.f;
        s2[t] = 0.f;
    }
    if (MODE == 1) load_gn_stats(st, fsums, b, (double)Cg * HW, eps);
    __syncthreads();
    const int p0 = blockIdx.x * P;
    const int p1 = min(HW, p0 + P);
    const int vec = C >> 2;
    const int qstride = min(vec, (int)blockDim.x);
    const int prow = max(1, (int)blockDim.x / vec);
    const int r = t / qstride;
    const int q0 = t - r * qstride;
    if (r < prow) {
        for (int q = q0; q < vec; q += qstride) {
            const int c = q << 2;
            float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
            float g4[4], b4[4], mu[4], rs[4];
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g4[j] = gamma[c + j];
                    b4[j] = beta[c + j];
                    mu[j] = st.mean[(c + j) / Cg];
                    rs[j] = st.rstd[(c + j) / Cg];
                }
            }
            for (int p = p0 + r; p < p1; p += prow) {
                const long row = (long)b * HW + p;
                const float4 xv = ldg4(x + row * ldx + c);
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        a1[j] += xs[j];
                        a2[j] += xs[j] * xs[j];
                    }
                } else {
                    const float4 dv = ldg4(dy + row * ldd + c);
                    const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float xh = (xs[j] - mu[j]) * rs[j];
                        float d = ds[j];
                        if (silu) {
                            const float y = xh * g4[j] + b4[j];
                            const float sg = sigmoidf_(y);
                            d *= sg * (1.f + y * (1.f - sg));
                        }
                        const float dxh = d * g4[j];
                        a1[j] += dxh;
                        a2[j] += dxh * xh;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&s1[(c + j) / Cg], a1[j]);
                atomicAdd(&s2[(c + j) / Cg], a2[j]);
            }
        }
    }
    __syncthreads();
    if (t < kGroups) {
        atomicAdd(&out[((long)b * kGroups + t) * 2 + 0], (double)s1[t]);
        atomicAdd(&out[((long)b * kGroups + t) * 2 + 1], (double)s2[t]);
    }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, long ldx, int HW, int C,
                                                       const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, __half* __restrict__ out16, long ld16,
                                                       __half* __restrict__ raw16, long ldraw) {
    __shared__ GnStat st;
    const int b = blockIdx.y;
    const int Cg = C / kGroups;
    load_gn_stats(st, sums, b, (double)Cg * HW, eps);
    __syncthreads();
    const int vec = C >> 2;
    const long total = (long)HW * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int p = (int)(idx / vec);
        const int c = (int)(idx - (long)p * vec) << 2;
        const long row = (long)b * HW + p;
        const float4 xv = ldg4(x + row * ldx + c);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = (c + j) / Cg;
            float v = (xs[j] - st.mean[g]) * st.rstd[g] * __ldg(gamma + c + j) + __ldg(beta + c + j);
            if (silu) v = v * sigmoidf_(v);
            y[j] = v;
        }
        *reinterpret_cast<uint2*>(out16 + row * ld16 + c) = pack_half4(y[0], y[1], y[2], y[3]);
        if (raw16) *reinterpret_cast<uint2*>(raw16 + row * ldraw + c) = pack_half4(xs[0], xs[1], xs[2], xs[3]);
    }
}

__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const float* __restrict__ dy, long ldd,
                                                           const float* __restrict__ x, long ldx, int HW, int C,
                                                           const double* __restrict__ sums,
                                                           const double* __restrict__ bsums,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, int silu,
                                                           const float* __restrict__ add, long ldadd,
                                                           float* __restrict__ dx32, long ld32,
                                                           __half* __restrict__ dx16, long ld16) {
    __shared__ GnStat st;
    __shared__ float m1[kGroups], m2[kGroups];
    const int b = blockIdx.y;
    const int Cg = C / kGroups;
    const double n = (double)Cg * HW;
    load_gn_stats(st, sums, b, n, eps);
    if (threadIdx.x < kGroups) {
        m1[threadIdx.x] = (float)(bsums[((long)b * kGroups + threadIdx.x) * 2 + 0] / n);
        m2[threadIdx.x] = (float)(bsums[((long)b * kGroups + threadIdx.x) * 2 + 1] / n);
    }
    __syncthreads();
    const int vec = C >> 2;
    const long total = (long)HW * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int p = (int)(idx / vec);
        const int c = (int)(idx - (long)p * vec) << 2;
        const long row = (long)b * HW + p;
        const float4 xv = ldg4(x + row * ldx + c);
        const float4 dv = ldg4(dy + row * ldd + c);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = (c + j) / Cg;
            const float gm = __ldg(gamma + c + j);
            const float xh = (xs[j] - st.mean[g]) * st.rstd[g];
            float d = ds[j];
            if (silu) {
                const float yv = xh * gm + __ldg(beta + c + j);
                const float sg = sigmoidf_(yv);
                d *= sg * (1.f + yv * (1.f - sg));
            }
            o[j] = st.rstd[g] * (d * gm - m1[g] - xh * m2[g]);
        }
        if (add) {
            const float4 av = ldg4(add + row * ldadd + c);
            o[0] += av.x; o[1] += av.y; o[2] += av.z; o[3] += av.w;
        }
        if (dx32) *reinterpret_cast<float4*>(dx32 + row * ld32 + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (dx16) *reinterpret_cast<uint2*>(dx16 + row * ld16 + c) = pack_half4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long ldx, long rows, int C,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, __half* __restrict__ out16, long ld16,
                                                     float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * ldx;
    const int vec = C >> 2;
    float s = 0.f;
    for (int q = lane; q < vec; q += 32) {
        const float4 v = ldg4(xr + 4 * q);
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
    for (int q = lane; q < vec; q += 32) {
        const float4 v = ldg4(xr + 4 * q);
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        ss += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    if (stats && lane == 0) {
        stats[row * 2 + 0] = mean;
        stats[row * 2 + 1] = rstd;
    }
    for (int q = lane; q < vec; q += 32) {
        const float4 v = ldg4(xr + 4 * q);
        const float4 g = ldg4(gamma + 4 * q);
        const float4 bb = ldg4(beta + 4 * q);
        *reinterpret_cast<uint2*>(out16 + row * ld16 + 4 * q) =
            pack_half4((v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y,
                       (v.z - mean) * rstd * g.z + bb.z, (v.w - mean) * rstd * g.w + bb.w);
    }
}

__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, long ldd,
                                                     const float* __restrict__ x, long ldx, long rows, int C,
                                                     const float* __restrict__ gamma, const float* __restrict__ stats,
                                                     const float* __restrict__ add, long ldadd,
                                                     float* __restrict__ dx32, long ld32, __half* __restrict__ dx16,
                                                     long ld16) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * ldx;
    const float* dr = dy + row * ldd;
    const float mean = stats[row * 2 + 0], rstd = stats[row * 2 + 1];
    const int vec = C >> 2;
    float a1 = 0.f, a2 = 0.f;
    for (int q = lane; q < vec; q += 32) {
        const float4 v = ldg4(xr + 4 * q);
        const float4 d = ldg4(dr + 4 * q);
        const float4 g = ldg4(gamma + 4 * q);
        const float h0 = d.x * g.x, h1 = d.y * g.y, h2 = d.z * g.z, h3 = d.w * g.w;
        a1 += (h0 + h1) + (h2 + h3);
        a2 += h0 * (v.x - mean) * rstd + h1 * (v.y - mean) * rstd + h2 * (v.z - mean) * rstd + h3 * (v.w - mean) * rstd;
    }
    const float m1 = warp_sum(a1) / C, m2 = warp_sum(a2) / C;
    for (int q = lane; q < vec; q += 32) {
        const float4 v = ldg4(xr + 4 * q);
        const float4 d = ldg4(dr + 4 * q);
        const float4 g = ldg4(gamma + 4 * q);
        float o0 = rstd * (d.x * g.x - m1 - (v.x - mean) * rstd * m2);
        float o1 = rstd * (d.y * g.y - m1 - (v.y - mean) * rstd * m2);
        float o2 = rstd * (d.z * g.z - m1 - (v.z - mean) * rstd * m2);
        float o3 = rstd * (d.w * g.w - m1 - (v.w - mean) * rstd * m2);
        if (add) {
            const float4 av = ldg4(add + row * ldadd + 4 * q);
            o0 += av.x; o1 += av.y; o2 += av.z; o3 += av.w;
        }
        if (dx32) *reinterpret_cast<float4*>(dx32 + row * ld32 + 4 * q) = make_float4(o0, o1, o2, o3);
        if (dx16) *reinterpret_cast<uint2*>(dx16 + row * ld16 + 4 * q) = pack_half4(o0, o1, o2, o3);
    }
}

// ------------------------------------------------------------------------------------------------ softmax
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ s, long lds, long rows, int n,
                                                          __half* __restrict__ p16, long ldp) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* sr = s + row * lds;
    __half* pr = p16 + row * ldp;
    const int n4 = ((lds & 3) == 0) ? (n >> 2) : 0;   // vectorisable prefix
    float mx = -INFINITY;
    for (int q = lane; q < n4; q += 32) {
        const float4 v = ldg4(sr + 4 * q);
        mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    for (int j = 4 * n4 + lane; j < n; j += 32) mx = fmaxf(mx, sr[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int q = lane; q < n4; q += 32) {
        const float4 v = ldg4(sr + 4 * q);
        sum += (__expf(v.x - mx) + __expf(v.y - mx)) + (__expf(v.z - mx) + __expf(v.w - mx));
    }
    for (int j = 4 * n4 + lane; j < n; j += 32) sum += __expf(sr[j] - mx);
    const float inv = 1.f / warp_sum(sum);
    const bool vst = (ldp & 3) == 0;
    for (int q = lane; q < n4; q += 32) {
        const float4 v = ldg4(sr + 4 * q);
        const float e0 = __expf(v.x - mx) * inv, e1 = __expf(v.y - mx) * inv, e2 = __expf(v.z - mx) * inv,
                    e3 = __expf(v.w - mx) * inv;
        if (vst) {
            *reinterpret_cast<uint2*>(pr + 4 * q) = pack_half4(e0, e1, e2, e3);
        } else {
            pr[4 * q + 0] = __float2half_rn(e0);
            pr[4 * q + 1] = __float2half_rn(e1);
            pr[4 * q + 2] = __float2half_rn(e2);
            pr[4 * q + 3] = __float2half_rn(e3);
        }
    }
    for (int j = 4 * n4 + lane; j < n; j += 32) pr[j] = __float2half_rn(__expf(sr[j] - mx) * inv);
}

__global__ void __launch_bounds__(256) softmax_bwd_kernel(const __half* __restrict__ p16, long ldp,
                                                          const float* __restrict__ dp, long lddp, long rows, int n,
                                                          float scale, __half* __restrict__ ds16, long ldds) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const __half* pr = p16 + row * ldp;
    const float* dr = dp + row * lddp;
    __half* or_ = ds16 + row * ldds;
    float acc = 0.f;
    for (int j = lane; j < n; j += 32) acc += __half2float(pr[j]) * dr[j];
    const float dot = warp_sum(acc);
    for (int j = lane; j < n; j += 32) or_[j] = __float2half_rn(scale * __half2float(pr[j]) * (dr[j] - dot));
}

// ------------------------------------------------------------------------------------------------ GEGLU
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

__global__ void __launch_bounds__(256) geglu_fwd_kernel(const float* __restrict__ ff, long ldf, long rows, int F,
                                                        __half* __restrict__ out16, long ld16) {
    const int vec = F >> 2;
    const long total = rows * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / vec;
        const int c = (int)(idx - r * vec) << 2;
        const float4 a = ldg4(ff + r * ldf + c);
        const float4 g = ldg4(ff + r * ldf + F + c);
        *reinterpret_cast<uint2*>(out16 + r * ld16 + c) =
            pack_half4(a.x * gelu_exact(g.x), a.y * gelu_exact(g.y), a.z * gelu_exact(g.z), a.w * gelu_exact(g.w));
    }
}

__global__ void __launch_bounds__(256) geglu_bwd_kernel(const float* __restrict__ dg, long ldg,
                                                        const float* __restrict__ ff, long ldf, long rows, int F,
                                                        __half* __restrict__ dff16, long ld16) {
    const int vec = F >> 2;
    const long total = rows * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / vec;
        const int c = (int)(idx - r * vec) << 2;
        const float4 d = ldg4(dg + r * ldg + c);
        const float4 a = ldg4(ff + r * ldf + c);
        const float4 g = ldg4(ff + r * ldf + F + c);
        *reinterpret_cast<uint2*>(dff16 + r * ld16 + c) =
            pack_half4(d.x * gelu_exact(g.x), d.y * gelu_exact(g.y), d.z * gelu_exact(g.z), d.w * gelu_exact(g.w));
        *reinterpret_cast<uint2*>(dff16 + r * ld16 + F + c) =
            pack_half4(d.x * a.x * gelu_grad(g.x), d.y * a.y * gelu_grad(g.y), d.z * a.z * gelu_grad(g.z),
                       d.w * a.w * gelu_grad(g.w));
    }
}

// ------------------------------------------------------------------------------------------------ movement
__global__ void __launch_bounds__(256) add2d_kernel(const float* __restrict__ a, long lda, const float* __restrict__ b,
                                                    long ldb, long rows, int cols, float* __restrict__ d32, long ld32,
                                                    __half* __restrict__ d16, long ld16) {
    const int vec = cols >> 2;
    const long total = rows * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / vec;
        const int c = (int)(idx - r * vec) << 2;
        float4 v = ldg4(a + r * lda + c);
        if (b) {
            const float4 w = ldg4(b + r * ldb + c);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        if (d32) *reinterpret_cast<float4*>(d32 + r * ld32 + c) = v;
        if (d16) *reinterpret_cast<uint2*>(d16 + r * ld16 + c) = pack_half4(v.x, v.y, v.z, v.w);
    }
}

__global__ void __launch_bounds__(256) cast2d_kernel(const float* __restrict__ a, long lda, long rows, int cols,
                                                     float mul, __half* __restrict__ d16, long ld16) {
    const int vec = cols >> 2;
    const long total = rows * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / vec;
        const int c = (int)(idx - r * vec) << 2;
        const float4 v = ldg4(a + r * lda + c);
        *reinterpret_cast<uint2*>(d16 + r * ld16 + c) = pack_half4(v.x * mul, v.y * mul, v.z * mul, v.w * mul);
    }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ x, long ldx, int B, int H, int W,
                                                         int C, __half* __restrict__ out16, long ld16) {
    const int vec = C >> 2;
    const long total = (long)B * 2 * H * 2 * W * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long pix = idx / vec;
        const int c = (int)(idx - pix * vec) << 2;
        const int ox = (int)(pix % (2 * W));
        const int oy = (int)((pix / (2 * W)) % (2 * H));
        const int b = (int)(pix / ((long)4 * W * H));
        const long src = ((long)b * H + (oy >> 1)) * W + (ox >> 1);
        const float4 v = ldg4(x + src * ldx + c);
        *reinterpret_cast<uint2*>(out16 + pix * ld16 + c) = pack_half4(v.x, v.y, v.z, v.w);
    }
}

__global__ void __launch_bounds__(256) sumpool2x_kernel(const float* __restrict__ d, long ldd, int B, int H, int W,
                                                        int C, float* __restrict__ dx, long ldx) {
    const int vec = C >> 2;
    const long total = (long)B * H * W * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long pix = idx / vec;
        const int c = (int)(idx - pix * vec) << 2;
        const int xx = (int)(pix % W);
        const int yy = (int)((pix / W) % H);
        const int b = (int)(pix / ((long)W * H));
        const long base = ((long)b * 2 * H + 2 * yy) * 2 * W + 2 * xx;
        const float4 v0 = ldg4(d + base * ldd + c);
        const float4 v1 = ldg4(d + (base + 1) * ldd + c);
        const float4 v2 = ldg4(d + (base + 2 * W) * ldd + c);
        const float4 v3 = ldg4(d + (base + 2 * W + 1) * ldd + c);
        *reinterpret_cast<float4*>(dx + pix * ldx + c) =
            make_float4((v0.x + v1.x) + (v2.x + v3.x), (v0.y + v1.y) + (v2.y + v3.y), (v0.z + v1.z) + (v2.z + v3.z),
                        (v0.w + v1.w) + (v2.w + v3.w));
    }
}

__global__ void __launch_bounds__(256) zero_insert2x_kernel(const float* __restrict__ d, long ldd, int B, int Ho,
                                                            int Wo, int C, __half* __restrict__ out16, long ld16) {
    const int vec = C >> 2;
    const long total = (long)B * 2 * Ho * 2 * Wo * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long pix = idx / vec;
        const int c = (int)(idx - pix * vec) << 2;
        const int ox = (int)(pix % (2 * Wo));
        const int oy = (int)((pix / (2 * Wo)) % (2 * Ho));
        const int b = (int)(pix / ((long)4 * Wo * Ho));
        uint2 o = make_uint2(0u, 0u);
        if (((ox | oy) & 1) == 0) {
            const long src = ((long)b * Ho + (oy >> 1)) * Wo + (ox >> 1);
            const float4 v = ldg4(d + src * ldd + c);
            o = pack_half4(v.x, v.y, v.z, v.w);
        }
        *reinterpret_cast<uint2*>(out16 + pix * ld16 + c) = o;
    }
}

__global__ void __launch_bounds__(256) im2col3x3_kernel(const float* __restrict__ x, long ldx, int B, int H, int W,
                                                        int C, int stride, int Ho, int Wo, __half* __restrict__ col,
                                                        long ldcol) {
    const long total = (long)B * Ho * Wo * ldcol;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long pix = idx / ldcol;
        const int k = (int)(idx - pix * ldcol);
        float v = 0.f;
        if (k < 9 * C) {
            const int tap = k / C, c = k - tap * C;
            const int ox = (int)(pix % Wo);
            const int oy = (int)((pix / Wo) % Ho);
            const int b = (int)(pix / ((long)Wo * Ho));
            const int iy = oy * stride + tap / 3 - 1;
            const int ix = ox * stride + tap % 3 - 1;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long)b * H + iy) * W + ix) * ldx + c];
        }
        col[idx] = __float2half_rn(v);
    }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int B, int C, int H, int W, float* __restrict__ dst,
                                    long ldn) {
    const long total = (long)B * H * W * ldn;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long pix = idx / ldn;
        const int c = (int)(idx - pix * ldn);
        const long hw = (long)H * W;
        const long b = pix / hw, p = pix - b * hw;
        dst[idx] = c < C ? src[(b * C + c) * hw + p] : 0.f;
    }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, long ldn, int B, int C, int H, int W,
                                    float* __restrict__ dst) {
    const long hw = (long)H * W;
    const long total = (long)B * C * hw;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long p = idx % hw;
        const int c = (int)((idx / hw) % C);
        const long b = idx / (hw * C);
        dst[idx] = src[(b * hw + p) * ldn + c];
    }
}

// ------------------------------------------------------------------------------------------------ time embedding
__global__ void __launch_bounds__(256) gemv_kernel(const float* __restrict__ x, int K, const __half* __restrict__ w,
                                                   const float* __restrict__ bias, int N, int silu_in,
                                                   float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    const __half* wr = w + (long)n * K;
    float acc = 0.f;
    for (int k = 2 * lane; k < K; k += 64) {
        const __half2 w2 = *reinterpret_cast<const __half2*>(wr + k);
        float x0 = x[k], x1 = x[k + 1];
        if (silu_in) {
            x0 = x0 * sigmoidf_(x0);
            x1 = x1 * sigmoidf_(x1);
        }
        acc += __low2float(w2) * x0 + __high2float(w2) * x1;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[n] = acc + (bias ? bias[n] : 0.f);
}

__global__ void timestep_embedding_kernel(float t, int dim, float* __restrict__ out) {
    const int half = dim / 2;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float freq = expf(-logf(10000.f) * (float)i / (float)half);
        const float arg = t * freq;
        out[i] = cosf(arg);          // flip_sin_to_cos: [cos | sin]
        out[half + i] = sinf(arg);
    }
}

}  // namespace

// ================================================================================================== launchers
#define S2I_REQ(cond, msg)                                       \
    do {                                                         \
        if (!(cond)) return set_error(S2I_ERR_ARG, "%s", msg);   \
    } while (0)

static int gn_chunk(int B, int HW) {
    int P = (int)(((long)HW * B) / 512);
    if (P < 4) P = 4;
    if (P > 128) P = 128;
    return P;
}

int gn_stats(const float* x, long ldx, int B, int HW, int C, double* sums, cudaStream_t st) {
    S2I_REQ(C % (4 * 1) == 0 && C % kGroups == 0 && (ldx & 3) == 0, "gn_stats: C must be a multiple of 32 and ld of 4");
    const int P = gn_chunk(B, HW);
    dim3 grid(ceil_div(HW, P), B);
    gn_reduce_kernel<0><<<grid, 256, 0, st>>>(x, ldx, nullptr, 0, HW, C, P, nullptr, nullptr, nullptr, 0.f, 0, sums);
    S2I_LAUNCH_CHECK();
    return 0;
}

int gn_apply(const float* x, long ldx, int B, int HW, int C, const double* sums, const float* gamma, const float* beta,
             float eps, int silu, void* out16, long ld16, void* raw16, long ldraw, cudaStream_t st) {
    S2I_REQ(C % kGroups == 0 && (ldx & 3) == 0 && (ld16 & 3) == 0 && (ldraw & 3) == 0, "gn_apply: alignment");
    dim3 grid(grid_for((long)HW * (C >> 2), 256, 148 * 8), B);
    gn_apply_kernel<<<grid, 256, 0, st>>>(x, ldx, HW, C, sums, gamma, beta, eps, silu, (__half*)out16, ld16,
                                         (__half*)raw16, ldraw);
    S2I_LAUNCH_CHECK();
    return 0;
}

int gn_bwd_stats(const float* dy, long ldd, const float* x, long ldx, int B, int HW, int C, const double* sums,
                 const float* gamma, const float* beta, float eps, int silu, double* bsums, cudaStream_t st) {
    S2I_REQ(C % kGroups == 0 && (ldx & 3) == 0 && (ldd & 3) == 0, "gn_bwd_stats: alignment");
    const int P = gn_chunk(B, HW);
    dim3 grid(ceil_div(HW, P), B);
    gn_reduce_kernel<1><<<grid, 256, 0, st>>>(x, ldx, dy, ldd, HW, C, P, sums, gamma, beta, eps, silu, bsums);
    S2I_LAUNCH_CHECK();
    return 0;
}

int gn_bwd_apply(const float* dy, long ldd, const float* x, long ldx, int B, int HW, int C, const double* sums,
                 const double* bsums, const float* gamma, const float* beta, float eps, int silu, const float* add,
                 long ldadd, float* dx32, long ld32, void* dx16, long ld16, cudaStream_t st) {
    S2I_REQ(C % kGroups == 0 && (ldx & 3) == 0 && (ldd & 3) == 0 && (ldadd & 3) == 0 && (ld32 & 3) == 0 && (ld16 & 3) == 0,
            "gn_bwd_apply: alignment");
    dim3 grid(grid_for((long)HW * (C >> 2), 256, 148 * 8), B);
    gn_bwd_apply_kernel<<<grid, 256, 0, st>>>(dy, ldd, x, ldx, HW, C, sums, bsums, gamma, beta, eps, silu, add, ldadd,
                                             dx32, ld32, (__half*)dx16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int ln_fwd(const float* x, long ldx, long rows, int C, const float* gamma, const float* beta, float eps, void* out16,
           long ld16, float* stats, cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldx & 3) == 0 && (ld16 & 3) == 0, "ln_fwd: alignment");
    ln_fwd_kernel<<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(x, ldx, rows, C, gamma, beta, eps, (__half*)out16, ld16,
                                                                stats);
    S2I_LAUNCH_CHECK();
    return 0;
}

int ln_bwd(const float* dy, long ldd, const float* x, long ldx, long rows, int C, const float* gamma,
           const float* stats, const float* add, long ldadd, float* dx32, long ld32, void* dx16, long ld16,
           cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldx & 3) == 0 && (ldd & 3) == 0 && (ldadd & 3) == 0 && (ld32 & 3) == 0 && (ld16 & 3) == 0,
            "ln_bwd: alignment");
    ln_bwd_kernel<<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(dy, ldd, x, ldx, rows, C, gamma, stats, add, ldadd, dx32,
                                                                ld32, (__half*)dx16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int softmax_fwd(const float* s, long lds, long rows, int n, void* p16, long ldp, cudaStream_t st) {
    softmax_fwd_kernel<<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(s, lds, rows, n, (__half*)p16, ldp);
    S2I_LAUNCH_CHECK();
    return 0;
}

int softmax_bwd(const void* p16, long ldp, const float* dp, long lddp, long rows, int n, float scale, void* ds16,
                long ldds, cudaStream_t st) {
    softmax_bwd_kernel<<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>((const __half*)p16, ldp, dp, lddp, rows, n, scale,
                                                                     (__half*)ds16, ldds);
    S2I_LAUNCH_CHECK();
    return 0;
}

int geglu_fwd(const float* ff, long ldf, long rows, int F, void* out16, long ld16, cudaStream_t st) {
    S2I_REQ((F & 3) == 0 && (ldf & 3) == 0 && (ld16 & 3) == 0, "geglu_fwd: alignment");
    geglu_fwd_kernel<<<grid_for(rows * (F >> 2), 256), 256, 0, st>>>(ff, ldf, rows, F, (__half*)out16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int geglu_bwd(const float* dg, long ldg, const float* ff, long ldf, long rows, int F, void* dff16, long ld16,
              cudaStream_t st) {
    S2I_REQ((F & 3) == 0 && (ldf & 3) == 0 && (ldg & 3) == 0 && (ld16 & 3) == 0, "geglu_bwd: alignment");
    geglu_bwd_kernel<<<grid_for(rows * (F >> 2), 256), 256, 0, st>>>(dg, ldg, ff, ldf, rows, F, (__half*)dff16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int add2d(const float* a, long lda, const float* b, long ldb, long rows, int cols, float* dst32, long ld32, void* dst16,
          long ld16, cudaStream_t st) {
    S2I_REQ((cols & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0 && (ld32 & 3) == 0 && (ld16 & 3) == 0, "add2d: alignment");
    add2d_kernel<<<grid_for(rows * (cols >> 2), 256), 256, 0, st>>>(a, lda, b, ldb, rows, cols, dst32, ld32,
                                                                   (__half*)dst16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int cast2d(const float* a, long lda, long rows, int cols, float mul, void* dst16, long ld16, cudaStream_t st) {
    S2I_REQ((cols & 3) == 0 && (lda & 3) == 0 && (ld16 & 3) == 0, "cast2d: alignment");
    cast2d_kernel<<<grid_for(rows * (cols >> 2), 256), 256, 0, st>>>(a, lda, rows, cols, mul, (__half*)dst16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int upsample2x(const float* x, long ldx, int B, int H, int W, int C, void* out16, long ld16, cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldx & 3) == 0 && (ld16 & 3) == 0, "upsample2x: alignment");
    upsample2x_kernel<<<grid_for((long)B * 4 * H * W * (C >> 2), 256), 256, 0, st>>>(x, ldx, B, H, W, C, (__half*)out16,
                                                                                    ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int sumpool2x(const float* d, long ldd, int B, int H, int W, int C, float* dx, long ldx, cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldx & 3) == 0 && (ldd & 3) == 0, "sumpool2x: alignment");
    sumpool2x_kernel<<<grid_for((long)B * H * W * (C >> 2), 256), 256, 0, st>>>(d, ldd, B, H, W, C, dx, ldx);
    S2I_LAUNCH_CHECK();
    return 0;
}

int zero_insert2x(const float* d, long ldd, int B, int Ho, int Wo, int C, void* out16, long ld16, cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldd & 3) == 0 && (ld16 & 3) == 0, "zero_insert2x: alignment");
    zero_insert2x_kernel<<<grid_for((long)B * 4 * Ho * Wo * (C >> 2), 256), 256, 0, st>>>(d, ldd, B, Ho, Wo, C,
                                                                                         (__half*)out16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int im2col3x3(const float* x, long ldx, int B, int H, int W, int C, int stride, void* col16, long ldcol,
              cudaStream_t st) {
    S2I_REQ(stride == 1 || stride == 2, "im2col3x3: stride");
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    S2I_REQ(ldcol >= 9L * C, "im2col3x3: ldcol too small");
    im2col3x3_kernel<<<grid_for((long)B * Ho * Wo * ldcol, 256), 256, 0, st>>>(x, ldx, B, H, W, C, stride, Ho, Wo,
                                                                              (__half*)col16, ldcol);
    S2I_LAUNCH_CHECK();
    return 0;
}

int nchw_to_nhwc(const float* src, int B, int C, int H, int W, float* dst, long ldn, cudaStream_t st) {
    nchw_to_nhwc_kernel<<<grid_for((long)B * H * W * ldn, 256), 256, 0, st>>>(src, B, C, H, W, dst, ldn);
    S2I_LAUNCH_CHECK();
    return 0;
}

int nhwc_to_nchw(const float* src, long ldn, int B, int C, int H, int W, float* dst, cudaStream_t st) {
    nhwc_to_nchw_kernel<<<grid_for((long)B * C * H * W, 256), 256, 0, st>>>(src, ldn, B, C, H, W, dst);
    S2I_LAUNCH_CHECK();
    return 0;
}

int gemv(const float* x, int K, const void* w16, const float* bias, int N, int silu_in, float* out, cudaStream_t st) {
    S2I_REQ((K & 1) == 0, "gemv: K must be even");
    gemv_kernel<<<ceil_div(N, 8), 256, 0, st>>>(x, K, (const __half*)w16, bias, N, silu_in, out);
    S2I_LAUNCH_CHECK();
    return 0;
}

int timestep_embedding(float t, int dim, float* out, cudaStream_t st) {
    timestep_embedding_kernel<<<1, 256, 0, st>>>(t, dim, out);
    S2I_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2i
