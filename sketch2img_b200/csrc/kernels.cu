// HBM-bound kernels (see kernels.cuh): 128-bit vectorised loads/stores, warp-shuffle reductions,
// shared-memory staging of per-group statistics.  Layouts: activations [rows = (b,y,x)][C] with a row stride.
#include "kernels.cuh"
#include "common.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <cstdlib>
#include <cstring>
#include <map>

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
// Statistics slot of one GroupNorm call: per sample 512 bytes = float2[32] (mean, rstd) -- or, for the backward
// reduction, (mean dxhat, mean dxhat*xhat) -- followed by the arrival counter of the reduction (zero between calls).
struct GnStat {
    float mean[kGroups];
    float rstd[kGroups];
};
constexpr int kSlotDoubles = kGroups * 2;   // doubles per sample in the caller's slot (512 bytes)

__device__ __forceinline__ void load_gn_stats(GnStat& s, const double* slot, int b) {
    if (threadIdx.x < kGroups) {
        const float2 v = reinterpret_cast<const float2*>(slot + (long)b * kSlotDoubles)[threadIdx.x];
        s.mean[threadIdx.x] = v.x;
        s.rstd[threadIdx.x] = v.y;
    }
}

// MODE 0: (sum x, sum x^2) -> (mean, rstd).  MODE 1: (sum dxhat, sum dxhat*xhat) -> their means (backward pass).
// Each block reduces a slab of pixels of one sample (register partials, 8 loads in flight per thread), writes its
// 32 x 2 group sums to `partial`; the last block of a sample to arrive adds the partials in block order (the result
// does not depend on scheduling) and writes the finished statistics.
template <int MODE>
__global__ void __launch_bounds__(256) gn_reduce_kernel(const float* __restrict__ x, long ldx,
                                                        const float* __restrict__ dy, long ldd, int HW, int C, int P,
                                                        const double* __restrict__ fslot,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, int silu, double* __restrict__ partial,
                                                        double* __restrict__ slot) {
    pdl_wait();
    pdl_launch();
    __shared__ float s1[kGroups], s2[kGroups];
    __shared__ GnStat st;
    __shared__ int is_last;
    const int b = blockIdx.y;
    const int nblk = gridDim.x;
    const int Cg = C / kGroups;
    const int t = threadIdx.x;
    if (t < kGroups) {
        s1[t] = 0.f;
        s2[t] = 0.f;
    }
    if (MODE == 1) load_gn_stats(st, fslot, b);
    __syncthreads();
    const int p0 = blockIdx.x * P;
    const int p1 = min(HW, p0 + P);
    const int vec = C >> 2;
    const int qstride = min(vec, (int)blockDim.x);
    const int prow = max(1, (int)blockDim.x / vec);
    const int r = t / qstride;
    const int q0 = t - r * qstride;
    constexpr int U = 8;
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
            for (int pb = p0 + r; pb < p1; pb += U * prow) {
                float4 xv[U], dv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pp = pb + u * prow;
                    const long row = (long)b * HW + (pp < p1 ? pp : p0);
                    xv[u] = ldg4(x + row * ldx + c);
                    if (MODE == 1) dv[u] = ldg4(dy + row * ldd + c);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (pb + u * prow >= p1) break;
                    const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
                    if (MODE == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            a1[j] += xs[j];
                            a2[j] += xs[j] * xs[j];
                        }
                    } else {
                        const float ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
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
            }
            if (Cg < 2) {   // degenerate: one channel per group
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(&s1[c + j], a1[j]);
                    atomicAdd(&s2[c + j], a2[j]);
                }
                continue;
            }
            // channels c..c+3 fall in at most two groups: merge before touching shared memory
            const int ga = c / Cg, gb = (c + 3) / Cg;
            float sa1 = 0.f, sa2 = 0.f, sb1 = 0.f, sb2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if ((c + j) / Cg == ga) {
                    sa1 += a1[j];
                    sa2 += a2[j];
                } else {
                    sb1 += a1[j];
                    sb2 += a2[j];
                }
            }
            atomicAdd(&s1[ga], sa1);
            atomicAdd(&s2[ga], sa2);
            if (gb != ga) {
                atomicAdd(&s1[gb], sb1);
                atomicAdd(&s2[gb], sb2);
            }
        }
    }
    __syncthreads();
    double* mine = partial + ((long)b * nblk + blockIdx.x) * (2 * kGroups);
    if (t < kGroups) {
        mine[2 * t] = (double)s1[t];
        mine[2 * t + 1] = (double)s2[t];
    }
    __threadfence();
    __syncthreads();
    unsigned int* counter = reinterpret_cast<unsigned int*>(slot + (long)b * kSlotDoubles + kGroups);
    if (t == 0) {
        const unsigned int old = atomicAdd(counter, 1u);
        is_last = (old == (unsigned int)(nblk - 1));
        if (is_last) *counter = 0u;     // ready for the next launch that reuses this slot
    }
    __syncthreads();
    if (is_last && t < kGroups) {
        __threadfence();
        const double* base = partial + (long)b * nblk * (2 * kGroups);
        double t1 = 0.0, t2 = 0.0;
        for (int k = 0; k < nblk; ++k) {
            t1 += __ldcg(base + (long)k * (2 * kGroups) + 2 * t);
            t2 += __ldcg(base + (long)k * (2 * kGroups) + 2 * t + 1);
        }
        const double n = (double)Cg * HW;
        float2 o;
        if (MODE == 0) {
            const double m = t1 / n;
            double var = t2 / n - m * m;
            if (var < 0.0) var = 0.0;
            o.x = (float)m;
            o.y = (float)(1.0 / sqrt(var + (double)eps));
        } else {
            o.x = (float)(t1 / n);
            o.y = (float)(t2 / n);
        }
        reinterpret_cast<float2*>(slot + (long)b * kSlotDoubles)[t] = o;
    }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, long ldx, int HW, int C,
                                                       const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, __half* __restrict__ out16, long ld16,
                                                       __half* __restrict__ raw16, long ldraw) {
    pdl_wait();
    pdl_launch();
    __shared__ GnStat st;
    const int b = blockIdx.y;
    const int Cg = C / kGroups;
    load_gn_stats(st, sums, b);
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
    pdl_wait();
    pdl_launch();
    __shared__ GnStat st;
    __shared__ float m1[kGroups], m2[kGroups];
    const int b = blockIdx.y;
    const int Cg = C / kGroups;
    load_gn_stats(st, sums, b);
    if (threadIdx.x < kGroups) {
        const float2 v = reinterpret_cast<const float2*>(bsums + (long)b * kSlotDoubles)[threadIdx.x];
        m1[threadIdx.x] = v.x;
        m2[threadIdx.x] = v.y;
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

// ---- fused statistics + apply -------------------------------------------------------------------------------------
// One launch per GroupNorm (forward: stats -> normalise/affine/SiLU -> fp16; backward: the two reduction terms -> dx).
// grid = (nblk, B) with nblk * B <= #SMs, so every block is resident and a grid-wide arrival counter (in the call's
// statistics slot, zeroed once per pass by the caller) can separate the reduction from the apply phase.  Each block
// reduces its slab of pixels, publishes 32 x 2 partial sums, waits for the grid, adds its sample's partials in block
// order (deterministic) and applies to the same slab (second read comes from L2).
struct GnFusedArgs {
    const float* x; long ldx;
    const float* dy; long ldd;               // backward only
    int HW, C, P;
    const double* fslot;                     // backward: the forward statistics slot
    const float* gamma; const float* beta;
    float eps; int silu;
    double* partial; double* slot;
    __half* out16; long ld16; __half* raw16; long ldraw;            // forward outputs
    const float* add; long ldadd; float* dx32; long ld32;           // backward outputs (dx16 = out16 / ld16)
};

template <int MODE>
__global__ void __launch_bounds__(512) gn_fused_kernel(const GnFusedArgs a) {
    pdl_wait();
    pdl_launch();
    __shared__ float s1[kGroups], s2[kGroups];
    __shared__ GnStat st;            // forward statistics (MODE 0: computed here; MODE 1: loaded)
    __shared__ float m1[kGroups], m2[kGroups];
    __shared__ double red[8][2 * kGroups];
    __shared__ __align__(16) float cs[5120];     // [2][prow][C] channel partials, prow * C <= max(2048, C) <= 2560
    const int b = blockIdx.y;
    const int nblk = gridDim.x;
    const int C = a.C, HW = a.HW;
    const int Cg = C / kGroups;
    const int t = threadIdx.x;
    if (t < kGroups) {
        s1[t] = 0.f;
        s2[t] = 0.f;
    }
    if (MODE == 1) load_gn_stats(st, a.fslot, b);
    __syncthreads();
    const int p0 = blockIdx.x * a.P;
    const int p1 = min(HW, p0 + a.P);
    const int vec = C >> 2;
    {   // ---------------- phase 1: slab reduction
        const int qstride = min(vec, (int)blockDim.x);
        const int prow = max(1, (int)blockDim.x / vec);
        const int r = t / qstride;
        const int q0 = t - r * qstride;
        constexpr int U = 8;
        if (r < prow) {
            for (int q = q0; q < vec; q += qstride) {
                const int c = q << 2;
                float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
                float g4[4], b4[4], mu[4], rs[4];
                if (MODE == 1) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        g4[j] = a.gamma[c + j];
                        b4[j] = a.beta[c + j];
                        mu[j] = st.mean[(c + j) / Cg];
                        rs[j] = st.rstd[(c + j) / Cg];
                    }
                }
                for (int pb = p0 + r; pb < p1; pb += U * prow) {
                    float4 xv[U], dv[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int pp = pb + u * prow;
                        const long row = (long)b * HW + (pp < p1 ? pp : p0);
                        xv[u] = ldg4(a.x + row * a.ldx + c);
                        if (MODE == 1) dv[u] = ldg4(a.dy + row * a.ldd + c);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (pb + u * prow >= p1) break;
                        const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
                        if (MODE == 0) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                a1[j] += xs[j];
                                a2[j] += xs[j] * xs[j];
                            }
                        } else {
                            const float ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float xh = (xs[j] - mu[j]) * rs[j];
                                float d = ds[j];
                                if (a.silu) {
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
                }
                // per-(pixel-row, channel) partials -> shared memory (no atomics); reduced per group below
                *reinterpret_cast<float4*>(cs + ((long)r * C + c)) = make_float4(a1[0], a1[1], a1[2], a1[3]);
                *reinterpret_cast<float4*>(cs + ((long)(prow + r) * C + c)) = make_float4(a2[0], a2[1], a2[2], a2[3]);
            }
        }
        __syncthreads();
        if (t < 2 * kGroups) {
            const int g = t >> 1, which = t & 1;
            float acc = 0.f;
            for (int rr = 0; rr < prow; ++rr) {
                const float* src = cs + ((long)(which * prow + rr) * C + g * Cg);
                for (int j = 0; j < Cg; ++j) acc += src[j];
            }
            if (which) s2[g] = acc; else s1[g] = acc;
        }
    }
    __syncthreads();
    double* mine = a.partial + ((long)b * nblk + blockIdx.x) * (2 * kGroups);
    if (t < kGroups) {
        mine[2 * t] = (double)s1[t];
        mine[2 * t + 1] = (double)s2[t];
    }
    __syncthreads();
    // ---------------- grid-wide arrival (all blocks are resident: grid <= #SMs)
    if (t == 0) {
        unsigned int* counter = reinterpret_cast<unsigned int*>(a.slot + kGroups);
        const unsigned int total = gridDim.x * gridDim.y;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < total);
    }
    __syncthreads();
    {   // ---------------- this sample's totals: 8 thread groups x 64 values, block order
        const int v = t & 63, k0 = t >> 6;
        const double* base = a.partial + (long)b * nblk * (2 * kGroups) + v;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        int k = k0;
        for (; k + 24 < nblk; k += 32) {      // four independent loads in flight
            const double v0 = __ldcg(base + (long)k * (2 * kGroups));
            const double v1 = __ldcg(base + (long)(k + 8) * (2 * kGroups));
            const double v2 = __ldcg(base + (long)(k + 16) * (2 * kGroups));
            const double v3 = __ldcg(base + (long)(k + 24) * (2 * kGroups));
            acc0 += v0; acc1 += v1; acc2 += v2; acc3 += v3;
        }
        for (; k < nblk; k += 8) acc0 += __ldcg(base + (long)k * (2 * kGroups));
        red[k0][v] = (acc0 + acc1) + (acc2 + acc3);
    }
    __syncthreads();
    if (t < kGroups) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            t1 += red[k][2 * t];
            t2 += red[k][2 * t + 1];
        }
        const double n = (double)Cg * HW;
        float2 o;
        if (MODE == 0) {
            const double m = t1 / n;
            double var = t2 / n - m * m;
            if (var < 0.0) var = 0.0;
            o.x = (float)m;
            o.y = (float)(1.0 / sqrt(var + (double)a.eps));
            st.mean[t] = o.x;
            st.rstd[t] = o.y;
        } else {
            o.x = (float)(t1 / n);
            o.y = (float)(t2 / n);
            m1[t] = o.x;
            m2[t] = o.y;
        }
        if (blockIdx.x == 0) {
            // slot layout: float2[32] per sample; sample 0's upper half holds the arrival counter (left untouched)
            reinterpret_cast<float2*>(a.slot + (long)b * kSlotDoubles)[t] = o;
        }
    }
    __syncthreads();
    // ---------------- phase 2: apply to the same slab (4 independent float4 per thread in flight)
    const long total = (long)(p1 - p0) * vec;
    constexpr int U2 = 4;
    for (long i0 = t; i0 < total; i0 += (long)U2 * blockDim.x) {
        float4 xv[U2], dv[U2], av[U2];
        long rowu[U2];
        int cu[U2];
#pragma unroll
        for (int u = 0; u < U2; ++u) {
            const long idx = i0 + (long)u * blockDim.x;
            const long id2 = idx < total ? idx : (long)t;
            const int p = p0 + (int)(id2 / vec);
            cu[u] = (int)(id2 % vec) << 2;
            rowu[u] = (long)b * HW + p;
            xv[u] = ldg4(a.x + rowu[u] * a.ldx + cu[u]);
            if (MODE == 1) {
                dv[u] = ldg4(a.dy + rowu[u] * a.ldd + cu[u]);
                av[u] = a.add ? ldg4(a.add + rowu[u] * a.ldadd + cu[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U2; ++u) {
            if (i0 + (long)u * blockDim.x >= total) break;
            const long row = rowu[u];
            const int c = cu[u];
            const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
            if (MODE == 0) {
                float y[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int g = (c + j) / Cg;
                    float v = (xs[j] - st.mean[g]) * st.rstd[g] * __ldg(a.gamma + c + j) + __ldg(a.beta + c + j);
                    if (a.silu) v = v * sigmoidf_(v);
                    y[j] = v;
                }
                *reinterpret_cast<uint2*>(a.out16 + row * a.ld16 + c) = pack_half4(y[0], y[1], y[2], y[3]);
                if (a.raw16) *reinterpret_cast<uint2*>(a.raw16 + row * a.ldraw + c) = pack_half4(xs[0], xs[1], xs[2], xs[3]);
            } else {
                const float ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int g = (c + j) / Cg;
                    const float gm = __ldg(a.gamma + c + j);
                    const float xh = (xs[j] - st.mean[g]) * st.rstd[g];
                    float d = ds[j];
                    if (a.silu) {
                        const float yv = xh * gm + __ldg(a.beta + c + j);
                        const float sg = sigmoidf_(yv);
                        d *= sg * (1.f + yv * (1.f - sg));
                    }
                    o[j] = st.rstd[g] * (d * gm - m1[g] - xh * m2[g]);
                }
                o[0] += av[u].x; o[1] += av[u].y; o[2] += av[u].z; o[3] += av[u].w;
                if (a.dx32) *reinterpret_cast<float4*>(a.dx32 + row * a.ld32 + c) = make_float4(o[0], o[1], o[2], o[3]);
                if (a.out16) *reinterpret_cast<uint2*>(a.out16 + row * a.ld16 + c) = pack_half4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// ---- cluster form: statistics + apply in one launch, no grid-wide barrier --------------------------------------------
// One thread-block CLUSTER owns `gpc` consecutive groups of one sample (W = gpc * C/32 channels, a column slab of the
// NHWC tensor); its `cs` CTAs split the pixels.  Each CTA reduces its [P pixels][W channels] slab from registers
// (float4 loads, 8 in flight per thread) while staging it in shared memory, the per-group sums are exchanged through
// distributed shared memory in rank order (deterministic), and the apply phase reads the slab back from shared memory:
// x (and dy) cross the L2 -> SM path once.  When the slab does not fit (cached = 0) the apply phase re-reads global.
struct GnClArgs {
    const float* x; long ldx;
    const float* dy; long ldd;               // backward only
    int HW, C, Cg, gpc, W, P, cs, cached;
    const double* fslot;                     // backward: the forward statistics slot
    const float* gamma; const float* beta;
    float eps; int silu;
    double* slot;
    __half* out16; long ld16; __half* raw16; long ldraw;            // forward outputs
    const float* add; long ldadd; float* dx32; long ld32;           // backward outputs (dx16 = out16 / ld16)
};
constexpr int kGnClT = 512;          // threads per CTA
constexpr int kGnClColp = 4096;      // floats: [2][prow][W] per-(pixel-lane, channel) partials, prow * W <= 2048
constexpr int kGnClRed = 2048;       // floats: [SUB][2 W] second-stage partials (W <= 1024)
constexpr int kGnClMaxW = 1024;

__device__ __forceinline__ void cluster_arrive_() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ double ld_cluster_f64(const double* local, uint32_t rank) {
    const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
    uint32_t ra;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(kGnClT) gn_cluster_kernel(const GnClArgs a) {
    extern __shared__ __align__(16) float gsm[];
    float* colp = gsm;
    float* red2 = gsm + kGnClColp;
    float* slab = gsm + kGnClColp + kGnClRed;     // MODE 0: x [P][W];  MODE 1: xhat [P][W] then dxhat [P][W]
    __shared__ double part[2 * kGroups];
    __shared__ float s_mean[kGroups], s_rstd[kGroups], s_m1[kGroups], s_m2[kGroups];
    // index arithmetic and the affine parameters (weights) first: they overlap the tail of the preceding kernel
    const int t = threadIdx.x, b = blockIdx.y;
    const int cs = a.cs;
    const int rank = blockIdx.x % cs;             // == %cluster_ctarank (clusters are 1-D along x)
    const int chunk = blockIdx.x / cs;
    const int W = a.W, Wq = W >> 2, Cg = a.Cg, gpc = a.gpc;
    const int c0 = chunk * W, g0 = chunk * gpc;
    const int p0 = rank * a.P, p1 = min(a.HW, p0 + a.P);
    const int prow = kGnClT / Wq;
    const int r = t / Wq, q = t - r * Wq;
    const bool active = r < prow;
    const int c = c0 + 4 * q;
    const long slabN = (long)a.P * W;
    int gl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) gl[j] = (4 * q + j) / Cg;
    float gm[4], bt[4], mu[4], rs[4];
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            gm[j] = __ldg(a.gamma + c + j);
            bt[j] = __ldg(a.beta + c + j);
        }
    }
    pdl_wait();
    pdl_launch();
    if (MODE == 1) {
        if (t < gpc) {
            const float2 v = reinterpret_cast<const float2*>(a.fslot + (long)b * kSlotDoubles)[g0 + t];
            s_mean[t] = v.x;
            s_rstd[t] = v.y;
        }
        __syncthreads();
    }
    if (active && MODE == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mu[j] = s_mean[gl[j]];
            rs[j] = s_rstd[gl[j]];
        }
    }
    // ---------------- phase 1: slab reduction (and staging)
    {
        float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
        constexpr int U = MODE == 0 ? 8 : 4;
        if (active) {
            for (int pb = p0 + r; pb < p1; pb += U * prow) {
                float4 xv[U], dv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pp = pb + u * prow;
                    const long row = (long)b * a.HW + (pp < p1 ? pp : pb);
                    xv[u] = ldg4(a.x + row * a.ldx + c);
                    if (MODE == 1) dv[u] = ldg4(a.dy + row * a.ldd + c);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pp = pb + u * prow;
                    if (pp < p1) {
                        const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
                        if (MODE == 0) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                a1[j] += xs[j];
                                a2[j] += xs[j] * xs[j];
                            }
                            if (a.cached) *reinterpret_cast<float4*>(slab + (long)(pp - p0) * W + 4 * q) = xv[u];
                        } else {
                            const float ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
                            float xh[4], dxh[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                xh[j] = (xs[j] - mu[j]) * rs[j];
                                float d = ds[j];
                                if (a.silu) {
                                    const float y = xh[j] * gm[j] + bt[j];
                                    const float sg = sigmoidf_(y);
                                    d *= sg * (1.f + y * (1.f - sg));
                                }
                                dxh[j] = d * gm[j];
                                a1[j] += dxh[j];
                                a2[j] += dxh[j] * xh[j];
                            }
                            if (a.cached) {
                                float* sp = slab + (long)(pp - p0) * W + 4 * q;
                                *reinterpret_cast<float4*>(sp) = make_float4(xh[0], xh[1], xh[2], xh[3]);
                                *reinterpret_cast<float4*>(sp + slabN) = make_float4(dxh[0], dxh[1], dxh[2], dxh[3]);
                            }
                        }
                    }
                }
            }
            *reinterpret_cast<float4*>(colp + (long)r * W + 4 * q) = make_float4(a1[0], a1[1], a1[2], a1[3]);
            *reinterpret_cast<float4*>(colp + (long)(prow + r) * W + 4 * q) = make_float4(a2[0], a2[1], a2[2], a2[3]);
        }
    }
    __syncthreads();
    {   // per-channel sums over the pixel lanes (SUB threads per channel), then per-group sums
        const int W2 = 2 * W;
        const int SUB = max(1, kGnClT / W2);
        for (int idx = t; idx < W2 * SUB; idx += kGnClT) {
            const int sub = idx / W2, col = idx - sub * W2;
            const int which = col >= W ? 1 : 0, cc = col - which * W;
            float acc = 0.f;
            for (int rr = sub; rr < prow; rr += SUB) acc += colp[(long)(which * prow + rr) * W + cc];
            red2[idx] = acc;
        }
        __syncthreads();
        if (t < 2 * gpc) {
            const int g = t >> 1, which = t & 1;
            float acc0 = 0.f, acc1 = 0.f;
            for (int sub = 0; sub < SUB; ++sub) {
                const float* src = red2 + (long)sub * W2 + which * W + g * Cg;
                for (int j = 0; j + 1 < Cg; j += 2) {
                    acc0 += src[j];
                    acc1 += src[j + 1];
                }
                if (Cg & 1) acc0 += src[Cg - 1];
            }
            part[t] = (double)acc0 + (double)acc1;
        }
    }
    // ---------------- cluster-wide totals through distributed shared memory, rank order
    cluster_arrive_();
    cluster_wait_();
    if (t < 64) {
        double tot = 0.0;
        if (t < 2 * gpc)
            for (int rk = 0; rk < cs; ++rk) tot += ld_cluster_f64(&part[t], (uint32_t)rk);
        const double other = __shfl_down_sync(0xffffffffu, tot, 1);
        if (t < 2 * gpc && !(t & 1)) {
            const int g = t >> 1;
            const double n = (double)Cg * a.HW;
            float2 o;
            if (MODE == 0) {
                const double m = tot / n;
                double var = other / n - m * m;
                if (var < 0.0) var = 0.0;
                o.x = (float)m;
                o.y = (float)(1.0 / sqrt(var + (double)a.eps));
                s_mean[g] = o.x;
                s_rstd[g] = o.y;
            } else {
                o.x = (float)(tot / n);
                o.y = (float)(other / n);
                s_m1[g] = o.x;
                s_m2[g] = o.y;
            }
            if (rank == 0) reinterpret_cast<float2*>(a.slot + (long)b * kSlotDoubles)[g0 + g] = o;
        }
    }
    cluster_arrive_();        // this CTA is done reading its peers' shared memory (matching wait before exit)
    __syncthreads();
    // ---------------- phase 2: apply
    if (active) {
        constexpr int U2 = 4;
        float k0[4], k1[4], k2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (MODE == 0) {
                k0[j] = s_mean[gl[j]];
                k1[j] = s_rstd[gl[j]] * gm[j];
                k2[j] = bt[j];
            } else {
                k0[j] = s_m1[gl[j]];
                k1[j] = s_m2[gl[j]];
                k2[j] = rs[j];
            }
        }
        for (int pb = p0 + r; pb < p1; pb += U2 * prow) {
            float4 v0[U2], v1[U2], av[U2];
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int pp = pb + u * prow;
                const int pc = pp < p1 ? pp : pb;
                const long row = (long)b * a.HW + pc;
                if (a.cached) {
                    const float* sp = slab + (long)(pc - p0) * W + 4 * q;
                    v0[u] = *reinterpret_cast<const float4*>(sp);
                    if (MODE == 1) v1[u] = *reinterpret_cast<const float4*>(sp + slabN);
                } else {
                    v0[u] = ldg4(a.x + row * a.ldx + c);
                    if (MODE == 1) v1[u] = ldg4(a.dy + row * a.ldd + c);
                }
                if (MODE == 1) av[u] = a.add ? ldg4(a.add + row * a.ldadd + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int pp = pb + u * prow;
                if (pp >= p1) continue;
                const long row = (long)b * a.HW + pp;
                const float xs[4] = {v0[u].x, v0[u].y, v0[u].z, v0[u].w};
                if (MODE == 0) {
                    float y[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float v = (xs[j] - k0[j]) * k1[j] + k2[j];
                        if (a.silu) v = v * sigmoidf_(v);
                        y[j] = v;
                    }
                    *reinterpret_cast<uint2*>(a.out16 + row * a.ld16 + c) = pack_half4(y[0], y[1], y[2], y[3]);
                    if (a.raw16) *reinterpret_cast<uint2*>(a.raw16 + row * a.ldraw + c) = pack_half4(xs[0], xs[1], xs[2], xs[3]);
                } else {
                    float xh[4], dxh[4];
                    if (a.cached) {
                        const float ds[4] = {v1[u].x, v1[u].y, v1[u].z, v1[u].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            xh[j] = xs[j];
                            dxh[j] = ds[j];
                        }
                    } else {
                        const float ds[4] = {v1[u].x, v1[u].y, v1[u].z, v1[u].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            xh[j] = (xs[j] - mu[j]) * rs[j];
                            float d = ds[j];
                            if (a.silu) {
                                const float y = xh[j] * gm[j] + bt[j];
                                const float sg = sigmoidf_(y);
                                d *= sg * (1.f + y * (1.f - sg));
                            }
                            dxh[j] = d * gm[j];
                        }
                    }
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = k2[j] * (dxh[j] - k0[j] - xh[j] * k1[j]);
                    o[0] += av[u].x; o[1] += av[u].y; o[2] += av[u].z; o[3] += av[u].w;
                    if (a.dx32) *reinterpret_cast<float4*>(a.dx32 + row * a.ld32 + c) = make_float4(o[0], o[1], o[2], o[3]);
                    if (a.out16) *reinterpret_cast<uint2*>(a.out16 + row * a.ld16 + c) = pack_half4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
    cluster_wait_();
}

// ---- statistics from the producer: normalise + SiLU in one streaming pass ------------------------------------------------
// The GEMM that wrote x also left per-channel partial sums (sum, sum of squares) of every block of rows it owned
// (gemm_tc.cuh, GemmDesc::colstat: [B][cap][2][ld] floats, `bps` blocks per sample used).  Each CTA here adds the partials of its
// channels in block order (fixed order: the result does not depend on timing), derives its groups' mean / rstd in double, and
// streams its [P pixels][W channels] slab once: x is read once and there is no reduction pass, no cluster, no barrier between
// CTAs.  A concatenated input (up path) brings two sources, one per producer.
struct GnNormArgs {
    const float* x; long ldx;
    int HW, C, Cg, gpc, W, P;
    const float* sp[2]; int scap[2], sbps[2], sc0[2], sc1[2]; long sld[2]; int nsrc;
    const float* gamma; const float* beta;
    float eps; int silu;
    double* slot;
    __half* out16; long ld16; __half* raw16; long ldraw;
};
constexpr int kGnNormT = 256;
constexpr int kGnNormMaxW = 512;

__global__ void __launch_bounds__(kGnNormT) gn_norm_kernel(const GnNormArgs a) {
    __shared__ double chs[2 * kGnNormMaxW];
    __shared__ float s_mean[kGroups], s_rstd[kGroups];
    const int t = threadIdx.x, b = blockIdx.z;
    const int W = a.W, Wq = W >> 2, Cg = a.Cg, gpc = a.gpc;
    const int c0 = blockIdx.x * W, g0 = blockIdx.x * gpc;
    const int p0 = blockIdx.y * a.P, p1 = min(a.HW, p0 + a.P);
    const int prow = kGnNormT / Wq;
    const int r = t / Wq, q = t - r * Wq;
    const bool active = r < prow;
    const int c = c0 + 4 * q;
    int gl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) gl[j] = (4 * q + j) / Cg;
    float gm[4], bt[4];
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            gm[j] = __ldg(a.gamma + c + j);
            bt[j] = __ldg(a.beta + c + j);
        }
    }
    pdl_wait();
    pdl_launch();
    // channel sums of this CTA's W channels (both statistics), blocks added in order
    for (int idx = t; idx < 2 * W; idx += kGnNormT) {
        const int st = idx >= W ? 1 : 0, col = idx - st * W;
        const int ch = c0 + col;
        const int k = (a.nsrc > 1 && ch >= a.sc0[1]) ? 1 : 0;
        const float* src = a.sp[k] + ((long)b * a.scap[k] * 2 + st) * a.sld[k] + (ch - a.sc0[k]);
        const long step = 2 * a.sld[k];
        double acc = 0.0;
        const int n = a.sbps[k];
        int i = 0;
        for (; i + 4 <= n; i += 4) {
            const float v0 = __ldcg(src + (long)i * step), v1 = __ldcg(src + (long)(i + 1) * step);
            const float v2 = __ldcg(src + (long)(i + 2) * step), v3 = __ldcg(src + (long)(i + 3) * step);
            acc += (double)v0; acc += (double)v1; acc += (double)v2; acc += (double)v3;
        }
        for (; i < n; ++i) acc += (double)__ldcg(src + (long)i * step);
        chs[idx] = acc;
    }
    __syncthreads();
    if (t < gpc) {
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < Cg; ++j) {
            s1 += chs[t * Cg + j];
            s2 += chs[W + t * Cg + j];
        }
        const double n = (double)Cg * a.HW;
        const double m = s1 / n;
        double var = s2 / n - m * m;
        if (var < 0.0) var = 0.0;
        float2 o;
        o.x = (float)m;
        o.y = (float)(1.0 / sqrt(var + (double)a.eps));
        s_mean[t] = o.x;
        s_rstd[t] = o.y;
        if (blockIdx.y == 0) reinterpret_cast<float2*>(a.slot + (long)b * kSlotDoubles)[g0 + t] = o;
    }
    __syncthreads();
    if (!active) return;
    float k0[4], k1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        k0[j] = s_mean[gl[j]];
        k1[j] = s_rstd[gl[j]] * gm[j];
    }
    constexpr int U = 4;
    for (int pb = p0 + r; pb < p1; pb += U * prow) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = pb + u * prow;
            const long row = (long)b * a.HW + (pp < p1 ? pp : pb);
            v[u] = ldg4(a.x + row * a.ldx + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = pb + u * prow;
            if (pp >= p1) continue;
            const long row = (long)b * a.HW + pp;
            const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            float y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float w = (xs[j] - k0[j]) * k1[j] + bt[j];
                if (a.silu) w = w * sigmoidf_(w);
                y[j] = w;
            }
            *reinterpret_cast<uint2*>(a.out16 + row * a.ld16 + c) = pack_half4(y[0], y[1], y[2], y[3]);
            if (a.raw16) *reinterpret_cast<uint2*>(a.raw16 + row * a.ldraw + c) = pack_half4(xs[0], xs[1], xs[2], xs[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row; the row is read ONCE into registers (all loads in flight together: NV float4 per lane cover C <= 128 NV
// channels), so the kernel is one memory round trip + two warp reductions instead of three dependent passes.
template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long ldx, long rows, int C,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float eps, __half* __restrict__ out16, long ld16,
                                                     float* __restrict__ stats) {
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * ldx;
    const int vec = C >> 2;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int q = lane + 32 * i;
        v[i] = q < vec ? ldg4(xr + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + 32 * i < vec) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / C + eps);
    if (stats && lane == 0) {
        stats[row * 2 + 0] = mean;
        stats[row * 2 + 1] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int q = lane + 32 * i;
        if (q < vec) {
            const float4 g = ldg4(gamma + 4 * q);
            const float4 bb = ldg4(beta + 4 * q);
            *reinterpret_cast<uint2*>(out16 + row * ld16 + 4 * q) =
                pack_half4((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y,
                           (v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, long ldd,
                                                     const float* __restrict__ x, long ldx, long rows, int C,
                                                     const float* __restrict__ gamma, const float* __restrict__ stats,
                                                     const float* __restrict__ add, long ldadd,
                                                     float* __restrict__ dx32, long ld32, __half* __restrict__ dx16,
                                                     long ld16) {
    pdl_wait();
    pdl_launch();
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * ldx;
    const float* dr = dy + row * ldd;
    const int vec = C >> 2;
    // xh = normalised input, hd = dy * gamma: both kept in registers between the reduction and the apply
    float4 xh[NV], hd[NV];
    const float mean = stats[row * 2 + 0], rstd = stats[row * 2 + 1];
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int q = lane + 32 * i;
        if (q < vec) {
            const float4 v = ldg4(xr + 4 * q);
            const float4 d = ldg4(dr + 4 * q);
            const float4 g = ldg4(gamma + 4 * q);
            hd[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
            xh[i] = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
            a1 += (hd[i].x + hd[i].y) + (hd[i].z + hd[i].w);
            a2 += hd[i].x * xh[i].x + hd[i].y * xh[i].y + hd[i].z * xh[i].z + hd[i].w * xh[i].w;
        } else {
            hd[i] = xh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float m1 = warp_sum(a1) / C, m2 = warp_sum(a2) / C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int q = lane + 32 * i;
        if (q < vec) {
            float o0 = rstd * (hd[i].x - m1 - xh[i].x * m2);
            float o1 = rstd * (hd[i].y - m1 - xh[i].y * m2);
            float o2 = rstd * (hd[i].z - m1 - xh[i].z * m2);
            float o3 = rstd * (hd[i].w - m1 - xh[i].w * m2);
            if (add) {
                const float4 av = ldg4(add + row * ldadd + 4 * q);
                o0 += av.x; o1 += av.y; o2 += av.z; o3 += av.w;
            }
            if (dx32) *reinterpret_cast<float4*>(dx32 + row * ld32 + 4 * q) = make_float4(o0, o1, o2, o3);
            if (dx16) *reinterpret_cast<uint2*>(dx16 + row * ld16 + 4 * q) = pack_half4(o0, o1, o2, o3);
        }
    }
}

// ------------------------------------------------------------------------------------------------ softmax
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ s, long lds, long rows, int n,
                                                          __half* __restrict__ p16, long ldp) {
    pdl_wait();
    pdl_launch();
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
    pdl_wait();
    pdl_launch();
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

__device__ __forceinline__ float4 ldh4(const __half* p) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__global__ void __launch_bounds__(256) geglu_fwd_kernel(const __half* __restrict__ ff, long ldf, long rows, int F,
                                                        __half* __restrict__ out16, long ld16) {
    pdl_wait();
    pdl_launch();
    const int vec = F >> 2;
    const long total = rows * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / vec;
        const int c = (int)(idx - r * vec) << 2;
        // interleaved projection: feature c's value sits at column 64 (c / 32) + c % 32, its gate 32 columns further
        const int ca = ((c >> 5) << 6) + (c & 31);
        const float4 a = ldh4(ff + r * ldf + ca);
        const float4 g = ldh4(ff + r * ldf + ca + 32);
        *reinterpret_cast<uint2*>(out16 + r * ld16 + c) =
            pack_half4(a.x * gelu_exact(g.x), a.y * gelu_exact(g.y), a.z * gelu_exact(g.z), a.w * gelu_exact(g.w));
    }
}

__global__ void __launch_bounds__(256) geglu_bwd_kernel(const float* __restrict__ dg, long ldg,
                                                        const __half* __restrict__ ff, long ldf, long rows, int F,
                                                        __half* __restrict__ dff16, long ld16) {
    pdl_wait();
    pdl_launch();
    const int vec = F >> 2;
    const long total = rows * vec;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx / vec;
        const int c = (int)(idx - r * vec) << 2;
        const int ca = ((c >> 5) << 6) + (c & 31);      // interleaved layout (see geglu_fwd_kernel)
        const float4 d = ldg4(dg + r * ldg + c);
        const float4 a = ldh4(ff + r * ldf + ca);
        const float4 g = ldh4(ff + r * ldf + ca + 32);
        *reinterpret_cast<uint2*>(dff16 + r * ld16 + ca) =
            pack_half4(d.x * gelu_exact(g.x), d.y * gelu_exact(g.y), d.z * gelu_exact(g.z), d.w * gelu_exact(g.w));
        *reinterpret_cast<uint2*>(dff16 + r * ld16 + ca + 32) =
            pack_half4(d.x * a.x * gelu_grad(g.x), d.y * a.y * gelu_grad(g.y), d.z * a.z * gelu_grad(g.z),
                       d.w * a.w * gelu_grad(g.w));
    }
}

// ------------------------------------------------------------------------------------------------ movement
__global__ void __launch_bounds__(256) add2d_kernel(const float* __restrict__ a, long lda, const float* __restrict__ b,
                                                    long ldb, long rows, int cols, float* __restrict__ d32, long ld32,
                                                    __half* __restrict__ d16, long ld16) {
    pdl_wait();
    pdl_launch();
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
    pdl_wait();
    pdl_launch();
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
    pdl_wait();
    pdl_launch();
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
    pdl_wait();
    pdl_launch();
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
    pdl_wait();
    pdl_launch();
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

// General form (any channel count: the 3-channel image of the VAE encoder, the 1-channel sketch): one element per thread.
__global__ void __launch_bounds__(256) im2col3x3_scalar_kernel(const float* __restrict__ x, long ldx, int B, int H, int W,
                                                               int C, int stride, int pad, int Ho, int Wo,
                                                               __half* __restrict__ col, long ldcol) {
    pdl_wait();
    pdl_launch();
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
            const int iy = oy * stride + tap / 3 - pad;
            const int ix = ox * stride + tap % 3 - pad;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long)b * H + iy) * W + ix) * ldx + c];
        }
        col[idx] = __float2half_rn(v);
    }
}

// One thread per four channels of one tap of one output pixel (C % 4 == 0, ldcol % 4 == 0): a float4 load, an 8-byte store, 32-bit
// index arithmetic (the element-per-thread form spent its time in seven 64-bit divisions per fp16 written).
__global__ void __launch_bounds__(256) im2col3x3_kernel(const float* __restrict__ x, long ldx, int B, int H, int W,
                                                        int C, int stride, int pad, int Ho, int Wo, __half* __restrict__ col,
                                                        long ldcol) {
    pdl_wait();
    pdl_launch();
    const int qrow = (int)(ldcol >> 2);                    // quads per output row
    const int total = B * Ho * Wo * qrow;
    const int kmax = 9 * C;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int pix = idx / qrow;
        const int k = (idx - pix * qrow) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kmax) {
            const int tap = k / C, c = k - tap * C;
            const int t3 = tap / 3;
            const int ox = pix % Wo;
            const int r = pix / Wo;
            const int oy = r % Ho;
            const int b = r / Ho;
            const int iy = oy * stride + t3 - pad;
            const int ix = ox * stride + (tap - 3 * t3) - pad;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = ldg4(x + (((long)b * H + iy) * W + ix) * ldx + c);
        }
        *reinterpret_cast<uint2*>(col + (long)pix * ldcol + k) = pack_half4(v.x, v.y, v.z, v.w);
    }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int B, int C, int H, int W, float* __restrict__ dst,
                                    long ldn) {
    pdl_wait();
    pdl_launch();
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
    pdl_wait();
    pdl_launch();
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
// out[n] = bias[n] + sum_k w[n][k] * act(x[k]): one warp per output row, 16-byte weight loads (all of a lane's loads independent and in
// flight together), the activated input vector staged once per block in shared memory.  K % 8 == 0, K <= 2048.
__global__ void __launch_bounds__(256) gemv_kernel(const float* __restrict__ x, int K, const __half* __restrict__ w,
                                                   const float* __restrict__ bias, int N, int silu_in,
                                                   float* __restrict__ out) {
    __shared__ __align__(16) float xs[2048];
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nv = K >> 3;                       // 16-byte units per weight row
    // the weights do not depend on the preceding kernel: request this lane's units before waiting for it
    uint4 wv[8];
    const uint4* wr = reinterpret_cast<const uint4*>(w + (long)(n < N ? n : 0) * K);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int u = lane + 32 * i;
        wv[i] = u < nv ? __ldg(wr + u) : make_uint4(0u, 0u, 0u, 0u);
    }
    pdl_wait();
    pdl_launch();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float v = x[k];
        if (silu_in) v = v * sigmoidf_(v);
        xs[k] = v;
    }
    __syncthreads();
    if (n >= N) return;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int u = lane + 32 * i;
        if (u < nv) {
            const __half2* h = reinterpret_cast<const __half2*>(&wv[i]);
            const float4 a = *reinterpret_cast<const float4*>(xs + 8 * u), c = *reinterpret_cast<const float4*>(xs + 8 * u + 4);
            const float2 w0 = __half22float2(h[0]), w1 = __half22float2(h[1]), w2 = __half22float2(h[2]), w3 = __half22float2(h[3]);
            acc += (w0.x * a.x + w0.y * a.y) + (w1.x * a.z + w1.y * a.w) + (w2.x * c.x + w2.y * c.y) + (w3.x * c.z + w3.y * c.w);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[n] = acc + (bias ? bias[n] : 0.f);
}

__global__ void timestep_embedding_kernel(float t, int dim, float* __restrict__ out) {
    pdl_wait();
    pdl_launch();
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

// pixels per reduction block: ~2 blocks per SM over the whole batch, at least 4 pixels each
static int gn_chunk(int B, int HW) {
    int nblk = (2 * kNumSMs) / (B > 0 ? B : 1);
    if (nblk < 1) nblk = 1;
    int P = ceil_div(HW, nblk);
    if (P < 4) P = 4;
    return P;
}

// per-block partial sums of the running reduction (one stream at a time uses the library)
static double* g_gn_partial = nullptr;
static size_t g_gn_partial_cap = 0;
static int g_gn_partial_dev = -1;
static int gn_partial(size_t doubles, double** out) {
    int dev = 0;
    S2I_CUDA(cudaGetDevice(&dev));
    if (dev != g_gn_partial_dev || doubles > g_gn_partial_cap) {
        if (g_gn_partial && dev == g_gn_partial_dev) {
            S2I_CUDA(cudaDeviceSynchronize());
            cudaFree(g_gn_partial);
        }
        size_t cap = doubles < (size_t)(1u << 18) ? (size_t)(1u << 18) : doubles * 2;
        void* p = nullptr;
        if (cudaMalloc(&p, cap * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return set_error(S2I_ERR_OOM, "gn: cannot allocate the partial-sum scratch");
        }
        ++g_alloc_gen;
        g_gn_partial = static_cast<double*>(p);
        g_gn_partial_cap = cap;
        g_gn_partial_dev = dev;
    }
    *out = g_gn_partial;
    return 0;
}

int gn_stats(const float* x, long ldx, int B, int HW, int C, float eps, double* sums, cudaStream_t st) {
    S2I_REQ(C % (4 * 1) == 0 && C % kGroups == 0 && (ldx & 3) == 0, "gn_stats: C must be a multiple of 32 and ld of 4");
    const int P = gn_chunk(B, HW);
    dim3 grid(ceil_div(HW, P), B);
    double* partial = nullptr;
    S2I_TRY(gn_partial((size_t)B * grid.x * 2 * kGroups, &partial));
    S2I_LAUNCH((gn_reduce_kernel<0>), grid, 256, 0, st, x, ldx, nullptr, 0, HW, C, P, nullptr, nullptr, nullptr, eps, 0, partial, sums);
    S2I_LAUNCH_CHECK();
    return 0;
}

int gn_apply(const float* x, long ldx, int B, int HW, int C, const double* sums, const float* gamma, const float* beta,
             float eps, int silu, void* out16, long ld16, void* raw16, long ldraw, cudaStream_t st) {
    S2I_REQ(C % kGroups == 0 && (ldx & 3) == 0 && (ld16 & 3) == 0 && (ldraw & 3) == 0, "gn_apply: alignment");
    dim3 grid(grid_for((long)HW * (C >> 2), 256, 148 * 8), B);
    S2I_LAUNCH((gn_apply_kernel), grid, 256, 0, st, x, ldx, HW, C, sums, gamma, beta, eps, silu, (__half*)out16, ld16,
                                         (__half*)raw16, ldraw);
    S2I_LAUNCH_CHECK();
    return 0;
}

int gn_bwd_stats(const float* dy, long ldd, const float* x, long ldx, int B, int HW, int C, const double* sums,
                 const float* gamma, const float* beta, float eps, int silu, double* bsums, cudaStream_t st) {
    S2I_REQ(C % kGroups == 0 && (ldx & 3) == 0 && (ldd & 3) == 0, "gn_bwd_stats: alignment");
    const int P = gn_chunk(B, HW);
    dim3 grid(ceil_div(HW, P), B);
    double* partial = nullptr;
    S2I_TRY(gn_partial((size_t)B * grid.x * 2 * kGroups, &partial));
    S2I_LAUNCH((gn_reduce_kernel<1>), grid, 256, 0, st, x, ldx, dy, ldd, HW, C, P, sums, gamma, beta, eps, silu, partial, bsums);
    S2I_LAUNCH_CHECK();
    return 0;
}

int gn_bwd_apply(const float* dy, long ldd, const float* x, long ldx, int B, int HW, int C, const double* sums,
                 const double* bsums, const float* gamma, const float* beta, float eps, int silu, const float* add,
                 long ldadd, float* dx32, long ld32, void* dx16, long ld16, cudaStream_t st) {
    S2I_REQ(C % kGroups == 0 && (ldx & 3) == 0 && (ldd & 3) == 0 && (ldadd & 3) == 0 && (ld32 & 3) == 0 && (ld16 & 3) == 0,
            "gn_bwd_apply: alignment");
    dim3 grid(grid_for((long)HW * (C >> 2), 256, 148 * 8), B);
    S2I_LAUNCH((gn_bwd_apply_kernel), grid, 256, 0, st, dy, ldd, x, ldx, HW, C, sums, bsums, gamma, beta, eps, silu, add, ldadd,
                                             dx32, ld32, (__half*)dx16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

// ---- cluster launchers (see gn_cluster_kernel)
static int g_gn_cluster = -1;        // S2I_GN_CLUSTER=0 in the environment selects the grid-barrier form instead
static const size_t kGnClSmemMax = 200 * 1024;
static int gn_cluster_prepare();
// How many clusters of `cs` CTAs with `smem` bytes each the device keeps resident at once (the launch's wave size).
// Measured on B200: 8-CTA clusters of this kernel reach 15, 4-CTA clusters 32+ -- large clusters leave SMs idle.
template <int MODE>
static int gn_cluster_capacity(int cs, size_t smem) {
    static std::map<long, int> cache;
    const long key = ((long)cs << 32) | (long)(smem >> 10);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64, 1, 1);
    cfg.blockDim = dim3(kGnClT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gn_cluster_kernel<MODE>, &cfg) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = cs <= 4 ? 128 / cs : 8;      // what B200 gave when this was written
    }
    cache[key] = n;
    return n;
}
// Picks groups per cluster (gpc), CTAs per cluster (cs) and whether the slab is staged in shared memory with a small
// cost model fitted to tools/gn_bench.py --sweep:  waves x (fixed + slab elements per CTA, x1.25 when the apply phase
// re-reads global memory), with rows shorter than 128 bytes charged for their wasted sectors.
static bool gn_cluster_geometry(int B, int HW, int C, int mode, GnClArgs* out, size_t* smem) {
    if (g_gn_cluster < 0) {
        const char* e = getenv("S2I_GN_CLUSTER");
        g_gn_cluster = (e && e[0] == '0') ? 0 : 1;
    }
    if (!g_gn_cluster || C % kGroups) return false;
    if (gn_cluster_prepare() != 0) return false;
    // tools/gn_bench.py --sweep: S2I_GN_GEOM="gpc,cs" forces one geometry (ignored when it is not legal for the shape)
    int f_gpc = 0, f_cs = 0;
    if (const char* e = getenv("S2I_GN_GEOM")) sscanf(e, "%d,%d", &f_gpc, &f_cs);
    struct Choice { GnClArgs a; size_t smem; bool ok; };
    static std::map<long, Choice> decided;
    const long key = ((long)B << 48) ^ ((long)HW << 24) ^ ((long)C << 4) ^ (long)mode;
    if (!f_gpc && !f_cs) {
        auto it = decided.find(key);
        if (it != decided.end()) {
            *out = it->second.a;
            *smem = it->second.smem;
            return it->second.ok;
        }
    }
    const int Cg = C / kGroups;
    double best = 1e30;
    bool found = false;
    const size_t fixed = (size_t)(kGnClColp + kGnClRed) * sizeof(float);
    for (int gpc = 1; gpc <= kGroups; gpc *= 2) {
        const int W = gpc * Cg;
        if (W % 4 || W > kGnClMaxW) continue;
        if (f_gpc && gpc != f_gpc) continue;
        for (int cs = 1; cs <= 16; cs *= 2) {
            if (f_cs ? cs != f_cs : cs > 8) continue;
            const int P = ceil_div(HW, cs);
            const double elems = (double)P * W * (mode ? 2 : 1);
            const size_t slab = (size_t)P * W * sizeof(float) * (mode ? 2 : 1);
            const bool fits = fixed + slab <= kGnClSmemMax;
            const size_t need = fixed + (fits ? slab : 0);
            const long clusters = (long)B * (kGroups / gpc);
            const int cap = mode ? gn_cluster_capacity<1>(cs, need) : gn_cluster_capacity<0>(cs, need);
            const double waves = (double)((clusters + cap - 1) / cap);
            double cost = waves * (25000.0 + elems * (fits ? 1.0 : 1.25));
            if (W * 4 < 128) cost *= 1.0 + (128 - W * 4) / 256.0;      // short rows waste sectors
            if (cost < best) {
                best = cost;
                found = true;
                memset(out, 0, sizeof(*out));
                out->HW = HW; out->C = C; out->Cg = Cg; out->gpc = gpc; out->W = W; out->P = P; out->cs = cs;
                out->cached = fits ? 1 : 0;
                *smem = need;
            }
        }
    }
    if (!f_gpc && !f_cs) decided[key] = Choice{*out, *smem, found};
    return found;
}
static int gn_cluster_prepare() {
    static int dev_done = -1;
    int dev = 0;
    S2I_CUDA(cudaGetDevice(&dev));
    if (dev_done == dev) return 0;
    S2I_CUDA(cudaFuncSetAttribute(gn_cluster_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnClSmemMax));
    S2I_CUDA(cudaFuncSetAttribute(gn_cluster_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnClSmemMax));
    S2I_CUDA(cudaFuncSetAttribute(gn_cluster_kernel<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    S2I_CUDA(cudaFuncSetAttribute(gn_cluster_kernel<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    dev_done = dev;
    return 0;
}

// ---- fused launchers (see gn_fused_kernel); fall back to the two-kernel form when the batch exceeds the SM count
static bool gn_fused_geometry(int B, int HW, int C, int* nblk, int* P) {
    if (B > kNumSMs || C > 2560) return false;
    int n = kNumSMs / B;
    int p = ceil_div(HW, n);
    if (p < 4) p = 4;
    n = ceil_div(HW, p);
    *nblk = n;
    *P = p;
    return true;
}

int gn_forward(const float* x, long ldx, int B, int HW, int C, double* slot, const float* gamma, const float* beta, float eps,
               int silu, void* out16, long ld16, void* raw16, long ldraw, cudaStream_t st) {
    S2I_REQ(C % 4 == 0 && C % kGroups == 0 && (ldx & 3) == 0 && (ld16 & 3) == 0 && (ldraw & 3) == 0, "gn_forward: alignment");
    GnClArgs ca;
    size_t csmem;
    if (gn_cluster_geometry(B, HW, C, 0, &ca, &csmem)) {
        ca.x = x; ca.ldx = ldx;
        ca.gamma = gamma; ca.beta = beta; ca.eps = eps; ca.silu = silu;
        ca.slot = slot;
        ca.out16 = (__half*)out16; ca.ld16 = ld16; ca.raw16 = (__half*)raw16; ca.ldraw = ldraw;
        launch_kernel_cluster(gn_cluster_kernel<0>, dim3((32 / ca.gpc) * ca.cs, B), dim3(kGnClT), csmem, ca.cs, st, ca);
        S2I_LAUNCH_CHECK();
        return 0;
    }
    int nblk, P;
    if (!gn_fused_geometry(B, HW, C, &nblk, &P)) {
        S2I_TRY(gn_stats(x, ldx, B, HW, C, eps, slot, st));
        return gn_apply(x, ldx, B, HW, C, slot, gamma, beta, eps, silu, out16, ld16, raw16, ldraw, st);
    }
    GnFusedArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.ldx = ldx; a.HW = HW; a.C = C; a.P = P;
    a.gamma = gamma; a.beta = beta; a.eps = eps; a.silu = silu;
    S2I_TRY(gn_partial((size_t)B * nblk * 2 * kGroups, &a.partial));
    a.slot = slot;
    a.out16 = (__half*)out16; a.ld16 = ld16; a.raw16 = (__half*)raw16; a.ldraw = ldraw;
    S2I_LAUNCH((gn_fused_kernel<0>), dim3(nblk, B), 512, 0, st, a);
    S2I_LAUNCH_CHECK();
    return 0;
}

bool gn_norm_supported(int C) {
    static const bool on = [] { const char* e = getenv("S2I_GN_FUSED_STATS"); return !(e && e[0] == '0'); }();
    if (!on || C % kGroups) return false;
    const int Cg = C / kGroups;
    for (int gpc = 1; gpc <= kGroups; gpc *= 2)
        if ((gpc * Cg) % 4 == 0 && gpc * Cg <= kGnNormMaxW) return true;
    return false;
}

int gn_norm(const float* x, long ldx, int B, int HW, int C, const GnStatSrc* src, int nsrc, double* slot, const float* gamma,
            const float* beta, float eps, int silu, void* out16, long ld16, void* raw16, long ldraw, cudaStream_t st) {
    S2I_REQ(C % 4 == 0 && C % kGroups == 0 && (ldx & 3) == 0 && (ld16 & 3) == 0 && (ldraw & 3) == 0, "gn_norm: alignment");
    S2I_REQ(nsrc >= 1 && nsrc <= 2, "gn_norm: one or two statistics sources");
    GnNormArgs a;
    memset(&a, 0, sizeof(a));
    const int Cg = C / kGroups;
    // channels per CTA: whole groups, rows of at least 256 bytes where the group size allows
    int gpc = 0;
    for (int g = 1; g <= kGroups; g *= 2) {
        const int W = g * Cg;
        if (W % 4 || W > kGnNormMaxW) continue;
        gpc = g;
        if (W >= 64) break;
    }
    S2I_REQ(gpc > 0, "gn_norm: no channel chunk fits this channel count");
    a.x = x; a.ldx = ldx; a.HW = HW; a.C = C; a.Cg = Cg; a.gpc = gpc; a.W = gpc * Cg;
    const int chunks = kGroups / gpc;
    // pixel chunks: about four CTAs per SM in total, at least 32 pixel rows each
    int np = ceil_div(4 * kNumSMs, B * chunks);
    const int prow = kGnNormT / (a.W / 4);
    const int max_np = ceil_div(HW, 4 * prow > 32 ? 4 * prow : 32);
    if (np > max_np) np = max_np;
    if (np < 1) np = 1;
    a.P = ceil_div(HW, np);
    np = ceil_div(HW, a.P);
    for (int k = 0; k < nsrc; ++k) {
        a.sp[k] = src[k].p; a.scap[k] = src[k].cap; a.sbps[k] = src[k].bps; a.sld[k] = src[k].ld;
        a.sc0[k] = src[k].c0; a.sc1[k] = src[k].c1;
    }
    a.nsrc = nsrc;
    a.gamma = gamma; a.beta = beta; a.eps = eps; a.silu = silu;
    a.slot = slot;
    a.out16 = (__half*)out16; a.ld16 = ld16; a.raw16 = (__half*)raw16; a.ldraw = ldraw;
    S2I_LAUNCH((gn_norm_kernel), dim3(chunks, np, B), kGnNormT, 0, st, a);
    S2I_LAUNCH_CHECK_TAG("gn_forward", 0.0, 0.0);
    return 0;
}

int gn_backward(const float* dy, long ldd, const float* x, long ldx, int B, int HW, int C, const double* fslot, double* bslot,
                const float* gamma, const float* beta, float eps, int silu, const float* add, long ldadd, float* dx32,
                long ld32, void* dx16, long ld16, cudaStream_t st) {
    S2I_REQ(C % 4 == 0 && C % kGroups == 0 && (ldx & 3) == 0 && (ldd & 3) == 0 && (ldadd & 3) == 0 && (ld32 & 3) == 0 &&
                (ld16 & 3) == 0, "gn_backward: alignment");
    GnClArgs ca;
    size_t csmem;
    if (gn_cluster_geometry(B, HW, C, 1, &ca, &csmem)) {
        ca.x = x; ca.ldx = ldx; ca.dy = dy; ca.ldd = ldd;
        ca.fslot = fslot;
        ca.gamma = gamma; ca.beta = beta; ca.eps = eps; ca.silu = silu;
        ca.slot = bslot;
        ca.add = add; ca.ldadd = ldadd; ca.dx32 = dx32; ca.ld32 = ld32; ca.out16 = (__half*)dx16; ca.ld16 = ld16;
        launch_kernel_cluster(gn_cluster_kernel<1>, dim3((32 / ca.gpc) * ca.cs, B), dim3(kGnClT), csmem, ca.cs, st, ca);
        S2I_LAUNCH_CHECK();
        return 0;
    }
    int nblk, P;
    if (!gn_fused_geometry(B, HW, C, &nblk, &P)) {
        S2I_TRY(gn_bwd_stats(dy, ldd, x, ldx, B, HW, C, fslot, gamma, beta, eps, silu, bslot, st));
        return gn_bwd_apply(dy, ldd, x, ldx, B, HW, C, fslot, bslot, gamma, beta, eps, silu, add, ldadd, dx32, ld32, dx16,
                            ld16, st);
    }
    GnFusedArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.ldx = ldx; a.dy = dy; a.ldd = ldd; a.HW = HW; a.C = C; a.P = P;
    a.fslot = fslot;
    a.gamma = gamma; a.beta = beta; a.eps = eps; a.silu = silu;
    S2I_TRY(gn_partial((size_t)B * nblk * 2 * kGroups, &a.partial));
    a.slot = bslot;
    a.add = add; a.ldadd = ldadd; a.dx32 = dx32; a.ld32 = ld32; a.out16 = (__half*)dx16; a.ld16 = ld16;
    S2I_LAUNCH((gn_fused_kernel<1>), dim3(nblk, B), 512, 0, st, a);
    S2I_LAUNCH_CHECK();
    return 0;
}

int ln_fwd(const float* x, long ldx, long rows, int C, const float* gamma, const float* beta, float eps, void* out16,
           long ld16, float* stats, cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldx & 3) == 0 && (ld16 & 3) == 0, "ln_fwd: alignment");
    S2I_REQ(C <= 2560, "ln_fwd: at most 2560 channels (the row is held in registers)");
    const unsigned grid = (unsigned)ceil_div_l(rows, 8);
    __half* o = (__half*)out16;
    // NV float4 per lane: 320 -> 3, 640 -> 5, 1280 -> 10, up to 2560 -> 20
    if (C <= 384) S2I_LAUNCH((ln_fwd_kernel<3>), grid, 256, 0, st, x, ldx, rows, C, gamma, beta, eps, o, ld16, stats);
    else if (C <= 640) S2I_LAUNCH((ln_fwd_kernel<5>), grid, 256, 0, st, x, ldx, rows, C, gamma, beta, eps, o, ld16, stats);
    else if (C <= 1280) S2I_LAUNCH((ln_fwd_kernel<10>), grid, 256, 0, st, x, ldx, rows, C, gamma, beta, eps, o, ld16, stats);
    else S2I_LAUNCH((ln_fwd_kernel<20>), grid, 256, 0, st, x, ldx, rows, C, gamma, beta, eps, o, ld16, stats);
    S2I_LAUNCH_CHECK();
    return 0;
}

int ln_bwd(const float* dy, long ldd, const float* x, long ldx, long rows, int C, const float* gamma,
           const float* stats, const float* add, long ldadd, float* dx32, long ld32, void* dx16, long ld16,
           cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldx & 3) == 0 && (ldd & 3) == 0 && (ldadd & 3) == 0 && (ld32 & 3) == 0 && (ld16 & 3) == 0,
            "ln_bwd: alignment");
    S2I_REQ(C <= 1280, "ln_bwd: at most 1280 channels (the row and its gradient are held in registers)");
    const unsigned grid = (unsigned)ceil_div_l(rows, 8);
    __half* o = (__half*)dx16;
    if (C <= 384)
        S2I_LAUNCH((ln_bwd_kernel<3>), grid, 256, 0, st, dy, ldd, x, ldx, rows, C, gamma, stats, add, ldadd, dx32, ld32, o, ld16);
    else if (C <= 640)
        S2I_LAUNCH((ln_bwd_kernel<5>), grid, 256, 0, st, dy, ldd, x, ldx, rows, C, gamma, stats, add, ldadd, dx32, ld32, o, ld16);
    else
        S2I_LAUNCH((ln_bwd_kernel<10>), grid, 256, 0, st, dy, ldd, x, ldx, rows, C, gamma, stats, add, ldadd, dx32, ld32, o, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int softmax_fwd(const float* s, long lds, long rows, int n, void* p16, long ldp, cudaStream_t st) {
    S2I_LAUNCH((softmax_fwd_kernel), (unsigned)ceil_div_l(rows, 8), 256, 0, st, s, lds, rows, n, (__half*)p16, ldp);
    S2I_LAUNCH_CHECK();
    return 0;
}

int softmax_bwd(const void* p16, long ldp, const float* dp, long lddp, long rows, int n, float scale, void* ds16,
                long ldds, cudaStream_t st) {
    S2I_LAUNCH((softmax_bwd_kernel), (unsigned)ceil_div_l(rows, 8), 256, 0, st, (const __half*)p16, ldp, dp, lddp, rows, n, scale,
                                                                     (__half*)ds16, ldds);
    S2I_LAUNCH_CHECK();
    return 0;
}

int geglu_fwd(const void* ff16, long ldf, long rows, int F, void* out16, long ld16, cudaStream_t st) {
    const __half* ff = static_cast<const __half*>(ff16);
    S2I_REQ((F & 31) == 0 && (ldf & 3) == 0 && (ld16 & 3) == 0, "geglu_fwd: alignment (interleaved 32-feature blocks)");
    S2I_LAUNCH((geglu_fwd_kernel), grid_for(rows * (F >> 2), 256), 256, 0, st, ff, ldf, rows, F, (__half*)out16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int geglu_bwd(const float* dg, long ldg, const void* ff16, long ldf, long rows, int F, void* dff16, long ld16,
              cudaStream_t st) {
    const __half* ff = static_cast<const __half*>(ff16);
    S2I_REQ((F & 31) == 0 && (ldf & 3) == 0 && (ldg & 3) == 0 && (ld16 & 3) == 0, "geglu_bwd: alignment (interleaved 32-feature blocks)");
    S2I_LAUNCH((geglu_bwd_kernel), grid_for(rows * (F >> 2), 256), 256, 0, st, dg, ldg, ff, ldf, rows, F, (__half*)dff16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int add2d(const float* a, long lda, const float* b, long ldb, long rows, int cols, float* dst32, long ld32, void* dst16,
          long ld16, cudaStream_t st) {
    S2I_REQ((cols & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0 && (ld32 & 3) == 0 && (ld16 & 3) == 0, "add2d: alignment");
    S2I_LAUNCH((add2d_kernel), grid_for(rows * (cols >> 2), 256), 256, 0, st, a, lda, b, ldb, rows, cols, dst32, ld32,
                                                                   (__half*)dst16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int cast2d(const float* a, long lda, long rows, int cols, float mul, void* dst16, long ld16, cudaStream_t st) {
    S2I_REQ((cols & 3) == 0 && (lda & 3) == 0 && (ld16 & 3) == 0, "cast2d: alignment");
    S2I_LAUNCH((cast2d_kernel), grid_for(rows * (cols >> 2), 256), 256, 0, st, a, lda, rows, cols, mul, (__half*)dst16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int upsample2x(const float* x, long ldx, int B, int H, int W, int C, void* out16, long ld16, cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldx & 3) == 0 && (ld16 & 3) == 0, "upsample2x: alignment");
    S2I_LAUNCH((upsample2x_kernel), grid_for((long)B * 4 * H * W * (C >> 2), 256), 256, 0, st, x, ldx, B, H, W, C, (__half*)out16,
                                                                                    ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int sumpool2x(const float* d, long ldd, int B, int H, int W, int C, float* dx, long ldx, cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldx & 3) == 0 && (ldd & 3) == 0, "sumpool2x: alignment");
    S2I_LAUNCH((sumpool2x_kernel), grid_for((long)B * H * W * (C >> 2), 256), 256, 0, st, d, ldd, B, H, W, C, dx, ldx);
    S2I_LAUNCH_CHECK();
    return 0;
}

int zero_insert2x(const float* d, long ldd, int B, int Ho, int Wo, int C, void* out16, long ld16, cudaStream_t st) {
    S2I_REQ((C & 3) == 0 && (ldd & 3) == 0 && (ld16 & 3) == 0, "zero_insert2x: alignment");
    S2I_LAUNCH((zero_insert2x_kernel), grid_for((long)B * 4 * Ho * Wo * (C >> 2), 256), 256, 0, st, d, ldd, B, Ho, Wo, C,
                                                                                         (__half*)out16, ld16);
    S2I_LAUNCH_CHECK();
    return 0;
}

int im2col3x3(const float* x, long ldx, int B, int H, int W, int C, int stride, void* col16, long ldcol,
              cudaStream_t st, int pad) {
    S2I_REQ(stride == 1 || stride == 2, "im2col3x3: stride");
    S2I_REQ(pad == 1 || (pad == 0 && stride == 2), "im2col3x3: pad 1, or the VAE's right / bottom padding (pad 0, stride 2)");
    // pad 1: symmetric (UNet).  pad 0: the window of output (oy, ox) starts at (2 oy, 2 ox) and the input is padded by one row /
    // column at the bottom / right only (diffusers Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then conv stride 2)
    const int Ho = pad ? (H + 2 - 3) / stride + 1 : (H + 1 - 3) / 2 + 1, Wo = pad ? (W + 2 - 3) / stride + 1 : (W + 1 - 3) / 2 + 1;
    S2I_REQ(ldcol >= 9L * C, "im2col3x3: ldcol too small");
    if ((C & 3) == 0 && (ldcol & 3) == 0 && (ldx & 3) == 0 && (long)B * Ho * Wo * (ldcol / 4) < (1L << 31))
        S2I_LAUNCH((im2col3x3_kernel), grid_for((long)B * Ho * Wo * (ldcol / 4), 256), 256, 0, st, x, ldx, B, H, W, C, stride, pad, Ho, Wo,
                   (__half*)col16, ldcol);
    else
        S2I_LAUNCH((im2col3x3_scalar_kernel), grid_for((long)B * Ho * Wo * ldcol, 256), 256, 0, st, x, ldx, B, H, W, C, stride, pad, Ho,
                   Wo, (__half*)col16, ldcol);
    S2I_LAUNCH_CHECK();
    return 0;
}

int nchw_to_nhwc(const float* src, int B, int C, int H, int W, float* dst, long ldn, cudaStream_t st) {
    S2I_LAUNCH((nchw_to_nhwc_kernel), grid_for((long)B * H * W * ldn, 256), 256, 0, st, src, B, C, H, W, dst, ldn);
    S2I_LAUNCH_CHECK();
    return 0;
}

int nhwc_to_nchw(const float* src, long ldn, int B, int C, int H, int W, float* dst, cudaStream_t st) {
    S2I_LAUNCH((nhwc_to_nchw_kernel), grid_for((long)B * C * H * W, 256), 256, 0, st, src, ldn, B, C, H, W, dst);
    S2I_LAUNCH_CHECK();
    return 0;
}

int gemv(const float* x, int K, const void* w16, const float* bias, int N, int silu_in, float* out, cudaStream_t st) {
    S2I_REQ((K & 7) == 0 && K <= 2048 && (reinterpret_cast<uintptr_t>(w16) & 15) == 0, "gemv: K must be a multiple of 8, at most 2048, and the weights 16-byte aligned");
    S2I_LAUNCH((gemv_kernel), ceil_div(N, 8), 256, 0, st, x, K, (const __half*)w16, bias, N, silu_in, out);
    S2I_LAUNCH_CHECK();
    return 0;
}

int timestep_embedding(float t, int dim, float* out, cudaStream_t st) {
    S2I_LAUNCH((timestep_embedding_kernel), 1, 256, 0, st, t, dim, out);
    S2I_LAUNCH_CHECK();
    return 0;
}

}  // namespace s2i
