// Fused attention backward on tcgen05 + TMA: dQ, dK, dV from (Q, K, V, O, dO, log-sum-exp) with the scores and
// probabilities RECOMPUTED on chip -- nothing of size Nq x Nk touches HBM.  Replaces, for the UNet's attention blocks,
// the part of `torch.autograd.grad(loss, latents_prev)` (modules/pipeline.py:159) that flows through
// softmax(scale Q K^T) V of diffusers' CrossAttention (app.py:43 / SURVEY A.4).
//
// One kernel, two modes; a CTA owns 128 "row" tokens and streams 64-token "column" tiles:
//   MODE_DQ   rows = queries (Q_i, dO_i resident), columns = keys   (K_j, V_j streamed):
//             S = Q K^T, dP = dO V^T, P = exp(scale S - lse_row), dS = scale P (dP - delta_row);  dQ_i += dS K_j
//   MODE_DKV  rows = keys    (K_j, V_j resident), columns = queries (Q_i, dO_i streamed):
//             S^T = K Q^T, dP^T = V dO^T, P^T = exp(scale S^T - lse_col), dS^T likewise;  dV_j += P^T dO_i, dK_j += dS^T Q_i
// Both products of a tile land in TMEM (2 x 64 fp32 columns each, double buffered); four softmax warps (one row per
// thread) turn them into fp16 P / dS tiles in 128B-swizzled shared memory, which the tensor core reads back as the
// A operand of the accumulating products (accumulators: TMEM columns 256.., fp32).
//   warp 0      TMA producer      warp 1  TMEM allocator + tcgen05.mma issuer      warps 2..5  softmax + epilogue
#include "attn.cuh"

#include <cudaTypedefs.h>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace s2i {

int encode_tmap_f16(CUtensorMap* m, int rank, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);   // gemm_tc.cu

namespace {

constexpr int kThreads = 192;
constexpr int kRows = 128;               // resident ("row") tokens per CTA
constexpr int kCols = 64;                // streamed ("column") tokens per tile
constexpr int kChunk16 = 128 * 64 * 2;   // 16 KB: 128 rows x 64 fp16 (one swizzle-128B K-major chunk)
constexpr int kChunk8 = 64 * 64 * 2;     //  8 KB:  64 rows x 64 fp16
constexpr int kMaxStages = 4;            // streamed-operand ring depth is chosen at launch (2 .. 4)

struct __align__(64) BwdParams {
    CUtensorMap mapR1, mapR2, mapC1, mapC2;   // resident (box 64 x 128) and streamed (box 64 x 64) operands
    int mode;                                 // 0 = dQ, 1 = dK/dV
    int Nrow, Ncol, heads, dp, nkc;
    int r1_c0, r2_c0, c1_c0, c2_c0;           // column of head 0 in each operand tensor
    int sbufs;                                // staging buffers per staged matrix (1 or 2)
    int stages;                               // streamed-operand ring depth
    int nT, tmem_cols;                        // tile-product buffers in TMEM (1 or 2) and the allocation that holds them
    uint32_t idesc_t, idesc_acc;
    float scale, scale_log2;
    const float* lse;                         // [B*heads][Nq]
    const float* delta;                       // [B*heads][Nq]: rowsum(dO * O); mode 1 reads what mode 0 wrote
    float* delta_w;                           // mode 0: where this launch writes it
    const __half* o_g; long ldo_g;            // mode 0: forward output and its gradient (global, head h at column h * dp)
    const __half* do_g; long lddo_g;
    int Nq;
    __half* out0; long ld0; int o0_c0;        // dQ (mode 0) or dV (mode 1)
    __half* out1; long ld1; int o1_c0;        // dK (mode 1)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kThreads, 2) attn_bwd_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int r_bytes = p.nkc * kChunk16;          // one resident operand tile
    const int c_bytes = p.nkc * kChunk8;           // one streamed operand tile
    const int stage_bytes = 2 * c_bytes;
    const int nstaged = p.mode ? 2 : 1;            // staged A operands per tile: dS (mode 0); P^T and dS^T (mode 1)
    uint8_t* sR1 = smem;
    uint8_t* sR2 = sR1 + r_bytes;
    uint8_t* sSt = sR2 + r_bytes;                  // [sbufs][nstaged] x 16 KB
    uint8_t* sC = sSt + p.sbufs * nstaged * kChunk16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sC + p.stages * stage_bytes);
    uint64_t* r_full = bars;            // resident operands landed
    uint64_t* acc_full = bars + 1;      // accumulators complete
    uint64_t* t_full = bars + 2;        // [2] both tile products complete
    uint64_t* t_empty = bars + 4;       // [2] softmax warps drained the TMEM tile buffers
    uint64_t* st_full = bars + 6;       // [2] staging written
    uint64_t* st_empty = bars + 8;      // [2] accumulating MMAs done with the staging buffer
    uint64_t* c_full = bars + 10;       // [kMaxStages]
    uint64_t* c_empty = bars + 14;      // [kMaxStages]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    float* sStat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bars + 20) + 15) & ~uintptr_t(15));   // mode 1: [2][2][64] (lse*log2e, delta)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * kRows;
    const int z = blockIdx.y;
    const int b = z / p.heads, h = z - b * p.heads;
    const int T = (p.Ncol + kCols - 1) / kCols;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.mapR1);
        ptx::prefetch_tmap(&p.mapR2);
        ptx::prefetch_tmap(&p.mapC1);
        ptx::prefetch_tmap(&p.mapC2);
    }
    if (warp == 1) {
        if (lane == 0) {
            ptx::mbar_init(r_full, 1);
            ptx::mbar_init(acc_full, 1);
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&t_full[i], 1);
                ptx::mbar_init(&t_empty[i], 4);
                ptx::mbar_init(&st_full[i], 4);
                ptx::mbar_init(&st_empty[i], 1);
            }
            for (int s = 0; s < p.stages; ++s) {
                ptx::mbar_init(&c_full[s], 1);
                ptx::mbar_init(&c_empty[s], 1);
            }
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();       // everything above is local setup; global memory of earlier kernels is touched only below
    pdl_launch();     // TMEM is held: dependents may become resident
    const uint32_t tmem_acc0 = tmem_base + (uint32_t)p.nT * 128u;    // T1 buffers | T2 buffers | accumulators
    const uint32_t tmem_acc1 = tmem_acc0 + (uint32_t)p.dp;

    if (warp == 0) {
        {
            // ------------------------------------------------ TMA producer (whole warp in the loop, one elected lane issues)
            const bool leader = ptx::elect_one();
            if (leader) {
                ptx::mbar_expect_tx(r_full, (uint32_t)(2 * r_bytes));
                for (int c = 0; c < p.nkc; ++c) {
                    ptx::tma_load_3d(sR1 + c * kChunk16, &p.mapR1, r_full, p.r1_c0 + h * p.dp + c * 64, row0, b);
                    ptx::tma_load_3d(sR2 + c * kChunk16, &p.mapR2, r_full, p.r2_c0 + h * p.dp + c * 64, row0, b);
                }
            }
            // ring counters advance incrementally (no division in the single-thread issue loops, cf. producer_loop in gemm_tc.cu)
            const int stages = p.stages, nkc = p.nkc;
            const int c1 = p.c1_c0 + h * p.dp, c2 = p.c2_c0 + h * p.dp;
            int stage = 0;
            uint32_t parity = 1;
            uint8_t* s1 = sC;
            for (int j = 0, col0 = 0; j < T; ++j, col0 += kCols) {
                ptx::mbar_wait(&c_empty[stage], parity);
                if (leader) {
                    ptx::mbar_expect_tx(&c_full[stage], (uint32_t)stage_bytes);
                    uint8_t* s2 = s1 + c_bytes;
                    for (int c = 0; c < nkc; ++c) {
                        ptx::tma_load_3d(s1 + c * kChunk8, &p.mapC1, &c_full[stage], c1 + c * 64, col0, b);
                        ptx::tma_load_3d(s2 + c * kChunk8, &p.mapC2, &c_full[stage], c2 + c * 64, col0, b);
                    }
                }
                s1 += stage_bytes;
                if (++stage == stages) {
                    stage = 0;
                    s1 = sC;
                    parity ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        {
            // ------------------------------------------------ MMA issuer (whole warp in the loop, one elected lane issues)
            const bool leader = ptx::elect_one();
            const int nks = p.dp >> 4;
            const int stages = p.stages, nT = p.nT, sbufs = p.sbufs, mode = p.mode;
            const uint32_t idesc_t = p.idesc_t, idesc_acc = p.idesc_acc;
            // descriptors of the first ring slots; the start-address field counts 16-byte units, so slots / K steps are additions
            const uint64_t dR1 = ptx::make_smem_desc_sw128(ptx::smem_u32(sR1), 16u, 1024u);
            const uint64_t dR2 = ptx::make_smem_desc_sw128(ptx::smem_u32(sR2), 16u, 1024u);
            const uint64_t dC1k = ptx::make_smem_desc_sw128(ptx::smem_u32(sC), 16u, 1024u);                        // K-major read (T products)
            const uint64_t dC1m = ptx::make_smem_desc_sw128(ptx::smem_u32(sC), (uint32_t)kChunk8, 1024u);          // MN-major read (accumulations)
            const uint64_t dSt0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sSt), 16u, 1024u);
            const uint64_t stage_step = (uint32_t)stage_bytes >> 4, c2_off = (uint32_t)c_bytes >> 4;
            const uint64_t st_step = (uint32_t)(nstaged * kChunk16) >> 4;
            int t_stage = 0, tb = 0;
            uint32_t t_c_par = 0, tb_par = 1;
            uint64_t t_coff = 0;
            // T1[g % nT] = R1 C1_g^T, T2[g % nT] = R2 C2_g^T
            auto issue_T = [&]() {
                ptx::mbar_wait(&c_full[t_stage], t_c_par);
                ptx::mbar_wait(&t_empty[tb], tb_par);
                ptx::tc_fence_after();
                const uint32_t t1 = tmem_base + (uint32_t)tb * 64u;
                const uint32_t t2 = tmem_base + (uint32_t)nT * 64u + (uint32_t)tb * 64u;
                if (leader) {
                    for (int k = 0; k < nks; ++k) {
                        const uint64_t ks = (uint64_t)((k & 3) * 2);
                        ptx::umma_f16(t1, dR1 + (uint64_t)((k >> 2) * (kChunk16 >> 4)) + ks,
                                      dC1k + t_coff + (uint64_t)((k >> 2) * (kChunk8 >> 4)) + ks, idesc_t, k != 0 ? 1u : 0u);
                    }
                    for (int k = 0; k < nks; ++k) {
                        const uint64_t ks = (uint64_t)((k & 3) * 2);
                        ptx::umma_f16(t2, dR2 + (uint64_t)((k >> 2) * (kChunk16 >> 4)) + ks,
                                      dC1k + t_coff + c2_off + (uint64_t)((k >> 2) * (kChunk8 >> 4)) + ks, idesc_t, k != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&t_full[tb]);
                }
                t_coff += stage_step;
                if (++t_stage == stages) {
                    t_stage = 0;
                    t_coff = 0;
                    t_c_par ^= 1u;
                }
                if (++tb == nT) {
                    tb = 0;
                    tb_par ^= 1u;
                }
            };
            ptx::mbar_wait(r_full, 0);
            issue_T();
            int stage = 0, sb = 0;
            uint32_t sb_par = 0, accumulate = 0;
            uint64_t coff = 0, stoff = 0;
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) issue_T();           // keep the softmax warps fed while tile j's staging is produced
                ptx::mbar_wait(&st_full[sb], sb_par);
                ptx::tc_fence_after();
                const uint64_t dSt = dSt0 + stoff, dC1 = dC1m + coff, dC2 = dC1 + c2_off;
                if (!leader) {
                } else if (mode == 0) {
                    // dQ += dS K_j   (B = K_j read MN-major: N = head channels, K = keys)
                    for (int kk = 0; kk < 4; ++kk)
                        ptx::umma_f16(tmem_acc0, dSt + (uint64_t)(kk * 2), dC1 + (uint64_t)(kk * 128), idesc_acc, accumulate | (uint32_t)(kk != 0));
                } else {
                    // dV += P^T dO_i ; dK += dS^T Q_i
                    for (int kk = 0; kk < 4; ++kk)
                        ptx::umma_f16(tmem_acc0, dSt + (uint64_t)(kk * 2), dC2 + (uint64_t)(kk * 128), idesc_acc, accumulate | (uint32_t)(kk != 0));
                    for (int kk = 0; kk < 4; ++kk)
                        ptx::umma_f16(tmem_acc1, dSt + (uint64_t)((kChunk16 >> 4) + kk * 2), dC1 + (uint64_t)(kk * 128), idesc_acc,
                                      accumulate | (uint32_t)(kk != 0));
                }
                accumulate = 1u;
                if (leader) {
                    ptx::umma_commit(&c_empty[stage]);
                    ptx::umma_commit(&st_empty[sb]);
                }
                coff += stage_step;
                if (++stage == stages) {
                    stage = 0;
                    coff = 0;
                }
                stoff += st_step;
                if (++sb == sbufs) {
                    sb = 0;
                    stoff = 0;
                    sb_par ^= 1u;
                }
            }
            if (leader) ptx::umma_commit(acc_full);
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------- softmax warps: thread <-> row token
        const int e = threadIdx.x - 64;
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const bool row_ok = (row0 + r) < p.Nrow;
        const float kLog2e = 1.4426950408889634f;
        float lse_r = 0.f, delta_r = 0.f;
        if (p.mode == 0 && row_ok) {
            lse_r = p.lse[(long)z * p.Nq + row0 + r] * kLog2e;
            // delta_row = <dO_row, O_row>, computed here (each thread owns one query row; the two 2 dp-byte rows are read
            // while the first operand tiles are still in flight) and left in global memory for the dK / dV launch
            const uint4* po = reinterpret_cast<const uint4*>(p.o_g + ((long)b * p.Nq + row0 + r) * p.ldo_g + h * p.dp);
            const uint4* pd = reinterpret_cast<const uint4*>(p.do_g + ((long)b * p.Nq + row0 + r) * p.lddo_g + h * p.dp);
            float acc = 0.f;
            for (int c = 0; c < p.dp / 8; ++c) {
                const uint4 a = __ldg(po + c), g = __ldg(pd + c);
                const __half2* ha = reinterpret_cast<const __half2*>(&a);
                const __half2* hg = reinterpret_cast<const __half2*>(&g);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 x = __half22float2(ha[i]), y = __half22float2(hg[i]);
                    acc = fmaf(x.x, y.x, acc);
                    acc = fmaf(x.y, y.y, acc);
                }
            }
            delta_r = acc;
            p.delta_w[(long)z * p.Nq + row0 + r] = acc;
        }
        const int nT = p.nT, sbufs = p.sbufs;
        int tb = 0, sb = 0;
        uint32_t tb_par = 0, sb_par = 1;
        for (int j = 0; j < T; ++j) {
            const int col0 = j * kCols;
            if (p.mode == 1) {
                // per-column statistics of this query tile -> shared memory (double buffered by tile parity)
                float* st = sStat + (j & 1) * 128;
                const int c = e & 63;
                const bool ok = (col0 + c) < p.Ncol;
                const long idx = (long)z * p.Nq + col0 + c;
                st[e] = ok ? (e < 64 ? p.lse[idx] * kLog2e : p.delta[idx]) : 0.f;
                ptx::named_bar_sync(1, 128);
            }
            ptx::mbar_wait(&t_full[tb], tb_par);
            ptx::tc_fence_after();
            const uint32_t t1 = tmem_base + (uint32_t)tb * 64u + lane_addr;
            const uint32_t t2 = t1 + (uint32_t)nT * 64u;
            const int nvalid = p.Ncol - col0;          // columns >= nvalid are padding
            const float* stl = sStat + (j & 1) * 128;
            ptx::mbar_wait(&st_empty[sb], sb_par);   // the MMAs that read this staging buffer are done
            // K-major 128B-swizzled tile: 16-byte unit u of row r lives at r * 128 + ((u ^ (r & 7)) * 16)
            uint8_t* st0 = sSt + sb * nstaged * kChunk16 + r * 128;
#pragma unroll
            for (int c0 = 0; c0 < kCols; c0 += 32) {
                uint32_t s[32], d[32];
                ptx::tmem_ld_32x32(t1 + (uint32_t)c0, s);
                ptx::tmem_ld_32x32(t2 + (uint32_t)c0, d);
                ptx::tmem_ld_wait();
                if (c0 + 32 == kCols) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&t_empty[tb]);      // both tile products are in registers
                }
                uint32_t pk[16], dk[16];
                // issue-bound loop (one warp per scheduler): 4 elements per iteration, column statistics (mode 1) fetched
                // as two 128-bit shared loads, padding handled only in the last tile
                const bool ragged = (c0 + 32) > nvalid;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 l4 = make_float4(lse_r, lse_r, lse_r, lse_r), d4 = make_float4(delta_r, delta_r, delta_r, delta_r);
                    if (p.mode == 1) {
                        l4 = *reinterpret_cast<const float4*>(stl + c0 + i);
                        d4 = *reinterpret_cast<const float4*>(stl + 64 + c0 + i);
                    }
                    float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), p.scale_log2, -l4.x));
                    float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -l4.y));
                    float p2 = ex2_approx(fmaf(__uint_as_float(s[i + 2]), p.scale_log2, -l4.z));
                    float p3 = ex2_approx(fmaf(__uint_as_float(s[i + 3]), p.scale_log2, -l4.w));
                    if (ragged) {
                        if (c0 + i >= nvalid) p0 = 0.f;
                        if (c0 + i + 1 >= nvalid) p1 = 0.f;
                        if (c0 + i + 2 >= nvalid) p2 = 0.f;
                        if (c0 + i + 3 >= nvalid) p3 = 0.f;
                    }
                    const float ps0 = p0 * p.scale, ps1 = p1 * p.scale, ps2 = p2 * p.scale, ps3 = p3 * p.scale;
                    const float g0 = ps0 * (__uint_as_float(d[i]) - d4.x);
                    const float g1 = ps1 * (__uint_as_float(d[i + 1]) - d4.y);
                    const float g2 = ps2 * (__uint_as_float(d[i + 2]) - d4.z);
                    const float g3 = ps3 * (__uint_as_float(d[i + 3]) - d4.w);
                    const __half2 pa = __floats2half2_rn(p0, p1), pb = __floats2half2_rn(p2, p3);
                    const __half2 ga = __floats2half2_rn(g0, g1), gb = __floats2half2_rn(g2, g3);
                    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pa);
                    pk[(i >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&pb);
                    dk[i >> 1] = *reinterpret_cast<const uint32_t*>(&ga);
                    dk[(i >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&gb);
                }
                const int u0 = c0 >> 3;      // first 16-byte unit of this half (8 fp16 per unit)
                if (p.mode == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        *reinterpret_cast<uint4*>(st0 + (((u0 + u) ^ (r & 7)) << 4)) =
                            make_uint4(dk[u * 4], dk[u * 4 + 1], dk[u * 4 + 2], dk[u * 4 + 3]);
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        *reinterpret_cast<uint4*>(st0 + (((u0 + u) ^ (r & 7)) << 4)) =
                            make_uint4(pk[u * 4], pk[u * 4 + 1], pk[u * 4 + 2], pk[u * 4 + 3]);
                        *reinterpret_cast<uint4*>(st0 + kChunk16 + (((u0 + u) ^ (r & 7)) << 4)) =
                            make_uint4(dk[u * 4], dk[u * 4 + 1], dk[u * 4 + 2], dk[u * 4 + 3]);
                    }
                }
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&st_full[sb]);
            if (++tb == nT) {
                tb = 0;
                tb_par ^= 1u;
            }
            if (++sb == sbufs) {
                sb = 0;
                sb_par ^= 1u;
            }
        }
        // epilogue: accumulators -> fp16
        ptx::mbar_wait(acc_full, 0);
        ptx::tc_fence_after();
        const int nout = p.mode ? 2 : 1;
        for (int o = 0; o < nout; ++o) {
            __half* base = o == 0 ? p.out0 : p.out1;
            const long ld = o == 0 ? p.ld0 : p.ld1;
            const int c0 = o == 0 ? p.o0_c0 : p.o1_c0;
            __half* orow = base + ((long)b * p.Nrow + row0 + r) * ld + c0 + h * p.dp;
            const uint32_t tacc = (o == 0 ? tmem_acc0 : tmem_acc1) + lane_addr;
            for (int c = 0; c < p.dp; c += 16) {
                uint32_t raw[16];
                ptx::tmem_ld_32x16(tacc + (uint32_t)c, raw);
                ptx::tmem_ld_wait();
                if (row_ok) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const __half2 h2 = __floats2half2_rn(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
                        w[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                    }
                    *reinterpret_cast<uint4*>(orow + c) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

int make_map(CUtensorMap* m, const __half* ptr, long ld, int N, int B, int box_rows) {
    uint64_t dims[3] = {(uint64_t)ld, (uint64_t)N, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)N * ld * 2};
    uint32_t box[3] = {64, (uint32_t)box_rows, 1};
    return encode_tmap_f16(m, 3, ptr, dims, str, box);
}

}  // namespace

bool attn_bwd_supported(int Nq, int Nk, int dp) {
    // TMEM: the tile products (128 columns single-, 256 double-buffered) + dp (dQ) or 2 dp (dK | dV) accumulator columns must
    // fit 512; shared memory: 2 resident + 2 x 2 streamed operand tiles of ceil(dp / 64) chunks + staging must fit 227 KB.
    // dp = 160 (SD1.5's 16 x 16 level) runs the dK / dV mode single-buffered in 448 columns.
    return Nq >= kRows && Nk >= kCols && dp % 16 == 0 && dp >= 16 && dp <= 192;
}

int attn_bwd_launch(const AttnBwdDesc& d, cudaStream_t stream) {
    if (!attn_bwd_supported(d.Nq, d.Nk, d.dp)) return set_error(S2I_ERR_ARG, "attn_bwd: unsupported shape Nq=%d Nk=%d dp=%d", d.Nq, d.Nk, d.dp);
    if (!d.q || !d.kv || !d.dO || !d.o || !d.lse || !d.delta || !d.dq) return set_error(S2I_ERR_ARG, "attn_bwd: null argument");
    const int Z = d.B * d.heads;
    if ((d.ldo % 8) != 0 || (d.lddo % 8) != 0 || (d.dp % 8) != 0 || (reinterpret_cast<uintptr_t>(d.o) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(d.dO) & 15) != 0)
        return set_error(S2I_ERR_ARG, "attn_bwd: O / dO rows must be 16-byte aligned per head");
    static bool attr_set = false;
    if (!attr_set) {
        S2I_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int nkc = (d.dp + 63) / 64;
    for (int mode = 0; mode < 2; ++mode) {
        if (mode == 1 && !d.dk) break;
        BwdParams p;
        memset(&p, 0, sizeof(p));
        p.mode = mode;
        p.heads = d.heads; p.dp = d.dp; p.nkc = nkc; p.Nq = d.Nq;
        p.scale = d.scale;
        p.scale_log2 = d.scale * 1.4426950408889634f;
        p.lse = d.lse; p.delta = d.delta; p.delta_w = d.delta;        // delta = rowsum(dO * O): written by the dQ launch
        p.o_g = d.o; p.ldo_g = d.ldo; p.do_g = d.dO; p.lddo_g = d.lddo;
        p.idesc_t = ptx::make_idesc_f16(128, kCols, 0, 0, 0);
        p.idesc_acc = ptx::make_idesc_f16(128, (uint32_t)d.dp, 0, 0, 1);
        if (mode == 0) {
            p.Nrow = d.Nq; p.Ncol = d.Nk;
            S2I_TRY(make_map(&p.mapR1, d.q, d.ldq, d.Nq, d.B, kRows));   p.r1_c0 = d.q_c0;
            S2I_TRY(make_map(&p.mapR2, d.dO, d.lddo, d.Nq, d.B, kRows)); p.r2_c0 = 0;
            S2I_TRY(make_map(&p.mapC1, d.kv, d.ldkv, d.Nk, d.B, kCols)); p.c1_c0 = d.k_c0;
            S2I_TRY(make_map(&p.mapC2, d.kv, d.ldkv, d.Nk, d.B, kCols)); p.c2_c0 = d.v_c0;
            p.out0 = d.dq; p.ld0 = d.lddq; p.o0_c0 = d.dq_c0;
        } else {
            p.Nrow = d.Nk; p.Ncol = d.Nq;
            S2I_TRY(make_map(&p.mapR1, d.kv, d.ldkv, d.Nk, d.B, kRows)); p.r1_c0 = d.k_c0;
            S2I_TRY(make_map(&p.mapR2, d.kv, d.ldkv, d.Nk, d.B, kRows)); p.r2_c0 = d.v_c0;
            S2I_TRY(make_map(&p.mapC1, d.q, d.ldq, d.Nq, d.B, kCols));   p.c1_c0 = d.q_c0;
            S2I_TRY(make_map(&p.mapC2, d.dO, d.lddo, d.Nq, d.B, kCols)); p.c2_c0 = 0;
            p.out0 = d.dv; p.ld0 = d.lddkv; p.o0_c0 = d.dv_c0;
            p.out1 = d.dk; p.ld1 = d.lddkv; p.o1_c0 = d.dk_c0;
        }
        const int nstaged = mode ? 2 : 1;
        // TMEM: double-buffered tile products need 256 + nacc*dp columns (512 allocation, one CTA per SM); single-buffered
        // they fit 256 columns and two CTAs share the SM -- the softmax warps are issue-bound, so a second CTA's warps
        // on every scheduler are worth more than overlapping this CTA's own tile products
        const int nacc = mode ? 2 : 1;
        p.nT = (128 + nacc * d.dp <= 256 || 256 + nacc * d.dp > 512) ? 1 : 2;
        p.tmem_cols = (p.nT * 128 + nacc * d.dp <= 256) ? 256 : 512;
        if (p.nT * 128 + nacc * d.dp > 512) return set_error(S2I_ERR_ARG, "attn_bwd: head dim %d does not fit TMEM", d.dp);
        const int stage_b = 2 * nkc * kChunk8;
        const int fixed = 2 * nkc * kChunk16 + 2 * stage_b + 2048 + 1024;   // operands (2-stage ring) + barriers/stats + slack
        p.sbufs = (fixed + 2 * nstaged * kChunk16 <= 227 * 1024) ? 2 : 1;
        // prefer two co-resident CTAs when a single staging buffer makes the footprint fit half an SM
        if (fixed + 2 * nstaged * kChunk16 > 113 * 1024 && fixed + nstaged * kChunk16 <= 113 * 1024) p.sbufs = 1;
        size_t smem_bytes = (size_t)fixed + (size_t)p.sbufs * nstaged * kChunk16;
        if (smem_bytes > 227u * 1024u) return set_error(S2I_ERR_ARG, "attn_bwd: head dim %d does not fit shared memory", d.dp);
        // A streamed tile's stage is only reloaded after the accumulating products that read it, one softmax pass later:
        // deepen the ring as far as the footprint's occupancy class (two CTAs per SM, or one) allows, like the forward's
        // "score tiles at most stages - 2 ahead" rule.  S2I_ATTNB_STAGES caps it (tools A/B).
        const size_t budget = smem_bytes <= 113u * 1024u ? 113u * 1024u : 227u * 1024u;
        int cap = kMaxStages;
        if (const char* e = getenv("S2I_ATTNB_STAGES")) cap = atoi(e) < 2 ? 2 : (atoi(e) > kMaxStages ? kMaxStages : atoi(e));
        p.stages = 2;
        const int Tcols = (p.Ncol + kCols - 1) / kCols;
        while (p.stages < cap && p.stages < Tcols && smem_bytes + (size_t)stage_b <= budget) {
            ++p.stages;
            smem_bytes += (size_t)stage_b;
        }
        dim3 grid((unsigned)((p.Nrow + kRows - 1) / kRows), (unsigned)Z, 1);
        S2I_LAUNCH((attn_bwd_kernel), grid, kThreads, smem_bytes, stream, p);
        // algorithmic work of the reference's backward: dP, dQ (mode 0) and dV, dK (mode 1) products at the true head dim
        S2I_LAUNCH_CHECK_TAG("attn_bwd", 4.0 * Z * (double)d.Nq * d.Nk * d.d_true,
                             2.0 * Z * d.d_true * (3.0 * d.Nq + 3.0 * d.Nk));      // Q, dO, dQ | K, V, dK/dV once (fp16)
    }
    return 0;
}

}  // namespace s2i
