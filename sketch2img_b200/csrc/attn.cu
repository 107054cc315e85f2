// Fused attention forward  O = softmax(scale * Q K^T) V  on tcgen05 + TMA (reference: the xformers /
// diffusers CrossAttention call inside `self.unet(...)`, modules/pipeline.py:96, app.py:43).
//
// One CTA = one (batch, head, 128-query tile).  S = Q K_j^T lands in TMEM (128 lanes x 128 fp32 columns); four
// softmax warps (one query row per thread) read it with tcgen05.ld, exponentiate, and write P as fp16 into shared
// memory in the 128B-swizzled K-major layout the tensor core reads; O += P V_j accumulates in TMEM (V is the
// MN-major B operand straight from its [tokens][channels] layout).  The scores never touch HBM.
//
// One pass over the key tiles with an online softmax whose reference maximum moves lazily: probabilities are formed as
// p = exp2(s * scale * log2 e - m_ref); m_ref is only raised (and the TMEM accumulator and row sum rescaled by
// exp2(m_old - m_new)) when a tile's maximum exceeds it by more than 8 (p stays <= 2^8, exact in fp16 / fp32).  After the
// first tiles this almost never happens, so the accumulator is free of per-tile rescaling.  S tiles are issued up to
// nS - 1 tiles ahead of the softmax warps (TMEM: nS x 64 score columns + dp output columns).  Final O / rowsum -> fp16.
//
//   warp 0      TMA producer (Q once, then K+V tiles through a stage ring)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  softmax + epilogue (TMEM lane quarter = warp & 3)
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
constexpr int kTileQ = 128;
constexpr int kTileK = 64;               // keys per tile: one 64-row TMA box serves as K (K-major) and V (MN-major)
constexpr int kChunk16 = 128 * 64 * 2;   // 16 KB: 128 rows x 64 fp16 (one swizzle-128B K-major chunk of Q or P)
constexpr int kChunk8 = 64 * 64 * 2;     // 8 KB: 64 rows x 64 fp16 (one chunk of a K or V tile)
constexpr int kMaxStages = 8;
constexpr int kMaxS = 4;

struct __align__(64) AttnParams {
    CUtensorMap mapQ, mapKV;
    int Nq, Nk, heads, dp;
    int q_c0, k_c0, v_c0;
    int nkc;            // 64-wide chunks covering dp
    int stages, pbufs, tmem_cols, nS;
    uint32_t idesc_s, idesc_o;
    float scale_log2;   // softmax scale * log2(e)
    float scale;
    __half* out;
    long ldo;
    float* lse;         // optional [B*heads][Nq]: scale * max + ln(rowsum)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) attn_fwd_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int q_bytes = p.nkc * kChunk16;
    const int k_bytes = p.nkc * kChunk8;
    const int stage_bytes = 2 * k_bytes;          // K tile + V tile
    uint8_t* sQ = smem;
    uint8_t* sP = sQ + q_bytes;                   // pbufs x 16 KB
    uint8_t* sKV = sP + p.pbufs * kChunk16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + p.stages * stage_bytes);
    uint64_t* q_full = bars;
    uint64_t* o_full = bars + 1;
    uint64_t* p_full = bars + 2;     // [2]
    uint64_t* p_empty = bars + 4;    // [2]
    uint64_t* s_full = bars + 6;     // [kMaxS]
    uint64_t* s_empty = bars + 10;   // [kMaxS]
    uint64_t* kv_full = bars + 14;   // [kMaxStages]
    uint64_t* kv_empty = kv_full + kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + kMaxStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kTileQ;
    const int z = blockIdx.y;
    const int b = z / p.heads, h = z - b * p.heads;
    const int T = (p.Nk + kTileK - 1) / kTileK;      // key tiles

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.mapQ);
        ptx::prefetch_tmap(&p.mapKV);
    }
    if (warp == 1) {
        if (lane == 0) {
            ptx::mbar_init(q_full, 1);
            ptx::mbar_init(o_full, 1);
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(&p_full[i], 4);
                ptx::mbar_init(&p_empty[i], 1);
            }
            for (int i = 0; i < kMaxS; ++i) {
                ptx::mbar_init(&s_full[i], 1);
                ptx::mbar_init(&s_empty[i], 4);
            }
            for (int s = 0; s < p.stages; ++s) {
                ptx::mbar_init(&kv_full[s], 1);
                ptx::mbar_init(&kv_empty[s], 1);
            }
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, p.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();       // everything above is local setup; global memory of earlier kernels is touched only below
    pdl_launch();     // TMEM is held: dependents may become resident
    const uint32_t tmem_O = tmem_base + (uint32_t)p.nS * 64u;        // S buffers: nS x 64 columns, then O

    if (warp == 0) {
        {
            // ------------------------------------------------ TMA producer: the whole warp runs the (warp-uniform) loop, one
            // elected lane issues -- a single-lane loop keeps its state in per-thread registers and converts it for every copy
            const bool leader = ptx::elect_one();
            if (leader) {
                ptx::mbar_expect_tx(q_full, (uint32_t)q_bytes);
                for (int c = 0; c < p.nkc; ++c)
                    ptx::tma_load_3d(sQ + c * kChunk16, &p.mapQ, q_full, p.q_c0 + h * p.dp + c * 64, q0, b);
            }
            // (ring counters advance incrementally: a division per tile in these single-thread loops is a few hundred cycles
            // of dependent latency, see producer_loop in gemm_tc.cu)
            const int stages = p.stages, nkc = p.nkc;
            const int kc0 = p.k_c0 + h * p.dp, vc0 = p.v_c0 + h * p.dp;
            int stage = 0;
            uint32_t parity = 1;
            uint8_t* sK = sKV;
            for (int j = 0, key0 = 0; j < T; ++j, key0 += kTileK) {
                ptx::mbar_wait(&kv_empty[stage], parity);
                if (leader) {
                    ptx::mbar_expect_tx(&kv_full[stage], (uint32_t)stage_bytes);
                    uint8_t* sV = sK + k_bytes;
                    for (int c = 0; c < nkc; ++c) {
                        ptx::tma_load_3d(sK + c * kChunk8, &p.mapKV, &kv_full[stage], kc0 + c * 64, key0, b);
                        ptx::tma_load_3d(sV + c * kChunk8, &p.mapKV, &kv_full[stage], vc0 + c * 64, key0, b);
                    }
                }
                sK += stage_bytes;
                if (++stage == stages) {
                    stage = 0;
                    sK = sKV;
                    parity ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        {
            // ------------------------------------------------ MMA issuer (whole warp in the loop, one elected lane issues)
            const bool leader = ptx::elect_one();
            const int nks = p.dp >> 4;             // K steps of the QK^T product
            const int stages = p.stages, nS = p.nS, pbufs = p.pbufs;
            const uint32_t idesc_s = p.idesc_s, idesc_o = p.idesc_o;
            // descriptors of the first ring slots; the start-address field counts 16-byte units, so slots / K steps are additions
            const uint64_t dQ = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ), 16u, 1024u);
            const uint64_t dK0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sKV), 16u, 1024u);
            const uint64_t dV0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sKV + k_bytes), (uint32_t)kChunk8, 1024u);
            const uint64_t dP0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sP), 16u, 1024u);
            const uint64_t stage_step = (uint32_t)stage_bytes >> 4;
            // score-tile issue state: K ring slot + parity, S buffer + parity
            int s_stage = 0, s_buf = 0;
            uint32_t s_kv_par = 0, s_buf_par = 1;
            uint64_t s_koff = 0;
            // S[g % nS] = Q K_g^T once the tile has landed and the softmax warps have drained that S buffer
            auto issue_S = [&]() {
                ptx::mbar_wait(&kv_full[s_stage], s_kv_par);
                ptx::mbar_wait(&s_empty[s_buf], s_buf_par);
                ptx::tc_fence_after();
                const uint32_t tS = tmem_base + (uint32_t)s_buf * 64u;
                if (leader) {
                    for (int k = 0; k < nks; ++k) {
                        // K step k: 64-column chunk k >> 2 (16 KB apart in Q, 8 KB in K), 32 bytes per step inside it
                        const uint64_t ks = (uint64_t)((k & 3) * 2);
                        ptx::umma_f16(tS, dQ + (uint64_t)((k >> 2) * (kChunk16 >> 4)) + ks,
                                      dK0 + s_koff + (uint64_t)((k >> 2) * (kChunk8 >> 4)) + ks, idesc_s, k != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&s_full[s_buf]);
                }
                s_koff += stage_step;
                if (++s_stage == stages) {
                    s_stage = 0;
                    s_koff = 0;
                    s_kv_par ^= 1u;
                }
                if (++s_buf == nS) {
                    s_buf = 0;
                    s_buf_par ^= 1u;
                }
            };
            ptx::mbar_wait(q_full, 0);
            int next_s = 0;
            int stage = 0, pb = 0;
            uint32_t pb_par = 0;
            uint64_t voff = 0, poff = 0;
            uint32_t accumulate = 0;
            for (int j = 0; j < T; ++j) {
                // keep score tiles issued up to nS tiles ahead of the P V products: buffer (j + nS) % nS is the one the
                // softmax warps drained for tile j (before P_j exists), and its K tile's stage was freed by P V (j + nS -
                // stages) <= j - 1 because nS <= stages - 1 -- no wait below depends on a later step of this loop
                while (next_s < T && next_s <= j + nS) {
                    issue_S();
                    ++next_s;
                }
                ptx::mbar_wait(&p_full[pb], pb_par);
                ptx::tc_fence_after();
                if (leader) {
                    for (int kk = 0; kk < 4; ++kk)
                        ptx::umma_f16(tmem_O, dP0 + poff + (uint64_t)(kk * 2), dV0 + voff + (uint64_t)(kk * (2048 >> 4)), idesc_o,
                                      accumulate | (uint32_t)(kk != 0));
                    ptx::umma_commit(&kv_empty[stage]);
                    ptx::umma_commit(&p_empty[pb]);
                }
                accumulate = 1u;
                voff += stage_step;
                if (++stage == stages) {
                    stage = 0;
                    voff = 0;
                }
                poff += (uint64_t)(kChunk16 >> 4);
                if (++pb == pbufs) {
                    pb = 0;
                    poff = 0;
                    pb_par ^= 1u;
                }
            }
            if (leader) ptx::umma_commit(o_full);
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------- softmax + epilogue (one query row per thread)
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const bool row_ok = (q0 + r) < p.Nq;
        float mref = 0.f;       // reference maximum of this row, in log2 units (score * scale * log2 e)
        float l = 0.f;          // sum of p relative to mref
        const int nS = p.nS, pbufs = p.pbufs;
        const float scale_log2 = p.scale_log2;
        // ring state (no divisions in the per-tile chain): S buffer + parity, P buffer + parity, and the previous tile's P buffer
        int sbuf = 0, pb = 0, prev_pb = 0;
        uint32_t s_par = 0, pb_par = 1, prev_par = 0;       // pb_par: parity of p_empty to wait for before writing P_j
        for (int j = 0; j < T; ++j) {
            ptx::mbar_wait(&s_full[sbuf], s_par);
            ptx::tc_fence_after();
            const uint32_t tS = tmem_base + (uint32_t)sbuf * 64u + lane_addr;
            uint32_t s[kTileK];
#pragma unroll
            for (int c = 0; c < kTileK; c += 32) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tS + (uint32_t)c, raw);
#pragma unroll
                for (int i = 0; i < 32; ++i) s[c + i] = raw[i];
            }
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s_empty[sbuf]);     // S is in registers: TMEM buffer may be overwritten
            if (++sbuf == nS) {
                sbuf = 0;
                s_par ^= 1u;
            }

            const int kvalid = p.Nk - j * kTileK;                // columns >= kvalid are padding (last tile only)
            float mx = -INFINITY;
            if (kvalid >= kTileK) {
#pragma unroll
                for (int i = 0; i < kTileK; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
            } else {
#pragma unroll
                for (int i = 0; i < kTileK; ++i)
                    if (i < kvalid) mx = fmaxf(mx, __uint_as_float(s[i]));
            }
            const float mt = mx * scale_log2;
            // lazy reference update: only when this tile would push p above 2^8 (always on the first tile)
            const bool need = (j == 0) || (mt > mref + 8.f);
            if (__any_sync(0xffffffffu, need)) {
                const float mnew = need ? mt : mref;
                const float factor = (j == 0) ? 0.f : ex2_approx(mref - mnew);     // 1 for rows that keep their reference
                l *= factor;
                mref = mnew;
                if (j > 0) {
                    // every P V product issued so far must have landed before the accumulator rows are rescaled
                    ptx::mbar_wait(&p_empty[prev_pb], prev_par);
                    ptx::tc_fence_after();
                    for (int c = 0; c < p.dp; c += 16) {
                        uint32_t o[16];
                        ptx::tmem_ld_32x16(tmem_O + lane_addr + (uint32_t)c, o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                        ptx::tmem_st_32x16(tmem_O + lane_addr + (uint32_t)c, o);
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                }
            }
            const float mneg = -mref;
            uint32_t pk[kTileK / 2];
            // The softmax warps are issue-bound (one warp per scheduler): keep the per-element work at FFMA + EX2 + half
            // a pack + one add.  Padding columns only exist in the last tile; four independent partial row sums.
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
            if (kvalid >= kTileK) {
#pragma unroll
                for (int i = 0; i < kTileK; i += 4) {
                    const float e0 = ex2_approx(fmaf(__uint_as_float(s[i]), scale_log2, mneg));
                    const float e1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), scale_log2, mneg));
                    const float e2 = ex2_approx(fmaf(__uint_as_float(s[i + 2]), scale_log2, mneg));
                    const float e3 = ex2_approx(fmaf(__uint_as_float(s[i + 3]), scale_log2, mneg));
                    l0 += e0; l1 += e1; l2 += e2; l3 += e3;
                    const __half2 ha = __floats2half2_rn(e0, e1), hb = __floats2half2_rn(e2, e3);
                    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&ha);
                    pk[(i >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&hb);
                }
            } else {
#pragma unroll
                for (int i = 0; i < kTileK; i += 2) {
                    float e0 = ex2_approx(fmaf(__uint_as_float(s[i]), scale_log2, mneg));
                    float e1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), scale_log2, mneg));
                    if (i >= kvalid) e0 = 0.f;
                    if (i + 1 >= kvalid) e1 = 0.f;
                    l0 += e0; l1 += e1;
                    const __half2 h2 = __floats2half2_rn(e0, e1);
                    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                }
            }
            l += (l0 + l1) + (l2 + l3);
            ptx::mbar_wait(&p_empty[pb], pb_par);    // the P V product that read this buffer is done
            // row r of the K-major swizzled tile: 16-byte unit u lives at r * 128 + ((u ^ (r & 7)) * 16)
            uint8_t* prow = sP + pb * kChunk16 + r * 128;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                *reinterpret_cast<uint4*>(prow + ((u ^ (r & 7)) << 4)) =
                    make_uint4(pk[u * 4], pk[u * 4 + 1], pk[u * 4 + 2], pk[u * 4 + 3]);
            ptx::fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[pb]);
            // P_j's product completes phase (j / pbufs) of p_empty[pb]: what the next tile's rescale waits for
            prev_pb = pb;
            prev_par = pb_par ^ 1u;
            if (++pb == pbufs) {
                pb = 0;
                pb_par ^= 1u;
            }
        }
        // epilogue: O / rowsum -> fp16
        ptx::mbar_wait(o_full, 0);
        ptx::tc_fence_after();
        const float inv = l > 0.f ? 1.f / l : 0.f;
        __half* orow = p.out + ((long)b * p.Nq + q0 + r) * p.ldo + h * p.dp;
        for (int c = 0; c < p.dp; c += 16) {
            uint32_t raw[16];
            ptx::tmem_ld_32x16(tmem_O + lane_addr + (uint32_t)c, raw);
            ptx::tmem_ld_wait();
            if (row_ok) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const __half2 h2 = __floats2half2_rn(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv);
                    o[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                *reinterpret_cast<uint4*>(orow + c) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        if (p.lse && row_ok) p.lse[(long)z * p.Nq + q0 + r] = mref * 0.6931471805599453f + logf(l);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace

bool attn_fwd_supported(int Nq, int Nk, int dp) {
    return Nq >= kTileQ && Nk >= kTileK && dp % 16 == 0 && dp >= 16 && dp <= 256;   // TMA boxes never exceed the tensor
}

int attn_fwd_launch(const AttnDesc& d, cudaStream_t stream) {
    if (!attn_fwd_supported(d.Nq, d.Nk, d.dp)) return set_error(S2I_ERR_ARG, "attn_fwd: unsupported shape Nq=%d Nk=%d dp=%d", d.Nq, d.Nk, d.dp);
    if ((d.ldo % 8) != 0 || (reinterpret_cast<uintptr_t>(d.out) & 15) != 0 || (d.dp % 8) != 0)
        return set_error(S2I_ERR_ARG, "attn_fwd: output must be 16-byte aligned per head");
    AttnParams p;
    memset(&p, 0, sizeof(p));
    p.Nq = d.Nq; p.Nk = d.Nk; p.heads = d.heads; p.dp = d.dp;
    p.q_c0 = d.q_c0; p.k_c0 = d.k_c0; p.v_c0 = d.v_c0;
    p.nkc = (d.dp + 63) / 64;
    p.scale = d.scale;
    p.scale_log2 = d.scale * 1.4426950408889634f;
    p.out = d.out; p.ldo = d.ldo; p.lse = d.lse;
    p.idesc_s = ptx::make_idesc_f16(128, kTileK, 0, 0, 0);
    p.idesc_o = ptx::make_idesc_f16(128, (uint32_t)d.dp, 0, 0, 1);
    // Two configurations.  Long key sequences (self-attention): ONE CTA per SM with a deep K/V ring -- a stage is only
    // recycled after its P V product, a full load -> S -> softmax -> P V round trip (~4 k cycles) later, so keeping the
    // softmax warps fed at ~600 cycles per tile takes 6+ stages -- and up to four score tiles in flight (512 TMEM columns).
    // Short ones (cross-attention to 77 tokens): the small footprint that lets two CTAs share an SM.
    const int q_bytes = p.nkc * kChunk16, stage_bytes = 2 * p.nkc * kChunk8;
    const int overhead = 512 + 1024;     // barriers + alignment slack
    const int half_sm = 113 * 1024, full_sm = 226 * 1024;
    const int T = (d.Nk + kTileK - 1) / kTileK;
    static const int deep_mode = [] {
        const char* e = getenv("S2I_ATTN_DEEP");
        return e ? atoi(e) : 0;      // measured on B200: two co-resident shallow CTAs beat one deep CTA (212 vs 353 us at N=4096)
    }();
    static const int triple_mode = [] {
        const char* e = getenv("S2I_ATTN_TRIPLE");
        return e ? atoi(e) : -1;      // -1: automatic
    }();
    int pbufs = 0, stages = 0;
    bool triple = false;
    // measured (tools/attn_bench.py): three CTAs per SM win for 64-wide heads (N = 9216: 530 vs 728 us), lose for 48 (234 vs 220)
    const bool want_triple = triple_mode > 0 || (triple_mode < 0 && d.dp > 48 && T >= 32);
    if (want_triple && T >= 8 && d.dp <= 64 && q_bytes + kChunk16 + 2 * stage_bytes + overhead <= 75 * 1024) {
        // three CTAs per SM (128 TMEM columns each: one score tile + O): three softmax warps per scheduler
        triple = true;
        pbufs = 1;
        stages = 2;
        p.tmem_cols = 128;
    } else if (deep_mode && T >= 8) {
        pbufs = 2;
        stages = (full_sm - overhead - q_bytes - pbufs * kChunk16) / stage_bytes;
        if (stages > kMaxStages) stages = kMaxStages;
        if (stages > T) stages = T;
        p.tmem_cols = 512;
    } else {
        p.tmem_cols = d.dp <= 128 ? 256 : 512;      // 256 columns let two CTAs share an SM
        const int tries[4][2] = {{2, 3}, {2, 2}, {1, 2}, {0, 0}};
        for (int i = 0; tries[i][0]; ++i) {
            if (q_bytes + tries[i][0] * kChunk16 + tries[i][1] * stage_bytes + overhead <= half_sm) {
                pbufs = tries[i][0];
                stages = tries[i][1];
                break;
            }
        }
    }
    if (!pbufs) {
        pbufs = 2;
        stages = (full_sm - overhead - q_bytes - pbufs * kChunk16) / stage_bytes;
        if (stages > kMaxStages) stages = kMaxStages;
    }
    int ns_force = 0;
    if (const char* e = getenv("S2I_ATTN_CFG")) {       // tools/attn_bench.py --sweep: "pbufs,stages[,nS]" (two-CTA form only)
        int fp = 0, fs = 0, fn = 0;
        sscanf(e, "%d,%d,%d", &fp, &fs, &fn);
        if (!triple && fp >= 1 && fp <= 2 && fs >= 2 && fs <= kMaxStages &&
            q_bytes + fp * kChunk16 + fs * stage_bytes + overhead <= full_sm) {
            pbufs = fp;
            stages = fs;
            ns_force = fn;
        }
    }
    if (stages < 2) return set_error(S2I_ERR_ARG, "attn_fwd: head dim %d does not fit shared memory", d.dp);
    p.stages = stages;
    p.pbufs = pbufs;
    p.nS = (p.tmem_cols - d.dp) / 64;
    if (p.nS > kMaxS) p.nS = kMaxS;
    if (p.nS > stages - 1) p.nS = stages - 1;   // score tiles run nS key tiles ahead of the P V products (see the issuer)
    // measured (tools/attn_bench.py --sweep): a K/V stage is only reloaded after its P V product, so score tiles running
    // more than stages - 2 ahead starve the loads (N = 4096, d = 40: 3 stages, nS 2 -> 1: 199 -> 182 us)
    if (stages >= 3 && p.nS > stages - 2) p.nS = stages - 2;
    if (ns_force >= 1 && ns_force <= kMaxS && ns_force <= stages - 1 && ns_force <= (p.tmem_cols - d.dp) / 64) p.nS = ns_force;
    if (p.nS < 1) return set_error(S2I_ERR_ARG, "attn_fwd: head dim %d leaves no TMEM for the score tiles", d.dp);
    const size_t smem_bytes = (size_t)q_bytes + (size_t)pbufs * kChunk16 + (size_t)stages * stage_bytes + overhead;

    {
        uint64_t dims[3] = {(uint64_t)d.ldq, (uint64_t)d.Nq, (uint64_t)d.B};
        uint64_t str[2] = {(uint64_t)d.ldq * 2, (uint64_t)d.Nq * d.ldq * 2};
        uint32_t box[3] = {64, (uint32_t)kTileQ, 1};
        S2I_TRY(encode_tmap_f16(&p.mapQ, 3, d.q, dims, str, box));
    }
    {
        uint64_t dims[3] = {(uint64_t)d.ldkv, (uint64_t)d.Nk, (uint64_t)d.B};
        uint64_t str[2] = {(uint64_t)d.ldkv * 2, (uint64_t)d.Nk * d.ldkv * 2};
        uint32_t box[3] = {64, (uint32_t)kTileK, 1};
        S2I_TRY(encode_tmap_f16(&p.mapKV, 3, d.kv, dims, str, box));
    }
    static bool attr_set = false;
    if (!attr_set) {
        S2I_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        S2I_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)((d.Nq + kTileQ - 1) / kTileQ), (unsigned)(d.B * d.heads), 1);
    if (triple) S2I_LAUNCH((attn_fwd_kernel<3>), grid, kThreads, smem_bytes, stream, p);
    else S2I_LAUNCH((attn_fwd_kernel<1>), grid, kThreads, smem_bytes, stream, p);
    // algorithmic work: QK^T + PV at the true head dim
    S2I_LAUNCH_CHECK_TAG("attn_fwd", 4.0 * d.B * d.heads * (double)d.Nq * d.Nk * d.d_true,
                         2.0 * d.B * d.heads * d.d_true * (2.0 * d.Nq + 2.0 * d.Nk));      // Q, O, K, V once (fp16)
    return 0;
}

}  // namespace s2i
