// tcgen05 + TMA implicit GEMM (see gemm_tc.cuh).  One CTA computes one 128 x BN output tile:
//   warp 0      TMA producer  (cp.async.bulk.tensor, 128B-swizzled operand tiles, zero-filled halos)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (accumulator: 128 lanes x BN fp32 columns)
//   warps 2..5  epilogue: tcgen05.ld -> bias / per-sample vector / ReLU / residual -> fp32 and/or fp16 stores
// Stage ring: full[] (TMA -> MMA, transaction bytes) and empty[] (tcgen05.commit -> TMA).
#include "gemm_tc.cuh"
#include "common.cuh"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <string>

namespace s2i {

namespace {

constexpr int kThreads = 192;
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;             // elements; 128 bytes = one swizzle row
constexpr int kAStageBytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kChunkBytes = 64 * kBlockK * 2;         // 8 KB: one 64-row (or 64-col, MN-major) box
constexpr int kMaxCluster = 16;     // split-K cluster size: 8 is the portable maximum, 16 needs the non-portable attribute

struct __align__(64) GemmParams {
    CUtensorMap mapA;
    CUtensorMap mapB;
    int tw, th, tb, tiles_x, tiles_y;
    int tw_shift, th_shift;  // tw and th are powers of two: tile row -> pixel by shifts
    int W, H, Bn, M, N;
    int k_chunks, k_last_steps, taps, a_mn, b_mn;
    int BN, stages, tmem_cols;
    int a_c0, a_hoff, a_zmode, b_c0, b_hoff, b_zmode, zh;
    uint32_t idesc, tx_bytes;
    float alpha;
    const float* bias;
    const float* rowvec;
    int rowvec_ld;
    const float* residual;
    long res_ld;
    float* out32;
    long ld32;
    void* out16;
    long ld16;
    int out16_bf16;
    long c_sb, c_sh;
    int relu;
    float qscale, qinv;
    int fast_epi;            // aligned operands: coalesced (smem-transposed) epilogue
    int splits, iters_per_split;
    float* ws;               // split-K partial tiles [split][tile][128][BN] fp32
    unsigned int* counters;  // one arrival counter per output tile (self-resetting)
    // ---- TMA epilogue (gemm_tma_kernel): residual tile in, result tiles out through cp.async.bulk.tensor
    CUtensorMap mapRes;      // fp32 residual  (N, W, H, B), box (32, tw, th, tb), 128B swizzle
    CUtensorMap mapO32;      // fp32 output, same geometry
    CUtensorMap mapO16;      // fp16 output, box (32, tw, th, tb), 64B swizzle
    CUtensorMap mapGlu;      // fp16 gated-GELU output [rows][N / 2], same box / swizzle
    int has_res, has_o32, has_o16, has_glu;
    int split_add;           // split-K by fp32 reduce-add into a zeroed output (split 0 carries bias + residual)
    int pipe_bytes;          // shared-memory bytes in front of the barriers (max of pipeline, epilogue staging, partial tile)
    // Deterministic split-K (dsplit != 0): splits = csplit * groups CTAs share an output tile.  The csplit CTAs of a thread-block
    // cluster (1, 1, csplit) reduce their partial tiles through distributed shared memory in rank order (CTA rank r owns the tile
    // rows [r * 128 / csplit, ...)); with groups > 1 each cluster writes its reduced rows to ws [group][tile][128][BN] and the CTA
    // that arrives LAST for its (tile, row slice) adds the groups' rows in group order and runs the epilogue.  No zero-fill, no
    // floating-point atomics: the result does not depend on timing.
    int dsplit, csplit, groups;
    int msub;                // M sub-tiles per CTA (1 or 2): two 128-row A tiles share every B tile (two TMEM accumulators)
    // CTA pair (pair != 0): the CTAs of two adjacent M tiles (cluster x = 2) run ONE tcgen05.mma.cta_group::2 stream of 256 x BN
    // instructions, issued by the even CTA.  Each CTA loads its own 128 A rows and HALF of the B tile (b_rows = BN / 2), so a tile
    // costs 16 KB + BN * 64 B per K step instead of 16 KB + BN * 128 B from L2; accumulators, epilogue and split-K are per CTA.
    int pair, b_rows;
    // Column statistics of the result for a following GroupNorm (GemmDesc::colstat): cstat [B][cst_cap][2][cst_ld]
    float* cstat;
    long cst_ld;
    int cst_cap, hw_shift;   // hw_shift = log2(tw * th): tile row >> hw_shift = sample within the tile
    int split_issue;         // A and B tiles issued by two warps
    int early_b;             // B is static (weights): its first tiles are requested before griddepcontrol.wait
    int tiles_m;             // number of 128-row M tiles of the problem
    unsigned long long* trace;   // optional [ctas][16] %globaltimer stamps of the kernel's phases (tools/gemm_trace.py)
};

__device__ __forceinline__ void stamp(const GemmParams& p, int slot) {
    if (p.trace) {
        // SM cycle counter: reading %globaltimer costs ~1 us (it distorted every phase it was meant to measure)
        const unsigned long long t = (unsigned long long)clock64();
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        p.trace[(long)cta * 16 + slot] = t;
    }
}

// ---- thread-block cluster helpers (split-K inside a cluster) ---------------------------------------------------------
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_smem_addr), "r"(rank));
    return ra;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

constexpr int kEpiLd = 36;   // floats per row of the per-warp 32x32 transpose buffer (16-byte aligned, conflict-free)

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcg_f4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ uint16_t to_half_bits(float v, int bf16) {
    if (bf16) {
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        return *reinterpret_cast<uint16_t*>(&h);
    }
    __half h = __float2half_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
}

struct TileCtx {
    int x0, y0, b0, m0, n0, z, zb, zhd, split, it_begin, it_end;
    int x1, y1, b1;          // second M sub-tile (msub == 2)
    int mt0;                 // index of the first M tile of this CTA
};

__device__ __forceinline__ void tile_xyb(const GemmParams& p, int mt, int& x0, int& y0, int& b0) {
    const int tx = mt % p.tiles_x;
    const int ty = (mt / p.tiles_x) % p.tiles_y;
    const int tbi = mt / (p.tiles_x * p.tiles_y);
    x0 = tx * p.tw;
    y0 = ty * p.th;
    b0 = tbi * p.tb;
}

__device__ __forceinline__ TileCtx tile_ctx(const GemmParams& p) {
    TileCtx t;
    const int msub = p.msub > 1 ? p.msub : 1;
    const int mt = blockIdx.x * msub;
    t.mt0 = mt;
    t.n0 = blockIdx.y * p.BN;
    t.z = blockIdx.z / p.splits;
    t.split = blockIdx.z - t.z * p.splits;      // cluster split-K: clusters are (1, 1, splits), so this is also the cluster rank
    t.zb = t.z / p.zh;
    t.zhd = t.z - t.zb * p.zh;
    t.x0 = t.y0 = t.b0 = t.m0 = 0;
    t.x1 = t.y1 = t.b1 = 0;
    if (!p.a_mn) {
        tile_xyb(p, mt, t.x0, t.y0, t.b0);
        if (msub > 1) tile_xyb(p, mt + 1, t.x1, t.y1, t.b1);
    } else {
        t.m0 = mt * kBlockM;
    }
    const int total_iters = p.taps * p.k_chunks;
    t.it_begin = t.split * p.iters_per_split;
    t.it_end = min(total_iters, t.it_begin + p.iters_per_split);
    return t;
}

// Stream the A / B operand tiles of this CTA's K range through the stage ring.  Called by ALL lanes of the producer warp: the loop
// state is warp-uniform (the compiler keeps it in uniform registers) and one elected lane issues the barrier arrive and the copies.
// The loop body is a dependent instruction chain that paces the whole CTA.  Measured (tools/gemm_trace.py, clock64 stamps): with the
// stage index / phase / tap / K chunk recomputed by integer division every trip a stage was issued every ~1000 cycles whatever the
// pipeline depth; with incremental counters in a single-lane (divergent) loop ~700; hence everything loop-invariant is hoisted,
// counters advance incrementally, and the loop runs converged.
// role: 0 = this warp issues both operands; 1 = A tiles (and arms the barriers), 2 = B tiles only -- two warps then issue in
// parallel (a tensor copy costs its issuing thread ~200 cycles, and one thread issuing both operands paced the ring).
template <bool PAIR = false>
__device__ __forceinline__ void producer_loop(const GemmParams& p, const TileCtx& t, uint8_t* smem, int stage_bytes,
                                              uint64_t* full, uint64_t* empty, bool early_b = false, int role = 0) {
    const bool leader = ptx::elect_one();
    const bool do_a = role != 2, do_b = role != 1;
    const int a_inner = p.a_c0 + (p.a_zmode ? 0 : t.zhd * p.a_hoff);
    const int a_bz = p.a_zmode ? t.z : t.zb;
    const int b_inner = p.b_c0 + (p.b_zmode ? 0 : t.zhd * p.b_hoff);
    const int b_bz = p.b_zmode ? t.z : t.zb;
    const int nchunks_b = (p.BN + 63) >> 6;
    const int pr = PAIR ? (int)(cluster_ctarank() & 1u) : 0;
    const int stages = p.stages, k_chunks = p.k_chunks;
    const bool a_mn = p.a_mn != 0, b_mn = p.b_mn != 0, two = p.msub > 1, taps9 = p.taps == 9, traced = p.trace != nullptr;
    const uint32_t tx = p.tx_bytes;
    const int a_bytes = (two ? 2 : 1) * kAStageBytes;
    const int ab = t.b0 + a_bz, ab1 = t.b1 + a_bz;
    const int bn0 = t.n0 + pr * p.b_rows;
    // running state: ring slot, K chunk within the tap, tap offsets, K coordinates
    int stage = 0;
    uint32_t parity = 1;                 // empty[] parity to wait for: the ring starts free
    uint8_t* sa = smem;
    int tap = t.it_begin / k_chunks;
    int kc = t.it_begin - tap * k_chunks;
    int dx = 0, dy = 0;
    if (taps9) {
        dy = tap / 3 - 1;
        dx = tap - (dy + 1) * 3 - 1;
    }
    int ak = a_inner + kc * kBlockK;     // A's K coordinate (K-major A: restarts with every tap)
    int lk = t.it_begin * kBlockK;       // linear K coordinate (B, MN-major A)
    const int n_it = t.it_end - t.it_begin;
    // Static B (weights): the first ring's worth of B tiles does not depend on the preceding kernel -- request them, THEN wait for
    // it (programmatic dependent launch); the A tiles of those slots follow in the loop below.
    int pre = 0;
    if (early_b && role == 1) pdl_wait();       // the B warp streams the weights on its own; A depends on the preceding kernel
    if (early_b && role == 0) {
        pre = n_it < stages ? n_it : stages;
        if (leader) {
            uint8_t* sbp = smem + a_bytes;
            int lkp = lk;
            for (int s = 0; s < pre; ++s, sbp += stage_bytes, lkp += kBlockK) {
                if constexpr (PAIR) {
                    if (pr == 0) ptx::mbar_expect_tx(&full[s], tx);
                    ptx::tma_load_3d_2sm(sbp, &p.mapB, &full[s], b_inner + lkp, bn0, b_bz);
                } else {
                    ptx::mbar_expect_tx(&full[s], tx);
                    ptx::tma_load_3d(sbp, &p.mapB, &full[s], b_inner + lkp, bn0, b_bz);
                }
            }
        }
        pdl_wait();
    }
    for (int li = 0; li < n_it; ++li) {
        const bool b_done = li < pre;          // this slot's barrier is armed and its B tile requested
        if (!b_done) ptx::mbar_wait(&empty[stage], parity);
        if (leader) {
            uint8_t* sb = sa + a_bytes;
            if constexpr (PAIR) {
                // both CTAs' boxes complete on the even CTA's barrier (tx counts the pair's four boxes)
                if (pr == 0 && !b_done && do_a) ptx::mbar_expect_tx(&full[stage], tx);
                if (do_a) ptx::tma_load_4d_2sm(sa, &p.mapA, &full[stage], ak, t.x0 + dx, t.y0 + dy, ab);
                if (!b_done && do_b) ptx::tma_load_3d_2sm(sb, &p.mapB, &full[stage], b_inner + lk, bn0, b_bz);
            } else {
                if (!b_done && do_a) ptx::mbar_expect_tx(&full[stage], tx);
                if (!do_a) {
                } else if (!a_mn) {
                    ptx::tma_load_4d(sa, &p.mapA, &full[stage], ak, t.x0 + dx, t.y0 + dy, ab);
                    if (two) ptx::tma_load_4d(sa + kAStageBytes, &p.mapA, &full[stage], ak, t.x1 + dx, t.y1 + dy, ab1);
                } else {
                    ptx::tma_load_4d(sa, &p.mapA, &full[stage], a_inner + t.m0, lk, 0, a_bz);
                    ptx::tma_load_4d(sa + kChunkBytes, &p.mapA, &full[stage], a_inner + t.m0 + 64, lk, 0, a_bz);
                }
                if (b_done || !do_b) {
                    // requested before the wait (K-major static B) / the other warp's
                } else if (!b_mn) {
                    ptx::tma_load_3d(sb, &p.mapB, &full[stage], b_inner + lk, bn0, b_bz);
                } else {
                    for (int j = 0; j < nchunks_b; ++j)
                        ptx::tma_load_3d(sb + j * kChunkBytes, &p.mapB, &full[stage], b_inner + t.n0 + j * 64, lk, b_bz);
                }
            }
            if (traced && li < 3 && do_a) stamp(p, 10 + li);
        }
        lk += kBlockK;
        ak += kBlockK;
        if (++kc == k_chunks) {
            kc = 0;
            ak = a_inner;
            if (taps9 && ++dx == 2) {
                dx = -1;
                ++dy;
            }
        }
        sa += stage_bytes;
        if (++stage == stages) {
            stage = 0;
            sa = smem;
            parity ^= 1u;
        }
    }
    __syncwarp();
}

// Issue the tcgen05.mma stream of this CTA's K range; accum_full fires when the accumulator is complete.  Called by ALL lanes of the
// MMA warp (warp-uniform loop state, one elected lane issues), same discipline as producer_loop: incremental ring / chunk counters,
// descriptors built by adding to a per-stage base.
template <bool PAIR = false>
__device__ __forceinline__ void mma_loop(const GemmParams& p, const TileCtx& t, uint8_t* smem, int stage_bytes,
                                         uint64_t* full, uint64_t* empty, uint64_t* accum_full, uint32_t tmem_base) {
    const bool leader = ptx::elect_one();
    const uint32_t a_kstep = p.a_mn ? 2048u : 32u;   // bytes per UMMA_K = 16 elements
    const uint32_t b_kstep = p.b_mn ? 2048u : 32u;
    const uint32_t a_lbo = p.a_mn ? (uint32_t)kChunkBytes : 16u;
    const uint32_t b_lbo = p.b_mn ? (uint32_t)kChunkBytes : 16u;
    const uint32_t pair_mask = PAIR ? (3u << (cluster_ctarank() & ~1u)) : 0u;      // this pair's two cluster ranks
    const int stages = p.stages, k_chunks = p.k_chunks, k_last = p.k_last_steps;
    const bool two = p.msub > 1;
    const uint32_t idesc = p.idesc;
    const uint32_t a_bytes = (two ? 2u : 1u) * kAStageBytes;
    const uint32_t tmem1 = tmem_base + (uint32_t)p.BN;
    // descriptors of ring slot 0, K step 0; the start-address field counts 16-byte units, so slots and K steps are additions
    const uint32_t s0 = ptx::smem_u32(smem);
    const uint64_t adesc0 = ptx::make_smem_desc_sw128(s0, a_lbo, 1024u);
    const uint64_t bdesc0 = ptx::make_smem_desc_sw128(s0 + a_bytes, b_lbo, 1024u);
    const uint64_t a_step = a_kstep >> 4, b_step = b_kstep >> 4, slot_step = (uint32_t)stage_bytes >> 4;
    const uint64_t a1_off = (uint32_t)kAStageBytes >> 4;
    int stage = 0;
    uint32_t parity = 0;
    uint64_t slot_off = 0;
    int kc = t.it_begin % k_chunks;
    uint32_t accumulate = 0;
    const int n_it = t.it_end - t.it_begin;
    for (int li = 0; li < n_it; ++li) {
        ptx::mbar_wait(&full[stage], parity);
        ptx::tc_fence_after();
        // The last K chunk of a tap may be partial: head-sliced operands must not read past Kc.
        const int ksteps = (kc == k_chunks - 1) ? k_last : kBlockK / 16;
        if (leader) {
            uint64_t adesc = adesc0 + slot_off, bdesc = bdesc0 + slot_off;
            uint32_t acc = accumulate;
            for (int k = 0; k < ksteps; ++k) {
                if constexpr (PAIR) {
                    ptx::umma_f16_2sm(tmem_base, adesc, bdesc, idesc, acc);
                } else {
                    ptx::umma_f16(tmem_base, adesc, bdesc, idesc, acc);
                    // second 128-row sub-tile against the same B tile -> second accumulator
                    if (two) ptx::umma_f16(tmem1, adesc + a1_off, bdesc, idesc, acc);
                }
                acc = 1u;
                adesc += a_step;
                bdesc += b_step;
            }
            if constexpr (PAIR) ptx::umma_commit_2sm(&empty[stage], pair_mask);   // frees the stage in both CTAs
            else ptx::umma_commit(&empty[stage]);                                 // frees the smem stage once these MMAs have read it
        }
        accumulate = 1u;
        if (++kc == k_chunks) kc = 0;
        slot_off += slot_step;
        if (++stage == stages) {
            stage = 0;
            slot_off = 0;
            parity ^= 1u;
        }
    }
    if (leader) {
        if constexpr (PAIR) ptx::umma_commit_2sm(accum_full, pair_mask);
        else ptx::umma_commit(accum_full);     // accumulator complete
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------------------
// gemm_tma_kernel: same mainloop, epilogue entirely through TMA.  Per 32-column chunk each epilogue thread (one tile
// row) does: tcgen05.ld -> + bias (staged in smem) -> + residual (tile loaded by TMA into the idle pipeline stages,
// 128B-swizzled, read and overwritten in place) -> fp32 / fp16 staging -> one thread issues the bulk tensor store.
// No per-thread global addressing, no predicates: ~100 instructions per chunk instead of ~1000 in gemm_tc_kernel.
// Supports K-major operands, Z = 1, N % 32 == 0, alpha = 1, no ReLU / rounding emulation.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }   // F.gelu

constexpr int kChunk32Bytes = kBlockM * 32 * 4;   // 16 KB: 128 rows x 32 fp32 columns
constexpr int kChunk16Bytes = kBlockM * 32 * 2;   //  8 KB: 128 rows x 32 fp16 columns
constexpr int kMaxChunks = 8;

// ESETS epilogue warp quartets (warps 2..5, and 6..9 when ESETS = 2) take the tile's 32-column chunks in turn: the chunks
// are independent (own residual / staging region, own bulk store), and one quartet alone works through them at ~0.75 us
// per chunk -- a latency chain (TMEM load, shared-memory round trips, proxy fence, barrier), not a bandwidth limit.
// GLU: the gated-GELU epilogue (its own instantiation, so the plain epilogue's register budget is untouched).
// PAIR: the CTA-pair form (GemmParams::pair) -- its own instantiations: a kernel containing cta_group::2 instructions can only be
// launched as clusters of an even number of CTAs.
template <int ESETS, bool GLU, bool PAIR = false>
__global__ void __launch_bounds__(64 + 128 * ESETS, ESETS == 2 ? 2 : 1)
gemm_tma_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_stage_bytes = PAIR ? ((p.b_rows * 128 + 1023) & ~1023) : ((p.BN + 63) >> 6) * kChunkBytes;
    const int msub = p.msub > 1 ? 2 : 1;
    const int stage_bytes = msub * kAStageBytes + b_stage_bytes;
    const int nch = p.BN >> 5;
    const int use32 = p.has_res | p.has_o32;
    const bool in_cluster = PAIR || (p.dsplit && p.csplit > 1);
    const uint32_t crank = in_cluster ? cluster_ctarank() : 0u;
    const bool mma_leader = !PAIR || (crank & 1u) == 0;
    // pipeline stages; aliased after the mainloop by the epilogue staging / the cluster split-K partial tile (host: launch_tma)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.pipe_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* accum_full = empty + p.stages;
    uint64_t* r_full = accum_full + 1;                      // [kMaxChunks]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + kMaxChunks);
    float* bias_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));   // [BN]
    float* cst_s = bias_s + ((p.BN + 8 + 3) & ~3);          // [ESETS][4 quarters][2][32]: column statistics being combined

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const TileCtx t = tile_ctx(p);
    if (threadIdx.x == 0) stamp(p, 0);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.mapA);
        ptx::prefetch_tmap(&p.mapB);
        if (p.has_res) ptx::prefetch_tmap(&p.mapRes);
        if (p.has_o32) ptx::prefetch_tmap(&p.mapO32);
        if (p.has_o16) ptx::prefetch_tmap(&p.mapO16);
        if (p.has_glu) ptx::prefetch_tmap(&p.mapGlu);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; ++s) {
                ptx::mbar_init(&full[s], 1);
                ptx::mbar_init(&empty[s], 1);
            }
            ptx::mbar_init(accum_full, 1);
            for (int c = 0; c < kMaxChunks; ++c) ptx::mbar_init(&r_full[c], 1);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        if constexpr (PAIR) {
            ptx::tmem_alloc_2sm(tmem_slot, p.tmem_cols);
            ptx::tmem_relinquish_2sm();
        } else {
            ptx::tmem_alloc(tmem_slot, p.tmem_cols);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if constexpr (PAIR) {
        // the peer's loads complete on, and the even CTA's commits arrive at, barriers of the OTHER CTA: both must be initialised
        __syncwarp();
        cluster_arrive();
        cluster_wait();
    } else {
        __syncthreads();
    }
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Everything above is local setup; global memory of earlier kernels is touched only below.  With static B (weights) the wait
    // moves into the roles that touch global memory: the producer (after requesting its first B tiles) and the epilogue warps.
    const bool early_b = p.early_b != 0;
    if (!early_b) pdl_wait();
    pdl_launch();     // TMEM is held: dependents may become resident
    if (threadIdx.x == 0) stamp(p, 1);
    const bool add_res = p.has_res && (!p.split_add || t.split == 0);
    const bool add_bias = p.dsplit || !p.split_add || t.split == 0;

    const bool split_issue = p.split_issue != 0;      // warp 0 issues the A tiles, warp 2 the B tiles (then joins the epilogue)
    if (warp == 0) {
        producer_loop<PAIR>(p, t, smem, stage_bytes, full, empty, early_b, split_issue ? 1 : 0);       // all lanes: warp-uniform loop, one elected issuer
        if (lane == 0) {
            stamp(p, 2);
            if (add_res) {
                // the pipeline stages are idle once the accumulator is complete: land the residual tile there
                ptx::mbar_wait(accum_full, 0);
                for (int c = 0; c < nch; ++c) {
                    if (t.n0 + c * 32 >= p.N) break;
                    ptx::mbar_expect_tx(&r_full[c], (uint32_t)(p.tw * p.th * p.tb * 128));
                    ptx::tma_load_4d(smem + c * kChunk32Bytes, &p.mapRes, &r_full[c], t.n0 + c * 32, t.x0, t.y0, t.b0);
                }
            }
        }
    } else if (warp == 1) {
        if (mma_leader) mma_loop<PAIR>(p, t, smem, stage_bytes, full, empty, accum_full, tmem_base);     // all lanes, one elected issuer
        if (lane == 0 && mma_leader) stamp(p, 3);
    } else {
        // ---------------------------------------------------- epilogue (warps 2..5), thread <-> tile row
        if (split_issue && warp == 2) producer_loop<PAIR>(p, t, smem, stage_bytes, full, empty, early_b, 2);
        if (early_b) pdl_wait();
        const int e = threadIdx.x - 64;
        const int eset = e >> 7;                // warp quartet: chunks eset, eset + ESETS, ...
        const int el = e & 127;
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;
        for (int i = e; i < p.BN; i += 128 * ESETS) {
            const int n = t.n0 + i;
            float b = 0.f;
            if (add_bias && n < p.N) {
                if (p.bias) b = __ldg(p.bias + n);
                if (p.rowvec) b += __ldg(p.rowvec + n);
            }
            bias_s[i] = b;
        }
        ptx::named_bar_sync(1, 128 * ESETS);
        if (e == 0) stamp(p, 4);
        ptx::mbar_wait(accum_full, 0);
        ptx::tc_fence_after();
        if (e == 0) stamp(p, 5);
        const uint32_t sw128 = (uint32_t)(r & 7);
        const uint32_t sw64 = (uint32_t)((r >> 1) & 3);
        uint8_t* base16 = smem + (use32 ? nch * kChunk32Bytes : 0);
        if (p.dsplit) {
            // deterministic split-K: this CTA's partial tile (its K range) -> shared memory [128][BN + 4] fp32 (the idle pipeline
            // stages); the cluster reduces it below, after the role branches
            float* part = reinterpret_cast<float*>(smem);
            const int ldp = p.BN + 4;
            for (int c = eset; c < nch; c += ESETS) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), raw);
                ptx::tmem_ld_wait();
                float* dst = part + (long)r * ldp + c * 32;
    #pragma unroll
                for (int u = 0; u < 8; ++u)
                    *reinterpret_cast<float4*>(dst + 4 * u) =
                        make_float4(__uint_as_float(raw[4 * u]), __uint_as_float(raw[4 * u + 1]), __uint_as_float(raw[4 * u + 2]),
                                    __uint_as_float(raw[4 * u + 3]));
            }
            if (e == 0) stamp(p, 6);       // split-K: partial tile parked
        } else
        for (int sub = 0; sub < msub; ++sub) {
            if (t.mt0 + sub >= p.tiles_m) break;          // odd tile count: the last CTA has one sub-tile
            const int xs = sub ? t.x1 : t.x0, ys = sub ? t.y1 : t.y0, bs = sub ? t.b1 : t.b0;
            if (sub > 0) {
                // the staging buffers are reused: the first sub-tile's stores must have read them; its residual tile
                // is then loaded over them by this thread (second phase of the r_full barriers)
                if (el == 0) ptx::tma_store_wait_read0();      // each quartet's issuing thread waits for its own bulk groups
                ptx::named_bar_sync(1, 128 * ESETS);
                if (e == 0 && add_res) {
                    for (int c = 0; c < nch; ++c) {
                        if (t.n0 + c * 32 >= p.N) break;
                        ptx::mbar_expect_tx(&r_full[c], (uint32_t)(p.tw * p.th * p.tb * 128));
                        ptx::tma_load_4d(smem + c * kChunk32Bytes, &p.mapRes, &r_full[c], t.n0 + c * 32, xs, ys, bs);
                    }
                }
            }
            if constexpr (GLU) {
                // gated GELU: chunk 2 pc holds 32 value columns, chunk 2 pc + 1 their gates -> 32 output columns
                uint8_t* baseG = base16 + (p.has_o16 ? nch * kChunk16Bytes : 0);
                for (int pc = eset; pc < (nch >> 1); pc += ESETS) {
                    if (t.n0 + pc * 64 >= p.N) break;
                    const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * p.BN + pc * 64);
                    uint8_t* rowG = baseG + pc * kChunk16Bytes + r * 64;
                    uint8_t* rowA = base16 + (2 * pc) * kChunk16Bytes + r * 64;
                    uint8_t* rowB = rowA + kChunk16Bytes;
    #pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {            // 16 value + 16 gate columns at a time (register budget)
                        uint32_t ra[16], rg[16];
                        ptx::tmem_ld_32x16(tcol + (uint32_t)(hh * 16), ra);
                        ptx::tmem_ld_32x16(tcol + 32u + (uint32_t)(hh * 16), rg);
                        ptx::tmem_ld_wait();
    #pragma unroll
                        for (int u = 0; u < 2; ++u) {           // 8 columns -> one 16-byte unit of each staging row
                            float a[8], g[8];
    #pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                a[j] = __uint_as_float(ra[8 * u + j]) + bias_s[pc * 64 + hh * 16 + 8 * u + j];
                                g[j] = __uint_as_float(rg[8 * u + j]) + bias_s[pc * 64 + 32 + hh * 16 + 8 * u + j];
                            }
                            uint32_t ho[4], hv[4], hw[4];
    #pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const __half2 o = __floats2half2_rn(a[2 * j] * gelu_erf(g[2 * j]), a[2 * j + 1] * gelu_erf(g[2 * j + 1]));
                                const __half2 v = __floats2half2_rn(a[2 * j], a[2 * j + 1]);
                                const __half2 w = __floats2half2_rn(g[2 * j], g[2 * j + 1]);
                                ho[j] = *reinterpret_cast<const uint32_t*>(&o);
                                hv[j] = *reinterpret_cast<const uint32_t*>(&v);
                                hw[j] = *reinterpret_cast<const uint32_t*>(&w);
                            }
                            const uint32_t unit = (uint32_t)(hh * 2 + u);
                            *reinterpret_cast<uint4*>(rowG + ((unit ^ sw64) << 4)) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
                            if (p.has_o16) {
                                *reinterpret_cast<uint4*>(rowA + ((unit ^ sw64) << 4)) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
                                *reinterpret_cast<uint4*>(rowB + ((unit ^ sw64) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                            }
                        }
                    }
                    ptx::fence_proxy_async();
                    ptx::named_bar_sync(2 + eset, 128);
                    if (el == 0) {
                        ptx::tma_store_4d(&p.mapGlu, baseG + pc * kChunk16Bytes, (t.n0 >> 1) + pc * 32, xs, ys, bs);
                        if (p.has_o16) {
                            ptx::tma_store_4d(&p.mapO16, base16 + (2 * pc) * kChunk16Bytes, t.n0 + pc * 64, xs, ys, bs);
                            ptx::tma_store_4d(&p.mapO16, base16 + (2 * pc + 1) * kChunk16Bytes, t.n0 + pc * 64 + 32, xs, ys, bs);
                        }
                        ptx::tma_store_commit();
                    }
                }
                continue;
            }
            for (int c = eset; c < nch; c += ESETS) {
                if (t.n0 + c * 32 >= p.N) break;
                uint32_t raw[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * p.BN + c * 32), raw);
                if (add_res) ptx::mbar_wait(&r_full[c], (uint32_t)(sub & 1));
                ptx::tmem_ld_wait();
                if (e == 0 && c == 0) stamp(p, 6);
                uint8_t* row32 = smem + c * kChunk32Bytes + r * 128;
                uint32_t hp[16];
    #pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + u * 4);
                    float4 v = make_float4(__uint_as_float(raw[4 * u]) + b4.x, __uint_as_float(raw[4 * u + 1]) + b4.y,
                                           __uint_as_float(raw[4 * u + 2]) + b4.z, __uint_as_float(raw[4 * u + 3]) + b4.w);
                    float4* slot = reinterpret_cast<float4*>(row32 + ((u ^ sw128) << 4));
                    if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (p.qscale != 0.f) {      // mimic unscaled fp16 autograd rounding (LGP backward)
                        v.x = __half2float(__float2half_rn(v.x * p.qinv)) * p.qscale;
                        v.y = __half2float(__float2half_rn(v.y * p.qinv)) * p.qscale;
                        v.z = __half2float(__float2half_rn(v.z * p.qinv)) * p.qscale;
                        v.w = __half2float(__float2half_rn(v.w * p.qinv)) * p.qscale;
                    }
                    if (add_res) {
                        const float4 x = *slot;
                        v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
                    }
                    if (p.has_o32) *slot = v;
                    if (p.has_o16) {
                        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
                        hp[2 * u] = *reinterpret_cast<const uint32_t*>(&h0);
                        hp[2 * u + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                    }
                }
                if (p.has_o16) {
                    uint8_t* row16 = base16 + c * kChunk16Bytes + r * 64;
    #pragma unroll
                    for (int w = 0; w < 4; ++w)
                        *reinterpret_cast<uint4*>(row16 + ((w ^ sw64) << 4)) =
                            make_uint4(hp[4 * w], hp[4 * w + 1], hp[4 * w + 2], hp[4 * w + 3]);
                }
                // the chunk's bulk store overlaps the next chunk's arithmetic (the copy engine, like the loads, runs at the
                // chip's ~10 TB/s L2 rate: a late burst of all stores would only lengthen the tail)
                ptx::fence_proxy_async();              // generic-proxy smem writes -> visible to the bulk-copy engine
                ptx::named_bar_sync(2 + eset, 128);
                if (el == 0) {
                    const int nc = t.n0 + c * 32;
                    if (p.has_o32) {
                        if (p.split_add) ptx::tma_reduce_add_4d(&p.mapO32, smem + c * kChunk32Bytes, nc, xs, ys, bs);
                        else ptx::tma_store_4d(&p.mapO32, smem + c * kChunk32Bytes, nc, xs, ys, bs);
                    }
                    if (p.has_o16) ptx::tma_store_4d(&p.mapO16, base16 + c * kChunk16Bytes, nc, xs, ys, bs);
                    ptx::tma_store_commit();
                    if (e == 0 && c < 3) stamp(p, 13 + c);
                }
                if (p.cstat) {
                    // column sums / sums of squares of this chunk for the GroupNorm that follows: each warp takes its 32 rows
                    // from the staged fp32 tile (lane = column: conflict-free under the 128-byte swizzle), the quarters of one
                    // sample are combined in quarter order, one float per (sample, tile, statistic, column) goes out
                    const uint8_t* c32 = smem + c * kChunk32Bytes;
                    const uint32_t cu = (uint32_t)lane >> 2, cw = ((uint32_t)lane & 3u) << 2;
                    float s1 = 0.f, s2 = 0.f;
    #pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) {
                        const uint32_t row = (uint32_t)(q * 32 + rr);
                        const float v = *reinterpret_cast<const float*>(c32 + row * 128u + (((cu ^ (row & 7u)) << 4) | cw));
                        s1 += v;
                        s2 = fmaf(v, v, s2);
                    }
                    float* cs = cst_s + eset * 256;
                    cs[q * 64 + lane] = s1;
                    cs[q * 64 + 32 + lane] = s2;
                    ptx::named_bar_sync(2 + eset, 128);
                    const int qps = (1 << p.hw_shift) >> 5;                  // quarters per sample (tw * th >= 32)
                    const int bi = (q * 32) >> p.hw_shift;                   // sample within the tile
                    if ((q & (qps - 1)) == 0 && bi < p.tb) {
                        float t1 = 0.f, t2 = 0.f;
                        for (int k = 0; k < qps; ++k) {
                            t1 += cs[(q + k) * 64 + lane];
                            t2 += cs[(q + k) * 64 + 32 + lane];
                        }
                        const int tile_xy = (ys >> p.th_shift) * p.tiles_x + (xs >> p.tw_shift);
                        float* dst = p.cstat + (((long)(bs + bi) * p.cst_cap + tile_xy) * 2) * p.cst_ld + t.n0 + c * 32 + lane;
                        dst[0] = t1;
                        dst[p.cst_ld] = t2;
                    }
                    ptx::named_bar_sync(2 + eset, 128);                      // the scratch is reused by this quartet's next chunk
                }
            }
        }
        if (el == 0) {
            if (e == 0) stamp(p, 7);
            ptx::tma_store_wait_read0();   // shared memory stays valid until the stores have read it
            if (e == 0) stamp(p, 8);
        }
    }

    if (p.dsplit) {
        // ---------------------------------------------------- deterministic split-K reduction (see GemmParams::dsplit)
        const int CS = p.csplit, G = p.groups;
        const int rank = t.split % CS, group = t.split / CS;
        ptx::tc_fence_before();
        if (CS > 1 || PAIR) {
            __syncwarp();
            cluster_arrive();
            cluster_wait();
        } else {
            __syncthreads();
        }
        if (warp >= 2) {
            const int e = threadIdx.x - 64;
            if (e == 0) stamp(p, 7);       // split-K: every partial of the cluster is visible
            const int rows_per = kBlockM / CS;
            const int row0 = rank * rows_per;
            const int nq = p.BN >> 2;
            const int ldp = p.BN + 4;
            const int tile_id = blockIdx.x + gridDim.x * blockIdx.y;
            const int num_tiles = gridDim.x * gridDim.y;
            const uint32_t part_local = ptx::smem_u32(smem);
            uint32_t base_k[kMaxCluster];
    #pragma unroll
            for (int k = 0; k < kMaxCluster; ++k)       // split k of this tile: cluster rank k, or 2 k + (M tile parity) in pair mode
                base_k[k] = map_to_rank(part_local, PAIR ? (uint32_t)(k < CS ? 2 * k : 0) + (crank & 1u) : (uint32_t)(k < CS ? k : 0));
            // bias / ReLU / rounding emulation / residual, then the fp32 and / or fp16 rows of the output
            float* part_own = reinterpret_cast<float*>(smem);
            auto finish = [&](float4 acc, int c4, int n, long grow, const float4& res4, int row) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c4);
                acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
                if (p.relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
                if (p.qscale != 0.f) {
                    acc.x = __half2float(__float2half_rn(acc.x * p.qinv)) * p.qscale;
                    acc.y = __half2float(__float2half_rn(acc.y * p.qinv)) * p.qscale;
                    acc.z = __half2float(__float2half_rn(acc.z * p.qinv)) * p.qscale;
                    acc.w = __half2float(__float2half_rn(acc.w * p.qinv)) * p.qscale;
                }
                if (p.residual) { acc.x += res4.x; acc.y += res4.y; acc.z += res4.z; acc.w += res4.w; }
                // column statistics: park the finished values in this CTA's own rows of its partial tile (no peer reads those)
                if (p.cstat) *reinterpret_cast<float4*>(part_own + (long)row * ldp + c4) = acc;
                if (p.out32) *reinterpret_cast<float4*>(p.out32 + grow * p.ld32 + n) = acc;
                if (p.out16) {
                    const __half2 h0 = __floats2half2_rn(acc.x, acc.y), h1 = __floats2half2_rn(acc.z, acc.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<const uint32_t*>(&h1);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out16) + grow * p.ld16 + n) = pk;
                }
            };
            // global row of tile row `row`, or -1 when it falls outside the tensor
            const int tw_shift = p.tw_shift, th_shift = p.th_shift;
            auto global_row = [&](int row) -> long {
                const int xi = row & (p.tw - 1);
                const int yi = (row >> tw_shift) & (p.th - 1);
                const int bi = row >> (tw_shift + th_shift);
                if (bi >= p.tb || t.x0 + xi >= p.W || t.y0 + yi >= p.H || t.b0 + bi >= p.Bn) return -1;
                return ((long)(t.b0 + bi) * p.H + (t.y0 + yi)) * p.W + (t.x0 + xi);
            };
            float* wsT = G > 1 ? p.ws + ((long)group * num_tiles + tile_id) * kBlockM * p.BN : nullptr;
            // kU float4 units per thread per trip: up to 8 (distributed) shared-memory loads and kU residual loads in flight (more
            // would spill under the two-CTAs-per-SM register bound)
            auto reduce_rows = [&](auto cs_tag) {
                constexpr int kCS = decltype(cs_tag)::value;
                constexpr int kU = kCS >= 8 ? 1 : 2;
                constexpr int kL = kCS > 8 ? 8 : kCS;        // loads in flight per unit (a cluster of 16 takes two passes, in rank order)
                const int total = rows_per * nq;
                // unit index -> (row, column unit), advanced without divisions: one per thread up front
                constexpr int kStep = 128 * ESETS;
                const int step_r = kStep / nq, step_c = kStep - step_r * nq;
                int ur = e / nq, uc = e - ur * nq;
                for (int base = e; base < total; base += kU * 128 * ESETS) {
                    float4 v[8], rs[kU];
                    long grow[kU];
                    int c4s[kU], rows[kU];
    #pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        grow[u] = -1;
                        const int idx = base + u * 128 * ESETS;
                        const int rr = ur;
                        const int cc = uc;
                        ur += step_r;
                        uc += step_c;
                        if (uc >= nq) {
                            uc -= nq;
                            ++ur;
                        }
                        if (idx < total) {
                            c4s[u] = cc << 2;
                            rows[u] = row0 + rr;
                            if (t.n0 + c4s[u] < p.N) grow[u] = global_row(rows[u]);
                            if (grow[u] >= 0) {
                                const uint32_t off = (uint32_t)((rows[u] * ldp + c4s[u]) * 4);
    #pragma unroll
                                for (int k = 0; k < kL; ++k) v[u * kL + k] = ld_cluster_f4(base_k[k] + off);
                                if (G == 1 && p.residual) rs[u] = ldg_f4(p.residual + grow[u] * p.res_ld + t.n0 + c4s[u]);
                            }
                        }
                    }
                    if (e == 0 && base == 0) stamp(p, 14);       // split-K: first trip's loads issued
    #pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        if (grow[u] >= 0) {
                            float4 acc = v[u * kL];
    #pragma unroll
                            for (int k = 1; k < kL; ++k) {
                                const float4 x = v[u * kL + k];
                                acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
                            }
                            if (kCS > 8) {
                                const uint32_t off = (uint32_t)((rows[u] * ldp + c4s[u]) * 4);
                                float4 w[8];
    #pragma unroll
                                for (int k = 0; k < 8; ++k) w[k] = ld_cluster_f4(base_k[8 + k] + off);
    #pragma unroll
                                for (int k = 0; k < 8; ++k) { acc.x += w[k].x; acc.y += w[k].y; acc.z += w[k].z; acc.w += w[k].w; }
                            }
                            if (G > 1) *reinterpret_cast<float4*>(wsT + (long)rows[u] * p.BN + c4s[u]) = acc;   // this group's rows, raw sums
                            else finish(acc, c4s[u], t.n0 + c4s[u], grow[u], rs[u], rows[u]);
                        }
                    }
                    if (e == 0 && base == 0) stamp(p, 15);       // split-K: first trip finished
                }
            };
            if (e == 0) stamp(p, 13);      // split-K: reduction starts (addresses mapped)
            if (CS == 16) reduce_rows(std::integral_constant<int, 16>());
            else if (CS == 8) reduce_rows(std::integral_constant<int, 8>());
            else if (CS == 4) reduce_rows(std::integral_constant<int, 4>());
            else if (CS == 2) reduce_rows(std::integral_constant<int, 2>());
            else reduce_rows(std::integral_constant<int, 1>());
            if (e == 0) stamp(p, 8);       // split-K: this CTA's rows written
            if (p.cstat && G == 1) {
                // column statistics of this CTA's row slice (all rows of one sample: host guarantees rows_per <= tw * th)
                ptx::named_bar_sync(1, 128 * ESETS);
                const int bi = row0 >> p.hw_shift;
                if (bi < p.tb) {
                    const int pps = (1 << p.hw_shift) / rows_per;                          // row slices per sample in this tile
                    const int tile_xy = (t.y0 >> p.th_shift) * p.tiles_x + (t.x0 >> p.tw_shift);
                    const int blk = tile_xy * pps + ((row0 & ((1 << p.hw_shift) - 1)) / rows_per);
                    for (int col = e; col < p.BN; col += 128 * ESETS) {
                        if (t.n0 + col >= p.N) continue;
                        float s1 = 0.f, s2 = 0.f;
                        for (int rr = 0; rr < rows_per; ++rr) {
                            const float v = part_own[(long)(row0 + rr) * ldp + col];
                            s1 += v;
                            s2 = fmaf(v, v, s2);
                        }
                        float* dst = p.cstat + (((long)(t.b0 + bi) * p.cst_cap + blk) * 2) * p.cst_ld + t.n0 + col;
                        dst[0] = s1;
                        dst[p.cst_ld] = s2;
                    }
                }
            }
            if (G > 1) {
                // the CTA arriving last for this (tile, row slice) adds the groups' rows in group order
                uint32_t* flag = reinterpret_cast<uint32_t*>(bias_s + p.BN);
                __threadfence();
                ptx::named_bar_sync(1, 128 * ESETS);
                if (e == 0) {
                    unsigned int* ctr = p.counters + tile_id * kMaxCluster + rank;
                    const unsigned int old = atomicAdd(ctr, 1u);
                    const bool last = old == (unsigned int)(G - 1);
                    if (last) *ctr = 0u;              // self-reset for the next launch
                    *flag = last ? 1u : 0u;
                }
                ptx::named_bar_sync(1, 128 * ESETS);
                if (*flag) {
                    __threadfence();
                    const float* ws0 = p.ws + (long)tile_id * kBlockM * p.BN;
                    const long gstride = (long)num_tiles * kBlockM * p.BN;
                    for (int idx = e; idx < rows_per * nq; idx += 128 * ESETS) {
                        const int rr = idx / nq;
                        const int c4 = (idx - rr * nq) << 2;
                        const int row = row0 + rr;
                        const int n = t.n0 + c4;
                        if (n >= p.N) continue;
                        const long grow = global_row(row);
                        if (grow < 0) continue;
                        const float* src = ws0 + (long)row * p.BN + c4;
                        float4 acc = ldcg_f4(src);
                        for (int g = 1; g < G; ++g) {
                            const float4 x = ldcg_f4(src + g * gstride);
                            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
                        }
                        float4 res4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.residual) res4 = ldg_f4(p.residual + grow * p.res_ld + n);
                        finish(acc, c4, n, grow, res4, row);
                    }
                }
            }
        }
        if (CS > 1 && !PAIR) {
            __syncwarp();
            cluster_arrive();         // this CTA no longer reads its peers' shared memory ...
            cluster_wait();           // ... and no peer reads this CTA's: it may exit
        }
    }

    ptx::tc_fence_before();
    if constexpr (PAIR) {
        // both CTAs are done with the pair's tensor memory and with each other's shared memory (barriers, split-K partial tiles)
        __syncwarp();
        cluster_arrive();
        cluster_wait();
        if (threadIdx.x == 0) stamp(p, 9);
        if (warp == 1) ptx::tmem_dealloc_2sm(tmem_base, p.tmem_cols);
        return;
    }
    __syncthreads();
    if (threadIdx.x == 0) stamp(p, 9);
    if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_stage_bytes = ((p.BN + 63) >> 6) * kChunkBytes;
    const int stage_bytes = kAStageBytes + b_stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* accum_full = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const TileCtx t = tile_ctx(p);
    const int n0 = t.n0, zb = t.zb, zhd = t.zhd, split = t.split;
    const int x0 = t.x0, y0 = t.y0, b0 = t.b0, m0 = t.m0;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&p.mapA);
        ptx::prefetch_tmap(&p.mapB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; ++s) {
                ptx::mbar_init(&full[s], 1);
                ptx::mbar_init(&empty[s], 1);
            }
            ptx::mbar_init(accum_full, 1);
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

    if (warp == 0) {
        producer_loop(p, t, smem, stage_bytes, full, empty);
    } else if (warp == 1) {
        mma_loop(p, t, smem, stage_bytes, full, empty, accum_full, tmem_base);
    } else {
        // ---------------------------------------------------- epilogue (warps 2..5)
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;            // tile row owned by this thread in the TMEM layout
        bool valid;
        long row;
        int sample = 0;
        if (!p.a_mn) {
            const int xi = r % p.tw;
            const int yi = (r / p.tw) % p.th;
            const int bi = r / (p.tw * p.th);
            valid = (bi < p.tb) && (x0 + xi < p.W) && (y0 + yi < p.H) && (b0 + bi < p.Bn);
            row = ((long)(b0 + bi) * p.H + (y0 + yi)) * p.W + (x0 + xi);
            sample = b0 + bi;
        } else {
            valid = (m0 + r) < p.M;
            row = m0 + r;
        }
        const long zoff = (long)zb * p.c_sb + (long)zhd * p.c_sh;

        ptx::mbar_wait(accum_full, 0);
        ptx::tc_fence_after();

        if (p.fast_epi) {
            // Coalesced path.  TMEM hands each thread one ROW (32 consecutive columns per load); global memory wants
            // a warp instruction to cover whole 128-byte row segments.  Each warp therefore transposes its 32x32 block
            // through shared memory (the pipeline stages are idle once the accumulator is complete) and then works with
            // lane -> (row i*4 + lane/8, columns 4*(lane%8)..+3): 8 lanes cover one 128-byte segment.  Bias / residual
            // loads for a chunk are issued before the TMEM data is needed, so their latency overlaps the tcgen05.ld.
            float* sT = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * kEpiLd);
            const int cq = (lane & 7) * 4;
            const int rsub = lane >> 3;
            const long my_row = valid ? row : -1;
            long rows8[8];
            int smp8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rows8[i] = __shfl_sync(0xffffffffu, my_row, i * 4 + rsub);
                smp8[i] = __shfl_sync(0xffffffffu, sample, i * 4 + rsub);
            }
            const int tile_id = blockIdx.x + gridDim.x * blockIdx.y;
            const int num_tiles = gridDim.x * gridDim.y;
            float* wsT = p.splits > 1
                             ? p.ws + ((long)(split * num_tiles + tile_id) * kBlockM + q * 32) * p.BN
                             : nullptr;

            // applies alpha / bias / per-sample vector / ReLU / fp16-rounding emulation / residual and stores 4 columns
            auto finish4 = [&](float4 v, const float4& b4, const float4& rs, long ro, int smp, int n) {
                v.x = v.x * p.alpha + b4.x; v.y = v.y * p.alpha + b4.y; v.z = v.z * p.alpha + b4.z; v.w = v.w * p.alpha + b4.w;
                if (p.rowvec && p.rowvec_ld != 0) {
                    const float4 rv = ldg_f4(p.rowvec + (long)smp * p.rowvec_ld + n);
                    v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
                }
                if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                if (p.qscale != 0.f) {
                    v.x = __half2float(__float2half_rn(v.x * p.qinv)) * p.qscale;
                    v.y = __half2float(__float2half_rn(v.y * p.qinv)) * p.qscale;
                    v.z = __half2float(__float2half_rn(v.z * p.qinv)) * p.qscale;
                    v.w = __half2float(__float2half_rn(v.w * p.qinv)) * p.qscale;
                }
                v.x += rs.x; v.y += rs.y; v.z += rs.z; v.w += rs.w;
                if (p.out32) *reinterpret_cast<float4*>(p.out32 + zoff + ro * p.ld32 + n) = v;
                if (p.out16) {
                    uint2 pk;
                    pk.x = (uint32_t)to_half_bits(v.x, p.out16_bf16) | ((uint32_t)to_half_bits(v.y, p.out16_bf16) << 16);
                    pk.y = (uint32_t)to_half_bits(v.z, p.out16_bf16) | ((uint32_t)to_half_bits(v.w, p.out16_bf16) << 16);
                    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out16) + zoff + ro * p.ld16 + n) = pk;
                }
            };
            auto load_b4 = [&](int n, bool colok) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (colok && p.bias) b4 = ldg_f4(p.bias + n);
                if (colok && p.rowvec && p.rowvec_ld == 0) {
                    const float4 t = ldg_f4(p.rowvec + n);
                    b4.x += t.x; b4.y += t.y; b4.z += t.z; b4.w += t.w;
                }
                return b4;
            };

            for (int c = 0; c < p.BN; c += 32) {
                const int ncols = min(min(32, p.BN - c), p.N - (n0 + c));     // multiple of 4 on this path
                if (ncols <= 0) break;
                uint32_t raw[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
                if (c + 32 <= p.BN) {
                    ptx::tmem_ld_32x32(taddr, raw);
                } else {   // BN is a multiple of 16: 16-column tail
                    uint32_t lo[16];
                    ptx::tmem_ld_32x16(taddr, lo);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        raw[j] = lo[j];
                        raw[16 + j] = 0u;
                    }
                }
                const bool colok = cq < ncols;
                const int n = n0 + c + cq;
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 rs[8];
                if (p.splits == 1) {
                    b4 = load_b4(n, colok);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.residual && colok && rows8[i] >= 0) rs[i] = ldg_f4(p.residual + zoff + rows8[i] * p.res_ld + n);
                    }
                }
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(sT + lane * kEpiLd + 4 * j) =
                        make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                    __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(sT + (i * 4 + rsub) * kEpiLd + cq);
                    if (p.splits == 1) {
                        if (colok && rows8[i] >= 0) finish4(v, b4, rs[i], rows8[i], smp8[i], n);
                    } else if (colok) {
                        *reinterpret_cast<float4*>(wsT + (long)(i * 4 + rsub) * p.BN + c + cq) = v;   // raw partial sums
                    }
                }
                __syncwarp();
            }

            if (p.splits > 1) {
                // The last CTA to arrive for this output tile reduces every split's partial tile (in split order, so the
                // result does not depend on arrival order) and runs the real epilogue.
                uint32_t* flag = reinterpret_cast<uint32_t*>(sT + 4 * 32 * kEpiLd - (warp - 2) * (32 * kEpiLd));
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (warp == 2 && lane == 0) {
                    const unsigned int old = atomicAdd(p.counters + tile_id, 1u);
                    const bool last = (old == (unsigned int)(p.splits - 1));
                    if (last) p.counters[tile_id] = 0u;      // self-reset for the next launch
                    *flag = last ? 1u : 0u;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (*flag) {
                    __threadfence();
                    const float* ws0 = p.ws + ((long)tile_id * kBlockM + q * 32) * p.BN;
                    const long split_stride = (long)num_tiles * kBlockM * p.BN;
                    for (int c = 0; c < p.BN; c += 32) {
                        const int ncols = min(min(32, p.BN - c), p.N - (n0 + c));
                        if (ncols <= 0) break;
                        const bool colok = cq < ncols;
                        const int n = n0 + c + cq;
                        if (!colok) continue;
                        const float4 b4 = load_b4(n, true);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (rows8[i] < 0) continue;
                            float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.residual) rs = ldg_f4(p.residual + zoff + rows8[i] * p.res_ld + n);
                            const float* src = ws0 + (long)(i * 4 + rsub) * p.BN + c + cq;
                            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int sp = 0; sp < p.splits; ++sp) {
                                const float4 t = ldcg_f4(src + sp * split_stride);
                                acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                            }
                            finish4(acc, b4, rs, rows8[i], smp8[i], n);
                        }
                    }
                }
            }
        } else {
            // Generic path (odd widths / unaligned outputs, e.g. N not a multiple of 4): one thread per row.
            float* o32 = p.out32 ? p.out32 + zoff + row * p.ld32 : nullptr;
            uint16_t* o16 = p.out16 ? reinterpret_cast<uint16_t*>(p.out16) + zoff + row * p.ld16 : nullptr;
            const float* res = p.residual ? p.residual + zoff + row * p.res_ld : nullptr;
            const float* rvec = p.rowvec ? p.rowvec + (long)sample * p.rowvec_ld : nullptr;
            for (int c = 0; c < p.BN; c += 32) {
                uint32_t raw[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
                if (c + 32 <= p.BN) {
                    ptx::tmem_ld_32x32(taddr, raw);
                } else {
                    uint32_t lo[16];
                    ptx::tmem_ld_32x16(taddr, lo);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        raw[j] = lo[j];
                        raw[16 + j] = 0u;
                    }
                }
                ptx::tmem_ld_wait();
                if (!valid) continue;
                const int nbase = n0 + c;
                const int ncols = min(min(32, p.BN - c), p.N - nbase);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j >= ncols) break;
                    const int n = nbase + j;
                    float v = __uint_as_float(raw[j]) * p.alpha;
                    if (p.bias) v += __ldg(p.bias + n);
                    if (rvec) v += __ldg(rvec + n);
                    if (p.relu) v = fmaxf(v, 0.f);
                    if (p.qscale != 0.f) v = __half2float(__float2half_rn(v * p.qinv)) * p.qscale;
                    if (res) v += res[n];
                    if (o32) o32[n] = v;
                    if (o16) o16[n] = to_half_bits(v, p.out16_bf16);
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
    });
    return fn;
}

int encode_map(CUtensorMap* m, int dtype, int rank, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes = 128) {
    auto fn = get_encode_fn();
    if (!fn) return set_error(S2I_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(S2I_ERR_CUDA,
                         "cuTensorMapEncodeTiled failed (%d): rank %d ptr %p dims [%llu %llu %llu %llu] strides [%llu %llu "
                         "%llu] box [%u %u %u %u]",
                         (int)r, rank, ptr, (unsigned long long)dims[0], (unsigned long long)dims[1],
                         (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                         (unsigned long long)strides_bytes[0], (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
                         (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
                         rank > 3 ? box[3] : 0);
    return 0;
}

int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}

// ---- tile-shape / split-K selection ------------------------------------------------------------------------------
// Rough per-CTA cycle model (B300_MICROARCH.md pacing law: one 128 x BN x 16 MMA takes ~BN/2 cycles; L2 -> SM operand
// traffic at ~64 B/clk/SM when every SM pulls; ~6 k cycles of fixed launch / prologue / first-load latency; epilogue
// ~8 cycles per column).  The problem sizes on this path are at most a few waves, so filling the 148 SMs (x2 resident
// CTAs when shared memory allows) matters as much as the inner-loop rate.
constexpr size_t kWsBytes = 96u << 20;      // split-K scratch: partial tiles (gemm_tc_kernel) / group rows (gemm_tma_kernel)
constexpr int kMaxSplitTiles = 4096;        // arrival counters

struct TileChoice {
    int BN, splits;
    int msub = 1;
    int cs = 1;       // deterministic split-K: cluster size (splits = cs * groups)
};

double model_cycles(int N, long tiles_m, int Z, int iters, int BN, int splits) {
    const int tiles_n = ceil_div(N, BN);
    const long ctas = tiles_m * tiles_n * Z * splits;
    const int stage_bytes = kAStageBytes + ceil_div(BN, 64) * kChunkBytes;
    const int occ = (3 * stage_bytes <= 100 * 1024) ? 2 : 1;
    const long slots = (long)kNumSMs * occ;
    const long waves = ceil_div_l(ctas, slots);
    const int it = ceil_div(iters, splits);
    const double resident = (double)(ctas < slots ? ceil_div_l(ctas, kNumSMs) : occ);   // CTAs sharing one SM
    const double mma = 2.0 * BN * resident;                                  // cycles per 64-wide K chunk, SM shared
    const double tma = (double)stage_bytes / 64.0 * resident;
    const double per_iter = mma > tma ? mma : tma;
    double fixed = 6000.0 + 8.0 * BN;
    if (splits > 1) fixed += 1500.0 + 2.0 * BN * splits;                     // partial store + last-CTA reduction
    return (double)waves * (it * per_iter + fixed);
}

TileChoice choose_tiles(int N, long tiles_m, int Z, int iters, bool allow_split) {
    const int nr = (int)round_up_l(N, 16);
    static const int cands[] = {256, 192, 160, 128, 96, 80, 64, 48, 32, 16};
    TileChoice best{nr <= 256 ? nr : 128, 1};
    double best_c = 1e30;
    for (int c : cands) {
        if (c > nr && c != 16) {
            if (nr > 256 || c != cands[0]) continue;
        }
        const int bn = c > nr ? nr : c;
        const long base = tiles_m * ceil_div(N, bn) * Z;
        for (int sp = 1; sp <= 32; sp *= 2) {
            if (sp > 1 && (!allow_split || base >= kNumSMs || iters / sp < 4)) break;
            if ((long)(sp - 1) * ceil_div(iters, sp) >= iters) continue;     // an empty split
            const double cyc = model_cycles(N, tiles_m, Z, iters, bn, sp);
            if (cyc < best_c) {
                best_c = cyc;
                best = TileChoice{bn, sp};
            }
        }
    }
    return best;
}

// ---- gemm_tma_kernel tile choice ----------------------------------------------------------------------------------
// Same pacing model with this kernel's constants: prologue + first-load latency ~4.5 k cycles, ~450 cycles per
// 32-column epilogue chunk (+ the residual tile's L2 round trip), operands arriving at ~48 B/clk per SM when every SM
// pulls from L2.  Split-K partial tiles are reduce-added into the (zeroed) output by the bulk-copy engine.
double model_cycles_tma(int N, long tiles_m_all, int iters, int BN, int splits, int epi, int msub = 1, bool cta_pair = false) {
    const bool residual = (epi & 1) != 0;
    const long tiles_m = ceil_div_l(tiles_m_all, msub);
    // Calibrated on B200 with tools/gemm_bench.py (profiles/r1_gemm_bench_v2.txt): every CTA pays ~9 k cycles of launch,
    // prologue, first-load and drain latency, so small problems want MANY short CTAs (two co-resident CTAs per SM hide
    // each other's latencies); operand tiles arrive from L2 at ~40 B/clk per SM when the whole chip pulls.
    const int tiles_n = ceil_div(N, BN);
    const long ctas = tiles_m * tiles_n * splits;
    const int stage_bytes = msub * kAStageBytes + ceil_div(BN, 64) * kChunkBytes;
    // two CTAs share an SM only if the pipeline (>= 2 stages) AND the epilogue staging that aliases it fit half of it,
    // and their accumulators fit the 512 TMEM columns together
    const int epi_bytes = (BN / 32) * (((epi & 2) ? 16384 : 0) + ((epi & 4) ? 8192 : 0));
    const int occ = (2 * stage_bytes <= 100 * 1024 && epi_bytes <= 100 * 1024 && msub * BN <= 256) ? 2 : 1;
    const long slots = (long)kNumSMs * occ;
    const long waves = ceil_div_l(ctas, slots);
    const int it = ceil_div(iters, splits);
    const double resident = (double)(ctas <= kNumSMs ? 1 : occ);
    const double mma = 2.0 * BN * resident * msub;
    // measured operand arrival: ~31 B/clk for a CTA alone on its SM, ~42 B/clk shared by two co-resident CTAs
    // (the chip-wide L2 rate: what matters is bytes per FLOP, hence the two-sub-tile form for K-heavy problems)
    // (a CTA pair loads half of each B tile per CTA)
    const double tma = (double)(msub * kAStageBytes + BN * (cta_pair ? 64 : 128)) / (resident > 1.0 ? 21.0 : 31.0);
    const double per_iter = mma > tma ? mma : tma;
    double fixed = 9000.0 + msub * 350.0 * (BN / 32) + (residual ? 900.0 * msub : 0.0) + (msub - 1) * 500.0 + (cta_pair ? 800.0 : 0.0);
    double total = (double)waves * (it * per_iter + fixed);
    if (splits > 1) total += 4500.0;      // zero-fill node + reduce-add traffic
    return total;
}

bool g_allow_msub = [] {
    const char* e = getenv("S2I_GEMM_MSUB");
    return !(e && e[0] == '0');
}();

// Split-K mode of gemm_tma_kernel.  Default: inside a thread-block cluster (1, 1, S), S in {2, 4, 8}: the partial tiles are
// reduced through distributed shared memory in rank order -- deterministic, no zero-fill launch, no fp32 scratch + cast for
// fp16 outputs.  S2I_SPLITK_ADD=1 selects the previous form (up to 32 splits reduce-added into a zeroed output by
// cp.reduce.async.bulk: the sum order, hence the low bits of the result, varies from run to run).
bool g_split_add_mode = [] {
    const char* e = getenv("S2I_SPLITK_ADD");
    return e && e[0] == '1';
}();

// How many clusters of `cs` CTAs of gemm_tma_kernel can be resident at once, for CTAs sized to share an SM in pairs
// (occ = 2: <= 112 KB of shared memory) or not (occ = 1).  Thread-block clusters are gang-scheduled inside a GPC, so this
// is less than #SMs * occ / cs (measured in round 1 for the GroupNorm kernel: 15 clusters of 8, not 18); a grid with more
// clusters than this runs in two waves.  Queried once per (cs, occ) from the occupancy API, filled in by launch_tma.
int g_cluster_cap[kMaxCluster + 1][3] = {};
int cluster_capacity(int cs, int occ);

// deterministic split-K (cluster size cs, `groups` clusters per tile): model of one CTA's cycles, times the number of waves
double model_cycles_dsplit(int N, long tiles_m, int iters, int BN, int cs, int groups, int epi, bool cta_pair = false) {
    const int splits = cs * groups;
    const int tiles_n = ceil_div(N, BN);
    const long ctas = tiles_m * tiles_n * splits;
    const int stage_bytes = kAStageBytes + ceil_div(BN, 64) * kChunkBytes;
    const long part_bytes = (long)kBlockM * (BN + 4) * 4;
    const int occ = (2 * stage_bytes <= 100 * 1024 && part_bytes <= 108 * 1024 && BN <= 256) ? 2 : 1;
    const long n_clusters = tiles_m * tiles_n * groups;
    const int cw = cta_pair ? 2 : 1;                      // CTA pairs: clusters (2, 1, cs) span two tiles
    const int cap1 = cluster_capacity(cs * cw, 1) * cw;   // ... in tile-clusters
    // CTAs are sized to pair up on an SM when the clusters do not fit one per SM (launch_tma does the same).  The occupancy API
    // reports the same cluster count for such CTAs as for unpaired ones; measured, twice that many clusters still run as one wave
    const int use_occ = (n_clusters > cap1 && occ == 2) ? 2 : 1;
    const long cap = (long)cap1 * use_occ;
    const long waves = ceil_div_l(n_clusters, cap > 0 ? cap : 1);
    const int it = ceil_div(iters, splits);
    const double resident = (double)((ctas <= kNumSMs && use_occ == 1) ? 1 : use_occ);
    const double mma = 2.0 * BN * resident;
    const double tma = (double)(kAStageBytes + BN * (cta_pair ? 64 : 128)) / (resident > 1.0 ? 21.0 : 31.0);
    const double per_iter = mma > tma ? mma : tma;
    const double rows_per = (double)kBlockM / cs;
    // prologue / first load / drain as in model_cycles_tma; TMEM -> smem 60 cycles per chunk; barriers; one tile's worth of
    // fp32 through (distributed) shared memory at ~48 B/clk; the row slice's global epilogue; with groups: the slice out to
    // and back from L2 (groups + 1 passes), fence + counter
    // (+ 3000: measured, a split costs ~1.5 us more than the terms below add up to -- cluster launch and barrier skew)
    double fixed = 12000.0 + 60.0 * (BN / 32) + (cs > 1 ? 1200.0 : 300.0) + (double)kBlockM * BN * 4 / 48.0 +
                   rows_per * BN * 8.0 / 24.0 + ((epi & 1) ? 600.0 : 0.0);
    if (groups > 1) fixed += 2500.0 + rows_per * BN * 4.0 * (groups + 1) / 24.0;
    return (double)waves * (it * per_iter + fixed);
}

// no_msub: the caller will run CTA pairs (which take the place of the two-sub-tile form).  The tile width and cluster size are
// chosen with the single-CTA model -- the sweeps of both forms (tools/gemm_bench.py psweep) agree on them except where a
// cluster of 16 would not fit (launch_tma adjusts those).
TileChoice choose_tiles_tma(int N, long tiles_m, int iters, bool allow_split, int epi, bool no_msub = false) {
    static const int cands[] = {256, 192, 160, 128, 96, 64, 32};
    TileChoice best{N >= 128 ? 128 : N, 1};
    double best_c = 1e30;
    for (int c : cands) {
        if (c > N && c != 32) continue;
        if (!g_split_add_mode && allow_split) {
            const long base = tiles_m * ceil_div(N, c);
            for (int cs = 2; cs * (no_msub ? 2 : 1) <= kMaxCluster; cs *= 2) {
                // one cluster per tile: measured (tools/gemm_bench.py dsweep, profiles/r2_gemm_dsweep_v1.txt) the best
                // configuration of every shape of the step has groups = 1 -- the extra trip through L2 (group rows out, fence,
                // counter, rows back in) costs more than the added CTAs bring; groups > 1 stays available to callers
                for (int groups = 1; groups <= 1; ++groups) {
                    const int sp = cs * groups;
                    if (sp == 1) continue;
                    if (iters / sp < 2 || base * sp > 2 * kNumSMs + 64 || base * kMaxCluster > kMaxSplitTiles) break;
                    if ((long)(sp - 1) * ceil_div(iters, sp) >= iters) continue;      // an empty split
                    if (groups > 1 && (size_t)groups * base * kBlockM * c * sizeof(float) > kWsBytes) continue;
                    const double cyc = model_cycles_dsplit(N, tiles_m, iters, c, cs, groups, epi);
                    if (cyc < best_c) {
                        best_c = cyc;
                        best = TileChoice{c, sp, 1, cs};
                    }
                }
            }
        }
        for (int ms = 1; ms <= 2; ++ms) {
            if (ms == 2 && (tiles_m < 2 || ms * c > 512 || !g_allow_msub || no_msub)) break;
            const long base = ceil_div_l(tiles_m, ms) * ceil_div(N, c);
            // powers of two, and every count above 8: 180 K iterations split 30 ways (6 each) where 32 would leave empty
            // splits (measured, tools/gemm_bench.py sweep: conv 1280 @ 8x8 20.3 -> 14.3 us; small odd counts measured worse
            // than the model predicts)
            for (int sp = 1; sp <= 32; ++sp) {
                if (sp > 1 && (!allow_split || !g_split_add_mode || base * sp > 2 * kNumSMs + 64 || iters / sp < 2)) break;
                if (sp < 8 && (sp & (sp - 1)) != 0) continue;
                if ((long)(sp - 1) * ceil_div(iters, sp) >= iters) continue;
                const double cyc = model_cycles_tma(N, tiles_m, iters, c, sp, epi, ms);
                if (cyc < best_c) {
                    best_c = cyc;
                    best = TileChoice{c, sp, ms};
                }
            }
        }
    }
    return best;
}

unsigned long long* g_trace = nullptr;
// CTA pairs in gemm_tma_kernel: 1 = wherever legal, 0 = never, -1 = the cost model's choice
int g_pair_mode = [] { const char* e = getenv("S2I_GEMM_PAIR"); return e ? atoi(e) : -1; }();
int g_force_msub = 0;      // tools / tests: 1 or 2 forces the M sub-tile count of gemm_tma_kernel, 0 = model's choice
int g_tma_epi = -1;
bool tma_epilogue_enabled() {
    if (g_tma_epi < 0) {
        const char* e = getenv("S2I_GEMM_TMA_EPI");
        g_tma_epi = (e && e[0] == '0') ? 0 : 1;
    }
    return g_tma_epi != 0;
}

// split-K scratch: partial tiles + per-tile arrival counters (one stream at a time uses the library)
float* g_ws = nullptr;
unsigned int* g_counters = nullptr;
int g_ws_device = -1;

int ensure_ws() {
    int dev = 0;
    S2I_CUDA(cudaGetDevice(&dev));
    if (g_ws && dev == g_ws_device) return 0;
    void* a = nullptr;
    void* b = nullptr;
    if (cudaMalloc(&a, kWsBytes) != cudaSuccess || cudaMalloc(&b, kMaxSplitTiles * sizeof(unsigned int)) != cudaSuccess) {
        cudaGetLastError();
        return set_error(S2I_ERR_OOM, "gemm: cannot allocate the split-K workspace");
    }
    S2I_CUDA(cudaMemset(b, 0, kMaxSplitTiles * sizeof(unsigned int)));
    ++g_alloc_gen;
    g_ws = static_cast<float*>(a);
    g_counters = static_cast<unsigned int*>(b);
    g_ws_device = dev;
    return 0;
}

int build_operand_maps(GemmParams& p, const GemmDesc& d, int BN) {      // BN: rows of one CTA's B box
    {
        const long sw = d.a_sw > 0 ? d.a_sw : d.aC;
        const long sh = d.a_sh > 0 ? d.a_sh : sw * d.aW;
        const long sb = d.a_sb > 0 ? d.a_sb : sh * d.aH;
        uint64_t dims[4] = {(uint64_t)d.aC, (uint64_t)d.aW, (uint64_t)d.aH, (uint64_t)d.aB};
        uint64_t str[3] = {(uint64_t)sw * 2, (uint64_t)sh * 2, (uint64_t)sb * 2};
        uint32_t box[4];
        if (!d.a_mn) {
            box[0] = kBlockK; box[1] = (uint32_t)p.tw; box[2] = (uint32_t)p.th; box[3] = (uint32_t)p.tb;
        } else {
            box[0] = 64; box[1] = kBlockK; box[2] = 1; box[3] = 1;
        }
        S2I_TRY(encode_map(&p.mapA, d.bf16, 4, d.A, dims, str, box));
    }
    {
        const long sr = d.b_sr > 0 ? d.b_sr : d.bI;
        const long sz = d.b_sz > 0 ? d.b_sz : sr * d.bR;
        uint64_t dims[3] = {(uint64_t)d.bI, (uint64_t)d.bR, (uint64_t)(d.bZ > 0 ? d.bZ : 1)};
        uint64_t str[2] = {(uint64_t)sr * 2, (uint64_t)sz * 2};
        uint32_t box[3];
        if (!d.b_mn) {
            box[0] = kBlockK; box[1] = (uint32_t)BN; box[2] = 1;
        } else {
            box[0] = 64; box[1] = kBlockK; box[2] = 1;
        }
        S2I_TRY(encode_map(&p.mapB, d.bf16, 3, d.B, dims, str, box));
    }
    return 0;
}

// Output-side map: [rows...][N] tensor with the A operand's pixel geometry, 32-column boxes.
int build_out_map(CUtensorMap* m, const GemmParams& p, const GemmDesc& d, const void* ptr, long ld, int dtype) {
    const int esz = dtype == 2 ? 4 : 2;
    uint64_t dims[4] = {(uint64_t)d.N, (uint64_t)d.aW, (uint64_t)d.aH, (uint64_t)d.aB};
    uint64_t str[3] = {(uint64_t)ld * esz, (uint64_t)ld * d.aW * esz, (uint64_t)ld * d.aW * d.aH * esz};
    uint32_t box[4] = {32, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
    return encode_map(m, dtype, 4, ptr, dims, str, box, dtype == 2 ? 128 : 64);
}

// Zero-fill of a split-K output [rows][n4 * 4] (row pitch ld) as a KERNEL: unlike a memset node it keeps the programmatic
// dependent launch chain (the GEMM's prologue overlaps it) -- ~150 split-K GEMMs per denoising step each had one.
__global__ void __launch_bounds__(256) zero_rows_kernel(float* __restrict__ dst, long ld, long rows, int n4) {
    pdl_wait();
    pdl_launch();
    const long total = rows * n4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / n4;
        const int c = (int)(i - r * n4) * 4;
        *reinterpret_cast<float4*>(dst + r * ld + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Every range of a zero plan in one launch: blockIdx.y = range.
__global__ void __launch_bounds__(256) zero_ranges_kernel(const ZeroRange* __restrict__ ranges) {
    pdl_wait();
    pdl_launch();
    const ZeroRange r = ranges[blockIdx.y];
    const long total = r.rows * r.n4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long row = i / r.n4;
        const int c = (int)(i - row * r.n4) * 4;
        *reinterpret_cast<float4*>(r.p + row * r.ld + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

ZeroMode g_zero_mode = ZeroMode::kOff;
std::vector<ZeroRange>* g_zero_plan = nullptr;

// fp32 [rows][N] (row pitch lds) -> fp16 [rows][N] (row pitch ldd); N % 4 == 0
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ src, long lds, __half* __restrict__ dst,
                                                        long ldd, long rows, int n4) {
    pdl_wait();
    pdl_launch();
    const long total = rows * n4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / n4;
        const int c = (int)(i - r * n4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&a);
        pk.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dst + r * ldd + c) = pk;
    }
}

// gemm_tma_kernel as thread-block clusters (1, 1, cz) along the split-K dimension
template <typename... KArgs, typename... Args>
inline void launch_kernel_cluster_z(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, int cx, int cz, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cx;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)cz;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_pdl && g_prev_kernel) ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    g_prev_kernel = true;
}

int cluster_capacity(int cs, int occ) {
    if (cs < 1 || cs > kMaxCluster || occ < 1 || occ > 2) return kNumSMs;
    int& slot = g_cluster_cap[cs][occ];
    if (slot > 0) return slot;
    int cap = kNumSMs * occ / cs;                 // fallback when the query is unavailable (e.g. no device at build time)
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(gemm_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(gemm_tma_kernel<2, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute(gemm_tma_kernel<1, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute(gemm_tma_kernel<2, false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute(gemm_tma_kernel<1, false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute(gemm_tma_kernel<2, true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, (unsigned)(cs * 64));
    cfg.blockDim = dim3(kThreads + 128);
    cfg.dynamicSmemBytes = occ == 2 ? 104 * 1024 : 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)cs;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tma_kernel<2, false>, &cfg) == cudaSuccess && n > 0) cap = n;
    else cudaGetLastError();
    slot = cap;
    if (getenv("S2I_GEMM_DEBUG")) fprintf(stderr, "gemm_tma: %d clusters of %d CTAs can be resident (%d CTA(s) per SM)\n", cap, cs, occ);
    return cap;
}

int launch_tma(GemmParams& p, const GemmDesc& d_in, long tiles_m, int num_iters, cudaStream_t stream) {
    GemmDesc d = d_in;
    const long rows_total = (long)d.aW * d.aH * d.aB;
    const bool glu = d.out_glu != nullptr;
    const bool nonlinear = d.relu || d.qscale != 0.f || glu;      // ReLU / rounding emulation / gating: the epilogue needs the full K sum
    const bool cluster_mode = !g_split_add_mode;
    // cluster split-K applies the epilogue to the reduced tile, so any output form may split; the reduce-add form can only
    // split a plain fp32 output (fp16 outputs go through an fp32 scratch + cast)
    bool can_split = cluster_mode ? (d.splits >= 0 && !glu) : (d.splits >= 0 && d.out32 && !d.out16 && !nonlinear);
    bool via_scratch = false;
    const int epi = (d.residual ? 1 : 0) | ((d.residual || d.out32) ? 2 : 0) | ((d.out16 || glu) ? 4 : 0);
    // CTA pairs pay where the K loop is long enough for the halved B traffic to outweigh the cluster launch (measured with
    // tools/gemm_bench.py pair, profiles/r2_gemm_pair_v1.txt: -10..-25 % on the convolutions and the K-heavy projections,
    // +0.3..0.6 us on the 5- and 10-step projections)
    const bool pair_wanted = cluster_mode && tiles_m % 2 == 0 && d.N >= 64 &&
                             (g_pair_mode == 1 ||
                              (g_pair_mode < 0 && (num_iters >= 40 || (num_iters >= 20 && tiles_m >= 8) || d.N >= 2048)));
    TileChoice tc = choose_tiles_tma(d.N, tiles_m, num_iters, can_split, epi, pair_wanted);
    if (glu && tc.BN % 64 != 0) {
        // value / gate columns come in 32 + 32 pairs: the tile width must hold whole pairs
        static const int cands[] = {256, 192, 128, 64};
        double best_c = 1e30;
        for (int c : cands) {
            if (c > d.N) continue;
            for (int ms = 1; ms <= 2; ++ms) {
                if (ms == 2 && (tiles_m < 2 || ms * c > 512 || pair_wanted)) break;
                const double cyc = model_cycles_tma(d.N, tiles_m, num_iters, c, 1, epi, ms);
                if (cyc < best_c) {
                    best_c = cyc;
                    tc = TileChoice{c, 1, ms};
                }
            }
        }
    }
    if (!cluster_mode && d.splits >= 0 && !nonlinear && d.out16 && !d.out32 && (size_t)rows_total * d.N * sizeof(float) <= kWsBytes) {
        // fp16-only output of a K-heavy small problem: split K into an fp32 scratch tile matrix, then convert
        const int epi_s = (d.residual ? 1 : 0) | 2;
        const TileChoice ts = choose_tiles_tma(d.N, tiles_m, num_iters, true, epi_s);
        if (ts.splits > 1 && model_cycles_tma(d.N, tiles_m, num_iters, ts.BN, ts.splits, epi_s, ts.msub) + 5000.0 <
                                 model_cycles_tma(d.N, tiles_m, num_iters, tc.BN, 1, epi, tc.msub)) {
            if (!d_in.scratch32) S2I_TRY(ensure_ws());
            tc = ts;
            via_scratch = can_split = true;
            d.out32 = d_in.scratch32 ? d_in.scratch32 : g_ws;
            d.ld32 = d.N;
            d.out16 = nullptr;
        }
    }
    if (d.BN > 0) tc.BN = d.BN;
    if (glu && tc.BN % 64 != 0) return set_error(S2I_ERR_ARG, "gemm: the gated-GELU epilogue needs a tile width that is a multiple of 64 (got %d)", tc.BN);
    if (d.splits > 0 && can_split) tc.splits = d.splits;
    if (!can_split) tc.splits = 1;
    if (g_force_msub > 0) tc.msub = (g_force_msub == 2 && tiles_m >= 2 && 2 * tc.BN <= 512) ? 2 : 1;
    while (tc.splits > 1 && (long)(tc.splits - 1) * ceil_div(num_iters, tc.splits) >= num_iters) --tc.splits;
    int groups = 1;
    if (cluster_mode && tc.splits > 1) {
        // splits = cluster size (a power of two up to the portable maximum of 8, dividing the 128 tile rows) x groups
        int cs = tc.cs;
        if ((d.splits > 0 || d.BN > 0) || cs < 1 || tc.splits % cs != 0) {          // forced by the caller: largest fitting cluster
            cs = 1;
            while (cs * 2 <= 8 && tc.splits % (cs * 2) == 0) cs *= 2;
        }
        if (const char* e = getenv("S2I_GEMM_CS")) {                                  // tools: force the cluster size
            const int f = atoi(e);
            if ((f == 1 || f == 2 || f == 4 || f == 8 || f == 16) && tc.splits % f == 0) cs = f;
        }
        groups = tc.splits / cs;
        const long tiles = tiles_m * ceil_div(d.N, tc.BN);
        if (tiles * kMaxCluster > kMaxSplitTiles || (groups > 1 && (size_t)groups * tiles * kBlockM * tc.BN * sizeof(float) > kWsBytes)) {
            groups = 1;                               // does not fit the scratch: one cluster per tile
            tc.splits = cs;
        }
        tc.cs = cs;
        tc.msub = 1;                                  // the partial tile of the reduction is one 128-row accumulator
        if (pair_wanted && d.splits <= 0 && d.BN <= 0 && groups == 1) {
            // CTA pairs make the K loop cheaper, so fewer, longer splits win (tools/gemm_bench.py psweep, profiles/r2_gemm_pair_v2.txt):
            // one wave of CTAs (two when every split still has >= 64 K steps), and 192-wide tiles for the 1280-channel levels
            // (7 column tiles instead of 8: fewer re-reads of A)
            if (d.N >= 1280 && tc.BN == 160) tc.BN = 192;
            const long base = tiles_m * ceil_div(d.N, tc.BN);
            int c2 = 1;
            while (c2 * 2 <= cs && (base * c2 * 2 <= kNumSMs + 4 || (base * c2 * 2 <= 2 * kNumSMs && num_iters / (c2 * 2) >= 64))) c2 *= 2;
            if (tiles_m >= 4) tc.cs = tc.splits = c2;
        }
    }
    const int BN = tc.BN;
    const int msub = tc.msub;
    const bool csplit = cluster_mode && tc.splits > 1;
    // CTA pairs (GemmParams::pair): adjacent M tiles share every B tile through tcgen05.mma.cta_group::2
    const bool cta_pair = pair_wanted && msub == 1 && BN % 32 == 0 && BN >= 64 && (tc.splits == 1 || csplit) &&
                          (csplit ? tc.cs : 1) * 2 <= kMaxCluster;
    static const bool early_ok = [] { const char* e = getenv("S2I_GEMM_EARLY_B"); return !(e && e[0] == '0'); }();
    p.early_b = (d.b_static && early_ok) ? 1 : 0;
    static const bool split_ok = [] { const char* e = getenv("S2I_GEMM_SPLIT_ISSUE"); return !(e && e[0] == '0'); }();
    p.split_issue = split_ok ? 1 : 0;
    p.pair = cta_pair ? 1 : 0;
    p.b_rows = cta_pair ? BN / 2 : BN;
    p.msub = msub;
    p.tiles_m = (int)tiles_m;
    p.BN = BN;
    p.splits = tc.splits;
    p.split_add = (!cluster_mode && tc.splits > 1) ? 1 : 0;
    p.dsplit = csplit ? 1 : 0;
    p.csplit = csplit ? tc.cs : 1;
    p.groups = csplit ? groups : 1;
    if (csplit && groups > 1) {
        S2I_TRY(ensure_ws());
        p.ws = g_ws;
        p.counters = g_counters;
    }
    p.iters_per_split = ceil_div(num_iters, tc.splits);
    p.has_res = (d.residual && !csplit) ? 1 : 0;
    p.has_o32 = (d.out32 && !csplit) ? 1 : 0;
    p.has_o16 = (d.out16 && !csplit) ? 1 : 0;
    p.has_glu = glu ? 1 : 0;
    // column statistics for a following GroupNorm (GemmDesc::colstat): possible when every tile lies inside the tensor, a block of
    // result rows (a warp's 32 rows, or one CTA's row slice of a split-K cluster) belongs to one sample, and the blocks fit `cap`
    p.cstat = nullptr;
    if (d.colstat_bps) *d.colstat_bps = 0;
    if (d.colstat && d.colstat_bps && !glu && !p.split_add && msub >= 1 && d.aW % p.tw == 0 && d.aH % p.th == 0 && d.aB % p.tb == 0) {
        const int hw = p.tw * p.th;
        const int tiles_xy = p.tiles_x * p.tiles_y;
        int bps = 0;
        if (!csplit) {
            if (d.out32 && hw >= 32) bps = tiles_xy;
        } else if (groups == 1 && kBlockM / tc.cs <= hw) {
            bps = tiles_xy * (hw / (kBlockM / tc.cs));
        }
        if (bps > 0 && bps <= d.colstat_cap) {
            p.cstat = d.colstat;
            p.cst_ld = (long)d.colstat_ld;
            p.cst_cap = d.colstat_cap;
            p.hw_shift = p.tw_shift + p.th_shift;
            *d.colstat_bps = bps;
        }
    }
    const int tiles_n = ceil_div(d.N, BN);
    p.tmem_cols = 32;
    while (p.tmem_cols < msub * BN) p.tmem_cols *= 2;

    const int stage_bytes = msub * kAStageBytes + (cta_pair ? (int)round_up_l(p.b_rows * 128, 1024) : ceil_div(BN, 64) * kChunkBytes);
    const int nch = BN / 32;
    const size_t epi_bytes = csplit ? (size_t)kBlockM * (BN + 4) * 4
                                    : (size_t)((p.has_res || p.has_o32) ? nch * kChunk32Bytes : 0) + (p.has_o16 ? nch * kChunk16Bytes : 0) +
                                          (glu ? (nch / 2) * kChunk16Bytes : 0);
    const long grid_m = ceil_div_l(tiles_m, msub);
    const long ctas = grid_m * tiles_n * tc.splits;
    const size_t tail = (size_t)(2 * 8 + 1 + kMaxChunks) * 8 + 16 + (size_t)BN * 4 + 64 + (d.colstat ? 2048 : 0) + 1024 + 64;   // barriers, TMEM slot, bias, flag, column statistics, slack
    // aim for two co-resident CTAs (<= 112 KB each) when more than one wave is coming and they can actually share an SM
    const bool can_pair = 2 * (size_t)stage_bytes + tail <= 112u * 1024u && epi_bytes + tail <= 112u * 1024u && msub * BN <= 256;
    // ... or when the split-K clusters do not fit one CTA per SM (clusters are gang-scheduled: one too many means a second wave)
    const int cluster_ctas = (csplit ? tc.cs : 1) * (cta_pair ? 2 : 1);
    const bool pair = can_pair && (ctas > kNumSMs || (cluster_ctas > 1 && ctas / cluster_ctas > cluster_capacity(cluster_ctas, 1)));
    // S2I_GEMM_SMEM_CAP (KB, experiment): shared-memory budget of a CTA that has the SM to itself
    static const unsigned solo_kb = [] { const char* e = getenv("S2I_GEMM_SMEM_CAP"); const int v = e ? atoi(e) : 0; return v >= 64 && v <= 220 ? (unsigned)v : 220u; }();
    const size_t budget = (pair ? 112u : solo_kb) * 1024u - tail;
    int stages = (int)(budget / stage_bytes);
    if (stages < 2) stages = 2;
    if (stages > 8) stages = 8;
    if (stages > p.iters_per_split) stages = p.iters_per_split < 1 ? 1 : p.iters_per_split;
    p.stages = stages;
    size_t pipe_bytes = (size_t)stages * stage_bytes;
    if (pipe_bytes < epi_bytes) pipe_bytes = epi_bytes;
    p.pipe_bytes = (int)pipe_bytes;
    const size_t smem_bytes = pipe_bytes + tail;
    if (smem_bytes > 227u * 1024u) return set_error(S2I_ERR_ARG, "gemm(tma): %zu bytes of shared memory needed (BN %d)", smem_bytes, BN);
    static const bool dbg = getenv("S2I_GEMM_DEBUG") != nullptr;       // tools: print the chosen configuration
    if (dbg)
        fprintf(stderr, "gemm_tma %s: M-tiles %ld N %d iters %d -> BN %d msub %d%s splits %d%s stages %d ctas %ld smem %zu%s\n", d.tag,
                tiles_m, d.N, num_iters, BN, msub, cta_pair ? " CTA pairs" : "", tc.splits, csplit ? (" (cluster " + std::to_string(tc.cs) + " x " + std::to_string(groups) + " groups)").c_str() : "", stages, ctas, smem_bytes,
                via_scratch ? " (via scratch)" : "");

    p.a_c0 = d.a_c0; p.a_hoff = d.a_hoff; p.a_zmode = d.a_zmode;
    p.b_c0 = d.b_c0; p.b_hoff = d.b_hoff; p.b_zmode = d.b_zmode;
    p.idesc = ptx::make_idesc_f16(cta_pair ? 2 * kBlockM : kBlockM, BN, d.bf16, 0, 0);
    p.tx_bytes = (uint32_t)(msub * p.tw * p.th * p.tb * kBlockK * 2) + (uint32_t)(p.b_rows * kBlockK * 2);
    if (cta_pair) p.tx_bytes *= 2;          // the pair's boxes all complete on the even CTA's barrier
    p.alpha = 1.f;
    p.bias = d.bias;
    p.rowvec = d.rowvec;
    p.relu = d.relu;
    p.qscale = d.qscale;
    p.qinv = d.qscale != 0.f ? 1.f / d.qscale : 0.f;
    S2I_TRY(build_operand_maps(p, d, p.b_rows));
    if (csplit) {
        // the reducing CTAs address global memory directly (rows of the tile's pixel box)
        p.residual = d.residual; p.res_ld = d.res_ld;
        p.out32 = d.out32; p.ld32 = d.ld32;
        p.out16 = d.out16; p.ld16 = d.ld16;
    } else {
        if (d.residual) S2I_TRY(build_out_map(&p.mapRes, p, d, d.residual, d.res_ld, 2));
        if (d.out32) S2I_TRY(build_out_map(&p.mapO32, p, d, d.out32, d.ld32, 2));
        if (d.out16) S2I_TRY(build_out_map(&p.mapO16, p, d, d.out16, d.ld16, 0));
    }
    if (glu) {
        GemmDesc dg = d;
        dg.N = d.N / 2;
        S2I_TRY(build_out_map(&p.mapGlu, p, dg, d.out_glu, d.ld_glu, 0));
    }
    if (p.split_add) {
        const long rows = (long)d.aW * d.aH * d.aB;
        const long n4 = d.N / 4, total = rows * n4;
        bool planned = false;
        const bool shared_ws = via_scratch && !d_in.scratch32;      // the library's scratch is reused within a step: never planned
        if (g_zero_mode == ZeroMode::kApply && g_zero_plan && !shared_ws) {
            for (const ZeroRange& z : *g_zero_plan)
                if (z.p == d.out32 && z.ld == (long)d.ld32 && z.rows == rows && z.n4 == (int)n4) {
                    planned = true;      // zeroed by the plan's single launch at the top of the captured step
                    break;
                }
        }
        if (!planned) {
            S2I_LAUNCH((zero_rows_kernel), (unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, stream,
                       d.out32, (long)d.ld32, rows, (int)n4);
            S2I_LAUNCH_CHECK_TAG("gemm_split_zero", 0.0, 0.0);
            if (g_zero_mode == ZeroMode::kRecord && g_zero_plan && !shared_ws)
                g_zero_plan->push_back(ZeroRange{d.out32, (long)d.ld32, rows, (int)n4});
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        S2I_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        S2I_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        S2I_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        S2I_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        S2I_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        S2I_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)grid_m, (unsigned)tiles_n, (unsigned)tc.splits);
    p.trace = g_trace;
    // two epilogue warp quartets whenever the tile has at least two 32-column chunks (S2I_GEMM_ESETS=1: always one);
    // read per call so tools can A/B in one process
    int esets = (BN >> 5) >= 2 ? 2 : 1;
    if (const char* e = getenv("S2I_GEMM_ESETS")) esets = atoi(e) == 1 ? 1 : esets;
    if (cta_pair) {
        const int cz = csplit ? tc.cs : 1;
        if (glu) launch_kernel_cluster_z(gemm_tma_kernel<2, true, true>, grid, dim3(kThreads + 128), smem_bytes, 2, cz, stream, p);
        else if (esets == 2) launch_kernel_cluster_z(gemm_tma_kernel<2, false, true>, grid, dim3(kThreads + 128), smem_bytes, 2, cz, stream, p);
        else launch_kernel_cluster_z(gemm_tma_kernel<1, false, true>, grid, dim3(kThreads), smem_bytes, 2, cz, stream, p);
    } else if (cluster_ctas > 1) {
        if (esets == 2) launch_kernel_cluster_z(gemm_tma_kernel<2, false>, grid, dim3(kThreads + 128), smem_bytes, 1, tc.cs, stream, p);
        else launch_kernel_cluster_z(gemm_tma_kernel<1, false>, grid, dim3(kThreads), smem_bytes, 1, tc.cs, stream, p);
    } else if (glu) S2I_LAUNCH((gemm_tma_kernel<2, true>), grid, kThreads + 128, smem_bytes, stream, p);
    else if (esets == 2) S2I_LAUNCH((gemm_tma_kernel<2, false>), grid, kThreads + 128, smem_bytes, stream, p);
    else S2I_LAUNCH((gemm_tma_kernel<1, false>), grid, kThreads, smem_bytes, stream, p);
    const double m_rows = (double)d.aW * d.aH * d.aB;
    // algorithmic bytes: the activation operand and the weights once (fp16), the fp32 residual, the outputs
    const double alg_bytes = 2.0 * (m_rows * d.Kc + (double)d.N * d.Kc * d.taps) + m_rows * d.N * ((d.residual ? 4.0 : 0.0) +
                             (d_in.out32 ? 4.0 : 0.0) + (d_in.out16 ? 2.0 : 0.0) + (d_in.out_glu ? 1.0 : 0.0));
    S2I_LAUNCH_CHECK_TAG(d.tag, 2.0 * m_rows * d.N * d.Kc * d.taps, alg_bytes);
    if (via_scratch) {
        const long total = rows_total * (d.N / 4);
        S2I_LAUNCH((cast_rows_kernel), (unsigned)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048), 256, 0, stream, 
            d.out32, d.N, static_cast<__half*>(d_in.out16), d_in.ld16, rows_total, d.N / 4);
        S2I_LAUNCH_CHECK_TAG("gemm_split_cast", 0.0, 0.0);
    }
    return 0;
}

}  // namespace

int encode_tmap_f16(CUtensorMap* m, int rank, const void* ptr, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
    return encode_map(m, 0, rank, ptr, dims, strides_bytes, box);
}

long g_launches = 0;
long gemm_launch_count() { return g_launches; }

void gemm_zero_plan(ZeroMode mode, std::vector<ZeroRange>* plan) {
    g_zero_mode = mode;
    g_zero_plan = plan;
}

int gemm_zero_ranges(const ZeroRange* ranges, int n, cudaStream_t stream) {
    if (n <= 0) return 0;
    // the ranges differ 20x in size: enough blocks per range that the large ones are not the tail
    S2I_LAUNCH((zero_ranges_kernel), dim3(148, (unsigned)n), 256, 0, stream, ranges);
    S2I_LAUNCH_CHECK_TAG("gemm_split_zero", 0.0, 0.0);
    return 0;
}
void gemm_set_tma_epilogue(int on) { g_tma_epi = on ? 1 : 0; }
bool gemm_split_add_mode() { return g_split_add_mode; }
void gemm_set_trace(unsigned long long* buf) { g_trace = buf; }
void gemm_force_msub(int msub) { g_force_msub = msub; }
void gemm_set_pair(int mode) { g_pair_mode = mode; }

int gemm_launch(const GemmDesc& d, cudaStream_t stream) {
    if (!d.A || !d.B) return set_error(S2I_ERR_ARG, "gemm: null operand");
    if (d.taps != 1 && d.taps != 9) return set_error(S2I_ERR_ARG, "gemm: taps must be 1 or 9");
    if (d.a_mn && d.taps != 1) return set_error(S2I_ERR_ARG, "gemm: MN-major A has no taps");
    if (d.taps == 9 && (d.Kc % kBlockK) != 0) return set_error(S2I_ERR_ARG, "gemm: conv3x3 needs Cin %% 64 == 0");
    if (!d.out32 && !d.out16 && !d.out_glu) return set_error(S2I_ERR_ARG, "gemm: no output");

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.a_mn = d.a_mn;
    p.b_mn = d.b_mn;
    p.taps = d.taps;
    p.N = d.N;
    p.k_chunks = ceil_div(d.Kc, kBlockK);
    p.k_last_steps = ceil_div(d.Kc - (p.k_chunks - 1) * kBlockK, 16);
    p.W = d.aW;
    p.H = d.aH;
    p.Bn = d.aB;
    p.zh = d.zh > 0 ? d.zh : 1;

    // ---- M tiling
    long tiles_m;
    if (!d.a_mn) {
        int tw;
        if (d.aW >= kBlockM) {
            tw = kBlockM;
        } else {
            tw = pow2_floor(d.aW);
            if (d.aW % tw != 0) {
                int t = tw;
                while (t > 1 && d.aW % t != 0) t /= 2;
                if (t >= 4) tw = t;   // prefer an exact divisor when it is not tiny
            }
        }
        int th = pow2_floor(kBlockM / tw);
        if (th > d.aH) th = pow2_floor(d.aH) < d.aH ? pow2_floor(d.aH) * 2 : pow2_floor(d.aH);
        if (th > kBlockM / tw) th = kBlockM / tw;
        int tb = kBlockM / (tw * th);
        if (tb > d.aB) tb = d.aB;
        if (tb < 1) tb = 1;
        p.tw = tw;
        p.th = th;
        p.tb = tb;
        for (p.tw_shift = 0; (1 << p.tw_shift) < tw; ++p.tw_shift) {}
        for (p.th_shift = 0; (1 << p.th_shift) < th; ++p.th_shift) {}
        p.tiles_x = ceil_div(d.aW, tw);
        p.tiles_y = ceil_div(d.aH, th);
        const int tiles_b = d.Z > 1 ? 1 : ceil_div(d.aB, tb);
        if (d.Z > 1) {
            p.tb = 1;
            p.Bn = 1 << 30;   // batch comes from z; no batch masking
        }
        tiles_m = (long)p.tiles_x * p.tiles_y * tiles_b;
        p.M = 0;
    } else {
        p.M = d.aC;
        tiles_m = ceil_div(d.aC, kBlockM);
        p.tw = kBlockM;
        p.th = p.tb = 1;
        p.tw_shift = 7;
        p.th_shift = 0;
        p.tiles_x = (int)tiles_m;
        p.tiles_y = 1;
    }
    const int Z = d.Z > 0 ? d.Z : 1;

    const int num_iters = p.taps * p.k_chunks;
    // coalesced epilogue needs 16-byte aligned rows; everything on the sampling path qualifies
    const long zo_align = (d.c_sb | d.c_sh);
    bool fast = (d.N % 4 == 0) && (zo_align % 4 == 0);
    if (d.out32) fast = fast && (d.ld32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.out32) & 15) == 0);
    if (d.out16) fast = fast && (d.ld16 % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.out16) & 7) == 0);
    if (d.residual) fast = fast && (d.res_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.residual) & 15) == 0);
    if (d.bias) fast = fast && ((reinterpret_cast<uintptr_t>(d.bias) & 15) == 0);
    if (d.rowvec) fast = fast && (d.rowvec_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.rowvec) & 15) == 0);
    p.fast_epi = fast ? 1 : 0;

    // Epilogue through TMA (gemm_tma_kernel) whenever the shape allows: K-major operands, one problem, plain
    // bias / shared per-column vector / fp32 residual epilogue, 32-column boxes, 16-byte aligned rows.
    bool tma = tma_epilogue_enabled() && !d.a_mn && !d.b_mn && Z == 1 && d.alpha == 1.f && d.N % 32 == 0 && !d.out16_bf16 && !(d.rowvec && d.rowvec_ld != 0) && d.c_sb == 0 && d.c_sh == 0 &&
               d.a_hoff == 0 && d.b_hoff == 0;
    if (d.out32) tma = tma && (d.ld32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.out32) & 15) == 0);
    if (d.out16) tma = tma && (d.ld16 % 8 == 0) && ((reinterpret_cast<uintptr_t>(d.out16) & 15) == 0);
    if (d.residual) tma = tma && (d.res_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.residual) & 15) == 0);
    if (d.out_glu) {
        if (!tma || d.N % 64 != 0 || d.residual || d.out32 || d.relu || d.qscale != 0.f || (d.ld_glu % 8) != 0 ||
            (reinterpret_cast<uintptr_t>(d.out_glu) & 15) != 0)
            return set_error(S2I_ERR_ARG, "gemm: the gated-GELU epilogue needs K-major operands, Z = 1, N %% 64 == 0, aligned fp16 "
                                          "outputs and no residual / fp32 output / ReLU");
    }
    if (tma) return launch_tma(p, d, tiles_m, num_iters, stream);

    TileChoice tc = choose_tiles(d.N, tiles_m, Z, num_iters, fast && Z == 1 && d.splits >= 0);
    if (d.BN > 0) tc.BN = d.BN;
    if (d.splits > 0) tc.splits = d.splits;
    if (!fast || Z != 1) tc.splits = 1;
    int BN = tc.BN;
    if (BN % 16 != 0 || BN < 16 || BN > 256) return set_error(S2I_ERR_ARG, "gemm: BN %d invalid", BN);
    p.BN = BN;
    if (tc.splits > 1) {
        const long tiles = tiles_m * ceil_div(d.N, BN);
        while (tc.splits > 1 && ((long)(tc.splits - 1) * ceil_div(num_iters, tc.splits) >= num_iters ||
                                 (size_t)tc.splits * tiles * kBlockM * BN * sizeof(float) > kWsBytes ||
                                 tiles > kMaxSplitTiles))
            --tc.splits;
    }
    p.splits = tc.splits;
    p.iters_per_split = ceil_div(num_iters, tc.splits);
    if (tc.splits > 1) {
        S2I_TRY(ensure_ws());
        p.ws = g_ws;
        p.counters = g_counters;
    }
    const int tiles_n = ceil_div(d.N, BN);
    p.tmem_cols = 32;
    while (p.tmem_cols < BN) p.tmem_cols *= 2;

    const int b_stage_bytes = ceil_div(BN, 64) * kChunkBytes;
    const int stage_bytes = kAStageBytes + b_stage_bytes;
    int stages = (100 * 1024) / stage_bytes;   // <= ~100 KB so two CTAs can share an SM
    if (stages < 3) stages = 3;
    if (stages > 6) stages = 6;
    if (stages > p.iters_per_split) stages = p.iters_per_split < 1 ? 1 : p.iters_per_split;
    p.stages = stages;
    // the epilogue reuses stage memory for its 4 x (32 x 36) fp32 transpose buffers (+ flag): keep at least 19 KB
    size_t pipe_bytes = (size_t)stages * stage_bytes;
    if (pipe_bytes < 19 * 1024) pipe_bytes = 19 * 1024;
    const size_t smem_bytes = pipe_bytes + (2 * stages + 1) * 8 + 16 + 1024;

    p.a_c0 = d.a_c0;
    p.a_hoff = d.a_hoff;
    p.a_zmode = d.a_zmode;
    p.b_c0 = d.b_c0;
    p.b_hoff = d.b_hoff;
    p.b_zmode = d.b_zmode;
    p.idesc = ptx::make_idesc_f16(kBlockM, BN, d.bf16, d.a_mn, d.b_mn);
    const uint32_t a_bytes = d.a_mn ? 2 * kChunkBytes : (uint32_t)(p.tw * p.th * p.tb * kBlockK * 2);
    const uint32_t b_bytes = d.b_mn ? (uint32_t)(ceil_div(BN, 64) * kChunkBytes) : (uint32_t)(BN * kBlockK * 2);
    p.tx_bytes = a_bytes + b_bytes;

    p.alpha = d.alpha;
    p.bias = d.bias;
    p.rowvec = d.rowvec;
    p.rowvec_ld = d.rowvec_ld;
    p.residual = d.residual;
    p.res_ld = d.res_ld;
    p.out32 = d.out32;
    p.ld32 = d.ld32;
    p.out16 = d.out16;
    p.ld16 = d.ld16;
    p.out16_bf16 = d.out16_bf16;
    p.c_sb = d.c_sb;
    p.c_sh = d.c_sh;
    p.relu = d.relu;
    p.qscale = d.qscale;
    p.qinv = d.qscale != 0.f ? 1.f / d.qscale : 0.f;

    S2I_TRY(build_operand_maps(p, d, BN));

    static bool attr_set = false;
    if (!attr_set) {
        S2I_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)tiles_m, (unsigned)tiles_n, (unsigned)(Z * p.splits));
    S2I_LAUNCH((gemm_tc_kernel), grid, kThreads, smem_bytes, stream, p);
    const double m_rows = d.a_mn ? (double)d.aC : (double)d.aW * d.aH * (d.Z > 1 ? 1 : d.aB);
    const double alg_bytes = Z * (2.0 * (m_rows * d.Kc + (double)d.N * d.Kc * d.taps) + m_rows * d.N * ((d.residual ? 4.0 : 0.0) +
                             (d.out32 ? 4.0 : 0.0) + (d.out16 ? 2.0 : 0.0)));
    S2I_LAUNCH_CHECK_TAG(d.tag, 2.0 * m_rows * d.N * d.Kc * d.taps * Z, alg_bytes);
    return 0;
}

}  // namespace s2i
