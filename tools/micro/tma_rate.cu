// Microbenchmark (B200): how fast can ONE CTA pull bytes from L2 / HBM into shared memory with bulk async copies, as a function of the
// bytes in flight (ring depth x copy size), the number of co-resident CTAs per SM, and the number of issuing threads?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu ; run: ./tma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_load(void* s, const void* g, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s)), "l"(g), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// Each CTA streams `iters` copies of `bytes` bytes through a ring of `depth` slots; `nthr` warps each run their own ring (separate
// issuing threads).  A slot is re-armed as soon as its copy has landed (no consumer work).
__global__ void __launch_bounds__(128) rate_kernel(const uint8_t* src, size_t span, int bytes, int depth, int iters, int nthr, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[4][16];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 16; ++j) mbar_init(&bars[i][j], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    if (w < nthr && lane == 0) {
        uint8_t* ring = smem + (size_t)w * depth * bytes;
        size_t off = ((size_t)blockIdx.x * nthr + w) * (size_t)iters * bytes % span;
        int s = 0;
        uint32_t par = 1;                                   // parity of the PREVIOUS use of slot s (first pass: nothing to wait for)
        uint8_t* dst = ring;
        for (int i = 0; i < iters; ++i) {                   // no divisions in the issue loop (they cost more than the copy instruction)
            if (i >= depth) mbar_wait(&bars[w][s], par);
            mbar_expect(&bars[w][s], bytes);
            bulk_load(dst, src + off, bytes, &bars[w][s]);
            off += bytes;
            if (off + bytes > span) off = 0;
            dst += bytes;
            if (++s == depth) { s = 0; dst = ring; par ^= 1u; }
        }
        // drain: the last `depth` copies (slot order continues from s)
        const int tail = iters < depth ? iters : depth;
        for (int k = 0; k < tail; ++k) {
            mbar_wait(&bars[w][s], par);
            if (++s == depth) { s = 0; par ^= 1u; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t span_l2 = 64u << 20, span_hbm = 2048ull << 20;
    uint8_t* buf;
    cudaMalloc(&buf, span_hbm);
    cudaMemset(buf, 1, span_hbm);
    long long* cyc;
    cudaMalloc(&cyc, 4096 * sizeof(long long));
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("%-6s %5s %6s %6s %5s %5s | %10s %12s %10s\n", "where", "ctas", "bytes", "depth", "nthr", "perSM", "B/clk/CTA", "B/clk/SM", "TB/s chip");
    for (int hbm = 0; hbm < 2; ++hbm)
        for (int per_sm : {1, 2, 4})
            for (int nthr : {1, 2})
                for (int bytes : {4096, 16384, 32768})
                    for (int depth : {1, 2, 4, 8}) {
                        const size_t smem = (size_t)nthr * depth * bytes;
                        if (smem > 200u * 1024u / per_sm - 2048) continue;
                        const int ctas = 148 * per_sm;
                        const int iters = (8 << 20) / bytes / nthr;      // 8 MB per CTA
                        const size_t span = hbm ? span_hbm : span_l2;
                        // shared memory sized so that exactly per_sm CTAs fit an SM
                        const size_t smem_req = per_sm == 1 ? 120u * 1024u : (per_sm == 2 ? 100u * 1024u : 50u * 1024u);
                        const size_t smem_launch = smem > smem_req ? smem : smem_req;
                        for (int rep = 0; rep < 2; ++rep) rate_kernel<<<ctas, 128, smem_launch>>>(buf, span, bytes, depth, iters, nthr, cyc);
                        cudaEvent_t e0, e1;
                        cudaEventCreate(&e0); cudaEventCreate(&e1);
                        cudaEventRecord(e0);
                        rate_kernel<<<ctas, 128, smem_launch>>>(buf, span, bytes, depth, iters, nthr, cyc);
                        cudaEventRecord(e1);
                        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                        float ms; cudaEventElapsedTime(&ms, e0, e1);
                        long long h[4096]; cudaMemcpy(h, cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
                        double avg = 0; for (int i = 0; i < ctas; ++i) avg += (double)h[i]; avg /= ctas;
                        const double per_cta = (double)iters * nthr * bytes / avg;
                        printf("%-6s %5d %6d %6d %5d %5d | %10.1f %12.1f %10.2f\n", hbm ? "HBM" : "L2", ctas, bytes, depth, nthr, per_sm, per_cta, per_cta * per_sm,
                               (double)ctas * iters * nthr * bytes / (ms * 1e-3) / 1e12);
                    }
    return 0;
}
