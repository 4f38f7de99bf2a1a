// Micro-benchmark: the chain CTA of walk_stream_kernel in isolation.  Warp 0 adds (one column per lane, stages of
// 128 messages x 16 floats), warp 1 (one lane) refills an 11-stage ring with 8 KB bulk copies from a product buffer in
// global memory and waits on the `empty` barriers.  Reports cycles per message of the consumer for a few variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/stream_chain.bin scripts/micro/stream_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kStages = 11, kMsgs = 128, kCols = 16;
constexpr int kStageFloats = kMsgs * kCols;       // 8 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int SLEEP>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (;;) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (SLEEP > 0) __nanosleep(SLEEP);
    }
}
__device__ __forceinline__ void mbar_wait_ptx(uint64_t* bar, uint32_t parity) {      // the product's wait loop
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// MODE 0: consumer as shipped (wait / 128 x (LDS, FADD) / arrive), loader spinning on try_wait
// MODE 1: same, loader backs off 100 ns between polls
// MODE 2: consumer only: no loader, no barriers, the ring is filled once (the floor inside a CTA of this shape)
// MODE 3: as 0, but the consumer also loads the first 16 values of the NEXT stage before the arrive (pipeline never drains)
template <int MODE>
__global__ void __launch_bounds__(256) chain_cta(const float* __restrict__ prod0, int nstage, float* out, long long* cyc,
                                                  size_t region_stages) {
    const float* prod = prod0 + (size_t)blockIdx.x * region_stages * kStageFloats;
    extern __shared__ __align__(128) unsigned char raw[];
    float* ring = reinterpret_cast<float*>(raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kStages * kStageFloats);
    uint64_t* empty = full + kStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (MODE == 2) for (int i = threadIdx.x; i < kStages * kStageFloats; i += 256) ring[i] = prod[i];
    __syncthreads();
    if (warp == 1 && MODE != 2) {
        if (lane == 0) {
            for (int b = 0; b < nstage; ++b) {
                const int use = b / kStages, stage = b - use * kStages;
                if (use > 0) { if (MODE == 4) mbar_wait_ptx(&empty[stage], (use - 1) & 1); else mbar_wait<MODE == 1 ? 100 : 0>(&empty[stage], (use - 1) & 1); }
                mbar_expect_tx(&full[stage], kStageFloats * 4);
                bulk_g2s(ring + stage * kStageFloats, prod + (size_t)((size_t)b % region_stages) * kStageFloats, kStageFloats * 4, &full[stage]);
            }
        }
    } else if (warp == 0) {
        float acc = 0.f;
        const long long t0 = clock64();
        bool next_full = false;
        for (int b = 0; b < nstage; ++b) {
            const int use = b / kStages, stage = b - use * kStages;
            if (MODE != 2) {
                if (!next_full) { if (MODE == 4) mbar_wait_ptx(&full[stage], use & 1); else mbar_wait<0>(&full[stage], use & 1); }
                next_full = false;
            }
            const float* xs = ring + stage * kStageFloats + lane;
            float va[32], vb[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) va[j] = xs[j * kCols];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float* cur = (q & 1) ? vb : va;
                float* nxt = (q & 1) ? va : vb;
                if (q + 1 < 4) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) nxt[j] = xs[(32 * (q + 1) + j) * kCols];
                } else if (MODE != 2 && b + 1 < nstage) {
                    const int use1 = (b + 1) / kStages;
                    next_full = mbar_test(&full[(b + 1) - use1 * kStages], use1 & 1);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) acc = __fadd_rn(acc, cur[j]);
            }
            if (MODE != 2) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
        }
        const long long t1 = clock64();
        out[lane] = acc;
        if (lane == 0) cyc[blockIdx.x] = t1 - t0;
    }
    if (MODE == 4) __syncthreads();          // warps 2..7 are parked here while warps 0 and 1 work (as in the product)
}

template <int MODE>
void run(const char* name, const float* prod, float* out, long long* cyc, int nstage, int grid, size_t region_stages) {
    const int smem = kStages * (kStageFloats * 4 + 16) + 16;
    cudaFuncSetAttribute(chain_cta<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[256];
    for (int rep = 0; rep < 2; ++rep) {
        chain_cta<MODE><<<grid, 256, smem>>>(prod, nstage, out, cyc, region_stages);
        cudaMemcpy(h, cyc, 8 * grid, cudaMemcpyDeviceToHost);
    }
    long long worst = 0;
    for (int i = 0; i < grid; ++i) worst = h[i] > worst ? h[i] : worst;
    printf("%-58s grid %3d, %5zu stages per CTA region: %.2f cycles/message (%s)\n", name, grid, region_stages,
           (double)worst / ((double)nstage * kMsgs), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *prod, *out;
    long long* cyc;
    const size_t total_stages = 262144;                       // 2 GB of products
    cudaMalloc(&prod, total_stages * kStageFloats * 4);
    cudaMemset(prod, 0, total_stages * kStageFloats * 4);
    cudaMalloc(&out, 4096);
    cudaMalloc(&cyc, 8 * 256);
    const int nstage = 2200;
    run<2>("consumer alone, ring filled once, no barriers", prod, out, cyc, nstage, 1, 4096);
    run<0>("loader + consumer as shipped, L2-resident products", prod, out, cyc, nstage, 1, 4096);
    run<1>("... loader backing off 100 ns between polls", prod, out, cyc, nstage, 1, 4096);
    run<0>("loader + consumer, every stage from HBM", prod, out, cyc, nstage, 1, 2200);
    run<4>("... other warps parked at the CTA barrier, PTX wait loops", prod, out, cyc, nstage, 1, 2200);
    run<0>("42 chain CTAs, every stage from HBM", prod, out, cyc, nstage, 42, 2200);
    run<0>("148 chain CTAs (one per SM), every stage from HBM", prod, out, cyc, 1700, 148, 1700);
    return 0;
}
