// Micro-benchmark for the hub walker's supply side (DESIGN.md section 4 "Hubs", section 9 item 2):
// how many 64-byte / 128-byte message slices per cycle can ONE SM pull from random rows of a large
// buffer into shared memory, by path:
//   ldg     : LDG.128 into registers + STS.128 (what walk_hub2_kernel's producers do), P warps
//   ldgsts  : cp.async 16 B (LDGSTS) straight into shared memory, P warps
//   bulk    : cp.async.bulk (TMA, UBLKCP) of one whole slice per lane, completion on an mbarrier, P warps
// One CTA per SM, every warp owns a private stage of 128 slices and refills it `iters` times.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_paths scripts/micro/gather_paths.cu && /tmp/gather_paths
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kMsgs = 128;                 // slices per stage

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// rows[i] = row index of message i (random); a slice = `slice_bytes` at column offset 0 of a row of `row_bytes`
template <int MODE>
__global__ void __launch_bounds__(512, 1)
gather_kernel(const char* __restrict__ data, size_t row_bytes, const uint32_t* __restrict__ rows, int slice_bytes,
              int iters, long long* cycles, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned char* stage = smem + (size_t)warp * kMsgs * slice_bytes;
    if (lane == 0) mbar_init(&bars[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t* myrows = rows + ((size_t)blockIdx.x * nw + warp) * (size_t)iters * kMsgs;
    const int lpm = slice_bytes / 16;                 // lanes per message (16-byte pieces): 4 or 8
    const int mpi = 32 / lpm;                         // messages per warp-wide instruction
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t* r = myrows + (size_t)it * kMsgs;
        if (MODE == 2) {
            if (lane == 0) mbar_expect_tx(&bars[warp], (uint32_t)(kMsgs * slice_bytes));
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kMsgs / 32; ++k) {
                const int m = k * 32 + lane;
                bulk_g2s(stage + (size_t)m * slice_bytes, data + (size_t)r[m] * row_bytes, (uint32_t)slice_bytes, &bars[warp]);
            }
            mbar_wait(&bars[warp], it & 1);
        } else {
            const int grp = lane / lpm, sub = lane % lpm;
            if (MODE == 1) {
                for (int i = 0; i < kMsgs / mpi; ++i) {
                    const int m = i * mpi + grp;
                    cp_async16(stage + (size_t)m * slice_bytes + sub * 16, data + (size_t)r[m] * row_bytes + sub * 16);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            } else {
                for (int i0 = 0; i0 < kMsgs / mpi; i0 += 16) {          // 16 loads in flight per lane, like the hub walker
                    float4 v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int m = (i0 + i) * mpi + grp;
                        v[i] = *reinterpret_cast<const float4*>(data + (size_t)r[m] * row_bytes + sub * 16);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int m = (i0 + i) * mpi + grp;
                        *reinterpret_cast<float4*>(stage + (size_t)m * slice_bytes + sub * 16) = v[i];
                    }
                }
            }
        }
        __syncwarp();
        acc += *reinterpret_cast<const float*>(stage + lane * 4);      // consume something
        __syncwarp();
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t row_bytes = 3456, nrows = 3000000;                   // ~10 GB of rows: every access a DRAM / TLB miss
    char* data;
    if (cudaMalloc(&data, row_bytes * nrows) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(data, 0, row_bytes * nrows);
    const int iters = 64, maxw = 16;
    const size_t nidx = (size_t)sms * maxw * iters * kMsgs;
    uint32_t* h = (uint32_t*)malloc(nidx * 4);
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < nidx; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = (uint32_t)(x % nrows); }
    uint32_t* rows;
    cudaMalloc(&rows, nidx * 4);
    cudaMemcpy(rows, h, nidx * 4, cudaMemcpyHostToDevice);
    long long* cyc;
    float* sink;
    cudaMalloc(&cyc, sms * 8);
    cudaMalloc(&sink, (size_t)sms * 512 * 4);
    long long hc[256];
    const char* names[3] = {"ldg+sts", "ldgsts ", "bulk   "};
    for (int slice = 64; slice <= 128; slice *= 2) {
        for (int nw = 4; nw <= 16; nw *= 2) {
            const int smem = nw * kMsgs * slice;
            for (int mode = 0; mode < 3; ++mode) {
                auto k = mode == 0 ? gather_kernel<0> : (mode == 1 ? gather_kernel<1> : gather_kernel<2>);
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                for (int rep = 0; rep < 2; ++rep) k<<<sms, nw * 32, smem>>>(data, row_bytes, rows, slice, iters, cyc, sink);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed: %s\n", names[mode], cudaGetErrorString(cudaGetLastError())); return 1; }
                cudaMemcpy(hc, cyc, sms * 8, cudaMemcpyDeviceToHost);
                double mean = 0;
                for (int i = 0; i < sms; ++i) mean += (double)hc[i];
                mean /= sms;
                const double msgs = (double)nw * iters * kMsgs;
                printf("slice %3d B  %2d warps  %s : %6.2f cycles/message/SM  %6.1f B/cycle/SM  (all %d SMs busy)\n", slice, nw,
                       names[mode], mean / msgs, msgs * slice / mean, sms);
            }
        }
    }
    return 0;
}
