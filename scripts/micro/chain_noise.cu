// Micro-benchmark: how much does an add chain (warp 0) slow down when other warps of the CTA are busy,
// and does it matter which warps (warp % 4 == 0 share the chain's sub-partition if the mapping is w % 4)?
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kMsgs = 128, kStages = 64;

__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ in, float* out, long long* cyc, unsigned noise_mask, int noise_kind) {
    __shared__ float buf[kMsgs * 16];
    __shared__ __align__(16) float scratch[16][32 * 4 * 2];
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kMsgs * 16; i += blockDim.x) buf[i] = in[i];
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        float acc = 0.f;
        const long long t0 = clock64();
        for (int s = 0; s < kStages; ++s) {
            const float* xs = buf + (lane & 15);
            for (int j0 = 0; j0 < kMsgs; j0 += 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc = __fadd_rn(acc, xs[(j0 + j) * 16]);
            }
        }
        const long long t1 = clock64();
        out[lane] = acc;
        if (lane == 0) { cyc[0] = t1 - t0; stop = 1; }
    } else if ((noise_mask >> warp) & 1u) {
        float4 v = make_float4(lane, 1.f, 2.f, 3.f);
        float4* p = reinterpret_cast<float4*>(&scratch[warp][0]) + lane;
        float a = lane;
        while (!stop) {
            if (noise_kind == 0) {            // shared-memory traffic + shuffles + fp32 math
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    *p = v;
                    v = p[32 * (i & 1)];
                    v.x = __fmul_rn(v.x, 1.0001f) + __shfl_xor_sync(0xffffffffu, v.y, 1);
                }
            } else if (noise_kind == 1) {     // pure ALU
#pragma unroll
                for (int i = 0; i < 32; ++i) a = fmaf(a, 1.0001f, 0.5f);
            } else {                          // global loads
                a += in[(lane * 97 + (int)a) & 4095];
            }
        }
        out[threadIdx.x] = v.x + a;
    }
}

int main() {
    float *in, *out;
    long long* cyc;
    cudaMalloc(&in, 1 << 16);
    cudaMemset(in, 0, 1 << 16);
    cudaMalloc(&out, 1 << 16);
    cudaMalloc(&cyc, 64);
    long long h = 0;
    const double n = (double)kMsgs * kStages;
    const unsigned all = 0xfffe, others = 0xeeee, same = 0x1110;   // warps 1..15 | warps with w%4 != 0 | warps 4, 8, 12
    const char* kinds[3] = {"smem+shfl+fmul", "pure ALU", "global loads"};
    for (int kind = 0; kind < 3; ++kind) {
        const unsigned masks[4] = {0u, all, others, same};
        const char* names[4] = {"no noise", "warps 1..15 busy", "warps w%4!=0 busy (12)", "warps 4,8,12 busy (3)"};
        for (int m = 0; m < 4; ++m) {
            k<<<1, 512>>>(in, out, cyc, masks[m], kind);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-16s %-26s : %.2f cycles/message\n", kinds[kind], names[m], h / n);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
