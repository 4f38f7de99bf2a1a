// Micro-benchmark: cycles per message of a sequential fp32 add chain fed from shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/chain_bench scripts/micro/chain_bench.cu && /tmp/chain_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kMsgs = 128;          // messages per stage
constexpr int kStages = 64;

// (a) layout [msg][16 floats], one LDS.32 + one FADD per message, as the compiler schedules it
__global__ void chain_lds32(const float* __restrict__ in, float* out, long long* cyc) {
    __shared__ float buf[kMsgs * 16];
    for (int i = threadIdx.x; i < kMsgs * 16; i += 32) buf[i] = in[i];
    __syncwarp();
    const int lane = threadIdx.x;
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
    if (lane == 0) cyc[0] = t1 - t0;
}

// (b) transposed layout [col][130], LDS.64 (2 messages), next block of 32 loaded before the adds
__global__ void chain_lds64_db(const float* __restrict__ in, float* out, long long* cyc) {
    __shared__ __align__(16) float buf[16 * 130];
    for (int i = threadIdx.x; i < 16 * 130; i += 32) buf[i] = in[i];
    __syncwarp();
    const int lane = threadIdx.x;
    float acc = 0.f;
    const float2* col = reinterpret_cast<const float2*>(buf + (lane & 15) * 130);
    const long long t0 = clock64();
    float2 a[16], b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = col[j];
    for (int s = 0; s < kStages; ++s) {
#pragma unroll
        for (int blk = 0; blk < 4; blk += 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) b[j] = col[(blk + 1) * 16 + j];
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc = __fadd_rn(acc, a[j].x); acc = __fadd_rn(acc, a[j].y); }
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = col[((blk + 2) & 3) * 16 + j];
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc = __fadd_rn(acc, b[j].x); acc = __fadd_rn(acc, b[j].y); }
        }
    }
    const long long t1 = clock64();
    out[lane] = acc;
    if (lane == 0) cyc[0] = t1 - t0;
}

// (c) pure register chain: the floor
__global__ void chain_regs(const float* __restrict__ in, float* out, long long* cyc) {
    const int lane = threadIdx.x;
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = in[j * 32 + lane];
    float acc = 0.f;
    const long long t0 = clock64();
    for (int s = 0; s < kStages * 4; ++s) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = __fadd_rn(acc, x[j]);
    }
    const long long t1 = clock64();
    out[lane] = acc;
    if (lane == 0) cyc[0] = t1 - t0;
}

// (d) layout [msg][16 floats] as (a), but explicitly double-buffered in registers (32 LDS.32 ahead)
__global__ void chain_lds32_db(const float* __restrict__ in, float* out, long long* cyc) {
    __shared__ float buf[kMsgs * 16];
    for (int i = threadIdx.x; i < kMsgs * 16; i += 32) buf[i] = in[i];
    __syncwarp();
    const int lane = threadIdx.x;
    float acc = 0.f;
    const float* xs = buf + (lane & 15);
    const long long t0 = clock64();
    float a[32], b[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = xs[j * 16];
    for (int s = 0; s < kStages; ++s) {
#pragma unroll
        for (int blk = 0; blk < 4; blk += 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) b[j] = xs[((blk + 1) * 32 + j) * 16];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc = __fadd_rn(acc, a[j]);
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = xs[(((blk + 2) & 3) * 32 + j) * 16];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc = __fadd_rn(acc, b[j]);
        }
    }
    const long long t1 = clock64();
    out[lane] = acc;
    if (lane == 0) cyc[0] = t1 - t0;
}

// (e) transposed layout [col][128 messages], 16-byte slots swizzled by the column (slot ^ (col & 7)): one LDS.128 per
// 4 messages, conflict-free for every quarter-warp; 8 loads (32 messages) ahead
__global__ void chain_lds128_t(const float* __restrict__ in, float* out, long long* cyc) {
    __shared__ __align__(16) float buf[16 * kMsgs];
    for (int i = threadIdx.x; i < 16 * kMsgs; i += 32) buf[i] = in[i];
    __syncwarp();
    const int lane = threadIdx.x;
    const int c = lane & 15;
    float acc = 0.f;
    const float4* col = reinterpret_cast<const float4*>(buf + c * kMsgs);
    const long long t0 = clock64();
    float4 a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = col[j ^ (c & 7)];
    for (int s = 0; s < kStages; ++s) {
#pragma unroll
        for (int blk = 0; blk < 4; blk += 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = col[((blk + 1) * 8 + j) ^ (c & 7)];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc = __fadd_rn(acc, a[j].x); acc = __fadd_rn(acc, a[j].y); acc = __fadd_rn(acc, a[j].z); acc = __fadd_rn(acc, a[j].w);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = col[(((blk + 2) & 3) * 8 + j) ^ (c & 7)];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc = __fadd_rn(acc, b[j].x); acc = __fadd_rn(acc, b[j].y); acc = __fadd_rn(acc, b[j].z); acc = __fadd_rn(acc, b[j].w);
            }
        }
    }
    const long long t1 = clock64();
    out[lane] = acc;
    if (lane == 0) cyc[0] = t1 - t0;
}

int main() {
    float *in, *out;
    long long* cyc;
    cudaMalloc(&in, 1 << 16);
    cudaMemset(in, 0, 1 << 16);
    cudaMalloc(&out, 4096);
    cudaMalloc(&cyc, 64);
    long long h = 0;
    const double n = (double)kMsgs * kStages;
    for (int rep = 0; rep < 2; ++rep) {
        chain_lds32<<<1, 32>>>(in, out, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("lds32 interleaved      : %.2f cycles/message\n", h / n);
        chain_lds32_db<<<1, 32>>>(in, out, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("lds32 double-buffered  : %.2f cycles/message\n", h / n);
        chain_lds64_db<<<1, 32>>>(in, out, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("lds64 transposed, dbuf : %.2f cycles/message\n", h / n);
        chain_lds128_t<<<1, 32>>>(in, out, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("lds128 transposed+swz  : %.2f cycles/message\n", h / n);
        chain_regs<<<1, 32>>>(in, out, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("register chain (floor) : %.2f cycles/message\n", h / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
