// Micro-benchmark behind the sharded data plane: what does a kernel on GPU 0 get out of / into
// GPU 1's HBM over NVLink, for the access shapes the routed exchange can use?
//   copy      : contiguous 16-byte loads from the peer (upper bound of a pull)
//   pull rows : one warp per random row block (3,456 B), 4 x 16-byte requests in flight per lane — tpn_pull_rows' shape
//   pull deep : same rows, the whole block requested before the first use (7 requests in flight per lane)
//   push rows : the owner writes random row blocks INTO the peer (stores are posted: no round trip per request)
// Build + run on a box with >= 2 GPUs:  nvcc -O3 -arch=sm_100a scripts/micro/peer_paths.cu -o /tmp/peer_paths && /tmp/peer_paths
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kRow4 = 216;      // float4 per row block: 4 layers x 216 floats

__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

template <int DEPTH>
__global__ void gather_rows(const float4* __restrict__ src, float4* __restrict__ dst, const int* __restrict__ rows, int n) {
    const int lane = threadIdx.x & 31;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n; s += warps) {
        const float4* p = src + (size_t)rows[s] * kRow4;
        float4* q = dst + (size_t)s * kRow4;
        for (int c0 = lane; c0 < kRow4; c0 += 32 * DEPTH) {
            float4 v[DEPTH];
#pragma unroll
            for (int k = 0; k < DEPTH; ++k) if (c0 + 32 * k < kRow4) v[k] = p[c0 + 32 * k];
#pragma unroll
            for (int k = 0; k < DEPTH; ++k) if (c0 + 32 * k < kRow4) q[c0 + 32 * k] = v[k];
        }
    }
}

// local rows -> remote slots
__global__ void scatter_rows(const float4* __restrict__ src, float4* __restrict__ dst, const int* __restrict__ rows, int n) {
    const int lane = threadIdx.x & 31;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n; s += warps) {
        const float4* p = src + (size_t)rows[s] * kRow4;
        float4* q = dst + (size_t)s * kRow4;
        for (int c0 = lane; c0 < kRow4; c0 += 128) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) if (c0 + 32 * k < kRow4) v[k] = p[c0 + 32 * k];
#pragma unroll
            for (int k = 0; k < 4; ++k) if (c0 + 32 * k < kRow4) q[c0 + 32 * k] = v[k];
        }
    }
}

template <class F>
float timed(F f, int reps = 10) {
    f();
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
    const size_t state_rows = 2000000;                       // 6.9 GB per GPU
    const size_t bytes = state_rows * kRow4 * 16;
    float4 *local, *remote, *stage_local, *stage_remote;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, bytes)); CK(cudaMemset(remote, 1, bytes));
    const int n = 100000;
    CK(cudaMalloc(&stage_remote, (size_t)n * kRow4 * 16));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&local, bytes)); CK(cudaMemset(local, 2, bytes));
    CK(cudaMalloc(&stage_local, (size_t)n * kRow4 * 16));
    int* h = (int*)malloc(n * sizeof(int));
    srand(1);
    for (int i = 0; i < n; ++i) h[i] = (int)(((size_t)rand() * 2654435761u) % state_rows);
    int* rows;
    CK(cudaMalloc(&rows, n * sizeof(int)));
    CK(cudaMemcpy(rows, h, n * sizeof(int), cudaMemcpyHostToDevice));
    const double gb = (double)n * kRow4 * 16 / 1e9;
    const int grid = 148 * 4;
    float ms;
    ms = timed([&] { copy_kernel<<<148 * 8, 256>>>(remote, local, (size_t)n * kRow4); });
    printf("copy   peer -> local, contiguous        : %8.1f GB/s\n", gb / (ms * 1e-3));
    ms = timed([&] { copy_kernel<<<148 * 8, 256>>>(local, stage_remote, (size_t)n * kRow4); });
    printf("copy   local -> peer, contiguous        : %8.1f GB/s\n", gb / (ms * 1e-3));
    for (int cnt : {10000, 100000}) {
        const double g = (double)cnt * kRow4 * 16 / 1e9;
        ms = timed([&] { gather_rows<4><<<grid, 256>>>(local, stage_local, rows, cnt); });
        printf("%6d rows  local gather (depth 4)        : %8.1f GB/s  %7.1f us\n", cnt, g / (ms * 1e-3), ms * 1e3);
        ms = timed([&] { gather_rows<4><<<grid, 256>>>(remote, stage_local, rows, cnt); });
        printf("%6d rows  PULL from peer (depth 4)      : %8.1f GB/s  %7.1f us\n", cnt, g / (ms * 1e-3), ms * 1e3);
        ms = timed([&] { gather_rows<7><<<grid, 256>>>(remote, stage_local, rows, cnt); });
        printf("%6d rows  PULL from peer (depth 7)      : %8.1f GB/s  %7.1f us\n", cnt, g / (ms * 1e-3), ms * 1e3);
        ms = timed([&] { gather_rows<7><<<148 * 8, 256>>>(remote, stage_local, rows, cnt); });
        printf("%6d rows  PULL depth 7, 8 CTAs per SM   : %8.1f GB/s  %7.1f us\n", cnt, g / (ms * 1e-3), ms * 1e3);
        ms = timed([&] { scatter_rows<<<grid, 256>>>(local, stage_remote, rows, cnt); });
        printf("%6d rows  PUSH into peer                : %8.1f GB/s  %7.1f us\n", cnt, g / (ms * 1e-3), ms * 1e3);
    }
    return 0;
}
