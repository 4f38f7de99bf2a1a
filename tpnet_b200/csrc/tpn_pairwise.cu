// tpn_pairwise — input of `self.mlp` in RandomProjectionModule.get_pair_wise_feature
// (reference: models/TPNet.py:112-129) for sm_100a.
//
// Per pair (a, b) the reference stacks the R = 2L+2 rows [a:P_0..P_L, b:P_0..P_L]
// (TPNet.py:119-121), forms the R x R Gram matrix with a batched fp32 GEMM (:122-123),
// clamps at 0 and takes log(x + 1.0) (:127-128).  The Gram blocks are 4x4 .. 10x10 over
// d ~ 100-200: ~3 flop/byte, far below the fp32 ridge, so this is a gather-bound kernel
// and tensor cores are deliberately not used.
//
// Mapping: G = 8 lanes per pair, 4 pairs per warp.  Lane j of a group reads float4
// columns j, j+8, ... of all R rows (the group reads 128 contiguous bytes per row per
// step; a node's L+1 rows are one contiguous block in the node-major state), keeps the
// R(R+1)/2 unique dot products in registers, reduces them over the 8 lanes with shuffles,
// applies the epilogue once per unique entry and mirrors it through shared memory so the
// warp writes its 4 * R*R outputs as coalesced 128-bit streaming stores.
// Lazy-decay mode: rows of layers >= 1 are brought current in registers (replay of the
// logged fp32 factors) before they enter the products; nothing is written back.
#include "tpn_common.cuh"

namespace tpn {
namespace {

constexpr int kPairThreads = 128;
constexpr int kGroup = 8;
constexpr int kPairsPerWarp = 32 / kGroup;
constexpr int kPairsPerBlock = kPairThreads / kGroup;

template <int LAYERS, bool LAZY>
__global__ void __launch_bounds__(kPairThreads)
pairwise_kernel(StateView st, const long long* __restrict__ a_ids, const long long* __restrict__ b_ids,
                long long n, int apply_log_scale, float* __restrict__ out, int ds4) {
    constexpr int H = LAYERS + 1;          // rows per endpoint
    constexpr int R = 2 * H;               // rows per pair
    constexpr int F = R * R;               // outputs per pair
    __shared__ __align__(16) float tile[kPairThreads / 32][kPairsPerWarp * F];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int sub = lane / kGroup;         // pair slot inside the warp
    const int gl = lane % kGroup;          // lane inside the group
    const long long pair0 = ((long long)blockIdx.x * (kPairThreads / 32) + warp) * kPairsPerWarp;
    if (pair0 >= n) return;                // whole warp out of range
    const long long pair = pair0 + sub;
    const long long pc = pair < n ? pair : n - 1;          // clamp: inactive groups redo the last pair, never store

    long long ida = a_ids[pc], idb = b_ids[pc];
    // ids are validated on the host for numpy inputs; clamp so a bad device id can never fault
    ida = ida < 0 ? 0 : (ida >= st.num_nodes ? st.num_nodes - 1 : ida);
    idb = idb < 0 ? 0 : (idb >= st.num_nodes ? st.num_nodes - 1 : idb);
    const float* pa = st.data + ida * st.node_stride;
    const float* pb = st.data + idb * st.node_stride;

    long long stamp[R];
    if (LAZY) {
#pragma unroll
        for (int l = 1; l < H; ++l) {
            stamp[l] = st.stamps[ida * LAYERS + (l - 1)];
            stamp[H + l] = st.stamps[idb * LAYERS + (l - 1)];
        }
    }

    float acc[R * (R + 1) / 2];
#pragma unroll
    for (int i = 0; i < R * (R + 1) / 2; ++i) acc[i] = 0.f;

    for (int c = gl; c < ds4; c += kGroup) {
        float4 x[R];
#pragma unroll
        for (int l = 0; l < H; ++l) {
            x[l] = ld4(pa + (long long)l * st.row_stride + 4 * c);
            x[H + l] = ld4(pb + (long long)l * st.row_stride + 4 * c);
        }
        if (LAZY) {
#pragma unroll
            for (int l = 1; l < H; ++l) {
                float4 one[1];
                if (stamp[l] >= 0) {
                    one[0] = x[l];
                    replay<1>(one, st.decay_log, LAYERS, l - 1, stamp[l], st.epoch);
                    x[l] = one[0];
                }
                if (stamp[H + l] >= 0) {
                    one[0] = x[H + l];
                    replay<1>(one, st.decay_log, LAYERS, l - 1, stamp[H + l], st.epoch);
                    x[H + l] = one[0];
                }
            }
        }
        int e = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int q = r; q < R; ++q) {
                float s = acc[e];
                s = fmaf(x[r].x, x[q].x, s);
                s = fmaf(x[r].y, x[q].y, s);
                s = fmaf(x[r].z, x[q].z, s);
                s = fmaf(x[r].w, x[q].w, s);
                acc[e] = s;
                ++e;
            }
        }
    }

    // reduce over the 8 lanes of the group
#pragma unroll
    for (int i = 0; i < R * (R + 1) / 2; ++i) {
        float s = acc[i];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        acc[i] = s;
    }

    // epilogue: unique entry e is finished by lane (e mod 8) of the group and mirrored
    float* mine = &tile[warp][sub * F];
    {
        int e = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int q = r; q < R; ++q) {
                if ((e % kGroup) == gl) {
                    float g = acc[e];
                    if (apply_log_scale) {
                        g = g < 0.f ? 0.f : g;                 // random_feature[random_feature < 0] = 0
                        g = logf(__fadd_rn(g, 1.0f));          // torch.log(x + 1.0), not log1p
                    }
                    mine[r * R + q] = g;
                    mine[q * R + r] = g;
                }
                ++e;
            }
        }
    }
    __syncwarp();
    // coalesced write-out of the warp's (up to) 4 pairs: 4*F floats, contiguous in `out`
    const long long left = n - pair0;
    const int valid = (int)(left < kPairsPerWarp ? left : kPairsPerWarp) * F;     // multiple of 4 (F = 4 H^2)
    float* dst = out + pair0 * F;
    const float* srcs = &tile[warp][0];
    for (int i = lane * 4; i < valid; i += 32 * 4) {
        const float4 v = *reinterpret_cast<const float4*>(srcs + i);
        __stcs(reinterpret_cast<float4*>(dst + i), v);
    }
}

template <int LAYERS>
void launch(const StateView& v, const long long* a, const long long* b, long long n, int scale, float* out,
            cudaStream_t s) {
    const unsigned grid = (unsigned)((n + kPairsPerBlock - 1) / kPairsPerBlock);
    const int ds4 = (int)(v.row_stride / 4);
    if (v.stamps != nullptr)
        pairwise_kernel<LAYERS, true><<<grid, kPairThreads, 0, s>>>(v, a, b, n, scale, out, ds4);
    else
        pairwise_kernel<LAYERS, false><<<grid, kPairThreads, 0, s>>>(v, a, b, n, scale, out, ds4);
}

}  // namespace
}  // namespace tpn

extern "C" int tpn_pairwise(const tpn_state_t* st, const int64_t* a_ids_dev, const int64_t* b_ids_dev, int64_t n,
                            int apply_log_scale, float* out_dev, void* stream_v) {
    using namespace tpn;
    int rc = validate_state(st);
    if (rc != TPN_OK) return rc;
    if (n < 0 || n > (int64_t)1 << 34) return TPN_ERR_INVALID_ARGUMENT;
    if (n == 0) return TPN_OK;
    if (a_ids_dev == nullptr || b_ids_dev == nullptr || out_dev == nullptr ||
        (reinterpret_cast<uintptr_t>(out_dev) & 15) != 0)
        return TPN_ERR_INVALID_ARGUMENT;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const StateView v = make_view(st);
    const long long* a = reinterpret_cast<const long long*>(a_ids_dev);
    const long long* b = reinterpret_cast<const long long*>(b_ids_dev);
    switch (st->num_layer) {
        case 1: launch<1>(v, a, b, n, apply_log_scale, out_dev, stream); break;
        case 2: launch<2>(v, a, b, n, apply_log_scale, out_dev, stream); break;
        case 3: launch<3>(v, a, b, n, apply_log_scale, out_dev, stream); break;
        case 4: launch<4>(v, a, b, n, apply_log_scale, out_dev, stream); break;
        default: return TPN_ERR_UNSUPPORTED;
    }
    return check_launch();
}
