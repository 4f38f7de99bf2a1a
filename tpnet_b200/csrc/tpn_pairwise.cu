// tpn_pairwise — input of `self.mlp` in RandomProjectionModule.get_pair_wise_feature
// (reference: models/TPNet.py:112-129) for sm_100a.
//
// Per pair (a, b) the reference stacks the R = 2L+2 rows [a:P_0..P_L, b:P_0..P_L]
// (TPNet.py:119-121), forms the R x R Gram matrix with a batched fp32 GEMM (:122-123),
// clamps at 0 and takes log(x + 1.0) (:127-128).  The Gram blocks are 4x4 .. 10x10 over
// d ~ 100-200: ~3 flop/byte, far below the fp32 ridge, so this is a gather-bound kernel
// and tensor cores are deliberately not used.
//
// Mapping: G lanes per pair (G = 8: 4 pairs per warp, throughput shape; G = 32: one pair
// per warp, latency shape for the small decoder calls).  Lane j of a group reads float4
// columns j, j+G, ... of all R rows (the group reads G*16 contiguous bytes per row per
// step; a node's L+1 rows are one contiguous block in the node-major state) and keeps the
// R(R+1)/2 unique dot products in registers as packed (even, odd) float2 partial sums fed
// by the sm_100 packed FFMA2.  The sums are reduced over the group with a transposing
// butterfly (each shuffle step halves the number of live values), so every lane ends up
// owning NP/G finished entries, applies the epilogue to just those and mirrors them
// through shared memory; the warp then writes its pairs' R*R outputs as coalesced
// 128-bit streaming stores.
// Lazy-decay mode: rows of layers >= 1 are brought current in registers (one multiply by the
// row's pending decay factor) before they enter the products; nothing is written back.
#include "tpn_pairwise.cuh"

namespace tpn {
namespace {

thread_local int g_dev_slot = 0;      // device of the call in progress (set by the entry point)


template <int LAYERS, int G, bool LAZY>
__global__ void __launch_bounds__(kPairThreads)
pairwise_kernel(StateView st, const long long* __restrict__ a_ids, const long long* __restrict__ b_ids,
                long long n, int apply_log_scale, float* __restrict__ out, int ds4, const int* __restrict__ n_dev) {
    if (n_dev != nullptr) {                // routed calls: the count lives on the device
        if (*n_dev > n && st.err != nullptr) *st.err = 4;       // more pairs than the call was sized for
        n = min(n, (long long)max(*n_dev, 0));
    }
    constexpr int H = LAYERS + 1;          // rows per endpoint
    constexpr int R = 2 * H;               // rows per pair
    constexpr int F = R * R;               // outputs per pair
    constexpr int NU = R * (R + 1) / 2;    // unique Gram entries
    constexpr int PPW = 32 / G;            // pairs per warp
    __shared__ __align__(16) float tile[kPairThreads / 32][PPW * F];
    __shared__ unsigned short tab[kPairThreads / 32][NU];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;              // pair slot inside the warp
    const int gl = lane % G;               // lane inside the group
    const long long pair0 = ((long long)blockIdx.x * (kPairThreads / 32) + warp) * PPW;
    if (pair0 >= n) return;                // whole warp out of range
    const long long pair = pair0 + sub;
    const long long pc = pair < n ? pair : n - 1;          // inactive groups redo the last pair, never store

    long long ida = a_ids[pc], idb = b_ids[pc];
    fill_entry_table<R>(tab[warp], lane);
    // ids are validated on the host for numpy inputs; clamp so a bad device id can never fault
    ida = resolve_id(st, ida, true);
    idb = resolve_id(st, idb, true);
    const int rs4 = (int)(st.row_stride >> 2);
    const float4* pa = reinterpret_cast<const float4*>(st.data + ida * st.node_stride);
    const float4* pb = reinterpret_cast<const float4*>(st.data + idb * st.node_stride);

    // lazy decay: one pending factor per row (stamp < 0: the row is all zero in memory)
    float fac[R];
    if (LAZY) {
#pragma unroll
        for (int l = 1; l < H; ++l) {
            const long long sa = st.stamps[ida * LAYERS + (l - 1)];
            const long long sb = st.stamps[idb * LAYERS + (l - 1)];
            fac[l] = sa >= 0 ? decay_factor(st, l - 1, sa) : 1.0f;
            fac[H + l] = sb >= 0 ? decay_factor(st, l - 1, sb) : 1.0f;
        }
    }

    float2 acc[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) acc[i] = make_float2(0.f, 0.f);

    for (int c = gl; c < ds4; c += G) {
        float4 x[R];
#pragma unroll
        for (int l = 0; l < H; ++l) {
            x[l] = pa[l * rs4 + c];
            x[H + l] = pb[l * rs4 + c];
        }
        if (LAZY) {
#pragma unroll
            for (int l = 1; l < H; ++l) {
                scale4(x[l], fac[l]);
                scale4(x[H + l], fac[H + l]);
            }
        }
        gram_step<R>(acc, x);
    }

    __syncwarp();                          // entry table written by this warp is visible
    finish_pair<R, G>(acc, gl, &tile[warp][sub * F], tab[warp], apply_log_scale);
    __syncwarp();
    // coalesced write-out of the warp's pairs: PPW*F floats, contiguous in `out`
    const long long left = n - pair0;
    const int valid = (int)(left < PPW ? left : PPW) * F;     // multiple of 4 (F = 4 H^2)
    float* dst = out + pair0 * F;
    const float* srcs = &tile[warp][0];
    for (int i = lane * 4; i < valid; i += 32 * 4) {
        const float4 v = *reinterpret_cast<const float4*>(srcs + i);
        __stcs(reinterpret_cast<float4*>(dst + i), v);
    }
}

// ---------------------------------------------------------------- TMA (bulk-copy) variant
// Same math, different data movement: the L+1 rows of a node are ONE contiguous block of
// (L+1)*row_stride*4 bytes in the node-major state, so a warp fetches everything its 4 pairs
// need with at most 8 `cp.async.bulk` copies (global -> shared, completion counted on a
// per-warp mbarrier) issued by one lane.  The whole working set of the warp is in flight at
// once (no per-lane load instructions, no register staging, no dependent round trips per
// column step), and b-endpoints repeated by consecutive pairs — the encoder's pair lists
// repeat each b K times (TPNet.py:313-316) — are fetched once per warp.
template <int LAYERS, bool LAZY>
__global__ void __launch_bounds__(kPairThreads)
pairwise_tma_kernel(StateView st, const long long* __restrict__ a_ids, const long long* __restrict__ b_ids,
                    long long n, int apply_log_scale, float* __restrict__ out, int ds4, const int* __restrict__ n_dev) {
    if (n_dev != nullptr) {
        if (*n_dev > n && st.err != nullptr) *st.err = 4;
        n = min(n, (long long)max(*n_dev, 0));
    }
    constexpr int G = 8;
    constexpr int H = LAYERS + 1;
    constexpr int R = 2 * H;
    constexpr int F = R * R;
    constexpr int NU = R * (R + 1) / 2;
    constexpr int PPW = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar[kPairThreads / 32];
    __shared__ unsigned short tab[kPairThreads / 32][NU];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int gl = lane % G;
    const long long pair0 = ((long long)blockIdx.x * (kPairThreads / 32) + warp) * PPW;
    if (pair0 >= n) return;
    const long long pair = pair0 + sub;
    const long long pc = pair < n ? pair : n - 1;

    long long ida = a_ids[pc], idb = b_ids[pc];
    uint64_t* bar = &mbar[warp];
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fill_entry_table<R>(tab[warp], lane);
    ida = resolve_id(st, ida, true);
    idb = resolve_id(st, idb, true);

    const int rs4 = (int)(st.row_stride >> 2);
    const uint32_t block_bytes = (uint32_t)(H * st.row_stride * 4);       // rows P_0..P_L of one node
    unsigned char* wbase = smem_raw + (size_t)warp * (2 * PPW) * block_bytes;

    // b endpoints of the warp's 4 pairs; a run of equal ids shares one slot
    long long bid[PPW];
#pragma unroll
    for (int s = 0; s < PPW; ++s) bid[s] = __shfl_sync(0xffffffffu, idb, s * G);
    int sb = sub;
#pragma unroll
    for (int s = PPW - 1; s >= 1; --s)
        if (sb == s && bid[s] == bid[s - 1]) sb = s - 1;
    __syncwarp();
    if (lane == 0) {
        uint32_t copies = PPW;
#pragma unroll
        for (int s = 0; s < PPW; ++s) copies += (s == 0 || bid[s] != bid[s - 1]) ? 1u : 0u;
        mbar_expect_tx(bar, copies * block_bytes);
    }
    __syncwarp();                                   // expect_tx is registered before any copy can complete
    // lanes 0, 8, 16, 24 own one pair each: fetch its a block, and its b block unless shared
    if (gl == 0) {
        bulk_g2s(wbase + (size_t)sub * block_bytes, st.data + ida * st.node_stride, block_bytes, bar);
        if (sb == sub)
            bulk_g2s(wbase + (size_t)(PPW + sub) * block_bytes, st.data + idb * st.node_stride, block_bytes, bar);
    }

    // lazy decay: one pending factor per row (stamp < 0: the row is all zero in memory)
    float fac[R];
    if (LAZY) {
#pragma unroll
        for (int l = 1; l < H; ++l) {
            const long long sa = st.stamps[ida * LAYERS + (l - 1)];
            const long long sb = st.stamps[idb * LAYERS + (l - 1)];
            fac[l] = sa >= 0 ? decay_factor(st, l - 1, sa) : 1.0f;
            fac[H + l] = sb >= 0 ? decay_factor(st, l - 1, sb) : 1.0f;
        }
    }
    float2 acc[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) acc[i] = make_float2(0.f, 0.f);

    const float4* pa = reinterpret_cast<const float4*>(wbase + (size_t)sub * block_bytes);
    const float4* pb = reinterpret_cast<const float4*>(wbase + (size_t)(PPW + sb) * block_bytes);
    mbar_wait(bar, 0);

    for (int c = gl; c < ds4; c += G) {
        float4 x[R];
#pragma unroll
        for (int l = 0; l < H; ++l) {
            x[l] = pa[l * rs4 + c];
            x[H + l] = pb[l * rs4 + c];
        }
        if (LAZY) {
#pragma unroll
            for (int l = 1; l < H; ++l) {
                scale4(x[l], fac[l]);
                scale4(x[H + l], fac[H + l]);
            }
        }
        gram_step<R>(acc, x);
    }

    __syncwarp();                                   // every lane is done reading the row buffers
    float* tile = reinterpret_cast<float*>(wbase);  // reuse them as the output staging tile
    finish_pair<R, G>(acc, gl, tile + sub * F, tab[warp], apply_log_scale);
    __syncwarp();
    const long long left = n - pair0;
    const int valid = (int)(left < PPW ? left : PPW) * F;
    float* dst = out + pair0 * F;
    for (int i = lane * 4; i < valid; i += 32 * 4) {
        const float4 v = *reinterpret_cast<const float4*>(tile + i);
        __stcs(reinterpret_cast<float4*>(dst + i), v);
    }
}

template <int LAYERS>
bool launch_tma(const StateView& v, const long long* a, const long long* b, long long n, const int* n_dev, int scale, float* out,
                cudaStream_t s) {
    constexpr int R = 2 * (LAYERS + 1);
    const size_t block_bytes = (size_t)(LAYERS + 1) * v.row_stride * 4;
    const size_t smem = (size_t)(kPairThreads / 32) * 8 * block_bytes;
    // the staging tile (4 pairs * R*R floats) aliases the row buffers; keep >= 2 CTAs per SM
    if (smem > 110 * 1024 || 4 * R * R * sizeof(float) > 8 * block_bytes || (block_bytes & 15) != 0 ||
        (v.node_stride & 3) != 0)
        return false;
    static bool configured_tab[kMaxDevices][2];          // per device: the opt-in is a per-device attribute
    bool* configured = configured_tab[g_dev_slot];
    const bool lazy = v.stamps != nullptr;
    auto kernel = lazy ? pairwise_tma_kernel<LAYERS, true> : pairwise_tma_kernel<LAYERS, false>;
    if (!configured[lazy ? 1 : 0]) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
        configured[lazy ? 1 : 0] = true;
    }
    const unsigned grid = (unsigned)((n + 15) / 16);
    kernel<<<grid, kPairThreads, smem, s>>>(v, a, b, n, scale, out, (int)(v.row_stride / 4), n_dev);
    return true;
}

template <int LAYERS, int G>
void launch_g(const StateView& v, const long long* a, const long long* b, long long n, const int* n_dev, int scale, float* out,
              cudaStream_t s) {
    constexpr int pairs_per_block = kPairThreads / G;
    const unsigned grid = (unsigned)((n + pairs_per_block - 1) / pairs_per_block);
    const int ds4 = (int)(v.row_stride / 4);
    if (v.stamps != nullptr)
        pairwise_kernel<LAYERS, G, true><<<grid, kPairThreads, 0, s>>>(v, a, b, n, scale, out, ds4, n_dev);
    else
        pairwise_kernel<LAYERS, G, false><<<grid, kPairThreads, 0, s>>>(v, a, b, n, scale, out, ds4, n_dev);
}

template <int LAYERS>
void launch(const StateView& v, const long long* a, const long long* b, long long n, const int* n_dev, int scale, float* out,
            cudaStream_t s) {
    // few pairs (decoder calls): one warp per pair fills more SMs and needs fewer
    // dependent round trips per pair; many pairs: 8 lanes per pair wastes no lanes on d ~ 140
    if (n <= 4096) {
        launch_g<LAYERS, 32>(v, a, b, n, n_dev, scale, out, s);
    } else if (!launch_tma<LAYERS>(v, a, b, n, n_dev, scale, out, s)) {
        launch_g<LAYERS, 8>(v, a, b, n, n_dev, scale, out, s);          // rows too wide for the shared-memory staging
    }
}

}  // namespace
}  // namespace tpn

extern "C" int tpn_pairwise(const tpn_state_t* st, const int64_t* a_ids_dev, const int64_t* b_ids_dev, int64_t n,
                            const int32_t* n_dev, int apply_log_scale, float* out_dev, void* stream_v) {
    using namespace tpn;
    int rc = validate_state(st);
    if (rc != TPN_OK) return rc;
    if (n < 0 || n > (int64_t)1 << 34) return TPN_ERR_INVALID_ARGUMENT;
    if (n == 0) return TPN_OK;
    if (a_ids_dev == nullptr || b_ids_dev == nullptr || out_dev == nullptr ||
        (reinterpret_cast<uintptr_t>(out_dev) & 15) != 0)
        return TPN_ERR_INVALID_ARGUMENT;
    DeviceScope scope(st->data);
    g_dev_slot = scope.slot();
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const StateView v = make_view(st);
    const long long* a = reinterpret_cast<const long long*>(a_ids_dev);
    const long long* b = reinterpret_cast<const long long*>(b_ids_dev);
    const int* nd = reinterpret_cast<const int*>(n_dev);
    switch (st->num_layer) {
        case 1: launch<1>(v, a, b, n, nd, apply_log_scale, out_dev, stream); break;
        case 2: launch<2>(v, a, b, n, nd, apply_log_scale, out_dev, stream); break;
        case 3: launch<3>(v, a, b, n, nd, apply_log_scale, out_dev, stream); break;
        case 4: launch<4>(v, a, b, n, nd, apply_log_scale, out_dev, stream); break;
        default: return TPN_ERR_UNSUPPORTED;
    }
    return check_launch();
}
