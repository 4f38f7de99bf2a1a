// tpn_update — RandomProjectionModule.update (reference: models/TPNet.py:67-99) for sm_100a.
//
// Pipeline of one call (all on the caller's stream, no host sync):
//   1. prep      : w_j = exp(f32(-lambda) * (f32(t_last) - f32(t_j)))            (TPNet.py:77-78)
//                  2B messages  m <  B : target src[m], source dst[m]            (first scatter_add_, :93)
//                               m >= B : target dst[m-B], source src[m-B]        (second scatter_add_, :95)
//                  lazy mode: append c_1..c_L to the decay log as a new epoch.
//   2. sort      : stable sort of the message indices by target id, so that each
//                  target's messages are contiguous and in the reference's
//                  accumulation order (batch order within the src role, then the
//                  dst role).  B <= 2048: one CTA, bitonic network in shared memory
//                  on (target << 32 | m).  Larger: LSD radix sort, 8-bit digits,
//                  warp-match ranking (integer atomics only on histogram counts).
//   3. sweep     : eager mode only — P_l *= c_l over the whole state          (TPNet.py:83-85)
//   4. walk      : one launch per layer i = L..1 (top-down, TPNet.py:90, so layer i
//                  reads the pre-batch layer i-1).  A group of G lanes owns one target
//                  row: reads it once, replays pending decay (lazy), adds its messages
//                  sequentially as fadd_rn(acc, fmul_rn(P_{i-1}[v], w)), writes it once.
//                  128-bit loads/stores; consecutive lanes read consecutive float4.
//                  No float atomics anywhere: one owner per target row.
#include "tpn_common.cuh"

namespace tpn {

namespace {

constexpr int kSmallMaxMsgs = 4096;      // 2B <= 4096 -> single-CTA sort
constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 8;
constexpr int kRadixTile = kRadixThreads * kRadixItems;   // 2048 keys per block
constexpr int kRadixBins = 256;
constexpr int kWalkThreads = 256;

struct DecayArgs {
    float c[TPN_MAX_LAYERS];
    int has_decay;
};

struct Workspace {
    float* w;          // [B]
    uint32_t* key_a;   // [E]
    uint32_t* key_b;   // [E]
    uint32_t* val_a;   // [E]
    uint32_t* val_b;   // [E]
    uint32_t* ssrc;    // [E] source node of the p-th sorted message
    float* sw;         // [E] weight of the p-th sorted message
    uint32_t* hist;    // [256 * nblk]
    size_t bytes;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

Workspace carve(void* base, int64_t batch) {
    const size_t E = 2 * (size_t)batch;
    const size_t nblk = (E + kRadixTile - 1) / kRadixTile;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    Workspace ws;
    auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
    ws.w = reinterpret_cast<float*>(take(sizeof(float) * batch));
    ws.key_a = reinterpret_cast<uint32_t*>(take(4 * E));
    ws.key_b = reinterpret_cast<uint32_t*>(take(4 * E));
    ws.val_a = reinterpret_cast<uint32_t*>(take(4 * E));
    ws.val_b = reinterpret_cast<uint32_t*>(take(4 * E));
    ws.ssrc = reinterpret_cast<uint32_t*>(take(4 * E));
    ws.sw = reinterpret_cast<float*>(take(4 * E));
    ws.hist = reinterpret_cast<uint32_t*>(take(4 * kRadixBins * (nblk + 1)));
    ws.bytes = off;
    return ws;
}

__device__ __forceinline__ float edge_weight(double t, float t_last_f, float neg_lambda) {
    // fp32 subtract of the fp32-cast timestamps, fp32 multiply, then exp.  exp is
    // evaluated in f64 and rounded once (torch's CPU exp is a 1-ulp fp32 routine;
    // the correctly rounded value is the closest reproducible target).
    const float diff = __fsub_rn(t_last_f, (float)t);
    const float arg = __fmul_rn(neg_lambda, diff);
    return (float)exp((double)arg);
}

// ---------------------------------------------------------------- small path
__global__ void __launch_bounds__(1024)
prep_small_kernel(const long long* __restrict__ src, const long long* __restrict__ dst,
                  const double* __restrict__ t, int B, float t_last_f, float neg_lambda, long long num_nodes,
                  uint32_t* __restrict__ skey, uint32_t* __restrict__ ssrc, float* __restrict__ sw,
                  int* __restrict__ err_flag, float* decay_log, int L, long long new_epoch, DecayArgs decay) {
    __shared__ unsigned long long comp[kSmallMaxMsgs];
    __shared__ float wsm[kSmallMaxMsgs / 2];
    const int E = 2 * B;
    int P = 1;
    while (P < E) P <<= 1;
    if (threadIdx.x == 0 && decay_log != nullptr && decay.has_decay) {
        for (int l = 0; l < L; ++l) decay_log[new_epoch * L + l] = decay.c[l];
    }
    for (int m = threadIdx.x; m < P; m += blockDim.x) {
        unsigned long long c = ~0ull;                       // padding sorts last
        if (m < E) {
            const int j = m < B ? m : m - B;
            const long long tgt = m < B ? src[j] : dst[j];
            const long long oth = m < B ? dst[j] : src[j];
            const bool ok = tgt >= 0 && tgt < num_nodes && oth >= 0 && oth < num_nodes;
            if (!ok && err_flag != nullptr) *err_flag = 1;
            const uint32_t key = ok ? (uint32_t)tgt : (uint32_t)num_nodes;    // sentinel: dropped by walk
            c = ((unsigned long long)key << 32) | (uint32_t)m;
            if (m < B) wsm[m] = edge_weight(t[m], t_last_f, neg_lambda);
        }
        comp[m] = c;
    }
    __syncthreads();
    // bitonic network on the composite (target, message index): total order == stable order
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = comp[i], b = comp[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { comp[i] = b; comp[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int p = threadIdx.x; p < E; p += blockDim.x) {
        const unsigned long long c = comp[p];
        const uint32_t m = (uint32_t)c;
        const int j = m < (uint32_t)B ? m : m - B;
        skey[p] = (uint32_t)(c >> 32);
        ssrc[p] = (uint32_t)(m < (uint32_t)B ? dst[j] : src[j]);
        sw[p] = wsm[j];
    }
}

// ---------------------------------------------------------------- large path
__global__ void __launch_bounds__(256)
prep_large_kernel(const long long* __restrict__ src, const long long* __restrict__ dst,
                  const double* __restrict__ t, long long B, float t_last_f, float neg_lambda, long long num_nodes,
                  float* __restrict__ w, uint32_t* __restrict__ key, uint32_t* __restrict__ val,
                  int* __restrict__ err_flag, float* decay_log, int L, long long new_epoch, DecayArgs decay) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0 && decay_log != nullptr && decay.has_decay) {
        for (int l = 0; l < L; ++l) decay_log[new_epoch * L + l] = decay.c[l];
    }
    if (j >= B) return;
    const long long s = src[j], d = dst[j];
    const bool ok = s >= 0 && s < num_nodes && d >= 0 && d < num_nodes;
    if (!ok && err_flag != nullptr) *err_flag = 1;
    w[j] = edge_weight(t[j], t_last_f, neg_lambda);
    key[j] = ok ? (uint32_t)s : (uint32_t)num_nodes;
    key[B + j] = ok ? (uint32_t)d : (uint32_t)num_nodes;
    val[j] = (uint32_t)j;
    val[B + j] = (uint32_t)(B + j);
}

__global__ void __launch_bounds__(kRadixThreads)
radix_hist_kernel(const uint32_t* __restrict__ key, int E, int shift, uint32_t* __restrict__ hist, int nblk) {
    __shared__ uint32_t bins[kRadixBins];
    bins[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kRadixTile;
#pragma unroll
    for (int i = 0; i < kRadixItems; ++i) {
        const int idx = base + i * kRadixThreads + threadIdx.x;
        if (idx < E) atomicAdd(&bins[(key[idx] >> shift) & 0xff], 1u);     // integer count: order-independent
    }
    __syncthreads();
    hist[threadIdx.x * nblk + blockIdx.x] = bins[threadIdx.x];
}

// exclusive scan of `total` uint32 in place, one CTA
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ hist, int total) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    const int per = (total + 1023) / 1024;
    const int lo = tid * per, hi = min(lo + per, total);
    uint32_t local = 0;
    for (int i = lo; i < hi; ++i) local += hist[i];
    uint32_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_sum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t v = warp_sum[lane];
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += n;
        }
        warp_sum[lane] = s - v;     // exclusive
    }
    __syncthreads();
    uint32_t run = warp_sum[wid] + inc - local;
    for (int i = lo; i < hi; ++i) {
        const uint32_t v = hist[i];
        hist[i] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(kRadixThreads)
radix_scatter_kernel(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                     uint32_t* __restrict__ kout, uint32_t* __restrict__ vout, int E, int shift,
                     const uint32_t* __restrict__ offs, int nblk) {
    constexpr int kWarps = kRadixThreads / 32;
    __shared__ uint32_t wcount[kWarps][kRadixBins + 1];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < kWarps * (kRadixBins + 1); i += kRadixThreads) (&wcount[0][0])[i] = 0;
    __syncthreads();
    // warp `wid` owns the contiguous chunk [base, base + 32*items): order inside the
    // tile is (warp, item, lane), which is the input order — the pass is stable.
    const int base = blockIdx.x * kRadixTile + wid * (32 * kRadixItems);
    uint32_t k[kRadixItems], v[kRadixItems], rank[kRadixItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kRadixItems; ++i) {
        const int idx = base + i * 32 + lane;
        const bool valid = idx < E;
        k[i] = valid ? kin[idx] : 0u;
        v[i] = valid ? vin[idx] : 0u;
        const uint32_t dgt = valid ? ((k[i] >> shift) & 0xff) : (uint32_t)kRadixBins;
        const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
        const uint32_t before = wcount[wid][dgt];
        __syncwarp();
        if ((peers & lt_mask) == 0) wcount[wid][dgt] = before + __popc(peers);    // lowest peer lane updates
        __syncwarp();
        rank[i] = before + __popc(peers & lt_mask);
    }
    __syncthreads();
    {   // per digit: global offset of this block + exclusive prefix over warps
        const int dgt = tid;    // kRadixThreads == kRadixBins
        uint32_t run = offs[dgt * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = wcount[w][dgt];
            wcount[w][dgt] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRadixItems; ++i) {
        const int idx = base + i * 32 + lane;
        if (idx < E) {
            const uint32_t dgt = (k[i] >> shift) & 0xff;
            const uint32_t pos = wcount[wid][dgt] + rank[i];
            kout[pos] = k[i];
            vout[pos] = v[i];
        }
    }
}

__global__ void __launch_bounds__(256)
payload_kernel(const uint32_t* __restrict__ order, const long long* __restrict__ src,
               const long long* __restrict__ dst, const float* __restrict__ w, long long B, int E,
               uint32_t* __restrict__ ssrc, float* __restrict__ sw) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    const uint32_t m = order[p];
    const long long j = m < B ? m : m - B;
    ssrc[p] = (uint32_t)(m < B ? dst[j] : src[j]);
    sw[p] = w[j];
}

// ---------------------------------------------------------------- eager decay sweep
__global__ void __launch_bounds__(256)
sweep_decay_kernel(StateView st, DecayArgs decay, long long total4, int ds4) {
    // layers 1..L of a node are contiguous right after its layer-0 row
    const long long per_node4 = (long long)st.num_layer * ds4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const long long node = i / per_node4;
        const int r = (int)(i - node * per_node4);
        const int li = r / ds4;
        float* p = st.data + node * st.node_stride + st.row_stride + (long long)r * 4;
        float4 x = ld4(p);
        scale4(x, decay.c[li]);
        st4(p, x);
    }
}

// ---------------------------------------------------------------- the walk update, one layer
template <int G, int VPL, bool LAZY>
__global__ void __launch_bounds__(kWalkThreads)
walk_update_layer_kernel(StateView st, int layer, const uint32_t* __restrict__ skey,
                         const uint32_t* __restrict__ ssrc, const float* __restrict__ sw, int E, int ds4,
                         int write_stamp) {
    const int gid = (blockIdx.x * kWalkThreads + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    if (gid >= E) return;
    const uint32_t key = skey[gid];
    if ((long long)key >= st.num_nodes) return;          // dropped edge (id out of range)
    if (gid > 0 && skey[gid - 1] == key) return;         // not the head of its segment
    const int col0 = blockIdx.y * (G * VPL) + lane;      // float4 column of register 0
    const int L = st.num_layer;

    float* trow = st.data + (long long)key * st.node_stride + (long long)layer * st.row_stride;
    float4 acc[VPL];
    long long tstamp = 0;
    if (LAZY) tstamp = st.stamps[(long long)key * L + (layer - 1)];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int c = col0 + k * G;
        acc[k] = (c < ds4 && (!LAZY || tstamp >= 0)) ? ld4(trow + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (LAZY && tstamp >= 0) replay<VPL>(acc, st.decay_log, L, layer - 1, tstamp, st.epoch);

    const bool src_decays = LAZY && layer >= 2;          // P_0 never decays
    int p = gid;
    uint32_t v = ssrc[p];
    float w = sw[p];
    float4 x[VPL];
    long long vstamp = 0;
    {
        const float* srow = st.data + (long long)v * st.node_stride + (long long)(layer - 1) * st.row_stride;
        if (src_decays) vstamp = st.stamps[(long long)v * L + (layer - 2)];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int c = col0 + k * G;
            x[k] = (c < ds4 && vstamp >= 0) ? ld4(srow + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    while (true) {
        const int pn = p + 1;
        const bool more = pn < E && skey[pn] == key;
        float4 xn[VPL];
        uint32_t vn = 0;
        float wn = 0.f;
        long long vnstamp = 0;
        if (more) {                                       // issue the next row's loads before consuming this one
            vn = ssrc[pn];
            wn = sw[pn];
            const float* srow = st.data + (long long)vn * st.node_stride + (long long)(layer - 1) * st.row_stride;
            if (src_decays) vnstamp = st.stamps[(long long)vn * L + (layer - 2)];
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int c = col0 + k * G;
                xn[k] = (c < ds4 && vnstamp >= 0) ? ld4(srow + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (src_decays && vstamp >= 0) replay<VPL>(x, st.decay_log, L, layer - 2, vstamp, st.epoch);
#pragma unroll
        for (int k = 0; k < VPL; ++k) axpy4_rn(acc[k], x[k], w);
        if (!more) break;
#pragma unroll
        for (int k = 0; k < VPL; ++k) x[k] = xn[k];
        v = vn; w = wn; vstamp = vnstamp; p = pn;
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int c = col0 + k * G;
        if (c < ds4) st4(trow + 4 * c, acc[k]);
    }
    if (LAZY && write_stamp && lane == 0) st.stamps[(long long)key * L + (layer - 1)] = (int)st.epoch;
}

__global__ void __launch_bounds__(256)
stamp_targets_kernel(StateView st, int layer, const uint32_t* __restrict__ skey, int E) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    const uint32_t key = skey[p];
    if ((long long)key >= st.num_nodes) return;
    if (p > 0 && skey[p - 1] == key) return;
    st.stamps[(long long)key * st.num_layer + (layer - 1)] = (int)st.epoch;
}

template <int G, int VPL>
void launch_walk(const StateView& v, int layer, const Workspace& ws, int E, int ds4, int col_tiles, bool lazy,
                 cudaStream_t stream) {
    const long long groups = E;
    dim3 grid((unsigned)((groups * G + kWalkThreads - 1) / kWalkThreads), (unsigned)col_tiles);
    const int write_stamp = col_tiles == 1;
    if (lazy)
        walk_update_layer_kernel<G, VPL, true><<<grid, kWalkThreads, 0, stream>>>(v, layer, ws.key_a, ws.ssrc, ws.sw,
                                                                                  E, ds4, write_stamp);
    else
        walk_update_layer_kernel<G, VPL, false><<<grid, kWalkThreads, 0, stream>>>(v, layer, ws.key_a, ws.ssrc, ws.sw,
                                                                                   E, ds4, write_stamp);
}

template <int G>
void dispatch_walk(int vpl, const StateView& v, int layer, const Workspace& ws, int E, int ds4, int col_tiles,
                   bool lazy, cudaStream_t s) {
    switch (vpl) {
        case 1: launch_walk<G, 1>(v, layer, ws, E, ds4, col_tiles, lazy, s); break;
        case 2: launch_walk<G, 2>(v, layer, ws, E, ds4, col_tiles, lazy, s); break;
        case 3: launch_walk<G, 3>(v, layer, ws, E, ds4, col_tiles, lazy, s); break;
        case 4: launch_walk<G, 4>(v, layer, ws, E, ds4, col_tiles, lazy, s); break;
        case 5: launch_walk<G, 5>(v, layer, ws, E, ds4, col_tiles, lazy, s); break;
        case 6: launch_walk<G, 6>(v, layer, ws, E, ds4, col_tiles, lazy, s); break;
        case 7: launch_walk<G, 7>(v, layer, ws, E, ds4, col_tiles, lazy, s); break;
        default: launch_walk<G, 8>(v, layer, ws, E, ds4, col_tiles, lazy, s); break;
    }
}

}  // namespace

}  // namespace tpn

extern "C" size_t tpn_update_workspace_bytes(int64_t batch) {
    if (batch < 1) batch = 1;
    return tpn::carve(nullptr, batch).bytes;
}

extern "C" int tpn_update(tpn_state_t* st, const int64_t* src_dev, const int64_t* dst_dev, const double* t_dev,
                          int64_t batch, double t_last, float neg_lambda, const float* decay, void* ws_dev,
                          size_t ws_bytes, int32_t* err_flag_dev, void* stream_v) {
    using namespace tpn;
    int rc = validate_state(st);
    if (rc != TPN_OK) return rc;
    if (batch < 1 || batch > (int64_t)0x3fffffff || src_dev == nullptr || dst_dev == nullptr || t_dev == nullptr ||
        ws_dev == nullptr || st->num_nodes >= (int64_t)0xffffffffll)
        return TPN_ERR_INVALID_ARGUMENT;
    Workspace ws = carve(ws_dev, batch);
    if (ws.bytes > ws_bytes) return TPN_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const bool lazy = st->stamps != nullptr;
    const int L = st->num_layer;

    DecayArgs dargs;
    dargs.has_decay = 0;
    for (int l = 0; l < TPN_MAX_LAYERS; ++l) dargs.c[l] = 1.0f;
    if (decay != nullptr) {
        bool all_one = true;
        for (int l = 0; l < L; ++l) {
            dargs.c[l] = decay[l];
            all_one = all_one && decay[l] == 1.0f;
        }
        dargs.has_decay = all_one ? 0 : 1;      // x * 1.0f == x exactly: nothing to do
    }
    long long new_epoch = st->epoch;
    if (lazy && dargs.has_decay) {
        if (st->epoch + 1 >= st->log_capacity) return TPN_ERR_LOG_FULL;
        new_epoch = st->epoch + 1;
    }

    const int E = (int)(2 * batch);
    const float t_last_f = (float)t_last;
    const long long* src = reinterpret_cast<const long long*>(src_dev);
    const long long* dst = reinterpret_cast<const long long*>(dst_dev);
    float* log_w = lazy ? st->decay_log : nullptr;

    if (E <= kSmallMaxMsgs) {
        prep_small_kernel<<<1, 1024, 0, stream>>>(src, dst, t_dev, (int)batch, t_last_f, neg_lambda, st->num_nodes,
                                                  ws.key_a, ws.ssrc, ws.sw, err_flag_dev, log_w, L, new_epoch, dargs);
    } else {
        prep_large_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, stream>>>(
            src, dst, t_dev, batch, t_last_f, neg_lambda, st->num_nodes, ws.w, ws.key_a, ws.val_a, err_flag_dev, log_w,
            L, new_epoch, dargs);
        int bits = 0;
        while ((1ll << bits) <= st->num_nodes) ++bits;    // keys are in [0, num_nodes] (num_nodes = dropped)
        const int passes = (bits + 7) / 8;
        const int nblk = (E + kRadixTile - 1) / kRadixTile;
        uint32_t *kin = ws.key_a, *kout = ws.key_b, *vin = ws.val_a, *vout = ws.val_b;
        for (int p = 0; p < passes; ++p) {
            radix_hist_kernel<<<nblk, kRadixThreads, 0, stream>>>(kin, E, 8 * p, ws.hist, nblk);
            radix_scan_kernel<<<1, 1024, 0, stream>>>(ws.hist, kRadixBins * nblk);
            radix_scatter_kernel<<<nblk, kRadixThreads, 0, stream>>>(kin, vin, kout, vout, E, 8 * p, ws.hist, nblk);
            uint32_t* tk = kin; kin = kout; kout = tk;
            uint32_t* tv = vin; vin = vout; vout = tv;
        }
        if (kin != ws.key_a) {       // odd number of passes: sorted keys live in key_b
            cudaMemcpyAsync(ws.key_a, kin, sizeof(uint32_t) * E, cudaMemcpyDeviceToDevice, stream);
        }
        payload_kernel<<<(E + 255) / 256, 256, 0, stream>>>(vin, src, dst, ws.w, batch, E, ws.ssrc, ws.sw);
    }
    st->epoch = new_epoch;
    StateView view = make_view(st);

    const int ds4 = (int)(st->row_stride / 4);
    if (!lazy && dargs.has_decay) {
        const long long total4 = st->num_nodes * (long long)L * ds4;
        const long long want = (total4 + 255) / 256;
        const unsigned grid = (unsigned)(want < 148 * 16 ? (want < 1 ? 1 : want) : 148 * 16);
        sweep_decay_kernel<<<grid, 256, 0, stream>>>(view, dargs, total4, ds4);
    }

    // lanes per target row: 8 for rows up to 256 floats, else a full warp; <= 8 float4 per lane per column tile
    const int G = ds4 <= 64 ? 8 : 32;
    int vpl = (ds4 + G - 1) / G;
    int col_tiles = 1;
    if (vpl > 8) {
        col_tiles = (vpl + 7) / 8;
        vpl = 8;
    }
    for (int layer = L; layer >= 1; --layer) {
        if (G == 8) dispatch_walk<8>(vpl, view, layer, ws, E, ds4, col_tiles, lazy, stream);
        else dispatch_walk<32>(vpl, view, layer, ws, E, ds4, col_tiles, lazy, stream);
        if (lazy && col_tiles > 1)
            stamp_targets_kernel<<<(E + 255) / 256, 256, 0, stream>>>(view, layer, ws.key_a, E);
    }
    return check_launch();
}
