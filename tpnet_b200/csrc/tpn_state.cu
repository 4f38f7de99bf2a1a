// Row gather (RandomProjectionModule.get_random_projections, reference models/TPNet.py:101-110),
// state maintenance for the lazy-decay bookkeeping, and the small C-ABI utilities.
#include <string.h>

#include "tpn_common.cuh"

namespace tpn {

static thread_local char g_cuda_error[256] = "";

void set_cuda_error(cudaError_t e) {
    strncpy(g_cuda_error, cudaGetErrorString(e), sizeof(g_cuda_error) - 1);
    g_cuda_error[sizeof(g_cuda_error) - 1] = 0;
}

int check_launch() {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_cuda_error(e);
        return TPN_ERR_CUDA;
    }
    return TPN_OK;
}

DeviceScope::DeviceScope(const void* device_ptr) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
        (void)cudaGetLastError();
        prev = -1;
    }
    dev = prev;
    if (device_ptr != nullptr) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, device_ptr) == cudaSuccess &&
            (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged))
            dev = attr.device;
        else
            (void)cudaGetLastError();
    }
    if (dev != prev && dev >= 0) {
        if (cudaSetDevice(dev) == cudaSuccess) switched = true;
        else {
            (void)cudaGetLastError();
            dev = prev;
        }
    }
}

DeviceScope::~DeviceScope() {
    if (switched && prev >= 0) (void)cudaSetDevice(prev);
}

int device_sm_count() {
    static int table[kMaxDevices];          // 0 = not queried yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        (void)cudaGetLastError();
        return 148;
    }
    if (table[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
            (void)cudaGetLastError();
            n = 148;
        }
        table[dev] = n;
    }
    return table[dev];
}

namespace {

// One warp per requested id; lanes stride over the float4 columns of each layer row.
// out is [L+1][n][dim] (dim need not be a multiple of 4, so stores are scalar but
// consecutive lanes write consecutive 16-byte pieces).
template <bool LAZY>
__global__ void __launch_bounds__(256)
gather_kernel(StateView st, const long long* __restrict__ ids, long long n, float* __restrict__ out, int ds4) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    long long id = ids[i];
    id = resolve_id(st, id, true);
    const float* base = st.data + id * st.node_stride;
    const int d = st.dim;
    for (int l = 0; l <= st.num_layer; ++l) {
        long long stamp = -1;
        float fac = 1.0f;
        if (LAZY && l >= 1) {
            stamp = st.stamps[id * st.num_layer + (l - 1)];
            if (stamp >= 0) fac = decay_factor(st, l - 1, stamp);
        }
        float* o = out + ((long long)l * n + i) * d;
        for (int c = lane; c < ds4; c += 32) {
            float4 x[1];
            x[0] = ld4(base + (long long)l * st.row_stride + 4 * c);
            if (LAZY && stamp >= 0) scale4(x[0], fac);
            const int k = 4 * c;
            if (k + 3 < d) {
                o[k] = x[0].x; o[k + 1] = x[0].y; o[k + 2] = x[0].z; o[k + 3] = x[0].w;
            } else {
                if (k < d) o[k] = x[0].x;
                if (k + 1 < d) o[k + 1] = x[0].y;
                if (k + 2 < d) o[k + 2] = x[0].z;
            }
        }
    }
}

// Sharded exchange, sender side: out[i] = the whole node block (rows 0..L, node_stride floats)
// of ids[i], brought current in lazy mode.  One warp per id.
template <bool LAZY>
__global__ void __launch_bounds__(256)
gather_blocks_kernel(StateView st, const long long* __restrict__ ids, long long n, float* __restrict__ out, int ds4) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    long long id = ids[i];
    id = resolve_id(st, id, true);
    const float* base = st.data + id * st.node_stride;
    float* o = out + i * st.node_stride;
    const int block4 = (int)(st.node_stride >> 2);
    for (int c = lane; c < block4; c += 32) {
        const int l = c / ds4;                       // rows are row_stride apart; tail padding belongs to "row" > L
        float4 x[1];
        x[0] = ld4(base + 4 * (long long)c);
        if (LAZY && l >= 1 && l <= st.num_layer) {
            const long long stamp = st.stamps[id * st.num_layer + (l - 1)];
            if (stamp >= 0) scale4(x[0], decay_factor(st, l - 1, stamp));
        }
        st4(o + 4 * (long long)c, x[0]);
    }
}

// Lazy mode: bring every written row current and stamp it with the current epoch.
__global__ void __launch_bounds__(256) materialize_kernel(StateView st, long long rows, int ds4) {
    // one warp per (node, layer>=1) row
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const long long node = r / st.num_layer;
    const int li = (int)(r - node * st.num_layer);
    const long long stamp = st.stamps[r];
    if (stamp < 0 || stamp == st.epoch) return;       // all-zero row, or already current
    float* row = st.data + node * st.node_stride + (long long)(li + 1) * st.row_stride;
    const float fac = decay_factor(st, li, stamp);
    for (int c = lane; c < ds4; c += 32) {
        float4 x = ld4(row + 4 * c);
        scale4(x, fac);
        st4(row + 4 * c, x);
    }
    __syncwarp();
    if (lane == 0) st.stamps[r] = (int)st.epoch;
}

__global__ void __launch_bounds__(256) restart_stamps_kernel(int* __restrict__ stamps, long long rows) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows && stamps[r] >= 0) stamps[r] = 0;
}

__global__ void __launch_bounds__(256) clear_layers_kernel(StateView st, long long total4, int ds4) {
    const long long per_node4 = (long long)st.num_layer * ds4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const long long node = i / per_node4;
        const long long r = i - node * per_node4;
        st4(st.data + node * st.node_stride + st.row_stride + r * 4, z);
    }
}

__global__ void __launch_bounds__(256) fill_int_kernel(int* __restrict__ p, long long n, int v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

inline unsigned capped_grid(long long work_items, int per_block) {
    long long g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    const long long cap = (long long)device_sm_count() * 32;
    return (unsigned)(g > cap ? cap : g);
}

}  // namespace
}  // namespace tpn

extern "C" int tpn_version(void) { return TPN_ABI_VERSION; }

extern "C" const char* tpn_error_string(int code) {
    switch (code) {
        case TPN_OK: return "ok";
        case TPN_ERR_INVALID_ARGUMENT: return "invalid argument";
        case TPN_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case TPN_ERR_LOG_FULL: return "decay log full: call tpn_materialize + tpn_reset_epoch";
        case TPN_ERR_CUDA: return "CUDA runtime error";
        case TPN_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}

extern "C" const char* tpn_last_cuda_error(void) { return tpn::g_cuda_error; }

extern "C" int tpn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
        tpn::set_cuda_error(e);
        return TPN_ERR_CUDA;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return TPN_OK;
}

extern "C" int tpn_gather(const tpn_state_t* st, const int64_t* ids_dev, int64_t n, float* out_dev, void* stream_v) {
    using namespace tpn;
    int rc = validate_state(st);
    if (rc != TPN_OK) return rc;
    if (n < 0) return TPN_ERR_INVALID_ARGUMENT;
    if (n == 0) return TPN_OK;
    if (ids_dev == nullptr || out_dev == nullptr) return TPN_ERR_INVALID_ARGUMENT;
    DeviceScope scope(st->data);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const StateView v = make_view(st);
    const int ds4 = (int)(st->row_stride / 4);
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    const long long* ids = reinterpret_cast<const long long*>(ids_dev);
    if (v.stamps != nullptr) gather_kernel<true><<<grid, 256, 0, stream>>>(v, ids, n, out_dev, ds4);
    else gather_kernel<false><<<grid, 256, 0, stream>>>(v, ids, n, out_dev, ds4);
    return check_launch();
}

extern "C" int tpn_gather_blocks(const tpn_state_t* st, const int64_t* ids_dev, int64_t n, float* out_dev,
                                 void* stream_v) {
    using namespace tpn;
    int rc = validate_state(st);
    if (rc != TPN_OK) return rc;
    if (n < 0) return TPN_ERR_INVALID_ARGUMENT;
    if (n == 0) return TPN_OK;
    if (ids_dev == nullptr || out_dev == nullptr || (reinterpret_cast<uintptr_t>(out_dev) & 15) != 0)
        return TPN_ERR_INVALID_ARGUMENT;
    DeviceScope scope(st->data);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const StateView v = make_view(st);
    const int ds4 = (int)(st->row_stride / 4);
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    const long long* ids = reinterpret_cast<const long long*>(ids_dev);
    if (v.stamps != nullptr) gather_blocks_kernel<true><<<grid, 256, 0, stream>>>(v, ids, n, out_dev, ds4);
    else gather_blocks_kernel<false><<<grid, 256, 0, stream>>>(v, ids, n, out_dev, ds4);
    return check_launch();
}

extern "C" int tpn_materialize(tpn_state_t* st, void* stream_v) {
    using namespace tpn;
    int rc = validate_state(st);
    if (rc != TPN_OK) return rc;
    if (st->stamps == nullptr || st->epoch == 0) return TPN_OK;
    DeviceScope scope(st->data);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const StateView v = make_view(st);
    const long long rows = st->num_nodes * st->num_layer;
    const unsigned grid = (unsigned)((rows * 32 + 255) / 256);
    materialize_kernel<<<grid, 256, 0, stream>>>(v, rows, (int)(st->row_stride / 4));
    return check_launch();
}

extern "C" int tpn_reset_epoch(tpn_state_t* st, void* stream_v) {
    using namespace tpn;
    int rc = validate_state(st);
    if (rc != TPN_OK) return rc;
    if (st->stamps == nullptr) return TPN_OK;
    DeviceScope scope(st->data);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const long long rows = st->num_nodes * st->num_layer;
    restart_stamps_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, stream>>>(st->stamps, rows);
    st->epoch = 0;
    st->cum_floor = 1.0;
    return check_launch();
}

extern "C" int tpn_clear_walk_layers(tpn_state_t* st, void* stream_v) {
    using namespace tpn;
    int rc = validate_state(st);
    if (rc != TPN_OK) return rc;
    DeviceScope scope(st->data);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    const StateView v = make_view(st);
    const int ds4 = (int)(st->row_stride / 4);
    const long long total4 = st->num_nodes * (long long)st->num_layer * ds4;
    clear_layers_kernel<<<capped_grid(total4, 256), 256, 0, stream>>>(v, total4, ds4);
    if (st->stamps != nullptr) {
        const long long rows = st->num_nodes * st->num_layer;
        fill_int_kernel<<<capped_grid(rows, 256), 256, 0, stream>>>(st->stamps, rows, -1);
        st->epoch = 0;
        st->cum_floor = 1.0;
    }
    return check_launch();
}
