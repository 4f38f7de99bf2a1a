// Host->device staging of the per-call id / timestamp arrays.
//
// The reference converts its numpy id arrays with `torch.from_numpy(...).to(device)` on every
// call (models/TPNet.py:74-77 and the implicit conversion at :109) — a pageable, synchronous
// copy.  Here one C call copies the arrays into a ring of pinned host slots (validating the
// ids on the way, which replaces the bounds check torch indexing performs), issues ONE
// cudaMemcpyAsync per call and records an event so that a slot is never overwritten before its
// copy has left.  The copy runs on the stager's OWN stream: the caller's stream only waits for the
// event, so the H2D copy of call k+1 overlaps the kernels of call k instead of queueing behind them
// (a copy issued on the compute stream is serialised with every kernel launched before it).  A device
// slot is rewritten `slots` calls later; the copy stream first waits for an event recorded on the
// caller's stream one call after the slot's consumers were enqueued.  No Python-level stream/event objects.
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#include "tpn_common.cuh"

struct tpn_stager {
    int slots;
    size_t slot_bytes;
    int cursor;
    char** host;          // pinned
    char** dev;
    cudaEvent_t* done;    // copy of slot i has left the pinned buffer (recorded on the copy stream)
    cudaEvent_t* seen;    // seen[j % slots]: recorded on the caller's stream at the START of call j
    bool* used;
    cudaStream_t copy_stream;
    long long calls;
};

namespace {

void release(tpn_stager* sg) {
    if (sg == nullptr) return;
    for (int i = 0; i < sg->slots; ++i) {
        if (sg->host && sg->host[i]) cudaFreeHost(sg->host[i]);
        if (sg->dev && sg->dev[i]) cudaFree(sg->dev[i]);
        if (sg->done && sg->done[i]) cudaEventDestroy(sg->done[i]);
        if (sg->seen && sg->seen[i]) cudaEventDestroy(sg->seen[i]);
    }
    if (sg->copy_stream != nullptr) cudaStreamDestroy(sg->copy_stream);
    free(sg->host);
    free(sg->dev);
    free(sg->done);
    free(sg->seen);
    free(sg->used);
    sg->host = nullptr;
    sg->dev = nullptr;
    sg->done = nullptr;
    sg->seen = nullptr;
    sg->used = nullptr;
    sg->copy_stream = nullptr;
}

// host threads of the staging pass: TPN_STAGE_THREADS, else min(8, cores / 2) — and never more than the
// process allows itself (OMP_NUM_THREADS / omp_set_num_threads: torchrun gives every rank 1 thread, the
// reference scripts ask for 3)
int stage_threads() {
    static int base = 0;
    if (base == 0) {
        const char* env = getenv("TPN_STAGE_THREADS");
        int v = env != nullptr ? atoi(env) : 0;
        if (v < 1) {
            v = omp_get_num_procs() / 2;
            v = v > 8 ? 8 : v;
        }
        base = v < 1 ? 1 : (v > 64 ? 64 : v);
    }
    if (getenv("TPN_STAGE_THREADS") != nullptr) return base;
    const int allowed = omp_get_max_threads();
    return allowed < base ? (allowed < 1 ? 1 : allowed) : base;
}

int allocate(tpn_stager* sg, size_t slot_bytes, int slots) {
    sg->slots = slots;
    sg->slot_bytes = slot_bytes;
    sg->cursor = 0;
    sg->host = (char**)calloc(slots, sizeof(char*));
    sg->dev = (char**)calloc(slots, sizeof(char*));
    sg->done = (cudaEvent_t*)calloc(slots, sizeof(cudaEvent_t));
    sg->seen = (cudaEvent_t*)calloc(slots, sizeof(cudaEvent_t));
    sg->used = (bool*)calloc(slots, sizeof(bool));
    sg->calls = 0;
    sg->copy_stream = nullptr;
    if (!sg->host || !sg->dev || !sg->done || !sg->seen || !sg->used) return TPN_ERR_INVALID_ARGUMENT;
    // TPN_STAGE_COPY_STREAM=0: copies on the caller's stream (A/B measurements)
    const char* env = getenv("TPN_STAGE_COPY_STREAM");
    if (slots >= 4 && !(env != nullptr && env[0] == '0')) {
        const cudaError_t e = cudaStreamCreateWithFlags(&sg->copy_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            sg->copy_stream = nullptr;               // fall back to the caller's stream
        }
    }
    for (int i = 0; i < slots; ++i) {
        cudaError_t e = cudaHostAlloc((void**)&sg->host[i], slot_bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc((void**)&sg->dev[i], slot_bytes);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sg->done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sg->seen[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            tpn::set_cuda_error(e);
            return TPN_ERR_CUDA;
        }
    }
    return TPN_OK;
}

}  // namespace

extern "C" int tpn_stager_create(tpn_stager_t** out, size_t slot_bytes, int slots) {
    if (out == nullptr || slots < 1 || slots > 64) return TPN_ERR_INVALID_ARGUMENT;
    if (slot_bytes < 4096) slot_bytes = 4096;
    tpn_stager* sg = (tpn_stager*)calloc(1, sizeof(tpn_stager));
    if (sg == nullptr) return TPN_ERR_INVALID_ARGUMENT;
    const int rc = allocate(sg, slot_bytes, slots);
    if (rc != TPN_OK) {
        release(sg);
        free(sg);
        return rc;
    }
    *out = sg;
    return TPN_OK;
}

extern "C" void tpn_stager_destroy(tpn_stager_t* sg) {
    if (sg == nullptr) return;
    release(sg);
    free(sg);
}

extern "C" int tpn_stage(tpn_stager_t* sg, const void* const* host, const int64_t* elems, const int* kinds,
                         int count, int64_t num_nodes, void** dev_out, void* stream_v) {
    if (sg == nullptr || host == nullptr || elems == nullptr || kinds == nullptr || dev_out == nullptr ||
        count < 1 || count > 8)
        return TPN_ERR_INVALID_ARGUMENT;
    tpn::DeviceScope scope(sg->dev[0]);               // the ring lives on the device it was created on
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    size_t total = 0;
    for (int i = 0; i < count; ++i) {
        if (elems[i] < 0 || host[i] == nullptr) return TPN_ERR_INVALID_ARGUMENT;
        total += (size_t)elems[i] * 8;
    }
    if (total > sg->slot_bytes) {                     // grow: rare (first call with a bigger batch)
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess && sg->copy_stream != nullptr) e = cudaStreamSynchronize(sg->copy_stream);
        if (e != cudaSuccess) {
            tpn::set_cuda_error(e);
            return TPN_ERR_CUDA;
        }
        const int slots = sg->slots;
        size_t want = sg->slot_bytes;
        while (want < total) want *= 2;
        release(sg);
        const int rc = allocate(sg, want, slots);
        if (rc != TPN_OK) return rc;
    }
    const int k = sg->cursor;
    sg->cursor = (k + 1) % sg->slots;
    if (sg->used[k]) {
        const cudaError_t e = cudaEventSynchronize(sg->done[k]);     // the copy out of this slot has left
        if (e != cudaSuccess) {
            tpn::set_cuda_error(e);
            return TPN_ERR_CUDA;
        }
    }
    char* dst = sg->host[k];
    size_t off = 0;
    // big batches (100k-edge calls move megabytes): the validate-and-copy pass is split over a few
    // host threads (OpenMP's persistent pool); small ones stay on the calling thread
    const int threads = total >= ((size_t)1 << 19) ? stage_threads() : 1;
    for (int i = 0; i < count; ++i) {
        const int64_t n = elems[i];
        if (kinds[i] == TPN_STAGE_RAW) {
            const char* src = reinterpret_cast<const char*>(host[i]);
            const int64_t bytes = n * 8, chunk = (bytes / threads + 63) & ~(int64_t)63;
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
            for (int t = 0; t < threads; ++t) {
                const int64_t b0 = (int64_t)t * chunk, b1 = b0 + chunk < bytes ? b0 + chunk : bytes;
                if (b0 < b1) memcpy(dst + off + b0, src + b0, (size_t)(b1 - b0));
            }
        } else {
            const int64_t* src = reinterpret_cast<const int64_t*>(host[i]);
            int64_t* out64 = reinterpret_cast<int64_t*>(dst + off);
            const int64_t lo = kinds[i] == TPN_STAGE_ID_WRAP ? -num_nodes : 0;   // indexing wraps, scatter does not
            int64_t mn = 0, mx = 0;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(min : mn) reduction(max : mx) if (threads > 1)
            for (int64_t j = 0; j < n; ++j) {
                const int64_t v = src[j];
                mn = v < mn ? v : mn;
                mx = v > mx ? v : mx;
                out64[j] = v < 0 ? v + num_nodes : v;
            }
            if (mn < lo || mx >= num_nodes) return TPN_ERR_INDEX;
        }
        dev_out[i] = sg->dev[k] + off;
        off += (size_t)n * 8;
    }
    cudaError_t e = cudaSuccess;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (sg->copy_stream != nullptr && cudaStreamIsCapturing(stream, &cap) != cudaSuccess) {
        (void)cudaGetLastError();
        cap = cudaStreamCaptureStatusNone;
    }
    if (sg->copy_stream != nullptr && cap == cudaStreamCaptureStatusNone) {
        // call j (slot k = j % slots).  seen[j % slots] <- "everything the caller enqueued before call j".  The device
        // slot was last used by call j - slots; its consumers were enqueued before call j - slots + 1 — or, when they
        // ran on a side stream of the caller (update_prepare), were joined into the caller's stream a call or two later.
        // The copy waits for the event of call j - 2: it still overlaps the kernels of the two calls before it, and
        // with slots >= 4 everything that ever read the slot is covered.
        const int S = sg->slots;
        const int js = (int)(sg->calls % S);
        e = cudaEventRecord(sg->seen[js], stream);
        if (e == cudaSuccess && sg->calls >= 2)
            e = cudaStreamWaitEvent(sg->copy_stream, sg->seen[(int)((sg->calls - 2) % S)], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(sg->dev[k], sg->host[k], total, cudaMemcpyHostToDevice, sg->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(sg->done[k], sg->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, sg->done[k], 0);
    } else {
        e = cudaMemcpyAsync(sg->dev[k], sg->host[k], total, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaEventRecord(sg->done[k], stream);
    }
    sg->calls += 1;
    if (e != cudaSuccess) {
        tpn::set_cuda_error(e);
        return TPN_ERR_CUDA;
    }
    sg->used[k] = true;
    return TPN_OK;
}
