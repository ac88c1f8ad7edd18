// Workspace, error reporting, memory helpers and profiling counters of the C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "peeb_common.cuh"

namespace peeb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return PEEB_E_CUDA;
}

int scratch_reserve(Scratch& s, size_t bytes, bool pinned_host) {
    if (bytes <= s.cap) return PEEB_OK;
    scratch_free(s, pinned_host);
    size_t want = align_up(bytes + bytes / 4, 1 << 16);
    if (pinned_host) PEEB_CUDA(cudaHostAlloc(&s.ptr, want, cudaHostAllocDefault));
    else PEEB_CUDA(cudaMalloc(&s.ptr, want));
    s.cap = want;
    return PEEB_OK;
}

void scratch_free(Scratch& s, bool pinned_host) {
    if (s.ptr) {
        if (pinned_host) cudaFreeHost(s.ptr); else cudaFree(s.ptr);
    }
    s.ptr = nullptr;
    s.cap = 0;
}

}  // namespace peeb

using namespace peeb;

extern "C" {

int peeb_abi_version(void) { return PEEB_ABI_VERSION; }

const char* peeb_last_error(void) { return g_err; }

int peeb_device_count(int* n) {
    PEEB_REQUIRE(n != nullptr, "peeb_device_count: null pointer");
    PEEB_CUDA(cudaGetDeviceCount(n));
    return PEEB_OK;
}

int peeb_ws_create(int device, peeb_ws** out) {
    PEEB_REQUIRE(out != nullptr, "peeb_ws_create: null pointer");
    *out = nullptr;
    int ndev = 0;
    PEEB_CUDA(cudaGetDeviceCount(&ndev));
    PEEB_REQUIRE(device >= 0 && device < ndev, "peeb_ws_create: device %d out of range (%d visible)", device, ndev);
    PEEB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PEEB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("libpeeb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return PEEB_E_UNSUPPORTED;
    }
    peeb_ws* ws = new peeb_ws();
    ws->device = device;
    ws->sm_count = prop.multiProcessorCount;
    ws->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    PEEB_CUDA(cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking));
    {   // peeb_dev_alloc / peeb_dev_free: keep freed blocks in the pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    PEEB_CUDA(cudaStreamCreateWithFlags(&ws->stream2, cudaStreamNonBlocking));
    PEEB_CUDA(cudaStreamCreateWithFlags(&ws->stream3, cudaStreamNonBlocking));
    for (int i = 0; i < peeb_ws::kPipeEvents; ++i) PEEB_CUDA(cudaEventCreateWithFlags(&ws->pipe_ev[i], cudaEventDisableTiming));
    if (const char* pr = getenv("PEEB_PIPE_ROLES")) ws->pipe_roles = pr[0] != '0';
    for (int i = 0; i < 4; ++i) PEEB_CUDA(cudaEventCreateWithFlags(&ws->ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) PEEB_CUDA(cudaEventCreate(&ws->prof_ev[i]));
    for (int i = 0; i < 2; ++i)
        for (int k = 0; k < peeb_ws::kTableRing; ++k) PEEB_CUDA(cudaEventCreateWithFlags(&ws->pev[i][k], cudaEventDisableTiming));
    const char* nb = getenv("PEEB_NO_BULK");
    ws->use_bulk = !(nb && nb[0] == '1');
    *out = ws;
    return PEEB_OK;
}

int peeb_ws_destroy(peeb_ws* ws) {
    if (!ws) return PEEB_OK;
    cudaSetDevice(ws->device);
    if (ws->stream) cudaStreamSynchronize(ws->stream);
    if (ws->stream2) cudaStreamSynchronize(ws->stream2);
    if (ws->stream3) cudaStreamSynchronize(ws->stream3);
    scratch_free(ws->tables);
    scratch_free(ws->tables_h, true);
    scratch_free(ws->stage);
    scratch_free(ws->stage2);
    scratch_free(ws->bits);
    scratch_free(ws->info_h, true);
    for (int i = 0; i < 2; ++i) {
        scratch_free(ws->ptables[i]);
        for (int k = 0; k < peeb_ws::kTableRing; ++k) scratch_free(ws->ptables_h[i][k], true);
        scratch_free(ws->pbits[i]);
        if (i == 0) { scratch_free(ws->step_counters); scratch_free(ws->hist); }
        for (int k = 0; k < peeb_ws::kTableRing; ++k) if (ws->pev[i][k]) cudaEventDestroy(ws->pev[i][k]);
    }
    for (int i = 0; i < 4; ++i) if (ws->ev[i]) cudaEventDestroy(ws->ev[i]);
    for (int i = 0; i < 2; ++i) if (ws->prof_ev[i]) cudaEventDestroy(ws->prof_ev[i]);
    if (ws->stream) cudaStreamDestroy(ws->stream);
    if (ws->stream2) cudaStreamDestroy(ws->stream2);
    if (ws->stream3) cudaStreamDestroy(ws->stream3);
    for (int i = 0; i < peeb_ws::kPipeEvents; ++i) if (ws->pipe_ev[i]) cudaEventDestroy(ws->pipe_ev[i]);
    delete ws;
    return PEEB_OK;
}

int peeb_ws_sync(peeb_ws* ws) {
    PEEB_REQUIRE(ws != nullptr, "peeb_ws_sync: null workspace");
    PEEB_CUDA(cudaSetDevice(ws->device));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream2));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream3));
    return PEEB_OK;
}

void* peeb_ws_stream(peeb_ws* ws) { return ws ? (void*)ws->stream : nullptr; }

int peeb_ws_set_option(peeb_ws* ws, int option, int value) {
    PEEB_REQUIRE(ws != nullptr, "peeb_ws_set_option: null workspace");
    if (option == PEEB_OPT_BULK) ws->use_bulk = value ? 1 : 0;
    else { set_error("peeb_ws_set_option: unknown option %d", option); return PEEB_E_INVALID; }
    return PEEB_OK;
}

int peeb_host_alloc(size_t bytes, void** ptr) {
    PEEB_REQUIRE(ptr != nullptr, "peeb_host_alloc: null pointer");
    PEEB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return PEEB_OK;
}

int peeb_host_free(void* ptr) {
    if (ptr) PEEB_CUDA(cudaFreeHost(ptr));
    return PEEB_OK;
}

int peeb_dev_alloc(peeb_ws* ws, size_t bytes, void** ptr) {
    PEEB_REQUIRE(ws && ptr, "peeb_dev_alloc: null pointer");
    PEEB_CUDA(cudaSetDevice(ws->device));
    // stream-ordered allocation from the device's pool (kept resident, see peeb_ws_create): a chain of calls
    // that allocates its intermediates every time does not pay cudaMalloc / cudaFree (and their device-wide
    // synchronisation) per buffer.  Every use of these buffers goes through the workspace stream.
    PEEB_CUDA(cudaMallocAsync(ptr, bytes ? bytes : 1, ws->stream));
    return PEEB_OK;
}

int peeb_dev_free(peeb_ws* ws, void* ptr) {
    PEEB_REQUIRE(ws != nullptr, "peeb_dev_free: null workspace");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (ptr) PEEB_CUDA(cudaFreeAsync(ptr, ws->stream));
    return PEEB_OK;
}

int peeb_memcpy_h2d(peeb_ws* ws, void* dst, const void* src, size_t bytes, void* stream) {
    PEEB_REQUIRE(ws != nullptr, "peeb_memcpy_h2d: null workspace");
    PEEB_CUDA(cudaSetDevice(ws->device));
    PEEB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return PEEB_OK;
}

int peeb_memcpy_d2h(peeb_ws* ws, void* dst, const void* src, size_t bytes, void* stream) {
    PEEB_REQUIRE(ws != nullptr, "peeb_memcpy_d2h: null workspace");
    PEEB_CUDA(cudaSetDevice(ws->device));
    PEEB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PEEB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PEEB_OK;
}

int peeb_prof_enable(peeb_ws* ws, int on) {
    PEEB_REQUIRE(ws != nullptr, "peeb_prof_enable: null workspace");
    ws->prof_on = on ? 1 : 0;
    for (int i = 0; i < PEEB_PROF_SLOTS; ++i) { ws->prof_ms[i] = 0.f; ws->prof_calls[i] = 0; }
    return PEEB_OK;
}

int peeb_prof_get(peeb_ws* ws, int slot, double* total_ms, long long* launches) {
    PEEB_REQUIRE(ws && slot >= 0 && slot < PEEB_PROF_SLOTS, "peeb_prof_get: bad arguments");
    if (total_ms) *total_ms = ws->prof_ms[slot];
    if (launches) *launches = ws->prof_calls[slot];
    return PEEB_OK;
}

// Counters of the PEE embed kernel's warp-steps: {all, generic code at a border column, redone by the generic code
// after the fast code saw a value leave the range}.  on = 1 starts counting from zero, on = 0 stops; out3 may be null.
int peeb_pee_step_counters(peeb_ws* ws, int on, uint64_t* out3) {
    PEEB_REQUIRE(ws != nullptr, "peeb_pee_step_counters: null workspace");
    PEEB_CUDA(cudaSetDevice(ws->device));
    int rc = scratch_reserve(ws->step_counters, 64);
    if (rc) return rc;
    PEEB_CUDA(cudaDeviceSynchronize());
    if (out3) PEEB_CUDA(cudaMemcpy(out3, ws->step_counters.ptr, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (on) PEEB_CUDA(cudaMemset(ws->step_counters.ptr, 0, 64));
    ws->step_counters_on = on ? 1 : 0;
    return PEEB_OK;
}

const char* peeb_prof_name(int slot) {
    static const char* names[PEEB_PROF_SLOTS] = {
        "moments", "hist_planes", "tile_moments", "lsb_embed", "planes_pack", "planes_unpack", "compact_bits",
        "pee_count", "pee_embed", "pee_extract", "pee_gather", "pee_hist", "pee_finalize", "lsb_recover", "lsb_extract",
        "bitmap_encode", "bitmap_decode", "", "", ""};
    return (slot >= 0 && slot < PEEB_PROF_SLOTS) ? names[slot] : "";
}

}  // extern "C"
