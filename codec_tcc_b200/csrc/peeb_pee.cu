// Row a10: reversible Prediction-Error Expansion on sm_100a -- entry points, the
// host-buffer (two-stream, chunked) pipeline, threshold-selection histogram.  The embed / extract kernels
// are in peeb_pee2.cu (rhombus predictor, two-pass checkerboard) and peeb_pee_med.cu (causal MED predictor,
// wavefront extract).
#include <algorithm>
#include <cstdlib>

#include <vector>

#include "peeb_common.cuh"
#include "peeb_pee.cuh"

namespace peeb {


__global__ void pee_finalize_kernel(PeeBatch bt) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= bt.n_units) return;
    long long* info = bt.info + (long long)u * PEEB_INFO;
    info[0] = bt.T[u];
    info[1] = bt.n_bits[u];
    info[2] = info[3] + info[4];
    info[7] = ((long long)bt.n_bits[u] > info[2]) ? PEEB_E_CAPACITY : 0;
}

// ------------------------------------------------------------------ host side
// upload T / n_bits (host arrays) into table set `slot` of the workspace; returns device pointers
int upload_unit_tables(peeb_ws* ws, int slot, int n_units, const int32_t* T, const int64_t* n_bits,
                              int bit_depth, size_t extra_bytes, cudaStream_t st, int** dT, unsigned** dN,
                              char** extra) {
    const size_t head = align_up((size_t)n_units * 8, 256);
    int rc = scratch_reserve(ws->ptables[slot], head + extra_bytes + 256);
    if (rc) return rc;
    // the pinned mirrors are a ring: wait until the upload that last used this entry (four calls ago) has been consumed
    const int ring = ws->ptable_seq[slot]++ % peeb_ws::kTableRing;
    PEEB_CUDA(cudaEventSynchronize(ws->pev[slot][ring]));
    rc = scratch_reserve(ws->ptables_h[slot][ring], head + 256, true);  // (+ a pinned word behind the tables for callers)
    if (rc) return rc;
    ws->ptable_h_cur[slot] = ws->ptables_h[slot][ring].ptr;
    int* hT = (int*)ws->ptables_h[slot][ring].ptr;
    unsigned* hN = (unsigned*)(hT + n_units);
    const int tmax = 1 << (bit_depth - 1);
    for (int u = 0; u < n_units; ++u) {
        PEEB_REQUIRE(T[u] >= 1 && T[u] <= tmax, "pee: T[%d]=%d outside 1..%d", u, T[u], tmax);
        PEEB_REQUIRE(n_bits[u] >= 0 && n_bits[u] < (1ll << 31), "pee: n_bits[%d] out of range", u);
        hT[u] = T[u];
        hN[u] = (unsigned)n_bits[u];
    }
    PEEB_CUDA(cudaMemcpyAsync(ws->ptables[slot].ptr, hT, (size_t)n_units * 8, cudaMemcpyHostToDevice, st));
    PEEB_CUDA(cudaEventRecord(ws->pev[slot][ring], st));
    *dT = (int*)ws->ptables[slot].ptr;
    *dN = (unsigned*)((int*)ws->ptables[slot].ptr + n_units);
    *extra = (char*)ws->ptables[slot].ptr + head;
    return PEEB_OK;
}

// The kernels live in peeb_pee2.cu (row-pair layout); images without an interior (h < 3 or w < 3) carry
// nothing: marked == source, empty location map, capacity 0.
static int embed_batch_impl(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                            int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload,
                            int64_t payload_stride, void* marked, int64_t marked_stride, uint8_t* lm,
                            int64_t lm_stride, int64_t* info, cudaStream_t st, int slot = 0) {
    PEEB_REQUIRE(ws && src && n_bits && payload && info, "peeb_pee_embed_batch: null pointer");  // T may be null: chosen on the device
    PEEB_REQUIRE(n_units >= 1, "peeb_pee_embed_batch: n_units must be >= 1");
    PEEB_REQUIRE(((uintptr_t)payload & 3) == 0 && (payload_stride & 3) == 0, "peeb_pee_embed_batch: payload must be 4-byte aligned");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "pee: itemsize must be 1 or 2");
    PEEB_REQUIRE(bit_depth >= 1 && bit_depth <= 8 * itemsize, "pee: bit_depth %d out of range for itemsize %d", bit_depth, itemsize);
    PEEB_REQUIRE(h >= 1 && w >= 1 && (long long)h * w < (1ll << 31), "pee: image size %dx%d unsupported", h, w);
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (h >= 3 && w >= 3)
        return embed_batch_impl2(ws, src, src_stride, n_units, h, w, itemsize, bit_depth, T, n_bits, payload, payload_stride,
                                 marked, marked_stride, lm, lm_stride, info, st, slot);
    int* dT; unsigned* dN; char* extra;
    std::vector<int32_t> ones;
    if (!T) { ones.assign((size_t)n_units, 1); T = ones.data(); }  // nothing can be embedded: T = 1 is as good as any
    int rc = upload_unit_tables(ws, slot, n_units, T, n_bits, bit_depth, 256, st, &dT, &dN, &extra);
    if (rc) return rc;
    PEEB_CUDA(cudaMemsetAsync(info, 0, sizeof(int64_t) * PEEB_INFO * n_units, st));
    for (int u = 0; u < n_units; ++u) {
        if (marked) PEEB_CUDA(cudaMemcpyAsync((char*)marked + u * marked_stride, (const char*)src + u * src_stride,
                                              (size_t)h * w * itemsize, cudaMemcpyDeviceToDevice, st));
        if (lm) PEEB_CUDA(cudaMemsetAsync(lm + u * lm_stride, 0, (size_t)h * ((w + 7) / 8), st));
    }
    PeeBatch bt{};
    bt.T = dT; bt.n_bits = dN; bt.info = (long long*)info; bt.n_units = n_units;
    pee_finalize_kernel<<<(n_units + 127) / 128, 128, 0, st>>>(bt);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

static int extract_batch_impl(peeb_ws* ws, const void* marked, int64_t marked_stride, int n_units, int h, int w,
                              int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits,
                              const uint8_t* lm, int64_t lm_stride, uint8_t* payload_out, int64_t payload_stride,
                              void* recovered, int64_t recovered_stride, int64_t* info, cudaStream_t st,
                              int slot = 0) {
    PEEB_REQUIRE(ws && marked && T && n_bits && lm && payload_out && info, "peeb_pee_extract_batch: null pointer");
    PEEB_REQUIRE(n_units >= 1 && n_units <= 65535, "peeb_pee_extract_batch: n_units must be 1..65535");
    PEEB_REQUIRE(((uintptr_t)payload_out & 3) == 0 && (payload_stride & 3) == 0, "peeb_pee_extract_batch: payload_out must be 4-byte aligned");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "pee: itemsize must be 1 or 2");
    PEEB_REQUIRE(bit_depth >= 1 && bit_depth <= 8 * itemsize, "pee: bit_depth %d out of range for itemsize %d", bit_depth, itemsize);
    PEEB_REQUIRE(h >= 1 && w >= 1 && (long long)h * w < (1ll << 31), "pee: image size %dx%d unsupported", h, w);
    PEEB_CUDA(cudaSetDevice(ws->device));
    for (int u = 0; u < n_units; ++u) {
        const size_t pb = peeb_payload_bytes(n_bits[u]);
        PEEB_REQUIRE(n_units == 1 || (int64_t)pb <= payload_stride, "peeb_pee_extract_batch: payload_stride too small for unit %d", u);
    }
    // every unit's output words start from zero: the extract kernel clears them itself (the gather kernel
    // ORs the boundary words in); the path for images without an interior uses a memset
    if (h >= 3 && w >= 3)
        return extract_batch_impl2(ws, marked, marked_stride, n_units, h, w, itemsize, bit_depth, T, n_bits, lm, lm_stride,
                                   payload_out, payload_stride, recovered, recovered_stride, info, st, slot);
    if (n_units == 1) PEEB_CUDA(cudaMemsetAsync(payload_out, 0, peeb_payload_bytes(n_bits[0]), st));
    else PEEB_CUDA(cudaMemsetAsync(payload_out, 0, (size_t)payload_stride * n_units, st));
    int* dT; unsigned* dN; char* extra;
    int rc = upload_unit_tables(ws, slot, n_units, T, n_bits, bit_depth, 256, st, &dT, &dN, &extra);
    if (rc) return rc;
    PEEB_CUDA(cudaMemsetAsync(info, 0, sizeof(int64_t) * PEEB_INFO * n_units, st));
    for (int u = 0; u < n_units; ++u)
        if (recovered) PEEB_CUDA(cudaMemcpyAsync((char*)recovered + u * recovered_stride, (const char*)marked + u * marked_stride,
                                                 (size_t)h * w * itemsize, cudaMemcpyDeviceToDevice, st));
    PeeBatch bt{};
    bt.T = dT; bt.n_bits = dN; bt.info = (long long*)info; bt.n_units = n_units;
    pee_finalize_kernel<<<(n_units + 127) / 128, 128, 0, st>>>(bt);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

static int hist_batch_impl(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                           int bit_depth, uint32_t* hist, cudaStream_t st) {
    PEEB_REQUIRE(ws && src && hist, "peeb_pee_hist_batch: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_pee_hist_batch: itemsize must be 1 or 2");
    PEEB_REQUIRE(bit_depth >= 1 && bit_depth <= 8 * itemsize, "peeb_pee_hist_batch: bad bit_depth");
    PEEB_REQUIRE(n_units >= 1 && n_units <= 65535 && h >= 1 && w >= 1, "peeb_pee_hist_batch: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const int tmax = 1 << (bit_depth - 1);
    PEEB_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 4 * (size_t)tmax * n_units, st));
    if (h < 3 || w < 3) return PEEB_OK;
    return hist_batch_impl2(ws, src, src_stride, n_units, h, w, itemsize, bit_depth, hist, st);
}

}  // namespace peeb

using namespace peeb;

extern "C" {

size_t peeb_payload_bytes(int64_t n_bits) {
    if (n_bits < 0) n_bits = 0;
    return align_up((size_t)((n_bits + 7) / 8), 4) + 8;
}

int peeb_pee_embed_batch(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                         int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload,
                         int64_t payload_stride, void* marked, int64_t marked_stride, uint8_t* lm, int64_t lm_stride,
                         int64_t* info, void* stream) {
    return embed_batch_impl(ws, src, src_stride, n_units, h, w, itemsize, bit_depth, T, n_bits, payload, payload_stride,
                            marked, marked_stride, lm, lm_stride, info, (cudaStream_t)stream);
}

int peeb_pee_extract_batch(peeb_ws* ws, const void* marked, int64_t marked_stride, int n_units, int h, int w,
                           int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* lm,
                           int64_t lm_stride, uint8_t* payload_out, int64_t payload_stride, void* recovered,
                           int64_t recovered_stride, int64_t* info, void* stream) {
    return extract_batch_impl(ws, marked, marked_stride, n_units, h, w, itemsize, bit_depth, T, n_bits, lm, lm_stride,
                              payload_out, payload_stride, recovered, recovered_stride, info, (cudaStream_t)stream);
}

int peeb_pee_hist_batch(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                        int bit_depth, uint32_t* hist, void* stream) {
    return hist_batch_impl(ws, src, src_stride, n_units, h, w, itemsize, bit_depth, hist, (cudaStream_t)stream);
}

// ---- host-buffer variants: staging + copies on the workspace stream, synchronous ----
static int64_t max_payload_bytes(int n_units, const int64_t* n_bits) {
    int64_t m = 0;
    for (int u = 0; u < n_units; ++u) {
        const int64_t b = (int64_t)peeb_payload_bytes(n_bits[u]);
        if (b > m) m = b;
    }
    return m;
}

// longest payload of a batch in bytes (the part of a payload row that has to cross the PCIe link)
static int64_t max_bits_bytes(int n_units, const int64_t* n_bits) {
    int64_t m = 0;
    for (int u = 0; u < n_units; ++u) m = std::max<int64_t>(m, (n_bits[u] + 7) / 8);
    return m;
}

// Units per chunk of a host batch: small enough that several chunks overlap their PCIe copies
// with each other's kernels, large enough to fill the GPU.
static int chunk_units(int n_units, size_t unit_bytes) {
    size_t target = 32u << 20;  // B200 + PCIe gen5: 32 MB measured best (scripts/e2e_chunk_sweep.sh); ~40 us of host work per chunk
    if (const char* e = getenv("PEEB_CHUNK_MB")) target = (size_t)atoi(e) << 20;  // tuning experiments
    long long c = (long long)((target + unit_bytes - 1) / unit_bytes);
    if (c < 1) c = 1;
    if (c > n_units) c = n_units;
    return (int)c;
}

// Chunk sizes (in units) of a host batch: uniform steady-state chunks.  PEEB_CHUNK_FIRST_MB=k ramps up
// from k MB chunks (and down again at the end) to shorten the first copy-in and the last copy-out, which
// overlap nothing; measured on B200 / PCIe gen5 it gains nothing (ct512 round trip 14.02 ms against
// 13.88 ms uniform, the bidirectional copy floor being 13.1-13.8 ms), so it is off by default.
static std::vector<int> chunk_plan(int n_units, size_t unit_bytes, int cmax_limit = 1 << 30) {
    const int cmax = std::min(chunk_units(n_units, unit_bytes), cmax_limit);
    size_t first = 0;
    if (const char* e = getenv("PEEB_CHUNK_FIRST_MB")) first = (size_t)atoi(e) << 20;  // 0: uniform chunks
    std::vector<int> up;
    if (first > 0) {
        const long long cmin = std::max<long long>(1, (long long)((first + unit_bytes - 1) / unit_bytes));
        long long used = 0;
        for (long long c = cmin; c < cmax && 2 * (used + c) + cmax <= n_units; c *= 2) { up.push_back((int)c); used += c; }
    }
    std::vector<int> plan(up);
    long long rest = n_units;
    for (int c : up) rest -= 2 * c;
    while (rest > 0) { const int c = (int)std::min<long long>(rest, cmax); plan.push_back(c); rest -= c; }
    plan.insert(plan.end(), up.rbegin(), up.rend());
    return plan;
}

static int peeb_pee_embed_h_impl(peeb_ws* ws, const void* src_host, int shared_flags, int n_units, int h, int w, int itemsize,
                     int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload_host,
                     int64_t payload_stride, void* marked_host, uint8_t* lm_host, int64_t* info_host) {
    PEEB_REQUIRE(ws && src_host && n_bits && info_host, "peeb_pee_embed_h: null pointer");  // T may be null
    PEEB_REQUIRE(n_units >= 1 && h >= 1 && w >= 1 && (itemsize == 1 || itemsize == 2), "peeb_pee_embed_h: bad sizes");
    PEEB_REQUIRE(payload_stride >= 0, "peeb_pee_embed_h: negative payload stride");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const int shared_src = shared_flags & 1, shared_pay = (shared_flags >> 1) & 1;
    const size_t img = (size_t)h * w * itemsize, img_al = align_up(img, 256);
    const size_t lmb = (size_t)h * ((w + 7) / 8), lm_al = align_up(lmb, 256);
    for (int u = 0; u < n_units; ++u)
        PEEB_REQUIRE(n_bits[u] >= 0 && (n_bits[u] + 7) / 8 <= payload_stride, "peeb_pee_embed_h: payload %d shorter than n_bits", u);
    const size_t pstride = align_up((size_t)std::max<int64_t>(payload_stride, max_payload_bytes(n_units, n_bits)), 16);
    const size_t n_src = shared_src ? 1 : (size_t)n_units;
    const size_t o_marked = n_src * img_al;
    int rc = scratch_reserve(ws->stage, o_marked + (marked_host ? (size_t)n_units * img_al : 0) + 256); if (rc) return rc;
    const size_t n_pay = shared_pay ? 1 : (size_t)n_units;
    const size_t o_lm = n_pay * pstride, o_info = o_lm + (size_t)n_units * lm_al;
    rc = scratch_reserve(ws->stage2, o_info + (size_t)n_units * PEEB_INFO * 8 + 256); if (rc) return rc;
    rc = scratch_reserve(ws->info_h, (size_t)n_units * PEEB_INFO * 8, true); if (rc) return rc;
    int64_t* info_pin = (int64_t*)ws->info_h.ptr;
    char* d1 = (char*)ws->stage.ptr;
    char* d2 = (char*)ws->stage2.ptr;
    cudaStream_t streams[2] = {ws->stream, ws->stream2};
    if (shared_src || shared_pay) {
        if (shared_src) PEEB_CUDA(cudaMemcpyAsync(d1, src_host, img, cudaMemcpyHostToDevice, streams[0]));
        if (shared_pay && payload_stride > 0 && payload_host)
            PEEB_CUDA(cudaMemcpyAsync(d2, payload_host, (size_t)payload_stride, cudaMemcpyHostToDevice, streams[0]));
        PEEB_CUDA(cudaEventRecord(ws->ev[1], streams[0]));
        PEEB_CUDA(cudaStreamWaitEvent(streams[1], ws->ev[1], 0));
    }
    // Two pipelines.  Roles (default): one stream per PCIe direction and one for the kernels, chained by
    // events, so the copies of each direction queue back to back.  Alternating: chunks alternate between
    // two streams, each doing its own copy-in, kernels and copy-out.
    const bool roles = ws->pipe_roles != 0;
    const std::vector<int> plan = chunk_plan(n_units, img);
    const size_t pay_width = std::min<size_t>((size_t)payload_stride, align_up((size_t)max_bits_bytes(n_units, n_bits), 16));
    for (int u0 = 0, c = 0; c < (int)plan.size(); u0 += plan[c], ++c) {
        const int n = plan[c], slot = c & 1;
        cudaStream_t st_in = roles ? ws->stream : streams[slot], st = roles ? ws->stream2 : streams[slot];
        cudaStream_t st_out = roles ? ws->stream3 : streams[slot];
        if (!shared_src)
            PEEB_CUDA(cudaMemcpy2DAsync(d1 + (size_t)u0 * img_al, img_al, (const char*)src_host + (size_t)u0 * img, img, img, n,
                                        cudaMemcpyHostToDevice, st_in));
        if (!shared_pay && pay_width > 0 && payload_host)  // only the bytes that hold payload bits travel
            PEEB_CUDA(cudaMemcpy2DAsync(d2 + (size_t)u0 * pstride, pstride, payload_host + (size_t)u0 * payload_stride,
                                        (size_t)payload_stride, pay_width, n, cudaMemcpyHostToDevice, st_in));
        if (roles) {
            cudaEvent_t e = ws->pipe_ev[(2 * c) % peeb_ws::kPipeEvents];
            PEEB_CUDA(cudaEventRecord(e, st_in));
            PEEB_CUDA(cudaStreamWaitEvent(st, e, 0));
        }
        rc = embed_batch_impl(ws, shared_src ? d1 : d1 + (size_t)u0 * img_al, shared_src ? 0 : (int64_t)img_al, n, h, w,
                              itemsize, bit_depth, T ? T + u0 : nullptr, n_bits + u0, (const uint8_t*)(shared_pay ? d2 : d2 + (size_t)u0 * pstride),
                              shared_pay ? 0 : (int64_t)pstride, marked_host ? d1 + o_marked + (size_t)u0 * img_al : nullptr, (int64_t)img_al,
                              lm_host ? (uint8_t*)(d2 + o_lm + (size_t)u0 * lm_al) : nullptr, (int64_t)lm_al,
                              (int64_t*)(d2 + o_info) + (size_t)u0 * PEEB_INFO, st, slot);
        if (rc) return rc;
        if (roles) {
            cudaEvent_t e = ws->pipe_ev[(2 * c + 1) % peeb_ws::kPipeEvents];
            PEEB_CUDA(cudaEventRecord(e, st));
            PEEB_CUDA(cudaStreamWaitEvent(st_out, e, 0));
        }
        if (marked_host)
            PEEB_CUDA(cudaMemcpy2DAsync((char*)marked_host + (size_t)u0 * img, img, d1 + o_marked + (size_t)u0 * img_al, img_al,
                                        img, n, cudaMemcpyDeviceToHost, st_out));
        if (lm_host)
            PEEB_CUDA(cudaMemcpy2DAsync(lm_host + (size_t)u0 * lmb, lmb, d2 + o_lm + (size_t)u0 * lm_al, lm_al, lmb, n,
                                        cudaMemcpyDeviceToHost, st_out));
        PEEB_CUDA(cudaMemcpyAsync(info_pin + (size_t)u0 * PEEB_INFO, (int64_t*)(d2 + o_info) + (size_t)u0 * PEEB_INFO,
                                  (size_t)n * PEEB_INFO * 8, cudaMemcpyDeviceToHost, st_out));
    }
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream2));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream3));
    memcpy(info_host, info_pin, (size_t)n_units * PEEB_INFO * 8);
    return PEEB_OK;
}

int peeb_pee_embed_h(peeb_ws* ws, const void* src_host, int shared_flags, int n_units, int h, int w, int itemsize,
                     int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload_host,
                     int64_t payload_stride, void* marked_host, uint8_t* lm_host, int64_t* info_host) {
    // a failed call returns only after the copies of earlier chunks have stopped touching the caller's buffers
    const int rc = peeb_pee_embed_h_impl(ws, src_host, shared_flags, n_units, h, w, itemsize, bit_depth, T, n_bits, payload_host, payload_stride, marked_host, lm_host, info_host);
    if (rc != PEEB_OK && ws) {
        cudaStreamSynchronize(ws->stream); cudaStreamSynchronize(ws->stream2); cudaStreamSynchronize(ws->stream3);
    }
    return rc;
}


static int peeb_pee_extract_h_impl(peeb_ws* ws, const void* marked_host, int n_units, int h, int w, int itemsize, int bit_depth,
                       const int32_t* T, const int64_t* n_bits, const uint8_t* lm_host, uint8_t* payload_out_host,
                       int64_t payload_stride, void* recovered_host, int64_t* info_host) {
    PEEB_REQUIRE(ws && marked_host && T && n_bits && lm_host && payload_out_host && info_host, "peeb_pee_extract_h: null pointer");
    PEEB_REQUIRE(n_units >= 1 && h >= 1 && w >= 1 && (itemsize == 1 || itemsize == 2), "peeb_pee_extract_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const size_t img = (size_t)h * w * itemsize, img_al = align_up(img, 256);
    const size_t lmb = (size_t)h * ((w + 7) / 8), lm_al = align_up(lmb, 256);
    for (int u = 0; u < n_units; ++u)
        PEEB_REQUIRE(n_bits[u] >= 0 && (n_bits[u] + 7) / 8 <= payload_stride, "peeb_pee_extract_h: payload_out %d shorter than n_bits", u);
    const size_t pstride = align_up((size_t)std::max<int64_t>(payload_stride, max_payload_bytes(n_units, n_bits)), 16);
    const size_t o_rec = (size_t)n_units * img_al;
    int rc = scratch_reserve(ws->stage, o_rec + (recovered_host ? (size_t)n_units * img_al : 0) + 256); if (rc) return rc;
    const size_t o_lm = (size_t)n_units * pstride, o_info = o_lm + (size_t)n_units * lm_al;
    rc = scratch_reserve(ws->stage2, o_info + (size_t)n_units * PEEB_INFO * 8 + 256); if (rc) return rc;
    rc = scratch_reserve(ws->info_h, (size_t)n_units * PEEB_INFO * 8, true); if (rc) return rc;
    int64_t* info_pin = (int64_t*)ws->info_h.ptr;
    char* d1 = (char*)ws->stage.ptr;
    char* d2 = (char*)ws->stage2.ptr;
    cudaStream_t streams[2] = {ws->stream, ws->stream2};
    const bool roles = ws->pipe_roles != 0;  // see peeb_pee_embed_h
    const std::vector<int> plan = chunk_plan(n_units, img, 65535);
    const size_t pay_width = std::min<size_t>((size_t)payload_stride, align_up((size_t)max_bits_bytes(n_units, n_bits), 16));
    for (int u0 = 0, c = 0; c < (int)plan.size(); u0 += plan[c], ++c) {
        const int n = plan[c], slot = c & 1;
        cudaStream_t st_in = roles ? ws->stream : streams[slot], st = roles ? ws->stream2 : streams[slot];
        cudaStream_t st_out = roles ? ws->stream3 : streams[slot];
        PEEB_CUDA(cudaMemcpy2DAsync(d1 + (size_t)u0 * img_al, img_al, (const char*)marked_host + (size_t)u0 * img, img, img, n,
                                    cudaMemcpyHostToDevice, st_in));
        PEEB_CUDA(cudaMemcpy2DAsync(d2 + o_lm + (size_t)u0 * lm_al, lm_al, lm_host + (size_t)u0 * lmb, lmb, lmb, n,
                                    cudaMemcpyHostToDevice, st_in));
        if (roles) {
            cudaEvent_t e = ws->pipe_ev[(2 * c) % peeb_ws::kPipeEvents];
            PEEB_CUDA(cudaEventRecord(e, st_in));
            PEEB_CUDA(cudaStreamWaitEvent(st, e, 0));
        }
        rc = extract_batch_impl(ws, d1 + (size_t)u0 * img_al, (int64_t)img_al, n, h, w, itemsize, bit_depth, T + u0, n_bits + u0,
                                (const uint8_t*)(d2 + o_lm + (size_t)u0 * lm_al), (int64_t)lm_al,
                                (uint8_t*)(d2 + (size_t)u0 * pstride), (int64_t)pstride,
                                recovered_host ? d1 + o_rec + (size_t)u0 * img_al : nullptr, (int64_t)img_al,
                                (int64_t*)(d2 + o_info) + (size_t)u0 * PEEB_INFO, st, slot);
        if (rc) return rc;
        if (roles) {
            cudaEvent_t e = ws->pipe_ev[(2 * c + 1) % peeb_ws::kPipeEvents];
            PEEB_CUDA(cudaEventRecord(e, st));
            PEEB_CUDA(cudaStreamWaitEvent(st_out, e, 0));
        }
        if (recovered_host)
            PEEB_CUDA(cudaMemcpy2DAsync((char*)recovered_host + (size_t)u0 * img, img, d1 + o_rec + (size_t)u0 * img_al, img_al,
                                        img, n, cudaMemcpyDeviceToHost, st_out));
        if (pay_width > 0)  // the bytes that can hold payload bits; the rest of every row is zero filled on the host below
            PEEB_CUDA(cudaMemcpy2DAsync(payload_out_host + (size_t)u0 * payload_stride, (size_t)payload_stride,
                                        d2 + (size_t)u0 * pstride, pstride, pay_width, n, cudaMemcpyDeviceToHost, st_out));
        PEEB_CUDA(cudaMemcpyAsync(info_pin + (size_t)u0 * PEEB_INFO, (int64_t*)(d2 + o_info) + (size_t)u0 * PEEB_INFO,
                                  (size_t)n * PEEB_INFO * 8, cudaMemcpyDeviceToHost, st_out));
    }
    // zero padding of the rows past the copied bytes (while the copies are still in flight: other bytes)
    if ((size_t)payload_stride > pay_width)
        for (int u = 0; u < n_units; ++u)
            memset(payload_out_host + (size_t)u * payload_stride + pay_width, 0, (size_t)payload_stride - pay_width);
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream2));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream3));
    memcpy(info_host, info_pin, (size_t)n_units * PEEB_INFO * 8);
    return PEEB_OK;
}

int peeb_pee_extract_h(peeb_ws* ws, const void* marked_host, int n_units, int h, int w, int itemsize, int bit_depth,
                       const int32_t* T, const int64_t* n_bits, const uint8_t* lm_host, uint8_t* payload_out_host,
                       int64_t payload_stride, void* recovered_host, int64_t* info_host) {
    // a failed call returns only after the copies of earlier chunks have stopped touching the caller's buffers
    const int rc = peeb_pee_extract_h_impl(ws, marked_host, n_units, h, w, itemsize, bit_depth, T, n_bits, lm_host, payload_out_host, payload_stride, recovered_host, info_host);
    if (rc != PEEB_OK && ws) {
        cudaStreamSynchronize(ws->stream); cudaStreamSynchronize(ws->stream2); cudaStreamSynchronize(ws->stream3);
    }
    return rc;
}


int peeb_pee_hist_h(peeb_ws* ws, const void* src_host, int n_units, int h, int w, int itemsize, int bit_depth,
                    uint32_t* hist_host) {
    PEEB_REQUIRE(ws && src_host && hist_host, "peeb_pee_hist_h: null pointer");
    PEEB_REQUIRE(n_units >= 1 && h >= 1 && w >= 1 && (itemsize == 1 || itemsize == 2), "peeb_pee_hist_h: bad sizes");
    PEEB_REQUIRE(bit_depth >= 1 && bit_depth <= 8 * itemsize, "peeb_pee_hist_h: bad bit_depth");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = ws->stream;
    const size_t img = (size_t)h * w * itemsize;
    const size_t hb = (size_t)4 * (1u << (bit_depth - 1)) * sizeof(uint32_t) * n_units;
    int rc = scratch_reserve(ws->stage, img * n_units + 256); if (rc) return rc;
    rc = scratch_reserve(ws->stage2, hb + 256); if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(ws->stage.ptr, src_host, img * n_units, cudaMemcpyHostToDevice, st));
    rc = hist_batch_impl(ws, ws->stage.ptr, (int64_t)img, n_units, h, w, itemsize, bit_depth, (uint32_t*)ws->stage2.ptr, st);
    if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(hist_host, ws->stage2.ptr, hb, cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaStreamSynchronize(st));
    return PEEB_OK;
}

}  // extern "C"
