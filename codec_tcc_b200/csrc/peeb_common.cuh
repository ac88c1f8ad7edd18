// Shared device/host helpers for libpeeb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/peeb200.h"

namespace peeb {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PEEB_CUDA(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return ::peeb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define PEEB_REQUIRE(cond, ...)                                                \
    do {                                                                       \
        if (!(cond)) { ::peeb::set_error(__VA_ARGS__); return PEEB_E_INVALID; } \
    } while (0)

// ---------------------------------------------------------------- workspace
// One per (device, caller thread).  Owns grow-only device scratch, a pinned
// host mirror for small tables, two streams and a few events.  Thread
// compatible, not thread safe.
struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
};

}  // namespace peeb

struct peeb_ws {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;   // used by the *_h (host buffer) entry points
    cudaStream_t stream2 = nullptr;  // second stream for copy/compute overlap
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t stream3 = nullptr;  // host batches: copy-out stream (stream = copy-in, stream2 = kernels)
    static constexpr int kPipeEvents = 32;
    cudaEvent_t pipe_ev[kPipeEvents] = {};  // ring: copy-in done / kernels done, per chunk of a host batch
    int pipe_roles = 1;              // 1: one stream per direction + one for kernels; 0: chunks alternate between two streams
    peeb::Scratch tables;            // device: per-launch tables (tickets, status, counts)
    peeb::Scratch tables_h;          // pinned host mirror
    peeb::Scratch stage;             // device: staging for *_h entry points
    peeb::Scratch stage2;            // device: second staging area
    peeb::Scratch bits;              // device: compaction scratch (decode_message)
    // PEE launches use one of two independent table sets, so that two chunks of a host batch
    // can be in flight on the two streams at once
    peeb::Scratch ptables[2];        // device: T/n_bits per unit, band counts, look-back status, tickets
    static constexpr int kTableRing = 4;
    peeb::Scratch ptables_h[2][kTableRing];  // pinned host mirrors of the per-unit tables: a ring per table set, so that a call
                                             // only waits for the upload issued four calls earlier (the host can run ahead)
    int ptable_seq[2] = {0, 0};
    void* ptable_h_cur[2] = {nullptr, nullptr};  // the mirror the latest upload used (callers keep a pinned word behind it)
    peeb::Scratch pbits[2];          // device: extract per-band bit staging
    cudaEvent_t pev[2][kTableRing] = {};
    peeb::Scratch info_h;            // pinned landing zone for per-unit info rows (keeps every copy of a
                                     // host batch asynchronous even when the caller's info array is pageable)
    int use_bulk = 1;                // TMA bulk copies (PEEB_NO_BULK=1 disables)
    peeb::Scratch step_counters;     // device: 3 x uint64 counters of the PEE embed kernel's step kinds
    peeb::Scratch hist;              // device: prediction-error histograms of a batch (threshold selection)
    int step_counters_on = 0;
    // profiling: accumulate per-kernel device time with events when enabled
    int prof_on = 0;
    float prof_ms[PEEB_PROF_SLOTS] = {0};
    long long prof_calls[PEEB_PROF_SLOTS] = {0};
    cudaEvent_t prof_ev[2] = {nullptr, nullptr};
};

namespace peeb {

int scratch_reserve(Scratch& s, size_t bytes, bool pinned_host = false);
void scratch_free(Scratch& s, bool pinned_host = false);

// RAII-less profiling bracket used around kernel launches inside the library.
struct ProfScope {
    peeb_ws* ws; int slot; cudaStream_t st;
    ProfScope(peeb_ws* w, int s, cudaStream_t stream) : ws(w), slot(s), st(stream) {
        if (ws && ws->prof_on) cudaEventRecord(ws->prof_ev[0], st);
    }
    ~ProfScope() {
        if (ws && ws->prof_on) {
            cudaEventRecord(ws->prof_ev[1], st);
            cudaEventSynchronize(ws->prof_ev[1]);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ws->prof_ev[0], ws->prof_ev[1]);
            ws->prof_ms[slot] += ms;
            ws->prof_calls[slot] += 1;
        }
    }
};

__host__ __device__ static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- device
#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 128-bit streaming global accesses (read once / write once data)
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(int4* p, const int4& v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- mbarrier + TMA bulk (1-D) copies: cp.async.bulk, SASS UBLKCP ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");  // (suspend-time hint: sleep, do not spin)
}
// global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global (bulk async-group)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy before a bulk store
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Exclusive scan of data[0..n) in place by the whole block (any n); returns the
// total to every thread.  warp_sums: >= 33 ints of shared memory.
__device__ __forceinline__ int block_excl_scan(int* data, int n, int* warp_sums) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int idx = base + tid;
        const int v = idx < n ? data[idx] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int ws = lane < nwarps ? warp_sums[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sums[lane] = ws;  // inclusive over warps
        }
        __syncthreads();
        const int before = warp > 0 ? warp_sums[warp - 1] : 0;
        if (idx < n) data[idx] = carry + before + incl - v;
        const int tile_total = warp_sums[nwarps - 1];
        __syncthreads();
        carry += tile_total;
    }
    return carry;
}

#endif  // __CUDACC__

}  // namespace peeb
