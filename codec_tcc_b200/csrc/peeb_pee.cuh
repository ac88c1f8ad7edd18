// Shared between the PEE translation units (peeb_pee.cu: 4-pixel-per-lane kernels and the host side;
// peeb_pee8.cu: the 8-pixel-per-lane two-phase kernels): geometry, batch descriptor, band staging.
#pragma once
#include "peeb_common.cuh"

namespace peeb {

constexpr unsigned long long ST_AGG = 1ull << 62, ST_PFX = 2ull << 62, ST_MASK = 3ull << 62;

constexpr int STRIP = 128;  // columns per warp work item (4 per lane)


struct PeeGeom {
    int h, w, itemsize;
    int R;         // band height (rows written per CTA)
    int nb;        // bands per unit
    int S;         // strips per row
    int pitch;     // shared-memory row pitch (bytes)
    int rowbytes;  // w * itemsize
    int lmw;       // ceil(w/8): global location-map row bytes
    int lmpitch;   // shared location-map row pitch (bytes, multiple of 4)
    int bulk;      // TMA bulk copies usable (rowbytes % 16 == 0)
    int maxval;
    int bandwords; // extract staging: 32-bit words per (unit, pass, band)
    int threads;   // CTA size: 256 (4 CTAs/SM), 512 (2) or 1024 (1), by how much of an SM's shared memory a band needs
};

struct PeeBatch {
    const unsigned char* src; long long src_stride;    // bytes
    unsigned char* dst; long long dst_stride;          // may be null
    unsigned char* lm; long long lm_stride;            // may be null (embed) / const (extract)
    const unsigned char* payload; long long payload_stride;
    unsigned char* payload_out;                        // extract
    const int* T;                                      // device, per unit
    const unsigned* n_bits;                            // device, per unit
    long long* info;                                   // device, per unit x 8
    int n_units;
};

// location-map bit of column j inside a 32-bit little-endian word of a packbits row
__device__ __forceinline__ unsigned lm_bitmask(int j) { return 1u << (8 * ((j >> 3) & 3) + 7 - (j & 7)); }

// Cooperative copy of image rows [lo, hi) of a unit into the band buffer.
template <typename PixT>
__device__ __forceinline__ void load_rows(const PeeGeom& g, const unsigned char* unit_src, unsigned char* simg,
                                          int r_first /* image row of smem row 0 */, int lo, int hi,
                                          uint64_t* bar) {
    if (hi <= lo) return;
    unsigned char* dst = simg + (size_t)(lo - r_first) * g.pitch;
    const unsigned char* src = unit_src + (size_t)lo * g.rowbytes;
    if (g.bulk) {
        // rows are contiguous in both spaces (pitch == rowbytes): bulk copies of <= 64 KB
        const unsigned total = (unsigned)(hi - lo) * (unsigned)g.rowbytes;
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, total);
            for (unsigned off = 0; off < total; off += 65536u) {
                const unsigned n = min(65536u, total - off);
                bulk_g2s(dst + off, src + off, n, bar);
            }
        }
        mbar_wait(bar, 0);
    } else {
        const int n = (hi - lo);
        for (int r = threadIdx.x >> 5; r < n; r += blockDim.x >> 5) {
            const PixT* s = reinterpret_cast<const PixT*>(src + (size_t)r * g.rowbytes);
            PixT* d = reinterpret_cast<PixT*>(dst + (size_t)r * g.pitch);
            for (int c = threadIdx.x & 31; c < g.w; c += 32) d[c] = s[c];
        }
        __syncthreads();
    }
}

template <typename PixT>
__device__ __forceinline__ void store_rows(const PeeGeom& g, unsigned char* unit_dst, const unsigned char* simg,
                                           int r_first, int lo, int hi) {
    if (hi <= lo) return;
    const unsigned char* src = simg + (size_t)(lo - r_first) * g.pitch;
    unsigned char* dst = unit_dst + (size_t)lo * g.rowbytes;
    if (g.bulk) {
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned total = (unsigned)(hi - lo) * (unsigned)g.rowbytes;
            for (unsigned off = 0; off < total; off += 65536u) bulk_s2g(dst + off, src + off, min(65536u, total - off));
            bulk_commit();
            bulk_wait_read0();
        }
    } else {
        __syncthreads();
        const int n = (hi - lo);
        for (int r = threadIdx.x >> 5; r < n; r += blockDim.x >> 5) {
            const PixT* s = reinterpret_cast<const PixT*>(src + (size_t)r * g.pitch);
            PixT* d = reinterpret_cast<PixT*>(dst + (size_t)r * g.rowbytes);
            for (int c = threadIdx.x & 31; c < g.w; c += 32) d[c] = s[c];
        }
    }
}

// Expands payload bits [first, first + count) of an MSB-first packed stream into one byte
// per bit in shared memory (bits at or past n_bits read as 0, Appendix A's zero padding).
// The stream must be readable 8 bytes past the word holding bit n_bits-1.
__device__ __forceinline__ void expand_payload(const unsigned* __restrict__ pay, unsigned first, int count,
                                               unsigned n_bits, unsigned char* out /* 4-byte aligned */) {
    const unsigned sh = first & 31, w0 = first >> 5;
    unsigned* out4 = reinterpret_cast<unsigned*>(out);
    for (int j = threadIdx.x; j * 32 < count; j += blockDim.x) {
        const unsigned start = first + 32u * (unsigned)j;  // stream index of this thread's first bit
        unsigned win = 0;
        if (start < n_bits) {
            const unsigned a = __byte_perm(__ldg(pay + w0 + j), 0, 0x0123);
            const unsigned b = __byte_perm(__ldg(pay + w0 + j + 1), 0, 0x0123);
            win = __funnelshift_l(b, a, sh);                 // bit `start` is the MSB
            if (n_bits - start < 32u) win &= ~(0xffffffffu >> (n_bits - start));
        }
        const unsigned lsb = __brev(win);                    // bit `start + k` at bit k
#pragma unroll
        for (int n = 0; n < 8; ++n)
            out4[j * 8 + n] = (((lsb >> (4 * n)) & 0xfu) * 0x00204081u) & 0x01010101u;
    }
}

// host side shared by the two translation units
int upload_unit_tables(peeb_ws* ws, int slot, int n_units, const int32_t* T, const int64_t* n_bits, int bit_depth,
                       size_t extra_bytes, cudaStream_t st, int** dT, unsigned** dN, char** extra);
int embed_batch_impl2(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                      int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload,
                      int64_t payload_stride, void* marked, int64_t marked_stride, uint8_t* lm, int64_t lm_stride,
                      int64_t* info, cudaStream_t st, int slot);
int extract_batch_impl2(peeb_ws* ws, const void* marked, int64_t marked_stride, int n_units, int h, int w,
                        int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* lm,
                        int64_t lm_stride, uint8_t* payload_out, int64_t payload_stride, void* recovered,
                        int64_t recovered_stride, int64_t* info, cudaStream_t st, int slot);

}  // namespace peeb
