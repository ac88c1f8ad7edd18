// N2: coding of the side bitmaps on the device.
//
// The reference ships the XOR side bitmaps of its LSB embedders as zlib over ONE BYTE PER PIXEL
// (src/codec.py:888-889 `zlib.compress(np.stack(bitmaps).tobytes())`, inverse :820-821) -- s*h*w bytes through
// a serial host coder, the next host bottleneck after the hot path.  Deflate is not available on the device
// offline; what the maps need is bit packing (8x) and the removal of their long zero runs (a 304-bit message
// touches 304 of 36 M positions; a PEE location map flags a few hundred pixels).  Format "PBR1", three levels
// of 32-way zero-run elimination over the packed bits:
//
//   L0[j]  32 elements per 32-bit word, bytes in np.packbits order (element 32j+8k+i -> bit 7-i of byte k)
//   L1[i]  bit t (LSB first) = (L0[32i+t] != 0)           L2[k]  bit t = (L1[32k+t] != 0)
//   C0[k]  number of non-zero L0 words under L2[k] (<= 1024): lets the decoder find every CTA's words with one scan
//   blob = "PBR1" | u32 0 | u64 n | u32 nz1 | u32 nz0 | L2[all] | C0[all] | non-zero L1 words | non-zero L0 words   (LE)
//
// A run of 1024 zero elements costs one bit, a run of 32768 eight bytes; a dense random map costs n/8 * (1 + 1/32
// + 1/512) bytes.  One CTA of 1024 threads owns one L2 word = 32768 elements: warp ballots give L1 and L2,
// the two compactions take one small scan over the CTAs' counts.  HBM-bound: the encoder reads one byte per
// element (or one bit, for maps that are already packed such as the PEE location map) and writes <= 1/8.
// (The CPU restatement the tests check this against lives with the test infrastructure.)
#include "peeb_common.cuh"

namespace peeb {

constexpr int PBR_HEADER = 24;
constexpr int PBR_CTA = 1024;

struct PbrSizes {
    long long n, n0, n1, n2;
};
static inline PbrSizes pbr_sizes(long long n) {
    PbrSizes s;
    s.n = n;
    s.n0 = (n + 31) / 32;
    s.n1 = (s.n0 + 31) / 32;
    s.n2 = (s.n1 + 31) / 32;
    return s;
}

// four bytes of a map -> four bits, MSB first (byte 0 -> bit 3): non-zero test per byte, then one multiply
// gathers the four flags (the partial products land on distinct bits: no carries)
__device__ __forceinline__ unsigned nibble_of(unsigned x) {
    const unsigned m = ((x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u) >> 7;
    return ((m * 0x08040201u) >> 24) & 0xfu;
}
// four bits -> four 0/1 bytes, MSB first
__device__ __forceinline__ unsigned bytes_of(unsigned nib) { return ((nib * 0x08040201u) >> 3) & 0x01010101u; }

// the L0 word of elements [32j, 32j + 32) of a byte map (elements >= n read as 0)
template <bool VEC>
__device__ __forceinline__ unsigned l0_from_bytes(const unsigned char* __restrict__ src, long long j, long long n) {
    const long long e0 = j * 32;
    if (e0 >= n) return 0u;
    unsigned w = 0;
    if (VEC && e0 + 32 <= n) {
        const int4 a = ldg_stream(reinterpret_cast<const int4*>(src + e0));
        const int4 b = ldg_stream(reinterpret_cast<const int4*>(src + e0) + 1);
        w = (nibble_of(a.x) << 4 | nibble_of(a.y)) | (nibble_of(a.z) << 4 | nibble_of(a.w)) << 8 |
            (nibble_of(b.x) << 4 | nibble_of(b.y)) << 16 | (nibble_of(b.z) << 4 | nibble_of(b.w)) << 24;
    } else {
        const int m = (int)min(32ll, n - e0);
        for (int i = 0; i < m; ++i)
            if (src[e0 + i]) w |= 1u << (8 * (i >> 3) + 7 - (i & 7));
    }
    return w;
}
// ... of a packed map (np.packbits bytes; bits past n are masked off)
__device__ __forceinline__ unsigned l0_from_packed(const unsigned char* __restrict__ src, long long j, long long n,
                                                   bool aligned) {
    const long long e0 = j * 32;
    if (e0 >= n) return 0u;
    unsigned w = 0;
    const long long nbytes = (n + 7) / 8;
    if (aligned && 4 * j + 4 <= nbytes) w = __ldg(reinterpret_cast<const unsigned*>(src) + j);
    else
        for (int k = 0; k < 4 && 4 * j + k < nbytes; ++k) w |= (unsigned)src[4 * j + k] << (8 * k);
    const long long left = n - e0;  // valid elements of this word
    if (left < 32) {
        unsigned keep = 0;
        for (int i = 0; i < (int)left; ++i) keep |= 1u << (8 * (i >> 3) + 7 - (i & 7));
        w &= keep;
    }
    return w;
}

// K1: L0 and L1 to scratch, L2 into the blob, per-CTA counts of non-zero L1 / L0 words
template <bool PACKED, bool VEC>
__global__ void __launch_bounds__(PBR_CTA) pbr_pack_kernel(const unsigned char* __restrict__ src, long long n, long long n0,
                                                           long long n1, unsigned* __restrict__ l0, unsigned* __restrict__ l1,
                                                           unsigned* __restrict__ l2, int2* __restrict__ counts) {
    __shared__ unsigned s_l1[32];
    const long long b = blockIdx.x, j = b * PBR_CTA + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned w = PACKED ? l0_from_packed(src, j, n, VEC) : l0_from_bytes<VEC>(src, j, n);
    if (j < n0) l0[j] = w;
    const unsigned p1 = __ballot_sync(0xffffffffu, w != 0u);
    if (lane == 0) {
        s_l1[warp] = p1;
        if (b * 32 + warp < n1) l1[b * 32 + warp] = p1;
    }
    __syncthreads();
    if (warp == 0) {
        const unsigned v = s_l1[lane];
        const unsigned p2 = __ballot_sync(0xffffffffu, v != 0u);
        int c0 = __popc(v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        if (lane == 0) {
            l2[b] = p2;
            l2[gridDim.x + b] = (unsigned)c0;   // C0 follows L2 in the blob (gridDim.x = number of L2 words)
            counts[b] = make_int2(__popc(p2), c0);
        }
    }
}

// K2 (one CTA): exclusive scans of the counts -> offsets; totals into the header and a result slot
__global__ void __launch_bounds__(1024) pbr_scan_kernel(int2* __restrict__ counts, long long n2, long long n,
                                                        unsigned* __restrict__ header, unsigned* __restrict__ totals) {
    __shared__ unsigned s_w[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned carry1 = 0, carry0 = 0;
    for (long long base = 0; base < n2; base += 1024) {
        const long long k = base + threadIdx.x;
        const int2 c = k < n2 ? counts[k] : make_int2(0, 0);
        unsigned i1 = (unsigned)c.x, i0 = (unsigned)c.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t1 = __shfl_up_sync(0xffffffffu, i1, o), t0 = __shfl_up_sync(0xffffffffu, i0, o);
            if (lane >= o) { i1 += t1; i0 += t0; }
        }
        if (lane == 31) { s_w[0][warp] = i1; s_w[1][warp] = i0; }
        __syncthreads();
        unsigned b1 = 0, b0 = 0, t1 = 0, t0 = 0;
        for (int q = 0; q < 32; ++q) {
            const unsigned a1 = s_w[0][q], a0 = s_w[1][q];
            if (q < warp) { b1 += a1; b0 += a0; }
            t1 += a1; t0 += a0;
        }
        if (k < n2) counts[k] = make_int2((int)(carry1 + b1 + i1 - (unsigned)c.x), (int)(carry0 + b0 + i0 - (unsigned)c.y));
        carry1 += t1; carry0 += t0;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        header[0] = 0x31524250u;  // "PBR1"
        header[1] = 0u;
        header[2] = (unsigned)((unsigned long long)n & 0xffffffffull);
        header[3] = (unsigned)((unsigned long long)n >> 32);
        header[4] = carry1;
        header[5] = carry0;
        totals[0] = carry1;
        totals[1] = carry0;
    }
}

// K3: the non-zero L1 and L0 words of a CTA, in order, behind those of the CTAs before it
__global__ void __launch_bounds__(PBR_CTA) pbr_compact_kernel(const unsigned* __restrict__ l0, const unsigned* __restrict__ l1,
                                                              long long n0, long long n1, long long n2,
                                                              const int2* __restrict__ offs, const unsigned* __restrict__ totals,
                                                              unsigned* __restrict__ body /* behind the header */) {
    const long long b = blockIdx.x, j = b * PBR_CTA + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int2 off = offs[b];
    const unsigned nz1 = totals[0];
    unsigned* out1 = body + 2 * n2;
    unsigned* out0 = body + 2 * n2 + nz1;
    const unsigned mine = (b * 32 + lane < n1) ? l1[b * 32 + lane] : 0u;  // lane k: L1 word k of this CTA
    int before = lane < warp ? __popc(mine) : 0;                          // non-zero L0 words of the warps before this one
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    const unsigned p1 = __shfl_sync(0xffffffffu, mine, warp);
    const unsigned w = j < n0 ? l0[j] : 0u;
    if (w != 0u) out0[(unsigned)off.y + (unsigned)before + __popc(p1 & lanemask_lt())] = w;
    if (warp == 0) {
        const unsigned p2 = __ballot_sync(0xffffffffu, mine != 0u);
        if (mine != 0u) out1[(unsigned)off.x + __popc(p2 & lanemask_lt())] = mine;
    }
}

// D1 (one CTA): per L2 word, where its non-zero L1 words and their non-zero L0 words start = exclusive scans of
// popc(L2[k]) and C0[k]; err = the totals disagree with the header
__global__ void __launch_bounds__(1024) pbr_plan_kernel(const unsigned* __restrict__ body, long long n2, unsigned nz1,
                                                        unsigned nz0, int2* __restrict__ offs, int* __restrict__ err) {
    __shared__ unsigned s_w[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned carry1 = 0, carry0 = 0;
    for (long long base = 0; base < n2; base += 1024) {
        const long long k = base + threadIdx.x;
        const unsigned c1 = k < n2 ? (unsigned)__popc(body[k]) : 0u;
        unsigned c0 = k < n2 ? body[n2 + k] : 0u;
        if (c0 > 1024u) { *err = 1; c0 = 0u; }
        unsigned i1 = c1, i0 = c0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t1 = __shfl_up_sync(0xffffffffu, i1, o), t0 = __shfl_up_sync(0xffffffffu, i0, o);
            if (lane >= o) { i1 += t1; i0 += t0; }
        }
        if (lane == 31) { s_w[0][warp] = i1; s_w[1][warp] = i0; }
        __syncthreads();
        unsigned b1 = 0, b0 = 0, t1 = 0, t0 = 0;
        for (int q = 0; q < 32; ++q) {
            const unsigned a1 = s_w[0][q], a0 = s_w[1][q];
            if (q < warp) { b1 += a1; b0 += a0; }
            t1 += a1; t0 += a0;
        }
        if (k < n2) offs[k] = make_int2((int)(carry1 + b1 + i1 - c1), (int)(carry0 + b0 + i0 - c0));
        carry1 += t1; carry0 += t0;
        __syncthreads();
    }
    if (threadIdx.x == 0 && (carry1 != nz1 || carry0 != nz0)) *err = 1;
}

// D2: one CTA per L2 word -> its 32768 elements as 0/1 bytes or as packed bits
template <bool PACKED, bool VEC>
__global__ void __launch_bounds__(PBR_CTA) pbr_expand_kernel(const unsigned* __restrict__ body, long long n, long long n0,
                                                             long long n2, unsigned nz1, unsigned nz0,
                                                             const int2* __restrict__ offs, unsigned char* __restrict__ dst,
                                                             int* __restrict__ err) {
    const long long b = blockIdx.x, j = b * PBR_CTA + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned* nzl1 = body + 2 * n2;
    const unsigned* nzl0 = body + 2 * n2 + nz1;
    const unsigned p2 = body[b];
    const int2 off = offs[b];
    unsigned mine = 0;  // lane k: L1 word k of this CTA
    if ((p2 >> lane) & 1u) {
        const unsigned at = (unsigned)off.x + __popc(p2 & lanemask_lt());
        if (at < nz1) mine = nzl1[at]; else *err = 1;
    }
    int before = lane < warp ? __popc(mine) : 0, all = __popc(mine);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        all += __shfl_xor_sync(0xffffffffu, all, o);
    }
    if (threadIdx.x == 0 && (unsigned)all != body[n2 + b]) *err = 1;  // the count the offsets were built from
    const unsigned p1 = __shfl_sync(0xffffffffu, mine, warp);
    unsigned w = 0;
    if ((p1 >> lane) & 1u) {
        const unsigned at = (unsigned)off.y + (unsigned)before + __popc(p1 & lanemask_lt());
        if (at < nz0) w = nzl0[at]; else *err = 1;
    }
    if (j >= n0) return;
    const long long e0 = j * 32;
    if (PACKED) {
        const long long nbytes = (n + 7) / 8;
        if (VEC && 4 * j + 4 <= nbytes) reinterpret_cast<unsigned*>(dst)[j] = w;
        else
            for (int k = 0; k < 4 && 4 * j + k < nbytes; ++k) dst[4 * j + k] = (unsigned char)(w >> (8 * k));
    } else if (VEC && e0 + 32 <= n) {
        int4 a, c;
        a.x = (int)bytes_of((w >> 4) & 0xfu);  a.y = (int)bytes_of(w & 0xfu);
        a.z = (int)bytes_of((w >> 12) & 0xfu); a.w = (int)bytes_of((w >> 8) & 0xfu);
        c.x = (int)bytes_of((w >> 20) & 0xfu); c.y = (int)bytes_of((w >> 16) & 0xfu);
        c.z = (int)bytes_of((w >> 28) & 0xfu); c.w = (int)bytes_of((w >> 24) & 0xfu);
        stg_stream(reinterpret_cast<int4*>(dst + e0), a);
        stg_stream(reinterpret_cast<int4*>(dst + e0) + 1, c);
    } else {
        const int m = (int)min(32ll, n - e0);
        for (int i = 0; i < m; ++i) dst[e0 + i] = (unsigned char)((w >> (8 * (i >> 3) + 7 - (i & 7))) & 1u);
    }
}

static int parse_header(const unsigned char* h, int64_t blob_bytes, int64_t n, PbrSizes& sz, unsigned& nz1, unsigned& nz0) {
    PEEB_REQUIRE(blob_bytes >= PBR_HEADER && memcmp(h, "PBR1", 4) == 0, "peeb_bitmap_decode: not a PBR1 blob");
    unsigned f[6];
    memcpy(f, h, sizeof f);
    PEEB_REQUIRE(f[1] == 0u, "peeb_bitmap_decode: unknown PBR1 flags %u", f[1]);
    const unsigned long long nn = (unsigned long long)f[2] | ((unsigned long long)f[3] << 32);
    PEEB_REQUIRE((long long)nn == n, "peeb_bitmap_decode: the blob holds %llu elements, %lld expected", nn, (long long)n);
    sz = pbr_sizes(n);
    nz1 = f[4]; nz0 = f[5];
    PEEB_REQUIRE((long long)nz1 <= sz.n1 && (long long)nz0 <= sz.n0, "peeb_bitmap_decode: corrupt header");
    PEEB_REQUIRE(blob_bytes == PBR_HEADER + 4 * (2 * sz.n2 + (long long)nz1 + (long long)nz0),
                 "peeb_bitmap_decode: blob size %lld does not match its header", (long long)blob_bytes);
    return PEEB_OK;
}

static int decode_dev(peeb_ws* ws, const unsigned char* blob, const PbrSizes& sz, unsigned nz1, unsigned nz0,
                      unsigned char* dst, int packed_output, cudaStream_t st) {
    int rc = scratch_reserve(ws->bits, align_up((size_t)sz.n2 * sizeof(int2), 256) + 256);
    if (rc) return rc;
    int2* offs = (int2*)ws->bits.ptr;
    int* err = (int*)((char*)ws->bits.ptr + align_up((size_t)sz.n2 * sizeof(int2), 256));
    rc = scratch_reserve(ws->tables_h, 256, true);
    if (rc) return rc;
    int* err_h = (int*)ws->tables_h.ptr;
    const unsigned* body = (const unsigned*)(blob + PBR_HEADER);
    PEEB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    {
        ProfScope p(ws, PEEB_K_BITMAP_DECODE, st);
        pbr_plan_kernel<<<1, 1024, 0, st>>>(body, sz.n2, nz1, nz0, offs, err);
        const bool vec = ((uintptr_t)dst & 15) == 0;
        const unsigned grid = (unsigned)sz.n2;
        if (packed_output) {
            if (vec) pbr_expand_kernel<true, true><<<grid, PBR_CTA, 0, st>>>(body, sz.n, sz.n0, sz.n2, nz1, nz0, offs, dst, err);
            else pbr_expand_kernel<true, false><<<grid, PBR_CTA, 0, st>>>(body, sz.n, sz.n0, sz.n2, nz1, nz0, offs, dst, err);
        } else {
            if (vec) pbr_expand_kernel<false, true><<<grid, PBR_CTA, 0, st>>>(body, sz.n, sz.n0, sz.n2, nz1, nz0, offs, dst, err);
            else pbr_expand_kernel<false, false><<<grid, PBR_CTA, 0, st>>>(body, sz.n, sz.n0, sz.n2, nz1, nz0, offs, dst, err);
        }
    }
    PEEB_CUDA(cudaGetLastError());
    PEEB_CUDA(cudaMemcpyAsync(err_h, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaStreamSynchronize(st));
    PEEB_REQUIRE(*err_h == 0, "peeb_bitmap_decode: corrupt blob (the level tables disagree with the header)");
    return PEEB_OK;
}

}  // namespace peeb

using namespace peeb;

extern "C" {

size_t peeb_bitmap_blob_bound(int64_t n) {
    if (n < 0) n = 0;
    const PbrSizes sz = pbr_sizes(n);
    return (size_t)PBR_HEADER + 4 * (size_t)(2 * sz.n2 + sz.n1 + sz.n0);
}

int peeb_bitmap_encode(peeb_ws* ws, const uint8_t* src, int64_t n, int packed_input, uint8_t* blob,
                       int64_t blob_capacity, int64_t* blob_bytes, void* stream) {
    PEEB_REQUIRE(ws && blob && blob_bytes && (src || n == 0), "peeb_bitmap_encode: null pointer");
    PEEB_REQUIRE(n >= 0 && n < (1ll << 40), "peeb_bitmap_encode: bad element count");
    PEEB_REQUIRE(((uintptr_t)blob & 3) == 0, "peeb_bitmap_encode: blob must be 4-byte aligned");
    PEEB_REQUIRE(blob_capacity >= (int64_t)peeb_bitmap_blob_bound(n), "peeb_bitmap_encode: blob_capacity below peeb_bitmap_blob_bound(n)");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    const PbrSizes sz = pbr_sizes(n);
    PEEB_REQUIRE(sz.n2 < (1ll << 31) && sz.n0 < (1ll << 32), "peeb_bitmap_encode: too many elements");
    const size_t o_l1 = align_up((size_t)sz.n0 * 4, 256), o_cnt = o_l1 + align_up((size_t)sz.n1 * 4, 256),
                 o_tot = o_cnt + align_up((size_t)sz.n2 * sizeof(int2), 256);
    int rc = scratch_reserve(ws->bits, o_tot + 256);
    if (rc) return rc;
    rc = scratch_reserve(ws->tables_h, 256, true);
    if (rc) return rc;
    char* d = (char*)ws->bits.ptr;
    unsigned* l0 = (unsigned*)d; unsigned* l1 = (unsigned*)(d + o_l1);
    int2* counts = (int2*)(d + o_cnt); unsigned* totals = (unsigned*)(d + o_tot);
    unsigned* totals_h = (unsigned*)ws->tables_h.ptr;
    unsigned* header = (unsigned*)blob;
    unsigned* body = header + PBR_HEADER / 4;
    {
        ProfScope p(ws, PEEB_K_BITMAP_ENCODE, st);
        if (sz.n2 > 0) {
            const unsigned grid = (unsigned)sz.n2;
            if (packed_input) {
                if (((uintptr_t)src & 3) == 0) pbr_pack_kernel<true, true><<<grid, PBR_CTA, 0, st>>>(src, n, sz.n0, sz.n1, l0, l1, body, counts);
                else pbr_pack_kernel<true, false><<<grid, PBR_CTA, 0, st>>>(src, n, sz.n0, sz.n1, l0, l1, body, counts);
            } else {
                if (((uintptr_t)src & 15) == 0) pbr_pack_kernel<false, true><<<grid, PBR_CTA, 0, st>>>(src, n, sz.n0, sz.n1, l0, l1, body, counts);
                else pbr_pack_kernel<false, false><<<grid, PBR_CTA, 0, st>>>(src, n, sz.n0, sz.n1, l0, l1, body, counts);
            }
        }
        pbr_scan_kernel<<<1, 1024, 0, st>>>(counts, sz.n2, n, header, totals);
        if (sz.n2 > 0)
            pbr_compact_kernel<<<(unsigned)sz.n2, PBR_CTA, 0, st>>>(l0, l1, sz.n0, sz.n1, sz.n2, counts, totals, body);
    }
    PEEB_CUDA(cudaGetLastError());
    PEEB_CUDA(cudaMemcpyAsync(totals_h, totals, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaStreamSynchronize(st));
    *blob_bytes = PBR_HEADER + 4 * (2 * sz.n2 + (int64_t)totals_h[0] + (int64_t)totals_h[1]);
    return PEEB_OK;
}

int peeb_bitmap_encode_h(peeb_ws* ws, const uint8_t* src_host, int64_t n, int packed_input, uint8_t* blob_host,
                         int64_t blob_capacity, int64_t* blob_bytes) {
    PEEB_REQUIRE(ws && blob_host && blob_bytes && (src_host || n == 0), "peeb_bitmap_encode_h: null pointer");
    PEEB_REQUIRE(n >= 0 && n < (1ll << 40), "peeb_bitmap_encode_h: bad element count");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const size_t in_bytes = packed_input ? (size_t)((n + 7) / 8) : (size_t)n;
    const size_t bound = peeb_bitmap_blob_bound(n);
    const size_t o_blob = align_up(in_bytes, 256);
    int rc = scratch_reserve(ws->stage, o_blob + align_up(bound, 256) + 256);
    if (rc) return rc;
    char* d = (char*)ws->stage.ptr;
    if (in_bytes) PEEB_CUDA(cudaMemcpyAsync(d, src_host, in_bytes, cudaMemcpyHostToDevice, ws->stream));
    int64_t got = 0;
    rc = peeb_bitmap_encode(ws, (const uint8_t*)d, n, packed_input, (uint8_t*)(d + o_blob), (int64_t)bound, &got, ws->stream);
    if (rc) { cudaStreamSynchronize(ws->stream); return rc; }
    PEEB_REQUIRE(blob_capacity >= got, "peeb_bitmap_encode_h: blob_capacity %lld below the %lld bytes of this map",
                 (long long)blob_capacity, (long long)got);
    PEEB_CUDA(cudaMemcpyAsync(blob_host, d + o_blob, (size_t)got, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    *blob_bytes = got;
    return PEEB_OK;
}

int peeb_bitmap_decode(peeb_ws* ws, const uint8_t* blob, int64_t blob_bytes, uint8_t* dst, int64_t n,
                       int packed_output, void* stream) {
    PEEB_REQUIRE(ws && blob && (dst || n == 0), "peeb_bitmap_decode: null pointer");
    PEEB_REQUIRE(n >= 0 && n < (1ll << 40), "peeb_bitmap_decode: bad element count");
    PEEB_REQUIRE(((uintptr_t)blob & 3) == 0, "peeb_bitmap_decode: blob must be 4-byte aligned");
    PEEB_REQUIRE(blob_bytes >= PBR_HEADER, "peeb_bitmap_decode: not a PBR1 blob");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char h[PBR_HEADER];
    PEEB_CUDA(cudaMemcpyAsync(h, blob, PBR_HEADER, cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaStreamSynchronize(st));
    PbrSizes sz; unsigned nz1, nz0;
    int rc = parse_header(h, blob_bytes, n, sz, nz1, nz0);
    if (rc) return rc;
    if (n == 0) return PEEB_OK;
    return decode_dev(ws, blob, sz, nz1, nz0, dst, packed_output, st);
}

int peeb_bitmap_decode_h(peeb_ws* ws, const uint8_t* blob_host, int64_t blob_bytes, uint8_t* dst_host, int64_t n,
                         int packed_output) {
    PEEB_REQUIRE(ws && blob_host && (dst_host || n == 0), "peeb_bitmap_decode_h: null pointer");
    PEEB_REQUIRE(n >= 0 && n < (1ll << 40), "peeb_bitmap_decode_h: bad element count");
    PEEB_CUDA(cudaSetDevice(ws->device));
    PbrSizes sz; unsigned nz1, nz0;
    int rc = parse_header(blob_host, blob_bytes, n, sz, nz1, nz0);
    if (rc) return rc;
    if (n == 0) return PEEB_OK;
    const size_t out_bytes = packed_output ? (size_t)((n + 7) / 8) : (size_t)n;
    const size_t o_out = align_up((size_t)blob_bytes, 256);
    rc = scratch_reserve(ws->stage, o_out + align_up(out_bytes, 256) + 256);
    if (rc) return rc;
    char* d = (char*)ws->stage.ptr;
    PEEB_CUDA(cudaMemcpyAsync(d, blob_host, (size_t)blob_bytes, cudaMemcpyHostToDevice, ws->stream));
    rc = decode_dev(ws, (const unsigned char*)d, sz, nz1, nz0, (unsigned char*)(d + o_out), packed_output, ws->stream);
    if (rc) { cudaStreamSynchronize(ws->stream); return rc; }
    PEEB_CUDA(cudaMemcpyAsync(dst_host, d + o_out, out_bytes, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

}  // extern "C"
