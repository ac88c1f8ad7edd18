// Rows a5-a9: the reference's bit-plane LSB path on sm_100a.
//   hist_planes   : image histogram (+ bit-plane population counts derived from it)
//   planes_unpack : bit-plane split          planes_pack : merge_modalities
//   tile_moments  : per-tile sum / sum of squares for the variance argmax
//   lsb_embed     : per-plane wrapped raster overwrite + XOR side bitmaps
//   compact_bits  : decode_message's "LSBs where the bitmap is set" gather
// All of it is streaming byte/integer work: 128-bit coalesced accesses, grids
// sized in multiples of the SM count, no tensor cores.
#include <algorithm>

#include "peeb_common.cuh"

namespace peeb {

static inline unsigned grid_for(peeb_ws* ws, long long work_items, int per_block, int waves = 8) {
    long long want = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)ws->sm_count * waves;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (unsigned)want;
}

// ------------------------------------------------------------------ histogram
// Shared-memory privatised histogram over values [0, WIN); larger values (only
// possible for 16-bit data above the window) go to global atomics.
template <int ITEM>
__global__ void __launch_bounds__(256) hist_kernel(const unsigned char* __restrict__ img, long long n, int win,
                                                   unsigned* __restrict__ hist) {
    extern __shared__ unsigned sh[];
    for (int k = threadIdx.x; k < win; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    constexpr int PER = 16 / ITEM;
    const bool aligned = (((uintptr_t)img) & 15) == 0;
    const long long nvec = aligned ? n / PER : 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    auto put = [&](unsigned v) {
        if ((int)v < win) atomicAdd(sh + v, 1u);
        else atomicAdd(hist + v, 1u);
    };
    const int4* q = reinterpret_cast<const int4*>(img);
    for (long long i = tid; i < nvec; i += nthr) {
        const int4 v = ldg_stream(q + i);
        const unsigned wv[4] = {(unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (ITEM == 2) { put(wv[k] & 0xffffu); put(wv[k] >> 16); }
            else { put(wv[k] & 0xffu); put((wv[k] >> 8) & 0xffu); put((wv[k] >> 16) & 0xffu); put(wv[k] >> 24); }
        }
    }
    for (long long e = nvec * PER + tid; e < n; e += nthr)
        put(ITEM == 2 ? (unsigned)reinterpret_cast<const unsigned short*>(img)[e] : (unsigned)img[e]);
    __syncthreads();
    for (int k = threadIdx.x; k < win; k += blockDim.x) {
        const unsigned c = sh[k];
        if (c) atomicAdd(hist + k, c);
    }
}

// 16-bit data: one CTA of 1024 threads per SM with 128 KB of shared counters, used in one of two ways that
// the CTA picks from a strided sample of its own part of the image (no host round trip, exact either way):
//   window  32-bit counters for the bins [0, 32768) -- 8..15-bit data in uint16 containers; a value above
//           the window (the sample missed it) goes to the global histogram with an atomic;
//   packed  all 65 536 bins as packed 16-bit counters.  The CTA counts at most 57 344 pixels (7 x 8 per
//           thread) between two flushes to the global histogram, so no 16-bit counter can wrap and no
//           carry can reach its neighbour.
__global__ void __launch_bounds__(1024) hist16_kernel(const unsigned char* __restrict__ img, long long n,
                                                      unsigned* __restrict__ hist, int force_window) {
    extern __shared__ unsigned sh[];
    constexpr int ROUNDS = 7;  // 7 * 1024 threads * 8 pixels < 65 536
    const unsigned short* px = reinterpret_cast<const unsigned short*>(img);
    const bool aligned = (((uintptr_t)img) & 15) == 0;
    const long long nvec = aligned ? n / 8 : 0;
    const int4* q = reinterpret_cast<const int4*>(img);
    const long long chunk = (long long)ROUNDS * blockDim.x;
    // sample: one pixel per thread out of each of the first chunks this CTA will read
    unsigned top = 0;
    if (!force_window) {
        const long long first = (long long)blockIdx.x * chunk * 8, step = (long long)gridDim.x * chunk * 8;
        for (int k = 0; k < 4; ++k) {
            const long long e = first + k * step + (long long)threadIdx.x * (ROUNDS * 8) + 5;
            if (e < n) top = max(top, (unsigned)px[e]);
        }
    }
    for (int k = threadIdx.x; k < 32768; k += blockDim.x) sh[k] = 0;
    const bool packed = __syncthreads_or(top >= 32768u) != 0;  // also orders the zeroing before the counting
    auto flush = [&]() {
        __syncthreads();
        for (int k = threadIdx.x; k < 32768; k += blockDim.x) {
            const unsigned c = sh[k];
            if (c) {
                if (packed) {
                    if (c & 0xffffu) atomicAdd(hist + 2 * k, c & 0xffffu);
                    if (c >> 16) atomicAdd(hist + 2 * k + 1, c >> 16);
                } else {
                    atomicAdd(hist + k, c);
                }
                sh[k] = 0;
            }
        }
        __syncthreads();
    };
    auto put = [&](unsigned v) {
        if (packed) atomicAdd(sh + (v >> 1), (v & 1u) ? 0x10000u : 1u);
        else if (v < 32768u) atomicAdd(sh + v, 1u);
        else atomicAdd(hist + v, 1u);
    };
    for (long long base = (long long)blockIdx.x * chunk; base < nvec; base += (long long)gridDim.x * chunk) {
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const long long i = base + (long long)r * blockDim.x + threadIdx.x;
            if (i < nvec) {
                const int4 v = ldg_stream(q + i);
                const unsigned wv[4] = {(unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) { put(wv[k] & 0xffffu); put(wv[k] >> 16); }
            }
        }
        if (packed) flush();  // 32-bit window counters cannot wrap (n < 2^32): they are flushed once, at the end
    }
    // what the vector loop left (unaligned image: everything), in slices of 57 344 pixels per CTA
    const long long rest0 = nvec * 8, slice = chunk * 8;
    for (long long base = rest0 + (long long)blockIdx.x * slice; base < n; base += (long long)gridDim.x * slice) {
        for (long long e = base + threadIdx.x; e < min(base + slice, n); e += blockDim.x) put((unsigned)px[e]);
        if (packed) flush();
    }
    if (!packed) flush();
}

// plane_ones[b] += sum over values v with bit b set of hist[v]; one CTA per 1024 bins (plane_ones zeroed by the caller;
// the first version walked all 65 536 bins with one CTA: 77 us, the longest launch of a small call)
__global__ void __launch_bounds__(256) hist_to_planes_kernel(const unsigned* __restrict__ hist, int nbins,
                                                             unsigned long long* __restrict__ plane_ones) {
    unsigned long long acc[16];
#pragma unroll
    for (int b = 0; b < 16; ++b) acc[b] = 0;
    const int v_end = min(nbins, ((int)blockIdx.x + 1) * 1024);
    for (int v = (int)blockIdx.x * 1024 + threadIdx.x; v < v_end; v += blockDim.x) {
        const unsigned long long c = hist[v];
#pragma unroll
        for (int b = 0; b < 16; ++b) acc[b] += ((v >> b) & 1) ? c : 0ull;
    }
    __shared__ unsigned long long sm[16][8];
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const unsigned long long t = warp_sum_u64(acc[b]);
        if ((threadIdx.x & 31) == 0) sm[b][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; ++k) t += sm[threadIdx.x][k];
        if (t) atomicAdd(plane_ones + threadIdx.x, t);
    }
}

static int launch_hist(peeb_ws* ws, const void* img, int64_t n, int itemsize, uint32_t* hist, uint64_t* ones,
                       cudaStream_t st) {
    const int nbins = itemsize == 1 ? 256 : 65536;
    PEEB_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * 65536, st));
    const int win = itemsize == 1 ? 256 : 16384;  // 64 KB of shared counters for 16-bit data
    const size_t smem = (size_t)win * sizeof(unsigned);
    {
        ProfScope p(ws, PEEB_K_HIST_PLANES, st);
        if (n > 0) {
            const unsigned grid = grid_for(ws, n, 256 * (16 / itemsize) * 8, 3);
            if (itemsize == 2 && (size_t)ws->max_smem_optin >= 131072) {
                PEEB_CUDA(cudaFuncSetAttribute(hist16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
                const long long per_cta = 7 * 1024 * 8;
                const unsigned g16 = (unsigned)std::max<long long>(1, std::min<long long>(ws->sm_count, (n + per_cta - 1) / per_cta));
                hist16_kernel<<<g16, 1024, 131072, st>>>((const unsigned char*)img, n, hist, getenv("PEEB_HIST_WINDOW_ONLY") ? 1 : 0);
            } else if (itemsize == 2) {
                PEEB_CUDA(cudaFuncSetAttribute(hist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                hist_kernel<2><<<grid, 256, smem, st>>>((const unsigned char*)img, n, win, hist);
            } else {
                hist_kernel<1><<<grid, 256, smem, st>>>((const unsigned char*)img, n, win, hist);
            }
        }
        PEEB_CUDA(cudaMemsetAsync(ones, 0, 16 * sizeof(uint64_t), st));
        hist_to_planes_kernel<<<(nbins + 1023) / 1024, 256, 0, st>>>(hist, nbins, (unsigned long long*)ones);
    }
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

// ------------------------------------------------------------------ plane split / merge
template <int ITEM>
__global__ void __launch_bounds__(256) planes_unpack_kernel(const unsigned char* __restrict__ img, long long n,
                                                            int first, int np, unsigned char* __restrict__ out) {
    constexpr int PER = 16 / ITEM;
    const unsigned lsb = ITEM == 2 ? 0x00010001u : 0x01010101u;
    const bool aligned = ((((uintptr_t)img) | (uintptr_t)out | (uintptr_t)(n * ITEM)) & 15) == 0;
    const long long nvec = aligned ? n / PER : 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < nvec; i += nthr) {
        const int4 v = ldg_stream(reinterpret_cast<const int4*>(img) + i);
        for (int k = 0; k < np; ++k) {
            const int sh = first + k;
            int4 o;
            if (sh >= 8 * ITEM) { o = make_int4(0, 0, 0, 0); }
            else {
                o.x = (int)(((unsigned)v.x >> sh) & lsb); o.y = (int)(((unsigned)v.y >> sh) & lsb);
                o.z = (int)(((unsigned)v.z >> sh) & lsb); o.w = (int)(((unsigned)v.w >> sh) & lsb);
            }
            stg_stream(reinterpret_cast<int4*>(out + (size_t)k * n * ITEM) + i, o);
        }
    }
    for (long long e = nvec * PER + tid; e < n; e += nthr) {
        const unsigned v = ITEM == 2 ? (unsigned)reinterpret_cast<const unsigned short*>(img)[e] : (unsigned)img[e];
        for (int k = 0; k < np; ++k) {
            const int sh = first + k;
            const unsigned b = sh >= 8 * ITEM ? 0u : (v >> sh) & 1u;
            if (ITEM == 2) reinterpret_cast<unsigned short*>(out + (size_t)k * n * 2)[e] = (unsigned short)b;
            else (out + (size_t)k * n)[e] = (unsigned char)b;
        }
    }
}

// out[i] = OR_k trunc_out(planes[k][i]) << k ; IN_ITEM / OUT_ITEM in {1,2}
template <int IN_ITEM, int OUT_ITEM>
__global__ void __launch_bounds__(256) planes_pack_kernel(const unsigned char* __restrict__ planes, long long n,
                                                          int np, unsigned char* __restrict__ out) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    constexpr unsigned OMASK = OUT_ITEM == 2 ? 0xffffu : 0xffu;
    // 8 pixels per thread and step: 16 (8) input bytes per plane, 16 (8) output bytes
    const bool aligned = ((((uintptr_t)planes) | (uintptr_t)out | (uintptr_t)(n * IN_ITEM)) & 15) == 0;
    const long long ngrp = aligned ? n / 8 : 0;
    for (long long gi = tid; gi < ngrp; gi += nthr) {
        unsigned acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0;
        for (int k = 0; k < np; ++k) {
            const unsigned char* pl = planes + (size_t)k * n * IN_ITEM;
            unsigned px[8];
            if (IN_ITEM == 2) {
                const int4 v = ldg_stream(reinterpret_cast<const int4*>(pl) + gi);
                const unsigned wv[4] = {(unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { px[2 * j] = wv[j] & 0xffffu; px[2 * j + 1] = wv[j] >> 16; }
            } else {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(pl) + gi);
#pragma unroll
                for (int j = 0; j < 4; ++j) { px[j] = (v.x >> (8 * j)) & 0xffu; px[4 + j] = (v.y >> (8 * j)) & 0xffu; }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] |= ((px[j] & OMASK) << k) & OMASK;
        }
        if (OUT_ITEM == 2) {
            int4 o;
            o.x = (int)(acc[0] | (acc[1] << 16)); o.y = (int)(acc[2] | (acc[3] << 16));
            o.z = (int)(acc[4] | (acc[5] << 16)); o.w = (int)(acc[6] | (acc[7] << 16));
            stg_stream(reinterpret_cast<int4*>(out) + gi, o);
        } else {
            uint2 o;
            o.x = acc[0] | (acc[1] << 8) | (acc[2] << 16) | (acc[3] << 24);
            o.y = acc[4] | (acc[5] << 8) | (acc[6] << 16) | (acc[7] << 24);
            reinterpret_cast<uint2*>(out)[gi] = o;
        }
    }
    for (long long e = ngrp * 8 + tid; e < n; e += nthr) {
        unsigned acc = 0;
        for (int k = 0; k < np; ++k) {
            const unsigned char* pl = planes + (size_t)k * n * IN_ITEM;
            const unsigned v = IN_ITEM == 2 ? (unsigned)reinterpret_cast<const unsigned short*>(pl)[e] : (unsigned)pl[e];
            acc |= ((v & OMASK) << k) & OMASK;
        }
        if (OUT_ITEM == 2) reinterpret_cast<unsigned short*>(out)[e] = (unsigned short)acc;
        else out[e] = (unsigned char)acc;
    }
}

// ------------------------------------------------------------------ tile moments
// One warp per tile: lanes stride over the tile's pixels in raster order.
template <int ITEM>
__global__ void __launch_bounds__(256) tile_moments_kernel(const unsigned char* __restrict__ plane, int h, int w,
                                                           int sbs, int tiles_x, long long ntiles,
                                                           long long* __restrict__ out) {
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    const bool fast16 = sbs == 16 && (w & 7) == 0 && (((uintptr_t)plane) & 15) == 0;
    for (long long t = warp0; t < ntiles; t += nwarps) {
        const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
        const int y0 = ty * sbs, x0 = tx * sbs;
        const int th = min(sbs, h - y0), tw = min(sbs, w - x0);
        unsigned long long s1 = 0, s2 = 0;
        const long long npx = (long long)th * tw;
        if (ITEM == 2 && fast16 && th == 16 && tw == 16) {
            // full 16x16 tile of 16-bit pixels: one 16-byte load per lane (row = lane / 2, half a row each)
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(plane) +
                                                                  (long long)(y0 + (lane >> 1)) * w + x0 + 8 * (lane & 1)));
            const unsigned wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned lo = wv[k] & 0xffffu, hi = wv[k] >> 16;
                s1 += lo + hi;
                s2 += (unsigned long long)lo * lo + (unsigned long long)hi * hi;
            }
        } else
        for (long long k = lane; k < npx; k += 32) {
            const int yy = (int)(k / tw), xx = (int)(k % tw);
            const long long at = (long long)(y0 + yy) * w + x0 + xx;
            const unsigned v = ITEM == 2 ? (unsigned)reinterpret_cast<const unsigned short*>(plane)[at] : (unsigned)plane[at];
            s1 += v;
            s2 += (unsigned long long)v * v;
        }
        s1 = warp_sum_u64(s1);
        s2 = warp_sum_u64(s2);
        if (lane == 0) { out[2 * t] = (long long)s1; out[2 * t + 1] = (long long)s2; }
    }
}

// ------------------------------------------------------------------ LSB embed
struct LsbSeg { long long start, len, bit_off; };
struct LsbSegs { LsbSeg s[16]; };

template <int ITEM>
__global__ void __launch_bounds__(256) lsb_embed_kernel(const unsigned char* __restrict__ in, long long n, LsbSegs segs,
                                                        const unsigned char* __restrict__ payload,
                                                        unsigned char* __restrict__ out,
                                                        unsigned char* __restrict__ bitmaps) {
    const int p = blockIdx.y;
    const LsbSeg sg = segs.s[p];
    const unsigned char* src = in + (size_t)p * n * ITEM;
    unsigned char* dst = out + (size_t)p * n * ITEM;
    unsigned char* bm = bitmaps + (size_t)p * n;
    constexpr int PER = 8;  // pixels per thread and step
    const bool aligned = ((((uintptr_t)in) | (uintptr_t)out | (uintptr_t)bitmaps | (uintptr_t)(n * ITEM) | (uintptr_t)n) & 15) == 0;
    const long long ngrp = aligned ? n / PER : 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    auto one = [&](long long i, unsigned v, unsigned& nv, unsigned& x) {
        long long rel = i - sg.start;
        if (rel < 0) rel += n;
        if (rel < sg.len) {
            const long long bp = sg.bit_off + rel;
            const unsigned bit = (__ldg(payload + (bp >> 3)) >> (7 - (int)(bp & 7))) & 1u;
            nv = (v & 0xFEu) | bit;                 // src/codec.py:306 / :472 on the plane's own dtype
            x = (v ^ nv) & 0xffu;                   // uint8 side bitmap, :310-311
        } else { nv = v; x = 0; }
    };
    for (long long gi = tid; gi < ngrp; gi += nthr) {
        unsigned px[8], nv[8], xr[8];
        if (ITEM == 2) {
            const int4 v = ldg_stream(reinterpret_cast<const int4*>(src) + gi);
            const unsigned wv[4] = {(unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { px[2 * j] = wv[j] & 0xffffu; px[2 * j + 1] = wv[j] >> 16; }
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(src) + gi);
#pragma unroll
            for (int j = 0; j < 4; ++j) { px[j] = (v.x >> (8 * j)) & 0xffu; px[4 + j] = (v.y >> (8 * j)) & 0xffu; }
        }
        // a group that lies entirely inside the plane's segment (or entirely outside) needs one position
        // test and one 8-bit payload window; groups on a segment boundary take the per-pixel code
        long long rel0 = gi * 8 - sg.start;
        if (rel0 < 0) rel0 += n;
        if (rel0 + 7 < sg.len) {
            const long long bp = sg.bit_off + rel0;
            const int sh = (int)(bp & 7);
            unsigned two = (unsigned)__ldg(payload + (bp >> 3)) << 8;
            if (sh) two |= (unsigned)__ldg(payload + (bp >> 3) + 1);
            const unsigned bits8 = (two >> (8 - sh)) & 0xffu;  // bit of pixel j at position 7 - j
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned bit = (bits8 >> (7 - j)) & 1u;
                nv[j] = (px[j] & 0xFEu) | bit;
                xr[j] = (px[j] ^ nv[j]) & 0xffu;
            }
        } else if (rel0 >= sg.len && rel0 + 7 < n) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { nv[j] = px[j]; xr[j] = 0; }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) one(gi * 8 + j, px[j], nv[j], xr[j]);
        }
        if (ITEM == 2) {
            int4 o;
            o.x = (int)(nv[0] | (nv[1] << 16)); o.y = (int)(nv[2] | (nv[3] << 16));
            o.z = (int)(nv[4] | (nv[5] << 16)); o.w = (int)(nv[6] | (nv[7] << 16));
            stg_stream(reinterpret_cast<int4*>(dst) + gi, o);
        } else {
            uint2 o;
            o.x = nv[0] | (nv[1] << 8) | (nv[2] << 16) | (nv[3] << 24);
            o.y = nv[4] | (nv[5] << 8) | (nv[6] << 16) | (nv[7] << 24);
            reinterpret_cast<uint2*>(dst)[gi] = o;
        }
        uint2 b;
        b.x = xr[0] | (xr[1] << 8) | (xr[2] << 16) | (xr[3] << 24);
        b.y = xr[4] | (xr[5] << 8) | (xr[6] << 16) | (xr[7] << 24);
        reinterpret_cast<uint2*>(bm)[gi] = b;
    }
    for (long long e = ngrp * PER + tid; e < n; e += nthr) {
        const unsigned v = ITEM == 2 ? (unsigned)reinterpret_cast<const unsigned short*>(src)[e] : (unsigned)src[e];
        unsigned nv, x;
        one(e, v, nv, x);
        if (ITEM == 2) reinterpret_cast<unsigned short*>(dst)[e] = (unsigned short)nv;
        else dst[e] = (unsigned char)nv;
        bm[e] = (unsigned char)x;
    }
}

// ------------------------------------------------------------------ compaction (decode_message)
constexpr int CB_BLOCK = 256, CB_PER = 8, CB_TILE = CB_BLOCK * CB_PER;  // 2048 positions per block

__global__ void __launch_bounds__(CB_BLOCK) compact_count_kernel(const unsigned char* __restrict__ bitmap, long long n,
                                                                 int* __restrict__ block_cnt) {
    const long long base = (long long)blockIdx.x * CB_TILE;
    int c = 0;
    for (int j = 0; j < CB_PER; ++j) {
        const long long i = base + (long long)j * CB_BLOCK + threadIdx.x;
        c += (i < n && bitmap[i] != 0) ? 1 : 0;
    }
    c = (int)warp_sum_i64(c);
    __shared__ int sm[CB_BLOCK / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < CB_BLOCK / 32; ++k) t += sm[k];
        block_cnt[blockIdx.x] = t;
    }
}

// exclusive scan of block counts in place (64-bit offsets out), one block
__global__ void __launch_bounds__(1024) compact_scan_kernel(const int* __restrict__ block_cnt, int nblocks,
                                                            long long* __restrict__ block_off, long long limit,
                                                            long long* __restrict__ count_out) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += blockDim.x) {
        const int idx = base + threadIdx.x;
        const long long v = idx < nblocks ? block_cnt[idx] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long ws = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_tot[lane] = ws;
        }
        __syncthreads();
        const long long before = carry_s + (warp > 0 ? warp_tot[warp - 1] : 0);
        if (idx < nblocks) block_off[idx] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = carry_s < limit ? carry_s : limit;
}

template <int ITEM>
__global__ void __launch_bounds__(CB_BLOCK) compact_write_kernel(const unsigned char* __restrict__ plane,
                                                                 const unsigned char* __restrict__ bitmap, long long n,
                                                                 const long long* __restrict__ block_off,
                                                                 long long limit, unsigned* __restrict__ bits_out) {
    const long long base = (long long)blockIdx.x * CB_TILE;
    long long running = block_off[blockIdx.x];
    if (running >= limit) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    __shared__ int wcnt[CB_BLOCK / 32];
    for (int j = 0; j < CB_PER; ++j) {
        const long long i = base + (long long)j * CB_BLOCK + threadIdx.x;
        const bool f = i < n && bitmap[i] != 0;
        unsigned bit = 0;
        if (f) bit = (ITEM == 2 ? (unsigned)reinterpret_cast<const unsigned short*>(plane)[i] : (unsigned)plane[i]) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wcnt[warp] = __popc(m);
        __syncthreads();
        long long wbase = running;
        int total = 0;
        for (int k = 0; k < CB_BLOCK / 32; ++k) { if (k < warp) wbase += wcnt[k]; total += wcnt[k]; }
        // this warp's set positions occupy output bits [wbase, wbase + popc(m)): at most two words
        const long long gpos = wbase + __popc(m & lt);
        const bool live = f && gpos < limit;
        const long long w0 = wbase >> 5;
        unsigned c0 = 0, c1 = 0;
        if (live && bit) {
            const unsigned mask = 1u << (8 * ((int)(gpos >> 3) & 3) + 7 - (int)(gpos & 7));  // MSB-first bytes
            if ((gpos >> 5) == w0) c0 = mask; else c1 = mask;
        }
        c0 = __reduce_or_sync(0xffffffffu, c0);
        c1 = __reduce_or_sync(0xffffffffu, c1);
        if (lane == 0) {
            if (c0) atomicOr(bits_out + w0, c0);
            if (c1) atomicOr(bits_out + w0 + 1, c1);
        }
        running += total;
        __syncthreads();
    }
}

// ---- N4: the true inverse of the LSB path ---------------------------------------
// cover = stego ^ sum_p (bitmap_p << p): src/codec.py keeps bitmap_p = orig_plane ^ stego_plane (0/1),
// SURVEY F3.3.  8 pixels per thread, 128-bit image accesses, 64-bit bitmap loads.
template <int ITEM>
__global__ void __launch_bounds__(256) lsb_recover_kernel(const unsigned char* __restrict__ stego,
                                                          const unsigned char* __restrict__ bitmaps, long long n, int s,
                                                          unsigned char* __restrict__ cover) {
    const bool aligned = ((((uintptr_t)stego) | (uintptr_t)cover | (uintptr_t)bitmaps | (uintptr_t)n) & 15) == 0;
    const long long ngrp = aligned ? n / 8 : 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (long long gi = tid; gi < ngrp; gi += nthr) {
        unsigned m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int p = 0; p < s; ++p) {
            const uint2 b = __ldg(reinterpret_cast<const uint2*>(bitmaps + (size_t)p * n) + gi);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                m[j] |= ((b.x >> (8 * j)) & 1u) << p;
                m[4 + j] |= ((b.y >> (8 * j)) & 1u) << p;
            }
        }
        if (ITEM == 2) {
            int4 v = ldg_stream(reinterpret_cast<const int4*>(stego) + gi);
            v.x ^= (int)(m[0] | (m[1] << 16)); v.y ^= (int)(m[2] | (m[3] << 16));
            v.z ^= (int)(m[4] | (m[5] << 16)); v.w ^= (int)(m[6] | (m[7] << 16));
            stg_stream(reinterpret_cast<int4*>(cover) + gi, v);
        } else {
            uint2 v = __ldg(reinterpret_cast<const uint2*>(stego) + gi);
            v.x ^= m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
            v.y ^= m[4] | (m[5] << 8) | (m[6] << 16) | (m[7] << 24);
            reinterpret_cast<uint2*>(cover)[gi] = v;
        }
    }
    for (long long i = ngrp * 8 + tid; i < n; i += nthr) {
        unsigned mk = 0;
        for (int p = 0; p < s; ++p) mk |= (unsigned)(bitmaps[(size_t)p * n + i] & 1u) << p;
        if (ITEM == 2) reinterpret_cast<unsigned short*>(cover)[i] = (unsigned short)(reinterpret_cast<const unsigned short*>(stego)[i] ^ mk);
        else cover[i] = (unsigned char)(stego[i] ^ mk);
    }
}

// bits_out bit (bit_off[p] + k) = bit p of stego[(start[p] + k) mod n], 0 <= k < len[p]; MSB-first bytes.
//
// A warp item = 1024 consecutive bits of one segment = 32 output chunks of 32 bits, one per lane; the items of all
// planes form one list.  Fast form (the item's pixels do not wrap and the image is 16-byte aligned): the pixels are
// read in ALIGNED blocks of eight (one 16- / 8-byte load per lane, four rounds cover 1024 pixels), a block's
// plane-p bits are gathered into a byte with one multiply per word, four lanes join their bytes into a word of the
// aligned bit stream, and the segment's own alignment (start mod 8) is dealt with afterwards, in bits: chunk =
// funnel shift of two neighbouring words.  About 150 instructions per item.  Generic form (wrap-around, the tail of
// the image, unaligned images): every lane reads the 32 pixels of round i, a ballot makes chunk i's word -- about
// 1000 instructions per item, which is what bound this kernel when it was the only form (0.16 of the HBM peak).
// History: one thread per chunk walking its own 32 pixels, a grid row per plane: 43 us for 18 Mbit; ballots: 37;
// loads before ballots: 33; one list of items: 27.
template <int ITEM>
__device__ __forceinline__ unsigned plane_bits8(const unsigned char* __restrict__ stego, long long pos, int p) {
    // plane-p bits of pixels pos .. pos + 7 (pos a multiple of 8), the first pixel in bit 7
    if (ITEM == 2) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(stego + 2 * pos));
        constexpr unsigned K = (1u << 17) | 1u;  // bits 0 and 16 of a masked word -> bits 17 and 16 of the product
        const unsigned a = ((((v.x >> p) & 0x00010001u) * K) >> 16) & 3u, b = ((((v.y >> p) & 0x00010001u) * K) >> 16) & 3u;
        const unsigned c = ((((v.z >> p) & 0x00010001u) * K) >> 16) & 3u, d = ((((v.w >> p) & 0x00010001u) * K) >> 16) & 3u;
        return (a << 6) | (b << 4) | (c << 2) | d;
    } else {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(stego + pos));
        constexpr unsigned K = (1u << 27) | (1u << 18) | (1u << 9) | 1u;  // bits 0, 8, 16, 24 -> bits 27 .. 24, no carries
        const unsigned a = ((((v.x >> p) & 0x01010101u) * K) >> 24) & 15u, b = ((((v.y >> p) & 0x01010101u) * K) >> 24) & 15u;
        return (a << 4) | b;
    }
}
// the bytes of four neighbouring lanes (lane & 3 = 0 first) as one word, in all four lanes
__device__ __forceinline__ unsigned quad_word(unsigned byte, int lane) {
    unsigned t = byte << (8 * (3 - (lane & 3)));
    t |= __shfl_xor_sync(0xffffffffu, t, 1);
    t |= __shfl_xor_sync(0xffffffffu, t, 2);
    return t;
}
template <int ITEM>
__global__ void __launch_bounds__(256) lsb_extract_kernel(const unsigned char* __restrict__ stego, long long n, int s,
                                                          LsbSegs segs, unsigned* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const bool base_aligned = ((uintptr_t)stego & 15) == 0;
    long long total = 0;
    for (int q = 0; q < s; ++q) total += (segs.s[q].len + 1023) >> 10;
    for (long long item = warp; item < total; item += nwarps) {
        int p = 0;
        long long grp = item;
        for (; p < s - 1; ++p) {
            const long long ng = (segs.s[p].len + 1023) >> 10;
            if (grp < ng) break;
            grp -= ng;
        }
        const LsbSeg sg = segs.s[p];
        const long long nchunks = (sg.len + 31) >> 5;
        const long long kbase = grp << 10;
        const long long ch = (grp << 5) + lane;  // this lane's chunk
        unsigned v = 0;                          // its word, first bit on top
        const long long a0 = (sg.start & ~7ll) + kbase;  // aligned pixel the item's bit stream starts at
        const int r = (int)(sg.start & 7);               // ... and the segment's offset inside it, in bits
        if (base_aligned && a0 + 1024 + 32 <= n) {
            unsigned w[4];  // w[j]: word 8 j + (lane >> 2) of the aligned stream
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = quad_word(plane_bits8<ITEM>(stego, a0 + 256 * j + 8 * lane, p), lane);
            // one more word for the last chunk when the segment does not start on a block
            const unsigned e = quad_word((r != 0 && lane < 4) ? plane_bits8<ITEM>(stego, a0 + 1024 + 8 * lane, p) : 0u, lane);
            unsigned cur = 0, nxt = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned x = __shfl_sync(0xffffffffu, w[j], 4 * (lane & 7));
                const unsigned y = __shfl_sync(0xffffffffu, w[j], 4 * ((lane + 1) & 7));
                if ((lane >> 3) == j) cur = x;
                if (((lane + 1) >> 3) == j) nxt = y;
            }
            const unsigned e0 = __shfl_sync(0xffffffffu, e, 0);
            if (lane == 31) nxt = e0;
            v = r ? __funnelshift_l(nxt, cur, r) : cur;
            const long long rem = sg.len - (ch << 5);  // bits of the chunk inside the segment
            if (rem <= 0) v = 0u;
            else if (rem < 32) v &= ~(0xffffffffu >> (int)rem);
        } else {
            unsigned px[32];  // (all loads before the first ballot: 32 memory round trips in a row otherwise)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const long long k = kbase + 32 * i + lane;
                long long pos = sg.start + k;  // start < n and k < len <= n: one wrap at most
                if (pos >= n) pos -= n;
                px[i] = 0u;
                if (k < sg.len) px[i] = ITEM == 2 ? __ldg(reinterpret_cast<const unsigned short*>(stego) + pos) : __ldg(stego + pos);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const unsigned wd = __brev(__ballot_sync(0xffffffffu, ((px[i] >> p) & 1u) != 0u));
                if (lane == i) v = wd;
            }
        }
        if (ch < nchunks && v) {
            const long long b = sg.bit_off + (ch << 5);  // stream bit of the chunk's first bit
            const int sh = (int)(b & 31);
            const unsigned hi = v >> sh, lo = sh ? v << (32 - sh) : 0u;
            if (hi) atomicOr(out + (b >> 5), __byte_perm(hi, 0, 0x0123));
            if (lo) atomicOr(out + (b >> 5) + 1, __byte_perm(lo, 0, 0x0123));
        }
    }
}

}  // namespace peeb

using namespace peeb;

extern "C" {

int peeb_hist_planes(peeb_ws* ws, const void* img, int64_t n, int itemsize, uint32_t* hist, uint64_t* plane_ones,
                     void* stream) {
    PEEB_REQUIRE(ws && img && hist && plane_ones, "peeb_hist_planes: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_hist_planes: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0, "peeb_hist_planes: negative size");
    PEEB_CUDA(cudaSetDevice(ws->device));
    return launch_hist(ws, img, n, itemsize, hist, plane_ones, (cudaStream_t)stream);
}

int peeb_hist_planes_h(peeb_ws* ws, const void* img_host, int64_t n, int itemsize, uint32_t* hist_host,
                       uint64_t* plane_ones_host) {
    PEEB_REQUIRE(ws && img_host && hist_host && plane_ones_host, "peeb_hist_planes_h: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_hist_planes_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0, "peeb_hist_planes_h: negative size");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const size_t bytes = align_up((size_t)n * itemsize, 256);
    int rc = scratch_reserve(ws->stage, bytes + 256); if (rc) return rc;
    rc = scratch_reserve(ws->tables, 65536 * 4 + 256); if (rc) return rc;
    uint32_t* dh = (uint32_t*)ws->tables.ptr;
    uint64_t* dp = (uint64_t*)((char*)ws->tables.ptr + 65536 * 4);
    PEEB_CUDA(cudaMemcpyAsync(ws->stage.ptr, img_host, (size_t)n * itemsize, cudaMemcpyHostToDevice, ws->stream));
    rc = launch_hist(ws, ws->stage.ptr, n, itemsize, dh, dp, ws->stream); if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(hist_host, dh, 65536 * 4, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaMemcpyAsync(plane_ones_host, dp, 16 * 8, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

int peeb_planes_unpack(peeb_ws* ws, const void* img, int64_t n, int itemsize, int first_plane, int n_planes,
                       void* planes_out, void* stream) {
    PEEB_REQUIRE(ws && img && planes_out, "peeb_planes_unpack: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_planes_unpack: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && first_plane >= 0 && n_planes >= 0 && n_planes <= 64, "peeb_planes_unpack: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (n == 0 || n_planes == 0) return PEEB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope p(ws, PEEB_K_PLANES_UNPACK, st);
    const unsigned grid = grid_for(ws, n, 256 * (16 / itemsize) * 2);
    if (itemsize == 2) planes_unpack_kernel<2><<<grid, 256, 0, st>>>((const unsigned char*)img, n, first_plane, n_planes, (unsigned char*)planes_out);
    else planes_unpack_kernel<1><<<grid, 256, 0, st>>>((const unsigned char*)img, n, first_plane, n_planes, (unsigned char*)planes_out);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int peeb_planes_unpack_h(peeb_ws* ws, const void* img_host, int64_t n, int itemsize, int first_plane, int n_planes,
                         void* planes_out_host) {
    PEEB_REQUIRE(ws && img_host && (planes_out_host || n_planes == 0), "peeb_planes_unpack_h: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_planes_unpack_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && n_planes >= 0 && n_planes <= 64, "peeb_planes_unpack_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (n == 0 || n_planes == 0) return PEEB_OK;
    const size_t ib = align_up((size_t)n * itemsize, 256);
    int rc = scratch_reserve(ws->stage, ib); if (rc) return rc;
    rc = scratch_reserve(ws->stage2, (size_t)n * itemsize * n_planes + 256); if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(ws->stage.ptr, img_host, (size_t)n * itemsize, cudaMemcpyHostToDevice, ws->stream));
    rc = peeb_planes_unpack(ws, ws->stage.ptr, n, itemsize, first_plane, n_planes, ws->stage2.ptr, ws->stream);
    if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(planes_out_host, ws->stage2.ptr, (size_t)n * itemsize * n_planes, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

int peeb_planes_pack(peeb_ws* ws, const void* planes, int64_t n, int in_itemsize, int n_planes, void* img_out,
                     void* stream) {
    PEEB_REQUIRE(ws && planes && img_out, "peeb_planes_pack: null pointer");
    PEEB_REQUIRE(in_itemsize == 1 || in_itemsize == 2, "peeb_planes_pack: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && n_planes >= 1 && n_planes <= 16, "peeb_planes_pack: n_planes must be 1..16");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (n == 0) return PEEB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int out_item = n_planes > 8 ? 2 : 1;  // src/codec.py:221
    ProfScope p(ws, PEEB_K_PLANES_PACK, st);
    const unsigned grid = grid_for(ws, n, 256 * 8 * 2);
    const unsigned char* pl = (const unsigned char*)planes;
    unsigned char* o = (unsigned char*)img_out;
    if (in_itemsize == 2 && out_item == 2) planes_pack_kernel<2, 2><<<grid, 256, 0, st>>>(pl, n, n_planes, o);
    else if (in_itemsize == 2) planes_pack_kernel<2, 1><<<grid, 256, 0, st>>>(pl, n, n_planes, o);
    else if (out_item == 2) planes_pack_kernel<1, 2><<<grid, 256, 0, st>>>(pl, n, n_planes, o);
    else planes_pack_kernel<1, 1><<<grid, 256, 0, st>>>(pl, n, n_planes, o);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int peeb_planes_pack_h(peeb_ws* ws, const void* const* plane_ptrs_host, int64_t n, int in_itemsize, int n_planes,
                       void* img_out_host) {
    PEEB_REQUIRE(ws && plane_ptrs_host && img_out_host, "peeb_planes_pack_h: null pointer");
    PEEB_REQUIRE(in_itemsize == 1 || in_itemsize == 2, "peeb_planes_pack_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && n_planes >= 1 && n_planes <= 16, "peeb_planes_pack_h: n_planes must be 1..16");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (n == 0) return PEEB_OK;
    const size_t pb = align_up((size_t)n * in_itemsize, 256);  // keeps every staged plane 16-byte aligned
    const int out_item = n_planes > 8 ? 2 : 1;
    int rc = scratch_reserve(ws->stage, pb * n_planes); if (rc) return rc;
    rc = scratch_reserve(ws->stage2, (size_t)n * out_item + 256); if (rc) return rc;
    // planes are staged back to back at stride n*itemsize (the kernel's layout), so copy them to a packed area
    char* d = (char*)ws->stage.ptr;
    for (int k = 0; k < n_planes; ++k) {
        PEEB_REQUIRE(plane_ptrs_host[k] != nullptr, "peeb_planes_pack_h: plane %d is null", k);
        PEEB_CUDA(cudaMemcpyAsync(d + (size_t)k * n * in_itemsize, plane_ptrs_host[k], (size_t)n * in_itemsize,
                                  cudaMemcpyHostToDevice, ws->stream));
    }
    rc = peeb_planes_pack(ws, d, n, in_itemsize, n_planes, ws->stage2.ptr, ws->stream); if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(img_out_host, ws->stage2.ptr, (size_t)n * out_item, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

int peeb_tile_moments(peeb_ws* ws, const void* plane, int h, int w, int itemsize, int sbs, int64_t* sums_out,
                      void* stream) {
    PEEB_REQUIRE(ws && plane && sums_out, "peeb_tile_moments: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_tile_moments: itemsize must be 1 or 2");
    PEEB_REQUIRE(h >= 1 && w >= 1 && sbs >= 1, "peeb_tile_moments: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles_x = (w + sbs - 1) / sbs, tiles_y = (h + sbs - 1) / sbs;
    const long long ntiles = (long long)tiles_x * tiles_y;
    ProfScope p(ws, PEEB_K_TILE_MOMENTS, st);
    const unsigned grid = grid_for(ws, ntiles, 8);
    if (itemsize == 2) tile_moments_kernel<2><<<grid, 256, 0, st>>>((const unsigned char*)plane, h, w, sbs, tiles_x, ntiles, (long long*)sums_out);
    else tile_moments_kernel<1><<<grid, 256, 0, st>>>((const unsigned char*)plane, h, w, sbs, tiles_x, ntiles, (long long*)sums_out);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int peeb_tile_moments_h(peeb_ws* ws, const void* plane_host, int h, int w, int itemsize, int sbs,
                        int64_t* sums_out_host) {
    PEEB_REQUIRE(ws && plane_host && sums_out_host, "peeb_tile_moments_h: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_tile_moments_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(h >= 1 && w >= 1 && sbs >= 1, "peeb_tile_moments_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const size_t ib = (size_t)h * w * itemsize;
    const long long ntiles = (long long)((w + sbs - 1) / sbs) * ((h + sbs - 1) / sbs);
    int rc = scratch_reserve(ws->stage, ib + 256); if (rc) return rc;
    rc = scratch_reserve(ws->stage2, (size_t)ntiles * 16 + 256); if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(ws->stage.ptr, plane_host, ib, cudaMemcpyHostToDevice, ws->stream));
    rc = peeb_tile_moments(ws, ws->stage.ptr, h, w, itemsize, sbs, (int64_t*)ws->stage2.ptr, ws->stream); if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(sums_out_host, ws->stage2.ptr, (size_t)ntiles * 16, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

static int check_segs(int64_t n, int s, const int64_t* start, const int64_t* len, const int64_t* bit_off,
                      int64_t payload_bits, LsbSegs& segs) {
    PEEB_REQUIRE(s >= 1 && s <= 16, "peeb_lsb_embed: s must be 1..16");
    for (int p = 0; p < s; ++p) {
        PEEB_REQUIRE(len[p] >= 0 && len[p] <= n, "peeb_lsb_embed: len[%d] out of range", p);
        PEEB_REQUIRE(len[p] == 0 || (start[p] >= 0 && start[p] < n), "peeb_lsb_embed: start[%d] out of range", p);
        PEEB_REQUIRE(len[p] == 0 || (bit_off[p] >= 0 && bit_off[p] + len[p] <= payload_bits),
                     "peeb_lsb_embed: segment %d reads past the payload", p);
        segs.s[p].start = len[p] ? start[p] : 0;
        segs.s[p].len = len[p];
        segs.s[p].bit_off = bit_off[p];
    }
    return PEEB_OK;
}

int peeb_lsb_embed(peeb_ws* ws, const void* planes_in, int64_t n, int itemsize, int s, const int64_t* start,
                   const int64_t* len, const int64_t* bit_off, const uint8_t* payload, int64_t payload_bits,
                   void* planes_out, uint8_t* bitmaps, void* stream) {
    PEEB_REQUIRE(ws && planes_in && start && len && bit_off && planes_out && bitmaps, "peeb_lsb_embed: null pointer");
    PEEB_REQUIRE(payload || payload_bits == 0, "peeb_lsb_embed: null payload");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_lsb_embed: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0, "peeb_lsb_embed: negative size");
    LsbSegs segs;
    int rc = check_segs(n, s, start, len, bit_off, payload_bits, segs); if (rc) return rc;
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (n == 0) return PEEB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope p(ws, PEEB_K_LSB_EMBED, st);
    long long bx = ((long long)ws->sm_count * 8 + s - 1) / s;
    const long long need = (n + 256 * 8 - 1) / (256 * 8);
    if (bx > need) bx = need;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)s);
    if (itemsize == 2) lsb_embed_kernel<2><<<grid, 256, 0, st>>>((const unsigned char*)planes_in, n, segs, payload, (unsigned char*)planes_out, bitmaps);
    else lsb_embed_kernel<1><<<grid, 256, 0, st>>>((const unsigned char*)planes_in, n, segs, payload, (unsigned char*)planes_out, bitmaps);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int peeb_lsb_embed_h(peeb_ws* ws, const void* const* plane_ptrs_host, int64_t n, int itemsize, int s,
                     const int64_t* start, const int64_t* len, const int64_t* bit_off, const uint8_t* payload_host,
                     int64_t payload_bits, void* planes_out_host, uint8_t* bitmaps_out_host) {
    PEEB_REQUIRE(ws && plane_ptrs_host && planes_out_host && bitmaps_out_host, "peeb_lsb_embed_h: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_lsb_embed_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && s >= 1 && s <= 16, "peeb_lsb_embed_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (n == 0) return PEEB_OK;
    const size_t pbytes = (size_t)n * itemsize * s, bbytes = (size_t)n * s, paybytes = (size_t)(payload_bits + 7) / 8;
    const size_t o_in = 0, o_out = align_up(pbytes, 256), o_bm = o_out + align_up(pbytes, 256);
    int rc = scratch_reserve(ws->stage, o_bm + align_up(bbytes, 256)); if (rc) return rc;
    rc = scratch_reserve(ws->stage2, paybytes + 256); if (rc) return rc;
    char* d = (char*)ws->stage.ptr;
    for (int k = 0; k < s; ++k) {
        PEEB_REQUIRE(plane_ptrs_host[k] != nullptr, "peeb_lsb_embed_h: plane %d is null", k);
        PEEB_CUDA(cudaMemcpyAsync(d + o_in + (size_t)k * n * itemsize, plane_ptrs_host[k], (size_t)n * itemsize,
                                  cudaMemcpyHostToDevice, ws->stream));
    }
    if (paybytes) PEEB_CUDA(cudaMemcpyAsync(ws->stage2.ptr, payload_host, paybytes, cudaMemcpyHostToDevice, ws->stream));
    rc = peeb_lsb_embed(ws, d + o_in, n, itemsize, s, start, len, bit_off, (const uint8_t*)ws->stage2.ptr, payload_bits,
                        d + o_out, (uint8_t*)(d + o_bm), ws->stream);
    if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(planes_out_host, d + o_out, pbytes, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaMemcpyAsync(bitmaps_out_host, d + o_bm, bbytes, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

int peeb_lsb_recover(peeb_ws* ws, const void* stego, const uint8_t* bitmaps, int64_t n, int itemsize, int s,
                     void* cover_out, void* stream) {
    PEEB_REQUIRE(ws && stego && bitmaps && cover_out, "peeb_lsb_recover: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_lsb_recover: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && s >= 1 && s <= 8 * itemsize, "peeb_lsb_recover: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (n == 0) return PEEB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long bx = (n / 8 + 255) / 256;
    const long long cap = (long long)ws->sm_count * 16;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    ProfScope prof(ws, PEEB_K_LSB_RECOVER, st);
    if (itemsize == 2) lsb_recover_kernel<2><<<(unsigned)bx, 256, 0, st>>>((const unsigned char*)stego, bitmaps, n, s, (unsigned char*)cover_out);
    else lsb_recover_kernel<1><<<(unsigned)bx, 256, 0, st>>>((const unsigned char*)stego, bitmaps, n, s, (unsigned char*)cover_out);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int peeb_lsb_recover_h(peeb_ws* ws, const void* stego_host, const uint8_t* const* bitmap_ptrs_host, int64_t n,
                       int itemsize, int s, void* cover_out_host) {
    PEEB_REQUIRE(ws && stego_host && bitmap_ptrs_host && cover_out_host, "peeb_lsb_recover_h: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_lsb_recover_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && s >= 1 && s <= 8 * itemsize, "peeb_lsb_recover_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (n == 0) return PEEB_OK;
    const size_t ib = (size_t)n * itemsize, o_out = align_up(ib, 256), o_bm = o_out + align_up(ib, 256);
    const size_t bpitch = align_up((size_t)n, 16);  // keeps every bitmap 16-byte aligned when n is
    int rc = scratch_reserve(ws->stage, o_bm + bpitch * s + 256); if (rc) return rc;
    char* d = (char*)ws->stage.ptr;
    PEEB_CUDA(cudaMemcpyAsync(d, stego_host, ib, cudaMemcpyHostToDevice, ws->stream));
    // the kernel addresses bitmap p at p*n: pack them back to back (n % 16 == 0 keeps the fast path)
    for (int k = 0; k < s; ++k) {
        PEEB_REQUIRE(bitmap_ptrs_host[k] != nullptr, "peeb_lsb_recover_h: bitmap %d is null", k);
        PEEB_CUDA(cudaMemcpyAsync(d + o_bm + (size_t)k * n, bitmap_ptrs_host[k], (size_t)n, cudaMemcpyHostToDevice, ws->stream));
    }
    rc = peeb_lsb_recover(ws, d, (const uint8_t*)(d + o_bm), n, itemsize, s, d + o_out, ws->stream);
    if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(cover_out_host, d + o_out, ib, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

int peeb_lsb_extract(peeb_ws* ws, const void* stego, int64_t n, int itemsize, int s, const int64_t* start,
                     const int64_t* len, const int64_t* bit_off, int64_t total_bits, uint8_t* bits_out, void* stream) {
    PEEB_REQUIRE(ws && stego && start && len && bit_off && bits_out, "peeb_lsb_extract: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_lsb_extract: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && total_bits >= 0 && s >= 1 && s <= 8 * itemsize, "peeb_lsb_extract: bad sizes");
    PEEB_REQUIRE(((uintptr_t)bits_out & 3) == 0, "peeb_lsb_extract: bits_out must be 4-byte aligned");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    LsbSegs segs;
    int rc = check_segs(n, s, start, len, bit_off, total_bits, segs);
    if (rc) return rc;
    // bits_out: ceil(total_bits/8) bytes rounded up to whole words + one word of slack
    PEEB_CUDA(cudaMemsetAsync(bits_out, 0, align_up((size_t)(total_bits + 7) / 8, 4) + 4, st));
    if (n == 0 || total_bits == 0) return PEEB_OK;
    long long groups = 0;  // a warp item = 1024 consecutive bits of one segment
    for (int p = 0; p < s; ++p) groups += (len[p] + 1023) / 1024;
    long long bx = (groups + 7) / 8;
    const long long cap = (long long)ws->sm_count * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    ProfScope prof(ws, PEEB_K_LSB_EXTRACT, st);
    if (itemsize == 2) lsb_extract_kernel<2><<<(unsigned)bx, 256, 0, st>>>((const unsigned char*)stego, n, s, segs, (unsigned*)bits_out);
    else lsb_extract_kernel<1><<<(unsigned)bx, 256, 0, st>>>((const unsigned char*)stego, n, s, segs, (unsigned*)bits_out);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int peeb_lsb_extract_h(peeb_ws* ws, const void* stego_host, int64_t n, int itemsize, int s, const int64_t* start,
                       const int64_t* len, const int64_t* bit_off, int64_t total_bits, uint8_t* bits_out_host) {
    PEEB_REQUIRE(ws && stego_host && bits_out_host, "peeb_lsb_extract_h: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_lsb_extract_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && total_bits >= 0, "peeb_lsb_extract_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const size_t ib = (size_t)n * itemsize, ob = align_up((size_t)(total_bits + 7) / 8, 4) + 4;
    int rc = scratch_reserve(ws->stage, ib + 256); if (rc) return rc;
    rc = scratch_reserve(ws->stage2, ob + 256); if (rc) return rc;
    if (ib) PEEB_CUDA(cudaMemcpyAsync(ws->stage.ptr, stego_host, ib, cudaMemcpyHostToDevice, ws->stream));
    rc = peeb_lsb_extract(ws, ws->stage.ptr, n, itemsize, s, start, len, bit_off, total_bits, (uint8_t*)ws->stage2.ptr, ws->stream);
    if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(bits_out_host, ws->stage2.ptr, (size_t)(total_bits + 7) / 8, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

int peeb_compact_bits(peeb_ws* ws, const void* plane, const uint8_t* bitmap, int64_t n, int itemsize, int64_t limit,
                      uint8_t* bits_out, int64_t* count_out, void* stream) {
    PEEB_REQUIRE(ws && plane && bitmap && bits_out && count_out, "peeb_compact_bits: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_compact_bits: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && limit >= 0, "peeb_compact_bits: bad sizes");
    PEEB_REQUIRE(((uintptr_t)bits_out & 3) == 0, "peeb_compact_bits: bits_out must be 4-byte aligned");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (limit > n) limit = n;
    // bits_out: ceil(limit/8) bytes promised to the caller, rounded up to whole words here
    PEEB_CUDA(cudaMemsetAsync(bits_out, 0, align_up((size_t)(limit + 7) / 8, 4), st));
    const int nblocks = (int)((n + CB_TILE - 1) / CB_TILE);
    if (nblocks == 0) { PEEB_CUDA(cudaMemsetAsync(count_out, 0, 8, st)); return PEEB_OK; }
    const size_t need = align_up((size_t)nblocks * 4, 256) + (size_t)nblocks * 8 + 256;
    int rc = scratch_reserve(ws->bits, need); if (rc) return rc;
    int* bc = (int*)ws->bits.ptr;
    long long* bo = (long long*)((char*)ws->bits.ptr + align_up((size_t)nblocks * 4, 256));
    ProfScope p(ws, PEEB_K_COMPACT, st);
    compact_count_kernel<<<nblocks, CB_BLOCK, 0, st>>>(bitmap, n, bc);
    compact_scan_kernel<<<1, 1024, 0, st>>>(bc, nblocks, bo, limit, (long long*)count_out);
    if (itemsize == 2) compact_write_kernel<2><<<nblocks, CB_BLOCK, 0, st>>>((const unsigned char*)plane, bitmap, n, bo, limit, (unsigned*)bits_out);
    else compact_write_kernel<1><<<nblocks, CB_BLOCK, 0, st>>>((const unsigned char*)plane, bitmap, n, bo, limit, (unsigned*)bits_out);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int peeb_compact_bits_h(peeb_ws* ws, const void* plane_host, const uint8_t* bitmap_host, int64_t n, int itemsize,
                        int64_t limit, uint8_t* bits_out_host, int64_t* count_out_host) {
    PEEB_REQUIRE(ws && plane_host && bitmap_host && bits_out_host && count_out_host, "peeb_compact_bits_h: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_compact_bits_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && limit >= 0, "peeb_compact_bits_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (limit > n) limit = n;
    const size_t pb = align_up((size_t)n * itemsize, 256), mb = align_up((size_t)n, 256);
    const size_t ob = align_up((size_t)(limit + 7) / 8, 256) + 256;
    int rc = scratch_reserve(ws->stage, pb + mb + 256); if (rc) return rc;
    rc = scratch_reserve(ws->stage2, ob + 256); if (rc) return rc;
    char* d = (char*)ws->stage.ptr;
    PEEB_CUDA(cudaMemcpyAsync(d, plane_host, (size_t)n * itemsize, cudaMemcpyHostToDevice, ws->stream));
    PEEB_CUDA(cudaMemcpyAsync(d + pb, bitmap_host, (size_t)n, cudaMemcpyHostToDevice, ws->stream));
    uint8_t* dbits = (uint8_t*)ws->stage2.ptr;
    int64_t* dcnt = (int64_t*)((char*)ws->stage2.ptr + ob);
    rc = peeb_compact_bits(ws, d, (const uint8_t*)(d + pb), n, itemsize, limit, dbits, dcnt, ws->stream); if (rc) return rc;
    if (limit) PEEB_CUDA(cudaMemcpyAsync(bits_out_host, dbits, (size_t)(limit + 7) / 8, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaMemcpyAsync(count_out_host, dcnt, 8, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

}  // extern "C"
