// N1: the causal predictor family of Prediction-Error Expansion -- DESIGN.md "Appendix A2".
//
//   predictor  MED (JPEG-LS): a = W, b = N, c = NW, p = clamp(a + b - c, min(a, b), max(a, b));
//   domain     1 <= i < h, 1 <= j < w (row 0 and column 0 never change), ONE pass in raster order;
//   classes, overflow flags, carrier rule, zero padding: as in Appendix A (peeb_pee2.cu).
//
// Embedding predicts from ORIGINAL pixels, so it is embarrassingly parallel: a count kernel, a scan of
// the per-(row, 256-column chunk) counts and an apply kernel, all streaming from global memory.
//
// Extraction predicts from pixels it has already RECOVERED (W, N, NW): an anti-diagonal wavefront.
// One CTA per image; a warp owns 32 consecutive rows (lane = row) and runs them skewed by one column
// per lane, so the N / NW neighbours of a lane are what the lane above produced one and two
// iterations earlier (two shuffles).  Consecutive 32-row groups are pipelined through the warps of
// the CTA: the last row of a group goes to a shared-memory line buffer whose entries carry a tag that
// the first lane of the next group polls (value and ready flag in one word: no fence).  Carrier bits are collected MSB-first per row and
// concatenated at the end by the same CTA.
#include <algorithm>
#include <cstdlib>

#include <vector>

#include "peeb_common.cuh"
#include "peeb_pee.cuh"

namespace peeb {

constexpr int MCHUNK = 256;  // pixels per warp item (8 per lane)

struct MedGeom {
    int h, w, itemsize, maxval;
    int nchunk;  // chunks per row
    int lmw;
    int rw;      // extract staging: 32-bit words per row
};

__device__ __forceinline__ int med3(int a, int b, int c) {
    const int lo = min(a, b), hi = max(a, b);
    return max(min(a + b - c, hi), lo);
}

// eight pixels of a row starting at column j0 (0 beyond the row end); vector load when the rows are aligned
template <typename PixT>
__device__ __forceinline__ void load8(const PixT* row, int j0, int w, bool vec, int (&v)[8]) {
    if (vec && j0 + 8 <= w) {
        if (sizeof(PixT) == 2) {
            const uint4 q = *reinterpret_cast<const uint4*>(row + j0);
            v[0] = q.x & 0xffff; v[1] = q.x >> 16; v[2] = q.y & 0xffff; v[3] = q.y >> 16;
            v[4] = q.z & 0xffff; v[5] = q.z >> 16; v[6] = q.w & 0xffff; v[7] = q.w >> 16;
        } else {
            const uint2 q = *reinterpret_cast<const uint2*>(row + j0);
#pragma unroll
            for (int k = 0; k < 4; ++k) { v[k] = (q.x >> (8 * k)) & 0xff; v[4 + k] = (q.y >> (8 * k)) & 0xff; }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = j0 + k < w ? (int)row[j0 + k] : 0;
    }
}
template <typename PixT>
__device__ __forceinline__ void store8(PixT* row, int j0, int w, bool vec, const int (&v)[8]) {
    if (vec && j0 + 8 <= w) {
        if (sizeof(PixT) == 2) {
            uint4 q;
            q.x = v[0] | (v[1] << 16); q.y = v[2] | (v[3] << 16); q.z = v[4] | (v[5] << 16); q.w = v[6] | (v[7] << 16);
            *reinterpret_cast<uint4*>(row + j0) = q;
        } else {
            uint2 q;
            q.x = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
            q.y = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
            *reinterpret_cast<uint2*>(row + j0) = q;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) if (j0 + k < w) row[j0 + k] = (PixT)v[k];
    }
}

// eight pixels of a row starting at column j0, still packed (16-bit pixels: x..w; 8-bit pixels: x, y); 0 beyond
// the row end.  Split from the unpacking so that the next row can be in flight while a row is processed.
template <typename PixT>
__device__ __forceinline__ uint4 fetch8(const PixT* row, int j0, int w, bool vec) {
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (vec && j0 + 8 <= w) {
        if (sizeof(PixT) == 2) q = *reinterpret_cast<const uint4*>(row + j0);
        else { const uint2 t = *reinterpret_cast<const uint2*>(row + j0); q.x = t.x; q.y = t.y; }
    } else {
        unsigned v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = j0 + k < w ? (unsigned)row[j0 + k] : 0u;
        if (sizeof(PixT) == 2) { q.x = v[0] | (v[1] << 16); q.y = v[2] | (v[3] << 16); q.z = v[4] | (v[5] << 16); q.w = v[6] | (v[7] << 16); }
        else { q.x = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); q.y = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24); }
    }
    return q;
}
template <typename PixT>
__device__ __forceinline__ void unpack8(const uint4& q, int (&v)[8]) {
    if (sizeof(PixT) == 2) {
        v[0] = q.x & 0xffff; v[1] = q.x >> 16; v[2] = q.y & 0xffff; v[3] = q.y >> 16;
        v[4] = q.z & 0xffff; v[5] = q.z >> 16; v[6] = q.w & 0xffff; v[7] = q.w >> 16;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = (q.x >> (8 * k)) & 0xff; v[4 + k] = (q.y >> (8 * k)) & 0xff; }
    }
}

// One warp item = (unit, block of `rb` consecutive rows, chunk); lane -> 8 pixels of every row of the block.
// The warp walks down its rows: the row above a row is the row it has just read (each pixel is
// loaded once), SSE / flag statistics are reduced once per item.  APPLY = false: count the carriers
// of every (row, chunk); APPLY = true: write marked pixels, the location-map byte of every 8 columns,
// SSE / flag statistics.  cnt / off are indexed [(unit * h + row) * nchunk + chunk] (raster order).
//
// A row first runs the FAST code: it assumes that no pixel over/underflows (every expandable pixel is
// a carrier, the location-map byte is 0) and only watches for a value leaving [0, maxval); pixels
// outside the domain (column 0, row 0, lanes past the row end) run with T = 0, which makes them
// "shifted by 0".  If any lane of the warp sees a value leave the range, or a lane straddles the row
// end, the row is redone by the GENERIC code (per-pixel domain test, flags, location map).
struct MedRow {
    int ncar;
    unsigned carmask, lmbyte;
    int nflag;
};
template <bool APPLY>
__device__ __forceinline__ void med_row_generic(const int (&x)[8], const int (&up)[8], int left, int upleft, unsigned rmask,
                                                int T, int maxval, int (&nv)[8], MedRow& r) {
    const int T2 = 2 * T;
    r.ncar = 0; r.carmask = 0; r.lmbyte = 0; r.nflag = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int a = k ? x[k - 1] : left, c = k ? up[k - 1] : upleft;
        const int t = x[k] - med3(a, up[k], c) + T;                      // e + T
        const bool expd = (unsigned)t < (unsigned)T2;                    // -T <= e < T
        const int nv0 = x[k] + max(min(t, T2), 0) - T;                   // x + e | x + T | x - T
        const bool ok = (unsigned)nv0 <= (unsigned)maxval - (expd ? 1u : 0u);
        const bool in = (rmask >> k) & 1u;
        if (in && expd && ok) { ++r.ncar; r.carmask |= 1u << k; }
        nv[k] = (in && ok) ? nv0 : x[k];
        if (APPLY && in && !ok) { r.lmbyte |= 0x80u >> k; ++r.nflag; }
    }
}
// returns true when some value left [0, maxval): the caller falls back to the generic code
__device__ __forceinline__ bool med_row_fast(const int (&x)[8], const int (&up)[8], int left, int upleft, int T0, int T,
                                             int maxval, int (&nv)[8], MedRow& r) {
    unsigned top = 0;
    r.ncar = 0; r.carmask = 0; r.lmbyte = 0; r.nflag = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int a = k ? x[k - 1] : left, c = k ? up[k - 1] : upleft;
        const int Tk = k ? T : T0;
        const int t = x[k] - med3(a, up[k], c) + Tk;
        nv[k] = x[k] + max(min(t, 2 * Tk), 0) - Tk;
        top = max(top, (unsigned)nv[k]);
        if ((unsigned)t < (unsigned)(2 * Tk)) { ++r.ncar; r.carmask |= 1u << k; }
    }
    return top >= (unsigned)maxval;
}

template <typename PixT, bool APPLY>
__global__ void __launch_bounds__(256) med_embed_kernel(MedGeom g, PeeBatch bt, int rb, unsigned short* __restrict__ cnt,
                                                        const unsigned* __restrict__ off) {
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nblk = (g.h + rb - 1) / rb;
    const long long per_unit = (long long)nblk * g.nchunk;
    if (item >= per_unit * bt.n_units) return;
    const int unit = (int)(item / per_unit);
    if (bt.active && !bt.active[unit]) return;  // threshold search: this unit is done
    const int rem = (int)(item - (long long)unit * per_unit);
    const int blk = rem / g.nchunk, chunk = rem - blk * g.nchunk;
    const int row0 = blk * rb, row1 = min(row0 + rb, g.h);
    const unsigned char* ubase = bt.src + (long long)unit * bt.src_stride;
    const bool vec_in = ((((uintptr_t)ubase) | ((uintptr_t)g.w * sizeof(PixT))) & (8 * sizeof(PixT) - 1)) == 0;
    unsigned char* obase = APPLY && bt.dst ? bt.dst + (long long)unit * bt.dst_stride : nullptr;
    const bool vec_out = ((((uintptr_t)obase) | ((uintptr_t)g.w * sizeof(PixT))) & (8 * sizeof(PixT) - 1)) == 0;
    const int j0 = chunk * MCHUNK + 8 * lane;
    const int T = bt.T[unit];
    const unsigned n_bits = bt.n_bits[unit];
    const unsigned* pay = reinterpret_cast<const unsigned*>(bt.payload + (long long)unit * bt.payload_stride);
    unsigned vmask = 0;  // pixels of this lane inside the domain 1 <= j < w
#pragma unroll
    for (int k = 0; k < 8; ++k) vmask |= (j0 + k >= 1 && j0 + k < g.w) ? 1u << k : 0u;
    // fast code: a lane is either inside the row or past its end (T = 0 there); its first pixel may be column 0
    const bool straddle = __any_sync(0xffffffffu, j0 < g.w && j0 + 8 > g.w);
    const int Tl = j0 + 8 <= g.w ? T : 0, Tl0 = j0 >= 1 ? Tl : 0;
    const PixT* cur = reinterpret_cast<const PixT*>(ubase) + (size_t)row0 * g.w;
    int up[8], upleft = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) up[k] = 0;
    if (row0 >= 1) {
        load8<PixT>(cur - g.w, j0, g.w, vec_in, up);
        upleft = __shfl_up_sync(0xffffffffu, up[7], 1);
        if (lane == 0) upleft = j0 > 0 ? (int)cur[j0 - 1 - g.w] : 0;
    }
    long long sse = 0;
    int nflag = 0;
    long long e = ((long long)unit * g.h + row0) * g.nchunk + chunk;  // index of (row, chunk) in cnt / off
    // the next row is fetched while a row is processed (one 16-byte load per lane and row would
    // otherwise leave the memory system mostly idle)
    uint4 qc = fetch8<PixT>(cur, j0, g.w, vec_in);
    int lc = (lane == 0 && j0 > 0) ? (int)cur[j0 - 1] : 0;
    for (int row = row0; row < row1; ++row, cur += g.w, e += g.nchunk) {
        int x[8], nv[8];
        uint4 qn = make_uint4(0u, 0u, 0u, 0u);
        int ln = 0;
        if (row + 1 < row1) {
            qn = fetch8<PixT>(cur + g.w, j0, g.w, vec_in);
            if (lane == 0 && j0 > 0) ln = (int)cur[g.w + j0 - 1];
        }
        unpack8<PixT>(qc, x);
        // W of the lane's first pixel: the previous lane's last pixel (lane 0: one scalar load)
        int left = __shfl_up_sync(0xffffffffu, x[7], 1);
        if (lane == 0) left = lc;
        MedRow r;
        bool redo = straddle;
        if (!redo) {
            const bool inrow = row >= 1;  // row 0 never changes
            redo = __any_sync(0xffffffffu, med_row_fast(x, up, left, upleft, inrow ? Tl0 : 0, inrow ? Tl : 0, g.maxval, nv, r));
        }
        if (redo) med_row_generic<APPLY>(x, up, left, upleft, row >= 1 ? vmask : 0u, T, g.maxval, nv, r);
        int ncar = r.ncar;
        if constexpr (!APPLY) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ncar += __shfl_xor_sync(0xffffffffu, ncar, o);
            if (lane == 0) cnt[e] = (unsigned short)ncar;
        } else {
            // rank of the lane's first carrier inside the chunk, then the chunk's offset in the unit's raster order
            int incl = ncar;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (ncar) {
                const unsigned base = off[e] + (unsigned)(incl - ncar);
                unsigned win = 0;  // bits base .. base+7 on top, zero past n_bits
                if (base < n_bits) {
                    const unsigned w0 = __byte_perm(__ldg(pay + (base >> 5)), 0, 0x0123), w1 = __byte_perm(__ldg(pay + (base >> 5) + 1), 0, 0x0123);
                    win = __funnelshift_l(w1, w0, base & 31);
                    if (n_bits - base < 32u) win &= ~(0xffffffffu >> (n_bits - base));
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (r.carmask & (1u << k)) { nv[k] += (int)(win >> 31); win <<= 1; }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) { const int d = nv[k] - x[k]; sse += (long long)d * d; }
            nflag += r.nflag;
            if (obase) store8<PixT>(reinterpret_cast<PixT*>(obase) + (size_t)row * g.w, j0, g.w, vec_out, nv);
            if (bt.lm && j0 < g.w) bt.lm[(long long)unit * bt.lm_stride + (size_t)row * g.lmw + (j0 >> 3)] = (unsigned char)r.lmbyte;
        }
        upleft = left;
#pragma unroll
        for (int k = 0; k < 8; ++k) up[k] = x[k];
        qc = qn; lc = ln;
    }
    if constexpr (APPLY) {
        sse = warp_sum_i64(sse);
        nflag = (int)warp_sum_i64(nflag);
        if (lane == 0) {
            long long* info = bt.info + (long long)unit * PEEB_INFO;
            if (sse) atomicAdd(reinterpret_cast<unsigned long long*>(info + 6), (unsigned long long)sse);
            if (nflag) atomicAdd(reinterpret_cast<unsigned long long*>(info + 5), (unsigned long long)nflag);
        }
    }
}

// one CTA per unit: exclusive scan of the chunk counts (raster order), capacity into info
__global__ void __launch_bounds__(1024) med_scan_kernel(MedGeom g, PeeBatch bt, const unsigned short* __restrict__ cnt,
                                                        unsigned* __restrict__ off) {
    __shared__ int ws[33];
    const int unit = blockIdx.x;
    if (bt.active && !bt.active[unit]) return;  // threshold search: this unit is done
    const long long ne = (long long)g.h * g.nchunk;
    const unsigned short* c = cnt + unit * ne;
    unsigned* o = off + unit * ne;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned carry = 0;
    for (long long base = 0; base < ne; base += blockDim.x) {
        const long long idx = base + threadIdx.x;
        const int v = idx < ne ? c[idx] : 0;
        int incl = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += t;
        }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int x = ws[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, x, s);
                if (lane >= s) x += t;
            }
            ws[lane] = x;
            if (lane == 31) ws[32] = x;
        }
        __syncthreads();
        if (idx < ne) o[idx] = carry + (unsigned)((warp ? ws[warp - 1] : 0) + incl - v);
        carry += (unsigned)ws[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        long long* info = bt.info + (long long)unit * PEEB_INFO;
        info[0] = bt.T[unit]; info[1] = bt.n_bits[unit]; info[2] = carry; info[3] = carry; info[4] = 0;
        info[7] = ((long long)bt.n_bits[unit] > (long long)carry) ? PEEB_E_CAPACITY : 0;
    }
}

// ------------------------------------------------------------------ wavefront extract
struct MedSmem {
    size_t line, prog, rowoff, misc, total;
};
__host__ __device__ inline MedSmem med_layout(const MedGeom& g, int nwarps) {
    MedSmem L{};
    size_t o = 0;
    L.line = o; o += align_up((size_t)(nwarps + 1) * g.w * sizeof(unsigned), 16);
    L.prog = o;
    L.rowoff = o; o += align_up((size_t)(g.h + 1) * sizeof(int), 16);
    L.misc = o; o += 64 * sizeof(int);
    L.total = o;
    return L;
}

__device__ __forceinline__ uint4 lds128_volatile(const unsigned* p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts128_volatile(unsigned* p, const uint4& v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// eight pixels of a row as four words of two 16-bit values (8-bit pixels are widened), and back
template <typename PixT> __device__ __forceinline__ uint4 ld8q(const PixT* p) {
    if (sizeof(PixT) == 2) return *reinterpret_cast<const uint4*>(p);
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    return make_uint4(__byte_perm(v.x, 0, 0x4140), __byte_perm(v.x, 0, 0x4342), __byte_perm(v.y, 0, 0x4140), __byte_perm(v.y, 0, 0x4342));
}
template <typename PixT> __device__ __forceinline__ void st8q(PixT* p, const uint4& q) {
    if (sizeof(PixT) == 2) *reinterpret_cast<uint4*>(p) = q;
    else *reinterpret_cast<uint2*>(p) = make_uint2(__byte_perm(q.x, q.y, 0x6420), __byte_perm(q.z, q.w, 0x6420));
}

// VEC: rows are 16-byte (8-bit pixels: 8-byte) aligned and w % 8 == 0: a lane streams its row through
// two register queues (eight pixels in, eight out, the next block and location-map byte prefetched
// one block ahead), so no memory latency sits on the wavefront's dependency chain.
// Cluster form (CS = 2): the 32-row groups of an image are dealt to the warps of TWO CTAs (blocks of nwarps
// consecutive groups per CTA), so that a few large images use twice as many SMs.  Every CTA keeps the whole ring of
// line buffers; a group's last row is written into the shared memory of the CTA that runs the next group
// (st.shared::cluster when that is the peer: one group in nwarps), which polls it locally as before.  The per-row
// carrier counts go to both CTAs, which then share the rows of the final concatenation.  CS = 1 is the plain launch.
__device__ __forceinline__ unsigned med_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void med_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned med_peer_addr(const void* own, unsigned cta) {
    unsigned remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(own)), "r"(cta));
    return remote;
}
__device__ __forceinline__ void st_peer_v4(unsigned addr, const uint4& v) {
    asm volatile("st.volatile.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_peer_u32(unsigned addr, unsigned v) {
    asm volatile("st.volatile.shared::cluster.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}

template <typename PixT, bool VEC>
__global__ void __launch_bounds__(512) med_extract_kernel(MedGeom g, PeeBatch bt, unsigned* __restrict__ stage_bits, int CS) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int rank = CS > 1 ? (int)med_cluster_rank() : 0, allwarps = CS * nwarps;
    const MedSmem L = med_layout(g, allwarps);
    // line buffers: one 32-bit entry per column = recovered value | tag << 16.  The tag (group + 2; row 0:
    // 1; never written: 0) makes an entry its own "ready" flag, so handing a row to the next group needs
    // neither a fence nor a separate progress counter.
    unsigned* line = reinterpret_cast<unsigned*>(smem_raw + L.line);
    int* rowoff = reinterpret_cast<int*>(smem_raw + L.rowoff);
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);
    const int unit = blockIdx.x / CS, h = g.h, w = g.w, nslot = allwarps + 1;
    const PixT* marked = reinterpret_cast<const PixT*>(bt.src + (long long)unit * bt.src_stride);
    PixT* rec = bt.dst ? reinterpret_cast<PixT*>(bt.dst + (long long)unit * bt.dst_stride) : nullptr;
    const unsigned char* lm = bt.lm + (long long)unit * bt.lm_stride;
    unsigned* stage = stage_bits + (long long)unit * h * g.rw;
    const int T = bt.T[unit], T2 = 2 * T, T4 = 4 * T;
    for (int k = threadIdx.x; k <= h; k += blockDim.x) rowoff[k] = 0;
    for (int k = threadIdx.x; k < nslot * w; k += blockDim.x) line[k] = 0u;
    __syncthreads();
    // row 0 never changes: it is the "group -1" line (slot nslot-1, tag 1) and goes straight to the output
    for (int j = threadIdx.x; j < w; j += blockDim.x) {
        const unsigned v = marked[j];
        line[(size_t)(nslot - 1) * w + j] = v | (1u << 16);
        if (rec && rank == 0) rec[j] = (PixT)v;
    }
    __syncthreads();
    if (CS > 1) med_cluster_sync();  // the peer's line buffers are cleared before anything is written into them

    const int ngroups = (h - 1 + 31) / 32;
    for (int grp = rank * nwarps + warp; grp < ngroups; grp += allwarps) {
        const int rowi = 1 + 32 * grp + lane;
        const bool valid = rowi < h;
        const int rowc = valid ? rowi : 1;
        const int lastl = min(31, h - 2 - 32 * grp);  // last lane that owns a row
        const PixT* mrow = marked + (size_t)rowc * w;
        PixT* rrow = rec ? rec + (size_t)rowc * w : nullptr;
        const unsigned char* lrow = lm + (size_t)rowc * g.lmw;
        unsigned* srow = stage + (size_t)rowc * g.rw;
        volatile unsigned* myline = line + (size_t)(grp % nslot) * w;
        const volatile unsigned* upline = line + (size_t)((grp + nslot - 1) % nslot) * w;
        const unsigned mytag = (unsigned)(grp + 2) << 16, uptag = (unsigned)(grp + 1);
        const bool writer = lane == lastl;
        // the next group runs in the CTA of warp (grp + 1) mod allwarps: its copy of this group's line is written
        const int cons = ((grp + 1) % allwarps) / nwarps;
        const bool remote = cons != rank;
        const unsigned peerline = remote ? med_peer_addr(const_cast<const unsigned*>(myline), (unsigned)cons) : 0u;
        const unsigned peeroff = CS > 1 ? med_peer_addr(rowoff, (unsigned)(rank ^ 1)) : 0u;
        if (VEC) {
            // ---- four columns per lane and step (skew of four columns per lane): the step overhead -- shuffles,
            // the poll of the line above, queue handling, loop control -- is paid once per four pixels.
            // A lane holds its last four recovered values packed in two words (curA: columns jb, jb+1;
            // curB: jb+2, jb+3) and the one before them (plast): exactly what the lane below needs next step.
            unsigned curA = 0, curB = 0, plast = 0;
            int a = 0, ulast = 0, ncar = 0, nW = 0;
            unsigned long long Wq = 0;
            unsigned lmbyte = 0, lmnext = 0;
            uint4 q = make_uint4(0, 0, 0, 0), nx = q;
            unsigned oA = 0, oB = 0;
            if (valid) {
                q = ld8q<PixT>(mrow);
                lmbyte = lrow[0];
                if (w > 8) { nx = ld8q<PixT>(mrow + 8); lmnext = lrow[1]; }
            }
            // running pointers of the block to fetch next, its location-map byte and the block to write next
            // (kept in registers: the addresses are not rebuilt from the row index at every block)
            const PixT* mnx = mrow + 16;
            const unsigned char* lnx = lrow + 2;
            PixT* rout = rrow;
            const int nsteps = w >> 2;
            int jb = -4 * lane;
            for (int t = 0; t < nsteps + 31; ++t, jb += 4) {
                unsigned uA = __shfl_up_sync(0xffffffffu, curA, 1), uB = __shfl_up_sync(0xffffffffu, curB, 1);
                int um1 = (int)__shfl_up_sync(0xffffffffu, plast, 1);
                if (lane == 0 && jb < w) {
                    uint4 v;
                    const unsigned tg = uptag << 16;
                    for (;;) {
                        v = lds128_volatile(const_cast<const unsigned*>(upline) + jb);
                        // all four tags equal uptag <=> no bit differs in any upper half
                        if (((((v.x ^ tg) | (v.y ^ tg)) | ((v.z ^ tg) | (v.w ^ tg))) & 0xffff0000u) == 0u) break;
                        __nanosleep(20);  // the group above is about one step away: do not burn issue slots on the poll
                    }
                    uA = (v.x & 0xffffu) | (v.y << 16); uB = (v.z & 0xffffu) | (v.w << 16);
                    um1 = ulast; ulast = (int)(v.w & 0xffffu);
                }
                __syncwarp();
                if (valid && jb >= 0 && jb < w) {
                    const bool second = jb & 4;
                    const unsigned xa = second ? q.z : q.x, xb = second ? q.w : q.y;
                    // location-map nibble of these four columns (bit 3 = column jb); column 0 is border: treated as flagged
                    unsigned fl = (lmbyte >> (second ? 0 : 4)) & 0xfu;
                    if (jb == 0) fl |= 8u;
                    const int xs[4] = {(int)(xa & 0xffffu), (int)(xa >> 16), (int)(xb & 0xffffu), (int)(xb >> 16)};
                    const int us[4] = {(int)(uA & 0xffffu), (int)(uA >> 16), (int)(uB & 0xffffu), (int)(uB >> 16)};
                    int vals[4];
                    unsigned wb = 0;
                    int nb = 0, c = um1;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool flagged = (fl >> (3 - k)) & 1u;
                        const int Tl = flagged ? 0 : T, Tl2 = flagged ? 0 : T2;
                        const int u = xs[k] - med3(a, us[k], c) + Tl2;                // e' + 2T
                        const int cc = max(min((u + 1) >> 1, Tl2), 0);                // clamp(ceil(e'/2), -T, T) + T
                        const int val = xs[k] - cc + Tl;
                        if ((unsigned)u < (unsigned)(flagged ? 0 : T4)) { wb = (wb << 1) | (unsigned)(u & 1); ++nb; }
                        vals[k] = val; a = val; c = us[k];
                    }
                    plast = curB >> 16;
                    curA = (unsigned)vals[0] | ((unsigned)vals[1] << 16);
                    curB = (unsigned)vals[2] | ((unsigned)vals[3] << 16);
                    if (nb) {
                        Wq = (Wq << nb) | wb; nW += nb; ncar += nb;
                        if (nW >= 32) { *srow++ = (unsigned)(Wq >> (nW - 32)); nW -= 32; }
                    }
                    if (writer) {
                        uint4 e;
                        e.x = (unsigned)vals[0] | mytag; e.y = (unsigned)vals[1] | mytag; e.z = (unsigned)vals[2] | mytag; e.w = (unsigned)vals[3] | mytag;
                        if (remote) st_peer_v4(peerline + 4u * (unsigned)jb, e);
                        else sts128_volatile(const_cast<unsigned*>(myline) + jb, e);
                    }
                    if (!second) { oA = curA; oB = curB; }
                    else {  // block of eight done: write it, switch to the prefetched one, fetch the one after
                        if (rrow) st8q<PixT>(rout, make_uint4(oA, oB, curA, curB));
                        rout += 8;
                        q = nx; lmbyte = lmnext;
                        if (jb + 12 < w) { nx = ld8q<PixT>(mnx); lmnext = *lnx; }
                        mnx += 8; ++lnx;
                    }
                }
            }
            if (valid) {
                if (nW) *srow = (unsigned)(Wq << (32 - nW));
                rowoff[rowi] = ncar;
                if (CS > 1) st_peer_u32(peeroff + 4u * (unsigned)rowi, (unsigned)ncar);
            }
            continue;
        }
        int cur1 = 0, cur2 = 0, aprev = 0, bprev0 = 0;
        unsigned W = 0, lmbyte = 0, lmnext = 0;
        int nW = 0, ncar = 0;
        uint4 q = make_uint4(0, 0, 0, 0), nx = q, o = q;
        if (VEC && valid) {
            q = ld8q<PixT>(mrow);
            lmbyte = lrow[0];
            if (w > 8) { nx = ld8q<PixT>(mrow + 8); lmnext = lrow[1]; }
        }
        int j = -lane;
        for (int t = 0; t < w + 31; ++t, ++j) {
            int b = __shfl_up_sync(0xffffffffu, cur1, 1), c = __shfl_up_sync(0xffffffffu, cur2, 1);
            if (lane == 0 && j < w) {
                unsigned v;
                do { v = upline[j]; } while ((v >> 16) != uptag);
                c = bprev0; b = (int)(v & 0xffffu); bprev0 = b;
            }
            __syncwarp();
            if (valid && j >= 0 && j < w) {
                int x;
                if (VEC) {
                    x = (int)(q.x & 0xffffu);
                    q.x = __funnelshift_r(q.x, q.y, 16); q.y = __funnelshift_r(q.y, q.z, 16);
                    q.z = __funnelshift_r(q.z, q.w, 16); q.w >>= 16;
                } else {
                    x = mrow[j];
                    if ((j & 7) == 0) lmbyte = lrow[j >> 3];
                }
                int val = x;
                if (j >= 1) {
                    const bool flagged = (lmbyte >> (7 - (j & 7))) & 1u;           // flagged pixels were left alone
                    const int Tl = flagged ? 0 : T, Tl2 = flagged ? 0 : T2;
                    const int u = x - med3(aprev, b, c) + Tl2;                     // e' + 2T
                    const int cc = max(min((u + 1) >> 1, Tl2), 0);                 // clamp(ceil(e'/2), -T, T) + T
                    val = x - cc + Tl;
                    if ((unsigned)u < (unsigned)(flagged ? 0 : T4)) {              // carrier: -2T <= e' < 2T
                        W = (W << 1) | (unsigned)(u & 1);
                        ++ncar;
                        if (++nW == 32) { *srow++ = W; nW = 0; W = 0; }
                    }
                }
                cur2 = cur1; cur1 = val; aprev = val;
                if (writer) {
                    if (remote) st_peer_u32(peerline + 4u * (unsigned)j, (unsigned)val | mytag);
                    else myline[j] = (unsigned)val | mytag;
                }
                if (VEC) {
                    o.x = __funnelshift_r(o.x, o.y, 16); o.y = __funnelshift_r(o.y, o.z, 16);
                    o.z = __funnelshift_r(o.z, o.w, 16); o.w = (o.w >> 16) | ((unsigned)val << 16);
                    if ((j & 7) == 7) {  // block done: write it, switch to the prefetched one, fetch the one after
                        if (rrow) st8q<PixT>(rrow + j - 7, o);
                        q = nx; lmbyte = lmnext;
                        if (j + 9 < w) { nx = ld8q<PixT>(mrow + j + 9); lmnext = lrow[(j + 9) >> 3]; }
                    }
                } else if (rrow) {
                    rrow[j] = (PixT)val;
                }
            }
        }
        if (valid) {
            if (nW) *srow = W << (32 - nW);
            rowoff[rowi] = ncar;
            if (CS > 1) st_peer_u32(peeroff + 4u * (unsigned)rowi, (unsigned)ncar);
        }
    }
    __syncthreads();
    if (CS > 1) med_cluster_sync();  // both CTAs now hold every row's count; the staged bits are visible to the peer

    // ---- concatenate the rows' bit streams (raster order) into the unit's payload, MSB-first
    const int total = block_excl_scan(rowoff, h, misc);
    __syncthreads();
    const long long n_bits = bt.n_bits[unit];
    if (threadIdx.x == 0 && rank == 0) {
        long long* info = bt.info + (long long)unit * PEEB_INFO;
        info[0] = T; info[1] = n_bits; info[2] = total; info[3] = total; info[4] = 0; info[5] = 0; info[6] = 0;
        info[7] = n_bits > total ? PEEB_E_CAPACITY : 0;
    }
    unsigned* out = reinterpret_cast<unsigned*>(bt.payload_out + (long long)unit * bt.payload_stride);
    for (int r = 1 + rank * nwarps + warp; r < h; r += allwarps) {
        const long long before = rowoff[r];
        const int cnt = (r + 1 < h ? rowoff[r + 1] : total) - rowoff[r];
        if (cnt == 0 || before >= n_bits) continue;
        const unsigned* src = stage + (size_t)r * g.rw;
        const int nsrc = (cnt + 31) >> 5;
        const long long first = before >> 5, last = (before + cnt - 1) >> 5;
        const int sh = (int)(before & 31);
        for (long long mw = first + lane; mw <= last; mw += 32) {
            const int i = (int)(mw - first);
            const unsigned cur = i < nsrc ? src[i] : 0u;
            const unsigned prev = (i >= 1 && i - 1 < nsrc) ? src[i - 1] : 0u;
            unsigned val = __funnelshift_r(cur, prev, sh);
            // the last staged word of a row may carry stale low bits only if cnt is a multiple of 32: none are written
            const long long bit0 = mw << 5;
            const long long endbit = before + cnt;                       // bits of this row end here
            if (bit0 + 32 > endbit) val &= ~(0xffffffffu >> (int)(endbit - bit0));   // drop what is past the row's bits
            if (bit0 + 32 > n_bits) {
                const int keep = (int)(n_bits - bit0);
                val = keep <= 0 ? 0u : (val & ~(0xffffffffu >> keep));
            }
            if (val == 0) continue;
            atomicOr(out + mw, __byte_perm(val, 0, 0x0123));
        }
    }
}

// ------------------------------------------------------------------ host side
static int med_geom(int h, int w, int itemsize, int bit_depth, MedGeom& g) {
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "pee_med: itemsize must be 1 or 2");
    PEEB_REQUIRE(bit_depth >= 1 && bit_depth <= 8 * itemsize, "pee_med: bit_depth %d out of range for itemsize %d", bit_depth, itemsize);
    PEEB_REQUIRE(h >= 1 && w >= 1 && (long long)h * w < (1ll << 31), "pee_med: image size %dx%d unsupported", h, w);
    PEEB_REQUIRE(w <= 65535, "pee_med: width %d unsupported (progress counters)", w);
    g.h = h; g.w = w; g.itemsize = itemsize; g.maxval = (1 << bit_depth) - 1;
    g.nchunk = (w + MCHUNK - 1) / MCHUNK;
    g.lmw = (w + 7) / 8;
    g.rw = (w + 31) / 32 + 1;
    return PEEB_OK;
}

}  // namespace peeb

using namespace peeb;

extern "C" {

int peeb_pee_med_embed_batch(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                             int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload,
                             int64_t payload_stride, void* marked, int64_t marked_stride, uint8_t* lm, int64_t lm_stride,
                             int64_t* info, void* stream) {
    PEEB_REQUIRE(ws && src && n_bits && payload && info, "peeb_pee_med_embed_batch: null pointer");  // T may be null: searched on the device
    PEEB_REQUIRE(n_units >= 1, "peeb_pee_med_embed_batch: n_units must be >= 1");
    PEEB_REQUIRE(((uintptr_t)payload & 3) == 0 && (payload_stride & 3) == 0, "peeb_pee_med_embed_batch: payload must be 4-byte aligned");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    MedGeom g;
    int rc = med_geom(h, w, itemsize, bit_depth, g);
    if (rc) return rc;
    const long long ne = (long long)n_units * h * g.nchunk;
    PEEB_REQUIRE(ne < (1ll << 31), "peeb_pee_med_embed_batch: batch too large");
    int* dT; unsigned* dN; char* extra;
    const size_t cnt_bytes = align_up((size_t)ne * sizeof(unsigned short), 256), off_bytes = align_up((size_t)ne * sizeof(unsigned), 256);
    // T == NULL: every unit starts at T = 1 and is embedded again at T + 1 while its payload does not
    // fit -- on the device, only the units that fall short take part in a round (the causal predictor has no histogram estimate)
    const bool auto_T = T == nullptr;
    const size_t act_bytes = auto_T ? align_up((size_t)(n_units + 1) * sizeof(int), 256) : 0;
    std::vector<int32_t> ones;
    if (auto_T) ones.assign((size_t)n_units, 1);
    rc = upload_unit_tables(ws, 0, n_units, auto_T ? ones.data() : T, n_bits, bit_depth, cnt_bytes + off_bytes + act_bytes, st, &dT, &dN, &extra);
    if (rc) return rc;
    unsigned short* cnt = (unsigned short*)extra;
    unsigned* off = (unsigned*)(extra + cnt_bytes);
    int* active = auto_T ? (int*)(extra + cnt_bytes + off_bytes) : nullptr;
    PEEB_CUDA(cudaMemsetAsync(info, 0, sizeof(int64_t) * PEEB_INFO * n_units, st));
    PeeBatch bt{};
    bt.src = (const unsigned char*)src; bt.src_stride = src_stride;
    bt.dst = (unsigned char*)marked; bt.dst_stride = marked_stride;
    bt.lm = lm; bt.lm_stride = lm_stride;
    bt.payload = payload; bt.payload_stride = payload_stride;
    bt.T = dT; bt.n_bits = dN; bt.info = (long long*)info; bt.n_units = n_units;
    bt.active = active;
    if (auto_T) PEEB_CUDA(cudaMemsetAsync(active, 0x01, (size_t)n_units * sizeof(int), st));  // every unit takes part in round one
    // rows per warp item: long enough to amortise the set-up, short enough to leave ~8 items per
    // resident warp slot for load balance
    int rb = 16;
    while (rb > 1 && (long long)n_units * ((h + rb - 1) / rb) * g.nchunk < (long long)ws->sm_count * 64 * 4) rb >>= 1;
    if (const char* e = getenv("PEEB_MED_ROWS")) rb = std::max(1, atoi(e));
    const long long nitems = (long long)n_units * ((h + rb - 1) / rb) * g.nchunk;
    const unsigned blocks = (unsigned)((nitems + 7) / 8);
    const int tmax = 1 << (bit_depth - 1);
    int* remaining_h = auto_T ? (int*)((char*)ws->ptable_h_cur[0] + align_up((size_t)n_units * 8, 256)) : nullptr;
    for (int round = 0; round <= (auto_T ? tmax : 0); ++round) {
        { ProfScope p(ws, PEEB_K_PEE_COUNT, st);
          if (itemsize == 2) med_embed_kernel<unsigned short, false><<<blocks, 256, 0, st>>>(g, bt, rb, cnt, off);
          else med_embed_kernel<unsigned char, false><<<blocks, 256, 0, st>>>(g, bt, rb, cnt, off); }
        med_scan_kernel<<<n_units, 1024, 0, st>>>(g, bt, cnt, off);
        { ProfScope p(ws, PEEB_K_PEE_EMBED, st);
          if (itemsize == 2) med_embed_kernel<unsigned short, true><<<blocks, 256, 0, st>>>(g, bt, rb, cnt, off);
          else med_embed_kernel<unsigned char, true><<<blocks, 256, 0, st>>>(g, bt, rb, cnt, off); }
        PEEB_CUDA(cudaGetLastError());
        if (!auto_T) break;
        rc = threshold_retry_round(n_units, tmax, (long long*)info, dT, active, active + n_units, remaining_h, st);
        if (rc) return rc;
        if (*remaining_h == 0) break;
    }
    return PEEB_OK;
}

int peeb_pee_med_extract_batch(peeb_ws* ws, const void* marked, int64_t marked_stride, int n_units, int h, int w,
                               int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* lm,
                               int64_t lm_stride, uint8_t* payload_out, int64_t payload_stride, void* recovered,
                               int64_t recovered_stride, int64_t* info, void* stream) {
    PEEB_REQUIRE(ws && marked && T && n_bits && lm && payload_out && info, "peeb_pee_med_extract_batch: null pointer");
    PEEB_REQUIRE(n_units >= 1, "peeb_pee_med_extract_batch: n_units must be >= 1");
    PEEB_REQUIRE(((uintptr_t)payload_out & 3) == 0 && (payload_stride & 3) == 0, "peeb_pee_med_extract_batch: payload_out must be 4-byte aligned");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    MedGeom g;
    int rc = med_geom(h, w, itemsize, bit_depth, g);
    if (rc) return rc;
    for (int u = 0; u < n_units; ++u)
        PEEB_REQUIRE(n_units == 1 || (int64_t)peeb_payload_bytes(n_bits[u]) <= payload_stride, "peeb_pee_med_extract_batch: payload_stride too small for unit %d", u);
    int* dT; unsigned* dN; char* extra;
    rc = upload_unit_tables(ws, 0, n_units, T, n_bits, bit_depth, 256, st, &dT, &dN, &extra);
    if (rc) return rc;
    rc = scratch_reserve(ws->pbits[0], (size_t)n_units * h * g.rw * sizeof(unsigned) + 256);
    if (rc) return rc;
    if (n_units == 1) PEEB_CUDA(cudaMemsetAsync(payload_out, 0, peeb_payload_bytes(n_bits[0]), st));
    else PEEB_CUDA(cudaMemsetAsync(payload_out, 0, (size_t)payload_stride * n_units, st));
    PeeBatch bt{};
    bt.src = (const unsigned char*)marked; bt.src_stride = marked_stride;
    bt.dst = (unsigned char*)recovered; bt.dst_stride = recovered_stride;
    bt.lm = const_cast<uint8_t*>(lm); bt.lm_stride = lm_stride;
    bt.payload_out = payload_out; bt.payload_stride = payload_stride;
    bt.T = dT; bt.n_bits = dN; bt.info = (long long*)info; bt.n_units = n_units;
    // warps per CTA: one per 32-row group up to 16; fewer when the line buffers would not fit
    // Groups are pipelined 33 steps apart, so a CTA with as many warps as groups keeps only about half of
    // them busy (fill and drain); fewer warps taking the groups round-robin stay busier and more images
    // share an SM -- but few large images need the warps.  Aim at ~16 wavefront warps per SM over the batch
    // (measured: 512 slices of 512x512 -> 5 warps, 64 radiographs of 3000x3000 -> 16; scripts/bench_med.py).
    int maxw = std::max(4, std::min(16, (ws->sm_count * 16 + n_units - 1) / n_units));
    if (const char* e = getenv("PEEB_MED_WARPS")) maxw = std::max(1, std::min(16, atoi(e)));
    int nwarps = std::max(1, std::min(maxw, (h - 1 + 31) / 32));
    while (nwarps > 1 && med_layout(g, nwarps).total > (size_t)ws->max_smem_optin) nwarps /= 2;
    // few large images: two CTAs (a cluster) per image, the same ring of line buffers in each; up to 32 warps per image
    // when the lines are short enough
    // (measured, scripts/bench_med.py: 64 radiographs of 3000x3000 6.23 -> 5.26 ms; 16 images of 1024x1024 0.95 -> 0.98 ms:
    // only large images take it)
    int CS = (2 * n_units <= ws->sm_count && nwarps >= 8 && (long long)h * w >= (4ll << 20)) ? 2 : 1;
    if (const char* e = getenv("PEEB_MED_CLUSTER")) CS = atoi(e) == 2 ? 2 : 1;
    int allwarps = nwarps;
    if (CS == 2) {
        const int want = std::min(32, (h - 1 + 31) / 32);
        while (allwarps * 2 <= want && med_layout(g, allwarps * 2).total <= (size_t)ws->max_smem_optin) allwarps *= 2;
        if (allwarps < 2) CS = 1; else nwarps = allwarps / 2;
    }
    const size_t smem = med_layout(g, CS * nwarps).total;
    PEEB_REQUIRE(smem <= (size_t)ws->max_smem_optin, "peeb_pee_med_extract_batch: image %dx%d needs more shared memory than one SM has", h, w);
    // vector path: every row of every unit starts on a 16-byte (8-bit pixels: 8-byte) boundary
    const uintptr_t al = 8 * (uintptr_t)itemsize - 1;
    const bool vec = (w % 8 == 0) && ((((uintptr_t)marked) | (uintptr_t)recovered | (uintptr_t)marked_stride | (uintptr_t)recovered_stride) & al) == 0 &&
                     !getenv("PEEB_MED_SCALAR");
    ProfScope p(ws, PEEB_K_PEE_EXTRACT, st);
#define PEEB_MED_LAUNCH(PIXT, VEC)                                                                                          \
    do {                                                                                                                    \
        PEEB_CUDA(cudaFuncSetAttribute(med_extract_kernel<PIXT, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cudaLaunchConfig_t cfg{};                                                                                           \
        cfg.gridDim = dim3((unsigned)(n_units * CS)); cfg.blockDim = dim3((unsigned)(nwarps * 32));                         \
        cfg.dynamicSmemBytes = smem; cfg.stream = st;                                                                       \
        cudaLaunchAttribute attr{};                                                                                         \
        attr.id = cudaLaunchAttributeClusterDimension;                                                                      \
        attr.val.clusterDim.x = (unsigned)CS; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;                         \
        cfg.attrs = &attr; cfg.numAttrs = 1;                                                                                \
        PEEB_CUDA(cudaLaunchKernelEx(&cfg, med_extract_kernel<PIXT, VEC>, g, bt, (unsigned*)ws->pbits[0].ptr, CS));         \
    } while (0)
    if (itemsize == 2) { if (vec) PEEB_MED_LAUNCH(unsigned short, true); else PEEB_MED_LAUNCH(unsigned short, false); }
    else { if (vec) PEEB_MED_LAUNCH(unsigned char, true); else PEEB_MED_LAUNCH(unsigned char, false); }
#undef PEEB_MED_LAUNCH
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

// ---- host-buffer variants: the whole batch is staged at once on the workspace stream (synchronous)
static int peeb_pee_med_embed_h_impl(peeb_ws* ws, const void* src_host, int n_units, int h, int w, int itemsize, int bit_depth,
                         const int32_t* T, const int64_t* n_bits, const uint8_t* payload_host, int64_t payload_stride,
                         void* marked_host, uint8_t* lm_host, int64_t* info_host) {
    PEEB_REQUIRE(ws && src_host && n_bits && info_host, "peeb_pee_med_embed_h: null pointer");
    PEEB_REQUIRE(n_units >= 1 && h >= 1 && w >= 1 && (itemsize == 1 || itemsize == 2) && payload_stride >= 0, "peeb_pee_med_embed_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = ws->stream;
    const size_t img = (size_t)h * w * itemsize, img_al = align_up(img, 256), lmb = (size_t)h * ((w + 7) / 8), lm_al = align_up(lmb, 256);
    int64_t maxpb = 0;
    for (int u = 0; u < n_units; ++u) {
        PEEB_REQUIRE(n_bits[u] >= 0 && (n_bits[u] + 7) / 8 <= payload_stride, "peeb_pee_med_embed_h: payload %d shorter than n_bits", u);
        maxpb = std::max<int64_t>(maxpb, (int64_t)peeb_payload_bytes(n_bits[u]));
    }
    const size_t pstride = align_up((size_t)std::max<int64_t>(payload_stride, maxpb), 16);
    int rc = scratch_reserve(ws->stage, 2 * (size_t)n_units * img_al + 256); if (rc) return rc;
    const size_t o_lm = (size_t)n_units * pstride, o_info = o_lm + (size_t)n_units * lm_al;
    rc = scratch_reserve(ws->stage2, o_info + (size_t)n_units * PEEB_INFO * 8 + 256); if (rc) return rc;
    char* d1 = (char*)ws->stage.ptr; char* d2 = (char*)ws->stage2.ptr;
    PEEB_CUDA(cudaMemcpy2DAsync(d1, img_al, src_host, img, img, n_units, cudaMemcpyHostToDevice, st));
    PEEB_CUDA(cudaMemsetAsync(d2, 0, o_lm, st));
    if (payload_stride > 0 && payload_host)
        PEEB_CUDA(cudaMemcpy2DAsync(d2, pstride, payload_host, (size_t)payload_stride, (size_t)payload_stride, n_units, cudaMemcpyHostToDevice, st));
    char* dm = d1 + (size_t)n_units * img_al;
    rc = peeb_pee_med_embed_batch(ws, d1, (int64_t)img_al, n_units, h, w, itemsize, bit_depth, T, n_bits, (const uint8_t*)d2, (int64_t)pstride,
                                  dm, (int64_t)img_al, (uint8_t*)(d2 + o_lm), (int64_t)lm_al, (int64_t*)(d2 + o_info), st);
    if (rc) return rc;
    if (marked_host) PEEB_CUDA(cudaMemcpy2DAsync(marked_host, img, dm, img_al, img, n_units, cudaMemcpyDeviceToHost, st));
    if (lm_host) PEEB_CUDA(cudaMemcpy2DAsync(lm_host, lmb, d2 + o_lm, lm_al, lmb, n_units, cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaMemcpyAsync(info_host, d2 + o_info, (size_t)n_units * PEEB_INFO * 8, cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaStreamSynchronize(st));
    return PEEB_OK;
}

int peeb_pee_med_embed_h(peeb_ws* ws, const void* src_host, int n_units, int h, int w, int itemsize, int bit_depth,
                         const int32_t* T, const int64_t* n_bits, const uint8_t* payload_host, int64_t payload_stride,
                         void* marked_host, uint8_t* lm_host, int64_t* info_host) {
    // a failed call returns only after the copies of earlier chunks have stopped touching the caller's buffers
    const int rc = peeb_pee_med_embed_h_impl(ws, src_host, n_units, h, w, itemsize, bit_depth, T, n_bits, payload_host, payload_stride, marked_host, lm_host, info_host);
    if (rc != PEEB_OK && ws) {
        cudaStreamSynchronize(ws->stream); cudaStreamSynchronize(ws->stream2); cudaStreamSynchronize(ws->stream3);
    }
    return rc;
}


static int peeb_pee_med_extract_h_impl(peeb_ws* ws, const void* marked_host, int n_units, int h, int w, int itemsize, int bit_depth,
                           const int32_t* T, const int64_t* n_bits, const uint8_t* lm_host, uint8_t* payload_out_host,
                           int64_t payload_stride, void* recovered_host, int64_t* info_host) {
    PEEB_REQUIRE(ws && marked_host && T && n_bits && lm_host && payload_out_host && info_host, "peeb_pee_med_extract_h: null pointer");
    PEEB_REQUIRE(n_units >= 1 && h >= 1 && w >= 1 && (itemsize == 1 || itemsize == 2), "peeb_pee_med_extract_h: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = ws->stream;
    const size_t img = (size_t)h * w * itemsize, img_al = align_up(img, 256), lmb = (size_t)h * ((w + 7) / 8), lm_al = align_up(lmb, 256);
    int64_t maxpb = 0;
    for (int u = 0; u < n_units; ++u) {
        PEEB_REQUIRE(n_bits[u] >= 0 && (n_bits[u] + 7) / 8 <= payload_stride, "peeb_pee_med_extract_h: payload_out %d shorter than n_bits", u);
        maxpb = std::max<int64_t>(maxpb, (int64_t)peeb_payload_bytes(n_bits[u]));
    }
    const size_t pstride = align_up((size_t)std::max<int64_t>(payload_stride, maxpb), 16);
    int rc = scratch_reserve(ws->stage, 2 * (size_t)n_units * img_al + 256); if (rc) return rc;
    const size_t o_lm = (size_t)n_units * pstride, o_info = o_lm + (size_t)n_units * lm_al;
    rc = scratch_reserve(ws->stage2, o_info + (size_t)n_units * PEEB_INFO * 8 + 256); if (rc) return rc;
    char* d1 = (char*)ws->stage.ptr; char* d2 = (char*)ws->stage2.ptr;
    PEEB_CUDA(cudaMemcpy2DAsync(d1, img_al, marked_host, img, img, n_units, cudaMemcpyHostToDevice, st));
    PEEB_CUDA(cudaMemcpy2DAsync(d2 + o_lm, lm_al, lm_host, lmb, lmb, n_units, cudaMemcpyHostToDevice, st));
    char* dr = d1 + (size_t)n_units * img_al;
    rc = peeb_pee_med_extract_batch(ws, d1, (int64_t)img_al, n_units, h, w, itemsize, bit_depth, T, n_bits, (const uint8_t*)(d2 + o_lm), (int64_t)lm_al,
                                    (uint8_t*)d2, (int64_t)pstride, recovered_host ? dr : nullptr, (int64_t)img_al, (int64_t*)(d2 + o_info), st);
    if (rc) return rc;
    if (recovered_host) PEEB_CUDA(cudaMemcpy2DAsync(recovered_host, img, dr, img_al, img, n_units, cudaMemcpyDeviceToHost, st));
    if (payload_stride > 0)
        PEEB_CUDA(cudaMemcpy2DAsync(payload_out_host, (size_t)payload_stride, d2, pstride, (size_t)payload_stride, n_units, cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaMemcpyAsync(info_host, d2 + o_info, (size_t)n_units * PEEB_INFO * 8, cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaStreamSynchronize(st));
    return PEEB_OK;
}

int peeb_pee_med_extract_h(peeb_ws* ws, const void* marked_host, int n_units, int h, int w, int itemsize, int bit_depth,
                           const int32_t* T, const int64_t* n_bits, const uint8_t* lm_host, uint8_t* payload_out_host,
                           int64_t payload_stride, void* recovered_host, int64_t* info_host) {
    // a failed call returns only after the copies of earlier chunks have stopped touching the caller's buffers
    const int rc = peeb_pee_med_extract_h_impl(ws, marked_host, n_units, h, w, itemsize, bit_depth, T, n_bits, lm_host, payload_out_host, payload_stride, recovered_host, info_host);
    if (rc != PEEB_OK && ws) {
        cudaStreamSynchronize(ws->stream); cudaStreamSynchronize(ws->stream2); cudaStreamSynchronize(ws->stream3);
    }
    return rc;
}


}  // extern "C"
