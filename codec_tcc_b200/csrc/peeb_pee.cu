// Row a10: reversible Prediction-Error Expansion on sm_100a.
//
// Work decomposition
//   unit  = one image of a batch (units are independent);
//   band  = R consecutive rows of a unit, full width.  One CTA owns one band: it
//           stages rows [r0-2, r0+R+2) in shared memory (one TMA bulk copy when the
//           row size allows), runs both colour passes on chip and writes the band
//           back once.  Two halo rows on each side let the CTA recompute, locally,
//           the colour-0 result of the rows next to the band that its colour-1
//           pixels predict from -- no second trip through global memory;
//   strip = 128 columns of a band; a warp walks a strip downwards, 4 pixels per lane
//           (one 64-bit / 32-bit shared load per row), so N/S neighbours are the
//           previous / next loads and W/E neighbours come from the lane's own word or
//           one shuffle.
//
// Payload bit index of a carrier = exclusive prefix count of carriers in raster order.
//   - per (row, strip) counts come from warp ballots, a block scan orders them;
//   - across bands: pass 0 uses the per-band counts of pee_count_kernel (which also
//     yields cap0, the base of pass 1); pass 1 uses a decoupled look-back over the
//     earlier bands of the same unit.  Bands are handed out by an atomic ticket, so a
//     band only ever waits on bands that are already running.
// Extraction needs no inter-band ordering for the pixels; every band writes its
// carrier bits to a staging area and pee_gather_kernel concatenates them.
#include <algorithm>
#include <cstdlib>

#include "peeb_common.cuh"
#include "peeb_pee.cuh"

namespace peeb {


// ------------------------------------------------------------------ pixels
// Four neighbouring pixels of one lane, kept packed in registers (one 64-bit / 32-bit
// shared-memory word); fields are extracted on demand.
template <typename PixT> struct Row4;
template <> struct Row4<unsigned short> {
    uint2 v;
    static constexpr int ITEM = 2;
    __device__ __forceinline__ void load(const unsigned char* p) { v = *reinterpret_cast<const uint2*>(p); }
    __device__ __forceinline__ void store(unsigned char* p) const { *reinterpret_cast<uint2*>(p) = v; }
    // Note (round 1 experiment): the kernels are bound by the ALU pipe (LOP3/SHF/ISETP/SEL, ncu ~70-80 %
    // busy) while the FMA pipe idles (~20 %).  Extracting halfwords with IMAD.HI / IMAD instead
    // (-DPEEB_FMA_EXTRACT) moves 10 ops per row step to the FMA pipe but needs two instructions
    // for the low half and was not faster on B200 (0.394 vs 0.389 ms), so LOP3/SHF stays.
    template <int K> __device__ __forceinline__ int f() const {
#ifndef PEEB_FMA_EXTRACT
        return K == 0 ? (int)(v.x & 0xffffu) : K == 1 ? (int)(v.x >> 16) : K == 2 ? (int)(v.y & 0xffffu) : (int)(v.y >> 16);
#else
        const unsigned wv = K < 2 ? v.x : v.y;
        const unsigned hi = __umulhi(wv, 0x10000u);
        return (K & 1) ? (int)hi : (int)(wv + hi * 0xffff0000u);
#endif
    }
    template <int K> __device__ __forceinline__ void set(int a) {
        if (K == 0) v.x = __byte_perm(v.x, (unsigned)a, 0x3254);
        else if (K == 1) v.x = __byte_perm(v.x, (unsigned)a, 0x5410);
        else if (K == 2) v.y = __byte_perm(v.y, (unsigned)a, 0x3254);
        else v.y = __byte_perm(v.y, (unsigned)a, 0x5410);
    }
    __device__ static __forceinline__ int load1(const unsigned char* row, int c) {
        return *reinterpret_cast<const unsigned short*>(row + 2 * c);
    }
};
template <> struct Row4<unsigned char> {
    unsigned v;
    static constexpr int ITEM = 1;
    __device__ __forceinline__ void load(const unsigned char* p) { v = *reinterpret_cast<const unsigned*>(p); }
    __device__ __forceinline__ void store(unsigned char* p) const { *reinterpret_cast<unsigned*>(p) = v; }
    template <int K> __device__ __forceinline__ int f() const { return (int)((v >> (8 * K)) & 0xffu); }
    template <int K> __device__ __forceinline__ void set(int a) {
        // byte K <- a (a <= 255)
        constexpr unsigned sel = K == 0 ? 0x3214u : K == 1 ? 0x3240u : K == 2 ? 0x3410u : 0x4210u;
        v = __byte_perm(v, (unsigned)a, sel);
    }
    __device__ static __forceinline__ int load1(const unsigned char* row, int c) { return row[c]; }
};

// Appendix A, embed side, for one pixel: value with a zero payload bit, whether
// it carries a bit.
__device__ __forceinline__ void classify_embed(int x, int p, int T, int maxval, int& nv0, bool& carrier) {
    const int e = x - p;
    const int t = e + T;
    const int v = x + e;                                            // p + 2e
    const bool expd = (unsigned)t < (unsigned)(2 * T);              // -T <= e < T
    carrier = expd && (unsigned)v < (unsigned)maxval;               // 0 <= v and v + 1 <= maxval
    const int xs = x + (t < 0 ? -T : T);                            // e < -T : e >= T (when not expandable)
    const bool shifted = !expd && (unsigned)xs <= (unsigned)maxval; // x-T >= 0 / x+T <= maxval
    nv0 = carrier ? v : (shifted ? xs : x);
}
// A pixel goes to the location map when it could neither be expanded nor shifted.  Shifts move by
// T >= 1, so "value unchanged and not a carrier" says exactly that (a carrier with e = 0 also keeps
// its value); deriving the flag this way keeps it out of the common path.
// (T == 0 marks a border column, which is never flagged.)
__device__ __forceinline__ bool flagged_embed(int x, int nv0, bool carrier, int T) {
    return nv0 == x && !carrier && T != 0;
}

// Appendix A, extract side.
__device__ __forceinline__ void classify_extract(int x, int p, int T, bool lmflag, int& orig, bool& carrier,
                                                 int& bit) {
    const int ee = x - p;
    const bool car = (unsigned)(ee + 2 * T) < (unsigned)(4 * T);  // -2T <= e' < 2T
    const int e = car ? (ee >> 1) : (ee >= 2 * T ? ee - T : ee + T);
    bit = ee & 1;
    carrier = car && !lmflag;
    orig = lmflag ? x : p + e;
}

// 64 payload bits starting at bit K0 of an MSB-first packed stream; bit K0 is the
// MSB of `hi`.  The stream must be readable 12 bytes past the word holding K0.
__device__ __forceinline__ void payload_window(const unsigned* __restrict__ pay, unsigned K0, unsigned& hi,
                                               unsigned& lo) {
    const unsigned wi = K0 >> 5, sh = K0 & 31;
    const unsigned w0 = __byte_perm(__ldg(pay + wi), 0, 0x0123);
    const unsigned w1 = __byte_perm(__ldg(pay + wi + 1), 0, 0x0123);
    const unsigned w2 = __byte_perm(__ldg(pay + wi + 2), 0, 0x0123);
    hi = __funnelshift_l(w1, w0, sh);
    lo = __funnelshift_l(w2, w1, sh);
}
__device__ __forceinline__ int window_bit(unsigned hi, unsigned lo, int k) {  // 0 <= k < 64
    return (int)(((k < 32 ? hi : lo) >> (31 - (k & 31))) & 1u);
}


// ------------------------------------------------------------------ shared layout
struct SmemLayout {
    size_t img, lm, tab, bits, misc, bar, xb, pk, tab0, stream, total;
};
__host__ __device__ inline SmemLayout band_layout(const PeeGeom& g, int kind /*0 count, 1 embed, 2 extract*/) {
    SmemLayout L{};
    size_t o = 0;
    L.img = o; o += align_up((size_t)(g.R + 4) * g.pitch + 512, 16);
    L.lm = o; o += align_up((size_t)(g.R + 2) * g.lmpitch, 16);
    L.tab = o; o += align_up((size_t)(g.R + 2) * g.S * sizeof(int), 16);
    L.misc = o; o += 64 * sizeof(int);
    L.bar = o; o += 16;
    L.bits = L.xb = L.pk = L.tab0 = L.stream = o;
    if (kind == 1) {
        L.bits = o; o += align_up((size_t)(g.R + 2) * g.S * 64 + 64, 16);           // payload bits of the band, one byte each
    } else if (kind == 2) {
        L.xb = o; o += align_up((size_t)g.R * g.S * 64, 16);
        L.pk = o; o += align_up((size_t)2 * g.R * g.S * sizeof(unsigned long long), 16);
        L.tab0 = o; o += align_up((size_t)g.R * g.S * sizeof(int), 16);
        L.stream = o; o += align_up((size_t)2 * g.bandwords * sizeof(unsigned), 16);
    }
    L.total = o;
    return L;
}


// ------------------------------------------------------------------ the row walk
// A warp item = (strip, chunk of rows).  The warp walks its strip downwards keeping
// three packed rows in registers; the body sees row i as (up, mid, down) and the
// compile-time column parity Q of the colour being processed (colour pixels sit at
// lane columns Q and Q+2).  Rows are taken two at a time so that the parity and the
// register roles are static inside the loop body.
struct ItemCtx {
    int s;        // strip
    int c0;       // first column of this lane
    int lane;
    unsigned lt;  // lanemask_lt
    bool va[2], vb[2];  // validity (interior column) of pixel A / B for Q = 0 / 1
    bool in_image;      // c0 < w: this lane's word holds image columns and may be written back
    int Ta[2], Tb[2];   // the threshold for pixel A / B, 0 where the column is not interior: with T = 0
                        // a pixel is never expandable and its shift is by 0, i.e. it is left alone
};

// rhombus predictions for the two colour pixels of this lane in row `mid`
template <typename PixT, int Q>
__device__ __forceinline__ void predict_pair(const ItemCtx& c, const Row4<PixT>& U, const Row4<PixT>& M,
                                             const Row4<PixT>& D, const unsigned char* midp, int w, int& xa,
                                             int& pa, int& xb, int& pb) {
    if (Q == 0) {
        int left = __shfl_up_sync(0xffffffffu, M.template f<3>(), 1);
        if (c.lane == 0) left = c.c0 > 0 ? Row4<PixT>::load1(midp, -1) : 0;
        const int m1 = M.template f<1>();
        xa = M.template f<0>(); pa = (U.template f<0>() + D.template f<0>() + left + m1) >> 2;
        xb = M.template f<2>(); pb = (U.template f<2>() + D.template f<2>() + m1 + M.template f<3>()) >> 2;
    } else {
        int right = __shfl_down_sync(0xffffffffu, M.template f<0>(), 1);
        if (c.lane == 31) right = c.c0 + 4 < w ? Row4<PixT>::load1(midp, 4) : 0;
        const int m2 = M.template f<2>();
        xa = M.template f<1>(); pa = (U.template f<1>() + D.template f<1>() + M.template f<0>() + m2) >> 2;
        xb = M.template f<3>(); pb = (U.template f<3>() + D.template f<3>() + m2 + right) >> 2;
    }
}

template <typename PixT, class Body>
__device__ __forceinline__ void walk_rows(const PeeGeom& g, unsigned char* simg, int r_first, int colour, int ra,
                                          int rb, const ItemCtx& c, Body& body) {
    if (ra >= rb) return;
    unsigned char* p = simg + (size_t)(ra - 1 - r_first) * g.pitch + (size_t)c.c0 * Row4<PixT>::ITEM;  // row ra-1
    const int pitch = g.pitch;
    Row4<PixT> U, M, D0, D1;
    U.load(p);
    M.load(p + pitch);
    p += pitch;  // p -> row `i` (mid) at this lane's columns
    int i = ra;
    if ((i + colour) & 1) {  // peel one row so that the loop starts on parity 0
        D0.load(p + pitch);
        body.template step<1>(i, U, M, D0, p);
        U = M; M = D0; p += pitch; ++i;
    }
    for (; i + 1 < rb; i += 2) {
        D0.load(p + pitch);
        body.template step<0>(i, U, M, D0, p);
        D1.load(p + 2 * pitch);
        body.template step<1>(i + 1, M, D0, D1, p + pitch);
        U = D0; M = D1; p += 2 * pitch;
    }
    if (i < rb) {
        D0.load(p + pitch);
        body.template step<0>(i, U, M, D0, p);
    }
}

// Splits rows [row_lo, row_hi) x strips into warp items and walks them.
template <typename PixT, class Body>
__device__ __forceinline__ void sweep(const PeeGeom& g, unsigned char* simg, int r_first, int colour, int row_lo,
                                      int row_hi, int T, Body& body) {
    const int nrows = row_hi - row_lo;
    if (nrows <= 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int G = max(1, nwarps / g.S);  // row chunks per strip
    const int RC = (nrows + G - 1) / G;  // rows per chunk
    const int nitems = g.S * G;
    ItemCtx c;
    c.lane = lane;
    c.lt = lanemask_lt();
    for (int item = warp; item < nitems; item += nwarps) {
        const int chunk = item / g.S;
        c.s = item - chunk * g.S;
        c.c0 = c.s * STRIP + 4 * lane;
        c.in_image = c.c0 < g.w;
        const int ra = row_lo + chunk * RC, rb = min(ra + RC, row_hi);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            c.va[q] = c.c0 + q >= 1 && c.c0 + q <= g.w - 2;
            c.vb[q] = c.c0 + q + 2 <= g.w - 2;
            c.Ta[q] = c.va[q] ? T : 0;
            c.Tb[q] = c.vb[q] ? T : 0;
            // keep them as register values: otherwise they are re-derived from c0 and w in every row step
            asm volatile("" : "+r"(c.Ta[q]), "+r"(c.Tb[q]));
        }
        body.begin_item(c, ra);
        walk_rows<PixT>(g, simg, r_first, colour, ra, rb, c, body);
    }
}

struct EmbedStats {
    unsigned long long sse = 0;
    unsigned flagged = 0;
};

// ---- body: count carriers of one colour per (row, strip) ---------------------------------
// GLOBAL=true  (pee_count_kernel): byte counts to rowcnt[row*S + strip] in global memory;
// GLOBAL=false (pass 1 of the embed kernel): counts to the shared table.
// Only the carrier predicate is live here, so the sweep is about a third of a full apply.
template <typename PixT, bool GLOBAL>
struct CountBody {
    const PeeGeom& g;
    const ItemCtx* c;
    int row0;                    // image row of table row 0
    unsigned char* rowcnt;       // GLOBAL
    int* tab;                    // !GLOBAL
    int total;                   // carriers seen by this warp (lane 0)
    int idx;                     // running table index of (row, strip)
    __device__ __forceinline__ void begin_item(const ItemCtx& ctx, int ra) { c = &ctx; idx = (ra - row0) * g.S + ctx.s; }
    template <int Q>
    __device__ __forceinline__ void step(int i, const Row4<PixT>& U, Row4<PixT>& M, const Row4<PixT>& D,
                                         unsigned char* midp) {
        int xa, pa, xb, pb, na, nb;
        bool cara, carb;
        predict_pair<PixT, Q>(*c, U, M, D, midp, g.w, xa, pa, xb, pb);
        classify_embed(xa, pa, c->Ta[Q], g.maxval, na, cara);
        classify_embed(xb, pb, c->Tb[Q], g.maxval, nb, carb);
        const unsigned ma = __ballot_sync(0xffffffffu, cara), mb = __ballot_sync(0xffffffffu, carb);
        if (c->lane == 0) {
            const int n = __popc(ma) + __popc(mb);
            if (GLOBAL) { rowcnt[idx] = (unsigned char)n; total += n; }
            else tab[idx] = n;
        }
        idx += g.S;
        (void)i;
    }
};

// ---- body: full apply of one colour (pass 0 of the embed kernel) --------------------------
// tab[(i-row0)*S + strip] = index into `bits` (one byte per payload bit of this band, zero
// padded) of the first carrier of that (row, strip).
template <typename PixT, bool ALLOWN>
struct ApplyBody {
    const PeeGeom& g;
    const ItemCtx* c;
    int row0, own_lo, own_hi;
    const int* tab;
    const unsigned char* bits;
    unsigned* slm; int lm_row0;
    EmbedStats* st;
    int idx;
    __device__ __forceinline__ void begin_item(const ItemCtx& ctx, int ra) { c = &ctx; idx = (ra - row0) * g.S + ctx.s; }
    template <int Q>
    __device__ __forceinline__ void step(int i, const Row4<PixT>& U, Row4<PixT>& M, const Row4<PixT>& D,
                                         unsigned char* midp) {
        int xa, pa, xb, pb, na, nb;
        bool cara, carb;
        predict_pair<PixT, Q>(*c, U, M, D, midp, g.w, xa, pa, xb, pb);
        classify_embed(xa, pa, c->Ta[Q], g.maxval, na, cara);
        classify_embed(xb, pb, c->Tb[Q], g.maxval, nb, carb);
        const unsigned ma = __ballot_sync(0xffffffffu, cara), mb = __ballot_sync(0xffffffffu, carb);
        const unsigned char* bp = bits + tab[idx] + __popc(ma & c->lt) + __popc(mb & c->lt);
        idx += g.S;
        const bool fla = flagged_embed(xa, na, cara, c->Ta[Q]), flb = flagged_embed(xb, nb, carb, c->Tb[Q]);
        if (cara) { na += bp[0]; ++bp; }
        if (carb) nb += bp[0];
        if (ALLOWN || (i >= own_lo && i < own_hi)) {
            const int da = na - xa, db = nb - xb;
            st->sse += (unsigned long long)((unsigned)(da * da) + (unsigned)(db * db));
            if (fla | flb) {  // rare; only interior columns may enter the map
                unsigned* lrow = slm + (size_t)(i - lm_row0) * (g.lmpitch >> 2);
                const int ca = c->c0 + Q, cb = ca + 2;
                if (fla) { atomicOr(lrow + (ca >> 5), lm_bitmask(ca)); ++st->flagged; }
                if (flb) { atomicOr(lrow + (cb >> 5), lm_bitmask(cb)); ++st->flagged; }
            }
        }
        M.template set<Q>(na);
        M.template set<Q + 2>(nb);
        if (c->in_image) M.store(midp);
    }
};

// ------------------------------------------------------------------ K_A: pass-0 counts
// grid = n_units * nb.  rowcnt[(unit*h + row)*S + strip] = pass-0 carriers of that row
// segment (<= 64, one byte); band_cnt[unit*nb + band] = their sum over the band's own rows;
// info[unit][3] (cap0) accumulates the unit total.
template <typename PixT, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) pee_count_kernel(PeeGeom g, PeeBatch bt, int* __restrict__ band_cnt,
                                                          unsigned char* __restrict__ rowcnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = band_layout(g, 0);
    unsigned char* simg = smem_raw + L.img;
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    const int unit = blockIdx.x / g.nb, band = blockIdx.x % g.nb;
    if (threadIdx.x == 0) { misc[0] = 0; if (g.bulk) { mbar_init(bar, 1); fence_mbar_init(); } }
    __syncthreads();
    const int r0 = band * g.R, r_first = r0 - 2;
    const unsigned char* usrc = bt.src + (long long)unit * bt.src_stride;
    load_rows<PixT>(g, usrc, simg, r_first, max(r0 - 1, 0), min(r0 + g.R + 1, g.h), bar);
    const int own_lo = max(r0, 1), own_hi = min(r0 + g.R, g.h - 1);
    CountBody<PixT, true> body{g, nullptr, 0, rowcnt + (long long)unit * g.h * g.S, nullptr, 0, 0};
    sweep<PixT>(g, simg, r_first, 0, own_lo, own_hi, bt.T[unit], body);
    if ((threadIdx.x & 31) == 0 && body.total) atomicAdd(misc, body.total);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tot = misc[0];
        band_cnt[unit * g.nb + band] = tot;
        if (tot) atomicAdd(reinterpret_cast<unsigned long long*>(bt.info + (long long)unit * PEEB_INFO + 3), (unsigned long long)tot);
    }
}

// ------------------------------------------------------------------ K_B: fused two-pass embed

template <typename PixT, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) pee_embed_kernel(PeeGeom g, PeeBatch bt, const int* __restrict__ band_cnt,
                                                          const unsigned char* __restrict__ rowcnt,
                                                          unsigned* __restrict__ ticket,
                                                          unsigned long long* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = band_layout(g, 1);
    unsigned char* simg = smem_raw + L.img;
    unsigned* slm = reinterpret_cast<unsigned*>(smem_raw + L.lm);
    int* tab = reinterpret_cast<int*>(smem_raw + L.tab);
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    unsigned char* bits = smem_raw + L.bits;

    // in-order ticket: a band only ever waits on bands with smaller tickets
    if (threadIdx.x == 0) {
        misc[40] = (int)atomicAdd(ticket, 1u);
        if (g.bulk) { mbar_init(bar, 1); fence_mbar_init(); }
    }
    __syncthreads();
    const int tk = misc[40];
    const int unit = tk / g.nb, band = tk - unit * g.nb;
    const int r0 = band * g.R, r_first = r0 - 2;
    const unsigned char* usrc = bt.src + (long long)unit * bt.src_stride;
    const int T = bt.T[unit];
    const unsigned n_bits = bt.n_bits[unit];
    const unsigned* payload = reinterpret_cast<const unsigned*>(bt.payload + (long long)unit * bt.payload_stride);
    long long* info = bt.info + (long long)unit * PEEB_INFO;
    const int own_lo = max(r0, 1), own_hi = min(r0 + g.R, g.h - 1);
    const int p0_lo = max(r0 - 1, 1), p0_hi = min(r0 + g.R + 1, g.h - 1);
    const int n0 = max(p0_hi - p0_lo, 0) * g.S, n1 = max(own_hi - own_lo, 0) * g.S;

    load_rows<PixT>(g, usrc, simg, r_first, max(r0 - 2, 0), min(r0 + g.R + 2, g.h), bar);
    for (int k = threadIdx.x; k < (g.R * g.lmpitch) >> 2; k += blockDim.x) slm[k] = 0;
    // pass-0 counts of the rows this band touches (own rows + one halo row each side)
    {
        const unsigned char* rc = rowcnt + ((long long)unit * g.h + p0_lo) * g.S;
        for (int k = threadIdx.x; k < n0; k += blockDim.x) tab[k] = rc[k];
    }
    // pass-0 prefix of this band and cap0 of the unit from the count kernel
    if (threadIdx.x < 32) {
        int before = 0, all = 0;
        for (int k = threadIdx.x; k < g.nb; k += 32) {
            const int cc = band_cnt[unit * g.nb + k];
            all += cc;
            if (k < band) before += cc;
        }
        before = (int)warp_sum_i64(before);
        all = (int)warp_sum_i64(all);
        if (threadIdx.x == 0) { misc[41] = before; misc[42] = all; }
    }
    __syncthreads();
    EmbedStats st;

    // ---- pass 0 (colour 0): band rows and one halo row on each side, one full sweep
    {
        // carriers of the halo row above precede this band in raster order
        int halo_top = 0;
        if (p0_lo < own_lo) for (int k = 0; k < g.S; ++k) halo_top += tab[k];
        __syncthreads();
        const int total0 = block_excl_scan(tab, n0, misc);   // tab: index into `bits`
        expand_payload(payload, (unsigned)(misc[41] - halo_top), total0, n_bits, bits);
        __syncthreads();
        ApplyBody<PixT, false> body{g, nullptr, p0_lo, own_lo, own_hi, tab, bits, slm, r0, &st, 0};
        sweep<PixT>(g, simg, r_first, 0, p0_lo, p0_hi, T, body);
    }
    __syncthreads();

    // ---- pass 1 (colour 1) over the band rows: a cheap counting sweep orders the carriers,
    //      the same full apply as pass 0 then rewrites the pixels
    {
        CountBody<PixT, false> body{g, nullptr, own_lo, nullptr, tab, 0, 0};
        sweep<PixT>(g, simg, r_first, 1, own_lo, own_hi, T, body);
    }
    __syncthreads();
    {
        const int total = block_excl_scan(tab, n1, misc);
        if (threadIdx.x == 0) {
            unsigned long long* stt = status + (long long)unit * g.nb;
            atomicExch(stt + band, ST_AGG | (unsigned)total);
            unsigned before = 0;
            for (int k = band - 1; k >= 0; --k) {
                unsigned long long v;
                do { v = *reinterpret_cast<volatile unsigned long long*>(stt + k); } while ((v & ST_MASK) == 0);
                before += (unsigned)(v & 0xffffffffu);
                if ((v & ST_MASK) == ST_PFX) break;
            }
            atomicExch(stt + band, ST_PFX | (unsigned long long)(before + (unsigned)total));
            misc[43] = (int)before;
            if (total) atomicAdd(reinterpret_cast<unsigned long long*>(info + 4), (unsigned long long)total);
        }
        __syncthreads();
        expand_payload(payload, (unsigned)(misc[42] + misc[43]), total, n_bits, bits);  // from cap0 + earlier bands' pass-1 carriers
        __syncthreads();
        ApplyBody<PixT, true> body{g, nullptr, own_lo, own_lo, own_hi, tab, bits, slm, r0, &st, 0};
        sweep<PixT>(g, simg, r_first, 1, own_lo, own_hi, T, body);
    }

    // ---- statistics
    {
        const unsigned long long sse = warp_sum_u64(st.sse);
        const unsigned long long fl = warp_sum_u64(st.flagged);
        if ((threadIdx.x & 31) == 0) {
            if (sse) atomicAdd(reinterpret_cast<unsigned long long*>(info + 6), sse);
            if (fl) atomicAdd(reinterpret_cast<unsigned long long*>(info + 5), fl);
        }
    }
    // ---- write back the band (image rows, location-map rows)
    const int b_lo = r0, b_hi = min(r0 + g.R, g.h);
    if (bt.dst) store_rows<PixT>(g, bt.dst + (long long)unit * bt.dst_stride, simg, r_first, b_lo, b_hi);
    else __syncthreads();
    if (bt.lm) {
        unsigned char* glm = bt.lm + (long long)unit * bt.lm_stride + (size_t)b_lo * g.lmw;
        const int nrows = b_hi - b_lo;
        if ((g.lmw & 3) == 0 && ((uintptr_t)glm & 3) == 0) {
            const int wpr = g.lmw >> 2;
            for (int k = threadIdx.x; k < nrows * wpr; k += blockDim.x)
                reinterpret_cast<unsigned*>(glm)[k] = slm[(k / wpr) * (g.lmpitch >> 2) + (k % wpr)];
        } else {
            const unsigned char* sb = reinterpret_cast<const unsigned char*>(slm);
            for (int k = threadIdx.x; k < nrows * g.lmw; k += blockDim.x)
                glm[k] = sb[(k / g.lmw) * g.lmpitch + (k % g.lmw)];
        }
    }
}

__global__ void pee_finalize_kernel(PeeBatch bt, int extract) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= bt.n_units) return;
    long long* info = bt.info + (long long)u * PEEB_INFO;
    info[0] = bt.T[u];
    info[1] = bt.n_bits[u];
    info[2] = info[3] + info[4];
    info[7] = ((long long)bt.n_bits[u] > info[2]) ? PEEB_E_CAPACITY : 0;
    (void)extract;
}

// ------------------------------------------------------------------ K_X: extract
// One sweep per colour; pixels are restored in place, carrier bits of the band's own
// rows are compacted per (row, strip) with ballots and a warp OR-reduction.
template <typename PixT, bool ALLOWN>
struct ExtractBody {
    const PeeGeom& g;
    const ItemCtx* c;
    int own_lo, own_hi;
    const unsigned* slm; int lm_row0;
    int* cnt;                 // carriers per (row, strip) of the band's own rows
    unsigned char* xbytes;    // their payload bits, one byte each, 64-byte slot per (row, strip)
    int lmshift;              // nibble position of this lane's 4 columns inside its location-map word
    const unsigned* lmp;      // running pointer to this lane's location-map word of row i
    int idx;                  // running table index
    __device__ __forceinline__ void begin_item(const ItemCtx& ctx, int ra) {
        c = &ctx;
        lmshift = 8 * ((ctx.c0 >> 3) & 3) + ((ctx.c0 & 4) ? 0 : 4);
        lmp = slm + (size_t)(ra - lm_row0) * (g.lmpitch >> 2) + (ctx.c0 >> 5);
        idx = (ra - own_lo) * g.S + ctx.s;
    }
    template <int Q>
    __device__ __forceinline__ void step(int i, const Row4<PixT>& U, Row4<PixT>& M, const Row4<PixT>& D,
                                         unsigned char* midp) {
        int xa, pa, xb, pb, oa, ob, bita, bitb;
        bool cara, carb;
        predict_pair<PixT, Q>(*c, U, M, D, midp, g.w, xa, pa, xb, pb);
        const unsigned nib = *lmp >> lmshift;  // bit 3 = column c0 ... bit 0 = column c0 + 3
        lmp += g.lmpitch >> 2;
        classify_extract(xa, pa, c->Ta[Q], (nib >> (3 - Q)) & 1u, oa, cara, bita);
        classify_extract(xb, pb, c->Tb[Q], (nib >> (1 - Q)) & 1u, ob, carb, bitb);
        if (ALLOWN || (i >= own_lo && i < own_hi)) {
            const unsigned ma = __ballot_sync(0xffffffffu, cara), mb = __ballot_sync(0xffffffffu, carb);
            unsigned char* sp = xbytes + idx * 64 + __popc(ma & c->lt) + __popc(mb & c->lt);  // raster rank
            if (cara) { sp[0] = (unsigned char)bita; ++sp; }
            if (carb) sp[0] = (unsigned char)bitb;
            if (c->lane == 0) cnt[idx] = __popc(ma) + __popc(mb);
        }
        idx += g.S;
        M.template set<Q>(oa);
        M.template set<Q + 2>(ob);
        if (c->in_image) M.store(midp);
    }
};

// Packs the first `n` (<= 64) 0/1 bytes of a 64-byte slot into a 64-bit word, first byte at bit 0.
__device__ __forceinline__ unsigned long long pack_slot(const unsigned char* slot, int n) {
    const unsigned* w = reinterpret_cast<const unsigned*>(slot);
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // (b0 | b1<<8 | b2<<16 | b3<<24) * 0x10204080 >> 28 = b0 | b1<<1 | b2<<2 | b3<<3
        lo |= (((w[k] & 0x01010101u) * 0x10204080u) >> 28) << (4 * k);
        hi |= (((w[k + 8] & 0x01010101u) * 0x10204080u) >> 28) << (4 * k);
    }
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    if (n < 64) v &= (1ull << n) - 1ull;  // bytes past the count are stale
    return v;
}

// grid = n_units * nb (no inter-band dependency).  stage_bits: per (unit, pass, band)
// `bandwords` 32-bit words, carrier bit k at word k>>5, bit k&31; stage_cnt likewise.
template <typename PixT, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) pee_extract_kernel(PeeGeom g, PeeBatch bt, unsigned* __restrict__ stage_bits,
                                                            int* __restrict__ stage_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = band_layout(g, 2);
    unsigned char* simg = smem_raw + L.img;
    unsigned* slm = reinterpret_cast<unsigned*>(smem_raw + L.lm);
    int* cnt1 = reinterpret_cast<int*>(smem_raw + L.tab);   // colour 1 table
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    unsigned char* xbytes = smem_raw + L.xb;                   // 64-byte slot per (row, strip), reused by both colours
    unsigned long long* pk1 = reinterpret_cast<unsigned long long*>(smem_raw + L.pk);  // packed pieces, colour 1 then 0
    unsigned long long* pk0 = pk1 + g.R * g.S;
    int* cnt0 = reinterpret_cast<int*>(smem_raw + L.tab0);
    unsigned* stream = reinterpret_cast<unsigned*>(smem_raw + L.stream);  // [2][bandwords]: pass 0, pass 1

    const int unit = blockIdx.x / g.nb, band = blockIdx.x % g.nb;
    const int r0 = band * g.R, r_first = r0 - 2;
    if (g.bulk && threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    const unsigned char* usrc = bt.src + (long long)unit * bt.src_stride;
    const int T = bt.T[unit];
    load_rows<PixT>(g, usrc, simg, r_first, max(r0 - 2, 0), min(r0 + g.R + 2, g.h), bar);
    // location-map rows [r0-1, r0+R+1) -> slm (row pitch lmpitch)
    {
        const int l_lo = max(r0 - 1, 0), l_hi = min(r0 + g.R + 1, g.h);
        const unsigned char* glm = bt.lm + (long long)unit * bt.lm_stride;
        unsigned char* sb = reinterpret_cast<unsigned char*>(slm);
        if ((g.lmw & 3) == 0 && ((uintptr_t)glm & 3) == 0) {
            const int wpr = g.lmw >> 2;
            const unsigned* gw = reinterpret_cast<const unsigned*>(glm + (size_t)l_lo * g.lmw);
            for (int k = threadIdx.x; k < (l_hi - l_lo) * wpr; k += blockDim.x)
                slm[(size_t)(l_lo - (r0 - 1) + k / wpr) * (g.lmpitch >> 2) + (k % wpr)] = gw[k];
        } else {
            for (int k = threadIdx.x; k < (l_hi - l_lo) * g.lmw; k += blockDim.x)
                sb[(size_t)(l_lo - (r0 - 1) + k / g.lmw) * g.lmpitch + (k % g.lmw)] = glm[(size_t)l_lo * g.lmw + k];
        }
        for (int k = threadIdx.x; k < 2 * g.bandwords; k += blockDim.x) stream[k] = 0;
    }
    __syncthreads();
    const int own_lo = max(r0, 1), own_hi = min(r0 + g.R, g.h - 1);
    const int p1_lo = max(r0 - 1, 1), p1_hi = min(r0 + g.R + 1, g.h - 1);
    // colour 1 first (band rows + one halo row each side), then colour 0 (band rows)
    const int n = max(own_hi - own_lo, 0) * g.S;
    {
        ExtractBody<PixT, false> body{g, nullptr, own_lo, own_hi, slm, r0 - 1, cnt1, xbytes, 0, nullptr, 0};
        sweep<PixT>(g, simg, r_first, 1, p1_lo, p1_hi, T, body);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) pk1[k] = pack_slot(xbytes + k * 64, cnt1[k]);
    __syncthreads();
    {
        ExtractBody<PixT, true> body{g, nullptr, own_lo, own_hi, slm, r0 - 1, cnt0, xbytes, 0, nullptr, 0};
        sweep<PixT>(g, simg, r_first, 0, own_lo, own_hi, T, body);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) pk0[k] = pack_slot(xbytes + k * 64, cnt0[k]);
    __syncthreads();

    // concatenate the per-(row,strip) pieces into one bit stream per pass
    for (int pass = 0; pass < 2; ++pass) {
        int* cnt = pass == 0 ? cnt0 : cnt1;
        const unsigned long long* pk = pass == 0 ? pk0 : pk1;
        unsigned* out = stream + (size_t)pass * g.bandwords;
        // the scan turns counts into bit offsets; a piece's size is the next offset minus its own
        const int total = block_excl_scan(cnt, n, misc);
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int o = cnt[k];
            const int cc = (k + 1 < n ? cnt[k + 1] : total) - o;
            if (cc > 0) {
                const unsigned long long v = pk[k];
                const int wi = o >> 5, sh = o & 31;
                atomicOr(out + wi, (unsigned)(v << sh));
                const unsigned long long hi = sh ? (v >> (32 - sh)) : (v >> 32);
                if ((unsigned)hi) atomicOr(out + wi + 1, (unsigned)hi);
                if (sh && (unsigned)(hi >> 32)) atomicOr(out + wi + 2, (unsigned)(hi >> 32));
            }
        }
        __syncthreads();
        const long long slot = ((long long)unit * 2 + pass) * g.nb + band;
        if (threadIdx.x == 0) stage_cnt[slot] = total;
        unsigned* gout = stage_bits + slot * g.bandwords;
        const int nw = (total + 31) >> 5;
        for (int k = threadIdx.x; k < nw; k += blockDim.x) gout[k] = out[k];
        __syncthreads();
    }
    if (bt.dst)
        store_rows<PixT>(g, bt.dst + (long long)unit * bt.dst_stride, simg, r_first, r0, min(r0 + g.R, g.h));
}

// ------------------------------------------------------------------ K_G: payload assembly
// grid = (2*nb, n_units).  Piece p of a unit = pass-0 band p (p < nb) or pass-1 band
// p-nb; its global bit offset is the sum of the earlier pieces' counts.  Output is
// MSB-first packed, truncated to n_bits; payload_out is zeroed by the caller.
__global__ void __launch_bounds__(128) pee_gather_kernel(PeeGeom g, PeeBatch bt, const unsigned* __restrict__ stage_bits,
                                                         const int* __restrict__ stage_cnt) {
    const int unit = blockIdx.y, piece = blockIdx.x;
    const int* cnts = stage_cnt + (long long)unit * 2 * g.nb;
    long long before = 0, all = 0;
    for (int k = threadIdx.x; k < 2 * g.nb; k += blockDim.x) {
        const int c = cnts[k];
        all += c;
        if (k < piece) before += c;
    }
    before = warp_sum_i64(before);
    all = warp_sum_i64(all);
    __shared__ long long s_b[4], s_a[4];
    if ((threadIdx.x & 31) == 0) { s_b[threadIdx.x >> 5] = before; s_a[threadIdx.x >> 5] = all; }
    __syncthreads();
    before = s_b[0] + s_b[1] + s_b[2] + s_b[3];
    all = s_a[0] + s_a[1] + s_a[2] + s_a[3];
    const long long n_bits = bt.n_bits[unit];
    long long* info = bt.info + (long long)unit * PEEB_INFO;
    if (piece == 0 && threadIdx.x == 0) {
        long long c0 = 0;
        for (int k = 0; k < g.nb; ++k) c0 += cnts[k];
        info[0] = bt.T[unit]; info[1] = n_bits; info[2] = all; info[3] = c0; info[4] = all - c0;
        info[5] = 0; info[6] = 0; info[7] = n_bits > all ? PEEB_E_CAPACITY : 0;
    }
    const int cnt = cnts[piece];
    if (cnt == 0 || before >= n_bits) return;
    const unsigned* src = stage_bits + ((long long)unit * 2 * g.nb + piece) * g.bandwords;
    unsigned* out = reinterpret_cast<unsigned*>(bt.payload_out + (long long)unit * bt.payload_stride);
    const int nsrc = (cnt + 31) >> 5;
    const long long first = before >> 5, last = (before + cnt - 1) >> 5;
    const int sh = (int)(before & 31);
    for (long long mw = first + threadIdx.x; mw <= last; mw += blockDim.x) {
        const int i = (int)(mw - first);
        const unsigned cur = i < nsrc ? src[i] : 0u;
        const unsigned prev = (i >= 1 && i - 1 < nsrc) ? src[i - 1] : 0u;
        unsigned val = sh ? ((cur << sh) | (prev >> (32 - sh))) : cur;
        // drop bits at or past n_bits
        const long long bit0 = mw << 5;
        if (bit0 + 32 > n_bits) {
            const int keep = (int)(n_bits - bit0);
            val = keep <= 0 ? 0u : (val & (0xffffffffu >> (32 - keep)));
        }
        if (val == 0) continue;
        const unsigned packed = __byte_perm(__brev(val), 0, 0x0123);  // LSB-first word -> MSB-first bytes
        if (mw == first || mw == last) atomicOr(out + mw, packed);
        else out[mw] = packed;
    }
}

// ------------------------------------------------------------------ prediction-error histogram
// Appendix A threshold selection: hist[u][c][e + tmax] over interior pixels not flagged
// for expansion.  One pixel per thread; |e| < 512 goes through warp-aggregated
// shared-memory atomics, the rare rest straight to global memory.
constexpr int HWIN = 512;
template <typename PixT>
__global__ void __launch_bounds__(256) pee_hist_kernel(const unsigned char* __restrict__ src, long long src_stride,
                                                       int h, int w, int maxval, int tmax,
                                                       unsigned* __restrict__ hist) {
    __shared__ unsigned sh[2][2 * HWIN];
    for (int k = threadIdx.x; k < 4 * HWIN; k += blockDim.x) (&sh[0][0])[k] = 0;
    __syncthreads();
    const int unit = blockIdx.y;
    const PixT* img = reinterpret_cast<const PixT*>(src + (long long)unit * src_stride);
    unsigned* uh = hist + (long long)unit * 4 * tmax;
    const long long npx = (long long)(h - 2) * (w - 2);
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp0 * 32; base < npx; base += nwarps * 32) {  // warp-uniform trip count
        const long long t = base + lane;
        bool ok = t < npx;
        int e = 0, c = 0;
        if (ok) {
            const int i = 1 + (int)(t / (w - 2)), j = 1 + (int)(t % (w - 2));
            const long long at = (long long)i * w + j;
            const int x = img[at];
            const int p = ((int)img[at - w] + (int)img[at + w] + (int)img[at - 1] + (int)img[at + 1]) >> 2;
            e = x - p;
            const int v = x + e;
            c = (i + j) & 1;
            ok = !((unsigned)v >= (unsigned)maxval) && e >= -tmax && e < tmax;
        }
        const bool in_win = ok && e >= -HWIN && e < HWIN;
        const unsigned active = __ballot_sync(0xffffffffu, in_win);
        if (in_win) {
            const int key = c * 2 * HWIN + e + HWIN;
            const unsigned peers = __match_any_sync(active, key);
            if ((int)(__ffs(peers) - 1) == lane) atomicAdd(&sh[0][0] + key, __popc(peers));
        } else if (ok) {
            atomicAdd(uh + (long long)c * 2 * tmax + e + tmax, 1u);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * HWIN; k += blockDim.x) {
        const unsigned v = (&sh[0][0])[k];
        if (v) {
            const int c = k / (2 * HWIN), e = k % (2 * HWIN) - HWIN;
            if (e >= -tmax && e < tmax) atomicAdd(uh + (long long)c * 2 * tmax + e + tmax, v);
        }
    }
}

// ------------------------------------------------------------------ host side
static int make_geom(peeb_ws* ws, int h, int w, int itemsize, int bit_depth, int kind, PeeGeom& g) {
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "pee: itemsize must be 1 or 2");
    PEEB_REQUIRE(bit_depth >= 1 && bit_depth <= 8 * itemsize, "pee: bit_depth %d out of range for itemsize %d", bit_depth, itemsize);
    PEEB_REQUIRE(h >= 1 && w >= 1 && (long long)h * w < (1ll << 31), "pee: image size %dx%d unsupported", h, w);
    g.h = h; g.w = w; g.itemsize = itemsize;
    g.rowbytes = w * itemsize;
    g.bulk = ws->use_bulk && (g.rowbytes % 16 == 0);
    g.pitch = g.bulk ? g.rowbytes : (int)align_up((size_t)g.rowbytes, 16);
    g.S = (w + STRIP - 1) / STRIP;
    g.lmw = (w + 7) / 8;
    g.lmpitch = (int)align_up((size_t)g.lmw, 4) + 4;  // +4: partial strips may peek one word past the row
    g.maxval = (1 << bit_depth) - 1;
    // Band height and CTA size.  Narrow images: 256-thread CTAs, four per SM, bands as tall as a
    // quarter of the SM's shared memory allows (capped at 64 rows).  When that leaves fewer than 24
    // rows (wide rows), the two halo rows per side cost too much: take half or all of the SM's
    // shared memory with 512- / 1024-thread CTAs instead (same warps per SM).  Rows are then spread
    // evenly over the bands.
    const size_t sm_total = (size_t)ws->max_smem_optin + 1024;  // 227 KB usable + 1 KB reserved per CTA
    auto fits = [&](int r, size_t budget) {
        g.R = r; g.bandwords = (r * ((w + 1) / 2) + 31) / 32 + 2;
        return band_layout(g, kind).total <= budget;
    };
    auto tallest = [&](size_t budget) {
        int r = 64;
        while (r > 1 && !fits(r, budget)) r -= (r > 16 ? 4 : 1);
        return fits(r, budget) ? r : 0;
    };
    int R = 0;
    g.threads = 256;
    if (const char* e = getenv("PEEB_BAND_KB")) {  // tuning experiments
        R = tallest((size_t)atoi(e) * 1024);
    } else {
        const int r4 = tallest(sm_total / 4 - 1024 - 1024);
        if (r4 >= 24) R = r4;
        else {
            const int r2 = tallest(sm_total / 2 - 1024 - 1024);
            if (r2 >= 24) { R = r2; g.threads = 512; }
            else {
                const int r1 = tallest((size_t)ws->max_smem_optin - 1024);
                R = r1; g.threads = 1024;
                if (r2 >= r1 && r2 > 0) { R = r2; g.threads = 512; }
                if (r4 >= R && r4 > 0) { R = r4; g.threads = 256; }
            }
        }
    }
    if (R <= 0) {
        set_error("pee: image width %d needs more shared memory than one SM has", w);
        return PEEB_E_UNSUPPORTED;
    }
    if (R > h) R = h;
    const int nb = (h + R - 1) / R;
    R = (h + nb - 1) / nb;
    fits(R, (size_t)ws->max_smem_optin);
    g.nb = (h + g.R - 1) / g.R;
    return PEEB_OK;
}

// first-generation kernels stay selectable for A/B runs (PEEB_PEE_V1=1)
static bool use_v1() {
    static const bool v = getenv("PEEB_PEE_V1") != nullptr;
    return v;
}

// upload T / n_bits (host arrays) into table set `slot` of the workspace; returns device pointers
int upload_unit_tables(peeb_ws* ws, int slot, int n_units, const int32_t* T, const int64_t* n_bits,
                              int bit_depth, size_t extra_bytes, cudaStream_t st, int** dT, unsigned** dN,
                              char** extra) {
    const size_t head = align_up((size_t)n_units * 8, 256);
    int rc = scratch_reserve(ws->ptables[slot], head + extra_bytes + 256);
    if (rc) return rc;
    // the pinned mirror is reused by the next call: wait until the previous upload has been consumed
    PEEB_CUDA(cudaEventSynchronize(ws->pev[slot]));
    rc = scratch_reserve(ws->ptables_h[slot], head, true);
    if (rc) return rc;
    int* hT = (int*)ws->ptables_h[slot].ptr;
    unsigned* hN = (unsigned*)(hT + n_units);
    const int tmax = 1 << (bit_depth - 1);
    for (int u = 0; u < n_units; ++u) {
        PEEB_REQUIRE(T[u] >= 1 && T[u] <= tmax, "pee: T[%d]=%d outside 1..%d", u, T[u], tmax);
        PEEB_REQUIRE(n_bits[u] >= 0 && n_bits[u] < (1ll << 31), "pee: n_bits[%d] out of range", u);
        hT[u] = T[u];
        hN[u] = (unsigned)n_bits[u];
    }
    PEEB_CUDA(cudaMemcpyAsync(ws->ptables[slot].ptr, hT, (size_t)n_units * 8, cudaMemcpyHostToDevice, st));
    PEEB_CUDA(cudaEventRecord(ws->pev[slot], st));
    *dT = (int*)ws->ptables[slot].ptr;
    *dN = (unsigned*)((int*)ws->ptables[slot].ptr + n_units);
    *extra = (char*)ws->ptables[slot].ptr + head;
    return PEEB_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    PEEB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PEEB_OK;
}

// launch dispatch over (pixel type, CTA size)
template <typename PixT, int NT>
static int launch_embed(peeb_ws* ws, const PeeGeom& g, const PeeBatch& bt, long long nbands, size_t smem, int* band_cnt,
                        unsigned char* rowcnt, unsigned* ticket, unsigned long long* status, cudaStream_t st) {
    int rc = set_smem(pee_count_kernel<PixT, NT>, smem); if (rc) return rc;
    rc = set_smem(pee_embed_kernel<PixT, NT>, smem); if (rc) return rc;
    { ProfScope p(ws, PEEB_K_PEE_COUNT, st);
      pee_count_kernel<PixT, NT><<<(unsigned)nbands, NT, smem, st>>>(g, bt, band_cnt, rowcnt); }
    { ProfScope p(ws, PEEB_K_PEE_EMBED, st);
      pee_embed_kernel<PixT, NT><<<(unsigned)nbands, NT, smem, st>>>(g, bt, band_cnt, rowcnt, ticket, status); }
    return PEEB_OK;
}
template <typename PixT, int NT>
static int launch_extract(peeb_ws* ws, const PeeGeom& g, const PeeBatch& bt, long long nbands, size_t smem,
                          unsigned* stage_bits, int* stage_cnt, cudaStream_t st) {
    int rc = set_smem(pee_extract_kernel<PixT, NT>, smem); if (rc) return rc;
    ProfScope p(ws, PEEB_K_PEE_EXTRACT, st);
    pee_extract_kernel<PixT, NT><<<(unsigned)nbands, NT, smem, st>>>(g, bt, stage_bits, stage_cnt);
    return PEEB_OK;
}
#define PEEB_DISPATCH(FN, ...)                                                                     \
    (g.itemsize == 2 ? (g.threads == 256 ? FN<unsigned short, 256>(__VA_ARGS__)                     \
                        : g.threads == 512 ? FN<unsigned short, 512>(__VA_ARGS__)                   \
                                           : FN<unsigned short, 1024>(__VA_ARGS__))                 \
                     : (g.threads == 256 ? FN<unsigned char, 256>(__VA_ARGS__)                      \
                        : g.threads == 512 ? FN<unsigned char, 512>(__VA_ARGS__)                    \
                                           : FN<unsigned char, 1024>(__VA_ARGS__)))

static int embed_batch_impl(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                            int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload,
                            int64_t payload_stride, void* marked, int64_t marked_stride, uint8_t* lm,
                            int64_t lm_stride, int64_t* info, cudaStream_t st, int slot = 0) {
    PEEB_REQUIRE(ws && src && T && n_bits && payload && info, "peeb_pee_embed_batch: null pointer");
    PEEB_REQUIRE(n_units >= 1, "peeb_pee_embed_batch: n_units must be >= 1");
    PEEB_REQUIRE(((uintptr_t)payload & 3) == 0 && (payload_stride & 3) == 0, "peeb_pee_embed_batch: payload must be 4-byte aligned");
    PEEB_CUDA(cudaSetDevice(ws->device));
    if (h >= 3 && w >= 3 && !use_v1())
        return embed_batch_impl2(ws, src, src_stride, n_units, h, w, itemsize, bit_depth, T, n_bits, payload, payload_stride,
                                 marked, marked_stride, lm, lm_stride, info, st, slot);
    PeeGeom g;
    int rc = make_geom(ws, h, w, itemsize, bit_depth, 1, g);
    if (rc) return rc;
    if (g.bulk && ((((uintptr_t)src) | (uintptr_t)marked | (uint64_t)src_stride | (uint64_t)marked_stride) & 15)) {
        g.bulk = 0;  // unaligned user buffers: plain copies
        g.pitch = (int)align_up((size_t)g.rowbytes, 16);
    }
    const long long nbands = (long long)n_units * g.nb;
    PEEB_REQUIRE(nbands < (1ll << 30), "peeb_pee_embed_batch: too many bands");
    int* dT; unsigned* dN; char* extra;
    const size_t cnt_bytes = align_up((size_t)nbands * sizeof(int), 256);
    const size_t st_bytes = align_up((size_t)nbands * sizeof(unsigned long long), 256);
    const size_t rc_bytes = align_up((size_t)n_units * h * g.S, 256);
    rc = upload_unit_tables(ws, slot, n_units, T, n_bits, bit_depth, cnt_bytes + st_bytes + 256 + rc_bytes, st, &dT, &dN, &extra);
    if (rc) return rc;
    int* band_cnt = (int*)extra;
    unsigned long long* status = (unsigned long long*)(extra + cnt_bytes);
    unsigned* ticket = (unsigned*)(extra + cnt_bytes + st_bytes);
    unsigned char* rowcnt = (unsigned char*)(extra + cnt_bytes + st_bytes + 256);
    PEEB_CUDA(cudaMemsetAsync(status, 0, st_bytes + 256, st));
    PEEB_CUDA(cudaMemsetAsync(info, 0, sizeof(int64_t) * PEEB_INFO * n_units, st));
    PeeBatch bt{};
    bt.src = (const unsigned char*)src; bt.src_stride = src_stride;
    bt.dst = (unsigned char*)marked; bt.dst_stride = marked_stride;
    bt.lm = lm; bt.lm_stride = lm_stride;
    bt.payload = payload; bt.payload_stride = payload_stride;
    bt.payload_out = nullptr; bt.T = dT; bt.n_bits = dN; bt.info = (long long*)info; bt.n_units = n_units;
    const size_t smem = band_layout(g, 1).total;
    if (h >= 3 && w >= 3) {
        rc = PEEB_DISPATCH(launch_embed, ws, g, bt, nbands, smem, band_cnt, rowcnt, ticket, status, st);
        if (rc) return rc;
        PEEB_CUDA(cudaGetLastError());
    } else {
        // no interior: nothing can be embedded, marked == source, empty location map
        for (int u = 0; u < n_units; ++u) {
            if (marked) PEEB_CUDA(cudaMemcpyAsync((char*)marked + u * marked_stride, (const char*)src + u * src_stride,
                                                  (size_t)h * w * itemsize, cudaMemcpyDeviceToDevice, st));
            if (lm) PEEB_CUDA(cudaMemsetAsync(lm + u * lm_stride, 0, (size_t)h * g.lmw, st));
        }
    }
    { ProfScope p(ws, PEEB_K_PEE_FINAL, st);
      pee_finalize_kernel<<<(n_units + 127) / 128, 128, 0, st>>>(bt, 0); }
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
    PEEB_CUDA(cudaSetDevice(ws->device));
    for (int u = 0; u < n_units; ++u) {
        const size_t pb = peeb_payload_bytes(n_bits[u]);
        PEEB_REQUIRE(n_units == 1 || (int64_t)pb <= payload_stride, "peeb_pee_extract_batch: payload_stride too small for unit %d", u);
    }
    if (h >= 3 && w >= 3 && !use_v1()) {
        // zero every unit's output words (the gather kernel ORs the boundary words in)
        if (n_units == 1) PEEB_CUDA(cudaMemsetAsync(payload_out, 0, peeb_payload_bytes(n_bits[0]), st));
        else PEEB_CUDA(cudaMemsetAsync(payload_out, 0, (size_t)payload_stride * n_units, st));
        return extract_batch_impl2(ws, marked, marked_stride, n_units, h, w, itemsize, bit_depth, T, n_bits, lm, lm_stride,
                                   payload_out, payload_stride, recovered, recovered_stride, info, st, slot);
    }
    PeeGeom g;
    int rc = make_geom(ws, h, w, itemsize, bit_depth, 2, g);
    if (rc) return rc;
    if (g.bulk && ((((uintptr_t)marked) | (uintptr_t)recovered | (uint64_t)marked_stride | (uint64_t)recovered_stride) & 15)) {
        g.bulk = 0;
        g.pitch = (int)align_up((size_t)g.rowbytes, 16);
    }
    const long long nbands = (long long)n_units * g.nb;
    int* dT; unsigned* dN; char* extra;
    const size_t cnt_bytes = align_up((size_t)nbands * 2 * sizeof(int), 256);
    rc = upload_unit_tables(ws, slot, n_units, T, n_bits, bit_depth, cnt_bytes, st, &dT, &dN, &extra);
    if (rc) return rc;
    int* stage_cnt = (int*)extra;
    rc = scratch_reserve(ws->pbits[slot], (size_t)nbands * 2 * g.bandwords * sizeof(unsigned) + 256);
    if (rc) return rc;
    unsigned* stage_bits = (unsigned*)ws->pbits[slot].ptr;
    PeeBatch bt{};
    bt.src = (const unsigned char*)marked; bt.src_stride = marked_stride;
    bt.dst = (unsigned char*)recovered; bt.dst_stride = recovered_stride;
    bt.lm = const_cast<uint8_t*>(lm); bt.lm_stride = lm_stride;
    bt.payload = nullptr; bt.payload_stride = payload_stride; bt.payload_out = payload_out;
    bt.T = dT; bt.n_bits = dN; bt.info = (long long*)info; bt.n_units = n_units;
    for (int u = 0; u < n_units; ++u) {
        const size_t pb = peeb_payload_bytes(n_bits[u]);
        PEEB_REQUIRE(n_units == 1 || (int64_t)pb <= payload_stride, "peeb_pee_extract_batch: payload_stride too small for unit %d", u);
    }
    // zero every unit's output words (the gather kernel ORs the boundary words in)
    if (n_units == 1) PEEB_CUDA(cudaMemsetAsync(payload_out, 0, peeb_payload_bytes(n_bits[0]), st));
    else PEEB_CUDA(cudaMemsetAsync(payload_out, 0, (size_t)payload_stride * n_units, st));
    if (h >= 3 && w >= 3) {
        const size_t smem = band_layout(g, 2).total;
        rc = PEEB_DISPATCH(launch_extract, ws, g, bt, nbands, smem, stage_bits, stage_cnt, st);
        if (rc) return rc;
        PEEB_CUDA(cudaGetLastError());
    } else {
        PEEB_CUDA(cudaMemsetAsync(stage_cnt, 0, cnt_bytes, st));
        for (int u = 0; u < n_units; ++u)
            if (recovered) PEEB_CUDA(cudaMemcpyAsync((char*)recovered + u * recovered_stride, (const char*)marked + u * marked_stride,
                                                     (size_t)h * w * itemsize, cudaMemcpyDeviceToDevice, st));
    }
    {
        ProfScope p(ws, PEEB_K_PEE_GATHER, st);
        dim3 grid((unsigned)(2 * g.nb), (unsigned)n_units);
        pee_gather_kernel<<<grid, 128, 0, st>>>(g, bt, stage_bits, stage_cnt);
    }
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
    const long long npx = (long long)(h - 2) * (w - 2);
    long long blocks = (npx + 256 * 8 - 1) / (256 * 8);
    long long cap = (long long)ws->sm_count * 8 / n_units;
    if (cap < 1) cap = 1;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)n_units);
    ProfScope p(ws, PEEB_K_PEE_HIST, st);
    if (itemsize == 2)
        pee_hist_kernel<unsigned short><<<grid, 256, 0, st>>>((const unsigned char*)src, src_stride, h, w, (1 << bit_depth) - 1, tmax, hist);
    else
        pee_hist_kernel<unsigned char><<<grid, 256, 0, st>>>((const unsigned char*)src, src_stride, h, w, (1 << bit_depth) - 1, tmax, hist);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
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

int peeb_pee_embed_h(peeb_ws* ws, const void* src_host, int shared_flags, int n_units, int h, int w, int itemsize,
                     int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload_host,
                     int64_t payload_stride, void* marked_host, uint8_t* lm_host, int64_t* info_host) {
    PEEB_REQUIRE(ws && src_host && T && n_bits && info_host, "peeb_pee_embed_h: null pointer");
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
    // chunks alternate between the two streams: copy-in, kernels and copy-out of neighbouring
    // chunks overlap (PCIe is full duplex, the copy engines run beside the SMs)
    const int cu = chunk_units(n_units, img);
    for (int u0 = 0, c = 0; u0 < n_units; u0 += cu, ++c) {
        const int n = std::min(cu, n_units - u0), slot = c & 1;
        cudaStream_t st = streams[slot];
        if (!shared_src)
            PEEB_CUDA(cudaMemcpy2DAsync(d1 + (size_t)u0 * img_al, img_al, (const char*)src_host + (size_t)u0 * img, img, img, n,
                                        cudaMemcpyHostToDevice, st));
        if (!shared_pay && payload_stride > 0 && payload_host)
            PEEB_CUDA(cudaMemcpy2DAsync(d2 + (size_t)u0 * pstride, pstride, payload_host + (size_t)u0 * payload_stride,
                                        (size_t)payload_stride, (size_t)payload_stride, n, cudaMemcpyHostToDevice, st));
        rc = embed_batch_impl(ws, shared_src ? d1 : d1 + (size_t)u0 * img_al, shared_src ? 0 : (int64_t)img_al, n, h, w,
                              itemsize, bit_depth, T + u0, n_bits + u0, (const uint8_t*)(shared_pay ? d2 : d2 + (size_t)u0 * pstride),
                              shared_pay ? 0 : (int64_t)pstride, marked_host ? d1 + o_marked + (size_t)u0 * img_al : nullptr, (int64_t)img_al,
                              lm_host ? (uint8_t*)(d2 + o_lm + (size_t)u0 * lm_al) : nullptr, (int64_t)lm_al,
                              (int64_t*)(d2 + o_info) + (size_t)u0 * PEEB_INFO, st, slot);
        if (rc) return rc;
        if (marked_host)
            PEEB_CUDA(cudaMemcpy2DAsync((char*)marked_host + (size_t)u0 * img, img, d1 + o_marked + (size_t)u0 * img_al, img_al,
                                        img, n, cudaMemcpyDeviceToHost, st));
        if (lm_host)
            PEEB_CUDA(cudaMemcpy2DAsync(lm_host + (size_t)u0 * lmb, lmb, d2 + o_lm + (size_t)u0 * lm_al, lm_al, lmb, n,
                                        cudaMemcpyDeviceToHost, st));
        PEEB_CUDA(cudaMemcpyAsync(info_pin + (size_t)u0 * PEEB_INFO, (int64_t*)(d2 + o_info) + (size_t)u0 * PEEB_INFO,
                                  (size_t)n * PEEB_INFO * 8, cudaMemcpyDeviceToHost, st));
    }
    PEEB_CUDA(cudaStreamSynchronize(streams[0]));
    PEEB_CUDA(cudaStreamSynchronize(streams[1]));
    memcpy(info_host, info_pin, (size_t)n_units * PEEB_INFO * 8);
    return PEEB_OK;
}

int peeb_pee_extract_h(peeb_ws* ws, const void* marked_host, int n_units, int h, int w, int itemsize, int bit_depth,
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
    const int cu = std::min(chunk_units(n_units, img), 65535);
    for (int u0 = 0, c = 0; u0 < n_units; u0 += cu, ++c) {
        const int n = std::min(cu, n_units - u0), slot = c & 1;
        cudaStream_t st = streams[slot];
        PEEB_CUDA(cudaMemcpy2DAsync(d1 + (size_t)u0 * img_al, img_al, (const char*)marked_host + (size_t)u0 * img, img, img, n,
                                    cudaMemcpyHostToDevice, st));
        PEEB_CUDA(cudaMemcpy2DAsync(d2 + o_lm + (size_t)u0 * lm_al, lm_al, lm_host + (size_t)u0 * lmb, lmb, lmb, n,
                                    cudaMemcpyHostToDevice, st));
        rc = extract_batch_impl(ws, d1 + (size_t)u0 * img_al, (int64_t)img_al, n, h, w, itemsize, bit_depth, T + u0, n_bits + u0,
                                (const uint8_t*)(d2 + o_lm + (size_t)u0 * lm_al), (int64_t)lm_al,
                                (uint8_t*)(d2 + (size_t)u0 * pstride), (int64_t)pstride,
                                recovered_host ? d1 + o_rec + (size_t)u0 * img_al : nullptr, (int64_t)img_al,
                                (int64_t*)(d2 + o_info) + (size_t)u0 * PEEB_INFO, st, slot);
        if (rc) return rc;
        if (recovered_host)
            PEEB_CUDA(cudaMemcpy2DAsync((char*)recovered_host + (size_t)u0 * img, img, d1 + o_rec + (size_t)u0 * img_al, img_al,
                                        img, n, cudaMemcpyDeviceToHost, st));
        if (payload_stride > 0)
            PEEB_CUDA(cudaMemcpy2DAsync(payload_out_host + (size_t)u0 * payload_stride, (size_t)payload_stride,
                                        d2 + (size_t)u0 * pstride, pstride, (size_t)payload_stride, n, cudaMemcpyDeviceToHost, st));
        PEEB_CUDA(cudaMemcpyAsync(info_pin + (size_t)u0 * PEEB_INFO, (int64_t*)(d2 + o_info) + (size_t)u0 * PEEB_INFO,
                                  (size_t)n * PEEB_INFO * 8, cudaMemcpyDeviceToHost, st));
    }
    PEEB_CUDA(cudaStreamSynchronize(streams[0]));
    PEEB_CUDA(cudaStreamSynchronize(streams[1]));
    memcpy(info_host, info_pin, (size_t)n_units * PEEB_INFO * 8);
    return PEEB_OK;
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
