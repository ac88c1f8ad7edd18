// Row a10: the PEE kernels, "row pair" layout.
//
// A first generation of these kernels (in the history of this repository; profiles/r01_ncu_full_pee_kernels.md)
// walked 128-column strips downwards and ranked carriers with warp ballots; ncu showed it bound by the
// integer ALU pipe at ~60 executed instructions per colour pixel.  This layout needs about a third less:
//
//   * a LANE owns two adjacent image rows of a cell (<= 64 columns) and walks them left to right in
//     16-byte steps, so its carriers come in raster order: the rank of a carrier inside a
//     (row, cell) is a lane-local running count, the payload bits of a cell are a 32-bit window in
//     a register (no ballots, no POPC, no per-bit shared-memory bytes);
//   * the two rows of a lane have opposite colour parity, so the code of a step is static; rows
//     above/below come from shared memory (4 x LDS.128 per 2 row-steps);
//   * the rhombus predictor never unpacks pixels: 4x - (N+S+W+E) is accumulated straight from the
//     packed words with IDP.2A / IDP.4A (integer dot product with byte weights, FMA pipe), and with
//     the accumulator preloaded with 3 + 4T one arithmetic shift gives e + T
//     (e = x - floor(S/4) = ceil((4x - S)/4));
//   * classification is branch-free: delta = clamp(e + T, 0, 2T) - T covers expand / shift+ / shift-
//     (SURVEY Appendix A), one unsigned compare gives the overflow flag for all three classes;
//   * rows whose pixels must not change (outside the pass, outside the image interior) run with
//     T = 0, which makes every pixel a shift by 0: no special cases in the step code.
//
// Band / look-back / staging structure is unchanged: one CTA per band of R rows (+2 halo rows per
// side), TMA bulk row copies into a padded, bank-conflict-free shared layout, pass-0 counts from a
// separate count kernel, decoupled look-back for pass 1, per-band bit staging + gather for extract.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "peeb_common.cuh"
#include "peeb_pee.cuh"

namespace peeb {

#ifdef PEEB_PHASE_TIMING
// development aid: per-phase clock64 totals of thread 0 of every embed CTA (see scripts/phase_timing.py)
__device__ unsigned long long g_phase[32];
#define PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); atomicAdd(&g_phase[i], (unsigned long long)(_t - _t0)); _t0 = _t; } } while (0)
#define PHASE_INIT long long _t0 = clock64()
#else
#define PHASE_MARK(i) do {} while (0)
#define PHASE_INIT do {} while (0)
#endif

#ifdef PEEB_DEBUG_BOUNDS
// Bounds-checked build (-DPEEB_DEBUG_BOUNDS; codec_tcc_b200/build.py --bounds, tests/test_gpu_bounds_build.py): the
// memory checker of the toolkit cannot be run on the boxes these kernels are developed on, so every index the band
// kernels derive from the geometry -- staged rows and the slack reads around a cell, count tables, location-map
// rows, band streams, staging slots, payload words -- is compared with the size of the region it points into.  A
// violation is counted (the first one is kept: site, byte offset, limit) and the access still happens; the parity
// suite is then run against this build and peeb_debug_bounds() must report none.
__device__ unsigned long long g_bounds[6];  // violations, first site, first offset, its limit, checks done (items), -
__device__ unsigned long long g_bounds_site[128];  // violations per site (printed by peeb_debug_bounds)
__device__ __noinline__ void bounds_fail(int site, long long off, long long lim) {
    atomicAdd(&g_bounds_site[site & 127], 1ull);
    if (atomicAdd(&g_bounds[0], 1ull) == 0ull) { g_bounds[1] = (unsigned long long)site; g_bounds[2] = (unsigned long long)off; g_bounds[3] = (unsigned long long)lim; }
}
// bytes [off, off + n) must lie inside [0, lim)
#define BOUNDS(site, off, n, lim) do { const long long _o = (long long)(off), _l = (long long)(lim); \
    if (_o < 0 || _o + (long long)(n) > _l) bounds_fail(site, _o, _l); } while (0)
#define BOUNDS_TICK() do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_bounds[4], 1ull); } while (0)
#else
#define BOUNDS(site, off, n, lim) do {} while (0)
#define BOUNDS_TICK() do {} while (0)
#endif

struct Geom2 {
    int h, w, itemsize, maxval;
    int R, nb;              // band height, bands per unit
    int rowbytes, pitch;    // pitch = align_up(rowbytes, 128) + 16: consecutive rows fall 16 bytes apart in the banks
    int bulk;
    int cws;                // 16-byte steps per cell
    int CW;                 // cell width in pixels (<= 64, so a cell holds <= 32 carriers of one colour)
    int ncol;               // cells per row
    int rpw, rpw_log2;      // row pairs per warp item (power of two); 32/rpw cells side by side
    int nic;                // warp items of a sweep: ceil(ncol / (32/rpw)) -- a sweep has at most rpw row pairs
    int tpitch;             // bytes per row of the carrier-count tables (one byte per cell, multiple of 16)
    int lmw, lmpitch;       // location map: global row bytes, shared row pitch (bytes, multiple of 4)
    int bandwords;          // extract staging: 32-bit words per (unit, pass, band)
    int threads;
    int minb;               // CTAs per SM the launch is sized for
    int lm_direct;          // embed: location-map bits go straight to global memory (no shared copy)
};

struct Smem2 {
    size_t img, lm, tab, misc, bar, tn0, tn1, tw0, tw1, stream, tab0, xch, pre, total;
};
constexpr int CLUSTER_MAX = 16;  // CTAs of one image's cluster (small-image path)
constexpr int HWIN = 512;  // errors -HWIN <= e < HWIN of the threshold-selection histogram are counted in shared memory
__host__ __device__ inline Smem2 layout2(const Geom2& g, int kind_ /*0 count, 1 embed, 2 extract, 3 histogram, 4 / 5 embed / extract of the cluster path*/) {
    Smem2 L{};
    size_t o = 0;
    const int kind = kind_ == 4 ? 1 : kind_ == 5 ? 2 : kind_;
    L.img = o; o += align_up((size_t)16 + (size_t)(g.R + 5) * g.pitch + 192, 16);
    L.lm = o;
    if (kind == 2 || (kind == 1 && !g.lm_direct)) o += align_up((size_t)(g.R + 2) * g.lmpitch + 16 + (kind_ == 2 ? 48 : 0), 16);
    L.tab = o;
    if (kind == 1) o += (size_t)(g.R + 2) * g.tpitch;   // pass-1 carriers per (row, cell), one byte each
    if (kind == 3) o += (size_t)4 * HWIN * sizeof(unsigned);  // [colour][e + HWIN]
    L.misc = o; o += 64 * sizeof(int);
    L.bar = o; o += 16;
    L.tn0 = L.tn1 = L.tw0 = L.tw1 = L.stream = o;
    if (kind == 2) {
        const size_t cells = (size_t)g.R * g.ncol;
        L.tn0 = o; o += (size_t)g.R * g.tpitch;
        L.tn1 = o; o += (size_t)g.R * g.tpitch;
        L.tw0 = o; o += align_up(cells * sizeof(unsigned), 16);
        L.tw1 = o; o += align_up(cells * sizeof(unsigned), 16);
        L.stream = o; o += align_up((size_t)2 * g.bandwords * sizeof(unsigned), 16);
    }
    L.pre = o;
    // wide images (more than 32 cells per row): per-(row, table word) prefixes and row totals of the current pass
    if (kind_ == 1 && g.tpitch > 32) o += align_up((size_t)(g.R + 2) * (g.tpitch >> 2) * sizeof(unsigned short) + (size_t)(g.R + 2) * sizeof(int), 16);
    L.tab0 = L.xch = o;
    if (kind_ == 4) { L.tab0 = o; o += (size_t)(g.R + 2) * g.tpitch; }   // pass-0 carriers per (row, cell)
    if (kind_ >= 4) { L.xch = o; o += (size_t)8 * CLUSTER_MAX * sizeof(unsigned); }  // what the CTAs of a cluster tell each other
    L.total = o;
    return L;
}

__device__ __forceinline__ long long img_region_bytes(const Geom2& g) { return (long long)align_up((size_t)16 + (size_t)(g.R + 5) * g.pitch + 192, 16); }
// byte offset of shared row `rs` (0 = image row r_first); every second group of 8 rows is shifted
// by 16 bytes so that lanes two rows apart (the row-pair layout) still hit distinct banks
__device__ __forceinline__ int row_off(const Geom2& g, int rs) { return 16 + rs * g.pitch + ((rs & 8) << 1); }

// ------------------------------------------------------------------ band staging
// Staging is split in two so that table loads and other set-up overlap the copy: issue_rows2 starts
// it (TMA bulk row copies by the lanes of warp 0, or plain loads), wait_rows2 makes the rows visible.
template <typename PixT>
__device__ __forceinline__ void issue_rows2(const Geom2& g, const unsigned char* usrc, unsigned char* simg, int r_first,
                                            int lo, int hi, uint64_t* bar) {
    if (hi <= lo) return;
    BOUNDS(12, lo, 0, g.h + 1); BOUNDS(12, hi, 0, g.h + 1);
    BOUNDS(13, row_off(g, lo - r_first), g.rowbytes, img_region_bytes(g));
    BOUNDS(13, row_off(g, hi - 1 - r_first), g.rowbytes, img_region_bytes(g));
    if (g.bulk) {
        // a warp issues its bulk copies one lane after the other: spread the rows over the warps,
        // four lanes each (the transaction count may be posted after the first copies complete)
        if (threadIdx.x == 0) mbar_expect_tx(bar, (unsigned)(hi - lo) * (unsigned)g.rowbytes);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        if (lane < 4)
            for (int r = lo + warp * 4 + lane; r < hi; r += nwarps * 4)
                bulk_g2s(simg + row_off(g, r - r_first), usrc + (size_t)r * g.rowbytes, (unsigned)g.rowbytes, bar);
    } else {
        for (int r = lo + (int)(threadIdx.x >> 5); r < hi; r += (int)(blockDim.x >> 5)) {
            const PixT* s = reinterpret_cast<const PixT*>(usrc + (size_t)r * g.rowbytes);
            PixT* d = reinterpret_cast<PixT*>(simg + row_off(g, r - r_first));
            for (int c = threadIdx.x & 31; c < g.w; c += 32) d[c] = s[c];
        }
    }
}
#ifdef PEEB_AB_POLLALL
#define POLL_THREAD true
#else
#define POLL_THREAD (threadIdx.x < 32)
#endif
// (one warp polls the mbarrier -- its acquire, followed by the CTA barrier, orders the copied rows before every
// thread's reads; seven more polling warps would only take issue slots from the CTAs that are computing)
__device__ __forceinline__ void wait_rows2(const Geom2& g, int lo, int hi, uint64_t* bar) {
    if (g.bulk && hi > lo && POLL_THREAD) mbar_wait(bar, 0);
    __syncthreads();
}
#define WAIT_ROWS2_MARKED(g, lo, hi, bar, slot) do { if ((g).bulk && (hi) > (lo) && POLL_THREAD) mbar_wait(bar, 0); PHASE_MARK(slot); __syncthreads(); } while (0)
template <typename PixT>
__device__ __forceinline__ void load_rows2(const Geom2& g, const unsigned char* usrc, unsigned char* simg, int r_first,
                                           int lo, int hi, uint64_t* bar) {
    issue_rows2<PixT>(g, usrc, simg, r_first, lo, hi, bar);
    wait_rows2(g, lo, hi, bar);
}

// ------------------------------------------------------------------ payload order without block scans
// Carrier counts live in byte tables, one byte per (row, cell), rows `tpitch` bytes apart (bytes past ncol are 0).
// A lane of the row-pair layout owns the cells (rowa, cell) and (rowa + 1, cell) of a sweep over rows
// [row_lo, row_hi); the stream position of its first carrier is
//     (carriers of the rows above, all cells)  +  (carriers of the cells to its left in its own row).
// Both terms come from the lane's own two table rows: IDP.4A sums the bytes (all of them / those left of
// the cell), a warp scan over the row pairs adds the rows above.  No block-wide scan, no barrier beyond the
// one that completes the table.  Every lane of the warp must call this (shuffles); lanes whose rows or
// cell are outside the sweep pass acta / actb = false and get offsets they never use.
__device__ __forceinline__ int idp4_sum(unsigned a, unsigned wsel, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(wsel), "r"(c));
    return d;
}
struct CellPrefix {
    int offa, offb;   // carriers before (rowa, cell) / (rowa + 1, cell) in the sweep's raster order
    int row0_total;   // carriers of the first row of the sweep (the halo row above a band, when there is one)
    int total;        // carriers of the whole sweep
};
// warp scan over the row pairs: (carriers of row a, row b, cells left of the cell in a, in b) -> offsets
__device__ __forceinline__ CellPrefix prefix_scan(const Geom2& g, int tota, int totb, int prea, int preb) {
    const int rp = threadIdx.x & (g.rpw - 1);
    const int v = tota + totb;
    int incl = v;
    for (int o = 1; o < g.rpw; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o, g.rpw);
        if (rp >= o) incl += t;
    }
    CellPrefix r;
    r.offa = incl - v + prea;
    r.offb = incl - totb + preb;
    r.row0_total = __shfl_sync(0xffffffffu, tota, 0, g.rpw);
    r.total = __shfl_sync(0xffffffffu, incl, g.rpw - 1, g.rpw);
    return r;
}
// tables of at most 16 cells per row (one 16-byte word per row): the two rows of the lane, already loaded
__device__ __forceinline__ CellPrefix cell_prefix16(const Geom2& g, const uint4& va, const uint4& vb, int cell) {
    const int jc = cell >> 2;
    const unsigned partial = 0x01010101u & ((1u << (8 * (cell & 3))) - 1u);
    const unsigned one = 0x01010101u;
    const unsigned m0 = jc > 0 ? one : partial, m1 = jc > 1 ? one : jc == 1 ? partial : 0u,
                   m2 = jc > 2 ? one : jc == 2 ? partial : 0u, m3 = jc == 3 ? partial : 0u;
    const int tota = idp4_sum(va.w, one, idp4_sum(va.z, one, idp4_sum(va.y, one, idp4_sum(va.x, one, 0))));
    const int totb = idp4_sum(vb.w, one, idp4_sum(vb.z, one, idp4_sum(vb.y, one, idp4_sum(vb.x, one, 0))));
    const int prea = idp4_sum(va.w, m3, idp4_sum(va.z, m2, idp4_sum(va.y, m1, idp4_sum(va.x, m0, 0))));
    const int preb = idp4_sum(vb.w, m3, idp4_sum(vb.z, m2, idp4_sum(vb.y, m1, idp4_sum(vb.x, m0, 0))));
    return prefix_scan(g, tota, totb, prea, preb);
}
template <bool GLOBAL>
__device__ __forceinline__ uint4 table_row16(const unsigned char* t, bool in) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (in) v = GLOBAL ? __ldg(reinterpret_cast<const uint4*>(t)) : *reinterpret_cast<const uint4*>(t);
    return v;
}
template <bool GLOBAL>
__device__ __forceinline__ CellPrefix cell_prefix_inl(const Geom2& g, const unsigned char* ta, bool rowa_in, bool rowb_in,
                                                      int cell) {
    // ta: table row of rowa (row of rowa + 1 follows at ta + tpitch); shared or global memory
    if (g.tpitch == 16)
        return cell_prefix16(g, table_row16<GLOBAL>(ta, rowa_in), table_row16<GLOBAL>(ta + 16, rowb_in), cell);
    const int jc = cell >> 2;
    const unsigned partial = 0x01010101u & ((1u << (8 * (cell & 3))) - 1u);
    int tota = 0, totb = 0, prea = 0, preb = 0;
    for (int j4 = 0; j4 < g.tpitch; j4 += 16) {
        const uint4 va = table_row16<GLOBAL>(ta + j4, rowa_in), vb = table_row16<GLOBAL>(ta + g.tpitch + j4, rowb_in);
        const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = (j4 >> 2) + k;
            const int sa = idp4_sum(wa[k], 0x01010101u, 0), sb = idp4_sum(wb[k], 0x01010101u, 0);
            tota += sa; totb += sb;
            if (j < jc) { prea += sa; preb += sb; }
            if (j == jc) { prea = idp4_sum(wa[k], partial, prea); preb = idp4_sum(wb[k], partial, preb); }
        }
    }
    return prefix_scan(g, tota, totb, prea, preb);
}

// (not inlined in the embed kernel: called once per warp item from four sweeps; keeps the kernel's code, and its
// instruction-cache footprint, small)
template <bool GLOBAL>
__device__ __noinline__ CellPrefix cell_prefix(const Geom2& g, const unsigned char* ta, bool rowa_in, bool rowb_in,
                                               int cell) {
    return cell_prefix_inl<GLOBAL>(g, ta, rowa_in, rowb_in, cell);
}

// Wide images: a row of the count table has more than 16 cells, and summing it per lane and item (cell_prefix_inl) was
// 14 % of the embed kernel's instructions on 3000-pixel rows.  Once per pass the CTA turns the table into what a lane
// needs: wpre[row][word] = carriers of the cells left of that table word (4 cells), rowtot[row] = carriers of the row;
// a lane then reads its own table word, one prefix and one total per row.  Every thread of the CTA calls this (one
// barrier at the end); nrows <= R + 2 rows starting at `tab`.
template <bool GLOBAL>
__device__ __forceinline__ void build_prefix(const Geom2& g, const unsigned char* tab, int nrows, unsigned short* wpre, int* rowtot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, wpr = g.tpitch >> 2;
    for (int r = warp; r < nrows; r += nwarps) {
        const unsigned* row = reinterpret_cast<const unsigned*>(tab + (size_t)r * g.tpitch);
        int carry = 0;
        for (int j0 = 0; j0 < wpr; j0 += 32) {
            const int j = j0 + lane;
            unsigned w = 0u;
            if (j < wpr) w = GLOBAL ? __ldg(row + j) : row[j];
            const int sum = idp4_sum(w, 0x01010101u, 0);
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (j < wpr) wpre[r * wpr + j] = (unsigned short)(carry + incl - sum);
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) rowtot[r] = carry;
    }
    __syncthreads();
}

// Decoupled look-back over the bands of a unit, run by a whole warp (every warp of the CTA may run it: all of
// them get the same answer, only `publish` = true writes this band's words).  Aggregates are summed back to the
// nearest band that already knows its inclusive prefix; returns the carriers of the earlier bands.
// look-back status words: the value is the whole message, so relaxed gpu-scope accesses suffice
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned warp_lookback(unsigned long long* stt, int band, unsigned total, bool publish) {
    const int lane = threadIdx.x & 31;
    if (publish && lane == 0) st_relaxed_gpu(stt + band, ST_AGG | total);
    unsigned before = 0;
    for (int k0 = band - 1; k0 >= 0; k0 -= 32) {
        const int k = k0 - lane;
        unsigned long long v = ST_PFX;  // lanes before band 0: prefix 0
        if (k >= 0) do { v = ld_relaxed_gpu(stt + k); } while ((v & ST_MASK) == 0);
        const unsigned pfx = __ballot_sync(0xffffffffu, (v & ST_MASK) == ST_PFX);
        const int first = __ffs(pfx) - 1;  // nearest band with a prefix (-1: none in this window)
        unsigned val = (first < 0 || lane <= first) ? (unsigned)(v & 0xffffffffu) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        before += val;
        if (first >= 0) break;
    }
    if (publish && lane == 0) st_relaxed_gpu(stt + band, ST_PFX | (unsigned long long)(before + total));
    return before;
}

// 32 payload bits starting at stream bit p (bit p on top), straight from the packed MSB-first payload in
// global memory; bits at or past n_bits read as 0 (zero padding), words past the payload are never loaded.
__device__ __forceinline__ unsigned payload_window(const unsigned* __restrict__ pay, unsigned p, unsigned n_bits) {
    const unsigned wi = p >> 5;
    unsigned w0 = 0u, w1 = 0u;
    // (a payload row holds peeb_payload_bytes(n_bits) = the bytes of its bits rounded up to whole words)
    if ((wi << 5) < n_bits) BOUNDS(11, 4ll * wi, 4, 4ll * ((n_bits + 31u) >> 5));
    if (((wi + 1u) << 5) < n_bits) BOUNDS(11, 4ll * wi + 4, 4, 4ll * ((n_bits + 31u) >> 5));
    if ((wi << 5) < n_bits) w0 = __byte_perm(__ldg(pay + wi), 0, 0x0123);
    if (((wi + 1u) << 5) < n_bits) w1 = __byte_perm(__ldg(pay + wi + 1), 0, 0x0123);
    const unsigned r0 = n_bits - (wi << 5);             // valid bits from the start of w0 (when w0 was loaded)
    if (r0 < 32u) w0 &= ~(0xffffffffu >> r0);
    else if (r0 < 64u) w1 &= ~(0xffffffffu >> (r0 - 32u));
    return __funnelshift_l(w1, w0, p & 31u);
}

// Write-back of band rows, split so that the tail work of a kernel overlaps it.  Callers synchronise
// the block before store_rows2_issue (all pixel writes done) and call store_rows2_wait before the
// shared buffer is reused or the kernel ends.
template <typename PixT>
__device__ __forceinline__ void store_rows2_issue(const Geom2& g, unsigned char* udst, const unsigned char* simg,
                                                  int r_first, int lo, int hi) {
    if (hi <= lo) return;
    BOUNDS(14, lo, 0, g.h + 1); BOUNDS(14, hi, 0, g.h + 1);
    BOUNDS(15, row_off(g, lo - r_first), g.rowbytes, img_region_bytes(g));
    BOUNDS(15, row_off(g, hi - 1 - r_first), g.rowbytes, img_region_bytes(g));
    if (g.bulk) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        if (lane < 4) {
            fence_async_smem();
            for (int r = lo + warp * 4 + lane; r < hi; r += nwarps * 4)
                bulk_s2g(udst + (size_t)r * g.rowbytes, simg + row_off(g, r - r_first), (unsigned)g.rowbytes);
            bulk_commit();
        }
    } else {
        for (int r = lo + (int)(threadIdx.x >> 5); r < hi; r += (int)(blockDim.x >> 5)) {
            const PixT* s = reinterpret_cast<const PixT*>(simg + row_off(g, r - r_first));
            PixT* d = reinterpret_cast<PixT*>(udst + (size_t)r * g.rowbytes);
            for (int c = threadIdx.x & 31; c < g.w; c += 32) d[c] = s[c];
        }
    }
}
__device__ __forceinline__ void store_rows2_wait(const Geom2& g) {
    if (g.bulk && (threadIdx.x & 31) < 4) bulk_wait_read0();
}

// ------------------------------------------------------------------ packed-pixel arithmetic
__device__ __forceinline__ int idp2(unsigned a, unsigned b, int c) {  // c + a.lo16*b.s8[0] + a.hi16*b.s8[1]
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int idp4(unsigned a, unsigned b, int c) {  // c + sum a.u8[k]*b.s8[k]
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <int J> __device__ __forceinline__ unsigned comp(const uint4& v) {
    return J == 0 ? v.x : J == 1 ? v.y : J == 2 ? v.z : v.w;
}
template <int J> __device__ __forceinline__ unsigned& compr(uint4& v) {
    if constexpr (J == 0) return v.x; else if constexpr (J == 1) return v.y; else if constexpr (J == 2) return v.z; else return v.w;
}
// word J of the step, or the last word of the previous step (J = -1) / first of the next (J = 4)
template <int J> __device__ __forceinline__ unsigned compx(const uint4& v, unsigned prev, unsigned next) {
    if constexpr (J < 0) return prev; else if constexpr (J > 3) return next; else return comp<J>(v);
}

// A 16-byte step holds NS colour pixels of a row; slot S sits at column 2S + Q of the step, Q the
// column parity of the colour in that row.  qsum = init + 4x - (N + S + W + E) from packed words.
template <typename PixT> struct PixOps;
template <> struct PixOps<unsigned short> {
    static constexpr int ITEM = 2, NS = 4, PXS = 8;
    template <int Q, int S>
    static __device__ __forceinline__ int qsum(const uint4& M, unsigned prev, unsigned next, const uint4& U,
                                               const uint4& D, int init) {
        if constexpr (Q == 0) {  // pixel = low half of word S: W = high half of word S-1, E = high half of word S
            int q = idp2(comp<S>(D), 0x00ffu, init);
            q = idp2(comp<S>(U), 0x00ffu, q);
            q = idp2(compx<S - 1>(M, prev, next), 0xff00u, q);
            return idp2(comp<S>(M), 0xff04u, q);
        } else {                 // pixel = high half of word S: W = low half of word S, E = low half of word S+1
            int q = idp2(comp<S>(D), 0xff00u, init);
            q = idp2(comp<S>(U), 0xff00u, q);
            q = idp2(compx<S + 1>(M, prev, next), 0x00ffu, q);
            return idp2(comp<S>(M), 0x04ffu, q);
        }
    }
    template <int Q, int S> static __device__ __forceinline__ int getx(const uint4& M) {
        return idp2(comp<S>(M), Q == 0 ? 0x0001u : 0x0100u, 0);
    }
    template <int Q, int S> static __device__ __forceinline__ void setx(uint4& M, int v) {
        unsigned& wv = compr<S>(M);
        wv = __byte_perm(wv, (unsigned)v, Q == 0 ? 0x3254 : 0x5410);
    }
    template <int Q, int S> static __device__ __forceinline__ int add4x(const uint4& M, int acc) {  // acc + 4x
        return idp2(comp<S>(M), Q == 0 ? 0x0004u : 0x0400u, acc);
    }
    template <int Q, int S> static __device__ __forceinline__ int addx(const uint4& M, int acc) {   // acc + x
        return idp2(comp<S>(M), Q == 0 ? 0x0001u : 0x0100u, acc);
    }
    // pixel += d inside the packed word (exact while the new value stays inside the pixel's range)
    template <int Q, int S> static __device__ __forceinline__ void addpacked(uint4& M, int d) {
        unsigned& wv = compr<S>(M);
        // (add and shift kept off the FMA pipe: VIADDMNMX with a neutral max, PRMT as the 16-bit shift)
        wv = __viaddmax_u32(wv, Q == 0 ? (unsigned)d : __byte_perm((unsigned)d, 0u, 0x1044), 0u);
    }
    // pixel = (pixel + d) mod 2^16, the neighbour in the word untouched whatever the value (extract: any input)
    template <int Q, int S> static __device__ __forceinline__ void addwrap(uint4& M, int d) {
        unsigned& wv = compr<S>(M);
        if constexpr (Q == 0) wv = __byte_perm(wv + (unsigned)d, wv, 0x7610);
        else wv += (unsigned)d << 16;
    }
};
template <> struct PixOps<unsigned char> {
    static constexpr int ITEM = 1, NS = 8, PXS = 16;
    static constexpr __host__ __device__ unsigned wcentre(int B) {
        unsigned wv = 4u << (8 * B);
        if (B > 0) wv |= 0xffu << (8 * (B - 1));
        if (B < 3) wv |= 0xffu << (8 * (B + 1));
        return wv;
    }
    template <int Q, int S>
    static __device__ __forceinline__ int qsum(const uint4& M, unsigned prev, unsigned next, const uint4& U,
                                               const uint4& D, int init) {
        constexpr int J = S / 2, B = 2 * (S % 2) + Q;  // word, byte inside the word
        constexpr unsigned wv = 0xffu << (8 * B);
        int q = idp4(comp<J>(D), wv, init);
        q = idp4(comp<J>(U), wv, q);
        if constexpr (B == 0) q = idp4(compx<J - 1>(M, prev, next), 0xff000000u, q);
        if constexpr (B == 3) q = idp4(compx<J + 1>(M, prev, next), 0x000000ffu, q);
        return idp4(comp<J>(M), wcentre(B), q);
    }
    template <int Q, int S> static __device__ __forceinline__ int getx(const uint4& M) {
        constexpr int J = S / 2, B = 2 * (S % 2) + Q;
        return idp4(comp<J>(M), 1u << (8 * B), 0);
    }
    template <int Q, int S> static __device__ __forceinline__ void setx(uint4& M, int v) {
        constexpr int J = S / 2, B = 2 * (S % 2) + Q;
        constexpr unsigned sel = B == 0 ? 0x3214u : B == 1 ? 0x3240u : B == 2 ? 0x3410u : 0x4210u;
        unsigned& wv = compr<J>(M);
        wv = __byte_perm(wv, (unsigned)v, sel);
    }
    template <int Q, int S> static __device__ __forceinline__ int add4x(const uint4& M, int acc) {  // acc + 4x
        constexpr int J = S / 2, B = 2 * (S % 2) + Q;
        return idp4(comp<J>(M), 4u << (8 * B), acc);
    }
    template <int Q, int S> static __device__ __forceinline__ int addx(const uint4& M, int acc) {   // acc + x
        constexpr int J = S / 2, B = 2 * (S % 2) + Q;
        return idp4(comp<J>(M), 1u << (8 * B), acc);
    }
    template <int Q, int S> static __device__ __forceinline__ void addpacked(uint4& M, int d) {
        constexpr int J = S / 2, B = 2 * (S % 2) + Q;
        unsigned& wv = compr<J>(M);
        wv += (unsigned)d << (8 * B);
    }
    template <int Q, int S> static __device__ __forceinline__ void addwrap(uint4& M, int d) {  // (pixel + d) mod 2^8
        constexpr int J = S / 2, B = 2 * (S % 2) + Q;
        constexpr unsigned sel = B == 0 ? 0x7650u : B == 1 ? 0x7610u : B == 2 ? 0x7210u : 0x3210u;
        unsigned& wv = compr<J>(M);
        const unsigned sum = wv + ((unsigned)d << (8 * B));
        wv = B == 3 ? sum : __byte_perm(sum, wv, sel);
    }
};

template <int I, int N, class F> __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ uint4 lds128(const unsigned char* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void sts128(unsigned char* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// per-row constants of the embed side (T = 0: every pixel is "shifted by 0", i.e. left alone)
struct KE {
    int init, T8, T2, negT, T;
};
__device__ __forceinline__ KE make_ke(int T) { return KE{3 + 4 * T, 8 * T, 2 * T, -T, T}; }
// extract side
struct KX {
    int init, T16, T2, T;
};
__device__ __forceinline__ KX make_kx(int T) { return KX{3 + 8 * T, 16 * T, 2 * T, T}; }

struct Stats2 {
    long long sse = 0;
    unsigned flagged = 0;
    // warp-steps of the apply sweeps (the same in every lane of a warp): all of them, those that took the generic
    // code because they touch a border column, and those redone after the fast code saw a value leave the range
    unsigned steps = 0, steps_edge = 0, steps_redone = 0;
};

// ---- predicated tails (explicit PTX: the compiler turns these into select chains otherwise) ----
// embed: a carrier (q < 8T unsigned, i.e. -T <= e < T; range already checked) takes the next payload bit,
// the top bit of W (tested as the sign: one predicated add, no 64-bit multiply-high and its register pair)
__device__ __forceinline__ void take_bit(int& nv, unsigned& W, int q, int T8) {
    asm("{\n\t.reg .pred p, pb;\n\t"
        "setp.lt.u32 p, %2, %3;\n\t"
        "setp.lt.and.s32 pb, %1, 0, p;\n\t"
        "@pb add.s32 %0, %0, 1;\n\t"
        "@p shf.l.wrap.b32 %1, 0, %1, 1;\n\t}"
        : "+r"(nv), "+r"(W) : "r"(q), "r"(T8));
}
// count: carrier <=> q < 8T and 0 <= v < maxval (v = x + e)
__device__ __forceinline__ void count_if(int& n, int q, int T8, int v, int maxval) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.u32 p, %1, %2;\n\t"
        "setp.lt.and.u32 p, %3, %4, p;\n\t"
        "@p add.s32 %0, %0, 1;\n\t}"
        : "+r"(n) : "r"(q), "r"(T8), "r"(v), "r"(maxval));
}
// extract: a carrier (q < 16T, i.e. -2T <= e' < 2T) appends bit 0 of e' + 2T = bit 2 of q to W
__device__ __forceinline__ void collect_bit(unsigned& W, int& n, int q, int T16) {
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "setp.lt.u32 p, %2, %3;\n\t"
        "shl.b32 t, %2, 29;\n\t"
        "@p shf.l.wrap.b32 %0, t, %0, 1;\n\t"
        "@p add.s32 %1, %1, 1;\n\t}"
        : "+r"(W), "+r"(n) : "r"(q), "r"(T16));
}

// ------------------------------------------------------------------ the sweep
// Rows [row_lo, row_hi) of the staged band (at most 2 * rpw of them: make_geom2 sizes the bands so), colour
// with column parity QA in row row_lo.  A warp item = (the sweep's row pairs) x (32/rpw neighbouring cells);
// lane -> (row pair, cell).
// lane -> (row pair, cell of warp item `item`) of a sweep over rows [row_lo, row_hi)
struct Lane2 {
    int rowa, cell;
    bool rowa_in, rowb_in, acta, actb;
};
__device__ __forceinline__ Lane2 lane_of(const Geom2& g, int row_lo, int row_hi, int item) {
    const int lane = threadIdx.x & 31;
    const int rp = lane & (g.rpw - 1), part = lane >> g.rpw_log2, parts = 32 >> g.rpw_log2;
    Lane2 l;
    l.rowa = row_lo + 2 * rp;
    l.rowa_in = l.rowa < row_hi; l.rowb_in = l.rowa + 1 < row_hi;
    if (!l.rowa_in) l.rowa = row_lo;  // idle lane: any staged row will do, nothing is written
    l.cell = item * parts + part;
    const bool colv = l.cell < g.ncol;
    l.acta = colv && l.rowa_in; l.actb = colv && l.rowb_in;
    if (!colv) l.cell = g.ncol - 1;
    return l;
}
// What a body needs before its first item (payload order, payload bits: global loads) can be fetched ahead of the
// sweep, e.g. while the band's rows are still on their way: sweep2_prime_order, then body.begin_bits() once the
// body knows its base B; sweep2 then skips begin() for that item.
template <class Body>
__device__ __forceinline__ void sweep2_prime_order(const Geom2& g, int row_lo, int row_hi, int T, Body& body) {
    const int warp = threadIdx.x >> 5;
    if (row_hi <= row_lo || warp >= g.nic) return;
    const Lane2 l = lane_of(g, row_lo, row_hi, warp);
    body.begin_order(l.rowa, l.cell, l.acta, l.actb, l.rowa_in, l.rowb_in, T);
    body.primed = true;
}
template <typename PixT, int QA, class Body>
__device__ __forceinline__ void sweep2(const Geom2& g, unsigned char* simg, int r_first, int row_lo, int row_hi, int T,
                                       Body& body) {
    using P = PixOps<PixT>;
    if (row_hi <= row_lo) return;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int item = warp; item < g.nic; item += nwarps) {
        const Lane2 l = lane_of(g, row_lo, row_hi, item);
        const int rs = l.rowa - r_first;
        const int oa = row_off(g, rs);
        unsigned char* pa = simg + oa + l.cell * g.CW * P::ITEM;
        const unsigned char* pu = pa + (row_off(g, rs - 1) - oa);
        unsigned char* pb = pa + (row_off(g, rs + 1) - oa);
        const unsigned char* pd = pa + (row_off(g, rs + 2) - oa);
        const int c0 = l.cell * g.CW;
#ifdef PEEB_DEBUG_BOUNDS
        {   // every step reads 16 bytes of four rows, the word before the first step and the word after each step,
            // and stores 16 bytes of the two middle rows: the whole walk must stay inside the staged band
            const long long lim = img_region_bytes(g), walk = 16ll * g.cws;
            BOUNDS(1, pu - simg, walk, lim); BOUNDS(1, pd - simg, walk, lim);
            BOUNDS(2, (pa - simg) - 4, walk + 4 + 4, lim); BOUNDS(2, (pb - simg) - 4, walk + 4 + 4, lim);
            BOUNDS(3, rs - 1, 0, g.R + 5); BOUNDS(3, rs + 2, 1, g.R + 5);
            BOUNDS_TICK();
        }
#endif
        if (!body.primed) body.begin(l.rowa, l.cell, l.acta, l.actb, l.rowa_in, l.rowb_in, T);
        body.primed = false;
        // Steps that touch a border column (or, for extract, cells with location-map bits) take the generic code:
        // one warp-uniform bit per step, worked out once per item.
        unsigned spm;
        {
            constexpr int PXL = P::PXS == 8 ? 3 : 4;
            const int t = g.w - 1 - c0 - P::PXS;                 // step s is at the right border iff s * PXS > t
            const int s_edge = t < 0 ? 0 : (t >> PXL) + 1;
            unsigned m = s_edge >= 32 ? 0u : (0xffffffffu << s_edge);
            if (c0 == 0) m |= 1u;
            if (body.item_special()) m = 0xffffffffu;
            spm = __reduce_or_sync(0xffffffffu, m);
        }
        // the row whose colour sits on even columns looks one word back, the other one word ahead
        unsigned prev = *reinterpret_cast<const unsigned*>((QA == 0 ? pa : pb) - 4);
#ifdef PEEB_UNROLL2
#pragma unroll 2
#else
#pragma unroll 1
#endif
        for (int s = 0; s < g.cws; ++s) {
            const uint4 U = lds128(pu), D = lds128(pd);
            uint4 A = lds128(pa), B = lds128(pb);
            const unsigned next = *reinterpret_cast<const unsigned*>((QA == 0 ? pb : pa) + 16);
            const unsigned newprev = QA == 0 ? A.w : B.w;
            body.template step<QA>(c0 + s * P::PXS, s, (spm >> s) & 1u, U, A, B, D, prev, next, pa, pb);
            prev = newprev;
            pu += 16; pa += 16; pb += 16; pd += 16;
        }
        body.end();
    }
}
template <typename PixT, class Body>
__device__ __forceinline__ void sweep2_colour(const Geom2& g, unsigned char* simg, int r_first, int colour, int row_lo,
                                              int row_hi, int T, Body& body) {
    if ((row_lo + colour) & 1) sweep2<PixT, 1>(g, simg, r_first, row_lo, row_hi, T, body);
    else sweep2<PixT, 0>(g, simg, r_first, row_lo, row_hi, T, body);
}


// ---- count carriers of one colour per (row, cell) ----------------------------------------------
template <typename PixT, bool GLOBAL>
struct Count2 {
    using P = PixOps<PixT>;
    const Geom2& g;
    int row0;               // image row of table row 0
    unsigned char* tab;     // byte table, [(row - row0) * tpitch + cell]: global (GLOBAL) or shared memory
    int total;              // GLOBAL: carriers seen by this lane
    KE ka, kb;
    int na, nb, ia;
    bool acta, actb, primed;
    __device__ __forceinline__ void begin(int rowa, int cell, bool a, bool b, bool, bool, int T) {
        acta = a; actb = b;
        ka = kb = make_ke(T);  // inactive rows compute like the others; nothing of theirs is kept
        na = nb = 0;
        ia = (rowa - row0) * g.tpitch + cell;
    }
    __device__ __forceinline__ bool item_special() const { return false; }
    template <int Q, bool EDGE>
    __device__ __forceinline__ void row(const uint4& M, unsigned prev, unsigned next, const uint4& U, const uint4& D,
                                        const KE& k, int c, int& n) {
        static_for<0, P::NS>([&](auto Sc) {
            constexpr int S = decltype(Sc)::value;
            KE kk = k;
            if (EDGE) {
                const int col = c + 2 * S + Q;
                kk = make_ke((col >= 1 && col <= g.w - 2) ? k.T : 0);
            }
            const int q = P::template qsum<Q, S>(M, prev, next, U, D, kk.init);  // 4(e + T) + r, 0 <= r <= 3
            // 4(x + e) + r = q + 4x - 4T: the carrier value x + e is in [0, maxval) iff that is in [0, 4 maxval)
            count_if(n, q, kk.T8, P::template add4x<Q, S>(M, q) + 4 * kk.negT, 4 * g.maxval);
        });
    }
    template <int QA>
    __device__ __forceinline__ void step(int c, int, bool special, const uint4& U, uint4& A, uint4& B, const uint4& D,
                                         unsigned prev, unsigned next, unsigned char*, unsigned char*) {
        if (special) {
            row<QA, true>(A, prev, next, U, B, ka, c, na);
            row<1 - QA, true>(B, prev, next, A, D, kb, c, nb);
        } else {
            row<QA, false>(A, prev, next, U, B, ka, c, na);
            row<1 - QA, false>(B, prev, next, A, D, kb, c, nb);
        }
    }
    __device__ __forceinline__ void end() {
        // (the table of the count kernel has a row per image row, a band's shared table R + 2 rows)
        if (acta) BOUNDS(4, ia, 1, (long long)(GLOBAL ? g.h : g.R + 2) * g.tpitch);
        if (actb) BOUNDS(4, ia + g.tpitch, 1, (long long)(GLOBAL ? g.h : g.R + 2) * g.tpitch);
        if (acta) tab[ia] = (unsigned char)na;
        if (actb) tab[ia + g.tpitch] = (unsigned char)nb;
        if (GLOBAL) total += (acta ? na : 0) + (actb ? nb : 0);
    }
};

// ---- full apply of one colour ---------------------------------------------------------------------
// The stream position of a lane's first carrier comes from the byte table of the pass (cell_prefix: pass 0 reads
// the count kernel's table in global memory, pass 1 the shared table of the band), its 32 payload bits straight
// from global memory (payload_window).
// A step first runs the fast code, which assumes that no pixel over/underflows (every expandable
// pixel is a carrier, nothing goes to the location map) and only watches for a value leaving
// [0, maxval); if any lane of the warp sees one, or the step touches a border column, the step is
// redone from the saved words by the generic code.
template <typename PixT, bool GTAB>
struct Apply2 {
    using P = PixOps<PixT>;
    const Geom2& g;
    int row0, own_lo, own_hi;
    const unsigned char* tab;   // byte table, row `row0` first
    const unsigned* pay;        // the unit's packed payload
    unsigned n_bits;
    unsigned B;                 // stream index of the first carrier of the sweep's first own row
    bool halo;                  // the sweep starts one row above the band: that row's carriers precede B
    unsigned* slm;  // location-map rows (row lm_row0 first), lmwords words apart: shared copy of the band, or
                    // the unit's rows in global memory (null: the caller wants no map)
    int lm_row0, lmwords;
    Stats2* st;
    const unsigned short* wpre = nullptr;  // wide images: build_prefix's tables of this pass (row pre_row0 first), else null
    const int* rowtot = nullptr;
    int pre_row0 = 0;
    KE ka, kb;
    unsigned Wa, Wb;
    long long ssea, sseb;
    bool sta, stb, owna, ownb, primed;
    int offa, offb;
    unsigned* lma;
    // begin = begin_order (where the lane's carriers sit in the pass: table rows -> offsets) + begin_bits (their
    // payload bits); a caller that primes the first item may set B between the two
    __device__ __forceinline__ void begin_order(int rowa, int cell, bool a, bool b, bool rowa_in, bool rowb_in, int T) {
        sta = a; stb = b;
        ka = kb = make_ke(T);
        CellPrefix cp;
        {
            [[maybe_unused]] const long long tlim = (long long)(GTAB ? g.h : g.R + 2) * g.tpitch;
            if (rowa_in) BOUNDS(5, (long long)(rowa - row0) * g.tpitch, g.tpitch, tlim);
            if (rowb_in) BOUNDS(5, (long long)(rowa + 1 - row0) * g.tpitch, g.tpitch, tlim);
            BOUNDS(5, cell, 1, g.tpitch);
        }
        if (wpre) {
            const int ra = rowa - pre_row0, jc = cell >> 2, wpr = g.tpitch >> 2;
            const unsigned partial = 0x01010101u & ((1u << (8 * (cell & 3))) - 1u);
            const unsigned* ta = reinterpret_cast<const unsigned*>(tab + (long long)(rowa - row0) * g.tpitch) + jc;
            int tota = 0, totb = 0, prea = 0, preb = 0;
            if (rowa_in) {
                tota = rowtot[ra];
                prea = idp4_sum(GTAB ? __ldg(ta) : *ta, partial, (int)wpre[ra * wpr + jc]);
            }
            if (rowb_in) {
                totb = rowtot[ra + 1];
                preb = idp4_sum(GTAB ? __ldg(ta + wpr) : ta[wpr], partial, (int)wpre[(ra + 1) * wpr + jc]);
            }
            cp = prefix_scan(g, tota, totb, prea, preb);
        } else {
            cp = cell_prefix<GTAB>(g, tab + (long long)(rowa - row0) * g.tpitch, rowa_in, rowb_in, cell);
        }
        offa = cp.offa - (halo ? cp.row0_total : 0);
        offb = cp.offb - (halo ? cp.row0_total : 0);
        owna = a && rowa >= own_lo && rowa < own_hi;
        ownb = b && rowa + 1 >= own_lo && rowa + 1 < own_hi;
        ssea = sseb = 0;
        lma = slm ? slm + (long long)(rowa - lm_row0) * lmwords : nullptr;
    }
    __device__ __forceinline__ void begin_bits() {
        Wa = sta ? payload_window(pay, B + (unsigned)offa, n_bits) : 0u;
        Wb = stb ? payload_window(pay, B + (unsigned)offb, n_bits) : 0u;
    }
    __device__ __forceinline__ void begin(int rowa, int cell, bool a, bool b, bool rowa_in, bool rowb_in, int T) {
        begin_order(rowa, cell, a, b, rowa_in, rowb_in, T);
        begin_bits();
    }
    __device__ __forceinline__ bool item_special() const { return false; }
    // The fast code of a step has two halves.  probe: predictions and differences of all colour pixels of the two
    // rows from the words as loaded, nothing modified, and the watch for a value leaving [0, maxval) -- the warp
    // votes on that before any state changes, so the generic code needs no saved copies to start over from.
    // commit: payload bits, SSE, and the differences added inside the packed words (O += d << position of the
    // pixel is exact because the new value stays in range) -- no pixel is unpacked or re-inserted.
    template <int Q>
    __device__ __forceinline__ bool probe(const uint4& M, unsigned prev, unsigned next, const uint4& U, const uint4& D,
                                          const KE& k, int (&q)[P::NS], int (&d)[P::NS]) {
        bool bad = false;
        static_for<0, P::NS>([&](auto Sc) {
            constexpr int S = decltype(Sc)::value;
            q[S] = P::template qsum<Q, S>(M, prev, next, U, D, k.init);
            // clamp(e, -T, T) = e | +T | -T, with the add inside the min (VIADDMNMX, ALU pipe: the FMA pipe is
            // where the IDPs of the predictor run, the busiest unit of this kernel)
            d[S] = max(__viaddmin_s32(q[S] >> 2, k.negT, k.T), k.negT);
            const int nv0 = P::template addx<Q, S>(M, d[S]);        // x + e | x + T | x - T
            bad |= (unsigned)nv0 >= (unsigned)g.maxval;             // (maxval itself is fine for a shift: rare, generic code sorts it out)
        });
        return bad;
    }
    template <int Q>
    __device__ __forceinline__ void commit(uint4& O, const KE& k, const int (&q)[P::NS], int (&d)[P::NS], unsigned& W,
                                           long long& sse) {
        static_for<0, P::NS>([&](auto Sc) {
            constexpr int S = decltype(Sc)::value;
            take_bit(d[S], W, q[S], k.T8);
            sse += (long long)d[S] * (long long)d[S];
            P::template addpacked<Q, S>(O, d[S]);
        });
    }
    template <int Q>
    __device__ __forceinline__ void generic(const uint4& M, uint4& O, unsigned prev, unsigned next, const uint4& U,
                                            const uint4& D, const KE& k, int c, unsigned& W, long long& sse, bool own,
                                            unsigned* lmrow) {
        static_for<0, P::NS>([&](auto Sc) {
            constexpr int S = decltype(Sc)::value;
            const int col = c + 2 * S + Q;
            const KE kk = make_ke((col >= 1 && col <= g.w - 2) ? k.T : 0);
            const int q = P::template qsum<Q, S>(M, prev, next, U, D, kk.init);
            const int x = P::template getx<Q, S>(M);
            const bool expd = (unsigned)q < (unsigned)kk.T8;        // -T <= e < T
            const int cc = max(min(q >> 2, kk.T2), 0);
            const int nv0 = x + cc + kk.negT;
            const unsigned lim = (unsigned)g.maxval - (expd ? 1u : 0u);
            const bool ok = (unsigned)nv0 <= lim;                   // else: location map, pixel unchanged
            int nv = ok ? nv0 : x;
            if (expd && ok) {                                       // carrier: next payload bit of the cell
                nv += (int)(W >> 31);
                W <<= 1;
            }
            const int d = nv - x;
            sse += (long long)d * (long long)d;
            if (own && !ok && kk.T != 0) {
                if (lmrow) {
                    BOUNDS(6, col >> 5, 1, lmwords);
                    BOUNDS(6, (lmrow - slm) / lmwords, 1, g.lm_direct ? g.h : g.R + 2);
                    atomicOr(lmrow + (col >> 5), lm_bitmask(col));
                }
                ++st->flagged;
            }
            P::template setx<Q, S>(O, nv);
        });
    }
    template <int QA>
    __device__ __forceinline__ void step(int c, int, bool special, const uint4& U, uint4& A, uint4& B, const uint4& D,
                                         unsigned prev, unsigned next, unsigned char* pa, unsigned char* pb) {
        int qa[P::NS], da[P::NS], qb[P::NS], db[P::NS];
        bool redo = special;
        if (!special) {
            bool bad = probe<QA>(A, prev, next, U, B, ka, qa, da);
            bad |= probe<1 - QA>(B, prev, next, A, D, kb, qb, db);
            redo = __any_sync(0xffffffffu, bad);
        }
        if (!redo) {
            commit<QA>(A, ka, qa, da, Wa, ssea);
            commit<1 - QA>(B, kb, qb, db, Wb, sseb);
        } else {
            const uint4 A0 = A, B0 = B;
            generic<QA>(A0, A, prev, next, U, B0, ka, c, Wa, ssea, owna, lma);
            generic<1 - QA>(B0, B, prev, next, A0, D, kb, c, Wb, sseb, ownb, lma ? lma + lmwords : nullptr);
            if (special) ++st->steps_edge; else ++st->steps_redone;
        }
        // (steps past the end of the row run the generic code with T = 0 everywhere: they store back what they read,
        // into the padding of the shared rows)
        if (sta) sts128(pa, A);
        if (stb) sts128(pb, B);
    }
    __device__ __forceinline__ void end() {
        if (owna) st->sse += ssea;
        if (ownb) st->sse += sseb;
        st->steps += g.cws;
    }
};

// carriers of a whole byte table of `nrows` rows, summed by one warp (every lane gets the total)
__device__ __forceinline__ int table_total(const Geom2& g, const unsigned char* tab, int nrows) {
    int sum = 0;
    const int wpr = g.tpitch >> 2;
    for (int k = threadIdx.x & 31; k < nrows * wpr; k += 32) sum = idp4_sum(reinterpret_cast<const unsigned*>(tab)[k], 0x01010101u, sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    return sum;
}

// ------------------------------------------------------------------ threshold selection (Appendix A)
// Prediction-error histogram of the ORIGINAL image, per colour, over interior pixels that are not flagged for
// expansion (0 <= x + e < maxval) with -tmax <= e < tmax: hist[unit][colour][e + tmax].  Same band staging and
// row-pair sweep as the count kernel (the errors of a lane's pixels come from the same IDP accumulations);
// |e| < HWIN is counted in a shared window of the CTA, the rare rest straight in global memory.
template <typename PixT>
struct Hist2 {
    using P = PixOps<PixT>;
    const Geom2& g;
    unsigned* sh;   // shared window of this colour: [e + HWIN]
    unsigned* gh;   // the unit's global histogram of this colour: [e + tmax]
    int tmax;
    bool acta, actb, primed;
    __device__ __forceinline__ void begin(int, int, bool a, bool b, bool, bool, int) { acta = a; actb = b; }
    __device__ __forceinline__ bool item_special() const { return false; }
    template <int Q, bool EDGE>
    __device__ __forceinline__ void row(const uint4& M, unsigned prev, unsigned next, const uint4& U, const uint4& D,
                                        int c, bool act) {
        static_for<0, P::NS>([&](auto Sc) {
            constexpr int S = decltype(Sc)::value;
            bool ok = act;
            if (EDGE) {
                const int col = c + 2 * S + Q;
                ok = ok && col >= 1 && col <= g.w - 2;
            }
            const int q = P::template qsum<Q, S>(M, prev, next, U, D, 3);  // 4e + r, 0 <= r <= 3
            const int e = q >> 2;
            ok = ok && (unsigned)P::template add4x<Q, S>(M, q) < 4u * (unsigned)g.maxval;  // 0 <= x + e < maxval
            ok = ok && e >= -tmax && e < tmax;
#ifdef PEEB_HIST_MATCH
            // warp-aggregated form: lanes with the same error elect one of them to add their number
            const bool inwin = ok && e >= -HWIN && e < HWIN;
            const unsigned key = inwin ? (unsigned)(e + HWIN) : (0x80000000u | (threadIdx.x & 31));
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            if (inwin && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(sh + e + HWIN, (unsigned)__popc(peers));
            if (ok && !inwin) atomicAdd(gh + e + tmax, 1u);
#else
            if (ok) {
                if (e >= -HWIN && e < HWIN) atomicAdd(sh + e + HWIN, 1u);
                else atomicAdd(gh + e + tmax, 1u);
            }
#endif
        });
    }
    template <int QA>
    __device__ __forceinline__ void step(int c, int, bool special, const uint4& U, uint4& A, uint4& B, const uint4& D,
                                         unsigned prev, unsigned next, unsigned char*, unsigned char*) {
        if (special) {
            row<QA, true>(A, prev, next, U, B, c, acta);
            row<1 - QA, true>(B, prev, next, A, D, c, actb);
        } else {
            row<QA, false>(A, prev, next, U, B, c, acta);
            row<1 - QA, false>(B, prev, next, A, D, c, actb);
        }
    }
    __device__ __forceinline__ void end() {}
};

template <typename PixT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) pee2_hist_kernel(Geom2 g, const unsigned char* __restrict__ src,
                                                             long long src_stride, int tmax, unsigned* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem2 L = layout2(g, 3);
    unsigned char* simg = smem_raw + L.img;
    unsigned* sh = reinterpret_cast<unsigned*>(smem_raw + L.tab);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    const int unit = blockIdx.x / g.nb, band = blockIdx.x % g.nb;
    if (threadIdx.x == 0 && g.bulk) { mbar_init(bar, 1); fence_mbar_init(); }
    for (int k = threadIdx.x; k < 4 * HWIN; k += blockDim.x) sh[k] = 0u;
    __syncthreads();
    const int r0 = band * g.R, r_first = r0 - 2;
    const unsigned char* usrc = src + (long long)unit * src_stride;
    load_rows2<PixT>(g, usrc, simg, r_first, max(r0 - 1, 0), min(r0 + g.R + 1, g.h), bar);
    const int own_lo = max(r0, 1), own_hi = min(r0 + g.R, g.h - 1);
    unsigned* uh = hist + (long long)unit * 4 * tmax;
    for (int colour = 0; colour < 2; ++colour) {
        Hist2<PixT> body{g, sh + colour * 2 * HWIN, uh + (long long)colour * 2 * tmax, tmax};
        sweep2_colour<PixT>(g, simg, r_first, colour, own_lo, own_hi, 0, body);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * HWIN; k += blockDim.x) {
        const unsigned v = sh[k];
        const int colour = k / (2 * HWIN), e = k % (2 * HWIN) - HWIN;
        if (v && e >= -tmax && e < tmax) atomicAdd(uh + (long long)colour * 2 * tmax + e + tmax, v);
    }
}

// One CTA per unit: T0 = min{T >= 1 : sum_c sum_{-T <= e < T} hist_c[e] >= n_bits}, tmax + 1 if there is none.
__global__ void __launch_bounds__(256) pee2_pick_T_kernel(const unsigned* __restrict__ hist, int tmax,
                                                          const unsigned* __restrict__ n_bits, int* __restrict__ T,
                                                          int* __restrict__ active) {
    const int unit = blockIdx.x;
    const unsigned* h0 = hist + (long long)unit * 4 * tmax;
    const unsigned* h1 = h0 + 2 * tmax;
    const unsigned long long want = n_bits[unit];
    __shared__ unsigned long long s_warp[8];
    __shared__ int s_found;
    if (threadIdx.x == 0) s_found = tmax + 1;
    __syncthreads();
    unsigned long long before = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t0 = 1; t0 <= tmax; t0 += 256) {
        const int t = t0 + (int)threadIdx.x;  // this thread adds the bins that T = t brings in: e = -t and e = t - 1
        unsigned long long v = 0;
        if (t <= tmax) v = (unsigned long long)h0[tmax - t] + h0[tmax + t - 1] + h1[tmax - t] + h1[tmax + t - 1];
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long wbefore = 0, total = 0;
        for (int k = 0; k < 8; ++k) { if (k < warp) wbefore += s_warp[k]; total += s_warp[k]; }
        const unsigned long long est = before + wbefore + incl;
        if (t <= tmax && est >= want) atomicMin(&s_found, t);
        __syncthreads();
        if (s_found <= tmax) break;
        before += total;
    }
    // no T whose estimate holds the payload (Appendix A: an error, even if the real capacity at tmax turned out larger
    // than its estimate): the unit is embedded once at tmax, zero padded, and keeps the capacity status (active = 2)
    if (threadIdx.x == 0) { T[unit] = min(s_found, tmax); active[unit] = s_found <= tmax ? 1 : 2; }
}

// After an embed with device-side thresholds: units whose payload did not fit (status PEEB_E_CAPACITY) and whose
// T can still grow get T + 1 and stay active for the next round; everything else is done.  remaining: units to redo.
__global__ void pee2_retry_kernel(int n_units, int tmax, long long* __restrict__ info, int* __restrict__ T,
                                  int* __restrict__ active, int* __restrict__ remaining) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    int again = 0;
    if (active[u] == 2) info[(long long)u * PEEB_INFO + 7] = PEEB_E_CAPACITY;
    else if (active[u] && info[(long long)u * PEEB_INFO + 7] == PEEB_E_CAPACITY && T[u] < tmax) { T[u] += 1; again = 1; }
    if (again) { info[(long long)u * PEEB_INFO + 5] = 0; info[(long long)u * PEEB_INFO + 6] = 0; }  // summed with atomics by the causal kernels
    active[u] = again;
    if (again) atomicAdd(remaining, 1);
}

// host side of a verify-and-increment round, for the other translation unit (causal predictor): units that fell
// short get T + 1 and stay active; *remaining_h (pinned) = how many; synchronises the stream
int threshold_retry_round(int n_units, int tmax, long long* info, int* T, int* active, int* remaining, int* remaining_h, cudaStream_t st) {
    PEEB_CUDA(cudaMemsetAsync(remaining, 0, sizeof(int), st));
    pee2_retry_kernel<<<(n_units + 255) / 256, 256, 0, st>>>(n_units, tmax, info, T, active, remaining);
    PEEB_CUDA(cudaMemcpyAsync(remaining_h, remaining, sizeof(int), cudaMemcpyDeviceToHost, st));
    PEEB_CUDA(cudaStreamSynchronize(st));
    return PEEB_OK;
}

// ------------------------------------------------------------------ K_A: pass-0 counts
// rowcnt: one byte per (row, cell) of every unit, rows `tpitch` bytes apart.
template <typename PixT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) pee2_count_kernel(Geom2 g, PeeBatch bt, int* __restrict__ band_cnt,
                                                              unsigned char* __restrict__ rowcnt,
                                                              unsigned* __restrict__ ticket,
                                                              unsigned long long* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem2 L = layout2(g, 0);
    unsigned char* simg = smem_raw + L.img;
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    const int unit = blockIdx.x / g.nb, band = blockIdx.x % g.nb;
    if (threadIdx.x == 0) { misc[0] = 0; if (g.bulk) { mbar_init(bar, 1); fence_mbar_init(); } }
    __syncthreads();
    const int r0 = band * g.R, r_first = r0 - 2;
    const unsigned char* usrc = bt.src + (long long)unit * bt.src_stride;
    issue_rows2<PixT>(g, usrc, simg, r_first, max(r0 - 1, 0), min(r0 + g.R + 1, g.h), bar);
    // this launch also resets what the embed kernel behind it starts from (no memsets on the stream):
    // the band's look-back status word, the ticket counter, the unit's summary row
    if (threadIdx.x == 0) {
        status[blockIdx.x] = 0ull;
        if (blockIdx.x == 0) *ticket = 0u;
    }
    const bool skip = bt.active && !bt.active[unit];  // threshold search: this unit is done
    if (!skip && band == 0 && threadIdx.x < PEEB_INFO) bt.info[(long long)unit * PEEB_INFO + threadIdx.x] = 0;
    const int own_lo = max(r0, 1), own_hi = min(r0 + g.R, g.h - 1);
    if (!skip) {
        // bytes past ncol of this band's table rows are summed with the rest by the embed kernel: the rows start from
        // zero (16-byte stores; the barrier below orders them before the sweep's byte stores)
        uint4* rt = reinterpret_cast<uint4*>(rowcnt + ((long long)unit * g.h + own_lo) * g.tpitch);
        for (int k = threadIdx.x; k < ((own_hi - own_lo) * g.tpitch) >> 4; k += blockDim.x) rt[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    wait_rows2(g, max(r0 - 1, 0), min(r0 + g.R + 1, g.h), bar);
    if (skip) return;
    Count2<PixT, true> body{g, 0, rowcnt + (long long)unit * g.h * g.tpitch, 0};
    sweep2_colour<PixT>(g, simg, r_first, 0, own_lo, own_hi, bt.T[unit], body);
    int tot = body.total;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(misc, tot);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int all = misc[0];
        band_cnt[unit * g.nb + band] = all;  // cap0 of the unit = their sum (the embed kernel adds them up)
    }
}

// ------------------------------------------------------------------ K_B: fused two-pass embed
// Barriers per band: ticket, rows staged, pass 0 applied, pass-1 table complete, pass 1 applied.  Everything
// between them is per warp: payload order (cell_prefix), look-back (every warp runs it, warp 0 publishes),
// payload bits (payload_window).
template <typename PixT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) pee2_embed_kernel(Geom2 g, PeeBatch bt, const int* __restrict__ band_cnt,
                                                              const unsigned char* __restrict__ rowcnt,
                                                              unsigned* __restrict__ ticket,
                                                              unsigned long long* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem2 L = layout2(g, 1);
    unsigned char* simg = smem_raw + L.img;
    unsigned* slm = reinterpret_cast<unsigned*>(smem_raw + L.lm);
    unsigned char* tab = smem_raw + L.tab;
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool use_pre = g.tpitch > 32;  // (32 cells per row: two 16-byte loads per table row are as cheap as the prefix pass; measured on the 2048-wide sweep)
    int* rowtot = reinterpret_cast<int*>(smem_raw + L.pre);
    unsigned short* wpre = reinterpret_cast<unsigned short*>(smem_raw + L.pre + (size_t)(g.R + 2) * sizeof(int));

    // in-order ticket: a band only ever waits on bands with smaller tickets
    if (threadIdx.x == 0) {
        misc[40] = (int)atomicAdd(ticket, 1u);
        if (g.bulk) { mbar_init(bar, 1); fence_mbar_init(); }
    }
    // bytes of the pass-1 table past ncol are summed with the rest: keep them 0
    for (int k = threadIdx.x; k < ((g.R + 2) * g.tpitch) >> 4; k += blockDim.x) reinterpret_cast<uint4*>(tab)[k] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // Tickets run band-major over the batch (band 0 of every unit, then band 1, ...): with more units
    // than resident CTAs the earlier bands of a unit have finished when a band looks back, so the
    // look-back finds an inclusive prefix at once instead of waiting on bands that run beside it.
    const int tk = misc[40];
    const int band = tk / bt.n_units, unit = tk - band * bt.n_units;
    if (bt.active && !bt.active[unit]) return;  // threshold search: this unit is done
    const int r0 = band * g.R, r_first = r0 - 2;
    const unsigned char* usrc = bt.src + (long long)unit * bt.src_stride;
    const int T = bt.T[unit];
    const unsigned n_bits = bt.n_bits[unit];
    const unsigned* payload = reinterpret_cast<const unsigned*>(bt.payload + (long long)unit * bt.payload_stride);
    long long* info = bt.info + (long long)unit * PEEB_INFO;
    const int own_lo = max(r0, 1), own_hi = min(r0 + g.R, g.h - 1);
    const int p0_lo = max(r0 - 1, 1), p0_hi = min(r0 + g.R + 1, g.h - 1);
    const int s_lo = max(r0 - 2, 0), s_hi = min(r0 + g.R + 2, g.h);
    PHASE_INIT;
#ifndef PEEB_AB_NOPREFETCH
    // the table rows of the warp's first item and the unit's band totals are wanted right after this: their lines are
    // asked for BEFORE the band's bulk copies, behind which small loads would otherwise wait (phase timing: 8 % of a
    // band's wall time went to the first item's table rows, which arrived when the 60-200 KB of rows had)
    if (warp < g.nic && p0_hi > p0_lo) {
        const Lane2 l = lane_of(g, p0_lo, p0_hi, warp);
        const unsigned char* ta = rowcnt + ((long long)unit * g.h + l.rowa) * g.tpitch;
        if (l.rowa_in) asm volatile("prefetch.global.L1 [%0];" :: "l"(ta));
        if (l.rowb_in) asm volatile("prefetch.global.L1 [%0];" :: "l"(ta + g.tpitch));
    }
    if (lane * 32 < g.nb) asm volatile("prefetch.global.L1 [%0];" :: "l"(band_cnt + unit * g.nb + lane * 32));
#endif
    // wide rows: the pass-0 prefix tables are built from the count kernel's table BEFORE the band's bulk copies are issued --
    // its loads do not wait behind 200 KB of rows, and the payload order is known while the rows arrive (dx3000 embed
    // 1.53 -> 1.48 ms against building them after the copies were issued)
    if (use_pre) build_prefix<true>(g, rowcnt + ((long long)unit * g.h + p0_lo) * g.tpitch, max(p0_hi - p0_lo, 0), wpre, rowtot);
    issue_rows2<PixT>(g, usrc, simg, r_first, s_lo, s_hi, bar);
    PHASE_MARK(16);  // copies issued
    // every warp: carriers of pass 0 in the earlier bands, and in the whole unit (cap0), from the count kernel's band
    // totals.  Their loads are in flight together with the table rows of the warp's first item (sweep2_prime_order),
    // the payload bits follow: two global round trips while the band's rows arrive, not three.
    int ccv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ccv[i] = lane + 32 * i < g.nb ? __ldg(band_cnt + unit * g.nb + lane + 32 * i) : 0;
    if (!g.lm_direct) {
        for (int k = threadIdx.x; k < ((g.R + 1) * g.lmpitch) >> 2; k += blockDim.x) slm[k] = 0;
    } else if (bt.lm) {
        // flagged pixels are rare: their bits are OR-ed straight into the (zeroed) global rows of this band
        const int b_lo = r0, b_hi = min(r0 + g.R, g.h);
        unsigned* glm = reinterpret_cast<unsigned*>(bt.lm + (long long)unit * bt.lm_stride + (size_t)b_lo * g.lmw);
        if ((g.lmw & 15) == 0 && ((uintptr_t)glm & 15) == 0) {
            uint4* gq = reinterpret_cast<uint4*>(glm);
            for (int k = threadIdx.x; k < (b_hi - b_lo) * (g.lmw >> 4); k += blockDim.x) gq[k] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            for (int k = threadIdx.x; k < (b_hi - b_lo) * (g.lmw >> 2); k += blockDim.x) glm[k] = 0u;
        }
    }
    PHASE_MARK(0);  // set-up
    Stats2 st;
    unsigned* lmbase = slm;
    int lmrow0 = r0, lmwords = g.lmpitch >> 2;
    if (g.lm_direct) {
        lmbase = bt.lm ? reinterpret_cast<unsigned*>(bt.lm + (long long)unit * bt.lm_stride) : nullptr;
        lmrow0 = 0; lmwords = g.lmw >> 2;
    }
    int cap0 = 0;
    // ---- pass 0 (colour 0): band rows and one halo row on each side; order from the count kernel's table.
    {
        Apply2<PixT, true> body{g, 0, own_lo, own_hi, rowcnt + (long long)unit * g.h * g.tpitch, payload, n_bits,
                                0u, p0_lo < own_lo, lmbase, lmrow0, lmwords, &st};
        if (use_pre) {  // wide rows: the prefixes built above
            body.wpre = wpre; body.rowtot = rowtot; body.pre_row0 = p0_lo;
        }
        sweep2_prime_order(g, p0_lo, p0_hi, T, body);
        PHASE_MARK(17);  // pass-0 order of the first item
        int before0 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cap0 += ccv[i];
            if (lane + 32 * i < band) before0 += ccv[i];
        }
        for (int k = lane + 128; k < g.nb; k += 32) {
            const int cc = __ldg(band_cnt + unit * g.nb + k);
            cap0 += cc;
            if (k < band) before0 += cc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            before0 += __shfl_xor_sync(0xffffffffu, before0, o);
            cap0 += __shfl_xor_sync(0xffffffffu, cap0, o);
        }
        body.B = (unsigned)before0;
        if (body.primed) body.begin_bits();
        PHASE_MARK(18);  // pass-0 order and bits of the first item
        WAIT_ROWS2_MARKED(g, s_lo, s_hi, bar, 19);
        PHASE_MARK(1);  // band copy wait
        sweep2_colour<PixT>(g, simg, r_first, 0, p0_lo, p0_hi, T, body);
        PHASE_MARK(2);  // apply 0
    }
    __syncthreads();
    PHASE_MARK(3);  // barrier

    // ---- pass 1 (colour 1) over the band rows: count, order (look-back over earlier bands), apply
    {
        Count2<PixT, false> body{g, own_lo, tab, 0};
        sweep2_colour<PixT>(g, simg, r_first, 1, own_lo, own_hi, T, body);
    }
    PHASE_MARK(4);  // count 1
    __syncthreads();
    PHASE_MARK(5);  // barrier
    if (use_pre) build_prefix<false>(g, tab, max(own_hi - own_lo, 0), wpre, rowtot);  // (pass 0 is done with them: barrier above)
    {
        const unsigned total = (unsigned)table_total(g, tab, max(own_hi - own_lo, 0));
        const unsigned before1 = warp_lookback(status + (long long)unit * g.nb, band, total, warp == 0);
        if (warp == 0 && lane == 0 && band == g.nb - 1) {
            // the last band knows both pass totals: it writes the unit's summary (no separate kernel);
            // n_flagged and sse (info[5], info[6]) are summed by every band with atomics
            const long long cap1 = (long long)before1 + total;
            info[0] = T; info[1] = n_bits; info[2] = cap0 + cap1; info[3] = cap0; info[4] = cap1;
            info[7] = ((long long)n_bits > cap0 + cap1) ? PEEB_E_CAPACITY : 0;
        }
        PHASE_MARK(6);  // look-back
        Apply2<PixT, false> body{g, own_lo, own_lo, own_hi, tab, payload, n_bits, (unsigned)cap0 + before1, false,
                                 lmbase, lmrow0, lmwords, &st};
        if (use_pre) { body.wpre = wpre; body.rowtot = rowtot; body.pre_row0 = own_lo; }
        sweep2_colour<PixT>(g, simg, r_first, 1, own_lo, own_hi, T, body);
        PHASE_MARK(7);  // apply 1
    }

    __syncthreads();
    PHASE_MARK(20);  // barrier after pass 1
    const int b_lo = r0, b_hi = min(r0 + g.R, g.h);
    if (bt.dst) store_rows2_issue<PixT>(g, bt.dst + (long long)unit * bt.dst_stride, simg, r_first, b_lo, b_hi);
    PHASE_MARK(21);  // stores issued
    {   // statistics and the location map go out while the rows drain
        const long long sse = warp_sum_i64(st.sse);
        const long long fl = warp_sum_i64((long long)st.flagged);
        if ((threadIdx.x & 31) == 0) {
            if (sse) atomicAdd(reinterpret_cast<unsigned long long*>(info + 6), (unsigned long long)sse);
            if (fl) atomicAdd(reinterpret_cast<unsigned long long*>(info + 5), (unsigned long long)fl);
            if (bt.steps) {  // optional counters of the generic-code share (peeb_pee_step_counters)
                atomicAdd(bt.steps, (unsigned long long)st.steps);
                if (st.steps_edge) atomicAdd(bt.steps + 1, (unsigned long long)st.steps_edge);
                if (st.steps_redone) atomicAdd(bt.steps + 2, (unsigned long long)st.steps_redone);
            }
        }
    }
    PHASE_MARK(8);  // stats
    if (bt.lm && !g.lm_direct) {
        unsigned char* glm = bt.lm + (long long)unit * bt.lm_stride + (size_t)b_lo * g.lmw;
        const int nrows = b_hi - b_lo;
        const int nwarps = blockDim.x >> 5;
        if ((g.lmw & 3) == 0 && ((uintptr_t)glm & 3) == 0) {
            const int wpr = g.lmw >> 2;
            for (int r = warp; r < nrows; r += nwarps)
                for (int k = lane; k < wpr; k += 32)
                    reinterpret_cast<unsigned*>(glm)[r * wpr + k] = slm[r * (g.lmpitch >> 2) + k];
        } else {
            const unsigned char* sb = reinterpret_cast<const unsigned char*>(slm);
            for (int r = warp; r < nrows; r += nwarps)
                for (int k = lane; k < g.lmw; k += 32) glm[(size_t)r * g.lmw + k] = sb[(size_t)r * g.lmpitch + k];
        }
    }
    store_rows2_wait(g);
    PHASE_MARK(9);  // location map + store drain
}

// ------------------------------------------------------------------ K_X: extract
// Carrier bits of a (row, cell) are collected MSB-first in a lane register (first carrier ends up
// in the highest of the n used bits); tn/tw tables hold count and bits per cell of the own rows.
// the location-map bytes (nbytes <= 8) of a lane's cell in its two rows, fetched once per cell ahead of use: the
// common all-zero cell then costs one test per step.  (A plain function returning values: called from the rare
// path only, and the body's state stays in registers.)
__device__ __noinline__ ulonglong2 fetch_lm_bytes(const unsigned char* p, int pitch, int nbytes) {
    ulonglong2 v = make_ulonglong2(0ull, 0ull);
#pragma unroll 1
    for (int k = 0; k < nbytes; ++k) {
        v.x |= (unsigned long long)p[k] << (8 * k);
        v.y |= (unsigned long long)p[pitch + k] << (8 * k);
    }
    return v;
}

template <typename PixT>
struct Extract2 {
    using P = PixOps<PixT>;
    const Geom2& g;
    int own_lo, own_hi;
    const unsigned char* slm;  // location-map rows, row lm_row0 first
    int lm_row0;
    unsigned char* tn;         // carriers per (own row, cell), one byte each, rows tpitch apart
    unsigned* tw;              // their bits, rows ncol words apart
    bool has_lm;               // some pixel of the band's (or its halo rows') location-map rows is flagged
    KX ka, kb;
    unsigned Wa, Wb;
    int na, nb, ia, in;
    bool sta, stb, reca, recb, primed;
    unsigned long long la, lb;  // location-map bytes of this lane's cell in rows a and b (byte k = columns 8k..8k+7 of the cell)
               // some pixel of the band's (or its halo rows') location-map rows is flagged
    __device__ __forceinline__ void begin(int rowa, int cell, bool a, bool b, bool, bool rowb_in, int T) {
        sta = a; stb = b;
        ka = kb = make_kx(T);
        Wa = Wb = 0u; na = nb = 0;
        ia = (rowa - own_lo) * g.ncol + cell;
        in = (rowa - own_lo) * g.tpitch + cell;
        reca = a && rowa >= own_lo && rowa < own_hi;
        recb = b && rowa + 1 >= own_lo && rowa + 1 < own_hi;
        la = lb = 0ull;
        if (has_lm) {  // flagged pixels are rare: most bands have none at all
            // (rows a and b of the shared copy: R + 2 rows of lmpitch bytes and the slack layout2 leaves behind them)
            // (a sweep with an odd number of rows has no row b in its last pair: row a is read twice, nothing of that
            // row b is kept.  The copy has 32 bytes of slack behind its last row for cells that stick out of the image.)
            const int to_b = rowb_in ? g.lmpitch : 0;
            BOUNDS(7, (long long)(rowa - lm_row0) * g.lmpitch + ((cell * g.CW) >> 3), to_b + g.cws * (P::PXS >> 3),
                   (long long)(g.R + 2) * g.lmpitch + 32);
            const ulonglong2 v = fetch_lm_bytes(slm + (size_t)(rowa - lm_row0) * g.lmpitch + ((cell * g.CW) >> 3), to_b,
                                                g.cws * (P::PXS >> 3));
            la = v.x; lb = v.y;
        }
    }
    __device__ __forceinline__ unsigned lmbits(unsigned long long v, int s) const {
        return (unsigned)(v >> (s * P::PXS)) & (P::PXS == 8 ? 0xffu : 0xffffu);
    }
    __device__ __forceinline__ bool item_special() const { return (la | lb) != 0ull; }
    template <int Q, bool SPECIAL>
    __device__ __forceinline__ void row(const uint4& M, uint4& O, unsigned prev, unsigned next, const uint4& U,
                                        const uint4& D, const KX& k, int c, unsigned lmb, unsigned& W, int& n) {
        static_for<0, P::NS>([&](auto Sc) {
            constexpr int S = decltype(Sc)::value;
            KX kk = k;
            if (SPECIAL) {
                constexpr int cs = 2 * S + Q;  // column inside the step
                const int col = c + cs;
                // packbits: column j of a byte at bit 7-j; 16-pixel steps read two bytes little-endian
                const unsigned flag = (lmb >> ((cs & 8) + 7 - (cs & 7))) & 1u;
                kk = make_kx((col >= 1 && col <= g.w - 2 && !flag) ? k.T : 0);
            }
            const int q = P::template qsum<Q, S>(M, prev, next, U, D, kk.init);  // 4(e' + 2T) + r, e' = x' - p
            const int cc = max(min((q + 4) >> 3, kk.T2), 0);                       // clamp(ceil(e'/2), -T, T) + T
            collect_bit(W, n, q, kk.T16);
            // the original pixel is x' - (cc - T): subtracted inside the packed word (modulo the pixel width, like the
            // specification's cast to the pixel type, so that a tampered image cannot disturb the neighbour in the word)
            P::template addwrap<Q, S>(O, kk.T - cc);
        });
    }
    template <int QA>
    __device__ __forceinline__ void step(int c, int s, bool special, const uint4& U, uint4& A, uint4& B, const uint4& D,
                                         unsigned prev, unsigned next, unsigned char* pa, unsigned char* pb) {
        const uint4 A0 = A, B0 = B;  // predictions read the words as loaded: independent chains per pixel
        if (special) {
            row<QA, true>(A0, A, prev, next, U, B0, ka, c, lmbits(la, s), Wa, na);
            row<1 - QA, true>(B0, B, prev, next, A0, D, kb, c, lmbits(lb, s), Wb, nb);
        } else {
            row<QA, false>(A0, A, prev, next, U, B0, ka, c, 0u, Wa, na);
            row<1 - QA, false>(B0, B, prev, next, A0, D, kb, c, 0u, Wb, nb);
        }
        if (sta) sts128(pa, A);
        if (stb) sts128(pb, B);
    }
    __device__ __forceinline__ void end() {
        if (reca) { BOUNDS(8, in, 1, (long long)g.R * g.tpitch); BOUNDS(8, ia, 1, (long long)g.R * g.ncol); }
        if (recb) { BOUNDS(8, in + g.tpitch, 1, (long long)g.R * g.tpitch); BOUNDS(8, ia + g.ncol, 1, (long long)g.R * g.ncol); }
        if (reca) { tn[in] = (unsigned char)na; tw[ia] = Wa; }
        if (recb) { tn[in + g.tpitch] = (unsigned char)nb; tw[ia + g.ncol] = Wb; }
    }
};

// The carrier bits of the band in payload order, both passes at once.  A thread takes one 32-bit word of a count
// table (four cells of a row; tn1 follows tn0, rows tpitch bytes apart, bytes of cells >= ncol are 0): block scan
// of the word sums in raster order -- one warp scan, one barrier, the warps of a pass never mix with the other's
// (pass 1 starts at a multiple of 32) -- then the bits of its cells (tw) go to their offsets with shared-memory
// atomics.  scan: >= 64 ints of shared memory.  Every thread gets the carriers of the band per pass.
__device__ __forceinline__ void assemble_streams(const Geom2& g, int nrows, const unsigned char* tn0, const unsigned* tw0,
                                                 const unsigned* tw1, unsigned* stream, int* scan, int& total0, int& total1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int wu = (g.ncol + 3) >> 2;             // table words per row that hold cells
    const int per = nrows * wu;                   // words per pass
    const int p1 = (per + 31) & ~31;              // first index of pass 1
    int carry0 = 0, carry1 = 0;
    int flip = 0;
    for (int base = 0; base < p1 + per; base += blockDim.x, flip ^= 32) {
        const int k = base + threadIdx.x;
        const int pass = k >= p1 ? 1 : 0;
        const int kk = k - (pass ? p1 : 0);
        const bool in = kk < per;
        int row = 0, j = 0;
        unsigned cw = 0u;
        if (in) {
            row = kk / wu; j = kk - row * wu;
            cw = *reinterpret_cast<const unsigned*>(tn0 + (size_t)(pass * g.R + row) * g.tpitch + 4 * j);
        }
        const int tot = idp4_sum(cw, 0x01010101u, 0);
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) scan[flip + warp] = incl;
        __syncthreads();
        int before = 0, all0 = 0, all1 = 0;
        for (int q = 0; q < nwarps; ++q) {
            const int a = scan[flip + q];
            const bool q1 = (base + 32 * q) >= p1;
            if (q1 == (pass != 0) && q < warp) before += a;
            if (q1) all1 += a; else all0 += a;
        }
        if (cw != 0u) {
            int o = (pass ? carry1 : carry0) + before + incl - tot;
            unsigned* out = stream + (size_t)pass * g.bandwords;
            const unsigned* tw = (pass ? tw1 : tw0) + (size_t)row * g.ncol + 4 * j;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cc = (int)((cw >> (8 * i)) & 0xffu);
                if (cc > 0) {
                    const int sh = o & 31;
                    const unsigned long long v = (unsigned long long)tw[i] << (64 - cc - sh);
                    BOUNDS(9, o >> 5, (unsigned)v ? 2 : 1, g.bandwords);
                    atomicOr(out + (o >> 5), (unsigned)(v >> 32));
                    if ((unsigned)v) atomicOr(out + (o >> 5) + 1, (unsigned)v);
                    o += cc;
                }
            }
        }
        carry0 += all0; carry1 += all1;
    }
    total0 = carry0; total1 = carry1;
}

// grid = n_units * nb.  stage_bits: per (unit, pass, band) `bandwords` words, stream bit k at bit
// 31-(k&31) of word k>>5; stage_cnt: carriers per (unit, pass, band).
template <typename PixT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) pee2_extract_kernel(Geom2 g, PeeBatch bt, unsigned* __restrict__ stage_bits,
                                                                int* __restrict__ stage_cnt, long long zero_words) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem2 L = layout2(g, 2);
    unsigned char* simg = smem_raw + L.img;
    unsigned char* slm = smem_raw + L.lm + 16;  // (16 bytes of slack in front: the unaligned copy may start a piece early)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    unsigned char* tn0 = smem_raw + L.tn0;
    unsigned char* tn1 = smem_raw + L.tn1;
    unsigned* tw0 = reinterpret_cast<unsigned*>(smem_raw + L.tw0);
    unsigned* tw1 = reinterpret_cast<unsigned*>(smem_raw + L.tw1);
    unsigned* stream = reinterpret_cast<unsigned*>(smem_raw + L.stream);  // [2][bandwords]: pass 0, pass 1

    const int unit = blockIdx.x / g.nb, band = blockIdx.x % g.nb;
    const int r0 = band * g.R, r_first = r0 - 2;
    if (g.bulk && threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    const unsigned char* usrc = bt.src + (long long)unit * bt.src_stride;
    const int T = bt.T[unit];
    const int s_lo = max(r0 - 2, 0), s_hi = min(r0 + g.R + 2, g.h);
    PHASE_INIT;
    int lm_any = 0;
    {   // location-map rows [r0-1, r0+R+1) -> slm (shared row r = image row r0-1+r, lmpitch bytes, zero padded); the
        // loads are issued before the band's bulk copies (small reads would otherwise queue behind them)
        const int l_lo = max(r0 - 1, 0), l_hi = min(r0 + g.R + 1, g.h);
        const unsigned char* glm = bt.lm + (long long)unit * bt.lm_stride;
        // bounds-checked build: global reads stay inside the unit's map (site 18; whole 16-byte pieces may reach up to
        // 15 bytes into the aligned pieces around it), shared writes inside the copy's region with its slack (site 19)
        [[maybe_unused]] const long long lm_glim = (long long)g.h * g.lmw;
        [[maybe_unused]] const unsigned char* const slm0 = slm - 16;
        [[maybe_unused]] const long long lm_slim = (long long)(g.R + 2) * g.lmpitch + 16 + 48;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        (void)lane; (void)warp; (void)nwarps;
        if (g.lmpitch == g.lmw && (g.lmw & 15) == 0 && ((uintptr_t)glm & 15) == 0) {
            // rows of whole 16-byte pieces, no padding between them: one flat copy of the rows inside the image, two
            // pieces per thread in flight (a cell that sticks out of its row then sees the first bytes of the next
            // row: columns past the image run with T = 0 whatever their flag says)
            const int npieces = (l_hi - l_lo) * (g.lmw >> 4);
            const uint4* gq = reinterpret_cast<const uint4*>(glm + (size_t)l_lo * g.lmw);
            uint4* sq = reinterpret_cast<uint4*>(slm + (size_t)(l_lo - (r0 - 1)) * g.lmw);
            BOUNDS(18, (long long)l_lo * g.lmw, 16ll * npieces, lm_glim);
            BOUNDS(19, reinterpret_cast<unsigned char*>(sq) - slm0, 16ll * npieces, lm_slim);
            BOUNDS(19, slm - slm0, 16ll * (g.R + 2) * (g.lmw >> 4), lm_slim);
            const int k0 = threadIdx.x, k1 = threadIdx.x + blockDim.x;
            const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
            const uint4 v0 = k0 < npieces ? __ldg(gq + k0) : zero4, v1 = k1 < npieces ? __ldg(gq + k1) : zero4;
            issue_rows2<PixT>(g, usrc, simg, r_first, s_lo, s_hi, bar);
            if (k0 < npieces) sq[k0] = v0;
            if (k1 < npieces) sq[k1] = v1;
            lm_any = (v0.x | v0.y | v0.z | v0.w | v1.x | v1.y | v1.z | v1.w) != 0u;
            for (int k = threadIdx.x + 2 * blockDim.x; k < npieces; k += blockDim.x) {
                const uint4 v = __ldg(gq + k);
                sq[k] = v;
                lm_any |= (v.x | v.y | v.z | v.w) != 0u;
            }
            // rows of the copy outside the image (first / last band) are read by idle lanes only: zero
            const int q = g.lmw >> 4;
            if (l_lo > r0 - 1 && (int)threadIdx.x < q) reinterpret_cast<uint4*>(slm)[threadIdx.x] = zero4;
            for (int k = (l_hi - (r0 - 1)) * q + threadIdx.x; k < (g.R + 2) * q; k += blockDim.x)
                reinterpret_cast<uint4*>(slm)[k] = zero4;
        } else if ((g.lmw & 15) == 0 && (g.lmpitch & 15) == 0 && ((uintptr_t)glm & 15) == 0) {
            // padded rows of 16-byte pieces: piece k of the shared copy -> (row, piece of the row)
            const int q = g.lmpitch >> 4, qv = g.lmw >> 4, npieces = (g.R + 2) * q;
            auto fetch = [&](int k) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                const int r = k / q, j = k - r * q, row = r0 - 1 + r;
                if (k < npieces && j < qv && row >= l_lo && row < l_hi) BOUNDS(18, (long long)row * g.lmw + 16 * j, 16, lm_glim);
                if (k < npieces && j < qv && row >= l_lo && row < l_hi)
                    v = __ldg(reinterpret_cast<const uint4*>(glm + (size_t)row * g.lmw) + j);
                return v;
            };
            const int k0 = threadIdx.x, k1 = threadIdx.x + blockDim.x;
            const uint4 v0 = fetch(k0), v1 = fetch(k1);
            issue_rows2<PixT>(g, usrc, simg, r_first, s_lo, s_hi, bar);
            uint4* sq = reinterpret_cast<uint4*>(slm);
            BOUNDS(19, slm - slm0, 16ll * npieces, lm_slim);
            if (k0 < npieces) sq[k0] = v0;
            if (k1 < npieces) sq[k1] = v1;
            lm_any = (v0.x | v0.y | v0.z | v0.w | v1.x | v1.y | v1.z | v1.w) != 0u;
            for (int k = threadIdx.x + 2 * blockDim.x; k < npieces; k += blockDim.x) {
                const uint4 v = fetch(k);
                sq[k] = v;
                lm_any |= (v.x | v.y | v.z | v.w) != 0u;
            }
        } else if (g.lmpitch == g.lmw) {
            // any row length and alignment (3000 columns: 375 bytes per row): the rows inside the image are one byte
            // range in global memory; it is copied in whole 16-byte pieces to a shared copy with the SAME alignment
            // modulo 16 (slm is moved by up to 15 bytes), rows back to back.  The few foreign bytes of the first and
            // last piece belong to neighbouring rows of the map and can only make the vote say "look" needlessly.
            const unsigned char* gfirst = glm + (size_t)l_lo * g.lmw;
            const int first = l_lo - (r0 - 1);                                   // shared row of image row l_lo
            slm += (unsigned)(((uintptr_t)gfirst - (uintptr_t)first * g.lmw) & 15);   // S(r) = slm + r * lmw, S(first) = gfirst mod 16
            const int head = (int)((uintptr_t)gfirst & 15);
            const int nbytes = (l_hi - l_lo) * g.lmw;
            const int npieces = nbytes > 0 ? (head + nbytes + 15) >> 4 : 0;
            const uint4* gq = reinterpret_cast<const uint4*>(gfirst - head);
            uint4* sq = reinterpret_cast<uint4*>(slm + (size_t)first * g.lmw - head);
            BOUNDS(18, (long long)l_lo * g.lmw - head + 15, 16ll * npieces - 15 - (npieces ? 15 : 0), lm_glim);
            BOUNDS(19, reinterpret_cast<unsigned char*>(sq) - slm0, 16ll * npieces, lm_slim);
            BOUNDS(19, slm - slm0, (long long)(g.R + 2) * g.lmw, lm_slim);
            const int k0 = threadIdx.x, k1 = threadIdx.x + blockDim.x;
            const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
            const uint4 v0 = k0 < npieces ? __ldg(gq + k0) : zero4, v1 = k1 < npieces ? __ldg(gq + k1) : zero4;
            issue_rows2<PixT>(g, usrc, simg, r_first, s_lo, s_hi, bar);
            if (k0 < npieces) sq[k0] = v0;
            if (k1 < npieces) sq[k1] = v1;
            lm_any = (v0.x | v0.y | v0.z | v0.w | v1.x | v1.y | v1.z | v1.w) != 0u;
            for (int k = threadIdx.x + 2 * blockDim.x; k < npieces; k += blockDim.x) {
                const uint4 v = __ldg(gq + k);
                sq[k] = v;
                lm_any |= (v.x | v.y | v.z | v.w) != 0u;
            }
            // rows of the copy outside the image (first / last band): read by idle lanes only, zero.  (Bytes, behind
            // a barrier-free guarantee: they do not overlap the pieces above except in the pieces' foreign bytes,
            // which nobody needs.)
            if (first > 0)
                for (int k = threadIdx.x; k < first * g.lmw - head; k += blockDim.x) slm[k] = 0;
            for (int k = (l_hi - (r0 - 1)) * g.lmw + 16 + (int)threadIdx.x; k < (g.R + 2) * g.lmw; k += blockDim.x) slm[k] = 0;
        } else {
            issue_rows2<PixT>(g, usrc, simg, r_first, s_lo, s_hi, bar);
            for (int r = warp; r < g.R + 2; r += nwarps) {
                const int row = r0 - 1 + r;
                const bool valid = row >= l_lo && row < l_hi;
                for (int k = lane; k < g.lmpitch; k += 32) {
                    if (valid && k < g.lmw) BOUNDS(18, (long long)row * g.lmw + k, 1, lm_glim);
                    BOUNDS(19, (slm - slm0) + (long long)r * g.lmpitch + k, 1, lm_slim);
                    const unsigned char v = (valid && k < g.lmw) ? glm[(size_t)row * g.lmw + k] : (unsigned char)0;
                    slm[(size_t)r * g.lmpitch + k] = v;
                    lm_any |= v != 0;
                }
            }
        }
        PHASE_MARK(24);  // location map + copies issued
        // the band streams (assembled with atomics) and the count tables (tn0 and tn1 are adjacent; the bytes
        // past ncol are summed with the rest) start from zero: 16-byte stores
        {
            uint4* z = reinterpret_cast<uint4*>(stream);
            const int nz = (int)(align_up((size_t)2 * g.bandwords * sizeof(unsigned), 16) >> 4);
            for (int k = threadIdx.x; k < nz; k += blockDim.x) z[k] = make_uint4(0u, 0u, 0u, 0u);
            uint4* t = reinterpret_cast<uint4*>(tn0);
            for (int k = threadIdx.x; k < (2 * g.R * g.tpitch) >> 4; k += blockDim.x) t[k] = make_uint4(0u, 0u, 0u, 0u);
        }
        // the unit's output words start from zero (the gather kernel ORs the boundary words in and skips
        // zero words): every band clears its share, instead of a memset on the stream
        unsigned* pout = reinterpret_cast<unsigned*>(bt.payload_out + (long long)unit * bt.payload_stride);
#ifdef PEEB_AB_ZWORD
        if (false) {
#else
        if ((zero_words & 3) == 0 && ((uintptr_t)pout & 15) == 0) {
#endif
            const int nq = (int)(zero_words >> 2);
            const int q_lo = (int)((long long)nq * band / g.nb), q_hi = (int)((long long)nq * (band + 1) / g.nb);
            uint4* pq = reinterpret_cast<uint4*>(pout);
            for (int k = q_lo + (int)threadIdx.x; k < q_hi; k += blockDim.x) pq[k] = make_uint4(0u, 0u, 0u, 0u);
        } else {
            const long long z_lo = zero_words * band / g.nb, z_hi = zero_words * (band + 1) / g.nb;
            for (long long k = z_lo + threadIdx.x; k < z_hi; k += blockDim.x) pout[k] = 0u;
        }
    }
    PHASE_MARK(25);  // tables cleared
    if (g.bulk && s_hi > s_lo && POLL_THREAD) mbar_wait(bar, 0);  // (see wait_rows2)
    PHASE_MARK(26);
    const bool has_lm = __syncthreads_or(lm_any) != 0;  // (also the barrier that makes the staged rows visible)
    PHASE_MARK(0);  // load
    const int own_lo = max(r0, 1), own_hi = min(r0 + g.R, g.h - 1);
    const int p1_lo = max(r0 - 1, 1), p1_hi = min(r0 + g.R + 1, g.h - 1);
    const int nrows = max(own_hi - own_lo, 0);
    {   // colour 1 first (band rows + one halo row each side), then colour 0 (band rows)
        Extract2<PixT> body{g, own_lo, own_hi, slm, r0 - 1, tn1, tw1, has_lm};
        sweep2_colour<PixT>(g, simg, r_first, 1, p1_lo, p1_hi, T, body);
        PHASE_MARK(1);  // sweep colour 1
    }
    __syncthreads();
    PHASE_MARK(2);
    {
        Extract2<PixT> body{g, own_lo, own_hi, slm, r0 - 1, tn0, tw0, has_lm};
        sweep2_colour<PixT>(g, simg, r_first, 0, own_lo, own_hi, T, body);
        PHASE_MARK(3);  // sweep colour 0
    }
    __syncthreads();
    PHASE_MARK(4);

    // the recovered rows leave while the carrier bits are assembled
    if (bt.dst)
        store_rows2_issue<PixT>(g, bt.dst + (long long)unit * bt.dst_stride, simg, r_first, r0, min(r0 + g.R, g.h));
    PHASE_MARK(27);  // stores issued
    int total0, total1;
    assemble_streams(g, nrows, tn0, tw0, tw1, stream, reinterpret_cast<int*>(smem_raw + L.misc), total0, total1);
    PHASE_MARK(28);  // assembled
    __syncthreads();
    PHASE_MARK(29);  // barrier
    for (int pass = 0; pass < 2; ++pass) {
        const int total = pass ? total1 : total0;
        const long long slot = ((long long)unit * 2 + pass) * g.nb + band;
        if (threadIdx.x == 0) stage_cnt[slot] = total;
        unsigned* gout = stage_bits + slot * g.bandwords;
        const unsigned* out = stream + (size_t)pass * g.bandwords;
        const int nw = (total + 31) >> 5;
        BOUNDS(10, 0, nw, g.bandwords);
        for (int k = threadIdx.x; k < nw; k += blockDim.x) gout[k] = out[k];
    }
    PHASE_MARK(5);  // stream assembly + staging writes
    store_rows2_wait(g);
    PHASE_MARK(6);  // store
}

// ------------------------------------------------------------------ K_G: payload assembly
// grid = (blocks per unit, n_units); a block takes `ppb` consecutive pieces of a unit.  Piece p of a unit =
// pass-0 band p (p < nb) or pass-1 band p-nb; its global bit offset is the sum of the earlier pieces'
// counts.  Output is MSB-first packed, truncated to n_bits; the output words were cleared by K_X.
constexpr int GATHER_PPB = 32;
__global__ void __launch_bounds__(256) pee2_gather_kernel(int nb, int bandwords, int ppb, PeeBatch bt,
                                                          const unsigned* __restrict__ stage_bits,
                                                          const int* __restrict__ stage_cnt) {
    const int unit = blockIdx.y, p0 = blockIdx.x * ppb, p1 = min(p0 + ppb, 2 * nb);
    const int* cnts = stage_cnt + (long long)unit * 2 * nb;
    __shared__ long long s_tot[3];
    __shared__ int s_cnt[GATHER_PPB];
    if (threadIdx.x < 32) {
        long long before = 0, all = 0, c0 = 0;
        for (int k = threadIdx.x; k < 2 * nb; k += 32) {
            const int c = cnts[k];
            all += c;
            if (k < p0) before += c;
            if (k < nb) c0 += c;
        }
        before = warp_sum_i64(before); all = warp_sum_i64(all); c0 = warp_sum_i64(c0);
        if (threadIdx.x == 0) { s_tot[0] = before; s_tot[1] = all; s_tot[2] = c0; }
    } else if (threadIdx.x < 32 + GATHER_PPB) {
        const int k = p0 + (int)threadIdx.x - 32;
        s_cnt[threadIdx.x - 32] = k < p1 ? cnts[k] : 0;
    }
    __syncthreads();
    long long before = s_tot[0];
    const long long all = s_tot[1], n_bits = bt.n_bits[unit];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long* info = bt.info + (long long)unit * PEEB_INFO;
        info[0] = bt.T[unit]; info[1] = n_bits; info[2] = all; info[3] = s_tot[2]; info[4] = all - s_tot[2];
        info[5] = 0; info[6] = 0; info[7] = n_bits > all ? PEEB_E_CAPACITY : 0;
    }
    unsigned* out = reinterpret_cast<unsigned*>(bt.payload_out + (long long)unit * bt.payload_stride);
    // (a warp per piece instead of the whole block per piece measured slower on the 512-slice batch: 21 against 18 us)
    for (int piece = p0; piece < p1; ++piece) {
        const int cnt = s_cnt[piece - p0];
        if (cnt == 0) continue;
        if (before >= n_bits) break;
        const unsigned* src = stage_bits + ((long long)unit * 2 * nb + piece) * bandwords;
        const int nsrc = (cnt + 31) >> 5;
        const long long first = before >> 5, last = (before + cnt - 1) >> 5;
        const int sh = (int)(before & 31);
        for (long long mw = first + threadIdx.x; mw <= last; mw += blockDim.x) {
            const int i = (int)(mw - first);
            BOUNDS(16, nsrc, 0, bandwords + 1);
            const unsigned cur = i < nsrc ? src[i] : 0u;
            const unsigned prev = (i >= 1 && i - 1 < nsrc) ? src[i - 1] : 0u;
            unsigned val = __funnelshift_r(cur, prev, sh);  // stream bits [32 mw, 32 mw + 32), first bit on top
            const long long bit0 = mw << 5;
            if (bit0 + 32 > n_bits) {  // drop bits at or past n_bits
                const int keep = (int)(n_bits - bit0);
                val = keep <= 0 ? 0u : (val & ~(0xffffffffu >> keep));
            }
            if (val == 0) continue;
            const unsigned packed = __byte_perm(val, 0, 0x0123);
            BOUNDS(17, 4 * mw, 4, 4 * ((n_bits + 31) >> 5));  // never past the words that hold the unit's n_bits
            if (mw == first || mw == last) atomicOr(out + mw, packed);
            else out[mw] = packed;
        }
        before += cnt;
    }
}

// ------------------------------------------------------------------ small-image path: one thread-block cluster per image
// A single slice (or a handful) cannot fill the GPU with the band kernels above, and their order of events -- count
// kernel, embed kernel with a look-back chain through global memory, extract kernel, gather kernel -- is four
// launches of mostly latency.  Here one CLUSTER of up to 16 CTAs owns an image: CTA k stages band k in its shared
// memory as before, and what the bands must tell each other (carriers per pass, statistics) goes through
// distributed shared memory: every CTA writes its word into the exchange array of every other CTA
// (st.shared::cluster) and one cluster barrier later everybody knows all of them -- no status words, no tickets, no
// pre-zeroed counters, no second kernel.  Embed = 1 launch (count, scan, apply for both passes, summary written by
// CTA 0), extract = 1 launch (the bands write their bits straight to their place in the unit's payload).
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// word `idx` of the exchange array of CTA `cta` of this cluster (same offset as in this CTA's own shared memory)
__device__ __forceinline__ void st_cluster(unsigned* own_addr, unsigned cta, unsigned v) {
    unsigned remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(own_addr)), "r"(cta));
    asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(remote), "r"(v) : "memory");
}
// every CTA of the cluster publishes `v` in slot [row][its rank] of everybody's exchange array (a warp does it, one
// destination per lane); the caller runs a cluster barrier before reading
__device__ __forceinline__ void cluster_publish(unsigned* xch, int row, unsigned rank, int C, unsigned v) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32 && lane < C) st_cluster(xch + row * CLUSTER_MAX + rank, (unsigned)lane, v);
}
__device__ __forceinline__ void cluster_sums(const unsigned* xch, int row, int rank, int C, unsigned& before, unsigned& all) {
    before = 0u; all = 0u;
    for (int k = 0; k < C; ++k) {
        const unsigned v = xch[row * CLUSTER_MAX + k];
        if (k < rank) before += v;
        all += v;
    }
}

template <typename PixT, int NT>
__global__ void __launch_bounds__(NT, 1) pee2_cluster_embed_kernel(Geom2 g, PeeBatch bt, int C) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem2 L = layout2(g, 4);
    unsigned char* simg = smem_raw + L.img;
    unsigned char* tab1 = smem_raw + L.tab;
    unsigned char* tab0 = smem_raw + L.tab0;
    unsigned* xch = reinterpret_cast<unsigned*>(smem_raw + L.xch);
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    const int lane = threadIdx.x & 31;
    const int band = (int)cluster_rank(), unit = blockIdx.x / C;
    PHASE_INIT;
    cluster_arrive();  // (waited for before the first remote store: every CTA of the cluster is then running)
    if (threadIdx.x == 0) {
        misc[0] = misc[1] = misc[2] = misc[3] = 0;
        if (g.bulk) { mbar_init(bar, 1); fence_mbar_init(); }
    }
    for (int k = threadIdx.x; k < ((g.R + 2) * g.tpitch) >> 4; k += blockDim.x) {
        reinterpret_cast<uint4*>(tab0)[k] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(tab1)[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    // (CTAs past the last band of the image own no rows: they publish zeros and take part in the barriers)
    const int r0 = min(band * g.R, g.h), r_first = r0 - 2;
    const int r1 = min(r0 + g.R, g.h);
    const unsigned char* usrc = bt.src + (long long)unit * bt.src_stride;
    const int T = bt.T[unit];
    const unsigned n_bits = bt.n_bits[unit];
    const unsigned* payload = reinterpret_cast<const unsigned*>(bt.payload + (long long)unit * bt.payload_stride);
    long long* info = bt.info + (long long)unit * PEEB_INFO;
    const bool live = r1 > r0;
    const int own_lo = max(r0, 1), own_hi = live ? min(r1, g.h - 1) : own_lo;
    const int p0_lo = max(r0 - 1, 1), p0_hi = live ? min(r1 + 1, g.h - 1) : p0_lo;
    const int s_lo = max(r0 - 2, 0), s_hi = live ? min(r1 + 2, g.h) : s_lo;
    issue_rows2<PixT>(g, usrc, simg, r_first, s_lo, s_hi, bar);
    unsigned* lmbase = bt.lm ? reinterpret_cast<unsigned*>(bt.lm + (long long)unit * bt.lm_stride) : nullptr;
    const int lmwords = g.lmw >> 2;
    if (lmbase)   // flagged pixels are rare: their bits are OR-ed straight into the (zeroed) global rows of this band
        for (int k = threadIdx.x; k < (r1 - r0) * lmwords; k += blockDim.x) lmbase[(size_t)r0 * lmwords + k] = 0u;
    PHASE_MARK(0);  // set-up
    wait_rows2(g, s_lo, s_hi, bar);
    PHASE_MARK(1);  // band copy wait
    Stats2 st;
    // ---- pass 0: carriers per (row, cell) of the band and its halo rows, the cluster's scan, apply
    {
        Count2<PixT, false> body{g, p0_lo, tab0, 0};
        sweep2_colour<PixT>(g, simg, r_first, 0, p0_lo, p0_hi, T, body);
    }
    __syncthreads();
    PHASE_MARK(2);  // count 0
    const unsigned total0 = (unsigned)table_total(g, tab0 + (size_t)(own_lo - p0_lo) * g.tpitch, max(own_hi - own_lo, 0));
    cluster_wait();
    cluster_publish(xch, 0, (unsigned)band, C, total0);
    cluster_arrive(); cluster_wait();
    unsigned before0, cap0;
    cluster_sums(xch, 0, band, C, before0, cap0);
    PHASE_MARK(3);  // exchange 0
    {
        Apply2<PixT, false> body{g, p0_lo, own_lo, own_hi, tab0, payload, n_bits, before0, p0_lo < own_lo, lmbase, 0, lmwords, &st};
        sweep2_colour<PixT>(g, simg, r_first, 0, p0_lo, p0_hi, T, body);
    }
    __syncthreads();
    PHASE_MARK(4);  // apply 0
    // ---- pass 1 over the band rows
    {
        Count2<PixT, false> body{g, own_lo, tab1, 0};
        sweep2_colour<PixT>(g, simg, r_first, 1, own_lo, own_hi, T, body);
    }
    __syncthreads();
    PHASE_MARK(5);  // count 1
    const unsigned total1 = (unsigned)table_total(g, tab1, max(own_hi - own_lo, 0));
    cluster_publish(xch, 1, (unsigned)band, C, total1);
    cluster_arrive(); cluster_wait();
    unsigned before1, cap1;
    cluster_sums(xch, 1, band, C, before1, cap1);
    PHASE_MARK(6);  // exchange 1
    {
        Apply2<PixT, false> body{g, own_lo, own_lo, own_hi, tab1, payload, n_bits, cap0 + before1, false, lmbase, 0, lmwords, &st};
        sweep2_colour<PixT>(g, simg, r_first, 1, own_lo, own_hi, T, body);
    }
    __syncthreads();
    PHASE_MARK(7);  // apply 1
    if (bt.dst) store_rows2_issue<PixT>(g, bt.dst + (long long)unit * bt.dst_stride, simg, r_first, r0, r1);
    {   // statistics: per CTA in shared memory, then to CTA 0 of the cluster, which writes the unit's summary
        const long long sse = warp_sum_i64(st.sse);
        const unsigned fl = (unsigned)warp_sum_i64((long long)st.flagged);
        if (lane == 0) {
            if (sse) atomicAdd(reinterpret_cast<unsigned long long*>(misc), (unsigned long long)sse);
            if (fl) atomicAdd(reinterpret_cast<unsigned*>(misc) + 2, fl);
            if (bt.steps) {
                atomicAdd(bt.steps, (unsigned long long)st.steps);
                if (st.steps_edge) atomicAdd(bt.steps + 1, (unsigned long long)st.steps_edge);
                if (st.steps_redone) atomicAdd(bt.steps + 2, (unsigned long long)st.steps_redone);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            st_cluster(xch + 2 * CLUSTER_MAX + band, 0u, (unsigned)misc[0]);
            st_cluster(xch + 3 * CLUSTER_MAX + band, 0u, (unsigned)misc[1]);
            st_cluster(xch + 4 * CLUSTER_MAX + band, 0u, (unsigned)misc[2]);
        }
        cluster_arrive(); cluster_wait();
        if (band == 0 && threadIdx.x == 0) {
            unsigned long long sse_all = 0, fl_all = 0;
            for (int k = 0; k < C; ++k) {
                sse_all += (unsigned long long)xch[2 * CLUSTER_MAX + k] | ((unsigned long long)xch[3 * CLUSTER_MAX + k] << 32);
                fl_all += xch[4 * CLUSTER_MAX + k];
            }
            const long long cap = (long long)cap0 + cap1;
            info[0] = T; info[1] = n_bits; info[2] = cap; info[3] = cap0; info[4] = cap1;
            info[5] = (long long)fl_all; info[6] = (long long)sse_all;
            info[7] = ((long long)n_bits > cap) ? PEEB_E_CAPACITY : 0;
        }
    }
    PHASE_MARK(8);  // statistics + summary
    store_rows2_wait(g);
    PHASE_MARK(9);  // store drain
}

template <typename PixT, int NT>
__global__ void __launch_bounds__(NT, 1) pee2_cluster_extract_kernel(Geom2 g, PeeBatch bt, int C, long long zero_words) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem2 L = layout2(g, 5);
    unsigned char* simg = smem_raw + L.img;
    unsigned char* slm = smem_raw + L.lm;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
    unsigned char* tn0 = smem_raw + L.tn0;
    unsigned char* tn1 = smem_raw + L.tn1;
    unsigned* tw0 = reinterpret_cast<unsigned*>(smem_raw + L.tw0);
    unsigned* tw1 = reinterpret_cast<unsigned*>(smem_raw + L.tw1);
    unsigned* stream = reinterpret_cast<unsigned*>(smem_raw + L.stream);
    unsigned* xch = reinterpret_cast<unsigned*>(smem_raw + L.xch);
    const int band = (int)cluster_rank(), unit = blockIdx.x / C;
    cluster_arrive();
    if (g.bulk && threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    const int r0 = min(band * g.R, g.h), r_first = r0 - 2, r1 = min(r0 + g.R, g.h);
    const bool live = r1 > r0;
    const unsigned char* usrc = bt.src + (long long)unit * bt.src_stride;
    const int T = bt.T[unit];
    const long long n_bits = bt.n_bits[unit];
    const int s_lo = max(r0 - 2, 0), s_hi = live ? min(r1 + 2, g.h) : s_lo;
    int lm_any = 0;
    {
        const int l_lo = max(r0 - 1, 0), l_hi = live ? min(r1 + 1, g.h) : l_lo;
        const unsigned char* glm = bt.lm + (long long)unit * bt.lm_stride;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        issue_rows2<PixT>(g, usrc, simg, r_first, s_lo, s_hi, bar);
        if ((g.lmw & 3) == 0 && (g.lmpitch & 3) == 0 && ((uintptr_t)glm & 3) == 0) {
            const int wpr = g.lmw >> 2, wpp = g.lmpitch >> 2;
            for (int r = warp; r < g.R + 2; r += nwarps) {
                const int row = r0 - 1 + r;
                const bool valid = row >= l_lo && row < l_hi;
                for (int k = lane; k < wpp; k += 32) {
                    const unsigned v = (valid && k < wpr) ? __ldg(reinterpret_cast<const unsigned*>(glm + (size_t)row * g.lmw) + k) : 0u;
                    reinterpret_cast<unsigned*>(slm + (size_t)r * g.lmpitch)[k] = v;
                    lm_any |= v != 0u;
                }
            }
        } else {
            for (int r = warp; r < g.R + 2; r += nwarps) {
                const int row = r0 - 1 + r;
                const bool valid = row >= l_lo && row < l_hi;
                for (int k = lane; k < g.lmpitch; k += 32) {
                    const unsigned char v = (valid && k < g.lmw) ? glm[(size_t)row * g.lmw + k] : (unsigned char)0;
                    slm[(size_t)r * g.lmpitch + k] = v;
                    lm_any |= v != 0;
                }
            }
        }
        uint4* z = reinterpret_cast<uint4*>(stream);
        const int nz = (int)(align_up((size_t)2 * g.bandwords * sizeof(unsigned), 16) >> 4);
        for (int k = threadIdx.x; k < nz; k += blockDim.x) z[k] = make_uint4(0u, 0u, 0u, 0u);
        uint4* t = reinterpret_cast<uint4*>(tn0);
        for (int k = threadIdx.x; k < (2 * g.R * g.tpitch) >> 4; k += blockDim.x) t[k] = make_uint4(0u, 0u, 0u, 0u);
        // the unit's output words start from zero: every CTA of the cluster clears its share (the cluster barrier
        // before the bits are written orders these stores before the other CTAs' atomics on boundary words)
        unsigned* pout = reinterpret_cast<unsigned*>(bt.payload_out + (long long)unit * bt.payload_stride);
        const long long z_lo = zero_words * band / C, z_hi = zero_words * (band + 1) / C;
        for (long long k = z_lo + threadIdx.x; k < z_hi; k += blockDim.x) pout[k] = 0u;
    }
    if (g.bulk && s_hi > s_lo && POLL_THREAD) mbar_wait(bar, 0);
    const bool has_lm = __syncthreads_or(lm_any) != 0;
    const int own_lo = max(r0, 1), own_hi = live ? min(r1, g.h - 1) : own_lo;
    const int p1_lo = max(r0 - 1, 1), p1_hi = live ? min(r1 + 1, g.h - 1) : p1_lo;
    const int nrows = max(own_hi - own_lo, 0);
    {
        Extract2<PixT> body{g, own_lo, own_hi, slm, r0 - 1, tn1, tw1, has_lm};
        sweep2_colour<PixT>(g, simg, r_first, 1, p1_lo, p1_hi, T, body);
    }
    __syncthreads();
    {
        Extract2<PixT> body{g, own_lo, own_hi, slm, r0 - 1, tn0, tw0, has_lm};
        sweep2_colour<PixT>(g, simg, r_first, 0, own_lo, own_hi, T, body);
    }
    __syncthreads();
    if (bt.dst) store_rows2_issue<PixT>(g, bt.dst + (long long)unit * bt.dst_stride, simg, r_first, r0, r1);
    int total0, total1;
    assemble_streams(g, nrows, tn0, tw0, tw1, stream, reinterpret_cast<int*>(smem_raw + L.misc), total0, total1);
    __syncthreads();
    cluster_wait();
    cluster_publish(xch, 0, (unsigned)band, C, (unsigned)total0);
    cluster_publish(xch, 1, (unsigned)band, C, (unsigned)total1);
    cluster_arrive(); cluster_wait();
    unsigned b0, all0, b1, all1;
    cluster_sums(xch, 0, band, C, b0, all0);
    cluster_sums(xch, 1, band, C, b1, all1);
    const long long all = (long long)all0 + all1;
    if (band == 0 && threadIdx.x == 0) {
        long long* info = bt.info + (long long)unit * PEEB_INFO;
        info[0] = T; info[1] = n_bits; info[2] = all; info[3] = all0; info[4] = all1;
        info[5] = 0; info[6] = 0; info[7] = n_bits > all ? PEEB_E_CAPACITY : 0;
    }
    // the band's two streams go straight to their place in the unit's payload (MSB-first bytes, cut at n_bits)
    unsigned* out = reinterpret_cast<unsigned*>(bt.payload_out + (long long)unit * bt.payload_stride);
    for (int pass = 0; pass < 2; ++pass) {
        const int cnt = pass ? total1 : total0;
        const long long before = pass ? (long long)all0 + b1 : (long long)b0;
        if (cnt == 0 || before >= n_bits) continue;
        const unsigned* src = stream + (size_t)pass * g.bandwords;
        const int nsrc = (cnt + 31) >> 5;
        const long long first = before >> 5, last = (before + cnt - 1) >> 5;
        const int sh = (int)(before & 31);
        for (long long mw = first + threadIdx.x; mw <= last; mw += blockDim.x) {
            const int i = (int)(mw - first);
            const unsigned cur = i < nsrc ? src[i] : 0u;
            const unsigned prev = (i >= 1 && i - 1 < nsrc) ? src[i - 1] : 0u;
            unsigned val = __funnelshift_r(cur, prev, sh);
            const long long bit0 = mw << 5;
            if (bit0 + 32 > n_bits) {
                const int keep = (int)(n_bits - bit0);
                val = keep <= 0 ? 0u : (val & ~(0xffffffffu >> keep));
            }
            if (val == 0) continue;
            const unsigned packed = __byte_perm(val, 0, 0x0123);
            if (mw == first || mw == last) atomicOr(out + mw, packed);
            else out[mw] = packed;
        }
    }
    store_rows2_wait(g);
}

// ------------------------------------------------------------------ host side
static int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) ++l; return l; }

static int make_geom2(peeb_ws* ws, int h, int w, int itemsize, int bit_depth, int kind, bool lm_direct_ok, Geom2& g) {
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "pee: itemsize must be 1 or 2");
    PEEB_REQUIRE(bit_depth >= 1 && bit_depth <= 8 * itemsize, "pee: bit_depth %d out of range for itemsize %d", bit_depth, itemsize);
    PEEB_REQUIRE(h >= 1 && w >= 1 && (long long)h * w < (1ll << 31), "pee: image size %dx%d unsupported", h, w);
    g.h = h; g.w = w; g.itemsize = itemsize;
    g.maxval = (1 << bit_depth) - 1;
    g.rowbytes = w * itemsize;
    g.bulk = ws->use_bulk && (g.rowbytes % 16 == 0);
    g.pitch = (int)align_up((size_t)g.rowbytes, 128) + 16;
    g.lmw = (w + 7) / 8;
    // extract copies rows of whole 16-byte pieces flat, without padding (they are read once per cell: bank conflicts do
    // not matter); otherwise 12 bytes of pad spread the rows of the lanes of a warp over the banks
    // (one 16-byte piece of padding per row measured 6 % faster on the 512-slice batch than rows back to back:
    // scripts/ab_variants.sh, gpurun_out/r02_ab11.log; PEEB_LM_PAD overrides it for such runs)
    g.lmpitch = (kind == 2 && (g.lmw & 15) == 0) ? g.lmw + (getenv("PEEB_LM_PAD") ? atoi(getenv("PEEB_LM_PAD")) : 16)
                : (kind == 2 && !getenv("PEEB_LM_BYTES")) ? g.lmw   // rows back to back: one unaligned range, copied in 16-byte pieces
                                                          : (int)align_up((size_t)g.lmw, 4) + 12;
    g.lm_direct = (kind == 1 && lm_direct_ok && (g.lmw & 3) == 0 && !getenv("PEEB_LM_SHARED")) ? 1 : 0;
    const int pxs = 16 / itemsize;
    const int nsteps = (g.rowbytes + 15) / 16;
    const size_t sm_total = (size_t)ws->max_smem_optin + 1024;  // 227 KB usable + 1 KB reserved per CTA
    // Candidates: band heights R with R + 2 = 2 * rpw (a pass over the band and its halo rows fills
    // whole warp items), CTA sizes, cell widths; score = modelled lane-steps per useful pixel.
    double best = 1e30;
    Geom2 bestg = g;
    bool found = false;
    const int forceR = getenv("PEEB_BAND_R") ? atoi(getenv("PEEB_BAND_R")) : 0;
    const int forceT = getenv("PEEB_CTA_THREADS") ? atoi(getenv("PEEB_CTA_THREADS")) : 0;
    const int forceCW = getenv("PEEB_CELL_W") ? atoi(getenv("PEEB_CELL_W")) : 0;
    const int forceB = getenv("PEEB_CTA_MINB") ? atoi(getenv("PEEB_CTA_MINB")) : 0;
    const bool old_model = getenv("PEEB_MODEL_R01") != nullptr;  // A/B runs against the first round's fit
    for (int R : {62, 58, 54, 30, 26, 14, 12, 10, 6, 2}) {
        if (forceR && R != forceR) continue;
        Geom2 t = g;
        t.R = R;
        t.rpw_log2 = ilog2((R + 2) / 2);
        if ((1 << t.rpw_log2) < (R + 2) / 2) ++t.rpw_log2;
        t.rpw = 1 << t.rpw_log2;
        const int parts = 32 / t.rpw;
        for (int cws = 64 / pxs; cws >= 1; --cws) {
            t.cws = cws; t.CW = cws * pxs;
            if (forceCW && t.CW != forceCW) continue;
            t.ncol = (nsteps + cws - 1) / cws;
            // a shared row holds whole cells: the steps of the last cell that lie past the end of the image row
            // read and store back padding of their own row, never the start of the next one
            t.pitch = (int)align_up((size_t)std::max(g.rowbytes, t.ncol * t.CW * itemsize), 128) + 16;
            t.tpitch = (int)align_up((size_t)t.ncol, 16);
            t.nic = (t.ncol + parts - 1) / parts;
            t.bandwords = (R * ((w + 1) / 2) + 31) / 32 + 2;
            const size_t smem = layout2(t, kind).total;
            if (smem > (size_t)ws->max_smem_optin) continue;
            const int nic = t.nic;  // warp items of a sweep
            for (int threads : {128, 256, 512, 1024}) {
                if (forceT ? threads != forceT : threads == 128) continue;  // 128-thread CTAs: experiments only (PEEB_CTA_THREADS)
                const int nwarps = threads / 32;
                int cps = (int)(sm_total / (smem + 1024));
                cps = std::min(cps, 2048 / threads);
                if (cps < 1) continue;
                // registers: 64K per SM; the row-pair step wants ~80 per thread
                const int regs = 65536 / (cps * threads);
                if (regs < 64) cps = std::max(1, 65536 / (64 * threads));
                if (forceB) cps = std::min(cps, forceB);
                const int rounds = (nic + nwarps - 1) / nwarps;
                // Model (fitted to B200 runs): lane-steps issued per useful row-step
                // (idle lanes, halo rows, unbalanced rounds), divided by how well the resident warps hide
                // the per-band latencies (24 warps per SM ~ saturation), with a penalty when the
                // register budget per thread falls under what the row-pair step needs without spilling.
                const double work = (double)rounds * nwarps * 32 * (cws + 2.5) * 2;  // + per-item set-up, ~2.5 steps
                // per-band latencies (tables, look-back, store) in lane-steps per pass; the extract kernel has no
                // look-back and no payload fetches, its bands can be smaller
                const double fixed = (kind == 2 && !old_model) ? 2048.0 : 4096.0;
                const double useful = (double)R * nsteps;
                const double warps_sm = (double)cps * threads / 32.0;
                // fewer, larger CTAs hide each other's barriers and latencies less well than three small ones; a single
                // CTA per SM overlaps nothing of its load and store phases (3000-wide radiographs, extract: 6-row bands
                // with three 256-thread CTAs per SM measured 1.12 ms against 1.33 ms for one 1024-thread CTA with 26 rows)
                const double occ = std::pow(std::min(1.0, warps_sm / 24.0), 0.7) *
                                   (cps >= 3 ? 1.0 : (cps == 1 && kind == 2 && !old_model) ? 0.6 : 0.85);
                // the embed step spills under ~72 registers per thread; count and extract fit 64
                const double regpen = (kind == 1 && 65536 / (cps * threads) < 72) ? 1.45 : 1.0;
                const double score = (work + fixed) / useful / occ * regpen;
                if (score < best) {
                    best = score; bestg = t; bestg.threads = threads; bestg.minb = cps; found = true;
                }
            }
        }
    }
    if (!found) {
        set_error("pee: image width %d needs more shared memory than one SM has", w);
        return PEEB_E_UNSUPPORTED;
    }
    g = bestg;
    g.nb = (h + g.R - 1) / g.R;
    g.R = (h + g.nb - 1) / g.nb;  // same number of bands, rows spread evenly
    g.bandwords = (g.R * ((w + 1) / 2) + 31) / 32 + 2;
    if (getenv("PEEB_DEBUG_GEOM")) {
        static int printed[4] = {0, 0, 0, 0};
        if (printed[kind]++ < 1)
            fprintf(stderr, "[peeb] %dx%dx%d kind %d: R=%d rpw=%d CW=%d ncol=%d threads=%d minb=%d smem=%zu score=%.3f\n", h, w, itemsize,
                    kind, g.R, g.rpw, g.CW, g.ncol, g.threads, g.minb, layout2(g, kind).total, best);
    }
    return PEEB_OK;
}

template <typename K>
static int set_smem2(K kernel, size_t bytes) {
    PEEB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PEEB_OK;
}

template <typename PixT, int NT, int MINB>
static int launch_embed2(peeb_ws* ws, const Geom2& g, const PeeBatch& bt, long long nbands, int* band_cnt,
                         unsigned char* rowcnt, unsigned* ticket, unsigned long long* status, cudaStream_t st) {
    const size_t smem0 = layout2(g, 0).total, smem1 = layout2(g, 1).total;
    int rc = set_smem2(pee2_count_kernel<PixT, NT, MINB>, smem0); if (rc) return rc;
    rc = set_smem2(pee2_embed_kernel<PixT, NT, MINB>, smem1); if (rc) return rc;
    { ProfScope p(ws, PEEB_K_PEE_COUNT, st);
      pee2_count_kernel<PixT, NT, MINB><<<(unsigned)nbands, NT, smem0, st>>>(g, bt, band_cnt, rowcnt, ticket, status); }
    { ProfScope p(ws, PEEB_K_PEE_EMBED, st);
      pee2_embed_kernel<PixT, NT, MINB><<<(unsigned)nbands, NT, smem1, st>>>(g, bt, band_cnt, rowcnt, ticket, status); }
    return PEEB_OK;
}
template <typename PixT, int NT, int MINB>
static int launch_extract2(peeb_ws* ws, const Geom2& g, const PeeBatch& bt, long long nbands, unsigned* stage_bits,
                           int* stage_cnt, long long zero_words, cudaStream_t st) {
    const size_t smem = layout2(g, 2).total;
    int rc = set_smem2(pee2_extract_kernel<PixT, NT, MINB>, smem); if (rc) return rc;
    ProfScope p(ws, PEEB_K_PEE_EXTRACT, st);
    pee2_extract_kernel<PixT, NT, MINB><<<(unsigned)nbands, NT, smem, st>>>(g, bt, stage_bits, stage_cnt, zero_words);
    return PEEB_OK;
}
// (CTA size, CTAs per SM) pairs the kernels are compiled for: registers per thread = 64K / (NT * MINB)
#define PEEB_DISPATCH2_T(FN, PIXT, ...)                                                     \
    (g.threads == 128 ? (g.minb >= 6 ? FN<PIXT, 128, 6>(__VA_ARGS__) : g.minb == 5 ? FN<PIXT, 128, 5>(__VA_ARGS__) \
                                                                    : FN<PIXT, 128, 4>(__VA_ARGS__))               \
     : g.threads == 256 ? (g.minb >= 3 ? FN<PIXT, 256, 3>(__VA_ARGS__) : g.minb == 2 ? FN<PIXT, 256, 2>(__VA_ARGS__) \
                                                                    : FN<PIXT, 256, 1>(__VA_ARGS__))               \
     : g.threads == 512 ? (g.minb >= 2 ? FN<PIXT, 512, 2>(__VA_ARGS__) : FN<PIXT, 512, 1>(__VA_ARGS__))           \
                        : FN<PIXT, 1024, 1>(__VA_ARGS__))
#ifdef PEEB_DEV_ONE_VARIANT  // development builds (SASS inspection): only the 16-bit, 256-thread, 3-CTA kernels
#undef PEEB_DISPATCH2_T
#define PEEB_DISPATCH2_T(FN, PIXT, ...) FN<PIXT, 256, 3>(__VA_ARGS__)
#define PEEB_DISPATCH2(FN, ...) PEEB_DISPATCH2_T(FN, unsigned short, __VA_ARGS__)
#else
#define PEEB_DISPATCH2(FN, ...) \
    (g.itemsize == 2 ? PEEB_DISPATCH2_T(FN, unsigned short, __VA_ARGS__) : PEEB_DISPATCH2_T(FN, unsigned char, __VA_ARGS__))
#endif

template <typename PixT, int NT, int MINB>
static int launch_hist2(peeb_ws* ws, const Geom2& g, long long nbands, const unsigned char* src, long long src_stride,
                        int tmax, unsigned* hist, cudaStream_t st) {
    const size_t smem = layout2(g, 3).total;
    int rc = set_smem2(pee2_hist_kernel<PixT, NT, MINB>, smem); if (rc) return rc;
    ProfScope p(ws, PEEB_K_PEE_HIST, st);
    pee2_hist_kernel<PixT, NT, MINB><<<(unsigned)nbands, NT, smem, st>>>(g, src, src_stride, tmax, hist);
    return PEEB_OK;
}

int hist_batch_impl2(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                     int bit_depth, uint32_t* hist, cudaStream_t st) {
    Geom2 g;
    int rc = make_geom2(ws, h, w, itemsize, bit_depth, 3, false, g);
    if (rc) return rc;
    if (g.bulk && ((((uintptr_t)src) | (uint64_t)src_stride) & 15)) g.bulk = 0;
    const long long nbands = (long long)n_units * g.nb;
    PEEB_REQUIRE(nbands < (1ll << 30), "peeb_pee_hist_batch: too many bands");
    rc = PEEB_DISPATCH2(launch_hist2, ws, g, nbands, (const unsigned char*)src, (long long)src_stride, 1 << (bit_depth - 1), hist, st);
    if (rc) return rc;
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

// ---- small-image path: geometry and launch of the cluster kernels
// A batch takes it when it is one or two small images (the band kernels would run a dozen CTAs; measured: one
// 512x512 slice 51 against 58 us per round trip, but eight slices 107 against 60 us -- clusters of 16 CTAs are
// placed one per GPC at a time) and an image fits one cluster: C CTAs of R = ceil(h / C) rows each, R + 2 <= 64
// (one sweep covers a band and its halo rows).  PEEB_CLUSTER=0 / 1 forces the choice (tests, A/B runs).
static bool make_geom_cluster(peeb_ws* ws, int n_units, int h, int w, int itemsize, int bit_depth, int kind /*4, 5*/,
                              int regular_bands, int max_c, Geom2& g, int& C) {
    const char* force = getenv("PEEB_CLUSTER");
    if (force && atoi(force) == 0) return false;
    if (!(force && atoi(force) == 1) && (long long)n_units * regular_bands * 8 > ws->sm_count) return false;
    if (h < 3 || w < 3) return false;
    g = Geom2{};
    g.h = h; g.w = w; g.itemsize = itemsize;
    g.maxval = (1 << bit_depth) - 1;
    g.rowbytes = w * itemsize;
    g.bulk = ws->use_bulk && (g.rowbytes % 16 == 0);
    g.lmw = (w + 7) / 8;
    g.lmpitch = (int)align_up((size_t)g.lmw, 4) + 12;
    g.lm_direct = 1;
    // (512 threads per CTA: more, smaller warp items -- 28 against 31 us for the embed of one 512x512 slice)
    g.threads = getenv("PEEB_CLUSTER_THREADS") ? atoi(getenv("PEEB_CLUSTER_THREADS")) : 512; g.minb = 1;
    if (g.threads != 256 && g.threads != 512) g.threads = 512;
    const int pxs = 16 / itemsize, nsteps = (g.rowbytes + 15) / 16;
    for (int c : {CLUSTER_MAX, 8}) {
        const int R = (h + c - 1) / c;
        if (c > max_c || R + 2 > 64) continue;
        g.R = std::max(R, 1);
        g.nb = (h + g.R - 1) / g.R;
        g.rpw_log2 = 0;
        while ((1 << g.rpw_log2) < (g.R + 3) / 2) ++g.rpw_log2;
        g.rpw = 1 << g.rpw_log2;
        const int parts = 32 / g.rpw;
        // about one warp item per warp and sweep
        int cws = nsteps / (parts * (g.threads / 32));
        cws = std::max(1, std::min(cws, 64 / pxs));
        g.cws = cws; g.CW = cws * pxs;
        g.ncol = (nsteps + cws - 1) / cws;
        g.pitch = (int)align_up((size_t)std::max(g.rowbytes, g.ncol * g.CW * itemsize), 128) + 16;
        g.tpitch = (int)align_up((size_t)g.ncol, 16);
        g.nic = (g.ncol + parts - 1) / parts;
        g.bandwords = (g.R * ((w + 1) / 2) + 31) / 32 + 2;
        if (layout2(g, kind).total > (size_t)ws->max_smem_optin) continue;
        C = c;
        if (getenv("PEEB_DEBUG_GEOM")) {
            static int printed[2] = {0, 0};
            if (printed[kind - 4]++ < 1)
                fprintf(stderr, "[peeb] %dx%dx%d kind %d (cluster): C=%d R=%d nb=%d rpw=%d CW=%d ncol=%d smem=%zu\n", h, w, itemsize, kind, C,
                        g.R, g.nb, g.rpw, g.CW, g.ncol, layout2(g, kind).total);
        }
        return true;
    }
    return false;
}

template <typename K, typename... Args>
static int launch_cluster(K kernel, int n_units, int C, int threads, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(n_units * C)); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)C; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    // attributes and the "can a cluster of this shape be placed at all" query once per (kernel, shape): they cost tens of
    // microseconds of host time, which a single-slice call would wait for
    struct Seen { const void* k; int C, threads; size_t smem; int ok; };
    static thread_local Seen seen[8];
    static thread_local int nseen = 0;
    int known = -1;
    for (int i = 0; i < nseen; ++i)
        if (seen[i].k == (const void*)kernel && seen[i].C == C && seen[i].threads == threads && seen[i].smem >= smem) known = seen[i].ok;
    if (known < 0) {
        int ok = 1;
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) ok = 0;
        if (ok && C > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ok = 0;
        int nclusters = 0;
        if (ok && (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1)) ok = 0;
        if (!ok) cudaGetLastError();
        if (nseen < 8) seen[nseen++] = Seen{(const void*)kernel, C, threads, smem, ok};
        known = ok;
    }
    if (!known) return 1;
    if (cudaLaunchKernelEx(&cfg, kernel, args...) != cudaSuccess) { cudaGetLastError(); return 1; }
    return 0;
}

// 0: launched; 1: this batch does not take the cluster path (the caller goes on with the band kernels)
static int try_cluster_embed(peeb_ws* ws, const Geom2& gr, PeeBatch bt, int h, int w, int itemsize, int bit_depth, cudaStream_t st) {
    if (bt.active) return 1;
    if (bt.lm && ((((uintptr_t)bt.lm | (uint64_t)bt.lm_stride) & 3) || (((w + 7) / 8) & 3))) return 1;
    for (int max_c : {CLUSTER_MAX, 8}) {  // 16 CTAs per cluster is more than the portable size: fall back to 8
        Geom2 g; int C = 0;
        if (!make_geom_cluster(ws, bt.n_units, h, w, itemsize, bit_depth, 4, gr.nb, max_c, g, C)) return 1;
        g.bulk = g.bulk && gr.bulk;
        const size_t smem = layout2(g, 4).total;
        ProfScope p(ws, PEEB_K_PEE_EMBED, st);
        const int rc = itemsize == 2 ? (g.threads == 512 ? launch_cluster(pee2_cluster_embed_kernel<unsigned short, 512>, bt.n_units, C, 512, smem, st, g, bt, C)
                                                         : launch_cluster(pee2_cluster_embed_kernel<unsigned short, 256>, bt.n_units, C, 256, smem, st, g, bt, C))
                                     : (g.threads == 512 ? launch_cluster(pee2_cluster_embed_kernel<unsigned char, 512>, bt.n_units, C, 512, smem, st, g, bt, C)
                                                         : launch_cluster(pee2_cluster_embed_kernel<unsigned char, 256>, bt.n_units, C, 256, smem, st, g, bt, C));
        if (rc == 0) return 0;
        if (C <= 8) return 1;
    }
    return 1;
}
static int try_cluster_extract(peeb_ws* ws, const Geom2& gr, PeeBatch bt, int h, int w, int itemsize, int bit_depth,
                               long long zero_words, cudaStream_t st) {
    for (int max_c : {CLUSTER_MAX, 8}) {
        Geom2 g; int C = 0;
        if (!make_geom_cluster(ws, bt.n_units, h, w, itemsize, bit_depth, 5, gr.nb, max_c, g, C)) return 1;
        g.bulk = g.bulk && gr.bulk;
        const size_t smem = layout2(g, 5).total;
        ProfScope p(ws, PEEB_K_PEE_EXTRACT, st);
        const int rc = itemsize == 2 ? (g.threads == 512 ? launch_cluster(pee2_cluster_extract_kernel<unsigned short, 512>, bt.n_units, C, 512, smem, st, g, bt, C, zero_words)
                                                         : launch_cluster(pee2_cluster_extract_kernel<unsigned short, 256>, bt.n_units, C, 256, smem, st, g, bt, C, zero_words))
                                     : (g.threads == 512 ? launch_cluster(pee2_cluster_extract_kernel<unsigned char, 512>, bt.n_units, C, 512, smem, st, g, bt, C, zero_words)
                                                         : launch_cluster(pee2_cluster_extract_kernel<unsigned char, 256>, bt.n_units, C, 256, smem, st, g, bt, C, zero_words));
        if (rc == 0) return 0;
        if (C <= 8) return 1;
    }
    return 1;
}

// T == nullptr: the thresholds are chosen on the device (SURVEY Appendix A "threshold selection"): error histogram of
// every unit -> smallest T whose estimate holds the payload -> embed -> units whose real capacity falls short get
// T + 1 and are embedded again (only they), until every unit fits or has reached tmax.  The host only reads one
// counter per round.  The chosen T comes back in info[u][0].
int embed_batch_impl2(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                      int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload,
                      int64_t payload_stride, void* marked, int64_t marked_stride, uint8_t* lm, int64_t lm_stride,
                      int64_t* info, cudaStream_t st, int slot) {
    Geom2 g;
    int rc = make_geom2(ws, h, w, itemsize, bit_depth, 1, (((uintptr_t)lm | (uint64_t)lm_stride) & 3) == 0, g);
    if (rc) return rc;
    if (g.bulk && ((((uintptr_t)src) | (uintptr_t)marked | (uint64_t)src_stride | (uint64_t)marked_stride) & 15))
        g.bulk = 0;  // unaligned user buffers: plain copies
    const long long nbands = (long long)n_units * g.nb;
    PEEB_REQUIRE(nbands < (1ll << 30), "peeb_pee_embed_batch: too many bands");
    const bool auto_T = T == nullptr;
    PEEB_REQUIRE(!auto_T || src_stride != 0, "peeb_pee_embed_batch: threshold selection needs one cover per unit");
    int* dT; unsigned* dN; char* extra;
    const size_t cnt_bytes = align_up((size_t)nbands * sizeof(int), 256);
    const size_t st_bytes = align_up((size_t)nbands * sizeof(unsigned long long), 256);
    const size_t rc_bytes = align_up((size_t)n_units * h * g.tpitch, 256);
    const size_t act_bytes = auto_T ? align_up((size_t)(n_units + 1) * sizeof(int), 256) : 0;
    std::vector<int32_t> ones;
    if (auto_T) ones.assign((size_t)n_units, 1);
    rc = upload_unit_tables(ws, slot, n_units, auto_T ? ones.data() : T, n_bits, bit_depth,
                            cnt_bytes + st_bytes + 256 + rc_bytes + act_bytes, st, &dT, &dN, &extra);
    if (rc) return rc;
    int* band_cnt = (int*)extra;
    unsigned long long* status = (unsigned long long*)(extra + cnt_bytes);
    unsigned* ticket = (unsigned*)(extra + cnt_bytes + st_bytes);
    unsigned char* rowcnt = (unsigned char*)(extra + cnt_bytes + st_bytes + 256);
    int* active = auto_T ? (int*)(extra + cnt_bytes + st_bytes + 256 + rc_bytes) : nullptr;
    PeeBatch bt{};
    bt.src = (const unsigned char*)src; bt.src_stride = src_stride;
    bt.dst = (unsigned char*)marked; bt.dst_stride = marked_stride;
    bt.lm = lm; bt.lm_stride = lm_stride;
    bt.payload = payload; bt.payload_stride = payload_stride;
    bt.payload_out = nullptr; bt.T = dT; bt.n_bits = dN; bt.info = (long long*)info; bt.n_units = n_units;
    bt.steps = ws->step_counters_on ? (unsigned long long*)ws->step_counters.ptr : nullptr;
    bt.active = active;
    if (!auto_T) {
        if (try_cluster_embed(ws, g, bt, h, w, itemsize, bit_depth, st) == 0) return PEEB_OK;
        rc = PEEB_DISPATCH2(launch_embed2, ws, g, bt, nbands, band_cnt, rowcnt, ticket, status, st);
        if (rc) return rc;
        PEEB_CUDA(cudaGetLastError());
        return PEEB_OK;
    }
    const int tmax = 1 << (bit_depth - 1);
    rc = scratch_reserve(ws->hist, (size_t)n_units * 4 * tmax * sizeof(unsigned));
    if (rc) return rc;
    unsigned* hist = (unsigned*)ws->hist.ptr;
    PEEB_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_units * 4 * tmax * sizeof(unsigned), st));
    rc = hist_batch_impl2(ws, src, src_stride, n_units, h, w, itemsize, bit_depth, hist, st);
    if (rc) return rc;
    pee2_pick_T_kernel<<<n_units, 256, 0, st>>>(hist, tmax, dN, dT, active);  // every unit takes part in round one
    int* remaining_h = (int*)((char*)ws->ptable_h_cur[slot] + align_up((size_t)n_units * 8, 256));  // pinned, behind the tables
    int* remaining = active + n_units;
    for (int round = 0; round <= tmax; ++round) {
        rc = PEEB_DISPATCH2(launch_embed2, ws, g, bt, nbands, band_cnt, rowcnt, ticket, status, st);
        if (rc) return rc;
        PEEB_CUDA(cudaMemsetAsync(remaining, 0, sizeof(int), st));
        pee2_retry_kernel<<<(n_units + 255) / 256, 256, 0, st>>>(n_units, tmax, (long long*)info, dT, active, remaining);
        PEEB_CUDA(cudaMemcpyAsync(remaining_h, remaining, sizeof(int), cudaMemcpyDeviceToHost, st));
        PEEB_CUDA(cudaStreamSynchronize(st));
        if (*remaining_h == 0) break;
    }
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int extract_batch_impl2(peeb_ws* ws, const void* marked, int64_t marked_stride, int n_units, int h, int w,
                        int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* lm,
                        int64_t lm_stride, uint8_t* payload_out, int64_t payload_stride, void* recovered,
                        int64_t recovered_stride, int64_t* info, cudaStream_t st, int slot) {
    Geom2 g;
    int rc = make_geom2(ws, h, w, itemsize, bit_depth, 2, false, g);
    if (rc) return rc;
    if (g.bulk && ((((uintptr_t)marked) | (uintptr_t)recovered | (uint64_t)marked_stride | (uint64_t)recovered_stride) & 15))
        g.bulk = 0;
    const long long nbands = (long long)n_units * g.nb;
    PEEB_REQUIRE(nbands < (1ll << 30), "peeb_pee_extract_batch: too many bands");
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
    // bytes of every unit's output that start from zero: the whole row (the documented zero padding); a single unit
    // may come with a buffer that is only as long as its own payload (stride unused)
    const size_t pb0 = peeb_payload_bytes(n_bits[0]);
    const long long zero_words = (long long)((n_units == 1 && (size_t)payload_stride < pb0) ? pb0 : (size_t)payload_stride) / 4;
    if (try_cluster_extract(ws, g, bt, h, w, itemsize, bit_depth, zero_words, st) == 0) return PEEB_OK;
    rc = PEEB_DISPATCH2(launch_extract2, ws, g, bt, nbands, stage_bits, stage_cnt, zero_words, st);
    if (rc) return rc;
    PEEB_CUDA(cudaGetLastError());
    {
        ProfScope p(ws, PEEB_K_PEE_GATHER, st);
        // pieces per block: about four blocks per SM over the batch, pieces spread evenly over a unit's blocks
        const int pieces = 2 * g.nb;
        long long want = ((long long)pieces * n_units) / ((long long)ws->sm_count * 4);
        want = std::max<long long>(1, std::min<long long>(want, GATHER_PPB));
        const int nblk = (int)((pieces + want - 1) / want), ppb = (pieces + nblk - 1) / nblk;
        dim3 grid((unsigned)nblk, (unsigned)n_units);
        pee2_gather_kernel<<<grid, 256, 0, st>>>(g.nb, g.bandwords, ppb, bt, stage_bits, stage_cnt);
    }
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

}  // namespace peeb

#ifdef PEEB_DEBUG_BOUNDS
namespace peeb {
// reset = 2: one check that holds and one that does not (4 bytes at offset 16 of a 16-byte region), so that a test
// can see the checker report before it trusts a clean run
__global__ void bounds_selftest_kernel() {
    BOUNDS(99, 0, 16, 16);
    BOUNDS(99, 16, 4, 16);
}
}  // namespace peeb
#endif
// peeb200.h: what the bounds-checked build found (all zeros, out6[5] = 0, in a normal build)
extern "C" __attribute__((visibility("default"))) int peeb_debug_bounds(unsigned long long* out6, int reset) {
    if (!out6) return PEEB_E_INVALID;
    for (int i = 0; i < 6; ++i) out6[i] = 0ull;
#ifdef PEEB_DEBUG_BOUNDS
    if (reset == 2) peeb::bounds_selftest_kernel<<<1, 1>>>();
    if (cudaDeviceSynchronize() != cudaSuccess) return PEEB_E_CUDA;
    if (cudaMemcpyFromSymbol(out6, peeb::g_bounds, sizeof(unsigned long long) * 6) != cudaSuccess) return PEEB_E_CUDA;
    out6[5] = 1ull;
    unsigned long long per_site[128];
    if (cudaMemcpyFromSymbol(per_site, peeb::g_bounds_site, sizeof(per_site)) != cudaSuccess) return PEEB_E_CUDA;
    for (int i = 0; i < 128; ++i)
        if (per_site[i] && reset != 2) fprintf(stderr, "peeb bounds: site %d: %llu violations\n", i, per_site[i]);
    if (reset) {
        const unsigned long long z[128] = {0};
        if (cudaMemcpyToSymbol(peeb::g_bounds, z, sizeof(unsigned long long) * 6) != cudaSuccess) return PEEB_E_CUDA;
        if (cudaMemcpyToSymbol(peeb::g_bounds_site, z, sizeof(z)) != cudaSuccess) return PEEB_E_CUDA;
    }
#else
    (void)reset;
#endif
    return PEEB_OK;
}

#ifdef PEEB_PHASE_TIMING
extern "C" __attribute__((visibility("default"))) int peeb_debug_phases(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, peeb::g_phase, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(peeb::g_phase, z, sizeof(z)); }
    return 0;
}
#endif
