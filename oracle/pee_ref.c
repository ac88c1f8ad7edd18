/* TEST INFRASTRUCTURE ONLY -- scalar C restatement of the PEE specification in
 * SURVEY.md Appendix A (row a10).  *** PARITY UNPINNED ***: the reference
 * repository contains no PEE code, so this follows the written
 * specification, not a reference source file.  It is an independent second
 * formulation of oracle/pee_numpy.py (sequential raster walk with a running
 * carrier counter instead of masks + cumsum); tests cross-check the two.
 *
 * Never linked into, loaded by or called from the product library.  Used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs only.
 *
 * Build: oracle/build_oracle.py  ->  oracle/libpee_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline int64_t px_get(const void* img, int itemsize, int64_t idx) {
    return itemsize == 1 ? (int64_t)((const uint8_t*)img)[idx] : (int64_t)((const uint16_t*)img)[idx];
}
static inline void px_set(void* img, int itemsize, int64_t idx, int64_t v) {
    if (itemsize == 1) ((uint8_t*)img)[idx] = (uint8_t)v; else ((uint16_t*)img)[idx] = (uint16_t)v;
}
static inline int payload_bit(const uint8_t* payload, int64_t n_bits, int64_t k) {
    if (k >= n_bits) return 0;                       /* zero padding past the payload */
    return (payload[k >> 3] >> (7 - (k & 7))) & 1;   /* most significant bit first    */
}

/* info[8] = {T, n_bits, capacity, cap0, cap1, n_flagged, sse, status}
 * status 0 ok, -2 payload larger than capacity (outputs are still the
 * zero-padded embed of the first `capacity` bits, as pee_sweep wants).
 * `marked` must not alias `img`. lm_packed: h * ceil(w/8) bytes. */
int pee_ref_embed(const void* img, int h, int w, int itemsize, int64_t maxval, int T,
                  const uint8_t* payload, int64_t n_bits,
                  void* marked, uint8_t* lm_packed, int64_t* info)
{
    const int lmw = (w + 7) / 8;
    memcpy(marked, img, (size_t)h * w * itemsize);
    memset(lm_packed, 0, (size_t)h * lmw);
    int64_t cap[2] = {0, 0}, flagged = 0, sse = 0;
    if (h >= 3 && w >= 3) {
        int64_t next_bit = 0;
        for (int colour = 0; colour < 2; ++colour) {
            for (int i = 1; i <= h - 2; ++i) {
                for (int j = 1; j <= w - 2; ++j) {
                    if (((i + j) & 1) != colour) continue;
                    const int64_t at = (int64_t)i * w + j;
                    const int64_t x = px_get(marked, itemsize, at);
                    const int64_t p = (px_get(marked, itemsize, at - w) + px_get(marked, itemsize, at + w) +
                                       px_get(marked, itemsize, at - 1) + px_get(marked, itemsize, at + 1)) >> 2;
                    const int64_t e = x - p;
                    int64_t y = x;
                    int flag = 0;
                    if (e >= -T && e < T) {
                        const int64_t v = p + 2 * e;
                        if (v < 0 || v + 1 > maxval) flag = 1;
                        else { y = v + payload_bit(payload, n_bits, next_bit); ++next_bit; ++cap[colour]; }
                    } else if (e >= T) {
                        if (x + T > maxval) flag = 1; else y = x + T;
                    } else {
                        if (x - T < 0) flag = 1; else y = x - T;
                    }
                    if (flag) { lm_packed[(int64_t)i * lmw + (j >> 3)] |= (uint8_t)(0x80u >> (j & 7)); ++flagged; }
                    px_set(marked, itemsize, at, y);
                }
            }
        }
        for (int64_t n = 0; n < (int64_t)h * w; ++n) {
            const int64_t d = px_get(marked, itemsize, n) - px_get(img, itemsize, n);
            sse += d * d;
        }
    }
    info[0] = T; info[1] = n_bits; info[2] = cap[0] + cap[1]; info[3] = cap[0]; info[4] = cap[1];
    info[5] = flagged; info[6] = sse; info[7] = (n_bits > cap[0] + cap[1]) ? -2 : 0;
    return (int)info[7];
}

/* payload_out: ceil(n_bits/8) bytes (zeroed here); recovered must not alias marked.
 * returns 0, or -2 if fewer than n_bits carriers were found. */
int pee_ref_extract(const void* marked, int h, int w, int itemsize, int T,
                    const uint8_t* lm_packed, int64_t n_bits,
                    uint8_t* payload_out, void* recovered)
{
    const int lmw = (w + 7) / 8;
    memcpy(recovered, marked, (size_t)h * w * itemsize);
    memset(payload_out, 0, (size_t)((n_bits + 7) / 8));
    if (h < 3 || w < 3) return n_bits > 0 ? -2 : 0;
    /* carrier bits per colour, raster order */
    uint8_t* bits[2];
    int64_t nb[2] = {0, 0};
    bits[0] = (uint8_t*)malloc((size_t)h * w / 2 + 8);
    bits[1] = (uint8_t*)malloc((size_t)h * w / 2 + 8);
    for (int colour = 1; colour >= 0; --colour) {
        for (int i = 1; i <= h - 2; ++i) {
            for (int j = 1; j <= w - 2; ++j) {
                if (((i + j) & 1) != colour) continue;
                if (lm_packed[(int64_t)i * lmw + (j >> 3)] & (0x80u >> (j & 7))) continue;
                const int64_t at = (int64_t)i * w + j;
                const int64_t x = px_get(recovered, itemsize, at);
                const int64_t p = (px_get(recovered, itemsize, at - w) + px_get(recovered, itemsize, at + w) +
                                   px_get(recovered, itemsize, at - 1) + px_get(recovered, itemsize, at + 1)) >> 2;
                const int64_t ee = x - p;
                int64_t e;
                if (ee >= -2 * (int64_t)T && ee < 2 * (int64_t)T) {
                    bits[colour][nb[colour]++] = (uint8_t)(ee & 1);
                    e = ee >> 1;                      /* arithmetic shift == floor */
                } else if (ee >= 2 * (int64_t)T) e = ee - T;
                else e = ee + T;
                px_set(recovered, itemsize, at, p + e);
            }
        }
    }
    int rc = 0;
    if (n_bits > nb[0] + nb[1]) rc = -2;
    else {
        for (int64_t k = 0; k < n_bits; ++k) {
            const int b = k < nb[0] ? bits[0][k] : bits[1][k - nb[0]];
            if (b) payload_out[k >> 3] |= (uint8_t)(0x80u >> (k & 7));
        }
    }
    free(bits[0]); free(bits[1]);
    return rc;
}

/* hist: 2 * (2*tmax) int64, hist[c][e + tmax]; errors outside [-tmax, tmax) dropped. */
void pee_ref_hist(const void* img, int h, int w, int itemsize, int64_t maxval, int64_t tmax, int64_t* hist)
{
    memset(hist, 0, sizeof(int64_t) * 4 * (size_t)tmax);
    for (int i = 1; i <= h - 2; ++i)
        for (int j = 1; j <= w - 2; ++j) {
            const int64_t at = (int64_t)i * w + j;
            const int64_t x = px_get(img, itemsize, at);
            const int64_t p = (px_get(img, itemsize, at - w) + px_get(img, itemsize, at + w) +
                               px_get(img, itemsize, at - 1) + px_get(img, itemsize, at + 1)) >> 2;
            const int64_t e = x - p, v = p + 2 * e;
            if (v < 0 || v + 1 > maxval || e < -tmax || e >= tmax) continue;
            hist[(size_t)((i + j) & 1) * 2 * tmax + (size_t)(e + tmax)] += 1;
        }
}

/* Batches: images are independent, so the loop over them is the only
 * parallelism a CPU port has.  Contiguous images, per-image payload stride. */
int pee_ref_embed_batch(const void* imgs, int n, int h, int w, int itemsize, int64_t maxval, int T,
                        const uint8_t* payloads, int64_t payload_stride, const int64_t* n_bits,
                        void* marked, uint8_t* lm_packed, int64_t* info, int threads)
{
    const size_t isz = (size_t)h * w * itemsize, lsz = (size_t)h * ((w + 7) / 8);
    int worst = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic) reduction(min:worst)
#endif
    for (int u = 0; u < n; ++u) {
        int rc = pee_ref_embed((const char*)imgs + u * isz, h, w, itemsize, maxval, T,
                               payloads + (size_t)u * payload_stride, n_bits[u],
                               (char*)marked + u * isz, lm_packed + u * lsz, info + 8 * (size_t)u);
        if (rc < worst) worst = rc;
    }
    return worst;
}

int pee_ref_extract_batch(const void* marked, int n, int h, int w, int itemsize, int T,
                          const uint8_t* lm_packed, const int64_t* n_bits,
                          uint8_t* payloads_out, int64_t payload_stride, void* recovered, int threads)
{
    const size_t isz = (size_t)h * w * itemsize, lsz = (size_t)h * ((w + 7) / 8);
    int worst = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic) reduction(min:worst)
#endif
    for (int u = 0; u < n; ++u) {
        int rc = pee_ref_extract((const char*)marked + u * isz, h, w, itemsize, T,
                                 lm_packed + u * lsz, n_bits[u],
                                 payloads_out + (size_t)u * payload_stride, (char*)recovered + u * isz);
        if (rc < worst) worst = rc;
    }
    return worst;
}

int pee_ref_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}


/* ------------------------------------------------------------------------------------------------
 * N1: the causal predictor family.  DESIGN.md "Appendix A2" -- *** PARITY UNPINNED ***.
 * MED predictor (JPEG-LS): a = W, b = N, c = NW, p = clamp(a + b - c, min(a,b), max(a,b)).
 * Domain: 1 <= i < h, 1 <= j < w (row 0 and column 0 never change).  ONE pass in raster order;
 * the embedder predicts from ORIGINAL pixels (so it is fully parallel), the extractor from the
 * pixels it has already recovered (W, N, NW come earlier in raster order: anti-diagonal wavefront).
 * Classes, flags, carrier rule and zero padding are those of Appendix A; info cap0 = capacity, cap1 = 0.
 * ------------------------------------------------------------------------------------------------ */
static inline int64_t med_predict(int64_t a, int64_t b, int64_t c) {
    const int64_t lo = a < b ? a : b, hi = a < b ? b : a;
    int64_t p = a + b - c;
    if (p < lo) p = lo;
    if (p > hi) p = hi;
    return p;
}

int pee_med_ref_embed(const void* img, int h, int w, int itemsize, int64_t maxval, int T,
                      const uint8_t* payload, int64_t n_bits, void* marked, uint8_t* lm_packed, int64_t* info)
{
    const int lmw = (w + 7) / 8;
    memcpy(marked, img, (size_t)h * w * itemsize);
    memset(lm_packed, 0, (size_t)h * lmw);
    int64_t cap = 0, flagged = 0, sse = 0, next_bit = 0;
    for (int i = 1; i < h; ++i)
        for (int j = 1; j < w; ++j) {
            const int64_t at = (int64_t)i * w + j;
            const int64_t x = px_get(img, itemsize, at);
            const int64_t p = med_predict(px_get(img, itemsize, at - 1), px_get(img, itemsize, at - w),
                                          px_get(img, itemsize, at - w - 1));
            const int64_t e = x - p;
            int64_t y = x;
            int flag = 0;
            if (e >= -T && e < T) {
                const int64_t v = p + 2 * e;
                if (v < 0 || v + 1 > maxval) flag = 1;
                else { y = v + payload_bit(payload, n_bits, next_bit); ++next_bit; ++cap; }
            } else if (e >= T) {
                if (x + T > maxval) flag = 1; else y = x + T;
            } else {
                if (x - T < 0) flag = 1; else y = x - T;
            }
            if (flag) { lm_packed[(int64_t)i * lmw + (j >> 3)] |= (uint8_t)(0x80u >> (j & 7)); ++flagged; }
            px_set(marked, itemsize, at, y);
            sse += (y - x) * (y - x);
        }
    info[0] = T; info[1] = n_bits; info[2] = cap; info[3] = cap; info[4] = 0;
    info[5] = flagged; info[6] = sse; info[7] = (n_bits > cap) ? -2 : 0;
    return (int)info[7];
}

int pee_med_ref_extract(const void* marked, int h, int w, int itemsize, int T, const uint8_t* lm_packed,
                        int64_t n_bits, uint8_t* payload_out, void* recovered)
{
    const int lmw = (w + 7) / 8;
    memcpy(recovered, marked, (size_t)h * w * itemsize);
    memset(payload_out, 0, (size_t)((n_bits + 7) / 8));
    int64_t k = 0;
    for (int i = 1; i < h; ++i)
        for (int j = 1; j < w; ++j) {
            if (lm_packed[(int64_t)i * lmw + (j >> 3)] & (0x80u >> (j & 7))) continue;
            const int64_t at = (int64_t)i * w + j;
            const int64_t x = px_get(recovered, itemsize, at);
            const int64_t p = med_predict(px_get(recovered, itemsize, at - 1), px_get(recovered, itemsize, at - w),
                                          px_get(recovered, itemsize, at - w - 1));
            const int64_t ee = x - p;
            int64_t e;
            if (ee >= -2 * (int64_t)T && ee < 2 * (int64_t)T) {
                if (k < n_bits && (ee & 1)) payload_out[k >> 3] |= (uint8_t)(0x80u >> (k & 7));
                ++k;
                e = ee >> 1;
            } else if (ee >= 2 * (int64_t)T) e = ee - T;
            else e = ee + T;
            px_set(recovered, itemsize, at, p + e);
        }
    return n_bits > k ? -2 : 0;
}
