/* libpeeb200 -- B200 (sm_100a) kernels for the pixel-array hot path of
 * wesleyfn/codec-tcc, behind a flat C ABI.
 *
 * The reference has no FFI of its own: its boundary is a set of Python
 * callables on numpy arrays.  Each entry point below
 * names the reference callable (file:line under the reference root) whose
 * arithmetic it carries out; the Python package `codec_tcc_b200` binds them
 * with ctypes and keeps the reference's names, argument meaning and error
 * behaviour (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - every function returns 0 (PEEB_OK) or a negative PEEB_E_* code;
 *     peeb_last_error() gives the message (thread local).  No exceptions cross
 *     the ABI.
 *   - `*_h` entry points take HOST pointers, do their own host<->device copies
 *     on the workspace's stream and return after synchronising it.
 *     All other data pointers are DEVICE pointers on the workspace's device and
 *     the call only enqueues work on `stream` (a cudaStream_t, may be 0).
 *   - images are C-contiguous, row-major, itemsize 1 (uint8) or 2 (uint16).
 *   - bit strings (payloads, location maps) are packed most-significant-bit
 *     first within each byte (numpy.packbits default; same bit order as
 *     message_to_bits, src/codec.py:239-240).
 *   - the library owns no user buffer.  A workspace is thread compatible, not
 *     thread safe: one per (device, calling thread).
 */
#ifndef PEEB200_H
#define PEEB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden: only this header is exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define PEEB_OK 0
#define PEEB_E_CUDA (-1)     /* a CUDA runtime call failed                       */
#define PEEB_E_CAPACITY (-2) /* PEE: payload larger than capacity / carriers     */
#define PEEB_E_INVALID (-3)  /* bad argument                                     */
#define PEEB_E_UNSUPPORTED (-4)

#define PEEB_ABI_VERSION 1

typedef struct peeb_ws peeb_ws; /* opaque workspace */

/* ---- lifecycle -------------------------------------------------------- */
int peeb_abi_version(void);
const char* peeb_last_error(void);
int peeb_device_count(int* n);
int peeb_ws_create(int device, peeb_ws** ws);
int peeb_ws_destroy(peeb_ws* ws);
int peeb_ws_sync(peeb_ws* ws);                 /* synchronise the workspace streams  */
void* peeb_ws_stream(peeb_ws* ws);             /* the workspace's own cudaStream_t   */
/* tuning / debugging switch: PEEB_OPT_BULK (TMA bulk copies for band staging, default 1) */
#define PEEB_OPT_BULK 0
int peeb_ws_set_option(peeb_ws* ws, int option, int value);
/* pinned host memory for callers that want full-speed PCIe copies */
int peeb_host_alloc(size_t bytes, void** ptr);
int peeb_host_free(void* ptr);
/* plain device memory (callers without torch) */
int peeb_dev_alloc(peeb_ws* ws, size_t bytes, void** ptr);
int peeb_dev_free(peeb_ws* ws, void* ptr);
int peeb_memcpy_h2d(peeb_ws* ws, void* dst_dev, const void* src_host, size_t bytes, void* stream);
int peeb_memcpy_d2h(peeb_ws* ws, void* dst_host, const void* src_dev, size_t bytes, void* stream);

/* per-kernel device time accounting (CUDA events around each launch; adds a
 * sync per launch -- for bench.py's roofline leg, never for a timed `value`) */
#define PEEB_PROF_SLOTS 20
enum {
    PEEB_K_MOMENTS = 0, PEEB_K_HIST_PLANES = 1, PEEB_K_TILE_MOMENTS = 2, PEEB_K_LSB_EMBED = 3,
    PEEB_K_PLANES_PACK = 4, PEEB_K_PLANES_UNPACK = 5, PEEB_K_COMPACT = 6,
    PEEB_K_PEE_COUNT = 7, PEEB_K_PEE_EMBED = 8, PEEB_K_PEE_EXTRACT = 9, PEEB_K_PEE_GATHER = 10,
    PEEB_K_PEE_HIST = 11, PEEB_K_PEE_FINAL = 12, PEEB_K_LSB_RECOVER = 13, PEEB_K_LSB_EXTRACT = 14,
    PEEB_K_BITMAP_ENCODE = 15, PEEB_K_BITMAP_DECODE = 16
};
int peeb_prof_enable(peeb_ws* ws, int on);  /* also resets the counters */
int peeb_prof_get(peeb_ws* ws, int slot, double* total_ms, long long* launches);
const char* peeb_prof_name(int slot);
/* warp-steps of the PEE embed kernel {all, generic code at a border column, redone by the generic code after the
 * fast code saw a value leave [0, maxval)}: on = 1 starts counting from zero, on = 0 stops; out3 (3 x uint64) gets
 * the counts so far (may be NULL).  Synchronises the device. */
int peeb_pee_step_counters(peeb_ws* ws, int on, uint64_t* out3);

/* ---- a1-a4: distortion moments ---------------------------------------- *
 * One pass over two images gives every integer the metrics need
 * (AnalisadorMSE.calcular_mse src/mse.py:74-116, calcular_ssim_simples
 * :135-179, difference statistics :202-209):
 *   out[0]=sum (a-b)^2   out[1]=sum |a-b|   out[2]=max |a-b|   out[3]=#(a!=b)
 *   out[4]=sum a         out[5]=sum b       out[6]=sum a^2     out[7]=sum b^2
 *   out[8]=sum a*b       out[9]=max a       out[10]=max b      out[11]=n
 * `out` : n_images x 12 int64 (device for the plain call, host for _h).
 * stride_a/stride_b are in elements between consecutive images.            */
#define PEEB_MOMENTS 12
int peeb_moments_batch(peeb_ws* ws, const void* a, const void* b, int64_t n_per_image, int itemsize,
                       int n_images, int64_t stride_a, int64_t stride_b, int64_t* out, void* stream);
int peeb_moments_h(peeb_ws* ws, const void* a_host, const void* b_host, int64_t n, int itemsize,
                   int64_t* out_host);
/* The MSE-only subset -- out[0] (SSE), out[9], out[10] (maxima), out[11] (n); the other slots are 0.
 * It is all calcular_mse (src/mse.py:74-116) needs when both maxima agree, and it is HBM bound where
 * the full set is bound by integer issue. */
int peeb_sse_batch(peeb_ws* ws, const void* a, const void* b, int64_t n_per_image, int itemsize,
                   int n_images, int64_t stride_a, int64_t stride_b, int64_t* out, void* stream);
int peeb_sse_h(peeb_ws* ws, const void* a_host, const void* b_host, int64_t n, int itemsize,
               int64_t* out_host);

/* a1 / a3 / a4 on float64 pixel data: the reference converts any input with
 * np.array(img, dtype=np.float64) (src/mse.py:85,91); integer-valued 8/16-bit data takes
 * the exact integer kernels above, everything else (fractional, negative, wider values)
 * these.  out (DEVICE, 12 doubles) = {sum u, sum v, sum (u-mean_a)^2, sum (v-mean_b)^2,
 * sum (u-mean_a)(v-mean_b), sum (u-v)^2, sum |a-b|, max |a-b|, max a, max b, #(a != b), n}
 * with u = scaled ? (a / div_a) * mul_a : a (the operation order of src/mse.py:104-105),
 * v likewise.  The _h form uploads once and runs both passes the metrics need: out_host
 * (26 doubles) = plain pass | normalised + centred pass (scaled only when the maxima
 * differ, src/mse.py:101) | the two centring values used.  Tolerance vs numpy: 1e-9 rel. */
int peeb_moments_f64(peeb_ws* ws, const double* a, const double* b, int64_t n, int scaled, double div_a,
                     double mul_a, double div_b, double mul_b, double mean_a, double mean_b, double* out,
                     void* stream);
int peeb_moments_f64_h(peeb_ws* ws, const double* a_host, const double* b_host, int64_t n, double* out_host);

/* ---- a5: histogram + bit-plane population counts ---------------------- *
 * Everything adaptive_modalities_decomposition / calculate_entropy /
 * calculate_mutual_information (src/codec.py:561-599, 489-502, 504-559) need
 * from the pixels: hist[65536] uint32 (256 used for itemsize 1) and the number
 * of set bits in each of the 16 bit-planes.  The float64 entropy sums stay on
 * the host so that numpy's summation order is reproduced exactly.           */
int peeb_hist_planes(peeb_ws* ws, const void* img, int64_t n, int itemsize, uint32_t* hist,
                     uint64_t* plane_ones, void* stream);
int peeb_hist_planes_h(peeb_ws* ws, const void* img_host, int64_t n, int itemsize,
                       uint32_t* hist_host, uint64_t* plane_ones_host);

/* ---- a8: bit-plane pack / unpack -------------------------------------- *
 * extract_local_planes (src/codec.py:789-793) and the plane split of
 * adaptive_modalities_decomposition (:571): planes[k][i] = (img[i] >> (first+k)) & 1,
 * itemsize preserved.  merge_modalities (:215-237): out = OR_k (trunc(planes[k]) << k),
 * out_itemsize = 2 when n_planes > 8 else 1.                                */
int peeb_planes_unpack(peeb_ws* ws, const void* img, int64_t n, int itemsize, int first_plane,
                       int n_planes, void* planes_out, void* stream);
int peeb_planes_unpack_h(peeb_ws* ws, const void* img_host, int64_t n, int itemsize, int first_plane,
                         int n_planes, void* planes_out_host);
int peeb_planes_pack(peeb_ws* ws, const void* planes, int64_t n, int in_itemsize, int n_planes,
                     void* img_out, void* stream);
/* planes given as n_planes separate host arrays (the reference passes a list) */
int peeb_planes_pack_h(peeb_ws* ws, const void* const* plane_ptrs_host, int64_t n, int in_itemsize,
                       int n_planes, void* img_out_host);

/* ---- a6: tile statistics for the hybrid embedder ---------------------- *
 * lsb_embed_block_then_multiplane scans sbs x sbs tiles (edge tiles clipped) of
 * local plane 0 for the largest float(np.var(tile)) (src/codec.py:441-450).
 * The variance of a tile of n values is (n*sum(v^2) - sum(v)^2) / n^2, an exact
 * rational; the device returns sum(v) and sum(v^2) per tile (row-major over
 * tiles, 2 x int64 each) and the host does the exact comparison (and settles
 * float ties on the few candidate tiles, see codec.py in the package).       */
int peeb_tile_moments(peeb_ws* ws, const void* plane, int h, int w, int itemsize, int sbs,
                      int64_t* sums_out, void* stream);
int peeb_tile_moments_h(peeb_ws* ws, const void* plane_host, int h, int w, int itemsize, int sbs,
                        int64_t* sums_out_host);

/* ---- a6/a7: LSB embed with XOR side bitmaps --------------------------- *
 * lsb_embed_multi_plane (src/codec.py:276-318) and the embedding loop of
 * lsb_embed_block_then_multiplane (:456-485).  For plane p (0 <= p < s) the
 * raster positions (start[p] + k) mod n, 0 <= k < len[p], receive
 * (v & 0xFE) | bit where bit = payload bit (bit_off[p] + k); every other
 * position is copied.  bitmaps[p][i] = uint8(v ^ new) (0 where untouched).
 * planes_in/planes_out: s x n elements of `itemsize`; bitmaps: s x n uint8.  */
int peeb_lsb_embed(peeb_ws* ws, const void* planes_in, int64_t n, int itemsize, int s,
                   const int64_t* start, const int64_t* len, const int64_t* bit_off,
                   const uint8_t* payload, int64_t payload_bits, void* planes_out, uint8_t* bitmaps,
                   void* stream);
int peeb_lsb_embed_h(peeb_ws* ws, const void* const* plane_ptrs_host, int64_t n, int itemsize, int s,
                     const int64_t* start, const int64_t* len, const int64_t* bit_off,
                     const uint8_t* payload_host, int64_t payload_bits, void* planes_out_host,
                     uint8_t* bitmaps_out_host);

/* ---- a9: decode_message's gather -------------------------------------- *
 * src/codec.py:767-771: the least significant bits of `plane` at the first
 * `limit` positions where `bitmap` is non-zero, packed MSB first into
 * bits_out (ceil(limit/8) bytes rounded up to a multiple of 4, 4-byte aligned, zero padded); *count_out = how many were
 * found (<= limit).                                                         */
int peeb_compact_bits(peeb_ws* ws, const void* plane, const uint8_t* bitmap, int64_t n, int itemsize,
                      int64_t limit, uint8_t* bits_out, int64_t* count_out, void* stream);
int peeb_compact_bits_h(peeb_ws* ws, const void* plane_host, const uint8_t* bitmap_host, int64_t n,
                        int itemsize, int64_t limit, uint8_t* bits_out_host, int64_t* count_out_host);

/* ---- N4: the true inverse of the LSB path ------------------ *
 * Not in the reference: decode_bin re-saves the stego image as "recovered"
 * (src/codec.py:816-817,838-842) and decode_message is not an inverse of any
 * embedder.  These two are the inverse of
 * lsb_embed_multi_plane / lsb_embed_block_then_multiplane (src/codec.py:276-318,
 * 412-487) on the merged stego image (bit p of a pixel = local plane p):
 *   recover: cover[i] = stego[i] ^ sum_p ((bitmaps[p][i] & 1) << p)   (:309-311 inverted)
 *   extract: bits_out bit (bit_off[p] + k) = bit p of stego[(start[p] + k) mod n],
 *            0 <= k < len[p]   (:299-306, :465-472 inverted), MSB-first bytes;
 *            bits_out: ceil(total_bits/8) bytes rounded up to a multiple of 4, + 4, 4-byte aligned.
 * bitmaps: s x n uint8 back to back (device) / s separate arrays (host).        */
int peeb_lsb_recover(peeb_ws* ws, const void* stego, const uint8_t* bitmaps, int64_t n, int itemsize, int s,
                     void* cover_out, void* stream);
int peeb_lsb_recover_h(peeb_ws* ws, const void* stego_host, const uint8_t* const* bitmap_ptrs_host, int64_t n,
                       int itemsize, int s, void* cover_out_host);
int peeb_lsb_extract(peeb_ws* ws, const void* stego, int64_t n, int itemsize, int s, const int64_t* start,
                     const int64_t* len, const int64_t* bit_off, int64_t total_bits, uint8_t* bits_out,
                     void* stream);
int peeb_lsb_extract_h(peeb_ws* ws, const void* stego_host, int64_t n, int itemsize, int s,
                       const int64_t* start, const int64_t* len, const int64_t* bit_off, int64_t total_bits,
                       uint8_t* bits_out_host);

/* ---- a10: Prediction-Error Expansion ---------------------------------- *
 * Not in the reference; specified in SURVEY.md Appendix A.
 * Units (images) are independent; a batch is n_units images of h x w with byte
 * strides between them.  src_stride may be 0 (every unit embeds into the same
 * cover -- the threshold sweep).  T[u], n_bits[u] are HOST arrays.  T may be NULL in the embed calls: every unit
 * then gets the smallest threshold that holds its payload, chosen on the device (error histogram of the unit ->
 * estimate -> embed -> T + 1 for the units that fall short, only those are embedded again); it comes back in
 * info[u][0], a unit that fits at no threshold up to 2^(bit_depth-1) keeps status PEEB_E_CAPACITY.
 * info: n_units x 8 int64 = {T, n_bits, capacity, cap0, cap1, n_flagged, sse, status}
 * with status 0 or PEEB_E_CAPACITY (the embed is still the zero-padded embed of
 * the first `capacity` bits, which is what a sweep wants).  The function's own
 * return value reports only launch problems.
 * payload: per unit, at least peeb_payload_bytes(n_bits[u]) readable bytes at
 * payload + u*payload_stride, 4-byte aligned.  marked / lm may be NULL (sweep:
 * statistics only).  lm: h x ceil(w/8) bytes per unit, np.packbits(axis=1).  */
#define PEEB_INFO 8
size_t peeb_payload_bytes(int64_t n_bits); /* round_up(ceil(n_bits/8),4)+8 */
int peeb_pee_embed_batch(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w,
                         int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits,
                         const uint8_t* payload, int64_t payload_stride, void* marked,
                         int64_t marked_stride, uint8_t* lm, int64_t lm_stride, int64_t* info,
                         void* stream);
/* payload_out: n_units rows of payload_stride writable bytes each (>= peeb_payload_bytes(n_bits[u])), 4-byte
 * aligned; every row comes back zero padded up to the stride.  One unit: payload_stride may be smaller than
 * peeb_payload_bytes(n_bits[0]) (e.g. 0), the buffer then is peeb_payload_bytes(n_bits[0]) bytes long. */
int peeb_pee_extract_batch(peeb_ws* ws, const void* marked, int64_t marked_stride, int n_units, int h,
                           int w, int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits,
                           const uint8_t* lm, int64_t lm_stride, uint8_t* payload_out,
                           int64_t payload_stride, void* recovered, int64_t recovered_stride,
                           int64_t* info, void* stream);
/* hist: n_units x 2 x (2*tmax) uint32, hist[u][c][e+tmax], tmax = 2^(bit_depth-1):
 * prediction errors of the original image per colour over interior pixels not
 * flagged for expansion (Appendix A, threshold selection). */
int peeb_pee_hist_batch(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w,
                        int itemsize, int bit_depth, uint32_t* hist, void* stream);

/* host-buffer variants: contiguous batches, copies inside, synchronous.
 * payload_stride as above (host bytes); lm_stride = h*ceil(w/8).
 * shared_flags: bit 0 = every unit embeds into the one image at src_host, bit 1 = every unit reads
 * the one payload row at payload_host (both: the threshold sweep of one image). */
int peeb_pee_embed_h(peeb_ws* ws, const void* src_host, int shared_flags, int n_units, int h, int w,
                     int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits,
                     const uint8_t* payload_host, int64_t payload_stride, void* marked_host,
                     uint8_t* lm_host, int64_t* info_host);
int peeb_pee_extract_h(peeb_ws* ws, const void* marked_host, int n_units, int h, int w, int itemsize,
                       int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* lm_host,
                       uint8_t* payload_out_host, int64_t payload_stride, void* recovered_host,
                       int64_t* info_host);
int peeb_pee_hist_h(peeb_ws* ws, const void* src_host, int n_units, int h, int w, int itemsize,
                    int bit_depth, uint32_t* hist_host);

/* ---- N1: PEE with the causal MED predictor ----------------- *
 * Not in the reference; specified in DESIGN.md "Appendix A2":
 * p = clamp(W + N - NW, min(W, N), max(W, N)) (JPEG-LS MED), domain i >= 1, j >= 1,
 * ONE raster-order pass; classes / flags / carriers / padding as in Appendix A.
 * The embedder predicts from original pixels (parallel); the extractor from
 * recovered ones (anti-diagonal wavefront, one CTA per image).  Same argument
 * lists as peeb_pee_embed_batch / peeb_pee_extract_batch and their host-buffer
 * forms (src_stride != 0: no shared cover); info: cap0 = capacity, cap1 = 0.
 * T may be NULL in the embed calls: every unit is embedded at T = 1, 2, ... until its
 * payload fits (only the units that fall short take part in a round); T in info[u][0]. */
int peeb_pee_med_embed_batch(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w,
                             int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits,
                             const uint8_t* payload, int64_t payload_stride, void* marked, int64_t marked_stride,
                             uint8_t* lm, int64_t lm_stride, int64_t* info, void* stream);
int peeb_pee_med_extract_batch(peeb_ws* ws, const void* marked, int64_t marked_stride, int n_units, int h,
                               int w, int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits,
                               const uint8_t* lm, int64_t lm_stride, uint8_t* payload_out, int64_t payload_stride,
                               void* recovered, int64_t recovered_stride, int64_t* info, void* stream);
int peeb_pee_med_embed_h(peeb_ws* ws, const void* src_host, int n_units, int h, int w, int itemsize, int bit_depth,
                         const int32_t* T, const int64_t* n_bits, const uint8_t* payload_host, int64_t payload_stride,
                         void* marked_host, uint8_t* lm_host, int64_t* info_host);
int peeb_pee_med_extract_h(peeb_ws* ws, const void* marked_host, int n_units, int h, int w, int itemsize,
                           int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* lm_host,
                           uint8_t* payload_out_host, int64_t payload_stride, void* recovered_host, int64_t* info_host);

/* ---- N2: coding of the side bitmaps --------------------------- *
 * Replaces, as an opt-in format, the reference's blob steps around its container:
 * zlib.compress(np.stack(bitmaps).tobytes()) (src/codec.py:888-889) and
 * np.frombuffer(zlib.decompress(blob), uint8) (src/codec.py:820-821), which push one
 * BYTE per pixel and plane through a serial host coder.  Format "PBR1" (bit packing +
 * three levels of 32-way zero-run elimination; layout in csrc/peeb_bitcode.cu and
 * oracle/bitcode_numpy.py).  src: n map elements, one uint8 each (non-zero = 1), or
 * with packed_input != 0 the np.packbits form (ceil(n/8) bytes, e.g. a PEE location
 * map).  The blob needs peeb_bitmap_blob_bound(n) bytes of capacity; the encoded size
 * comes back in *blob_bytes (HOST pointer; the call synchronises the stream).
 * decode writes n bytes of 0/1 (or ceil(n/8) packed bytes) and fails with
 * PEEB_E_INVALID on a blob whose header, size or level tables are inconsistent.     */
size_t peeb_bitmap_blob_bound(int64_t n);
int peeb_bitmap_encode(peeb_ws* ws, const uint8_t* src, int64_t n, int packed_input, uint8_t* blob,
                       int64_t blob_capacity, int64_t* blob_bytes, void* stream);
int peeb_bitmap_decode(peeb_ws* ws, const uint8_t* blob, int64_t blob_bytes, uint8_t* dst, int64_t n,
                       int packed_output, void* stream);
int peeb_bitmap_encode_h(peeb_ws* ws, const uint8_t* src_host, int64_t n, int packed_input, uint8_t* blob_host,
                         int64_t blob_capacity, int64_t* blob_bytes);
int peeb_bitmap_decode_h(peeb_ws* ws, const uint8_t* blob_host, int64_t blob_bytes, uint8_t* dst_host, int64_t n,
                         int packed_output);

/* ---- bounds-checked build (development aid; no reference counterpart) ---------------------------------------
 * A library compiled with -DPEEB_DEBUG_BOUNDS (python -m codec_tcc_b200.build --bounds ->
 * codec_tcc_b200/lib/bounds/libpeeb200.so) compares every shared- and global-memory index of the PEE band kernels
 * with the size of the region it points into, counts violations and carries on.  out6 = {violations, site of the
 * first one, its byte offset, its limit, warp items checked, 1 if this build checks at all}; the call synchronises
 * the device.  A normal build answers all zeros.  reset != 0 clears the counters afterwards; reset == 2 first runs
 * the checker's self-test (one check that holds, one that does not: exactly one more violation, site 99). */
int peeb_debug_bounds(unsigned long long* out6, int reset);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* PEEB200_H */
