// Shared between the PEE translation units (peeb_pee.cu: entry points, host-buffer pipeline, histogram;
// peeb_pee2.cu: rhombus-predictor kernels; peeb_pee_med.cu: causal-predictor kernels): batch descriptor,
// look-back status encoding, host helpers.
#pragma once
#include "peeb_common.cuh"

namespace peeb {

constexpr unsigned long long ST_AGG = 1ull << 62, ST_PFX = 2ull << 62, ST_MASK = 3ull << 62;

struct PeeBatch {
    const unsigned char* src; long long src_stride;    // bytes
    unsigned char* dst; long long dst_stride;          // may be null
    unsigned char* lm; long long lm_stride;            // may be null (embed) / const (extract)
    const unsigned char* payload; long long payload_stride;
    unsigned char* payload_out;                        // extract
    const int* T;                                      // device, per unit
    const unsigned* n_bits;                            // device, per unit
    long long* info;                                   // device, per unit x 8
    unsigned long long* steps;                         // embed, optional: {warp-steps, at a border, redone} counters
    const int* active;                                 // embed, optional: units with 0 are skipped (threshold search)
    int n_units;
};

// location-map bit of column j inside a 32-bit little-endian word of a packbits row
__device__ __forceinline__ unsigned lm_bitmask(int j) { return 1u << (8 * ((j >> 3) & 3) + 7 - (j & 7)); }

// host side shared by the two translation units
int upload_unit_tables(peeb_ws* ws, int slot, int n_units, const int32_t* T, const int64_t* n_bits, int bit_depth,
                       size_t extra_bytes, cudaStream_t st, int** dT, unsigned** dN, char** extra);
int embed_batch_impl2(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                      int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* payload,
                      int64_t payload_stride, void* marked, int64_t marked_stride, uint8_t* lm, int64_t lm_stride,
                      int64_t* info, cudaStream_t st, int slot);
// T = NULL: thresholds are chosen on the device (Appendix A: histogram estimate, then verify and increment)
int hist_batch_impl2(peeb_ws* ws, const void* src, int64_t src_stride, int n_units, int h, int w, int itemsize,
                     int bit_depth, uint32_t* hist, cudaStream_t st);
int threshold_retry_round(int n_units, int tmax, long long* info, int* T, int* active, int* remaining, int* remaining_h, cudaStream_t st);
int extract_batch_impl2(peeb_ws* ws, const void* marked, int64_t marked_stride, int n_units, int h, int w,
                        int itemsize, int bit_depth, const int32_t* T, const int64_t* n_bits, const uint8_t* lm,
                        int64_t lm_stride, uint8_t* payload_out, int64_t payload_stride, void* recovered,
                        int64_t recovered_stride, int64_t* info, cudaStream_t st, int slot);

}  // namespace peeb
