// K5: fused integer moment reduction for MSE / PSNR / global SSIM / difference
// statistics (rows a1-a4).  One streaming pass over both images, 128-bit loads,
// exact 64-bit integer sums; HBM bound (2*itemsize algorithmic bytes per pixel).
#include <algorithm>
#include <cmath>

#include "peeb_common.cuh"

namespace peeb {

struct Acc {
    unsigned long long sse = 0, sad = 0, ne = 0, sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
    unsigned maxd = 0, maxa = 0, maxb = 0;
    // FULL=false keeps only what calcular_mse needs when both maxima agree: SSE and the two maxima
    template <bool FULL>
    __device__ __forceinline__ void add(unsigned a, unsigned b) {
        const int d = (int)a - (int)b;
        if (!FULL) {
            sse += (unsigned long long)((long long)d * d);
            maxa = max(maxa, a);
            maxb = max(maxb, b);
            return;
        }
        const unsigned ad = (unsigned)(d < 0 ? -d : d);
        sse += (unsigned long long)ad * ad;
        saa += (unsigned long long)a * a;
        sbb += (unsigned long long)b * b;
        sab += (unsigned long long)a * b;
        sad32 += ad;
        sa32 += a;
        sb32 += b;
        ne32 += (d != 0);
        maxd = max(maxd, ad);
        maxa = max(maxa, a);
        maxb = max(maxb, b);
    }
    // 32-bit partial sums, folded every <= 32768 elements (65535 * 32768 < 2^32)
    unsigned sad32 = 0, sa32 = 0, sb32 = 0, ne32 = 0;
    __device__ __forceinline__ void fold() {
        sad += sad32; sa += sa32; sb += sb32; ne += ne32;
        sad32 = sa32 = sb32 = ne32 = 0;
    }
};

template <int ITEM, bool FULL>
__device__ __forceinline__ void add_vec(Acc& acc, const int4& va, const int4& vb) {
    const unsigned wa[4] = {(unsigned)va.x, (unsigned)va.y, (unsigned)va.z, (unsigned)va.w};
    const unsigned wb[4] = {(unsigned)vb.x, (unsigned)vb.y, (unsigned)vb.z, (unsigned)vb.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (ITEM == 2) {
            acc.template add<FULL>(wa[k] & 0xffffu, wb[k] & 0xffffu);
            acc.template add<FULL>(wa[k] >> 16, wb[k] >> 16);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc.template add<FULL>((wa[k] >> (8 * j)) & 0xffu, (wb[k] >> (8 * j)) & 0xffu);
        }
    }
}

// grid = (blocks_per_image, n_images); out zeroed by the caller.
template <int ITEM, bool FULL>
__global__ void __launch_bounds__(256) moments_kernel(const unsigned char* __restrict__ a,
                                                      const unsigned char* __restrict__ b, long long n,
                                                      long long stride_a, long long stride_b,
                                                      unsigned long long* __restrict__ out) {
    const unsigned char* pa = a + (long long)blockIdx.y * stride_a * ITEM;
    const unsigned char* pb = b + (long long)blockIdx.y * stride_b * ITEM;
    unsigned long long* o = out + (long long)blockIdx.y * PEEB_MOMENTS;
    constexpr int PER_VEC = 16 / ITEM;
    Acc acc;
    const bool aligned = ((((uintptr_t)pa) | ((uintptr_t)pb)) & 15) == 0;
    const long long nvec = aligned ? n / PER_VEC : 0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthr = (long long)gridDim.x * blockDim.x;
    const int4* qa = reinterpret_cast<const int4*>(pa);
    const int4* qb = reinterpret_cast<const int4*>(pb);
    int since_fold = 0;
    long long i = tid;
    // 4 independent 128-bit loads per array in flight per thread
    for (; i + 3 * nthr < nvec; i += 4 * nthr) {
        int4 va[4], vb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { va[u] = ldg_stream(qa + i + u * nthr); vb[u] = ldg_stream(qb + i + u * nthr); }
#pragma unroll
        for (int u = 0; u < 4; ++u) add_vec<ITEM, FULL>(acc, va[u], vb[u]);
        if (++since_fold == 256) { acc.fold(); since_fold = 0; }
    }
    for (; i < nvec; i += nthr) add_vec<ITEM, FULL>(acc, ldg_stream(qa + i), ldg_stream(qb + i));
    acc.fold();
    // scalar tail (or everything, when unaligned)
    for (long long e = nvec * PER_VEC + tid; e < n; e += nthr) {
        unsigned x, y;
        if (ITEM == 2) { x = reinterpret_cast<const unsigned short*>(pa)[e]; y = reinterpret_cast<const unsigned short*>(pb)[e]; }
        else { x = pa[e]; y = pb[e]; }
        acc.template add<FULL>(x, y);
        acc.fold();
    }

    // block reduction: warp shuffles, then one atomic per block and quantity
    unsigned long long sums[8] = {acc.sse, acc.sad, acc.ne, acc.sa, acc.sb, acc.saa, acc.sbb, acc.sab};
    unsigned maxs[3] = {acc.maxd, acc.maxa, acc.maxb};
#pragma unroll
    for (int k = 0; k < 8; ++k) sums[k] = warp_sum_u64(sums[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) maxs[k] = max(maxs[k], __shfl_xor_sync(0xffffffffu, maxs[k], ofs));
    __shared__ unsigned long long s_sum[8][8];
    __shared__ unsigned s_max[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s_sum[k][warp] = sums[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) s_max[k][warp] = maxs[k];
    }
    __syncthreads();
    if (threadIdx.x < 11) {
        const int k = threadIdx.x;
        // out order: sse sad maxd ne sa sb saa sbb sab maxa maxb
        if (k < 8) {
            unsigned long long t = 0;
            for (int wv = 0; wv < 8; ++wv) t += s_sum[k][wv];
            const int slot = (k == 0) ? 0 : (k == 1) ? 1 : (k == 2) ? 3 : k + 1;  // ne->3, sa->4 ... sab->8
            atomicAdd(o + slot, t);
        } else {
            const int m = k - 8;
            unsigned t = 0;
            for (int wv = 0; wv < 8; ++wv) t = max(t, s_max[m][wv]);
            const int slot = (m == 0) ? 2 : (m == 1) ? 9 : 10;
            atomicMax(o + slot, (unsigned long long)t);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) o[11] = (unsigned long long)n;
}

static int launch_moments(peeb_ws* ws, const void* a, const void* b, int64_t n, int itemsize, int n_images,
                          int64_t stride_a, int64_t stride_b, int64_t* out, cudaStream_t st, bool full = true) {
    PEEB_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t) * PEEB_MOMENTS * n_images, st));
    if (n <= 0) return PEEB_OK;
    // enough blocks to fill the machine (multiples of the SM count), split over the images
    const long long per_block = 256LL * (16 / itemsize) * 16;
    long long want = (n + per_block - 1) / per_block;
    long long cap = (long long)ws->sm_count * 8 / (n_images < 1 ? 1 : n_images);
    if (cap < 1) cap = 1;
    if (want > cap) want = cap;
    dim3 grid((unsigned)want, (unsigned)n_images);
    ProfScope prof(ws, PEEB_K_MOMENTS, st);
    const unsigned char* pa = (const unsigned char*)a;
    const unsigned char* pb = (const unsigned char*)b;
    unsigned long long* po = (unsigned long long*)out;
    if (itemsize == 2 && full) moments_kernel<2, true><<<grid, 256, 0, st>>>(pa, pb, n, stride_a, stride_b, po);
    else if (itemsize == 2) moments_kernel<2, false><<<grid, 256, 0, st>>>(pa, pb, n, stride_a, stride_b, po);
    else if (full) moments_kernel<1, true><<<grid, 256, 0, st>>>(pa, pb, n, stride_a, stride_b, po);
    else moments_kernel<1, false><<<grid, 256, 0, st>>>(pa, pb, n, stride_a, stride_b, po);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

}  // namespace peeb

using namespace peeb;

extern "C" {

int peeb_moments_batch(peeb_ws* ws, const void* a, const void* b, int64_t n, int itemsize, int n_images,
                       int64_t stride_a, int64_t stride_b, int64_t* out, void* stream) {
    PEEB_REQUIRE(ws && a && b && out, "peeb_moments_batch: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_moments_batch: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && n_images >= 1 && n_images <= 65535, "peeb_moments_batch: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    return launch_moments(ws, a, b, n, itemsize, n_images, stride_a, stride_b, out, (cudaStream_t)stream);
}

int peeb_sse_batch(peeb_ws* ws, const void* a, const void* b, int64_t n, int itemsize, int n_images,
                   int64_t stride_a, int64_t stride_b, int64_t* out, void* stream) {
    PEEB_REQUIRE(ws && a && b && out, "peeb_sse_batch: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_sse_batch: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0 && n_images >= 1 && n_images <= 65535, "peeb_sse_batch: bad sizes");
    PEEB_CUDA(cudaSetDevice(ws->device));
    return launch_moments(ws, a, b, n, itemsize, n_images, stride_a, stride_b, out, (cudaStream_t)stream, false);
}

static int moments_host(peeb_ws* ws, const void* a_host, const void* b_host, int64_t n, int itemsize,
                        int64_t* out_host, bool full);

int peeb_moments_h(peeb_ws* ws, const void* a_host, const void* b_host, int64_t n, int itemsize,
                   int64_t* out_host) {
    return moments_host(ws, a_host, b_host, n, itemsize, out_host, true);
}

int peeb_sse_h(peeb_ws* ws, const void* a_host, const void* b_host, int64_t n, int itemsize, int64_t* out_host) {
    return moments_host(ws, a_host, b_host, n, itemsize, out_host, false);
}

static int moments_host(peeb_ws* ws, const void* a_host, const void* b_host, int64_t n, int itemsize,
                        int64_t* out_host, bool full) {
    PEEB_REQUIRE(ws && a_host && b_host && out_host, "peeb_moments_h: null pointer");
    PEEB_REQUIRE(itemsize == 1 || itemsize == 2, "peeb_moments_h: itemsize must be 1 or 2");
    PEEB_REQUIRE(n >= 0, "peeb_moments_h: negative size");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const size_t bytes = align_up((size_t)n * itemsize, 256);
    int rc = scratch_reserve(ws->stage, 2 * bytes + 256);
    if (rc) return rc;
    rc = scratch_reserve(ws->tables, 4096);
    if (rc) return rc;
    char* da = (char*)ws->stage.ptr;
    char* db = da + bytes;
    int64_t* dout = (int64_t*)ws->tables.ptr;
    // the two uploads ride different streams so that they overlap on the copy engines
    PEEB_CUDA(cudaMemcpyAsync(da, a_host, (size_t)n * itemsize, cudaMemcpyHostToDevice, ws->stream));
    PEEB_CUDA(cudaMemcpyAsync(db, b_host, (size_t)n * itemsize, cudaMemcpyHostToDevice, ws->stream2));
    PEEB_CUDA(cudaEventRecord(ws->ev[0], ws->stream2));
    PEEB_CUDA(cudaStreamWaitEvent(ws->stream, ws->ev[0], 0));
    rc = launch_moments(ws, da, db, n, itemsize, 1, 0, 0, dout, ws->stream, full);
    if (rc) return rc;
    PEEB_CUDA(cudaMemcpyAsync(out_host, dout, sizeof(int64_t) * PEEB_MOMENTS, cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    return PEEB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ a1 / a3 / a4 on float64 pixel data
// The reference converts whatever it is given with np.array(img, dtype=np.float64) (src/mse.py:85,91) and
// works element-wise in float64 from there; integer-valued pixel data takes the exact integer kernel above,
// everything else (fractional, negative or > 16-bit values) comes here.  One kernel, run twice by the Python
// side: first plain (maxima, sums, difference statistics), then with the reference's range normalisation
// u = (a / div_a) * mul_a (src/mse.py:104-105, same operation order) and the means of pass one, which gives
// the squared difference and the centred second moments (np.var / covariance of src/mse.py:166-168) without
// cancellation.  Per-CTA partial sums, added up in a fixed order by one CTA: results do not depend on timing.
namespace peeb {

constexpr int F64_OUT = 12;  // sum u, sum v, sum (u-mu)^2, sum (v-mv)^2, sum (u-mu)(v-mv), sum (u-v)^2,
                             // sum |a-b|, max |a-b|, max a, max b, #(a != b), n

struct F64Acc {
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    double maxd = 0.0, maxa = -INFINITY, maxb = -INFINITY;
    double ne = 0.0;
};

__global__ void __launch_bounds__(256) moments_f64_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                          long long n, double div_a, double mul_a, double div_b,
                                                          double mul_b, double mu, double mv, int scaled,
                                                          double* __restrict__ partial) {
    F64Acc acc;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = a[i], y = b[i];
        const double u = scaled ? (x / div_a) * mul_a : x, v = scaled ? (y / div_b) * mul_b : y;
        const double du = u - mu, dv = v - mv, d = u - v, r = fabs(x - y);
        acc.s[0] += u; acc.s[1] += v;
        acc.s[2] += du * du; acc.s[3] += dv * dv; acc.s[4] += du * dv;
        acc.s[5] += d * d; acc.s[6] += r;
        acc.maxd = fmax(acc.maxd, r); acc.maxa = fmax(acc.maxa, x); acc.maxb = fmax(acc.maxb, y);
        acc.ne += (x != y) ? 1.0 : 0.0;
    }
    __shared__ double sh[8][F64_OUT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double vals[F64_OUT] = {acc.s[0], acc.s[1], acc.s[2], acc.s[3], acc.s[4], acc.s[5], acc.s[6],
                            acc.maxd, acc.maxa, acc.maxb, acc.ne, 0.0};
#pragma unroll
    for (int k = 0; k < F64_OUT; ++k) {
        double v = vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, v, o);
            v = (k >= 7 && k <= 9) ? fmax(v, t) : v + t;
        }
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < F64_OUT) {
        const int k = threadIdx.x;
        double v = sh[0][k];
        for (int w = 1; w < 8; ++w) v = (k >= 7 && k <= 9) ? fmax(v, sh[w][k]) : v + sh[w][k];
        partial[(long long)blockIdx.x * F64_OUT + k] = v;
    }
}

__global__ void moments_f64_final_kernel(const double* __restrict__ partial, int nblocks, long long n,
                                         double* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= F64_OUT) return;
    double v = partial[k];
    for (int b = 1; b < nblocks; ++b) {
        const double t = partial[(long long)b * F64_OUT + k];
        v = (k >= 7 && k <= 9) ? fmax(v, t) : v + t;
    }
    out[k] = k == 11 ? (double)n : v;
}

}  // namespace peeb

extern "C" {

int peeb_moments_f64(peeb_ws* ws, const double* a, const double* b, int64_t n, int scaled, double div_a, double mul_a,
                     double div_b, double mul_b, double mean_a, double mean_b, double* out, void* stream) {
    PEEB_REQUIRE(ws && a && b && out, "peeb_moments_f64: null pointer");
    PEEB_REQUIRE(n >= 1, "peeb_moments_f64: n must be >= 1");
    PEEB_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    long long want = (n + 256 * 8 - 1) / (256 * 8);
    const int nblocks = (int)std::min<long long>(std::max<long long>(want, 1), (long long)ws->sm_count * 8);
    int rc = scratch_reserve(ws->tables, (size_t)nblocks * F64_OUT * sizeof(double) + 256);
    if (rc) return rc;
    double* partial = (double*)ws->tables.ptr;
    ProfScope prof(ws, PEEB_K_MOMENTS, st);
    moments_f64_kernel<<<nblocks, 256, 0, st>>>(a, b, n, div_a, mul_a, div_b, mul_b, mean_a, mean_b, scaled, partial);
    moments_f64_final_kernel<<<1, 32, 0, st>>>(partial, nblocks, n, out);
    PEEB_CUDA(cudaGetLastError());
    return PEEB_OK;
}

int peeb_moments_f64_h(peeb_ws* ws, const double* a_host, const double* b_host, int64_t n, double* out_host) {
    PEEB_REQUIRE(ws && a_host && b_host && out_host, "peeb_moments_f64_h: null pointer");
    PEEB_REQUIRE(n >= 1, "peeb_moments_f64_h: n must be >= 1");
    PEEB_CUDA(cudaSetDevice(ws->device));
    const size_t bytes = (size_t)n * sizeof(double), o_b = align_up(bytes, 256), o_out = 2 * o_b;
    int rc = scratch_reserve(ws->stage, o_out + 256);
    if (rc) return rc;
    char* d = (char*)ws->stage.ptr;
    const double* da = (const double*)d;
    const double* db = (const double*)(d + o_b);
    double* dout = (double*)(d + o_out);
    PEEB_CUDA(cudaMemcpyAsync(d, a_host, bytes, cudaMemcpyHostToDevice, ws->stream));
    PEEB_CUDA(cudaMemcpyAsync(d + o_b, b_host, bytes, cudaMemcpyHostToDevice, ws->stream));
    // pass one: plain values (maxima, sums, difference statistics)
    rc = peeb_moments_f64(ws, da, db, n, 0, 1.0, 1.0, 1.0, 1.0, 0.0, 0.0, dout, ws->stream);
    if (rc) { cudaStreamSynchronize(ws->stream); return rc; }
    PEEB_CUDA(cudaMemcpyAsync(out_host, dout, F64_OUT * sizeof(double), cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    // pass two: the reference's range normalisation (only when the maxima differ, src/mse.py:101) and moments
    // centred on (an estimate of) the means -- the caller corrects with the exact means of this pass
    const double maxa = out_host[8], maxb = out_host[9], top = maxa > maxb ? maxa : maxb;
    const int scaled = maxa != maxb;
    const double mu = scaled ? ((out_host[0] / (double)n) / maxa) * top : out_host[0] / (double)n;
    const double mv = scaled ? ((out_host[1] / (double)n) / maxb) * top : out_host[1] / (double)n;
    rc = peeb_moments_f64(ws, da, db, n, scaled, maxa, top, maxb, top, mu, mv, dout, ws->stream);
    if (rc) { cudaStreamSynchronize(ws->stream); return rc; }
    PEEB_CUDA(cudaMemcpyAsync(out_host + F64_OUT, dout, F64_OUT * sizeof(double), cudaMemcpyDeviceToHost, ws->stream));
    PEEB_CUDA(cudaStreamSynchronize(ws->stream));
    out_host[2 * F64_OUT] = mu;
    out_host[2 * F64_OUT + 1] = mv;
    return PEEB_OK;
}

}  // extern "C"
