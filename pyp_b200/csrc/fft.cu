// fft.cu — batched small-box FFTs written for sm_100a (no cuFFT on the product path).
//
// Building block: mixed-radix (4/2/3) Stockham autosort passes on lines held in shared memory,
// ping-ponging between two buffers, twiddles from a double-precision-built table.
//   * rows kernel : two real rows are packed into one complex line (Z = a + i b) and separated
//                   after the transform, so an n x n real image costs n/2 complex FFTs.
//   * lines kernel: strided complex lines (columns of a 2-D spectrum, y / z lines of a 3-D one);
//                   a CTA takes a tile of T adjacent lines so every global access is a
//                   T*8-byte contiguous segment.
// Replaces the FFT stage inside external/cistem2/{refine3d,reconstruct3d,merge3d}
// (closed; contract src/pyp/refine/frealign/frealign.py:3918-3994,1780-1824,2075-2093).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include "fft_smem.cuh"
#include "internal.cuh"

namespace {

using namespace fftsm;

// ---------------------------------------------------------------- strided complex lines
// line (o, t): element e at data[o*ostride + e*estride + t], t in [0, ninner)
// sign_mode 1: multiply output by (-1)^(e + t%inner_w + t/inner_w)  (real-space centring)
template <int DIR>
__global__ void fft_lines_kernel(float2 *__restrict__ data, int n, long long estride, int ninner,
                                 long long ostride, int T, int ntiles, Radices rad,
                                 const float2 *__restrict__ tw_g, float scale, int sign_mode,
                                 int inner_w) {
    extern __shared__ float2 smem[];
    const int pitch = n + 1;
    float2 *tw = smem;
    float2 *bufa = smem + n;
    float2 *bufb = bufa + (size_t)T * pitch;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int outer = blockIdx.x / ntiles;
    const int t0 = (blockIdx.x - outer * ntiles) * T;
    const int nl = min(T, ninner - t0);
    float2 *base = data + (long long)outer * ostride + t0;
    for (int i = tid; i < n; i += nt) tw[i] = tw_g[i];
    for (int idx = tid; idx < n * T; idx += nt) {
        const int e = idx / T, t = idx - e * T;
        if (t < nl) bufa[t * pitch + e] = base[(long long)e * estride + t];
    }
    __syncthreads();
    float2 *res = fft_lines_smem<DIR>(bufa, bufb, pitch, nl, n, rad, tw, tid, nt);
    for (int idx = tid; idx < n * T; idx += nt) {
        const int e = idx / T, t = idx - e * T;
        if (t < nl) {
            float2 v = res[t * pitch + e];
            float s = scale;
            if (sign_mode) {
                const int ti = t0 + t;
                if ((e + ti % inner_w + ti / inner_w) & 1) s = -s;
            }
            v.x *= s;
            v.y *= s;
            base[(long long)e * estride + t] = v;
        }
    }
}

// ---------------------------------------------------------------- real rows -> half spectrum
// rows are contiguous: row r at in[r*n]; out row r at out[r*(n/2+1)]; a CTA takes PR row pairs.
__global__ void fft_rows_r2c_kernel(const float *__restrict__ in, float2 *__restrict__ out, int n,
                                    long long n_rows, int PR, Radices rad,
                                    const float2 *__restrict__ tw_g, int rows_per_image,
                                    const float *__restrict__ offs, const float *__restrict__ scls) {
    extern __shared__ float2 smem[];
    const int pitch = n + 1;
    float2 *tw = smem;
    float2 *bufa = smem + n;
    float2 *bufb = bufa + (size_t)PR * pitch;
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long pair0 = (long long)blockIdx.x * PR;
    const long long n_pairs = n_rows / 2;
    const int np_ = (int)min((long long)PR, n_pairs - pair0);
    for (int i = tid; i < n; i += nt) tw[i] = tw_g[i];
    for (int idx = tid; idx < np_ * n; idx += nt) {
        const int p = idx / n, e = idx - p * n;
        const long long ra = (pair0 + p) * 2;
        float a = in[ra * n + e], b = in[(ra + 1) * n + e];
        if (scls) {
            const long long img = ra / rows_per_image;
            const float o = offs[img], s = scls[img];
            a = (a - o) * s;
            b = (b - o) * s;
        }
        bufa[p * pitch + e] = make_float2(a, b);
    }
    __syncthreads();
    float2 *res = fft_lines_smem<-1>(bufa, bufb, pitch, np_, n, rad, tw, tid, nt);
    const int nh = n / 2 + 1;
    for (int idx = tid; idx < np_ * nh; idx += nt) {
        const int p = idx / nh, k = idx - p * nh;
        const float2 za = res[p * pitch + k];
        float2 zb = res[p * pitch + ((n - k) % n)];
        zb.y = -zb.y;
        const float2 fa = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y + zb.y));
        const float2 df = make_float2(za.x - zb.x, za.y - zb.y);
        const float2 fb = make_float2(0.5f * df.y, -0.5f * df.x);
        const long long ra = (pair0 + p) * 2;
        out[ra * nh + k] = fa;
        out[(ra + 1) * nh + k] = fb;
    }
}

// half spectrum rows -> real rows (unnormalised inverse)
__global__ void fft_rows_c2r_kernel(const float2 *__restrict__ in, float *__restrict__ out, int n,
                                    long long n_rows, int PR, Radices rad,
                                    const float2 *__restrict__ tw_g, float scale) {
    extern __shared__ float2 smem[];
    const int pitch = n + 1;
    float2 *tw = smem;
    float2 *bufa = smem + n;
    float2 *bufb = bufa + (size_t)PR * pitch;
    const int tid = threadIdx.x, nt = blockDim.x;
    const long long pair0 = (long long)blockIdx.x * PR;
    const long long n_pairs = n_rows / 2;
    const int np_ = (int)min((long long)PR, n_pairs - pair0);
    const int nh = n / 2 + 1;
    for (int i = tid; i < n; i += nt) tw[i] = tw_g[i];
    for (int idx = tid; idx < np_ * nh; idx += nt) {
        const int p = idx / nh, k = idx - p * nh;
        const long long ra = (pair0 + p) * 2;
        float2 fa = in[ra * nh + k], fb = in[(ra + 1) * nh + k];
        if (k == 0 || 2 * k == n) {  // self-conjugate bins carry no imaginary part
            fa.y = 0.f;
            fb.y = 0.f;
        }
        // Z[k] = fa + i fb ; Z[n-k] = conj(fa) + i conj(fb)
        bufa[p * pitch + k] = make_float2(fa.x - fb.y, fa.y + fb.x);
        if (k > 0 && 2 * k < n) bufa[p * pitch + n - k] = make_float2(fa.x + fb.y, -fa.y + fb.x);
    }
    __syncthreads();
    float2 *res = fft_lines_smem<+1>(bufa, bufb, pitch, np_, n, rad, tw, tid, nt);
    for (int idx = tid; idx < np_ * n; idx += nt) {
        const int p = idx / n, e = idx - p * n;
        const long long ra = (pair0 + p) * 2;
        const float2 v = res[p * pitch + e];
        out[ra * n + e] = v.x * scale;
        out[(ra + 1) * n + e] = v.y * scale;
    }
}

int pick_tile(int n) {
    // keep 2 buffers of T lines within ~96 KB so two CTAs fit per SM
    int T = 16;
    while (T > 4 && (size_t)2 * T * (n + 1) * 8 > 96 * 1024) T /= 2;
    return T;
}

template <class K> int set_smem(cspb_ctx *ctx, K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        CU_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

int launch_lines(cspb_ctx *ctx, float2 *data, int n, long long estride, int ninner, long long ostride,
                 int n_outer, int dir, float scale, int sign_mode, int inner_w) {
    Radices rad;
    if (!factor(n, rad)) return cspb_fail(ctx, CSPB_E_ARG, "FFT length %d is not 2^a*3^b", n);
    const float2 *tw;
    int rc = fft_get_twiddles(ctx, n, &tw);
    if (rc) return rc;
    const int T = pick_tile(n);
    const size_t smem = ((size_t)n + (size_t)2 * T * (n + 1)) * sizeof(float2);
    const int ntiles = ceil_div(ninner, T);
    const unsigned grid = (unsigned)((long long)ntiles * n_outer);
    if (dir < 0) {
        if ((rc = set_smem(ctx, fft_lines_kernel<-1>, smem))) return rc;
        fft_lines_kernel<-1><<<grid, 256, smem, ctx->stream>>>(data, n, estride, ninner, ostride, T, ntiles, rad, tw,
                                                               scale, sign_mode, inner_w);
    } else {
        if ((rc = set_smem(ctx, fft_lines_kernel<+1>, smem))) return rc;
        fft_lines_kernel<+1><<<grid, 256, smem, ctx->stream>>>(data, n, estride, ninner, ostride, T, ntiles, rad, tw,
                                                               scale, sign_mode, inner_w);
    }
    KERNEL_CHECK(ctx);
    return 0;
}

int launch_rows_r2c(cspb_ctx *ctx, const float *in, float2 *out, int n, long long n_rows, int rows_per_image,
                    const float *offs, const float *scls) {
    Radices rad;
    if (!factor(n, rad) || (n & 1)) return cspb_fail(ctx, CSPB_E_ARG, "FFT length %d unsupported", n);
    const float2 *tw;
    int rc = fft_get_twiddles(ctx, n, &tw);
    if (rc) return rc;
    const int PR = pick_tile(n);
    const size_t smem = ((size_t)n + (size_t)2 * PR * (n + 1)) * sizeof(float2);
    if ((rc = set_smem(ctx, fft_rows_r2c_kernel, smem))) return rc;
    const long long n_pairs = n_rows / 2;
    fft_rows_r2c_kernel<<<ceil_div(n_pairs, PR), 256, smem, ctx->stream>>>(in, out, n, n_rows, PR, rad, tw,
                                                                          rows_per_image, offs, scls);
    KERNEL_CHECK(ctx);
    return 0;
}

int launch_rows_c2r(cspb_ctx *ctx, const float2 *in, float *out, int n, long long n_rows, float scale) {
    Radices rad;
    if (!factor(n, rad) || (n & 1)) return cspb_fail(ctx, CSPB_E_ARG, "FFT length %d unsupported", n);
    const float2 *tw;
    int rc = fft_get_twiddles(ctx, n, &tw);
    if (rc) return rc;
    const int PR = pick_tile(n);
    const size_t smem = ((size_t)n + (size_t)2 * PR * (n + 1)) * sizeof(float2);
    if ((rc = set_smem(ctx, fft_rows_c2r_kernel, smem))) return rc;
    const long long n_pairs = n_rows / 2;
    fft_rows_c2r_kernel<<<ceil_div(n_pairs, PR), 256, smem, ctx->stream>>>(in, out, n, n_rows, PR, rad, tw, scale);
    KERNEL_CHECK(ctx);
    return 0;
}

}  // namespace

int fft_get_twiddles(cspb_ctx *ctx, int n, const float2 **tw_out) {
    for (size_t i = 0; i < ctx->tw_n.size(); ++i)
        if (ctx->tw_n[i] == n) {
            *tw_out = ctx->d_tw.as<float2>() + ctx->tw_off[i];
            return 0;
        }
    const size_t cap = 16384;  // float2 entries; enough for a handful of sizes up to 2048
    if (!ctx->d_tw.p) RESERVE(ctx, ctx->d_tw, cap * sizeof(float2));
    if (ctx->tw_used + (size_t)n > cap) return cspb_fail(ctx, CSPB_E_NOMEM, "twiddle cache full");
    std::vector<float2> h(n);
    for (int k = 0; k < n; ++k) {
        const double a = -2.0 * CSPB_PI_D * (double)k / (double)n;
        h[k] = make_float2((float)cos(a), (float)sin(a));
    }
    float2 *dst = ctx->d_tw.as<float2>() + ctx->tw_used;
    CU_TRY(ctx, cudaMemcpyAsync(dst, h.data(), n * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->tw_n.push_back(n);
    ctx->tw_off.push_back(ctx->tw_used);
    ctx->tw_used += n;
    *tw_out = dst;
    return 0;
}

int fft2_r2c_dev(cspb_ctx *ctx, const float *in, float2 *out, int n, int batch, const float *offs,
                 const float *scls) {
    int rc = launch_rows_r2c(ctx, in, out, n, (long long)batch * n, n, offs, scls);
    if (rc) return rc;
    const int nh = n / 2 + 1;
    return launch_lines(ctx, out, n, nh, nh, (long long)n * nh, batch, -1, 1.f, 0, nh);
}

int fft2_c2r_dev(cspb_ctx *ctx, float2 *inout_c, float *out, int n, int batch) {
    const int nh = n / 2 + 1;
    int rc = launch_lines(ctx, inout_c, n, nh, nh, (long long)n * nh, batch, +1, 1.f, 0, nh);
    if (rc) return rc;
    return launch_rows_c2r(ctx, inout_c, out, n, (long long)batch * n, 1.f);
}

int fft3_r2c_dev(cspb_ctx *ctx, const float *in, float2 *out, int np) {
    const int xh = np / 2 + 1;
    int rc = launch_rows_r2c(ctx, in, out, np, (long long)np * np, np * np, nullptr, nullptr);
    if (rc) return rc;
    // y lines: element stride xh, inner = x, outer = z
    rc = launch_lines(ctx, out, np, xh, xh, (long long)xh * np, np, -1, 1.f, 0, xh);
    if (rc) return rc;
    // z lines: element stride xh*np, inner = (y,x) flattened
    return launch_lines(ctx, out, np, (long long)xh * np, xh * np, 0, 1, -1, 1.f, 0, xh);
}

int fft3_c2r_dev(cspb_ctx *ctx, float2 *inout_c, float *out, int np) {
    const int xh = np / 2 + 1;
    int rc = launch_lines(ctx, inout_c, np, (long long)xh * np, xh * np, 0, 1, +1, 1.f, 0, xh);
    if (rc) return rc;
    rc = launch_lines(ctx, inout_c, np, xh, xh, (long long)xh * np, np, +1, 1.f, 0, xh);
    if (rc) return rc;
    return launch_rows_c2r(ctx, inout_c, out, np, (long long)np * np, 1.f);
}
