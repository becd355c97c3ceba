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
#include "device_math.cuh"
#include "fft_fast.cuh"
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
    const int pitch = line_pitch(n);
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
        if (t < nl) bufa[t * pitch + skew(e)] = base[(long long)e * estride + t];
    }
    __syncthreads();
    float2 *res = fft_lines_smem<DIR>(bufa, bufb, pitch, nl, n, rad, tw, tid, nt);
    for (int idx = tid; idx < n * T; idx += nt) {
        const int e = idx / T, t = idx - e * T;
        if (t < nl) {
            float2 v = res[t * pitch + skew(e)];
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
    const int pitch = line_pitch(n);
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
        bufa[p * pitch + skew(e)] = make_float2(a, b);
    }
    __syncthreads();
    float2 *res = fft_lines_smem<-1>(bufa, bufb, pitch, np_, n, rad, tw, tid, nt);
    const int nh = n / 2 + 1;
    for (int idx = tid; idx < np_ * nh; idx += nt) {
        const int p = idx / nh, k = idx - p * nh;
        const float2 za = res[p * pitch + skew(k)];
        float2 zb = res[p * pitch + skew((n - k) % n)];
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
    const int pitch = line_pitch(n);
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
        bufa[p * pitch + skew(k)] = make_float2(fa.x - fb.y, fa.y + fb.x);
        if (k > 0 && 2 * k < n) bufa[p * pitch + skew(n - k)] = make_float2(fa.x + fb.y, -fa.y + fb.x);
    }
    __syncthreads();
    float2 *res = fft_lines_smem<+1>(bufa, bufb, pitch, np_, n, rad, tw, tid, nt);
    for (int idx = tid; idx < np_ * n; idx += nt) {
        const int p = idx / n, e = idx - p * n;
        const long long ra = (pair0 + p) * 2;
        const float2 v = res[p * pitch + skew(e)];
        out[ra * n + e] = v.x * scale;
        out[(ra + 1) * n + e] = v.y * scale;
    }
}

// ================================================================ fast path: n = R1 * R2
// strided complex lines, tile of TL = 16 adjacent lines per CTA (lane = line: 128-byte segments)
// threads per line = max(R1, R2) (R1 <= R2), TL lines per CTA, TL * R2 threads
template <int R2> struct FastTile { static constexpr int TL = R2 <= 16 ? 16 : 256 / R2; };
template <int R1, int R2, int DIR>
__global__ void __launch_bounds__(256) fft_lines_fast_kernel(float2 *__restrict__ data, long long estride, int ninner,
                                                             long long ostride, int ntiles, const float2 *__restrict__ tw_g,
                                                             float scale, int sign_mode, int inner_w,
                                                             const float *__restrict__ filt, int filt_w) {
    constexpr int N = R1 * R2, TL = FastTile<R2>::TL, P = N + (N >> 4) + 1, NT = TL * (R2 < 16 ? 16 : R2);
    __shared__ float2 S[TL * P];
    __shared__ float2 tw[N];
    const int tid = threadIdx.x;
    const int l = tid % TL, t = tid / TL;  // R2 threads per line in stage 1
    const int outer = blockIdx.x / ntiles;
    const int t0 = (blockIdx.x - outer * ntiles) * TL;
    const bool live = t0 + l < ninner;
    float2 *base = data + (long long)outer * ostride + t0 + l;
    for (int i = tid; i < N; i += NT) tw[i] = tw_g[i];
    float2 v[R1 > R2 ? R1 : R2];
    if (t < R2) {
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = live ? base[(long long)(t + R2 * r) * estride] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    if (t < R2) fftfast::stage1<R1, R2, DIR>(v, t, S + l * P, tw);
    __syncthreads();
    if (t < R1) {
        fftfast::stage2<R1, R2, DIR>(v, t, S + l * P);
        if (live) {
            const int ti = t0 + l;
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                const int e = t + R1 * k2;
                float s = scale;
                if (sign_mode && ((e + ti % inner_w + ti / inner_w) & 1)) s = -s;
                if (filt) {  // radial filter: line index = i (x frequency), element = j (y frequency)
                    const int j = e >= N / 2 ? e - N : e;
                    const int i = ti % filt_w;
                    s *= filt[(int)(sqrtf((float)(i * i + j * j)) + 0.5f)];
                }
                base[(long long)e * estride] = make_float2(v[k2].x * s, v[k2].y * s);
            }
        }
    }
}

// real rows -> half spectrum, PR = 16 row pairs per CTA (lane = element: contiguous rows)
template <int R1, int R2>
__global__ void __launch_bounds__(256) fft_rows_r2c_fast_kernel(const float *__restrict__ in, float2 *__restrict__ out,
                                                                long long n_rows, const float2 *__restrict__ tw_g,
                                                                int rows_per_image, const float *__restrict__ offs,
                                                                const float *__restrict__ scls) {
    constexpr int N = R1 * R2, PR = 256 / R2, P = N + (N >> 4) + 1, NH = N / 2 + 1, NT = PR * R2;
    __shared__ float2 S[PR * P];
    __shared__ float2 tw[N];
    const int tid = threadIdx.x;
    const int t = tid % R2, p = tid / R2;
    const long long pair = (long long)blockIdx.x * PR + p;
    const bool live = pair < n_rows / 2;
    for (int i = tid; i < N; i += NT) tw[i] = tw_g[i];
    float2 v[R1 > R2 ? R1 : R2];
    {
        const float *ra = in + pair * 2 * N, *rb = ra + N;
        float o = 0.f, sc = 1.f;
        if (live && scls) {
            const long long img = pair * 2 / rows_per_image;
            o = offs[img];
            sc = scls[img];
        }
#pragma unroll
        for (int r = 0; r < R1; ++r)
            v[r] = live ? make_float2((ra[t + R2 * r] - o) * sc, (rb[t + R2 * r] - o) * sc) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    fftfast::stage1<R1, R2, -1>(v, t, S + p * P, tw);
    __syncthreads();
    if (t < R1) fftfast::stage2<R1, R2, -1>(v, t, S + p * P);
    __syncthreads();
    // natural-order spectrum of the packed line back to shared memory, then untangle the two rows
    if (t < R1) {
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) S[p * P + fftsm::skew(t + R1 * k2)] = v[k2];
    }
    __syncthreads();
    for (int idx = tid; idx < PR * NH; idx += NT) {
        const int pp = idx / NH, k = idx - pp * NH;
        const long long pr = (long long)blockIdx.x * PR + pp;
        if (pr >= n_rows / 2) break;
        const float2 za = S[pp * P + fftsm::skew(k)];
        float2 zb = S[pp * P + fftsm::skew(k ? N - k : 0)];
        zb.y = -zb.y;
        const float2 fa = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y + zb.y));
        const float2 df = make_float2(za.x - zb.x, za.y - zb.y);
        const float2 fb = make_float2(0.5f * df.y, -0.5f * df.x);
        out[pr * 2 * NH + k] = fa;
        out[(pr * 2 + 1) * NH + k] = fb;
    }
}

// half spectrum rows -> real rows (unnormalised inverse), optional real-space multiplier per pixel:
// mask(r) = scale * cosine_edge(r, mask_radius, mask_width) with r from the image centre
template <int R1, int R2>
__global__ void __launch_bounds__(256) fft_rows_c2r_fast_kernel(const float2 *__restrict__ in, float *__restrict__ out,
                                                                long long n_rows, const float2 *__restrict__ tw_g, float scale,
                                                                float mask_radius, float mask_width) {
    constexpr int N = R1 * R2, PR = 256 / R2, P = N + (N >> 4) + 1, NH = N / 2 + 1, NT = PR * R2;
    __shared__ float2 S[PR * P];
    float2 *Z = S;  // the packed spectrum is staged in the same buffer (extra barrier below)
    __shared__ float2 tw[N];
    const int tid = threadIdx.x;
    const int t = tid % R2, p = tid / R2;
    const long long pair = (long long)blockIdx.x * PR + p;
    const bool live = pair < n_rows / 2;
    for (int i = tid; i < N; i += NT) tw[i] = tw_g[i];
    // Z[k] = fa + i fb ; Z[n-k] = conj(fa) + i conj(fb)
    for (int idx = tid; idx < PR * NH; idx += NT) {
        const int pp = idx / NH, k = idx - pp * NH;
        const long long pr = (long long)blockIdx.x * PR + pp;
        if (pr >= n_rows / 2) break;
        float2 fa = in[pr * 2 * NH + k], fb = in[(pr * 2 + 1) * NH + k];
        if (k == 0 || 2 * k == N) {
            fa.y = 0.f;
            fb.y = 0.f;
        }
        Z[pp * P + fftsm::skew(k)] = make_float2(fa.x - fb.y, fa.y + fb.x);
        if (k > 0 && 2 * k < N) Z[pp * P + fftsm::skew(N - k)] = make_float2(fa.x + fb.y, -fa.y + fb.x);
    }
    __syncthreads();
    float2 v[R1 > R2 ? R1 : R2];
#pragma unroll
    for (int r = 0; r < R1; ++r) v[r] = live ? Z[p * P + fftsm::skew(t + R2 * r)] : make_float2(0.f, 0.f);
    __syncthreads();
    fftfast::stage1<R1, R2, +1>(v, t, S + p * P, tw);
    __syncthreads();
    if (t < R1 && live) {
        fftfast::stage2<R1, R2, +1>(v, t, S + p * P);
        float *ra = out + pair * 2 * N, *rb = ra + N;
        const int ya = (int)((pair * 2) % N) - N / 2, yb = ya + 1;
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
            const int e = t + R1 * k2;
            float sa = scale, sb = scale;
            if (mask_width > 0.f) {
                const float x = (float)(e - N / 2);
                sa *= cosine_edge(sqrtf(x * x + (float)(ya * ya)), mask_radius, mask_width);
                sb *= cosine_edge(sqrtf(x * x + (float)(yb * yb)), mask_radius, mask_width);
            }
            ra[e] = v[k2].x * sa;
            rb[e] = v[k2].y * sb;
        }
    }
}

// ---- fused passes of the particle preprocessing (refine.cu, cspb_refine_load_images): the whitening and masking round
// trips stay on the SM.  Same butterflies in the same order as the separate passes above — results are bit-identical —
// with one more exchange through shared memory where a forward transform hands over to an inverse one (the second stage
// leaves element u + R1 k2 in thread u, the first stage of the next transform wants element t + R2 r in thread t).

// columns: forward FFT -> radial filter (whitening) -> inverse FFT, in place.  Persistent CTAs (grid = a multiple of the SM
// count) walk the column tiles; the loads of the NEXT tile are issued into a second register set before the current tile is
// transformed, so that every CTA has a tile's worth of DRAM reads in flight all the time (the one-tile-per-CTA version was
// latency bound at 2.8 TB/s of traffic: profiles/r02p_prep_ncu.txt).
template <int R1, int R2, bool KEEP>
__global__ void __launch_bounds__(256, 2) fft_cols_filter_fast_kernel(float2 *__restrict__ data, long long estride, int ninner, long long ostride,
                                                                   int ntiles, int total_tiles, const float2 *__restrict__ tw_g,
                                                                   const float *__restrict__ filt, int filt_w, float2 *__restrict__ keep) {
    constexpr int N = R1 * R2, TL = FastTile<R2>::TL, P = N + (N >> 4) + 1, NT = TL * (R2 < 16 ? 16 : R2);
    __shared__ float2 S[TL * P];
    __shared__ float2 tw[N];
    const int tid = threadIdx.x;
    const int l = tid % TL, t = tid / TL;
    for (int i = tid; i < N; i += NT) tw[i] = tw_g[i];
    float2 v[R1 > R2 ? R1 : R2], vn[R1];
    int tile = blockIdx.x;
    {
        const int outer = tile / ntiles, t0 = (tile - outer * ntiles) * TL;
        const float2 *b = data + (long long)outer * ostride + t0 + l;
        if (t < R2) {
#pragma unroll
            for (int r = 0; r < R1; ++r) vn[r] = (tile < total_tiles && t0 + l < ninner) ? b[(long long)(t + R2 * r) * estride] : make_float2(0.f, 0.f);
        }
    }
    for (; tile < total_tiles; tile += gridDim.x) {
        const int outer = tile / ntiles;
        const int t0 = (tile - outer * ntiles) * TL;
        const bool live = t0 + l < ninner;
        float2 *base = data + (long long)outer * ostride + t0 + l;
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = vn[r];
        {   // prefetch
            const int nt_ = tile + gridDim.x;
            const int no = nt_ / ntiles, n0 = (nt_ - no * ntiles) * TL;
            const float2 *b = data + (long long)no * ostride + n0 + l;
            if (t < R2 && nt_ < total_tiles && n0 + l < ninner) {
#pragma unroll
                for (int r = 0; r < R1; ++r) vn[r] = b[(long long)(t + R2 * r) * estride];
            } else {
#pragma unroll
                for (int r = 0; r < R1; ++r) vn[r] = make_float2(0.f, 0.f);
            }
        }
        __syncthreads();  // tw loaded (first tile) / the previous tile's last reads of S are done
        if (t < R2) fftfast::stage1<R1, R2, -1>(v, t, S + l * P, tw);
        __syncthreads();
        if (t < R1) fftfast::stage2<R1, R2, -1>(v, t, S + l * P);
        __syncthreads();
        if (t < R1) {
            const int i = (t0 + l) % filt_w;
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                const int e = t + R1 * k2;
                const int j = e >= N / 2 ? e - N : e;
                const float s = filt[(int)(sqrtf((float)(i * i + j * j)) + 0.5f)];
                // KEEP: the plain forward transform goes to a second buffer, for the insertion (cspb_refine_keep_spectra)
                if (KEEP && live) keep[(long long)outer * ostride + t0 + l + (long long)e * estride] = v[k2];
                v[k2] = make_float2(__fmul_rn(v[k2].x, s), __fmul_rn(v[k2].y, s));  // rounded product, as the separate passes store it
                if (R1 != R2) S[l * P + fftsm::skew(e)] = v[k2];
            }
        }
        if (R1 != R2) {  // square factorisation: thread t already holds elements t + R2 r, the inverse transform's input
            __syncthreads();
            if (t < R2) {
#pragma unroll
                for (int r = 0; r < R1; ++r) v[r] = S[l * P + fftsm::skew(t + R2 * r)];
            }
            __syncthreads();
        }
        if (t < R2) fftfast::stage1<R1, R2, +1>(v, t, S + l * P, tw);
        __syncthreads();
        if (t < R1) {
            fftfast::stage2<R1, R2, +1>(v, t, S + l * P);
            if (live) {
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) base[(long long)(t + R1 * k2) * estride] = v[k2];
            }
        }
    }
}

// rows: half spectrum -> real (inverse) -> x scale x soft circular mask -> half spectrum (forward), in place
template <int R1, int R2>
__global__ void __launch_bounds__(256, R2 <= 16 ? 4 : 3) fft_rows_mask_fast_kernel(float2 *__restrict__ data, long long n_rows, const float2 *__restrict__ tw_g,
                                                                 float scale, float mask_radius, float mask_width) {
    constexpr int N = R1 * R2, PR = 256 / R2, P = N + (N >> 4) + 1, NH = N / 2 + 1, NT = PR * R2;
    __shared__ float2 S[PR * P];
    __shared__ float2 tw[N];
    const int tid = threadIdx.x;
    const int t = tid % R2, p = tid / R2;
    const long long pair = (long long)blockIdx.x * PR + p;
    const bool live = pair < n_rows / 2;
    for (int i = tid; i < N; i += NT) tw[i] = tw_g[i];
    for (int idx = tid; idx < PR * NH; idx += NT) {
        const int pp = idx / NH, k = idx - pp * NH;
        const long long pr = (long long)blockIdx.x * PR + pp;
        if (pr >= n_rows / 2) break;
        float2 fa = data[pr * 2 * NH + k], fb = data[(pr * 2 + 1) * NH + k];
        if (k == 0 || 2 * k == N) {
            fa.y = 0.f;
            fb.y = 0.f;
        }
        S[pp * P + fftsm::skew(k)] = make_float2(fa.x - fb.y, fa.y + fb.x);
        if (k > 0 && 2 * k < N) S[pp * P + fftsm::skew(N - k)] = make_float2(fa.x + fb.y, -fa.y + fb.x);
    }
    __syncthreads();
    float2 v[R1 > R2 ? R1 : R2];
#pragma unroll
    for (int r = 0; r < R1; ++r) v[r] = live ? S[p * P + fftsm::skew(t + R2 * r)] : make_float2(0.f, 0.f);
    __syncthreads();
    fftfast::stage1<R1, R2, +1>(v, t, S + p * P, tw);
    __syncthreads();
    if (t < R1) fftfast::stage2<R1, R2, +1>(v, t, S + p * P);
    __syncthreads();
    if (t < R1) {  // the two real rows of the pair sit in .x / .y: scale, mask, hand over to the forward transform
        const int ya = (int)((pair * 2) % N) - N / 2, yb = ya + 1;
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
            const int e = t + R1 * k2;
            const float x = (float)(e - N / 2);
            const float sa = scale * cosine_edge(sqrtf(x * x + (float)(ya * ya)), mask_radius, mask_width);
            const float sb = scale * cosine_edge(sqrtf(x * x + (float)(yb * yb)), mask_radius, mask_width);
            v[k2] = make_float2(__fmul_rn(v[k2].x, sa), __fmul_rn(v[k2].y, sb));
            if (R1 != R2) S[p * P + fftsm::skew(e)] = v[k2];
        }
    }
    if (R1 != R2) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = S[p * P + fftsm::skew(t + R2 * r)];
        __syncthreads();
    }
    fftfast::stage1<R1, R2, -1>(v, t, S + p * P, tw);
    __syncthreads();
    if (t < R1) fftfast::stage2<R1, R2, -1>(v, t, S + p * P);
    __syncthreads();
    if (t < R1) {
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) S[p * P + fftsm::skew(t + R1 * k2)] = v[k2];
    }
    __syncthreads();
    for (int idx = tid; idx < PR * NH; idx += NT) {
        const int pp = idx / NH, k = idx - pp * NH;
        const long long pr = (long long)blockIdx.x * PR + pp;
        if (pr >= n_rows / 2) break;
        const float2 za = S[pp * P + fftsm::skew(k)];
        float2 zb = S[pp * P + fftsm::skew(k ? N - k : 0)];
        zb.y = -zb.y;
        const float2 fa = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y + zb.y));
        const float2 df = make_float2(za.x - zb.x, za.y - zb.y);
        const float2 fb = make_float2(0.5f * df.y, -0.5f * df.x);
        data[pr * 2 * NH + k] = fa;
        data[(pr * 2 + 1) * NH + k] = fb;
    }
}

// columns: forward FFT, then the band samples go straight to their slots of the packed image (slot_of: n * nh table,
// -1 outside the band); sign of the centred origin and the optional per-ring weight applied on the way
template <int R1, int R2>
__global__ void __launch_bounds__(256) fft_cols_pack_fast_kernel(const float2 *__restrict__ data, long long estride, int ninner, long long ostride,
                                                                 int ntiles, const float2 *__restrict__ tw_g, const int32_t *__restrict__ slot_of,
                                                                 const float *__restrict__ ringw, float2 *__restrict__ packed, int n_slots) {
    constexpr int N = R1 * R2, TL = FastTile<R2>::TL, P = N + (N >> 4) + 1, NT = TL * (R2 < 16 ? 16 : R2);
    __shared__ float2 S[TL * P];
    __shared__ float2 tw[N];
    const int tid = threadIdx.x;
    const int l = tid % TL, t = tid / TL;
    const int outer = blockIdx.x / ntiles;
    const int t0 = (blockIdx.x - outer * ntiles) * TL;
    const bool live = t0 + l < ninner;
    const float2 *base = data + (long long)outer * ostride + t0 + l;
    for (int i = tid; i < N; i += NT) tw[i] = tw_g[i];
    float2 v[R1 > R2 ? R1 : R2];
    if (t < R2) {
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = live ? base[(long long)(t + R2 * r) * estride] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    if (t < R2) fftfast::stage1<R1, R2, -1>(v, t, S + l * P, tw);
    __syncthreads();
    if (t < R1) {
        fftfast::stage2<R1, R2, -1>(v, t, S + l * P);
        if (live) {
            const int i = t0 + l;
            float2 *o = packed + (long long)outer * n_slots;
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) {
                const int e = t + R1 * k2;
                const int slot = __ldg(slot_of + e * ninner + i);
                if (slot < 0) continue;
                const int j = e >= N / 2 ? e - N : e;
                float w = ((i + j) & 1) ? -1.f : 1.f;
                if (ringw) w *= ringw[(int)sqrtf((float)(i * i + j * j))];
                o[slot] = make_float2(v[k2].x * w, v[k2].y * w);
            }
        }
    }
}

__global__ void zero_slots_kernel(float2 *__restrict__ packed, int n_slots, const int32_t *__restrict__ list, int n_list, int count) {
    const long long total = (long long)n_list * count;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x)
        packed[(k / n_list) * n_slots + list[k % n_list]] = make_float2(0.f, 0.f);
}

int pick_tile(int n) {
    // keep 2 buffers of T lines within ~96 KB so two CTAs fit per SM
    int T = 16;
    while (T > 4 && (size_t)2 * T * line_pitch(n) * 8 > 96 * 1024) T /= 2;
    return T;
}

template <class K> int set_smem(cspb_ctx *ctx, K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        CU_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

int launch_lines(cspb_ctx *ctx, float2 *data, int n, long long estride, int ninner, long long ostride,
                 int n_outer, int dir, float scale, int sign_mode, int inner_w, const float *filt = nullptr, int filt_w = 1) {
    Radices rad;
    if (!factor(n, rad)) return cspb_fail(ctx, CSPB_E_ARG, "FFT length %d is not 2^a*3^b", n);
    const float2 *tw;
    int rc = fft_get_twiddles(ctx, n, &tw);
    if (rc) return rc;
    if (fft_has_fast_path(n)) {
        const int TL = n == 512 ? 8 : (n == 384 ? 10 : 16);
        const int nthreads = n == 384 ? 240 : 256;
        const int ntiles = ceil_div(ninner, TL);
        const unsigned grid = (unsigned)((long long)ntiles * n_outer);
#define CSPB_LINES_FAST(R1_, R2_)                                                                                       \
    do {                                                                                                                \
        if (dir < 0)                                                                                                    \
            fft_lines_fast_kernel<R1_, R2_, -1><<<grid, nthreads, 0, ctx->stream>>>(data, estride, ninner, ostride, ntiles, tw, scale, \
                                                                               sign_mode, inner_w, filt, filt_w);       \
        else                                                                                                            \
            fft_lines_fast_kernel<R1_, R2_, +1><<<grid, nthreads, 0, ctx->stream>>>(data, estride, ninner, ostride, ntiles, tw, scale, \
                                                                               sign_mode, inner_w, filt, filt_w);       \
    } while (0)
        if (n == 512) CSPB_LINES_FAST(16, 32);
        else if (n == 384) CSPB_LINES_FAST(16, 24);
        else if (n == 256) CSPB_LINES_FAST(16, 16);
        else if (n == 128) CSPB_LINES_FAST(8, 16);
        else CSPB_LINES_FAST(8, 8);
#undef CSPB_LINES_FAST
        KERNEL_CHECK(ctx);
        return 0;
    }
    if (filt) return cspb_fail(ctx, CSPB_E_ARG, "fused radial filter needs the fast FFT path (n = 64/128/256/384/512)");
    const int T = pick_tile(n);
    const size_t smem = ((size_t)n + (size_t)2 * T * line_pitch(n)) * sizeof(float2);
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
    const long long n_pairs = n_rows / 2;
    if (fft_has_fast_path(n)) {
        if (n == 512) fft_rows_r2c_fast_kernel<16, 32><<<ceil_div(n_pairs, 8), 256, 0, ctx->stream>>>(in, out, n_rows, tw, rows_per_image, offs, scls);
        else if (n == 384) fft_rows_r2c_fast_kernel<16, 24><<<ceil_div(n_pairs, 10), 240, 0, ctx->stream>>>(in, out, n_rows, tw, rows_per_image, offs, scls);
        else if (n == 256) fft_rows_r2c_fast_kernel<16, 16><<<ceil_div(n_pairs, 16), 256, 0, ctx->stream>>>(in, out, n_rows, tw, rows_per_image, offs, scls);
        else if (n == 128) fft_rows_r2c_fast_kernel<8, 16><<<ceil_div(n_pairs, 16), 256, 0, ctx->stream>>>(in, out, n_rows, tw, rows_per_image, offs, scls);
        else fft_rows_r2c_fast_kernel<8, 8><<<ceil_div(n_pairs, 32), 256, 0, ctx->stream>>>(in, out, n_rows, tw, rows_per_image, offs, scls);
        KERNEL_CHECK(ctx);
        return 0;
    }
    const int PR = pick_tile(n);
    const size_t smem = ((size_t)n + (size_t)2 * PR * line_pitch(n)) * sizeof(float2);
    if ((rc = set_smem(ctx, fft_rows_r2c_kernel, smem))) return rc;
    fft_rows_r2c_kernel<<<ceil_div(n_pairs, PR), 256, smem, ctx->stream>>>(in, out, n, n_rows, PR, rad, tw,
                                                                          rows_per_image, offs, scls);
    KERNEL_CHECK(ctx);
    return 0;
}

int launch_rows_c2r(cspb_ctx *ctx, const float2 *in, float *out, int n, long long n_rows, float scale,
                    float mask_radius = 0.f, float mask_width = 0.f) {
    Radices rad;
    if (!factor(n, rad) || (n & 1)) return cspb_fail(ctx, CSPB_E_ARG, "FFT length %d unsupported", n);
    const float2 *tw;
    int rc = fft_get_twiddles(ctx, n, &tw);
    if (rc) return rc;
    const long long n_pairs = n_rows / 2;
    if (fft_has_fast_path(n)) {
        if (n == 512) fft_rows_c2r_fast_kernel<16, 32><<<ceil_div(n_pairs, 8), 256, 0, ctx->stream>>>(in, out, n_rows, tw, scale, mask_radius, mask_width);
        else if (n == 384) fft_rows_c2r_fast_kernel<16, 24><<<ceil_div(n_pairs, 10), 240, 0, ctx->stream>>>(in, out, n_rows, tw, scale, mask_radius, mask_width);
        else if (n == 256) fft_rows_c2r_fast_kernel<16, 16><<<ceil_div(n_pairs, 16), 256, 0, ctx->stream>>>(in, out, n_rows, tw, scale, mask_radius, mask_width);
        else if (n == 128) fft_rows_c2r_fast_kernel<8, 16><<<ceil_div(n_pairs, 16), 256, 0, ctx->stream>>>(in, out, n_rows, tw, scale, mask_radius, mask_width);
        else fft_rows_c2r_fast_kernel<8, 8><<<ceil_div(n_pairs, 32), 256, 0, ctx->stream>>>(in, out, n_rows, tw, scale, mask_radius, mask_width);
        KERNEL_CHECK(ctx);
        return 0;
    }
    if (mask_width > 0.f) return cspb_fail(ctx, CSPB_E_ARG, "fused mask needs the fast FFT path (n = 64/128/256/384/512)");
    const int PR = pick_tile(n);
    const size_t smem = ((size_t)n + (size_t)2 * PR * line_pitch(n)) * sizeof(float2);
    if ((rc = set_smem(ctx, fft_rows_c2r_kernel, smem))) return rc;
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

bool fft_has_fast_path(int n) { return n == 64 || n == 128 || n == 256 || n == 384 || n == 512; }

int fft2_r2c_dev(cspb_ctx *ctx, const float *in, float2 *out, int n, int batch, const float *offs,
                 const float *scls, const float *radial_filter) {
    int rc = launch_rows_r2c(ctx, in, out, n, (long long)batch * n, n, offs, scls);
    if (rc) return rc;
    const int nh = n / 2 + 1;
    return launch_lines(ctx, out, n, nh, nh, (long long)n * nh, batch, -1, 1.f, 0, nh, radial_filter, nh);
}

int fft2_c2r_dev(cspb_ctx *ctx, float2 *inout_c, float *out, int n, int batch, float scale, float mask_radius,
                 float mask_width) {
    const int nh = n / 2 + 1;
    int rc = launch_lines(ctx, inout_c, n, nh, nh, (long long)n * nh, batch, +1, 1.f, 0, nh);
    if (rc) return rc;
    return launch_rows_c2r(ctx, inout_c, out, n, (long long)batch * n, scale, mask_radius, mask_width);
}

// The whole particle preprocessing of the fast path in four passes (refine.cu): rows r2c with normalisation | columns
// forward x whitening x inverse | rows inverse x 1/n^2 x mask x forward | columns forward + band pack.  `spec` = work buffer
// of batch half spectra; dummy slots of the band plan are zeroed from `dummy_list`.
int fft2_whiten_mask_pack_dev(cspb_ctx *ctx, const float *in, float2 *spec, int n, int batch, const float *offs, const float *scls,
                              const float *radial_filter, float scale, float mask_radius, float mask_width, const int32_t *slot_of,
                              const float *ringw, const int32_t *dummy_list, int n_dummy, float2 *packed, int n_slots, float2 *keep_forward) {
    if (!fft_has_fast_path(n)) return cspb_fail(ctx, CSPB_E_ARG, "fused preprocessing needs the fast FFT path");
    int rc = launch_rows_r2c(ctx, in, spec, n, (long long)batch * n, n, offs, scls);
    if (rc) return rc;
    const float2 *tw;
    if ((rc = fft_get_twiddles(ctx, n, &tw))) return rc;
    const int nh = n / 2 + 1;
    const int TL = n == 512 ? 8 : (n == 384 ? 10 : 16), nthreads = n == 384 ? 240 : 256;
    const int ntiles = ceil_div(nh, TL);
    const unsigned grid = (unsigned)((long long)ntiles * batch);
    const long long n_pairs = (long long)batch * n / 2;
    const long long ostride = (long long)n * nh;
    const unsigned pgrid = grid < (unsigned)(2 * ctx->sm_count) ? grid : (unsigned)(2 * ctx->sm_count);  // persistent: 2 CTAs per SM
#define CSPB_FUSED(R1_, R2_, PR_)                                                                                                        \
    do {                                                                                                                                 \
        if (keep_forward)                                                                                                                \
            fft_cols_filter_fast_kernel<R1_, R2_, true><<<pgrid, nthreads, 0, ctx->stream>>>(spec, nh, nh, ostride, ntiles, (int)grid, tw,    \
                                                                                             radial_filter, nh, keep_forward);          \
        else                                                                                                                             \
            fft_cols_filter_fast_kernel<R1_, R2_, false><<<pgrid, nthreads, 0, ctx->stream>>>(spec, nh, nh, ostride, ntiles, (int)grid, tw,   \
                                                                                              radial_filter, nh, nullptr);              \
        KERNEL_CHECK(ctx);                                                                                                               \
        fft_rows_mask_fast_kernel<R1_, R2_><<<ceil_div(n_pairs, PR_), nthreads, 0, ctx->stream>>>(spec, (long long)batch * n, tw, scale, \
                                                                                                 mask_radius, mask_width);              \
        KERNEL_CHECK(ctx);                                                                                                               \
        fft_cols_pack_fast_kernel<R1_, R2_><<<grid, nthreads, 0, ctx->stream>>>(spec, nh, nh, ostride, ntiles, tw, slot_of, ringw, packed, \
                                                                               n_slots);                                                \
        KERNEL_CHECK(ctx);                                                                                                               \
    } while (0)
    if (n == 512) CSPB_FUSED(16, 32, 8);
    else if (n == 384) CSPB_FUSED(16, 24, 10);
    else if (n == 256) CSPB_FUSED(16, 16, 16);
    else if (n == 128) CSPB_FUSED(8, 16, 16);
    else CSPB_FUSED(8, 8, 32);
#undef CSPB_FUSED
    if (n_dummy > 0) {
        zero_slots_kernel<<<ceil_div((long long)n_dummy * batch, 256) < 1024 ? ceil_div((long long)n_dummy * batch, 256) : 1024, 256, 0, ctx->stream>>>(
            packed, n_slots, dummy_list, n_dummy, batch);
        KERNEL_CHECK(ctx);
    }
    return 0;
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
