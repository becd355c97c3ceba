// fft_smem.cuh — shared-memory FFT building block: mixed-radix (4/2/3) Stockham autosort passes
// over lines held in shared memory, ping-ponging between two buffers; twiddles W_n^k from a table.
#pragma once
#include <cuda_runtime.h>

namespace fftsm {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// DIR = -1 forward (exp(-i..)), +1 inverse
template <int R, int DIR> struct Butterfly;
template <int DIR> struct Butterfly<2, DIR> {
    __device__ __forceinline__ static void run(float2 *v) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <int DIR> struct Butterfly<4, DIR> {
    __device__ __forceinline__ static void run(float2 *v) {
        float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        float2 t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        float2 t3 = (DIR < 0) ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};
template <int DIR> struct Butterfly<3, DIR> {
    __device__ __forceinline__ static void run(float2 *v) {
        const float c = 0.86602540378443864676f;
        float2 s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
        float2 m = make_float2(v[0].x - 0.5f * s.x, v[0].y - 0.5f * s.y);
        float2 q = (DIR < 0) ? make_float2(d.y * c, -d.x * c) : make_float2(-d.y * c, d.x * c);
        v[0] = cadd(v[0], s);
        v[1] = cadd(m, q);
        v[2] = csub(m, q);
    }
};

template <int R, int DIR>
__device__ __forceinline__ void stockham_pass(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                              int pitch, int nlines, int n, int Ns,
                                              const float2 *__restrict__ tw, int tid, int nthreads) {
    const int per_line = n / R;
    const int total = per_line * nlines;
    const int tws = n / (Ns * R);
    for (int w = tid; w < total; w += nthreads) {
        const int line = w / per_line;
        const int j = w - line * per_line;
        const float2 *s = src + line * pitch;
        float2 *d = dst + line * pitch;
        const int k = j % Ns;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = s[j + r * per_line];
#pragma unroll
        for (int r = 1; r < R; ++r) {
            float2 t = tw[r * k * tws];
            if (DIR > 0) t.y = -t.y;
            v[r] = cmul(v[r], t);
        }
        Butterfly<R, DIR>::run(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) d[j0 + r * Ns] = v[r];
    }
}

struct Radices {
    int count;
    int r[12];
};

// run all passes; returns pointer to the buffer holding the result
template <int DIR>
__device__ __forceinline__ float2 *fft_lines_smem(float2 *a, float2 *b, int pitch, int nlines, int n,
                                                  const Radices &rad, const float2 *tw, int tid,
                                                  int nthreads) {
    int Ns = 1;
    float2 *src = a, *dst = b;
    for (int p = 0; p < rad.count; ++p) {
        const int R = rad.r[p];
        if (R == 4)
            stockham_pass<4, DIR>(src, dst, pitch, nlines, n, Ns, tw, tid, nthreads);
        else if (R == 2)
            stockham_pass<2, DIR>(src, dst, pitch, nlines, n, Ns, tw, tid, nthreads);
        else
            stockham_pass<3, DIR>(src, dst, pitch, nlines, n, Ns, tw, tid, nthreads);
        Ns *= R;
        __syncthreads();
        float2 *t = src;
        src = dst;
        dst = t;
    }
    return src;
}


inline bool factor(int n, Radices &rad) {
    rad.count = 0;
    int m = n;
    while (m % 4 == 0) { rad.r[rad.count++] = 4; m /= 4; }
    while (m % 2 == 0) { rad.r[rad.count++] = 2; m /= 2; }
    while (m % 3 == 0) { rad.r[rad.count++] = 3; m /= 3; }
    return m == 1 && rad.count <= 12;
}

}  // namespace fftsm
