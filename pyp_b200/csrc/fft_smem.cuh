// fft_smem.cuh — shared-memory FFT building block: mixed-radix (16/8/4/2/3) Stockham autosort passes
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

// W_16^k = exp(DIR * 2 pi i k / 16), k = 0..3 (higher powers by symmetry)
template <int DIR> __device__ __forceinline__ float2 mul_w16(float2 v, int k) {
    const float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f, h = 0.70710678118654752440f;
    // (c + i DIR s)
    switch (k) {
    case 0: return v;
    case 1: return make_float2(v.x * c1 - DIR * v.y * s1, v.y * c1 + DIR * v.x * s1);
    case 2: return make_float2((v.x - DIR * v.y) * h, (v.y + DIR * v.x) * h);
    case 3: return make_float2(v.x * s1 - DIR * v.y * c1, v.y * s1 + DIR * v.x * c1);
    case 4: return make_float2(-DIR * v.y, DIR * v.x);
    case 6: return make_float2((-v.x - DIR * v.y) * h, (-v.y + DIR * v.x) * h);
    case 9: return make_float2(-v.x * c1 + DIR * v.y * s1, -v.y * c1 - DIR * v.x * s1);
    default: return v;
    }
}
// radix 8 = 2 x 4: n = a + 2m, k = b + 4c
template <int DIR> struct Butterfly<8, DIR> {
    __device__ __forceinline__ static void run(float2 *v) {
        float2 t[2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            float2 u[4] = {v[a], v[a + 2], v[a + 4], v[a + 6]};
            Butterfly<4, DIR>::run(u);
#pragma unroll
            for (int b = 0; b < 4; ++b) t[a][b] = a ? mul_w16<DIR>(u[b], 2 * b) : u[b];  // W_8^(ab) = W_16^(2ab)
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            v[b] = cadd(t[0][b], t[1][b]);
            v[b + 4] = csub(t[0][b], t[1][b]);
        }
    }
};
// radix 16 = 4 x 4: n = a + 4m, k = b + 4c
template <int DIR> struct Butterfly<16, DIR> {
    __device__ __forceinline__ static void run(float2 *v) {
        float2 t[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float2 u[4] = {v[a], v[a + 4], v[a + 8], v[a + 12]};
            Butterfly<4, DIR>::run(u);
#pragma unroll
            for (int b = 0; b < 4; ++b) t[a][b] = mul_w16<DIR>(u[b], a * b);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float2 u[4] = {t[0][b], t[1][b], t[2][b], t[3][b]};
            Butterfly<4, DIR>::run(u);
#pragma unroll
            for (int c = 0; c < 4; ++c) v[b + 4 * c] = u[c];
        }
    }
};

// cos / sin of 2 pi k / 96: compile-time twiddles for the radix-24 (stride 4) and radix-32 (stride 3) butterflies
__device__ constexpr float W96C[96] = {1.000000000e+00f, 9.978589232e-01f, 9.914448614e-01f, 9.807852804e-01f, 9.659258263e-01f, 9.469301295e-01f, 9.238795325e-01f, 8.968727415e-01f, 8.660254038e-01f, 8.314696123e-01f, 7.933533403e-01f, 7.518398075e-01f, 7.071067812e-01f, 6.593458151e-01f, 6.087614290e-01f, 5.555702330e-01f, 5.000000000e-01f, 4.422886902e-01f, 3.826834324e-01f, 3.214394653e-01f, 2.588190451e-01f, 1.950903220e-01f, 1.305261922e-01f, 6.540312923e-02f, 6.123233996e-17f, -6.540312923e-02f, -1.305261922e-01f, -1.950903220e-01f, -2.588190451e-01f, -3.214394653e-01f, -3.826834324e-01f, -4.422886902e-01f, -5.000000000e-01f, -5.555702330e-01f, -6.087614290e-01f, -6.593458151e-01f, -7.071067812e-01f, -7.518398075e-01f, -7.933533403e-01f, -8.314696123e-01f, -8.660254038e-01f, -8.968727415e-01f, -9.238795325e-01f, -9.469301295e-01f, -9.659258263e-01f, -9.807852804e-01f, -9.914448614e-01f, -9.978589232e-01f, -1.000000000e+00f, -9.978589232e-01f, -9.914448614e-01f, -9.807852804e-01f, -9.659258263e-01f, -9.469301295e-01f, -9.238795325e-01f, -8.968727415e-01f, -8.660254038e-01f, -8.314696123e-01f, -7.933533403e-01f, -7.518398075e-01f, -7.071067812e-01f, -6.593458151e-01f, -6.087614290e-01f, -5.555702330e-01f, -5.000000000e-01f, -4.422886902e-01f, -3.826834324e-01f, -3.214394653e-01f, -2.588190451e-01f, -1.950903220e-01f, -1.305261922e-01f, -6.540312923e-02f, -1.836970199e-16f, 6.540312923e-02f, 1.305261922e-01f, 1.950903220e-01f, 2.588190451e-01f, 3.214394653e-01f, 3.826834324e-01f, 4.422886902e-01f, 5.000000000e-01f, 5.555702330e-01f, 6.087614290e-01f, 6.593458151e-01f, 7.071067812e-01f, 7.518398075e-01f, 7.933533403e-01f, 8.314696123e-01f, 8.660254038e-01f, 8.968727415e-01f, 9.238795325e-01f, 9.469301295e-01f, 9.659258263e-01f, 9.807852804e-01f, 9.914448614e-01f, 9.978589232e-01f};
__device__ constexpr float W96S[96] = {0.000000000e+00f, 6.540312923e-02f, 1.305261922e-01f, 1.950903220e-01f, 2.588190451e-01f, 3.214394653e-01f, 3.826834324e-01f, 4.422886902e-01f, 5.000000000e-01f, 5.555702330e-01f, 6.087614290e-01f, 6.593458151e-01f, 7.071067812e-01f, 7.518398075e-01f, 7.933533403e-01f, 8.314696123e-01f, 8.660254038e-01f, 8.968727415e-01f, 9.238795325e-01f, 9.469301295e-01f, 9.659258263e-01f, 9.807852804e-01f, 9.914448614e-01f, 9.978589232e-01f, 1.000000000e+00f, 9.978589232e-01f, 9.914448614e-01f, 9.807852804e-01f, 9.659258263e-01f, 9.469301295e-01f, 9.238795325e-01f, 8.968727415e-01f, 8.660254038e-01f, 8.314696123e-01f, 7.933533403e-01f, 7.518398075e-01f, 7.071067812e-01f, 6.593458151e-01f, 6.087614290e-01f, 5.555702330e-01f, 5.000000000e-01f, 4.422886902e-01f, 3.826834324e-01f, 3.214394653e-01f, 2.588190451e-01f, 1.950903220e-01f, 1.305261922e-01f, 6.540312923e-02f, 1.224646799e-16f, -6.540312923e-02f, -1.305261922e-01f, -1.950903220e-01f, -2.588190451e-01f, -3.214394653e-01f, -3.826834324e-01f, -4.422886902e-01f, -5.000000000e-01f, -5.555702330e-01f, -6.087614290e-01f, -6.593458151e-01f, -7.071067812e-01f, -7.518398075e-01f, -7.933533403e-01f, -8.314696123e-01f, -8.660254038e-01f, -8.968727415e-01f, -9.238795325e-01f, -9.469301295e-01f, -9.659258263e-01f, -9.807852804e-01f, -9.914448614e-01f, -9.978589232e-01f, -1.000000000e+00f, -9.978589232e-01f, -9.914448614e-01f, -9.807852804e-01f, -9.659258263e-01f, -9.469301295e-01f, -9.238795325e-01f, -8.968727415e-01f, -8.660254038e-01f, -8.314696123e-01f, -7.933533403e-01f, -7.518398075e-01f, -7.071067812e-01f, -6.593458151e-01f, -6.087614290e-01f, -5.555702330e-01f, -5.000000000e-01f, -4.422886902e-01f, -3.826834324e-01f, -3.214394653e-01f, -2.588190451e-01f, -1.950903220e-01f, -1.305261922e-01f, -6.540312923e-02f};
// v * exp(DIR * 2 pi i k / 96), k a compile-time constant after unrolling
template <int DIR> __device__ __forceinline__ float2 mul_w96(float2 v, int k) {
    const float c = W96C[k % 96], s = DIR * W96S[k % 96];
    return make_float2(v.x * c - v.y * s, v.y * c + v.x * s);
}
// radix 24 = 3 x 8: n = a + 3m, k = b + 8c
template <int DIR> struct Butterfly<24, DIR> {
    __device__ __forceinline__ static void run(float2 *v) {
        float2 t[3][8];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float2 u[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) u[m] = v[a + 3 * m];
            Butterfly<8, DIR>::run(u);
#pragma unroll
            for (int b = 0; b < 8; ++b) t[a][b] = (a * b) ? mul_w96<DIR>(u[b], 4 * a * b) : u[b];
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            float2 u[3] = {t[0][b], t[1][b], t[2][b]};
            Butterfly<3, DIR>::run(u);
#pragma unroll
            for (int c = 0; c < 3; ++c) v[b + 8 * c] = u[c];
        }
    }
};
// radix 32 = 4 x 8: n = a + 4m, k = b + 8c
template <int DIR> struct Butterfly<32, DIR> {
    __device__ __forceinline__ static void run(float2 *v) {
        float2 t[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float2 u[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) u[m] = v[a + 4 * m];
            Butterfly<8, DIR>::run(u);
#pragma unroll
            for (int b = 0; b < 8; ++b) t[a][b] = (a * b) ? mul_w96<DIR>(u[b], 3 * a * b) : u[b];
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            float2 u[4] = {t[0][b], t[1][b], t[2][b], t[3][b]};
            Butterfly<4, DIR>::run(u);
#pragma unroll
            for (int c = 0; c < 4; ++c) v[b + 8 * c] = u[c];
        }
    }
};

// skewed shared-memory index: one padding element per 16 keeps the stride-R stores of a radix-8/16
// Stockham pass and the unit-stride loads of the next one free of bank conflicts
__host__ __device__ __forceinline__ int skew(int a) { return a + (a >> 4); }
// elements a line of n complex values occupies in shared memory (skewed, +1 against pitch conflicts)
__host__ __device__ __forceinline__ int line_pitch(int n) { return n + (n >> 4) + 1; }

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
        for (int r = 0; r < R; ++r) v[r] = s[skew(j + r * per_line)];
#pragma unroll
        for (int r = 1; r < R; ++r) {
            float2 t = tw[r * k * tws];
            if (DIR > 0) t.y = -t.y;
            v[r] = cmul(v[r], t);
        }
        Butterfly<R, DIR>::run(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) d[skew(j0 + r * Ns)] = v[r];
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
        if (R == 16)
            stockham_pass<16, DIR>(src, dst, pitch, nlines, n, Ns, tw, tid, nthreads);
        else if (R == 8)
            stockham_pass<8, DIR>(src, dst, pitch, nlines, n, Ns, tw, tid, nthreads);
        else if (R == 4)
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
    while (m % 16 == 0) { rad.r[rad.count++] = 16; m /= 16; }
    while (m % 8 == 0) { rad.r[rad.count++] = 8; m /= 8; }
    while (m % 4 == 0) { rad.r[rad.count++] = 4; m /= 4; }
    while (m % 2 == 0) { rad.r[rad.count++] = 2; m /= 2; }
    while (m % 3 == 0) { rad.r[rad.count++] = 3; m /= 3; }
    return m == 1 && rad.count <= 12;
}

}  // namespace fftsm
