// fft_fast.cuh — two-stage register FFT for n = R1 * R2 (256 = 16 x 16, 128 = 8 x 16, 64 = 8 x 8):
// every thread loads R1 elements straight from global memory into registers, runs a radix-R1
// butterfly, multiplies by W_n^(t k1), exchanges ONCE through shared memory (skewed, conflict free)
// and finishes with a radix-R2 butterfly whose results go straight back to global memory.
// One shared buffer, one barrier, R1 independent global loads in flight per thread — against the
// generic Stockham path (fft_smem.cuh) that ping-pongs between two buffers once per radix.
#pragma once
#include "fft_smem.cuh"

namespace fftfast {
using namespace fftsm;

// stage 1 of line FFT: v[r] = x[t + (n/R1) r]  ->  S[k1 * R2 + t] = W_n^(DIR t k1) * sum_r v[r] W_R1^(DIR r k1)
template <int R1, int R2, int DIR>
__device__ __forceinline__ void stage1(float2 *v, int t, float2 *S, const float2 *__restrict__ tw) {
    Butterfly<R1, DIR>::run(v);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
        float2 w = v[k1];
        if (k1 > 0) {
            float2 c = tw[t * k1];
            if (DIR > 0) c.y = -c.y;
            w = cmul(w, c);
        }
        S[skew(k1 * R2 + t)] = w;
    }
}
// stage 2: thread u (< R1) gathers S[u * R2 + t'], t' < R2, and leaves X[u + R1 k2] in w[k2]
template <int R1, int R2, int DIR>
__device__ __forceinline__ void stage2(float2 *w, int u, const float2 *S) {
#pragma unroll
    for (int tp = 0; tp < R2; ++tp) w[tp] = S[skew(u * R2 + tp)];
    Butterfly<R2, DIR>::run(w);
}

}  // namespace fftfast
