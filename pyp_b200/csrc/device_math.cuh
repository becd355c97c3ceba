// device_math.cuh — small inline math shared by the kernels.  Definitions follow
// oracle/SEMANTICS.md (CTF, Euler convention, trilinear gather); the Euler convention is the
// FREALIGN/cisTEM ZYZ matrix that pyp's Python side decodes at
// src/pyp/analysis/geometry/core.py:174-234.
#pragma once
#include <cuda_runtime.h>
#include "internal.cuh"

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// block-wide sum, result valid in every thread; red needs >= 33 floats
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// 1 inside radius-width/2, 0 outside radius+width/2, raised cosine between
__host__ __device__ __forceinline__ float cosine_edge(float r, float radius, float width) {
    const float lo = radius - 0.5f * width, hi = radius + 0.5f * width;
    if (r <= lo) return 1.f;
    if (r >= hi) return 0.f;
    return 0.5f * (1.f + cospif((r - lo) / width));
}

// sinc^2(pi d / np): real-space footprint of trilinear interpolation on an np-point Fourier grid
__host__ __device__ __forceinline__ float sinc2_corr(int d, int np) {
    if (d == 0) return 1.f;
    const float a = CSPB_PI_F * (float)d / (float)np;
    const float s = sinf(a) / a;
    return s * s;
}

// R = Rz(phi) Ry(theta) Rz(psi), row-major r[9]; a 2-D frequency (i, j) maps to
// (r0 i + r1 j, r3 i + r4 j, r6 i + r7 j).  Angles in degrees.
__host__ __device__ __forceinline__ void euler_matrix(float psi, float theta, float phi, float *r) {
    float sps, cps, sth, cth, sph, cph;
    sincosf(psi * (CSPB_PI_F / 180.f), &sps, &cps);
    sincosf(theta * (CSPB_PI_F / 180.f), &sth, &cth);
    sincosf(phi * (CSPB_PI_F / 180.f), &sph, &cph);
    r[0] = cph * cth * cps - sph * sps;
    r[1] = -cph * cth * sps - sph * cps;
    r[2] = cph * sth;
    r[3] = sph * cth * cps + cph * sps;
    r[4] = -sph * cth * sps + cph * cps;
    r[5] = sph * sth;
    r[6] = -sth * cps;
    r[7] = sth * sps;
    r[8] = cth;
}

// CTF coefficients with the pixel size folded in (see CtfCoef in internal.cuh)
__host__ __device__ __forceinline__ CtfCoef make_ctf_coef(float d1, float d2, float ast_deg, float phase_shift,
                                                          float pixel, float kv, float cs_mm, float ampl, int box) {
    CtfCoef c;
    const float v = kv * 1000.f;
    const float lambda = 12.2639f / sqrtf(v + 0.97845e-6f * v * v);
    const float s2u = 1.f / (((float)box * pixel) * ((float)box * pixel));
    c.a = CSPB_PI_F * lambda * 0.5f * (d1 + d2) * s2u;
    c.b = CSPB_PI_F * lambda * 0.5f * (d1 - d2) * s2u;
    float s2, c2;
    sincosf(2.f * ast_deg * (CSPB_PI_F / 180.f), &s2, &c2);
    c.cos2ast = c2;
    c.sin2ast = s2;
    c.c4 = -0.5f * CSPB_PI_F * lambda * lambda * lambda * (cs_mm * 1.0e7f) * s2u * s2u;
    c.ph0 = phase_shift + atanf(ampl / sqrtf(1.f - ampl * ampl));
    c.dstep = CSPB_PI_F * lambda * s2u;
    c.pad_ = 0.f;
    return c;
}

// phase chi at frequency index (fi, fj), r2 = fi^2 + fj^2
__device__ __forceinline__ float ctf_chi(const CtfCoef &c, float fi, float fj, float r2) {
    // r2 cos 2(alpha) = fi^2 - fj^2 and r2 sin 2(alpha) = 2 fi fj: no division by r2
    return r2 * c.a + c.b * ((fi * fi - fj * fj) * c.cos2ast + 2.f * fi * fj * c.sin2ast) + c.c4 * r2 * r2 + c.ph0;
}

// trilinear gather from the cropped centred x-paired reference (Friedel flip for x < 0)
__device__ __forceinline__ float2 gather_trilinear(const float4 *__restrict__ ref4, int sx, int sy, int rc, float x,
                                                   float y, float z) {
    const bool flip = x < 0.f;
    if (flip) { x = -x; y = -y; z = -z; }
    const float x0 = floorf(x), y0 = floorf(y), z0 = floorf(z);
    const float fx = x - x0, fy = y - y0, fz = z - z0;
    const long long sxy = (long long)sx * sy;
    const long long idx = ((long long)((int)z0 + rc) * sy + ((int)y0 + rc)) * sx + (int)x0;
    const float4 a00 = __ldg(ref4 + idx), a10 = __ldg(ref4 + idx + sx);
    const float4 a01 = __ldg(ref4 + idx + sxy), a11 = __ldg(ref4 + idx + sxy + sx);
    const float v00x = a00.x + fx * (a00.z - a00.x), v00y = a00.y + fx * (a00.w - a00.y);
    const float v10x = a10.x + fx * (a10.z - a10.x), v10y = a10.y + fx * (a10.w - a10.y);
    const float v01x = a01.x + fx * (a01.z - a01.x), v01y = a01.y + fx * (a01.w - a01.y);
    const float v11x = a11.x + fx * (a11.z - a11.x), v11y = a11.y + fx * (a11.w - a11.y);
    const float v0x = v00x + fy * (v10x - v00x), v0y = v00y + fy * (v10y - v00y);
    const float v1x = v01x + fy * (v11x - v01x), v1y = v01y + fy * (v11y - v01y);
    float2 p = make_float2(v0x + fz * (v1x - v0x), v0y + fz * (v1y - v0y));
    if (flip) p.y = -p.y;
    return p;
}

// sigma / logP from the band sums {num, signed X, A, B} (oracle/SEMANTICS.md §score)
__host__ __device__ __forceinline__ void score_stats(float4 v, int n_samples, float *sigma, float *logp) {
    const float X = v.y, A = v.z, B = v.w;
    const float alpha = B > 0.f ? X / B : 0.f;
    float resid = A - 2.f * alpha * X + alpha * alpha * B;
    if (resid < 0.f) resid = 0.f;
    const float sig2 = alpha * alpha * B;
    *sigma = sig2 > 0.f ? sqrtf(resid / sig2) : 0.f;
    const float ns = (float)(n_samples > 0 ? n_samples : 1);
    const float var = resid / ns;
    *logp = var > 0.f ? -0.5f * ns * (1.f + logf(2.f * CSPB_PI_F * var)) : 0.f;
}

// mean and 1/sigma of the pixels outside radius R (clamped to the half box, as the reference does) of one
// n x n image per CTA — the normalisation of analysis/image.py:320-338,406-417.  Rows are walked by
// warps with 16-byte loads (no integer division); the second pass re-reads the image from L2.
// DUAL: the same two passes also give the statistics for a second radius (the reconstruction's, when the forward
// transforms are kept for it: cspb_refine_keep_spectra) — own accumulators in the same order, so each pair of results is
// bit-identical to a single-radius call.
template <bool DUAL>
__device__ __forceinline__ void image_edge_stats_t(const float *__restrict__ p, int n, float radius, float radius2, int invert, int invert2,
                                                   float *off_out, float *scl_out, float *off2_out, float *scl2_out,
                                                   float *red /* >= 64 floats */) {
    __shared__ float s_mean_[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    // a radius beyond the half box is clamped to it (analysis/image.py:324-331): the corners are the background
    if (radius > 0.5f * (float)n) radius = 0.5f * (float)n;
    if (radius2 > 0.5f * (float)n) radius2 = 0.5f * (float)n;
    const float r2lim = radius * radius, r2lim2 = radius2 * radius2;
    const float r2skip = DUAL ? fminf(r2lim, r2lim2) : r2lim;  // groups wholly inside every radius are never loaded
    const int c = n / 2;
    const bool vec = (n & 3) == 0;
    float s = 0.f, cnt = 0.f, s2 = 0.f, cnt2 = 0.f;
    for (int y = warp; y < n; y += nw) {
        const float dy2 = (float)((y - c) * (y - c));
        const float *row = p + (long long)y * n;
        if (vec) {
            for (int x = lane * 4; x < n; x += 128) {
                // four pixels wholly inside the radius are not loaded at all (dx^2 is convex: the end pixels decide)
                const float dxa = (float)(x - c), dxb = (float)(x + 3 - c);
                if (fmaxf(dxa * dxa, dxb * dxb) + dy2 <= r2skip) continue;
                const float4 v = *reinterpret_cast<const float4 *>(row + x);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float dx = (float)(x + k - c);
                    if (dx * dx + dy2 > r2lim) { s += vv[k]; cnt += 1.f; }
                    if (DUAL && dx * dx + dy2 > r2lim2) { s2 += vv[k]; cnt2 += 1.f; }
                }
            }
        } else {
            for (int x = lane; x < n; x += 32) {
                const float dx = (float)(x - c);
                if (dx * dx + dy2 > r2lim) { s += row[x]; cnt += 1.f; }
                if (DUAL && dx * dx + dy2 > r2lim2) { s2 += row[x]; cnt2 += 1.f; }
            }
        }
    }
    s = block_sum(s, red);
    cnt = block_sum(cnt, red);
    if (DUAL) {
        s2 = block_sum(s2, red);
        cnt2 = block_sum(cnt2, red);
    }
    if (tid == 0) {
        s_mean_[0] = cnt > 0.f ? s / cnt : 0.f;
        s_mean_[1] = cnt2 > 0.f ? s2 / cnt2 : 0.f;
    }
    __syncthreads();
    const float mean = s_mean_[0], mean2 = s_mean_[1];
    float v2 = 0.f, w2 = 0.f;
    for (int y = warp; y < n; y += nw) {
        const float dy2 = (float)((y - c) * (y - c));
        const float *row = p + (long long)y * n;
        if (vec) {
            for (int x = lane * 4; x < n; x += 128) {
                const float dxa = (float)(x - c), dxb = (float)(x + 3 - c);
                if (fmaxf(dxa * dxa, dxb * dxb) + dy2 <= r2skip) continue;
                const float4 v = *reinterpret_cast<const float4 *>(row + x);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float dx = (float)(x + k - c);
                    if (dx * dx + dy2 > r2lim) { const float d = vv[k] - mean; v2 += d * d; }
                    if (DUAL && dx * dx + dy2 > r2lim2) { const float d = vv[k] - mean2; w2 += d * d; }
                }
            }
        } else {
            for (int x = lane; x < n; x += 32) {
                const float dx = (float)(x - c);
                if (dx * dx + dy2 > r2lim) { const float d = row[x] - mean; v2 += d * d; }
                if (DUAL && dx * dx + dy2 > r2lim2) { const float d = row[x] - mean2; w2 += d * d; }
            }
        }
    }
    v2 = block_sum(v2, red);
    if (DUAL) w2 = block_sum(w2, red);
    if (tid == 0) {
        const float sgn = invert ? -1.f : 1.f;
        const float var = cnt > 0.f ? v2 / cnt : 0.f;
        *off_out = mean;
        *scl_out = var > 0.f ? sgn * rsqrtf(var) : sgn;
        if (DUAL) {
            const float sgn2 = invert2 ? -1.f : 1.f;
            const float var2 = cnt2 > 0.f ? w2 / cnt2 : 0.f;
            *off2_out = mean2;
            *scl2_out = var2 > 0.f ? sgn2 * rsqrtf(var2) : sgn2;
        }
    }
}

__device__ __forceinline__ void image_edge_stats(const float *__restrict__ p, int n, float radius, int normalize, int invert,
                                                 float *off_out, float *scl_out, float *red /* >= 64 floats */) {
    if (!normalize) {
        if (threadIdx.x == 0) { *off_out = 0.f; *scl_out = invert ? -1.f : 1.f; }
        return;
    }
    image_edge_stats_t<false>(p, n, radius, radius, invert, invert, off_out, scl_out, nullptr, nullptr, red);
}
