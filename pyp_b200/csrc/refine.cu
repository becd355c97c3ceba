// refine.cu — the projection-matching scorer: reference preparation, particle preprocessing,
// the fused central-slice x CTF x weighted-correlation kernel and the batched local optimiser.
//
// Replaces the numerics of external/cistem2/refine3d (closed LFS binary); the contract is the
// stdin answer list built at src/pyp/refine/frealign/frealign.py:3918-3994 and the `.cistem`
// rows of src/pyp/inout/metadata/cistem_star_file.py:596-628.  Semantics that the reference
// tree cannot pin (score definition, whitening, mask) are fixed in oracle/SEMANTICS.md and
// restated on the CPU in oracle/cspb_oracle.c, against which these kernels are tested.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "device_math.cuh"
#include "internal.cuh"
#include "opt.cuh"

// ================================================================== reference preparation
namespace {

// real n^3 -> padded np^3 (centred, zero padded)
__global__ void pad_volume_kernel(const float *__restrict__ vol, float *__restrict__ out, int n, int np) {
    const long long total = (long long)np * np * np;
    const int off = (np - n) / 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % np);
        const int y = (int)((idx / np) % np);
        const int z = (int)(idx / ((long long)np * np));
        const int sx = x - off, sy = y - off, sz = z - off;
        float v = 0.f;
        if (sx >= 0 && sx < n && sy >= 0 && sy < n && sz >= 0 && sz < n) {
            // pre-compensate the trilinear interpolation of the slice gather (sinc^2 per axis)
            const float g = sinc2_corr(x - np / 2, np) * sinc2_corr(y - np / 2, np) * sinc2_corr(z - np / 2, np);
            v = vol[((long long)sz * n + sy) * n + sx] / g;
        }
        out[idx] = v;
    }
}

// FFT-ordered half spectrum [z][y][xh] -> cropped centred x-paired float4 volume
__global__ void crop_pair_kernel(const float2 *__restrict__ spec, float4 *__restrict__ ref4, int np, int rc,
                                 int sx, int sy) {
    const long long total = (long long)sx * sy * sy;
    const int xh = np / 2 + 1;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % sx);
        const int yy = (int)((idx / sx) % sy);
        const int zz = (int)(idx / ((long long)sx * sy));
        const int y = yy - rc, z = zz - rc;
        float2 v[2];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int xx = x + d;
            float2 t = make_float2(0.f, 0.f);
            if (xx < xh && y >= -np / 2 && y < np / 2 && z >= -np / 2 && z < np / 2) {
                const int iy = y < 0 ? y + np : y, iz = z < 0 ? z + np : z;
                t = spec[((long long)iz * np + iy) * xh + xx];
                if ((xx + y + z) & 1) {  // real-space centre at np/2 -> (-1)^(x+y+z)
                    t.x = -t.x;
                    t.y = -t.y;
                }
            }
            v[d] = t;
        }
        ref4[idx] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    }
}

// FFT-ordered half spectrum -> cropped centred volume of (y,z) 2x2 quads for the scorer:
// ref8[(z*sy + y)*sx8 + x] = { V(x,y,z), V(x,y+1,z), V(x,y,z+1), V(x,y+1,z+1) }  (32 bytes, one LDG.256)
__global__ void crop_quad_kernel(const float2 *__restrict__ spec, RefQuad *__restrict__ ref8, int np, int rc, int sx8, int sy) {
    const long long total = (long long)sx8 * sy * sy;
    const int xh = np / 2 + 1;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % sx8);
        const int yy = (int)((idx / sx8) % sy);
        const int zz = (int)(idx / ((long long)sx8 * sy));
        float2 v[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int y = yy - rc + (d & 1), z = zz - rc + (d >> 1);
            float2 t = make_float2(0.f, 0.f);
            if (x < xh && y >= -np / 2 && y < np / 2 && z >= -np / 2 && z < np / 2) {
                const int iy = y < 0 ? y + np : y, iz = z < 0 ? z + np : z;
                t = spec[((long long)iz * np + iy) * xh + x];
                if ((x + y + z) & 1) {  // real-space centre at np/2 -> (-1)^(x+y+z)
                    t.x = -t.x;
                    t.y = -t.y;
                }
            }
            v[d] = t;
        }
        RefQuad q;
        q.v00 = v[0]; q.v10 = v[1]; q.v01 = v[2]; q.v11 = v[3];
        ref8[idx] = q;
    }
}

// ================================================================== particle preprocessing
// per-image mean / variance of the pixels outside radius R (all pixels if none are outside)
__global__ void image_stats_kernel(const float *__restrict__ img, int n, float radius, int normalize,
                                   int invert, float *__restrict__ offs, float *__restrict__ scls) {
    __shared__ float red[64];
    image_edge_stats(img + (long long)blockIdx.x * n * n, n, radius, normalize, invert, offs + blockIdx.x, scls + blockIdx.x, red);
}

// the same, plus the statistics for a second radius from the same two passes (kept spectra: the reconstruction's radius)
__global__ void image_stats_dual_kernel(const float *__restrict__ img, int n, float radius, float radius2, int invert, int invert2,
                                        float *__restrict__ offs, float *__restrict__ scls, float *__restrict__ offs2,
                                        float *__restrict__ scls2) {
    __shared__ float red[64];
    image_edge_stats_t<true>(img + (long long)blockIdx.x * n * n, n, radius, radius2, invert, invert2, offs + blockIdx.x, scls + blockIdx.x,
                             offs2 + blockIdx.x, scls2 + blockIdx.x, red);
}

// per-image ring power: one warp per ring walks a host-built CSR list of the half-plane samples
// (deterministic order).  ring = nearest integer radius.  out[img*n_rings + ring] = sum |F|^2
__global__ void ring_power_kernel(const float2 *__restrict__ spec, int n, int img_step, const int *__restrict__ ring_off,
                                  const int *__restrict__ ring_idx, int n_rings, float *__restrict__ out) {
    const int nh = n / 2 + 1;
    const float2 *f = spec + (long long)blockIdx.x * img_step * n * nh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = warp; r < n_rings; r += nw) {
        float s = 0.f;
        for (int k = ring_off[r] + lane; k < ring_off[r + 1]; k += 32) {
            const float2 v = f[ring_idx[k]];
            s += v.x * v.x + v.y * v.y;
        }
        s = warp_sum(s);
        if (lane == 0) out[(long long)blockIdx.x * n_rings + r] = s;
    }
}

__global__ void sum_over_images_kernel(const float *__restrict__ in, int n_img, int n_rings, float *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rings) return;
    float s = 0.f;
    for (int i = 0; i < n_img; ++i) s += in[(long long)i * n_rings + r];
    out[r] = s;
}

// multiply a half spectrum by a radial filter indexed by nearest ring
__global__ void radial_filter_kernel(float2 *__restrict__ spec, int n, int batch, const float *__restrict__ filt) {
    const int nh = n / 2 + 1;
    const long long total = (long long)batch * n * nh;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % nh);
        int j = (int)((idx / nh) % n);
        if (j >= n / 2) j -= n;
        const int ring = (int)(sqrtf((float)(i * i + j * j)) + 0.5f);
        const float w = filt[ring];
        float2 v = spec[idx];
        v.x *= w;
        v.y *= w;
        spec[idx] = v;
    }
}

// real-space soft circular mask (cosine edge of width `width` centred on `radius`) and 1/n^2
__global__ void mask_kernel(float *__restrict__ img, int n, int batch, float radius, float width, float scale) {
    const long long total = (long long)batch * n * n;
    const int c = n / 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % n) - c, y = (int)((idx / n) % n) - c;
        const float r = sqrtf((float)(x * x + y * y));
        img[idx] *= scale * cosine_edge(r, radius, width);
    }
}

// band-pack: packed[img][slot] = sign(i+j) * F[img][j][i] * filt[nearest ring] * ringw[ring]
__global__ void pack_kernel(const float2 *__restrict__ spec, int n, int n_slots, const int32_t *__restrict__ slot_ij,
                            const float *__restrict__ filt, const float *__restrict__ ringw,
                            float2 *__restrict__ packed) {
    const int nh = n / 2 + 1;
    const float2 *f = spec + (long long)blockIdx.y * n * nh;
    float2 *o = packed + (long long)blockIdx.y * n_slots;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
        const int32_t ij = slot_ij[s];
        const int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
        float2 v = make_float2(0.f, 0.f);
        if (i != CSPB_DUMMY_I) {
            const int jj = j < 0 ? j + n : j;
            v = f[jj * nh + i];
            const float r = sqrtf((float)(i * i + j * j));
            float w = ((i + j) & 1) ? -1.f : 1.f;
            if (filt) w *= filt[(int)(r + 0.5f)];
            if (ringw) w *= ringw[(int)r];
            v.x *= w;
            v.y *= w;
        }
        o[s] = v;
    }
}

// ================================================================== CTF coefficients
__global__ void ctf_coef_kernel(const cspb_row *__restrict__ rows, int n_rows, int box, CtfCoef *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_rows) return;
    out[k] = make_ctf_coef(rows[k].defocus_1, rows[k].defocus_2, rows[k].defocus_angle, rows[k].phase_shift,
                           rows[k].pixel_size, rows[k].voltage_kv, rows[k].cs_mm, rows[k].amplitude_contrast, box);
}

// ================================================================== the scorer
struct ScoreArgs {
    const RefQuad *ref8;
    int sx, sy, rc;   // sx = x extent of the quad volume (rc + 2)
    float padf;
    const int32_t *slot_ij;
    const BandDesc *bands;
    int n_bands;
    int n_slots;
    const float2 *packed;
    const CtfCoef *ctf;
    const ScoreUnit *units;
    int n_units;
    const float *poses6;  // per eval: psi, theta, phi (deg), shift x, y (Angstrom), defocus delta (Angstrom)
    float inv_npx2;       // 2 / (n * pixel): shift (A) * frequency index -> phase in units of pi
    int limit_ring;       // rings above use |X_r| (signed-CC limit); INT_MAX = always signed
    int ring_cut;         // coarse-to-fine stages (CUT kernels): rings above do not enter the sums; INT_MAX = whole band
    float4 *out;          // per eval: {numerator, signed X total, A, B}
    float *gout;          // score_grad_kernel: 28 floats per eval {num, X, A, B, dnum[5], dB[3], jtj[15], 0}
};

// one 256-bit read-only load (LDG.E.ENL2.256.CONSTANT, sm_100+)
__device__ __forceinline__ RefQuad ldg_quad(const RefQuad *p) {
    RefQuad q;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(q.v00.x), "=f"(q.v00.y), "=f"(q.v10.x), "=f"(q.v10.y), "=f"(q.v01.x), "=f"(q.v01.y), "=f"(q.v11.x), "=f"(q.v11.y)
        : "l"(p));
    return q;
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FFMA2 / FMUL2 on an aligned register pair).  A complex
// value lives in one 64-bit register; a trilinear interpolation is 7 FADD2 + 7 FFMA2 instead of 28 scalar
// instructions, with the same roundings (a + f (b - a) per component).
typedef unsigned long long f32x2;
struct QuadP { f32x2 v00, v10, v01, v11; };  // register image of a RefQuad
__device__ __forceinline__ QuadP ldg_quadp(const RefQuad *p) {
    QuadP q;
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q.v00), "=l"(q.v10), "=l"(q.v01), "=l"(q.v11) : "l"(p));
    return q;
}
// predicated: lanes with `take` false issue no memory traffic and keep `q`
__device__ __forceinline__ void ldg_quadp_if(const RefQuad *p, bool take, QuadP &q) {
    asm("{\n\t.reg .pred t;\n\tsetp.ne.s32 t, %5, 0;\n\t@t ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];\n\t}"
        : "+l"(q.v00), "+l"(q.v10), "+l"(q.v01), "+l"(q.v11)
        : "l"(p), "r"((int)take));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 pack2(float x, float y) { f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ float sum2(f32x2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x + y; }
__device__ __forceinline__ f32x2 lerp2(f32x2 a, f32x2 b, f32x2 f) { return fma2(f, sub2(b, a), a); }

// sin / cos of pi * v on the MUFU pipe: v is reduced to [-1, 1] first, where sin.approx / cos.approx
// are good to 2^-21 absolute — the size of the fp32 rounding of the phase itself
__device__ __forceinline__ void sincospi_fast(float v, float *sn, float *cs) {
    const float M = 12582912.f;
    const float k = (v * 0.5f + M) - M;  // nearest integer of v / 2
    const float a = (v - 2.f * k) * CSPB_PI_F;
    *sn = __sinf(a);
    *cs = __cosf(a);
}

// -sin(chi) of the CTF on the MUFU pipe: chi / pi is reduced to [-1, 1] exactly, sin.approx is good to 2^-21
// there — below the fp32 rounding of chi itself (chi reaches hundreds of radians at the band edge)
__device__ __forceinline__ float ctf_from_chi_fast(float chi) {
    const float v = chi * (1.f / CSPB_PI_F);
    const float M = 12582912.f;
    const float k = (v * 0.5f + M) - M;
    return -__sinf((v - 2.f * k) * CSPB_PI_F);
}

// One warp per unit = (image, exactly PB poses).  Lanes walk the polar-patch band plan:
// coalesced 8-byte reads of the packed image, CTF synthesised once per sample (MUFU) and shared by the
// unit's poses, trilinear gather as 2 x 32-byte loads from the quad reference, interpolation in packed
// fp32x2 (one complex value per register pair), per-ring sums closed with 3 shuffles per ring band
// (lane%4 = ring offset inside the 4-ring band).  The pose loop is branch free and the slot loop is
// unrolled twice so that the 4 * PB gathers of two samples are in flight together: the kernel waits on
// gather latency, and 16 warps x 2 samples in flight (128 registers, r01g) beat 24 warps x 1 sample
// (80 registers) by 6 %; 3 samples / 12 warps and 2 samples / 20 warps (spills) were slower.
template <int PB, bool DDEF, int MODE, bool CUT = false>
#ifndef CSPB_SCORE_MINB
#define CSPB_SCORE_MINB 4
#endif
__global__ void __launch_bounds__(128, DDEF ? 4 : CSPB_SCORE_MINB) score_kernel(const ScoreArgs A) {
    // MODE 0: every pose has its own rotation and shift.  MODE 1 (SHARED): rotation and CTF of pose 0 hold
    // for all poses (the optimiser's +-x, +-y evaluations): one gather, PB phase ramps.  MODE 2 (SAMESHIFT):
    // the shift of pose 0 holds for all poses (the centre and the +-angle / +-defocus evaluations): the
    // image is shifted once per sample, PB gathers.
    constexpr bool SHARED = MODE == 1, SAMESHIFT = MODE == 2;
    __shared__ float s_pose[4][PB][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = blockIdx.x * 4 + warp;
    if (u >= A.n_units) return;
    const ScoreUnit un = A.units[u];
    const CtfCoef cc = A.ctf[un.image];
    if (lane < PB) {
        float m[8];
        const float *q = A.poses6 + (long long)(un.first_eval + lane) * 6;
        float r[9];
        euler_matrix(q[0], q[1], q[2], r);
        m[0] = r[0] * A.padf; m[1] = r[1] * A.padf;   // x = m00 i + m01 j
        m[2] = r[3] * A.padf; m[3] = r[4] * A.padf;   // y = m10 i + m11 j
        m[4] = r[6] * A.padf; m[5] = r[7] * A.padf;   // z = m20 i + m21 j
        m[6] = q[3] * A.inv_npx2;
        m[7] = q[4] * A.inv_npx2;
#pragma unroll
        for (int k = 0; k < 8; ++k) s_pose[warp][lane][k] = m[k];
    }
    __syncwarp();
    float ddef[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p)
        ddef[p] = DDEF ? A.poses6[(long long)(un.first_eval + p) * 6 + 5] * cc.dstep : 0.f;

    const float2 *img = A.packed + (long long)un.image * A.n_slots;
    const int origin = (A.rc * A.sy + A.rc) * A.sx;
    // per-ring-offset running sums {numerator, signed total} live in shared memory (lanes 0..3 of the
    // warp own them): they are touched once per ring band and would otherwise pin 2 PB registers
    __shared__ float2 s_ring[4][PB][4];
    f32x2 accB[PB];  // {sum px^2, sum py^2}
#pragma unroll
    for (int p = 0; p < PB; ++p) {
        accB[p] = 0ull;
        if (lane < 4) s_ring[warp][p][lane] = make_float2(0.f, 0.f);
    }
    f32x2 accA = 0ull;

    for (int b = 0; b < A.n_bands; ++b) {
        const BandDesc bd = A.bands[b];
        // CUT (coarse-to-fine stages of the analytic optimiser): this lane's ring is in or out for the whole band
        bool ring_on = true;
        if (CUT) {
            const int rp_ = (lane & 2) ? bd.rings23 : bd.rings01;
            ring_on = ((lane & 1) ? (rp_ >> 16) : (rp_ & 0xFFFF)) <= A.ring_cut;
            if (!__any_sync(0xffffffffu, ring_on)) continue;
        }
        f32x2 accX[PB];  // {sum gr px, sum gi py}
#pragma unroll
        for (int p = 0; p < PB; ++p) accX[p] = 0ull;
#ifndef CSPB_SCORE_UNROLL
#define CSPB_SCORE_UNROLL 2
#endif
#define CSPB_PRAGMA_(x) _Pragma(#x)
#define CSPB_UNROLL_(n) CSPB_PRAGMA_(unroll n)
        CSPB_UNROLL_(CSPB_SCORE_UNROLL)
        for (int it = 0; it < bd.n_iter; ++it) {
            const int slot = bd.slot_start + it * 32 + lane;
            const int32_t ij = __ldg(A.slot_ij + slot);
            f32x2 Fp = __ldcs(reinterpret_cast<const f32x2 *>(img + slot));  // read once per unit: streaming
            if (CUT && !ring_on) Fp = 0ull;
            const float2 F = make_float2(__uint_as_float((unsigned)Fp), __uint_as_float((unsigned)(Fp >> 32)));
            int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
            const bool valid = (i != CSPB_DUMMY_I) && (!CUT || ring_on);
            if (!valid) { i = 0; j = 0; }
            const float fi = (float)i, fj = (float)j;
            const float r2 = fi * fi + fj * fj;
            const float chi0 = ctf_chi(cc, fi, fj, r2);
            float ctfv = 0.f;
            if (!DDEF) ctfv = valid ? ctf_from_chi_fast(chi0) : 0.f;
            accA = fma2(Fp, Fp, accA);
            f32x2 G0 = 0ull;
            if (SAMESHIFT) {
                const float2 mc = *reinterpret_cast<const float2 *>(&s_pose[warp][0][6]);
                float sn, cs;
                sincospi_fast(fi * mc.x + fj * mc.y, &sn, &cs);
                G0 = pack2(F.x * cs - F.y * sn, F.x * sn + F.y * cs);
            }
            // SHARED units (the optimiser's +-x, +-y evaluations): rotation and CTF of pose 0 hold for
            // every pose of the unit — one gather, PB phase ramps
            constexpr int NG = SHARED ? 1 : PB;
            // The poses of a unit are neighbours (stencil +-h, line search): more often than not pose p
            // lands in the voxel pose 0 already gathered — those lanes reuse its quads and issue no
            // load (predicated), which removes their share of the L1 wavefronts.
            QuadP qa0, qa1;  // quads of pose 0, kept for the poses that land in the same voxel
            int off0 = 0;
#pragma unroll
            for (int p = 0; p < NG; ++p) {
                const float4 ma = *reinterpret_cast<const float4 *>(&s_pose[warp][p][0]);
                const float2 mb = *reinterpret_cast<const float2 *>(&s_pose[warp][p][4]);
                float x = ma.x * fi + ma.y * fj;
                float y = ma.z * fi + ma.w * fj;
                float z = mb.x * fi + mb.y * fj;
                // Friedel mate for the negative half space: s = sign(x)
                const float sgn = __int_as_float((__float_as_int(x) & 0x80000000) | 0x3f800000);
                x = fabsf(x); y *= sgn; z *= sgn;
                const int ix = __float2int_rd(x), iy = __float2int_rd(y), iz = __float2int_rd(z);
                const float f = x - (float)ix, fyv = y - (float)iy, fzv = z - (float)iz;
                const int off = origin + (iz * A.sy + iy) * A.sx + ix;
                const RefQuad *q = A.ref8 + off;
                QuadP q0, q1;
                if (p == 0) {
                    off0 = off;
                    qa0 = ldg_quadp(q);
                    qa1 = ldg_quadp(q + 1);
                    q0 = qa0;
                    q1 = qa1;
                } else {
                    q0 = qa0;
                    q1 = qa1;
                    ldg_quadp_if(q, off != off0, q0);
                    ldg_quadp_if(q + 1, off != off0, q1);
                }
                const f32x2 fx2 = pack2(f, f), fy2 = pack2(fyv, fyv), fz2 = pack2(fzv, fzv);
                const f32x2 v00 = lerp2(q0.v00, q1.v00, fx2), v10 = lerp2(q0.v10, q1.v10, fx2);
                const f32x2 v01 = lerp2(q0.v01, q1.v01, fx2), v11 = lerp2(q0.v11, q1.v11, fx2);
                const f32x2 v0 = lerp2(v00, v10, fy2), v1 = lerp2(v01, v11, fy2);
                float cv = ctfv;
                if (DDEF) cv = valid ? ctf_from_chi_fast(chi0 + r2 * ddef[p]) : 0.f;
                // CTF on both components, Friedel sign on the imaginary one
                const f32x2 P = mul2(lerp2(v0, v1, fz2), pack2(cv, cv * sgn));
                accB[p] = fma2(P, P, accB[p]);
                if (SAMESHIFT) {
                    accX[p] = fma2(G0, P, accX[p]);
                } else {
#pragma unroll
                    for (int pp = 0; pp < (SHARED ? PB : 1); ++pp) {
                        const int e = SHARED ? pp : p;
                        const float2 mc = *reinterpret_cast<const float2 *>(&s_pose[warp][e][6]);
                        float sn, cs;
                        sincospi_fast(fi * mc.x + fj * mc.y, &sn, &cs);
                        accX[e] = fma2(pack2(F.x * cs - F.y * sn, F.x * sn + F.y * cs), P, accX[e]);
                    }
                }
            }
        }
        const int rpair = (lane & 2) ? bd.rings23 : bd.rings01;
        const int ring = (lane & 1) ? (rpair >> 16) : (rpair & 0xFFFF);
#pragma unroll
        for (int p = 0; p < PB; ++p) {
            float v = sum2(accX[p]);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 4) {
                float2 t = s_ring[warp][p][lane];
                t.x += (ring > A.limit_ring) ? fabsf(v) : v;
                t.y += v;
                s_ring[warp][p][lane] = t;
            }
        }
    }
    const float asum = warp_sum(sum2(accA));
#pragma unroll
    for (int p = 0; p < PB; ++p) {
        const float bsum = warp_sum(sum2(accB[SHARED ? 0 : p]));
        const float2 t = s_ring[warp][p][lane & 3];
        float nv = t.x, xv = t.y;
        nv += __shfl_xor_sync(0xffffffffu, nv, 1);
        nv += __shfl_xor_sync(0xffffffffu, nv, 2);
        xv += __shfl_xor_sync(0xffffffffu, xv, 1);
        xv += __shfl_xor_sync(0xffffffffu, xv, 2);
        if (lane == 0) A.out[un.first_eval + p] = make_float4(nv, xv, asum, bsum);
    }
}

// ---- analytic-gradient evaluation (oracle/SEMANTICS.md §7c; CPU restatement: orc_score_grad).
// One warp per (image, pose).  The eight corners a sample gathers give the value AND the spatial gradient of the
// trilinear interpolant; the chain rule through R(psi, theta, phi) and the phase ramp turn them into
// dP/d(psi, theta, phi, x, y).  Besides the score sums the warp accumulates d num / d p (with the signed / absolute ring
// rule, hence per ring), d B / d angles and the 15 entries of J^T J — everything the Gauss-Newton step needs from ONE
// gather per sample instead of the 11 evaluations of the central-difference stencil.  Constant factors (degrees,
// 2 pi / (n pixel)) are applied once at the end.
#define CSPB_GOUT 28
__global__ void __launch_bounds__(128, 3) score_grad_kernel(const ScoreArgs A) {
    __shared__ float s_pose[4][16];     // 0..5 first two columns of pad*R, 6..11 of pad*dR/dtheta, 12..13 shift phase per index
    __shared__ float s_tot[4][4][8];    // per warp and ring track: {num, X, dnum[5]} running totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = blockIdx.x * 4 + warp;
    if (u >= A.n_units) return;
    const ScoreUnit un = A.units[u];
    const CtfCoef cc = A.ctf[un.image];
    if (lane == 0) {
        const float *q = A.poses6 + (long long)un.first_eval * 6;
        float r[9];
        euler_matrix(q[0], q[1], q[2], r);
        float sps, cps, sth, cth, sph, cph;
        sincosf(q[0] * (CSPB_PI_F / 180.f), &sps, &cps);
        sincosf(q[1] * (CSPB_PI_F / 180.f), &sth, &cth);
        sincosf(q[2] * (CSPB_PI_F / 180.f), &sph, &cph);
        float *o = s_pose[warp];
        o[0] = r[0] * A.padf; o[1] = r[1] * A.padf; o[2] = r[3] * A.padf; o[3] = r[4] * A.padf; o[4] = r[6] * A.padf; o[5] = r[7] * A.padf;
        o[6] = -cph * sth * cps * A.padf; o[7] = cph * sth * sps * A.padf;
        o[8] = -sph * sth * cps * A.padf; o[9] = sph * sth * sps * A.padf;
        o[10] = -cth * cps * A.padf;      o[11] = cth * sps * A.padf;
        o[12] = q[3] * A.inv_npx2; o[13] = q[4] * A.inv_npx2;
    }
    if (lane < 4)
#pragma unroll
        for (int k = 0; k < 8; ++k) s_tot[warp][lane][k] = 0.f;
    __syncwarp();
    const float *sp = s_pose[warp];
    const float m0 = sp[0], m1 = sp[1], m2 = sp[2], m3 = sp[3], m4 = sp[4], m5 = sp[5];
    const float t0 = sp[6], t1 = sp[7], t2 = sp[8], t3 = sp[9], t4 = sp[10], t5 = sp[11];
    const float mx = sp[12], my = sp[13];
    const float2 *img = A.packed + (long long)un.image * A.n_slots;
    const int origin = (A.rc * A.sy + A.rc) * A.sx;
    f32x2 aA = 0ull, aB = 0ull, adB[3] = {0ull, 0ull, 0ull}, aJ[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) aJ[k] = 0ull;
    for (int b = 0; b < A.n_bands; ++b) {
        const BandDesc bd = A.bands[b];
        const int rpair = (lane & 2) ? bd.rings23 : bd.rings01;
        const int ring = (lane & 1) ? (rpair >> 16) : (rpair & 0xFFFF);
        const bool ring_on = ring <= A.ring_cut;
        if (!__any_sync(0xffffffffu, ring_on)) continue;
        f32x2 aX = 0ull, adX[3] = {0ull, 0ull, 0ull}, aSx = 0ull, aSy = 0ull, aAr = 0ull, aBr = 0ull;
#ifndef CSPB_GRAD_UNROLL
#define CSPB_GRAD_UNROLL 2  // two samples in flight per lane at 159 registers / 12 warps per SM: 5 % faster than one at 126 / 16 (r02e)
#endif
        CSPB_UNROLL_(CSPB_GRAD_UNROLL)
        for (int it = 0; it < bd.n_iter; ++it) {
            const int slot = bd.slot_start + it * 32 + lane;
            const int32_t ij = __ldg(A.slot_ij + slot);
            f32x2 Fp = __ldcs(reinterpret_cast<const f32x2 *>(img + slot));
            if (!ring_on) Fp = 0ull;
            const float2 F = make_float2(__uint_as_float((unsigned)Fp), __uint_as_float((unsigned)(Fp >> 32)));
            int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
            const bool valid = (i != CSPB_DUMMY_I) && ring_on;
            if (i == CSPB_DUMMY_I) { i = 0; j = 0; }
            const float fi = (float)i, fj = (float)j;
            const float r2 = fi * fi + fj * fj;
            const float cv = valid ? ctf_from_chi_fast(ctf_chi(cc, fi, fj, r2)) : 0.f;
            float x = m0 * fi + m1 * fj, y = m2 * fi + m3 * fj, z = m4 * fi + m5 * fj;
            // coordinate velocities per radian (before the Friedel flip): psi = R (-j, i), theta = dR/dtheta (i, j), phi = (-y, x, 0)
            const float vpx = m1 * fi - m0 * fj, vpy = m3 * fi - m2 * fj, vpz = m5 * fi - m4 * fj;
            const float vtx = t0 * fi + t1 * fj, vty = t2 * fi + t3 * fj, vtz = t4 * fi + t5 * fj;
            const float vfx = -y, vfy = x;
            const float sgn = __int_as_float((__float_as_int(x) & 0x80000000) | 0x3f800000);
            x = fabsf(x); y *= sgn; z *= sgn;
            const int ix = __float2int_rd(x), iy = __float2int_rd(y), iz = __float2int_rd(z);
            const float f = x - (float)ix, fyv = y - (float)iy, fzv = z - (float)iz;
            const RefQuad *q = A.ref8 + origin + (iz * A.sy + iy) * A.sx + ix;
            const QuadP q0 = ldg_quadp(q), q1 = ldg_quadp(q + 1);
            const f32x2 fx2 = pack2(f, f), fy2 = pack2(fyv, fyv), fz2 = pack2(fzv, fzv);
            // along x first: a = c0 + fx (c1 - c0); the four x-differences interpolate to dT/dx
            const f32x2 d00 = sub2(q1.v00, q0.v00), d10 = sub2(q1.v10, q0.v10), d01 = sub2(q1.v01, q0.v01), d11 = sub2(q1.v11, q0.v11);
            const f32x2 a00 = fma2(fx2, d00, q0.v00), a10 = fma2(fx2, d10, q0.v10), a01 = fma2(fx2, d01, q0.v01), a11 = fma2(fx2, d11, q0.v11);
            const f32x2 dx0 = lerp2(d00, d10, fy2), dx1 = lerp2(d01, d11, fy2);
            const f32x2 gx = lerp2(dx0, dx1, fz2);
            const f32x2 dy0 = sub2(a10, a00), dy1 = sub2(a11, a01);
            const f32x2 v0 = fma2(fy2, dy0, a00), v1 = fma2(fy2, dy1, a01);
            const f32x2 gy = lerp2(dy0, dy1, fz2);
            const f32x2 gz = sub2(v1, v0);
            const f32x2 val = fma2(fz2, gz, v0);
            const f32x2 cP = pack2(cv, cv * sgn), cD = pack2(cv * sgn, cv);
            const f32x2 P = mul2(val, cP);
            const f32x2 dPp = mul2(fma2(gz, pack2(vpz, vpz), fma2(gy, pack2(vpy, vpy), mul2(gx, pack2(vpx, vpx)))), cD);
            const f32x2 dPt = mul2(fma2(gz, pack2(vtz, vtz), fma2(gy, pack2(vty, vty), mul2(gx, pack2(vtx, vtx)))), cD);
            const f32x2 dPf = mul2(fma2(gy, pack2(vfy, vfy), mul2(gx, pack2(vfx, vfx))), cD);
            float sn, cs;
            sincospi_fast(fi * mx + fj * my, &sn, &cs);
            const f32x2 G = pack2(F.x * cs - F.y * sn, F.x * sn + F.y * cs);
            // Q = -i P: d/dshift of the equivalent projection shift (up to 2 pi / (n pixel) times the index)
            const float Pre = __uint_as_float((unsigned)P), Pim = __uint_as_float((unsigned)(P >> 32));
            const f32x2 Q = pack2(Pim, -Pre);
            const f32x2 fi2 = pack2(fi, fi), fj2 = pack2(fj, fj);
            aAr = fma2(Fp, Fp, aAr);
            aBr = fma2(P, P, aBr);
            aX = fma2(G, P, aX);
            adX[0] = fma2(G, dPp, adX[0]); adX[1] = fma2(G, dPt, adX[1]); adX[2] = fma2(G, dPf, adX[2]);
            const f32x2 GQ = mul2(G, Q);
            aSx = fma2(fi2, GQ, aSx); aSy = fma2(fj2, GQ, aSy);
            adB[0] = fma2(P, dPp, adB[0]); adB[1] = fma2(P, dPt, adB[1]); adB[2] = fma2(P, dPf, adB[2]);
            // J^T J, upper triangle row major over (psi, theta, phi, x, y)
            const f32x2 Qx = mul2(Q, fi2), Qy = mul2(Q, fj2);
            aJ[0] = fma2(dPp, dPp, aJ[0]); aJ[1] = fma2(dPp, dPt, aJ[1]); aJ[2] = fma2(dPp, dPf, aJ[2]); aJ[3] = fma2(dPp, Qx, aJ[3]); aJ[4] = fma2(dPp, Qy, aJ[4]);
            aJ[5] = fma2(dPt, dPt, aJ[5]); aJ[6] = fma2(dPt, dPf, aJ[6]); aJ[7] = fma2(dPt, Qx, aJ[7]); aJ[8] = fma2(dPt, Qy, aJ[8]);
            aJ[9] = fma2(dPf, dPf, aJ[9]); aJ[10] = fma2(dPf, Qx, aJ[10]); aJ[11] = fma2(dPf, Qy, aJ[11]);
            aJ[12] = fma2(Qx, Qx, aJ[12]); aJ[13] = fma2(Qx, Qy, aJ[13]); aJ[14] = fma2(Qy, Qy, aJ[14]);
        }
        // close the band: ring sums of X, its five derivatives, A_r and B_r (lanes with the same lane % 4 share a ring)
        float rs[8] = {sum2(aX), sum2(adX[0]), sum2(adX[1]), sum2(adX[2]), sum2(aSx), sum2(aSy), sum2(aAr), sum2(aBr)};
        aA = pack2(__uint_as_float((unsigned)aA) + rs[6], __uint_as_float((unsigned)(aA >> 32)));
        aB = pack2(__uint_as_float((unsigned)aB) + rs[7], __uint_as_float((unsigned)(aB >> 32)));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            rs[k] += __shfl_xor_sync(0xffffffffu, rs[k], 4);
            rs[k] += __shfl_xor_sync(0xffffffffu, rs[k], 8);
            rs[k] += __shfl_xor_sync(0xffffffffu, rs[k], 16);
        }
        if (lane < 4) {
            // |X_r| above the signed-CC limit; its derivative with the soft sign X_r / sqrt(X_r^2 + (0.02)^2 A_r B_r)
            // (SEMANTICS.md §7c): continuous where the objective has its kinks
            float sg = 1.f;
            const bool absr = ring > A.limit_ring;
            if (absr) {
                const float q = rs[0] * rs[0] + 4e-4f * rs[6] * rs[7];
                sg = q > 0.f ? rs[0] * rsqrtf(q) : 0.f;
            }
            float *t = s_tot[warp][lane];
            t[0] += absr ? fabsf(rs[0]) : rs[0];
            t[1] += rs[0];
#pragma unroll
            for (int k = 0; k < 5; ++k) t[2 + k] += sg * rs[1 + k];
        }
    }
    __syncwarp();
    float o[CSPB_GOUT];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        float v = s_tot[warp][lane & 3][k];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        o[k < 2 ? k : k + 2] = v;  // num, X, then dnum at 4..8
    }
    o[2] = warp_sum(sum2(aA));
    o[3] = warp_sum(sum2(aB));
#pragma unroll
    for (int k = 0; k < 3; ++k) o[9 + k] = warp_sum(sum2(adB[k]));
#pragma unroll
    for (int k = 0; k < 15; ++k) o[12 + k] = warp_sum(sum2(aJ[k]));
    o[27] = 0.f;
    if (lane == 0) {
        // units: angles per degree, shifts per Angstrom (phase = pi * inv_npx2 * index * shift)
        const float d2r = CSPB_PI_F / 180.f, k2 = CSPB_PI_F * A.inv_npx2;
        const float su[5] = {d2r, d2r, d2r, k2, k2};
#pragma unroll
        for (int k = 0; k < 5; ++k) o[4 + k] *= su[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) o[9 + k] *= 2.f * d2r;
        int t = 12;
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int b2 = a; b2 < 5; ++b2) o[t++] *= su[a] * su[b2];
        float4 *dst = reinterpret_cast<float4 *>(A.gout + (long long)un.first_eval * CSPB_GOUT);
#pragma unroll
        for (int k = 0; k < CSPB_GOUT / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
}

// Gather-load census of one scorer launch (roofline bookkeeping, never on the product path): the same
// units and address arithmetic as score_kernel, no loads — counts the 32-byte quad loads that kernel
// issues per lane (2 for the unit's first pose, 2 more for every other pose that leaves its voxel;
// shared-gather units gather once).  bench.py runs one untimed step with the census on and reports
// loaded bytes / s against the measured gather peak.
__global__ void __launch_bounds__(128) score_census_kernel(const ScoreArgs A, int PB, int mode, unsigned long long *__restrict__ total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = blockIdx.x * 4 + warp;
    if (u >= A.n_units) return;
    const ScoreUnit un = A.units[u];
    __shared__ float s_m[4][4][6];
    if (lane < PB) {
        const float *q = A.poses6 + (long long)(un.first_eval + lane) * 6;
        float r[9];
        euler_matrix(q[0], q[1], q[2], r);
        s_m[warp][lane][0] = r[0] * A.padf; s_m[warp][lane][1] = r[1] * A.padf;
        s_m[warp][lane][2] = r[3] * A.padf; s_m[warp][lane][3] = r[4] * A.padf;
        s_m[warp][lane][4] = r[6] * A.padf; s_m[warp][lane][5] = r[7] * A.padf;
    }
    __syncwarp();
    const int origin = (A.rc * A.sy + A.rc) * A.sx;
    const int NG = mode == 1 ? 1 : PB;
    unsigned cnt = 0, slots = 0;
    for (int b = 0; b < A.n_bands; ++b) {
        const BandDesc bd = A.bands[b];
        const int rp_ = (lane & 2) ? bd.rings23 : bd.rings01;
        if (!__any_sync(0xffffffffu, ((lane & 1) ? (rp_ >> 16) : (rp_ & 0xFFFF)) <= A.ring_cut)) continue;  // band above the stage's cut
        for (int it = 0; it < bd.n_iter; ++it) {
            const int32_t ij = A.slot_ij[bd.slot_start + it * 32 + lane];
            int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
            if (i == CSPB_DUMMY_I) { i = 0; j = 0; }
            const float fi = (float)i, fj = (float)j;
            int off0 = 0;
            ++slots;  // one 8-byte read of the packed image per lane and slot
            for (int p = 0; p < NG; ++p) {
                const float *m = s_m[warp][p];
                float x = m[0] * fi + m[1] * fj, y = m[2] * fi + m[3] * fj, z = m[4] * fi + m[5] * fj;
                const float sgn = x < 0.f ? -1.f : 1.f;
                x = fabsf(x); y *= sgn; z *= sgn;
                const int off = origin + (__float2int_rd(z) * A.sy + __float2int_rd(y)) * A.sx + __float2int_rd(x);
                if (p == 0) { off0 = off; cnt += 2; }
                else if (off != off0) cnt += 2;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        slots += __shfl_xor_sync(0xffffffffu, slots, o);
    }
    if (lane == 0) {
        atomicAdd(total, (unsigned long long)cnt);
        atomicAdd(total + 1, (unsigned long long)slots);
    }
}

// central slice on the full half-plane grid (building block / test helper)
__global__ void project_kernel(const float4 *__restrict__ ref4, int sx, int sy, int rc, float padf, int n,
                               float r_hi, float psi, float theta, float phi, float2 *__restrict__ out) {
    const int nh = n / 2 + 1;
    float r[9];
    euler_matrix(psi, theta, phi, r);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * nh; idx += gridDim.x * blockDim.x) {
        const int i = idx % nh;
        int j = idx / nh;
        if (j >= n / 2) j -= n;
        float2 v = make_float2(0.f, 0.f);
        const float fi = (float)i, fj = (float)j;
        if (fi * fi + fj * fj <= r_hi * r_hi) {
            float x = (r[0] * fi + r[1] * fj) * padf, y = (r[3] * fi + r[4] * fj) * padf,
                  z = (r[6] * fi + r[7] * fj) * padf;
            v = gather_trilinear(ref4, sx, sy, rc, x, y, z);
        }
        out[idx] = v;
    }
}

__global__ void ctf_image_kernel(CtfCoef cc, int n, float *__restrict__ out) {
    const int nh = n / 2 + 1;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * nh; idx += gridDim.x * blockDim.x) {
        const int i = idx % nh;
        int j = idx / nh;
        if (j >= n / 2) j -= n;
        const float fi = (float)i, fj = (float)j;
        out[idx] = -sinpif(ctf_chi(cc, fi, fj, fi * fi + fj * fj) * (1.f / CSPB_PI_F));
    }
}

// ================================================================== batched local optimiser (shared pieces in opt.cuh)
struct SearchHit {
    float score, sx, sy;
    int orient;
};

// state k belongs to image k / K.  With hits (global search) the state starts at candidate k % K.
__global__ void opt_init_kernel(const cspb_row *__restrict__ rows, int n_states, int K, const SearchHit *__restrict__ hits,
                                const float *__restrict__ angles3, OptState *__restrict__ st, float h_ang, float h_shift,
                                float h_def, float lam_scale) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_states) return;
    const cspb_row row = rows[k / K];
    OptState s;
    s.x[0] = row.psi; s.x[1] = row.theta; s.x[2] = row.phi;
    s.x[3] = row.x_shift; s.x[4] = row.y_shift; s.x[5] = 0.f;
    if (hits) {
        const SearchHit h = hits[k];
        if (h.orient >= 0) {
            s.x[0] = angles3[3 * h.orient]; s.x[1] = angles3[3 * h.orient + 1]; s.x[2] = angles3[3 * h.orient + 2];
            s.x[3] = h.sx; s.x[4] = h.sy;
        }
    }
    s.h[0] = s.h[1] = s.h[2] = h_ang;
    s.h[3] = s.h[4] = h_shift;
    s.h[5] = h_def;
    for (int m = 0; m < OPT_NP; ++m) { s.d[m] = 0.f; s.x0[m] = s.x[m]; }
    s.f = 0.f; s.lam = lam_scale * row.sigma * row.sigma; s.pad_[0] = s.pad_[1] = 0.f;
    st[k] = s;
}

// final: evaluate the refined pose and the starting pose (2 evals per image)
__global__ void opt_finish_eval_kernel(const OptState *__restrict__ st, int n, int K, float *__restrict__ poses6,
                                       ScoreUnit *__restrict__ units) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    for (int m = 0; m < OPT_NP; ++m) {
        poses6[(long long)k * 12 + m] = st[k].x[m];
        poses6[(long long)k * 12 + 6 + m] = st[k].x0[m];
    }
    ScoreUnit un;
    un.image = k / K; un.first_eval = 2 * k; un.count = 2; un.pad_ = 0;
    units[k] = un;
}

// one thread per image: among its K refined candidates (and each candidate's starting pose) keep
// the best score; ties keep the earlier candidate
__global__ void opt_write_rows_kernel(const OptState *__restrict__ st, int n_images, int K, const float4 *__restrict__ sc,
                                      int n_samples, int refine_defocus, cspb_row *__restrict__ rows,
                                      cspb_row *__restrict__ changes, const OptPrior pr, float *__restrict__ alpha_out) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= n_images) return;
    cspb_row r = rows[img];
    const cspb_row old = r;
    float best = -1e30f, best_obj = -1e30f;
    float4 vbest = make_float4(0.f, 0.f, 0.f, 0.f);
    float xb[OPT_NP] = {r.psi, r.theta, r.phi, r.x_shift, r.y_shift, 0.f};
    float start_score = 0.f;
    for (int c = 0; c < K; ++c) {
        const int k = img * K + c;
        const OptState s = st[k];
        float4 v = sc[2 * k];
        const float4 v0 = sc[2 * k + 1];
        const float *x = s.x;
        // never return a worse pose (objective = score - shift restraint; the restraint is 0 without priors)
        if (100.f * cc_of(v) - 100.f * prior_pen(pr, s.lam, s.x[3], s.x[4]) <
            100.f * cc_of(v0) - 100.f * prior_pen(pr, s.lam, s.x0[3], s.x0[4])) { v = v0; x = s.x0; }
        if (c == 0) start_score = 100.f * cc_of(v0);
        const float f = 100.f * cc_of(v);
        const float obj = cc_of(v) - prior_pen(pr, s.lam, x[3], x[4]);
        if (obj > best_obj) {
            best_obj = obj;
            best = f;
            vbest = v;
            for (int m = 0; m < OPT_NP; ++m) xb[m] = x[m];
        }
    }
    r.psi = wrap360(xb[0]);
    r.theta = xb[1];
    r.phi = wrap360(xb[2]);
    r.x_shift = xb[3];
    r.y_shift = xb[4];
    if (refine_defocus) {
        r.defocus_1 += xb[5];
        r.defocus_2 += xb[5];
    }
    float sigma, logp;
    score_stats(vbest, n_samples, &sigma, &logp);
    r.score = best;
    r.sigma = sigma;
    r.logp = logp;
    rows[img] = r;
    if (alpha_out) alpha_out[img] = vbest.w > 0.f ? vbest.y / vbest.w : 0.f;
    if (changes) {
        cspb_row c = r;
        // angular changes wrapped into [-180, 180): a refinement that crosses 0/360 is a small change
        c.psi = wrap180(r.psi - old.psi); c.theta = wrap180(r.theta - old.theta); c.phi = wrap180(r.phi - old.phi);
        c.x_shift = r.x_shift - old.x_shift; c.y_shift = r.y_shift - old.y_shift;
        c.defocus_1 = r.defocus_1 - old.defocus_1; c.defocus_2 = r.defocus_2 - old.defocus_2;
        c.score = r.score - start_score;
        c.logp = r.logp - old.logp; c.sigma = r.sigma - old.sigma;
        changes[img] = c;
    }
}

__global__ void scores_from_out_kernel(const float4 *__restrict__ sc, int n, float *__restrict__ scores) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) scores[k] = 100.f * cc_of(sc[k]);
}

// build units from an image-sorted eval list: consecutive evals of one image, <= PB per unit,
// bucketed by count (units[c] holds the units with exactly c poses, c = 1..4)
void build_units_host(const int32_t *image_index, int n_evals, int PB, std::vector<ScoreUnit> *units /* [5] */) {
    for (int c = 0; c <= 4; ++c) units[c].clear();
    int e = 0;
    while (e < n_evals) {
        int c = 1;
        while (c < PB && e + c < n_evals && image_index[e + c] == image_index[e]) ++c;
        ScoreUnit un;
        un.image = image_index[e]; un.first_eval = e; un.count = c; un.pad_ = 0;
        units[c].push_back(un);
        e += c;
    }
}

int grid_for(long long total, int block, int sm) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)sm * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// every unit in [d_units, d_units + n_units) has exactly `count` poses (1..4)
static void fill_score_args(cspb_ctx *ctx, ScoreArgs &a, const ScoreUnit *d_units, int n_units, const float *d_poses6, const CtfCoef *d_ctf);

// analytic-gradient evaluation of one pose per unit (score_grad_kernel); gout: CSPB_GOUT floats per evaluation
int launch_score_grad(cspb_ctx *ctx, const ScoreUnit *d_units, int n_units, const float *d_poses6, const CtfCoef *d_ctf, float *d_gout,
                      int ring_cut) {
    if (n_units <= 0) return 0;
    ScoreArgs a;
    fill_score_args(ctx, a, d_units, n_units, d_poses6, d_ctf);
    a.ring_cut = ring_cut;
    a.out = nullptr;
    a.gout = d_gout;
    prof_begin(ctx, CSPB_PROF_SCORE, n_units);
    score_grad_kernel<<<ceil_div(n_units, 4), 128, 0, ctx->stream>>>(a);
    prof_end(ctx);
    KERNEL_CHECK(ctx);
    if (ctx->count_loads) {  // one gather (two quad loads) per lane and slot of the bands inside the cut
        score_census_kernel<<<ceil_div(n_units, 4), 128, 0, ctx->stream>>>(a, 1, 0, ctx->d_load_count.as<unsigned long long>());
        KERNEL_CHECK(ctx);
        ctx->census_evals += n_units;
    }
    return 0;
}

static void fill_score_args(cspb_ctx *ctx, ScoreArgs &a, const ScoreUnit *d_units, int n_units, const float *d_poses6, const CtfCoef *d_ctf) {
    a.ref8 = ctx->ref.d_ref8.as<RefQuad>();
    a.sx = ctx->ref.sx8; a.sy = ctx->ref.sy; a.rc = ctx->ref.rc;
    a.padf = (float)ctx->ref.pad;
    a.slot_ij = ctx->plan.d_slot_ij.as<int32_t>();
    a.bands = ctx->plan.d_bands.as<BandDesc>();
    a.n_bands = ctx->plan.n_bands;
    a.n_slots = ctx->plan.n_slots;
    a.packed = ctx->d_packed.as<float2>();
    a.ctf = d_ctf;
    a.units = d_units;
    a.n_units = n_units;
    a.poses6 = d_poses6;
    a.inv_npx2 = 2.f / ((float)ctx->rcfg.box * ctx->rcfg.pixel_size);
    const float lim = ctx->rcfg.signed_cc_limit;
    a.limit_ring = lim > 0.f ? (int)floorf((float)ctx->rcfg.box * ctx->rcfg.pixel_size / lim) : 0x7fffffff;
    a.ring_cut = 0x7fffffff;
    a.out = nullptr;
    a.gout = nullptr;
}

int launch_score(cspb_ctx *ctx, const ScoreUnit *d_units, int n_units, int count, const float *d_poses6,
                 const CtfCoef *d_ctf, float4 *d_out, bool ddef, int64_t n_evals, int mode, int ring_cut) {
    if (n_units <= 0) return 0;
    if (count < 1 || count > 4) return cspb_fail(ctx, CSPB_E_ARG, "launch_score: bad unit size %d", count);
    if (ring_cut != 0x7fffffff && (count != 1 || ddef || mode == 1)) return cspb_fail(ctx, CSPB_E_ARG, "launch_score: ring cut needs single-pose units");
    ScoreArgs a;
    a.ref8 = ctx->ref.d_ref8.as<RefQuad>();
    a.sx = ctx->ref.sx8; a.sy = ctx->ref.sy; a.rc = ctx->ref.rc;
    a.padf = (float)ctx->ref.pad;
    a.slot_ij = ctx->plan.d_slot_ij.as<int32_t>();
    a.bands = ctx->plan.d_bands.as<BandDesc>();
    a.n_bands = ctx->plan.n_bands;
    a.n_slots = ctx->plan.n_slots;
    a.packed = ctx->d_packed.as<float2>();
    a.ctf = d_ctf;
    a.units = d_units;
    a.n_units = n_units;
    a.poses6 = d_poses6;
    a.inv_npx2 = 2.f / ((float)ctx->rcfg.box * ctx->rcfg.pixel_size);
    const float lim = ctx->rcfg.signed_cc_limit;
    a.limit_ring = lim > 0.f ? (int)floorf((float)ctx->rcfg.box * ctx->rcfg.pixel_size / lim) : 0x7fffffff;
    a.ring_cut = ring_cut;
    a.out = d_out;
    a.gout = nullptr;
    const int grid = ceil_div(n_units, 4);
    prof_begin(ctx, CSPB_PROF_SCORE, n_evals);
    if (ring_cut != 0x7fffffff) {
        score_kernel<1, false, 0, true><<<grid, 128, 0, ctx->stream>>>(a);
    } else
#define CSPB_LAUNCH_SCORE(PB_)                                                              \
    do {                                                                                    \
        if (mode == 1) {                                                                    \
            if (ddef) score_kernel<PB_, true, 1><<<grid, 128, 0, ctx->stream>>>(a);         \
            else score_kernel<PB_, false, 1><<<grid, 128, 0, ctx->stream>>>(a);             \
        } else if (mode == 2 && PB_ > 1) {                                                  \
            if (ddef) score_kernel<PB_, true, 2><<<grid, 128, 0, ctx->stream>>>(a);         \
            else score_kernel<PB_, false, 2><<<grid, 128, 0, ctx->stream>>>(a);             \
        } else {                                                                            \
            if (ddef) score_kernel<PB_, true, 0><<<grid, 128, 0, ctx->stream>>>(a);         \
            else score_kernel<PB_, false, 0><<<grid, 128, 0, ctx->stream>>>(a);             \
        }                                                                                   \
    } while (0)
    switch (count) {
    case 1: CSPB_LAUNCH_SCORE(1); break;
    case 2: CSPB_LAUNCH_SCORE(2); break;
    case 3: CSPB_LAUNCH_SCORE(3); break;
    default: CSPB_LAUNCH_SCORE(4); break;
    }
#undef CSPB_LAUNCH_SCORE
    prof_end(ctx);
    KERNEL_CHECK(ctx);
    if (ctx->count_loads) {
        score_census_kernel<<<grid, 128, 0, ctx->stream>>>(a, count, (mode == 2 && count == 1) ? 0 : mode, ctx->d_load_count.as<unsigned long long>());
        KERNEL_CHECK(ctx);
        ctx->census_evals += n_evals;
    }
    return 0;
}

// units in the class layout of opt_stencil_kernel / csp_expand_kernel:
// [A full: n*(nA/PB)] [A tail: n if nA%PB] [S full: n*(nS/PB)] [S tail: n if nS%PB]; S = shared units
int launch_score_classes(cspb_ctx *ctx, const ScoreUnit *d_units, int n_groups, int nA, int nS, int PB, const float *d_poses6,
                         const CtfCoef *d_ctf, float4 *d_out, bool ddef, bool a_same_shift) {
    const int cls[4][2] = {{nA / PB, PB}, {nA % PB ? 1 : 0, nA % PB}, {nS / PB, PB}, {nS % PB ? 1 : 0, nS % PB}};
    size_t off = 0;
    for (int c = 0; c < 4; ++c) {
        const int n_units = n_groups * cls[c][0];
        if (n_units == 0) continue;
        int rc = launch_score(ctx, d_units + off, n_units, cls[c][1], d_poses6, d_ctf, d_out, ddef, (int64_t)n_units * cls[c][1],
                              c >= 2 ? 1 : (a_same_shift ? 2 : 0));
        if (rc) return rc;
        off += n_units;
    }
    return 0;
}

namespace {

// nearest-ring CSR of the full half plane (for the noise power curve)
void build_ring_csr(int n, std::vector<int> &off, std::vector<int> &idx, std::vector<int> &count) {
    const int nh = n / 2 + 1, n_rings = n + 1;
    count.assign(n_rings, 0);
    std::vector<int> ring_of((size_t)n * nh);
    for (int jj = 0; jj < n; ++jj)
        for (int i = 0; i < nh; ++i) {
            const int j = jj >= n / 2 ? jj - n : jj;
            const int r = (int)(sqrtf((float)(i * i + j * j)) + 0.5f);
            ring_of[(size_t)jj * nh + i] = r;
            count[r]++;
        }
    off.assign(n_rings + 1, 0);
    for (int r = 0; r < n_rings; ++r) off[r + 1] = off[r] + count[r];
    idx.resize((size_t)n * nh);
    std::vector<int> cur(off.begin(), off.end() - 1);
    for (size_t k = 0; k < ring_of.size(); ++k) idx[cur[ring_of[k]]++] = (int)k;
}

}  // namespace

// ================================================================== C-ABI: refine
extern "C" int cspb_refine_cfg_default(cspb_refine_cfg *cfg, int box, float pixel_size) {
    if (!cfg || box <= 0 || pixel_size <= 0.f) return CSPB_E_ARG;
    memset(cfg, 0, sizeof *cfg);
    cfg->box = box;
    cfg->pad = 1;
    cfg->pixel_size = pixel_size;
    cfg->mask_radius = 0.375f * box * pixel_size;
    cfg->low_res_limit = 100.f;   // config/pyp_config.toml refine.rlref default
    cfg->high_res_limit = 2.5f * pixel_size;
    cfg->signed_cc_limit = 30.f;  // frealign.py:3879
    cfg->search_mask_radius = 1.5f * cfg->mask_radius;
    cfg->search_high_res = 8.f * pixel_size;
    cfg->angular_step = 20.f;
    cfg->best_matches = 20;
    cfg->search_range_x = cfg->search_range_y = 0.125f * box * pixel_size;
    cfg->defocus_range = 500.f;
    cfg->defocus_step = 50.f;
    cfg->global_search = 0;
    cfg->local_refine = 1;
    cfg->refine_psi = cfg->refine_theta = cfg->refine_phi = cfg->refine_x = cfg->refine_y = 1;
    cfg->refine_defocus = 0;
    cfg->apply_mask = 1;
    cfg->normalize = 1;
    cfg->invert_contrast = 0;
    cfg->whiten = 1;
    cfg->symmetry_order = 1;
    cfg->local_iterations = 8;
    return 0;
}

extern "C" int cspb_refine_configure(cspb_ctx *ctx, const cspb_refine_cfg *cfg) {
    CSPB_ENTER(ctx);
    if (!ctx || !cfg) return CSPB_E_ARG;
    if (cfg->box < 16 || (cfg->box & 1) || cfg->pixel_size <= 0.f || (cfg->pad != 1 && cfg->pad != 2))
        return cspb_fail(ctx, CSPB_E_ARG, "bad box/pixel/pad");
    if (cfg->high_res_limit <= 0.f || cfg->low_res_limit <= cfg->high_res_limit)
        return cspb_fail(ctx, CSPB_E_ARG, "need low_res_limit > high_res_limit > 0");
    ctx->rcfg = *cfg;
    const float npx = (float)cfg->box * cfg->pixel_size;
    float r_lo = npx / cfg->low_res_limit, r_hi = npx / cfg->high_res_limit;
    const float r_cap = (float)(cfg->box / 2 - 2);  // keep every trilinear corner below Nyquist
    if (r_hi > r_cap) r_hi = r_cap;
    // the half-sphere of 32-byte quads the scorer touches: beyond the L2 (126 MB) the band keeps the radial ring order
    const double rc_q = ceil((cfg->pad >= 2 ? 2.0 : 1.0) * r_hi) + 2.0;
    const bool radial = (2.0 / 3.0) * 3.14159265 * rc_q * rc_q * rc_q * 32.0 > 110e6;
    if (!build_band_plan(ctx->plan, cfg->box, r_lo, r_hi, radial)) return cspb_fail(ctx, CSPB_E_ARG, "empty band");
    BandPlan &pl = ctx->plan;
    RESERVE(ctx, pl.d_slot_ij, pl.slot_ij.size() * sizeof(int32_t));
    RESERVE(ctx, pl.d_bands, pl.bands.size() * sizeof(BandDesc));
    CU_TRY(ctx, cudaMemcpyAsync(pl.d_slot_ij.p, pl.slot_ij.data(), pl.slot_ij.size() * sizeof(int32_t),
                                cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(pl.d_bands.p, pl.bands.data(), pl.bands.size() * sizeof(BandDesc),
                                cudaMemcpyHostToDevice, ctx->stream));
    {   // inverse map of the band plan for the fused preprocessing (fft2_whiten_mask_pack_dev)
        const int n = cfg->box, nh = n / 2 + 1;
        std::vector<int32_t> slot_of((size_t)n * nh, -1), dummy;
        for (size_t sl = 0; sl < pl.slot_ij.size(); ++sl) {
            const int32_t ij = pl.slot_ij[sl];
            const int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
            if (i == CSPB_DUMMY_I) { dummy.push_back((int32_t)sl); continue; }
            slot_of[(size_t)(j < 0 ? j + n : j) * nh + i] = (int32_t)sl;
        }
        pl.n_dummy = (int)dummy.size();
        RESERVE(ctx, pl.d_slot_of, slot_of.size() * sizeof(int32_t));
        RESERVE(ctx, pl.d_dummy, (dummy.size() + 1) * sizeof(int32_t));
        CU_TRY(ctx, cudaMemcpyAsync(pl.d_slot_of.p, slot_of.data(), slot_of.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        if (!dummy.empty())
            CU_TRY(ctx, cudaMemcpyAsync(pl.d_dummy.p, dummy.data(), dummy.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->refine_ready = true;
    ctx->ref.ready = false;
    ctx->n_images = 0;
    ctx->have_noise = false;
    ctx->have_ring_w = false;
    return 0;
}

// Forget the loaded images, the whitening curve and the ring weights but keep the band plan and the
// transformed reference: what a resident engine does between two front-end calls that share a reference
// (a fresh process would re-estimate the whitening curve from its own images — so must the next call).
extern "C" int cspb_refine_reset_images(cspb_ctx *ctx) {
    if (!ctx) return CSPB_E_ARG;
    if (!ctx->refine_ready) return cspb_fail(ctx, CSPB_E_STATE, "cspb_refine_configure first");
    ctx->n_images = 0;
    ctx->have_noise = false;
    ctx->have_ring_w = false;
    return 0;
}

// units (warps) of the scoring kernel that are resident at once = one full wave over the GPU
extern "C" int cspb_wave_units(cspb_ctx *ctx) {
    if (!ctx) return CSPB_E_ARG;
    if (ctx->wave_units <= 0) {
        int ctas = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, score_kernel<4, false, false>, 128, 0) != cudaSuccess || ctas <= 0) {
            cudaGetLastError();
            ctas = 6;
        }
        ctx->wave_units = ctx->sm_count * ctas * 4;
    }
    return ctx->wave_units;
}

extern "C" int cspb_band_counts(const cspb_ctx *ctx, int *n_band, int *n_slots) {
    if (!ctx || !ctx->refine_ready) return CSPB_E_STATE;
    if (n_band) *n_band = ctx->plan.n_band;
    if (n_slots) *n_slots = ctx->plan.n_slots;
    return 0;
}

extern "C" int cspb_refine_set_ring_weights(cspb_ctx *ctx, const float *w, int n_rings) {
    CSPB_ENTER(ctx);
    if (!ctx || !ctx->refine_ready) return CSPB_E_STATE;
    if (!w) { ctx->have_ring_w = false; return 0; }
    const int need = ctx->rcfg.box + 1;
    std::vector<float> h(need, 0.f);
    for (int i = 0; i < need; ++i) h[i] = i < n_rings ? w[i] : (n_rings > 0 ? w[n_rings - 1] : 1.f);
    RESERVE(ctx, ctx->d_ring_w, need * sizeof(float));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_ring_w.p, h.data(), need * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_ring_w = true;
    return 0;
}

static bool is_lattice_op(const float *m) {
    for (int k = 0; k < 9; ++k) {
        const float v = m[k], r = roundf(v);
        if (fabsf(v - r) > 1e-4f || fabsf(r) > 1.f) return false;
    }
    return true;
}

extern "C" int cspb_set_symmetry(cspb_ctx *ctx, const float *mats, int n_mats) {
    CSPB_ENTER(ctx);
    if (!ctx || !mats || n_mats < 1) return CSPB_E_ARG;
    if (ctx->recon_ready && ctx->raw_dirty) {
        int rc = recon_flush_deferred(ctx);  // pending inserts belong to the previous group
        if (rc) return rc;
    }
    ctx->sym.assign(mats, mats + 9 * (size_t)n_mats);
    ctx->n_sym = n_mats;
    // split G into lattice-preserving H and right-coset representatives R (g = h r)
    std::vector<int> H;
    for (int g = 0; g < n_mats; ++g)
        if (is_lattice_op(mats + 9 * g)) H.push_back(g);
    std::vector<int> R;
    if (H.size() > 1) {
        for (int g = 0; g < n_mats; ++g) {
            bool covered = false;
            for (int r : R) {
                // q = g * r^T ; covered if q is in H
                float q[9];
                const float *G = mats + 9 * g, *Rm = mats + 9 * r;
                for (int a = 0; a < 3; ++a)
                    for (int c = 0; c < 3; ++c) q[3 * a + c] = G[3 * a] * Rm[3 * c] + G[3 * a + 1] * Rm[3 * c + 1] + G[3 * a + 2] * Rm[3 * c + 2];
                if (!is_lattice_op(q)) continue;
                for (int h : H) {
                    float d = 0.f;
                    for (int k = 0; k < 9; ++k) d = fmaxf(d, fabsf(q[k] - mats[9 * h + k]));
                    if (d < 1e-3f) { covered = true; break; }
                }
                if (covered) break;
            }
            if (!covered) R.push_back(g);
        }
    }
    ctx->sym_lit.clear();
    ctx->sym_lat_t.clear();
    if (H.size() > 1 && H.size() * R.size() == (size_t)n_mats) {
        for (int r : R) ctx->sym_lit.insert(ctx->sym_lit.end(), mats + 9 * r, mats + 9 * r + 9);
        for (int h : H)
            for (int a = 0; a < 3; ++a)
                for (int c = 0; c < 3; ++c) ctx->sym_lat_t.push_back((int)roundf(mats[9 * h + 3 * c + a]));  // transpose
        ctx->n_lit = (int)R.size();
        ctx->n_lat = (int)H.size();
    } else {
        ctx->sym_lit = ctx->sym;
        ctx->n_lit = n_mats;
        ctx->n_lat = 1;
        const int id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        ctx->sym_lat_t.assign(id, id + 9);
    }
    RESERVE(ctx, ctx->d_sym, ctx->sym.size() * sizeof(float));
    RESERVE(ctx, ctx->d_sym_lit, ctx->sym_lit.size() * sizeof(float));
    RESERVE(ctx, ctx->d_sym_lat, ctx->sym_lat_t.size() * sizeof(int));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_sym.p, ctx->sym.data(), ctx->sym.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_sym_lit.p, ctx->sym_lit.data(), ctx->sym_lit.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_sym_lat.p, ctx->sym_lat_t.data(), ctx->sym_lat_t.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->recon_ready && ctx->n_lat > 1) {
        const int np = ctx->rnp, xh = np / 2 + 1;
        const size_t bytes = (size_t)xh * np * np * sizeof(float4);
        for (int h = 0; h < 2; ++h)
            if (ctx->d_raw[h].bytes < bytes) {
                RESERVE(ctx, ctx->d_raw[h], bytes);
                CU_TRY(ctx, cudaMemsetAsync(ctx->d_raw[h].p, 0, bytes, ctx->stream));
            }
    }
    return 0;
}

extern "C" int cspb_set_reference(cspb_ctx *ctx, const float *vol, int n, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !vol) return CSPB_E_ARG;
    if (!ctx->refine_ready) return cspb_fail(ctx, CSPB_E_STATE, "cspb_refine_configure first");
    if (n != ctx->rcfg.box) return cspb_fail(ctx, CSPB_E_ARG, "reference edge %d != box %d", n, ctx->rcfg.box);
    RefVolume &rv = ctx->ref;
    rv.n = n;
    rv.pad = ctx->rcfg.pad;
    rv.np = n * rv.pad;
    const int np = rv.np, xh = np / 2 + 1;
    rv.rc = (int)ceilf(rv.pad * ctx->plan.r_hi) + 2;
    rv.sx = rv.rc + 1;
    rv.sy = 2 * rv.rc + 1;
    const size_t vol_bytes = (size_t)n * n * n * sizeof(float);
    const float *d_vol = vol;
    if (loc == CSPB_HOST) {
        RESERVE(ctx, ctx->d_stage, vol_bytes);
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_stage.p, vol, vol_bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_vol = ctx->d_stage.as<float>();
    }
    RESERVE(ctx, ctx->d_work0, (size_t)np * np * np * sizeof(float));
    RESERVE(ctx, ctx->d_work1, (size_t)xh * np * np * sizeof(float2));
    pad_volume_kernel<<<grid_for((long long)np * np * np, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
        d_vol, ctx->d_work0.as<float>(), n, np);
    KERNEL_CHECK(ctx);
    int rc = fft3_r2c_dev(ctx, ctx->d_work0.as<float>(), ctx->d_work1.as<float2>(), np);
    if (rc) return rc;
    const size_t ref_bytes = (size_t)rv.sx * rv.sy * rv.sy * sizeof(float4);
    RESERVE(ctx, rv.d_ref4, ref_bytes);
    crop_pair_kernel<<<grid_for((long long)rv.sx * rv.sy * rv.sy, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
        ctx->d_work1.as<float2>(), rv.d_ref4.as<float4>(), np, rv.rc, rv.sx, rv.sy);
    KERNEL_CHECK(ctx);
    rv.sx8 = rv.rc + 2;
    RESERVE(ctx, rv.d_ref8, (size_t)rv.sx8 * rv.sy * rv.sy * sizeof(RefQuad));
    crop_quad_kernel<<<grid_for((long long)rv.sx8 * rv.sy * rv.sy, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
        ctx->d_work1.as<float2>(), rv.d_ref8.as<RefQuad>(), np, rv.rc, rv.sx8, rv.sy);
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    rv.ready = true;
    return 0;
}

// FFT a chunk of images held on the device into d_work1 (half spectra), optionally estimating
// the noise curve.  Returns device pointer of the spectra.
static int preprocess_chunk(cspb_ctx *ctx, const float *d_img, int count, float2 **spec_out, const float *fused_filter) {
    const cspb_refine_cfg &c = ctx->rcfg;
    const int n = c.box, nh = n / 2 + 1;
    RESERVE(ctx, ctx->d_stats, (size_t)2 * count * sizeof(float));
    float *offs = ctx->d_stats.as<float>(), *scls = offs + count;
    image_stats_kernel<<<count, 256, 0, ctx->stream>>>(d_img, n, c.mask_radius / c.pixel_size, c.normalize,
                                                       c.invert_contrast, offs, scls);
    KERNEL_CHECK(ctx);
    RESERVE(ctx, ctx->d_work1, (size_t)count * n * nh * sizeof(float2));
    float2 *spec = ctx->d_work1.as<float2>();
    int rc = fft2_r2c_dev(ctx, d_img, spec, n, count, offs, scls, fused_filter);
    if (rc) return rc;
    *spec_out = spec;
    return 0;
}

static int estimate_noise_from_spectra(cspb_ctx *ctx, const float2 *spec, int count) {
    const int n = ctx->rcfg.box, n_rings = n + 1;
    std::vector<int> off, idx, cnt;
    build_ring_csr(n, off, idx, cnt);
    DevBuf d_off, d_idx;
    RESERVE(ctx, d_off, off.size() * sizeof(int));
    RESERVE(ctx, d_idx, idx.size() * sizeof(int));
    CU_TRY(ctx, cudaMemcpyAsync(d_off.p, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(d_idx.p, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    // sample: every step-th of the first min(count, 4096) images, so that the curve does not depend on
    // how the caller chunks the stack (every chunking of the staged and streamed APIs starts with >= 4096
    // images or with the whole stack)
    if (count > 4096) count = 4096;
    const int step = count > 1024 ? count / 1024 : 1;
    const int n_s = (count + step - 1) / step;
    // sample every `step`-th image: gather pointers by launching per sampled image
    RESERVE(ctx, ctx->d_work2, ((size_t)n_s * n_rings + n_rings) * sizeof(float));
    float *per = ctx->d_work2.as<float>();
    ring_power_kernel<<<n_s, 256, 0, ctx->stream>>>(spec, n, step, d_off.as<int>(), d_idx.as<int>(), n_rings, per);
    KERNEL_CHECK(ctx);
    float *tot = per + (long long)n_s * n_rings;
    sum_over_images_kernel<<<ceil_div(n_rings, 128), 128, 0, ctx->stream>>>(per, n_s, n_rings, tot);
    KERNEL_CHECK(ctx);
    std::vector<float> h(n_rings);
    CU_TRY(ctx, cudaMemcpyAsync(h.data(), tot, n_rings * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->noise_curve.assign(n_rings, 0.f);
    for (int r = 0; r < n_rings; ++r)
        ctx->noise_curve[r] = cnt[r] > 0 ? h[r] / ((float)cnt[r] * (float)n_s) : 0.f;
    return cspb_refine_set_noise_curve(ctx, ctx->noise_curve.data(), n_rings);
}

extern "C" int cspb_refine_set_noise_curve(cspb_ctx *ctx, const float *curve, int n_rings) {
    CSPB_ENTER(ctx);
    if (!ctx || !ctx->refine_ready) return CSPB_E_STATE;
    const int need = ctx->rcfg.box + 1;
    if (!curve || n_rings != need) return cspb_fail(ctx, CSPB_E_ARG, "noise curve needs box+1 = %d rings", need);
    ctx->noise_curve.assign(curve, curve + need);
    std::vector<float> filt(need);
    for (int r = 0; r < need; ++r) filt[r] = curve[r] > 0.f ? 1.f / sqrtf(curve[r]) : 0.f;
    RESERVE(ctx, ctx->d_noise, need * sizeof(float));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_noise.p, filt.data(), need * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_noise = true;
    return 0;
}

extern "C" int cspb_refine_get_noise_curve(cspb_ctx *ctx, float *curve_out, int n_rings) {
    CSPB_ENTER(ctx);
    if (!ctx || !ctx->refine_ready || !ctx->have_noise) return CSPB_E_STATE;
    if (!curve_out || n_rings != ctx->rcfg.box + 1) return CSPB_E_ARG;
    memcpy(curve_out, ctx->noise_curve.data(), n_rings * sizeof(float));
    return 0;
}

extern "C" int cspb_refine_num_images(const cspb_ctx *ctx) { return ctx ? ctx->n_images : 0; }

extern "C" int cspb_refine_load_images(cspb_ctx *ctx, const float *images, int n_images, int loc, int append) {
    CSPB_ENTER(ctx);
    if (!ctx || !images || n_images < 0) return CSPB_E_ARG;
    if (!ctx->refine_ready) return cspb_fail(ctx, CSPB_E_STATE, "cspb_refine_configure first");
    const cspb_refine_cfg &c = ctx->rcfg;
    const int n = c.box, nh = n / 2 + 1, n_slots = ctx->plan.n_slots;
    const int base = append ? ctx->n_images : 0;
    const int total = base + n_images;
    // capacity is kept in images of the CURRENT band plan: a reconfiguration (other box or band) changes
    // n_slots, so the byte size is checked as well
    if (total > ctx->img_capacity || (size_t)total * n_slots * sizeof(float2) > ctx->d_packed.bytes) {
        const int cap = append ? (total * 3 / 2 > 1024 ? total * 3 / 2 : 1024) : total;
        DevBuf nb;
        RESERVE(ctx, nb, (size_t)cap * n_slots * sizeof(float2));
        if (base > 0)
            CU_TRY(ctx, cudaMemcpyAsync(nb.p, ctx->d_packed.p, (size_t)base * n_slots * sizeof(float2),
                                        cudaMemcpyDeviceToDevice, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d_packed.release();
        ctx->d_packed.p = nb.p;
        ctx->d_packed.bytes = nb.bytes;
        nb.p = nullptr;
        nb.bytes = 0;
        ctx->img_capacity = cap;
    }
    // cspb_refine_keep_spectra: the plain forward transforms (and their normalisation) stay on the device for the insertion
    const float *span_src = images;
    int span_count = n_images;
    long long span_off = 0;
    if (ctx->keep_span_src && loc == CSPB_DEVICE && images >= ctx->keep_span_src) {
        const size_t d = (size_t)(images - ctx->keep_span_src), per = (size_t)n * n;
        if (d % per == 0 && d / per + (size_t)n_images <= (size_t)ctx->keep_span_count) {
            span_src = ctx->keep_span_src;
            span_count = ctx->keep_span_count;
            span_off = (long long)(d / per);
        }
    }
    const bool continuing = span_off > 0 && ctx->keep_src == span_src && ctx->keep_count == span_off && ctx->keep_box == n;
    if (!continuing) {
        ctx->keep_count = 0;
        ctx->keep_src = nullptr;
    }
    bool keep = ctx->keep_on && loc == CSPB_DEVICE && !append && n_images > 0 && (span_off == 0 || continuing);
    if (keep) {
        const size_t need = (size_t)span_count * n * nh * sizeof(float2);
        if (need > ctx->d_keep_spec.bytes) {  // an optimisation only: never at the price of the memory the caller needs
            size_t free_b = 0, total_b = 0;
            CU_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
            if (continuing || need + ((size_t)8 << 30) > free_b + ctx->d_keep_spec.bytes) keep = false;
        }
        if (keep) {
            RESERVE(ctx, ctx->d_keep_spec, need);
            RESERVE(ctx, ctx->d_keep_stats, (size_t)4 * span_count * sizeof(float));  // offs, scls (refine) | offs, scls (recon)
        }
    }
    // ... and, when the reconstruction is already configured, its normalisation comes out of the same pass over the pixels
    float recon_radius = 0.f;
    bool dual = keep && ctx->recon_ready && ctx->ccfg.box == n && c.normalize && ctx->ccfg.normalize;
    if (dual) {
        recon_radius = ctx->ccfg.mask_radius / ctx->ccfg.pixel_size;
        if (recon_radius > 0.5f * (float)n) recon_radius = 0.5f * (float)n;
        if (continuing && !(ctx->keep_recon_valid && ctx->keep_recon_radius == recon_radius &&
                            (ctx->keep_recon_invert != 0) == (ctx->ccfg.invert_contrast != 0)))
            dual = false;
    }
    const int chunk = chunk_images(n, n_images);
    if (loc == CSPB_HOST && n_images > 0) {
        // host stack: chunk k+1 is copied on the copy stream into the next staging buffer while chunk k is preprocessed
        if (!ctx->pipe_copy) {
            CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->pipe_copy, cudaStreamNonBlocking));
            for (int k = 0; k < CSPB_PIPE_STAGES; ++k) {
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_ready[k], cudaEventDisableTiming));
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_freed[k], cudaEventDisableTiming));
            }
        }
        const int n_buf = n_images > chunk ? CSPB_PIPE_STAGES : 1;
        for (int k = 0; k < n_buf; ++k) RESERVE(ctx, ctx->pipe_stage[k], (size_t)chunk * n * n * sizeof(float));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // whoever used the staging buffers before this call is done
    }
    int chunk_index = 0;
    for (int s = 0; s < n_images; s += chunk, ++chunk_index) {
        const int cnt = n_images - s < chunk ? n_images - s : chunk;
        const float *d_img = images + (size_t)s * n * n;
        const int sb = chunk_index % CSPB_PIPE_STAGES;
        if (loc == CSPB_HOST) {
            if (chunk_index >= CSPB_PIPE_STAGES) CU_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_copy, ctx->pipe_freed[sb], 0));
            CU_TRY(ctx, cudaMemcpyAsync(ctx->pipe_stage[sb].p, d_img, (size_t)cnt * n * n * sizeof(float), cudaMemcpyHostToDevice, ctx->pipe_copy));
            CU_TRY(ctx, cudaEventRecord(ctx->pipe_ready[sb], ctx->pipe_copy));
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_ready[sb], 0));
            d_img = ctx->pipe_stage[sb].as<float>();
        }
        float2 *spec = nullptr;
        // whitening filter and soft mask ride on the FFT passes when the box has the fast path and
        // the noise curve is already known (every chunk but the very first)
        const bool fast = fft_has_fast_path(n);
        const bool fused_filt = fast && c.whiten && ctx->have_noise;
        const char *fuse_env = getenv("CSPB_PREP_FUSED");  // "0": the separate passes (A/B and the bit-identity test)
        const bool no_fuse = fuse_env && fuse_env[0] == '0';
        if (fused_filt && c.apply_mask && !no_fuse) {
            // the common case (every batch but the one the whitening curve is estimated on): four passes instead of seven,
            // the whitening and masking round trips never leave the SM and the band pack rides on the last column pass
            // CSPB_PREP_SUB=<images>: sub-chunks whose half spectra stay in the L2 from one pass to the next — measured SLOWER
            // (profiles/r02_notes.md: the passes are not DRAM bound and every dependent launch costs ~11 us), so the default
            // is one set of launches per chunk
            const char *sub_env = getenv("CSPB_PREP_SUB");
            int sub = sub_env ? atoi(sub_env) : cnt;
            if (sub < 8) sub = 8;
            RESERVE(ctx, ctx->d_stats, (size_t)2 * cnt * sizeof(float));
            RESERVE(ctx, ctx->d_work1, (size_t)(sub < cnt ? sub : cnt) * n * nh * sizeof(float2));
            for (int q = 0; q < cnt; q += sub) {
                const int m = cnt - q < sub ? cnt - q : sub;
                float *offs = ctx->d_stats.as<float>() + 2 * q, *scls = offs + m;
                if (keep) {  // the normalisation goes where the insertion finds it
                    offs = ctx->d_keep_stats.as<float>() + span_off + s + q;
                    scls = offs + span_count;
                }
                const float *img_q = d_img + (size_t)q * n * n;
                if (dual)
                    image_stats_dual_kernel<<<m, 256, 0, ctx->stream>>>(img_q, n, c.mask_radius / c.pixel_size, recon_radius, c.invert_contrast,
                                                                        ctx->ccfg.invert_contrast, offs, scls, offs + 2 * span_count,
                                                                        scls + 2 * span_count);
                else
                    image_stats_kernel<<<m, 256, 0, ctx->stream>>>(img_q, n, c.mask_radius / c.pixel_size, c.normalize, c.invert_contrast, offs, scls);
                KERNEL_CHECK(ctx);
                int rc = fft2_whiten_mask_pack_dev(ctx, img_q, ctx->d_work1.as<float2>(), n, m, offs, scls, ctx->d_noise.as<float>(),
                                                   1.f / ((float)n * (float)n), c.mask_radius / c.pixel_size, 20.f / c.pixel_size,
                                                   ctx->plan.d_slot_of.as<int32_t>(), ctx->have_ring_w ? ctx->d_ring_w.as<float>() : nullptr,
                                                   ctx->plan.d_dummy.as<int32_t>(), ctx->plan.n_dummy,
                                                   ctx->d_packed.as<float2>() + (size_t)(base + s + q) * n_slots, n_slots,
                                                   keep ? ctx->d_keep_spec.as<float2>() + (size_t)(span_off + s + q) * n * nh : nullptr);
                if (rc) return rc;
            }
            if (loc == CSPB_HOST) CU_TRY(ctx, cudaEventRecord(ctx->pipe_freed[sb], ctx->stream));
            continue;
        }
        int rc = preprocess_chunk(ctx, d_img, cnt, &spec, fused_filt ? ctx->d_noise.as<float>() : nullptr);
        if (rc) return rc;
        if (keep && fused_filt) keep = false;  // the whitening rode on the forward pass: no plain transform to keep
        dual = false;                          // separate passes: the insertion computes its own normalisation
        if (keep) {
            CU_TRY(ctx, cudaMemcpyAsync(ctx->d_keep_spec.as<float2>() + (size_t)(span_off + s) * n * nh, spec, (size_t)cnt * n * nh * sizeof(float2),
                                        cudaMemcpyDeviceToDevice, ctx->stream));
            float *ks = ctx->d_keep_stats.as<float>() + span_off + s;
            CU_TRY(ctx, cudaMemcpyAsync(ks, ctx->d_stats.as<float>(), (size_t)cnt * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
            CU_TRY(ctx, cudaMemcpyAsync(ks + span_count, ctx->d_stats.as<float>() + cnt, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToDevice,
                                        ctx->stream));
        }
        if (c.whiten && !ctx->have_noise) {
            rc = estimate_noise_from_spectra(ctx, spec, cnt);
            if (rc) return rc;
        }
        const float *filt = (c.whiten && !fused_filt) ? ctx->d_noise.as<float>() : nullptr;
        const long long tot_c = (long long)cnt * n * nh;
        if (c.apply_mask) {
            if (filt) {
                radial_filter_kernel<<<grid_for(tot_c, 256, ctx->sm_count), 256, 0, ctx->stream>>>(spec, n, cnt, filt);
                KERNEL_CHECK(ctx);
            }
            RESERVE(ctx, ctx->d_work0, (size_t)cnt * n * n * sizeof(float));
            float *real = ctx->d_work0.as<float>();
            const float inv_n2 = 1.f / ((float)n * (float)n);
            if (fast) {
                rc = fft2_c2r_dev(ctx, spec, real, n, cnt, inv_n2, c.mask_radius / c.pixel_size, 20.f / c.pixel_size);
                if (rc) return rc;
            } else {
                rc = fft2_c2r_dev(ctx, spec, real, n, cnt);
                if (rc) return rc;
                mask_kernel<<<grid_for((long long)cnt * n * n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
                    real, n, cnt, c.mask_radius / c.pixel_size, 20.f / c.pixel_size, inv_n2);
                KERNEL_CHECK(ctx);
            }
            rc = fft2_r2c_dev(ctx, real, spec, n, cnt, nullptr, nullptr);
            if (rc) return rc;
            filt = nullptr;
        }
        dim3 grid(ceil_div(n_slots, 256), cnt);
        pack_kernel<<<grid, 256, 0, ctx->stream>>>(spec, n, n_slots, ctx->plan.d_slot_ij.as<int32_t>(), filt,
                                                   ctx->have_ring_w ? ctx->d_ring_w.as<float>() : nullptr,
                                                   ctx->d_packed.as<float2>() + (size_t)(base + s) * n_slots);
        KERNEL_CHECK(ctx);
        if (loc == CSPB_HOST) CU_TRY(ctx, cudaEventRecord(ctx->pipe_freed[sb], ctx->stream));  // staging buffer reuse
    }
    if (loc == CSPB_HOST && n_images > 0) {  // the caller's host buffer is free again, the packed images are complete
        CU_TRY(ctx, cudaStreamSynchronize(ctx->pipe_copy));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->n_images = total;
    if (!keep) {
        ctx->keep_count = 0;
        ctx->keep_src = nullptr;
    } else {
        ctx->keep_src = span_src;
        ctx->keep_count = (int)span_off + n_images;
        ctx->keep_total = span_count;
        ctx->keep_box = n;
        ctx->keep_radius = c.mask_radius / c.pixel_size;
        if (ctx->keep_radius > 0.5f * (float)n) ctx->keep_radius = 0.5f * (float)n;
        ctx->keep_normalize = c.normalize;
        ctx->keep_invert = c.invert_contrast;
        ctx->keep_recon_valid = dual;
        ctx->keep_recon_radius = recon_radius;
        ctx->keep_recon_invert = ctx->ccfg.invert_contrast;
    }
    return 0;
}

extern "C" int cspb_refine_keep_spectra(cspb_ctx *ctx, int on) {
    if (!ctx) return CSPB_E_ARG;
    ctx->keep_on = on != 0;
    if (!on) {
        ctx->keep_count = 0;
        ctx->keep_src = nullptr;
    }
    return 0;
}

int upload_rows(cspb_ctx *ctx, const cspb_row *rows, int n, cspb_row **d_rows, CtfCoef **d_ctf) {
    RESERVE(ctx, ctx->d_rows, (size_t)n * (sizeof(cspb_row) + sizeof(CtfCoef)));
    *d_rows = ctx->d_rows.as<cspb_row>();
    *d_ctf = reinterpret_cast<CtfCoef *>(*d_rows + n);
    CU_TRY(ctx, cudaMemcpyAsync(*d_rows, rows, (size_t)n * sizeof(cspb_row), cudaMemcpyHostToDevice, ctx->stream));
    ctf_coef_kernel<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(*d_rows, n, ctx->rcfg.box, *d_ctf);
    KERNEL_CHECK(ctx);
    return 0;
}

extern "C" int cspb_refine_score_poses(cspb_ctx *ctx, const cspb_row *rows, int n_rows, const int32_t *image_index,
                                       const float *poses6, int n_evals, float *scores_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows || !image_index || !poses6 || !scores_out || n_evals < 0) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (n_rows != ctx->n_images) return cspb_fail(ctx, CSPB_E_ARG, "rows (%d) != loaded images (%d)", n_rows, ctx->n_images);
    for (int e = 0; e < n_evals; ++e)
        if (image_index[e] < 0 || image_index[e] >= n_rows) return cspb_fail(ctx, CSPB_E_ARG, "image index out of range");
    if (n_evals == 0) return 0;
    cspb_row *d_rows;
    CtfCoef *d_ctf;
    int rc = upload_rows(ctx, rows, n_rows, &d_rows, &d_ctf);
    if (rc) return rc;
    std::vector<ScoreUnit> units[5];
    bool ddef = false;
    for (int e = 0; e < n_evals && !ddef; ++e) ddef = poses6[(size_t)e * 6 + 5] != 0.f;
    const int PB = (n_evals >= 2 * n_rows) ? 4 : 1;
    build_units_host(image_index, n_evals, PB, units);
    std::vector<ScoreUnit> flat;
    for (int c = 1; c <= 4; ++c) flat.insert(flat.end(), units[c].begin(), units[c].end());
    const int n_units = (int)flat.size();
    RESERVE(ctx, ctx->d_evals, (size_t)n_evals * 6 * sizeof(float));
    RESERVE(ctx, ctx->d_units, (size_t)n_units * sizeof(ScoreUnit));
    RESERVE(ctx, ctx->d_out, (size_t)n_evals * (sizeof(float4) + sizeof(float)));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_evals.p, poses6, (size_t)n_evals * 6 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_units.p, flat.data(), (size_t)n_units * sizeof(ScoreUnit), cudaMemcpyHostToDevice, ctx->stream));
    float4 *d_out = ctx->d_out.as<float4>();
    float *d_sc = reinterpret_cast<float *>(d_out + n_evals);
    size_t off = 0;
    for (int c = 1; c <= 4; ++c) {
        rc = launch_score(ctx, ctx->d_units.as<ScoreUnit>() + off, (int)units[c].size(), c, ctx->d_evals.as<float>(), d_ctf, d_out, ddef,
                          (int64_t)units[c].size() * c, 0);
        if (rc) return rc;
        off += units[c].size();
    }
    scores_from_out_kernel<<<ceil_div(n_evals, 256), 256, 0, ctx->stream>>>(d_out, n_evals, d_sc);
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaMemcpyAsync(scores_out, d_sc, (size_t)n_evals * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cspb_refine_score(cspb_ctx *ctx, const cspb_row *rows, int n, float *scores_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows || !scores_out || n < 0) return CSPB_E_ARG;
    std::vector<int32_t> idx(n);
    std::vector<float> poses((size_t)n * 6);
    for (int k = 0; k < n; ++k) {
        idx[k] = k;
        float *q = &poses[(size_t)k * 6];
        q[0] = rows[k].psi; q[1] = rows[k].theta; q[2] = rows[k].phi;
        q[3] = rows[k].x_shift; q[4] = rows[k].y_shift; q[5] = 0.f;
    }
    return cspb_refine_score_poses(ctx, rows, n, idx.data(), poses.data(), n, scores_out);
}

// building block of the analytic optimiser (oracle/SEMANTICS.md §7c): value and derivatives at the poses of the rows;
// out = 28 floats per row {num, X, A, B, d num / d (psi, theta, phi [deg], x, y [A]), d B / d angles, J^T J upper triangle, 0}
extern "C" int cspb_refine_score_grad(cspb_ctx *ctx, const cspb_row *rows, int n, int ring_cut, float *out28) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows || !out28 || n < 0) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (n != ctx->n_images) return cspb_fail(ctx, CSPB_E_ARG, "rows (%d) != loaded images (%d)", n, ctx->n_images);
    if (n == 0) return 0;
    cspb_row *d_rows;
    CtfCoef *d_ctf;
    int rc = upload_rows(ctx, rows, n, &d_rows, &d_ctf);
    if (rc) return rc;
    std::vector<float> poses((size_t)n * 6);
    std::vector<ScoreUnit> units(n);
    for (int k = 0; k < n; ++k) {
        float *q = &poses[(size_t)k * 6];
        q[0] = rows[k].psi; q[1] = rows[k].theta; q[2] = rows[k].phi; q[3] = rows[k].x_shift; q[4] = rows[k].y_shift; q[5] = 0.f;
        units[k].image = k; units[k].first_eval = k; units[k].count = 1; units[k].pad_ = 0;
    }
    RESERVE(ctx, ctx->d_evals, poses.size() * sizeof(float));
    RESERVE(ctx, ctx->d_units, units.size() * sizeof(ScoreUnit));
    RESERVE(ctx, ctx->d_out, (size_t)n * CSPB_GOUT * sizeof(float));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_evals.p, poses.data(), poses.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_units.p, units.data(), units.size() * sizeof(ScoreUnit), cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_score_grad(ctx, ctx->d_units.as<ScoreUnit>(), n, ctx->d_evals.as<float>(), d_ctf, ctx->d_out.as<float>(),
                           ring_cut > 0 ? ring_cut : 0x7fffffff);
    if (rc) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(out28, ctx->d_out.p, (size_t)n * CSPB_GOUT * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ================================================================== 2-D focus mask (prompts 29-32, 44)
// LOGP over the projected focus sphere (oracle/SEMANTICS.md 6b): residual spectrum (shifted image -
// alpha CTF slice) on the scoring band -> inverse FFT -> variance inside the disc.
__global__ void focus_residual_kernel(const float4 *__restrict__ ref4, int sx, int sy, int rc, float padf,
                                      const int32_t *__restrict__ slot_ij, int n_slots, const float2 *__restrict__ packed,
                                      const CtfCoef *__restrict__ ctf, const cspb_row *__restrict__ rows,
                                      const float *__restrict__ alpha, int n, float inv_npx2, float2 *__restrict__ spec) {
    __shared__ float m[9];
    const int img = blockIdx.y, nh = n / 2 + 1;
    const cspb_row row = rows[img];
    if (threadIdx.x == 0) euler_matrix(row.psi, row.theta, row.phi, m);
    __syncthreads();
    const CtfCoef cc = ctf[img];
    const float a = alpha[img], mx = row.x_shift * inv_npx2, my = row.y_shift * inv_npx2;
    const float2 *F = packed + (long long)img * n_slots;
    float2 *o = spec + (long long)img * n * nh;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
        const int32_t ij = slot_ij[s];
        const int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
        if (i == CSPB_DUMMY_I) continue;
        const float fi = (float)i, fj = (float)j;
        const float2 P = gather_trilinear(ref4, sx, sy, rc, (m[0] * fi + m[1] * fj) * padf, (m[3] * fi + m[4] * fj) * padf,
                                          (m[6] * fi + m[7] * fj) * padf);
        const float cv = -sinpif(ctf_chi(cc, fi, fj, fi * fi + fj * fj) * (1.f / CSPB_PI_F));
        float sn, cs;
        sincospif(fi * mx + fj * my, &sn, &cs);
        const float2 f = F[s];
        const float gr = f.x * cs - f.y * sn, gi = f.x * sn + f.y * cs;
        const float w = ((i + j) & 1) ? -1.f : 1.f;  // centred phase origin -> image corner
        const int jj = j < 0 ? j + n : j;
        o[jj * nh + i] = make_float2(w * (gr - a * cv * P.x), w * (gi - a * cv * P.y));
    }
}

// one CTA per image: variance of the residual inside the projected disc -> LOGP
__global__ void focus_logp_kernel(const float *__restrict__ real, int n, float pixel, float fx, float fy, float fz, float frad,
                                  cspb_row *__restrict__ rows, cspb_row *__restrict__ changes) {
    __shared__ float s_sum[32];
    __shared__ int s_cnt[32];
    const int img = blockIdx.x;
    const cspb_row row = rows[img];
    float m[9];
    euler_matrix(row.psi, row.theta, row.phi, m);
    const float h = (float)(n / 2);
    const float x = fx / pixel - h, y = fy / pixel - h, z = fz / pixel - h;
    const float cx = h + m[0] * x + m[3] * y + m[6] * z, cy = h + m[1] * x + m[4] * y + m[7] * z;
    const float rad = frad / pixel;
    const int x0 = max(0, (int)floorf(cx - rad)), x1 = min(n - 1, (int)ceilf(cx + rad));
    const int y0 = max(0, (int)floorf(cy - rad)), y1 = min(n - 1, (int)ceilf(cy + rad));
    const int wbox = x1 - x0 + 1, hbox = y1 - y0 + 1;
    const float *p = real + (long long)img * n * n;
    float ss = 0.f;
    int cnt = 0;
    if (wbox > 0 && hbox > 0)
        for (int t = threadIdx.x; t < wbox * hbox; t += blockDim.x) {
            const int px = x0 + t % wbox, py = y0 + t / wbox;
            const float dx = (float)px - cx, dy = (float)py - cy;
            if (dx * dx + dy * dy > rad * rad) continue;
            const float v = p[py * n + px];
            ss += v * v;
            ++cnt;
        }
    ss = warp_sum(ss);
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_sum[warp] = ss; s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        int c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tot += s_sum[w]; c += s_cnt[w]; }
        float logp = 0.f;
        if (c > 0) {
            const float var = tot / (float)c;
            if (var > 0.f) logp = -0.5f * (float)c * (1.f + logf(2.f * CSPB_PI_F * var));
        }
        if (changes) changes[img].logp += logp - row.logp;
        rows[img].logp = logp;
    }
}

// matching projections (refine3d answers 8 / 43, frealign.py:3929-3931): the CTF-multiplied central slice at the pose of
// every row, moved by the row's shift into the frame of the particle image, on the half plane up to the scoring limit
__global__ void match_spectrum_kernel(const float4 *__restrict__ ref4, int sx, int sy, int rc, float padf, const CtfCoef *__restrict__ ctf,
                                      const cspb_row *__restrict__ rows, int n, float r_hi, float inv_npx2, float2 *__restrict__ spec) {
    __shared__ float m[9];
    const int img = blockIdx.y, nh = n / 2 + 1;
    const cspb_row row = rows[img];
    if (threadIdx.x == 0) euler_matrix(row.psi, row.theta, row.phi, m);
    __syncthreads();
    const CtfCoef cc = ctf[img];
    const float mx = row.x_shift * inv_npx2, my = row.y_shift * inv_npx2;
    float2 *o = spec + (long long)img * n * nh;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * nh; idx += gridDim.x * blockDim.x) {
        const int i = idx % nh;
        int j = idx / nh;
        if (j >= n / 2) j -= n;
        float2 v = make_float2(0.f, 0.f);
        const float fi = (float)i, fj = (float)j, r2 = fi * fi + fj * fj;
        if (r2 <= r_hi * r_hi && j != -n / 2) {
            const float2 P = gather_trilinear(ref4, sx, sy, rc, (m[0] * fi + m[1] * fj) * padf, (m[3] * fi + m[4] * fj) * padf,
                                              (m[6] * fi + m[7] * fj) * padf);
            const float cv = -sinpif(ctf_chi(cc, fi, fj, r2) * (1.f / CSPB_PI_F));
            float sn, cs;
            sincospif(fi * mx + fj * my, &sn, &cs);   // the image is moved by +shift onto the reference: the projection goes back by -shift
            const float w = ((i + j) & 1) ? -cv : cv;  // centred phase origin -> image corner
            v = make_float2(w * (P.x * cs + P.y * sn), w * (P.y * cs - P.x * sn));
        }
        o[idx] = v;
    }
}

extern "C" int cspb_refine_matching_projections(cspb_ctx *ctx, const cspb_row *rows, int n_rows, float *out_images) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows || !out_images || n_rows < 0) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (n_rows == 0) return 0;
    const cspb_refine_cfg &c = ctx->rcfg;
    const int n = c.box, nh = n / 2 + 1;
    int chunk = chunk_images(n, n_rows);
    if (chunk > 2048) chunk = 2048;
    RESERVE(ctx, ctx->d_work0, (size_t)chunk * n * n * sizeof(float));
    RESERVE(ctx, ctx->d_work1, (size_t)chunk * n * nh * sizeof(float2));
    for (int s = 0; s < n_rows; s += chunk) {
        const int cnt = n_rows - s < chunk ? n_rows - s : chunk;
        cspb_row *d_rows;
        CtfCoef *d_ctf;
        int rc = upload_rows(ctx, rows + s, cnt, &d_rows, &d_ctf);
        if (rc) return rc;
        dim3 grid(ceil_div(n * nh, 256) < 64 ? ceil_div(n * nh, 256) : 64, cnt);
        match_spectrum_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->ref.d_ref4.as<float4>(), ctx->ref.sx, ctx->ref.sy, ctx->ref.rc, (float)ctx->ref.pad,
                                                             d_ctf, d_rows, n, ctx->plan.r_hi, 2.f / ((float)n * c.pixel_size), ctx->d_work1.as<float2>());
        KERNEL_CHECK(ctx);
        rc = fft2_c2r_dev(ctx, ctx->d_work1.as<float2>(), ctx->d_work0.as<float>(), n, cnt, 1.f / ((float)n * (float)n));
        if (rc) return rc;
        CU_TRY(ctx, cudaMemcpyAsync(out_images + (size_t)s * n * n, ctx->d_work0.p, (size_t)cnt * n * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

static int focus_logp_enqueue(cspb_ctx *ctx, cspb_row *d_rows, CtfCoef *d_ctf, int n_img, cspb_row *d_changes) {
    const cspb_refine_cfg &c = ctx->rcfg;
    const int n = c.box, nh = n / 2 + 1, n_slots = ctx->plan.n_slots;
    // the rows hold the refined defocus: refresh the CTF coefficients
    ctf_coef_kernel<<<ceil_div(n_img, 128), 128, 0, ctx->stream>>>(d_rows, n_img, n, d_ctf);
    KERNEL_CHECK(ctx);
    int chunk = chunk_images(n, n_img);
    if (chunk > 2048) chunk = 2048;
    RESERVE(ctx, ctx->d_work0, (size_t)chunk * n * n * sizeof(float));
    RESERVE(ctx, ctx->d_work1, (size_t)chunk * n * nh * sizeof(float2));
    float *real = ctx->d_work0.as<float>();
    float2 *spec = ctx->d_work1.as<float2>();
    const float inv_npx2 = 2.f / ((float)n * c.pixel_size);
    for (int s = 0; s < n_img; s += chunk) {
        const int cnt = n_img - s < chunk ? n_img - s : chunk;
        CU_TRY(ctx, cudaMemsetAsync(spec, 0, (size_t)cnt * n * nh * sizeof(float2), ctx->stream));
        dim3 grid(ceil_div(n_slots, 256), cnt);
        focus_residual_kernel<<<grid, 256, 0, ctx->stream>>>(
            ctx->ref.d_ref4.as<float4>(), ctx->ref.sx, ctx->ref.sy, ctx->ref.rc, (float)ctx->ref.pad,
            ctx->plan.d_slot_ij.as<int32_t>(), n_slots, ctx->d_packed.as<float2>() + (size_t)s * n_slots, d_ctf + s, d_rows + s,
            ctx->d_alpha.as<float>() + s, n, inv_npx2, spec);
        KERNEL_CHECK(ctx);
        int rc = fft2_c2r_dev(ctx, spec, real, n, cnt, 1.f / ((float)n * (float)n));
        if (rc) return rc;
        focus_logp_kernel<<<cnt, 256, 0, ctx->stream>>>(real, n, c.pixel_size, ctx->focus[0], ctx->focus[1], ctx->focus[2],
                                                         ctx->focus[3], d_rows + s, d_changes ? d_changes + s : nullptr);
        KERNEL_CHECK(ctx);
    }
    return 0;
}

extern "C" int cspb_refine_set_focus_mask(cspb_ctx *ctx, float x, float y, float z, float radius) {
    if (!ctx) return CSPB_E_ARG;
    ctx->focus[0] = x; ctx->focus[1] = y; ctx->focus[2] = z; ctx->focus[3] = radius > 0.f ? radius : 0.f;
    return 0;
}

// ================================================================== beam-tilt phase sum (refine_ctf answer 23)
// S(i,j) = sum over the images of G * conj(CTF * slice) on the scoring band (oracle/SEMANTICS.md §12): one
// thread per band slot walks a strided subset of the images (coalesced reads of the packed spectra,
// one gather per image — the cost of one score evaluation per particle); NG image groups run in
// parallel and are reduced in a fixed order (deterministic).
__global__ void pose_matrix_kernel(const cspb_row *__restrict__ rows, int n, float inv_npx2, float *__restrict__ m8) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float r[9];
    euler_matrix(rows[k].psi, rows[k].theta, rows[k].phi, r);
    float *o = m8 + (long long)k * 8;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[3]; o[3] = r[4]; o[4] = r[6]; o[5] = r[7];
    o[6] = rows[k].x_shift * inv_npx2; o[7] = rows[k].y_shift * inv_npx2;
}

__global__ void phase_sum_kernel(const float4 *__restrict__ ref4, int sx, int sy, int rc, float padf,
                                 const int32_t *__restrict__ slot_ij, int n_slots, const float2 *__restrict__ packed,
                                 const CtfCoef *__restrict__ ctf, const float *__restrict__ m8, int n_img, int NG,
                                 float2 *__restrict__ partial) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int32_t ij = slot_ij[s];
    const int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
    float2 acc = make_float2(0.f, 0.f);
    if (i != CSPB_DUMMY_I) {
        const float fi = (float)i, fj = (float)j, r2 = fi * fi + fj * fj;
        for (int img = blockIdx.y; img < n_img; img += NG) {
            const float4 ma = __ldg(reinterpret_cast<const float4 *>(m8 + (long long)img * 8));
            const float4 mb = __ldg(reinterpret_cast<const float4 *>(m8 + (long long)img * 8 + 4));
            const float2 P = gather_trilinear(ref4, sx, sy, rc, (ma.x * fi + ma.y * fj) * padf, (ma.z * fi + ma.w * fj) * padf,
                                              (mb.x * fi + mb.y * fj) * padf);
            const float cv = -sinpif(ctf_chi(ctf[img], fi, fj, r2) * (1.f / CSPB_PI_F));
            float sn, cs;
            sincospif(fi * mb.z + fj * mb.w, &sn, &cs);
            const float2 f = __ldcs(packed + (long long)img * n_slots + s);
            const float gr = f.x * cs - f.y * sn, gi = f.x * sn + f.y * cs;
            acc.x += cv * (gr * P.x + gi * P.y);
            acc.y += cv * (gi * P.x - gr * P.y);
        }
    }
    partial[(long long)blockIdx.y * n_slots + s] = acc;
}

__global__ void phase_sum_scatter_kernel(const float2 *__restrict__ partial, int NG, const int32_t *__restrict__ slot_ij, int n_slots,
                                         int n, float2 *__restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const int32_t ij = slot_ij[s];
    const int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
    if (i == CSPB_DUMMY_I) return;
    float2 t = make_float2(0.f, 0.f);
    for (int g = 0; g < NG; ++g) {
        const float2 v = partial[(long long)g * n_slots + s];
        t.x += v.x; t.y += v.y;
    }
    const int jj = j < 0 ? j + n : j;
    out[jj * (n / 2 + 1) + i] = t;
}

extern "C" int cspb_refine_phase_sum(cspb_ctx *ctx, const cspb_row *rows, int n_rows, float *out_complex) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows || !out_complex || n_rows < 0) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (n_rows != ctx->n_images) return cspb_fail(ctx, CSPB_E_ARG, "rows (%d) != loaded images (%d)", n_rows, ctx->n_images);
    const cspb_refine_cfg &c = ctx->rcfg;
    const int n = c.box, nh = n / 2 + 1, n_slots = ctx->plan.n_slots;
    const size_t out_bytes = (size_t)n * nh * sizeof(float2);
    if (n_rows == 0) { memset(out_complex, 0, out_bytes); return 0; }
    cspb_row *d_rows;
    CtfCoef *d_ctf;
    int rc = upload_rows(ctx, rows, n_rows, &d_rows, &d_ctf);
    if (rc) return rc;
    int NG = (8 * ctx->sm_count) / ceil_div(n_slots, 256);
    if (NG < 1) NG = 1;
    if (NG > n_rows) NG = n_rows;
    RESERVE(ctx, ctx->d_evals, (size_t)n_rows * 8 * sizeof(float));
    RESERVE(ctx, ctx->d_out, (size_t)NG * n_slots * sizeof(float2));
    RESERVE(ctx, ctx->d_work2, out_bytes);
    float *m8 = ctx->d_evals.as<float>();
    pose_matrix_kernel<<<ceil_div(n_rows, 128), 128, 0, ctx->stream>>>(d_rows, n_rows, 2.f / ((float)n * c.pixel_size), m8);
    KERNEL_CHECK(ctx);
    dim3 grid(ceil_div(n_slots, 256), NG);
    phase_sum_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->ref.d_ref4.as<float4>(), ctx->ref.sx, ctx->ref.sy, ctx->ref.rc,
                                                    (float)ctx->ref.pad, ctx->plan.d_slot_ij.as<int32_t>(), n_slots,
                                                    ctx->d_packed.as<float2>(), d_ctf, m8, n_rows, NG, ctx->d_out.as<float2>());
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaMemsetAsync(ctx->d_work2.p, 0, out_bytes, ctx->stream));
    phase_sum_scatter_kernel<<<ceil_div(n_slots, 256), 256, 0, ctx->stream>>>(ctx->d_out.as<float2>(), NG,
                                                                             ctx->plan.d_slot_ij.as<int32_t>(), n_slots, n,
                                                                             ctx->d_work2.as<float2>());
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaMemcpyAsync(out_complex, ctx->d_work2.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int search_enqueue(cspb_ctx *ctx, const CtfCoef *d_ctf, const float *d_angles3, int n_orient, int K, void *d_hits);

// resolution stage of iteration `it` of the analytic optimiser (same code as lm_stage in oracle/cspb_oracle.c):
// returns f >= 1 and the last ring scored (INT_MAX for the full band)
static float lm_stage(int it, int iters, float r_lo, float r_hi, int *ring_cut) {
    const int ramp = iters - 3;
    float f = 1.f;
    if (ramp > 0 && it < ramp) f = powf(6.f, 1.f - (float)it / (float)ramp);
    *ring_cut = 0x7fffffff;
    if (f > 1.f) {
        int rc = (int)floorf(r_hi / f);
        int least = (int)floorf(r_lo) + 4;
        if (least < 12) least = 12;
        if (rc < least) rc = least;
        if ((float)rc >= r_hi) { f = 1.f; } else { *ring_cut = rc; f = r_hi / (float)rc; }
    }
    return f;
}

// local refinement with the analytic optimiser (oracle/SEMANTICS.md §7c): per iteration ONE gradient evaluation
// (score_grad_kernel) and ONE trial evaluation per state, coarse to fine over the rings of the band
// With hits (global search: K states per image, state k starts at candidate k % K of image k / K) every hit is refined and
// opt_write_rows_kernel keeps the best; the first, coarse stages (rings below r_hi / 6, trust region 16 stencil steps of
// that resolution) span the half grid step a hit may sit away from its optimum.
static int refine_lm_enqueue(cspb_ctx *ctx, cspb_row *d_rows, const CtfCoef *d_ctf, int n_img, cspb_row *d_changes, int free_mask,
                             int64_t *n_evals_out, int K = 1, const SearchHit *d_hits = nullptr, const float *d_angles = nullptr,
                             int64_t evals_before = 0) {
    const cspb_refine_cfg &c = ctx->rcfg;
    const int iters = (free_mask & 31) ? (c.local_iterations > 0 ? c.local_iterations : 8) : 0;
    const int n = n_img * K;  // optimiser states
    RESERVE(ctx, ctx->d_opt, (size_t)n * sizeof(OptState));
    RESERVE(ctx, ctx->d_evals, (size_t)n * 2 * 6 * sizeof(float));
    RESERVE(ctx, ctx->d_units, (size_t)n * sizeof(ScoreUnit));
    RESERVE(ctx, ctx->d_out, (size_t)n * (2 * sizeof(float4) + CSPB_GOUT * sizeof(float)));
    OptState *st = ctx->d_opt.as<OptState>();
    float *ev = ctx->d_evals.as<float>();
    ScoreUnit *un = ctx->d_units.as<ScoreUnit>();
    float4 *out = ctx->d_out.as<float4>();
    float *gout = reinterpret_cast<float *>(out + (size_t)2 * n);
    const float r_hi = ctx->plan.r_hi, r_lo = ctx->plan.r_lo;
    const int g = ceil_div(n, 128);
    const bool focus_on = ctx->focus[3] > 0.f;
    if (focus_on) RESERVE(ctx, ctx->d_alpha, (size_t)n_img * sizeof(float));
    OptPrior pr{};
    float lam_scale = 0.f;
    if (c.use_priors) {
        const float rad = c.mask_radius / c.pixel_size;
        const float nmask = fmaxf(3.14159265f * rad * rad, 1.f);
        lam_scale = 1.f / nmask;
        pr.on = 1;
        pr.mx = c.prior_mean_x; pr.my = c.prior_mean_y;
        pr.wx = c.prior_var_x > 0.f ? 0.5f / c.prior_var_x : 0.f;
        pr.wy = c.prior_var_y > 0.f ? 0.5f / c.prior_var_y : 0.f;
    }
    opt_init_kernel<<<g, 128, 0, ctx->stream>>>(d_rows, n, K, d_hits, d_angles, st, 0.f, 0.f, 0.f, lam_scale);
    KERNEL_CHECK(ctx);
    int64_t evals = evals_before;
    for (int it = 0; it < iters; ++it) {
        int ring_cut;
        const float f = lm_stage(it, iters, r_lo, r_hi, &ring_cut);
        const float h_ang = 0.35f * 57.29578f * f / r_hi, h_shift = 0.07f * (float)c.box * f / r_hi * c.pixel_size;
        lm_pose_kernel<<<g, 128, 0, ctx->stream>>>(st, n, K, ev, un);
        KERNEL_CHECK(ctx);
        int rc = launch_score_grad(ctx, un, n, ev, d_ctf, gout, ring_cut);
        if (rc) return rc;
        lm_step_kernel<<<g, 128, 0, ctx->stream>>>(st, n, K, free_mask, gout, 16.f * h_ang, 16.f * h_shift, ev, un, pr);
        KERNEL_CHECK(ctx);
        rc = launch_score(ctx, un, n, 1, ev, d_ctf, out, false, n, 0, ring_cut);
        if (rc) return rc;
        lm_select_kernel<<<g, 128, 0, ctx->stream>>>(st, n, out, pr);
        KERNEL_CHECK(ctx);
        evals += 2 * (int64_t)n;
    }
    opt_finish_eval_kernel<<<g, 128, 0, ctx->stream>>>(st, n, K, ev, un);
    KERNEL_CHECK(ctx);
    int rc = launch_score(ctx, un, n, 2, ev, d_ctf, out, false, 2 * (int64_t)n, 0);
    if (rc) return rc;
    evals += 2 * (int64_t)n;
    opt_write_rows_kernel<<<ceil_div(n_img, 128), 128, 0, ctx->stream>>>(st, n_img, K, out, ctx->plan.n_band, 0, d_rows, d_changes, pr,
                                                                         focus_on ? ctx->d_alpha.as<float>() : nullptr);
    KERNEL_CHECK(ctx);
    if (focus_on) {
        rc = focus_logp_enqueue(ctx, d_rows, const_cast<CtfCoef *>(d_ctf), n_img, d_changes);
        if (rc) return rc;
    }
    if (n_evals_out) *n_evals_out = evals;
    return 0;
}

// enqueue the whole refinement on the stream; rows/ctf already on the device.  With global search
// the grid is searched first and the K best hits per image are refined as separate states.
static int refine_local_enqueue(cspb_ctx *ctx, cspb_row *d_rows, const CtfCoef *d_ctf, int n, cspb_row *d_changes,
                                int64_t *n_evals_out) {
    const cspb_refine_cfg &c = ctx->rcfg;
    int free_mask = 0;
    if (c.refine_psi) free_mask |= 1;
    if (c.refine_theta) free_mask |= 2;
    if (c.refine_phi) free_mask |= 4;
    if (c.refine_x) free_mask |= 8;
    if (c.refine_y) free_mask |= 16;
    if (c.refine_defocus) free_mask |= 32;
    // optimiser 0 (default): analytic gradient + Gauss-Newton step for the five pose parameters — plain local refinement
    // here, the hits of a global search below; the stencil optimiser serves the defocus refinement and optimizer = 1
    if (c.optimizer == 0 && !c.global_search && !c.refine_defocus && c.local_refine)
        return refine_lm_enqueue(ctx, d_rows, d_ctf, n, d_changes, free_mask, n_evals_out);
    int64_t evals = 0;
    int K = 1;
    const SearchHit *d_hits = nullptr;
    const float *d_angles = nullptr;
    if (c.global_search) {
        if (ctx->n_grid <= 0) return cspb_fail(ctx, CSPB_E_STATE, "global search needs cspb_refine_set_search_grid first");
        K = c.best_matches > 0 ? c.best_matches : 20;
        if (K > ctx->n_grid) K = ctx->n_grid;
        RESERVE(ctx, ctx->d_hits, (size_t)n * K * sizeof(SearchHit));
        d_angles = ctx->d_grid.as<float>();
        int rc = search_enqueue(ctx, d_ctf, d_angles, ctx->n_grid, K, ctx->d_hits.p);
        if (rc) return rc;
        d_hits = ctx->d_hits.as<SearchHit>();
        evals += (int64_t)n * ctx->n_grid;
        free_mask |= 31;  // the hits are refined in all five pose parameters
        if (c.optimizer == 0 && !c.refine_defocus)
            return refine_lm_enqueue(ctx, d_rows, d_ctf, n, d_changes, free_mask, n_evals_out, K, d_hits, d_angles, evals);
    }
    int n_free = 0;
    for (int m = 0; m < OPT_NP; ++m) n_free += (free_mask >> m) & 1;
    const int ns = n * K;  // optimiser states
    const int NE = 1 + 2 * n_free, PB = 4;
    const int shift_mask = free_mask & 24;  // x, y: pure image shifts, scored from the centre's gather
    const int nS = 2 * (((shift_mask >> 3) & 1) + ((shift_mask >> 4) & 1));
    const int upi = (NE + PB - 1) / PB + 1;
    const bool ddef = c.refine_defocus != 0;
    const bool do_local = (c.local_refine || c.global_search) && n_free > 0;
    const int iters = do_local ? (c.local_iterations > 0 ? c.local_iterations : 8) : 0;
    const int late = iters / 2 + 1;  // stencil steps stay constant for the first half, then shrink
    RESERVE(ctx, ctx->d_opt, (size_t)ns * sizeof(OptState));
    RESERVE(ctx, ctx->d_evals, (size_t)ns * (NE + OPT_NL + 2) * 6 * sizeof(float));
    RESERVE(ctx, ctx->d_units, (size_t)ns * (upi + 1) * sizeof(ScoreUnit));
    RESERVE(ctx, ctx->d_out, (size_t)ns * (NE + OPT_NL + 2) * sizeof(float4));
    OptState *st = ctx->d_opt.as<OptState>();
    float *ev = ctx->d_evals.as<float>(), *ev_ls = ev + (size_t)ns * NE * 6;
    ScoreUnit *un = ctx->d_units.as<ScoreUnit>(), *un_ls = un + (size_t)ns * upi;
    float4 *out = ctx->d_out.as<float4>(), *out_ls = out + (size_t)ns * NE;
    const float r_hi = ctx->plan.r_hi;
    // hits of the global search sit up to half a grid step from the optimum: their pose steps start
    // r_hi / r_search times larger, so that the first stencils span the search resolution
    float coarse = 1.f;
    if (c.global_search) {
        const float npx = (float)c.box * c.pixel_size;
        float r_s = c.search_high_res > 0.f ? npx / c.search_high_res : r_hi;
        if (r_s > r_hi) r_s = r_hi;
        if (r_s < ctx->plan.r_lo + 2.f) r_s = fminf(ctx->plan.r_lo + 2.f, r_hi);
        coarse = r_hi / r_s;
    }
    const float h_ang = coarse * 0.35f * 57.29578f / r_hi;                      // ~1/3 of the angular resolution at r_hi
    const float h_shift = coarse * 0.07f * (float)c.box / r_hi * c.pixel_size;  // Angstrom
    const float h_def = c.defocus_step > 0.f ? c.defocus_step : 50.f;
    const int g = ceil_div(ns, 128);
    const char *ss_env = getenv("CSPB_SAMESHIFT");  // =0: score the A class with the general kernel (A/B measurements)
    const bool same_shift = !(ss_env && ss_env[0] == '0');
    const bool focus_on = ctx->focus[3] > 0.f;
    if (focus_on) RESERVE(ctx, ctx->d_alpha, (size_t)n * sizeof(float));
    OptPrior pr{};
    float lam_scale = 0.f;
    if (c.use_priors) {
        const float rad = c.mask_radius / c.pixel_size;
        const float nmask = fmaxf(3.14159265f * rad * rad, 1.f);
        lam_scale = 1.f / nmask;
        pr.on = 1;
        pr.mx = c.prior_mean_x; pr.my = c.prior_mean_y;
        pr.wx = c.prior_var_x > 0.f ? 0.5f / c.prior_var_x : 0.f;
        pr.wy = c.prior_var_y > 0.f ? 0.5f / c.prior_var_y : 0.f;
    }
    opt_init_kernel<<<g, 128, 0, ctx->stream>>>(d_rows, ns, K, d_hits, d_angles, st, h_ang, h_shift, h_def, lam_scale);
    KERNEL_CHECK(ctx);
    for (int it = 0; it < iters; ++it) {
        opt_stencil_kernel<<<g, 128, 0, ctx->stream>>>(st, ns, K, free_mask, shift_mask, NE, PB, ev, un);
        KERNEL_CHECK(ctx);
        // the A class (centre, +-angles, +-defocus) keeps the centre's shift: the image is shifted once per sample
        int rc = launch_score_classes(ctx, un, ns, NE - nS, nS, PB, ev, d_ctf, out, ddef, same_shift);
        if (rc) return rc;
        opt_step_kernel<<<g, 128, 0, ctx->stream>>>(st, ns, K, free_mask, NE, out, ev_ls, un_ls, pr);
        KERNEL_CHECK(ctx);
        rc = launch_score(ctx, un_ls, ns, OPT_NL, ev_ls, d_ctf, out_ls, ddef, (int64_t)ns * OPT_NL, 0);
        if (rc) return rc;
        opt_select_kernel<<<g, 128, 0, ctx->stream>>>(st, ns, out_ls, it + 1 >= late ? 0.6f : 1.f, pr);
        KERNEL_CHECK(ctx);
        evals += (int64_t)ns * (NE + OPT_NL);
    }
    opt_finish_eval_kernel<<<g, 128, 0, ctx->stream>>>(st, ns, K, ev, un);
    KERNEL_CHECK(ctx);
    int rc = launch_score(ctx, un, ns, 2, ev, d_ctf, out, ddef, 2 * (int64_t)ns, 0);
    if (rc) return rc;
    evals += 2 * (int64_t)ns;
    opt_write_rows_kernel<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(st, n, K, out, ctx->plan.n_band, c.refine_defocus, d_rows, d_changes, pr,
                                                                         focus_on ? ctx->d_alpha.as<float>() : nullptr);
    KERNEL_CHECK(ctx);
    if (focus_on) {
        rc = focus_logp_enqueue(ctx, d_rows, const_cast<CtfCoef *>(d_ctf), n, d_changes);
        if (rc) return rc;
    }
    if (n_evals_out) *n_evals_out = evals;
    return 0;
}

extern "C" int cspb_refine_set_search_grid(cspb_ctx *ctx, const float *angles3, int n_orient) {
    CSPB_ENTER(ctx);
    if (!ctx || (!angles3 && n_orient > 0) || n_orient < 0) return CSPB_E_ARG;
    ctx->n_grid = n_orient;
    if (n_orient == 0) return 0;
    RESERVE(ctx, ctx->d_grid, (size_t)n_orient * 3 * sizeof(float));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_grid.p, angles3, (size_t)n_orient * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cspb_refine_run_device(cspb_ctx *ctx, cspb_row *rows_dev, int n, int64_t *n_evals_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows_dev) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (n != ctx->n_images) return cspb_fail(ctx, CSPB_E_ARG, "rows (%d) != loaded images (%d)", n, ctx->n_images);
    RESERVE(ctx, ctx->d_rows, (size_t)n * (sizeof(cspb_row) + sizeof(CtfCoef)));
    CtfCoef *d_ctf = reinterpret_cast<CtfCoef *>(ctx->d_rows.as<cspb_row>() + n);
    ctf_coef_kernel<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(rows_dev, n, ctx->rcfg.box, d_ctf);
    KERNEL_CHECK(ctx);
    return refine_local_enqueue(ctx, rows_dev, d_ctf, n, nullptr, n_evals_out);
}

extern "C" int cspb_refine_run(cspb_ctx *ctx, cspb_row *rows, int n, cspb_row *changes_out, int64_t *n_evals_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (n != ctx->n_images) return cspb_fail(ctx, CSPB_E_ARG, "rows (%d) != loaded images (%d)", n, ctx->n_images);
    if (n == 0) { if (n_evals_out) *n_evals_out = 0; return 0; }
    cspb_row *d_rows;
    CtfCoef *d_ctf;
    int rc = upload_rows(ctx, rows, n, &d_rows, &d_ctf);
    if (rc) return rc;
    DevBuf d_chg;
    if (changes_out) RESERVE(ctx, d_chg, (size_t)n * sizeof(cspb_row));
    rc = refine_local_enqueue(ctx, d_rows, d_ctf, n, changes_out ? d_chg.as<cspb_row>() : nullptr, n_evals_out);
    if (rc) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(rows, d_rows, (size_t)n * sizeof(cspb_row), cudaMemcpyDeviceToHost, ctx->stream));
    if (changes_out)
        CU_TRY(ctx, cudaMemcpyAsync(changes_out, d_chg.p, (size_t)n * sizeof(cspb_row), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ================================================================== building blocks
extern "C" int cspb_ctf_image(cspb_ctx *ctx, const cspb_row *row, int n, float *out) {
    CSPB_ENTER(ctx);
    if (!ctx || !row || !out || n < 2) return CSPB_E_ARG;
    const int nh = n / 2 + 1;
    RESERVE(ctx, ctx->d_work2, (size_t)n * nh * sizeof(float));
    const CtfCoef cc = make_ctf_coef(row->defocus_1, row->defocus_2, row->defocus_angle, row->phase_shift,
                                     row->pixel_size, row->voltage_kv, row->cs_mm, row->amplitude_contrast, n);
    ctf_image_kernel<<<grid_for((long long)n * nh, 256, ctx->sm_count), 256, 0, ctx->stream>>>(cc, n, ctx->d_work2.as<float>());
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_work2.p, (size_t)n * nh * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cspb_project(cspb_ctx *ctx, float psi, float theta, float phi, float *out_complex) {
    CSPB_ENTER(ctx);
    if (!ctx || !out_complex) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    const int n = ctx->rcfg.box, nh = n / 2 + 1;
    RESERVE(ctx, ctx->d_work2, (size_t)n * nh * sizeof(float2));
    project_kernel<<<grid_for((long long)n * nh, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
        ctx->ref.d_ref4.as<float4>(), ctx->ref.sx, ctx->ref.sy, ctx->ref.rc, (float)ctx->ref.pad, n, ctx->plan.r_hi,
        psi, theta, phi, ctx->d_work2.as<float2>());
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaMemcpyAsync(out_complex, ctx->d_work2.p, (size_t)n * nh * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ================================================================== gather microbenchmark
// The denominator SURVEY.md §8d asks for next to the HBM peak: how fast can 32-lane gathers of
// 32-byte items (the scorer's access) be served from a window of `window_bytes` — small window =
// L1-resident, tens of MB = L2-resident, GBs = HBM.  Every lane walks its own LCG sequence.
namespace {
__global__ void __launch_bounds__(128, 8) gather_peak_kernel(const RefQuad *__restrict__ buf, unsigned n_items, int iters,
                                                            unsigned per_cta_window, float *__restrict__ sink) {
    unsigned state = (blockIdx.x * 128u + threadIdx.x) * 2654435761u + 12345u;
    // per-CTA window (L1 test) or the whole buffer
    const unsigned base = per_cta_window ? (unsigned)(((unsigned long long)blockIdx.x * per_cta_window) % (n_items - per_cta_window)) : 0u;
    const unsigned span = per_cta_window ? per_cta_window : n_items;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        RefQuad q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            state = state * 1664525u + 1013904223u;
            q[k] = ldg_quad(buf + base + (state >> 8) % span);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += q[k].v00.x + q[k].v10.y + q[k].v01.x + q[k].v11.y;
    }
    if (acc == 123.456f) sink[0] = acc;
}
}  // namespace

extern "C" int cspb_gather_peak(cspb_ctx *ctx, size_t window_bytes, int per_cta, float *gbs_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !gbs_out || window_bytes < 4096) return CSPB_E_ARG;
    const size_t buf_bytes = per_cta ? ((size_t)256 << 20) : window_bytes;
    DevBuf buf, sink;
    RESERVE(ctx, buf, buf_bytes);
    RESERVE(ctx, sink, 16);
    CU_TRY(ctx, cudaMemsetAsync(buf.p, 0, buf_bytes, ctx->stream));
    const unsigned n_items = (unsigned)(buf_bytes / sizeof(RefQuad));
    const unsigned win = per_cta ? (unsigned)(window_bytes / sizeof(RefQuad)) : 0u;
    const int grid = ctx->sm_count * 8 * 4, iters = 2000;
    cudaEvent_t e0, e1;
    CU_TRY(ctx, cudaEventCreate(&e0));
    CU_TRY(ctx, cudaEventCreate(&e1));
    gather_peak_kernel<<<grid, 128, 0, ctx->stream>>>(buf.as<RefQuad>(), n_items, 50, win, sink.as<float>());  // warm-up
    CU_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    gather_peak_kernel<<<grid, 128, 0, ctx->stream>>>(buf.as<RefQuad>(), n_items, iters, win, sink.as<float>());
    CU_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gbs_out = (float)((double)grid * 128.0 * iters * 4.0 * sizeof(RefQuad) / (ms * 1e-3) / 1e9);
    return 0;
}
