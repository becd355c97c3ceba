// recon.cu — Fourier-space 3-D reconstruction: CTF-weighted trilinear insertion with CTF^2
// weight accumulation (reconstruct3d), accumulator dump / add (local_merge3d) and the final
// statistics + optimal-filter normalisation + gridding correction (merge3d).
//
// Replaces external/cistem2/{reconstruct3d,local_merge3d,merge3d} (closed LFS binaries); the
// contracts are src/pyp/refine/frealign/frealign.py:1780-1824, :1878-1888, :2075-2093.
// Semantics: oracle/SEMANTICS.md §reconstruction; CPU restatement: oracle/cspb_oracle.c.
//
// Accumulator layout (ours): one float4 {sum re, sum im, sum ctf^2 w, 0} per voxel of the
// Hermitian half-volume, x in [0, np/2], y and z centred (index y + np/2), x fastest.  One
// 16-byte vector atomic (red.global.add.v4.f32, sm_90+) per trilinear corner.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "device_math.cuh"
#include "internal.cuh"

namespace {

struct InsertArgs {
    const float2 *spec;   // chunk of half spectra (unnormalised, origin at pixel 0)
    const cspb_row *rows; // chunk rows (device)
    int n, count;
    int np, xh;
    float padf;
    float rmax2;          // resolution limit in Fourier pixels, squared
    float bfac_k;         // score_bfactor / 4 * s2u  (0 = no score weighting)
    float avg_score;
    float score_threshold;
    int per_particle;
    const float *sym;     // n_sym row-major 3x3
    int n_sym;
    float4 *acc0, *acc1;
    int tiles;            // tiles per image
    const float2 *aux;    // optional per row {weight, cut radius in Fourier pixels}: data-driven dose weighting (SEMANTICS.md §10)
    float aux_width;      // width of the raised-cosine edge at the cut radius
    const float2 *rescale; // optional per row {factor, DC term}: `spec` holds transforms normalised for refinement (kept spectra)
};

__device__ __forceinline__ void add_corner(float4 *acc, int np, int xh, int x, int y, int z, float w, float re,
                                           float im, float wt) {
    const int c = np / 2;
    if (x > c || y < -c || y >= c || z < -c || z >= c) return;
    const long long idx = ((long long)(z + c) * np + (y + c)) * xh + x;
    atomicAdd(acc + idx, make_float4(w * re, w * im, w * wt, 0.f));
}

#define INSERT_MAX_SYM 64
__global__ void __launch_bounds__(256) insert_kernel(const InsertArgs A) {
    __shared__ float s_mat[INSERT_MAX_SYM][6];  // first two columns of pad * S_k * M
    __shared__ CtfCoef s_ctf;
    // one launch covers both halves without idle CTAs: the grid walks the even-indexed images of the chunk first and the
    // odd-indexed ones after, so that with the usual alternating split (odd / even POSITION_IN_STACK) the first half of the
    // launch touches one accumulator and the second half the other — each half-sphere (70 MB at 256 px) stays L2-resident
    // while it is being hit, as with one pass per half, at half the CTAs
    const int k = blockIdx.x / A.tiles, tile = blockIdx.x - k * A.tiles;
    const int n_even = (A.count + 1) >> 1;
    const int img = k < n_even ? 2 * k : 2 * (k - n_even) + 1;
    const cspb_row row = A.rows[img];
    if (!(row.occupancy > 0.f) || row.score < A.score_threshold) return;
    const int half = A.per_particle ? (row.pind & 1) : ((row.position_in_stack & 1u) ? 0 : 1);
    const int n = A.n, nh = n / 2 + 1;
    if (threadIdx.x < A.n_sym) {
        float m[9];
        euler_matrix(row.psi, row.theta, row.phi, m);
        const float *S = A.sym + 9 * threadIdx.x;
        float *o = s_mat[threadIdx.x];
        o[0] = (S[0] * m[0] + S[1] * m[3] + S[2] * m[6]) * A.padf;
        o[1] = (S[0] * m[1] + S[1] * m[4] + S[2] * m[7]) * A.padf;
        o[2] = (S[3] * m[0] + S[4] * m[3] + S[5] * m[6]) * A.padf;
        o[3] = (S[3] * m[1] + S[4] * m[4] + S[5] * m[7]) * A.padf;
        o[4] = (S[6] * m[0] + S[7] * m[3] + S[8] * m[6]) * A.padf;
        o[5] = (S[6] * m[1] + S[7] * m[4] + S[8] * m[7]) * A.padf;
    }
    if (threadIdx.x == 255)
        s_ctf = make_ctf_coef(row.defocus_1, row.defocus_2, row.defocus_angle, row.phase_shift, row.pixel_size,
                              row.voltage_kv, row.cs_mm, row.amplitude_contrast, n);
    __syncthreads();
#ifndef CSPB_INSERT_PAIR
#define CSPB_INSERT_PAIR 0
#endif
    // CSPB_INSERT_PAIR=1 (A/B build, not the default): two lanes per sample, lane parity = the x corner (x0 / x0 + 1), so that
    // the two neighbouring float4 cells of an x pair leave in ONE RED instruction.  Measured slower (r02m: 32.4 vs 29.3 ms per
    // 32 768 particles): vector REDs of adjacent lanes are not merged into one sector operation, and the doubled per-sample
    // arithmetic is not free.
    const int lane_dx = CSPB_INSERT_PAIR ? (threadIdx.x & 1) : 0;
    const int idx = CSPB_INSERT_PAIR ? tile * (blockDim.x >> 1) + (threadIdx.x >> 1) : tile * blockDim.x + threadIdx.x;
    if (idx >= n * nh) return;
    const int i = idx % nh;
    int j = idx / nh;
    if (j >= n / 2) j -= n;
    if (i == 0 && j < 0) return;  // Friedel mates of (0, -j): inserted once
    const float fi = (float)i, fj = (float)j;
    const float r2 = fi * fi + fj * fj;
    if (r2 > A.rmax2) return;
    float2 F = __ldcs(A.spec + (long long)img * n * nh + idx);  // streamed once: do not displace the accumulators in L2
    if (A.rescale) {  // F_recon = (scl_r / scl_f) F_refine + scl_r (off_f - off_r) n^2 delta(0): the transform is linear
        const float2 rs = A.rescale[img];
        F.x *= rs.x;
        F.y *= rs.x;
        if (idx == 0) F.x += rs.y;
    }
    if ((i + j) & 1) { F.x = -F.x; F.y = -F.y; }  // box centre at n/2
    const float ctf = -sinpif(ctf_chi(s_ctf, fi, fj, r2) * (1.f / CSPB_PI_F));
    float w = row.occupancy * 0.01f;
    if (A.bfac_k != 0.f) w *= expf(-A.bfac_k * (A.avg_score - row.score) * r2);
    if (A.aux) {
        const float2 wc = A.aux[img];
        w *= wc.x;
        if (wc.y > 0.f) w *= cosine_edge(sqrtf(r2), wc.y, A.aux_width);
    }
    // undo the particle shift: multiply by exp(+2 pi i (i sx + j sy) / n), shifts in pixels
    const float k2 = 2.f / ((float)n * row.pixel_size);
    float sn, cs;
    sincospif((fi * row.x_shift + fj * row.y_shift) * k2, &sn, &cs);
    const float re = (F.x * cs - F.y * sn) * ctf, im = (F.x * sn + F.y * cs) * ctf;
    const float wt = ctf * ctf;
    float4 *acc = half ? A.acc1 : A.acc0;
    for (int s = 0; s < A.n_sym; ++s) {
        const float *o = s_mat[s];
        float x = o[0] * fi + o[1] * fj, y = o[2] * fi + o[3] * fj, z = o[4] * fi + o[5] * fj;
        float vim = im;
        if (x < 0.f) { x = -x; y = -y; z = -z; vim = -im; }
        const float x0f = floorf(x), y0f = floorf(y), z0f = floorf(z);
        const float fx = x - x0f, fy = y - y0f, fz = z - z0f;
        const int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
        const float wx0 = 1.f - fx, wy0 = 1.f - fy, wz0 = 1.f - fz;
        if (CSPB_INSERT_PAIR) {
            const float wx = w * (lane_dx ? fx : wx0);
            const int xc = x0 + lane_dx;
            add_corner(acc, A.np, A.xh, xc, y0, z0, wx * wy0 * wz0, re, vim, wt);
            add_corner(acc, A.np, A.xh, xc, y0 + 1, z0, wx * fy * wz0, re, vim, wt);
            add_corner(acc, A.np, A.xh, xc, y0, z0 + 1, wx * wy0 * fz, re, vim, wt);
            add_corner(acc, A.np, A.xh, xc, y0 + 1, z0 + 1, wx * fy * fz, re, vim, wt);
        } else {
            add_corner(acc, A.np, A.xh, x0, y0, z0, w * wx0 * wy0 * wz0, re, vim, wt);
            add_corner(acc, A.np, A.xh, x0 + 1, y0, z0, w * fx * wy0 * wz0, re, vim, wt);
            add_corner(acc, A.np, A.xh, x0, y0 + 1, z0, w * wx0 * fy * wz0, re, vim, wt);
            add_corner(acc, A.np, A.xh, x0 + 1, y0 + 1, z0, w * fx * fy * wz0, re, vim, wt);
            add_corner(acc, A.np, A.xh, x0, y0, z0 + 1, w * wx0 * wy0 * fz, re, vim, wt);
            add_corner(acc, A.np, A.xh, x0 + 1, y0, z0 + 1, w * fx * wy0 * fz, re, vim, wt);
            add_corner(acc, A.np, A.xh, x0, y0 + 1, z0 + 1, w * wx0 * fy * fz, re, vim, wt);
            add_corner(acc, A.np, A.xh, x0 + 1, y0 + 1, z0 + 1, w * fx * fy * fz, re, vim, wt);
        }
    }
}

// ---- pull-based insertion (CSPB_INSERT=pull; NOT the default, see cspb_recon_insert_weighted): one RED per touched voxel
// instead of eight per sample.
// A central slice is a plane through the lattice: seen along its dominant axis w (the largest component of the plane
// normal) it is a height field over the (u, v) lattice, and every (u, v) column receives weight on at most six
// consecutive w levels around the plane.  One CTA owns a 16 x 16 tile of columns of one (projection, operator) plane:
//   1. the samples that can reach the tile (preimage of the tile + 1 under the 2 x 2 map (i, j) -> (u, v)) are staged in
//      shared memory ONCE with everything that does not depend on the corner: CTF, un-shift phase, weights, the Friedel
//      rule (the half volume keeps x >= 0: of a sample and its mate -t exactly one lands there);
//   2. every thread owns a column, walks the few samples within one voxel of it, and accumulates the bilinear (u, v) x
//      linear (w) weights into registers — no atomics, no shared-memory conflicts;
//   3. the non-empty levels go out with one vector RED each: ~2.5 per column against 8 per sample, about 3x fewer L2
//      atomic operations for the same sums (the sums are re-associated, nothing else changes).
#define PULL_T 16
#define PULL_MAXS (48 * 48)
struct PullGeo {
    float auu, aub, ava, avb;      // (u, v) = A (i, j)
    float i00, i01, i10, i11;      // A^-1
    float ci, cj;                  // w = ci i + cj j
    float ax, bx;                  // x = ax i + bx j (the Friedel rule)
    float alpha, beta;             // plane height over the columns: w = alpha u + beta v
    int pu, pv, pw;                // axis of u, v, w in (x, y, z)
};

__global__ void __launch_bounds__(256) insert_pull_kernel(const InsertArgs A, int tu, int rb) {
    __shared__ PullGeo G;
    __shared__ CtfCoef s_ctf;
    __shared__ float4 s_v[PULL_MAXS];
    __shared__ int s_box[4];
    const int tiles = tu * tu;
    int bid = blockIdx.x;
    const int tile = bid % tiles;
    bid /= tiles;
    const int op = bid % A.n_sym, k = bid / A.n_sym;
    const int n_even = (A.count + 1) >> 1;
    const int img = k < n_even ? 2 * k : 2 * (k - n_even) + 1;  // even-indexed images first: one half at a time stays L2-resident
    const cspb_row row = A.rows[img];
    if (!(row.occupancy > 0.f) || row.score < A.score_threshold) return;
    const int half = A.per_particle ? (row.pind & 1) : ((row.position_in_stack & 1u) ? 0 : 1);
    const int n = A.n, nh = n / 2 + 1;
    const int u0 = -rb + PULL_T * (tile % tu), v0 = -rb + PULL_T * (tile / tu);
    {   // tile outside the disc of inserted samples (+1 for the trilinear footprint)
        const int nu = u0 > 0 ? u0 : (u0 + PULL_T - 1 < 0 ? u0 + PULL_T - 1 : 0);
        const int nv = v0 > 0 ? v0 : (v0 + PULL_T - 1 < 0 ? v0 + PULL_T - 1 : 0);
        if (nu * nu + nv * nv > rb * rb) return;
    }
    if (threadIdx.x == 0) {
        float m[9];
        euler_matrix(row.psi, row.theta, row.phi, m);
        const float *S = A.sym + 9 * op;
        float a[3], b[3];
        for (int r = 0; r < 3; ++r) {
            a[r] = (S[3 * r] * m[0] + S[3 * r + 1] * m[3] + S[3 * r + 2] * m[6]) * A.padf;
            b[r] = (S[3 * r] * m[1] + S[3 * r + 1] * m[4] + S[3 * r + 2] * m[7]) * A.padf;
        }
        const float nrm[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        int w = 0;
        if (fabsf(nrm[1]) > fabsf(nrm[w])) w = 1;
        if (fabsf(nrm[2]) > fabsf(nrm[w])) w = 2;
        PullGeo g;
        g.pw = w; g.pu = (w + 1) % 3; g.pv = (w + 2) % 3;
        g.auu = a[g.pu]; g.aub = b[g.pu]; g.ava = a[g.pv]; g.avb = b[g.pv];
        const float det = g.auu * g.avb - g.aub * g.ava, inv = 1.f / det;
        g.i00 = g.avb * inv; g.i01 = -g.aub * inv; g.i10 = -g.ava * inv; g.i11 = g.auu * inv;
        g.ci = a[w]; g.cj = b[w];
        g.ax = a[0]; g.bx = b[0];
        g.alpha = g.ci * g.i00 + g.cj * g.i10;
        g.beta = g.ci * g.i01 + g.cj * g.i11;
        G = g;
        // samples that can reach the tile: preimage of [u0 - 1, u0 + T] x [v0 - 1, v0 + T]
        float ilo = 1e30f, ihi = -1e30f, jlo = 1e30f, jhi = -1e30f;
        for (int c = 0; c < 4; ++c) {
            const float uu = (float)((c & 1) ? u0 + PULL_T : u0 - 1), vv = (float)((c & 2) ? v0 + PULL_T : v0 - 1);
            const float fi = g.i00 * uu + g.i01 * vv, fj = g.i10 * uu + g.i11 * vv;
            ilo = fminf(ilo, fi); ihi = fmaxf(ihi, fi); jlo = fminf(jlo, fj); jhi = fmaxf(jhi, fj);
        }
        s_box[0] = (int)floorf(ilo); s_box[1] = (int)ceilf(ihi) - s_box[0] + 1;
        s_box[2] = (int)floorf(jlo); s_box[3] = (int)ceilf(jhi) - s_box[2] + 1;
    }
    if (threadIdx.x == 255)
        s_ctf = make_ctf_coef(row.defocus_1, row.defocus_2, row.defocus_angle, row.phase_shift, row.pixel_size,
                              row.voltage_kv, row.cs_mm, row.amplitude_contrast, n);
    __syncthreads();
    // the x axis of the half volume among the column axes: columns on the negative side receive nothing
    if (G.pu == 0 && u0 + PULL_T - 1 < 0) return;
    if (G.pv == 0 && v0 + PULL_T - 1 < 0) return;
    const int i_lo = s_box[0], ni = s_box[1], j_lo = s_box[2];
    int nj = s_box[3];
    if (ni * nj > PULL_MAXS) nj = PULL_MAXS / ni;  // cannot happen for pad >= 1 (|det A| >= pad^2 / sqrt 3), kept as a bound
    const float k2 = 2.f / ((float)n * row.pixel_size);
    const float2 *spec = A.spec + (long long)img * n * nh;
    float wrow = row.occupancy * 0.01f, cutr = 0.f;
    if (A.aux) { const float2 wc = A.aux[img]; wrow *= wc.x; cutr = wc.y; }
    for (int s = threadIdx.x; s < ni * nj; s += 256) {
        const int i = i_lo + s % ni, j = j_lo + s / ni;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        const float fi = (float)i, fj = (float)j;
        const float r2 = fi * fi + fj * fj;
        // of the sample t = (i, j) and its Friedel mate -t exactly one lies in the stored half space x >= 0; on x = 0 the
        // stored representative (i > 0, or i = 0 and j >= 0) is the one inserted
        const float px = G.ax * fi + G.bx * fj;
        const bool rep = i > 0 || (i == 0 && j >= 0);
        if (r2 <= A.rmax2 && (px > 0.f || (px == 0.f && rep))) {
            const int si = rep ? i : -i, sj = rep ? j : -j;   // stored sample
            float2 F = __ldcs(spec + (sj < 0 ? sj + n : sj) * nh + si);
            if ((si + sj) & 1) { F.x = -F.x; F.y = -F.y; }  // box centre at n/2
            if (!rep) F.y = -F.y;                             // Hermitian extension
            const float ctf = -sinpif(ctf_chi(s_ctf, fi, fj, r2) * (1.f / CSPB_PI_F));
            float w = wrow;
            if (A.bfac_k != 0.f) w *= expf(-A.bfac_k * (A.avg_score - row.score) * r2);
            if (cutr > 0.f) w *= cosine_edge(sqrtf(r2), cutr, A.aux_width);
            float sn, cs;
            sincospif((fi * row.x_shift + fj * row.y_shift) * k2, &sn, &cs);
            val = make_float4(w * (F.x * cs - F.y * sn) * ctf, w * (F.x * sn + F.y * cs) * ctf, w * ctf * ctf, 1.f);
        }
        s_v[s] = val;
    }
    __syncthreads();
    const int u = u0 + (threadIdx.x & (PULL_T - 1)), v = v0 + threadIdx.x / PULL_T;
    const float fu = (float)u, fv = (float)v;
    const float ic = G.i00 * fu + G.i01 * fv, jc = G.i10 * fu + G.i11 * fv;
    const float ei = fabsf(G.i00) + fabsf(G.i01), ej = fabsf(G.i10) + fabsf(G.i11);
    int ia = (int)ceilf(ic - ei), ib = (int)floorf(ic + ei), ja = (int)ceilf(jc - ej), jb = (int)floorf(jc + ej);
    ia = max(ia, i_lo); ib = min(ib, i_lo + ni - 1); ja = max(ja, j_lo); jb = min(jb, j_lo + nj - 1);
    const int zb = (int)floorf(G.alpha * fu + G.beta * fv) - 2;
    float acc[6][3];
#pragma unroll
    for (int l = 0; l < 6; ++l) acc[l][0] = acc[l][1] = acc[l][2] = 0.f;
    for (int j = ja; j <= jb; ++j)
        for (int i = ia; i <= ib; ++i) {
            const float4 sv = s_v[(j - j_lo) * ni + (i - i_lo)];
            if (sv.w == 0.f) continue;
            const float fi = (float)i, fj = (float)j;
            const float du = fabsf(G.auu * fi + G.aub * fj - fu), dv = fabsf(G.ava * fi + G.avb * fj - fv);
            if (du >= 1.f || dv >= 1.f) continue;
            const float wuv = (1.f - du) * (1.f - dv);
            const float h = G.ci * fi + G.cj * fj;
            const float z0f = floorf(h);
            const float fz = h - z0f;
            const int l0 = (int)z0f - zb;
            const float w0 = wuv * (1.f - fz), w1 = wuv * fz;
            if (l0 < 0 || l0 > 4) {
                // |alpha| + |beta| <= 2 bounds the level to [0, 4]; a rounding tie in the dominant-axis choice can leave it
                // one off: such a sample goes out directly
                const float4 *sv4 = &sv;
                for (int q = 0; q < 2; ++q) {
                    const float wl = q ? w1 : w0;
                    const int pw_ = (int)z0f + q;
                    const int x = G.pu == 0 ? u : (G.pv == 0 ? v : pw_), y = G.pu == 1 ? u : (G.pv == 1 ? v : pw_), z = G.pu == 2 ? u : (G.pv == 2 ? v : pw_);
                    const int c_ = A.np / 2;
                    if (wl != 0.f && x >= 0 && x <= c_ && y >= -c_ && y < c_ && z >= -c_ && z < c_)
                        atomicAdd((half ? A.acc1 : A.acc0) + ((long long)(z + c_) * A.np + (y + c_)) * A.xh + x,
                                  make_float4(wl * sv4->x, wl * sv4->y, wl * sv4->z, 0.f));
                }
                continue;
            }
#pragma unroll
            for (int l = 0; l < 6; ++l) {
                const float wl = (l == l0) ? w0 : ((l == l0 + 1) ? w1 : 0.f);
                acc[l][0] += wl * sv.x; acc[l][1] += wl * sv.y; acc[l][2] += wl * sv.z;
            }
        }
    float4 *dst = half ? A.acc1 : A.acc0;
    const int c = A.np / 2;
#pragma unroll
    for (int l = 0; l < 6; ++l) {
        if (acc[l][0] == 0.f && acc[l][1] == 0.f && acc[l][2] == 0.f) continue;
        const int pw_ = zb + l;
        const int x = G.pu == 0 ? u : (G.pv == 0 ? v : pw_), y = G.pu == 1 ? u : (G.pv == 1 ? v : pw_), z = G.pu == 2 ? u : (G.pv == 2 ? v : pw_);
        if (x < 0 || x > c || y < -c || y >= c || z < -c || z >= c) continue;
        atomicAdd(dst + ((long long)(z + c) * A.np + (y + c)) * A.xh + x, make_float4(acc[l][0], acc[l][1], acc[l][2], 0.f));
    }
}

__global__ void add_volume_kernel(float4 *__restrict__ dst, const float4 *__restrict__ src, long long nvox) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nvox; k += (long long)gridDim.x * blockDim.x) {
        float4 a = dst[k];
        const float4 b = src[k];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        dst[k] = a;
    }
}

// x = 0 plane of the Hermitian half-volume: every inserted sample also stands for its Friedel
// mate, which lands on (0,-y,-z) conjugated.  Readers fold the mate in on the fly, so the
// accumulators themselves stay pure sums (dumps can be merged in any order).
__device__ __forceinline__ float4 load_sym(const float4 *__restrict__ acc, int np, int xh, int x, int y, int z) {
    const int c = np / 2;
    float4 a = acc[((long long)(z + c) * np + (y + c)) * xh + x];
    if (x == 0 && y != -c && z != -c) {
        const float4 b = acc[((long long)(-z + c) * np + (-y + c)) * xh];
        a.x += b.x;
        a.y -= b.y;
        a.z += b.z;
    }
    return a;
}

// acc[v] += sum_h raw[h^-1 v] over the lattice-preserving symmetry operators (exact: trilinear
// weights are invariant under signed axis permutations).  Sources are read through load_sym so the
// x = 0 plane arrives complete; destinations on x = 0 are stored halved so that load_sym's later
// folding restores the total.
__global__ void lattice_sym_kernel(const float4 *__restrict__ raw, float4 *__restrict__ acc, int np, int xh,
                                   const int *__restrict__ ht, int n_lat) {
    const int c = np / 2;
    const long long total = (long long)xh * np * np;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(k % xh), y = (int)((k / xh) % np) - c, z = (int)(k / ((long long)xh * np)) - c;
        float sx = 0.f, sy = 0.f, sw = 0.f;
        for (int h = 0; h < n_lat; ++h) {
            const int *m = ht + 9 * h;
            int ux = m[0] * x + m[1] * y + m[2] * z, uy = m[3] * x + m[4] * y + m[5] * z, uz = m[6] * x + m[7] * y + m[8] * z;
            float sg = 1.f;
            if (ux < 0) { ux = -ux; uy = -uy; uz = -uz; sg = -1.f; }
            if (ux > c || uy < -c || uy >= c || uz < -c || uz >= c) continue;
            const float4 v = load_sym(raw, np, xh, ux, uy, uz);
            sx += v.x;
            sy += sg * v.y;
            sw += v.z;
        }
        if (x == 0 && y != -c && z != -c) { sx *= 0.5f; sy *= 0.5f; sw *= 0.5f; }
        float4 a = acc[k];
        a.x += sx; a.y += sy; a.z += sw;
        acc[k] = a;
    }
}

// per-shell sums for FSC: {sum Re(V1 V2*), sum |V1|^2, sum |V2|^2, sum (W1+W2), count, sumW1, sumW2}
#define SHELL_Q 7
__global__ void shell_stats_kernel(const float4 *__restrict__ a0, const float4 *__restrict__ a1, int np, int xh,
                                   float inv_pad, int n_shells, double *__restrict__ out) {
    extern __shared__ float sh[];  // n_shells * SHELL_Q
    for (int k = threadIdx.x; k < n_shells * SHELL_Q; k += blockDim.x) sh[k] = 0.f;
    __syncthreads();
    const int c = np / 2;
    const long long total = (long long)xh * np * np;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(k % xh), y = (int)((k / xh) % np) - c, z = (int)(k / ((long long)xh * np)) - c;
        const float r = sqrtf((float)(x * x + y * y + z * z)) * inv_pad;
        const int s = (int)(r + 0.5f);
        if (s >= n_shells) continue;
        const float4 p = load_sym(a0, np, xh, x, y, z), q = load_sym(a1, np, xh, x, y, z);
        if (!(p.z > 0.f) || !(q.z > 0.f)) continue;
        const float mult = (x == 0) ? 0.5f : 1.f;  // x = 0 plane stores both Friedel mates
        const float v1x = p.x / p.z, v1y = p.y / p.z, v2x = q.x / q.z, v2y = q.y / q.z;
        float *d = sh + s * SHELL_Q;
        atomicAdd(d + 0, mult * (v1x * v2x + v1y * v2y));
        atomicAdd(d + 1, mult * (v1x * v1x + v1y * v1y));
        atomicAdd(d + 2, mult * (v2x * v2x + v2y * v2y));
        atomicAdd(d + 3, mult * (p.z + q.z));
        atomicAdd(d + 4, mult);
        atomicAdd(d + 5, mult * p.z);
        atomicAdd(d + 6, mult * q.z);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_shells * SHELL_Q; k += blockDim.x)
        if (sh[k] != 0.f) atomicAdd(out + k, (double)sh[k]);
}

// per-shell FSC -> SSNR -> Wiener terms and the 7-column statistics row (one thread per shell)
__global__ void shell_terms_kernel(const double *__restrict__ sh, int ns, double frac, float box_a, int n,
                                   float *__restrict__ term, float *__restrict__ stats) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const double *q = sh + (size_t)s * SHELL_Q;
    double fsc = 0.0;
    if (q[1] > 0.0 && q[2] > 0.0) fsc = q[0] / sqrt(q[1] * q[2]);
    if (s == 0 && q[4] > 0.0) fsc = 1.0;
    double f = fsc;
    if (f > 0.9999) f = 0.9999;
    if (f < 0.0) f = 0.0;
    const double rec_ssnr = 2.0 * f / (1.0 - f);
    const double part_ssnr = rec_ssnr / frac;
    const double part_fsc = part_ssnr / (2.0 + part_ssnr);
    const double cntv = q[4] > 0.0 ? q[4] : 1.0;
    const double mw = q[3] / cntv, mw0 = q[5] / cntv, mw1 = q[6] / cntv;
    const double ssnr_floor = 1e-4;
    term[s] = (float)(mw / (rec_ssnr > ssnr_floor ? rec_ssnr : ssnr_floor));
    const double half_ssnr = 0.5 * rec_ssnr > ssnr_floor ? 0.5 * rec_ssnr : ssnr_floor;
    term[ns + s] = (float)(mw0 / half_ssnr);
    term[2 * ns + s] = (float)(mw1 / half_ssnr);
    float *o = stats + (size_t)s * 7;
    o[0] = (float)s;
    o[1] = s > 0 ? box_a / (float)s : 0.f;
    o[2] = (float)s / (float)n;
    o[3] = (float)fsc;
    o[4] = (float)part_fsc;
    o[5] = (float)sqrt(part_ssnr);
    o[6] = (float)sqrt(rec_ssnr);
}

// accumulators -> FFT-ordered half spectrum ready for the inverse transform
// mode 0: (a0+a1)/(w0+w1+term), 1: a0/(w0+term), 2: a1/(w1+term); term indexed by shell
__global__ void filter_to_fft_kernel(const float4 *__restrict__ a0, const float4 *__restrict__ a1, int np, int xh,
                                     float inv_pad, int n_shells, const float *__restrict__ term, int mode,
                                     float2 *__restrict__ out) {
    const int c = np / 2;
    const long long total = (long long)xh * np * np;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(k % xh), iy = (int)((k / xh) % np), iz = (int)(k / ((long long)xh * np));
        const int y = iy >= c ? iy - np : iy, z = iz >= c ? iz - np : iz;
        const float r = sqrtf((float)(x * x + y * y + z * z)) * inv_pad;
        const int s = (int)(r + 0.5f);
        float2 v = make_float2(0.f, 0.f);
        if (s < n_shells) {
            float re, im, w;
            if (mode == 0) {
                const float4 p = load_sym(a0, np, xh, x, y, z), q = load_sym(a1, np, xh, x, y, z);
                re = p.x + q.x; im = p.y + q.y; w = p.z + q.z;
            } else {
                const float4 p = load_sym(mode == 1 ? a0 : a1, np, xh, x, y, z);
                re = p.x; im = p.y; w = p.z;
            }
            const float den = w + term[s];
            if (w > 0.f && den > 0.f) {
                const float sg = ((x + y + z) & 1) ? -1.f : 1.f;  // centre the real-space box at np/2
                v = make_float2(sg * re / den, sg * im / den);
            }
        }
        out[k] = v;
    }
}

// crop centre n^3 of the np^3 real volume, scale, gridding (sinc^2) correction, soft outer mask
__global__ void post_real_kernel(const float *__restrict__ big, int np, int n, float scale, float mask_radius_px,
                                 float mask_width_px, float *__restrict__ out) {
    const long long total = (long long)n * n * n;
    const int off = (np - n) / 2;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(k % n), y = (int)((k / n) % n), z = (int)(k / ((long long)n * n));
        const int dx = x - n / 2, dy = y - n / 2, dz = z - n / 2;
        float v = big[((long long)(z + off) * np + (y + off)) * np + (x + off)] * scale;
        v /= sinc2_corr(dx, np) * sinc2_corr(dy, np) * sinc2_corr(dz, np);
        const float r = sqrtf((float)(dx * dx + dy * dy + dz * dz));
        v *= cosine_edge(r, mask_radius_px, mask_width_px);
        out[k] = v;
    }
}

int grid_for(long long total, int block, int sm) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)sm * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// kept spectra (cspb_refine_keep_spectra): per image the factor and the DC term that turn the refinement normalisation
// (off_f, scl_f) into the reconstruction's (off_r, scl_r)
__global__ void keep_rescale_kernel(const float *__restrict__ off_r, const float *__restrict__ scl_r, const float *__restrict__ off_f,
                                    const float *__restrict__ scl_f, int count, float n2, float2 *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    out[k] = make_float2(scl_r[k] / scl_f[k], scl_r[k] * (off_f[k] - off_r[k]) * n2);
}

__global__ void image_stats_recon_kernel(const float *__restrict__ img, int n, float radius, int normalize, int invert,
                                         float *__restrict__ offs, float *__restrict__ scls) {
    __shared__ float red[64];
    image_edge_stats(img + (long long)blockIdx.x * n * n, n, radius, normalize, invert, offs + blockIdx.x, scls + blockIdx.x, red);
}

}  // namespace

extern "C" int cspb_recon_cfg_default(cspb_recon_cfg *cfg, int box, float pixel_size) {
    if (!cfg || box <= 0 || pixel_size <= 0.f) return CSPB_E_ARG;
    memset(cfg, 0, sizeof *cfg);
    cfg->box = box;
    cfg->pad = 1;                                  // frealign.py:1769
    cfg->pixel_size = pixel_size;
    cfg->mask_radius = pixel_size * box / 2.f;     // frealign.py:1650-1654 rad_rec default
    cfg->resolution_limit = 2.f * pixel_size;      // frealign.py:1644-1649 res_rec default
    cfg->score_bfactor = 2.f;                      // refine_bsc default
    cfg->score_weighting = 0;
    cfg->score_threshold = 0.f;
    cfg->normalize = 1;
    cfg->invert_contrast = 0;
    cfg->per_particle_split = 0;
    cfg->average_score = 0.f;
    return 0;
}

extern "C" int cspb_recon_begin(cspb_ctx *ctx, const cspb_recon_cfg *cfg) {
    CSPB_ENTER(ctx);
    if (!ctx || !cfg) return CSPB_E_ARG;
    if (cfg->box < 16 || (cfg->box & 1) || (cfg->pad != 1 && cfg->pad != 2) || cfg->pixel_size <= 0.f)
        return cspb_fail(ctx, CSPB_E_ARG, "bad recon box/pad/pixel");
    ctx->ccfg = *cfg;
    ctx->rnp = cfg->box * cfg->pad;
    const int np = ctx->rnp, xh = np / 2 + 1;
    const size_t bytes = (size_t)xh * np * np * sizeof(float4);
    for (int h = 0; h < 2; ++h) {
        RESERVE(ctx, ctx->d_acc[h], bytes);
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_acc[h].p, 0, bytes, ctx->stream));
    }
    if (ctx->sym.empty()) {
        const float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        int rc = cspb_set_symmetry(ctx, id, 1);
        if (rc) return rc;
    }
    if (ctx->n_lat > 1)
        for (int h = 0; h < 2; ++h) {
            RESERVE(ctx, ctx->d_raw[h], bytes);
            CU_TRY(ctx, cudaMemsetAsync(ctx->d_raw[h].p, 0, bytes, ctx->stream));
        }
    ctx->raw_dirty = false;
    ctx->recon_ready = true;
    ctx->recon_inserted = 0;
    return 0;
}

int recon_flush_deferred(cspb_ctx *ctx) {
    if (!ctx->recon_ready || !ctx->raw_dirty) return 0;
    const int np = ctx->rnp, xh = np / 2 + 1;
    const long long nvox = (long long)xh * np * np;
    for (int h = 0; h < 2; ++h) {
        lattice_sym_kernel<<<grid_for(nvox, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
            ctx->d_raw[h].as<float4>(), ctx->d_acc[h].as<float4>(), np, xh, ctx->d_sym_lat.as<int>(), ctx->n_lat);
        KERNEL_CHECK(ctx);
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_raw[h].p, 0, (size_t)nvox * sizeof(float4), ctx->stream));
    }
    ctx->raw_dirty = false;
    return 0;
}

extern "C" int cspb_recon_dims(const cspb_ctx *ctx, int *np_out, int64_t *floats_per_half_out) {
    if (!ctx || !ctx->recon_ready) return CSPB_E_STATE;
    const int np = ctx->rnp, xh = np / 2 + 1;
    if (np_out) *np_out = np;
    if (floats_per_half_out) *floats_per_half_out = (int64_t)xh * np * np * 4;
    return 0;
}

extern "C" int cspb_recon_device_ptr(cspb_ctx *ctx, int half, void **ptr_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !ctx->recon_ready || half < 0 || half > 1 || !ptr_out) return CSPB_E_ARG;
    int rcf = recon_flush_deferred(ctx);
    if (rcf) return rcf;
    *ptr_out = ctx->d_acc[half].p;
    return 0;
}

extern "C" int cspb_recon_insert(cspb_ctx *ctx, const float *images, const cspb_row *rows, int n_images, int loc) {
    return cspb_recon_insert_weighted(ctx, images, rows, n_images, loc, nullptr);
}

extern "C" int cspb_recon_insert_weighted(cspb_ctx *ctx, const float *images, const cspb_row *rows, int n_images, int loc,
                                          const float *weight_cut) {
    CSPB_ENTER(ctx);
    if (!ctx || !images || !rows || n_images < 0) return CSPB_E_ARG;
    if (!ctx->recon_ready) return cspb_fail(ctx, CSPB_E_STATE, "cspb_recon_begin first");
    const cspb_recon_cfg &c = ctx->ccfg;
    const int n = c.box, nh = n / 2 + 1, np = ctx->rnp, xh = np / 2 + 1;
    const int chunk = chunk_images(n, n_images);
    if (ctx->n_lit > INSERT_MAX_SYM) return cspb_fail(ctx, CSPB_E_ARG, "more than %d symmetry operators", INSERT_MAX_SYM);
    const bool deferred = ctx->n_lat > 1;
    static const bool pull = getenv("CSPB_INSERT") && !strcmp(getenv("CSPB_INSERT"), "pull");
    // the same pixels cspb_refine_load_images has just transformed (cspb_refine_keep_spectra): no second transform
    const float2 *kept = nullptr;
    long long kept0 = 0;
    if (loc == CSPB_DEVICE && !pull && ctx->keep_count > 0 && ctx->keep_box == n && images >= ctx->keep_src) {
        const size_t d = (size_t)(images - ctx->keep_src), per = (size_t)n * n;
        if (d % per == 0 && d / per + (size_t)n_images <= (size_t)ctx->keep_count) {
            kept0 = (long long)(d / per);
            kept = ctx->d_keep_spec.as<float2>() + kept0 * n * nh;
        }
    }
    if (loc == CSPB_HOST && n_images > 0) {
        // host stack: rows (and dose weights) go up once; chunk k+1 of the images is copied on the copy stream into the next
        // staging buffer while chunk k is transformed and inserted
        if (!ctx->pipe_copy) {
            CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->pipe_copy, cudaStreamNonBlocking));
            for (int k = 0; k < CSPB_PIPE_STAGES; ++k) {
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_ready[k], cudaEventDisableTiming));
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_freed[k], cudaEventDisableTiming));
            }
        }
        const int n_buf = n_images > chunk ? CSPB_PIPE_STAGES : 1;
        for (int k = 0; k < n_buf; ++k) RESERVE(ctx, ctx->pipe_stage[k], (size_t)chunk * n * n * sizeof(float));
        RESERVE(ctx, ctx->d_rows, (size_t)n_images * sizeof(cspb_row));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // whoever used the staging buffers before this call is done
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_rows.p, rows, (size_t)n_images * sizeof(cspb_row), cudaMemcpyHostToDevice, ctx->stream));
        if (weight_cut) {
            RESERVE(ctx, ctx->d_aux, (size_t)n_images * sizeof(float2));
            CU_TRY(ctx, cudaMemcpyAsync(ctx->d_aux.p, weight_cut, (size_t)n_images * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    float r_norm = c.mask_radius / c.pixel_size;
    if (r_norm > 0.5f * (float)n) r_norm = 0.5f * (float)n;  // the clamp of image_edge_stats
    const bool same_norm = kept && r_norm == ctx->keep_radius && (c.normalize != 0) == (ctx->keep_normalize != 0) &&
                           (c.invert_contrast != 0) == (ctx->keep_invert != 0);
    // the refinement's pass over the pixels already produced this normalisation (image_stats_dual_kernel)
    const bool have_stats = kept && ctx->keep_recon_valid && c.normalize && r_norm == ctx->keep_recon_radius &&
                            (c.invert_contrast != 0) == (ctx->keep_recon_invert != 0);
    int chunk_index = 0;
    for (int s = 0; s < n_images; s += chunk, ++chunk_index) {
        const int cnt = n_images - s < chunk ? n_images - s : chunk;
        const float *d_img = images + (size_t)s * n * n;
        const cspb_row *d_rows = rows + s;
        const int sb = chunk_index % CSPB_PIPE_STAGES;
        if (loc == CSPB_HOST) {
            if (chunk_index >= CSPB_PIPE_STAGES) CU_TRY(ctx, cudaStreamWaitEvent(ctx->pipe_copy, ctx->pipe_freed[sb], 0));
            CU_TRY(ctx, cudaMemcpyAsync(ctx->pipe_stage[sb].p, d_img, (size_t)cnt * n * n * sizeof(float), cudaMemcpyHostToDevice, ctx->pipe_copy));
            CU_TRY(ctx, cudaEventRecord(ctx->pipe_ready[sb], ctx->pipe_copy));
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_ready[sb], 0));
            d_img = ctx->pipe_stage[sb].as<float>();
            d_rows = ctx->d_rows.as<cspb_row>() + s;
        }
        const float2 *d_aux = nullptr;
        if (weight_cut) d_aux = (loc == CSPB_HOST ? ctx->d_aux.as<float2>() : reinterpret_cast<const float2 *>(weight_cut)) + s;
        RESERVE(ctx, ctx->d_stats, (size_t)2 * chunk * sizeof(float));
        float *offs = ctx->d_stats.as<float>(), *scls = offs + cnt;
        const float2 *rescale = nullptr;
        const float *ks = ctx->d_keep_stats.as<float>();
        if (have_stats && !same_norm) {
            offs = const_cast<float *>(ks) + 2 * (long long)ctx->keep_total + kept0 + s;
            scls = offs + ctx->keep_total;
        } else if (!same_norm) {
            image_stats_recon_kernel<<<cnt, 256, 0, ctx->stream>>>(d_img, n, c.mask_radius / c.pixel_size, c.normalize,
                                                                 c.invert_contrast, offs, scls);
            KERNEL_CHECK(ctx);
        }
        if (kept) {
            if (!same_norm) {
                RESERVE(ctx, ctx->d_keep_scale, (size_t)chunk * sizeof(float2));
                keep_rescale_kernel<<<ceil_div(cnt, 256), 256, 0, ctx->stream>>>(offs, scls, ks + kept0 + s, ks + ctx->keep_total + kept0 + s, cnt,
                                                                               (float)n * (float)n, ctx->d_keep_scale.as<float2>());
                KERNEL_CHECK(ctx);
                rescale = ctx->d_keep_scale.as<float2>();
            }
        } else {
            RESERVE(ctx, ctx->d_work1, (size_t)chunk * n * nh * sizeof(float2));
            int rc = fft2_r2c_dev(ctx, d_img, ctx->d_work1.as<float2>(), n, cnt, offs, scls);
            if (rc) return rc;
        }
        InsertArgs a;
        a.spec = kept ? kept + (long long)s * n * nh : ctx->d_work1.as<float2>();
        a.rescale = rescale;
        a.rows = d_rows;
        a.n = n; a.count = cnt; a.np = np; a.xh = xh;
        a.padf = (float)c.pad;
        float rmax = (float)n * c.pixel_size / (c.resolution_limit > 0.f ? c.resolution_limit : 2.f * c.pixel_size);
        if (rmax > (float)(n / 2 - 1)) rmax = (float)(n / 2 - 1);  // no weight ever lands on the Nyquist planes
        a.rmax2 = rmax * rmax;
        const float s2u = 1.f / (((float)n * c.pixel_size) * ((float)n * c.pixel_size));
        a.bfac_k = c.score_weighting ? c.score_bfactor * 0.25f * s2u : 0.f;
        a.avg_score = c.average_score;
        a.score_threshold = c.score_threshold;
        a.per_particle = c.per_particle_split;
        a.sym = ctx->d_sym_lit.as<float>();
        a.n_sym = ctx->n_lit;
        a.acc0 = (deferred ? ctx->d_raw[0] : ctx->d_acc[0]).as<float4>();
        a.acc1 = (deferred ? ctx->d_raw[1] : ctx->d_acc[1]).as<float4>();
        a.tiles = ceil_div((long long)n * nh, CSPB_INSERT_PAIR ? 128 : 256);
        a.aux = d_aux;
        a.aux_width = 0.1f * 0.5f * (float)n;
        // one launch, ordered by half (see insert_kernel): the voxels one half touches (a half-sphere of radius np/2,
        // 16 B each) fit in L2 and the vector atomics do not spill to HBM
        prof_begin(ctx, CSPB_PROF_INSERT, (int64_t)cnt * ctx->n_lit);
        // default: the per-sample kernel (8 vector REDs per sample).  CSPB_INSERT=pull selects the pull-based kernel below for
        // A/B runs: it issues ~3x fewer L2 atomics but measured 2.3x SLOWER (r02k: 67.9 vs 29.2 ms per 32 768 particles) — the
        // staging redundancy (every sample is prepared by ~2.7 tiles), the candidate walks and the rejected tiles cost more
        // than the atomics they save; profiles/r02_notes.md
        if (!pull) {
            insert_kernel<<<(unsigned)((long long)cnt * a.tiles), 256, 0, ctx->stream>>>(a);
        } else {
            const int rb = (int)ceilf(rmax * (float)c.pad) + 1;        // columns -rb .. rb hold every trilinear footprint
            const int tu = ceil_div(2 * rb + 1, PULL_T);
            insert_pull_kernel<<<(unsigned)((long long)cnt * ctx->n_lit * tu * tu), 256, 0, ctx->stream>>>(a, tu, rb);
        }
        prof_end(ctx);
        KERNEL_CHECK(ctx);
        if (loc == CSPB_HOST) CU_TRY(ctx, cudaEventRecord(ctx->pipe_freed[sb], ctx->stream));
    }
    if (loc == CSPB_HOST && n_images > 0) {  // the caller's host buffers are free again
        CU_TRY(ctx, cudaStreamSynchronize(ctx->pipe_copy));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (deferred && n_images > 0) ctx->raw_dirty = true;
    ctx->recon_inserted += n_images;
    return 0;
}

extern "C" int cspb_recon_get_dump(cspb_ctx *ctx, int half, float *out, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !ctx->recon_ready || half < 0 || half > 1 || !out) return CSPB_E_ARG;
    int rcf = recon_flush_deferred(ctx);
    if (rcf) return rcf;
    const int np = ctx->rnp, xh = np / 2 + 1;
    const size_t bytes = (size_t)xh * np * np * sizeof(float4);
    CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_acc[half].p, bytes, loc == CSPB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cspb_recon_add_dump(cspb_ctx *ctx, int half, const float *in, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !ctx->recon_ready || half < 0 || half > 1 || !in) return CSPB_E_ARG;
    const int np = ctx->rnp, xh = np / 2 + 1;
    const long long nvox = (long long)xh * np * np;
    const float4 *src = reinterpret_cast<const float4 *>(in);
    if (loc == CSPB_HOST) {
        RESERVE(ctx, ctx->d_work1, (size_t)nvox * sizeof(float4));
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_work1.p, in, (size_t)nvox * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
        src = ctx->d_work1.as<float4>();
    }
    add_volume_kernel<<<grid_for(nvox, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->d_acc[half].as<float4>(), src, nvox);
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cspb_recon_finalize(cspb_ctx *ctx, float molecular_mass_kda, float outer_radius_a, float *half1,
                                   float *half2, float *map, float *stats, int n_shells, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !ctx->recon_ready) return CSPB_E_STATE;
    const cspb_recon_cfg &c = ctx->ccfg;
    const int n = c.box, np = ctx->rnp, xh = np / 2 + 1;
    const int ns = n / 2 + 1;
    if (stats && n_shells != ns) return cspb_fail(ctx, CSPB_E_ARG, "stats needs box/2+1 = %d shells", ns);
    int rcf = recon_flush_deferred(ctx);
    if (rcf) return rcf;
    const long long nvox = (long long)xh * np * np;
    const float4 *a0 = ctx->d_acc[0].as<float4>(), *a1 = ctx->d_acc[1].as<float4>();
    // shell statistics -> Wiener terms, all on the device (oracle/SEMANTICS.md §merge3d)
    const size_t sh_bytes = (size_t)ns * SHELL_Q * sizeof(double);
    RESERVE(ctx, ctx->d_shell, sh_bytes + (size_t)(3 + 7) * ns * sizeof(float));
    double *d_sh = ctx->d_shell.as<double>();
    float *d_term = reinterpret_cast<float *>(d_sh + (size_t)ns * SHELL_Q);
    float *d_stats = d_term + (size_t)3 * ns;
    CU_TRY(ctx, cudaMemsetAsync(d_sh, 0, sh_bytes, ctx->stream));
    shell_stats_kernel<<<grid_for(nvox, 256, ctx->sm_count) / 4 + 1, 256, ns * SHELL_Q * sizeof(float), ctx->stream>>>(
        a0, a1, np, xh, 1.f / (float)c.pad, ns, d_sh);
    KERNEL_CHECK(ctx);
    const float box_a = (float)n * c.pixel_size;
    const float rad_a = outer_radius_a > 0.f ? outer_radius_a : 0.5f * box_a;
    double mask_vol = 4.0 / 3.0 * CSPB_PI_D * (double)rad_a * rad_a * rad_a;
    const double box_vol = (double)box_a * box_a * box_a;
    if (mask_vol > box_vol) mask_vol = box_vol;
    double frac = molecular_mass_kda > 0.f ? ((double)molecular_mass_kda * 1000.0 / 0.81) / mask_vol : 1.0;
    if (frac > 1.0 || frac <= 0.0) frac = 1.0;
    shell_terms_kernel<<<ceil_div(ns, 128), 128, 0, ctx->stream>>>(d_sh, ns, frac, box_a, n, d_term, d_stats);
    KERNEL_CHECK(ctx);
    if (stats)
        CU_TRY(ctx, cudaMemcpyAsync(stats, d_stats, (size_t)ns * 7 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    RESERVE(ctx, ctx->d_work1, (size_t)nvox * sizeof(float2));
    RESERVE(ctx, ctx->d_work0, (size_t)np * np * np * sizeof(float));
    if (loc == CSPB_HOST) RESERVE(ctx, ctx->d_work2, (size_t)n * n * n * sizeof(float));
    float *outs[3] = {map, half1, half2};
    for (int mode = 0; mode < 3; ++mode) {
        if (!outs[mode]) continue;
        filter_to_fft_kernel<<<grid_for(nvox, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
            a0, a1, np, xh, 1.f / (float)c.pad, ns, d_term + (size_t)mode * ns, mode, ctx->d_work1.as<float2>());
        KERNEL_CHECK(ctx);
        int rc = fft3_c2r_dev(ctx, ctx->d_work1.as<float2>(), ctx->d_work0.as<float>(), np);
        if (rc) return rc;
        float *dst = loc == CSPB_HOST ? ctx->d_work2.as<float>() : outs[mode];
        post_real_kernel<<<grid_for((long long)n * n * n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
            ctx->d_work0.as<float>(), np, n, 1.f / ((float)np * (float)np * (float)np), rad_a / c.pixel_size,
            20.f / c.pixel_size, dst);
        KERNEL_CHECK(ctx);
        if (loc == CSPB_HOST) {
            CU_TRY(ctx, cudaMemcpyAsync(outs[mode], dst, (size_t)n * n * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // d_work2 is reused by the next map
        }
    }
    if (loc == CSPB_HOST || stats) CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cspb_recon_end(cspb_ctx *ctx) {
    CSPB_ENTER(ctx);
    if (!ctx) return CSPB_E_ARG;
    ctx->d_acc[0].release();
    ctx->d_acc[1].release();
    ctx->d_raw[0].release();
    ctx->d_raw[1].release();
    ctx->raw_dirty = false;
    ctx->recon_ready = false;
    return 0;
}
