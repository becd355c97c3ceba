// search.cu — refine3d's global search (prompt 36): exhaustive orientation grid with an FFT
// cross-correlation shift search in a resolution-reduced box and parabolic sub-pixel peaks.
//
// Replaces the "global search yes" branch of external/cistem2/refine3d (closed; contract
// src/pyp/refine/frealign/frealign.py:3866-3871,3957-3962: angular step `refine_dang`, search
// ranges, high-res limit for search, 20 best matches refined locally).  Semantics:
// oracle/SEMANTICS.md §11; CPU restatement: oracle/cspb_oracle.c (orc_global_search).
//
// Structure (B200-first): the slices of the reference at every grid orientation are particle
// independent, so they are extracted once per launch into HBM (n_orient x n_ss complex, a few
// tens of MB, L2-resident); one CTA then owns one particle, keeps its CTF-multiplied search-band
// spectrum in shared memory and streams the orientations through a shared-memory batched
// small-box FFT (columns: complex Stockham lines; rows: two real rows per complex line, only the
// rows inside the shift window), finds the peak with a block arg-max and keeps the top-K hits.
#include <math.h>
#include <string.h>
#include "device_math.cuh"
#include "fft_smem.cuh"
#include "internal.cuh"

using namespace fftsm;

namespace {

struct SearchDev {
    int n, nb, n_ss, i_max;   // box, reduced box, search samples, largest i in the band
    int wx, wy;               // shift window half-widths in reduced pixels
    int n_rowpairs;           // row pairs covering dy in [-wy-1, wy+1]
    const int *ss_slot;       // slot of the sample in the packed band
    const int *ss_ij;         // i | j << 16
    const float *ss_mult;     // 2 for i > 0, 1 for the i = 0 column
};

// G[p][s] = F * ctf, C2[p][s] = ctf^2, A[p] = sum mult |F|^2
__global__ void search_prep_kernel(SearchDev S, const float2 *__restrict__ packed, int n_slots,
                                   const CtfCoef *__restrict__ ctf, float2 *__restrict__ G, float *__restrict__ C2,
                                   float *__restrict__ A) {
    __shared__ float red[64];
    const int p = blockIdx.x;
    const CtfCoef cc = ctf[p];
    float a = 0.f;
    for (int s = threadIdx.x; s < S.n_ss; s += blockDim.x) {
        const int ij = S.ss_ij[s];
        const float fi = (float)(short)(ij & 0xFFFF), fj = (float)(short)(ij >> 16);
        const float2 F = packed[(long long)p * n_slots + S.ss_slot[s]];
        const float c = -sinpif(ctf_chi(cc, fi, fj, fi * fi + fj * fj) * (1.f / CSPB_PI_F));
        G[(long long)p * S.n_ss + s] = make_float2(F.x * c, F.y * c);
        C2[(long long)p * S.n_ss + s] = c * c;
        a += S.ss_mult[s] * (F.x * F.x + F.y * F.y);
    }
    a = block_sum(a, red);
    if (threadIdx.x == 0) A[p] = a;
}

// Pall[o][s] = central slice at orientation o (no CTF)
__global__ void search_slices_kernel(SearchDev S, const float4 *__restrict__ ref4, int sx, int sy, int rc, float padf,
                                     const float *__restrict__ angles3, int n_orient, float2 *__restrict__ Pall) {
    const long long total = (long long)n_orient * S.n_ss;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(k / S.n_ss), s = (int)(k - (long long)o * S.n_ss);
        float r[9];
        euler_matrix(angles3[3 * o], angles3[3 * o + 1], angles3[3 * o + 2], r);
        const int ij = S.ss_ij[s];
        const float fi = (float)(short)(ij & 0xFFFF), fj = (float)(short)(ij >> 16);
        Pall[k] = gather_trilinear(ref4, sx, sy, rc, (r[0] * fi + r[1] * fj) * padf, (r[3] * fi + r[4] * fj) * padf,
                                   (r[6] * fi + r[7] * fj) * padf);
    }
}

struct Hit {
    float score, sx, sy;
    int orient;
};

// one CTA per particle; streams all orientations; keeps the K best hits (sorted, best first)
__global__ void __launch_bounds__(256) search_kernel(SearchDev S, Radices rad, const float2 *__restrict__ tw_g,
                                                     const float2 *__restrict__ G, const float *__restrict__ C2,
                                                     const float *__restrict__ A, const float2 *__restrict__ Pall,
                                                     int n_orient, int K, float shift_scale, Hit *__restrict__ hits) {
    extern __shared__ float2 sm[];
    const int nb = S.nb, pitch = fftsm::line_pitch(nb), nl = S.i_max + 1, nrp = S.n_rowpairs;
    float2 *tw = sm;                              // nb
    float2 *g = tw + nb;                          // n_ss
    float2 *bufa = g + S.n_ss;                    // nl * pitch   (columns, line = fixed i)
    float2 *bufb = bufa + (size_t)nl * pitch;     // nl * pitch
    float2 *rowa = bufb + (size_t)nl * pitch;     // nrp * pitch  (row pairs)
    float2 *rowb = rowa + (size_t)nrp * pitch;    // nrp * pitch
    float *c2m = reinterpret_cast<float *>(rowb + (size_t)nrp * pitch);  // n_ss: mult * ctf^2
    int *pos = reinterpret_cast<int *>(c2m + S.n_ss);                    // n_ss: index into bufa
    Hit *top = reinterpret_cast<Hit *>(pos + S.n_ss);                    // K
    __shared__ float red[64];
    const int tid = threadIdx.x, nt = blockDim.x, p = blockIdx.x;
    for (int k = tid; k < nb; k += nt) tw[k] = tw_g[k];
    for (int s = tid; s < S.n_ss; s += nt) {
        g[s] = G[(long long)p * S.n_ss + s];
        c2m[s] = C2[(long long)p * S.n_ss + s] * S.ss_mult[s];
        const int ij = S.ss_ij[s];
        const int i = (int)(short)(ij & 0xFFFF), j = (int)(short)(ij >> 16);
        pos[s] = i * pitch + fftsm::skew(j < 0 ? j + nb : j);
    }
    for (int k = tid; k < K; k += nt) {
        top[k].score = -1e30f; top[k].sx = 0.f; top[k].sy = 0.f; top[k].orient = -1;
    }
    const float Ap = A[p];
    __syncthreads();
    const int wxs = 2 * S.wx + 1, wys = 2 * S.wy + 1;
    for (int o = 0; o < n_orient; ++o) {
        const float2 *P = Pall + (long long)o * S.n_ss;
        for (int k = tid; k < nl * pitch; k += nt) bufa[k] = make_float2(0.f, 0.f);
        __syncthreads();
        float b = 0.f;
        for (int s = tid; s < S.n_ss; s += nt) {
            const float2 pv = __ldg(P + s), gv = g[s];
            bufa[pos[s]] = make_float2(gv.x * pv.x + gv.y * pv.y, gv.y * pv.x - gv.x * pv.y);  // G conj(P)
            b += c2m[s] * (pv.x * pv.x + pv.y * pv.y);
        }
        b = block_sum(b, red);  // also the barrier after the fill
        float2 *T = fft_lines_smem<+1>(bufa, bufb, pitch, nl, nb, rad, tw, tid, nt);
        // rows inside the window, two per complex line: pair q holds rows y0 = first + 2q, y0 + 1
        const int first = -S.wy - 1;
        for (int k = tid; k < nrp * pitch; k += nt) rowa[k] = make_float2(0.f, 0.f);
        __syncthreads();
        for (int k = tid; k < nrp * nl; k += nt) {
            const int q = k / nl, i = k - q * nl;
            const int ya = ((first + 2 * q) % nb + nb) % nb, yb = ((first + 2 * q + 1) % nb + nb) % nb;
            float2 fa = T[i * pitch + fftsm::skew(ya)], fb = T[i * pitch + fftsm::skew(yb)];
            if (i == 0) { fa.y = 0.f; fb.y = 0.f; }  // the DC column of a real image is real
            rowa[q * pitch + fftsm::skew(i)] = make_float2(fa.x - fb.y, fa.y + fb.x);
            if (i > 0) rowa[q * pitch + fftsm::skew(nb - i)] = make_float2(fa.x + fb.y, -fa.y + fb.x);
        }
        __syncthreads();
        float2 *R = fft_lines_smem<+1>(rowa, rowb, pitch, nrp, nb, rad, tw, tid, nt);
        // R[q][x].x = row first+2q, .y = row first+2q+1
        float best = -1e30f;
        int bidx = 0x7fffffff;
        for (int k = tid; k < wxs * wys; k += nt) {
            const int dy = k / wxs - S.wy, dx = k % wxs - S.wx;
            const int rr = dy - first, q = rr >> 1;
            const float2 v = R[q * pitch + fftsm::skew((dx + nb) % nb)];
            const float val = (rr & 1) ? v.y : v.x;
            if (val > best) { best = val; bidx = k; }
        }
        // block arg-max, ties -> smallest index
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        __shared__ float wbest[8];
        __shared__ int widx[8];
        if ((tid & 31) == 0) { wbest[tid >> 5] = best; widx[tid >> 5] = bidx; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (nt >> 5); ++w)
                if (wbest[w] > best || (wbest[w] == best && widx[w] < bidx)) { best = wbest[w]; bidx = widx[w]; }
            // parabolic sub-pixel fit along x and y
            const int dy = bidx / wxs - S.wy, dx = bidx % wxs - S.wx;
            auto at = [&](int ddx, int ddy) {
                const int rr = ddy - first, q = rr >> 1;
                const float2 v = R[q * pitch + fftsm::skew((ddx + nb) % nb)];
                return (rr & 1) ? v.y : v.x;
            };
            const float v0 = best, xm = at(dx - 1, dy), xp = at(dx + 1, dy), ym = at(dx, dy - 1), yp = at(dx, dy + 1);
            float ox = 0.f, oy = 0.f, peak = v0;
            const float cx = xm - 2.f * v0 + xp, cy = ym - 2.f * v0 + yp;
            if (cx < 0.f) { ox = fminf(fmaxf(0.5f * (xm - xp) / cx, -0.5f), 0.5f); peak -= 0.25f * (xm - xp) * ox; }
            if (cy < 0.f) { oy = fminf(fmaxf(0.5f * (ym - yp) / cy, -0.5f), 0.5f); peak -= 0.25f * (ym - yp) * oy; }
            const float den = Ap * b;
            const float score = den > 0.f ? 100.f * peak * rsqrtf(den) : 0.f;
            // insert into the sorted top-K (best first; equal scores keep the earlier orientation)
            if (score > top[K - 1].score) {
                int k = K - 1;
                while (k > 0 && score > top[k - 1].score) { top[k] = top[k - 1]; --k; }
                top[k].score = score;
                top[k].sx = ((float)dx + ox) * shift_scale;
                top[k].sy = ((float)dy + oy) * shift_scale;
                top[k].orient = o;
            }
        }
        __syncthreads();
    }
    for (int k = tid; k < K; k += nt) hits[(long long)p * K + k] = top[k];
}

int grid_for(long long total, int block, int sm) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)sm * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int pick_reduced_box(int need, int n) {
    const int sizes[] = {16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024};
    for (int s : sizes)
        if (s >= need) return s < n ? s : n;
    return n;
}

}  // namespace

// Run the global search over all loaded images.  angles3: n_orient x (psi, theta, phi) degrees on
// the device.  hits_out (device): n_images x K {score, shift_x A, shift_y A, orientation index}.
int search_enqueue(cspb_ctx *ctx, const CtfCoef *d_ctf, const float *d_angles3, int n_orient, int K, void *d_hits) {
    const cspb_refine_cfg &c = ctx->rcfg;
    const BandPlan &pl = ctx->plan;
    const int n = c.box, P = ctx->n_images;
    const float npx = (float)n * c.pixel_size;
    float r_s = c.search_high_res > 0.f ? npx / c.search_high_res : pl.r_hi;
    if (r_s > pl.r_hi) r_s = pl.r_hi;
    if (r_s < pl.r_lo + 2.f) r_s = fminf(pl.r_lo + 2.f, pl.r_hi);
    // search samples = band samples with r <= r_s
    std::vector<int> slot, ij;
    std::vector<float> mult;
    int i_max = 0;
    for (int s = 0; s < pl.n_slots; ++s) {
        const int32_t v = pl.slot_ij[s];
        const int i = (int)(short)(v & 0xFFFF), j = (int)(short)(v >> 16);
        if (i == CSPB_DUMMY_I) continue;
        if ((float)(i * i + j * j) > r_s * r_s) continue;
        slot.push_back(s);
        ij.push_back(v);
        mult.push_back(i > 0 ? 2.f : 1.f);
        if (i > i_max) i_max = i;
        if (abs(j) > i_max) i_max = abs(j);
    }
    const int n_ss = (int)slot.size();
    if (n_ss == 0) return cspb_fail(ctx, CSPB_E_ARG, "empty search band");
    const int nb = pick_reduced_box(2 * (i_max + 2), n);
    Radices rad;
    if (!factor(nb, rad)) return cspb_fail(ctx, CSPB_E_ARG, "reduced box %d unsupported", nb);
    const float red = (float)nb / (float)n;  // reduced pixels per full pixel
    int wx = (int)ceilf(c.search_range_x / c.pixel_size * red), wy = (int)ceilf(c.search_range_y / c.pixel_size * red);
    const int wmax = nb / 2 - 2;
    if (wx < 1) wx = 1;
    if (wy < 1) wy = 1;
    if (wx > wmax) wx = wmax;
    if (wy > wmax) wy = wmax;
    SearchDev S;
    S.n = n; S.nb = nb; S.n_ss = n_ss; S.i_max = i_max; S.wx = wx; S.wy = wy;
    S.n_rowpairs = (2 * wy + 3 + 1) / 2;
    DevBuf d_tab;
    RESERVE(ctx, d_tab, (size_t)n_ss * 12);
    int *d_slot = d_tab.as<int>(), *d_ij = d_slot + n_ss;
    float *d_mult = reinterpret_cast<float *>(d_ij + n_ss);
    CU_TRY(ctx, cudaMemcpyAsync(d_slot, slot.data(), n_ss * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(d_ij, ij.data(), n_ss * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(d_mult, mult.data(), n_ss * 4, cudaMemcpyHostToDevice, ctx->stream));
    S.ss_slot = d_slot; S.ss_ij = d_ij; S.ss_mult = d_mult;
    const float2 *tw;
    int rc = fft_get_twiddles(ctx, nb, &tw);
    if (rc) return rc;
    RESERVE(ctx, ctx->d_work0, (size_t)P * n_ss * 12 + (size_t)P * 4);
    float2 *G = ctx->d_work0.as<float2>();
    float *C2 = reinterpret_cast<float *>(G + (size_t)P * n_ss);
    float *A = C2 + (size_t)P * n_ss;
    RESERVE(ctx, ctx->d_work2, (size_t)n_orient * n_ss * sizeof(float2));
    float2 *Pall = ctx->d_work2.as<float2>();
    search_prep_kernel<<<P, 128, 0, ctx->stream>>>(S, ctx->d_packed.as<float2>(), pl.n_slots, d_ctf, G, C2, A);
    KERNEL_CHECK(ctx);
    search_slices_kernel<<<grid_for((long long)n_orient * n_ss, 256, ctx->sm_count), 256, 0, ctx->stream>>>(
        S, ctx->ref.d_ref4.as<float4>(), ctx->ref.sx, ctx->ref.sy, ctx->ref.rc, (float)ctx->ref.pad, d_angles3, n_orient, Pall);
    KERNEL_CHECK(ctx);
    const int pitch = fftsm::line_pitch(nb), nl = i_max + 1;
    const size_t smem = ((size_t)nb + n_ss + 2 * (size_t)nl * pitch + 2 * (size_t)S.n_rowpairs * pitch) * sizeof(float2) +
                        (size_t)n_ss * 8 + (size_t)K * sizeof(Hit);
    if (smem > 220 * 1024) return cspb_fail(ctx, CSPB_E_ARG, "global-search box %d needs %zu B of shared memory", nb, smem);
    if (smem > 48 * 1024)
        CU_TRY(ctx, cudaFuncSetAttribute(search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float shift_scale = c.pixel_size / red;  // reduced pixel -> Angstrom
    search_kernel<<<P, 256, smem, ctx->stream>>>(S, rad, tw, G, C2, A, Pall, n_orient, K, shift_scale,
                                                 reinterpret_cast<Hit *>(d_hits));
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // d_tab goes out of scope
    return 0;
}
