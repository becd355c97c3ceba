/*
 * cspb_oracle_fast.c — the SAME algorithms as cspb_oracle.c (scorer of SEMANTICS.md §6, analytic optimiser of §7c,
 * reconstruct3d insertion of §8) written the way a production CPU code would be: TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * cspb_oracle.c is a clarity-first restatement (it recomputes the CTF and walks the whole half plane per evaluation,
 * resolves Friedel symmetry and the FFT-order wrap per trilinear corner, allocates per evaluation and inserts every
 * symmetry operator literally).  A CPU/GPU ratio against that port says little, so bench.py times THIS file as the CPU
 * arm and quotes the naive port beside it.  What differs, results unchanged (tests/test_cpu_oracle.py compares the two):
 *   - the band is a precomputed sample list (i, j, ring) shared by all particles;
 *   - the CTF of a particle is evaluated once per sample, not once per sample and evaluation;
 *   - the reference is a centred crop (x in [0, R], y, z in [-R, R]): one Friedel flip per SAMPLE, direct corner indexing;
 *   - no allocation inside an evaluation; sincosf for the phase ramp;
 *   - reconstruct3d inserts the right-coset representatives of the lattice-preserving subgroup only (all 24 operators
 *     of O permute the voxel lattice) and applies the lattice operators once to the accumulated volume — exact, the
 *     same decomposition the CUDA path uses (recon.cu, lattice_sym_kernel).
 */
#define _GNU_SOURCE /* sincosf */
#include <math.h>
#include "cspb_oracle.c"

/* ---- single-precision iterative radix-2 FFT (n a power of two; other sizes keep the reference transform) ---- */
static float *g_tw = NULL;  /* cos, sin of 2 pi k / n, k < n/2 */
static int g_tw_n = 0;
static int *g_rev = NULL;
#pragma omp threadprivate(g_tw, g_tw_n, g_rev)

static void ffast_plan(int n) {
    if (g_tw_n == n) return;
    free(g_tw); free(g_rev);
    g_tw = (float *)malloc(sizeof(float) * n);
    g_rev = (int *)malloc(sizeof(int) * n);
    for (int k = 0; k < n / 2; ++k) { g_tw[2 * k] = (float)cos(2.0 * M_PI * k / n); g_tw[2 * k + 1] = (float)sin(2.0 * M_PI * k / n); }
    int bits = 0;
    while ((1 << bits) < n) ++bits;
    for (int k = 0; k < n; ++k) {
        int r = 0;
        for (int b = 0; b < bits; ++b) if (k & (1 << b)) r |= 1 << (bits - 1 - b);
        g_rev[k] = r;
    }
    g_tw_n = n;
}

/* in-place complex FFT of one line (stride 1), sign = -1 forward, +1 inverse (unnormalised) */
static void ffast_line(float *d, int n, int sign) {
    for (int k = 0; k < n; ++k) {
        const int r = g_rev[k];
        if (r > k) { float a = d[2 * k], b = d[2 * k + 1]; d[2 * k] = d[2 * r]; d[2 * k + 1] = d[2 * r + 1]; d[2 * r] = a; d[2 * r + 1] = b; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int s0 = 0; s0 < n; s0 += len)
            for (int k = 0; k < half; ++k) {
                const float wr = g_tw[2 * k * step], wi = (float)sign * g_tw[2 * k * step + 1];
                float *a = d + 2 * (s0 + k), *b = d + 2 * (s0 + k + half);
                const float tr = b[0] * wr - b[1] * wi, ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr; b[1] = a[1] - ti;
                a[0] += tr; a[1] += ti;
            }
    }
}

static void ffast_r2c(const float *img, int n, float *out_c) {
    if (n & (n - 1)) { orc_fft2_r2c(img, n, out_c); return; }
    ffast_plan(n);
    const int nh = n / 2 + 1;
    float *row = (float *)malloc(sizeof(float) * 2 * n), *col = (float *)malloc(sizeof(float) * 2 * n);
    for (int y = 0; y < n; ++y) {  /* rows: real input as complex, keep x <= n/2 */
        for (int x = 0; x < n; ++x) { row[2 * x] = img[(size_t)y * n + x]; row[2 * x + 1] = 0.f; }
        ffast_line(row, n, -1);
        memcpy(out_c + 2 * (size_t)y * nh, row, sizeof(float) * 2 * nh);
    }
    for (int x = 0; x < nh; ++x) {  /* columns */
        for (int y = 0; y < n; ++y) { col[2 * y] = out_c[2 * ((size_t)y * nh + x)]; col[2 * y + 1] = out_c[2 * ((size_t)y * nh + x) + 1]; }
        ffast_line(col, n, -1);
        for (int y = 0; y < n; ++y) { out_c[2 * ((size_t)y * nh + x)] = col[2 * y]; out_c[2 * ((size_t)y * nh + x) + 1] = col[2 * y + 1]; }
    }
    free(row); free(col);
}

static void ffast_c2r(const float *in_c, int n, float *out) {
    if (n & (n - 1)) { orc_fft2_c2r(in_c, n, out); return; }
    ffast_plan(n);
    const int nh = n / 2 + 1;
    float *work = (float *)malloc(sizeof(float) * 2 * (size_t)n * nh), *line = (float *)malloc(sizeof(float) * 2 * n);
    memcpy(work, in_c, sizeof(float) * 2 * (size_t)n * nh);
    for (int x = 0; x < nh; ++x) {
        for (int y = 0; y < n; ++y) { line[2 * y] = work[2 * ((size_t)y * nh + x)]; line[2 * y + 1] = work[2 * ((size_t)y * nh + x) + 1]; }
        ffast_line(line, n, +1);
        for (int y = 0; y < n; ++y) { work[2 * ((size_t)y * nh + x)] = line[2 * y]; work[2 * ((size_t)y * nh + x) + 1] = line[2 * y + 1]; }
    }
    for (int y = 0; y < n; ++y) {  /* rows: Hermitian extension, inverse, real part */
        for (int x = 0; x < nh; ++x) { line[2 * x] = work[2 * ((size_t)y * nh + x)]; line[2 * x + 1] = work[2 * ((size_t)y * nh + x) + 1]; }
        for (int x = nh; x < n; ++x) { line[2 * x] = line[2 * (n - x)]; line[2 * x + 1] = -line[2 * (n - x) + 1]; }
        ffast_line(line, n, +1);
        for (int x = 0; x < n; ++x) out[(size_t)y * n + x] = line[2 * x];
    }
    free(work); free(line);
}

/* particle preprocessing / insertion of the optimised leg: the reference code with the fast transforms */
void orc_prepare_image_fast(const float *img, const orc_refine_cfg *cfg, const float *noise_curve, const float *ring_weights, float *spec) {
    g_fft2_r2c = ffast_r2c; g_fft2_c2r = ffast_c2r;
    orc_prepare_image(img, cfg, noise_curve, ring_weights, spec);
    g_fft2_r2c = orc_fft2_r2c; g_fft2_c2r = orc_fft2_c2r;
}

typedef struct {
    int n_s, ring_max;
    int *i, *j, *bin;
} fband;

static fband *fband_make(const orc_refine_cfg *cfg) {
    const int n = cfg->box;
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    fband *b = (fband *)calloc(1, sizeof(fband));
    const int cap = orc_band_count(cfg);
    b->i = (int *)malloc(sizeof(int) * cap); b->j = (int *)malloc(sizeof(int) * cap); b->bin = (int *)malloc(sizeof(int) * cap);
    for (int j = -n / 2; j < n / 2; ++j)   /* the order of cspb_oracle.c: identical summation order */
        for (int i = 0; i <= n / 2; ++i) {
            const float r2 = (float)(i * i + j * j);
            if (r2 < lo * lo || r2 > hi * hi) continue;
            b->i[b->n_s] = i; b->j[b->n_s] = j; b->bin[b->n_s] = (int)sqrtf(r2);
            if (b->bin[b->n_s] > b->ring_max) b->ring_max = b->bin[b->n_s];
            b->n_s++;
        }
    return b;
}
static void fband_free(fband *b) { if (b) { free(b->i); free(b->j); free(b->bin); free(b); } }

typedef struct {
    int R, sx, sy;   /* crop radius; x extent R + 2; y, z extent 2R + 3 */
    float *v;        /* complex [z][y][x] */
    int pad;
} fcrop;

static fcrop *fcrop_make(const orc_ref *r, float r_hi) {
    fcrop *c = (fcrop *)calloc(1, sizeof(fcrop));
    c->pad = r->pad;
    c->R = (int)ceilf(r->pad * r_hi) + 2;
    c->sx = c->R + 2; c->sy = 2 * c->R + 3;
    c->v = (float *)malloc(sizeof(float) * 2 * (size_t)c->sx * c->sy * c->sy);
    for (int z = -c->R - 1; z <= c->R + 1; ++z)
        for (int y = -c->R - 1; y <= c->R + 1; ++y)
            for (int x = 0; x < c->sx; ++x) {
                float *d = c->v + 2 * (((size_t)(z + c->R + 1) * c->sy + (y + c->R + 1)) * c->sx + x);
                ref_at(r, x, y, z, d, d + 1);
            }
    return c;
}
static void fcrop_free(fcrop *c) { if (c) { free(c->v); free(c); } }

/* the eight corners at (x >= 0, y, z) */
static inline const float *fcrop_at(const fcrop *c, int x, int y, int z) {
    return c->v + 2 * (((size_t)(z + c->R + 1) * c->sy + (y + c->R + 1)) * c->sx + x);
}

/* value (+ derivatives when dnum != NULL) of one pose; ctf[] = the particle's CTF per band sample.  Same arithmetic
 * as score_cut / score_grad_cut.  ring[] = scratch of (ring_max + 1) * 8 floats. */
static float fast_eval(const fcrop *c, const fband *b, const float *spec, int n, const float *ctfv, const float *pose6, float pixel,
                       const orc_refine_cfg *cfg, int ring_cut, float *out4, float *dnum, float *dB, float *jtj, float *ring) {
    const int nh = n / 2 + 1, nr = b->ring_max + 1;
    const int grad = dnum != NULL;
    memset(ring, 0, sizeof(float) * (size_t)nr * 8);
    float suma = 0.f, sumb = 0.f, m[9], dth[6];
    orc_euler_matrix(pose6[0], pose6[1], pose6[2], m);
    if (grad) {
        euler_derivatives(pose6[0], pose6[1], pose6[2], dth);
        for (int a = 0; a < 3; ++a) dB[a] = 0.f;
        for (int a = 0; a < NJ; ++a) jtj[a] = 0.f;
    }
    const float k2 = 2.f * PI_F / ((float)n * pixel), d2r = PI_F / 180.f, pad = (float)c->pad;
    for (int s = 0; s < b->n_s; ++s) {
        const int bin = b->bin[s];
        if (bin > ring_cut) continue;
        const int i = b->i[s], j = b->j[s];
        const int jj = j < 0 ? j + n : j;
        const float fr = spec[2 * ((size_t)jj * nh + i)], fim = spec[2 * ((size_t)jj * nh + i) + 1];
        const float fi = (float)i, fj = (float)j;
        float x = (m[0] * fi + m[1] * fj) * pad, y = (m[3] * fi + m[4] * fj) * pad, z = (m[6] * fi + m[7] * fj) * pad;
        float vel[3][3];
        if (grad) {
            vel[0][0] = (m[1] * fi - m[0] * fj) * pad; vel[0][1] = (m[4] * fi - m[3] * fj) * pad; vel[0][2] = (m[7] * fi - m[6] * fj) * pad;
            vel[1][0] = (dth[0] * fi + dth[1] * fj) * pad; vel[1][1] = (dth[2] * fi + dth[3] * fj) * pad; vel[1][2] = (dth[4] * fi + dth[5] * fj) * pad;
            vel[2][0] = -y; vel[2][1] = x; vel[2][2] = 0.f;
        }
        const float sg = x < 0.f ? -1.f : 1.f;
        x *= sg; y *= sg; z *= sg;
        const int x0 = (int)floorf(x), y0 = (int)floorf(y), z0 = (int)floorf(z);
        const float fx = x - x0, fy = y - y0, fz = z - z0;
        const float *c000 = fcrop_at(c, x0, y0, z0), *c010 = fcrop_at(c, x0, y0 + 1, z0), *c001 = fcrop_at(c, x0, y0, z0 + 1),
                    *c011 = fcrop_at(c, x0, y0 + 1, z0 + 1);
        float v[2], gx[2], gy[2], gz[2];
        for (int k = 0; k < 2; ++k) {
            const float d00 = c000[2 + k] - c000[k], d01 = c010[2 + k] - c010[k], d10 = c001[2 + k] - c001[k], d11 = c011[2 + k] - c011[k];
            const float a00 = c000[k] + fx * d00, a01 = c010[k] + fx * d01, a10 = c001[k] + fx * d10, a11 = c011[k] + fx * d11;
            const float dx0 = d00 + fy * (d01 - d00), dx1 = d10 + fy * (d11 - d10);
            gx[k] = dx0 + fz * (dx1 - dx0);
            const float dy0 = a01 - a00, dy1 = a11 - a10;
            const float v0 = a00 + fy * dy0, v1 = a10 + fy * dy1;
            gy[k] = dy0 + fz * (dy1 - dy0);
            gz[k] = v1 - v0;
            v[k] = v0 + fz * gz[k];
        }
        const float ctf = ctfv[s];
        const float pr = ctf * v[0], pi = ctf * sg * v[1];
        const float ph = (fi * pose6[3] + fj * pose6[4]) * k2;
        float sn, cs;
        sincosf(ph, &sn, &cs);
        const float gr = fr * cs - fim * sn, gi = fr * sn + fim * cs;
        float *rg = ring + (size_t)bin * 8;
        rg[0] += gr * pr + gi * pi;
        suma += fr * fr + fim * fim;
        sumb += pr * pr + pi * pi;
        if (grad) {
            float dp[NG][2];
            for (int a = 0; a < 3; ++a) {
                const float dre = gx[0] * vel[a][0] + gy[0] * vel[a][1] + gz[0] * vel[a][2];
                const float dim = gx[1] * vel[a][0] + gy[1] * vel[a][1] + gz[1] * vel[a][2];
                dp[a][0] = ctf * sg * dre * d2r;
                dp[a][1] = ctf * dim * d2r;
            }
            dp[3][0] = k2 * fi * pi; dp[3][1] = -k2 * fi * pr;
            dp[4][0] = k2 * fj * pi; dp[4][1] = -k2 * fj * pr;
            for (int a = 0; a < NG; ++a) rg[1 + a] += gr * dp[a][0] + gi * dp[a][1];
            rg[6] += fr * fr + fim * fim;
            rg[7] += pr * pr + pi * pi;
            for (int a = 0; a < 3; ++a) dB[a] += 2.f * (pr * dp[a][0] + pi * dp[a][1]);
            int t = 0;
            for (int a = 0; a < NG; ++a)
                for (int q = a; q < NG; ++q) jtj[t++] += dp[a][0] * dp[q][0] + dp[a][1] * dp[q][1];
        }
    }
    const int limit = cfg->signed_cc_limit > 0.f ? (int)floorf((float)n * cfg->pixel_size / cfg->signed_cc_limit) : 0x7fffffff;
    float num = 0.f, xs = 0.f;
    if (grad) for (int a = 0; a < NG; ++a) dnum[a] = 0.f;
    for (int q = 0; q < nr; ++q) {
        const float *rg = ring + (size_t)q * 8;
        xs += rg[0];
        num += (q > limit) ? fabsf(rg[0]) : rg[0];
        if (grad) {
            float sgn = 1.f;
            if (q > limit) {
                const float w = rg[0] * rg[0] + LM_SOFT * LM_SOFT * rg[6] * rg[7];
                sgn = w > 0.f ? rg[0] / sqrtf(w) : 0.f;
            }
            for (int a = 0; a < NG; ++a) dnum[a] += sgn * rg[1 + a];
        }
    }
    out4[0] = num; out4[1] = xs; out4[2] = suma; out4[3] = sumb;
    const float den = suma * sumb;
    return den > 0.f ? 100.f * num / sqrtf(den) : 0.f;
}

/* orc_refine_local with optimizer 0 (analytic, §7c) on the fast evaluation; the stencil optimiser, the defocus
 * refinement and the focus mask are left to cspb_oracle.c */
long long orc_refine_local_fast(const orc_ref *r, const float *specs, orc_row *rows, int n_img, const orc_refine_cfg *cfg) {
    if (cfg->optimizer != 0 || cfg->refine_defocus || cfg->focus_radius > 0.f) return orc_refine_local(r, specs, rows, n_img, cfg);
    const int n = cfg->box, nh = n / 2 + 1;
    const int freem[NP] = {cfg->refine_psi, cfg->refine_theta, cfg->refine_phi, cfg->refine_x, cfg->refine_y, 0};
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    fband *b = fband_make(cfg);
    fcrop *c = fcrop_make(r, hi);
    const int nband = b->n_s;
    int n_free = 0;
    for (int m = 0; m < NG; ++m) n_free += freem[m] ? 1 : 0;
    const int iters = n_free > 0 ? (cfg->local_iterations > 0 ? cfg->local_iterations : 8) : 0;
    long long evals = 0;
#pragma omp parallel reduction(+ : evals)
    {
        float *ctfv = (float *)malloc(sizeof(float) * (size_t)b->n_s);
        float *ring = (float *)malloc(sizeof(float) * (size_t)(b->ring_max + 1) * 8);
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < n_img; ++k) {
            const float *spec = specs + 2 * (size_t)k * n * nh;
            orc_row *row = &rows[k];
            const ctfc cc = ctf_make(row, n);
            for (int s = 0; s < b->n_s; ++s) ctfv[s] = ctf_eval(&cc, b->i[s], b->j[s], 0.f);
            float x[NP] = {row->psi, row->theta, row->phi, row->x_shift, row->y_shift, 0.f}, o4[4];
            const float x_start[NP] = {x[0], x[1], x[2], x[3], x[4], x[5]};
            for (int it = 0; it < iters; ++it) {
                int ring_cut;
                const float f = lm_stage(it, iters, lo, hi, &ring_cut);
                const float h_ang = 0.35f * 57.29578f * f / hi, h_shift = 0.07f * (float)n * f / hi * cfg->pixel_size;
                const float trust[NG] = {16.f * h_ang, 16.f * h_ang, 16.f * h_ang, 16.f * h_shift, 16.f * h_shift};
                float dnum[NG], dB[3], jtj[NJ], d[NG], slope, q[NP];
                fast_eval(c, b, spec, n, ctfv, x, row->pixel_size, cfg, ring_cut, o4, dnum, dB, jtj, ring);
                const float f0 = lm_step(o4, dnum, dB, jtj, freem, trust, cfg, row, x, d, &slope);
                memcpy(q, x, sizeof q);
                for (int m = 0; m < NG; ++m) q[m] = x[m] + d[m];
                const float f1 = fast_eval(c, b, spec, n, ctfv, q, row->pixel_size, cfg, ring_cut, o4, NULL, NULL, NULL, ring) * 0.01f - prior_pen(cfg, row, q);
                evals += 2;
                const float t = lm_line(f0, slope, f1);
                for (int m = 0; m < NG; ++m) x[m] += t * d[m];
            }
            float o4s[4];
            float sc = fast_eval(c, b, spec, n, ctfv, x, row->pixel_size, cfg, 0x7fffffff, o4, NULL, NULL, NULL, ring);
            const float sc_start = fast_eval(c, b, spec, n, ctfv, x_start, row->pixel_size, cfg, 0x7fffffff, o4s, NULL, NULL, NULL, ring);
            evals += 2;
            const float obj = sc * 0.01f - prior_pen(cfg, row, x), obj_start = sc_start * 0.01f - prior_pen(cfg, row, x_start);
            if (obj < obj_start) { memcpy(x, x_start, sizeof x_start); memcpy(o4, o4s, sizeof o4s); sc = sc_start; }
            write_row(row, x, sc, o4, nband, 0);
        }
        free(ctfv);
        free(ring);
    }
    fband_free(b);
    fcrop_free(c);
    return evals;
}

/* ---- reconstruct3d with deferred lattice symmetry ---- */
static int is_lattice_op(const float *m) {
    for (int k = 0; k < 9; ++k) {
        const float r = roundf(m[k]);
        if (fabsf(m[k] - r) > 1e-4f || fabsf(r) > 1.f) return 0;
    }
    return 1;
}

/* value of the complete (x = 0 plane folded) raw accumulator at lattice point (x >= 0, y, z) */
static void raw_at(const float *raw, int np, int xh, int x, int y, int z, float *o3) {
    const int c = np / 2;
    const float *a = raw + 4 * (((size_t)(z + c) * np + (y + c)) * xh + x);
    o3[0] = a[0]; o3[1] = a[1]; o3[2] = a[2];
    if (x == 0 && y != -c && z != -c) {
        const float *m = raw + 4 * (((size_t)(-z + c) * np + (-y + c)) * xh);
        o3[0] += m[0]; o3[1] -= m[1]; o3[2] += m[2];
    }
}

/* acc[v] += sum_h raw[h^-1 v] over the lattice operators (transposes = inverses given in ht); destinations on the
 * x = 0 plane are stored halved so that the later Friedel folding of that plane restores the total (recon.cu) */
static void lattice_symmetrize(const float *raw, float *acc, int np, int xh, const int *ht, int n_lat) {
    const int c = np / 2;
#pragma omp parallel for schedule(static)
    for (int zz = 0; zz < np; ++zz)
        for (int yy = 0; yy < np; ++yy)
            for (int x = 0; x < xh; ++x) {
                const int y = yy - c, z = zz - c;
                float s[3] = {0.f, 0.f, 0.f};
                for (int h = 0; h < n_lat; ++h) {
                    const int *m = ht + 9 * h;
                    int ux = m[0] * x + m[1] * y + m[2] * z, uy = m[3] * x + m[4] * y + m[5] * z, uz = m[6] * x + m[7] * y + m[8] * z;
                    float sg = 1.f;
                    if (ux < 0) { ux = -ux; uy = -uy; uz = -uz; sg = -1.f; }
                    if (ux > c || uy < -c || uy >= c || uz < -c || uz >= c) continue;
                    float v[3];
                    raw_at(raw, np, xh, ux, uy, uz, v);
                    s[0] += v[0]; s[1] += sg * v[1]; s[2] += v[2];
                }
                if (x == 0 && y != -c && z != -c) { s[0] *= 0.5f; s[1] *= 0.5f; s[2] *= 0.5f; }
                float *d = acc + 4 * (((size_t)zz * np + yy) * xh + x);
                d[0] += s[0]; d[1] += s[1]; d[2] += s[2];
            }
}

/* orc_recon_insert with G = H R: only the coset representatives R are inserted per sample, into a raw accumulator pair
 * that lives next to rc's; `finish` != 0 (or a later orc_recon_finish_fast) applies H, the operators that permute the
 * lattice, once to the raw sums and folds them into rc's accumulators.  In production the fold belongs to the merge
 * (once per reconstruction, after the dumps of all ranges are summed), which is why bench.py leaves it outside the
 * timed region like the transform of the reference. */
static float *g_raw[2] = {NULL, NULL};
static const orc_recon *g_raw_owner = NULL;
static int *g_ht = NULL, g_nh = 0;

/* drop the pending raw sums without folding them (bench.py: the fold is outside the timed region and its result unused) */
void orc_recon_discard_fast(orc_recon *rc) {
    if (g_raw_owner != rc || !g_raw[0]) return;
    for (int h = 0; h < 2; ++h) { free(g_raw[h]); g_raw[h] = NULL; }
    free(g_ht);
    g_ht = NULL;
    g_raw_owner = NULL;
}

void orc_recon_finish_fast(orc_recon *rc) {
    if (g_raw_owner != rc || !g_raw[0]) return;
    for (int h = 0; h < 2; ++h) {
        lattice_symmetrize(g_raw[h], rc->acc[h], rc->np, rc->xh, g_ht, g_nh);
        free(g_raw[h]);
        g_raw[h] = NULL;
    }
    free(g_ht);
    g_ht = NULL;
    g_raw_owner = NULL;
}

void orc_recon_insert_fast(orc_recon *rc, const float *imgs, const orc_row *rows, int count, const float *sym, int n_sym, int finish) {
    const float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!sym || n_sym < 1) { sym = id; n_sym = 1; }
    int *H = (int *)malloc(sizeof(int) * n_sym), *Rr = (int *)malloc(sizeof(int) * n_sym), nH = 0, nR = 0;
    for (int g = 0; g < n_sym; ++g)
        if (is_lattice_op(sym + 9 * g)) H[nH++] = g;
    if (nH > 1)
        for (int g = 0; g < n_sym; ++g) {
            int covered = 0;
            for (int t = 0; t < nR && !covered; ++t) {
                float q[9];
                const float *G = sym + 9 * g, *Rm = sym + 9 * Rr[t];
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) q[3 * a + b] = G[3 * a] * Rm[3 * b] + G[3 * a + 1] * Rm[3 * b + 1] + G[3 * a + 2] * Rm[3 * b + 2];
                if (!is_lattice_op(q)) continue;
                for (int h = 0; h < nH && !covered; ++h) {
                    float d = 0.f;
                    for (int k = 0; k < 9; ++k) d = fmaxf(d, fabsf(q[k] - sym[9 * H[h] + k]));
                    if (d < 1e-3f) covered = 1;
                }
            }
            if (!covered) Rr[nR++] = g;
        }
    if (nH <= 1 || nH * nR != n_sym) { /* no useful decomposition: literal insertion */
        g_fft2_r2c = ffast_r2c;
        orc_recon_insert(rc, imgs, rows, count, sym, n_sym);
        g_fft2_r2c = orc_fft2_r2c;
        free(H); free(Rr);
        return;
    }
    if (g_raw_owner && g_raw_owner != rc) orc_recon_finish_fast((orc_recon *)g_raw_owner);
    const size_t nf = 4 * (size_t)rc->xh * rc->np * rc->np;
    if (!g_raw[0]) {
        for (int h = 0; h < 2; ++h) g_raw[h] = (float *)calloc(nf, sizeof(float));
        g_raw_owner = rc;
        g_nh = nH;
        g_ht = (int *)malloc(sizeof(int) * 9 * nH);
        for (int h = 0; h < nH; ++h)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) g_ht[9 * h + 3 * a + b] = (int)roundf(sym[9 * H[h] + 3 * b + a]);
    }
    float *lit = (float *)malloc(sizeof(float) * 9 * nR);
    for (int t = 0; t < nR; ++t) memcpy(lit + 9 * t, sym + 9 * Rr[t], sizeof(float) * 9);
    float *keep[2] = {rc->acc[0], rc->acc[1]};
    rc->acc[0] = g_raw[0]; rc->acc[1] = g_raw[1];
    g_fft2_r2c = ffast_r2c;
    orc_recon_insert(rc, imgs, rows, count, lit, nR);
    g_fft2_r2c = orc_fft2_r2c;
    rc->acc[0] = keep[0]; rc->acc[1] = keep[1];
    free(lit); free(H); free(Rr);
    if (finish) orc_recon_finish_fast(rc);
}
