/*
 * cspb_oracle.c — CPU restatement of the CSP refine3d / reconstruct3d hot path (plain C).
 * TEST INFRASTRUCTURE ONLY — see the header of cspb_oracle.h.  "Parity unpinned": the
 * reference arithmetic is in closed LFS binaries; definitions are in oracle/SEMANTICS.md.
 *
 * Deliberately written differently from the CUDA path so the two check each other:
 *   - natural (j, i) half-plane loops with a per-pixel bin index, the shape of cisTEM's
 *     Image::GetWeightedCorrelationWithImage (published source; contract at
 *     src/pyp/refine/frealign/frealign.py:3918-3994), not the GPU's polar-patch band plan;
 *   - explicit 8-weight trilinear interpolation on a plain FFT-ordered padded volume;
 *   - a recursive mixed-radix FFT in double precision.
 * Arithmetic is fp32 where cisTEM's is (images, CTF, correlation sums), as the reference is.
 */
#include "cspb_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define PI_D 3.14159265358979323846
#define PI_F 3.14159265358979323846f

/* ================================================================ FFT (double, mixed radix) */
typedef struct { double re, im; } cd;

static void fft_rec(const cd *in, cd *out, int n, int stride, int sign, const cd *tw, int tw_n) {
    /* decimation in time; radix = smallest prime factor of n; tw = exp(sign*2 pi i k / tw_n) */
    if (n == 1) { out[0] = in[0]; return; }
    int p = 2;
    while (n % p) ++p;
    const int m = n / p;
    for (int q = 0; q < p; ++q) fft_rec(in + (size_t)q * stride, out + (size_t)q * m, m, stride * p, sign, tw, tw_n);
    cd *tmp = (cd *)malloc(sizeof(cd) * (size_t)n);
    const int tstep = tw_n / n;
    for (int k = 0; k < m; ++k)
        for (int r = 0; r < p; ++r) {
            /* X[k + r m] = sum_q W_n^{q (k + r m)} E_q[k] */
            double sr = 0, si = 0;
            for (int q = 0; q < p; ++q) {
                const int e = (int)(((long long)q * (k + r * m)) % n) * tstep;
                const cd w = tw[e];
                const cd v = out[(size_t)q * m + k];
                sr += w.re * v.re - w.im * v.im;
                si += w.re * v.im + w.im * v.re;
            }
            tmp[k + r * m].re = sr;
            tmp[k + r * m].im = si;
        }
    memcpy(out, tmp, sizeof(cd) * (size_t)n);
    free(tmp);
}

static cd *make_tw(int n, int sign) {
    cd *tw = (cd *)malloc(sizeof(cd) * (size_t)n);
    for (int k = 0; k < n; ++k) {
        const double a = sign * 2.0 * PI_D * k / n;
        tw[k].re = cos(a);
        tw[k].im = sin(a);
    }
    return tw;
}

/* in-place strided 1-D transform of `count` lines */
static void fft_lines(cd *data, int n, size_t estride, size_t lstride, size_t count, int sign) {
    cd *tw = make_tw(n, sign);
#pragma omp parallel
    {
        cd *a = (cd *)malloc(sizeof(cd) * (size_t)n), *b = (cd *)malloc(sizeof(cd) * (size_t)n);
#pragma omp for schedule(static)
        for (long long l = 0; l < (long long)count; ++l) {
            cd *base = data + (size_t)l * lstride;
            for (int e = 0; e < n; ++e) a[e] = base[(size_t)e * estride];
            fft_rec(a, b, n, 1, sign, tw, n);
            for (int e = 0; e < n; ++e) base[(size_t)e * estride] = b[e];
        }
        free(a);
        free(b);
    }
    free(tw);
}

/* full complex n-D transforms on double arrays (x fastest) */
static void fft2_full(cd *d, int n, int sign) {
    fft_lines(d, n, 1, (size_t)n, (size_t)n, sign);
    for (int x = 0; x < n; ++x) fft_lines(d + x, n, (size_t)n, 0, 1, sign);
}

void orc_fft2_r2c(const float *img, int n, float *out_c) {
    const int nh = n / 2 + 1;
    cd *d = (cd *)malloc(sizeof(cd) * (size_t)n * n);
    for (int k = 0; k < n * n; ++k) { d[k].re = img[k]; d[k].im = 0; }
    fft2_full(d, n, -1);
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < nh; ++x) {
            out_c[2 * ((size_t)y * nh + x)] = (float)d[(size_t)y * n + x].re;
            out_c[2 * ((size_t)y * nh + x) + 1] = (float)d[(size_t)y * n + x].im;
        }
    free(d);
}

/* the 2-D transforms of the particle preprocessing and of the insertion go through these hooks: cspb_oracle_fast.c points
 * them at its single-precision iterative FFT for the optimised CPU leg; everything else keeps the double-precision ones */
static void (*g_fft2_r2c)(const float *, int, float *) = orc_fft2_r2c;
void orc_fft2_c2r(const float *in_c, int n, float *out);
static void (*g_fft2_c2r)(const float *, int, float *) = orc_fft2_c2r;

void orc_fft2_c2r(const float *in_c, int n, float *out) {
    const int nh = n / 2 + 1;
    cd *d = (cd *)malloc(sizeof(cd) * (size_t)n * n);
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            if (x < nh) {
                d[(size_t)y * n + x].re = in_c[2 * ((size_t)y * nh + x)];
                d[(size_t)y * n + x].im = in_c[2 * ((size_t)y * nh + x) + 1];
            } else {
                const int my = (n - y) % n, mx = n - x;
                d[(size_t)y * n + x].re = in_c[2 * ((size_t)my * nh + mx)];
                d[(size_t)y * n + x].im = -in_c[2 * ((size_t)my * nh + mx) + 1];
            }
        }
    fft2_full(d, n, +1);
    for (int k = 0; k < n * n; ++k) out[k] = (float)d[k].re;
    free(d);
}

/* 3-D transforms on a full complex cube */
static void fft3_full(cd *d, int n, int sign) {
    fft_lines(d, n, 1, (size_t)n, (size_t)n * n, sign);                      /* x */
    for (int z = 0; z < n; ++z)                                              /* y */
        fft_lines(d + (size_t)z * n * n, n, (size_t)n, 1, (size_t)n, sign);
    fft_lines(d, n, (size_t)n * n, 1, (size_t)n * n, sign);                  /* z */
}

/* ================================================================ Euler, CTF */
void orc_euler_matrix(float psi, float theta, float phi, float *r) {
    /* FREALIGN / cisTEM ZYZ: R = Rz(phi) Ry(theta) Rz(psi); pyp decodes the same angles from its
       left-handed matrix at src/pyp/analysis/geometry/core.py:222-247 */
    const float d2r = PI_F / 180.f;
    const float cps = cosf(psi * d2r), sps = sinf(psi * d2r);
    const float cth = cosf(theta * d2r), sth = sinf(theta * d2r);
    const float cph = cosf(phi * d2r), sph = sinf(phi * d2r);
    r[0] = cph * cth * cps - sph * sps;  r[1] = -cph * cth * sps - sph * cps;  r[2] = cph * sth;
    r[3] = sph * cth * cps + cph * sps;  r[4] = -sph * cth * sps + cph * cps;  r[5] = sph * sth;
    r[6] = -sth * cps;                   r[7] = sth * sps;                     r[8] = cth;
}

typedef struct { float a, b, c2, s2, c4, ph0, dstep; } ctfc;

static ctfc ctf_make(const orc_row *row, int box) {
    /* units: src/pyp/inout/metadata/core.py:2891-2923 (defocus A, voltage kV, Cs mm, pixel A) */
    ctfc c;
    const float v = row->voltage_kv * 1000.f;
    const float lambda = 12.2639f / sqrtf(v + 0.97845e-6f * v * v);
    const float l = (float)box * row->pixel_size;
    const float s2u = 1.f / (l * l);
    c.a = PI_F * lambda * 0.5f * (row->defocus_1 + row->defocus_2) * s2u;
    c.b = PI_F * lambda * 0.5f * (row->defocus_1 - row->defocus_2) * s2u;
    c.c2 = cosf(2.f * row->defocus_angle * (PI_F / 180.f));
    c.s2 = sinf(2.f * row->defocus_angle * (PI_F / 180.f));
    c.c4 = -0.5f * PI_F * lambda * lambda * lambda * (row->cs_mm * 1.0e7f) * s2u * s2u;
    c.ph0 = row->phase_shift + atanf(row->amplitude_contrast / sqrtf(1.f - row->amplitude_contrast * row->amplitude_contrast));
    c.dstep = PI_F * lambda * s2u;
    return c;
}

static float ctf_eval(const ctfc *c, int i, int j, float ddef_a) {
    const float fi = (float)i, fj = (float)j, r2 = fi * fi + fj * fj;
    float cos2 = 0.f, sin2 = 0.f;
    if (r2 > 0.f) { cos2 = (fi * fi - fj * fj) / r2; sin2 = 2.f * fi * fj / r2; }
    const float chi = r2 * (c->a + ddef_a * c->dstep + c->b * (cos2 * c->c2 + sin2 * c->s2)) + c->c4 * r2 * r2 + c->ph0;
    return -sinf(chi);
}

void orc_ctf_image(const orc_row *row, int n, float *out) {
    const int nh = n / 2 + 1;
    const ctfc c = ctf_make(row, n);
    for (int jj = 0; jj < n; ++jj)
        for (int i = 0; i < nh; ++i) {
            const int j = jj >= n / 2 ? jj - n : jj;
            out[(size_t)jj * nh + i] = ctf_eval(&c, i, j, 0.f);
        }
}

float orc_band_limits(const orc_refine_cfg *cfg, float *r_lo, float *r_hi) {
    const float npx = (float)cfg->box * cfg->pixel_size;
    float lo = npx / cfg->low_res_limit, hi = npx / cfg->high_res_limit;
    const float cap = (float)(cfg->box / 2 - 2);
    if (hi > cap) hi = cap;
    if (r_lo) *r_lo = lo;
    if (r_hi) *r_hi = hi;
    return hi;
}

int orc_band_count(const orc_refine_cfg *cfg) {
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    const int n = cfg->box;
    int c = 0;
    for (int j = -n / 2; j < n / 2; ++j)
        for (int i = 0; i <= n / 2; ++i) {
            const float r2 = (float)(i * i + j * j);
            if (r2 >= lo * lo && r2 <= hi * hi) ++c;
        }
    return c;
}

/* ================================================================ reference volume */
struct orc_ref {
    int n, pad, np, xh;
    float *v; /* complex [z][y][x], x in [0,np/2], y,z FFT order, centred phase applied */
};

static float sinc2c(int d, int np) {
    if (d == 0) return 1.f;
    const float a = PI_F * (float)d / (float)np;
    const float s = sinf(a) / a;
    return s * s;
}

orc_ref *orc_ref_create(const float *vol, int n, int pad) {
    orc_ref *r = (orc_ref *)calloc(1, sizeof(orc_ref));
    r->n = n; r->pad = pad; r->np = n * pad; r->xh = r->np / 2 + 1;
    const int np = r->np, off = (np - n) / 2;
    cd *d = (cd *)calloc((size_t)np * np * np, sizeof(cd));
    for (int z = 0; z < n; ++z)
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const int X = x + off, Y = y + off, Z = z + off;
                const float g = sinc2c(X - np / 2, np) * sinc2c(Y - np / 2, np) * sinc2c(Z - np / 2, np);
                d[((size_t)Z * np + Y) * np + X].re = vol[((size_t)z * n + y) * n + x] / g;
            }
    fft3_full(d, np, -1);
    r->v = (float *)malloc(sizeof(float) * 2 * (size_t)r->xh * np * np);
    for (int z = 0; z < np; ++z)
        for (int y = 0; y < np; ++y)
            for (int x = 0; x < r->xh; ++x) {
                const cd c = d[((size_t)z * np + y) * np + x];
                const float sg = ((x + y + z) & 1) ? -1.f : 1.f; /* box centre at np/2 */
                const size_t o = 2 * (((size_t)z * np + y) * r->xh + x);
                r->v[o] = sg * (float)c.re;
                r->v[o + 1] = sg * (float)c.im;
            }
    free(d);
    return r;
}

void orc_ref_free(orc_ref *r) { if (r) { free(r->v); free(r); } }

static void ref_at(const orc_ref *r, int x, int y, int z, float *re, float *im) {
    /* integer lattice lookup with Friedel symmetry and FFT-order wrap */
    float sg = 1.f;
    if (x < 0) { x = -x; y = -y; z = -z; sg = -1.f; }
    const int np = r->np;
    if (x > np / 2 || y < -np / 2 || y > np / 2 || z < -np / 2 || z > np / 2) { *re = 0; *im = 0; return; }
    const int iy = ((y % np) + np) % np, iz = ((z % np) + np) % np;
    const size_t o = 2 * (((size_t)iz * np + iy) * r->xh + x);
    *re = r->v[o];
    *im = sg * r->v[o + 1];
}

static void ref_interp(const orc_ref *r, float x, float y, float z, float *re, float *im) {
    /* cisTEM ExtractSlice: trilinear interpolation over the 8 neighbours */
    const int x0 = (int)floorf(x), y0 = (int)floorf(y), z0 = (int)floorf(z);
    const float fx = x - x0, fy = y - y0, fz = z - z0;
    float sr = 0.f, si = 0.f;
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                const float w = (dx ? fx : 1.f - fx) * (dy ? fy : 1.f - fy) * (dz ? fz : 1.f - fz);
                float a, b;
                ref_at(r, x0 + dx, y0 + dy, z0 + dz, &a, &b);
                sr += w * a;
                si += w * b;
            }
    *re = sr;
    *im = si;
}

void orc_project(const orc_ref *r, float psi, float theta, float phi, float r_hi, float *out_c) {
    const int n = r->n, nh = n / 2 + 1;
    float m[9];
    orc_euler_matrix(psi, theta, phi, m);
    for (int jj = 0; jj < n; ++jj)
        for (int i = 0; i < nh; ++i) {
            const int j = jj >= n / 2 ? jj - n : jj;
            float re = 0.f, im = 0.f;
            if ((float)(i * i + j * j) <= r_hi * r_hi)
                ref_interp(r, (m[0] * i + m[1] * j) * r->pad, (m[3] * i + m[4] * j) * r->pad, (m[6] * i + m[7] * j) * r->pad, &re, &im);
            out_c[2 * ((size_t)jj * nh + i)] = re;
            out_c[2 * ((size_t)jj * nh + i) + 1] = im;
        }
}

/* ================================================================ particle preprocessing */
static float cos_edge(float r, float radius, float width) {
    const float lo = radius - 0.5f * width, hi = radius + 0.5f * width;
    if (r <= lo) return 1.f;
    if (r >= hi) return 0.f;
    return 0.5f * (1.f + cosf(PI_F * (r - lo) / width));
}

/* mean / std of the pixels outside `radius` (analysis/image.py:320-338,406-417 convention) */
/* mean / variance of the pixels farther than `radius` (pixels) from the box centre n/2 — the background of
 * src/pyp/analysis/image.py:320-338 (extract_background): a radius beyond the half box is clamped to it */
static void edge_stats(const float *img, int n, float radius, double *mean, double *var) {
    const int c = n / 2;
    if (radius > 0.5f * (float)n) radius = 0.5f * (float)n;
    const int use_all = 0;
    double s = 0, cnt = 0;
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x)
            if (use_all || (float)((x - c) * (x - c) + (y - c) * (y - c)) > radius * radius) { s += img[y * n + x]; cnt += 1; }
    const double m = cnt > 0 ? s / cnt : 0;
    double v = 0;
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x)
            if (use_all || (float)((x - c) * (x - c) + (y - c) * (y - c)) > radius * radius) { const double d = img[y * n + x] - m; v += d * d; }
    *mean = m;
    *var = cnt > 0 ? v / cnt : 0;
}

static void normalized_copy(const float *img, int n, float radius, int normalize, int invert, float *out) {
    float off = 0.f, scl = invert ? -1.f : 1.f;
    if (normalize) {
        double m, v;
        edge_stats(img, n, radius, &m, &v);
        off = (float)m;
        if (v > 0) scl *= (float)(1.0 / sqrt(v));
    }
    for (int k = 0; k < n * n; ++k) out[k] = (img[k] - off) * scl;
}

/* image.py:406-417 normalize_image: (image - background mean) / background std (population std; a zero std
 * leaves the scale alone); `invert` multiplies by -1 (refine3d prompt 47) */
void orc_normalize(const float *img, int n, float radius_px, int normalize, int invert, float *out) {
    normalized_copy(img, n, radius_px, normalize, invert, out);
}

void orc_noise_curve(const float *imgs, int count, const orc_refine_cfg *cfg, float *curve) {
    /* mean |F|^2 per nearest-integer ring over every step-th of the first min(count, 4096) images
       (step = that count / 1024, >= 1) */
    if (count > 4096) count = 4096;
    const int n = cfg->box, nh = n / 2 + 1, nr = n + 1;
    const int step = count > 1024 ? count / 1024 : 1;
    double *sum = (double *)calloc(nr, sizeof(double));
    double *cnt = (double *)calloc(nr, sizeof(double));
    float *tmp = (float *)malloc(sizeof(float) * (size_t)n * n);
    float *spec = (float *)malloc(sizeof(float) * 2 * (size_t)n * nh);
    int ns = 0;
    for (int k = 0; k < count; k += step, ++ns) {
        normalized_copy(imgs + (size_t)k * n * n, n, cfg->mask_radius / cfg->pixel_size, cfg->normalize, cfg->invert_contrast, tmp);
        orc_fft2_r2c(tmp, n, spec);
        for (int jj = 0; jj < n; ++jj)
            for (int i = 0; i < nh; ++i) {
                const int j = jj >= n / 2 ? jj - n : jj;
                const int ring = (int)(sqrtf((float)(i * i + j * j)) + 0.5f);
                const float re = spec[2 * ((size_t)jj * nh + i)], im = spec[2 * ((size_t)jj * nh + i) + 1];
                sum[ring] += (double)(re * re + im * im);
                if (ns == 0) cnt[ring] += 1;
            }
    }
    for (int r = 0; r < nr; ++r) curve[r] = cnt[r] > 0 ? (float)(sum[r] / (cnt[r] * ns)) : 0.f;
    free(sum); free(cnt); free(tmp); free(spec);
}

void orc_prepare_image(const float *img, const orc_refine_cfg *cfg, const float *noise_curve, const float *ring_weights, float *spec) {
    /* normalise -> FFT -> whiten -> (inverse FFT, soft mask, FFT) -> centre -> ring weights */
    const int n = cfg->box, nh = n / 2 + 1;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)n * n);
    normalized_copy(img, n, cfg->mask_radius / cfg->pixel_size, cfg->normalize, cfg->invert_contrast, tmp);
    g_fft2_r2c(tmp, n, spec);
    const int whiten = cfg->whiten && noise_curve;
    if (whiten)
        for (int jj = 0; jj < n; ++jj)
            for (int i = 0; i < nh; ++i) {
                const int j = jj >= n / 2 ? jj - n : jj;
                const int ring = (int)(sqrtf((float)(i * i + j * j)) + 0.5f);
                const float w = noise_curve[ring] > 0.f ? 1.f / sqrtf(noise_curve[ring]) : 0.f;
                spec[2 * ((size_t)jj * nh + i)] *= w;
                spec[2 * ((size_t)jj * nh + i) + 1] *= w;
            }
    if (cfg->apply_mask) {
        g_fft2_c2r(spec, n, tmp);
        const float rad = cfg->mask_radius / cfg->pixel_size, wid = 20.f / cfg->pixel_size;
        const float sc = 1.f / ((float)n * (float)n);
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                const float r = sqrtf((float)((x - n / 2) * (x - n / 2) + (y - n / 2) * (y - n / 2)));
                tmp[y * n + x] *= sc * cos_edge(r, rad, wid);
            }
        g_fft2_r2c(tmp, n, spec);
    }
    for (int jj = 0; jj < n; ++jj)
        for (int i = 0; i < nh; ++i) {
            const int j = jj >= n / 2 ? jj - n : jj;
            float w = ((i + j) & 1) ? -1.f : 1.f;
            if (ring_weights) w *= ring_weights[(int)sqrtf((float)(i * i + j * j))];
            spec[2 * ((size_t)jj * nh + i)] *= w;
            spec[2 * ((size_t)jj * nh + i) + 1] *= w;
        }
    free(tmp);
}

/* ================================================================ score */
/* `ring_cut`: only the rings floor(r) <= ring_cut enter the sums (coarse-to-fine stages of §7c); INT_MAX = the whole band */
static float score_cut(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg, float *out4, int ring_cut);
float orc_score(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg, float *out4) {
    return score_cut(r, spec, row, pose6, cfg, out4, 0x7fffffff);
}

static float score_cut(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg, float *out4, int ring_cut) {
    const int n = cfg->box, nh = n / 2 + 1;
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    const int nb = 2 * n;
    float *cross = (float *)calloc(nb, sizeof(float));
    float suma = 0.f, sumb = 0.f;
    float m[9];
    orc_euler_matrix(pose6[0], pose6[1], pose6[2], m);
    const ctfc c = ctf_make(row, n);
    const float k2 = 2.f * PI_F / ((float)n * row->pixel_size);
    for (int j = -n / 2; j < n / 2; ++j)
        for (int i = 0; i <= n / 2; ++i) {
            const float r2 = (float)(i * i + j * j);
            if (r2 < lo * lo || r2 > hi * hi) continue;
            const int bin = (int)sqrtf(r2);
            if (bin > ring_cut) continue;
            const int jj = j < 0 ? j + n : j;
            const float fr = spec[2 * ((size_t)jj * nh + i)], fim = spec[2 * ((size_t)jj * nh + i) + 1];
            float pr, pi;
            ref_interp(r, (m[0] * i + m[1] * j) * r->pad, (m[3] * i + m[4] * j) * r->pad, (m[6] * i + m[7] * j) * r->pad, &pr, &pi);
            const float ctf = ctf_eval(&c, i, j, pose6[5]);
            pr *= ctf;
            pi *= ctf;
            const float ph = (i * pose6[3] + j * pose6[4]) * k2;
            const float cs = cosf(ph), sn = sinf(ph);
            const float gr = fr * cs - fim * sn, gi = fr * sn + fim * cs;
            cross[bin] += gr * pr + gi * pi;
            suma += fr * fr + fim * fim;
            sumb += pr * pr + pi * pi;
        }
    const int limit = cfg->signed_cc_limit > 0.f ? (int)floorf((float)n * cfg->pixel_size / cfg->signed_cc_limit) : 0x7fffffff;
    float num = 0.f, xs = 0.f;
    for (int b = 0; b < nb; ++b) {
        xs += cross[b];
        num += (b > limit) ? fabsf(cross[b]) : cross[b];
    }
    free(cross);
    if (out4) { out4[0] = num; out4[1] = xs; out4[2] = suma; out4[3] = sumb; }
    const float den = suma * sumb;
    return den > 0.f ? 100.f * num / sqrtf(den) : 0.f;
}

/* ================================================================ beam-tilt phase sum (refine_ctf) */
void orc_phase_sum(const orc_ref *r, const float *specs, const orc_row *rows, int n_img, const orc_refine_cfg *cfg, float *out_c) {
    const int n = cfg->box, nh = n / 2 + 1;
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    double *acc = (double *)calloc((size_t)2 * n * nh, sizeof(double));
    for (int k = 0; k < n_img; ++k) {
        const float *spec = specs + 2 * (size_t)k * n * nh;
        const orc_row *row = &rows[k];
        float m[9];
        orc_euler_matrix(row->psi, row->theta, row->phi, m);
        const ctfc c = ctf_make(row, n);
        const float k2 = 2.f * PI_F / ((float)n * row->pixel_size);
        for (int j = -n / 2; j < n / 2; ++j)
            for (int i = 0; i <= n / 2; ++i) {
                const float r2 = (float)(i * i + j * j);
                if (r2 < lo * lo || r2 > hi * hi) continue;
                const int jj = j < 0 ? j + n : j;
                const float fr = spec[2 * ((size_t)jj * nh + i)], fim = spec[2 * ((size_t)jj * nh + i) + 1];
                float pr, pi;
                ref_interp(r, (m[0] * i + m[1] * j) * r->pad, (m[3] * i + m[4] * j) * r->pad, (m[6] * i + m[7] * j) * r->pad, &pr, &pi);
                const float ctf = ctf_eval(&c, i, j, 0.f);
                const float ph = (i * row->x_shift + j * row->y_shift) * k2;
                const float cs = cosf(ph), sn = sinf(ph);
                const float gr = fr * cs - fim * sn, gi = fr * sn + fim * cs;
                acc[2 * ((size_t)jj * nh + i)] += (double)(ctf * (gr * pr + gi * pi));
                acc[2 * ((size_t)jj * nh + i) + 1] += (double)(ctf * (gi * pr - gr * pi));
            }
    }
    for (size_t t = 0; t < (size_t)2 * n * nh; ++t) out_c[t] = (float)acc[t];
    free(acc);
}

/* ================================================================ local refinement */
#define NP 6
#define NL 3
static float wrap360(float a) { a = fmodf(a, 360.f); if (a < 0.f) a += 360.f; return a; }

static void stats_from(const float *v, int ns, float *sigma, float *logp) {
    const float X = v[1], A = v[2], B = v[3];
    const float alpha = B > 0.f ? X / B : 0.f;
    float resid = A - 2.f * alpha * X + alpha * alpha * B;
    if (resid < 0.f) resid = 0.f;
    const float sig2 = alpha * alpha * B;
    *sigma = sig2 > 0.f ? sqrtf(resid / sig2) : 0.f;
    const float n = (float)(ns > 0 ? ns : 1);
    const float var = resid / n;
    *logp = var > 0.f ? -0.5f * n * (1.f + logf(2.f * PI_F * var)) : 0.f;
}


/* diagonal Newton step from a central-difference stencil, continuous in (f0, fp, fm):
   d = g / max(-c, |g| / dmax) with dmax = 4h (trust region) */
static float newton_step(float f0, float fp, float fm, float h) {
    const float g = (fp - fm) / (2.f * h);
    const float c = (fp - 2.f * f0 + fm) / (h * h);
    const float dmax = 4.f * h;
    float den = -c;
    const float floor_ = fabsf(g) / dmax;
    if (den < floor_) den = floor_;
    return den > 0.f ? g / den : 0.f;
}

/* step length from f(0) and f at t = 0.5, 1, 2: least-squares parabola, maximiser clamped to
   [0, 2.5]; a convex fit takes the better end of the interval */
static const float LS_T[3] = {0.5f, 1.f, 2.f};
static float line_step(float f0, const float *fl) {
    /* fit f(t) - f0 = a t^2 + b t on the three samples (normal equations, 2x2) */
    float s22 = 0.f, s23 = 0.f, s33 = 0.f, r2 = 0.f, r3 = 0.f;
    for (int l = 0; l < 3; ++l) {
        const float t = LS_T[l], y = fl[l] - f0;
        s22 += t * t; s23 += t * t * t; s33 += t * t * t * t;
        r2 += t * y; r3 += t * t * y;
    }
    const float det = s22 * s33 - s23 * s23;
    const float b = (r2 * s33 - r3 * s23) / det;
    const float a = (r3 * s22 - r2 * s23) / det;
    if (a < 0.f) {
        float t = -b / (2.f * a);
        if (t < 0.f) t = 0.f;
        if (t > 2.5f) t = 2.5f;
        return t;
    }
    return (a * 2.5f + b > 0.f) ? 2.5f : 0.f; /* convex fit: best end of [0, 2.5] under the model */
}

/* refine one starting pose x[6] in place; returns the final score (x100) and its band sums in o4.
   The better of {refined, start} is kept.  *evals is incremented per objective evaluation. */
/* Shift restraint of the local refinement (refine3d prompt 7; SEMANTICS.md §7b), in CC units:
 * (sigma^2 / N_mask) * sum_{k in x,y} (p_k - mean_k)^2 / (2 var_k), sigma = the row's SIGMA column,
 * N_mask = pixels inside the mask radius.  Angles are not restrained (cisTEM restrains the shifts only). */
static float prior_pen(const orc_refine_cfg *cfg, const orc_row *row, const float *x) {
    if (!cfg->use_priors) return 0.f;
    const float rad = cfg->mask_radius / cfg->pixel_size;
    float nmask = 3.14159265f * rad * rad;
    if (nmask < 1.f) nmask = 1.f;
    const float lam = row->sigma * row->sigma / nmask;
    const float wx = cfg->prior_var_x > 0.f ? 0.5f / cfg->prior_var_x : 0.f;
    const float wy = cfg->prior_var_y > 0.f ? 0.5f / cfg->prior_var_y : 0.f;
    const float dx = x[3] - cfg->prior_mean_x, dy = x[4] - cfg->prior_mean_y;
    return lam * (wx * dx * dx + wy * dy * dy);
}

/* returns the score (100 CC) of the pose kept; *obj_out = CC - restraint of that pose */
static float refine_one(const orc_ref *r, const float *spec, const orc_row *row, float *x, const int *freem,
                        const orc_refine_cfg *cfg, float *o4, long long *evals, float coarse, float *obj_out) {
    const int n = cfg->box;
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    int n_free = 0;
    for (int m = 0; m < NP; ++m) n_free += freem[m] ? 1 : 0;
    const float h_ang = 0.35f * 57.29578f / hi;                    /* ~1/3 of the angular resolution at r_hi */
    const float h_shift = 0.07f * (float)n / hi * cfg->pixel_size; /* Angstrom */
    const float h_def = cfg->defocus_step > 0.f ? cfg->defocus_step : 50.f;
    const int iters = n_free > 0 ? (cfg->local_iterations > 0 ? cfg->local_iterations : 8) : 0;
    const int late = iters / 2 + 1; /* stencil steps stay constant for the first half, then shrink */
    /* coarse >= 1: the pose steps start `coarse` times larger (hits of the global search sit up to
       half a grid step from the optimum; the stencil then spans the search resolution) */
    float h[NP] = {h_ang * coarse, h_ang * coarse, h_ang * coarse, h_shift * coarse, h_shift * coarse, h_def};
    float d[NP], q[NP];
    const float x_start[NP] = {x[0], x[1], x[2], x[3], x[4], x[5]};
    for (int it = 0; it < iters; ++it) {
        const float f0 = orc_score(r, spec, row, x, cfg, o4) * 0.01f - prior_pen(cfg, row, x);
        (*evals)++;
        for (int m = 0; m < NP; ++m) {
            d[m] = 0.f;
            if (!freem[m]) continue;
            memcpy(q, x, sizeof q);
            q[m] = x[m] + h[m];
            const float fp = orc_score(r, spec, row, q, cfg, o4) * 0.01f - prior_pen(cfg, row, q);
            q[m] = x[m] - h[m];
            const float fm = orc_score(r, spec, row, q, cfg, o4) * 0.01f - prior_pen(cfg, row, q);
            (*evals) += 2;
            d[m] = newton_step(f0, fp, fm, h[m]);
        }
        float fl[NL];
        for (int l = 0; l < NL; ++l) {
            for (int m = 0; m < NP; ++m) q[m] = x[m] + LS_T[l] * d[m];
            fl[l] = orc_score(r, spec, row, q, cfg, o4) * 0.01f - prior_pen(cfg, row, q);
            (*evals)++;
        }
        const float t = line_step(f0, fl);
        for (int m = 0; m < NP; ++m) x[m] += t * d[m];
        if (it + 1 >= late)
            for (int m = 0; m < NP; ++m) h[m] *= 0.6f;
    }
    /* final: score the refined and the starting pose, never return a worse one */
    float o4s[4];
    float sc = orc_score(r, spec, row, x, cfg, o4);
    const float sc_start = orc_score(r, spec, row, x_start, cfg, o4s);
    (*evals) += 2;
    float obj = sc * 0.01f - prior_pen(cfg, row, x);
    const float obj_start = sc_start * 0.01f - prior_pen(cfg, row, x_start);
    if (obj < obj_start) { memcpy(x, x_start, sizeof x_start); memcpy(o4, o4s, sizeof o4s); sc = sc_start; obj = obj_start; }
    if (obj_out) *obj_out = obj;
    return sc;
}

/* ================================================================ analytic-gradient local refinement (SEMANTICS.md §7c)
 * One evaluation gives the score AND its derivatives: the eight corners of the trilinear interpolant yield the
 * value and the spatial gradient of the slice sample, the chain rule through R(psi, theta, phi) gives dP/dangle,
 * the phase ramp gives d/dshift.  Besides the gradient the evaluation accumulates J^T J (J = dP/dparameters): for data
 * G = alpha P(p*) + noise the curvature of CC around p* is -(CC / B) J^T J (Gauss-Newton: second derivatives of the
 * interpolant dropped), which is the Hessian the step uses. */
#define NG 5
#define LM_SOFT 0.02f
#define NJ 15 /* upper triangle of the 5x5 J^T J, row major: (0,0) (0,1) ... (0,4) (1,1) ... (4,4) */

static void euler_derivatives(float psi, float theta, float phi, float *dth6) {
    /* first two columns of dR/dtheta (rows x, y, z), per radian; dR/dpsi and dR/dphi need no table:
       d(x,y,z)/dpsi = R (-j, i, 0)^T and d(x,y,z)/dphi = (-y, x, 0) */
    const float d2r = PI_F / 180.f;
    const float cps = cosf(psi * d2r), sps = sinf(psi * d2r);
    const float cth = cosf(theta * d2r), sth = sinf(theta * d2r);
    const float cph = cosf(phi * d2r), sph = sinf(phi * d2r);
    dth6[0] = -cph * sth * cps; dth6[1] = cph * sth * sps;
    dth6[2] = -sph * sth * cps; dth6[3] = sph * sth * sps;
    dth6[4] = -cth * cps;       dth6[5] = cth * sps;
}

/* trilinear value and spatial gradient at (x >= 0 after the Friedel flip done by the caller) */
static void ref_interp_grad(const orc_ref *r, float x, float y, float z, float *v /*2*/, float *gx, float *gy, float *gz) {
    const int x0 = (int)floorf(x), y0 = (int)floorf(y), z0 = (int)floorf(z);
    const float fx = x - x0, fy = y - y0, fz = z - z0;
    float c[2][2][2][2];
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) ref_at(r, x0 + dx, y0 + dy, z0 + dz, &c[dz][dy][dx][0], &c[dz][dy][dx][1]);
    for (int k = 0; k < 2; ++k) {
        /* along x first (the GPU's order): a = c0 + fx (c1 - c0) */
        float a[2][2], d[2][2];
        for (int dz = 0; dz < 2; ++dz)
            for (int dy = 0; dy < 2; ++dy) {
                d[dz][dy] = c[dz][dy][1][k] - c[dz][dy][0][k];
                a[dz][dy] = c[dz][dy][0][k] + fx * d[dz][dy];
            }
        const float dx0 = d[0][0] + fy * (d[0][1] - d[0][0]), dx1 = d[1][0] + fy * (d[1][1] - d[1][0]);
        gx[k] = dx0 + fz * (dx1 - dx0);
        const float dy0 = a[0][1] - a[0][0], dy1 = a[1][1] - a[1][0];
        const float v0 = a[0][0] + fy * dy0, v1 = a[1][0] + fy * dy1;
        gy[k] = dy0 + fz * (dy1 - dy0);
        gz[k] = v1 - v0;
        v[k] = v0 + fz * gz[k];
    }
}

/* score (x100) with derivatives.  out4 = {num, X, A, B}; dnum[5] = d num / d (psi, theta, phi [deg], x, y [A]) with the
 * signed / absolute ring rule applied; dB[3] = d B / d angles [deg]; jtj[15] = J^T J in the same units. */
static float score_grad_cut(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg,
                            float *out4, float *dnum, float *dB, float *jtj, int ring_cut);
float orc_score_grad_cut(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg,
                         float *out4, float *dnum, float *dB, float *jtj, int ring_cut) {
    return score_grad_cut(r, spec, row, pose6, cfg, out4, dnum, dB, jtj, ring_cut > 0 ? ring_cut : 0x7fffffff);
}
float orc_score_grad(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg,
                     float *out4, float *dnum, float *dB, float *jtj) {
    return score_grad_cut(r, spec, row, pose6, cfg, out4, dnum, dB, jtj, 0x7fffffff);
}

static float score_grad_cut(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg,
                            float *out4, float *dnum, float *dB, float *jtj, int ring_cut) {
    const int n = cfg->box, nh = n / 2 + 1;
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    const int nb = 2 * n;
    float *ring = (float *)calloc((size_t)nb * (3 + NG), sizeof(float)); /* per ring: X, its five derivatives, A_r, B_r */
    float suma = 0.f, sumb = 0.f;
    float m[9], dth[6];
    orc_euler_matrix(pose6[0], pose6[1], pose6[2], m);
    euler_derivatives(pose6[0], pose6[1], pose6[2], dth);
    const ctfc c = ctf_make(row, n);
    const float k2 = 2.f * PI_F / ((float)n * row->pixel_size);
    const float d2r = PI_F / 180.f, pad = (float)r->pad;
    for (int a = 0; a < 3; ++a) dB[a] = 0.f;
    for (int a = 0; a < NJ; ++a) jtj[a] = 0.f;
    for (int j = -n / 2; j < n / 2; ++j)
        for (int i = 0; i <= n / 2; ++i) {
            const float r2 = (float)(i * i + j * j);
            if (r2 < lo * lo || r2 > hi * hi) continue;
            const int bin = (int)sqrtf(r2);
            if (bin > ring_cut) continue;
            const int jj = j < 0 ? j + n : j;
            const float fr = spec[2 * ((size_t)jj * nh + i)], fim = spec[2 * ((size_t)jj * nh + i) + 1];
            const float fi = (float)i, fj = (float)j;
            float x = (m[0] * fi + m[1] * fj) * pad, y = (m[3] * fi + m[4] * fj) * pad, z = (m[6] * fi + m[7] * fj) * pad;
            /* coordinate velocities per radian of psi, theta, phi */
            float vel[3][3] = {{(m[1] * fi - m[0] * fj) * pad, (m[4] * fi - m[3] * fj) * pad, (m[7] * fi - m[6] * fj) * pad},
                               {(dth[0] * fi + dth[1] * fj) * pad, (dth[2] * fi + dth[3] * fj) * pad, (dth[4] * fi + dth[5] * fj) * pad},
                               {-y, x, 0.f}};
            const float sg = x < 0.f ? -1.f : 1.f; /* Friedel mate for the negative half space */
            x *= sg; y *= sg; z *= sg;
            float v[2], gx[2], gy[2], gz[2];
            ref_interp_grad(r, x, y, z, v, gx, gy, gz);
            const float ctf = ctf_eval(&c, i, j, 0.f);
            const float pr = ctf * v[0], pi = ctf * sg * v[1];
            float dp[NG][2]; /* dP / d parameter; the shift columns are those of the equivalent projection shift */
            for (int a = 0; a < 3; ++a) {
                const float dre = gx[0] * vel[a][0] + gy[0] * vel[a][1] + gz[0] * vel[a][2];
                const float dim = gx[1] * vel[a][0] + gy[1] * vel[a][1] + gz[1] * vel[a][2];
                dp[a][0] = ctf * sg * dre * d2r; /* coordinates move by sg * vel */
                dp[a][1] = ctf * dim * d2r;
            }
            dp[3][0] = k2 * fi * pi; dp[3][1] = -k2 * fi * pr; /* -i k P: shifting the image by +s = shifting P by -s */
            dp[4][0] = k2 * fj * pi; dp[4][1] = -k2 * fj * pr;
            const float ph = (fi * pose6[3] + fj * pose6[4]) * k2;
            const float cs = cosf(ph), sn = sinf(ph);
            const float gr = fr * cs - fim * sn, gi = fr * sn + fim * cs;
            float *rg = ring + (size_t)bin * (3 + NG);
            rg[0] += gr * pr + gi * pi;
            for (int a = 0; a < NG; ++a) rg[1 + a] += gr * dp[a][0] + gi * dp[a][1];
            rg[1 + NG] += fr * fr + fim * fim;
            rg[2 + NG] += pr * pr + pi * pi;
            suma += fr * fr + fim * fim;
            sumb += pr * pr + pi * pi;
            for (int a = 0; a < 3; ++a) dB[a] += 2.f * (pr * dp[a][0] + pi * dp[a][1]);
            int t = 0;
            for (int a = 0; a < NG; ++a)
                for (int b = a; b < NG; ++b) jtj[t++] += dp[a][0] * dp[b][0] + dp[a][1] * dp[b][1];
        }
    const int limit = cfg->signed_cc_limit > 0.f ? (int)floorf((float)n * cfg->pixel_size / cfg->signed_cc_limit) : 0x7fffffff;
    float num = 0.f, xs = 0.f;
    for (int a = 0; a < NG; ++a) dnum[a] = 0.f;
    for (int b = 0; b < nb; ++b) {
        const float *rg = ring + (size_t)b * (3 + NG);
        xs += rg[0];
        num += (b > limit) ? fabsf(rg[0]) : rg[0];
        /* derivative of |X_r| with a soft sign X_r / sqrt(X_r^2 + eps_r^2), eps_r = 0.02 sqrt(A_r B_r): a ring whose
           correlation is below 0.02 (far inside its noise) contributes in proportion instead of flipping with the
           rounding noise of X_r — the gradient stays continuous where the objective has its kinks */
        float sg = 1.f;
        if (b > limit) {
            const float q = rg[0] * rg[0] + LM_SOFT * LM_SOFT * rg[1 + NG] * rg[2 + NG];
            sg = q > 0.f ? rg[0] / sqrtf(q) : 0.f;
        }
        for (int a = 0; a < NG; ++a) dnum[a] += sg * rg[1 + a];
    }
    free(ring);
    if (out4) { out4[0] = num; out4[1] = xs; out4[2] = suma; out4[3] = sumb; }
    const float den = suma * sumb;
    return den > 0.f ? 100.f * num / sqrtf(den) : 0.f;
}

/* solve the symmetric positive definite 5x5 system H d = g (Cholesky, fp32); returns 0 if H is not positive definite */
static int solve_spd5(const float *H /*25*/, const float *g, float *d) {
    float L[NG][NG];
    memset(L, 0, sizeof L);
    for (int i = 0; i < NG; ++i)
        for (int j = 0; j <= i; ++j) {
            float s = H[i * NG + j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            if (i == j) {
                if (!(s > 0.f)) return 0;
                L[i][i] = sqrtf(s);
            } else
                L[i][j] = s / L[j][j];
        }
    float yv[NG];
    for (int i = 0; i < NG; ++i) {
        float s = g[i];
        for (int k = 0; k < i; ++k) s -= L[i][k] * yv[k];
        yv[i] = s / L[i][i];
    }
    for (int i = NG - 1; i >= 0; --i) {
        float s = yv[i];
        for (int k = i + 1; k < NG; ++k) s -= L[k][i] * d[k];
        d[i] = s / L[i][i];
    }
    return 1;
}

/* Levenberg-Marquardt-style step from one gradient evaluation (continuous in its inputs: no accept / reject).
 * o4, dnum, dB, jtj as returned by orc_score_grad; lam / prior as in prior_pen.  Returns f0 (objective at x); writes the
 * step d[5] (trust region applied) and *slope = directional derivative of the objective along d. */
#define LM_DAMP 0.05f
#define LM_CC_FLOOR 0.01f
static float lm_step(const float *o4, const float *dnum, const float *dB, const float *jtj, const int *freem, const float *trust,
                     const orc_refine_cfg *cfg, const orc_row *row, const float *x, float *d, float *slope) {
    const float A = o4[2], B = o4[3];
    const float den = A * B;
    const float rs = den > 0.f ? 1.f / sqrtf(den) : 0.f;
    const float cc = o4[0] * rs;
    float g[NG], H[NG * NG];
    for (int a = 0; a < NG; ++a) g[a] = dnum[a] * rs - (a < 3 && B > 0.f ? 0.5f * cc * dB[a] / B : 0.f);
    const float cce = cc > LM_CC_FLOOR ? cc : LM_CC_FLOOR;
    const float hs = B > 0.f ? cce / B : 0.f;
    int t = 0;
    for (int a = 0; a < NG; ++a)
        for (int b = a; b < NG; ++b) { H[a * NG + b] = H[b * NG + a] = hs * jtj[t]; ++t; }
    float f0 = cc;
    if (cfg->use_priors) { /* restraint: value, gradient and (exact) curvature */
        const float rad = cfg->mask_radius / cfg->pixel_size;
        float nmask = 3.14159265f * rad * rad;
        if (nmask < 1.f) nmask = 1.f;
        const float lam = row->sigma * row->sigma / nmask;
        const float w[2] = {cfg->prior_var_x > 0.f ? 0.5f / cfg->prior_var_x : 0.f, cfg->prior_var_y > 0.f ? 0.5f / cfg->prior_var_y : 0.f};
        const float dx[2] = {x[3] - cfg->prior_mean_x, x[4] - cfg->prior_mean_y};
        for (int k = 0; k < 2; ++k) {
            f0 -= lam * w[k] * dx[k] * dx[k];
            g[3 + k] -= 2.f * lam * w[k] * dx[k];
            H[(3 + k) * NG + 3 + k] += 2.f * lam * w[k];
        }
    }
    for (int a = 0; a < NG; ++a) {
        if (!freem[a]) {
            for (int b = 0; b < NG; ++b) H[a * NG + b] = H[b * NG + a] = 0.f;
            H[a * NG + a] = 1.f;
            g[a] = 0.f;
        }
    }
    for (int a = 0; a < NG; ++a) H[a * NG + a] *= 1.f + LM_DAMP;
    if (!solve_spd5(H, g, d))
        for (int a = 0; a < NG; ++a) d[a] = H[a * NG + a] > 0.f ? g[a] / H[a * NG + a] : 0.f;
    float worst = 1.f; /* trust region: scale the whole step so that no component exceeds its radius */
    for (int a = 0; a < NG; ++a) {
        const float q = fabsf(d[a]) / trust[a];
        if (q > worst) worst = q;
    }
    *slope = 0.f;
    for (int a = 0; a < NG; ++a) {
        d[a] /= worst;
        *slope += g[a] * d[a];
    }
    return f0;
}

/* step length from f(0), f'(0) along d and f(1): parabola through the three, maximiser clamped to [0, 2] — blended
 * towards the plain Gauss-Newton step t = 1 as the predicted gain `slope` falls to the rounding noise of the scores
 * (weight slope^2 / (slope^2 + tau^2), tau = 2e-5 in CC units): a converged state takes t = 1 instead of a fit to noise */
#define LM_TAU 2e-5f
static float lm_line(float f0, float slope, float f1) {
    const float c = f1 - f0 - slope;
    float t;
    if (c < 0.f) {
        t = -slope / (2.f * c);
        if (t < 0.f) t = 0.f;
        if (t > 2.f) t = 2.f;
    } else
        t = f1 > f0 ? 2.f : 0.f;
    const float w = slope * slope / (slope * slope + LM_TAU * LM_TAU);
    return 1.f + w * (t - 1.f);
}

/* resolution stage of iteration `it`: returns f >= 1 and the last ring used (INT_MAX for the full band) */
static float lm_stage(int it, int iters, float r_lo, float r_hi, int *ring_cut) {
    const int ramp = iters - 3;
    float f = 1.f;
    if (ramp > 0 && it < ramp) f = powf(6.f, 1.f - (float)it / (float)ramp);
    *ring_cut = 0x7fffffff;
    if (f > 1.f) {
        int rc = (int)floorf(r_hi / f);
        int least = (int)floorf(r_lo) + 4; /* never fewer than a handful of rings, nor below 12 Fourier pixels */
        if (least < 12) least = 12;
        if (rc < least) rc = least;
        if ((float)rc >= r_hi) { f = 1.f; } else { *ring_cut = rc; f = r_hi / (float)rc; }
    }
    return f;
}

/* refine_one with the analytic optimiser: per iteration ONE gradient evaluation and ONE trial evaluation */
static float refine_one_lm(const orc_ref *r, const float *spec, const orc_row *row, float *x, const int *freem, const orc_refine_cfg *cfg,
                           float *o4, long long *evals, float *obj_out) {
    const int n = cfg->box;
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    int n_free = 0;
    for (int m = 0; m < NG; ++m) n_free += freem[m] ? 1 : 0;
    const int iters = n_free > 0 ? (cfg->local_iterations > 0 ? cfg->local_iterations : 8) : 0;
    const float x_start[NP] = {x[0], x[1], x[2], x[3], x[4], x[5]};
    for (int it = 0; it < iters; ++it) {
        /* coarse to fine: stage `it` scores the rings up to r_hi / f, f falling geometrically from 6 to 1 over the first
           iters - 3 stages (frequency marching: a wide basin first, the full band for the last three); the trust region
           follows the resolution of the stage */
        int ring_cut;
        const float f = lm_stage(it, iters, lo, hi, &ring_cut);
        const float h_ang = 0.35f * 57.29578f * f / hi, h_shift = 0.07f * (float)n * f / hi * cfg->pixel_size;
        const float trust[NG] = {16.f * h_ang, 16.f * h_ang, 16.f * h_ang, 16.f * h_shift, 16.f * h_shift};
        float dnum[NG], dB[3], jtj[NJ], d[NG], slope, q[NP];
        score_grad_cut(r, spec, row, x, cfg, o4, dnum, dB, jtj, ring_cut);
        const float f0 = lm_step(o4, dnum, dB, jtj, freem, trust, cfg, row, x, d, &slope);
        memcpy(q, x, sizeof q);
        for (int m = 0; m < NG; ++m) q[m] = x[m] + d[m];
        const float f1 = score_cut(r, spec, row, q, cfg, o4, ring_cut) * 0.01f - prior_pen(cfg, row, q);
        (*evals) += 2;
        const float t = lm_line(f0, slope, f1);
        if (getenv("ORC_LM_DEBUG"))
            fprintf(stderr, "lm it %d cut %d f0 %.7f f1 %.7f slope %.3e c %.3e t %.4f d %.5f %.5f %.5f %.5f %.5f x %.5f %.5f %.5f %.5f %.5f\n", it, ring_cut, f0, f1, slope,
                    f1 - f0 - slope, t, d[0], d[1], d[2], d[3], d[4], x[0], x[1], x[2], x[3], x[4]);
        for (int m = 0; m < NG; ++m) x[m] += t * d[m];
    }
    float o4s[4];
    float sc = orc_score(r, spec, row, x, cfg, o4);
    const float sc_start = orc_score(r, spec, row, x_start, cfg, o4s);
    (*evals) += 2;
    float obj = sc * 0.01f - prior_pen(cfg, row, x);
    const float obj_start = sc_start * 0.01f - prior_pen(cfg, row, x_start);
    if (obj < obj_start) { memcpy(x, x_start, sizeof x_start); memcpy(o4, o4s, sizeof o4s); sc = sc_start; obj = obj_start; }
    if (obj_out) *obj_out = obj;
    return sc;
}

/* ---- 2-D focus mask (refine3d prompts 29-32, 44; SEMANTICS.md §6b).  The sphere centre c (Angstrom from
 * the corner of the map) is projected with the particle's pose: image axes are the first two columns of
 * the rotation matrix (the slice geometry of orc_score), so the centre lands at n/2 + (col0.c', col1.c'),
 * c' = c / pixel - n/2.  LOGP = -N/2 (1 + ln(2 pi var)), var = mean of the squared real-space residual
 * (shifted image - alpha CTF projection, band-limited like the score) over the N pixels of the disc. */
void orc_focus_center(const orc_refine_cfg *cfg, const float *pose6, float *cx, float *cy) {
    float m[9];
    orc_euler_matrix(pose6[0], pose6[1], pose6[2], m);
    const float h = (float)(cfg->box / 2);
    const float x = cfg->focus_x / cfg->pixel_size - h, y = cfg->focus_y / cfg->pixel_size - h, z = cfg->focus_z / cfg->pixel_size - h;
    *cx = h + m[0] * x + m[3] * y + m[6] * z;
    *cy = h + m[1] * x + m[4] * y + m[7] * z;
}

float orc_focus_logp(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg,
                     const float *o4) {
    const int n = cfg->box, nh = n / 2 + 1;
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    float m[9];
    orc_euler_matrix(pose6[0], pose6[1], pose6[2], m);
    const ctfc c = ctf_make(row, n);
    const float k2 = 2.f * PI_F / ((float)n * row->pixel_size);
    const float alpha = o4[3] > 0.f ? o4[1] / o4[3] : 0.f;
    float *D = (float *)calloc((size_t)2 * n * nh, sizeof(float));
    float *real = (float *)malloc(sizeof(float) * (size_t)n * n);
    for (int j = -n / 2; j < n / 2; ++j)
        for (int i = 0; i <= n / 2; ++i) {
            const float r2 = (float)(i * i + j * j);
            if (r2 < lo * lo || r2 > hi * hi) continue;
            const int jj = j < 0 ? j + n : j;
            const float fr = spec[2 * ((size_t)jj * nh + i)], fim = spec[2 * ((size_t)jj * nh + i) + 1];
            float pr, pi;
            ref_interp(r, (m[0] * i + m[1] * j) * r->pad, (m[3] * i + m[4] * j) * r->pad, (m[6] * i + m[7] * j) * r->pad, &pr, &pi);
            const float ctf = ctf_eval(&c, i, j, pose6[5]);
            const float ph = (i * pose6[3] + j * pose6[4]) * k2;
            const float cs = cosf(ph), sn = sinf(ph);
            const float gr = fr * cs - fim * sn, gi = fr * sn + fim * cs;
            const float w = ((i + j) & 1) ? -1.f : 1.f; /* back from the centred phase origin to the image corner */
            D[2 * ((size_t)jj * nh + i)] = w * (gr - alpha * ctf * pr);
            D[2 * ((size_t)jj * nh + i) + 1] = w * (gi - alpha * ctf * pi);
        }
    orc_fft2_c2r(D, n, real);
    float cx, cy;
    orc_focus_center(cfg, pose6, &cx, &cy);
    const float rad = cfg->focus_radius / cfg->pixel_size, inv = 1.f / ((float)n * (float)n);
    double ss = 0.0;
    long cnt = 0;
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            const float dx = (float)x - cx, dy = (float)y - cy;
            if (dx * dx + dy * dy > rad * rad) continue;
            const float v = real[(size_t)y * n + x] * inv;
            ss += (double)v * v;
            ++cnt;
        }
    free(D);
    free(real);
    if (cnt == 0) return 0.f;
    const float var = (float)(ss / (double)cnt);
    return var > 0.f ? -0.5f * (float)cnt * (1.f + logf(2.f * PI_F * var)) : 0.f;
}

static void write_row(orc_row *row, const float *x, float sc, const float *o4, int nband, int refine_defocus) {
    row->psi = wrap360(x[0]);
    row->theta = x[1];
    row->phi = wrap360(x[2]);
    row->x_shift = x[3];
    row->y_shift = x[4];
    if (refine_defocus) { row->defocus_1 += x[5]; row->defocus_2 += x[5]; }
    row->score = sc;
    stats_from(o4, nband, &row->sigma, &row->logp);
}

long long orc_refine_local(const orc_ref *r, const float *specs, orc_row *rows, int n_img, const orc_refine_cfg *cfg) {
    const int n = cfg->box, nh = n / 2 + 1;
    const int freem[NP] = {cfg->refine_psi, cfg->refine_theta, cfg->refine_phi, cfg->refine_x, cfg->refine_y, cfg->refine_defocus};
    const int nband = orc_band_count(cfg);
    long long evals = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : evals)
    for (int k = 0; k < n_img; ++k) {
        const float *spec = specs + 2 * (size_t)k * n * nh;
        orc_row *row = &rows[k];
        float x[NP] = {row->psi, row->theta, row->phi, row->x_shift, row->y_shift, 0.f};
        float o4[4];
        long long ev = 0;
        /* optimiser 0 (default): analytic gradient + Gauss-Newton step (SEMANTICS.md §7c) for the five pose parameters;
           the stencil optimiser of §7 when asked for (optimizer = 1) or when the defocus is refined as well */
        const int analytic = cfg->optimizer == 0 && !cfg->refine_defocus;
        const float sc = analytic ? refine_one_lm(r, spec, row, x, freem, cfg, o4, &ev, NULL)
                                  : refine_one(r, spec, row, x, freem, cfg, o4, &ev, 1.f, NULL);
        evals += ev;
        const orc_row before = *row;
        write_row(row, x, sc, o4, nband, cfg->refine_defocus);
        if (cfg->focus_radius > 0.f) row->logp = orc_focus_logp(r, spec, &before, x, cfg, o4);
    }
    return evals;
}

/* ================================================================ global search */
static int pick_reduced_box(int need, int n) {
    const int sizes[] = {16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024};
    for (unsigned k = 0; k < sizeof sizes / sizeof sizes[0]; ++k)
        if (sizes[k] >= need) return sizes[k] < n ? sizes[k] : n;
    return n;
}

typedef struct { float score, sx, sy; int orient; } hit_t;

/* cisTEM-style global search (refine3d prompt 36): every grid orientation is scored by the peak
   of the cross-correlation map over shifts, computed by FFT in a box reduced to the search
   resolution; the K best hits are refined locally and the best refined pose wins. */
long long orc_global_search(const orc_ref *r, const float *specs, orc_row *rows, int n_img, const orc_refine_cfg *cfg,
                            const float *angles3, int n_orient) {
    const int n = cfg->box, nh = n / 2 + 1;
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    const float npx = (float)n * cfg->pixel_size;
    float r_s = cfg->search_high_res > 0.f ? npx / cfg->search_high_res : hi;
    if (r_s > hi) r_s = hi;
    if (r_s < lo + 2.f) r_s = fminf(lo + 2.f, hi);
    /* search samples */
    int n_ss = 0, i_max = 0;
    int *si = (int *)malloc(sizeof(int) * (size_t)n * nh), *sj = (int *)malloc(sizeof(int) * (size_t)n * nh);
    for (int j = -n / 2; j < n / 2; ++j)
        for (int i = 0; i <= n / 2; ++i) {
            const float r2 = (float)(i * i + j * j);
            if (r2 < lo * lo || r2 > hi * hi || r2 > r_s * r_s) continue;
            si[n_ss] = i; sj[n_ss] = j; ++n_ss;
            if (i > i_max) i_max = i;
            if (abs(j) > i_max) i_max = abs(j);
        }
    const int nb = pick_reduced_box(2 * (i_max + 2), n);
    const float red = (float)nb / (float)n;
    int wx = (int)ceilf(cfg->search_range_x / cfg->pixel_size * red), wy = (int)ceilf(cfg->search_range_y / cfg->pixel_size * red);
    const int wmax = nb / 2 - 2;
    if (wx < 1) wx = 1;
    if (wy < 1) wy = 1;
    if (wx > wmax) wx = wmax;
    if (wy > wmax) wy = wmax;
    const int wxs = 2 * wx + 1, wys = 2 * wy + 1;
    int K = cfg->best_matches > 0 ? cfg->best_matches : 20;
    if (K > n_orient) K = n_orient;
    const float shift_scale = cfg->pixel_size / red;
    const int nband = orc_band_count(cfg);
    /* slices are particle independent */
    float *Pall = (float *)malloc(sizeof(float) * 2 * (size_t)n_orient * n_ss);
    for (int o = 0; o < n_orient; ++o) {
        float m[9];
        orc_euler_matrix(angles3[3 * o], angles3[3 * o + 1], angles3[3 * o + 2], m);
        for (int s = 0; s < n_ss; ++s)
            ref_interp(r, (m[0] * si[s] + m[1] * sj[s]) * r->pad, (m[3] * si[s] + m[4] * sj[s]) * r->pad,
                       (m[6] * si[s] + m[7] * sj[s]) * r->pad, &Pall[2 * ((size_t)o * n_ss + s)], &Pall[2 * ((size_t)o * n_ss + s) + 1]);
    }
    long long evals = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : evals)
    for (int k = 0; k < n_img; ++k) {
        const float *spec = specs + 2 * (size_t)k * n * nh;
        orc_row *row = &rows[k];
        const ctfc c = ctf_make(row, n);
        float *G = (float *)malloc(sizeof(float) * 2 * (size_t)n_ss), *c2m = (float *)malloc(sizeof(float) * (size_t)n_ss);
        float A = 0.f;
        for (int s = 0; s < n_ss; ++s) {
            const int jj = sj[s] < 0 ? sj[s] + n : sj[s];
            const float fr = spec[2 * ((size_t)jj * nh + si[s])], fi = spec[2 * ((size_t)jj * nh + si[s]) + 1];
            const float ctf = ctf_eval(&c, si[s], sj[s], 0.f);
            const float mult = si[s] > 0 ? 2.f : 1.f;
            G[2 * s] = fr * ctf; G[2 * s + 1] = fi * ctf;
            c2m[s] = ctf * ctf * mult;
            A += mult * (fr * fr + fi * fi);
        }
        hit_t *top = (hit_t *)malloc(sizeof(hit_t) * (size_t)K);
        for (int t = 0; t < K; ++t) { top[t].score = -1e30f; top[t].sx = top[t].sy = 0.f; top[t].orient = -1; }
        cd *S = (cd *)malloc(sizeof(cd) * (size_t)nb * nb);
        for (int o = 0; o < n_orient; ++o) {
            const float *P = Pall + 2 * (size_t)o * n_ss;
            memset(S, 0, sizeof(cd) * (size_t)nb * nb);
            float B = 0.f;
            for (int s = 0; s < n_ss; ++s) {
                const float pr = P[2 * s], pi = P[2 * s + 1];
                const float xr = G[2 * s] * pr + G[2 * s + 1] * pi, xi = G[2 * s + 1] * pr - G[2 * s] * pi;
                const int i = si[s], j = sj[s];
                const int iy = j < 0 ? j + nb : j;
                S[(size_t)iy * nb + i].re = xr;
                S[(size_t)iy * nb + i].im = xi;
                if (i > 0) { /* Friedel mate fills the other half plane */
                    const int my = (nb - iy) % nb, mx = nb - i;
                    S[(size_t)my * nb + mx].re = xr;
                    S[(size_t)my * nb + mx].im = -xi;
                }
                B += c2m[s] * (pr * pr + pi * pi);
            }
            fft2_full(S, nb, +1);
            float best = -1e30f;
            int bidx = 0;
            for (int w = 0; w < wxs * wys; ++w) {
                const int dy = w / wxs - wy, dx = w % wxs - wx;
                const float v = (float)S[(size_t)((dy + nb) % nb) * nb + ((dx + nb) % nb)].re;
                if (v > best) { best = v; bidx = w; }
            }
            const int dy = bidx / wxs - wy, dx = bidx % wxs - wx;
#define AT(ddx, ddy) ((float)S[(size_t)(((ddy) + nb) % nb) * nb + (((ddx) + nb) % nb)].re)
            const float v0 = best, xm = AT(dx - 1, dy), xp = AT(dx + 1, dy), ym = AT(dx, dy - 1), yp = AT(dx, dy + 1);
#undef AT
            float ox = 0.f, oy = 0.f, peak = v0;
            const float cx = xm - 2.f * v0 + xp, cy = ym - 2.f * v0 + yp;
            if (cx < 0.f) { ox = 0.5f * (xm - xp) / cx; if (ox < -0.5f) ox = -0.5f; if (ox > 0.5f) ox = 0.5f; peak -= 0.25f * (xm - xp) * ox; }
            if (cy < 0.f) { oy = 0.5f * (ym - yp) / cy; if (oy < -0.5f) oy = -0.5f; if (oy > 0.5f) oy = 0.5f; peak -= 0.25f * (ym - yp) * oy; }
            const float den = A * B;
            const float score = den > 0.f ? 100.f * peak / sqrtf(den) : 0.f;
            if (score > top[K - 1].score) {
                int t = K - 1;
                while (t > 0 && score > top[t - 1].score) { top[t] = top[t - 1]; --t; }
                top[t].score = score; top[t].sx = ((float)dx + ox) * shift_scale; top[t].sy = ((float)dy + oy) * shift_scale; top[t].orient = o;
            }
        }
        long long ev = n_orient;
        /* refine every hit in all five pose parameters; keep the best */
        const int freem[NP] = {1, 1, 1, 1, 1, cfg->refine_defocus};
        float bestsc = -1e30f, bestobj = -1e30f, xb[NP] = {row->psi, row->theta, row->phi, row->x_shift, row->y_shift, 0.f}, ob[4] = {0, 0, 0, 0};
        for (int t = 0; t < K; ++t) {
            float x[NP] = {row->psi, row->theta, row->phi, row->x_shift, row->y_shift, 0.f}, o4[4];
            if (top[t].orient >= 0) {
                x[0] = angles3[3 * top[t].orient]; x[1] = angles3[3 * top[t].orient + 1]; x[2] = angles3[3 * top[t].orient + 2];
                x[3] = top[t].sx; x[4] = top[t].sy;
            }
            float obj;
            /* the default optimiser refines the hits too (SEMANTICS.md §7c); stencil optimiser with steps widened to the search
               resolution for optimizer = 1 or a defocus refinement */
            const float sc = cfg->optimizer == 0 && !cfg->refine_defocus ? refine_one_lm(r, spec, row, x, freem, cfg, o4, &ev, &obj)
                                                                         : refine_one(r, spec, row, x, freem, cfg, o4, &ev, hi / r_s, &obj);
            if (obj > bestobj) { bestobj = obj; bestsc = sc; memcpy(xb, x, sizeof xb); memcpy(ob, o4, sizeof ob); }
        }
        evals += ev;
        const orc_row before = *row;
        write_row(row, xb, bestsc, ob, nband, cfg->refine_defocus);
        if (cfg->focus_radius > 0.f) row->logp = orc_focus_logp(r, spec, &before, xb, cfg, ob);
        free(S); free(top); free(G); free(c2m);
    }
    free(Pall); free(si); free(sj);
    return evals;
}

/* ================================================================ CSP (external/CSP/csp)
 * Constrained refinement of tilt series: every projection pose is a function of its particle's
 * parameters (PPSI, PTHETA, PPHI, PSHIFT_X/Y/Z) and its tilt's parameters (TILTANG, TILTAXIS,
 * TSHIFT_X/Y) — composition pinned to src/pyp/analysis/geometry/core.py:1081-1217 by
 * tests/golden/csp_euler.npy — and the objective of an entity is the mean score of its
 * projections inside the exposure window (cistem_star_file.py:936-986, update_particle_score).
 * Modes as passed on argv (local_run.py:332-335,411-431; align/core.py:1015-1023):
 *   0 tilt angle+axis, 1 particle angles, 2 particle shifts, 3 tilt shifts, 4 tilt defocus offset,
 *   5 particle angles+shifts, 6 tilt angle+axis+shifts.  SEMANTICS.md §11. */
static void csp_decode(const float *m, float *psi, float *theta, float *phi) {
    const float sth = hypotf(m[6], m[7]);
    const float r2d = 180.f / PI_F;
    if (sth > 1e-6f) {
        *theta = atan2f(sth, m[8]) * r2d;
        *psi = atan2f(m[7], -m[6]) * r2d;
        *phi = atan2f(m[5], m[2]) * r2d;
    } else if (m[8] > 0.f) {
        *theta = 0.f; *psi = 0.f; *phi = atan2f(m[3], m[0]) * r2d;
    } else {
        *theta = 180.f; *psi = 0.f; *phi = atan2f(-m[3], -m[0]) * r2d;
    }
}

/* first two rows of Rz(axis) Ry(angle) */
static void csp_projector(float angle, float axis, float *a6) {
    const float d2r = PI_F / 180.f;
    const float c = cosf(angle * d2r), s = sinf(angle * d2r), cb = cosf(axis * d2r), sb = sinf(axis * d2r);
    a6[0] = cb * c; a6[1] = -sb; a6[2] = cb * s;
    a6[3] = sb * c; a6[4] = cb;  a6[5] = sb * s;
}

void orc_csp_compose(const orc_particle *p, const orc_particle *p0, const orc_tilt *t, const orc_tilt *t0,
                     const float *centre3, float pixel, float bx, float by, float *out5) {
    float e[9], q[9], m[9];
    orc_euler_matrix(-p->psi, -p->theta, -p->phi, e);
    const float d2r = PI_F / 180.f;
    const float c = cosf(t->angle * d2r), s = sinf(t->angle * d2r), cb = cosf(t->axis * d2r), sb = sinf(t->axis * d2r);
    /* Ry(-angle) Rz(-axis) */
    q[0] = c * cb; q[1] = c * sb; q[2] = -s;
    q[3] = -sb;    q[4] = cb;     q[5] = 0.f;
    q[6] = s * cb; q[7] = s * sb; q[8] = c;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) m[3 * a + b] = e[3 * a] * q[b] + e[3 * a + 1] * q[3 + b] + e[3 * a + 2] * q[6 + b];
    csp_decode(m, &out5[0], &out5[1], &out5[2]);
    float A[6], A0[6];
    csp_projector(t->angle, t->axis, A);
    csp_projector(t0->angle, t0->axis, A0);
    const float X[3] = {(p->x_position_3d - centre3[0]) * pixel, (p->y_position_3d - centre3[1]) * pixel,
                        (p->z_position_3d - centre3[2]) * pixel};
    const float v[3] = {X[0] - p->shift_x, X[1] - p->shift_y, X[2] - p->shift_z};
    const float v0[3] = {X[0] - p0->shift_x, X[1] - p0->shift_y, X[2] - p0->shift_z};
    out5[3] = bx + (A[0] * v[0] + A[1] * v[1] + A[2] * v[2]) - (A0[0] * v0[0] + A0[1] * v0[1] + A0[2] * v0[2]) + (t->shift_x - t0->shift_x);
    out5[4] = by + (A[3] * v[0] + A[4] * v[1] + A[5] * v[2]) - (A0[3] * v0[0] + A0[4] * v0[1] + A0[5] * v0[2]) + (t->shift_y - t0->shift_y);
}

static uint32_t csp_mix(uint32_t a) {
    a ^= a >> 16; a *= 0x7feb352du; a ^= a >> 15; a *= 0x846ca68bu; a ^= a >> 16;
    return a;
}
/* uniform in [-1, 1): counter-based, identical on the GPU */
static float csp_uniform(uint32_t seed, uint32_t ent, uint32_t k, uint32_t dim) {
    const uint32_t h = csp_mix(seed ^ csp_mix(ent * 0x9E3779B9u + k) ^ (dim * 0x85EBCA6Bu + 0x27d4eb2fu));
    return (float)(h >> 8) * (1.f / 8388608.f) - 1.f;
}

typedef struct {
    int kind;            /* 0 particle entity, 1 tilt entity */
    int freem[NP];
    float tol[NP], h[NP], gstep[NP];
} csp_plan;

static int csp_make_plan(const orc_refine_cfg *cfg, const orc_csp_cfg *c, csp_plan *pl) {
    float lo, hi;
    orc_band_limits(cfg, &lo, &hi);
    const float h_ang = 0.35f * 57.29578f / hi;
    const float h_shift = 0.07f * (float)cfg->box / hi * cfg->pixel_size;
    const float h_def = cfg->defocus_step > 0.f ? cfg->defocus_step : 50.f;
    memset(pl, 0, sizeof *pl);
    switch (c->mode) {
    case 1: case 2: case 5: pl->kind = 0; break;
    case 0: case 3: case 4: case 6: pl->kind = 1; break;
    default: return -1;
    }
    if (pl->kind == 0) {
        const float tol[NP] = {c->tol_particle_psi, c->tol_particle_theta, c->tol_particle_phi, c->tol_particle_shift,
                               c->tol_particle_shift, c->tol_particle_shift};
        const float h[NP] = {h_ang, h_ang, h_ang, h_shift, h_shift, h_shift};
        const float g[NP] = {c->angle_step, c->angle_step, c->angle_step, c->shift_step, c->shift_step, c->shift_step};
        memcpy(pl->tol, tol, sizeof tol); memcpy(pl->h, h, sizeof h); memcpy(pl->gstep, g, sizeof g);
        const int ang = c->mode == 1 || c->mode == 5, sh = c->mode == 2 || c->mode == 5;
        for (int m = 0; m < 3; ++m) { pl->freem[m] = ang; pl->freem[3 + m] = sh; }
    } else {
        const float tol[NP] = {c->tol_tilt_angle, c->tol_tilt_axis, c->tol_tilt_shift, c->tol_tilt_shift, c->tol_defocus, 0.f};
        const float h[NP] = {h_ang, h_ang, h_shift, h_shift, h_def, 0.f};
        const float g[NP] = {c->angle_step, c->angle_step, c->shift_step, c->shift_step, 0.25f * c->tol_defocus, 0.f};
        memcpy(pl->tol, tol, sizeof tol); memcpy(pl->h, h, sizeof h); memcpy(pl->gstep, g, sizeof g);
        const int ang = c->mode == 0 || c->mode == 6, sh = c->mode == 3 || c->mode == 6;
        pl->freem[0] = pl->freem[1] = ang;
        pl->freem[2] = pl->freem[3] = sh;
        pl->freem[4] = c->mode == 4;
    }
    return 0;
}

/* number of exhaustive candidates (candidate 0 = the input parameters) */
static long long csp_num_candidates(const csp_plan *pl, const orc_csp_cfg *c, int *counts) {
    if (!c->grid_search) {
        for (int m = 0; m < NP; ++m) counts[m] = 1;
        return c->random_evals > 0 ? c->random_evals : 0;
    }
    long long total = 1;
    for (int m = 0; m < NP; ++m) {
        counts[m] = 1;
        if (pl->freem[m] && pl->gstep[m] > 0.f && pl->tol[m] > 0.f) counts[m] = 2 * (int)floorf(pl->tol[m] / pl->gstep[m]) + 1;
        total *= counts[m];
    }
    return total;
}

static void csp_candidate(const csp_plan *pl, const orc_csp_cfg *c, const int *counts, const float *x0, uint32_t ent,
                          long long k, float *x) {
    memcpy(x, x0, NP * sizeof(float));
    if (k == 0) return;
    if (!c->grid_search) {
        for (int m = 0; m < NP; ++m)
            if (pl->freem[m]) x[m] = x0[m] + pl->tol[m] * csp_uniform(c->seed, ent, (uint32_t)k, (uint32_t)m);
    } else {
        /* candidate 0 is the centre; k >= 1 walks the lattice in mixed radix, skipping nothing
           (the centre is visited twice, harmless) */
        long long q = k - 1;
        for (int m = 0; m < NP; ++m) {
            const int d = (int)(q % counts[m]);
            q /= counts[m];
            x[m] = x0[m] + (float)(d - counts[m] / 2) * pl->gstep[m];
        }
    }
}

typedef struct {
    const orc_ref *r; const float *specs; const orc_row *rows; const orc_refine_cfg *cfg;
    const orc_particle *particles; const orc_tilt *tilts;   /* input tables (p0 / t0) */
    const int *row_part, *row_tilt;                         /* per row: index into the tables */
    const float *centre3;
    int kind, ent;                                          /* entity being refined */
    const int *members; int n_members;                      /* row indices, window rows first */
    int n_window;
} csp_group;

/* pose6 of member row k for entity parameters x */
static void csp_member_pose(const csp_group *g, int row, const float *x, float *pose6) {
    const orc_row *rw = &g->rows[row];
    orc_particle p = g->particles[g->row_part[row]];
    orc_tilt t = g->tilts[g->row_tilt[row]];
    const orc_particle p0 = p;
    const orc_tilt t0 = t;
    float ddef = 0.f;
    if (g->kind == 0) {
        p.psi = x[0]; p.theta = x[1]; p.phi = x[2]; p.shift_x = x[3]; p.shift_y = x[4]; p.shift_z = x[5];
    } else {
        t.angle = x[0]; t.axis = x[1]; t.shift_x = x[2]; t.shift_y = x[3]; ddef = x[4];
    }
    float o5[5];
    orc_csp_compose(&p, &p0, &t, &t0, g->centre3, rw->pixel_size, rw->x_shift, rw->y_shift, o5);
    memcpy(pose6, o5, 5 * sizeof(float));
    pose6[5] = ddef;
}

/* mean cc over the first `count` members; optionally keeps the per-member band sums */
static float csp_objective(const csp_group *g, const float *x, int count, float *o4_all, long long *evals) {
    const int n = g->cfg->box, nh = n / 2 + 1;
    float sum = 0.f;
    for (int k = 0; k < count; ++k) {
        const int row = g->members[k];
        float pose[6], o4[4];
        csp_member_pose(g, row, x, pose);
        const float sc = orc_score(g->r, g->specs + 2 * (size_t)row * n * nh, &g->rows[row], pose, g->cfg, o4);
        if (o4_all) memcpy(o4_all + 4 * k, o4, sizeof o4);
        if (k < g->n_window) sum += sc * 0.01f;
        (*evals)++;
    }
    return g->n_window > 0 ? sum / (float)g->n_window : 0.f;
}

static void csp_clamp(const csp_plan *pl, const float *x0, float *x) {
    for (int m = 0; m < NP; ++m) {
        if (!pl->freem[m] || pl->tol[m] <= 0.f) continue;
        if (x[m] > x0[m] + pl->tol[m]) x[m] = x0[m] + pl->tol[m];
        if (x[m] < x0[m] - pl->tol[m]) x[m] = x0[m] - pl->tol[m];
    }
}

long long orc_csp_run(const orc_ref *r, const float *specs, orc_row *rows, int n_rows, orc_particle *particles,
                      int n_particles, orc_tilt *tilts, int n_tilts, const orc_refine_cfg *cfg,
                      const orc_csp_cfg *csp, int first, int last) {
    csp_plan pl;
    if (csp_make_plan(cfg, csp, &pl)) return -1;
    const int nband = orc_band_count(cfg);
    /* row -> table indices */
    int *row_part = (int *)malloc(sizeof(int) * (size_t)(n_rows > 0 ? n_rows : 1));
    int *row_tilt = (int *)malloc(sizeof(int) * (size_t)(n_rows > 0 ? n_rows : 1));
    long long bad = 0;
    for (int k = 0; k < n_rows; ++k) {
        row_part[k] = row_tilt[k] = -1;
        for (int a = 0; a < n_particles; ++a)
            if (particles[a].pind == rows[k].pind) { row_part[k] = a; break; }
        for (int a = 0; a < n_tilts; ++a)
            if (tilts[a].tind == rows[k].tind && tilts[a].rind == rows[k].rind) { row_tilt[k] = a; break; }
        if (row_part[k] < 0 || row_tilt[k] < 0) bad++;
    }
    if (bad) { free(row_part); free(row_tilt); return -2; }
    float centre3[3] = {0.f, 0.f, 0.f};
    for (int a = 0; a < n_particles; ++a) {
        centre3[0] += particles[a].x_position_3d; centre3[1] += particles[a].y_position_3d; centre3[2] += particles[a].z_position_3d;
    }
    if (n_particles > 0) for (int d = 0; d < 3; ++d) centre3[d] /= (float)n_particles;
    /* the search reads the INPUT tables and rows; results go to copies */
    orc_row *rows_out = (orc_row *)malloc(sizeof(orc_row) * (size_t)(n_rows > 0 ? n_rows : 1));
    memcpy(rows_out, rows, sizeof(orc_row) * (size_t)n_rows);
    const int n_ent = pl.kind == 0 ? n_particles : n_tilts;
    float *ent_x = (float *)malloc(sizeof(float) * NP * (size_t)(n_ent > 0 ? n_ent : 1));
    float *ent_f = (float *)malloc(sizeof(float) * (size_t)(n_ent > 0 ? n_ent : 1));
    char *ent_done = (char *)calloc((size_t)(n_ent > 0 ? n_ent : 1), 1);
    int counts[NP];
    const long long n_cand = csp_num_candidates(&pl, csp, counts);
    int n_free = 0;
    for (int m = 0; m < NP; ++m) n_free += pl.freem[m];
    const int iters = n_free > 0 && csp->iterations > 0 ? csp->iterations : 0;
    const int late = iters / 2 + 1;
    long long evals = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : evals)
    for (int e = 0; e < n_ent; ++e) {
        const int id = pl.kind == 0 ? particles[e].pind : tilts[e].tind;
        if (id < first || (last >= 0 && id > last)) continue;
        /* members: window rows first, both in row order */
        int *members = (int *)malloc(sizeof(int) * (size_t)(n_rows > 0 ? n_rows : 1));
        int nm = 0, nw = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (int k = 0; k < n_rows; ++k) {
                if ((pl.kind == 0 ? row_part[k] : row_tilt[k]) != e) continue;
                const int t = rows[k].tind;
                const int in_w = !(t < csp->window_min || (csp->window_max != -1 && t > csp->window_max));
                if (in_w == (pass == 0)) { members[nm++] = k; if (pass == 0) nw++; }
            }
        if (nm == 0) { free(members); continue; }
        csp_group g = {r, specs, rows, cfg, particles, tilts, row_part, row_tilt, centre3, pl.kind, e, members, nm, nw};
        float x0[NP], x[NP];
        if (pl.kind == 0) {
            const orc_particle *p = &particles[e];
            const float v[NP] = {p->psi, p->theta, p->phi, p->shift_x, p->shift_y, p->shift_z};
            memcpy(x0, v, sizeof v);
        } else {
            const orc_tilt *t = &tilts[e];
            const float v[NP] = {t->angle, t->axis, t->shift_x, t->shift_y, 0.f, 0.f};
            memcpy(x0, v, sizeof v);
        }
        memcpy(x, x0, sizeof x0);
        long long ev = 0;
        const int active = nw > 0 && nw >= csp->min_projections;
        if (active) {
            /* exhaustive stage: best of n_cand candidates, ties keep the lowest index */
            float best = -1e30f;
            long long kbest = 0;
            for (long long k = 0; k < n_cand; ++k) {
                float xc[NP];
                csp_candidate(&pl, csp, counts, x0, (uint32_t)id, k, xc);
                const float f = csp_objective(&g, xc, nw, NULL, &ev);
                if (f > best) { best = f; kbest = k; }
            }
            if (n_cand > 0) csp_candidate(&pl, csp, counts, x0, (uint32_t)id, kbest, x);
            /* local stage: the batched stencil/Newton/line-search optimiser of refine3d in entity space */
            float h[NP], d[NP], q[NP];
            memcpy(h, pl.h, sizeof h);
            for (int it = 0; it < iters; ++it) {
                const float f0 = csp_objective(&g, x, nw, NULL, &ev);
                for (int m = 0; m < NP; ++m) {
                    d[m] = 0.f;
                    if (!pl.freem[m]) continue;
                    memcpy(q, x, sizeof q);
                    q[m] = x[m] + h[m];
                    const float fp = csp_objective(&g, q, nw, NULL, &ev);
                    q[m] = x[m] - h[m];
                    const float fm = csp_objective(&g, q, nw, NULL, &ev);
                    d[m] = newton_step(f0, fp, fm, h[m]);
                }
                float fl[NL];
                for (int l = 0; l < NL; ++l) {
                    for (int m = 0; m < NP; ++m) q[m] = x[m] + LS_T[l] * d[m];
                    fl[l] = csp_objective(&g, q, nw, NULL, &ev);
                }
                const float t = line_step(f0, fl);
                for (int m = 0; m < NP; ++m) x[m] += t * d[m];
                csp_clamp(&pl, x0, x);
                if (it + 1 >= late)
                    for (int m = 0; m < NP; ++m) h[m] *= 0.6f;
            }
        }
        /* final: refined and input parameters over ALL members; never return a worse objective */
        float *o4a = (float *)malloc(sizeof(float) * 4 * (size_t)nm), *o4b = (float *)malloc(sizeof(float) * 4 * (size_t)nm);
        float fa = csp_objective(&g, x, nm, o4a, &ev);
        const float fb = csp_objective(&g, x0, nm, o4b, &ev);
        if (fa < fb) { memcpy(x, x0, sizeof x0); memcpy(o4a, o4b, sizeof(float) * 4 * (size_t)nm); fa = fb; }
        for (int k = 0; k < nm; ++k) {
            const int row = members[k];
            float pose[6];
            csp_member_pose(&g, row, x, pose);
            orc_row *o = &rows_out[row];
            o->psi = wrap360(pose[0]); o->theta = pose[1]; o->phi = wrap360(pose[2]);
            o->x_shift = pose[3]; o->y_shift = pose[4];
            o->defocus_1 += pose[5]; o->defocus_2 += pose[5];
            const float *v = o4a + 4 * k;
            const float den = v[2] * v[3];
            o->score = den > 0.f ? 100.f * v[0] / sqrtf(den) : 0.f;
            stats_from(v, nband, &o->sigma, &o->logp);
        }
        memcpy(ent_x + NP * (size_t)e, x, sizeof x);
        ent_f[e] = fa;
        ent_done[e] = 1;
        evals += ev;
        free(o4a); free(o4b); free(members);
    }
    memcpy(rows, rows_out, sizeof(orc_row) * (size_t)n_rows);
    for (int e = 0; e < n_ent; ++e) {
        if (!ent_done[e]) continue;
        const float *x = ent_x + NP * (size_t)e;
        if (pl.kind == 0) {
            orc_particle *p = &particles[e];
            p->psi = x[0]; p->theta = x[1]; p->phi = x[2]; p->shift_x = x[3]; p->shift_y = x[4]; p->shift_z = x[5];
            p->score = 100.f * ent_f[e];
        } else {
            orc_tilt *t = &tilts[e];
            t->angle = x[0]; t->axis = x[1]; t->shift_x = x[2]; t->shift_y = x[3];
        }
    }
    free(rows_out); free(ent_x); free(ent_f); free(ent_done); free(row_part); free(row_tilt);
    return evals;
}

/* ================================================================ reconstruction */
struct orc_recon {
    orc_recon_cfg cfg;
    int np, xh;
    float *acc[2]; /* float4 per voxel, centred y,z */
};

orc_recon *orc_recon_create(const orc_recon_cfg *cfg) {
    orc_recon *rc = (orc_recon *)calloc(1, sizeof(orc_recon));
    rc->cfg = *cfg;
    rc->np = cfg->box * cfg->pad;
    rc->xh = rc->np / 2 + 1;
    for (int h = 0; h < 2; ++h) rc->acc[h] = (float *)calloc((size_t)rc->xh * rc->np * rc->np * 4, sizeof(float));
    return rc;
}

void orc_recon_free(orc_recon *rc) { if (rc) { free(rc->acc[0]); free(rc->acc[1]); free(rc); } }

static void add_corner(orc_recon *rc, float *acc, int x, int y, int z, float w, float re, float im, float wt) {
    const int c = rc->np / 2;
    if (x > c || y < -c || y >= c || z < -c || z >= c) return;
    float *p = acc + 4 * (((size_t)(z + c) * rc->np + (y + c)) * rc->xh + x);
    p[0] += w * re;
    p[1] += w * im;
    p[2] += w * wt;
}

void orc_recon_insert(orc_recon *rc, const float *imgs, const orc_row *rows, int count, const float *sym, int n_sym) {
    orc_recon_insert_weighted(rc, imgs, rows, count, sym, n_sym, NULL);
}

/* `aux` (may be NULL): two floats per row {weight, cut radius in Fourier pixels} of the data-driven dose weighting
 * (SEMANTICS.md §10): the row's samples are weighted by `weight` and, beyond the cut radius, by a raised-cosine edge of
 * width DOSE_EDGE * n/2 centred on it (cut radius <= 0: no low-pass) */
#define DOSE_EDGE 0.1f
void orc_recon_insert_weighted(orc_recon *rc, const float *imgs, const orc_row *rows, int count, const float *sym, int n_sym, const float *aux) {
    /* cisTEM Reconstruct3D::InsertSliceWithCTF shape: i = 0 column only for j >= 0 */
    const orc_recon_cfg *c = &rc->cfg;
    const int n = c->box, nh = n / 2 + 1;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)n * n);
    float *spec = (float *)malloc(sizeof(float) * 2 * (size_t)n * nh);
    float rmax = (float)n * c->pixel_size / (c->resolution_limit > 0.f ? c->resolution_limit : 2.f * c->pixel_size);
    if (rmax > (float)(n / 2 - 1)) rmax = (float)(n / 2 - 1); /* no weight ever lands on the Nyquist planes */
    const float l = (float)n * c->pixel_size, s2u = 1.f / (l * l);
    const float bk = c->score_weighting ? c->score_bfactor * 0.25f * s2u : 0.f;
    const float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!sym || n_sym < 1) { sym = id; n_sym = 1; }
    for (int k = 0; k < count; ++k) {
        const orc_row *row = &rows[k];
        if (!(row->occupancy > 0.f) || row->score < c->score_threshold) continue;
        normalized_copy(imgs + (size_t)k * n * n, n, c->mask_radius / c->pixel_size, c->normalize, c->invert_contrast, tmp);
        g_fft2_r2c(tmp, n, spec);
        const ctfc cc = ctf_make(row, n);
        float m[9];
        orc_euler_matrix(row->psi, row->theta, row->phi, m);
        const int half = c->per_particle_split ? (row->pind & 1) : ((row->position_in_stack & 1u) ? 0 : 1);
        float *acc = rc->acc[half];
        const float k2 = 2.f * PI_F / ((float)n * row->pixel_size);
        for (int j = -n / 2; j < n / 2; ++j)
            for (int i = 0; i <= n / 2; ++i) {
                if (i == 0 && j < 0) continue;
                const float r2 = (float)(i * i + j * j);
                if (r2 > rmax * rmax) continue;
                const int jj = j < 0 ? j + n : j;
                const float sg = ((i + j) & 1) ? -1.f : 1.f;
                const float fr = sg * spec[2 * ((size_t)jj * nh + i)], fim = sg * spec[2 * ((size_t)jj * nh + i) + 1];
                const float ctf = ctf_eval(&cc, i, j, 0.f);
                float w = row->occupancy * 0.01f;
                if (bk != 0.f) w *= expf(-bk * (c->average_score - row->score) * r2);
                if (aux) {
                    w *= aux[2 * k];
                    if (aux[2 * k + 1] > 0.f) w *= cos_edge(sqrtf(r2), aux[2 * k + 1], DOSE_EDGE * 0.5f * (float)n);
                }
                const float ph = (i * row->x_shift + j * row->y_shift) * k2;
                const float cs = cosf(ph), sn = sinf(ph);
                const float re = (fr * cs - fim * sn) * ctf, im = (fr * sn + fim * cs) * ctf;
                const float wt = ctf * ctf;
                for (int s = 0; s < n_sym; ++s) {
                    const float *S = sym + 9 * s;
                    float R[9];
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b) R[3 * a + b] = S[3 * a] * m[b] + S[3 * a + 1] * m[3 + b] + S[3 * a + 2] * m[6 + b];
                    float x = (R[0] * i + R[1] * j) * c->pad, y = (R[3] * i + R[4] * j) * c->pad, z = (R[6] * i + R[7] * j) * c->pad;
                    float vim = im;
                    if (x < 0.f) { x = -x; y = -y; z = -z; vim = -im; }
                    const int x0 = (int)floorf(x), y0 = (int)floorf(y), z0 = (int)floorf(z);
                    const float fx = x - x0, fy = y - y0, fz = z - z0;
                    for (int dz = 0; dz < 2; ++dz)
                        for (int dy = 0; dy < 2; ++dy)
                            for (int dx = 0; dx < 2; ++dx) {
                                const float ww = (dx ? fx : 1.f - fx) * (dy ? fy : 1.f - fy) * (dz ? fz : 1.f - fz);
                                add_corner(rc, acc, x0 + dx, y0 + dy, z0 + dz, w * ww, re, vim, wt);
                            }
                }
            }
    }
    free(tmp);
    free(spec);
}

void orc_recon_get_dump(const orc_recon *rc, int half, float *out) {
    memcpy(out, rc->acc[half], sizeof(float) * 4 * (size_t)rc->xh * rc->np * rc->np);
}

static void symmetrize_x0(float *acc, int np, int xh) {
    const int c = np / 2;
    for (int z = -c + 1; z < c; ++z)
        for (int y = -c + 1; y < c; ++y) {
            if (!(z > 0 || (z == 0 && y >= 0))) continue;
            float *p = acc + 4 * (((size_t)(z + c) * np + (y + c)) * xh);
            float *q = acc + 4 * (((size_t)(-z + c) * np + (-y + c)) * xh);
            const float re = p[0] + q[0], im = p[1] - q[1], w = p[2] + q[2];
            p[0] = re; p[1] = im; p[2] = w;
            q[0] = re; q[1] = -im; q[2] = w;
        }
}

void orc_recon_finalize(orc_recon *rc, float mw_kda, float outer_radius_a, float *half1, float *half2, float *map, float *stats) {
    const orc_recon_cfg *c = &rc->cfg;
    const int n = c->box, np = rc->np, xh = rc->xh, ns = n / 2 + 1, cen = np / 2;
    const size_t nvox = (size_t)xh * np * np;
    float *a[2];
    for (int h = 0; h < 2; ++h) {
        a[h] = (float *)malloc(sizeof(float) * 4 * nvox);
        memcpy(a[h], rc->acc[h], sizeof(float) * 4 * nvox);
        symmetrize_x0(a[h], np, xh);
    }
    double *sh = (double *)calloc((size_t)ns * 7, sizeof(double));
    for (int z = -cen; z < cen; ++z)
        for (int y = -cen; y < cen; ++y)
            for (int x = 0; x < xh; ++x) {
                const float r = sqrtf((float)(x * x + y * y + z * z)) / (float)c->pad;
                const int s = (int)(r + 0.5f);
                if (s >= ns) continue;
                const size_t o = 4 * (((size_t)(z + cen) * np + (y + cen)) * xh + x);
                const float *p = a[0] + o, *q = a[1] + o;
                if (!(p[2] > 0.f) || !(q[2] > 0.f)) continue;
                const double mult = x == 0 ? 0.5 : 1.0;
                const float v1x = p[0] / p[2], v1y = p[1] / p[2], v2x = q[0] / q[2], v2y = q[1] / q[2];
                double *d = sh + (size_t)s * 7;
                d[0] += mult * (v1x * v2x + v1y * v2y);
                d[1] += mult * (v1x * v1x + v1y * v1y);
                d[2] += mult * (v2x * v2x + v2y * v2y);
                d[3] += mult * (p[2] + q[2]);
                d[4] += mult;
                d[5] += mult * p[2];
                d[6] += mult * q[2];
            }
    const float box_a = (float)n * c->pixel_size;
    const float rad_a = outer_radius_a > 0.f ? outer_radius_a : 0.5f * box_a;
    double mask_vol = 4.0 / 3.0 * PI_D * (double)rad_a * rad_a * rad_a;
    const double box_vol = (double)box_a * box_a * box_a;
    if (mask_vol > box_vol) mask_vol = box_vol;
    double frac = mw_kda > 0.f ? ((double)mw_kda * 1000.0 / 0.81) / mask_vol : 1.0;
    if (frac > 1.0 || frac <= 0.0) frac = 1.0;
    float *term = (float *)calloc((size_t)3 * ns, sizeof(float));
    for (int s = 0; s < ns; ++s) {
        const double *q = sh + (size_t)s * 7;
        double fsc = 0.0;
        if (q[1] > 0.0 && q[2] > 0.0) fsc = q[0] / sqrt(q[1] * q[2]);
        if (s == 0 && q[4] > 0.0) fsc = 1.0;
        double f = fsc;
        if (f > 0.9999) f = 0.9999;
        if (f < 0.0) f = 0.0;
        const double rec = 2.0 * f / (1.0 - f), part = rec / frac, pfsc = part / (2.0 + part);
        const double cnt = q[4] > 0.0 ? q[4] : 1.0;
        const double floor_ = 1e-4;
        term[s] = (float)((q[3] / cnt) / (rec > floor_ ? rec : floor_));
        const double hs = 0.5 * rec > floor_ ? 0.5 * rec : floor_;
        term[ns + s] = (float)((q[5] / cnt) / hs);
        term[2 * ns + s] = (float)((q[6] / cnt) / hs);
        if (stats) {
            float *o = stats + (size_t)s * 7;
            o[0] = (float)s; o[1] = s > 0 ? box_a / (float)s : 0.f; o[2] = (float)s / (float)n;
            o[3] = (float)fsc; o[4] = (float)pfsc; o[5] = (float)sqrt(part); o[6] = (float)sqrt(rec);
        }
    }
    float *outs[3] = {map, half1, half2};
    cd *d = (cd *)malloc(sizeof(cd) * (size_t)np * np * np);
    for (int mode = 0; mode < 3; ++mode) {
        if (!outs[mode]) continue;
        for (int iz = 0; iz < np; ++iz)
            for (int iy = 0; iy < np; ++iy)
                for (int ix = 0; ix < np; ++ix) {
                    /* full Hermitian cube from the stored half */
                    const int x = ix > cen ? ix - np : ix;
                    const int y = iy >= cen ? iy - np : iy, z = iz >= cen ? iz - np : iz;
                    float sg = 1.f;
                    int hx = x, hy = y, hz = z;
                    if (x < 0) { hx = -x; hy = -y; hz = -z; sg = -1.f; }
                    if (hy == cen) hy = -cen; /* Nyquist index is its own mate */
                    if (hz == cen) hz = -cen;
                    double re = 0, im = 0;
                    if (hy >= -cen && hy < cen && hz >= -cen && hz < cen) {
                        const float r = sqrtf((float)(hx * hx + hy * hy + hz * hz)) / (float)c->pad;
                        const int s = (int)(r + 0.5f);
                        if (s < ns) {
                            const size_t o = 4 * (((size_t)(hz + cen) * np + (hy + cen)) * xh + hx);
                            float vr, vi, w;
                            if (mode == 0) { vr = a[0][o] + a[1][o]; vi = a[0][o + 1] + a[1][o + 1]; w = a[0][o + 2] + a[1][o + 2]; }
                            else { const float *p = a[mode - 1] + o; vr = p[0]; vi = p[1]; w = p[2]; }
                            const float den = w + term[(size_t)mode * ns + s];
                            if (w > 0.f && den > 0.f) {
                                const float cs = ((hx + hy + hz) & 1) ? -1.f : 1.f;
                                re = cs * vr / den;
                                im = sg * cs * vi / den;
                            }
                        }
                    }
                    d[((size_t)iz * np + iy) * np + ix].re = re;
                    d[((size_t)iz * np + iy) * np + ix].im = im;
                }
        fft3_full(d, np, +1);
        const int off = (np - n) / 2;
        const float scale = 1.f / ((float)np * (float)np * (float)np);
        const float rad = rad_a / c->pixel_size, wid = 20.f / c->pixel_size;
        for (int z = 0; z < n; ++z)
            for (int y = 0; y < n; ++y)
                for (int x = 0; x < n; ++x) {
                    const int dx = x - n / 2, dy = y - n / 2, dz = z - n / 2;
                    float v = (float)d[((size_t)(z + off) * np + (y + off)) * np + (x + off)].re * scale;
                    v /= sinc2c(dx, np) * sinc2c(dy, np) * sinc2c(dz, np);
                    v *= cos_edge(sqrtf((float)(dx * dx + dy * dy + dz * dz)), rad, wid);
                    outs[mode][((size_t)z * n + y) * n + x] = v;
                }
    }
    free(d);
    free(term);
    free(sh);
    free(a[0]);
    free(a[1]);
}

void orc_fsc(const float *va, const float *vb, int n, float *fsc_out) {
    const int ns = n / 2 + 1;
    cd *a = (cd *)malloc(sizeof(cd) * (size_t)n * n * n), *b = (cd *)malloc(sizeof(cd) * (size_t)n * n * n);
    for (size_t k = 0; k < (size_t)n * n * n; ++k) { a[k].re = va[k]; a[k].im = 0; b[k].re = vb[k]; b[k].im = 0; }
    fft3_full(a, n, -1);
    fft3_full(b, n, -1);
    double *s = (double *)calloc((size_t)ns * 3, sizeof(double));
    for (int iz = 0; iz < n; ++iz)
        for (int iy = 0; iy < n; ++iy)
            for (int ix = 0; ix < n; ++ix) {
                const int x = ix >= n / 2 ? ix - n : ix, y = iy >= n / 2 ? iy - n : iy, z = iz >= n / 2 ? iz - n : iz;
                const int sh = (int)(sqrt((double)(x * x + y * y + z * z)) + 0.5);
                if (sh >= ns) continue;
                const cd p = a[((size_t)iz * n + iy) * n + ix], q = b[((size_t)iz * n + iy) * n + ix];
                s[3 * sh] += p.re * q.re + p.im * q.im;
                s[3 * sh + 1] += p.re * p.re + p.im * p.im;
                s[3 * sh + 2] += q.re * q.re + q.im * q.im;
            }
    for (int k = 0; k < ns; ++k) fsc_out[k] = (s[3 * k + 1] > 0 && s[3 * k + 2] > 0) ? (float)(s[3 * k] / sqrt(s[3 * k + 1] * s[3 * k + 2])) : 0.f;
    free(s); free(a); free(b);
}
