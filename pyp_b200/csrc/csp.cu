// csp.cu — constrained single-particle (tilt-series) refinement on top of the scorer.
//
// Replaces the numerics of external/CSP/csp (closed LFS binary).  Contract: argv built at
// src/pyp/system/local_run.py:306-467, driver src/pyp/align/core.py:883-1248, extended tables
// src/pyp/inout/metadata/cistem_star_file.py:247-248.  The pose of projection (particle p, tilt t)
// is composed exactly as src/pyp/analysis/geometry/core.py:1081-1217 does (pinned by
// tests/golden/csp_euler.npy); the objective of an entity is the mean score of its projections in
// the exposure window (cistem_star_file.py:936-986).  Search strategy and every other choice:
// oracle/SEMANTICS.md §11, restated on the CPU in oracle/cspb_oracle.c (orc_csp_run).
//
// Device layout: one optimiser state per entity (particle or tilt).  A "list" is the concatenation
// of the member rows of all selected entities (window members first inside each entity).  Per
// evaluation round every (list entry, candidate) pair is expanded into a pose of the scorer
// (csp_expand_kernel), scored by the same score_kernel refine3d uses, and folded back into one
// objective value per (entity, candidate) (csp_reduce_kernel).
#include <math.h>
#include <string.h>
#include <unordered_map>
#include "device_math.cuh"
#include "internal.cuh"
#include "opt.cuh"

namespace {

struct CspEntry {
    int row;    // index of the projection row / loaded image
    int group;  // entity (state) index
};

struct CspGroup {
    int ent;         // index into the particle / tilt table
    int id;          // PIND / TIND (seeds the candidate generator)
    int off_search;  // first entry in the search list
    int n_search;    // window members that enter the objective (0 = entity is not refined)
    int off_all;     // first entry in the all-members list
    int n_all;
    int n_window;    // window members among n_all (they come first)
    int pad_;
};

struct CspPlan {
    int kind;  // 0 particle entity, 1 tilt entity
    int free_mask;
    float tol[OPT_NP], h[OPT_NP], gstep[OPT_NP];
    int counts[OPT_NP];
    int grid_search;
    uint32_t seed;
};

struct CspTables {
    const cspb_row *rows;
    const cspb_particle *particles;
    const cspb_tilt *tilts;
    const int *row_part;
    const int *row_tilt;
    float cx, cy, cz;
};

__host__ __device__ __forceinline__ void csp_decode(const float *m, float *psi, float *theta, float *phi) {
    const float sth = hypotf(m[6], m[7]);
    const float r2d = 180.f / CSPB_PI_F;
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

// first two rows of Rz(axis) Ry(angle)
__host__ __device__ __forceinline__ void csp_projector(float angle, float axis, float *a6) {
    const float d2r = CSPB_PI_F / 180.f;
    const float c = cosf(angle * d2r), s = sinf(angle * d2r), cb = cosf(axis * d2r), sb = sinf(axis * d2r);
    a6[0] = cb * c; a6[1] = -sb; a6[2] = cb * s;
    a6[3] = sb * c; a6[4] = cb;  a6[5] = sb * s;
}

// M_row = M(-PPSI,-PTHETA,-PPHI) Ry(-TILTANG) Rz(-TILTAXIS); shift = base + A(t)(X - p) - A(t0)(X - p0) + dT
__host__ __device__ __forceinline__ void csp_compose(const cspb_particle &p, const cspb_particle &p0, const cspb_tilt &t,
                                                     const cspb_tilt &t0, float cx, float cy, float cz, float pixel,
                                                     float bx, float by, float *out5) {
    float e[9], q[9], m[9];
    euler_matrix(-p.psi, -p.theta, -p.phi, e);
    const float d2r = CSPB_PI_F / 180.f;
    const float c = cosf(t.angle * d2r), s = sinf(t.angle * d2r), cb = cosf(t.axis * d2r), sb = sinf(t.axis * d2r);
    q[0] = c * cb; q[1] = c * sb; q[2] = -s;
    q[3] = -sb;    q[4] = cb;     q[5] = 0.f;
    q[6] = s * cb; q[7] = s * sb; q[8] = c;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) m[3 * a + b] = e[3 * a] * q[b] + e[3 * a + 1] * q[3 + b] + e[3 * a + 2] * q[6 + b];
    csp_decode(m, &out5[0], &out5[1], &out5[2]);
    float A[6], A0[6];
    csp_projector(t.angle, t.axis, A);
    csp_projector(t0.angle, t0.axis, A0);
    const float X[3] = {(p.x_position_3d - cx) * pixel, (p.y_position_3d - cy) * pixel, (p.z_position_3d - cz) * pixel};
    const float v[3] = {X[0] - p.shift_x, X[1] - p.shift_y, X[2] - p.shift_z};
    const float v0[3] = {X[0] - p0.shift_x, X[1] - p0.shift_y, X[2] - p0.shift_z};
    out5[3] = bx + (A[0] * v[0] + A[1] * v[1] + A[2] * v[2]) - (A0[0] * v0[0] + A0[1] * v0[1] + A0[2] * v0[2]) + (t.shift_x - t0.shift_x);
    out5[4] = by + (A[3] * v[0] + A[4] * v[1] + A[5] * v[2]) - (A0[3] * v0[0] + A0[4] * v0[1] + A0[5] * v0[2]) + (t.shift_y - t0.shift_y);
}

__host__ __device__ __forceinline__ uint32_t csp_mix(uint32_t a) {
    a ^= a >> 16; a *= 0x7feb352du; a ^= a >> 15; a *= 0x846ca68bu; a ^= a >> 16;
    return a;
}
// uniform in [-1, 1), counter based (identical in the oracle)
__host__ __device__ __forceinline__ float csp_uniform(uint32_t seed, uint32_t ent, uint32_t k, uint32_t dim) {
    const uint32_t h = csp_mix(seed ^ csp_mix(ent * 0x9E3779B9u + k) ^ (dim * 0x85EBCA6Bu + 0x27d4eb2fu));
    return (float)(h >> 8) * (1.f / 8388608.f) - 1.f;
}

__device__ __forceinline__ void csp_member_pose(const CspTables &T, int kind, int row, const float *x, float *pose6) {
    const cspb_row rw = T.rows[row];
    cspb_particle p = T.particles[T.row_part[row]];
    cspb_tilt t = T.tilts[T.row_tilt[row]];
    const cspb_particle p0 = p;
    const cspb_tilt t0 = t;
    float ddef = 0.f;
    if (kind == 0) {
        p.psi = x[0]; p.theta = x[1]; p.phi = x[2]; p.shift_x = x[3]; p.shift_y = x[4]; p.shift_z = x[5];
    } else {
        t.angle = x[0]; t.axis = x[1]; t.shift_x = x[2]; t.shift_y = x[3]; ddef = x[4];
    }
    csp_compose(p, p0, t, t0, T.cx, T.cy, T.cz, rw.pixel_size, rw.x_shift, rw.y_shift, pose6);
    pose6[5] = ddef;
}

__global__ void csp_init_kernel(const CspGroup *__restrict__ groups, int n_groups, CspTables T, CspPlan pl,
                                OptState *__restrict__ st) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    OptState s;
    if (pl.kind == 0) {
        const cspb_particle p = T.particles[groups[g].ent];
        s.x[0] = p.psi; s.x[1] = p.theta; s.x[2] = p.phi; s.x[3] = p.shift_x; s.x[4] = p.shift_y; s.x[5] = p.shift_z;
    } else {
        const cspb_tilt t = T.tilts[groups[g].ent];
        s.x[0] = t.angle; s.x[1] = t.axis; s.x[2] = t.shift_x; s.x[3] = t.shift_y; s.x[4] = 0.f; s.x[5] = 0.f;
    }
    for (int m = 0; m < OPT_NP; ++m) { s.h[m] = pl.h[m]; s.d[m] = 0.f; s.x0[m] = s.x[m]; }
    s.f = 0.f; s.lam = 0.f; s.pad_[0] = s.pad_[1] = 0.f;
    st[g] = s;
}

// exhaustive-stage candidates k0 .. k0+nc-1 of every group -> params[(g*nc + c)*6]
__global__ void csp_candidates_kernel(const OptState *__restrict__ st, const CspGroup *__restrict__ groups, int n_groups,
                                      CspPlan pl, long long k0, int nc, float *__restrict__ params) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_groups * nc) return;
    const int g = idx / nc, c = idx % nc;
    const long long k = k0 + c;
    const OptState s = st[g];
    float x[OPT_NP];
    for (int m = 0; m < OPT_NP; ++m) x[m] = s.x0[m];
    if (k > 0) {
        if (!pl.grid_search) {
            for (int m = 0; m < OPT_NP; ++m)
                if ((pl.free_mask >> m) & 1) x[m] = s.x0[m] + pl.tol[m] * csp_uniform(pl.seed, (uint32_t)groups[g].id, (uint32_t)k, (uint32_t)m);
        } else {
            long long q = k - 1;
            for (int m = 0; m < OPT_NP; ++m) {
                const int d = (int)(q % pl.counts[m]);
                q /= pl.counts[m];
                x[m] = s.x0[m] + (float)(d - pl.counts[m] / 2) * pl.gstep[m];
            }
        }
    }
    for (int m = 0; m < OPT_NP; ++m) params[(long long)idx * 6 + m] = x[m];
}

// every (list entry, candidate) -> pose of the scorer; units of <= PB candidates per entry.  The last
// nS candidates vary pure shifts only (same rotation and CTF as the centre): they form "shared" units
// scored from one gather.  Unit layout by class as in opt.cuh: [A full][A tail][S full][S tail].
__global__ void csp_expand_kernel(const CspEntry *__restrict__ list, int n_entries, int nc, int nS, int PB, CspTables T, int kind,
                                  const float *__restrict__ params, float *__restrict__ poses6, ScoreUnit *__restrict__ units) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)n_entries * nc) return;
    const int j = (int)(idx / nc), c = (int)(idx % nc);
    const CspEntry en = list[j];
    float x[OPT_NP];
    for (int m = 0; m < OPT_NP; ++m) x[m] = params[((long long)en.group * nc + c) * 6 + m];
    float pose[6];
    csp_member_pose(T, kind, en.row, x, pose);
    for (int m = 0; m < 6; ++m) poses6[idx * 6 + m] = pose[m];
    const int nA = nc - nS;
    const bool shared = c >= nA;
    const int cl = shared ? c - nA : c, ncl = shared ? nS : nA;  // candidate index / count inside its class
    if (cl % PB == 0) {
        const int fullA = nA / PB, tailA = nA % PB ? 1 : 0, fullS = nS / PB;
        const int nfull = ncl / PB, cu = cl / PB;
        long long base = shared ? (long long)n_entries * (fullA + tailA) : 0;
        ScoreUnit un;
        un.image = en.row;
        un.first_eval = (int)idx;
        un.count = min(PB, ncl - cl);
        un.pad_ = shared ? 1 : 0;
        if (cu < nfull) units[base + (long long)j * nfull + cu] = un;
        else units[base + (long long)n_entries * (shared ? fullS : fullA) + j] = un;
    }
}

// objective of (group, candidate) = mean cc over the first n_mean entries of the group, in list order
__global__ void csp_reduce_kernel(const CspGroup *__restrict__ groups, int n_groups, int nc, int use_all,
                                  const float4 *__restrict__ sc, float4 *__restrict__ obj) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_groups * nc) return;
    const int g = idx / nc, c = idx % nc;
    const CspGroup gr = groups[g];
    const int off = use_all ? gr.off_all : gr.off_search;
    const int n_mean = use_all ? gr.n_window : gr.n_search;
    float sum = 0.f;
    for (int k = 0; k < n_mean; ++k) sum += cc_of(sc[(long long)(off + k) * nc + c]);
    obj[idx] = make_float4(n_mean > 0 ? sum / (float)n_mean : 0.f, 0.f, 1.f, 1.f);
}

struct CspBest {
    float f;
    int pad_;
    long long k;
};

__global__ void csp_best_kernel(const float4 *__restrict__ obj, int n_groups, int nc, long long k0, int first_chunk,
                                CspBest *__restrict__ best) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    CspBest b;
    if (first_chunk) { b.f = -1e30f; b.k = 0; b.pad_ = 0; } else b = best[g];
    for (int c = 0; c < nc; ++c) {
        const float f = obj[(long long)g * nc + c].x;
        if (f > b.f) { b.f = f; b.k = k0 + c; }
    }
    best[g] = b;
}

// x <- best candidate (regenerated from its index)
__global__ void csp_take_best_kernel(OptState *__restrict__ st, const CspGroup *__restrict__ groups, int n_groups, CspPlan pl,
                                     const CspBest *__restrict__ best) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups || groups[g].n_search <= 0) return;
    OptState s = st[g];
    const long long k = best[g].k;
    if (k > 0) {
        if (!pl.grid_search) {
            for (int m = 0; m < OPT_NP; ++m)
                if ((pl.free_mask >> m) & 1) s.x[m] = s.x0[m] + pl.tol[m] * csp_uniform(pl.seed, (uint32_t)groups[g].id, (uint32_t)k, (uint32_t)m);
        } else {
            long long q = k - 1;
            for (int m = 0; m < OPT_NP; ++m) {
                const int d = (int)(q % pl.counts[m]);
                q /= pl.counts[m];
                s.x[m] = s.x0[m] + (float)(d - pl.counts[m] / 2) * pl.gstep[m];
            }
        }
    }
    st[g] = s;
}

__global__ void csp_clamp_kernel(OptState *__restrict__ st, int n_groups, CspPlan pl) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    OptState s = st[g];
    for (int m = 0; m < OPT_NP; ++m) {
        if (!((pl.free_mask >> m) & 1) || pl.tol[m] <= 0.f) continue;
        s.x[m] = fminf(fmaxf(s.x[m], s.x0[m] - pl.tol[m]), s.x0[m] + pl.tol[m]);
    }
    st[g] = s;
}

// final parameters: candidate 0 = refined, 1 = input
__global__ void csp_final_params_kernel(const OptState *__restrict__ st, int n_groups, float *__restrict__ params) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    for (int m = 0; m < OPT_NP; ++m) {
        params[(long long)g * 12 + m] = st[g].x[m];
        params[(long long)g * 12 + 6 + m] = st[g].x0[m];
    }
}

// choose refined vs input (never a worse objective), update the entity tables
__global__ void csp_finish_kernel(const OptState *__restrict__ st, const CspGroup *__restrict__ groups, int n_groups, int kind,
                                  const float4 *__restrict__ obj, int *__restrict__ choice, cspb_particle *__restrict__ particles_out,
                                  cspb_tilt *__restrict__ tilts_out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const float fa = obj[2 * g].x, fb = obj[2 * g + 1].x;
    const int c = fa < fb ? 1 : 0;
    choice[g] = c;
    const float *x = c ? st[g].x0 : st[g].x;
    if (kind == 0) {
        cspb_particle p = particles_out[groups[g].ent];
        p.psi = x[0]; p.theta = x[1]; p.phi = x[2]; p.shift_x = x[3]; p.shift_y = x[4]; p.shift_z = x[5];
        p.score = 100.f * (c ? fb : fa);
        particles_out[groups[g].ent] = p;
    } else {
        cspb_tilt t = tilts_out[groups[g].ent];
        t.angle = x[0]; t.axis = x[1]; t.shift_x = x[2]; t.shift_y = x[3];
        tilts_out[groups[g].ent] = t;
    }
}

// rows of every member: re-composed pose at the chosen parameters, its score and statistics
__global__ void csp_write_rows_kernel(const CspEntry *__restrict__ list, int n_entries, CspTables T, int kind,
                                      const OptState *__restrict__ st, const int *__restrict__ choice,
                                      const float4 *__restrict__ sc, int n_samples, cspb_row *__restrict__ rows_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_entries) return;
    const CspEntry en = list[j];
    const int c = choice[en.group];
    const float *x = c ? st[en.group].x0 : st[en.group].x;
    float pose[6];
    csp_member_pose(T, kind, en.row, x, pose);
    cspb_row r = rows_out[en.row];
    r.psi = wrap360(pose[0]); r.theta = pose[1]; r.phi = wrap360(pose[2]);
    r.x_shift = pose[3]; r.y_shift = pose[4];
    r.defocus_1 += pose[5]; r.defocus_2 += pose[5];
    const float4 v = sc[(long long)j * 2 + c];
    r.score = 100.f * cc_of(v);
    score_stats(v, n_samples, &r.sigma, &r.logp);
    rows_out[en.row] = r;
}

// csp mode -2: cut boxes out of the tilt images with bin x bin real-space averaging
__global__ void csp_extract_kernel(const float *__restrict__ images, int nx, int ny, int n_tilt, const cspb_row *__restrict__ rows,
                                   const float *__restrict__ img_mean, int box_in, int bin, float *__restrict__ out) {
    const int bo = box_in / bin;
    const cspb_row r = rows[blockIdx.x];
    const int t = r.imind;
    float *o = out + (long long)blockIdx.x * bo * bo;
    const bool ok = t >= 0 && t < n_tilt;
    const float *img = images + (long long)(ok ? t : 0) * nx * ny;
    const int x0 = (int)floorf(r.original_x) - box_in / 2, y0 = (int)floorf(r.original_y) - box_in / 2;
    // a box with no pixel inside the image is all zeros, a partial one is padded with the mean of its inside part
    // (extract/core.py:100-155); the inside mean is approximated by the image mean
    const bool any_inside = ok && x0 < nx && y0 < ny && x0 + box_in > 0 && y0 + box_in > 0;
    const float fill = any_inside ? img_mean[t] : 0.f;
    const float inv = 1.f / (float)(bin * bin);
    for (int idx = threadIdx.x; idx < bo * bo; idx += blockDim.x) {
        const int ox = idx % bo, oy = idx / bo;
        float s = 0.f;
        for (int dy = 0; dy < bin; ++dy)
            for (int dx = 0; dx < bin; ++dx) {
                const int x = x0 + ox * bin + dx, y = y0 + oy * bin + dy;
                s += (any_inside && x >= 0 && x < nx && y >= 0 && y < ny) ? __ldg(img + (long long)y * nx + x) : fill;
            }
        o[idx] = s * inv;
    }
}

__global__ void image_mean_kernel(const float *__restrict__ images, long long npix, float *__restrict__ mean) {
    __shared__ float red[64];
    const float *p = images + (long long)blockIdx.x * npix;
    float s = 0.f;
    for (long long k = threadIdx.x; k < npix; k += blockDim.x) s += p[k];
    s = block_sum(s, red);
    if (threadIdx.x == 0) mean[blockIdx.x] = s / (float)npix;
}

int make_plan(const cspb_ctx *ctx, const cspb_csp_cfg *c, CspPlan *pl) {
    const float hi = ctx->plan.r_hi;
    const float h_ang = 0.35f * 57.29578f / hi;
    const float h_shift = 0.07f * (float)ctx->rcfg.box / hi * ctx->rcfg.pixel_size;
    const float h_def = ctx->rcfg.defocus_step > 0.f ? ctx->rcfg.defocus_step : 50.f;
    memset(pl, 0, sizeof *pl);
    int freem[OPT_NP] = {0, 0, 0, 0, 0, 0};
    switch (c->mode) {
    case 1: case 2: case 5: pl->kind = 0; break;
    case 0: case 3: case 4: case 6: pl->kind = 1; break;
    default: return -1;
    }
    if (pl->kind == 0) {
        const float tol[OPT_NP] = {c->tol_particle_psi, c->tol_particle_theta, c->tol_particle_phi, c->tol_particle_shift,
                                   c->tol_particle_shift, c->tol_particle_shift};
        const float h[OPT_NP] = {h_ang, h_ang, h_ang, h_shift, h_shift, h_shift};
        const float g[OPT_NP] = {c->angle_step, c->angle_step, c->angle_step, c->shift_step, c->shift_step, c->shift_step};
        memcpy(pl->tol, tol, sizeof tol); memcpy(pl->h, h, sizeof h); memcpy(pl->gstep, g, sizeof g);
        const int ang = c->mode == 1 || c->mode == 5, sh = c->mode == 2 || c->mode == 5;
        for (int m = 0; m < 3; ++m) { freem[m] = ang; freem[3 + m] = sh; }
    } else {
        const float tol[OPT_NP] = {c->tol_tilt_angle, c->tol_tilt_axis, c->tol_tilt_shift, c->tol_tilt_shift, c->tol_defocus, 0.f};
        const float h[OPT_NP] = {h_ang, h_ang, h_shift, h_shift, h_def, 0.f};
        const float g[OPT_NP] = {c->angle_step, c->angle_step, c->shift_step, c->shift_step, 0.25f * c->tol_defocus, 0.f};
        memcpy(pl->tol, tol, sizeof tol); memcpy(pl->h, h, sizeof h); memcpy(pl->gstep, g, sizeof g);
        const int ang = c->mode == 0 || c->mode == 6, sh = c->mode == 3 || c->mode == 6;
        freem[0] = freem[1] = ang;
        freem[2] = freem[3] = sh;
        freem[4] = c->mode == 4;
    }
    for (int m = 0; m < OPT_NP; ++m) {
        if (freem[m]) pl->free_mask |= 1 << m;
        pl->counts[m] = 1;
        if (c->grid_search && freem[m] && pl->gstep[m] > 0.f && pl->tol[m] > 0.f)
            pl->counts[m] = 2 * (int)floorf(pl->tol[m] / pl->gstep[m]) + 1;
    }
    pl->grid_search = c->grid_search ? 1 : 0;
    pl->seed = c->seed;
    return 0;
}

}  // namespace

extern "C" int cspb_csp_cfg_default(cspb_csp_cfg *c) {
    if (!c) return CSPB_E_ARG;
    memset(c, 0, sizeof *c);
    // defaults of config/pyp_config.toml [tabs.csp]
    c->mode = 5;
    c->window_min = 0;
    c->window_max = 20;
    c->iterations = 5;
    c->random_evals = 0;
    c->grid_search = 0;
    c->angle_step = 20.f;
    c->shift_step = 6.f;
    c->tol_particle_psi = c->tol_particle_theta = c->tol_particle_phi = 30.f;
    c->tol_particle_shift = 20.f;
    c->tol_tilt_angle = 1.5f;
    c->tol_tilt_axis = 1.f;
    c->tol_tilt_shift = 100.f;
    c->tol_defocus = 750.f;
    c->seed = 0;
    c->min_projections = 0;
    return 0;
}

extern "C" int cspb_csp_compose(const cspb_particle *p, const cspb_particle *p0, const cspb_tilt *t, const cspb_tilt *t0,
                                const float *centre3, float pixel_size, float base_x, float base_y, float *out5) {
    if (!p || !p0 || !t || !t0 || !centre3 || !out5) return CSPB_E_ARG;
    csp_compose(*p, *p0, *t, *t0, centre3[0], centre3[1], centre3[2], pixel_size, base_x, base_y, out5);
    return 0;
}

#define CSP_UP(buf, vec)                                                                                         \
    do {                                                                                                         \
        RESERVE(ctx, buf, (vec).size() * sizeof((vec)[0]) + 16);                                                 \
        if (!(vec).empty())                                                                                      \
            CU_TRY(ctx, cudaMemcpyAsync((buf).p, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice, ctx->stream)); \
    } while (0)

extern "C" int cspb_csp_run(cspb_ctx *ctx, cspb_row *rows, int n_rows, cspb_particle *particles, int n_particles,
                            cspb_tilt *tilts, int n_tilts, const cspb_csp_cfg *cfg, int first, int last, int64_t *n_evals_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows || !particles || !tilts || !cfg || n_rows < 0 || n_particles < 0 || n_tilts < 0) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (n_rows != ctx->n_images) return cspb_fail(ctx, CSPB_E_ARG, "rows (%d) != loaded images (%d)", n_rows, ctx->n_images);
    CspPlan pl;
    if (make_plan(ctx, cfg, &pl)) return cspb_fail(ctx, CSPB_E_ARG, "unknown csp mode %d", cfg->mode);
    if (n_evals_out) *n_evals_out = 0;
    if (n_rows == 0) return 0;

    // ---- host: row -> table indices, entity selection, member lists (window members first)
    std::unordered_map<int, int> pmap;
    std::unordered_map<long long, int> tmap;
    for (int a = n_particles - 1; a >= 0; --a) pmap[particles[a].pind] = a;  // first occurrence wins, as in the oracle
    for (int a = n_tilts - 1; a >= 0; --a) tmap[((long long)tilts[a].tind << 32) ^ (uint32_t)tilts[a].rind] = a;
    std::vector<int> row_part(n_rows), row_tilt(n_rows);
    for (int k = 0; k < n_rows; ++k) {
        auto ip = pmap.find(rows[k].pind);
        auto it = tmap.find(((long long)rows[k].tind << 32) ^ (uint32_t)rows[k].rind);
        if (ip == pmap.end() || it == tmap.end())
            return cspb_fail(ctx, CSPB_E_ARG, "row %d: PIND %d / (TIND %d, RIND %d) missing from the extended tables", k,
                             rows[k].pind, rows[k].tind, rows[k].rind);
        row_part[k] = ip->second;
        row_tilt[k] = it->second;
    }
    // float accumulation in table order, as the oracle does
    float cxf = 0.f, cyf = 0.f, czf = 0.f;
    for (int a = 0; a < n_particles; ++a) { cxf += particles[a].x_position_3d; cyf += particles[a].y_position_3d; czf += particles[a].z_position_3d; }
    if (n_particles > 0) { cxf /= (float)n_particles; cyf /= (float)n_particles; czf /= (float)n_particles; }
    const int n_ent = pl.kind == 0 ? n_particles : n_tilts;
    std::vector<std::vector<int>> members(n_ent);
    for (int k = 0; k < n_rows; ++k) members[pl.kind == 0 ? row_part[k] : row_tilt[k]].push_back(k);
    std::vector<CspGroup> groups;
    std::vector<CspEntry> list_search, list_all;
    for (int e = 0; e < n_ent; ++e) {
        const int id = pl.kind == 0 ? particles[e].pind : tilts[e].tind;
        if (id < first || (last >= 0 && id > last) || members[e].empty()) continue;
        CspGroup g;
        g.ent = e; g.id = id; g.off_search = (int)list_search.size(); g.off_all = (int)list_all.size();
        g.n_all = (int)members[e].size(); g.n_window = 0; g.n_search = 0; g.pad_ = 0;
        const int gi = (int)groups.size();
        for (int pass = 0; pass < 2; ++pass)
            for (int k : members[e]) {
                const int t = rows[k].tind;
                const bool in_w = !(t < cfg->window_min || (cfg->window_max != -1 && t > cfg->window_max));
                if (in_w == (pass == 0)) {
                    list_all.push_back({k, gi});
                    if (pass == 0) g.n_window++;
                }
            }
        if (g.n_window > 0 && g.n_window >= cfg->min_projections) {
            g.n_search = g.n_window;
            for (int k = 0; k < g.n_window; ++k) list_search.push_back(list_all[g.off_all + k]);
        }
        groups.push_back(g);
    }
    const int G = (int)groups.size();
    if (G == 0) return 0;
    const int n_search = (int)list_search.size(), n_all = (int)list_all.size();

    // ---- device tables
    cspb_row *d_rows;
    CtfCoef *d_ctf;
    int rc = upload_rows(ctx, rows, n_rows, &d_rows, &d_ctf);
    if (rc) return rc;
    DevBuf *const cb = ctx->csp_buf;  // persistent: no allocation in the steady state
    DevBuf &b_rows_out = cb[0], &b_part = cb[1], &b_part_out = cb[2], &b_tilt = cb[3], &b_tilt_out = cb[4], &b_rp = cb[5], &b_rt = cb[6],
           &b_groups = cb[7], &b_ls = cb[8], &b_la = cb[9], &b_best = cb[10], &b_choice = cb[11], &b_params = cb[12], &b_obj = cb[13],
           &b_dummy = cb[14];
    RESERVE(ctx, b_rows_out, (size_t)n_rows * sizeof(cspb_row));
    CU_TRY(ctx, cudaMemcpyAsync(b_rows_out.p, d_rows, (size_t)n_rows * sizeof(cspb_row), cudaMemcpyDeviceToDevice, ctx->stream));
    std::vector<cspb_particle> vp(particles, particles + n_particles);
    std::vector<cspb_tilt> vt(tilts, tilts + n_tilts);
    CSP_UP(b_part, vp); CSP_UP(b_part_out, vp); CSP_UP(b_tilt, vt); CSP_UP(b_tilt_out, vt);
    CSP_UP(b_rp, row_part); CSP_UP(b_rt, row_tilt); CSP_UP(b_groups, groups); CSP_UP(b_ls, list_search); CSP_UP(b_la, list_all);
    CspTables T;
    T.rows = d_rows; T.particles = b_part.as<cspb_particle>(); T.tilts = b_tilt.as<cspb_tilt>();
    T.row_part = b_rp.as<int>(); T.row_tilt = b_rt.as<int>(); T.cx = cxf; T.cy = cyf; T.cz = czf;
    const CspGroup *d_groups = b_groups.as<CspGroup>();

    int n_free = 0;
    for (int m = 0; m < OPT_NP; ++m) n_free += (pl.free_mask >> m) & 1;
    const int NE = 1 + 2 * n_free, PB = 4;
    // tilt shifts only move the projections in plane: their stencil evaluations share the centre's
    // rotation -> one gather (measured +9 % in mode 6).  Particle shifts have the same property, but
    // there the quad reuse of the plain units already removes the loads and the extra unit classes
    // cost more than they save (measured -10 % in mode 5), so they stay plain.
    const int shift_mask = pl.kind == 1 ? (pl.free_mask & 0x0C) : 0;
    int nS = 0;
    for (int m = 0; m < OPT_NP; ++m) nS += 2 * ((shift_mask >> m) & 1);
    const int iters = n_free > 0 && cfg->iterations > 0 ? cfg->iterations : 0;
    const int late = iters / 2 + 1;
    long long n_cand = 0;
    if (cfg->grid_search) { n_cand = 1; for (int m = 0; m < OPT_NP; ++m) n_cand *= pl.counts[m]; }
    else n_cand = cfg->random_evals > 0 ? cfg->random_evals : 0;
    // candidates per exhaustive chunk: bound the eval list to ~8 M poses
    int CH = 64;
    while (CH > 4 && (long long)n_search * CH > (8ll << 20)) CH /= 2;
    int max_nc = NE > CH ? NE : CH;
    if (max_nc < OPT_NL) max_nc = OPT_NL;
    const long long max_list = n_search > n_all ? n_search : n_all;
    const long long max_evals = (long long)n_search * max_nc > 2ll * n_all ? (long long)n_search * max_nc : 2ll * n_all;
    (void)max_list;
    RESERVE(ctx, ctx->d_opt, (size_t)G * sizeof(OptState));
    RESERVE(ctx, ctx->d_evals, (size_t)max_evals * 6 * sizeof(float));
    RESERVE(ctx, ctx->d_units, (size_t)(max_evals / 1 + 4) * sizeof(ScoreUnit));
    RESERVE(ctx, ctx->d_out, (size_t)max_evals * sizeof(float4));
    RESERVE(ctx, b_params, (size_t)G * (max_nc + OPT_NL + 2) * 6 * sizeof(float));
    RESERVE(ctx, b_obj, (size_t)G * (max_nc + OPT_NL + 2) * sizeof(float4));
    RESERVE(ctx, b_dummy, (size_t)G * ((NE + PB - 1) / PB + 3) * sizeof(ScoreUnit));
    RESERVE(ctx, b_best, (size_t)G * sizeof(CspBest));
    RESERVE(ctx, b_choice, (size_t)G * sizeof(int));
    OptState *st = ctx->d_opt.as<OptState>();
    float *poses = ctx->d_evals.as<float>();
    ScoreUnit *units = ctx->d_units.as<ScoreUnit>();
    float4 *out = ctx->d_out.as<float4>();
    float *params = b_params.as<float>();
    float4 *obj = b_obj.as<float4>();
    ScoreUnit *dummy = b_dummy.as<ScoreUnit>();
    const bool ddef = cfg->mode == 4;
    const int gg = ceil_div(G, 128);
    int64_t evals = 0;

    // evaluate `nc` candidates (params[g][c]) of every group over a list, objective -> obj[g][c]
    auto evaluate = [&](const CspEntry *list, int n_entries, int nc, int use_all, int n_shared = 0) -> int {
        if (n_entries > 0) {
            const long long tot = (long long)n_entries * nc;
            csp_expand_kernel<<<ceil_div(tot, 256), 256, 0, ctx->stream>>>(list, n_entries, nc, n_shared, PB, T, pl.kind, params, poses, units);
            KERNEL_CHECK(ctx);
            int r = launch_score_classes(ctx, units, n_entries, nc - n_shared, n_shared, PB, poses, d_ctf, out, ddef, false);
            if (r) return r;
            evals += tot;
        }
        csp_reduce_kernel<<<ceil_div((long long)G * nc, 128), 128, 0, ctx->stream>>>(d_groups, G, nc, use_all, out, obj);
        KERNEL_CHECK(ctx);
        return 0;
    };

    csp_init_kernel<<<gg, 128, 0, ctx->stream>>>(d_groups, G, T, pl, st);
    KERNEL_CHECK(ctx);
    if (n_search > 0) {
        // exhaustive stage
        for (long long k0 = 0; k0 < n_cand; k0 += CH) {
            const int nc = (int)(n_cand - k0 < CH ? n_cand - k0 : CH);
            csp_candidates_kernel<<<ceil_div((long long)G * nc, 128), 128, 0, ctx->stream>>>(st, d_groups, G, pl, k0, nc, params);
            KERNEL_CHECK(ctx);
            rc = evaluate(b_ls.as<CspEntry>(), n_search, nc, 0);
            if (rc) return rc;
            csp_best_kernel<<<gg, 128, 0, ctx->stream>>>(obj, G, nc, k0, k0 == 0, b_best.as<CspBest>());
            KERNEL_CHECK(ctx);
        }
        if (n_cand > 0) {
            csp_take_best_kernel<<<gg, 128, 0, ctx->stream>>>(st, d_groups, G, pl, b_best.as<CspBest>());
            KERNEL_CHECK(ctx);
        }
        // local stage: refine3d's stencil / Newton / line-search optimiser in entity space
        float *params_ls = params + (size_t)G * NE * 6;
        for (int it = 0; it < iters; ++it) {
            opt_stencil_kernel<<<gg, 128, 0, ctx->stream>>>(st, G, 1, pl.free_mask, shift_mask, NE, PB, params, dummy);
            KERNEL_CHECK(ctx);
            rc = evaluate(b_ls.as<CspEntry>(), n_search, NE, 0, nS);
            if (rc) return rc;
            opt_step_kernel<<<gg, 128, 0, ctx->stream>>>(st, G, 1, pl.free_mask, NE, obj, params_ls, dummy, OptPrior{});
            KERNEL_CHECK(ctx);
            // the line-search candidates are evaluated from the head of the params buffer
            CU_TRY(ctx, cudaMemcpyAsync(params, params_ls, (size_t)G * OPT_NL * 6 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
            rc = evaluate(b_ls.as<CspEntry>(), n_search, OPT_NL, 0);
            if (rc) return rc;
            opt_select_kernel<<<gg, 128, 0, ctx->stream>>>(st, G, obj, it + 1 >= late ? 0.6f : 1.f, OptPrior{});
            KERNEL_CHECK(ctx);
            csp_clamp_kernel<<<gg, 128, 0, ctx->stream>>>(st, G, pl);
            KERNEL_CHECK(ctx);
        }
    }
    // final: refined vs input parameters over all members
    csp_final_params_kernel<<<gg, 128, 0, ctx->stream>>>(st, G, params);
    KERNEL_CHECK(ctx);
    rc = evaluate(b_la.as<CspEntry>(), n_all, 2, 1);
    if (rc) return rc;
    csp_finish_kernel<<<gg, 128, 0, ctx->stream>>>(st, d_groups, G, pl.kind, obj, b_choice.as<int>(), b_part_out.as<cspb_particle>(),
                                                   b_tilt_out.as<cspb_tilt>());
    KERNEL_CHECK(ctx);
    csp_write_rows_kernel<<<ceil_div(n_all, 128), 128, 0, ctx->stream>>>(b_la.as<CspEntry>(), n_all, T, pl.kind, st, b_choice.as<int>(),
                                                                         out, ctx->plan.n_band, b_rows_out.as<cspb_row>());
    KERNEL_CHECK(ctx);
    CU_TRY(ctx, cudaMemcpyAsync(rows, b_rows_out.p, (size_t)n_rows * sizeof(cspb_row), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_particles > 0)
        CU_TRY(ctx, cudaMemcpyAsync(particles, b_part_out.p, (size_t)n_particles * sizeof(cspb_particle), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_tilts > 0)
        CU_TRY(ctx, cudaMemcpyAsync(tilts, b_tilt_out.p, (size_t)n_tilts * sizeof(cspb_tilt), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_evals_out) *n_evals_out = evals;
    return 0;
}

extern "C" int cspb_csp_extract(cspb_ctx *ctx, const float *images, int nx, int ny, int n_tilt, const cspb_row *rows,
                                int n_rows, int box_in, int bin, float *stack_out, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !images || !rows || !stack_out || nx <= 0 || ny <= 0 || n_tilt <= 0 || n_rows < 0) return CSPB_E_ARG;
    if (box_in <= 0 || bin <= 0 || box_in % bin) return cspb_fail(ctx, CSPB_E_ARG, "box %d must be a multiple of bin %d", box_in, bin);
    if (n_rows == 0) return 0;
    const int bo = box_in / bin;
    const size_t img_bytes = (size_t)nx * ny * n_tilt * sizeof(float), out_bytes = (size_t)n_rows * bo * bo * sizeof(float);
    DevBuf b_img, b_out, b_rows, b_mean;
    const float *d_img = images;
    float *d_out = stack_out;
    if (loc == CSPB_HOST) {
        RESERVE(ctx, b_img, img_bytes);
        RESERVE(ctx, b_out, out_bytes);
        CU_TRY(ctx, cudaMemcpyAsync(b_img.p, images, img_bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_img = b_img.as<float>();
        d_out = b_out.as<float>();
    }
    RESERVE(ctx, b_rows, (size_t)n_rows * sizeof(cspb_row));
    RESERVE(ctx, b_mean, (size_t)n_tilt * sizeof(float));
    CU_TRY(ctx, cudaMemcpyAsync(b_rows.p, rows, (size_t)n_rows * sizeof(cspb_row), cudaMemcpyHostToDevice, ctx->stream));
    image_mean_kernel<<<n_tilt, 1024, 0, ctx->stream>>>(d_img, (long long)nx * ny, b_mean.as<float>());
    KERNEL_CHECK(ctx);
    csp_extract_kernel<<<n_rows, 256, 0, ctx->stream>>>(d_img, nx, ny, n_tilt, b_rows.as<cspb_row>(), b_mean.as<float>(), box_in, bin, d_out);
    KERNEL_CHECK(ctx);
    if (loc == CSPB_HOST) CU_TRY(ctx, cudaMemcpyAsync(stack_out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}


// ================================================================== SPA particle extraction
// Box cutting of src/pyp/extract/core.py:360-511 (extract_particles_non_mpi, one micrograph): the box of particle k starts
// at floor(coord / coordinate_binning - floor(box / 2)) on both axes; a box that leaves the micrograph is padded with the
// mean of its inside part, a box entirely outside is zero.  The reference's clipping is kept to the letter, including
// its treatment of the upper edge: a box whose end reaches nx (maxX >= nx) loses its last line even when it fits exactly
// (core.py:471-483).  Empty boxes (constant, or < 1 % of the pixels different from the median — approximated by "constant")
// are replaced by unit white noise; the reference draws it from an unseeded generator (image.py:461-471), here it comes from
// a counter-based hash of (particle, pixel) so that runs repeat.  Normalisation (image.py:320-417) is NOT done here: the
// loader fuses it into the first FFT pass (fft.cu), so extraction -> normalisation -> FFT never writes a normalised stack.
namespace {
__device__ __forceinline__ float hash_normal(unsigned a, unsigned b) {
    unsigned x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA6Bu;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    unsigned y = x * 0x9E3779B1u + 0x6A09E667u;
    y ^= y >> 16; y *= 0x7FEB352Du; y ^= y >> 15; y *= 0x846CA68Bu; y ^= y >> 16;
    const float u1 = ((float)(x >> 8) + 1.f) * (1.f / 16777217.f), u2 = (float)(y >> 8) * (1.f / 16777216.f);
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

__global__ void __launch_bounds__(256) spa_extract_kernel(const float *__restrict__ img, int nx, int ny, const float *__restrict__ xy, int box,
                                                          float cbin, float *__restrict__ out) {
    __shared__ float red[34];
    __shared__ float s_lo, s_hi;
    const int k = blockIdx.x;
    float *o = out + (long long)k * box * box;
    // reference names: X runs over the rows of the array (coordinate y), Y over its columns (coordinate x)
    int minX = (int)floorf(xy[2 * k + 1] / cbin - floorf((float)box / 2.f)), minY = (int)floorf(xy[2 * k] / cbin - floorf((float)box / 2.f));
    int maxX = minX + box, maxY = minY + box;
    int minx = 0, miny = 0, maxx = box, maxy = box;
    if (minX < 0) { minx = -minX; minX = 0; } else if (maxX >= ny) { maxx = box - (maxX - ny + 1); maxX = ny - 1; }
    if (minY < 0) { miny = -minY; minY = 0; } else if (maxY >= nx) { maxy = box - (maxY - nx + 1); maxY = nx - 1; }
    const int h = maxX - minX, w = maxY - minY;  // inside part as the reference slices it
    const bool any = h > 0 && w > 0 && minX < ny && minY < nx;
    float sum = 0.f, lo = INFINITY, hi = -INFINITY;
    if (any)
        for (int t = threadIdx.x; t < h * w; t += blockDim.x) {
            const float v = __ldg(img + (long long)(minX + t / w) * nx + minY + t % w);
            sum += v; lo = fminf(lo, v); hi = fmaxf(hi, v);
        }
    sum = block_sum(sum, red);
    for (int q = 16; q > 0; q >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, q)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, q)); }
    __shared__ float w_lo[8], w_hi[8];
    if ((threadIdx.x & 31) == 0) { w_lo[threadIdx.x >> 5] = lo; w_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = w_lo[0], b = w_hi[0];
        for (int q = 1; q < 8; ++q) { a = fminf(a, w_lo[q]); b = fmaxf(b, w_hi[q]); }
        s_lo = a; s_hi = b;
    }
    __syncthreads();
    const float fill = any ? sum / (float)(h * w) : 0.f;
    // an empty frame (image.py:461-471: constant, or hardly any pixel off the median — taken here as "constant inside part")
    const bool empty = any && s_lo == s_hi;
    for (int t = threadIdx.x; t < box * box; t += blockDim.x) {
        const int r = t / box, c = t % box;
        float v = fill;
        if (any && r >= minx && r < maxx && c >= miny && c < maxy) v = __ldg(img + (long long)(minX + r - minx) * nx + minY + c - miny);
        if (empty) v = hash_normal((unsigned)k, (unsigned)t);
        o[t] = v;
    }
}
}  // namespace

extern "C" int cspb_spa_extract(cspb_ctx *ctx, const float *image, int nx, int ny, const float *coords_xy, int n, int box,
                                float coordinate_binning, float *stack_out, int image_loc, int out_loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !image || !coords_xy || !stack_out || nx <= 0 || ny <= 0 || n < 0 || box <= 0 || !(coordinate_binning > 0.f)) return CSPB_E_ARG;
    if (n == 0) return 0;
    const size_t img_bytes = (size_t)nx * ny * sizeof(float), out_bytes = (size_t)n * box * box * sizeof(float);
    DevBuf b_img, b_out, b_xy;
    const float *d_img = image;
    float *d_out = stack_out;
    if (image_loc == CSPB_HOST) {
        RESERVE(ctx, b_img, img_bytes);
        CU_TRY(ctx, cudaMemcpyAsync(b_img.p, image, img_bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_img = b_img.as<float>();
    }
    if (out_loc == CSPB_HOST) {
        RESERVE(ctx, b_out, out_bytes);
        d_out = b_out.as<float>();
    }
    RESERVE(ctx, b_xy, (size_t)n * 2 * sizeof(float));
    CU_TRY(ctx, cudaMemcpyAsync(b_xy.p, coords_xy, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    spa_extract_kernel<<<n, 256, 0, ctx->stream>>>(d_img, nx, ny, b_xy.as<float>(), box, coordinate_binning, d_out);
    KERNEL_CHECK(ctx);
    if (out_loc == CSPB_HOST) CU_TRY(ctx, cudaMemcpyAsync(stack_out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
