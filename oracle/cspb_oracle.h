/*
 * cspb_oracle.h — CPU restatement of the CSP refine3d / reconstruct3d hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may import, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY STATUS: "parity unpinned" for the numerics.  The reference's arithmetic for this path
 * lives in closed git-LFS binaries (external/cistem2/{refine3d,reconstruct3d,merge3d},
 * external/CSP/csp — 133-byte pointer stubs in /root/reference) built from the
 * un-vendored pyp fork of cisTEM 2.0.0-alpha (banner quoted at
 * src/pyp/inout/image/mrc.py:646-651).  This file restates the published cisTEM/FREALIGN
 * algorithm (Grant, Rohou & Grigorieff, eLife 2018; Grigorieff, Methods Enzymol. 2016) with
 * every choice written down in oracle/SEMANTICS.md.  What IS pinned against the reference
 * tree: the `.cistem` row layout (cistem_star_file.py:596-628), the MRC container
 * (mrc.py:113-156), the Euler decode (geometry/core.py:222-247) and the prompt orders
 * (frealign.py:3918-3994,1780-1824) — see tests/golden/.
 */
#ifndef CSPB_ORACLE_H
#define CSPB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the 128-byte .cistem projection row — cistem_star_file.py:596-628 */
typedef struct orc_row {
    uint32_t position_in_stack;
    float psi, theta, phi, x_shift, y_shift;
    float defocus_1, defocus_2, defocus_angle, phase_shift;
    int32_t image_is_active;
    float occupancy, logp, sigma, score;
    float pixel_size, voltage_kv, cs_mm, amplitude_contrast;
    float beam_tilt_x, beam_tilt_y, image_shift_x, image_shift_y;
    float original_x, original_y;
    int32_t imind, pind, tind, rind, find;
    float fshift_x, fshift_y;
} orc_row;

/* refine3d answers that matter to the numerics — frealign.py:3918-3994 (same fields as
 * cspb_refine_cfg in include/cspb200.h) */
typedef struct orc_refine_cfg {
    int32_t box, pad;
    float pixel_size, mask_radius, low_res_limit, high_res_limit, signed_cc_limit;
    float defocus_step;
    int32_t refine_psi, refine_theta, refine_phi, refine_x, refine_y, refine_defocus;
    int32_t apply_mask, normalize, invert_contrast, whiten, local_iterations;
    /* global search (prompts 24-28, 36) */
    float search_high_res, search_range_x, search_range_y;
    int32_t best_matches, global_search;
    /* shift restraint (prompt 7 "use priors"; SEMANTICS.md §7b): mean / variance of X_SHIFT, Y_SHIFT in
     * Angstrom (rows 0 / 1 of <name>_stat.cistem, particle_cspt.py:1009-1016); variance <= 0: unrestrained */
    int32_t use_priors;
    float prior_mean_x, prior_mean_y, prior_var_x, prior_var_y;
    /* 2-D focus mask (prompts 29-32 + 44, class_focusmask "x,y,z,radius" in Angstrom from the corner of
     * the map, frealign.py:3845-3848,3883-3885; SEMANTICS.md §6b): LOGP over the projected sphere */
    float focus_x, focus_y, focus_z, focus_radius;
    /* local optimiser: 0 = analytic gradient + Gauss-Newton step (SEMANTICS.md §7c; used when the defocus is not
     * refined), 1 = central-difference stencil (§7) */
    int32_t optimizer;
} orc_refine_cfg;

typedef struct orc_recon_cfg {
    int32_t box, pad;
    float pixel_size, mask_radius, resolution_limit, score_bfactor;
    int32_t score_weighting;
    float score_threshold;
    int32_t normalize, invert_contrast, per_particle_split;
    float average_score;
} orc_recon_cfg;

typedef struct orc_ref orc_ref;      /* padded centred Fourier reference */
typedef struct orc_recon orc_recon;  /* two half accumulators */

/* ---- FFT (double precision inside) */
void orc_fft2_r2c(const float *img, int n, float *out_c /* n*(n/2+1) complex */);
void orc_fft2_c2r(const float *in_c, int n, float *out /* unnormalised inverse */);

/* ---- Euler / CTF */
void orc_euler_matrix(float psi, float theta, float phi, float *r9);
void orc_ctf_image(const orc_row *row, int n, float *out /* n*(n/2+1) */);
float orc_band_limits(const orc_refine_cfg *cfg, float *r_lo, float *r_hi); /* returns r_hi; also n_band via orc_band_count */
int orc_band_count(const orc_refine_cfg *cfg);
/* score with analytic derivatives: dnum[5] (psi, theta, phi per degree; x, y per Angstrom), dB[3], jtj[15] (upper triangle) */
float orc_score_grad(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg,
                     float *out4, float *dnum, float *dB, float *jtj);
/* the same restricted to the rings <= ring_cut (ring_cut <= 0: the whole band) */
float orc_score_grad_cut(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg,
                         float *out4, float *dnum, float *dB, float *jtj, int ring_cut);

/* ---- reference */
orc_ref *orc_ref_create(const float *vol, int n, int pad);
void orc_ref_free(orc_ref *r);
void orc_project(const orc_ref *r, float psi, float theta, float phi, float r_hi, float *out_c);

/* ---- particle preprocessing: image -> prepared half spectrum n*(n/2+1) complex */
void orc_noise_curve(const float *imgs, int count, const orc_refine_cfg *cfg, float *curve /* n+1 */);
void orc_prepare_image(const float *img, const orc_refine_cfg *cfg, const float *noise_curve /* or NULL */,
                       const float *ring_weights /* or NULL */, float *spec_out);

/* ---- score: out4 = {numerator, signed cross, image power, projection power}; returns 100*CC */
float orc_score(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6,
                const orc_refine_cfg *cfg, float *out4);

/* ---- batched local refinement of n particles (specs: n prepared spectra).  OpenMP over
 * particles when built with -fopenmp.  Returns number of objective evaluations. */
/* refine_ctf beam-tilt input (frealign.py:3995-4041, answer 23): sum over the images of
 * G * conj(CTF * slice) on the scoring band at the pose of each row (G = shifted prepared spectrum of
 * orc_score); out_c = n * (n/2+1) complex, zero outside the band (SEMANTICS.md §12) */
/* particle normalisation of src/pyp/analysis/image.py:320-338,406-417 (radius in pixels) */
void orc_normalize(const float *img, int n, float radius_px, int normalize, int invert, float *out);
void orc_phase_sum(const orc_ref *r, const float *specs, const orc_row *rows, int n_img, const orc_refine_cfg *cfg, float *out_c);
/* centre (pixels, image coordinates) of the projected focus sphere for a pose */
void orc_focus_center(const orc_refine_cfg *cfg, const float *pose6, float *cx, float *cy);
/* LOGP of the real-space residual inside the projected focus sphere; o4 = band sums of the pose */
float orc_focus_logp(const orc_ref *r, const float *spec, const orc_row *row, const float *pose6, const orc_refine_cfg *cfg,
                     const float *o4);
long long orc_refine_local(const orc_ref *r, const float *specs, orc_row *rows, int n,
                           const orc_refine_cfg *cfg);

/* ---- global search over the orientation grid angles3 (n_orient x psi,theta,phi), then local
 * refinement of the best_matches hits.  Returns number of objective evaluations. */
long long orc_global_search(const orc_ref *r, const float *specs, orc_row *rows, int n,
                            const orc_refine_cfg *cfg, const float *angles3, int n_orient);

/* ---- constrained single-particle refinement (external/CSP/csp; argv contract at
 * src/pyp/system/local_run.py:364-376,392-404,451-463).  Extended tables as in
 * cistem_star_file.py:247-248.  Same fields as cspb_particle / cspb_tilt / cspb_csp_cfg. */
typedef struct orc_particle {
    int32_t pind;
    float shift_x, shift_y, shift_z, psi, theta, phi, x_position_3d, y_position_3d, z_position_3d, score, occ;
} orc_particle;
typedef struct orc_tilt {
    int32_t tind, rind;
    float shift_x, shift_y, angle, axis;
} orc_tilt;
typedef struct orc_csp_cfg {
    int32_t mode, window_min, window_max, iterations, random_evals, grid_search;
    float angle_step, shift_step;
    float tol_particle_psi, tol_particle_theta, tol_particle_phi, tol_particle_shift;
    float tol_tilt_angle, tol_tilt_axis, tol_tilt_shift, tol_defocus;
    uint32_t seed;
    int32_t min_projections;
    int32_t reserved[6];
} orc_csp_cfg;
/* pose of one projection from its particle and tilt parameters (geometry/core.py:1081-1217):
 * out5 = psi, theta, phi (deg), x, y (same unit as the inputs) for base shift (bx, by);
 * p0/t0 are the input (extraction-time) parameters the base shift belongs to. */
void orc_csp_compose(const orc_particle *p, const orc_particle *p0, const orc_tilt *t, const orc_tilt *t0,
                     const float *centre3, float pixel, float bx, float by, float *out5);
/* refine entities first..last (PIND for particle modes 1/2/5, TIND for tilt modes 0/3/4/6; last < 0 = open)
 * in place; rows[k] belongs to specs[k].  Returns the number of objective evaluations, < 0 on error. */
long long orc_csp_run(const orc_ref *r, const float *specs, orc_row *rows, int n_rows, orc_particle *particles,
                      int n_particles, orc_tilt *tilts, int n_tilts, const orc_refine_cfg *cfg,
                      const orc_csp_cfg *csp, int first, int last);

/* ---- reconstruction */
orc_recon *orc_recon_create(const orc_recon_cfg *cfg);
void orc_recon_free(orc_recon *rc);
void orc_recon_insert(orc_recon *rc, const float *imgs, const orc_row *rows, int count,
                      const float *sym /* n_sym*9 */, int n_sym);
/* the same with the per-row {weight, cut radius} pairs of the dose weighting (SEMANTICS.md §10); aux may be NULL */
void orc_recon_insert_weighted(orc_recon *rc, const float *imgs, const orc_row *rows, int count, const float *sym, int n_sym, const float *aux);
/* raw accumulators, same layout as the GPU dump: [z][y][x] float4 {re, im, w, 0}, centred y,z */
void orc_recon_get_dump(const orc_recon *rc, int half, float *out);
void orc_recon_finalize(orc_recon *rc, float molecular_mass_kda, float outer_radius_a, float *half1,
                        float *half2, float *map, float *stats /* (n/2+1)*7 */);

/* FSC curve between two real n^3 volumes (n/2+1 shells, nearest-integer shells) */
void orc_fsc(const float *a, const float *b, int n, float *fsc_out);

/* ---- optimised CPU leg (cspb_oracle_fast.c): the same algorithms, production-style; results equal the functions above */
long long orc_refine_local_fast(const orc_ref *r, const float *specs, orc_row *rows, int n_img, const orc_refine_cfg *cfg);
void orc_recon_insert_fast(orc_recon *rc, const float *imgs, const orc_row *rows, int count, const float *sym, int n_sym, int finish);
void orc_recon_finish_fast(orc_recon *rc);
void orc_recon_discard_fast(orc_recon *rc);
void orc_prepare_image_fast(const float *img, const orc_refine_cfg *cfg, const float *noise_curve, const float *ring_weights, float *spec);

#ifdef __cplusplus
}
#endif
#endif
