/*
 * cspb200.h — C-ABI of libcspb200.so, the B200 (sm_100a) drop-in engine behind nextPYP's
 * closed CPU binaries  external/cistem2/{refine3d,reconstruct3d,local_merge3d,merge3d}  and
 * external/CSP/csp.
 *
 * The reference has NO in-process API for this path: pyp spawns those executables and talks
 * to them through argv / a stdin heredoc / files.  The entry points below are
 * therefore what a cisTEM-style executable (or pyp itself, through ctypes) binds INSTEAD of the
 * numerics that live inside the binaries.  Each one cites the reference interface it replaces
 * (paths relative to /root/reference/).
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success or a negative CSPB_E_*;
 *     cspb_last_error(ctx) gives the text.  No exceptions cross the boundary.
 *   - caller allocates every buffer.  `loc` arguments say where a buffer lives:
 *     CSPB_HOST (pageable or pinned host memory) or CSPB_DEVICE (device pointer on ctx's GPU).
 *   - parameter rows are the packed 128-byte little-endian rows of a `.cistem` projection table
 *     (src/pyp/inout/metadata/cistem_star_file.py:596-628, layout in SURVEY.md Appendix B);
 *     `cspb_row` below is that exact layout, so a file can be mmapped and handed over unchanged.
 *   - image stacks are MRC mode-2 payloads: float32, x fastest, n*n per projection
 *     (src/pyp/inout/image/mrc.py:113-156,537-559).
 *   - a context is bound to one GPU; one process per GPU; a context is not thread-safe,
 *     different contexts are independent.
 *   - there is no CPU fallback: every call fails with CSPB_E_CUDA when no sm_100 device is usable.
 */
#ifndef CSPB200_H
#define CSPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSPB_ABI_VERSION 1

enum { CSPB_HOST = 0, CSPB_DEVICE = 1 };

enum {
    CSPB_OK = 0,
    CSPB_E_ARG = -1,    /* invalid argument / state              */
    CSPB_E_CUDA = -2,   /* CUDA runtime error / no usable device */
    CSPB_E_NOMEM = -3,  /* host or device allocation failed      */
    CSPB_E_STATE = -4,  /* call order violated (e.g. no reference set) */
};

/* One projection row of a .cistem table — cistem_star_file.py:596-628 (column order),
 * :127-185 (types).  Units: angles deg, shifts Angstrom, defocus Angstrom, pixel Angstrom,
 * voltage kV, Cs mm (src/pyp/inout/metadata/core.py:2891-2923). */
typedef struct cspb_row {
    uint32_t position_in_stack; /* 1-based */
    float psi, theta, phi;
    float x_shift, y_shift;
    float defocus_1, defocus_2, defocus_angle, phase_shift;
    int32_t image_is_active;    /* pyp re-uses it as film index (cistem_star_file.py:1516) */
    float occupancy, logp, sigma, score;
    float pixel_size, voltage_kv, cs_mm, amplitude_contrast;
    float beam_tilt_x, beam_tilt_y, image_shift_x, image_shift_y;
    float original_x, original_y;
    int32_t imind, pind, tind, rind, find;
    float fshift_x, fshift_y;
} cspb_row; /* sizeof == 128 */

typedef struct cspb_ctx cspb_ctx;

/* ------------------------------------------------------------------ lifecycle */
int cspb_abi_version(void);
/* Visible CUDA devices (0 when there is none): the front-ends map particle ranges to GPUs with it. */
int cspb_device_count(void);
/* Bind a context to CUDA device `device`. */
int cspb_create(int device, cspb_ctx **out);
int cspb_destroy(cspb_ctx *ctx);
const char *cspb_last_error(const cspb_ctx *ctx);
/* Block until all work queued on the context's stream has finished. */
int cspb_sync(cspb_ctx *ctx);
/* Raw cudaStream_t of the context (for CUDA-event timing and NCCL plumbing on the same stream). */
int cspb_stream(cspb_ctx *ctx, void **stream_out);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t cspb_launch_count(const cspb_ctx *ctx);

/* Live kernel timing for the roofline (bench.py): when enabled, every launch of the scoring
 * kernel (kind 0) and of the insertion kernel (kind 1) is bracketed by CUDA events on the
 * context stream.  cspb_profile_get synchronises and returns the summed device time, the
 * number of launches and the units processed (kind 0: (image,pose) evaluations; kind 1:
 * projection x symmetry-operator insertions) since the last enable. */
enum { CSPB_PROF_SCORE = 0, CSPB_PROF_INSERT = 1, CSPB_PROF_KINDS = 2 };
int cspb_profile_enable(cspb_ctx *ctx, int on);
int cspb_profile_get(cspb_ctx *ctx, int kind, double *total_ms, int64_t *launches, int64_t *units);
/* Gather-load census of the scoring kernel (roofline bookkeeping): while on, every scorer launch is
 * followed by a counting launch over the same units that adds up the 32-byte reference loads the scorer
 * issues and the 8-byte reads of the packed image spectra (one per lane and band slot visited).
 * cspb_profile_get_loads returns both counts and the evaluations they cover since the census was switched on.
 * Run it on an untimed step; it has no reference counterpart. */
int cspb_profile_count_loads(cspb_ctx *ctx, int on);
int cspb_profile_get_loads(cspb_ctx *ctx, int64_t *quad_loads, int64_t *slot_reads, int64_t *evals);

/* ------------------------------------------------------------------ refine3d
 * Replaces the numerics of external/cistem2/refine3d as driven by
 * src/pyp/refine/frealign/frealign.py:3918-3994 (prompt order in SURVEY.md Appendix A.2).
 * Field comments give the prompt number that carries the value. */
typedef struct cspb_refine_cfg {
    int32_t box;               /* n: image edge in pixels (from the stack header)              */
    int32_t pad;               /* prompt 35 padding factor refine_iblow (1 or 2)               */
    float pixel_size;          /* prompt 15, Angstrom                                          */
    float mask_radius;         /* prompt 18 outer mask radius (particle_rad), Angstrom         */
    float low_res_limit;       /* prompt 19 refine_rlref, Angstrom                             */
    float high_res_limit;      /* prompt 20 get_rhref(), Angstrom                              */
    float signed_cc_limit;     /* prompt 21 (30.0 or refine_fboostlim), Angstrom; 0 = always signed */
    float search_mask_radius;  /* prompt 23 global-search mask radius, Angstrom                */
    float search_high_res;     /* prompt 24 high-res limit for the global search, Angstrom     */
    float angular_step;        /* prompt 25 refine_dang, degrees                               */
    int32_t best_matches;      /* prompt 26 number of global-search matches to refine (20)     */
    float search_range_x;      /* prompt 27, Angstrom                                          */
    float search_range_y;      /* prompt 28, Angstrom                                          */
    float defocus_range;       /* prompt 33, Angstrom                                          */
    float defocus_step;        /* prompt 34, Angstrom                                          */
    int32_t global_search;     /* prompt 36                                                    */
    int32_t local_refine;      /* prompt 37                                                    */
    int32_t refine_psi, refine_theta, refine_phi, refine_x, refine_y; /* prompts 38-42         */
    int32_t refine_defocus;    /* prompt 45                                                    */
    int32_t apply_mask;        /* soft circular particle mask of radius mask_radius (default 1);   */
                               /* prompt 44 switches the FOCUS mask: cspb_refine_set_focus_mask    */
    int32_t normalize;         /* prompt 46 normalise particles                                */
    int32_t invert_contrast;   /* prompt 47                                                    */
    int32_t whiten;            /* 1 = whiten with the stack's noise power curve (cisTEM default) */
    int32_t symmetry_order;    /* number of symmetry matrices handed to cspb_set_symmetry (1 = C1) */
    int32_t local_iterations;  /* batched local-optimiser iterations (ours; default 8)         */
    int32_t use_priors;        /* prompt 7: restrain X/Y shifts to the data set's distribution     */
    float prior_mean_x, prior_mean_y; /* Angstrom: row 0 of <name>_stat.cistem (particle_cspt.py:1009-1016) */
    float prior_var_x, prior_var_y;   /* Angstrom^2: row 1; <= 0 leaves that shift unrestrained     */
    int32_t optimizer;         /* local optimiser (ours), also for the hits of a global search: 0 = analytic gradient +
                                  Gauss-Newton step, coarse to fine (SEMANTICS.md §7c), 1 = central-difference stencil (§7;
                                  always used for the defocus refinement) */
    int32_t reserved[1];
} cspb_refine_cfg;

/* Fill a config with the defaults pyp passes for a plain local refinement. */
int cspb_refine_cfg_default(cspb_refine_cfg *cfg, int box, float pixel_size);

/* Configure the scorer; builds the polar-patch band plan for (box, low_res, high_res). */
int cspb_refine_configure(cspb_ctx *ctx, const cspb_refine_cfg *cfg);
/* Drop the loaded images, the whitening curve and the ring weights; the configuration and the transformed
 * reference stay.  For a resident engine serving several front-end invocations against one reference. */
int cspb_refine_reset_images(cspb_ctx *ctx);

/* Optional per-ring SSNR weights (prompt 5/6, statistics_rNN.txt part_SSNR column mapped to
 * rings of the box); n_rings = box/2+1.  NULL resets to all-ones. */
int cspb_refine_set_ring_weights(cspb_ctx *ctx, const float *w, int n_rings);

/* 2-D focus mask (prompts 29-32 with prompt 44 = yes; class_focusmask "x,y,z,radius",
 * frealign.py:3845-3848,3883-3885): sphere centre in Angstrom from the corner of the map and radius in
 * Angstrom.  With radius > 0 cspb_refine_run* writes into LOGP the log-likelihood of the real-space
 * residual inside the projected sphere (oracle/SEMANTICS.md 6b) instead of the whole-band value.
 * radius <= 0 switches it off (default). */
int cspb_refine_set_focus_mask(cspb_ctx *ctx, float x, float y, float z, float radius);

/* refine_ctf answer 23 (estimate beam tilt, frealign.py:3995-4041): sum over the loaded images of
 * G * conj(CTF * slice) on the scoring band at the pose of each row; out_complex = box * (box/2+1)
 * complex64 (re, im interleaved), zero outside the band.  The host fits the coma phase to it
 * (pyp_b200/beamtilt.py; oracle/SEMANTICS.md 12). */
int cspb_refine_phase_sum(cspb_ctx *ctx, const cspb_row *rows, int n_rows, float *out_complex);

/* Upload the reference map (prompt 4, n^3 float32, x fastest) and build the padded, centred,
 * band-cropped Fourier half-volume in HBM. */
int cspb_set_reference(cspb_ctx *ctx, const float *vol, int n, int loc);

/* Symmetry matrices (row-major 3x3 each) for prompt 11 (refine) / 9 (reconstruct). */
int cspb_set_symmetry(cspb_ctx *ctx, const float *mats, int n_mats);

/* Load `n_images` projections (stack slices first..last) and preprocess them on the GPU:
 * normalise -> FFT -> whiten -> mask -> band-pack.  `images` is n_images*box*box float32.
 * May be called repeatedly with append=1 to stream a stack in chunks. */
int cspb_refine_load_images(cspb_ctx *ctx, const float *images, int n_images, int loc, int append);

/* One forward transform per projection for refine3d AND reconstruct3d (the reference's two programs each read and
 * transform the stack: frealign.py:3918-3994, 1780-1824).  on != 0: a device-resident, non-appending
 * cspb_refine_load_images also keeps the plain forward transforms and their normalisation (264 KB per 256-px image, skipped
 * silently when the device lacks the room); a later cspb_recon_insert[_weighted] whose `images` pointer lies inside that
 * same device buffer rescales them to its own normalisation (the transform is linear: factor scl_r/scl_f, DC term
 * scl_r (off_f - off_r) n^2) instead of transforming again.  THE CALLER PROMISES the pixels have not changed in between.
 * Any other load, or on = 0, drops what was kept.  The streamed pipelines below do this internally. */
int cspb_refine_keep_spectra(cspb_ctx *ctx, int on);
int cspb_refine_num_images(const cspb_ctx *ctx);

/* Score every loaded image at the pose in its row (one evaluation per image).
 * rows[k] belongs to loaded image k.  scores_out: n floats (SCORE column, 100*CC). */
int cspb_refine_score(cspb_ctx *ctx, const cspb_row *rows, int n, float *scores_out);

/* Generic evaluation list: eval e scores image image_index[e] at
 * poses[e] = {psi, theta, phi (deg), shift_x, shift_y (Angstrom), defocus delta (Angstrom)}.
 * Evaluations should be sorted by image for best locality.  This is the objective function
 * that csp (external/CSP/csp, src/pyp/system/local_run.py:364-376) sums over tilts. */
int cspb_refine_score_poses(cspb_ctx *ctx, const cspb_row *rows, int n_rows,
                            const int32_t *image_index, const float *poses6, int n_evals,
                            float *scores_out);
/* Value and analytic derivatives of the score at the poses of `rows` (the evaluation the default local optimiser is
 * built on, oracle/SEMANTICS.md §7c): out28 = 28 floats per row {num, X, A, B, d num / d (psi, theta, phi per degree,
 * x, y per Angstrom), d B / d angles, the 15 entries of the upper triangle of J^T J, 0}.  ring_cut > 0 restricts the
 * sums to the rings <= ring_cut (coarse-to-fine stages); 0 = the whole band. */
int cspb_refine_score_grad(cspb_ctx *ctx, const cspb_row *rows, int n, int ring_cut, float *out28);
/* Matching projections (refine3d answers 8 and 43, `<name>_match.mrc_<first>_<last>`, frealign.py:3929-3931): for every
 * row the CTF-multiplied projection of the reference at the row's pose, displaced by the row's shift into the frame of
 * the particle image, band limited at the scoring limit; out_images = n_rows * box * box floats (host). */
int cspb_refine_matching_projections(cspb_ctx *ctx, const cspb_row *rows, int n_rows, float *out_images);

/* Orientation grid of the global search (prompt 36/25): n_orient x {psi, theta, phi} degrees.
 * The host builds it from the angular step and the symmetry symbol (pyp_b200/search_grid.py),
 * the same grid is handed to the CPU oracle. */
int cspb_refine_set_search_grid(cspb_ctx *ctx, const float *angles3, int n_orient);

/* Full refine3d pass over the loaded images: optional global search, then batched local
 * refinement of the masked parameters.  rows are updated in place (PSI, THETA, PHI, X_SHIFT,
 * Y_SHIFT, [DEFOCUS], LOGP, SIGMA, SCORE, score change is returned in changes_out if non-NULL).
 * n_evals_out receives the number of (image,pose) objective evaluations performed. */
int cspb_refine_run(cspb_ctx *ctx, cspb_row *rows, int n, cspb_row *changes_out,
                    int64_t *n_evals_out);

/* Asynchronous resident variant used by the benchmark: rows live in device memory, nothing is
 * copied, the work is only enqueued on the context stream. */
int cspb_refine_run_device(cspb_ctx *ctx, cspb_row *rows_dev, int n, int64_t *n_evals_out);

/* Noise power curve used for whitening (box/2+1 floats); estimated by load_images when the
 * context has none yet, or set explicitly (multi-GPU: all-reduce the partial sums first). */
int cspb_refine_get_noise_curve(cspb_ctx *ctx, float *curve_out, int n_rings);
int cspb_refine_set_noise_curve(cspb_ctx *ctx, const float *curve, int n_rings);

/* ------------------------------------------------------------------ csp
 * Replaces the numerics of external/CSP/csp (closed LFS binary) as driven by
 * src/pyp/system/local_run.py:306-467 (argv: par, extended par, mode, first, last, flag, images,
 * stack) and src/pyp/align/core.py:883-1248.  The extended tables are the two blocks of
 * `<par>_extended.cistem` (src/pyp/inout/metadata/cistem_star_file.py:247-248); the structs below
 * are those packed rows, so the file blocks can be handed over unchanged. */
typedef struct cspb_particle {
    int32_t pind;
    float shift_x, shift_y, shift_z;   /* PSHIFT_X/Y/Z, Angstrom */
    float psi, theta, phi;             /* PPSI, PTHETA, PPHI, degrees */
    float x_position_3d, y_position_3d, z_position_3d; /* ORIGINAL_*_POSITION_3D, pixels */
    float score, occ;                  /* PSCORE, POCC */
} cspb_particle; /* sizeof == 48 */

typedef struct cspb_tilt {
    int32_t tind, rind;
    float shift_x, shift_y;            /* TSHIFT_X/Y, Angstrom */
    float angle, axis;                 /* TILTANG, TILTAXIS, degrees */
} cspb_tilt; /* sizeof == 24 */

/* csp_* keys of .pyp_config.toml (config/pyp_config.toml [tabs.csp], lines 6241-6640) */
typedef struct cspb_csp_cfg {
    int32_t mode;             /* argv mode: 0 tilt angle+axis, 1 particle angles, 2 particle shifts,
                                 3 tilt shifts, 4 tilt defocus offset, 5 particle angles+shifts,
                                 6 tilt angle+axis+shifts (align/core.py:1015-1023, local_run.py:332-335) */
    int32_t window_min;       /* csp_UseImagesForRefinementMin (TIND window of the objective)       */
    int32_t window_max;       /* csp_UseImagesForRefinementMax, -1 = open (cistem_star_file.py:965) */
    int32_t iterations;       /* csp_OptimizerMaxIter                                               */
    int32_t random_evals;     /* csp_NumberOfRandomIterations: exhaustive-stage candidates          */
    int32_t grid_search;      /* csp_GridSearch: lattice instead of random candidates               */
    float angle_step;         /* csp_AngleStep, degrees                                             */
    float shift_step;         /* csp_ShiftStep, Angstrom                                            */
    float tol_particle_psi, tol_particle_theta, tol_particle_phi; /* csp_ToleranceParticles{Psi,Theta,Phi} */
    float tol_particle_shift; /* csp_ToleranceParticlesShifts                                       */
    float tol_tilt_angle;     /* csp_ToleranceMicrographTiltAngles                                  */
    float tol_tilt_axis;      /* csp_ToleranceMicrographTiltAxisAngles                              */
    float tol_tilt_shift;     /* csp_ToleranceMicrographShifts                                      */
    float tol_defocus;        /* csp_ToleranceMicrographDefocus1                                    */
    uint32_t seed;            /* random-search seed (counter-based generator, same on GPU and oracle) */
    int32_t min_projections;  /* csp_RefineProjectionCutoff                                         */
    int32_t reserved[6];
} cspb_csp_cfg;

int cspb_csp_cfg_default(cspb_csp_cfg *cfg);

/* Constrained refinement of the entities first..last (PIND for the particle modes 1/2/5, TIND for
 * the tilt modes 0/3/4/6; last < 0 = no upper bound).  rows[k] belongs to loaded image k
 * (cspb_refine_load_images); every row's PIND and (TIND, RIND) must exist in the tables.  The
 * objective of an entity is the mean score of its projections inside the exposure window
 * (update_particle_score, cistem_star_file.py:936-986).  On return the refined entities, the
 * poses / scores of all their rows and PSCORE are updated in place. */
int cspb_csp_run(cspb_ctx *ctx, cspb_row *rows, int n_rows, cspb_particle *particles, int n_particles,
                 cspb_tilt *tilts, int n_tilts, const cspb_csp_cfg *cfg, int first, int last,
                 int64_t *n_evals_out);

/* Pose of one projection from its particle and tilt (src/pyp/analysis/geometry/core.py:1081-1217):
 * out5 = psi, theta, phi (deg), x, y shift (Angstrom).  Host-side helper, no GPU needed. */
int cspb_csp_compose(const cspb_particle *p, const cspb_particle *p0, const cspb_tilt *t,
                     const cspb_tilt *t0, const float *centre3, float pixel_size, float base_x,
                     float base_y, float *out5);

/* csp mode -2 (align/core.py:958-964, local_run.py:449): cut n_rows boxes of edge `box_in` out of a
 * tilt series (n_tilt images of nx*ny, float32), centred on (ORIGINAL_X_POSITION, ORIGINAL_Y_POSITION)
 * of each row on image IMIND, bin them by bin x bin real-space averaging (the reference's own
 * "real" method, src/pyp/extract/core.py:180-203) and write box_in/bin-pixel particles to
 * `stack_out` (n_rows * (box_in/bin)^2 floats).  Pixels outside the image are filled with the
 * image mean; a box entirely outside is zero (extract/core.py:100-155). */
int cspb_csp_extract(cspb_ctx *ctx, const float *images, int nx, int ny, int n_tilt, const cspb_row *rows,
                     int n_rows, int box_in, int bin, float *stack_out, int loc);

/* SPA box cutting (src/pyp/extract/core.py:360-511, one micrograph): n boxes of edge `box` from image (ny rows of nx
 * float32), box k starting at floor(coords_xy[2k + {0,1}] / coordinate_binning - floor(box / 2)) along x / y; a box that
 * leaves the micrograph is padded with the mean of its inside part (the reference's clipping rule to the letter), a box
 * entirely outside is zero, a constant box is replaced by reproducible unit white noise (image.py:461-471).  The particles
 * are NOT normalised here — cspb_refine_load_images / cspb_recon_insert fuse normalize_image (image.py:320-417) into their
 * first FFT pass — so with out_loc = CSPB_DEVICE the boxes go from the micrograph to the scorer without a stack on the host. */
int cspb_spa_extract(cspb_ctx *ctx, const float *image, int nx, int ny, const float *coords_xy, int n, int box,
                     float coordinate_binning, float *stack_out, int image_loc, int out_loc);

/* ------------------------------------------------------------------ reconstruct3d
 * Replaces external/cistem2/reconstruct3d as driven by frealign.py:1780-1824
 * (SURVEY.md Appendix A.3). */
typedef struct cspb_recon_cfg {
    int32_t box;               /* n                                                     */
    int32_t pad;               /* prompt 25 padding (pyp always passes 1)               */
    float pixel_size;          /* prompt 12                                             */
    float mask_radius;         /* prompt 15 outer radius rad_rec, Angstrom              */
    float resolution_limit;    /* prompt 16 res_rec, Angstrom (2*pixel = Nyquist)       */
    float score_bfactor;       /* prompt 18 refine_bsc, score -> B-factor constant      */
    int32_t score_weighting;   /* prompt 19                                             */
    float score_threshold;     /* prompt 23                                             */
    int32_t normalize;         /* prompt 26                                             */
    int32_t invert_contrast;   /* prompt 28                                             */
    int32_t per_particle_split;/* prompt 32: halves by PIND parity instead of stack parity */
    float average_score;       /* mean SCORE of the contributing rows (for score weighting) */
    int32_t reserved[8];
} cspb_recon_cfg;

int cspb_recon_cfg_default(cspb_recon_cfg *cfg, int box, float pixel_size);

/* Allocate and zero the two Fourier half-accumulators {re, im, ctf^2 weight, pad} per voxel. */
int cspb_recon_begin(cspb_ctx *ctx, const cspb_recon_cfg *cfg);

/* Insert n_images projections (OCC > 0 rows only contribute) with CTF multiplication and
 * CTF^2 weight accumulation, for every symmetry matrix set with cspb_set_symmetry. */
int cspb_recon_insert(cspb_ctx *ctx, const float *images, const cspb_row *rows, int n_images,
                      int loc);
/* The same with the data-driven dose weighting of prompt 22 (frealign.py:1731-1753; oracle/SEMANTICS.md §10): weight_cut
 * holds two floats per projection {weight, cut radius in Fourier pixels}; the projection's samples are weighted by
 * `weight` and, beyond the cut radius, by a raised-cosine edge of width 0.05 * box (cut radius <= 0: no low-pass).
 * The pairs come from pyp_b200.tables.dose_weight_pairs (host). weight_cut lives where `loc` says; NULL = cspb_recon_insert. */
int cspb_recon_insert_weighted(cspb_ctx *ctx, const float *images, const cspb_row *rows, int n_images, int loc,
                               const float *weight_cut);

/* Accumulator geometry: voxels = (np/2+1)*np*np, 4 floats per voxel. */
int cspb_recon_dims(const cspb_ctx *ctx, int *np_out, int64_t *floats_per_half_out);
/* Device pointer of half h (0/1) — hand it to ncclReduce / torch.distributed. */
int cspb_recon_device_ptr(cspb_ctx *ctx, int half, void **ptr_out);
/* Dump / add-dump: the file-mode collective of local_merge3d (frealign.py:1878-1888). */
int cspb_recon_get_dump(cspb_ctx *ctx, int half, float *out, int loc);
int cspb_recon_add_dump(cspb_ctx *ctx, int half, const float *in, int loc);

/* merge3d (frealign.py:2075-2093): symmetrise, per-shell statistics, optimal-filter
 * normalisation, inverse FFT, gridding correction, outer mask.
 * half1/half2/map: n^3 float32 each (may be NULL); stats: n_shells*7 floats, rows
 * {shell, resolution A, ring radius (1/A * ... ), FSC, Part_FSC, Part_SSNR^.5, Rec_SSNR^.5}. */
int cspb_recon_finalize(cspb_ctx *ctx, float molecular_mass_kda, float outer_radius_a,
                        float *half1, float *half2, float *map, float *stats, int n_shells,
                        int loc);
int cspb_recon_end(cspb_ctx *ctx);

/* Batch sizes the streamed pipeline below uses for a stack of n_images of `box` pixels when the scorer's wave is wave_units
 * (cspb_wave_units): returns their number and writes up to max_sizes of them.  Host-only (no device needed). */
int cspb_pipeline_batches(int n_images, int box, int wave_units, int *sizes_out, int max_sizes);

/* ------------------------------------------------------------------ streamed host pipeline
 * refine3d and / or reconstruct3d over a HOST stack (pinned memory for full overlap) with ONE upload
 * per projection: chunk k+1 is copied on a second stream while chunk k is preprocessed, refined
 * (rows updated in place, as cspb_refine_run) and inserted (as cspb_recon_insert, with the refined
 * rows) on the context stream.  Chunks are those of cspb_refine_load_images, so the results are
 * identical to load_images + refine_run + recon_insert.  Needs cspb_refine_configure +
 * cspb_set_reference for CSPB_DO_REFINE and cspb_recon_begin for CSPB_DO_INSERT. */
enum { CSPB_DO_REFINE = 1, CSPB_DO_INSERT = 2 };
int cspb_refine_reconstruct(cspb_ctx *ctx, const float *images_host, cspb_row *rows_host, int n_images,
                            int flags, int64_t *n_evals_out);

/* ------------------------------------------------------------------ between the stages
 * Score shaping and class occupancies on device-resident tables, so that refine -> select -> reconstruct never leaves
 * the GPU.  Replaces the host step pyp runs between its two binaries:
 *   src/pyp/analysis/scores.py:300-761 shape_phase_residuals (as called at :766-825 and particle_cspt.py:747-755):
 *   OCCUPANCY := 0 for projections below the score threshold of `cutoff` (SPA: that quantile of the scores; tilt series
 *   — any |tilt| > 0 — the same quantile of the per-particle mean scores over |tilt| <= 12, applied to the means over
 *   |tilt| < 10), outside the min/max score, defocus, azimuth, frame and tilt windows; POSITION_IN_STACK := 1..n;
 *   src/pyp/analysis/occupancies.py:173-208 occupancy_extended.
 * The host restatement pyp_b200/select.py is pinned bit for bit against the reference's outputs (tests/golden/shape_*). */
typedef struct cspb_select_cfg {
    float cutoff;              /* reconstruct_cutoff as a fraction in (0, 1]; 1 keeps everything                       */
    float mindef, maxdef;      /* Angstrom (scores.py:646-650)                                                         */
    int32_t firstframe, lastframe; /* TIND window, lastframe < 0 = off (:672-680)                                      */
    float mintilt, maxtilt;    /* degrees (:682-690)                                                                   */
    float minazh, maxazh;      /* THETA mod 180 window, [0, 180] = off (:652-670)                                      */
    float minscore, maxscore;  /* fractions of the score range when <= 1, absolute scores otherwise (:519-530)          */
    int32_t renumber;          /* rewrite POSITION_IN_STACK = 1..n like `_used.cistem` files (:757-759)                */
    float threshold_override;  /* not NaN: use this threshold (the automatic two-Gaussian cutoff, reconstruct_cutoff = 0,
                                  is fitted on the host: pyp_b200/select.py optimal_threshold)                         */
    int32_t reserved[4];
} cspb_select_cfg;
int cspb_select_cfg_default(cspb_select_cfg *cfg);
/* rows: n projection rows (device or host per `loc`), tilt_angle: per-row tilt angle in degrees or NULL (single
 * particle: all zero).  threshold_out (may be NULL) receives the threshold used (NaN if none). */
int cspb_select_scores(cspb_ctx *ctx, cspb_row *rows, int n, const float *tilt_angle, const cspb_select_cfg *cfg, int loc,
                       double *threshold_out);
/* LogP -> occupancy over K classes: logp, sigma (K x n, class major), K previous mean occupancies ->
 * occ_out (K x n, percent) and sigma_out (n). */
int cspb_class_occupancies(cspb_ctx *ctx, const float *logp, const float *sigma, const double *class_average_occ, int n_classes,
                           int n, float *occ_out, float *sigma_out, int loc);
/* Data-driven dose weights: mean SCORE per scan-order index (TIND) over the projections with
 * OCCUPANCY > 0, -1 where there is none — the content of pyp's global_weight.txt (src/pyp/inout/metadata/core.py:3039-3075)
 * that reconstruct3d's prompt 22 reads, computed on the rows the scorer left on the device.  weights_out: n_idx doubles;
 * *n_used_out = 1 + the largest index with projections (the length of the file). */
int cspb_global_weights(cspb_ctx *ctx, const cspb_row *rows, int n, int loc, double *weights_out, int n_idx, int *n_used_out);
/* refine3d -> score shaping -> reconstruct3d in one call over a HOST stack: every projection is uploaded once into a
 * resident device buffer and refined as its batch arrives (as cspb_refine_reconstruct); when all rows are refined the
 * selection runs on the device table and the whole stack is inserted with the shaped occupancies.  Needs
 * cspb_refine_configure + cspb_set_reference + cspb_recon_begin.  rows_host returns the refined, shaped rows. */
int cspb_refine_select_reconstruct(cspb_ctx *ctx, const float *images_host, cspb_row *rows_host, int n_images,
                                   const cspb_select_cfg *cfg, int64_t *n_evals_out, double *threshold_out);

/* ------------------------------------------------------------------ building blocks
 * Exposed for parity tests against the oracle and cuFFT. */
/* Batched 2-D real-to-complex FFT, unnormalised, output n*(n/2+1) complex per image. */
int cspb_fft2_r2c(cspb_ctx *ctx, const float *in, float *out_complex, int n, int batch, int loc);
int cspb_fft2_c2r(cspb_ctx *ctx, const float *in_complex, float *out, int n, int batch, int loc);
/* Same through cuFFT (comparison baseline only; never on the product path). */
int cspb_cufft2_r2c(cspb_ctx *ctx, const float *in, float *out_complex, int n, int batch, int loc);
/* CTF values on the half-plane grid for one row: out n*(n/2+1) floats. */
int cspb_ctf_image(cspb_ctx *ctx, const cspb_row *row, int n, float *out);
/* Central slice of the current reference at one pose: out n*(n/2+1) complex (zero outside band). */
int cspb_project(cspb_ctx *ctx, float psi, float theta, float phi, float *out_complex);
/* Gather microbenchmark (the roofline denominator of the scoring kernel next to the HBM peak,
 * SURVEY.md §8d): GB/s of 32-lane gathers of 32-byte items at random positions of a window of
 * `window_bytes` — per_cta = 1: every CTA has its own window (L1-resident for <= 64 KB),
 * per_cta = 0: one shared window (L2-resident for tens of MB, HBM for GBs). */
int cspb_gather_peak(cspb_ctx *ctx, size_t window_bytes, int per_cta, float *gbs_out);
/* Units (image x <=4 poses, one warp each) of the scoring kernel resident at once on this GPU = one
 * full wave; stacks whose particle count is a multiple of it lose nothing to wave quantisation. */
int cspb_wave_units(cspb_ctx *ctx);
/* Band plan introspection: number of lattice samples inside the band / padded slots. */
int cspb_band_counts(const cspb_ctx *ctx, int *n_band, int *n_slots);

#ifdef __cplusplus
}
#endif
#endif /* CSPB200_H */
