// internal.cuh — shared declarations of libcspb200 (not part of the C-ABI).
// Everything here is ours; the reference has no source for this path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/cspb200.h"

#define CSPB_PI_F 3.14159265358979323846f
#define CSPB_PI_D 3.14159265358979323846

// ---------------------------------------------------------------- device buffer (grow-only)
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    // returns false on allocation failure
    bool reserve(size_t want) {
        if (want <= bytes) return true;
        release();
        if (cudaMalloc(&p, want) != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            return false;
        }
        bytes = want;
        return true;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ---------------------------------------------------------------- band plan
// The half-plane lattice samples inside the scoring band, re-ordered into "polar patches":
// ring-bands of 4 rings; inside a band slot = base + a*4 + k holds the a-th sample (by angle) of
// the band's k-th ring, so a warp's 32 slots are an 8(angle) x 4(ring) patch and lane%4 is the
// ring offset.  Rings are padded to a common length with dummy slots; the 4 rings of a band are
// neighbours in sample COUNT (rings sorted by count, a local permutation of the radial order:
// lattice ring counts fluctuate by +-10 around pi*r), which halves the padding of radial bands.
#define CSPB_DUMMY_I 0x7FFF
struct BandDesc {
    int slot_start;  // multiple of 32
    int n_iter;      // slots/32
    int rings01;     // ring index of track 0 (low 16 bits) and track 1 (high 16 bits)
    int rings23;     // tracks 2 and 3
};
struct BandPlan {
    int n = 0;
    float r_lo = 0, r_hi = 0;
    int ring_min = 0, ring_max = 0;
    int n_band = 0;   // real samples
    int n_slots = 0;  // padded slots (multiple of 32)
    int n_bands = 0;
    std::vector<int32_t> slot_ij;  // i | (j << 16)
    std::vector<BandDesc> bands;
    DevBuf d_slot_ij, d_bands;
    DevBuf d_slot_of;   // inverse map of the band plan: (row jj, column i) of the half spectrum -> slot, -1 outside the band
    DevBuf d_dummy;     // the padding slots of the plan
    int n_dummy = 0;
};
bool build_band_plan(BandPlan &plan, int n, float r_lo, float r_hi, bool radial_order = false);

// ---------------------------------------------------------------- reference volume (scoring)
// Cropped, centred (y,z shifted by +rc), x-paired Fourier half-volume:
// ref4[(z*sy + y)*sx + x] = { V(x,y,z), V(x+1,y,z) } as float4.
// The scorer reads a second copy in (y,z) 2x2 quads, one 32-byte item per voxel, so that the 8
// trilinear neighbours of a sample are two LDG.256 at x and x+1 (half the L1 wavefronts of 4 x 16 B):
// ref8[(z*sy + y)*sx8 + x] = { V(x,y,z), V(x,y+1,z), V(x,y,z+1), V(x,y+1,z+1) }.
struct __align__(32) RefQuad {
    float2 v00, v10, v01, v11;
};
struct RefVolume {
    int n = 0, pad = 1, np = 0;
    int rc = 0;          // centre offset: y,z in [-rc, rc]
    int sx = 0, sy = 0;  // sx = rc+1, sy = sz = 2rc+1
    int sx8 = 0;         // x extent of the quad volume: rc+2
    DevBuf d_ref4;       // x-paired copy (global search, projections)
    DevBuf d_ref8;       // quad copy (scoring kernel)
    bool ready = false;
};

// per-image CTF coefficients (device): chi = s2*(a + b*(c2a*cos2ast + s2a*sin2ast)) + c4*s2*s2 + ph0
// with s2 = (i^2+j^2) in Fourier-pixel^2 units (pixel size folded in).
struct CtfCoef {
    float a, b, cos2ast, sin2ast, c4, ph0, dstep, pad_;  // dstep: d(a)/d(defocus Angstrom)
};

// one warp of the scoring kernel = one unit = (image, up to PB candidate poses)
struct ScoreUnit {
    int image;       // index into packed images / ctf coefficients
    int first_eval;  // index of the unit's first evaluation
    int count;       // number of poses (<= PB)
    int pad_;
};

struct ProfRec {
    cudaEvent_t a, b;
    int kind;
    int64_t units;
};

// ---------------------------------------------------------------- context
#define CSPB_PIPE_STAGES 3

struct cspb_ctx {
    bool prof_on = false;
    std::vector<ProfRec> prof;
    // gather-load census of the scorer (cspb_profile_count_loads): quad loads the launches issue
    bool count_loads = false;
    DevBuf d_load_count;
    int64_t census_evals = 0;

    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int sm_count = 148;
    int wave_units = 0;  // resident warps of the scoring kernel over the whole GPU (cspb_wave_units)

    // refine state
    bool refine_ready = false;
    cspb_refine_cfg rcfg{};
    BandPlan plan;
    RefVolume ref;
    DevBuf d_ring_w;        // per-ring weights (n/2+1 .. n)
    bool have_ring_w = false;
    DevBuf d_noise;         // whitening filter 1/sqrt(P) per ring (n+1 floats)
    std::vector<float> noise_curve;  // P(ring)
    bool have_noise = false;
    int n_images = 0;
    int img_capacity = 0;
    DevBuf d_packed;        // n_images * n_slots float2
    std::vector<float> sym; // 9*n_sym, the full group as given
    DevBuf d_sym;
    int n_sym = 1;
    // reconstruction-side decomposition G = H * R: H = operators that permute the voxel lattice
    // (signed permutation matrices), applied once to the accumulated volume; R = right-coset
    // representatives, applied per sample by the insertion kernel.
    std::vector<float> sym_lit;  // 9*n_lit
    std::vector<int> sym_lat_t;  // 9*n_lat, transposed (= inverse) integer matrices
    DevBuf d_sym_lit, d_sym_lat;
    int n_lit = 1, n_lat = 1;

    // 2-D focus mask (cspb_refine_set_focus_mask): sphere centre x, y, z and radius in Angstrom; radius <= 0 = off
    float focus[4] = {0.f, 0.f, 0.f, 0.f};
    DevBuf d_alpha;  // per image: signal scale X / B of the pose kept by the refinement

    // global-search orientation grid (psi, theta, phi) and hit buffer
    DevBuf d_grid, d_hits;
    int n_grid = 0;

    // scratch
    DevBuf d_stage, d_work0, d_work1, d_work2, d_stats, d_rows, d_evals, d_units, d_out, d_opt;
    DevBuf d_tw;  // twiddle tables
    std::vector<int> tw_n;      // sizes cached
    std::vector<size_t> tw_off; // offsets (in float2)
    size_t tw_used = 0;

    // streamed host pipeline (pipeline.cu): copy stream, events, two staging buffers, device rows — kept
    // between calls (allocating and freeing ~10 GB per call costs tens of milliseconds)
    cudaStream_t pipe_copy = nullptr;
    // three staging buffers: with two, the copy of batch k+2 waits for the kernels of batch k, and a large batch followed by
    // smaller ones leaves the copy engine idle (r02za timeline: copies done at 168-180 ms instead of 155)
    cudaEvent_t pipe_ready[CSPB_PIPE_STAGES] = {}, pipe_freed[CSPB_PIPE_STAGES] = {};
    DevBuf pipe_stage[CSPB_PIPE_STAGES], pipe_rows;
    DevBuf pipe_all;  // resident stack of cspb_refine_select_reconstruct
    DevBuf csp_buf[15];  // tables and work buffers of cspb_csp_run, kept between calls (cudaMalloc / cudaFree per call cost
                         // up to hundreds of ms of jitter per step: r02zp)
    // forward transforms kept for the insertion (cspb_refine_keep_spectra): half spectra of the images of the last
    // device-resident, non-appending cspb_refine_load_images, normalised as for refinement, plus that normalisation
    // (offset, scale per image); cspb_recon_insert of the same pixels rescales them instead of transforming again
    bool keep_on = false;
    DevBuf d_keep_spec, d_keep_stats, d_keep_scale;
    const float *keep_src = nullptr;
    int keep_count = 0, keep_box = 0, keep_total = 0;   // images kept so far / box / images of the span (stats layout)
    const float *keep_span_src = nullptr;               // pipeline.cu: the loads that follow are consecutive pieces of this
    int keep_span_count = 0;                            // resident stack (kept spectra accumulate instead of being replaced)
    float keep_radius = 0.f;          // normalisation of the kept spectra: radius in pixels, normalize, invert
    int keep_normalize = 0, keep_invert = 0;
    bool keep_recon_valid = false;    // d_keep_stats also holds the reconstruction's normalisation (same image pass), made for
    float keep_recon_radius = 0.f;    // this radius / contrast sign
    int keep_recon_invert = 0;

    // recon state
    bool recon_ready = false;
    cspb_recon_cfg ccfg{};
    int rnp = 0;
    DevBuf d_acc[2];
    DevBuf d_raw[2];      // deferred-symmetry accumulators (only when n_lat > 1)
    bool raw_dirty = false;
    DevBuf d_shell;  // per-shell sums, Wiener terms, statistics rows
    DevBuf d_aux;    // per-row {weight, cut radius} of the dose weighting (host callers)
    int64_t recon_inserted = 0;
};

// The current CUDA device is per host thread and other code in the process (torch, a second context on
// another GPU) may change it: every C-ABI entry that allocates or launches makes the context's device
// current for the duration of the call and restores the caller's afterwards.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const cspb_ctx *ctx) {
        if (!ctx) return;
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != ctx->device && cudaSetDevice(ctx->device) == cudaSuccess) switched = prev >= 0;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define CSPB_ENTER(ctx) DeviceGuard cspb_device_guard_(ctx)

int cspb_fail(cspb_ctx *ctx, int code, const char *fmt, ...);
extern "C" int cspb_wave_units(cspb_ctx *ctx);
// event bracket helpers (no-ops unless profiling is enabled)
void prof_begin(cspb_ctx *ctx, int kind, int64_t units);
void prof_end(cspb_ctx *ctx);
// fold pending deferred-symmetry inserts into the main accumulators (recon.cu)
int recon_flush_deferred(cspb_ctx *ctx);

#define CU_TRY(ctx, expr)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return cspb_fail(ctx, CSPB_E_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                             cudaGetErrorString(e_), __FILE__, __LINE__);                  \
    } while (0)

#define KERNEL_CHECK(ctx)                                                                  \
    do {                                                                                   \
        (ctx)->launches++;                                                                 \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess)                                                             \
            return cspb_fail(ctx, CSPB_E_CUDA, "kernel launch failed: %s (%s:%d)",         \
                             cudaGetErrorString(e_), __FILE__, __LINE__);                  \
    } while (0)

#define RESERVE(ctx, buf, bytes_)                                                          \
    do {                                                                                   \
        if (!(buf).reserve(bytes_))                                                        \
            return cspb_fail(ctx, CSPB_E_NOMEM, "device allocation of %zu bytes failed",   \
                             (size_t)(bytes_));                                            \
    } while (0)

// ---------------------------------------------------------------- scorer (refine.cu)
// enqueue the scoring kernel over `n_units` units; poses6 per eval = psi, theta, phi (deg), shift x, y
// (Angstrom), defocus delta (Angstrom); out per eval = {numerator, signed X, A, B}
// every unit of the range has exactly `count` poses (one template instance per count, branch free);
// mode 1 (shared) = rotation and CTF of the unit's first pose apply to all its poses (pure shift variations);
// mode 2 (same shift) = the shift of the unit's first pose applies to all its poses (pure rotation / defocus variations)
// ring_cut != INT_MAX (coarse-to-fine stages of the analytic optimiser): only rings <= ring_cut are scored; single-pose units only
int launch_score(cspb_ctx *ctx, const ScoreUnit *d_units, int n_units, int count, const float *d_poses6,
                 const CtfCoef *d_ctf, float4 *d_out, bool ddef, int64_t n_evals, int mode, int ring_cut = 0x7fffffff);
// units in class layout [A full][A tail][S full][S tail] (opt.cuh); nA plain + nS shared evals per group
// a_same_shift: every pose of an A unit carries the shift of the unit's first pose (refine3d stencils)
int launch_score_classes(cspb_ctx *ctx, const ScoreUnit *d_units, int n_groups, int nA, int nS, int PB, const float *d_poses6,
                         const CtfCoef *d_ctf, float4 *d_out, bool ddef, bool a_same_shift);
// upload rows next to their CTF coefficients (ctx->d_rows)
int upload_rows(cspb_ctx *ctx, const cspb_row *rows, int n, cspb_row **d_rows, CtfCoef **d_ctf);

// ---------------------------------------------------------------- FFT (fft.cu)
// twiddle table W_n^k = exp(-2 pi i k / n), k < n, cached per n on the device
int fft_get_twiddles(cspb_ctx *ctx, int n, const float2 **tw_out);
// batched 2-D R2C / C2R on device buffers; out pitch = n/2+1 complex per row
// scale/offset: optional per-image affine applied while loading rows: v = (x - off[b]) * scl[b]
// radial_filter (fast path only, see fft_has_fast_path): multiply the spectrum by filt[nearest ring]
int fft2_r2c_dev(cspb_ctx *ctx, const float *in, float2 *out, int n, int batch,
                 const float *offs, const float *scls, const float *radial_filter = nullptr);
// scale and (fast path only) a soft circular real-space mask are applied while the rows are written
int fft2_c2r_dev(cspb_ctx *ctx, float2 *inout_c, float *out, int n, int batch, float scale = 1.f,
                 float mask_radius = 0.f, float mask_width = 0.f);
bool fft_has_fast_path(int n);
// fused preprocessing of the fast path (fft.cu): normalise + r2c | whiten round trip | mask round trip | pack; see refine.cu
int fft2_whiten_mask_pack_dev(cspb_ctx *ctx, const float *in, float2 *spec, int n, int batch, const float *offs, const float *scls,
                              const float *radial_filter, float scale, float mask_radius, float mask_width, const int32_t *slot_of,
                              const float *ringw, const int32_t *dummy_list, int n_dummy, float2 *packed, int n_slots,
                              float2 *keep_forward = nullptr);
// 3-D R2C / C2R of an np^3 volume (in-place complex work buffer of (np/2+1)*np*np)
int fft3_r2c_dev(cspb_ctx *ctx, const float *in, float2 *out, int np);
int fft3_c2r_dev(cspb_ctx *ctx, float2 *inout_c, float *out, int np);

// ---------------------------------------------------------------- helpers
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// images per processing chunk (staging + half spectra within ~4.5 GB, at most 8192, even chunks).
// Large chunks matter: a scorer launch over 2048 particles fills less than half of the 148 SMs.
static inline int chunk_images(int n, int n_images) {
    const size_t per_img = (size_t)n * n * 4 + (size_t)n * (n / 2 + 1) * 8;
    long long chunk = (long long)(((size_t)9 << 29) / per_img);
    if (chunk < 1) chunk = 1;
    if (chunk > 8192) chunk = 8192;
    if (n_images > chunk) chunk = ceil_div(n_images, ceil_div(n_images, chunk));
    return (int)chunk;
}
