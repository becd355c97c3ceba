// select.cu — the step between refine3d and reconstruct3d on the device: score shaping (which projections enter the
// reconstruction) and LogP -> occupancy over classes.  SURVEY.md §8f rank 1.
//
// Restates on device-resident rows what pyp does on the host between its two binaries:
//   src/pyp/analysis/scores.py:300-761 `shape_phase_residuals` as called by `call_shape_phase_residuals` (:766-825) — one
//   angular and one defocus group, scores, a cutoff fraction in (0, 1] (the automatic cutoff of reconstruct_cutoff = 0 is
//   a two-Gaussian fit the caller does on the host, pyp_b200/select.py, and hands over as `threshold_override`);
//   src/pyp/analysis/occupancies.py:173-208 `occupancy_extended`.
// The host restatement pyp_b200/select.py is pinned bit for bit against the reference's outputs (tests/golden/shape_*);
// these kernels are compared with it on the same tables (tests/test_gpu_select.py).  Decisions are taken in float64
// like numpy does: scores are widened, per-particle means are float64 sums (exact for <= 2^29 terms of comparable
// magnitude, hence independent of the order of the atomics).
#include <cub/cub.cuh>
#include <math.h>
#include "internal.cuh"

namespace {

__global__ void sel_minmax_tomo_kernel(const cspb_row *__restrict__ rows, const float *__restrict__ tilt, int n, float *__restrict__ mm /*min,max*/,
                                       int *__restrict__ flags /*is_tomo, max pind*/) {
    float lo = INFINITY, hi = -INFINITY;
    int tomo = 0, pmax = -1;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const float s = rows[k].score;
        lo = fminf(lo, s);
        hi = fmaxf(hi, s);
        if (tilt && fabsf(tilt[k]) > 0.f) tomo = 1;
        pmax = max(pmax, rows[k].pind);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        tomo |= __shfl_xor_sync(0xffffffffu, tomo, o);
        pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // float atomics on the bit pattern: scores may be negative, so compare as floats through CAS
        int *ilo = reinterpret_cast<int *>(mm), *ihi = ilo + 1;
        int old = *ilo;
        while (lo < __int_as_float(old)) { const int prev = atomicCAS(ilo, old, __float_as_int(lo)); if (prev == old) break; old = prev; }
        old = *ihi;
        while (hi > __int_as_float(old)) { const int prev = atomicCAS(ihi, old, __float_as_int(hi)); if (prev == old) break; old = prev; }
        if (tomo) atomicOr(flags, 1);
        atomicMax(flags + 1, pmax);
    }
}

__global__ void sel_scores_kernel(const cspb_row *__restrict__ rows, int n, float *__restrict__ keys) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) keys[k] = rows[k].score;
}

// per-particle float64 sums / counts of the scores of the rows with |tilt| <= lim (inclusive) or < lim (strict)
__global__ void sel_particle_sums_kernel(const cspb_row *__restrict__ rows, const float *__restrict__ tilt, int n, float lim, int strict,
                                         double *__restrict__ sums, int *__restrict__ cnt) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float t = fabsf(tilt[k]);
    if (strict ? !(t < lim) : !(t <= lim)) return;
    const int p = rows[k].pind;
    if (p < 0) return;
    atomicAdd(sums + p, (double)rows[k].score);
    atomicAdd(cnt + p, 1);
}

// means of the particles that have rows, compacted in pind order (np.unique order) — one thread, tiny
__global__ void sel_compact_means_kernel(const double *__restrict__ sums, const int *__restrict__ cnt, int n_p, double *__restrict__ means,
                                         int *__restrict__ n_out) {
    if (blockIdx.x || threadIdx.x) return;
    int m = 0;
    for (int p = 0; p < n_p; ++p)
        if (cnt[p] > 0) means[m++] = sums[p] / (double)cnt[p];
    *n_out = m;
}

struct SelParams {
    double threshold, lo, hi;
    int tomo_rule;          // threshold applied to the near-tilt particle means instead of the row scores
    int cutoff_is_one;
    float mindef, maxdef, mintilt, maxtilt, minazh, maxazh;
    int firstframe, lastframe, renumber;
};

__global__ void sel_apply_kernel(cspb_row *__restrict__ rows, const float *__restrict__ tilt, int n, const SelParams P,
                                 const double *__restrict__ sums10, const int *__restrict__ cnt10) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    cspb_row r = rows[k];
    const double s = (double)r.score;
    bool drop = false;
    if (P.tomo_rule) {
        if (!P.cutoff_is_one && r.pind >= 0 && cnt10[r.pind] > 0) {
            const double mean = sums10[r.pind] / (double)cnt10[r.pind];
            if (!(mean >= P.threshold)) drop = true;
        }
        if (s < P.lo || s > P.hi) drop = true;
    } else if (s < P.threshold || s < P.lo || s > P.hi)
        drop = true;
    const double d1 = (double)r.defocus_1;
    if (d1 < (double)P.mindef || d1 > (double)P.maxdef) drop = true;
    if (P.maxazh < 180.f || P.minazh > 0.f) {
        double az = fmod((double)r.theta, 180.0);
        if (az < 0.0) az += 180.0;
        if (az < (double)P.minazh || az > (double)P.maxazh) drop = true;
    }
    if (P.lastframe > -1 && (r.tind < P.firstframe || r.tind > P.lastframe)) drop = true;
    const double ta = tilt ? (double)tilt[k] : 0.0;
    if (ta < (double)P.mintilt || ta > (double)P.maxtilt) drop = true;
    if (drop) r.occupancy = 0.f;
    if (P.renumber) r.position_in_stack = (uint32_t)(k + 1);
    rows[k] = r;
}

__global__ void occ_kernel(const float *__restrict__ logp, const float *__restrict__ sigma, const double *__restrict__ avg, int K, int n,
                           float *__restrict__ occ, float *__restrict__ sig_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double mx = -INFINITY;
    for (int c = 0; c < K; ++c) mx = fmax(mx, (double)logp[(size_t)c * n + k]);
    double total = 0.0;
    for (int c = 0; c < K; ++c) {
        const double d = mx - (double)logp[(size_t)c * n + k];
        if (d < 10.0) total += exp(-d) * avg[c];
    }
    double sg = 0.0;
    for (int c = 0; c < K; ++c) {
        const double d = mx - (double)logp[(size_t)c * n + k];
        const double o = d < 10.0 ? exp(-d) * avg[c] * 100.0 / total : 0.0;
        occ[(size_t)c * n + k] = (float)o;
        sg += (double)sigma[(size_t)c * n + k] * o / 100.0;
    }
    sig_out[k] = (float)sg;
}

// per scan-order index: float64 sum of the scores and count of the projections with OCCUPANCY > 0
__global__ void weights_kernel(const cspb_row *__restrict__ rows, int n, int n_idx, double *__restrict__ sums, int *__restrict__ cnt) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const cspb_row r = rows[k];
    if (!(r.occupancy > 0.f) || r.tind < 0 || r.tind >= n_idx) return;
    atomicAdd(sums + r.tind, (double)r.score);
    atomicAdd(cnt + r.tind, 1);
}

}  // namespace

// Data-driven dose weights: mean SCORE per scan-order index (TIND) over the projections with
// OCCUPANCY > 0, -1 where none — what pyp's compute_global_weights writes to global_weight.txt
// (src/pyp/inout/metadata/core.py:3039-3075) for reconstruct3d's prompt 22, here on the device-resident rows the scorer
// just produced.  weights_out: n_idx doubles; *n_used_out = 1 + the largest index that has projections (the file's length).
extern "C" int cspb_global_weights(cspb_ctx *ctx, const cspb_row *rows, int n, int loc, double *weights_out, int n_idx, int *n_used_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows || !weights_out || n < 0 || n_idx <= 0) return CSPB_E_ARG;
    const cspb_row *d_rows = rows;
    DevBuf b_rows, b_sums, b_cnt;
    if (loc == CSPB_HOST && n > 0) {
        RESERVE(ctx, b_rows, (size_t)n * sizeof(cspb_row));
        CU_TRY(ctx, cudaMemcpyAsync(b_rows.p, rows, (size_t)n * sizeof(cspb_row), cudaMemcpyHostToDevice, ctx->stream));
        d_rows = b_rows.as<cspb_row>();
    }
    RESERVE(ctx, b_sums, (size_t)n_idx * sizeof(double));
    RESERVE(ctx, b_cnt, (size_t)n_idx * sizeof(int));
    CU_TRY(ctx, cudaMemsetAsync(b_sums.p, 0, (size_t)n_idx * sizeof(double), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(b_cnt.p, 0, (size_t)n_idx * sizeof(int), ctx->stream));
    if (n > 0) {
        weights_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_rows, n, n_idx, b_sums.as<double>(), b_cnt.as<int>());
        KERNEL_CHECK(ctx);
    }
    std::vector<double> hs(n_idx);
    std::vector<int> hc(n_idx);
    CU_TRY(ctx, cudaMemcpyAsync(hs.data(), b_sums.p, (size_t)n_idx * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(hc.data(), b_cnt.p, (size_t)n_idx * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int used = 0;
    for (int t = 0; t < n_idx; ++t) {
        weights_out[t] = hc[t] > 0 ? hs[t] / (double)hc[t] : -1.0;
        if (hc[t] > 0) used = t + 1;
    }
    if (n_used_out) *n_used_out = used;
    return 0;
}

extern "C" int cspb_select_cfg_default(cspb_select_cfg *c) {
    if (!c) return CSPB_E_ARG;
    memset(c, 0, sizeof *c);
    c->cutoff = 1.f;            // reconstruct_cutoff: keep everything
    c->mindef = 0.f; c->maxdef = 100000.f;
    c->firstframe = 0; c->lastframe = -1;
    c->mintilt = -90.f; c->maxtilt = 90.f;
    c->minazh = 0.f; c->maxazh = 180.f;
    c->minscore = 0.f; c->maxscore = 1.f;
    c->renumber = 1;
    c->threshold_override = NAN;
    return 0;
}

extern "C" int cspb_select_scores(cspb_ctx *ctx, cspb_row *rows, int n, const float *tilt_angle, const cspb_select_cfg *cfg, int loc,
                                  double *threshold_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !rows || !cfg || n < 0) return CSPB_E_ARG;
    const bool have_override = !isnan(cfg->threshold_override);
    if (!have_override && !(cfg->cutoff > 0.f && cfg->cutoff <= 1.f))
        return cspb_fail(ctx, CSPB_E_ARG, "cutoff must be a fraction in (0, 1]; for the automatic cutoff (0) fit the score populations on the host and pass threshold_override");
    if (threshold_out) *threshold_out = NAN;
    if (n == 0) return 0;
    cspb_row *d_rows = rows;
    const float *d_tilt = tilt_angle;
    DevBuf b_rows, b_tilt;
    if (loc == CSPB_HOST) {
        RESERVE(ctx, b_rows, (size_t)n * sizeof(cspb_row));
        CU_TRY(ctx, cudaMemcpyAsync(b_rows.p, rows, (size_t)n * sizeof(cspb_row), cudaMemcpyHostToDevice, ctx->stream));
        d_rows = b_rows.as<cspb_row>();
        if (tilt_angle) {
            RESERVE(ctx, b_tilt, (size_t)n * sizeof(float));
            CU_TRY(ctx, cudaMemcpyAsync(b_tilt.p, tilt_angle, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            d_tilt = b_tilt.as<float>();
        }
    }
    const int g = ceil_div(n, 256);
    DevBuf b_small;
    RESERVE(ctx, b_small, 64);
    float *d_mm = b_small.as<float>();
    int *d_flags = reinterpret_cast<int *>(d_mm + 2);
    const float init_mm[2] = {INFINITY, -INFINITY};
    const int init_fl[3] = {0, -1, 0};
    CU_TRY(ctx, cudaMemcpyAsync(d_mm, init_mm, sizeof init_mm, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(d_flags, init_fl, sizeof init_fl, cudaMemcpyHostToDevice, ctx->stream));
    sel_minmax_tomo_kernel<<<g < 1024 ? g : 1024, 256, 0, ctx->stream>>>(d_rows, d_tilt, n, d_mm, d_flags);
    KERNEL_CHECK(ctx);
    float h_mm[2];
    int h_fl[3];
    CU_TRY(ctx, cudaMemcpyAsync(h_mm, d_mm, sizeof h_mm, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(h_fl, d_flags, sizeof h_fl, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const bool is_tomo = h_fl[0] != 0;
    const int n_p = h_fl[1] + 1;
    const double smin = (double)h_mm[0], smax = (double)h_mm[1];
    SelParams P;
    P.lo = cfg->minscore < 1.f ? smin + (double)cfg->minscore * (smax - smin) : (double)cfg->minscore;       // scores.py:519-523
    P.hi = cfg->maxscore <= 1.f ? smax - (1.0 - (double)cfg->maxscore) * (smax - smin) : (double)cfg->maxscore;  // :525-530
    P.cutoff_is_one = cfg->cutoff == 1.f;
    P.mindef = cfg->mindef; P.maxdef = cfg->maxdef; P.mintilt = cfg->mintilt; P.maxtilt = cfg->maxtilt;
    P.minazh = cfg->minazh; P.maxazh = cfg->maxazh; P.firstframe = cfg->firstframe; P.lastframe = cfg->lastframe;
    P.renumber = cfg->renumber;
    double threshold = NAN;
    DevBuf b_sums, b_cnt, b_keys, b_sorted, b_tmp;
    double *d_sums10 = nullptr;
    int *d_cnt10 = nullptr;
    if (is_tomo) {
        if (n_p <= 0) return cspb_fail(ctx, CSPB_E_ARG, "tilt-series table without PIND");
        // two particle tables: |tilt| <= 12 for the threshold (scores.py:455-470), |tilt| < 10 for the decision (:571-585)
        RESERVE(ctx, b_sums, (size_t)3 * n_p * sizeof(double));
        RESERVE(ctx, b_cnt, ((size_t)2 * n_p + 4) * sizeof(int));
        CU_TRY(ctx, cudaMemsetAsync(b_sums.p, 0, (size_t)3 * n_p * sizeof(double), ctx->stream));
        CU_TRY(ctx, cudaMemsetAsync(b_cnt.p, 0, ((size_t)2 * n_p + 4) * sizeof(int), ctx->stream));
        double *s12 = b_sums.as<double>(), *s10 = s12 + n_p, *means = s10 + n_p;
        int *c12 = b_cnt.as<int>(), *c10 = c12 + n_p, *d_m = c10 + n_p;
        sel_particle_sums_kernel<<<g, 256, 0, ctx->stream>>>(d_rows, d_tilt, n, 12.f, 0, s12, c12);
        KERNEL_CHECK(ctx);
        sel_particle_sums_kernel<<<g, 256, 0, ctx->stream>>>(d_rows, d_tilt, n, 10.f, 1, s10, c10);
        KERNEL_CHECK(ctx);
        d_sums10 = s10;
        d_cnt10 = c10;
        if (!have_override) {
            sel_compact_means_kernel<<<1, 1, 0, ctx->stream>>>(s12, c12, n_p, means, d_m);
            KERNEL_CHECK(ctx);
            int m = 0;
            CU_TRY(ctx, cudaMemcpyAsync(&m, d_m, sizeof m, cudaMemcpyDeviceToHost, ctx->stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            if (m > 0) {
                RESERVE(ctx, b_sorted, (size_t)m * sizeof(double));
                size_t tmp = 0;
                cub::DeviceRadixSort::SortKeys(nullptr, tmp, means, b_sorted.as<double>(), m, 0, 64, ctx->stream);
                RESERVE(ctx, b_tmp, tmp);
                cub::DeviceRadixSort::SortKeys(b_tmp.p, tmp, means, b_sorted.as<double>(), m, 0, 64, ctx->stream);
                KERNEL_CHECK(ctx);
                const int idx = (int)((double)(m - 1) * (1.0 - (double)cfg->cutoff));
                CU_TRY(ctx, cudaMemcpyAsync(&threshold, b_sorted.as<double>() + idx, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            }
        }
    } else if (!have_override) {
        RESERVE(ctx, b_keys, (size_t)2 * n * sizeof(float));
        float *keys = b_keys.as<float>(), *sorted = keys + n;
        sel_scores_kernel<<<g, 256, 0, ctx->stream>>>(d_rows, n, keys);
        KERNEL_CHECK(ctx);
        size_t tmp = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys, sorted, n, 0, 32, ctx->stream);
        RESERVE(ctx, b_tmp, tmp);
        cub::DeviceRadixSort::SortKeys(b_tmp.p, tmp, keys, sorted, n, 0, 32, ctx->stream);
        KERNEL_CHECK(ctx);
        const int idx = (int)((double)(n - 1) * (1.0 - (double)cfg->cutoff));
        float t = 0.f;
        CU_TRY(ctx, cudaMemcpyAsync(&t, sorted + idx, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        threshold = (double)t;
    }
    if (have_override) threshold = (double)cfg->threshold_override;
    P.threshold = threshold;
    P.tomo_rule = is_tomo && threshold > 0.0;  // scores.py:571 (a NaN threshold takes the plain branch and compares false)
    sel_apply_kernel<<<g, 256, 0, ctx->stream>>>(d_rows, d_tilt, n, P, d_sums10, d_cnt10);
    KERNEL_CHECK(ctx);
    if (loc == CSPB_HOST) CU_TRY(ctx, cudaMemcpyAsync(rows, d_rows, (size_t)n * sizeof(cspb_row), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (threshold_out) *threshold_out = threshold;
    return 0;
}

extern "C" int cspb_class_occupancies(cspb_ctx *ctx, const float *logp, const float *sigma, const double *class_average_occ, int n_classes,
                                      int n, float *occ_out, float *sigma_out, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !logp || !sigma || !class_average_occ || !occ_out || !sigma_out || n_classes < 1 || n < 0) return CSPB_E_ARG;
    if (n == 0) return 0;
    const size_t kn = (size_t)n_classes * n;
    DevBuf b_in, b_out, b_avg;
    const float *d_logp = logp, *d_sigma = sigma;
    float *d_occ = occ_out, *d_sig = sigma_out;
    RESERVE(ctx, b_avg, (size_t)n_classes * sizeof(double));
    CU_TRY(ctx, cudaMemcpyAsync(b_avg.p, class_average_occ, (size_t)n_classes * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (loc == CSPB_HOST) {
        RESERVE(ctx, b_in, 2 * kn * sizeof(float));
        RESERVE(ctx, b_out, (kn + n) * sizeof(float));
        CU_TRY(ctx, cudaMemcpyAsync(b_in.p, logp, kn * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(ctx, cudaMemcpyAsync(b_in.as<float>() + kn, sigma, kn * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        d_logp = b_in.as<float>(); d_sigma = d_logp + kn;
        d_occ = b_out.as<float>(); d_sig = d_occ + kn;
    }
    occ_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(d_logp, d_sigma, b_avg.as<double>(), n_classes, n, d_occ, d_sig);
    KERNEL_CHECK(ctx);
    if (loc == CSPB_HOST) {
        CU_TRY(ctx, cudaMemcpyAsync(occ_out, d_occ, kn * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaMemcpyAsync(sigma_out, d_sig, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
