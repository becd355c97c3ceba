// pipeline.cu — host-stack pipeline: one upload per projection, copies overlapped with compute.
//
// pyp runs refine3d and reconstruct3d as separate processes that each read the particle stack
// (src/pyp/refine/frealign/frealign.py:3918-3994, 1780-1824).  In process, the two stages can
// share one host->device copy of every chunk: while chunk k is preprocessed, refined and inserted on
// the compute stream, chunk k+1 is copied on a second stream into the other staging buffer.
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <math.h>
#include <vector>
#include "internal.cuh"

namespace {
// restores the kept-spectra switches on every way out of a pipeline call
struct KeepScope {
    cspb_ctx *c;
    bool was;
    KeepScope(cspb_ctx *ctx, bool on, const float *span, int span_count) : c(ctx), was(ctx->keep_on) {
        c->keep_on = on;
        c->keep_span_src = span;
        c->keep_span_count = span_count;
    }
    ~KeepScope() {
        c->keep_on = was;
        c->keep_span_src = nullptr;
        c->keep_span_count = 0;
        c->keep_count = 0;  // the buffers the spectra were made from are about to be reused
        c->keep_src = nullptr;
    }
};
}  // namespace

// Batch schedule of the streamed pipeline.  Per-projection work does not depend on the batching (the whitening curve is
// estimated on the first min(n, 4096) images whatever the chunking), so the results equal those of the staged calls.
// Whole waves of the scorer (W = SMs x resident CTAs x 4 units, cspb_wave_units), 2W per batch — the kernels need ~0.65 of a
// batch's copy time, so every batch is done before the next has landed — after a first batch that holds the 4 096 images the
// whitening curve is estimated on, shrinking W, W, W/2 at the end: the call ends one batch-processing time after the last
// copy lands (a 2W batch straight before W, W/2 leaves the kernels 4 ms behind the copies).  r02zb timeline: with 4W batches
// in the middle the 9 472-image batch landed at 103 ms, took 30 ms and pushed the end of the call 19 ms behind the last copy.
// Bounded by ~6 GB of staging per buffer.
static std::vector<int> pipeline_batches(long long n_images, int n, long long W) {
    std::vector<int> sizes;
    if (n_images <= 0) return sizes;
    if (W < 1) W = 1;
    long long cap = (long long)(((size_t)6 << 30) / ((size_t)n * n * sizeof(float)));
    const long long want = 2 * W > 4096 ? 2 * W : 4096;
    if (cap > want) cap = want;
    if (cap < 1) cap = 1;
    std::vector<int> tail;
    long long rem = n_images;
    if (rem >= 8 * W && 2 * W <= cap && W >= 64) { tail = {(int)W, (int)W, (int)(W / 2)}; rem -= 2 * W + W / 2; }
    else if (rem >= 3 * W && W <= cap) { tail = {(int)W}; rem -= W; }
    long long first = W > 4096 ? W : 4096;
    if (first > cap) first = cap;
    if (first > rem) first = rem;
    rem -= first;
    const long long full = cap < 2 * W ? cap : 2 * W;
    const long long odd = rem % full;  // what the full batches leave over goes second (no crumbs: onto the first batch)
    if (odd > 0 && odd < W / 2) first += odd;
    sizes.push_back((int)first);
    if (odd >= W / 2) sizes.push_back((int)odd);
    for (long long k = 0; k < rem / full; ++k) sizes.push_back((int)full);
    sizes.insert(sizes.end(), tail.begin(), tail.end());
    return sizes;
}

// the schedule above for a given stack, box and scorer wave (no device needed: exposed for the host-side tests)
extern "C" int cspb_pipeline_batches(int n_images, int box, int wave_units, int *sizes_out, int max_sizes) {
    if (n_images < 0 || box < 2 || wave_units < 1 || (!sizes_out && max_sizes > 0)) return CSPB_E_ARG;
    const std::vector<int> s = pipeline_batches(n_images, box, wave_units);
    for (int k = 0; k < (int)s.size() && k < max_sizes; ++k) sizes_out[k] = s[k];
    return (int)s.size();
}

extern "C" int cspb_refine_reconstruct(cspb_ctx *ctx, const float *images_host, cspb_row *rows_host, int n_images, int flags,
                                       int64_t *n_evals_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !images_host || !rows_host || n_images < 0) return CSPB_E_ARG;
    const bool do_refine = flags & CSPB_DO_REFINE, do_insert = flags & CSPB_DO_INSERT;
    if (!do_refine && !do_insert) return cspb_fail(ctx, CSPB_E_ARG, "nothing to do (flags = %d)", flags);
    if (do_refine && (!ctx->refine_ready || !ctx->ref.ready)) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (do_insert && !ctx->recon_ready) return cspb_fail(ctx, CSPB_E_STATE, "cspb_recon_begin first");
    const int n = do_refine ? ctx->rcfg.box : ctx->ccfg.box;
    if (do_refine && do_insert && ctx->ccfg.box != ctx->rcfg.box) return cspb_fail(ctx, CSPB_E_ARG, "refine and reconstruct boxes differ");
    if (n_evals_out) *n_evals_out = 0;
    if (n_images == 0) return 0;
    const std::vector<int> sizes = pipeline_batches(n_images, n, cspb_wave_units(ctx));
    int chunk = 0;
    for (int v : sizes) chunk = v > chunk ? v : chunk;
    struct { cudaStream_t &copy; cudaEvent_t *ready, *freed; DevBuf *stage; DevBuf &rows; } P = {
        ctx->pipe_copy, ctx->pipe_ready, ctx->pipe_freed, ctx->pipe_stage, ctx->pipe_rows};
    const bool dbg = getenv("CSPB_PIPE_DEBUG") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    if (!P.copy) {
        CU_TRY(ctx, cudaStreamCreateWithFlags(&P.copy, cudaStreamNonBlocking));
        for (int k = 0; k < CSPB_PIPE_STAGES; ++k) {
            CU_TRY(ctx, cudaEventCreateWithFlags(&P.ready[k], cudaEventDisableTiming));
            CU_TRY(ctx, cudaEventCreateWithFlags(&P.freed[k], cudaEventDisableTiming));
        }
    }
    for (int k = 0; k < CSPB_PIPE_STAGES; ++k) RESERVE(ctx, P.stage[k], (size_t)chunk * n * n * sizeof(float));
    RESERVE(ctx, P.rows, (size_t)n_images * sizeof(cspb_row));
    if (dbg) fprintf(stderr, "pipe: buffers ready at %.1f ms\n", ms_since());
    CU_TRY(ctx, cudaMemcpyAsync(P.rows.p, rows_host, (size_t)n_images * sizeof(cspb_row), cudaMemcpyHostToDevice, ctx->stream));
    int64_t evals = 0;
    int rc = 0, idx = 0;
    // refine and insert see the same staged pixels: one forward transform per projection (cspb_refine_keep_spectra)
    KeepScope keep_scope(ctx, do_refine && do_insert, nullptr, 0);
    for (int s = 0; idx < (int)sizes.size() && !rc; s += sizes[idx], ++idx) {
        const int b = idx % CSPB_PIPE_STAGES;
        const int cnt = sizes[idx];
        if (idx >= CSPB_PIPE_STAGES) CU_TRY(ctx, cudaStreamWaitEvent(P.copy, P.freed[b], 0));
        CU_TRY(ctx, cudaMemcpyAsync(P.stage[b].p, images_host + (size_t)s * n * n, (size_t)cnt * n * n * sizeof(float),
                                    cudaMemcpyHostToDevice, P.copy));
        CU_TRY(ctx, cudaEventRecord(P.ready[b], P.copy));
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, P.ready[b], 0));
        cspb_row *d_rows = P.rows.as<cspb_row>() + s;
        if (do_refine) {
            rc = cspb_refine_load_images(ctx, P.stage[b].as<float>(), cnt, CSPB_DEVICE, 0);
            int64_t ev = 0;
            if (!rc) rc = cspb_refine_run_device(ctx, d_rows, cnt, &ev);
            evals += ev;
        }
        if (!rc && do_insert) rc = cspb_recon_insert(ctx, P.stage[b].as<float>(), d_rows, cnt, CSPB_DEVICE);
        if (!rc) CU_TRY(ctx, cudaEventRecord(P.freed[b], ctx->stream));
        if (dbg) fprintf(stderr, "pipe: chunk %d (%d images) enqueued at %.1f ms\n", idx, cnt, ms_since());
    }
    if (dbg) {
        cudaStreamSynchronize(P.copy);
        fprintf(stderr, "pipe: copies done at %.1f ms\n", ms_since());
    }
    if (!rc)
        CU_TRY(ctx, cudaMemcpyAsync(rows_host, P.rows.p, (size_t)n_images * sizeof(cspb_row), cudaMemcpyDeviceToHost, ctx->stream));
    cudaStreamSynchronize(P.copy);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (dbg) fprintf(stderr, "pipe: all done at %.1f ms\n", ms_since());
    if (n_evals_out) *n_evals_out = evals;
    return rc;
}


// refine3d -> score shaping (select.cu) -> reconstruct3d without leaving the device.  The score
// threshold is a quantile over ALL rows, so insertion cannot start before the last batch is refined: the stack is
// uploaded once into a resident buffer (no double buffering, the copy stream never waits for the kernels), batches are
// refined as they land, then the table is shaped and the whole resident stack inserted.
extern "C" int cspb_refine_select_reconstruct(cspb_ctx *ctx, const float *images_host, cspb_row *rows_host, int n_images,
                                              const cspb_select_cfg *cfg, int64_t *n_evals_out, double *threshold_out) {
    CSPB_ENTER(ctx);
    if (!ctx || !images_host || !rows_host || !cfg || n_images < 0) return CSPB_E_ARG;
    if (!ctx->refine_ready || !ctx->ref.ready) return cspb_fail(ctx, CSPB_E_STATE, "configure + set_reference first");
    if (!ctx->recon_ready) return cspb_fail(ctx, CSPB_E_STATE, "cspb_recon_begin first");
    if (ctx->ccfg.box != ctx->rcfg.box) return cspb_fail(ctx, CSPB_E_ARG, "refine and reconstruct boxes differ");
    if (n_evals_out) *n_evals_out = 0;
    if (threshold_out) *threshold_out = NAN;
    if (n_images == 0) return 0;
    const int n = ctx->rcfg.box;
    const size_t img_bytes = (size_t)n * n * sizeof(float);
    size_t free_b = 0, total_b = 0;
    CU_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
    if ((size_t)n_images * img_bytes > ctx->pipe_all.bytes && (size_t)n_images * img_bytes + ((size_t)8 << 30) > free_b + ctx->pipe_all.bytes)
        return cspb_fail(ctx, CSPB_E_NOMEM, "a resident stack of %d images (%.1f GB) does not fit the device: use smaller ranges", n_images,
                         (double)n_images * img_bytes / 1e9);
    RESERVE(ctx, ctx->pipe_all, (size_t)n_images * img_bytes);
    RESERVE(ctx, ctx->pipe_rows, (size_t)n_images * sizeof(cspb_row));
    if (!ctx->pipe_copy) {
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->pipe_copy, cudaStreamNonBlocking));
        for (int k = 0; k < CSPB_PIPE_STAGES; ++k) {
            CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_ready[k], cudaEventDisableTiming));
            CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->pipe_freed[k], cudaEventDisableTiming));
        }
    }
    float *d_all = ctx->pipe_all.as<float>();
    cspb_row *d_rows = ctx->pipe_rows.as<cspb_row>();
    CU_TRY(ctx, cudaMemcpyAsync(d_rows, rows_host, (size_t)n_images * sizeof(cspb_row), cudaMemcpyHostToDevice, ctx->stream));
    // batches: the first holds the whitening sample (4 096 images), then whole multiples of a scorer wave
    const long long W = cspb_wave_units(ctx);
    std::vector<int> sizes;
    {
        long long rem = n_images, s = W > 4096 ? W : 4096;
        while (rem > 0) {
            long long take = s < rem ? s : rem;
            if (rem - take < W / 2) take = rem;
            sizes.push_back((int)take);
            rem -= take;
            s = 4 * W;
        }
    }
    std::vector<cudaEvent_t> ready(sizes.size());
    int64_t evals = 0;
    int rc = 0;
    size_t off = 0;
    for (size_t k = 0; k < sizes.size(); ++k) {  // all copies are queued up front: the copy engine runs back to back
        CU_TRY(ctx, cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming));
        CU_TRY(ctx, cudaMemcpyAsync(d_all + off * n * n, images_host + off * n * n, (size_t)sizes[k] * img_bytes, cudaMemcpyHostToDevice, ctx->pipe_copy));
        CU_TRY(ctx, cudaEventRecord(ready[k], ctx->pipe_copy));
        off += sizes[k];
    }
    off = 0;
    // the batches are consecutive pieces of the resident stack: their forward transforms are kept for the insertion
    KeepScope keep_scope(ctx, true, d_all, n_images);
    for (size_t k = 0; k < sizes.size() && !rc; ++k) {
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ready[k], 0));
        rc = cspb_refine_load_images(ctx, d_all + off * n * n, sizes[k], CSPB_DEVICE, 0);
        int64_t ev = 0;
        if (!rc) rc = cspb_refine_run_device(ctx, d_rows + off, sizes[k], &ev);
        evals += ev;
        off += sizes[k];
    }
    if (!rc) rc = cspb_select_scores(ctx, d_rows, n_images, nullptr, cfg, CSPB_DEVICE, threshold_out);
    if (!rc) rc = cspb_recon_insert(ctx, d_all, d_rows, n_images, CSPB_DEVICE);
    if (!rc) CU_TRY(ctx, cudaMemcpyAsync(rows_host, d_rows, (size_t)n_images * sizeof(cspb_row), cudaMemcpyDeviceToHost, ctx->stream));
    cudaStreamSynchronize(ctx->pipe_copy);
    cudaStreamSynchronize(ctx->stream);
    for (auto e : ready) cudaEventDestroy(e);
    if (n_evals_out) *n_evals_out = evals;
    return rc;
}
