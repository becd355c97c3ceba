// capi.cu — context lifecycle and the building-block entry points of include/cspb200.h.
#include <cufft.h>
#include <stdio.h>
#include <string.h>
#include "internal.cuh"

static_assert(sizeof(cspb_row) == 128, "cspb_row must match the 128-byte .cistem row");

// number of visible CUDA devices (0 without a driver / GPU); lets the front-ends spread their
// particle ranges over the GPUs without importing a framework
extern "C" int cspb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int cspb_abi_version(void) { return CSPB_ABI_VERSION; }

extern "C" int cspb_create(int device, cspb_ctx **out) {
    if (!out) return CSPB_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return CSPB_E_CUDA;  // no CPU fallback by design
    }
    if (device < 0 || device >= count) return CSPB_E_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CSPB_E_CUDA;
    if (prop.major < 10) return CSPB_E_CUDA;  // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return CSPB_E_CUDA;
    cspb_ctx *ctx = new (std::nothrow) cspb_ctx();
    if (!ctx) return CSPB_E_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return CSPB_E_CUDA;
    }
    *out = ctx;
    return 0;
}

extern "C" int cspb_destroy(cspb_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->pipe_copy) {
        cudaStreamSynchronize(ctx->pipe_copy);
        for (int k = 0; k < CSPB_PIPE_STAGES; ++k) {
            if (ctx->pipe_ready[k]) cudaEventDestroy(ctx->pipe_ready[k]);
            if (ctx->pipe_freed[k]) cudaEventDestroy(ctx->pipe_freed[k]);
        }
        cudaStreamDestroy(ctx->pipe_copy);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

extern "C" const char *cspb_last_error(const cspb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int cspb_sync(cspb_ctx *ctx) {
    CSPB_ENTER(ctx);
    if (!ctx) return CSPB_E_ARG;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cspb_stream(cspb_ctx *ctx, void **stream_out) {
    if (!ctx || !stream_out) return CSPB_E_ARG;
    *stream_out = (void *)ctx->stream;
    return 0;
}

extern "C" int64_t cspb_launch_count(const cspb_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int cspb_fft2_r2c(cspb_ctx *ctx, const float *in, float *out_complex, int n, int batch, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !in || !out_complex || n < 2 || batch < 1) return CSPB_E_ARG;
    const int nh = n / 2 + 1;
    const size_t in_b = (size_t)batch * n * n * sizeof(float), out_b = (size_t)batch * n * nh * sizeof(float2);
    if (loc == CSPB_DEVICE) return fft2_r2c_dev(ctx, in, reinterpret_cast<float2 *>(out_complex), n, batch, nullptr, nullptr);
    RESERVE(ctx, ctx->d_work0, in_b);
    RESERVE(ctx, ctx->d_work1, out_b);
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_work0.p, in, in_b, cudaMemcpyHostToDevice, ctx->stream));
    int rc = fft2_r2c_dev(ctx, ctx->d_work0.as<float>(), ctx->d_work1.as<float2>(), n, batch, nullptr, nullptr);
    if (rc) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(out_complex, ctx->d_work1.p, out_b, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int cspb_fft2_c2r(cspb_ctx *ctx, const float *in_complex, float *out, int n, int batch, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !in_complex || !out || n < 2 || batch < 1) return CSPB_E_ARG;
    const int nh = n / 2 + 1;
    const size_t out_b = (size_t)batch * n * n * sizeof(float), in_b = (size_t)batch * n * nh * sizeof(float2);
    RESERVE(ctx, ctx->d_work1, in_b);
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_work1.p, in_complex, in_b,
                                loc == CSPB_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    if (loc == CSPB_DEVICE) return fft2_c2r_dev(ctx, ctx->d_work1.as<float2>(), out, n, batch);
    RESERVE(ctx, ctx->d_work0, out_b);
    int rc = fft2_c2r_dev(ctx, ctx->d_work1.as<float2>(), ctx->d_work0.as<float>(), n, batch);
    if (rc) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_work0.p, out_b, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// cuFFT twin of cspb_fft2_r2c: comparison baseline for tests / bench only.
extern "C" int cspb_cufft2_r2c(cspb_ctx *ctx, const float *in, float *out_complex, int n, int batch, int loc) {
    CSPB_ENTER(ctx);
    if (!ctx || !in || !out_complex || n < 2 || batch < 1) return CSPB_E_ARG;
    const int nh = n / 2 + 1;
    const size_t in_b = (size_t)batch * n * n * sizeof(float), out_b = (size_t)batch * n * nh * sizeof(float2);
    const float *d_in = in;
    float2 *d_out = reinterpret_cast<float2 *>(out_complex);
    if (loc == CSPB_HOST) {
        RESERVE(ctx, ctx->d_work0, in_b);
        RESERVE(ctx, ctx->d_work1, out_b);
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_work0.p, in, in_b, cudaMemcpyHostToDevice, ctx->stream));
        d_in = ctx->d_work0.as<float>();
        d_out = ctx->d_work1.as<float2>();
    }
    static cufftHandle plan = 0;
    static int plan_n = 0, plan_b = 0;
    if (plan_n != n || plan_b != batch) {
        if (plan) cufftDestroy(plan);
        int dims[2] = {n, n};
        if (cufftPlanMany(&plan, 2, dims, nullptr, 1, n * n, nullptr, 1, n * nh, CUFFT_R2C, batch) != CUFFT_SUCCESS) {
            plan = 0; plan_n = 0;
            return cspb_fail(ctx, CSPB_E_CUDA, "cufftPlanMany failed");
        }
        plan_n = n; plan_b = batch;
    }
    cufftSetStream(plan, ctx->stream);
    if (cufftExecR2C(plan, const_cast<float *>(d_in), reinterpret_cast<cufftComplex *>(d_out)) != CUFFT_SUCCESS)
        return cspb_fail(ctx, CSPB_E_CUDA, "cufftExecR2C failed");
    if (loc == CSPB_HOST) {
        CU_TRY(ctx, cudaMemcpyAsync(out_complex, d_out, out_b, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

// ---------------------------------------------------------------- live kernel timing
void prof_begin(cspb_ctx *ctx, int kind, int64_t units) {
    if (!ctx->prof_on) return;
    ProfRec r;
    r.kind = kind;
    r.units = units;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, ctx->stream);
    ctx->prof.push_back(r);
}

void prof_end(cspb_ctx *ctx) {
    if (!ctx->prof_on || ctx->prof.empty()) return;
    cudaEventRecord(ctx->prof.back().b, ctx->stream);
}

extern "C" int cspb_profile_enable(cspb_ctx *ctx, int on) {
    CSPB_ENTER(ctx);
    if (!ctx) return CSPB_E_ARG;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    ctx->prof_on = on != 0;
    return 0;
}

extern "C" int cspb_profile_get(cspb_ctx *ctx, int kind, double *total_ms, int64_t *launches, int64_t *units) {
    CSPB_ENTER(ctx);
    if (!ctx || kind < 0 || kind >= CSPB_PROF_KINDS) return CSPB_E_ARG;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double t = 0.0;
    int64_t n = 0, u = 0;
    for (auto &r : ctx->prof) {
        if (r.kind != kind) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) t += ms;
        ++n;
        u += r.units;
    }
    if (total_ms) *total_ms = t;
    if (launches) *launches = n;
    if (units) *units = u;
    return 0;
}

// Census of the scorer's gather loads (see score_census_kernel): while on, every scorer launch is followed
// by a counting launch over the same units.  For roofline bookkeeping only — switch it off before timing.
extern "C" int cspb_profile_count_loads(cspb_ctx *ctx, int on) {
    CSPB_ENTER(ctx);
    if (!ctx) return CSPB_E_ARG;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (on) {
        RESERVE(ctx, ctx->d_load_count, 2 * sizeof(unsigned long long));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_load_count.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
        ctx->census_evals = 0;
    }
    ctx->count_loads = on != 0;
    return 0;
}

extern "C" int cspb_profile_get_loads(cspb_ctx *ctx, int64_t *quad_loads, int64_t *slot_reads, int64_t *evals) {
    CSPB_ENTER(ctx);
    if (!ctx || !ctx->d_load_count.p) return CSPB_E_STATE;
    unsigned long long v[2] = {0, 0};
    CU_TRY(ctx, cudaMemcpyAsync(v, ctx->d_load_count.p, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (quad_loads) *quad_loads = (int64_t)v[0];
    if (slot_reads) *slot_reads = (int64_t)v[1];
    if (evals) *evals = ctx->census_evals;
    return 0;
}
