// opt.cuh — the batched, branch-free optimiser shared by refine3d (refine.cu: one state per image
// or search hit) and csp (csp.cu: one state per particle or tilt).  Semantics: oracle/SEMANTICS.md §7.
#pragma once
#include "internal.cuh"

namespace {

// Per image: x = {psi, theta, phi, shift x, shift y, defocus delta}, masked by `free`.
// Every iteration = central-difference stencil (1+2M evals) -> diagonal Newton step with a
// trust region -> 3-point line search -> keep the best point seen.  All images in lockstep.
#define OPT_NP 6
#define OPT_NL 3
struct OptState {
    float x[OPT_NP];
    float h[OPT_NP];
    float d[OPT_NP];
    float x0[OPT_NP];  // starting pose
    float f;           // objective at the current centre
    float lam;         // shift-restraint scale sigma^2 / N_mask of the state's image (0 without priors)
    float pad_[2];
};

// Shift restraint (refine3d prompt 7, oracle/SEMANTICS.md §7b): objective = CC - lam * (wx (sx - mx)^2 +
// wy (sy - my)^2), w = 1 / (2 var).  `on` = 0 (csp, refine3d without priors): the objective is the CC.
struct OptPrior {
    int on;
    float mx, my, wx, wy;
};
__device__ __forceinline__ float prior_pen(const OptPrior pr, float lam, float sx, float sy) {
    if (!pr.on) return 0.f;
    const float dx = sx - pr.mx, dy = sy - pr.my;
    return lam * (pr.wx * dx * dx + pr.wy * dy * dy);
}

__device__ __forceinline__ float wrap360(float a) {
    a = fmodf(a, 360.f);
    if (a < 0.f) a += 360.f;
    return a;
}

__device__ __forceinline__ float wrap180(float d) { return d - 360.f * rintf(d * (1.f / 360.f)); }

// Stencil evaluation order of the parameters: pure in-plane shifts last, so that the evaluations
// which share the centre's rotation and CTF are consecutive and can be scored from ONE gather.
__device__ __constant__ const int OPT_ORDER[OPT_NP] = {0, 1, 2, 5, 3, 4};

// free_mask bit m set -> parameter m is refined.  n_free = popcount.  evals per state NE = 1+2*n_free.
// shift_mask: free parameters that are pure image shifts (refine3d: x, y); their 2*n_shift evaluations
// form "shared" units (rotation and CTF of the unit's first pose apply to all of its poses).
// Unit layout by class, so that every scorer launch sees one kind of unit (launch_score_classes):
//   [A full: n*(nA/PB)] [A tail: n if nA%PB] [S full: n*(nS/PB)] [S tail: n if nS%PB],  nA = NE - nS.
__global__ void opt_stencil_kernel(const OptState *__restrict__ st, int n, int K, int free_mask, int shift_mask, int NE, int PB,
                                   float *__restrict__ poses6, ScoreUnit *__restrict__ units) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const OptState s = st[k];
    float *q = poses6 + (long long)k * NE * 6;
    for (int m = 0; m < OPT_NP; ++m) q[m] = s.x[m];
    int e = 1, nS = 0;
    for (int o = 0; o < OPT_NP; ++o) {
        const int m = OPT_ORDER[o];
        if (!((free_mask >> m) & 1)) continue;
        if ((shift_mask >> m) & 1) nS += 2;
        for (int sgn = 0; sgn < 2; ++sgn, ++e) {
            float *qe = q + e * 6;
            for (int t = 0; t < OPT_NP; ++t) qe[t] = s.x[t];
            qe[m] += sgn ? -s.h[m] : s.h[m];
        }
    }
    const int nA = NE - nS;
    const int fullA = nA / PB, tailA = nA % PB, fullS = nS / PB, tailS = nS % PB;
    long long base = 0;
    for (int c = 0; c < fullA; ++c) {
        ScoreUnit un;
        un.image = k / K; un.first_eval = k * NE + c * PB; un.count = PB; un.pad_ = 0;
        units[base + (long long)k * fullA + c] = un;
    }
    base += (long long)n * fullA;
    if (tailA) {
        ScoreUnit un;
        un.image = k / K; un.first_eval = k * NE + fullA * PB; un.count = tailA; un.pad_ = 0;
        units[base + k] = un;
        base += n;
    }
    for (int c = 0; c < fullS; ++c) {
        ScoreUnit un;
        un.image = k / K; un.first_eval = k * NE + nA + c * PB; un.count = PB; un.pad_ = 1;
        units[base + (long long)k * fullS + c] = un;
    }
    base += (long long)n * fullS;
    if (tailS) {
        ScoreUnit un;
        un.image = k / K; un.first_eval = k * NE + nA + fullS * PB; un.count = tailS; un.pad_ = 1;
        units[base + k] = un;
    }
}

__device__ __forceinline__ float cc_of(const float4 v) {
    const float den = v.z * v.w;
    return den > 0.f ? v.x * rsqrtf(den) : 0.f;
}

// diagonal Newton step, continuous in (f0, fp, fm): d = g / max(-c, |g|/dmax), dmax = 4h
__device__ __forceinline__ float newton_step(float f0, float fp, float fm, float h) {
    const float g = (fp - fm) / (2.f * h);
    const float c = (fp - 2.f * f0 + fm) / (h * h);
    const float dmax = 4.f * h;
    float den = -c;
    const float floor_ = fabsf(g) / dmax;
    if (den < floor_) den = floor_;
    return den > 0.f ? g / den : 0.f;
}

// step length from f(0) and f(0.5), f(1), f(2): least-squares parabola through the origin
// offset, maximiser clamped to [0, 2.5]; a convex fit takes the better end of the interval
__device__ __forceinline__ float line_step(float f0, const float *fl) {
    const float tl[OPT_NL] = {0.5f, 1.f, 2.f};
    float s22 = 0.f, s23 = 0.f, s33 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
    for (int l = 0; l < OPT_NL; ++l) {
        const float t = tl[l], y = fl[l] - f0;
        s22 += t * t; s23 += t * t * t; s33 += t * t * t * t;
        r2 += t * y; r3 += t * t * y;
    }
    const float det = s22 * s33 - s23 * s23;
    const float b = (r2 * s33 - r3 * s23) / det;
    const float a = (r3 * s22 - r2 * s23) / det;
    if (a < 0.f) {
        float t = -b / (2.f * a);
        return fminf(fmaxf(t, 0.f), 2.5f);
    }
    return (a * 2.5f + b > 0.f) ? 2.5f : 0.f;  // convex fit: better end of [0, 2.5] under the model
}

// consume stencil scores, propose the Newton direction, emit line-search poses
__global__ void opt_step_kernel(OptState *__restrict__ st, int n, int K, int free_mask, int NE, const float4 *__restrict__ sc,
                                float *__restrict__ poses6_ls, ScoreUnit *__restrict__ units_ls, const OptPrior pr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    OptState s = st[k];
    const float4 *v = sc + (long long)k * NE;
    const float f0 = cc_of(v[0]) - prior_pen(pr, s.lam, s.x[3], s.x[4]);
    int e = 1;
    for (int m = 0; m < OPT_NP; ++m) s.d[m] = 0.f;
    for (int o = 0; o < OPT_NP; ++o) {
        const int m = OPT_ORDER[o];
        if (!((free_mask >> m) & 1)) continue;
        const float hx = m == 3 ? s.h[m] : 0.f, hy = m == 4 ? s.h[m] : 0.f;
        const float fp = cc_of(v[e]) - prior_pen(pr, s.lam, s.x[3] + hx, s.x[4] + hy);
        const float fm = cc_of(v[e + 1]) - prior_pen(pr, s.lam, s.x[3] - hx, s.x[4] - hy);
        s.d[m] = newton_step(f0, fp, fm, s.h[m]);
        e += 2;
    }
    s.f = f0;
    st[k] = s;
    const float tl[OPT_NL] = {0.5f, 1.f, 2.f};
    float *q = poses6_ls + (long long)k * OPT_NL * 6;
    for (int l = 0; l < OPT_NL; ++l)
        for (int m = 0; m < OPT_NP; ++m) q[l * 6 + m] = s.x[m] + tl[l] * s.d[m];
    ScoreUnit un;
    un.image = k / K; un.first_eval = k * OPT_NL; un.count = OPT_NL; un.pad_ = 0;
    units_ls[k] = un;
}

__global__ void opt_select_kernel(OptState *__restrict__ st, int n, const float4 *__restrict__ sc_ls, float shrink,
                                  const OptPrior pr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    OptState s = st[k];
    float fl[OPT_NL];
    const float tl[OPT_NL] = {0.5f, 1.f, 2.f};
    for (int l = 0; l < OPT_NL; ++l)
        fl[l] = cc_of(sc_ls[(long long)k * OPT_NL + l]) - prior_pen(pr, s.lam, s.x[3] + tl[l] * s.d[3], s.x[4] + tl[l] * s.d[4]);
    const float t = line_step(s.f, fl);
    for (int m = 0; m < OPT_NP; ++m) s.x[m] += t * s.d[m];
    for (int m = 0; m < OPT_NP; ++m) s.h[m] *= shrink;
    st[k] = s;
}

// ------------------------------------------------------------------ analytic optimiser (oracle/SEMANTICS.md §7c)
// Same arithmetic, in the same order, as lm_step / lm_line / solve_spd5 of oracle/cspb_oracle.c.
#define LM_NG 5
#define LM_GOUT 28
#define LM_DAMP 0.05f
#define LM_CC_FLOOR 0.01f
#define LM_TAU 2e-5f

// one gradient evaluation per state at its current pose
__global__ void lm_pose_kernel(const OptState *__restrict__ st, int n, int K, float *__restrict__ poses6, ScoreUnit *__restrict__ units) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    for (int m = 0; m < OPT_NP; ++m) poses6[(long long)k * 6 + m] = st[k].x[m];
    ScoreUnit un;
    un.image = k / K; un.first_eval = k; un.count = 1; un.pad_ = 0;
    units[k] = un;
}

__device__ __forceinline__ bool lm_solve_spd5(const float *H, const float *g, float *d) {
    float L[LM_NG][LM_NG];
    for (int i = 0; i < LM_NG; ++i)
        for (int j = 0; j < LM_NG; ++j) L[i][j] = 0.f;
    for (int i = 0; i < LM_NG; ++i)
        for (int j = 0; j <= i; ++j) {
            float s = H[i * LM_NG + j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            if (i == j) {
                if (!(s > 0.f)) return false;
                L[i][i] = sqrtf(s);
            } else
                L[i][j] = s / L[j][j];
        }
    float y[LM_NG];
    for (int i = 0; i < LM_NG; ++i) {
        float s = g[i];
        for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
        y[i] = s / L[i][i];
    }
    for (int i = LM_NG - 1; i >= 0; --i) {
        float s = y[i];
        for (int k = i + 1; k < LM_NG; ++k) s -= L[k][i] * d[k];
        d[i] = s / L[i][i];
    }
    return true;
}

// consume the gradient evaluation: Gauss-Newton step with damping and trust region, trial pose x + d
__global__ void lm_step_kernel(OptState *__restrict__ st, int n, int K, int free_mask, const float *__restrict__ gout, float trust_ang,
                               float trust_shift, float *__restrict__ poses6, ScoreUnit *__restrict__ units, const OptPrior pr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    OptState s = st[k];
    const float *o = gout + (long long)k * LM_GOUT;
    const float A = o[2], B = o[3];
    const float den = A * B;
    const float rs = den > 0.f ? 1.f / sqrtf(den) : 0.f;
    const float cc = o[0] * rs;
    float g[LM_NG], H[LM_NG * LM_NG], d[LM_NG];
    for (int a = 0; a < LM_NG; ++a) g[a] = o[4 + a] * rs - (a < 3 && B > 0.f ? 0.5f * cc * o[9 + a] / B : 0.f);
    const float cce = cc > LM_CC_FLOOR ? cc : LM_CC_FLOOR;
    const float hs = B > 0.f ? cce / B : 0.f;
    int t = 0;
    for (int a = 0; a < LM_NG; ++a)
        for (int b = a; b < LM_NG; ++b) { H[a * LM_NG + b] = H[b * LM_NG + a] = hs * o[12 + t]; ++t; }
    float f0 = cc;
    if (pr.on) {
        const float w[2] = {pr.wx, pr.wy};
        const float dx[2] = {s.x[3] - pr.mx, s.x[4] - pr.my};
        for (int q = 0; q < 2; ++q) {
            f0 -= s.lam * w[q] * dx[q] * dx[q];
            g[3 + q] -= 2.f * s.lam * w[q] * dx[q];
            H[(3 + q) * LM_NG + 3 + q] += 2.f * s.lam * w[q];
        }
    }
    for (int a = 0; a < LM_NG; ++a)
        if (!((free_mask >> a) & 1)) {
            for (int b = 0; b < LM_NG; ++b) H[a * LM_NG + b] = H[b * LM_NG + a] = 0.f;
            H[a * LM_NG + a] = 1.f;
            g[a] = 0.f;
        }
    for (int a = 0; a < LM_NG; ++a) H[a * LM_NG + a] *= 1.f + LM_DAMP;
    if (!lm_solve_spd5(H, g, d))
        for (int a = 0; a < LM_NG; ++a) d[a] = H[a * LM_NG + a] > 0.f ? g[a] / H[a * LM_NG + a] : 0.f;
    float worst = 1.f;
    for (int a = 0; a < LM_NG; ++a) {
        const float q = fabsf(d[a]) / (a < 3 ? trust_ang : trust_shift);
        if (q > worst) worst = q;
    }
    float slope = 0.f;
    for (int a = 0; a < LM_NG; ++a) {
        d[a] /= worst;
        slope += g[a] * d[a];
    }
    for (int a = 0; a < LM_NG; ++a) s.d[a] = d[a];
    s.d[5] = 0.f;
    s.f = f0;
    s.pad_[0] = slope;
    st[k] = s;
    float *q = poses6 + (long long)k * 6;
    for (int m = 0; m < OPT_NP; ++m) q[m] = s.x[m] + s.d[m];
    ScoreUnit un;
    un.image = k / K; un.first_eval = k; un.count = 1; un.pad_ = 0;
    units[k] = un;
}

// consume the trial evaluation: parabola through f(0), f'(0), f(1); maximiser clamped to [0, 2]
__global__ void lm_select_kernel(OptState *__restrict__ st, int n, const float4 *__restrict__ sc, const OptPrior pr) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    OptState s = st[k];
    const float f1 = cc_of(sc[k]) - prior_pen(pr, s.lam, s.x[3] + s.d[3], s.x[4] + s.d[4]);
    const float slope = s.pad_[0];
    const float c = f1 - s.f - slope;
    float t;
    if (c < 0.f) {
        t = -slope / (2.f * c);
        t = fminf(fmaxf(t, 0.f), 2.f);
    } else
        t = f1 > s.f ? 2.f : 0.f;
    // towards the plain Gauss-Newton step as the predicted gain falls to the rounding noise of the scores
    const float w = slope * slope / (slope * slope + LM_TAU * LM_TAU);
    t = 1.f + w * (t - 1.f);
    for (int m = 0; m < LM_NG; ++m) s.x[m] += t * s.d[m];
    st[k] = s;
}

}  // namespace
