// plan.cu — host-side builders: polar-patch band plan and ring lists.
//
// The band is the set of half-plane lattice points (i in [0,n/2], j in [-n/2,n/2-1]) with
// r_lo <= sqrt(i^2+j^2) <= r_hi, r in Fourier pixels (= n*pixel/resolution), the same set
// cisTEM's weighted correlation loops over (bin index = int(r)); SURVEY.md §8d counts it as
// n_band.  Order is ours: ring-bands of 4 rings, angle-sorted, dummy-padded (see internal.cuh).
#include <algorithm>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "internal.cuh"

int cspb_fail(cspb_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

bool build_band_plan(BandPlan &plan, int n, float r_lo, float r_hi, bool radial_order) {
    plan.n = n;
    plan.r_lo = r_lo;
    plan.r_hi = r_hi;
    plan.slot_ij.clear();
    plan.bands.clear();
    const int nh = n / 2;
    const float lo2 = r_lo * r_lo, hi2 = r_hi * r_hi;
    const int ring_min = (int)floorf(r_lo), ring_max = std::min((int)floorf(r_hi), (int)(nh * 1.5f));
    plan.ring_min = ring_min;
    plan.ring_max = ring_max;
    if (ring_max < ring_min) return false;
    std::vector<std::vector<std::pair<float, int32_t>>> rings(ring_max - ring_min + 1);
    int n_band = 0;
    for (int j = -nh; j < nh; ++j)
        for (int i = 0; i <= nh; ++i) {
            const float r2 = (float)(i * i + j * j);
            if (r2 < lo2 || r2 > hi2) continue;
            const int ring = (int)floorf(sqrtf(r2));
            if (ring < ring_min || ring > ring_max) continue;
            const float ang = atan2f((float)j, (float)i);
            const int32_t packed = (int32_t)((uint32_t)(i & 0xFFFF) | ((uint32_t)(j & 0xFFFF) << 16));
            rings[ring - ring_min].push_back(std::make_pair(ang, packed));
            ++n_band;
        }
    plan.n_band = n_band;
    for (auto &r : rings) std::sort(r.begin(), r.end());
    const int32_t dummy = (int32_t)CSPB_DUMMY_I;  // i = 0x7FFF, j = 0
    const int n_rings = ring_max - ring_min + 1;
    int slot = 0;
    // Bands hold 4 rings of similar sample count: rings are ordered by count (ties by radius) — the
    // lattice ring counts scatter by about +-10 around pi*r, so radially consecutive rings pad each
    // other by 10 % at 256 px, count neighbours by 5 % (what is left is the rounding to 8 angles).
    // radial_order keeps radially consecutive rings together: chosen by the caller when the reference does
    // not fit in L2 (a warp's 8 x 4 patch is then compact in the volume, which is worth more DRAM-side than
    // the padding: 384 px, r01h, 0.281 s vs 0.288 s); CSPB_BAND_ORDER=radial|count forces either (A/B).  If the ring count is not a
    // multiple of 4 the partial band holds the few-sample rings at the head of the order.
    std::vector<int> order(n_rings);
    for (int k = 0; k < n_rings; ++k) order[k] = k;
    const char *env = getenv("CSPB_BAND_ORDER");
    if (env && strcmp(env, "radial") == 0) radial_order = true;
    if (env && strcmp(env, "count") == 0) radial_order = false;
    if (!radial_order)
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rings[a].size() < rings[b].size(); });
    const int rem = n_rings % 4;
    for (int r0 = rem ? rem - 4 : 0; r0 < n_rings; r0 += 4) {
        size_t lmax = 0;
        for (int k = 0; k < 4; ++k)
            if (r0 + k >= 0) lmax = std::max(lmax, rings[order[r0 + k]].size());
        if (lmax == 0) continue;
        const int L = (int)((lmax + 7) / 8) * 8;
        BandDesc bd;
        bd.slot_start = slot;
        bd.n_iter = 4 * L / 32;
        int rid[4];
        for (int k = 0; k < 4; ++k) rid[k] = ring_min + (r0 + k >= 0 ? order[r0 + k] : 0);
        bd.rings01 = rid[0] | (rid[1] << 16);
        bd.rings23 = rid[2] | (rid[3] << 16);
        plan.slot_ij.resize(slot + 4 * L, dummy);
        for (int k = 0; k < 4; ++k) {
            if (r0 + k < 0) continue;
            const auto &r = rings[order[r0 + k]];
            const int nk = (int)r.size();
            for (int m = 0; m < nk; ++m) {
                const int a = (int)(((long long)m * L) / nk);  // spread dummies evenly along the arc
                plan.slot_ij[slot + a * 4 + k] = r[m].second;
            }
        }
        slot += 4 * L;
        plan.bands.push_back(bd);
    }
    plan.n_slots = slot;
    plan.n_bands = (int)plan.bands.size();
    return plan.n_slots > 0;
}
