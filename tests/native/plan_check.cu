// Host-side check of the band plan (pyp_b200/csrc/plan.cu): built and run by tests/test_cpu_plan.py.
// Prints one line per configuration: n n_band real_samples n_slots n_bands bad radial
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>
#include "internal.cuh"

int main(int argc, char **argv) {
    const int cfgs[4][2] = {{128, 135}, {256, 100}, {384, 135}, {512, 100}};
    for (int radial = 0; radial < 2; ++radial)
        for (auto &c : cfgs) {
            const int n = c[0];
            const float px = c[1] / 100.f;
            BandPlan p;
            const bool ok = build_band_plan(p, n, n * px / 100.f, 0.4f * n, radial != 0);
            std::set<int> seen;
            int bad = ok ? 0 : 1, real = 0, covered = 0;
            for (auto &b : p.bands) {
                const int rid[4] = {b.rings01 & 0xFFFF, b.rings01 >> 16, b.rings23 & 0xFFFF, b.rings23 >> 16};
                if (b.slot_start != covered || b.slot_start % 32) ++bad;   // bands tile the slot range in order
                covered += b.n_iter * 32;
                for (int s = 0; s < b.n_iter * 32; ++s) {
                    const int ij = p.slot_ij[b.slot_start + s];
                    const int i = (short)(ij & 0xFFFF), j = (short)(ij >> 16);
                    if (i == CSPB_DUMMY_I) continue;
                    ++real;
                    const float r2 = (float)(i * i + j * j);
                    if ((int)floorf(sqrtf(r2)) != rid[s & 3]) ++bad;       // lane % 4 = the band's ring of that track
                    if (r2 < p.r_lo * p.r_lo || r2 > p.r_hi * p.r_hi || i < 0 || i > n / 2 || j < -n / 2 || j >= n / 2) ++bad;
                    if (!seen.insert(ij).second) ++bad;                    // every sample exactly once
                }
            }
            if (covered != p.n_slots || p.n_slots != (int)p.slot_ij.size()) ++bad;
            printf("%d %d %d %d %d %d %d\n", n, p.n_band, real, p.n_slots, p.n_bands, bad, radial);
        }
    return 0;
}
