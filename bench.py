#!/usr/bin/env python
"""bench.py — CSP refine3d + reconstruct3d throughput on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE.json
configs[1] (C2): SPA, O symmetry, 256-px box at 1.0 A/px, local angular search (refine_mode 1), scoring band
100 A .. 2.5 px:
  preprocess (normalise, FFT, whiten, mask, band-pack) -> batched local refinement (scorer)
  -> reconstruct3d insertion (all symmetry operators) -> [NCCL reduce of the half-volumes] -> merge3d finalise on rank 0.
`--config C1|C3|C4|C5` runs the other BASELINE configs at their real shapes (bounded particle counts per GPU).

  value     scored projections / s (C5: reconstruct3d particles / s) with the inputs resident in HBM
  e2e       the same metric through the public host-buffer C-ABI call (pinned host stack, H2D + D2H inside)
  roofline  the dominant kernel against the unit that binds it, measured live (CUDA events on the engine's
            stream, gather-load census, gather-peak microbenchmark)
  cpu_baseline   the CPU restatement of the same path (oracle/) on the host cores, bounded sample

`--impl reference` times the reference-side CPU implementation alone: real cisTEM/CSP binaries when
oracle/tier_a.py finds them (baseline/_ref, $PYP_DIR/external), else the oracle port — one single-thread
process per host core over contiguous particle ranges, the reference's own layout
(src/pyp/system/local_run.py:507-516).  It never loads libcspb200.so.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

UNIT = "scored projections/s"

# BASELINE.json configs (SURVEY.md §8d "Synthetic inputs"); `particles` = the config's own count,
# `per_gpu` = what one bench step holds per GPU (weak scaling)
CONFIGS = {
    "C1": dict(desc="SPA CSP refine3d+reconstruct3d, 5k synthetic particles, 128-px box, C1 (BASELINE configs[0])",
               kind="spa", box=128, pixel=1.35, sym="C1", particles=5000, per_gpu=5000),
    "C2": dict(desc="SPA apoferritin-like O symmetry, 256-px box at 1.0 A/px, local angular search, refine3d+reconstruct3d (BASELINE configs[1])",
               kind="spa", box=256, pixel=1.0, sym="O", particles=100000, per_gpu=32768),
    "C3": dict(desc="TOMO CSP sub-tomogram refinement, particles x 41 tilts, 128-px box, per-tilt CTF/defocus (BASELINE configs[2])",
               kind="tomo", box=128, pixel=1.35, sym="C1", particles=20000, per_gpu=2000, tilts=41),
    "C4": dict(desc="SPA global search, 384-px box, 20 degree grid, top-20 hits refined, reconstruct3d half-maps (BASELINE configs[3])",
               kind="spa", box=384, pixel=1.35, sym="C1", particles=500000, per_gpu=2048, global_search=True),
    "C5": dict(desc="Large-box reconstruction stress: 512-px box, 2x padded Fourier volume per GPU (BASELINE configs[4])",
               kind="recon", box=512, pixel=1.0, sym="C1", particles=200000, per_gpu=4096, pad=2),
}


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--particles", type=int, default=int(os.environ.get("CSPB_BENCH_PARTICLES", 0)), help="particles per GPU (0 = the config's default)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-size (strong scaling) leg of C2")
    ap.add_argument("--optimizer", default="analytic", choices=["analytic", "stencil"],
                    help="local optimiser: analytic gradient + Gauss-Newton (default, SEMANTICS.md 7c) or the central-difference stencil (7)")
    return ap.parse_args(argv)


# ------------------------------------------------------------------------------ workload definition (both arms)
def refine_params(c, optimizer="analytic"):
    """refine3d configuration of the workload as a plain dict (no library involved): the defaults of
    cspb_refine_cfg_default (tests/test_cpu_host.py checks they agree) with the benchmark's band."""
    box, px = c["box"], c["pixel"]
    mask = (0.25 if c.get("global_search") else 0.38) * box * px
    p = dict(box=box, pad=1, pixel_size=px, mask_radius=mask, low_res_limit=100.0, high_res_limit=2.5 * px, signed_cc_limit=30.0,
             search_mask_radius=1.5 * mask, search_high_res=8.0 * px, angular_step=20.0, best_matches=20,
             search_range_x=0.125 * box * px, search_range_y=0.125 * box * px, defocus_range=500.0, defocus_step=50.0,
             global_search=0, local_refine=1, refine_psi=1, refine_theta=1, refine_phi=1, refine_x=1, refine_y=1, refine_defocus=0,
             apply_mask=1, normalize=1, invert_contrast=0, whiten=1, symmetry_order=1, local_iterations=8,
             use_priors=0, prior_mean_x=0.0, prior_mean_y=0.0, prior_var_x=0.0, prior_var_y=0.0, optimizer=1 if optimizer == "stencil" else 0)
    if c.get("global_search"):  # a compact particle searched to 30 A: refine_dang's default 20 degree grid is ~2x the angular resolution of the search band
        p.update(global_search=1, search_high_res=30.0, search_range_x=20.0, search_range_y=20.0)
    return p


def recon_params(c):
    box, px = c["box"], c["pixel"]
    return dict(box=box, pad=c.get("pad", 1), pixel_size=px, mask_radius=px * box / 2.0, resolution_limit=2.0 * px, score_bfactor=2.0,
                score_weighting=0, score_threshold=0.0, normalize=1, invert_contrast=0, per_particle_split=0, average_score=0.0)


def fill(struct, params):
    for k, v in params.items():
        if hasattr(struct, k):
            setattr(struct, k, v)
    return struct


def band_count(box, pixel, lo_a, hi_a):
    """n_band of SURVEY.md §8d: half-plane lattice points with r_lo <= r <= r_hi (r_hi capped at box/2 - 2)."""
    r_lo, r_hi = box * pixel / lo_a, min(box * pixel / hi_a, box / 2 - 2)
    i = np.arange(0, box // 2 + 1, dtype=np.float32)[None, :]
    j = np.arange(-box // 2, box // 2, dtype=np.float32)[:, None]
    r2 = i * i + j * j
    return int(((r2 >= np.float32(r_lo) ** 2) & (r2 <= np.float32(r_hi) ** 2)).sum())


def recon_band(n):
    """Samples reconstruct3d inserts per projection at Nyquist (i = 0 column counted once)."""
    i = np.arange(0, n // 2 + 1)[None, :]
    j = np.arange(-n // 2, n // 2)[:, None]
    keep = (i * i + j * j <= (n // 2 - 1) ** 2) & ~((i == 0) & (j < 0))
    return int(keep.sum())


def workload_config(a, P):
    """The `config` object of the JSON line — identical in both arms (the sample the CPU arm times is
    described in its cpu_baseline.sample)."""
    c = CONFIGS[a.config]
    rp = refine_params(c)
    cfg = {"workload": c["desc"], "config": a.config, "box": c["box"], "pixel_A": c["pixel"], "symmetry": c["sym"],
           "particles_in_config": c["particles"], "particles_per_gpu": P, "band": "100A..2.5px",
           "n_band": band_count(c["box"], c["pixel"], rp["low_res_limit"], rp["high_res_limit"]),
           "l2_policy": f"inputs larger than L2 ({P * c.get('tilts', 1) * c['box'] ** 2 * 4 / 1e9:.2f} GB stack per GPU per step, read once)"}
    if c["kind"] == "spa":
        stencil = getattr(a, "optimizer", "analytic") == "stencil"
        cfg["search"] = (f"global: 20 deg grid + FFT shift search, top-20 hits refined locally (8 iterations of the {'stencil' if stencil else 'analytic-gradient'} "
                         f"optimiser, {114 if stencil else 18} evaluations per hit)" if c.get("global_search")
                         else ("local: 8 iterations of the central-difference stencil optimiser (114 evaluations per particle)" if getattr(a, "optimizer", "analytic") == "stencil"
                               else "local: 8 coarse-to-fine iterations of the analytic-gradient optimiser (one gradient + one trial evaluation each, 18 evaluations per particle)"))
    elif c["kind"] == "tomo":
        cfg["search"] = f"csp mode 5 (particle angles + shifts), {c['tilts']} tilts per particle, exposure window 0..20, 5 optimiser iterations"
    else:
        cfg["search"] = f"reconstruct3d only, padding {c.get('pad', 1)}"
    return cfg


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx = max(mx, float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ CPU arm (oracle port, persistent workers)
def _cpu_worker_main(conn):
    """One single-thread worker process = one contiguous particle range (local_run.py:507-516).  Holds its
    slice and the prepared reference; every "step" message re-runs preprocess + refine + insert on the slice."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle as O

    state = {}
    while True:
        msg = conn.recv()
        if msg[0] == "init":
            _, kind, vol, stack, rows, extra, rd, cd, curve, sym = msg
            state = dict(kind=kind, stack=stack, rows=rows, extra=extra, curve=curve, sym=sym)
            state["ocfg"] = O.RefineCfg(**rd) if rd else None
            state["occfg"] = O.ReconCfg(**cd) if cd else None
            # the transformed reference is a per-process constant amortised over thousands of particles in production: not timed
            state["ref"] = O.Reference(vol, rd["pad"]) if rd else None
            conn.send(("ready",))
        elif msg[0] == "step":
            s = state
            fast = len(msg) < 2 or msg[1] == "optimised"
            t0 = time.perf_counter()
            n_ev, rows_out = 0, s["rows"]
            specs = O.prepare_images(s["stack"], s["ocfg"], s["curve"], fast=fast) if s["ocfg"] is not None else None
            t1 = time.perf_counter()
            if s["kind"] == "spa":
                if s["ocfg"].global_search:
                    rows_out, n_ev = O.global_search(s["ref"], specs, s["rows"], s["ocfg"], s["extra"]["grid"])
                else:
                    rows_out, n_ev = (O.refine_local_fast if fast else O.refine_local)(s["ref"], specs, s["rows"], s["ocfg"])
            elif s["kind"] == "tomo":
                e = s["extra"]
                rows_out, _, _, n_ev = O.csp_run(s["ref"], specs, s["rows"], e["particles"], e["tilts"], s["ocfg"], O.CspCfg(**e["ccfg"]), e["first"], e["last"])
            t2 = time.perf_counter()
            if s["occfg"] is not None:
                rc = O.Recon(s["occfg"])
                if fast:  # the lattice operators are applied once per reconstruction at merge time: outside the timed region
                    rc.insert_fast(s["stack"], rows_out, s["sym"], finish=False)
                else:
                    rc.insert(s["stack"], rows_out, s["sym"])
            t3 = time.perf_counter()
            if s["occfg"] is not None:
                if fast:
                    rc.discard_fast()
                del rc
            conn.send((n_ev, int(s["rows"].size), t1 - t0, t2 - t1, t3 - t2))
        else:
            conn.close()
            return


class CpuArm:
    """The oracle on `cores` worker processes; data generated and distributed ONCE, `step()` re-times the compute."""

    def __init__(self, a, sample, cores, seed=0):
        import multiprocessing as mp

        from oracle import oracle as O
        from pyp_b200 import synth
        from pyp_b200.symmetry import symmetry_matrices

        c = CONFIGS[a.config]
        box, px, kind = c["box"], c["pixel"], c["kind"]
        self.kind, self.sample, self.cores = kind, sample, cores
        ph = synth.Phantom(box, n_blobs=40, seed=seed, sigma=4.0 if c.get("global_search") else 2.0, radius_frac=0.15 if c.get("global_search") else 0.35)
        vol = ph.volume()
        rd = refine_params(c, a.optimizer) if kind != "recon" else None
        cd = recon_params(c) if kind != "tomo" else None
        mats = symmetry_matrices(c["sym"])
        extra_all = {}
        if kind == "tomo":
            rows, particles, tilts = synth.make_tilt_series(sample, px, seed=1)
            rows = rows.astype(O.ROW_DTYPE)
            start_rows = rows
            ccfg = dict(mode=5, window_min=0, window_max=20, iterations=5, random_evals=0, grid_search=0, angle_step=20.0, shift_step=6.0,
                        tol_particle_psi=30.0, tol_particle_theta=30.0, tol_particle_phi=30.0, tol_particle_shift=20.0, tol_tilt_angle=1.5,
                        tol_tilt_axis=1.0, tol_tilt_shift=100.0, tol_defocus=750.0, seed=0, min_projections=0)
            extra_all = dict(particles=particles.astype(O.PARTICLE_DTYPE), tilts=tilts.astype(O.TILT_DTYPE), ccfg=ccfg)
        else:
            rows = synth.make_rows(sample, px, seed=1).astype(O.ROW_DTYPE)
            start_rows = rows if kind == "recon" else synth.perturb_rows(rows, 2.0, 1.0).astype(O.ROW_DTYPE)
            if c.get("global_search"):
                from pyp_b200.search_grid import search_grid

                start_rows = rows.copy()
                for k in ("psi", "theta", "phi", "x_shift", "y_shift"):
                    start_rows[k] = 0
                extra_all = dict(grid=search_grid(20.0, c["sym"]))
        stack = synth.make_stack(ph, rows, snr=0.05)
        curve = None
        if rd:
            ocfg = O.RefineCfg(**rd)
            curve = O.noise_curve(stack[:256], ocfg)  # a whitening curve from the head of the sample (set-up, not timed)
        units = sample  # particles (tomo: particles, each with all its tilts)
        per = units // cores + (1 if units % cores else 0)
        tilts_n = c.get("tilts", 1) if kind == "tomo" else 1
        self.workers = []
        ctx = mp.get_context("spawn")  # spawn: libgomp is not fork-safe
        for s in range(0, units, per):
            e = min(units, s + per)
            extra = dict(extra_all)
            if kind == "tomo":  # the worker sees the whole tables and refines particles s..e-1 of its own rows
                extra.update(first=s, last=e - 1)
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker_main, args=(child,), daemon=True)
            p.start()
            parent.send(("init", kind, vol, stack[s * tilts_n:e * tilts_n], start_rows[s * tilts_n:e * tilts_n], extra, rd, cd, curve, mats))
            self.workers.append((p, parent))
        for _, conn in self.workers:
            assert conn.recv()[0] == "ready"

    def step(self, variant="optimised"):
        """variant "optimised" = oracle/cspb_oracle_fast.c (production-style CPU code, the arm's value), "naive" = the
        clarity-first restatement oracle/cspb_oracle.c (same algorithm, same results)"""
        t0 = time.perf_counter()
        for _, conn in self.workers:
            conn.send(("step", variant))
        res = [conn.recv() for _, conn in self.workers]
        wall = time.perf_counter() - t0
        evals = sum(r[0] for r in res)
        slow = max(r[2] + r[3] + r[4] for r in res)
        return {"evals": evals, "particles": self.sample, "seconds": slow, "wall_s": wall,
                "prep_s_max": max(r[2] for r in res), "refine_s_max": max(r[3] for r in res), "recon_s_max": max(r[4] for r in res),
                "processes": len(self.workers)}

    def close(self):
        for p, conn in self.workers:
            try:
                conn.send(("quit",))
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()


def cpu_sample_size(a, cores):
    """Bounded sample: ~8-10 s of CPU work per step on `cores` cores (rates of the naive port, r01)."""
    if a.cpu_sample:
        return a.cpu_sample
    c = CONFIGS[a.config]
    per_core = {"C1": 300, "C2": 100, "C3": 6, "C4": 1, "C5": 24}[a.config]
    if c["kind"] == "recon":  # 2 x 8.6 GB of accumulators per worker process at 1024^3: few workers
        return per_core * min(cores, 2)
    return per_core * cores


def cpu_value(a, r):
    c = CONFIGS[a.config]
    if c["kind"] == "recon":
        return r["particles"] / r["seconds"], "particles/s"
    return r["evals"] / r["seconds"], UNIT


def metric_name(a):
    return "reconstruct3d_particles_per_sec" if CONFIGS[a.config]["kind"] == "recon" else "csp_scored_projections_per_sec"


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import tier_a

    c = CONFIGS[a.config]
    P = a.particles or c["per_gpu"]
    cores = os.cpu_count() or 1
    probe = tier_a.probe_report()
    reps = max(1, min(a.steps, 3))      # bounded: the whole run ends within ~2 minutes whatever K is
    warm = 1 if a.warmup > 0 else 0
    t_all = time.perf_counter()
    if probe["tier"] == "A" and c["kind"] == "spa" and not c.get("global_search"):
        line_extra = _run_tier_a(a, probe, cores, reps)
        kind, sample_txt, v, unit, secs = "reference", line_extra.pop("sample"), line_extra.pop("value"), UNIT, line_extra.pop("seconds")
    else:
        if c["kind"] == "recon":
            cores = min(cores, 2)
        sample = cpu_sample_size(a, os.cpu_count() or 1)
        arm = CpuArm(a, sample, cores)
        for _ in range(warm):
            arm.step()
        runs = [arm.step() for _ in range(reps)]
        naive = arm.step("naive")
        arm.close()
        vals = [cpu_value(a, r) for r in runs]
        v, unit = float(np.mean([x[0] for x in vals])), vals[0][1]
        secs = float(np.mean([r["seconds"] for r in runs]))
        kind = "port"
        sample_txt = (f"{sample} particles of the workload ({reps} timed repetitions of {secs:.1f} s after {warm} warm-up, data generated once), "
                      f"one single-thread process per core; optimised CPU code (oracle/cspb_oracle_fast.c: precomputed band list, CTF once per "
                      f"particle, cropped reference with one Friedel flip per sample, lattice symmetry operators applied once to the sums); "
                      f"process start-up and reference FFT not timed; reference binaries: {'LFS stubs only' if probe['stubs'] else 'not found'} "
                      f"({', '.join(probe['searched'])})")
        line_extra = {"reconstruct3d_particles_per_s": float(np.mean([r["particles"] / max(r["recon_s_max"], 1e-9) for r in runs])),
                      "per_core_value": v / cores,
                      "naive_port": {"value": cpu_value(a, naive)[0], "seconds": naive["seconds"],
                                     "what": "oracle/cspb_oracle.c, the clarity-first restatement the parity tests use: same algorithm, same results"}}
    line = {
        "impl": "reference", "metric": metric_name(a), "value": v, "unit": unit, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "timed_repetitions": reps, "ms_per_step": 1e3 * secs, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(a, P),
        "cpu_baseline": dict({"value": v, "unit": unit, "cores": cores, "kind": kind, "sample": sample_txt, "tier_a_probe": probe["tier"]}, **line_extra),
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "total_run_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


def _run_tier_a(a, probe, cores, reps):
    """Real reference binaries present: one pyp-style iteration on a sample of the workload, timed by wall clock
    from the first fork to the last output file (SURVEY.md §8d "CPU baseline timing")."""
    import tempfile

    from oracle import tier_a
    from pyp_b200 import synth

    c = CONFIGS[a.config]
    box, px = c["box"], c["pixel"]
    sample = cpu_sample_size(a, cores)
    ph = synth.Phantom(box, n_blobs=40, seed=0, sigma=2.0)
    rows = synth.make_rows(sample, px, seed=1)
    stack = synth.make_stack(ph, rows, snr=0.05)
    start = synth.perturb_rows(rows, 2.0, 1.0)
    secs = []
    for _ in range(reps):
        with tempfile.TemporaryDirectory() as d:
            r = tier_a.run_iteration(probe["binaries"], d, ph.volume(), stack, start, px, c["sym"], mw=440.0, cores=cores)
            secs.append(r["refine_s"] + r["reconstruct_s"] + r["merge_s"])
    s = float(np.mean(secs))
    # a closed binary does not report its evaluation count: quote its particles/s at our optimiser's evaluations per particle
    epp = 114.0 if a.optimizer == "stencil" else 18.0
    return {"value": epp * sample / s, "seconds": s, "particles_per_s": sample / s,
            "sample": f"{sample} particles through {sorted(probe['binaries'])} as pyp runs them ({cores} concurrent single-thread ranges), wall clock; "
                      f"scored projections/s = particles/s x {epp:.0f} (the evaluation count of our optimiser; the binary reports none)"}


# ------------------------------------------------------------------------------ B200 arm
def run_b200_arm(a):
    import torch
    import torch.distributed as dist

    from pyp_b200 import synth, synth_torch
    from pyp_b200._lib import PARTICLE_DTYPE, ROW_DTYPE, TILT_DTYPE
    from pyp_b200.engine import Engine

    c = CONFIGS[a.config]
    kind = c["kind"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:  # one process per GPU: stay on the GPU's NUMA node (pinned stack, copies)
        from pyp_b200.dist import bind_to_gpu_numa_node

        bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, px = c["box"], c["pixel"]
    P = a.particles or c["per_gpu"]          # particles per GPU (tomo: particles, each with `tilts` projections)
    n_tilt = c.get("tilts", 1) if kind == "tomo" else 1
    eng = Engine(local_rank)
    rcfg = fill(Engine.refine_defaults(n, px), refine_params(c, a.optimizer)) if kind != "recon" else None
    ccfg = fill(Engine.recon_defaults(n, px), recon_params(c)) if kind != "tomo" else None
    n_band = n_slots = 0
    if rcfg is not None:
        eng.refine_configure(rcfg)
        n_band, n_slots = eng.band_counts()
    n_sym = eng.set_symmetry(c["sym"])

    # ---- synthetic workload, generated straight into HBM.  C2 also holds the fixed 100 000-particle job of the
    # config (strong-scaling leg): rank r owns particles [r * Ps, (r + 1) * Ps) of it
    strong = (a.config == "C2" and not a.no_strong)
    Ps = -(-c["particles"] // world) if strong else 0
    P_alloc = max(P, Ps)
    glob = c.get("global_search")
    centres, amps, sigma = synth_torch.symmetric_phantom(n, c["sym"], n_base=9 if c["sym"] != "C1" else 40,
                                                        radius_frac=0.15 if glob else 0.35, sigma=4.0 if glob else 2.0)
    vol = synth_torch.volume(n, centres, amps, sigma, dev)
    tables = None
    if kind == "tomo":
        truth, particles, tilts = synth.make_tilt_series(P_alloc, px, seed=1000 + rank)
        start = truth
        tables = (np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE), np.ascontiguousarray(tilts, dtype=TILT_DTYPE))
    else:
        truth = synth.make_rows(P_alloc, px, seed=1000 + rank)
        if kind == "recon":
            start = truth
        elif glob:
            start = truth.copy()
            for k in ("psi", "theta", "phi", "x_shift", "y_shift"):
                start[k] = 0
        else:
            start = synth.perturb_rows(truth, 2.0, 1.0, seed=2000 + rank)
    n_proj = P * n_tilt
    stack_all = synth_torch.make_stack(n, centres, amps, sigma, truth, snr=0.05, seed=3000 + rank, device=dev)
    stack = stack_all[:n_proj]
    rows_host_all = np.ascontiguousarray(start, dtype=ROW_DTYPE)
    rows_host = rows_host_all[:n_proj]
    rows_init_all = torch.from_numpy(rows_host_all.view(np.uint8).reshape(-1, 128)).to(dev)
    rows_dev_all = rows_init_all.clone()
    torch.cuda.synchronize()
    if rcfg is not None:
        eng.set_reference(vol)
        if glob:
            from pyp_b200.search_grid import search_grid

            eng.set_search_grid(search_grid(20.0, c["sym"]))
    csp_cfg = None
    if kind == "tomo":
        csp_cfg = Engine.csp_defaults(5)
        csp_cfg.window_max, csp_cfg.iterations = 20, 5

    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    out_maps = [torch.empty((n, n, n), device=dev, dtype=torch.float32) for _ in range(3)] if (rank == 0 and ccfg is not None) else None
    stage_events, reduce_events = [], []

    def reduce_volumes(timed):
        """one ncclReduce(sum) per half onto rank 0, bracketed by its own events on torch's stream"""
        nfloats = eng.recon_dims()[1]
        eng.sync()
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_a.record()
        for h in (0, 1):
            t = _wrap_device(torch, eng.recon_device_ptr(h), nfloats, dev)
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        e_b.record()
        torch.cuda.synchronize()
        if timed:
            reduce_events.append((e_a, e_b))

    def device_step(timed=False, count=None):
        """inputs resident in HBM; everything enqueued on the engine stream"""
        cnt = n_proj if count is None else count
        st, rd_ = stack_all[:cnt], rows_dev_all[:cnt]
        evs = []

        def mark():
            if timed:
                e = torch.cuda.Event(enable_timing=True)
                e.record(ext)
                evs.append(e)

        rd_.copy_(rows_init_all[:cnt])
        torch.cuda.current_stream().synchronize()
        mark()
        n_ev = 0
        if rcfg is not None:
            eng.keep_spectra(ccfg is not None)  # the resident stack reaches recon_insert unchanged: one forward FFT per projection
            eng.load_images(st)
        mark()
        if kind == "spa":
            n_ev = eng.refine_device(rd_.data_ptr(), cnt)
        elif kind == "tomo":
            n_ev = eng.csp_run(rows_host_all[:cnt], tables[0][:cnt // n_tilt], tables[1], csp_cfg)[3]
        mark()
        if ccfg is not None:
            eng.recon_begin(ccfg)
            eng.recon_insert(st, rd_.data_ptr())
        mark()
        if ccfg is not None:
            if world > 1:
                reduce_volumes(timed)
            if rank == 0:
                eng.recon_finalize_device(out_maps, molecular_mass_kda=440.0)
        mark()
        if timed:
            stage_events.append(evs)
        return n_ev if kind != "recon" else cnt

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(local_rank)  # started before the warm-up steps, sampled through the timed region
    sampler.start()
    for _ in range(a.warmup):
        device_step()
    barrier()
    eng.profile_enable(True)
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    units = 0
    for _ in range(a.steps):
        units += device_step(timed=True)
    e1.record(ext)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    score_ms, score_launches, score_units = eng.profile_get(0)
    ins_ms, ins_launches, ins_units = eng.profile_get(1)
    eng.profile_enable(False)
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    units_total = sum_over_ranks(units)
    ms_per_step = ms / a.steps
    value = units_total / (ms * 1e-3)
    stage = np.zeros(4)
    for evs in stage_events:
        for k in range(4):
            stage[k] += evs[k].elapsed_time(evs[k + 1])
    stage /= max(1, len(stage_events))  # ms per step: prep, refine, insert, reduce+finalise
    reduce_ms = float(np.mean([x.elapsed_time(y) for x, y in reduce_events])) if reduce_events else 0.0

    # ---- gather-load census: one untimed step with the counting launches on (scorer configs only)
    loads_per_eval = slots_per_eval = None
    if rcfg is not None:
        eng.count_loads(True)
        device_step()
        eng.sync()
        q, sl, ev_c = eng.loads()
        eng.count_loads(False)
        loads_per_eval, slots_per_eval = q / max(1, ev_c), sl / max(1, ev_c)

    # ---- the two local optimisers on the same resident stack (untimed for `value`): final mean score, evaluations and
    # refinement time of each — the algorithmic gain of the analytic optimiser at a matched final score
    opt_cmp = None
    if kind == "spa" and not glob and rank == 0:
        opt_cmp = {}
        for name, code in (("analytic", 0), ("stencil", 1)):
            cfg2 = fill(Engine.refine_defaults(n, px), refine_params(c, name))
            eng.refine_configure(cfg2)
            eng.set_reference(vol)
            eng.keep_spectra(False)  # refinement only in this leg
            eng.load_images(stack)
            for timed_pass in (False, True):  # the first pass sizes the optimiser's work buffers (cudaMalloc), the second is timed
                rows_dev_all[:n_proj].copy_(rows_init_all[:n_proj])
                torch.cuda.synchronize()
                t_a, t_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t_a.record(ext)
                n_ev = eng.refine_device(rows_dev_all.data_ptr(), n_proj)
                t_b.record(ext)
                eng.sync()
            out_rows = rows_dev_all[:n_proj].cpu().numpy().view(ROW_DTYPE).reshape(-1)
            opt_cmp[name] = {"mean_final_score": float(out_rows["score"].mean()), "evals_per_particle": n_ev / n_proj,
                             "refine_ms": t_a.elapsed_time(t_b), "particles_per_s_refine_only": n_proj / (t_a.elapsed_time(t_b) * 1e-3)}
        eng.refine_configure(rcfg)
        eng.set_reference(vol)

    # ---- strong-scaling leg: the config's fixed 100 000-particle job over `world` ranks
    strong_line = None
    if strong:
        device_step(count=Ps)
        barrier()
        k3 = max(1, min(a.steps, 3))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(ext)
        ev_s = 0
        for _ in range(k3):
            ev_s += device_step(count=Ps)
        s1.record(ext)
        barrier()
        ms_s = max_over_ranks(s0.elapsed_time(s1))
        ev_s = sum_over_ranks(ev_s)
        strong_line = {"particles_total": Ps * world, "particles_per_gpu": Ps, "steps": k3, "ms_per_step": ms_s / k3, "value": ev_s / (ms_s * 1e-3),
                       "unit": UNIT, "scaling": "strong"}

    # ---- end to end through the host-buffer API (pinned stack, H2D + D2H inside the timed region)
    e2e = None
    if not a.no_e2e:
        pinned = True
        try:
            host_stack = torch.empty((n_proj, n, n), dtype=torch.float32, pin_memory=True)
        except RuntimeError:  # no room for a pinned stack on this host: pageable memory (slower copies), said in the JSON
            pinned = False
            host_stack = torch.empty((n_proj, n, n), dtype=torch.float32)
        host_stack.copy_(stack)
        torch.cuda.synchronize()
        hs = host_stack.numpy()
        host_maps = [torch.empty((n, n, n), dtype=torch.float32, pin_memory=True).numpy() for _ in range(3)] if (rank == 0 and ccfg is not None) else None

        def host_step():
            # the public host-buffer calls: one upload per projection, copies overlapped with compute
            n_ev = 0
            if ccfg is not None:
                eng.recon_begin(ccfg)
            if kind == "spa":
                _, n_ev = eng.refine_reconstruct(hs, rows_host)
            elif kind == "tomo":
                eng.load_images(hs)
                n_ev = eng.csp_run(rows_host, tables[0][:P], tables[1], csp_cfg)[3]
            else:
                eng.recon_insert(hs, rows_host)
                n_ev = n_proj
            if ccfg is not None:
                if world > 1:
                    reduce_volumes(False)
                if rank == 0:
                    eng.recon_finalize(molecular_mass_kda=440.0, want_halves=True, out=host_maps)
            return n_ev

        sampler2 = ClockSampler(local_rank)  # started before the warm-up call: spawning nvidia-smi stalls the driver briefly
        sampler2.start()
        for _ in range(2):  # two untimed calls: the second one has been seen 15 % slower than the steady state (r02zt)
            host_step()
        barrier()
        k2 = max(1, min(a.steps, 5))
        ev2, per_step = 0, []
        t0 = time.perf_counter()
        for _ in range(k2):
            t1 = time.perf_counter()
            ev2 += host_step()
            per_step.append(1e3 * (time.perf_counter() - t1))
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        ev2 = sum_over_ranks(ev2)
        clocks2 = sampler2.stop()
        h2d = n_proj * n * n * 4 + n_proj * 128        # every projection is uploaded once
        d2h = n_proj * 128 + (3 * n * n * n * 4 if (rank == 0 and ccfg is not None) else 0)
        e2e = {"value": ev2 / dt, "unit": UNIT if kind != "recon" else "particles/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": k2, "ms_per_step": 1e3 * dt / k2, "ms_each_step_rank0": per_step, "clocks": clocks2,
               "host_memory": "pinned" if pinned else "pageable", "h2d_gbs_per_gpu_if_copy_bound": h2d / (dt / k2) / 1e9}
        del host_stack

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    roof = {}
    if rcfg is not None:
        # the scorer's own ceiling: random 32-byte gathers (cspb_gather_peak) from a window the size of the reference
        # half-sphere the band touches (L2-resident up to ~100 MB, HBM beyond) and from an L1-resident window
        rc_q = math.ceil(rcfg.pad * min(n * px / rcfg.high_res_limit, n / 2 - 2)) + 2
        touched = int((2.0 / 3.0) * math.pi * rc_q ** 3 * 32)
        try:
            g_ref, g_l1 = eng.gather_peak(max(touched, 16 << 20)), eng.gather_peak(32 << 10, per_cta=True)
        except Exception:
            g_ref = g_l1 = None
        traffic = l1_frac = None
        try:  # DRAM bytes per launch of the scoring kernel and its L1 data-pipe utilisation, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = float(tj["score_dram_bytes_per_eval"]) * score_units / max(1, score_launches) if a.config == "C2" else None
            l1_frac = tj.get("score_l1_data_pipe_frac") if a.config == "C2" else None
        except Exception:
            pass
        alg_bytes = 72.0 * n_band                                   # 64 B gather + 8 B image per band sample
        alg_gbs = alg_bytes * score_units / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
        loaded_bytes = 32.0 * loads_per_eval + 8.0 * slots_per_eval  # per evaluation: reference quads really loaded + the packed image slots read
        loaded_gbs = loaded_bytes * score_units / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
        # the binding unit in both regimes is the SM's L1 data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 72-80 % for the plain
        # scorer at 256 px AND at 384 px where the reference exceeds the L2 — profiles/r02_notes.md): the ceiling is the
        # L1-resident gather peak; the peak from a window of the reference's size is the random-access (no locality) figure
        roof = {"bound": "l1_data_pipe", "kernel": "score_grad_kernel + score_kernel<PB,DDEF,MODE> (all scorer launches of the step)",
                "achieved": loaded_gbs, "peak": g_l1, "unit": "GB/s", "frac": (loaded_gbs / g_l1) if g_l1 else None,
                "traffic": traffic,
                "peak_source": "live cspb_gather_peak: random 32-byte gathers by 32 warps/SM from an L1-resident window (one 32-byte wavefront per clock and SM)",
                "gather_peak_reference_sized_window": g_ref, "reference_touched_mb": touched / 1e6,
                "achieved_source": "live: (32 B x quad loads counted by the census launch + 8 B x band slots) per evaluation x evaluations / summed CUDA-event time of the scorer launches",
                "loaded_bytes_per_unit": loaded_bytes, "quad_loads_per_slot_read": loads_per_eval / max(1.0, slots_per_eval),
                "band_fraction_per_evaluation": slots_per_eval / max(1, n_slots),
                "frac_of_reference_window_peak": (loaded_gbs / g_ref) if g_ref else None,
                "l1_data_pipe_frac_ncu": l1_frac,
                "algorithmic": {"bytes_per_unit": alg_bytes, "gbs": alg_gbs, "x_hbm_peak": alg_gbs / peak,
                                "note": "SURVEY.md §8d counts 64 B of gather per band sample and pose; the kernel loads less (poses of a unit share voxels, shift evaluations share one gather) and those bytes are served by L1/L2, not HBM"},
                "hbm": {"peak": peak, "peak_source": peak_src, "dram_frac": (traffic / (score_ms / max(1, score_launches) * 1e-3) / 1e9 / peak) if traffic else None},
                "units_per_launch": score_units / max(1, score_launches), "avg_launch_ms": score_ms / max(1, score_launches),
                "share_of_step": (score_ms / a.steps) / ms_per_step}
    ins_roof = None
    if ccfg is not None:
        ins_bytes = 96.0 * (ccfg.pad ** 2) * recon_band(n)  # per projection per literally inserted operator
        ins_gbs = (ins_bytes * ins_units) / (ins_ms * 1e-3) / 1e9 if ins_ms > 0 else 0.0
        ins_roof = {"bound": "hbm", "kernel": "insert_kernel", "achieved": ins_gbs, "peak": peak, "unit": "GB/s", "frac": ins_gbs / peak,
                    "traffic": None, "peak_source": peak_src, "bytes_per_unit": ins_bytes, "avg_launch_ms": ins_ms / max(1, ins_launches),
                    "share_of_step": (ins_ms / a.steps) / ms_per_step,
                    "note": "achieved = algorithmic bytes (96 pad^2 n_band per projection and operator) / kernel time; the atomics are served by L2 while the touched half-sphere fits"}
    unit = "particles/s" if kind == "recon" else UNIT
    line = {
        "metric": metric_name(a), "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a, P),
        "details": {"symmetry_operators": n_sym, "scorer_wave_units": eng.wave_units if rcfg is not None else None, "n_slots": n_slots,
                    "evals_per_particle": (units / (a.steps * P)) if kind != "recon" else None,
                    "parallelism": f"particle shards x{world}, NCCL reduce of half-volumes" if world > 1 else "single GPU"},
        "particles_per_sec": world * P / (ms_per_step * 1e-3),
        "stage_ms_per_step": {"preprocess": stage[0], "refine": stage[1], "insert": stage[2], "reduce_finalize": stage[3],
                              "nccl_reduce": reduce_ms, "score_kernels": score_ms / a.steps, "insert_kernel": ins_ms / a.steps},
        "roofline": roof if roof else ins_roof,
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if ccfg is not None:
        line["reconstruct3d_particles_per_sec"] = world * P / max((stage[2] + stage[3]) * 1e-3, 1e-9)
    if roof and ins_roof:
        line["roofline_insert"] = ins_roof
    if strong_line:
        line["strong_scaling"] = strong_line
    if opt_cmp:
        line["optimizers"] = opt_cmp
    if e2e:
        # how much of the device-resident rate survives the host-to-device copy; the copy-bound ceiling is
        # h2d_bytes / the pinned H2D bandwidth of the platform at this N (tools/h2d_scaling.py)
        e2e["e2e_efficiency"] = ms_per_step / e2e["ms_per_step"]
        line["e2e"] = e2e
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        if kind == "recon":
            cores = min(cores, 2)
        sample = cpu_sample_size(a, os.cpu_count() or 1)
        arm = CpuArm(a, sample, cores)
        r = arm.step()
        arm.close()
        v, u = cpu_value(a, r)
        line["cpu_baseline"] = {"value": v, "unit": u, "cores": cores, "kind": "port",
                                "sample": f"{sample} particles of the same workload, one single-thread process per core, {r['seconds']:.1f} s (optimised CPU code, oracle/cspb_oracle_fast.c; start-up and reference FFT not timed; `--impl reference` also times the naive port)",
                                "per_core_value": v / cores,
                                "reconstruct3d_particles_per_s": (r["particles"] / r["recon_s_max"]) if r["recon_s_max"] > 0 else None}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _wrap_device(torch, ptr, nfloats, dev):
    """torch view of an engine-owned device buffer (plumbing for NCCL)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(nfloats),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=dev)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
