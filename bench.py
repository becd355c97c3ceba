#!/usr/bin/env python
"""bench.py — CSP refine3d + reconstruct3d throughput on B200 (BASELINE.json metric).

One "step" = one refinement iteration of the hot path over one resident synthetic stack:
  preprocess (normalise, FFT, whiten, mask, band-pack) -> batched local refinement (scorer)
  -> reconstruct3d insertion (all symmetry operators) -> [NCCL reduce of the half-volumes]
  -> merge3d finalise on rank 0.
Workload = BASELINE.json configs[1]: SPA, O symmetry, 256-px box at 1.0 A/px, local angular
search (refine_mode 1), scoring band 100 A .. 2.5 A; `--particles` per GPU (weak scaling).

  value   scored projections / s   (objective evaluations of the whole job / step time)
  e2e     same metric through the public host-buffer API (pinned host stack, H2D inside)
  extras  reconstruct3d particles/s, per-stage times, roofline of the scoring kernel

`--impl reference` times the CPU restatement of the same path (oracle/, one single-thread
process per host core over contiguous particle ranges — the reference's own layout,
src/pyp/system/local_run.py:507-516) on a bounded sample of the same workload.
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

METRIC = "csp_scored_projections_per_sec"
UNIT = "scored projections/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=int(os.environ.get("CSPB_BENCH_PARTICLES", 32768)), help="particles per GPU")
    ap.add_argument("--box", type=int, default=256)
    ap.add_argument("--pixel", type=float, default=1.0)
    ap.add_argument("--sym", default="O")
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_cfgs(box, pixel):
    from pyp_b200.engine import Engine

    rcfg = Engine.refine_defaults(box, pixel)
    rcfg.low_res_limit = 100.0            # refine_rlref default
    rcfg.high_res_limit = 2.5 * pixel     # fixed benchmark band, SURVEY.md §8d
    rcfg.mask_radius = 0.38 * box * pixel
    rcfg.local_iterations = 8
    ccfg = Engine.recon_defaults(box, pixel)
    return rcfg, ccfg


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


# ------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One single-thread process over a contiguous particle range (local_run.py:507-516)."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle as O

    vol, stack, rows, rcfg_d, ccfg_d, curve, sym = args
    ocfg = O.RefineCfg(**rcfg_d)
    occfg = O.ReconCfg(**ccfg_d)
    ref = O.Reference(vol, ocfg.pad)  # per-process constant, amortised over thousands of particles in production: not timed
    t0 = time.perf_counter()
    specs = O.prepare_images(stack, ocfg, curve)
    t1 = time.perf_counter()
    out, n_ev = O.refine_local(ref, specs, rows, ocfg)
    t2 = time.perf_counter()
    rc = O.Recon(occfg)
    rc.insert(stack, out, sym)
    t3 = time.perf_counter()
    return n_ev, rows.size, t1 - t0, t2 - t1, t3 - t2


def cpu_reference_run(box, pixel, sym, sample, cores, seed=0):
    """Times the oracle on `sample` particles of the workload with `cores` processes.
    Returns dict(evals_per_s, particles_per_s, seconds, ...)."""
    import multiprocessing as mp

    from oracle import oracle as O
    from pyp_b200 import synth
    from pyp_b200.symmetry import symmetry_matrices

    rcfg, ccfg = workload_cfgs(box, pixel)
    ph = synth.Phantom(box, n_blobs=40, seed=seed, sigma=2.0)
    vol = ph.volume()
    rows = synth.make_rows(sample, pixel, seed=1).astype(O.ROW_DTYPE)
    stack = synth.make_stack(ph, rows, snr=0.05)
    start = synth.perturb_rows(rows, 2.0, 1.0).astype(O.ROW_DTYPE)
    ocfg = O.refine_cfg_from(rcfg)
    curve = O.noise_curve(stack, ocfg)
    mats = symmetry_matrices(sym)
    rd = {k: getattr(ocfg, k) for k, _ in O.RefineCfg._fields_}
    cd = {k: getattr(O.recon_cfg_from(ccfg), k) for k, _ in O.ReconCfg._fields_}
    inc = int(math.ceil(sample / cores))
    jobs = []
    for s in range(0, sample, inc):
        jobs.append((vol, stack[s:s + inc], start[s:s + inc], rd, cd, curve, mats))
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(len(jobs)) as pool:  # spawn: libgomp is not fork-safe
        res = pool.map(_cpu_worker, jobs)
    spawn_wall = time.perf_counter() - t0
    wall = max(r[2] + r[3] + r[4] for r in res)  # slowest worker's compute time (process start-up excluded)
    evals = sum(r[0] for r in res)
    return {
        "evals": evals, "particles": sample, "seconds": wall,
        "evals_per_s": evals / wall, "particles_per_s": sample / wall,
        "refine_s_max": max(r[3] for r in res), "recon_s_max": max(r[4] for r in res), "prep_s_max": max(r[2] for r in res),
        "processes": len(jobs), "wall_with_spawn_s": spawn_wall,
    }


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = a.cpu_sample or 48 * cores  # ~10 s of CPU work per step
    rcfg, _ = workload_cfgs(a.box, a.pixel)
    for _ in range(max(0, min(a.warmup, 1))):
        cpu_reference_run(a.box, a.pixel, a.sym, cores, cores)
    runs = [cpu_reference_run(a.box, a.pixel, a.sym, sample, cores) for _ in range(max(1, a.steps))]
    t = float(np.mean([r["seconds"] for r in runs]))
    v = float(np.mean([r["evals_per_s"] for r in runs]))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SPA local refine3d+reconstruct3d, {a.sym} symmetry, {a.box}-px box at {a.pixel} A/px (BASELINE configs[1])",
                   "sample_particles": sample, "band": "100A..2.5px"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} particles of the workload, one single-thread process per core (oracle/cspb_oracle.c; reference binaries are source-less LFS stubs)",
                         "reconstruct3d_particles_per_s": float(np.mean([r["particles"] / r["recon_s_max"] for r in runs]))},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ B200 arm
def run_b200_arm(a):
    import torch
    import torch.distributed as dist

    from pyp_b200 import synth, synth_torch
    from pyp_b200._lib import ROW_DTYPE
    from pyp_b200.engine import Engine

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

    n, px = a.box, a.pixel
    rcfg, ccfg = workload_cfgs(n, px)
    eng = Engine(local_rank)
    P = a.particles  # (a wave-aligned count, eng.wave_units x k, was measured: no gain — CTAs do not finish in lockstep)
    eng.refine_configure(rcfg)
    n_sym = eng.set_symmetry(a.sym)
    n_band, n_slots = eng.band_counts()

    # ---- synthetic workload, generated straight into HBM
    centres, amps, sigma = synth_torch.symmetric_phantom(n, a.sym)
    vol = synth_torch.volume(n, centres, amps, sigma, dev)
    truth = synth.make_rows(P, px, seed=1000 + rank)
    start = synth.perturb_rows(truth, 2.0, 1.0, seed=2000 + rank)
    stack = synth_torch.make_stack(n, centres, amps, sigma, truth, snr=0.05, seed=3000 + rank, device=dev)
    rows_host = np.ascontiguousarray(start, dtype=ROW_DTYPE)
    rows_init = torch.from_numpy(rows_host.view(np.uint8).reshape(P, 128)).to(dev)
    rows_dev = rows_init.clone()
    torch.cuda.synchronize()
    eng.set_reference(vol)

    ext = torch.cuda.ExternalStream(eng.stream, device=dev)

    out_maps = [torch.empty((n, n, n), device=dev, dtype=torch.float32) for _ in range(3)] if rank == 0 else None
    stage_events = []

    def device_step(timed=False):
        """inputs resident in HBM; everything enqueued on the engine stream"""
        evs = []

        def mark():
            if timed:
                e = torch.cuda.Event(enable_timing=True)
                e.record(ext)
                evs.append(e)

        rows_dev.copy_(rows_init)
        torch.cuda.current_stream().synchronize()
        mark()
        eng.load_images(stack)
        mark()
        n_ev = eng.refine_device(rows_dev.data_ptr(), P)
        mark()
        eng.recon_begin(ccfg)
        eng.recon_insert(stack, rows_dev.data_ptr())
        mark()
        if world > 1:
            nfloats = eng.recon_dims()[1]
            eng.sync()
            for h in (0, 1):
                t = _wrap_device(torch, eng.recon_device_ptr(h), nfloats, dev)
                dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
        if rank == 0:
            eng.recon_finalize_device(out_maps, molecular_mass_kda=440.0)
        mark()
        if timed:
            stage_events.append(evs)
        return n_ev

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)  # started before the warm-up steps (see above), sampled through the timed region
    sampler.start()
    for _ in range(a.warmup):
        device_step()
    barrier()
    eng.profile_enable(True)
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    evals = 0
    for _ in range(a.steps):
        evals += device_step(timed=True)
    e1.record(ext)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    score_ms, score_launches, score_units = eng.profile_get(0)
    ins_ms, ins_launches, ins_units = eng.profile_get(1)
    eng.profile_enable(False)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ev = torch.tensor([float(evals)], device=dev, dtype=torch.float64)
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
        evals_total = float(ev.item())
    else:
        evals_total = float(evals)
    ms_per_step = ms / a.steps
    value = evals_total / (ms * 1e-3)
    stage = np.zeros(4)
    for evs in stage_events:
        for k in range(4):
            stage[k] += evs[k].elapsed_time(evs[k + 1])
    stage /= max(1, len(stage_events))  # ms per step: prep, refine, insert, reduce+finalise

    # ---- end to end through the host-buffer API (pinned stack, H2D + D2H inside the timed region)
    e2e = None
    if not a.no_e2e:
        pinned = True
        try:
            host_stack = torch.empty((P, n, n), dtype=torch.float32, pin_memory=True)
        except RuntimeError:  # no room for a pinned stack on this host: pageable memory (slower copies), said in the JSON
            pinned = False
            host_stack = torch.empty((P, n, n), dtype=torch.float32)
        host_stack.copy_(stack)
        torch.cuda.synchronize()
        hs = host_stack.numpy()
        host_maps = [torch.empty((n, n, n), dtype=torch.float32, pin_memory=True).numpy() for _ in range(3)] if rank == 0 else None

        def host_step():
            # the public host-buffer call: one upload per projection, copies overlapped with compute
            tt = [time.perf_counter()]
            eng.recon_begin(ccfg)
            out, n_ev = eng.refine_reconstruct(hs, rows_host)
            tt.append(time.perf_counter())
            if world > 1:
                eng.sync()
                for h in (0, 1):
                    t = _wrap_device(torch, eng.recon_device_ptr(h), eng.recon_dims()[1], dev)
                    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
                torch.cuda.synchronize()
            if rank == 0:
                eng.recon_finalize(molecular_mass_kda=440.0, want_halves=True, out=host_maps)
            tt.append(time.perf_counter())
            if os.environ.get("CSPB_BENCH_DEBUG"):
                sys.stderr.write(f"e2e: refine_reconstruct {1e3 * (tt[1] - tt[0]):.1f} ms, reduce+finalize {1e3 * (tt[2] - tt[1]):.1f} ms\n")
            return n_ev

        sampler2 = ClockSampler(local_rank)  # started before the warm-up call: spawning nvidia-smi stalls the driver briefly
        sampler2.start()
        host_step()
        barrier()
        k2 = max(1, min(a.steps, 3))
        ev2, per_step = 0, []
        t0 = time.perf_counter()
        for _ in range(k2):
            t1 = time.perf_counter()
            ev2 += host_step()
            per_step.append(1e3 * (time.perf_counter() - t1))
        barrier()
        dt = time.perf_counter() - t0
        clocks2 = sampler2.stop()
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = P * n * n * 4 + P * 128                  # every projection is uploaded once (cspb_refine_reconstruct)
        d2h = P * 128 + (3 * n * n * n * 4 if rank == 0 else 0)
        e2e = {"value": world * ev2 / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": k2, "ms_per_step": 1e3 * dt / k2, "ms_each_step_rank0": per_step, "clocks": clocks2,
               "host_memory": "pinned" if pinned else "pageable"}
        del host_stack

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    # gather roofline of the scoring kernel: random 32-byte gathers, L2- and L1-resident
    try:
        gather_l2, gather_l1 = eng.gather_peak(64 << 20), eng.gather_peak(32 << 10, per_cta=True)
    except Exception:
        gather_l2 = gather_l1 = None
    traffic = l1_frac = None
    try:  # DRAM bytes per launch of the scoring kernel, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic = float(tj["score_dram_bytes_per_eval"]) * score_units / max(1, score_launches)
        l1_frac = tj.get("score_l1_data_pipe_frac")
    except Exception:
        traffic = l1_frac = None
    bytes_per_eval = 72.0 * n_band                       # 64 B gather + 8 B image per band sample
    score_gbs = (bytes_per_eval * score_units) / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
    ins_bytes = 96.0 * (ccfg.pad ** 2) * _recon_band(n)  # per projection per literally inserted operator
    ins_gbs = (ins_bytes * ins_units) / (ins_ms * 1e-3) / 1e9 if ins_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SPA local refine3d+reconstruct3d, {a.sym} symmetry ({n_sym} ops), {n}-px box at {px} A/px (BASELINE configs[1])",
                   "particles_per_gpu": P, "scorer_wave_units": eng.wave_units, "band": "100A..2.5px", "n_band": n_band, "n_slots": n_slots,
                   "evals_per_particle": evals / (a.steps * P), "l2_policy": f"inputs larger than L2 ({P * n * n * 4 / 1e9:.1f} GB stack per step)",
                   "parallelism": f"particle shards x{world}, NCCL reduce of half-volumes" if world > 1 else "single GPU"},
        "reconstruct3d_particles_per_sec": world * P / ((stage[2] + stage[3]) * 1e-3),
        "refine3d_scored_projections_per_sec": evals_total / a.steps / ((stage[0] + stage[1]) * 1e-3),
        "stage_ms_per_step": {"preprocess": stage[0], "refine": stage[1], "insert": stage[2], "reduce_finalize": stage[3],
                              "score_kernels": score_ms / a.steps, "insert_kernel": ins_ms / a.steps},
        "roofline": {"bound": "hbm", "kernel": "score_kernel<4,false,2>", "achieved": score_gbs, "peak": peak, "unit": "GB/s",
                     "frac": score_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                     "note": "gather served by L1/L2 (reference volume L2-resident): algorithmic bytes exceed the HBM peak; the kernel's own ceiling is the SM data pipe, measured live as gather_peak_* (random 32-byte gathers); gather_achieved counts the 64 B/sample of the algorithm, of which the kernel really loads about half (neighbouring poses reuse quads, shift evaluations share one gather)",
                     "gather_peak_l2_resident": gather_l2, "gather_peak_l1_resident": gather_l1,
                     "gather_achieved": score_gbs * 64.0 / 72.0, "gather_frac": (score_gbs * 64.0 / 72.0 / gather_l2) if gather_l2 else None,
                     "l1_data_pipe_frac_ncu": l1_frac,  # the binding unit of this kernel, from the committed ncu capture
                     "bytes_per_unit": bytes_per_eval, "units_per_launch": score_units / max(1, score_launches),
                     "avg_launch_ms": score_ms / max(1, score_launches)},
        "roofline_insert": {"bound": "hbm", "kernel": "insert_kernel", "achieved": ins_gbs, "peak": peak, "unit": "GB/s",
                            "frac": ins_gbs / peak, "bytes_per_unit": ins_bytes, "avg_launch_ms": ins_ms / max(1, ins_launches)},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = a.cpu_sample or 48 * cores  # ~10 s of CPU work per step
        r = cpu_reference_run(n, px, a.sym, sample, cores)
        line["cpu_baseline"] = {"value": r["evals_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{sample} particles of the same workload, one single-thread process per core, {r['seconds']:.1f} s",
                                "reconstruct3d_particles_per_s": r["particles"] / r["recon_s_max"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _recon_band(n):
    c = 0
    for j in range(-n // 2, n // 2):
        for i in range(0, n // 2 + 1):
            if i == 0 and j < 0:
                continue
            if i * i + j * j <= (n // 2 - 1) ** 2:
                c += 1
    return c


def _wrap_device(torch, ptr, nfloats, dev):
    """torch view of an engine-owned device buffer (plumbing for NCCL)."""
    import ctypes

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
