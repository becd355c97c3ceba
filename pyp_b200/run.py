"""Multi-GPU product entry: one refine3d -> [score shaping] -> reconstruct3d -> merge3d iteration over a whole
stack on N GPUs of one box, one process per GPU.

    python -m pyp_b200.run --gpus 8 --refine refine3d.in --reconstruct reconstruct3d.in --merge merge3d.in

Each `.in` file holds exactly the answers pyp feeds the corresponding binary on stdin (frealign.py:3918-3994,
1780-1824, 2075-2093) with `first`/`last` covering the whole stack: what pyp does with `slurm_tasks` processes per
stage and dump files in between (split_refinement frealign.py:3014-3193, split_reconstruction :1622-1835,
merge_reconstructions :1910-2136) happens here in one launch:

  * rank r refines the contiguous shard `dist.shard_range(first, last, r, N)` of the stack (no communication);
    the whitening curve is estimated once, by rank 0, on the images a single process would use, and broadcast;
  * the refined rows are gathered on rank 0 and written as the ONE range file pyp expects for [first, last]
    (`<name>_<first:07d>_<last:07d>.cistem` + `_changes`), sorted like Parameters.merge;
  * every rank inserts its shard into private half-volume accumulators; one ncclReduce(sum) per half over NVLink
    replaces the `local_merge3d` / `merge3d` file sums; rank 0 writes either the dump pair (reconstruct3d's own
    outputs, when no --merge is given) or merge3d's maps + statistics + log table.

Any stage can be left out (`--refine` only = a multi-GPU refine3d; `--reconstruct` only = a multi-GPU
reconstruct3d reading its `_used.cistem` from disk).  When refine and reconstruct run together and the
reconstruction's parameter file does not exist yet, the rows refined in this launch are inserted directly
(optionally after the score shaping of `--cutoff`, analysis/scores.py:300-761).
Launched without torchrun and with --gpus > 1 it re-executes itself under `python -m torch.distributed.run`.
"""
import argparse
import io
import os
import subprocess
import sys
import time

import numpy as np


def _parse_args(argv):
    ap = argparse.ArgumentParser(prog="python -m pyp_b200.run", description=__doc__.split("\n\n")[0])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--refine", help="file with refine3d's stdin answers (whole range)")
    ap.add_argument("--reconstruct", help="file with reconstruct3d's stdin answers (whole range)")
    ap.add_argument("--merge", help="file with merge3d's stdin answers (its dump seeds are ignored: the sum happens over NVLink)")
    ap.add_argument("--cutoff", type=float, default=None, help="score shaping between the stages: reconstruct_cutoff fraction (e.g. 0.75)")
    ap.add_argument("--keep-dumps", action="store_true", help="with --merge: also write reconstruct3d's dump pair (answers 37/38)")
    # option names must not be prefixes of torchrun's own options (--log-dir, --master-port ...): the relaunch below hands
    # this argv to `python -m torch.distributed.run ... -m pyp_b200.run <argv>` whose parser matches prefixes
    ap.add_argument("--out-log", default=None, help="log file of rank 0 (default: stdout)")
    ap.add_argument("--port", type=int, default=29533, help="rendezvous port of the relaunch under torchrun")
    return ap.parse_args(argv)


def _read(path):
    with open(path) as f:
        return f.read()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    a = _parse_args(argv)
    if not (a.refine or a.reconstruct):
        sys.stderr.write("pyp_b200.run: nothing to do (give --refine and/or --reconstruct)\n")
        return 2
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(a.port), "-m", "pyp_b200.run"] + argv
        return subprocess.call(cmd)
    try:
        return _worker(a)
    except Exception as e:  # the word pyp greps for (particle_cspt.py:812-818)
        sys.stderr.write(f"pyp_b200.run: caught error: {type(e).__name__}: {e}\n")
        return 1


def _worker(a):
    import torch
    import torch.distributed as tdist

    from . import dist
    from .cli import merge3d as m_cli, prompts, reconstruct3d as rc_cli, refine3d as rf_cli
    from .engine import ROW_DTYPE, Engine
    from .formats import cistem, dump, mrc, statistics

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        tdist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
    log = io.StringIO()
    t0 = time.time()
    eng = Engine(local_rank)
    refined = None        # this rank's refined rows (shard)
    shard = None
    timing = {}

    def images_of(path, pos):
        _, data = mrc.read(path, first=int(pos.min()), last=int(pos.max()))
        return np.ascontiguousarray(data[pos - pos.min()])

    # ------------------------------------------------------------------ refine3d over the shard
    if a.refine:
        p = rf_cli.parse(prompts.Answers(_read(a.refine), "refine3d"))
        hdr = mrc.read_header(p["stack"])
        box = hdr["nx"]
        first, last = p["first"], (min(p["last"], hdr["nz"]) if p["last"] > 0 else hdr["nz"])
        if first < 1 or first > last:
            raise ValueError(f"particle range {first}..{last} outside the stack (1..{hdr['nz']})")
        rows_all = cistem.read_parameters(p["parameters"])
        rows_rng = rows_all[rf_cli.select_rows(rows_all, first, last)]
        lo, hi = dist.shard_range(0, rows_rng.size - 1, rank, world)
        rows = rows_rng[lo:hi + 1]
        shard = (lo, hi)
        cfg = rf_cli.build_cfg(p, box)
        if p["use_priors"]:
            cfg.use_priors = 1
            cfg.prior_mean_x, cfg.prior_mean_y, cfg.prior_var_x, cfg.prior_var_y = rf_cli.shift_prior(p, rows_all)
        eng.ensure_reference(cfg, p["reference"], lambda: mrc.read(p["reference"])[1])
        ring_w = None
        if p["use_statistics"] and os.path.exists(p["statistics"]) and os.path.getsize(p["statistics"]) > 0:
            st = statistics.read_statistics(p["statistics"])
            if st.size:
                ring_w = statistics.ring_weights_from_statistics(st, box, p["pixel_size"])
                eng.set_ring_weights(ring_w)
        focus = p["apply_2d_masking"] and p["mask_2d"][3] > 0
        eng.set_focus_mask(*(p["mask_2d"] if focus else (0, 0, 0, 0)))
        eng.set_symmetry(p["symmetry"])
        if p["global_search"]:
            from .search_grid import search_grid

            eng.set_search_grid(search_grid(p["angular_step"], p["symmetry"]))
        # partition-invariant whitening: rank 0 estimates the curve on the images a single process would use (the
        # head of the range) and every rank installs it before loading its own shard
        if cfg.whiten and world > 1:
            curve = torch.zeros(box + 1, dtype=torch.float32)
            if rank == 0 and rows_rng.size:
                head = rows_rng["position_in_stack"][: min(rows_rng.size, 16384)].astype(np.int64)
                eng.load_images(images_of(p["stack"], head))
                curve = torch.from_numpy(eng.noise_curve())
            tdist.broadcast(curve, src=0)
            eng.refine_reset_images()
            if ring_w is not None:
                eng.set_ring_weights(ring_w)
            eng.set_noise_curve(curve.numpy())
        t1 = time.time()
        pos = rows["position_in_stack"].astype(np.int64)
        for s in range(0, rows.size, 16384):
            eng.load_images(images_of(p["stack"], pos[s:s + 16384]), append=s > 0)
        refined, changes, n_evals = eng.refine(rows, want_changes=True) if rows.size else (rows.copy(), rows.copy(), 0)
        timing["refine_s"] = time.time() - t1
        all_rows = dist.gather_rows(refined) if world > 1 else refined[np.argsort(refined["position_in_stack"], kind="stable")]
        all_chg = dist.gather_rows(changes) if world > 1 else changes[np.argsort(changes["position_in_stack"], kind="stable")]
        if world > 1:
            ev = torch.tensor([float(n_evals)], dtype=torch.float64)
            tdist.all_reduce(ev)
            n_evals = int(ev.item())
        if rank == 0:
            if all_rows.size:
                cistem.write_parameters(p["out_parameters"], all_rows)
                cistem.write_parameters(p["out_changes"], all_chg)
            log.write(prompts.banner("Refine3D") + f"\nRefined particles {first} to {last} ({all_rows.size} rows) on {world} GPU(s), box {box}; "
                      f"{n_evals} projections scored, {timing['refine_s']:.2f} s on rank 0\n")
            if all_rows.size:
                log.write(f"Mean score {float(all_rows['score'].mean()):.4f}, mean change {float(all_chg['score'].mean()):+.4f}\n")
            prompts.write_notes(log, "refine3d", rf_cli.ignored_answers(p))
            log.write("\nRefine3D: Normal termination\n")

    # ------------------------------------------------------------------ reconstruct3d over the shard + NVLink sum
    if a.reconstruct:
        p = rc_cli.parse(prompts.Answers(_read(a.reconstruct), "reconstruct3d"))
        hdr = mrc.read_header(p["stack"])
        box = hdr["nx"]
        first, last = p["first"], (min(p["last"], hdr["nz"]) if p["last"] > 0 else hdr["nz"])
        if refined is not None and not os.path.exists(p["parameters"]):
            rows = refined  # the rows of this launch, already sharded
            if a.cutoff is not None:
                # score shaping between the stages (scores.py:300-761, one cluster, SPA windows at their defaults):
                # the threshold is a property of the whole table, so rank 0 shapes the gathered rows and every
                # rank takes the occupancies of its own shard
                from . import select

                gathered = dist.gather_rows(rows) if world > 1 else rows
                total = torch.zeros(1, dtype=torch.int64)
                if rank == 0:
                    total[0] = gathered.size
                if world > 1:
                    tdist.broadcast(total, src=0)
                occ = torch.zeros(int(total[0]), dtype=torch.float32)
                if rank == 0:
                    scfg = Engine.select_defaults(a.cutoff)
                    scfg.renumber = 0
                    if a.cutoff == 0:  # automatic cutoff: the two-Gaussian fit of the score populations stays on the host
                        scfg.threshold_override = 1.075 * select.optimal_threshold(gathered["score"]) if gathered.size > 20 else float("nan")
                    shaped, _ = eng.select_scores(gathered, scfg)  # select.cu: the same decisions as scores.py:300-761
                    occ = torch.from_numpy(np.ascontiguousarray(shaped["occupancy"], dtype=np.float32))
                if world > 1:
                    tdist.broadcast(occ, src=0)
                rows = rows.copy()
                rows["occupancy"] = occ.numpy()[shard[0]:shard[1] + 1]
            used_all = None
        else:
            used_all = cistem.read_parameters(p["parameters"])
            rng_rows = used_all[rf_cli.select_rows(used_all, first, last)]
            lo, hi = dist.shard_range(0, rng_rows.size - 1, rank, world)
            rows = rng_rows[lo:hi + 1].copy()
        # dose weights and the average score are properties of the WHOLE range: computed before sharding matters
        all_for_weights = used_all if used_all is not None else (dist.gather_rows(rows) if world > 1 else rows)
        if world > 1 and used_all is None:  # the weights are a property of the whole table: every rank needs it
            holder = [all_for_weights]
            tdist.broadcast_object_list(holder, src=0)
            all_for_weights = holder[0]
        dw, dw_note = rc_cli.dose_weights(p, rows, all_for_weights, box)
        if dw is not None:
            rows["occupancy"] = np.where(dw[:, 0] > 0, rows["occupancy"], 0)
        used = rows[rows["occupancy"] > 0]
        mom = torch.tensor([float(used["score"].astype(np.float64).sum()), float(used.size), float(rows.size)], dtype=torch.float64)
        if world > 1:
            tdist.all_reduce(mom)
        cfg = Engine.recon_defaults(box, p["pixel_size"])
        cfg.pad = 2 if p["padding"] >= 1.5 else 1
        cfg.mask_radius = p["outer_mask_radius"]
        cfg.resolution_limit = p["resolution_limit"]
        cfg.score_bfactor, cfg.score_weighting, cfg.score_threshold = p["score_bfactor"], int(p["score_weighting"]), p["score_threshold"]
        cfg.normalize, cfg.invert_contrast = int(p["normalize"]), int(p["invert"])
        cfg.per_particle_split = int(p["per_particle_split"])
        cfg.average_score = float(mom[0] / mom[1]) if mom[1] > 0 else 0.0
        eng.set_symmetry(p["symmetry"])
        eng.recon_begin(cfg)
        t1 = time.time()
        pos = rows["position_in_stack"].astype(np.int64)
        for s in range(0, rows.size, 8192):
            e = min(rows.size, s + 8192)
            eng.recon_insert(images_of(p["stack"], pos[s:e]), rows[s:e], None if dw is None else dw[s:e])
        timing["insert_s"] = time.time() - t1
        if world > 1:
            nfloats = eng.recon_dims()[1]
            eng.sync()
            for h in (0, 1):
                dist.reduce_sum(dist.device_tensor(eng.recon_device_ptr(h), nfloats, dev), dst=0)
            torch.cuda.synchronize()
        if rank == 0:
            n_used = int(mom[1])
            log.write(prompts.banner("Reconstruct3D") + f"\nInserted {n_used} of {int(mom[2])} particles ({first}..{last}) on {world} GPU(s), "
                      f"symmetry {p['symmetry']}, box {box}, padding {cfg.pad}\n")
            if dw_note:
                log.write(dw_note + "\n")
            prompts.write_notes(log, "reconstruct3d", rc_cli.ignored_answers(p))
            if a.merge and a.keep_dumps:
                for h, path in ((0, p["dump1"]), (1, p["dump2"])):
                    dump.write(path, eng.recon_get_dump(h), box, cfg.pad, h, p["pixel_size"], n_used)
            if a.merge:
                pm = m_cli.parse(prompts.Answers(_read(a.merge), "merge3d"))
                vol, h1, h2, st = eng.recon_finalize(pm["molecular_mass"], pm["outer_radius"])
                mrc.write(pm["half1"], h1, p["pixel_size"])
                mrc.write(pm["half2"], h2, p["pixel_size"])
                mrc.write(pm["filtered"], vol, p["pixel_size"])
                with open(pm["statistics"], "w") as f:
                    f.write(statistics.HEADER + statistics.format_table(st) + "\n")
                log.write("\nReconstruct3D: Normal termination\n" + prompts.banner("Merge3D") +
                          f"\nMerged the accumulators of {world} GPU(s) over NVLink, {n_used} particles\n\n" + statistics.merge3d_log(st))
            elif p["dump"]:
                for h, path in ((0, p["dump1"]), (1, p["dump2"])):
                    dump.write(path, eng.recon_get_dump(h), box, cfg.pad, h, p["pixel_size"], n_used)
                log.write("\nReconstruct3D: Normal termination\n")
            else:
                vol, h1, h2, st = eng.recon_finalize(p["molecular_mass"], p["outer_mask_radius"])
                mrc.write(p["out_map1"], h1, p["pixel_size"])
                mrc.write(p["out_map2"], h2, p["pixel_size"])
                mrc.write(p["out_filtered"], vol, p["pixel_size"])
                with open(p["out_statistics"], "w") as f:
                    f.write(statistics.HEADER + statistics.format_table(st) + "\n")
                log.write("\nReconstruct3D: Normal termination\n")
        eng.recon_end()
    if world > 1:
        tdist.barrier()
    if rank == 0:
        log.write(f"pyp_b200.run: {world} GPU(s), {time.time() - t0:.2f} s, stages {timing}\n")
        text = log.getvalue()
        if a.out_log:
            with open(a.out_log, "a") as f:
                f.write(text)
        else:
            sys.stdout.write(text)
    eng.close()
    if world > 1:
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
