"""Golden fixtures for the stdin contracts: the command strings the REFERENCE itself assembles for
`refine3d` and `refine_ctf` (src/pyp/refine/frealign/frealign.py:3771-4043 `mrefine_version`), for a local
refinement, a global search with focus mask / priors, and the beam-tilt variant.  The heredoc between
`<< eot` and `eot` is what our front-ends must parse.  Run in the build container only:

    python tests/golden/make_golden_prompts.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402,F401

os.environ["PYP_DIR"] = "/root/reference"
from pyp.refine.frealign import frealign as F  # noqa: E402
from pyp.system import project_params  # noqa: E402


class Loud(dict):
    """A parameter table that names the key the reference asks for when it is missing."""

    def __missing__(self, key):
        raise KeyError(f"reference asked for parameter {key!r}")


def base_mp():
    return Loud(
        data_set="T20S", refine_dataset="T20S", scope_pixel=1.35, data_bin=1, extract_bin=1, scope_mag=10000, scope_voltage=300.0,
        scope_cs=2.7, scope_wgh=0.07, reconstruct_rrec="0", refine_mask="1,1,1,1,1", class_num=1, refine_global_stat=False,
        refine_fssnr=True, refine_priors=False, class_focusmask="0,0,0,0", refine_fdef="F", csp_ToleranceMicrographDefocus1=2000,
        refine_fmatch="F", refine_invert=False, refine_mode="1", refine_srad=0, particle_rad=80.0, refine_fboost=False,
        refine_fboostlim=25.0, extract_fmt="frealign", data_mode="spr", reconstruct_norm=True, particle_sym="O", particle_mw=700.0,
        refine_rlref="100.0", refine_rhref="8.0", class_rhcls="8.0", refine_dang="20.0", refine_searchx="0", refine_searchy="0",
        refine_iblow="1", refine_beamtilt=False, refine_metric="cc3m", refine_iter=2, refine_maxiter=8, refine_skip=False)


def heredoc(cmd):
    body = cmd.split("<< eot", 1)[1]
    body = body.split("\n", 1)[1]
    return body.rsplit("eot", 1)[0]


def main():
    cases = {
        "local": {},
        "global_focus_priors": dict(refine_mode="4", refine_priors=True, class_focusmask="120.5,98.0,77.25,45.0", refine_mask="1,0,1,1,0",
                                    refine_fboost=True, refine_srad=110.0, refine_fdef="T", refine_invert=True),
        "beamtilt": dict(_beam=True),
    }
    out = {}
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    os.environ["PYP_SCRATCH"] = tmp
    open("statistics_r01.txt", "w").write("C header\n   1.00000   50.00000   0.90000   0.99000   0.98000   30.00000   2.00000\n")
    try:
        for tag, kw in cases.items():
            mp = base_mp()
            kw = dict(kw)
            beam = kw.pop("_beam", False)
            mp.update(kw)
            F.project_params.load_pyp_parameters = lambda path=".", _mp=mp: _mp
            cmd = F.mrefine_version(mp, 1, 100, 3, 1, tmp, "T20S_r01", "0000001_0000100", "refine.log", tmp, refine_beam_tilt=beam)
            out[tag] = {"program": cmd.split(" << eot")[0].split("/")[-1].strip(),
                        "heredoc": heredoc(cmd).replace(tmp, "$SCRATCH"),
                        "parameters": {k: mp[k] for k in sorted(mp)}}
    finally:
        os.chdir(cwd)
    json.dump(out, open(os.path.join(HERE, "prompts_refine.json"), "w"), indent=1)
    for k, v in out.items():
        print(k, v["program"], len(v["heredoc"].splitlines()), "answers")
    # ---- reconstruct3d: frealign.py:1622-1835 split_reconstruction(run=False) returns the command
    rec = {}
    rcases = {
        "plain": {},
        "dose_blur_split": dict(reconstruct_dose_weighting_enable=True, reconstruct_dose_weighting_weights="",
                                reconstruct_dose_weighting_multiply=True, reconstruct_dose_weighting_fraction=4,
                                reconstruct_dose_weighting_transition=0.75, reconstruct_lblur=True,
                                reconstruct_per_particle_splitting=True, refine_score_weighting=True, reconstruct_apply_symmetry=False),
    }
    os.chdir(tmp)
    os.makedirs(os.path.join(tmp, "log"), exist_ok=True)
    try:
        for tag, kw in rcases.items():
            mp = base_mp()
            mp.update(slurm_tasks=7, reconstruct_radrec="0", extract_box=128, reconstruct_dose_weighting_enable=False,
                      refine_score_weighting=False, reconstruct_per_particle_splitting=False, reconstruct_lblur=False,
                      reconstruct_apply_symmetry=True, refine_bsc=2.0, reconstruct_cutoff="1", reconstruct_norm=True,
                      refine_adjust="F", refine_crop="F", reconstruct_scratch_copy_stack=False)
            mp.update(kw)
            cmd = F.split_reconstruction(mp, 1, 50, 3, 1, 1, 0, 0, dump_intermediate="yes", num_frames=1, run=False)
            rec[tag] = {"program": cmd.split(" << eot")[0].split("/")[-1].strip(), "heredoc": heredoc(cmd).replace(tmp, "$SCRATCH"),
                        "parameters": {k: mp[k] for k in sorted(mp)}}
    finally:
        os.chdir(cwd)
    json.dump(rec, open(os.path.join(HERE, "prompts_reconstruct.json"), "w"), indent=1)
    for k, v in rec.items():
        print(k, v["program"], len(v["heredoc"].splitlines()), "answers")
    # ---- merge3d / local_merge3d: the functions run their command; the executor is replaced by a recorder
    class Captured(Exception):
        pass

    def recorder(command, *a, **k):
        raise Captured(command)

    mer = {}
    F.local_run.stream_shell_command = recorder
    F.local_run.run_shell_command = recorder
    work = os.path.join(tmp, "merge")
    os.makedirs(os.path.join(work, "swarm"), exist_ok=True)
    os.makedirs(os.path.join(work, "log"), exist_ok=True)
    os.environ["PYP_SCRATCH"] = os.path.join(work, "scratch")
    os.makedirs(os.environ["PYP_SCRATCH"], exist_ok=True)
    for k in (1, 2, 3):
        for h in (1, 2):
            open(os.path.join(os.environ["PYP_SCRATCH"], f"T20S_r01_map{h}_n{k}.mrc"), "w").close()
            open(os.path.join(work, "swarm", f"T20S_r01_02_map{h}_n{k}.mrc"), "w").close()
    os.chdir(os.path.join(work, "swarm"))
    try:
        mp = base_mp()
        mp.update(refine_metric="new", reconstruct_radrec="0", extract_box=128, reconstruct_num_frames="1", refine_merge_normalize=False)
        try:
            F.merge_reconstructions(mp, 3, 1)
        except Captured as e:
            cmd = str(e)
            mer["merge3d"] = {"program": cmd.split(" << eot")[0].split("/")[-1].strip(),
                              "heredoc": heredoc(cmd).replace(os.environ["PYP_SCRATCH"], "$SCRATCH")}
        try:
            F.local_merge_reconstruction()
        except Captured as e:
            cmd = str(e)
            mer["local_merge3d"] = {"program": cmd.split(" << eot")[0].split("/")[-1].strip(), "heredoc": heredoc(cmd)}
    finally:
        os.chdir(cwd)
    # ---- csp argv: src/pyp/system/local_run.py:306-467 create_csp_split_commands returns the command lines
    from pyp.system import local_run as LR
    import pyp.inout.image.core as IC

    IC.get_image_dimensions = lambda path: [512, 512, 5]      # the tilt series / a movie: only sizes the chunks
    os.chdir(tmp)
    open("frames_csp.txt", "w").write("TS_01_000.tif\n")

    cp = Loud(extract_box=128, slurm_tasks=4, slurm_memory_per_task=16, csp_NumberOfRandomIterations=0, csp_frame_refinement=False)
    ptl, scan = list(range(0, 9)), list(range(0, 5))
    csp = {}
    for tag, mode, frames in (("extract", -2, False), ("particles", 2, False), ("micrographs", 3, False), ("frames", 3, True)):
        cmds, count, movies = LR.create_csp_split_commands("$PYP_DIR/external/CSP/csp", "TS_01_r01_02.cistem", mode, 4, "TS_01_r01",
                                                           "TS_01_stack.mrc", ptl, scan, [0], cp, use_frames=frames)
        csp[tag] = {"driver_mode": mode, "commands": cmds}
    os.chdir(cwd)
    json.dump(csp, open(os.path.join(HERE, "prompts_csp.json"), "w"), indent=1)
    for k, v in csp.items():
        print(k, len(v["commands"]), v["commands"][0])
    json.dump(mer, open(os.path.join(HERE, "prompts_merge.json"), "w"), indent=1)
    for k, v in mer.items():
        print(k, v["program"], v["heredoc"].splitlines())


if __name__ == "__main__":
    main()
