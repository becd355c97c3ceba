import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pyp_b200 import synth, synth_torch
from pyp_b200._lib import ROW_DTYPE
from pyp_b200.engine import Engine
P, n, px = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 256, 1.0
dev = torch.device("cuda", 0)
c = bench.CONFIGS["C2"]
rcfg = bench.fill(Engine.refine_defaults(n, px), bench.refine_params(c))
ccfg = bench.fill(Engine.recon_defaults(n, px), bench.recon_params(c))
eng = Engine(0); eng.refine_configure(rcfg); eng.set_symmetry("O")
centres, amps, sigma = synth_torch.symmetric_phantom(n, "O")
vol = synth_torch.volume(n, centres, amps, sigma, dev)
truth = synth.make_rows(P, px, seed=1000); start = synth.perturb_rows(truth, 2.0, 1.0, seed=2000)
stack = synth_torch.make_stack(n, centres, amps, sigma, truth, snr=0.05, seed=3000, device=dev)
rows_host = np.ascontiguousarray(start, dtype=ROW_DTYPE)
eng.set_reference(vol)
host_stack = torch.empty((P, n, n), dtype=torch.float32, pin_memory=True); host_stack.copy_(stack); torch.cuda.synchronize()
hs = host_stack.numpy()
def T(f, *a, **k):
    eng.sync(); t0 = time.perf_counter(); r = f(*a, **k); eng.sync(); return (time.perf_counter() - t0) * 1e3, r
for rep in range(3):
    t_begin, _ = T(eng.recon_begin, ccfg)
    t_rr, _ = T(eng.refine_reconstruct, hs, rows_host)
    t_fin, _ = T(eng.recon_finalize, molecular_mass_kda=440.0, want_halves=True)
    # pure copy
    d = torch.empty((P, n, n), dtype=torch.float32, device=dev)
    t_cp, _ = T(lambda: (d.copy_(host_stack, non_blocking=True), torch.cuda.synchronize()))
    del d
    t_r, _ = T(eng.refine_reconstruct, hs, rows_host, insert=False)
    print(f"rep {rep}: recon_begin {t_begin:.1f} ms, refine_reconstruct {t_rr:.1f} ms, refine only {t_r:.1f}, finalize(host) {t_fin:.1f} ms, pure H2D {t_cp:.1f} ms ({P*n*n*4/t_cp/1e6:.1f} GB/s)")
