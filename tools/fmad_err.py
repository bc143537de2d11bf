"""Tolerance-mode A/B: the same density + enthalpy pass with the strict
library (-fmad=false, bit-identical to the reference) and an FMA-contraction build
(exp/fmad/libmxadv.so, `FMAD=true scripts/exp_build.sh fmad`), on one seeded 128^3 x 13-species box.
Run once per library (MXADV_LIB selects it) with `dump <file>`, then `compare <a> <b>`: the worst
relative difference per output array against north_star's 1e-12."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def dump(path):
    import torch
    from maestroex_b200 import lib as mxlib
    from maestroex_b200.devsynth import make_device_problem
    from maestroex_b200.fusedcalls import density_edge_flux, enthalpy_edge_flux
    from maestroex_b200.synth import FIRSTSPEC, RHOH
    n = 128
    prob = make_device_problem(n=(n, n, n), nspec=13, ng=3, bc_kind="reacting_bubble", seed=4242, device="cuda:0")
    calls = mxlib.calls()
    out = (prob.sedge, prob.sflux, prob.etarhoflux)
    density_edge_flux(calls, prob, prob.edge_params(FIRSTSPEC, 13, ppm_type=1), prob.flux_params(), outputs=out)
    enthalpy_edge_flux(calls, prob, prob.edge_params(RHOH, 1, ppm_type=1), prob.flux_params(), outputs=out)
    torch.cuda.synchronize()
    arrs = {f"sedge{d}": prob.sedge[d].data.cpu().numpy() for d in range(3)}
    arrs.update({f"sflux{d}": prob.sflux[d].data.cpu().numpy() for d in range(3)})
    arrs["eta"] = prob.etarhoflux.data.cpu().numpy()
    np.savez(path, **arrs)


def compare(a, b):
    A, B = np.load(a), np.load(b)
    worst = 0.0
    for k in A.files:
        x, y = A[k], B[k]
        # relative to the array's own scale per component (zones where a quantity passes
        # through zero make a pointwise ratio meaningless: SURVEY.md hard part 2)
        for c in range(x.shape[0]):
            scale = np.abs(x[c]).max()
            if scale == 0.0:
                continue
            e = np.abs(x[c] - y[c]).max() / scale
            worst = max(worst, e)
        pt = np.abs(x - y) / np.maximum(np.abs(x), 1e-300)
        print(f"{k:8s} max |a-b|/max|a| per comp = {max(np.abs(x[c]-y[c]).max()/max(np.abs(x[c]).max(),1e-300) for c in range(x.shape[0])):.3e}"
              f"   worst pointwise rel = {pt.max():.3e}   identical zones = {100.0*np.mean(x==y):.2f} %")
    print(f"WORST scaled relative difference {worst:.3e}  (north_star tolerance 1e-12)")


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2])
    else:
        compare(sys.argv[2], sys.argv[3])
