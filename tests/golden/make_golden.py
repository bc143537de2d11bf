#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/*.npz from the REFERENCE itself.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
Each fixture holds the complete inputs of one small single-box case (state, force, MAC
velocities, base-state arrays, BC table, dt, rel_eps ...) and the outputs of the reference's
own Maestro::MakeEdgeScal -> MakeRhoXFlux -> MakeRhoHFlux (and CelltoEdge, Makew0), obtained from
oracle/_ref/libmaestro_ref{2,3}d.so = Source/Maestro{MakeEdgeScalars,PPM,Slopes,MakeFlux,
CelltoEdge}.cpp compiled unmodified (oracle/Makefile).  The fixtures travel to the GPU box,
the reference does not: tests/test_golden.py checks the C oracle (CPU) and libmxadv.so (GPU)
against them bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import run_edge, run_rhoh, run_rhoX  # noqa: E402
from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH, make_problem  # noqa: E402

# name -> (dim, n, nspec, ppm_type, slope_order, bc_kind, spt, ept, seed)
CASES = {
    "bubble3d_ppm1": (3, (14, 12, 10), 3, 1, 4, "reacting_bubble", 1, 1, 101),
    "bubble3d_ppm0": (3, (12, 10, 9), 3, 0, 4, "reacting_bubble", 1, 1, 102),
    "bubble3d_ppm2": (3, (12, 11, 10), 3, 2, 4, "reacting_bubble", 1, 1, 103),
    "mixed3d_ppm1_rho_and_X_h": (3, (10, 9, 8), 2, 1, 4, "mixed", 3, 2, 104),
    "mixed3d_ppm2_slope2": (3, (9, 9, 9), 2, 2, 2, "mixed", 1, 0, 105),
    "outflow3d_ppm0_slope2": (3, (10, 8, 9), 3, 0, 2, "outflow", 2, 5, 106),
    "bubble2d_ppm1": (2, (24, 20), 3, 1, 4, "reacting_bubble", 1, 1, 107),
    "bubble2d_ppm0": (2, (20, 16), 3, 0, 4, "reacting_bubble", 1, 1, 108),
    "mixed2d_ppm2": (2, (18, 14), 2, 2, 4, "mixed", 3, 2, 109),
}


def build(name):
    dim, n, nspec, ppm, so, bck, spt, ept, seed = CASES[name]
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=dim, n=n, nspec=nspec, ng=ng, bc_kind=bck, seed=seed)
    return prob, dict(ppm_type=ppm, slope_order=so), dict(species_pred_type=spt,
                                                          enthalpy_pred_type=ept)


def run_case(calls, prob, ekw, fkw):
    """The density + enthalpy pass as DensityAdvance / EnthalpyAdvance call it."""
    se = run_edge(calls, prob, prob.edge_params(FIRSTSPEC, prob.nspec, **ekw))
    run_edge(calls, prob, prob.edge_params(RHO, 1, **ekw), sedge=se)
    run_edge(calls, prob, prob.edge_params(RHOH, 1, **ekw), sedge=se)
    sf, eta = run_rhoX(calls, prob, se, prob.flux_params(**fkw))
    run_rhoh(calls, prob, se, prob.flux_params(**fkw), sflux=sf)
    return se, sf, eta


def save(name, prob, se, sf, eta):
    d = {"state": prob.state.data, "force": prob.force.data, "bc": prob.bc,
         "dt": np.float64(prob.dt), "rel_eps": np.float64(prob.rel_eps),
         "eta_in": prob.etarhoflux.data, "eta_out": eta.data}
    for i, u in enumerate(prob.umac):
        d[f"umac{i}"] = u.data
    for k, b in prob.base.items():
        d[f"base_{k}"] = b.data
    for i in range(len(se)):
        d[f"sedge{i}"] = se[i].data
        d[f"sflux{i}"] = sf[i].data
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def main():
    from oracle import backends
    for name in CASES:
        prob, ekw, fkw = build(name)
        ref = backends.reference(prob.geom.dim)
        assert ref is not None, "oracle/_ref is not built: needs /root/reference"
        se, sf, eta = run_case(ref, prob, ekw, fkw)
        save(name, prob, se, sf, eta)
        print(name, "ok")
    # CelltoEdge
    ref = backends.reference(3)
    rng = np.random.default_rng(7)
    nr = 48
    cell = 1.0e6 * np.exp(-6.0 * (np.arange(nr) + 0.5) / nr) * (1.0 + 0.05 * rng.random(nr))
    out = {}
    for lo, hi in ((0, nr - 1), (8, 39)):
        edge = np.full(nr + 1, -3.0)
        ref.cell_to_edge(cell, edge, nr, lo, hi)
        out[f"edge_{lo}_{hi}"] = edge
    np.savez_compressed(os.path.join(HERE, "cell_to_edge.npz"), cell=cell, **out)
    print("cell_to_edge ok")
    # Makew0 (Source/MaestroMakew0.cpp + MaestroMakeGrav.cpp compiled unmodified)
    from maestroex_b200.synth import W0_CASES, make_w0_problem, run_w0
    for name in W0_CASES:
        sph, par, d = make_w0_problem(name)
        w0, force, pb = run_w0(ref, sph, par, d)
        if name.startswith("irreg"):
            force[0] = 0.0  # undefined in the reference (out-of-bounds read, :926): not compared
        np.savez_compressed(os.path.join(HERE, "w0_" + name + ".npz"), w0=w0, w0_force=force,
                            p0bdot=np.float64(pb), **{"in_" + k: v for k, v in d.items()})
        print("w0", name, "ok")


if __name__ == "__main__":
    main()
