mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from maestroex_b200 import lib as mxlib
from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH, make_problem
from maestroex_b200.fusedcalls import density_edge_flux, enthalpy_edge_flux
from gpu_util import to_device
import torch
calls = mxlib.calls()
for ppm, bc, sph in ((1, "reacting_bubble", False), (2, "mixed", False), (0, "outflow", True)):
    ng = 4 if ppm == 2 else 3
    prob = to_device(make_problem(dim=3, n=(37, 19, 40), nspec=3, ng=ng, bc_kind=bc, seed=5, spherical=sph))
    fpar = prob.flux_params(spherical=1 if sph else 0)
    density_edge_flux(calls, prob, prob.edge_params(FIRSTSPEC, 3, ppm_type=ppm), fpar)
    if not sph:
        enthalpy_edge_flux(calls, prob, prob.edge_params(RHOH, 1, ppm_type=ppm), fpar)
    torch.cuda.synchronize()
print("sanitizer workload done")
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=k_stream4 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|done" | head -8
done
