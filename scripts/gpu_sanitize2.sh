cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from maestroex_b200 import lib as mxlib
from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH, make_problem
from maestroex_b200.fusedcalls import density_edge_flux
from gpu_util import to_device
import torch
calls = mxlib.calls()
prob = to_device(make_problem(dim=3, n=(37, 19, 40), nspec=3, ng=3, bc_kind="periodic", seed=5))
density_edge_flux(calls, prob, prob.edge_params(FIRSTSPEC, 3, ppm_type=1), prob.flux_params())
torch.cuda.synchronize()
PY
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=k_stream4 python /tmp/san.py 2>&1 | grep -v "^=========     Saved host\|Host Frame\|\[0x" | head -60
