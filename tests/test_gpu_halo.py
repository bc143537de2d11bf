"""mx_halo_pack / mx_halo_unpack (device) against the numpy slab copies used by the gloo test:
the FillBoundary pack/unpack step (AMReX_FabArrayCommI.H:49-150) is bit-exact data movement."""
import numpy as np
import pytest

from gpu_util import to_device, to_host
from maestroex_b200.synth import make_problem
from test_decomp_gloo import NumpyPacker

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ng", [3, 4])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_pack_unpack_roundtrip(d, ng):
    import torch
    from maestroex_b200 import lib as mxlib
    from maestroex_b200.decomp import DevicePacker
    prob = make_problem(dim=3, n=(20, 12, 9), nspec=3, ng=ng, bc_kind="periodic", seed=3)
    dprob = to_device(prob)
    dp, npk = DevicePacker(mxlib.calls(), torch.device("cuda:0")), NumpyPacker()
    c0, nc = 2, 4
    n = dp.count(prob.geom, dprob.state, d, ng, nc)
    assert n == npk.count(prob.geom, prob.state, d, ng, nc)
    for side in (0, 1):
        dbuf, hbuf = dp.alloc(n), npk.alloc(n)
        dp.pack(prob.geom, dprob.state, d, side, ng, c0, nc, dbuf)
        npk.pack(prob.geom, prob.state, d, side, ng, c0, nc, hbuf)
        torch.cuda.synchronize()
        assert np.array_equal(dbuf.cpu().numpy(), hbuf.numpy())
        # periodic self-fill: my `side` interior becomes my opposite ghosts
        dp.unpack(prob.geom, dprob.state, d, 1 - side, ng, c0, nc, dbuf)
        npk.unpack(prob.geom, prob.state, d, 1 - side, ng, c0, nc, hbuf)
    got = to_host([dprob.state])[0]
    assert np.array_equal(got.data, prob.state.data)
