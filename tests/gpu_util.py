"""Helpers for the GPU parity tests: move a synthetic Problem onto the device (torch)."""
import copy

import numpy as np

from maestroex_b200.abi import Base, Fab


def to_device(prob):
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())  # one rank per GPU under torchrun

    def fab(f):
        return Fab(torch.from_numpy(np.ascontiguousarray(f.data)).to(dev), f.lo, f.hi)

    p = copy.copy(prob)
    p.state = fab(prob.state)
    p.force = fab(prob.force)
    p.umac = [fab(f) for f in prob.umac]
    p.sedge = [fab(f) for f in prob.sedge]
    p.sflux = [fab(f) for f in prob.sflux]
    p.etarhoflux = fab(prob.etarhoflux)
    p.base = {k: Base(torch.from_numpy(np.ascontiguousarray(v.data)).to(dev))
              for k, v in prob.base.items()}
    p.mac = {k: [fab(f) for f in v] for k, v in prob.mac.items()}

    def clone_outputs():
        cp = lambda f: Fab(f.data.clone(), f.lo, f.hi)  # noqa: E731
        return [cp(f) for f in p.sedge], [cp(f) for f in p.sflux], cp(p.etarhoflux)

    p.clone_outputs = clone_outputs
    return p


def to_host(fabs):
    import torch
    torch.cuda.synchronize()
    out = []
    for f in fabs:
        out.append(Fab(f.data.cpu().numpy(), f.lo, f.hi))
    return out
