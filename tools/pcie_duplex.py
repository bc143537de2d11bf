"""PCIe micro-benchmark for the host-fab path: H2D alone, D2H alone, both at once (pinned)."""
import time
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_in, do_out, reps=4, chunk=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if chunk is None:
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        else:
            for o in range(0, n, chunk):
                if do_in:
                    with torch.cuda.stream(s1):
                        d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
                if do_out:
                    with torch.cuda.stream(s2):
                        h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return (int(do_in) + int(do_out)) * n * reps / dt / 1e9


run(True, True)
print("H2D alone  GB/s", round(run(True, False), 1))
print("D2H alone  GB/s", round(run(False, True), 1))
print("both       GB/s (aggregate)", round(run(True, True), 1))
print("both, 2 MiB chunks GB/s (aggregate)", round(run(True, True, chunk=2 << 20), 1))
print("D2H alone, 2 MiB chunks GB/s", round(run(False, True, chunk=2 << 20), 1))
