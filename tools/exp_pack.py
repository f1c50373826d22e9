"""Experiment: time the exchange kernels alone (pack, packed RMSProp) at bench-like sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svox2_b200 import csrc as C

N, R = 13_388_269, 2_900_000
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
rows = torch.sort(torch.randperm(N, generator=g, device=dev)[:R])[0].contiguous()
mk = lambda: torch.rand(N, 28, device=dev)[:, :27]
gsh, sh, rms = mk(), mk(), mk()
gd, den, drms = (torch.rand(N, 1, device=dev) for _ in range(3))
packed = torch.empty(R, 28, device=dev)
mask = torch.zeros(N, dtype=torch.bool, device=dev)
mask[rows] = True

def timeit(f, reps=10):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

print("pack            : %.1f us" % timeit(lambda: C.pack_grad_rows(gsh, gd, rows, packed)))
print("packed rmsprop  : %.1f us" % timeit(lambda: C.rmsprop_step_packed(sh, rms, den, drms, rows, packed, mask, mask, 0.95, 1e-2, 30.0, 1e-8, grad_sh=gsh, grad_density=gd)))
print("masked rmsprop sh : %.1f us" % timeit(lambda: C.rmsprop_step(sh, rms, gsh, mask, 0.95, 1e-2, 1e-8, -1e9, 1e-2)))
print("nonzero         : %.1f us" % timeit(lambda: torch.nonzero(mask)))
