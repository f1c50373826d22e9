"""Experiment: masked RMSProp on three separate [N,28] arrays vs one interleaved [N,3,28] buffer."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svox2_b200 import _lib

N = 13_388_269
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
seg = torch.rand((N + 63) // 64, generator=g, device=dev) < 0.25
mask = (seg.repeat_interleave(64)[:N] & (torch.rand(N, generator=g, device=dev) < 0.5)).contiguous()
print("touched rows", int(mask.sum()), "of", N)
L = _lib.lib()

def run(data, rms, grad, stride, reps=20):
    def once():
        _lib.check(L.svox2_rmsprop_step(data.data_ptr(), rms.data_ptr(), grad.data_ptr(), N, 27, stride, 1,
                                        mask.data_ptr(), None, 0, 0.95, 1e-2, 1e-8, -1e9, -1.0,
                                        torch.cuda.current_stream().cuda_stream))
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

sep = [torch.rand(N, 28, device=dev) for _ in range(3)]
print("separate   : %.1f us" % run(sep[0], sep[1], sep[2], 28))
del sep
inter = torch.rand(N, 3, 28, device=dev)
print("interleaved: %.1f us" % run(inter[:, 0], inter[:, 1], inter[:, 2], 84))
sep = [torch.rand(N, 32, device=dev) for _ in range(3)]
print("separate, 128-B rows   : %.1f us" % run(sep[0], sep[1], sep[2], 32))
del sep
inter = torch.rand(N, 3, 32, device=dev)
print("interleaved, 128-B rows: %.1f us" % run(inter[:, 0], inter[:, 1], inter[:, 2], 96))
