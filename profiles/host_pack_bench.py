import sys, os, time, ctypes as C
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import vibo_b200
from vibo_b200 import _lib, kernels as K
lib = _lib.load()
print("threads", lib.vibo_host_threads(), "cpus", os.cpu_count())
P, I = 1000000, 1000
resp = torch.empty(P, I, dtype=torch.float32).pin_memory()
mask = torch.empty(P, I, dtype=torch.uint8).pin_memory()
resp.bernoulli_(0.5); mask.fill_(1)
out = torch.empty(P, I, dtype=torch.int8).pin_memory()
d = K.make_desc(P, I, 1, 2, False)
for rep in range(4):
    t = time.time()
    lib.vibo_pack_host(C.byref(d), resp.data_ptr(), mask.data_ptr(), out.data_ptr())
    dt = time.time() - t
    print("pack", round(dt * 1e3, 1), "ms", round(P * I * 5 / dt / 1e9, 1), "GB/s raw")
# with a concurrent H2D of the raw rows
dev = torch.device("cuda")
dr = torch.empty(P // 4, I, dtype=torch.float32, device=dev)
s = torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize()
    t = time.time()
    with torch.cuda.stream(s):
        for k in range(8):
            dr.copy_(resp[(k % 4) * (P // 4):(k % 4 + 1) * (P // 4)], non_blocking=True)
    lib.vibo_pack_host(C.byref(d), resp.data_ptr(), mask.data_ptr(), out.data_ptr())
    tp = time.time() - t
    torch.cuda.synchronize()
    tc = time.time() - t
    print("concurrent: pack", round(tp * 1e3, 1), "ms =", round(P * I * 5 / tp / 1e9, 1), "GB/s; 8 GB H2D done at", round(tc * 1e3, 1), "ms =", round(8 / tc, 1), "GB/s")
