"""How fast is pinning a caller's pageable numpy array in place (cudaHostRegister) versus copying it through a pinned
ring with host threads?  Decides the design of the pageable-host path of b200gate_run."""
import time
import numpy as np
import torch

rt = torch.cuda.cudart()
n = 64 * 28_800_000
a = np.empty(n, dtype=np.float32)
a[:] = 1.0                                           # faulted-in input
b = np.empty(n, dtype=np.float32)                    # fresh output: pages not faulted yet
for name, arr in (("input (faulted)", a), ("output (fresh np.empty)", b)):
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
    dt = time.perf_counter() - t0
    print(f"cudaHostRegister {name}: rc={rc} {dt*1e3:.1f} ms  {arr.nbytes/dt/1e9:.1f} GB/s", flush=True)
    t0 = time.perf_counter()
    rt.cudaHostUnregister(arr.ctypes.data)
    print(f"  unregister {(time.perf_counter()-t0)*1e3:.1f} ms", flush=True)
# threaded memcpy into a pinned buffer
import threading
pin = torch.empty(n, dtype=torch.float32, pin_memory=True).numpy()
for nt in (1, 4, 8, 16, 32):
    parts = np.array_split(np.arange(n), nt)
    def work(i):
        lo, hi = parts[i][0], parts[i][-1] + 1
        np.copyto(pin[lo:hi], a[lo:hi])
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"memcpy pageable->pinned with {nt} threads: {dt*1e3:.1f} ms {a.nbytes/dt/1e9:.1f} GB/s", flush=True)
d = torch.empty(n, dtype=torch.float32, device="cuda")
t0 = time.perf_counter(); d.copy_(torch.from_numpy(a)); torch.cuda.synchronize()
print(f"pageable H2D (torch copy_): {(time.perf_counter()-t0)*1e3:.1f} ms")
t0 = time.perf_counter(); d.copy_(torch.from_numpy(pin), non_blocking=True); torch.cuda.synchronize()
print(f"pinned H2D: {(time.perf_counter()-t0)*1e3:.1f} ms")
