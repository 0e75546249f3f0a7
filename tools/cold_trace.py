"""First solve of a model in a fresh process with the library's own timeline on stderr (verbosity 1)."""
import sys
import time

sys.path.insert(0, ".")
from stcsp_solver_b200 import binding, instances

names = sys.argv[1:] or ["juggling_b6_f6_nosym", "partialorder_14", "digitinvader9"]
import ctypes
binding.lib().stcsp_gpu_device_count()
t0 = time.perf_counter()
ctypes.CDLL("libcudart.so").cudaFree(0)             # context creation, outside the first solve
print("context %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
for name in names:
    model = binding.Model(instances.by_name(name))
    for i in range(3):
        t0 = time.perf_counter()
        a = binding.solve(model, binding.default_options(verbosity=1 if i == 0 else 0))
        w = (time.perf_counter() - t0) * 1e3
        print("%s solve %d: e2e %.3f ms device %.3f ms launches %d waves %d" % (name, i, w, a.c.solve_ms, a.c.n_kernel_launches, a.c.n_waves), flush=True)
        del a
