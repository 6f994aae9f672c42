import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
from galax_b200 import _lib
from quick_perf import ev_time
L = _lib.lib(); sink = torch.zeros(8, dtype=torch.float64, device="cuda"); n = C.c_int64()
for sign, nm in ((1, "1-reg-operand"), (-1, "3-reg-operand")):
    for blocks, threads in ((148 * 8, 256), (148 * 4, 128), (148 * 16, 128)):
        f = lambda: L.gx_bench_dfma(sign * blocks, threads, 20000, sink.data_ptr(), C.byref(n), None)
        best, med = ev_time(f)
        print(f"dfma {nm} blocks={blocks} threads={threads}: {2.0 * n.value * blocks * threads / best / 1e12:.2f} TFLOP/s")
