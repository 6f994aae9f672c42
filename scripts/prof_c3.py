"""C3 phase timing (progenitor orbit, release, stream integration) with torch's profiler for a host/device split."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
from torch.profiler import profile, ProfilerActivity

pot = gp.MilkyWayPotential(); M = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
ts = np.linspace(0.0, 3000.0, M)
w0 = gd.PhaseSpaceCoordinate(np.array([30.0, 10, 20]), np.array([10.0, -150, -20]) * gp.KMS, 0.0)
draws = np.random.default_rng(3).standard_normal((4, M))
gen = gd.MockStreamGenerator(gd.FardalStreamDF(), pot)
gen.run(draws[:, :1000], ts[:1000], w0, 1e4)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stream, prog = gen.run(draws, ts, w0, 1e4)
    torch.cuda.synchronize(); print("total", time.perf_counter() - t0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    stream, prog = gen.run(draws, ts, w0, 1e4)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=10, max_name_column_width=60))
