"""HBM ceilings for the traffic patterns of this path, measured with plain torch ops on large buffers:
copy (1 read : 1 write, what MEASURED_PEAKS.json quotes), write-only (fill), read-only (sum), and a 1:4 read:write
mix like K1's acc + Hessian output (24 B in, 96 B out per point)."""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from quick_perf import ev_time
n = 1_500_000_000  # doubles: 12 GB
a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty(n, dtype=torch.float64, device="cuda")
a.fill_(1.0); b.fill_(2.0)
t, _ = ev_time(lambda: b.copy_(a), reps=5, warm=2); print(f"copy  (1r:1w): {2*n*8/t/1e9:.0f} GB/s")
t, _ = ev_time(lambda: a.fill_(3.0), reps=5, warm=2); print(f"fill  (0r:1w): {n*8/t/1e9:.0f} GB/s")
t, _ = ev_time(lambda: a.sum(), reps=5, warm=2); print(f"sum   (1r:0w): {n*8/t/1e9:.0f} GB/s")
# 1 read : 4 writes: out[4, m] = in[m] broadcast (expand + copy into a 4x larger buffer)
m = n // 4
src = a[:m]; dst = b[: 4 * m].view(4, m)
t, _ = ev_time(lambda: dst.copy_(src.expand(4, m)), reps=5, warm=2); print(f"mix   (1r:4w): {5*m*8/t/1e9:.0f} GB/s")
