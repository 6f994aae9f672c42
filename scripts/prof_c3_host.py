"""cProfile of MockStreamGenerator.run at C3 size: where the host time goes."""
import cProfile, pstats, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import galax_b200.dynamics as gd, galax_b200.potential as gp
pot = gp.MilkyWayPotential(); M = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
ts = np.linspace(0.0, 3000.0, M)
w0 = gd.PhaseSpaceCoordinate(np.array([30.0, 10, 20]), np.array([10.0, -150, -20]) * gp.KMS, 0.0)
draws = np.random.default_rng(3).standard_normal((4, M))
gen = gd.MockStreamGenerator(gd.FardalStreamDF(), pot)
gen.run(draws[:, :1000], ts[:1000], w0, 1e4); gen.run(draws, ts, w0, 1e4)
pr = cProfile.Profile(); pr.enable()
stream, prog = gen.run(draws, ts, w0, 1e4); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
