"""cProfile of the vision env's host side on the GPU box: where do the milliseconds of env.step() go (1024 envs)?"""
import cProfile, pstats, os, sys, io
import numpy as np
sys.path.insert(0, os.getcwd())
from flybody_b200 import fly_envs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = fly_envs.vision_guided_flight(n_envs=N, seed=1, terrain_bank=64)
env.reset()
rs = np.random.RandomState(0)
acts = rs.uniform(-0.2, 0.2, (40, N, 12))
for k in range(5): env.step(acts[k])
pr = cProfile.Profile(); pr.enable()
for k in range(5, 35): env.step(acts[k])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
