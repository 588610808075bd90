"""Host-side breakdown of one env.step() (public API, 4096 envs): cProfile over 30 steps."""
import cProfile, pstats, sys, os, io
import numpy as np
sys.path.insert(0, os.getcwd())
from flybody_b200 import fly_envs
N = 4096
env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=N, reset_noise=0.05, seed=1)
env.reset()
rs = np.random.RandomState(0)
acts = rs.uniform(-0.5, 0.5, (40, N, 59)).astype(np.float32)
for k in range(5):
    env.step(acts[k])
pr = cProfile.Profile(); pr.enable()
for k in range(5, 35):
    env.step(acts[k])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22); print(s.getvalue()[:4500])
