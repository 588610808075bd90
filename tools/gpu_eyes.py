"""Throughput of the eye ray caster: 2 x 32 x 32 rays per env over per-env 401 x 401 heightfields (the arena of
vision_guided_flight), flies at their flight_imitation start pose:  python tools/gpu_eyes.py [n_envs]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from flybody_b200 import arenas, fly_envs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = fly_envs.flight_imitation(n_envs=N, device_task=True)
env.reset()
nrow, ncol = arenas.grid_shape(20, 10)
env.enable_eyes(size=32, fovy=150.0, terrain_shape=(nrow, ncol), half_size=20.0, z_offset=-0.01)
gen = arenas.SineBumps()
terr = np.stack([gen.generate(np.random.RandomState(k)) for k in range(8)]).astype(np.float32)
for k in range(0, N, 8):
    ids = np.arange(k, min(k + 8, N))
    env.set_terrain(ids, terr[:len(ids)])
sim = env._sim
for _ in range(3):
    sim._lib.fb_render_eyes(sim._h)
sim.sync()
t0 = time.perf_counter()
K = 20
for _ in range(K):
    sim._lib.fb_render_eyes(sim._h)
sim.sync()
dt = (time.perf_counter() - t0) / K
img = env.render_eyes()
px = np.concatenate([v.reshape(-1, 3) for v in img.values()]).astype(np.float64)
print(f'N={N}: {dt * 1e3:.3f} ms per render of {N} x 2 x 32 x 32 rays = {N * 2048 / dt / 1e9:.2f} Grays/s; pixel mean {px.mean():.1f} std {px.std():.1f}', flush=True)
