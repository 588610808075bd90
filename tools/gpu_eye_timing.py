"""Eye ray caster on the GPU box: render time of 1024 envs x 2 eyes x 32 x 32 with the block-maximum skipping of the terrain
march (default) and without (FB_EYE_NO_SKIP=1), and that both give the same images."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from flybody_b200 import fly_envs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
imgs = {}
for mode in ('skip', 'noskip'):
    if mode == 'noskip':
        os.environ['FB_EYE_NO_SKIP'] = '1'
    else:
        os.environ.pop('FB_EYE_NO_SKIP', None)
    env = fly_envs.vision_guided_flight(n_envs=N, seed=1, terrain_bank=64)
    env.reset()
    rs = np.random.RandomState(0)
    for k in range(5):
        env.step(rs.uniform(-0.2, 0.2, (N, 12)))
    stream = torch.cuda.ExternalStream(env._sim.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(5):
        env._sim.render_eyes_async()
    stream.synchronize()
    e0.record(stream)
    for k in range(50):
        env._sim.render_eyes_async()
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / 50
    imgs[mode] = env.eyes_device().cpu().numpy().copy()
    print(f'{mode}: {ms:.4f} ms per render of {N} envs ({N * 2 * 32 * 32 / ms / 1e6:.2f} G rays/s)', flush=True)
    env.close()
d = np.abs(imgs['skip'].astype(np.int64) - imgs['noskip'].astype(np.int64))
print('images equal:', bool((d == 0).all()), 'max diff', int(d.max()), 'frac differing pixels', float((d.max(-1) > 0).mean()))
