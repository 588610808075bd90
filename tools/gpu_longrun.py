"""Steady-state measurement: walk_imitation, N envs, random policy, task hooks on the device (auto-reset
at termination / episode end), 1000 control steps after a 50-step warm-up FROM A STANDING RESET (no staggered pre-roll: this
log shows what the pre-roll of bench.py is for).  Prints, per block of 25 control steps: ms per step (CUDA events), auto-resets,
mean / max constraint rows, share of envs on the global-memory solver path (nefc > 32), mean contacts, flagged envs.
    python tools/gpu_longrun.py [n_envs] [steps] [terminal_com_dist]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from flybody_b200 import fly_envs, stepper as st

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tcd = float(sys.argv[3]) if len(sys.argv) > 3 else float('inf')
env = fly_envs.walk_imitation(terminal_com_dist=tcd, n_envs=N, reset_noise=0.05, seed=1234, device_task=True)
env.reset()
sim = env.physics.stepper
stream = torch.cuda.ExternalStream(sim.stream)
gen = torch.Generator(device='cuda'); gen.manual_seed(1234)
acts = (torch.rand((64, N, 59), device='cuda', generator=gen) - 0.5)
BLK = 25
print(json.dumps({'n_envs': N, 'steps': K, 'terminal_com_dist': tcd, 'episode_steps': int(env._episode_steps[0]), 'block': BLK}))
tot_ms, ep_prev = 0.0, env.device_reset_count()
hist_all = np.zeros(161, np.int64)
for blk in range((50 + K) // BLK):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for k in range(BLK):
            env.step_device(acts[(blk * BLK + k) % 64])
        e1.record(stream)
    sim.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / BLK
    nefc = sim.get(st.NEFC)[:, 0].astype(np.int64); ncon = sim.get(st.NCON)[:, 0]; flags = sim.get(st.FLAGS)[:, 0].astype(np.int64)
    ep = env.device_reset_count()
    step0 = blk * BLK
    if step0 >= 50:
        tot_ms += ms * BLK; hist_all += np.bincount(np.minimum(nefc, 160), minlength=161)
    print(json.dumps({'step': step0, 'ms_per_step': round(ms, 3), 'env_steps_per_s': round(N / ms * 1e3), 'resets': int(ep - ep_prev), 'nefc_mean': round(float(nefc.mean()), 2),
                      'nefc_p99': int(np.percentile(nefc, 99)), 'nefc_max': int(nefc.max()), 'share_nefc_gt32': round(float((nefc > 32).mean()), 4), 'ncon_mean': round(float(ncon.mean()), 2),
                      'diverged': int((flags & 1).sum()), 'overflow': int(((flags & 6) != 0).sum())}), flush=True)
    ep_prev = ep
n_timed = ((50 + K) // BLK) * BLK - 50
cum = np.cumsum(hist_all) / max(hist_all.sum(), 1)
print(json.dumps({'summary': True, 'timed_steps': n_timed, 'mean_ms_per_step': round(tot_ms / n_timed, 3), 'env_steps_per_s': round(N * n_timed / tot_ms * 1e3),
                  'nefc_cdf': {str(b): round(float(cum[b]), 4) for b in (8, 16, 24, 32, 48, 64, 96, 128, 159)}}))
