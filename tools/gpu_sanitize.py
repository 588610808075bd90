"""compute-sanitizer target: 64 envs x 1-2 control steps of every model variant through every launch path of the stepper
(stage kernels with and without the CUDA-graph replay, device-side task hooks, observation program, heightfield kernel,
eye renderer).  Run as
    compute-sanitizer --tool racecheck|synccheck|memcheck|initcheck python tools/gpu_sanitize.py
The host emulation runs a warp's lanes one after the other, so a missing __syncwarp / shared-memory hazard is invisible to the
CPU tests; this is the check that sees it."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from flybody_b200 import fly_envs, stepper as st
from flybody_b200.flymodel import load_model

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rs = np.random.RandomState(0)
for variant, n_sub, scale in (('walk', 10, 0.5), ('flight', 4, 0.2)):
    m = load_model(variant)
    s = st.BatchedStepper(m, N)
    q = np.tile(m.qpos0, (N, 1))
    hinge = [m.jnt_qposadr[j] for j in range(m.njnt) if m.jnt_type[j] == 3]
    q[:, hinge] += rs.uniform(-0.08, 0.08, (N, len(hinge)))          # deep interpenetrations: generic convex pairs, many rows
    s.reset(q, rs.uniform(-1, 1, (N, m.nv)))
    for k in range(3):                                               # 1st plain launches, 2nd captures the graph, 3rd replays it
        s.set_control(rs.uniform(-scale, scale, (N, m.nu)).astype(np.float32)); s.step(n_sub)
    s.pack_obs(); s.sync()
    print(variant, 'stepper ok: flags', int((s.get(st.FLAGS) != 0).sum()), 'nefc max', int(s.get(st.NEFC).max()), flush=True)
    s.close()
for make, na, scale in ((lambda: fly_envs.walk_imitation(n_envs=N, terminal_com_dist=0.02, device_task=True, reset_noise=0.05), 59, 0.5),
                        (lambda: fly_envs.flight_imitation(n_envs=N, terminal_com_dist=0.02, device_task=True, seed=2), 12, 0.2),
                        (lambda: fly_envs.walk_imitation(n_envs=N, terminal_com_dist=0.02), 59, 0.5)):
    env = make(); env.reset()
    for k in range(14):                                              # through terminations and auto-resets
        ts = env.step(rs.uniform(-scale, scale, (N, na)).astype(np.float32))
    print('env ok', na, 'first seen' if env.n_resets > N else 'no auto-reset yet', flush=True)
    env.close()
env = fly_envs.vision_guided_flight(n_envs=min(N, 32), seed=1, terrain_bank=4)
env.reset()
for k in range(3):
    ts = env.step(rs.uniform(-0.2, 0.2, (env.n_envs, 12)))
print('vision ok', {k: v.shape for k, v in ts.observation.items() if 'eye' in k}, flush=True)
env.close()
env = fly_envs.vision_guided_flight(n_envs=min(N, 32), seed=1, terrain_bank=4, device_task=True, time_limit=0.0008, target_height_range=(0.12, 0.6))
env.reset()
for k in range(8):                                                   # device-side vision task: terrain bank copies, contact / time-limit resets
    ts = env.step(rs.uniform(-0.2, 0.2, (env.n_envs, 12)))
print('vision device task ok, resets', env.n_resets, flush=True)
env.close()
print('SANITIZE_TARGET_DONE')
