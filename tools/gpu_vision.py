"""First GPU run of vision_guided_flight (not yet executed on a B200 when this was written): (1) terrain contacts of the CUDA
heightfield kernel against the host-emulation build on the same states, (2) env-steps/s of `fly_envs.vision_guided_flight`
through the public API (host task code, eyes rendered every step), (3) stage times.
    python tools/gpu_vision.py [n_envs] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
from flybody_b200 import arenas, fly_envs, stepper as st
from flybody_b200.flymodel import load_model

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ge.build()
# (1) contact parity GPU vs emulation: flies dipped into a terrain
m = load_model('vision')
terr = arenas.SineBumps().generate(np.random.RandomState(4)).astype(np.float32)
q = np.tile(m.qpos0, (8, 1)); v = np.zeros((8, m.nv))
for e in range(8):
    x, y = -5.0 + e, 0.3 * e
    q[e, :3] = [x, y, float(arenas.hfield_height(terr, [x], [y], 20.0)[0]) - 0.01 + 0.12 + 0.02 * e]
out = []
for lib in (None, ge.EMU):
    sim = st.BatchedStepper(m, 8, lib_path=lib)
    sim.hfield_collision(m.meta['hf_geom'], m.hf_size, 401, 401, m.hf_pair_geom)
    sim.hfield_write(np.arange(8), np.tile(terr, (8, 1, 1)))
    sim.reset(q, v); sim.set_control(np.zeros((8, m.nu), np.float32)); sim.forward()
    out.append((sim.get(st.NCON)[:, 0].copy(), sim.get(st.CONTACT).copy(), sim.get(st.QACC).copy()))
    sim.close()
print('ncon gpu', out[0][0], 'emu', out[1][0])
same = np.array_equal(out[0][0], out[1][0])
print('contact counts equal:', same, ' max |d dist|:', float(np.abs(out[0][1] - out[1][1]).max()) if same else None,
      ' qacc rel err:', float(np.abs(out[0][2] - out[1][2]).max() / np.abs(out[1][2]).max()))
# (2) env throughput
env = fly_envs.vision_guided_flight(n_envs=N, seed=1, terrain_bank=64)
t0 = time.perf_counter(); env.reset(); print(f'reset of {N} envs (terrain generation on the host): {time.perf_counter() - t0:.1f} s')
rs = np.random.RandomState(0)
acts = rs.uniform(-0.2, 0.2, (K + 5, N, 12))
for k in range(5):
    env.step(acts[k])
t0 = time.perf_counter(); n_last = 0
for k in range(5, K + 5):
    ts = env.step(acts[k]); n_last += int((np.asarray(ts.step_type) == 2).sum())
dt = (time.perf_counter() - t0) / K
print(f'vision_guided_flight N={N}: {dt * 1e3:.3f} ms/step {N / dt:.0f} env-steps/s (host task code, eyes every step)  terminations {n_last}  mean reward {float(np.mean(ts.reward)):.3f}')
# (3) stage times
sim = env._sim
sim.profile(True)
for k in range(5):
    env.step(acts[k])
p = sim.profile_read()
print({k: round(v[0] / 5, 3) for k, v in p.items() if v[1]})
