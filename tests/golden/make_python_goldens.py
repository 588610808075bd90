"""Golden vectors from the reference's *pure-Python* task helpers (they import only numpy, so they run here
even though mujoco / dm_control do not):

  flybody/tasks/pattern_generators.py   WingBeatPatternGenerator   -> wbpg_* arrays
  flybody/tasks/task_utils.py           com2root / root2com        -> com2root_*, root2com_*
  flybody/quaternions.py                get_dquat_local, quat_dist_short_arc, rotate_vec_with_quat

The modules are loaded by file path (the package __init__ pulls in dm_control).  Run in the build container:
    python tests/golden/make_python_goldens.py        -> tests/golden/python_goldens.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = '/root/reference/flybody'
OUT = os.path.dirname(os.path.abspath(__file__))


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    for pkg in ('flybody', 'flybody.tasks'):
        sys.modules[pkg] = types.ModuleType(pkg)
        sys.modules[pkg].__path__ = []
    load('flybody.tasks.constants', f'{REF}/tasks/constants.py')
    quats = load('flybody.quaternions', f'{REF}/quaternions.py')
    pg = load('flybody.tasks.pattern_generators', f'{REF}/tasks/pattern_generators.py')
    tu = load('flybody.tasks.task_utils', f'{REF}/tasks/task_utils.py')

    g = {}
    rs = np.random.RandomState(7)
    # --- wing beat pattern generator: 4 independent generators, 300 control steps of varying requested frequency
    phases = np.array([0.0, 0.3, 0.62, 0.97])
    acts = rs.uniform(-1, 1, (300, 4))
    acts[100:150] = 1.0
    acts[150:200] = -1.0
    out_q, out_v, out_seq = [], [], []
    for k in range(4):
        w = pg.WingBeatPatternGenerator()
        q, v = w.reset(initial_phase=phases[k], return_qvel=True)
        out_q.append(q.copy()); out_v.append(v.copy())
        seq = []
        for t in range(300):
            seq.append(w.step(ctrl_freq=w.base_beat_freq * (1 + w.rel_freq_range * acts[t, k])).copy())
        out_seq.append(np.array(seq))
    g['wbpg_phases'] = phases
    g['wbpg_actions'] = acts
    g['wbpg_reset_qpos'] = np.array(out_q)
    g['wbpg_reset_qvel'] = np.array(out_v)
    g['wbpg_steps'] = np.array(out_seq)            # [4, 300, 6]
    # --- CoM <-> root
    q = rs.normal(size=(16, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    com = rs.normal(size=(16, 3))
    g['quat'] = q
    g['com'] = com
    g['com2root'] = tu.com2root(com, q)
    g['root2com'] = np.array([tu.root2com(np.concatenate([com[i], q[i]])) for i in range(16)])
    # --- quaternion helpers used by the observables / rewards
    q2 = rs.normal(size=(16, 4)); q2 /= np.linalg.norm(q2, axis=1, keepdims=True)
    g['quat2'] = q2
    g['dquat_local'] = quats.get_dquat_local(q, q2)
    g['quat_dist_short_arc'] = quats.quat_dist_short_arc(q, q2)
    g['rotate_vec'] = quats.rotate_vec_with_quat(com, q)
    # --- ellipsoid fluid model, local-frame forces (the reference's Python restatement of engine_passive.c)
    for pkg in ('dm_control', 'dm_control.mujoco', 'dm_control.mjcf'):
        sys.modules[pkg] = types.ModuleType(pkg)
    sys.modules['dm_control'].mujoco = sys.modules['dm_control.mujoco']
    sys.modules['dm_control'].mjcf = sys.modules['dm_control.mjcf']
    sys.modules['dm_control.mjcf'].Physics = object
    efm = load('flybody.ellipsoid_fluid_model', f'{REF}/ellipsoid_fluid_model.py')
    size = np.array([0.0005, 0.0551, 0.114])                 # wing fluid geom, fruitfly.xml:388,399
    coefs = np.array([1.0, 0.5, 1.5, 1.7, 1.0])              # blunt, slender, angular, kutta, magnus (constants.py:28)
    vmass = np.array([2.9e-3, 1.1e-5, 3.0e-6])               # arbitrary positive virtual mass / inertia
    vinert = np.array([4.0e-6, 7.0e-9, 2.0e-8])
    rho, eta = 0.00128, 0.000185
    vels = rs.normal(size=(32, 6)) * np.array([300, 300, 300, 60, 60, 60])
    vels[0] = 0.0
    vels[1, 3:] = 0.0
    out = []
    for lv in vels:
        f = np.zeros(6)
        efm.mj_addedMassForces(lv, None, rho, vmass, vinert, f)
        efm.mj_viscousForces(lv, rho, eta, size, coefs[4], coefs[3], coefs[0], coefs[1], coefs[2], f)
        out.append(f)
    g['fluid_size'], g['fluid_coefs'], g['fluid_vmass'], g['fluid_vinert'] = size, coefs, vmass, vinert
    g['fluid_density_viscosity'] = np.array([rho, eta])
    g['fluid_local_vels'] = vels                             # [angular; linear], geom frame
    g['fluid_local_force'] = np.array(out)                   # [torque; force]
    np.savez_compressed(os.path.join(OUT, 'python_goldens.npz'), **g)
    print({k: v.shape for k, v in g.items()})


if __name__ == '__main__':
    main()
