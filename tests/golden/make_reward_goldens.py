"""Golden vectors for the walking-imitation reward from the reference's own pure-Python functions
(`flybody/tasks/rewards.py`, `flybody/quaternions.py`; they import only numpy).  Modules are loaded by file path
because the package __init__ pulls in dm_control.  Run in the build container:
    python tests/golden/make_reward_goldens.py        -> tests/golden/reward_goldens.npz
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
    quats = load('flybody.quaternions', f'{REF}/quaternions.py')
    sys.modules['flybody'].quaternions = quats
    rw = load('flybody.tasks.rewards', f'{REF}/tasks/rewards.py')
    rs = np.random.RandomState(11)
    g = {}
    n_env, nj, ns, T = 5, 9, 4, 7
    # joint orientation quaternions, with the z-axis edge cases of quat_z2vec
    axes = rs.normal(size=(12, 3))
    axes[0] = [0, 0, 1.0]; axes[1] = [0, 0, -2.0]; axes[2] = [0, 0, 0]
    ang = rs.uniform(-2, 2, 12)
    g['axes'], g['angles'] = axes, ang
    g['quat_z2vec'] = quats.quat_z2vec(axes)
    g['joint_orientation_quat'] = quats.joint_orientation_quat(axes[3:], ang[3:])
    # egocentric vectors
    rootp = rs.normal(size=(n_env, 1, 3)); rootq = rs.normal(size=(n_env, 1, 4)); rootq /= np.linalg.norm(rootq, axis=-1, keepdims=True)
    sites = rs.normal(size=(n_env, ns, 3))
    g['root_pos'], g['root_quat'], g['sites'] = rootp[:, 0], rootq[:, 0], sites
    g['egocentric'] = np.array([quats.get_egocentric_vec(rootp[e, 0], sites[e], rootq[e, 0]) for e in range(n_env)])
    # a synthetic walking snippet and per-env walker states -> reward factors, one env at a time as the reference does
    snippet = {'qpos': rs.normal(size=(T, 7 + nj)), 'qvel': rs.normal(size=(T, 6 + nj)) * 20,
               'root2site': rs.normal(size=(T, ns, 3)) * 0.1, 'joint_quat': rs.normal(size=(T, nj, 4))}
    snippet['qpos'][:, 3:7] /= np.linalg.norm(snippet['qpos'][:, 3:7], axis=1, keepdims=True)
    snippet['joint_quat'] /= np.linalg.norm(snippet['joint_quat'], axis=-1, keepdims=True)
    steps = np.array([0, 3, 6, 2, 5])
    wq = snippet['qpos'][steps] + rs.normal(size=(n_env, 7 + nj)) * 0.05
    wv = snippet['qvel'][steps] + rs.normal(size=(n_env, 6 + nj)) * 5
    wsite = snippet['root2site'][steps] + rs.normal(size=(n_env, ns, 3)) * 0.02
    waxes = rs.normal(size=(n_env, nj, 3))                      # joint axes already in the root frame
    factors, diffs = [], []
    for e in range(n_env):
        jq = quats.joint_orientation_quat(waxes[e], wq[e, 7:])
        wf = {'com': wq[e, :3], 'qvel': wv[e], 'root2site': wsite[e], 'joint_quat': np.vstack((wq[e, 3:7], jq))}
        rf = rw.get_reference_features(snippet, int(steps[e]))
        factors.append(rw.reward_factors_deep_mimic(walker_features=wf, reference_features=rf, weights=(20, 1, 1, 1)))
        d = rw.compute_diffs(wf, rf, n=2)
        diffs.append([d[k] for k in ('com', 'qvel', 'root2site', 'joint_quat')])
    for k, v in snippet.items():
        g['snippet_' + k] = v
    g['steps'], g['walker_qpos'], g['walker_qvel'], g['walker_root2site'], g['walker_axes_ego'] = steps, wq, wv, wsite, waxes
    g['deep_mimic_factors'] = np.array(factors)                 # [n_env, 4]
    g['deep_mimic_diffs'] = np.array(diffs)
    np.savez_compressed(os.path.join(OUT, 'reward_goldens.npz'), **g)
    print({k: v.shape for k, v in g.items()})


if __name__ == '__main__':
    main()
