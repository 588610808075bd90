"""A/B of the launch groupings (FB_FUSE=0..3): ms per control step of walk_imitation, graph replay on, no profiling.
Each mode runs in its own process (the mode is read at fb_create):  python tools/gpu_fuse_ab.py [n_envs] [steps]"""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))


def one(n_envs, steps):
    from flybody_b200.flymodel import load_model
    from flybody_b200 import stepper as st
    from conftest import walk_reset_qpos
    m = load_model('walk'); rs = np.random.RandomState(0)
    s = st.BatchedStepper(m, n_envs)
    qq = np.tile(walk_reset_qpos(m), (n_envs, 1)); qq[:, 7:109] += rs.uniform(-0.05, 0.05, (n_envs, 102)); s.reset(qq)
    ctrl = [rs.uniform(-0.5, 0.5, (n_envs, m.nu)).astype(np.float32) for _ in range(4)]
    for it in range(6):
        s.set_control(ctrl[it % 4]); s.step(10)
    s.sync()
    dev = []
    t0 = time.perf_counter()
    for it in range(steps):
        s.set_control(ctrl[it % 4]); s.step(10); s.sync(); dev.append(s.last_step_ms)
    wall = (time.perf_counter() - t0) / steps * 1e3
    q = s.get(st.QPOS)
    print(f"FB_FUSE={os.environ.get('FB_FUSE')} FB_SPLIT={os.environ.get('FB_SPLIT')} FB_STAGGER={os.environ.get('FB_STAGGER')} n={n_envs} device ms/step median {np.median(dev):.3f} min {np.min(dev):.3f} wall {wall:.3f} "
          f"env-steps/s {n_envs / np.median(dev) * 1e3:.0f} checksum {float(np.abs(q).sum()):.6f} finite {bool(np.isfinite(q).all())}", flush=True)


if __name__ == '__main__':
    if os.environ.get('FB_AB_CHILD'):
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        n = sys.argv[1] if len(sys.argv) > 1 else '4096'; k = sys.argv[2] if len(sys.argv) > 2 else '20'
        # variants: ';'-separated lists of VAR=VALUE settings, e.g. FB_AB_VARIANTS="FB_SPLIT=1;FB_SPLIT=2 FB_STAGGER=3"
        variants = os.environ.get('FB_AB_VARIANTS')
        variants = [v.split() for v in variants.split(';')] if variants else [[f'FB_FUSE={m}'] for m in '0123']
        for v in variants:
            env = dict(os.environ, FB_AB_CHILD='1', **dict(x.split('=') for x in v))
            print('==', ' '.join(v), flush=True)
            subprocess.run([sys.executable, __file__, n, k], env=env, timeout=300)
