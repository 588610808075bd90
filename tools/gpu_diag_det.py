"""Diagnostic: find the first stage/field where a shuffled batch differs from the shuffled result."""
import numpy as np, sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from flybody_b200 import stepper as st
from flybody_b200.flymodel import load_model
from parity_common import reset_qpos
m = load_model('walk'); N = 4096
rs = np.random.RandomState(0)
q0 = reset_qpos(m)
qq = np.tile(q0, (N, 1)); qq[:, 7:109] += rs.uniform(-0.05, 0.05, (N, 102))
perm = rs.permutation(N)
c = rs.uniform(-0.5, 0.5, (N, m.nu)).astype(np.float32)
names = ['QPOS','QVEL','ACT','QACC','SENSORDATA','NCON','NEFC','NITER','EFC_FORCE','QFRC_CONSTRAINT','QFRC_ACTUATOR','QFRC_BIAS','QFRC_PASSIVE','XPOS','CVEL']
def cmp(tag, s, s2):
    for nm in names:
        if not hasattr(st, nm): continue
        a = s.get(getattr(st, nm))[perm]; b = s2.get(getattr(st, nm))
        if a.shape != b.shape: continue
        bad = np.argwhere(a != b)
        if len(bad): print(tag, nm, 'differs in', len(np.unique(bad[:,0])), 'envs; first', bad[0], a[tuple(bad[0])], b[tuple(bad[0])])
for pre in (0, 1):
    s = st.BatchedStepper(m, N); s2 = st.BatchedStepper(m, N)
    if pre:
        s.reset(qq)
        for k in range(3): s.set_control(rs.uniform(-0.5, 0.5, (N, m.nu))); s.step(10)
    s.reset(qq); s2.reset(qq[perm]); cmp(f'pre{pre} reset', s, s2)
    s.set_control(c); s2.set_control(c[perm])
    for k in range(10):
        s.step(1); s2.step(1); cmp(f'pre{pre} sub{k}', s, s2)
print('done')
