"""ncu target: warm the stepper up (5 control steps so contacts/warm starts are in steady state), then run a few
substeps.  Use with `ncu --launch-skip 357 --launch-count 7` to capture exactly one steady-state substep."""
import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from flybody_b200.flymodel import load_model
from flybody_b200 import stepper as st
from conftest import walk_reset_qpos
m = load_model(os.environ.get('FB_VARIANT', 'walk')); N = int(os.environ.get('FB_N', 4096)); rs = np.random.RandomState(0)
s = st.BatchedStepper(m, N)
q0 = walk_reset_qpos(m) if m.nq == 116 else m.qpos0
qq = np.tile(q0, (N, 1))
if m.nq == 116:
    qq[:, 7:109] += rs.uniform(-0.05, 0.05, (N, 102))
l0 = s.launch_count
s.reset(qq)
print('launches in reset:', s.launch_count - l0)
for k in range(5):
    s.set_control(rs.uniform(-0.5, 0.5, (N, m.nu)).astype(np.float32))
    s.step(10)
print('launches before capture:', s.launch_count - l0)
s.set_control(rs.uniform(-0.5, 0.5, (N, m.nu)).astype(np.float32))
s.step(2)
s.sync()
print('done', s.launch_count - l0)
