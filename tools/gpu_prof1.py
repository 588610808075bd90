import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
from flybody_b200.flymodel import load_model
from flybody_b200 import stepper as st
from conftest import walk_reset_qpos
m = load_model('walk'); N=int(os.environ.get('FB_N','4096')); rs=np.random.RandomState(0)
s = st.BatchedStepper(m, N)
q0 = walk_reset_qpos(m); qq = np.tile(q0,(N,1)); qq[:,7:109] += rs.uniform(-0.05,0.05,(N,102)); s.reset(qq)
for it in range(int(os.environ.get('FB_STEPS','3'))):
    s.set_control(rs.uniform(-0.5,0.5,(N,m.nu)).astype(np.float32)); s.step(10); s.sync()
print('ms', s.last_step_ms)
