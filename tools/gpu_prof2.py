import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
from flybody_b200.flymodel import load_model
from flybody_b200 import stepper as st
from conftest import walk_reset_qpos
m = load_model('walk'); N=4096; rs=np.random.RandomState(0)
s = st.BatchedStepper(m, N)
q0 = walk_reset_qpos(m); qq = np.tile(q0,(N,1)); qq[:,7:109] += rs.uniform(-0.05,0.05,(N,102)); s.reset(qq)
for it in range(4):
    s.set_control(rs.uniform(-0.5,0.5,(N,m.nu)).astype(np.float32)); s.step(10); s.sync()
s.profile(True)
for it in range(5):
    s.set_control(rs.uniform(-0.5,0.5,(N,m.nu)).astype(np.float32)); s.step(10); s.sync()
p = s.profile_read()
print(os.environ.get('FB_SOLVE_SMEM_KB'), 'ms/step', s.last_step_ms, {k: round(v[0]/5,2) for k,v in p.items() if v[1]})
nefc = s.get(st.NEFC)[:,0]; print('nefc mean', nefc.mean(), 'max', nefc.max(), 'block-max mean', nefc.reshape(-1,32).max(1).mean(), 'niter mean', s.get(st.SOLVER_NITER).mean(), 'blockmax', s.get(st.SOLVER_NITER)[:,0].reshape(-1,32).max(1).mean())
