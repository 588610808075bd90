import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
from flybody_b200.flymodel import load_model
from flybody_b200 import stepper as st
from conftest import walk_reset_qpos
m = load_model('walk'); N=4096; rs=np.random.RandomState(0)
s = st.BatchedStepper(m, N)
q0 = walk_reset_qpos(m); qq = np.tile(q0,(N,1)); qq[:,7:109] += rs.uniform(-0.05,0.05,(N,102)); s.reset(qq)
s.profile(True)
for it in range(3):
    s.forward()
p = s.profile_read()
print(os.environ.get('FB_POS_TRUNC'), {k: round(v[0]/v[1]*1e3,1) for k,v in p.items() if v[1]})
