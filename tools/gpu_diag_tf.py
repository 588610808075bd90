import numpy as np, sys
import os; sys.path.insert(0,os.getcwd()); sys.path.insert(0,os.path.join(os.getcwd(),'tests'))
from flybody_b200.flymodel import load_model
from flybody_b200 import stepper as st
from oracle import fly_oracle as fo
from parity_common import reset_qpos
m = load_model('walk')
sim = st.BatchedStepper(m, 1, lib_path=None)
o = fo.Oracle(m, tolerance=1e-12)
q0 = reset_qpos(m); o.reset(q0); sim.reset(q0)
rs = np.random.RandomState(0)
for k in range(200):
    sim.set(st.QPOS, o.qpos); sim.set(st.QVEL, o.qvel); sim.set(st.ACT, o.get(fo.ACT)); sim.set(st.QACC_WARMSTART, o.get(fo.QACC_WARMSTART)); sim.forward()
    ctrl = rs.uniform(-0.5,0.5,m.nu); sim.set_control(ctrl); o.set(fo.CTRL, ctrl)
    # substep-by-substep
    worst=0
    for ss in range(10):
        o.control_step(1); sim.step(1)
        ev = np.abs(sim.get(st.QVEL)[0]-o.qvel).max()
        nco, nce = int(o.get(fo.NCON)[0]), int(sim.get(st.NCON)[0,0]); nfo,nfe = int(o.get(fo.NEFC)[0]), int(sim.get(st.NEFC)[0,0])
        if ev>2e-3 and worst==0:
            worst=1
            print(f'step {k} sub {ss}: qvel err {ev:.2e} argmax dof {np.abs(sim.get(st.QVEL)[0]-o.qvel).argmax()} ncon {nco}/{nce} nefc {nfo}/{nfe} niter {o.get(fo.SOLVER_NITER)[0]}/{sim.get(st.SOLVER_NITER)[0,0]}')
            co = o.get(fo.CONTACT).reshape(-1,16); ce = sim.get(st.CONTACT)[0].reshape(-1,16)
            print('   oracle contacts (dist,g1,g2,incl):', [(round(c[0],7),int(c[7]),int(c[8]),int(c[10])) for c in co[:nco]])
            print('   emu    contacts (dist,g1,g2,incl):', [(round(float(c[0]),7),int(c[7]),int(c[8]),int(c[10])) for c in ce[:nce]])
