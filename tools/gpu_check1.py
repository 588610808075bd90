import numpy as np, time, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
from flybody_b200.flymodel import load_model
from flybody_b200 import stepper as st
from oracle import fly_oracle as fo
from conftest import walk_reset_qpos
try:
    import mujoco; print('MUJOCO AVAILABLE', mujoco.__version__)
except Exception as e: print('no mujoco:', e)
try:
    import dm_control; print('DM_CONTROL AVAILABLE')
except Exception as e: print('no dm_control:', e)
os.system('nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv; nproc; lscpu | grep "Model name"')
m = load_model('walk')
N=64
s = st.BatchedStepper(m, N)
print(s.version())
o = fo.Oracle(m, tolerance=1e-12)
q0 = walk_reset_qpos(m)
rs = np.random.RandomState(0)
q = q0.copy(); q[7:109] += rs.uniform(-0.1,0.1,102)
v = rs.uniform(-1,1,m.nv)
o.reset(q, v); s.reset(q, v)
def cmp(name, f):
    a = o.get(f); b = s.get(f)[0].astype(np.float64); b2 = s.get(f)[N-1].astype(np.float64)
    n=min(len(a),len(b)); err = np.abs(a[:n]-b[:n]).max(); ref = np.abs(a[:n]).max()
    print(f'{name:18s} err {err:.3e} ref {ref:.3e} rel {err/(ref+1e-30):.2e}  env0-vs-envN {np.abs(b-b2).max():.1e}')
for nm,f in [('xpos',fo.XPOS),('M',fo.QM_DENSE),('bias',fo.QFRC_BIAS),('passive',fo.QFRC_PASSIVE),('smooth',fo.QFRC_SMOOTH),('efc_force',fo.EFC_FORCE),('qfrc_con',fo.QFRC_CONSTRAINT),('qacc',fo.QACC),('sensordata',fo.SENSORDATA)]: cmp(nm,f)
print('ncon', o.get(fo.NCON), s.get(st.NCON)[:2,0], 'nefc', o.get(fo.NEFC), s.get(st.NEFC)[:2,0], 'flags', s.get(st.FLAGS)[:2,0])
# teacher-forced steps
o.reset(q0); s.reset(q0)
acts = rs.uniform(-0.5,0.5,(40,m.nu)); errs=[]
for k in range(40):
    s.set(st.QPOS,o.qpos); s.set(st.QVEL,o.qvel); s.set(st.ACT,o.get(fo.ACT)); s.set(st.QACC_WARMSTART,o.get(fo.QACC_WARMSTART)); s.forward()
    s.set_control(acts[k]); o.set(fo.CTRL,acts[k]); o.control_step(10); s.step(10)
    errs.append((np.abs(s.get(st.QPOS)[0][:109]-o.qpos[:109]).max(), np.abs(s.get(st.QVEL)[0][:108]-o.qvel[:108]).max()))
e=np.array(errs); print('teacher-forced 40 control steps: max qpos err %.2e qvel err %.2e'%(e[:,0].max(), e[:,1].max()))
# timing
for N in (4096,):
    s = st.BatchedStepper(m, N)
    qq = np.tile(q0,(N,1)); qq[:,7:109] += rs.uniform(-0.05,0.05,(N,102))
    s.reset(qq)
    A = rs.uniform(-0.5,0.5,(N,m.nu)).astype(np.float32)
    for it in range(5):
        s.set_control(A); s.step(10); s.sync()
    t=time.time(); ms=[]
    for it in range(10):
        s.set_control(rs.uniform(-0.5,0.5,(N,m.nu)).astype(np.float32)); s.step(10); s.sync(); ms.append(s.last_step_ms)
    dt=time.time()-t
    print('N',N,'ms/control step (events)', np.mean(ms), 'wall', dt/10*1e3, 'env-steps/s', N/np.mean(ms)*1e3, 'flags nonzero', (s.get(st.FLAGS)[:,0]!=0).sum(), 'mean ncon', s.get(st.NCON).mean(), 'mean nefc', s.get(st.NEFC).mean(), 'niter', s.get(st.SOLVER_NITER).mean())
