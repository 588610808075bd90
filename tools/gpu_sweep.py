"""Throughput of the stepper (device-resident random controls, CUDA graph replay) over batch sizes and model variants."""
import numpy as np, sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import torch
from flybody_b200.flymodel import load_model
from flybody_b200 import stepper as st
from conftest import walk_reset_qpos

def run(variant, N, nsub, steps=20, scale=0.5):
    m = load_model(variant)
    s = st.BatchedStepper(m, N)
    rs = np.random.RandomState(0)
    if variant == 'walk':
        q = np.tile(walk_reset_qpos(m), (N, 1)); q[:, 7:109] += rs.uniform(-0.05, 0.05, (N, 102))
        s.reset(q)
    else:
        q = np.tile(m.qpos0, (N, 1)); q[:, 2] = 1.0
        s.reset(q)
    acts = (torch.rand((steps + 5, N, m.nu), device='cuda') - 0.5) * 2 * scale
    for k in range(5):
        s.set_control_device(acts[k].data_ptr()); s.step(nsub)
    s.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(5, steps + 5):
        s.set_control_device(acts[k].data_ptr()); s.step(nsub)
    s.sync()
    dt = (time.perf_counter() - t0) / steps
    bad = int((s.get(st.FLAGS)[:, 0] != 0).sum())
    print(f'{variant:7s} N={N:6d}  {dt * 1e3:8.3f} ms/control step  {N / dt:12.0f} env-steps/s  flagged {bad}', flush=True)
    s.close()

for variant, N, nsub, scale in [('walk', 256, 10, 0.5), ('walk', 1024, 10, 0.5), ('walk', 4096, 10, 0.5), ('walk', 8192, 10, 0.5), ('walk', 16384, 10, 0.5),
                                ('flight', 4096, 4, 0.2), ('flight', 16384, 4, 0.2)]:
    run(variant, N, nsub, scale=scale)
