"""Single-warp latency profile: build with -DFB_CLK (see _variants/lib_clk.so), step a small batch and print the cycles
warp 0 spends in every stage of every kernel of one steady-state substep."""
import ctypes as C, numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
os.environ['FB_NO_GRAPH'] = '1'
from flybody_b200.flymodel import load_model
from flybody_b200 import stepper as st
from conftest import walk_reset_qpos
lib = os.path.join(os.getcwd(), '_variants', 'lib_clk.so')
N = int(os.environ.get('FB_N', 32))
m = load_model('walk'); s = st.BatchedStepper(m, N, lib_path=lib); rs = np.random.RandomState(0)
q = np.tile(walk_reset_qpos(m), (N, 1)); q[:, 7:109] += rs.uniform(-0.05, 0.05, (N, 102)); s.reset(q)
for k in range(4):
    s.set_control(rs.uniform(-0.5, 0.5, (N, m.nu)).astype(np.float32)); s.step(10)
s.set_control(rs.uniform(-0.5, 0.5, (N, m.nu)).astype(np.float32)); s.sync()
l0 = s.launch_count
s.step(2); s.sync()
buf = np.zeros(32 * 4096, np.int64)
s._lib.fb_clk_read.argtypes = [C.c_void_p, C.c_void_p]
s._lib.fb_clk_read(s._h, buf.ctypes.data)
buf = buf.reshape(4096, 32)
names = ['smooth', 'solve', 'finish', 'pos', 'col', 'proj', 'vel']
stages = {'smooth': ['act0', 'act1', 'adh', 'adh_b', 'qfrc', 'L^-T', 'out', 'kref'], 'solve': ['solve'],
          'finish': ['f1', 'L^-1', 'f5', 'f6', 'f7', 'euler', 'f8', 'f9'], 'pos': ['p0', 'p1 kin', 'p1b', 'p2 crb', 'p3', 'p4 M', 'factor', 'wr', 'reinit', 'factor2', 'wr2'],
          'col': ['stage', 'broad', 'narrow', 'mpr', 'compact'], 'proj': ['c0', 'c1', 'c2', 'c3', 'J,Z', 'A'], 'vel': ['v0', 'v1', 'v1b', 'v2', 'v3', 'v3b', 'v4']}
tot = 0
for j in range(7):                       # second substep of the step(2) call: launches l0+7 .. l0+13
    row = buf[(l0 + 7 + j) % 4096]
    n = len(stages[names[j]])
    dts = np.diff(row[:n + 1])
    tot += dts.sum()
    print(f'{names[j]:7s} {dts.sum():8d} cyc  ' + '  '.join(f'{a}:{int(b)}' for a, b in zip(stages[names[j]], dts)))
print('sum of in-kernel cycles of warp 0:', tot, '=', tot / 1.965e3, 'us per substep')
