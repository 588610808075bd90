"""Golden heightfields from the reference's own terrain functions (`flybody/tasks/arenas/hills.py`), loaded by file path with
dm_control stubbed out (the pure functions need only numpy / scipy).  A small arena keeps the fixture small.
    python tests/golden/make_terrain_goldens.py        -> tests/golden/terrain_goldens.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = '/root/reference/flybody'
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    for name in ('dm_control', 'dm_control.composer', 'dm_control.locomotion', 'dm_control.locomotion.arenas', 'dm_control.locomotion.arenas.assets',
                 'dm_control.mujoco', 'dm_control.mujoco.wrapper', 'dm_control.mujoco.wrapper.mjbindings'):
        sys.modules[name] = types.ModuleType(name)
    sys.modules['dm_control'].composer = sys.modules['dm_control.composer']
    sys.modules['dm_control.composer'].Arena = object
    sys.modules['dm_control.locomotion.arenas'].assets = sys.modules['dm_control.locomotion.arenas.assets']
    sys.modules['dm_control.mujoco.wrapper'].mjbindings = sys.modules['dm_control.mujoco.wrapper.mjbindings']
    sys.modules['dm_control.mujoco.wrapper.mjbindings'].mjlib = None
    spec = importlib.util.spec_from_file_location('hills', f'{REF}/tasks/arenas/hills.py')
    hills = importlib.util.module_from_spec(spec); spec.loader.exec_module(hills)

    dim, dens = 6, 5
    nrow = ncol = ((2 * dens * dim) // 2) * 2 + 1

    class FakeModel:
        hfield_size = np.array([[dim, dim, 1.0, 0.05]]); hfield_nrow = np.array([nrow]); hfield_ncol = np.array([ncol])

    class FakePhysics:
        model = FakeModel()
    g = {'dim': np.array(dim), 'grid_density': np.array(dens)}
    g['bowl'] = hills.terrain_bowl(FakePhysics(), elevation_z=4.3, random_state=np.random.RandomState(3))
    size = FakeModel.hfield_size[0, :2]
    g['bumps'] = hills.add_sine_bumps(g['bowl'], size, wavelength=3.7, phase=0.9, height=0.8)
    t, s = hills.add_sine_trench(g['bowl'], size, wavelength=2.5, phase=1.1, amplitude=0.45, start_x=-2.0, end_x=2.5, width=1.3, height=1.3, sigma=0.2)
    g['trench'], g['trench_sine'] = t, s
    pts = np.array([[0.0, 0.0], [-5.9, 2.2], [3.3, -4.4], [5.99, 5.99], [-6.0, -6.0]])
    g['idx_points'] = pts
    g['idx'] = np.array([hills.pos_to_terrain_idx(x, y, size, nrow, ncol) for x, y in pts])
    np.savez_compressed(os.path.join(OUT, 'terrain_goldens.npz'), **g)
    print({k: v.shape for k, v in g.items()})


if __name__ == '__main__':
    main()
