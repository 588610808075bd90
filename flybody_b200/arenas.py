"""Terrain generators of the vision-guided flight arenas (reference `flybody/tasks/arenas/hills.py:13-130, 259-281, 333-392,
442-473`): a bowl of smooth random bumps rising to "horizon mountains", with sinusoidal ridges (`SineBumps`) or a smoothed
sinusoidal trench (`SineTrench`) on top.  GROUNDWORK for SURVEY.md 8(f).1 (`vision_guided_flight`): the heightfield collision
and the eye ray-caster that consume these heightfields are not built yet; what is here is pinned against the reference's own
functions (`tests/golden/make_terrain_goldens.py`).

A terrain is a float array [nrow, ncol] of heights in world units (the reference stores it as MuJoCo `hfield_data` with
elevation scale 1, `hills.py:170-176`); row index = y, column index = x, both spanning [-size, +size].
"""
import numpy as np
from scipy import ndimage


def grid_shape(dim=20, grid_density=10):
    """(nrow, ncol) of the arena heightfield (reference `hills.py:172-175`): an odd number of points per axis."""
    size = dim if isinstance(dim, tuple) else (dim, dim)
    dens = grid_density if isinstance(grid_density, tuple) else (grid_density, grid_density)
    return ((2 * dens[1] * size[1]) // 2) * 2 + 1, ((2 * dens[0] * size[0]) // 2) * 2 + 1


def pos_to_terrain_idx(x, y, size, nrow, ncol):
    """(idx_x, idx_y) of world position (x, y) (reference `hills.py:13-17`; note size[0] scales y and size[1] scales x)."""
    idx_y = int((y / size[0]) * (nrow / 2) + nrow / 2)
    idx_x = int((x / size[1]) * (ncol / 2) + ncol / 2)
    return idx_x, idx_y


def terrain_bowl(size, nrow, ncol, bump_scale=2.0, elevation_z=4.0, tanh_rel_radius=0.7, tanh_sharpness=8.0, random_state=None):
    """reference `hills.py:20-63`: uniform random bumps (one per bump_scale length units), spline-zoomed to the grid,
    normalised to [0, elevation_z], multiplied by a tanh bowl that is 0 in the middle and 1 at the rim."""
    assert nrow == ncol
    bump_res = int(2 * size[0] / bump_scale)
    bumps = random_state.uniform(0, 1, (bump_res, bump_res))
    terrain = ndimage.zoom(bumps, nrow / float(bump_res))
    terrain -= np.min(terrain)
    terrain /= np.max(terrain)
    terrain *= elevation_z
    axis = np.linspace(-1, 1, terrain.shape[0])
    xv, yv = np.meshgrid(axis, axis)
    r = np.sqrt(xv ** 2 + yv ** 2)
    return terrain * (0.5 * np.tanh(tanh_sharpness * (r - tanh_rel_radius)) + 0.5)


def add_sine_bumps(terrain, arena_size, wavelength=5.0, phase=0.0, height=1.0):
    """reference `hills.py:66-86`: ridges along y, max-combined with the terrain."""
    ncol = terrain.shape[1]
    x_axis = np.linspace(-arena_size[0], arena_size[0], ncol)
    bumps = height * 0.5 * (np.sin(2 * np.pi / wavelength * x_axis + phase) + 1)
    return np.maximum(bumps[None, :], terrain)


def add_sine_trench(terrain, arena_size, wavelength=5, phase=0.0, amplitude=1.0, start_x=0, end_x=10.0, width=1.0, height=1.0,
                    sigma=0.2):
    """reference `hills.py:89-130`: a plateau of `height` between start_x and end_x with a sinusoidal corridor of `width` cut
    into it, Gaussian-smoothed, max-combined with the terrain.  Returns (terrain, sine) with sine the corridor centre line."""
    nrow, ncol = terrain.shape
    idx_from, _ = pos_to_terrain_idx(start_x, 0, arena_size, nrow, ncol)
    idx_to, _ = pos_to_terrain_idx(end_x, 0, arena_size, nrow, ncol)
    delta, _ = pos_to_terrain_idx(-arena_size[0] + width / 2, 0, arena_size, nrow, ncol)
    x_axis = np.linspace(0, end_x - start_x, idx_to - idx_from + 1)
    sine = amplitude * np.sin(2 * np.pi / wavelength * x_axis + phase)
    sine -= sine[0]
    trench = np.zeros_like(terrain)
    trench[:, idx_from:idx_to] = height
    for idx_x in range(idx_from, idx_to):
        _, idx_y = pos_to_terrain_idx(0, sine[idx_x - idx_from], arena_size, nrow, ncol)
        trench[idx_y - delta:idx_y + delta + 1, idx_x] = 0.0
    trench = ndimage.gaussian_filter(trench, sigma=sigma / width * delta * 2)
    return np.maximum(trench, terrain), sine


def hfield_height(terrain, x, y, half_size):
    """height at the grid point nearest to (x, y) (reference `tasks/vision_flight.py:81-95`: argmin of |axis - x| over the grid axis
    `linspace(-half, half, n)`, first minimum on ties; x, y arrays broadcast) -- computed in closed form instead of an [N, n] table."""
    ncol = terrain.shape[-1]
    step = 2.0 * half_size / (ncol - 1)
    nearest = lambda v: np.clip(np.ceil((np.asarray(v, np.float64).reshape(-1) + half_size) / step - 0.5), 0, ncol - 1).astype(np.int64)
    xi, yi = nearest(x), nearest(y)
    return terrain[..., yi, xi] if terrain.ndim == 2 else terrain[np.arange(len(xi)), yi, xi]


class SineBumps:
    """`hills.py:395-473`: per-episode parameters drawn in the reference's order (elevation, bowl bumps, wavelength, phase, height)."""

    def __init__(self, dim=20, grid_density=10, elevation_z_range=(4.0, 5.0), phase_range=(0.0, 2 * np.pi),
                 wavelength_range=(10.0, 15.0), height_range=(0.5, 1.0)):
        self.size = (dim, dim)
        self.nrow, self.ncol = grid_shape(dim, grid_density)
        self._elev, self._phase, self._wl, self._h = elevation_z_range, phase_range, wavelength_range, height_range
        self.trench_specs = None

    def generate(self, random_state):
        elevation_z = random_state.uniform(*self._elev)
        bowl = terrain_bowl(self.size, self.nrow, self.ncol, elevation_z=elevation_z, random_state=random_state)
        return add_sine_bumps(bowl, self.size, wavelength=random_state.uniform(*self._wl), phase=random_state.uniform(*self._phase),
                              height=random_state.uniform(*self._h))


class SineTrench:
    """`hills.py:284-392`: per-episode parameters in the reference's order (elevation, bowl bumps, start, length, amplitude,
    width factor, phase, wavelength, height, sigma); `trench_specs` holds the corridor centre line for the reward."""

    def __init__(self, dim=20, grid_density=10, elevation_z_range=(4.0, 5.0), start_offset_range=(-5.0, -3.0), trench_len_range=(4.0, 10.0),
                 phase_range=(0.0, 2 * np.pi), wavelength_range=(5.0, 8.0), amplitude_range=(0.35, 0.6), width_range=(0.5, 1.0),
                 height_range=(1.3, 1.3), sigma_range=(0.2, 0.2)):
        self.size = (dim, dim)
        self.nrow, self.ncol = grid_shape(dim, grid_density)
        self._r = dict(elev=elevation_z_range, start=start_offset_range, length=trench_len_range, phase=phase_range, wl=wavelength_range,
                       amp=amplitude_range, width=width_range, height=height_range, sigma=sigma_range)
        self.trench_specs = None

    def generate(self, random_state):
        r = self._r
        elevation_z = random_state.uniform(*r['elev'])
        bowl = terrain_bowl(self.size, self.nrow, self.ncol, elevation_z=elevation_z, random_state=random_state)
        start_x = random_state.uniform(*r['start'])
        end_x = start_x + random_state.uniform(*r['length'])
        amplitude = random_state.uniform(*r['amp'])
        width = 2 * amplitude + 0.604 * random_state.uniform(*r['width'])          # 0.604 = wing span: no straight fly-through
        terrain, sine = add_sine_trench(bowl, self.size, start_x=start_x, end_x=end_x, phase=random_state.uniform(*r['phase']),
                                        wavelength=random_state.uniform(*r['wl']), amplitude=amplitude, width=width,
                                        height=random_state.uniform(*r['height']), sigma=random_state.uniform(*r['sigma']))
        self.trench_specs = {'x_coords': np.linspace(start_x, end_x, sine.shape[0]), 'y_coords': sine}
        return terrain
