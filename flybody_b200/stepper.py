"""ctypes binding of the CUDA stepper's C ABI (`include/flybody_b200.h`).

`BatchedStepper` is the batched stand-in for the `physics` object that dm_control hands to the
reference's task hooks: `set_control`, `step`, state reads.  It fails loudly
when the CUDA library is missing -- there is no CPU fallback in the product path.
"""
import ctypes as C
import os

import numpy as np

from .flymodel import FbModel, FlyModel

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libflybody_b200.so')

# enum FbField
(QPOS, QVEL, ACT, CTRL, QACC, QACC_WARMSTART, SENSORDATA, SENSOR_MEAN, XPOS, XMAT, SITE_XPOS, SITE_XMAT,
 SUBTREE_COM, NCON, NEFC, TIME, QFRC_SMOOTH, QM_DENSE, QFRC_CONSTRAINT, SOLVER_NITER, QFRC_PASSIVE,
 QFRC_BIAS, QFRC_ACTUATOR, CONTACT, EFC_FORCE, FLAGS) = range(26)
MAXCON, MAXEFC = 64, 160

EXPORTS = ['fb_create', 'fb_destroy', 'fb_reset', 'fb_reset_hold', 'fb_set_ctrl', 'fb_set_action_map', 'fb_write_state', 'fb_step', 'fb_forward',
           'fb_get', 'fb_field_size', 'fb_record_stride', 'fb_set', 'fb_obs_ptr', 'fb_n_envs', 'fb_n_envs_padded', 'fb_stream',
           'fb_sync', 'fb_pack_obs', 'fb_read_obs', 'fb_obs_program', 'fb_ref_slots', 'fb_ref_slot_write', 'fb_task_program', 'fb_task_step', 'fb_task_reset_all', 'fb_task_request_reset', 'fb_task_set_reset_noise', 'fb_task_episodes', 'fb_task_uniforms', 'fb_task_uniform_rows', 'fb_hfield_bank', 'fb_task_ptrs', 'fb_task_read', 'fb_eye_program', 'fb_hfield_collision', 'fb_hfield_write', 'fb_render_eyes', 'fb_eyes_ptr', 'fb_eyes_read', 'fb_task_inputs', 'fb_read_task_obs', 'fb_profile', 'fb_profile_read', 'fb_profile_name', 'fb_launch_count', 'fb_last_step_ms', 'fb_set_solver', 'fb_last_error', 'fb_version']


# enum FbObsItem
(OBS_SENSOR_MEAN, OBS_SENSOR_NOW, OBS_ACT, OBS_QPOS, OBS_QVEL, OBS_SITES_EGO, OBS_ROOT_ZAXIS, OBS_REF_DISP,
 OBS_REF_QUAT, OBS_SCALARS, OBS_ROOT_POSE, OBS_SUBTREE_COM, OBS_DOF_AXIS_EGO, OBS_WORLD_CONTACT, OBS_TASK_TARGET) = range(15)


class FbObsProgram(C.Structure):
    _fields_ = [('n_items', C.c_int32), ('kind', C.POINTER(C.c_int32)), ('a', C.POINTER(C.c_int32)),
                ('b', C.POINTER(C.c_int32)), ('n_list', C.c_int32), ('list', C.POINTER(C.c_int32)),
                ('root_body', C.c_int32), ('n_sub', C.c_int32), ('ref_len', C.c_int32),
                ('ref_qpos', C.POINTER(C.c_float))]


class FbTaskProgram(C.Structure):
    """include/flybody_b200.h: FbTaskProgram (device-side task logic)."""
    _fields_ = [('kind', C.c_int32), ('root_qadr', C.c_int32), ('root_vadr', C.c_int32), ('ghost_qadr', C.c_int32), ('ghost_vadr', C.c_int32),
                ('user_col', C.c_int32), ('ghost_offset', C.c_float * 3), ('control_timestep', C.c_float), ('time_limit', C.c_float),
                ('terminal_com_dist', C.c_float), ('terminal_linvel', C.c_float), ('terminal_angvel', C.c_float), ('terminal_qacc', C.c_float),
                ('terminal_height', C.c_float), ('velocimeter_adr', C.c_int32), ('gyro_adr', C.c_int32), ('com_body', C.c_int32),
                ('episode_steps', C.c_int32), ('ref_len', C.c_int32), ('ref_qpos', C.POINTER(C.c_float)), ('ref_qvel', C.POINTER(C.c_float)),
                ('obs_refdisp_off', C.c_int32), ('obs_refquat_off', C.c_int32), ('reset_qpos', C.POINTER(C.c_float)),
                ('n_noise', C.c_int32), ('noise_qadr', C.POINTER(C.c_int32)), ('noise_amp', C.c_float), ('seed', C.c_uint32),
                ('n_wing', C.c_int32), ('wing_qadr', C.POINTER(C.c_int32)), ('wing_vadr', C.POINTER(C.c_int32)), ('wing_ctrl', C.POINTER(C.c_int32)),
                ('n_freq', C.c_int32), ('tab_len', C.c_int32), ('wb_traj', C.POINTER(C.c_float)), ('wb_phase', C.POINTER(C.c_float)),
                ('wb_phase_mod', C.POINTER(C.c_float)), ('wb_freqs', C.POINTER(C.c_float)), ('wb_len', C.POINTER(C.c_int32)),
                ('wb_base_freq', C.c_float), ('wb_rel_range', C.c_float), ('wb_rate', C.c_float), ('com_offset', C.c_float * 3),
                ('target_height_range', C.c_float * 2), ('target_speed_range', C.c_float * 2), ('init_x_range', C.c_float * 2), ('init_y_range', C.c_float * 2),
                ('hover_quat', C.c_float * 4), ('target_zaxis', C.c_float * 3), ('floor_contacts_fatal', C.c_int32),
                ('trench_cap', C.c_int32), ('trench_x', C.POINTER(C.c_float)), ('trench_len', C.POINTER(C.c_int32)), ('trench_y', C.POINTER(C.c_float))]


class FbEyeProgram(C.Structure):
    """include/flybody_b200.h: FbEyeProgram (eye cameras)."""
    _fields_ = [('n_cam', C.c_int32), ('body', C.c_int32 * 2), ('pos', (C.c_float * 3) * 2), ('quat', (C.c_float * 4) * 2),
                ('fovy_deg', C.c_float), ('size', C.c_int32), ('nrow', C.c_int32), ('ncol', C.c_int32), ('half_size', C.c_float),
                ('z_offset', C.c_float), ('zfar', C.c_float), ('sky_top', C.c_float * 3), ('sky_horizon', C.c_float * 3),
                ('ground', C.c_float * 3), ('ambient', C.c_float), ('diffuse', C.c_float)]


class StepperError(RuntimeError):
    pass


def load_library(path=None):
    path = path or os.environ.get('FLYBODY_B200_LIB') or LIB_PATH      # the env override is for A/B timing of kernel variants
    if not os.path.exists(path):
        raise StepperError(
            f'CUDA stepper library not found at {path}. Build it with `python -c "import __graft_entry__ as g; '
            f'g.build()"` (nvcc, sm_100a). There is no CPU fallback.')
    lib = C.CDLL(path)
    lib.fb_create.argtypes = [C.POINTER(FbModel), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.fb_destroy.argtypes = [C.c_void_p]
    lib.fb_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.fb_reset_hold.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.fb_set_ctrl.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.fb_set_action_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.fb_write_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.fb_step.argtypes = [C.c_void_p, C.c_int]
    lib.fb_forward.argtypes = [C.c_void_p]
    lib.fb_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.fb_field_size.argtypes = [C.c_void_p, C.c_int]
    lib.fb_record_stride.argtypes = [C.c_void_p]
    lib.fb_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.fb_obs_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.fb_obs_program.argtypes = [C.c_void_p, C.c_void_p]
    lib.fb_task_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fb_ref_slots.argtypes = [C.c_void_p, C.c_int]
    lib.fb_task_program.argtypes = [C.c_void_p, C.c_void_p]
    lib.fb_task_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.fb_task_reset_all.argtypes = [C.c_void_p]
    lib.fb_task_uniforms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.fb_task_request_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.fb_task_episodes.argtypes = [C.c_void_p, C.c_void_p]
    lib.fb_task_set_reset_noise.argtypes = [C.c_void_p, C.c_float]
    lib.fb_task_uniform_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.fb_hfield_bank.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.fb_task_ptrs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    lib.fb_task_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fb_eye_program.argtypes = [C.c_void_p, C.c_void_p]
    lib.fb_hfield_write.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.fb_hfield_collision.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.fb_render_eyes.argtypes = [C.c_void_p]
    lib.fb_eyes_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.fb_eyes_read.argtypes = [C.c_void_p, C.c_void_p]
    lib.fb_ref_slot_write.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.fb_read_task_obs.argtypes = [C.c_void_p, C.c_void_p]
    lib.fb_pack_obs.argtypes = [C.c_void_p]
    lib.fb_read_obs.argtypes = [C.c_void_p, C.c_void_p]
    lib.fb_profile.argtypes = [C.c_void_p, C.c_int]
    lib.fb_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.fb_profile_name.argtypes = [C.c_int]
    lib.fb_profile_name.restype = C.c_char_p
    lib.fb_n_envs.argtypes = [C.c_void_p]
    lib.fb_n_envs_padded.argtypes = [C.c_void_p]
    lib.fb_stream.argtypes = [C.c_void_p]
    lib.fb_stream.restype = C.c_void_p
    lib.fb_sync.argtypes = [C.c_void_p]
    lib.fb_launch_count.argtypes = [C.c_void_p]
    lib.fb_launch_count.restype = C.c_longlong
    lib.fb_last_step_ms.argtypes = [C.c_void_p]
    lib.fb_last_step_ms.restype = C.c_float
    lib.fb_set_solver.argtypes = [C.c_void_p, C.c_float, C.c_int]
    lib.fb_last_error.argtypes = [C.c_void_p]
    lib.fb_last_error.restype = C.c_char_p
    lib.fb_version.restype = C.c_char_p
    return lib


class BatchedStepper:
    """N fly environments stepped in lock-step on one device."""

    def __init__(self, model: FlyModel, n_envs: int, device: int = 0, lib_path: str | None = None):
        self.model = model
        self.n_envs = int(n_envs)
        self.device = int(device)
        self._lib = load_library(lib_path)
        h = C.c_void_p()
        rc = self._lib.fb_create(C.byref(model.c), self.n_envs, int(device), C.byref(h))
        self._h = h
        if rc != 0:
            msg = self._lib.fb_last_error(h).decode() if h else 'no handle'
            raise StepperError(f'fb_create failed ({rc}): {msg}')

    def close(self):
        if getattr(self, '_h', None):
            self._lib.fb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise StepperError(f'{what} failed ({rc}): {self._lib.fb_last_error(self._h).decode()}')

    # -- state access ---------------------------------------------------------------------
    def get(self, field):
        n = self._lib.fb_field_size(self._h, field)
        if n < 0:
            raise KeyError(field)
        out = np.zeros((self.n_envs, max(n, 1)), np.float32)
        self._check(self._lib.fb_get(self._h, field, out.ctypes.data, 0), 'fb_get')
        return out[:, :n]

    def set(self, field, values):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(values, np.float32), (self.n_envs, np.asarray(values).shape[-1])))
        self._check(self._lib.fb_set(self._h, field, v.ctypes.data), 'fb_set')

    def set_control(self, ctrl):
        """physics.set_control (reference fruitfly.py:540-544), ctrl [N, nu] (or [nu], broadcast)."""
        k = getattr(self, '_n_ctrl_cols', self.model.nu)      # action columns after set_action_map, else nu
        c = np.ascontiguousarray(np.broadcast_to(np.asarray(ctrl, np.float32), (self.n_envs, k)))
        self._check(self._lib.fb_set_ctrl(self._h, c.ctypes.data, 0), 'fb_set_ctrl')

    def write_state(self, field, idx, vals):
        idx = np.ascontiguousarray(idx, np.int32)
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(vals, np.float32), (self.n_envs, len(idx))))
        self._check(self._lib.fb_write_state(self._h, field, idx.ctypes.data, len(idx), v.ctypes.data), 'fb_write_state')

    def reset(self, qpos=None, qvel=None, env_ids=None):
        n = self.n_envs if env_ids is None else len(env_ids)
        ids = None if env_ids is None else np.ascontiguousarray(env_ids, np.int32)
        qp = None if qpos is None else np.ascontiguousarray(np.broadcast_to(np.asarray(qpos, np.float32), (n, self.model.nq)))
        qv = None if qvel is None else np.ascontiguousarray(np.broadcast_to(np.asarray(qvel, np.float32), (n, self.model.nv)))
        self._check(self._lib.fb_reset(self._h, None if ids is None else ids.ctypes.data, n,
                                       None if qp is None else qp.ctypes.data, None if qv is None else qv.ctypes.data), 'fb_reset')

    def reset_hold(self, env_ids, qpos, qvel=None):
        ids = np.ascontiguousarray(env_ids, np.int32)
        qp = np.ascontiguousarray(np.broadcast_to(np.asarray(qpos, np.float32), (len(ids), self.model.nq)))
        qv = None if qvel is None else np.ascontiguousarray(np.broadcast_to(np.asarray(qvel, np.float32), (len(ids), self.model.nv)))
        self._check(self._lib.fb_reset_hold(self._h, ids.ctypes.data, len(ids), qp.ctypes.data,
                                            None if qv is None else qv.ctypes.data), 'fb_reset_hold')

    def set_action_map(self, ctrl_index):
        """fb_set_action_map: afterwards set_control takes rows in action order (see the header)."""
        if ctrl_index is None:
            self._n_ctrl_cols = self.model.nu
            self._check(self._lib.fb_set_action_map(self._h, None, 0), 'fb_set_action_map')
            return
        idx = np.ascontiguousarray(ctrl_index, np.int32)
        self._check(self._lib.fb_set_action_map(self._h, idx.ctypes.data, len(idx)), 'fb_set_action_map')
        self._n_ctrl_cols = len(idx)

    def step(self, n_substeps):
        self._check(self._lib.fb_step(self._h, int(n_substeps)), 'fb_step')

    def forward(self):
        self._check(self._lib.fb_forward(self._h), 'fb_forward')

    def sync(self):
        self._check(self._lib.fb_sync(self._h), 'fb_sync')

    def set_solver(self, tolerance=0.0, max_iter=0):
        self._lib.fb_set_solver(self._h, float(tolerance), int(max_iter))

    @property
    def launch_count(self):
        return int(self._lib.fb_launch_count(self._h))

    @property
    def last_step_ms(self):
        return float(self._lib.fb_last_step_ms(self._h))

    @property
    def stream(self):
        return self._lib.fb_stream(self._h)

    def obs_ptr(self):
        p = C.c_void_p()
        n = C.c_int()
        self._check(self._lib.fb_obs_ptr(self._h, C.byref(p), C.byref(n)), 'fb_obs_ptr')
        return p.value, n.value

    def obs_program(self, items, lists, root_body, n_sub, ref_qpos=None):
        """items: [(kind, a, b)], lists: flat int list referenced by QPOS/QVEL/SITES items. Returns row length."""
        kind = np.ascontiguousarray([i[0] for i in items], np.int32)
        a = np.ascontiguousarray([i[1] for i in items], np.int32)
        b = np.ascontiguousarray([i[2] for i in items], np.int32)
        lst = np.ascontiguousarray(lists if len(lists) else [0], np.int32)
        ref = None if ref_qpos is None else np.ascontiguousarray(ref_qpos, np.float32)
        p = FbObsProgram(len(items), kind.ctypes.data_as(C.POINTER(C.c_int32)), a.ctypes.data_as(C.POINTER(C.c_int32)),
                         b.ctypes.data_as(C.POINTER(C.c_int32)), len(lists), lst.ctypes.data_as(C.POINTER(C.c_int32)),
                         int(root_body), int(n_sub), 0 if ref is None else ref.shape[0],
                         None if ref is None else ref.ctypes.data_as(C.POINTER(C.c_float)))
        dim = self._lib.fb_obs_program(self._h, C.byref(p))
        if dim < 0:
            raise StepperError(f'fb_obs_program failed ({dim}): {self._lib.fb_last_error(self._h).decode()}')
        self._tobs_dim = dim
        return dim

    def ref_slots(self, slot_len):
        """one reference table [slot_len][7] per env instead of the shared one (fb_ref_slots)."""
        self._check(self._lib.fb_ref_slots(self._h, int(slot_len)), 'fb_ref_slots')
        self._slot_len = int(slot_len)

    def ref_slot_write(self, env_ids, rows):
        ids = np.ascontiguousarray(env_ids, np.int32)
        r = np.ascontiguousarray(rows, np.float32)
        assert r.shape == (len(ids), self._slot_len, 7), r.shape
        self._check(self._lib.fb_ref_slot_write(self._h, ids.ctypes.data, len(ids), r.ctypes.data), 'fb_ref_slot_write')

    # ---- device-side task logic (fb_task_*)
    def task_program(self, **kw):
        """upload an FbTaskProgram; array arguments are numpy arrays (kept alive for the duration of the call)."""
        keep = []

        def fp(a):
            a = np.ascontiguousarray(a, np.float32); keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_float))

        def ip(a):
            a = np.ascontiguousarray(a, np.int32); keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int32))
        p = FbTaskProgram()
        for name, ctype in FbTaskProgram._fields_:
            if name not in kw or kw[name] is None:
                continue
            v = kw[name]
            if ctype is C.POINTER(C.c_float):
                v = fp(v)
            elif ctype is C.POINTER(C.c_int32):
                v = ip(v)
            elif ctype in (C.c_float * 2, C.c_float * 3, C.c_float * 4):
                v = ctype(*[float(x) for x in v])
            setattr(p, name, v)
        self._check(self._lib.fb_task_program(self._h, C.byref(p)), 'fb_task_program')

    def task_step(self, action, n_substeps, is_device=False):
        if is_device:
            self._check(self._lib.fb_task_step(self._h, C.c_void_p(int(action)), 1, int(n_substeps)), 'fb_task_step')
        else:
            a = np.ascontiguousarray(action, np.float32)
            self._keep_action = a                  # the copy is asynchronous
            self._check(self._lib.fb_task_step(self._h, a.ctypes.data, 0, int(n_substeps)), 'fb_task_step')

    def task_reset_all(self):
        self._check(self._lib.fb_task_reset_all(self._h), 'fb_task_reset_all')

    def task_request_reset(self, env_ids):
        ids = np.ascontiguousarray(env_ids, np.int32)
        self._check(self._lib.fb_task_request_reset(self._h, ids.ctypes.data, len(ids)), 'fb_task_request_reset')

    def task_uniform_rows(self, env_ids, rows):
        ids = np.ascontiguousarray(env_ids, np.int32); u = np.ascontiguousarray(rows, np.float32).reshape(len(ids), 8)
        self._check(self._lib.fb_task_uniform_rows(self._h, ids.ctypes.data, len(ids), u.ctypes.data), 'fb_task_uniform_rows')

    def hfield_bank(self, heights):
        h = np.ascontiguousarray(heights, np.float32).reshape(len(heights), -1)
        assert h.shape[1] == self._hfield_cells, (h.shape, self._hfield_cells)
        self._check(self._lib.fb_hfield_bank(self._h, h.shape[0], h.ctypes.data), 'fb_hfield_bank')

    def task_set_reset_noise(self, amp):
        self._check(self._lib.fb_task_set_reset_noise(self._h, float(amp)), 'fb_task_set_reset_noise')

    def task_episodes(self):
        out = np.zeros(self.n_envs, np.int32)
        self._check(self._lib.fb_task_episodes(self._h, out.ctypes.data), 'fb_task_episodes')
        return out

    def task_uniforms(self, env_ids, u):
        ids = np.ascontiguousarray(env_ids, np.int32); uu = np.ascontiguousarray(u, np.float32)
        self._check(self._lib.fb_task_uniforms(self._h, ids.ctypes.data, len(ids), uu.ctypes.data), 'fb_task_uniforms')

    def task_ptrs(self):
        o, n, r = C.c_void_p(), C.c_int(), C.c_void_p()
        self._check(self._lib.fb_task_ptrs(self._h, C.byref(o), C.byref(n), C.byref(r)), 'fb_task_ptrs')
        return o.value, n.value, r.value

    def task_read(self, obs_out, out4):
        self._check(self._lib.fb_task_read(self._h, obs_out.ctypes.data, out4.ctypes.data), 'fb_task_read')

    # ---- eye cameras (fb_eye_program / fb_render_eyes)
    def eye_program(self, bodies, pos, quat, fovy_deg=150.0, size=32, nrow=0, ncol=0, half_size=20.0, z_offset=0.0, zfar=50.0,
                    sky_top=(0.33, 0.47, 0.70), sky_horizon=(0.50, 0.56, 0.65), ground=(0.15, 0.11, 0.08), ambient=0.4, diffuse=0.8):
        """The default palette is calibrated, not copied: MuJoCo's GL image cannot be reproduced, but the gray-level statistics the
        reference's vision network assumes can -- `VisNet` normalises with mean 77 / std 56 "from the trench task"
        (network_factory_vis.py:158-162); with these colours the eyes of `vision_guided_flight('trench')` measure 76 / 56 (bumps: 80 / 57),
        tests/test_eyes.py::test_eye_statistics_match_the_vision_network_normalisation."""
        p = FbEyeProgram()
        p.n_cam = len(bodies)
        for c in range(p.n_cam):
            p.body[c] = int(bodies[c])
            for k in range(3):
                p.pos[c][k] = float(pos[c][k])
            for k in range(4):
                p.quat[c][k] = float(quat[c][k])
        p.fovy_deg, p.size, p.nrow, p.ncol, p.half_size, p.z_offset, p.zfar = fovy_deg, size, nrow, ncol, half_size, z_offset, zfar
        for k in range(3):
            p.sky_top[k], p.sky_horizon[k], p.ground[k] = sky_top[k], sky_horizon[k], ground[k]
        p.ambient, p.diffuse = ambient, diffuse
        self._check(self._lib.fb_eye_program(self._h, C.byref(p)), 'fb_eye_program')
        self._eye_shape = (self.n_envs, p.n_cam, size, size, 3)
        self._hfield_cells = nrow * ncol

    def hfield_collision(self, geom, size, nrow, ncol, pair_geom):
        """turn on the heightfield narrowphase for the model's terrain geom (fb_hfield_collision)."""
        sz = np.ascontiguousarray(size, np.float32); pg = np.ascontiguousarray(pair_geom, np.int32)
        self._check(self._lib.fb_hfield_collision(self._h, int(geom), sz.ctypes.data, int(nrow), int(ncol), pg.ctypes.data, len(pg)), 'fb_hfield_collision')
        self._hfield_cells = int(nrow) * int(ncol)

    def hfield_write(self, env_ids, heights):
        ids = np.ascontiguousarray(env_ids, np.int32)
        h = np.ascontiguousarray(heights, np.float32).reshape(len(ids), -1)
        assert h.shape[1] == self._hfield_cells, (h.shape, self._hfield_cells)
        self._check(self._lib.fb_hfield_write(self._h, ids.ctypes.data, len(ids), h.ctypes.data), 'fb_hfield_write')

    def eyes_ptr(self):
        """(device pointer, bytes per env) of the eye-camera images (valid after eye_program; rewritten by every render_eyes)"""
        p, n = C.c_void_p(), C.c_int()
        self._check(self._lib.fb_eyes_ptr(self._h, C.byref(p), C.byref(n)), 'fb_eyes_ptr')
        return p.value, n.value

    def render_eyes_async(self):
        """launch the eye renderer on the stepper's stream; the images stay on the device (eyes_ptr)"""
        self._check(self._lib.fb_render_eyes(self._h), 'fb_render_eyes')

    def render_eyes(self, out=None):
        """-> uint8 [n_envs, n_cam, size, size, 3] rendered from the current body poses."""
        self._check(self._lib.fb_render_eyes(self._h), 'fb_render_eyes')
        out = np.empty(self._eye_shape, np.uint8) if out is None else out
        self._check(self._lib.fb_eyes_read(self._h, out.ctypes.data), 'fb_eyes_read')
        return out

    def task_inputs(self, step_idx, first):
        si = np.ascontiguousarray(step_idx, np.int32)
        fi = np.ascontiguousarray(first, np.uint8)
        self._check(self._lib.fb_task_inputs(self._h, si.ctypes.data, fi.ctypes.data), 'fb_task_inputs')

    def read_task_obs(self, out):
        self._check(self._lib.fb_pack_obs(self._h), 'fb_pack_obs')
        self._check(self._lib.fb_read_task_obs(self._h, out.ctypes.data), 'fb_read_task_obs')
        return out

    def profile(self, enable=True):
        self._lib.fb_profile(self._h, 1 if enable else 0)

    def profile_read(self):
        ms = np.zeros(16, np.float64)
        cnt = np.zeros(16, np.int64)
        n = self._lib.fb_profile_read(self._h, ms.ctypes.data, cnt.ctypes.data, 16)
        return {self._lib.fb_profile_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n)}

    def set_control_device(self, dev_ptr):
        """ctrl already resident on the device as contiguous rows [n_envs][nu] fp32."""
        self._check(self._lib.fb_set_ctrl(self._h, C.c_void_p(dev_ptr), 1), 'fb_set_ctrl')

    def pack_obs(self):
        self._check(self._lib.fb_pack_obs(self._h), 'fb_pack_obs')

    @property
    def record_bytes(self):
        """bytes of device state per env (one record: persistent state + per-substep intermediates)."""
        return 4 * int(self._lib.fb_record_stride(self._h))

    @property
    def n_envs_padded(self):
        return int(self._lib.fb_n_envs_padded(self._h))

    def read_obs(self, out=None):
        """Pack + copy the per-env observation record [N, obs_dim] (see fb_pack_obs in the header)."""
        _, dim = self.obs_ptr()
        if out is None:
            out = np.empty((self.n_envs, dim), np.float32)
        self._check(self._lib.fb_pack_obs(self._h), 'fb_pack_obs')
        self._check(self._lib.fb_read_obs(self._h, out.ctypes.data), 'fb_read_obs')
        return out

    def obs_layout(self):
        m = self.model
        off, lay = 0, {}
        for name, n in (('qpos', m.nq), ('qvel', m.nv), ('act', m.na), ('sensor_mean', m.nsensordata),
                        ('sensordata', m.nsensordata), ('root_xpos', 3), ('root_xmat', 9),
                        ('site_xpos', 3 * m.nsite), ('flags', 1), ('qacc_sq', 1), ('time', 1)):
            lay[name] = slice(off, off + n)
            off += n
        return lay

    def version(self):
        return self._lib.fb_version().decode()
