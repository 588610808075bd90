"""Drop-in, batched counterparts of the reference's environment factories
(`flybody/fly_envs.py:100-155` walk_imitation, `:30-97` flight_imitation).

`walk_imitation(n_envs=N)` returns a `BatchedFlyEnv` with the `dm_env` surface of the reference
(`reset()`, `step(action)`, `action_spec()`, `observation_spec()`, `control_timestep()`,
`physics.timestep()`, `task._traj_generator.set_next_trajectory`) where every array carries a
leading batch dimension N.  With `n_envs=None` a single environment without the batch dimension
is returned (what the reference's tests construct).  The physics is the CUDA stepper; the task
logic mirrors `flybody/tasks/walk_imitation.py` / `tasks/base.py` in vectorised numpy.
"""
import collections
import hashlib

import numpy as np

from . import rewards as rw
from . import stepper as st
from .dm_env_shim import Array, BoundedArray, StepType, TimeStep
from .flymodel import load_model

from .synthetic import (_WALK_CONTROL_TIMESTEP, _WALK_PHYSICS_TIMESTEP, _TERMINAL_LINVEL, _TERMINAL_ANGVEL,  # noqa: F401
                        _FLY_CONTROL_TIMESTEP, _FLY_PHYSICS_TIMESTEP, _TERMINAL_HEIGHT, _TERMINAL_QACC, _ACTION_CLASS_ORDER,
                        mult_quat, reciprocal_quat, quat_dist_short_arc, linear_tolerance, constant_speed_trajectory,
                        rotate_vec_with_quat, com2root, root2com, _COM_OFFSET)
from .trajectory_loaders import (HDF5FlightTrajectoryLoader, HDF5WalkingTrajectoryLoader,  # noqa: F401
                                 InferenceFlightTrajectoryLoader, InferenceWalkingTrajectoryLoader)


class BatchedWingBeatPatternGenerator:
    """Vectorised `WingBeatPatternGenerator` (reference `tasks/pattern_generators.py:8-203`): per-env phase-preserving
    frequency modulation of one wing-beat cycle.  Tables are built once (one resampled, repeated cycle per discrete beat
    frequency, padded to a common length); the per-env state is (filtered frequency, table index, position in table)."""

    def __init__(self, n_envs, base_pattern_path=None, base_beat_freq=218.0, rel_freq_range=0.05, num_freqs=201,
                 min_repeats=10, max_repeats=20, dt_ctrl=_FLY_CONTROL_TIMESTEP, ctrl_filter=0.5 / 218.0):
        if base_pattern_path is None:       # the reference's synthetic stand-in cycle (pattern_generators.py:53-59)
            x = np.linspace(0, 2 * np.pi, 500)
            cycle = np.stack([1.1 * np.sin(x - np.pi / 2) + 0.3, 0.25 * np.sin(1.5 * x) - 0.1, 1.35 * np.sin(x) + 0.8], 1)
        else:
            cycle = np.load(base_pattern_path)
        cycle = np.concatenate([cycle, cycle], 1)              # both wings
        n_pts = cycle.shape[0]
        self.base_beat_freq, self.rel_freq_range, self.ctrl_filter, self._dt = base_beat_freq, rel_freq_range, ctrl_filter, dt_ctrl
        self._rate = np.exp(-dt_ctrl / ctrl_filter) if ctrl_filter != 0.0 else 0.0
        self.beat_freqs = np.linspace((1 - rel_freq_range) * base_beat_freq, (1 + rel_freq_range) * base_beat_freq, num_freqs)
        trajs, phases = [], []
        reps = np.arange(min_repeats, max_repeats + 1)
        for f in self.beat_freqs:
            period = 1.0 / f
            err = ((reps * period) % dt_ctrl) / dt_ctrl        # mismatch of the seam, in control steps
            i_over, i_under = int(np.argmin(err)), int(np.argmin(np.abs(1 - err)))
            if err[i_over] < abs(1 - err[i_under]):
                pick, shift = i_over, dt_ctrl
            else:
                pick, shift = i_under, 0.0
            n_rep = pick + 1                                    # (sic) the reference repeats argmin+1 cycles
            data = np.tile(cycle, (n_rep, 1))
            ph = np.linspace(0, n_rep, n_rep * n_pts, endpoint=False)
            total = data.shape[0] * (period / n_pts)
            t_data = np.linspace(0, total, data.shape[0])
            t_ctrl = np.arange(0, total - shift, dt_ctrl)
            trajs.append(np.stack([np.interp(t_ctrl, t_data, data[:, k]) for k in range(data.shape[1])], 1))
            phases.append(np.interp(t_ctrl, t_data, ph))
        self.lengths = np.array([t.shape[0] for t in trajs])
        L = int(self.lengths.max())
        self.traj = np.zeros((num_freqs, L, cycle.shape[1]))
        self.phase = np.full((num_freqs, L), np.inf)            # padding never wins an argmin
        for i, (t, ph) in enumerate(zip(trajs, phases)):
            self.traj[i, :t.shape[0]] = t
            self.phase[i, :t.shape[0]] = ph
        self.phase_mod = np.where(np.isfinite(self.phase), np.mod(self.phase, 1.0, where=np.isfinite(self.phase), out=np.zeros_like(self.phase)), np.inf)
        self.freq = np.full(n_envs, base_beat_freq)
        self.idx = np.zeros(n_envs, np.int64)
        self.pos = np.zeros(n_envs, np.int64)

    def _nearest(self, freq):
        return np.abs(self.beat_freqs[None, :] - np.asarray(freq)[:, None]).argmin(1)

    def reset(self, ids, initial_phase):
        """-> (wing qpos [n, 6], wing qvel [n, 6]) at the requested phase of the base frequency."""
        ids = np.asarray(ids)
        self.freq[ids] = self.base_beat_freq
        idx = self._nearest(self.freq[ids])
        pos = np.abs(np.asarray(initial_phase)[:, None] - self.phase[idx]).argmin(1)
        self.idx[ids], self.pos[ids] = idx, pos
        nxt = np.minimum(pos + 1, self.lengths[idx] - 1)
        q = self.traj[idx, pos]
        return q, (self.traj[idx, nxt] - q) / self._dt

    def step(self, ctrl_freq, active=None):
        """advance every (active) env by one control step at its requested beat frequency -> wing angles [N, 6]."""
        act = np.ones(self.freq.shape[0], bool) if active is None else np.asarray(active, bool)
        pos = np.where(act, (self.pos + 1) % self.lengths[self.idx], self.pos)
        freq = ctrl_freq if self.ctrl_filter == 0.0 else self.freq * self._rate + ctrl_freq * (1 - self._rate)
        self.freq = np.where(act, freq, self.freq)
        new = self._nearest(self.freq)
        sw = np.nonzero(act & (new != self.idx))[0]
        if sw.size:                                              # keep the phase within the beat when changing table
            cur = self.phase_mod[self.idx[sw], pos[sw]]
            pos[sw] = np.abs(cur[:, None] - self.phase_mod[new[sw]]).argmin(1)
            self.idx[sw] = new[sw]
        self.pos = pos
        return self.traj[self.idx, self.pos]



class _FeatureBank:
    """Full-body reference features of the walking snippets in use, concatenated row-wise (append-only, cached by
    trajectory index) so that the per-env reward gathers are single fancy-index reads."""

    def __init__(self):
        self._off, self._n = {}, 0
        self.qpos = self.qvel = self.root2site = self.joint_quat = None

    @staticmethod
    def _append(buf, n_used, new):
        new = np.asarray(new, np.float64)
        if buf is None:
            buf = np.zeros((max(4 * new.shape[0], 1024),) + new.shape[1:])
        while n_used + new.shape[0] > buf.shape[0]:
            buf = np.concatenate([buf, np.zeros_like(buf)], 0)
        buf[n_used:n_used + new.shape[0]] = new
        return buf

    def offset_of(self, idx, snip):
        key = None if idx is None else int(idx)
        if key is not None and key in self._off:
            return self._off[key]
        T = snip['qpos'].shape[0]
        self.qpos = self._append(self.qpos, self._n, snip['qpos'])
        self.qvel = self._append(self.qvel, self._n, snip['qvel'])
        self.root2site = self._append(self.root2site, self._n, snip['root2site'])
        self.joint_quat = self._append(self.joint_quat, self._n, snip['joint_quat'])
        off = self._n
        self._n += T
        if key is not None:
            self._off[key] = off
        return off


class _PhysicsFacade:
    """The slice of `dm_control.mjcf.Physics` the reference's callers touch."""

    def __init__(self, env):
        self._env = env

    def timestep(self):
        return self._env._physics_timestep

    def time(self):
        return self._env._time.copy() if self._env._batched else float(self._env._time[0])

    @property
    def stepper(self):
        return self._env._sim


class _ImitationTask:
    """What the reference's callers reach through `env.task` (`tasks/base.py:22-268`)."""
    name = 'FruitFlyTask'

    def __init__(self, env, traj_generator, terminal_com_dist, future_steps, time_limit, control_timestep):
        self._env = env
        self._traj_generator = traj_generator
        self._terminal_com_dist = terminal_com_dist
        self._future_steps = future_steps
        self._time_limit = time_limit
        self._max_episode_steps = round(time_limit / control_timestep) + 1
        self._ghost_offset = np.zeros(3)
        self._next_traj_idx = None

    def set_next_trajectory_index(self, idx):
        self._next_traj_idx = idx


class WalkImitationTask(_ImitationTask):
    """Batched `WalkImitation` (reference `tasks/walk_imitation.py:19-203`, `tasks/base.py:367-428`)."""


class FlightImitationTask(_ImitationTask):
    """Batched `FlightImitationWBPG` (reference `tasks/flight_imitation.py:16-212`, `tasks/base.py:271-364`)."""

    def __init__(self, env, wbpg, *a, **k):
        super().__init__(env, *a, **k)
        self._wbpg = wbpg


_VARIANTS = {
    # control dt, time limit, observation names in spec order (hidden '_' entries feed termination / reward)
    'walk': dict(dt=_WALK_CONTROL_TIMESTEP, obs=('accelerometer', 'actuator_activation', 'appendages_pos', 'force', 'gyro',
                                                 'joints_pos', 'joints_vel', 'touch', 'velocimeter', 'world_zaxis',
                                                 'ref_displacement', 'ref_root_quat')),
    'flight': dict(dt=_FLY_CONTROL_TIMESTEP, obs=('accelerometer', 'actuator_activation', 'gyro', 'joints_pos', 'joints_vel',
                                                  'velocimeter', 'world_zaxis', 'ref_displacement', 'ref_root_quat')),
}


class BatchedFlyEnv:
    """dm_env-shaped environment over N lock-stepped flies (composer.Environment stand-in)."""

    def __init__(self, variant, n_envs, device=0, terminal_com_dist=0.3, time_limit=10.0, future_steps=64,
                 lib_path=None, reset_noise=0.0, seed=0, traj_generator=None, wpg_pattern_path=None, inference_mode=True,
                 max_reference_steps=None, device_task=False, model=None):
        assert variant in _VARIANTS
        self._variant = variant
        self._batched = n_envs is not None
        self.n_envs = int(n_envs) if self._batched else 1
        self.model = model if model is not None else load_model(variant)      # `model`: a variant from flymodel.model_for
        m = self.model
        self._device = int(device)
        self._sim = st.BatchedStepper(m, self.n_envs, device=device, lib_path=lib_path)
        self._control_timestep = _VARIANTS[variant]['dt']
        self._physics_timestep = float(m.opt_timestep)
        self._n_sub = int(round(self._control_timestep / self._physics_timestep))
        self._time_limit = time_limit
        self.physics = _PhysicsFacade(self)
        if variant == 'walk':
            self.task = WalkImitationTask(self, traj_generator or InferenceWalkingTrajectoryLoader(), terminal_com_dist,
                                          future_steps, time_limit, self._control_timestep)
            self._n_user = 0
        else:
            self._wbpg = BatchedWingBeatPatternGenerator(self.n_envs, base_pattern_path=wpg_pattern_path)
            self.task = FlightImitationTask(self, self._wbpg, traj_generator or InferenceFlightTrajectoryLoader(),
                                            terminal_com_dist, future_steps, time_limit, self._control_timestep)
            self._n_user = 1                                          # flight_imitation.py:38 num_user_actions=1
        self._rs = np.random.RandomState(seed)
        self._reset_noise = reset_noise
        N = self.n_envs
        # --- action <-> ctrl maps (reference fruitfly.py:342-379, 532-579): actuator classes in fixed order, then user
        ci = m.meta['ctrl_indices']
        idx, self._action_indices = [], {}
        for key in _ACTION_CLASS_ORDER:
            if ci.get(key):
                self._action_indices[key] = np.arange(len(idx), len(idx) + len(ci[key]))
                idx.extend(ci[key])
        self._ctrl_of_action = np.asarray(idx, np.int64)
        # walking: actions go to the device as they are, the action -> ctrl permutation and NaN -> 0 run in the scatter
        # kernel (fb_set_action_map); flight keeps the host path because the wing actions are modified by the WBPG first
        self._device_task = bool(device_task)       # task hooks evaluated on the device (fb_task_*): no host logic in step()
        self._device_action_map = variant == 'walk' or self._device_task
        if self._device_action_map:
            # user actions (flight: the beat frequency) have no ctrl slot: column -> -1, read by the device-side task code
            self._sim.set_action_map(np.concatenate([self._ctrl_of_action, -np.ones(self._n_user, np.int64)]) if self._device_task
                                     else self._ctrl_of_action)
        names = [m.meta['actuator_names'][i].split('/')[-1] for i in idx] + [f'user_{i}' for i in range(self._n_user)]
        rng = m.actuator_ctrlrange[idx]
        lo = np.concatenate([rng[:, 0], -np.ones(self._n_user)])
        hi = np.concatenate([rng[:, 1], np.ones(self._n_user)])
        self._action_spec = BoundedArray((len(names),), np.float64, lo, hi, name='\t'.join(names))
        # --- index tables
        jn = m.meta['jnt_names']
        self._root_q = m.jnt_qposadr_of('walker/')
        self._root_v = m.jnt_dofadr_of('walker/')
        self._ghost_q = m.jnt_qposadr_of('ghost/')
        self._ghost_v = m.jnt_dofadr_of('ghost/')
        obsj = [jn.index(n) for n in m.meta['observable_joints']]
        self._obs_qadr = m.jnt_qposadr[obsj]
        self._obs_vadr = m.jnt_dofadr[obsj]
        wing_names = [f'walker/wing_{a}_{s}' for s in ('left', 'right') for a in ('yaw', 'roll', 'pitch')]
        self._wing_qadr = np.array([m.jnt_qposadr_of(n) for n in wing_names])
        self._wing_vadr = np.array([m.jnt_dofadr_of(n) for n in wing_names])
        self._wing_spring = m.qpos_spring[self._wing_qadr]
        sn = m.meta['site_names']
        sens = m.meta['sensor_names']

        def sd(names_):
            out = []
            for n_ in names_:
                i = sens.index('walker/' + n_)
                out.extend(range(m.sensor_adr[i], m.sensor_adr[i] + m.sensor_dim[i]))
            return np.array(out)
        self._sd = dict(accelerometer=sd(['accelerometer']), gyro=sd(['gyro']), velocimeter=sd(['velocimeter']))
        self._has_legs = 'walker/force_tarsus_T1_left' in sens      # Walking always; Flying with disable_legs=False (base.py:360-364)
        if self._has_legs:
            legs = [f'T{k}_{s}' for k in (1, 2, 3) for s in ('left', 'right')]
            self._sd.update(force=sd([f'force_tarsus_{l}' for l in legs]), touch=sd([f'touch_claw_{l}' for l in legs]))
            app = [f'walker/claw_T{k}_{s}' for k in (1, 2, 3) for s in ('left', 'right')] + ['walker/head']
            self._app_sites = np.array([sn.index(n) for n in app])
        else:
            self._app_sites = np.zeros(0, np.int64)
        if variant != 'walk':
            # position of the wing joints inside the joints_pos observable (the host reads wing angles from there)
            self._wing_in_obs = np.array([m.meta['observable_joints'].index(n) for n in wing_names])
            self._wing_qpos_host = np.zeros((N, 6))
        self._leg_act_qadr = np.array([m.jnt_qposadr[m.actuator_trnid[i]] for i in range(m.nu)
                                       if m.actuator_trntype[i] == 0 and any(t in m.meta['actuator_names'][i] for t in ('T1', 'T2', 'T3'))],
                                      np.int64)
        self._future = future_steps + 1
        self._rec = None
        self._program_ref_id = None
        self._program_switched = False
        self.n_capacity_overflows = 0            # env-steps whose contact / constraint-row lists hit their capacity (FB_FLAGS bits 1, 2)
        # One reference per env (dataset loaders hand out a different snippet per episode) or one shared by all envs (the
        # Inference* loaders hold a single trajectory): the shared case keeps one device table, the per-env case one slot per env.
        tg = self.task._traj_generator
        self._per_env_ref = not isinstance(tg, (InferenceWalkingTrajectoryLoader, InferenceFlightTrajectoryLoader))
        self._max_reference_steps = max_reference_steps
        self._ref_rows = None                       # per-env mode: [N, slot_len, 13] fp32 (root qpos 7, root qvel 6), padded with the last row
        self._ref_len = np.zeros(N, np.int64)
        self._episode_steps = np.zeros(N, np.int64)
        # full-body imitation reward (reference walk_imitation.py:152-177): mocap joints / sites named by the dataset
        self._inference_mode = bool(inference_mode) or variant != 'walk'
        if self._device_task and (self._per_env_ref or not self._inference_mode):
            raise NotImplementedError('device_task covers the shared-reference (inference-mode) tasks; dataset mode keeps the '
                                      'host-side task code')
        self._seed = seed
        if not self._inference_mode:
            jnames, snames = tg.get_joint_names(), tg.get_site_names()
            mj = [jn.index('walker/' + n) for n in jnames]
            self._mocap_qadr, self._mocap_vadr = m.jnt_qposadr[mj].astype(np.int64), m.jnt_dofadr[mj].astype(np.int64)
            self._mocap_sites = np.array([sn.index('walker/' + n) for n in snames], np.int64)
            self._bank = _FeatureBank()
            self._env_bank_off = np.zeros(N, np.int64)
        # --- per-env episode state
        self._step_counter = np.zeros(N, np.int64)
        self._time = np.zeros(N)
        self._needs_reset = np.ones(N, bool)
        self._ref_qpos = None
        self.h2d_bytes_per_step = 0
        self.d2h_bytes_per_step = 0
        self.n_resets = 0

    # ---------------------------------------------------------------------------------- specs
    def action_spec(self):
        return self._action_spec

    def _obs_table(self):
        """(name, width, shape, program item) per observable, in spec order; hidden entries last."""
        m, sd, f = self.model, self._sd, self._future
        nq, napp = len(self._obs_qadr), len(self._app_sites)
        o_app, o_q, o_v = 0, napp, napp + nq
        full = {
            'accelerometer': (3, (3,), (st.OBS_SENSOR_MEAN, sd['accelerometer'][0], 3)),
            'actuator_activation': (m.na, (m.na,), (st.OBS_ACT, 0, m.na)),
            'appendages_pos': (3 * napp, (3 * napp,), (st.OBS_SITES_EGO, o_app, napp)),
            'gyro': (3, (3,), (st.OBS_SENSOR_MEAN, sd['gyro'][0], 3)),
            'joints_pos': (nq, (nq,), (st.OBS_QPOS, o_q, nq)),
            'joints_vel': (nq, (nq,), (st.OBS_QVEL, o_v, nq)),
            'velocimeter': (3, (3,), (st.OBS_SENSOR_MEAN, sd['velocimeter'][0], 3)),
            'world_zaxis': (3, (3,), (st.OBS_ROOT_ZAXIS, 0, 3)),
            'ref_displacement': (3 * f, (f, 3), (st.OBS_REF_DISP, 0, f)),
            'ref_root_quat': (4 * f, (f, 4), (st.OBS_REF_QUAT, 0, f)),
        }
        if 'force' in sd:
            full['force'] = (len(sd['force']), (len(sd['force']),), (st.OBS_SENSOR_MEAN, sd['force'][0], len(sd['force'])))
            full['touch'] = (len(sd['touch']), (len(sd['touch']),), (st.OBS_SENSOR_MEAN, sd['touch'][0], len(sd['touch'])))
        names = list(_VARIANTS[self._variant]['obs'])
        if self._has_legs and 'force' not in names:      # Flying with legs: appendages_pos, force, touch join the walker's observables,
            walker = sorted(names[:-2] + ['appendages_pos', 'force', 'touch'])      # which dm_control lists alphabetically, task ones last
            names = walker + names[-2:]
        rows = [('walker/' + k,) + full[k] for k in names]
        rows += [('_velocimeter_now', 3, (3,), (st.OBS_SENSOR_NOW, sd['velocimeter'][0], 3)),
                 ('_gyro_now', 3, (3,), (st.OBS_SENSOR_NOW, sd['gyro'][0], 3)),
                 ('_scalars', 3, (3,), (st.OBS_SCALARS, 0, 3))]
        if self._variant == 'flight':
            rows += [('_root_pose', 7, (7,), (st.OBS_ROOT_POSE, 0, 7)),
                     ('_subtree_com', 3, (3,), (st.OBS_SUBTREE_COM, m.body_id('walker/thorax'), 3)),
                     ('_ghost_pose', 7, (7,), (st.OBS_QPOS, napp + 2 * nq, 7))]     # the ghost where it is after the step
        if not self._inference_mode:                 # what get_walker_features reads (tasks/rewards.py:37-63)
            nj, ns = len(self._mocap_qadr), len(self._mocap_sites)
            o = napp + 2 * nq
            rows += [('_root_pose', 7, (7,), (st.OBS_ROOT_POSE, 0, 7)),
                     ('_root_qvel', 6, (6,), (st.OBS_QVEL, o, 6)),
                     ('_mocap_qpos', nj, (nj,), (st.OBS_QPOS, o + 6, nj)),
                     ('_mocap_qvel', nj, (nj,), (st.OBS_QVEL, o + 6 + nj, nj)),
                     ('_mocap_axes', 3 * nj, (nj, 3), (st.OBS_DOF_AXIS_EGO, o + 6 + nj, nj)),
                     ('_mocap_sites', 3 * ns, (ns, 3), (st.OBS_SITES_EGO, o + 6 + 2 * nj, ns)),
                     ('_wing_qpos', 6, (6,), (st.OBS_QPOS, o + 6 + 2 * nj + ns, 6))]
        return rows

    def observation_spec(self):
        lead = (self.n_envs,) if self._batched else ()
        return collections.OrderedDict((n, Array(lead + shp, np.float32, name=n)) for n, _, shp, _ in self._obs_table()
                                       if not n.startswith('_'))

    def reward_spec(self):
        return Array((self.n_envs,) if self._batched else (), np.float64, name='reward')

    def discount_spec(self):
        return BoundedArray((self.n_envs,) if self._batched else (), np.float64, 0.0, 1.0, name='discount')

    def control_timestep(self):
        return self._control_timestep

    # -------------------------------------------------------------------------------- episode
    def _upload_program(self):
        """Observation program = the task's observables in spec order, evaluated on the device (fb_obs_program)."""
        m = self.model
        rows = self._obs_table()
        lists = list(self._app_sites) + list(self._obs_qadr) + list(self._obs_vadr)
        if self._variant == 'flight':
            lists += list(range(self._ghost_q, self._ghost_q + 7))
        if not self._inference_mode:
            lists += list(range(self._root_v, self._root_v + 6)) + list(self._mocap_qadr) + list(self._mocap_vadr) \
                + list(self._mocap_sites) + list(self._wing_qadr)
        shared = self._ref_qpos[:, :7] if not self._per_env_ref else np.zeros((self._future, 7))
        dim = self._sim.obs_program([r[3] for r in rows], lists, m.body_id('walker/thorax'), self._n_sub, shared)
        if self._per_env_ref:
            self._sim.ref_slots(self._slot_len)
        widths = [r[1] for r in rows]
        assert sum(widths) == dim, (sum(widths), dim)
        off = np.concatenate([[0], np.cumsum(widths)])
        self._obs_slices = {r[0]: slice(int(off[i]), int(off[i + 1])) for i, r in enumerate(rows)}
        self._obs_shapes = {r[0]: r[2] for r in rows}
        N = self.n_envs
        try:
            import torch
            self._rec = torch.empty((N, dim), dtype=torch.float32).pin_memory().numpy() if torch.cuda.is_available() \
                else np.empty((N, dim), np.float32)
        except Exception:
            self._rec = np.empty((N, dim), np.float32)
        if self._device_task:
            self._upload_task_program()
            self._dev_views = None

    def _upload_task_program(self):
        """fb_task_program: the task hooks of this env as a device-side program (same constants as the host path below)."""
        m, t, sl = self.model, self.task, self._obs_slices
        q0 = m.qpos0.copy()
        if self._variant == 'walk':
            q0[self._wing_qadr] = self._wing_spring
        kw = dict(kind=0 if self._variant == 'walk' else 1, root_qadr=self._root_q, root_vadr=self._root_v, ghost_qadr=self._ghost_q,
                  ghost_vadr=self._ghost_v, user_col=len(self._ctrl_of_action) if self._n_user else -1, ghost_offset=t._ghost_offset,
                  control_timestep=self._control_timestep, time_limit=self._time_limit, terminal_com_dist=min(t._terminal_com_dist, 3e38),
                  terminal_linvel=_TERMINAL_LINVEL, terminal_angvel=_TERMINAL_ANGVEL, terminal_qacc=_TERMINAL_QACC, terminal_height=_TERMINAL_HEIGHT,
                  velocimeter_adr=int(self._sd['velocimeter'][0]), gyro_adr=int(self._sd['gyro'][0]), com_body=m.body_id('walker/thorax'),
                  episode_steps=int(self._steps_of(self._ref_qpos.shape[0])), ref_len=self._ref_qpos.shape[0], ref_qpos=self._ref_qpos[:, :7],
                  ref_qvel=self._ref_qvel[:, :6], obs_refdisp_off=sl['walker/ref_displacement'].start, obs_refquat_off=sl['walker/ref_root_quat'].start,
                  reset_qpos=q0, n_noise=len(self._leg_act_qadr) if (self._variant == 'walk' and self._reset_noise > 0) else 0,
                  noise_qadr=self._leg_act_qadr if (self._variant == 'walk' and self._reset_noise > 0) else None,
                  noise_amp=self._reset_noise, seed=int(self._seed) & 0xffffffff, com_offset=_COM_OFFSET)
        if self._variant == 'flight':
            wb = self._wbpg
            kw.update(n_wing=6, wing_qadr=self._wing_qadr, wing_vadr=self._wing_vadr, wing_ctrl=self._ctrl_of_action[self._action_indices['wings']],
                      n_freq=wb.traj.shape[0], tab_len=wb.traj.shape[1], wb_traj=wb.traj, wb_phase=np.where(np.isfinite(wb.phase), wb.phase, 3e38),
                      wb_phase_mod=np.where(np.isfinite(wb.phase_mod), wb.phase_mod, 3e38), wb_freqs=wb.beat_freqs, wb_len=wb.lengths,
                      wb_base_freq=wb.base_beat_freq, wb_rel_range=wb.rel_freq_range, wb_rate=wb._rate)
        self._sim.task_program(**kw)
        self._out4 = np.zeros((self.n_envs, 4), np.float32)

    def _snippet_root(self, snip):
        """root-joint reference (qpos [T,7], qvel [T,6]) of one loader snippet."""
        if self._variant == 'walk':
            return np.asarray(snip['qpos'])[:, :7], np.asarray(snip['qvel'])[:, :6]
        com_qpos, qvel = snip                       # flight data is a CoM trajectory (flight_imitation.py:94-99)
        com_qpos = np.asarray(com_qpos)
        return np.concatenate([com2root(com_qpos[:, :3], com_qpos[:, 3:7]), com_qpos[:, 3:7]], 1), np.asarray(qvel)[:, :6]

    def _steps_of(self, n_rows):
        t = self.task
        if self._variant == 'walk':
            return min(t._max_episode_steps, n_rows - t._future_steps - 1)                               # walk_imitation.py:104-105
        return min(n_rows, round(self._time_limit / self._control_timestep)) - (t._future_steps + 1)     # flight_imitation.py:101-106

    def _load_snippet(self, ids):
        """initialize_episode_mjcf: pick the reference of the episode that starts now (walk_imitation.py:93-110,
        flight_imitation.py:88-111).  Shared mode: ONE trajectory for every env, re-uploaded when the loader returns different
        values (content digest, so in-place edits count).  The switch is global: envs in mid-episode follow the new trajectory from
        their current step on, and with the task hooks on the device the re-upload restarts every env (`_program_switched`) -- a
        batch whose envs need independent references uses the per-env mode (`ref_path=` datasets), where each env of `ids` draws
        its own snippet and only its device slot is rewritten."""
        t = self.task
        tg = t._traj_generator
        if not self._per_env_ref:
            snip = tg.get_trajectory(traj_idx=t._next_traj_idx)
            t._next_traj_idx = None
            qpos, qvel = self._snippet_root(snip)
            qpos, qvel = np.asarray(qpos, np.float64), np.asarray(qvel, np.float64)
            key = (qpos.shape, hashlib.blake2b(qpos.tobytes() + qvel.tobytes(), digest_size=16).digest())
            if self._program_ref_id != key:
                self._program_switched = self._program_ref_id is not None
                self._program_ref_id = key
                self._ref_qpos, self._ref_qvel = qpos, qvel
                self._upload_program()
            self._ref_len[:] = self._ref_qpos.shape[0]
            self._episode_steps[:] = self._steps_of(self._ref_qpos.shape[0])
            return
        if self._ref_rows is None:                   # slot length: the longest reference an episode can consume
            cap = self._max_reference_steps or (t._max_episode_steps + t._future_steps + 1)
            if hasattr(tg, 'trajectory_len') and hasattr(tg, 'num_trajectories'):
                cap = min(cap, max(int(tg.trajectory_len(i)) for i in range(tg.num_trajectories)))
            self._slot_len = int(cap)
            self._ref_rows = np.zeros((self.n_envs, self._slot_len, 13), np.float32)
            self._upload_program()
        L = self._slot_len
        for e in ids:
            idx = t._next_traj_idx
            if idx is None and hasattr(tg, 'traj_indices') and hasattr(tg, '_random_state'):
                idx = tg._random_state.choice(tg.traj_indices)             # the draw the loader itself would make
            snip = tg.get_trajectory(traj_idx=idx)
            qpos, qvel = self._snippet_root(snip)
            n = min(qpos.shape[0], L)
            self._ref_rows[e, :n, :7], self._ref_rows[e, :n, 7:] = qpos[:n], qvel[:n]
            self._ref_rows[e, n:] = self._ref_rows[e, n - 1]
            self._ref_len[e] = n
            self._episode_steps[e] = self._steps_of(n)
            if not self._inference_mode:
                self._env_bank_off[e] = self._bank.offset_of(idx, snip)
        t._next_traj_idx = None
        self._sim.ref_slot_write(ids, self._ref_rows[ids, :, :7])

    def _ref_at(self, step, ids=None):
        """[n, 13] root reference (qpos 7, qvel 6) of the listed envs at their own steps."""
        if not self._per_env_ref:
            return np.concatenate([self._ref_qpos[step, :7], self._ref_qvel[step, :6]], 1)
        ids = np.arange(self.n_envs) if ids is None else ids
        return self._ref_rows[ids, step].astype(np.float64)

    def _reset_envs(self, ids, hold=False):
        """initialize_episode: walk_imitation.py:112-136 (root <- ref_qpos[0], wings retracted, ghost placed);
        flight_imitation.py:113-144 (root pose + linear velocity from the reference, wings on the beat pattern at a
        random phase)."""
        m = self.model
        ids = np.asarray(ids)
        self._load_snippet(ids)
        n = len(ids)
        qpos = np.tile(m.qpos0, (n, 1))
        qvel = None
        ref0 = self._ref_at(np.zeros(n, np.int64), ids)
        qpos[:, self._root_q:self._root_q + 7] = ref0[:, :7]
        qpos[:, self._ghost_q:self._ghost_q + 7] = ref0[:, :7] + np.concatenate([self.task._ghost_offset, np.zeros(4)])
        if self._variant == 'walk':
            if not self._inference_mode:          # full-body start pose from the snippet (walk_imitation.py:117-118)
                qpos[:, self._mocap_qadr] = self._bank.qpos[self._env_bank_off[ids], 7:]
            qpos[:, self._wing_qadr] = self._wing_spring
            if self._reset_noise > 0:
                qpos[:, self._leg_act_qadr] += self._rs.uniform(-self._reset_noise, self._reset_noise, (n, len(self._leg_act_qadr)))
        else:
            wq, wv = self._wbpg.reset(ids, self._rs.uniform(size=n))
            qpos[:, self._wing_qadr] = wq
            qvel = np.zeros((n, m.nv))
            qvel[:, self._wing_vadr] = wv
            qvel[:, self._root_v:self._root_v + 3] = ref0[:, 7:10]
            self._wing_qpos_host[ids] = wq
        if hold:
            self._sim.reset_hold(ids, qpos, qvel)
        else:
            self._sim.reset(qpos=qpos, qvel=qvel, env_ids=None if n == self.n_envs else ids)
        self._step_counter[ids] = 0
        self._time[ids] = 0.0
        self._needs_reset[ids] = False
        self.n_resets += n

    def _device_step(self, action):
        """one control step with the task hooks on the device: actions in, observation rows + (reward, discount, step_type) out."""
        resetting = self._needs_reset.copy()
        if resetting.any():
            self._load_snippet(np.nonzero(resetting)[0])            # a trajectory set since the last reset re-uploads the programs
            if self._program_switched:                              # ... which restarts every env on the device (fb_task_reset_all)
                resetting[:] = True
                self._program_switched = False
        if self._variant == 'flight' and resetting.any():           # the wing-beat phases the host path would draw at these resets
            ids = np.nonzero(resetting)[0]
            self._sim.task_uniforms(ids, self._rs.uniform(size=len(ids)))
        self._sim.task_step(action, self._n_sub)
        self._sim.task_read(self._rec, self._out4)
        flags = self._rec[:, self._obs_slices['_scalars']][:, 0].astype(np.int64)      # bits 1, 2: capacity overflows (counted, not a termination)
        self.n_capacity_overflows += int(((flags & 6) != 0).sum())
        self.h2d_bytes_per_step = action.nbytes
        self.d2h_bytes_per_step = self._rec.nbytes + self._out4.nbytes
        self._step_counter = np.where(resetting, 0, self._step_counter + 1)
        self._time = np.where(resetting, 0.0, self._time + self._control_timestep)
        self.n_resets += int(resetting.sum())
        step_type = self._out4[:, 2].astype(np.int64)
        self._needs_reset = step_type == int(StepType.LAST)
        return TimeStep(step_type,
                        self._out4[:, 0].astype(np.float64), self._out4[:, 1].astype(np.float64), self._observation(self._rec))

    def request_reset(self, env_ids):
        """Restart the listed envs at the next step (their action of that step is dropped and they report FIRST), whatever their
        episode state -- an actor restarting single environments."""
        ids = np.asarray(env_ids, np.int64).reshape(-1)
        if len(ids) == 0:
            return
        if self._device_task:
            self._sim.task_request_reset(ids)
        self._needs_reset[ids] = True

    def set_reset_noise(self, amp):
        """U(-amp, amp) rad added to the actuated leg joints of the start pose at the resets from now on (0: the reference's exact
        start pose).  `reset_noise=` of the factory sets the initial value; a batch is typically started with noise once, so that the
        envs decorrelate, and reset exactly afterwards."""
        if self._device_task and amp > 0 and not self._reset_noise > 0:
            raise ValueError('the device task program was uploaded without a noise joint list: create the env with reset_noise > 0')
        if self._device_task:
            self._sim.task_set_reset_noise(amp)
        else:
            self._reset_noise = float(amp)

    def device_reset_count(self):
        """episodes started so far, summed over the envs (device-side task logic: read from the device's episode counters)"""
        return int(self._sim.task_episodes().sum()) if self._device_task else int(self.n_resets)

    def step_device(self, action):
        """Device-resident control step for a policy that lives on the GPU: `action` is a CUDA tensor / array exposing
        `__cuda_array_interface__` (fp32, [n_envs, n_action], on this env's device); returns zero-copy torch views
        (observation rows [n_envs, obs_dim], out [n_envs, 4] = reward, discount, step_type, 0) of the library's buffers, valid
        until the next step.  Nothing is copied to the host and the call does not synchronise: work is ordered on
        `env.physics.stepper.stream` (wrap it in `torch.cuda.ExternalStream` to order a policy after it).  Requires
        `device_task=True`; `observation_layout()` names the columns."""
        if not self._device_task:
            raise RuntimeError('step_device needs device_task=True (task hooks on the device)')
        import torch
        if self._rec is None:
            self._load_snippet(np.arange(self.n_envs))
        cai = action.__cuda_array_interface__
        assert tuple(cai['shape']) == (self.n_envs, self._action_spec.shape[0]) and cai['typestr'] == '<f4', cai
        self._sim.task_step(cai['data'][0], self._n_sub, is_device=True)
        self._needs_reset[:] = False            # (host mirror of the device's reset flags: not read back on this path)
        if getattr(self, '_dev_views', None) is None:
            obs_ptr, dim, out_ptr = self._sim.task_ptrs()

            class _View:
                def __init__(self, ptr, shape):
                    self.__cuda_array_interface__ = {'shape': shape, 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}
            dev = f'cuda:{self._device}'
            self._dev_views = (torch.as_tensor(_View(obs_ptr, (self.n_envs, dim)), device=dev),
                               torch.as_tensor(_View(out_ptr, (self.n_envs, 4)), device=dev))
        return self._dev_views

    # ---------------------------------------------------------------------------------- eyes
    # cameras `eye_right` / `eye_left` on the head (reference fruitfly.xml:335-336; the head body keeps its XML frame)
    _EYE_CAMERAS = (('walker/right_eye', (0.0219, 0.0131, 0.0), (0.474, 0.688, -0.344, -0.429)),
                    ('walker/left_eye', (-0.0219, 0.0131, 0.0), (0.474, 0.688, 0.344, 0.429)))

    def enable_eyes(self, size=32, fovy=150.0, terrain_shape=None, half_size=20.0, z_offset=0.0):
        """Turn on the eye-camera observables `walker/right_eye`, `walker/left_eye` (reference `fruitfly.py:729-745`; 32 x 32,
        fovy 150 in `tasks/vision_flight.py:23-24`), rendered on the device by a ray caster over an optional per-env
        heightfield `terrain_shape = (nrow, ncol)` spanning [-half_size, half_size]^2 (`flybody_b200.arenas`), the ground plane
        and a sky.  `render_eyes()` returns them for the current state; they are not part of `step()`'s observation dict."""
        quats = [np.asarray(q, np.float64) / np.linalg.norm(q) for _, _, q in self._EYE_CAMERAS]
        head = self.model.body_id('walker/head')
        nrow, ncol = terrain_shape if terrain_shape is not None else (0, 0)
        self._sim.eye_program([head, head], [p for _, p, _ in self._EYE_CAMERAS], quats, fovy_deg=fovy, size=size, nrow=nrow, ncol=ncol,
                              half_size=half_size, z_offset=z_offset)
        self._eyes_on = True

    def set_terrain(self, env_ids, heights):
        """heights [n, nrow, ncol] (world units; row = y, col = x) of the listed envs' terrains, as seen by the eyes."""
        self._sim.hfield_write(env_ids, heights)

    def render_eyes(self):
        """OrderedDict {'walker/right_eye', 'walker/left_eye'}: uint8 [n_envs, size, size, 3] from the current poses."""
        if not getattr(self, '_eyes_on', False):
            raise RuntimeError('call enable_eyes() first')
        img = self._sim.render_eyes()
        return collections.OrderedDict((name, img[:, k]) for k, (name, _, _) in enumerate(self._EYE_CAMERAS))

    def observation_layout(self):
        """{observable name: (column slice, shape)} of the observation rows `step_device` returns."""
        return {k: (sl, self._obs_shapes[k]) for k, sl in self._obs_slices.items() if not k.startswith('_')}

    def reset(self):
        if self._device_task:
            self._load_snippet(np.arange(self.n_envs))
            self._sim.task_reset_all()
            self._needs_reset[:] = True
            ts = self._device_step(np.zeros((self.n_envs, self._action_spec.shape[0]), np.float32))    # every env held: FIRST
            return self._unbatch(ts, first=True)
        self._reset_envs(np.arange(self.n_envs))
        self._sim.task_inputs(self._step_counter, np.ones(self.n_envs, np.uint8))
        rec = self._sim.read_task_obs(self._rec)
        obs = self._observation(rec)
        N = self.n_envs
        ts = TimeStep(np.full(N, StepType.FIRST), np.zeros(N), np.ones(N), obs)
        return self._unbatch(ts, first=True)

    def step(self, action):
        m = self.model
        N = self.n_envs
        if self._device_action_map:
            action = np.asarray(action, np.float32).reshape(N, -1)          # not modified on the host
        else:
            action = np.array(action, np.float64, copy=True).reshape(N, -1)
        assert action.shape[1] == self._action_spec.shape[0], f'action must have {self._action_spec.shape[0]} entries'
        if self._device_task:
            return self._unbatch(self._device_step(action))
        # auto-reset of envs whose last step was LAST (composer.Environment semantics); their action is ignored
        resetting = self._needs_reset.copy()
        if resetting.any():
            self._reset_envs(np.nonzero(resetting)[0], hold=True)
        # before_step (walk_imitation.py:138-150, flight_imitation.py:146-168, base.py:197-201)
        step = np.round(self._time / self._control_timestep).astype(np.int64)
        step = np.minimum(step, self._ref_len - 1)
        step = np.where(resetting, 0, step)
        ghost = self._ref_at(step)
        ghost[:, :3] += self.task._ghost_offset
        ghost = ghost.astype(np.float32)
        ghost[resetting, 7:] = 0.0
        if self._variant == 'flight':
            # wing-beat pattern at the requested frequency, as a position target turned into a force command
            wb = self._wbpg
            freq = wb.base_beat_freq * (1 + wb.rel_freq_range * action[:, -1])
            target = wb.step(freq, active=~resetting)
            wi = self._action_indices['wings']
            action[:, wi] += target - self._wing_qpos_host
        self._sim.write_state(st.QPOS, np.arange(self._ghost_q, self._ghost_q + 7), ghost[:, :7])
        self._sim.write_state(st.QVEL, np.arange(self._ghost_v, self._ghost_v + 6), ghost[:, 7:])
        self._step_counter += np.where(resetting, 0, 1)
        if self._device_action_map:
            self._sim.set_control(action)
            self.h2d_bytes_per_step = action.nbytes + ghost.nbytes
        else:
            action[np.isnan(action)] = 0.0
            ctrl = np.zeros((N, m.nu), np.float32)
            ctrl[:, self._ctrl_of_action] = action[:, :len(self._ctrl_of_action)]
            self._sim.set_control(ctrl)
            self.h2d_bytes_per_step = ctrl.nbytes + ghost.nbytes
        # n_sub_steps x physics.step()
        self._sim.task_inputs(self._step_counter, resetting)
        self._sim.step(self._n_sub)
        rec = self._sim.read_task_obs(self._rec)
        self.d2h_bytes_per_step = rec.nbytes
        self._time = np.where(resetting, 0.0, self._time + self._control_timestep)
        obs = self._observation(rec)
        step_type, reward, discount, self._needs_reset = self._task_after(rec, obs, resetting, self._time, ghost)
        return self._unbatch(TimeStep(step_type, reward, discount, obs))

    def _task_after(self, rec, obs, resetting, time, ghost):
        """check_termination / reward / discount on the observation record of the step just taken (walk_imitation.py:152-203,
        flight_imitation.py:170-226, base.py:203-225) -> step_type, reward, discount, needs_reset.  (The device-side task
        program `ktask_after` is the same logic; tests/test_device_task.py holds the two against each other.)"""
        N = self.n_envs
        sl = self._obs_slices
        step_now = np.round(time / self._control_timestep).astype(np.int64)
        com_dist = np.linalg.norm(obs['walker/ref_displacement'][:, 0], axis=1)
        reached_end = step_now == self._episode_steps
        scal = rec[:, sl['_scalars']]
        # FB_FLAGS bit 0 = non-finite / diverged state (reference base.py:222-225 terminates on it); bits 1, 2 = contact /
        # constraint-row capacity overflows, which are counted, not treated as bad physics
        flags = scal[:, 0].astype(np.int64)
        self.n_capacity_overflows += int(((flags & 6) != 0).sum())
        bad = ((flags & 1) != 0) | ~(np.sqrt(scal[:, 1].astype(np.float64)) <= _TERMINAL_QACC)
        if self._variant == 'walk':
            linvel = np.linalg.norm(rec[:, sl['_velocimeter_now']], axis=1)
            angvel = np.linalg.norm(rec[:, sl['_gyro_now']], axis=1)
            terminate = (linvel > _TERMINAL_LINVEL) | (angvel > _TERMINAL_ANGVEL) | reached_end | \
                        (com_dist > self.task._terminal_com_dist) | bad
            if self._inference_mode:
                reward = np.ones(N)                           # reward factors == (1,)
            else:
                reward = np.prod(self._walk_reward_factors(rec, np.minimum(step_now, self._ref_len - 1)), axis=1)
        else:
            self._wing_qpos_host = rec[:, sl['walker/joints_pos']][:, self._wing_in_obs].astype(np.float64)
            height = rec[:, sl['_root_pose']][:, 2]
            terminate = (height < _TERMINAL_HEIGHT) | (com_dist > self.task._terminal_com_dist) | reached_end | bad
            # reward factors: CoM displacement and orientation error to the ghost (legs are disabled: third factor == 1)
            # (the ghost is a free body that coasts with its reference velocity during the substeps; the reference reads its
            # pose after the step, flight_imitation.py:174-176)
            ghost_com = root2com(rec[:, sl['_ghost_pose']].astype(np.float64))
            disp = np.linalg.norm(ghost_com - rec[:, sl['_subtree_com']], axis=1)
            qd = quat_dist_short_arc(np.array([1.0, 0, 0, 0]), obs['walker/ref_root_quat'][:, 0].astype(np.float64))
            reward = linear_tolerance(disp, 0.4) * linear_tolerance(qd, np.pi)
        discount = np.where(terminate & ~reached_end, 0.0, 1.0)
        last = terminate | (time >= self._time_limit - 1e-9)
        step_type = np.where(last, StepType.LAST, StepType.MID)
        # rows that were reset this call report FIRST (their action was ignored)
        step_type = np.where(resetting, StepType.FIRST, step_type)
        reward = np.where(resetting, 0.0, reward)
        discount = np.where(resetting, 1.0, discount)
        return step_type, reward, discount, last & ~resetting

    def _walk_reward_factors(self, rec, step):
        """[N, 4 + 6]: DeepMimic factors (weights 20, 1, 1, 1) and one wing-retraction factor per wing joint
        (reference walk_imitation.py:152-177)."""
        sl = self._obs_slices
        nj, ns = len(self._mocap_qadr), len(self._mocap_sites)
        f64 = lambda k, shape=None: rec[:, sl[k]].astype(np.float64).reshape((self.n_envs,) + (shape or (-1,)))
        wf = rw.get_walker_features(f64('_root_pose'), f64('_mocap_qpos'), f64('_root_qvel'), f64('_mocap_qvel'),
                                    f64('_mocap_sites', (ns, 3)), f64('_mocap_axes', (nj, 3)))
        rows = self._env_bank_off + step
        b = self._bank
        rf = {'com': b.qpos[rows, :3], 'qvel': b.qvel[rows], 'root2site': b.root2site[rows],
              'joint_quat': np.concatenate([b.qpos[rows, None, 3:7], b.joint_quat[rows]], 1)}
        factors = rw.reward_factors_deep_mimic(wf, rf, weights=(20, 1, 1, 1))
        retract = linear_tolerance(f64('_wing_qpos') - self._wing_spring, 3.0)
        return np.concatenate([factors, retract], 1)

    # ---------------------------------------------------------------------------- observations
    def _observation(self, rec):
        """Views into the device-evaluated observation rows (fp32; the reference returns float64 copies)."""
        N = self.n_envs
        obs = collections.OrderedDict()
        for k, sl in self._obs_slices.items():
            if not k.startswith('_'):
                obs[k] = rec[:, sl].reshape((N,) + self._obs_shapes[k])
        return obs

    def _unbatch(self, ts, first=False):
        if self._batched:
            return ts
        obs = collections.OrderedDict((k, v[0]) for k, v in ts.observation.items())
        if first or ts.step_type[0] == StepType.FIRST:
            return TimeStep(StepType.FIRST, None, None, obs)
        return TimeStep(StepType(int(ts.step_type[0])), float(ts.reward[0]), float(ts.discount[0]), obs)

    def close(self):
        self._sim.close()


def walk_imitation(ref_path=None, force_actuators=False, disable_wings=True, traj_indices=None, random_state=None,
                   terminal_com_dist=0.3, joint_filter=0.01, n_envs=None, device=0, lib_path=None, reset_noise=0.0,
                   seed=0, max_reference_steps=None, device_task=False):
    """Batched `flybody.fly_envs.walk_imitation` (reference `fly_envs.py:100-155`).  With `ref_path` (an HDF5 walking
    dataset, or its `.npz` conversion, see `trajectory_loaders`) every env tracks its own snippet, starts from the
    snippet's full-body pose and is rewarded with the DeepMimic factors; without it the task runs in inference mode on
    the synthetic straight walk, reward 1 (`fly_envs.py:127-135`)."""
    model = None
    if force_actuators or not disable_wings or joint_filter != 0.01:
        from .flymodel import model_for
        model = model_for('walk', force_actuators=force_actuators, use_wings=True if not disable_wings else None,
                          joint_filter=None if joint_filter == 0.01 else joint_filter)
    tg = None
    if ref_path is not None:
        tg = HDF5WalkingTrajectoryLoader(path=ref_path, random_state=random_state, traj_indices=traj_indices)
    return BatchedFlyEnv('walk', n_envs, device=device, terminal_com_dist=terminal_com_dist, time_limit=10.0,
                         future_steps=64, lib_path=lib_path, reset_noise=reset_noise, seed=seed, traj_generator=tg,
                         inference_mode=ref_path is None, max_reference_steps=max_reference_steps, device_task=device_task, model=model)


def flight_imitation(ref_path=None, wpg_pattern_path=None, force_actuators=False, disable_legs=True, traj_indices=None,
                     randomize_start_step=True, joint_filter=0.0, future_steps=5, random_state=None, terminal_com_dist=2.0,
                     n_envs=None, device=0, lib_path=None, seed=0, device_task=False):
    """Batched `flybody.fly_envs.flight_imitation` (reference `fly_envs.py:30-97`): wing-beat-pattern-generator flight
    tracking, 4 substeps of 5e-5 s per control step, 12 actions (head 3, wings 6, abdomen 2, beat frequency 1).  With
    `ref_path` every env tracks its own (randomly cut) CoM trajectory of the flight dataset."""
    model = None
    if force_actuators or not disable_legs or joint_filter != 0.0:      # a non-default FruitFly: compiled on first use (flymodel.model_for)
        from .flymodel import model_for
        model = model_for('flight', force_actuators=force_actuators, use_legs=True if not disable_legs else None,
                          joint_filter=None if joint_filter == 0.0 else joint_filter)
    tg = None
    if ref_path is not None:
        tg = HDF5FlightTrajectoryLoader(path=ref_path, traj_indices=traj_indices, randomize_start_step=randomize_start_step,
                                        random_state=random_state)
    return BatchedFlyEnv('flight', n_envs, device=device, terminal_com_dist=terminal_com_dist, time_limit=0.6,
                         future_steps=future_steps, lib_path=lib_path, seed=seed, wpg_pattern_path=wpg_pattern_path,
                         traj_generator=tg, device_task=device_task, model=model)


def vision_guided_flight(wpg_pattern_path=None, bumps_or_trench='bumps', force_actuators=False, disable_legs=True, random_state=None,
                         joint_filter=0.0, n_envs=None, device=0, lib_path=None, seed=0, **kwargs_arena):
    """Batched `flybody.fly_envs.vision_guided_flight` (reference `fly_envs.py:194-246`): 'bumps' or 'trench' terrain, eye
    cameras, wing-beat pattern generator, fatal ground contacts -- `flybody_b200.vision_env.BatchedVisionFlightEnv` (SURVEY.md
    8(f).1; terrain contacts against the fp64 oracle and eyes against the camera model in tests/test_hfield.py, test_eyes.py and on the
    B200 in tests/test_gpu_parity.py).  `terrain_bank=K, device_task=True` run the task's hooks on the device (fb_task_* kind 2)."""
    if force_actuators or not disable_legs or joint_filter != 0.0:
        raise NotImplementedError('only the default vision_guided_flight model variant is compiled (flybody_b200/assets/fly_vision.npz)')
    from .vision_env import BatchedVisionFlightEnv
    return BatchedVisionFlightEnv(n_envs, bumps_or_trench=bumps_or_trench, wpg_pattern_path=wpg_pattern_path, device=device, lib_path=lib_path,
                                  seed=seed, **kwargs_arena)


def walk_on_ball(*args, **kwargs):
    """`flybody.fly_envs.walk_on_ball` (reference `fly_envs.py:158-191`): outside the hot-path scope."""
    raise NotImplementedError(walk_on_ball.__doc__)


def template_task(*args, **kwargs):
    """`flybody.fly_envs.template_task` (reference `fly_envs.py:249-310`): outside the hot-path scope."""
    raise NotImplementedError(template_task.__doc__)
