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

import numpy as np

from . import stepper as st
from .dm_env_shim import Array, BoundedArray, StepType, TimeStep
from .flymodel import load_model

_WALK_CONTROL_TIMESTEP = 2e-3      # reference tasks/constants.py:10-13
_WALK_PHYSICS_TIMESTEP = 2e-4
_TERMINAL_LINVEL = 50.0
_TERMINAL_ANGVEL = 200.0
_FLY_CONTROL_TIMESTEP = 2e-4       # tasks/constants.py:16-19
_FLY_PHYSICS_TIMESTEP = 5e-5
_TERMINAL_HEIGHT = 0.2
_TERMINAL_QACC = 1e14              # tasks/constants.py:21
_ACTION_CLASS_ORDER = ('adhesion', 'head', 'mouth', 'antennae', 'wings', 'abdomen', 'legs', 'user')  # fruitfly.py:25-32


# --- quaternion helpers on [..., 4] arrays (reference flybody/quaternions.py:13-76) -----------
def mult_quat(a, b):
    aw, ax, ay, az = np.moveaxis(a, -1, 0)
    bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def reciprocal_quat(q):
    return q * np.array([1.0, -1, -1, -1]) / np.sum(q * q, -1, keepdims=True)


def constant_speed_trajectory(n_steps, speed, yaw_speed=0.0, init_pos=(0, 0, 0.1278), init_heading=0.0,
                              body_rot_angle_y=0.0, body_rot_angle_x=0.0, control_timestep=0.002):
    """reference `tasks/synthetic_trajectories.py:10-70` (mju_quat2Vel restated: `quaternions.py:358-382`)."""
    qpos = np.zeros((n_steps, 7))
    qvel = np.zeros((n_steps, 6))
    qpos[0, :3] = init_pos
    qpos[:, 2] = init_pos[2]
    ya, xa = np.deg2rad(body_rot_angle_y), np.deg2rad(body_rot_angle_x)
    qpos[0, 3:] = [np.cos(ya / 2), 0.0, np.sin(ya / 2), 0.0]
    qpos[0, 3:] = mult_quat(np.array([np.cos(xa / 2), np.sin(xa / 2), 0.0, 0]), qpos[0, 3:])
    dq = np.array([np.cos(init_heading / 2), 0, 0, np.sin(init_heading / 2)])
    qpos[0, 3:] = mult_quat(dq, qpos[0, 3:])
    qvel[0, :2] = speed * np.array([np.cos(init_heading), np.sin(init_heading)])
    dtheta = yaw_speed * control_timestep
    dq = np.array([np.cos(dtheta / 2), 0, 0, np.sin(dtheta / 2)])
    axis = dq[1:]
    sin_a_2 = np.linalg.norm(axis)
    vel = np.zeros(3)
    if sin_a_2 > 0:
        speed_ang = 2 * np.arctan2(sin_a_2, dq[0])
        if speed_ang > np.pi:
            speed_ang -= 2 * np.pi
        vel = axis / sin_a_2 * speed_ang / 1.0
    qvel[:, 3:] = vel
    M = np.array([[np.cos(dtheta), -np.sin(dtheta)], [np.sin(dtheta), np.cos(dtheta)]])
    for i in range(1, n_steps):
        qvel[i, :2] = M @ qvel[i - 1, :2]
        qpos[i, :2] = qpos[i - 1, :2] + qvel[i, :2] * control_timestep
        qpos[i, 3:] = mult_quat(dq, qpos[i - 1, 3:])
    return qpos, qvel


class InferenceWalkingTrajectoryLoader:
    """reference `tasks/trajectory_loaders.py:267-309`."""

    def __init__(self):
        qpos, qvel = constant_speed_trajectory(n_steps=300, speed=2, init_pos=(0, 0, 0.1278),
                                               control_timestep=_WALK_CONTROL_TIMESTEP)
        self.set_next_trajectory(qpos, qvel)

    def set_next_trajectory(self, qpos, qvel):
        self._snippet = {'qpos': np.asarray(qpos, np.float64), 'qvel': np.asarray(qvel, np.float64)}

    def get_trajectory(self, traj_idx=None):
        return self._snippet

    def get_joint_names(self):
        return []

    def get_site_names(self):
        return []


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


class WalkImitationTask:
    """Batched `WalkImitation` (reference `tasks/walk_imitation.py:19-203`, `tasks/base.py:22-268,367-428`)."""

    def __init__(self, env, traj_generator, terminal_com_dist, future_steps, time_limit, inference_mode=True):
        self._env = env
        self._traj_generator = traj_generator
        self._terminal_com_dist = terminal_com_dist
        self._future_steps = future_steps
        self._time_limit = time_limit
        self._inference_mode = inference_mode
        self._max_episode_steps = round(time_limit / _WALK_CONTROL_TIMESTEP) + 1
        self._ghost_offset = np.zeros(3)

    name = 'FruitFlyTask'


class BatchedFlyEnv:
    """dm_env-shaped environment over N lock-stepped flies (composer.Environment stand-in)."""

    def __init__(self, variant, n_envs, device=0, terminal_com_dist=0.3, time_limit=10.0, future_steps=64,
                 lib_path=None, reset_noise=0.0, seed=0, traj_generator=None):
        assert variant == 'walk'
        self._batched = n_envs is not None
        self.n_envs = int(n_envs) if self._batched else 1
        self.model = load_model(variant)
        m = self.model
        self._sim = st.BatchedStepper(m, self.n_envs, device=device, lib_path=lib_path)
        self._control_timestep = _WALK_CONTROL_TIMESTEP
        self._physics_timestep = float(m.opt_timestep)
        self._n_sub = int(round(self._control_timestep / self._physics_timestep))
        self._time_limit = time_limit
        self.physics = _PhysicsFacade(self)
        self.task = WalkImitationTask(self, traj_generator or InferenceWalkingTrajectoryLoader(), terminal_com_dist,
                                      future_steps, time_limit)
        self._rs = np.random.RandomState(seed)
        self._reset_noise = reset_noise
        N = self.n_envs
        # --- action <-> ctrl maps (reference fruitfly.py:342-379, 532-579)
        ci = m.meta['ctrl_indices']
        idx = []
        for key in _ACTION_CLASS_ORDER:
            if ci.get(key):
                idx.extend(ci[key])
        self._ctrl_of_action = np.asarray(idx, np.int64)
        names = [m.meta['actuator_names'][i].split('/')[-1] for i in idx]
        rng = m.actuator_ctrlrange[idx]
        self._action_spec = BoundedArray((len(idx),), np.float64, rng[:, 0], rng[:, 1], name='\t'.join(names))
        # --- index tables
        jn = m.meta['jnt_names']
        self._root_q = m.jnt_qposadr_of('walker/')
        self._root_v = m.jnt_dofadr_of('walker/')
        self._ghost_q = m.jnt_qposadr_of('ghost/')
        self._ghost_v = m.jnt_dofadr_of('ghost/')
        obsj = [jn.index(n) for n in m.meta['observable_joints']]
        self._obs_qadr = m.jnt_qposadr[obsj]
        self._obs_vadr = m.jnt_dofadr[obsj]
        self._wing_qadr = np.array([m.jnt_qposadr_of(f'walker/wing_{a}_{s}') for s in ('left', 'right') for a in ('yaw', 'roll', 'pitch')])
        self._wing_spring = m.qpos_spring[self._wing_qadr]
        sn = m.meta['site_names']
        app = [f'walker/claw_T{k}_{s}' for k in (1, 2, 3) for s in ('left', 'right')] + ['walker/head']
        self._app_sites = np.array([sn.index(n) for n in app])
        sens = m.meta['sensor_names']
        def sd(names_):
            out = []
            for n_ in names_:
                i = sens.index('walker/' + n_)
                out.extend(range(m.sensor_adr[i], m.sensor_adr[i] + m.sensor_dim[i]))
            return np.array(out)
        legs = [f'T{k}_{s}' for k in (1, 2, 3) for s in ('left', 'right')]
        self._sd = dict(accelerometer=sd(['accelerometer']), gyro=sd(['gyro']), velocimeter=sd(['velocimeter']),
                        force=sd([f'force_tarsus_{l}' for l in legs]), touch=sd([f'touch_claw_{l}' for l in legs]))
        self._leg_act_qadr = np.array([m.jnt_qposadr[m.actuator_trnid[i]] for i in range(m.nu)
                                       if m.actuator_trntype[i] == 0 and any(t in m.meta['actuator_names'][i] for t in ('T1', 'T2', 'T3'))])
        self._future = future_steps + 1
        self._rec = None
        self._program_ref_id = None
        # --- per-env episode state
        self._step_counter = np.zeros(N, np.int64)
        self._time = np.zeros(N)
        self._needs_reset = np.ones(N, bool)
        self._first_after_reset = np.zeros(N, bool)
        self._ref_qpos = None
        self.h2d_bytes_per_step = 0
        self.d2h_bytes_per_step = 0
        self.n_resets = 0

    # ---------------------------------------------------------------------------------- specs
    def action_spec(self):
        return self._action_spec

    def observation_spec(self):
        f = self.task._future_steps + 1
        shapes = collections.OrderedDict([
            ('walker/accelerometer', (3,)), ('walker/actuator_activation', (self.model.na,)),
            ('walker/appendages_pos', (21,)), ('walker/force', (18,)), ('walker/gyro', (3,)),
            ('walker/joints_pos', (len(self._obs_qadr),)), ('walker/joints_vel', (len(self._obs_vadr),)),
            ('walker/touch', (6,)), ('walker/velocimeter', (3,)), ('walker/world_zaxis', (3,)),
            ('walker/ref_displacement', (f, 3)), ('walker/ref_root_quat', (f, 4))])
        lead = (self.n_envs,) if self._batched else ()
        return collections.OrderedDict((k, Array(lead + v, np.float32, name=k)) for k, v in shapes.items())

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
        sd = self._sd
        lists = list(self._app_sites) + list(self._obs_qadr) + list(self._obs_vadr)
        o_app, o_q, o_v = 0, len(self._app_sites), len(self._app_sites) + len(self._obs_qadr)
        f = self._future
        items = [(st.OBS_SENSOR_MEAN, sd['accelerometer'][0], 3), (st.OBS_ACT, 0, m.na), (st.OBS_SITES_EGO, o_app, len(self._app_sites)),
                 (st.OBS_SENSOR_MEAN, sd['force'][0], len(sd['force'])), (st.OBS_SENSOR_MEAN, sd['gyro'][0], 3),
                 (st.OBS_QPOS, o_q, len(self._obs_qadr)), (st.OBS_QVEL, o_v, len(self._obs_vadr)),
                 (st.OBS_SENSOR_MEAN, sd['touch'][0], len(sd['touch'])), (st.OBS_SENSOR_MEAN, sd['velocimeter'][0], 3),
                 (st.OBS_ROOT_ZAXIS, 0, 3), (st.OBS_REF_DISP, 0, f), (st.OBS_REF_QUAT, 0, f),
                 (st.OBS_SENSOR_NOW, sd['velocimeter'][0], 3), (st.OBS_SENSOR_NOW, sd['gyro'][0], 3), (st.OBS_SCALARS, 0, 3)]
        dim = self._sim.obs_program(items, lists, m.body_id('walker/thorax'), self._n_sub, self._ref_qpos[:, :7])
        names = ['walker/accelerometer', 'walker/actuator_activation', 'walker/appendages_pos', 'walker/force', 'walker/gyro',
                 'walker/joints_pos', 'walker/joints_vel', 'walker/touch', 'walker/velocimeter', 'walker/world_zaxis',
                 'walker/ref_displacement', 'walker/ref_root_quat', '_velocimeter_now', '_gyro_now', '_scalars']
        widths = [3, m.na, 3 * len(self._app_sites), len(sd['force']), 3, len(self._obs_qadr), len(self._obs_vadr),
                  len(sd['touch']), 3, 3, 3 * f, 4 * f, 3, 3, 3]
        assert sum(widths) == dim
        off = np.concatenate([[0], np.cumsum(widths)])
        self._obs_slices = {n_: slice(int(off[i]), int(off[i + 1])) for i, n_ in enumerate(names)}
        N = self.n_envs
        try:
            import torch
            self._rec = torch.empty((N, dim), dtype=torch.float32).pin_memory().numpy() if torch.cuda.is_available() \
                else np.empty((N, dim), np.float32)
        except Exception:
            self._rec = np.empty((N, dim), np.float32)

    def _load_snippet(self):
        snip = self.task._traj_generator.get_trajectory(traj_idx=None)
        self._ref_qpos = snip['qpos']
        self._ref_qvel = snip['qvel']
        if self._program_ref_id is not self._ref_qpos:
            self._program_ref_id = self._ref_qpos
            self._upload_program()
        snippet_steps = self._ref_qpos.shape[0] - self.task._future_steps - 1
        self._episode_steps = min(self.task._max_episode_steps, snippet_steps)      # walk_imitation.py:104-105

    def _reset_envs(self, ids, hold=False):
        """initialize_episode (walk_imitation.py:112-136): root <- ref_qpos[0], wings retracted, ghost placed."""
        m = self.model
        self._load_snippet()
        n = len(ids)
        qpos = np.tile(m.qpos0, (n, 1))
        qpos[:, self._root_q:self._root_q + 7] = self._ref_qpos[0, :7]
        qpos[:, self._wing_qadr] = self._wing_spring
        qpos[:, self._ghost_q:self._ghost_q + 7] = self._ref_qpos[0, :7]
        if self._reset_noise > 0:
            qpos[:, self._leg_act_qadr] += self._rs.uniform(-self._reset_noise, self._reset_noise, (n, len(self._leg_act_qadr)))
        if hold:
            self._sim.reset_hold(ids, qpos)
        else:
            self._sim.reset(qpos=qpos, qvel=None, env_ids=None if n == self.n_envs else ids)
        self._step_counter[ids] = 0
        self._time[ids] = 0.0
        self._needs_reset[ids] = False
        self.n_resets += n

    def reset(self):
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
        action = np.array(action, np.float64, copy=True).reshape(N, -1)
        # auto-reset of envs whose last step was LAST (composer.Environment semantics); their action is ignored
        resetting = self._needs_reset.copy()
        if resetting.any():
            self._reset_envs(np.nonzero(resetting)[0], hold=True)
        # before_step (walk_imitation.py:138-150, base.py:197-201)
        step = np.round(self._time / self._control_timestep).astype(np.int64)
        step = np.minimum(step, self._ref_qpos.shape[0] - 1)
        step = np.where(resetting, 0, step)
        ghost = np.concatenate([self._ref_qpos[step, :7], self._ref_qvel[step, :6]], 1).astype(np.float32)
        ghost[resetting, 7:] = 0.0
        self._sim.write_state(st.QPOS, np.arange(self._ghost_q, self._ghost_q + 7), ghost[:, :7])
        self._sim.write_state(st.QVEL, np.arange(self._ghost_v, self._ghost_v + 6), ghost[:, 7:])
        action[np.isnan(action)] = 0.0
        self._step_counter += np.where(resetting, 0, 1)
        ctrl = np.zeros((N, m.nu), np.float32)
        ctrl[:, self._ctrl_of_action] = action
        self._sim.set_control(ctrl)
        self.h2d_bytes_per_step = ctrl.nbytes + ghost.nbytes
        # n_sub_steps x physics.step()
        self._sim.task_inputs(self._step_counter, resetting)
        self._sim.step(self._n_sub)
        rec = self._sim.read_task_obs(self._rec)
        self.d2h_bytes_per_step = rec.nbytes
        self._time = np.where(resetting, 0.0, self._time + self._control_timestep)
        obs = self._observation(rec)
        sl = self._obs_slices
        # check_termination / reward / discount (walk_imitation.py:152-203, base.py:203-225)
        linvel = np.linalg.norm(rec[:, sl['_velocimeter_now']], axis=1)
        angvel = np.linalg.norm(rec[:, sl['_gyro_now']], axis=1)
        step_now = np.round(self._time / self._control_timestep).astype(np.int64)
        com_dist = np.linalg.norm(obs['walker/ref_displacement'][:, 0], axis=1)
        reached_end = step_now == self._episode_steps
        scal = rec[:, sl['_scalars']]
        bad = (scal[:, 0] != 0) | ~(np.sqrt(scal[:, 1].astype(np.float64)) <= _TERMINAL_QACC)
        terminate = (linvel > _TERMINAL_LINVEL) | (angvel > _TERMINAL_ANGVEL) | reached_end | \
                    (com_dist > self.task._terminal_com_dist) | bad
        reward = np.ones(N)                                   # inference mode: reward factors == (1,)
        discount = np.where(terminate & ~reached_end, 0.0, 1.0)
        last = terminate | (self._time >= self._time_limit - 1e-9)
        step_type = np.where(last, StepType.LAST, StepType.MID)
        # rows that were reset this call report FIRST (their action was ignored)
        step_type = np.where(resetting, StepType.FIRST, step_type)
        reward = np.where(resetting, 0.0, reward)
        discount = np.where(resetting, 1.0, discount)
        self._needs_reset = last & ~resetting
        return self._unbatch(TimeStep(step_type, reward, discount, obs))

    # ---------------------------------------------------------------------------- observations
    def _observation(self, rec):
        """Views into the device-evaluated observation rows (fp32; the reference returns float64 copies)."""
        N, f = self.n_envs, self._future
        obs = collections.OrderedDict()
        for k, sl in self._obs_slices.items():
            if k.startswith('_'):
                continue
            v = rec[:, sl]
            if k == 'walker/ref_displacement':
                v = v.reshape(N, f, 3)
            elif k == 'walker/ref_root_quat':
                v = v.reshape(N, f, 4)
            obs[k] = v
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
                   seed=0):
    """Batched `flybody.fly_envs.walk_imitation` (reference `fly_envs.py:100-155`)."""
    if ref_path is not None:
        raise NotImplementedError('HDF5 reference datasets (h5py) are a "next" row (SURVEY.md 8(f).3); '
                                  'use env.task._traj_generator.set_next_trajectory(qpos, qvel)')
    if force_actuators or not disable_wings or joint_filter != 0.01:
        raise NotImplementedError('only the default walk_imitation model variant is compiled '
                                  '(flybody_b200/assets/fly_walk.npz); recompile with compiler.compile_variant')
    return BatchedFlyEnv('walk', n_envs, device=device, terminal_com_dist=terminal_com_dist, time_limit=10.0,
                         future_steps=64, lib_path=lib_path, reset_noise=reset_noise, seed=seed)
