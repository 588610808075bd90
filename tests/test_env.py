"""dm_env surface of the batched walk_imitation env (contracts of reference tests/test_walking_env.py),
run on the host-emulation build (no GPU here); tests/test_gpu_parity.py repeats the smoke on the B200."""
import json
import os

import numpy as np
import pytest

import __graft_entry__ as ge
from flybody_b200 import fly_envs
from flybody_b200.dm_env_shim import StepType

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'reference_goldens.json')))

# trajectory of reference tests/test_walking_env.py:27-35
n_steps, ctrl_timestep = 200, 0.002
qpos = np.zeros((n_steps, 7))
qpos[:, 0] = np.arange(0, n_steps * ctrl_timestep, ctrl_timestep)
qpos[:, [2, 3]] = [0.14355, 1.]
qvel = np.zeros((n_steps, 6))
qvel[:, 0] = 1.


@pytest.fixture(scope='module')
def emu():
    ge.build()
    return ge.EMU


def test_can_create_env_inference_mode(emu):
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), lib_path=emu)
    assert list(env.observation_spec()) == G['walk_obs_names']
    assert env.action_spec().shape == (G['walk_num_act'],)
    assert env.action_spec().name.split('\t') == G['action_names']
    env.task._traj_generator.set_next_trajectory(qpos, qvel)
    ts = env.reset()
    assert ts.step_type == StepType.FIRST
    for name in G['walk_obs_names']:
        assert isinstance(ts.observation[name], (float, np.ndarray))
        assert ts.observation[name].shape == env.observation_spec()[name].shape
    assert np.isclose(env.control_timestep(), 2e-3)
    assert np.isclose(env.physics.timestep(), 2e-4)


def test_can_step_env_inference_mode(emu):
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), lib_path=emu)
    env.task._traj_generator.set_next_trajectory(qpos, qvel)
    env.reset()
    rs = np.random.RandomState(0)
    for _ in range(30):
        ts = env.step(rs.uniform(-0.5, 0.5, 59))
        assert ts.reward == 1.
        assert all(np.all(np.isfinite(v)) for v in ts.observation.values())


def test_batched_env_autoreset_semantics(emu):
    env = fly_envs.walk_imitation(terminal_com_dist=0.05, n_envs=3, lib_path=emu)    # ghost leaves 0.05 cm after ~13 steps
    ts = env.reset()
    assert ts.observation['walker/joints_pos'].shape == (3, 85)
    rs = np.random.RandomState(1)
    seen_last = seen_first = False
    prev_last = np.zeros(3, bool)
    for _ in range(40):
        ts = env.step(rs.uniform(-0.3, 0.3, (3, 59)))
        # the step after LAST is FIRST with the reset observation (composer.Environment semantics)
        assert np.all(ts.step_type[prev_last] == StepType.FIRST)
        if prev_last.any():
            seen_first = True
            d0 = np.linalg.norm(ts.observation['walker/ref_displacement'][prev_last, 0], axis=1)
            assert np.all(d0 < 1e-6)
        prev_last = ts.step_type == StepType.LAST
        if prev_last.any():
            seen_last = True
            assert np.all(ts.discount[prev_last] == 0.0)       # fatal termination (com distance)
    assert seen_last and seen_first


def test_nan_action_is_zeroed(emu):
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=2, lib_path=emu)
    env.reset()
    a = np.zeros((2, 59))
    a[0, 5] = np.nan                                           # walk_imitation.py:147-148
    ts = env.step(a)
    assert np.all(np.isfinite(ts.observation['walker/joints_pos']))


def test_device_observation_program_matches_numpy_formulas(emu):
    """The device-evaluated observables equal the reference's formulas (fruitfly.py:674-684, base.py:245-268)
    evaluated in numpy on the raw state record."""
    from flybody_b200.fly_envs import mult_quat, reciprocal_quat
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=3, lib_path=emu)
    env.reset()
    rs = np.random.RandomState(3)
    for _ in range(4):
        ts = env.step(rs.uniform(-0.5, 0.5, (3, 59)))
    sim, m = env._sim, env.model
    raw = sim.read_obs()
    lay = sim.obs_layout()
    qpos, qvel = raw[:, lay['qpos']].astype(np.float64), raw[:, lay['qvel']].astype(np.float64)
    xpos = raw[:, lay['root_xpos']].astype(np.float64)
    xmat = raw[:, lay['root_xmat']].astype(np.float64).reshape(3, 3, 3)
    sites = raw[:, lay['site_xpos']].astype(np.float64).reshape(3, -1, 3)
    app = np.einsum('nsi,nij->nsj', sites[:, env._app_sites] - xpos[:, None], xmat).reshape(3, -1)
    f = 65
    idx = np.minimum(env._step_counter[:, None] + np.arange(f)[None], env._ref_qpos.shape[0] - 1)
    ref = env._ref_qpos[idx]
    disp = np.einsum('nfi,nij->nfj', ref[:, :, :3] - qpos[:, None, :3], xmat)
    rq = mult_quat(np.broadcast_to(reciprocal_quat(qpos[:, 3:7])[:, None], (3, f, 4)), ref[:, :, 3:7])
    o = ts.observation
    assert np.allclose(o['walker/appendages_pos'], app, atol=2e-6)
    assert np.allclose(o['walker/ref_displacement'], disp, atol=2e-6)
    assert np.allclose(o['walker/ref_root_quat'], rq, atol=2e-6)
    assert np.allclose(o['walker/joints_pos'], qpos[:, env._obs_qadr], atol=0)
    assert np.allclose(o['walker/joints_vel'], qvel[:, env._obs_vadr], atol=0)
    assert np.allclose(o['walker/world_zaxis'], xmat[:, 2, :], atol=1e-7)
    sm = raw[:, lay['sensor_mean']]
    assert np.allclose(o['walker/accelerometer'], sm[:, env._sd['accelerometer']], rtol=1e-6)
    assert np.allclose(o['walker/force'], sm[:, env._sd['force']], rtol=1e-6, atol=1e-9)
    assert np.allclose(o['walker/touch'], sm[:, env._sd['touch']], rtol=1e-6, atol=1e-9)


def check_sensor_observables_against_oracle(lib, device_task):
    """The five buffered observables (accelerometer, gyro, velocimeter, force, touch: `buffer_size = n_sub, aggregator = mean`,
    reference fruitfly.py:626-665) returned by env.reset() / env.step() against the fp64 oracle driven with the same controls:
    at FIRST the buffer holds one sample (the forward pass after the reset) padded with zeros (dm_control `Buffer`, SURVEY.md
    App. C), afterwards the mean over the control step's substeps.  Also the instantaneous observables that read the state."""
    from oracle import fly_oracle as fo
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=2, lib_path=lib, device_task=device_task)
    m = env.model
    o = fo.Oracle(m, tolerance=1e-12)
    ts = env.reset()
    q0 = env._sim.get(fly_envs.st.QPOS)[0].astype(np.float64)
    o.reset(q0)
    n_sub = env._n_sub
    names = ('accelerometer', 'gyro', 'velocimeter', 'force', 'touch')
    sd = o.get(fo.SENSORDATA)
    for k in names:
        want = sd[env._sd[k]] / n_sub
        got = np.asarray(ts.observation['walker/' + k], np.float64)[0]
        assert np.allclose(got, want, rtol=2e-4, atol=2e-4 * (np.abs(want).max() + 1e-3)), ('FIRST', k, got, want)
    rs = np.random.RandomState(5)
    sim = env._sim
    events = 0
    for step in range(4):                                   # teacher-forced: each control step starts from the oracle's state
        sim.set(fly_envs.st.QPOS, np.tile(o.qpos, (2, 1))); sim.set(fly_envs.st.QVEL, np.tile(o.qvel, (2, 1)))
        sim.set(fly_envs.st.ACT, np.tile(o.get(fo.ACT), (2, 1))); sim.forward()
        a = rs.uniform(-0.5, 0.5, (2, 59))
        a[1] = a[0]
        ts = env.step(a)
        ctrl = np.zeros(m.nu); ctrl[env._ctrl_of_action] = a[0]
        o.set(fo.CTRL, ctrl)
        mean = o.control_step(n_sub)
        # a contact that opens / closes inside the control step can land on different substeps in fp32 and fp64 (an "event" step,
        # tests/parity_common.py): such a step is bounded loosely, and at most one of the four may be one
        ok = True
        for k in names:
            want = mean[env._sd[k]]
            got = np.asarray(ts.observation['walker/' + k], np.float64)[0]
            ok &= bool(np.allclose(got, want, rtol=2e-3, atol=2e-3 * (np.abs(want).max() + 1e-3)))
            assert np.allclose(got, want, rtol=0.2, atol=0.2 * (np.abs(want).max() + 1e-3)), (step, k, got, want)
        ok &= bool(np.allclose(ts.observation['walker/joints_vel'][0], o.qvel[env._obs_vadr], atol=5e-3))
        events += 0 if ok else 1
        assert np.allclose(ts.observation['walker/joints_pos'][0], o.qpos[env._obs_qadr], atol=2e-4), step
        assert np.allclose(ts.observation['walker/joints_vel'][0], o.qvel[env._obs_vadr], atol=1.0), step
        assert np.allclose(ts.observation['walker/actuator_activation'][0], o.get(fo.ACT), atol=2e-5), step
        assert np.array_equal(ts.observation['walker/force'][0], ts.observation['walker/force'][1])
    assert events <= 1, events
    env.close()


@pytest.mark.parametrize('device_task', [False, True])
def test_sensor_observables_match_the_oracle_mean_incl_first_step(emu, device_task):
    check_sensor_observables_against_oracle(emu, device_task)


# ------------------------------------------------------------------------------------ flight_imitation
FLIGHT_OBS = ['walker/accelerometer', 'walker/actuator_activation', 'walker/gyro', 'walker/joints_pos', 'walker/joints_vel',
              'walker/velocimeter', 'walker/world_zaxis', 'walker/ref_displacement', 'walker/ref_root_quat']
FLIGHT_ACTIONS = ['head_abduct', 'head_twist', 'head', 'wing_yaw_left', 'wing_roll_left', 'wing_pitch_left',
                  'wing_yaw_right', 'wing_roll_right', 'wing_pitch_right', 'abdomen_abduct', 'abdomen', 'user_0']


def test_flight_env_contract(emu):
    """flight_imitation(): 12 actions (docs/sensory-input-tracking.ipynb:183 order), obs shapes `:152-160`
    (3, 0, 3, 25, 25, 3, 3, 6x3, 6x4), control / physics timesteps of tasks/constants.py:16-17."""
    env = fly_envs.flight_imitation(lib_path=emu)
    spec = env.observation_spec()
    assert list(spec) == FLIGHT_OBS
    assert [spec[k].shape for k in FLIGHT_OBS] == [(3,), (0,), (3,), (25,), (25,), (3,), (3,), (6, 3), (6, 4)]
    a = env.action_spec()
    assert a.shape == (12,) and a.name.split('\t') == FLIGHT_ACTIONS
    assert a.minimum[-1] == -1 and a.maximum[-1] == 1
    assert np.isclose(env.control_timestep(), 2e-4) and np.isclose(env.physics.timestep(), 5e-5)
    ts = env.reset()
    assert ts.step_type == StepType.FIRST
    # the fly starts on the reference: zero displacement, identity relative orientation
    assert np.abs(ts.observation['walker/ref_displacement'][0]).max() < 1e-5
    assert np.allclose(ts.observation['walker/ref_root_quat'][0], [1, 0, 0, 0], atol=1e-5)


def test_flight_env_steps_with_wbpg(emu):
    """random policy of tasks/task_utils.py:58-65 (U(-0.2, 0.2)); reward is the product of the CoM and orientation
    tracking factors (flight_imitation.py:170-201), 1 at reset and decaying smoothly."""
    env = fly_envs.flight_imitation(n_envs=2, lib_path=emu, seed=3)
    env.reset()
    rs = np.random.RandomState(0)
    rew = []
    for k in range(40):
        ts = env.step(rs.uniform(-0.2, 0.2, (2, 12)))
        assert all(np.all(np.isfinite(v)) for v in ts.observation.values())
        assert np.all(ts.step_type == StepType.MID)
        rew.append(ts.reward.copy())
    rew = np.array(rew)
    assert np.all((rew > 0.8) & (rew <= 1.0)), rew[-1]
    # wings are driven along the beat pattern: stroke angle spans most of the synthetic cycle within one beat (~23 steps)
    assert env._wing_qpos_host.shape == (2, 6)
    q = env._sim.get(fly_envs.st.QPOS)
    assert np.abs(np.linalg.norm(q[:, 3:7], axis=1) - 1).max() < 1e-5


def test_flight_env_good_termination_at_trajectory_end(emu):
    """episode ends with discount 1 when the reference runs out (flight_imitation.py:203-226): 200-step default
    trajectory - (future_steps + 1) = step 194."""
    env = fly_envs.flight_imitation(terminal_com_dist=float('inf'), lib_path=emu)
    env.reset()
    n = 0
    while True:
        ts = env.step(np.zeros(12))
        n += 1
        if ts.step_type == StepType.LAST:
            break
        assert n < 400
    # either the end of the reference (discount 1) or, for a passive fly that sinks, the height limit (discount 0)
    assert (n == 194 and ts.discount == 1.0) or (n < 194 and ts.discount == 0.0)


def test_action_map_on_device_matches_host_permutation(emu):
    """fb_set_action_map (FruitFly.apply_action on the device): actions in action order + NaN -> 0 give the same ctrl as
    the host-side permutation (fruitfly.py:532-544, tasks/base.py:197-201)."""
    from flybody_b200 import stepper as st
    from flybody_b200.flymodel import load_model
    m = load_model('walk')
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=3, lib_path=emu)
    env.reset()
    rs = np.random.RandomState(1)
    a = rs.uniform(-0.5, 0.5, (3, 59))
    a[1, 7] = np.nan
    env.step(a)
    ctrl = env._sim.get(st.CTRL)
    want = np.zeros((3, m.nu), np.float32)
    want[:, env._ctrl_of_action] = np.nan_to_num(a, nan=0.0)
    rng = m.actuator_ctrlrange
    assert np.allclose(ctrl, want, atol=1e-7)
    assert want[1, env._ctrl_of_action[7]] == 0.0
