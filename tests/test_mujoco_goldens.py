"""MuJoCo pin kit, consumer side.  `tools/dump_mujoco_goldens.py` (run wherever the reference + mujoco exist) writes
`tests/golden/mujoco_{walk,flight,vision}.npz`; these tests then pin, against a REAL MuJoCo,

  (a) the model compiler                       test_compiler_matches_mjmodel           arrays by NAME, 1e-9 relative
  (b) the fp64 oracle, stage by stage          test_oracle_stage_fields                smooth 1e-8, constraint 1e-5 relative
  (c) the fp64 oracle over the trajectory      test_oracle_teacher_forced_trajectory   |dqpos| 1e-8, |dqvel| 1e-5, sensor mean 1e-5 rel
  (d) the kernel source / CUDA stepper         test_stepper_* (emulation here, -m gpu on the B200)   fp32 gates of parity_common

They SKIP while the files are absent (mujoco / dm_control are not installable in the build container or on the GPU boxes: no wheel,
no network), and DESIGN.md says "parity unpinned" until they are there.  `test_consumer_plumbing_*` always runs: it writes a file of
the same layout FROM THE ORACLE (meta source = 'oracle-selftest', names and index maps deliberately permuted like MuJoCo's differ
from ours) and pushes it through every checker, so the consumer code itself is exercised on every run.

Index maps: MuJoCo's model of the task has the 67-body ghost fly where ours has one fused free body, and dm_control names the
floor geom differently; everything is therefore matched BY NAME over the `walker/` objects (bodies, joints -> dofs / qpos addresses,
collision geoms, sites, sensors, actuators)."""
import json
import os

import numpy as np
import pytest

import __graft_entry__ as ge
from flybody_b200 import stepper as st
from flybody_b200.flymodel import load_model
from oracle import fly_oracle as fo
from parity_common import rel_err, sensor_mean_error

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = dict(model=1e-9, smooth=1e-8, constraint=1e-5, tf_q=1e-8, tf_v=1e-5, tf_s=1e-5,            # fp64 oracle vs MuJoCo
           dev_smooth=2e-6, dev_constraint=1e-3, dev_q=2e-6, dev_v=5e-4, dev_s=2e-4)              # fp32 stepper vs MuJoCo


class Golden:
    def __init__(self, path):
        z = np.load(path)
        self.a = {k: z[k] for k in z.files if k != '__meta__'}
        self.meta = json.loads(bytes(z['__meta__']).decode())
        self.names = self.meta['names']

    def __getitem__(self, k):
        return self.a[k]

    def __contains__(self, k):
        return k in self.a

    @property
    def n_steps(self):
        return self.a['traj/qpos'].shape[0] - 1


def golden(variant):
    p = os.path.join(GOLDEN_DIR, f'mujoco_{variant}.npz')
    if not os.path.exists(p):
        pytest.skip(f'{os.path.relpath(p)} absent: run tools/dump_mujoco_goldens.py where mujoco + dm_control + flybody are installed')
    return Golden(p)


class Maps:
    """name-matched index maps between our compiled model `m` and the recorded mjModel"""

    def __init__(self, m, g):
        self.m, self.g = m, g
        ours, theirs = m.meta, g.names

        def match(our_names, their_names, what, keep=lambda s: s.startswith('walker/')):
            idx = {s: i for i, s in enumerate(their_names)}
            pairs = [(i, idx[s]) for i, s in enumerate(our_names) if keep(s) and s in idx]
            missing = [s for s in our_names if keep(s) and s not in idx]
            assert not missing, f'{what}: {len(missing)} of our names are not in the recorded model, e.g. {missing[:4]}'
            return np.array([p[0] for p in pairs], np.int64), np.array([p[1] for p in pairs], np.int64)
        self.body = match(ours['body_names'], theirs['body'], 'bodies')
        self.site = match(ours['site_names'], theirs['site'], 'sites')
        self.sensor = match(ours['sensor_names'], theirs['sensor'], 'sensors')
        self.act = match(ours['actuator_names'], theirs['actuator'], 'actuators')
        self.geom = match(ours['geom_names'], theirs['geom'], 'geoms')
        # joints: the root free joint is named after the attachment frame by dm_control ('walker/'); hinges by their XML name
        self.jnt = match(ours['jnt_names'], theirs['jnt'], 'joints')
        jo, jt = self.jnt
        t_qadr, t_dadr, t_type = g['model/jnt_qposadr'], g['model/jnt_dofadr'], g['model/jnt_type']
        q_o, q_t, d_o, d_t = [], [], [], []
        for a, b in zip(jo, jt):
            assert int(m.jnt_type[a]) == int(t_type[b]), (ours['jnt_names'][a], m.jnt_type[a], t_type[b])
            nq, nd = (7, 6) if int(m.jnt_type[a]) == 0 else (1, 1)
            q_o += list(range(int(m.jnt_qposadr[a]), int(m.jnt_qposadr[a]) + nq)); q_t += list(range(int(t_qadr[b]), int(t_qadr[b]) + nq))
            d_o += list(range(int(m.jnt_dofadr[a]), int(m.jnt_dofadr[a]) + nd)); d_t += list(range(int(t_dadr[b]), int(t_dadr[b]) + nd))
        self.q = (np.array(q_o), np.array(q_t)); self.d = (np.array(d_o), np.array(d_t))
        so, stt = self.sensor
        s_o, s_t = [], []
        for a, b in zip(so, stt):
            n = int(m.sensor_dim[a]); assert n == int(g['model/sensor_dim'][b])
            s_o += list(range(int(m.sensor_adr[a]), int(m.sensor_adr[a]) + n)); s_t += list(range(int(g['model/sensor_adr'][b]), int(g['model/sensor_adr'][b]) + n))
        self.sd = (np.array(s_o), np.array(s_t))
        # activations follow the actuators that have one
        ao, at = self.act
        a_o = [int(m.actuator_actadr[a]) for a in ao if int(m.actuator_actadr[a]) >= 0]
        a_t = [int(g['model/actuator_actadr'][b]) for a, b in zip(ao, at) if int(m.actuator_actadr[a]) >= 0]
        self.a = (np.array(a_o, np.int64), np.array(a_t, np.int64))
        # the floor: our 'floor' <-> the recorded model's plane geom
        planes = [i for i, t in enumerate(g['model/geom_type']) if int(t) == 0]
        self.floor = (ours['geom_names'].index('floor') if 'floor' in ours['geom_names'] else -1, planes[0] if planes else -1)
        # the terrain of vision_guided_flight: our heightfield geom <-> the recorded model's hfield geom (mjGEOM_HFIELD = 1)
        hfs = [i for i, t in enumerate(g['model/geom_type']) if int(t) == 1]
        self.terrain = (int(ours['hf_geom']) if 'hf_geom' in ours else -1, hfs[0] if hfs else -1)

    def to_ours(self, pair, their_vec, n_ours, fill=None):
        out = np.zeros(n_ours) if fill is None else np.array(fill, np.float64).copy()
        out[pair[0]] = np.asarray(their_vec, np.float64)[pair[1]]
        return out


# ------------------------------------------------------------------------------------------------ (a) compiler
def check_compiler(m, g):
    mp = Maps(m, g)
    assert int(g['model/nu']) == m.nu and int(g['model/na']) == m.na
    assert len(mp.d[0]) == sum(1 for s in m.meta['jnt_names'] if s.startswith('walker/')) + 5       # every walker dof found (free joint: 6)
    cmp = lambda ours, theirs, what, tol=TOL['model']: (_ for _ in ()).throw(AssertionError((what, rel_err(theirs, ours)))) if rel_err(theirs, ours) > tol else None
    bo, bt = mp.body
    for f in ('body_pos', 'body_quat', 'body_ipos', 'body_mass', 'body_inertia'):
        k = np.asarray(getattr(m, f)).reshape(m.nbody, -1).shape[1]
        cmp(np.asarray(getattr(m, f)).reshape(m.nbody, k)[bo], g['model/' + f].reshape(-1, k)[bt], f)
    # inertial frames are defined up to the order / sign of the principal axes: compare the inertia tensors in the body frame
    def tensor(quat, diag):
        w, x, y, z = quat
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        return R @ np.diag(diag) @ R.T
    for a, b in zip(bo, bt):
        To, Tt = tensor(m.body_iquat.reshape(-1, 4)[a], m.body_inertia.reshape(-1, 3)[a]), tensor(g['model/body_iquat'][b], g['model/body_inertia'][b])
        assert np.abs(To - Tt).max() <= 1e-8 * (np.abs(Tt).max() + 1e-30), ('inertia tensor', m.meta['body_names'][a])
    jo, jt = mp.jnt
    for f, k in (('jnt_pos', 3), ('jnt_axis', 3), ('jnt_stiffness', 1), ('jnt_range', 2), ('jnt_solref', 2), ('jnt_solimp', 5), ('jnt_margin', 1)):
        cmp(np.asarray(getattr(m, f)).reshape(-1, k)[jo], g['model/' + f].reshape(-1, k)[jt], f)
    assert np.array_equal(np.asarray(m.jnt_limited)[jo] != 0, g['model/jnt_limited'][jt] != 0)
    do, dt = mp.d
    for f in ('dof_armature', 'dof_damping', 'dof_invweight0'):
        cmp(np.asarray(getattr(m, f))[do], g['model/' + f][dt], f, 1e-7 if f == 'dof_invweight0' else TOL['model'])
    cmp(np.asarray(m.qpos0)[mp.q[0]], g['model/qpos0'][mp.q[1]], 'qpos0'); cmp(np.asarray(m.qpos_spring)[mp.q[0]], g['model/qpos_spring'][mp.q[1]], 'qpos_spring')
    go, gt = mp.geom
    for f, k in (('geom_size', 3), ('geom_pos', 3), ('geom_quat', 4), ('geom_rbound', 1), ('geom_friction', 3), ('geom_solmix', 1), ('geom_solref', 2),
                 ('geom_solimp', 5), ('geom_margin', 1), ('geom_gap', 1)):
        cmp(np.asarray(getattr(m, f)).reshape(-1, k)[go], g['model/' + f].reshape(-1, k)[gt], f)
    for f in ('geom_type', 'geom_condim'):
        assert np.array_equal(np.asarray(getattr(m, f))[go], g['model/' + f][gt]), f
    ao, at = mp.act
    for f, k in (('actuator_gainprm', 3), ('actuator_biasprm', 3), ('actuator_dynprm', 3), ('actuator_ctrlrange', 2), ('actuator_forcerange', 2)):
        ours = np.asarray(getattr(m, f)).reshape(m.nu, -1)[ao][:, :k]
        cmp(ours, g['model/' + f].reshape(int(g['model/nu']), -1)[at][:, :ours.shape[1]], f)
    for f in ('actuator_dyntype', 'actuator_biastype', 'actuator_trntype'):
        assert np.array_equal(np.asarray(getattr(m, f))[ao], g['model/' + f][at]), f
    assert np.isclose(m.opt_timestep, float(g['model/opt_timestep'])) and np.isclose(m.opt_impratio, float(g['model/opt_impratio']))
    assert np.isclose(m.opt_density, float(g['model/opt_density'])) and np.isclose(m.opt_viscosity, float(g['model/opt_viscosity']))
    assert int(m.opt_noslip_iterations) == int(g['model/opt_noslip_iterations']) and int(m.opt_cone_elliptic) == int(g['model/opt_cone'])
    assert np.isclose(m.stat_meaninertia, float(g['model/stat_meaninertia']), rtol=1e-6)
    return mp


# ------------------------------------------------------------------------------------------------ backends
class OracleBackend:
    kind = 'oracle'

    def __init__(self, m):
        self.m, self.o = m, fo.Oracle(m, tolerance=1e-12)

    def load_state(self, qpos, qvel, act, ctrl, warm):
        self.o.reset(qpos, qvel)
        if self.m.na:
            self.o.set(fo.ACT, act)
        self.o.set(fo.CTRL, ctrl); self.o.set(fo.QACC_WARMSTART, warm)
        self.o.forward()

    def set_terrain(self, heights):
        self.o.set_hfield(self.m.meta['hf_geom'], self.m.hf_size, heights, self.m.hf_pair_geom)

    def get(self, f):
        return self.o.get(f)

    def control_step(self, n_sub):
        mean = self.o.control_step(n_sub)
        return self.o.qpos, self.o.qvel, mean


class StepperBackend:
    kind = 'stepper'

    def __init__(self, m, lib):
        self.m, self.s = m, st.BatchedStepper(m, 2, lib_path=lib)

    def load_state(self, qpos, qvel, act, ctrl, warm):
        self.s.reset(qpos, qvel)
        if self.m.na:
            self.s.set(st.ACT, act)
        self.s.set_control(ctrl); self.s.forward()

    def set_terrain(self, heights):
        h = np.asarray(heights, np.float32)
        self.s.hfield_collision(self.m.meta['hf_geom'], self.m.hf_size, h.shape[0], h.shape[1], self.m.hf_pair_geom)
        self.s.hfield_write(np.arange(2), np.stack([h, h]))

    def get(self, f):
        return self.s.get(f)[1].astype(np.float64)

    def control_step(self, n_sub):
        self.s.step(n_sub)
        return self.get(st.QPOS), self.get(st.QVEL), self.get(st.SENSOR_MEAN)


def dev_con_tol(variant):
    """constraint-stage gate of the fp32 stepper: terrain contacts come from MPR against prisms of the heightfield (fp32 portal
    refinement), their forces agree with the fp64 oracle to 5e-3 relative where primitive contacts reach 1e-3"""
    return TOL['dev_constraint'] * (5 if variant == 'vision' else 1)


def with_terrain(be, g):
    """vision_guided_flight files carry the episode's heightfield (hfield/data: mjModel.hfield_data, normalised; hfield/size: radius x,
    radius y, elevation z, base z): world heights = data * elevation"""
    if 'hfield/data' in g:
        be.set_terrain(np.asarray(g['hfield/data'], np.float64) * float(g['hfield/size'][2]))
    return be


def our_state(m, mp, g, prefix, row=None):
    """(qpos, qvel, act, ctrl, warm) in OUR index space from a recorded state; objects MuJoCo does not share with us (ghost) keep
    their model defaults"""
    get = (lambda f: g[prefix + f][row]) if row is not None else (lambda f: g[prefix + f])
    qpos = mp.to_ours(mp.q, get('qpos'), m.nq, fill=m.qpos0)
    qvel = mp.to_ours(mp.d, get('qvel'), m.nv)
    act = mp.to_ours(mp.a, get('act'), m.na) if m.na else np.zeros(0)
    ctrl = mp.to_ours(mp.act, get('ctrl'), m.nu)
    warm = mp.to_ours(mp.d, get('qacc_warmstart'), m.nv)
    return qpos, qvel, act, ctrl, warm


# ------------------------------------------------------------------------------------------------ (b) stages
def check_stage_fields(m, g, mp, be, tol_smooth, tol_con):
    n_stage = len(g['stage_steps'])
    assert n_stage > 0
    worst = {}
    for k in range(n_stage):
        P = f'stage/{k}/'
        be.load_state(*our_state(m, mp, g, P))
        bo, bt = mp.body
        do, dt = mp.d

        def chk(name, ours, theirs, tol):
            e = rel_err(theirs, ours)
            worst[name] = max(worst.get(name, 0.0), e)
            assert e < tol, (k, name, e)
        chk('xpos', be.get(fo.XPOS).reshape(-1, 3)[bo], g[P + 'xpos'][bt], tol_smooth)
        chk('xmat', be.get(fo.XMAT).reshape(-1, 9)[bo], g[P + 'xmat'].reshape(-1, 9)[bt], tol_smooth)
        chk('site_xpos', be.get(fo.SITE_XPOS).reshape(-1, 3)[mp.site[0]], g[P + 'site_xpos'][mp.site[1]], tol_smooth)
        M = be.get(fo.QM_DENSE).reshape(m.nv, m.nv)
        chk('qM', M[np.ix_(do, do)], g[P + 'qM_dense'][np.ix_(dt, dt)], tol_smooth)
        for name, f in (('qfrc_bias', fo.QFRC_BIAS), ('qfrc_passive', fo.QFRC_PASSIVE), ('qfrc_actuator', fo.QFRC_ACTUATOR), ('qfrc_smooth', fo.QFRC_SMOOTH)):
            if np.abs(g[P + name][dt]).max() > 0:
                chk(name, be.get(f)[do], g[P + name][dt], tol_smooth)
        # contacts between walker geoms / the floor: same set of geom pairs, same distances
        gmap = {int(b): int(a) for a, b in zip(*mp.geom)}
        if mp.floor[0] >= 0:
            gmap[mp.floor[1]] = mp.floor[0]
        if mp.terrain[0] >= 0 and mp.terrain[1] >= 0:
            gmap[mp.terrain[1]] = mp.terrain[0]
        theirs = sorted((gmap[int(a)], gmap[int(b)], float(dd)) for a, b, dd in zip(g[P + 'con_geom1'], g[P + 'con_geom2'], g[P + 'con_dist'])
                        if int(a) in gmap and int(b) in gmap)
        ncon = int(be.get(fo.NCON)[0])
        con = be.get(fo.CONTACT).reshape(-1, 16)[:ncon]
        ours = sorted((int(c[7]), int(c[8]), float(c[0])) for c in con)
        assert [t[:2] for t in theirs] == [o[:2] for o in ours], (k, 'contact pairs', theirs, ours)
        if ours:
            dist_err = max(abs(a[2] - b[2]) for a, b in zip(theirs, ours))
            worst['con_dist'] = max(worst.get('con_dist', 0.0), dist_err)
            assert dist_err < (1e-7 if be.kind == 'oracle' else 5e-5), (k, 'contact distance', dist_err)
        for name, f in (('qfrc_constraint', fo.QFRC_CONSTRAINT), ('qacc', fo.QACC)):
            if np.abs(g[P + name][dt]).max() > 0:
                chk(name, be.get(f)[do], g[P + name][dt], tol_con)
        chk('sensordata', be.get(fo.SENSORDATA)[mp.sd[0]], g[P + 'sensordata'][mp.sd[1]], tol_con)
    return worst


# ------------------------------------------------------------------------------------------------ (c) trajectory
def check_teacher_forced(m, g, mp, be, tol_q, tol_v, tol_s, max_events=0.05, n_steps=None):
    n_sub = int(g.meta['n_sub'])
    T = g.n_steps if n_steps is None else min(n_steps, g.n_steps)
    eq, ev, es = [], [], []
    per_sub = 'sub/sensordata' in g
    for k in range(T):
        qpos, qvel, act, _, warm = our_state(m, mp, g, 'traj/', k)
        ctrl = mp.to_ours(mp.act, g['traj/ctrl'][k + 1], m.nu)          # the controls the task's before_step wrote for this step
        be.load_state(qpos, qvel, act, ctrl, warm)
        q1, v1, mean = be.control_step(n_sub)
        eq.append(np.abs(q1[mp.q[0]] - g['traj/qpos'][k + 1][mp.q[1]]).max())
        ev.append(np.abs(v1[mp.d[0]] - g['traj/qvel'][k + 1][mp.d[1]]).max())
        if per_sub and k < int(g.meta['n_substep_steps']):               # the samples the observation buffers average
            want = g['sub/sensordata'][k * n_sub:(k + 1) * n_sub].mean(0)
            sub = Maps.to_ours(mp, mp.sd, want, m.nsensordata)
            es.append(sensor_mean_error(m, mean, sub))
    eq, ev, es = np.array(eq), np.array(ev), np.array(es)
    res = dict(p90_q=float(np.percentile(eq, 90)), p90_v=float(np.percentile(ev, 90)), max_q=float(eq.max()), max_v=float(ev.max()),
               events=int(((eq > tol_q) | (ev > tol_v)).sum()), steps=T, p90_s=float(np.percentile(es, 90)) if len(es) else None)
    assert res['p90_q'] < tol_q and res['p90_v'] < tol_v, res
    assert res['events'] <= max_events * T, res
    if len(es):
        assert res['p90_s'] < tol_s, res
    return res


def check_observations(m, g, mp):
    """the buffered observables of the recorded TimeSteps are the per-substep means (and, at FIRST, one sample padded with zeros)"""
    n_sub = int(g.meta['n_sub'])
    nss = int(g.meta['n_substep_steps'])
    if 'sub/sensordata' not in g or nss == 0:
        pytest.skip('no per-substep rows in this file')
    names = g.names['sensor']
    for key, sname in (('obs/walker/accelerometer', 'walker/accelerometer'), ('obs/walker/gyro', 'walker/gyro'), ('obs/walker/velocimeter', 'walker/velocimeter')):
        if key not in g or sname not in names:
            continue
        s = names.index(sname); a = int(g['model/sensor_adr'][s])
        first = g['traj/sensordata'][0][a:a + 3] / n_sub
        assert np.allclose(g[key][0], first, rtol=1e-9, atol=1e-12), ('FIRST padding rule (SURVEY App. C) does not hold for', key, g[key][0], first)
        for k in range(nss):
            want = g['sub/sensordata'][k * n_sub:(k + 1) * n_sub, a:a + 3].mean(0)
            assert np.allclose(g[key][k + 1], want, rtol=1e-9, atol=1e-12), (key, k)


# ------------------------------------------------------------------------------------------------ tests on the real files
@pytest.mark.parametrize('variant', ['walk', 'flight', 'vision'])
def test_compiler_matches_mjmodel(variant):
    g = golden(variant)
    check_compiler(load_model(variant), g)


@pytest.mark.parametrize('variant', ['walk', 'flight', 'vision'])
def test_recorded_observations_follow_the_buffer_rules(variant):
    g = golden(variant)
    m = load_model(variant)
    check_observations(m, g, Maps(m, g))


@pytest.mark.parametrize('variant', ['walk', 'flight', 'vision'])
def test_oracle_stage_fields(variant):
    g = golden(variant); m = load_model(variant); mp = Maps(m, g)
    print(check_stage_fields(m, g, mp, with_terrain(OracleBackend(m), g), TOL['smooth'], TOL['constraint']))


@pytest.mark.parametrize('variant', ['walk', 'flight', 'vision'])
def test_oracle_teacher_forced_trajectory(variant):
    g = golden(variant); m = load_model(variant); mp = Maps(m, g)
    print(check_teacher_forced(m, g, mp, with_terrain(OracleBackend(m), g), TOL['tf_q'], TOL['tf_v'], TOL['tf_s']))


@pytest.mark.parametrize('variant', ['walk', 'flight', 'vision'])
def test_stepper_kernel_source_against_mujoco(variant):
    """the kernel source (host emulation) against MuJoCo: 20 control steps + every recorded stage"""
    g = golden(variant); m = load_model(variant); mp = Maps(m, g)
    ge.build()
    be = with_terrain(StepperBackend(m, ge.EMU), g)
    print(check_stage_fields(m, g, mp, be, TOL['dev_smooth'], dev_con_tol(variant)))
    print(check_teacher_forced(m, g, mp, be, TOL['dev_q'], 4 * TOL['dev_v'], 5 * TOL['dev_s'], max_events=0.15, n_steps=20))


@pytest.mark.gpu
@pytest.mark.parametrize('variant', ['walk', 'flight', 'vision'])
def test_cuda_stepper_against_mujoco(variant):
    """BASELINE.json config 1 on the B200 against the recorded CPU MuJoCo trajectory (north_star: "outputs match the reference CPU
    MuJoCo qpos/qvel trajectories on identical action sequences within a stated fp32 tolerance")."""
    g = golden(variant); m = load_model(variant); mp = Maps(m, g)
    be = with_terrain(StepperBackend(m, None), g)
    print(check_stage_fields(m, g, mp, be, TOL['dev_smooth'], dev_con_tol(variant)))
    print(check_teacher_forced(m, g, mp, be, TOL['dev_q'], TOL['dev_v'], TOL['dev_s'], max_events=0.10))


# ------------------------------------------------------------------------------------------------ self-test of the consumer
def write_selftest_file(path, variant, n_steps=6, n_stage=3, seed=0):
    """A file of the dump tool's layout produced by THE ORACLE, with the object orders permuted and a multi-body ghost's worth of
    extra (unnamed-by-us) objects appended, the way the real mjModel differs from ours.  Not a parity statement -- plumbing only."""
    m = load_model(variant)
    rs = np.random.RandomState(seed)
    o = fo.Oracle(m, tolerance=1e-12)
    meta_names = m.meta
    walker = lambda names: [s for s in names if s.startswith('walker/')]
    # "their" model: walker objects in a rotated order + some foreign objects in front; the joint / dof / qpos spaces keep our order
    # within the walker but are shifted behind a foreign free joint (7 qpos / 6 dofs)
    def perm(names, n_foreign, tag):
        ours = list(range(len(names)))
        keep = [i for i in ours if names[i].startswith('walker/') or names[i] in ('floor', 'terrain')]
        order = keep[1:] + keep[:1]                                       # rotate
        their_names = [f'{tag}/foreign{i}' for i in range(n_foreign)] + [('groundplane' if names[i] == 'floor' else names[i]) for i in order]
        return order, their_names, n_foreign
    out = {}
    names = {}
    b_order, names['body'], b_sh = perm(meta_names['body_names'], 2, 'ghostb')
    g_order, names['geom'], g_sh = perm(meta_names['geom_names'], 3, 'ghostg')
    s_order, names['site'], s_sh = perm(meta_names['site_names'], 1, 'ghosts')
    n_order, names['sensor'], n_sh = perm(meta_names['sensor_names'], 0, 'x')
    a_order, names['actuator'], a_sh = perm(meta_names['actuator_names'], 0, 'x')
    names['tendon'] = list(meta_names.get('tendon_names', []))
    wj = [i for i, s in enumerate(meta_names['jnt_names']) if s.startswith('walker/')]
    names['jnt'] = ['foreign/'] + [meta_names['jnt_names'][i] for i in wj]
    QS, DS = 7, 6                                                          # shift of the walker's qpos / dof addresses
    nq_t, nv_t = QS + sum(7 if m.jnt_type[i] == 0 else 1 for i in wj), DS + sum(6 if m.jnt_type[i] == 0 else 1 for i in wj)
    t_qadr = [0] + [int(m.jnt_qposadr[i]) + QS for i in wj]; t_dadr = [0] + [int(m.jnt_dofadr[i]) + DS for i in wj]
    assert all(int(m.jnt_qposadr[i]) < nq_t - QS for i in wj), 'walker joints must come first in our model'
    nb_t, ng_t, ns_t = len(names['body']), len(names['geom']), len(names['site'])

    def rows(field, k, order, shift, n_t, fill=0.0):
        src = np.asarray(getattr(m, field), np.float64).reshape(-1, k)
        dst = np.full((n_t, k), fill)
        dst[shift:shift + len(order)] = src[order]
        return dst.squeeze(-1) if k == 1 else dst
    for f, k in (('body_pos', 3), ('body_quat', 4), ('body_ipos', 3), ('body_iquat', 4), ('body_mass', 1), ('body_inertia', 3)):
        out['model/' + f] = rows(f, k, b_order, b_sh, nb_t, 1.0 if f.endswith('quat') else 0.0)
    for f, k in (('geom_size', 3), ('geom_pos', 3), ('geom_quat', 4), ('geom_rbound', 1), ('geom_friction', 3), ('geom_solmix', 1), ('geom_solref', 2),
                 ('geom_solimp', 5), ('geom_margin', 1), ('geom_gap', 1)):
        out['model/' + f] = rows(f, k, g_order, g_sh, ng_t)
    for f in ('geom_type', 'geom_condim'):
        out['model/' + f] = rows(f, 1, g_order, g_sh, ng_t, 5).astype(np.int32)
    jsel = np.array(wj)
    for f, k in (('jnt_pos', 3), ('jnt_axis', 3), ('jnt_stiffness', 1), ('jnt_range', 2), ('jnt_solref', 2), ('jnt_solimp', 5), ('jnt_margin', 1)):
        src = np.asarray(getattr(m, f), np.float64).reshape(-1, k)[jsel]
        out['model/' + f] = np.concatenate([np.zeros((1, k)), src]).squeeze(-1) if k == 1 else np.concatenate([np.zeros((1, k)), src])
    out['model/jnt_limited'] = np.concatenate([[0], np.asarray(m.jnt_limited)[jsel]]).astype(np.int32)
    out['model/jnt_type'] = np.concatenate([[0], np.asarray(m.jnt_type)[jsel]]).astype(np.int32)
    out['model/jnt_qposadr'] = np.array(t_qadr, np.int32); out['model/jnt_dofadr'] = np.array(t_dadr, np.int32)
    nq_w, nv_w = nq_t - QS, nv_t - DS

    def qspace(v, n_w, shift, n_t, fill=0.0):
        o_ = np.full(n_t, fill); o_[shift:shift + n_w] = np.asarray(v, np.float64)[:n_w]; return o_
    for f in ('dof_armature', 'dof_damping', 'dof_invweight0'):
        out['model/' + f] = qspace(getattr(m, f), nv_w, DS, nv_t)
    out['model/qpos0'] = qspace(m.qpos0, nq_w, QS, nq_t); out['model/qpos_spring'] = qspace(m.qpos_spring, nq_w, QS, nq_t)
    for f, k in (('actuator_gainprm', 3), ('actuator_biasprm', 3), ('actuator_dynprm', 3), ('actuator_ctrlrange', 2), ('actuator_forcerange', 2)):
        src = np.asarray(getattr(m, f), np.float64).reshape(m.nu, -1)[a_order]
        out['model/' + f] = np.concatenate([src, np.zeros((m.nu, 10 - src.shape[1]))], 1) if k == 3 else src      # mjNGAIN = 10 columns
    for f in ('actuator_dyntype', 'actuator_biastype', 'actuator_trntype'):
        out['model/' + f] = np.asarray(getattr(m, f))[a_order].astype(np.int32)
    act_t = np.full(m.nu, -1, np.int32); c = 0
    for i_t, i_o in enumerate(a_order):
        if int(m.actuator_actadr[i_o]) >= 0:
            act_t[i_t] = c; c += 1
    out['model/actuator_actadr'] = act_t
    sadr, c = [], 0
    for i_o in n_order:
        sadr.append(c); c += int(m.sensor_dim[i_o])
    out['model/sensor_adr'] = np.array(sadr, np.int32); out['model/sensor_dim'] = np.asarray(m.sensor_dim)[n_order].astype(np.int32)
    for f, v in (('nu', m.nu), ('na', m.na), ('opt_timestep', m.opt_timestep), ('opt_impratio', m.opt_impratio), ('opt_density', m.opt_density),
                 ('opt_viscosity', m.opt_viscosity), ('opt_noslip_iterations', m.opt_noslip_iterations), ('opt_cone', m.opt_cone_elliptic),
                 ('stat_meaninertia', m.stat_meaninertia)):
        out['model/' + f] = np.array(v)
    # ---- states in "their" index space
    inv = lambda order: {o_: i for i, o_ in enumerate(order)}
    sens_t = lambda sd: np.concatenate([sd[int(m.sensor_adr[i]):int(m.sensor_adr[i]) + int(m.sensor_dim[i])] for i in n_order]) if m.nsensor else np.zeros(0)
    act_of = [int(m.actuator_actadr[i]) for i in a_order if int(m.actuator_actadr[i]) >= 0]

    def their_state():
        return dict(qpos=qspace(o.qpos, nq_w, QS, nq_t), qvel=qspace(o.qvel, nv_w, DS, nv_t), act=o.get(fo.ACT)[act_of] if m.na else np.zeros(0),
                    ctrl=o.get(fo.CTRL)[a_order], qacc=qspace(o.get(fo.QACC), nv_w, DS, nv_t),
                    qacc_warmstart=qspace(o.get(fo.QACC_WARMSTART), nv_w, DS, nv_t), sensordata=sens_t(o.get(fo.SENSORDATA)))
    n_sub = 10 if variant == 'walk' else 4
    q0 = m.qpos0.copy()
    if variant == 'vision':          # an episode's terrain, the fly low enough over it that its body geoms meet the heightfield
        from flybody_b200 import arenas
        nrow, ncol = int(m.meta['hf_nrow']), int(m.meta['hf_ncol'])
        terr = arenas.SineBumps(dim=int(m.hf_size[0]), grid_density=(nrow - 1) // (2 * int(m.hf_size[0]))).generate(rs)
        assert terr.shape == (nrow, ncol), terr.shape
        o.set_hfield(m.meta['hf_geom'], m.hf_size, terr, m.hf_pair_geom)
        out['hfield/data'] = np.asarray(terr, np.float64) / float(m.hf_size[2]); out['hfield/size'] = np.asarray(m.hf_size, np.float64)
        x, y = -5.0, 0.0
        q0[0], q0[1], q0[2] = x, y, float(arenas.hfield_height(terr[None], [x], [y], float(m.hf_size[0]))[0]) + 0.1
    if variant == 'walk':
        for side in ('left', 'right'):
            for dof, val in (('yaw', 1.5), ('roll', 0.7), ('pitch', -1.0)):
                q0[m.jnt_qposadr_of(f'walker/wing_{dof}_{side}')] = val
    o.reset(q0)
    traj, sub, stages, stage_steps = [their_state()], [], {}, []
    scale = 0.5 if variant == 'walk' else 0.2
    seen_terrain_contact = False
    for k in range(n_steps):
        ctrl = rs.uniform(-scale, scale, m.nu)
        if k % max(1, n_steps // n_stage) == 0 and len(stage_steps) < n_stage:
            i = len(stage_steps); stage_steps.append(k); P = f'stage/{i}/'
            o.set(fo.CTRL, ctrl); o.forward()
            for kk, v in their_state().items():
                stages[P + kk] = v
            X = lambda f, kdim, order, shift, n_t: (lambda a: (lambda d_: (d_.__setitem__(slice(shift, shift + len(order)), a[order]), d_)[1])(np.zeros((n_t, kdim))))(o.get(f).reshape(-1, kdim))
            stages[P + 'xpos'] = X(fo.XPOS, 3, b_order, b_sh, nb_t); stages[P + 'xmat'] = X(fo.XMAT, 9, b_order, b_sh, nb_t)
            stages[P + 'site_xpos'] = X(fo.SITE_XPOS, 3, s_order, s_sh, ns_t)
            M = np.zeros((nv_t, nv_t)); M[DS:, DS:] = o.get(fo.QM_DENSE).reshape(m.nv, m.nv)[:nv_w, :nv_w]; M[:DS, :DS] = np.eye(DS)
            stages[P + 'qM_dense'] = M
            for name, f in (('qfrc_bias', fo.QFRC_BIAS), ('qfrc_passive', fo.QFRC_PASSIVE), ('qfrc_actuator', fo.QFRC_ACTUATOR), ('qfrc_smooth', fo.QFRC_SMOOTH),
                            ('qfrc_constraint', fo.QFRC_CONSTRAINT)):
                stages[P + name] = qspace(o.get(f), nv_w, DS, nv_t)
            nc = int(o.get(fo.NCON)[0]); con = o.get(fo.CONTACT).reshape(-1, 16)[:nc]
            gi = inv(g_order)
            stages[P + 'con_geom1'] = np.array([gi[int(c[7])] + g_sh for c in con], np.int32); stages[P + 'con_geom2'] = np.array([gi[int(c[8])] + g_sh for c in con], np.int32)
            stages[P + 'con_dist'] = np.array([c[0] for c in con])
            seen_terrain_contact |= any(int(c[7]) == m.meta.get('hf_geom', -9) or int(c[8]) == m.meta.get('hf_geom', -9) for c in con)
        o.set(fo.CTRL, ctrl)
        ns = m.nsensordata
        o_sum = np.zeros(ns)
        for _ in range(n_sub):
            o.step2(); o.step1()
            sub.append(their_state())
        traj.append(their_state())
    assert variant != 'vision' or seen_terrain_contact, 'the self-test file should exercise heightfield contacts'
    for f in traj[0]:
        out['traj/' + f] = np.stack([r[f] for r in traj]); out['sub/' + f] = np.stack([r[f] for r in sub])
    # observations the way dm_control's buffers would report them
    for sname in ('walker/accelerometer', 'walker/gyro', 'walker/velocimeter'):
        if sname in names['sensor']:
            a = int(out['model/sensor_adr'][names['sensor'].index(sname)])
            sd, ssub = out['traj/sensordata'], out['sub/sensordata']
            out['obs/' + sname] = np.stack([sd[0, a:a + 3] / n_sub] + [ssub[k * n_sub:(k + 1) * n_sub, a:a + 3].mean(0) for k in range(n_steps)])
    out['stage_steps'] = np.array(stage_steps, np.int32)
    out.update(stages)
    meta = dict(task=variant, mujoco='none', dm_control='none', n_sub=n_sub, names=names, n_substep_steps=n_steps, source='oracle-selftest', format=1)
    out['__meta__'] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(path, **out)


@pytest.mark.parametrize('variant', ['walk', 'flight', 'vision'])
def test_consumer_plumbing_with_an_oracle_generated_file(variant, tmp_path):
    """every checker above, run on a file of the dump tool's layout that the oracle produced (permuted / shifted index spaces):
    proves the name maps, the teacher-forcing and the comparisons execute and close to round-off -- NOT a parity statement."""
    p = str(tmp_path / f'mujoco_{variant}.npz')
    write_selftest_file(p, variant)
    g = Golden(p)
    assert g.meta['source'] == 'oracle-selftest'
    m = load_model(variant)
    mp = check_compiler(m, g)
    check_observations(m, g, mp)
    w = check_stage_fields(m, g, mp, with_terrain(OracleBackend(m), g), 1e-10, 1e-7)
    assert 'qM' in w and 'xpos' in w
    r = check_teacher_forced(m, g, mp, with_terrain(OracleBackend(m), g), 1e-10, 1e-8, 1e-8)
    assert r['steps'] == 6 and r['p90_s'] is not None
    ge.build()
    be = with_terrain(StepperBackend(m, ge.EMU), g)
    check_stage_fields(m, g, mp, be, TOL['dev_smooth'], dev_con_tol(variant))
    check_teacher_forced(m, g, mp, be, TOL['dev_q'], 10 * TOL['dev_v'], 10 * TOL['dev_s'], max_events=0.34)
