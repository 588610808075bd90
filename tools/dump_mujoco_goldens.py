#!/usr/bin/env python
"""MuJoCo pin kit, producer side: dump golden vectors from the REFERENCE (flybody + dm_control + mujoco).

Run this ONCE on any machine where the reference runs (`pip install flybody` or a checkout of TuragaLab/flybody with its
dependencies; no GPU needed):

    python tools/dump_mujoco_goldens.py [--reference /path/to/flybody/checkout] [--out tests/golden] [--steps 200]

It writes `tests/golden/mujoco_walk.npz`, `mujoco_flight.npz` and `mujoco_vision.npz` (a few MB each; the vision file carries the
episode's heightfield as hfield/data + hfield/size and needs an OpenGL backend for the eye cameras).  Commit them; from then on
`pytest tests/test_mujoco_goldens.py` pins (a) the model compiler, (b) the fp64 oracle and (c, `-m gpu`) the CUDA stepper against a
real MuJoCo, and DESIGN.md's "parity unpinned" caveat can go.  Nothing in this repository imports mujoco at test or run time: the
`.npz` files are the only thing that travels.  (The build container and the GPU boxes have no mujoco / dm_control and no network:
`pip download mujoco` fails there, which is why this kit exists instead of the files.)

What is recorded, per task (`walk_imitation(terminal_com_dist=inf)` as `tests/test_walking_env.py:40`, `flight_imitation()`):

  meta (json)        mujoco / dm_control versions, names of bodies / joints / geoms / sites / actuators / sensors, action names,
                     n_sub, timesteps, the seed of the action stream
  model/<field>      the mjModel arrays the step reads (sizes, options, body / joint / dof / geom / site / tendon / actuator /
                     sensor tables) -> checks `flybody_b200.compiler`
  traj/actions       [T, A]   np.random.RandomState(0).uniform(-.5, .5, (T, A))  (walk; flight: U(-.2, .2), task_utils.py:58-65)
  traj/<x>           [T+1, .] state after reset (row 0) and after every control step: qpos qvel act ctrl qacc qacc_warmstart
                     sensordata time ncon nefc, plus every observation of the TimeStep as obs/<name>, reward, discount, step_type
  sub/<x>            [S, .]   the same state fields after EVERY physics substep of the first `--substep-steps` control steps
                     (the per-substep sensor samples the observation buffers average; reference fruitfly.py:626-665)
  stage/<k>/<x>      intermediates of mj_forward on `--stages` recorded states: xpos xmat xipos ximat geom_xpos geom_xmat
                     site_xpos site_xmat subtree_com cvel cdof cinert qM (dense) qLD qfrc_bias qfrc_passive qfrc_actuator
                     actuator_force qfrc_smooth qacc_smooth contacts (dist pos frame geom1 geom2 dim efc_address includemargin
                     friction solref solimp) efc_J (dense) efc_type efc_pos efc_margin efc_D efc_R efc_aref efc_b efc_force
                     qfrc_constraint qacc sensordata solver_niter

Field names follow mjData / mjModel.  Everything is float64 / int32.
"""
import argparse
import json
import os
import sys

import numpy as np


def _names(model, objtype, n):
    import mujoco
    out = []
    for i in range(n):
        s = mujoco.mj_id2name(model, objtype, i)
        out.append(s if s is not None else '')
    return out


MODEL_FIELDS = [
    'body_parentid', 'body_rootid', 'body_jntadr', 'body_jntnum', 'body_dofadr', 'body_dofnum', 'body_pos', 'body_quat',
    'body_ipos', 'body_iquat', 'body_mass', 'body_inertia', 'body_invweight0', 'body_subtreemass', 'body_gravcomp',
    'jnt_type', 'jnt_qposadr', 'jnt_dofadr', 'jnt_bodyid', 'jnt_limited', 'jnt_pos', 'jnt_axis', 'jnt_stiffness', 'jnt_range',
    'jnt_solref', 'jnt_solimp', 'jnt_margin', 'qpos0', 'qpos_spring',
    'dof_bodyid', 'dof_jntid', 'dof_parentid', 'dof_Madr', 'dof_armature', 'dof_damping', 'dof_invweight0', 'dof_frictionloss',
    'geom_type', 'geom_bodyid', 'geom_contype', 'geom_conaffinity', 'geom_condim', 'geom_priority', 'geom_size', 'geom_pos',
    'geom_quat', 'geom_rbound', 'geom_friction', 'geom_solmix', 'geom_solref', 'geom_solimp', 'geom_margin', 'geom_gap',
    'geom_fluid', 'geom_dataid',
    'site_bodyid', 'site_type', 'site_pos', 'site_quat', 'site_size',
    'tendon_adr', 'tendon_num', 'wrap_type', 'wrap_objid', 'wrap_prm', 'tendon_stiffness', 'tendon_damping', 'tendon_limited',
    'actuator_trntype', 'actuator_trnid', 'actuator_dyntype', 'actuator_gaintype', 'actuator_biastype', 'actuator_ctrllimited',
    'actuator_forcelimited', 'actuator_actlimited', 'actuator_actadr', 'actuator_dynprm', 'actuator_gainprm', 'actuator_biasprm',
    'actuator_ctrlrange', 'actuator_forcerange', 'actuator_gear',
    'sensor_type', 'sensor_objtype', 'sensor_objid', 'sensor_adr', 'sensor_dim', 'sensor_cutoff', 'sensor_noise',
    'exclude_signature', 'hfield_size', 'hfield_nrow', 'hfield_ncol',
]
OPT_FIELDS = ['timestep', 'gravity', 'wind', 'density', 'viscosity', 'impratio', 'tolerance', 'ls_tolerance', 'noslip_tolerance',
              'iterations', 'ls_iterations', 'noslip_iterations', 'cone', 'jacobian', 'solver', 'integrator', 'disableflags',
              'enableflags', 'o_margin']
STATE_FIELDS = ['qpos', 'qvel', 'act', 'ctrl', 'qacc', 'qacc_warmstart', 'sensordata']


def dump_model(model):
    out = {}
    for f in MODEL_FIELDS:
        try:
            out['model/' + f] = np.array(getattr(model, f))
        except Exception as e:                      # field renamed / absent in this mujoco version
            print(f'  (model.{f} not available: {e})')
    for f in OPT_FIELDS:
        try:
            out['model/opt_' + f] = np.array(getattr(model.opt, f))
        except Exception as e:
            print(f'  (model.opt.{f} not available: {e})')
    for f in ('meaninertia', 'meanmass', 'meansize', 'extent'):
        out['model/stat_' + f] = np.array(getattr(model.stat, f))
    for f in ('nq', 'nv', 'nu', 'na', 'nbody', 'njnt', 'ngeom', 'nsite', 'ntendon', 'nwrap', 'nsensor', 'nsensordata', 'nM',
              'nmesh', 'nexclude', 'npair', 'nhfield'):
        out['model/' + f] = np.array(getattr(model, f))
    return out


def state_row(data):
    row = {f: np.array(getattr(data, f), np.float64) for f in STATE_FIELDS}
    row['time'] = np.array([data.time])
    row['ncon'] = np.array([data.ncon], np.int32)
    row['nefc'] = np.array([data.nefc], np.int32)
    return row


def dense_efc_J(model, data):
    """efc_J as [nefc, nv] whatever the Jacobian storage is."""
    import mujoco
    nefc, nv = data.nefc, model.nv
    if nefc == 0:
        return np.zeros((0, nv))
    if mujoco.mj_isSparse(model):
        J = np.zeros((nefc, nv))
        nnz, adr, col, val = (np.array(data.efc_J_rownnz), np.array(data.efc_J_rowadr), np.array(data.efc_J_colind),
                              np.array(data.efc_J).ravel())
        for r in range(nefc):
            J[r, col[adr[r]:adr[r] + nnz[r]]] = val[adr[r]:adr[r] + nnz[r]]
        return J
    return np.array(data.efc_J).reshape(nefc, nv)


def dump_stage(model, src, prefix):
    """mj_forward on a copy of (model, src state) with intermediates; the copy keeps the caller's data untouched."""
    import mujoco
    d = mujoco.MjData(model)
    d.qpos[:] = src.qpos; d.qvel[:] = src.qvel; d.ctrl[:] = src.ctrl; d.time = src.time
    if model.na:
        d.act[:] = src.act
    d.qacc_warmstart[:] = src.qacc_warmstart
    mujoco.mj_forward(model, d)
    out = {}
    put = lambda k, v, dt=np.float64: out.__setitem__(prefix + k, np.array(v, dt))
    for f in ('qpos', 'qvel', 'act', 'ctrl', 'qacc_warmstart', 'xpos', 'xquat', 'xmat', 'xipos', 'ximat', 'geom_xpos', 'geom_xmat',
              'site_xpos', 'site_xmat', 'subtree_com', 'cvel', 'cdof', 'cinert', 'qLD', 'qLDiagInv', 'actuator_length', 'actuator_velocity',
              'actuator_force', 'act_dot', 'qfrc_bias', 'qfrc_passive', 'qfrc_actuator', 'qfrc_smooth', 'qacc_smooth', 'qfrc_constraint',
              'qacc', 'sensordata', 'cacc', 'cfrc_int', 'cfrc_ext'):
        try:
            put(f, getattr(d, f))
        except Exception as e:
            print(f'  (data.{f} not available: {e})')
    try:
        mujoco.mj_rnePostConstraint(model, d)      # cacc / cfrc_int / cfrc_ext as the force / accelerometer sensors see them
        put('cacc', d.cacc); put('cfrc_int', d.cfrc_int); put('cfrc_ext', d.cfrc_ext)
    except Exception as e:
        print(f'  (mj_rnePostConstraint: {e})')
    M = np.zeros((model.nv, model.nv))
    mujoco.mj_fullM(model, M, d.qM)
    put('qM_dense', M)
    nc = d.ncon
    put('ncon', [nc], np.int32)
    con = d.contact
    put('con_dist', [con[i].dist for i in range(nc)]); put('con_pos', [con[i].pos for i in range(nc)] or np.zeros((0, 3)))
    put('con_frame', [con[i].frame for i in range(nc)] or np.zeros((0, 9)))
    put('con_geom1', [con[i].geom1 for i in range(nc)], np.int32); put('con_geom2', [con[i].geom2 for i in range(nc)], np.int32)
    put('con_dim', [con[i].dim for i in range(nc)], np.int32); put('con_efc_address', [con[i].efc_address for i in range(nc)], np.int32)
    put('con_includemargin', [con[i].includemargin for i in range(nc)]); put('con_friction', [con[i].friction for i in range(nc)] or np.zeros((0, 5)))
    put('con_solref', [con[i].solref for i in range(nc)] or np.zeros((0, 2))); put('con_solimp', [con[i].solimp for i in range(nc)] or np.zeros((0, 5)))
    put('con_exclude', [con[i].exclude for i in range(nc)], np.int32)
    put('nefc', [d.nefc], np.int32)
    put('efc_J', dense_efc_J(model, d))
    for f in ('efc_type', 'efc_id'):
        put(f, np.array(getattr(d, f))[:d.nefc], np.int32)
    for f in ('efc_pos', 'efc_margin', 'efc_D', 'efc_R', 'efc_aref', 'efc_b', 'efc_force', 'efc_diagApprox', 'efc_KBIP', 'efc_vel'):
        try:
            a = np.array(getattr(d, f))
            put(f, a[:d.nefc] if a.ndim == 1 else a.reshape(-1, a.shape[-1])[:d.nefc])
        except Exception as e:
            print(f'  (data.{f} not available: {e})')
    try:
        put('solver_niter', np.array(d.solver_niter).ravel()[:1], np.int32)
    except Exception:
        pass
    return out


def record_task(name, env, actions, n_substep_steps, stage_every, mujoco_mod):
    physics = env.physics
    model, data = physics.model.ptr, physics.data.ptr
    n_sub = int(round(env.control_timestep() / physics.timestep()))
    out = dump_model(model)
    mj = mujoco_mod
    names = dict(body=_names(model, mj.mjtObj.mjOBJ_BODY, model.nbody), jnt=_names(model, mj.mjtObj.mjOBJ_JOINT, model.njnt),
                 geom=_names(model, mj.mjtObj.mjOBJ_GEOM, model.ngeom), site=_names(model, mj.mjtObj.mjOBJ_SITE, model.nsite),
                 actuator=_names(model, mj.mjtObj.mjOBJ_ACTUATOR, model.nu), sensor=_names(model, mj.mjtObj.mjOBJ_SENSOR, model.nsensor),
                 tendon=_names(model, mj.mjtObj.mjOBJ_TENDON, model.ntendon))
    spec = env.action_spec()
    sub_rows = []
    recording = {'on': False}

    def after_substep(*_a, **_k):
        if recording['on']:
            sub_rows.append(state_row(data))
    try:
        env.add_extra_hook('after_substep', after_substep)
    except Exception as e:                         # older dm_control: no extra hooks -> no per-substep rows
        print('  (no after_substep hook:', e, ')')
    ts = env.reset()
    if int(model.nhfield) > 0:       # the episode's terrain (vision_guided_flight: regenerated by initialize_episode_mjcf at every reset)
        nrow, ncol, adr = int(model.hfield_nrow[0]), int(model.hfield_ncol[0]), int(model.hfield_adr[0])
        out['hfield/data'] = np.array(model.hfield_data[adr:adr + nrow * ncol], np.float64).reshape(nrow, ncol)      # normalised: x elevation = world
        out['hfield/size'] = np.array(model.hfield_size[0], np.float64)                                           # radius x, radius y, elevation z, base z
    rows = [state_row(data)]
    obs_rows = [{k: np.array(v, np.float64) for k, v in ts.observation.items()}]
    rew, disc, stype = [0.0], [1.0], [int(ts.step_type)]
    stages = {}
    stage_ids = []
    for k in range(len(actions)):
        if k % stage_every == 0:
            stage_ids.append(k)
            stages.update(dump_stage(model, data, f'stage/{len(stage_ids) - 1}/'))
        recording['on'] = k < n_substep_steps
        ts = env.step(actions[k])
        rows.append(state_row(data))
        obs_rows.append({kk: np.array(v, np.float64) for kk, v in ts.observation.items()})
        rew.append(float(ts.reward) if ts.reward is not None else 0.0)
        disc.append(float(ts.discount) if ts.discount is not None else 1.0)
        stype.append(int(ts.step_type))
        if ts.last():
            print(f'  {name}: episode ended at control step {k + 1} (step_type LAST); the remaining actions are not replayed')
            actions = actions[:k + 1]
            break
    for f in rows[0]:
        out['traj/' + f] = np.stack([r[f] for r in rows])
    for f in obs_rows[0]:
        out['obs/' + f] = np.stack([r[f] for r in obs_rows])
    if sub_rows:
        for f in sub_rows[0]:
            out['sub/' + f] = np.stack([r[f] for r in sub_rows])
    out['traj/actions'] = np.asarray(actions, np.float64)
    out['traj/reward'] = np.array(rew); out['traj/discount'] = np.array(disc); out['traj/step_type'] = np.array(stype, np.int32)
    out['stage_steps'] = np.array(stage_ids, np.int32)
    out.update(stages)
    import dm_control
    meta = dict(task=name, mujoco=mj.__version__, dm_control=getattr(dm_control, '__version__', 'unknown'), n_sub=n_sub,
                physics_timestep=float(physics.timestep()), control_timestep=float(env.control_timestep()), names=names,
                action_names=str(spec.name).split('\t'), action_min=np.asarray(spec.minimum).tolist(), action_max=np.asarray(spec.maximum).tolist(),
                obs_names=list(ts.observation.keys()), n_substep_steps=int(min(n_substep_steps, len(actions))), source='mujoco',
                format=1)
    out['__meta__'] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--reference', default=None, help='checkout of TuragaLab/flybody to import instead of an installed `flybody`')
    ap.add_argument('--out', default=os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden'))
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--substep-steps', type=int, default=20, help='control steps whose every physics substep is recorded')
    ap.add_argument('--stages', type=int, default=20, help='number of states with mj_forward intermediates')
    args = ap.parse_args()
    if args.reference:
        sys.path.insert(0, args.reference)
    import mujoco
    from flybody.fly_envs import walk_imitation, flight_imitation
    os.makedirs(args.out, exist_ok=True)
    T = args.steps
    every = max(1, T // args.stages)
    # BASELINE.json configs[0]: walk_imitation, 200 random-action steps, terminal_com_dist = inf as
    # tests/test_walking_env.py:40 so that the 2 cm/s ghost walking away does not end the episode
    env = walk_imitation(terminal_com_dist=float('inf'))
    A = env.action_spec().shape[0]
    acts = np.random.RandomState(0).uniform(-0.5, 0.5, (T, A))
    out = record_task('walk', env, acts, args.substep_steps, every, mujoco)
    np.savez_compressed(os.path.join(args.out, 'mujoco_walk.npz'), **out)
    print('wrote mujoco_walk.npz:', len(out), 'arrays,', out['traj/qpos'].shape[0] - 1, 'control steps')
    # flight_imitation defaults (synthetic 200-step trajectory); random policy of tasks/task_utils.py:58-65
    env = flight_imitation()
    A = env.action_spec().shape[0]
    acts = np.random.RandomState(0).uniform(-0.2, 0.2, (min(T, 150), A))
    out = record_task('flight', env, acts, args.substep_steps, max(1, len(acts) // args.stages), mujoco)
    np.savez_compressed(os.path.join(args.out, 'mujoco_flight.npz'), **out)
    print('wrote mujoco_flight.npz:', len(out), 'arrays,', out['traj/qpos'].shape[0] - 1, 'control steps')
    # vision_guided_flight ('bumps' arena, synthetic wing-beat pattern): terrain contacts (mjc_ConvexHField) and the task's observables.
    # The eye cameras render through MuJoCo's OpenGL context: on a headless machine set MUJOCO_GL=egl (or osmesa) first.
    try:
        from flybody.fly_envs import vision_guided_flight
        env = vision_guided_flight(bumps_or_trench='bumps')
        A = env.action_spec().shape[0]
        acts = np.random.RandomState(0).uniform(-0.2, 0.2, (min(T, 100), A))
        out = record_task('vision', env, acts, args.substep_steps, max(1, len(acts) // args.stages), mujoco)
        np.savez_compressed(os.path.join(args.out, 'mujoco_vision.npz'), **out)
        print('wrote mujoco_vision.npz:', len(out), 'arrays,', out['traj/qpos'].shape[0] - 1, 'control steps')
    except Exception as e:
        print('vision_guided_flight NOT recorded (the eye cameras need an OpenGL backend: MUJOCO_GL=egl or osmesa):', repr(e))
    print('now: git add tests/golden/mujoco_*.npz && python -m pytest tests/test_mujoco_goldens.py -q')


if __name__ == '__main__':
    main()
