/* flybody_b200.h -- C ABI of the B200-native batched fruit-fly physics stepper.
 *
 * The reference (TuragaLab/flybody) has no FFI of its own: the hot path sits behind
 * dm_control's `Physics.step()` which `composer.Environment.step()` calls n_sub_steps times per
 * control step (reference flybody/fly_envs.py:152-155).  These entry
 * points are what a ctypes binding replacing that `physics.step()` loop binds to; each one
 * names the reference-side call it stands in for.
 *
 * Conventions: plain C, no torch types.  Return 0 on success, negative code on error
 * (message via fb_last_error).  Caller owns host buffers; the library owns device buffers.
 * One handle = one device + one stream; calls on a handle are not re-entrant.
 * Device side: one fp32/int32 record per env (one warp steps one env); host-side buffers passed to fb_*
 * are rows [env][component] (what numpy/dm_env code holds).
 */
#ifndef FLYBODY_B200_H_
#define FLYBODY_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* enums shared by compiler (Python), oracle (C) and kernels (CUDA) */
enum { FB_GEOM_PLANE = 0, FB_GEOM_SPHERE = 2, FB_GEOM_CAPSULE = 3, FB_GEOM_ELLIPSOID = 4,
       FB_GEOM_CYLINDER = 5, FB_GEOM_BOX = 6 };
enum { FB_JNT_FREE = 0, FB_JNT_HINGE = 3 };
enum { FB_TRN_JOINT = 0, FB_TRN_TENDON = 1, FB_TRN_BODY = 2 };
enum { FB_DYN_NONE = 0, FB_DYN_FILTER = 2, FB_DYN_FILTEREXACT = 3 };
enum { FB_SENS_TOUCH = 0, FB_SENS_ACCELEROMETER = 1, FB_SENS_VELOCIMETER = 2, FB_SENS_GYRO = 3,
       FB_SENS_FORCE = 4 };

/* Flat model = the subset of mjModel the fly's mj_step touches (generated from
 * flybody_b200/flymodel.py:FIELDS; tests/test_model.py keeps the two in sync). */
/*@FBMODEL_BEGIN*/
typedef struct FbModel {
  int32_t nq;
  int32_t nv;
  int32_t nu;
  int32_t na;
  int32_t nbody;
  int32_t njnt;
  int32_t ngeom;
  int32_t npair;
  int32_t nsite;
  int32_t ntendon;
  int32_t nwrap;
  int32_t nsensor;
  int32_t nsensordata;
  int32_t nM;
  int32_t nfluid;
  int32_t opt_iterations;
  int32_t opt_ls_iterations;
  int32_t opt_noslip_iterations;
  int32_t opt_cone_elliptic;
  double opt_timestep;
  double opt_gravity[3];
  double opt_density;
  double opt_viscosity;
  double opt_wind[3];
  double opt_impratio;
  double opt_tolerance;
  double opt_ls_tolerance;
  double opt_noslip_tolerance;
  double stat_meaninertia;
  const int32_t* body_parentid;
  const int32_t* body_rootid;
  const int32_t* body_jntadr;
  const int32_t* body_jntnum;
  const int32_t* body_dofadr;
  const int32_t* body_dofnum;
  const int32_t* body_lastdof;
  const int32_t* body_fluid_ellipsoid;
  const double* body_pos;
  const double* body_quat;
  const double* body_ipos;
  const double* body_iquat;
  const double* body_mass;
  const double* body_inertia;
  const double* body_invweight0;
  const double* body_subtreemass;
  const int32_t* jnt_type;
  const int32_t* jnt_qposadr;
  const int32_t* jnt_dofadr;
  const int32_t* jnt_bodyid;
  const int32_t* jnt_limited;
  const double* jnt_pos;
  const double* jnt_axis;
  const double* jnt_stiffness;
  const double* jnt_range;
  const double* jnt_solref;
  const double* jnt_solimp;
  const double* jnt_margin;
  const double* qpos0;
  const double* qpos_spring;
  const int32_t* dof_bodyid;
  const int32_t* dof_jntid;
  const int32_t* dof_parentid;
  const int32_t* dof_Madr;
  const double* dof_armature;
  const double* dof_damping;
  const double* dof_invweight0;
  const int32_t* geom_type;
  const int32_t* geom_bodyid;
  const int32_t* geom_condim;
  const int32_t* geom_priority;
  const double* geom_size;
  const double* geom_pos;
  const double* geom_quat;
  const double* geom_rbound;
  const double* geom_friction;
  const double* geom_solmix;
  const double* geom_solref;
  const double* geom_solimp;
  const double* geom_margin;
  const double* geom_gap;
  const int32_t* pair_geom1;
  const int32_t* pair_geom2;
  const int32_t* fluid_bodyid;
  const double* fluid_pos;
  const double* fluid_quat;
  const double* fluid_size;
  const double* fluid_coef;
  const int32_t* site_bodyid;
  const int32_t* site_type;
  const double* site_pos;
  const double* site_quat;
  const double* site_size;
  const int32_t* tendon_adr;
  const int32_t* tendon_num;
  const int32_t* wrap_dofid;
  const int32_t* wrap_qposadr;
  const double* wrap_coef;
  const int32_t* actuator_trntype;
  const int32_t* actuator_trnid;
  const int32_t* actuator_dyntype;
  const int32_t* actuator_biastype;
  const int32_t* actuator_ctrllimited;
  const int32_t* actuator_forcelimited;
  const int32_t* actuator_actadr;
  const double* actuator_dynprm;
  const double* actuator_gainprm;
  const double* actuator_biasprm;
  const double* actuator_ctrlrange;
  const double* actuator_forcerange;
  const int32_t* sensor_type;
  const int32_t* sensor_objid;
  const int32_t* sensor_adr;
  const int32_t* sensor_dim;
} FbModel;
/*@FBMODEL_END*/

/* fields for fb_get / fb_set (per-env arrays; n = floats per env) */
enum FbField {
  FB_QPOS = 0,        /* nq   : physics.data.qpos                                   */
  FB_QVEL = 1,        /* nv   : physics.data.qvel                                   */
  FB_ACT = 2,         /* na   : physics.data.act                                    */
  FB_CTRL = 3,        /* nu   : physics.data.ctrl (physics.set_control)             */
  FB_QACC = 4,        /* nv   : physics.data.qacc (tasks/base.py:224)               */
  FB_QACC_WARMSTART = 5,
  FB_SENSORDATA = 6,  /* nsensordata : last-substep sensordata                      */
  FB_SENSOR_MEAN = 7, /* nsensordata : mean over the substeps of the last fb_step
                         (observable buffer_size + aggregator='mean', fruitfly.py:626-665) */
  FB_XPOS = 8,        /* nbody*3  : physics.bind(bodies).xpos                       */
  FB_XMAT = 9,        /* nbody*9  : physics.bind(bodies).xmat                       */
  FB_SITE_XPOS = 10,  /* nsite*3                                                    */
  FB_SITE_XMAT = 11,  /* nsite*9                                                    */
  FB_SUBTREE_COM = 12,/* nbody*3  : physics.data.subtree_com                        */
  FB_NCON = 13,       /* 1 (as float): number of detected contacts                  */
  FB_NEFC = 14,       /* 1 (as float): number of constraint rows                    */
  FB_TIME = 15,       /* 1 : physics.data.time                                      */
  FB_QFRC_SMOOTH = 16,/* nv : passive - bias + actuator                             */
  FB_QM_DENSE = 17,   /* nv*nv : dense joint-space inertia (debug / parity)         */
  FB_QFRC_CONSTRAINT = 18, /* nv                                                    */
  FB_SOLVER_NITER = 19,    /* 1                                                     */
  FB_QFRC_PASSIVE = 20,    /* nv                                                    */
  FB_QFRC_BIAS = 21,       /* nv                                                    */
  FB_QFRC_ACTUATOR = 22,   /* nv                                                    */
  FB_CONTACT = 23,    /* FB_MAXCON * 16: per contact [dist, pos3, normal3, geom1, geom2, dim, incl, mu, efc_adr, pad3] */
  FB_EFC_FORCE = 24,  /* FB_MAXEFC                                                  */
  FB_FLAGS = 25,      /* 1 (as float): bit0 bad state (nan/inf or |qacc|>1e14), bit1 contact overflow, bit2 efc overflow */
  FB_NFIELDS = 26
};

#define FB_MAXCON 64     /* contact slots per env   */
#define FB_MAXEFC 160    /* constraint rows per env */

typedef struct FbSim* FbHandle;

/* mjcf.Physics.from_mjcf_model + mj_makeData, batched: uploads the model, allocates SoA state
 * for n_envs environments on `device`, all envs at qpos0 / zero velocity.                     */
int fb_create(const FbModel* m, int n_envs, int device, FbHandle* out);
int fb_destroy(FbHandle h);

/* physics.reset_context()/bind(...).qpos = ... (walk_imitation.py:112-136): overwrite qpos/qvel
 * of the listed envs (AoS host arrays [n][nq], [n][nv]; NULL = model defaults), zero act /
 * warm start / time, then mj_forward for those envs.  env_ids == NULL means all envs.        */
int fb_reset(FbHandle h, const int32_t* env_ids, int n, const float* qpos, const float* qvel);

/* Auto-reset of a subset of envs inside a batched rollout (composer.Environment resets an env on the step()
 * after a LAST timestep and returns the reset observation without stepping): writes the reset state of the
 * listed envs on the device and marks them "held" -- the next fb_step recomputes their forward quantities
 * every substep but does not integrate them; the hold is cleared when that fb_step completes.            */
int fb_reset_hold(FbHandle h, const int32_t* env_ids, int n, const float* qpos, const float* qvel);

/* physics.set_control(ctrl) (fruitfly.py:540-544).  ctrl: contiguous rows [N][nu], on the host
 * (is_device=0, copied asynchronously on the handle's stream) or already on the device (is_device=1). */
int fb_set_ctrl(FbHandle h, const float* ctrl, int is_device);
/* FruitFly.apply_action (fruitfly.py:532-544) on the device: after this call fb_set_ctrl takes rows [N][n_action] in the
 * environment's ACTION order; ctrl_index[c] is the ctrl slot of action c (-1: no actuator, e.g. a user action), NaN
 * actions are applied as 0 (tasks/base.py:197-201).  ctrl_index == NULL restores plain ctrl rows.                  */
int fb_set_action_map(FbHandle h, const int32_t* ctrl_index, int n_action);

/* walker.set_pose / set_velocity on a subset of coordinates (walk_imitation.py:141-145):
 * field in {FB_QPOS, FB_QVEL, FB_ACT}; vals is [N][k] AoS host, written to coordinates idx[k]
 * of every env.  No forward pass is run (dm_control does not run one either).               */
int fb_write_state(FbHandle h, int field, const int32_t* idx, int k, const float* vals);

/* n_substeps x physics.step() with legacy_step=True, i.e. (mj_step2; mj_step1) per substep
 * (SURVEY.md App. A), accumulating the per-substep sensor mean.  Asynchronous on the handle's
 * stream.                                                                                     */
int fb_step(FbHandle h, int n_substeps);

/* mj_forward on all envs (after external state writes).                                       */
int fb_forward(FbHandle h);

/* Read a per-env field.  is_device=0: dst is host [N][n] AoS float32, synchronises.
 * is_device=1: *(void**)dst receives the borrowed device pointer of env 0's slots; consecutive envs
 * are fb_record_stride() 4-byte slots apart.                                                  */
int fb_get(FbHandle h, int field, void* dst, int is_device);
int fb_field_size(FbHandle h, int field);   /* floats per env, <0 on error */
int fb_record_stride(FbHandle h);           /* 4-byte slots between the records of consecutive envs */
int fb_set(FbHandle h, int field, const float* src);  /* host [N][n] AoS -> device, all envs */

/* Packed observation buffer of the last control step, [N][floats_per_env] AoS on device, for the
 * NCCL gather to rank 0.  Layout: qpos, qvel, act, sensor_mean, xpos/xmat of
 * root, site_xpos.                                                                            */
int fb_obs_ptr(FbHandle h, void** dev_ptr, int* floats_per_env);
/* Task observation program: the observables of the reference task (FruitFlyObservables, fruitfly.py:585-756;
 * ref_displacement / ref_root_quat, tasks/base.py:245-268) evaluated on the device right after fb_step, one
 * fp32 row per env in the task's final observation layout.  Items are concatenated in order.           */
enum FbObsItem {
  FB_OBS_SENSOR_MEAN = 0, /* a = sensordata adr, b = len : per-substep mean (first[e]: single sample / n_sub) */
  FB_OBS_SENSOR_NOW = 1,  /* a = adr, b = len           : last sample (termination checks)                   */
  FB_OBS_ACT = 2,         /* a = first act, b = len                                                          */
  FB_OBS_QPOS = 3,        /* a = offset into list, b = len : qpos[list[a..a+b)]                              */
  FB_OBS_QVEL = 4,
  FB_OBS_SITES_EGO = 5,   /* a = offset into list (site ids), b = count : (site_xpos - root_xpos) @ root_xmat */
  FB_OBS_ROOT_ZAXIS = 6,  /* root xmat[6:9]                                                                  */
  FB_OBS_REF_DISP = 7,    /* b = future_steps+1 : (ref_pos[step+i] - root_pos) @ root_xmat                   */
  FB_OBS_REF_QUAT = 8,    /* b = future_steps+1 : root_quat^-1 * ref_quat[step+i]                            */
  FB_OBS_SCALARS = 9,     /* flags, |qacc|^2, time                                                           */
  FB_OBS_ROOT_POSE = 10,  /* root xpos[3] + quat[4]                                                          */
  FB_OBS_SUBTREE_COM = 11,/* a = body id : physics.data.subtree_com[body]                                    */
  FB_OBS_DOF_AXIS_EGO = 12,/* a = offset into list (dof ids), b = count : world joint axis (physics.bind(joints).xaxis,
                              tasks/rewards.py:48-50) rotated into the root frame                                */
  FB_OBS_TASK_TARGET = 14,   /* b floats of the device task's per-episode targets (kind 2: target height, target speed = the reference's
                               `task_input` observable, tasks/vision_flight.py:56-76)                                     */
  FB_OBS_WORLD_CONTACT = 13 /* 1 float: 1 if an active contact (efc_address >= 0) involves a geom of the world body (ground plane, terrain):
                               the reference's check_floor_contact (tasks/vision_flight.py:235-247)                        */
};
typedef struct FbObsProgram {
  int32_t n_items; const int32_t* kind; const int32_t* a; const int32_t* b;
  int32_t n_list; const int32_t* list;
  int32_t root_body, n_sub;
  int32_t ref_len; const float* ref_qpos;      /* [ref_len][7] reference root trajectory */
} FbObsProgram;
/* Upload the program (and reference table); returns the row length in floats (<0 on error).             */
int fb_obs_program(FbHandle h, const FbObsProgram* p);
/* Per-env reference tables: every env tracks its own reference trajectory (one dataset snippet per episode, reference
 * tasks/walk_imitation.py:93-105, tasks/flight_imitation.py:88-106).  Called after fb_obs_program, fb_ref_slots switches
 * FB_OBS_REF_DISP / FB_OBS_REF_QUAT to a device table [n_envs][slot_len][7]; fb_ref_slot_write fills the slots of the
 * listed envs from rows [n][slot_len][7] (the caller pads a shorter snippet with its last row).  The step index given to
 * fb_task_inputs then counts inside the env's own slot.                                                  */
int fb_ref_slots(FbHandle h, int slot_len);
int fb_ref_slot_write(FbHandle h, const int32_t* env_ids, int n, const float* rows);
/* Per control step: index into the reference table and "first step after reset" flag of every env.      */
int fb_task_inputs(FbHandle h, const int32_t* step_idx, const uint8_t* first);
/* Copy the task observation rows [N][row_len] to a (pinned) host buffer (fb_pack_obs must have run).    */
int fb_read_task_obs(FbHandle h, float* host_dst);
/* Launch the pack kernel (after fb_step) / copy the packed rows to a (pinned) host buffer [N][floats_per_env]:
 * qpos, qvel, act, sensor_mean, sensordata, root xpos[3], root xmat[9], site_xpos[3*nsite], flags, |qacc|^2, time.
 * Stands in for the observation_updater reads of composer.Environment.step.        */
int fb_pack_obs(FbHandle h);
int fb_read_obs(FbHandle h, float* host_dst);

/* Device-side task logic: everything the reference's task hooks do around the physics step -- composer
 * auto-reset + initialize_episode (tasks/walk_imitation.py:112-136, tasks/flight_imitation.py:113-144), before_step (ghost
 * placement walk_imitation.py:138-150; wing-beat pattern generator flight_imitation.py:146-168, pattern_generators.py:131-203),
 * check_termination / get_reward / get_discount (tasks/base.py:203-225, walk_imitation.py:179-203,
 * flight_imitation.py:170-226) -- evaluated on the device, so that a control step needs no host round trip: actions in
 * (host or device pointer), observation rows + (reward, discount, step_type) out.  Covers the shared-reference (inference-mode)
 * tasks; dataset mode keeps the host-side task code.  Requires fb_set_action_map and fb_obs_program first.          */
typedef struct FbTaskProgram {
  int32_t kind;                        /* 0 walk_imitation, 1 flight_imitation, 2 vision_guided_flight (fields at the end of the struct) */
  int32_t root_qadr, root_vadr, ghost_qadr, ghost_vadr;   /* free-joint slots of the walker root and of the ghost */
  int32_t user_col;                    /* column of the beat-frequency action in the action row, -1 if none */
  float ghost_offset[3];
  float control_timestep, time_limit;
  float terminal_com_dist, terminal_linvel, terminal_angvel, terminal_qacc, terminal_height;
  int32_t velocimeter_adr, gyro_adr;   /* sensordata addresses */
  int32_t com_body;                    /* body whose subtree CoM the flight reward tracks */
  int32_t episode_steps;               /* step index that ends the episode (end of the reference) */
  int32_t ref_len; const float* ref_qpos /* [ref_len][7] */; const float* ref_qvel /* [ref_len][6] */;
  int32_t obs_refdisp_off, obs_refquat_off;   /* offsets of ref_displacement / ref_root_quat inside the observation row */
  const float* reset_qpos;             /* [nq] episode start pose (root / ghost slots are overwritten from the reference) */
  int32_t n_noise; const int32_t* noise_qadr; float noise_amp; uint32_t seed;   /* U(-amp, amp) on these qpos at reset */
  int32_t n_wing; const int32_t* wing_qadr; const int32_t* wing_vadr; const int32_t* wing_ctrl;
  int32_t n_freq, tab_len;             /* wing-beat tables: one resampled pattern per discrete beat frequency */
  const float* wb_traj /* [n_freq][tab_len][n_wing] */; const float* wb_phase; const float* wb_phase_mod /* [n_freq][tab_len], +inf padded */;
  const float* wb_freqs /* [n_freq] */; const int32_t* wb_len /* [n_freq] */;
  float wb_base_freq, wb_rel_range, wb_rate;
  float com_offset[3];                 /* root -> CoM offset in the root frame (tasks/task_utils.py:237) */
  /* kind 2, vision_guided_flight (tasks/vision_flight.py:97-254): no ghost / reference; per episode a target height and speed, a start
   * point, a wing-beat phase and a terrain of the device bank (fb_hfield_bank) are drawn; reward = product of the height / forward speed /
   * speed / side speed / body axis factors and, for 'trench' arenas, the distance to the corridor's centre line (vision_flight.py:214-226:
   * y of the centre line at trench_len[k] equally spaced x in [trench_x[k][0], trench_x[k][1]] for bank terrain k);
   * termination on a bad state or, if floor_contacts_fatal, an active contact with a world geom.                                */
  float target_height_range[2], target_speed_range[2], init_x_range[2], init_y_range[2];
  float hover_quat[4], target_zaxis[3];
  int32_t floor_contacts_fatal;
  int32_t trench_cap;                  /* 0: no centre-line factor; else the row stride of trench_y */
  const float* trench_x /* [n_terrain][2] first / last x */; const int32_t* trench_len /* [n_terrain] samples */; const float* trench_y /* [n_terrain][trench_cap] */;
} FbTaskProgram;
int fb_task_program(FbHandle h, const FbTaskProgram* p);
/* One control step with the task logic on the device: [auto-reset] -> action -> ctrl -> before_step -> n_substeps x physics ->
 * observation program -> termination / reward.  `action` rows [n_envs][n_action] (fb_set_action_map), host or device.  */
int fb_task_step(FbHandle h, const float* action, int is_device, int n_substeps);
/* Mark every env for reset at the next fb_task_step (env.reset()).                                                    */
int fb_task_reset_all(FbHandle h);
/* Amplitude of the U(-a, a) joint noise the device-side reset adds to the listed start-pose joints (FbTaskProgram.noise_amp), changed
 * without re-uploading the program: e.g. noise for the very first reset only (decorrelated start states), exact start pose at the
 * auto-resets, as the reference's initialize_episode has it.                                                            */
int fb_task_set_reset_noise(FbHandle h, float amp);
/* Mark the listed envs for reset at the next fb_task_step, whatever their episode state (an actor restarting single
 * environments; bench.py's pre-roll uses it to spread the envs over the phases of an episode).                          */
int fb_task_request_reset(FbHandle h, const int32_t* env_ids, int n);
/* Per-env count of episodes started so far (resets done by the device-side task logic), int32 [n_envs] to the host.    */
int fb_task_episodes(FbHandle h, int32_t* dst);
/* Terrain bank for the device-side vision task: K heightfields [K][nrow * ncol] (world units, the grid of fb_hfield_collision /
 * fb_eye_program); a resetting env copies one of them into its own heightfield (collision + eyes).                      */
int fb_hfield_bank(FbHandle h, int n_terrain, const float* heights);
/* Rows of 8 uniform numbers in [0,1) consumed by the listed envs' next reset instead of the device's counter hash (kind 2: target
 * height, target speed, start x, start y, wing-beat phase, terrain pick; kind 1: wing-beat phase) -- lets a test feed the device the
 * draws of the host-side task code.                                                                                    */
int fb_task_uniform_rows(FbHandle h, const int32_t* env_ids, int n, const float* u);
/* Uniform numbers in [0,1) consumed by the listed envs' next reset (flight: wing-beat phase); without them the device
 * draws from a counter hash of (seed, env, episode).                                                                 */
int fb_task_uniforms(FbHandle h, const int32_t* env_ids, int n, const float* u);
/* Results of the last fb_task_step: borrowed device pointers (obs rows [n_envs][obs_dim], out rows [n_envs][4] = reward,
 * discount, step_type 0 FIRST / 1 MID / 2 LAST, 0) and the host copy (synchronises).                                  */
int fb_task_ptrs(FbHandle h, void** obs_dev, int* obs_dim, void** out_dev);
int fb_task_read(FbHandle h, float* obs_host, float* out_host);
/* Eye cameras (reference FruitFlyObservables.right_eye / left_eye, fruitfly/fruitfly.py:729-745; cameras fruitfly.xml:335-336;
 * 32 x 32 RGB, fovy 150 deg in tasks/vision_flight.py:23-24): a ray caster over a per-env heightfield terrain
 * (tasks/arenas/hills.py), the ground plane and a sky, from the current body poses.  MuJoCo's camera model (pinhole, -z forward,
 * +y up, vertical fovy, row 0 at the top); not MuJoCo's OpenGL image.                                                  */
typedef struct FbEyeProgram {
  int32_t n_cam;                         /* 1 or 2 */
  int32_t body[2]; float pos[2][3]; float quat[2][4];   /* camera frames in the frame of the body they are attached to */
  float fovy_deg; int32_t size;          /* vertical field of view; images are size x size */
  int32_t nrow, ncol; float half_size;   /* heightfield grid over [-half_size, half_size]^2 (row = y, col = x); nrow = 0: none */
  float z_offset, zfar;                  /* height of the ground plane / terrain base (hills.py:210), far clipping distance */
  float sky_top[3], sky_horizon[3], ground[3], ambient, diffuse;   /* colours in [0,1]; headlight terms (hills.py:248-250) */
} FbEyeProgram;
int fb_eye_program(FbHandle h, const FbEyeProgram* p);
/* Heightfield collision (MuJoCo mjc_ConvexHField; the terrain of tasks/arenas/hills.py): `geom` is the model's heightfield geom
 * (pose and contact parameters come from the model), size = (x half extent, y half extent, elevation scale, base depth), the grid
 * is nrow x ncol, pair_geom lists the geoms that can touch it.  Adds one kernel between collision and constraint rows.  The
 * per-env heights are the ones of fb_hfield_write (shared with the eye cameras), as fractions of the elevation scale.       */
int fb_hfield_collision(FbHandle h, int geom, const float* size, int nrow, int ncol, const int32_t* pair_geom, int npair);
/* Heights (world units) of the listed envs' terrains, rows [n][nrow * ncol]; envs never written are flat.            */
int fb_hfield_write(FbHandle h, const int32_t* env_ids, int n, const float* heights);
/* Render every env's eyes from the current poses into the library's buffer [n_envs][n_cam][size][size][3] (uint8).   */
int fb_render_eyes(FbHandle h);
int fb_eyes_ptr(FbHandle h, void** dev_ptr, int* bytes_per_env);
int fb_eyes_read(FbHandle h, uint8_t* host_dst);
int fb_n_envs(FbHandle h);
int fb_n_envs_padded(FbHandle h);
void* fb_stream(FbHandle h);                 /* cudaStream_t the handle launches on */
int fb_sync(FbHandle h);
long long fb_launch_count(FbHandle h);       /* kernels launched by this handle so far */
/* device time (ms) of the last fb_step measured with CUDA events on the handle's stream */
float fb_last_step_ms(FbHandle h);
/* solver configuration: tolerance and iteration cap of the constraint solver */
/* per-kernel device timing (CUDA events around every launch while enabled): cumulative ms and launch
 * counts per pipeline stage; fb_profile_name(kind) names the stage.  Used by bench.py's roofline block. */
int fb_profile(FbHandle h, int enable);
int fb_profile_read(FbHandle h, double* ms, long long* counts, int n);
const char* fb_profile_name(int kind);
int fb_set_solver(FbHandle h, float tolerance, int max_iter);
const char* fb_last_error(FbHandle h);
const char* fb_version(void);

#ifdef __cplusplus
}
#endif
#endif
