"""`FlyModel`: the flat model the stepper reads, and its ctypes mirror of `FbModel`
(`include/flybody_b200.h`).  The field table below is the single source of truth: the C struct
text in the header is generated from it (`c_struct_text`) and a test keeps them in sync.

Arrays are the subset of `mjModel` the fly's `mj_step` pipeline touches.
"""
import ctypes as C
import json
import os

import numpy as np

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets')

# (name, kind) kind: 'i' int32 scalar, 'd' double scalar, 'd3' double[3], 'pi' const int32*, 'pd' const double*
FIELDS = [
    # sizes
    ('nq', 'i'), ('nv', 'i'), ('nu', 'i'), ('na', 'i'), ('nbody', 'i'), ('njnt', 'i'), ('ngeom', 'i'),
    ('npair', 'i'), ('nsite', 'i'), ('ntendon', 'i'), ('nwrap', 'i'), ('nsensor', 'i'),
    ('nsensordata', 'i'), ('nM', 'i'), ('nfluid', 'i'),
    # options
    ('opt_iterations', 'i'), ('opt_ls_iterations', 'i'), ('opt_noslip_iterations', 'i'),
    ('opt_cone_elliptic', 'i'),
    ('opt_timestep', 'd'), ('opt_gravity', 'd3'), ('opt_density', 'd'), ('opt_viscosity', 'd'),
    ('opt_wind', 'd3'), ('opt_impratio', 'd'), ('opt_tolerance', 'd'), ('opt_ls_tolerance', 'd'),
    ('opt_noslip_tolerance', 'd'), ('stat_meaninertia', 'd'),
    # bodies
    ('body_parentid', 'pi'), ('body_rootid', 'pi'), ('body_jntadr', 'pi'), ('body_jntnum', 'pi'),
    ('body_dofadr', 'pi'), ('body_dofnum', 'pi'), ('body_lastdof', 'pi'), ('body_fluid_ellipsoid', 'pi'),
    ('body_pos', 'pd'), ('body_quat', 'pd'), ('body_ipos', 'pd'), ('body_iquat', 'pd'),
    ('body_mass', 'pd'), ('body_inertia', 'pd'), ('body_invweight0', 'pd'), ('body_subtreemass', 'pd'),
    # joints / dofs
    ('jnt_type', 'pi'), ('jnt_qposadr', 'pi'), ('jnt_dofadr', 'pi'), ('jnt_bodyid', 'pi'),
    ('jnt_limited', 'pi'),
    ('jnt_pos', 'pd'), ('jnt_axis', 'pd'), ('jnt_stiffness', 'pd'), ('jnt_range', 'pd'),
    ('jnt_solref', 'pd'), ('jnt_solimp', 'pd'), ('jnt_margin', 'pd'), ('qpos0', 'pd'), ('qpos_spring', 'pd'),
    ('dof_bodyid', 'pi'), ('dof_jntid', 'pi'), ('dof_parentid', 'pi'), ('dof_Madr', 'pi'),
    ('dof_armature', 'pd'), ('dof_damping', 'pd'), ('dof_invweight0', 'pd'),
    # collision geoms
    ('geom_type', 'pi'), ('geom_bodyid', 'pi'), ('geom_condim', 'pi'), ('geom_priority', 'pi'),
    ('geom_size', 'pd'), ('geom_pos', 'pd'), ('geom_quat', 'pd'), ('geom_rbound', 'pd'),
    ('geom_friction', 'pd'), ('geom_solmix', 'pd'), ('geom_solref', 'pd'), ('geom_solimp', 'pd'),
    ('geom_margin', 'pd'), ('geom_gap', 'pd'),
    ('pair_geom1', 'pi'), ('pair_geom2', 'pi'),
    # ellipsoid-fluid geoms
    ('fluid_bodyid', 'pi'), ('fluid_pos', 'pd'), ('fluid_quat', 'pd'), ('fluid_size', 'pd'),
    ('fluid_coef', 'pd'),
    # sites
    ('site_bodyid', 'pi'), ('site_type', 'pi'), ('site_pos', 'pd'), ('site_quat', 'pd'), ('site_size', 'pd'),
    # tendons
    ('tendon_adr', 'pi'), ('tendon_num', 'pi'), ('wrap_dofid', 'pi'), ('wrap_qposadr', 'pi'),
    ('wrap_coef', 'pd'),
    # actuators
    ('actuator_trntype', 'pi'), ('actuator_trnid', 'pi'), ('actuator_dyntype', 'pi'),
    ('actuator_biastype', 'pi'), ('actuator_ctrllimited', 'pi'), ('actuator_forcelimited', 'pi'),
    ('actuator_actadr', 'pi'),
    ('actuator_dynprm', 'pd'), ('actuator_gainprm', 'pd'), ('actuator_biasprm', 'pd'),
    ('actuator_ctrlrange', 'pd'), ('actuator_forcerange', 'pd'),
    # sensors
    ('sensor_type', 'pi'), ('sensor_objid', 'pi'), ('sensor_adr', 'pi'), ('sensor_dim', 'pi'),
]

_CT = {'i': C.c_int32, 'd': C.c_double, 'd3': C.c_double * 3,
       'pi': C.POINTER(C.c_int32), 'pd': C.POINTER(C.c_double)}
_CDECL = {'i': 'int32_t %s;', 'd': 'double %s;', 'd3': 'double %s[3];',
          'pi': 'const int32_t* %s;', 'pd': 'const double* %s;'}


class FbModel(C.Structure):
    _fields_ = [(n, _CT[k]) for n, k in FIELDS]


def c_struct_text():
    lines = ['typedef struct FbModel {']
    lines += ['  ' + _CDECL[k] % n for n, k in FIELDS]
    lines.append('} FbModel;')
    return '\n'.join(lines)


class FlyModel:
    """Host-side model: numpy arrays + names + a live ctypes `FbModel` view."""

    def __init__(self, arrays, meta):
        self.meta = meta
        self.a = {}
        for k, v in arrays.items():
            self.a[k] = v
        for k, v in meta.items():
            if k not in self.a:
                self.a[k] = v
        self._keep = []
        self.c = self._build_c()

    def __getattr__(self, k):
        a = self.__dict__.get('a')
        if a is not None and k in a:
            return a[k]
        raise AttributeError(k)

    def _build_c(self):
        s = FbModel()
        for n, k in FIELDS:
            v = self.a[n]
            if k == 'i':
                setattr(s, n, int(v))
            elif k == 'd':
                setattr(s, n, float(v))
            elif k == 'd3':
                setattr(s, n, (C.c_double * 3)(*[float(x) for x in np.asarray(v).ravel()]))
            elif k == 'pi':
                arr = np.ascontiguousarray(np.asarray(v, dtype=np.int32).ravel())
                if arr.size == 0:
                    arr = np.zeros(1, np.int32)
                self._keep.append(arr)
                setattr(s, n, arr.ctypes.data_as(C.POINTER(C.c_int32)))
            else:
                arr = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
                if arr.size == 0:
                    arr = np.zeros(1, np.float64)
                self._keep.append(arr)
                setattr(s, n, arr.ctypes.data_as(C.POINTER(C.c_double)))
        return s

    def rebuild(self):
        """Re-create the ctypes view after editing arrays in `self.a`."""
        self._keep = []
        self.c = self._build_c()
        return self

    def copy(self):
        import copy as _copy
        return FlyModel({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.a.items()
                         if isinstance(v, np.ndarray)}, _copy.deepcopy(self.meta))

    # name lookups ------------------------------------------------------------------------
    def body_id(self, name):
        return self.meta['body_names'].index(name)

    def jnt_id(self, name):
        return self.meta['jnt_names'].index(name)

    def site_id(self, name):
        return self.meta['site_names'].index(name)

    def jnt_qposadr_of(self, name):
        return int(self.a['jnt_qposadr'][self.jnt_id(name)])

    def jnt_dofadr_of(self, name):
        return int(self.a['jnt_dofadr'][self.jnt_id(name)])


def load_model(variant='walk', path=None):
    path = path or os.path.join(ASSETS, f'fly_{variant}.npz')
    d = np.load(path)
    meta = json.loads(bytes(d['__meta__']).decode())
    arrays = {k: d[k] for k in d.files if k != '__meta__'}
    return FlyModel(arrays, meta)


def from_compiled(m):
    arrays = {k: v for k, v in m.items() if isinstance(v, np.ndarray)}
    meta = {k: v for k, v in m.items() if not isinstance(v, np.ndarray) and not k.startswith('_')}
    return FlyModel(arrays, meta)


# ------------------------------------------------------------------------------------------------ model variants on demand
def reference_assets_dir():
    """Directory holding the reference's `fruitfly.xml` + meshes: $FLYBODY_ASSETS, an installed `flybody` package, or the
    reference checkout of the build container.  None if none is present (only the shipped, pre-compiled variants are available)."""
    cand = [os.environ.get('FLYBODY_ASSETS')]
    try:
        import importlib.util
        spec = importlib.util.find_spec('flybody')
        if spec is not None and spec.submodule_search_locations:
            cand.append(os.path.join(list(spec.submodule_search_locations)[0], 'fruitfly', 'assets'))
    except Exception:
        pass
    cand.append('/root/reference/flybody/fruitfly/assets')
    for c in cand:
        if c and os.path.exists(os.path.join(c, 'fruitfly.xml')):
            return c
    return None


def variant_name(variant, force_actuators=False, use_wings=None, use_legs=None, joint_filter=None):
    name = f'fly_{variant}'
    if force_actuators:
        name += '_force'
    if use_wings is not None:
        name += '_wings' if use_wings else '_nowings'
    if use_legs is not None:
        name += '_legs' if use_legs else '_nolegs'
    if joint_filter is not None:
        name += f'_jf{joint_filter:g}'
    return name


def model_for(variant, force_actuators=False, use_wings=None, use_legs=None, joint_filter=None, cache_dir=None):
    """The compiled model of a task (`variant` = walk / flight / vision) with the `FruitFly` switches the reference's env factories
    expose (reference fly_envs.py:100-246: force_actuators, disable_wings, disable_legs, joint_filter; None = the task's default).
    The default models ship pre-compiled (`assets/fly_*.npz`); any other combination is compiled from the reference's
    `fruitfly.xml` on first use (`flybody_b200.compiler`, ~2 s) and cached."""
    if not force_actuators and use_wings is None and use_legs is None and joint_filter is None:
        return load_model(variant)
    name = variant_name(variant, force_actuators, use_wings, use_legs, joint_filter)
    cache_dir = cache_dir or os.environ.get('FLYBODY_B200_CACHE') or os.path.join(os.path.expanduser('~'), '.cache', 'flybody_b200')
    for d in (ASSETS, cache_dir):
        p = os.path.join(d, name + '.npz')
        if os.path.exists(p):
            return load_model(path=p)
    src = reference_assets_dir()
    if src is None:
        raise NotImplementedError(
            f'model variant {name} is not among the pre-compiled ones ({sorted(f for f in os.listdir(ASSETS) if f.endswith(".npz"))}) and the '
            "reference's fruitfly.xml was not found to compile it: install flybody or set FLYBODY_ASSETS to its fruitfly/assets directory")
    from .compiler import compile_model as cm
    m = cm.compile_variant(variant, assets_dir=src, mesh_cache=cm.load_mesh_cache(src), force_actuators=force_actuators, use_wings=use_wings,
                           use_legs=use_legs, joint_filter=joint_filter)
    os.makedirs(cache_dir, exist_ok=True)
    p = os.path.join(cache_dir, name + '.npz')
    cm.save_model(m, p)
    return load_model(path=p)
