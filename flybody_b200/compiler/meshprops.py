"""Volume / centre of mass / inertia of the fly's OBJ meshes.

The reference model gives most bodies their mass through `density x mesh volume`
(`fruitfly.xml:14,37,80,94`; goldens `tests/test_flybare.py:27-36`).  The OBJ files are
un-indexed triangle soups (`v/vt/vn`, `f a/a/a b/b/b c/c/c`).

Two volume algorithms are provided because the MuJoCo compiler has had both
(SURVEY.md App. D.2); `compile_model` selects the one that reproduces the golden masses:

* ``exact``  : signed tetrahedra against the origin (exact for closed, consistently
               oriented surfaces);
* ``legacy`` : |signed| tetrahedra against the area-weighted surface centroid (exact only
               for convex meshes);
* ``legacy2``: CoM from the ``legacy`` pass, then volume and inertia from |signed| tetrahedra
               against that CoM.  THIS is the variant that reproduces the reference's golden
               masses (total 2e-9, head 1e-8, abdomen 3e-9 relative; `tests/test_flybare.py:27-36`)
               and is therefore the default.

Results are for unit density, in the (already scaled) mesh frame, and are cached in
``assets/mesh_props.json`` so that neither the GPU box nor the tests need the 154 MB of OBJ.
"""
import json
import os

import numpy as np


def load_obj_triangles(path):
    """Return (T,3,3) float64 triangle vertex array from a triangle-soup OBJ."""
    verts = []
    faces = []
    with open(path, 'r') as f:
        for line in f:
            if line.startswith('v '):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith('f '):
                p = line.split()[1:]
                idx = [int(t.split('/')[0]) for t in p]
                # fan-triangulate (faces are triangles in this asset set)
                for k in range(1, len(idx) - 1):
                    faces.append((idx[0], idx[k], idx[k + 1]))
    v = np.asarray(verts, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    f = np.where(f > 0, f - 1, f + len(v))
    return v[f]


def _tet_integrals(a, b, c, vol6):
    """Sum over tetrahedra (origin,a,b,c) of volume, first and second moments.

    vol6: (T,) signed (or absolute) 6*volume per tetrahedron."""
    vol = vol6 / 6.0
    V = vol.sum()
    # first moment: centroid of tet = (a+b+c)/4
    m1 = ((a + b + c) / 4.0 * vol[:, None]).sum(0)
    # second moments: integral of x_i x_j over tet (origin,a,b,c) = vol/20 * (sum_k p_k p_k^T + (sum p)(sum p)^T)
    s = a + b + c
    outer = (np.einsum('ti,tj->tij', a, a) + np.einsum('ti,tj->tij', b, b)
             + np.einsum('ti,tj->tij', c, c) + np.einsum('ti,tj->tij', s, s))
    m2 = (outer * (vol / 20.0)[:, None, None]).sum(0)
    return V, m1, m2


def mesh_props(tri, mode):
    """Unit-density (volume, com[3], inertia 3x3 about com) of triangle array `tri` (T,3,3)."""
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    if mode == 'exact':
        origin = np.zeros(3)
        vol6 = np.einsum('ti,ti->t', a, np.cross(b, c))
        V, m1, m2 = _tet_integrals(a, b, c, vol6)
    elif mode == 'legacy':
        n = np.cross(b - a, c - a)
        area = 0.5 * np.linalg.norm(n, axis=1)
        origin = (((a + b + c) / 3.0) * area[:, None]).sum(0) / area.sum()
        a, b, c = a - origin, b - origin, c - origin
        vol6 = np.abs(np.einsum('ti,ti->t', a, np.cross(b, c)))
        V, m1, m2 = _tet_integrals(a, b, c, vol6)
    elif mode == 'legacy2':
        # |signed| tetrahedra against the centre of mass found by the `legacy` pass
        _, origin, _ = mesh_props(tri, 'legacy')
        a, b, c = a - origin, b - origin, c - origin
        vol6 = np.abs(np.einsum('ti,ti->t', a, np.cross(b, c)))
        V, m1, m2 = _tet_integrals(a, b, c, vol6)
        # the pass-1 centre is kept as the mesh CoM and the inertia is taken about it
        I0 = np.trace(m2) * np.eye(3) - m2
        return float(V), origin, I0
    else:
        raise ValueError(mode)
    com_rel = m1 / V
    # inertia about the integration origin, then shift to com
    I0 = np.trace(m2) * np.eye(3) - m2
    Ic = I0 - V * (np.dot(com_rel, com_rel) * np.eye(3) - np.outer(com_rel, com_rel))
    return float(V), com_rel + origin, Ic


def build_cache(assets_dir, mesh_files, scale, out_path):
    """mesh_files: {mesh_name: file}. Writes {name: {mode: {volume, com, inertia}}}."""
    cache = {}
    for name, fn in sorted(mesh_files.items()):
        tri = load_obj_triangles(os.path.join(assets_dir, fn)) * np.asarray(scale)[None, None, :]
        cache[name] = {}
        for mode in ('exact', 'legacy', 'legacy2'):
            V, com, I = mesh_props(tri, mode)
            cache[name][mode] = {'volume': V, 'com': com.tolist(), 'inertia': I.tolist()}
    with open(out_path, 'w') as f:
        json.dump(cache, f, indent=0)
    return cache
