// fb_constraint.h -- collision, constraint rows, Delassus projection, actuation, the dual-space
// Newton solver with elliptic cones + noslip, acceleration-stage sensors and the Euler integrator.
#pragma once
#include "fb_tree.h"

// contact record: con_pos[3], con_frame[9] per slot (positions relative to ref)
#define CON_F(arr, slot, k, K) AT(arr, (slot) * (K) + (k))
#define EFC(arr, r) AT(arr, r)
#define EJ(arr, r, dof) AT(arr, (size_t)(r) * m.nv + (dof))
#define EA(arr, r, c) AT(arr, (size_t)(r) * FB_MAXEFC + (c))
#define EW(slot, r) AT(d.efc_w, (slot) * FB_MAXEFC + (r))

// ---------------------------------------------------------------------------------------------
// K9 collision: blockDim = (32 envs, nchunk).  Each thread scans a contiguous chunk of the
// candidate pair list (bounding-sphere broadphase, then the analytic narrowphase) into a private
// staging area; a prefix sum over the chunk counts makes the final contact order deterministic
// (pair-list order), which the Gauss-Seidel noslip sweeps depend on.
struct ShCol { int cnt[FB_MAXCHUNK][32]; };
struct RawCon { float dist; V3 pos, n, t; };

FB_DEV int raw_sphere_sphere(RawCon* c, float margin, V3 p1, float r1, V3 p2, float r2) {
  V3 dif = p2 - p1; float cd = dot(dif, dif), lim = margin + r1 + r2;
  if (cd > lim * lim) return 0;
  float len = sqrtf(cd);
  c->dist = len - r1 - r2;
  c->n = (len < 1e-20f) ? v3(1, 0, 0) : dif * (1.0f / len);
  c->pos = p1 + c->n * (r1 + 0.5f * c->dist);
  c->t = v3(0, 0, 0);
  return 1;
}
FB_DEV int raw_plane_sphere(RawCon* c, float margin, V3 pp, V3 n, V3 sp, float r) {
  float cd = dot(sp - pp, n);
  if (cd > margin + r) return 0;
  c->dist = cd - r; c->n = n; c->pos = sp - n * (r + 0.5f * c->dist); c->t = v3(0, 0, 0);
  return 1;
}
FB_DEV int col_plane_capsule(RawCon* c, float margin, V3 pp, V3 n, V3 cp, const M3& cm, V3 size) {
  V3 axis = col(cm, 2), seg = axis * size.y; int k = 0;
  k += raw_plane_sphere(c + k, margin, pp, n, cp + seg, size.x);
  k += raw_plane_sphere(c + k, margin, pp, n, cp - seg, size.x);
  for (int i = 0; i < k; i++) c[i].t = axis;
  return k;
}
FB_DEV int col_plane_cylinder(RawCon* c, float margin, V3 pp, V3 n, V3 cp, const M3& cm, V3 size) {
  V3 axis = col(cm, 2);
  float dist0 = dot(cp - pp, n), prjaxis = dot(n, axis);
  if (prjaxis > 0) { axis = axis * -1.0f; prjaxis = -prjaxis; }
  V3 vec = axis * prjaxis - n;
  float len2 = dot(vec, vec);
  if (len2 >= 1e-12f) vec = vec * (size.x / sqrtf(len2)); else vec = col(cm, 0) * size.x;
  float prjvec = dot(vec, n);
  axis = axis * size.y; prjaxis *= size.y;
  int cnt = 0;
  if (dist0 + prjaxis + prjvec <= margin) {
    c[cnt].dist = dist0 + prjaxis + prjvec; c[cnt].pos = cp + vec + axis - n * (0.5f * c[cnt].dist); cnt++;
  } else return 0;
  if (dist0 - prjaxis + prjvec <= margin) {
    c[cnt].dist = dist0 - prjaxis + prjvec; c[cnt].pos = cp + vec - axis - n * (0.5f * c[cnt].dist); cnt++;
  }
  float prjvec1 = -prjvec * 0.5f;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    V3 vec1 = normalized(cross(vec, axis)) * (size.x * 0.8660254038f);
    for (int s = 0; s < 2; s++) {
      c[cnt].dist = dist0 + prjaxis + prjvec1;
      c[cnt].pos = cp + axis + vec1 * (s ? -1.0f : 1.0f) - vec * 0.5f - n * (0.5f * c[cnt].dist);
      cnt++;
    }
  }
  for (int i = 0; i < cnt; i++) { c[i].n = n; c[i].t = v3(0, 0, 0); }
  return cnt;
}
FB_DEV int col_plane_ellipsoid(RawCon* c, float margin, V3 pp, V3 n, V3 ep, const M3& em, V3 size) {
  V3 ln = mulT(em, n);
  V3 s = v3(-size.x * size.x * ln.x, -size.y * size.y * ln.y, -size.z * size.z * ln.z);
  float den = sqrtf(size.x * size.x * ln.x * ln.x + size.y * size.y * ln.y * ln.y + size.z * size.z * ln.z * ln.z);
  s = s * (1.0f / fmaxf(den, 1e-30f));
  V3 sup = mul(em, s) + ep;
  float dist = dot(sup - pp, n);
  if (dist > margin) return 0;
  c->dist = dist; c->n = n; c->pos = sup - n * (0.5f * dist); c->t = v3(0, 0, 0);
  return 1;
}
FB_DEV int col_sphere_capsule(RawCon* c, float margin, V3 sp, float sr, V3 cp, const M3& cm, V3 csize) {
  V3 axis = col(cm, 2);
  float x = clampf(dot(axis, sp - cp), -csize.y, csize.y);
  return raw_sphere_sphere(c, margin, sp, sr, cp + axis * x, csize.x);
}
FB_DEV int col_capsule_capsule(RawCon* c, float margin, V3 p1, const M3& m1, V3 s1, V3 p2, const M3& m2, V3 s2) {
  V3 a1 = col(m1, 2), a2 = col(m2, 2), dif = p1 - p2;
  float ma = dot(a1, a1), mb = -dot(a1, a2), mc = dot(a2, a2), u = -dot(a1, dif), v = dot(a2, dif);
  float det = ma * mc - mb * mb;
  // fp32: 1 - cos^2 loses its digits long before the fp64 threshold (1e-15) of the reference
  if (fabsf(det) >= 1e-6f) {
    float x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > s1.y) { x1 = s1.y; x2 = (v - mb * s1.y) / mc; }
    else if (x1 < -s1.y) { x1 = -s1.y; x2 = (v + mb * s1.y) / mc; }
    if (x2 > s2.y) { x2 = s2.y; x1 = clampf((u - mb * s2.y) / ma, -s1.y, s1.y); }
    else if (x2 < -s2.y) { x2 = -s2.y; x1 = clampf((u + mb * s2.y) / ma, -s1.y, s1.y); }
    return raw_sphere_sphere(c, margin, p1 + a1 * x1, s1.x, p2 + a2 * x2, s2.x);
  }
  int k = 0; float x;
  x = clampf((v - mb * s1.y) / mc, -s2.y, s2.y); k += raw_sphere_sphere(c + k, margin, p1 + a1 * s1.y, s1.x, p2 + a2 * x, s2.x);
  x = clampf((v + mb * s1.y) / mc, -s2.y, s2.y); k += raw_sphere_sphere(c + k, margin, p1 - a1 * s1.y, s1.x, p2 + a2 * x, s2.x);
  if (k >= 2) return k;
  x = clampf((u - mb * s2.y) / ma, -s1.y, s1.y); k += raw_sphere_sphere(c + k, margin, p1 + a1 * x, s1.x, p2 + a2 * s2.y, s2.x);
  if (k >= 2) return k;
  x = clampf((u + mb * s2.y) / ma, -s1.y, s1.y); k += raw_sphere_sphere(c + k, margin, p1 + a1 * x, s1.x, p2 - a2 * s2.y, s2.x);
  return k;
}
FB_DEV void make_frame(V3 n, V3 t, V3& f1, V3& f2) {   // mju_makeFrame
  if (norm(t) < 0.5f) { t = (n.y < 0.5f && n.y > -0.5f) ? v3(0, 1, 0) : v3(0, 0, 1); }
  t = normalized(t - n * dot(n, t));
  f1 = t; f2 = cross(n, t);
}

#define FB_COL_ARGS const DevModel& m, const DevData& d, ShCol& sh, int e, int lane, int y
FB_DEV void kcol_p0(FB_COL_ARGS) {
  int cnt = 0;
  int p0 = m.chunk_start[y], p1 = m.chunk_start[y + 1];
  V3 ref = v3(AT(d.ref, 0), AT(d.ref, 1), AT(d.ref, 2));
  (void)ref;
  for (int k = p0; k < p1; k++) {
    int g1 = m.pair_geom1[k], g2 = m.pair_geom2[k];
    int t1 = m.geom_type[g1], t2 = m.geom_type[g2];
    float margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]);
    V3 x1 = ld3(d.geom_xpos, g1, d, e), x2 = ld3(d.geom_xpos, g2, d, e);
    RawCon rc[4]; int n = 0;
    if (t1 == FB_GEOM_PLANE) {
      V3 pn = v3(AT(d.geom_xmat, 9 * g1 + 2), AT(d.geom_xmat, 9 * g1 + 5), AT(d.geom_xmat, 9 * g1 + 8));
      if (dot(x2 - x1, pn) > margin + m.geom_rbound[g2]) continue;
      V3 s2 = mld3(m.geom_size, g2);
      if (t2 == FB_GEOM_SPHERE) n = raw_plane_sphere(rc, margin, x1, pn, x2, s2.x);
      else { M3 R2 = ld9(d.geom_xmat, g2, d, e);
        if (t2 == FB_GEOM_CAPSULE) n = col_plane_capsule(rc, margin, x1, pn, x2, R2, s2);
        else if (t2 == FB_GEOM_CYLINDER) n = col_plane_cylinder(rc, margin, x1, pn, x2, R2, s2);
        else if (t2 == FB_GEOM_ELLIPSOID) n = col_plane_ellipsoid(rc, margin, x1, pn, x2, R2, s2); }
    } else {
      V3 df = x2 - x1; float r = margin + m.geom_rbound[g1] + m.geom_rbound[g2];
      if (dot(df, df) > r * r) continue;
      V3 s1 = mld3(m.geom_size, g1), s2 = mld3(m.geom_size, g2);
      if (t1 == FB_GEOM_SPHERE && t2 == FB_GEOM_SPHERE) n = raw_sphere_sphere(rc, margin, x1, s1.x, x2, s2.x);
      else if (t1 == FB_GEOM_SPHERE && t2 == FB_GEOM_CAPSULE) { M3 R2 = ld9(d.geom_xmat, g2, d, e); n = col_sphere_capsule(rc, margin, x1, s1.x, x2, R2, s2); }
      else if (t1 == FB_GEOM_CAPSULE && t2 == FB_GEOM_CAPSULE) { M3 R1 = ld9(d.geom_xmat, g1, d, e), R2 = ld9(d.geom_xmat, g2, d, e); n = col_capsule_capsule(rc, margin, x1, R1, s1, x2, R2, s2); }
      else n = 0;    // generic convex pairs (ellipsoid / cylinder vs non-plane): next row, DESIGN.md
    }
    for (int i = 0; i < n; i++) {
      if (cnt >= FB_CHUNKCAP) { FB_FLAG_OR(2); break; }
      int slot = y * FB_CHUNKCAP + cnt;
      V3 f1, f2; make_frame(rc[i].n, rc[i].t, f1, f2);
      CON_F(d.tmp_con, slot, 0, 13) = rc[i].dist;
      CON_F(d.tmp_con, slot, 1, 13) = rc[i].pos.x; CON_F(d.tmp_con, slot, 2, 13) = rc[i].pos.y; CON_F(d.tmp_con, slot, 3, 13) = rc[i].pos.z;
      CON_F(d.tmp_con, slot, 4, 13) = rc[i].n.x; CON_F(d.tmp_con, slot, 5, 13) = rc[i].n.y; CON_F(d.tmp_con, slot, 6, 13) = rc[i].n.z;
      CON_F(d.tmp_con, slot, 7, 13) = f1.x; CON_F(d.tmp_con, slot, 8, 13) = f1.y; CON_F(d.tmp_con, slot, 9, 13) = f1.z;
      CON_F(d.tmp_con, slot, 10, 13) = f2.x; CON_F(d.tmp_con, slot, 11, 13) = f2.y; CON_F(d.tmp_con, slot, 12, 13) = f2.z;
      AT(d.tmp_geom, 2 * slot) = g1; AT(d.tmp_geom, 2 * slot + 1) = g2;
      cnt++;
    }
  }
  sh.cnt[y][lane] = cnt;
}
FB_DEV void kcol_p1(FB_COL_ARGS) {
  int off = 0;
  for (int yy = 0; yy < y; yy++) off += sh.cnt[yy][lane];
  int cnt = sh.cnt[y][lane];
  for (int i = 0; i < cnt; i++) {
    int dst = off + i, src = y * FB_CHUNKCAP + i;
    if (dst >= FB_MAXCON) { FB_FLAG_OR(2); break; }
    AT(d.con_dist, dst) = CON_F(d.tmp_con, src, 0, 13);
    for (int k = 0; k < 3; k++) CON_F(d.con_pos, dst, k, 3) = CON_F(d.tmp_con, src, 1 + k, 13);
    for (int k = 0; k < 9; k++) CON_F(d.con_frame, dst, k, 9) = CON_F(d.tmp_con, src, 4 + k, 13);
    AT(d.con_geom1, dst) = AT(d.tmp_geom, 2 * src); AT(d.con_geom2, dst) = AT(d.tmp_geom, 2 * src + 1);
  }
  if (y == m.nchunk - 1) { int tot = off + cnt; AT(d.ncon, 0) = tot > FB_MAXCON ? FB_MAXCON : tot; }
}

// ---------------------------------------------------------------------------------------------
// K10 constraint rows (MuJoCo mj_makeConstraint + mj_makeImpedance), lane = env
FB_DEV float impedance(const float* si, float pos, float margin) {
  if (si[0] == si[1] || si[2] <= FB_MINVAL) return 0.5f * (si[0] + si[1]);
  float x = fabsf((pos - margin) / si[2]);
  if (x >= 1 || x <= 0) return x >= 1 ? si[1] : si[0];
  float yv;
  if (si[4] == 1) yv = x;
  else if (x <= si[3]) yv = powf(x, si[4]) / powf(si[3], si[4] - 1);
  else yv = 1 - powf(1 - x, si[4]) / powf(1 - si[3], si[4] - 1);
  return si[0] + yv * (si[1] - si[0]);
}
FB_DEV void row_params(const DevModel& m, const DevData& d, int e, int r, const float* sr_in, const float* si, float pos, float margin,
                       float diagApprox, bool isfric) {
  float sr0 = sr_in[0], sr1 = sr_in[1];
  if (sr0 > 0 && sr0 < 2 * m.timestep) sr0 = 2 * m.timestep;    // refsafe
  float imp = clampf(impedance(si, pos, margin), 0.0001f, 0.9999f);
  float R = fmaxf(FB_MINVAL, (1 - imp) * diagApprox / imp);
  float K, B, dmax = si[1];
  if (sr0 > 0) { K = 1 / fmaxf(FB_MINVAL, dmax * dmax * sr0 * sr0 * sr1 * sr1); B = 2 / fmaxf(FB_MINVAL, dmax * sr0); }
  else { K = -sr0 / fmaxf(FB_MINVAL, dmax * dmax); B = -sr1 / fmaxf(FB_MINVAL, dmax); }
  if (isfric) K = 0;
  EFC(d.efc_R, r) = R; EFC(d.efc_D, r) = 1.0f / R; EFC(d.efc_K, r) = K; EFC(d.efc_B, r) = B; EFC(d.efc_imp, r) = imp;
}
FB_DEV void kcon(const DevModel& m, const DevData& d, int e) {
  int n = 0;
  for (int j = 0; j < m.njnt; j++) {
    if (!m.jnt_limited[j] || m.jnt_type[j] != FB_JNT_HINGE) continue;
    float value = AT(d.qpos, m.jnt_qposadr[j]), margin = m.jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      float dist = side * (m.jnt_range[2 * j + (side + 1) / 2] - value);
      if (dist < margin && n < FB_MAXEFC) {
        EFC(d.efc_type, n) = FB_CT_LIMIT; EFC(d.efc_id, n) = (side < 0) ? j : -(j + 1);   // sign encodes the side
        EFC(d.efc_pos, n) = dist; EFC(d.efc_margin, n) = margin;
        row_params(m, d, e, n, m.jnt_solref + 2 * j, m.jnt_solimp + 5 * j, dist, margin, m.dof_invweight0[m.jnt_dofadr[j]], false);
        n++;
      }
    }
  }
  int ncon = AT(d.ncon, 0);
  for (int ci = 0; ci < ncon; ci++) {
    int g1 = AT(d.con_geom1, ci), g2 = AT(d.con_geom2, ci);
    float margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]), gap = fmaxf(m.geom_gap[g1], m.geom_gap[g2]);
    float includemargin = margin - gap, dist = AT(d.con_dist, ci);
    int dim = m.geom_condim[g1] > m.geom_condim[g2] ? m.geom_condim[g1] : m.geom_condim[g2];
    dim = dim >= 3 ? 3 : 1;
    float fr0 = fmaxf(m.geom_friction[3 * g1], m.geom_friction[3 * g2]);
    AT(d.con_dim, ci) = dim; AT(d.con_mu, ci) = fr0 * sqrtf(1.0f / m.impratio);
    CON_F(d.con_fric, ci, 0, 2) = fr0; CON_F(d.con_fric, ci, 1, 2) = fr0;
    AT(d.con_efcadr, ci) = -1;
    if (!(dist < includemargin)) continue;          // detected, but inside the gap: adhesion only
    if (n + dim > FB_MAXEFC) { FB_FLAG_OR(4); continue; }
    float mix1 = m.geom_solmix[g1], mix2 = m.geom_solmix[g2], mix;
    if (mix1 >= FB_MINVAL && mix2 >= FB_MINVAL) mix = mix1 / (mix1 + mix2);
    else if (mix1 < FB_MINVAL && mix2 < FB_MINVAL) mix = 0.5f; else mix = (mix1 < FB_MINVAL) ? 0.0f : 1.0f;
    float sr[2], si[5];
    if (m.geom_solref[2 * g1] > 0 && m.geom_solref[2 * g2] > 0) for (int q = 0; q < 2; q++) sr[q] = mix * m.geom_solref[2 * g1 + q] + (1 - mix) * m.geom_solref[2 * g2 + q];
    else for (int q = 0; q < 2; q++) sr[q] = fminf(m.geom_solref[2 * g1 + q], m.geom_solref[2 * g2 + q]);
    for (int q = 0; q < 5; q++) si[q] = mix * m.geom_solimp[5 * g1 + q] + (1 - mix) * m.geom_solimp[5 * g2 + q];
    int b1 = m.geom_bodyid[g1], b2 = m.geom_bodyid[g2];
    float tran = m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2];
    AT(d.con_efcadr, ci) = n;
    for (int r = 0; r < dim; r++) {
      EFC(d.efc_type, n + r) = dim == 1 ? FB_CT_FRICTIONLESS : FB_CT_ELLIPTIC; EFC(d.efc_id, n + r) = ci;
      EFC(d.efc_pos, n + r) = r == 0 ? dist : 0.0f; EFC(d.efc_margin, n + r) = r == 0 ? includemargin : 0.0f;
      row_params(m, d, e, n + r, sr, si, dist, includemargin, tran, r > 0);
    }
    if (dim == 3) {   // friction rows: R_t = R_n / impratio (both tangents share mu here)
      float Rn = EFC(d.efc_R, n), Rt = Rn / fmaxf(FB_MINVAL, m.impratio);
      EFC(d.efc_R, n + 1) = Rt; EFC(d.efc_D, n + 1) = 1.0f / Rt; EFC(d.efc_R, n + 2) = Rt; EFC(d.efc_D, n + 2) = 1.0f / Rt;
    }
    n += dim;
  }
  AT(d.nefc, 0) = n;
}

// ---------------------------------------------------------------------------------------------
// Row dof sets.  A limit row touches the ancestor chain of its dof; a contact row the union of the
// chains of its two bodies.  `RowDofs` walks that union in descending dof order.
struct RowChains { int la, lb; };
FB_DEV RowChains row_chains(const DevModel& m, const DevData& d, int e, int r, int& ci, int& frow, float& sign) {
  RowChains rc; int tp = EFC(d.efc_type, r), id = EFC(d.efc_id, r);
  if (tp == FB_CT_LIMIT) {
    int j = id >= 0 ? id : -(id + 1);
    sign = id >= 0 ? 1.0f : -1.0f;        // lower limit: J = +1, upper: J = -1
    rc.la = m.jnt_dofadr[j]; rc.lb = -1; ci = -1; frow = 0;
  } else {
    ci = id; frow = r - AT(d.con_efcadr, ci); sign = 1.0f;
    rc.la = m.body_lastdof[m.geom_bodyid[AT(d.con_geom2, ci)]];
    rc.lb = m.body_lastdof[m.geom_bodyid[AT(d.con_geom1, ci)]];
  }
  return rc;
}
FB_DEV bool in_chain(const DevModel& m, int dof, int last) { return last >= 0 && dof <= last && last <= m.dof_subend[dof]; }

// Jacobian entry of contact `ci`, frame row `frow`, on dof k of chain side s (+1 body2, -1 body1)
FB_DEV float contact_J(const DevModel& m, const DevData& d, int e, V3 f, V3 pos, int k) {
  V3 t = ld3(d.Slin, k, d, e) + cross(ld3(d.Sang, k, d, e), pos);
  return dot(f, t);
}

// K10b projection: blockDim = (32, FB_ROWPAR): rows are distributed over threadIdx.y.
//   phase 0: J (dense-by-dof storage), Z = D^-1/2 L^-T J^T restricted to the row's dof set
//   phase 1: A = Z Z^T  (= J M^-1 J^T, the unregularised Delassus matrix)
struct ShNone { int dummy; };
#define FB_ROW_ARGS const DevModel& m, const DevData& d, ShNone& sh, int e, int lane, int y
FB_DEV void kproj_p0(FB_ROW_ARGS) {
  int n = AT(d.nefc, 0);
  for (int r = y; r < n; r += FB_ROWPAR) {
    int ci, frow; float sign;
    RowChains rc = row_chains(m, d, e, r, ci, frow, sign);
    V3 f = v3(0, 0, 0), pos = v3(0, 0, 0);
    if (ci >= 0) { f = v3(CON_F(d.con_frame, ci, 3 * frow, 9), CON_F(d.con_frame, ci, 3 * frow + 1, 9), CON_F(d.con_frame, ci, 3 * frow + 2, 9));
                   pos = v3(CON_F(d.con_pos, ci, 0, 3), CON_F(d.con_pos, ci, 1, 3), CON_F(d.con_pos, ci, 2, 3)); }
    // J over the dof set
    int la = rc.la, lb = rc.lb;
    while (la >= 0 || lb >= 0) {
      int k = la > lb ? la : lb; float v = 0;
      if (ci < 0) v = (k == rc.la) ? sign : 0.0f;
      else { if (la == k) v += contact_J(m, d, e, f, pos, k); if (lb == k) v -= contact_J(m, d, e, f, pos, k); }
      EJ(d.efc_J, r, k) = v; EJ(d.efc_Z, r, k) = v;
      if (la == k) la = m.dof_parentid[la];
      if (lb == k) lb = m.dof_parentid[lb];
    }
    // Z <- L^-T sweep (descending dofs), then scale by D^-1/2
    la = rc.la; lb = rc.lb;
    while (la >= 0 || lb >= 0) {
      int k = la > lb ? la : lb;
      if (la == k) la = m.dof_parentid[la];
      if (lb == k) lb = m.dof_parentid[lb];
      float zk = EJ(d.efc_Z, r, k);
      int adrk = m.dof_Madr[k], t = 1;
      for (int i = m.dof_parentid[k]; i >= 0; i = m.dof_parentid[i], t++) EJ(d.efc_Z, r, i) -= AT(d.qLD, adrk + t) * zk;
      EJ(d.efc_Z, r, k) = zk / sqrtf(AT(d.qLD, adrk));
    }
  }
}
FB_DEV void kproj_p1(FB_ROW_ARGS) {
  int n = AT(d.nefc, 0);
  for (int r = y; r < n; r += FB_ROWPAR) {
    int ci, frow; float sign;
    RowChains rr = row_chains(m, d, e, r, ci, frow, sign);
    for (int c = 0; c <= r; c++) {
      RowChains rc = row_chains(m, d, e, c, ci, frow, sign);
      float s = 0; int la = rc.la, lb = rc.lb;
      while (la >= 0 || lb >= 0) {
        int k = la > lb ? la : lb;
        if (la == k) la = m.dof_parentid[la];
        if (lb == k) lb = m.dof_parentid[lb];
        if (in_chain(m, k, rr.la) || in_chain(m, k, rr.lb)) s += EJ(d.efc_Z, r, k) * EJ(d.efc_Z, c, k);
      }
      EA(d.efc_A, r, c) = s; EA(d.efc_A, c, r) = s;
    }
  }
}
// J . x for every row (x: qvel, qacc_smooth, qacc_warmstart): aref, b, jar at the warm start
FB_DEV void kref(FB_ROW_ARGS) {
  int n = AT(d.nefc, 0);
  for (int r = y; r < n; r += FB_ROWPAR) {
    int ci, frow; float sign;
    RowChains rc = row_chains(m, d, e, r, ci, frow, sign);
    float vel = 0, as = 0, ws = 0; int la = rc.la, lb = rc.lb;
    while (la >= 0 || lb >= 0) {
      int k = la > lb ? la : lb;
      if (la == k) la = m.dof_parentid[la];
      if (lb == k) lb = m.dof_parentid[lb];
      float J = EJ(d.efc_J, r, k);
      vel += J * AT(d.qvel, k); as += J * AT(d.qacc_smooth, k); ws += J * AT(d.qacc_warmstart, k);
    }
    float aref = -EFC(d.efc_B, r) * vel - EFC(d.efc_K, r) * EFC(d.efc_imp, r) * (EFC(d.efc_pos, r) - EFC(d.efc_margin, r));
    EFC(d.efc_aref, r) = aref; EFC(d.efc_b, r) = as - aref; EFC(d.efc_jarws, r) = ws - aref;
  }
}

// ---------------------------------------------------------------------------------------------
// K3 + K8 transmission and actuation (MuJoCo mj_transmission, mj_fwdActuation), lane = env
FB_DEV void kact(const DevModel& m, const DevData& d, int e) {
  for (int k = 0; k < m.nv; k++) AT(d.qfrc_actuator, k) = 0;
  int ncon = AT(d.ncon, 0);
  for (int i = 0; i < m.nu; i++) {
    float ctrl = AT(d.ctrl, i);
    if (m.actuator_ctrllimited[i]) ctrl = clampf(ctrl, m.actuator_ctrlrange[2 * i], m.actuator_ctrlrange[2 * i + 1]);
    int id = m.actuator_trnid[i], tt = m.actuator_trntype[i];
    float len = 0, vel = 0;
    if (tt == FB_TRN_JOINT) { len = AT(d.qpos, m.jnt_qposadr[id]); vel = AT(d.qvel, m.jnt_dofadr[id]); }
    else if (tt == FB_TRN_TENDON) {
      for (int w = m.tendon_adr[id]; w < m.tendon_adr[id] + m.tendon_num[id]; w++) { len += m.wrap_coef[w] * AT(d.qpos, m.wrap_qposadr[w]); vel += m.wrap_coef[w] * AT(d.qvel, m.wrap_dofid[w]); }
    }
    float input = ctrl; int aa = m.actuator_actadr[i];
    if (aa >= 0) { AT(d.act_dot, aa) = (ctrl - AT(d.act, aa)) / fmaxf(FB_MINVAL, m.actuator_dynprm[3 * i]); input = AT(d.act, aa); }
    float force = m.actuator_gainprm[3 * i] * input;
    if (m.actuator_biastype[i] == 1) force += m.actuator_biasprm[3 * i] + m.actuator_biasprm[3 * i + 1] * len + m.actuator_biasprm[3 * i + 2] * vel;
    if (m.actuator_forcelimited[i]) force = clampf(force, m.actuator_forcerange[2 * i], m.actuator_forcerange[2 * i + 1]);
    AT(d.actuator_force, i) = force;
    if (tt == FB_TRN_JOINT) AT(d.qfrc_actuator, m.jnt_dofadr[id]) += force;
    else if (tt == FB_TRN_TENDON) { for (int w = m.tendon_adr[id]; w < m.tendon_adr[id] + m.tendon_num[id]; w++) AT(d.qfrc_actuator, m.wrap_dofid[w]) += m.wrap_coef[w] * force; }
    else {
      // adhesion: moment = - mean of the contact-normal Jacobians of all detected contacts of the body
      int cnt = 0;
      for (int ci = 0; ci < ncon; ci++) { if (m.geom_bodyid[AT(d.con_geom1, ci)] == id || m.geom_bodyid[AT(d.con_geom2, ci)] == id) cnt++; }
      if (cnt == 0 || force == 0.0f) continue;
      float sc = -force / cnt;
      for (int ci = 0; ci < ncon; ci++) {
        int b1 = m.geom_bodyid[AT(d.con_geom1, ci)], b2 = m.geom_bodyid[AT(d.con_geom2, ci)];
        if (b1 != id && b2 != id) continue;
        V3 f = v3(CON_F(d.con_frame, ci, 0, 9), CON_F(d.con_frame, ci, 1, 9), CON_F(d.con_frame, ci, 2, 9));
        V3 pos = v3(CON_F(d.con_pos, ci, 0, 3), CON_F(d.con_pos, ci, 1, 3), CON_F(d.con_pos, ci, 2, 3));
        for (int k = m.body_lastdof[b2]; k >= 0; k = m.dof_parentid[k]) AT(d.qfrc_actuator, k) += sc * contact_J(m, d, e, f, pos, k);
        for (int k = m.body_lastdof[b1]; k >= 0; k = m.dof_parentid[k]) AT(d.qfrc_actuator, k) -= sc * contact_J(m, d, e, f, pos, k);
      }
    }
  }
  for (int k = 0; k < m.nv; k++) {
    float s = AT(d.qfrc_passive, k) - AT(d.qfrc_bias, k) + AT(d.qfrc_actuator, k);
    AT(d.qfrc_smooth, k) = s; AT(d.qacc_smooth, k) = s;
  }
}

// ---------------------------------------------------------------------------------------------
// K11 constraint solve in the dual (force) space, lane = env.
//
// With a = qacc_smooth + M^-1 J^T lam the primal cost of MuJoCo's Newton solver becomes
//   c(lam) = 1/2 lam^T A lam + s(b + A lam),   A = J M^-1 J^T, b = J qacc_smooth - aref,
// s = the same per-row cost (half-quadratic for limits / frictionless contacts, three-zone elliptic
// cone for frictional contacts).  Newton steps on the primal problem map to
//   dlam = -(I + C A)^-1 (lam - f(lam)),   C = Hessian of s = E E^T,
// solved through the small SPD system G = I + E^T A E (Cholesky), followed by the same exact line
// search.  The minimiser is the one MuJoCo's primal Newton converges to (strictly convex problem).
enum { W_LAM = 0, W_JAR = 1, W_F = 2, W_R = 3, W_U = 4, W_DL = 5, W_ADL = 6, W_P = 7 };
// E columns: col c has entries on rows row..row+n-1 with values val[0..n-1]
#define ECOL_ROW(c) AT(d.efc_ecol, (c))
#define ECOL_VAL(c, k) AT(d.efc_eval, 3 * (c) + (k))

struct ConeInfo { float mu, f1, f2; };

// per-row cost/force evaluation at jar (W_JAR) -> W_F; optionally builds E columns.  returns cost.
FB_DEV float constraint_update(const DevModel& m, const DevData& d, int e, int n, bool build, int* ncol_out) {
  float cost = 0; int nc = 0;
  for (int i = 0; i < n; i++) {
    int tp = EFC(d.efc_type, i);
    float jar = EW(W_JAR, i), D = EFC(d.efc_D, i);
    if (tp != FB_CT_ELLIPTIC) {
      if (jar < 0) { EW(W_F, i) = -D * jar; cost += 0.5f * D * jar * jar; if (build) { ECOL_ROW(nc) = i; ECOL_VAL(nc, 0) = sqrtf(D); ECOL_VAL(nc, 1) = 0; ECOL_VAL(nc, 2) = 0; nc++; } }
      else EW(W_F, i) = 0;
    } else {
      int ci = EFC(d.efc_id, i);
      float mu = AT(d.con_mu, ci), f1 = CON_F(d.con_fric, ci, 0, 2), f2 = CON_F(d.con_fric, ci, 1, 2);
      float j1 = EW(W_JAR, i + 1), j2 = EW(W_JAR, i + 2), D1 = EFC(d.efc_D, i + 1), D2 = EFC(d.efc_D, i + 2);
      float U0 = jar * mu, U1 = j1 * f1, U2 = j2 * f2, N = U0, T = sqrtf(U1 * U1 + U2 * U2);
      if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        EW(W_F, i) = -D * jar; EW(W_F, i + 1) = -D1 * j1; EW(W_F, i + 2) = -D2 * j2;
        cost += 0.5f * (D * jar * jar + D1 * j1 * j1 + D2 * j2 * j2);
        if (build) { for (int r = 0; r < 3; r++) { ECOL_ROW(nc) = i + r; ECOL_VAL(nc, 0) = sqrtf(EFC(d.efc_D, i + r)); ECOL_VAL(nc, 1) = 0; ECOL_VAL(nc, 2) = 0; nc++; } }
      } else if (N >= mu * T || (T <= 0 && N >= 0)) {
        EW(W_F, i) = 0; EW(W_F, i + 1) = 0; EW(W_F, i + 2) = 0;
      } else {
        float Dm = D / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
        cost += 0.5f * Dm * NmT * NmT;
        float f0 = -Dm * NmT * mu;
        EW(W_F, i) = f0; EW(W_F, i + 1) = -f0 / T * U1 * f1; EW(W_F, i + 2) = -f0 / T * U2 * f2;
        if (build) {
          // C = Dm [ (S g)(S g)^T + (-f mu / T)(S t)(S t)^T ],  g = (1, -mu U1/T, -mu U2/T), t = (0, -U2, U1)/T
          float sD = sqrtf(Dm);
          ECOL_ROW(nc) = -(i + 1);            // negative: 3-row column starting at row i
          ECOL_VAL(nc, 0) = sD * mu; ECOL_VAL(nc, 1) = -sD * f1 * mu * U1 / T; ECOL_VAL(nc, 2) = -sD * f2 * mu * U2 / T; nc++;
          float k2 = sqrtf(fmaxf(0.0f, Dm * (-NmT) * mu / T));
          ECOL_ROW(nc) = -(i + 1);
          ECOL_VAL(nc, 0) = 0; ECOL_VAL(nc, 1) = -k2 * f1 * U2 / T; ECOL_VAL(nc, 2) = k2 * f2 * U1 / T; nc++;
        }
      }
      i += 2;
    }
  }
  if (ncol_out) *ncol_out = nc;
  return cost;
}
// 1-D cost along lam + alpha*dlam: constraint part only (value, first and second derivative)
FB_DEV void ls_eval(const DevModel& m, const DevData& d, int e, int n, float alpha, float& c, float& g, float& h) {
  for (int i = 0; i < n; i++) {
    int tp = EFC(d.efc_type, i);
    float jv = EW(W_ADL, i), x = EW(W_JAR, i) + alpha * jv, D = EFC(d.efc_D, i);
    if (tp != FB_CT_ELLIPTIC) { if (x < 0) { c += 0.5f * D * x * x; g += D * x * jv; h += D * jv * jv; } }
    else {
      int ci = EFC(d.efc_id, i);
      float mu = AT(d.con_mu, ci), f1 = CON_F(d.con_fric, ci, 0, 2), f2 = CON_F(d.con_fric, ci, 1, 2);
      float jv1 = EW(W_ADL, i + 1), jv2 = EW(W_ADL, i + 2);
      float x1 = EW(W_JAR, i + 1) + alpha * jv1, x2 = EW(W_JAR, i + 2) + alpha * jv2;
      float D1 = EFC(d.efc_D, i + 1), D2 = EFC(d.efc_D, i + 2);
      float U0 = x * mu, U1 = x1 * f1, U2 = x2 * f2, dU0 = jv * mu, dU1 = jv1 * f1, dU2 = jv2 * f2;
      float N = U0, T = sqrtf(U1 * U1 + U2 * U2);
      if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        c += 0.5f * (D * x * x + D1 * x1 * x1 + D2 * x2 * x2); g += D * x * jv + D1 * x1 * jv1 + D2 * x2 * jv2;
        h += D * jv * jv + D1 * jv1 * jv1 + D2 * jv2 * jv2;
      } else if (N >= mu * T || (T <= 0 && N >= 0)) {
      } else {
        float Dm = D / (mu * mu * (1 + mu * mu)), f = N - mu * T;
        float dT = (U1 * dU1 + U2 * dU2) / T, ddT = (dU1 * dU1 + dU2 * dU2 - dT * dT) / T;
        float fp = dU0 - mu * dT, fpp = -mu * ddT;
        c += 0.5f * Dm * f * f; g += Dm * f * fp; h += Dm * (fp * fp + f * fpp);
      }
      i += 2;
    }
  }
}
FB_DEV int qcqp2(float* res, const float* A, const float* b, float d0, float d1, float r) {   // mju_QCQP2
  float A11 = A[0] * d0 * d0, A22 = A[3] * d1 * d1, A12 = A[1] * d0 * d1, b1 = b[0] * d0, b2 = b[1] * d1;
  float la = 0, v1 = 0, v2 = 0;
  for (int it = 0; it < 20; it++) {
    float det = (A11 + la) * (A22 + la) - A12 * A12;
    if (det < 1e-10f) { res[0] = 0; res[1] = 0; return 0; }
    float di = 1 / det, P11 = (A22 + la) * di, P22 = (A11 + la) * di, P12 = -A12 * di;
    v1 = -P11 * b1 - P12 * b2; v2 = -P12 * b1 - P22 * b2;
    float val = v1 * v1 + v2 * v2 - r * r;
    if (val < 1e-10f) break;
    float deriv = -2 * (P11 * v1 * v1 + 2 * P12 * v1 * v2 + P22 * v2 * v2);
    float delta = -val / deriv;
    if (delta < 1e-10f) break;
    la += delta;
  }
  res[0] = v1 * d0; res[1] = v2 * d1;
  return la != 0;
}

FB_DEV void ksolve(const DevModel& m, const DevData& d, int e) {
  int n = AT(d.nefc, 0);
  AT(d.niter, 0) = 0;
  if (n == 0) { for (int k = 0; k < m.nv; k++) AT(d.qfrc_constraint, k) = 0; return; }
  float scale = 1.0f / (m.meaninertia * (m.nv > 1 ? m.nv : 1));
  // ---- warm start: forces implied by the previous qacc, kept if cheaper than lam = 0
  for (int i = 0; i < n; i++) EW(W_JAR, i) = EFC(d.efc_jarws, i);
  constraint_update(m, d, e, n, false, nullptr);
  for (int i = 0; i < n; i++) EW(W_LAM, i) = EW(W_F, i);
  float cost_ws, cost0;
  {
    float q = 0;
    for (int i = 0; i < n; i++) { float s = EFC(d.efc_b, i); for (int j = 0; j < n; j++) s += EA(d.efc_A, i, j) * EW(W_LAM, j); EW(W_JAR, i) = s; q += 0.5f * EW(W_LAM, i) * (s - EFC(d.efc_b, i)); }
    cost_ws = q + constraint_update(m, d, e, n, false, nullptr);
    for (int i = 0; i < n; i++) EW(W_JAR, i) = EFC(d.efc_b, i);
    cost0 = constraint_update(m, d, e, n, false, nullptr);
    if (!(cost_ws < cost0)) for (int i = 0; i < n; i++) EW(W_LAM, i) = 0;
  }
  int iter = 0;
  for (; iter < m.max_iter; iter++) {
    // jar = b + A lam, forces, E columns
    float quad = 0;
    for (int i = 0; i < n; i++) { float s = EFC(d.efc_b, i); for (int j = 0; j < n; j++) s += EA(d.efc_A, i, j) * EW(W_LAM, j); EW(W_JAR, i) = s; quad += 0.5f * EW(W_LAM, i) * (s - EFC(d.efc_b, i)); }
    int nc = 0;
    float cost = quad + constraint_update(m, d, e, n, true, &nc);
    float rr = 0, ll = 0;
    for (int i = 0; i < n; i++) { float r = EW(W_LAM, i) - EW(W_F, i); EW(W_R, i) = r; rr += r * r; ll += EW(W_F, i) * EW(W_F, i); }
    if (rr <= 1e-12f * (ll + 1e-30f)) break;
    // u = A r
    for (int i = 0; i < n; i++) { float s = 0; for (int j = 0; j < n; j++) s += EA(d.efc_A, i, j) * EW(W_R, j); EW(W_U, i) = s; }
    // G = I + E^T A E (lower triangle), p = E^T u
    for (int p = 0; p < nc; p++) {
      int rp = ECOL_ROW(p), np = 1; if (rp < 0) { rp = -rp - 1; np = 3; }
      float pv = 0; for (int a = 0; a < np; a++) pv += ECOL_VAL(p, a) * EW(W_U, rp + a);
      EW(W_P, p) = pv;
      for (int q = 0; q <= p; q++) {
        int rq = ECOL_ROW(q), nq = 1; if (rq < 0) { rq = -rq - 1; nq = 3; }
        float s = (p == q) ? 1.0f : 0.0f;
        for (int a = 0; a < np; a++) { float va = ECOL_VAL(p, a); if (va == 0.0f) continue; for (int bb = 0; bb < nq; bb++) s += va * EA(d.efc_A, rp + a, rq + bb) * ECOL_VAL(q, bb); }
        EA(d.efc_G, p, q) = s;
      }
    }
    // Cholesky G = L L^T in place, solve G q = p
    for (int j = 0; j < nc; j++) {
      float s = EA(d.efc_G, j, j);
      for (int k = 0; k < j; k++) { float l = EA(d.efc_G, j, k); s -= l * l; }
      s = sqrtf(fmaxf(s, 1e-12f)); EA(d.efc_G, j, j) = s; float inv = 1.0f / s;
      for (int i = j + 1; i < nc; i++) { float t = EA(d.efc_G, i, j); for (int k = 0; k < j; k++) t -= EA(d.efc_G, i, k) * EA(d.efc_G, j, k); EA(d.efc_G, i, j) = t * inv; }
    }
    for (int i = 0; i < nc; i++) { float s = EW(W_P, i); for (int k = 0; k < i; k++) s -= EA(d.efc_G, i, k) * EW(W_P, k); EW(W_P, i) = s / EA(d.efc_G, i, i); }
    for (int i = nc - 1; i >= 0; i--) { float s = EW(W_P, i); for (int k = i + 1; k < nc; k++) s -= EA(d.efc_G, k, i) * EW(W_P, k); EW(W_P, i) = s / EA(d.efc_G, i, i); }
    // dlam = -r + E q
    for (int i = 0; i < n; i++) EW(W_DL, i) = -EW(W_R, i);
    for (int p = 0; p < nc; p++) { int rp = ECOL_ROW(p), np = 1; if (rp < 0) { rp = -rp - 1; np = 3; } float qv = EW(W_P, p); for (int a = 0; a < np; a++) EW(W_DL, rp + a) += ECOL_VAL(p, a) * qv; }
    // A dlam, quadratic coefficients of the Gauss term
    float q1 = 0, q2 = 0;
    for (int i = 0; i < n; i++) { float s = 0; for (int j = 0; j < n; j++) s += EA(d.efc_A, i, j) * EW(W_DL, j); EW(W_ADL, i) = s; q1 += EW(W_DL, i) * (EW(W_JAR, i) - EFC(d.efc_b, i)); q2 += 0.5f * EW(W_DL, i) * s; }
    // exact line search (safeguarded Newton on the derivative)
    float c0 = quad, g0 = q1, h0 = 2 * q2;
    ls_eval(m, d, e, n, 0.0f, c0, g0, h0);
    if (!(g0 < 0) || !(h0 > 0)) break;
    float alpha = -g0 / h0, lo = 0, hi = -1, cbest = c0;
    for (int ls = 0; ls < m.ls_iter; ls++) {
      float c = quad + alpha * q1 + alpha * alpha * q2, g = q1 + 2 * alpha * q2, h = 2 * q2;
      ls_eval(m, d, e, n, alpha, c, g, h);
      cbest = c;
      if (fabsf(g) < 1e-6f * fabsf(g0)) break;
      if (g < 0) lo = alpha; else hi = alpha;
      float na = alpha - g / h;
      if (hi >= 0 && (na <= lo || na >= hi)) na = 0.5f * (lo + hi);
      else if (hi < 0 && na <= lo) na = 2 * alpha;
      if (fabsf(na - alpha) <= 1e-7f * fabsf(alpha)) { alpha = na; break; }
      alpha = na;
    }
    for (int i = 0; i < n; i++) EW(W_LAM, i) += alpha * EW(W_DL, i);
    float improvement = scale * (cost - cbest);
    if (improvement < m.tolerance) { iter++; break; }
  }
  AT(d.niter, 0) = iter;
  // final forces at the solution: lam = f(b + A lam)
  for (int i = 0; i < n; i++) { float s = EFC(d.efc_b, i); for (int j = 0; j < n; j++) s += EA(d.efc_A, i, j) * EW(W_LAM, j); EW(W_JAR, i) = s; }
  constraint_update(m, d, e, n, false, nullptr);
  for (int i = 0; i < n; i++) EFC(d.efc_force, i) = EW(W_F, i);
  // ---- noslip (MuJoCo mj_solNoSlip): Gauss-Seidel on the friction rows with the unregularised A
  if (m.noslip_iterations > 0) {
    for (int it = 0; it < m.noslip_iterations; it++) {
      float improvement = 0;
      if (it == 0) for (int i = 0; i < n; i++) improvement += 0.5f * EFC(d.efc_force, i) * EFC(d.efc_force, i) * EFC(d.efc_R, i);
      bool any = false;
      for (int i = 0; i < n; i++) {
        if (EFC(d.efc_type, i) != FB_CT_ELLIPTIC) continue;
        any = true;
        int ci = EFC(d.efc_id, i);
        float fn = EFC(d.efc_force, i), old0 = EFC(d.efc_force, i + 1), old1 = EFC(d.efc_force, i + 2);
        float res[2], Ac[4], bc[2], v[2];
        for (int r = 0; r < 2; r++) { float s = EFC(d.efc_b, i + 1 + r); for (int j = 0; j < n; j++) s += EA(d.efc_A, i + 1 + r, j) * EFC(d.efc_force, j); res[r] = s; }
        Ac[0] = EA(d.efc_A, i + 1, i + 1); Ac[1] = EA(d.efc_A, i + 1, i + 2); Ac[2] = EA(d.efc_A, i + 2, i + 1); Ac[3] = EA(d.efc_A, i + 2, i + 2);
        bc[0] = res[0] - Ac[0] * old0 - Ac[1] * old1; bc[1] = res[1] - Ac[2] * old0 - Ac[3] * old1;
        float fr0 = CON_F(d.con_fric, ci, 0, 2), fr1 = CON_F(d.con_fric, ci, 1, 2);
        if (fn < FB_MINVAL) { v[0] = 0; v[1] = 0; }
        else {
          int active = qcqp2(v, Ac, bc, fr0, fr1, fn);
          if (active) { float s = (v[0] / fr0) * (v[0] / fr0) + (v[1] / fr1) * (v[1] / fr1); s = sqrtf(fn * fn / fmaxf(FB_MINVAL, s)); v[0] *= s; v[1] *= s; }
        }
        float d0 = v[0] - old0, d1 = v[1] - old1;
        float change = 0.5f * (d0 * (Ac[0] * d0 + Ac[1] * d1) + d1 * (Ac[2] * d0 + Ac[3] * d1)) + d0 * res[0] + d1 * res[1];
        if (change > 1e-10f) { v[0] = old0; v[1] = old1; change = 0; }
        EFC(d.efc_force, i + 1) = v[0]; EFC(d.efc_force, i + 2) = v[1];
        improvement -= change;
        i += 2;
      }
      if (!any) break;
      if (improvement * scale < m.noslip_tolerance) break;
    }
  }
  // qfrc_constraint = J^T f
  for (int k = 0; k < m.nv; k++) AT(d.qfrc_constraint, k) = 0;
  for (int r = 0; r < n; r++) {
    float f = EFC(d.efc_force, r);
    if (f == 0.0f) continue;
    int ci, frow; float sign;
    RowChains rc = row_chains(m, d, e, r, ci, frow, sign);
    int la = rc.la, lb = rc.lb;
    while (la >= 0 || lb >= 0) {
      int k = la > lb ? la : lb;
      if (la == k) la = m.dof_parentid[la];
      if (lb == k) lb = m.dof_parentid[lb];
      AT(d.qfrc_constraint, k) += EJ(d.efc_J, r, k) * f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K12 acceleration-stage sensors + K13 Euler (tree kernel phases, see fb_kernels.cu for the order)
// phase: qtmp <- qfrc_constraint (rhs of M x = J^T f)
FB_DEV void kfin_copy(FB_PHASE_ARGS) {
  if (y == 0) for (int r = 0; r < m.nroot; r++) { int b = m.root_body[r]; for (int k = 0; k < m.body_dofnum[b]; k++) { int i = m.body_dofadr[b] + k; AT(d.qtmp, i) = AT(d.qfrc_constraint, i); } }
  if (y >= m.nlist) return;
  FB_LIST_LOOP_FWD { for (int k = 0; k < m.body_dofnum[b]; k++) { int i = m.body_dofadr[b] + k; AT(d.qtmp, i) = AT(d.qfrc_constraint, i); } }
}
FB_DEV void kfin_solve_a(FB_PHASE_ARGS) { solve_a(m, d, sh, e, lane, y, d.qLD, d.qtmp); }
FB_DEV void kfin_solve_b(FB_PHASE_ARGS) { solve_b(m, d, sh, e, lane, y, d.qLD, d.qtmp); }
FB_DEV void kfin_solve_c(FB_PHASE_ARGS) {
  solve_c(m, d, sh, e, lane, y, d.qLD, d.qtmp);
  // qacc = qacc_smooth + M^-1 J^T f for own dofs (root dofs by y == 0 were finalised in phase b)
  if (y == 0) for (int r = 0; r < m.nroot; r++) { int b = m.root_body[r]; for (int k = 0; k < m.body_dofnum[b]; k++) { int i = m.body_dofadr[b] + k; AT(d.qacc, i) = AT(d.qacc_smooth, i) + AT(d.qtmp, i); } }
  if (y >= m.nlist) return;
  FB_LIST_LOOP_FWD { for (int k = 0; k < m.body_dofnum[b]; k++) { int i = m.body_dofadr[b] + k; AT(d.qacc, i) = AT(d.qacc_smooth, i) + AT(d.qtmp, i); } }
}
// sensors: full RNE with qacc, minus contact forces -> cfrc_int (MuJoCo mj_rnePostConstraint)
FB_DEV void kfin_sens_root(FB_PHASE_ARGS) {
  if (y != 0) return;
  // external (contact) wrench per body into bfl (about ref), scattered sequentially
  for (int b = 0; b < m.nbody; b++) { S6 z; z.a = v3(0, 0, 0); z.l = v3(0, 0, 0); st6(d.bfl, b, d, e, z); }
  int ncon = AT(d.ncon, 0);
  for (int ci = 0; ci < ncon; ci++) {
    int adr = AT(d.con_efcadr, ci); if (adr < 0) continue;
    int dim = AT(d.con_dim, ci);
    V3 F = v3(0, 0, 0);
    for (int r = 0; r < dim; r++) F = F + v3(CON_F(d.con_frame, ci, 3 * r, 9), CON_F(d.con_frame, ci, 3 * r + 1, 9), CON_F(d.con_frame, ci, 3 * r + 2, 9)) * EFC(d.efc_force, adr + r);
    V3 pos = v3(CON_F(d.con_pos, ci, 0, 3), CON_F(d.con_pos, ci, 1, 3), CON_F(d.con_pos, ci, 2, 3));
    V3 tq = cross(pos, F);
    int b1 = m.geom_bodyid[AT(d.con_geom1, ci)], b2 = m.geom_bodyid[AT(d.con_geom2, ci)];
    S6 w2 = ld6(d.bfl, b2, d, e); w2.a = w2.a + tq; w2.l = w2.l + F; st6(d.bfl, b2, d, e, w2);
    S6 w1 = ld6(d.bfl, b1, d, e); w1.a = w1.a - tq; w1.l = w1.l - F; st6(d.bfl, b1, d, e, w1);
  }
  for (int r = 0; r < m.nroot; r++) body_vel_acc(m, d, e, m.root_body[r], d.qacc, d.bvel, d.bacc);
}
FB_DEV void kfin_sens_fwd(FB_PHASE_ARGS) {
  if (y >= m.nlist) return;
  FB_LIST_LOOP_FWD {
    body_vel_acc(m, d, e, b, d.qacc, d.bvel, d.bacc);
    S6 f = body_inertial_force(m, d, e, b, d.bvel, d.bacc), x = ld6(d.bfl, b, d, e);
    f.a = f.a - x.a; f.l = f.l - x.l;
    st6(d.bfrc, b, d, e, f);
  }
}
FB_DEV void kfin_sens_bwd(FB_PHASE_ARGS) {
  if (y >= m.nlist) return;
  FB_LIST_LOOP_REV {
    int p = m.body_parentid[b];
    if (m.body_isroot[p]) continue;       // cfrc_int of root bodies is not read by any fly sensor
    S6 f = ld6(d.bfrc, b, d, e), pf = ld6(d.bfrc, p, d, e);
    pf.a = pf.a + f.a; pf.l = pf.l + f.l; st6(d.bfrc, p, d, e, pf);
  }
}
FB_DEV float ray_quad(float a, float b, float c, float* x) {
  float det = b * b - a * c;
  if (det < 1e-30f) { x[0] = -1; x[1] = -1; return -1; }
  det = sqrtf(det); x[0] = (-b - det) / a; x[1] = (-b + det) / a;
  if (x[0] >= 0) return x[0]; if (x[1] >= 0) return x[1]; return -1;
}
FB_DEV float ray_capsule(V3 pos, const M3& mat, V3 size, V3 pnt, V3 vec) {
  V3 dif = pnt - pos; float xx[2]; float ssz = size.x + size.y;
  if (ray_quad(dot(vec, vec), dot(vec, dif), dot(dif, dif) - ssz * ssz, xx) < 0) return -1;
  V3 lp = mulT(mat, dif), lv = mulT(mat, vec);
  float x = -1;
  float sol = ray_quad(lv.x * lv.x + lv.y * lv.y, lv.x * lp.x + lv.y * lp.y, lp.x * lp.x + lp.y * lp.y - size.x * size.x, xx);
  if (sol >= 0 && fabsf(lp.z + sol * lv.z) <= size.y) { if (x < 0 || sol < x) x = sol; }
  V3 ld = v3(lp.x, lp.y, lp.z - size.y);
  ray_quad(dot(lv, lv), dot(lv, ld), dot(ld, ld) - size.x * size.x, xx);
  for (int i = 0; i < 2; i++) if (xx[i] >= 0 && lp.z + xx[i] * lv.z >= size.y) { if (x < 0 || xx[i] < x) x = xx[i]; }
  ld.z = lp.z + size.y;
  ray_quad(dot(lv, lv), dot(lv, ld), dot(ld, ld) - size.x * size.x, xx);
  for (int i = 0; i < 2; i++) if (xx[i] >= 0 && lp.z + xx[i] * lv.z <= -size.y) { if (x < 0 || xx[i] < x) x = xx[i]; }
  return x;
}
FB_DEV void kfin_sens_out(FB_PHASE_ARGS) {
  if (y != 0) return;
  int ncon = AT(d.ncon, 0);
  for (int s = 0; s < m.nsensor; s++) {
    int tp = m.sensor_type[s], site = m.sensor_objid[s], adr = m.sensor_adr[s], b = m.site_bodyid[site];
    if (tp == FB_SENS_ACCELEROMETER) {
      S6 a = ld6(d.bacc, b, d, e), v = ld6(d.bvel, b, d, e);
      V3 p = ld3(d.site_xpos, site, d, e);
      V3 vp = v.l + cross(v.a, p);
      V3 acc = a.l + cross(a.a, p) + cross(v.a, vp);
      V3 out = mulT(ld9(d.site_xmat, site, d, e), acc);
      AT(d.sensordata, adr) = out.x; AT(d.sensordata, adr + 1) = out.y; AT(d.sensordata, adr + 2) = out.z;
    } else if (tp == FB_SENS_FORCE) {
      V3 out = mulT(ld9(d.site_xmat, site, d, e), ld3(d.bfrc, 2 * b + 1, d, e));
      AT(d.sensordata, adr) = out.x; AT(d.sensordata, adr + 1) = out.y; AT(d.sensordata, adr + 2) = out.z;
    } else if (tp == FB_SENS_TOUCH) {
      float sum = 0;
      for (int ci = 0; ci < ncon; ci++) {
        int ea = AT(d.con_efcadr, ci); if (ea < 0) continue;
        int b1 = m.geom_bodyid[AT(d.con_geom1, ci)], b2 = m.geom_bodyid[AT(d.con_geom2, ci)];
        if (b != b1 && b != b2) continue;
        float fn = EFC(d.efc_force, ea); if (fn <= 0) continue;
        V3 ray = v3(CON_F(d.con_frame, ci, 0, 9), CON_F(d.con_frame, ci, 1, 9), CON_F(d.con_frame, ci, 2, 9));
        if (b == b2) ray = ray * -1.0f;
        V3 pos = v3(CON_F(d.con_pos, ci, 0, 3), CON_F(d.con_pos, ci, 1, 3), CON_F(d.con_pos, ci, 2, 3));
        V3 ssz = mld3(m.site_size, site);
        float hit;
        if (m.site_type[site] == FB_GEOM_CAPSULE) hit = ray_capsule(ld3(d.site_xpos, site, d, e), ld9(d.site_xmat, site, d, e), ssz, pos, ray);
        else { float xx[2]; V3 dif = pos - ld3(d.site_xpos, site, d, e); hit = ray_quad(dot(ray, ray), dot(ray, dif), dot(dif, dif) - ssz.x * ssz.x, xx); }
        if (hit >= 0) sum += fn;
      }
      AT(d.sensordata, adr) = sum;
    }
  }
  // per-substep sensor accumulation + state check (|qacc| > 1e14 or non-finite: reference tasks/base.py:222-225)
  float s2 = 0; bool bad = false;
  for (int k = 0; k < m.nv; k++) { float a = AT(d.qacc, k); s2 += a * a; if (!isfinite(a) || !isfinite(AT(d.qvel, k))) bad = true; }
  if (bad || !(s2 < 1e28f)) FB_FLAG_OR(1);
}
// Euler: qtmp <- qfrc_smooth + qfrc_constraint, solve with the (M + h D) factor, integrate
FB_DEV void keul_rhs(FB_PHASE_ARGS) {
  if (y == 0) for (int r = 0; r < m.nroot; r++) { int b = m.root_body[r]; for (int k = 0; k < m.body_dofnum[b]; k++) { int i = m.body_dofadr[b] + k; AT(d.qtmp, i) = AT(d.qfrc_smooth, i) + AT(d.qfrc_constraint, i); } }
  if (y >= m.nlist) return;
  FB_LIST_LOOP_FWD { for (int k = 0; k < m.body_dofnum[b]; k++) { int i = m.body_dofadr[b] + k; AT(d.qtmp, i) = AT(d.qfrc_smooth, i) + AT(d.qfrc_constraint, i); } }
}
FB_DEV void keul_solve_a(FB_PHASE_ARGS) { solve_a(m, d, sh, e, lane, y, d.qLDe, d.qtmp); }
FB_DEV void keul_solve_b(FB_PHASE_ARGS) { solve_b(m, d, sh, e, lane, y, d.qLDe, d.qtmp); }
FB_DEV void integrate_body(const DevModel& m, const DevData& d, int e, int b) {
  float h = m.timestep;
  for (int k = 0; k < m.body_jntnum[b]; k++) {
    int j = m.body_jntadr[b] + k, qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    if (m.jnt_type[j] == FB_JNT_FREE) {
      for (int i = 0; i < 6; i++) { AT(d.qvel, da + i) += h * AT(d.qtmp, da + i); AT(d.qacc_warmstart, da + i) = AT(d.qacc, da + i); }
      for (int i = 0; i < 3; i++) AT(d.qpos, qa + i) += h * AT(d.qvel, da + i);
      V3 w = v3(AT(d.qvel, da + 3), AT(d.qvel, da + 4), AT(d.qvel, da + 5));
      float ang = norm(w) * h;
      Q4 q = q4(AT(d.qpos, qa + 3), AT(d.qpos, qa + 4), AT(d.qpos, qa + 5), AT(d.qpos, qa + 6));
      if (ang > 0) q = qmul(q, axisangle(normalized(w), ang));
      q = qnormalize(q);
      AT(d.qpos, qa + 3) = q.w; AT(d.qpos, qa + 4) = q.x; AT(d.qpos, qa + 5) = q.y; AT(d.qpos, qa + 6) = q.z;
    } else {
      AT(d.qvel, da) += h * AT(d.qtmp, da); AT(d.qacc_warmstart, da) = AT(d.qacc, da);
      AT(d.qpos, qa) += h * AT(d.qvel, da);
    }
  }
}
FB_DEV void keul_solve_c_integrate(FB_PHASE_ARGS) {
  solve_c(m, d, sh, e, lane, y, d.qLDe, d.qtmp);
  if (AT(d.hold, 0)) return;          // env staged for reset: recompute (forward) but do not integrate
  if (y == 0) {
    for (int r = 0; r < m.nroot; r++) integrate_body(m, d, e, m.root_body[r]);
    for (int i = 0; i < m.na; i++) AT(d.act, i) += m.timestep * AT(d.act_dot, i);
    AT(d.time, 0) += m.timestep;
  }
  if (y >= m.nlist) return;
  FB_LIST_LOOP_FWD integrate_body(m, d, e, b);
}
// accumulate sensor sums (after step1 of the substep: vel sensors are from the new state)
FB_DEV void ksens_accum(const DevModel& m, const DevData& d, int e, int first) {
  for (int i = 0; i < m.nsensordata; i++) AT(d.sensor_sum, i) = (first ? 0.0f : AT(d.sensor_sum, i)) + AT(d.sensordata, i);
}

// ---------------------------------------------------------------------------------------------
// packed per-env observation record (AoS, one row per env) for the host task code / NCCL gather:
//   qpos[nq] qvel[nv] act[na] sensor_mean[nsd] sensordata[nsd] root_xpos[3] root_xmat[9] site_xpos[3*nsite]
//   flags[1] qacc_sq[1] time[1]
FB_DEV void kpack(const DevModel& m, const DevData& d, int e, float inv_nsub) {
  float* o = d.obs + (size_t)e * d.obs_dim;
  int k = 0;
  for (int i = 0; i < m.nq; i++) o[k++] = AT(d.qpos, i);
  for (int i = 0; i < m.nv; i++) o[k++] = AT(d.qvel, i);
  for (int i = 0; i < m.na; i++) o[k++] = AT(d.act, i);
  for (int i = 0; i < m.nsensordata; i++) o[k++] = AT(d.sensor_sum, i) * inv_nsub;
  for (int i = 0; i < m.nsensordata; i++) o[k++] = AT(d.sensordata, i);
  int rb = m.root_body[0];
  V3 ref = v3(AT(d.ref, 0), AT(d.ref, 1), AT(d.ref, 2));
  V3 rp = ld3(d.xpos, rb, d, e) + ref;
  o[k++] = rp.x; o[k++] = rp.y; o[k++] = rp.z;
  for (int i = 0; i < 9; i++) o[k++] = AT(d.xmat, 9 * rb + i);
  for (int s = 0; s < m.nsite; s++) { V3 p = ld3(d.site_xpos, s, d, e) + ref; o[k++] = p.x; o[k++] = p.y; o[k++] = p.z; }
  o[k++] = (float)AT(d.flags, 0);
  float s2 = 0; for (int i = 0; i < m.nv; i++) { float a = AT(d.qacc, i); s2 += a * a; }
  o[k++] = s2; o[k++] = AT(d.time, 0);
}

// partial reset staged by fb_reset_hold: thread k (global index) rewrites env rst_ids[k]
FB_DEV void kreset_scatter(const DevModel& m, const DevData& d, int k) {
  if (k >= d.rst_n) return;
  int e = d.rst_ids[k];
  for (int i = 0; i < m.nq; i++) AT(d.qpos, i) = d.rst_qpos[(size_t)k * m.nq + i];
  for (int i = 0; i < m.nv; i++) { AT(d.qvel, i) = d.rst_has_qvel ? d.rst_qvel[(size_t)k * m.nv + i] : 0.0f; AT(d.qacc, i) = 0; AT(d.qacc_warmstart, i) = 0; }
  for (int i = 0; i < m.na; i++) AT(d.act, i) = 0;
  AT(d.time, 0) = 0; AT(d.flags, 0) = 0; AT(d.hold, 0) = 1;
}
FB_DEV void kclear_hold(const DevModel& m, const DevData& d, int e) { AT(d.hold, 0) = 0; }
