// fb_solver.h -- block-cooperative version of the constraint solve (K11).
//
// blockDim = (32 envs, FB_SOLVE_Y).  lane == env as everywhere else; the FB_SOLVE_Y threads that
// share a lane split the rows of that env's dual problem (matrix-vector products with the Delassus
// matrix, the Hessian-factor assembly, Cholesky columns, line-search partial sums) and meet at block
// barriers.  Control flow is block-uniform: loops run to the block-wide maximum trip count and envs
// that are already converged idle through the barriers.
//
// The code is written as a sequence of PAR sections (each ends in a barrier); on the GPU every thread
// executes the uniform code between sections redundantly, in the host emulation a PAR section is a
// loop over (y, lane).  Per-env scalars therefore live in shared memory (sh.sc / sh.isc).
#pragma once
#include "fb_constraint.h"

#define FB_SOLVE_Y 16

struct ShSolve {
  float red[FB_SOLVE_Y][4][32];
  float sc[20][32];
  int isc[8][32];
};
enum { SC_COST = 0, SC_QUAD, SC_Q1, SC_Q2, SC_ALPHA, SC_LO, SC_HI, SC_G0, SC_CBEST, SC_DIAG, SC_COSTWS, SC_COST0, SC_RR, SC_LL, SC_CPREV };
enum { I_N = 0, I_NC, I_DONE, I_ITER, I_LSDONE };
// extra solver vectors (beyond W_LAM..W_P of fb_constraint.h) live in efc_w2
enum { X_E0 = 8, X_E1, X_XQ, X_OUT,
       // row constants staged once per solve (so the Newton phases never chase global metadata)
       S_D, S_B, S_R, S_TYPE /* 0 plain, 1 elliptic head, 2 elliptic tail */, S_MU, S_F1, S_F2, S_LA, S_LB,
       S_STATE, S_COLIDX, S_ECROW, S_ECKIND, S_NSLOT };
// solver vectors: 12 slots (W_LAM..W_P = 0..7, X_* = 8..11).  When they fit, they live in shared memory
// ([slot][row][lane], row stride = block-wide max nefc); otherwise in the global efc_w / efc_w2 arrays.
struct SolveCtx { float* vsh; int vstride; float* gsh; int gcap; int gmode; float* ash; int amode; };
#define FB_SOLVE_DYN_FLOATS (190 * 256)
#define SWG(slot, r) ((slot) < 8 ? &AT(d.efc_w, (slot) * FB_MAXEFC + (r)) : &AT(d.efc_w2, ((slot) - 8) * FB_MAXEFC + (r)))
#define SW(slot, r) (*(cx.vsh ? &cx.vsh[((slot) * cx.vstride + (r)) * 32 + lane] : SWG(slot, r)))
// Delassus matrix: packed lower triangle in shared memory when it fits, else the global array
#define TRI(r, c) ((r) * ((r) + 1) / 2 + (c))
#define AM(r, c) (cx.amode ? cx.ash[((r) >= (c) ? TRI(r, c) : TRI(c, r)) * 32 + lane] : EA(d.efc_A, r, c))
#define GM(p, q) (*(cx.gmode ? &cx.gsh[((p) * cx.gcap + (q)) * 32 + lane] : &EA(d.efc_G, p, q)))
#define ESTATE(r) SW(S_STATE, r)
#define ECOLIDX(r) SW(S_COLIDX, r)
#define ECROW(p) SW(S_ECROW, p)
#define ECKIND(p) SW(S_ECKIND, p)

#ifdef __CUDACC__
#define PAR_BEGIN { const int lane = threadIdx.x, y = threadIdx.y; const int e = blk * 32 + lane; (void)y; (void)e; (void)lane;
#define PAR_END } __syncthreads();
#define FB_BLOCKFN __device__ __noinline__
#else
#define PAR_BEGIN for (int y = 0; y < FB_SOLVE_Y; y++) for (int lane = 0; lane < 32; lane++) { const int e = blk * 32 + lane; (void)e;
#define PAR_END }
#define FB_BLOCKFN static
#endif
#define MY_N (sh.isc[I_N][lane])
#define ACTIVE (!sh.isc[I_DONE][lane])

FB_DEV float red_sum(const ShSolve& sh, int k, int lane) { float s_ = 0; for (int yy = 0; yy < FB_SOLVE_Y; yy++) s_ += sh.red[yy][k][lane]; return s_; }
// block-uniform helpers (read shared after a barrier)
FB_DEV int blk_max_i(const ShSolve& sh, int slot) { int v = 0; for (int l = 0; l < 32; l++) v = sh.isc[slot][l] > v ? sh.isc[slot][l] : v; return v; }
FB_DEV int blk_any_active(const ShSolve& sh) { int v = 0; for (int l = 0; l < 32; l++) v |= !sh.isc[I_DONE][l]; return v; }
FB_DEV int blk_any_ls(const ShSolve& sh) { int v = 0; for (int l = 0; l < 32; l++) v |= (!sh.isc[I_DONE][l] && !sh.isc[I_LSDONE][l]); return v; }

// forces / cost of the rows headed at r (a non-elliptic row, or the first row of an elliptic contact).
// Returns the cost; writes W_F; with build: state + E values.
FB_DEV float head_update(const DevModel& m, const DevData& d, const SolveCtx& cx, int lane, int e, int r, bool build) {
  int tp = (int)SW(S_TYPE, r);
  float jar = SW(W_JAR, r), D = SW(S_D, r), cost = 0;
  if (tp == 0) {
    if (jar < 0) { SW(W_F, r) = -D * jar; cost = 0.5f * D * jar * jar; if (build) { ESTATE(r) = 1; SW(X_E0, r) = sqrtf(D); } }
    else { SW(W_F, r) = 0; if (build) ESTATE(r) = 0; }
    return cost;
  }
  float mu = SW(S_MU, r), f1 = SW(S_F1, r), f2 = SW(S_F2, r);
  float j1 = SW(W_JAR, r + 1), j2 = SW(W_JAR, r + 2), D1 = SW(S_D, r + 1), D2 = SW(S_D, r + 2);
  float U0 = jar * mu, U1 = j1 * f1, U2 = j2 * f2, N = U0, T = sqrtf(U1 * U1 + U2 * U2);
  if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
    SW(W_F, r) = -D * jar; SW(W_F, r + 1) = -D1 * j1; SW(W_F, r + 2) = -D2 * j2;
    cost = 0.5f * (D * jar * jar + D1 * j1 * j1 + D2 * j2 * j2);
    if (build) { ESTATE(r) = 1; ESTATE(r + 1) = 1; ESTATE(r + 2) = 1; SW(X_E0, r) = sqrtf(D); SW(X_E0, r + 1) = sqrtf(D1); SW(X_E0, r + 2) = sqrtf(D2); }
  } else if (N >= mu * T || (T <= 0 && N >= 0)) {
    SW(W_F, r) = 0; SW(W_F, r + 1) = 0; SW(W_F, r + 2) = 0;
    if (build) { ESTATE(r) = 0; ESTATE(r + 1) = 0; ESTATE(r + 2) = 0; }
  } else {
    float Dm = D / (mu * mu * (1 + mu * mu)), NmT = N - mu * T;
    cost = 0.5f * Dm * NmT * NmT;
    float f0 = -Dm * NmT * mu;
    SW(W_F, r) = f0; SW(W_F, r + 1) = -f0 / T * U1 * f1; SW(W_F, r + 2) = -f0 / T * U2 * f2;
    if (build) {
      float sD = sqrtf(Dm), k2 = sqrtf(fmaxf(0.0f, Dm * (-NmT) * mu / T));
      ESTATE(r) = 2; ESTATE(r + 1) = 3; ESTATE(r + 2) = 3;
      SW(X_E0, r) = sD * mu; SW(X_E0, r + 1) = -sD * f1 * mu * U1 / T; SW(X_E0, r + 2) = -sD * f2 * mu * U2 / T;
      SW(X_E1, r) = 0; SW(X_E1, r + 1) = -k2 * f1 * U2 / T; SW(X_E1, r + 2) = k2 * f2 * U1 / T;
    }
  }
  return cost;
}
#define IS_HEAD(r) (SW(S_TYPE, r) < 1.5f)
// line-search contribution of the rows headed at r
FB_DEV void head_ls(const DevModel& m, const DevData& d, const SolveCtx& cx, int lane, int e, int r, float alpha, float& c, float& g, float& h) {
  int tp = (int)SW(S_TYPE, r);
  float jv = SW(W_ADL, r), x = SW(W_JAR, r) + alpha * jv, D = SW(S_D, r);
  if (tp == 0) { if (x < 0) { c += 0.5f * D * x * x; g += D * x * jv; h += D * jv * jv; } return; }
  float mu = SW(S_MU, r), f1 = SW(S_F1, r), f2 = SW(S_F2, r);
  float jv1 = SW(W_ADL, r + 1), jv2 = SW(W_ADL, r + 2);
  float x1 = SW(W_JAR, r + 1) + alpha * jv1, x2 = SW(W_JAR, r + 2) + alpha * jv2;
  float D1 = SW(S_D, r + 1), D2 = SW(S_D, r + 2);
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
}
// column p of E: value on row `row` (0 if the column does not touch it)
FB_DEV float ecol_val(const DevData& d, const SolveCtx& cx, int lane, int e, int p, int a) {   // a-th entry of column p
  int kind = (int)ECKIND(p), r = (int)ECROW(p);
  return kind == 2 ? SW(X_E1, r + a) : SW(X_E0, r + a);
}

#define ROWS_BEGIN for (int r = y; r < MY_N; r += FB_SOLVE_Y) {
#define ROWS_END }
// sum the per-thread partials of slot k into per-env scalar `dst` (call inside a y==0 guard)
#define RED_SUM(k) red_sum(sh, k, lane)

FB_BLOCKFN void ksolve_block(const DevModel& m, const DevData& d, ShSolve& sh, int blk) {
  const float scale = 1.0f / (m.meaninertia * (m.nv > 1 ? m.nv : 1));
  PAR_BEGIN
    if (y == 0) { int n = AT(d.nefc, 0); sh.isc[I_N][lane] = n; sh.isc[I_DONE][lane] = (n == 0); sh.isc[I_ITER][lane] = 0; sh.isc[I_NC][lane] = 0; sh.isc[I_LSDONE][lane] = 1; }
  PAR_END
  SolveCtx cx; cx.vsh = nullptr; cx.vstride = 0; cx.gsh = nullptr; cx.gcap = 0; cx.gmode = 0; cx.ash = nullptr; cx.amode = 0;
  if (blk_max_i(sh, I_N) > 0) {
    {
      int nmaxb = blk_max_i(sh, I_N);
      float* dyn = sh_dyn(sh);
      int rem = m.solve_dyn_floats / 32;          // floats per lane: vectors first, then A (packed), then G
      if (S_NSLOT * nmaxb <= rem) {
        cx.vsh = dyn; cx.vstride = nmaxb; rem -= S_NSLOT * nmaxb; dyn += (size_t)S_NSLOT * nmaxb * 32;
        int tri = nmaxb * (nmaxb + 1) / 2;
        if (tri <= rem) { cx.ash = dyn; cx.amode = 1; rem -= tri; dyn += (size_t)tri * 32; }
        int g = 0; while ((g + 1) * (g + 1) <= rem) g++;
        cx.gsh = dyn; cx.gcap = g;
      }
    }
    // ---------------- stage row constants (and A) into shared memory
    PAR_BEGIN ROWS_BEGIN
      int tp = EFC(d.efc_type, r), tcode = 0; float mu = 0, f1 = 0, f2 = 0;
      if (tp == FB_CT_ELLIPTIC) { int ci = EFC(d.efc_id, r); tcode = (AT(d.con_efcadr, ci) == r) ? 1 : 2; mu = AT(d.con_mu, ci); f1 = CON_F(d.con_fric, ci, 0, 2); f2 = CON_F(d.con_fric, ci, 1, 2); }
      SW(S_TYPE, r) = (float)tcode; SW(S_MU, r) = mu; SW(S_F1, r) = f1; SW(S_F2, r) = f2;
      SW(S_D, r) = EFC(d.efc_D, r); SW(S_B, r) = EFC(d.efc_b, r); SW(S_R, r) = EFC(d.efc_R, r);
      SW(S_LA, r) = (float)AT(d.efc_la, r); SW(S_LB, r) = (float)AT(d.efc_lb, r);
      if (cx.amode) for (int c = 0; c <= r; c++) cx.ash[TRI(r, c) * 32 + lane] = EA(d.efc_A, r, c);
    ROWS_END PAR_END
    // ---------------- warm start
    PAR_BEGIN float c = 0; ROWS_BEGIN SW(W_JAR, r) = EFC(d.efc_jarws, r); ROWS_END sh.red[y][0][lane] = c; PAR_END
    PAR_BEGIN ROWS_BEGIN if (IS_HEAD(r)) head_update(m, d, cx, lane, e, r, false); ROWS_END PAR_END
    PAR_BEGIN ROWS_BEGIN SW(W_LAM, r) = SW(W_F, r); ROWS_END PAR_END
    PAR_BEGIN float q = 0; int n = MY_N;
      ROWS_BEGIN float s = SW(S_B, r); for (int j = 0; j < n; j++) s += AM(r, j) * SW(W_LAM, j); SW(W_JAR, r) = s; q += 0.5f * SW(W_LAM, r) * (s - SW(S_B, r)); ROWS_END
      sh.red[y][0][lane] = q; PAR_END
    PAR_BEGIN float c = 0; ROWS_BEGIN if (IS_HEAD(r)) c += head_update(m, d, cx, lane, e, r, false); ROWS_END sh.red[y][1][lane] = c; PAR_END
    PAR_BEGIN if (y == 0) sh.sc[SC_COSTWS][lane] = RED_SUM(0) + RED_SUM(1); ROWS_BEGIN SW(W_JAR, r) = SW(S_B, r); ROWS_END PAR_END
    PAR_BEGIN float c = 0; ROWS_BEGIN if (IS_HEAD(r)) c += head_update(m, d, cx, lane, e, r, false); ROWS_END sh.red[y][0][lane] = c; PAR_END
    PAR_BEGIN if (y == 0) sh.sc[SC_COST0][lane] = RED_SUM(0); PAR_END
    PAR_BEGIN if (!(sh.sc[SC_COSTWS][lane] < sh.sc[SC_COST0][lane])) { ROWS_BEGIN SW(W_LAM, r) = 0; ROWS_END } PAR_END
    // ---------------- Newton iterations
    for (int iter = 0; iter < m.max_iter; iter++) {
      if (!blk_any_active(sh)) break;
      // jar = b + A lam
      PAR_BEGIN float q = 0; int n = MY_N;
        if (ACTIVE) { ROWS_BEGIN float s = SW(S_B, r); for (int j = 0; j < n; j++) s += AM(r, j) * SW(W_LAM, j); SW(W_JAR, r) = s; q += 0.5f * SW(W_LAM, r) * (s - SW(S_B, r)); ROWS_END }
        sh.red[y][0][lane] = q; PAR_END
      PAR_BEGIN float c = 0; if (ACTIVE) { ROWS_BEGIN if (IS_HEAD(r)) c += head_update(m, d, cx, lane, e, r, true); ROWS_END } sh.red[y][1][lane] = c; PAR_END
      // residual r = lam - f, column bookkeeping (sequential prefix over rows by y == 0)
      PAR_BEGIN float rr = 0, ll = 0;
        if (ACTIVE) { ROWS_BEGIN float f = SW(W_F, r), rv = SW(W_LAM, r) - f; SW(W_R, r) = rv; rr += rv * rv; ll += f * f; ROWS_END }
        sh.red[y][2][lane] = rr; sh.red[y][3][lane] = ll;
        if (y == 0 && ACTIVE) {
          sh.sc[SC_QUAD][lane] = RED_SUM(0); sh.sc[SC_COST][lane] = RED_SUM(0) + RED_SUM(1);
          int n = MY_N, nc = 0;
          for (int r = 0; r < n; r++) {
            int stt = (int)ESTATE(r);
            if (stt == 1) { ECOLIDX(r) = nc; ECROW(nc) = r; ECKIND(nc) = 0; nc++; }
            else if (stt == 2) { ECOLIDX(r) = nc; ECOLIDX(r + 1) = nc; ECOLIDX(r + 2) = nc; ECROW(nc) = r; ECKIND(nc) = 1; ECROW(nc + 1) = r; ECKIND(nc + 1) = 2; nc += 2; }
            else if (stt == 0) ECOLIDX(r) = -1;
          }
          sh.isc[I_NC][lane] = nc;
        }
      PAR_END
      PAR_BEGIN if (y == 0 && ACTIVE) { float rr = RED_SUM(2), ll = RED_SUM(3); if (rr <= 1e-12f * (ll + 1e-30f)) sh.isc[I_DONE][lane] = 1; } PAR_END
      if (!blk_any_active(sh)) break;
      // u = A r
      PAR_BEGIN int n = MY_N; if (ACTIVE) { ROWS_BEGIN float s = 0; for (int j = 0; j < n; j++) s += AM(r, j) * SW(W_R, j); SW(W_U, r) = s; ROWS_END } PAR_END
      int ncmax = 0; for (int l = 0; l < 32; l++) if (!sh.isc[I_DONE][l] && sh.isc[I_NC][l] > ncmax) ncmax = sh.isc[I_NC][l];
      cx.gmode = (cx.gsh != nullptr && ncmax <= cx.gcap) ? 1 : 0;
      // p = E^T u ; G = I + E^T A E (lower triangle)
      PAR_BEGIN if (ACTIVE) { int nc = sh.isc[I_NC][lane];
        for (int p = y; p < nc; p += FB_SOLVE_Y) {
          int rp = (int)ECROW(p), np = ECKIND(p) == 0 ? 1 : 3;
          float pv = 0; for (int a = 0; a < np; a++) pv += ecol_val(d, cx, lane, e, p, a) * SW(W_U, rp + a);
          SW(W_P, p) = pv;
          for (int q = 0; q <= p; q++) {
            int rq = (int)ECROW(q), nq = ECKIND(q) == 0 ? 1 : 3;
            float s = (p == q) ? 1.0f : 0.0f;
            for (int a = 0; a < np; a++) { float va = ecol_val(d, cx, lane, e, p, a); if (va == 0.0f) continue; for (int bb = 0; bb < nq; bb++) s += va * AM(rp + a, rq + bb) * ecol_val(d, cx, lane, e, q, bb); }
            GM(p, q) = s;
          }
        } }
      PAR_END
      // Cholesky G = L L^T (left-looking, two barriers per column), then the two triangular solves
      for (int j = 0; j < ncmax; j++) {
        PAR_BEGIN if (ACTIVE) { int nc = sh.isc[I_NC][lane];
          if (j < nc) for (int i = j + ((y - j % FB_SOLVE_Y + FB_SOLVE_Y) % FB_SOLVE_Y); i < nc; i += FB_SOLVE_Y) {
            float t = GM(i, j); for (int k = 0; k < j; k++) t -= GM(i, k) * GM(j, k);
            GM(i, j) = t; if (i == j) sh.sc[SC_DIAG][lane] = sqrtf(fmaxf(t, 1e-12f));
          } }
        PAR_END
        PAR_BEGIN if (ACTIVE) { int nc = sh.isc[I_NC][lane]; float dg = sh.sc[SC_DIAG][lane];
          if (j < nc) for (int i = j + ((y - j % FB_SOLVE_Y + FB_SOLVE_Y) % FB_SOLVE_Y); i < nc; i += FB_SOLVE_Y) GM(i, j) = (i == j) ? dg : GM(i, j) / dg; }
        PAR_END
      }
      for (int j = 0; j < ncmax; j++) {     // forward: L xq = p
        PAR_BEGIN if (ACTIVE) { int nc = sh.isc[I_NC][lane];
          if (j < nc) { float xj = SW(W_P, j) / GM(j, j);
            for (int i = j + 1 + ((y - (j + 1) % FB_SOLVE_Y + FB_SOLVE_Y) % FB_SOLVE_Y); i < nc; i += FB_SOLVE_Y) SW(W_P, i) -= GM(i, j) * xj;
            if (y == 0) SW(X_XQ, j) = xj; } }
        PAR_END
      }
      for (int j = ncmax - 1; j >= 0; j--) {   // backward: L^T out = xq
        PAR_BEGIN if (ACTIVE) { int nc = sh.isc[I_NC][lane];
          if (j < nc) { float xj = SW(X_XQ, j) / GM(j, j);
            for (int i = y; i < j; i += FB_SOLVE_Y) SW(X_XQ, i) -= GM(j, i) * xj;
            if (y == 0) SW(X_OUT, j) = xj; } }
        PAR_END
      }
      // dlam = -r + E q
      PAR_BEGIN if (ACTIVE) { ROWS_BEGIN
          float v = -SW(W_R, r); int stt = (int)ESTATE(r), c0 = (int)ECOLIDX(r);
          if (stt == 1) v += SW(X_E0, r) * SW(X_OUT, c0);
          else if (stt >= 2) v += SW(X_E0, r) * SW(X_OUT, c0) + SW(X_E1, r) * SW(X_OUT, c0 + 1);
          SW(W_DL, r) = v;
        ROWS_END } PAR_END
      // A dlam, q1, q2
      PAR_BEGIN float q1 = 0, q2 = 0; int n = MY_N;
        if (ACTIVE) { ROWS_BEGIN float s = 0; for (int j = 0; j < n; j++) s += AM(r, j) * SW(W_DL, j); SW(W_ADL, r) = s;
          q1 += SW(W_DL, r) * (SW(W_JAR, r) - SW(S_B, r)); q2 += 0.5f * SW(W_DL, r) * s; ROWS_END }
        sh.red[y][0][lane] = q1; sh.red[y][1][lane] = q2; PAR_END
      PAR_BEGIN if (y == 0 && ACTIVE) { sh.sc[SC_Q1][lane] = RED_SUM(0); sh.sc[SC_Q2][lane] = RED_SUM(1); sh.sc[SC_ALPHA][lane] = 0; sh.sc[SC_LO][lane] = 0; sh.sc[SC_HI][lane] = -1; sh.isc[I_LSDONE][lane] = 0; } PAR_END
      // exact line search: evaluation 0 at alpha = 0, then safeguarded Newton on the derivative
      for (int ls = 0; ls <= m.ls_iter; ls++) {
        if (!blk_any_ls(sh)) break;
        PAR_BEGIN float c = 0, g = 0, h = 0;
          if (ACTIVE && !sh.isc[I_LSDONE][lane]) { float alpha = sh.sc[SC_ALPHA][lane]; ROWS_BEGIN if (IS_HEAD(r)) head_ls(m, d, cx, lane, e, r, alpha, c, g, h); ROWS_END }
          sh.red[y][0][lane] = c; sh.red[y][1][lane] = g; sh.red[y][2][lane] = h; PAR_END
        PAR_BEGIN if (y == 0 && ACTIVE && !sh.isc[I_LSDONE][lane]) {
          float alpha = sh.sc[SC_ALPHA][lane], quad = sh.sc[SC_QUAD][lane], q1 = sh.sc[SC_Q1][lane], q2 = sh.sc[SC_Q2][lane];
          float c = quad + alpha * q1 + alpha * alpha * q2 + RED_SUM(0), g = q1 + 2 * alpha * q2 + RED_SUM(1), h = 2 * q2 + RED_SUM(2);
          if (ls == 0) {
            sh.sc[SC_G0][lane] = g; sh.sc[SC_CBEST][lane] = c;
            if (!(g < 0) || !(h > 0)) { sh.isc[I_LSDONE][lane] = 1; sh.isc[I_DONE][lane] = 1; sh.sc[SC_ALPHA][lane] = 0; }
            else sh.sc[SC_ALPHA][lane] = -g / h;
          } else {
            sh.sc[SC_CBEST][lane] = c;
            float g0 = sh.sc[SC_G0][lane], lo = sh.sc[SC_LO][lane], hi = sh.sc[SC_HI][lane];
            if (fabsf(g) < 1e-6f * fabsf(g0) || ls == m.ls_iter) sh.isc[I_LSDONE][lane] = 1;
            else {
              if (g < 0) lo = alpha; else hi = alpha;
              float na = alpha - g / h;
              if (hi >= 0 && (na <= lo || na >= hi)) na = 0.5f * (lo + hi);
              else if (hi < 0 && na <= lo) na = 2 * alpha;
              if (fabsf(na - alpha) <= 1e-7f * fabsf(alpha)) sh.isc[I_LSDONE][lane] = 1;
              sh.sc[SC_ALPHA][lane] = na; sh.sc[SC_LO][lane] = lo; sh.sc[SC_HI][lane] = hi;
            }
          } }
        PAR_END
      }
      PAR_BEGIN if (ACTIVE) { float alpha = sh.sc[SC_ALPHA][lane]; ROWS_BEGIN SW(W_LAM, r) += alpha * SW(W_DL, r); ROWS_END } PAR_END
      PAR_BEGIN if (y == 0 && ACTIVE) { sh.isc[I_ITER][lane] = iter + 1; float imp = scale * (sh.sc[SC_COST][lane] - sh.sc[SC_CBEST][lane]); if (imp < m.tolerance) sh.isc[I_DONE][lane] = 1; } PAR_END
    }
    // ---------------- forces at the solution
    PAR_BEGIN int n = MY_N; ROWS_BEGIN float s = SW(S_B, r); for (int j = 0; j < n; j++) s += AM(r, j) * SW(W_LAM, j); SW(W_JAR, r) = s; ROWS_END PAR_END
    PAR_BEGIN ROWS_BEGIN if (IS_HEAD(r)) head_update(m, d, cx, lane, e, r, false); ROWS_END PAR_END
    // ---------------- noslip: inherently sequential Gauss-Seidel over the friction rows (y == 0), on W_F
    if (m.noslip_iterations > 0) {
      PAR_BEGIN if (y == 0) { int n = MY_N; AT(d.niter, 0) = sh.isc[I_ITER][lane];
        for (int it = 0; it < m.noslip_iterations && n > 0; it++) {
          float improvement = 0;
          if (it == 0) for (int i = 0; i < n; i++) improvement += 0.5f * SW(W_F, i) * SW(W_F, i) * SW(S_R, i);
          bool any = false;
          for (int i = 0; i < n; i++) {
            if (SW(S_TYPE, i) < 0.5f) continue;
            any = true;
            float fn = SW(W_F, i), old0 = SW(W_F, i + 1), old1 = SW(W_F, i + 2);
            float res[2], Ac[4], bc[2], v[2];
            for (int rr = 0; rr < 2; rr++) { float s = SW(S_B, i + 1 + rr); for (int j = 0; j < n; j++) s += AM(i + 1 + rr, j) * SW(W_F, j); res[rr] = s; }
            Ac[0] = AM(i + 1, i + 1); Ac[1] = AM(i + 1, i + 2); Ac[2] = Ac[1]; Ac[3] = AM(i + 2, i + 2);
            bc[0] = res[0] - Ac[0] * old0 - Ac[1] * old1; bc[1] = res[1] - Ac[2] * old0 - Ac[3] * old1;
            float fr0 = SW(S_F1, i), fr1 = SW(S_F2, i);
            if (fn < FB_MINVAL) { v[0] = 0; v[1] = 0; }
            else {
              int active = qcqp2(v, Ac, bc, fr0, fr1, fn);
              if (active) { float s = (v[0] / fr0) * (v[0] / fr0) + (v[1] / fr1) * (v[1] / fr1); s = sqrtf(fn * fn / fmaxf(FB_MINVAL, s)); v[0] *= s; v[1] *= s; }
            }
            float d0 = v[0] - old0, d1 = v[1] - old1;
            float change = 0.5f * (d0 * (Ac[0] * d0 + Ac[1] * d1) + d1 * (Ac[2] * d0 + Ac[3] * d1)) + d0 * res[0] + d1 * res[1];
            if (change > 1e-10f) { v[0] = old0; v[1] = old1; change = 0; }
            SW(W_F, i + 1) = v[0]; SW(W_F, i + 2) = v[1];
            improvement -= change;
            i += 2;
          }
          if (!any) break;
          if (improvement * scale < m.noslip_tolerance) break;
        } }
      PAR_END
    }
    PAR_BEGIN ROWS_BEGIN EFC(d.efc_force, r) = SW(W_F, r); ROWS_END if (y == 0) AT(d.niter, 0) = sh.isc[I_ITER][lane]; PAR_END
  }
  // ---------------- qfrc_constraint = J^T f, gathered per dof (race free)
  PAR_BEGIN int n = MY_N;
    for (int k = y; k < m.nv; k += FB_SOLVE_Y) {
      float s = 0;
      for (int r = 0; r < n; r++) {
        float f = SW(W_F, r);
        if (f == 0.0f) continue;
        if (in_chain(m, k, (int)SW(S_LA, r)) || in_chain(m, k, (int)SW(S_LB, r))) s += EJ(d.efc_J, r, k) * f;
      }
      AT(d.qfrc_constraint, k) = s;
    }
  PAR_END
}
