// fb_render.h -- egocentric eye cameras (reference FruitFlyObservables.right_eye / left_eye, fruitfly/fruitfly.py:729-745:
// 32 x 32 RGB images from the cameras `eye_right` / `eye_left` on the head, fruitfly.xml:335-336, fovy set by the task,
// tasks/vision_flight.py:23-24,78-79).  The reference rasterises the scene with MuJoCo's OpenGL renderer; this is a ray
// caster over what the flight arenas contain -- a heightfield terrain (tasks/arenas/hills.py), the ground plane around it and
// a sky -- so pixel parity with the GL image is neither attainable nor attempted; the camera model
// (pinhole, -z forward, +y up, vertical field of view, square image, row 0 at the top) is MuJoCo's and is what the tests check.
//
// One thread block per (env, eye); a thread casts the rays of its pixels: march in steps of half a grid cell until the ray
// is under the bilinear terrain surface, bisect, shade with the headlight (ambient + diffuse along the view direction,
// hills.py:248-250) on a 1 x 1 checker so that optic flow is visible; rays that leave the arena hit the ground plane or the sky.
#pragma once
#include "fb_math.h"

#define FB_MAXCAM 2
struct DevEye {
  int n_cam, size, nrow, ncol;                      // nrow == 0: no heightfield, flat ground at z_offset
  int body[FB_MAXCAM]; float pos[FB_MAXCAM][3]; float quat[FB_MAXCAM][4];    // camera frames in their body's frame
  float tan_half, half_size, z_offset, zfar;
  float sky_top[3], sky_horizon[3], ground[3], ambient, diffuse;
  const float* hfield;                              // [N][nrow * ncol] heights (world units), row = y, col = x
  const float* hmax;                                // [N] highest point of the env's terrain
  const float* cmax;                                // [N][nbr * nbc] highest grid point of every FB_EYE_BLOCK x FB_EYE_BLOCK cell block (incl. its border points)
  int nbr, nbc;
  unsigned char* out;                               // [N][n_cam][size][size][3]
};
#define FB_EYE_BLOCK 8

FB_DEV float eye_height(const DevEye& p, const float* h, float x, float y) {      // bilinear terrain height inside the arena
  const float fx = (x + p.half_size) * (float)(p.ncol - 1) / (2.0f * p.half_size), fy = (y + p.half_size) * (float)(p.nrow - 1) / (2.0f * p.half_size);
  int ix = (int)floorf(fx), iy = (int)floorf(fy);
  ix = ix < 0 ? 0 : (ix > p.ncol - 2 ? p.ncol - 2 : ix); iy = iy < 0 ? 0 : (iy > p.nrow - 2 ? p.nrow - 2 : iy);
  const float tx = fx - (float)ix, ty = fy - (float)iy;
  const float* r0 = h + (size_t)iy * p.ncol + ix; const float* r1 = r0 + p.ncol;
  return (r0[0] * (1.0f - tx) + r0[1] * tx) * (1.0f - ty) + (r1[0] * (1.0f - tx) + r1[1] * tx) * ty + p.z_offset;
}
FB_DEV unsigned char eye_u8(float v) { v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); return (unsigned char)(v * 255.0f + 0.5f); }

// colour of the ray from `o` along the unit vector `dir`
FB_DEV void eye_cast(const DevEye& p, const float* h, const float* cm, float hmax, V3 o, V3 dir, float* rgb) {
  float t_hit = -1.0f; V3 n = v3(0, 0, 1);
  if (p.nrow > 0) {
    const float S = p.half_size, cell = 2.0f * S / (float)(p.ncol - 1);
    // the part of the ray above the arena footprint and below the highest terrain point
    float t0 = 0.0f, t1 = p.zfar;
    const float dd[2] = {dir.x, dir.y}, oo[2] = {o.x, o.y};
    for (int a = 0; a < 2; a++) {
      if (fabsf(dd[a]) < 1e-9f) { if (fabsf(oo[a]) > S) t1 = -1.0f; }
      else { float ta = (-S - oo[a]) / dd[a], tb = (S - oo[a]) / dd[a]; if (ta > tb) { float s = ta; ta = tb; tb = s; } t0 = fmaxf(t0, ta); t1 = fminf(t1, tb); }
    }
    const float ztop = hmax + p.z_offset;
    if (o.z > ztop) { if (dir.z >= 0.0f) t1 = -1.0f; else t0 = fmaxf(t0, (ztop - o.z) / dir.z); }
    else if (dir.z > 0.0f) t1 = fminf(t1, (ztop - o.z) / dir.z);
    if (t1 > t0) {
      const float hstep = 0.5f * cell / fmaxf(sqrtf(dir.x * dir.x + dir.y * dir.y), 0.05f);       // <= half a cell sideways per step
      // samples at t0 + k hstep.  The bilinear surface inside a block of FB_EYE_BLOCK^2 cells never exceeds the block's highest grid
      // point, so while the ray is above that height from here to where it leaves the block, every sample on the way is above the
      // surface and is skipped (conservative: the first sample found below the surface is the one plain marching finds).
      const float bw = (float)FB_EYE_BLOCK * cell, inv_cell = 1.0f / cell;
      int k = 0; bool below = false; float t = t0;
      for (int it = 0; it < 8192; it++) {
        t = t0 + (float)k * hstep;
        if (t > t1) break;
        const V3 q = o + dir * t;
        if (cm) {
          int ix = (int)floorf((q.x + S) * inv_cell), iy = (int)floorf((q.y + S) * inv_cell);
          ix = ix < 0 ? 0 : (ix > p.ncol - 2 ? p.ncol - 2 : ix); iy = iy < 0 ? 0 : (iy > p.nrow - 2 ? p.nrow - 2 : iy);
          const int bx = ix / FB_EYE_BLOCK, by = iy / FB_EYE_BLOCK;
          const float top = cm[by * p.nbc + bx] + p.z_offset;
          if (q.z > top) {
            float te = t1;                           // where the ray leaves the block's footprint
            if (dir.x > 1e-9f) te = fminf(te, ((float)(bx + 1) * bw - S - o.x) / dir.x); else if (dir.x < -1e-9f) te = fminf(te, ((float)bx * bw - S - o.x) / dir.x);
            if (dir.y > 1e-9f) te = fminf(te, ((float)(by + 1) * bw - S - o.y) / dir.y); else if (dir.y < -1e-9f) te = fminf(te, ((float)by * bw - S - o.y) / dir.y);
            if (te > t && o.z + dir.z * te > top) { int kn = (int)floorf((te - t0) / hstep) + 1; k = kn > k ? kn : k + 1; continue; }
          }
        }
        if (q.z < eye_height(p, h, q.x, q.y)) { below = true; break; }
        k++;
      }
      const float tp = k > 0 ? t0 + (float)(k - 1) * hstep : t0;
      if (below) {
        float lo = tp, hi = t;
        for (int it = 0; it < 8; it++) { const float mid = 0.5f * (lo + hi); const V3 q = o + dir * mid; if (q.z < eye_height(p, h, q.x, q.y)) hi = mid; else lo = mid; }
        t_hit = hi;
        const V3 q = o + dir * t_hit; const float e = 0.5f * cell;
        n = normalized(v3(eye_height(p, h, q.x - e, q.y) - eye_height(p, h, q.x + e, q.y), eye_height(p, h, q.x, q.y - e) - eye_height(p, h, q.x, q.y + e), 2.0f * e));
      }
    }
  }
  if (t_hit < 0.0f && dir.z < -1e-9f) { const float tg = (p.z_offset - o.z) / dir.z; if (tg > 0.0f && tg <= p.zfar) { t_hit = tg; n = v3(0, 0, 1); } }   // ground plane
  if (t_hit < 0.0f) {                                  // sky: horizon colour blending into the zenith colour
    const float s = dir.z > 0.0f ? dir.z : 0.0f;
    for (int c = 0; c < 3; c++) rgb[c] = p.sky_horizon[c] + (p.sky_top[c] - p.sky_horizon[c]) * s;
    return;
  }
  const V3 q = o + dir * t_hit;
  const int cx = (int)floorf(q.x), cy = (int)floorf(q.y);
  const float tex = ((cx + cy) & 1) ? 1.0f : 0.75f;
  float lam = -(dir.x * n.x + dir.y * n.y + dir.z * n.z); lam = lam < 0.0f ? 0.0f : lam;
  const float shade = (p.ambient + p.diffuse * lam) * tex;
  for (int c = 0; c < 3; c++) rgb[c] = p.ground[c] * shade;
}
// pixel (i, j) of eye `cam` of env e (row 0 = top of the image)
FB_DEV void eye_pixel(const DevData& d, const DevEye& p, int e, int cam, int i, int j) {
  const int b = p.body[cam];
  const M3 Rb = ld9(d.xmat, b, d, e);
  const V3 o = ld3(d.xpos, b, d, e) + v3(AT(d.ref, 0), AT(d.ref, 1), AT(d.ref, 2)) + mul(Rb, v3(p.pos[cam][0], p.pos[cam][1], p.pos[cam][2]));
  const M3 Rc = q2m(q4(p.quat[cam][0], p.quat[cam][1], p.quat[cam][2], p.quat[cam][3]));
  const float u = (2.0f * ((float)j + 0.5f) / (float)p.size - 1.0f) * p.tan_half, v = (1.0f - 2.0f * ((float)i + 0.5f) / (float)p.size) * p.tan_half;
  const V3 dir = mul(Rb, mul(Rc, normalized(v3(u, v, -1.0f))));
  float rgb[3];
  eye_cast(p, p.nrow > 0 ? p.hfield + (size_t)e * p.nrow * p.ncol : nullptr, (p.nrow > 0 && p.cmax) ? p.cmax + (size_t)e * p.nbr * p.nbc : nullptr,
           p.nrow > 0 ? p.hmax[e] : 0.0f, o, dir, rgb);
  unsigned char* px = p.out + ((((size_t)e * p.n_cam + cam) * p.size + i) * p.size + j) * 3;
  px[0] = eye_u8(rgb[0]); px[1] = eye_u8(rgb[1]); px[2] = eye_u8(rgb[2]);
}
