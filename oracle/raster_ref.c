/*
 * ORACLE (test infrastructure only — never linked into or called by the product path).
 *
 * CPU restatement of the OpenGL part of the reference's novel-view warp, which cannot run here (no moderngl / EGL / GL
 * driver) — "parity unpinned" for exactly this file: the GL rasteriser is third-party (driver) arithmetic.
 *   reference: rgbd_3d/moderngl_renderer.py:197-202  GL state: depth '<', no culling, no blending, CCW front faces
 *              rgbd_3d/moderngl_renderer.py:302-316  per source mesh: clear, draw TRIANGLES, aggregation compute pass
 *              rgbd_3d/shaders/aggregation.vsh:19-29  vertex stage
 *              rgbd_3d/shaders/aggregation.fsh:19-52  fragment stage (view-angle weight, flags, back faces)
 *              rgbd_3d/shaders/aggregation.csh:12-44  cross-view accumulation, "farther wins" for low-confidence pixels
 *              rgbd_3d/shaders/clear.csh
 *              rgbd_3d/moderngl_renderer.py:11-148 + shaders/simple.{vsh,fsh}  SimpleRenderer (training-pair warp):
 *                 same GL state, no back-face discard, alpha = (edge flag varying > 0.999 ? 0 : 1)
 * Rasterisation rules that GL leaves to the implementation are fixed here (and mirrored by the CUDA kernels):
 *   - vertex stage in fp32, clip = (P*MV) * v with P*MV formed on the host in double and rounded to fp32
 *   - near-plane clipping in clip space (z + w >= 0), polygons split as a fan
 *   - window coordinates snapped to 1/256 pixel, pixel centres at +0.5, coverage from exact 64-bit edge functions with
 *     a top-left style tie rule; sequential draw order, strict '<' depth test on fp32 window depth in (0,1)
 *   - perspective-correct varyings, NEAREST / clamp-to-edge texture fetch, gl_FrontFacing from the snapped signed area
 * Compile with -ffp-contract=off (the CUDA side uses -fmad=false) so both sides round identically.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float clip[4];     /* clip-space position */
  float pos[3];      /* world position (varying) */
  float nrm[3];      /* normalised vertex normal (varying) */
  float uv[2];
  float edge, pad, ero;   /* flag bits as 0/1 varyings (aggregation.vsh:26-28) */
} Vtx;

static void lerp_vtx(const Vtx* a, const Vtx* b, float t, Vtx* o) {
  const float* pa = (const float*)a;
  const float* pb = (const float*)b;
  float* po = (float*)o;
  for (int i = 0; i < (int)(sizeof(Vtx) / sizeof(float)); ++i) po[i] = pa[i] + (pb[i] - pa[i]) * t;
}

/* clip a triangle against z + w >= 0; returns polygon vertex count (0, 3 or 4) */
static int clip_near(const Vtx* in, Vtx* out) {
  float d[3];
  int n = 0;
  for (int i = 0; i < 3; ++i) d[i] = in[i].clip[2] + in[i].clip[3];
  if (d[0] >= 0.f && d[1] >= 0.f && d[2] >= 0.f) {
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
    return 3;
  }
  if (d[0] < 0.f && d[1] < 0.f && d[2] < 0.f) return 0;
  for (int i = 0; i < 3; ++i) {
    const int j = (i + 1) % 3;
    if (d[i] >= 0.f) out[n++] = in[i];
    if ((d[i] >= 0.f) != (d[j] >= 0.f)) {
      const float t = d[i] / (d[i] - d[j]);
      lerp_vtx(&in[i], &in[j], t, &out[n++]);
    }
  }
  return n;
}

static float shade_weight(const float* pos, const float* nrm, const float* cam, float edge, float pad, float ero) {
  /* aggregation.fsh:28-49 */
  float dir[3] = {cam[0] - pos[0], cam[1] - pos[1], cam[2] - pos[2]};
  float dl = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  float nl = sqrtf(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
  float dt = (dir[0] * nrm[0] + dir[1] * nrm[1] + dir[2] * nrm[2]) / (dl * nl);
  float w = dt < 0.f ? 0.f : (dt > 1.f ? 1.f : dt);
  w = acosf(w);
  w = fmaxf(-w * 20.f, -50.f);
  w = expf(w);
  w = fmaxf(w, 1e-4f);
  if (!(ero < 0.999f)) w *= 1e-8f;
  if (pad > 0.001f || edge > 0.999f) w = 1e-16f;
  return fmaxf(w, 1e-16f);
}

static void raster_tri(const Vtx* v, int S, const float* tex, int T, const float* cam, float* color_fb, float* depth_fb, int simple) {
  int64_t X[3], Y[3];
  float zw[3], iw[3];
  for (int i = 0; i < 3; ++i) {
    const float w = v[i].clip[3];
    const float xn = v[i].clip[0] / w, yn = v[i].clip[1] / w, zn = v[i].clip[2] / w;
    const float xw = (xn * 0.5f + 0.5f) * (float)S, yw = (yn * 0.5f + 0.5f) * (float)S;
    X[i] = (int64_t)floorf(xw * 256.f + 0.5f);
    Y[i] = (int64_t)floorf(yw * 256.f + 0.5f);
    zw[i] = zn * 0.5f + 0.5f;
    iw[i] = 1.0f / w;
  }
  int64_t area = (X[1] - X[0]) * (Y[2] - Y[0]) - (Y[1] - Y[0]) * (X[2] - X[0]);
  if (area == 0) return;
  const int front = area > 0;              /* CCW in window space (y up) */
  const int64_t sgn = front ? 1 : -1;
  int64_t minx = X[0], maxx = X[0], miny = Y[0], maxy = Y[0];
  for (int i = 1; i < 3; ++i) {
    if (X[i] < minx) minx = X[i];
    if (X[i] > maxx) maxx = X[i];
    if (Y[i] < miny) miny = Y[i];
    if (Y[i] > maxy) maxy = Y[i];
  }
  int64_t px0 = (minx - 128 + 255) / 256, px1 = (maxx - 128) / 256, py0 = (miny - 128 + 255) / 256, py1 = (maxy - 128) / 256;
  if (minx - 128 < 0) px0 = 0;
  if (miny - 128 < 0) py0 = 0;
  if (px0 < 0) px0 = 0;
  if (py0 < 0) py0 = 0;
  if (px1 > S - 1) px1 = S - 1;
  if (py1 > S - 1) py1 = S - 1;
  const float farea = (float)(sgn * area);
  for (int64_t py = py0; py <= py1; ++py) {
    for (int64_t px = px0; px <= px1; ++px) {
      const int64_t cx = px * 256 + 128, cy = py * 256 + 128;
      int64_t E[3];
      int inside = 1;
      for (int i = 0; i < 3 && inside; ++i) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;   /* edge opposite vertex i: a -> b */
        int64_t dx = (X[b] - X[a]) * sgn, dy = (Y[b] - Y[a]) * sgn;
        int64_t e = ((X[b] - X[a]) * (cy - Y[a]) - (Y[b] - Y[a]) * (cx - X[a])) * sgn;
        const int tie_ok = (dy > 0) || (dy == 0 && dx < 0);
        if (e < 0 || (e == 0 && !tie_ok)) inside = 0;
        E[i] = e;
      }
      if (!inside) continue;
      const float l0 = (float)E[0] / farea, l1 = (float)E[1] / farea, l2 = (float)E[2] / farea;
      const float z = (l0 * zw[0] + l1 * zw[1]) + l2 * zw[2];
      if (!(z > 0.f && z < 1.f)) continue;
      const float b0 = l0 * iw[0], b1 = l1 * iw[1], b2 = l2 * iw[2];
      const float bs = (b0 + b1) + b2;
      const float c0 = b0 / bs, c1 = b1 / bs, c2 = b2 / bs;
      const size_t idx = (size_t)py * S + (size_t)px;
      const float pad = (c0 * v[0].pad + c1 * v[1].pad) + c2 * v[2].pad;
      if (!simple && !front && pad > 0.001f) continue;      /* discard: no depth write (aggregation.fsh:23) */
      if (!(z < depth_fb[idx])) continue;                   /* depth func '<' */
      depth_fb[idx] = z;
      float* o = color_fb + idx * 4;
      if (!front) { o[0] = o[1] = o[2] = o[3] = 0.f; continue; }
      float pos[3], nrm[3];
      for (int k = 0; k < 3; ++k) {
        pos[k] = (c0 * v[0].pos[k] + c1 * v[1].pos[k]) + c2 * v[2].pos[k];
        nrm[k] = (c0 * v[0].nrm[k] + c1 * v[1].nrm[k]) + c2 * v[2].nrm[k];
      }
      const float uu = (c0 * v[0].uv[0] + c1 * v[1].uv[0]) + c2 * v[2].uv[0];
      const float vv = (c0 * v[0].uv[1] + c1 * v[1].uv[1]) + c2 * v[2].uv[1];
      const float edge = (c0 * v[0].edge + c1 * v[1].edge) + c2 * v[2].edge;
      const float ero = (c0 * v[0].ero + c1 * v[1].ero) + c2 * v[2].ero;
      int tx = (int)floorf(uu * (float)T), ty = (int)floorf(vv * (float)T);
      tx = tx < 0 ? 0 : (tx > T - 1 ? T - 1 : tx);
      ty = ty < 0 ? 0 : (ty > T - 1 ? T - 1 : ty);
      const float* tc = tex + ((size_t)ty * T + tx) * 3;
      o[0] = tc[0]; o[1] = tc[1]; o[2] = tc[2];
      o[3] = simple ? (edge > 0.999f ? 0.f : 1.f)          /* simple.fsh:19 */
                    : shade_weight(pos, nrm, cam, edge, pad, ero);
    }
  }
}

/* Draw one source mesh into a cleared framebuffer (color rgba = 0, depth = 1).
 *   verts  [V][9] float32: position(3) normal(3) uv(2) flag(1)   (moderngl_renderer.py:284-289)
 *   faces  [F][3] uint32;  tex [T][T][3] float32 (row 0 = image top, sampled at the unflipped uv)
 *   mvp    [16] float32 row-major (P*MV);  cam [3] = source camera position in world space (u_sample_camera) */
void raster_draw_mesh(const float* verts, int V, const uint32_t* faces, int F, const float* tex, int T, const float* mvp,
                      const float* cam, int S, float* color_fb, float* depth_fb) {
  (void)V;
  for (size_t i = 0; i < (size_t)S * S; ++i) { depth_fb[i] = 1.0f; }
  memset(color_fb, 0, sizeof(float) * 4 * (size_t)S * S);
  for (int f = 0; f < F; ++f) {
    Vtx in[3], poly[4];
    for (int k = 0; k < 3; ++k) {
      const float* a = verts + (size_t)faces[f * 3 + k] * 9;
      for (int r = 0; r < 4; ++r)
        in[k].clip[r] = ((mvp[r * 4 + 0] * a[0] + mvp[r * 4 + 1] * a[1]) + mvp[r * 4 + 2] * a[2]) + mvp[r * 4 + 3];
      in[k].pos[0] = a[0]; in[k].pos[1] = a[1]; in[k].pos[2] = a[2];
      const float nl = sqrtf((a[3] * a[3] + a[4] * a[4]) + a[5] * a[5]);     /* v_normal = normalize(i_normal) */
      in[k].nrm[0] = a[3] / nl; in[k].nrm[1] = a[4] / nl; in[k].nrm[2] = a[5] / nl;
      in[k].uv[0] = a[6]; in[k].uv[1] = a[7];
      const int flag = (int)a[8];
      in[k].edge = (float)(flag & 1); in[k].pad = (float)((flag >> 1) & 1); in[k].ero = (float)((flag >> 2) & 1);
    }
    const int n = clip_near(in, poly);
    if (n >= 3) {
      Vtx t0[3] = {poly[0], poly[1], poly[2]};
      raster_tri(t0, S, tex, T, cam, color_fb, depth_fb, 0);
    }
    if (n == 4) {
      Vtx t1[3] = {poly[0], poly[2], poly[3]};
      raster_tri(t1, S, tex, T, cam, color_fb, depth_fb, 0);
    }
  }
}

/* SimpleRenderer.render for one modelview (moderngl_renderer.py:94-146): clear, draw TRIANGLES with simple.vsh / simple.fsh.
 *   verts [V][6] float32: position(3) uv(2) flag(1)  (moderngl_renderer.py:113-117); v_is_edge = mod(flag, 2) */
void raster_draw_simple(const float* verts, int V, const uint32_t* faces, int F, const float* tex, int T, const float* mvp,
                        int S, float* color_fb, float* depth_fb) {
  (void)V;
  const float cam[3] = {0.f, 0.f, 0.f};
  for (size_t i = 0; i < (size_t)S * S; ++i) { depth_fb[i] = 1.0f; }
  memset(color_fb, 0, sizeof(float) * 4 * (size_t)S * S);
  for (int f = 0; f < F; ++f) {
    Vtx in[3], poly[4];
    for (int k = 0; k < 3; ++k) {
      const float* a = verts + (size_t)faces[f * 3 + k] * 6;
      memset(&in[k], 0, sizeof(Vtx));
      for (int r = 0; r < 4; ++r)
        in[k].clip[r] = ((mvp[r * 4 + 0] * a[0] + mvp[r * 4 + 1] * a[1]) + mvp[r * 4 + 2] * a[2]) + mvp[r * 4 + 3];
      in[k].pos[0] = a[0]; in[k].pos[1] = a[1]; in[k].pos[2] = a[2];
      in[k].uv[0] = a[3]; in[k].uv[1] = a[4];
      in[k].edge = fmodf(a[5], 2.0f);
    }
    const int n = clip_near(in, poly);
    if (n >= 3) {
      Vtx t0[3] = {poly[0], poly[1], poly[2]};
      raster_tri(t0, S, tex, T, cam, color_fb, depth_fb, 1);
    }
    if (n == 4) {
      Vtx t1[3] = {poly[0], poly[2], poly[3]};
      raster_tri(t1, S, tex, T, cam, color_fb, depth_fb, 1);
    }
  }
}

/* aggregation.csh:12-44 applied to every pixel: accumulate one rendered source view into the aggregation images. */
void raster_aggregate(const float* color_fb, const float* depth_fb, int S, float* agg_color /*[S*S][4]*/,
                      float* agg_depth /*[S*S][2]*/, float* agg_mask /*[S*S][2]*/) {
  for (size_t i = 0; i < (size_t)S * S; ++i) {
    const float* c = color_fb + i * 4;
    const float depth = depth_fb[i];
    const float wc = c[3];
    const float wd = c[3] > 1e-14f ? 1.0f : (c[3] > 0.0f ? 1e-8f : 0.0f);
    const float mc = c[3] > 1e-6f ? 1.0f : 0.0f;
    const float md = c[3] > 1e-14f ? 1.0f : 0.0f;
    float* pc = agg_color + i * 4;
    float* pd = agg_depth + i * 2;
    float* pm = agg_mask + i * 2;
    if (fabsf(pd[1] - 1e-8f) < 1e-8f && fabsf(wd - 1e-8f) < 1e-8f) {
      if (depth * 1e-8f > pd[0]) {
        pd[0] = depth * 1e-8f;
        pd[1] = 1e-8f;
        pc[0] = c[0] * wc; pc[1] = c[1] * wc; pc[2] = c[2] * wc; pc[3] = wc;
      }
    } else {
      pd[0] += depth * wd; pd[1] += wd;
      pc[0] += c[0] * wc; pc[1] += c[1] * wc; pc[2] += c[2] * wc; pc[3] += wc;
    }
    pm[0] += md; pm[1] += mc;
  }
}
