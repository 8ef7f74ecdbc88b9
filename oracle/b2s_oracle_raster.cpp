// TEST INFRASTRUCTURE ONLY.  CPU restatement of the camera path the reference gets from SAPIEN's Vulkan rasteriser:
// `camera_group.take_picture()` (mani_skill/utils/structs/render_camera.py:269-273) with the "minimal" shader pack whose
// render targets are `Color` rgba8 and `PositionSegmentation` 4 x int16 = (x, y, z in mm, OpenGL camera frame,
// segmentation id) (mani_skill/render/shaders.py:68-84, docs/source/user_guide/concepts/observation.md).
//
// SAPIEN's renderer is not in /root/reference and cannot run here, and the reference ships no reference images or
// masks: PARITY UNPINNED for absolute pixel values.  What this oracle pins is the CUDA rasteriser's own definition:
//   * pinhole camera, sapien camera frame (x forward, y left, z up), pixel centres at (u + 0.5, v + 0.5), v = 0 at the top
//   * convex hulls, and boxes whose eight corners are in front of the near plane, are triangle meshes: vertices projected to the
//     screen, sample inside iff the three edge functions are >= 0 after orienting the triangle counter-clockwise, 1/depth
//     interpolated linearly in screen space, quantised to a 23-bit reversed-z key; back faces culled
//   * flat faces tested per pixel (half-spaces; the camera-facing faces of a box that reaches behind the near plane; camera-facing box
//     faces whose screen rectangle exceeds PATCH_PIXELS, which are then not rasterised): 1/depth of the plane hit is linear in the ray
//     direction (no division), the hit is inside the face iff |o_b inv + d_b| <= h_b inv (1 + 1e-5) for the two in-plane axes;
//     spheres: 1 / ray parameter
//   * the smallest (key << 9 | box face << 6 | visual index) wins, depth = 1 / dequantised(1/depth); flat faces (boxes: the face of the
//     winning triangle / the face the ray enters through; half-spaces) have one shaded colour per image; spheres: normal = (hit - centre) / r;
//     hulls: normal from the depth neighbourhood
//   * segmentation = per_scene_id of the winning visual, 0 = background; position = hit point in mm (round to nearest even)
// Plain loops, float32, compiled with -ffp-contract=off; the CUDA translation unit is compiled with -fmad=false, so the
// integer outputs (segmentation, position) are expected to agree bit for bit.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct F3 {
  float x, y, z;
};
inline F3 f3(float x, float y, float z) { F3 r = {x, y, z}; return r; }
inline F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline F3 operator-(F3 a) { return f3(-a.x, -a.y, -a.z); }
inline F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
inline float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 cross(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float norm(F3 a) { return sqrtf(dot(a, a)); }

struct Q {
  float w, x, y, z;
};
struct P7 {
  F3 p;
  Q q;
};
inline Q qmul(Q a, Q b) {
  Q r = {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
         a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
  return r;
}
inline Q qnorm(Q q) {
  float s = 1.f / sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  Q r = {q.w * s, q.x * s, q.y * s, q.z * s};
  return r;
}
inline F3 qrot(Q q, F3 v) {
  F3 u = f3(q.x, q.y, q.z);
  F3 t = cross(u, v) * 2.f;
  return v + t * q.w + cross(u, t);
}
inline void qmat(Q q, float* m) {
  float w = q.w, x = q.x, y = q.y, z = q.z;
  m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
inline P7 pmul(P7 a, P7 b) {
  P7 c;
  c.p = a.p + qrot(a.q, b.p);
  c.q = qmul(a.q, b.q);
  return c;
}
inline P7 p7(const float* f) {
  P7 p;
  p.p = f3(f[0], f[1], f[2]);
  Q q = {f[3], f[4], f[5], f[6]};
  p.q = q;
  return p;
}
inline P7 ident() {
  P7 p;
  p.p = f3(0, 0, 0);
  Q q = {1, 0, 0, 0};
  p.q = q;
  return p;
}

const float DEPTH_MAX = 8388607.0f;
const int KEY_SHIFT = 9;
const int PATCH_PIXELS = 1024;  // maniskill_b200/csrc/b2s_raster.cuh B2S_PATCH_PIXELS
inline unsigned make_key(unsigned dk, int face, int v) { return (dk << KEY_SHIFT) | ((unsigned)face << 6) | (unsigned)v; }
// reversed-z quantisation on 1/depth; per-camera constants computed once (maniskill_b200/csrc/b2s_raster.cuh DepthMap)
struct DepthMap {
  float invn, invf, range, scale, inv_max;
};
inline DepthMap depth_map(float nearp, float farp) {
  DepthMap m;
  m.invn = 1.0f / nearp;
  m.invf = 1.0f / farp;
  m.range = m.invn - m.invf;
  m.scale = 1.0f / m.range;
  m.inv_max = 1.0f / DEPTH_MAX;
  return m;
}
inline bool inv_depth_in_range(float inv, const DepthMap& m) { return inv < m.invn && inv > m.invf; }
inline unsigned depth_key_inv(float inv, const DepthMap& m) {
  float t = (inv - m.invf) * m.scale;
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  float q = DEPTH_MAX - t * DEPTH_MAX;
  return (unsigned)q;
}
inline float key_depth(unsigned k, const DepthMap& m) {
  float t = (DEPTH_MAX - (float)k) * m.inv_max;
  float inv = m.invf + t * m.range;
  return 1.0f / inv;
}
inline bool ray_sphere(F3 o, F3 dv, float r, float& t_hit) {
  float a = dot(dv, dv), b = dot(o, dv), c = dot(o, o) - r * r;
  float disc = b * b - a * c;
  if (disc < 0.0f) return false;
  float t = (-b - sqrtf(disc)) / a;
  if (t <= 0.0f) return false;
  t_hit = t;
  return true;
}
inline uint8_t to_u8(float x) {
  float c = fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f + 0.5f;
  return (uint8_t)c;
}
inline int16_t to_mm(float x) {
  float v = x * 1000.0f;
  v = fminf(fmaxf(v, -32768.0f), 32767.0f);
  return (int16_t)rintf(v);
}
inline unsigned shade_rgb(F3 n_world, const float* col) {
  const float kk = 0.57735026f;
  F3 l1 = f3(-kk, -kk, kk), l2 = f3(0.0f, 0.0f, 1.0f);
  float w = 0.3f + 0.5f * fmaxf(dot(n_world, l1), 0.0f) + 0.5f * fmaxf(dot(n_world, l2), 0.0f);
  return (unsigned)to_u8(col[0] * w) | ((unsigned)to_u8(col[1] * w) << 8) | ((unsigned)to_u8(col[2] * w) << 16);
}
inline F3 mulm(const float* m, F3 v) { return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z); }
inline F3 tmulm(const float* m, F3 v) { return f3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z, m[2] * v.x + m[5] * v.y + m[8] * v.z); }

}  // namespace

extern "C" {

// Renders one camera of one sub-scene.
//   vis_*: render-shape table (same arrays the C-ABI B2SVisualTable carries); per-env size/pose already resolved by the caller
//   vert_* / tri_*: indexed triangle geometry of the hull and box visuals (box vertices = corners of the unit cube)
//   body:  [n_rows*13] float32 rows of this sub-scene
//   cam:   w h fx fy cx cy near far mount_row local_pose(7)  as 16 floats
//   out:   color [h*w*4] uint8, posseg [h*w*4] int16
void b2o_render_one(int n_vis, const int* vis_type, const int* vis_row, const float* vis_pose, const float* vis_size,
                    const float* vis_color, const int* vis_seg, int n_vert, const float* vert_local, const int* vert_vis, int n_tri,
                    const int* tri_idx, const int* tri_vis, int n_rows, const float* body, const float* cam, uint8_t* color, int16_t* posseg) {
  const int W = (int)cam[0], H = (int)cam[1];
  const float fx = cam[2], fy = cam[3], cx = cam[4], cy = cam[5], nearp = cam[6], farp = cam[7];
  const DepthMap dm = depth_map(nearp, farp);
  const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy;
  const int mount = (int)cam[8];
  const unsigned NO_HIT = 0xFFFFFFFFu;
  auto body_pose = [&](int row) {
    if (row < 0) return ident();
    return p7(body + (size_t)row * 13);
  };
  P7 Xc = pmul(body_pose(mount), p7(cam + 9));
  Xc.q = qnorm(Xc.q);
  float Rc[9];
  qmat(Xc.q, Rc);
  std::vector<float> vR(n_vis * 9), vt(n_vis * 3), vo(n_vis * 3);
  struct Patch { int x0, x1, y0, y1; float cc; int v, face, bounded; };
  std::vector<Patch> patches;
  std::vector<unsigned> face_patch(n_vis, 0u);
  std::vector<unsigned> face_rgb(n_vis * 6);
  std::vector<int> vmode(n_vis, 1);  // 0 rasterised, 1 analytic
  for (int v = 0; v < n_vis; v++) {
    P7 Xv = pmul(body_pose(vis_row[v]), p7(vis_pose + 7 * v));
    Xv.q = qnorm(Xv.q);
    float Rv[9];
    qmat(Xv.q, Rv);
    float* Rm = &vR[v * 9];
    // Rcv = Rc^T Rv
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rm[3 * i + j] = Rc[i] * Rv[j] + Rc[3 + i] * Rv[3 + j] + Rc[6 + i] * Rv[6 + j];
    F3 t = tmulm(Rc, Xv.p - Xc.p);
    vt[v * 3] = t.x; vt[v * 3 + 1] = t.y; vt[v * 3 + 2] = t.z;
    for (int f = 0; f < 6; f++) {
      const int a = f >> 1;
      const float sgn = (f & 1) ? 1.0f : -1.0f;
      face_rgb[v * 6 + f] = shade_rgb(f3(Rv[a] * sgn, Rv[3 + a] * sgn, Rv[6 + a] * sgn), vis_color + 4 * v);
    }
    vo[v * 3] = -(Rm[0] * t.x + Rm[3] * t.y + Rm[6] * t.z);
    vo[v * 3 + 1] = -(Rm[1] * t.x + Rm[4] * t.y + Rm[7] * t.z);
    vo[v * 3 + 2] = -(Rm[2] * t.x + Rm[5] * t.y + Rm[8] * t.z);
    const int ty = vis_type[v];
    if (ty == 4) vmode[v] = 0;
    if (ty == 1) {
      bool behind = false;
      for (int c = 0; c < 8 && !behind; c++) {
        F3 l = f3((c & 1) ? vis_size[3 * v] : -vis_size[3 * v], (c & 2) ? vis_size[3 * v + 1] : -vis_size[3 * v + 1],
                  (c & 4) ? vis_size[3 * v + 2] : -vis_size[3 * v + 2]);
        F3 pc = mulm(Rm, l) + t;
        if (pc.x <= nearp) behind = true;
      }
      if (!behind) vmode[v] = 0;
    }
  }
  for (int v = 0; v < n_vis; v++)
    for (int f = 0; f < 6; f++) {
      const int a = f >> 1, ty = vis_type[v];
      const float sgn = (f & 1) ? 1.0f : -1.0f;
      float ha;
      if (ty == 0) { if (f != 1) continue; ha = 0.0f; }
      else if (ty == 1) ha = vis_size[3 * v + a];
      else continue;
      const float oa = vo[v * 3 + a];
      if (!(oa * sgn > ha)) continue;
      int rx0 = 0, rx1 = W - 1, ry0 = 0, ry1 = H - 1;
      if (ty == 1 && vmode[v] == 0) {
        const int b = a == 2 ? 0 : a + 1, c = b == 2 ? 0 : b + 1;
        float mnx = 1e30f, mxx = -1e30f, mny = 1e30f, mxy = -1e30f;
        const float* Rm = &vR[v * 9];
        for (int k = 0; k < 4; k++) {
          float l[3];
          l[a] = sgn * ha; l[b] = (k & 1) ? vis_size[3 * v + b] : -vis_size[3 * v + b]; l[c] = (k & 2) ? vis_size[3 * v + c] : -vis_size[3 * v + c];
          const float px_ = Rm[0] * l[0] + Rm[1] * l[1] + Rm[2] * l[2] + vt[v * 3];
          const float py_ = Rm[3] * l[0] + Rm[4] * l[1] + Rm[5] * l[2] + vt[v * 3 + 1];
          const float pz_ = Rm[6] * l[0] + Rm[7] * l[1] + Rm[8] * l[2] + vt[v * 3 + 2];
          const float u = cx - fx * py_ / px_, w = cy - fy * pz_ / px_;
          mnx = fminf(mnx, u); mxx = fmaxf(mxx, u); mny = fminf(mny, w); mxy = fmaxf(mxy, w);
        }
        rx0 = std::max(0, (int)floorf(mnx) - 1); rx1 = std::min(W - 1, (int)ceilf(mxx) + 1);
        ry0 = std::max(0, (int)floorf(mny) - 1); ry1 = std::min(H - 1, (int)ceilf(mxy) + 1);
        if (rx0 > rx1 || ry0 > ry1 || (rx1 - rx0 + 1) * (ry1 - ry0 + 1) <= PATCH_PIXELS) continue;
      }
      if (ty == 1) face_patch[v] |= 1u << f;
      Patch P = {rx0, rx1, ry0, ry1, 1.0f / (sgn * ha - oa), v, f, ty == 1};
      patches.push_back(P);
    }
  // vertices: screen x, y, 1/depth (0 = not in front of the near plane)
  std::vector<float> vert((size_t)n_vert * 3 + 3);
  for (int i = 0; i < n_vert; i++) {
    const int v = vert_vis[i];
    float lx = vert_local[3 * i], ly = vert_local[3 * i + 1], lz = vert_local[3 * i + 2];
    if (vis_type[v] == 1) { lx = lx * vis_size[3 * v]; ly = ly * vis_size[3 * v + 1]; lz = lz * vis_size[3 * v + 2]; }
    const float* Rm = &vR[v * 9];
    float xc = Rm[0] * lx + Rm[1] * ly + Rm[2] * lz + vt[v * 3];
    float yc = Rm[3] * lx + Rm[4] * ly + Rm[5] * lz + vt[v * 3 + 1];
    float zc = Rm[6] * lx + Rm[7] * ly + Rm[8] * lz + vt[v * 3 + 2];
    float* o = &vert[(size_t)3 * i];
    if (xc <= nearp) { o[0] = o[1] = o[2] = 0.0f; continue; }
    float inv = 1.0f / xc;
    o[0] = cx - fx * yc * inv;
    o[1] = cy - fy * zc * inv;
    o[2] = inv;
  }
  const int npix = W * H;
  std::vector<unsigned> zkey(npix, NO_HIT);
  for (int t = 0; t < n_tri; t++) {
    int v = tri_vis[t];
    if (vmode[v] != 0) continue;
    int face = 0;
    if (vis_type[v] == 1) {  // the face of the unit cube this triangle lies in
      const float* a = vert_local + 3 * (size_t)tri_idx[3 * t];
      const float* b = vert_local + 3 * (size_t)tri_idx[3 * t + 1];
      const float* c = vert_local + 3 * (size_t)tri_idx[3 * t + 2];
      for (int k = 0; k < 3; k++)
        if (a[k] == b[k] && b[k] == c[k]) face = 2 * k + (a[k] > 0.0f ? 1 : 0);
      if ((face_patch[v] >> face) & 1u) continue;  // tested per pixel below
    }
    float px[3], py[3], pd[3];
    bool ok = true;
    for (int k = 0; k < 3; k++) {
      const float* p = &vert[(size_t)3 * tri_idx[3 * t + k]];
      if (p[2] == 0.0f) ok = false;
      px[k] = p[0]; py[k] = p[1]; pd[k] = p[2];
    }
    if (!ok) continue;
    float area = (px[1] - px[0]) * (py[2] - py[0]) - (px[2] - px[0]) * (py[1] - py[0]);
    if (!(area < 0.0f)) continue;
    float tx = px[1]; px[1] = px[2]; px[2] = tx;
    float ty = py[1]; py[1] = py[2]; py[2] = ty;
    float td = pd[1]; pd[1] = pd[2]; pd[2] = td;
    area = -area;
    float minx = fminf(px[0], fminf(px[1], px[2])), maxx = fmaxf(px[0], fmaxf(px[1], px[2]));
    float miny = fminf(py[0], fminf(py[1], py[2])), maxy = fmaxf(py[0], fmaxf(py[1], py[2]));
    int x0 = (int)floorf(minx - 0.5f), x1 = (int)ceilf(maxx - 0.5f), y0 = (int)floorf(miny - 0.5f), y1 = (int)ceilf(maxy - 0.5f);
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > W - 1) x1 = W - 1;
    if (y1 > H - 1) y1 = H - 1;
    if (x0 > x1 || y0 > y1) continue;
    float inv_area = 1.0f / area;
    for (int y = y0; y <= y1; y++)
      for (int x = x0; x <= x1; x++) {
        float sx = (float)x + 0.5f, sy = (float)y + 0.5f;
        float w0 = (px[2] - px[1]) * (sy - py[1]) - (py[2] - py[1]) * (sx - px[1]);
        float w1 = (px[0] - px[2]) * (sy - py[2]) - (py[0] - py[2]) * (sx - px[2]);
        float w2 = (px[1] - px[0]) * (sy - py[0]) - (py[1] - py[0]) * (sx - px[0]);
        if (w0 < 0.0f || w1 < 0.0f || w2 < 0.0f) continue;
        float inv = (w0 * pd[0] + w1 * pd[1] + w2 * pd[2]) * inv_area;
        if (!inv_depth_in_range(inv, dm)) continue;
        unsigned key = make_key(depth_key_inv(inv, dm), face, v);
        if (key < zkey[y * W + x]) zkey[y * W + x] = key;
      }
  }
  for (int i = 0; i < npix; i++) {
    int x = i % W, y = i / W;
    float ry = -((float)x + 0.5f - cx) * inv_fx, rz = -((float)y + 0.5f - cy) * inv_fy;
    F3 rdir = f3(1.0f, ry, rz);
    unsigned best = zkey[i];
    for (const Patch& P : patches) {
      if (x < P.x0 || x > P.x1 || y < P.y0 || y > P.y1) continue;
      const float* Rm = &vR[P.v * 9];
      const int a = P.face >> 1;
      const float inv = (Rm[a] * rdir.x + Rm[3 + a] * ry + Rm[6 + a] * rz) * P.cc;
      if (!inv_depth_in_range(inv, dm)) continue;
      if (P.bounded) {
        const int b = a == 2 ? 0 : a + 1, c = b == 2 ? 0 : b + 1;
        const float lb = vo[P.v * 3 + b] * inv + (Rm[b] * rdir.x + Rm[3 + b] * ry + Rm[6 + b] * rz);
        if (fabsf(lb) > vis_size[3 * P.v + b] * inv * 1.00001f) continue;
        const float lc = vo[P.v * 3 + c] * inv + (Rm[c] * rdir.x + Rm[3 + c] * ry + Rm[6 + c] * rz);
        if (fabsf(lc) > vis_size[3 * P.v + c] * inv * 1.00001f) continue;
      }
      unsigned key = make_key(depth_key_inv(inv, dm), P.face, P.v);
      if (key < best) best = key;
    }
    for (int v = 0; v < n_vis; v++) {
      if (vis_type[v] != 2) continue;
      const float* Rm = &vR[v * 9];
      F3 o = f3(vo[v * 3], vo[v * 3 + 1], vo[v * 3 + 2]);
      F3 dl = f3(Rm[0] * rdir.x + Rm[3] * rdir.y + Rm[6] * rdir.z, Rm[1] * rdir.x + Rm[4] * rdir.y + Rm[7] * rdir.z,
                 Rm[2] * rdir.x + Rm[5] * rdir.y + Rm[8] * rdir.z);
      float th = 0.0f;
      if (!ray_sphere(o, dl, vis_size[3 * v], th)) continue;
      float inv = 1.0f / th;
      if (inv_depth_in_range(inv, dm)) {
        unsigned key = make_key(depth_key_inv(inv, dm), 0, v);
        if (key < best) best = key;
      }
    }
    uint8_t* c4 = color + (size_t)i * 4;
    int16_t* p4 = posseg + (size_t)i * 4;
    c4[0] = c4[1] = c4[2] = 0; c4[3] = 255;
    p4[0] = p4[1] = p4[2] = p4[3] = 0;
    if (best != NO_HIT) {
      const int bv = (int)(best & 63u);
      const float depth = key_depth(best >> KEY_SHIFT, dm);
      F3 pc = rdir * depth;
      const int ty = vis_type[bv];
      unsigned rgbw;
      if (ty == 1 || ty == 0) {
        rgbw = face_rgb[bv * 6 + (int)((best >> 6) & 7u)];
      } else {
        F3 n_cam = f3(-1, 0, 0);
        if (ty == 4) {
          int xn = x + 1 < W ? x + 1 : x - 1, yn = y + 1 < H ? y + 1 : y - 1;
          unsigned kx = zkey[y * W + xn], ky = zkey[yn * W + x];
          if (kx != NO_HIT && ky != NO_HIT && (int)(kx & 63u) == bv && (int)(ky & 63u) == bv) {
            float dx_ = key_depth(kx >> KEY_SHIFT, dm), dy_ = key_depth(ky >> KEY_SHIFT, dm);
            F3 pxn = f3(1.0f, -((float)xn + 0.5f - cx) * inv_fx, rz) * dx_;
            F3 pyn = f3(1.0f, ry, -((float)yn + 0.5f - cy) * inv_fy) * dy_;
            F3 e1 = pxn - pc, e2 = pyn - pc;
            if (xn < x) e1 = -e1;
            if (yn < y) e2 = -e2;
            F3 nn = cross(e2, e1);
            float l = norm(nn);
            if (l > 1e-20f) n_cam = nn * (1.0f / l);
            if (n_cam.x > 0.0f) n_cam = -n_cam;
          }
        } else {
          n_cam = f3(pc.x - vt[bv * 3], pc.y - vt[bv * 3 + 1], pc.z - vt[bv * 3 + 2]) * (1.0f / vis_size[3 * bv]);
        }
        rgbw = shade_rgb(mulm(Rc, n_cam), vis_color + 4 * bv);
      }
      c4[0] = (uint8_t)(rgbw & 255u); c4[1] = (uint8_t)((rgbw >> 8) & 255u); c4[2] = (uint8_t)((rgbw >> 16) & 255u);
      p4[0] = to_mm(-pc.y); p4[1] = to_mm(pc.z); p4[2] = to_mm(-pc.x); p4[3] = (int16_t)vis_seg[bv];
    }
  }
}
}
