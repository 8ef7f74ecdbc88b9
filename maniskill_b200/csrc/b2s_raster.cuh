// b200sim batched rasteriser -- replaces `camera_group.take_picture()` of the reference
// (mani_skill/utils/structs/render_camera.py:269-273; camera group creation mani_skill/envs/scene.py:1087-1106; pose
// sync physics->renderer mani_skill/envs/scene.py:404-427) for the "minimal" shader pack
// (mani_skill/render/shaders.py:68-84): per camera and sub-scene it produces
//     Color                 [H, W, 4] uint8  rgba
//     PositionSegmentation  [H, W, 4] int16  x, y, z in millimetres (OpenGL camera frame: x right, y up, z backward,
//                                            so depth = -z) and the segmentation id (= per_scene_id, 0 = background)
//
// A CTA renders one (sub-scene, camera) image at a time (two CTAs per SM take images from a ticket counter); the 32-bit depth/id
// buffer of the whole image lives in shared memory (23-bit reversed-z key of 1/depth | 3-bit box face | 6-bit visual index; 64 KB for
// 128 x 128).  Everything that can hide something goes through that key:
//   stage 0  per visual: camera-from-visual transform (body pose read once from rigid_body_data); the Lambert-shaded colour of the six
//            faces of a box / of a half-space (ambient 0.3 + the two directional lights of mani_skill/envs/sapien_env.py:845-853)
//   stage 1  the vertices of all rasterised visuals (hulls, boxes as 12 triangles) are projected ONCE into a shared-memory cache (screen
//            x, y, 1/depth: one division per vertex instead of three per triangle); flat faces that are tested per pixel are picked:
//            half-spaces, camera-facing faces of boxes that reach behind the near plane, camera-facing box faces with a large rectangle
//   stage 2  indexed triangles: back-face cull by the sign of the screen area, ">= 0" edge rule on counter-clockwise triangles, 1/depth
//            interpolated in screen space, shared-memory atomicMin on the key; small triangles by the thread that set them up,
//            larger ones queued (their setups kept in shared memory) and rasterised by a warp, huge ones by the CTA
//   stage 3  per pixel: the flat-face patches (1/depth of a ray-plane hit is linear in the ray: no division; inside test on the two
//            in-plane axes) and spheres produce keys of the same form, the smallest key wins, depth = 1 / (dequantised 1/depth): ONE
//            division per covered pixel; box / half-space pixels look their colour up by (visual, face), sphere pixels shade with
//            (hit - centre) / r, hull pixels with a normal from the depth neighbourhood; 32-bit stores assembled per warp.
// HBM traffic per image: the body rows in; out 12 B per pixel for the raw render targets, or only the textures the observation mode
// delivers (rgb 3 B + depth 2 B [+ segmentation 2 B] per pixel, written as whole 32-bit words assembled per warp).  The arithmetic is restated operation by operation in
// oracle/b2s_oracle_raster.cpp; this translation unit is compiled with -fmad=false so that both produce identical integers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/b200sim.h"
#include "b2s_step.cuh"

namespace b2s {

struct RasterModel {
  int n_envs, n_cam, n_vis, n_vert, n_tri, n_rows, n_ov;
  const int* vis_type;
  const int* vis_row;
  const float* vis_pose;   // [n_vis*7]
  const float* vis_size;   // [n_vis*3]
  const float* vis_color;  // [n_vis*4]
  const int* vis_seg;
  const int* vis_ov;       // per-env override slot or -1
  const float* ov_size;    // SoA [n_ov*3][N]
  const float* ov_pose;    // SoA [n_ov*7][N]
  // indexed geometry of the rasterised visuals (convex hulls: local vertices; boxes: the corners of the unit cube, scaled by the size)
  const float* vert_local; // [n_vert*3]
  const int* vert_vis;     // [n_vert]
  const int* tri_idx;      // [n_tri*3] vertex indices, wound outwards
  const int* tri_vis;      // [n_tri]
  // cameras
  const int* cam_w;
  const int* cam_h;
  const float* cam_intr;   // [n_cam*6] fx fy cx cy near far
  const int* cam_mount;    // row or -1
  const float* cam_pose;   // [n_cam*7]
  size_t* cam_offset;      // [n_cam] pixel offset of camera c inside one env's block
  size_t pixels_per_env;
};

// what a camera group writes per pixel (include/b200sim.h B2S_OUT_*): the shader pack's raw render targets and / or the compact
// textures the observation modes deliver (render/shaders.py:74-83: rgb = Color[..., :3], depth = -PositionSegmentation[..., 2],
// segmentation = PositionSegmentation[..., 3])
struct RasterTargets {
  unsigned mask;
  uint8_t* color;   // [N, pixels, 4] u8
  int16_t* posseg;  // [N, pixels, 4] i16
  uint8_t* rgb;     // [N, pixels, 3] u8
  int16_t* depth;   // [N, pixels]    i16 (mm)
  int16_t* seg;     // [N, pixels]    i16
};

struct RasterGroup {
  int* work;      // [2] image ticket counter + finished-CTA counter (device)
  int n_ctas;     // CTAs of a launch: two per SM (shared-memory bound), at most one per image
  RasterModel R;
  std::vector<void*> allocs;
  RasterTargets T;
  int max_pixels;  // largest camera image (pixels) -> shared memory size
};

// depth-buffer key: 23-bit reversed-z depth | 3-bit face of a box (2 * axis + (outward normal positive)) | 6-bit visual index
#define B2S_DEPTH_BITS 23
#define B2S_DEPTH_MAX 8388607.0f
#define B2S_KEY_SHIFT 9
#define B2S_NO_HIT 0xFFFFFFFFu
B2S_HD unsigned raster_key(unsigned depth_key, int face, int v) { return (depth_key << B2S_KEY_SHIFT) | ((unsigned)face << 6) | (unsigned)v; }
B2S_HD int key_visual(unsigned key) { return (int)(key & 63u); }
B2S_HD int key_face(unsigned key) { return (int)((key >> 6) & 7u); }

// Reversed-z quantisation on 1/depth.  The constants of a camera are computed once (DepthMap).
struct DepthMap {
  float invn, invf, range, scale, inv_max;  // 1/near, 1/far, invn - invf, 1 / range, 1 / B2S_DEPTH_MAX
};
B2S_HD DepthMap depth_map(float nearp, float farp) {
  DepthMap m;
  m.invn = 1.0f / nearp;
  m.invf = 1.0f / farp;
  m.range = m.invn - m.invf;
  m.scale = 1.0f / m.range;
  m.inv_max = 1.0f / B2S_DEPTH_MAX;
  return m;
}
B2S_HD bool inv_depth_in_range(float inv, const DepthMap& m) { return inv < m.invn && inv > m.invf; }
B2S_HD unsigned depth_key_inv(float inv, const DepthMap& m) {
  float t = (inv - m.invf) * m.scale;  // 1 at near, 0 at far
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  float q = B2S_DEPTH_MAX - t * B2S_DEPTH_MAX;
  return (unsigned)q;
}
B2S_HD float key_depth(unsigned k, const DepthMap& m) {
  float t = (B2S_DEPTH_MAX - (float)k) * m.inv_max;
  float inv = m.invf + t * m.range;
  return 1.0f / inv;
}

B2S_HD bool ray_sphere(v3 o, v3 dv, float r, float& t_hit) {
  float a = dot(dv, dv), b = dot(o, dv), c = dot(o, o) - r * r;
  float disc = b * b - a * c;
  if (disc < 0.0f) return false;
  float t = (-b - sqrtf(disc)) / a;
  if (t <= 0.0f) return false;
  t_hit = t;
  return true;
}
B2S_HD uint8_t to_u8(float x) {
  float c = fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f + 0.5f;
  return (uint8_t)c;
}
B2S_HD float shade_weight(v3 n_world) {
  // ambient 0.3 + directional [1,1,-1] + directional [0,0,-1] (white), Lambert
  const float k = 0.57735026f;
  v3 l1 = mk3(-k, -k, k), l2 = mk3(0.0f, 0.0f, 1.0f);
  return 0.3f + 0.5f * fmaxf(dot(n_world, l1), 0.0f) + 0.5f * fmaxf(dot(n_world, l2), 0.0f);
}
// flat colour of a surface with world normal n: rgb bytes packed r | g << 8 | b << 16
B2S_HD unsigned shade_rgb(v3 n_world, const float* col) {
  const float w = shade_weight(n_world);
  return (unsigned)to_u8(col[0] * w) | ((unsigned)to_u8(col[1] * w) << 8) | ((unsigned)to_u8(col[2] * w) << 16);
}
B2S_HD int16_t to_mm(float x) {
  float v = x * 1000.0f;
  v = fminf(fmaxf(v, -32768.0f), 32767.0f);
  return (int16_t)rintf(v);
}

#if defined(__CUDACC__) && defined(B2S_RASTER_IMPL)

__device__ __forceinline__ pose raster_body_pose(const float* body_data, int n_rows, int env, int row) {
  if (row < 0) return pose_ident();
  const float* o = body_data + ((size_t)env * n_rows + row) * 13;
  pose P;
  P.p = mk3(o[0], o[1], o[2]);
  P.q = mkq(o[3], o[4], o[5], o[6]);
  return P;
}

#define B2S_RASTER_THREADS 512
#define B2S_MAX_VIS 64
#define B2S_VERT_CACHE 1536       // projected vertices kept in shared memory (18 KB); vertices beyond are projected on use
#define B2S_MAX_BIG_TRIS 1024
#define B2S_MAX_HUGE_TRIS 128
#define B2S_BIG_CACHE 160         // screen-space setups of the first queued warp-sized triangles are kept (60 B each) ...
#define B2S_HUGE_CACHE 32         // ... and of the first CTA-sized ones; triangles queued beyond are set up again from the vertex cache
#define B2S_BIG_TRI_PIXELS 24     // bounding boxes above this many pixels leave the one-thread path
#define B2S_HUGE_TRI_PIXELS 2048  // ... and above this many are shared by the whole CTA instead of one warp
#define B2S_PATCH_PIXELS 1024     // faces of a projected box with a larger screen rectangle are tested per pixel instead of rasterised
#define B2S_MAX_PATCH 192         // flat faces tested per pixel (at most three faces of a box face the camera)
#define B2S_MAX_SPHERES 16

enum { VM_RASTER = 0, VM_ANALYTIC = 1 };

// screen-space setup of a triangle
struct TriSetup {
  float px[3], py[3], pd[3], inv_area;
  int x0, y0, x1, y1, vf;  // vf = the low 9 key bits: face << 6 | visual
};

// a flat face tested per pixel instead of being rasterised: the plane (local axis a = face >> 1) = s h_a of visual v, bounded by the
// other two half extents (boxes) or unbounded (half-spaces); 1/depth of the ray hit = (ray direction)_a * cc
struct FacePatch {
  short x0, x1, y0, y1;  // screen rectangle that contains it
  float cc;              // 1 / (s h_a - (camera origin)_a)
  short v, face;
  int bounded;
};

struct RasterShared {
  float vis_R[B2S_MAX_VIS][9];   // camera-from-visual rotation (row major)
  float vis_t[B2S_MAX_VIS][3];   // camera-from-visual translation
  float vis_col[B2S_MAX_VIS][3]; // base colour (per-pixel shading of spheres and hulls)
  unsigned face_rgb[B2S_MAX_VIS][6];  // shaded colour of the six faces of a box / of a half-space (face 1 = +x), packed r | g << 8 | b << 16
  short vis_seg[B2S_MAX_VIS];
  float vis_sz[B2S_MAX_VIS][3];
  float vis_o[B2S_MAX_VIS][3];   // camera origin in the visual's frame (ray origin of the analytic tests)
  int vis_rect[B2S_MAX_VIS][4];  // conservative screen rectangle (x0, x1, y0, y1) of the analytic primitives
  int vis_kind[B2S_MAX_VIS];
  int vis_mode[B2S_MAX_VIS];
  float vert[B2S_VERT_CACHE][3];  // screen x, screen y, 1/depth (0 = at or behind the near plane)
  unsigned short big[B2S_MAX_BIG_TRIS];    // triangles rasterised by one warp each
  unsigned short huge[B2S_MAX_HUGE_TRIS];  // triangles rasterised by the whole CTA
  TriSetup big_setup[B2S_BIG_CACHE];
  TriSetup huge_setup[B2S_HUGE_CACHE];
  FacePatch patch[B2S_MAX_PATCH];
  unsigned face_patch[B2S_MAX_VIS];  // bit f: face f of this box is a patch (its two triangles are skipped)
  unsigned char sphere[B2S_MAX_SPHERES];
  int n_patch, n_sphere;
  unsigned rgb_stage[B2S_RASTER_THREADS / 32][24];  // 32 pixels x 3 bytes of a warp, regrouped into 24 words
  int n_big, n_huge;
};

// local vertex i of the rasterised geometry -> screen x, y, 1/depth (out[2] = 0 when it is not in front of the near plane)
__device__ __forceinline__ void project_vertex(const RasterModel& R, const RasterShared& sh, int i, float fx, float fy, float cx, float cy, float nearp,
                                               float* out) {
  const int v = R.vert_vis[i];
  float lx = R.vert_local[3 * i], ly = R.vert_local[3 * i + 1], lz = R.vert_local[3 * i + 2];
  if (sh.vis_kind[v] == SH_BOX) { lx = lx * sh.vis_sz[v][0]; ly = ly * sh.vis_sz[v][1]; lz = lz * sh.vis_sz[v][2]; }
  const float xc = sh.vis_R[v][0] * lx + sh.vis_R[v][1] * ly + sh.vis_R[v][2] * lz + sh.vis_t[v][0];
  const float yc = sh.vis_R[v][3] * lx + sh.vis_R[v][4] * ly + sh.vis_R[v][5] * lz + sh.vis_t[v][1];
  const float zc = sh.vis_R[v][6] * lx + sh.vis_R[v][7] * ly + sh.vis_R[v][8] * lz + sh.vis_t[v][2];
  if (xc <= nearp) { out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; return; }
  const float inv = 1.0f / xc;
  out[0] = cx - fx * yc * inv;
  out[1] = cy - fy * zc * inv;
  out[2] = inv;
}

// screen-space setup of triangle t: false when it is culled (behind the near plane, back facing, off screen)
__device__ __forceinline__ bool setup_triangle(const RasterModel& R, const RasterShared& sh, int t, int W, int H, float fx, float fy, float cx, float cy,
                                               float nearp, TriSetup& T) {
  const int vf = R.tri_vis[t];  // visual | box face << 8 (packed by raster_create)
  const int v = vf & 255;
  if (v >= B2S_MAX_VIS || sh.vis_mode[v] != VM_RASTER) return false;
  if ((sh.face_patch[v] >> (vf >> 8)) & 1u) return false;  // this face is tested per pixel (stage 3)
  T.vf = ((vf >> 8) << 6) | v;
  for (int k = 0; k < 3; k++) {
    const int i = R.tri_idx[3 * t + k];
    float p[3];
    if (i < B2S_VERT_CACHE) { p[0] = sh.vert[i][0]; p[1] = sh.vert[i][1]; p[2] = sh.vert[i][2]; }
    else project_vertex(R, sh, i, fx, fy, cx, cy, nearp, p);
    if (p[2] == 0.0f) return false;  // triangles touching the near plane are dropped (robot links never get that close)
    T.px[k] = p[0]; T.py[k] = p[1]; T.pd[k] = p[2];
  }
  float area = (T.px[1] - T.px[0]) * (T.py[2] - T.py[0]) - (T.px[2] - T.px[0]) * (T.py[1] - T.py[0]);
  // triangles are wound outwards: with screen x to the right and y down a front face has negative area; back faces are culled
  // (the solids are closed: they are hidden by the front faces), front faces are made counter-clockwise
  if (!(area < 0.0f)) return false;
  float s;
  s = T.px[1]; T.px[1] = T.px[2]; T.px[2] = s;
  s = T.py[1]; T.py[1] = T.py[2]; T.py[2] = s;
  s = T.pd[1]; T.pd[1] = T.pd[2]; T.pd[2] = s;
  area = -area;
  const float minx = fminf(T.px[0], fminf(T.px[1], T.px[2])), maxx = fmaxf(T.px[0], fmaxf(T.px[1], T.px[2]));
  const float miny = fminf(T.py[0], fminf(T.py[1], T.py[2])), maxy = fmaxf(T.py[0], fmaxf(T.py[1], T.py[2]));
  T.x0 = max(0, (int)floorf(minx - 0.5f)); T.x1 = min(W - 1, (int)ceilf(maxx - 0.5f));
  T.y0 = max(0, (int)floorf(miny - 0.5f)); T.y1 = min(H - 1, (int)ceilf(maxy - 0.5f));
  if (T.x0 > T.x1 || T.y0 > T.y1) return false;
  T.inv_area = 1.0f / area;
  return true;
}

// one coverage + depth sample: edge functions at the pixel centre, 1/depth interpolated in screen space, atomicMin on the key
__device__ __forceinline__ void raster_sample(unsigned* zkey, int W, int x, int y, const TriSetup& T, const DepthMap& dm) {
  const float sx = (float)x + 0.5f, sy = (float)y + 0.5f;
  const float w0 = (T.px[2] - T.px[1]) * (sy - T.py[1]) - (T.py[2] - T.py[1]) * (sx - T.px[1]);
  const float w1 = (T.px[0] - T.px[2]) * (sy - T.py[2]) - (T.py[0] - T.py[2]) * (sx - T.px[2]);
  const float w2 = (T.px[1] - T.px[0]) * (sy - T.py[0]) - (T.py[1] - T.py[0]) * (sx - T.px[0]);
  if (w0 < 0.0f || w1 < 0.0f || w2 < 0.0f) return;
  const float inv = (w0 * T.pd[0] + w1 * T.pd[1] + w2 * T.pd[2]) * T.inv_area;
  if (!inv_depth_in_range(inv, dm)) return;
  const unsigned key = (depth_key_inv(inv, dm) << B2S_KEY_SHIFT) | (unsigned)T.vf;
  atomicMin(&zkey[y * W + x], key);
}

// env_mask (nullable): only sub-scenes with env_mask[env] != 0 are rendered (re-render after a partial reset); the others keep
// their previous picture.
// RAW: the shader pack's Color / PositionSegmentation targets are written (the hit position is needed); otherwise only the compact textures
template <bool RAW>
__device__ __forceinline__ void raster_image(const RasterModel& R, const float* __restrict__ body_data, const RasterTargets& O, int env, int cam,
                                             int big_tri_pixels, int patch_pixels, unsigned* zkey, RasterShared& sh) {
  if (threadIdx.x == 0) { sh.n_big = 0; sh.n_huge = 0; sh.n_patch = 0; sh.n_sphere = 0; }
  const int W = R.cam_w[cam], H = R.cam_h[cam];
  const float fx = R.cam_intr[6 * cam], fy = R.cam_intr[6 * cam + 1], cx = R.cam_intr[6 * cam + 2], cy = R.cam_intr[6 * cam + 3];
  const float nearp = R.cam_intr[6 * cam + 4], farp = R.cam_intr[6 * cam + 5];
  const DepthMap dm = depth_map(nearp, farp);
  const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy;
  const int npix = W * H;
  for (int i = threadIdx.x; i < npix; i += blockDim.x) zkey[i] = B2S_NO_HIT;
  // ---------------- stage 0: camera pose in the sub-scene frame, camera-from-visual transforms
  pose Xc = pmul(raster_body_pose(body_data, R.n_rows, env, R.cam_mount[cam]), pose7(R.cam_pose + 7 * cam));
  Xc.q = qnormalized(Xc.q);
  const m3 Rc = qmat(Xc.q);
  const int nv = R.n_vis < B2S_MAX_VIS ? R.n_vis : B2S_MAX_VIS;
  for (int v = threadIdx.x; v < nv; v += blockDim.x) {
    float lp[7];
    const int ov = R.vis_ov[v];
    if (ov >= 0) {
      for (int k = 0; k < 7; k++) lp[k] = R.ov_pose[(size_t)(ov * 7 + k) * R.n_envs + env];
      for (int k = 0; k < 3; k++) sh.vis_sz[v][k] = R.ov_size[(size_t)(ov * 3 + k) * R.n_envs + env];
    } else {
      for (int k = 0; k < 7; k++) lp[k] = R.vis_pose[7 * v + k];
      for (int k = 0; k < 3; k++) sh.vis_sz[v][k] = R.vis_size[3 * v + k];
    }
    pose Xv = pmul(raster_body_pose(body_data, R.n_rows, env, R.vis_row[v]), pose7(lp));
    Xv.q = qnormalized(Xv.q);
    const m3 Rv = qmat(Xv.q);
    const m3 Rcv = mul(transpose(Rc), Rv);
    const v3 tcv = tmul(Rc, Xv.p - Xc.p);
    for (int k = 0; k < 9; k++) sh.vis_R[v][k] = Rcv.m[k];
    {
      // Lambert shading is constant on a flat face: the six face colours of a box (a half-space: face 1, its +x) are computed once
      const float* col = R.vis_color + 4 * v;
      for (int k = 0; k < 3; k++) sh.vis_col[v][k] = col[k];
      sh.vis_seg[v] = (short)R.vis_seg[v];
      for (int f = 0; f < 6; f++) {
        const int a = f >> 1;
        const float sgn = (f & 1) ? 1.0f : -1.0f;
        sh.face_rgb[v][f] = shade_rgb(mk3(Rv.m[a] * sgn, Rv.m[3 + a] * sgn, Rv.m[6 + a] * sgn), col);
      }
    }
    sh.vis_t[v][0] = tcv.x; sh.vis_t[v][1] = tcv.y; sh.vis_t[v][2] = tcv.z;
    const float ox = -(Rcv.m[0] * tcv.x + Rcv.m[3] * tcv.y + Rcv.m[6] * tcv.z);
    sh.vis_o[v][0] = ox;
    sh.vis_o[v][1] = -(Rcv.m[1] * tcv.x + Rcv.m[4] * tcv.y + Rcv.m[7] * tcv.z);
    sh.vis_o[v][2] = -(Rcv.m[2] * tcv.x + Rcv.m[5] * tcv.y + Rcv.m[8] * tcv.z);
    const int ty = R.vis_type[v];
    sh.vis_kind[v] = ty;
    sh.face_patch[v] = 0u;
    // screen rectangle that surely contains the primitive (whole image when it reaches behind the near plane)
    int rx0 = 0, rx1 = W - 1, ry0 = 0, ry1 = H - 1;
    int mode = ty == SH_CONVEX ? VM_RASTER : VM_ANALYTIC;
    if (ty == SH_BOX || ty == SH_SPHERE) {
      const float hx = sh.vis_sz[v][0], hy = ty == SH_BOX ? sh.vis_sz[v][1] : sh.vis_sz[v][0], hz = ty == SH_BOX ? sh.vis_sz[v][2] : sh.vis_sz[v][0];
      float mnx = 1e30f, mxx = -1e30f, mny = 1e30f, mxy = -1e30f;
      bool behind = false;
      for (int c = 0; c < 8; c++) {
        const v3 l = mk3((c & 1) ? hx : -hx, (c & 2) ? hy : -hy, (c & 4) ? hz : -hz);
        const v3 pc = mul(Rcv, l) + tcv;
        if (pc.x <= nearp) { behind = true; break; }
        const float u = cx - fx * pc.y / pc.x, w = cy - fy * pc.z / pc.x;
        mnx = fminf(mnx, u); mxx = fmaxf(mxx, u); mny = fminf(mny, w); mxy = fmaxf(mxy, w);
      }
      if (!behind) {
        rx0 = max(0, (int)floorf(mnx) - 1); rx1 = min(W - 1, (int)ceilf(mxx) + 1);
        ry0 = max(0, (int)floorf(mny) - 1); ry1 = min(H - 1, (int)ceilf(mxy) + 1);
        if (ty == SH_BOX) mode = VM_RASTER;  // all eight corners in front of the near plane: its twelve triangles are rasterised
      }
    }
    sh.vis_mode[v] = mode;
    sh.vis_rect[v][0] = rx0; sh.vis_rect[v][1] = rx1; sh.vis_rect[v][2] = ry0; sh.vis_rect[v][3] = ry1;
    if (ty == SH_SPHERE) {
      const int slot = atomicAdd(&sh.n_sphere, 1);
      if (slot < B2S_MAX_SPHERES) sh.sphere[slot] = (unsigned char)v;
    }
  }
  __syncthreads();
  // ---------------- stage 1: project the vertices once; pick the flat faces that are tested per pixel
  {
    const int nc = R.n_vert < B2S_VERT_CACHE ? R.n_vert : B2S_VERT_CACHE;
    for (int i = threadIdx.x; i < nc; i += blockDim.x) project_vertex(R, sh, i, fx, fy, cx, cy, nearp, sh.vert[i]);
    // flat faces: half-spaces; the camera-facing faces of a box that reaches behind the near plane (it cannot be projected); the
    // camera-facing faces of the other boxes whose screen rectangle is large (two bounding-box walks of two huge triangles cost
    // more than one plane test per pixel, and the per-pixel pass is perfectly balanced)
    for (int vf = threadIdx.x; vf < nv * 6; vf += blockDim.x) {
      const int v = vf / 6, f = vf - 6 * v, a = f >> 1, ty = sh.vis_kind[v];
      const float sgn = (f & 1) ? 1.0f : -1.0f;
      float ha;
      if (ty == SH_PLANE) { if (f != 1) continue; ha = 0.0f; }
      else if (ty == SH_BOX) ha = sh.vis_sz[v][a];
      else continue;
      const float oa = sh.vis_o[v][a];
      if (!(oa * sgn > ha)) continue;  // the camera is behind this face
      int rx0 = 0, rx1 = W - 1, ry0 = 0, ry1 = H - 1;
      if (ty == SH_BOX && sh.vis_mode[v] == VM_RASTER) {
        const int b = a == 2 ? 0 : a + 1, c = b == 2 ? 0 : b + 1;
        float mnx = 1e30f, mxx = -1e30f, mny = 1e30f, mxy = -1e30f;
        for (int k = 0; k < 4; k++) {
          float l[3];
          l[a] = sgn * ha; l[b] = (k & 1) ? sh.vis_sz[v][b] : -sh.vis_sz[v][b]; l[c] = (k & 2) ? sh.vis_sz[v][c] : -sh.vis_sz[v][c];
          const float* Rm = sh.vis_R[v];
          const float px_ = Rm[0] * l[0] + Rm[1] * l[1] + Rm[2] * l[2] + sh.vis_t[v][0];
          const float py_ = Rm[3] * l[0] + Rm[4] * l[1] + Rm[5] * l[2] + sh.vis_t[v][1];
          const float pz_ = Rm[6] * l[0] + Rm[7] * l[1] + Rm[8] * l[2] + sh.vis_t[v][2];
          const float u = cx - fx * py_ / px_, w = cy - fy * pz_ / px_;
          mnx = fminf(mnx, u); mxx = fmaxf(mxx, u); mny = fminf(mny, w); mxy = fmaxf(mxy, w);
        }
        rx0 = max(0, (int)floorf(mnx) - 1); rx1 = min(W - 1, (int)ceilf(mxx) + 1);
        ry0 = max(0, (int)floorf(mny) - 1); ry1 = min(H - 1, (int)ceilf(mxy) + 1);
        if (rx0 > rx1 || ry0 > ry1 || (rx1 - rx0 + 1) * (ry1 - ry0 + 1) <= patch_pixels) continue;  // stays a pair of triangles
      } else if (ty == SH_BOX) {
        // the box reaches behind the near plane: its face is clipped against that plane first (a quadrilateral against one plane: at most
        // five corners) and the screen rectangle of what is left prunes the per-pixel test.  Pruning only: the rectangle is widened by two
        // pixels and the hit test decides, so the CPU restatement may (and does) test every pixel for these faces.
        const int b = a == 2 ? 0 : a + 1, c = b == 2 ? 0 : b + 1;
        const float* Rm = sh.vis_R[v];
        float qx[4], qy[4], qz[4];
        for (int k = 0; k < 4; k++) {
          float l[3];
          const int kb = (k == 1 || k == 2), kc = (k >= 2);  // corners in order around the face
          l[a] = sgn * ha; l[b] = kb ? sh.vis_sz[v][b] : -sh.vis_sz[v][b]; l[c] = kc ? sh.vis_sz[v][c] : -sh.vis_sz[v][c];
          qx[k] = Rm[0] * l[0] + Rm[1] * l[1] + Rm[2] * l[2] + sh.vis_t[v][0];
          qy[k] = Rm[3] * l[0] + Rm[4] * l[1] + Rm[5] * l[2] + sh.vis_t[v][1];
          qz[k] = Rm[6] * l[0] + Rm[7] * l[1] + Rm[8] * l[2] + sh.vis_t[v][2];
        }
        float mnx = 1e30f, mxx = -1e30f, mny = 1e30f, mxy = -1e30f;
        int kept = 0;
        for (int k = 0; k < 4; k++) {
          const int k1 = (k + 1) & 3;
          const bool in0 = qx[k] > nearp, in1 = qx[k1] > nearp;
          if (in0) {
            const float u = cx - fx * qy[k] / qx[k], w = cy - fy * qz[k] / qx[k];
            mnx = fminf(mnx, u); mxx = fmaxf(mxx, u); mny = fminf(mny, w); mxy = fmaxf(mxy, w); kept++;
          }
          if (in0 != in1) {  // the edge crosses the near plane
            const float tt = (nearp - qx[k]) / (qx[k1] - qx[k]);
            const float ey = qy[k] + tt * (qy[k1] - qy[k]), ez = qz[k] + tt * (qz[k1] - qz[k]);
            const float u = cx - fx * ey / nearp, w = cy - fy * ez / nearp;
            mnx = fminf(mnx, u); mxx = fmaxf(mxx, u); mny = fminf(mny, w); mxy = fmaxf(mxy, w); kept++;
          }
        }
        if (kept == 0) continue;  // the whole face is behind the near plane
        // points on the near plane project far outside the image: clamp before the float -> int conversion
        mnx = fminf(fmaxf(mnx, -4.0f), (float)W + 4.0f); mxx = fminf(fmaxf(mxx, -4.0f), (float)W + 4.0f);
        mny = fminf(fmaxf(mny, -4.0f), (float)H + 4.0f); mxy = fminf(fmaxf(mxy, -4.0f), (float)H + 4.0f);
        rx0 = max(0, (int)floorf(mnx) - 2); rx1 = min(W - 1, (int)ceilf(mxx) + 2);
        ry0 = max(0, (int)floorf(mny) - 2); ry1 = min(H - 1, (int)ceilf(mxy) + 2);
        if (rx0 > rx1 || ry0 > ry1) continue;
      }
      const int slot = atomicAdd(&sh.n_patch, 1);
      if (slot >= B2S_MAX_PATCH) continue;
      if (ty == SH_BOX) atomicOr(&sh.face_patch[v], 1u << f);
      FacePatch P;
      P.x0 = (short)rx0; P.x1 = (short)rx1; P.y0 = (short)ry0; P.y1 = (short)ry1;
      P.cc = 1.0f / (sgn * ha - oa);
      P.v = (short)v; P.face = (short)f; P.bounded = ty == SH_BOX;
      sh.patch[slot] = P;
    }
  }
  __syncthreads();
  // ---------------- stage 2: triangles (camera frame: x forward, y left, z up)
  for (int t = threadIdx.x; t < R.n_tri; t += blockDim.x) {
    TriSetup T;
    if (!setup_triangle(R, sh, t, W, H, fx, fy, cx, cy, nearp, T)) continue;
    const int cnt = (T.x1 - T.x0 + 1) * (T.y1 - T.y0 + 1);
    if (cnt > big_tri_pixels && t < 65536) {
      // larger on-screen triangle: queued and rasterised by a whole warp (by the whole CTA when huge) so that the one-thread path
      // stays short and balanced
      if (cnt > B2S_HUGE_TRI_PIXELS) {
        const int slot = atomicAdd(&sh.n_huge, 1);
        if (slot < B2S_MAX_HUGE_TRIS) {
          sh.huge[slot] = (unsigned short)t;
          if (slot < B2S_HUGE_CACHE) sh.huge_setup[slot] = T;
          continue;
        }
      }
      const int slot = atomicAdd(&sh.n_big, 1);
      if (slot < B2S_MAX_BIG_TRIS) {
        sh.big[slot] = (unsigned short)t;
        if (slot < B2S_BIG_CACHE) sh.big_setup[slot] = T;
        continue;
      }
    }
    for (int y = T.y0; y <= T.y1; y++)
      for (int x = T.x0; x <= T.x1; x++) raster_sample(zkey, W, x, y, T, dm);
  }
  __syncthreads();
  {
    // pixel walk over a bounding box without a division per sample: a lane starts at pixel `first` and advances by `stride` pixels
    const int warp = threadIdx.x >> 5, n_warp = blockDim.x >> 5, lane = threadIdx.x & 31;
    // warp-sized triangles, one warp each; the setup of the first pass is read back from shared memory (broadcast loads)
    const int nb = sh.n_big < B2S_MAX_BIG_TRIS ? sh.n_big : B2S_MAX_BIG_TRIS;
    for (int b = warp; b < nb; b += n_warp) {
      TriSetup T;
      if (b < B2S_BIG_CACHE) T = sh.big_setup[b];
      else if (!setup_triangle(R, sh, (int)sh.big[b], W, H, fx, fy, cx, cy, nearp, T)) continue;
      const int bw = T.x1 - T.x0 + 1, cnt = bw * (T.y1 - T.y0 + 1);
      const int sy_ = 32 / bw, sx_ = 32 - sy_ * bw;
      int yy = lane / bw, xx = lane - yy * bw;
      for (int p = lane; p < cnt; p += 32) {
        raster_sample(zkey, W, T.x0 + xx, T.y0 + yy, T, dm);
        xx += sx_; yy += sy_;
        if (xx >= bw) { xx -= bw; yy++; }
      }
    }
    // huge triangles (close-ups, table faces): the whole CTA walks the bounding box of each
    const int nh = sh.n_huge < B2S_MAX_HUGE_TRIS ? sh.n_huge : B2S_MAX_HUGE_TRIS;
    for (int b = 0; b < nh; b++) {
      TriSetup T;
      if (b < B2S_HUGE_CACHE) T = sh.huge_setup[b];
      else if (!setup_triangle(R, sh, (int)sh.huge[b], W, H, fx, fy, cx, cy, nearp, T)) continue;
      const int bw = T.x1 - T.x0 + 1, cnt = bw * (T.y1 - T.y0 + 1);
      const int first = (int)threadIdx.x, stride = (int)blockDim.x;
      const int sy_ = stride / bw, sx_ = stride - sy_ * bw;
      int yy = first / bw, xx = first - yy * bw;
      for (int p = first; p < cnt; p += stride) {
        raster_sample(zkey, W, T.x0 + xx, T.y0 + yy, T, dm);
        xx += sx_; yy += sy_;
        if (xx >= bw) { xx -= bw; yy++; }
      }
    }
  }
  __syncthreads();
  // ---------------- stage 3: per pixel analytic primitives + merge + shade + store
  const size_t pix0 = (size_t)env * R.pixels_per_env + R.cam_offset[cam];  // first pixel of this image in the targets
  uint8_t* cbase = O.color + pix0 * 4;
  int16_t* pbase = O.posseg + pix0 * 4;
  // the compact textures are written as 32-bit words assembled per warp (32 consecutive pixels = 96 B of rgb, 64 B of depth): needs
  // whole warps on consecutive pixels, i.e. a pixel count that is a multiple of the CTA size; otherwise element-wise stores
  const bool packed = (npix % (int)blockDim.x) == 0 && (pix0 % 32) == 0;
  const int warp_id = threadIdx.x >> 5, lane_id = threadIdx.x & 31;
  // pixel walk: i = tid + k * blockDim; x, y advance incrementally (no division per pixel)
  int x = threadIdx.x % W, y = threadIdx.x / W;
  const int step_x = blockDim.x % W, step_y = blockDim.x / W;
  const int n_patch = sh.n_patch < B2S_MAX_PATCH ? sh.n_patch : B2S_MAX_PATCH, n_sphere = sh.n_sphere < B2S_MAX_SPHERES ? sh.n_sphere : B2S_MAX_SPHERES;
  // running output pointers of the packed stores: per iteration a warp advances by blockDim pixels
  unsigned* rgb_w = reinterpret_cast<unsigned*>(O.rgb + pix0 * 3) + (size_t)warp_id * 24 + lane_id;
  unsigned* depth_w = reinterpret_cast<unsigned*>(O.depth + pix0) + (threadIdx.x >> 1);
  unsigned* seg_w = reinterpret_cast<unsigned*>(O.seg + pix0) + (threadIdx.x >> 1);
  const int rgb_step = (int)(blockDim.x >> 5) * 24, half_step = (int)(blockDim.x >> 1);
  for (int i = threadIdx.x; i < npix; i += blockDim.x, x += step_x, y += step_y, rgb_w += rgb_step, depth_w += half_step, seg_w += half_step) {
    if (x >= W) { x -= W; y++; }
    const float ry = -((float)x + 0.5f - cx) * inv_fx, rz = -((float)y + 0.5f - cy) * inv_fy;
    const v3 rdir = mk3(1.0f, ry, rz);  // camera frame, depth = distance along x
    unsigned best = zkey[i];
    for (int k = 0; k < n_patch; k++) {
      const FacePatch P = sh.patch[k];
      if (x < P.x0 || x > P.x1 || y < P.y0 || y > P.y1) continue;
      const float* Rm = sh.vis_R[P.v];
      const int a = P.face >> 1;
      // ray direction in the visual's frame is Rcv^T rdir; 1/depth of the hit with the face plane is linear in it: no division
      const float inv = (Rm[a] * rdir.x + Rm[3 + a] * ry + Rm[6 + a] * rz) * P.cc;
      if (!inv_depth_in_range(inv, dm)) continue;
      if (P.bounded) {
        // hit point (o + dl / inv) inside the face: |o_b inv + dl_b| <= h_b inv for the other two axes (inv > 0)
        const int b = a == 2 ? 0 : a + 1, c = b == 2 ? 0 : b + 1;
        const float lb = sh.vis_o[P.v][b] * inv + (Rm[b] * rdir.x + Rm[3 + b] * ry + Rm[6 + b] * rz);
        if (fabsf(lb) > sh.vis_sz[P.v][b] * inv * 1.00001f) continue;
        const float lc = sh.vis_o[P.v][c] * inv + (Rm[c] * rdir.x + Rm[3 + c] * ry + Rm[6 + c] * rz);
        if (fabsf(lc) > sh.vis_sz[P.v][c] * inv * 1.00001f) continue;
      }
      const unsigned key = raster_key(depth_key_inv(inv, dm), P.face, P.v);
      best = key < best ? key : best;
    }
    for (int k = 0; k < n_sphere; k++) {
      const int v = sh.sphere[k];
      if (x < sh.vis_rect[v][0] || x > sh.vis_rect[v][1] || y < sh.vis_rect[v][2] || y > sh.vis_rect[v][3]) continue;
      // ray in the visual's frame: o = Rcv^T (0 - t), d = Rcv^T rdir
      const v3 o = mk3(sh.vis_o[v][0], sh.vis_o[v][1], sh.vis_o[v][2]);
      const v3 dl = mk3(sh.vis_R[v][0] * rdir.x + sh.vis_R[v][3] * rdir.y + sh.vis_R[v][6] * rdir.z,
                        sh.vis_R[v][1] * rdir.x + sh.vis_R[v][4] * rdir.y + sh.vis_R[v][7] * rdir.z,
                        sh.vis_R[v][2] * rdir.x + sh.vis_R[v][5] * rdir.y + sh.vis_R[v][8] * rdir.z);
      float th = 0.0f;
      if (!ray_sphere(o, dl, sh.vis_sz[v][0], th)) continue;
      const float inv = 1.0f / th;
      if (inv_depth_in_range(inv, dm)) {
        const unsigned key = raster_key(depth_key_inv(inv, dm), 0, v);
        best = key < best ? key : best;
      }
    }
    uchar4 c4 = make_uchar4(0, 0, 0, 255);
    short4 p4 = make_short4(0, 0, 0, 0);
    if (best != B2S_NO_HIT) {
      const int bv = key_visual(best);
      const float depth = key_depth(best >> B2S_KEY_SHIFT, dm);
      const int ty = sh.vis_kind[bv];
      unsigned rgbw;
      if (ty == SH_BOX || ty == SH_PLANE) {
        rgbw = sh.face_rgb[bv][key_face(best)];  // flat face: shaded once per image in stage 0
      } else {
        const v3 pc = rdir * depth;  // camera-frame hit point
        v3 n_cam = mk3(-1, 0, 0);
        if (ty == SH_CONVEX) {
          // screen-space normal from neighbouring depths of the same visual (flat-ish shading of hull faces)
          const int xn = x + 1 < W ? x + 1 : x - 1, yn = y + 1 < H ? y + 1 : y - 1;
          const unsigned kx = zkey[y * W + xn], ky = zkey[yn * W + x];
          if (kx != B2S_NO_HIT && ky != B2S_NO_HIT && key_visual(kx) == bv && key_visual(ky) == bv) {
            const float dx_ = key_depth(kx >> B2S_KEY_SHIFT, dm), dy_ = key_depth(ky >> B2S_KEY_SHIFT, dm);
            const v3 pxn = mk3(1.0f, -((float)xn + 0.5f - cx) * inv_fx, rz) * dx_;
            const v3 pyn = mk3(1.0f, ry, -((float)yn + 0.5f - cy) * inv_fy) * dy_;
            v3 e1 = pxn - pc, e2 = pyn - pc;
            if (xn < x) e1 = -e1;
            if (yn < y) e2 = -e2;
            const v3 nn = cross(e2, e1);  // screen x runs to -y_cam, screen y to -z_cam: e2 x e1 faces the camera
            const float l = norm(nn);
            if (l > 1e-20f) n_cam = nn * (1.0f / l);
            if (n_cam.x > 0.0f) n_cam = -n_cam;
          }
        } else {
          // sphere: (hit point - centre) / radius, in the camera frame
          n_cam = mk3(pc.x - sh.vis_t[bv][0], pc.y - sh.vis_t[bv][1], pc.z - sh.vis_t[bv][2]) * (1.0f / sh.vis_sz[bv][0]);
        }
        rgbw = shade_rgb(mul(Rc, n_cam), sh.vis_col[bv]);
      }
      c4 = make_uchar4((uint8_t)(rgbw & 255u), (uint8_t)((rgbw >> 8) & 255u), (uint8_t)((rgbw >> 16) & 255u), 255);
      // OpenGL camera frame: x right = -y_cam, y up = z_cam, z backward = -x_cam (the hit point is ray * depth, the ray's x is 1)
      p4.z = to_mm(-depth);
      p4.w = sh.vis_seg[bv];
      if (RAW) { p4.x = to_mm(-(ry * depth)); p4.y = to_mm(rz * depth); }
    }
    if (O.mask & B2S_OUT_COLOR) *reinterpret_cast<uchar4*>(cbase + (size_t)i * 4) = c4;
    if (O.mask & B2S_OUT_POSSEG) *reinterpret_cast<short4*>(pbase + (size_t)i * 4) = p4;
    if (O.mask & (B2S_OUT_RGB | B2S_OUT_DEPTH | B2S_OUT_SEG)) {
      const short dmm = (short)(-(int)p4.z);  // depth = -z of the OpenGL frame, with int16 negation semantics
      if (packed) {
        if (O.mask & B2S_OUT_RGB) {
          uint8_t* wb = reinterpret_cast<uint8_t*>(sh.rgb_stage[warp_id]);
          wb[3 * lane_id] = c4.x; wb[3 * lane_id + 1] = c4.y; wb[3 * lane_id + 2] = c4.z;
          __syncwarp();
          if (lane_id < 24) *rgb_w = sh.rgb_stage[warp_id][lane_id];
          __syncwarp();
        }
        if (O.mask & B2S_OUT_DEPTH) {
          const unsigned lo = (unsigned)(unsigned short)dmm, hi = __shfl_down_sync(0xffffffffu, lo, 1);
          if (!(lane_id & 1)) *depth_w = lo | (hi << 16);
        }
        if (O.mask & B2S_OUT_SEG) {
          const unsigned lo = (unsigned)(unsigned short)p4.w, hi = __shfl_down_sync(0xffffffffu, lo, 1);
          if (!(lane_id & 1)) *seg_w = lo | (hi << 16);
        }
      } else {
        if (O.mask & B2S_OUT_RGB) { uint8_t* o3 = O.rgb + (pix0 + i) * 3; o3[0] = c4.x; o3[1] = c4.y; o3[2] = c4.z; }
        if (O.mask & B2S_OUT_DEPTH) O.depth[pix0 + i] = dmm;
        if (O.mask & B2S_OUT_SEG) O.seg[pix0 + i] = p4.w;
      }
    }
  }
}


// The images of a launch are handed out through a counter in global memory: `work[0]` = next ticket, `work[1]` = CTAs that have finished
// (the last one rewinds both for the next launch).  A CTA keeps taking tickets until none is left; images of sub-scenes that are masked out
// are skipped by the one thread that takes the ticket, so a masked re-render after a step on which nothing finished costs 128 atomics
// instead of thousands of CTA launches with 110 KB of shared memory each.
template <bool RAW>
__global__ void __launch_bounds__(B2S_RASTER_THREADS, 2) raster_kernel(RasterModel R, const float* __restrict__ body_data, RasterTargets O,
                                                                    const uint8_t* __restrict__ env_mask, int big_tri_pixels, int patch_pixels, int* work) {
  extern __shared__ unsigned zkey[];
  __shared__ RasterShared sh;
  __shared__ int s_img, s_cur, s_end;
  const int n_img = R.n_envs * R.n_cam;
  // a ticket is one image of a full render (best balance) and 32 consecutive images of a masked one (the taker scans their flags)
  const int chunk = env_mask ? 32 : 1;
  if (threadIdx.x == 0) { s_cur = 0; s_end = 0; }
  for (;;) {
    __syncthreads();  // the previous image's shared state is no longer read
    if (threadIdx.x == 0) {
      int img = n_img;
      for (;;) {
        if (s_cur >= s_end) {
          s_cur = atomicAdd(&work[0], chunk);
          s_end = s_cur + chunk < n_img ? s_cur + chunk : n_img;
          if (s_cur >= n_img) break;
        }
        const int c = s_cur++;
        if (!env_mask || env_mask[c / R.n_cam]) { img = c; break; }
      }
      s_img = img;
    }
    __syncthreads();
    const int img = s_img;
    if (img >= n_img) break;
    raster_image<RAW>(R, body_data, O, img / R.n_cam, img % R.n_cam, big_tri_pixels, patch_pixels, zkey, sh);
  }
  if (threadIdx.x == 0 && atomicAdd(&work[1], 1) == (int)gridDim.x - 1) { work[0] = 0; work[1] = 0; __threadfence(); }
}

#endif  // __CUDACC__ && B2S_RASTER_IMPL

const char* raster_create(const DevModel& M, const DevState& S, const B2SModel& host, const B2SCameraDesc* cams, int n_cam,
                          const B2SVisualTable* vis, unsigned outputs, RasterGroup** out, B2SRenderTargets* targets);
const char* raster_run(const DevModel& M, const DevState& S, RasterGroup* g, const uint8_t* env_mask, cudaStream_t st);
void raster_destroy(RasterGroup* g);

}  // namespace b2s
