// b200sim batched tiled rasteriser -- replaces `camera_group.take_picture()` of the reference
// (mani_skill/utils/structs/render_camera.py:269-273; camera group creation mani_skill/envs/scene.py:1087-1106; pose
// sync physics->renderer mani_skill/envs/scene.py:404-427) for the "minimal" shader pack
// (mani_skill/render/shaders.py:68-84): per camera and sub-scene it produces
//     Color                 [H, W, 4] uint8  rgba
//     PositionSegmentation  [H, W, 4] int16  x, y, z in millimetres (OpenGL camera frame: x right, y up, z backward,
//                                            so depth = -z) and the segmentation id (= per_scene_id, 0 = background)
//
// One CTA renders one (sub-scene, camera) image.  The whole depth/id buffer of the image lives in shared memory
// (H*W 32-bit keys: 24-bit reversed-z depth | 8-bit visual id; 64 KB for 128x128):
//   pass 1  all threads stride over the triangles of the convex-hull visuals: body pose (read once from
//           rigid_body_data) x local pose -> camera frame, project, top-left fill rule, shared-memory atomicMin
//   pass 2  each thread owns pixels: analytic ray tests against boxes / spheres / the ground half-space, merge with the
//           rasterised key, shade (ambient 0.3 + two directional lights, mani_skill/envs/sapien_env.py:845-853), write
//           the two render targets straight to HBM with 4-byte (rgba8) and 8-byte (4 x int16) stores, coalesced by row.
// HBM traffic per image: the body rows once in, 12 B per pixel out.  The per-pixel arithmetic is restated in
// oracle/b2s_oracle_raster.cpp; this translation unit is compiled with -fmad=false so both produce identical masks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/b200sim.h"
#include "b2s_step.cuh"

namespace b2s {

struct RasterModel {
  int n_envs, n_cam, n_vis, n_tri_total, n_rows, n_ov;
  const int* vis_type;
  const int* vis_row;
  const float* vis_pose;   // [n_vis*7]
  const float* vis_size;   // [n_vis*3]
  const int* vis_hull;
  const float* vis_color;  // [n_vis*4]
  const int* vis_seg;
  const int* vis_ov;       // per-env override slot or -1
  const float* ov_size;    // SoA [n_ov*3][N]
  const float* ov_pose;    // SoA [n_ov*7][N]
  // triangle soup of the hull visuals
  const int* tri_vis;      // [n_tri_total] owning visual
  const float* tri_verts;  // [n_tri_total*9] local (hull frame) vertices
  // cameras
  const int* cam_w;
  const int* cam_h;
  const float* cam_intr;   // [n_cam*6] fx fy cx cy near far
  const int* cam_mount;    // row or -1
  const float* cam_pose;   // [n_cam*7]
  size_t* cam_offset;      // [n_cam] pixel offset of camera c inside one env's block
  size_t pixels_per_env;
};

struct RasterGroup {
  RasterModel R;
  std::vector<void*> allocs;
  uint8_t* color;
  int16_t* posseg;
  int max_pixels;  // largest camera image (pixels) -> shared memory size
};

#define B2S_DEPTH_BITS 24
#define B2S_DEPTH_MAX 16777215.0f

// Reversed-z quantisation shared by the raster and the analytic pass, on 1/depth.  The constants of a camera are computed once
// (DepthMap), a sample costs no division: the key comes from the interpolated 1/depth directly, the near/far test is done on it.
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

// ray (origin o, direction dvec, both in the box frame) against an axis-aligned box of half extents h: entry distance
B2S_HD bool ray_box(v3 o, v3 dv, v3 h, float& t_hit, v3& n_local) {
  float tmin = -1e30f, tmax = 1e30f;
  int axis = 0;
  float sgn = 1.0f;
  float oo[3] = {o.x, o.y, o.z}, dd[3] = {dv.x, dv.y, dv.z}, hh[3] = {h.x, h.y, h.z};
  for (int k = 0; k < 3; k++) {
    if (fabsf(dd[k]) < 1e-12f) {
      if (fabsf(oo[k]) > hh[k]) return false;
    } else {
      float inv = 1.0f / dd[k];
      float t0 = (-hh[k] - oo[k]) * inv, t1 = (hh[k] - oo[k]) * inv;
      float s = -1.0f;
      if (t0 > t1) { float tt = t0; t0 = t1; t1 = tt; s = 1.0f; }
      if (t0 > tmin) { tmin = t0; axis = k; sgn = s; }
      if (t1 < tmax) tmax = t1;
    }
  }
  if (tmin > tmax || tmax <= 0.0f || tmin <= 0.0f) return false;
  t_hit = tmin;
  n_local = mk3(axis == 0 ? sgn : 0.0f, axis == 1 ? sgn : 0.0f, axis == 2 ? sgn : 0.0f);
  return true;
}
B2S_HD bool ray_sphere(v3 o, v3 dv, float r, float& t_hit, v3& n_local) {
  float a = dot(dv, dv), b = dot(o, dv), c = dot(o, o) - r * r;
  float disc = b * b - a * c;
  if (disc < 0.0f) return false;
  float t = (-b - sqrtf(disc)) / a;
  if (t <= 0.0f) return false;
  t_hit = t;
  n_local = (o + dv * t) * (1.0f / r);
  return true;
}

B2S_HD uint8_t to_u8(float x) {
  float c = fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f + 0.5f;
  return (uint8_t)c;
}
B2S_HD v3 shade(v3 base, v3 n_world) {
  // ambient 0.3 + directional [1,1,-1] + directional [0,0,-1] (white), Lambert
  const float k = 0.57735026f;
  v3 l1 = mk3(-k, -k, k), l2 = mk3(0.0f, 0.0f, 1.0f);
  float w = 0.3f + 0.5f * fmaxf(dot(n_world, l1), 0.0f) + 0.5f * fmaxf(dot(n_world, l2), 0.0f);
  return base * w;
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

#define B2S_MAX_BIG_TRIS 512
#define B2S_BIG_TRI_PIXELS 24     // bounding boxes above this many pixels leave the one-thread path
#define B2S_HUGE_TRI_PIXELS 2048  // ... and above this many are shared by the whole CTA instead of one warp

// one coverage + depth sample: edge functions at the pixel centre, 1/depth interpolated in screen space, atomicMin on the key
__device__ __forceinline__ void raster_sample(unsigned* zkey, int W, int x, int y, const float* px, const float* py, const float* pd, float inv_area,
                                              int v, const DepthMap& dm) {
  float sx = (float)x + 0.5f, sy = (float)y + 0.5f;
  float w0 = (px[2] - px[1]) * (sy - py[1]) - (py[2] - py[1]) * (sx - px[1]);
  float w1 = (px[0] - px[2]) * (sy - py[2]) - (py[0] - py[2]) * (sx - px[2]);
  float w2 = (px[1] - px[0]) * (sy - py[0]) - (py[1] - py[0]) * (sx - px[0]);
  if (w0 < 0.0f || w1 < 0.0f || w2 < 0.0f) return;
  float inv = (w0 * pd[0] + w1 * pd[1] + w2 * pd[2]) * inv_area;
  if (!inv_depth_in_range(inv, dm)) return;
  unsigned key = (depth_key_inv(inv, dm) << 8) | (unsigned)v;
  atomicMin(&zkey[y * W + x], key);
}

__global__ void __launch_bounds__(512) raster_kernel(RasterModel R, const float* __restrict__ body_data, uint8_t* __restrict__ color,
                                                     int16_t* __restrict__ posseg) {
  extern __shared__ unsigned zkey[];
  __shared__ float vis_R[64][9];   // camera-from-visual rotation (row major)
  __shared__ float vis_t[64][3];   // camera-from-visual translation
  __shared__ float vis_Rw[64][9];  // world-from-visual rotation (lighting)
  __shared__ float vis_sz[64][3];
  __shared__ int vis_rect[64][4];  // conservative screen rectangle (x0, x1, y0, y1) of the ray-cast primitives
  __shared__ int vis_kind[64];
  __shared__ float vis_o[64][3];   // camera origin in the visual's frame (ray origin of the analytic tests)
  __shared__ float big_tri[B2S_MAX_BIG_TRIS][10];
  __shared__ int big_box[B2S_MAX_BIG_TRIS][7];  // x0 y0 width count visual, walk step (x, y) of the unit that rasterises it
  __shared__ int n_big;
  if (threadIdx.x == 0) n_big = 0;
  const int env = blockIdx.x / R.n_cam, cam = blockIdx.x % R.n_cam;
  const int W = R.cam_w[cam], H = R.cam_h[cam];
  const float fx = R.cam_intr[6 * cam], fy = R.cam_intr[6 * cam + 1], cx = R.cam_intr[6 * cam + 2], cy = R.cam_intr[6 * cam + 3];
  const float nearp = R.cam_intr[6 * cam + 4], farp = R.cam_intr[6 * cam + 5];
  const DepthMap dm = depth_map(nearp, farp);
  const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy;
  const int npix = W * H;
  for (int i = threadIdx.x; i < npix; i += blockDim.x) zkey[i] = 0xFFFFFFFFu;
  // camera pose in the sub-scene frame
  pose Xc = pmul(raster_body_pose(body_data, R.n_rows, env, R.cam_mount[cam]), pose7(R.cam_pose + 7 * cam));
  Xc.q = qnormalized(Xc.q);
  m3 Rc = qmat(Xc.q);
  const int nv = R.n_vis < 64 ? R.n_vis : 64;
  for (int v = threadIdx.x; v < nv; v += blockDim.x) {
    float lp[7];
    int ov = R.vis_ov[v];
    if (ov >= 0) {
      for (int k = 0; k < 7; k++) lp[k] = R.ov_pose[(size_t)(ov * 7 + k) * R.n_envs + env];
      for (int k = 0; k < 3; k++) vis_sz[v][k] = R.ov_size[(size_t)(ov * 3 + k) * R.n_envs + env];
    } else {
      for (int k = 0; k < 7; k++) lp[k] = R.vis_pose[7 * v + k];
      for (int k = 0; k < 3; k++) vis_sz[v][k] = R.vis_size[3 * v + k];
    }
    pose Xv = pmul(raster_body_pose(body_data, R.n_rows, env, R.vis_row[v]), pose7(lp));
    Xv.q = qnormalized(Xv.q);
    m3 Rv = qmat(Xv.q);
    m3 Rcv = mul(transpose(Rc), Rv);
    v3 tcv = tmul(Rc, Xv.p - Xc.p);
    for (int k = 0; k < 9; k++) { vis_R[v][k] = Rcv.m[k]; vis_Rw[v][k] = Rv.m[k]; }
    vis_t[v][0] = tcv.x; vis_t[v][1] = tcv.y; vis_t[v][2] = tcv.z;
    vis_o[v][0] = -(Rcv.m[0] * tcv.x + Rcv.m[3] * tcv.y + Rcv.m[6] * tcv.z);
    vis_o[v][1] = -(Rcv.m[1] * tcv.x + Rcv.m[4] * tcv.y + Rcv.m[7] * tcv.z);
    vis_o[v][2] = -(Rcv.m[2] * tcv.x + Rcv.m[5] * tcv.y + Rcv.m[8] * tcv.z);
    // screen rectangle that surely contains the primitive (whole image when it reaches behind the near plane)
    int rx0 = 0, rx1 = W - 1, ry0 = 0, ry1 = H - 1;
    int ty = R.vis_type[v];
    vis_kind[v] = ty;
    if (ty == SH_BOX || ty == SH_SPHERE) {
      float hx = vis_sz[v][0], hy = ty == SH_BOX ? vis_sz[v][1] : vis_sz[v][0], hz = ty == SH_BOX ? vis_sz[v][2] : vis_sz[v][0];
      float mnx = 1e30f, mxx = -1e30f, mny = 1e30f, mxy = -1e30f;
      bool behind = false;
      for (int c = 0; c < 8; c++) {
        v3 l = mk3((c & 1) ? hx : -hx, (c & 2) ? hy : -hy, (c & 4) ? hz : -hz);
        v3 pc = mul(Rcv, l) + tcv;
        if (pc.x <= nearp) { behind = true; break; }
        float u = cx - fx * pc.y / pc.x, w = cy - fy * pc.z / pc.x;
        mnx = fminf(mnx, u); mxx = fmaxf(mxx, u); mny = fminf(mny, w); mxy = fmaxf(mxy, w);
      }
      if (!behind) {
        rx0 = max(0, (int)floorf(mnx) - 1); rx1 = min(W - 1, (int)ceilf(mxx) + 1);
        ry0 = max(0, (int)floorf(mny) - 1); ry1 = min(H - 1, (int)ceilf(mxy) + 1);
      }
    }
    vis_rect[v][0] = rx0; vis_rect[v][1] = rx1; vis_rect[v][2] = ry0; vis_rect[v][3] = ry1;
  }
  __syncthreads();
  // ---------------- pass 1: rasterise hull triangles (camera frame: x forward, y left, z up)
  for (int t = threadIdx.x; t < R.n_tri_total; t += blockDim.x) {
    int v = R.tri_vis[t];
    if (v >= nv) continue;
    const float* tv = R.tri_verts + 9 * (size_t)t;
    float px[3], py[3], pd[3];
    bool ok = true;
    for (int k = 0; k < 3; k++) {
      v3 l = mk3(tv[3 * k], tv[3 * k + 1], tv[3 * k + 2]);
      float xc = vis_R[v][0] * l.x + vis_R[v][1] * l.y + vis_R[v][2] * l.z + vis_t[v][0];
      float yc = vis_R[v][3] * l.x + vis_R[v][4] * l.y + vis_R[v][5] * l.z + vis_t[v][1];
      float zc = vis_R[v][6] * l.x + vis_R[v][7] * l.y + vis_R[v][8] * l.z + vis_t[v][2];
      if (xc <= nearp) ok = false;
      float inv = 1.0f / xc;
      px[k] = cx - fx * yc * inv;
      py[k] = cy - fy * zc * inv;
      pd[k] = inv;
    }
    if (!ok) continue;  // triangles touching the near plane are dropped (robot links never get that close)
    float area = (px[1] - px[0]) * (py[2] - py[0]) - (px[2] - px[0]) * (py[1] - py[0]);
    // hull triangles are wound outwards (render.py build_visual_table): with screen x to the right and y down a front face has
    // negative area; back faces are culled (hulls are closed: they are hidden by the front faces), front faces made counter-clockwise
    if (!(area < 0.0f)) continue;
    float tx = px[1]; px[1] = px[2]; px[2] = tx;
    float ty = py[1]; py[1] = py[2]; py[2] = ty;
    float td = pd[1]; pd[1] = pd[2]; pd[2] = td;
    area = -area;
    float minx = fminf(px[0], fminf(px[1], px[2])), maxx = fmaxf(px[0], fmaxf(px[1], px[2]));
    float miny = fminf(py[0], fminf(py[1], py[2])), maxy = fmaxf(py[0], fmaxf(py[1], py[2]));
    int x0 = max(0, (int)floorf(minx - 0.5f)), x1 = min(W - 1, (int)ceilf(maxx - 0.5f));
    int y0 = max(0, (int)floorf(miny - 0.5f)), y1 = min(H - 1, (int)ceilf(maxy - 0.5f));
    if (x0 > x1 || y0 > y1) continue;
    float inv_area = 1.0f / area;
    const int bw = x1 - x0 + 1, cnt = bw * (y1 - y0 + 1);
    if (cnt > B2S_BIG_TRI_PIXELS) {
      // larger on-screen triangle: queued and rasterised by a whole warp (by the whole CTA when huge: close-ups of the wrist camera)
      // so that the one-thread path stays short and balanced
      int slot = atomicAdd(&n_big, 1);
      if (slot < B2S_MAX_BIG_TRIS) {
        float* o = big_tri[slot];
        o[0] = px[0]; o[1] = px[1]; o[2] = px[2]; o[3] = py[0]; o[4] = py[1]; o[5] = py[2];
        o[6] = pd[0]; o[7] = pd[1]; o[8] = pd[2]; o[9] = inv_area;
        const int stride = cnt > B2S_HUGE_TRI_PIXELS ? (int)blockDim.x : 32;  // pixels between two samples of one lane
        const int sy_ = stride / bw;
        big_box[slot][0] = x0; big_box[slot][1] = y0; big_box[slot][2] = bw; big_box[slot][3] = cnt; big_box[slot][4] = v;
        big_box[slot][5] = stride - sy_ * bw; big_box[slot][6] = sy_;
        continue;
      }
    }
    for (int y = y0; y <= y1; y++)
      for (int x = x0; x <= x1; x++) raster_sample(zkey, W, x, y, px, py, pd, inv_area, v, dm);
  }
  __syncthreads();
  {
    // pixel walk over a bounding box without a division per sample: a lane starts at pixel `first` and advances by `stride`
    // pixels, (xx, yy) follow incrementally with the precomputed (stride % width, stride / width)
    const int nb = n_big < B2S_MAX_BIG_TRIS ? n_big : B2S_MAX_BIG_TRIS;
    const int warp = threadIdx.x >> 5, n_warp = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int pass = 0; pass < 2; pass++) {  // 0: warp-sized triangles, one warp each; 1: huge ones, all threads
      for (int b = pass == 0 ? warp : 0; b < nb; b += pass == 0 ? n_warp : 1) {
        const int cnt = big_box[b][3];
        if ((cnt > B2S_HUGE_TRI_PIXELS) != (pass == 1)) continue;
        const float* o = big_tri[b];
        const int x0 = big_box[b][0], y0 = big_box[b][1], bw = big_box[b][2], v = big_box[b][4], sx_ = big_box[b][5], sy_ = big_box[b][6];
        const int first = pass == 0 ? lane : (int)threadIdx.x, stride = pass == 0 ? 32 : (int)blockDim.x;
        int yy = first / bw, xx = first - yy * bw;
        for (int p = first; p < cnt; p += stride) {
          raster_sample(zkey, W, x0 + xx, y0 + yy, o, o + 3, o + 6, o[9], v, dm);
          xx += sx_; yy += sy_;
          if (xx >= bw) { xx -= bw; yy++; }
        }
      }
    }
  }
  __syncthreads();
  // ---------------- pass 2: per pixel analytic primitives + merge + shade + store
  uint8_t* cbase = color + ((size_t)env * R.pixels_per_env + R.cam_offset[cam]) * 4;
  int16_t* pbase = posseg + ((size_t)env * R.pixels_per_env + R.cam_offset[cam]) * 4;
  // pixel walk: i = tid + k * blockDim; x, y advance incrementally (no division per pixel).  With W dividing blockDim the
  // column of a thread is fixed, so the set of ray-cast primitives whose screen rectangle covers that column is one
  // 64-bit mask computed once per column.
  int x = threadIdx.x % W, y = threadIdx.x / W;
  const int step_x = blockDim.x % W, step_y = blockDim.x / W;
  int mask_x = -1;
  unsigned long long xmask = 0ull;
  for (int i = threadIdx.x; i < npix; i += blockDim.x, x += step_x, y += step_y) {
    if (x >= W) { x -= W; y++; }
    if (x != mask_x) {
      xmask = 0ull;
      for (int v = 0; v < nv; v++)
        if (vis_kind[v] != SH_CONVEX && x >= vis_rect[v][0] && x <= vis_rect[v][1]) xmask |= 1ull << v;
      mask_x = x;
    }
    float ry = -((float)x + 0.5f - cx) * inv_fx, rz = -((float)y + 0.5f - cy) * inv_fy;
    v3 rdir = mk3(1.0f, ry, rz);  // camera frame, depth = distance along x
    float best = 1e30f;
    int best_v = -1;
    v3 best_n = mk3(0, 0, 0);  // world normal
    unsigned k = zkey[i];
    if (k != 0xFFFFFFFFu) {
      best = key_depth(k >> 8, dm);
      best_v = (int)(k & 255u);
    }
    bool raster_hit = best_v >= 0;
    for (unsigned long long m = xmask; m != 0ull; m &= m - 1ull) {
      const int v = __ffsll((long long)m) - 1;
      if (y < vis_rect[v][2] || y > vis_rect[v][3]) continue;
      const int ty = vis_kind[v];
      // ray in the visual's frame: o = Rcv^T (0 - t), d = Rcv^T rdir
      v3 o = mk3(vis_o[v][0], vis_o[v][1], vis_o[v][2]);
      v3 dl = mk3(vis_R[v][0] * rdir.x + vis_R[v][3] * rdir.y + vis_R[v][6] * rdir.z, vis_R[v][1] * rdir.x + vis_R[v][4] * rdir.y + vis_R[v][7] * rdir.z,
                  vis_R[v][2] * rdir.x + vis_R[v][5] * rdir.y + vis_R[v][8] * rdir.z);
      float th;
      v3 nl;
      bool hit = false;
      if (ty == SH_BOX) hit = ray_box(o, dl, mk3(vis_sz[v][0], vis_sz[v][1], vis_sz[v][2]), th, nl);
      else if (ty == SH_SPHERE) hit = ray_sphere(o, dl, vis_sz[v][0], th, nl);
      else if (ty == SH_PLANE) {  // half-space, normal = +x of the visual frame
        if (dl.x < -1e-9f && o.x > 0.0f) { th = -o.x / dl.x; nl = mk3(1, 0, 0); hit = true; }
      }
      if (hit && th > nearp && th < farp && th < best) {
        best = th;
        best_v = v;
        raster_hit = false;
        best_n = mk3(vis_Rw[v][0] * nl.x + vis_Rw[v][1] * nl.y + vis_Rw[v][2] * nl.z, vis_Rw[v][3] * nl.x + vis_Rw[v][4] * nl.y + vis_Rw[v][5] * nl.z,
                     vis_Rw[v][6] * nl.x + vis_Rw[v][7] * nl.y + vis_Rw[v][8] * nl.z);
      }
    }
    uchar4 c4 = make_uchar4(0, 0, 0, 255);
    short4 p4 = make_short4(0, 0, 0, 0);
    if (best_v >= 0) {
      v3 pc = rdir * best;  // camera-frame hit point
      if (raster_hit) {
        // screen-space normal from neighbouring depths of the same visual (flat-ish shading of hull faces)
        int xn = x + 1 < W ? x + 1 : x - 1, yn = y + 1 < H ? y + 1 : y - 1;
        unsigned kx = zkey[y * W + xn], ky = zkey[yn * W + x];
        v3 n_cam = mk3(-1, 0, 0);
        if (kx != 0xFFFFFFFFu && ky != 0xFFFFFFFFu && (int)(kx & 255u) == best_v && (int)(ky & 255u) == best_v) {
          float dx_ = key_depth(kx >> 8, dm), dy_ = key_depth(ky >> 8, dm);
          v3 pxn = mk3(1.0f, -((float)xn + 0.5f - cx) * inv_fx, rz) * dx_;
          v3 pyn = mk3(1.0f, ry, -((float)yn + 0.5f - cy) * inv_fy) * dy_;
          v3 e1 = pxn - pc, e2 = pyn - pc;
          if (xn < x) e1 = -e1;
          if (yn < y) e2 = -e2;
          v3 nn = cross(e2, e1);  // screen x runs to -y_cam, screen y to -z_cam: e2 x e1 faces the camera
          float l = norm(nn);
          if (l > 1e-20f) n_cam = nn * (1.0f / l);
          if (n_cam.x > 0.0f) n_cam = -n_cam;
        }
        best_n = mul(Rc, n_cam);
      }
      const float* col = R.vis_color + 4 * best_v;
      v3 rgb = shade(mk3(col[0], col[1], col[2]), best_n);
      c4 = make_uchar4(to_u8(rgb.x), to_u8(rgb.y), to_u8(rgb.z), 255);
      // OpenGL camera frame: x right = -y_cam, y up = z_cam, z backward = -x_cam
      p4 = make_short4(to_mm(-pc.y), to_mm(pc.z), to_mm(-pc.x), (short)R.vis_seg[best_v]);
    }
    *reinterpret_cast<uchar4*>(cbase + (size_t)i * 4) = c4;
    *reinterpret_cast<short4*>(pbase + (size_t)i * 4) = p4;
  }
}

#endif  // __CUDACC__ && B2S_RASTER_IMPL

const char* raster_create(const DevModel& M, const DevState& S, const B2SModel& host, const B2SCameraDesc* cams, int n_cam,
                          const B2SVisualTable* vis, RasterGroup** out, B2SRenderTargets* targets);
const char* raster_run(const DevModel& M, const DevState& S, RasterGroup* g, cudaStream_t st);
void raster_destroy(RasterGroup* g);

}  // namespace b2s
