// Host side of the rasteriser + the kernel instantiation.  Compiled with -fmad=false (see Makefile) so the per-pixel
// arithmetic is plain IEEE mul/add/div/sqrt and the segmentation mask is reproducible bit for bit by the CPU oracle.
#define B2S_RASTER_IMPL
#include "b2s_raster.cuh"

#include <stdlib.h>
#include <string.h>

namespace b2s {

namespace {
template <class T>
T* dev_copy(RasterGroup* g, const T* src, size_t n) {
  if (n == 0) n = 1, src = nullptr;
  void* d = nullptr;
  if (cudaMalloc(&d, n * sizeof(T)) != cudaSuccess) return nullptr;
  if (src) cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyHostToDevice);
  else cudaMemset(d, 0, n * sizeof(T));
  g->allocs.push_back(d);
  return (T*)d;
}
}  // namespace

const char* raster_create(const DevModel& M, const DevState& S, const B2SModel& host, const B2SCameraDesc* cams, int n_cam,
                          const B2SVisualTable* vis, unsigned outputs, RasterGroup** out, B2SRenderTargets* targets) {
  if (outputs == 0 || (outputs & ~(unsigned)(B2S_OUT_COLOR | B2S_OUT_POSSEG | B2S_OUT_RGB | B2S_OUT_DEPTH | B2S_OUT_SEG))) return "bad render output mask";
  if (vis->n_visual > 64) return "more than 64 render shapes per sub-scene";
  RasterGroup* g = new RasterGroup();
  RasterModel& R = g->R;
  memset(&R, 0, sizeof(R));
  const size_t N = M.n_envs;
  R.n_envs = M.n_envs; R.n_cam = n_cam; R.n_vis = vis->n_visual; R.n_vert = vis->n_vert; R.n_tri = vis->n_tri; R.n_rows = M.n_rows; R.n_ov = vis->n_ov;
  R.vis_type = dev_copy(g, vis->type, vis->n_visual); R.vis_row = dev_copy(g, vis->row, vis->n_visual);
  R.vis_pose = dev_copy(g, vis->pose, (size_t)vis->n_visual * 7); R.vis_size = dev_copy(g, vis->size, (size_t)vis->n_visual * 3);
  R.vis_color = dev_copy(g, vis->color, (size_t)vis->n_visual * 4); R.vis_seg = dev_copy(g, vis->seg_id, vis->n_visual);
  R.vis_ov = dev_copy(g, vis->ov_slot, vis->n_visual);
  {
    std::vector<float> sz((size_t)vis->n_ov * 3 * N + 1), ps((size_t)vis->n_ov * 7 * N + 1);
    for (size_t e = 0; e < N; e++) {
      for (int k = 0; k < vis->n_ov * 3; k++) sz[(size_t)k * N + e] = vis->ov_size[e * vis->n_ov * 3 + k];
      for (int k = 0; k < vis->n_ov * 7; k++) ps[(size_t)k * N + e] = vis->ov_pose[e * vis->n_ov * 7 + k];
    }
    R.ov_size = dev_copy(g, sz.data(), sz.size());
    R.ov_pose = dev_copy(g, ps.data(), ps.size());
  }
  for (int t = 0; t < vis->n_tri * 3; t++)
    if (vis->tri_idx[t] < 0 || vis->tri_idx[t] >= vis->n_vert) { raster_destroy(g); return "triangle vertex index out of range"; }
  R.vert_local = dev_copy(g, vis->vert_local, (size_t)vis->n_vert * 3);
  R.vert_vis = dev_copy(g, vis->vert_vis, vis->n_vert);
  R.tri_idx = dev_copy(g, vis->tri_idx, (size_t)vis->n_tri * 3);
  {
    // owning visual | face << 8: a triangle of a box lies in the face whose coordinate its three unit-cube corners share
    // (face = 2 * axis + (that coordinate > 0)); hull triangles carry face 0
    std::vector<int> packed(vis->n_tri + 1);
    for (int t = 0; t < vis->n_tri; t++) {
      const int v = vis->tri_vis[t];
      if (v < 0 || v >= vis->n_visual) { raster_destroy(g); return "triangle visual index out of range"; }
      int face = 0;
      if (vis->type[v] == SH_BOX) {
        const float* a = vis->vert_local + 3 * (size_t)vis->tri_idx[3 * t];
        const float* b = vis->vert_local + 3 * (size_t)vis->tri_idx[3 * t + 1];
        const float* c = vis->vert_local + 3 * (size_t)vis->tri_idx[3 * t + 2];
        for (int k = 0; k < 3; k++)
          if (a[k] == b[k] && b[k] == c[k]) face = 2 * k + (a[k] > 0.0f ? 1 : 0);
      }
      packed[t] = v | (face << 8);
    }
    R.tri_vis = dev_copy(g, packed.data(), vis->n_tri);
  }
  std::vector<int> w(n_cam), h(n_cam), mount(n_cam);
  std::vector<float> intr(n_cam * 6), cp(n_cam * 7);
  std::vector<size_t> off(n_cam);
  size_t pix = 0;
  g->max_pixels = 0;
  for (int c = 0; c < n_cam; c++) {
    w[c] = cams[c].width; h[c] = cams[c].height; mount[c] = cams[c].mount_row;
    intr[6 * c] = cams[c].fx; intr[6 * c + 1] = cams[c].fy; intr[6 * c + 2] = cams[c].cx; intr[6 * c + 3] = cams[c].cy;
    intr[6 * c + 4] = cams[c].near_; intr[6 * c + 5] = cams[c].far_;
    for (int k = 0; k < 7; k++) cp[7 * c + k] = cams[c].local_pose[k];
    off[c] = pix;
    pix += (size_t)w[c] * h[c];
    if (w[c] * h[c] > g->max_pixels) g->max_pixels = w[c] * h[c];
  }
  if ((size_t)g->max_pixels * 4 > 190 * 1024) { raster_destroy(g); return "camera image does not fit the shared-memory depth buffer (max ~220x220)"; }
  R.cam_w = dev_copy(g, w.data(), n_cam); R.cam_h = dev_copy(g, h.data(), n_cam); R.cam_mount = dev_copy(g, mount.data(), n_cam);
  R.cam_intr = dev_copy(g, intr.data(), intr.size()); R.cam_pose = dev_copy(g, cp.data(), cp.size());
  R.cam_offset = dev_copy(g, off.data(), n_cam);
  R.pixels_per_env = pix;
  memset(&g->T, 0, sizeof(g->T));
  g->T.mask = outputs;
  if (outputs & B2S_OUT_COLOR) g->T.color = dev_copy<uint8_t>(g, nullptr, N * pix * 4);
  if (outputs & B2S_OUT_POSSEG) g->T.posseg = dev_copy<int16_t>(g, nullptr, N * pix * 4);
  if (outputs & B2S_OUT_RGB) g->T.rgb = dev_copy<uint8_t>(g, nullptr, N * pix * 3 + 4);
  if (outputs & B2S_OUT_DEPTH) g->T.depth = dev_copy<int16_t>(g, nullptr, N * pix + 2);
  if (outputs & B2S_OUT_SEG) g->T.seg = dev_copy<int16_t>(g, nullptr, N * pix + 2);
  g->work = dev_copy<int>(g, nullptr, 2);
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long n_img = (long)N * n_cam;
    g->n_ctas = (int)(n_img < 2L * sms ? n_img : 2L * sms);
  }
  for (void* p : g->allocs)
    if (!p) { raster_destroy(g); return "rasteriser allocation failed"; }
  if (cudaFuncSetAttribute(raster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g->max_pixels * 4) != cudaSuccess ||
      cudaFuncSetAttribute(raster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g->max_pixels * 4) != cudaSuccess) {
    raster_destroy(g);
    return "cannot reserve shared memory for the depth buffer";
  }
  targets->color = g->T.color;
  targets->position_seg = g->T.posseg;
  targets->rgb = g->T.rgb;
  targets->depth = g->T.depth;
  targets->segmentation = g->T.seg;
  (void)S; (void)host;
  *out = g;
  return nullptr;
}

const char* raster_run(const DevModel& M, const DevState& S, RasterGroup* g, const uint8_t* env_mask, cudaStream_t st) {
  const int grid = g->n_ctas;
  // 512 threads = 16 warps per image: two images per SM (shared-memory bound) keep 32 warps in flight
  // bounding boxes above this many pixels leave the one-thread path for the warp path; neighbouring triangles of the list belong to the
  // same hull and have similar sizes, so the lanes of a warp stay balanced well beyond one warp's worth of pixels (measured)
  static int big = getenv("B2S_RASTER_BIG") ? atoi(getenv("B2S_RASTER_BIG")) : B2S_BIG_TRI_PIXELS;
  // faces of a box whose screen rectangle is larger than this are tested per pixel; part of the rasteriser's definition (the CPU
  // restatement uses the same constant; the override exists for tuning runs only)
  static int patch = getenv("B2S_RASTER_PATCH") ? atoi(getenv("B2S_RASTER_PATCH")) : B2S_PATCH_PIXELS;
  if (g->T.mask & (B2S_OUT_COLOR | B2S_OUT_POSSEG)) raster_kernel<true><<<grid, B2S_RASTER_THREADS, (size_t)g->max_pixels * 4, st>>>(g->R, S.body_data, g->T, env_mask, big, patch, g->work);
  else raster_kernel<false><<<grid, B2S_RASTER_THREADS, (size_t)g->max_pixels * 4, st>>>(g->R, S.body_data, g->T, env_mask, big, patch, g->work);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

void raster_destroy(RasterGroup* g) {
  if (!g) return;
  for (void* p : g->allocs)
    if (p) cudaFree(p);
  delete g;
}

}  // namespace b2s
