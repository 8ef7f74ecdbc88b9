// libb200sim.so -- C-ABI (include/b200sim.h) over the sm_100a kernels.
//
// Launch geometry for the fused substep kernel: one lane per sub-scene, 32-lane CTAs so that n_envs = 4096 spreads as
// 128 CTAs over the 148 SMs (one resident warp per SM; the per-lane scratch of a substep is several KB, so residency
// is bounded by local memory traffic, not by warps).  State is env-major SoA so every global access of a warp is one
// fully-coalesced 128 B line per slot.
#include <cuda_runtime.h>
#include <stdio.h>

#include <mutex>
#include <unordered_map>

#include "b2s_raster.cuh"
#include "b2s_world.inl"

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}
#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) return fail(B2S_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e_)); \
  } while (0)

struct DevMem {
  static void* alloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, n) != cudaSuccess) return nullptr;
    return p;
  }
  static void upload(void* d, const void* s, size_t n) { cudaMemcpy(d, s, n, cudaMemcpyHostToDevice); }
  static void zero(void* d, size_t n) { cudaMemset(d, 0, n); }
  static void release(void* p) { cudaFree(p); }
};

struct Query {
  int n;
  int* rows_dev;
};

struct World : b2s::WorldT<DevMem> {
  int device;
  std::vector<Query> queries;
  std::vector<b2s::RasterGroup*> groups;
};

std::mutex g_mu;
std::unordered_map<uint64_t, World*> g_worlds;
uint64_t g_next = 1;

World* get(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_worlds.find(h);
  return it == g_worlds.end() ? nullptr : it->second;
}

template <class C, int ND>
__global__ void __launch_bounds__(32) step_kernel(b2s::DevModel M, b2s::DevState S, int substeps, unsigned fetch_mask) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  b2s::step_env<C, ND>(M, S, env, substeps, fetch_mask);
}

template <class C>
__global__ void fetch_kernel(b2s::DevModel M, b2s::DevState S, unsigned mask) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  b2s::fetch_env<C>(M, S, env, mask);
}

__global__ void apply_kernel(b2s::DevModel M, b2s::DevState S, unsigned mask) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  b2s::apply_env(M, S, env, mask);
}

// out[env, q, 3] = sum of manifold impulses between rows (a,b), sign = acting on a
__global__ void query_kernel(b2s::DevModel M, b2s::DevState S, const int* rows, int nq, float* out) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  const size_t N = M.n_envs;
  int nm = S.man_count[env];
  for (int q = 0; q < nq; q++) {
    int ra = rows[2 * q], rb = rows[2 * q + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int m = 0; m < nm; m++) {
      const float* o = S.man + (size_t)(m * 8) * N + env;
      int a = (int)o[0], b = (int)o[N];
      float sg = (a == ra && b == rb) ? 1.f : ((a == rb && b == ra) ? -1.f : 0.f);
      sx += sg * o[2 * N]; sy += sg * o[3 * N]; sz += sg * o[4 * N];
    }
    float* w = out + ((size_t)env * nq + q) * 3;
    w[0] = sx; w[1] = sy; w[2] = sz;
  }
}

}  // namespace

extern "C" {

const char* b2s_last_error(void) { return g_err; }
int32_t b2s_version(void) { return 1; }

int32_t b2s_world_create(const B2SModel* model, int32_t device, uint64_t* world) {
  if (!model || !world) return fail(B2S_ERR_INVALID, "null argument");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(B2S_ERR_NO_DEVICE, "no CUDA device: b200sim has no CPU path");
  if (device < 0 || device >= count) return fail(B2S_ERR_INVALID, "bad device index");
  CK(cudaSetDevice(device));
  World* w = new World();
  w->device = device;
  const char* err = w->build(*model);
  if (err) {
    w->release();
    delete w;
    return fail(B2S_ERR_CAPACITY, "%s", err);
  }
  CK(cudaDeviceSynchronize());
  // the per-lane scratch of the fused substep lives in local memory; make sure the stack limit allows it
  size_t need = 96 * 1024;
  size_t cur = 0;
  cudaDeviceGetLimit(&cur, cudaLimitStackSize);
  if (cur < need) cudaDeviceSetLimit(cudaLimitStackSize, need);
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t h = g_next++;
  g_worlds[h] = w;
  *world = h;
  // initial fetch so the exposed buffers are valid right after create (px.gpu_init semantics)
  int N = w->M.n_envs;
  if (w->caps == 0) fetch_kernel<b2s::CapsS><<<(N + 63) / 64, 64>>>(w->M, w->S, 0xFFFFFFFFu);
  else fetch_kernel<b2s::CapsL><<<(N + 63) / 64, 64>>>(w->M, w->S, 0xFFFFFFFFu);
  CK(cudaDeviceSynchronize());
  return B2S_OK;
}

int32_t b2s_world_destroy(uint64_t world) {
  World* w = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_worlds.find(world);
    if (it == g_worlds.end()) return fail(B2S_ERR_INVALID, "unknown world");
    w = it->second;
    g_worlds.erase(it);
  }
  cudaSetDevice(w->device);
  cudaDeviceSynchronize();
  for (auto& q : w->queries) cudaFree(q.rows_dev);
  for (auto* g : w->groups) b2s::raster_destroy(g);
  w->release();
  delete w;
  return B2S_OK;
}

int32_t b2s_world_buffers(uint64_t world, B2SBufferTable* out) {
  World* w = get(world);
  if (!w || !out) return fail(B2S_ERR_INVALID, "unknown world");
  out->rigid_body_data = w->S.body_data;
  out->qpos = w->S.xq; out->qvel = w->S.xqd; out->qacc = w->S.xqacc; out->qf = w->S.xqf;
  out->target_qpos = w->S.xtq; out->target_qvel = w->S.xtqd;
  out->n_rows = w->M.n_rows;
  out->max_dof = w->M.max_dof_per_art;
  out->contact_count = w->S.man_count;
  out->overflow_flag = w->S.overflow;
  return B2S_OK;
}

int32_t b2s_step(uint64_t world, int32_t substeps, uint32_t fetch_mask, void* stream) {
  World* w = get(world);
  if (!w) return fail(B2S_ERR_INVALID, "unknown world");
  if (substeps < 1) return fail(B2S_ERR_INVALID, "substeps < 1");
  int N = w->M.n_envs;
  cudaStream_t st = (cudaStream_t)stream;
  if (w->caps == 0 && w->M.n_dof == 9) step_kernel<b2s::CapsS, 9><<<(N + 31) / 32, 32, 0, st>>>(w->M, w->S, substeps, fetch_mask);
  else if (w->caps == 0) step_kernel<b2s::CapsS, 0><<<(N + 31) / 32, 32, 0, st>>>(w->M, w->S, substeps, fetch_mask);
  else step_kernel<b2s::CapsL, 0><<<(N + 31) / 32, 32, 0, st>>>(w->M, w->S, substeps, fetch_mask);
  CK(cudaGetLastError());
  return B2S_OK;
}

int32_t b2s_apply(uint64_t world, uint32_t mask, void* stream) {
  World* w = get(world);
  if (!w) return fail(B2S_ERR_INVALID, "unknown world");
  int N = w->M.n_envs;
  apply_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->M, w->S, mask);
  CK(cudaGetLastError());
  return B2S_OK;
}

int32_t b2s_fetch(uint64_t world, uint32_t mask, void* stream) {
  World* w = get(world);
  if (!w) return fail(B2S_ERR_INVALID, "unknown world");
  int N = w->M.n_envs;
  cudaStream_t st = (cudaStream_t)stream;
  if (w->caps == 0) fetch_kernel<b2s::CapsS><<<(N + 63) / 64, 64, 0, st>>>(w->M, w->S, mask);
  else fetch_kernel<b2s::CapsL><<<(N + 63) / 64, 64, 0, st>>>(w->M, w->S, mask);
  CK(cudaGetLastError());
  return B2S_OK;
}

int32_t b2s_update_kinematics(uint64_t world, void* stream) {
  // link poses are a pure function of (root pose, q): refreshing them is the fetch of the link rows
  return b2s_fetch(world, b2s::BUF_LINK, stream);
}

int32_t b2s_contact_query_create(uint64_t world, const int32_t* rows, int32_t n_query, uint64_t* query) {
  World* w = get(world);
  if (!w || !rows || !query || n_query < 1) return fail(B2S_ERR_INVALID, "bad contact query");
  Query q;
  q.n = n_query;
  CK(cudaMalloc(&q.rows_dev, sizeof(int) * 2 * n_query));
  CK(cudaMemcpy(q.rows_dev, rows, sizeof(int) * 2 * n_query, cudaMemcpyHostToDevice));
  w->queries.push_back(q);
  *query = w->queries.size();
  return B2S_OK;
}

int32_t b2s_contact_query_run(uint64_t world, uint64_t query, float* out_dev, void* stream) {
  World* w = get(world);
  if (!w || query < 1 || query > w->queries.size() || !out_dev) return fail(B2S_ERR_INVALID, "bad contact query");
  const Query& q = w->queries[query - 1];
  int N = w->M.n_envs;
  query_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->M, w->S, q.rows_dev, q.n, out_dev);
  CK(cudaGetLastError());
  return B2S_OK;
}

int32_t b2s_camera_group_create(uint64_t world, const B2SCameraDesc* cams, int32_t n_cam, const B2SVisualTable* vis,
                                uint64_t* group, B2SRenderTargets* out) {
  World* w = get(world);
  if (!w || !cams || !vis || !group || !out || n_cam < 1) return fail(B2S_ERR_INVALID, "bad camera group");
  b2s::RasterGroup* g = nullptr;
  const char* err = b2s::raster_create(w->M, w->S, w->host, cams, n_cam, vis, &g, out);
  if (err) return fail(B2S_ERR_INVALID, "%s", err);
  w->groups.push_back(g);
  *group = w->groups.size();
  return B2S_OK;
}

int32_t b2s_render(uint64_t world, uint64_t group, void* stream) {
  World* w = get(world);
  if (!w || group < 1 || group > w->groups.size()) return fail(B2S_ERR_INVALID, "bad camera group");
  const char* err = b2s::raster_run(w->M, w->S, w->groups[group - 1], (cudaStream_t)stream);
  if (err) return fail(B2S_ERR_CUDA, "%s", err);
  return B2S_OK;
}

}  // extern "C"
