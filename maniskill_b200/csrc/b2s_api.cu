// libb200sim.so -- C-ABI (include/b200sim.h) over the sm_100a kernels.
//
// b2s_step launches the pipelined substep (b2s_pipe.cuh, b2s_solve.cuh): per physics substep kin -> collide -> manifest ->
// rowfill -> solve, each at its own width (lane per sub-scene / per candidate pair / per row / 4 lanes per sub-scene), exchanging
// L2-resident env-major struct-of-arrays buffers.  State is env-major SoA so a warp's access to a slot is one coalesced line.
// B2S_FUSED=1 selects the single-kernel form (one lane per sub-scene, everything in per-lane scratch) kept for comparison.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "b2s_raster.cuh"
#include "b2s_world.inl"
#include "b2s_pipe.cuh"

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

#define B2S_IK_MAX_ELEM 32
#define B2S_IK_MAX_CTRL 12
struct IKChainDev {
  int n_elem, n_ctrl;
  float lambda, alpha;
  float *origin, *axis;   // device
  int *kind, *column;
  unsigned char* controlled;
};

struct PickTaskDev {
  int n_action;
  int *dof_action, *dof_use_delta, *dof_normalize;
  float *dof_low, *dof_high;
  B2SPickTask task;
  B2SPickReset reset;
  int has_reset;
};

struct World : b2s::WorldT<DevMem> {
  int device;
  // the dynamics half of kin (ABA, M~^-1) runs on a side stream while collide + manifest run on the caller's stream
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // captured control steps, keyed by (substeps, fetch_mask); `cap` is the stream they are captured on
  cudaStream_t cap = nullptr;
  std::vector<IKChainDev> iks;
  bool pdl_refused = false;
  // dynamics half of kin: 0 = one lane per sub-scene (kin_kernel<.., 2>, default), 1 = eight lanes per sub-scene (kin_dyn_kernel).  The group
  // kernel alone is 1.5x (9 joints) to 2.6x (24 joints) faster, but next to collide + manifest it slows the control step
  int kin_group = 0;  // capture / instantiation with programmatic dependencies failed once: plain edges from then on
  std::unordered_map<uint64_t, cudaGraphExec_t> graphs;
  std::vector<Query> queries;
  std::vector<b2s::RasterGroup*> groups;
  std::vector<PickTaskDev> pick_tasks;
};

std::mutex g_mu;
std::unordered_map<uint64_t, World*> g_worlds;
uint64_t g_next = 1;

// Every entry point runs on the world's own device whatever the caller's current device is (the reference's
// `physx_cuda:n` backend, mani_skill/envs/utils/system/backend.py:46-68), and leaves the caller's device as it found it.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

World* get(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_worlds.find(h);
  return it == g_worlds.end() ? nullptr : it->second;
}

// ---- programmatic dependent launch (sm_90+): the kernels of a control step form a chain; launched with the
// programmatic-stream-serialization attribute, the next kernel of the chain is scheduled while its predecessor still runs and parks at
// griddepcontrol.wait until the predecessor has completed and flushed its writes -- the launch latency and block scheduling of 31 graph
// nodes no longer sit between the kernels.  Every chained kernel starts with pdl_enter(): wait first, then allow ITS dependents to be
// scheduled (one kernel of look-ahead, not the whole graph).  Without the attribute both instructions are no-ops.
__device__ int g_pdl_early_trigger = 0;
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (g_pdl_early_trigger) asm volatile("griddepcontrol.launch_dependents;");
}
static int pdl_enabled() {
  static int on = getenv("B2S_PDL") ? atoi(getenv("B2S_PDL")) : 0;
  return on;
}
template <class... KArgs, class... Args>
static cudaError_t launch_chain(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <class C, int ND>
__global__ void __launch_bounds__(32) step_kernel(b2s::DevModel M, b2s::DevState S, int substeps, unsigned fetch_mask) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  b2s::step_env<C, ND>(M, S, env, substeps, fetch_mask);
}

// ---- phase B: 4 lanes per sub-scene run the Gauss-Seidel sweeps and integrate (b2s_solve.cuh).  Every sub-scene is resident at
// once, so the launch lasts as long as the slowest one: what counts is the dependent chain of a row visit.  Measured on B200
// (PickCube-v1, ms per control step, 4096 / 16384 envs): 4 lanes 0.93 / 1.78, 8 lanes 0.93 / 1.82, 16 lanes 1.00 / 2.21; one lane
// per sub-scene with instruction-level parallelism instead of shuffles 1.4 / 2.2.
#define B2S_SOLVE_EPB 16  // sub-scenes per block (2 warps of 8 groups with 4 lanes; the impulse table of a block is EPB x MAXROW x 8 B of shared memory)
template <int NUQ, int L>
__global__ void __launch_bounds__(B2S_SOLVE_EPB * L) solve_kernel(b2s::DevModel M, b2s::DevState S) {
  pdl_enter();
  constexpr int MR = b2s::CapsS::MAXROW;
  constexpr int EPB = B2S_SOLVE_EPB;
  __shared__ b2s::LamTot s_lt[EPB][MR];
  __shared__ float s_stage[EPB][2 * NUQ];
  const int g = threadIdx.x / L, lane = threadIdx.x % L;
  const int env = blockIdx.x * EPB + g;
  const bool valid = env < M.n_envs;
  int n_row = valid ? S.sol_nrow[env] : 0;
  // the row loop bound must be uniform over the warp (the group reduction uses full-warp shuffles)
  int nmax = n_row;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  b2s::solve_env<L, NUQ, MR>(M, S, env, lane, valid, nmax, s_lt[g], s_stage[g]);
}

static int solve_lanes() {
  static int forced = getenv("B2S_SOLVE_L") ? atoi(getenv("B2S_SOLVE_L")) : 0;
  return forced == 2 ? 2 : 4;
}
static cudaError_t launch_solve(const b2s::DevModel& M, const b2s::DevState& S, cudaStream_t st, bool pdl) {
  const int N = M.n_envs;
  const int L = solve_lanes();
  const dim3 grid((N + B2S_SOLVE_EPB - 1) / B2S_SOLVE_EPB), block(B2S_SOLVE_EPB * L);
  if (L == 2) {
    if (M.n_u > 16) return launch_chain(pdl, solve_kernel<32, 2>, grid, block, 0, st, M, S);
    return launch_chain(pdl, solve_kernel<16, 2>, grid, block, 0, st, M, S);
  }
  if (M.n_u > 16) return launch_chain(pdl, solve_kernel<32, 4>, grid, block, 0, st, M, S);
  return launch_chain(pdl, solve_kernel<16, 4>, grid, block, 0, st, M, S);
}

// ---- pipelined phase A (b2s_pipe.cuh): kin (lane per sub-scene) -> collide (lane per candidate pair x sub-scene) ->
// manifest (lane per sub-scene) -> rowfill (lane per row x sub-scene); phase B is the solve_kernel above
template <class C, int ND, int PART>
__global__ void __launch_bounds__(32) kin_kernel(b2s::DevModel M, b2s::DevState S) {
  pdl_enter();
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  b2s::kin_env<C, ND, PART>(M, S, env);
}

// the dynamics half of kin with 8 lanes per sub-scene (b2s_pipe.cuh kin_dyn_group): the articulated inertias in shared memory, lane r
// owns row r of the 6x6 sweeps, reductions by shuffles inside the 8-lane group, one M~^-1 column per lane
// Sub-scenes per CTA: the scratch is 7.5 KB (12 joints) / 15 KB (24 joints) per sub-scene; 30 KB per CTA keeps 7 CTAs per SM so that the 1024 CTAs of
// 4096 (2048 large) sub-scenes are resident in one wave -- with 8 sub-scenes per CTA the launch ran in two waves and was slower than the one-lane kernel
#define B2S_KIN_G 8
template <class C, int EPB>
__global__ void __launch_bounds__(B2S_KIN_G * EPB) kin_dyn_kernel(b2s::DevModel M, b2s::DevState S) {
  constexpr int B2S_KIN_EPB = EPB;
  pdl_enter();
  extern __shared__ __align__(16) unsigned char kin_dyn_smem[];
  b2s::KinDynScratch<C, B2S_KIN_G>* W = reinterpret_cast<b2s::KinDynScratch<C, B2S_KIN_G>*>(kin_dyn_smem);
  const int g = threadIdx.x / B2S_KIN_G, lane = threadIdx.x % B2S_KIN_G;
  const int env = blockIdx.x * B2S_KIN_EPB + g;
  if (env >= M.n_envs) return;
  b2s::kin_dyn_group<C, B2S_KIN_G>(M, S, env, lane, b2s::group_mask<B2S_KIN_G>(threadIdx.x & 31), W[g]);
}
template <class C, int EPB>
static cudaError_t launch_kin_dyn(const b2s::DevModel& M, const b2s::DevState& S, cudaStream_t st) {
  const size_t smem = sizeof(b2s::KinDynScratch<C, B2S_KIN_G>) * EPB;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kin_dyn_kernel<C, EPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_chain(false, kin_dyn_kernel<C, EPB>, dim3((M.n_envs + EPB - 1) / EPB), dim3(B2S_KIN_G * EPB), smem, st, M, S);
}

#define B2S_COLLIDE_THREADS 64
__global__ void __launch_bounds__(B2S_COLLIDE_THREADS) collide_kernel(b2s::DevModel M, b2s::DevState S) {
  pdl_enter();
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  b2s::collide_env(M, S, env, blockIdx.y);
}

template <class C>
__global__ void __launch_bounds__(64) manifest_kernel(b2s::DevModel M, b2s::DevState S) {
  pdl_enter();
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  b2s::manifest_env<C>(M, S, env);
}

template <class C, int ND, int NUQ, int L>
__global__ void __launch_bounds__(128) rowfill_kernel(b2s::DevModel M, b2s::DevState S) {
  pdl_enter();
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  const int r = blockIdx.y;
  if (r >= S.sol_nrow[env]) return;
  b2s::rowfill_env<C, ND, NUQ, L>(M, S, env, r);
}

// gpu_fetch_*: internal env-major SoA state -> the exposed AoS buffers.  One lane computes one sub-scene (FK + link velocities); its
// [n_rows][13] block of rigid_body_data (936 B for PickCube-v1) is written to a shared-memory tile, and the tile of the CTA's 32
// sub-scenes -- contiguous in HBM, 32 x n_rows x 52 B, a multiple of 16 -- leaves with ONE bulk asynchronous copy
// (cp.async.bulk.global.shared::cta, the TMA engine) instead of 32 lanes storing single floats 936 B apart.  The staged path needs
// the whole rows to be rewritten (mask has BUF_LINK and BUF_RIGID) and a full CTA; otherwise the lanes store directly.
#define B2S_FETCH_EPB 32
template <class C>
__global__ void __launch_bounds__(B2S_FETCH_EPB) fetch_kernel(b2s::DevModel M, b2s::DevState S, unsigned mask, const uint8_t* __restrict__ only) {
  pdl_enter();
  extern __shared__ __align__(128) float fetch_tile[];
  const int env0 = blockIdx.x * B2S_FETCH_EPB, env = env0 + threadIdx.x;
  // `only` (nullable): CTAs none of whose sub-scenes is flagged leave their rows as they are (refresh after a partial reset)
  if (only && !__syncthreads_or(env < M.n_envs && only[env])) return;
  const int per_env = M.n_rows * 13;
  const unsigned tile_bytes = (unsigned)(B2S_FETCH_EPB * per_env * sizeof(float));
  const bool staged = (mask & b2s::BUF_LINK) && (mask & b2s::BUF_RIGID) && env0 + B2S_FETCH_EPB <= M.n_envs && (tile_bytes % 16u) == 0;
  if (env < M.n_envs) b2s::fetch_env<C>(M, S, env, mask, staged ? fetch_tile + (size_t)threadIdx.x * per_env : nullptr);
  if (!staged) return;
  // generic-proxy writes to shared memory must be visible to the async proxy before the bulk copy reads them
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    float* dst = S.body_data + (size_t)env0 * per_env;
    const unsigned src = (unsigned)__cvta_generic_to_shared(fetch_tile);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(tile_bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile must stay alive until the engine has read it
  }
}
static size_t fetch_smem(const b2s::DevModel& M) { return (size_t)B2S_FETCH_EPB * M.n_rows * 13 * sizeof(float); }

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
      const bool any = rb == B2S_ANY_BODY;  // net impulse on ra: every patch that has ra on one side
      float sg = (a == ra && (any || b == rb)) ? 1.f : ((b == ra && (any || a == rb)) ? -1.f : 0.f);
      sx += sg * o[2 * N]; sy += sg * o[3 * N]; sz += sg * o[4 * N];
    }
    float* w = out + ((size_t)env * nq + q) * 3;
    w[0] = sx; w[1] = sy; w[2] = sz;
  }
}

// ---- fused control step, launch 1: joint-space controller (pd_joint_pos.py:76-93,207-228; gym_utils.py:104-108).
// Reads the internal qpos, writes the internal target (what the substep kernel consumes) and the exposed target_qpos.
__global__ void controller_kernel(b2s::DevModel M, b2s::DevState S, PickTaskDev T, const float* __restrict__ actions) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  const size_t N = M.n_envs;
  const int md = M.max_dof_per_art;
  for (int a = 0; a < M.n_art; a++) {
    int d0 = M.art_dof_start[a], d1 = M.art_dof_start[a + 1];
    for (int i = d0; i < d1; i++) {
      int col = T.dof_action[i];
      if (col < 0) continue;
      float act = actions[(size_t)env * T.n_action + col];
      if (T.dof_normalize[i]) {
        act = fminf(fmaxf(act, -1.f), 1.f);
        float lo = T.dof_low[i], hi = T.dof_high[i];
        act = 0.5f * (hi + lo) + 0.5f * (hi - lo) * act;
      }
      float tq = T.dof_use_delta[i] ? S.q[i * N + env] + act : act;
      S.tq[i * N + env] = tq;
      S.xtq[((size_t)env * M.n_art + a) * md + i - d0] = tq;
    }
  }
}

// ---- fused control step, launch 3: evaluate + reward + observation of the pick task (pick_cube.py:132-191,
// panda.py:237-269, base_agent.py:339-347).  Reads the freshly fetched exposed buffers, so the python path sees the same data.
// `only` (nullable): second use after a device-side auto-reset -- only the sub-scenes with only[env] != 0 are visited and only their
// observation row is rewritten (reward, flags and elapsed keep describing the finished step / were set by the reset).
__global__ void pick_epilogue_kernel(b2s::DevModel M, b2s::DevState S, PickTaskDev T, float* __restrict__ obs, float* __restrict__ reward,
                                     uint8_t* __restrict__ flags, int* __restrict__ elapsed, const uint8_t* __restrict__ only) {
  using namespace b2s;
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  if (only && !only[env]) return;
  const size_t N = M.n_envs;
  const B2SPickTask& K = T.task;
  const int nr = M.n_rows, nd = M.n_dof, md = M.max_dof_per_art;
  const float* body = S.body_data + (size_t)env * nr * 13;
  const float* tcp = body + K.tcp_row * 13;
  const float* obj = body + K.obj_row * 13;
  const float* goal = body + K.goal_row * 13;
  // contact forces finger <-> object (sum of patch impulses / dt, acting on the finger)
  v3 lf = mk3(0, 0, 0), rf = mk3(0, 0, 0);
  int nm = S.man_count[env];
  for (int m = 0; m < nm; m++) {
    const float* o = S.man + (size_t)(m * 8) * N + env;
    int a = (int)o[0], b = (int)o[N];
    v3 imp = mk3(o[2 * N], o[3 * N], o[4 * N]);
    if (a == K.lfinger_row && b == K.obj_row) lf = lf + imp;
    else if (a == K.obj_row && b == K.lfinger_row) lf = lf - imp;
    if (a == K.rfinger_row && b == K.obj_row) rf = rf + imp;
    else if (a == K.obj_row && b == K.rfinger_row) rf = rf - imp;
  }
  float inv_dt = 1.f / M.dt;
  lf = lf * inv_dt;
  rf = rf * inv_dt;
  float lforce = norm(lf), rforce = norm(rf);
  const float* l13 = body + K.lfinger_row * 13;
  const float* r13 = body + K.rfinger_row * 13;
  v3 ldir = col(qmat(mkq(l13[3], l13[4], l13[5], l13[6])), 1);
  v3 rdir = -col(qmat(mkq(r13[3], r13[4], r13[5], r13[6])), 1);
  // common.compute_angle_between: normalise (zero stays zero), clip, acos
  v3 lfn = lforce < 1e-6f ? mk3(0, 0, 0) : lf * (1.f / lforce), rfn = rforce < 1e-6f ? mk3(0, 0, 0) : rf * (1.f / rforce);
  float ldn = norm(ldir), rdn = norm(rdir);
  v3 ldu = ldn < 1e-6f ? mk3(0, 0, 0) : ldir * (1.f / ldn), rdu = rdn < 1e-6f ? mk3(0, 0, 0) : rdir * (1.f / rdn);
  const float rad2deg = 57.29577951308232f;
  float langle = acosf(fminf(fmaxf(dot(ldu, lfn), -1.f), 1.f)) * rad2deg;
  float rangle = acosf(fminf(fmaxf(dot(rdu, rfn), -1.f), 1.f)) * rad2deg;
  bool is_grasped = (lforce >= K.min_force && langle <= K.max_angle_deg) && (rforce >= K.min_force && rangle <= K.max_angle_deg);
  v3 tcp_p = mk3(tcp[0], tcp[1], tcp[2]), obj_p = mk3(obj[0], obj[1], obj[2]), goal_p = mk3(goal[0], goal[1], goal[2]);
  float obj_to_goal = norm(goal_p - obj_p);
  bool is_obj_placed = obj_to_goal <= K.goal_thresh;
  const float* qv = S.xqd + (size_t)env * M.n_art * md;  // articulation 0 = the robot
  const float* qp = S.xq + (size_t)env * M.n_art * md;
  float vmax = 0.f, vsq = 0.f;
  for (int i = 0; i < K.n_static_dof; i++) { vmax = fmaxf(vmax, fabsf(qv[i])); vsq += qv[i] * qv[i]; }
  bool is_static = vmax <= K.static_thresh;
  bool success = is_obj_placed && is_static;
  // dense reward (pick_cube.py:161-191)
  float tcp_to_obj = norm(obj_p - tcp_p);
  float r = 1.f - tanhf(5.f * tcp_to_obj);
  r += is_grasped ? 1.f : 0.f;
  r += (1.f - tanhf(5.f * obj_to_goal)) * (is_grasped ? 1.f : 0.f);
  r += (1.f - tanhf(5.f * sqrtf(vsq))) * (is_obj_placed ? 1.f : 0.f);
  if (success) r = 5.f;
  if (K.normalized_reward) r = r / 5.f;
  if (!only) {
    reward[env] = r;
    int el = elapsed[env] + 1;
    elapsed[env] = el;
    uint8_t* f = flags + (size_t)env * 6;
    f[0] = success; f[1] = is_obj_placed; f[2] = is_static; f[3] = is_grasped; f[4] = success;
    f[5] = (K.max_episode_steps > 0 && el >= K.max_episode_steps) ? 1 : 0;
  }
  // observation: qpos, qvel, is_grasped, tcp_pose, goal_pos, obj_pose, tcp_to_obj_pos, obj_to_goal_pos
  int d0 = M.art_dof_start[0], d1 = M.art_dof_start[1], nq = d1 - d0;
  float* o = obs + (size_t)env * (2 * nq + 24);
  for (int i = 0; i < nq; i++) { o[i] = qp[i]; o[nq + i] = qv[i]; }
  o += 2 * nq;
  o[0] = is_grasped ? 1.f : 0.f;
  for (int k = 0; k < 7; k++) o[1 + k] = tcp[k];
  for (int k = 0; k < 3; k++) o[8 + k] = goal[k];
  for (int k = 0; k < 7; k++) o[11 + k] = obj[k];
  v3 t2o = obj_p - tcp_p, o2g = goal_p - obj_p;
  o[18] = t2o.x; o[19] = t2o.y; o[20] = t2o.z;
  o[21] = o2g.x; o[22] = o2g.y; o[23] = o2g.z;
  (void)nd;
}

// ---- device-side auto-reset of the pick task (include/b200sim.h B2SPickAutoReset): one lane per sub-scene; lanes whose episode
// goes on return at once.  Writes the INTERNAL state (env-major SoA); the fetch that follows refreshes the exposed buffers.
__global__ void pick_reset_kernel(b2s::DevModel M, b2s::DevState S, PickTaskDev T, const float* __restrict__ obs, uint8_t* __restrict__ flags,
                                  int* __restrict__ elapsed, const float* __restrict__ rnd, float* __restrict__ final_obs,
                                  uint8_t* __restrict__ done, int ignore_terminations, int max_episode_steps) {
  using namespace b2s;
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= M.n_envs) return;
  const size_t N = M.n_envs;
  uint8_t* f = flags + (size_t)env * 6;
  if (max_episode_steps > 0) f[5] = elapsed[env] >= max_episode_steps ? 1 : 0;  // TimeLimit truncation
  const bool d = (f[4] && !ignore_terminations) || f[5];
  done[env] = d ? 1 : 0;
  if (!d) return;
  const B2SPickReset& R = T.reset;
  const int d0 = M.art_dof_start[0], d1 = M.art_dof_start[1], nq = d1 - d0;
  const int od = 2 * nq + 24;
  for (int k = 0; k < od; k++) final_obs[(size_t)env * od + k] = obs[(size_t)env * od + k];
  const float* u = rnd + (size_t)env * 24;
  // robot: rest configuration + Gaussian noise (Box-Muller on the caller's uniforms), fingers exact, zero velocity / force,
  // drive targets hold the reset configuration until the first action
  for (int i = 0; i < nq; i++) {
    float q = R.rest_qpos[i];
    if (i < nq - 2) {
      const float u1 = fmaxf(u[6 + 2 * i], 1.0e-7f), u2 = u[7 + 2 * i];
      q += R.robot_qpos_noise * sqrtf(-2.f * logf(u1)) * cosf(6.283185307179586f * u2);
    }
    const size_t o = (size_t)(d0 + i) * N + env;
    S.q[o] = q; S.qd[o] = 0.f; S.qf[o] = 0.f; S.tq[o] = q; S.tqd[o] = 0.f; S.qacc[o] = 0.f;
  }
  // object: xy uniform in the spawn square, resting on the table, random yaw (random_quaternions(lock_x, lock_y)); zero velocity
  const float hs = R.cube_spawn_half_size;
  const float cx = R.cube_spawn_center[0] + (2.f * u[0] - 1.f) * hs, cy = R.cube_spawn_center[1] + (2.f * u[1] - 1.f) * hs, cz = R.cube_half_size;
  const float yaw = 6.283185307179586f * u[2];
  float* ob = S.fb + (size_t)(R.obj_fb * 13) * N + env;
  const float po[13] = {cx, cy, cz, cosf(0.5f * yaw), 0.f, 0.f, sinf(0.5f * yaw), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 13; k++) ob[(size_t)k * N] = po[k];
  // goal: xy uniform in the same square, height uniform in [0, max_goal_height] above the object's centre
  float* gb = S.fb + (size_t)(R.goal_fb * 13) * N + env;
  const float pg[13] = {R.cube_spawn_center[0] + (2.f * u[3] - 1.f) * hs, R.cube_spawn_center[1] + (2.f * u[4] - 1.f) * hs,
                        u[5] * R.max_goal_height + cz, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 13; k++) gb[(size_t)k * N] = pg[k];
  elapsed[env] = 0;
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
  DeviceGuard guard_(device);
  World* w = new World();
  w->device = device;
  w->kin_group = getenv("B2S_KIN_GROUP") ? atoi(getenv("B2S_KIN_GROUP")) : 0;  // read per world (tests compare both forms in one process)
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
  if (fetch_smem(w->M) > 48 * 1024) {
    cudaFuncSetAttribute(fetch_kernel<b2s::CapsS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fetch_smem(w->M));
    cudaFuncSetAttribute(fetch_kernel<b2s::CapsL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fetch_smem(w->M));
  }
  if (w->caps == 0) fetch_kernel<b2s::CapsS><<<(N + B2S_FETCH_EPB - 1) / B2S_FETCH_EPB, B2S_FETCH_EPB, fetch_smem(w->M)>>>(w->M, w->S, 0xFFFFFFFFu, nullptr);
  else fetch_kernel<b2s::CapsL><<<(N + B2S_FETCH_EPB - 1) / B2S_FETCH_EPB, B2S_FETCH_EPB, fetch_smem(w->M)>>>(w->M, w->S, 0xFFFFFFFFu, nullptr);
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
  DeviceGuard guard_(w->device);
  cudaDeviceSynchronize();
  for (auto& kv : w->graphs) cudaGraphExecDestroy(kv.second);
  if (w->cap) cudaStreamDestroy(w->cap);
  if (w->side) cudaStreamDestroy(w->side);
  if (w->ev_fork) cudaEventDestroy(w->ev_fork);
  if (w->ev_join) cudaEventDestroy(w->ev_join);
  for (auto& q : w->queries) cudaFree(q.rows_dev);
  for (auto& k : w->iks) { cudaFree(k.origin); cudaFree(k.axis); cudaFree(k.kind); cudaFree(k.column); cudaFree(k.controlled); }
  for (auto* g : w->groups) b2s::raster_destroy(g);
  w->release();
  delete w;
  return B2S_OK;
}

int32_t b2s_world_buffers(uint64_t world, B2SBufferTable* out) {
  World* w = get(world);
  if (!w || !out) return fail(B2S_ERR_INVALID, "unknown world");
  DeviceGuard guard_(w->device);
  out->rigid_body_data = w->S.body_data;
  out->qpos = w->S.xq; out->qvel = w->S.xqd; out->qacc = w->S.xqacc; out->qf = w->S.xqf;
  out->target_qpos = w->S.xtq; out->target_qvel = w->S.xtqd;
  out->n_rows = w->M.n_rows;
  out->max_dof = w->M.max_dof_per_art;
  out->contact_count = w->S.man_count;
  out->overflow_flag = w->S.overflow;
  return B2S_OK;
}

// Enqueues the kernels of `substeps` physics substeps (+ the fetch) on `st`.  The dynamics half of kin runs on the world's side
// stream between a fork and a join event; under stream capture the same calls become the dependency edges of the graph.
static int enqueue_step(World* w, int substeps, unsigned fetch_mask, cudaStream_t st, bool pdl = false) {
  const int N = w->M.n_envs;
  static int fused = getenv("B2S_FUSED") ? atoi(getenv("B2S_FUSED")) : 0;
  if (!fused && w->M.n_u <= 32) {
    const int MR = b2s::CapsS::MAXROW;
    static int overlap_on = getenv("B2S_OVERLAP") ? atoi(getenv("B2S_OVERLAP")) : 1;
    if (overlap_on && !w->side) {
      CK(cudaStreamCreateWithFlags(&w->side, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&w->ev_fork, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&w->ev_join, cudaEventDisableTiming));
    }
    const bool overlap = overlap_on != 0;
    const b2s::DevModel& M = w->M;
    const b2s::DevState& S = w->S;
    for (int sidx = 0; sidx < substeps; sidx++) {
      const dim3 pg((N + 31) / 32), b32(32);
      // the chain on `st` is launched with programmatic dependencies (pdl); the dynamics half of kin runs on the side stream behind a
      // full event dependency
#define B2S_KIN(PART_, STREAM_, PDL_)                                                                                        \
  {                                                                                                                          \
    if (w->caps == 0 && M.n_dof == 9) { CK(launch_chain(PDL_, kin_kernel<b2s::CapsS, 9, PART_>, pg, b32, 0, STREAM_, M, S)); } \
    else if (w->caps == 0) { CK(launch_chain(PDL_, kin_kernel<b2s::CapsS, 0, PART_>, pg, b32, 0, STREAM_, M, S)); } \
    else { CK(launch_chain(PDL_, kin_kernel<b2s::CapsL, 0, PART_>, pg, b32, 0, STREAM_, M, S)); } \
  }
      if (overlap) {
        B2S_KIN(1, st, pdl)
        CK(cudaEventRecord(w->ev_fork, st));
        CK(cudaStreamWaitEvent(w->side, w->ev_fork, 0));
        if (w->kin_group) {
          if (w->caps == 0) { CK((launch_kin_dyn<b2s::CapsS, 4>(M, S, w->side))); }
          else { CK((launch_kin_dyn<b2s::CapsL, 2>(M, S, w->side))); }
        } else {
          B2S_KIN(2, w->side, false)
        }
        CK(cudaEventRecord(w->ev_join, w->side));
      } else {
        B2S_KIN(0, st, pdl)
      }
#undef B2S_KIN
      if (M.n_pair > 0)
        CK(launch_chain(pdl, collide_kernel, dim3((N + B2S_COLLIDE_THREADS - 1) / B2S_COLLIDE_THREADS, M.n_pair), dim3(B2S_COLLIDE_THREADS), 0, st, M, S));
      if (w->caps == 0) { CK(launch_chain(pdl, manifest_kernel<b2s::CapsS>, pg, b32, 0, st, M, S)); }
      else { CK(launch_chain(pdl, manifest_kernel<b2s::CapsL>, pg, b32, 0, st, M, S)); }
      if (overlap) CK(cudaStreamWaitEvent(st, w->ev_join, 0));
      const dim3 rg((N + 127) / 128, MR), b128(128);
#define B2S_ROWFILL(C_, ND_, NUQ_)                                                                                   \
  {                                                                                                                 \
    if (solve_lanes() == 2) { CK(launch_chain(pdl, rowfill_kernel<C_, ND_, NUQ_, 2>, rg, b128, 0, st, M, S)); } \
    else { CK(launch_chain(pdl, rowfill_kernel<C_, ND_, NUQ_, 4>, rg, b128, 0, st, M, S)); } \
  }
      if (w->caps == 0 && M.n_dof == 9 && M.n_u <= 16) B2S_ROWFILL(b2s::CapsS, 9, 16)
      else if (w->caps == 0 && M.n_u <= 16) B2S_ROWFILL(b2s::CapsS, 0, 16)
      else if (w->caps == 0) B2S_ROWFILL(b2s::CapsS, 0, 32)
      else if (M.n_u <= 16) B2S_ROWFILL(b2s::CapsL, 0, 16)
      else B2S_ROWFILL(b2s::CapsL, 0, 32)
#undef B2S_ROWFILL
      CK(launch_solve(M, S, st, pdl));
    }
    if (fetch_mask) {
      const dim3 fg((N + B2S_FETCH_EPB - 1) / B2S_FETCH_EPB), fb(B2S_FETCH_EPB);
      if (w->caps == 0) { CK(launch_chain(pdl, fetch_kernel<b2s::CapsS>, fg, fb, fetch_smem(M), st, M, S, fetch_mask, (const uint8_t*)nullptr)); }
      else { CK(launch_chain(pdl, fetch_kernel<b2s::CapsL>, fg, fb, fetch_smem(M), st, M, S, fetch_mask, (const uint8_t*)nullptr)); }
    }
    CK(cudaGetLastError());
    return B2S_OK;
  }
  // single-kernel form: one lane per sub-scene, 32-lane CTAs (fewer lanes per warp measured slower: the ~0.5 MB kernel thrashes
  // the instruction caches when several warps per SM run different phases)
  int grid = (N + 31) / 32;
  if (w->caps == 0 && w->M.n_dof == 9) step_kernel<b2s::CapsS, 9><<<grid, 32, 0, st>>>(w->M, w->S, substeps, fetch_mask);
  else if (w->caps == 0) step_kernel<b2s::CapsS, 0><<<grid, 32, 0, st>>>(w->M, w->S, substeps, fetch_mask);
  else step_kernel<b2s::CapsL, 0><<<grid, 32, 0, st>>>(w->M, w->S, substeps, fetch_mask);
  CK(cudaGetLastError());
  return B2S_OK;
}

// The launch sequence of a control step (5 x 6 kernels + fetch) never changes for a world: it is captured once per
// (substeps, fetch_mask) into a CUDA graph on an internal stream and replayed on the caller's stream with one launch.
// B2S_GRAPH=0 issues the kernels directly.
int32_t b2s_step(uint64_t world, int32_t substeps, uint32_t fetch_mask, void* stream) {
  World* w = get(world);
  if (!w) return fail(B2S_ERR_INVALID, "unknown world");
  DeviceGuard guard_(w->device);
  if (substeps < 1) return fail(B2S_ERR_INVALID, "substeps < 1");
  cudaStream_t st = (cudaStream_t)stream;
  static int graph_on = getenv("B2S_GRAPH") ? atoi(getenv("B2S_GRAPH")) : 1;
  if (!graph_on) return enqueue_step(w, substeps, fetch_mask, st);
  const uint64_t key = ((uint64_t)(uint32_t)substeps << 32) | fetch_mask;
  auto it = w->graphs.find(key);
  if (it == w->graphs.end()) {
    if (!w->cap) CK(cudaStreamCreateWithFlags(&w->cap, cudaStreamNonBlocking));
    cudaGraphExec_t ex = nullptr;
    // first with programmatic dependencies between the kernels of the chain; a driver that refuses them in a capture gets plain edges
    if (pdl_enabled()) {
      int early = pdl_enabled() >= 2;  // B2S_PDL=1: dependents are released when a kernel's blocks exit; 2: right after its own wait
      CK(cudaMemcpyToSymbol(g_pdl_early_trigger, &early, sizeof(int)));
    }
    for (int attempt = (pdl_enabled() && !w->pdl_refused) ? 0 : 1; attempt < 2 && !ex; attempt++) {
      const bool pdl = attempt == 0;
      CK(cudaStreamBeginCapture(w->cap, cudaStreamCaptureModeRelaxed));
      int rc = enqueue_step(w, substeps, fetch_mask, w->cap, pdl);
      cudaGraph_t g = nullptr;
      cudaError_t e = cudaStreamEndCapture(w->cap, &g);
      if (rc == B2S_OK && e == cudaSuccess) e = cudaGraphInstantiate(&ex, g, 0);
      if (g) cudaGraphDestroy(g);
      if (rc != B2S_OK || e != cudaSuccess) {
        ex = nullptr;
        cudaGetLastError();
        if (pdl) { w->pdl_refused = true; continue; }
        if (rc != B2S_OK) return rc;
        return fail(B2S_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(e));
      }
    }
    it = w->graphs.emplace(key, ex).first;
  }
  CK(cudaGraphLaunch(it->second, st));
  return B2S_OK;
}

int32_t b2s_apply(uint64_t world, uint32_t mask, void* stream) {
  World* w = get(world);
  if (!w) return fail(B2S_ERR_INVALID, "unknown world");
  DeviceGuard guard_(w->device);
  int N = w->M.n_envs;
  apply_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->M, w->S, mask);
  CK(cudaGetLastError());
  return B2S_OK;
}

static int32_t fetch_rows(World* w, uint32_t mask, const uint8_t* only, cudaStream_t st) {
  const int N = w->M.n_envs;
  if (w->caps == 0) fetch_kernel<b2s::CapsS><<<(N + B2S_FETCH_EPB - 1) / B2S_FETCH_EPB, B2S_FETCH_EPB, fetch_smem(w->M), st>>>(w->M, w->S, mask, only);
  else fetch_kernel<b2s::CapsL><<<(N + B2S_FETCH_EPB - 1) / B2S_FETCH_EPB, B2S_FETCH_EPB, fetch_smem(w->M), st>>>(w->M, w->S, mask, only);
  CK(cudaGetLastError());
  return B2S_OK;
}

int32_t b2s_fetch(uint64_t world, uint32_t mask, void* stream) {
  World* w = get(world);
  if (!w) return fail(B2S_ERR_INVALID, "unknown world");
  DeviceGuard guard_(w->device);
  return fetch_rows(w, mask, nullptr, (cudaStream_t)stream);
}

int32_t b2s_update_kinematics(uint64_t world, void* stream) {
  // link poses are a pure function of (root pose, q): refreshing them is the fetch of the link rows
  return b2s_fetch(world, b2s::BUF_LINK, stream);
}

int32_t b2s_contact_query_create(uint64_t world, const int32_t* rows, int32_t n_query, uint64_t* query) {
  World* w = get(world);
  if (!w || !rows || !query || n_query < 1) return fail(B2S_ERR_INVALID, "bad contact query");
  DeviceGuard guard_(w->device);
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
  DeviceGuard guard_(w->device);
  const Query& q = w->queries[query - 1];
  int N = w->M.n_envs;
  query_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w->M, w->S, q.rows_dev, q.n, out_dev);
  CK(cudaGetLastError());
  return B2S_OK;
}

int32_t b2s_camera_group_create(uint64_t world, const B2SCameraDesc* cams, int32_t n_cam, const B2SVisualTable* vis,
                                uint64_t* group, B2SRenderTargets* out) {
  return b2s_camera_group_create_outputs(world, cams, n_cam, vis, B2S_OUT_RAW, group, out);
}

int32_t b2s_camera_group_create_outputs(uint64_t world, const B2SCameraDesc* cams, int32_t n_cam, const B2SVisualTable* vis,
                                        uint32_t outputs, uint64_t* group, B2SRenderTargets* out) {
  World* w = get(world);
  if (!w || !cams || !vis || !group || !out || n_cam < 1) return fail(B2S_ERR_INVALID, "bad camera group");
  DeviceGuard guard_(w->device);
  b2s::RasterGroup* g = nullptr;
  const char* err = b2s::raster_create(w->M, w->S, w->host, cams, n_cam, vis, outputs, &g, out);
  if (err) return fail(B2S_ERR_INVALID, "%s", err);
  w->groups.push_back(g);
  *group = w->groups.size();
  return B2S_OK;
}

int32_t b2s_render(uint64_t world, uint64_t group, void* stream) {
  World* w = get(world);
  if (!w || group < 1 || group > w->groups.size()) return fail(B2S_ERR_INVALID, "bad camera group");
  DeviceGuard guard_(w->device);
  const char* err = b2s::raster_run(w->M, w->S, w->groups[group - 1], nullptr, (cudaStream_t)stream);
  if (err) return fail(B2S_ERR_CUDA, "%s", err);
  return B2S_OK;
}

int32_t b2s_render_masked(uint64_t world, uint64_t group, const uint8_t* env_mask_dev, void* stream) {
  World* w = get(world);
  if (!w || group < 1 || group > w->groups.size()) return fail(B2S_ERR_INVALID, "bad camera group");
  DeviceGuard guard_(w->device);
  const char* err = b2s::raster_run(w->M, w->S, w->groups[group - 1], env_mask_dev, (cudaStream_t)stream);
  if (err) return fail(B2S_ERR_CUDA, "%s", err);
  return B2S_OK;
}

int32_t b2s_pick_task_create(uint64_t world, const B2SJointController* c, const B2SPickTask* task, uint64_t* handle) {
  World* w = get(world);
  if (!w || !c || !task || !handle) return fail(B2S_ERR_INVALID, "bad pick task");
  DeviceGuard guard_(w->device);
  int nd = w->M.n_dof, nr = w->M.n_rows;
  const int rows[5] = {task->tcp_row, task->obj_row, task->goal_row, task->lfinger_row, task->rfinger_row};
  for (int r : rows)
    if (r < 0 || r >= nr) return fail(B2S_ERR_INVALID, "pick task row out of range");
  if (w->M.n_art < 1 || task->n_static_dof > w->M.max_dof_per_art) return fail(B2S_ERR_INVALID, "pick task needs the robot as articulation 0");
  PickTaskDev T;
  T.n_action = c->n_action;
  T.task = *task;
  T.has_reset = 0;
  memset(&T.reset, 0, sizeof(T.reset));
  T.dof_action = (int*)w->up(c->dof_action, nd); T.dof_use_delta = (int*)w->up(c->dof_use_delta, nd);
  T.dof_normalize = (int*)w->up(c->dof_normalize, nd);
  T.dof_low = (float*)w->up(c->dof_low, nd); T.dof_high = (float*)w->up(c->dof_high, nd);
  if (!T.dof_action || !T.dof_use_delta || !T.dof_normalize || !T.dof_low || !T.dof_high) return fail(B2S_ERR_CUDA, "allocation failed");
  w->pick_tasks.push_back(T);
  *handle = w->pick_tasks.size();
  return B2S_OK;
}

int32_t b2s_pick_task_step(uint64_t world, uint64_t handle, const float* actions_dev, int32_t substeps, const B2SPickOutputs* out,
                           void* stream) {
  World* w = get(world);
  if (!w || handle < 1 || handle > w->pick_tasks.size() || !out || !out->obs || !out->reward || !out->flags || !out->elapsed)
    return fail(B2S_ERR_INVALID, "bad pick task step");
  DeviceGuard guard_(w->device);
  const PickTaskDev& T = w->pick_tasks[handle - 1];
  cudaStream_t st = (cudaStream_t)stream;
  int N = w->M.n_envs;
  if (actions_dev) controller_kernel<<<(N + 127) / 128, 128, 0, st>>>(w->M, w->S, T, actions_dev);
  int32_t rc = b2s_step(world, substeps, 0xFFFFFFFFu, stream);
  if (rc != B2S_OK) return rc;
  pick_epilogue_kernel<<<(N + 127) / 128, 128, 0, st>>>(w->M, w->S, T, out->obs, out->reward, out->flags, out->elapsed, nullptr);
  CK(cudaGetLastError());
  return B2S_OK;
}

int32_t b2s_pick_task_set_reset(uint64_t world, uint64_t handle, const B2SPickReset* reset) {
  World* w = get(world);
  if (!w || handle < 1 || handle > w->pick_tasks.size() || !reset) return fail(B2S_ERR_INVALID, "bad pick task reset");
  const int nq = w->M.n_art > 0 ? w->host.max_dof_per_art : 0;
  if (reset->n_rest < 2 || reset->n_rest > 16 || reset->n_rest > nq) return fail(B2S_ERR_INVALID, "pick task reset: bad rest configuration length");
  if (reset->obj_fb < 0 || reset->obj_fb >= w->M.n_fb || reset->goal_fb < 0 || reset->goal_fb >= w->M.n_fb)
    return fail(B2S_ERR_INVALID, "pick task reset: free-body index out of range");
  PickTaskDev& T = w->pick_tasks[handle - 1];
  T.reset = *reset;
  T.has_reset = 1;
  return B2S_OK;
}

int32_t b2s_pick_task_autoreset(uint64_t world, uint64_t handle, const B2SPickOutputs* out, const B2SPickAutoReset* ar, void* stream) {
  World* w = get(world);
  if (!w || handle < 1 || handle > w->pick_tasks.size() || !out || !out->obs || !out->reward || !out->flags || !out->elapsed || !ar || !ar->rand ||
      !ar->final_obs || !ar->done)
    return fail(B2S_ERR_INVALID, "bad pick task auto-reset");
  if (!w->pick_tasks[handle - 1].has_reset) return fail(B2S_ERR_INVALID, "b2s_pick_task_set_reset has not been called");
  int32_t rc = B2S_OK;
  DeviceGuard guard_(w->device);
  const PickTaskDev& T = w->pick_tasks[handle - 1];
  cudaStream_t st = (cudaStream_t)stream;
  const int N = w->M.n_envs;
  pick_reset_kernel<<<(N + 127) / 128, 128, 0, st>>>(w->M, w->S, T, out->obs, out->flags, out->elapsed, ar->rand, ar->final_obs, ar->done,
                                                     ar->ignore_terminations, ar->max_episode_steps);
  // exposed buffers of the reset sub-scenes from their new internal state (CTAs of 32 sub-scenes none of which was reset exit at once),
  // then their observation rows
  rc = fetch_rows(w, 0xFFFFFFFFu, ar->done, st);
  if (rc != B2S_OK) return rc;
  pick_epilogue_kernel<<<(N + 127) / 128, 128, 0, st>>>(w->M, w->S, T, out->obs, out->reward, out->flags, out->elapsed, ar->done);
  CK(cudaGetLastError());
  return B2S_OK;
}

int32_t b2s_pick_task_step_autoreset(uint64_t world, uint64_t handle, const float* actions_dev, int32_t substeps, const B2SPickOutputs* out,
                                     const B2SPickAutoReset* ar, void* stream) {
  int32_t rc = b2s_pick_task_step(world, handle, actions_dev, substeps, out, stream);
  if (rc != B2S_OK) return rc;
  return b2s_pick_task_autoreset(world, handle, out, ar, stream);
}

// dst[env] = src[env] (row_bytes each, 16-byte multiples) for the sub-scenes with mask[env] != 0.  A fixed grid of CTAs strides over the
// sub-scenes, so a step on which nothing finished costs one small launch instead of tens of thousands of blocks that exit at once.
__global__ void __launch_bounds__(256) masked_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t row_vec, const uint8_t* __restrict__ mask,
                                                          int n_envs) {
  for (int env = blockIdx.x; env < n_envs; env += gridDim.x) {
    if (!mask[env]) continue;
    const size_t base = (size_t)env * row_vec;
    for (size_t i = threadIdx.x; i < row_vec; i += blockDim.x) dst[base + i] = src[base + i];
  }
}

int32_t b2s_masked_copy(uint64_t world, void* dst_dev, const void* src_dev, uint64_t row_bytes, const uint8_t* mask_dev, void* stream) {
  World* w = get(world);
  if (!w || !dst_dev || !src_dev || !mask_dev || row_bytes == 0 || row_bytes % 16 != 0) return fail(B2S_ERR_INVALID, "bad masked copy");
  DeviceGuard guard_(w->device);
  const size_t row_vec = row_bytes / 16;
  const int grid = w->M.n_envs < 148 * 8 ? w->M.n_envs : 148 * 8;
  masked_copy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint4*)dst_dev, (const uint4*)src_dev, row_vec, mask_dev, w->M.n_envs);
  CK(cudaGetLastError());
  return B2S_OK;
}


}  // extern "C"

// One damped least-squares IK step of a serial chain, one lane per sub-scene: forward kinematics down the chain (joint frames, axes and
// anchor points in the root frame), geometric Jacobian columns of the controlled joints, M = J J^T + lambda I (6 x 6, SPD) by Cholesky,
// dq = J^T M^-1 delta.
__global__ void __launch_bounds__(128) ik_dls_kernel(int n_envs, IKChainDev K, const float* __restrict__ delta, const float* __restrict__ qpos, int stride,
                                                     float* __restrict__ target) {
  using namespace b2s;
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n_envs) return;
  v3 p = mk3(0, 0, 0);
  m3 R;
  for (int k = 0; k < 9; k++) R.m[k] = (k % 4 == 0) ? 1.f : 0.f;
  v3 jp[B2S_IK_MAX_CTRL], ja[B2S_IK_MAX_CTRL];
  float qc[B2S_IK_MAX_CTRL];
  bool rev[B2S_IK_MAX_CTRL];
  int nc = 0;
  for (int e = 0; e < K.n_elem; e++) {
    const float* o = K.origin + 7 * e;
    p = p + mul(R, mk3(o[0], o[1], o[2]));
    R = mul(R, qmat(mkq(o[3], o[4], o[5], o[6])));
    const int kind = K.kind[e];
    if (kind == 0) continue;
    const v3 ax = mk3(K.axis[3 * e], K.axis[3 * e + 1], K.axis[3 * e + 2]);
    const v3 aw = mul(R, ax);
    const float q = qpos[(size_t)env * stride + K.column[e]];
    if (K.controlled[e] && nc < B2S_IK_MAX_CTRL) { jp[nc] = p; ja[nc] = aw; qc[nc] = q; rev[nc] = kind == 1; nc++; }
    if (kind == 1) R = mul(R, qmat(qaxis_angle(ax, q)));
    else p = p + aw * q;
  }
  // Jacobian columns (linear rows first, then angular: the pytorch_kinematics convention)
  float J[6][B2S_IK_MAX_CTRL];
  for (int c = 0; c < nc; c++) {
    const v3 lin = rev[c] ? cross(ja[c], p - jp[c]) : ja[c];
    const v3 ang = rev[c] ? ja[c] : mk3(0, 0, 0);
    J[0][c] = lin.x; J[1][c] = lin.y; J[2][c] = lin.z; J[3][c] = ang.x; J[4][c] = ang.y; J[5][c] = ang.z;
  }
  float M[6][6], y[6];
  for (int i = 0; i < 6; i++) {
    y[i] = delta[(size_t)env * 6 + i];
    for (int j = 0; j <= i; j++) {
      float s = i == j ? K.lambda : 0.f;
      for (int c = 0; c < nc; c++) s += J[i][c] * J[j][c];
      M[i][j] = s;
    }
  }
  // Cholesky M = L L^T in place (lower triangle), then L z = delta, L^T y = z
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j <= i; j++) {
      float s = M[i][j];
      for (int k = 0; k < j; k++) s -= M[i][k] * M[j][k];
      M[i][j] = i == j ? sqrtf(fmaxf(s, 1e-20f)) : s / M[j][j];
    }
  }
  for (int i = 0; i < 6; i++) {
    float s = y[i];
    for (int k = 0; k < i; k++) s -= M[i][k] * y[k];
    y[i] = s / M[i][i];
  }
  for (int i = 5; i >= 0; i--) {
    float s = y[i];
    for (int k = i + 1; k < 6; k++) s -= M[k][i] * y[k];
    y[i] = s / M[i][i];
  }
  for (int c = 0; c < nc; c++) {
    float dq = 0.f;
    for (int i = 0; i < 6; i++) dq += J[i][c] * y[i];
    target[(size_t)env * K.n_ctrl + c] = qc[c] + K.alpha * dq;
  }
}

extern "C" {

int32_t b2s_ik_create(uint64_t world, const B2SChainDesc* chain, uint64_t* ik) {
  World* w = get(world);
  if (!w || !chain || !ik || chain->n_elem < 1 || chain->n_elem > B2S_IK_MAX_ELEM || !chain->origin || !chain->axis || !chain->kind || !chain->qpos_column ||
      !chain->controlled)
    return fail(B2S_ERR_INVALID, "bad IK chain (1..32 elements, all tables given)");
  DeviceGuard guard_(w->device);
  IKChainDev K;
  K.n_elem = chain->n_elem;
  K.n_ctrl = 0;
  for (int e = 0; e < chain->n_elem; e++) {
    if (chain->kind[e] < 0 || chain->kind[e] > 2) return fail(B2S_ERR_INVALID, "IK chain: joint kind must be 0, 1 or 2");
    if (chain->kind[e] != 0 && chain->controlled[e]) K.n_ctrl++;
  }
  if (K.n_ctrl < 1 || K.n_ctrl > B2S_IK_MAX_CTRL) return fail(B2S_ERR_INVALID, "IK chain: 1..12 controlled joints");
  K.lambda = chain->lambda;
  K.alpha = chain->alpha;
  const size_t n = chain->n_elem;
  CK(cudaMalloc(&K.origin, n * 7 * sizeof(float))); CK(cudaMalloc(&K.axis, n * 3 * sizeof(float)));
  CK(cudaMalloc(&K.kind, n * sizeof(int))); CK(cudaMalloc(&K.column, n * sizeof(int))); CK(cudaMalloc(&K.controlled, n));
  CK(cudaMemcpy(K.origin, chain->origin, n * 7 * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(K.axis, chain->axis, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(K.kind, chain->kind, n * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(K.column, chain->qpos_column, n * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(K.controlled, chain->controlled, n, cudaMemcpyHostToDevice));
  w->iks.push_back(K);
  *ik = w->iks.size();
  return B2S_OK;
}

int32_t b2s_ik_step(uint64_t world, uint64_t ik, const float* delta_pose_dev, const float* qpos_dev, int32_t qpos_stride, float* target_dev, void* stream) {
  World* w = get(world);
  if (!w || ik < 1 || ik > w->iks.size() || !delta_pose_dev || !qpos_dev || !target_dev || qpos_stride < 1) return fail(B2S_ERR_INVALID, "bad IK step");
  DeviceGuard guard_(w->device);
  const int N = w->M.n_envs;
  ik_dls_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, w->iks[ik - 1], delta_pose_dev, qpos_dev, qpos_stride, target_dev);
  CK(cudaGetLastError());
  return B2S_OK;
}

}  // extern "C"
