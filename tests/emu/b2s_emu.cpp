// TEST TOOL ONLY: compiles the device per-env code (maniskill_b200/csrc/*.cuh, all __host__ __device__) for the host
// and loops it over envs, so the CUDA kernel's logic can be compared with the oracle on a machine without a GPU.
// Never loaded by the product; maniskill_b200.backend only loads libb200sim.so and raises when no GPU is present.
#include <stdio.h>

#include "../../maniskill_b200/csrc/b2s_world.inl"
#include "../../maniskill_b200/csrc/b2s_pipe.cuh"

namespace {
struct HostMem {
  static void* alloc(size_t n) { return malloc(n); }
  static void upload(void* d, const void* s, size_t n) { memcpy(d, s, n); }
  static void zero(void* d, size_t n) { memset(d, 0, n); }
  static void release(void* p) { free(p); }
};
typedef b2s::WorldT<HostMem> World;
}  // namespace

// pipelined substep (kin -> collide per pair -> manifest -> rowfill per row -> group solve with L = 1), what the CUDA library runs
template <class C, int ND, int NUQ>
static void pipe_substep(World* w) {
  const int MR = b2s::CapsS::MAXROW;
  b2s::LamTot lamtot[MR];
  float stage[2 * 32];
  for (int e = 0; e < w->M.n_envs; e++) {
    // forward kinematics + joint rows, then the dynamics in the group form the CUDA library runs (here a group of one lane)
    b2s::kin_env<C, ND, 1>(w->M, w->S, e);
    {
      static b2s::KinDynScratch<C, 1> scratch;
      b2s::kin_dyn_group<C, 1>(w->M, w->S, e, 0, 1u, scratch);
    }
    for (int k = 0; k < w->M.n_pair; k++) b2s::collide_env(w->M, w->S, e, k);
    b2s::manifest_env<C>(w->M, w->S, e);
    for (int r = 0; r < w->S.sol_nrow[e]; r++) b2s::rowfill_env<C, ND, NUQ, 1>(w->M, w->S, e, r);
    b2s::solve_env<1, NUQ, MR>(w->M, w->S, e, 0, true, w->S.sol_nrow[e], lamtot, stage);
  }
}

extern "C" {
void* emu_create(const B2SModel* m) {
  World* w = new World();
  const char* err = w->build(*m);
  if (err) { fprintf(stderr, "emu_create: %s\n", err); delete w; return nullptr; }
  return w;
}
void emu_destroy(void* h) { World* w = (World*)h; w->release(); delete w; }
void emu_step(void* h, int substeps, unsigned fetch_mask) {
  World* w = (World*)h;
  for (int e = 0; e < w->M.n_envs; e++) {
    if (w->caps == 0 && w->M.n_dof == 9) b2s::step_env<b2s::CapsS, 9>(w->M, w->S, e, substeps, fetch_mask);
    else if (w->caps == 0) b2s::step_env<b2s::CapsS, 0>(w->M, w->S, e, substeps, fetch_mask);
    else b2s::step_env<b2s::CapsL, 0>(w->M, w->S, e, substeps, fetch_mask);
  }
}
void emu_step_pipe(void* h, int substeps, unsigned fetch_mask) {
  World* w = (World*)h;
  for (int s = 0; s < substeps; s++) {
    const bool small_u = w->M.n_u <= 16;
    if (w->caps == 0 && w->M.n_dof == 9 && small_u) pipe_substep<b2s::CapsS, 9, 16>(w);
    else if (w->caps == 0 && small_u) pipe_substep<b2s::CapsS, 0, 16>(w);
    else if (w->caps == 0) pipe_substep<b2s::CapsS, 0, 32>(w);
    else if (small_u) pipe_substep<b2s::CapsL, 0, 16>(w);
    else pipe_substep<b2s::CapsL, 0, 32>(w);
  }
  if (fetch_mask) {
    for (int e = 0; e < w->M.n_envs; e++) {
      if (w->caps == 0) b2s::fetch_env<b2s::CapsS>(w->M, w->S, e, fetch_mask);
      else b2s::fetch_env<b2s::CapsL>(w->M, w->S, e, fetch_mask);
    }
  }
}
void emu_apply(void* h, unsigned mask) {
  World* w = (World*)h;
  for (int e = 0; e < w->M.n_envs; e++) b2s::apply_env(w->M, w->S, e, mask);
}
void emu_fetch(void* h, unsigned mask) {
  World* w = (World*)h;
  for (int e = 0; e < w->M.n_envs; e++) {
    if (w->caps == 0) b2s::fetch_env<b2s::CapsS>(w->M, w->S, e, mask);
    else b2s::fetch_env<b2s::CapsL>(w->M, w->S, e, mask);
  }
}
// raw pointers to the exposed (AoS) buffers, host memory here
float* emu_buffer(void* h, int which) {
  World* w = (World*)h;
  switch (which) {
    case 0: return w->S.body_data;
    case 1: return w->S.xq;
    case 2: return w->S.xqd;
    case 3: return w->S.xqacc;
    case 4: return w->S.xqf;
    case 5: return w->S.xtq;
    case 6: return w->S.xtqd;
    case 7: return w->S.man;
  }
  return nullptr;
}
int* emu_man_count(void* h) { return ((World*)h)->S.man_count; }
int emu_overflow(void* h) { return *((World*)h)->S.overflow; }
}
extern "C" int* emu_nrow(void* h) { return ((World*)h)->S.sol_nrow; }
// stand-alone narrowphase probe (device collide_pair compiled for the host), same arguments as the oracle's b2o_collide
extern "C" int emu_collide(int ta, const double* pa, const double* sa, const float* va, int nva, int tb, const double* pb, const double* sb,
                           const float* vb, int nvb, double margin, double* out) {
  using namespace b2s;
  WShape A, B;
  A.type = ta; A.X.p = mk3((float)pa[0], (float)pa[1], (float)pa[2]); A.X.q = qnormalized(mkq((float)pa[3], (float)pa[4], (float)pa[5], (float)pa[6]));
  A.R = qmat(A.X.q); A.size = mk3((float)sa[0], (float)sa[1], (float)sa[2]); A.verts = va; A.nverts = nva;
  B.type = tb; B.X.p = mk3((float)pb[0], (float)pb[1], (float)pb[2]); B.X.q = qnormalized(mkq((float)pb[3], (float)pb[4], (float)pb[5], (float)pb[6]));
  B.R = qmat(B.X.q); B.size = mk3((float)sb[0], (float)sb[1], (float)sb[2]); B.verts = vb; B.nverts = nvb;
  CPoint c[4];
  int n = collide_pair(A, B, (float)margin, c);
  for (int i = 0; i < n; i++) {
    out[7 * i] = c[i].p.x; out[7 * i + 1] = c[i].p.y; out[7 * i + 2] = c[i].p.z;
    out[7 * i + 3] = c[i].n.x; out[7 * i + 4] = c[i].n.y; out[7 * i + 5] = c[i].n.z; out[7 * i + 6] = c[i].sep;
  }
  return n;
}
