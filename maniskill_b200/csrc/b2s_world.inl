// World construction shared by the CUDA library (Mem = device memory) and the host-side emulation used by the CPU
// tests (Mem = malloc).  Copies the compiled model tables, transposes per-env overrides to struct-of-arrays and
// allocates the state.  Included by b2s_api.cu and tests/emu/b2s_emu.cpp.
#pragma once
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/b200sim.h"
#include "b2s_step.cuh"

namespace b2s {

template <class Mem>
struct WorldT {
  DevModel M;
  DevState S;
  B2SModel host;  // scalar copy (pointers invalid after create)
  std::vector<void*> allocs;
  int caps;  // 0 small, 1 large

  template <class T>
  const T* up(const T* src, size_t n) {
    if (n == 0) n = 1, src = nullptr;
    void* d = Mem::alloc(n * sizeof(T));
    if (!d) return nullptr;
    if (src) Mem::upload(d, src, n * sizeof(T));
    else Mem::zero(d, n * sizeof(T));
    allocs.push_back(d);
    return (const T*)d;
  }
  template <class T>
  T* zeros(size_t n) {
    if (n == 0) n = 1;
    void* d = Mem::alloc(n * sizeof(T));
    if (!d) return nullptr;
    Mem::zero(d, n * sizeof(T));
    allocs.push_back(d);
    return (T*)d;
  }
  // env-major [N][per] host array -> SoA [per][N]
  const float* up_soa(const float* src, size_t per, size_t N) {
    std::vector<float> t(per * N ? per * N : 1, 0.f);
    for (size_t e = 0; e < N; e++)
      for (size_t k = 0; k < per; k++) t[k * N + e] = src[e * per + k];
    return up(t.data(), per * N);
  }

  const char* build(const B2SModel& m) {
    host = m;
    const size_t N = m.n_envs;
    if (m.n_dof > 32) return "n_dof > 32";
    memset(&M, 0, sizeof(M));
    memset(&S, 0, sizeof(S));
    M.n_envs = m.n_envs; M.n_art = m.n_art; M.n_dof = m.n_dof; M.n_link = m.n_link; M.n_fb = m.n_fb; M.n_shape = m.n_shape;
    M.n_pair = m.n_pair; M.n_hull = m.n_hull; M.n_eq = m.n_eq; M.n_ov_shape = m.n_ov_shape; M.n_ov_fb = m.n_ov_fb;
    M.max_contacts = m.max_contacts; M.max_manifolds = m.max_manifolds; M.n_pos_iters = m.n_pos_iters;
    M.n_vel_iters = m.n_vel_iters; M.max_dof_per_art = m.max_dof_per_art; M.n_rows = m.n_link + m.n_fb;
    M.dt = m.dt; M.gx = m.gravity[0]; M.gy = m.gravity[1]; M.gz = m.gravity[2]; M.contact_offset = m.contact_offset;
    M.rest_offset = m.rest_offset; M.max_depen_vel = m.max_depen_vel; M.contact_hertz = m.contact_hertz;
    M.contact_zeta = m.contact_zeta; M.margin_min = m.margin_min;
    const int nd = m.n_dof;
    M.dof_parent = up(m.dof_parent, nd); M.dof_art = up(m.dof_art, nd); M.dof_type = up(m.dof_type, nd);
    M.dof_T0 = up(m.dof_T0, nd * 7); M.dof_axis = up(m.dof_axis, nd * 3); M.dof_mass = up(m.dof_mass, nd);
    M.dof_com = up(m.dof_com, nd * 3); M.dof_inertia = up(m.dof_inertia, nd * 6); M.dof_gravity = up(m.dof_gravity, nd);
    M.dof_limit = up(m.dof_limit, nd * 2); M.dof_drive = up(m.dof_drive, nd * 4); M.dof_passive = up(m.dof_passive, nd * 4);
    M.dof_anc_mask = up(m.dof_anc_mask, nd);
    M.link_dof = up(m.link_dof, m.n_link); M.link_offset = up(m.link_offset, m.n_link * 7);
    M.art_dof_start = up(m.art_dof_start, m.n_art + 1); M.art_link_start = up(m.art_link_start, m.n_art + 1);
    M.eq_dof = up(m.eq_dof, m.n_eq * 2); M.eq_param = up(m.eq_param, m.n_eq * 4);
    M.fb_type = up(m.fb_type, m.n_fb); M.fb_mass = up(m.fb_mass, m.n_fb); M.fb_com = up(m.fb_com, m.n_fb * 3);
    M.fb_inertia = up(m.fb_inertia, m.n_fb * 6); M.fb_damping = up(m.fb_damping, m.n_fb * 2);
    M.fb_gravity = up(m.fb_gravity, m.n_fb); M.fb_ov = up(m.fb_ov, m.n_fb);
    {
      std::vector<int> slot(m.n_fb > 0 ? m.n_fb : 1, -1);
      int nu = nd;
      for (int b = 0; b < m.n_fb; b++)
        if (m.fb_type[b] == 0) { slot[b] = nu; nu += 6; }
      M.fb_slot = up(slot.data(), m.n_fb);
      M.n_u = nu;
    }
    M.shape_type = up(m.shape_type, m.n_shape); M.shape_owner_kind = up(m.shape_owner_kind, m.n_shape);
    M.shape_owner = up(m.shape_owner, m.n_shape); M.shape_row = up(m.shape_row, m.n_shape);
    M.shape_hull = up(m.shape_hull, m.n_shape); M.shape_ov = up(m.shape_ov, m.n_shape);
    M.shape_pose = up(m.shape_pose, m.n_shape * 7); M.shape_size = up(m.shape_size, m.n_shape * 3);
    M.shape_mu = up(m.shape_mu, m.n_shape); M.shape_bound = up(m.shape_bound, m.n_shape * 4);
    M.shape_patch = up(m.shape_patch, m.n_shape);
    M.hull_offset = up(m.hull_offset, m.n_hull + 1); M.hull_verts = up(m.hull_verts, (size_t)m.n_hull_verts * 3);
    M.hull_aabb = up(m.hull_aabb, (size_t)m.n_hull * 6);
    M.pair_a = up(m.pair_a, m.n_pair); M.pair_b = up(m.pair_b, m.n_pair);
    M.ov_shape_size = up_soa(m.ov_shape_size, (size_t)m.n_ov_shape * 3, N);
    M.ov_shape_pose = up_soa(m.ov_shape_pose, (size_t)m.n_ov_shape * 7, N);
    M.ov_shape_bound = up_soa(m.ov_shape_bound, (size_t)m.n_ov_shape * 4, N);
    M.ov_fb_mass = up_soa(m.ov_fb_mass, (size_t)m.n_ov_fb * 10, N);
    // state
    S.q = zeros<float>(nd * N); S.qd = zeros<float>(nd * N); S.tq = zeros<float>(nd * N); S.tqd = zeros<float>(nd * N);
    S.qf = zeros<float>(nd * N); S.qacc = zeros<float>(nd * N);
    {
      std::vector<float> t((size_t)m.n_art * 7 * N + 1);
      for (int a = 0; a < m.n_art; a++)
        for (int k = 0; k < 7; k++)
          for (size_t e = 0; e < N; e++) t[(size_t)(a * 7 + k) * N + e] = m.art_root_pose[a * 7 + k];
      S.root = (float*)up(t.data(), (size_t)m.n_art * 7 * N);
    }
    {
      std::vector<float> t((size_t)m.n_fb * 13 * N + 1, 0.f);
      for (int b = 0; b < m.n_fb; b++)
        for (int k = 0; k < 7; k++)
          for (size_t e = 0; e < N; e++) t[(size_t)(b * 13 + k) * N + e] = m.fb_init_pose[b * 7 + k];
      S.fb = (float*)up(t.data(), (size_t)m.n_fb * 13 * N);
    }
    S.man = zeros<float>((size_t)Caps<12, 4, 32, 1>::MAXMAN * 8 * N);
    S.man_count = zeros<int>(N);
    S.sol_rows = zeros<float>(N * (size_t)Caps<12, 4, 32, 1>::MAXROW * (M.n_u <= 16 ? 44 : 76));
    S.sol_nrow = zeros<int>(N);
    S.sol_qdd = zeros<float>((size_t)nd * N);
    S.kin_link = zeros<float>((size_t)nd * 19 * N);
    S.kin_minv = zeros<float>((size_t)nd * nd * N);
    S.kin_fb = zeros<float>((size_t)m.n_fb * 13 * N);
    S.col_mask = zeros<unsigned>((size_t)((m.n_pair + 31) / 32) * N);
    S.col_data = zeros<float>((size_t)m.n_pair * 20 * N);
    S.row_desc = zeros<float>(N * (size_t)Caps<12, 4, 32, 1>::MAXROW * 16);
    S.overflow = zeros<int>(1);
    S.body_data = zeros<float>(N * M.n_rows * 13);
    size_t nq = N * m.n_art * (m.max_dof_per_art > 0 ? m.max_dof_per_art : 1);
    S.xq = zeros<float>(nq); S.xqd = zeros<float>(nq); S.xqacc = zeros<float>(nq); S.xqf = zeros<float>(nq);
    S.xtq = zeros<float>(nq); S.xtqd = zeros<float>(nq);
    for (void* p : allocs)
      if (!p) return "allocation failed";
    caps = (nd <= 12 && m.n_fb <= 4 && m.n_shape <= 32 && m.n_art <= 1) ? 0 : 1;
    if (nd > 24 || m.n_fb > 6 || m.n_shape > 56 || m.n_art > 2) return "model exceeds compiled capacities (24 dof, 6 bodies, 56 shapes, 2 articulations)";
    return nullptr;
  }
  void release() {
    for (void* p : allocs)
      if (p) Mem::release(p);
    allocs.clear();
  }
};

typedef Caps<12, 4, 32, 1> CapsS;
typedef Caps<24, 6, 56, 2> CapsL;

}  // namespace b2s
