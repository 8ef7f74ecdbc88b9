// b200sim fused substep (device): one lane advances one sub-scene.  This is the B200 replacement for
// `PhysxGpuSystem.step()` (reference call site mani_skill/envs/scene.py:379-380 inside the 5x loop at
// mani_skill/envs/sapien_env.py:1123-1128) plus, when requested, the `gpu_fetch_*` copies that follow it
// (mani_skill/envs/scene.py:968-986):
//
//   joint-space implicit PD drive  ->  ABA (Featherstone) with cached factorisation  ->  M~^-1 columns
//   bounding-sphere broadphase over the baked candidate list  ->  narrowphase (b2s_collide.cuh)
//   constraint rows (tendon couplings, joint limits, contact normals, patch friction, torsion)
//   sub-stepped temporal Gauss-Seidel with soft penetration recovery  ->  integrate  ->  export
//
// State lives in env-major struct-of-arrays ([slot][n_envs]) so the 32 lanes of a warp touch 32 consecutive
// floats per slot; the model tables are shared by every lane.  Everything a substep needs is kept in per-lane
// scratch for all `substeps` iterations, so HBM sees each state word once in and once out per launch.
#pragma once
#include "b2s_collide.cuh"

// loops with large bodies stay rolled: the kernel is instruction-fetch sensitive (one warp per SM streams the whole
// substep through the instruction caches), the small inner loops (dot products over the dofs) are the ones worth unrolling
#if defined(__CUDACC__)
#define B2S_NO_UNROLL _Pragma("unroll 1")
#else
#define B2S_NO_UNROLL
#endif

namespace b2s {

struct DevModel {
  int n_envs, n_art, n_dof, n_link, n_fb, n_shape, n_pair, n_hull, n_eq, n_ov_shape, n_ov_fb;
  int max_contacts, max_manifolds, n_pos_iters, n_vel_iters, max_dof_per_art, n_rows;
  float dt, gx, gy, gz, contact_offset, rest_offset, max_depen_vel, contact_hertz, contact_zeta, margin_min;
  const int *dof_parent, *dof_art, *dof_type;
  const float *dof_T0, *dof_axis, *dof_mass, *dof_com, *dof_inertia, *dof_gravity, *dof_limit, *dof_drive, *dof_passive;
  const unsigned* dof_anc_mask;
  const int* link_dof;
  const float* link_offset;
  const int *art_dof_start, *art_link_start;
  const int* eq_dof;
  const float* eq_param;
  const int* fb_type;
  const float *fb_mass, *fb_com, *fb_inertia, *fb_damping, *fb_gravity;
  const int* fb_ov;
  const int* fb_slot;  // first slot of the body in the unified velocity vector of the split solve, -1 for kinematic bodies
  int n_u;             // n_dof + 6 * (dynamic free bodies)
  const int *shape_type, *shape_owner_kind, *shape_owner, *shape_row, *shape_hull, *shape_ov;
  const float *shape_pose, *shape_size, *shape_mu, *shape_bound, *shape_patch;
  const int* hull_offset;
  const float* hull_verts;
  const float* hull_aabb;
  const int *pair_a, *pair_b;
  // per-env overrides, SoA [slot][n_envs]
  const float *ov_shape_size, *ov_shape_pose, *ov_shape_bound, *ov_fb_mass;
};

struct DevState {
  // internal state, SoA [slot][n_envs]
  float *q, *qd, *tq, *tqd, *qf, *qacc;  // [n_dof]
  float* root;                           // [n_art*7]
  float* fb;                             // [n_fb*13]  pose7, linvel(com)3, angvel3
  float* man;                            // [max_manifolds*8] rowA rowB ix iy iz npts sep pad  (last substep)
  int* man_count;                        // [n_envs]
  int* overflow;                         // [1]
  // exposed env-major AoS buffers (the reference's px.cuda_* layout)
  float* body_data;                      // [n_envs, n_rows, 13]
  float *xq, *xqd, *xqacc, *xqf, *xtq, *xtqd;  // [n_envs*n_art, max_dof_per_art]
  // pipelined substep exchange, phase A -> phase B (b2s_pipe.cuh -> b2s_solve.cuh)
  float* sol_rows;                       // [n_envs, MAXROW, 44 | 68] unified constraint rows (AoS)
  int* sol_nrow;                         // [n_envs]
  float* sol_qdd;                        // [n_dof][n_envs] free joint accelerations
  // pipelined substep exchange (b2s_pipe.cuh), SoA [slot][n_envs] unless noted
  float* kin_link;                       // [n_dof*19]  joint frame pose, spatial velocity, motion axis
  float* kin_minv;                       // [n_dof*n_dof]  M~^-1
  float* kin_fb;                         // [n_fb*13]  world com, 1/m, world inverse inertia
  unsigned* col_mask;                    // [ceil(n_pair/32)]  hit bitmap over the candidate pairs
  float* col_data;                       // [n_pair*20]  point count, normal, 4 x (point, separation); valid where the hit bit is set
  float* row_desc;                       // [n_envs, MAXROW, 16]  row descriptors (AoS)
};

enum { OWNER_STATIC = 0, OWNER_LINK = 1, OWNER_BODY = 2 };
// Reasons recorded in the sticky overflow word (include/b200sim.h B2S_OVF_*): which fixed capacity dropped a contact / row.  The
// reference's PhysX reports its buffer-capacity overflows the same way: loudly, with the knob to raise
// (mani_skill/utils/structs/types.py:16-32, GPUMemoryConfig).
enum { OVF_MANIFOLDS = 1, OVF_CONTACTS = 2, OVF_ROWS = 4, OVF_ART_ROWS = 8, OVF_LIMITS = 16, OVF_EQ = 32 };
B2S_HD void raise_overflow(int* word, int code) {
#if defined(__CUDA_ARCH__)
  if ((*(volatile int*)word & code) != code) atomicOr(word, code);
#else
  *word |= code;
#endif
}
enum { ROW_CONTACT_N = 0, ROW_FRICTION = 1, ROW_LIMIT = 2, ROW_EQ = 3, ROW_PATCH_START = 0x10 /* flag: first normal row of a patch */ };
enum {
  BUF_RIGID = 1u << 0, BUF_ROOT_POSE = 1u << 1, BUF_QPOS = 1u << 2, BUF_QVEL = 1u << 3, BUF_QF = 1u << 4,
  BUF_TARGET_QPOS = 1u << 5, BUF_TARGET_QVEL = 1u << 6, BUF_QACC = 1u << 7, BUF_LINK = 1u << 8
};

template <int MAXD_, int MAXFB_, int MAXSH_, int MAXART_>
struct Caps {
  static constexpr int MAXD = MAXD_, MAXFB = MAXFB_, MAXSH = MAXSH_, MAXART = MAXART_;
  static constexpr int MAXMAN = 24, MAXCP = 64, MAXLIM = 24, MAXEQ = 2;  // MAXLIM: one active limit per joint
  static constexpr int MAXROW = MAXEQ + MAXLIM + MAXCP + 3 * MAXMAN;
  static constexpr int MAXAR = MAXROW;  // rows that touch an articulation (no separate limit)
};

template <class C>
struct Lane {
  float q[C::MAXD], qd[C::MAXD], tq[C::MAXD], tqd[C::MAXD], qf[C::MAXD], qacc[C::MAXD];
  pose root[C::MAXART];
  pose fbX[C::MAXFB];
  v3 fbv[C::MAXFB], fbw[C::MAXFB];
  // last-substep manifold summary
  int n_man;
  int man_rowA[C::MAXMAN], man_rowB[C::MAXMAN], man_npts[C::MAXMAN];
  v3 man_imp[C::MAXMAN];
  float man_sep[C::MAXMAN];
  // FK cache (valid after fk())
  pose X[C::MAXD];
  v6 V[C::MAXD];
};

B2S_HD v6 m6mul(const float* A, v6 x) {
  float xv[6] = {x.a.x, x.a.y, x.a.z, x.l.x, x.l.y, x.l.z};
  float y[6];
  for (int i = 0; i < 6; i++) {
    float s = 0.f;
    for (int j = 0; j < 6; j++) s += A[6 * i + j] * xv[j];
    y[i] = s;
  }
  return mk6(mk3(y[0], y[1], y[2]), mk3(y[3], y[4], y[5]));
}

B2S_HD void spatial_inertia(float* I, float m, v3 c, const m3& Ic) {
  float cc = dot(c, c);
  float cv[3] = {c.x, c.y, c.z};
  for (int i = 0; i < 36; i++) I[i] = 0.f;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) I[6 * i + j] = Ic.m[3 * i + j] + m * ((i == j ? cc : 0.f) - cv[i] * cv[j]);
  float cx[3][3] = {{0, -c.z, c.y}, {c.z, 0, -c.x}, {-c.y, c.x, 0}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      I[6 * i + 3 + j] = m * cx[i][j];
      I[6 * (3 + i) + j] = m * cx[j][i];
    }
  for (int i = 0; i < 3; i++) I[6 * (3 + i) + 3 + i] = m;
}

// forward kinematics only (used by fetch / update_kinematics)
template <class C>
B2S_HDN void fk(const DevModel& M, Lane<C>& L) {
  for (int i = 0; i < M.n_dof; i++) {
    int p = M.dof_parent[i], a = M.dof_art[i];
    pose Xp = p >= 0 ? L.X[p] : L.root[a];
    pose Xj = pmul(Xp, pose7(M.dof_T0 + 7 * i));
    v3 ax = mk3(M.dof_axis[3 * i], M.dof_axis[3 * i + 1], M.dof_axis[3 * i + 2]);
    pose mo = pose_ident();
    bool rev = M.dof_type[i] == 0;
    if (rev) mo.q = qaxis_angle(ax, L.q[i]);
    else mo.p = ax * L.q[i];
    L.X[i] = pmul(Xj, mo);
    L.X[i].q = qnormalized(L.X[i].q);
    v3 aw = qrot(Xj.q, ax);
    v3 Oa = L.root[a].p;
    v6 S = rev ? mk6(aw, cross(L.X[i].p - Oa, aw)) : mk6(mk3(0, 0, 0), aw);
    v6 Vp = p >= 0 ? L.V[p] : zero6();
    L.V[i] = Vp + S * L.qd[i];
  }
}

// ND > 0 fixes the dof count at compile time (loops over joints unroll, the velocity vectors live in registers);
// ND == 0 reads it from the model.
template <class C, int ND>
B2S_HDN void substep(const DevModel& M, Lane<C>& L, int env, int* overflow) {
  const int N = M.n_envs;
  const int nd = ND > 0 ? ND : M.n_dof;
  const float dt = M.dt;
  const v3 grav = mk3(M.gx, M.gy, M.gz);
  // ---------------------------------------------------------------- 1. FK + spatial quantities + drive
  v6 S[C::MAXD], cvp[C::MAXD], pA[C::MAXD], U[C::MAXD];
  float IA[C::MAXD][36];
  float Dinv[C::MAXD], u[C::MAXD], tau[C::MAXD], arm[C::MAXD];
  v3 cW[C::MAXD];
  m3 IwW[C::MAXD];
  v6 fextW[C::MAXD];
  pose* X = L.X;
  v6* V = L.V;
  B2S_NO_UNROLL
  for (int i = 0; i < nd; i++) {
    int p = M.dof_parent[i], a = M.dof_art[i];
    pose Xp = p >= 0 ? X[p] : L.root[a];
    pose Xj = pmul(Xp, pose7(M.dof_T0 + 7 * i));
    v3 ax = mk3(M.dof_axis[3 * i], M.dof_axis[3 * i + 1], M.dof_axis[3 * i + 2]);
    pose mo = pose_ident();
    bool rev = M.dof_type[i] == 0;
    if (rev) mo.q = qaxis_angle(ax, L.q[i]);
    else mo.p = ax * L.q[i];
    X[i] = pmul(Xj, mo);
    X[i].q = qnormalized(X[i].q);
    v3 aw = qrot(Xj.q, ax);
    v3 Oa = L.root[a].p;
    S[i] = rev ? mk6(aw, cross(X[i].p - Oa, aw)) : mk6(mk3(0, 0, 0), aw);
    v6 Vp = p >= 0 ? V[p] : zero6();
    v6 vj = S[i] * L.qd[i];
    V[i] = Vp + vj;
    cvp[i] = crm(V[i], vj);
    m3 Rm = qmat(X[i].q);
    v3 com = mk3(M.dof_com[3 * i], M.dof_com[3 * i + 1], M.dof_com[3 * i + 2]);
    v3 c = X[i].p + mul(Rm, com) - Oa;
    const float* in6 = M.dof_inertia + 6 * i;
    m3 Iw = mul(mul(Rm, sym6(in6[0], in6[1], in6[2], in6[3], in6[4], in6[5])), transpose(Rm));
    float mass = M.dof_mass[i];
    cW[i] = c;
    IwW[i] = Iw;
    v3 fg = grav * (mass * M.dof_gravity[i]);
    fextW[i] = mk6(cross(c, fg), fg);
    float kp = M.dof_drive[4 * i], kd = M.dof_drive[4 * i + 1];
    float damp = M.dof_passive[4 * i], armature = M.dof_passive[4 * i + 2];
    // pass 0 of the drive: every drive implicit (see the ABA loop below for the force-limited pass)
    tau[i] = kp * (L.tq[i] - L.q[i] - dt * L.qd[i]) + kd * (L.tqd[i] - L.qd[i]) + L.qf[i] - damp * L.qd[i];
    arm[i] = armature + dt * kd + dt * dt * kp + dt * damp;
  }
  // ---------------------------------------------------------------- 2. collision detection
  const int ns = M.n_shape;
  v3 bc[C::MAXSH];
  float br[C::MAXSH], bv[C::MAXSH];
  pose SX[C::MAXSH];
  v3 Ssize[C::MAXSH];
  B2S_NO_UNROLL
  for (int s = 0; s < ns; s++) {
    int kind = M.shape_owner_kind[s], ow = M.shape_owner[s];
    pose own;
    if (kind == OWNER_STATIC) own = pose_ident();
    else if (kind == OWNER_LINK) own = ow >= 0 ? X[ow] : L.root[-ow - 1];
    else own = L.fbX[ow];
    int ov = M.shape_ov[s];
    float lp[7], bd[4];
    if (ov >= 0) {
      for (int k = 0; k < 7; k++) lp[k] = M.ov_shape_pose[(size_t)(ov * 7 + k) * N + env];
      for (int k = 0; k < 4; k++) bd[k] = M.ov_shape_bound[(size_t)(ov * 4 + k) * N + env];
      Ssize[s] = mk3(M.ov_shape_size[(size_t)(ov * 3) * N + env], M.ov_shape_size[(size_t)(ov * 3 + 1) * N + env],
                     M.ov_shape_size[(size_t)(ov * 3 + 2) * N + env]);
    } else {
      for (int k = 0; k < 7; k++) lp[k] = M.shape_pose[7 * s + k];
      for (int k = 0; k < 4; k++) bd[k] = M.shape_bound[4 * s + k];
      Ssize[s] = mk3(M.shape_size[3 * s], M.shape_size[3 * s + 1], M.shape_size[3 * s + 2]);
    }
    SX[s] = pmul(own, pose7(lp));
    SX[s].q = qnormalized(SX[s].q);
    bc[s] = own.p + qrot(own.q, mk3(bd[0], bd[1], bd[2]));
    br[s] = bd[3];
    if (kind == OWNER_LINK && ow >= 0) {
      v3 w = V[ow].a;
      v3 vc = V[ow].l + cross(w, bc[s] - L.root[M.dof_art[ow]].p);
      bv[s] = norm(vc) + norm(w) * br[s];
    } else if (kind == OWNER_BODY) {
      bv[s] = norm(L.fbv[ow]) + norm(L.fbw[ow]) * (br[s] + norm(bc[s] - own.p));
    } else {
      bv[s] = 0.f;
    }
  }
  int n_man = 0, n_points = 0;
  int man_sa[C::MAXMAN], man_sb[C::MAXMAN], man_np[C::MAXMAN];
  v3 man_n[C::MAXMAN], man_p[C::MAXMAN][4];
  float man_s[C::MAXMAN][4], man_mu[C::MAXMAN], man_patch[C::MAXMAN];
  const float margin_cap = 2.f * M.contact_offset;
  const int max_man = M.max_manifolds < C::MAXMAN ? M.max_manifolds : C::MAXMAN;
  const int max_cp = M.max_contacts < C::MAXCP ? M.max_contacts : C::MAXCP;
  B2S_NO_UNROLL
  for (int k = 0; k < M.n_pair; k++) {
    int a = M.pair_a[k], b = M.pair_b[k];
    const float margin = fminf(margin_cap, M.margin_min + 2.f * dt * (bv[a] + bv[b]));
    int ta = M.shape_type[a], tb = M.shape_type[b];
    if (ta == SH_PLANE || tb == SH_PLANE) {
      int pl = ta == SH_PLANE ? a : b, ot = pl == a ? b : a;
      if (M.shape_type[ot] == SH_PLANE) continue;
      m3 Rp = qmat(SX[pl].q);
      float d = dot(bc[ot] - SX[pl].p, col(Rp, 0)) - br[ot];
      if (d > margin) continue;
    } else {
      v3 dd = bc[a] - bc[b];
      float rr = br[a] + br[b] + margin;
      if (dot(dd, dd) > rr * rr) continue;
      // bounding sphere of one against the oriented box of the other (both ways): a table-sized box has a huge
      // bounding sphere, this keeps GJK for the pairs that are really close.  Conservative => results unchanged.
      bool cull = false;
      for (int w = 0; w < 2 && !cull; w++) {
        int sb = w == 0 ? b : a, ss = w == 0 ? a : b;
        int tbx = M.shape_type[sb];
        v3 cl = mk3(0, 0, 0), hl;
        if (tbx == SH_BOX) hl = Ssize[sb];
        else if (tbx == SH_SPHERE) hl = mk3(Ssize[sb].x, Ssize[sb].x, Ssize[sb].x);
        else if (tbx == SH_CAPSULE) hl = mk3(Ssize[sb].x + Ssize[sb].y, Ssize[sb].x, Ssize[sb].x);
        else {
          const float* bx = M.hull_aabb + 6 * M.shape_hull[sb];
          cl = mk3(bx[0], bx[1], bx[2]);
          hl = mk3(bx[3], bx[4], bx[5]);
        }
        q4 qi = mkq(SX[sb].q.w, -SX[sb].q.x, -SX[sb].q.y, -SX[sb].q.z);
        v3 pl = qrot(qi, bc[ss] - SX[sb].p) - cl;
        v3 ex = mk3(fmaxf(fabsf(pl.x) - hl.x, 0.f), fmaxf(fabsf(pl.y) - hl.y, 0.f), fmaxf(fabsf(pl.z) - hl.z, 0.f));
        float lim = br[ss] + margin;
        if (dot(ex, ex) > lim * lim) cull = true;
      }
      if (cull) continue;
      // oriented box against oriented box on the six face axes (a lower bound of the true distance): robot links
      // hovering a few centimetres above the table stop here instead of running GJK on 64-vertex hulls
      {
        v3 cA = mk3(0, 0, 0), hA, cB = mk3(0, 0, 0), hB;
        if (ta == SH_BOX) hA = Ssize[a];
        else if (ta == SH_SPHERE) hA = mk3(Ssize[a].x, Ssize[a].x, Ssize[a].x);
        else if (ta == SH_CAPSULE) hA = mk3(Ssize[a].x + Ssize[a].y, Ssize[a].x, Ssize[a].x);
        else { const float* bx = M.hull_aabb + 6 * M.shape_hull[a]; cA = mk3(bx[0], bx[1], bx[2]); hA = mk3(bx[3], bx[4], bx[5]); }
        if (tb == SH_BOX) hB = Ssize[b];
        else if (tb == SH_SPHERE) hB = mk3(Ssize[b].x, Ssize[b].x, Ssize[b].x);
        else if (tb == SH_CAPSULE) hB = mk3(Ssize[b].x + Ssize[b].y, Ssize[b].x, Ssize[b].x);
        else { const float* bx = M.hull_aabb + 6 * M.shape_hull[b]; cB = mk3(bx[0], bx[1], bx[2]); hB = mk3(bx[3], bx[4], bx[5]); }
        m3 RA = qmat(SX[a].q), RB = qmat(SX[b].q);
        v3 dAB = (SX[b].p + mul(RB, cB)) - (SX[a].p + mul(RA, cA));
        float hAa[3] = {hA.x, hA.y, hA.z}, hBa[3] = {hB.x, hB.y, hB.z};
        float sepmax = -1e30f;
        for (int i = 0; i < 3; i++) {
          v3 ax = col(RA, i);
          float rB = hBa[0] * fabsf(dot(ax, col(RB, 0))) + hBa[1] * fabsf(dot(ax, col(RB, 1))) + hBa[2] * fabsf(dot(ax, col(RB, 2)));
          sepmax = fmaxf(sepmax, fabsf(dot(ax, dAB)) - hAa[i] - rB);
          v3 bxs = col(RB, i);
          float rA = hAa[0] * fabsf(dot(bxs, col(RA, 0))) + hAa[1] * fabsf(dot(bxs, col(RA, 1))) + hAa[2] * fabsf(dot(bxs, col(RA, 2)));
          sepmax = fmaxf(sepmax, fabsf(dot(bxs, dAB)) - hBa[i] - rA);
        }
        if (sepmax > margin + 1e-5f) continue;
      }
    }
    WShape WA, WB;
    WA.type = ta; WA.X = SX[a]; WA.R = qmat(SX[a].q); WA.size = Ssize[a]; WA.verts = nullptr; WA.nverts = 0;
    WB.type = tb; WB.X = SX[b]; WB.R = qmat(SX[b].q); WB.size = Ssize[b]; WB.verts = nullptr; WB.nverts = 0;
    if (ta == SH_CONVEX) {
      int h = M.shape_hull[a];
      WA.verts = M.hull_verts + 3 * M.hull_offset[h];
      WA.nverts = M.hull_offset[h + 1] - M.hull_offset[h];
    }
    if (tb == SH_CONVEX) {
      int h = M.shape_hull[b];
      WB.verts = M.hull_verts + 3 * M.hull_offset[h];
      WB.nverts = M.hull_offset[h + 1] - M.hull_offset[h];
    }
    CPoint out[4];
    int n = collide_pair(WA, WB, margin, out);
    if (n == 0) continue;
    // patches of the same two bodies with (nearly) the same normal are one friction patch (4 finger boxes on the table
    // give 4 points + 3 friction rows, not 16 + 12)
    int merge = -1;
    for (int mi = 0; mi < n_man; mi++) {
      int qa = man_sa[mi], qb = man_sb[mi];
      if (M.shape_owner_kind[qa] == M.shape_owner_kind[a] && M.shape_owner[qa] == M.shape_owner[a] &&
          M.shape_owner_kind[qb] == M.shape_owner_kind[b] && M.shape_owner[qb] == M.shape_owner[b] &&
          M.shape_row[qa] == M.shape_row[a] && M.shape_row[qb] == M.shape_row[b] && dot(man_n[mi], out[0].n) > 0.995f) {
        merge = mi;
        break;
      }
    }
    if (merge >= 0) {
      v3 cp[8];
      float cs[8];
      int nc = 0;
      for (int i = 0; i < man_np[merge]; i++) { cp[nc] = man_p[merge][i]; cs[nc] = man_s[merge][i]; nc++; }
      for (int i = 0; i < n; i++) { cp[nc] = out[i].p; cs[nc] = out[i].sep - M.rest_offset; nc++; }
      int keep[4];
      int k = reduce4(nc, cp, cs, keep);
      n_points += k - man_np[merge];
      man_np[merge] = k;
      for (int i = 0; i < k; i++) { man_p[merge][i] = cp[keep[i]]; man_s[merge][i] = cs[keep[i]]; }
      man_patch[merge] = fmaxf(man_patch[merge], fmaxf(M.shape_patch[a], M.shape_patch[b]));
      continue;
    }
    if (n_man >= max_man || n_points + n > max_cp) { *overflow |= n_man >= max_man ? OVF_MANIFOLDS : OVF_CONTACTS; continue; }
    man_sa[n_man] = a; man_sb[n_man] = b; man_np[n_man] = n; man_n[n_man] = out[0].n;
    for (int i = 0; i < n; i++) { man_p[n_man][i] = out[i].p; man_s[n_man][i] = out[i].sep - M.rest_offset; }
    man_mu[n_man] = 0.5f * (M.shape_mu[a] + M.shape_mu[b]);
    man_patch[n_man] = fmaxf(M.shape_patch[a], M.shape_patch[b]);
    n_man++;
    n_points += n;
  }
  // ---------------------------------------------------------------- 3. ABA
  // Drives are implicit springs with a force limit: pass 0 treats every drive implicitly, a drive whose implicit
  // force would exceed its limit is re-run as a constant force at the limit (pass 1, rare).
  float qdd[C::MAXD];
  for (int pass = 0; pass < 2; pass++) {
    B2S_NO_UNROLL
    for (int i = 0; i < nd; i++) {
      spatial_inertia(IA[i], M.dof_mass[i], cW[i], IwW[i]);
      pA[i] = crf(V[i], m6mul(IA[i], V[i])) - fextW[i];
    }
    B2S_NO_UNROLL
    for (int i = nd - 1; i >= 0; i--) {
      U[i] = m6mul(IA[i], S[i]);
      float D = dot6(S[i], U[i]) + arm[i];
      Dinv[i] = 1.f / D;
      u[i] = tau[i] - dot6(S[i], pA[i]);
      int p = M.dof_parent[i];
      if (p >= 0) {
        float Ia[36];
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 6; c++) Ia[6 * r + c] = IA[i][6 * r + c] - get6(U[i], r) * get6(U[i], c) * Dinv[i];
        v6 pa = pA[i] + m6mul(Ia, cvp[i]) + U[i] * (u[i] * Dinv[i]);
        for (int r = 0; r < 36; r++) IA[p][r] += Ia[r];
        pA[p] = pA[p] + pa;
      }
    }
    {
      v6 acc[C::MAXD];
      B2S_NO_UNROLL
      for (int i = 0; i < nd; i++) {
        int p = M.dof_parent[i];
        v6 ap = (p >= 0 ? acc[p] : zero6()) + cvp[i];
        qdd[i] = (u[i] - dot6(U[i], ap)) * Dinv[i];
        acc[i] = ap + S[i] * qdd[i];
      }
    }
    if (pass == 1) break;
    bool any_sat = false;
    for (int i = 0; i < nd; i++) {
      float kp = M.dof_drive[4 * i], kd = M.dof_drive[4 * i + 1], fl = M.dof_drive[4 * i + 2];
      float damp = M.dof_passive[4 * i], armature = M.dof_passive[4 * i + 2];
      float qd1 = L.qd[i] + dt * qdd[i];
      float f = kp * (L.tq[i] - L.q[i] - dt * qd1) + kd * (L.tqd[i] - qd1);
      if (fabsf(f) > fl) {
        any_sat = true;
        tau[i] = (f > 0.f ? fl : -fl) + L.qf[i] - damp * L.qd[i];
        arm[i] = armature + dt * damp;
      }
    }
    if (!any_sat) break;
  }
  float Minv[C::MAXD * C::MAXD];
  for (int i = 0; i < nd * nd; i++) Minv[i] = 0.f;
  B2S_NO_UNROLL
  for (int j = 0; j < nd; j++) {
    float uu[C::MAXD];
    for (int i = 0; i < nd; i++) uu[i] = 0.f;
    uu[j] = 1.f;
    v6 carry = U[j] * Dinv[j];
    int p = M.dof_parent[j];
    while (p >= 0) {
      uu[p] = -dot6(S[p], carry);
      carry = carry + U[p] * (uu[p] * Dinv[p]);
      p = M.dof_parent[p];
    }
    v6 aa[C::MAXD];
    B2S_NO_UNROLL
    for (int i = 0; i < nd; i++) {
      if (M.dof_art[i] != M.dof_art[j]) continue;
      int pi = M.dof_parent[i];
      v6 ap = pi >= 0 ? aa[pi] : zero6();
      float qd2 = (uu[i] - dot6(U[i], ap)) * Dinv[i];
      aa[i] = ap + S[i] * qd2;
      Minv[i * nd + j] = qd2;
    }
  }
  float v[C::MAXD];
  for (int i = 0; i < nd; i++) v[i] = L.qd[i];
  const int nfb = M.n_fb;
  v3 fv[C::MAXFB], fw[C::MAXFB], fcom[C::MAXFB], fcoml[C::MAXFB];
  float finvm[C::MAXFB];
  m3 fIinv[C::MAXFB];
  for (int b = 0; b < nfb; b++) {
    int ov = M.fb_ov[b];
    float mass, in6[6];
    v3 com;
    if (ov >= 0) {
      mass = M.ov_fb_mass[(size_t)(ov * 10) * N + env];
      com = mk3(M.ov_fb_mass[(size_t)(ov * 10 + 1) * N + env], M.ov_fb_mass[(size_t)(ov * 10 + 2) * N + env],
                M.ov_fb_mass[(size_t)(ov * 10 + 3) * N + env]);
      for (int k = 0; k < 6; k++) in6[k] = M.ov_fb_mass[(size_t)(ov * 10 + 4 + k) * N + env];
    } else {
      mass = M.fb_mass[b];
      com = mk3(M.fb_com[3 * b], M.fb_com[3 * b + 1], M.fb_com[3 * b + 2]);
      for (int k = 0; k < 6; k++) in6[k] = M.fb_inertia[6 * b + k];
    }
    m3 Rm = qmat(L.fbX[b].q);
    fcoml[b] = com;
    fcom[b] = L.fbX[b].p + mul(Rm, com);
    fv[b] = L.fbv[b];
    fw[b] = L.fbw[b];
    if (M.fb_type[b] == 0) {
      finvm[b] = 1.f / mass;
      fIinv[b] = mul(mul(Rm, inverse3(sym6(in6[0], in6[1], in6[2], in6[3], in6[4], in6[5]))), transpose(Rm));
    } else {
      finvm[b] = 0.f;
      for (int k = 0; k < 9; k++) fIinv[b].m[k] = 0.f;
    }
  }
  // ---------------------------------------------------------------- 4. rows
  int n_row = 0, n_ar = 0;
  signed char r_type[C::MAXROW], r_fb0[C::MAXROW], r_fb1[C::MAXROW], r_ncount[C::MAXROW], r_man[C::MAXROW], r_angonly[C::MAXROW];
  short r_art[C::MAXROW], r_nrow[C::MAXROW];
  v3 r_dir[C::MAXROW], r_ang0[C::MAXROW], r_ang1[C::MAXROW], r_Bang0[C::MAXROW], r_Bang1[C::MAXROW];
  float r_dinv[C::MAXROW], r_gamma[C::MAXROW], r_s0[C::MAXROW], r_mu[C::MAXROW], r_lambda[C::MAXROW], r_total[C::MAXROW];
  float JB[C::MAXAR][2 * C::MAXD];
  const int npos = M.n_pos_iters;
  const float h = dt / npos;

  // finishes row `ri`: effective mass from the articulation block (if any) and the free-body sides
#define B2S_FINISH_ROW(ri)                                                                                      \
  {                                                                                                             \
    float d_ = 0.f;                                                                                             \
    if (r_art[ri] >= 0) {                                                                                       \
      float* J_ = JB[r_art[ri]];                                                                                \
      float* B_ = J_ + C::MAXD;                                                                                 \
      for (int i_ = 0; i_ < nd; i_++) {                                                                         \
        float s_ = 0.f;                                                                                         \
        for (int j_ = 0; j_ < nd; j_++) s_ += Minv[i_ * nd + j_] * J_[j_];                                      \
        B_[i_] = s_;                                                                                            \
      }                                                                                                         \
      for (int j_ = 0; j_ < nd; j_++) d_ += J_[j_] * B_[j_];                                                    \
    }                                                                                                           \
    if (r_fb0[ri] >= 0) {                                                                                       \
      int b_ = r_fb0[ri];                                                                                       \
      v3 lin_ = r_angonly[ri] ? mk3(0, 0, 0) : r_dir[ri];                                                       \
      v3 Bl_ = lin_ * finvm[b_];                                                                                \
      r_Bang0[ri] = mul(fIinv[b_], r_ang0[ri]);                                                                 \
      d_ += dot(lin_, Bl_) + dot(r_ang0[ri], r_Bang0[ri]);                                                      \
    }                                                                                                           \
    if (r_fb1[ri] >= 0) {                                                                                       \
      int b_ = r_fb1[ri];                                                                                       \
      v3 lin_ = r_angonly[ri] ? mk3(0, 0, 0) : -r_dir[ri];                                                      \
      v3 Bl_ = lin_ * finvm[b_];                                                                                \
      r_Bang1[ri] = mul(fIinv[b_], r_ang1[ri]);                                                                 \
      d_ += dot(lin_, Bl_) + dot(r_ang1[ri], r_Bang1[ri]);                                                      \
    }                                                                                                           \
    r_dinv[ri] = (d_ + r_gamma[ri]) > 1e-12f ? 1.f / (d_ + r_gamma[ri]) : 0.f;                                  \
    r_lambda[ri] = 0.f;                                                                                         \
    r_total[ri] = 0.f;                                                                                          \
  }
#define B2S_BLANK_ROW(ri, ty)                                                                                   \
  {                                                                                                             \
    r_type[ri] = ty; r_fb0[ri] = -1; r_fb1[ri] = -1; r_art[ri] = -1; r_nrow[ri] = -1; r_ncount[ri] = 0;          \
    r_man[ri] = -1; r_angonly[ri] = 0; r_gamma[ri] = 0.f; r_s0[ri] = 0.f; r_mu[ri] = 0.f;                       \
    r_dir[ri] = mk3(0, 0, 0); r_ang0[ri] = mk3(0, 0, 0); r_ang1[ri] = mk3(0, 0, 0);                              \
    r_Bang0[ri] = mk3(0, 0, 0); r_Bang1[ri] = mk3(0, 0, 0);                                                      \
  }

  for (int e = 0; e < M.n_eq && e < C::MAXEQ; e++) {
    if (n_ar >= C::MAXAR || n_row >= C::MAXROW) { *overflow |= OVF_EQ; break; }
    int ri = n_row++;
    B2S_BLANK_ROW(ri, ROW_EQ);
    int a = M.eq_dof[2 * e], b = M.eq_dof[2 * e + 1];
    float mult = M.eq_param[4 * e], off = M.eq_param[4 * e + 1], k = M.eq_param[4 * e + 2];
    r_art[ri] = (short)n_ar++;
    float* J = JB[r_art[ri]];
    for (int j = 0; j < nd; j++) J[j] = 0.f;
    J[b] = 1.f;
    J[a] = -mult;
    r_s0[ri] = L.q[b] - mult * L.q[a] - off;
    r_gamma[ri] = 1.f / (h * h * k);
    B2S_FINISH_ROW(ri);
  }
  int n_lim = 0;
  B2S_NO_UNROLL
  for (int i = 0; i < nd; i++) {
    float lo = M.dof_limit[2 * i], hi = M.dof_limit[2 * i + 1];
    const float limit_margin = 0.005f + 2.f * dt * fabsf(L.qd[i]);  // only while the limit is reachable within this step
    B2S_NO_UNROLL
    for (int side = 0; side < 2; side++) {
      bool act = side == 0 ? (lo > -1e29f && L.q[i] - lo < limit_margin) : (hi < 1e29f && hi - L.q[i] < limit_margin);
      if (!act) continue;
      if (n_lim >= C::MAXLIM || n_ar >= C::MAXAR || n_row >= C::MAXROW) { *overflow |= n_lim >= C::MAXLIM ? OVF_LIMITS : OVF_ROWS; continue; }
      n_lim++;
      int ri = n_row++;
      B2S_BLANK_ROW(ri, ROW_LIMIT);
      r_art[ri] = (short)n_ar++;
      float* J = JB[r_art[ri]];
      for (int j = 0; j < nd; j++) J[j] = 0.f;
      J[i] = side == 0 ? 1.f : -1.f;
      r_s0[ri] = side == 0 ? L.q[i] - lo : hi - L.q[i];
      B2S_FINISH_ROW(ri);
    }
  }
  L.n_man = 0;
  B2S_NO_UNROLL
  for (int mi = 0; mi < n_man; mi++) {
    v3 n = man_n[mi];
    v3 t1 = fabsf(n.x) < 0.57735f ? normalized(cross(n, mk3(1, 0, 0))) : normalized(cross(n, mk3(0, 1, 0)));
    v3 t2 = cross(n, t1);
    v3 cen = mk3(0, 0, 0);
    float minsep = 1e30f;
    int np = man_np[mi];
    for (int i = 0; i < np; i++) { cen = cen + man_p[mi][i]; minsep = fminf(minsep, man_s[mi][i]); }
    cen = cen * (1.f / np);
    float rad = 0.f;
    for (int i = 0; i < np; i++) rad += norm(man_p[mi][i] - cen);
    rad = fmaxf(rad / np, man_patch[mi]);
    int first = n_row;
    int sh[2] = {man_sa[mi], man_sb[mi]};
    bool any_art = false;
    for (int sde = 0; sde < 2; sde++)
      if (M.shape_owner_kind[sh[sde]] == OWNER_LINK && M.shape_owner[sh[sde]] >= 0) any_art = true;
    int nrows_needed = np + 2 + (rad > 0.f ? 1 : 0);
    if (n_row + nrows_needed > C::MAXROW || (any_art && n_ar + nrows_needed > C::MAXAR)) { *overflow |= n_row + nrows_needed > C::MAXROW ? OVF_ROWS : OVF_ART_ROWS; continue; }
    B2S_NO_UNROLL
    for (int k = 0; k < nrows_needed; k++) {
      int ri = n_row++;
      bool is_n = k < np;
      bool tors = k == np + 2;
      B2S_BLANK_ROW(ri, is_n ? ROW_CONTACT_N : ROW_FRICTION);
      v3 pt = is_n ? man_p[mi][k] : cen;
      v3 dir = is_n ? n : (k == np ? t1 : (k == np + 1 ? t2 : n));
      r_dir[ri] = dir;
      r_angonly[ri] = tors ? 1 : 0;
      if (any_art) {
        r_art[ri] = (short)n_ar++;
        float* J = JB[r_art[ri]];
        for (int j = 0; j < nd; j++) J[j] = 0.f;
      }
      for (int sde = 0; sde < 2; sde++) {
        float sg = sde == 0 ? 1.f : -1.f;
        int kind = M.shape_owner_kind[sh[sde]], ow = M.shape_owner[sh[sde]];
        if (kind == OWNER_LINK && ow >= 0) {
          int a = M.dof_art[ow];
          v3 rr = pt - L.root[a].p;
          v6 F = tors ? mk6(dir, mk3(0, 0, 0)) : mk6(cross(rr, dir), dir);
          unsigned mask = M.dof_anc_mask[ow];
          float* J = JB[r_art[ri]];
          for (int j = 0; j < nd; j++)
            if (mask & (1u << j)) J[j] += sg * dot6(S[j], F);
        } else if (kind == OWNER_BODY) {
          v3 rr = pt - fcom[ow];
          v3 ang = tors ? dir * sg : cross(rr, dir) * sg;
          if (sde == 0) { r_fb0[ri] = (signed char)ow; r_ang0[ri] = ang; }
          else { r_fb1[ri] = (signed char)ow; r_ang1[ri] = ang; }
        }
      }
      if (is_n) {
        r_man[ri] = (signed char)mi;
        r_s0[ri] = man_s[mi][k];
      } else {
        r_man[ri] = tors ? -1 : (signed char)mi;
        r_mu[ri] = tors ? man_mu[mi] * rad : man_mu[mi];
        r_nrow[ri] = (short)first;
        r_ncount[ri] = (signed char)np;
      }
      B2S_FINISH_ROW(ri);
    }
    int mo = L.n_man++;
    L.man_rowA[mo] = M.shape_row[man_sa[mi]];
    L.man_rowB[mo] = M.shape_row[man_sb[mi]];
    L.man_npts[mo] = np;
    L.man_sep[mo] = minsep;
    L.man_imp[mo] = mk3(0, 0, 0);
    // remember which output slot this manifold uses (mi may differ from mo if rows overflowed)
    man_np[mi] = -1 - mo;
  }
  // ---------------------------------------------------------------- 5. sub-stepped soft TGS
  float dq[C::MAXD], vfree[C::MAXD], ac[C::MAXD];
  for (int j = 0; j < nd; j++) { dq[j] = 0.f; ac[j] = 0.f; }
  v3 dx[C::MAXFB], dth[C::MAXFB], fvfree[C::MAXFB], fwfree[C::MAXFB], acv[C::MAXFB], acw[C::MAXFB];
  for (int b = 0; b < nfb; b++) { dx[b] = mk3(0, 0, 0); dth[b] = mk3(0, 0, 0); }
  const float kPi = 3.14159265358979323846f;
  const float omega = 2.f * kPi * fminf(M.contact_hertz, 0.25f / h), zeta = M.contact_zeta;
  const float sa1 = 2.f * zeta + h * omega, sa2 = h * omega * sa1, sa3 = 1.f / (1.f + sa2);
  const float soft_rate = omega / sa1, soft_mass = sa2 * sa3, soft_imp = sa3;

#define B2S_ROW_JV(ri, VQ, LV, AV, out)                                                                         \
  {                                                                                                             \
    float s_ = 0.f;                                                                                             \
    if (r_art[ri] >= 0) {                                                                                       \
      const float* J_ = JB[r_art[ri]];                                                                          \
      for (int j_ = 0; j_ < nd; j_++) s_ += J_[j_] * VQ[j_];                                                    \
    }                                                                                                           \
    if (r_fb0[ri] >= 0) {                                                                                       \
      v3 lin_ = r_angonly[ri] ? mk3(0, 0, 0) : r_dir[ri];                                                       \
      s_ += dot(lin_, LV[r_fb0[ri]]) + dot(r_ang0[ri], AV[r_fb0[ri]]);                                          \
    }                                                                                                           \
    if (r_fb1[ri] >= 0) {                                                                                       \
      v3 lin_ = r_angonly[ri] ? mk3(0, 0, 0) : -r_dir[ri];                                                      \
      s_ += dot(lin_, LV[r_fb1[ri]]) + dot(r_ang1[ri], AV[r_fb1[ri]]);                                          \
    }                                                                                                           \
    out = s_;                                                                                                   \
  }
#define B2S_ROW_APPLY(ri, dl)                                                                                   \
  {                                                                                                             \
    if (r_art[ri] >= 0) {                                                                                       \
      const float* B_ = JB[r_art[ri]] + C::MAXD;                                                                \
      for (int j_ = 0; j_ < nd; j_++) v[j_] += B_[j_] * (dl);                                                   \
    }                                                                                                           \
    if (r_fb0[ri] >= 0) {                                                                                       \
      int b_ = r_fb0[ri];                                                                                       \
      v3 lin_ = r_angonly[ri] ? mk3(0, 0, 0) : r_dir[ri];                                                       \
      fv[b_] = fv[b_] + (lin_ * finvm[b_]) * (dl);                                                              \
      fw[b_] = fw[b_] + r_Bang0[ri] * (dl);                                                                     \
    }                                                                                                           \
    if (r_fb1[ri] >= 0) {                                                                                       \
      int b_ = r_fb1[ri];                                                                                       \
      v3 lin_ = r_angonly[ri] ? mk3(0, 0, 0) : -r_dir[ri];                                                      \
      fv[b_] = fv[b_] + (lin_ * finvm[b_]) * (dl);                                                              \
      fw[b_] = fw[b_] + r_Bang1[ri] * (dl);                                                                     \
    }                                                                                                           \
  }

  for (int it = 0; it < npos + M.n_vel_iters; it++) {
    const bool relax = it >= npos;
    if (!relax) {
      for (int j = 0; j < nd; j++) v[j] += h * qdd[j];
      for (int b = 0; b < nfb; b++) {
        if (M.fb_type[b] != 0) continue;
        fv[b] = (fv[b] + grav * (h * M.fb_gravity[b])) * fmaxf(0.f, 1.f - h * M.fb_damping[2 * b]);
        fw[b] = fw[b] * fmaxf(0.f, 1.f - h * M.fb_damping[2 * b + 1]);
      }
      // warm start: re-apply the velocity change the constraints produced in the previous sub-step
      for (int j = 0; j < nd; j++) vfree[j] = v[j];
      for (int b = 0; b < nfb; b++) { fvfree[b] = fv[b]; fwfree[b] = fw[b]; }
      if (it > 0) {
        for (int j = 0; j < nd; j++) v[j] += ac[j];
        for (int b = 0; b < nfb; b++) { fv[b] = fv[b] + acv[b]; fw[b] = fw[b] + acw[b]; }
      }
    } else {
      for (int ri = 0; ri < n_row; ri++) r_total[ri] -= r_lambda[ri];
    }
    for (int ri = 0; ri < n_row; ri++) {
      // J.v (velocity error) and J.dq (position error of the row) in one pass: two independent accumulation chains
      // and one read of the row
      float jv = 0.f, sd = 0.f;
      if (r_art[ri] >= 0) {
        const float* J_ = JB[r_art[ri]];
        for (int j_ = 0; j_ < nd; j_++) {
          float Jj = J_[j_];
          jv += Jj * v[j_];
          sd += Jj * dq[j_];
        }
      }
      if (r_fb0[ri] >= 0) {
        int b_ = r_fb0[ri];
        v3 lin_ = r_angonly[ri] ? mk3(0, 0, 0) : r_dir[ri];
        jv += dot(lin_, fv[b_]) + dot(r_ang0[ri], fw[b_]);
        sd += dot(lin_, dx[b_]) + dot(r_ang0[ri], dth[b_]);
      }
      if (r_fb1[ri] >= 0) {
        int b_ = r_fb1[ri];
        v3 lin_ = r_angonly[ri] ? mk3(0, 0, 0) : -r_dir[ri];
        jv += dot(lin_, fv[b_]) + dot(r_ang1[ri], fw[b_]);
        sd += dot(lin_, dx[b_]) + dot(r_ang1[ri], dth[b_]);
      }
      float nl;
      const float lam = r_lambda[ri];
      const int ty = r_type[ri];
      if (ty == ROW_FRICTION) {
        float nsum = 0.f;
        for (int k = 0; k < r_ncount[ri]; k++) nsum += r_lambda[r_nrow[ri] + k];
        float lim = r_mu[ri] * nsum;
        nl = fmaxf(-lim, fminf(lim, lam - jv * r_dinv[ri]));
      } else if (ty == ROW_EQ) {
        float s = r_s0[ri] + sd;
        float bias = relax ? 0.f : s / h;
        nl = lam - (jv + bias + r_gamma[ri] * lam) * r_dinv[ri];
      } else {
        float s = r_s0[ri] + sd;
        float bias, ms = 1.f, is = 0.f;
        if (s > 0.f) bias = s / h;
        else if (relax) bias = 0.f;
        else { bias = fmaxf(soft_rate * s, -M.max_depen_vel); ms = soft_mass; is = soft_imp; }
        nl = fmaxf(0.f, lam - r_dinv[ri] * ms * (jv + bias) - is * lam);
      }
      float dl = nl - r_lambda[ri];
      r_lambda[ri] = nl;
      if (dl != 0.f) B2S_ROW_APPLY(ri, dl);
    }
    // joint velocity clamp (PhysX maxJointVelocity, dof_drive[4 j + 3]; 0 = none) after every sweep
    for (int j = 0; j < nd; j++) {
      const float vmax = M.dof_drive[4 * j + 3];
      if (vmax > 0.f) v[j] = fmaxf(-vmax, fminf(vmax, v[j]));
    }
    for (int ri = 0; ri < n_row; ri++) r_total[ri] += r_lambda[ri];
    if (!relax) {
      for (int j = 0; j < nd; j++) ac[j] = v[j] - vfree[j];
      for (int b = 0; b < nfb; b++) { acv[b] = fv[b] - fvfree[b]; acw[b] = fw[b] - fwfree[b]; }
      for (int j = 0; j < nd; j++) dq[j] += h * v[j];
      for (int b = 0; b < nfb; b++) {
        dx[b] = dx[b] + fv[b] * h;
        dth[b] = dth[b] + fw[b] * h;
      }
    }
  }
  // ---------------------------------------------------------------- 6. integrate + export
  for (int i = 0; i < nd; i++) {
    L.qacc[i] = (v[i] - L.qd[i]) / dt;
    L.q[i] += dq[i];
    L.qd[i] = v[i];
  }
  for (int b = 0; b < nfb; b++) {
    if (M.fb_type[b] != 0) continue;
    v3 cnew = fcom[b] + dx[b];
    q4 qn = qnormalized(qmul(qexp(dth[b]), L.fbX[b].q));
    L.fbX[b].q = qn;
    L.fbX[b].p = cnew - qrot(qn, fcoml[b]);
    L.fbv[b] = fv[b];
    L.fbw[b] = fw[b];
  }
  for (int ri = 0; ri < n_row; ri++) {
    if (r_man[ri] >= 0) {
      int mo = -1 - man_np[r_man[ri]];
      L.man_imp[mo] = L.man_imp[mo] + r_dir[ri] * r_total[ri];
    }
  }
#undef B2S_FINISH_ROW
#undef B2S_BLANK_ROW
#undef B2S_ROW_JV
#undef B2S_ROW_APPLY
}

// ------------------------------------------------------------------------------------------------ state <-> lane
template <class C>
B2S_HDN void load_lane(const DevModel& M, const DevState& St, int env, Lane<C>& L) {
  const size_t N = M.n_envs;
  for (int i = 0; i < M.n_dof; i++) {
    L.q[i] = St.q[i * N + env]; L.qd[i] = St.qd[i * N + env]; L.tq[i] = St.tq[i * N + env];
    L.tqd[i] = St.tqd[i * N + env]; L.qf[i] = St.qf[i * N + env]; L.qacc[i] = St.qacc[i * N + env];
  }
  for (int a = 0; a < M.n_art; a++) {
    float f[7];
    for (int k = 0; k < 7; k++) f[k] = St.root[(size_t)(a * 7 + k) * N + env];
    L.root[a] = pose7(f);
  }
  for (int b = 0; b < M.n_fb; b++) {
    float f[13];
    for (int k = 0; k < 13; k++) f[k] = St.fb[(size_t)(b * 13 + k) * N + env];
    L.fbX[b] = pose7(f);
    L.fbv[b] = mk3(f[7], f[8], f[9]);
    L.fbw[b] = mk3(f[10], f[11], f[12]);
  }
  L.n_man = 0;
}

template <class C>
B2S_HDN void store_lane(const DevModel& M, const DevState& St, int env, const Lane<C>& L) {
  const size_t N = M.n_envs;
  for (int i = 0; i < M.n_dof; i++) {
    St.q[i * N + env] = L.q[i]; St.qd[i * N + env] = L.qd[i]; St.qacc[i * N + env] = L.qacc[i];
  }
  for (int b = 0; b < M.n_fb; b++) {
    if (M.fb_type[b] != 0) continue;
    float f[13] = {L.fbX[b].p.x, L.fbX[b].p.y, L.fbX[b].p.z, L.fbX[b].q.w, L.fbX[b].q.x, L.fbX[b].q.y, L.fbX[b].q.z,
                   L.fbv[b].x, L.fbv[b].y, L.fbv[b].z, L.fbw[b].x, L.fbw[b].y, L.fbw[b].z};
    for (int k = 0; k < 13; k++) St.fb[(size_t)(b * 13 + k) * N + env] = f[k];
  }
  St.man_count[env] = L.n_man;
  for (int m = 0; m < L.n_man; m++) {
    float* o = St.man + (size_t)(m * 8) * N + env;
    o[0] = (float)L.man_rowA[m]; o[N] = (float)L.man_rowB[m];
    o[2 * N] = L.man_imp[m].x; o[3 * N] = L.man_imp[m].y; o[4 * N] = L.man_imp[m].z;
    o[5 * N] = (float)L.man_npts[m]; o[6 * N] = L.man_sep[m];
  }
}

// exposed env-major AoS buffers <- lane (the fused gpu_fetch_*).  Needs a valid FK cache for BUF_LINK.
// body_out: where the [n_rows][13] block of this sub-scene goes -- its place in St.body_data, or a shared-memory staging tile that the
// kernel then moves to HBM with one bulk copy (fetch_kernel in b2s_api.cu)
template <class C>
B2S_HDN void fetch_lane(const DevModel& M, const DevState& St, int env, const Lane<C>& L, unsigned mask, float* body_out = nullptr) {
  const int nr = M.n_rows;
  if (!body_out) body_out = St.body_data + (size_t)env * nr * 13;
  if (mask & BUF_LINK) {
    for (int l = 0; l < M.n_link; l++) {
      int d = M.link_dof[l];
      pose base = d >= 0 ? L.X[d] : L.root[-d - 1];
      pose P = pmul(base, pose7(M.link_offset + 7 * l));
      P.q = qnormalized(P.q);
      v3 lv = mk3(0, 0, 0), av = mk3(0, 0, 0);
      if (d >= 0) {
        int a = M.dof_art[d];
        av = L.V[d].a;
        lv = L.V[d].l + cross(av, P.p - L.root[a].p);
      }
      float* o = body_out + (size_t)l * 13;
      o[0] = P.p.x; o[1] = P.p.y; o[2] = P.p.z; o[3] = P.q.w; o[4] = P.q.x; o[5] = P.q.y; o[6] = P.q.z;
      o[7] = lv.x; o[8] = lv.y; o[9] = lv.z; o[10] = av.x; o[11] = av.y; o[12] = av.z;
    }
  }
  if (mask & BUF_RIGID) {
    for (int b = 0; b < M.n_fb; b++) {
      float* o = body_out + (size_t)(M.n_link + b) * 13;
      o[0] = L.fbX[b].p.x; o[1] = L.fbX[b].p.y; o[2] = L.fbX[b].p.z;
      o[3] = L.fbX[b].q.w; o[4] = L.fbX[b].q.x; o[5] = L.fbX[b].q.y; o[6] = L.fbX[b].q.z;
      o[7] = L.fbv[b].x; o[8] = L.fbv[b].y; o[9] = L.fbv[b].z; o[10] = L.fbw[b].x; o[11] = L.fbw[b].y; o[12] = L.fbw[b].z;
    }
  }
  const int md = M.max_dof_per_art;
  for (int a = 0; a < M.n_art; a++) {
    int d0 = M.art_dof_start[a], d1 = M.art_dof_start[a + 1];
    size_t base = ((size_t)env * M.n_art + a) * md;
    for (int i = d0; i < d1; i++) {
      if (mask & BUF_QPOS) St.xq[base + i - d0] = L.q[i];
      if (mask & BUF_QVEL) St.xqd[base + i - d0] = L.qd[i];
      if (mask & BUF_QACC) St.xqacc[base + i - d0] = L.qacc[i];
      if (mask & BUF_TARGET_QPOS) St.xtq[base + i - d0] = L.tq[i];
      if (mask & BUF_TARGET_QVEL) St.xtqd[base + i - d0] = L.tqd[i];
      if (mask & BUF_QF) St.xqf[base + i - d0] = L.qf[i];
    }
  }
}

// exposed buffers -> internal state (gpu_apply_*)
B2S_HDN inline void apply_env(const DevModel& M, const DevState& St, int env, unsigned mask) {
  const size_t N = M.n_envs;
  const int nr = M.n_rows;
  if (mask & BUF_RIGID) {
    for (int b = 0; b < M.n_fb; b++) {
      const float* o = St.body_data + ((size_t)env * nr + M.n_link + b) * 13;
      float qn = 1.f / sqrtf(o[3] * o[3] + o[4] * o[4] + o[5] * o[5] + o[6] * o[6]);
      for (int k = 0; k < 13; k++) St.fb[(size_t)(b * 13 + k) * N + env] = (k >= 3 && k < 7) ? o[k] * qn : o[k];
    }
  }
  if (mask & BUF_ROOT_POSE) {
    for (int a = 0; a < M.n_art; a++) {
      const float* o = St.body_data + ((size_t)env * nr + M.art_link_start[a]) * 13;
      float qn = 1.f / sqrtf(o[3] * o[3] + o[4] * o[4] + o[5] * o[5] + o[6] * o[6]);
      for (int k = 0; k < 7; k++) St.root[(size_t)(a * 7 + k) * N + env] = k >= 3 ? o[k] * qn : o[k];
    }
  }
  const int md = M.max_dof_per_art;
  for (int a = 0; a < M.n_art; a++) {
    int d0 = M.art_dof_start[a], d1 = M.art_dof_start[a + 1];
    size_t base = ((size_t)env * M.n_art + a) * md;
    for (int i = d0; i < d1; i++) {
      if (mask & BUF_QPOS) St.q[i * N + env] = St.xq[base + i - d0];
      if (mask & BUF_QVEL) St.qd[i * N + env] = St.xqd[base + i - d0];
      if (mask & BUF_QF) St.qf[i * N + env] = St.xqf[base + i - d0];
      if (mask & BUF_TARGET_QPOS) St.tq[i * N + env] = St.xtq[base + i - d0];
      if (mask & BUF_TARGET_QVEL) St.tqd[i * N + env] = St.xtqd[base + i - d0];
    }
  }
}

template <class C, int ND>
B2S_HDN void step_env(const DevModel& M, const DevState& St, int env, int substeps, unsigned fetch_mask) {
  Lane<C> L;
  load_lane<C>(M, St, env, L);
  int ovf = 0;
  for (int s = 0; s < substeps; s++) substep<C, ND>(M, L, env, &ovf);
  if (ovf) raise_overflow(St.overflow, ovf);
  store_lane<C>(M, St, env, L);
  if (fetch_mask) {
    fk<C>(M, L);
    fetch_lane<C>(M, St, env, L, fetch_mask);
  }
}

template <class C>
B2S_HDN void fetch_env(const DevModel& M, const DevState& St, int env, unsigned mask, float* body_out = nullptr) {
  Lane<C> L;
  load_lane<C>(M, St, env, L);
  fk<C>(M, L);
  fetch_lane<C>(M, St, env, L, mask, body_out);
}

}  // namespace b2s
