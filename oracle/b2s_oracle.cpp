// TEST INFRASTRUCTURE ONLY.  CPU oracle for the ManiSkill `px.step()` hot path.
//
// This file restates, in plain scalar C++ (array-of-structs, one env at a time, REAL = float or double), the
// physics the reference obtains from `self.px.step()` (mani_skill/envs/scene.py:379-380, called 5x per control
// step from mani_skill/envs/sapien_env.py:1123-1128) together with the buffer exchange around it
// (mani_skill/envs/scene.py:950-986).  The arithmetic lives in the third-party `sapien>=3.0.0` / NVIDIA PhysX 5
// packages (setup.py:47-49) which are NOT in /root/reference and cannot be installed here, and the reference's
// tests hold no golden vectors for this path:
//
//      >>>  PARITY UNPINNED for absolute trajectories.  <<<
//
// What IS pinned: (1) the configuration the reference fixes (TGS, 15 position + 1 velocity iteration, contact
// offset 0.02, dt = 1/100, default friction 0.3, implicit PD drives, gravity-free robot links:
// mani_skill/utils/structs/types.py:38-74, mani_skill/agents/base_agent.py:278-282,
// mani_skill/agents/robots/panda/panda.py:60-98), (2) analytic known-answer tests in tests/test_oracle_kat.py,
// and (3) the published algorithms: Featherstone's articulated-body algorithm (RBDA 2008, table 7.1), the
// temporal Gauss-Seidel solver of Macklin et al. 2019 "Small Steps in Physics Simulation", GJK/EPA.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// One substep (everything in the sub-scene frame):
//   1. forward kinematics of every articulation (joint q -> abody poses, spatial axes S_i, velocities)
//   2. shape placement + candidate-pair narrowphase -> contact points (normal from B to A, separation)
//   3. ABA with the implicit PD drive folded into the joint-space diagonal (D_i += dt*kd + dt^2*kp) -> free
//      acceleration; free bodies get gravity + damping; M~^-1 columns from unit-torque ABA solves
//   4. constraint rows: joint couplings (fixed tendons), joint limits, contacts (normal + 2 friction)
//   5. TGS: n_pos_iters sub-steps of h = dt/n_pos_iters (add h*free acceleration, warm start, one sweep with soft
//      penetration bias, advance the linearised configuration); then n_vel_iters relaxation sweeps without bias
//   6. integrate, export link rows / qacc / per-contact impulses
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/b200sim.h"
#include "b2s_oracle_collide.h"

namespace {

struct V6 {
  R v[6];
  V6() { for (int i = 0; i < 6; i++) v[i] = 0; }
  V6(V3 a, V3 b) { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = b.x; v[4] = b.y; v[5] = b.z; }
  V3 top() const { return V3(v[0], v[1], v[2]); }
  V3 bot() const { return V3(v[3], v[4], v[5]); }
};
static inline V6 operator+(const V6& a, const V6& b) { V6 c; for (int i = 0; i < 6; i++) c.v[i] = a.v[i] + b.v[i]; return c; }
static inline V6 operator-(const V6& a, const V6& b) { V6 c; for (int i = 0; i < 6; i++) c.v[i] = a.v[i] - b.v[i]; return c; }
static inline V6 operator*(const V6& a, R s) { V6 c; for (int i = 0; i < 6; i++) c.v[i] = a.v[i] * s; return c; }
static inline R dot6(const V6& a, const V6& b) { R s = 0; for (int i = 0; i < 6; i++) s += a.v[i] * b.v[i]; return s; }
// spatial cross products (motion x motion, motion x* force), Featherstone RBDA eq. 2.31 / 2.32
static inline V6 crm(const V6& v, const V6& m) { return V6(cross(v.top(), m.top()), cross(v.top(), m.bot()) + cross(v.bot(), m.top())); }
static inline V6 crf(const V6& v, const V6& f) { return V6(cross(v.top(), f.top()) + cross(v.bot(), f.bot()), cross(v.top(), f.bot())); }

struct M6 {
  R m[6][6];
  M6() { for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) m[i][j] = 0; }
};
static inline V6 operator*(const M6& A, const V6& x) {
  V6 y;
  for (int i = 0; i < 6; i++) { R s = 0; for (int j = 0; j < 6; j++) s += A.m[i][j] * x.v[j]; y.v[i] = s; }
  return y;
}

// spatial inertia about the reference point: mass m, com c (relative to the reference point), Ic about com
static M6 spatial_inertia(R m, V3 c, const M3& Ic) {
  M6 I;
  R cc = dot(c, c);
  R cv[3] = {c.x, c.y, c.z};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) I.m[i][j] = Ic.m[i][j] + m * ((i == j ? cc : 0) - cv[i] * cv[j]);
  // top-right = m [c]x ; bottom-left = m [c]x^T
  R cx[3][3] = {{0, -c.z, c.y}, {c.z, 0, -c.x}, {-c.y, c.x, 0}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      I.m[i][3 + j] = m * cx[i][j];
      I.m[3 + i][j] = m * cx[j][i];
    }
  for (int i = 0; i < 3; i++) I.m[3 + i][3 + i] = m;
  return I;
}

enum { ROW_CONTACT_N = 0, ROW_FRICTION = 1, ROW_LIMIT = 2, ROW_EQ = 3 };

struct Row {
  int type;
  R J[32], B[32];       // articulation part over the env's dof vector (n_dof <= 32)
  bool has_art;
  int fb[2];            // free body per side or -1
  V3 lin[2], ang[2];    // jacobian of each free-body side (already signed)
  V3 Blin[2], Bang[2];  // M^-1 J^T of each free-body side
  R dinv, gamma;
  R s0;
  R mu;
  int nrow, ncount;     // first normal row of the manifold and their count (friction rows)
  R lambda, total;      // impulse of the current sub-step (warm started), sum over the step
  int contact;          // contact index or -1
  V3 dir;               // world direction of the row (contacts)
};

struct ContactOut {
  int rowA, rowB;  // exposed body rows (-1 static)
  V3 p, n;
  R sep;
  int npts;
  V3 impulse;      // total impulse applied on A during the last substep
};

struct Env {
  std::vector<R> q, qd, tq, tqd, qf, qacc;
  std::vector<Pose> root;
  std::vector<Pose> fbX;
  std::vector<V3> fbv, fbw;
  std::vector<ContactOut> contacts;
  // per-env overrides
  std::vector<R> ov_size, ov_pose, ov_bound, ov_mass;
  // FK cache for fetch
  std::vector<Pose> X;
  std::vector<V6> V;
};

struct Oracle {
  B2SModel m;
  // owned copies of the model tables
  std::vector<int32_t> dof_parent, dof_art, dof_type, link_dof, art_dof_start, art_link_start, eq_dof, fb_type, fb_ov,
      shape_type, shape_owner_kind, shape_owner, shape_row, shape_hull, shape_ov, hull_offset, pair_a, pair_b;
  std::vector<uint32_t> dof_anc_mask;
  std::vector<float> dof_T0, dof_axis, dof_mass, dof_com, dof_inertia, dof_gravity, dof_limit, dof_drive, dof_passive,
      link_offset, art_root_pose, eq_param, fb_mass, fb_com, fb_inertia, fb_damping, fb_gravity, fb_init_pose, shape_pose,
      shape_size, shape_mu, shape_bound, hull_verts, shape_patch, hull_aabb;
  std::vector<Env> envs;
  int overflow;
};

template <class T>
static void cp(std::vector<T>& dst, const T* src, size_t n) {
  dst.assign(n, T());
  if (src && n) memcpy(dst.data(), src, n * sizeof(T));
}

static void fk_env(const Oracle& O, Env& E) {
  const B2SModel& m = O.m;
  E.X.resize(m.n_dof);
  E.V.resize(m.n_dof);
  for (int i = 0; i < m.n_dof; i++) {
    int p = O.dof_parent[i];
    int a = O.dof_art[i];
    Pose Xp = p >= 0 ? E.X[p] : E.root[a];
    Pose Xj = pmul(Xp, pose_from7(&O.dof_T0[7 * i]));
    V3 ax(O.dof_axis[3 * i], O.dof_axis[3 * i + 1], O.dof_axis[3 * i + 2]);
    Pose mo;
    if (O.dof_type[i] == B2S_JOINT_REVOLUTE) mo.q = qaxis_angle(ax, E.q[i]);
    else mo.p = ax * E.q[i];
    E.X[i] = pmul(Xj, mo);
    E.X[i].q = qnormalized(E.X[i].q);
    V3 aw = qrot(Xj.q, ax);
    V3 Oa = E.root[a].p;
    V6 S = O.dof_type[i] == B2S_JOINT_REVOLUTE ? V6(aw, cross(E.X[i].p - Oa, aw)) : V6(V3(), aw);
    V6 Vp = p >= 0 ? E.V[p] : V6();
    E.V[i] = Vp + S * E.qd[i];
  }
}

static void step_env(Oracle& O, Env& E) {
  const B2SModel& m = O.m;
  const int nd = m.n_dof;
  const R dt = m.dt;
  const V3 grav(m.gravity[0], m.gravity[1], m.gravity[2]);
  // ------------------------------------------------------------ 1. FK + spatial quantities
  std::vector<Pose> X(nd);
  std::vector<V6> S(nd), V(nd), cvp(nd), pA(nd), U(nd);
  std::vector<M6> IA(nd);
  std::vector<R> Dinv(nd), u(nd), tau(nd), arm(nd);
  for (int i = 0; i < nd; i++) {
    int p = O.dof_parent[i], a = O.dof_art[i];
    Pose Xp = p >= 0 ? X[p] : E.root[a];
    Pose Xj = pmul(Xp, pose_from7(&O.dof_T0[7 * i]));
    V3 ax(O.dof_axis[3 * i], O.dof_axis[3 * i + 1], O.dof_axis[3 * i + 2]);
    Pose mo;
    if (O.dof_type[i] == B2S_JOINT_REVOLUTE) mo.q = qaxis_angle(ax, E.q[i]);
    else mo.p = ax * E.q[i];
    X[i] = pmul(Xj, mo);
    X[i].q = qnormalized(X[i].q);
    V3 aw = qrot(Xj.q, ax);
    V3 Oa = E.root[a].p;
    S[i] = O.dof_type[i] == B2S_JOINT_REVOLUTE ? V6(aw, cross(X[i].p - Oa, aw)) : V6(V3(), aw);
    V6 Vp = p >= 0 ? V[p] : V6();
    V6 vj = S[i] * E.qd[i];
    V[i] = Vp + vj;
    cvp[i] = crm(V[i], vj);
    M3 Rm = qmat(X[i].q);
    V3 com(O.dof_com[3 * i], O.dof_com[3 * i + 1], O.dof_com[3 * i + 2]);
    V3 c = X[i].p + Rm * com - Oa;
    R in6[6];
    for (int k = 0; k < 6; k++) in6[k] = O.dof_inertia[6 * i + k];
    M3 Iw = Rm * sym_from6(in6) * transpose(Rm);
    R mass = O.dof_mass[i];
    IA[i] = spatial_inertia(mass, c, Iw);
    V3 fg = grav * (mass * O.dof_gravity[i]);
    V6 fext(cross(c, fg), fg);
    pA[i] = crf(V[i], IA[i] * V[i]) - fext;
    // implicit PD drive (docs/source/user_guide/concepts/controllers.md:134; articulation_joint.py:187-195)
    R kp = O.dof_drive[4 * i], kd = O.dof_drive[4 * i + 1], fl = O.dof_drive[4 * i + 2];
    R damp = O.dof_passive[4 * i], armature = O.dof_passive[4 * i + 2];
    (void)fl;
    // first pass: every drive implicit.  tau = kp (tq - q - dt qd) + kd (tqd - qd), joint-space diagonal += dt kd + dt^2 kp
    tau[i] = kp * (E.tq[i] - E.q[i] - dt * E.qd[i]) + kd * (E.tqd[i] - E.qd[i]) + E.qf[i] - damp * E.qd[i];
    arm[i] = armature + dt * kd + dt * dt * kp + dt * damp;
  }
  // ------------------------------------------------------------ 2. collision detection
  const int ns = m.n_shape;
  std::vector<WShape> W(ns);
  std::vector<V3> bc(ns);
  std::vector<R> br(ns), bv(ns);
  const int nov = m.n_ov_shape;
  for (int s = 0; s < ns; s++) {
    Pose own;
    int kind = O.shape_owner_kind[s], ow = O.shape_owner[s];
    if (kind == B2S_OWNER_STATIC) own = Pose();
    else if (kind == B2S_OWNER_LINK) own = ow >= 0 ? X[ow] : E.root[-ow - 1];
    else own = E.fbX[ow];
    int ov = O.shape_ov[s];
    float lp[7], sz[3], bd[4];
    for (int k = 0; k < 7; k++) lp[k] = ov >= 0 ? (float)E.ov_pose[ov * 7 + k] : O.shape_pose[7 * s + k];
    for (int k = 0; k < 3; k++) sz[k] = ov >= 0 ? (float)E.ov_size[ov * 3 + k] : O.shape_size[3 * s + k];
    for (int k = 0; k < 4; k++) bd[k] = ov >= 0 ? (float)E.ov_bound[ov * 4 + k] : O.shape_bound[4 * s + k];
    (void)nov;
    W[s].type = O.shape_type[s];
    W[s].X = pmul(own, pose_from7(lp));
    W[s].X.q = qnormalized(W[s].X.q);
    W[s].Rm = qmat(W[s].X.q);
    W[s].size = V3(sz[0], sz[1], sz[2]);
    W[s].verts = nullptr;
    W[s].nverts = 0;
    if (W[s].type == B2S_SHAPE_CONVEX) {
      int h = O.shape_hull[s];
      W[s].verts = &O.hull_verts[3 * O.hull_offset[h]];
      W[s].nverts = O.hull_offset[h + 1] - O.hull_offset[h];
    }
    bc[s] = own.p + qrot(own.q, V3(bd[0], bd[1], bd[2]));
    br[s] = bd[3];
    // speed bound of any point of the shape (for the speculative margin)
    if (kind == B2S_OWNER_LINK && ow >= 0) {
      V3 w = V[ow].top();
      V3 vc = V[ow].bot() + cross(w, bc[s] - E.root[O.dof_art[ow]].p);
      bv[s] = norm(vc) + norm(w) * br[s];
    } else if (kind == B2S_OWNER_BODY) {
      bv[s] = norm(E.fbv[ow]) + norm(E.fbw[ow]) * (br[s] + norm(bc[s] - own.p));
    } else {
      bv[s] = 0;
    }
  }
  // contacts of one shape pair share a normal and form a friction patch (manifold)
  struct Manifold {
    int sa, sb, npts;
    V3 n, p[4];
    R sep[4];
    R mu, patch;
  };
  std::vector<Manifold> mans;
  int n_points = 0;
  // contacts are generated up to the distance a pair can close within this step, capped by the sum of the two
  // contact offsets (types.py:44): margin = min(2*contact_offset, margin_min + 2*dt*(speed bound A + speed bound B))
  const R margin_cap = 2 * m.contact_offset;
  for (int k = 0; k < m.n_pair; k++) {
    int a = O.pair_a[k], b = O.pair_b[k];
    const R margin = std::fmin(margin_cap, m.margin_min + 2 * dt * (bv[a] + bv[b]));
    // broadphase: bounding spheres (planes: distance of the sphere to the half-space)
    if (W[a].type == B2S_SHAPE_PLANE || W[b].type == B2S_SHAPE_PLANE) {
      int pl = W[a].type == B2S_SHAPE_PLANE ? a : b, ot = pl == a ? b : a;
      if (W[ot].type == B2S_SHAPE_PLANE) continue;
      R d = dot(bc[ot] - W[pl].X.p, W[pl].Rm.col(0)) - br[ot];
      if (d > margin) continue;
    } else {
      V3 dd = bc[a] - bc[b];
      R rr = br[a] + br[b] + margin;
      if (dot(dd, dd) > rr * rr) continue;
    }
    Contact out[4];
    int n = collide_pair(W[a], W[b], margin, out);
    if (n == 0) continue;
    // Patches of the same two bodies with (nearly) the same normal are one friction patch: the four collision boxes of a
    // Panda finger lying on the table give 4 points + 3 friction rows, not 16 + 12.
    int merge = -1;
    for (size_t mi = 0; mi < mans.size(); mi++) {
      const Manifold& Q = mans[mi];
      if (O.shape_owner_kind[Q.sa] == O.shape_owner_kind[a] && O.shape_owner[Q.sa] == O.shape_owner[a] &&
          O.shape_owner_kind[Q.sb] == O.shape_owner_kind[b] && O.shape_owner[Q.sb] == O.shape_owner[b] &&
          O.shape_row[Q.sa] == O.shape_row[a] && O.shape_row[Q.sb] == O.shape_row[b] && dot(Q.n, out[0].n) > R(0.995)) {
        merge = (int)mi;
        break;
      }
    }
    if (merge >= 0) {
      Manifold& Q = mans[merge];
      V3 cp[8];
      R cs[8];
      int nc = 0;
      for (int i = 0; i < Q.npts; i++) { cp[nc] = Q.p[i]; cs[nc] = Q.sep[i]; nc++; }
      for (int i = 0; i < n; i++) { cp[nc] = out[i].p; cs[nc] = out[i].sep - m.rest_offset; nc++; }
      int keep[4];
      int k = reduce4(nc, cp, cs, keep);
      n_points += k - Q.npts;
      Q.npts = k;
      for (int i = 0; i < k; i++) { Q.p[i] = cp[keep[i]]; Q.sep[i] = cs[keep[i]]; }
      Q.patch = std::fmax(Q.patch, (R)std::fmax(O.shape_patch[a], O.shape_patch[b]));
      continue;
    }
    if ((int)mans.size() >= m.max_manifolds || n_points + n > m.max_contacts) { O.overflow = 1; continue; }
    Manifold M;
    M.sa = a; M.sb = b; M.npts = n; M.n = out[0].n;
    for (int i = 0; i < n; i++) { M.p[i] = out[i].p; M.sep[i] = out[i].sep - m.rest_offset; }
    M.mu = R(0.5) * (O.shape_mu[a] + O.shape_mu[b]);  // PhysX default combine mode: average
    M.patch = std::fmax(O.shape_patch[a], O.shape_patch[b]);
    mans.push_back(M);
    n_points += n;
  }
  // ------------------------------------------------------------ 3. ABA (Featherstone RBDA table 7.1) in sub-scene axes
  // Drives are implicit springs with a force limit (PhysxArticulationJoint.set_drive_properties(force_limit=...),
  // mani_skill/agents/controllers/pd_joint_pos.py:38-52).  Pass 0 treats every drive implicitly; a drive whose
  // implicit force would exceed its limit is re-run as a constant force at the limit (pass 1).
  std::vector<M6> IAa;
  std::vector<V6> pAa;
  std::vector<V6> acc(nd);
  std::vector<R> qdd(nd);
  for (int pass = 0; pass < 2; pass++) {
    IAa = IA;
    pAa = pA;
    for (int i = nd - 1; i >= 0; i--) {
      U[i] = IAa[i] * S[i];
      R D = dot6(S[i], U[i]) + arm[i];
      Dinv[i] = R(1) / D;
      u[i] = tau[i] - dot6(S[i], pAa[i]);
      int p = O.dof_parent[i];
      if (p >= 0) {
        M6 Ia = IAa[i];
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 6; c++) Ia.m[r][c] -= U[i].v[r] * U[i].v[c] * Dinv[i];
        V6 pa = pAa[i] + Ia * cvp[i] + U[i] * (u[i] * Dinv[i]);
        for (int r = 0; r < 6; r++) {
          for (int c = 0; c < 6; c++) IAa[p].m[r][c] += Ia.m[r][c];
          pAa[p].v[r] += pa.v[r];
        }
      }
    }
    for (int i = 0; i < nd; i++) {
      int p = O.dof_parent[i];
      V6 ap = (p >= 0 ? acc[p] : V6()) + cvp[i];
      qdd[i] = (u[i] - dot6(U[i], ap)) * Dinv[i];
      acc[i] = ap + S[i] * qdd[i];
    }
    if (pass == 1) break;
    bool any_sat = false;
    for (int i = 0; i < nd; i++) {
      R kp = O.dof_drive[4 * i], kd = O.dof_drive[4 * i + 1], fl = O.dof_drive[4 * i + 2];
      R damp = O.dof_passive[4 * i], armature = O.dof_passive[4 * i + 2];
      R qd1 = E.qd[i] + dt * qdd[i];
      R f = kp * (E.tq[i] - E.q[i] - dt * qd1) + kd * (E.tqd[i] - qd1);
      if (std::fabs(f) > fl) {
        any_sat = true;
        tau[i] = (f > 0 ? fl : -fl) + E.qf[i] - damp * E.qd[i];
        arm[i] = armature + dt * damp;
      }
    }
    if (!any_sat) break;
  }
  // M~^-1 columns: unit joint torque, zero velocity / bias
  std::vector<R> Minv(nd * nd, 0);
  for (int j = 0; j < nd; j++) {
    std::vector<R> uu(nd, 0);
    std::vector<V6> pp(nd);
    std::vector<char> has(nd, 0);
    uu[j] = 1;
    int k = j;
    V6 carry = U[j] * Dinv[j];
    int p = O.dof_parent[k];
    while (p >= 0) {
      uu[p] = -dot6(S[p], carry);
      carry = carry + U[p] * (uu[p] * Dinv[p]);
      k = p;
      p = O.dof_parent[k];
    }
    std::vector<V6> aa(nd);
    for (int i = 0; i < nd; i++) {
      if (O.dof_art[i] != O.dof_art[j]) continue;
      int pi = O.dof_parent[i];
      V6 ap = pi >= 0 ? aa[pi] : V6();
      R qd2 = (uu[i] - dot6(U[i], ap)) * Dinv[i];
      aa[i] = ap + S[i] * qd2;
      Minv[i * nd + j] = qd2;
    }
  }
  // unconstrained velocities
  std::vector<R> v(nd);
  for (int i = 0; i < nd; i++) v[i] = E.qd[i];  // h * qdd is added sub-step by sub-step
  const int nfb = m.n_fb;
  std::vector<V3> fv(nfb), fw(nfb), fcom(nfb);
  std::vector<R> finvm(nfb);
  std::vector<M3> fIinv(nfb);
  for (int b = 0; b < nfb; b++) {
    int ov = O.fb_ov[b];
    R mass = ov >= 0 ? E.ov_mass[ov * 10] : (R)O.fb_mass[b];
    V3 com = ov >= 0 ? V3(E.ov_mass[ov * 10 + 1], E.ov_mass[ov * 10 + 2], E.ov_mass[ov * 10 + 3]) : V3(O.fb_com[3 * b], O.fb_com[3 * b + 1], O.fb_com[3 * b + 2]);
    R in6[6];
    for (int k = 0; k < 6; k++) in6[k] = ov >= 0 ? E.ov_mass[ov * 10 + 4 + k] : (R)O.fb_inertia[6 * b + k];
    M3 Rm = qmat(E.fbX[b].q);
    fcom[b] = E.fbX[b].p + Rm * com;
    fv[b] = E.fbv[b];
    fw[b] = E.fbw[b];
    if (O.fb_type[b] == B2S_BODY_DYNAMIC) {
      finvm[b] = R(1) / mass;
      fIinv[b] = Rm * inverse_sym(sym_from6(in6)) * transpose(Rm);
    } else {
      finvm[b] = 0;
      fIinv[b] = M3();
    }
  }
  // ------------------------------------------------------------ 4. rows
  std::vector<Row> rows;
  rows.reserve(96);
  auto finish_row = [&](Row& r) {
    R d = 0;
    if (r.has_art) {
      for (int i = 0; i < nd; i++) r.B[i] = 0;
      for (int i = 0; i < nd; i++) {
        R s = 0;
        for (int j = 0; j < nd; j++) s += Minv[i * nd + j] * r.J[j];
        r.B[i] = s;
      }
      for (int j = 0; j < nd; j++) d += r.J[j] * r.B[j];
    }
    for (int sde = 0; sde < 2; sde++) {
      int b = r.fb[sde];
      if (b < 0) continue;
      r.Blin[sde] = r.lin[sde] * finvm[b];
      r.Bang[sde] = fIinv[b] * r.ang[sde];
      d += dot(r.lin[sde], r.Blin[sde]) + dot(r.ang[sde], r.Bang[sde]);
    }
    r.dinv = (d + r.gamma) > R(1e-12) ? R(1) / (d + r.gamma) : 0;
    r.lambda = 0;
    r.total = 0;
  };
  auto blank_row = [&](int type) {
    Row r;
    r.type = type;
    r.has_art = false;
    r.fb[0] = r.fb[1] = -1;
    r.gamma = 0;
    r.s0 = 0;
    r.mu = 0;
    r.nrow = -1;
    r.ncount = 0;
    r.contact = -1;
    return r;
  };
  const int npos = m.n_pos_iters;
  const R h = dt / npos;
  for (int e = 0; e < m.n_eq; e++) {
    Row r = blank_row(ROW_EQ);
    int a = O.eq_dof[2 * e], b = O.eq_dof[2 * e + 1];
    R mult = O.eq_param[4 * e], off = O.eq_param[4 * e + 1], k = O.eq_param[4 * e + 2];
    r.has_art = true;
    for (int j = 0; j < nd; j++) r.J[j] = 0;
    r.J[b] = 1;
    r.J[a] = -mult;
    r.s0 = E.q[b] - mult * E.q[a] - off;
    r.gamma = R(1) / (h * h * k);
    finish_row(r);
    rows.push_back(r);
  }
  for (int i = 0; i < nd; i++) {
    R lo = O.dof_limit[2 * i], hi = O.dof_limit[2 * i + 1];
    // a limit row exists only while the joint can reach the limit within this step
    const R limit_margin = R(0.005) + 2 * dt * std::fabs(E.qd[i]);
    if (lo > R(-1e29) && E.q[i] - lo < limit_margin) {
      Row r = blank_row(ROW_LIMIT);
      r.has_art = true;
      for (int j = 0; j < nd; j++) r.J[j] = 0;
      r.J[i] = 1;
      r.s0 = E.q[i] - lo;
      finish_row(r);
      rows.push_back(r);
    }
    if (hi < R(1e29) && hi - E.q[i] < limit_margin) {
      Row r = blank_row(ROW_LIMIT);
      r.has_art = true;
      for (int j = 0; j < nd; j++) r.J[j] = 0;
      r.J[i] = -1;
      r.s0 = hi - E.q[i];
      finish_row(r);
      rows.push_back(r);
    }
  }
  E.contacts.clear();
  // one Jacobian row along `dir` at point `pt` (angular_only: pure relative rotation about dir) between the owners
  auto contact_row = [&](int type, int sa, int sb, V3 pt, V3 dir, bool angular_only) {
    Row r = blank_row(type);
    r.dir = dir;
    for (int j = 0; j < nd; j++) r.J[j] = 0;
    int sh[2] = {sa, sb};
    for (int sde = 0; sde < 2; sde++) {
      R sg = sde == 0 ? R(1) : R(-1);
      int kind = O.shape_owner_kind[sh[sde]], ow = O.shape_owner[sh[sde]];
      if (kind == B2S_OWNER_LINK && ow >= 0) {
        int a = O.dof_art[ow];
        V3 rr = pt - E.root[a].p;
        V6 F = angular_only ? V6(dir, V3()) : V6(cross(rr, dir), dir);
        uint32_t mask = O.dof_anc_mask[ow];
        for (int j = 0; j < nd; j++)
          if (mask & (1u << j)) r.J[j] += sg * dot6(S[j], F);
        r.has_art = true;
      } else if (kind == B2S_OWNER_BODY) {
        r.fb[sde] = ow;
        V3 rr = pt - fcom[ow];
        r.lin[sde] = angular_only ? V3() : dir * sg;
        r.ang[sde] = angular_only ? dir * sg : cross(rr, dir) * sg;
      }
    }
    return r;
  };
  for (size_t mi = 0; mi < mans.size(); mi++) {
    const Manifold& M = mans[mi];
    V3 n = M.n;
    V3 t1 = std::fabs(n.x) < R(0.57735) ? normalized(cross(n, V3(1, 0, 0))) : normalized(cross(n, V3(0, 1, 0)));
    V3 t2 = cross(n, t1);
    V3 cen;
    R minsep = 1e30;
    for (int i = 0; i < M.npts; i++) { cen = cen + M.p[i]; minsep = std::fmin(minsep, M.sep[i]); }
    cen = cen * (R(1) / M.npts);
    R rad = 0;
    for (int i = 0; i < M.npts; i++) rad += norm(M.p[i] - cen);
    rad = std::fmax(rad / M.npts, M.patch);
    int first = (int)rows.size();
    for (int i = 0; i < M.npts; i++) {
      Row r = contact_row(ROW_CONTACT_N, M.sa, M.sb, M.p[i], n, false);
      r.contact = (int)mi;
      r.s0 = M.sep[i];
      finish_row(r);
      rows.push_back(r);
    }
    V3 td[2] = {t1, t2};
    for (int k = 0; k < 2; k++) {
      Row r = contact_row(ROW_FRICTION, M.sa, M.sb, cen, td[k], false);
      r.contact = (int)mi;
      r.mu = M.mu;
      r.nrow = first;
      r.ncount = M.npts;
      finish_row(r);
      rows.push_back(r);
    }
    if (rad > 0) {
      Row r = contact_row(ROW_FRICTION, M.sa, M.sb, cen, n, true);
      r.contact = -1;  // torsional: no net force
      r.mu = M.mu * rad;
      r.nrow = first;
      r.ncount = M.npts;
      finish_row(r);
      rows.push_back(r);
    }
    ContactOut co;
    co.rowA = O.shape_row[M.sa];
    co.rowB = O.shape_row[M.sb];
    co.p = cen;
    co.n = n;
    co.sep = minsep;
    co.npts = M.npts;
    co.impulse = V3();
    E.contacts.push_back(co);
  }
  // ------------------------------------------------------------ 5. TGS, sub-stepped with soft contacts
  // n_pos_iters sub-steps of h = dt/n_pos_iters; each: add h * free acceleration, re-apply the previous sub-step's
  // impulses (warm start), one Gauss-Seidel sweep, advance the linearised configuration.  Penetration is removed by a
  // soft constraint (natural frequency contact_hertz, damping ratio contact_zeta; Catto's "soft step" coefficients),
  // speculative contacts (s > 0) are exact.  n_vel_iters relaxation sweeps without penetration bias follow.
  std::vector<R> dq(nd, 0), vfree(nd, 0), ac(nd, 0);
  std::vector<V3> dx(nfb), dth(nfb), fvfree(nfb), fwfree(nfb), acv(nfb), acw(nfb);
  const R kPi = R(3.14159265358979323846);
  const R omega = 2 * kPi * std::fmin((R)m.contact_hertz, R(0.25) / h), zeta = m.contact_zeta;
  const R sa1 = 2 * zeta + h * omega, sa2 = h * omega * sa1, sa3 = R(1) / (R(1) + sa2);
  const R soft_rate = omega / sa1, soft_mass = sa2 * sa3, soft_imp = sa3;
  auto row_jv = [&](const Row& r, const std::vector<R>& vq, const std::vector<V3>& lv, const std::vector<V3>& av) {
    R s = 0;
    if (r.has_art)
      for (int j = 0; j < nd; j++) s += r.J[j] * vq[j];
    for (int sde = 0; sde < 2; sde++)
      if (r.fb[sde] >= 0) s += dot(r.lin[sde], lv[r.fb[sde]]) + dot(r.ang[sde], av[r.fb[sde]]);
    return s;
  };
  auto apply = [&](const Row& r, R dl) {
    if (r.has_art)
      for (int j = 0; j < nd; j++) v[j] += r.B[j] * dl;
    for (int sde = 0; sde < 2; sde++) {
      int b = r.fb[sde];
      if (b < 0) continue;
      fv[b] = fv[b] + r.Blin[sde] * dl;
      fw[b] = fw[b] + r.Bang[sde] * dl;
    }
  };
  auto sweep = [&](bool relax) {
    for (size_t ri = 0; ri < rows.size(); ri++) {
      Row& r = rows[ri];
      R jv = row_jv(r, v, fv, fw);
      R nl;
      if (r.type == ROW_FRICTION) {
        R nsum = 0;
        for (int k = 0; k < r.ncount; k++) nsum += rows[r.nrow + k].lambda;
        R lim = r.mu * nsum;
        nl = std::fmax(-lim, std::fmin(lim, r.lambda - jv * r.dinv));
      } else if (r.type == ROW_EQ) {
        R s = r.s0 + row_jv(r, dq, dx, dth);
        R bias = relax ? R(0) : s / h;
        nl = r.lambda - (jv + bias + r.gamma * r.lambda) * r.dinv;
      } else {
        R s = r.s0 + row_jv(r, dq, dx, dth);
        R bias, ms = 1, is = 0;
        if (s > 0) bias = s / h;  // speculative: the gap may close, not overshoot
        else if (relax) bias = 0;
        else { bias = std::fmax(soft_rate * s, -(R)m.max_depen_vel); ms = soft_mass; is = soft_imp; }
        nl = std::fmax(R(0), r.lambda - r.dinv * ms * (jv + bias) - is * r.lambda);
      }
      R dl = nl - r.lambda;
      r.lambda = nl;
      if (dl != 0) apply(r, dl);
    }
  };
  // PhysX clamps articulation joint velocities to PxArticulationJointReducedCoordinate::maxJointVelocity (default 100 rad/s | m/s,
  // which sapien leaves untouched); here: after every sweep, dof_drive[4 j + 3] (0 = no clamp)
  auto clamp_joint_velocity = [&]() {
    for (int j = 0; j < nd; j++) {
      const R vmax = O.dof_drive[4 * j + 3];
      if (vmax > 0) v[j] = std::fmax(-vmax, std::fmin(vmax, v[j]));
    }
  };
  for (int it = 0; it < npos; it++) {
    for (int j = 0; j < nd; j++) v[j] += h * qdd[j];
    for (int b = 0; b < nfb; b++) {
      if (O.fb_type[b] != B2S_BODY_DYNAMIC) continue;
      fv[b] = (fv[b] + grav * (h * O.fb_gravity[b])) * std::fmax(R(0), R(1) - h * O.fb_damping[2 * b]);
      fw[b] = fw[b] * std::fmax(R(0), R(1) - h * O.fb_damping[2 * b + 1]);
    }
    // warm start: re-apply the velocity change the constraints produced in the previous sub-step (= sum_r B_r lambda_r)
    vfree = v; fvfree = fv; fwfree = fw;
    if (it > 0) {
      for (int j = 0; j < nd; j++) v[j] += ac[j];
      for (int b = 0; b < nfb; b++) { fv[b] = fv[b] + acv[b]; fw[b] = fw[b] + acw[b]; }
    }
    sweep(false);
    clamp_joint_velocity();
    for (int j = 0; j < nd; j++) ac[j] = v[j] - vfree[j];
    for (int b = 0; b < nfb; b++) { acv[b] = fv[b] - fvfree[b]; acw[b] = fw[b] - fwfree[b]; }
    for (size_t ri = 0; ri < rows.size(); ri++) rows[ri].total += rows[ri].lambda;
    for (int j = 0; j < nd; j++) dq[j] += h * v[j];
    for (int b = 0; b < nfb; b++) {
      dx[b] = dx[b] + fv[b] * h;
      dth[b] = dth[b] + fw[b] * h;
    }
  }
  for (int it = 0; it < m.n_vel_iters; it++) {
    for (size_t ri = 0; ri < rows.size(); ri++) rows[ri].total -= rows[ri].lambda;
    sweep(true);
    clamp_joint_velocity();
    for (size_t ri = 0; ri < rows.size(); ri++) rows[ri].total += rows[ri].lambda;
  }
  // ------------------------------------------------------------ 6. integrate + export
  for (int i = 0; i < nd; i++) {
    E.qacc[i] = (v[i] - E.qd[i]) / dt;
    E.q[i] += dq[i];
    E.qd[i] = v[i];
  }
  for (int b = 0; b < nfb; b++) {
    if (O.fb_type[b] != B2S_BODY_DYNAMIC) continue;
    int ov = O.fb_ov[b];
    V3 com = ov >= 0 ? V3(E.ov_mass[ov * 10 + 1], E.ov_mass[ov * 10 + 2], E.ov_mass[ov * 10 + 3]) : V3(O.fb_com[3 * b], O.fb_com[3 * b + 1], O.fb_com[3 * b + 2]);
    V3 cnew = fcom[b] + dx[b];
    Q4 qn = qnormalized(qmul(qexp(dth[b]), E.fbX[b].q));
    E.fbX[b].q = qn;
    E.fbX[b].p = cnew - qrot(qn, com);
    E.fbv[b] = fv[b];
    E.fbw[b] = fw[b];
  }
  for (size_t ri = 0; ri < rows.size(); ri++) {
    const Row& r = rows[ri];
    if (r.contact >= 0) E.contacts[r.contact].impulse = E.contacts[r.contact].impulse + r.dir * r.total;
  }
}

}  // namespace

extern "C" {

void* b2o_create(const B2SModel* mp) {
  Oracle* O = new Oracle();
  const B2SModel& m = *mp;
  O->m = m;
  O->overflow = 0;
  int nd = m.n_dof;
  cp(O->dof_parent, m.dof_parent, nd); cp(O->dof_art, m.dof_art, nd); cp(O->dof_type, m.dof_type, nd);
  cp(O->dof_T0, m.dof_T0, nd * 7); cp(O->dof_axis, m.dof_axis, nd * 3); cp(O->dof_mass, m.dof_mass, nd);
  cp(O->dof_com, m.dof_com, nd * 3); cp(O->dof_inertia, m.dof_inertia, nd * 6); cp(O->dof_gravity, m.dof_gravity, nd);
  cp(O->dof_limit, m.dof_limit, nd * 2); cp(O->dof_drive, m.dof_drive, nd * 4); cp(O->dof_passive, m.dof_passive, nd * 4);
  cp(O->dof_anc_mask, m.dof_anc_mask, nd);
  cp(O->link_dof, m.link_dof, m.n_link); cp(O->link_offset, m.link_offset, m.n_link * 7);
  cp(O->art_root_pose, m.art_root_pose, m.n_art * 7); cp(O->art_dof_start, m.art_dof_start, m.n_art + 1);
  cp(O->art_link_start, m.art_link_start, m.n_art + 1);
  cp(O->eq_dof, m.eq_dof, m.n_eq * 2); cp(O->eq_param, m.eq_param, m.n_eq * 4);
  cp(O->fb_type, m.fb_type, m.n_fb); cp(O->fb_mass, m.fb_mass, m.n_fb); cp(O->fb_com, m.fb_com, m.n_fb * 3);
  cp(O->fb_inertia, m.fb_inertia, m.n_fb * 6); cp(O->fb_damping, m.fb_damping, m.n_fb * 2); cp(O->fb_gravity, m.fb_gravity, m.n_fb);
  cp(O->fb_init_pose, m.fb_init_pose, m.n_fb * 7); cp(O->fb_ov, m.fb_ov, m.n_fb);
  cp(O->shape_type, m.shape_type, m.n_shape); cp(O->shape_owner_kind, m.shape_owner_kind, m.n_shape);
  cp(O->shape_owner, m.shape_owner, m.n_shape); cp(O->shape_row, m.shape_row, m.n_shape);
  cp(O->shape_pose, m.shape_pose, m.n_shape * 7); cp(O->shape_size, m.shape_size, m.n_shape * 3);
  cp(O->shape_hull, m.shape_hull, m.n_shape); cp(O->shape_mu, m.shape_mu, m.n_shape); cp(O->shape_bound, m.shape_bound, m.n_shape * 4);
  cp(O->shape_ov, m.shape_ov, m.n_shape); cp(O->shape_patch, m.shape_patch, m.n_shape);
  cp(O->hull_offset, m.hull_offset, m.n_hull + 1); cp(O->hull_verts, m.hull_verts, m.n_hull_verts * 3); cp(O->hull_aabb, m.hull_aabb, m.n_hull * 6);
  cp(O->pair_a, m.pair_a, m.n_pair); cp(O->pair_b, m.pair_b, m.n_pair);
  O->envs.resize(m.n_envs);
  for (int e = 0; e < m.n_envs; e++) {
    Env& E = O->envs[e];
    E.q.assign(nd, 0); E.qd.assign(nd, 0); E.tq.assign(nd, 0); E.tqd.assign(nd, 0); E.qf.assign(nd, 0); E.qacc.assign(nd, 0);
    E.root.resize(m.n_art);
    for (int a = 0; a < m.n_art; a++) E.root[a] = pose_from7(&O->art_root_pose[7 * a]);
    E.fbX.resize(m.n_fb); E.fbv.assign(m.n_fb, V3()); E.fbw.assign(m.n_fb, V3());
    for (int b = 0; b < m.n_fb; b++) E.fbX[b] = pose_from7(&O->fb_init_pose[7 * b]);
    E.ov_size.resize(m.n_ov_shape * 3); E.ov_pose.resize(m.n_ov_shape * 7); E.ov_bound.resize(m.n_ov_shape * 4);
    for (int k = 0; k < m.n_ov_shape * 3; k++) E.ov_size[k] = m.ov_shape_size[(size_t)e * m.n_ov_shape * 3 + k];
    for (int k = 0; k < m.n_ov_shape * 7; k++) E.ov_pose[k] = m.ov_shape_pose[(size_t)e * m.n_ov_shape * 7 + k];
    for (int k = 0; k < m.n_ov_shape * 4; k++) E.ov_bound[k] = m.ov_shape_bound[(size_t)e * m.n_ov_shape * 4 + k];
    E.ov_mass.resize(m.n_ov_fb * 10);
    for (int k = 0; k < m.n_ov_fb * 10; k++) E.ov_mass[k] = m.ov_fb_mass[(size_t)e * m.n_ov_fb * 10 + k];
  }
  return O;
}

void b2o_destroy(void* h) { delete (Oracle*)h; }

// state exchange, all arrays env-major float64.
// which: 0 q, 1 qd, 2 target q, 3 target qd, 4 qf, 5 qacc   [n_envs, n_dof]
void b2o_set_joint(void* h, int which, const double* src) {
  Oracle* O = (Oracle*)h;
  int nd = O->m.n_dof;
  for (int e = 0; e < O->m.n_envs; e++) {
    Env& E = O->envs[e];
    std::vector<R>* a[6] = {&E.q, &E.qd, &E.tq, &E.tqd, &E.qf, &E.qacc};
    for (int i = 0; i < nd; i++) (*a[which])[i] = (R)src[(size_t)e * nd + i];
  }
}
void b2o_get_joint(void* h, int which, double* dst) {
  Oracle* O = (Oracle*)h;
  int nd = O->m.n_dof;
  for (int e = 0; e < O->m.n_envs; e++) {
    Env& E = O->envs[e];
    std::vector<R>* a[6] = {&E.q, &E.qd, &E.tq, &E.tqd, &E.qf, &E.qacc};
    for (int i = 0; i < nd; i++) dst[(size_t)e * nd + i] = (double)(*a[which])[i];
  }
}
// free bodies: [n_envs, n_fb, 13] pos quat(wxyz) linvel(com) angvel
void b2o_set_bodies(void* h, const double* src) {
  Oracle* O = (Oracle*)h;
  int nb = O->m.n_fb;
  for (int e = 0; e < O->m.n_envs; e++)
    for (int b = 0; b < nb; b++) {
      const double* s = src + ((size_t)e * nb + b) * 13;
      Env& E = O->envs[e];
      E.fbX[b].p = V3((R)s[0], (R)s[1], (R)s[2]);
      E.fbX[b].q = qnormalized(Q4((R)s[3], (R)s[4], (R)s[5], (R)s[6]));
      E.fbv[b] = V3((R)s[7], (R)s[8], (R)s[9]);
      E.fbw[b] = V3((R)s[10], (R)s[11], (R)s[12]);
    }
}
void b2o_get_bodies(void* h, double* dst) {
  Oracle* O = (Oracle*)h;
  int nb = O->m.n_fb;
  for (int e = 0; e < O->m.n_envs; e++)
    for (int b = 0; b < nb; b++) {
      double* s = dst + ((size_t)e * nb + b) * 13;
      const Env& E = O->envs[e];
      s[0] = E.fbX[b].p.x; s[1] = E.fbX[b].p.y; s[2] = E.fbX[b].p.z;
      s[3] = E.fbX[b].q.w; s[4] = E.fbX[b].q.x; s[5] = E.fbX[b].q.y; s[6] = E.fbX[b].q.z;
      s[7] = E.fbv[b].x; s[8] = E.fbv[b].y; s[9] = E.fbv[b].z;
      s[10] = E.fbw[b].x; s[11] = E.fbw[b].y; s[12] = E.fbw[b].z;
    }
}
// articulation roots [n_envs, n_art, 7]
void b2o_set_roots(void* h, const double* src) {
  Oracle* O = (Oracle*)h;
  int na = O->m.n_art;
  for (int e = 0; e < O->m.n_envs; e++)
    for (int a = 0; a < na; a++) {
      const double* s = src + ((size_t)e * na + a) * 7;
      O->envs[e].root[a].p = V3((R)s[0], (R)s[1], (R)s[2]);
      O->envs[e].root[a].q = qnormalized(Q4((R)s[3], (R)s[4], (R)s[5], (R)s[6]));
    }
}
// link rows [n_envs, n_link, 13] from forward kinematics of the current state (pose + velocity of the link origin)
void b2o_get_links(void* h, double* dst) {
  Oracle* O = (Oracle*)h;
  const B2SModel& m = O->m;
  for (int e = 0; e < m.n_envs; e++) {
    Env& E = O->envs[e];
    fk_env(*O, E);
    for (int l = 0; l < m.n_link; l++) {
      int d = O->link_dof[l];
      Pose base = d >= 0 ? E.X[d] : E.root[-d - 1];
      Pose P = pmul(base, pose_from7(&O->link_offset[7 * l]));
      P.q = qnormalized(P.q);
      V3 lv, av;
      if (d >= 0) {
        int a = O->dof_art[d];
        av = E.V[d].top();
        lv = E.V[d].bot() + cross(av, P.p - E.root[a].p);
      }
      double* s = dst + ((size_t)e * m.n_link + l) * 13;
      s[0] = P.p.x; s[1] = P.p.y; s[2] = P.p.z; s[3] = P.q.w; s[4] = P.q.x; s[5] = P.q.y; s[6] = P.q.z;
      s[7] = lv.x; s[8] = lv.y; s[9] = lv.z; s[10] = av.x; s[11] = av.y; s[12] = av.z;
    }
  }
}

void b2o_step(void* h, int substeps) {
  Oracle* O = (Oracle*)h;
  for (int s = 0; s < substeps; s++)
    for (int e = 0; e < O->m.n_envs; e++) step_env(*O, O->envs[e]);
}
// range variant for multi-threaded timing of the CPU baseline (each thread owns a slice of envs)
void b2o_step_range(void* h, int substeps, int env_begin, int env_end) {
  Oracle* O = (Oracle*)h;
  for (int s = 0; s < substeps; s++)
    for (int e = env_begin; e < env_end; e++) step_env(*O, O->envs[e]);
}

// all envs, `nthreads` std::threads each owning a contiguous slice (CPU baseline timing)
void b2o_step_mt(void* h, int substeps, int nthreads) {
  Oracle* O = (Oracle*)h;
  int n = O->m.n_envs;
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    int lo = (int)((long long)n * t / nthreads), hi = (int)((long long)n * (t + 1) / nthreads);
    if (hi > lo) th.emplace_back([=]() { for (int s = 0; s < substeps; s++) for (int e = lo; e < hi; e++) step_env(*O, O->envs[e]); });
  }
  for (auto& t : th) t.join();
}

int b2o_contact_count(void* h, int env) { return (int)((Oracle*)h)->envs[env].contacts.size(); }
// per manifold: rowA rowB px py pz nx ny nz sep ix iy iz  (12 doubles; p = patch centroid, sep = min over points)
void b2o_get_contacts(void* h, int env, double* dst) {
  const Env& E = ((Oracle*)h)->envs[env];
  for (size_t i = 0; i < E.contacts.size(); i++) {
    const ContactOut& c = E.contacts[i];
    double* s = dst + i * 12;
    s[0] = c.rowA; s[1] = c.rowB; s[2] = c.p.x; s[3] = c.p.y; s[4] = c.p.z; s[5] = c.n.x; s[6] = c.n.y; s[7] = c.n.z;
    s[8] = c.sep; s[9] = c.impulse.x; s[10] = c.impulse.y; s[11] = c.impulse.z;
  }
}
// sum of last-substep contact impulses between rows a and b acting on a  -> out[n_envs,3]
void b2o_pair_impulse(void* h, int rowA, int rowB, double* out) {
  Oracle* O = (Oracle*)h;
  for (int e = 0; e < O->m.n_envs; e++) {
    V3 s;
    for (const ContactOut& c : O->envs[e].contacts) {
      const bool any = rowB == -2;  // B2S_ANY_BODY: net impulse on rowA
      if (c.rowA == rowA && (any || c.rowB == rowB)) s = s + c.impulse;
      else if (c.rowB == rowA && (any || c.rowA == rowB)) s = s - c.impulse;
    }
    out[3 * e] = s.x; out[3 * e + 1] = s.y; out[3 * e + 2] = s.z;
  }
}
int b2o_overflow(void* h) { return ((Oracle*)h)->overflow; }
int b2o_real_size(void) { return (int)sizeof(R); }

// stand-alone narrowphase probe for unit tests: two shapes given as (type, pose7, size3, verts, nverts)
int b2o_collide(int ta, const double* pa, const double* sa, const float* va, int nva, int tb, const double* pb, const double* sb,
                const float* vb, int nvb, double margin, double* out) {
  WShape A, B;
  A.type = ta; A.X.p = V3((R)pa[0], (R)pa[1], (R)pa[2]); A.X.q = qnormalized(Q4((R)pa[3], (R)pa[4], (R)pa[5], (R)pa[6]));
  A.Rm = qmat(A.X.q); A.size = V3((R)sa[0], (R)sa[1], (R)sa[2]); A.verts = va; A.nverts = nva;
  B.type = tb; B.X.p = V3((R)pb[0], (R)pb[1], (R)pb[2]); B.X.q = qnormalized(Q4((R)pb[3], (R)pb[4], (R)pb[5], (R)pb[6]));
  B.Rm = qmat(B.X.q); B.size = V3((R)sb[0], (R)sb[1], (R)sb[2]); B.verts = vb; B.nverts = nvb;
  Contact c[4];
  int n = collide_pair(A, B, (R)margin, c);
  for (int i = 0; i < n; i++) {
    double* s = out + 7 * i;
    s[0] = c[i].p.x; s[1] = c[i].p.y; s[2] = c[i].p.z; s[3] = c[i].n.x; s[4] = c[i].n.y; s[5] = c[i].n.z; s[6] = c[i].sep;
  }
  return n;
}
}
