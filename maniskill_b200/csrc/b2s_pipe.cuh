// b200sim pipelined substep, phase A: the part of the substep before the constraint solve (steps 1-4 of the single-lane
// `substep` in b2s_step.cuh) cut into four kernels so that each part runs at its own natural width instead of one lane per
// sub-scene:
//
//   kin      one lane per sub-scene        FK, spatial axes/velocities, implicit-PD ABA, M~^-1, free-body mass properties
//   collide  one lane per (candidate pair, sub-scene)   shape placement, broadphase culls, narrowphase (b2s_collide.cuh)
//   manifest one lane per sub-scene        friction patches in candidate order (merge / capacity), 64-byte row descriptors
//   rowfill  one lane per (row, sub-scene) Jacobian, response B = M~^-1 J^T, effective mass -> unified rows of b2s_solve.cuh
//
// followed by the unchanged phase B (`solve_env`).  The arithmetic per pair / per row is the one of `substep` (same operation
// order, so the host-side emulation of both paths agrees to rounding); what changes is where the intermediate results live:
// env-major struct-of-arrays exchange buffers ([slot][n_envs], L2 resident) instead of ~27 KB of per-lane scratch.
//
// Replaces (with b2s_solve.cuh) `PhysxGpuSystem.step()` -- reference call site mani_skill/envs/scene.py:379-380.
#pragma once
#include "b2s_solve.cuh"

namespace b2s {

#define B2S_KL 19  // kin_link floats per joint: X(7) V(ang3, lin3) S(ang3, lin3)
#define B2S_KF 13  // kin_fb floats per free body: com(3) 1/m fIinv(9)
#define B2S_CD 20  // col_data floats per candidate pair: point count, normal(3), 4 x (point(3), separation)
#define B2S_RD 16  // row descriptor floats: meta sides flags pt(3) dir(3) s0 mu gamma dofA coefA dofB coefB

B2S_HD float int_as_float(int i) {
  union { int i; float f; } c;
  c.i = i;
  return c.f;
}

// 64-byte row descriptor (manifest / kin -> rowfill), written with four 16-byte stores
B2S_HD void emit_desc(float* RD, int ri, int ty, int nrow, int ncount, int slot, int sides, int flags, v3 pt, v3 dir, float s0, float mu,
                      float gamma, int dA, float cA, int dB, float cB) {
  struct alignas(16) f4 { float x, y, z, w; };
  f4* D = reinterpret_cast<f4*>(RD + (size_t)ri * B2S_RD);
  f4 a, b, c, d;
  a.x = int_as_float((ty & 0xff) | (nrow << 8) | ((ncount & 0xff) << 16) | (((slot + 1) & 0xff) << 24));
  a.y = int_as_float(sides); a.z = int_as_float(flags); a.w = pt.x;
  b.x = pt.y; b.y = pt.z; b.z = dir.x; b.w = dir.y;
  c.x = dir.z; c.y = s0; c.z = mu; c.w = gamma;
  d.x = int_as_float(dA); d.y = cA; d.z = int_as_float(dB); d.w = cB;
  D[0] = a; D[1] = b; D[2] = c; D[3] = d;
}

// The rows that do not depend on the collision pass -- tendon couplings and the joint limits that can be reached within this step --
// are emitted by the FK kernel, which holds q and qd in registers; manifest appends the contact rows after them.
template <class C>
B2S_HD void emit_joint_rows(const DevModel& M, const DevState& St, int env, const float* q, const float* qd, int nd) {
  const size_t N = M.n_envs;
  const float dt = M.dt, h = dt / M.n_pos_iters;
  float* RD = St.row_desc + (size_t)env * C::MAXROW * B2S_RD;
  const v3 zero3 = mk3(0, 0, 0);
  int n_row = 0, ovf = 0;
  for (int e = 0; e < M.n_eq && e < C::MAXEQ; e++) {
    if (n_row >= C::MAXAR || n_row >= C::MAXROW) { ovf |= OVF_EQ; break; }
    const int ri = n_row++;
    const int a = M.eq_dof[2 * e], b = M.eq_dof[2 * e + 1];
    const float mult = M.eq_param[4 * e], off = M.eq_param[4 * e + 1], kk = M.eq_param[4 * e + 2];
    emit_desc(RD, ri, ROW_EQ, 0, 0, -1, 0, 0, zero3, zero3, q[b] - mult * q[a] - off, 0.f, 1.f / (h * h * kk), b, 1.f, a, -mult);
  }
  int n_lim = 0;
  B2S_NO_UNROLL
  for (int i = 0; i < nd; i++) {
    const float lo = M.dof_limit[2 * i], hi = M.dof_limit[2 * i + 1];
    const float qi = q[i];
    const float limit_margin = 0.005f + 2.f * dt * fabsf(qd[i]);  // only while the limit is reachable within this step
    B2S_NO_UNROLL
    for (int side = 0; side < 2; side++) {
      const bool act = side == 0 ? (lo > -1e29f && qi - lo < limit_margin) : (hi < 1e29f && hi - qi < limit_margin);
      if (!act) continue;
      if (n_lim >= C::MAXLIM || n_row >= C::MAXAR || n_row >= C::MAXROW) { ovf |= n_lim >= C::MAXLIM ? OVF_LIMITS : OVF_ROWS; continue; }
      n_lim++;
      const int ri = n_row++;
      emit_desc(RD, ri, ROW_LIMIT, 0, 0, -1, 0, 0, zero3, zero3, side == 0 ? qi - lo : hi - qi, 0.f, 0.f, i, side == 0 ? 1.f : -1.f, -1, 0.f);
    }
  }
  St.sol_nrow[env] = n_row;  // manifest continues from here
  if (ovf) raise_overflow(St.overflow, ovf);
  (void)N;
}

// free body b: world centre of mass, inverse mass, world inverse inertia (zero for kinematic bodies) -> kin_fb
B2S_HDN void kin_free_body(const DevModel& M, const DevState& St, int env, int b) {
  const size_t N = M.n_envs;
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
  float f[7];
  for (int k = 0; k < 7; k++) f[k] = St.fb[(size_t)(b * 13 + k) * N + env];
  pose Xb = pose7(f);
  m3 Rm = qmat(Xb.q);
  v3 fcom = Xb.p + mul(Rm, com);
  float finvm = 0.f;
  m3 fI;
  for (int k = 0; k < 9; k++) fI.m[k] = 0.f;
  if (M.fb_type[b] == 0) {
    finvm = 1.f / mass;
    fI = mul(mul(Rm, inverse3(sym6(in6[0], in6[1], in6[2], in6[3], in6[4], in6[5]))), transpose(Rm));
  }
  float* o = St.kin_fb + (size_t)(b * B2S_KF) * N + env;
  o[0] = fcom.x; o[N] = fcom.y; o[2 * N] = fcom.z; o[3 * N] = finvm;
  for (int k = 0; k < 9; k++) o[(4 + k) * N] = fI.m[k];
}

// ------------------------------------------------------------------------------------------------ kin
// PART 0: everything.  The library runs it as two kernels so that the dynamics overlap the collision pass: PART 1 = forward
// kinematics only (joint frames, velocities, motion axes -> kin_link; clears the hit bitmap), PART 2 = the rest (reloads kin_link).
template <class C, int ND, int PART = 0>
B2S_HDN void kin_env(const DevModel& M, const DevState& St, int env) {
  const size_t N = M.n_envs;
  const int nd = ND > 0 ? ND : M.n_dof;
  const float dt = M.dt;
  const v3 grav = mk3(M.gx, M.gy, M.gz);
  float q[C::MAXD], qd[C::MAXD], tq[C::MAXD], tqd[C::MAXD], qf[C::MAXD];
  pose root[C::MAXART];
  for (int i = 0; i < nd; i++) {
    q[i] = St.q[i * N + env]; qd[i] = St.qd[i * N + env]; tq[i] = St.tq[i * N + env];
    tqd[i] = St.tqd[i * N + env]; qf[i] = St.qf[i * N + env];
  }
  for (int a = 0; a < M.n_art; a++) {
    float f[7];
    for (int k = 0; k < 7; k++) f[k] = St.root[(size_t)(a * 7 + k) * N + env];
    root[a] = pose7(f);
  }
  v6 S[C::MAXD], cvp[C::MAXD], pA[C::MAXD], U[C::MAXD], V[C::MAXD];
  pose X[C::MAXD];
  float IA[C::MAXD][36];
  float Dinv[C::MAXD], u[C::MAXD], tau[C::MAXD], arm[C::MAXD];
  v3 cW[C::MAXD];
  m3 IwW[C::MAXD];
  v6 fextW[C::MAXD];
  B2S_NO_UNROLL
  for (int i = 0; i < nd; i++) {
    const int p = M.dof_parent[i], a = M.dof_art[i];
    const v3 Oa = root[a].p;
    float* o = St.kin_link + (size_t)(i * B2S_KL) * N + env;
    if (PART != 2) {
      pose Xp = p >= 0 ? X[p] : root[a];
      pose Xj = pmul(Xp, pose7(M.dof_T0 + 7 * i));
      v3 ax = mk3(M.dof_axis[3 * i], M.dof_axis[3 * i + 1], M.dof_axis[3 * i + 2]);
      pose mo = pose_ident();
      bool rev = M.dof_type[i] == 0;
      if (rev) mo.q = qaxis_angle(ax, q[i]);
      else mo.p = ax * q[i];
      X[i] = pmul(Xj, mo);
      X[i].q = qnormalized(X[i].q);
      v3 aw = qrot(Xj.q, ax);
      S[i] = rev ? mk6(aw, cross(X[i].p - Oa, aw)) : mk6(mk3(0, 0, 0), aw);
      v6 Vp = p >= 0 ? V[p] : zero6();
      V[i] = Vp + S[i] * qd[i];
      const float w[B2S_KL] = {X[i].p.x, X[i].p.y, X[i].p.z, X[i].q.w, X[i].q.x, X[i].q.y, X[i].q.z,
                               V[i].a.x, V[i].a.y, V[i].a.z, V[i].l.x, V[i].l.y, V[i].l.z,
                               S[i].a.x, S[i].a.y, S[i].a.z, S[i].l.x, S[i].l.y, S[i].l.z};
      for (int k = 0; k < B2S_KL; k++) o[k * N] = w[k];
    } else {
      float w[B2S_KL];
      for (int k = 0; k < B2S_KL; k++) w[k] = o[k * N];
      X[i] = pose7(w);
      V[i] = mk6(mk3(w[7], w[8], w[9]), mk3(w[10], w[11], w[12]));
      S[i] = mk6(mk3(w[13], w[14], w[15]), mk3(w[16], w[17], w[18]));
    }
    if (PART == 1) continue;
    v6 vj = S[i] * qd[i];
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
    tau[i] = kp * (tq[i] - q[i] - dt * qd[i]) + kd * (tqd[i] - qd[i]) + qf[i] - damp * qd[i];
    arm[i] = armature + dt * kd + dt * dt * kp + dt * damp;
  }
  if (PART != 2) {
    for (int kw = 0; kw < (M.n_pair + 31) >> 5; kw++) St.col_mask[(size_t)kw * N + env] = 0u;
    emit_joint_rows<C>(M, St, env, q, qd, nd);
  }
  if (PART == 1) return;
  // ABA with the implicit drive in the joint diagonal; second pass for force-limited drives (see substep, part 3)
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
      float qd1 = qd[i] + dt * qdd[i];
      float f = kp * (tq[i] - q[i] - dt * qd1) + kd * (tqd[i] - qd1);
      if (fabsf(f) > fl) {
        any_sat = true;
        tau[i] = (f > 0.f ? fl : -fl) + qf[i] - damp * qd[i];
        arm[i] = armature + dt * damp;
      }
    }
    if (!any_sat) break;
  }
  for (int j = 0; j < nd; j++) St.sol_qdd[j * N + env] = qdd[j];
  // M~^-1 columns from unit-torque solves on the cached factorisation
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
      float qd2 = 0.f;
      if (M.dof_art[i] == M.dof_art[j]) {
        int pi = M.dof_parent[i];
        v6 ap = pi >= 0 ? aa[pi] : zero6();
        qd2 = (uu[i] - dot6(U[i], ap)) * Dinv[i];
        aa[i] = ap + S[i] * qd2;
      }
      St.kin_minv[(size_t)(i * nd + j) * N + env] = qd2;
    }
  }
  for (int b = 0; b < M.n_fb; b++) kin_free_body(M, St, env, b);
}

// ------------------------------------------------------------------------------------------------ kin, dynamics half, G lanes per sub-scene
// The same arithmetic as kin_env<C, ND, 2> shared by a group of G lanes (device: G = 8, four sub-scenes per warp; host emulation:
// G = 1).  What is parallel: the per-joint preparation (lanes stride over the joints), the 6x6 articulated-inertia sweeps (lane r owns
// ROW r of every 6x6 matrix and component r of every spatial vector: U = IA S, the rank-one downdate and the shift to the parent are
// row-local, the scalars D = S.U and u = tau - S.pA are reductions over the group by shuffles), the forward acceleration sweep (one
// reduction per joint), the unit-torque solves for the columns of M~^-1 (one column per lane) and the free bodies (one per lane).
// The matrices live in the group's scratch (shared memory on the device) instead of 324+ floats of local memory per lane.
template <int G>
B2S_HD unsigned group_mask(int lane_in_warp) {
  return G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane_in_warp & ~(G - 1)));
}
template <int G>
B2S_HD float group_sum_m(float x, unsigned mask) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 1; o < G; o <<= 1) x += __shfl_xor_sync(mask, x, o);
#endif
  (void)mask;
  return x;
}
B2S_HD void group_sync_m(unsigned mask) {
#if defined(__CUDA_ARCH__)
  __syncwarp(mask);
#endif
  (void)mask;
}

template <class C, int G>
struct KinDynScratch {
  float S[C::MAXD][6], V[C::MAXD][6], cvp[C::MAXD][6], pA[C::MAXD][6], U[C::MAXD][6], acc[C::MAXD][6], fext[C::MAXD][6];
  float IA[C::MAXD][36];
  float Iw[C::MAXD][9], cW[C::MAXD][3];
  float Dinv[C::MAXD], u[C::MAXD], tau[C::MAXD], arm[C::MAXD], qdd[C::MAXD], q[C::MAXD], qd[C::MAXD], tq[C::MAXD], tqd[C::MAXD], qf[C::MAXD];
  float col_uu[G][C::MAXD], col_aa[G][C::MAXD][6];
  int sat;
};

B2S_HD v6 load6(const float* p) { return mk6(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5])); }
B2S_HD void store6(float* p, v6 x) { p[0] = x.a.x; p[1] = x.a.y; p[2] = x.a.z; p[3] = x.l.x; p[4] = x.l.y; p[5] = x.l.z; }

// lane: 0 .. G-1 within the group; mask: the group's lanes of the warp (group_mask)
template <class C, int G>
B2S_HDN void kin_dyn_group(const DevModel& M, const DevState& St, int env, int lane, unsigned mask, KinDynScratch<C, G>& W) {
  const size_t N = M.n_envs;
  const int nd = M.n_dof;
  const float dt = M.dt;
  const v3 grav = mk3(M.gx, M.gy, M.gz);
  // ---- per joint: state, FK results of part 1, bias forces, world inertia, drive torque and implicit diagonal
  for (int i = lane; i < nd; i += G) {
    const float q = St.q[i * N + env], qd = St.qd[i * N + env], tq = St.tq[i * N + env], tqd = St.tqd[i * N + env], qf = St.qf[i * N + env];
    W.q[i] = q; W.qd[i] = qd; W.tq[i] = tq; W.tqd[i] = tqd; W.qf[i] = qf;
    const float* o = St.kin_link + (size_t)(i * B2S_KL) * N + env;
    float w[B2S_KL];
    for (int k = 0; k < B2S_KL; k++) w[k] = o[k * N];
    const pose X = pose7(w);
    const v6 V = mk6(mk3(w[7], w[8], w[9]), mk3(w[10], w[11], w[12]));
    const v6 S = mk6(mk3(w[13], w[14], w[15]), mk3(w[16], w[17], w[18]));
    const int a = M.dof_art[i];
    const v3 Oa = mk3(St.root[(size_t)(a * 7) * N + env], St.root[(size_t)(a * 7 + 1) * N + env], St.root[(size_t)(a * 7 + 2) * N + env]);
    store6(W.S[i], S); store6(W.V[i], V);
    store6(W.cvp[i], crm(V, S * qd));
    const m3 Rm = qmat(X.q);
    const v3 com = mk3(M.dof_com[3 * i], M.dof_com[3 * i + 1], M.dof_com[3 * i + 2]);
    const v3 c = X.p + mul(Rm, com) - Oa;
    const float* in6 = M.dof_inertia + 6 * i;
    const m3 Iw = mul(mul(Rm, sym6(in6[0], in6[1], in6[2], in6[3], in6[4], in6[5])), transpose(Rm));
    const float mass = M.dof_mass[i];
    W.cW[i][0] = c.x; W.cW[i][1] = c.y; W.cW[i][2] = c.z;
    for (int k = 0; k < 9; k++) W.Iw[i][k] = Iw.m[k];
    const v3 fg = grav * (mass * M.dof_gravity[i]);
    store6(W.fext[i], mk6(cross(c, fg), fg));
    const float kp = M.dof_drive[4 * i], kd = M.dof_drive[4 * i + 1];
    const float damp = M.dof_passive[4 * i], armature = M.dof_passive[4 * i + 2];
    W.tau[i] = kp * (tq - q - dt * qd) + kd * (tqd - qd) + qf - damp * qd;
    W.arm[i] = armature + dt * kd + dt * dt * kp + dt * damp;
  }
  group_sync_m(mask);
  // ---- ABA with the implicit drive in the joint diagonal; second pass for force-limited drives
  for (int pass = 0; pass < 2; pass++) {
    for (int i = lane; i < nd; i += G) {
      float I[36];
      m3 Iw;
      for (int k = 0; k < 9; k++) Iw.m[k] = W.Iw[i][k];
      spatial_inertia(I, M.dof_mass[i], mk3(W.cW[i][0], W.cW[i][1], W.cW[i][2]), Iw);
      for (int k = 0; k < 36; k++) W.IA[i][k] = I[k];
      const v6 V = load6(W.V[i]);
      store6(W.pA[i], crf(V, m6mul(I, V)) - load6(W.fext[i]));
    }
    if (lane == 0) W.sat = 0;
    group_sync_m(mask);
    B2S_NO_UNROLL
    for (int i = nd - 1; i >= 0; i--) {
      const int p = M.dof_parent[i];
      float d = 0.f, e = 0.f;
      for (int r = lane; r < 6; r += G) {
        float ur = 0.f;
        for (int c = 0; c < 6; c++) ur += W.IA[i][6 * r + c] * W.S[i][c];
        W.U[i][r] = ur;
        d += W.S[i][r] * ur;
        e += W.S[i][r] * W.pA[i][r];
      }
      d = group_sum_m<G>(d, mask);
      e = group_sum_m<G>(e, mask);
      const float Dinv = 1.f / (d + W.arm[i]);
      const float ui = W.tau[i] - e;
      if (lane == 0) { W.Dinv[i] = Dinv; W.u[i] = ui; }
      group_sync_m(mask);  // U of this joint is complete
      if (p >= 0) {
        for (int r = lane; r < 6; r += G) {
          const float ur = W.U[i][r];
          float par = W.pA[i][r] + ur * (ui * Dinv);
          for (int c = 0; c < 6; c++) {
            const float ia = W.IA[i][6 * r + c] - ur * W.U[i][c] * Dinv;
            par += ia * W.cvp[i][c];
            W.IA[p][6 * r + c] += ia;
          }
          W.pA[p][r] += par;  // rows are owned by lanes: the parent's row r is only ever touched by this lane
        }
      }
    }
    B2S_NO_UNROLL
    for (int i = 0; i < nd; i++) {
      const int p = M.dof_parent[i];
      float t = 0.f, apr[6];
      for (int r = lane; r < 6; r += G) {
        apr[r] = (p >= 0 ? W.acc[p][r] : 0.f) + W.cvp[i][r];
        t += W.U[i][r] * apr[r];
      }
      t = group_sum_m<G>(t, mask);
      const float qdd = (W.u[i] - t) * W.Dinv[i];
      for (int r = lane; r < 6; r += G) W.acc[i][r] = apr[r] + W.S[i][r] * qdd;
      if (lane == 0) W.qdd[i] = qdd;
    }
    group_sync_m(mask);
    if (pass == 1) break;
    for (int i = lane; i < nd; i += G) {
      const float kp = M.dof_drive[4 * i], kd = M.dof_drive[4 * i + 1], fl = M.dof_drive[4 * i + 2];
      const float damp = M.dof_passive[4 * i], armature = M.dof_passive[4 * i + 2];
      const float qd1 = W.qd[i] + dt * W.qdd[i];
      const float f = kp * (W.tq[i] - W.q[i] - dt * qd1) + kd * (W.tqd[i] - qd1);
      if (fabsf(f) > fl) {
        W.sat = 1;
        W.tau[i] = (f > 0.f ? fl : -fl) + W.qf[i] - damp * W.qd[i];
        W.arm[i] = armature + dt * damp;
      }
    }
    group_sync_m(mask);
    const int any_sat = W.sat;
    group_sync_m(mask);  // everybody has read the flag before the next pass clears it
    if (!any_sat) break;
  }
  for (int j = lane; j < nd; j += G) St.sol_qdd[j * N + env] = W.qdd[j];
  // ---- M~^-1 columns from unit-torque solves on the cached factorisation: one column per lane
  for (int j = lane; j < nd; j += G) {
    float* uu = W.col_uu[lane];
    for (int i = 0; i < nd; i++) uu[i] = 0.f;
    uu[j] = 1.f;
    v6 carry = load6(W.U[j]) * W.Dinv[j];
    int p = M.dof_parent[j];
    while (p >= 0) {
      uu[p] = -dot6(load6(W.S[p]), carry);
      carry = carry + load6(W.U[p]) * (uu[p] * W.Dinv[p]);
      p = M.dof_parent[p];
    }
    B2S_NO_UNROLL
    for (int i = 0; i < nd; i++) {
      float qd2 = 0.f;
      if (M.dof_art[i] == M.dof_art[j]) {
        const int pi = M.dof_parent[i];
        const v6 ap = pi >= 0 ? load6(W.col_aa[lane][pi]) : zero6();
        qd2 = (uu[i] - dot6(load6(W.U[i]), ap)) * W.Dinv[i];
        store6(W.col_aa[lane][i], ap + load6(W.S[i]) * qd2);
      }
      St.kin_minv[(size_t)(i * nd + j) * N + env] = qd2;
    }
  }
  // ---- free bodies: world centre of mass, inverse mass, world inverse inertia (zero for kinematic bodies)
  for (int b = lane; b < M.n_fb; b += G) kin_free_body(M, St, env, b);
}

// ------------------------------------------------------------------------------------------------ collide
struct PlacedShape {
  pose X;      // world pose of the shape
  v3 size, bc; // half extents / radii, bounding-sphere centre
  float br, bv;  // bounding-sphere radius, speed bound of any point of the shape
  int type;
};

B2S_HDN inline void place_shape(const DevModel& M, const DevState& St, int env, int s, PlacedShape& o) {
  const size_t N = M.n_envs;
  const int kind = M.shape_owner_kind[s], ow = M.shape_owner[s];
  pose own = pose_ident();
  if (kind == OWNER_LINK) {
    float f[7];
    const float* src = ow >= 0 ? St.kin_link + (size_t)(ow * B2S_KL) * N + env : St.root + (size_t)((-ow - 1) * 7) * N + env;
    for (int k = 0; k < 7; k++) f[k] = src[k * N];
    own = pose7(f);
  } else if (kind == OWNER_BODY) {
    float f[7];
    for (int k = 0; k < 7; k++) f[k] = St.fb[(size_t)(ow * 13 + k) * N + env];
    own = pose7(f);
  }
  const int ov = M.shape_ov[s];
  float lp[7], bd[4];
  if (ov >= 0) {
    for (int k = 0; k < 7; k++) lp[k] = M.ov_shape_pose[(size_t)(ov * 7 + k) * N + env];
    for (int k = 0; k < 4; k++) bd[k] = M.ov_shape_bound[(size_t)(ov * 4 + k) * N + env];
    o.size = mk3(M.ov_shape_size[(size_t)(ov * 3) * N + env], M.ov_shape_size[(size_t)(ov * 3 + 1) * N + env],
                 M.ov_shape_size[(size_t)(ov * 3 + 2) * N + env]);
  } else {
    for (int k = 0; k < 7; k++) lp[k] = M.shape_pose[7 * s + k];
    for (int k = 0; k < 4; k++) bd[k] = M.shape_bound[4 * s + k];
    o.size = mk3(M.shape_size[3 * s], M.shape_size[3 * s + 1], M.shape_size[3 * s + 2]);
  }
  o.type = M.shape_type[s];
  o.X = pmul(own, pose7(lp));
  o.X.q = qnormalized(o.X.q);
  o.bc = own.p + qrot(own.q, mk3(bd[0], bd[1], bd[2]));
  o.br = bd[3];
  if (kind == OWNER_LINK && ow >= 0) {
    const float* kl = St.kin_link + (size_t)(ow * B2S_KL) * N + env;
    v3 w = mk3(kl[7 * N], kl[8 * N], kl[9 * N]), vl = mk3(kl[10 * N], kl[11 * N], kl[12 * N]);
    const float* rp = St.root + (size_t)(M.dof_art[ow] * 7) * N + env;
    v3 vc = vl + cross(w, o.bc - mk3(rp[0], rp[N], rp[2 * N]));
    o.bv = norm(vc) + norm(w) * o.br;
  } else if (kind == OWNER_BODY) {
    const float* fb = St.fb + (size_t)(ow * 13) * N + env;
    v3 v = mk3(fb[7 * N], fb[8 * N], fb[9 * N]), w = mk3(fb[10 * N], fb[11 * N], fb[12 * N]);
    o.bv = norm(v) + norm(w) * (o.br + norm(o.bc - own.p));
  } else {
    o.bv = 0.f;
  }
}

// oriented bounding box of a shape in its own frame (centre, half extents)
B2S_HDN inline void shape_obb(const DevModel& M, int s, const PlacedShape& P, v3& c, v3& h) {
  c = mk3(0, 0, 0);
  if (P.type == SH_BOX) h = P.size;
  else if (P.type == SH_SPHERE) h = mk3(P.size.x, P.size.x, P.size.x);
  else if (P.type == SH_CAPSULE) h = mk3(P.size.x + P.size.y, P.size.x, P.size.x);
  else {
    const float* bx = M.hull_aabb + 6 * M.shape_hull[s];
    c = mk3(bx[0], bx[1], bx[2]);
    h = mk3(bx[3], bx[4], bx[5]);
  }
}

// candidate pair k of sub-scene env -> col_n / col_data
B2S_HDN inline void collide_env(const DevModel& M, const DevState& St, int env, int k) {
  const size_t N = M.n_envs;
  const int a = M.pair_a[k], b = M.pair_b[k];
  PlacedShape A, B;
  place_shape(M, St, env, a, A);
  place_shape(M, St, env, b, B);
  const float margin = fminf(2.f * M.contact_offset, M.margin_min + 2.f * M.dt * (A.bv + B.bv));
  const int ta = A.type, tb = B.type;
  if (ta == SH_PLANE || tb == SH_PLANE) {
    if (ta == SH_PLANE && tb == SH_PLANE) return;
    const PlacedShape& P = ta == SH_PLANE ? A : B;
    const PlacedShape& O = ta == SH_PLANE ? B : A;
    m3 Rp = qmat(P.X.q);
    float d = dot(O.bc - P.X.p, col(Rp, 0)) - O.br;
    if (d > margin) return;
  } else {
    v3 dd = A.bc - B.bc;
    float rr = A.br + B.br + margin;
    if (dot(dd, dd) > rr * rr) return;
    v3 cA, hA, cB, hB;
    shape_obb(M, a, A, cA, hA);
    shape_obb(M, b, B, cB, hB);
    // bounding sphere of one against the oriented box of the other, both ways
    for (int w = 0; w < 2; w++) {
      const PlacedShape& Bx = w == 0 ? B : A;
      const PlacedShape& Sp = w == 0 ? A : B;
      v3 cl = w == 0 ? cB : cA, hl = w == 0 ? hB : hA;
      q4 qi = mkq(Bx.X.q.w, -Bx.X.q.x, -Bx.X.q.y, -Bx.X.q.z);
      v3 pl = qrot(qi, Sp.bc - Bx.X.p) - cl;
      v3 ex = mk3(fmaxf(fabsf(pl.x) - hl.x, 0.f), fmaxf(fabsf(pl.y) - hl.y, 0.f), fmaxf(fabsf(pl.z) - hl.z, 0.f));
      float lim = Sp.br + margin;
      if (dot(ex, ex) > lim * lim) return;
    }
    // oriented box against oriented box on the six face axes
    m3 RA = qmat(A.X.q), RB = qmat(B.X.q);
    v3 dAB = (B.X.p + mul(RB, cB)) - (A.X.p + mul(RA, cA));
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
    if (sepmax > margin + 1e-5f) return;
  }
  WShape WA, WB;
  WA.type = ta; WA.X = A.X; WA.R = qmat(A.X.q); WA.size = A.size; WA.verts = nullptr; WA.nverts = 0;
  WB.type = tb; WB.X = B.X; WB.R = qmat(B.X.q); WB.size = B.size; WB.verts = nullptr; WB.nverts = 0;
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
  const int n = collide_pair(WA, WB, margin, out);
  if (n == 0) return;
  // hit bitmap of the sub-scene (cleared by kin): the manifest pass visits the set bits in candidate order
#if defined(__CUDA_ARCH__)
  atomicOr(St.col_mask + (size_t)(k >> 5) * N + env, 1u << (k & 31));
#else
  St.col_mask[(size_t)(k >> 5) * N + env] |= 1u << (k & 31);
#endif
  float* o = St.col_data + (size_t)(k * B2S_CD) * N + env;
  o[0] = (float)n; o[N] = out[0].n.x; o[2 * N] = out[0].n.y; o[3 * N] = out[0].n.z;
  for (int i = 0; i < n; i++) {
    float* w = o + (size_t)(4 + 4 * i) * N;
    w[0] = out[i].p.x; w[N] = out[i].p.y; w[2 * N] = out[i].p.z; w[3 * N] = out[i].sep;
  }
}

// ------------------------------------------------------------------------------------------------ manifest
template <class C>
B2S_HDN void manifest_env(const DevModel& M, const DevState& St, int env) {
  const size_t N = M.n_envs;
  int ovf = 0;
  int n_man = 0, n_points = 0;
  int man_sa[C::MAXMAN], man_sb[C::MAXMAN], man_np[C::MAXMAN];
  v3 man_n[C::MAXMAN], man_p[C::MAXMAN][4];
  float man_s[C::MAXMAN][4], man_mu[C::MAXMAN], man_patch[C::MAXMAN];
  const int max_man = M.max_manifolds < C::MAXMAN ? M.max_manifolds : C::MAXMAN;
  const int max_cp = M.max_contacts < C::MAXCP ? M.max_contacts : C::MAXCP;
  const int n_words = (M.n_pair + 31) >> 5;
  B2S_NO_UNROLL
  for (int kw = 0; kw < n_words; kw++)
  for (unsigned bits = St.col_mask[(size_t)kw * N + env]; bits; bits &= bits - 1) {
    int k = kw * 32;  // + index of the lowest set bit
#if defined(__CUDA_ARCH__)
    k += __ffs((int)bits) - 1;
#else
    for (unsigned t = (bits & (0u - bits)) >> 1; t; t >>= 1) k++;
#endif
    const int a = M.pair_a[k], b = M.pair_b[k];
    // the whole record of the pair is fetched at once (one round trip to L2), the point count selects what is used
    const float* o = St.col_data + (size_t)(k * B2S_CD) * N + env;
    float cd[B2S_CD];
#pragma unroll
    for (int i = 0; i < B2S_CD; i++) cd[i] = o[(size_t)i * N];
    const int n = (int)cd[0];
    const v3 nrm = mk3(cd[1], cd[2], cd[3]);
    v3 op[4];
    float os[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      op[i] = mk3(cd[4 + 4 * i], cd[5 + 4 * i], cd[6 + 4 * i]);
      os[i] = cd[7 + 4 * i] - M.rest_offset;
    }
    // patches of the same two bodies with (nearly) the same normal are one friction patch
    int merge = -1;
    for (int mi = 0; mi < n_man; mi++) {
      int qa = man_sa[mi], qb = man_sb[mi];
      if (M.shape_owner_kind[qa] == M.shape_owner_kind[a] && M.shape_owner[qa] == M.shape_owner[a] &&
          M.shape_owner_kind[qb] == M.shape_owner_kind[b] && M.shape_owner[qb] == M.shape_owner[b] &&
          M.shape_row[qa] == M.shape_row[a] && M.shape_row[qb] == M.shape_row[b] && dot(man_n[mi], nrm) > 0.995f) {
        merge = mi;
        break;
      }
    }
    if (merge >= 0) {
      v3 cp[8];
      float cs[8];
      int nc = 0;
      for (int i = 0; i < man_np[merge]; i++) { cp[nc] = man_p[merge][i]; cs[nc] = man_s[merge][i]; nc++; }
      for (int i = 0; i < n; i++) { cp[nc] = op[i]; cs[nc] = os[i]; nc++; }
      int keep[4];
      int kk = reduce4(nc, cp, cs, keep);
      n_points += kk - man_np[merge];
      man_np[merge] = kk;
      for (int i = 0; i < kk; i++) { man_p[merge][i] = cp[keep[i]]; man_s[merge][i] = cs[keep[i]]; }
      man_patch[merge] = fmaxf(man_patch[merge], fmaxf(M.shape_patch[a], M.shape_patch[b]));
      continue;
    }
    if (n_man >= max_man || n_points + n > max_cp) { ovf |= n_man >= max_man ? OVF_MANIFOLDS : OVF_CONTACTS; continue; }
    man_sa[n_man] = a; man_sb[n_man] = b; man_np[n_man] = n; man_n[n_man] = nrm;
    for (int i = 0; i < n; i++) { man_p[n_man][i] = op[i]; man_s[n_man][i] = os[i]; }
    man_mu[n_man] = 0.5f * (M.shape_mu[a] + M.shape_mu[b]);
    man_patch[n_man] = fmaxf(M.shape_patch[a], M.shape_patch[b]);
    n_man++;
    n_points += n;
  }
  // ---- row descriptors
  // the joint rows (tendons, reachable limits) are already there (emit_joint_rows, FK kernel); all of them touch the articulation
  int n_row = St.sol_nrow[env], n_ar = n_row;
  float* RD = St.row_desc + (size_t)env * C::MAXROW * B2S_RD;
  int n_out = 0;
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
    int sides = 0;
    for (int sde = 0; sde < 2; sde++) {
      int kind = M.shape_owner_kind[sh[sde]], ow = M.shape_owner[sh[sde]];
      if (kind == OWNER_LINK && ow >= 0) { any_art = true; sides |= (ow + 1) << (16 * sde); }
      else if (kind == OWNER_BODY) sides |= (ow + 1) << (16 * sde + 8);
    }
    int nrows_needed = np + 2 + (rad > 0.f ? 1 : 0);
    if (n_row + nrows_needed > C::MAXROW || (any_art && n_ar + nrows_needed > C::MAXAR)) { ovf |= n_row + nrows_needed > C::MAXROW ? OVF_ROWS : OVF_ART_ROWS; continue; }
    const int mo = n_out++;
    B2S_NO_UNROLL
    for (int k = 0; k < nrows_needed; k++) {
      int ri = n_row++;
      if (any_art) n_ar++;
      bool is_n = k < np;
      bool tors = k == np + 2;
      v3 pt = is_n ? man_p[mi][k] : cen;
      v3 dir = is_n ? n : (k == np ? t1 : (k == np + 1 ? t2 : n));
      if (is_n) {
        emit_desc(RD, ri, ROW_CONTACT_N | (k == 0 ? ROW_PATCH_START : 0), 0, 0, mo, sides, 0, pt, dir, man_s[mi][k], 0.f, 0.f, -1, 0.f, -1, 0.f);
      } else {
        emit_desc(RD, ri, ROW_FRICTION, first, np, tors ? -1 : mo, sides, tors ? 1 : 0, pt, dir, 0.f, tors ? man_mu[mi] * rad : man_mu[mi], 0.f, -1, 0.f,
                  -1, 0.f);
      }
    }
    float* o = St.man + (size_t)(mo * 8) * N + env;
    o[0] = (float)M.shape_row[man_sa[mi]]; o[N] = (float)M.shape_row[man_sb[mi]];
    o[2 * N] = 0.f; o[3 * N] = 0.f; o[4 * N] = 0.f;
    o[5 * N] = (float)np; o[6 * N] = minsep;
  }
  St.sol_nrow[env] = n_row;
  St.man_count[env] = n_out;
  if (ovf) raise_overflow(St.overflow, ovf);
}

// ------------------------------------------------------------------------------------------------ rowfill
// row r of sub-scene env: descriptor -> unified row record (J_u | B_u | 12 scalars), see b2s_solve.cuh
// L = lanes per sub-scene of the solve that will read the record (fixes the lane-major order of the two vectors, b2s_solve.cuh)
template <class C, int ND, int NUQ, int L>
B2S_HDN void rowfill_env(const DevModel& M, const DevState& St, int env, int r) {
  constexpr int RF = 2 * NUQ + B2S_ROW_SCALARS;
  constexpr int JD = ND > 0 ? ND : C::MAXD;
  const size_t N = M.n_envs;
  const int nd = ND > 0 ? ND : M.n_dof;
  const float* D = St.row_desc + ((size_t)env * C::MAXROW + r) * B2S_RD;
  float* R = St.sol_rows + ((size_t)env * C::MAXROW + r) * RF;
  const int meta = as_int(D[0]), sides = as_int(D[1]);
  const bool tors = (as_int(D[2]) & 1) != 0;
  const int ty = meta & 0x0f;
  const v3 pt = mk3(D[3], D[4], D[5]), dir = mk3(D[6], D[7], D[8]);
  const float gamma = D[11];
  float J[JD];
  bool has_art = false;
  if (ty == ROW_EQ || ty == ROW_LIMIT) {
    const int dA = as_int(D[12]), dB = as_int(D[14]);
    const float cA = D[13], cB = D[15];
#pragma unroll
    for (int j = 0; j < JD; j++) J[j] = j == dA ? cA : (j == dB ? cB : 0.f);
    has_art = true;
  } else {
#pragma unroll
    for (int j = 0; j < JD; j++) J[j] = 0.f;
    for (int sde = 0; sde < 2; sde++) {
      const int ow = ((sides >> (16 * sde)) & 0xff) - 1;
      if (ow < 0) continue;
      has_art = true;
      const float sg = sde == 0 ? 1.f : -1.f;
      const float* rp = St.root + (size_t)(M.dof_art[ow] * 7) * N + env;
      const v3 rr = pt - mk3(rp[0], rp[N], rp[2 * N]);
      const v6 F = tors ? mk6(dir, mk3(0, 0, 0)) : mk6(cross(rr, dir), dir);
      const unsigned mask = M.dof_anc_mask[ow];
#pragma unroll
      for (int j = 0; j < JD; j++) {
        if (j < nd && (mask & (1u << j))) {
          const float* kl = St.kin_link + (size_t)(j * B2S_KL + 13) * N + env;
          const v6 Sj = mk6(mk3(kl[0], kl[N], kl[2 * N]), mk3(kl[3 * N], kl[4 * N], kl[5 * N]));
          J[j] += sg * dot6(Sj, F);
        }
      }
    }
  }
  // zero the two vectors (16-byte stores), then scatter the non-zero blocks
  {
    struct alignas(16) f4 { float x, y, z, w; };
    f4* R4 = reinterpret_cast<f4*>(R);
    f4 z;
    z.x = 0.f; z.y = 0.f; z.z = 0.f; z.w = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * NUQ / 4; k++) R4[k] = z;
  }
  float d = 0.f;
  if (has_art) {
#pragma unroll
    for (int i = 0; i < JD; i++) {
      if (i < nd) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < JD; j++)
          if (j < nd) s += St.kin_minv[(size_t)(i * nd + j) * N + env] * J[j];
        R[row_pos<L, NUQ>(i)] = J[i];
        R[NUQ + row_pos<L, NUQ>(i)] = s;
        d += J[i] * s;  // effective mass of the articulation block, summed in joint order like the fused substep
      }
    }
  }
  float ckin = 0.f;
  for (int sde = 0; sde < 2; sde++) {
    const int b = ((sides >> (16 * sde + 8)) & 0xff) - 1;
    if (b < 0) continue;
    const float sg = sde == 0 ? 1.f : -1.f;
    const float* kf = St.kin_fb + (size_t)(b * B2S_KF) * N + env;
    const v3 fcom = mk3(kf[0], kf[N], kf[2 * N]);
    const float finvm = kf[3 * N];
    m3 fI;
#pragma unroll
    for (int k = 0; k < 9; k++) fI.m[k] = kf[(4 + k) * N];
    const v3 lin = tors ? mk3(0, 0, 0) : dir * sg;
    const v3 ang = tors ? dir * sg : cross(pt - fcom, dir) * sg;
    const v3 Bl = lin * finvm;
    const v3 Ba = mul(fI, ang);
    d += dot(lin, Bl) + dot(ang, Ba);
    const int o = M.fb_slot[b];
    if (o < 0) {
      const float* fb = St.fb + (size_t)(b * 13) * N + env;
      ckin += dot(lin, mk3(fb[7 * N], fb[8 * N], fb[9 * N])) + dot(ang, mk3(fb[10 * N], fb[11 * N], fb[12 * N]));
      continue;
    }
    const float jl[6] = {lin.x, lin.y, lin.z, ang.x, ang.y, ang.z}, bl[6] = {Bl.x, Bl.y, Bl.z, Ba.x, Ba.y, Ba.z};
#pragma unroll
    for (int k = 0; k < 6; k++) { R[row_pos<L, NUQ>(o + k)] = jl[k]; R[NUQ + row_pos<L, NUQ>(o + k)] = bl[k]; }
  }
  float* Sc = R + 2 * NUQ;
  const float sc[12] = {(d + gamma) > 1e-12f ? 1.f / (d + gamma) : 0.f, gamma, D[9], D[10], D[0], ckin, 0.f, 0.f, dir.x, dir.y, dir.z, 0.f};
#pragma unroll
  for (int k = 0; k < 12; k++) Sc[k] = sc[k];
}

}  // namespace b2s
