// b200sim pipelined substep, phase B: the sub-stepped soft TGS solve + integration of one sub-scene by a GROUP of L lanes.
//
// Phase A (b2s_pipe.cuh: kin -> collide -> manifest -> rowfill) leaves, per sub-scene, a table of constraint rows in a
// unified layout: the generalised velocity of the sub-scene is one vector u[NUQ] = [joint velocities (n_dof) | per DYNAMIC free
// body: linear(3) angular(3)] (NUQ = 16 or 28 slots), and every row is two dense NUQ-vectors (Jacobian Ju, response
// Bu = M~^-1 Ju^T) plus 12 scalars (16-byte aligned records of 176 or 272 bytes).  Phase B runs the 15 + 1 Gauss-Seidel sweeps
// with the vector sliced over the L lanes of a group (slot s belongs to lane s % L): a row visit is 2 x NUQ/L multiply-adds, a
// two-stage shuffle reduction shared by both dot products, the scalar update done redundantly by the group, NUQ/L multiply-adds
// for the response; the next row is fetched while the current one is reduced.  That shortens the dependent chain of a row visit,
// puts 4x more warps on the machine where the 1-lane kernel leaves SMs at one warp, and the kernel is small enough to stay resident
// in the instruction caches.  The same code with L = 1 is what the host-side test emulation runs.
//
// Replaces (together with phase A) `PhysxGpuSystem.step()` -- reference call site mani_skill/envs/scene.py:379-380.
#pragma once
#include "b2s_step.cuh"

namespace b2s {

#define B2S_ROW_SCALARS 12  // dinv gamma s0 mu meta dir(3) ckin pad(3)

template <int L>
B2S_HD float group_sum(float x) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 1; o < L; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
#endif
  return x;
}

B2S_HD int as_int(float f) {
  union { int i; float f; } c;
  c.f = f;
  return c.i;
}

B2S_HD void group_sync() {
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
}

// One sub-scene, executed by the L lanes of its group (lane = 0..L-1).  `valid` is false for padding groups past n_envs (they only
// take part in the shuffles).  lam / tot / stage are group-private scratch (shared memory on the device).
template <int L, int NUQ, int MAXROW>
B2S_HDN void solve_env(const DevModel& M, const DevState& St, int env, int lane, bool valid, int nrow_max, float* lam, float* tot, float* stage) {
  constexpr int SL = NUQ / L;                       // slots per lane
  constexpr int RF = 2 * NUQ + B2S_ROW_SCALARS;     // floats per row record
  const size_t N = M.n_envs;
  const int nd = M.n_dof, nfb = M.n_fb;
  const int npos = M.n_pos_iters;
  const float dt = M.dt, h = dt / npos;
  const float kPi = 3.14159265358979323846f;
  const float omega = 2.f * kPi * fminf(M.contact_hertz, 0.25f / h), zeta = M.contact_zeta;
  const float sa1 = 2.f * zeta + h * omega, sa2 = h * omega * sa1, sa3 = 1.f / (1.f + sa2);
  const float soft_rate = omega / sa1, soft_mass = sa2 * sa3, soft_imp = sa3;
  const float inv_h = 1.f / h, max_depen = M.max_depen_vel;
  const int n_row = valid ? St.sol_nrow[env] : 0;
  const float* rows = St.sol_rows + (size_t)(valid ? env : 0) * MAXROW * RF;
  float u[SL], du[SL], uf[SL], ac[SL], af[SL], dm[SL];
#pragma unroll
  for (int k = 0; k < SL; k++) {
    const int s = k * L + lane;
    u[k] = 0.f; du[k] = 0.f; ac[k] = 0.f; af[k] = 0.f; dm[k] = 1.f; uf[k] = 0.f;
    if (!valid) continue;
    if (s < nd) {
      u[k] = St.qd[s * N + env];
      af[k] = h * St.sol_qdd[s * N + env];
    } else if (s < M.n_u) {
      // slot -> (dynamic body, component)
      int b = 0;
      for (int bb = 0; bb < nfb; bb++)
        if (M.fb_slot[bb] >= 0 && M.fb_slot[bb] <= s) b = bb;
      const int c = s - M.fb_slot[b];
      u[k] = St.fb[(size_t)(b * 13 + 7 + c) * N + env];
      if (c < 3) {
        const float g = c == 0 ? M.gx : (c == 1 ? M.gy : M.gz);
        af[k] = g * (h * M.fb_gravity[b]);
        dm[k] = fmaxf(0.f, 1.f - h * M.fb_damping[2 * b]);
      } else {
        dm[k] = fmaxf(0.f, 1.f - h * M.fb_damping[2 * b + 1]);
      }
    }
  }
  for (int r = lane; r < n_row; r += L) { lam[r] = 0.f; tot[r] = 0.f; }
  group_sync();
  for (int it = 0; it < npos + M.n_vel_iters; it++) {
    const bool relax = it >= npos;
    if (!relax) {
#pragma unroll
      for (int k = 0; k < SL; k++) {
        u[k] = (u[k] + af[k]) * dm[k];
        uf[k] = u[k];
        if (it > 0) u[k] += ac[k];
      }
    } else {
      for (int r = lane; r < n_row; r += L) tot[r] -= lam[r];
      group_sync();
    }
    // software pipeline: the Jacobian slice, the response slice and the scalars of row r+1 are in flight while row r is reduced
    float Jn[SL], Bn[SL], scn[6];
    if (n_row > 0) {
#pragma unroll
      for (int k = 0; k < SL; k++) { Jn[k] = rows[k * L + lane]; Bn[k] = rows[NUQ + k * L + lane]; }
#pragma unroll
      for (int k = 0; k < 5; k++) scn[k] = rows[2 * NUQ + k];
      scn[5] = rows[2 * NUQ + 8];
    }
    float pn_sum = 0.f;
#pragma unroll 2
    for (int r = 0; r < nrow_max; r++) {
      const bool act = r < n_row;
      float Jc[SL], Bc[SL], sc[6];
#pragma unroll
      for (int k = 0; k < SL; k++) { Jc[k] = Jn[k]; Bc[k] = Bn[k]; }
#pragma unroll
      for (int k = 0; k < 6; k++) sc[k] = scn[k];
      if (r + 1 < n_row) {
        const float* Rn = rows + (size_t)(r + 1) * RF;
#pragma unroll
        for (int k = 0; k < SL; k++) { Jn[k] = Rn[k * L + lane]; Bn[k] = Rn[NUQ + k * L + lane]; }
#pragma unroll
        for (int k = 0; k < 5; k++) scn[k] = Rn[2 * NUQ + k];
        scn[5] = Rn[2 * NUQ + 8];
      }
      // branch-free row visit (the lanes of a warp hold rows of different types).  The friction bound mu x (sum of the normal
      // impulses of the patch) comes from a running sum: the normal rows of a patch directly precede its friction rows.
      const int meta = as_int(sc[4]);
      const int ty = meta & 0x0f;
      const bool fric = ty == ROW_FRICTION, eq = ty == ROW_EQ;
      if (meta & ROW_PATCH_START) pn_sum = 0.f;
      const float lamr = lam[act ? r : 0];
      float jp[2] = {0.f, 0.f}, sp[2] = {0.f, 0.f};
#pragma unroll
      for (int k = 0; k < SL; k++) {
        jp[k & 1] += Jc[k] * u[k];
        sp[k & 1] += Jc[k] * du[k];
      }
      const float jv = group_sum<L>(jp[0] + jp[1]) + sc[5];  // + constant contribution of kinematic bodies
      const float sep = sc[2] + group_sum<L>(sp[0] + sp[1]);
      const float dinv = sc[0], gamma = sc[1], mu = sc[3];
      // nl = clamp(lamr - dinv * ms * (jv + bias) - c * lamr, lo, hi)
      //   contact / limit: open (sep > 0): exact bias sep/h; else soft constraint during the position sweeps, no bias in the relaxation
      //   tendon: bias sep/h (position sweeps), regularised by gamma;   friction: no bias, box bound
      const bool open = sep > 0.f;
      const float sh = sep * inv_h;
      const float bias_c = open ? sh : (relax ? 0.f : fmaxf(soft_rate * sep, -max_depen));
      const float bias = fric ? 0.f : (eq ? (relax ? 0.f : sh) : bias_c);
      const bool use_soft = !fric && !eq && !open && !relax;
      const float ms = use_soft ? soft_mass : 1.f;
      const float c = eq ? gamma * dinv : (use_soft ? soft_imp : 0.f);
      const float lim = mu * pn_sum;
      const float lo = fric ? -lim : (eq ? -3.0e38f : 0.f), hi = fric ? lim : 3.0e38f;
      float nl = lamr - dinv * ms * (jv + bias) - c * lamr;
      nl = fminf(fmaxf(nl, lo), hi);
      const float dl = act ? nl - lamr : 0.f;
      if (act) lam[r] = nl;  // every lane of the group writes the same value
      if (ty == ROW_CONTACT_N) pn_sum += nl;
#pragma unroll
      for (int k = 0; k < SL; k++) u[k] += Bc[k] * dl;
    }
    group_sync();
    for (int r = lane; r < n_row; r += L) tot[r] += lam[r];
    if (!relax) {
#pragma unroll
      for (int k = 0; k < SL; k++) {
        ac[k] = u[k] - uf[k];
        du[k] += h * u[k];
      }
    }
  }
  // ---- integrate + export (the first lane of the group; the vector is gathered through the staging area)
#pragma unroll
  for (int k = 0; k < SL; k++) {
    stage[k * L + lane] = u[k];
    stage[NUQ + k * L + lane] = du[k];
  }
  group_sync();
  if (!valid) return;
  // the lanes of the group share the export: joints and bodies strided by lane, patch impulses by output slot
  for (int i = lane; i < nd; i += L) {
    const float v1 = stage[i], qd0 = St.qd[i * N + env];
    St.qacc[i * N + env] = (v1 - qd0) / dt;
    St.q[i * N + env] += stage[NUQ + i];
    St.qd[i * N + env] = v1;
  }
  for (int b = lane; b < nfb; b += L) {
    const int o_ = M.fb_slot[b];
    if (o_ < 0) continue;
    float f[7];
    for (int k = 0; k < 7; k++) f[k] = St.fb[(size_t)(b * 13 + k) * N + env];
    pose X = pose7(f);
    v3 com;
    const int ov = M.fb_ov[b];
    if (ov >= 0) com = mk3(M.ov_fb_mass[(size_t)(ov * 10 + 1) * N + env], M.ov_fb_mass[(size_t)(ov * 10 + 2) * N + env], M.ov_fb_mass[(size_t)(ov * 10 + 3) * N + env]);
    else com = mk3(M.fb_com[3 * b], M.fb_com[3 * b + 1], M.fb_com[3 * b + 2]);
    const float* sv = stage + o_;
    const float* sdv = stage + NUQ + o_;
    v3 fcom = X.p + mul(qmat(X.q), com);
    v3 cnew = fcom + mk3(sdv[0], sdv[1], sdv[2]);
    q4 qn = qnormalized(qmul(qexp(mk3(sdv[3], sdv[4], sdv[5])), X.q));
    v3 pn = cnew - qrot(qn, com);
    const float o[13] = {pn.x, pn.y, pn.z, qn.w, qn.x, qn.y, qn.z, sv[0], sv[1], sv[2], sv[3], sv[4], sv[5]};
    for (int k = 0; k < 13; k++) St.fb[(size_t)(b * 13 + k) * N + env] = o[k];
  }
  // contact patch impulses: sum over the rows of a patch of (row direction x impulse accumulated over the step); a lane owns the
  // output slots congruent to it, so the accumulation order per slot is the row order
  for (int r = 0; r < n_row; r++) {
    const float* Sc = rows + (size_t)r * RF + 2 * NUQ;
    const int slot = ((as_int(Sc[4]) >> 24) & 0xff) - 1;
    if (slot < 0 || slot % L != lane) continue;
    float* o = St.man + (size_t)(slot * 8) * N + env;
    const float t = tot[r];
    o[2 * N] += Sc[5] * t; o[3 * N] += Sc[6] * t; o[4 * N] += Sc[7] * t;
  }
}

}  // namespace b2s
