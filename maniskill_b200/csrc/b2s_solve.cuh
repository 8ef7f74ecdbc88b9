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

// Position of slot s of the unified vector inside a row record: lane-major, so that the SL slots of a lane are contiguous and a
// lane fetches its slice of J (and of B) with 16-byte loads.  Identity for L = 1 (host emulation).
template <int L, int NUQ>
B2S_HD int row_pos(int s) {
  return (s % L) * (NUQ / L) + s / L;
}

// n consecutive floats (n % 4 == 0 and 16-byte aligned on the device)
template <int n>
B2S_HD void load_vec(const float* p, float* out) {
#if defined(__CUDA_ARCH__)
  if (n % 4 == 0) {
#pragma unroll
    for (int k = 0; k < n / 4; k++) {
      const float4 v = *reinterpret_cast<const float4*>(p + 4 * k);
      out[4 * k] = v.x; out[4 * k + 1] = v.y; out[4 * k + 2] = v.z; out[4 * k + 3] = v.w;
    }
    return;
  }
#endif
#pragma unroll
  for (int k = 0; k < n; k++) out[k] = p[k];
}

// Row record (RF = 2 NUQ + 12 floats, 16-byte aligned): J (lane-major) | B (lane-major) | dinv gamma s0 mu | meta ckin 0 0 | dir(3) 0
struct alignas(8) LamTot {
  float lam, tot;  // impulse of a row, its accumulation over the step
};

#define B2S_SC_META 4
#define B2S_SC_CKIN 5
#define B2S_SC_DIR 8

// One sub-scene, executed by the L lanes of its group (lane = 0..L-1).  `valid` is false for padding groups past n_envs (they only
// take part in the shuffles).  lamtot (impulse, accumulated impulse pairs) / stage are group-private scratch (shared memory on the
// device).
template <int L, int NUQ, int MAXROW>
B2S_HDN void solve_env(const DevModel& M, const DevState& St, int env, int lane, bool valid, int nrow_max, LamTot* lamtot, float* stage) {
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
  float u[SL], du[SL], uf[SL], ac[SL], af[SL], dm[SL], umax[SL];
#pragma unroll
  for (int k = 0; k < SL; k++) {
    const int s = k * L + lane;
    u[k] = 0.f; du[k] = 0.f; ac[k] = 0.f; af[k] = 0.f; dm[k] = 1.f; uf[k] = 0.f; umax[k] = 3.0e38f;
    if (!valid) continue;
    if (s < nd) {
      u[k] = St.qd[s * N + env];
      af[k] = h * St.sol_qdd[s * N + env];
      const float vmax = M.dof_drive[4 * s + 3];  // PhysX maxJointVelocity (0 = no clamp)
      if (vmax > 0.f) umax[k] = vmax;
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
  for (int r = lane; r < n_row; r += L) { lamtot[r].lam = 0.f; lamtot[r].tot = 0.f; }
  group_sync();
  const float* myJ = rows + lane * SL;        // this lane's slice of J; its slice of B is NUQ floats further
  for (int it = 0; it < npos + M.n_vel_iters; it++) {
    const bool relax = it >= npos;
    if (!relax) {
#pragma unroll
      for (int k = 0; k < SL; k++) {
        u[k] = (u[k] + af[k]) * dm[k];
        uf[k] = u[k];
        if (it > 0) u[k] += ac[k];
      }
    }
    // Software pipeline over two register sets (A: even rows, B: odd rows): the Jacobian slice, the response slice and the scalars
    // of the next row are in flight while the current one is reduced.  Fetches are unconditional (index clamped to the last row).
    float JA[SL], BA[SL], scA[4], mcA[2], JB[SL], BB[SL], scB[4], mcB[2];
    const int r_last = n_row > 0 ? n_row - 1 : 0;
#define B2S_FETCH_ROW(rr, J_, B_, sc_, mc_)                                      \
  {                                                                              \
    const int rc_ = (rr) < r_last ? (rr) : r_last;                               \
    const float* Rn_ = myJ + (size_t)rc_ * RF;                                   \
    load_vec<SL>(Rn_, J_);                                                       \
    load_vec<SL>(Rn_ + NUQ, B_);                                                 \
    if ((rr) >= n_row) { _Pragma("unroll") for (int k = 0; k < SL; k++) B_[k] = 0.f; }  \
    const float* Sn_ = rows + (size_t)rc_ * RF + 2 * NUQ;                        \
    load_vec<4>(Sn_, sc_);                                                       \
    mc_[0] = Sn_[B2S_SC_META]; mc_[1] = Sn_[B2S_SC_CKIN];                        \
  }
    // Branch-free row visit (the lanes of a warp hold rows of different types).  The friction bound mu x (sum of the normal
    // impulses of the patch) comes from a running sum: the normal rows of a patch directly precede its friction rows.
    //   nl = clamp(lamr - dinv * ms * (jv + bias) - c * lamr, lo, hi)
    //   contact / limit: open (sep > 0): exact bias sep/h; else soft constraint during the position sweeps, no bias in the relaxation
    //   tendon: bias sep/h (position sweeps), regularised by gamma;   friction: no bias, box bound
    // Impulse and its accumulation over the step (position sweeps add the sub-step impulse, the relaxation its correction) are
    // one 8-byte pair; every lane of the group writes the same pair.
#define B2S_VISIT_ROW(rr, J_, B_, sc_, mc_)                                                          \
  {                                                                                                  \
    const bool act_ = (rr) < n_row;                                                                  \
    const int meta_ = as_int(mc_[0]);                                                                \
    const int ty_ = meta_ & 0x0f;                                                                    \
    const bool fric_ = ty_ == ROW_FRICTION, eq_ = ty_ == ROW_EQ;                                     \
    if (meta_ & ROW_PATCH_START) pn_sum = 0.f;                                                       \
    LamTot* lt_ = lamtot + (act_ ? (rr) : 0);                                                        \
    const LamTot lt0_ = *lt_;                                                                        \
    const float lamr_ = lt0_.lam;                                                                    \
    float jp_[2] = {0.f, 0.f}, sp_[2] = {0.f, 0.f};                                                  \
    _Pragma("unroll") for (int k = 0; k < SL; k++) {                                                 \
      jp_[k & 1] += J_[k] * u[k];                                                                    \
      sp_[k & 1] += J_[k] * du[k];                                                                   \
    }                                                                                                \
    const float jvr_ = group_sum<L>(jp_[0] + jp_[1]);  /* critical chain: u -> dot -> reduce -> update -> u */                  \
    const float sep_ = sc_[2] + group_sum<L>(sp_[0] + sp_[1]);                                       \
    const float dinv_ = sc_[0];                                                                      \
    const bool open_ = sep_ > 0.f;                                                                   \
    const float sh_ = sep_ * inv_h;                                                                  \
    const float bias_c_ = open_ ? sh_ : (relax ? 0.f : fmaxf(soft_rate * sep_, -max_depen));         \
    const float bias_ = fric_ ? 0.f : (eq_ ? (relax ? 0.f : sh_) : bias_c_);                         \
    const bool use_soft_ = !fric_ && !eq_ && !open_ && !relax;                                       \
    const float dm_ = use_soft_ ? dinv_ * soft_mass : dinv_;                                         \
    const float c_ = eq_ ? sc_[1] * dinv_ : (use_soft_ ? soft_imp : 0.f);                            \
    const float lim_ = sc_[3] * pn_sum;                                                              \
    const float lo_ = fric_ ? -lim_ : (eq_ ? -3.0e38f : 0.f), hi_ = fric_ ? lim_ : 3.0e38f;          \
    /* everything that does not depend on the reduced dot product is folded beforehand: nl = t - dm * (jv + bk) */ \
    const float bk_ = bias_ + mc_[1];               /* bias + constant contribution of kinematic bodies */ \
    const float t_ = lamr_ - c_ * lamr_;                                                             \
    float nl_ = t_ - dm_ * (jvr_ + bk_);                                                             \
    nl_ = fminf(fmaxf(nl_, lo_), hi_);                                                               \
    const float dl_ = nl_ - lamr_;                                                                   \
    if (act_) {                                                                                      \
      LamTot lt1_;                                                                                   \
      lt1_.lam = nl_; lt1_.tot = lt0_.tot + (relax ? dl_ : nl_);                                     \
      *lt_ = lt1_;                                                                                   \
    }                                                                                                \
    if (ty_ == ROW_CONTACT_N) pn_sum += nl_;                                                         \
    _Pragma("unroll") for (int k = 0; k < SL; k++) u[k] += B_[k] * dl_;  /* B_ is zero for padding rows */  \
  }
    float pn_sum = 0.f;
    B2S_FETCH_ROW(0, JA, BA, scA, mcA)
    B2S_NO_UNROLL
    for (int r = 0; r < nrow_max; r += 2) {
      B2S_FETCH_ROW(r + 1, JB, BB, scB, mcB)
      B2S_VISIT_ROW(r, JA, BA, scA, mcA)
      B2S_FETCH_ROW(r + 2, JA, BA, scA, mcA)
      B2S_VISIT_ROW(r + 1, JB, BB, scB, mcB)
    }
#undef B2S_FETCH_ROW
#undef B2S_VISIT_ROW
#pragma unroll
    for (int k = 0; k < SL; k++) u[k] = fmaxf(-umax[k], fminf(umax[k], u[k]));  // joint velocity clamp after every sweep
    if (!relax) {
#pragma unroll
      for (int k = 0; k < SL; k++) {
        ac[k] = u[k] - uf[k];
        du[k] += h * u[k];
      }
    }
  }
  // ---- integrate + export (the vector is gathered through the staging area)
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
  // output slots congruent to it (the rows of a slot are consecutive), so the accumulation order per slot is the row order
  int cur = -1;
  v3 acc = mk3(0, 0, 0);
  for (int r = 0; r <= n_row; r++) {
    int slot = -1;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < n_row) {
      const float* Sc = rows + (size_t)r * RF + 2 * NUQ;
      slot = ((as_int(Sc[B2S_SC_META]) >> 24) & 0xff) - 1;
      if (slot >= 0 && slot % L == lane) load_vec<4>(Sc + B2S_SC_DIR, d);
      else slot = -1;
    }
    if (slot != cur && cur >= 0) {
      float* o = St.man + (size_t)(cur * 8) * N + env;
      o[2 * N] = acc.x; o[3 * N] = acc.y; o[4 * N] = acc.z;
      acc = mk3(0, 0, 0);
    }
    if (slot >= 0) {
      const float t = lamtot[r].tot;
      acc = acc + mk3(d[0], d[1], d[2]) * t;
    }
    cur = slot;
  }
}

}  // namespace b2s
