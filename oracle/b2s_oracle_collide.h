// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md). Never linked into the product.
//
// Narrowphase restatement.  The reference delegates this to PhysX PCM inside
// `px.step()` (mani_skill/envs/scene.py:379-380); PhysX is not in /root/reference, so the algorithms follow
// the literature: SAT + face clipping for box-box, vertex tests against half-spaces, GJK (Gilbert-Johnson-
// Keerthi, Ericson's simplex sub-algorithms) + EPA (van den Bergen) for everything else that is convex.
// Configuration follows the reference: contact generated when distance < contact_offset_A + contact_offset_B
// (mani_skill/utils/structs/types.py:44-45), at most 4 points per pair.
#pragma once
#include <utility>
#include <vector>
#include "b2s_oracle_math.h"

#define SHAPE_PLANE 0
#define SHAPE_BOX 1
#define SHAPE_SPHERE 2
#define SHAPE_CAPSULE 3
#define SHAPE_CONVEX 4

struct Contact {
  int sa, sb;  // shape ids
  V3 p;        // contact point (midway between the two surfaces), sub-scene frame
  V3 n;        // unit normal pointing from B towards A
  R sep;       // signed distance (negative = penetration)
};

struct WShape {  // shape placed in the sub-scene frame
  int type;
  Pose X;
  M3 Rm;
  V3 size;
  const float* verts;  // convex: hull vertices (local)
  int nverts;
};

// ---------------------------------------------------------------- manifold reduction
// keep <=4 of n candidate points: deepest, farthest from it, largest triangle, farthest from that triangle.
// Keeps <= 4 of n candidate points: the deepest one, then the points that spread the support polygon most, where a candidate
// pays for lying above the deepest one: every length is reduced by (REDUCE4_LAMBDA x height above the deepest point), so the
// near-contact set is preferred over far speculative points (a many-vertex hull resting on a facet keeps that facet).
#define REDUCE4_LAMBDA 20
static inline int reduce4(int n, const V3* p, const R* d, int* keep) {
  if (n <= 4) {
    for (int i = 0; i < n; i++) keep[i] = i;
    return n;
  }
  int i0 = 0;
  for (int i = 1; i < n; i++)
    if (d[i] < d[i0] - R(1e-6)) i0 = i;  // ties (symmetric features) go to the lower index
  const R lam2 = (R)(REDUCE4_LAMBDA * REDUCE4_LAMBDA);
  int i1 = -1;
  R best = (R)-1e30;
  for (int i = 0; i < n; i++) {
    if (i == i0) continue;
    V3 e = p[i] - p[i0];
    R h = d[i] - d[i0];
    R v = dot(e, e) - lam2 * h * h;
    if (v > best + R(1e-4) * std::fabs(best) + R(1e-12)) { best = v; i1 = i; }
  }
  int i2 = -1;
  best = (R)-1e30;
  const V3 base = p[i1] - p[i0];
  const R base2 = dot(base, base);
  for (int i = 0; i < n; i++) {
    if (i == i0 || i == i1) continue;
    V3 c = cross(base, p[i] - p[i0]);
    R h = d[i] - d[i0];
    R v = dot(c, c) - base2 * lam2 * h * h;  // (triangle height)^2 - (lambda h)^2, times base^2
    if (v > best + R(1e-4) * std::fabs(best) + R(1e-12)) { best = v; i2 = i; }
  }
  int i3 = -1;
  best = (R)-1e30;
  for (int i = 0; i < n; i++) {
    if (i == i0 || i == i1 || i == i2) continue;
    V3 e0 = p[i] - p[i0], e1 = p[i] - p[i1], e2 = p[i] - p[i2];
    R h = d[i] - d[i0];
    R v = std::fmin(dot(e0, e0), std::fmin(dot(e1, e1), dot(e2, e2))) - lam2 * h * h;
    if (v > best + R(1e-4) * std::fabs(best) + R(1e-12)) { best = v; i3 = i; }
  }
  keep[0] = i0; keep[1] = i1; keep[2] = i2; keep[3] = i3;
  return 4;
}

// ---------------------------------------------------------------- plane vs vertex sets
static inline int collide_plane_points(const WShape& P, int npts, const V3* pts, R radius, R margin, Contact* out) {
  V3 n = P.Rm.col(0);
  V3 cand[64];
  R dist[64];
  int m = 0;
  for (int i = 0; i < npts && m < 64; i++) {
    R d = dot(pts[i] - P.X.p, n) - radius;
    if (d < margin) {
      cand[m] = pts[i] - n * (radius + d * R(0.5));
      dist[m] = d;
      m++;
    }
  }
  int keep[4];
  int k = reduce4(m, cand, dist, keep);
  for (int i = 0; i < k; i++) {
    out[i].p = cand[keep[i]];
    out[i].n = n;
    out[i].sep = dist[keep[i]];
  }
  return k;
}

static inline void box_corners(const WShape& B, V3* c) {
  for (int i = 0; i < 8; i++) {
    V3 l((i & 1) ? B.size.x : -B.size.x, (i & 2) ? B.size.y : -B.size.y, (i & 4) ? B.size.z : -B.size.z);
    c[i] = B.X.p + B.Rm * l;
  }
}

// A = any non-plane shape, P = plane.  normal from plane towards A.
static inline int collide_plane_any(const WShape& A, const WShape& P, R margin, Contact* out) {
  V3 pts[64];
  if (A.type == SHAPE_BOX) {
    box_corners(A, pts);
    return collide_plane_points(P, 8, pts, 0, margin, out);
  } else if (A.type == SHAPE_SPHERE) {
    pts[0] = A.X.p;
    return collide_plane_points(P, 1, pts, A.size.x, margin, out);
  } else if (A.type == SHAPE_CAPSULE) {
    V3 ax = A.Rm.col(0) * A.size.y;
    pts[0] = A.X.p + ax;
    pts[1] = A.X.p - ax;
    return collide_plane_points(P, 2, pts, A.size.x, margin, out);
  } else {
    int n = A.nverts < 64 ? A.nverts : 64;
    for (int i = 0; i < n; i++) pts[i] = A.X.p + A.Rm * V3(A.verts[3 * i], A.verts[3 * i + 1], A.verts[3 * i + 2]);
    return collide_plane_points(P, n, pts, 0, margin, out);
  }
}

// ---------------------------------------------------------------- box vs box (SAT + clipping)
static inline int clip_poly(int n, const V3* in, V3* out, int axis, R sign, R lim) {
  // keep points with sign*coord <= lim ; V3 = (u, v, depth)
  int m = 0;
  for (int i = 0; i < n; i++) {
    V3 a = in[i], b = in[(i + 1) % n];
    R da = sign * a[axis] - lim, db = sign * b[axis] - lim;
    if (da <= 0) out[m++] = a;
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      R t = da / (da - db);
      out[m++] = a + (b - a) * t;
    }
  }
  return m;
}

static inline int collide_box_box(const WShape& A, const WShape& B, R margin, Contact* out) {
  const M3& RA = A.Rm;
  const M3& RB = B.Rm;
  V3 d = B.X.p - A.X.p;
  V3 dA = tmul(RA, d), dB = tmul(RB, d);
  R C[3][3], AC[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      C[i][j] = dot(RA.col(i), RB.col(j));
      AC[i][j] = std::fabs(C[i][j]);
    }
  R hA[3] = {A.size.x, A.size.y, A.size.z}, hB[3] = {B.size.x, B.size.y, B.size.z};
  R best_face = -1e30;
  int face_code = -1;  // 0..2 = A axis, 3..5 = B axis
  for (int i = 0; i < 3; i++) {
    R s = std::fabs(dA[i]) - (hA[i] + hB[0] * AC[i][0] + hB[1] * AC[i][1] + hB[2] * AC[i][2]);
    if (s > margin) return 0;
    if (s > best_face) { best_face = s; face_code = i; }
  }
  for (int j = 0; j < 3; j++) {
    R s = std::fabs(dB[j]) - (hB[j] + hA[0] * AC[0][j] + hA[1] * AC[1][j] + hA[2] * AC[2][j]);
    if (s > margin) return 0;
    if (s > best_face) { best_face = s; face_code = 3 + j; }
  }
  R best_edge = -1e30;
  int ei = -1, ej = -1;
  V3 edgeL;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      R len2 = R(1) - C[i][j] * C[i][j];
      if (len2 < R(1e-6)) continue;
      R len = std::sqrt(len2);
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      R rA = hA[i1] * AC[i2][j] + hA[i2] * AC[i1][j];
      R rB = hB[j1] * AC[i][j2] + hB[j2] * AC[i][j1];
      R dl = dA[i2] * C[i1][j] - dA[i1] * C[i2][j];
      R s = (std::fabs(dl) - (rA + rB)) / len;
      if (s > margin) return 0;
      if (s > best_edge) {
        best_edge = s; ei = i; ej = j;
        V3 L = cross(RA.col(i), RB.col(j)) * (R(1) / len);
        edgeL = (dot(L, d) < 0) ? -L : L;  // oriented from A to B
      }
    }
  if (ei >= 0 && best_edge > best_face + R(1e-4)) {
    // edge-edge: one point
    V3 L = edgeL;
    V3 pA = A.X.p, pB = B.X.p;
    for (int k = 0; k < 3; k++) {
      if (k != ei) pA = pA + RA.col(k) * ((dot(RA.col(k), L) > 0 ? hA[k] : -hA[k]));
      if (k != ej) pB = pB - RB.col(k) * ((dot(RB.col(k), L) > 0 ? hB[k] : -hB[k]));
    }
    V3 ua = RA.col(ei), ub = RB.col(ej);
    V3 r = pB - pA;
    R uaub = dot(ua, ub), q1 = dot(ua, r), q2 = -dot(ub, r);
    R dd = R(1) - uaub * uaub;
    R s = 0, t = 0;
    if (dd > R(1e-9)) {
      s = (q1 + uaub * q2) / dd;
      t = (uaub * q1 + q2) / dd;
    }
    s = std::fmax(-hA[ei], std::fmin(hA[ei], s));
    t = std::fmax(-hB[ej], std::fmin(hB[ej], t));
    V3 ca = pA + ua * s, cb = pB + ub * t;
    out[0].p = (ca + cb) * R(0.5);
    out[0].n = -L;
    out[0].sep = dot(cb - ca, L);
    return 1;
  }
  // face contact: reference box = owner of the separating face
  bool refA = face_code < 3;
  const WShape& Rf = refA ? A : B;
  const WShape& In = refA ? B : A;
  const M3& RR = Rf.Rm;
  const M3& RI = In.Rm;
  const R* hR = refA ? hA : hB;
  const R* hI = refA ? hB : hA;
  int ax = refA ? face_code : face_code - 3;
  V3 dRI = In.X.p - Rf.X.p;
  V3 nref = RR.col(ax);
  if (dot(nref, dRI) < 0) nref = -nref;  // from reference towards incident
  // incident face: axis of In most anti-parallel to nref
  int jx = 0;
  R bestd = -1;
  for (int j = 0; j < 3; j++) {
    R v = std::fabs(dot(RI.col(j), nref));
    if (v > bestd) { bestd = v; jx = j; }
  }
  R sgn = dot(RI.col(jx), nref) > 0 ? R(-1) : R(1);
  int j1 = (jx + 1) % 3, j2 = (jx + 2) % 3;
  V3 fc = In.X.p + RI.col(jx) * (sgn * hI[jx]);
  V3 e1 = RI.col(j1) * hI[j1], e2 = RI.col(j2) * hI[j2];
  V3 quad[4] = {fc + e1 + e2, fc - e1 + e2, fc - e1 - e2, fc + e1 - e2};
  int u1 = (ax + 1) % 3, u2 = (ax + 2) % 3;
  V3 U = RR.col(u1), Vv = RR.col(u2);
  V3 poly[16], tmp[16];
  for (int i = 0; i < 4; i++) {
    V3 rel = quad[i] - Rf.X.p;
    poly[i] = V3(dot(rel, U), dot(rel, Vv), dot(rel, nref) - hR[ax]);
  }
  int n = 4;
  n = clip_poly(n, poly, tmp, 0, R(1), hR[u1]);
  n = clip_poly(n, tmp, poly, 0, R(-1), hR[u1]);
  n = clip_poly(n, poly, tmp, 1, R(1), hR[u2]);
  n = clip_poly(n, tmp, poly, 1, R(-1), hR[u2]);
  V3 cand[16];
  R dist[16];
  int m = 0;
  for (int i = 0; i < n; i++) {
    if (poly[i].z < margin) {
      V3 w = Rf.X.p + U * poly[i].x + Vv * poly[i].y + nref * (hR[ax] + poly[i].z * R(0.5));
      cand[m] = w;
      dist[m] = poly[i].z;
      m++;
    }
  }
  int keep[4];
  int k = reduce4(m, cand, dist, keep);
  V3 nBA = refA ? -nref : nref;
  for (int i = 0; i < k; i++) {
    out[i].p = cand[keep[i]];
    out[i].n = nBA;
    out[i].sep = dist[keep[i]];
  }
  return k;
}

// ---------------------------------------------------------------- GJK / EPA on convex cores
static inline V3 support_core(const WShape& S, V3 dir) {
  V3 dl = tmul(S.Rm, dir);
  V3 l;
  switch (S.type) {
    case SHAPE_BOX:
      l = V3(dl.x >= 0 ? S.size.x : -S.size.x, dl.y >= 0 ? S.size.y : -S.size.y, dl.z >= 0 ? S.size.z : -S.size.z);
      break;
    case SHAPE_SPHERE:
      l = V3(0, 0, 0);
      break;
    case SHAPE_CAPSULE:
      l = V3(dl.x >= 0 ? S.size.y : -S.size.y, 0, 0);
      break;
    default: {
      int bi = 0;
      R bd = -1e30;
      for (int i = 0; i < S.nverts; i++) {
        R v = dl.x * S.verts[3 * i] + dl.y * S.verts[3 * i + 1] + dl.z * S.verts[3 * i + 2];
        if (v > bd) { bd = v; bi = i; }
      }
      l = V3(S.verts[3 * bi], S.verts[3 * bi + 1], S.verts[3 * bi + 2]);
    }
  }
  return S.X.p + S.Rm * l;
}
static inline R core_radius(const WShape& S) { return (S.type == SHAPE_SPHERE || S.type == SHAPE_CAPSULE) ? S.size.x : R(0); }

struct SVert {
  V3 w, a, b;
};

static inline SVert mink_support(const WShape& A, const WShape& B, V3 dir) {
  SVert s;
  s.a = support_core(A, dir);
  s.b = support_core(B, -dir);
  s.w = s.a - s.b;
  return s;
}

// closest point to the origin on triangle (a,b,c); writes barycentrics; returns bitmask of used vertices
static inline int closest_tri(V3 a, V3 b, V3 c, R* bary) {
  V3 ab = b - a, ac = c - a, ap = -a;
  R d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { bary[0] = 1; bary[1] = 0; bary[2] = 0; return 1; }
  V3 bp = -b;
  R d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { bary[0] = 0; bary[1] = 1; bary[2] = 0; return 2; }
  R vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) {
    R v = d1 / (d1 - d3);
    bary[0] = 1 - v; bary[1] = v; bary[2] = 0;
    return 3;
  }
  V3 cp = -c;
  R d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { bary[0] = 0; bary[1] = 0; bary[2] = 1; return 4; }
  R vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) {
    R w = d2 / (d2 - d6);
    bary[0] = 1 - w; bary[1] = 0; bary[2] = w;
    return 5;
  }
  R va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    R w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    bary[0] = 0; bary[1] = 1 - w; bary[2] = w;
    return 6;
  }
  R denom = R(1) / (va + vb + vc);
  R v = vb * denom, w = vc * denom;
  bary[0] = 1 - v - w; bary[1] = v; bary[2] = w;
  return 7;
}

struct GjkOut {
  int status;  // 0 separated (dist>0), 1 overlapping
  R dist;
  V3 pa, pb;   // witness points on the cores
  V3 dir;      // unit vector from B towards A (valid when separated)
  SVert simplex[4];
  int ns;
};

// reduce simplex to the feature closest to the origin; returns squared distance, writes v and barycentrics
static inline bool simplex_closest(SVert* s, int& n, V3& v, R* lam) {
  if (n == 1) {
    lam[0] = 1; v = s[0].w;
    return false;
  }
  if (n == 2) {
    V3 a = s[0].w, b = s[1].w, ab = b - a;
    R t = -dot(a, ab), dd = dot(ab, ab);
    if (t <= 0 || dd <= 0) { n = 1; lam[0] = 1; v = a; return false; }
    if (t >= dd) { s[0] = s[1]; n = 1; lam[0] = 1; v = b; return false; }
    t /= dd;
    lam[0] = 1 - t; lam[1] = t; v = a + ab * t;
    return false;
  }
  if (n == 3) {
    R bary[3];
    int mask = closest_tri(s[0].w, s[1].w, s[2].w, bary);
    SVert t[3];
    int m = 0;
    for (int i = 0; i < 3; i++)
      if (mask & (1 << i)) { t[m] = s[i]; lam[m] = bary[i]; m++; }
    for (int i = 0; i < m; i++) s[i] = t[i];
    n = m;
    v = V3(0, 0, 0);
    for (int i = 0; i < n; i++) v = v + s[i].w * lam[i];
    return false;
  }
  // tetrahedron: test the four faces whose outside contains the origin
  static const int F[4][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 3, 1}, {1, 2, 3, 0}};
  R bestd = 1e30;
  int bestf = -1;
  R bestb[3] = {0, 0, 0};
  int bestmask = 0;
  bool degenerate = false;
  for (int f = 0; f < 4; f++) {
    V3 a = s[F[f][0]].w, b = s[F[f][1]].w, c = s[F[f][2]].w, dpt = s[F[f][3]].w;
    V3 nrm = cross(b - a, c - a);
    R sd = dot(nrm, dpt - a), so = dot(nrm, -a);
    if (std::fabs(sd) < R(1e-20)) { degenerate = true; }
    // origin is outside this face if it is on the other side than the 4th vertex
    if (so * sd < 0 || std::fabs(sd) < R(1e-20)) {
      R bary[3];
      int mask = closest_tri(a, b, c, bary);
      V3 pt = a * bary[0] + b * bary[1] + c * bary[2];
      R d2 = dot(pt, pt);
      if (d2 < bestd) { bestd = d2; bestf = f; bestmask = mask; bestb[0] = bary[0]; bestb[1] = bary[1]; bestb[2] = bary[2]; }
    }
  }
  (void)degenerate;
  if (bestf < 0) return true;  // origin enclosed
  SVert t[3];
  int m = 0;
  for (int i = 0; i < 3; i++)
    if (bestmask & (1 << i)) { t[m] = s[F[bestf][i]]; lam[m] = bestb[i]; m++; }
  for (int i = 0; i < m; i++) s[i] = t[i];
  n = m;
  v = V3(0, 0, 0);
  for (int i = 0; i < n; i++) v = v + s[i].w * lam[i];
  return false;
}

static inline void gjk(const WShape& A, const WShape& B, GjkOut& o) {
  V3 v = A.X.p - B.X.p;
  if (dot(v, v) < R(1e-12)) v = V3(1, 0, 0);
  SVert s[4];
  int n = 0;
  R lam[4] = {1, 0, 0, 0};
  s[0] = mink_support(A, B, v);
  n = 1;
  v = s[0].w;
  o.status = 0;
  for (int iter = 0; iter < 40; iter++) {
    R vv = dot(v, v);
    if (vv < R(1e-14)) { o.status = 1; break; }
    SVert w = mink_support(A, B, -v);
    R vw = dot(v, w.w);
    if (vv - vw <= R(1e-6) * vv) break;  // no more progress towards the origin
    bool dup = false;
    for (int i = 0; i < n; i++) {
      V3 e = s[i].w - w.w;
      if (dot(e, e) < R(1e-14)) dup = true;
    }
    if (dup) break;
    s[n++] = w;
    V3 nv;
    if (simplex_closest(s, n, nv, lam)) { o.status = 1; break; }
    if (dot(nv, nv) >= vv) { v = nv; break; }  // numerical stall
    v = nv;
  }
  o.ns = n;
  for (int i = 0; i < n; i++) o.simplex[i] = s[i];
  if (o.status == 0) {
    R d = norm(v);
    o.dist = d;
    o.pa = V3(0, 0, 0);
    o.pb = V3(0, 0, 0);
    for (int i = 0; i < n; i++) {
      o.pa = o.pa + s[i].a * lam[i];
      o.pb = o.pb + s[i].b * lam[i];
    }
    o.dir = d > 0 ? v * (R(1) / d) : V3(1, 0, 0);
  }
}

#define EPA_MAXV 40
#define EPA_MAXF 96
struct EpaFace {
  int i0, i1, i2;
  V3 n;
  R d;
  bool alive;
};

// returns true on success: depth >= 0, normal nBA (from B to A), witness points on the cores
static inline bool epa(const WShape& A, const WShape& B, const GjkOut& g, R& depth, V3& nBA, V3& pa, V3& pb) {
  if (g.ns < 4) return false;
  SVert V[EPA_MAXV];
  EpaFace F[EPA_MAXF];
  int nv = 4, nf = 0;
  for (int i = 0; i < 4; i++) V[i] = g.simplex[i];
  static const int T[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
  V3 cen = (V[0].w + V[1].w + V[2].w + V[3].w) * R(0.25);
  for (int f = 0; f < 4; f++) {
    EpaFace& fc = F[nf];
    fc.i0 = T[f][0]; fc.i1 = T[f][1]; fc.i2 = T[f][2];
    V3 nrm = cross(V[fc.i1].w - V[fc.i0].w, V[fc.i2].w - V[fc.i0].w);
    R l = norm(nrm);
    if (l < R(1e-14)) return false;
    nrm = nrm * (R(1) / l);
    if (dot(nrm, V[fc.i0].w - cen) < 0) {
      int t = fc.i1; fc.i1 = fc.i2; fc.i2 = t;
      nrm = -nrm;
    }
    fc.n = nrm;
    fc.d = dot(nrm, V[fc.i0].w);
    fc.alive = true;
    nf++;
  }
  int bestf = 0;
  for (int iter = 0; iter < 32; iter++) {
    bestf = -1;
    R bd = 1e30;
    for (int f = 0; f < nf; f++)
      if (F[f].alive && F[f].d < bd) { bd = F[f].d; bestf = f; }
    if (bestf < 0) return false;
    SVert w = mink_support(A, B, F[bestf].n);
    R dw = dot(F[bestf].n, w.w);
    if (dw - bd < R(1e-6) || nv >= EPA_MAXV) break;
    // remove faces visible from w, collect horizon
    int eA[EPA_MAXF], eB[EPA_MAXF];
    int ne = 0;
    for (int f = 0; f < nf; f++) {
      if (!F[f].alive) continue;
      if (dot(F[f].n, w.w - V[F[f].i0].w) > R(0)) {
        F[f].alive = false;
        int ed[3][2] = {{F[f].i0, F[f].i1}, {F[f].i1, F[f].i2}, {F[f].i2, F[f].i0}};
        for (int k = 0; k < 3; k++) {
          int found = -1;
          for (int e = 0; e < ne; e++)
            if (eA[e] == ed[k][1] && eB[e] == ed[k][0]) { found = e; break; }
          if (found >= 0) {
            eA[found] = eA[ne - 1]; eB[found] = eB[ne - 1];
            ne--;
          } else if (ne < EPA_MAXF) {
            eA[ne] = ed[k][0]; eB[ne] = ed[k][1];
            ne++;
          }
        }
      }
    }
    if (ne == 0) break;
    int wi = nv;
    V[nv++] = w;
    bool full = false;
    for (int e = 0; e < ne; e++) {
      int slot = -1;
      for (int f = 0; f < nf; f++)
        if (!F[f].alive) { slot = f; break; }
      if (slot < 0) {
        if (nf >= EPA_MAXF) { full = true; break; }
        slot = nf++;
      }
      EpaFace& fc = F[slot];
      fc.i0 = eA[e]; fc.i1 = eB[e]; fc.i2 = wi;
      V3 nrm = cross(V[fc.i1].w - V[fc.i0].w, V[fc.i2].w - V[fc.i0].w);
      R l = norm(nrm);
      if (l < R(1e-14)) { fc.alive = true; fc.n = F[bestf].n; fc.d = 1e30; continue; }
      fc.n = nrm * (R(1) / l);
      fc.d = dot(fc.n, V[fc.i0].w);
      fc.alive = true;
    }
    if (full) break;
  }
  bestf = -1;
  R bd = 1e30;
  for (int f = 0; f < nf; f++)
    if (F[f].alive && F[f].d < bd) { bd = F[f].d; bestf = f; }
  if (bestf < 0) return false;
  const EpaFace& fc = F[bestf];
  // barycentrics of the origin's projection on the face
  V3 proj = fc.n * fc.d;
  V3 a = V[fc.i0].w - proj, b = V[fc.i1].w - proj, c = V[fc.i2].w - proj;
  R bary[3];
  closest_tri(a, b, c, bary);
  pa = V[fc.i0].a * bary[0] + V[fc.i1].a * bary[1] + V[fc.i2].a * bary[2];
  pb = V[fc.i0].b * bary[0] + V[fc.i1].b * bary[1] + V[fc.i2].b * bary[2];
  depth = fc.d < 0 ? 0 : fc.d;
  nBA = -fc.n;
  return true;
}

static inline int collide_convex_generic(const WShape& A, const WShape& B, R margin, Contact* out) {
  GjkOut g;
  gjk(A, B, g);
  R rA = core_radius(A), rB = core_radius(B);
  if (g.status == 0) {
    R dist = g.dist - rA - rB;
    if (dist >= margin) return 0;
    V3 n = g.dir;
    if (g.dist < R(1e-9)) n = normalized(A.X.p - B.X.p);
    V3 sa = g.pa - n * rA, sb = g.pb + n * rB;
    out[0].p = (sa + sb) * R(0.5);
    out[0].n = n;
    out[0].sep = dist;
    return 1;
  }
  R depth;
  V3 n, pa, pb;
  if (!epa(A, B, g, depth, n, pa, pb)) {
    n = normalized(A.X.p - B.X.p);
    out[0].p = (A.X.p + B.X.p) * R(0.5);
    out[0].n = n;
    out[0].sep = -(rA + rB);
    return 1;
  }
  V3 sa = pa - n * rA, sb = pb + n * rB;
  out[0].p = (sa + sb) * R(0.5);
  out[0].n = n;
  out[0].sep = -depth - rA - rB;
  return 1;
}

// Convex mesh (vertex cloud H) against a box Bx: a multi-point patch around the single GJK / EPA contact.  PhysX accumulates up to
// four points of a pair over frames (persistent manifold); this path is stateless, so the patch is generated at once: every hull
// vertex within `margin` of the box's supporting plane along the contact normal, whose foot point lies on the box, is a candidate
// (separation measured along the normal); reduce4 keeps the deepest one and the spread of the near-contact set.  The normal and the
// first candidate are the GJK / EPA result, so a vertex-less contact (edge against edge) degrades to the single point.
static inline int hull_box_patch(const WShape& H, const WShape& Bx, V3 n_out, bool hull_is_a, const Contact& c0, R margin, Contact* out) {
  const V3 nbh = hull_is_a ? n_out : -n_out;  // from the box towards the hull
  const R hb[3] = {Bx.size.x, Bx.size.y, Bx.size.z};
  R smax = dot(Bx.X.p, nbh);
  for (int k = 0; k < 3; k++) smax += hb[k] * std::fabs(dot(Bx.Rm.col(k), nbh));
  V3 cand[65];
  R dist[65];
  int m = 0;
  cand[m] = c0.p; dist[m] = c0.sep; m++;
  const R tol = R(1e-3);
  for (int i = 0; i < H.nverts && m < 65; i++) {
    V3 l(H.verts[3 * i], H.verts[3 * i + 1], H.verts[3 * i + 2]);
    V3 vw = H.X.p + H.Rm * l;
    R sp = dot(vw, nbh) - smax;
    if (!(sp < margin)) continue;
    V3 q = vw - nbh * sp;  // foot point on the supporting plane
    V3 rel = q - Bx.X.p;
    bool inside = true;
    for (int k = 0; k < 3; k++)
      if (std::fabs(dot(Bx.Rm.col(k), rel)) > hb[k] + tol) inside = false;
    if (!inside) continue;
    V3 cp = vw - nbh * (sp * R(0.5));
    V3 dc = cp - c0.p;
    if (dot(dc, dc) < R(1e-6)) {  // the GJK / EPA point itself (within 1 mm): keep one of the two, the vertex
      cand[0] = cp; dist[0] = sp;
      continue;
    }
    cand[m] = cp;
    dist[m] = sp;
    m++;
  }
  int keep[4];
  int k = reduce4(m, cand, dist, keep);
  for (int i = 0; i < k; i++) {
    out[i].p = cand[keep[i]];
    out[i].n = n_out;
    out[i].sep = dist[keep[i]];
  }
  return k;
}

// ---- convex mesh against convex mesh: multi-point patch from the two support faces around the GJK / EPA normal (what PhysX's
// persistent contact manifold accumulates over frames).  support face = the hull's vertices within tol of its supporting plane;
// candidates = support-face vertices of one hull whose foot point lies inside the other hull's support-face polygon.
struct FacePolygon {
  std::vector<std::pair<R, R>> pts;  // counter-clockwise in the basis (t1, t2) around c
  V3 c, t1, t2;
};
static inline R support_face_points(const WShape& H, V3 dir, R tol, std::vector<V3>& pts) {
  R smax = R(-3.0e38);
  std::vector<V3> world(H.nverts);
  for (int i = 0; i < H.nverts; i++) {
    world[i] = H.X.p + H.Rm * V3(H.verts[3 * i], H.verts[3 * i + 1], H.verts[3 * i + 2]);
    smax = std::fmax(smax, dot(world[i], dir));
  }
  pts.clear();
  for (int i = 0; i < H.nverts && (int)pts.size() < 24; i++)
    if (smax - dot(world[i], dir) < tol) pts.push_back(world[i]);
  return smax;
}
static inline bool build_face_polygon(const std::vector<V3>& pts, V3 dir, FacePolygon& P) {
  P.pts.clear();
  const int n = (int)pts.size();
  if (n < 3) return false;
  V3 c(0, 0, 0);
  for (const V3& p : pts) c = c + p;
  c = c * (R(1) / n);
  V3 t1 = std::fabs(dir.x) < R(0.57735) ? normalized(cross(dir, V3(1, 0, 0))) : normalized(cross(dir, V3(0, 1, 0)));
  V3 t2 = cross(dir, t1);
  P.c = c; P.t1 = t1; P.t2 = t2;
  for (const V3& p : pts) P.pts.push_back({dot(p - c, t1), dot(p - c, t2)});
  auto half = [](const std::pair<R, R>& a) { return (a.second < 0 || (a.second == 0 && a.first < 0)) ? 1 : 0; };
  // the same insertion order as the device code (ties resolve identically)
  for (int i = 1; i < n; i++) {
    std::pair<R, R> cur = P.pts[i];
    int j = i - 1;
    while (j >= 0) {
      bool after = half(P.pts[j]) > half(cur) || (half(P.pts[j]) == half(cur) && P.pts[j].first * cur.second - P.pts[j].second * cur.first < 0);
      if (!after) break;
      P.pts[j + 1] = P.pts[j];
      j--;
    }
    P.pts[j + 1] = cur;
  }
  return true;
}
static inline bool polygon_contains(const FacePolygon& P, V3 q, R tol) {
  R qx = dot(q - P.c, P.t1), qy = dot(q - P.c, P.t2);
  const int n = (int)P.pts.size();
  for (int i = 0; i < n; i++) {
    const auto& a = P.pts[i];
    const auto& b = P.pts[(i + 1) % n];
    R ex = b.first - a.first, ey = b.second - a.second;
    R len = std::sqrt(ex * ex + ey * ey);
    if (len < R(1e-9)) continue;
    if (ex * (qy - a.second) - ey * (qx - a.first) < -tol * len) return false;
  }
  return true;
}
static inline int hull_hull_patch(const WShape& A, const WShape& B, V3 n_out, const Contact& c0, R margin, Contact* out) {
  const R tol = R(1e-3);
  std::vector<V3> fa, fb;
  const R sa = support_face_points(A, -n_out, tol, fa);
  const R sb = support_face_points(B, n_out, tol, fb);
  FacePolygon PA, PB;
  const bool okA = build_face_polygon(fa, n_out, PA), okB = build_face_polygon(fb, n_out, PB);
  if (!okA && !okB) return 1;
  if (!(-sa - sb < margin)) return 1;
  V3 cand[49];
  R dist[49];
  int m = 0;
  cand[m] = c0.p; dist[m] = c0.sep; m++;
  if (okB)
    for (const V3& v : fa) {
      R sp = dot(v, n_out) - sb;
      if (!(sp < margin)) continue;
      if (!polygon_contains(PB, v - n_out * sp, tol)) continue;
      V3 cp = v - n_out * (sp * R(0.5));
      if (dot(cp - c0.p, cp - c0.p) < R(1e-6)) { cand[0] = cp; dist[0] = sp; continue; }
      cand[m] = cp; dist[m] = sp; m++;
    }
  if (okA)
    for (const V3& v : fb) {
      R sp = -sa - dot(v, n_out);
      if (!(sp < margin)) continue;
      if (!polygon_contains(PA, v + n_out * sp, tol)) continue;
      V3 cp = v + n_out * (sp * R(0.5));
      if (dot(cp - c0.p, cp - c0.p) < R(1e-6)) { cand[0] = cp; dist[0] = sp; continue; }
      cand[m] = cp; dist[m] = sp; m++;
    }
  if (m == 1) return 1;
  int keep[4];
  int k = reduce4(m, cand, dist, keep);
  for (int i = 0; i < k; i++) { out[i].p = cand[keep[i]]; out[i].n = n_out; out[i].sep = dist[keep[i]]; }
  return k;
}

// dispatch; out normals point from shape b towards shape a
static inline int collide_pair(const WShape& a, const WShape& b, R margin, Contact* out) {
  if (a.type == SHAPE_PLANE && b.type == SHAPE_PLANE) return 0;
  if (b.type == SHAPE_PLANE) return collide_plane_any(a, b, margin, out);
  if (a.type == SHAPE_PLANE) {
    int k = collide_plane_any(b, a, margin, out);
    for (int i = 0; i < k; i++) out[i].n = -out[i].n;
    return k;
  }
  if (a.type == SHAPE_BOX && b.type == SHAPE_BOX) return collide_box_box(a, b, margin, out);
  int k = collide_convex_generic(a, b, margin, out);
  if (k == 1 && a.type == SHAPE_CONVEX && b.type == SHAPE_BOX) {
    Contact c0 = out[0];
    return hull_box_patch(a, b, c0.n, true, c0, margin, out);
  }
  if (k == 1 && a.type == SHAPE_BOX && b.type == SHAPE_CONVEX) {
    Contact c0 = out[0];
    return hull_box_patch(b, a, c0.n, false, c0, margin, out);
  }
  if (k == 1 && a.type == SHAPE_CONVEX && b.type == SHAPE_CONVEX) {
    Contact c0 = out[0];
    return hull_hull_patch(a, b, c0.n, c0, margin, out);
  }
  return k;
}
