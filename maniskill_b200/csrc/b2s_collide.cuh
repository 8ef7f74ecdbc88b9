// b200sim narrowphase (device): plane/box/sphere/capsule/convex-hull contact generation for one shape pair, run by
// one lane per sub-scene.  Replaces the PCM narrowphase inside `px.step()` (reference call site
// mani_skill/envs/scene.py:379-380; configuration mani_skill/utils/structs/types.py:44-45).
//   box-box              SAT over 15 axes + Sutherland-Hodgman clipping of the incident face, <=4 points
//   plane-{box,hull,..}  vertex tests, <=4 points (deepest / farthest / max area / farthest-from-triangle)
//   everything else      GJK distance on the convex cores (+ sphere/capsule radius), EPA when the cores overlap
// Hull vertices are shared by all sub-scenes and read through the read-only path; all scratch is per-lane.
#pragma once
#include "b2s_math.cuh"

namespace b2s {

enum { SH_PLANE = 0, SH_BOX = 1, SH_SPHERE = 2, SH_CAPSULE = 3, SH_CONVEX = 4 };

struct WShape {
  int type;
  pose X;
  m3 R;
  v3 size;
  const float* verts;
  int nverts;
};

struct CPoint {
  v3 p, n;
  float sep;
};

// Keeps <= 4 of n candidate points: the deepest one, then the points that spread the support polygon most, where a candidate
// pays for lying above the deepest one: every length is reduced by (B2S_REDUCE4_LAMBDA x height above the deepest point), so the
// near-contact set is preferred over far speculative points (a many-vertex hull resting on a facet keeps that facet).
#define B2S_REDUCE4_LAMBDA 20
B2S_HD int reduce4(int n, const v3* p, const float* d, int* keep) {
  if (n <= 4) {
    for (int i = 0; i < n; i++) keep[i] = i;
    return n;
  }
  int i0 = 0;
  for (int i = 1; i < n; i++)
    if (d[i] < d[i0] - 1e-6f) i0 = i;  // ties (symmetric features) go to the lower index
  const float lam2 = (float)(B2S_REDUCE4_LAMBDA * B2S_REDUCE4_LAMBDA);
  int i1 = -1;
  float best = -1e30f;
  for (int i = 0; i < n; i++) {
    if (i == i0) continue;
    v3 e = p[i] - p[i0];
    float h = d[i] - d[i0];
    float v = dot(e, e) - lam2 * h * h;
    if (v > best + 1e-4f * fabsf(best) + 1e-12f) { best = v; i1 = i; }
  }
  int i2 = -1;
  best = -1e30f;
  const v3 base = p[i1] - p[i0];
  const float base2 = dot(base, base);
  for (int i = 0; i < n; i++) {
    if (i == i0 || i == i1) continue;
    v3 c = cross(base, p[i] - p[i0]);
    float h = d[i] - d[i0];
    float v = dot(c, c) - base2 * lam2 * h * h;  // (triangle height)^2 - (lambda h)^2, times base^2
    if (v > best + 1e-4f * fabsf(best) + 1e-12f) { best = v; i2 = i; }
  }
  int i3 = -1;
  best = -1e30f;
  for (int i = 0; i < n; i++) {
    if (i == i0 || i == i1 || i == i2) continue;
    v3 e0 = p[i] - p[i0], e1 = p[i] - p[i1], e2 = p[i] - p[i2];
    float h = d[i] - d[i0];
    float v = fminf(dot(e0, e0), fminf(dot(e1, e1), dot(e2, e2))) - lam2 * h * h;
    if (v > best + 1e-4f * fabsf(best) + 1e-12f) { best = v; i3 = i; }
  }
  keep[0] = i0; keep[1] = i1; keep[2] = i2; keep[3] = i3;
  return 4;
}

// streaming variant of "collide_plane_points": candidates are produced by a functor-free loop in the callers
B2S_HDN inline int plane_points(const WShape& P, int npts, const v3* pts, float radius, float margin, CPoint* out) {
  v3 n = col(P.R, 0);
  v3 cand[64];
  float dist[64];
  int m = 0;
  for (int i = 0; i < npts && m < 64; i++) {
    float d = dot(pts[i] - P.X.p, n) - radius;
    if (d < margin) {
      cand[m] = pts[i] - n * (radius + d * 0.5f);
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

B2S_HDN inline int collide_plane_any(const WShape& A, const WShape& P, float margin, CPoint* out) {
  v3 pts[64];
  if (A.type == SH_BOX) {
    for (int i = 0; i < 8; i++) {
      v3 l = mk3((i & 1) ? A.size.x : -A.size.x, (i & 2) ? A.size.y : -A.size.y, (i & 4) ? A.size.z : -A.size.z);
      pts[i] = A.X.p + mul(A.R, l);
    }
    return plane_points(P, 8, pts, 0.f, margin, out);
  } else if (A.type == SH_SPHERE) {
    pts[0] = A.X.p;
    return plane_points(P, 1, pts, A.size.x, margin, out);
  } else if (A.type == SH_CAPSULE) {
    v3 ax = col(A.R, 0) * A.size.y;
    pts[0] = A.X.p + ax;
    pts[1] = A.X.p - ax;
    return plane_points(P, 2, pts, A.size.x, margin, out);
  } else {
    int n = A.nverts < 64 ? A.nverts : 64;
    for (int i = 0; i < n; i++) pts[i] = A.X.p + mul(A.R, mk3(A.verts[3 * i], A.verts[3 * i + 1], A.verts[3 * i + 2]));
    return plane_points(P, n, pts, 0.f, margin, out);
  }
}

B2S_HD int clip_poly(int n, const v3* in, v3* out, int axis, float sign, float lim) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    v3 a = in[i], b = in[(i + 1) % n];
    float da = sign * comp(a, axis) - lim, db = sign * comp(b, axis) - lim;
    if (da <= 0.f) out[m++] = a;
    if ((da < 0.f && db > 0.f) || (da > 0.f && db < 0.f)) {
      float t = da / (da - db);
      out[m++] = a + (b - a) * t;
    }
  }
  return m;
}

B2S_HDN inline int collide_box_box(const WShape& A, const WShape& B, float margin, CPoint* out) {
  const m3& RA = A.R;
  const m3& RB = B.R;
  v3 d = B.X.p - A.X.p;
  v3 dA = tmul(RA, d), dB = tmul(RB, d);
  float C[3][3], AC[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      C[i][j] = dot(col(RA, i), col(RB, j));
      AC[i][j] = fabsf(C[i][j]);
    }
  float hA[3] = {A.size.x, A.size.y, A.size.z}, hB[3] = {B.size.x, B.size.y, B.size.z};
  float dAa[3] = {dA.x, dA.y, dA.z}, dBa[3] = {dB.x, dB.y, dB.z};
  float best_face = -1e30f;
  int face_code = -1;
  for (int i = 0; i < 3; i++) {
    float s = fabsf(dAa[i]) - (hA[i] + hB[0] * AC[i][0] + hB[1] * AC[i][1] + hB[2] * AC[i][2]);
    if (s > margin) return 0;
    if (s > best_face) { best_face = s; face_code = i; }
  }
  for (int j = 0; j < 3; j++) {
    float s = fabsf(dBa[j]) - (hB[j] + hA[0] * AC[0][j] + hA[1] * AC[1][j] + hA[2] * AC[2][j]);
    if (s > margin) return 0;
    if (s > best_face) { best_face = s; face_code = 3 + j; }
  }
  float best_edge = -1e30f;
  int ei = -1, ej = -1;
  v3 edgeL = mk3(0, 0, 0);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float len2 = 1.f - C[i][j] * C[i][j];
      if (len2 < 1e-6f) continue;
      float len = sqrtf(len2);
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      float rA = hA[i1] * AC[i2][j] + hA[i2] * AC[i1][j];
      float rB = hB[j1] * AC[i][j2] + hB[j2] * AC[i][j1];
      float dl = dAa[i2] * C[i1][j] - dAa[i1] * C[i2][j];
      float s = (fabsf(dl) - (rA + rB)) / len;
      if (s > margin) return 0;
      if (s > best_edge) {
        best_edge = s; ei = i; ej = j;
        v3 L = cross(col(RA, i), col(RB, j)) * (1.f / len);
        edgeL = (dot(L, d) < 0.f) ? -L : L;
      }
    }
  if (ei >= 0 && best_edge > best_face + 1e-4f) {
    v3 L = edgeL;
    v3 pA = A.X.p, pB = B.X.p;
    for (int k = 0; k < 3; k++) {
      if (k != ei) pA = pA + col(RA, k) * ((dot(col(RA, k), L) > 0.f ? hA[k] : -hA[k]));
      if (k != ej) pB = pB - col(RB, k) * ((dot(col(RB, k), L) > 0.f ? hB[k] : -hB[k]));
    }
    v3 ua = col(RA, ei), ub = col(RB, ej);
    v3 r = pB - pA;
    float uaub = dot(ua, ub), q1 = dot(ua, r), q2 = -dot(ub, r);
    float dd = 1.f - uaub * uaub;
    float s = 0.f, t = 0.f;
    if (dd > 1e-9f) {
      s = (q1 + uaub * q2) / dd;
      t = (uaub * q1 + q2) / dd;
    }
    s = fmaxf(-hA[ei], fminf(hA[ei], s));
    t = fmaxf(-hB[ej], fminf(hB[ej], t));
    v3 ca = pA + ua * s, cb = pB + ub * t;
    out[0].p = (ca + cb) * 0.5f;
    out[0].n = -L;
    out[0].sep = dot(cb - ca, L);
    return 1;
  }
  bool refA = face_code < 3;
  const WShape& Rf = refA ? A : B;
  const WShape& In = refA ? B : A;
  const m3& RR = Rf.R;
  const m3& RI = In.R;
  const float* hR = refA ? hA : hB;
  const float* hI = refA ? hB : hA;
  int ax = refA ? face_code : face_code - 3;
  v3 dRI = In.X.p - Rf.X.p;
  v3 nref = col(RR, ax);
  if (dot(nref, dRI) < 0.f) nref = -nref;
  int jx = 0;
  float bestd = -1.f;
  for (int j = 0; j < 3; j++) {
    float v = fabsf(dot(col(RI, j), nref));
    if (v > bestd) { bestd = v; jx = j; }
  }
  float sgn = dot(col(RI, jx), nref) > 0.f ? -1.f : 1.f;
  int j1 = (jx + 1) % 3, j2 = (jx + 2) % 3;
  v3 fc = In.X.p + col(RI, jx) * (sgn * hI[jx]);
  v3 e1 = col(RI, j1) * hI[j1], e2 = col(RI, j2) * hI[j2];
  v3 quad[4] = {fc + e1 + e2, fc - e1 + e2, fc - e1 - e2, fc + e1 - e2};
  int u1 = (ax + 1) % 3, u2 = (ax + 2) % 3;
  v3 U = col(RR, u1), V = col(RR, u2);
  v3 poly[16], tmp[16];
  for (int i = 0; i < 4; i++) {
    v3 rel = quad[i] - Rf.X.p;
    poly[i] = mk3(dot(rel, U), dot(rel, V), dot(rel, nref) - hR[ax]);
  }
  int n = 4;
  n = clip_poly(n, poly, tmp, 0, 1.f, hR[u1]);
  n = clip_poly(n, tmp, poly, 0, -1.f, hR[u1]);
  n = clip_poly(n, poly, tmp, 1, 1.f, hR[u2]);
  n = clip_poly(n, tmp, poly, 1, -1.f, hR[u2]);
  v3 cand[16];
  float dist[16];
  int m = 0;
  for (int i = 0; i < n; i++) {
    if (poly[i].z < margin) {
      cand[m] = Rf.X.p + U * poly[i].x + V * poly[i].y + nref * (hR[ax] + poly[i].z * 0.5f);
      dist[m] = poly[i].z;
      m++;
    }
  }
  int keep[4];
  int k = reduce4(m, cand, dist, keep);
  v3 nBA = refA ? -nref : nref;
  for (int i = 0; i < k; i++) {
    out[i].p = cand[keep[i]];
    out[i].n = nBA;
    out[i].sep = dist[keep[i]];
  }
  return k;
}

// ------------------------------------------------------------------------------------------------ GJK / EPA
B2S_HD v3 support_core(const WShape& S, v3 dir) {
  v3 dl = tmul(S.R, dir);
  v3 l;
  if (S.type == SH_BOX) {
    l = mk3(dl.x >= 0.f ? S.size.x : -S.size.x, dl.y >= 0.f ? S.size.y : -S.size.y, dl.z >= 0.f ? S.size.z : -S.size.z);
  } else if (S.type == SH_SPHERE) {
    l = mk3(0, 0, 0);
  } else if (S.type == SH_CAPSULE) {
    l = mk3(dl.x >= 0.f ? S.size.y : -S.size.y, 0, 0);
  } else {
    int bi = 0;
    float bd = -1e30f;
    for (int i = 0; i < S.nverts; i++) {
      float v = dl.x * S.verts[3 * i] + dl.y * S.verts[3 * i + 1] + dl.z * S.verts[3 * i + 2];
      if (v > bd) { bd = v; bi = i; }
    }
    l = mk3(S.verts[3 * bi], S.verts[3 * bi + 1], S.verts[3 * bi + 2]);
  }
  return S.X.p + mul(S.R, l);
}
B2S_HD float core_radius(const WShape& S) { return (S.type == SH_SPHERE || S.type == SH_CAPSULE) ? S.size.x : 0.f; }

struct SVert {
  v3 w, a, b;
};
B2S_HD SVert mink_support(const WShape& A, const WShape& B, v3 dir) {
  SVert s;
  s.a = support_core(A, dir);
  s.b = support_core(B, -dir);
  s.w = s.a - s.b;
  return s;
}

B2S_HD int closest_tri(v3 a, v3 b, v3 c, float* bary) {
  v3 ab = b - a, ac = c - a, ap = -a;
  float d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0.f && d2 <= 0.f) { bary[0] = 1; bary[1] = 0; bary[2] = 0; return 1; }
  v3 bp = -b;
  float d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0.f && d4 <= d3) { bary[0] = 0; bary[1] = 1; bary[2] = 0; return 2; }
  float vc = d1 * d4 - d3 * d2;
  if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
    float v = d1 / (d1 - d3);
    bary[0] = 1 - v; bary[1] = v; bary[2] = 0;
    return 3;
  }
  v3 cp = -c;
  float d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0.f && d5 <= d6) { bary[0] = 0; bary[1] = 0; bary[2] = 1; return 4; }
  float vb = d5 * d2 - d1 * d6;
  if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
    float w = d2 / (d2 - d6);
    bary[0] = 1 - w; bary[1] = 0; bary[2] = w;
    return 5;
  }
  float va = d3 * d6 - d5 * d4;
  if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
    float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    bary[0] = 0; bary[1] = 1 - w; bary[2] = w;
    return 6;
  }
  float denom = 1.f / (va + vb + vc);
  float v = vb * denom, w = vc * denom;
  bary[0] = 1 - v - w; bary[1] = v; bary[2] = w;
  return 7;
}

struct GjkOut {
  int status;
  float dist;
  v3 pa, pb, dir;
  SVert simplex[4];
  int ns;
};

B2S_HDN inline bool simplex_closest(SVert* s, int& n, v3& v, float* lam) {
  if (n == 1) {
    lam[0] = 1; v = s[0].w;
    return false;
  }
  if (n == 2) {
    v3 a = s[0].w, b = s[1].w, ab = b - a;
    float t = -dot(a, ab), dd = dot(ab, ab);
    if (t <= 0.f || dd <= 0.f) { n = 1; lam[0] = 1; v = a; return false; }
    if (t >= dd) { s[0] = s[1]; n = 1; lam[0] = 1; v = b; return false; }
    t /= dd;
    lam[0] = 1 - t; lam[1] = t; v = a + ab * t;
    return false;
  }
  if (n == 3) {
    float bary[3];
    int mask = closest_tri(s[0].w, s[1].w, s[2].w, bary);
    SVert t[3];
    int m = 0;
    for (int i = 0; i < 3; i++)
      if (mask & (1 << i)) { t[m] = s[i]; lam[m] = bary[i]; m++; }
    for (int i = 0; i < m; i++) s[i] = t[i];
    n = m;
    v = mk3(0, 0, 0);
    for (int i = 0; i < n; i++) v = v + s[i].w * lam[i];
    return false;
  }
  const int F[4][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 3, 1}, {1, 2, 3, 0}};
  float bestd = 1e30f;
  int bestf = -1;
  float bestb[3] = {0, 0, 0};
  int bestmask = 0;
  for (int f = 0; f < 4; f++) {
    v3 a = s[F[f][0]].w, b = s[F[f][1]].w, c = s[F[f][2]].w, dpt = s[F[f][3]].w;
    v3 nrm = cross(b - a, c - a);
    float sd = dot(nrm, dpt - a), so = dot(nrm, -a);
    if (so * sd < 0.f || fabsf(sd) < 1e-20f) {
      float bary[3];
      int mask = closest_tri(a, b, c, bary);
      v3 pt = a * bary[0] + b * bary[1] + c * bary[2];
      float d2 = dot(pt, pt);
      if (d2 < bestd) { bestd = d2; bestf = f; bestmask = mask; bestb[0] = bary[0]; bestb[1] = bary[1]; bestb[2] = bary[2]; }
    }
  }
  if (bestf < 0) return true;
  SVert t[3];
  int m = 0;
  for (int i = 0; i < 3; i++)
    if (bestmask & (1 << i)) { t[m] = s[F[bestf][i]]; lam[m] = bestb[i]; m++; }
  for (int i = 0; i < m; i++) s[i] = t[i];
  n = m;
  v = mk3(0, 0, 0);
  for (int i = 0; i < n; i++) v = v + s[i].w * lam[i];
  return false;
}

B2S_HDN inline void gjk(const WShape& A, const WShape& B, GjkOut& o) {
  v3 v = A.X.p - B.X.p;
  if (dot(v, v) < 1e-12f) v = mk3(1, 0, 0);
  SVert s[4];
  int n = 0;
  float lam[4] = {1, 0, 0, 0};
  s[0] = mink_support(A, B, v);
  n = 1;
  v = s[0].w;
  o.status = 0;
  for (int iter = 0; iter < 40; iter++) {
    float vv = dot(v, v);
    if (vv < 1e-14f) { o.status = 1; break; }
    SVert w = mink_support(A, B, -v);
    float vw = dot(v, w.w);
    if (vv - vw <= 1e-6f * vv) break;
    bool dup = false;
    for (int i = 0; i < n; i++) {
      v3 e = s[i].w - w.w;
      if (dot(e, e) < 1e-14f) dup = true;
    }
    if (dup) break;
    s[n++] = w;
    v3 nv;
    if (simplex_closest(s, n, nv, lam)) { o.status = 1; break; }
    if (dot(nv, nv) >= vv) { v = nv; break; }
    v = nv;
  }
  o.ns = n;
  for (int i = 0; i < n; i++) o.simplex[i] = s[i];
  if (o.status == 0) {
    float d = norm(v);
    o.dist = d;
    o.pa = mk3(0, 0, 0);
    o.pb = mk3(0, 0, 0);
    for (int i = 0; i < n; i++) {
      o.pa = o.pa + s[i].a * lam[i];
      o.pb = o.pb + s[i].b * lam[i];
    }
    o.dir = d > 0.f ? v * (1.f / d) : mk3(1, 0, 0);
  }
}

#define B2S_EPA_MAXV 40
#define B2S_EPA_MAXF 96

B2S_HDN inline bool epa(const WShape& A, const WShape& B, const GjkOut& g, float& depth, v3& nBA, v3& pa, v3& pb) {
  if (g.ns < 4) return false;
  SVert V[B2S_EPA_MAXV];
  int Fi0[B2S_EPA_MAXF], Fi1[B2S_EPA_MAXF], Fi2[B2S_EPA_MAXF];
  v3 Fn[B2S_EPA_MAXF];
  float Fd[B2S_EPA_MAXF];
  bool Fal[B2S_EPA_MAXF];
  int nv = 4, nf = 0;
  for (int i = 0; i < 4; i++) V[i] = g.simplex[i];
  const int T[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
  v3 cen = (V[0].w + V[1].w + V[2].w + V[3].w) * 0.25f;
  for (int f = 0; f < 4; f++) {
    int i0 = T[f][0], i1 = T[f][1], i2 = T[f][2];
    v3 nrm = cross(V[i1].w - V[i0].w, V[i2].w - V[i0].w);
    float l = norm(nrm);
    if (l < 1e-14f) return false;
    nrm = nrm * (1.f / l);
    if (dot(nrm, V[i0].w - cen) < 0.f) {
      int t = i1; i1 = i2; i2 = t;
      nrm = -nrm;
    }
    Fi0[nf] = i0; Fi1[nf] = i1; Fi2[nf] = i2;
    Fn[nf] = nrm;
    Fd[nf] = dot(nrm, V[i0].w);
    Fal[nf] = true;
    nf++;
  }
  int bestf = 0;
  for (int iter = 0; iter < 32; iter++) {
    bestf = -1;
    float bd = 1e30f;
    for (int f = 0; f < nf; f++)
      if (Fal[f] && Fd[f] < bd) { bd = Fd[f]; bestf = f; }
    if (bestf < 0) return false;
    SVert w = mink_support(A, B, Fn[bestf]);
    float dw = dot(Fn[bestf], w.w);
    if (dw - bd < 1e-6f || nv >= B2S_EPA_MAXV) break;
    int eA[B2S_EPA_MAXF], eB[B2S_EPA_MAXF];
    int ne = 0;
    for (int f = 0; f < nf; f++) {
      if (!Fal[f]) continue;
      if (dot(Fn[f], w.w - V[Fi0[f]].w) > 0.f) {
        Fal[f] = false;
        int ed[3][2] = {{Fi0[f], Fi1[f]}, {Fi1[f], Fi2[f]}, {Fi2[f], Fi0[f]}};
        for (int k = 0; k < 3; k++) {
          int found = -1;
          for (int e = 0; e < ne; e++)
            if (eA[e] == ed[k][1] && eB[e] == ed[k][0]) { found = e; break; }
          if (found >= 0) {
            eA[found] = eA[ne - 1]; eB[found] = eB[ne - 1];
            ne--;
          } else if (ne < B2S_EPA_MAXF) {
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
        if (!Fal[f]) { slot = f; break; }
      if (slot < 0) {
        if (nf >= B2S_EPA_MAXF) { full = true; break; }
        slot = nf++;
      }
      Fi0[slot] = eA[e]; Fi1[slot] = eB[e]; Fi2[slot] = wi;
      v3 nrm = cross(V[Fi1[slot]].w - V[Fi0[slot]].w, V[Fi2[slot]].w - V[Fi0[slot]].w);
      float l = norm(nrm);
      if (l < 1e-14f) { Fal[slot] = true; Fn[slot] = Fn[bestf]; Fd[slot] = 1e30f; continue; }
      Fn[slot] = nrm * (1.f / l);
      Fd[slot] = dot(Fn[slot], V[Fi0[slot]].w);
      Fal[slot] = true;
    }
    if (full) break;
  }
  bestf = -1;
  float bd = 1e30f;
  for (int f = 0; f < nf; f++)
    if (Fal[f] && Fd[f] < bd) { bd = Fd[f]; bestf = f; }
  if (bestf < 0) return false;
  v3 proj = Fn[bestf] * Fd[bestf];
  v3 a = V[Fi0[bestf]].w - proj, b = V[Fi1[bestf]].w - proj, c = V[Fi2[bestf]].w - proj;
  float bary[3];
  closest_tri(a, b, c, bary);
  pa = V[Fi0[bestf]].a * bary[0] + V[Fi1[bestf]].a * bary[1] + V[Fi2[bestf]].a * bary[2];
  pb = V[Fi0[bestf]].b * bary[0] + V[Fi1[bestf]].b * bary[1] + V[Fi2[bestf]].b * bary[2];
  depth = Fd[bestf] < 0.f ? 0.f : Fd[bestf];
  nBA = -Fn[bestf];
  return true;
}

B2S_HDN inline int collide_convex_generic(const WShape& A, const WShape& B, float margin, CPoint* out) {
  GjkOut g;
  gjk(A, B, g);
  float rA = core_radius(A), rB = core_radius(B);
  if (g.status == 0) {
    float dist = g.dist - rA - rB;
    if (dist >= margin) return 0;
    v3 n = g.dir;
    if (g.dist < 1e-9f) n = normalized(A.X.p - B.X.p);
    v3 sa = g.pa - n * rA, sb = g.pb + n * rB;
    out[0].p = (sa + sb) * 0.5f;
    out[0].n = n;
    out[0].sep = dist;
    return 1;
  }
  float depth;
  v3 n, pa, pb;
  if (!epa(A, B, g, depth, n, pa, pb)) {
    n = normalized(A.X.p - B.X.p);
    out[0].p = (A.X.p + B.X.p) * 0.5f;
    out[0].n = n;
    out[0].sep = -(rA + rB);
    return 1;
  }
  v3 sa = pa - n * rA, sb = pb + n * rB;
  out[0].p = (sa + sb) * 0.5f;
  out[0].n = n;
  out[0].sep = -depth - rA - rB;
  return 1;
}

// Convex mesh (vertex cloud H) against a box Bx: a multi-point patch around the single GJK / EPA contact.  PhysX accumulates up to
// four points of a pair over frames (persistent manifold); this path is stateless, so the patch is generated at once: every hull
// vertex within `margin` of the box's supporting plane along the contact normal, whose foot point lies on the box, is a candidate
// (separation measured along the normal); reduce4 keeps the deepest one and the spread of the near-contact set.  The normal and the
// first candidate are the GJK / EPA result, so a vertex-less contact (edge against edge) degrades to the single point.
B2S_HDN inline int hull_box_patch(const WShape& H, const WShape& Bx, v3 n_out, bool hull_is_a, const CPoint& c0, float margin, CPoint* out) {
  const v3 nbh = hull_is_a ? n_out : -n_out;  // from the box towards the hull
  const float hb[3] = {Bx.size.x, Bx.size.y, Bx.size.z};
  float smax = dot(Bx.X.p, nbh);
  for (int k = 0; k < 3; k++) smax += hb[k] * fabsf(dot(col(Bx.R, k), nbh));
  v3 cand[65];
  float dist[65];
  int m = 0;
  cand[m] = c0.p; dist[m] = c0.sep; m++;
  const float tol = 1e-3f;
  for (int i = 0; i < H.nverts && m < 65; i++) {
    v3 l = mk3(H.verts[3 * i], H.verts[3 * i + 1], H.verts[3 * i + 2]);
    v3 vw = H.X.p + mul(H.R, l);
    float sp = dot(vw, nbh) - smax;
    if (!(sp < margin)) continue;
    v3 q = vw - nbh * sp;  // foot point on the supporting plane
    v3 rel = q - Bx.X.p;
    bool inside = true;
    for (int k = 0; k < 3; k++)
      if (fabsf(dot(col(Bx.R, k), rel)) > hb[k] + tol) inside = false;
    if (!inside) continue;
    v3 cp = vw - nbh * (sp * 0.5f);
    v3 dc = cp - c0.p;
    if (dot(dc, dc) < 1e-6f) {  // the GJK / EPA point itself (within 1 mm): keep one of the two, the vertex
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

// ---- convex mesh against convex mesh: a multi-point patch where PhysX's persistent manifold would have accumulated one over frames
// (`enable_pcm`, mani_skill/utils/structs/types.py:50).  Around the GJK / EPA normal each hull has a SUPPORT FACE: its vertices within
// `tol` of its supporting plane.  A support-face vertex of one hull whose foot point lies inside the other hull's support-face polygon
// (both directions are tried) is a contact candidate, the separation measured along the normal; reduce4 keeps the deepest one and the
// spread.  Faces with fewer than three vertices (edge / vertex contacts) and crossing edges fall back to the single GJK / EPA point.
#define B2S_FACE_MAX 24
struct FacePoly {
  int n;
  v3 c, t1, t2;        // centroid, in-plane basis
  float x[B2S_FACE_MAX], y[B2S_FACE_MAX];  // vertices in the basis, counter-clockwise
};
// vertices of H within tol of its extreme along dir -> world points (at most B2S_FACE_MAX), returns the extreme value
B2S_HDN inline float support_face(const WShape& H, v3 dir, float tol, v3* pts, int& n) {
  float smax = -3.0e38f;
  for (int i = 0; i < H.nverts; i++) {
    v3 vw = H.X.p + mul(H.R, mk3(H.verts[3 * i], H.verts[3 * i + 1], H.verts[3 * i + 2]));
    smax = fmaxf(smax, dot(vw, dir));
  }
  n = 0;
  for (int i = 0; i < H.nverts && n < B2S_FACE_MAX; i++) {
    v3 vw = H.X.p + mul(H.R, mk3(H.verts[3 * i], H.verts[3 * i + 1], H.verts[3 * i + 2]));
    if (smax - dot(vw, dir) < tol) pts[n++] = vw;
  }
  return smax;
}
B2S_HDN inline bool face_polygon(const v3* pts, int n, v3 dir, FacePoly& P) {
  P.n = 0;
  if (n < 3) return false;
  v3 c = mk3(0, 0, 0);
  for (int i = 0; i < n; i++) c = c + pts[i];
  c = c * (1.f / n);
  v3 t1 = fabsf(dir.x) < 0.57735f ? normalized(cross(dir, mk3(1, 0, 0))) : normalized(cross(dir, mk3(0, 1, 0)));
  v3 t2 = cross(dir, t1);
  P.c = c; P.t1 = t1; P.t2 = t2; P.n = n;
  for (int i = 0; i < n; i++) { v3 d = pts[i] - c; P.x[i] = dot(d, t1); P.y[i] = dot(d, t2); }
  // insertion sort by angle around the centroid: lower half-plane after the upper one, inside a half-plane by the sign of the cross product
  for (int i = 1; i < n; i++) {
    float xi = P.x[i], yi = P.y[i];
    int hi_ = (yi < 0.f || (yi == 0.f && xi < 0.f)) ? 1 : 0;
    int j = i - 1;
    while (j >= 0) {
      int hj = (P.y[j] < 0.f || (P.y[j] == 0.f && P.x[j] < 0.f)) ? 1 : 0;
      bool after = hj > hi_ || (hj == hi_ && P.x[j] * yi - P.y[j] * xi < 0.f);  // j comes after i
      if (!after) break;
      P.x[j + 1] = P.x[j]; P.y[j + 1] = P.y[j];
      j--;
    }
    P.x[j + 1] = xi; P.y[j + 1] = yi;
  }
  return true;
}
B2S_HDN inline bool face_contains(const FacePoly& P, v3 q, float tol) {
  v3 d = q - P.c;
  float qx = dot(d, P.t1), qy = dot(d, P.t2);
  for (int i = 0; i < P.n; i++) {
    int k = i + 1 < P.n ? i + 1 : 0;
    float ex = P.x[k] - P.x[i], ey = P.y[k] - P.y[i];
    float len = sqrtf(ex * ex + ey * ey);
    if (len < 1e-9f) continue;
    if ((ex * (qy - P.y[i]) - ey * (qx - P.x[i])) < -tol * len) return false;
  }
  return true;
}
B2S_HDN inline int hull_hull_patch(const WShape& A, const WShape& B, v3 n_out, const CPoint& c0, float margin, CPoint* out) {
  // n_out points from B towards A: A's support face is its extreme along -n_out, B's along +n_out
  const float tol = 1e-3f;
  v3 fa[B2S_FACE_MAX], fb[B2S_FACE_MAX];
  int na = 0, nb = 0;
  const float sa = support_face(A, -n_out, tol, fa, na);   // max of dot(v, -n) over A  -> A's lowest extent along n is -sa
  const float sb = support_face(B, n_out, tol, fb, nb);    // B's highest extent along n
  FacePoly PA, PB;
  const bool okA = face_polygon(fa, na, n_out, PA), okB = face_polygon(fb, nb, n_out, PB);
  if (!okA && !okB) return 1;
  v3 cand[2 * B2S_FACE_MAX + 1];
  float dist[2 * B2S_FACE_MAX + 1];
  int m = 0;
  cand[m] = c0.p; dist[m] = c0.sep; m++;
  const float gap = -sa - sb;  // separation of the two supporting planes along n (negative when they overlap)
  if (!(gap < margin)) return 1;
  if (okB) {  // vertices of A's face over B's face polygon
    for (int i = 0; i < na; i++) {
      const float sp = dot(fa[i], n_out) - sb;
      if (!(sp < margin)) continue;
      const v3 q = fa[i] - n_out * sp;
      if (!face_contains(PB, q, tol)) continue;
      const v3 cp = fa[i] - n_out * (sp * 0.5f);
      const v3 dc = cp - c0.p;
      if (dot(dc, dc) < 1e-6f) { cand[0] = cp; dist[0] = sp; continue; }
      cand[m] = cp; dist[m] = sp; m++;
    }
  }
  if (okA) {  // vertices of B's face under A's face polygon
    for (int i = 0; i < nb; i++) {
      const float sp = -sa - dot(fb[i], n_out);
      if (!(sp < margin)) continue;
      const v3 q = fb[i] + n_out * sp;
      if (!face_contains(PA, q, tol)) continue;
      const v3 cp = fb[i] + n_out * (sp * 0.5f);
      const v3 dc = cp - c0.p;
      if (dot(dc, dc) < 1e-6f) { cand[0] = cp; dist[0] = sp; continue; }
      cand[m] = cp; dist[m] = sp; m++;
    }
  }
  if (m == 1) return 1;
  int keep[4];
  const int k = reduce4(m, cand, dist, keep);
  for (int i = 0; i < k; i++) { out[i].p = cand[keep[i]]; out[i].n = n_out; out[i].sep = dist[keep[i]]; }
  return k;
}

B2S_HDN inline int collide_pair(const WShape& a, const WShape& b, float margin, CPoint* out) {
  if (a.type == SH_PLANE && b.type == SH_PLANE) return 0;
  if (b.type == SH_PLANE) return collide_plane_any(a, b, margin, out);
  if (a.type == SH_PLANE) {
    int k = collide_plane_any(b, a, margin, out);
    for (int i = 0; i < k; i++) out[i].n = -out[i].n;
    return k;
  }
  if (a.type == SH_BOX && b.type == SH_BOX) return collide_box_box(a, b, margin, out);
  int k = collide_convex_generic(a, b, margin, out);
  if (k == 1 && a.type == SH_CONVEX && b.type == SH_BOX) {
    CPoint c0 = out[0];
    return hull_box_patch(a, b, c0.n, true, c0, margin, out);
  }
  if (k == 1 && a.type == SH_BOX && b.type == SH_CONVEX) {
    CPoint c0 = out[0];
    return hull_box_patch(b, a, c0.n, false, c0, margin, out);
  }
  if (k == 1 && a.type == SH_CONVEX && b.type == SH_CONVEX) {
    CPoint c0 = out[0];
    return hull_hull_patch(a, b, c0.n, c0, margin, out);
  }
  return k;
}

}  // namespace b2s
