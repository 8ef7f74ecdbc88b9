// b200sim device math: float3-style vectors, row-major 3x3, wxyz quaternions.  All functions are __host__
// __device__ so the same per-env code can be exercised by the host-side emulation used in CPU tests
// (tests/emu); the product only ever launches it on the GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B2S_HD __host__ __device__ __forceinline__
#define B2S_HDN __host__ __device__
#else
#define B2S_HD inline
#define B2S_HDN
#endif

namespace b2s {

struct v3 {
  float x, y, z;
};
B2S_HD v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
B2S_HD v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
B2S_HD v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
B2S_HD v3 operator-(v3 a) { return mk3(-a.x, -a.y, -a.z); }
B2S_HD v3 operator*(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
B2S_HD v3 operator*(float s, v3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
B2S_HD float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B2S_HD v3 cross(v3 a, v3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
B2S_HD float norm(v3 a) { return sqrtf(dot(a, a)); }
B2S_HD v3 normalized(v3 a) {
  float n = norm(a);
  return n > 0.f ? a * (1.f / n) : mk3(1.f, 0.f, 0.f);
}
B2S_HD float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct m3 {
  float m[9];  // row major
};
B2S_HD v3 col(const m3& A, int j) { return mk3(A.m[j], A.m[3 + j], A.m[6 + j]); }
B2S_HD v3 row(const m3& A, int i) { return mk3(A.m[3 * i], A.m[3 * i + 1], A.m[3 * i + 2]); }
B2S_HD v3 mul(const m3& A, v3 v) { return mk3(dot(row(A, 0), v), dot(row(A, 1), v), dot(row(A, 2), v)); }
B2S_HD v3 tmul(const m3& A, v3 v) { return mk3(dot(col(A, 0), v), dot(col(A, 1), v), dot(col(A, 2), v)); }
B2S_HD m3 mul(const m3& A, const m3& B) {
  m3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
B2S_HD m3 transpose(const m3& A) {
  m3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * j + i];
  return C;
}
B2S_HD m3 sym6(float xx, float yy, float zz, float xy, float xz, float yz) {
  m3 A;
  A.m[0] = xx; A.m[4] = yy; A.m[8] = zz;
  A.m[1] = A.m[3] = xy;
  A.m[2] = A.m[6] = xz;
  A.m[5] = A.m[7] = yz;
  return A;
}
B2S_HD m3 inverse3(const m3& A) {
  m3 C;
  float a = A.m[0], b = A.m[1], c = A.m[2], d = A.m[3], e = A.m[4], f = A.m[5], g = A.m[6], h = A.m[7], i = A.m[8];
  float det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  float id = 1.f / det;
  C.m[0] = (e * i - f * h) * id; C.m[1] = (c * h - b * i) * id; C.m[2] = (b * f - c * e) * id;
  C.m[3] = (f * g - d * i) * id; C.m[4] = (a * i - c * g) * id; C.m[5] = (c * d - a * f) * id;
  C.m[6] = (d * h - e * g) * id; C.m[7] = (b * g - a * h) * id; C.m[8] = (a * e - b * d) * id;
  return C;
}

struct q4 {
  float w, x, y, z;
};
B2S_HD q4 mkq(float w, float x, float y, float z) { q4 q; q.w = w; q.x = x; q.y = y; q.z = z; return q; }
B2S_HD q4 qmul(q4 a, q4 b) {
  return mkq(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
             a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
B2S_HD q4 qnormalized(q4 q) {
  float s = 1.f / sqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return mkq(q.w * s, q.x * s, q.y * s, q.z * s);
}
B2S_HD m3 qmat(q4 q) {
  m3 A;
  float w = q.w, x = q.x, y = q.y, z = q.z;
  A.m[0] = 1 - 2 * (y * y + z * z); A.m[1] = 2 * (x * y - w * z); A.m[2] = 2 * (x * z + w * y);
  A.m[3] = 2 * (x * y + w * z); A.m[4] = 1 - 2 * (x * x + z * z); A.m[5] = 2 * (y * z - w * x);
  A.m[6] = 2 * (x * z - w * y); A.m[7] = 2 * (y * z + w * x); A.m[8] = 1 - 2 * (x * x + y * y);
  return A;
}
B2S_HD v3 qrot(q4 q, v3 v) {
  v3 u = mk3(q.x, q.y, q.z);
  v3 t = cross(u, v) * 2.f;
  return v + t * q.w + cross(u, t);
}
B2S_HD q4 qaxis_angle(v3 axis, float ang) {
  float h = ang * 0.5f;
  float s = sinf(h);
  return mkq(cosf(h), axis.x * s, axis.y * s, axis.z * s);
}
B2S_HD q4 qexp(v3 rv) {
  float th = norm(rv);
  if (th < 1e-12f) return qnormalized(mkq(1.f, rv.x * 0.5f, rv.y * 0.5f, rv.z * 0.5f));
  float s = sinf(th * 0.5f) / th;
  return mkq(cosf(th * 0.5f), rv.x * s, rv.y * s, rv.z * s);
}

struct pose {
  v3 p;
  q4 q;
};
B2S_HD pose pmul(const pose& a, const pose& b) {
  pose c;
  c.p = a.p + qrot(a.q, b.p);
  c.q = qmul(a.q, b.q);
  return c;
}
B2S_HD pose pose_ident() {
  pose P;
  P.p = mk3(0, 0, 0);
  P.q = mkq(1, 0, 0, 0);
  return P;
}
B2S_HD pose pose7(const float* f) {
  pose P;
  P.p = mk3(f[0], f[1], f[2]);
  P.q = mkq(f[3], f[4], f[5], f[6]);
  return P;
}

// 6-vectors: [angular(3); linear(3)] for motions, [moment(3); force(3)] for forces
struct v6 {
  v3 a, l;
};
B2S_HD v6 mk6(v3 a, v3 l) { v6 r; r.a = a; r.l = l; return r; }
B2S_HD v6 zero6() { return mk6(mk3(0, 0, 0), mk3(0, 0, 0)); }
B2S_HD v6 operator+(v6 x, v6 y) { return mk6(x.a + y.a, x.l + y.l); }
B2S_HD v6 operator-(v6 x, v6 y) { return mk6(x.a - y.a, x.l - y.l); }
B2S_HD v6 operator*(v6 x, float s) { return mk6(x.a * s, x.l * s); }
B2S_HD float dot6(v6 x, v6 y) { return dot(x.a, y.a) + dot(x.l, y.l); }
B2S_HD v6 crm(v6 v, v6 m) { return mk6(cross(v.a, m.a), cross(v.a, m.l) + cross(v.l, m.a)); }
B2S_HD v6 crf(v6 v, v6 f) { return mk6(cross(v.a, f.a) + cross(v.l, f.l), cross(v.a, f.l)); }
B2S_HD float get6(const v6& x, int i) { return i < 3 ? comp(x.a, i) : comp(x.l, i - 3); }

}  // namespace b2s
