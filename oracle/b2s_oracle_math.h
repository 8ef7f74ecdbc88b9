// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md). Never linked into the product.
// Small fixed-size linear algebra for the oracle, templated on the scalar through the REAL macro.
#pragma once
#include <cmath>
#include <cstdint>

#ifndef REAL
#define REAL double
#endif
typedef REAL R;

struct V3 {
  R x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(R a, R b, R c) : x(a), y(b), z(c) {}
  R operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  R& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, R s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(R s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
static inline R dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline R norm(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalized(V3 a) {
  R n = norm(a);
  return n > R(0) ? a * (R(1) / n) : V3(1, 0, 0);
}

struct M3 {  // row major
  R m[3][3];
  M3() {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) m[i][j] = 0;
  }
  V3 col(int j) const { return V3(m[0][j], m[1][j], m[2][j]); }
  V3 row(int i) const { return V3(m[i][0], m[i][1], m[i][2]); }
};
static inline V3 operator*(const M3& A, V3 v) { return V3(dot(A.row(0), v), dot(A.row(1), v), dot(A.row(2), v)); }
static inline V3 tmul(const M3& A, V3 v) { return V3(dot(A.col(0), v), dot(A.col(1), v), dot(A.col(2), v)); }  // A^T v
static inline M3 operator*(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
  return C;
}
static inline M3 transpose(const M3& A) {
  M3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A.m[j][i];
  return C;
}
static inline M3 operator+(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = A.m[i][j] + B.m[i][j];
  return C;
}
static inline M3 sym_from6(const R* s) {  // xx yy zz xy xz yz
  M3 A;
  A.m[0][0] = s[0]; A.m[1][1] = s[1]; A.m[2][2] = s[2];
  A.m[0][1] = A.m[1][0] = s[3];
  A.m[0][2] = A.m[2][0] = s[4];
  A.m[1][2] = A.m[2][1] = s[5];
  return A;
}
static inline M3 inverse_sym(const M3& A) {
  // general 3x3 inverse by cofactors
  M3 C;
  R a = A.m[0][0], b = A.m[0][1], c = A.m[0][2], d = A.m[1][0], e = A.m[1][1], f = A.m[1][2], g = A.m[2][0], h = A.m[2][1], i = A.m[2][2];
  R det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  R id = R(1) / det;
  C.m[0][0] = (e * i - f * h) * id; C.m[0][1] = (c * h - b * i) * id; C.m[0][2] = (b * f - c * e) * id;
  C.m[1][0] = (f * g - d * i) * id; C.m[1][1] = (a * i - c * g) * id; C.m[1][2] = (c * d - a * f) * id;
  C.m[2][0] = (d * h - e * g) * id; C.m[2][1] = (b * g - a * h) * id; C.m[2][2] = (a * e - b * d) * id;
  return C;
}

struct Q4 {  // w x y z
  R w, x, y, z;
  Q4() : w(1), x(0), y(0), z(0) {}
  Q4(R a, R b, R c, R d) : w(a), x(b), y(c), z(d) {}
};
static inline Q4 qmul(Q4 a, Q4 b) {
  return Q4(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
static inline Q4 qnormalized(Q4 q) {
  R n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  R s = R(1) / n;
  return Q4(q.w * s, q.x * s, q.y * s, q.z * s);
}
static inline M3 qmat(Q4 q) {
  M3 A;
  R w = q.w, x = q.x, y = q.y, z = q.z;
  A.m[0][0] = 1 - 2 * (y * y + z * z); A.m[0][1] = 2 * (x * y - w * z); A.m[0][2] = 2 * (x * z + w * y);
  A.m[1][0] = 2 * (x * y + w * z); A.m[1][1] = 1 - 2 * (x * x + z * z); A.m[1][2] = 2 * (y * z - w * x);
  A.m[2][0] = 2 * (x * z - w * y); A.m[2][1] = 2 * (y * z + w * x); A.m[2][2] = 1 - 2 * (x * x + y * y);
  return A;
}
static inline V3 qrot(Q4 q, V3 v) {
  // v + 2 w (u x v) + 2 u x (u x v)
  V3 u(q.x, q.y, q.z);
  V3 t = cross(u, v) * R(2);
  return v + t * q.w + cross(u, t);
}
static inline Q4 qaxis_angle(V3 axis, R ang) {
  R h = ang * R(0.5);
  R s = std::sin(h);
  return Q4(std::cos(h), axis.x * s, axis.y * s, axis.z * s);
}
// rotation-vector exponential (used to apply the accumulated TGS angular displacement)
static inline Q4 qexp(V3 rv) {
  R th = norm(rv);
  if (th < R(1e-12)) return qnormalized(Q4(1, rv.x * R(0.5), rv.y * R(0.5), rv.z * R(0.5)));
  R s = std::sin(th * R(0.5)) / th;
  return Q4(std::cos(th * R(0.5)), rv.x * s, rv.y * s, rv.z * s);
}

struct Pose {
  V3 p;
  Q4 q;
};
static inline Pose pmul(const Pose& a, const Pose& b) {
  Pose c;
  c.p = a.p + qrot(a.q, b.p);
  c.q = qmul(a.q, b.q);
  return c;
}
static inline Pose pose_from7(const float* f) {
  Pose P;
  P.p = V3(f[0], f[1], f[2]);
  P.q = Q4(f[3], f[4], f[5], f[6]);
  return P;
}
