// Register-resident small linear algebra for the per-instance solve kernels.
// Semantics follow /root/reference/src/smplfitter/pt/rotation.py (cited per function); the
// algorithms are our own (one-sided Jacobi SVD instead of a library SVD, packed Cholesky).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace sf {

// ---- 3x3 helpers (row-major float[9]) -------------------------------------------------
__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
// C = A^T B
__device__ __forceinline__ void mat3_tmul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
__device__ __forceinline__ void mat3_vec(const float* A, const float* x, float* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = A[i * 3] * x[0] + A[i * 3 + 1] * x[1] + A[i * 3 + 2] * x[2];
}

__device__ __forceinline__ float div_no_nan(float a, float b) {  // pt/rotation.py:8-11
  return b == 0.f ? 0.f : a / b;
}

// Rodrigues formula with the element grouping of pt/rotation.py:236-258.
__device__ __forceinline__ void rotvec2mat(const float* rv, float* m) {
  const float angle = sqrtf(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  const float ax = div_no_nan(rv[0], angle), ay = div_no_nan(rv[1], angle), az = div_no_nan(rv[2], angle);
  float s, c;
  sincosf(angle, &s, &c);
  const float sx = s * ax, sy = s * ay, sz = s * az;
  const float c1x = (1.f - c) * ax, c1y = (1.f - c) * ay, c1z = (1.f - c) * az;
  float t = c1x * ay;
  m[1] = t - sz;
  m[3] = t + sz;
  t = c1x * az;
  m[2] = t + sy;
  m[6] = t - sy;
  t = c1y * az;
  m[5] = t - sx;
  m[7] = t + sx;
  m[0] = c1x * ax + c;
  m[4] = c1y * ay + c;
  m[8] = c1z * az + c;
}

// Rotation matrix -> rotation vector through the quaternion with the branch order of
// pt/rotation.py:261-289 (trace > 0, then r00 largest, then r11 > r22).
__device__ __forceinline__ void mat2rotvec(const float* r, float* rv) {
  const float r00 = r[0], r01 = r[1], r02 = r[2], r10 = r[3], r11 = r[4], r12 = r[5], r20 = r[6],
              r21 = r[7], r22 = r[8];
  const float trace = r00 + r11 + r22;
  float x, y, z, w;
  if (trace > 0.f) {
    x = r21 - r12; y = r02 - r20; z = r10 - r01; w = 1.f + trace;
  } else if (r00 > r11 && r00 > r22) {
    x = (1.f - r22) + (r00 - r11); y = r10 + r01; z = r02 + r20; w = r21 - r12;
  } else if (r11 > r22) {
    x = r10 + r01; y = (1.f - r22) - (r00 - r11); z = r21 + r12; w = r02 - r20;
  } else {
    x = r02 + r20; y = r21 + r12; z = (1.f + r22) - (r00 + r11); w = r10 - r01;
  }
  const float n = sqrtf(x * x + y * y + z * z);
  const float k = div_no_nan(2.f, n) * atan2f(n, w);
  rv[0] = k * x; rv[1] = k * y; rv[2] = k * z;
}

// Rotation taking unit vector a onto unit vector b (pt/rotation.py:210-224).
__device__ __forceinline__ void align_unit_vectors(const float* a, const float* b, float* R) {
  const float cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
  const float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
  const float s = sqrtf(cx * cx + cy * cy + cz * cz);
  const float ang = atan2f(s, dot);
  float rv[3] = {div_no_nan(cx * ang, s), div_no_nan(cy * ang, s), div_no_nan(cz * ang, s)};
  rotvec2mat(rv, R);
}

// Closest rotation to A (Frobenius): R = U diag(1,1,det(U V^T)) V^T, i.e. the SVD projection
// with the *last* (smallest) singular direction flipped on reflections (pt/rotation.py:100-110).
// One-sided (Hestenes) Jacobi in double, entirely in registers: rotate the columns of G = A V
// until mutually orthogonal, order them by norm, complete both bases right-handed with cross
// products (which bakes in the reflection fix).
__device__ inline void proj_so3(const float* Af, float* Rf) {
  double g[3][3], v[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      g[i][j] = (double)Af[i * 3 + j];
      v[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 12; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      const double al = g[0][p] * g[0][p] + g[1][p] * g[1][p] + g[2][p] * g[2][p];
      const double be = g[0][q] * g[0][q] + g[1][q] * g[1][q] + g[2][q] * g[2][q];
      const double ga = g[0][p] * g[0][q] + g[1][p] * g[1][q] + g[2][p] * g[2][q];
      if (fabs(ga) > 1e-15 * sqrt(al * be) && ga != 0.0) {
        rotated = true;
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double gp = g[i][p], gq = g[i][q];
          g[i][p] = c * gp - s * gq;
          g[i][q] = s * gp + c * gq;
          const double vp = v[i][p], vq = v[i][q];
          v[i][p] = c * vp - s * vq;
          v[i][q] = s * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
  double n[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) n[j] = g[0][j] * g[0][j] + g[1][j] * g[1][j] + g[2][j] * g[2][j];
  // indices of the two largest columns (i0 >= i1), the smallest is completed by cross products
  int i0 = 0;
  if (n[1] > n[i0]) i0 = 1;
  if (n[2] > n[i0]) i0 = 2;
  int i1 = (i0 == 0) ? 1 : 0;
  const int i2c = 3 - i0 - i1;
  if (n[i2c] > n[i1]) i1 = i2c;
  double u1[3], u2[3], v1[3], v2[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    // dynamic column pick without local-memory indexing
    u1[i] = (i0 == 0) ? g[i][0] : (i0 == 1) ? g[i][1] : g[i][2];
    u2[i] = (i1 == 0) ? g[i][0] : (i1 == 1) ? g[i][1] : g[i][2];
    v1[i] = (i0 == 0) ? v[i][0] : (i0 == 1) ? v[i][1] : v[i][2];
    v2[i] = (i1 == 0) ? v[i][0] : (i1 == 1) ? v[i][1] : v[i][2];
  }
  const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  if (!(n1 > 1e-300)) {  // A == 0 (or NaN): identity / NaN propagation
    const float nanv = (n1 == n1) ? 0.f : nanf("");
#pragma unroll
    for (int i = 0; i < 9; ++i) Rf[i] = ((i % 4) == 0 ? 1.f : 0.f) + nanv;
    return;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) u1[i] /= n1;
  const double d = u2[0] * u1[0] + u2[1] * u1[1] + u2[2] * u1[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) u2[i] -= d * u1[i];
  double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  if (n2 > 1e-14 * n1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) u2[i] /= n2;
  } else {  // rank one: any direction orthogonal to u1 (the projection is not unique there)
    const double ax = fabs(u1[0]), ay = fabs(u1[1]), az = fabs(u1[2]);
    double e[3] = {0, 0, 0};
    if (ax <= ay && ax <= az) e[0] = 1; else if (ay <= az) e[1] = 1; else e[2] = 1;
    u2[0] = u1[1] * e[2] - u1[2] * e[1];
    u2[1] = u1[2] * e[0] - u1[0] * e[2];
    u2[2] = u1[0] * e[1] - u1[1] * e[0];
    n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) u2[i] /= n2;
  }
  const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rf[i * 3 + j] = (float)(u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j]);
}

// In-place Cholesky solve of the SPD system (n <= NMAX) stored dense row-major in double.
// Mirrors torch.linalg.cholesky_ex + cholesky_solve (pt/bodyfitter.py:1083-1084): no pivoting,
// failures surface as NaN.
template <int NMAX>
__device__ inline void chol_solve(double (*G)[NMAX], double* rhs, int n) {
  for (int j = 0; j < n; ++j) {
    double d = G[j][j];
    for (int k = 0; k < j; ++k) d -= G[j][k] * G[j][k];
    d = sqrt(d);
    G[j][j] = d;
    const double inv = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double s = G[i][j];
      for (int k = 0; k < j; ++k) s -= G[i][k] * G[j][k];
      G[i][j] = s * inv;
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = rhs[i];
    for (int k = 0; k < i; ++k) s -= G[i][k] * rhs[k];
    rhs[i] = s / G[i][i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = rhs[i];
    for (int k = i + 1; k < n; ++k) s -= G[k][i] * rhs[k];
    rhs[i] = s / G[i][i];
  }
}

}  // namespace sf
