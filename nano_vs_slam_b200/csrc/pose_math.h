// Two-view pose math shared by the pose kernels (csrc/pose.cu) and the host unit harness of the tests
// (tests/host/pose_host.cpp compiles this header with g++ to check the algebra against OpenCV without a GPU).
//
// Replaces, for a batch of frame pairs, what the reference does per pair on the CPU with OpenCV
// (src/visual_odometry/visual_odometry.py:383-412): cv2.findEssentialMat(kpn_cur, kpn_ref, focal=1, pp=(0,0),
// RANSAC, prob=0.999, threshold=0.0003) followed by cv2.recoverPose.  OpenCV's source is not part of the reference
// tree; the algorithm restated here is the published one it implements: D. Nister's five-point solver (tenth-degree
// polynomial in z), Sampson-distance consensus, and the cheirality test over the four (R, t) decompositions.
// Conventions (checked against cv2 4.13 in tests/test_pose_host.py): points1 = current frame, points2 = reference
// frame, p2^T E p1 = 0, inlier iff squared Sampson distance <= threshold^2, result x2 ~ R x1 + t with |t| = 1.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define NVS_HD __host__ __device__ inline
#else
#define NVS_HD inline
#endif

namespace nvs_pose {

// ---- monomials of degree <= 3 in (x, y, z), Nister's column order ------------------------------------------------
// 0 x^3  1 y^3  2 x^2y  3 xy^2  4 x^2z  5 x^2  6 y^2z  7 y^2  8 xyz  9 xy | 10 xz^2 11 xz 12 x 13 yz^2 14 yz 15 y
// 16 z^3 17 z^2 18 z 19 1
NVS_HD int mono_idx(int a, int b, int c) {
  switch (a * 16 + b * 4 + c) {
    case 48: return 0;
    case 12: return 1;
    case 36: return 2;
    case 24: return 3;
    case 33: return 4;
    case 32: return 5;
    case 9: return 6;
    case 8: return 7;
    case 21: return 8;
    case 20: return 9;
    case 18: return 10;
    case 17: return 11;
    case 16: return 12;
    case 6: return 13;
    case 5: return 14;
    case 4: return 15;
    case 3: return 16;
    case 2: return 17;
    case 1: return 18;
    default: return 19;
  }
}

NVS_HD void mat3_mul(const double* a, const double* b, double* c) {  // c = a b
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
NVS_HD void mat3_mul_bt(const double* a, const double* b, double* c) {  // c = a b^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      c[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
NVS_HD double det3(const double* r0, const double* r1, const double* r2) {
  return r0[0] * (r1[1] * r2[2] - r1[2] * r2[1]) - r0[1] * (r1[0] * r2[2] - r1[2] * r2[0]) +
         r0[2] * (r1[0] * r2[1] - r1[1] * r2[0]);
}

// ---- null space of the 5x9 epipolar constraint matrix ------------------------------------------------------------
// rows q = [x2x1, x2y1, x2, y2x1, y2y1, y2, x1, y1, 1] (E row-major, p2^T E p1 = 0).  Gauss-Jordan with complete
// pivoting; basis[k][9], k = 0..3 (any basis of the null space works for the polynomial system).
NVS_HD bool null_space_5x9(const double* p1, const double* p2, double basis[4][9]) {
  double A[5][9];
  int perm[9];
  for (int j = 0; j < 9; ++j) perm[j] = j;
  for (int i = 0; i < 5; ++i) {
    const double x1 = p1[2 * i], y1 = p1[2 * i + 1], x2 = p2[2 * i], y2 = p2[2 * i + 1];
    A[i][0] = x2 * x1; A[i][1] = x2 * y1; A[i][2] = x2;
    A[i][3] = y2 * x1; A[i][4] = y2 * y1; A[i][5] = y2;
    A[i][6] = x1;      A[i][7] = y1;      A[i][8] = 1.0;
  }
  for (int k = 0; k < 5; ++k) {
    int pi = k, pj = k;
    double best = 0.0;
    for (int i = k; i < 5; ++i)
      for (int j = k; j < 9; ++j)
        if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pi = i; pj = j; }
    if (best < 1e-14) return false;
    for (int j = 0; j < 9; ++j) { double t = A[k][j]; A[k][j] = A[pi][j]; A[pi][j] = t; }
    for (int i = 0; i < 5; ++i) { double t = A[i][k]; A[i][k] = A[i][pj]; A[i][pj] = t; }
    { int t = perm[k]; perm[k] = perm[pj]; perm[pj] = t; }
    const double inv = 1.0 / A[k][k];
    for (int j = 0; j < 9; ++j) A[k][j] *= inv;
    for (int i = 0; i < 5; ++i) {
      if (i == k) continue;
      const double f = A[i][k];
      if (f != 0.0)
        for (int j = 0; j < 9; ++j) A[i][j] -= f * A[k][j];
    }
  }
  for (int f = 0; f < 4; ++f) {
    double v[9];
    for (int j = 0; j < 9; ++j) v[j] = 0.0;
    v[5 + f] = 1.0;
    for (int i = 0; i < 5; ++i) v[i] = -A[i][5 + f];
    double nrm = 0.0;
    for (int j = 0; j < 9; ++j) nrm += v[j] * v[j];
    nrm = 1.0 / sqrt(nrm);
    for (int j = 0; j < 9; ++j) basis[f][perm[j]] = v[j] * nrm;
  }
  return true;
}

// ---- real roots of a polynomial (coefficients low -> high) by derivative interlacing -----------------------------
NVS_HD double poly_eval(const double* p, int deg, double x) {
  double v = p[deg];
  for (int i = deg - 1; i >= 0; --i) v = v * x + p[i];
  return v;
}

template <int MAXDEG>
NVS_HD int poly_real_roots(const double* p_in, int deg, double* roots) {
  while (deg > 0 && fabs(p_in[deg]) < 1e-300) --deg;
  if (deg == 0) return 0;
  // der[k] = k-th derivative, degree deg-k, stored consecutively
  double der[(MAXDEG + 1) * (MAXDEG + 2) / 2];
  int off[MAXDEG + 1];
  int o = 0;
  for (int k = 0; k < deg; ++k) { off[k] = o; o += deg - k + 1; }
  for (int i = 0; i <= deg; ++i) der[i] = p_in[i];
  for (int k = 1; k < deg; ++k)
    for (int i = 0; i <= deg - k; ++i) der[off[k] + i] = der[off[k - 1] + i + 1] * (i + 1);
  double prev[MAXDEG + 2], cur[MAXDEG + 2];
  int nprev = 0;
  for (int k = deg - 1; k >= 0; --k) {  // polynomial of degree d = deg-k
    const double* q = der + off[k];
    const int d = deg - k;
    double bound = 0.0;
    for (int i = 0; i < d; ++i) { double r = fabs(q[i] / q[d]); if (r > bound) bound = r; }
    bound += 1.0;
    int ncur = 0;
    double lo = -bound;
    double flo = poly_eval(q, d, lo);
    for (int s = 0; s <= nprev; ++s) {
      double hi = (s < nprev) ? prev[s] : bound;
      if (hi > bound) hi = bound;
      if (hi < lo) hi = lo;
      double fhi = poly_eval(q, d, hi);
      if (flo == 0.0) {
        cur[ncur++] = lo;
      } else if ((flo < 0.0) != (fhi < 0.0) && fhi != 0.0) {
        double a = lo, b = hi, fa = flo;
        for (int it = 0; it < 100; ++it) {
          const double m = 0.5 * (a + b);
          if (m == a || m == b) break;
          const double fm = poly_eval(q, d, m);
          if (fm == 0.0) { a = b = m; break; }
          if ((fm < 0.0) == (fa < 0.0)) { a = m; fa = fm; } else { b = m; }
        }
        cur[ncur++] = 0.5 * (a + b);
      }
      lo = hi;
      flo = fhi;
    }
    if (flo == 0.0 && (ncur == 0 || cur[ncur - 1] != lo)) cur[ncur++] = lo;
    nprev = ncur;
    for (int i = 0; i < ncur; ++i) prev[i] = cur[i];
  }
  for (int i = 0; i < nprev; ++i) roots[i] = prev[i];
  return nprev;
}

// ---- five-point solver: up to 10 essential matrices (row-major, Frobenius norm sqrt(2)) --------------------------
NVS_HD int five_point(const double* p1, const double* p2, double Es[10][9]) {
  double Bs[4][9];  // E = x B0 + y B1 + z B2 + B3
  if (!null_space_5x9(p1, p2, Bs)) return 0;
  double M[10][20];
  for (int r = 0; r < 10; ++r)
    for (int c = 0; c < 20; ++c) M[r][c] = 0.0;
  // det(E) and 2 E E^T E - tr(E E^T) E are trilinear in the coefficient vector (x, y, z, 1)
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double BBt[9];
      mat3_mul_bt(Bs[i], Bs[j], BBt);
      const double tr = BBt[0] + BBt[4] + BBt[8];
      for (int k = 0; k < 4; ++k) {
        const int a = (i == 0) + (j == 0) + (k == 0), b = (i == 1) + (j == 1) + (k == 1),
                  c = (i == 2) + (j == 2) + (k == 2);
        const int col = mono_idx(a, b, c);
        double G[9];
        mat3_mul(BBt, Bs[k], G);
        for (int e = 0; e < 9; ++e) M[e][col] += 2.0 * G[e] - tr * Bs[k][e];
        M[9][col] += det3(Bs[i], Bs[j] + 3, Bs[k] + 6);
      }
    }
  // Gauss-Jordan on the ten cubic-and-mixed columns
  for (int k = 0; k < 10; ++k) {
    int piv = k;
    double best = fabs(M[k][k]);
    for (int i = k + 1; i < 10; ++i)
      if (fabs(M[i][k]) > best) { best = fabs(M[i][k]); piv = i; }
    if (best < 1e-14) return 0;
    if (piv != k)
      for (int j = 0; j < 20; ++j) { double t = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = t; }
    const double inv = 1.0 / M[k][k];
    for (int j = k; j < 20; ++j) M[k][j] *= inv;
    for (int i = 0; i < 10; ++i) {
      if (i == k) continue;
      const double f = M[i][k];
      if (f != 0.0)
        for (int j = k; j < 20; ++j) M[i][j] -= f * M[k][j];
    }
  }
  // rows 4..9 (x^2z, x^2, y^2z, y^2, xyz, xy): <k> = <e> - z<f> etc. -> B(z) [x y 1]^T = 0
  double bx[3][4], by[3][4], bc[3][5];
  for (int r = 0; r < 3; ++r) {
    const double* e = &M[4 + 2 * r][10];
    const double* f = &M[5 + 2 * r][10];
    bx[r][3] = -f[0]; bx[r][2] = e[0] - f[1]; bx[r][1] = e[1] - f[2]; bx[r][0] = e[2];
    by[r][3] = -f[3]; by[r][2] = e[3] - f[4]; by[r][1] = e[4] - f[5]; by[r][0] = e[5];
    bc[r][4] = -f[6]; bc[r][3] = e[6] - f[7]; bc[r][2] = e[7] - f[8]; bc[r][1] = e[8] - f[9]; bc[r][0] = e[9];
  }
  double poly[11];
  for (int i = 0; i < 11; ++i) poly[i] = 0.0;
  // det B = sum over cyclic rows: bx[r0] * (by[r1] bc[r2] - bc[r1] by[r2])
  for (int r0 = 0; r0 < 3; ++r0) {
    const int r1 = (r0 + 1) % 3, r2 = (r0 + 2) % 3;
    double minor[8];
    for (int i = 0; i < 8; ++i) minor[i] = 0.0;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 5; ++j) minor[i + j] += by[r1][i] * bc[r2][j] - by[r2][i] * bc[r1][j];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 8; ++j) poly[i + j] += bx[r0][i] * minor[j];
  }
  double roots[12];
  const int nr = poly_real_roots<10>(poly, 10, roots);
  int n = 0;
  for (int r = 0; r < nr && n < 10; ++r) {
    const double z = roots[r];
    double Bz[3][3];
    for (int q = 0; q < 3; ++q) {
      Bz[q][0] = poly_eval(bx[q], 3, z);
      Bz[q][1] = poly_eval(by[q], 3, z);
      Bz[q][2] = poly_eval(bc[q], 4, z);
    }
    // null vector (x, y, 1): cross product of the best-conditioned pair of rows
    double bestw = 0.0, x = 0.0, y = 0.0;
    for (int q = 0; q < 3; ++q) {
      const double* u = Bz[q];
      const double* v = Bz[(q + 1) % 3];
      const double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cw = u[0] * v[1] - u[1] * v[0];
      if (fabs(cw) > fabs(bestw)) { bestw = cw; x = cx; y = cy; }
    }
    if (bestw == 0.0) continue;
    x /= bestw;
    y /= bestw;
    double nrm = 0.0;
    for (int e = 0; e < 9; ++e) {
      Es[n][e] = x * Bs[0][e] + y * Bs[1][e] + z * Bs[2][e] + Bs[3][e];
      nrm += Es[n][e] * Es[n][e];
    }
    if (!(nrm > 0.0) || !isfinite(nrm)) continue;
    nrm = sqrt(2.0 / nrm);
    for (int e = 0; e < 9; ++e) Es[n][e] *= nrm;
    ++n;
  }
  return n;
}

// one inlier-threshold unit of the integer consensus score (per-point cost = min(err / thr^2, 1) * POSE_SCORE_ONE)
#define POSE_SCORE_ONE 1073741824.0f

// ---- consensus -------------------------------------------------------------------------------------------------
template <typename T>
NVS_HD T sampson_sq(const T* E, T x1, T y1, T x2, T y2) {
  const T a0 = E[0] * x1 + E[1] * y1 + E[2], a1 = E[3] * x1 + E[4] * y1 + E[5], a2 = E[6] * x1 + E[7] * y1 + E[8];
  const T b0 = E[0] * x2 + E[3] * y2 + E[6], b1 = E[1] * x2 + E[4] * y2 + E[7];
  const T r = x2 * a0 + y2 * a1 + a2;
  return r * r / (a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1);
}

// counter-based generator (splitmix64 finaliser): sample s of pair p, draw d
NVS_HD uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
NVS_HD void sample5(uint64_t seed, int pair, int it, int n, int* idx) {
  uint64_t s = mix64(seed ^ (uint64_t(uint32_t(pair)) << 32) ^ uint64_t(uint32_t(it)));
  for (int k = 0; k < 5; ++k) {
    for (;;) {
      s = mix64(s);
      const int c = int((s >> 11) % uint64_t(n));
      bool dup = false;
      for (int j = 0; j < k; ++j) dup |= (idx[j] == c);
      if (!dup) { idx[k] = c; break; }
    }
  }
}

// ---- E -> four (R, t) candidates -----------------------------------------------------------------------------------
NVS_HD void jacobi_eig3(double A[3][3], double V[3][3]) {  // symmetric A -> eigenvalues on the diagonal, A = V D V^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double offd = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (offd < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

NVS_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// E = U diag(s, s, 0) V^T with det U = det V = +1; R1 = U W V^T, R2 = U W^T V^T, t = u3 (OpenCV decomposeEssentialMat)
NVS_HD bool decompose_essential(const double* E, double* R1, double* R2, double* t) {
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = E[i] * E[j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];  // E^T E
  jacobi_eig3(A, V);
  int o[3] = {0, 1, 2};  // eigenvalues descending
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (A[o[b]][o[b]] > A[o[a]][o[a]]) { int tt = o[a]; o[a] = o[b]; o[b] = tt; }
  double v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
  for (int k = 0; k < 3; ++k) { v1[k] = V[k][o[0]]; v2[k] = V[k][o[1]]; }
  cross3(v1, v2, v3);
  double n1 = 0, n2 = 0;
  for (int i = 0; i < 3; ++i) {
    u1[i] = E[3 * i] * v1[0] + E[3 * i + 1] * v1[1] + E[3 * i + 2] * v1[2];
    u2[i] = E[3 * i] * v2[0] + E[3 * i + 1] * v2[1] + E[3 * i + 2] * v2[2];
    n1 += u1[i] * u1[i];
  }
  if (!(n1 > 0.0)) return false;
  n1 = 1.0 / sqrt(n1);
  double d = 0;
  for (int i = 0; i < 3; ++i) { u1[i] *= n1; d += u1[i] * u2[i]; }
  for (int i = 0; i < 3; ++i) { u2[i] -= d * u1[i]; n2 += u2[i] * u2[i]; }
  if (!(n2 > 0.0)) return false;
  n2 = 1.0 / sqrt(n2);
  for (int i = 0; i < 3; ++i) u2[i] *= n2;
  cross3(u1, u2, u3);
  // U W V^T with W = [[0,1,0],[-1,0,0],[0,0,1]]: columns of U W = (-u2, u1, u3); of U W^T = (u2, -u1, u3)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      R1[3 * i + j] = -u2[i] * v1[j] + u1[i] * v2[j] + u3[i] * v3[j];
      R2[3 * i + j] = u2[i] * v1[j] - u1[i] * v2[j] + u3[i] * v3[j];
    }
  for (int i = 0; i < 3; ++i) t[i] = u3[i];
  return true;
}

// cheirality of one correspondence under x2 ~ R x1 + t: depths of the least-squares intersection of the two rays;
// in front of both cameras and closer than `far` (OpenCV recoverPose: distanceThresh = 50)
NVS_HD bool in_front(const double* R, const double* t, double x1, double y1, double x2, double y2, double far_) {
  const double a[3] = {R[0] * x1 + R[1] * y1 + R[2], R[3] * x1 + R[4] * y1 + R[5], R[6] * x1 + R[7] * y1 + R[8]};
  const double b[3] = {x2, y2, 1.0};
  // minimise |l2 b - l1 a - t|^2
  const double aa = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], bb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
  const double ab = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
  const double at = a[0] * t[0] + a[1] * t[1] + a[2] * t[2], bt = b[0] * t[0] + b[1] * t[1] + b[2] * t[2];
  const double den = aa * bb - ab * ab;
  if (!(fabs(den) > 1e-300)) return false;
  const double l1 = (ab * bt - bb * at) / den;
  const double l2 = (aa * bt - ab * at) / den;
  return l1 > 0.0 && l2 > 0.0 && l1 < far_ && l2 < far_;
}

// ---- local refinement on the essential manifold -------------------------------------------------------------------
// E = [t]x R with 5 degrees of freedom: R <- exp([w]x) R (w in R^3), t <- normalise(t + a b1 + b b2) with (b1, b2) an
// orthonormal basis of the tangent plane of the unit sphere at t.  Gauss-Newton on the signed Sampson distances of the
// current consensus set, Jacobian by forward differences (the kernels and the host harness run the same sequence).
constexpr int POSE_NPAR = 5;
constexpr int POSE_NACC = 1 + POSE_NPAR + POSE_NPAR * (POSE_NPAR + 1) / 2 + 1;  // cost, J^T r, upper J^T J, MSAC cost
#define POSE_FD_EPS 1e-6

NVS_HD void rodrigues_mul(const double* w, const double* R, double* out) {  // out = exp([w]x) R
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double a, b;  // exp = I + a K + b K^2, K = [w]x
  if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; } else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double K2[9], Ex[9];
  mat3_mul(K, K, K2);
  for (int e = 0; e < 9; ++e) Ex[e] = (e % 4 == 0 ? 1.0 : 0.0) + a * K[e] + b * K2[e];
  mat3_mul(Ex, R, out);
}

NVS_HD void essential_from_pose(const double* R, const double* t, double* E) {  // [t]x R
  const double T[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
  mat3_mul(T, R, E);
}

NVS_HD void tangent_basis(const double* t, double* b1, double* b2) {
  int k = 0;  // coordinate axis least aligned with t
  if (fabs(t[1]) < fabs(t[k])) k = 1;
  if (fabs(t[2]) < fabs(t[k])) k = 2;
  double ax[3] = {0, 0, 0};
  ax[k] = 1.0;
  cross3(t, ax, b1);
  const double n = 1.0 / sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
  for (int i = 0; i < 3; ++i) b1[i] *= n;
  cross3(t, b1, b2);
}

NVS_HD void perturb_pose(const double* R, const double* t, const double* d, double* R2, double* t2) {
  double b1[3], b2[3];
  tangent_basis(t, b1, b2);
  rodrigues_mul(d, R, R2);
  double n = 0.0;
  for (int i = 0; i < 3; ++i) { t2[i] = t[i] + d[3] * b1[i] + d[4] * b2[i]; n += t2[i] * t2[i]; }
  n = 1.0 / sqrt(n);
  for (int i = 0; i < 3; ++i) t2[i] *= n;
}

// base E and its five forward-difference neighbours, Es[6][9]
NVS_HD void refine_stencil(const double* R, const double* t, double Es[6][9]) {
  essential_from_pose(R, t, Es[0]);
  for (int k = 0; k < POSE_NPAR; ++k) {
    double d[POSE_NPAR] = {0, 0, 0, 0, 0}, R2[9], t2[3];
    d[k] = POSE_FD_EPS;
    perturb_pose(R, t, d, R2, t2);
    essential_from_pose(R2, t2, Es[1 + k]);
  }
}

NVS_HD double sampson_signed(const double* E, double x1, double y1, double x2, double y2) {
  const double a0 = E[0] * x1 + E[1] * y1 + E[2], a1 = E[3] * x1 + E[4] * y1 + E[5], a2 = E[6] * x1 + E[7] * y1 + E[8];
  const double b0 = E[0] * x2 + E[3] * y2 + E[6], b1 = E[1] * x2 + E[4] * y2 + E[7];
  return (x2 * a0 + y2 * a1 + a2) / sqrt(a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1);
}

// one correspondence's contribution to the normal equations (acc[POSE_NACC]); thr2 = squared inlier threshold
NVS_HD void refine_accumulate(const double Es[6][9], double x1, double y1, double x2, double y2, double thr2,
                              double* acc) {
  const double r0 = sampson_signed(Es[0], x1, y1, x2, y2);
  const double e = r0 * r0;
  acc[POSE_NACC - 1] += e <= thr2 ? e : thr2;
  if (!(e <= thr2)) return;
  double J[POSE_NPAR];
  for (int k = 0; k < POSE_NPAR; ++k) J[k] = (sampson_signed(Es[1 + k], x1, y1, x2, y2) - r0) * (1.0 / POSE_FD_EPS);
  acc[0] += e;
  int o = 1 + POSE_NPAR;
  for (int k = 0; k < POSE_NPAR; ++k) {
    acc[1 + k] += J[k] * r0;
    for (int l = k; l < POSE_NPAR; ++l) acc[o++] += J[k] * J[l];
  }
}

// solve (J^T J + lambda diag) d = -J^T r; false if singular
NVS_HD bool refine_solve(const double* acc, double* d) {
  double A[POSE_NPAR][POSE_NPAR + 1];
  int o = 1 + POSE_NPAR;
  for (int k = 0; k < POSE_NPAR; ++k)
    for (int l = k; l < POSE_NPAR; ++l) { A[k][l] = acc[o]; A[l][k] = acc[o]; ++o; }
  for (int k = 0; k < POSE_NPAR; ++k) {
    A[k][k] *= 1.0 + 1e-6;
    A[k][POSE_NPAR] = -acc[1 + k];
  }
  for (int k = 0; k < POSE_NPAR; ++k) {
    int piv = k;
    for (int i = k + 1; i < POSE_NPAR; ++i)
      if (fabs(A[i][k]) > fabs(A[piv][k])) piv = i;
    if (!(fabs(A[piv][k]) > 1e-300)) return false;
    if (piv != k)
      for (int j = 0; j <= POSE_NPAR; ++j) { double tt = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = tt; }
    for (int i = k + 1; i < POSE_NPAR; ++i) {
      const double f = A[i][k] / A[k][k];
      for (int j = k; j <= POSE_NPAR; ++j) A[i][j] -= f * A[k][j];
    }
  }
  for (int k = POSE_NPAR - 1; k >= 0; --k) {
    double v = A[k][POSE_NPAR];
    for (int j = k + 1; j < POSE_NPAR; ++j) v -= A[k][j] * d[j];
    d[k] = v / A[k][k];
  }
  for (int k = 0; k < POSE_NPAR; ++k)
    if (!isfinite(d[k])) return false;
  return true;
}

}  // namespace nvs_pose
