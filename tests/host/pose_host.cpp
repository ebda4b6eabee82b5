// Host unit harness for csrc/pose_math.h (TEST INFRASTRUCTURE: compiled by tests/test_pose_host.py with g++, never
// shipped or loaded by the product).  Runs the same hypothesis / consensus / cheirality sequence as csrc/pose.cu,
// single-threaded, so the algebra can be checked against OpenCV on a machine without a GPU, and so the GPU tests
// have a same-seed prediction of the kernels' output.
#include "../../nano_vs_slam_b200/csrc/pose_math.h"
#include <stdint.h>
#include <string.h>

using namespace nvs_pose;

extern "C" int nvs_host_five_point(const double* p1, const double* p2, double* Es) {
  double E[10][9];
  const int n = five_point(p1, p2, E);
  memcpy(Es, E, sizeof(double) * 9 * n);
  return n;
}

extern "C" int nvs_host_real_roots(const double* p, int deg, double* roots) { return poly_real_roots<10>(p, deg, roots); }

// cur / ref: (n,2) float normalised image coordinates.  Returns the inlier count; E, R (row-major), t, mask out.
extern "C" int nvs_host_pose(const float* cur, const float* ref, int n, float thr, int iters, uint64_t seed, int pair,
                             float* E_out, float* R_out, float* t_out, uint8_t* mask, int refine) {
  if (n < 5) return -1;
  const float thr2 = thr * thr;
  const float inv = 1.0f / thr2;
  unsigned long long best = ~0ull;
  float Eb[9] = {0};
  for (int it = 0; it < iters; ++it) {
    int idx[5];
    sample5(seed, pair, it, n, idx);
    double p1[10], p2[10], Es[10][9];
    for (int k = 0; k < 5; ++k) {
      p1[2 * k] = cur[2 * idx[k]]; p1[2 * k + 1] = cur[2 * idx[k] + 1];
      p2[2 * k] = ref[2 * idx[k]]; p2[2 * k + 1] = ref[2 * idx[k] + 1];
    }
    const int nc = five_point(p1, p2, Es);
    for (int c = 0; c < nc; ++c) {
      float Ef[9];
      for (int e = 0; e < 9; ++e) Ef[e] = float(Es[c][e]);
      unsigned long long score = 0;
      for (int i = 0; i < n; ++i) {
        const float err = sampson_sq<float>(Ef, cur[2 * i], cur[2 * i + 1], ref[2 * i], ref[2 * i + 1]);
        const float q = err <= thr2 ? err * inv : 1.0f;
        score += (unsigned long long)(q * POSE_SCORE_ONE);
      }
      if (score < best) { best = score; memcpy(Eb, Ef, sizeof(Eb)); }
    }
  }
  if (best == ~0ull) return -2;
  int ninl = 0;
  for (int i = 0; i < n; ++i) {
    mask[i] = sampson_sq<float>(Eb, cur[2 * i], cur[2 * i + 1], ref[2 * i], ref[2 * i + 1]) <= thr2;
    ninl += mask[i];
  }
  double Ed[9], R1[9], R2[9], t[3];
  for (int e = 0; e < 9; ++e) { Ed[e] = Eb[e]; E_out[e] = Eb[e]; }
  if (!decompose_essential(Ed, R1, R2, t)) return -3;
  int good[4] = {0, 0, 0, 0};
  const double tn[3] = {-t[0], -t[1], -t[2]};
  for (int i = 0; i < n; ++i) {
    const double x1 = cur[2 * i], y1 = cur[2 * i + 1], x2 = ref[2 * i], y2 = ref[2 * i + 1];
    good[0] += in_front(R1, t, x1, y1, x2, y2, 50.0);
    good[1] += in_front(R2, t, x1, y1, x2, y2, 50.0);
    good[2] += in_front(R1, tn, x1, y1, x2, y2, 50.0);
    good[3] += in_front(R2, tn, x1, y1, x2, y2, 50.0);
  }
  int b = 0;
  for (int c = 1; c < 4; ++c)
    if (good[c] > good[b]) b = c;
  const double* R = (b & 1) ? R2 : R1;
  const double* tt = (b & 2) ? tn : t;
  double Rb[9], tb[3];
  for (int e = 0; e < 9; ++e) Rb[e] = R[e];
  for (int e = 0; e < 3; ++e) tb[e] = tt[e];
  if (refine > 0) {
    // Gauss-Newton on the consensus set, keeping the pose with the lowest truncated cost (same sequence as
    // pose_refine_kernel)
    double Rc[9], tc[3], best_cost = -1.0;
    for (int e = 0; e < 9; ++e) Rc[e] = Rb[e];
    for (int e = 0; e < 3; ++e) tc[e] = tb[e];
    const double thr2d = double(thr2);
    for (int iter = 0; iter <= refine; ++iter) {
      double Es[6][9], acc[POSE_NACC];
      for (int k = 0; k < POSE_NACC; ++k) acc[k] = 0.0;
      refine_stencil(Rc, tc, Es);
      for (int i = 0; i < n; ++i) refine_accumulate(Es, cur[2 * i], cur[2 * i + 1], ref[2 * i], ref[2 * i + 1], thr2d, acc);
      const double cost = acc[POSE_NACC - 1];
      if (best_cost >= 0.0 && !(cost < best_cost)) break;
      best_cost = cost;
      for (int e = 0; e < 9; ++e) Rb[e] = Rc[e];
      for (int e = 0; e < 3; ++e) tb[e] = tc[e];
      double d[POSE_NPAR], R2[9], t2[3];
      if (iter == refine || !refine_solve(acc, d)) break;
      perturb_pose(Rc, tc, d, R2, t2);
      for (int e = 0; e < 9; ++e) Rc[e] = R2[e];
      for (int e = 0; e < 3; ++e) tc[e] = t2[e];
    }
    double Ed2[9];
    essential_from_pose(Rb, tb, Ed2);
    double nrm = 0.0;
    for (int e = 0; e < 9; ++e) nrm += Ed2[e] * Ed2[e];
    nrm = sqrt(2.0 / nrm);
    ninl = 0;
    for (int e = 0; e < 9; ++e) { Eb[e] = float(Ed2[e] * nrm); E_out[e] = Eb[e]; }
    for (int i = 0; i < n; ++i) {
      mask[i] = sampson_sq<float>(Eb, cur[2 * i], cur[2 * i + 1], ref[2 * i], ref[2 * i + 1]) <= thr2;
      ninl += mask[i];
    }
  }
  for (int e = 0; e < 9; ++e) R_out[e] = float(Rb[e]);
  for (int e = 0; e < 3; ++e) t_out[e] = float(tb[e]);
  return ninl;
}
