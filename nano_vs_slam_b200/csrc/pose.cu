// Batched relative-pose step of the VO loop: for P frame pairs at once, what the reference does
// per pair on the CPU with OpenCV (src/visual_odometry/visual_odometry.py:383-412):
//   unproject matched keypoints with the pinhole intrinsics -> findEssentialMat(five-point, consensus on the Sampson
//   distance, threshold 0.0003) -> recoverPose (cheirality over the four decompositions) -> R, t, inlier mask.
// Inputs are exactly what nvs_select_keypoints + nvs_match_batch leave on the device, so nothing returns to the host
// between the network and the pose.  The algebra lives in pose_math.h (shared with the host unit harness).
//
// Work split (all sizes tiny next to the network: P x iters five-point solves, P x iters x 10 x n Sampson terms):
//   gather_kernel      matched pixel coordinates -> normalised (cur, ref) pairs, one thread per match
//   hypotheses_kernel  one thread per (pair, sample): 5 distinct matches (counter-based RNG), Nister five-point in
//                      fp64 -> up to 10 candidate E (fp32) in the workspace
//   score_kernel       one CTA per (pair, sample): truncated (MSAC) Sampson cost of its candidates over all matches;
//                      per-point costs are quantised to 2^-30 of the squared threshold and summed as integers, so the score
//                      does not depend on the reduction order (deterministic arg-min, lowest index wins ties)
//   select_kernel      one CTA per pair: arg-min, inlier mask, decomposition, cheirality vote, outputs
//   refine_kernel      (refine > 0) one CTA per pair: Gauss-Newton polish of (R, t) on the consensus set
#include "common.cuh"
#include "pose_math.h"

namespace nvs {

using namespace nvs_pose;

constexpr int POSE_MAX_CAND = 10;
constexpr int SCORE_T = 256;
constexpr int SELECT_T = 256;

struct PoseIn {
  const float* pts;        // (F, kmax, 2) pixel or normalised coordinates
  const int32_t* pair_a;   // (P) current frame of each pair
  const int32_t* pair_b;   // (P) reference frame
  const int32_t* idx1;     // (P, kmax) match m -> keypoint of frame a, or null = identity
  const int32_t* idx2;     // (P, kmax) match m -> keypoint of frame b, or null = identity
  const int32_t* count;    // (P) matches per pair
  int kmax;
  float fx, fy, cx, cy;
};

__global__ void __launch_bounds__(256) pose_gather_kernel(PoseIn in, float2* __restrict__ cur, float2* __restrict__ ref) {
  const int p = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  int n = in.count[p];
  n = n < 0 ? 0 : (n > in.kmax ? in.kmax : n);
  if (m >= n) return;
  const size_t o = (size_t)p * in.kmax + m;
  const int i1 = in.idx1 ? in.idx1[o] : m, i2 = in.idx2 ? in.idx2[o] : m;
  const float2 a = reinterpret_cast<const float2*>(in.pts)[(size_t)in.pair_a[p] * in.kmax + i1];
  const float2 b = reinterpret_cast<const float2*>(in.pts)[(size_t)in.pair_b[p] * in.kmax + i2];
  // camera.unproject_points: (u - cx) / fx, (v - cy) / fy (visual_odometry.py:386-387)
  cur[o] = make_float2((a.x - in.cx) / in.fx, (a.y - in.cy) / in.fy);
  ref[o] = make_float2((b.x - in.cx) / in.fx, (b.y - in.cy) / in.fy);
}

__global__ void __launch_bounds__(32) pose_hypotheses_kernel(const float2* __restrict__ cur, const float2* __restrict__ ref,
                                                             const int32_t* __restrict__ count, int kmax, int iters,
                                                             unsigned long long seed, float* __restrict__ cand,
                                                             int32_t* __restrict__ ncand, int it0, int it1,
                                                             const int32_t* __restrict__ done) {
  // samples [it0, it1) of every pair that has not reached its adaptive sample count yet (done == nullptr: all)
  const int p = blockIdx.y;
  const int it = it0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= it1 || (done != nullptr && done[p])) return;
  int n = count[p];
  n = n > kmax ? kmax : n;
  const size_t h = (size_t)p * iters + it;
  if (n < 5) {
    ncand[h] = 0;
    return;
  }
  int idx[5];
  sample5(seed, p, it, n, idx);
  double p1[10], p2[10], Es[10][9];
  for (int k = 0; k < 5; ++k) {
    const float2 a = cur[(size_t)p * kmax + idx[k]], b = ref[(size_t)p * kmax + idx[k]];
    p1[2 * k] = a.x; p1[2 * k + 1] = a.y;
    p2[2 * k] = b.x; p2[2 * k + 1] = b.y;
  }
  const int nc = five_point(p1, p2, Es);
  ncand[h] = nc;
  float* out = cand + h * POSE_MAX_CAND * 9;
  for (int c = 0; c < nc; ++c)
    for (int e = 0; e < 9; ++e) out[c * 9 + e] = (float)Es[c][e];
}

__global__ void __launch_bounds__(SCORE_T) pose_score_kernel(const float2* __restrict__ cur, const float2* __restrict__ ref,
                                                             const int32_t* __restrict__ count, int kmax, int iters,
                                                             float thr2, const float* __restrict__ cand,
                                                             const int32_t* __restrict__ ncand,
                                                             unsigned long long* __restrict__ score, int it0,
                                                             const int32_t* __restrict__ done) {
  __shared__ float Es[POSE_MAX_CAND * 9];
  __shared__ unsigned long long red[SCORE_T / 32][POSE_MAX_CAND];
  const int p = blockIdx.y, it = it0 + blockIdx.x;
  if (done != nullptr && done[p]) return;
  const size_t h = (size_t)p * iters + it;
  const int nc = ncand[h];
  unsigned long long* sc = score + h * POSE_MAX_CAND;
  if (nc == 0) {
    if (threadIdx.x < POSE_MAX_CAND) sc[threadIdx.x] = ~0ull;
    return;
  }
  int n = count[p];
  n = n > kmax ? kmax : n;
  if (threadIdx.x < nc * 9) Es[threadIdx.x] = cand[h * POSE_MAX_CAND * 9 + threadIdx.x];
  __syncthreads();
  unsigned long long acc[POSE_MAX_CAND];
#pragma unroll
  for (int c = 0; c < POSE_MAX_CAND; ++c) acc[c] = 0;
  const float inv = 1.0f / thr2;
  for (int i = threadIdx.x; i < n; i += SCORE_T) {
    const float2 a = cur[(size_t)p * kmax + i], b = ref[(size_t)p * kmax + i];
#pragma unroll
    for (int c = 0; c < POSE_MAX_CAND; ++c) {
      if (c < nc) {
        const float err = sampson_sq<float>(Es + 9 * c, a.x, a.y, b.x, b.y);
        const float q = err <= thr2 ? err * inv : 1.0f;
        acc[c] += (unsigned long long)(q * POSE_SCORE_ONE);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < POSE_MAX_CAND; ++c) {
    unsigned long long v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < POSE_MAX_CAND) {
    unsigned long long v = ~0ull;
    if ((int)threadIdx.x < nc) {
      v = 0;
      for (int w = 0; w < SCORE_T / 32; ++w) v += red[w][threadIdx.x];
    }
    sc[threadIdx.x] = v;
  }
}

// Adaptive stopping (cv2.findEssentialMat's prob argument, visual_odometry.py:392: prob = 0.999): after the samples
// [it0, it1) of a round have been scored, one CTA per pair updates the pair's best truncated cost; the inlier ratio is
// bounded from below by w = 1 - cost / n (an inlier contributes less than one unit), the standard RANSAC sample count for
// confidence c is N = log(1 - c) / log(1 - w^5), and the pair is finished once it1 >= N.  Finished pairs are skipped by
// the following rounds (their hypotheses / score CTAs exit at once); `used` = samples the select step has to scan.
__global__ void __launch_bounds__(128) pose_adapt_kernel(const int32_t* __restrict__ count, int kmax, int iters, int it0,
                                                         int it1, float confidence,
                                                         const unsigned long long* __restrict__ score,
                                                         unsigned long long* __restrict__ best, int32_t* __restrict__ done,
                                                         int32_t* __restrict__ used) {
  __shared__ unsigned long long s_min[128];
  const int p = blockIdx.x, tid = threadIdx.x;
  if (done[p]) return;
  unsigned long long m = ~0ull;
  for (int h = it0 * POSE_MAX_CAND + tid; h < it1 * POSE_MAX_CAND; h += 128) {
    const unsigned long long s = score[(size_t)p * iters * POSE_MAX_CAND + h];
    m = s < m ? s : m;
  }
  s_min[tid] = m;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o && s_min[tid + o] < s_min[tid]) s_min[tid] = s_min[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    unsigned long long b = best[p];
    if (s_min[0] < b) b = s_min[0];
    best[p] = b;
    used[p] = it1;
    int n = count[p];
    n = n > kmax ? kmax : n;
    if (n < 5) {
      done[p] = 1;  // nothing to estimate
    } else if (b != ~0ull) {
      double w = 1.0 - (double)b / ((double)n * (double)POSE_SCORE_ONE);
      w = w < 1e-3 ? 1e-3 : (w > 1.0 - 1e-9 ? 1.0 - 1e-9 : w);
      const double w5 = w * w * w * w * w;
      const double need = log(1.0 - (double)confidence) / log(1.0 - w5);
      if ((double)it1 >= need) done[p] = 1;
    }
  }
}

__global__ void __launch_bounds__(SELECT_T) pose_select_kernel(const float2* __restrict__ cur, const float2* __restrict__ ref,
                                                               const int32_t* __restrict__ count, int kmax, int iters,
                                                               float thr2, const float* __restrict__ cand,
                                                               const unsigned long long* __restrict__ score,
                                                               float* __restrict__ out_E, float* __restrict__ out_R,
                                                               float* __restrict__ out_t, uint8_t* __restrict__ out_mask,
                                                               int32_t* __restrict__ out_inliers,
                                                               const int32_t* __restrict__ used) {
  __shared__ unsigned long long s_best[SELECT_T];
  __shared__ int s_arg[SELECT_T];
  __shared__ float s_E[9];
  __shared__ double s_R[2][9], s_t[2][3];
  __shared__ int s_cnt[5];
  __shared__ int s_ok;
  const int p = blockIdx.x, tid = threadIdx.x;
  int n = count[p];
  n = n < 0 ? 0 : (n > kmax ? kmax : n);
  const int total = iters * POSE_MAX_CAND;  // row pitch of the score / candidate arrays
  const int scan = (used != nullptr ? used[p] : iters) * POSE_MAX_CAND;  // samples actually evaluated for this pair
  unsigned long long best = ~0ull;
  int arg = total;
  for (int h = tid; h < scan; h += SELECT_T) {
    const unsigned long long s = score[(size_t)p * total + h];
    if (s < best) { best = s; arg = h; }
  }
  s_best[tid] = best;
  s_arg[tid] = arg;
  __syncthreads();
  for (int o = SELECT_T / 2; o > 0; o >>= 1) {
    if (tid < o) {
      const unsigned long long b2 = s_best[tid + o];
      const int a2 = s_arg[tid + o];
      if (b2 < s_best[tid] || (b2 == s_best[tid] && a2 < s_arg[tid])) { s_best[tid] = b2; s_arg[tid] = a2; }
    }
    __syncthreads();
  }
  const bool have = s_best[0] != ~0ull;
  if (tid < 9) s_E[tid] = have ? cand[(size_t)p * total * 9 + (size_t)s_arg[0] * 9 + tid] : 0.f;
  if (tid < 5) s_cnt[tid] = 0;
  __syncthreads();
  if (tid == 0) {
    double Ed[9];
    for (int e = 0; e < 9; ++e) Ed[e] = s_E[e];
    s_ok = have && decompose_essential(Ed, s_R[0], s_R[1], s_t[0]);
    for (int e = 0; e < 3; ++e) s_t[1][e] = -s_t[0][e];
  }
  __syncthreads();
  const bool ok = s_ok != 0;
  int ninl = 0, g0 = 0, g1 = 0, g2 = 0, g3 = 0;
  for (int i = tid; i < kmax; i += SELECT_T) {
    uint8_t m = 0;
    if (i < n && ok) {
      const float2 a = cur[(size_t)p * kmax + i], b = ref[(size_t)p * kmax + i];
      m = sampson_sq<float>(s_E, a.x, a.y, b.x, b.y) <= thr2;
      ninl += m;
      // recoverPose is called without a mask in the reference (:404): every match votes
      g0 += in_front(s_R[0], s_t[0], a.x, a.y, b.x, b.y, 50.0);
      g1 += in_front(s_R[1], s_t[0], a.x, a.y, b.x, b.y, 50.0);
      g2 += in_front(s_R[0], s_t[1], a.x, a.y, b.x, b.y, 50.0);
      g3 += in_front(s_R[1], s_t[1], a.x, a.y, b.x, b.y, 50.0);
    }
    out_mask[(size_t)p * kmax + i] = m;
  }
  atomicAdd(&s_cnt[0], ninl);
  atomicAdd(&s_cnt[1], g0);
  atomicAdd(&s_cnt[2], g1);
  atomicAdd(&s_cnt[3], g2);
  atomicAdd(&s_cnt[4], g3);
  __syncthreads();
  if (tid == 0) {
    int b = 0;
    for (int c = 1; c < 4; ++c)
      if (s_cnt[1 + c] > s_cnt[1 + b]) b = c;
    out_inliers[p] = ok ? s_cnt[0] : 0;
    for (int e = 0; e < 9; ++e) {
      out_E[p * 9 + e] = ok ? s_E[e] : 0.f;
      out_R[p * 9 + e] = ok ? (float)s_R[b & 1][e] : (float)(e % 4 == 0);
    }
    for (int e = 0; e < 3; ++e) out_t[p * 3 + e] = ok ? (float)s_t[b >> 1][e] : 0.f;
  }
}

// Optional local optimisation (the role LO + final polishing play in OpenCV's USAC): Gauss-Newton on the essential
// manifold over the consensus set of the selected model, keeping the pose with the lowest truncated cost.  One CTA per
// pair; the 22 sums of the normal equations are reduced in fp64 (warp shuffles, then shared memory).
constexpr int REFINE_T = 256;

__global__ void __launch_bounds__(REFINE_T) pose_refine_kernel(const float2* __restrict__ cur, const float2* __restrict__ ref,
                                                               const int32_t* __restrict__ count, int kmax, float thr2,
                                                               int refine, float* __restrict__ out_E,
                                                               float* __restrict__ out_R, float* __restrict__ out_t,
                                                               uint8_t* __restrict__ out_mask,
                                                               int32_t* __restrict__ out_inliers) {
  __shared__ double s_Es[6][9];
  __shared__ double s_red[REFINE_T / 32][POSE_NACC];
  __shared__ double s_Rc[9], s_tc[3], s_Rb[9], s_tb[3];
  __shared__ double s_best;
  __shared__ int s_stop, s_cnt;
  __shared__ float s_Ef[9];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (out_inliers[p] <= 0) return;  // no model for this pair (block-uniform)
  int n = count[p];
  n = n > kmax ? kmax : n;
  if (tid == 0) {
    for (int e = 0; e < 9; ++e) s_Rc[e] = s_Rb[e] = out_R[p * 9 + e];
    for (int e = 0; e < 3; ++e) s_tc[e] = s_tb[e] = out_t[p * 3 + e];
    s_best = -1.0;
    s_stop = 0;
    s_cnt = 0;
  }
  __syncthreads();
  const double thr2d = (double)thr2;
  for (int iter = 0; iter <= refine; ++iter) {
    if (tid == 0) refine_stencil(s_Rc, s_tc, s_Es);
    __syncthreads();
    double acc[POSE_NACC];
#pragma unroll
    for (int k = 0; k < POSE_NACC; ++k) acc[k] = 0.0;
    for (int i = tid; i < n; i += REFINE_T) {
      const float2 a = cur[(size_t)p * kmax + i], b = ref[(size_t)p * kmax + i];
      refine_accumulate(s_Es, a.x, a.y, b.x, b.y, thr2d, acc);
    }
#pragma unroll
    for (int k = 0; k < POSE_NACC; ++k) {
      double v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double tot[POSE_NACC];
      for (int k = 0; k < POSE_NACC; ++k) {
        double v = 0.0;
        for (int w = 0; w < REFINE_T / 32; ++w) v += s_red[w][k];
        tot[k] = v;
      }
      const double cost = tot[POSE_NACC - 1];
      if (s_best >= 0.0 && !(cost < s_best)) {
        s_stop = 1;
      } else {
        s_best = cost;
        for (int e = 0; e < 9; ++e) s_Rb[e] = s_Rc[e];
        for (int e = 0; e < 3; ++e) s_tb[e] = s_tc[e];
        double d[POSE_NPAR], R2[9], t2[3];
        if (iter == refine || !refine_solve(tot, d)) {
          s_stop = 1;
        } else {
          perturb_pose(s_Rc, s_tc, d, R2, t2);
          for (int e = 0; e < 9; ++e) s_Rc[e] = R2[e];
          for (int e = 0; e < 3; ++e) s_tc[e] = t2[e];
        }
      }
    }
    __syncthreads();
    if (s_stop) break;
  }
  if (tid == 0) {
    double E[9], nrm = 0.0;
    essential_from_pose(s_Rb, s_tb, E);
    for (int e = 0; e < 9; ++e) nrm += E[e] * E[e];
    nrm = sqrt(2.0 / nrm);
    for (int e = 0; e < 9; ++e) {
      s_Ef[e] = (float)(E[e] * nrm);
      out_E[p * 9 + e] = s_Ef[e];
      out_R[p * 9 + e] = (float)s_Rb[e];
    }
    for (int e = 0; e < 3; ++e) out_t[p * 3 + e] = (float)s_tb[e];
  }
  __syncthreads();
  int ninl = 0;
  for (int i = tid; i < n; i += REFINE_T) {
    const float2 a = cur[(size_t)p * kmax + i], b = ref[(size_t)p * kmax + i];
    const uint8_t m = sampson_sq<float>(s_Ef, a.x, a.y, b.x, b.y) <= thr2;
    out_mask[(size_t)p * kmax + i] = m;
    ninl += m;
  }
  atomicAdd(&s_cnt, ninl);
  __syncthreads();
  if (tid == 0) out_inliers[p] = s_cnt;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace nvs

using namespace nvs;

extern "C" size_t nvs_pose_workspace_bytes(int32_t n_pairs, int32_t kmax, int32_t iters) {
  if (n_pairs <= 0 || kmax <= 0 || iters <= 0) return 0;
  const size_t P = (size_t)n_pairs, K = (size_t)kmax, I = (size_t)iters;
  return 2 * align256(P * K * 8) + align256(P * I * POSE_MAX_CAND * 9 * 4) + align256(P * I * 4) +
         align256(P * I * POSE_MAX_CAND * 8) + align256(P * 8) + 2 * align256(P * 4) + 256;
}

static int pose_batch_impl(const float* pts, int32_t n_frames, int32_t kmax, const int32_t* pair_a,
                           const int32_t* pair_b, const int32_t* idx1, const int32_t* idx2, const int32_t* count,
                           int32_t n_pairs, float fx, float fy, float cx, float cy, float threshold, int32_t iters,
                           uint64_t seed, int32_t refine, float confidence, int32_t round_size, float* out_E, float* out_R,
                           float* out_t, uint8_t* out_mask, int32_t* out_inliers, int32_t* out_iters, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!pts || !pair_a || !pair_b || !count || !out_E || !out_R || !out_t || !out_mask || !out_inliers || !workspace)
    return NVS_ERR_ARG;
  if ((idx1 == nullptr) != (idx2 == nullptr)) return NVS_ERR_ARG;
  if (n_frames <= 0 || kmax <= 0 || n_pairs <= 0 || n_pairs > 65535 || iters <= 0 || iters > 65535) return NVS_ERR_ARG;
  if (!(threshold > 0.f) || fx == 0.f || fy == 0.f || refine < 0 || refine > 100) return NVS_ERR_ARG;
  if (workspace_bytes < nvs_pose_workspace_bytes(n_pairs, kmax, iters)) return NVS_ERR_ARG;
  if (((uintptr_t)workspace & 255) != 0) return NVS_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t P = (size_t)n_pairs, K = (size_t)kmax, I = (size_t)iters;
  char* w = static_cast<char*>(workspace);
  float2* cur = reinterpret_cast<float2*>(w); w += align256(P * K * 8);
  float2* ref = reinterpret_cast<float2*>(w); w += align256(P * K * 8);
  float* cand = reinterpret_cast<float*>(w); w += align256(P * I * POSE_MAX_CAND * 9 * 4);
  int32_t* ncand = reinterpret_cast<int32_t*>(w); w += align256(P * I * 4);
  unsigned long long* score = reinterpret_cast<unsigned long long*>(w); w += align256(P * I * POSE_MAX_CAND * 8);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(w); w += align256(P * 8);
  int32_t* done = reinterpret_cast<int32_t*>(w); w += align256(P * 4);
  int32_t* used = reinterpret_cast<int32_t*>(w);
  PoseIn in{pts, pair_a, pair_b, idx1, idx2, count, kmax, fx, fy, cx, cy};
  pose_gather_kernel<<<dim3((kmax + 255) / 256, n_pairs), 256, 0, st>>>(in, cur, ref);
  NVS_CHECK_LAUNCH();
  const float thr2 = threshold * threshold;
  const bool adaptive = confidence > 0.f;
  if (!adaptive) {
    pose_hypotheses_kernel<<<dim3((iters + 31) / 32, n_pairs), 32, 0, st>>>(cur, ref, count, kmax, iters, seed, cand, ncand,
                                                                            0, iters, nullptr);
    NVS_CHECK_LAUNCH();
    pose_score_kernel<<<dim3(iters, n_pairs), SCORE_T, 0, st>>>(cur, ref, count, kmax, iters, thr2, cand, ncand, score, 0,
                                                                nullptr);
    NVS_CHECK_LAUNCH();
  } else {
    // rounds of `round_size` samples; a pair drops out once its sample count reaches the RANSAC bound for `confidence`
    cudaMemsetAsync(best, 0xFF, P * 8, st);
    cudaMemsetAsync(done, 0, P * 4, st);
    cudaMemsetAsync(used, 0, P * 4, st);
    for (int it0 = 0; it0 < iters; it0 += round_size) {
      const int it1 = it0 + round_size < iters ? it0 + round_size : iters, r = it1 - it0;
      pose_hypotheses_kernel<<<dim3((r + 31) / 32, n_pairs), 32, 0, st>>>(cur, ref, count, kmax, iters, seed, cand, ncand,
                                                                        it0, it1, done);
      NVS_CHECK_LAUNCH();
      pose_score_kernel<<<dim3(r, n_pairs), SCORE_T, 0, st>>>(cur, ref, count, kmax, iters, thr2, cand, ncand, score, it0,
                                                              done);
      NVS_CHECK_LAUNCH();
      pose_adapt_kernel<<<n_pairs, 128, 0, st>>>(count, kmax, iters, it0, it1, confidence, score, best, done, used);
      NVS_CHECK_LAUNCH();
    }
    if (out_iters) cudaMemcpyAsync(out_iters, used, P * 4, cudaMemcpyDeviceToDevice, st);
  }
  pose_select_kernel<<<n_pairs, SELECT_T, 0, st>>>(cur, ref, count, kmax, iters, thr2, cand, score, out_E, out_R, out_t,
                                                   out_mask, out_inliers, adaptive ? used : nullptr);
  NVS_CHECK_LAUNCH();
  if (refine > 0) {
    pose_refine_kernel<<<n_pairs, REFINE_T, 0, st>>>(cur, ref, count, kmax, thr2, refine, out_E, out_R, out_t, out_mask,
                                                     out_inliers);
    NVS_CHECK_LAUNCH();
  }
  return NVS_OK;
}

extern "C" int nvs_pose_batch(const float* pts, int32_t n_frames, int32_t kmax, const int32_t* pair_a,
                              const int32_t* pair_b, const int32_t* idx1, const int32_t* idx2, const int32_t* count,
                              int32_t n_pairs, float fx, float fy, float cx, float cy, float threshold, int32_t iters,
                              uint64_t seed, int32_t refine, float* out_E, float* out_R, float* out_t,
                              uint8_t* out_mask, int32_t* out_inliers, void* workspace, size_t workspace_bytes,
                              void* stream) {
  return pose_batch_impl(pts, n_frames, kmax, pair_a, pair_b, idx1, idx2, count, n_pairs, fx, fy, cx, cy, threshold, iters,
                         seed, refine, 0.f, 0, out_E, out_R, out_t, out_mask, out_inliers, nullptr, workspace,
                         workspace_bytes, stream);
}

extern "C" int nvs_pose_batch_adaptive(const float* pts, int32_t n_frames, int32_t kmax, const int32_t* pair_a,
                                       const int32_t* pair_b, const int32_t* idx1, const int32_t* idx2,
                                       const int32_t* count, int32_t n_pairs, float fx, float fy, float cx, float cy,
                                       float threshold, int32_t max_iters, uint64_t seed, int32_t refine, float confidence,
                                       int32_t round_size, float* out_E, float* out_R, float* out_t, uint8_t* out_mask,
                                       int32_t* out_inliers, int32_t* out_iters, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  if (!(confidence > 0.f && confidence < 1.f) || round_size <= 0) return NVS_ERR_ARG;
  return pose_batch_impl(pts, n_frames, kmax, pair_a, pair_b, idx1, idx2, count, n_pairs, fx, fy, cx, cy, threshold,
                         max_iters, seed, refine, confidence, round_size, out_E, out_R, out_t, out_mask, out_inliers,
                         out_iters, workspace, workspace_bytes, stream);
}
