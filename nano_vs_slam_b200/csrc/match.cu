// Descriptor matching on device.
//   nvs_knn2  = cv2.BFMatcher(NORM_L2).knnMatch(des1, des2, k=2)      (feature_matcher.py:94-96)
//   nvs_match = + ratio test + goodMatchesOneToOne (feature_matcher.py:179-209), or mutual NN
//               (cv2 crossCheck=True, evaluation/descriptor.py:221).
// Distances are exact sum((a-b)^2) in fp32 (no |a|^2+|b|^2-2ab cancellation), sqrt at the end, like
// OpenCV's batchDistance.  The work is tiny (4000 x 4000 x 32) and latency bound: queries sit in
// registers, train rows stream through shared memory as warp-broadcast LDS.128, the train set is split
// over blockIdx.y for occupancy and the partial top-2 lists are merged in split order (ties -> lowest
// train index, as a sequential scan would give).
#include <limits.h>

#include "common.cuh"

namespace nvs {

constexpr int KNN_T = 128;   // queries per CTA
constexpr int KNN_TT = 64;   // train rows per shared tile

// Batched form (nvs_match_batch): blockIdx.z = pair; descriptors of all frames live in one (F, kmax, D) tensor and
// the per-frame keypoint counts stay on the device (no host round trip between selection and matching).
// counts == nullptr: the single-pair entry points, sizes come from the scalar arguments.
struct MatchBatch {
  const int32_t* counts;  // [F] keypoints per frame
  const int32_t* pa;      // [P] query frame of every pair
  const int32_t* pb;      // [P] train frame of every pair
  int kmax;               // rows per frame in the descriptor tensor = pitch of all per-pair arrays
  int swap;               // 1: roles of a and b exchanged (the reverse 2-NN of the mutual test)
};
__device__ __forceinline__ void batch_sizes(const MatchBatch& mb, int pair, int& n1, int& n2, int& fa, int& fb) {
  fa = mb.pa[pair];
  fb = mb.pb[pair];
  if (mb.swap) { const int t = fa; fa = fb; fb = t; }
  n1 = mb.counts[fa];
  n2 = mb.counts[fb];
}

template <int D>
__global__ void __launch_bounds__(KNN_T) knn2_partial_kernel(const float* __restrict__ des1,
                                                             const float* __restrict__ des2,
                                                             float* __restrict__ pd, int32_t* __restrict__ pi,
                                                             int n1, int n2, int per_split, MatchBatch mb) {
  __shared__ __align__(16) float ts[KNN_TT][D];
  if (mb.counts) {
    int fa, fb;
    batch_sizes(mb, blockIdx.z, n1, n2, fa, fb);
    if ((int)(blockIdx.x * KNN_T) >= n1) return;  // whole CTA past this pair's queries
    des1 += (size_t)fa * mb.kmax * D;
    des2 += (size_t)fb * mb.kmax * D;
    pd += (size_t)blockIdx.z * gridDim.y * mb.kmax * 2;
    pi += (size_t)blockIdx.z * gridDim.y * mb.kmax * 2;
    per_split = (n2 + (int)gridDim.y - 1) / (int)gridDim.y;
    if (per_split < 1) per_split = 1;
  }
  const int q = blockIdx.x * KNN_T + threadIdx.x;
  const int split = blockIdx.y;
  const int t_begin = split * per_split, t_end = min(n2, t_begin + per_split);
  float a[D];
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    const float4 v = q < n1 ? *reinterpret_cast<const float4*>(des1 + (size_t)q * D + c)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    a[c] = v.x; a[c + 1] = v.y; a[c + 2] = v.z; a[c + 3] = v.w;
  }
  float d0 = INFINITY, d1 = INFINITY;
  int i0 = -1, i1 = -1;
  for (int t0 = t_begin; t0 < t_end; t0 += KNN_TT) {
    const int nt = min(KNN_TT, t_end - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < KNN_TT * (D / 4); i += KNN_T) {
      const int r = i / (D / 4), c4 = i - r * (D / 4);
      const float4 v = r < nt ? *reinterpret_cast<const float4*>(des2 + (size_t)(t0 + r) * D + c4 * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&ts[r][c4 * 4]) = v;
    }
    __syncthreads();
    for (int r = 0; r < nt; ++r) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        const float4 tv = *reinterpret_cast<const float4*>(&ts[r][c]);
        const float e0 = a[c] - tv.x, e1 = a[c + 1] - tv.y, e2 = a[c + 2] - tv.z, e3 = a[c + 3] - tv.w;
        s = fmaf(e0, e0, s); s = fmaf(e1, e1, s); s = fmaf(e2, e2, s); s = fmaf(e3, e3, s);
      }
      if (s < d1) {
        if (s < d0) { d1 = d0; i1 = i0; d0 = s; i0 = t0 + r; }
        else { d1 = s; i1 = t0 + r; }
      }
    }
  }
  if (q < n1) {
    const size_t o = ((size_t)split * (mb.counts ? mb.kmax : n1) + q) * 2;
    pd[o] = d0; pd[o + 1] = d1; pi[o] = i0; pi[o + 1] = i1;
  }
}

__global__ void knn2_merge_kernel(const float* __restrict__ pd, const int32_t* __restrict__ pi,
                                  int32_t* __restrict__ idx, float* __restrict__ dist, int n1, int nsplit,
                                  MatchBatch mb) {
  int pitch = n1;
  if (mb.counts) {
    int n2, fa, fb;
    batch_sizes(mb, blockIdx.y, n1, n2, fa, fb);
    pitch = mb.kmax;
    pd += (size_t)blockIdx.y * nsplit * mb.kmax * 2;
    pi += (size_t)blockIdx.y * nsplit * mb.kmax * 2;
    idx += (size_t)blockIdx.y * mb.kmax * 2;
    dist += (size_t)blockIdx.y * mb.kmax * 2;
  }
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n1) return;
  float d0 = INFINITY, d1 = INFINITY;
  int i0 = -1, i1 = -1;
  for (int s = 0; s < nsplit; ++s) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float d = pd[((size_t)s * pitch + q) * 2 + j];
      const int i = pi[((size_t)s * pitch + q) * 2 + j];
      if (i < 0) continue;
      if (d < d1) {
        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
        else { d1 = d; i1 = i; }
      }
    }
  }
  idx[q * 2] = i0; idx[q * 2 + 1] = i1;
  dist[q * 2] = sqrtf(d0); dist[q * 2 + 1] = sqrtf(d1);
}

__device__ __forceinline__ int block_scan_incl1024(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  if (warp > 0) v += warp_tot[warp - 1];
  *total = warp_tot[31];
  return v;
}

// mode 0: literal goodMatchesOneToOne.  For train index t let F = first query (lowest index) that passes
// the ratio test with nn(q) = t.  The reference appends (F, t) when F claims t, and afterwards every later
// passing query with d(q) < d(F) overwrites the entry (dist_match[t] is never updated, :201-208), so the
// surviving query is the LAST such q.  Output order = order of first claims = ascending F.
__global__ void __launch_bounds__(1024) one_to_one_kernel(const int32_t* __restrict__ idx,
                                                          const float* __restrict__ dist, int n1, int n2,
                                                          double ratio, int32_t* first_q,
                                                          int32_t* last_q,
                                                          int32_t* __restrict__ out1, int32_t* __restrict__ out2,
                                                          float* __restrict__ outd, int32_t* __restrict__ out_count,
                                                          MatchBatch mb) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x;
  if (mb.counts) {
    int fa, fb;
    batch_sizes(mb, blockIdx.x, n1, n2, fa, fb);
    const size_t o = (size_t)blockIdx.x * mb.kmax;
    idx += o * 2; dist += o * 2; first_q += o; last_q += o; out1 += o; out2 += o; outd += o; out_count += blockIdx.x;
    if (n1 < 1 || n2 < 2) {  // knnMatch(k=2) needs two train rows: no matches for this pair
      if (tid == 0) *out_count = 0;
      return;
    }
  }
  for (int t = tid; t < n2; t += 1024) { first_q[t] = INT_MAX; last_q[t] = -1; }
  __syncthreads();
  for (int q = tid; q < n1; q += 1024) {
    const float m = dist[q * 2], n = dist[q * 2 + 1];
    if (!((double)m > ratio * (double)n)) atomicMin(&first_q[idx[q * 2]], q);
  }
  __syncthreads();
  for (int q = tid; q < n1; q += 1024) {
    const float m = dist[q * 2], n = dist[q * 2 + 1];
    if ((double)m > ratio * (double)n) continue;
    const int t = idx[q * 2], f = first_q[t];
    if (q != f && m < dist[f * 2]) atomicMax(&last_q[t], q);
  }
  __syncthreads();
  int base = 0;
  for (int q0 = 0; q0 < n1; q0 += 1024) {
    const int q = q0 + tid;
    bool is_first = false;
    int t = 0;
    if (q < n1) {
      const float m = dist[q * 2], n = dist[q * 2 + 1];
      t = idx[q * 2];
      is_first = !((double)m > ratio * (double)n) && first_q[t] == q;
    }
    int tot;
    const int incl = block_scan_incl1024(is_first ? 1 : 0, warp_tot, &tot);
    if (is_first) {
      const int pos = base + incl - 1;
      const int win = last_q[t] >= 0 ? last_q[t] : q;
      out1[pos] = win; out2[pos] = t; outd[pos] = dist[win * 2];
    }
    base += tot;
  }
  if (tid == 0) *out_count = base;
}

// mode 1: mutual nearest neighbours, output in ascending query order.
__global__ void __launch_bounds__(1024) mutual_kernel(const int32_t* __restrict__ idx12,
                                                      const float* __restrict__ dist12,
                                                      const int32_t* __restrict__ idx21, int n1,
                                                      int32_t* __restrict__ out1, int32_t* __restrict__ out2,
                                                      float* __restrict__ outd, int32_t* __restrict__ out_count,
                                                      MatchBatch mb) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x;
  if (mb.counts) {
    int n2, fa, fb;
    batch_sizes(mb, blockIdx.x, n1, n2, fa, fb);
    const size_t o = (size_t)blockIdx.x * mb.kmax;
    idx12 += o * 2; dist12 += o * 2; idx21 += o * 2; out1 += o; out2 += o; outd += o; out_count += blockIdx.x;
    if (n2 < 1) n1 = 0;
  }
  int base = 0;
  for (int q0 = 0; q0 < n1; q0 += 1024) {
    const int q = q0 + tid;
    bool keep = false;
    int t = 0;
    if (q < n1) {
      t = idx12[q * 2];
      keep = t >= 0 && idx21[t * 2] == q;
    }
    int tot;
    const int incl = block_scan_incl1024(keep ? 1 : 0, warp_tot, &tot);
    if (keep) {
      const int pos = base + incl - 1;
      out1[pos] = q; out2[pos] = t; outd[pos] = dist12[q * 2];
    }
    base += tot;
  }
  if (tid == 0) *out_count = base;
}

static inline int knn_splits(int n2) {
  int s = (n2 + 511) / 512;
  return s < 1 ? 1 : (s > 16 ? 16 : s);
}
static inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static int run_knn2(const float* des1, const float* des2, int32_t* idx, float* dist, int n1, int n2, int D,
                    float* pd, int32_t* pi, cudaStream_t st) {
  const int ns = knn_splits(n2);
  int per = (n2 + ns - 1) / ns;
  dim3 grid((n1 + KNN_T - 1) / KNN_T, ns);
  const MatchBatch none{nullptr, nullptr, nullptr, 0, 0};
  switch (D) {
    case 32: knn2_partial_kernel<32><<<grid, KNN_T, 0, st>>>(des1, des2, pd, pi, n1, n2, per, none); break;
    case 64: knn2_partial_kernel<64><<<grid, KNN_T, 0, st>>>(des1, des2, pd, pi, n1, n2, per, none); break;
    case 128: knn2_partial_kernel<128><<<grid, KNN_T, 0, st>>>(des1, des2, pd, pi, n1, n2, per, none); break;
    default: return NVS_ERR_UNSUPPORTED;
  }
  NVS_CHECK_LAUNCH();
  knn2_merge_kernel<<<(n1 + 127) / 128, 128, 0, st>>>(pd, pi, idx, dist, n1, ns, none);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

// all pairs in one launch each: grid.z / grid.y / grid.x = pair
static int run_knn2_batch(const float* des, int32_t* idx, float* dist, int D, float* pd, int32_t* pi, int n_pairs,
                          const MatchBatch& mb, cudaStream_t st) {
  const int ns = knn_splits(mb.kmax);
  dim3 grid((mb.kmax + KNN_T - 1) / KNN_T, ns, n_pairs);
  switch (D) {
    case 32: knn2_partial_kernel<32><<<grid, KNN_T, 0, st>>>(des, des, pd, pi, 0, 0, 0, mb); break;
    case 64: knn2_partial_kernel<64><<<grid, KNN_T, 0, st>>>(des, des, pd, pi, 0, 0, 0, mb); break;
    case 128: knn2_partial_kernel<128><<<grid, KNN_T, 0, st>>>(des, des, pd, pi, 0, 0, 0, mb); break;
    default: return NVS_ERR_UNSUPPORTED;
  }
  NVS_CHECK_LAUNCH();
  knn2_merge_kernel<<<dim3((mb.kmax + 127) / 128, n_pairs), 128, 0, st>>>(pd, pi, idx, dist, 0, ns, mb);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}

}  // namespace nvs

using namespace nvs;

// workspace layout: [partials d][partials i] sized for max(n1,n2) queries, [idx12][dist12][idx21][dist21],
// [first_q][last_q]
extern "C" size_t nvs_match_workspace_bytes(int32_t n1, int32_t n2) {
  if (n1 <= 0 || n2 <= 0) return 0;
  const size_t nmax = (size_t)(n1 > n2 ? n1 : n2);
  const size_t part = align256(16 * nmax * 2 * sizeof(float));
  return 2 * part + 2 * align256((size_t)n1 * 2 * 4) + 2 * align256((size_t)n2 * 2 * 4) +
         2 * align256((size_t)n2 * 4);
}

extern "C" int nvs_match(const float* des1, const float* des2, int32_t n1, int32_t n2, int32_t D, double ratio,
                         int32_t mode, int32_t* out_idx1, int32_t* out_idx2, float* out_dist,
                         int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  if (!des1 || !des2 || !out_idx1 || !out_idx2 || !out_dist || !out_count || !workspace) return NVS_ERR_ARG;
  if (n1 <= 0 || n2 <= 0) return NVS_ERR_ARG;
  if (workspace_bytes < nvs_match_workspace_bytes(n1, n2)) return NVS_ERR_ARG;
  if (mode == 0 && n2 < 2) return NVS_ERR_ARG;  // knnMatch(k=2) needs two train rows
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  const size_t nmax = (size_t)(n1 > n2 ? n1 : n2);
  const size_t part = align256(16 * nmax * 2 * sizeof(float));
  float* pd = reinterpret_cast<float*>(w); w += part;
  int32_t* pi = reinterpret_cast<int32_t*>(w); w += part;
  int32_t* idx12 = reinterpret_cast<int32_t*>(w); w += align256((size_t)n1 * 8);
  float* dist12 = reinterpret_cast<float*>(w); w += align256((size_t)n1 * 8);
  int32_t* idx21 = reinterpret_cast<int32_t*>(w); w += align256((size_t)n2 * 8);
  float* dist21 = reinterpret_cast<float*>(w); w += align256((size_t)n2 * 8);
  int32_t* first_q = reinterpret_cast<int32_t*>(w); w += align256((size_t)n2 * 4);
  int32_t* last_q = reinterpret_cast<int32_t*>(w);

  int rc = run_knn2(des1, des2, idx12, dist12, n1, n2, D, pd, pi, st);
  if (rc != NVS_OK) return rc;
  if (mode == 0) {
    one_to_one_kernel<<<1, 1024, 0, st>>>(idx12, dist12, n1, n2, ratio, first_q, last_q, out_idx1, out_idx2,
                                          out_dist, out_count, MatchBatch{nullptr, nullptr, nullptr, 0, 0});
  } else if (mode == 1) {
    rc = run_knn2(des2, des1, idx21, dist21, n2, n1, D, pd, pi, st);
    if (rc != NVS_OK) return rc;
    mutual_kernel<<<1, 1024, 0, st>>>(idx12, dist12, idx21, n1, out_idx1, out_idx2, out_dist, out_count,
                                      MatchBatch{nullptr, nullptr, nullptr, 0, 0});
  } else if (mode == 2) {
    // raw 2-NN: out_idx1 (n1,2) = idx, out_dist (n1,2) = dist; out_idx2 unused
    cudaError_t e = cudaMemcpyAsync(out_idx1, idx12, (size_t)n1 * 8, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist, dist12, (size_t)n1 * 8, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return nvs_set_cuda_error(e);
    return NVS_OK;
  } else {
    return NVS_ERR_ARG;
  }
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}


// ---- batched: P pairs of frames out of one (F, kmax, D) descriptor tensor, counts on the device
// workspace per pair: [partials d][partials i] (16 splits x kmax x 2), [idx12][dist12][idx21][dist21] (kmax x 2),
// [first_q][last_q] (kmax)
extern "C" size_t nvs_match_batch_workspace_bytes(int32_t n_pairs, int32_t kmax) {
  if (n_pairs <= 0 || kmax <= 0) return 0;
  const size_t per = 2 * (size_t)16 * kmax * 2 * 4 + 4 * (size_t)kmax * 2 * 4 + 2 * (size_t)kmax * 4;
  return align256(per * n_pairs) + 256;
}

extern "C" int nvs_match_batch(const float* des, const int32_t* counts, int32_t n_frames, int32_t kmax, int32_t D,
                               const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs, double ratio,
                               int32_t mode, int32_t* out_idx1, int32_t* out_idx2, float* out_dist,
                               int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  if (!des || !counts || !pair_a || !pair_b || !out_idx1 || !out_idx2 || !out_dist || !out_count || !workspace)
    return NVS_ERR_ARG;
  if (n_frames <= 0 || kmax <= 0 || n_pairs <= 0 || n_pairs > 65535) return NVS_ERR_ARG;
  if (mode != 0 && mode != 1) return NVS_ERR_ARG;
  if (workspace_bytes < nvs_match_batch_workspace_bytes(n_pairs, kmax)) return NVS_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t P = (size_t)n_pairs, K = (size_t)kmax;
  char* w = static_cast<char*>(workspace);
  float* pd = reinterpret_cast<float*>(w); w += P * 16 * K * 2 * 4;
  int32_t* pi = reinterpret_cast<int32_t*>(w); w += P * 16 * K * 2 * 4;
  int32_t* idx12 = reinterpret_cast<int32_t*>(w); w += P * K * 2 * 4;
  float* dist12 = reinterpret_cast<float*>(w); w += P * K * 2 * 4;
  int32_t* idx21 = reinterpret_cast<int32_t*>(w); w += P * K * 2 * 4;
  float* dist21 = reinterpret_cast<float*>(w); w += P * K * 2 * 4;
  int32_t* first_q = reinterpret_cast<int32_t*>(w); w += P * K * 4;
  int32_t* last_q = reinterpret_cast<int32_t*>(w);
  MatchBatch mb{counts, pair_a, pair_b, kmax, 0};
  int rc = run_knn2_batch(des, idx12, dist12, D, pd, pi, n_pairs, mb, st);
  if (rc != NVS_OK) return rc;
  if (mode == 0) {
    one_to_one_kernel<<<n_pairs, 1024, 0, st>>>(idx12, dist12, 0, 0, ratio, first_q, last_q, out_idx1, out_idx2,
                                                out_dist, out_count, mb);
  } else {
    MatchBatch rev = mb;
    rev.swap = 1;
    rc = run_knn2_batch(des, idx21, dist21, D, pd, pi, n_pairs, rev, st);
    if (rc != NVS_OK) return rc;
    mutual_kernel<<<n_pairs, 1024, 0, st>>>(idx12, dist12, idx21, 0, out_idx1, out_idx2, out_dist, out_count, mb);
  }
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
