// Keypoint selection on device = the numpy glue of KP2DtinyFrontend.run (frontend.py:94-126):
//   mask = score > thresh (strict) [& ~isin(label, classes_to_filter)]; if more than top_k pass, keep the
//   top_k by score (np.argpartition: an unordered set with arbitrary tie choice -> here ties go to the
//   lowest cell index and the output is compacted in ascending cell order, deterministic).
// One CTA per frame: a 4-pass 8-bit radix select over the score bits finds the k-th largest score, then
// one ordered compaction pass writes points / descriptors / labels.  Single launch, no host round trip.
#include "common.cuh"

namespace nvs {

constexpr int SEL_T = 1024;

__device__ __forceinline__ uint32_t float_key(float f) {  // monotone float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// inclusive block scan of one int per thread; returns inclusive value, total via *total
__device__ __forceinline__ int block_scan_incl(int v, int* warp_tot /*[32] shared*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();  // protect warp_tot reuse
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

__global__ void __launch_bounds__(SEL_T) select_kernel(
    const float* __restrict__ score, const float* __restrict__ coord, const float* __restrict__ feat,
    const int64_t* __restrict__ seg_cells, const int32_t* __restrict__ filter, int n_filter, float thresh,
    int top_k, float* __restrict__ out_pts, float* __restrict__ out_desc, float* __restrict__ out_score,
    int32_t* __restrict__ out_cell, int64_t* __restrict__ out_label, int32_t* __restrict__ out_count,
    int n_cells, int D) {
  __shared__ int hist[256];
  __shared__ int warp_tot[32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need;
  __shared__ int s_filter[64];

  const int b = blockIdx.x, tid = threadIdx.x;
  const float* sc = score + (size_t)b * n_cells;
  const int64_t* lab = seg_cells ? seg_cells + (size_t)b * n_cells : nullptr;
  const int nf = (filter && lab) ? min(n_filter, 64) : 0;
  if (tid < nf) s_filter[tid] = filter[tid];
  __syncthreads();

  auto passes = [&](int i) -> bool {
    const float v = sc[i];
    if (!(v > thresh)) return false;
    if (nf) {
      const int l = (int)lab[i];
      for (int f = 0; f < nf; ++f)
        if (s_filter[f] == l) return false;
    }
    return true;
  };

  // ---- count candidates ----
  int cnt = 0;
  for (int i = tid; i < n_cells; i += SEL_T) cnt += passes(i) ? 1 : 0;
  int total;
  block_scan_incl(cnt, warp_tot, &total);
  const int k = (top_k > 0) ? top_k : n_cells;

  uint32_t pivot = 0;  // keep keys > pivot, plus `need` keys == pivot (lowest cells first)
  int need = 0;
  const bool select = total > k;
  if (select) {
    // radix select the k-th largest key among candidates, MSB first
    if (tid == 0) {
      s_prefix = 0;
      s_need = k;
    }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < 256; i += SEL_T) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = tid; i < n_cells; i += SEL_T) {
        if (!passes(i)) continue;
        const uint32_t key = float_key(sc[i]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int rem = s_need, bin = 255;
        for (; bin > 0; --bin) {
          if (hist[bin] >= rem) break;
          rem -= hist[bin];
        }
        s_need = rem;  // rank inside the chosen bin
        s_prefix = prefix | ((uint32_t)bin << shift);
      }
      __syncthreads();
    }
    pivot = s_prefix;
    need = s_need;  // how many keys equal to pivot are kept
  }

  // ---- ordered compaction ----
  int base = 0, eq_seen = 0;
  for (int c0 = 0; c0 < n_cells; c0 += SEL_T) {
    const int i = c0 + tid;
    bool pass_ = false, eq = false, gt = false;
    if (i < n_cells && passes(i)) {
      pass_ = true;
      if (select) {
        const uint32_t key = float_key(sc[i]);
        gt = key > pivot;
        eq = key == pivot;
      }
    }
    bool keep = pass_;
    if (select) {
      int eq_tot;
      const int eq_incl = block_scan_incl(eq ? 1 : 0, warp_tot, &eq_tot);
      keep = gt || (eq && (eq_seen + eq_incl) <= need);
      eq_seen += eq_tot;
    }
    int keep_tot;
    const int pos_incl = block_scan_incl(keep ? 1 : 0, warp_tot, &keep_tot);
    if (keep) {
      const int pos = base + pos_incl - 1;
      const size_t o = (size_t)b * k + pos;
      out_pts[o * 2 + 0] = coord[((size_t)b * 2 + 0) * n_cells + i];
      out_pts[o * 2 + 1] = coord[((size_t)b * 2 + 1) * n_cells + i];
      out_score[o] = sc[i];
      out_cell[o] = i;
      if (out_label && lab) out_label[o] = lab[i];
      if (out_desc && feat)
        for (int d = 0; d < D; ++d) out_desc[o * D + d] = feat[((size_t)b * D + d) * n_cells + i];
    }
    base += keep_tot;
  }
  if (tid == 0) out_count[b] = base;
}

}  // namespace nvs

extern "C" int nvs_select_keypoints(const float* score, const float* coord, const float* feat,
                                    const int64_t* seg_cells, const int32_t* filter_classes, int32_t n_filter,
                                    float thresh, int32_t top_k, float* out_pts, float* out_desc,
                                    float* out_score, int32_t* out_cell, int64_t* out_label,
                                    int32_t* out_count, int32_t B, int32_t n_cells, int32_t D, void* stream) {
  if (!score || !coord || !out_pts || !out_score || !out_cell || !out_count) return NVS_ERR_ARG;
  if (B <= 0 || n_cells <= 0 || n_cells > (1 << 24)) return NVS_ERR_ARG;
  if (n_filter > 64) return NVS_ERR_UNSUPPORTED;
  if (out_desc && (!feat || D <= 0)) return NVS_ERR_ARG;
  nvs::select_kernel<<<B, nvs::SEL_T, 0, static_cast<cudaStream_t>(stream)>>>(
      score, coord, feat, seg_cells, filter_classes, n_filter, thresh, top_k, out_pts, out_desc, out_score,
      out_cell, out_label, out_count, n_cells, D);
  NVS_CHECK_LAUNCH();
  return NVS_OK;
}
