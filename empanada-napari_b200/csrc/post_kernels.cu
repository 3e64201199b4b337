// Per-slice post-processing kernels (pieces 2, 3, 4 of the hot path), batched over the slices
// of one plane. Bit-exact restatements of:
//   * _MedianQueue recursive median + harden   (empanada/inference/engines.py:47-90,115-121)
//   * find_instance_center                     (empanada/inference/postprocess.py:39-76)
//   * group_pixels / chunked_pixel_grouping    (postprocess.py:79-169, engines.py:258-275)
//   * merge_semantic_and_instance (thing path) (postprocess.py:224-296, engines.py:278-298)
// All of it is HBM-bound integer / compare work: coalesced, vectorised where the layout
// allows, centres staged through shared memory.
#include <cuda_runtime.h>
#include <cstdint>

#include "common.cuh"

namespace post {

// ------------------------------------------------------------------ sigmoid + median + harden
// One thread owns one pixel and walks the B new slices in order, carrying the (ks-1)-deep
// queue in registers: the median of a full window REPLACES the queued middle value, exactly
// like `output[key] = self.get_median(key)` does on the deque item (engines.py:79-82).
// hist  : [(ks-1)][HW] fp32 queue contents before this batch (valid entries: n_hist)
// hard  : [N][HW] u8 plane buffer, written at absolute slice index for every emission
// prob  : optional [N][HW] fp32 of the emitted (filtered) probabilities (tests / halo exchange)
template <int KS>
__global__ void median_harden_kernel(const float* __restrict__ logits, int B, long long HW,
                                     float* __restrict__ hist, int n_hist, int slice0,
                                     float conf_thr, int is_prob, uint8_t* __restrict__ hard,
                                     float* __restrict__ prob_out) {
  constexpr int MID = (KS - 1) / 2;
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= HW) return;
  float q[KS];  // q[0..nq) = queue, oldest first
  int nq = n_hist;
#pragma unroll
  for (int i = 0; i < KS - 1; ++i) q[i] = (i < n_hist) ? hist[i * HW + p] : 0.0f;
  for (int b = 0; b < B; ++b) {
    float v = logits[b * HW + p];
    if (!is_prob) v = 1.0f / (1.0f + expf(-v));
    if (nq == KS) {  // deque(maxlen): drop the oldest
#pragma unroll
      for (int i = 0; i < KS - 1; ++i) q[i] = q[i + 1];
      nq = KS - 1;
    }
#pragma unroll
    for (int i = 0; i < KS; ++i)
      if (i == nq) q[i] = v;
    ++nq;
    const int t = slice0 + b;  // absolute index of the slice just pushed
    if (nq <= MID) {
      hard[t * HW + p] = v >= conf_thr;
      if (prob_out) prob_out[t * HW + p] = v;
    } else if (nq == KS) {
      float s[KS];
#pragma unroll
      for (int i = 0; i < KS; ++i) s[i] = q[i];
      // partial selection sort up to the middle order statistic (exact compares only)
#pragma unroll
      for (int i = 0; i <= MID; ++i) {
#pragma unroll
        for (int j = i + 1; j < KS; ++j) {
          const float lo = fminf(s[i], s[j]), hi = fmaxf(s[i], s[j]);
          s[i] = lo; s[j] = hi;
        }
      }
      const float med = s[MID];
      q[MID] = med;
      hard[(t - MID) * HW + p] = med >= conf_thr;
      if (prob_out) prob_out[(t - MID) * HW + p] = med;
    }
  }
  // persist queue for the next batch: keep at most KS-1 newest entries
  if (nq == KS) {
#pragma unroll
    for (int i = 0; i < KS - 1; ++i) q[i] = q[i + 1];
    nq = KS - 1;
  }
#pragma unroll
  for (int i = 0; i < KS - 1; ++i)
    if (i < nq) hist[i * HW + p] = q[i];
}


// Vectorised form: one thread owns FOUR consecutive pixels (16-byte loads of the logits and of
// the queue, one 4-byte store of the hardened mask) and the slice loop is unrolled so that several
// independent 16-byte loads are in flight per thread. Same arithmetic per pixel as above.
template <int KS>
__global__ void __launch_bounds__(256)
median_harden_v4_kernel(const float4* __restrict__ logits, int B, long long HW4,
                        float4* __restrict__ hist, int n_hist, int slice0, float conf_thr,
                        int is_prob, uchar4* __restrict__ hard, float4* __restrict__ prob_out) {
  constexpr int MID = (KS - 1) / 2;
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= HW4) return;
  float q[KS][4];
  int nq = n_hist;
#pragma unroll
  for (int i = 0; i < KS - 1; ++i) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n_hist) t = hist[i * HW4 + p];
    q[i][0] = t.x; q[i][1] = t.y; q[i][2] = t.z; q[i][3] = t.w;
  }
  constexpr int UNROLL = 4;
  for (int b0 = 0; b0 < B; b0 += UNROLL) {
    float4 in[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (b0 + u < B) in[u] = __ldcs(&logits[(b0 + u) * HW4 + p]);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int b = b0 + u;
      if (b >= B) break;
      float v[4] = {in[u].x, in[u].y, in[u].z, in[u].w};
      if (!is_prob) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
      }
      if (nq == KS) {
#pragma unroll
        for (int i = 0; i < KS - 1; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) q[i][j] = q[i + 1][j];
        nq = KS - 1;
      }
#pragma unroll
      for (int i = 0; i < KS; ++i)
        if (i == nq) {
#pragma unroll
          for (int j = 0; j < 4; ++j) q[i][j] = v[j];
        }
      ++nq;
      const int t = slice0 + b;
      if (nq <= MID) {
        hard[t * HW4 + p] = make_uchar4(v[0] >= conf_thr, v[1] >= conf_thr, v[2] >= conf_thr, v[3] >= conf_thr);
        if (prob_out) prob_out[t * HW4 + p] = make_float4(v[0], v[1], v[2], v[3]);
      } else if (nq == KS) {
        float med[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float s[KS];
#pragma unroll
          for (int i = 0; i < KS; ++i) s[i] = q[i][j];
#pragma unroll
          for (int i = 0; i <= MID; ++i) {
#pragma unroll
            for (int k = i + 1; k < KS; ++k) {
              const float lo = fminf(s[i], s[k]), hi = fmaxf(s[i], s[k]);
              s[i] = lo; s[k] = hi;
            }
          }
          med[j] = s[MID];
          q[MID][j] = med[j];
        }
        hard[(t - MID) * HW4 + p] = make_uchar4(med[0] >= conf_thr, med[1] >= conf_thr, med[2] >= conf_thr, med[3] >= conf_thr);
        if (prob_out) prob_out[(t - MID) * HW4 + p] = make_float4(med[0], med[1], med[2], med[3]);
      }
    }
  }
  if (nq == KS) {
#pragma unroll
    for (int i = 0; i < KS - 1; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) q[i][j] = q[i + 1][j];
    nq = KS - 1;
  }
#pragma unroll
  for (int i = 0; i < KS - 1; ++i)
    if (i < nq) hist[i * HW4 + p] = make_float4(q[i][0], q[i][1], q[i][2], q[i][3]);
}

// end(): queue[mid+1:] are emitted unfiltered (engines.py:351-361)
__global__ void median_flush_kernel(const float* __restrict__ hist, int n_hist, int ks,
                                    long long HW, int n_slices_total, float conf_thr,
                                    uint8_t* __restrict__ hard, float* __restrict__ prob_out) {
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= HW) return;
  // after the last push the queue held min(ks, n) items; hist keeps the newest min(ks-1, n).
  // Items still to emit: queue positions mid+1.. -> the last (nq_full - mid - 1) slices.
  const int mid = (ks - 1) / 2;
  const int nq_full = n_slices_total < ks ? n_slices_total : ks;
  const int n_emit = nq_full - mid - 1;
  for (int e = 0; e < n_emit; ++e) {
    const int t = n_slices_total - n_emit + e;
    const float v = hist[(n_hist - n_emit + e) * HW + p];
    hard[t * HW + p] = v >= conf_thr;
    if (prob_out) prob_out[t * HW + p] = v;
  }
}

// ------------------------------------------------------------------ centre NMS + ordered select
// Two passes over 1024-pixel raster chunks (grid = chunks x slices): pass 0 counts the kept
// pixels per chunk, pass 1 recomputes the flags and writes each kept pixel at
// (kept pixels in earlier chunks) + (rank inside the chunk), i.e. torch.nonzero (row-major) order.
__device__ __forceinline__ bool nms_keep(const float* __restrict__ c, int p, int n, int h4, int w4,
                                         float thr, int k) {
  if (p >= n) return false;
  const int pad = k / 2;
  const int y = p / w4, x = p - y * w4;
  const float v0 = c[p];
  const float t = (v0 > thr) ? v0 : -1.0f;
  if (!(t > 0.0f)) return false;
  // window [y-pad, y-pad+k) x [x-pad, x-pad+k): for even k this is the pooled output cell (y, x)
  // of the (h+1, w+1) map whose last row/col the reference drops
  float m = -INFINITY;
  for (int dy = 0; dy < k; ++dy) {
    const int yy = y - pad + dy;
    if (yy < 0 || yy >= h4) continue;
    for (int dx = 0; dx < k; ++dx) {
      const int xx = x - pad + dx;
      if (xx < 0 || xx >= w4) continue;
      const float u = c[yy * w4 + xx];
      m = fmaxf(m, (u > thr) ? u : -1.0f);
    }
  }
  return t == m;
}
template <int PASS>
__global__ void __launch_bounds__(1024)
nms_centers_kernel(const float* __restrict__ ctr, int h4, int w4, float thr, int k,
                   int* __restrict__ centers, int cap, int* __restrict__ counts,
                   int* __restrict__ chunk_counts, int chunks) {
  const int b = blockIdx.y, ch = blockIdx.x;
  const int n = h4 * w4;
  const int p = ch * 1024 + threadIdx.x;
  const bool keep = nms_keep(ctr + static_cast<long long>(b) * n, p, n, h4, w4, thr, k);
  if (PASS == 0) {
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) chunk_counts[b * chunks + ch] = c;
    return;
  }
  __shared__ int warp_sums[32];
  __shared__ int base;
  // kept pixels of all earlier chunks of this slice (chunks <= a few thousand: one strided sum)
  int part = 0;
  for (int i = threadIdx.x; i < ch; i += 1024) part += chunk_counts[b * chunks + i];
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < 32; ++i) s += warp_sums[i];
    base = s;
  }
  __syncthreads();
  const int chunk_base = base;
  __syncthreads();
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_sums[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    int v = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    warp_sums[lane] = v;  // inclusive
  }
  __syncthreads();
  const int pos = chunk_base + ((warp == 0) ? 0 : warp_sums[warp - 1]) + __popc(ballot & ((1u << lane) - 1));
  if (keep && pos < cap) {
    const int y = p / w4, x = p - y * w4;
    centers[static_cast<long long>(b) * cap + pos] = (y << 16) | x;
  }
  if (ch == chunks - 1 && threadIdx.x == 0) counts[b] = chunk_base + warp_sums[31];  // may exceed cap
}

// ------------------------------------------------------------------ nearest-centre grouping
// cells4[b][y][x] = 1 + argmin_k sqrt(fma(dx,dx, dy*dy)) with first-minimum tie rule; the
// K > 20 path of the reference additionally leaves 0 where every distance is >= 1e5.
// Arithmetic replicates ATen's CPU vector_norm (sqrt(fma(dx, dx, fl(dy*dy)))), see DESIGN.md.
constexpr int GROUP_TILE = 1024;
__global__ void __launch_bounds__(256)
group_pixels_kernel(const float* __restrict__ off, const int* __restrict__ centers, int cap,
                    const int* __restrict__ counts, int h4, int w4, float step,
                    int* __restrict__ cells4) {
  const int b = blockIdx.y;
  const int n = h4 * w4;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int K = counts[b];
  if (K > cap) K = cap;
  __shared__ float cy[GROUP_TILE], cx[GROUP_TILE];
  float ly = 0.f, lx = 0.f;
  if (p < n) {
    const int y = p / w4, x = p - y * w4;
    const float* ob = off + static_cast<long long>(b) * 2 * n;
    ly = __fadd_rn(__fmul_rn(static_cast<float>(y), step), ob[p]);
    lx = __fadd_rn(__fmul_rn(static_cast<float>(x), step), ob[n + p]);
  }
  float best = INFINITY;
  int best_k = -1;
  const int* cb = centers + static_cast<long long>(b) * cap;
  for (int k0 = 0; k0 < K; k0 += GROUP_TILE) {
    const int kn = min(GROUP_TILE, K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kn; i += blockDim.x) {
      const int packed = cb[k0 + i];
      cy[i] = __fmul_rn(step, static_cast<float>(static_cast<unsigned>(packed) >> 16));
      cx[i] = __fmul_rn(step, static_cast<float>(packed & 0xFFFF));
    }
    __syncthreads();
    if (p < n) {
      for (int i = 0; i < kn; ++i) {
        const float dy = __fsub_rn(cy[i], ly);
        const float dx = __fsub_rn(cx[i], lx);
        const float d = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (d < best) { best = d; best_k = k0 + i; }
      }
    }
  }
  if (p < n) {
    int id = best_k + 1;                    // K == 0 -> 0
    if (K > 20 && !(best < 1e5f)) id = 0;   // chunked path: strict < against the 1e5 initialiser
    cells4[static_cast<long long>(b) * n + p] = id;
  }
}

// ------------------------------------------------------------------ semantic / instance merge
// Pass 1: which instance ids own at least one thing pixel.  present: [B][cap+1] (zeroed).
__global__ void merge_flags_kernel(const uint8_t* __restrict__ hard, const int* __restrict__ cells4,
                                   int B, int H, int W, int h, int w, int scale, int cap,
                                   int* __restrict__ present) {
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w || y >= h) return;
  // NOTE: the reference renumbers over the PADDED slice (postprocess runs before the crop,
  // engines.py:389-392), so ids present only in the padding still consume a number.
  (void)B;
  const int w4 = W / scale;
  const long long hw = static_cast<long long>(H) * W;
  if (hard[b * hw + static_cast<long long>(y) * W + x]) {
    const int id = cells4[(static_cast<long long>(b) * (H / scale) + y / scale) * w4 + x / scale];
    if (id > 0) present[static_cast<long long>(b) * (cap + 1) + id] = 1;
  }
}

// Pass 2 (one CTA per slice): exclusive scan of the flags -> dense new ids 1..K'.
__global__ void __launch_bounds__(1024)
rank_ids_kernel(int* __restrict__ present, int cap, int label_divisor, int class_id) {
  int* f = present + static_cast<long long>(blockIdx.x) * (cap + 1);
  __shared__ int warp_sums[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start <= cap; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const int v = (i <= cap) ? f[i] : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, v != 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
      int s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += u;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const int rank = base + ((warp == 0) ? 0 : warp_sums[warp - 1]) +
                     __popc(ballot & ((1u << lane) - 1)) + 1;
    if (i <= cap) f[i] = v ? (class_id * label_divisor + rank) : 0;
    __syncthreads();
    if (threadIdx.x == 0) base += warp_sums[31];
    __syncthreads();
  }
}

// Pass 3: pan[b][y][x] (cropped h x w, int32).
__global__ void write_pan_kernel(const uint8_t* __restrict__ hard, const int* __restrict__ cells4,
                                 const int* __restrict__ newid, int H, int W, int h, int w,
                                 int scale, int cap, int void_label, int* __restrict__ pan) {
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w || y >= h) return;
  const int w4 = W / scale;
  const long long hw = static_cast<long long>(H) * W;
  int v = void_label;
  if (hard[b * hw + static_cast<long long>(y) * W + x]) {
    const int id = cells4[(static_cast<long long>(b) * (H / scale) + y / scale) * w4 + x / scale];
    if (id > 0) v = newid[static_cast<long long>(b) * (cap + 1) + id];
  }
  pan[(static_cast<long long>(b) * h + y) * w + x] = v;
}

// Quad variants (W % 4 == 0, w % 4 == 0): one thread = four pixels of a row, `hard` read as one
// 32-bit word. Slices are mostly background, so most quads end after that single load (flags) or
// after one 16-byte store of void labels (pan).
__global__ void merge_flags_v4_kernel(const uint8_t* __restrict__ hard, const int* __restrict__ cells4,
                                      int H, int W, int scale, int cap, int* __restrict__ present) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (x0 >= W) return;
  const long long hw = static_cast<long long>(H) * W;
  const uint32_t hq = __ldg(reinterpret_cast<const uint32_t*>(hard + b * hw + static_cast<long long>(y) * W + x0));
  if (hq == 0) return;
  const int w4 = W / scale;
  const int* crow = cells4 + (static_cast<long long>(b) * (H / scale) + y / scale) * w4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (((hq >> (8 * j)) & 0xff) == 0) continue;
    const int id = crow[(x0 + j) / scale];
    if (id > 0) present[static_cast<long long>(b) * (cap + 1) + id] = 1;
  }
}
__global__ void write_pan_v4_kernel(const uint8_t* __restrict__ hard, const int* __restrict__ cells4,
                                    const int* __restrict__ newid, int H, int W, int h, int w,
                                    int scale, int cap, int void_label, int* __restrict__ pan) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (x0 >= w) return;
  const long long hw = static_cast<long long>(H) * W;
  const uint32_t hq = __ldg(reinterpret_cast<const uint32_t*>(hard + b * hw + static_cast<long long>(y) * W + x0));
  int v[4] = {void_label, void_label, void_label, void_label};
  if (hq != 0) {
    const int w4 = W / scale;
    const int* crow = cells4 + (static_cast<long long>(b) * (H / scale) + y / scale) * w4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (((hq >> (8 * j)) & 0xff) == 0) continue;
      const int id = crow[(x0 + j) / scale];
      if (id > 0) v[j] = newid[static_cast<long long>(b) * (cap + 1) + id];
    }
  }
  *reinterpret_cast<int4*>(pan + (static_cast<long long>(b) * h + y) * w + x0) = make_int4(v[0], v[1], v[2], v[3]);
}

}  // namespace post

// ------------------------------------------------------------------------------ launchers
extern "C" {

int be_median_push(const float* logits, int B, int H, int W, int ks, float* hist, int n_hist,
                   int slice0, float conf_thr, int is_prob, uint8_t* hard, float* prob_out,
                   cudaStream_t stream) {
  const long long HW = static_cast<long long>(H) * W;
  const int threads = 256;
  const bool vec = (HW % 4 == 0) && ks <= 7 &&
                   (((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(hist) |
                      reinterpret_cast<uintptr_t>(prob_out)) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(hard) & 3) == 0);
  if (vec) {
    const long long HW4 = HW / 4;
    const unsigned blocks4 = static_cast<unsigned>((HW4 + threads - 1) / threads);
#define LAUNCH_MED4(KS)                                                                          \
  post::median_harden_v4_kernel<KS><<<blocks4, threads, 0, stream>>>(                            \
      reinterpret_cast<const float4*>(logits), B, HW4, reinterpret_cast<float4*>(hist), n_hist,  \
      slice0, conf_thr, is_prob, reinterpret_cast<uchar4*>(hard), reinterpret_cast<float4*>(prob_out))
    switch (ks) {
      case 1: LAUNCH_MED4(1); break;
      case 3: LAUNCH_MED4(3); break;
      case 5: LAUNCH_MED4(5); break;
      case 7: LAUNCH_MED4(7); break;
      default: return be_set_error("median kernel size must be odd and <= 11");
    }
#undef LAUNCH_MED4
    return be_check_launch("median_harden_v4_kernel");
  }
  const unsigned blocks = static_cast<unsigned>((HW + threads - 1) / threads);
#define LAUNCH_MED(KS)                                                                         \
  post::median_harden_kernel<KS><<<blocks, threads, 0, stream>>>(logits, B, HW, hist, n_hist,  \
                                                                 slice0, conf_thr, is_prob,    \
                                                                 hard, prob_out)
  switch (ks) {
    case 1: LAUNCH_MED(1); break;
    case 3: LAUNCH_MED(3); break;
    case 5: LAUNCH_MED(5); break;
    case 7: LAUNCH_MED(7); break;
    case 9: LAUNCH_MED(9); break;
    case 11: LAUNCH_MED(11); break;
    default: return be_set_error("median kernel size must be odd and <= 11");
  }
#undef LAUNCH_MED
  return be_check_launch("median_harden_kernel");
}

int be_median_flush(const float* hist, int n_hist, int ks, int H, int W, int n_slices_total,
                    float conf_thr, uint8_t* hard, float* prob_out, cudaStream_t stream) {
  const long long HW = static_cast<long long>(H) * W;
  const int threads = 256;
  post::median_flush_kernel<<<static_cast<unsigned>((HW + threads - 1) / threads), threads, 0,
                              stream>>>(hist, n_hist, ks, HW, n_slices_total, conf_thr, hard,
                                        prob_out);
  return be_check_launch("median_flush_kernel");
}

// chunk_counts: scratch [B * ceil(h4*w4/1024)] int32
int be_centers(const float* ctr, int B, int h4, int w4, float thr, int k, int* centers, int cap,
               int* counts, int* chunk_counts, cudaStream_t stream) {
  if (h4 >= 65536 || w4 >= 65536) return be_set_error("head map too large for packed centres");
  const int chunks = (h4 * w4 + 1023) / 1024;
  post::nms_centers_kernel<0><<<dim3(chunks, B), 1024, 0, stream>>>(ctr, h4, w4, thr, k, centers, cap, counts, chunk_counts, chunks);
  post::nms_centers_kernel<1><<<dim3(chunks, B), 1024, 0, stream>>>(ctr, h4, w4, thr, k, centers, cap, counts, chunk_counts, chunks);
  return be_check_launch("nms_centers_kernel");
}

int be_group_pixels(const float* off, const int* centers, int cap, const int* counts, int B,
                    int h4, int w4, float step, int* cells4, cudaStream_t stream) {
  dim3 grid((h4 * w4 + 255) / 256, B);
  post::group_pixels_kernel<<<grid, 256, 0, stream>>>(off, centers, cap, counts, h4, w4, step,
                                                      cells4);
  return be_check_launch("group_pixels_kernel");
}

// in-place: presence flags [B][cap+1] -> dense per-class ids class*div + rank (0 where absent)
int be_rank_ids(int* present, int B, int cap, int label_divisor, int class_id, cudaStream_t stream) {
  post::rank_ids_kernel<<<B, 1024, 0, stream>>>(present, cap, label_divisor, class_id);
  return be_check_launch("rank_ids_kernel");
}

int be_merge_pan(const uint8_t* hard, const int* cells4, int B, int H, int W, int h, int w,
                 int scale, int cap, int label_divisor, int class_id, int void_label,
                 int* present, int* pan, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(present, 0, sizeof(int) * static_cast<size_t>(B) * (cap + 1), stream);
  if (e != cudaSuccess) return be_set_error(cudaGetErrorString(e));
  // flags over the PADDED extent (H x W), pan over the cropped extent (h x w)
  const bool quads = (W % 4 == 0) && (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(hard) & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(pan) & 15) == 0);
  if (quads) {
    dim3 gridp((W / 4 + 127) / 128, H, B);
    post::merge_flags_v4_kernel<<<gridp, 128, 0, stream>>>(hard, cells4, H, W, scale, cap, present);
  } else {
    dim3 gridp((W + 255) / 256, H, B);
    post::merge_flags_kernel<<<gridp, 256, 0, stream>>>(hard, cells4, B, H, W, H, W, scale, cap, present);
  }
  post::rank_ids_kernel<<<B, 1024, 0, stream>>>(present, cap, label_divisor, class_id);
  if (quads) {
    dim3 grid((w / 4 + 127) / 128, h, B);
    post::write_pan_v4_kernel<<<grid, 128, 0, stream>>>(hard, cells4, present, H, W, h, w, scale, cap, void_label, pan);
  } else {
    dim3 grid((w + 255) / 256, h, B);
    post::write_pan_kernel<<<grid, 256, 0, stream>>>(hard, cells4, present, H, W, h, w, scale, cap,
                                                     void_label, pan);
  }
  return be_check_launch("merge_pan kernels");
}

}  // extern "C"
