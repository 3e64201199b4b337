// Run-based per-plane post-processing (pieces 3, 5 and 6b of the hot path), batched over all
// slices of a plane. After the median queue has produced the hardened semantic mask, everything
// up to the painted label volume works on ROW RUNS instead of pixels:
//
//   group_flags   : nearest-centre grouping (postprocess.py:119-169, engines.py:258-275) evaluated
//                   only for head-grid cells that contain a thing pixel, fused with the "which
//                   instance ids own a thing pixel" flags of merge_semantic_and_instance
//                   (postprocess.py:224-296)
//   rowruns       : maximal horizontal runs of equal panoptic value inside one class range
//                   (the pan_seg of engines.py:278-298 restricted as in rle.py:60-66), extracted
//                   straight from the mask + cell ids: the dense pan_seg never reaches HBM
//   runs_cc       : 8-connected components of equal-valued pixels (skimage.measure.label semantics,
//                   rle.py:18-24) as a union-find over runs; ids in raster order of first pixel
//   runs_stats    : area + bounding box per component (regionprops, rle.py:75-83)
//   runs_overlap  : pixel overlap between components of adjacent slices (what rle_intersection
//                   measures for the matcher, array_utils.py:375-407) as run-interval intersections
//   runs_paint    : the relabelled (D,H,W) volume written exactly once, coalesced, by gathering the
//                   runs that cover each output row / tile (fill_volume, patterns.py:204-213)
//
// HBM traffic per pixel of a plane: 1 B mask + 0.25 B cell ids per dense pass (three passes) and
// one 4 B write of the painted volume; everything else is proportional to the number of runs.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdint>

#include "common.cuh"

namespace runs {

// ------------------------------------------------------------------ grouping + presence flags
// cells [B][H/scale][W/scale] and present [B][cap+1] must be zero on entry. Arithmetic identical
// to post::group_pixels_kernel (bit-exact restatement of ATen's CPU vector_norm).
constexpr int GROUP_TILE = 1024;
__global__ void __launch_bounds__(256)
group_flags_kernel(const uint8_t* __restrict__ hard, const float* __restrict__ off,
                   const int* __restrict__ centers, int cap, const int* __restrict__ counts, int H,
                   int W, int scale, int gstep, int* __restrict__ cells, int* __restrict__ present) {
  const int b = blockIdx.y;
  const int h4 = H / scale, w4 = W / scale;
  const int n = h4 * w4;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int K = counts[b];
  if (K > cap) K = cap;
  bool fg = false;
  int y = 0, x = 0;
  if (p < n && K > 0) {
    y = p / w4;
    x = p - y * w4;
    const uint8_t* hb = hard + static_cast<long long>(b) * H * W;
    if (scale == 4) {
      uint32_t any = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        any |= __ldg(reinterpret_cast<const uint32_t*>(hb + static_cast<long long>(4 * y + r) * W + 4 * x));
      fg = any != 0;
    } else {
      for (int r = 0; r < scale && !fg; ++r)
        for (int c = 0; c < scale; ++c)
          if (hb[static_cast<long long>(scale * y + r) * W + scale * x + c]) { fg = true; break; }
    }
  }
  if (!__syncthreads_or(fg)) return;
  __shared__ float cy[GROUP_TILE], cx[GROUP_TILE];
  // `gstep`: grid step of the head maps in model pixels (4, or 1 with fine boundaries); `scale`:
  // output pixels per cell (= gstep x upsampling, engines.py:263-275)
  const float step = static_cast<float>(gstep);
  float ly = 0.f, lx = 0.f;
  if (fg) {
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
      const unsigned packed = static_cast<unsigned>(cb[k0 + i]);
      cy[i] = __fmul_rn(step, static_cast<float>(packed >> 16));
      cx[i] = __fmul_rn(step, static_cast<float>(packed & 0xFFFFu));
    }
    __syncthreads();
    if (fg) {
      for (int i = 0; i < kn; ++i) {
        const float dy = __fsub_rn(cy[i], ly);
        const float dx = __fsub_rn(cx[i], lx);
        const float d = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (d < best) { best = d; best_k = k0 + i; }
      }
    }
  }
  if (fg) {
    int id = best_k + 1;
    if (K > 20 && !(best < 1e5f)) id = 0;   // chunked path of the reference (postprocess.py:79-116)
    if (id > 0) {
      cells[static_cast<long long>(b) * n + p] = id;
      present[static_cast<long long>(b) * (cap + 1) + id] = 1;
    }
  }
}

// ------------------------------------------------------------------ per-slice mask area
__global__ void __launch_bounds__(256)
slice_area_kernel(const uint8_t* __restrict__ hard, long long n16, int* __restrict__ area) {
  const uint4* src = reinterpret_cast<const uint4*>(hard) + static_cast<long long>(blockIdx.y) * n16;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = (static_cast<long long>(blockIdx.x) * 4 + k) * 256 + threadIdx.x;
    if (i < n16) {
      const uint4 v = __ldg(src + i);
      // bytes are 0 / 1: the popcount of the low bits is the number of set pixels
      cnt += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) + __popc(v.w & 0x01010101u);
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(area + blockIdx.y, cnt);
}

// ------------------------------------------------------------------ row runs
// The panoptic value of a pixel is  hard ? cellval[cell] : bg  (restricted to the class range):
// `resolve_cells_kernel` first turns the per-cell centre ids into final values in place
// (cell id -> per-class running counter, postprocess.py:273-288; void / out-of-range -> 0), so the
// per-pixel passes below chase one pointer instead of two.
__host__ __device__ __forceinline__ int cls_filter(int v, int lo, int hi) { return (v >= lo && v < hi && v != 0) ? v : 0; }
__global__ void __launch_bounds__(256)
resolve_cells_kernel(int* __restrict__ cells, const int* __restrict__ newid, long long per_slice, int cap,
                     int void_label, int lo, int hi, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int id = cells[i];
  const int v = id > 0 ? __ldg(newid + (i / per_slice) * (cap + 1) + id) : void_label;
  cells[i] = cls_filter(v, lo, hi);
}

struct Src {   // one plane's mask / resolved cell values
  const uint8_t* hard; const int* cellval;
  int H, W, h, w, scale, bg;
};
__device__ __forceinline__ int val_at(const Src& s, const uint8_t* hb, const int* cb, int y, int x) {
  return hb[static_cast<long long>(y) * s.W + x] ? __ldg(cb + (y / s.scale) * (s.W / s.scale) + x / s.scale) : s.bg;
}

// One thread = 16 consecutive pixels of a row (one 16-byte load of the mask, one 16-byte load of
// four cell values at scale 4); groups are numbered row-major over the cropped slice (gpr per row).
// PASS 0: heads / tails per 256-group chunk. PASS 1: run records at
// slice_off[b] + (scanned chunk offset) + (rank inside the chunk); k-th head and k-th tail of a
// slice delimit the same run. Also emits row_ptr (slice-local index of the first run of a row).
constexpr int RR_THREADS = 256;
constexpr int RR_PX = 16;
template <int PASS>
__global__ void __launch_bounds__(RR_THREADS)
rowruns_kernel(Src s, int gpr, int chunks, int* __restrict__ counts /*[2][B][chunks]*/, int B,
               const int* __restrict__ slice_off, int* __restrict__ row_ptr /*[B][h+1]*/,
               int2* __restrict__ run_yx, int* __restrict__ run_x1, int* __restrict__ run_val,
               int* __restrict__ L) {
  const int b = blockIdx.y, ch = blockIdx.x;
  const int g = ch * RR_THREADS + threadIdx.x;
  const int y = g / gpr, gi = g - y * gpr;
  const int x0 = RR_PX * gi;
  const bool valid = y < s.h;
  int v[RR_PX];
#pragma unroll
  for (int i = 0; i < RR_PX; ++i) v[i] = 0;
  unsigned hm = 0, tm = 0;
  if (valid) {
    const uint8_t* hb = s.hard + static_cast<long long>(b) * s.H * s.W;
    const int w4 = s.W / s.scale;
    const int* cb = s.cellval + static_cast<long long>(b) * (s.H / s.scale) * w4;
    const uint4 hq = __ldg(reinterpret_cast<const uint4*>(hb + static_cast<long long>(y) * s.W + x0));
    const unsigned hw[4] = {hq.x, hq.y, hq.z, hq.w};
    if ((hq.x | hq.y | hq.z | hq.w) != 0 || s.bg != 0) {
      const int* crow = cb + (y / s.scale) * w4;
      if (s.scale == 4) {
        const int4 cv = __ldg(reinterpret_cast<const int4*>(crow + x0 / 4));
        const int c4[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[4 * k + j] = (x0 + 4 * k + j < s.w) ? (((hw[k] >> (8 * j)) & 0xffu) ? c4[k] : s.bg) : 0;
      } else {
#pragma unroll
        for (int i = 0; i < RR_PX; ++i)
          if (x0 + i < s.w) v[i] = ((hw[i >> 2] >> (8 * (i & 3))) & 0xffu) ? __ldg(crow + (x0 + i) / s.scale) : s.bg;
      }
      const int left = (x0 > 0 && v[0] != 0) ? val_at(s, hb, cb, y, x0 - 1) : 0;
      const int right = (x0 + RR_PX < s.w && v[RR_PX - 1] != 0) ? val_at(s, hb, cb, y, x0 + RR_PX) : 0;
#pragma unroll
      for (int i = 0; i < RR_PX; ++i) {
        if (v[i] == 0) continue;
        const int pv = (i == 0) ? left : v[i - 1];
        const int nv = (i == RR_PX - 1) ? right : v[i + 1];
        if (x0 + i == 0 || pv != v[i]) hm |= 1u << i;
        if (x0 + i == s.w - 1 || nv != v[i]) tm |= 1u << i;
      }
    }
  }
  const long long cidx = static_cast<long long>(b) * chunks + ch;
  int* counts_t = counts + static_cast<long long>(B) * chunks;
  if (!__syncthreads_or(hm != 0)) {
    if (PASS == 0) {
      if (threadIdx.x == 0) { counts[cidx] = 0; counts_t[cidx] = 0; }
    } else if (valid && gi == 0) {
      row_ptr[static_cast<long long>(b) * (s.h + 1) + y] = counts[cidx];
    }
    return;
  }
  // heads in the low half, tails in the high half of one scanned word (<= 4096 each per chunk)
  const int packed = __popc(hm) | (__popc(tm) << 16);
  typedef cub::BlockScan<int, RR_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  int ex, tot;
  Scan(tmp).ExclusiveSum(packed, ex, tot);
  if (PASS == 0) {
    if (threadIdx.x == 0) { counts[cidx] = tot & 0xffff; counts_t[cidx] = tot >> 16; }
    return;
  }
  const int exh = ex & 0xffff, ext = ex >> 16;
  const int offh = counts[cidx], offt = counts_t[cidx];
  if (valid && gi == 0) row_ptr[static_cast<long long>(b) * (s.h + 1) + y] = offh + exh;
  const int so = slice_off[b];
  if (hm) {
    int pos = so + offh + exh;
#pragma unroll
    for (int i = 0; i < RR_PX; ++i)
      if ((hm >> i) & 1u) { run_yx[pos] = make_int2(y, x0 + i); run_val[pos] = v[i]; L[pos] = pos; ++pos; }
  }
  if (tm) {
    int pos = so + offt + ext;
#pragma unroll
    for (int i = 0; i < RR_PX; ++i)
      if ((tm >> i) & 1u) { run_x1[pos] = x0 + i + 1; ++pos; }
  }
}

// one CTA per slice: exclusive scan (in place) of the per-chunk head and tail counts;
// n_runs[b] = runs of the slice; row_ptr[b][h] = n_runs[b]
__global__ void __launch_bounds__(1024)
rowruns_scan_kernel(int* __restrict__ counts, int B, int chunks, int* __restrict__ n_runs,
                    int* __restrict__ row_ptr, int h) {
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int carry;
  const int b = blockIdx.x;
  for (int half = 0; half < 2; ++half) {
    int* c = counts + (static_cast<long long>(half) * B + b) * chunks;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int s0 = 0; s0 < chunks; s0 += 1024) {
      const int i = s0 + threadIdx.x;
      const int v = (i < chunks) ? c[i] : 0;
      int ex, total;
      Scan(tmp).ExclusiveSum(v, ex, total);
      if (i < chunks) c[i] = carry + ex;
      __syncthreads();
      if (threadIdx.x == 0) carry += total;
      __syncthreads();
    }
    if (half == 0 && threadIdx.x == 0) {
      n_runs[b] = carry;
      row_ptr[static_cast<long long>(b) * (h + 1) + h] = carry;
    }
    __syncthreads();
  }
}

// slice_off[0..B] = exclusive scan of n_runs; stats[0] = total, stats[1] = max per slice
__global__ void __launch_bounds__(1024)
slice_offsets_kernel(const int* __restrict__ n_runs, int B, int* __restrict__ slice_off,
                     int* __restrict__ stats) {
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int carry, mx;
  if (threadIdx.x == 0) { carry = 0; mx = 0; }
  __syncthreads();
  for (int s0 = 0; s0 < B; s0 += 1024) {
    const int i = s0 + threadIdx.x;
    const int v = (i < B) ? n_runs[i] : 0;
    int ex, total;
    Scan(tmp).ExclusiveSum(v, ex, total);
    if (i < B) slice_off[i] = carry + ex;
    atomicMax(&mx, v);
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) { slice_off[B] = carry; stats[0] = carry; stats[1] = mx; }
}

// ------------------------------------------------------------------ union-find over runs
__device__ __forceinline__ int find_root(const int* L, int a) {
  int p = L[a];
  while (p != a) { a = p; p = L[a]; }
  return a;
}
__device__ __forceinline__ void unite(int* L, int a, int b) {
  while (true) {
    a = find_root(L, a);
    b = find_root(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }   // the smaller (raster-earlier) run becomes the root
    const int old = atomicMin(&L[b], a);
    if (old == b) return;
    b = old;
  }
}

// each run links to the equal-valued runs of the previous row it touches (8-connectivity)
__global__ void runs_cc_merge_kernel(const int* __restrict__ row_ptr, const int2* __restrict__ run_yx,
                                     const int* __restrict__ run_x1, const int* __restrict__ run_val,
                                     const int* __restrict__ slice_off, int h, int* __restrict__ L) {
  const int b = blockIdx.y;
  const int so = slice_off[b], n = slice_off[b + 1] - so;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = so + i;
  const int2 yx = run_yx[r];
  if (yx.x == 0) return;
  const int x0 = yx.y, x1 = run_x1[r], v = run_val[r];
  const int* rp = row_ptr + static_cast<long long>(b) * (h + 1);
  const int p0 = so + rp[yx.x - 1], p1 = so + rp[yx.x];
  for (int q = p0; q < p1; ++q) {
    const int qx1 = run_x1[q];
    if (qx1 < x0) continue;            // ends left of x0 - 1
    const int qx0 = run_yx[q].y;
    if (qx0 > x1) break;               // starts right of the pixel after this run's last
    if (run_val[q] == v) unite(L, r, q);
  }
}

// one CTA per slice: flatten, number the roots in raster order (1-based), propagate
__global__ void __launch_bounds__(1024)
runs_number_kernel(int* __restrict__ L, const int* __restrict__ slice_off, int* __restrict__ run_cc,
                   int* __restrict__ n_cc) {
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int carry;
  const int b = blockIdx.x;
  const int so = slice_off[b], n = slice_off[b + 1] - so;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int s0 = 0; s0 < n; s0 += 1024) {
    const int i = s0 + threadIdx.x;
    int is_root = 0;
    if (i < n) {
      const int r = so + i;
      const int root = find_root(L, r);
      if (root != r) L[r] = root;
      is_root = root == r;
    }
    int ex, total;
    Scan(tmp).ExclusiveSum(is_root, ex, total);
    if (is_root) run_cc[so + i] = carry + ex + 1;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += 1024) {
    const int r = so + i;
    const int root = L[r];
    if (root != r) run_cc[r] = run_cc[root];
  }
  if (threadIdx.x == 0) n_cc[b] = carry;
}

// table [B][cap][5] = area, y0, x0, y1, x1 (half-open), pre-initialised by table_init_kernel
__global__ void table_init_kernel(int* __restrict__ table, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int f = static_cast<int>(i % 5);
  table[i] = (f == 1 || f == 2) ? 0x7fffffff : 0;
}
__global__ void runs_stats_kernel(const int2* __restrict__ run_yx, const int* __restrict__ run_x1,
                                  const int* __restrict__ run_cc, const int* __restrict__ slice_off,
                                  int cap, int* __restrict__ table) {
  const int b = blockIdx.y;
  const int so = slice_off[b], n = slice_off[b + 1] - so;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = so + i;
  const int id = run_cc[r];
  if (id <= 0 || id > cap) return;
  const int2 yx = run_yx[r];
  const int x1 = run_x1[r];
  int* t = table + (static_cast<long long>(b) * cap + (id - 1)) * 5;
  atomicAdd(&t[0], x1 - yx.y);
  atomicMin(&t[1], yx.x);
  atomicMin(&t[2], yx.y);
  atomicMax(&t[3], yx.x + 1);
  atomicMax(&t[4], x1);
}

// ------------------------------------------------------------------ adjacent-slice overlaps
constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
__device__ __forceinline__ bool hash_add(unsigned long long* keys, int* vals, unsigned long long cap_mask,
                                         unsigned long long key, int count) {
  unsigned long long slot = mix64(key) & cap_mask;
  for (unsigned long long probe = 0; probe <= cap_mask; ++probe) {
    const unsigned long long cur = keys[slot];
    if (cur == key) { atomicAdd(&vals[slot], count); return true; }
    if (cur == EMPTY_KEY) {
      const unsigned long long old = atomicCAS(&keys[slot], EMPTY_KEY, key);
      if (old == EMPTY_KEY || old == key) { atomicAdd(&vals[slot], count); return true; }
    }
    slot = (slot + 1) & cap_mask;
  }
  return false;
}
// slice b (>= 1) against slice b - 1 of the same run set; key = (key_s0 + b) << 40 | prev << 20 | cur.
// A run accumulates its overlap per previous component locally (runs of one previous component are
// usually consecutive) and touches the table once per (prev, cur) change.
__global__ void runs_overlap_kernel(const int* __restrict__ row_ptr, const int2* __restrict__ run_yx,
                                    const int* __restrict__ run_x1, const int* __restrict__ run_cc,
                                    const int* __restrict__ slice_off, int h, int key_s0,
                                    unsigned long long* __restrict__ keys, int* __restrict__ vals,
                                    unsigned long long cap_mask, int* __restrict__ overflow) {
  const int b = blockIdx.y + 1;
  const int so = slice_off[b], n = slice_off[b + 1] - so;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = so + i;
  const int2 yx = run_yx[r];
  const int x0 = yx.y, x1 = run_x1[r], c = run_cc[r];
  const int sp = slice_off[b - 1];
  const int* rp = row_ptr + static_cast<long long>(b - 1) * (h + 1);
  const int p0 = sp + rp[yx.x], p1 = sp + rp[yx.x + 1];
  int acc_q = 0, acc = 0;
  bool ok = true;
  for (int q = p0; q < p1; ++q) {
    const int qx1 = run_x1[q];
    if (qx1 <= x0) continue;
    const int qx0 = run_yx[q].y;
    if (qx0 >= x1) break;
    const int ov = min(x1, qx1) - max(x0, qx0);
    const int pc = run_cc[q];
    if (pc != acc_q) {
      if (acc > 0)
        ok &= hash_add(keys, vals, cap_mask, (static_cast<unsigned long long>(key_s0 + b) << 40) |
                       (static_cast<unsigned long long>(acc_q) << 20) | static_cast<unsigned long long>(c), acc);
      acc_q = pc; acc = 0;
    }
    acc += ov;
  }
  if (acc > 0)
    ok &= hash_add(keys, vals, cap_mask, (static_cast<unsigned long long>(key_s0 + b) << 40) |
                   (static_cast<unsigned long long>(acc_q) << 20) | static_cast<unsigned long long>(c), acc);
  if (!ok) atomicExch(overflow, 1);
}

// ------------------------------------------------------------------ paint
// label of a run of slice s: lut[s][cc] (lut == nullptr: the component id itself plus `add`)
__device__ __forceinline__ int run_label(const int* __restrict__ lut, int lut_stride, int s, int cc, int add) {
  if (cc <= 0) return 0;
  if (lut == nullptr) return cc + add;
  return (cc < lut_stride) ? lut[static_cast<long long>(s) * lut_stride + cc] : 0;
}

// xy / xz planes (rows contiguous in the destination): one warp per output row walks the row's
// runs once and stores the whole row (zeros included) with 16-byte vectors.
template <int VEC>
__global__ void __launch_bounds__(256)
paint_rows_kernel(const int* __restrict__ row_ptr, const int2* __restrict__ run_yx,
                  const int* __restrict__ run_x1, const int* __restrict__ run_cc,
                  const int* __restrict__ slice_off, const int* __restrict__ lut, int lut_stride, int add,
                  int b0, int h, int w, int* __restrict__ dst, long long stride_s, long long stride_y) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = b0 + blockIdx.y;
  if (row >= h) return;
  const int so = slice_off[b];
  const int* rp = row_ptr + static_cast<long long>(b) * (h + 1);
  int cur = so + rp[row];
  const int r1 = so + rp[row + 1];
  int* out = dst + b * stride_s + row * stride_y;
  for (int xs = 0; xs < w; xs += 128) {
    int v[4] = {0, 0, 0, 0};
    for (int r = cur; r < r1; ++r) {
      const int rx0 = run_yx[r].y;
      if (rx0 >= xs + 128) break;
      const int rx1 = run_x1[r];
      const int lab = run_label(lut, lut_stride, b, run_cc[r], add);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = VEC ? (xs + 4 * lane + j) : (xs + lane + 32 * j);
        if (x >= rx0 && x < rx1) v[j] = lab;
      }
      if (rx1 <= xs + 128) cur = r + 1;
    }
    if (VEC) {
      const int x = xs + 4 * lane;
      if (x + 3 < w) *reinterpret_cast<int4*>(out + x) = make_int4(v[0], v[1], v[2], v[3]);
      else
        for (int j = 0; j < 4; ++j) if (x + j < w) out[x + j] = v[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int x = xs + lane + 32 * j; if (x < w) out[x] = v[j]; }
    }
  }
}

// yz plane: slice index = x of the volume, slice row = z, slice column = y. One CTA gathers a
// (128 columns) x (32 slices) tile of row z through shared memory and stores 128-byte lines.
constexpr int YZ_TY = 128;
__global__ void __launch_bounds__(256)
paint_yz_kernel(const int* __restrict__ row_ptr, const int2* __restrict__ run_yx,
                const int* __restrict__ run_x1, const int* __restrict__ run_cc,
                const int* __restrict__ slice_off, const int* __restrict__ lut, int lut_stride, int add,
                int nb, int h, int w, int* __restrict__ dst, long long HW, int Wv) {
  __shared__ int tile[YZ_TY][33];
  const int z = blockIdx.y, x0 = blockIdx.x * 32, y0 = blockIdx.z * YZ_TY;
  for (int i = threadIdx.x; i < YZ_TY * 33; i += blockDim.x) (&tile[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int sb = warp; sb < 32; sb += 8) {
    const int b = x0 + sb;
    if (b >= nb) break;
    const int so = slice_off[b];
    const int* rp = row_ptr + static_cast<long long>(b) * (h + 1);
    const int r0 = so + rp[z], r1 = so + rp[z + 1];
    for (int r = r0 + lane; r < r1; r += 32) {
      const int c0 = max(run_yx[r].y, y0), c1 = min(run_x1[r], y0 + YZ_TY);
      if (c0 >= c1) continue;
      const int lab = run_label(lut, lut_stride, b, run_cc[r], add);
      for (int c = c0; c < c1; ++c) tile[c - y0][sb] = lab;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < YZ_TY * 32; i += blockDim.x) {
    const int col = i >> 5, sx = i & 31;
    if (y0 + col < w && x0 + sx < nb)
      dst[z * HW + static_cast<long long>(y0 + col) * Wv + x0 + sx] = tile[col][sx];
  }
}

}  // namespace runs

// ------------------------------------------------------------------------------ launchers
extern "C" {

int be_group_flags(const uint8_t* hard, const float* off, const int* centers, int cap,
                   const int* counts, int B, int H, int W, int scale, int step, int* cells, int* present,
                   cudaStream_t stream) {
  if (scale < 1 || step < 1 || H % scale || W % scale || W % 4) return be_set_error("group_flags: bad geometry");
  const int n = (H / scale) * (W / scale);
  dim3 grid((n + 255) / 256, B);
  runs::group_flags_kernel<<<grid, 256, 0, stream>>>(hard, off, centers, cap, counts, H, W, scale, step, cells, present);
  return be_check_launch("group_flags_kernel");
}

// area[b] = number of set pixels of hard[b] (H x W uint8, H * W % 16 == 0): the stuff-area test of
// merge_semantic_and_instance (postprocess.py:283-294) for semantic-only planes. area must be zero.
int be_slice_area(const uint8_t* hard, int B, int H, int W, int* area, cudaStream_t stream) {
  const long long n = 1LL * H * W;
  if (n % 16 || (reinterpret_cast<uintptr_t>(hard) & 15)) return be_set_error("slice_area: H * W must be a multiple of 16");
  dim3 grid(static_cast<unsigned>((n / 16 + 1023) / 1024), B);
  runs::slice_area_kernel<<<grid, 256, 0, stream>>>(hard, n / 16, area);
  return be_check_launch("slice_area_kernel");
}

// counts: workspace [2 * B * chunks] int32, chunks = ceil(h * ceil(w/16) / 256). Leaves the scanned
// chunk offsets in `counts`, n_runs[B], slice_off[B+1], stats[2] = {total runs, max runs per slice}
// and row_ptr[b][h] = n_runs[b]; the caller reads stats, allocates the run arrays and calls
// be_rowruns_write with the same arguments.
int be_rowruns_count(const uint8_t* hard, int* cells, const int* newid, int B, int H, int W,
                     int h, int w, int scale, int cap, int void_label, int lo, int hi, int* counts,
                     int* n_runs, int* slice_off, int* stats, int* row_ptr, cudaStream_t stream) {
  if (W % runs::RR_PX || (reinterpret_cast<uintptr_t>(hard) & 15) || scale < 1 || H % scale || W % scale)
    return be_set_error("rowruns: padded width must be a multiple of 16 (and of the cell size)");
  if (scale == 4 && ((reinterpret_cast<uintptr_t>(cells) & 15) || (W / 4) % 4))
    return be_set_error("rowruns: cell rows must be 16-byte aligned");
  // cells: centre ids -> final panoptic values, in place (be_rowruns_write expects them resolved)
  const long long per_slice = 1LL * (H / scale) * (W / scale), total = per_slice * B;
  runs::resolve_cells_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      cells, newid, per_slice, cap, void_label, lo, hi, total);
  runs::Src s{hard, cells, H, W, h, w, scale, runs::cls_filter(void_label, lo, hi)};
  const int gpr = (w + runs::RR_PX - 1) / runs::RR_PX;
  const int chunks = (h * gpr + runs::RR_THREADS - 1) / runs::RR_THREADS;
  runs::rowruns_kernel<0><<<dim3(chunks, B), runs::RR_THREADS, 0, stream>>>(s, gpr, chunks, counts, B, nullptr, nullptr,
                                                                          nullptr, nullptr, nullptr, nullptr);
  runs::rowruns_scan_kernel<<<B, 1024, 0, stream>>>(counts, B, chunks, n_runs, row_ptr, h);
  runs::slice_offsets_kernel<<<1, 1024, 0, stream>>>(n_runs, B, slice_off, stats);
  return be_check_launch("rowruns count kernels");
}

int be_rowruns_write(const uint8_t* hard, const int* cells, const int* newid, int B, int H, int W,
                     int h, int w, int scale, int cap, int void_label, int lo, int hi, int* counts,
                     const int* slice_off, int* row_ptr, int* run_yx, int* run_x1, int* run_val,
                     int* L, cudaStream_t stream) {
  (void)newid; (void)cap;
  runs::Src s{hard, cells, H, W, h, w, scale, runs::cls_filter(void_label, lo, hi)};
  const int gpr = (w + runs::RR_PX - 1) / runs::RR_PX;
  const int chunks = (h * gpr + runs::RR_THREADS - 1) / runs::RR_THREADS;
  runs::rowruns_kernel<1><<<dim3(chunks, B), runs::RR_THREADS, 0, stream>>>(
      s, gpr, chunks, counts, B, slice_off, row_ptr, reinterpret_cast<int2*>(run_yx), run_x1, run_val, L);
  return be_check_launch("rowruns write kernel");
}

// union-find over the runs (L was initialised by be_rowruns_write); run_cc <- raster-order
// component id (1-based), n_cc[b] <- components of slice b
int be_runs_cc(const int* row_ptr, const int* run_yx, const int* run_x1, const int* run_val,
               const int* slice_off, int B, int h, int max_runs, int* L, int* run_cc, int* n_cc,
               cudaStream_t stream) {
  if (max_runs > 0) {
    dim3 grid((max_runs + 255) / 256, B);
    runs::runs_cc_merge_kernel<<<grid, 256, 0, stream>>>(row_ptr, reinterpret_cast<const int2*>(run_yx), run_x1,
                                                         run_val, slice_off, h, L);
  }
  runs::runs_number_kernel<<<B, 1024, 0, stream>>>(L, slice_off, run_cc, n_cc);
  return be_check_launch("runs_cc kernels");
}

int be_runs_stats(const int* run_yx, const int* run_x1, const int* run_cc, const int* slice_off,
                  int B, int max_runs, int cap, int* table, cudaStream_t stream) {
  const long long tn = static_cast<long long>(B) * cap * 5;
  if (tn > 0) runs::table_init_kernel<<<static_cast<unsigned>((tn + 255) / 256), 256, 0, stream>>>(table, tn);
  if (max_runs > 0 && cap > 0) {
    dim3 grid((max_runs + 255) / 256, B);
    runs::runs_stats_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const int2*>(run_yx), run_x1, run_cc,
                                                      slice_off, cap, table);
  }
  return be_check_launch("runs_stats kernels");
}

int be_runs_overlap(const int* row_ptr, const int* run_yx, const int* run_x1, const int* run_cc,
                    const int* slice_off, int B, int h, int max_runs, int key_s0,
                    unsigned long long* keys, int* vals, unsigned long long cap, int* overflow,
                    cudaStream_t stream) {
  if (cap & (cap - 1)) return be_set_error("hash capacity must be a power of two");
  if (B < 2 || max_runs <= 0) return 0;
  dim3 grid((max_runs + 255) / 256, B - 1);
  runs::runs_overlap_kernel<<<grid, 256, 0, stream>>>(row_ptr, reinterpret_cast<const int2*>(run_yx), run_x1, run_cc,
                                                      slice_off, h, key_s0, keys, vals, cap - 1, overflow);
  return be_check_launch("runs_overlap_kernel");
}

// Paints slices [b0, b0 + nb) of the run set into dst. Element (slice b, row y, column x) lands at
// b*stride_s + y*stride_y + x*stride_x (xy: HW, W, 1 | xz: W, HW, 1 | yz: 1, HW, W of the
// destination volume); every element of those slices is written (background = 0). lut [.][lut_stride]
// indexed by slice b (NULL: label = component id + add).
int be_runs_paint(const int* row_ptr, const int* run_yx, const int* run_x1, const int* run_cc,
                  const int* slice_off, const int* lut, int lut_stride, int add, int b0, int nb, int h,
                  int w, int* dst, long long stride_s, long long stride_y, long long stride_x,
                  cudaStream_t stream) {
  if (nb <= 0) return 0;
  const int2* yx = reinterpret_cast<const int2*>(run_yx);
  if (stride_x == 1) {
    const bool vec = (stride_s % 4 == 0) && (stride_y % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    dim3 grid((h + 7) / 8, nb);
    if (vec)
      runs::paint_rows_kernel<1><<<grid, 256, 0, stream>>>(row_ptr, yx, run_x1, run_cc, slice_off, lut, lut_stride,
                                                          add, b0, h, w, dst, stride_s, stride_y);
    else
      runs::paint_rows_kernel<0><<<grid, 256, 0, stream>>>(row_ptr, yx, run_x1, run_cc, slice_off, lut, lut_stride,
                                                          add, b0, h, w, dst, stride_s, stride_y);
    return be_check_launch("paint_rows_kernel");
  }
  if (stride_s != 1 || b0 != 0) return be_set_error("runs_paint: unsupported destination layout");
  dim3 grid((nb + 31) / 32, h, (w + runs::YZ_TY - 1) / runs::YZ_TY);
  runs::paint_yz_kernel<<<grid, 256, 0, stream>>>(row_ptr, yx, run_x1, run_cc, slice_off, lut, lut_stride, add, nb, h,
                                                  w, dst, stride_y, static_cast<int>(stride_x));
  return be_check_launch("paint_yz_kernel");
}

}  // extern "C"
