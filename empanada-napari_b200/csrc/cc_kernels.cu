// Connected components, per-component tables, adjacent-slice overlap tables, LUT relabel and
// run-length extraction (pieces 5 and 6b of the hot path), batched over slices.
// Restates on the GPU what the reference does per slice on the CPU in
//   * connected_components / pan_seg_to_rle_seg (empanada/inference/rle.py:18-86):
//       8-connected components of EQUAL-valued pixels inside one class range, numbered in
//       raster order of their first pixel (skimage.measure.label semantics), bbox + area;
//   * rle_intersection between consecutive slices (empanada/array_utils.py:375-407): here a
//       sparse (cc_prev, cc_cur) -> overlapping-pixel-count table built in one pass;
//   * rle_encode (array_utils.py:213-239): maximal runs of equal label over FLAT indices
//       (runs may continue across row ends, exactly as the reference's flat-index encoding).
// Bandwidth-bound integer kernels: 4-byte coalesced accesses, atomics only at run heads.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdint>

#include "common.cuh"

namespace cc {

// ------------------------------------------------------------------ union-find labelling
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
    if (a > b) { const int t = a; a = b; b = t; }  // a < b: smaller index becomes the root
    const int old = atomicMin(&L[b], a);
    if (old == b) return;
    b = old;
  }
}

// value of pixel p restricted to the class range [lo, hi) (rle.py:60-66)
__device__ __forceinline__ int cls_val(const int* pan, long long i, int lo, int hi) {
  const int v = pan[i];
  return (v >= lo && v < hi && v != 0) ? v : 0;
}

// Pass 1 (row runs): every pixel of a horizontal run of equal class value points at the run's
// first pixel, computed with a segmented max-scan per 1024-pixel row chunk (no pointer chasing).
// Runs are cut at chunk borders; pass 2 stitches them.
constexpr int ROWCHUNK = 1024;
__global__ void __launch_bounds__(ROWCHUNK)
cc_init_kernel(const int* __restrict__ pan, int* __restrict__ L, int h, int w, int lo, int hi) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * ROWCHUNK + threadIdx.x;
  const long long base = static_cast<long long>(b) * h * w;
  const int p = y * w + x;
  const int v = (x < w) ? cls_val(pan, base + p, lo, hi) : 0;
  const int vl = (x < w && threadIdx.x > 0) ? cls_val(pan, base + p - 1, lo, hi) : 0;
  if (!__syncthreads_or(v != 0)) {  // nothing of this class in the row chunk
    if (x < w) L[base + p] = -1;
    return;
  }
  const bool head = v != 0 && (threadIdx.x == 0 || vl != v);
  // start index of the run containing this pixel = max head position <= x
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  const unsigned below = heads & (0xffffffffu >> (31 - lane));  // heads at lanes <= lane
  int start = below ? (warp * 32 + 31 - __clz(below)) : -1;     // chunk-relative, -1: none in warp
  __shared__ int warp_last[32];
  if (lane == 31) warp_last[warp] = start;
  __syncthreads();
  if (warp == 0) {
    int s = warp_last[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s = max(s, u);
    }
    warp_last[lane] = s;  // inclusive running max of "last head seen"
  }
  __syncthreads();
  if (start < 0 && warp > 0) start = warp_last[warp - 1];
  if (x < w) L[base + p] = v ? (y * w + blockIdx.x * ROWCHUNK + start) : -1;
}

// Pass 2: union run representatives across rows (8-connectivity) and across chunk borders.
// Only the first pixel of a run that touches a given upper run issues the union.
__global__ void cc_merge_kernel(const int* __restrict__ pan, int* __restrict__ L, int h, int w,
                                int lo, int hi) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const long long base = static_cast<long long>(b) * h * w;
  const int p = y * w + x;
  const int v = cls_val(pan, base + p, lo, hi);
  if (!v) return;
  int* Lb = L + base;
  const bool left_same = (x > 0) && (cls_val(pan, base + p - 1, lo, hi) == v);
  const bool right_same = (x + 1 < w) && (cls_val(pan, base + p + 1, lo, hi) == v);
  if (left_same && (x % ROWCHUNK) == 0) unite(Lb, p, p - 1);  // run cut at a chunk border
  if (y > 0) {
    const bool n_same = cls_val(pan, base + p - w, lo, hi) == v;
    const bool nw_same = (x > 0) && (cls_val(pan, base + p - w - 1, lo, hi) == v);
    const bool ne_same = (x + 1 < w) && (cls_val(pan, base + p - w + 1, lo, hi) == v);
    if (n_same) {
      // the pixel to the left already linked this pair of runs if it sits under the same upper run
      if (!(left_same && nw_same)) unite(Lb, p, p - w);
    } else {
      if (nw_same && !left_same) unite(Lb, p, p - w - 1);
      if (ne_same && !right_same) unite(Lb, p, p - w + 1);
    }
    // N equal and NE equal belong to the same upper run; N not equal but both diagonals equal are
    // two different upper runs, each handled above
    if (n_same == false && nw_same && left_same) { /* handled by the left pixel's N link */ }
  }
}

__global__ void cc_compress_kernel(int* __restrict__ L, long long hw_total, int hw) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= hw_total) return;
  const int v = L[i];
  if (v < 0) return;
  int* Lb = L + (i / hw) * hw;
  L[i] = find_root(Lb, v);
}

// roots per 1024-pixel chunk (raster order) -> chunk_counts[b][chunk]
constexpr int CHUNK = 1024;
__global__ void __launch_bounds__(CHUNK)
cc_count_roots_kernel(const int* __restrict__ L, int hw, int chunks, int* __restrict__ chunk_counts) {
  const int b = blockIdx.y, ch = blockIdx.x;
  const int p = ch * CHUNK + threadIdx.x;
  const bool root = (p < hw) && (L[static_cast<long long>(b) * hw + p] == p);
  const int n = __syncthreads_count(root);
  if (threadIdx.x == 0) chunk_counts[b * chunks + ch] = n;
}

// one CTA per slice: exclusive scan of chunk counts (in place), total -> n_cc[b]
__global__ void __launch_bounds__(1024)
cc_scan_chunks_kernel(int* __restrict__ chunk_counts, int chunks, int* __restrict__ n_cc) {
  typedef cub::BlockScan<int, 1024> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int carry;
  int* c = chunk_counts + static_cast<long long>(blockIdx.x) * chunks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int s = 0; s < chunks; s += 1024) {
    const int i = s + threadIdx.x;
    const int v = (i < chunks) ? c[i] : 0;
    int ex, total;
    Scan(tmp).ExclusiveSum(v, ex, total);
    if (i < chunks) c[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_cc[blockIdx.x] = carry;
}

// roots receive their raster-order id (1-based) in `out`
__global__ void __launch_bounds__(CHUNK)
cc_number_roots_kernel(const int* __restrict__ L, int hw, int chunks,
                       const int* __restrict__ chunk_offsets, int* __restrict__ out) {
  typedef cub::BlockScan<int, CHUNK> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const int b = blockIdx.y, ch = blockIdx.x;
  const int p = ch * CHUNK + threadIdx.x;
  const long long base = static_cast<long long>(b) * hw;
  const int root = (p < hw) && (L[base + p] == p);
  int ex;
  Scan(tmp).ExclusiveSum(root, ex);
  if (root) out[base + p] = chunk_offsets[b * chunks + ch] + ex + 1;
}

__global__ void cc_apply_ids_kernel(const int* __restrict__ L, int* __restrict__ out,
                                    long long hw_total, int hw) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= hw_total) return;
  const int r = L[i];
  if (r < 0) { out[i] = 0; return; }
  const long long base = (i / hw) * hw;
  if (base + r != i) out[i] = out[base + r];
}

// per-component area + bbox; one atomic group per horizontal run
// table layout per slice: [cap][5] = area, y0, x0, y1, x1 (half-open); pre-initialised.
__global__ void cc_table_init_kernel(int* __restrict__ table, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int f = static_cast<int>(i % 5);
  table[i] = (f == 1 || f == 2) ? 0x7fffffff : 0;
}
__global__ void cc_stats_kernel(const int* __restrict__ ccimg, int h, int w, int cap,
                                int* __restrict__ table) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const int* row = ccimg + (static_cast<long long>(b) * h + y) * w;
  const int id = row[x];
  if (id == 0 || id > cap) return;
  if (x > 0 && row[x - 1] == id) return;  // not a run head
  int e = x + 1;
  while (e < w && row[e] == id) ++e;
  int* t = table + (static_cast<long long>(b) * cap + (id - 1)) * 5;
  atomicAdd(&t[0], e - x);
  atomicMin(&t[1], y);
  atomicMin(&t[2], x);
  atomicMax(&t[3], y + 1);
  atomicMax(&t[4], e);
}

// ------------------------------------------------------------------ quad variants (w % 4 == 0)
// One thread = four consecutive pixels (16-byte vectors). Component images are mostly background,
// so every kernel leaves a background quad after its first vector load; passes are fused where
// the second pass only needs what the first one has in registers.
__device__ __forceinline__ int4 ldq(const int* p) { return *reinterpret_cast<const int4*>(p); }

// cc_merge over quads: a quad without any pixel of the class has nothing to union
__global__ void cc_merge_v4_kernel(const int* __restrict__ pan, int* __restrict__ L, int h, int w,
                                   int lo, int hi) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (x0 >= w) return;
  const long long base = static_cast<long long>(b) * h * w;
  const int4 q = ldq(pan + base + static_cast<long long>(y) * w + x0);
  const int vq[4] = {q.x, q.y, q.z, q.w};
  int* Lb = L + base;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int v = (vq[j] >= lo && vq[j] < hi && vq[j] != 0) ? vq[j] : 0;
    if (!v) continue;
    const int x = x0 + j, p = y * w + x;
    const bool left_same = (x > 0) && (cls_val(pan, base + p - 1, lo, hi) == v);
    const bool right_same = (x + 1 < w) && (cls_val(pan, base + p + 1, lo, hi) == v);
    if (left_same && (x % ROWCHUNK) == 0) unite(Lb, p, p - 1);
    if (y > 0) {
      const bool n_same = cls_val(pan, base + p - w, lo, hi) == v;
      const bool nw_same = (x > 0) && (cls_val(pan, base + p - w - 1, lo, hi) == v);
      const bool ne_same = (x + 1 < w) && (cls_val(pan, base + p - w + 1, lo, hi) == v);
      if (n_same) {
        if (!(left_same && nw_same)) unite(Lb, p, p - w);
      } else {
        if (nw_same && !left_same) unite(Lb, p, p - w - 1);
        if (ne_same && !right_same) unite(Lb, p, p - w + 1);
      }
    }
  }
}

// path compression fused with the per-chunk root count (CHUNK = 1024 pixels = 256 threads x 4)
__global__ void __launch_bounds__(CHUNK / 4)
cc_compress_count_v4_kernel(int* __restrict__ L, int hw, int chunks, int* __restrict__ chunk_counts) {
  __shared__ int total;
  const int b = blockIdx.y, ch = blockIdx.x;
  const int p0 = ch * CHUNK + 4 * threadIdx.x;
  int* Lb = L + static_cast<long long>(b) * hw;
  if (threadIdx.x == 0) total = 0;
  int roots = 0;
  if (p0 < hw) {
    const int4 q = ldq(Lb + p0);
    if ((q.x & q.y & q.z & q.w) >= 0) {   // some lane is not -1
      const int vq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (vq[j] < 0) continue;
        const int r = find_root(Lb, vq[j]);
        if (r != vq[j]) Lb[p0 + j] = r;
        roots += (r == p0 + j);
      }
    }
  }
  __syncthreads();
  if (__syncthreads_or(roots != 0)) {
    if (roots) atomicAdd(&total, roots);
    __syncthreads();
  }
  if (threadIdx.x == 0) chunk_counts[b * chunks + ch] = total;
}

// raster-order ids of the roots; chunks without a root (almost all) return before reading L
__global__ void __launch_bounds__(CHUNK / 4)
cc_number_roots_v4_kernel(const int* __restrict__ L, int hw, int chunks,
                          const int* __restrict__ chunk_offsets, const int* __restrict__ n_cc,
                          int* __restrict__ out) {
  typedef cub::BlockScan<int, CHUNK / 4> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const int b = blockIdx.y, ch = blockIdx.x;
  const int off = chunk_offsets[b * chunks + ch];
  const int next = (ch + 1 < chunks) ? chunk_offsets[b * chunks + ch + 1] : n_cc[b];
  if (next == off) return;
  const int p0 = ch * CHUNK + 4 * threadIdx.x;
  const long long base = static_cast<long long>(b) * hw;
  int root[4] = {0, 0, 0, 0};
  if (p0 < hw) {
    const int4 q = ldq(L + base + p0);
    root[0] = q.x == p0; root[1] = q.y == p0 + 1; root[2] = q.z == p0 + 2; root[3] = q.w == p0 + 3;
  }
  const int nr = root[0] + root[1] + root[2] + root[3];
  int ex;
  Scan(tmp).ExclusiveSum(nr, ex);
  if (nr) {
    int id = off + ex + 1;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (root[j]) out[base + p0 + j] = id++;
  }
}

// final ids for every pixel fused with the per-component area / bbox statistics
__global__ void cc_apply_stats_v4_kernel(const int* __restrict__ L, int* __restrict__ out, int h, int w,
                                         int cap, int* __restrict__ table) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (x0 >= w) return;
  const long long base = static_cast<long long>(b) * h * w;
  const long long i0 = base + static_cast<long long>(y) * w + x0;
  const int4 q = ldq(L + i0);
  if (q.x < 0 && q.y < 0 && q.z < 0 && q.w < 0) {
    *reinterpret_cast<int4*>(out + i0) = make_int4(0, 0, 0, 0);
    return;
  }
  const int rq[4] = {q.x, q.y, q.z, q.w};
  int id[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) id[j] = (rq[j] < 0) ? 0 : out[base + rq[j]];   // roots were numbered already
  *reinterpret_cast<int4*>(out + i0) = make_int4(id[0], id[1], id[2], id[3]);
  if (table == nullptr) return;
  const int before = (x0 > 0 && rq[0] >= 0) ? L[i0 - 1] : -1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (rq[j] < 0 || id[j] > cap) continue;
    const int pv = (j == 0) ? before : rq[j - 1];
    if (pv == rq[j]) continue;                  // not a run head (same root <=> same component)
    const int x = x0 + j;
    int e = x + 1;
    const int* Lrow = L + base + static_cast<long long>(y) * w;
    while (e < w && Lrow[e] == rq[j]) ++e;
    int* t = table + (static_cast<long long>(b) * cap + (id[j] - 1)) * 5;
    atomicAdd(&t[0], e - x);
    atomicMin(&t[1], y);
    atomicMin(&t[2], x);
    atomicMax(&t[3], y + 1);
    atomicMax(&t[4], e);
  }
}

// relabel through the LUT, rows contiguous in the destination (xy / xz planes)
__global__ void relabel_v4_kernel(const int* __restrict__ src, int h, int w, int s0,
                                  const int* __restrict__ lut, int lut_stride, int* __restrict__ dst,
                                  long long stride_s, long long stride_y) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (x0 >= w) return;
  const int4 q = ldq(src + (static_cast<long long>(b) * h + y) * w + x0);
  const int s = s0 + b;
  int4 o = make_int4(0, 0, 0, 0);
  if ((q.x | q.y | q.z | q.w) != 0) {
    const int* l = lut + static_cast<long long>(s) * lut_stride;
    if (q.x > 0 && q.x < lut_stride) o.x = l[q.x];
    if (q.y > 0 && q.y < lut_stride) o.y = l[q.y];
    if (q.z > 0 && q.z < lut_stride) o.z = l[q.z];
    if (q.w > 0 && q.w < lut_stride) o.w = l[q.w];
  }
  *reinterpret_cast<int4*>(dst + s * stride_s + y * stride_y + x0) = o;
}

// ------------------------------------------------------------------ adjacent-slice overlaps
// open-addressing table: key = slice(24) | prev_cc(20) | cur_cc(20); val = pixel count
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

// ccimg: plane buffer [N][h][w]; processes slices [s0, s1), pairing slice s with s-1 (s >= 1)
__global__ void pair_overlap_kernel(const int* __restrict__ ccimg, int h, int w, int s0,
                                    unsigned long long* __restrict__ keys, int* __restrict__ vals,
                                    unsigned long long cap_mask, int* __restrict__ overflow) {
  const int s = s0 + blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w || s < 1) return;
  const int* cur = ccimg + (static_cast<long long>(s) * h + y) * w;
  const int* prv = cur - static_cast<long long>(h) * w;
  const int c = cur[x], q = prv[x];
  if (c == 0 || q == 0) return;
  if (x > 0 && cur[x - 1] == c && prv[x - 1] == q) return;  // not the head of this pair run
  int e = x + 1;
  while (e < w && cur[e] == c && prv[e] == q) ++e;
  const unsigned long long key = (static_cast<unsigned long long>(s) << 40) |
                                 (static_cast<unsigned long long>(q) << 20) |
                                 static_cast<unsigned long long>(c);
  if (!hash_add(keys, vals, cap_mask, key, e - x)) atomicExch(overflow, 1);
}

// quad variant: a quad with no labelled pixel in either slice ends after two vector loads
__global__ void pair_overlap_v4_kernel(const int* __restrict__ ccimg, int h, int w, int s0,
                                       unsigned long long* __restrict__ keys, int* __restrict__ vals,
                                       unsigned long long cap_mask, int* __restrict__ overflow) {
  const int s = s0 + blockIdx.z, y = blockIdx.y;
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (x0 >= w || s < 1) return;
  const int* cur = ccimg + (static_cast<long long>(s) * h + y) * w;
  const int* prv = cur - static_cast<long long>(h) * w;
  const int4 cq = *reinterpret_cast<const int4*>(cur + x0);
  if ((cq.x | cq.y | cq.z | cq.w) == 0) return;
  const int4 pq = *reinterpret_cast<const int4*>(prv + x0);
  if ((pq.x | pq.y | pq.z | pq.w) == 0) return;
  const int cv[4] = {cq.x, cq.y, cq.z, cq.w}, pv[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cv[j], q = pv[j];
    if (c == 0 || q == 0) continue;
    const int x = x0 + j;
    const int cl = (j == 0) ? (x > 0 ? cur[x - 1] : 0) : cv[j - 1];
    const int ql = (j == 0) ? (x > 0 ? prv[x - 1] : 0) : pv[j - 1];
    if (x > 0 && cl == c && ql == q) continue;  // not the head of this pair run
    int e = x + 1;
    while (e < w && cur[e] == c && prv[e] == q) ++e;
    const unsigned long long key = (static_cast<unsigned long long>(s) << 40) |
                                   (static_cast<unsigned long long>(q) << 20) |
                                   static_cast<unsigned long long>(c);
    if (!hash_add(keys, vals, cap_mask, key, e - x)) atomicExch(overflow, 1);
  }
}

// generic compaction of a (key, count) table into dense arrays (order unspecified)
__global__ void hash_compact_kernel(const unsigned long long* __restrict__ keys,
                                    const int* __restrict__ vals, unsigned long long cap,
                                    unsigned long long* __restrict__ out_keys,
                                    int* __restrict__ out_vals, int out_cap,
                                    int* __restrict__ cursor) {
  const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= cap) return;
  const unsigned long long k = keys[i];
  if (k == EMPTY_KEY) return;
  const int pos = atomicAdd(cursor, 1);
  if (pos < out_cap) { out_keys[pos] = k; out_vals[pos] = vals[i]; }
}

// ------------------------------------------------------------------ LUT relabel into a volume
// src: cc image of slices [s0, s0+B) laid out [B][h][w]; lut: [N][lut_stride] (entry 0 unused);
// dst: (D,H,W)-ordered volume; element (slice s, row y, col x) lands at
//      s*stride_s + y*stride_y + x*stride_x  (xy: HW, W, 1 | xz: W, HW, 1 | yz: 1, HW, W)
__global__ void relabel_kernel(const int* __restrict__ src, int h, int w, int s0,
                               const int* __restrict__ lut, int lut_stride,
                               int* __restrict__ dst, long long stride_s, long long stride_y,
                               long long stride_x) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const int id = src[(static_cast<long long>(b) * h + y) * w + x];
  const int s = s0 + b;
  int v = 0;
  if (id > 0 && id < lut_stride) v = lut[static_cast<long long>(s) * lut_stride + id];
  dst[s * stride_s + y * stride_y + x * stride_x] = v;
}

// tiled transpose variant for the yz plane (stride_x == W, stride_s == 1): a 32x32 tile of
// (slice, x) is staged through shared memory so both the read and the write are coalesced.
__global__ void relabel_yz_kernel(const int* __restrict__ src, int h, int w, int s0, int nb,
                                  const int* __restrict__ lut, int lut_stride,
                                  int* __restrict__ dst, long long HW, int Wvol) {
  __shared__ int tile[32][33];
  const int y = blockIdx.y;                 // row of the slice = z of the volume
  const int x0 = blockIdx.x * 32;           // col of the slice = y of the volume
  const int b0 = blockIdx.z * 32;           // slice within the batch = x of the volume
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int b = b0 + j, x = x0 + threadIdx.x;
    int v = 0;
    if (b < nb && x < w) {
      const int id = src[(static_cast<long long>(b) * h + y) * w + x];
      if (id > 0 && id < lut_stride) v = lut[static_cast<long long>(s0 + b) * lut_stride + id];
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int x = x0 + j, b = b0 + threadIdx.x;
    if (b < nb && x < w) dst[y * HW + static_cast<long long>(x) * Wvol + (s0 + b)] = tile[threadIdx.x][j];
  }
}

// ------------------------------------------------------------------ run-length extraction
// Runs of equal non-zero label over the FLAT index space of each segment (segment = one slice
// for xy/xz trackers, the whole volume for yz / consensus). Two passes give raster order.
__device__ __forceinline__ bool is_run_head(const int* img, long long i, long long seg_len) {
  const int v = img[i];
  if (v == 0) return false;
  return (i % seg_len == 0) || (img[i - 1] != v);
}
__device__ __forceinline__ bool is_run_tail(const int* img, long long i, long long n, long long seg_len) {
  const int v = img[i];
  if (v == 0) return false;
  return ((i + 1) % seg_len == 0) || (i + 1 >= n) || (img[i + 1] != v);
}
// chunk_counts[c] = heads in chunk c, chunk_counts[chunks + 1 + c] = tails in chunk c. The k-th
// head and the k-th tail (global raster rank) delimit the same run, so no thread walks a run.
// One CTA = one chunk of RUN_CHUNK = 4096 elements, one thread = four consecutive elements read
// as one 16-byte vector; chunks without any label (most of a volume) leave after one vote.
constexpr int RUN_CHUNK = 4096;
constexpr int RUN_THREADS = RUN_CHUNK / 4;
struct Quad { int v[4]; int head[4]; int tail[4]; };
__device__ __forceinline__ bool load_quad(const int* __restrict__ img, long long i0, long long n,
                                          long long seg_len, Quad& q) {
#pragma unroll
  for (int j = 0; j < 4; ++j) { q.v[j] = 0; q.head[j] = 0; q.tail[j] = 0; }
  if (i0 >= n) return false;
  if (i0 + 4 <= n && (reinterpret_cast<uintptr_t>(img) & 15) == 0) {
    const int4 t = __ldg(reinterpret_cast<const int4*>(img + i0));
    q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
  } else {  // tail of the image, or a view that does not start on a 16-byte boundary
    for (int j = 0; j < 4 && i0 + j < n; ++j) q.v[j] = img[i0 + j];
  }
  if ((q.v[0] | q.v[1] | q.v[2] | q.v[3]) == 0) return false;
  const int before = (q.v[0] != 0 && i0 > 0) ? img[i0 - 1] : 0;
  const int after = (q.v[3] != 0 && i0 + 4 < n) ? img[i0 + 4] : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = i0 + j;
    if (q.v[j] == 0 || i >= n) continue;
    const int pv = (j == 0) ? before : q.v[j - 1];
    const int nv = (j == 3) ? after : q.v[j + 1];
    q.head[j] = (i % seg_len == 0) || (pv != q.v[j]);
    q.tail[j] = ((i + 1) % seg_len == 0) || (i + 1 >= n) || (nv != q.v[j]);
  }
  return true;
}
__global__ void __launch_bounds__(RUN_THREADS)
runs_count_kernel(const int* __restrict__ img, long long n, long long seg_len, int chunks,
                  int* __restrict__ chunk_counts) {
  __shared__ int sh[2];
  const long long i0 = blockIdx.x * static_cast<long long>(RUN_CHUNK) + 4LL * threadIdx.x;
  Quad q;
  const bool any = load_quad(img, i0, n, seg_len, q);
  if (threadIdx.x < 2) sh[threadIdx.x] = 0;
  if (!__syncthreads_or(any)) {
    if (threadIdx.x == 0) { chunk_counts[blockIdx.x] = 0; chunk_counts[chunks + 1 + blockIdx.x] = 0; }
    return;
  }
  int nh = q.head[0] + q.head[1] + q.head[2] + q.head[3];
  int nt = q.tail[0] + q.tail[1] + q.tail[2] + q.tail[3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { nh += __shfl_xor_sync(0xffffffffu, nh, o); nt += __shfl_xor_sync(0xffffffffu, nt, o); }
  if ((threadIdx.x & 31) == 0 && (nh | nt)) { atomicAdd(&sh[0], nh); atomicAdd(&sh[1], nt); }
  __syncthreads();
  if (threadIdx.x == 0) { chunk_counts[blockIdx.x] = sh[0]; chunk_counts[chunks + 1 + blockIdx.x] = sh[1]; }
}
// out_label / out_start by head rank, out_len (= tail - start + 1) by tail rank
__global__ void __launch_bounds__(RUN_THREADS)
runs_write_kernel(const int* __restrict__ img, long long n, long long seg_len, int chunks,
                  const long long* __restrict__ chunk_offsets, int* __restrict__ out_label,
                  long long* __restrict__ out_start, long long* __restrict__ out_end, long long out_cap) {
  typedef cub::BlockScan<int, RUN_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const long long i0 = blockIdx.x * static_cast<long long>(RUN_CHUNK) + 4LL * threadIdx.x;
  Quad q;
  const bool any = load_quad(img, i0, n, seg_len, q);
  if (!__syncthreads_or(any)) return;
  const int nh = q.head[0] + q.head[1] + q.head[2] + q.head[3];
  const int nt = q.tail[0] + q.tail[1] + q.tail[2] + q.tail[3];
  int exh, ext;
  Scan(tmp).ExclusiveSum(nh, exh);
  __syncthreads();
  Scan(tmp).ExclusiveSum(nt, ext);
  if (nh) {
    long long pos = chunk_offsets[blockIdx.x] + exh;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q.head[j]) { if (pos < out_cap) { out_label[pos] = q.v[j]; out_start[pos] = i0 + j; } ++pos; }
  }
  if (nt) {
    long long pos = chunk_offsets[chunks + 1 + blockIdx.x] + ext;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q.tail[j]) { if (pos < out_cap) out_end[pos] = i0 + j + 1; ++pos; }
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, unsigned long long v, unsigned long long n) {
  const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace cc

// ------------------------------------------------------------------------------ launchers
extern "C" {

// workspace: L [B*h*w] int32, chunk_counts [B*chunks] int32
int be_cc_label(const int* pan, int B, int h, int w, int lo, int hi, int* L, int* chunk_counts,
                int* cc_out, int* n_cc, int cap, int* table, cudaStream_t stream) {
  const int hw = h * w;
  const long long total = static_cast<long long>(B) * hw;
  dim3 grid((w + 255) / 256, h, B);
  dim3 grid_rows((w + cc::ROWCHUNK - 1) / cc::ROWCHUNK, h, B);
  cc::cc_init_kernel<<<grid_rows, cc::ROWCHUNK, 0, stream>>>(pan, L, h, w, lo, hi);
  const int chunks = (hw + cc::CHUNK - 1) / cc::CHUNK;
  const bool quads = (w % 4 == 0) && (((reinterpret_cast<uintptr_t>(pan) | reinterpret_cast<uintptr_t>(L) |
                                        reinterpret_cast<uintptr_t>(cc_out)) & 15) == 0);
  if (quads) {
    dim3 gridq((w / 4 + 127) / 128, h, B);
    cc::cc_merge_v4_kernel<<<gridq, 128, 0, stream>>>(pan, L, h, w, lo, hi);
    cc::cc_compress_count_v4_kernel<<<dim3(chunks, B), cc::CHUNK / 4, 0, stream>>>(L, hw, chunks, chunk_counts);
    cc::cc_scan_chunks_kernel<<<B, 1024, 0, stream>>>(chunk_counts, chunks, n_cc);
    cc::cc_number_roots_v4_kernel<<<dim3(chunks, B), cc::CHUNK / 4, 0, stream>>>(L, hw, chunks, chunk_counts, n_cc, cc_out);
    if (table != nullptr) {
      const long long tn = static_cast<long long>(B) * cap * 5;
      cc::cc_table_init_kernel<<<static_cast<unsigned>((tn + 255) / 256), 256, 0, stream>>>(table, tn);
    }
    cc::cc_apply_stats_v4_kernel<<<gridq, 128, 0, stream>>>(L, cc_out, h, w, cap, table);
    return be_check_launch("cc_label kernels (quad)");
  }
  cc::cc_merge_kernel<<<grid, 256, 0, stream>>>(pan, L, h, w, lo, hi);
  const unsigned nb = static_cast<unsigned>((total + 255) / 256);
  cc::cc_compress_kernel<<<nb, 256, 0, stream>>>(L, total, hw);
  cc::cc_count_roots_kernel<<<dim3(chunks, B), cc::CHUNK, 0, stream>>>(L, hw, chunks, chunk_counts);
  cc::cc_scan_chunks_kernel<<<B, 1024, 0, stream>>>(chunk_counts, chunks, n_cc);
  cc::cc_number_roots_kernel<<<dim3(chunks, B), cc::CHUNK, 0, stream>>>(L, hw, chunks, chunk_counts, cc_out);
  cc::cc_apply_ids_kernel<<<nb, 256, 0, stream>>>(L, cc_out, total, hw);
  if (table != nullptr) {
    const long long tn = static_cast<long long>(B) * cap * 5;
    cc::cc_table_init_kernel<<<static_cast<unsigned>((tn + 255) / 256), 256, 0, stream>>>(table, tn);
    cc::cc_stats_kernel<<<grid, 256, 0, stream>>>(cc_out, h, w, cap, table);
  }
  return be_check_launch("cc_label kernels");
}

int be_hash_clear(unsigned long long* keys, int* vals, unsigned long long cap, cudaStream_t stream) {
  cc::fill_u64_kernel<<<static_cast<unsigned>((cap + 255) / 256), 256, 0, stream>>>(keys, cc::EMPTY_KEY, cap);
  cudaError_t e = cudaMemsetAsync(vals, 0, sizeof(int) * cap, stream);
  if (e != cudaSuccess) return be_set_error(cudaGetErrorString(e));
  return be_check_launch("hash_clear");
}

int be_pair_overlap(const int* cc_plane, int h, int w, int s0, int s1, unsigned long long* keys,
                    int* vals, unsigned long long cap, int* overflow, cudaStream_t stream) {
  if (cap & (cap - 1)) return be_set_error("hash capacity must be a power of two");
  if (s1 <= s0) return 0;
  if (w % 4 == 0 && (reinterpret_cast<uintptr_t>(cc_plane) & 15) == 0) {
    dim3 gridq((w / 4 + 127) / 128, h, s1 - s0);
    cc::pair_overlap_v4_kernel<<<gridq, 128, 0, stream>>>(cc_plane, h, w, s0, keys, vals, cap - 1, overflow);
    return be_check_launch("pair_overlap_v4_kernel");
  }
  dim3 grid((w + 255) / 256, h, s1 - s0);
  cc::pair_overlap_kernel<<<grid, 256, 0, stream>>>(cc_plane, h, w, s0, keys, vals, cap - 1, overflow);
  return be_check_launch("pair_overlap_kernel");
}

int be_hash_compact(const unsigned long long* keys, const int* vals, unsigned long long cap,
                    unsigned long long* out_keys, int* out_vals, int out_cap, int* cursor,
                    cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(int), stream);
  if (e != cudaSuccess) return be_set_error(cudaGetErrorString(e));
  cc::hash_compact_kernel<<<static_cast<unsigned>((cap + 255) / 256), 256, 0, stream>>>(
      keys, vals, cap, out_keys, out_vals, out_cap, cursor);
  return be_check_launch("hash_compact_kernel");
}

int be_relabel(const int* cc_batch, int B, int h, int w, int s0, const int* lut, int lut_stride,
               int* dst, long long stride_s, long long stride_y, long long stride_x,
               cudaStream_t stream) {
  if (stride_s == 1 && stride_x > 1) {
    // yz plane: x of the slice is y of the volume (stride_x == W), slice index is x of the volume
    dim3 grid((w + 31) / 32, h, (B + 31) / 32);
    cc::relabel_yz_kernel<<<grid, dim3(32, 8), 0, stream>>>(cc_batch, h, w, s0, B, lut, lut_stride,
                                                            dst, stride_y, static_cast<int>(stride_x));
  } else if (stride_x == 1 && w % 4 == 0 && stride_s % 4 == 0 && stride_y % 4 == 0 &&
             ((reinterpret_cast<uintptr_t>(cc_batch) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    dim3 gridq((w / 4 + 127) / 128, h, B);
    cc::relabel_v4_kernel<<<gridq, 128, 0, stream>>>(cc_batch, h, w, s0, lut, lut_stride, dst, stride_s, stride_y);
  } else {
    dim3 grid((w + 255) / 256, h, B);
    cc::relabel_kernel<<<grid, 256, 0, stream>>>(cc_batch, h, w, s0, lut, lut_stride, dst, stride_s,
                                                 stride_y, stride_x);
  }
  return be_check_launch("relabel kernels");
}

// Pass 1 of run extraction: per-chunk counts (chunk = 1024 elements). The caller scans the
// counts (be_scan_i32_to_i64) and then calls be_runs_write.
// chunk_counts: [2 * (chunks + 1)] int32 (heads then tails; entry `chunks` of each half unused)
int be_runs_count(const int* img, long long n, long long seg_len, int* chunk_counts,
                  cudaStream_t stream) {
  const unsigned chunks = static_cast<unsigned>((n + cc::RUN_CHUNK - 1) / cc::RUN_CHUNK);
  cc::runs_count_kernel<<<chunks, cc::RUN_THREADS, 0, stream>>>(img, n, seg_len, static_cast<int>(chunks), chunk_counts);
  return be_check_launch("runs_count_kernel");
}

// exclusive scan int32 counts -> int64 offsets (+ total at offsets[n]) using cub
int be_scan_i32_to_i64(const int* counts, long long* offsets, long long n, void* temp,
                       size_t temp_bytes, size_t* temp_needed, cudaStream_t stream) {
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, counts, offsets, static_cast<int>(n + 1), stream);
  if (temp_needed) *temp_needed = need;
  if (temp == nullptr) return 0;
  if (temp_bytes < need) return be_set_error("scan temp storage too small");
  cudaError_t e = cub::DeviceScan::ExclusiveSum(temp, need, counts, offsets, static_cast<int>(n + 1), stream);
  if (e != cudaSuccess) return be_set_error(cudaGetErrorString(e));
  return 0;
}

// chunk_offsets: exclusive scans of the two halves of chunk_counts (same layout)
int be_runs_write(const int* img, long long n, long long seg_len, const long long* chunk_offsets,
                  int* out_label, long long* out_start, long long* out_end, long long out_cap,
                  cudaStream_t stream) {
  const unsigned chunks = static_cast<unsigned>((n + cc::RUN_CHUNK - 1) / cc::RUN_CHUNK);
  cc::runs_write_kernel<<<chunks, cc::RUN_THREADS, 0, stream>>>(img, n, seg_len, static_cast<int>(chunks),
                                                          chunk_offsets, out_label, out_start, out_end, out_cap);
  return be_check_launch("runs_write_kernel");
}

// stable sort of runs by 64-bit key (label rank << 40 | sequence) carrying (start, len)
int be_sort_runs(const unsigned long long* keys_in, unsigned long long* keys_out,
                 const int* idx_in, int* idx_out, int n, void* temp, size_t temp_bytes,
                 size_t* temp_needed, cudaStream_t stream) {
  size_t need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, idx_in, idx_out, n, 0, 64, stream);
  if (temp_needed) *temp_needed = need;
  if (temp == nullptr) return 0;
  if (temp_bytes < need) return be_set_error("sort temp storage too small");
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, need, keys_in, keys_out, idx_in, idx_out, n, 0, 64, stream);
  if (e != cudaSuccess) return be_set_error(cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
