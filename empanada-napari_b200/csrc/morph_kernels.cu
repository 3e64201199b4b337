// Tracker-level morphology (empanada/inference/filters.py:154-210,
// applied by Engine3d.infer_on_axis, empanada_napari/inference.py:560-570): the tracker's label
// volume is eroded / dilated with the 3-D cross (skimage.morphology.erosion / dilation with their
// default footprint = grey min / max over the 6-neighbourhood, borders reflected) or hole-filled
// slice by slice, then re-encoded by filters.pan_seg_to_rle_seg: labels outside the class range
// dropped, 26-connected components of equal-valued voxels renumbered in raster order.
//
//   morph3d        : one min / max pass over the dense int32 label volume (4 B in, 4 B out per voxel)
//   range_keep     : labels outside [lo, hi) -> 0 (in place)
//   runs3d_merge   : union-find over the ROW RUNS of the volume (be_runs_count / be_runs_write with
//                    seg_len = W): a run joins the equal-valued runs it touches (diagonals
//                    included) in the previous row of its slice and in rows y-1, y, y+1 of the
//                    previous slice; the smaller (raster-earlier) run index is the root, so the
//                    roots in index order ARE the raster order of the components' first voxels
//   fill_holes     : per (slice, label) in ascending label order, inside the label's bounding box
//                    of the ORIGINAL slice: every non-zero pixel and every hole (zero pixels not
//                    4-connected to the box border through zeros) becomes the label
//                    (filters.py:174-210 with scipy.ndimage.binary_fill_holes) - one CTA per slice
#include <cuda_runtime.h>
#include <cstdint>

#include "common.cuh"

namespace morph {

__global__ void __launch_bounds__(256)
morph3d_kernel(const int* __restrict__ src, int* __restrict__ dst, int D, int H, int W, int op) {
  const long long n = 1LL * D * H * W;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = static_cast<int>(i % W);
  const long long r = i / W;
  const int y = static_cast<int>(r % H), z = static_cast<int>(r / H);
  int v = __ldg(src + i);
  auto acc = [&](int u) { v = op ? max(v, u) : min(v, u); };
  if (x > 0) acc(__ldg(src + i - 1));
  if (x < W - 1) acc(__ldg(src + i + 1));
  if (y > 0) acc(__ldg(src + i - W));
  if (y < H - 1) acc(__ldg(src + i + W));
  if (z > 0) acc(__ldg(src + i - 1LL * H * W));
  if (z < D - 1) acc(__ldg(src + i + 1LL * H * W));
  dst[i] = v;
}

__global__ void __launch_bounds__(256)
range_keep_kernel(int* __restrict__ vol, long long n, int lo, int hi) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = vol[i];
  if (v != 0 && (v < lo || v >= hi)) vol[i] = 0;
}

// row_ptr[r] = index of the first run whose row (start / W) is >= r, r in [0, rows]
__global__ void __launch_bounds__(256)
row_ptr_kernel(const long long* __restrict__ run_start, int n_runs, int W, long long rows, int* __restrict__ row_ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_runs) return;
  const long long row_i = (i < n_runs) ? run_start[i] / W : rows;
  const long long row_p = (i > 0) ? run_start[i - 1] / W : -1;
  for (long long r = row_p + 1; r <= row_i; ++r) row_ptr[r] = i;
}

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
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[b], a);
    if (old == b) return;
    b = old;
  }
}

__global__ void __launch_bounds__(256)
runs3d_merge_kernel(const int* __restrict__ row_ptr, const long long* __restrict__ run_start,
                    const long long* __restrict__ run_end, const int* __restrict__ run_val, int n_runs,
                    int H, int W, int* __restrict__ L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_runs) return;
  const long long s = run_start[i];
  const long long row = s / W;
  const int x0 = static_cast<int>(s - row * W), x1 = x0 + static_cast<int>(run_end[i] - s);
  const int y = static_cast<int>(row % H);
  const long long z = row / H;
  const int v = run_val[i];
  auto scan_row = [&](long long r) {
    for (int q = row_ptr[r]; q < row_ptr[r + 1]; ++q) {
      const long long qs = run_start[q];
      const int qx0 = static_cast<int>(qs - r * W), qx1 = qx0 + static_cast<int>(run_end[q] - qs);
      if (qx1 < x0) continue;          // ends left of x0 - 1
      if (qx0 > x1) break;             // starts right of the voxel after this run's last
      if (run_val[q] == v) unite(L, i, q);
    }
  };
  if (y > 0) scan_row(row - 1);
  if (z > 0) {
    const long long base = row - H;
    if (y > 0) scan_row(base - 1);
    scan_row(base);
    if (y < H - 1) scan_row(base + 1);
  }
}

__global__ void __launch_bounds__(256)
runs3d_resolve_kernel(const int* __restrict__ L, int n_runs, int* __restrict__ root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_runs) root[i] = find_root(L, i);
}

// ------------------------------------------------------------------ fill holes, one CTA per slice
// img: the slice (H x W int32, modified in place); outside: scratch H x W bytes of the slice.
// labels / boxes: this slice's labels in ascending order with their ORIGINAL boxes (y0 x0 y1 x1).
__global__ void __launch_bounds__(512)
fill_holes_kernel(int* __restrict__ vol, uint8_t* __restrict__ scratch, int H, int W,
                  const int* __restrict__ slice_off, const int* __restrict__ labels,
                  const int* __restrict__ boxes) {
  const int z = blockIdx.x;
  int* img = vol + static_cast<long long>(z) * H * W;
  uint8_t* out = scratch + static_cast<long long>(z) * H * W;
  __shared__ int changed;
  for (int k = slice_off[z]; k < slice_off[z + 1]; ++k) {
    const int lab = labels[k];
    const int y0 = boxes[4 * k], x0 = boxes[4 * k + 1], y1 = boxes[4 * k + 2], x1 = boxes[4 * k + 3];
    const int bh = y1 - y0, bw = x1 - x0, n = bh * bw;
    // outside <- zero pixels on the box border
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int yy = i / bw, xx = i - yy * bw;
      const bool border = yy == 0 || xx == 0 || yy == bh - 1 || xx == bw - 1;
      out[(y0 + yy) * W + x0 + xx] = (border && img[(y0 + yy) * W + x0 + xx] == 0) ? 1 : 0;
    }
    __syncthreads();
    // grow through 4-connected zeros until nothing changes
    while (true) {
      if (threadIdx.x == 0) changed = 0;
      __syncthreads();
      int any = 0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int yy = i / bw, xx = i - yy * bw;
        const int p = (y0 + yy) * W + x0 + xx;
        if (out[p] || img[p] != 0) continue;
        if ((yy > 0 && out[p - W]) || (yy < bh - 1 && out[p + W]) || (xx > 0 && out[p - 1]) || (xx < bw - 1 && out[p + 1])) {
          out[p] = 1;
          any = 1;
        }
      }
      if (any) changed = 1;
      __syncthreads();
      const int c = changed;
      __syncthreads();
      if (!c) break;
    }
    // the box: every pixel that is not outside background takes the label, the rest is zero
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int yy = i / bw, xx = i - yy * bw;
      const int p = (y0 + yy) * W + x0 + xx;
      if (img[p] != 0 || !out[p]) img[p] = lab;
    }
    __syncthreads();
  }
}

}  // namespace morph

extern "C" {

int be_morph3d(const int* src, int* dst, int D, int H, int W, int op, cudaStream_t stream) {
  const long long n = 1LL * D * H * W;
  if (n <= 0) return 0;
  morph::morph3d_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, dst, D, H, W, op);
  return be_check_launch("morph3d_kernel");
}

int be_range_keep(int* vol, long long n, int lo, int hi, cudaStream_t stream) {
  if (n <= 0) return 0;
  morph::range_keep_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(vol, n, lo, hi);
  return be_check_launch("range_keep_kernel");
}

// run_* : row runs of a (D,H,W) label volume in raster order (be_runs_count / be_runs_write with
// seg_len = W). row_ptr: [D*H + 1] out. L: [n_runs] workspace. root: [n_runs] out, the index of the
// raster-first run of each run's 26-connected equal-value component.
int be_runs3d_cc(const long long* run_start, const long long* run_end, const int* run_val, int n_runs,
                 int D, int H, int W, int* row_ptr, int* L, int* root, cudaStream_t stream) {
  const long long rows = 1LL * D * H;
  morph::row_ptr_kernel<<<(n_runs + 1 + 255) / 256, 256, 0, stream>>>(run_start, n_runs, W, rows, row_ptr);
  if (n_runs > 0) {
    const unsigned blocks = (n_runs + 255) / 256;
    morph::runs3d_merge_kernel<<<blocks, 256, 0, stream>>>(row_ptr, run_start, run_end, run_val, n_runs, H, W, L);
    morph::runs3d_resolve_kernel<<<blocks, 256, 0, stream>>>(L, n_runs, root);
  }
  return be_check_launch("runs3d_cc kernels");
}

int be_fill_holes(int* vol, uint8_t* scratch, int D, int H, int W, const int* slice_off,
                  const int* labels, const int* boxes, cudaStream_t stream) {
  if (D <= 0) return 0;
  morph::fill_holes_kernel<<<D, 512, 0, stream>>>(vol, scratch, H, W, slice_off, labels, boxes);
  return be_check_launch("fill_holes_kernel");
}

}  // extern "C"
