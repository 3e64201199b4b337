// Orthoplane consensus on TRIPLE RUNS (piece 6a of the hot path). The voxel arithmetic of
// empanada/consensus.py:233-287 (pairwise overlaps between the xy/xz/yz instances), :449-460
// (per-cluster voxel vote, array_utils.py:563-639), merge_overlapping (:144-195), the final fill
// (array_utils.py:754-765) and the run-length tables of the result, computed from maximal x-runs
// of identical (xy, xz, yz) label triples:
//
//   triple_runs (2 dense passes, 12 B per voxel each): count run heads per chunk, then write the
//                runs (row, x0, x1, node triple) in raster order + a per-row pointer;
//   pairs / stats / records / value : one thread per run, proportional to the number of runs
//                (about 1 % of the voxels): pair-overlap table, per-candidate sizes and
//                candidate-pair overlaps, per-instance (id, start, length) records of every
//                claimed voxel range (overlapped ranges included, so no side list is needed),
//                and the painted value of every run;
//   paint      : the consensus volume written once (runs::paint_rows_kernel through be_runs_paint);
//   records    : sorted by (instance, start) and joined where they touch = the instance's RLE
//                exactly as the reference's sorted, merged ranges (array_utils.py:659-752).
//
// The graph decisions between these passes run on the host on the small tables they emit.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <cstdint>

#include "common.cuh"

namespace crun {

constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return k;
}
__device__ __forceinline__ bool hash_add(unsigned long long* keys, int* vals, unsigned long long mask,
                                         unsigned long long key, int count) {
  unsigned long long slot = mix64(key) & mask;
  for (unsigned long long probe = 0; probe <= mask; ++probe) {
    const unsigned long long cur = keys[slot];
    if (cur == key) { atomicAdd(&vals[slot], count); return true; }
    if (cur == EMPTY_KEY) {
      const unsigned long long old = atomicCAS(&keys[slot], EMPTY_KEY, key);
      if (old == EMPTY_KEY || old == key) { atomicAdd(&vals[slot], count); return true; }
    }
    slot = (slot + 1) & mask;
  }
  return false;
}

struct Vols {   // three label volumes (NULL = plane absent) and their label -> node LUTs
  const int* v[3]; const int* lut[3]; int nlut[3];
};
struct Triple { int a, b, c; };
__device__ __forceinline__ bool same(const Triple& x, const Triple& y) { return x.a == y.a && x.b == y.b && x.c == y.c; }
__device__ __forceinline__ bool nonzero(const Triple& t) { return (t.a | t.b | t.c) != 0; }
__device__ __forceinline__ int node_of(const Vols& V, int p, int label) {
  return (label > 0 && label < V.nlut[p]) ? __ldg(V.lut[p] + label) : 0;
}
__device__ __forceinline__ Triple triple_at(const Vols& V, long long i) {
  Triple t;
  t.a = V.v[0] ? node_of(V, 0, V.v[0][i]) : 0;
  t.b = V.v[1] ? node_of(V, 1, V.v[1][i]) : 0;
  t.c = V.v[2] ? node_of(V, 2, V.v[2][i]) : 0;
  return t;
}

// four consecutive voxels i0 .. i0+3 (flat index; rows of W voxels): node triples, run heads / tails
struct Quad { Triple t[4]; int head[4]; int tail[4]; };
__device__ __forceinline__ bool load_quad(const Vols& V, long long i0, long long n, int W, bool vec, Quad& q) {
#pragma unroll
  for (int j = 0; j < 4; ++j) { q.t[j] = Triple{0, 0, 0}; q.head[j] = 0; q.tail[j] = 0; }
  if (i0 >= n) return false;
  int lab[3][4];
  int any = 0;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
#pragma unroll
    for (int j = 0; j < 4; ++j) lab[p][j] = 0;
    if (V.v[p] == nullptr) continue;
    if (vec && i0 + 4 <= n) {
      const int4 t = __ldg(reinterpret_cast<const int4*>(V.v[p] + i0));
      lab[p][0] = t.x; lab[p][1] = t.y; lab[p][2] = t.z; lab[p][3] = t.w;
    } else {
      for (int j = 0; j < 4 && i0 + j < n; ++j) lab[p][j] = V.v[p][i0 + j];
    }
    any |= lab[p][0] | lab[p][1] | lab[p][2] | lab[p][3];
  }
  if (any == 0) return false;
  bool nz = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q.t[j].a = node_of(V, 0, lab[0][j]);
    q.t[j].b = node_of(V, 1, lab[1][j]);
    q.t[j].c = node_of(V, 2, lab[2][j]);
    nz |= nonzero(q.t[j]);
  }
  if (!nz) return false;
  const int x0 = static_cast<int>(i0 % W);
  Triple before{0, 0, 0}, after{0, 0, 0};
  if (nonzero(q.t[0]) && x0 > 0) before = triple_at(V, i0 - 1);
  if (nonzero(q.t[3]) && i0 + 4 < n && (x0 + 3) % W != W - 1) after = triple_at(V, i0 + 4);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (i0 + j >= n || !nonzero(q.t[j])) continue;
    const int x = (x0 + j) % W;                       // a quad may straddle a row end when W % 4 != 0
    const Triple pv = (j == 0) ? before : q.t[j - 1];
    const Triple nv = (j == 3) ? after : q.t[j + 1];
    q.head[j] = (x == 0) || !same(pv, q.t[j]);
    q.tail[j] = (x == W - 1) || (i0 + j + 1 >= n) || !same(nv, q.t[j]);
  }
  return true;
}

constexpr int TR_THREADS = 256;
// PASS 0: heads / tails per chunk of 1024 voxels -> counts[2][chunks].
// PASS 1: runs written at (scanned chunk offset + rank): run_yx (row, x0), run_abc (a, b, c, 0) by
//         heads, run_x1 by tails; row_ptr[row] = index of the first run at or after the row start.
template <int PASS>
__global__ void __launch_bounds__(TR_THREADS)
triple_runs_kernel(Vols V, long long n, int W, int vec, long long chunks, int* __restrict__ counts,
                   const long long* __restrict__ offsets, int* __restrict__ row_ptr,
                   int2* __restrict__ run_yx, int* __restrict__ run_x1, int4* __restrict__ run_abc) {
  const long long ch = blockIdx.x;
  const long long i0 = (ch * TR_THREADS + threadIdx.x) * 4;
  Quad q;
  const bool any = load_quad(V, i0, n, W, vec != 0, q);
  int nh = 0, nt = 0;
  if (any) {
    nh = q.head[0] + q.head[1] + q.head[2] + q.head[3];
    nt = q.tail[0] + q.tail[1] + q.tail[2] + q.tail[3];
  }
  const bool block_any = __syncthreads_or(any);
  if (PASS == 0) {
    if (!block_any) {
      if (threadIdx.x == 0) { counts[ch] = 0; counts[chunks + 1 + ch] = 0; }
      return;
    }
    typedef cub::BlockReduce<int, TR_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const int th = Reduce(tmp).Sum(nh);
    __syncthreads();
    const int tt = Reduce(tmp).Sum(nt);
    if (threadIdx.x == 0) { counts[ch] = th; counts[chunks + 1 + ch] = tt; }
    return;
  }
  typedef cub::BlockScan<int, TR_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  int exh = 0, ext = 0;
  if (block_any) {
    Scan(tmp).ExclusiveSum(nh, exh);
    __syncthreads();
    Scan(tmp).ExclusiveSum(nt, ext);
  }
  const long long offh = offsets[ch], offt = offsets[chunks + 1 + ch];
  // row pointers: every voxel that starts a row records how many runs precede it
  if (i0 < n) {
    const int x0 = static_cast<int>(i0 % W);
    int before = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j >= n) break;
      if ((x0 + j) % W == 0) row_ptr[(i0 + j) / W] = static_cast<int>(offh + exh + before);
      before += q.head[j];
    }
  }
  if (nh) {
    long long pos = offh + exh;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q.head[j]) {
        const long long i = i0 + j;
        run_yx[pos] = make_int2(static_cast<int>(i / W), static_cast<int>(i % W));
        run_abc[pos] = make_int4(q.t[j].a, q.t[j].b, q.t[j].c, 0);
        ++pos;
      }
  }
  if (nt) {
    long long pos = offt + ext;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q.tail[j]) { run_x1[pos] = static_cast<int>((i0 + j) % W) + 1; ++pos; }
  }
}

// ---------------------------------------------------------------- per-run kernels
// pair overlaps between instances of different planes; key = nodeA(32) | nodeB(32), A < B
__global__ void pairs_kernel(const int2* __restrict__ run_yx, const int* __restrict__ run_x1,
                             const int4* __restrict__ run_abc, long long n_runs,
                             unsigned long long* __restrict__ keys, int* __restrict__ vals,
                             unsigned long long mask, int* __restrict__ overflow) {
  const long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (r >= n_runs) return;
  const int4 t = run_abc[r];
  if (((t.x != 0) + (t.y != 0) + (t.z != 0)) < 2) return;
  const int len = run_x1[r] - run_yx[r].y;
  bool ok = true;
  if (t.x && t.y) ok &= hash_add(keys, vals, mask, (static_cast<unsigned long long>(t.x) << 32) | t.y, len);
  if (t.x && t.z) ok &= hash_add(keys, vals, mask, (static_cast<unsigned long long>(t.x) << 32) | t.z, len);
  if (t.y && t.z) ok &= hash_add(keys, vals, mask, (static_cast<unsigned long long>(t.y) << 32) | t.z, len);
  if (!ok) atomicExch(overflow, 1);
}

// memb_off[node] .. memb_off[node+1] indexes memb_list: the candidate output instances ("cids",
// 1-based) the node is a member of (usually exactly one). A voxel is claimed by cid when at
// least vote_thr of its (<= 3) nodes are members.
__device__ __forceinline__ bool in_list(const int* __restrict__ memb_off, const int* __restrict__ memb_list,
                                        int nd, int cid) {
  if (nd == 0) return false;
  for (int m = memb_off[nd]; m < memb_off[nd + 1]; ++m)
    if (memb_list[m] == cid) return true;
  return false;
}
constexpr int MAX_CLAIMS = 32;
__device__ __forceinline__ int collect_claims(const int4& t, const int* __restrict__ memb_off,
                                              const int* __restrict__ memb_list, int vote_thr, int* claims) {
  const int nodes[3] = {t.x, t.y, t.z};
  int n = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int nd = nodes[j];
    if (nd == 0) continue;
    for (int m = memb_off[nd]; m < memb_off[nd + 1]; ++m) {
      const int cid = memb_list[m];
      bool seen = false;
      for (int i = 0; i < j; ++i) seen |= in_list(memb_off, memb_list, nodes[i], cid);
      if (seen) continue;  // counted at its first node
      int votes = 1;
      for (int i = j + 1; i < 3; ++i) votes += in_list(memb_off, memb_list, nodes[i], cid) ? 1 : 0;
      if (votes >= vote_thr) {
        if (n == MAX_CLAIMS) return -1;
        claims[n++] = cid;
      }
    }
  }
  return n;
}

// sizes[cid] += run length for every claiming cid; overlaps between cids claiming the same run
__global__ void stats_kernel(const int2* __restrict__ run_yx, const int* __restrict__ run_x1,
                             const int4* __restrict__ run_abc, long long n_runs,
                             const int* __restrict__ memb_off, const int* __restrict__ memb_list,
                             int vote_thr, int* __restrict__ sizes,
                             unsigned long long* __restrict__ keys, int* __restrict__ vals,
                             unsigned long long mask, int* __restrict__ overflow) {
  const long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (r >= n_runs) return;
  int claims[MAX_CLAIMS];
  const int k = collect_claims(run_abc[r], memb_off, memb_list, vote_thr, claims);
  if (k < 0) { atomicExch(overflow, 2); return; }
  if (k == 0) return;
  const int len = run_x1[r] - run_yx[r].y;
  bool ok = true;
  for (int u = 0; u < k; ++u) {
    atomicAdd(&sizes[claims[u]], len);
    for (int v = u + 1; v < k; ++v) {
      const int lo = min(claims[u], claims[v]), hi = max(claims[u], claims[v]);
      ok &= hash_add(keys, vals, mask, (static_cast<unsigned long long>(lo) << 32) | hi, len);
    }
  }
  if (!ok) atomicExch(overflow, 1);
}

// distinct final instance ids claiming a run (cid_final[cid], 0 = no instance)
__device__ __forceinline__ int final_ids(const int4& t, const int* __restrict__ memb_off,
                                         const int* __restrict__ memb_list, int vote_thr,
                                         const int* __restrict__ cid_final, int* fin) {
  int claims[MAX_CLAIMS];
  const int k = collect_claims(t, memb_off, memb_list, vote_thr, claims);  // >= 0: checked by stats_kernel
  int nf = 0;
  for (int u = 0; u < k; ++u) {
    const int f = cid_final[claims[u]];
    if (f == 0) continue;
    bool dup = false;
    for (int v = 0; v < nf; ++v) dup |= (fin[v] == f);
    if (!dup) fin[nf++] = f;
  }
  return nf;
}

// MODE 0: rec_count[r] = number of final ids claiming run r; fsize[f] += length.
// MODE 1: records (key = f << 40 | flat start, len) at rec_off[r]..; run_val[r] = max KEPT final id
//         (later ids overwrite earlier ones in the reference's in-order fill).
template <int MODE>
__global__ void records_kernel(const int2* __restrict__ run_yx, const int* __restrict__ run_x1,
                               const int4* __restrict__ run_abc, long long n_runs, int W,
                               long long flat0, const int* __restrict__ memb_off,
                               const int* __restrict__ memb_list, int vote_thr,
                               const int* __restrict__ cid_final, int* __restrict__ rec_count,
                               int* __restrict__ fsize, const long long* __restrict__ rec_off,
                               const int* __restrict__ keep, unsigned long long* __restrict__ rec_key,
                               int* __restrict__ rec_len, int* __restrict__ run_val) {
  const long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (r >= n_runs) return;
  int fin[MAX_CLAIMS];
  const int nf = final_ids(run_abc[r], memb_off, memb_list, vote_thr, cid_final, fin);
  const int2 yx = run_yx[r];
  const int len = run_x1[r] - yx.y;
  if (MODE == 0) {
    rec_count[r] = nf;
    for (int v = 0; v < nf; ++v) atomicAdd(&fsize[fin[v]], len);
    return;
  }
  const unsigned long long start = static_cast<unsigned long long>(flat0 + static_cast<long long>(yx.x) * W + yx.y);
  long long pos = rec_off[r];
  int best = 0;
  for (int v = 0; v < nf; ++v) {
    rec_key[pos] = (static_cast<unsigned long long>(fin[v]) << 40) | start;
    rec_len[pos] = len;
    ++pos;
    if (keep[fin[v]]) best = max(best, fin[v]);
  }
  run_val[r] = best;
}

// sorted records -> head flags of the joined ranges: a record starts a new range unless it
// continues the previous record of the same instance (touching flat ranges are one run,
// array_utils.py:659-752)
__global__ void join_flags_kernel(const unsigned long long* __restrict__ key, const int* __restrict__ len,
                                  long long n, int* __restrict__ head) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  int h = 1;
  if (i > 0) {
    const unsigned long long k0 = key[i - 1], k1 = key[i];
    if ((k0 >> 40) == (k1 >> 40) && (k0 & ((1ull << 40) - 1)) + static_cast<unsigned long long>(len[i - 1]) == (k1 & ((1ull << 40) - 1)))
      h = 0;
  }
  head[i] = h;
}
// joined ranges: out_start / out_len / out_id at (inclusive head rank - 1); lengths accumulate
__global__ void join_write_kernel(const unsigned long long* __restrict__ key, const int* __restrict__ len,
                                  const int* __restrict__ head, const int* __restrict__ rank_incl,
                                  long long n, long long* __restrict__ out_start,
                                  unsigned long long* __restrict__ out_len_acc, int* __restrict__ out_id) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const long long o = static_cast<long long>(rank_incl[i]) - 1;
  if (head[i]) {
    out_start[o] = static_cast<long long>(key[i] & ((1ull << 40) - 1));
    out_id[o] = static_cast<int>(key[i] >> 40);
  }
  atomicAdd(&out_len_acc[o], static_cast<unsigned long long>(len[i]));
}

}  // namespace crun

// ------------------------------------------------------------------------------ launchers
extern "C" {

static crun::Vols make_vols(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                            const int* lc, int na, int nb, int nc) {
  crun::Vols V;
  V.v[0] = va; V.v[1] = vb; V.v[2] = vc;
  V.lut[0] = la; V.lut[1] = lb; V.lut[2] = lc;
  V.nlut[0] = na; V.nlut[1] = nb; V.nlut[2] = nc;
  return V;
}
static int vols_vec(const int* va, const int* vb, const int* vc) {
  return ((reinterpret_cast<uintptr_t>(va) | reinterpret_cast<uintptr_t>(vb) | reinterpret_cast<uintptr_t>(vc)) & 15) == 0;
}

// counts: [2 * (chunks + 1)] int32, chunks = ceil(n / 1024)
int be_triple_count(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                    const int* lc, int na, int nb, int nc, long long n, int W, int* counts,
                    cudaStream_t stream) {
  const long long chunks = (n + 1023) / 1024;
  crun::triple_runs_kernel<0><<<static_cast<unsigned>(chunks), crun::TR_THREADS, 0, stream>>>(
      make_vols(va, vb, vc, la, lb, lc, na, nb, nc), n, W, vols_vec(va, vb, vc), chunks, counts, nullptr, nullptr,
      nullptr, nullptr, nullptr);
  return be_check_launch("triple_runs_kernel<count>");
}

// offsets: exclusive scans of the two halves of counts (be_scan_i32_to_i64 per half);
// row_ptr: [rows + 1] int32 (entry `rows` is set by the caller to the total)
int be_triple_write(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                    const int* lc, int na, int nb, int nc, long long n, int W,
                    const long long* offsets, int* row_ptr, int* run_yx, int* run_x1, int* run_abc,
                    cudaStream_t stream) {
  const long long chunks = (n + 1023) / 1024;
  crun::triple_runs_kernel<1><<<static_cast<unsigned>(chunks), crun::TR_THREADS, 0, stream>>>(
      make_vols(va, vb, vc, la, lb, lc, na, nb, nc), n, W, vols_vec(va, vb, vc), chunks, nullptr, offsets, row_ptr,
      reinterpret_cast<int2*>(run_yx), run_x1, reinterpret_cast<int4*>(run_abc));
  return be_check_launch("triple_runs_kernel<write>");
}

int be_triple_pairs(const int* run_yx, const int* run_x1, const int* run_abc, long long n_runs,
                    unsigned long long* keys, int* vals, unsigned long long cap, int* overflow,
                    cudaStream_t stream) {
  if (cap & (cap - 1)) return be_set_error("hash capacity must be a power of two");
  if (n_runs <= 0) return 0;
  crun::pairs_kernel<<<static_cast<unsigned>((n_runs + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const int2*>(run_yx), run_x1, reinterpret_cast<const int4*>(run_abc), n_runs, keys, vals,
      cap - 1, overflow);
  return be_check_launch("pairs_kernel");
}

int be_triple_stats(const int* run_yx, const int* run_x1, const int* run_abc, long long n_runs,
                    const int* memb_off, const int* memb_list, int vote_thr, int* sizes,
                    unsigned long long* keys, int* vals, unsigned long long cap, int* overflow,
                    cudaStream_t stream) {
  if (cap & (cap - 1)) return be_set_error("hash capacity must be a power of two");
  if (n_runs <= 0) return 0;
  crun::stats_kernel<<<static_cast<unsigned>((n_runs + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const int2*>(run_yx), run_x1, reinterpret_cast<const int4*>(run_abc), n_runs, memb_off,
      memb_list, vote_thr, sizes, keys, vals, cap - 1, overflow);
  return be_check_launch("stats_kernel");
}

int be_triple_rec_count(const int* run_yx, const int* run_x1, const int* run_abc, long long n_runs,
                        const int* memb_off, const int* memb_list, int vote_thr, const int* cid_final,
                        int* rec_count, int* fsize, cudaStream_t stream) {
  if (n_runs <= 0) return 0;
  crun::records_kernel<0><<<static_cast<unsigned>((n_runs + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const int2*>(run_yx), run_x1, reinterpret_cast<const int4*>(run_abc), n_runs, 0, 0, memb_off,
      memb_list, vote_thr, cid_final, rec_count, fsize, nullptr, nullptr, nullptr, nullptr, nullptr);
  return be_check_launch("records_kernel<count>");
}

int be_triple_rec_write(const int* run_yx, const int* run_x1, const int* run_abc, long long n_runs,
                        int W, long long flat0, const int* memb_off, const int* memb_list,
                        int vote_thr, const int* cid_final, const long long* rec_off, const int* keep,
                        unsigned long long* rec_key, int* rec_len, int* run_val, cudaStream_t stream) {
  if (n_runs <= 0) return 0;
  crun::records_kernel<1><<<static_cast<unsigned>((n_runs + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const int2*>(run_yx), run_x1, reinterpret_cast<const int4*>(run_abc), n_runs, W, flat0,
      memb_off, memb_list, vote_thr, cid_final, nullptr, nullptr, rec_off, keep, rec_key, rec_len, run_val);
  return be_check_launch("records_kernel<write>");
}

// sort records by key (instance << 40 | start)
int be_sort_records(const unsigned long long* keys_in, unsigned long long* keys_out, const int* len_in,
                    int* len_out, long long n, void* temp, size_t temp_bytes, size_t* temp_needed,
                    cudaStream_t stream) {
  size_t need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, len_in, len_out, static_cast<int>(n), 0, 64, stream);
  if (temp_needed) *temp_needed = need;
  if (temp == nullptr) return 0;
  if (temp_bytes < need) return be_set_error("sort temp storage too small");
  const cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, need, keys_in, keys_out, len_in, len_out,
                                                        static_cast<int>(n), 0, 64, stream);
  return e == cudaSuccess ? 0 : be_set_error(cudaGetErrorString(e));
}

// head flags + inclusive ranks of the joined ranges (rank_incl[n-1] = number of joined ranges)
int be_join_flags(const unsigned long long* key, const int* len, long long n, int* head,
                  int* rank_incl, void* temp, size_t temp_bytes, size_t* temp_needed,
                  cudaStream_t stream) {
  size_t need = 0;
  cub::DeviceScan::InclusiveSum(nullptr, need, head, rank_incl, static_cast<int>(n), stream);
  if (temp_needed) *temp_needed = need;
  if (temp == nullptr) return 0;
  if (temp_bytes < need) return be_set_error("scan temp storage too small");
  if (n <= 0) return 0;
  crun::join_flags_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(key, len, n, head);
  const cudaError_t e = cub::DeviceScan::InclusiveSum(temp, need, head, rank_incl, static_cast<int>(n), stream);
  return e == cudaSuccess ? be_check_launch("join_flags_kernel") : be_set_error(cudaGetErrorString(e));
}

// out_len_acc must be zero on entry (uint64 per joined range)
int be_join_write(const unsigned long long* key, const int* len, const int* head,
                  const int* rank_incl, long long n, long long* out_start,
                  unsigned long long* out_len_acc, int* out_id, cudaStream_t stream) {
  if (n <= 0) return 0;
  crun::join_write_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(key, len, head, rank_incl, n,
                                                                                   out_start, out_len_acc, out_id);
  return be_check_launch("join_write_kernel");
}

}  // extern "C"
