// Orthoplane consensus kernels (piece 6a of the hot path): the voxel-level arithmetic of
// empanada/consensus.py:233-287 (pairwise instance overlaps between the xy/xz/yz trackers),
// :449-460 (per-cluster voxel vote, array_utils.py:563-639) and the final paint
// (array_utils.py:754-765), computed on three dense (D,H,W) int32 label volumes instead of
// sorted range lists. The graph decisions between these passes run on the host on the tiny
// tables these kernels emit. HBM-bound: 12 B read per voxel per pass (+4 B write when painting);
// atomics only at the head of each run of identical (a,b,c) label triples.
#include <cuda_runtime.h>
#include <cstdint>

#include "common.cuh"

namespace cons {

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

struct Triple { int a, b, c; };
// label -> node index (1-based, 0 = not an instance) through per-plane LUTs
__device__ __forceinline__ Triple load_triple(const int* va, const int* vb, const int* vc,
                                              const int* la, const int* lb, const int* lc,
                                              int na, int nb, int nc, long long i) {
  Triple t;
  int v = va ? va[i] : 0; t.a = (v > 0 && v < na) ? la[v] : 0;
  v = vb ? vb[i] : 0;     t.b = (v > 0 && v < nb) ? lb[v] : 0;
  v = vc ? vc[i] : 0;     t.c = (v > 0 && v < nc) ? lc[v] : 0;
  return t;
}
__device__ __forceinline__ bool same(const Triple& x, const Triple& y) {
  return x.a == y.a && x.b == y.b && x.c == y.c;
}

// ------------------------------------------------------------------ pass 1: pair overlaps
// key = nodeA(32) | nodeB(32) with nodeA < nodeB (global node numbering: xy, then xz, then yz)
__device__ __forceinline__ void plane_pairs_voxel(const int* __restrict__ va, const int* __restrict__ vb,
                                                  const int* __restrict__ vc, const int* __restrict__ la,
                                                  const int* __restrict__ lb, const int* __restrict__ lc,
                                                  int na, int nb, int nc, long long i, int W,
                                                  unsigned long long* __restrict__ keys, int* __restrict__ vals,
                                                  unsigned long long mask, int* __restrict__ overflow) {
  const Triple t = load_triple(va, vb, vc, la, lb, lc, na, nb, nc, i);
  const int nz = (t.a != 0) + (t.b != 0) + (t.c != 0);
  if (nz < 2) return;
  const int x = static_cast<int>(i % W);
  if (x > 0 && same(t, load_triple(va, vb, vc, la, lb, lc, na, nb, nc, i - 1))) return;
  int len = 1;
  while (x + len < W && same(t, load_triple(va, vb, vc, la, lb, lc, na, nb, nc, i + len))) ++len;
  bool ok = true;
  if (t.a && t.b) ok &= hash_add(keys, vals, mask, (static_cast<unsigned long long>(t.a) << 32) | t.b, len);
  if (t.a && t.c) ok &= hash_add(keys, vals, mask, (static_cast<unsigned long long>(t.a) << 32) | t.c, len);
  if (t.b && t.c) ok &= hash_add(keys, vals, mask, (static_cast<unsigned long long>(t.b) << 32) | t.c, len);
  if (!ok) atomicExch(overflow, 1);
}

// One thread = four consecutive voxels fetched as one 16-byte vector per plane. Label volumes are
// mostly background: a quad whose twelve labels are all zero (the common case) costs three
// vector loads and nothing else; only quads that touch an instance take the per-voxel path.
__device__ __forceinline__ int4 ld4(const int* __restrict__ v, long long i) {
  return v ? __ldg(reinterpret_cast<const int4*>(v + i)) : make_int4(0, 0, 0, 0);
}
__device__ __forceinline__ int any4(const int4& a) { return a.x | a.y | a.z | a.w; }

__global__ void plane_pairs_kernel(const int* __restrict__ va, const int* __restrict__ vb,
                                   const int* __restrict__ vc, const int* __restrict__ la,
                                   const int* __restrict__ lb, const int* __restrict__ lc,
                                   int na, int nb, int nc, long long n, int W,
                                   unsigned long long* __restrict__ keys, int* __restrict__ vals,
                                   unsigned long long mask, int* __restrict__ overflow) {
  const long long i0 = 4 * (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x);
  if (i0 >= n) return;
  if (i0 + 4 <= n) {
    const int4 A = ld4(va, i0), B = ld4(vb, i0), C = ld4(vc, i0);
    // at least two planes must be labelled somewhere in the quad
    if (((any4(A) != 0) + (any4(B) != 0) + (any4(C) != 0)) < 2) return;
  }
  for (long long i = i0; i < min(i0 + 4, n); ++i)
    plane_pairs_voxel(va, vb, vc, la, lb, lc, na, nb, nc, i, W, keys, vals, mask, overflow);
}

// ------------------------------------------------------------------ pass 2 / 3: votes
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
// returns the number of claiming cids written to claims[] (or -1 if more than MAX_CLAIMS claim)
__device__ __forceinline__ int collect_claims(const Triple& t, const int* __restrict__ memb_off,
                                              const int* __restrict__ memb_list, int vote_thr,
                                              int* claims) {
  const int nodes[3] = {t.a, t.b, t.c};
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

// MODE 0: statistics (size per cid, pairwise overlap between cids claiming the same voxel)
// MODE 1: paint. cid_final[cid] = final instance id (0 = dropped); voxel <- max final id;
//         voxels claimed by >1 distinct final ids are appended to the side list (voxel, id).
template <int MODE>
__device__ __forceinline__ int vote_voxel(const int* __restrict__ va, const int* __restrict__ vb,
                            const int* __restrict__ vc, const int* __restrict__ la,
                            const int* __restrict__ lb, const int* __restrict__ lc, int na, int nb,
                            int nc, long long i, int W, const int* __restrict__ memb_off,
                            const int* __restrict__ memb_list, int vote_thr,
                            int* __restrict__ sizes, unsigned long long* __restrict__ keys,
                            int* __restrict__ vals, unsigned long long mask,
                            int* __restrict__ overflow, const int* __restrict__ cid_final,
                            long long* __restrict__ side_voxel,
                            int* __restrict__ side_id, int side_cap, int* __restrict__ side_count) {
  const Triple t = load_triple(va, vb, vc, la, lb, lc, na, nb, nc, i);
  if (MODE == 0) {
    if ((t.a | t.b | t.c) == 0) return 0;
    const int x = static_cast<int>(i % W);
    if (x > 0 && same(t, load_triple(va, vb, vc, la, lb, lc, na, nb, nc, i - 1))) return 0;
    int len = 1;
    while (x + len < W && same(t, load_triple(va, vb, vc, la, lb, lc, na, nb, nc, i + len))) ++len;
    int claims[MAX_CLAIMS];
    const int k = collect_claims(t, memb_off, memb_list, vote_thr, claims);
    if (k < 0) { atomicExch(overflow, 2); return 0; }
    bool ok = true;
    for (int u = 0; u < k; ++u) {
      atomicAdd(&sizes[claims[u]], len);
      for (int v = u + 1; v < k; ++v) {
        const int lo = min(claims[u], claims[v]), hi = max(claims[u], claims[v]);
        ok &= hash_add(keys, vals, mask, (static_cast<unsigned long long>(lo) << 32) | hi, len);
      }
    }
    if (!ok) atomicExch(overflow, 1);
    return 0;
  } else {
    int best = 0;
    if ((t.a | t.b | t.c) != 0) {
      int claims[MAX_CLAIMS];
      const int k = collect_claims(t, memb_off, memb_list, vote_thr, claims);  // >= 0: checked by the stats pass
      int fin[MAX_CLAIMS], nf = 0;
      for (int u = 0; u < k; ++u) {
        const int f = cid_final[claims[u]];
        if (f == 0) continue;
        bool dup = false;
        for (int v = 0; v < nf; ++v) dup |= (fin[v] == f);
        if (!dup) fin[nf++] = f;
        best = max(best, f);
      }
      if (nf > 1) {
        for (int v = 0; v < nf; ++v) {
          const int pos = atomicAdd(side_count, 1);
          if (pos < side_cap) { side_voxel[pos] = i; side_id[pos] = fin[v]; }
        }
      }
    }
    return best;
  }
}

// Quad-vectorised like plane_pairs_kernel: all-background quads (the common case) cost three
// 16-byte loads and, when painting, one 16-byte store of zeros.
template <int MODE>
__global__ void vote_kernel(const int* __restrict__ va, const int* __restrict__ vb,
                            const int* __restrict__ vc, const int* __restrict__ la,
                            const int* __restrict__ lb, const int* __restrict__ lc, int na, int nb,
                            int nc, long long n, int W, const int* __restrict__ memb_off,
                            const int* __restrict__ memb_list, int vote_thr,
                            int* __restrict__ sizes, unsigned long long* __restrict__ keys,
                            int* __restrict__ vals, unsigned long long mask,
                            int* __restrict__ overflow, const int* __restrict__ cid_final,
                            int* __restrict__ out, long long* __restrict__ side_voxel,
                            int* __restrict__ side_id, int side_cap, int* __restrict__ side_count) {
  const long long i0 = 4 * (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x);
  if (i0 >= n) return;
  if (i0 + 4 <= n) {
    const int4 A = ld4(va, i0), B = ld4(vb, i0), C = ld4(vc, i0);
    if ((any4(A) | any4(B) | any4(C)) == 0) {
      if (MODE == 1) *reinterpret_cast<int4*>(out + i0) = make_int4(0, 0, 0, 0);
      return;
    }
    int r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      r[j] = vote_voxel<MODE>(va, vb, vc, la, lb, lc, na, nb, nc, i0 + j, W, memb_off, memb_list, vote_thr, sizes,
                              keys, vals, mask, overflow, cid_final, side_voxel, side_id, side_cap, side_count);
    if (MODE == 1) *reinterpret_cast<int4*>(out + i0) = make_int4(r[0], r[1], r[2], r[3]);
    return;
  }
  for (long long i = i0; i < n; ++i) {
    const int r = vote_voxel<MODE>(va, vb, vc, la, lb, lc, na, nb, nc, i, W, memb_off, memb_list, vote_thr, sizes,
                                   keys, vals, mask, overflow, cid_final, side_voxel, side_id, side_cap, side_count);
    if (MODE == 1) out[i] = r;
  }
}

// histogram of a label volume (instance sizes after painting)
__global__ void label_hist_kernel(const int* __restrict__ vol, long long n, int W, int nbins,
                                  int* __restrict__ hist) {
  const long long i0 = 4 * (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x);
  if (i0 >= n) return;
  if (i0 + 4 <= n && any4(ld4(vol, i0)) == 0) return;
  for (long long i = i0; i < min(i0 + 4, n); ++i) {
    const int v = vol[i];
    if (v <= 0 || v >= nbins) continue;
    const int x = static_cast<int>(i % W);
    if (x > 0 && vol[i - 1] == v) continue;
    int len = 1;
    while (x + len < W && vol[i + len] == v) ++len;
    atomicAdd(&hist[v], len);
  }
}

// vol[i] = lut[vol[i]] in place (drop filtered instances); untouched quads are not rewritten
__global__ void lut_inplace_kernel(int* __restrict__ vol, long long n, const int* __restrict__ lut,
                                   int nlut) {
  const long long i0 = 4 * (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x);
  if (i0 >= n) return;
  if (i0 + 4 <= n) {
    int4 q = *reinterpret_cast<const int4*>(vol + i0);
    if (any4(q) == 0) return;
    int4 o = q;
    if (q.x > 0) o.x = (q.x < nlut) ? lut[q.x] : 0;
    if (q.y > 0) o.y = (q.y < nlut) ? lut[q.y] : 0;
    if (q.z > 0) o.z = (q.z < nlut) ? lut[q.z] : 0;
    if (q.w > 0) o.w = (q.w < nlut) ? lut[q.w] : 0;
    if (o.x != q.x || o.y != q.y || o.z != q.z || o.w != q.w) *reinterpret_cast<int4*>(vol + i0) = o;
    return;
  }
  for (long long i = i0; i < n; ++i) {
    const int v = vol[i];
    if (v > 0) vol[i] = (v < nlut) ? lut[v] : 0;
  }
}

}  // namespace cons

extern "C" {

int be_plane_pairs(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                   const int* lc, int na, int nb, int nc, long long n, int W,
                   unsigned long long* keys, int* vals, unsigned long long cap, int* overflow,
                   cudaStream_t stream) {
  if (cap & (cap - 1)) return be_set_error("hash capacity must be a power of two");
  cons::plane_pairs_kernel<<<static_cast<unsigned>(((n + 3) / 4 + 255) / 256), 256, 0, stream>>>(
      va, vb, vc, la, lb, lc, na, nb, nc, n, W, keys, vals, cap - 1, overflow);
  return be_check_launch("plane_pairs_kernel");
}

int be_vote_stats(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                  const int* lc, int na, int nb, int nc, long long n, int W, const int* memb_off,
                  const int* memb_list, int vote_thr, int* sizes, unsigned long long* keys,
                  int* vals, unsigned long long cap, int* overflow, cudaStream_t stream) {
  if (cap & (cap - 1)) return be_set_error("hash capacity must be a power of two");
  cons::vote_kernel<0><<<static_cast<unsigned>(((n + 3) / 4 + 255) / 256), 256, 0, stream>>>(
      va, vb, vc, la, lb, lc, na, nb, nc, n, W, memb_off, memb_list, vote_thr, sizes, keys, vals,
      cap - 1, overflow, nullptr, nullptr, nullptr, nullptr, 0, nullptr);
  return be_check_launch("vote_kernel<stats>");
}

int be_vote_paint(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                  const int* lc, int na, int nb, int nc, long long n, int W, const int* memb_off,
                  const int* memb_list, int vote_thr, const int* cid_final, int* out,
                  long long* side_voxel, int* side_id, int side_cap, int* side_count,
                  cudaStream_t stream) {
  cons::vote_kernel<1><<<static_cast<unsigned>(((n + 3) / 4 + 255) / 256), 256, 0, stream>>>(
      va, vb, vc, la, lb, lc, na, nb, nc, n, W, memb_off, memb_list, vote_thr, nullptr, nullptr,
      nullptr, 0, nullptr, cid_final, out, side_voxel, side_id, side_cap, side_count);
  return be_check_launch("vote_kernel<paint>");
}

int be_label_hist(const int* vol, long long n, int W, int nbins, int* hist, cudaStream_t stream) {
  cons::label_hist_kernel<<<static_cast<unsigned>(((n + 3) / 4 + 255) / 256), 256, 0, stream>>>(vol, n, W, nbins, hist);
  return be_check_launch("label_hist_kernel");
}

int be_lut_inplace(int* vol, long long n, const int* lut, int nlut, cudaStream_t stream) {
  cons::lut_inplace_kernel<<<static_cast<unsigned>(((n + 3) / 4 + 255) / 256), 256, 0, stream>>>(vol, n, lut, nlut);
  return be_check_launch("lut_inplace_kernel");
}

}  // extern "C"
