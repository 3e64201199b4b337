// Host replay of the reference's sequential instance tracker on the sparse tables the GPU
// kernels produce (per-slice component areas/boxes + adjacent-slice overlap counts).
// Decision logic restated from:
//   RLEMatcher.__call__ / rle_matcher        empanada/inference/matcher.py:136-326
//   forward_matching / backward_matching     empanada/inference/patterns.py:55-121
//   InstanceTracker.update (boxes, order)    empanada/inference/tracker.py:11-100
// Geometry never enters here: IoU/IoA of merged objects are sums over their (disjoint)
// components, so the overlap table is sufficient. The Hungarian step is SciPy's
// linear_sum_assignment(maximize=True) (un-vendored dependency of the reference,
// matcher.py:213): the same shortest-augmenting-path algorithm (Crouse 2016) is run here on
// each connected block of the sparse IoU matrix, rows/columns kept in matrix order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_map>
#include <vector>

extern "C" int be_set_error(const char* msg);

namespace {

struct Obj {
  int label;
  long long area;
  int box[4];            // y0 x0 y1 x1 (half open)
  std::vector<int> ccs;  // component ids (1-based) of this slice that make up the object
};

// Rectangular LAP, minimisation, shortest augmenting paths (Crouse, "On implementing 2D
// rectangular assignment algorithms", 2016) - the algorithm behind scipy.optimize.linear_sum_assignment.
// cost is nr x nc row-major with nr <= nc. Returns col4row.
static void lap_min(int nr, int nc, const std::vector<double>& cost, std::vector<int>& col4row) {
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(nr, 0.0), v(nc, 0.0), shortest(nc);
  std::vector<int> path(nc, -1), row4col(nc, -1), remaining(nc);
  std::vector<char> SR(nr), SC(nc);
  col4row.assign(nr, -1);
  for (int cur = 0; cur < nr; ++cur) {
    double min_val = 0.0;
    int i = cur;
    int num_remaining = nc;
    for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
    std::fill(SR.begin(), SR.end(), 0);
    std::fill(SC.begin(), SC.end(), 0);
    std::fill(shortest.begin(), shortest.end(), INF);
    int sink = -1;
    while (sink == -1) {
      int index = -1;
      double lowest = INF;
      SR[i] = 1;
      for (int it = 0; it < num_remaining; ++it) {
        const int j = remaining[it];
        const double r = min_val + cost[static_cast<size_t>(i) * nc + j] - u[i] - v[j];
        if (r < shortest[j]) { path[j] = i; shortest[j] = r; }
        if (shortest[j] < lowest || (shortest[j] == lowest && row4col[j] == -1)) {
          lowest = shortest[j];
          index = it;
        }
      }
      min_val = lowest;
      if (min_val == INF) return;  // infeasible (cannot happen for finite costs)
      const int j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = 1;
      remaining[index] = remaining[--num_remaining];
    }
    u[cur] += min_val;
    for (int r = 0; r < nr; ++r)
      if (SR[r] && r != cur) u[r] += min_val - shortest[col4row[r]];
    for (int j = 0; j < nc; ++j)
      if (SC[j]) v[j] -= min_val - shortest[j];
    int j = sink;
    while (true) {
      const int r = path[j];
      row4col[j] = r;
      std::swap(col4row[r], j);
      if (r == cur) break;
    }
  }
}

struct Entry { int row, col; long long inter; };

// One matcher step: relabel `match` objects against `target` objects.
// entries: sparse intersections (row = target index, col = match index), duplicates allowed.
static void match_step(const std::vector<Obj>& target, const std::vector<Obj>& match,
                       std::vector<Entry>& entries, double iou_thr, float ioa_thr, bool assign_new,
                       int& next_label, std::vector<int>& new_labels) {
  const int n = static_cast<int>(target.size()), m = static_cast<int>(match.size());
  new_labels.assign(m, 0);
  std::vector<int> matched_row(m, -1);
  // aggregate duplicate (row, col)
  std::sort(entries.begin(), entries.end(), [](const Entry& a, const Entry& b) {
    return a.row != b.row ? a.row < b.row : a.col < b.col;
  });
  std::vector<Entry> agg;
  for (const Entry& e : entries) {
    if (!agg.empty() && agg.back().row == e.row && agg.back().col == e.col) agg.back().inter += e.inter;
    else agg.push_back(e);
  }
  std::vector<double> iou(agg.size());
  std::vector<float> ioa(agg.size());
  for (size_t k = 0; k < agg.size(); ++k) {
    const long long inter = agg[k].inter;
    const long long uni = target[agg[k].row].area + match[agg[k].col].area - inter;
    iou[k] = static_cast<double>(inter) / static_cast<double>(uni);
    ioa[k] = static_cast<float>(static_cast<double>(inter) / static_cast<double>(match[agg[k].col].area));
  }
  if (n > 0 && m > 0 && !agg.empty()) {
    // connected blocks of the bipartite graph (union-find over rows [0,n) and cols [n,n+m))
    std::vector<int> parent(n + m);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    for (const Entry& e : agg) {
      const int a = find(e.row), b = find(n + e.col);
      if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }
    std::unordered_map<int, std::vector<size_t>> blocks;
    for (size_t k = 0; k < agg.size(); ++k) blocks[find(agg[k].row)].push_back(k);
    for (auto& kv : blocks) {
      const std::vector<size_t>& idx = kv.second;
      if (idx.size() == 1) {  // isolated positive entry: always part of the optimum
        const Entry& e = agg[idx[0]];
        if (iou[idx[0]] >= iou_thr) matched_row[e.col] = e.row;
        continue;
      }
      std::vector<int> rows, cols;
      for (size_t k : idx) { rows.push_back(agg[k].row); cols.push_back(agg[k].col); }
      std::sort(rows.begin(), rows.end()); rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
      std::sort(cols.begin(), cols.end()); cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
      const int br = static_cast<int>(rows.size()), bc = static_cast<int>(cols.size());
      std::vector<double> dense(static_cast<size_t>(br) * bc, 0.0);
      for (size_t k : idx) {
        const int r = static_cast<int>(std::lower_bound(rows.begin(), rows.end(), agg[k].row) - rows.begin());
        const int c = static_cast<int>(std::lower_bound(cols.begin(), cols.end(), agg[k].col) - cols.begin());
        dense[static_cast<size_t>(r) * bc + c] = iou[k];
      }
      // scipy: maximise == minimise(-cost); tall matrices are transposed first
      const bool transpose = bc < br;
      const int nr = transpose ? bc : br, nc = transpose ? br : bc;
      std::vector<double> cost(static_cast<size_t>(nr) * nc);
      for (int r = 0; r < br; ++r)
        for (int c = 0; c < bc; ++c) {
          const double val = -dense[static_cast<size_t>(r) * bc + c];
          if (transpose) cost[static_cast<size_t>(c) * nc + r] = val; else cost[static_cast<size_t>(r) * nc + c] = val;
        }
      std::vector<int> col4row;
      lap_min(nr, nc, cost, col4row);
      for (int a = 0; a < nr; ++a) {
        if (col4row[a] < 0) continue;
        const int r = transpose ? col4row[a] : a, c = transpose ? a : col4row[a];
        if (dense[static_cast<size_t>(r) * bc + c] >= iou_thr) matched_row[cols[c]] = rows[r];
      }
    }
  }
  // per-column IoA maximum (float32 matrix semantics: zeros everywhere else, first max row wins)
  std::vector<float> ioa_max(m, 0.0f);
  std::vector<int> ioa_arg(m, 0);
  if (n > 0 && m > 0) {
    for (size_t k = 0; k < agg.size(); ++k) {  // agg is sorted by row, so ">" keeps the first max
      const int c = agg[k].col;
      if (ioa[k] > ioa_max[c]) { ioa_max[c] = ioa[k]; ioa_arg[c] = agg[k].row; }
    }
  }
  for (int i = 0; i < m; ++i) {
    if (matched_row[i] >= 0) {
      new_labels[i] = target[matched_row[i]].label;
    } else if (n > 0 && m > 0 && ioa_max[i] >= ioa_thr) {
      new_labels[i] = target[ioa_arg[i]].label;
    } else if (assign_new) {
      new_labels[i] = next_label++;
    } else {
      new_labels[i] = match[i].label;
    }
  }
}

static void merge_by_label(const std::vector<Obj>& match, const std::vector<int>& new_labels,
                           std::vector<Obj>& out) {
  out.clear();
  std::unordered_map<int, int> pos;
  for (size_t i = 0; i < match.size(); ++i) {
    auto it = pos.find(new_labels[i]);
    if (it == pos.end()) {
      pos[new_labels[i]] = static_cast<int>(out.size());
      Obj o = match[i];
      o.label = new_labels[i];
      out.push_back(std::move(o));
    } else {
      Obj& o = out[it->second];
      o.area += match[i].area;
      o.box[0] = std::min(o.box[0], match[i].box[0]);
      o.box[1] = std::min(o.box[1], match[i].box[1]);
      o.box[2] = std::max(o.box[2], match[i].box[2]);
      o.box[3] = std::max(o.box[3], match[i].box[3]);
      o.ccs.insert(o.ccs.end(), match[i].ccs.begin(), match[i].ccs.end());
    }
  }
}

}  // namespace

extern "C" {

// n_cc        [n_slices]              components per slice
// cc_table    [n_slices][cap][5]      area, y0, x0, y1, x1
// pair_keys   slice(24)|prev_cc(20)|cur_cc(20), pair_vals = overlapping pixels (slice vs slice-1)
// axis        0 xy, 1 xz, 2 yz        (tracker.py:11-23 box lifting)
// lut         [n_slices][lut_stride]  out: component id -> final tracked label
// inst_*      out, tracker insertion order (first arrival in the backward sweep)
int be_match_replay(int n_slices, const int* n_cc, const int* cc_table, int cap,
                    const unsigned long long* pair_keys, const int* pair_vals, long long n_pairs,
                    int class_id, int label_divisor, double iou_thr, double ioa_thr, int axis,
                    int* lut, int lut_stride, int* inst_labels, long long* inst_sizes,
                    int* inst_boxes, int max_inst, int* n_inst) {
  if (n_slices <= 0) { *n_inst = 0; return 0; }
  for (int s = 0; s < n_slices; ++s)
    if (n_cc[s] > cap || n_cc[s] >= lut_stride || n_cc[s] >= (1 << 20))
      return be_set_error("component count exceeds table capacity; re-run with a larger cap");
  // bucket pairs by slice
  std::vector<long long> pstart(n_slices + 1, 0);
  for (long long k = 0; k < n_pairs; ++k) {
    const int s = static_cast<int>(pair_keys[k] >> 40);
    if (s < 0 || s >= n_slices) return be_set_error("overlap table: slice index out of range");
    ++pstart[s + 1];
  }
  for (int s = 0; s < n_slices; ++s) pstart[s + 1] += pstart[s];
  std::vector<long long> porder(n_pairs), fill(pstart.begin(), pstart.end() - 1);
  for (long long k = 0; k < n_pairs; ++k) porder[fill[pair_keys[k] >> 40]++] = k;

  const int base_label = class_id * label_divisor;
  const float ioa_thr_f = static_cast<float>(ioa_thr);
  auto slice_ccs = [&](int s, std::vector<Obj>& out) {
    out.clear();
    const int* t = cc_table + static_cast<size_t>(s) * cap * 5;
    for (int c = 1; c <= n_cc[s]; ++c) {
      Obj o;
      o.label = base_label + c;
      o.area = t[(c - 1) * 5];
      o.box[0] = t[(c - 1) * 5 + 1]; o.box[1] = t[(c - 1) * 5 + 2];
      o.box[2] = t[(c - 1) * 5 + 3]; o.box[3] = t[(c - 1) * 5 + 4];
      o.ccs.push_back(c);
      out.push_back(std::move(o));
    }
  };

  // ---------------- forward pass (assign_new = True)
  std::vector<std::vector<Obj>> fwd(n_slices);
  int next_label = base_label + 1;
  std::vector<int> owner;  // component id (slice s-1) -> index into fwd[s-1]
  std::vector<Entry> entries;
  std::vector<int> new_labels;
  for (int s = 0; s < n_slices; ++s) {
    std::vector<Obj> cur;
    slice_ccs(s, cur);
    if (s == 0) {
      fwd[0] = cur;
      if (!cur.empty()) next_label = cur.back().label + 1;  // max(labels) + 1, labels ascend
    } else {
      const std::vector<Obj>& tgt = fwd[s - 1];
      owner.assign(n_cc[s - 1] + 1, -1);
      for (size_t i = 0; i < tgt.size(); ++i)
        for (int c : tgt[i].ccs) owner[c] = static_cast<int>(i);
      entries.clear();
      for (long long k = pstart[s]; k < pstart[s + 1]; ++k) {
        const unsigned long long key = pair_keys[porder[k]];
        const int q = static_cast<int>((key >> 20) & 0xFFFFF), c = static_cast<int>(key & 0xFFFFF);
        if (q < 1 || q > n_cc[s - 1] || c < 1 || c > n_cc[s]) return be_set_error("overlap table: component id out of range");
        entries.push_back({owner[q], c - 1, pair_vals[porder[k]]});
      }
      match_step(tgt, cur, entries, iou_thr, ioa_thr_f, true, next_label, new_labels);
      merge_by_label(cur, new_labels, fwd[s]);
    }
  }

  // ---------------- backward pass (assign_new = False) + tracker
  std::unordered_map<int, int> inst_pos;
  int ninst = 0;
  std::vector<Obj> bwd_next, bwd_cur;
  std::vector<int> owner_next, owner_cur;
  int dummy_next = 0;
  for (int s = n_slices - 1; s >= 0; --s) {
    if (s == n_slices - 1) {
      bwd_cur = fwd[s];
    } else {
      const std::vector<Obj>& match = fwd[s];
      owner_cur.assign(n_cc[s] + 1, -1);
      for (size_t i = 0; i < match.size(); ++i)
        for (int c : match[i].ccs) owner_cur[c] = static_cast<int>(i);
      owner_next.assign(n_cc[s + 1] + 1, -1);
      for (size_t i = 0; i < bwd_next.size(); ++i)
        for (int c : bwd_next[i].ccs) owner_next[c] = static_cast<int>(i);
      entries.clear();
      for (long long k = pstart[s + 1]; k < pstart[s + 2 > n_slices ? n_slices : s + 2]; ++k) {
        const unsigned long long key = pair_keys[porder[k]];
        const int q = static_cast<int>((key >> 20) & 0xFFFFF), c = static_cast<int>(key & 0xFFFFF);
        entries.push_back({owner_next[c], owner_cur[q], pair_vals[porder[k]]});
      }
      match_step(bwd_next, match, entries, iou_thr, ioa_thr_f, false, dummy_next, new_labels);
      merge_by_label(match, new_labels, bwd_cur);
    }
    // tracker.update(rle_seg, s): dict order of bwd_cur
    int* l = lut + static_cast<size_t>(s) * lut_stride;
    std::memset(l, 0, sizeof(int) * lut_stride);
    for (const Obj& o : bwd_cur) {
      for (int c : o.ccs) l[c] = o.label;
      int b3[6];
      if (axis == 0) { b3[0] = s; b3[1] = o.box[0]; b3[2] = o.box[1]; b3[3] = s + 1; b3[4] = o.box[2]; b3[5] = o.box[3]; }
      else if (axis == 1) { b3[0] = o.box[0]; b3[1] = s; b3[2] = o.box[1]; b3[3] = o.box[2]; b3[4] = s + 1; b3[5] = o.box[3]; }
      else { b3[0] = o.box[0]; b3[1] = o.box[1]; b3[2] = s; b3[3] = o.box[2]; b3[4] = o.box[3]; b3[5] = s + 1; }
      auto it = inst_pos.find(o.label);
      if (it == inst_pos.end()) {
        if (ninst >= max_inst) return be_set_error("instance table capacity exceeded; re-run with a larger max_inst");
        inst_pos[o.label] = ninst;
        inst_labels[ninst] = o.label;
        inst_sizes[ninst] = o.area;
        std::memcpy(inst_boxes + ninst * 6, b3, sizeof(b3));
        ++ninst;
      } else {
        const int p = it->second;
        inst_sizes[p] += o.area;
        int* bb = inst_boxes + p * 6;
        for (int d = 0; d < 3; ++d) { bb[d] = std::min(bb[d], b3[d]); bb[d + 3] = std::max(bb[d + 3], b3[d + 3]); }
      }
    }
    bwd_next.swap(bwd_cur);
  }
  *n_inst = ninst;
  return 0;
}

}  // extern "C"
