/* C-ABI of libb200_empanada.so - the B200 (sm_100a) implementation of empanada's panoptic
 * inference hot path. Plain pointers and sizes only; every device pointer is caller-owned; every
 * GPU entry point takes the cudaStream_t to launch on (passed as void*); no hidden synchronisation.
 * Convention: return 0 on success, negative on failure; be_last_error() returns the thread-local
 * message. Each group cites the reference interface (paths under volume-em/empanada-napari) it
 * replaces; the reference-side binding is shown in INTEGRATION.md.
 */
#ifndef B200_EMPANADA_H
#define B200_EMPANADA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* be_stream;  /* cudaStream_t */

int be_version(void);
const char* be_last_error(void);

/* ---- piece (1): network forward. Replaces `self.model(image, render_steps, interpolate_ins)`
 * (empanada/inference/engines.py:250) = QuantizablePanopticDeepLabPR.forward
 * (empanada/models/quantization/panoptic_deeplab.py:238). The forward pass of one network at one
 * (batch, H, W) is a launch list: be_op_* with list != NULL records, list == NULL launches now. */
int be_oplist_create(void** list);
int be_oplist_destroy(void* list);
int be_oplist_launches(void* list);
int be_oplist_size(void* list);
/* replay everything after the slice-gather prefix of the list as one CUDA graph (launch-bound
 * lists: single small tiles); the first replay after enabling runs plainly, the second captures */
int be_oplist_set_graph(void* list, int enable);
int be_oplist_run(void* list, const void* volume, long long stride_slice, long long stride_y,
                  long long stride_x, int first_slice, be_stream st);
int be_oplist_run_timed(void* list, const void* volume, long long stride_slice,
                        long long stride_y, long long stride_x, int first_slice, float* ms_per_op,
                        int max_ops, be_stream st);
/* implicit-GEMM convolution on tcgen05/TMEM fed by TMA (every nn.Conv2d / Conv1d with groups == 1:
 * encoders/resnet.py:24-33, decoders/aspp.py:16-22,67-86, blocks.py conv_bn_act, point_rend.py:160-166);
 * NHWC bf16 in/out, weights [Cout][R*S*Cin] bf16, fused bias / residual / ReLU|SiLU / 1x1 head */
int be_op_conv(void* list, const void* in, long long in_ld, int B, int Hi, int Wi, int Cin,
               const void* w, int Cout, int R, int S, int stride, int dil, int pad, int Ho, int Wo,
               void* out, long long out_ld, int out_coff, float* out_f32, long long out_f32_ld,
               int out_f32_planar, const float* bias, long long bias_img_stride,
               const void* residual, long long res_ld, int act, const float* head_w,
               const float* head_b, float* head_out, int head_n, be_stream st);
/* slice gather + Preprocessor.normalize + factor_pad + conv1 7x7/2 + BN + ReLU
 * (data/volume_dataset.py:37-53, empanada_napari/utils.py:170-201, inference/postprocess.py:26-36,
 * encoders/resnet.py:217-220). `elem` = element type of the volume: 0 u8, 1 i8, 2 u16, 3 i16,
 * 4 u32, 5 i32, 6 u64, 7 i64 (any integer dtype, utils.py:189-201), 8 = fp32 that is already
 * normalised (the engine-level API, engines.py:300-325; pass mean255 0 and inv_std255 1);
 * mean255 / inv_std255 are mean*iinfo.max and 1/(std*iinfo.max) of that dtype; strides are in
 * elements */
int be_op_stem(void* list, int B, int h, int w, int H, int W, float mean255, float inv_std255,
               const float* wt_49x64, const float* bias64, void* out, const void* vol,
               long long stride_slice, long long stride_y, long long stride_x, int first_slice,
               int elem, be_stream st);
int be_op_maxpool(void* list, const void* in, int B, int Hi, int Wi, int C, void* out, int Ho, int Wo,
                  be_stream st);                                   /* encoders/resnet.py:221 */
/* be_op_stem followed by be_op_maxpool in ONE kernel (encoders/resnet.py:217-221): `out` is the
 * quarter-resolution [B][H/4][W/4][64] map; the half-resolution map never reaches HBM */
int be_op_stem_pool(void* list, int B, int h, int w, int H, int W, float mean255, float inv_std255,
                    const float* wt_49x64, const float* bias64, void* out, const void* vol,
                    long long stride_slice, long long stride_y, long long stride_x, int first_slice,
                    int elem, be_stream st);
/* depthwise k x k (blocks.py:15-35); optional fused producer: channels [0,Cup) are the
 * align_corners=True bilinear upsampling of `up` (decoders/panoptic_deeplab.py:76-77) */
int be_op_dwconv(void* list, const void* in, long long in_ld, int B, int H, int W, int C, int k,
                 const float* wt, void* out, long long out_ld, const void* up, int Cup, int Hu,
                 int Wu, be_stream st);
/* ConvTranspose2d(k=2, s=2) + folded BN + activation (blocks.py conv_transpose_bn_act,
 * decoders/bifpn.py:213-218) as one 1x1 implicit GEMM with N = 4*Cout and a pixel-shuffle
 * epilogue; w rows n = (2*dy+dx)*Cout + co, bias [4*Cout] */
int be_op_convt2x2(void* list, const void* in, long long in_ld, int B, int Hi, int Wi, int Cin,
                   const void* w, int Cout, void* out, long long out_ld, int out_coff,
                   const float* bias, int act, be_stream st);
/* BiFPN fast-normalised fusion (decoders/bifpn.py:52-68,106-133):
 * out = (w1*R(a) + w2*b [+ w3*c]) / denom, R = identity (mode 0), nearest x2 (1), MaxPool2d(3,2,1) (2) */
int be_op_bifpn_fuse(void* list, const void* a, long long a_ld, int mode, int Ha, int Wa,
                     const void* b, long long b_ld, const void* c, long long c_ld, float w1,
                     float w2, float w3, float denom, int B, int H, int W, int C, void* out,
                     long long out_ld, be_stream st);
int be_op_bilinear(void* list, const void* in, long long in_ld, int B, int Hi, int Wi, int C,
                   void* out, long long out_ld, int out_coff, int Ho, int Wo,
                   be_stream st);                                  /* decoders/panoptic_deeplab.py:76 */
int be_op_aspp_pool_bias(void* list, const void* in, int B, int HW, int C, const float* w_pool,
                         int Cmid, const float* w_proj_pool, const float* bias_proj, int N,
                         float* pooled, float* mid, float* bias_out,
                         be_stream st);                            /* decoders/aspp.py:30-48,97-102 */
/* Interpolate2d(4, bilinear, align_corners=True) of ctr_hmp / offsets for `interpolate_ins`
 * (fine boundaries; quantization/panoptic_deeplab.py:233-234): planar fp32 [planes][h][w] -> [planes][4h][4w] */
int be_up4(const float* in, int planes, int h, int w, float* out, be_stream st);
/* `resize_by_factor` (empanada/data/utils/transforms.py:9-21; volume_dataset.py:42-47): the N
 * slices (h x w, uint8, element strides stride_y / stride_x, slice stride stride_s) of a volume
 * down-sampled to dh x dw = ceil(h/f) x ceil(w/f) with OpenCV's 8-bit INTER_LINEAR arithmetic
 * (third-party, restated; see csrc/model_kernels.cu). out: [N][dh][dw] uint8, contiguous. */
int be_resize_linear_u8(const uint8_t* vol, long long stride_s, long long stride_y, long long stride_x,
                        int N, int h, int w, int dh, int dw, uint8_t* out, be_stream st);
int be_op_up2(void* list, const float* in, int B, int h, int w, float* out, be_stream st);
int be_op_topk(void* list, const float* x, int B, int n, int k, unsigned* state, unsigned* hist,
               int* idx_out, be_stream st);                        /* point_rend.py:110-137 */
int be_op_pr_sample(void* list, const int* idx, int B, int k, int Hf, int Wf, const float* coarse,
                    const void* feat, int h4, int w4, int C, void* P, void* P2, int ldp,
                    float* coarse_pts, be_stream st);              /* point_rend.py:24-59,255-258 */
int be_op_pr_predict(void* list, const void* X, int ldp, int C, const float* coarse_pts,
                     const float* wp, float bias, const int* idx, int B, int k, int HWf, float* sem,
                     be_stream st);                                /* point_rend.py:181-188,260-267 */

/* ---- piece (4): _MedianQueue + logits_to_prob + _harden_seg (engines.py:22-30,47-90,115-121,351-361) */
int be_median_push(const float* logits, int B, int H, int W, int ks, float* hist, int n_hist,
                   int slice0, float conf_thr, int is_prob, uint8_t* hard, float* prob_out,
                   be_stream st);
int be_median_flush(const float* hist, int n_hist, int ks, int H, int W, int n_slices_total,
                    float conf_thr, uint8_t* hard, float* prob_out, be_stream st);
/* ---- piece (2): find_instance_center (postprocess.py:39-76) */
int be_centers(const float* ctr, int B, int h4, int w4, float thr, int k, int* centers, int cap,
               int* counts, int* chunk_counts_scratch, be_stream st);
/* ---- piece (3): group_pixels / get_instance_cells (postprocess.py:79-169, engines.py:258-275) and
 * get_panoptic_seg / merge_semantic_and_instance (engines.py:278-298, postprocess.py:224-296) */
int be_group_pixels(const float* off, const int* centers, int cap, const int* counts, int B, int h4,
                    int w4, float step, int* cells4, be_stream st);
int be_rank_ids(int* present, int B, int cap, int label_divisor, int class_id, be_stream st);
int be_merge_pan(const uint8_t* hard, const int* cells4, int B, int H, int W, int h, int w,
                 int scale, int cap, int label_divisor, int class_id, int void_label, int* present,
                 int* pan, be_stream st);
/* ---- piece (5): connected_components / pan_seg_to_rle_seg (inference/rle.py:18-86), overlap
 * tables standing for rle_intersection (array_utils.py:375-407), matcher + tracker replay
 * (matcher.py:136-326, patterns.py:55-121, tracker.py:11-123), relabel, rle_encode
 * (array_utils.py:213-239) */
int be_cc_label(const int* pan, int B, int h, int w, int lo, int hi, int* L, int* chunk_counts,
                int* cc_out, int* n_cc, int cap, int* table, be_stream st);
int be_hash_clear(unsigned long long* keys, int* vals, unsigned long long cap, be_stream st);
int be_pair_overlap(const int* cc_plane, int h, int w, int s0, int s1, unsigned long long* keys,
                    int* vals, unsigned long long cap, int* overflow, be_stream st);
int be_hash_compact(const unsigned long long* keys, const int* vals, unsigned long long cap,
                    unsigned long long* out_keys, int* out_vals, int out_cap, int* cursor,
                    be_stream st);
/* be_match_replay: the Hungarian step (scipy.optimize.linear_sum_assignment, matcher.py:213) is
 * solved per connected block of the sparse IoU matrix; a block whose optimum is not certified
 * unique by its dual variables makes the step replay SciPy's run on the full matrix, so exact IoU
 * ties resolve as in the reference. iou_thr must be > 0 (the engines use 0.25 as the reference). */
int be_match_replay(int n_slices, const int* n_cc, const int* cc_table, int cap,
                    const unsigned long long* pair_keys, const int* pair_vals, long long n_pairs,
                    int class_id, int label_divisor, double iou_thr, double ioa_thr, int axis,
                    int* lut, int lut_stride, int* inst_labels, long long* inst_sizes,
                    int* inst_boxes, int max_inst, int* n_inst);            /* host function */
/* diagnostics of the last be_match_replay call: out[0] matcher steps, out[1] steps with a
 * multi-entry block, out[2] steps replayed on the full matrix */
int be_match_replay_stats(long long* out);                                  /* host function */
int be_relabel(const int* cc_batch, int B, int h, int w, int s0, const int* lut, int lut_stride,
               int* dst, long long stride_s, long long stride_y, long long stride_x, be_stream st);
int be_runs_count(const int* img, long long n, long long seg_len, int* chunk_counts, be_stream st);
int be_scan_i32_to_i64(const int* counts, long long* offsets, long long n, void* temp,
                       size_t temp_bytes, size_t* temp_needed, be_stream st);
int be_runs_write(const int* img, long long n, long long seg_len, const long long* chunk_offsets,
                  int* out_label, long long* out_start, long long* out_end, long long out_cap,
                  be_stream st);
int be_sort_runs(const unsigned long long* keys_in, unsigned long long* keys_out, const int* idx_in,
                 int* idx_out, int n, void* temp, size_t temp_bytes, size_t* temp_needed,
                 be_stream st);
/* ---- pieces (3), (5), (6b) on ROW RUNS (csrc/run_kernels.cu), the batched plane path: grouping
 * only where a thing pixel exists + presence flags (postprocess.py:119-169,224-296); maximal row
 * runs of equal panoptic value inside a class range, straight from the mask + cell ids (the
 * pan_seg of engines.py:278-298 restricted as in rle.py:60-66 never reaches HBM); 8-connected
 * equal-value components in raster order as a union-find over runs (rle.py:18-24); area / bbox per
 * component (rle.py:75-83); adjacent-slice overlaps (array_utils.py:375-407); and the relabelled
 * volume written once from the runs (patterns.py:204-213). Run arrays are sized by the caller from
 * stats[0] (total runs) after be_rowruns_count. */
int be_group_flags(const uint8_t* hard, const float* off, const int* centers, int cap,
                   const int* counts, int B, int H, int W, int scale, int step, int* cells, int* present,
                   be_stream st);
/* set pixels per slice of the hardened mask: the stuff-area test of merge_semantic_and_instance
 * (postprocess.py:283-294) for semantic-only planes (engines.py thing_list = []) */
int be_slice_area(const uint8_t* hard, int B, int H, int W, int* area, be_stream st);
/* be_rowruns_count first resolves `cells` IN PLACE from centre ids to final panoptic values
 * (newid lookup, void label, class range); be_rowruns_write takes the resolved array. W must be a
 * multiple of 16. counts: workspace [2 * B * chunks], chunks = ceil(h * ceil(w/16) / 256). */
int be_rowruns_count(const uint8_t* hard, int* cells, const int* newid, int B, int H, int W,
                     int h, int w, int scale, int cap, int void_label, int lo, int hi, int* counts,
                     int* n_runs, int* slice_off, int* stats, int* row_ptr, be_stream st);
int be_rowruns_write(const uint8_t* hard, const int* cells, const int* newid, int B, int H, int W,
                     int h, int w, int scale, int cap, int void_label, int lo, int hi, int* counts,
                     const int* slice_off, int* row_ptr, int* run_yx, int* run_x1, int* run_val,
                     int* L, be_stream st);
int be_runs_cc(const int* row_ptr, const int* run_yx, const int* run_x1, const int* run_val,
               const int* slice_off, int B, int h, int max_runs, int* L, int* run_cc, int* n_cc,
               be_stream st);
int be_runs_stats(const int* run_yx, const int* run_x1, const int* run_cc, const int* slice_off,
                  int B, int max_runs, int cap, int* table, be_stream st);
int be_runs_overlap(const int* row_ptr, const int* run_yx, const int* run_x1, const int* run_cc,
                    const int* slice_off, int B, int h, int max_runs, int key_s0,
                    unsigned long long* keys, int* vals, unsigned long long cap, int* overflow,
                    be_stream st);
int be_runs_paint(const int* row_ptr, const int* run_yx, const int* run_x1, const int* run_cc,
                  const int* slice_off, const int* lut, int lut_stride, int add, int b0, int nb, int h,
                  int w, int* dst, long long stride_s, long long stride_y, long long stride_x,
                  be_stream st);
/* ---- piece (6): merge_objects_from_trackers (consensus.py:233-287,449-460), fill
 * (array_utils.py:754-765), filters on the painted volume */
int be_plane_pairs(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                   const int* lc, int na, int nb, int nc, long long n, int W,
                   unsigned long long* keys, int* vals, unsigned long long cap, int* overflow,
                   be_stream st);
int be_vote_stats(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                  const int* lc, int na, int nb, int nc, long long n, int W, const int* memb_off,
                  const int* memb_list, int vote_thr, int* sizes, unsigned long long* keys,
                  int* vals, unsigned long long cap, int* overflow, be_stream st);
int be_vote_paint(const int* va, const int* vb, const int* vc, const int* la, const int* lb,
                  const int* lc, int na, int nb, int nc, long long n, int W, const int* memb_off,
                  const int* memb_list, int vote_thr, const int* cid_final, int* out,
                  long long* side_voxel, int* side_id, int side_cap, int* side_count, be_stream st);
int be_label_hist(const int* vol, long long n, int W, int nbins, int* hist, be_stream st);
int be_lut_inplace(int* vol, long long n, const int* lut, int nlut, be_stream st);

/* ---- tracker-level morphology (csrc/morph_kernels.cu; empanada/inference/filters.py:154-210 as
 * applied by Engine3d.infer_on_axis, empanada_napari/inference.py:560-570). be_morph3d: one grey
 * erosion (op 0) / dilation (op 1) pass with the 3-D cross on a (D,H,W) int32 label volume.
 * be_range_keep: labels outside [lo, hi) -> 0 (filters.py:96-103). be_runs3d_cc: 26-connected
 * equal-value components over the row runs of the volume (skimage.measure.label in 3-D,
 * filters.py:15-20,105-107), root = raster-first run of the component. be_fill_holes: the
 * per-slice, per-label bounding-box hole filling of filters.py:174-210 (scipy binary_fill_holes),
 * labels ascending per slice in CSR form (slice_off [D+1], labels, boxes y0 x0 y1 x1). */
int be_morph3d(const int* src, int* dst, int D, int H, int W, int op, be_stream st);
int be_range_keep(int* vol, long long n, int lo, int hi, be_stream st);
int be_runs3d_cc(const long long* run_start, const long long* run_end, const int* run_val, int n_runs,
                 int D, int H, int W, int* row_ptr, int* L, int* root, be_stream st);
int be_fill_holes(int* vol, uint8_t* scratch, int D, int H, int W, const int* slice_off,
                  const int* labels, const int* boxes, be_stream st);

/* ---- host: cluster decisions of the connected components of the consensus instance graph
 * (csrc/cluster_graph.cpp; empanada/consensus.py:35-142 on graph.subgraph(comp)) for n_comp
 * components in CSR layout: nodes ascending per component, edges (a, b, iou, overlap) in the
 * instance graph's edge-insertion order. n_clusters [n_comp]; totals = {clusters, cluster members};
 * the lists are fetched with be_components_clusters_fetch (cluster graph node order; clusters may
 * share nodes). float_sum: 1 when the interpreter's builtin sum() is the compensated one of
 * Python >= 3.12 (the reference averages edge weights with it). */
int be_components_clusters(int n_comp, const int* node_off, const int* nodes, const int* edge_off,
                           const int* ea, const int* eb, const double* eiou, const long long* eov,
                           long long n_nodes_total, double cluster_iou_thr, double min_iou,
                           double min_overlap, int float_sum, int* n_clusters, long long* totals);
int be_components_clusters_fetch(int* cluster_sizes, int* members);

#ifdef __cplusplus
}
#endif
#endif
