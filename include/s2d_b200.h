/*
 * s2d_b200.h -- C ABI of libs2d_b200.so, the B200 (sm_100a) implementation of the
 * Sparse2Dense / CenterPoint hot path:  voxelize -> sparse 3-D conv backbone -> BEV.
 *
 * The reference has no C/FFI boundary of its own on this path: the
 * arithmetic sits behind three Python operator APIs.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch in
 *     this repo) owns every allocation, including workspaces sized by the *_workspace_bytes
 *     queries;
 *   - every call is asynchronous on the cudaStream_t passed as `stream` (void* here so the
 *     header needs no CUDA include) and never synchronises the device;
 *   - return value: 0 on success, negative on error (never exit()/abort(), unlike
 *     det3d/ops/iou3d_nms/src/iou3d_nms.cpp:14-25); s2d_last_error() gives the text of the
 *     last failure on the calling thread;
 *   - sparse tensors are (features f32 [N,C] row-major, coors i32 [N,4] = (batch,z,y,x));
 *   - rulebook tables are "k-major": tbl[k * tbl_stride + out_row] = in_row or -1, with k the
 *     row-major kernel offset (kz,ky,kx), exactly the offset order of spconv's
 *     weight[kD,kH,kW,Cin,Cout].
 */
#ifndef S2D_B200_H_
#define S2D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2D_OK 0
#define S2D_ERR_INVALID (-1)   /* bad argument                    */
#define S2D_ERR_CUDA (-2)      /* a CUDA runtime call failed      */
#define S2D_ERR_WORKSPACE (-3) /* workspace or capacity too small */
#define S2D_ERR_UNSUPPORTED (-4)

#define S2D_MAX_BATCH 64

int s2d_version(void);
const char* s2d_last_error(void);
/* kernels launched by this library in this process so far (bench.py reports the per-run delta) */
unsigned long long s2d_kernel_launches(void);
/* Copies n ints (row counts and similar control values) from device memory into MAPPED pinned host memory with a
 * kernel store, not a DMA copy, so that the read never queues behind a bulk device->host transfer; the caller
 * waits for the stream (an event) before reading dst_host_mapped. */
int s2d_export_i32(const int* src, int n, int* dst_host_mapped, void* stream);

/* ---------------------------------------------------------------------------------------
 * Voxelizer.  Replaces points_to_voxel(points, voxel_size, coors_range, max_points,
 * reverse_index=True, max_voxels) -- det3d/ops/point_cloud/point_cloud_ops.py:112-184
 * (kernel :7-55), as called by VoxelGenerator.generate (det3d/core/input/voxel_generator.py:
 * 19-30) -- for a whole batch at once, plus the batch-index column that collate_kitti
 * prepends (det3d/torchie/parallel/collate.py:137-144) and, optionally, the reader
 * VoxelFeatureExtractorV3.forward (det3d/models/readers/voxel_encoder.py:17-24).
 *
 * Semantics are bit-exact with the reference's serial loop: per scene, voxel ids follow the
 * order of first appearance of a voxel in the point list; a voxel keeps its first
 * max_points points in point order; once max_voxels voxels exist, points that would open a
 * new voxel are dropped while existing voxels keep filling.
 *
 *   points               f32 [n_points, F], scenes concatenated
 *   scene_offsets_host   HOST i32 [batch+1], scene b owns points [off[b], off[b+1])
 *   range_host, vsize_host   HOST f32 [6] (xmin,ymin,zmin,xmax,ymax,zmax), [3] (x,y,z)
 *   voxels      (nullable) f32 [batch*max_voxels, max_points, F], rows past the total untouched
 *   coors                i32 [batch*max_voxels, 4] (b,z,y,x)
 *   num_points           i32 [batch*max_voxels]
 *   mean        (nullable) f32 [batch*max_voxels, mean_channels]: sum of the kept points /
 *                        num_points over the first mean_channels features
 *   voxel_offsets        i32 [batch+1]: exclusive prefix of the per-scene voxel counts;
 *                        voxel_offsets[batch] is the total number of rows written
 * ------------------------------------------------------------------------------------- */
size_t s2d_voxelize_workspace_bytes(int n_points, int batch, int max_points, int max_voxels);
int s2d_voxelize(const float* points, const int* scene_offsets_host, int n_points, int batch, int F,
                 const float* range_host, const float* vsize_host, int max_points, int max_voxels,
                 float* voxels, int* coors, int* num_points, float* mean, int mean_channels,
                 int* voxel_offsets, void* workspace, size_t workspace_bytes, void* stream);

/* VoxelFeatureExtractorV3.forward alone (voxel_encoder.py:17-24) for voxels produced elsewhere. */
int s2d_voxel_mean(const float* voxels, const int* num_points, int n_voxels, int max_points, int F,
                   int channels, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Coordinate index of one sparse tensor: an occupancy bitmap over batch*D*H*W with a
 * per-word popcount prefix (and, when the rows are not in ascending flattened order, a
 * rank -> row permutation).  It is what spconv keeps as its grid/hash inside
 * indice_dict (spconv.SparseConvTensor.indice_dict, used at scn.py:105-152).
 *
 *   index memory = s2d_grid_index_bytes(batch, shape, n_rows_capacity), owned by the caller.
 *   s2d_grid_index_build: rows may be in any order and must be unique.
 * ------------------------------------------------------------------------------------- */
size_t s2d_grid_index_bytes(int batch, const int* shape_host, int n_rows_capacity);
/* n_rows_dev (nullable): when given, the row count is read from device memory at run time and
 * n_rows is only the launch bound / capacity -- lets the voxelizer feed the index without a
 * host round trip. */
int s2d_grid_index_build(const int* coors, int n_rows, const int* n_rows_dev, int batch,
                         const int* shape_host, void* index, size_t index_bytes, void* stream);

/* SubMConv3d rulebook (spconv.SubMConv3d, scn.py:16-39,105): output set == input set, same
 * row order.  tbl i32 [K, tbl_stride], K = prod(ksize); neighbour of row i under offset k is
 * the active voxel at x_i + (k - ksize/2)*dilation.  n_pairs (nullable): device u64, += the
 * number of (in,out) pairs. */
int s2d_rulebook_subm(const int* coors, int n_rows, int batch, const int* shape_host,
                      const int* ksize_host, const int* dilation_host, const void* index, int* tbl,
                      int tbl_stride, unsigned long long* n_pairs, void* stream);

/* SparseConv3d (spconv.SparseConv3d, scn.py:116-118,126-128,136-138,147-149), in two steps so
 * that the coordinate phase of a whole backbone can run without a host synchronisation:
 *
 * 1. s2d_sparse_out_coords: output site o is active iff x = o*s - p + k*d for an active input
 *    x and some offset k.  Writes out_coors i32 [out_capacity,4] in ascending flattened
 *    (b,z,y,x) order, the true count to *n_out (device; may exceed out_capacity, the caller
 *    checks) and builds index_out = the coordinate index of the output tensor
 *    (s2d_grid_index_bytes(batch, shape_out, out_capacity)), which serves the next layers.
 *    n_in_dev: as n_rows_dev above.
 * 2. s2d_rulebook_sparse: tbl[k][o] = input row at o*s - p + k*d or -1, looked up in index_in
 *    (the index of the INPUT tensor). */
int s2d_conv_out_shape(const int* shape_in_host, const int* ksize_host, const int* stride_host,
                       const int* pad_host, const int* dilation_host, int* shape_out_host);
int s2d_sparse_out_coords(const int* coors_in, int n_in, const int* n_in_dev, int batch,
                          const int* shape_in_host, const int* ksize_host, const int* stride_host,
                          const int* pad_host, const int* dilation_host, void* index_out,
                          size_t index_out_bytes, int* out_coors, int out_capacity, int* n_out,
                          void* stream);
int s2d_rulebook_sparse(const int* out_coors, int n_out, int batch, const int* shape_in_host,
                        const int* ksize_host, const int* stride_host, const int* pad_host,
                        const int* dilation_host, const void* index_in, int* tbl, int tbl_stride,
                        unsigned long long* n_pairs, void* stream);

/* ---------------------------------------------------------------------------------------
 * Sparse convolution forward, output-stationary gather-GEMM with fused epilogue.  Replaces
 * spconv's per-offset gather -> mm -> scatter-add and the BatchNorm1d / ReLU / residual
 * modules applied to .features (scn.py:69-85,104-152):
 *
 *   out[o,:] = act( (sum_k in[tbl[k][o],:] @ W[k]) * scale + shift (+ residual[o,:]) )
 *
 *   W f32 [K, Cin, Cout] (spconv layout [kD,kH,kW,Cin,Cout] flattened)
 *   scale, shift (nullable, both or neither) f32 [Cout] -- eval-mode BN folded by the host,
 *   or bias via shift with scale == NULL meaning 1
 *   residual (nullable) f32 [n_out, Cout]; relu != 0 applies max(0, .)
 *   precision: S2D_PRECISION_FP32 (SIMT FFMA) or a tcgen05 mode (fp32 accumulate in TMEM).
 *              For the tcgen05 modes W must be the PACKED image produced by s2d_spconv_pack_weights
 *              FOR THAT PRECISION (once per layer), not the raw tensor:
 *                TF32X3     = error-compensated split (hi*hi + lo*hi + hi*lo in TF32, ~fp32 accuracy),
 *                TF32_BF16C = hi*hi in TF32 + the two correction terms as one BF16 contraction
 *                             (same accuracy class, 2/3 of the tensor time; TF32 and TF32X3 share an image),
 *                TF32       = single pass.
 *              Shapes: Cin == 16 or Cin % 32 == 0; Cout % 16 == 0 (s2d_spconv_tf32_supported).
 * ------------------------------------------------------------------------------------- */
#define S2D_PRECISION_FP32 0
#define S2D_PRECISION_TF32 1
#define S2D_PRECISION_TF32X3 2
#define S2D_PRECISION_TF32_BF16C 3
#define S2D_PRECISION_AUTO 4 /* TF32_BF16C where the layer is tensor bound (Cout % 128 == 0), TF32X3 elsewhere */
#define S2D_PRECISION_BF16X2 5 /* operands pre-split into BF16 pairs (x = hi + lo, 16 mantissa bits): hi*w1 + hi*w2 + lo*w1,
                                  three BF16 MMAs with fp32 accumulation, ~4e-6 relative per layer.  Needs in_split. */
int s2d_spconv_tf32_supported(int Cin, int Cout);
size_t s2d_spconv_packed_bytes(int K, int Cin, int Cout);
int s2d_spconv_pack_weights(const float* W, int K, int Cin, int Cout, int precision, float* packed, void* stream);
int s2d_spconv_fwd(const float* in, int n_in, const float* W, const int* tbl, int tbl_stride,
                   int n_out, int Cin, int Cout, int K, const float* scale, const float* shift,
                   const float* residual, int relu, float* out, int precision, void* stream);

/* ---------------------------------------------------------------------------------------
 * General gather-GEMM convolution (the same kernels as s2d_spconv_fwd with every knob exposed).
 * It also serves the DENSE 2-D convolutions of the S2D neck / RPN / CenterHead
 * (det3d/models/necks/rpn.py:186-259,300-337; det3d/models/bbox_heads/center_head.py:209-244):
 * a BEV map in NHWC is a "sparse" tensor with every site active, rows = (b,y,x), and a
 * Conv2d / ConvTranspose2d is a gather-GEMM over a regular-grid neighbour table
 * (s2d_grid2d_table).  Replaces torch.nn.Conv2d/ConvTranspose2d + BatchNorm2d + GELU/ReLU
 * (+ residual add, + torch.cat through out_ld / channel-offset pointers).
 *
 *   out[r(o), :] = post( act( (sum_k in[tbl[k][o], :] @ W[k]) * scale + shift  [+ res] ) [+ res] )
 *
 *   in_ld / out_ld / res_ld : row strides in floats (multiples of 4); pointers may already be
 *                             offset to a channel slice of a wider buffer
 *   out_rows (nullable)     : r(o) = out_rows[o] (row remap, e.g. sub-pixel transposed conv)
 *   act                     : S2D_ACT_NONE / RELU / GELU (exact erf form, torch.nn.GELU default)
 *   res_after_act           : 0: residual added before the activation (ResNet), 1: after it
 *   weights                 : raw [K,Cin,Cout] for S2D_PRECISION_FP32, packed image otherwise
 * ------------------------------------------------------------------------------------- */
#define S2D_ACT_NONE 0
#define S2D_ACT_RELU 1
#define S2D_ACT_GELU 2
typedef struct s2d_conv_params {
  const float* in;
  const float* weights;
  const int* tbl;
  const float* scale;
  const float* shift;
  const float* residual;
  float* out;
  const int* out_rows;
  int in_ld, out_ld, res_ld;
  int tbl_stride, K;
  int n_in, n_out, Cin, Cout;
  int act, res_after_act, precision;
  /* S2D_PRECISION_BF16X2 only (zero / null otherwise).  A "split row" stores, per 32-channel chunk (or per 16-channel
   * row), [hi words | lo words], two BF16 per 32-bit word, at word offset = channel offset: the same geometry and byte
   * size as the fp32 row.  in_split: the input rows in that form (s2d_rows_split, or the out_split of the producing
   * layer; `in` may then be null).  out_split: optional second output for the next layer (`out` may be null).
   * tile_masks: optional i32 [ceil(n_out / 128)], bit k set iff some row of the 128-row tile has tbl[k][row] >= 0
   * (s2d_table_tile_masks); offsets whose bit is clear are skipped for the whole tile. */
  const void* in_split;
  void* out_split;
  const int* tile_masks;
  int in_split_ld, out_split_ld;
} s2d_conv_params;
int s2d_conv_fwd(const s2d_conv_params* params, void* stream);
/* Dense-grid form of s2d_conv_fwd for a regular map [B, H, W] of split rows (stride 1, k x k kernel, padding `pad`; the
 * Conv2d layers of det3d/models/necks/rpn.py and bbox_heads/center_head.py): no neighbour table -- the 128 rows a 16 x 8-pixel
 * tile needs for a kernel offset are ONE 4-D TMA box (cp.async.bulk.tensor, zero fill outside the map = the padding).
 * params: precision = S2D_PRECISION_BF16X2, in_split / in_split_ld = the map, n_in = B*H*W, tbl = NULL, K = k*k,
 * n_out = s2d_grid2d_tile_rows_count(B, H, W) (tile-order rows), out_rows = s2d_grid2d_tile_rows(...) (tile-order row ->
 * pixel row, -1 outside the map); everything else as for s2d_conv_fwd. */
int s2d_grid2d_tile_rows_count(int B, int H, int W);
int s2d_grid2d_tile_rows(int B, int H, int W, int* rows, void* stream);
int s2d_conv_fwd_grid(const s2d_conv_params* params, int B, int H, int W, int k, int pad, void* stream);
/* fp32 rows [n_rows, C] (row stride in_ld floats) -> split rows (row stride out_ld words); C == 16 or C % 32 == 0. */
int s2d_rows_split(const float* in, long long n_rows, int C, int in_ld, void* out, int out_ld, void* stream);
/* tile_masks[t] = OR over the rows r of tile t (128 rows) and the offsets k < K (<= 31) of (tbl[k][r] >= 0) << k. */
int s2d_table_tile_masks(const int* tbl, int tbl_stride, int K, int n_rows, int* tile_masks, void* stream);
/* Row grouping for the tile kernel (conv_bf2.cu): a stable counting sort of the table's rows by the 9-bit key "which
 * (dz, dy) offset triples k/3 have a neighbour" -> perm[p] = original row of position p, tbl_out[k][p] = tbl[k][perm[p]] and
 * the tile masks of tbl_out.  A conv launched with tbl_out, out_rows = perm and these masks writes bit-identical results
 * (each output row still sums its offsets in the same order) while skipping the (tile, offset) pairs grouping has emptied.
 * Replaces nothing in the reference (spconv executes per-offset pair lists, spconv/ops.py indice_conv); K <= 27. */
size_t s2d_rulebook_subm_grouped_workspace_bytes(int n_rows);
/* s2d_rulebook_subm (3 x 3 x 3) built directly in grouped row order: perm, tbl[k][p] = neighbour k of row perm[p], tile masks;
 * the key comes from the occupancy bitmap, so the scan-order table is never written (same result as s2d_rulebook_subm +
 * s2d_table_group_rows). */
int s2d_rulebook_subm_grouped(const int* coors, int n_rows, int batch, const int* shape_host, const int* dilation_host,
                              const void* index, int* perm, int* tbl, int tbl_stride, int* tile_masks, void* workspace,
                              size_t workspace_bytes, void* stream);
/* The strided 3 x 3 x 3 rulebook (s2d_rulebook_sparse) built directly in grouped row order; workspace as above. */
int s2d_rulebook_sparse_grouped(const int* out_coors, int n_out, int batch, const int* shape_in_host, const int* stride_host,
                                const int* pad_host, const int* dilation_host, const void* index_in, int* perm, int* tbl,
                                int tbl_stride, int* tile_masks, void* workspace, size_t workspace_bytes, void* stream);
size_t s2d_table_group_rows_workspace_bytes(int n_rows);
int s2d_table_group_rows(const int* tbl, int tbl_stride, int K, int n_rows, int* perm, int* tbl_out, int out_stride,
                         int* tile_masks, void* workspace, size_t workspace_bytes, void* stream);

/* Regular-grid neighbour tables for dense 2-D convolutions over NHWC rows (row = (b*H + y)*W + x).
 *   s2d_grid2d_table       : Conv2d(kh x kw, stride, pad): tbl i32 [kh*kw, B*Ho*Wo], -1 outside the map
 *                            (zero padding; nn.ZeroPad2d(1)+Conv2d(3) of rpn.py:128-131 is pad = 1)
 *   s2d_grid2d_tconv_table : ConvTranspose2d(stride 2) with (k,pad) = (4,1) or (2,0) as four sub-pixel
 *                            convolutions (rpn.py:224-238,84-92): class (py,px) owns outputs (2y+py, 2x+px);
 *                            tbl i32 [(kh/2)*(kw/2), B*H*W] over the INPUT grid, taps ky = ((py+pad)&1) + 2a,
 *                            out_rows i32 [B*H*W] = output row of each input-grid site for s2d_conv_fwd. */
int s2d_grid2d_table(int B, int H, int W, int kh, int kw, int stride, int pad, int* tbl, int tbl_stride,
                     void* stream);
int s2d_grid2d_tconv_table(int B, int H, int W, int kh, int kw, int pad, int py, int px, int* tbl,
                           int tbl_stride, int* out_rows, void* stream);

/* Layout changes at the module boundary: torch NCHW [B,C,HW] <-> rows [B*HW, ld] (channel fastest). */
int s2d_nchw_to_nhwc(const float* in, int B, int C, int HW, float* out, int out_ld, void* stream);
int s2d_nhwc_to_nchw(const float* in, int in_ld, int B, int C, int HW, float* out, void* stream);

/* ConvNeXt pieces of the S2D module (rpn.py:204-222): depthwise k x k conv (weight [C,k,k], bias [C]
 * nullable) and nn.LayerNorm([C,H,W], eps) over all C*H*W elements of a sample with affine
 * parameters stored [C,H,W]; data in NHWC rows. */
int s2d_dwconv2d(const float* in, const float* weight, const float* bias, int B, int H, int W, int C, int k,
                 int pad, float* out, void* stream);
size_t s2d_layernorm_workspace_bytes(int B);
int s2d_layernorm_chw(const float* in, const float* gamma, const float* beta, int B, int C, int HW, float eps,
                      float* out, void* workspace, size_t workspace_bytes, void* stream);

/* SparseConvTensor.dense() + view(N, C*D, H, W) (scn.py:173-176) written directly as NHWC rows
 * out f32 [B*H*W, out_ld], out[(b,y,x)][c*D + z] = feat[row, c]; zero-fills out itself. */
int s2d_dense_bev_nhwc(const float* feat, const int* coors, int n_rows, int C, int batch, int D, int H, int W,
                       float* out, int out_ld, void* stream);

/* SparseConvTensor.dense() + view(N, C*D, H, W) (scn.py:173-176): bev f32 [batch, C*D, H, W],
 * bev[b, c*D + z, y, x] = feat[row, c].  The kernel zero-fills bev itself. */
int s2d_dense_bev(const float* feat, const int* coors, int n_rows, int C, int batch, int D, int H, int W,
                  float* bev, void* stream);
/* Same result, output-stationary: a cell -> row map in `workspace` (s2d_dense_bev_workspace_bytes), then every global
 * access is a full 128 B line and empty cells are zero-filled in the same pass (no memset of the whole map). */
size_t s2d_dense_bev_workspace_bytes(int batch, int D, int H, int W);
int s2d_dense_bev_tiled(const float* feat, const int* coors, int n_rows, int C, int batch, int D, int H, int W,
                        float* bev, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * CenterHead.predict on the device (det3d/models/bbox_heads/center_head.py:293-495, one task, no
 * double flip): decode -> score / range masks -> top nms_pre_max_size by score -> rotated-BEV NMS
 * -> first nms_post_max_size survivors.  Replaces the eager torch ops of predict/post_processing,
 * rotate_nms_pcdet (det3d/core/bbox/box_torch_ops.py:449-470) and iou3d_nms_cuda.nms_gpu
 * (det3d/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17, iou3d_nms.cpp:90-136, kernel
 * iou3d_nms_kernel.cu:267-311), which copies the suppression mask to the host and sweeps it on
 * the CPU; nothing here leaves the device or synchronises.
 *
 * s2d_centerhead_decode: head maps as NHWC rows (row = (b*H + y)*W + x, row strides ld_* in floats,
 *   so the channel slices of one fused head-output buffer can be passed directly):
 *     score = max_c sigmoid(hm[c]) (first maximum), label = argmax, box = [x, y, height, exp(dim[0..2]),
 *     atan2(rot[0], rot[1])] with x = (col + reg[0]) * out_size_factor * voxel_x + pc_x (fp32 steps as in
 *     center_head.py:393-401); key = 0 unless score > score_threshold and the centre lies in `range`
 *     (post_center_limit_range), else (score bits << 32) | (0xFFFFFFFF - cell): descending key order =
 *     descending score, ties by ascending cell.
 *   boxes f32 [B*H*W,7], scores f32 [B*H*W], labels i32 [B*H*W], keys u64 [B*H*W].
 * s2d_centerhead_select: per sample the first `pre_max` (<= 4096) keys in descending order, NMS at
 *   iou_threshold, at most post_max survivors gathered to out_* [batch, post_max, ...] (rows past
 *   n_out[b] are zero / -1); out_cells (nullable) = BEV cell of each detection.
 * s2d_nms_sorted: drop-in for nms_gpu(boxes sorted by descending score, keep, thresh): keep i32 [n],
 *   *n_keep on the device; n <= 4096 (rotate_nms_pcdet never passes more than nms_pre_max_size).
 * s2d_iou_bev: pairwise rotated-BEV IoU [n_a, n_b] (boxes_iou_bev_gpu of the same extension).
 * ------------------------------------------------------------------------------------- */
typedef struct s2d_decode_params {
  const float* reg;
  const float* height;
  const float* dim;
  const float* rot;
  const float* hm;
  int ld_reg, ld_height, ld_dim, ld_rot, ld_hm;
  int B, H, W, num_cls;
  float out_size_factor, voxel_x, voxel_y, pc_x, pc_y;
  float score_threshold;
  float range[6];
} s2d_decode_params;
int s2d_centerhead_decode(const s2d_decode_params* params, float* boxes, float* scores, int* labels,
                          unsigned long long* keys, void* stream);
size_t s2d_centerhead_select_workspace_bytes(int batch, int pre_max, int post_max);
int s2d_centerhead_select(const unsigned long long* keys, const float* boxes, const float* scores,
                          const int* labels, int batch, int cells, int pre_max, float iou_threshold,
                          int post_max, float* out_boxes, float* out_scores, int* out_labels, int* out_cells,
                          int* n_out, void* workspace, size_t workspace_bytes, void* stream);
size_t s2d_nms_workspace_bytes(int n_boxes);
int s2d_nms_sorted(const float* boxes, int n_boxes, float iou_threshold, int* keep, int* n_keep,
                   void* workspace, size_t workspace_bytes, void* stream);
int s2d_iou_bev(const float* boxes_a, int n_a, const float* boxes_b, int n_b, float* ious, void* stream);

/* ---------------------------------------------------------------------------------------
 * Second stage of the two-stage CenterPoint (det3d/models/detectors/two_stage.py:154-199).
 *
 * s2d_bev_box_features: for every first-stage box its centre and (num_point == 5) the four face
 *   mid-points (two_stage.py:49-76, box_torch_ops.py:386-406) are sampled bilinearly from the BEV map
 *   (bird_eye_view.py:24-40, center_utils.py:93-122 incl. the clamp-before-weights behaviour) and
 *   concatenated per box: out f32 [batch*max_boxes, num_point*C] in the order centre, front, back,
 *   left, right; slots p >= n_boxes[b] are zero (reorder_first_stage_pred_and_feature, :78-119).
 *   bev: NHWC rows [batch*H*W, bev_ld] (no NCHW->NHWC copy of the map is needed); boxes f32
 *   [batch, max_boxes, 7]; n_boxes device i32 [batch].
 * s2d_roi_refine: generate_predicted_boxes (roi_head_template.py:153-183) + post_process
 *   (two_stage.py:121-151): out_boxes = rotate_z(rcnn_reg + roi(xyz = 0), roi_yaw) + roi_xyz,
 *   out_scores = sqrt(sigmoid(rcnn_cls) * roi_score); padded slots zero.
 * ------------------------------------------------------------------------------------- */
int s2d_bev_box_features(const float* bev, int bev_ld, int batch, int H, int W, int C, const float* boxes,
                         const int* n_boxes, int max_boxes, int num_point, const float* pc_start_host,
                         const float* voxel_size_host, float out_stride, float* out, void* stream);
int s2d_roi_refine(const float* rois, const float* roi_scores, const int* n_boxes, int batch, int max_boxes,
                   const float* rcnn_cls, int cls_ld, const float* rcnn_reg, int reg_ld, float* out_boxes,
                   float* out_scores, void* stream);

/* ---------------------------------------------------------------------------------------
 * PointPillars + S2D variant (BASELINE configs[3]).
 *
 * s2d_pfn_fwd: PillarFeatureNet.forward with num_filters = [64, 64]
 *   (det3d/models/readers/pillar_encoder.py:41-56,114-154) as ONE kernel: decorate each point with the offset
 *   to the pillar's point mean and to the pillar centre (x_offset = vx/2 + pc_range_x), zero the padded
 *   points, Linear(10->32, no bias) + BN1d (folded scale/shift) + ReLU, max over points, concat [x, max],
 *   Linear(64->64) + BN1d + ReLU, max -> out f32 [n_pillars, 64].
 *   voxels f32 [n_pillars, max_points, 5] zero padded, num_points i32, coors i32 [n_pillars,4] (b,z,y,x),
 *   W0 f32 [32,10], W1 f32 [64,64] (nn.Linear layout [out,in]).
 * s2d_maxpool2d_rows: nn.MaxPool2d(2,2) on NHWC rows (pillar_encoder.py:236).
 * s2d_gather_rows: out[o,:] = in[idx[o],:] (zero when idx < 0): nn.Upsample(nearest) with a host-made index map
 *   (pillar_encoder.py:283,295) and any other row remap.
 * s2d_grid2d_tconv_table_s: sub-pixel class (py,px) of ConvTranspose2d(k, stride, pad) with k % stride == 0 and
 *   k - stride == 2*pad (RPN deblocks with stride 2 and 4, det3d/models/necks/rpn.py:82-95); tbl i32
 *   [(k/stride)^2, B*H*W], out_rows as in s2d_grid2d_tconv_table.
 * The pillar scatter (PointPillarsScatter*.forward, pillar_encoder.py:337-371) is s2d_dense_bev_nhwc with D = 1.
 * ------------------------------------------------------------------------------------- */
int s2d_pfn_fwd(const float* voxels, const int* num_points, const int* coors, int n_pillars, int max_points, int F,
                float vx, float vy, float x_offset, float y_offset, const float* W0, const float* scale0,
                const float* shift0, const float* W1, const float* scale1, const float* shift1, float* out,
                void* stream);
int s2d_maxpool2d_rows(const float* in, int in_ld, int B, int H, int W, int C, float* out, int out_ld, void* stream);
int s2d_gather_rows(const float* in, int in_ld, const int* idx, long long n_out, int C, float* out, int out_ld,
                    void* stream);
int s2d_grid2d_tconv_table_s(int B, int H, int W, int k, int stride, int pad, int py, int px, int* tbl,
                             int tbl_stride, int* out_rows, void* stream);

/* ---------------------------------------------------------------------------------------
 * Loss values of the distillation training step (forward only).  Every function is a
 * single pass over its inputs with a deterministic two-level reduction in double; results are device doubles.
 * A map element (b, c, cell) lives at base + b*sb + c*sc + cell*scell floats (NCHW: sc = H*W, scell = 1; NHWC rows:
 * sc = 1, scell = row stride), so the head outputs can be read where they are.
 *
 * s2d_masked_mse: out4 = { sum_{t>0} (s-t)^2, #{t>0}, sum_{t<=0} (s-t)^2, #{t<=0} } over n elements -- the
 *   sparse2dense_loss terms F.mse_loss(F_S[F_D>0], F_D[F_D>0]) and F.mse_loss(F_S[~], F_D[~])
 *   (det3d/torchie/trainer/trainer.py:783-789) are out4[0]/out4[1] and out4[2]/out4[3].
 * s2d_focal_loss: FastFocalLoss.forward / fastfocalloss (det3d/models/losses/centernet_loss.py:27-54, trainer.py:38-58):
 *   out3 = { sum log(1-o) o^2 (1-t)^4 over the map, sum log(p)(1-p)^2 mask over the M peaks, sum mask };
 *   loss = -(out3[0] + out3[1]) / out3[2], or -out3[0] when out3[2] == 0.  out_is_logits applies
 *   CenterHead._sigmoid (clamped sigmoid, center_head.py:246-248), target_is_logits applies sigmoid (trainer.py:792).
 *   ind i64 [B,M] (cell), mask u8 [B,M], cat i64 [B,M].
 * s2d_gather_reg_loss: RegLoss.forward (centernet_loss.py:6-25; squared = 0) and distill_reg_loss (trainer.py:68-76;
 *   squared = 1, target_map = the teacher's anno_box map): out[d] = sum_{b,m} err(pred*mask, target*mask) for
 *   d < D <= 16 and out[16] = sum(mask); loss[d] = out[d] / (out[16] + 1e-4).  target_rows f32 [B,M,D] or target_map.
 * ------------------------------------------------------------------------------------- */
size_t s2d_loss_workspace_bytes(void);
int s2d_masked_mse(const float* f_student, const float* f_teacher, long long n, double* out4, void* workspace,
                   size_t workspace_bytes, void* stream);
int s2d_focal_loss(const float* out, long long out_sb, long long out_sc, long long out_scell, int out_is_logits,
                   const float* target, long long tgt_sb, long long tgt_sc, long long tgt_scell, int target_is_logits,
                   int B, int C, int HW, const long long* ind, const unsigned char* mask, const long long* cat, int M,
                   double* out3, void* workspace, size_t workspace_bytes, void* stream);
int s2d_gather_reg_loss(const float* pred, long long pred_sb, long long pred_sc, long long pred_scell,
                        const float* target_rows, const float* target_map, long long tgt_sb, long long tgt_sc,
                        long long tgt_scell, int B, int M, int D, int squared, const long long* ind,
                        const unsigned char* mask, double* out, void* workspace, size_t workspace_bytes, void* stream);

/* Gradients of the three losses w.r.t. the student's map.  out4 / out3 / sums: the device result of the forward call;
 * upstream: optional device float (per regression dimension: float[D]) multiplied in; d_*: gradient buffer addressed with
 * its own (sb, sc, scell) strides.  s2d_gather_reg_loss_bwd zero-fills d_pred[0 .. d_numel) first and then adds the <= B*M
 * peak contributions; s2d_focal_loss_bwd writes every element.  With squared = 0 the derivative of |e| at 0 is 0. */
int s2d_masked_mse_bwd(const float* f_student, const float* f_teacher, long long n, const double* out4, float w_pos,
                       float w_neg, const float* upstream, float* d_student, void* stream);
int s2d_focal_loss_bwd(const float* out, long long out_sb, long long out_sc, long long out_scell, int out_is_logits,
                       const float* target, long long tgt_sb, long long tgt_sc, long long tgt_scell, int target_is_logits,
                       int B, int C, int HW, const long long* ind, const unsigned char* mask, const long long* cat, int M,
                       const double* out3, const float* upstream, float* d_out, long long d_sb, long long d_sc,
                       long long d_scell, void* stream);
int s2d_gather_reg_loss_bwd(const float* pred, long long pred_sb, long long pred_sc, long long pred_scell,
                            const float* target_rows, const float* target_map, long long tgt_sb, long long tgt_sc,
                            long long tgt_scell, int B, int M, int D, int squared, const long long* ind,
                            const unsigned char* mask, const double* sums, const float* upstream, float* d_pred,
                            long long d_numel, long long d_sb, long long d_sc, long long d_scell, void* stream);

/* PCR losses, KD_VoxelNet.mask_offset_loss (det3d/models/detectors/voxelnet.py:171-185,229-249) without the dense gt / grid
 * tensors: predictions are rows in (b, y, x, z) order (row = ((b*H + y)*W + x)*D + z; mask_logits [n], offset [n,3]); the
 * reconstruction voxels are coors i32 [M,4] (b,z,y,x) + gt_feats f32 [M,5] (voxel means).  centre9 = {scale, minus, half} per
 * axis x, y, z: voxel centre = idx*scale - minus + half evaluated in fp32 in that order (the reference's grid expression).
 * sums6 = { sum softplus(x) over all cells, sum softplus(-x) over occupied, sum softplus(x) over occupied, #occupied,
 *           sum |offset - (gt_xyz - centre)| over the non-zero targets, their count };
 *   mask loss = (sums[0] - sums[2] + beta*sums[1]) / n with beta = (n - #occ) / #occ;  offset loss = sums[4] / sums[5].
 * s2d_pcr_loss_bwd: upstream2 = device {d mask loss, d offset loss}; writes d_mask_logits [n] and d_offset [n,3] (zero-filled). */
int s2d_pcr_loss(const float* mask_logits, const float* offset, int B, int D, int H, int W, const int* coors,
                 const float* gt_feats, long long M, const float* centre9, double* sums6, void* workspace,
                 size_t workspace_bytes, void* stream);
int s2d_pcr_loss_bwd(const float* mask_logits, const float* offset, int B, int D, int H, int W, const int* coors,
                     const float* gt_feats, long long M, const float* centre9, const double* sums6, const float* upstream2,
                     float* d_mask_logits, float* d_offset, void* stream);

/* ---------------------------------------------------------------------------------------
 * CenterPoint training targets for a batch: AssignLabel.__call__
 * (det3d/datasets/pipelines/preprocess.py:489-653; gaussian_radius / draw_umich_gaussian, det3d/core/utils/center_utils.py:18-64;
 * limit_period, det3d/core/bbox/box_np_ops.py:360-361) for ONE task.  gt_boxes f32 [B, max_objs, 9] (x y z w l h vx vy rot) in
 * the task's class-major object order, gt_classes i32 [B, max_objs] 1-based inside the task, num_objs i32 [B].
 * Outputs (zero-filled here): hm f32 [B, num_cls, H, W], anno_box f32 [B, max_objs, 10] (reg2, z, log wlh, vx, vy, sin, cos),
 * ind / cat i64 [B, max_objs], mask u8 [B, max_objs], optional gt_boxes_and_cls f32 [B, max_objs, 10] (x y z w l h rot vx vy,
 * class + cls_offset).  Objects with a non-positive size or a centre outside the map are skipped exactly as the reference does.
 * ------------------------------------------------------------------------------------- */
int s2d_assign_label(const float* gt_boxes, const int* gt_classes, const int* num_objs, int B, int max_objs, int num_cls,
                     int H, int W, float pc_x, float pc_y, float voxel_x, float voxel_y, int out_size_factor,
                     double gaussian_overlap, int min_radius, int cls_offset, float* hm, float* anno_box, long long* ind,
                     unsigned char* mask, long long* cat, float* gt_boxes_and_cls, void* stream);

/* ---------------------------------------------------------------------------------------
 * Training step (SURVEY.md section 8 rows a7 train-mode, a16 backward, a17): csrc/train.cu.
 *
 * Every convolution of this library is out[i] = sum_k in[tbl[k][i]] . W[k]; its backward is
 *   dIn  = s2d_conv_fwd over the TRANSPOSED table (s2d_table_transpose) with W[k]^T            (data gradient)
 *   dW[k][a][b] = sum_i G[tbl[k][i]][a] * D[d_rows ? d_rows[i] : i][b]    (s2d_conv_wgrad; G = layer input, D = dOut)
 * which replaces spconv's indice_conv_backward (per-offset gather -> mm -> scatter) and cuDNN's dgrad / wgrad.
 * s2d_table_transpose: inv[k][j] = (out_rows ? out_rows[i] : i) for every tbl[k][i] = j >= 0, -1 elsewhere.
 * s2d_conv_wgrad: deterministic (per-chunk partial sums, fixed-order reduction); accumulate != 0 adds to out.
 *
 * Rows normalisation [n, C] (BatchNorm1d over active voxels scn.py:100-107, BatchNorm2d over BEV pixels rpn.py:126-145,
 * det3d/models/utils/norm.py:59-108), training mode:
 *   s2d_bn_train_stats  -> mean, invstd (biased variance), folded scale = gamma*invstd, shift = beta - mean*scale, and the
 *                          running statistics update (momentum, unbiased variance) in place;
 *   s2d_rows_affine_act -> out = act(x*scale + shift (+ residual))  or  act(x*scale + shift) + residual;
 *   s2d_rows_affine_act_bwd -> dz = dy * act'(.) (also the residual gradient when the residual is added before the
 *                          activation), column sums {sum dz, sum dz*x} into the workspace, optional dshift = sum dz;
 *   s2d_bn_train_bwd    -> dx, dgamma, dbeta from the workspace the previous call filled.
 * s2d_layernorm_chw_bwd / s2d_dwconv2d_wgrad: backward of the ConvNeXt pieces (rpn.py:204-222); the depthwise data
 *   gradient is s2d_dwconv2d with the flipped kernel.  dweight of the depthwise conv is [C][k*k].
 * s2d_grad_norm_clip: out2 = { ||g||_2, min(1, max_norm / (norm + 1e-6)) } (clip_grad_norm_, hooks/optimizer.py:15-21).
 * s2d_adam_step: p *= 1 - weight_decay*lr, then Adam with bias correction (det3d/solver/fastai_optim.py:158-174 with
 *   true_wd, torch.optim.Adam arithmetic); grad_scale: optional device float multiplied into g (the clip coefficient).
 * ------------------------------------------------------------------------------------- */
int s2d_table_transpose(const int* tbl, int tbl_stride, int K, int n_out, const int* out_rows, int* inv, int inv_stride,
                        int n_in, void* stream);
size_t s2d_conv_wgrad_workspace_bytes(int n_rows, int K, int Cg, int Cd);
int s2d_conv_wgrad(const float* g, int g_ld, int n_g, int Cg, const float* d, int d_ld, const int* d_rows, int Cd,
                   const int* tbl, int tbl_stride, int n_rows, int K, float* out, int accumulate, void* workspace,
                   size_t workspace_bytes, void* stream);
/* The same weight gradient on the tensor cores (csrc/wgrad_tc.cu) for Cg, Cd multiples of 32: both operands are given as SPLIT
 * rows (s2d_rows_split: per 32-channel chunk [32 BF16 hi | 32 BF16 lo]), gathered straight into MN-major operand tiles; one
 * tcgen05.mma yields the hi/lo cross products as accumulator quadrants, summed by the epilogue (fp32-level result: both
 * operands carry 16 mantissa bits).  Deterministic (row chunks, fixed-order reduction). */
int s2d_conv_wgrad_bf2_supported(int Cg, int Cd);
size_t s2d_conv_wgrad_bf2_workspace_bytes(int n_rows, int K, int Cg, int Cd);
int s2d_conv_wgrad_bf2(const void* g_split, int g_ld, int n_g, int Cg, const void* d_split, int d_ld, const int* d_rows, int Cd,
                       const int* tbl, int tbl_stride, int n_rows, int K, float* out, int accumulate, void* workspace,
                       size_t workspace_bytes, void* stream);
size_t s2d_rows_workspace_bytes(int C);
int s2d_bn_train_stats(const float* x, int ld, int n, int C, float eps, float momentum, const float* gamma,
                       const float* beta, float* running_mean, float* running_var, float* mean, float* invstd,
                       float* scale, float* shift, void* workspace, size_t workspace_bytes, void* stream);
int s2d_rows_affine_act(const float* x, int ld, int n, int C, const float* scale, const float* shift,
                        const float* residual, int res_ld, int act, int res_after_act, float* out, int out_ld,
                        void* stream);
int s2d_rows_affine_act_bwd(const float* x, int ld, int n, int C, const float* scale, const float* shift,
                            const float* residual, int res_ld, int act, int res_after_act, const float* dy, int dy_ld,
                            float* dz, int dz_ld, float* dshift, void* workspace, size_t workspace_bytes, void* stream);
int s2d_bn_train_bwd(const float* x, int ld, int n, int C, const float* dz, int dz_ld, const float* mean,
                     const float* invstd, const float* gamma, float* dx, int dx_ld, float* dgamma, float* dbeta,
                     void* workspace, size_t workspace_bytes, void* stream);
/* SyncBatchNorm (tools/train.py:92-96 --sync_bn, apis/train.py:281-303): the same statistics / backward in two halves.
 * The local column sums (2*C doubles) sit at byte offset s2d_rows_workspace_sums_offset(C) of the workspace after
 * s2d_bn_train_sums / s2d_rows_affine_act_bwd; the caller all-reduces (sum) them together with its row count over the ranks
 * (one NCCL all-reduce of 2*C + 1 doubles) and hands the result back as device pointers, so no host synchronisation:
 * forward  s2d_bn_train_sums -> all-reduce -> s2d_bn_train_finalize(global sums, global rows);
 * backward s2d_rows_affine_act_bwd -> s2d_bn_train_bwd_params (dgamma / dbeta from the LOCAL sums) -> all-reduce ->
 *          s2d_bn_train_bwd_dx (global sums and rows).  n = 0 rows on a rank is allowed in the sums / dx calls. */
size_t s2d_rows_workspace_sums_offset(int C);
int s2d_bn_train_sums(const float* x, int ld, int n, int C, void* workspace, size_t workspace_bytes, void* stream);
int s2d_bn_train_finalize(const double* sums2c, const double* n_total_dev, int C, float eps, float momentum,
                          const float* gamma, const float* beta, float* running_mean, float* running_var, float* mean,
                          float* invstd, float* scale, float* shift, void* stream);
int s2d_bn_train_bwd_params(const double* sums2c_local, int C, const float* mean, const float* invstd, float* dgamma,
                            float* dbeta, void* stream);
int s2d_bn_train_bwd_dx(const float* x, int ld, int n, int C, const float* dz, int dz_ld, const float* mean,
                        const float* invstd, const float* gamma, const double* sums2c_global, const double* n_total_dev,
                        float* dx, int dx_ld, void* workspace, size_t workspace_bytes, void* stream);
size_t s2d_layernorm_bwd_workspace_bytes(int B);
int s2d_layernorm_chw_bwd(const float* x, const float* weight, int B, int C, int HW, float eps, const float* dy, float* dx,
                          float* dweight, float* dbias, void* workspace, size_t workspace_bytes, void* stream);
size_t s2d_dwconv2d_wgrad_workspace_bytes(int B, int H, int W, int C, int k);
int s2d_dwconv2d_wgrad(const float* x, const float* dy, int B, int H, int W, int C, int k, int pad, float* dweight,
                       void* workspace, size_t workspace_bytes, void* stream);
size_t s2d_grad_norm_workspace_bytes(void);
int s2d_grad_norm_clip(const float* g, long long n, float max_norm, float* out2, void* workspace, size_t workspace_bytes,
                       void* stream);
int s2d_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, const float* grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2D_B200_H_ */
