/* pvo_b200 -- C ABI of libpvo_b200.so (B200 / sm_100a CUDA implementation of PVO's VO_Module hot path).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point replaces one
 * function of the reference's pybind module `droid_backends` (VO_Module/src/droid.cpp:234-247) or one of the
 * library calls the reference's Python makes around it; the reference interface is cited at each declaration.
 * The Python module pvo_b200/droid_backends.py binds these with ctypes under the reference's own names, so
 * VO_Module/droid_slam runs on it unchanged.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers to contiguous row-major data; the library never allocates,
 *     frees or retains caller memory (scratch comes from a caller-provided workspace);
 *   - `stream` is a cudaStream_t (as void*); nothing synchronises the host, so every call can be captured in
 *     a CUDA graph;
 *   - return value: 0 ok, 1 invalid argument, 2 CUDA launch error, 3 unsupported, 4 workspace too small
 *     (the Python shim raises RuntimeError, mirroring TORCH_CHECK in droid.cpp:83-84);
 *   - dtype: 0 = float32, 1 = float16;
 *   - index tensors ii/jj are int64 (torch.long), exactly what the reference passes.
 */
#ifndef PVO_B200_H_
#define PVO_B200_H_
#ifdef __cplusplus
extern "C" {
#endif

#define PVO_DTYPE_F32 0
#define PVO_DTYPE_F16 1

int pvo_version(void);

/* ---- correlation ------------------------------------------------------------------------------------------- */

/* droid_backends.corr_index_forward(volume, coords, radius) -> [corr]
 *   src/droid.cpp:238, src/correlation_kernels.cu:19-70,126-155.   Bit-exact with the reference kernel.
 *   volume (N,h1,w1,h2,w2), coords (N,2,h1,w1) float32, out (N,2r+1,2r+1,h1,w1) (written, need not be zeroed).
 *   flags bit0: use 2-byte tap loads instead of the vectorised row fetch (for A/B measurement only). */
int pvo_corr_index_forward(const void* volume, const float* coords, void* out, int N, int h1, int w1, int h2, int w2,
                           int radius, int dtype, int flags, void* stream);

/* droid_backends.corr_index_backward(volume, coords, grad, radius) -> [grad_volume]
 *   src/droid.cpp:239, src/correlation_kernels.cu:73-124,157-185.  grad_volume must be zero-initialised. */
int pvo_corr_index_backward(const float* coords, const void* corr_grad, void* volume_grad, int N, int h1, int w1,
                            int h2, int w2, int radius, int dtype, void* stream);

/* CorrBlock.__call__ fused over all pyramid levels (droid_slam/modules/corr.py:40-50): one launch instead of
 *   4 x (CorrSampler + torch.zeros) + torch.cat.
 *   levels[l]: (E,h,w,level_h[l],level_w[l]); coords (E,h,w,2) float32 (NOT permuted; level l uses coords/2^l);
 *   layout 0: out (E, L*49, h, w) -- the reference's layout;
 *   layout 1: out (E*h*w, out_cstride) pixel-major, channel = l*49 + xoff*7 + yoff, pad channels zero-filled. */
int pvo_corr_lookup_pyramid(const void* const* levels, const int* level_h, const int* level_w, int num_levels,
                            const float* coords, void* out, int E, int h, int w, int radius, int dtype, int layout,
                            int out_cstride, int flags, void* stream);

/* CorrBlock.__init__: all-pairs correlation + 3x avg_pool2d pyramid (droid_slam/modules/corr.py:23-38,63-71).
 *   fmap1/fmap2: (E, h*w, 128) pixel-major float16 (the reference's (1,E,128,h,w) transposed), already the
 *   per-edge gather fmaps[ii] / fmaps[jj];  out level l: (E,h,w,h>>l,w>>l) float16.
 *   tcgen05 GEMM (fp32 accumulate in TMEM) with the 1/16 scale, the half rounding and all three pooling levels
 *   fused into the epilogue. */
/* same contraction on CUDA cores for any dtype (float32 volumes of the training path; fallback): fmaps (E,h*w,C) */
int pvo_corr_volume(const void* fmap1, const void* fmap2, void* const* levels_out, int E, int h, int w, int C,
                    int num_levels, int dtype, void* stream);
int pvo_corr_volume_f16(const void* fmap1, const void* fmap2, void* const* levels_out, int E, int h, int w,
                        int num_levels, void* stream);

/* The same volume + pyramid in the TILED private layout (what CorrBlock keeps on the production half path; the reference's
 *   `corr_pyramid` attribute, droid_slam/modules/corr.py:29-38, is materialised from it on demand by pvo_corr_untile):
 *   level l: (E, h*w, ceil((h>>l)/4), ceil((w>>l)/8), 4, 8) float16 -- each source pixel's (h>>l, w>>l) plane cut into tiles of
 *   4 rows x 8 columns (64 bytes = one DRAM access), tile-row-major, zero padded to whole tiles.  The 8 rows of a lookup window
 *   at any alignment lie in <= 2 tile columns: 16 (or 8) 16-byte pieces that share sectors, instead of 8 partially used rows. */
/*   fmap1 (E, h*w, 128); fmap2 (E, h*wpad, 128), image 2 zero padded to wpad in {40, 80, 104, 128} columns (the tensor-core tile
 *   width); level tensors are sized with wpad: (capacity, h*w, ceil((h>>l)/4), ceil((wpad>>l)/8), 4, 8).  slots (E int32, or NULL):
 *   edge e is written to plane block slots[e] -- FactorGraph keeps volumes in a pool, edges come and go without copies. */
int pvo_corr_volume_f16_tiled(const void* fmap1, const void* fmap2, void* const* levels_out, const int* slots, int E, int h, int w,
                              int wpad, int num_levels, void* stream);

/* CorrBlock.__call__ (droid_slam/modules/corr.py:40-50 -> src/correlation_kernels.cu:19-70) on the tiled layout: all four
 *   levels in one launch, radius 3, float16; bit-identical to pvo_corr_lookup_pyramid / the reference kernel on the same volume
 *   values.  levels / level_h / level_w: the tiled level tensors and the plane sizes (h>>l, wpad>>l); slots as above; coords (E,h,w,2) float32;
 *   layout / out / out_cstride as pvo_corr_lookup_pyramid.  The window rows are staged per pixel into shared memory with
 *   zero-filling cp.async (out-of-plane pieces are not fetched), one code path for interior, border and outside windows. */
int pvo_corr_lookup_tiled(const void* const* levels, const int* level_h, const int* level_w, int num_levels,
                          const int* slots, const float* coords, void* out, int E, int h, int w, int radius, int layout,
                          int out_cstride, void* stream);

/* layout conversion of ONE level between (nplanes, h2, w2) and (nplanes, ceil(h2/4), ceil(w2/8), 4, 8), float16 */
int pvo_corr_untile(const void* tiled, void* planar, long nplanes, int h2, int w2, int w2pad, void* stream);
int pvo_corr_tile(const void* planar, void* tiled, long nplanes, int h2, int w2, int w2pad, void* stream);

/* droid_backends.altcorr_forward(fmap1, fmap2, coords, radius) -> [corr]
 *   src/droid.cpp:240, src/altcorr_kernel.cu:27-149,290-319.
 *   fmap1 (B,H1,W1,C), fmap2 (B,H2,W2,C) float32, coords (B,S,H1,W1,2), out (B,S,(2r+1)^2,H1,W1). */
int pvo_altcorr_forward(const float* fmap1, const float* fmap2, const float* coords, float* out, int B, int S, int H1,
                        int W1, int H2, int W2, int C, int radius, void* stream);

/* droid_backends.altcorr_backward(fmap1, fmap2, coords, corr_grad, radius) -> [fmap1_grad, fmap2_grad, coords_grad]
 *   src/droid.cpp:241, src/altcorr_kernel.cu:152-286,321-356 (training only).  corr_grad (B,S,49,H1,W1); fmap1_grad (B,H1,W1,C) is
 *   written; fmap2_grad (B,H2,W2,C) must be zero-initialised (atomic accumulation); coords receive no gradient (the reference
 *   returns zeros). */
int pvo_altcorr_backward(const float* fmap1, const float* fmap2, const float* coords, const float* corr_grad,
                         float* fmap1_grad, float* fmap2_grad, int B, int S, int H1, int W1, int H2, int W2, int C, int radius,
                         void* stream);

/* cvx_upsample(data, mask) (droid_slam/droid_net.py:23-37; DepthVideo.upsample depth_video.py:139-144): convex 8x upsampling,
 *   softmax over the 9 logits of each of the 64 sub-pixels x 3x3 zero-padded neighbourhood.
 *   data (N,ht,wd,dim) float32; mask (N,576,ht,wd) float32|float16 (channel = k*64 + sy*8 + sx); out (N,8ht,8wd,dim) float32. */
int pvo_cvx_upsample(const float* data, const void* mask, float* out, int N, int ht, int wd, int dim, int mask_dtype, void* stream);
/* upsample_inter (droid_net.py:54-61, droid.py:106-113, factor_graph.py:431-438): bilinear x8, align_corners=True, channels last.
 *   in (N,ht,wd,dim) float32 -> out (N,8ht,8wd,dim) float32 */
int pvo_upsample_inter(const float* in, float* out, int N, int ht, int wd, int dim, void* stream);

/* ---- geometry ---------------------------------------------------------------------------------------------- */

/* DepthVideo.reproject(ii, jj) (droid_slam/depth_video.py:154-163 -> geom/projective_ops.py:102-130, lietorch
 *   SE3 mul/inv/act4) fused with the motion-feature construction of FactorGraph.update (factor_graph.py:231-237).
 *   poses (F,7), disps (F,ht,wd), intrinsics (F,4); coords out (E,ht,wd,2); valid out (E,ht,wd) or NULL;
 *   motn out (E*ht*wd, 8) in motn_dtype or NULL; when motn != NULL target/delta_dy/raw_mask are (E,ht,wd,2). */
int pvo_reproject(const float* poses, const float* disps, const float* intrinsics, const long* ii, const long* jj,
                  float* coords, float* valid, const float* target, const float* delta_dy, const float* raw_mask,
                  void* motn, int motn_dtype, int E, int ht, int wd, void* stream);

/* droid_backends.frame_distance(poses, disps, intrinsics, ii, jj, beta) -> dist
 *   src/droid.cpp:236, src/droid_kernels.cu:497-636.  intrinsics (4,). */
int pvo_frame_distance(const float* poses, const float* disps, const float* intrinsics, const long* ii,
                       const long* jj, float* dist, int n, int ht, int wd, float beta, void* stream);

/* ---- dense bundle adjustment ------------------------------------------------------------------------------- */

/* droid_backends.ba(poses, disps, intrinsics, targets, weights, eta, ii, jj, t0, t1, iterations, lm, ep,
 *                   motion_only) -> [dx, dz]       src/droid.cpp:87-114,235, src/droid_kernels.cu:1293-1410.
 *   In place on poses (nframes,7) / disps (nframes,ht,wd).  intrinsics (4,).  targets/weights (E,2,ht,wd)
 *   [interleaved = 0, the reference layout] or (E,ht,wd,2) [interleaved = 1].  eta (K,ht,wd),
 *   K = |unique(arange(t0,t1) U ii)|.  dx_out (t1-t0,6) / dz_out (K,ht*wd) may be NULL.
 *   workspace >= pvo_ba_workspace_bytes(...).  flags bit0: reuse the edge plan already in the workspace;
 *   bits 1-2: phase (multi-GPU, below); bit 3: geom/ba.py semantics; bits 8-19: max out-degree hint; bits 20-27: 1 + block
 *   bandwidth of the pose system (<= 2 max|ii-jj| over the whole graph; 0 = unknown, the solver scans for it). */
long pvo_ba_workspace_bytes(int E, int nframes, int P, int K, int HW);
int pvo_ba(float* poses, float* disps, const float* intrinsics, const float* targets, const float* weights,
           const float* eta, const long* ii, const long* jj, int E, int nframes, int ht, int wd, int t0, int t1, int K,
           int iterations, float lm, float ep, int motion_only, int interleaved, float* dx_out, float* dz_out,
           void* workspace, long workspace_bytes, int flags, void* stream);
/* byte range of {H, b} (float64) inside the workspace -- what pvo_b200/dist.py all-reduces between the two phases
 *   (flags bits 1-2 of pvo_ba: 1 = assemble only, 2 = solve + update only) when edges are sharded over GPUs */
int pvo_ba_system_range(int E, int nframes, int P, int K, int HW, long* offset, long* bytes);
/* the range the ranks exchange for a pvo_ba call with these `flags` / `motion_only`: when the flags carry the block bandwidth of the
 *   pose system (bits 20-27) and the system is too large for the dense shared-memory solve, phase 1 leaves a BAND-PACKED block
 *   {[6P][band + 6] lower-band rows | b | dA} (P = 63, radius 3: 133 KB instead of 1.1 MB) and phase 2 solves from it; otherwise the
 *   dense range of pvo_ba_system_range.  The reference has no multi-GPU path; the system is droid_kernels.cu:1096-1198's SparseBlock. */
int pvo_ba_exchange_range(int E, int nframes, int P, int K, int HW, int flags, int motion_only, long* offset, long* bytes);
/* copies {K seen on the device, error flag, max out-degree} to the host (synchronises; debugging only) */
int pvo_ba_read_meta(const void* workspace, int E, int nframes, int P, int K, int HW, int* meta_host3);
/* DepthVideo.ba's trailing disps.clamp_(min=0.001), depth_video.py:214 */
int pvo_clamp_min(float* x, long n, float lo, void* stream);

/* ---- update operator building blocks ----------------------------------------------------------------------- */

/* nn.Conv2d (+ fused torch.cat of inputs, bias, activation, ConvGRU gate arithmetic) on CUDA cores, fp32
 *   accumulate; droid_slam/droid_net.py:167-225, modules/gru.py:6-32.  See pvo_b200/csrc/conv_simt.cu. */
int pvo_conv2d(const void* const* seg_ptr, const int* seg_channels, const int* seg_stride, int nseg, const float* W,
               int wstride, const float* bias, int Cout, int KH, int KW, int pad, int nimg, int H, int Wd, int dtype,
               int epilogue, void* out, int ostride, int ooff, void* out2, const void* aux0, int aux_stride,
               const float* aux1, int stride, void* stream);
/* nn.InstanceNorm2d (affine=False) + ReLU [+ residual add + ReLU] of BasicEncoder's fnet
 *   (droid_slam/modules/extractor.py:28-32,47-55,129-130) on pixel-major (N,HW,C) data.  scratch: N*C*66 floats. */
int pvo_instance_norm(const void* in, void* out, const void* res, float* scratch, int N, int HW, int C, int relu, int dtype,
                      void* stream);
/* the same contraction on tcgen05 tensor cores (float16 operands via TMA, fp32 accumulate in TMEM), 3x3 / 1x1 /
 *   7x1 kernels with Cin % 16 == 0 and Cout % 64 == 0.  See pvo_b200/csrc/conv_tc.cu. */
int pvo_conv2d_tc(const void* const* seg_ptr, const int* seg_channels, const int* seg_stride, int nseg, const void* Wh,
                  const float* bias, int Cout, int KH, int KW, int nimg, int H, int Wd, int epilogue, void* out,
                  int ostride, int ooff, void* out2, const void* aux0, int aux_stride, const float* aux1, void* stream);
/* grouped 3x3 convolution with <= 2 outputs per group: the four 128->2 heads and GraphAgg's eta head */
int pvo_conv3x3_small(const void* in, int istride, int cin_g, const float* W, const float* bias, float* out,
                      int ostride, int G, int CG, int nimg, int H, int Wd, int dtype, int post, const void* Wfrag,
                      void* stream);
int pvo_colmean(const void* in, float* out, float* scratch /* E*32*C floats */, int E, int HW, int C, int dtype,
                void* stream);                                                                    /* gru.py:26 */
/* flow-encoder input for the tensor-core path: (M,8) half motion features -> (M,64) with channel dxi*8+c = feature c
 *   of the pixel dxi-3 columns away, so the 7x7 conv (droid_net.py:179) runs as a 7x1 conv over 64 channels */
int pvo_expand_x7(const void* motn, void* out, int nimg, int H, int Wd, void* stream);
/* tensor-core path: finish the spatial mean from the gate conv's per-tile sums (epilogue 2 of pvo_conv2d_tc) + mat-vecs */
int pvo_glo_reduce_matvec(const float* partial, int parts_per_img, int HW, const float* W, const float* b, float* out, int E,
                          void* stream);
int pvo_glo_matvec(const float* glo, const float* W, const float* b, float* out, int E, void* stream); /* gru.py:28-30 */
/* torch_scatter.scatter_mean over edges sharing ii (droid_net.py:83-88); ptr/idx/meta from pvo_group_edges */
int pvo_segment_mean(const void* in, int istride, int ioff, void* out, const int* ptr, const int* idx, const int* meta,
                     int Kmax, int HW, int C, int dtype, void* stream);
/* torch.unique(ii, return_inverse) + grouping, on the device: kx[Kcap], ptr[Kcap+1], idx[E], meta[3], frame2k[nframes] */
int pvo_group_edges(const long* ii, int E, int nframes, int t0, int t1, int Kcap, int* frame2k, int* kx, int* ptr,
                    int* idx, int* meta, void* stream);
int pvo_nchw_to_nhwc(const void* src, int sdtype, void* dst, int ddtype, int N, int C, int HW, int dstride, int doff,
                     void* stream);
int pvo_nhwc_to_nchw(const void* src, int sdtype, int sstride, int soff, void* dst, int ddtype, int N, int C, int HW,
                     void* stream);

/* ---- Panoptic-FPN forward helpers (VPS_Module/detectron2/modeling/...), pixel-major activations ---------------- */
int pvo_maxpool3s2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream);     /* resnet.py:356-359 */
/* 7x7/s2 stem as a 5x1 tensor-core conv: x (N,H,W,3) half -> 2x2 space-to-depth, x-expanded (N,H/2,W/2,64) half (resnet.py:340-354) */
int pvo_stem_expand(const void* x, void* E, int N, int H, int W, void* stream);
int pvo_subsample2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream);     /* strided 1x1 convs */
int pvo_upsample_add(void* lat, const void* top, int N, int H, int W, int C, int Ht, int Wt, int dtype, void* stream);  /* fpn.py:139-147 */
int pvo_upsample_bilinear(const void* in, void* out, int N, int H, int W, int C, int scale, int accumulate, int dtype,
                          void* stream);                                                                 /* semantic_seg.py:199-201,239-241 */
int pvo_group_norm(const void* in, void* out, const float* gamma, const float* beta, float* scratch, int N, int HW, int C,
                   int group_size, float eps, int relu, int dtype, void* stream);                        /* semantic_seg.py:186-196 */
/* paste_masks_in_image (layers/mask_ops.py:17-152, threshold 0.5): prob (R,S,S) float32 -> out (R,H,W) uint8 0/1; W % 4 == 0 */
int pvo_paste_masks(const float* prob, const float* boxes /*(R,4)*/, int R, int S, int H, int W, float thr, void* out, void* stream);
/* combine_semantic_and_instance_outputs (modeling/meta_arch/panoptic_fpn.py:497-588) for a batch of equal-size images,
 * one 8-CTA cluster per image, no host sync.  masks (R,H*W) uint8; boxes (R,4) or NULL (limits each scan to the box);
 * order: per-image ranks by descending score; sem (N,H*W) int64 labels -> pan (N,H*W) int32, info (N,max_info,5), counts (N) */
int pvo_panoptic_combine(const void* masks, const float* boxes, const float* scores, const long* classes, const long* order,
                         const int* inst_off, const long* sem, int* pan, int* info_out, int* counts_out, int N, int H, int W,
                         int num_sem, float overlap_thr, int stuff_area_thr, float score_thr, int thing_cat0, int max_info,
                         void* stream);

/* ROIPooler (modeling/poolers.py:189-235: level by box size, canonical 224 @ level 4) + ROIAlignV2
 * (layers/roi_align.py:49-58 -> torchvision roi_align aligned=True, sampling_ratio=0) over the 4 pixel-major pyramid
 * levels p2..p5 (strides 4..32).  rois (R,5) = {image index, x0, y0, x1, y1} float32; out (R, S, S, C) pixel-major. */
int pvo_roi_align_fpn(const void* const* feats, const int* Hs, const int* Ws, int C, const float* rois, int R, int S, void* out,
                      int dtype, void* stream);
/* find_top_rpn_proposals' decode stage (proposal_generator/proposal_utils.py:59-100) for all images and the 5 levels in one
 * launch: rpn_out[l] (N,h,w,16) half = objectness x3 | deltas 3x4, topk_idx[l] (N,ks[l]) indices into (h,w,3) ->
 * boxes (N,5,P,4) decoded + clipped to limits (N,4), scores (N,5,P) (-inf padding), valid (N,5,P).  cell_anchors: HOST (5,3,4). */
int pvo_rpn_decode(const void* const* rpn_out, const long* const* topk_idx, const int* hs, const int* ws, const int* ks, const int* strides,
                   const float* cell_anchors, const float* limits, int N, int P, float* boxes, float* scores, unsigned char* valid,
                   void* stream);
/* FastRCNNOutputLayers.predict_probs / predict_boxes (roi_heads/fast_rcnn.py:400-440): softmax + per-class decode
 * (weights 10,10,5,5) + clip; rois of image n are rows offsets[n]..offsets[n+1] -> class-major padded scores (N,K,P),
 * boxes (N,K,P,4), both zero-initialised by the caller. */
int pvo_det_decode(const float* cls /*(R,K+1)*/, const float* reg /*(R,4K)*/, const float* props /*(R,4)*/, const int* offsets,
                   const float* limits /*(N,4)*/, int R, int N, int P, int K, float* scores, float* boxes, void* stream);
/* Greedy NMS for G independent groups of n <= 1024 boxes sorted by descending score (layers/nms.py:9-22 batched_nms ->
 * torchvision nms per category): keep[g][i] = 1 iff box i is valid and no kept earlier box overlaps it with IoU > thr. */
int pvo_nms_groups(const float* boxes /*(G,n,4)*/, const unsigned char* valid /*(G,n)*/, int G, int n, float thr,
                   unsigned char* keep /*(G,n)*/, unsigned int* scratch /* G*n*32 words */, void* stream);

/* One-shot all-reduce (sum, fp64, in place) of the BA reduced camera system over NVLink peer memory: every rank pushes its
 * partial into a staging slot in each peer's symmetric buffer, flags it, and sums what arrived locally (one launch, graph
 * capturable).  stage_ptrs[i]: rank i's [2][world][n_max] doubles; flag_ptrs[i]: rank i's [world][256] ints (zeroed once);
 * epochs: local [256] ints (zeroed once).  Replaces the ncclAllReduce of SURVEY.md section 8(e). */
int pvo_allreduce_oneshot(double* data, void* const* stage_ptrs, void* const* flag_ptrs, int* epochs, int rank, int world, long n,
                          long n_max, void* stream);

/* heads' second convolution (droid_net.py:195-214), tensor-core form: T = block-diagonal 1x1 conv of the four heads' hidden
 * layers onto (tap, output) pairs, [M][128] half from pvo_conv2d_tc; this kernel sums the 9 shifted taps + bias -> out[M][ostride] f32 */
int pvo_heads_stencil(const void* T, const float* bias, float* out, int ostride, int nimg, int H, int W, void* stream);

/* Hardware probe (test infrastructure; no reference counterpart): one 128x64x64 tcgen05 MMA whose A descriptor starts row_off
 * rows into a swizzled shared-memory patch with 8-row groups sbo_bytes apart -- tools/probe_umma.py reads off which addressings work. */
int pvo_umma_probe(const void* G, const void* B, int rows_g, int row_off, int sbo_bytes, int base_offset, float* D, void* stream);

/* FactorGraph.update post-processing (factor_graph.py:250-283): target_cam, raw_mask, bin_mask, panoptic segment
 *   filter (:256-276, integer counting on the device instead of numpy on the host), delta_dy, sigmoid weights. */
long pvo_update_post_workspace_ints(int E);
int pvo_update_post(const float* coords1, const float* head_out /*(M,8)*/, const int* segm /*(E,ht,wd) or NULL*/,
                    double thresh, float* target, float* weight, float* raw_mask, float* delta_dy, float* full_flow,
                    int* seg_workspace, long seg_workspace_ints, int E, int ht, int wd, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVO_B200_H_ */
