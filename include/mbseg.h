/*
 * mbseg.h -- C ABI of libmbseg.so: the B200-native (sm_100a) segmentation hot path of microbeSEG.
 *
 * The reference (hip-satomi/microbeSEG) is pure Python; it has no FFI.  The drop-in boundary is
 * its L2 Python operator surface; the host-side mirror of that surface
 * lives in microbeseg_b200/ and binds the entry points below with ctypes.
 * Every entry point cites the reference code whose arithmetic it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = ok, nonzero = error, message via mbs_last_error() (thread local);
 *   - no allocation inside hot calls: scratch comes from a caller-owned workspace whose size is
 *     returned by the matching *_workspace_bytes() query;
 *   - activations are NHWC bf16; distance maps are float32 (H,W); masks are uint16 (H,W).
 */
#ifndef MBSEG_H_
#define MBSEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------- */
/* library                                                                                  */
/* ---------------------------------------------------------------------------------------- */
const char *mbs_last_error(void);
int mbs_version(void);
/* number of kernel launches issued by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
int64_t mbs_launch_count(int reset);
/* 1 if a tcgen05/TMA pipeline barrier timed out in any kernel since the last reset (the kernels
 * bail out instead of hanging the device); -1 if the flag could not be read. */
int mbs_debug_flags(int reset);

/* ---------------------------------------------------------------------------------------- */
/* U-Net building blocks (replace torch.nn/cuDNN calls made by src/utils/unets.py)           */
/* ---------------------------------------------------------------------------------------- */
enum { MBS_ACT_NONE = 0, MBS_ACT_RELU = 1, MBS_ACT_LEAKYRELU = 2, MBS_ACT_ELU = 3, MBS_ACT_MISH = 4 };
enum { MBS_IN_U8 = 0, MBS_IN_U16 = 1, MBS_IN_F32 = 2 };
enum { MBS_CONV3X3_S1 = 0, MBS_CONV3X3_S2 = 1, MBS_CONVT2X2_S2 = 2, MBS_CONV2X2_S2 = 3 /* 2x2 stride-2 conv: data
       gradient of the transposed conv; weights packed [Cout][4][Cin] */ };

/*
 * First encoder conv fused with min-max normalisation and top/left padding.
 * Replaces: frame min/max normalisation `2*(f32(img)-min)/(max-min)-1`
 *   (src/inference/infer.py:346, infer_script_local.py:130), zero_pad_model_input
 *   (src/utils/utils.py:124-163, pads TOP/LEFT with pad_val=frame_min, i.e. -1 after
 *   normalisation) and ConvBlock's first Conv2d(1,C,3,p=1)+act+BatchNorm2d(eval)
 *   (src/utils/unets.py:112-134).
 * img: raw frame (H,W) of in_dtype; output (Hp,Wp,C) bf16 with Hp=H+pad_y, Wp=W+pad_x.
 * weight: [C][9] f32 (ky*3+kx), bias/scale/shift: [C] f32 (scale/shift = folded eval BN).
 * norm_lo/norm_hi: frame min / max as floats; pad pixels take the value norm_lo.  norm_hi < norm_lo
 * means "img is already normalised" (values pass through; used by the drop-in net(x) call).
 * norm_hi == norm_lo reproduces the reference's unguarded 0/0 (NaN maps -> empty mask).
 * lohi_dev (optional): device float[2] = {min, max} from mbs_frame_minmax; overrides norm_lo/hi. */
int mbs_first_conv(const void *img, int in_dtype, int H, int W, int pad_y, int pad_x, float norm_lo,
                   float norm_hi, const float *lohi_dev, const float *weight, const float *bias, const float *scale,
                   const float *shift, int C, int act, void *out_nhwc_bf16, int out_ld, int out_coff,
                   void *stream);

/* Frame min / max on the device (replaces np.min/np.max of the raw frame, infer_script_local.py:124,
 * src/inference/infer.py:253): lohi_dev[0] = min, lohi_dev[1] = max as float; scratch8 = 8 bytes. */
int mbs_frame_minmax(const void *img, int in_dtype, long long n, float *lohi_dev, void *scratch8, void *stream);

/*
 * Implicit-GEMM convolution on the tcgen05 tensor cores (TMA-fed, TMEM accumulators, fused
 * epilogue bias -> activation -> BatchNorm(eval) affine -> bf16).
 * Replaces: nn.Conv2d(3x3,s1,p1) / nn.Conv2d(3x3,s2,p1) / nn.ConvTranspose2d(2x2,s2) followed by
 *   the activation and nn.BatchNorm2d of ConvBlock / ConvPool / TranspConvBlock
 *   (src/utils/unets.py:92-173, 176-226, 229-264), the channel concat torch.cat([up, skip],1)
 *   (src/utils/unets.py:492,502; expressed as two K sources), and optionally the final
 *   Conv2d(C,1,1) head (src/utils/unets.py:460-461,494,504) fused into the epilogue.
 */
typedef struct {
    int mode;            /* MBS_CONV3X3_S1 / MBS_CONV3X3_S2 / MBS_CONVT2X2_S2 */
    int N, H, W;         /* input batch / height / width (both sources) */
    /* source 0 and (optional) source 1: NHWC bf16 views with `ld` channels per pixel, the
     * source's channels start at element offset `coff` and number `C` (multiple of 64). */
    const void *src0; int C0, ld0, coff0;
    const void *src1; int C1, ld1, coff1;
    /* packed weights bf16: conv: [Cout][9][C0+C1]; convT: [4*Cout][C0] with row (dy*2+dx)*Cout+co */
    const void *weight;
    int Cout;            /* output channels (multiple of 64) */
    const float *bias, *scale, *shift;   /* [Cout] f32 */
    int act;
    /* destination NHWC bf16 view (may be NULL when only the head output is wanted) */
    void *dst; int ldd, coffd;
    /* optional fused 1x1 head(s) (requires Cout == 64, head_n <= 4):
     * head_out[n][h][y][x] = sum_c y_c * head_w[h][c] + head_b[h]   (Conv2d(64, ch_out, 1), unets.py:347,460-461) */
    const float *head_w; float head_b[4]; int head_n; float *head_out;
} mbs_conv_desc;

int mbs_conv_gemm(const mbs_conv_desc *d, void *stream);

/* The first layer FUSED into the second convolution of the top encoder block (enc0b, 64 -> 64 channels, 3x3 stride 1):
 * producer warps of the tensor-core kernel compute the first layer's halo patches straight into shared memory, so its
 * 64-channel full-resolution output never exists in global memory.  Results are bit-identical to mbs_first_conv followed
 * by mbs_conv_gemm(conv2).  img: N contiguous frames [N][H][W]; conv2: mode MBS_CONV3X3_S1, N, H + pad_y, W + pad_x,
 * C0 = 64, C1 = 0, Cout = 64, no head (src0 is ignored).  Replaces unets.py:112-134 of the first ConvBlock. */
int mbs_first_conv_halo64(const void *img, int in_dtype, int N, int H, int W, int pad_y, int pad_x, float norm_lo,
                          float norm_hi, const float *lohi_dev, const float *weight, const float *bias, const float *scale,
                          const float *shift, int act, const mbs_conv_desc *conv2, void *stream);
/* pack helpers: f32 reference-layout weights -> bf16 GEMM layout (device to device) */
int mbs_pack_conv3x3_weight(const float *w_oihw, int Cout, int Cin, void *packed_bf16, void *stream);
int mbs_pack_convT2x2_weight(const float *w_iohw, int Cin, int Cout, void *packed_bf16, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* distance post-processing (replaces src/inference/postprocessing.py:7-59)                 */
/* ---------------------------------------------------------------------------------------- */
size_t mbs_postproc_workspace_bytes(int H, int W);
/*
 * border, cell: float32 device arrays, element (y,x) at [y*ld + x] (ld in elements; a crop of a
 * padded network output is expressed through the base pointer and ld).  out: uint16 (H,W) dense.
 * info_host (optional, 8 x int64 on the host): [0]=#seed components before filtering,
 *   [1]=#markers, [2]=#minimax relaxation sweeps, [3]=1 if the exact sequential flood was needed
 *   (value ties), [4]=#ambiguous pixels, others reserved.  Passing info_host forces a stream sync.
 */
int mbs_distance_postprocessing(const float *border, const float *cell, int H, int W, int ld,
                                float th_seed, float th_cell, uint16_t *out, void *workspace,
                                size_t workspace_bytes, int64_t *info_host, void *stream);

/* Softmax over the three class planes of the boundary network, crop of the pads and channel-last layout in one pass
 * (replaces F.softmax(prediction, dim=1)[0, :, pads...].permute(1, 2, 0), src/inference/infer.py:371-374).
 * logits: 3 planes `plane_stride` floats apart, row pitch ld; (y0, x0): first kept pixel; prob: (H,W,3) float32. */
int mbs_softmax3_hwc(const float *logits, size_t plane_stride, int ld, int y0, int x0, int H, int W, float *prob, void *stream);
/* boundary method (replaces src/inference/postprocessing.py:62-90): prediction = softmax probabilities
 * (H,W,3) float32, channel-last as the reference passes them; the flood image is flat, so the result is
 * pure FIFO order and is produced by the exact sequential flood whenever two markers share a mask region. */
int mbs_boundary_postprocessing(const float *prediction_hwc, int H, int W, uint16_t *out, void *workspace,
                                size_t workspace_bytes, int64_t *info_host, void *stream);

/* individual stages, exposed for stage-level parity tests */
int mbs_pp_front(const float *border, const float *cell, int H, int W, int ld, float th_seed,
                 float th_cell, float *cell_smooth, uint8_t *mask, uint8_t *seed, void *stream);
int mbs_pp_label8(const uint8_t *binary, int H, int W, int32_t *labels, int32_t *n_out,
                  void *workspace, size_t workspace_bytes, void *stream);
/* skimage.measure.label of an INTEGER instance image (8-connectivity, equal non-zero values; labels 1..n in raster
 * order of each component's first pixel) -- what eval.py:261,313 apply to ground truth and prediction before AJI+.
 * workspace: mbs_postproc_workspace_bytes(H, W). */
int mbs_label8_instances(const uint16_t *image, int H, int W, int32_t *labels, int32_t *n_out,
                         void *workspace, size_t workspace_bytes, void *stream);
int mbs_pp_watershed(const float *image, const int32_t *markers, const uint8_t *mask, int H, int W,
                     int32_t *labels_out, void *workspace, size_t workspace_bytes,
                     int64_t *info_host, int force_sequential, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* training step (replaces torch autograd + cuDNN/ATen in src/training/train.py:460-493)     */
/* activations / activation gradients NHWC bf16; statistics and parameter gradients fp32      */
/* ---------------------------------------------------------------------------------------- */
/* BatchNorm2d, training mode (unets.py:128,153,206,246): batch mean / biased variance over the M = N*H*W rows of
 * a [M][C]; y = gamma*(a-mean)*invstd + beta.  sums_scratch: mbs_bn_scratch_floats(C) floats (per-block partial sums,
 * reduced without atomics: the statistics are deterministic).  running_mean / running_var / num_batches_tracked
 * (optional, NULL to skip) are updated like nn.BatchNorm2d in training mode (momentum, unbiased variance).
 * act = MBS_ACT_NONE: `a` is the post-activation tensor; MBS_ACT_MISH: `a` holds the PRE-activation z and mish(z)
 * (unets.py:81-89) is applied on load (training keeps one tensor per layer; mish' needs z). */
size_t mbs_bn_scratch_floats(int C);
int mbs_bn_train_fwd(const void *a, long long M, int C, const float *gamma, const float *beta, float eps, void *y,
                     float *sums_scratch, float *mean, float *invstd, float momentum, float *running_mean,
                     float *running_var, long long *num_batches_tracked, int act, void *stream);
/* backward of  y = BN(act(z)):  dz = act'(a) * gamma*invstd*(dy - dbeta/M - xhat*dgamma/M);  dgamma_dbeta: [2*C]
 * (dgamma then dbeta), dbias[c] = sum dz (gradient of the conv bias).  act: MBS_ACT_RELU (a = post-activation),
 * MBS_ACT_MISH (a = pre-activation z) or MBS_ACT_NONE.
 * scratch: mbs_bn_scratch_floats(C) floats. */
int mbs_bn_train_bwd(const void *dy, const void *a, long long M, int C, const float *mean, const float *invstd,
                     const float *gamma, int act, void *dz, float *dgamma_dbeta, float *dbias, float *scratch, void *stream);
/* nn.GroupNorm(groups, C) (normalization 'gn', groups = 8) / nn.InstanceNorm2d(C) ('in', groups = C, gamma = beta = NULL)
 * of ONE sample in eval mode (unets.py:129-132,155-158,207-210,247-250): a, y = [M = H*W][C] bf16 (y may alias a),
 * statistics over the sample's own pixels, biased variance.  scratch: mbs_bn_scratch_floats(C) floats. */
int mbs_sample_group_norm(const void *a, long long M, int C, int groups, const float *gamma, const float *beta, float eps,
                          void *y, float *scratch, void *stream);
/* Conv2d(C,1,1) head, SmoothL1Loss(beta=1,'mean') (losses.py:30-32) and their gradients */
int mbs_head_fwd(const void *y, long long M, int C, const float *w, const float *b_dev, float *pred, void *stream);
int mbs_smoothl1(const float *pred, const float *target, long long M, float *loss_accum, float *grad, void *stream);
/* the three distance-method criteria of get_loss (losses.py:24-35), reduction 'mean': kind 0 smooth_l1, 1 l1, 2 l2;
 * *loss_accum += loss (zeroed by the caller), grad = dloss/dpred */
int mbs_regression_loss(const float *pred, const float *target, long long M, int kind, float *loss_accum, float *grad,
                        void *stream);
/* scratch: >= 592 * (C + 1) floats (per-block partial sums, reduced in a fixed order: bitwise reproducible) */
int mbs_head_bwd(const float *g, const void *y, long long M, int C, const float *w, void *dy, float *dw_db, float *scratch,
                 void *stream);
/* layout / glue kernels of the backward pass */
/* data-gradient filter of a 3x3 conv: packed[ci][tap][co] = bf16(w[co][ci][8 - tap]) from the reference-layout weight */
int mbs_pack_conv3x3_dgrad(const float *w, int Cout, int Cin, void *packed, void *stream);
/* Every GEMM-packed bf16 weight of one training step in ONE launch (replaces the per-layer mbs_pack_* calls inside the
 * reference's step, src/training/train.py:473-493, where cuDNN re-reads the fp32 weights itself).  One job per layer:
 *   kind 0, Conv2d 3x3        w[cout][cin][3][3]: fwd[co][tap][ci] = w[co][ci][tap], dgrad[ci][tap][co] = w[co][ci][8 - tap]
 *   kind 1, ConvTranspose 2x2 w[cin][cout][2][2]: fwd[q*cout + co][ci] = w[ci][co][q], dgrad[ci][q][co] = w[ci][co][q]
 * cout and cin multiples of 32; fwd / dgrad may be NULL; one CTA per 32 x 32 channel tile, tile0 = prefix sum of the
 * jobs' tile counts (cout/32 * cin/32), total_tiles = their sum.  jobs_dev: DEVICE array. */
typedef struct {
    const void *w;
    void *fwd, *dgrad;
    int cout, cin, kind, tile0;
} mbs_pack_job;
int mbs_pack_train_weights(const mbs_pack_job *jobs_dev, int n_jobs, int total_tiles, void *stream);
/* weight gradient g[co][tap][ci] (mbs_conv_wgrad layout) -> reference layout out[co][ci][3][3] */
int mbs_unpack_conv3x3_grad(const float *g, int Cout, int Cin, float *out, void *stream);
/* nn.MaxPool2d(2, 2) on contiguous NHWC bf16 (build_unet pool_method = 'max', unets.py:306-307,363-364) */
int mbs_maxpool2x2(const void *src, int N, int H, int W, int C, void *dst, void *stream);
int mbs_zero_insert_up2(const void *src, int N, int H, int W, int C, void *dst, void *stream);
int mbs_add3_bf16(const void *a, const void *b, const void *c, long long n, void *out, void *stream);
/* scratch: >= 592 * C * 9 floats (mbs_bn_scratch_floats(2048) is enough); per-block partial sums, reduced in a fixed order */
int mbs_first_conv_wgrad(const float *x, const void *dz, int N, int H, int W, int C, float *dw, float *scratch, void *stream);
/* weight gradient on the tensor cores, straight from the NHWC bf16 activations (no layout copies):
 *   out[m][tap][out_coff + n] += sum over pixels o of a[sA*o + offA(tap)][m] * b[sB*o + offB(tap)][n]
 * kind 0/1: conv3x3 stride 1/2 -- a = dz [N,Ho,Wo,Cm], b = x [N,s*Ho,s*Wo,Cn] (autograd of unets.py:112-134,
 * 196-199); kind 2: transposed conv 2x2 -- a = d(up) [N,2Ho,2Wo,Cm], b = x [N,Ho,Wo,Cn] (unets.py:245).
 * ld / coff: pixel stride and channel offset (elements) of the views; out is fp32, zeroed by the caller. */
typedef struct {
    int kind, N, Ho, Wo;
    const void *a; int Cm, lda, coffa;
    const void *b; int Cn, ldb, coffb;
    float *out; int out_ld, out_coff;
    int partial;    /* 0: accumulate into the zeroed `out` with vector reductions (summation order varies from run to run);
                     * 1: deterministic -- `out` is a scratch buffer [splits][Cm][taps][Cn] (splits = mbs_conv_wgrad_splits(d),
                     *    out_ld / out_coff ignored), every K split stores its tile; mbs_wgrad_reduce sums them in order */
} mbs_wgrad_desc;
int mbs_conv_wgrad(const mbs_wgrad_desc *d, void *stream);
/* number of K splits mbs_conv_wgrad uses for this descriptor on the current device (-1: bad descriptor) */
int mbs_conv_wgrad_splits(const mbs_wgrad_desc *d);
/* Ordered sum over the K splits of up to two partial buffers (the two sources of a concatenated input; part1 may be NULL)
 * and the layout change to the reference's parameter layout in one pass:
 *   layout 0 (Conv2d 3x3):        out[co][coff_s + ci][3][3] = sum_k part_s[k][co][tap][ci]     (out: [Cm][cn0 + cn1][3][3])
 *   layout 1 (ConvTranspose 2x2): out[ci][co][2][2]          = sum_k part0[k][co][q][ci]        (out: [cn0][Cm][2][2]) */
int mbs_wgrad_reduce(const float *part0, int splits0, int cn0, const float *part1, int splits1, int cn1, int Cm, int layout,
                     float *out, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* Ranger optimizer step (RAdam + gradient centralisation + lookahead), replaces the per-tensor loop of    */
/* src/training/ranger2020.py:101-208 (the reference's default --optimizer) with one launch over a table   */
/* of all parameter tensors.  Rows: output-channel slices of dim>1 tensors (gc = 1: the gradient mean over  */
/* the row is subtracted first, :30-40) or 1024-element chunks of 1-D tensors (gc = 0); row_start = prefix  */
/* sum of `rows`.  step_lr = step_size*lr, use_denom = (N_sma > threshold) from the host schedule (:165-180)*/
/* ---------------------------------------------------------------------------------------- */
typedef struct {
    float *p, *g, *exp_avg, *exp_avg_sq, *slow;   /* fp32 device tensors of identical shape */
    long long numel;
    int rows, row_len, gc, row_start;
} mbs_ranger_tensor;
int mbs_ranger_step(const mbs_ranger_tensor *tensors_dev, int n_tensors, int total_rows, float beta1, float beta2,
                    float eps, float weight_decay, float step_lr, int use_denom, int lookahead, float alpha, void *stream);
/* Fused Adam / AMSGrad step for every parameter tensor in one launch: the reference's Adam recipe
 * torch.optim.Adam(lr=8e-4, betas=(0.9,0.999), eps=1e-8, weight_decay=0, amsgrad=True) (src/training/train.py:380-385),
 * arithmetic of torch/optim/adam.py::_single_tensor_adam.  Same table type as mbs_ranger_step (`slow` = max_exp_avg_sq,
 * gc = 0, rows = 1024-element chunks); step_size = lr / (1 - beta1^step), bias_correction2_sqrt = sqrt(1 - beta2^step). */
int mbs_adam_step(const mbs_ranger_tensor *tensors_dev, int n_tensors, int total_rows, float beta1, float beta2, float eps,
                  float weight_decay, float step_size, float bias_correction2_sqrt, int amsgrad, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* training-label generation, distance method (replaces src/training/train_data_representations */
/* .py:261-361 distance_label + :102-126 border_label + :40-72 bottom_hat_closing, and the     */
/* max_mal of src/training/train.py:74-79); batched over crops                                 */
/* ---------------------------------------------------------------------------------------- */
size_t mbs_labels_workspace_bytes(int n_crops, int H, int W, int max_id);
/* label types 'boundary' (mode 0, train_data_representations.py:75-99) and 'border' (mode 1, :102-126) of the
 * baseline boundary method: uint8 [n_crops][H][W], 0 background / 1 nucleus / 2 boundary resp. touching border */
int mbs_boundary_border_labels(const uint16_t *masks, int n_crops, int H, int W, int mode, uint8_t *out, void *stream);
/* j4_label (train_data_representations.py:158-190; 0 background, 1 cell, 2 touching, 3 gap): bottom hat of the binary mask with
 * disk(se_radius) (scipy binary_closing, border_value 0) and the "more than one instance in the (2k+1)^2 window" test of
 * compute_neighbor_instances (:193-217).  tmp: n_crops*H*W bytes of scratch. */
int mbs_j4_labels(const uint16_t *masks, int n_crops, int H, int W, int k_neighbors, int se_radius, uint8_t *out, uint8_t *tmp,
                  void *stream);
/* masks: uint16 [n_crops][H][W] instance ids (0 = background, ids <= max_id).
 * max_mal_out: int32 [n_crops] = int(ceil(max regionprops.major_axis_length)) per crop. */
int mbs_labels_max_mal(const uint16_t *masks, int n_crops, int H, int W, int max_id, int32_t *max_mal_out,
                       void *workspace, size_t workspace_bytes, void *stream);
/* search_radius >= 0: used for every crop (get_label passes int(ceil(0.75*max_mal)));
 * search_radius < 0: derived per crop from its own max_mal on the device; radius_hint (>= the
 * largest radius in the batch, or -1) only sizes the shared-memory window.
 * cell_dist / neighbor_dist: float32 [n_crops][H][W].  error_out (optional, int32 [n_crops]):
 * bit 0 = the bounding box of an instance wider than 58 px did not fit the closing kernel's shared memory
 * (> ~110 k pixels), bit 1 = more than 4096 gap components.  EDT search windows of any size are handled
 * (shared memory, or a per-crop global buffer for windows that do not fit). */
int mbs_distance_labels(const uint16_t *masks, int n_crops, int H, int W, int max_id, int search_radius,
                        int radius_hint, float *cell_dist, float *neighbor_dist, int32_t *max_mal_out,
                        int32_t *error_out, void *workspace, size_t workspace_bytes, void *stream);
/* The same with cell_clip > 0: cell_dist = clip(EDT, 0, cell_clip) / cell_clip instead of the per-instance normalisation --
 * cell_distance_label(apply_clipping=True, clip_val) of the 'cell_dist_clipped' label type (train_data_representations.py:246-256).
 * cell_clip <= 0 is mbs_distance_labels. */
int mbs_distance_labels_ex(const uint16_t *masks, int n_crops, int H, int W, int max_id, int search_radius,
                           int radius_hint, float cell_clip, float *cell_dist, float *neighbor_dist, int32_t *max_mal_out,
                           int32_t *error_out, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Threshold sweep of the evaluation (src/evaluation/eval.py:128-129, 395-412: one prediction post-processed for every
 * pair of product(th_cell, th_seed)).  Smoothing, seed image, 8-connected labelling and the area filter depend on
 * th_seed only and run once per seed threshold; each cell threshold then costs one watershed.
 * th_*_host: HOST float arrays; out: uint16 [n_seed][n_cell][H][W], each plane identical to
 * mbs_distance_postprocessing(.., th_seeds[is], th_cells[ic], ..); info_host (optional): int64 [n_seed*n_cell*8].
 * Workspace: mbs_postproc_workspace_bytes(H, W).
 */
int mbs_distance_postprocessing_sweep(const float *border, const float *cell, int H, int W, int ld, const float *th_seeds_host,
                                      int n_seed, const float *th_cells_host, int n_cell, uint16_t *out, void *workspace,
                                      size_t workspace_bytes, int64_t *info_host, void *stream);

/*
 * Criteria of the boundary method (src/training/losses.py:16-21, 71-96; 3 classes): nn.CrossEntropyLoss (with_dice = 0)
 * or ce_dice = cross entropy + 0.5 * sum_c c * dice_c over the channels 1, 2 of softmax(logits) (with_dice = 1).
 * logits: float32 planar [3][M] (M = N*H*W pixels of the batch), labels: uint8 [M] in {0,1,2};
 * loss_accum += loss (device float, caller zeroes it); grad: float32 [3][M] = dloss/dlogits; sums7: device double[7] scratch.
 */
int mbs_ce_dice_loss(const float *logits, const uint8_t *labels, long long M, int with_dice, float *loss_accum, float *grad,
                     double *sums7, void *stream);

/*
 * fp32 check mode: one convolution of the network as a plain CUDA-core fp32 direct convolution, no bf16,
 * no tensor cores.  Same operators as mbs_conv_gemm (Conv2d 3x3 s1 / s2, ConvTranspose2d 2x2 s2, 1x1 head; bias, activation,
 * folded eval BatchNorm; torch.cat as two sources), float32 NHWC in and out, weights float32 [tap][Cin][Cout].
 * mode: 0 = 3x3 s1, 1 = 3x3 s2, 2 = transposed 2x2 s2, 4 = 1x1.  Diagnostic path (microbeseg_b200.unets.forward_check_fp32).
 */
typedef struct {
    int mode, N, H, W;
    const float *src0; int C0;
    const float *src1; int C1;
    const float *weight; int Cout;
    const float *bias, *scale, *shift;       /* scale / shift may be NULL (identity) */
    int act;
    float *dst;
} mbs_convref_desc;
int mbs_conv_ref_f32(const mbs_convref_desc *d, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* mask -> polygon ROI encoding                                           */
/* ---------------------------------------------------------------------------------------- */
/*
 * Outer contour of every instance of a uint16 mask, in the point order of
 * cv2.findContours(RETR_TREE, CHAIN_APPROX_NONE) on the instance's own crop.
 * Replaces: get_indices_pandas (src/utils/hull_polygon.py:8-42) + cv2_countour (:45-89), called once per cell from
 *   src/inference/infer.py:273-289.  Scope: each id is one 8-connected component (watershed labels); for an
 *   instance with holes the outer contour is returned (the reference's result when its shapely test holds, :61-75).
 * mbs_contour_first: first[l] = linear index of the first raster pixel of id l+1 (INT32_MAX if absent), l < n_labels.
 * mbs_contour_trace: offsets == NULL -> counting pass, counts[l] = number of contour points of id l+1;
 *                    offsets != NULL -> writing pass, points_yx[2*(offsets[l]+k)] = y, [...+1] = x of point k.
 * overflow (device int32, caller zeroes it): set if a walk hit its step guard (cannot happen for a valid mask).
 */
int mbs_contour_first(const uint16_t *mask, int H, int W, int n_labels, int32_t *first, void *stream);
int mbs_contour_trace(const uint16_t *mask, int H, int W, int n_labels, const int32_t *first, const int64_t *offsets,
                      int32_t *counts, int32_t *points_yx, int32_t *overflow, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* per-frame analysis                                                     */
/* ---------------------------------------------------------------------------------------- */
/*
 * Area and major / minor axis length (skimage regionprops: 4*sqrt of the inertia tensor eigenvalues, float64) of every
 * instance id of every frame.  Replaces the regionprops loop of src/inference/analysis.py:153-166 (per-frame counts,
 * mean area, mean axis lengths of the analysis export).  Outputs are indexed [frame * (max_id+1) + id]; area 0 = id absent.
 * Workspace: (max_id+1) * n_frames * 96 bytes (rounded up to 256).
 */
int mbs_instance_stats(const uint16_t *masks, int n_frames, int H, int W, int max_id, int32_t *area, double *major,
                       double *minor, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* training-time augmentations, batched on the device (replace the per-sample NumPy / imgaug /  */
/* scipy transforms of src/training/mytransforms.py:13-406 that run in the DataLoader workers). */
/* The random decisions and parameters are drawn on the HOST in the reference's own call order  */
/* (microbeseg_b200/augment.py); these entry points apply them.  uint16 images [n][H][W].       */
/* ---------------------------------------------------------------------------------------- */
/* Flip (mytransforms.py:129-233, 8 dihedral maps: exact), Scaling (:326-379) and Rotate (:271-323) (imgaug Affine =
 * cv2.warpAffine about the image centre, constant border 0).  mats_dev: [n][6] float64 INVERSE maps,
 * (sx, sy) = M * (x, y, 1) for output pixel (x, y); modes_dev: [n] 0 copy, 1 nearest (order 0: uint8 labels, dihedral
 * maps), 2 bilinear (order 1: images and float labels).  dtype: MBS_IN_U8 / U16 / F32.  src != dst. */
int mbs_aug_warp(const void *src, void *dst, int dtype, int n, int H, int W, const double *mats_dev, const int *modes_dev,
                 void *stream);
size_t mbs_aug_workspace_bytes(int n);
/* Contrast (mytransforms.py:64-126) in place: modes_dev [n]: 0 none, 1 percentile stretch (np.percentile(img, (q0, q1))
 * + skimage rescale_intensity to the dtype range), 2 contrast + gamma adjustment; params_dev [n][4] float64 = {q0, q1,
 * factor, gamma}.  The CLAHE branch (skimage equalize_adapthist) is not built. */
int mbs_aug_contrast(uint16_t *img, int n, int H, int W, const int *modes_dev, const double *params_dev, void *workspace,
                     size_t workspace_bytes, void *stream);
/* Blur (mytransforms.py:38-61) in place: scipy.ndimage.gaussian_filter(img (H,W,1), sigma) bit for bit (float64
 * accumulation in scipy's order, 'reflect', truncation into uint16 after each of the three axis passes).
 * weights_dev: [n][17] float64 kernels from the host (numpy's arithmetic), radius_dev: [n] (0 = no blur, <= 8). */
int mbs_aug_blur(uint16_t *img, uint16_t *tmp, int n, int H, int W, const double *weights_dev, const int *radius_dev, void *stream);
/* Noise (mytransforms.py:236-268: additive Gaussian noise, sigma = noise_frac * max(img), samples rounded, clipped to
 * uint16; counter-based generator, reproducible for a seed) + ToTensor's min_max_normalization (utils.py:50-74):
 * out = 2 * (clip(img, min, max) - min) / (max - min) - 1 in float32.  img_out / out: either may be NULL. */
int mbs_aug_noise_normalize(const uint16_t *img, int n, int H, int W, const float *noise_frac_dev, unsigned long long seed,
                            float min_value, float max_value, uint16_t *img_out, float *out, void *workspace,
                            size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MBSEG_H_ */
