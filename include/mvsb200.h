/*
 * mvsb200.h -- C ABI of libmvsb200.so, the B200 (sm_100a) plane-sweep cost-volume engine.
 *
 * This is the drop-in boundary for the hot path of fdarmon/wild_deep_mvs.  The reference has
 * no FFI layer (it is pure PyTorch); each entry point below replaces a Python function or
 * nn.Module.forward of the reference, cited as file:line.  A reference maintainer binds these
 * with ctypes; wild_deep_mvs_b200/_lib.py is exactly that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented as "host";
 *   - all tensors are dense fp32;
 *   - feature maps are channels-last  [B, H, W, C]      (== torch NCHW tensor in channels_last);
 *   - volumes are channels-last       [B, D, H, W, C]   (== torch NCDHW tensor in channels_last_3d);
 *   - single-channel volumes/maps are therefore plain [B, D, H, W] / [B, H, W];
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs;
 *   - no entry point allocates device memory; the caller owns every buffer;
 *   - return value 0 = success, negative = MVSB200_E_* (message via mvsb200_last_error()).
 */
#ifndef MVSB200_H_
#define MVSB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MVSB200_ABI_VERSION 6

#define MVSB200_OK 0
#define MVSB200_E_INVALID (-1)  /* bad argument / unsupported shape */
#define MVSB200_E_CUDA (-2)     /* a CUDA runtime call or launch failed */
#define MVSB200_E_NODEVICE (-3) /* no sm_100-class device */

#define MVSB200_MAX_SRC 16 /* maximum number of source views per call */

typedef void *mvsb200_stream_t;

#if defined(__GNUC__)
#define MVSB200_API __attribute__((visibility("default")))
#else
#define MVSB200_API
#endif

MVSB200_API int mvsb200_abi_version(void);
/* Thread-local message describing the last failure on the calling thread ("" if none). */
MVSB200_API const char *mvsb200_last_error(void);
/* Fills SM count and compute capability of the current device. */
MVSB200_API int mvsb200_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------------------------------
 * Geometry prologues (one thread per (batch, source view); fp64 inside, fp32 out).
 * --------------------------------------------------------------------------------------- */

/* warp[b][s] = { rot[9] row-major, trans[3], 0,0,0,0 } with  [rot|trans] = (src_proj @ inv(ref_proj))[:3,:4]
 * Replaces `proj = torch.matmul(src_proj, torch.inverse(ref_proj))`
 *   models/MVSNet/module.py:128, models/CVP_MVSNet/models/modules.py:89-97,247-252.
 * ref_proj [B,4,4], src_proj [B,S,4,4], warp [B,S,16]. */
MVSB200_API int mvsb200_mvs_relative_proj(const float *ref_proj, const float *src_proj, float *warp, int B, int S,
                              mvsb200_stream_t stream);

/* Closed form of get_homographies (models/VisMVSNet/homography.py:23-74) after
 * scale_camera(cam, scale) (models/VisMVSNet/preproc.py:63-92):
 *   H_d p = A p - b (n.p) / (depth_d + 1e-9),  A = K_s R_s R_r^T K_r^-1,  b = K_s R_s (c_s - c_r),
 *   n = R_r[2,:] R_r^T K_r^-1.      warp[b][s] = { A[9], b[3], n[3], 0 }.
 * ref_cam [B,2,4,4], src_cam [B,S,2,4,4] ([.,0]=R|t, [.,1]=K as built by fill_cam_array,
 * models/VisMVSNet/frontend.py:14-24). */
MVSB200_API int mvsb200_vis_homography_params(const float *ref_cam, const float *src_cam, float scale, float *warp, int B,
                                  int S, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K1: fused homography warp + cross-view aggregation.  The warped volumes, sampling grids
 * and per-voxel homographies of the reference are never stored.
 * --------------------------------------------------------------------------------------- */
#define MVSB200_GEOM_MVS 0 /* homo_warping: models/MVSNet/module.py:111-169, CVP modules.py:74-128,229-286 */
#define MVSB200_GEOM_VIS 1 /* homography_warping + interpolate: models/VisMVSNet/homography.py:77-121 */

#define MVSB200_AGG_VARIANCE 0      /* M2/V - M1^2/V^2    models/MVSNet/model.py:113-139              */
#define MVSB200_AGG_VARIANCE_MEAN 1 /* M2/V - (M1/V)^2    CVP net.py:152, modules.py:289               */
#define MVSB200_AGG_SOFTMIN 2       /* MVSNet-s           models/MVSNet/model.py:141-173               */
#define MVSB200_AGG_GROUPCORR 3     /* 8-group correlation per source view, VisMVSNet/nn_utils.py:473-490,
                                       model_cas.py:176-186,340 */

#define MVSB200_DEPTH_VALUES 0       /* depth [B,D]            (module.py:140-141)                      */
#define MVSB200_DEPTH_VOLUME 1       /* depth [B,D,H,W]        (CVP modules.py:262-263)                 */
#define MVSB200_DEPTH_START 2        /* depth [B] start, interval [B]:  start + interval*d              */
#define MVSB200_DEPTH_START_MAP 3    /* depth [B,H,W] start, interval [B]  (homography.py:24-40)        */

typedef struct {
    int geom, agg, depth_mode;
    int B, S, C;  /* batch, number of SOURCE views (<= MVSB200_MAX_SRC), feature channels (8,16 or 32) */
    int D, H, W;  /* hypotheses and reference feature size */
    int groups;   /* GROUPCORR only (C/groups must be 4) */
    int src_h[MVSB200_MAX_SRC], src_w[MVSB200_MAX_SRC]; /* source maps may differ in size (module.py:119-125) */
    long long out_view_stride; /* GROUPCORR: element stride between the per-source output volumes */
} mvsb200_cost_volume_desc;

/* ref [B,H,W,C]; src: HOST array of S device pointers, src[s] is [B,src_h[s],src_w[s],C];
 * warp [B,S,16] from one of the geometry prologues; depth/interval per depth_mode; temp: device scalar
 * (SOFTMIN only, models/MVSNet/model.py:94-95).
 * out: [B,D,H,W,C] (VARIANCE*, SOFTMIN) or S volumes [B,D,H,W,groups] (GROUPCORR).
 * out_amax: optional device scalar, atomically max-ed with max|out| (zero it before the launch); the z-march conv
 * engine takes it as x_amax so the volume is not read a second time.
 * Alignment: ref, every src[s] and out (and, for the backward entry point, grad_out / grad_ref / grad_src[s]) must be
 * 32-byte aligned -- the kernels move a pixel's channels with 256-bit accesses; violated -> MVSB200_E_INVALID. */
MVSB200_API int mvsb200_build_cost_volume(const mvsb200_cost_volume_desc *desc, const float *ref, const float *const *src,
                              const float *warp, const float *depth, const float *interval, const float *temp,
                              float *out, float *out_amax, mvsb200_stream_t stream);

/* Backward of mvsb200_build_cost_volume with respect to the feature maps (the reference computes the sampling grid
 * under torch.no_grad(): models/MVSNet/module.py:127, VisMVSNet/homography.py:25,110 -- cameras and hypotheses get no
 * gradient; this replaces grid_sampler_2d_backward + the autograd graph of models/MVSNet/model.py:113-173,
 * VisMVSNet/model_cas.py:176-186, CVP modules.py:229-293).  Inputs as in the forward call; grad_out has the forward
 * output's layout.  grad_ref [B,H,W,C], grad_src[s] [B,src_h[s],src_w[s],C] (HOST array of S device pointers) and
 * grad_temp (device scalar, SOFTMIN only) must be ZEROED by the caller: the kernel adds into them (vector atomics). */
MVSB200_API int mvsb200_build_cost_volume_backward(const mvsb200_cost_volume_desc *desc, const float *ref, const float *const *src,
                                                   const float *warp, const float *depth, const float *interval,
                                                   const float *temp, const float *grad_out, float *grad_ref,
                                                   float *const *grad_src, float *grad_temp, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K2: 3-D convolution / transposed convolution with fused BN (scale,bias) + ReLU + skip.
 * Replaces ConvBnReLU3D (models/MVSNet/module.py:41-48), the nn.ConvTranspose3d+BN+ReLU blocks
 * and skip adds of CostRegNet (models/MVSNet/model.py:43-84), BasicBlock / UNet of Vis-MVSNet
 * (models/VisMVSNet/nn_utils.py:123-278) and CVP's CostRegNet (CVP_MVSNet/models/net.py:50-85).
 * --------------------------------------------------------------------------------------- */
#define MVSB200_SKIP_NONE 0
#define MVSB200_SKIP_BEFORE_RELU 1 /* relu(bn(conv) + skip)   BasicBlock, nn_utils.py:166-169        */
#define MVSB200_SKIP_AFTER_RELU 2  /* skip + relu(bn(deconv)) CostRegNet, MVSNet/model.py:79-81      */

typedef struct {
    int B, D, H, W;     /* INPUT spatial size */
    int Cin, Cin2;      /* channels of x and of the optional second input x2 (torch.cat([x, x2], 1)) */
    int Cout;
    int kd, kh, kw;     /* each 1 or 3; padding is k/2 */
    int stride;         /* 1 or 2 */
    int transposed;     /* 1: ConvTranspose3d(k=3, stride=2, padding=1, output_padding=1) */
    int relu;
    int skip_mode;
    int static_params;  /* 1: the caller guarantees that the layer's parameters (packed weights, scale, bias) were written
                         * before the PREVIOUS launch on the stream was issued -- they are not outputs of the preceding
                         * kernels.  The z-march engine is then launched as a programmatic dependent launch
                         * (cudaLaunchAttributeProgrammaticStreamSerialization): its prologue (barrier init, TMEM
                         * allocation, staging of the resident weights) overlaps the tail of the previous kernel and
                         * only the reads of x / x2 / skip / the abs-max scalars wait for it (griddepcontrol.wait).
                         * 0: ordinary stream-ordered launch. */
} mvsb200_conv3d_desc;

/* Output spatial size for a descriptor (PyTorch's rules). */
MVSB200_API int mvsb200_conv3d_out_shape(const mvsb200_conv3d_desc *desc, int *Do, int *Ho, int *Wo);

/* x [B,D,H,W,Cin]; x2 [B,D,H,W,Cin2] or NULL; w packed [kd*kh*kw][Cin+Cin2][Cout] (tap-major; for a
 * transposed conv the tap index is that of the ConvTranspose3d weight); scale,bias [Cout] or NULL
 * (y = conv*scale + bias); skip [B,Do,Ho,Wo,Cout] or NULL; y [B,Do,Ho,Wo,Cout]. */
MVSB200_API int mvsb200_conv3d(const mvsb200_conv3d_desc *desc, const float *x, const float *x2, const float *w,
                   const float *scale, const float *bias, const float *skip, float *y,
                   mvsb200_stream_t stream);

/* K2 on the tcgen05 tensor cores (k=3 layers with Cin % 8 == 0 and Cout in {1, 8, 16, 32, 64, ...}): implicit GEMM,
 * kind::tf32 MMAs with TMEM accumulators, same fused epilogue as mvsb200_conv3d.
 *   MVSB200_PRECISION_3XTF32  error-compensated split products, fp32-equivalent (the parity-tested default)
 *   MVSB200_PRECISION_TF32    single-pass TF32 (10-bit mantissa operands, fp32 accumulate); faster, outside the
 *                             1e-3 depth parity bar on peaked cost volumes
 * Weights are packed once per layer: `w` is the tap-major [27][Cin+Cin2][Cout] buffer mvsb200_conv3d takes,
 * `packed` a device buffer of mvsb200_conv3d_tc_packed_floats(desc) floats. */
#define MVSB200_PRECISION_3XTF32 0
#define MVSB200_PRECISION_TF32 1
MVSB200_API int mvsb200_conv3d_tc_supported(const mvsb200_conv3d_desc *desc);
MVSB200_API long long mvsb200_conv3d_tc_packed_floats(const mvsb200_conv3d_desc *desc);
MVSB200_API int mvsb200_conv3d_tc_pack(const mvsb200_conv3d_desc *desc, const float *w, float *packed,
                                       mvsb200_stream_t stream);
MVSB200_API int mvsb200_conv3d_tc(const mvsb200_conv3d_desc *desc, const float *x, const float *x2, const float *packed,
                                  const float *scale, const float *bias, const float *skip, float *y, int precision,
                                  mvsb200_stream_t stream);

/* K2 "z-march" engine (k=3 layers, Cin % 8 == 0, Cout == 8 or Cout % 16 == 0, packed weights resident in shared
 * memory): persistent warp-specialised tcgen05 kernel (kind::f16) that stages every input plane once and keeps the
 * accumulators of the output planes in flight in TMEM.  Operands are split into two fp16 pieces after a power-of-two
 * scaling taken from the tensor's abs-max (error-compensated products, fp32-equivalent to ~1e-6; the reference is
 * fp32: models/MVSNet/module.py:41-48, VisMVSNet/nn_utils.py:123-278, CVP_MVSNet/models/net.py:50-85).
 *   x_amax / x2_amax : device scalars holding max|x| (max|x2|) -- produced by mvsb200_absmax or by the y_amax output
 *                      of the kernel that wrote the tensor; they must bound the tensor (a too-small value overflows
 *                      fp16), a larger value only costs precision;
 *   y_amax           : optional device scalar, atomically max-ed with max|y| (zero it before the launch).
 * `packed` holds mvsb200_conv3d_zm_packed_bytes(desc) bytes (16-byte aligned) filled by mvsb200_conv3d_zm_pack from
 * the tap-major [27][Cin+Cin2][Cout] weights mvsb200_conv3d takes. */
MVSB200_API int mvsb200_absmax(const float *x, long long n, float *amax, mvsb200_stream_t stream);
MVSB200_API int mvsb200_conv3d_zm_supported(const mvsb200_conv3d_desc *desc);
MVSB200_API long long mvsb200_conv3d_zm_packed_bytes(const mvsb200_conv3d_desc *desc);
MVSB200_API int mvsb200_conv3d_zm_pack(const mvsb200_conv3d_desc *desc, const float *w, void *packed,
                                       mvsb200_stream_t stream);
MVSB200_API int mvsb200_conv3d_zm(const mvsb200_conv3d_desc *desc, const float *x, const float *x2, const void *packed,
                                  const float *scale, const float *bias, const float *skip, float *y,
                                  const float *x_amax, const float *x2_amax, float *y_amax, mvsb200_stream_t stream);
/* The same with x being channels [x_first_channel, x_first_channel + desc->Cin) of a tensor with x_channels channels
 * per voxel.  A layer whose packed weights do not fit the engine's shared memory (Cin * Cout >= 64 * 64) is run as two
 * launches over the two halves of its input channels, the second adding the first's output as its skip operand
 * (skip may alias y: every output element is read and written by the same thread). */
MVSB200_API int mvsb200_conv3d_zm_slice(const mvsb200_conv3d_desc *desc, const float *x, int x_channels, int x_first_channel,
                                        const float *x2, const void *packed, const float *scale, const float *bias,
                                        const float *skip, float *y, const float *x_amax, const float *x2_amax,
                                        float *y_amax, mvsb200_stream_t stream);

/* K2 single-output-channel head (3x3x3, stride 1, Cout == 1, Cin in {8,16,24,32}; MVSNet `prob`, Vis-MVSNet
 * `final_conv`, CVP `prob0`: models/MVSNet/model.py:72, VisMVSNet/model_cas.py:44,61, CVP_MVSNet/models/net.py:67) on the
 * CUDA cores in fp32: y = act(conv * scale + bias).  The weights are passed on the HOST (w_host: 27*Cin floats,
 * tap-major [27][Cin]) because they are handed to the kernel as launch parameters (constant bank); scale / bias are host
 * scalars.  x [B,D,H,W,Cin] device, y [B,D,H,W] device. */
MVSB200_API int mvsb200_conv3d_c1_supported(const mvsb200_conv3d_desc *desc);
MVSB200_API int mvsb200_conv3d_c1(const mvsb200_conv3d_desc *desc, const float *x, const float *w_host, float scale, float bias,
                                  float *y, mvsb200_stream_t stream);

/* Weight gradient of a 3x3x3 layer (training, row f2): R[ca][cb][27] += sum_v a[v,ca] * b[stride*v + tap - 1, cb], zero
 * outside the volume.  Conv3d(k=3,p=1,stride): a = grad_out [B,Da,Ha,Wa,Cout], b = input [B,Db,Hb,Wb,Cin] -> R = dW
 * [Cout,Cin,3,3,3]; ConvTranspose3d(k=3,p=1,stride,output_padding=stride-1): a = input, b = grad_out -> R = dW
 * [Cin,Cout,3,3,3] (autograd of models/MVSNet/module.py:41-48, MVSNet/model.py:59-72, VisMVSNet/nn_utils.py:123-278,
 * CVP net.py:50-74).  R must be ZEROED by the caller (blocks add partial sums).  The input gradient of the same layers is
 * a forward call of mvsb200_conv3d_zm / mvsb200_conv3d with re-packed weights (no symbol of its own). */
MVSB200_API int mvsb200_conv3d_wgrad(const float *a, const float *b, int B, int Da, int Ha, int Wa, int Ca, int Db, int Hb, int Wb,
                                     int Cb, int stride, float *r, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K3: softmax over D + depth regression + confidence (+ entropy, + probability volume).
 * Replaces F.softmax + depth_regression (+ photometric confidence): models/MVSNet/model.py:207-215,
 * module.py:174-182; soft_argmin / entropy: models/VisMVSNet/nn_utils.py:453-470; CVP net.py:161-162,
 * 203-219.
 * --------------------------------------------------------------------------------------- */
#define MVSB200_CONF_NONE 0
#define MVSB200_CONF_SUM4 1   /* sum of p over [idx-1, idx+2], idx = trunc(E[d])  (MVSNet/model.py:211-215) */
#define MVSB200_CONF_WINDOW 2 /* sum of p over |d - E[d]| <= 2                    (nn_utils.py:463-465)     */

/* score [B,D,H,W]; depth/interval as in K1 (depth_mode); outputs [B,H,W] (NULL to skip);
 * prob_out [B,D,H,W] or NULL. */
MVSB200_API int mvsb200_depth_regress(const float *score, int B, int D, int H, int W, int depth_mode, const float *depth,
                          const float *interval, int conf_mode, float *depth_out, float *conf_out,
                          float *entropy_out, float *prob_out, mvsb200_stream_t stream);

/* Backward of the regression output: grad_score[b,d,y,x] = grad_depth[b,y,x] p_d (h_d - depth), p = softmax(score)
 * (autograd of models/MVSNet/model.py:207-209, module.py:174-178, VisMVSNet/nn_utils.py:453-466; confidence and entropy
 * are produced under no_grad in the reference).  grad_depth [B,H,W]; grad_score [B,D,H,W], fully written.
 * grad_hyp (optional, DEPTH_VOLUME only): [B,D,H,W] <- grad_depth p_d, the gradient with respect to per-voxel hypotheses
 * (CVP-MVSNet's refinement levels regress over hypotheses built from the previous level's depth map:
 * CVP_MVSNet/models/net.py:176-182,203, modules.py:362-365). */
MVSB200_API int mvsb200_depth_regress_backward(const float *score, int B, int D, int H, int W, int depth_mode, const float *depth,
                                               const float *interval, const float *grad_depth, float *grad_score,
                                               float *grad_hyp, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K5: the per-pixel part of CVP-MVSNet's calDepthHypo (models/CVP_MVSNet/models/modules.py:131-226), fp64 inside:
 * abs_delta[b][y*W+x] = |depth change that moves the projection of pixel (x,y) at depth ref_depth[b][y][x] into the
 * first source view by one pixel along the epipolar line|, +inf where the reference's validity test fails
 * (modules.py:206-209).  The level's hypothesis interval is the (lower) median of the finite entries.
 * ref_depth [B,H,W] fp32; ref_in, src_in [B,3,3] (level-conditioned intrinsics, first source view); ref_ex, src_ex
 * [B,4,4]; abs_delta [B,H*W] fp64. */
MVSB200_API int mvsb200_cvp_depth_delta(const float *ref_depth, const float *ref_in, const float *src_in, const float *ref_ex,
                                        const float *src_ex, int B, int H, int W, double *abs_delta, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K4: visibility-weighted fusion of per-pair volumes (models/VisMVSNet/model_cas.py:354-357,385-386):
 *   fused = sum_s exp(-uncert_s) * interm_s / sum_s exp(-uncert_s)
 * interm, uncert: HOST arrays of S device pointers; interm[s] [B,D,H,W,G], uncert[s] [B,H,W]. */
MVSB200_API int mvsb200_vis_fuse(const float *const *interm, const float *const *uncert, int S, int B, int D, int H, int W,
                     int G, float *fused, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K6: UncertNet of Vis-MVSNet (models/VisMVSNet/model_cas.py:77-98) fused into one kernel:
 *   u = head( relu(bn2(conv2( relu(bn1(conv1(e))) ))) + e ),   e = per-pair entropy map, all convs 3x3, pad 1.
 * entropy, out: [N,H,W] device (N = pairs x batch).  params_host: HOST array of 752 floats
 *   w1[9][8] (tap-major, tap = ky*3+kx), scale1[8], bias1[8], w2[9][8 in][8 out], scale2[8], bias2[8], w_head[9][8]
 * (eval-mode BatchNorm folded into scale/bias); they become launch parameters. */
#define MVSB200_UNCERT_PARAMS 752
MVSB200_API int mvsb200_vis_uncert_net(const float *entropy, int N, int H, int W, const float *params_host, float *out,
                                       mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K7: 2-D convolution + folded BatchNorm + ReLU over channels-last maps (fp32, CUDA cores): ConvBnReLU of
 * models/MVSNet/module.py:9-17 as used by FeatureNet (models/MVSNet/model.py:21-41), row f1 of SURVEY.md 8.
 *   y = act(conv(x, w) * scale + bias),  padding k/2;  (k, stride) in {(3,1), (5,2)};  Cin in {4,8,16,32} (a 3-channel
 *   image is passed zero-padded to 4 channels), Cout in {8,16,32}.
 * x [B,H,W,Cin], w [k*k][Cin][Cout] (tap-major, tap = ky*k+kx), scale/bias [Cout] or NULL, y [B,Ho,Wo,Cout]. */
MVSB200_API int mvsb200_conv2d(int B, int H, int W, int Cin, int Cout, int k, int stride, int relu, const float *x, const float *w,
                               const float *scale, const float *bias, float *y, mvsb200_stream_t stream);

/* K7b: y = act(y * scale[c] + bias[c] (+ residual)) in place over a channels-last map y [n_pixels, C] (C % 4 == 0, 16-byte
 * aligned): the fused epilogue behind a convolution library call that has none -- CVP-MVSNet's `conv` = Conv2d(bias) +
 * LeakyReLU(0.1) (models/CVP_MVSNet/models/modules.py:24-28; slope 0.1), eval-mode BatchNorm + ReLU (+ skip) of the 2-D
 * BasicBlocks (slope 0).  scale, bias, residual may be NULL (not both bias and residual); slope = 1 is no activation -- with
 * only a residual it is the skip connection added AFTER the activation of a layer that ran as two launches.
 * amax (optional, a ZEROED device scalar) receives max|y| (the z-march engine's operand scale for the next layer). */
MVSB200_API int mvsb200_bias_act(float *y, long long n_pixels, int C, const float *scale, const float *bias, const float *residual,
                                 float slope, float *amax, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K8: geometric-consistency filter of one depth map against its N source depth maps (evaluation/filtering.py:59-84
 * with utils/utils_3D.py:64-73,98-132,162-178,243-273,296-311), row f3 of SURVEY.md 8.
 * depth [H,W]; src_depth: HOST array of N device pointers, map i is [src_h[i], src_w[i]] (src_h, src_w: host arrays);
 * K, R [1+N,3,3], t [1+N,3] (view 0 = the reference view, x_cam = R X + t, intrinsics already divided by the down-scale
 * factor); thresholds as the reference's --depth_threshold / --max_reproj_error / --min_tri_angle / --num_consistent.
 * Outputs (bytes, 0/1) [H,W]: mask_depth, mask_disp, geo_mask = at least num_consistent-1 sources agree;
 * votes [3,H,W] (optional): the number of agreeing sources behind each mask. */
MVSB200_API int mvsb200_geometric_filter(const float *depth, int H, int W, const float *const *src_depth, const int *src_h,
                                         const int *src_w, int N, const float *K, const float *R, const float *t,
                                         float depth_threshold, float max_reproj_error, float min_tri_angle, int num_consistent,
                                         unsigned char *mask_depth, unsigned char *mask_disp, unsigned char *geo_mask,
                                         unsigned char *votes, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * K9: the consumer of the all-gathered depth maps, row f3 of SURVEY.md 8 -- the geometric half of
 * masked_photometricloss (models/trainer.py:240-278) with get_flow_from_depthmap (:209-219),
 * flows_from_single_depthmap (utils/utils_3D.py:190-211) and normalize (:243-268): per pixel of this rank's depth map
 * and source view, re-project with the 4x4 projection matrices, sample the GATHERED depth map of that view
 * (grid_sample bilinear / zeros / align_corners=False) and keep the pixel where the relative depth difference is below
 * geom_clamping and the sample falls inside the image.
 * ref_depth [B,H,W]; gathered [B,N,H,W] (the all-gather's output; view `ref` is skipped); proj [B,N,4,4];
 * inv_ref [B,4,4] = proj[:, ref]^-1; imgs [B,N,C,H,W] or NULL.  Sources are ordered as the reference's src_idx (all
 * views but `ref`, ascending).  Outputs over [B,N-1,H,W]: mask (bytes 0/1, required); optional inside (bytes),
 * grid [..,2] (the normalised sampling grid), depth_src, warped_depth, warped [B,N-1,C,H,W] (the images sampled with the
 * same grid: what the SSIM term compares with the reference image). */
MVSB200_API int mvsb200_gathered_masks(int B, int N, int C, int H, int W, int ref, const float *ref_depth, const float *gathered,
                                       const float *proj, const float *inv_ref, const float *imgs, float geom_clamping,
                                       unsigned char *mask, unsigned char *inside, float *grid, float *depth_src,
                                       float *warped_depth, float *warped, mvsb200_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU: independent reference views per rank, ONE all-gather of the per-view depth maps -- the
 * collective the reference issues with `dist.all_gather(all_depthmaps, depth_est)` (models/trainer.py:246-247) on views
 * sharded as `reference_frame = rank` (:101) / DistributedSampler (depthmap_eval.py:95).
 * NCCL is bound at run time (dlopen("libnccl.so.2")); the library has no link-time dependency on it.
 *   mvsb200_gather_unique_id   rank 0: fills a 128-byte HOST buffer (ncclUniqueId); hand it to every rank by any means
 *   mvsb200_gather_init        every rank, current device: creates the communicator (*comm, opaque)
 *   mvsb200_allgather_depth    IN PLACE over all_maps [world * count_per_rank] floats (device): rank r owns the slice
 *                              [r * count_per_rank, (r + 1) * count_per_rank) -- pass that slice as `depth` to
 *                              mvsb200_depth_regress so the map is born where it is sent from; stream-ordered, no
 *                              allocation, legal inside a CUDA-graph capture (destroy captured graphs before the communicator)
 *   mvsb200_gather_destroy */
MVSB200_API int mvsb200_gather_unique_id(void *id128);
MVSB200_API int mvsb200_gather_init(const void *id128, int world, int rank, void **comm);
MVSB200_API int mvsb200_allgather_depth(void *comm, float *all_maps, long long count_per_rank, int rank, mvsb200_stream_t stream);
MVSB200_API int mvsb200_gather_destroy(void *comm);

#ifdef __cplusplus
}
#endif
#endif /* MVSB200_H_ */
