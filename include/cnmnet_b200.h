/*
 * cnmnet_b200 -- C ABI of the B200-native CNMNet inference hot path (libcnmnet_b200.so).
 *
 * The reference (xxlong0/CNMNet) has no FFI layer: its boundary is the torch.nn.Module API
 * of depthnet/ (SURVEY.md section 8b).  This header is the layer inserted UNDER that API:
 * every entry point replaces one chain of PyTorch library ops in the reference, cited
 * per function as <file>:<lines> relative to the reference tree.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers on the current
 *     device unless the name says `host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - returns 0 on success, a negative cnm_status otherwise; never throws, never allocates
 *     device memory, never synchronises (CUDA-graph capturable); cnm_last_error() gives
 *     a thread-local message for the last failing call;
 *   - inputs are never modified; outputs are fully overwritten.
 *
 * Public-boundary tensors are contiguous NCHW fp32 exactly as in the reference.  Network
 * activations between kernels use the private "act" layout: NHWC, channel count a multiple
 * of the K-chunk (64 for CNM_BF16, 32 for CNM_TF32), element type per cnm_dtype.
 */
#ifndef CNMNET_B200_H_
#define CNMNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cnm_status {
  CNM_OK = 0,
  CNM_ERR_INVALID_ARGUMENT = -1,
  CNM_ERR_UNSUPPORTED_SHAPE = -2,
  CNM_ERR_CUDA = -3,
  CNM_ERR_NO_DEVICE = -4
} cnm_status;

/* element type of private activation tensors and packed conv weights */
typedef enum cnm_dtype {
  CNM_BF16 = 0, /* bf16 storage, tcgen05 kind::f16, fp32 accumulate; K-chunk = 64 channels */
  CNM_TF32 = 1, /* fp32 storage, tcgen05 kind::tf32, fp32 accumulate; K-chunk = 32 channels */
  CNM_BF16X3 = 2, /* split bf16: every value stored as hi + lo bf16 (16 mantissa bits); a convolution runs as three
                    kind::f16 passes hi*Whi + lo*Whi + hi*Wlo.  Physical layout: logical channel c of a pixel at
                    physical channel (c/64)*128 + c%64 (hi) and 64 further (lo); act tensors have 2*C physical
                    channels.  Conv descriptors / sources / steps are in PHYSICAL channels; desc.Cout is logical. */
  CNM_F16F8 = 3   /* fp16 + fp8 corrections: a value v is stored as hi = fp16(v) plus two e4m3 bytes,
                    lo8 = e4m3((v - hi) * 2^7) and hi8 = e4m3(hi * 2^-4).  Same physical layout as CNM_BF16X3 in
                    2-byte units: per 64-channel logical chunk a 128-byte fp16 row, then a 128-byte fp8 row
                    [lo8 x 64 | hi8 x 64].  A convolution runs as TWO passes into one accumulator: kind::f16
                    hi * W16 and kind::f8f6f4 [lo8 | hi8] * [Whi8 | Wlo8] (K = 128 fp8 elements per 128-byte row,
                    twice the bf16 rate), with the power-of-two scales folded into the packed weights:
                    W16 = fp16(w * 2^e), Whi8 = e4m3(w * 2^(e-7)), Wlo8 = e4m3((w * 2^e - W16) * 2^4); the
                    accumulator is multiplied by 2^-e (desc.acc_exp = e) before bias / ReLU.  The FIRST half
                    of the K-steps (groups) are the fp8 rows, the second half the fp16 rows: the small correction terms
                    are accumulated while the sum is tiny, which halves the rounding of the tensor core's accumulator.  ~22 significand bits for the leading term, corrections good to
                    ~2^-17 relative: same accuracy class as CNM_BF16X3 at 2/3 of its tensor time. */
} cnm_dtype;

/* cost-volume output layouts */
typedef enum cnm_volume_layout {
  CNM_VOLUME_NCHW_F32 = 0, /* [P, D, H, W] fp32: the reference's getVolume() result */
  CNM_VOLUME_ACT = 1       /* [P, H, W, D] in cnm_dtype: conv1.0's operand, no torch.cat */
} cnm_volume_layout;

const char* cnm_last_error(void);
/* library / build identification, e.g. "cnmnet_b200 0.1 sm_100a" */
const char* cnm_version(void);
/* number of kernels launched by this library since load (all threads); used for gpu_launches */
int64_t cnm_launch_count(void);

/* Debug range check of CNM_F16F8 (activations are stored as fp16 and saturate at 65504 silently): when enabled -- by
 * cnm_set_range_check(1) or CNM_CHECK_RANGE=1 in the environment at first use -- every convolution epilogue counts the
 * (pixel, 64-channel chunk) pairs holding a value beyond the fp16 range in a per-device counter.
 * cnm_range_overflow_count synchronises the device and returns the count (-1 when the check is off); reset != 0
 * clears it.  Enable before the first step of a pipeline (the counter is allocated on first use, outside graphs). */
void cnm_set_range_check(int enable);

/* Programmatic dependent launch between the kernels of a pipeline step (on by default; CNM_PDL=0 in the environment
 * or cnm_set_pdl(0) turns it off): every step kernel is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and waits (griddepcontrol.wait) before its first global-memory
 * access, so its prologue overlaps the tail of the kernel before it.  Results do not depend on the setting.
 * Set it before a plan is captured into a CUDA graph. */
void cnm_set_pdl(int enable);
int64_t cnm_range_overflow_count(int reset);

/* ---- camera pre-processing -------------------------------------------------------------
 * replaces process_camera_parameters + get_pixel_coordinates  (depthnet/depth_util.py:13-56)
 * ref_cam [B,2,4,4], src_cam [V,B,2,4,4]  ->  hk [V*B, 12] = (Hm row-major 3x3, kt 3)
 * with Hm = K_s R K_r^-1, kt = K_s t, [R|t] = E_s E_r^-1; computed in fp64, stored fp32.  */
int cnm_camera_setup(const float* ref_cam, const float* src_cam, float* hk, int B, int V, void* stream);

/* ---- plane-sweep cost volume -----------------------------------------------------------
 * replaces depthNet.getVolume incl. F.grid_sample  (depthnet/depthNet_model.py:185-224)
 * cnm_pack_source: src [N,3,H,W] fp32 -> packed texels [N, H+3, W+4] x float4 (RGB0) with a
 *   zero border (row/col -1 and the trailing ones) so zero padding needs no predicates.
 * cnm_cost_volume: ref [B,3,H,W]; packed sources for pair p = v*B + b; hk [V*B,12];
 *   out per `layout`; D planes, z_d = 1/(idmin + d (idmax-idmin)/(D-1)) evaluated in fp64.  */
int64_t cnm_packed_source_bytes(int N, int H, int W);
int cnm_pack_source(const float* src, void* packed, int N, int H, int W, void* stream);
int cnm_cost_volume(const float* ref, const void* packed_src, const float* hk, void* out, int B, int V, int H,
                    int W, int D, double idepth_min, double idepth_max, int layout, int dtype, void* stream);

/* cnm_plane_sweep: the same cost volume, second-generation kernel (csrc/plane_sweep.cu).  Reads the source views IN
 *   PLACE: src [V*B,3,H,W] fp32 planar, pair p = v*B + b -- no cnm_pack_source pass.  Per (32x8 pixel tile, 8 planes)
 *   the source footprint is staged in shared memory by TMA (zero fill outside the image = grid_sample's zeros
 *   padding) and every bilinear tap is a shared-memory load; footprints that do not fit a 64 x 40 box are sampled from
 *   global memory.  Same arguments and outputs otherwise.  Needs W % 4 == 0, D % 8 == 0, D <= 128, src 16B aligned
 *   (CNM_ERR_UNSUPPORTED_SHAPE otherwise: use the cnm_pack_source + cnm_cost_volume pair).                        */
int cnm_plane_sweep(const float* ref, const float* src, const float* hk, void* out, int B, int V, int H, int W, int D,
                    double idepth_min, double idepth_max, int layout, int dtype, void* stream);

/* ---- depth -> normal -------------------------------------------------------------------
 * replaces pixel2cam (depthnet/inverse_warp.py:27-43) + Depth2normal.forward without planes
 * (depthnet/depth_util.py:149-203,240).  depth [B,H,W], Kinv [B,3,3] -> normals [B,3,H,W],
 * points [B,3,H,W] (may be NULL).  k odd, 3..15.  Box sums and the 3x3 solve run in fp64.   */
int cnm_depth2normal(const float* depth, const float* Kinv, float* normals, float* points, int B, int H, int W,
                     int k_size, void* stream);

/* replaces Kinv = inverse(ref_cam[1,:3,:3])  (eval.py:270-271, 453-454): cam [B,2,4,4] -> Kinv [B,3,3] */
int cnm_intrinsics_inverse(const float* cam, float* Kinv, int B, void* stream);

/* ---- Combined Normal Map ---------------------------------------------------------------
 * replaces the planes branch of Depth2normal.forward (depthnet/depth_util.py:205-238) and
 * get_normal_by_planes (:243-278).  normals [B,3,H,W] updated IN PLACE (the one exception to
 * the no-mutation rule: callers pass a private copy); segs [B,P,H,W] uint8; planes_num [B]
 * int32 on device; loss (device double, may be NULL) receives sum_b sum_i mean_hw(1-cos).
 * Planes are applied with the reference's sequential semantics (overlapping masks included) but
 * computed in parallel over pixels; P <= 32.  workspace: cnm_plane_regularize_workspace_bytes()
 * bytes of device memory, 256B aligned, contents irrelevant.                                  */
int64_t cnm_plane_regularize_workspace_bytes(int B, int P, int H, int W);
int cnm_plane_regularize(float* normals, const uint8_t* segs, const int32_t* planes_num, double* loss, void* workspace,
                         int B, int P, int H, int W, void* stream);

/* ---- depth evaluation metrics ------------------------------------------------------------
 * replaces the per-image body of cal_metrics (eval.py:1030-1047) and the functions it calls in
 * utils/metric.py:149-362.  pred, gt [B,H,W] fp32 (device); pred is clamped to [min_depth,
 * max_depth], valid pixels are min_depth < gt < max_depth and finite.  out [B,10] double =
 * {num_valid, l1, abs_relative, rmse, scale_invariant, sq_relative, rmse_log, a1, a2, a3}
 * (NaN metrics for an image without valid pixels, like the reference).  workspace:
 * cnm_depth_metrics_workspace_bytes() bytes, 8B aligned.                                    */
int64_t cnm_depth_metrics_workspace_bytes(int B, int H, int W);
int cnm_depth_metrics(const float* pred, const float* gt, double* out, void* workspace, int B, int H, int W,
                      float min_depth, float max_depth, void* stream);

/* ---- training losses, forward values --------------------------------------------------
 * replace depthnet/losses.py: IdepthLoss (:31-48) / IdepthwithProbLoss (:51-74) -> cnm_idepth_loss (prob NULL = plain
 * L1; log_mode: |log10 pred - log10 gt|, times 10 when prob is given); IdepthLoss_234 (:7-28) ->
 * cnm_idepth_loss_234 (gt [B,1,H,W], disp_k [B,1,H/2^(k-1),W/2^(k-1)]); surface_normal_loss (:77-122) ->
 * cnm_surface_normal_loss (valid [B,1,H,W] uint8 or NULL, prob [B,1,H,W] or NULL).  out: device doubles --
 * idepth {loss, masked pixel count}, 234 {loss}, normals {loss, mean angle in degrees, pixel count}.
 * workspace: cnm_loss_workspace_bytes() bytes, 8B aligned.  No gradients: the package is inference-only.        */
int64_t cnm_loss_workspace_bytes(void);
int cnm_idepth_loss(const float* pred, const float* gt, const float* prob, double* out, void* workspace, int64_t n,
                    int log_mode, void* stream);
int cnm_idepth_loss_234(const float* disp2, const float* disp3, const float* disp4, const float* gt, double* out,
                        void* workspace, int B, int H, int W, void* stream);
int cnm_surface_normal_loss(const float* pred, const float* gt_normal, const uint8_t* valid, const float* prob, double* out,
                            void* workspace, int B, int H, int W, void* stream);

/* ---- idepth -> depth -------------------------------------------------------------------
 * replaces torch.div(1.0, idepth)  (eval.py:269,452,934).  n elements.                      */
int cnm_reciprocal(const float* idepth, float* depth, int64_t n, void* stream);

/* ---- layout converters between the public NCHW fp32 boundary and private act tensors ---- */
int cnm_nchw_to_act(const float* src, void* dst, int N, int C, int H, int W, int c_dst_total, int c_dst_offset,
                    int dtype, void* stream);
int cnm_act_to_nchw(const void* src, float* dst, int N, int C, int H, int W, int c_src_total, int c_src_offset,
                    int dtype, void* stream);

/* ---- tap packers (side operands of the implicit GEMM) ------------------------------------
 * cnm_pack_row_taps: x [Nsrc,Cx,H,W] fp32 -> act [N,H,W,chunk]; image n reads x[n % Nsrc] (one ref image
 *   serves all its source views); channel s*Cx+c holds x[c][y][x+s-kw/2] (zero outside the image),
 *   for s < kw; used for the 3 image channels of conv1.0
 *   (torch.cat((left_image, volume),1), depthNet_model.py:233-235).
 * cnm_pack_disp_taps: disp [N,1,h,w] fp32 -> act [N,2h,2w,chunk]; channel 3r+s holds
 *   nearest_up2(disp)[y+r-1][x+s-1] (zero outside): F.upsample + cat (depthNet_model.py:247-258). */
/* cnm_pack_row_taps3: the denser form for the 2-byte dtypes -- dst [N, H+2, W, 64]: row j (j = 0 .. H+1, one apron row
 * above and below the image) is centred on image row j-1 and channel (q*kw + s)*Cx + c holds x[c][j-2+q][xx+s-kw/2],
 * q = 0..2 (zero outside the image): three kernel rows per K-step, a 7x7 layer needs 3 steps instead of 7.           */
int cnm_pack_row_taps3(const float* x, void* dst, int N, int Nsrc, int Cx, int H, int W, int kw, int dtype, void* stream);
int cnm_pack_row_taps(const float* x, void* dst, int N, int Nsrc, int Cx, int H, int W, int kw, int dtype,
                      void* stream);
int cnm_pack_disp_taps(const float* disp, void* dst, int N, int h, int w, int dtype, void* stream);

/* ---- refine-input aggregation ----------------------------------------------------------
 * replaces x = cat(id1, id2, |id1-id2|, iconv1+iconv2) (depthnet/depthNet_model.py:332-333) and
 * the view-group means that precede it (eval.py:659-663, 920-929).
 * idepth [G*2... see below]: the V = 2*group views are laid out view-major: idepth [V,B,H,W]
 * fp32, iconv act [V*B,H,W,64]; group g in {0,1} averages views g, g+2, g+4, ... (op order
 * ((x0+x2)+x4)/group).  Outputs: sum_act [B,H,W,64] = mean_0(iconv) + mean_1(iconv);
 * taps_act [B,H,W,chunk] channel 9c+3r+s = 3x3 taps of c in (id1,id2,|id1-id2|);
 * id_out [2,B,H,W] fp32 group-mean idepths (may be NULL).                                   */
int cnm_refine_aggregate(const float* idepth, const void* iconv, void* sum_act, void* taps_act, float* id_out,
                         int B, int H, int W, int group, int dtype, void* stream);
/* same, but iconv given at the public boundary as NCHW fp32 [2,B,64,H,W] (group = 1) */
int cnm_refine_aggregate_nchw(const float* idepth, const float* iconv_nchw, void* sum_act, void* taps_act, int B,
                              int H, int W, int dtype, void* stream);

/* ---- bilinear x2 upsample of an act tensor ----------------------------------------------
 * replaces nn.Upsample(scale_factor=2, mode='bilinear') (depthnet/depthNet_model.py:94,105)
 * [N,h,w,C] -> [N,2h,2w,C], align_corners=False.                                            */
int cnm_upsample2x(const void* src, void* dst, int N, int h, int w, int C, int dtype, void* stream);

/* ---- 1-channel 3x3 head ----------------------------------------------------------------
 * replaces depth_layer = conv3x3(C->1, bias) + sigmoid, times idepth_scale
 * (depthnet/depthNet_model.py:82-84,245,250,255,261,351,365).  x act [N,H,W,C]; w [9*C] fp32
 * ordered (r,s,c); out [N,1,H,W] fp32 = scale * sigmoid(conv + bias); logits (may be NULL).  */
int cnm_head3x3(const void* x, const float* w, float bias, float scale, float* out, float* logits, int N, int H,
                int W, int C, int dtype, void* stream);

/* The same head in two passes, the form the pipeline uses: cnm_head_partial reads every pixel once and writes the nine
 * per-tap dot products of its C-vector with w_tap_major [9][C] (fp32, tap t = 3r+s) as 12 floats per pixel
 * (partial [N,H,W,12], entries 9..11 zero); 64-channel layers get the same partials for free from the producing
 * convolution's epilogue (cnm_conv_desc.head_partial).  cnm_head_finish gathers the 3x3 neighbourhood.            */
int cnm_head_partial(const void* x, const float* w_tap_major, float* partial, int N, int H, int W, int C, int dtype,
                     void* stream);
/* second half of a head: out(y,x) = scale * sigmoid(bias + sum over the 3x3 neighbourhood of
 * sum_k partial[k][y+r-1][x+s-1][3r+s]), zero outside the image; partial = nslices stacked [N,H,W,12] tables (1 from
 * cnm_head_partial, Cout/64 from a fused head); logits (may be NULL) gets the pre-sigmoid sum                         */
int cnm_head_finish(const float* partial, int nslices, float bias, float scale, float* out, float* logits, int N, int H,
                    int W, void* stream);

/* ---- convolution as implicit GEMM --------------------------------------------------------
 * replaces nn.Conv2d + BatchNorm2d(eval) + ReLU and the torch.cat of its inputs
 * (depthnet/depthNet_model.py:19-112, 235-261, 335-365).
 *
 * The K loop is a list of steps; step i multiplies a [128 pixel x chunk] tile of source
 * tensor `src` shifted by (dx,dy) pixels, channels [c0, c0+chunk), with columns
 * [i*chunk, (i+1)*chunk) of the packed weight matrix wp [Cout, num_steps*chunk] (K-major,
 * BN folded).  Several sources = concatenation; taps = spatial kernel; stride-2 layers list
 * their taps against the four parity sub-images of the source (sub = py*2+px).
 */
typedef struct cnm_conv_src {
  const void* ptr; /* act tensor [N, H, W, C]; for a view: its first pixel */
  int32_t H, W, C;
  int32_t apron;   /* stride-1 layers: the tensor has `apron` extra rows above and below the image (H = Ho + 2 apron; row r
                      is image row r - apron); steps and groups address TENSOR rows.  0 for ordinary sources. */
  /* View (fold layers of cnm_conv2d_halo / _halo2 only; all four 0 = the whole contiguous tensor): the source is the
   * window rows [y_off, y_off + H) x columns [x_off, x_off + W) of images that are pitch_w pixels wide and pitch_hw pixels
   * apart; every pixel outside the window reads as zero (TMA bounds).  This is how a border row / column / corner of a
   * tensor becomes an operand of its own without being copied.                                                      */
  int32_t x_off, y_off, pitch_w, pitch_hw;
} cnm_conv_src;

typedef struct cnm_conv_step {
  int32_t src; /* index into srcs */
  int32_t sub; /* 0 for stride 1; parity sub-image py*2+px for stride 2 */
  int32_t dx, dy; /* tile shift in (sub-)image pixels */
  int32_t c0;  /* first channel */
  int32_t reserved[3];
} cnm_conv_step;

typedef struct cnm_conv_desc {
  int32_t N, Ho, Wo, Cout; /* output act tensor [N, Ho, Wo, c_out_total], written at c_out_offset */
  int32_t c_out_total, c_out_offset;
  int32_t stride; /* 1 or 2 */
  int32_t relu;
  int32_t num_srcs, num_steps;
  int32_t dtype; /* cnm_dtype */
  int32_t reserved; /* 0 = library picks the output-channel tile; 64 / 128 / 256 = caller-selected (autotuner) */
  int32_t acc_exp;  /* the accumulator is scaled by 2^-acc_exp before bias / ReLU (CNM_F16F8 weight scale; else 0) */
  int32_t ksplit;   /* cnm_conv2d_tc only: > 1 cuts the K loop of every output tile into ksplit ranges run by different
                       CTAs (layers with fewer tiles than SMs); partial accumulators go through splitk_ws and a second
                       kernel sums them in a fixed order.  0 / 1 = off. */
  int32_t cluster;  /* cnm_conv2d_tc only, 2-byte dtypes, column tile >= 128: 2 or 4 = thread-block clusters of that many
                       CTAs share every weight tile through TMA multicast (each CTA loads 1/cluster of it).  Same
                       results as 0 / 1 (off): the K order of a tile does not change. */
  int32_t tile_h;   /* cnm_conv2d_halo2 only, 2-byte dtypes: 8 = 16x8 pixel blocks per CTA instead of 16x16 (maps with few
                       blocks: twice the work items, two accumulator stages at a column tile of 256).  0 / 16 = 16x16. */
  int32_t fold;     /* cnm_conv2d_halo / _halo2, 2-byte dtypes: 1 = the layer is nn.Upsample(x2, bilinear) + conv3x3 folded into
                       ONE convolution on the low-res tensor (depthnet/depthNet_model.py:92-98, `upconv`).  Ho x Wo is the
                       LOW-RES size, Cout = 4 x C channels = one block of C per output phase (a, b), and the epilogue writes
                       block (a, b) of low-res pixel (y, x) to out[n][2y+a][2x+b] of the [N, 2Ho, 2Wo, c_out_total] output.
                       The caller supplies the phase-combined weights (bilinear taps 0.75 / 0.25 folded into the 3x3
                       kernel) and, as extra sources, one-row / one-column / one-pixel VIEWS of the input with the border
                       corrections (where the interpolation clamps and the convolution pads with zeros, the operator
                       differs from the interior one); groups carry a border class in bits 4-7 of ntaps and run only on
                       the tiles of that class.  Up to 12 sources, operand map = source index.  See engine.FoldedUpconvOp. */
  int32_t balance;  /* 1 = balanced (stream-K) schedule, 2-byte dtypes: the launch's K-steps (tiles x steps) are cut into one
                       equal contiguous range per SM instead of whole tiles per CTA, so layers whose tile count is not a
                       multiple of the SM count (120 or 240 tiles on 148 SMs) keep every SM busy.  A CTA whose range starts
                       inside a tile parks that piece's raw fp32 accumulators in splitk_ws; the CTA holding the tile's first
                       K-step adds the pieces in a fixed order in its epilogue (deterministic; last bits differ from the
                       one-chain schedule).  cnm_conv2d_tc (ksplit <= 1, cluster <= 1) and cnm_conv2d_halo / _halo2 (column
                       tile <= 128, 16x16 blocks).  splitk_ws: cnm_conv_balance_ws_bytes() bytes, 128B aligned, ZEROED once
                       by the caller (the kernels leave it zeroed); launches sharing it must be stream-ordered. */
  /* Optional fused 1-channel 3x3 head on this layer's output (depth_layer, depthnet/depthNet_model.py:82-84), for the
   * 2-byte dtypes (tensor-core kernels only): head_w = [9][Cout] fp32 tap-major weights (device, 16B aligned),
   * head_partial = [Cout/64][N, Ho, Wo, 12] fp32 (device, 16B aligned): every 64-channel chunk of the output writes,
   * per pixel, the nine per-tap dot products of its part of the fp32 output vector into its own slice (entries 9..11
   * are zero).  cnm_head_finish(nslices = Cout/64) turns them into the head's output.  Both NULL = no fused head.   */
  const float* head_w;
  float* head_partial;
  float* splitk_ws; /* [ksplit][N*Ho*Wo][Cout] fp32 workspace (device, 16B aligned) when ksplit > 1, else NULL */
  /* Optional 1-channel side input added in the epilogue: the nearest-upsampled disparity that depthNet_model.py:247-258
   * concatenates to the inputs of iconv3 / iconv2 / iconv1.  side_disp = [N, 1, Ho/2, Wo/2] fp32 (Ho, Wo even), side_w =
   * [9][Cout] fp32 tap-major = the BN-folded 3x3 weights of that input channel (16B aligned).  Before bias-ReLU the layer
   * adds sum_{r,s} side_w[3r+s][c] * up(side_disp)(y+r-1, x+s-1), up = nearest x2, zero outside the image -- in fp32, on
   * unrounded disparities, instead of spending a 64-channel K-step (and a 64-channel side tensor) on nine scalars per
   * pixel.  Tensor-core kernels, 2-byte dtypes, ksplit <= 1.  Both NULL = none.                                       */
  const float* side_disp;
  const float* side_w;
} cnm_conv_desc;

int64_t cnm_conv_balance_ws_bytes(void);

/* tcgen05/TMEM/TMA implicit GEMM (the product path).  `steps` is a DEVICE array of
 * num_steps entries (uploaded once when the layer plan is built); wp/bias/out are device.   */
int cnm_conv2d_tc(const cnm_conv_desc* desc, const cnm_conv_src* srcs, const cnm_conv_step* steps, const void* wp,
                  const float* bias, void* out, void* stream);
/* same contract on CUDA cores with fp32 FMA: validation twin of the tensor-core kernel        */
int cnm_conv2d_simt(const cnm_conv_desc* desc, const cnm_conv_src* srcs, const cnm_conv_step* steps, const void* wp,
                    const float* bias, void* out, void* stream);

/* Row-halo variant for layers with Ho, Wo >= 16 (csrc/conv_halo.cu): same act / weight layouts and descriptor; the
 * K loop is given as GROUPS.  A group loads one slab {chunk, 16, rows} of operand map `map` = src*4 + sub (sub = 0
 * for stride 1, the parity sub-image py*2+px for stride 2): channels [c0, c0+chunk), columns shifted by dx, rows
 * starting `pad` above the tile; it consumes `ntaps` consecutive weight steps, tap t reading the slab rows shifted
 * by t.  desc->num_steps is the total number of weight steps (sum of ntaps); map_info[2*map] = pad and
 * map_info[2*map+1] = rows (16 + vertical tap span; 0 = map unused) is a HOST array of 12 x 2 ints.             */
typedef struct cnm_conv_group {
  int32_t map, c0, dx;
  int32_t ntaps; /* bits 0-3: taps of the group; bits 4-7: border class the group needs (fold layers: 1 top, 2 bottom,
                    4 left, 8 right; 0 = every tile); bits 8-27: index of the group's first weight step (= sum of the taps
                    of the groups before it; read by the balanced and the fold schedules); bits 28-30: slab rows between
                    consecutive taps, minus one (0 = taps on consecutive rows) */
} cnm_conv_group;
int cnm_conv2d_halo(const cnm_conv_desc* desc, const cnm_conv_src* srcs, const int32_t* map_info,
                    const cnm_conv_group* groups, int num_groups, const void* wp, const float* bias, void* out,
                    void* stream);

/* Same contract as cnm_conv2d_halo on CTA pairs (tcgen05.mma.cta_group::2, M = 256: csrc/conv_halo2.cu); additionally
 * needs an even number of 16x16 pixel blocks.                                                                       */
int cnm_conv2d_halo2(const cnm_conv_desc* desc, const cnm_conv_src* srcs, const int32_t* map_info,
                     const cnm_conv_group* groups, int num_groups, const void* wp, const float* bias, void* out,
                     void* stream);

/* Debugging aid (CNM_CONV_TIMELINE=1 in the environment before the first convolution): CTA 0 of every cnm_conv2d_tc
 * launch leaves clock64 stamps -- kernel entry, prologue done, dependency wait done, first TMA issued, first
 * operands landed, last MMA of the first segment issued, first accumulator complete, last store drained, exit, and (balanced
 * schedule) contributor pieces gathered -- and this call copies the 12 slots (unused ones 0) of the last launch to the
 * host (it synchronises the device).  tools/conv_timeline.py.  */
int cnm_debug_conv_timeline(long long* out12);

/* ---- conv1.0 with the plane sweep fused in -----------------------------------------
 * replaces getVolume + torch.cat + conv1[0]  (depthnet/depthNet_model.py:185-224, 231-235) in ONE kernel: the cost
 * volume never exists as a tensor.  Same contract as cnm_conv2d_halo2 for a stride-1 CNM_F16F8 layer whose source
 * `vol_src` is the [N, H, W, D] volume (Cout one tile of <= 128, no fused head); srcs[vol_src].ptr is NOT read (pass
 * any 128B-aligned device pointer).  Eight extra warps per CTA compute, for every 16x16 output tile, the
 * (16 + k - 1)^2 region of the volume around it -- exactly the arithmetic of cnm_cost_volume on the packed source
 * views, zeros outside the image -- into one of the CTA's two scratch slots in the act layout, and the TMA producer
 * boxes the operand slabs out of that slot.  A slot is rewritten every other tile, so the scratch (37 MB at 64
 * planes) lives in L2.  The output is bit-identical to cnm_pack_source + cnm_cost_volume(CNM_VOLUME_ACT) +
 * cnm_conv2d_halo2 with the same tile.                                                                             */
typedef struct cnm_sweep_fuse {
  const float* ref;        /* [B,3,H,W] fp32 reference frames */
  const void* packed_src;  /* cnm_pack_source of the N = V*B source views (pair p = v*B + b) */
  const float* hk;         /* [N,12] from cnm_camera_setup */
  void* scratch;           /* device, 128B aligned */
  int64_t scratch_bytes;   /* >= cnm_sweep_scratch_bytes(D, k) */
  int32_t B, D;            /* reference frames; planes (64 or 128) */
  int32_t vol_src;         /* index of the volume among the layer's sources */
  int32_t reserved;
  double idepth_min, idepth_max;
} cnm_sweep_fuse;
int64_t cnm_sweep_scratch_bytes(int D, int k);
int cnm_conv2d_halo2_sweep(const cnm_conv_desc* desc, const cnm_conv_src* srcs, const int32_t* map_info,
                           const cnm_conv_group* groups, int num_groups, const void* wp, const float* bias, void* out,
                           const cnm_sweep_fuse* sweep, void* stream);

/* ---- depth-driven inverse warp ("next") ----------------------------
 * replaces inverse_warp + cam2pixel  (depthnet/inverse_warp.py:46-118).
 * feat [B,C,H,W], depth [B,H,W], pose [B,3,4], K [B,3,3], Kinv [B,3,3] -> out [B,C,H,W];
 * padding_mode 0 = zeros, 1 = border.                                                      */
int cnm_inverse_warp(const float* feat, const float* depth, const float* pose, const float* K, const float* Kinv,
                     float* out, int B, int C, int H, int W, int padding_mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNMNET_B200_H_ */
