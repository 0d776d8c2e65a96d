/*
 * hd_b200.h -- C-ABI of the B200-native HMMR video->SMPL hot path (libhd_b200.so).
 *
 * The reference (akanazawa/human_dynamics) is pure Python/TF1 and has no FFI layer; its
 * boundary for this path is the Python call surface of graph-building functions plus one
 * sess.run.  Each entry point below replaces the arithmetic of the cited
 * reference function; the Python shim under src/ (same module paths, names and argument
 * meaning as the reference) binds these with ctypes -- see INTEGRATION.md.
 *
 * Conventions: every pointer is a DEVICE pointer to fp32 (or int32 where typed) unless it
 * says "host"; tensors are row-major, activations NHWC; `stream` is a cudaStream_t passed
 * as void*; all calls are asynchronous on `stream`, never synchronise, never allocate,
 * and return an hd_status (0 = ok).  No torch types cross this boundary.
 */
#ifndef HD_B200_H_
#define HD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HD_OK = 0,
  HD_ERR_INVALID = 1,      /* bad shape / null pointer / unsupported combination */
  HD_ERR_WORKSPACE = 2,    /* workspace too small */
  HD_ERR_CUDA = 3,         /* a CUDA runtime / driver call failed (see hd_last_error) */
  HD_ERR_UNSUPPORTED = 4   /* device is not sm_100 or requested impl not available */
} hd_status;

int hd_version(void);
const char *hd_status_string(int status);
/* Last CUDA error text recorded by this thread's most recent failing call ("" if none). */
const char *hd_last_error(void);
/* Number of kernels this library has launched since load / since the last reset (host counter). */
long long hd_launch_count(void);
void hd_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------
 * Fused implicit-GEMM convolution / fully-connected layer.
 * Replaces every slim conv2d / fully_connected (+ the BatchNorm/GroupNorm/ReLU/bias/residual
 * ops around it) on the path: src/models.py:65-74 (resnet_v2_50 convs), :102-113 (IEF FCs),
 * :173-184,:209-221 (temporal convs), :283-294 (fc2_res).
 *
 *   a[n,iy,ix,ci] = in[n,iy,ix,ci]                                (zero outside the image)
 *   if pre_scale:  a = a*pre_scale[n*pre_img_stride+ci] + pre_shift[...]; if pre_relu: a=max(a,0)
 *        (prologue acts on real pixels only: padding stays 0, like padding relu(bn(x)) in TF)
 *   acc[n,oy,ox,co] = sum_{ky,kx,ci} a[n, oy*stride-pad_t+ky, ox*stride-pad_l+kx, ci] * W[(ky,kx,ci),co]
 *   v = acc*post_scale[co] + post_shift[co]   (NULL scale = 1, NULL shift = 0)
 *   if res: v += res[(n*res_H + oy*res_stride)*res_W + ox*res_stride][co]
 *   if post_relu: v = max(v,0)
 *   out[(n*Ho+oy)*Wo+ox][co] = v          (out may be NULL when out_hi/out_lo are given, see below)
 * ------------------------------------------------------------------------------------------ */
enum { HD_IMPL_SIMT = 0, HD_IMPL_TC_3XTF32 = 1, HD_IMPL_TC_1XTF32 = 2, HD_IMPL_TC_3XF16 = 3 };

typedef struct {
  const float *in;  long long in_ld;          /* floats between consecutive pixels (>= Cin) */
  int n_img, H, W, Cin;
  int Ho, Wo, KH, KW, stride, pad_t, pad_l;
  const float *w_kn;                          /* [K, Cout] row-major, K=(ky,kx,ci) (TF HWIO flattened) */
  const void *w_nk_hi;                        /* [Cout_pad, K] K-major head of the split weights: fp32 holding TF32 values (impl 1,2)
                                                 or fp16 (impl 3) */
  const void *w_nk_lo;                        /* remainder, same layout: RN_tf32(w - hi), or RN_f16((w - hi) * 2^11) */
  int Cout;  int K_pad;
  const float *pre_scale, *pre_shift;  int pre_img_stride;  int pre_relu;
  const float *post_scale, *post_shift;  int post_relu;
  const float *res;  long long res_ld;  int res_H, res_W, res_stride;
  float *out;  long long out_ld;
  int impl;
  const void *tmap_hi, *tmap_lo;              /* HOST pointers to 128-byte CUtensorMap blobs from hd_make_weight_tmap */
  /* Pre-split activations (impl 3 only).  When in_hi/in_lo are set the A operand is read from two fp16 arrays of the
   * same [pixels, in_ld] geometry as `in` (hi = RN_f16(a), lo = RN_f16((a - hi) * 2^11), a = the ALREADY pre-activated
   * input) with cp.async straight into the swizzled tile -- no register staging, no prologue (pre_scale must be NULL).
   * When out_hi/out_lo are set the epilogue additionally (or, with out == NULL, only) writes
   *   y = v * post2_scale[co] + post2_shift[co]; if post2_relu: y = max(y, 0)      (NULL scale/shift = 1 / 0)
   * as such a pair [pixels, out2_ld]: the next layer's pre-activated, pre-split A operand. */
  const void *in_hi, *in_lo;
  void *out_hi, *out_lo;  long long out2_ld;
  const float *post2_scale, *post2_shift;  int post2_relu;
  /* TMA epilogue (impl 3, pre-split input, K <= 256, Cout % 32 == 0, residual row == output row): HOST pointers to 128-byte
   * CUtensorMap blobs from hd_make_act_tmap over `res`, `out`, `out_hi`, `out_lo` (each required iff that pointer is set).
   * When given, the residual is loaded and all outputs are stored as 128-row x 32-column slabs by TMA; otherwise (or with
   * HD_CONV_NO_TMA_EPILOGUE in `flags`) the epilogue uses per-thread global accesses.  Results are identical. */
  const void *tmap_res, *tmap_out, *tmap_out_hi, *tmap_out_lo;
  int flags;
  /* optional second pair of weight maps with a 64-row box (hd_make_weight_tmap(..., box_rows = 64, ...)) for Cout > 64: lets a GEMM
   * with few 128x128 tiles (2 * tiles <= #SMs) run on 64-wide tiles, i.e. on twice as many CTAs. */
  const void *tmap_hi_n64, *tmap_lo_n64;
  /* out_subsample = s > 1 (TMA-epilogue path only; HD_ERR_UNSUPPORTED elsewhere): `out` is a dense [n_img, ceil(Ho/s), ceil(Wo/s), Cout]
   * tensor that receives only the output pixels with oy % s == 0 and ox % s == 0 -- i.e. x[:, ::s, ::s, :], slim's identity shortcut
   * of the NEXT, strided unit (max_pool2d(x, [1,1], stride), A.4), written by the producing conv3 itself: the full-resolution fp32
   * block output has no other reader (the next unit's convs read out_hi / out_lo), so 3/4 of it is never written and no separate
   * hd_subsample pass runs.  out_ld is the row pitch of that dense tensor; tmap_out is not used.  0 / 1 = off. */
  int out_subsample;
} hd_conv_desc;

enum {
  HD_CONV_NO_TMA_EPILOGUE = 1,
  /* in_hi / in_lo are the padded RGBX fp16 planes written by hd_pack_conv1_planes / hd_process_image_planes: the 7x7 stride-2
   * conv1 of the ResNet root is then described as Cin = 32 (one kernel row = 8 pixels x 4 channels, 7 real + 1 zero-weight),
   * KH = 8 (7 real + 1 zero-weight), KW = 1, stride = 2, pad = 0, H = Hi + 6, W = even(Wi + 8), in_ld = 4, K = 256. */
  HD_CONV_INPUT_PLANES = 2
};

int hd_conv_gemm(const hd_conv_desc *d, void *stream);
/* Same launch; additionally CTA (0,0) of the tensor-core kernel writes per-role clock64 counters to dbg[0..15]
 * (device int64; layout in conv_simt.cu).  Tuning aid, not part of the reference surface. */
int hd_conv_gemm_profile(const hd_conv_desc *d, void *stream, long long *dbg);

/* Encode the TMA descriptor (CUtensorMap, 128 B, written to host memory `tmap_out`) for a K-major weight matrix
 * [rows, k_pad] of elem_bytes-wide elements (4 = fp32/tf32 path, 2 = fp16 path) with a {128 bytes x box_rows} box and
 * 128-byte swizzle. */
int hd_make_weight_tmap(const void *w_nk, int rows, int k_pad, int box_rows, int elem_bytes, void *tmap_out);
/* Same for a row-major activation matrix [rows, cols] with leading dimension ld_elems (elem_bytes 4 = fp32, 2 = fp16): box
 * {32 columns x 128 rows}, 128-byte (fp32) / 64-byte (fp16) swizzle -- the slabs of the TMA epilogue. */
int hd_make_act_tmap(const void *base, long long rows, int cols, long long ld_elems, int elem_bytes, void *tmap_out);

/* ---- ResNet root / tail pieces (slim resnet_v2_50, called from src/models.py:65-74) ---- */
/* Input of the tensor-core conv1 (HD_CONV_INPUT_PLANES): img fp32 [N,H,W,3] -> two fp16 planes [N, H+6, WP, 4] (head and
 * 2^11-scaled remainder of every sample, channel 3 = 0), the image at row/column offset 3 inside a zero border that the
 * caller clears ONCE (the kernel writes the interior only).  WP = plane row length in pixels (even, >= W + 8). */
int hd_pack_conv1_planes(const float *img, void *plane_hi, void *plane_lo, int N, int H, int W, int WP, void *stream);
/* conv1: 7x7 stride 2, explicit zero pad 3+3, + bias.  in [N,H,W,3] -> out [N,H/2,W/2,64]; w [7*7*3,64]. */
int hd_conv1_7x7s2(const float *in, const float *w, const float *bias, float *out, int N, int H, int W, void *stream);
/* pool1: 3x3 stride 2 max pool, TF SAME padding (pad 0 top/left, 1 bottom/right for even sizes).
 * Optional second output (out_hi/out_lo non-NULL): relu(v*scale[c] + shift[c]) as an fp16 head/remainder pair
 * (the first bottleneck unit's pre-activation, pre-split for the tensor-core kernel); `out` may then be NULL (the first unit's
 * shortcut is a conv of the pre-activation, so nobody reads the fp32 pool output). */
int hd_maxpool3x3s2_same(const float *in, float *out, int N, int H, int W, int C, const float *scale, const float *shift,
                         void *out_hi, void *out_lo, void *stream);
/* slim's identity shortcut of a strided unit: max_pool2d(x, [1,1], stride) = x[:, ::s, ::s, :] (A.4).  in [N,H,W,C] -> out
 * [N,ceil(H/s),ceil(W/s),C]; lets the residual of the unit's conv3 be a plain row-aligned tensor (TMA epilogue). */
int hd_subsample(const float *in, float *out, int N, int H, int W, int C, int stride, void *stream);
/* postnorm BN+ReLU then global mean over HxW: in [N,HW,C] -> out [N,C]. */
int hd_bnrelu_avgpool(const float *in, const float *scale, const float *shift, float *out, int N, int HW, int C, void *stream);

/* ---- crop pre-processing, the caller of the path: process_image (src/evaluation/run_video.py:56-107) + resize_img
 * (src/util/common.py:7-14), batched.  frames uint8 [N,H,W,3]; geom int32 [N,4] (16-byte aligned) = {Hs, Ws, x0, y0}: size of the
 * cv2.resize'd frame and the top-left corner of the SxS crop in its coordinates (may lie outside: edge replication = the
 * reference's np.pad(mode='edge')); out fp32 [N,S,S,3] = crop of resize(2*(frame/255 - 0.5)) (bilinear, cv2 conventions).
 * plane_hi / plane_lo (optional, both or neither; then `out` may be NULL): the same crop written directly as the padded RGBX
 * fp16 planes [N,S+6,WP,4] the tensor-core conv1 reads (hd_pack_conv1_planes layout; border cleared once by the caller). */
int hd_process_image(const unsigned char *frames, int N, int H, int W, const int *geom, float *out, int S, void *plane_hi,
                     void *plane_lo, int WP, void *stream);
/* Host bookkeeping of process_image for one frame (no CUDA call): bbox = {cx, cy, scale} as float64 -> geom = the {Hs, Ws, x0, y0} row
 * hd_process_image takes, and optionally the reference's `center` (after the crop) and `start_pt` (in the edge-padded scaled image)
 * (run_video.py:69-100; floor / round-half-to-even in double like the reference's numpy code).  HD_ERR_INVALID when the reference would
 * return a crop smaller than img_size x img_size. */
int hd_crop_geometry(int H, int W, const double *bbox, int img_size, int *geom, int *center, int *start_pt);

/* ---- f_movie GroupNorm statistics (tf.contrib.layers.group_norm at src/models.py:155,188) ----
 * x [B,T,C]; per (clip, group) mean / biased variance over T*(C/groups) elements (two-pass);
 * gain[b,c] = rsqrt(var+eps)*gamma[c]; offset[b,c] = beta[c] - mean*gain[b,c]. */
int hd_groupnorm_stats(const float *x, const float *gamma, const float *beta, float *gain, float *offset,
                       int B, int T, int C, int groups, float eps, void *stream);

/* GroupNorm + ReLU straight to the tensor-core conv's A operand format: y = relu(group_norm(x)) as an fp16 head / 2^11-scaled
 * remainder pair [B*T, C] (same statistics and affine as hd_groupnorm_stats + the conv prologue).  T * C/groups <= 1280. */
int hd_groupnorm_relu_split(const float *x, const float *gamma, const float *beta, void *out_hi, void *out_lo, int B, int T, int C,
                            int groups, float eps, void *stream);
/* fp32 [n] -> fp16 head / remainder pair (n % 4 == 0, 16-byte aligned). */
int hd_split_f16(const float *x, void *hi, void *lo, long long n, void *stream);

/* ---- IEF pieces too small / too narrow for the tensor-core tile (src/models.py:101-113,400-413) ----
 * fc1, theta part: h1 = relu(P + theta . W) with P [N,C] = phi . W1[:2048] + b1 (hoisted), theta rows of K <= 96 at stride theta_ld,
 * W [K,C]; writes h1 as an fp16 head / remainder pair (fc2's A operand) and / or fp32. */
int hd_ief_fc1_theta(const float *P, const float *theta, int theta_ld, const float *W, int K, int C, void *out_hi, void *out_lo,
                     float *out_f32, int N, void *stream);
/* fc3 + IEF update: out[n, :D] = prev[n, :D] + h2[n] . W + bias, h2 [N,K] (K % 64 == 0), W [K,D], D <= 96; fixed summation order. */
int hd_ief_fc3(const float *h2, const float *W, const float *bias, const float *prev, int prev_ld, float *out, int out_ld, int N, int K,
               int D, void *stream);

/* ---- IEF glue (src/models.py:349-371): dst[n*dst_ld + :85] = [1, 0, 0, theta[n,3:75], theta[n,75:85]] ---- */
int hd_ief_delta_init(const float *theta, float *dst, int dst_ld, int N, void *stream);

/* ---- network-level entries: library-owned layer plans for the three networks on the path ----
 * A plan packs the TF-named weights once (BatchNorm folded, K-major fp16 head / remainder split, TMA descriptors) and owns its
 * activation buffers: `*_create` allocates device memory and copies weights (synchronous, once); `*_forward` is a fixed sequence of
 * the per-layer entries above on `stream` -- no allocation, no synchronisation -- and is bit-identical to the Python host plans
 * (human_dynamics_b200/nets.py).  Precision mode: HD_IMPL_TC_3XF16 (FP32-class).
 * Weights are pulled through a callback: get(user, "<TF variable name>", &numel) returns a HOST pointer to the fp32 array in the
 * TensorFlow layout (conv HWIO, FC [in,out]) -- e.g. "resnet_v2_50/block1/unit_1/bottleneck_v2/conv1/weights" -- or NULL if absent
 * (then create fails with HD_ERR_INVALID and hd_net_error names the variable); the arrays only have to stay valid during `*_create`.
 * A plan is bound to the device that was current at creation and is not re-entrant: one `*_forward` at a time per hd_net (its
 * activation buffers are the state); use one plan per stream for concurrency.  `*_forward` can be captured in a CUDA graph.
 * hd_net_destroy frees the device memory immediately: synchronise the streams that used the plan first. */
typedef struct hd_net hd_net;
typedef const float *(*hd_weight_fn)(void *user, const char *tf_name, long long *numel);
void hd_net_destroy(hd_net *net);
const char *hd_net_error(const hd_net *net);
long long hd_net_num_launches(const hd_net *net);
/* encoder_resnet (src/models.py:50-77): images [n_frames,size,size,3] fp32 in [-1,1] -> phi [n_frames,2048].  size even. */
int hd_resnet50_create(hd_weight_fn get, void *user, int n_frames, int size, hd_net **net);
int hd_resnet50_forward(hd_net *net, const float *images, float *phi, void *stream);
/* az_fc2_groupnorm (src/models.py:121-228): phi [B,T,2048] -> movie strips [B,T,2048] (out != phi).  T <= 20. */
int hd_fmovie_create(hd_weight_fn get, void *user, int B, int T, int num_conv_layers, hd_net **net);
int hd_fmovie_forward(hd_net *net, const float *phi, float *out, void *stream);
/* call_hmr_ief (src/models.py:299-415) as wired by tester.py:196-207 (scope single_view_ief, 3 stages, use_optcam, deltas started
 * from the main prediction): phi [N,2048] -> theta [N,85] and, for the num_delta non-zero delta_t values in ascending order,
 * deltas [N,num_delta,85] = [1,0,0 | pose | beta].  IEF starts from the checkpoint's `mean_param`. */
int hd_ief_create(hd_weight_fn get, void *user, int N, const int *delta_t, int num_delta, hd_net **net);
int hd_ief_forward(hd_net *net, const float *phi, float *theta, float *deltas, void *stream);

/* ---- SMPL (src/tf_smpl/batch_smpl.py:26-162, batch_lbs.py:15-60,133-194, projection.py:16-29) ---- */
typedef struct {
  int num_verts, num_kps, lbs_nnz, kp_nnz_total;
  const float *v_template;     /* [V*3] */
  const float *dirs;           /* [10+207, V*3]: shapedirs rows then posedirs rows (batch_smpl.py:45-48,60-63) */
  const float *J_template;     /* [24*3]  = J_regressor^T v_template */
  const float *J_shapedirs;    /* [10, 24*3] = J_regressor^T shapedirs  (exact refactoring of batch_smpl.py:110-118) */
  const int *lbs_idx;          /* [V, lbs_nnz] joint ids of the non-zero skinning weights (padded with weight 0) */
  const float *lbs_w;          /* [V, lbs_nnz] */
  const int *kp_ptr;           /* [K+1] CSC offsets into kp_vidx / kp_w (cocoplus_regressor, batch_smpl.py:76-82) */
  const int *kp_vidx;          /* [kp_nnz_total] */
  const float *kp_w;           /* [kp_nnz_total] */
  int parents[24];             /* kintree_table[0] as int32 (root -1), batch_smpl.py:66 */
} hd_smpl_consts;

/* Host-side packing of the SMPL model (the contents of the official pickle as dense float64 arrays) into the arrays hd_smpl_consts points
 * at -- what SMPL.__init__ does in the reference (batch_smpl.py:27-86) plus the layouts the kernels want.  Pure host code, no CUDA call:
 * the caller uploads the outputs and stores the device pointers in hd_smpl_consts.  All pointers are HOST memory.
 *   in : v_template [V,3], shapedirs [V,3,10], posedirs [V,3,207], J_regressor [24,V], weights [V,24], kp_regressor [K,V] (cocoplus_regressor,
 *        or its first 14 rows for joint_type 'lsp', batch_smpl.py:81-82), kintree_parents = kintree_table[0] as stored (uint32, root 4294967295)
 *   out: v_template_out [V*3] f32, dirs [217, V*3] f32, J_template [72] f32, J_shapedirs [10,72] f32, lbs_idx / lbs_w [V, lbs_nnz] (ELL: joint
 *        ids ascending, padding = joint 0 with weight 0), kp_ptr [K+1] / kp_vidx / kp_w [kp_nnz_total] (CSC over keypoints), parents int[24]
 * hd_smpl_pack_sizes reports lbs_nnz (4, or the per-vertex maximum rounded up to a multiple of 4, at most 24) and kp_nnz_total first. */
int hd_smpl_pack_sizes(int V, int K, const double *weights, const double *kp_regressor, int *lbs_nnz, int *kp_nnz_total);
int hd_smpl_pack(int V, int K, const double *v_template, const double *shapedirs, const double *posedirs, const double *J_regressor,
                 const double *weights, const double *kp_regressor, const unsigned int *kintree_parents, float *v_template_out, float *dirs,
                 float *J_template, float *J_shapedirs, int *lbs_idx, float *lbs_w, int lbs_nnz, int *kp_ptr, int *kp_vidx, float *kp_w,
                 int *parents);
size_t hd_smpl_workspace_bytes(int N);
/* beta rows of 10 at stride beta_ld, theta rows of 72 at stride theta_ld, cam rows of 3 at stride cam_ld (so the
 * three can alias columns [75:85], [3:75], [0:3] of one [N,85] omega buffer, src/omega.py:231-235).
 * verts [N,V,3], joints [N,K,3], Rs [N,24,3,3], Jtr [N,24,3]; cam + kps [N,K,2] optional (both or neither).
 * Any output pointer except verts may be NULL.  Pose n lands in output slot n*out_mul + out_off of every output
 * array (out_mul=1, out_off=0 = dense), so D delta heads can write the [B,T,D,...] stacking of tester.py:252-253
 * in place. */
int hd_smpl_forward(const hd_smpl_consts *c, const float *beta, int beta_ld, const float *theta, int theta_ld, int N,
                    float *verts, float *joints, float *Rs, float *Jtr,
                    const float *cam, int cam_ld, float *kps, int out_mul, int out_off,
                    void *ws, size_t ws_bytes, void *stream);
/* Staged form of the same computation for large batches: the dense shape/pose blend (batch_smpl.py:110-112,127-133) is one
 * [N,256] x [256, V*3] GEMM on the tensor cores (hd_conv_gemm over `coef` with the packed `dirs`, bias = v_template), so the
 * rest is HBM-shaped:  hd_smpl_pose (Rodrigues + FK; also writes the GEMM operand rows coef[n] = [beta | R-I | 0]) ->
 * hd_conv_gemm -> hd_smpl_lbs (skinning of the blended v_posed [N, vp_ld]) -> hd_smpl_joints (keypoints + projection).
 * ws of hd_smpl_pose: N*216 floats. */
int hd_smpl_pose(const hd_smpl_consts *c, const float *beta, int beta_ld, const float *theta, int theta_ld, int N, float *Rs,
                 float *Jtr, float *A12, float *coef /* fp32 rows, nullable */, int coef_ld,
                 void *coef_hi, void *coef_lo /* the same rows as an fp16 head / 2^11-scaled remainder pair, nullable */,
                 void *a12t_hi, void *a12t_lo /* [N,12,32] fp16 head / unscaled remainder of A transposed: operand of hd_smpl_lbs_tc, nullable */,
                 int out_mul, int out_off, void *ws, size_t ws_bytes, void *stream);
/* Skinning with the 6890x24 blend-weight x transform contraction on the tensor cores (batch_smpl.py:141-151):
 * w_hi / w_lo = dense skinning weights [roundup128(V), 32] (24 joints + zero padding) as an fp16 head / unscaled remainder pair;
 * a12t_hi / a12t_lo from hd_smpl_pose; v_posed [N, vp_ld] (vp_ld % 4 == 0, >= roundup4(3V)) -> verts slot n*out_mul + out_off. */
int hd_smpl_lbs_tc(const void *w_hi, const void *w_lo, const void *a12t_hi, const void *a12t_lo, const float *v_posed, long long vp_ld,
                   float *verts, int N, int V, int out_mul, int out_off, void *stream);
int hd_smpl_lbs(const hd_smpl_consts *c, const float *v_posed, long long vp_ld, const float *A12, float *verts, int N, int out_mul,
                int out_off, void *stream);
int hd_smpl_joints(const hd_smpl_consts *c, const float *verts, const float *cam, int cam_ld, float *joints, float *kps, int N,
                   int out_mul, int out_off, void *stream);
/* batch_rodrigues: theta [M,3] -> R [M,3,3]. */
int hd_rodrigues(const float *theta, float *R, int M, void *stream);
/* batch_rot2aa (src/tf_smpl/batch_lbs.py:63-105): R [M,3,3] -> axis-angle [M,3]. */
int hd_rot2aa(const float *Rs, float *aa, int M, void *stream);
/* batch_global_rigid_transformation: Rs [N,24,3,3], Js [N,24,3], parents host int[24] -> new_J [N,24,3], A [N,24,4,4]. */
int hd_global_rigid(const float *Rs, const float *Js, const int *parents_host, float *new_J, float *A44, int N,
                    int rotate_base, void *stream);
/* batch_orth_proj_idrot: X [N,P,3], cam [N,3] -> out [N,P,2]. */
int hd_orth_proj(const float *X, const float *cam, float *out, int N, int P, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HD_B200_H_ */
