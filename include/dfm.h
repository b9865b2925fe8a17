/*
 * dfm.h -- C ABI of libdfm.so, the B200 (sm_100a) deformation engine.
 *
 * This is the drop-in boundary for the deformation hot path of
 * ivadomed/multimodal-registration.  The reference reaches that path through the Python
 * API of voxelmorph / neurite (Keras layers and eager functions on channels-last tensors);
 * each entry point below names the reference interface it replaces.  Citations are
 * file:line under the reference checkout; "[UR]" marks interfaces that live in the
 * un-vendored packages voxelmorph@52dd120f / neurite@c7bb05d5 (reference README.md:35-37)
 * and are specified in SURVEY.md Appendix A.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller; the library allocates
 *     no device memory and keeps no device state between calls.  The only thing it remembers
 *     is a bounded host-side cache of texture DESCRIPTORS (cudaTextureObject_t over caller
 *     memory, keyed by device / base pointer / geometry; no copy of the data) used by the
 *     one-channel linear warps -- a descriptor names an address range, it does not own it;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation, no default-stream use -> capturable in CUDA graphs;
 *   - volumes are fp32 unless stated.  A tensor is either "planar"  [B][C][X][Y][Z]
 *     or "channels-last" (the reference layout) [B][X][Y][Z][C]; Z is fastest in both.
 *     `flags` says which (DFM_*_CL bits); a displacement field has C = 3 and component d
 *     displaces along axis d in voxels of its own grid ('ij' indexing);
 *   - return value 0 on success, a negative DFM_E* code otherwise; dfm_last_error() gives
 *     the message for the calling thread.  Nothing aborts or throws across the ABI.
 */
#ifndef DFM_H_
#define DFM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFM_VERSION 100 /* 0.1.0 */

/* error codes */
#define DFM_OK 0
#define DFM_EINVAL (-1)       /* bad shape / argument */
#define DFM_EALIGN (-2)       /* pointer not aligned as required */
#define DFM_EUNSUPPORTED (-3) /* valid in the reference but not implemented here */
#define DFM_ECUDA (-4)        /* CUDA runtime error (message has the detail) */

/* interpolation (interp_method of the reference: 'linear' | 'nearest') */
#define DFM_LINEAR 0
#define DFM_NEAREST 1

/* layout flags (bit set = channels-last, clear = planar) */
#define DFM_FIELD_IN_CL 1u  /* input displacement field(s)            */
#define DFM_FIELD_OUT_CL 2u /* output displacement field              */
#define DFM_IMG_CL 4u       /* image / volume, input and output alike */
#define DFM_LOC_ABSOLUTE 8u /* dfm_warp_fwd: `field` holds absolute sample locations (ne.utils.interpn)
                               instead of displacements added to the voxel grid */

int dfm_version(void);
/* 1: this build keeps the reference's op order with separately rounded fp32 ops (libdfm_exact.so,
 * bit-identical to the oracle); 0: fused/packed accumulation (libdfm.so, the default; differs by the
 * removed intermediate roundings only, ~1e-7 relative). */
int dfm_exact_order(void);
const char *dfm_last_error(void);

/* ---------------------------------------------------------------------------------------
 * SpatialTransformer / vxm.utils.transform / ne.utils.interpn  [UR]
 *   replaces: vxm.layers.SpatialTransformer (train_synthmorph.py:298),
 *             vxm.networks.Transform(...).predict (gen_apply_def_field.py:74-76,
 *             3d_reg.py:331-334,377-380, bids_registration.py:335-338,380-383,
 *             bids_two_steps_registration.py:338-341,354-355,400-401,444-447,496-499),
 *             vxm.utils.transform (train_synthmorph.py:67, channel-wise: pass B*C items of
 *             one channel each).
 *   out[b,c,p] = interp(img[b,c], p + field[b,:,p]), edge clamp; if has_fill, samples whose
 *   UNCLIPPED location is outside [0, dim-1] become `fill`.
 *   img: C channels of (Xi,Yi,Zi); field/out grid: (X,Y,Z).  elem_size: bytes per image
 *   element -- 4 (fp32) for DFM_LINEAR; 1, 2, 4 or 8 for DFM_NEAREST (values are moved,
 *   not interpreted; `fill` is then given as raw bits in fill_bits).
 * ------------------------------------------------------------------------------------- */
int dfm_warp_fwd(const void *img, const float *field, void *out,
                 int B, int C, int Xi, int Yi, int Zi, int X, int Y, int Z,
                 int interp, int elem_size, int has_fill, float fill, uint64_t fill_bits,
                 unsigned flags, void *stream);

/* Channel-wise vxm.utils.transform [UR] (train_synthmorph.py:61-67, generate_label_maps): every channel of a
 * channels-last volume moves by its own 3-vector field, both tensors in the reference's layout, no transposes.
 *   vol [B][Xi][Yi][Zi][C], shift [B][X][Y][Z][C][3] -> out[b,p,c] = interp(vol[b,...,c], p + shift[b,p,c,:])
 *   (linear, edge clamp, optional fill as in dfm_warp_fwd).  argmax = 0: out is float [B][X][Y][Z][C];
 *   argmax = 1 (C <= 256): the tf.argmax(im, axis=-1) that follows at train_synthmorph.py:68 is taken in the
 *   kernel (first maximum) and out is uint8 [B][X][Y][Z].  DFM_EUNSUPPORTED when X*Y*Z*C*C >= 2^32. */
int dfm_warp_channelwise_fwd(const float *vol, const float *shift, void *out, int B, int C, int Xi, int Yi, int Zi,
                             int X, int Y, int Z, int has_fill, float fill, int argmax, void *stream);

/* Fused RescaleTransform(factor >= 1) + linear SpatialTransformer of a one-channel image: the
 * deformation tail of VxmDense at inference (3d_reg.py:305,310; bids_*.py:311-322), where the
 * full-resolution warp is only an intermediate.  out[b,p] = interp(img[b], p + U[b,:,p]) with
 * U = resize(factor * coarse) evaluated on the fly, so U never touches HBM.  Where the shapes allow
 * (up-sampling whose coarse box fits the marching tile, image rows 32-byte multiples, X*Y <= 65000)
 * the field is marched like dfm_resize_fwd's up-sampler and the image corners are fetched by texture
 * gathers: the same arithmetic as dfm_resize_fwd followed by dfm_warp_fwd, bit for bit, in both
 * builds.  Other shapes: libdfm_exact.so the same bits, libdfm.so a separable evaluation a few ulp
 * from it.
 *   img [B][Xi][Yi][Zi], coarse [B][3][Xh][Yh][Zh] planar, out [B][X][Y][Z]; cx/cy/cz as in
 *   dfm_resize_fwd (X/Y/Z entries).  work: nullable scratch of B*3*X*Y*Z floats used when the
 *   fused kernel is not applicable (then the two kernels run back to back); without it such
 *   shapes return DFM_EUNSUPPORTED. */
int dfm_rescale_warp_fwd(const float *img, const float *coarse, float *out, const float *cx, const float *cy,
                         const float *cz, float *work, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh,
                         int X, int Y, int Z, float factor, int has_fill, float fill, void *stream);

/* The same fusion for NEAREST-neighbour warps of one-channel volumes of 4-byte elements (label maps and
 * segmentations: Transform(interp_method='nearest', rescale=scale) at 3d_reg.py:377-380,
 * bids_registration.py:380-383, bids_two_steps_registration.py:338-341,354-355): U is marched like
 * dfm_resize_fwd's up-sampler and each output picks img at round-half-even(p + U), clipped; values are
 * moved as raw bits (fill_bits: the 32-bit fill pattern).  Bit-identical to dfm_resize_fwd followed by
 * dfm_warp_fwd(DFM_NEAREST) in both builds.  work as in dfm_rescale_warp_fwd. */
int dfm_rescale_warp_nearest_fwd(const void *img, const float *coarse, void *out, const float *cx, const float *cy,
                                 const float *cz, float *work, int B, int Xi, int Yi, int Zi, int Xh, int Yh,
                                 int Zh, int X, int Y, int Z, float factor, int has_fill, uint32_t fill_bits,
                                 void *stream);

/* Linear warp of a ONE-HOT label map straight from the label map (forward and d/d field):
 *   replaces: pred = vxm.layers.SpatialTransformer(interp_method='linear')([map_1, flow]) at train_synthmorph.py:298,
 *             where map_1 is the one-hot output of ne.models.labels_to_image (:288-289) -- one-hot by construction.
 *   labels [B][Xi][Yi][Zi] uint8 (values >= C act as an all-zero row), field [B][3][X][Y][Z] (or channels-last,
 *   DFM_FIELD_IN_CL), out / gout [B][X][Y][Z][C] channels-last (the reference layout), C <= 256.
 *   out[b,p,c] = sum of the trilinear corner weights whose corner label is c, accumulated in corner order: bit-identical
 *   to dfm_warp_fwd on the materialised one-hot tensor (DFM_IMG_CL) in both builds, without building or reading it.
 *   dfm_warp_onehot_bwd: gfield [B][3][X][Y][Z] (DFM_FIELD_OUT_CL: channels-last) is overwritten with d/d field of
 *   sum(gout * out) (TensorFlow autodiff semantics as dfm_warp_bwd; labels carry no gradient); C <= 48. */
int dfm_warp_onehot_fwd(const uint8_t *labels, const float *field, float *out, int B, int C, int Xi, int Yi, int Zi,
                        int X, int Y, int Z, int has_fill, float fill, unsigned flags, void *stream);
int dfm_warp_onehot_bwd(const float *gout, const uint8_t *labels, const float *field, float *gfield, int B, int C,
                        int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, unsigned flags, void *stream);

/* Backward of the linear warp (TensorFlow autodiff semantics of the reference graph:
 * floor has zero gradient, clip passes gradient inside [0, max] inclusive).
 *   gimg   (nullable): [B,C,Xi,Yi,Zi] is ACCUMULATED INTO (caller zeroes it).
 *   gfield (nullable): [B,3,X,Y,Z] is overwritten.
 *   replaces: the gradient of `pred` at train_synthmorph.py:298,305-306 (d/dfield only) and
 *   of the SpatialTransformer inside VxmDense (train_synthmorph.py:296). */
int dfm_warp_bwd(const float *gout, const float *img, const float *field,
                 float *gimg, float *gfield,
                 int B, int C, int Xi, int Yi, int Zi, int X, int Y, int Z,
                 int has_fill, unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------
 * out = scale*own + interp(scale*src, p + scale*own)        (3-channel fields)
 *   src == own : one scaling-and-squaring step of vxm.utils.integrate_vec [UR]
 *   src != own : vxm.utils.compose([src, own]) [UR]
 *                (bids_two_steps_registration.py:324,346,369,484)
 *   src grid (Xs,Ys,Zs); own/out grid (X,Y,Z).  `scale` must be a power of two (1 for
 *   compose); it folds integrate_vec's  v / 2**nb_steps  into the first step.
 * ------------------------------------------------------------------------------------- */
int dfm_field_warp_add(const float *src, const float *own, float *out,
                       int B, int Xs, int Ys, int Zs, int X, int Y, int Z,
                       float scale, int interp, unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------
 * VecInt(method='ss', int_steps) -> vxm.utils.integrate_vec [UR]
 *   replaces: the VecInt layer inside VxmDense (3d_reg.py:305, bids_registration.py:311,
 *   bids_two_steps_registration.py:311,314, train_synthmorph.py:296) and labels_to_image
 *   (train_synthmorph.py:288-289).
 *   svf -> out on grid (X,Y,Z), nsteps >= 0.  `work`: caller scratch of
 *   dfm_vecint_workspace_bytes() bytes (planar fp32).  If save_steps != 0, work receives
 *   v_0 .. v_{nsteps-1} (the inputs of every step, v_0 = svf / 2**nsteps), which
 *   dfm_vecint_bwd needs.  The workspace ends with three 256-byte aligned arrays of B floats: the
 *   per-item maximum displacement measured by the first step and by the two steps before the last
 *   (kernel selection per item on the device, no host synchronisation; dfm_vecint_bwd reads the first).
 * ------------------------------------------------------------------------------------- */
size_t dfm_vecint_workspace_bytes(int B, int X, int Y, int Z, int nsteps, int save_steps);
int dfm_vecint_fwd(const float *svf, float *out, float *work,
                   int B, int X, int Y, int Z, int nsteps, int save_steps,
                   unsigned flags, void *stream);
/* gsvf = d loss / d svf given gout = d loss / d out and `saved` = the workspace dfm_vecint_fwd filled with
 * save_steps != 0 (the step inputs AND the trailing B-float displacement bound, which selects per item between the
 * atomics-free gather adjoint and the scatter adjoint of each step).  `scratch`: 2 * B*3*X*Y*Z floats.  All planar. */
int dfm_vecint_bwd(const float *gout, const float *saved, float *gsvf, float *scratch,
                   int B, int X, int Y, int Z, int nsteps, void *stream);

/* One SS step backward: given g = dL/dv' and the step input v (v' = v + interp(v, p+v)),
 * gv[b,:,p] = scale * ( g + dloc-term ) and the dvol-term is scatter-ADDED into gv
 * (so gv must not alias g).  Exposed for tests. */
int dfm_ss_step_bwd(const float *g, const float *v, float *gv,
                    int B, int X, int Y, int Z, float scale, void *stream);
/* The same with a displacement bound: item b satisfies |v| <= bound[b] * bscale (device array of B floats).  Items
 * bounded below one voxel take the gather formulation of the volume path (no atomics, deterministic); the others
 * scatter with red.global.add.  dfm_ss_step_bwd is this call without a bound. */
int dfm_ss_step_bwd_bounded(const float *g, const float *v, float *gv, const float *bound, float bscale,
                            int B, int X, int Y, int Z, float scale, void *stream);

/* ---------------------------------------------------------------------------------------
 * RescaleTransform(zoom) -> vxm.utils.rescale_dense_transform -> ne.utils.resize [UR]
 *   replaces: 3d_reg.py:394, bids_registration.py:398, bids_two_steps_registration.py:515,
 *   Transform(rescale=scale) at the sites listed under dfm_warp_fwd, RescaleTransform
 *   inside VxmDense / labels_to_image.
 *   out[b,c,jx,jy,jz] = post * interp(pre * in[b,c], (cx[jx], cy[jy], cz[jz]))
 *   cx/cy/cz: DEVICE arrays of Xo/Yo/Zo fp32 sample coordinates on the input grid
 *   (tf.linspace(0, n_in-1, n_out) in the reference; computed by the host shim so the
 *   coordinate convention stays a host decision).  factor >= 1: pre = factor, post = 1;
 *   factor < 1: pre = 1, post = factor.  C channels (3 for a field; any C for ne.utils.resize).
 * ------------------------------------------------------------------------------------- */
int dfm_resize_fwd(const float *in, float *out, const float *cx, const float *cy, const float *cz,
                   int B, int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo,
                   float pre, float post, int interp, unsigned flags, void *stream);
/* Adjoint of the linear dfm_resize_fwd, written as a gather (no atomics).  Planar only.
 *   gin[b,c,i] = pre*post * sum_kx sum_ky sum_kz xw[ix][kx] yw[iy][ky] zw[iz][kz] *
 *                gout[b,c, xlo[ix]+kx, ylo[iy]+ky, zlo[iz]+kz]
 * Per axis (DEVICE arrays, computed by the host from cx/cy/cz): lo[n_in] = first output index
 * whose interpolation support touches input i, cnt[n_in] = number of such outputs (contiguous),
 * w[n_in][k] = the weight output lo[i]+k puts on input i (row length k = kx/ky/kz). */
int dfm_resize_bwd(const float *gout, float *gin,
                   const int *xlo, const int *xcnt, const float *xw, int kx,
                   const int *ylo, const int *ycnt, const float *yw, int ky,
                   const int *zlo, const int *zcnt, const float *zw, int kz,
                   int B, int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo,
                   float pre, float post, void *stream);
/* The same with a caller-provided workspace of dfm_resize_bwd_workspace_bytes(B, C, Xi, Yi, Zo) bytes: up-sampling
 * adjoints (>= 3 taps per axis) then run as two separable passes (x,y then z: K^2 + K instead of K^3 gathers per input
 * sample).  work == NULL is dfm_resize_bwd. */
size_t dfm_resize_bwd_workspace_bytes(int B, int C, int Xi, int Yi, int Zo);
int dfm_resize_bwd_ws(const float *gout, float *gin, float *work,
                      const int *xlo, const int *xcnt, const float *xw, int kx,
                      const int *ylo, const int *ycnt, const float *yw, int ky,
                      const int *zlo, const int *zcnt, const float *zw, int kz,
                      int B, int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo,
                      float pre, float post, void *stream);

/* ---------------------------------------------------------------------------------------
 * Jacobian-determinant map  (eval_reg_with_jacobian.py:62-78)
 *   field: [3][X][Y][Z] planar or [X][Y][Z][3] channels-last (DFM_FIELD_IN_CL), fp32
 *          (in_f64 = 0) or fp64 (in_f64 = 1).
 *   det:   [(X-4)(Y-4)(Z-4)] fp32 (out_f64 = 0) or fp64 (out_f64 = 1); nullable.
 *   4th-order central differences on the interior, det(I + J), all arithmetic in fp64.
 *   stats (device, 4 doubles, nullable): n_negative (det < 0), sum(det), sum(det^2), n_total.
 *   partials: caller scratch of dfm_jacdet_workspace_bytes() bytes.  B fields per call;
 *   stats is then [B][4].
 * ------------------------------------------------------------------------------------- */
size_t dfm_jacdet_workspace_bytes(int B, int X, int Y, int Z);
int dfm_jacdet(const void *field, void *det, double *stats, void *partials,
               int B, int X, int Y, int Z, int in_f64, int out_f64,
               unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------
 * Sub-volume stitching  (3d_reg.py:214-259 get_def_field_from_subvol; duplicated in
 * bids_registration.py:226-271 and bids_two_steps_registration.py:226-271)
 *   tiles: T fields of shape (tx,ty,tz), channels-last [T][tx][ty][tz][3] (DFM_FIELD_IN_CL, what
 *          model.predict returns) or planar [T][3][tx][ty][tz]; mins: DEVICE int [T][3], the
 *          (x_min, y_min, z_min) of every tile in the volume; out: [X][Y][Z][3] (DFM_FIELD_OUT_CL)
 *          or planar [3][X][Y][Z], fp32 or fp64 (out_f64).
 *   out[p] = sum_t (w_t(p) / sum_t' w_t'(p)) * tile_t[p - min_t],  w = 1 - max(|x|,|y|,|z|)/(max+1)
 *   on the tile-centred grid [-s/2, s/2); float64 arithmetic in tile order (bit-identical to the
 *   reference); voxels no tile covers are 0.
 * ------------------------------------------------------------------------------------- */
int dfm_stitch_subvol(const float *tiles, const int *mins, void *out, int T, int tx, int ty, int tz,
                      int X, int Y, int Z, int out_f64, unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------
 * Losses adjacent to the warp (train_synthmorph.py:301-306) -- the voxel-level parts.
 *   Dice: vxm.losses.Dice().loss(y_true, y_pred) [UR] on one-hot maps, channels-last [B][N][C]
 *   (DFM_IMG_CL, the reference layout) or planar [B][C][N]:
 *     dfm_dice_sums:  sums[b][c] = (sum_n t*p, sum_n t+p) as doubles ([B][C][2], device); the loss is
 *                     -mean_{b,c} divide_no_nan(2*sums0, sums1), formed by the caller from B*C scalars.
 *     dfm_dice_bwd:   g_pred[b,n,c] = coef[b][c][0] * y_true[b,n,c] + coef[b][c][1]  (coef fp32, device).
 *   Grad: vxm.losses.Grad('l2', loss_mult).loss(None, flow) [UR] on a field [B][3][X][Y][Z] planar or
 *   channels-last (DFM_FIELD_IN_CL):
 *     dfm_grad_l2_sums: sums[b][axis] = sum of squared forward differences along the axis over all three
 *                     components ([B][3] doubles); loss[b] = loss_mult/3 * sum_axis sums/count_axis.
 *     dfm_grad_l2_bwd:  g[b,c,n] = sum_axis coef[b][axis] * ((v[n]-v[n-e]) - (v[n+e]-v[n])), missing
 *                     neighbours dropped (coef = 2*loss_mult*upstream / (3*count_axis), fp32, device).
 *   work: caller scratch of dfm_*_workspace_bytes() bytes (per-block partial sums, combined in a fixed
 *   order: results are deterministic).
 * ------------------------------------------------------------------------------------- */
size_t dfm_dice_workspace_bytes(int B, int C, size_t N);
int dfm_dice_sums(const float *y_true, const float *y_pred, double *sums, void *work, int B, int C, size_t N,
                  unsigned flags, void *stream);
int dfm_dice_bwd(const float *y_true, const float *coef, float *g_pred, int B, int C, size_t N, unsigned flags,
                 void *stream);
size_t dfm_grad_l2_workspace_bytes(int B, int X, int Y, int Z);
int dfm_grad_l2_sums(const float *flow, double *sums, void *work, int B, int X, int Y, int Z, unsigned flags,
                     void *stream);
int dfm_grad_l2_bwd(const float *flow, const float *coef, float *g, int B, int X, int Y, int Z, unsigned flags,
                    void *stream);

/* d Dice / d field through SpatialTransformer('linear') (train_synthmorph.py:298 + :305) with the Dice
 * gradient formed on the fly, g[b,n,c] = coef[b][c][0] * y_true[b,n,c] + coef[b][c][1], instead of being
 * materialised by dfm_dice_bwd and read back by dfm_warp_bwd.  Channels-last maps only (img [B][Ni][C],
 * y_true [B][N][C]); field planar or channels-last (DFM_FIELD_IN_CL); gfield planar, or channels-last with
 * DFM_FIELD_OUT_CL; 2 <= C <= 32, or even C <= 64 (else DFM_EUNSUPPORTED: run the two calls). */
int dfm_warp_dice_bwd(const float *y_true, const float *coef, const float *img, const float *field, float *gfield,
                      int B, int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill,
                      unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------
 * Layout conversion between the reference's channels-last tensors and planar tensors.
 *   cl [B][N][C]  <->  planar [B][C][N],  elem_size in {1, 2, 4, 8}.
 * ------------------------------------------------------------------------------------- */
int dfm_cl_to_planar(const void *cl, void *planar, int B, int C, size_t N, int elem_size, void *stream);
int dfm_planar_to_cl(const void *planar, void *cl, int B, int C, size_t N, int elem_size, void *stream);

/* ---------------------------------------------------------------------------------------
 * Voxel-level evaluation metrics.  Inputs are float32 (is_f64 = 0) or float64
 * (is_f64 = 1, what nibabel's get_fdata() hands the reference scripts) device arrays of n elements.
 *   dfm_minmax       out2 = {min, max};  work: dfm_metrics_workspace_bytes() of scratch.
 *   dfm_joint_hist   replaces np.histogramdd([a, b], bins) of normalized_mutual_information
 *                    (eval_reg_with_mi.py:66-69): edges np.linspace(min, max, bins + 1) per image in float64,
 *                    right-most edge inclusive; hist [bins][bins] uint64 (row = bin of a), overwritten.
 *                    minmax_a / minmax_b: DEVICE {min, max} pairs from dfm_minmax (no host round trip).
 *   dfm_axis_sums    the three plane sums of detect_zero_padding (eval_reg_with_mi.py:16-36):
 *                    xs[x] = sum_{y,z}, ys[y] = sum_{x,z}, zs[z] = sum_{x,y}, float64, overwritten.
 *   dfm_overlap_sums the masked sums of eval_reg_on_sc_seg.py:80-93: out5 = {sum(m[fx == 1]), sum(m[fx == 0]),
 *                    count(fx == 1), count(fx == 0), sum(m)} in float64; work as for dfm_minmax.
 * ------------------------------------------------------------------------------------- */
size_t dfm_metrics_workspace_bytes(void);
int dfm_minmax(const void *a, size_t n, int is_f64, double *out2, double *work, void *stream);
int dfm_joint_hist(const void *a, const void *b, size_t n, int is_f64, const double *minmax_a, const double *minmax_b,
                   int bins, unsigned long long *hist, void *stream);
int dfm_axis_sums(const void *im, int X, int Y, int Z, int is_f64, double *xs, double *ys, double *zs, void *stream);
int dfm_overlap_sums(const void *fx, const void *m, size_t n, int is_f64, double *out5, double *work, void *stream);

/* ---------------------------------------------------------------------------------------
 * Intensity model of ne.models.labels_to_image [UR] (train_synthmorph.py:258-289) -- what
 * follows the label-map deformation (dfm_vecint_fwd -> dfm_resize_fwd -> dfm_warp_fwd nearest, fill 0).
 *   dfm_synth_intensity  out[i] = means[l] + stds[l] * N(0,1), l = (int)labels[i] clamped to the table; Philox-4x32-10
 *                        keyed by `seed`, a pure function of (seed, i).  labels: float label map of n voxels.
 *   dfm_conv1d_axis      one pass of a separable blur: K (odd) taps along axis 0 / 1 / 2 of [B][X][Y][Z], zero padding.
 *   dfm_scale_exp_clip   out = clip(img * exp(logbias), lo, hi)   (logbias nullable: clip only).
 *   dfm_norm_gamma       out = ((img - min_b) / (max_b - min_b)) ** gamma_b per item; minmax: B device {min, max}
 *                        pairs (float64, from dfm_minmax), gamma: B floats (nullable: 1).
 *   dfm_onehot           out[i][c] = (lut[(int)labels[i]] == c), channels-last [n][C]; labels outside the table or
 *                        mapped to a negative entry give an all-zero voxel (out_label_list semantics).
 * ------------------------------------------------------------------------------------- */
int dfm_synth_intensity(const float *labels, const float *means, const float *stds, int nlabels, uint64_t seed, float *out,
                        size_t n, void *stream);
int dfm_conv1d_axis(const float *in, float *out, int B, int X, int Y, int Z, int axis, const float *taps, int K, void *stream);
int dfm_scale_exp_clip(const float *img, const float *logbias, float *out, size_t n, float lo, float hi, void *stream);
int dfm_norm_gamma(const float *img, const double *minmax, const float *gamma, float *out, int B, size_t n_per_item, void *stream);
int dfm_onehot(const float *labels, const int *lut, int nlut, int C, float *out, size_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DFM_H_ */
