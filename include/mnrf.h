/* mnrf.h -- C ABI of the B200-native MultiNeRF per-ray core (libmnrf_b200.so).
 *
 * The reference (google-research/multinerf) has NO native boundary: its operator API is
 * the Python signature `Model.__call__` (internal/models.py:75-312) plus the free
 * functions of internal/{stepfun,render,coord}.py, all lowered by XLA.  This header is
 * the boundary a maintainer would bind instead (ctypes stub in INTEGRATION.md); each entry
 * names the reference code it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (fp32 row-major, samples on the
 *    last axis, unless a parameter says bf16); entries only write their declared outputs;
 *  - entries enqueue work on `stream` and return immediately: no allocation, no host
 *    synchronisation, no global mutable state (except the tensor-map encoder lookup);
 *  - return 0 on success, non-zero on error; the message is in mnrf_last_error()
 *    (thread-local).  Python-side config errors of the reference (ValueError at trace
 *    time) stay Python-side; the ABI reports shape / alignment / launch failures;
 *  - bf16 buffers are raw uint16 storage (`mnrf_bf16`);
 *  - WORKSPACE POLICY: no entry needs scratch beyond its declared arguments (every intermediate is an
 *    argument the caller allocates: per-level activation / mask / gradient buffers are sized from the
 *    shapes documented per entry), so there are no `*_workspace_bytes` queries; kernels keep their
 *    staging in shared memory / TMEM.  The Python host (multinerf_b200/models.py `_level_state`)
 *    allocates each buffer once per (level, shape) and never frees it while a captured graph lives.
 */
#ifndef MNRF_H_
#define MNRF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t mnrf_bf16;
typedef void* mnrf_stream;   /* cudaStream_t */

#define MNRF_ABI_VERSION 1

/* ---- library ------------------------------------------------------------------------- */
int mnrf_abi_version(void);
const char* mnrf_last_error(void);
/* 1 if the current device is sm_100 (B200) and the tcgen05/TMA path can run. */
int mnrf_device_ok(void);
int mnrf_num_sms(void);

/* ---- hierarchical resampling ---------------------------------------------------------
 * Replaces, for one level: stepfun.max_dilate_weights (stepfun.py:116-128) + the [1:-1]
 * trim (models.py:170-171) + the annealed logits (models.py:183-185) + stepfun.
 * sample_intervals (stepfun.py:214-263: softmax, integrate_weights, sorted_interp
 * math.py:108-127, midpoints).  One warp owns one ray.
 *   sdist_prev [B, P+1], w_prev [B, P]           previous step function
 *   u_base     [S]      host-computed linspace grid of stepfun.py:194-206
 *   jitter     raw U[0,1): NULL | [B] (single_jitter) | [B, S]
 *   cw_in      optional [B, P'+1] CDF to use instead of the internally computed one
 *              (P' = 3P-2 with dilation, else P) -- lets tests pin the integer search
 *   sdist_out  [B, S+1]
 *   idx_out    optional int32 [B, S]: interval index #{cw <= u} - 1 (bit-exact contract)
 *   cw_out     optional [B, P'+1]; tdil_out/wdil_out optional [B, P'+1] / [B, P'] (trimmed)
 */
typedef struct {
  int32_t num_rays, num_prev, num_samples;
  int32_t use_dilation;
  float dilation, domain_lo, domain_hi;
  float anneal, resample_padding;
  int32_t jitter_mode;      /* 0 none, 1 per ray, 2 per sample */
  float max_jitter;
} mnrf_sample_desc;

int mnrf_sample_level(const mnrf_sample_desc* d, const float* sdist_prev, const float* w_prev,
                      const float* u_base, const float* jitter, const float* cw_in,
                      float* sdist_out, int32_t* idx_out, float* cw_out, float* tdil_out,
                      float* wdil_out, mnrf_stream stream);
/* Same, but the annealing exponent is read from device memory (`anneal_dev[0]`) at run time so
 * that a captured CUDA graph can be replayed while train_frac advances (models.py:174-179). */
int mnrf_sample_level_dyn(const mnrf_sample_desc* d, const float* sdist_prev, const float* w_prev,
                          const float* u_base, const float* jitter, const float* anneal_dev,
                          float* sdist_out, mnrf_stream stream);

/* ---- ray casting + integrated positional encoding ------------------------------------
 * Replaces coord.construct_ray_warps s_to_t (coord.py:63-99), render.cast_rays
 * (render.py:103-127, diag=False), coord.track_linearize(contract) (coord.py:21-60),
 * coord.lift_and_diagonalize (:129-133) and coord.integrated_pos_enc (:107-126).
 *   sdist [B, S+1]; origins/directions [B,3]; radii/near/far [B]; basis [K,3]
 *   feat_bf16  [B*S, ld_feat] row stride in elements; columns [2KL, feat_cols) zero-filled
 *   feat_f32   optional [B*S, 2KL] (fp32 copy for parity tests)
 *   tdist_out  optional [B, S+1]
 */
enum { MNRF_RAYDIST_NONE = 0, MNRF_RAYDIST_RECIPROCAL, MNRF_RAYDIST_LOG, MNRF_RAYDIST_EXP,
       MNRF_RAYDIST_SQRT, MNRF_RAYDIST_SQUARE, MNRF_RAYDIST_PIECEWISE };
enum { MNRF_RAY_CONE = 0, MNRF_RAY_CYLINDER = 1 };

typedef struct {
  int32_t num_rays, num_samples;
  int32_t raydist_fn, ray_shape, warp_contract, disable_integration;
  int32_t basis_k, min_deg, max_deg;
  int32_t ld_feat, feat_cols;
} mnrf_encode_desc;

int mnrf_encode(const mnrf_encode_desc* d, const float* sdist, const float* origins,
                const float* directions, const float* radii, const float* near,
                const float* far, const float* basis, mnrf_bf16* feat_bf16, float* feat_f32,
                float* tdist_out, mnrf_stream stream);
/* Same, plus the tangent features d(feature)/d(mean_x|y|z) as three stacked bf16 blocks
 * tfeat[dir*B*S + m, ld_tfeat] (input of the forward-mode density-normal chain that replaces
 * vmap(value_and_grad(predict_density)), models.py:473-492).  Contraction is not supported here. */
int mnrf_encode_tangent(const mnrf_encode_desc* d, const float* sdist, const float* origins,
                        const float* directions, const float* radii, const float* near,
                        const float* far, const float* basis, mnrf_bf16* feat_bf16,
                        mnrf_bf16* tfeat_bf16, int32_t ld_tfeat, mnrf_stream stream);

/* View-direction positional encoding, coord.pos_enc (coord.py:136-147) with
 * append_identity, broadcast over the S samples of each ray (models.py:550-554) and
 * written as bf16 into columns [col0, col0 + 3 + 6*deg) of a [B*S, ld] buffer; columns up
 * to col_end are zero-filled. */
int mnrf_viewdir_enc(int32_t num_rays, int32_t num_samples, int32_t deg, const float* viewdirs,
                     mnrf_bf16* out, int32_t ld, int32_t col0, int32_t col_end,
                     mnrf_stream stream);

/* ---- dense layers on tcgen05 ------------------------------------------------------------
 * One Dense layer of models.py:436-437,455-460 (y = act(x W + b)) and its two backward
 * GEMMs, bf16 operands, fp32 accumulation in TMEM.
 *   mode FWD  : out[M,N] bf16 = act(A[M,K] * Bt[N,K]^T + bias[N])           (A, Bt K-major)
 *   mode DGRAD: out[M,N] bf16 = (A[M,K] * Bt[N,K]^T + rowv[M]*colv[N]) masked by mask[M,N]>0
 *               (or by the 1-bit `maskbits`)
 *               (A = dY, Bt = W in [in,out] layout; mask = stored activation; all optional)
 *   mode WGRAD: out[Mo,N] fp32 += A[R,Mo]^T * B[R,N]   (A = X, B = dY, both row-major with
 *               the reduction index R on rows: "MN-major" operands), split over R, fp32 atomics
 * All leading dimensions are in elements.  K (or R) must be a multiple of 64 (16 for R),
 * M-tiles are 128 rows; N must be a multiple of 16.
 */
enum { MNRF_GEMM_FWD = 0, MNRF_GEMM_DGRAD = 1, MNRF_GEMM_WGRAD = 2 };
enum { MNRF_ACT_NONE = 0, MNRF_ACT_RELU = 1 };

typedef struct {
  int32_t mode, act;
  int64_t m;          /* rows of the output (FWD/DGRAD: samples; WGRAD: `in` features) */
  int32_t n, k;       /* output columns; reduction length (WGRAD: number of samples R) */
  int64_t lda, ldb, ldc, ldmask;
  int64_t ldmaskbits; /* row pitch of `maskbits` in 32-bit words */
  int64_t ldadd;      /* row pitch of `addend` */
  int64_t mask_mod;   /* > 0: mask row = output row mod mask_mod (the 3 stacked tangent streams of the
                         density-normal chain share the primal's ReLU masks); 0: mask row = output row */
  int32_t impl;       /* 0 = tcgen05 (product path); 1 = SIMT reference kernel (bring-up/tests) */
} mnrf_gemm_desc;

/* maskbits (optional): 1-bit ReLU masks, word w of row m covers columns [32w, 32w+32).  FWD with
 * MNRF_ACT_RELU writes them (bit = output > 0); DGRAD reads them instead of the bf16 `mask`
 * (16x less mask traffic).  The caller zero-fills nothing: every word of the tile is written.
 * colsum (optional, DGRAD only): colsum[N] += column sums of the output, i.e. the bias gradient
 * of the layer whose activation masks this dgrad; reduced from the epilogue registers (fp32,
 * before the bf16 rounding of the stored output) -- no separate pass over dY.
 * addend (optional, DGRAD only): out += addend[M, ldadd] (bf16), added after the mask -- the second
 * gradient contribution to an input consumed twice (skip connection of the view MLP). */
int mnrf_gemm(const mnrf_gemm_desc* d, const mnrf_bf16* a, const mnrf_bf16* b, const float* bias,
              const float* rowv, const float* colv, const mnrf_bf16* mask, uint32_t* maskbits,
              float* colsum, const mnrf_bf16* addend, void* out, mnrf_stream stream);

/* Weight gradient with side sums computed from the operand tiles the main loop stages (the epilogue warps
 * are idle there), so the bias gradient and the gradient of a Dense(1) head on the same activation cost no
 * extra pass over HBM:
 *   out[Mo,N] += A[R,Mo]^T B[R,N]                       (as mnrf_gemm, mode MNRF_GEMM_WGRAD; A = X, B = dY)
 *   bsum[N]   += sum_r B[r, :]                          (optional: bias gradient of the layer)
 *   side_aw[Mo] += sum_r side_w[r] * A[r, :]            (optional: dW of a Dense(1) head reading X, with
 *                                                        side_w = its d(raw output), models.py:460) */
int mnrf_gemm_wgrad(const mnrf_gemm_desc* d, const mnrf_bf16* a, const mnrf_bf16* b, float* bsum,
                    const float* side_w, float* side_aw, float* out, mnrf_stream stream);

/* ---- layer-chained 256-wide MLP trunk ------------------------------------------------------
 * ONE persistent launch walks 512-row units of samples through all Dense layers of a 256-wide trunk
 * (forward; models.py:441-465 incl. the skip concat) or through its whole input-gradient chain
 * (backward, the dgrad side of jax.value_and_grad, train_utils.py:316-317).  Activations stay in
 * shared memory between layers and are written out once per layer (for the weight-gradient GEMMs);
 * accumulators live in TMEM; weights stream from L2 (csrc/chain.cu).
 *
 * A layer multiplies up to two operands, both in 64-column k-blocks:
 *   resident  -- the previous layer's output held in shared memory: n_res = 4 k-blocks (0 for the
 *                first layer), against weight k-blocks [res_kb0, res_kb0 + 4);
 *   streamed  -- n_stream k-blocks of the `stream` tensor (columns stream_col0 + 64 s), against weight
 *                k-blocks [stream_kb0, stream_kb0 + n_stream): the IPE features of layer 0 and of a
 *                skip layer (forward), the incoming gradient of the first chained layer (backward).
 * `w` is K-major [256, ldw] bf16: the forward operand w_nk [out, in_pad] or the dgrad operand
 * w_kn [in_pad(first 256 rows used), out].
 *   FWD: out = relu(acc + bias) (bf16), maskbits written (1 bit per output, as mnrf_gemm);
 *        head_w/head_b/head_out (optional): head_out[m] = <bf16(out_last[m, :]), head_w> + head_b[0],
 *        the Dense(1) density head of models.py:460 computed in the last layer's epilogue.
 *   BWD: out = acc masked by maskbits (read; NULL = no mask); colsum[256] += column sums of out (the
 *        bias gradient of the layer whose activation the mask came from).
 * `out` may be NULL (FWD only: the activation is not needed later).  m is any row count; rows past m
 * are zero-filled on load and clipped on store.
 */
#define MNRF_CHAIN_MAX_LAYERS 8
enum { MNRF_CHAIN_FWD = 0, MNRF_CHAIN_BWD = 1 };

typedef struct {
  const mnrf_bf16* w;
  int64_t ldw;
  const float* bias;
  uint32_t* maskbits;
  int64_t ldmaskbits;         /* in 32-bit words */
  float* colsum;
  mnrf_bf16* out;
  int64_t ldo;
  int32_t n_stream, stream_col0, stream_kb0;
  int32_t n_res, res_kb0;
  int32_t reserved;
} mnrf_chain_layer;

typedef struct {
  int32_t mode, num_layers;
  int32_t width;              /* must be 256 */
  int32_t stream_cols;        /* columns of `stream` (multiple of 64) */
  int64_t m;                  /* sample rows */
  const mnrf_bf16* stream;    /* [m, ldstream] bf16 */
  int64_t ldstream;
  const float* head_w;        /* [256] fp32 or NULL */
  const float* head_b;        /* device scalar or NULL */
  float* head_out;            /* [m] fp32 */
  mnrf_chain_layer layer[MNRF_CHAIN_MAX_LAYERS];
} mnrf_chain_desc;

int mnrf_mlp_chain(const mnrf_chain_desc* d, mnrf_stream stream);
int mnrf_mlp_chain_max_layers(void);

/* ---- small heads (N <= 4 outputs): density / rgb / predicted normals ---------------------
 * raw[M, n_out] = X[M, K](bf16) * W[n_out, K](bf16) + b, fp32 accumulate; models.py:460,585.
 * Backward: dX[M, K] (bf16, optionally relu-masked by X > 0; optional accumulate is not
 * provided -- the trunk adds the density term through mnrf_gemm's rowv/colv),
 * dW[K, n_out] += (the fp32 master layout [in, out]), db[n_out] += (fp32 atomics);
 * dxsum[K] += column sums of dX (optional: bias gradient of the layer that produced X).
 */
int mnrf_head_fwd(int64_t m, int32_t k, int32_t n_out, const mnrf_bf16* x, int64_t ldx,
                  const mnrf_bf16* w, const float* b, float* raw, mnrf_stream stream);
int mnrf_head_bwd(int64_t m, int32_t k, int32_t n_out, const mnrf_bf16* x, int64_t ldx,
                  const mnrf_bf16* w, const float* draw, mnrf_bf16* dx, int64_t lddx,
                  int32_t relu_mask, float* dw, float* db, float* dxsum, mnrf_stream stream);

/* Column sums of a bf16 matrix into fp32 (bias gradients): out[N] += sum_m x[m, :]. */
int mnrf_colsum(int64_t m, int32_t n, const mnrf_bf16* x, int64_t ldx, float* out,
                mnrf_stream stream);

/* ---- compositing ------------------------------------------------------------------------
 * Forward: density activation (models.py:506) + rgb activation/padding (models.py:584-602)
 * + render.compute_alpha_weights (render.py:130-151) + render.volumetric_rendering
 * (render.py:154-213).  One warp owns one ray.
 *   raw_density [B,S]; raw_rgb [B,S,3] or NULL (PropMLP: disable_rgb -> rgb = 0)
 *   density_noise optional [B,S] N(0,1) draws (models.py:462-464)
 *   sdist [B,S+1]; directions [B,3]; near/far [B]; bg: scalar or NULL->bg_rgb [B,3]
 *   rgb_scale optional [B,3]: per-ray colour scale applied to the sample colours (RawNeRF
 *   exposure_values x learned exposure scaling, models.py:257-267)
 *   outputs: weights [B,S]; rgb_out [B,3]; optional density_out [B,S], rgb_samples [B,S,3]
 *   extras (compute_extras): acc [B], dist [B,4] = (mean, p5, median, p95) or NULL
 */
enum { MNRF_RGB_SIGMOID = 0, MNRF_RGB_SAFE_EXP = 1 };

typedef struct {
  int32_t num_rays, num_samples;
  int32_t raydist_fn, opaque_background;
  float density_bias, density_noise;
  int32_t rgb_act;
  float rgb_premult, rgb_bias, rgb_padding;
  float bg_const;
  int32_t rgb_mode;         /* 0: colour = act(raw_rgb); 1: diffuse + specular (models.py:588-599):
                               clip(linear_to_srgb(tint * act(raw_rgb) + sigmoid(raw_diffuse - log 3)), 0, 1),
                               tint = sigmoid(raw_tint) or 0.5 when raw_tint is NULL */
} mnrf_composite_desc;

int mnrf_composite_fwd(const mnrf_composite_desc* d, const float* raw_density,
                       const float* raw_rgb, const float* density_noise, const float* sdist,
                       const float* directions, const float* near, const float* far,
                       const float* bg_rgb, const float* rgb_scale, const float* raw_diffuse,
                       const float* raw_tint, float* weights, float* rgb_out,
                       float* density_out, float* rgb_samples, float* acc, float* dist,
                       mnrf_stream stream);

/* Losses + compositing backward for one level (train_utils.py:72-159 + the adjoint of
 * render.py:130-213).  Fuses: data loss (mse | charb | rawnerf) on this level's pixel,
 * distortion loss (final level), interlevel loss (proposal levels, against the final
 * level's (sdist, weights)), then the alpha-compositing adjoint, the density-activation
 * and rgb-activation derivatives.
 *   outputs: d_raw_density [B,S]; d_raw_rgb [B,S,3] or NULL; stats[8] += (fp32 atomics):
 *     [0] data loss (already weighted by data_mult)  [1] mse  [2] distortion  [3] interlevel
 */
enum { MNRF_LOSS_MSE = 0, MNRF_LOSS_CHARB = 1, MNRF_LOSS_RAWNERF = 2 };

typedef struct {
  mnrf_composite_desc c;
  int32_t loss_type;
  float charb_padding;
  float data_mult;          /* data_loss_mult (final) or data_coarse_loss_mult (proposal) */
  float distortion_mult;    /* 0 on proposal levels */
  float interlevel_mult;    /* 0 on the final level */
  int32_t num_samples_fine; /* S of the final level (interlevel) */
  int32_t lossmult_channels;/* 1 or 3 */
} mnrf_loss_desc;

int mnrf_composite_bwd(const mnrf_loss_desc* d, const float* raw_density, const float* raw_rgb,
                       const float* density_noise, const float* sdist, const float* directions,
                       const float* near, const float* far, const float* bg_rgb,
                       const float* rgb_scale, const float* raw_diffuse, const float* raw_tint,
                       const float* extra_dw /* [B,S] added to dL/dweights, or NULL */,
                       const float* target_rgb,
                       const float* lossmult, const float* inv_denom /* device scalar */,
                       const float* sdist_fine, const float* weights_fine,
                       float* d_raw_density, float* d_raw_rgb, float* d_rgb_scale /* [B,3] or NULL */,
                       float* d_raw_diffuse, float* d_raw_tint, float* stats, mnrf_stream stream);

/* ---- Ref-NeRF per-sample stage ---------------------------------------------------------------
 * Between the spatial trunk and the directional MLP: normals_pred / normals = -l2_normalize(.)
 * (models.py:488-499, ref_utils.py:40-42), roughness = softplus(raw + bias) (models.py:520-523),
 * refdirs = reflect(-viewdirs, normals) (ref_utils.py:22-37, models.py:545), the integrated
 * directional encoding (ref_utils.generate_ide_fn ref_utils.py:98-159; ide_mat [l_max+1, ide_n]
 * fp32 and ide_ml int32 [2, ide_n] = (m | l) come from multinerf_b200/ref_utils.py) or coord.pos_enc,
 * and n.v (models.py:560-563).  Writes the bf16 slab [col0, col_end) of the view-MLP input.
 * raw_grad_density / d_raw_grad_density are direction-major [3, M].
 * Backward adds train_utils.orientation_loss (:162-178) and predicted_normal_loss (:181-197):
 * orient_mult / prednorm_mult are the level's multipliers divided by the number of rays;
 * extra_dw [M] (forward) receives d(loss)/d(weights) of those two terms; stats[4], stats[5] their
 * values.  After consuming d_slab[:, col0:col_end) (the gradient of the encoding), the backward
 * overwrites those columns with the 11 head gradients (raw_density, grad_pred x3, raw_diffuse x3,
 * raw_tint x3, raw_roughness; bf16, zero-padded) so that [bottleneck grad | head grads] is the A
 * operand of a single dgrad GEMM into the trunk.
 */
typedef struct {
  int64_t M;
  int32_t num_samples;
  int32_t use_pred_normals, use_density_normals, use_reflections, use_ide, use_n_dot_v, use_roughness;
  int32_t deg_view, ide_n;
  float roughness_bias;
  int32_t ld, col0, col_end;
} mnrf_refdir_desc;

int mnrf_refdir_fwd(const mnrf_refdir_desc* d, const float* ide_mat, const int32_t* ide_ml,
                    const float* grad_pred, const float* raw_rough, const float* raw_grad_density,
                    const float* viewdirs, float* normals_pred, float* normals, float* roughness,
                    mnrf_bf16* slab, float orient_mult, float prednorm_mult, int32_t orient_on_pred,
                    float* extra_dw /* [M] or NULL */, mnrf_stream stream);
int mnrf_refdir_bwd(const mnrf_refdir_desc* d, const float* ide_mat, const int32_t* ide_ml,
                    const float* grad_pred, const float* raw_rough, const float* raw_grad_density,
                    const float* viewdirs, const float* weights, mnrf_bf16* d_slab,
                    int32_t ld_dslab, float orient_mult, float prednorm_mult, int32_t orient_on_pred,
                    const float* d_raw_density, const float* d_raw_diffuse, const float* d_raw_tint,
                    float* d_grad_pred, float* d_raw_rough, float* d_raw_grad_density,
                    float* stats, mnrf_stream stream);
/* out[r, n] bf16 = maskbit(r mod mask_mod, n) ? rowv[r] * colv[n] : 0 -- the first dY of the
 * density-normal (tangent) backward chain. */
int mnrf_outer_mask(int64_t rows, int32_t n, int64_t mask_mod, const float* rowv, const float* colv,
                    const uint32_t* maskbits, int64_t ldmaskbits, mnrf_bf16* out, int64_t ldo,
                    mnrf_stream stream);

/* ---- ray generation (the step before the path) --------------------------
 * camera_utils.pixels_to_rays (camera_utils.py:522-636) + the camera gather of
 * camera_utils.cast_ray_batch (:639-688), as the reference runs it on the device when
 * Config.cast_rays_in_train_step is set (train_utils.py:266-268): half-pixel offset, inverse
 * intrinsics, optional radial/tangential undistortion (_radial_and_tangential_undistort
 * :478-513, Newton steps), optional fisheye model, OpenCV->OpenGL flip, camera rotation,
 * optional NDC projection (convert_to_ndc :32-97), mip-NeRF cone radii from the dx/dy neighbours.
 * pixtocams [num_cameras, 3, 3] and camtoworlds [num_cameras, 3, 4] are row-major fp32;
 * cam_idx may be NULL when num_cameras == 1.  Outputs are [num_rays, 3|3|3|1|2] fp32.
 */
#define MNRF_CAM_PERSPECTIVE 0
#define MNRF_CAM_FISHEYE 1
typedef struct {
  int32_t num_rays;
  int32_t num_cameras;
  int32_t camtype;
  int32_t has_distortion;
  float k1, k2, k3, k4, p1, p2;
  float undistort_eps;        /* 1e-9 in the reference */
  int32_t undistort_iters;    /* 10 in the reference */
  int32_t has_ndc;
  float ndc_p02, ndc_p12;     /* pixtocam_ndc[0][2], pixtocam_ndc[1][2] */
  float ndc_near;             /* 1.0 in the reference */
} mnrf_camera_desc;

int mnrf_pixels_to_rays(const mnrf_camera_desc* d, const int32_t* pix_x, const int32_t* pix_y,
                        const int32_t* cam_idx, const float* pixtocams, const float* camtoworlds,
                        float* origins, float* directions, float* viewdirs, float* radii,
                        float* imageplane, mnrf_stream stream);

/* ---- optimizer ---------------------------------------------------------------------------
 * train_utils.clip_gradients (train_utils.py:200-218: value clip, then global-norm clip with
 * eps in the denominator), nan_to_num (:328) and optax.adam on one flat fp32 parameter
 * group (one top-level module).  norm_sq_scratch: device float[1], zeroed by the call.
 */
typedef struct {
  int64_t n;
  float grad_max_val, grad_max_norm;
  float lr, beta1, beta2, eps;
  int32_t step;             /* 1-based update count t */
  float grad_scale;         /* multiplies the raw gradient first (1/world_size for pmean) */
} mnrf_adam_desc;

int mnrf_clip_adam(const mnrf_adam_desc* d, float* params, const float* grads, float* mu,
                   float* nu, float* norm_sq_scratch, mnrf_stream stream);
/* Same, but (lr, 1-beta1^t, 1-beta2^t) are read from device memory `dyn[0..2]` (graph replay). */
int mnrf_clip_adam_dyn(const mnrf_adam_desc* d, float* params, const float* grads, float* mu,
                       float* nu, float* norm_sq_scratch, const float* dyn, mnrf_stream stream);

/* fp32 master [in_pad, out] (row-major) -> bf16 shadows: w_nk [out, in_pad] (K-major operand
 * of the forward GEMM) and w_kn [in_pad, out] (K-major operand of the dgrad GEMM). */
int mnrf_pack_weights(int32_t in_pad, int32_t out, const float* master, mnrf_bf16* w_nk,
                      mnrf_bf16* w_kn, mnrf_stream stream);

/* The same for `count` layers in one launch (the optimizer epilogue of a train step).  `items` is a
 * DEVICE array of mnrf_pack_item, ordered as the layers' 32x32 tiles are numbered: item i owns tiles
 * [tile0, tile0 + ceil(out/32) * ceil(in_pad/32)), tile0 of item i+1 = the end of item i;
 * total_tiles = the end of the last item. */
typedef struct {
  const float* master;
  mnrf_bf16* w_nk;          /* may be NULL */
  mnrf_bf16* w_kn;          /* may be NULL */
  int32_t in_pad, out;
  int32_t tile0;
  int32_t reserved;
} mnrf_pack_item;

int mnrf_pack_weights_batched(int32_t count, const mnrf_pack_item* items, int32_t total_tiles,
                              mnrf_stream stream);

#ifdef __cplusplus
}
#endif
#endif  /* MNRF_H_ */
