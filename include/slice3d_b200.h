/*
 * slice3d_b200 -- C ABI of the B200-native slice-to-3D hot path.
 *
 * The reference (yizhiwang96/Slice3D, reg_slices/) has NO native boundary on this
 * path: everything is PyTorch library calls made from Slices3DRegModel.forward
 * (reg_slices/src/models.py:48-94) and Generator3D.eval_points
 * (reg_slices/reconstruct.py:74-102).  The entry points below are what a
 * maintainer would bind (ctypes) to replace those calls:
 *
 *   s3d_model_create      <- Slices3DRegModel.load_state_dict + .cuda().eval()
 *                            (reconstruct.py:343-345): takes the tensors of the
 *                            244-key state_dict by name, keeps folded/packed copies.
 *   s3d_encoder_fwd       <- self.slices_generator(img_input)          (models.py:65,
 *                            unet_custom.py:40-69) + the fc_s projection hoisted onto
 *                            the planes (models.py:80; exact, bilinear sampling is linear).
 *   s3d_decoder_fwd       <- query flip / rotation (models.py:53-60), project_coord
 *                            (:28-36,69), sample_from_planes x5 (:38-46,71-78),
 *                            fc_p/fc_s (:79-80), att_decoder (:82-83), fc_out (:84),
 *                            and the negation done by eval_points (reconstruct.py:97).
 *   s3d_decoder_grid_fwd  <- the same over a make_3d_grid slab without materialising
 *                            the (nx*ny*nz,3) point tensor (src_convonet/common.py:145-164,
 *                            reconstruct.py:137-146).
 *   s3d_vgg_loss_fwd      <- self.vggptlossfunc(slices_rec, img_slices) (models.py:90-92,
 *                            vgg_perceptual_loss.py:51-71), evaluated in every forward.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is a CUDA device pointer on the
 *     device the model was created on; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), no implicit synchronisation.
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library owns
 *     only the packed weight copies inside s3d_model, freed by s3d_model_destroy.
 *   - return value: 0 on success, negative S3D_ERR_* otherwise; s3d_last_error()
 *     returns a thread-local message.  Nothing throws across the ABI.
 *   - one s3d_model per device; calls on one model must be externally serialised
 *     (the Python host holds the GIL).
 */
#ifndef SLICE3D_B200_H
#define SLICE3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3D_ABI_VERSION 1

#define S3D_OK 0
#define S3D_ERR_BAD_ARG (-1)
#define S3D_ERR_MISSING_TENSOR (-2)
#define S3D_ERR_CUDA (-3)
#define S3D_ERR_WORKSPACE (-4)
#define S3D_ERR_UNSUPPORTED (-5)

/* Decoder arithmetic.  All modes accumulate in fp32. */
#define S3D_PREC_FP32 0   /* CUDA-core fp32 (validation grade, slow)                     */
#define S3D_PREC_BF16X3 1 /* tcgen05, operands split into bf16 hi+lo, 3 MMA passes (<=1e-4) */
#define S3D_PREC_BF16 2   /* tcgen05, single bf16 pass (fast; measured 1.8e-2 max-abs)    */
#define S3D_PREC_FP16X3 3 /* tcgen05, operands split into fp16 hi+lo, 3 MMA passes: same speed as BF16X3, ~10x
                             smaller error (22 instead of 16 mantissa bits); decoder activations must stay below 65504 */
#define S3D_PREC_FP16F8 4 /* as FP16X3, but the two cross terms (x_lo.w_hi, x_hi.w_lo) of the FFN contractions -- 88 % of the
                             MMA work -- run as scaled E4M3 products on kind::f8f6f4 at twice the rate; the leading term
                             stays fp16.  ~5e-5 max-abs; FFN activations must stay below 511 */

typedef struct s3d_model s3d_model;

/* One tensor of the checkpoint: state_dict key, device pointer, element count.
 * dtype is float32 for every key the library reads (int64 num_batches_tracked is ignored). */
typedef struct {
  const char* name;
  const void* data_dev;
  int64_t numel;
} s3d_tensor;

/* Query grid descriptor: point (ix,iy,iz) = (px[ix], py[iy], pz[iz]), flat index
 * (ix*ny+iy)*nz+iz (x slowest, z fastest, as make_3d_grid).  px/py/pz are device
 * arrays holding the per-axis torch.linspace values times box_size. */
typedef struct {
  int32_t nx, ny, nz;
  const float* px_dev;
  const float* py_dev;
  const float* pz_dev;
} s3d_grid;

int s3d_abi_version(void);
const char* s3d_last_error(void);

/* Create / destroy.  `tensors` must contain every key under slices_generator.*,
 * att_decoder.*, fc_p.*, fc_s.*, fc_out.* (att_layer.* and num_batches_tracked are not read; vggptlossfunc.* is
 * optional) -- or the keys of a Slices3DGTModel checkpoint, see s3d_gt_encoder_fwd.  n_slices = K (12 in the reference). */
int s3d_model_create(s3d_model** out, const s3d_tensor* tensors, int32_t n_tensors, int32_t n_slices,
                     int32_t device, void* stream);
void s3d_model_destroy(s3d_model* m);
int s3d_model_n_slices(const s3d_model* m);

/* Encoder.  img_dev: (B,3,S,S) fp32 NCHW, S a multiple of 16.
 *   planes_dev      out, s3d_planes_bytes(B,K,S) bytes: per image, per scale s=0..4,
 *                   (K, S/16*2^s, S/16*2^s, 128) fp32 channels-last = fc_s_s . plane_s.
 *   feats_nchw_dev  optional (may be NULL, entries may be NULL): the five raw feature
 *                   planes (B*K, C_s, H_s, W_s) fp32 NCHW, C = 512,256,128,64,32.
 *   slices_rec_dev  optional: (B*K,3,S,S) fp32 NCHW, tanh output.
 *   workspace_dev   s3d_encoder_workspace_bytes(B,K,S) bytes, uninitialised, owned by the caller for the duration of
 *                   the call (activations + 16 MB of split-K partial tiles and their counters); two calls that may
 *                   overlap on different streams need two workspaces.
 */
size_t s3d_planes_bytes(int32_t B, int32_t K, int32_t S);
size_t s3d_encoder_workspace_bytes(int32_t B, int32_t K, int32_t S);
int s3d_encoder_fwd(const s3d_model* m, const float* img_dev, int32_t B, int32_t S, void* planes_dev,
                    float* const* feats_nchw_dev, float* slices_rec_dev, void* workspace_dev,
                    size_t workspace_bytes, void* stream);

/* Decoder over explicit points.  For image `b` of the encoder batch pass
 * planes_dev + b * s3d_planes_bytes(1,K,S).
 *   qry_dev    (n,3) fp32 query points (qry_norot).
 *   T_dev      (4,3) fp32 trans_mat_wo_rot_tp.
 *   rot_dev    (3,3) fp32 obj_rot_mat or NULL.  NULL => test mode: y,z are negated
 *              (models.py:55); with flip_in_place != 0 the negated values are also
 *              written back to qry_dev, reproducing the reference's in-place side effect.
 *   out_dev    (n) fp32 = out_scale * sdf_pred  (out_scale = -1 gives eval_points' values).
 */
size_t s3d_decoder_workspace_bytes(int64_t n, int32_t precision);
int s3d_decoder_fwd(const s3d_model* m, const void* planes_dev, int32_t S, float* qry_dev, int64_t n,
                    const float* T_dev, const float* rot_dev, int32_t flip_in_place, float out_scale,
                    float* out_dev, int32_t precision, void* workspace_dev, size_t workspace_bytes,
                    void* stream);

/* The same for a batch: B images whose planes lie back to back in planes_dev (the encoder's batch layout), qry_dev
 * (B, n_per_image, 3), T_dev (B,4,3), rot_dev (B,3,3) or NULL, out_dev (B, n_per_image): ONE launch for the whole
 * feed_dict of Slices3DRegModel.forward (models.py:48-94 with n_bs > 1, e.g. train.py:val_step). */
int s3d_decoder_batch_fwd(const s3d_model* m, const void* planes_dev, int32_t S, float* qry_dev, int32_t B,
                          int64_t n_per_image, const float* T_dev, const float* rot_dev, int32_t flip_in_place,
                          float out_scale, float* out_dev, int32_t precision, void* workspace_dev,
                          size_t workspace_bytes, void* stream);

/* Decoder over grid points [first, first+count) of `grid` (test-mode y,z flip applied
 * on the fly).  out_dev receives `count` values. */
int s3d_decoder_grid_fwd(const s3d_model* m, const void* planes_dev, int32_t S, const s3d_grid* grid,
                         int64_t first, int64_t count, const float* T_dev, float out_scale, float* out_dev,
                         int32_t precision, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Validation aid (fp32 path only): like s3d_decoder_fwd with out_scale = 1 and no in-place
 * flip, and additionally writes the token matrices (K+1 rows of 128 per query) after the
 * token build and after each of the three attention layers to tokens_dev, laid out
 * [4][n][K+1][128] fp32.  Workspace as for S3D_PREC_FP32. */
int s3d_decoder_debug_tokens(const s3d_model* m, const void* planes_dev, int32_t S, const float* qry_dev,
                             int64_t n, const float* T_dev, const float* rot_dev, float* out_dev,
                             float* tokens_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Hardware self-test of the tensor-core plumbing the decoder relies on (UMMA shared-memory and
 * instruction descriptors, 128-byte-swizzled operand tiles, bulk async copy, TMEM load): one
 * 128-row tile against one weight unit, passes = 1 (bf16), 3 (bf16 hi/lo split) or 4 (= three passes over
 * fp16 hi/lo pairs, the S3D_PREC_FP16X3 operands).
 *   d[128][128] = a[128][128] . w[128][128]^T;  mode 0: A operand in shared memory, mode 1: A operand in
 *   tensor memory (the two operand paths of the decoder).
 * a_dev, w_dev, d_dev are fp32 row-major device arrays.  Synchronises the stream. */
int s3d_selftest_umma(int32_t mode, int32_t passes, const float* a_dev, const float* w_dev, float* d_dev,
                      void* stream);

/* VGG19 perceptual loss between two image batches a, b (N,3,S,S) fp32 in [-1,1] (NCHW, device):
 * replaces VGGPerceptualLoss.forward (reg_slices/src/vgg_perceptual_loss.py:51-71), which the reference
 * evaluates inside every Slices3DRegModel.forward, also at test time (reg_slices/src/models.py:90-92).
 * loss_dev receives ONE float = sum_t w_t * mean|tap_t(a) - tap_t(b)| (the caller applies the 0.001 of
 * models.py:92).  Needs a model created WITH the vggptlossfunc.* tensors (S3D_ERR_MISSING_TENSOR otherwise);
 * S must be a multiple of 16. */
size_t s3d_vgg_loss_workspace_bytes(int32_t N, int32_t S);
int s3d_vgg_loss_fwd(const s3d_model* m, const float* a_dev, const float* b_dev, int32_t N, int32_t S, float* loss_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- Training (reg_slices/train.py:41-53: model(batch) under model.train(), loss.backward()) ----------------------
 * Forward and backward of the per-query half of Slices3DRegModel.forward in TRAIN mode (reg_slices/src/models.py:57-84):
 * project_coord, 5 x grid_sample of the NCHW feature planes the U-Net produced, fc_s / fc_p, the 3-layer
 * nn.TransformerEncoder WITH dropout (p = dropout_p on the attention weights, after the attention block, inside the FFN
 * and after it), fc_out on token 0 -- exact fp32, PyTorch parameter layouts used in place (no packing: the optimizer
 * rewrites them every step).  The U-Net / VGG19 convolutions keep their autograd on the caller's side.
 *   feats_dev[5]   (B*K, C_s, R_s, R_s) fp32 NCHW, C = 512,256,128,64,32, R_s = S/16 * 2^s.
 *   qry_dev        (B, n_qry, 3) fp32, already rotated (bmm with obj_rot_mat, models.py:60) or flipped.
 *   T_dev          (B, 4, 3) fp32.
 *   params_dev[42] fc_p.weight, fc_p.bias, fc_s.weight, fc_s.bias, then for each layer l = 0..2 of att_decoder.layers:
 *                  self_attn.in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias, linear1.weight, linear1.bias,
 *                  linear2.weight, linear2.bias, norm1.weight, norm1.bias, norm2.weight, norm2.bias; fc_out.0.weight, bias.
 *   saved_dev      s3d_train_decoder_saved_bytes(): activations kept for the backward pass (caller-owned).
 *   bwd: dsdf_dev (B, n_qry); dfeats_dev[5] ZERO-INITIALISED gradients of the feature planes (accumulated with
 *        atomics: queries share texels); dparams_dev[42] gradients, overwritten; workspace of
 *        s3d_train_decoder_bwd_workspace_bytes().  Dropout masks are regenerated from (seed, site, element). */
typedef struct {
  int32_t B, n_qry, K, S;
  float dropout_p;
  uint64_t seed;
} s3d_train_cfg;
size_t s3d_train_decoder_saved_bytes(const s3d_train_cfg* cfg);
size_t s3d_train_decoder_bwd_workspace_bytes(const s3d_train_cfg* cfg);
int s3d_train_decoder_fwd(const s3d_train_cfg* cfg, const float* const* feats_dev, const float* qry_dev, const float* T_dev,
                          const float* const* params_dev, float* sdf_dev, void* saved_dev, size_t saved_bytes, void* stream);
int s3d_train_decoder_bwd(const s3d_train_cfg* cfg, const float* qry_dev, const float* T_dev, const float* const* params_dev,
                          const float* dsdf_dev, void* saved_dev, size_t saved_bytes, float* const* dfeats_dev,
                          float* const* dparams_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- Slices3DGTModel (reg_slices/src/model_gt.py:12-111) ---------------------------------------------------------------
 * s3d_model_create recognises a GT checkpoint by its keys (img_encoder.*, fc_local.*, pts_feat_extractor.*, att_decoder.*,
 * fc_out.*) and returns a handle for the two entry points below; the regression entry points refuse it.
 *   s3d_gt_encoder_fwd   <- self.img_encoder(img_slices) (model_gt.py:81-84, vgg16bn_feats.py:44-58) + the first Linear of
 *                           fc_local hoisted onto the five taps (model_gt.py:97): img_slices_dev (B*K,3,S,S) fp32 NCHW ->
 *                           planes_dev in the layout of s3d_encoder_fwd (scale s holds tap 4 - s); taps_nchw_dev optional:
 *                           the raw taps conv1_2 .. conv5_3 (B*K, 64..512, S..S/16) for validation.
 *   s3d_gt_decoder_fwd   <- model_gt.py:69-79 (flip / rotation), :86-104 (projection, 5 x grid_sample, fc_local,
 *                           pts_feat_extractor, att_decoder, fc_out); arguments as s3d_decoder_batch_fwd. */
size_t s3d_gt_encoder_workspace_bytes(int32_t B, int32_t K, int32_t S);
int s3d_gt_encoder_fwd(const s3d_model* m, const float* img_slices_dev, int32_t B, int32_t S, void* planes_dev,
                       float* const* taps_nchw_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
size_t s3d_gt_decoder_workspace_bytes(int64_t n, int32_t precision);
int s3d_gt_decoder_fwd(const s3d_model* m, const void* planes_dev, int32_t S, float* qry_dev, int32_t B, int64_t n_per_image,
                       const float* T_dev, const float* rot_dev, int32_t flip_in_place, float out_scale, float* out_dev,
                       int32_t precision, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- Input pipeline (reg_slices/src/datasets.py:37,75-118) ----------------------------------------------------------
 * A batch of decoded RGBA images (N, H, W, 4) uint8 -> (N, 3, S, S) fp32 exactly as Slice3DDataset prepares them:
 * alpha compositing (png_2_whitebg when white_bg != 0, else png_2_rgb), Pillow's antialiased bilinear resize of the
 * 8-bit image to S x S (T.Resize on a PIL image) bit for bit, T.ToTensor, T.Normalize(0.5, 0.5).
 *   bounds_*_dev  int32 [S][2] = (first input index, tap count) per output column (h) / row (v);
 *   kk_*_dev      int32 [S][ksize] fixed-point (22-bit) filter coefficients -- Pillow's precompute_coeffs +
 *                 normalize_coeffs_8bpc, evaluated in float64 on the host (slice3d_b200/inputs.py:resample_tables). */
size_t s3d_preprocess_workspace_bytes(int32_t N, int32_t H, int32_t S);
int s3d_preprocess_rgba(const uint8_t* rgba_dev, int32_t N, int32_t H, int32_t W, int32_t S, int32_t white_bg,
                        const int32_t* bounds_h_dev, const int32_t* kk_h_dev, int32_t ksize_h, const int32_t* bounds_v_dev,
                        const int32_t* kk_v_dev, int32_t ksize_v, float* out_dev, void* workspace_dev, size_t workspace_bytes,
                        void* stream);

/* The same loss in TRAINING (reg_slices/train.py:41-53 back-propagates through ret['vgg_loss']): forward keeping the
 * activations, and the gradient with respect to the FIRST batch a (the network is frozen: data gradients only, each a
 * 3x3 convolution with the rotated / transposed weights on the same tcgen05 kernel).  A handle created from the
 * vggptlossfunc.* tensors ALONE is enough (and is what training uses: those weights never change).
 *   saved_dev   s3d_vgg_loss_train_bytes(N, S) bytes, kept by the caller between forward and backward.
 *   gout_dev    one float: the upstream gradient of the loss value (e.g. 0.001 from models.py:92, times d(total)/d(term)).
 *   grad_a_dev  (N,3,S,S) fp32 NCHW, overwritten. */
size_t s3d_vgg_loss_train_bytes(int32_t N, int32_t S);
int s3d_vgg_loss_train_fwd(const s3d_model* m, const float* a_dev, const float* b_dev, int32_t N, int32_t S, float* loss_dev,
                           void* saved_dev, size_t saved_bytes, void* stream);
int s3d_vgg_loss_train_bwd(const s3d_model* m, int32_t N, int32_t S, const float* gout_dev, void* saved_dev, size_t saved_bytes,
                           float* grad_a_dev, void* stream);

/* One MISE refinement step on dense device state, replacing MISE.subdivide_voxels (reg_slices/src_convonet/utils/
 * libmise/mise.pyx:184-283) after the caller has stored the new values: R = resolution0 << depth; value_dev / known_dev /
 * exists_dev are (R+1)^3 lattice arrays (float64 / bytes), cell_level_dev is the R^3 int8 array "level of the leaf voxel
 * containing this unit cell" (the octree), flags_dev a scratch of s3d_mise_scratch_ints() int32 that is zero on entry and
 * zero again on return.  Every leaf voxel below `depth` that is next to a known value >= threshold AND a known value <=
 * threshold is split (its cells move one level down, the 27 lattice points of its children start to exist). */
size_t s3d_mise_scratch_ints(int32_t resolution0, int32_t depth);
int s3d_mise_subdivide(int32_t resolution0, int32_t depth, double threshold, const double* value_dev, const uint8_t* known_dev,
                       int8_t* cell_level_dev, uint8_t* exists_dev, int32_t* flags_dev, void* stream);

/* `n_rounds` whole MISE rounds enqueued back to back with NO host round trip, replacing the loop of reconstruct.py:147-167
 * (query -> points to the device -> eval_points -> values to the host -> update): per round a deterministic compaction of
 * the lattice points that exist without a value (flat-index order), ONE tensor-core decoder launch whose query count is
 * read from device memory, the value store and s3d_mise_subdivide.  State arrays as for s3d_mise_subdivide.
 *   scratch_dev   s3d_sparse_scratch_bytes(resolution0, depth, capacity) bytes: block counts, point indices, points, values.
 *   counts_dev    int32 [n_rounds + 2]: entry r receives the number of points round r asked for (0 = the refinement had
 *                 already converged: the round was a no-op); [n_rounds] = live count, [n_rounds + 1] = overflow flag (a
 *                 round asked for more than `capacity` points: the state is then incomplete and the caller must re-run).
 * precision must be a tensor-core mode; decoder workspace as for s3d_decoder_fwd with n = capacity. */
size_t s3d_sparse_scratch_bytes(int32_t resolution0, int32_t depth, int64_t capacity);
int s3d_sparse_rounds(const s3d_model* m, const void* planes_dev, int32_t S, const float* T_dev, double box_size,
                      float out_scale, int32_t resolution0, int32_t depth, double threshold, double* value_dev,
                      uint8_t* known_dev, int8_t* cell_level_dev, uint8_t* exists_dev, int32_t* flags_dev, void* scratch_dev,
                      int64_t capacity, int32_t* counts_dev, int32_t n_rounds, int32_t precision, void* workspace_dev,
                      size_t workspace_bytes, void* stream);

/* Marching cubes over a float64 volume (nx,ny,nz), replacing libmcubes.marching_cubes (reg_slices/reconstruct.py:190;
 * src_convonet/utils/libmcubes/marchingcubes.h:22-196, pywrapper.cpp:90-128) in two passes around the caller's
 * exclusive prefix sums (the running vertex / triangle counters of the sequential reference):
 *   s3d_mc_count  per cell (x-major, (nx-1)(ny-1)(nz-1) of them): vcount = vertices the cell creates, tcount =
 *                 tri_count_dev[configuration] (256 entries), owned = which of its edges 6, 5, 10 are crossed.
 *   s3d_mc_emit   vbase / tbase = exclusive prefix sums of vcount / tcount; table_dev = 256 x 15 edge ids per
 *                 configuration (5 triangles, unused entries ignored); writes verts (n,3) float64 -- the reference's
 *                 vertex array bit for bit, in its order -- and tris (m,3) int64. */
int s3d_mc_count(const double* vol_dev, int32_t nx, int32_t ny, int32_t nz, double isovalue, const int32_t* tri_count_dev,
                 int32_t* vcount_dev, int32_t* tcount_dev, uint8_t* owned_dev, void* stream);
int s3d_mc_emit(const double* vol_dev, int32_t nx, int32_t ny, int32_t nz, double isovalue, const int8_t* table_dev,
                const int64_t* vbase_dev, const int64_t* tbase_dev, const int32_t* tcount_dev, const uint8_t* owned_dev,
                double* verts_dev, int64_t* tris_dev, void* stream);

/* Exclusive prefix sum of n int32 counts into int64 offsets (out_dev[i] = in[0] + ... + in[i-1]; *total_dev = the sum): the
 * running vertex / triangle counters of the sequential marching cubes between s3d_mc_count and s3d_mc_emit.  Three
 * kernels (block sums, one-block scan of the sums, ranked writes); scratch of s3d_scan_scratch_bytes(n). */
size_t s3d_scan_scratch_bytes(int64_t n);
int s3d_exclusive_scan(const int32_t* in_dev, int64_t n, int64_t* out_dev, int64_t* total_dev, void* scratch_dev, void* stream);

/* Debugging aid (tools/enc_check.py): simt != 0 makes s3d_encoder_fwd run the whole encoder on the fp32 CUDA-core GEMM. */
int s3d_debug_set_encoder(s3d_model* m, int32_t simt);
/* Timing experiments on the tensor-core decoder (tools/dec_bench.py): bit 0 = the weight producer skips its copies (the
 * results are garbage; shows what the L2 -> shared-memory weight stream costs).  0 = normal operation. */
int s3d_debug_set_decoder_flags(int32_t flags);

/* Instrumentation of the tensor-core decoder: 32 cycle counters (clock64 deltas summed over CTAs since the
 * last reset; index meaning in slice3d_b200/_native.py PROFILE_FIELDS).  Synchronises the device. */
int s3d_debug_profile(int64_t* out32, int32_t reset);

/* Instrumentation: number of kernels this library has launched since load (all models). */
int64_t s3d_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SLICE3D_B200_H */
