/*
 * lgrast.h -- C-ABI of liblgrast.so, the B200 (sm_100a) rasterizer for LightGaussian's hot path.
 *
 * Drop-in boundary.  The four work entry points replace, one for one, the static C++ API the
 * reference's torch binding calls (RAST = submodules/compress-diff-gaussian-rasterization):
 *
 *   lgr_forward        <- CudaRasterizer::Rasterizer::forward       RAST/cuda_rasterizer/rasterizer.h:35-57
 *   lgr_forward_count  <- CudaRasterizer::Rasterizer::forwardCount  RAST/cuda_rasterizer/rasterizer.h:60-84
 *   lgr_backward       <- CudaRasterizer::Rasterizer::backward      RAST/cuda_rasterizer/rasterizer.h:86-112
 *   lgr_mark_visible   <- CudaRasterizer::Rasterizer::markVisible   RAST/cuda_rasterizer/rasterizer.h:28-33
 *
 * Conventions kept from the reference: every data pointer is a DEVICE pointer to contiguous float32 /
 * int32 memory owned by the caller; a NULL pointer means "input absent" (shs / colors_precomp /
 * scales / rotations / cov3D_precomp); viewmatrix and projmatrix are the TRANSPOSED 4x4 matrices
 * (scene/cameras.py:70-84); out_color is planar [3,H,W]; the three opaque state blobs are obtained
 * through caller-supplied allocators in the order geometry -> image -> binning (the last one only
 * after the instance count is known) and are handed back unchanged to lgr_backward.
 *
 * Differences (all additive): plain function-pointer allocators instead of std::function, an explicit
 * cudaStream_t (the reference launches on the legacy default stream), an int status return with
 * lgr_last_error(), and gradient outputs that need NOT be zero-initialised by the caller.
 *
 * No torch / C++ types cross this boundary.
 */
#ifndef LGRAST_H_INCLUDED
#define LGRAST_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGR_ABI_VERSION 2

/* status codes */
#define LGR_OK 0
#define LGR_ERR_INVALID_ARG 1 /* bad sizes / required pointer missing */
#define LGR_ERR_CUDA 2        /* a CUDA runtime call or kernel failed; see lgr_last_error() */
#define LGR_ERR_ALLOC 3       /* an allocator callback returned NULL */

/* Replaces std::function<char*(size_t)> (RAST/cuda_rasterizer/rasterizer.h:36-38): must return a device
 * pointer to at least `bytes` bytes, aligned to 256 B, that stays valid until the matching backward call. */
typedef char* (*lgr_alloc_fn)(void* user, size_t bytes);

/* Per-view constants: the numeric fields of GaussianRasterizationSettings
 * (RAST/diff_gaussian_rasterization/__init__.py:248-261). */
typedef struct lgr_view {
    int32_t image_width;
    int32_t image_height;
    float tan_fovx;
    float tan_fovy;
    float scale_modifier;
    int32_t sh_degree;       /* active degree D (0..3) */
    int32_t prefiltered;     /* as the reference: a culled point traps the kernel when set */
    int32_t debug;           /* synchronise and check after every stage */
    const float* viewmatrix; /* device, 16 floats, transposed W2C */
    const float* projmatrix; /* device, 16 floats, transposed Proj*W2C */
    const float* campos;     /* device, 3 floats */
    const float* background; /* device, 3 floats */
} lgr_view;

/* Forward render.  P Gaussians, M stored SH coefficients per channel (0 when shs == NULL).
 * Writes out_color[3*H*W], radii[P]; *num_rendered receives the reference's instance count
 * (sum over Gaussians of the tile-rectangle area, rasterizer_impl.cu:278-282). */
int lgr_forward(const lgr_view* view, int P, int M,
                const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                const float* scales, const float* rotations, const float* cov3D_precomp,
                lgr_alloc_fn geometry_alloc, void* geometry_user,
                lgr_alloc_fn binning_alloc, void* binning_user,
                lgr_alloc_fn image_alloc, void* image_user,
                float* out_color, int32_t* radii, int32_t* num_rendered, void* cuda_stream);

/* Forward render + Global Significance accumulation (RAST/cuda_rasterizer/forward.cu:378-501).
 * gaussians_count[P] and important_score[P] are fully written (no zero-init needed):
 *   gaussians_count[i] = number of (pixel, i) pairs that were blended in this view (exact, deterministic),
 *   important_score[i] = opacity[i] * gaussians_count[i]. */
int lgr_forward_count(const lgr_view* view, int P, int M,
                      const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                      const float* scales, const float* rotations, const float* cov3D_precomp,
                      lgr_alloc_fn geometry_alloc, void* geometry_user,
                      lgr_alloc_fn binning_alloc, void* binning_user,
                      lgr_alloc_fn image_alloc, void* image_user,
                      float* out_color, int32_t* gaussians_count, float* important_score, int32_t* radii,
                      int32_t* num_rendered, void* cuda_stream);

/* Backward.  The three blobs and num_rendered come from the matching lgr_forward call.
 * Outputs (all fully written): dL_dmeans2D[P,3], dL_dcolors[P,3], dL_dopacity[P], dL_dmeans3D[P,3],
 * dL_dcov3D[P,6], dL_dsh[P,M,3] (may be NULL when M == 0), dL_dscales[P,3], dL_drotations[P,4]
 * -- the tuple RAST/rasterize_points.cu:298 returns. */
int lgr_backward(const lgr_view* view, int P, int M, int num_rendered,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, const float* rotations, const float* cov3D_precomp,
                 const int32_t* radii, char* geometry_blob, char* binning_blob, char* image_blob,
                 const float* dL_dout_color,
                 float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drotations, void* cuda_stream);

/* ---- fused-activation variants (used inside gaussian_renderer.render()) ----
 * The six parameter leaves of GaussianModel (scene/gaussian_model.py:46-56) are read directly and the activations of
 * its getters (:98-118: exp, normalize, sigmoid, cat) are applied in-kernel, bit-identically to the torch CUDA ops, so
 * the result equals lgr_forward on the activated tensors.  M counts ALL SH coefficients per channel
 * (features_dc holds 1, features_rest M-1).  features_* and rotation must be 16-byte aligned. */
typedef struct lgr_raw_params {
    const float* xyz;           /* [P,3]      */
    const float* features_dc;   /* [P,1,3]    */
    const float* features_rest; /* [P,M-1,3]  */
    const float* scaling;       /* [P,3] log-scale         -> exp        */
    const float* rotation;      /* [P,4] raw quaternion    -> normalize  */
    const float* opacity;       /* [P,1] logit             -> sigmoid    */
    int32_t features_rest_row_stride; /* floats between consecutive rows of features_rest; 0 = dense, i.e. (M-1)*3.
                                       * A larger value describes a row-strided view such as the distillation student's
                                       * _features_rest[:, :8, :] of a [P,15,3] tensor (scene/gaussian_model.py:129-136):
                                       * stride 45, 24 floats used.  The storage must hold P*stride floats from the pointer. */
} lgr_raw_params;

typedef struct lgr_raw_grads { /* dL/d(leaf), same shapes, fully written */
    float* xyz;
    float* features_dc;
    float* features_rest;
    float* scaling;
    float* rotation;
    float* opacity;
    float* rgb; /* optional [P,3]: clamp-masked dL/dRGB of this view.  When rgb != NULL and features_rest == NULL ("compact
                   mode") features_dc / features_rest are not written: see lgr_sh_grad_from_views. */
} lgr_raw_grads;

/* gaussians_count / important_score may both be NULL (plain forward) or both non-NULL (significance mode). */
int lgr_forward_raw(const lgr_view* view, int P, int M, const lgr_raw_params* params,
                    lgr_alloc_fn geometry_alloc, void* geometry_user,
                    lgr_alloc_fn binning_alloc, void* binning_user,
                    lgr_alloc_fn image_alloc, void* image_user,
                    float* out_color, int32_t* gaussians_count, float* important_score, int32_t* radii,
                    int32_t* num_rendered, void* cuda_stream);

int lgr_backward_raw(const lgr_view* view, int P, int M, int num_rendered, const lgr_raw_params* params,
                     const int32_t* radii, char* geometry_blob, char* binning_blob, char* image_blob,
                     const float* dL_dout_color, const lgr_raw_grads* grads, float* dL_dmeans2D, void* cuda_stream);

/* The same backward in two stages, so a caller can start exchanging this view's dL/dRGB (written to d_rgb[P,3] when non-NULL)
 * while the per-Gaussian stage runs:  begin = accumulator clear + blend backward (+ dRGB extraction),  end = K7+K8.
 * lgr_backward_raw == begin(d_rgb = NULL) followed by end.  In lgr_backward_raw_end, grads->features_rest == NULL selects the
 * compact mode (SH leaves not written). */
int lgr_backward_raw_begin(const lgr_view* view, int P, int num_rendered, const int32_t* radii, char* geometry_blob, char* binning_blob,
                           char* image_blob, const float* dL_dout_color, float* d_rgb, void* cuda_stream);
int lgr_backward_raw_end(const lgr_view* view, int P, int M, const lgr_raw_params* params, const int32_t* radii, char* geometry_blob,
                         const lgr_raw_grads* grads, float* dL_dmeans2D, void* cuda_stream);
/* stage 2 for Gaussians [first, first+count) only (first a multiple of 256): lets a view-parallel caller exchange one range's
 * gradients while the next range is computed */
int lgr_backward_raw_end_range(const lgr_view* v, int P, int M, const lgr_raw_params* params, const int32_t* radii, char* geometry_blob,
                               const lgr_raw_grads* grads, float* dL_dmeans2D, int first, int count, void* cuda_stream);

/* View-parallel training: for one view dL/dSH[k][c] = basis_k(dir) * dRGB[c] is rank-1 per Gaussian
 * (RAST/cuda_rasterizer/backward.cu:44-97), so ranks exchange dRGB (12 B/Gaussian/view, all-gather) instead of the
 * dense 12*M B/Gaussian gradient, and each rank rebuilds the SUM over views here:
 *   d_features_dc/rest[i] = sum_v basis(normalize(xyz[i] - campos[v])) (x) d_rgb[v][i]       (rows k >= (D+1)^2 are zero)
 * campos: [n_views,3] device; d_rgb: [n_views,P,3] device. */
int lgr_sh_grad_from_views(int P, int M, int sh_degree, int n_views, const float* xyz, const float* campos, const float* d_rgb,
                           float* d_features_dc, float* d_features_rest, void* cuda_stream);

/* All-reduce (sum) of n_floats floats over NVLink peer memory, for the view-parallel gradient exchange.
 * peer_buffers[r] (HOST array of `world` DEVICE pointers) is rank r's buffer as mapped into THIS process (symmetric /
 * IPC memory, 16-byte aligned); n_floats must be a multiple of 4.  This rank reduces slice `rank` from all peers with P2P
 * loads (fixed rank order: every rank obtains bit-identical sums) and stores it into all peers.  The caller must place a
 * cross-GPU barrier on the stream BEFORE (all ranks' data written) and AFTER (all peers' stores landed) this call. */
int lgr_peer_allreduce(float* const* peer_buffers, int rank, int world, size_t n_floats, void* cuda_stream);

/* Same contract as lgr_peer_allreduce, with the sum formed inside the NVSwitch: multicast_ptr is the NVLS multicast mapping
 * of the symmetric buffer (multimem.ld_reduce / multimem.st).  Barriers before and after are the caller's. */
int lgr_multimem_allreduce(float* multicast_ptr, int rank, int world, size_t n_floats, void* cuda_stream);

/* ---- fused image loss of the training loops (utils/loss_utils.py:18-85) ----
 * forward: out2[0] = mean|img - target| (l1_loss), out2[1] = ssim(img, target) (11x11 Gaussian window, sigma 1.5, zero padding,
 * C1 = 0.01^2, C2 = 0.03^2, mean over C*H*W).  dmaps (optional, [3,C,H,W]) receives what the backward needs.  All device
 * pointers; planar [C,H,W] float32.  workspace: lgr_image_loss_workspace_bytes() bytes, 8-byte aligned.
 * backward: d_img = s * d/d img ( g_l1 * l1 + g_ssim * ssim ),  s = *grad_scale (device scalar) or 1 when NULL.
 * The reference's training loss (1-l)*l1 + l*(1-ssim) (prune_finetune.py:160-164) is g_l1 = 1-l, g_ssim = -l. */
size_t lgr_image_loss_workspace_bytes(int C, int H, int W);
int lgr_image_loss_forward(const float* img, const float* target, int C, int H, int W, float* out2, float* dmaps, void* workspace,
                           void* cuda_stream);
/* L1 only: out2[0] = mean|img - target|, out2[1] = 0 (no SSIM filtering); same workspace; backward = lgr_image_loss_backward with g_ssim = 0 */
int lgr_image_l1_forward(const float* img, const float* target, int C, int H, int W, float* out2, void* workspace, void* cuda_stream);
int lgr_image_loss_backward(const float* img, const float* target, const float* dmaps, int C, int H, int W, float g_l1, float g_ssim,
                            const float* grad_scale, float* d_img, void* cuda_stream);

/* Push variant of lgr_backward_raw_sparse_pack for N ranks (N <= 8): every rank owns ONE exchange buffer of N slots of
 * lgr_sparse_exchange_bytes(P) bytes each, mapped into all peers; slot_of_this_rank[r] is the address of slot `self` inside the buffer
 * of rank r.  The packed view is written to all of them (plain stores over NVLink from inside the kernels), so after one cross-GPU
 * barrier lgr_backward_raw_sparse_accumulate runs on the N slots of the rank's OWN buffer: no remote load is on its critical path. */
int lgr_backward_raw_sparse_pack_push(const lgr_view* view, int P, int M, const lgr_raw_params* params, const int32_t* radii, char* geometry_blob,
                                      void* const* slot_of_this_rank, int world, int self, void* workspace, float* dL_dmeans2D,
                                      void* cuda_stream);

/* ---- sparse view-parallel gradient exchange over peer memory ----
 * Only ~13 % of the Gaussians get a non-zero gradient from one view.  lgr_backward_raw_sparse_pack (after lgr_backward_raw_begin) runs the
 * per-Gaussian backward on those only and publishes, in `exchange_buffer` (lgr_sparse_exchange_bytes(P) bytes, 256-byte aligned, mapped into
 * every peer): the view's camera position, a bitmap + prefix counts of the non-zero Gaussians and their 64-byte gradient rows.  After a
 * cross-GPU barrier, lgr_backward_raw_sparse_accumulate reads the rows of all `world` buffers (peer_buffers[r] = rank r's buffer as mapped
 * here; P2P loads) in rank order and writes the six dense leaf gradients, summed over the views, bit-identically on every rank.
 * workspace: lgr_sparse_workspace_bytes(P) bytes of local scratch.  dL_dmeans2D (local view only) is written dense. */
size_t lgr_sparse_exchange_bytes(int P);
size_t lgr_sparse_workspace_bytes(int P);
int lgr_backward_raw_sparse_pack(const lgr_view* v, int P, int M, const lgr_raw_params* params, const int32_t* radii, char* geometry_blob,
                                 void* exchange_buffer, void* workspace, float* dL_dmeans2D, void* cuda_stream);
int lgr_backward_raw_sparse_accumulate(int P, int M, int sh_degree, int world, const void* const* peer_buffers, const float* xyz,
                                       const lgr_raw_grads* grads, void* cuda_stream);

/* ---- optimizer side of the training loops ----
 * lgr_adamw_step: torch.optim.AdamW's default (foreach) update, amsgrad off, for up to 8 tensors in ONE launch; replaces
 * `gaussians.optimizer.step()` (prune_finetune.py:287, optimizer built at scene/gaussian_model.py:184-217).  `step` is the
 * tensor's step count AFTER the increment (t >= 1); lr is per tensor (one per param group), the rest is shared.  The host
 * scalars are formed in double precision exactly as torch/optim/adam.py does and rounded to fp32 at the kernel boundary. */
typedef struct lgr_adamw_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t numel;
    double lr;
    double step;
    int64_t row_elems;        /* 0: param is contiguous.  Otherwise param is a row-strided view: element e lives at     */
    int64_t param_row_stride; /* param[(e / row_elems) * param_row_stride + e % row_elems]; grad and moments stay dense */
} lgr_adamw_tensor;
int lgr_adamw_step(int n_tensors, const lgr_adamw_tensor* tensors, double beta1, double beta2, double eps, double weight_decay,
                   void* cuda_stream);

/* Row compaction of GaussianModel._prune_optimizer / prune_points (scene/gaussian_model.py:564-600): `keep` is a device byte
 * mask over P rows.  lgr_compact_plan writes the indices of the kept rows, ascending, to src_row[0..rows_out) and returns
 * rows_out through a host pointer (one stream synchronisation); lgr_compact_rows then gathers up to 24 row-major tensors
 * (row_words 4-byte words per row: parameters and both Adam moments of all groups) in ONE launch. */
typedef struct lgr_compact_tensor {
    const void* src;
    void* dst;
    int32_t row_words;
} lgr_compact_tensor;
size_t lgr_compact_workspace_bytes(int P);
int lgr_compact_plan(int P, const uint8_t* keep, int32_t* src_row, void* workspace, size_t workspace_bytes, int32_t* rows_out_host,
                     void* cuda_stream);
int lgr_compact_rows(int rows_out, const int32_t* src_row, int n_tensors, const lgr_compact_tensor* tensors, void* cuda_stream);

/* ---- VecTree vector quantisation of the SH features (vectree/vq.py:262-306, vectree/vectree.py:86-125,166-207) ----
 * All pointers are device pointers; x is [n,d] row-major float32, embed [K,d], d <= 64.
 * lgr_vq_assign: idx[i] = argmin_c |x_i - e_c|^2 (evaluated as |e_c|^2 - 2 x_i.e_c like torch.cdist's expansion; smallest index wins
 *   ties).  When cluster_batch / embed_sum are given they receive sum_i w_i [idx_i = c] and sum_i w_i x_i [idx_i = c], with
 *   w_i = weight_i * n / *weight_sum (the reference's normalisation), or 1 when weight is NULL: the quantities EuclideanCodebook.forward
 *   forms with F.one_hot and einsum.  workspace: lgr_vq_workspace_bytes(n) bytes, 8-byte aligned.
 * lgr_vq_ema_update: cluster_size <- decay*cluster_size + (1-decay)*cluster_batch; embed <- decay*embed + (1-decay)*embed_sum/smoothed,
 *   smoothed = (cluster_size+eps)/(sum+K*eps)*sum (vq.py:286-300).  scratch: one float.
 * lgr_vq_gather: out[i,:] = embed[idx[i],:] (batched_embedding, vq.py:161-165).
 * lgr_vq_pack_indices / lgr_vq_unpack_indices: `bits` bits per index, most significant first, bytes filled from the high bit
 *   (dec2bin + np.packbits, vectree.py:120-125; np.unpackbits + bin2dec, vectree/utils.py:33-39); out has (n*bits+7)/8 bytes. */
size_t lgr_vq_workspace_bytes(int64_t n);
/* lgr_vq_assign for d <= 32 and n*K >= 2^20 runs the scores on the tensor cores (csrc/lgr_vq_tc.cuh: bf16 hi/lo split operands, tcgen05.mma,
 * accumulators in TMEM) and re-scores in exact FP32 only the rows whose two best codes are closer than the split's error bound; indices are
 * the same as the FP32 kernel's.  lgr_set_vq_mode(1) forces the FP32 kernel (A/B measurements).  The tensor-core path keeps a grow-only
 * device scratch (operand tiles) inside the library. */
int lgr_set_vq_mode(int mode);
int lgr_vq_assign(int n, int d, int K, const float* x, const float* embed, const float* weight, const float* weight_sum, int32_t* idx,
                  float* cluster_batch, float* embed_sum, void* workspace, void* cuda_stream);
int lgr_vq_ema_update(int K, int d, double decay, double eps, float* cluster_size, float* embed, const float* cluster_batch,
                      const float* embed_sum, float* scratch, void* cuda_stream);
int lgr_vq_gather(int n, int d, const int32_t* idx, const float* embed, float* out, void* cuda_stream);
int lgr_vq_pack_indices(int64_t n, int bits, const int32_t* idx, uint8_t* out, void* cuda_stream);
int lgr_vq_unpack_indices(int64_t n, int bits, const uint8_t* in, int32_t* idx, void* cuda_stream);

/* present[i] = (view-space z of point i) > 0.2   (RAST/cuda_rasterizer/rasterizer_impl.cu:54-66, auxiliary.h:139-164) */
int lgr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                     void* cuda_stream);

/* ---- introspection (tests, diagnostics) ---- */
int lgr_abi_version(void);
const char* lgr_last_error(void); /* thread-local, valid until the next failing call on this thread */

/* Byte offsets of the per-Gaussian arrays inside a geometry blob for P points, so tests can read the
 * intermediates the way RAST/cuda_rasterizer/rasterizer_impl.cu:155-170 lays them out for the reference.
 * out[0]=depth(f32) out[1]=means2D(float2) out[2]=conic_opacity(float4) out[3]=rgb(float4, w unused)
 * out[4]=cov3D(6 f32) out[5]=clamped(u8 bitmask r=1,g=2,b=4) out[6]=tiles_touched(u32) out[7]=depth-sorted ids(u32)
 * Returns the total geometry blob size. */
size_t lgr_geometry_layout(int P, size_t* out, int n_out);
/* out[0]=final_T(f32[N]) out[1]=n_contrib(u32[N]) out[2]=ranges(uint2[tiles]).  Returns image blob size. */
size_t lgr_image_layout(int width, int height, size_t* out, int n_out);
/* out[0]=point_list(u32[R]) -- per-tile, depth-sorted Gaussian ids.  Returns binning blob size for R instances. */
size_t lgr_binning_layout(int num_rendered, int width, int height, size_t* out, int n_out);
/* Kernel launches issued by this library since load (for bench.py's gpu_launches). */
uint64_t lgr_launch_count(void);

/* Diagnostics / A-B measurements: which blend kernels forward and backward launch.  0 (default) = the shared-ring kernels
 * (csrc/lgr_blend.cuh: producer warp + TMA-staged per-instance records), 1 = the round-1 per-warp kernels.  Process-wide; a
 * backward must run in the mode its forward ran in (the ring backward streams the records the ring forward stored). */
int lgr_set_blend_mode(int mode);

/* Binning (per-tile depth-ordered instance lists; replaces the scan + 64-bit radix sort of RAST/cuda_rasterizer/rasterizer_impl.cu:278-319
 * and its blocking device-to-host copy of num_rendered at :282).  2 (default) = depth sort of the P Gaussians + stable tile bucketing with the
 * CUDA toolkit's radix sort / scan, one stream synchronisation for the instance count (one binning_alloc call, the reference's behaviour).
 * 0 = hand-written kernels (csrc/lgr_bin.cuh), no library and no GPU idle on the host: the binning allocator is called BEFORE the instance
 * count is known, with a size from a running estimate; the kernels bound their stores by that capacity, the host reads the count while the
 * blend kernel is already running and repeats scatter + blend with an exactly sized blob (a second binning_alloc call) when the estimate was
 * too small.  1 = the same kernels, blob sized exactly after a stream synchronisation.  All three produce bit-identical lists
 * (tests/test_gpu_binning.py); 2 is the default because it is the fastest measured; it is also used automatically above
 * 32 768 tiles.  Process-wide.  lgr_binning_overflows() = views that took the repeat. */
int lgr_set_binning_mode(int mode);
uint64_t lgr_binning_overflows(void);
void lgr_set_binning_estimate(uint64_t instances);   /* overwrite the running estimate of mode 0 (tests; 0 = forget) */

/* Diagnostics / A-B: the fused K7+K8 of lgr_backward_raw for a whole view with dense outputs.  0 (default) = one pass zero-fills all
 * output rows (TMA bulk stores) and lists the Gaussians whose accumulators are non-zero, a second kernel runs K7+K8 on that list;
 * 1 = the dense one-warp-per-32-Gaussians kernel. */
int lgr_set_kback_mode(int mode);

/* Exact tile-level culling at binning time (default on): (tile, Gaussian) instances in which no pixel can reach
 * alpha >= 1/255 are not listed.  Images, gradients and significance are unchanged; only the internal lists shrink.
 * Turn it off to obtain per-tile lists identical to the reference's (tests).  num_rendered always reports the
 * reference's value.  The geometry blob starts with int32 words: [0] instances listed, [1] the reference's num_rendered,
 * [2] instances the binning blob was sized for, [3] capacity overflow flag (0 once a forward call has returned). */
int lgr_set_tile_culling(int on);

/* Optional per-stage device timing: when enabled every launch is bracketed by CUDA events on its stream.
 * lgr_profile_collect() synchronises the device, writes the accumulated milliseconds and launch counts of each
 * stage (index < lgr_profile_stage_count()) and resets them.  Timing mode adds event overhead; do not use it
 * inside a throughput measurement. */
int lgr_profile_enable(int on);
int lgr_profile_stage_count(void);
const char* lgr_profile_stage_name(int stage);
int lgr_profile_collect(double* ms_out, uint64_t* launches_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* LGRAST_H_INCLUDED */
