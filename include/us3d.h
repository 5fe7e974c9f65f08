/* us3d.h — C ABI of the B200-native UnScene3D hot path (libus3d.so, sm_100a).
 *
 * This is the drop-in boundary: what the reference reaches through MinkowskiEngine's pybind layer
 * (MinkowskiEngine is an un-vendored dependency, /root/reference/.devcontainer/Dockerfile:50-51;
 * the reference-side call sites are listed per function) and through its own pybind extensions
 * (third_party/pointnet2/_ext_src/src/bindings.cpp:9-22, utils/cuda_utils/cuda_utils.cpp:49-54).
 *
 * Conventions (mirroring the reference's native boundaries):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _h;
 *   - the caller allocates every output and scratch buffer, kernels fill them in place;
 *   - dtypes are fixed: float32 features/weights, int32 coordinates and row indices, (b,x,y,z) rows;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless stated ("[sync]": the function returns a host value and waits for the stream);
 *   - return value 0 = ok, negative = error (us3d_last_error() gives the text, thread local).
 */
#ifndef US3D_H
#define US3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define US3D_ABI_VERSION 18
#define US3D_MAX_KVOL 27

int us3d_abi_version(void);
const char *us3d_last_error(void);
/* number of kernels launched by this library since load / since the last reset (bench.py gpu_launches) */
long long us3d_launch_count(void);
void us3d_reset_launch_count(void);

/* ---------------------------------------------------------------- coordinates (SURVEY §8(a) A1–A3)
 * Open-addressing hash table over packed 64-bit voxel keys (b:10|x:18|y:18|z:18, biased), stored in
 * global memory (L2-resident on B200): keys[cap] uint64, vals[cap] int32, cap a power of two.     */

/* power-of-two capacity for n keys (load factor <= 0.5) */
int us3d_hash_capacity(int n);

/* Quantise + de-duplicate coordinate rows; replaces ME's CoordinateMap insert / stride
 * (ME.SparseTensor(...) trainer/trainer.py:115-117; conv(..., stride=2) models/res16unet.py:51-116)
 * and ME.utils.sparse_quantize (datasets/utils.py:266-270, 403-408).
 *   coords[n,4]  int32 rows (b,x,y,z);  each spatial axis is floored to a multiple of tstride[a]
 *   out_coords[n,4] unique rows in order of FIRST OCCURRENCE; out_first[n] the input row of each
 *   unique row (ascending); inverse[n] = unique row of every input row; *out_count_h = #unique.
 *   keys/vals: hash table of capacity cap, left holding key -> unique row for later lookups.
 *   scratch: int32[2*n + 4096].                                                             [sync] */
int us3d_coords_unique(const int32_t *coords, int n, int tsx, int tsy, int tsz, uint64_t *keys, int32_t *vals,
                       int cap, int32_t *out_coords, int32_t *out_first, int32_t *inverse, int32_t *scratch,
                       int *out_count_h, void *stream);

/* HOST variant for the data-loader side: the reference calls ME.utils.sparse_quantize on CPU arrays inside forked DataLoader
 * workers (datasets/utils.py:266-270, 403-408; conf/data/indoor.yaml:24), where no CUDA context may be created.  Host
 * pointers only: unique rows of coords_h[n, d] (d <= 8) in order of first occurrence; first_h[u] = input row of unique row u,
 * inverse_h[i] = unique row of input row i.  Returns the number of unique rows (>= 0) or a negative error.            */
int us3d_coords_unique_h(const int32_t *coords_h, int n, int d, int32_t *first_h, int32_t *inverse_h);

/* Neighbour table ("kernel map", ME KernelGenerator HYPER_CUBE, models/modules/common.py:137-144):
 *   nbr[k*n_q + q] = row r of the hashed map with coords_r == query[q] + offsets_h[k]  (or -1).
 *   offsets_h[kvol,3] is a HOST array of int32 (already scaled by tensor stride / negated by caller).
 *   tile_mask (optional, may be NULL): uint32[ceil(n_q/tile_rows)], bit k set iff any row of that
 *   tile has a neighbour at offset k.                                                             */
int us3d_kernel_map(const int32_t *query, int n_q, const int32_t *offsets_h, int kvol, const uint64_t *keys,
                    const int32_t *vals, int cap, int32_t *nbr, uint32_t *tile_mask, int tile_rows, void *stream);

/* ---------------------------------------------------------------- sparse convolution (A4, A5)
 * Output-stationary gather convolution over a neighbour table:
 *     Y[orow(j)] (=|+=) sum_k  X[nbr[k*n_rows + j]] · Wk        j = 0..n_rows-1
 * with Wk = W[k] ([cin,cout], row-major) or, when transpose_w, W[kk]^T with W stored [kvol,cout,cin]
 * and kk = flip_k ? kvol-1-k : k   (that is the dX pass of the same map).  orow(j) = out_rows ?
 * out_rows[j] : j.  X rows have leading dimension ldx, Y rows ldy (so halves of a concatenation can
 * be read / written in place).  bias (optional, [cout]) is added once per output row.
 * Replaces MinkowskiConvolution / MinkowskiConvolutionTranspose forward and input-gradient
 * (models/modules/common.py:146-155, 179-188).                                                    */
int us3d_spconv_gather(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *w, int cin,
                       int cout, int transpose_w, int flip_k, const float *bias, const int32_t *out_rows, float *y,
                       int ldy, int accumulate, const uint32_t *tile_mask, void *stream);

/* Tensor-core (tcgen05 + TMEM) path for cin, cout multiples of 16, cout <= 256.
 *   passes = 1: bf16 x bf16 -> fp32;  passes = 3: three-term bf16 split (hi·hi + lo·hi + hi·lo), fp32-faithful.
 *   wpack: weights pre-packed into bf16 planes laid out as the swizzled shared-memory image of every (offset,
 *   64-channel chunk) slab [W_hi rows | W_lo rows]; us3d_spconv_packed_bytes gives its size.
 *   pack_weights(transpose=1, flip_k) produces the slabs of the input-gradient pass (W[kk]^T), in which case
 *   the GEMM runs with cin' = cout, cout' = cin.  (us3d_spconv_pack_pair / _pack_many below build the images of one /
 *   of all convolutions of a network in one launch.)                                               */
int us3d_spconv_tc_supported(int cin, int cout);
long long us3d_spconv_packed_bytes(int kvol, int kdim, int ndim, int passes);
int us3d_spconv_pack_weights(const float *w, int kvol, int cin, int cout, int transpose, int flip_k, int passes,
                             void *out, void *stream);

/* Weight gradient of the same map:  dW[k] += sum_j X[nbr[k*n_rows+j]]^T · dY[orow(j)]   ([kvol,cin,cout]).
 * dW must be zero-initialised (or hold the value to accumulate into) by the caller.  Exact-fp32 SIMT kernel. */
int us3d_spconv_wgrad(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *dy, int ldy,
                      const int32_t *out_rows, float *dw, int cin, int cout, void *stream);

/* fp32 rows -> bf16 planes hi (and lo = x - hi when lo != NULL), row-major [n, c], c % 8 == 0: the operand form of the
 * tensor-core kernels.  Activations get their planes from the pass that produces them (us3d_bn_apply_planes,
 * us3d_bn_backward_planes); this entry point is for tensors that come from elsewhere.                       */
int us3d_split_bf16(const float *x, int ldx, int n, int c, void *hi, void *lo, void *stream);

/* Production forward / input-gradient kernel (csrc/spconv_mt.cu): the activation rows are consumed as bf16 planes [n_in, cin];
 * neighbour rows are gathered with 16-byte cp.async into K-major SWIZZLE_128B slots (indices streamed ahead through a
 * shared-memory ring), up to T = 512 / acc_cols output tiles share every weight slab (one TMEM accumulator each), products
 * on tcgen05.  Work decomposition: every CTA owns a contiguous range of output tiles — `partition` (may be NULL; int32
 * [us3d_spconv_partition_size()], from us3d_spconv_partition on the table's tile masks) makes the ranges equal in cost
 * rather than in count — or, on small maps, the offsets of a tile are dealt to several CTAs whose partial tiles meet in a
 * caller-provided workspace.  out_rows (may be NULL): table column j writes row out_rows[j] (pattern-ordered tables).      */
int us3d_spconv_gather_mt(const void *x_hi, const void *x_lo, int n_in, const int32_t *nbr, int n_rows, int kvol,
                          const void *wpack, int cin, int cout, int passes, const float *bias, const int32_t *out_rows,
                          float *y, int ldy, int accumulate, const uint32_t *tile_mask, const int32_t *partition,
                          void *workspace, long long workspace_bytes, void *stream);
/* The same launch with the BatchNorm statistics of y folded in (MinkowskiConvolution -> MinkowskiBatchNorm,
 * /root/reference/models/modules/common.py:20-22, resnet_block.py:48-58): the epilogue (or, on split maps, the pass that sums the
 * partial tiles) adds the column sums and sums of squares of the rows it writes, the last CTA finalises mean / invstd and the
 * running statistics exactly like us3d_bn_stats_fused — one launch and one read of y fewer per normalised convolution.
 * bn == NULL: plain us3d_spconv_gather_mt.  bn->ws: the scratch of us3d_bn_workspace_bytes (zero on entry, zero on exit);
 * running_mean / running_var / num_batches_tracked may be NULL; accumulate must be 0. */
typedef struct us3d_bn_fuse {
    double *ws;
    float *mean, *invstd;                 /* out: [cout] */
    float *running_mean, *running_var;    /* in/out, may be NULL */
    long long *num_batches_tracked;       /* in/out, may be NULL */
    float eps, momentum;
} us3d_bn_fuse_t;
int us3d_spconv_gather_mt_bn(const void *x_hi, const void *x_lo, int n_in, const int32_t *nbr, int n_rows, int kvol,
                             const void *wpack, int cin, int cout, int passes, const float *bias, const int32_t *out_rows,
                             float *y, int ldy, int accumulate, const uint32_t *tile_mask, const int32_t *partition,
                             void *workspace, long long workspace_bytes, const us3d_bn_fuse_t *bn, void *stream);
/* Launch lists (csrc/executor.cu): the launches of a residual block's forward or backward pass
 * (/root/reference/models/modules/resnet_block.py:24-64: conv - norm - relu - conv - norm - (+ shortcut) - relu) issued by ONE call.
 * Every op names an entry point of this header and carries its resolved arguments — p[]: pointers, v[]: integers, f[]: floats, in
 * the order documented next to each kind in csrc/executor.cu (the entry point's own argument order).  us3d_run_ops calls them in
 * list order on `stream` and returns the first non-zero return code; results are those of the call-by-call route. */
enum { US3D_OP_CONV = 1,        /* us3d_spconv_gather_mt_bn   */
       US3D_OP_BN_APPLY = 2,    /* us3d_bn_apply_planes       */
       US3D_OP_BN_BACKWARD = 3, /* us3d_bn_backward_planes    */
       US3D_OP_WGRAD = 4,       /* us3d_spconv_wgrad_planes   */
       US3D_OP_ADD = 5 };       /* us3d_add                   */
typedef struct us3d_op {
    int kind;
    const void *p[16];
    long long v[10];
    float f[2];
} us3d_op_t;
int us3d_run_ops(const us3d_op_t *ops, int n_ops, void *stream);
/* the same list as flat arrays (one conversion for a ctypes caller): a = n_ops x [kind, p[16], v[10]] (int64), f = n_ops x [f0, f1] (double) */
int us3d_run_ops_flat(const long long *a, const double *f, int n_ops, void *stream);

/* workspace (may be NULL; 16-byte aligned, contents irrelevant): room for the partial tiles of the split mode,
 * [parts][n_rows][cout] fp32, summed in part order by a second launch (no atomics: results are bit-reproducible).
 * us3d_spconv_gather_mt_workspace_bytes = the most the launcher can use for a map (0: it would not split).             */
long long us3d_spconv_gather_mt_workspace_bytes(int n_rows, int kvol, int cout);
int us3d_spconv_partition_size(void);
int us3d_spconv_partition(const uint32_t *tile_mask, int n_tiles, int kvol, int32_t *partition, void *stream);

int us3d_spconv_wgrad_tc_supported(int cin, int cout);

/* Production weight-gradient kernel: operands from the bf16 planes of X and dY via cp.async; one CTA handles up
 * to 4 kernel offsets per dY tile (one TMEM accumulator each), so dY is streamed ceil(kvol/4) times, not kvol. */
int us3d_spconv_wgrad_planes(const void *x_hi, const void *x_lo, const void *dy_hi, const void *dy_lo, const int32_t *nbr,
                             int n_rows, int kvol, float *dw, int cin, int cout, int passes, const uint32_t *tile_mask,
                             const int32_t *dy_rows, void *stream);
/* dy_rows (may be NULL): table column j pairs with dY row dy_rows[j] — the table is in pattern order (below). */

/* bf16 planes [n, c] (c % 8 == 0) re-ordered by rows: out[j] = in[order[j]] (lo / lo_out may both be NULL).  Brings dY into the
 * row order of a pattern-ordered table so that us3d_spconv_wgrad_planes streams contiguous dY tiles (dy_rows = NULL). */
int us3d_permute_planes(const void *hi, const void *lo, const int32_t *order, int n, int c, void *hi_out, void *lo_out, void *stream);

/* Rows of a large map ordered by neighbour pattern.  The tcgen05 kernels skip a kernel offset for a 128-row tile only if
 * no row of the tile has a neighbour there; spatially consecutive rows of a voxelised surface leave every offset active,
 * rows grouped by presence pattern do not (200k-voxel scene: 100 % -> 67 % of the (tile, offset) pairs, pair density 49 %).
 *   neighbour_pattern_keys: keys[j] = presence pattern of row j (bit = offset present) with the bits permuted so that the
 *       rarest offset is the most significant; the caller sorts rows by key (stable) to get `order`.
 *       scratch: uint32[n_rows + 32].
 *   kernel_map_reorder: nbr_out[k, j] = nbr[k, order[j]] and the tile masks of the re-ordered table; the convolution then
 *       runs with out_rows = order (forward / input gradient) or dy_rows = order (weight gradient).            */
int us3d_neighbour_pattern_keys(const int32_t *nbr, int n_rows, int kvol, uint32_t *scratch, int32_t *keys, void *stream);
int us3d_kernel_map_reorder(const int32_t *nbr, int n_rows, int kvol, const int32_t *order, int32_t *nbr_out,
                            uint32_t *tile_mask, int tile_rows, void *stream);

/* ---------------------------------------------------------------- normalisation / elementwise (A6)
 * BatchNorm1d over all rows (ME.MinkowskiBatchNorm, models/modules/common.py:20-22), train mode:
 *   stats: double sum[c], sumsq[c] (caller zeroes) ; finalize -> mean[c], invstd[c] float and the
 *   running statistics update (momentum, unbiased variance) exactly like torch.nn.BatchNorm1d.   */
int us3d_bn_stats(const float *x, int ldx, int n, int c, double *sum, double *sumsq, void *stream);
int us3d_bn_finalize(const double *sum, const double *sumsq, int n, int c, float eps, float momentum, float *mean,
                     float *invstd, float *running_mean, float *running_var, void *stream);
/* y = [relu]( (x-mean)*invstd*gamma + beta [+ residual] ) */
int us3d_bn_apply(const float *x, int ldx, int n, int c, const float *mean, const float *invstd, const float *gamma,
                  const float *beta, const float *residual, int ldr, int relu, float *y, int ldy, void *stream);
/* backward of the fused op above.  g = dy * (relu ? y>0 : 1):  red[0..c) = sum g, red[c..2c) = sum g*xhat
 * (double, caller zeroes); then dx = gamma*invstd*(g - red0/n - xhat*red1/n), dres = g.            */
int us3d_bn_bwd_reduce(const float *dy, int lddy, const float *x, int ldx, const float *y, int ldy, int n, int c,
                       const float *mean, const float *invstd, int relu, double *red, void *stream);
int us3d_bn_bwd_apply(const float *dy, int lddy, const float *x, int ldx, const float *y, int ldy, int n, int c,
                      const float *mean, const float *invstd, const float *gamma, int relu, const double *red,
                      float *dx, int lddx, float *dres, int lddres, float *dgamma, float *dbeta, void *stream);
/* one-call forms of the two pairs above (ws: 2*c doubles of scratch, zeroed inside); batch_terms = 0 treats the
 * statistics as constants (inference): dx = g * invstd * gamma                                          */
int us3d_bn_batch_stats(const float *x, int ldx, int n, int c, float eps, float momentum, float *mean, float *invstd,
                        float *running_mean, float *running_var, double *ws, void *stream);
int us3d_bn_backward(const float *dy, int lddy, const float *x, int ldx, const float *y, int ldy, int n, int c,
                     const float *mean, const float *invstd, const float *gamma, int relu, int batch_terms, double *ws,
                     float *dx, int lddx, float *dres, int lddres, float *dgamma, float *dbeta, void *stream);
/* inference-mode BN is us3d_bn_apply with mean=running_mean, invstd=rsqrt(running_var+eps) (host computes) */

/* Fused forms used by the module surface (csrc/fused_ops.cu).  `ws` = us3d_bn_workspace_bytes(c) bytes that are ZERO on
 * entry and are left zero on exit (the block that finishes last cleans up), so no memset separates two layers.
 *   bn_stats_fused      one launch: column sums, then mean / invstd / running statistics (nn.BatchNorm1d momentum
 *                       rule, unbiased running variance) and num_batches_tracked += 1 (int64 scalar, may be NULL)
 *   bn_apply_planes     us3d_bn_apply that also writes the result as bf16 planes hi (and lo = y - hi unless NULL),
 *                       row-major [n, c] — what the tcgen05 convolution of the next layer gathers from (c % 8 == 0)
 *   bn_backward_planes  us3d_bn_backward that also writes dx as bf16 planes (read by the input-gradient and
 *                       weight-gradient kernels of the preceding convolution)
 * Replace MinkowskiBatchNorm + MinkowskiReLU + `out += residual` (models/modules/common.py:20-22,
 * models/modules/resnet_block.py:48-64). */
int us3d_bn_workspace_bytes(int c);
int us3d_bn_stats_fused(const float *x, int ldx, int n, int c, float eps, float momentum, float *mean, float *invstd,
                        float *running_mean, float *running_var, long long *num_batches_tracked, double *ws, void *stream);
int us3d_bn_apply_planes(const float *x, int ldx, int n, int c, const float *mean, const float *invstd, const float *gamma,
                         const float *beta, const float *residual, int ldr, int relu, float *y, int ldy, void *hi, void *lo,
                         void *stream);
int us3d_bn_backward_planes(const float *dy, int lddy, const float *x, int ldx, const float *y, int ldy, int n, int c,
                            const float *mean, const float *invstd, const float *gamma, int relu, int batch_terms, double *ws,
                            float *dx, int lddx, float *dres, int lddres, float *dgamma, float *dbeta, void *dx_hi, void *dx_lo,
                            void *stream);

/* Forward (out_fwd) and input-gradient (out_dgrad: W[k]^T, offsets reversed when flip_dgrad) weight images of one
 * convolution in one call; either pointer may be NULL.  Sizes: us3d_spconv_packed_bytes(kvol, cin, cout, passes) and
 * us3d_spconv_packed_bytes(kvol, cout, cin, passes). */
int us3d_spconv_pack_pair(const float *w, int kvol, int cin, int cout, int flip_dgrad, int passes, void *out_fwd,
                          void *out_dgrad, void *stream);
/* Every weight image of a network in one launch (what a training step does after the optimizer update).  items: DEVICE array
 * of n_items descriptors { const float *w; uint8_t *out; int32 kvol, w_cin, w_cout, w_ci0, kdim, ndim, transpose, flip; }
 * (48 bytes): image `out` of the [kvol, w_cin, w_cout] weight, K = kdim, N = ndim; transpose = 0: forward image of input
 * channels w_ci0 .. w_ci0 + kdim; transpose = 1: input-gradient image (W^T, offsets reversed when flip) of input channels
 * w_ci0 .. w_ci0 + ndim.  Layout of each image as us3d_spconv_pack_weights.                                      */
int us3d_spconv_pack_many(const void *items, int n_items, int passes, void *stream);

/* The stem convolution (conv0p1s1: 3 input channels, models/res16unet.py:219-221): its K = 3 contraction does not fill
 * a tensor-core tile, so it runs as warp = rows, lane = output channel with the weights in shared memory.
 * cin <= 4 forward; weight gradient for cin == 3, kvol in {1, 8, 27} (dw zero-initialised by the caller). */
int us3d_stem_conv_supported(int cin, int cout, int kvol);
int us3d_stem_conv_fwd(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *w, int cin, int cout,
                       const float *bias, float *y, int ldy, void *stream);
int us3d_stem_conv_wgrad(const float *x, int ldx, const int32_t *nbr, int n_rows, int kvol, const float *dy, int lddy, float *dw,
                         int cin, int cout, void *stream);

/* y = relu(x) ; dx = dy * (y > 0) ; z = a + b ; concat along channels / its inverse */
int us3d_relu(const float *x, float *y, long long numel, void *stream);
int us3d_relu_bwd(const float *dy, const float *y, float *dx, long long numel, void *stream);
int us3d_add(const float *a, const float *b, float *z, long long numel, void *stream);
int us3d_copy2d(const float *src, int lds, float *dst, int ldd, int n, int c, void *stream);

/* ---------------------------------------------------------------- pooling (A7)
 * MinkowskiAvg/Sum/MaxPooling(k2,s2) (models/mask3d.py:131,213,432): y[o] = reduce over PRESENT
 * children nbr[k*n_out+o]; mode 0 = avg, 1 = sum, 2 = max.  bwd scatters through the same table.   */
int us3d_pool_fwd(const float *x, int c, const int32_t *nbr, int n_out, int kvol, int mode, float *y, void *stream);
int us3d_pool_bwd(const float *dy, const float *x, const float *y, int c, const int32_t *nbr, int n_out, int kvol,
                  int mode, float *dx, void *stream);

/* ---------------------------------------------------------------- decoder helpers
 * Furthest point sampling, bit-compatible with third_party/pointnet2/_ext_src/src/sampling_gpu.cu:72-176
 * (start at row 0, rows with |p|^2 <= 1e-3 skipped, ties to the lowest thread slot then lowest row;
 * thread count = min(2^floor(log2 n), 512), _ext_src/include/cuda_utils.h:15-21).
 *   xyz[b,n,3] float, temp[b,n] float scratch (caller fills with 1e10), idx[b,m] int32.          */
int us3d_furthest_point_sampling(const float *xyz, int b, int n, int m, float *temp, int32_t *idx, void *stream);

/* torch_scatter.scatter_mean(src[n,c], index[n], dim=0) (models/mask3d.py:223): out[s,c] zeroed by
 * caller, count[s] float zeroed by caller; fwd accumulates then divides; bwd gathers dy/count.     */
int us3d_segment_mean_fwd(const float *src, const int64_t *index, int n, int c, int s, float *out, float *count,
                          void *stream);
/* Same mean with the sum taken in fp64 (`acc` [s, c] doubles and `count` zero-filled by the caller) and rounded to fp32 once:
 * the pseudo-mask path thresholds affinities computed from these means (pseudo_masks/unscene3d_pseudo_main.py:350-402, 89-119),
 * so they must not depend on the order of the atomics. */
int us3d_segment_mean_f64(const float *src, const int64_t *index, int n, int c, int s, double *acc, float *out, float *count,
                          void *stream);
int us3d_segment_mean_bwd(const float *dout, const int64_t *index, const float *count, int n, int c, float *dsrc,
                          void *stream);

/* Masked multi-head cross-attention core of the decoder (nn.MultiheadAttention inside CrossAttentionLayer,
 * models/mask3d.py:561-651, called at models/mask3d.py:355-365): out = softmax(scale * q k^T, hidden keys removed) v per
 * (scene, head).  q [nq, b, h*head_dim], k / v [nk, b, h*head_dim] and out / dout / dq / dk / dv alike: fp32, contiguous,
 * sequence-first as the reference passes them (already projected).  mask: uint8 / bool, non-zero = key hidden from query
 * (the reference's memory_mask), element (scene, head, query, key) at mask[scene*sb + head*sh + query*sq + key*sk] — the
 * decoder's own [b, nk, nq] tensor is passed with sh = 0, the reference layout [b*h, nq, nk] with sb = h*nq*nk; NULL = no
 * mask.  lse [b, h, nq] receives the row log-sum-exp (saved for backward).  ws: us3d_xattn_workspace_bytes() bytes.
 * Nothing of size nq x nk is written; no atomics (run-to-run deterministic).  head_dim 16 or 32.
 * A query whose keys are all hidden yields NaN, as torch's softmax does (the decoder un-hides such rows, mask3d.py:349). */
long long us3d_xattn_workspace_bytes(int b, int h, int nq, int nk, int head_dim);
int us3d_xattn_fwd(const float *q, const float *k, const float *v, const uint8_t *mask, long long mask_sb, long long mask_sh,
                   long long mask_sq, long long mask_sk, int b, int h, int nq, int nk, int head_dim, float scale, float *ws,
                   float *out, float *lse, void *stream);
int us3d_xattn_bwd(const float *q, const float *k, const float *v, const uint8_t *mask, long long mask_sb, long long mask_sh,
                   long long mask_sq, long long mask_sk, const float *out, const float *lse, const float *dout, int b, int h,
                   int nq, int nk, int head_dim, float scale, float *ws, float *dq, float *dk, float *dv, void *stream);

/* Hungarian matcher cost (models/matcher.py:97-168):  C[q,t] = w_mask * BCE + w_class * (-p[q, label_t])
 * + w_dice * dice,  logits[s,q] (pred_masks[b], row = segment/point), tgt[t,s] float {0,1},
 * prob[q,ncls] softmaxed class probabilities, labels[t] int64 (253 = ignore -> class cost -1).     */
int us3d_matcher_cost(const float *logits, int s, int q, const float *tgt, int t, const float *prob, int ncls,
                      const int64_t *labels, float w_class, float w_mask, float w_dice, float *cost, void *stream);

/* Attention masks of the decoder rounds (Mask3D.mask_module, models/mask3d.py:407-446: segment logits gathered to the voxels,
 * MinkowskiAvgPooling x 1..4, `sigmoid < 0.5`) as one sparse product: rowptr[n_rows+1] / col / val = CSR matrix A of the nested
 * average pooling composed with the voxel -> segment map (int64 indices, fp32 weights), seg[s_total, q] the segment logits;
 * bits[n_rows, q] (uint8) = sigmoid(A seg) < 0.5.                                                                  */
int us3d_pooled_mask_bits(const int64_t *rowptr, const int64_t *col, const float *val, int n_rows, const float *seg, int q,
                          uint8_t *bits, void *stream);

/* Fourier positional encoding of xyz rows (models/position_embedding.py:128-172 get_fourier_embeddings, called from
 * models/mask3d.py:183-198, 238-240): xyz[n, >=3] fp32 with leading dimension ld; lo[3] / hi[3] (device, both or neither) = the
 * input range the rows are normalised to the unit cube with; gauss_b[3, >= d_out] (leading dimension ldb) = the module's
 * `gauss_B` buffer; out[n, 2 d_out] row-major = [sin(2 pi t . B) | cos(2 pi t . B)] — the transpose of the reference's
 * [1, 2 d_out, n] result, which every caller permutes into this layout.                                              */
int us3d_fourier_posenc(const float *xyz, int n, int ld, const float *lo, const float *hi, const float *gauss_b, int ldb,
                        int d_out, float *out, void *stream);

/* Mask losses of the set criterion for the matched pairs of one scene (models/criterion.py:22-73 dice_loss / sigmoid_ce_loss,
 * called from loss_masks :168-216):  logits[s, q] fp32 (pred_masks of the scene), tgt[t_all, s] (float32, or uint8 / bool when
 * tgt_is_float == 0), qidx[t] / tidx[t] int64 = matched query / target of pair t, weights[t] (may be NULL = 1; the DropLoss
 * gate), n = the scene's normaliser.  fwd: stats[t, 4] (per-pair sums kept for backward), out[2] = (loss_mask, loss_dice).
 * bwd: gout[2] = upstream gradients of the two losses (device), dlogits[s, q] (zero outside the matched columns).      */
int us3d_mask_loss_fwd(const float *logits, int s, int q, const void *tgt, int tgt_is_float, const int64_t *qidx,
                       const int64_t *tidx, int t, const float *weights, float n, float *stats, float *out, void *stream);
int us3d_mask_loss_bwd(const float *logits, int s, int q, const void *tgt, int tgt_is_float, const int64_t *qidx,
                       const int64_t *tidx, int t, const float *weights, float n, const float *stats, const float *gout,
                       float *dlogits, void *stream);

/* ---------------------------------------------------------------- pseudo-mask NCut step (A19, A20)
 * get_affinity_matrix / second_smallest_eigenvector of pseudo_masks/unscene3d_pseudo_main.py:89-146.
 *   gram:      A[s,s] = normalised-row Gram matrix of f[s,d] (fp32) + stats {min over non-zero, max, any > 0} (ordered-uint
 *              encoded, uint32[3]) that normalize_mat (:82-86) needs; inv_norm: float[s] scratch.
 *   threshold: averages the two normalised matrices (Ab may be NULL), thresholds at tau into a bit matrix
 *              bits[s, ceil(s/32)] (W = eps 11^T + (1-eps) B), clears painted rows/columns (uint8[s], may be NULL) and
 *              writes the degrees of the UNPAINTED graph (double[s]) — the reference computes D before painting.
 *   matvec:    y = W x in fp64 (xsum = sum(x), 1 double on the device).                                      */
int us3d_ncut_gram(const float *f, int s, int d, float *inv_norm, float *A, uint32_t *stats, void *stream);
int us3d_ncut_threshold(const float *Aa, const float *Ab, int s, const uint32_t *stats_a, const uint32_t *stats_b, float tau,
                        double eps, const uint8_t *painted, uint32_t *bits, double *degree, void *stream);
int us3d_ncut_matvec(const uint32_t *bits, int s, double eps, const double *x, const double *xsum, double *y, void *stream);
/*   lanczos:   the spectral step of second_smallest_eigenvector (:138-146) as ONE cooperative launch: Lanczos steps [j0, j1) of
 *              M = D^-1/2 W D^-1/2 with two Gram-Schmidt passes per step against every earlier basis vector, entirely on the
 *              device (bit matrix, basis slices in shared memory, three grid barriers per step).  Q: double[(m + 2), s], row 0 =
 *              the deflated vector D^1/2 1 / |.|, row 1 = the unit start vector, rows 2.. are written (unit vectors on return);
 *              dinv = D^-1/2; alpha / beta: double[m]; *steps_done (device int) = steps completed so far — < j1 when the
 *              recurrence broke down (beta < breakdown).  The workspace is zeroed by the call with j0 == 0 and carries the
 *              state of later calls (j0 = previous *steps_done).  s <= 32 x (number of SMs).                              */
long long us3d_ncut_lanczos_workspace_bytes(int m);
int us3d_ncut_lanczos(const uint32_t *bits, int s, double eps, const double *dinv, double *Q, double *alpha, double *beta, int j0,
                      int j1, int m, double breakdown, void *workspace, long long workspace_bytes, int *steps_done, void *stream);

/* ---------------------------------------------------------------- 2D -> 3D feature lifting
 * project_features_cuda.project_features_cuda (utils/cuda_utils/project_image_cuda_kernel.cu:24-146, 183-256; caller
 * utils/cuda_utils/raycast_image.py:18-77): every pixel's camera ray is marched through the dense occupancy grid in steps of
 * ray_inc from depth_min to depth_max (voxel units); the pixel's feature row is added to the first occupied voxel (int32 labels:
 * maximum, when pred_mode) and the voxel's hit count incremented.  feats [B, V, H, W, C]; occ int64 [B, Z, Y, X], 0 = empty else
 * voxel index; view_inv float [B, V, 4, 4]; intr float [B, 4] = (fx, fy, mx, my); hit int32 scratch [B * V * H * W]; counts
 * int32 [n_vox] and out [n_vox, C] are accumulated into.                                                                */
int us3d_project_features_2d3d(const void *feats, const long long *occ, const float *view_inv, const float *intr, int B, int V, int H, int W,
                               int C, int Z, int Y, int X, float depth_min, float depth_max, float ray_inc, int pred_mode, int *hit,
                               int *counts, void *out, void *stream);

/* ---------------------------------------------------------------- Felzenszwalb mesh over-segmentation
 * HOST function.  felzenszwalb_cpp.segment_mesh (utils/cpp_utils/segmentator.cpp:17-262; callers datasets/freemask_semseg.py:212,
 * pseudo_masks/datasets/scannet.py:182): vertices / colors float[n_verts][3], faces int32[n_faces][3] -> comps int32[n_verts]
 * (segment ids 0..S-1 in the order of their representative vertices) and the sorted distinct directed pairs of adjacent segments
 * int32[<= cap][2].  Returns the number of pairs (the first `cap` are written) or a negative error code.                 */
int us3d_felzenszwalb_segment_h(const float *vertices, const int32_t *faces, const float *colors, int n_verts, int n_faces, float kthr,
                                int seg_min_verts, int32_t *comps, int32_t *pairs, int cap);

/* ---------------------------------------------------------------- tri-plane projection (noise-robust loss)
 * custom_cuda_utils.project_sparse_voxels_to_planes[_backward] (utils/cuda_utils/cuda_utils.cpp:26-53, kernels
 * cuda_utils_kernel.cu:371-433, 496-556; callers models/noise_robust_loss.py:28-31, 67-69).  coords int32 [n, 4] (batch, x, y, z)
 * centred at zero; pred / tgt float [n, inst]; planes float [x_dim, y_dim, inst], [x_dim, z_dim, inst], [y_dim, z_dim, inst]
 * and int32 counts [x_dim, y_dim] ... are ACCUMULATED into (the caller zero-fills, as the reference's Python does).  Voxels
 * with a coordinate >= its dim are skipped (the reference sizes the planes by the maximum coordinate).  Backward: grad[n, inst]
 * = mean of the non-zero plane gradients at the voxel's three cells; skipped voxels are left untouched.                  */
int us3d_project_voxels_to_planes(const int32_t *coords, const float *pred, const float *tgt, int n, int inst, int x_dim, int y_dim,
                                  int z_dim, float *pred_xy, float *pred_xz, float *pred_yz, float *tgt_xy, float *tgt_xz,
                                  float *tgt_yz, int32_t *num_xy, int32_t *num_xz, int32_t *num_yz, void *stream);
int us3d_project_voxels_to_planes_bwd(const int32_t *coords, int n, int inst, int x_dim, int y_dim, int z_dim, const float *grad_xy,
                                      const float *grad_xz, const float *grad_yz, float *grad, void *stream);

/* ---------------------------------------------------------------- FreeMask-style pseudo masks (A22)
 * Segment branch of the scene loop of pseudo_masks/freemask_main.py:203-417.
 *   soft_masks:     soft[s,s] = cosine_sim(f, f) (utils/freemask_utils.py:8-18: rows L2-normalised with eps 1e-9, Gram matrix,
 *                   per-row min subtracted, divided by row max + eps), columns of all-zero rows of f set to 0 (:266);
 *                   norm: float[s] scratch (receives the row norms).
 *   row_stats:      per row of soft[m,s] (leading dimension ld): count[m] = #(soft >= thr) (:267-268, 343-344), soft_sum[m] =
 *                   sum of those values (numerator of maskness, :353); optional points[m] = sum of weights[s] (int32 points per
 *                   segment = masks.sum(1) on the mapped masks, :395) and bbox[m,6] (min xyz, max xyz over seg_min/seg_max
 *                   [s,3] doubles = extent of the mapped mask, :382-383).  weights / seg_min / seg_max / points / bbox may be NULL.
 *   weighted_inter: inter[m,m] = sum_s weights[s] [soft[i,s] >= thr][soft[j,s] >= thr] = (mask_i * mask_j).sum() of
 *                   matrix_nms(kernel='mask') on the mapped masks (utils/pc_utils.py:746).
 *   separate_h:     HOST function (all pointers host memory): the blob separation of :289-326, see csrc/freemask.cu.   */
int us3d_freemask_soft_masks(const float *f, int s, int d, float *norm, float *soft, void *stream);
int us3d_freemask_row_stats(const float *soft, int ld, int m, int s, float thr, const int32_t *weights, const double *seg_min,
                            const double *seg_max, int32_t *count, float *soft_sum, long long *points, double *bbox, void *stream);
int us3d_freemask_weighted_inter(const float *soft, int ld, int m, int s, float thr, const int32_t *weights, int32_t *inter,
                                 void *stream);
int us3d_freemask_separate_h(const uint8_t *masks_h, int m, int s, const int32_t *adj_ptr_h, const int32_t *adj_h,
                             int32_t *blob_query_h, int32_t *blob_ptr_h, int32_t *blob_members_h, int max_blobs,
                             long long max_members);

#ifdef __cplusplus
}
#endif
#endif /* US3D_H */
