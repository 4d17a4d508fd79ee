/* nerfq -- C ABI of the B200-native NeRF ray-rendering hot path (libnerfq.so).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the comment says "host";
 *   - every call is enqueued on `stream` and returns immediately (no allocation, no sync);
 *   - return value: 0 ok, -1 bad argument, -2 kernel attribute error, -3 launch error;
 *   - empty inputs (n == 0) are accepted and do nothing;
 *   - tensors are dense row-major float32 / int32.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * checkout jihyounchoi/vanilla-nerf-model-compression-using-lsa-enhanced-nncodec).  The reference has
 * no FFI of its own on the render path (it is torch eager code); its one native boundary is the
 * pybind11 module `deepCABAC`, whose quantLayer/dequantLayer the quantiser entry points mirror.
 */
#ifndef NERFQ_H
#define NERFQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* nerfq_stream_t; /* == cudaStream_t */

/* ------------------------------------------------------------------------------------------------
 * Quantiser (replaces deepCABAC Encoder.quantLayer(dq_flag=0) / Decoder.dequantLayer)
 *   call sites: nnc_core/approximator/baseline.py:48-57 (quantLayer), :98 (dequantLayer)
 * ---------------------------------------------------------------------------------------------- */

/* delta(qp, qp_density): nnc_core/common.py:28-46.  `out` is a HOST pointer. */
int nerfq_stepsize(int qp, int qp_density, float* out);

/* level = sign(w) * (int)(|w|/delta + 0.5f).  The qp is raised until the largest level fits int32
 * (baseline.py:60-62) and written to *qp_used (device int, nullable).  workspace4: 4 bytes.
 * Non-finite input never hangs or traps: a NaN / Inf element gets level 0, and when the tensor's maximum is
 * non-finite the qp is left as requested (the clip search is bounded). */
int nerfq_quantize_urq(const float* w, int32_t* lvl, long long n, int qp, int qp_density, int* qp_used,
                       void* workspace4, nerfq_stream_t stream);

/* The same for `count` (<= 64) tensors in one launch pair -- every tensor of a model, as run_ft_and_lsa quantises
 * and reconstructs them before tuning (nnc_core/approximator/__init__.py:655-661).  w, lvl, rec, n, qp are HOST
 * arrays of length count (device pointers / element counts / per-tensor qp); rec may be NULL, rec[t] may be NULL or
 * alias w[t] (reconstruction level*delta in place).  workspace: 4*count bytes; qp_used: nullable device int[count]. */
int nerfq_quantize_batch(const float* const* w, int32_t* const* lvl, float* const* rec, const long long* n, const int* qp,
                         int count, int qp_density, int* qp_used, void* workspace, nerfq_stream_t stream);

/* w = (float)level * delta(qp). */
int nerfq_dequantize(const int32_t* lvl, float* w, long long n, int qp, int qp_density, nerfq_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Packed network (replaces the NeRF / ScaledLinear parameter tensors as the MLP kernels' input)
 *   utils.py:18-80 (NeRF), framework/applications/utils/transforms.py:84-111 (ScaledLinear)
 * Layer order of the 12 weight pointers: pts_linears.0..7, alpha_linear, feature_linear,
 * views_linears.0, rgb_linear.  Flat per-channel order (scale, bias, d_scale; 2436 entries):
 * pts0..7, feature, views, alpha, rgb.
 * ---------------------------------------------------------------------------------------------- */
unsigned long long nerfq_packed_net_bytes(void);
int nerfq_num_channels(void);

/* weights12: HOST array of 12 device pointers ([out,in] row-major; int32 levels if src_is_int32 else float32);
 * delta12: HOST array of the 12 step sizes (1.0 for unquantised weights).
 * Operand range: the MLP kernels multiply fp16 operands.  Integer levels up to +-2048 are exact; larger ones are rounded
 * to 11 significant bits (relative error <= 2^-12) and, beyond fp16's range, a power of two is folded into the layer's
 * delta so nothing overflows.  Float weights are normalised per layer by a power of two the same way. */
int nerfq_pack_net(void* packed, const void* const* weights12, const float* delta12, int src_is_int32,
                   nerfq_stream_t stream);

/* max |level| (or |weight|) of each of the 12 layers as found by the last nerfq_pack_net, copied to max_abs12
 * (DEVICE float[12]): > 2048 for an int32 source means that layer's operands were rounded; NaN / Inf means the
 * source held non-finite values. */
int nerfq_pack_status(const void* packed, float* max_abs12, nerfq_stream_t stream);

/* Epilogue constants {delta*scale, bias} per output channel; scale == NULL means no LSA (scale 1).  Must follow every
 * nerfq_pack_net and every change of the scales: it also rebuilds the backward weight image (level * delta * scale,
 * transforms.py:104-111 folded into the dgrad operand) that nerfq_mlp_backward streams. */
int nerfq_set_scale_bias(void* packed, const float* scale, const float* bias, nerfq_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused positional encoding + MLP (replaces run_network + NeRF.forward)
 *   framework/nerf_model/run_nerf.py:46-63,408,430; run_nerf_helpers.py:18-67; utils.py:57-80
 * rays [n_rays,11] = o(3) d(3) near far viewdir(3); z [n_rays,S]; raw [n_rays,S,4] = rgb logits, sigma.
 * The sample points o + d*z, their positional encodings and all layer activations stay on chip; operands are fp16
 * (weights: exact integer levels), accumulation is fp32, the epilogue applies delta*scale and bias in fp32.
 * save (nullable): nerfq_mlp_save_bytes(n_rays*S) bytes receiving the activations the backward needs.
 * max_ctas: 0 = one CTA per SM.  The result does not depend on max_ctas, on how the rays are split over calls, or on
 * the run: partial sums that meet in arbitrary order are accumulated in fixed point.
 * Ranges: activations are stored as fp16 and saturate at +-65504 instead of overflowing; the sigma logit before its bias
 * (the alpha head's sum over 256 channels) is accumulated as a 32-bit integer in units of 2^-18, i.e. it must lie within
 * +-8192 (partial sums may wrap, the total may not); a volume density that large saturates alpha long before.
 * ---------------------------------------------------------------------------------------------- */
int nerfq_mlp_forward(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray,
                      float* raw, void* save, int max_ctas, nerfq_stream_t stream);
unsigned long long nerfq_mlp_save_bytes(long long n_points);

/* Gradient of the loss w.r.t. the LSA scales (replaces torch autograd over NeRF.forward with only
 * weight_scaling trainable: framework/pytorch_model/__init__.py:1129-1145, run_nerf.py:756).
 * d_scale [2436] is ACCUMULATED (caller zeroes).  `packed` is not const: a 19.5 KB scratch area inside it holds the
 * partial sums of the launch as 64-bit fixed point (left zeroed again on completion), so one packed network must not
 * run two backward launches concurrently.  The sums are order-independent: the result is bit-reproducible.
 * Ranges: a per-channel sum s*ds beyond +-32768 saturates (2^-48 fixed point); a channel whose LSA scale is exactly 0
 * reports gradient 0 (the kernel recovers sum dY (W x) from y - b = s (W x), which vanishes with s). */
int nerfq_mlp_backward(void* packed, const float* d_raw, const float* raw, const void* save, long long n_points,
                       float* d_scale, int max_ctas, nerfq_stream_t stream);

/* The same in two halves, for data-parallel tuning (the reference is single-GPU, README.md:76): `partial`
 * accumulates s*ds into the caller's fixed-point buffer grad_fix (DEVICE int64[nerfq_mlp_grad_fix_bytes()/8], value *
 * 2^48, zeroed by the caller once); the ranks then sum their buffers as INTEGERS (e.g. ncclInt64 / ncclSum) and `finalize`
 * converts: d_scale[i] += grad_fix[i] / 2^48 / scale[i], leaving grad_fix zeroed.  Integer sums are order-independent, so
 * N ranks x B rays give bit-identical gradients to one rank x N*B rays (given the same per-ray loss weights). */
unsigned long long nerfq_mlp_grad_fix_bytes(void);
int nerfq_mlp_backward_partial(const void* packed, const float* d_raw, const float* raw, const void* save, long long n_points,
                               long long* grad_fix, int max_ctas, nerfq_stream_t stream);
int nerfq_mlp_backward_finalize(const void* packed, long long* grad_fix, float* d_scale, nerfq_stream_t stream);

/* The same with the all-reduce fused in, for the ranks of ONE node (NVLink / NVSwitch peer access): every rank maps a
 * region of nerfq_dp_peer_bytes() bytes of every other rank (e.g. torch.distributed._symmetric_memory, cuMem + fabric/fd
 * handles), zero-initialised before first use; `peers` is a DEVICE array of `world` pointers to these regions in rank
 * order.  One launch per step on every rank replaces {all-reduce, finalize(coarse), finalize(fine)}: it publishes
 * grad_fix2 (DEVICE int64[2][nerfq_mlp_grad_fix_bytes()/8]: coarse, fine; left zeroed) in the rank's own region, exchanges
 * epoch flags with all peers, reads every rank's sums over the interconnect and accumulates the converted result into
 * d_scale2 (DEVICE float[2][2436]).  `epoch` is a DEVICE counter owned by the rank (zero at start, advanced by the kernel,
 * so the launch can sit in a CUDA graph).  Every rank must issue the same sequence of calls; a peer that never arrives
 * makes the kernel trap after a bounded wait.  packed_fine == NULL: only the coarse network carries sums. */
unsigned long long nerfq_dp_peer_bytes(void);
int nerfq_mlp_backward_finalize_peers(const void* packed_coarse, const void* packed_fine, long long* grad_fix2, void* const* peers,
                                      int world, int rank, unsigned int* epoch, float* d_scale2, nerfq_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Ray-side kernels
 * ---------------------------------------------------------------------------------------------- */

/* Stratified depths, run_nerf.py:379-403.  t_rand [n_rays,S] nullable (perturb). */
int nerfq_coarse_depths(const float* rays, const float* t_rand, long long n_rays, int S, int lindisp, float* z_out,
                        nerfq_stream_t stream);

/* raw2outputs, run_nerf.py:285-345.  noise [n_rays,S] nullable; depth, weights nullable. */
int nerfq_composite_fwd(const float* raw, const float* z, const float* rays, const float* noise, int white_bkgd,
                        long long n_rays, int S, float* rgb, float* disp, float* acc, float* depth, float* weights,
                        nerfq_stream_t stream);

/* d loss / d raw given d loss / d rgb_map (the only output the LSA objective uses, run_nerf.py:741-751). */
int nerfq_composite_bwd(const float* raw, const float* z, const float* rays, const float* noise, int white_bkgd,
                        const float* d_rgb, long long n_rays, int S, float* d_raw, nerfq_stream_t stream);

/* sample_pdf (+ concat + sort + z_std), run_nerf_helpers.py:119-163 and run_nerf.py:423-429,451.
 * Either z_coarse [n,S] (bins = midpoints; z_out [n,S+Ni] = sorted union) or bins [n,S-1] (z_out NULL).
 * weights [n,S] (interior S-2 entries are read); u [n,Ni] nullable = deterministic linspace. */
int nerfq_sample_fine(const float* z_coarse, const float* bins, const float* weights, const float* u, long long n_rays,
                      int S, int Ni, float* z_out, float* z_std, float* z_samples, nerfq_stream_t stream);

/* get_rays + viewdirs + ndc_rays + ray packing, run_nerf_helpers.py:71-115 and run_nerf.py:108-142.
 * K4 = {fx, fy, cx, cy} and c2w12 (3x4 row-major) are HOST pointers. */
int nerfq_camera_rays(int H, int W, const float* K4, const float* c2w12, int ndc, float near, float far,
                      long long first_pixel, long long count, float* rays_out, nerfq_stream_t stream);
int nerfq_pack_rays(const float* rays_o, const float* rays_d, long long n, int ndc, int H, int W, float focal, float near,
                    float far, float* rays_out, nerfq_stream_t stream);

/* render_rays, forward only, as one call (run_nerf.py:348-457 with perturb = 0 and raw_noise_std = 0 -- the test-view
 * path of render_path, run_nerf.py:161-211): stratified depths -> MLP(coarse) -> composite -> sample_pdf + sort -> MLP(fine)
 * -> composite, enqueued back to back on `stream`.  rays [n_rays,11]; outputs rgb [n,3], disp [n], acc [n] and, when
 * Ni > 0, rgb0 / disp0 / acc0 / z_std (z_std nullable).  packed_fine == NULL uses the coarse network for the fine pass
 * (N_importance > 0 without network_fine).  workspace: nerfq_render_rays_workspace_bytes(n_rays, S, Ni) bytes holding the
 * depths, raw network outputs and coarse weights between the launches (at the reference's chunk of 32768 rays 137 MB). */
unsigned long long nerfq_render_rays_workspace_bytes(long long n_rays, int S, int Ni);
int nerfq_render_rays_fwd(const void* packed_coarse, const void* packed_fine, const float* rays, long long n_rays, int S, int Ni,
                          int lindisp, int white_bkgd, void* workspace, float* rgb, float* disp, float* acc, float* rgb0,
                          float* disp0, float* acc0, float* z_std, int max_ctas, nerfq_stream_t stream);

/* to8b, run_nerf_helpers.py:14: out[i] = (uint8)(255 * clip(x[i], 0, 1)), truncating like numpy's astype. */
int nerfq_to8b(const float* x, uint8_t* out, long long n, nerfq_stream_t stream);

/* Training-batch selection (run_nerf.py:690-735: np.random.choice(H*W, N_rand, replace=False), then rays and target
 * colours of those pixels): n DISTINCT pixels of one H x W image chosen by a keyed permutation of (seed, step), their
 * packed rays generated from the pose (as nerfq_camera_rays) and their colours gathered from image [H*W,3] -- one launch,
 * no host random numbers, no synchronisation.  K4 / c2w12 HOST pointers; image and target_out nullable together;
 * index_out [n] nullable.  The permutation is this library's own (the reference's host RNG stream is not reproduced). */
int nerfq_select_batch(int H, int W, const float* K4, const float* c2w12, int ndc, float near, float far, const float* image,
                       unsigned long long seed, unsigned long long step, long long n, float* rays_out, float* target_out,
                       int* index_out, nerfq_stream_t stream);

/* img2mse (x2) and gradient, run_nerf_helpers.py:12 and run_nerf.py:741-751.  The two means are ADDED to loss2[2] (the
 * caller zeroes it; several calls accumulate a chunked batch) by one thread in a fixed order: bit-reproducible.
 * n_norm: the number of rays the mean runs over (0 = n_rays); a data-parallel rank passes the GLOBAL batch size so that
 * its loss terms and gradients are the rank's share of the global mean.
 * workspace: nerfq_mse_grad_workspace_bytes() bytes, zeroed once by the caller (the kernel leaves it re-armed); must not
 * be shared by launches that may run concurrently. */
unsigned long long nerfq_mse_grad_workspace_bytes(void);
int nerfq_mse_grad(const float* rgb, const float* rgb0, const float* target, long long n_rays, long long n_norm, float* d_rgb,
                   float* d_rgb0, float* loss2, void* workspace, nerfq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFQ_H */
