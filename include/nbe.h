/* nbe.h -- C ABI of the B200-native N-body emulator forward pass (libnbe_b200.so).
 *
 * Plain pointers and sizes only; no torch / CUDA types in the signatures (streams are passed
 * as void* holding a cudaStream_t).  Every function returns 0 (NBE_OK) or a negative error
 * code and never throws; nbe_last_error() gives the message.  The caller owns every I/O
 * buffer; the context owns the packed weights, the activation arena and its streams.  One
 * context per (GPU, host thread); contexts are independent.
 *
 * The reference (/root/reference, pure JAX/Flax) has no FFI of its own: its operator
 * boundary is `jax.jit(model.apply)` (src/jax_nbody_emulator/subbox.py:137, called at
 * :221-233) and `NBodyEmulator.apply` (nbody_emulator.py:42-79).  Each entry point below
 * cites the reference interface it stands in for; INTEGRATION.md shows the jax.ffi / ctypes
 * stub a maintainer of the reference would add.
 */
#ifndef NBE_H_
#define NBE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBE_OK 0
#define NBE_ERR_ARG (-1)      /* bad argument (shape, dtype, null pointer, state)        */
#define NBE_ERR_CUDA (-2)     /* a CUDA runtime / driver call failed                     */
#define NBE_ERR_STATE (-3)    /* call order: params / modulation missing                 */
#define NBE_ERR_UNSUPPORTED (-4)

/* element types of caller-visible buffers */
enum { NBE_F32 = 0, NBE_F16 = 1, NBE_BF16 = 2 };

/* arithmetic of the conv layers.
 * NBE_PREC_SPLIT  (default): primal activations and weights are carried as fp16 hi+lo pairs
 *   (3 tensor-core products, ~22 significant bits), tangent in fp16; fp32 accumulation in
 *   TMEM.  Needed to meet rel-L2 <= 1e-3 on the *velocity*: the LeakyReLU tangent rule
 *   (layers_vel.py:185) is discontinuous in the primal sign, so a primal rounding of
 *   relative size d flips a fraction ~d of the masks and costs ~sqrt(d) in velocity.
 * NBE_PREC_FP16: single fp16 product (displacement ~5e-4, velocity ~1e-2 rel-L2).        */
enum { NBE_PREC_SPLIT = 0, NBE_PREC_FP16 = 1 };

typedef struct nbe_ctx nbe_ctx;

/* One conv layer of the parameter tree params['params'][block][layer]
 * (nbody_emulator.py:115-129; leaf shapes tests/test_style_layers_vel.py:392-436).
 * All pointers are HOST fp32, C-contiguous.  Style models: style_weight (cin,2) and
 * style_bias (cin) set, dweight NULL.  Premodulated models: style_* NULL, weight is the
 * demodulated weight and dweight its Dz-tangent (NULL when compute_vel == 0).            */
typedef struct nbe_layer_params {
  const char* block;          /* "conv_l00", "down_l0", ... "conv_r01" */
  const char* layer;          /* "skip", "conv_0", "conv_1"            */
  const float* weight;        /* (cout, cin, k, k, k)                  */
  const float* dweight;       /* (cout, cin, k, k, k) or NULL          */
  const float* bias;          /* (cout)                                */
  const float* style_weight;  /* (cin, 2) or NULL                      */
  const float* style_bias;    /* (cin) or NULL                         */
  int32_t cout, cin, k;
} nbe_layer_params;

/* Create / destroy a per-GPU context (replaces the implicit default-device state behind
 * jax.jit at subbox.py:137). */
int nbe_create(nbe_ctx** out, int device);
void nbe_destroy(nbe_ctx* ctx);
const char* nbe_last_error(const nbe_ctx* ctx);
const char* nbe_version(void);

/* Upload the 33-layer parameter tree (replaces `params` in model.apply(params, ...),
 * subbox.py:226-233).  premodulated: 0 = Style* models, 1 = NBodyEmulator(Vel)Core.
 * compute_vel: 1 = *VelCore.  eps: demodulation epsilon (style_layers_vel.py:33).        */
int nbe_set_params(nbe_ctx* ctx, const nbe_layer_params* layers, int n_layers, int premodulated,
                   int compute_vel, float eps);
int nbe_set_precision(nbe_ctx* ctx, int precision);

/* Per-sample weight modulation + demodulation + Dz-tangent for `batch` samples
 * (style_layers_vel.py:62-105; nbody_emulator.py:131-148, :189-219), one fused kernel over
 * all layers, writing the tensor-core operand layout.  Om may be NULL for premodulated
 * models (the weights are only re-packed).  Host arrays of length `batch`.               */
int nbe_modulate(nbe_ctx* ctx, const float* Om, const float* Dz, int batch, void* stream);

/* Read back the fp32 modulated weight / dweight of one layer and sample in OIDHW order
 * (backs modulate_emulator_parameters(_vel), nbody_emulator.py:150-187, :221-266).
 * dw_host may be NULL.  Requires a prior nbe_modulate.                                   */
int nbe_get_modulated(nbe_ctx* ctx, int layer_index, int sample, float* w_host, float* dw_host);

/* model.apply on device buffers (the four *Core.__call__; e.g.
 * style_nbody_emulator_vel_core.py:105-195): x_dev (batch,3,n0,n1,n2) NCDHW of in_dtype,
 * outputs (batch,3,n0-96,n1-96,n2-96) NCDHW of out_dtype.  vel_dev / vel_fac must be
 * non-NULL iff the parameters were set with compute_vel.  Weights come from the last
 * nbe_modulate (its batch must equal `batch`, or 1 = shared).  Asynchronous on `stream`. */
int nbe_forward(nbe_ctx* ctx, const void* x_dev, int in_dtype, int batch, const int32_t dims[3],
                const float* Dz, const float* vel_fac, void* disp_dev, void* vel_dev, int out_dtype,
                void* stream);

/* SubboxProcessor.process_box (subbox.py:139-219) on HOST buffers for subboxes
 * [sub_first, sub_first+sub_count): uploads the input box (3,size) once, gathers each
 * periodic padded crop on the GPU with the reference's integer tables
 * (SubboxConfig._get_crop_inds, subbox.py:81-97; `crop_idx`/`add_idx` below are those tables,
 * computed by the caller), runs the net, pastes into a device slab and copies the finished
 * slab back into disp_host / vel_host (3,size) of out_dtype.  Only the output voxels owned
 * by the given subboxes are written.  Synchronous.
 * crop_idx: int32 [n_sub_total][3][pad_len] with pad_len_d = crop_d + pad_lo_d + pad_hi_d
 *           stored as three consecutive arrays per subbox (lengths plen[0..2]);
 * add_idx0: int32 [n_sub_total][3] first output index per dim (add indices are contiguous,
 *           tests/test_subbox.py:167-180).                                               */
int nbe_process_box(nbe_ctx* ctx, const void* in_host, int in_dtype, const int32_t size[3],
                    const int32_t crop[3], const int32_t plen[3], const int32_t* crop_idx,
                    const int32_t* add_idx0, int sub_first, int sub_count, float Dz, float vel_fac,
                    void* disp_host, void* vel_host, int out_dtype);

/* nbe_process_box over ALL GPUs of the process in one call (what the reference user's single
 * processor.process_box(box, z, Om) call, subbox.py:139-219, becomes on an 8-GPU box): ctxs[0..ngpu) are
 * contexts on distinct devices, each with the same parameters set and modulated.  The range
 * [sub_first, sub_first+sub_count) is cut into ngpu contiguous shares (the first sub_count % ngpu
 * contexts get one more subbox); one host thread per context uploads its (D, H) window from the SAME
 * in_host and copies its finished blocks into the SAME disp_host / vel_host, which should be
 * page-locked (nbe_host_register, or cudaHostAlloc'ed by the caller).  The shares are disjoint: no
 * exchange step, no per-GPU copy of the box in host memory.  Synchronous; returns the first error
 * (message on ctxs[0]).
 * out_size0: D extent of disp_host / vel_host, 0 = size[0].  The tables are plain gather / paste
 * indices into whatever arrays are passed, so a box too large for host or device memory is streamed
 * as D-slabs: in_host holds the crop[0]+96 planes of one slab of subboxes (size[0] = that count, D
 * tables = 0..size[0]-1), the outputs its crop[0] planes (out_size0 = crop[0], D anchors 0) --
 * SubboxProcessor.process_box_streamed, BASELINE config 5.                                        */
int nbe_process_box_multi(nbe_ctx** ctxs, int ngpu, const void* in_host, int in_dtype, const int32_t size[3],
                          const int32_t crop[3], const int32_t plen[3], const int32_t* crop_idx,
                          const int32_t* add_idx0, int sub_first, int sub_count, float Dz, float vel_fac,
                          void* disp_host, void* vel_host, int out_dtype, int32_t out_size0);

/* nbe_process_box with the results left ON THE DEVICE in block layout: disp_blocks_dev / vel_blocks_dev
 * are (sub_count, 3, crop0, crop1, crop2) of out_dtype, one contiguous record per subbox in index order.
 * This is the send buffer of the optional output gather of a sharded box (one process per GPU):
 * ncclAllGather of the records over NVLink, no host round trip (SubboxProcessor.process_box(gather=...)). */
int nbe_process_box_blocks(nbe_ctx* ctx, const void* in_host, int in_dtype, const int32_t size[3],
                           const int32_t crop[3], const int32_t plen[3], const int32_t* crop_idx,
                           const int32_t* add_idx0, int sub_first, int sub_count, float Dz, float vel_fac,
                           void* disp_blocks_dev, void* vel_blocks_dev, int out_dtype);

/* Same decomposition with the box and the outputs resident in device memory ((3,size) each);
 * asynchronous on `stream`, no host<->device traffic besides the index tables.            */
int nbe_process_box_dev(nbe_ctx* ctx, const void* box_dev, int in_dtype, const int32_t size[3],
                        const int32_t crop[3], const int32_t plen[3], const int32_t* crop_idx,
                        const int32_t* add_idx0, int sub_first, int sub_count, float Dz, float vel_fac,
                        void* disp_dev, void* vel_dev, int out_dtype, void* stream);

/* Page-lock (cudaHostRegister) / unlock a caller-owned host buffer used with nbe_process_box.
 * nbe_host_register returns 1 if it registered the buffer, 0 if it already was page-locked.  */
int nbe_host_register(nbe_ctx* ctx, void* ptr, size_t bytes);
int nbe_host_unregister(nbe_ctx* ctx, void* ptr);

/* ---- the step after the path: displacement -> density contrast -> P(k) shells ----
 * Replaces the reference's third-party calls dj.get_delta_from_psi(psi, method="pm", res, worder,
 * deconvolve) (scripts/core.py:396-409, 446-458) and PKL.Pk(delta, boxsize, MAS=...)
 * (scripts/utils.py:1083-1090).  The FFT between the calls is the caller's (cuFFT): delta_k is the
 * half-complex cube (res, res, res/2+1) of float2, unnormalised.
 *
 * nbe_density_from_psi: particles of the (n0,n1,n2) lattice displaced by psi_dev (3,n0,n1,n2) fp32,
 * in the units of boxsize, are assigned to a periodic res^3 mesh with the B-spline of order worder
 * (1 NGP, 2 CIC, 3 TSC, 4 PCS); delta_dev (res^3 fp32) receives rho/mean(rho) - 1.
 * nbe_mas_deconvolve: delta_k /= prod_i sinc(pi k_i/res)^worder, in place.
 * nbe_pk_bins: out_dev[3][nbins] (double; zeroed here) = shell sums of |delta_k / W^mas_order|^2,
 * of |k|/k_F and of the mode count over the independent modes, shell = floor(|k|/k_F).           */
int nbe_density_from_psi(nbe_ctx* ctx, const float* psi_dev, const int32_t n[3], float boxsize, int32_t res,
                         int32_t worder, float* delta_dev, void* stream);
int nbe_mas_deconvolve(nbe_ctx* ctx, void* delta_k_dev, int32_t res, int32_t worder, void* stream);
int nbe_pk_bins(nbe_ctx* ctx, const void* delta_k_dev, int32_t res, int32_t mas_order, int32_t nbins,
                double* out_dev, void* stream);

/* ---- the step before the path: linear density -> Zel'dovich displacement ----
 * Replaces dj.with_lpt(n_order=1) + dj.evaluate_lpt_psi_at_a (scripts/core.py:396-397).
 * psi_k_dev receives three half-complex cubes psi_j(k) = i k_j / k^2 * delta(k) (k = 0 and the
 * Nyquist component of each derivative zeroed); the caller's inverse FFTs give psi in the units
 * of boxsize when delta_k is the unnormalised forward FFT and the inverse is normalised.       */
int nbe_za_psi_k(nbe_ctx* ctx, const void* delta_k_dev, int32_t res, float boxsize, void* psi_k_dev, void* stream);

/* Free the activation arena, cached plans and box staging buffers of the context (re-created on demand;
 * parameters and packed weights stay): for processes that share a GPU or alternate between shapes.   */
int nbe_release_workspace(nbe_ctx* ctx);

/* Bytes of device memory the context needs for one (n0,n1,n2) sample (activation arena). */
size_t nbe_workspace_bytes(nbe_ctx* ctx, const int32_t dims[3]);

/* Instrumentation: kernels launched by this context since the last reset, and per-launch
 * device timings: while profiling is enabled every launch is bracketed by CUDA events on the
 * launching stream (no synchronisation); nbe_get_profile resolves them and returns, per
 * launch slot, the name, the mean duration in ms over all samples run since profiling was
 * enabled and the algorithmic FLOPs of one launch.  Arrays hold up to `cap` entries; returns
 * the number of slots.                                                                    */
int64_t nbe_launch_count(nbe_ctx* ctx, int reset);
int nbe_set_profiling(nbe_ctx* ctx, int enable);
int nbe_get_profile(nbe_ctx* ctx, int cap, const char** names, float* ms, double* flops);

/* Debug aid: read back activation tensor `act` (internal id, NDHWC fp16; which: 0 hi, 1 lo,
 * 2 tangent) of the most recent plan.  host == NULL queries the size.                     */
long long nbe_debug_read_act(nbe_ctx* ctx, int act, int which, void* host, size_t cap, int32_t shape_out[4]);

/* Tangent folding: the reference's modulated tangent weights are
 * dW = W (.) (a_i + beta_o) (style_layers_vel.py:86-93), so x*dW + dx*W = (dx + a (.) x)*W + beta (.) (x*W).
 * When active (default; NBE_FOLD=0 disables, and a modulation close to zero switches it off by itself) the
 * 64-output 3^3 launches run 4 tensor-core products per layer instead of 5 and the stored tangent of the tensors
 * they read is dx' = dx + a (.) x.  nbe_fold_active reports the state after the last nbe_modulate;
 * nbe_debug_act_fold copies the fold vector a of activation `act` (zeros if it has none) and returns its
 * channel count.                                                                           */
int nbe_fold_active(nbe_ctx* ctx);
int nbe_debug_act_fold(nbe_ctx* ctx, int act, int sample, float* a_host, int cap);

#ifdef __cplusplus
}
#endif
#endif /* NBE_H_ */
