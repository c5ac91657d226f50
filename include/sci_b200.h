/*
 * sci_b200.h — C ABI of the B200-native AdaptivePnP_SCI hot path.
 *
 * The reference (xyvirtualgroup/AdaptivePnP_SCI) is pure Python/PyTorch and has
 * no FFI of its own; the boundary of the hot path is the set of Python call
 * signatures listed in SURVEY.md §8(b).  This header is the C-ABI a binding
 * (ctypes, see adaptivepnp_sci_b200/_lib.py and INTEGRATION.md) loads to replace
 * the ATen/cuDNN/skimage op sequences behind those signatures.  Every entry
 * point cites the reference code it replaces (paths relative to the reference).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless
 *     the name ends in _host; tensors stay owned by the caller (PyTorch);
 *   - stream-ordered: work is enqueued on `stream` (a cudaStream_t passed as
 *     void*), nothing synchronises, nothing allocates; kernels that need
 *     scratch take a workspace pointer whose size comes from a *_workspace_bytes
 *     query;
 *   - return 0 on success, a negative SCI_E* code on failure (never throws);
 *   - re-entrant across streams/devices (no global mutable state except the
 *     per-device TMA driver entry point cache).
 *
 * Device layouts ("planar"): the reference keeps cubes pixel-major
 * ([H,W,B], [h,w,B,4], [H,W,3,B]).  On the device every cube is FRAME-PLANAR
 * fp32:  Bayer-domain cubes  [B][H][W]   (full-resolution mosaic; the four
 * RGGB phases of the reference's [h,w,B,4] cubes are the (row&1,col&1) classes),
 * RGB cubes [B][3][H][W], planes [H][W].  Because the sensing operators are
 * pixel-wise, the per-phase loops of the reference (for ib in range(4)) are one
 * pass over the full-resolution planes.  sci_pixlast_to_planar /
 * sci_planar_to_pixlast convert at the boundary (bit-exact index remaps).
 */
#ifndef SCI_B200_H
#define SCI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCI_OK              0
#define SCI_EINVAL         -1   /* bad argument (null pointer, odd size, ...) */
#define SCI_ELAUNCH        -2   /* CUDA launch / runtime error               */
#define SCI_EUNSUPPORTED   -3   /* shape not supported by this build         */
#define SCI_EWORKSPACE     -4   /* workspace too small                       */

/* ABI version (major*1000+minor). */
int sci_version(void);
/* Last CUDA error string recorded by a failing call on this thread. */
const char* sci_last_error(void);

/* ---- layout remaps (bit-exact) -------------------------------------------
 * Replaces the strided-copy loops at dvp_linear_inv_2_stage_ADMM_tensor_online.py:66-82,
 * :170-172, :275-277, :312-314 and utils/utils_image.py:130-171.
 * pixel-last  in[P][C][B]  <->  planar out[B][C][P]   (C=1: Bayer cube, C=3: RGB cube) */
int sci_pixlast_to_planar(const float* in, float* out, int P, int C, int B, void* stream);
int sci_planar_to_pixlast(const float* in, float* out, int P, int C, int B, void* stream);

/* ---- K0: Bayer split + Phi-sum + initialisation --------------------------
 * dvp...online.py:59-82 / :347-370.  phi_hwb [H][W][B] (reference layout) ->
 * phi [B][H][W]; phisum [H][W] = sum_t phi, zeros replaced by 1 (:73);
 * theta0 [B][H][W] = y*phi_t (At_, utilspy.py:35-44) or, when x0_hwb != NULL,
 * the warm start x0_hwb [H][W][B] transposed (:82). */
int sci_bayer_split_init(const float* y, const float* phi_hwb, const float* x0_hwb,
                         float* phi, float* phisum, float* theta0,
                         int H, int W, int B, void* stream);

/* ---- sensing operators on arbitrary-stride views -------------------------
 * utilspy.py:28-33 (A_) and :35-44 (At_); strides in ELEMENTS so the
 * reference's Bayer-phase views ([h,w,B] slices of [h,w,B,4]) are accepted. */
int sci_A(const float* x, long sxh, long sxw, long sxt,
          const float* phi, long sph, long spw, long spt,
          float* y, long syh, long syw, int h, int w, int B, void* stream);
int sci_At(const float* y, long syh, long syw,
           const float* phi, long sph, long spw, long spt,
           float* x, long sxh, long sxw, long sxt, int h, int w, int B, void* stream);

/* ---- K1/K2: fused Euclidean projection ------------------------------------
 * stage 1 (dvp...online.py:389-391):  v = theta + b ;
 *     x = v + lambda * phi * ((y - sum_t phi*v) / (phisum + gamma))
 * stage 2 (dvp...online.py:128-140):  p = theta - (1/rho) b ;
 *     x = p + phi * ((y - sum_t phi*p) / (alpha*rho + phisum))
 * All cubes planar [B][npix].  If orig != NULL also accumulates
 * sse[0] += sum (x - orig)^2 (fp32 terms, fp64 sum) for the per-iteration PSNR
 * of stage 1 (:507-512). */
int sci_project_stage1(const float* theta, const float* b, const float* phi, const float* y,
                       const float* phisum, float* x, long npix, int B, float lambda_, float gamma,
                       const float* orig, double* sse, void* stream);
int sci_project_stage2(const float* theta, const float* b, const float* phi, const float* y,
                       const float* phisum, float* x, long npix, int B, float alpha, float rho,
                       void* stream);

/* ---- K4: TV prior (Chambolle projection) + fused dual update --------------
 * Replaces the D2H -> skimage.restoration.denoise_tv_chambolle(weight, n_iter_max=5,
 * multichannel=True) -> H2D round trip at dvp...online.py:153-160 / :403-407 and
 * the clip + dual update at :265-267 / :501-503.
 * Each of the 4B channels (frame t, Bayer phase) is an independent 2-D image of
 * (H/2)x(W/2) pixels living at stride 2 inside plane t.
 *     f     = x + c_b * b            (b may be NULL: f = x)
 *     theta = TV(f), clipped to [0,1] if clip != 0
 *     b_out = b + s_b * (x - theta)  (skipped if b == NULL)
 * b_out must not alias b (the early-stop fix-up pass re-reads b).
 * nstop_out (optional, may be NULL): int[4B], the iteration 1..4 whose iterate
 * was returned for channel t*4+phase.  workspace: sci_tv_workspace_bytes(). */
size_t sci_tv_workspace_bytes(int H, int W, int B);
int sci_tv_chambolle2d(const float* x, const float* b, float c_b, float* theta, float* b_out,
                       float s_b, int clip, int H, int W, int B, float weight, float eps,
                       int n_iter_max, void* workspace, size_t workspace_bytes, int* nstop_out,
                       void* stream);

/* ---- K6: Malvar-2004 demosaic + (-w/tau) ----------------------------------
 * packages/colour_demosaicing/bayer/demosaicing/malvar2004.py:169-246 applied to
 * every frame (dvp...online.py:186-191) of the mosaic  m = x + c_b*b  (b may be
 * NULL), followed by x_rgb_w = x_rgb - inv_tau*w (:198).  Outputs planar
 * x_rgb [B][3][H][W] and, if w != NULL, u [B][3][H][W]. */
int sci_malvar2004(const float* x, const float* b, float c_b, const float* w, float inv_tau,
                   float* x_rgb, float* u, int H, int W, int B, void* stream);

/* ---- K3/K5: Bayer sampling of the denoised RGB + dual updates --------------
 * dvp...online.py:206-209 (theta <- RGGB samples of xhat), :265 (clip), :267
 * (b += x - theta), :271 (w += x_rgb - xhat), :275-280 (PSNR of merged theta).
 * first_iter != 0 reproduces the k=0 aliasing of xall/theta_all (:87-89): x is
 * replaced by the unclipped samples, so b receives theta_unclipped - theta. */
int sci_dual_update_rgb(const float* xhat, const float* x_rgb, float* w, const float* x,
                        float* b, float* theta, int first_iter, int H, int W, int B,
                        const float* orig, double* sse, void* stream);

/* ---- K5: RGB cube -> Bayer samples / sparse 3-channel mosaic --------------
 * sci_rgb_to_bayer: dvp...online.py:206-209, packages/fastdvdnet/utils.py:69-78
 *   rgb [B][3][H][W] -> mosaic [B][H][W].
 * sci_bayer_to_rgb_sparse: utils/utils_image.py:153-161 (oneCh2ThreeCh). */
int sci_rgb_to_bayer(const float* rgb, float* mosaic, int H, int W, int B, void* stream);
int sci_bayer_to_rgb_sparse(const float* mosaic, float* rgb, int H, int W, int B, void* stream);

/* Reference-layout Bayer stack [h][w][B][4] <-> mosaic [H][W][B]
 * (utils/utils_image.py:130-151 fourCh2OneCh / oneCh2FourCh; B=1 covers the 3-D form). */
int sci_bayer4_to_mosaic(const float* stack, float* mosaic, int h, int w, int B, void* stream);
int sci_mosaic_to_bayer4(const float* mosaic, float* stack, int h, int w, int B, void* stream);

/* ---- K11: PSNR accumulation -------------------------------------------------
 * dvp...online.py:279,320: sse_per_frame[t] += sum_p (a[t][p]-orig[t][p])^2,
 * fp32 difference and square, fp64 accumulation (skimage PSNR restated). */
int sci_psnr_accum(const float* a, const float* orig, long npix, int B, double* sse_per_frame,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCI_B200_H */
