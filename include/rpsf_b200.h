/* rpsf_b200.h — C ABI of the B200-native regularizepsf correction path.
 *
 * The reference (punch-mission/regularizepsf) is pure Python and has no FFI; its boundary for
 * this path is the public Python API.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference repository root).  The Python mirror of that API
 * (regularizepsf_b200/transform.py, psf.py) is the only in-tree caller and binds these symbols
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success or a negative RPSF_E_* code;
 * rpsf_last_error() gives the message for the calling thread.  No C++ exceptions cross the
 * ABI.  Device pointers are caller-owned; `stream` is a cudaStream_t passed as void* (NULL =
 * legacy default stream).  Stream-taking calls enqueue work and return without
 * synchronising.  A plan owns one spectrum workspace, so use one plan per concurrent stream.
 */
#ifndef RPSF_B200_H
#define RPSF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPSF_ABI_VERSION 1

/* status codes */
#define RPSF_OK 0
#define RPSF_E_INVALID_ARGUMENT (-1)
#define RPSF_E_INVALID_COORDINATE (-2) /* maps to regularizepsf.exceptions.InvalidCoordinateError */
#define RPSF_E_INCORRECT_SHAPE (-3)    /* maps to regularizepsf.exceptions.IncorrectShapeError   */
#define RPSF_E_UNSUPPORTED (-4)        /* patch size / dtype / pad mode without a device path    */
#define RPSF_E_CUDA (-5)
#define RPSF_E_NO_KERNEL (-6)          /* apply before rpsf_transform_set_kernel                 */

/* element types */
#define RPSF_F32 0
#define RPSF_F64 1
#define RPSF_U8 2
#define RPSF_I16 3
#define RPSF_U16 4
#define RPSF_I32 5
#define RPSF_I64 6
#define RPSF_U32 7

/* np.pad modes with an on-device index map (transform.py:119-123, `pad_mode`) */
#define RPSF_PAD_SYMMETRIC 0
#define RPSF_PAD_REFLECT 1
#define RPSF_PAD_EDGE 2
#define RPSF_PAD_WRAP 3
#define RPSF_PAD_CONSTANT 4
/* every other np.pad mode (mean, median, maximum, minimum, linear_ramp, ...): the caller materialises the padded frame
 * (np.pad of 2P per side, as transform.py:119-123 does) and passes a pointer to the UNPADDED pixel (0, 0) inside it;
 * rows / columns in [-2P, dim + 2P) must then be addressable through the pitch.  Device-pointer entry points only. */
#define RPSF_PAD_MATERIALIZED 5

typedef struct rpsf_transform rpsf_transform;
typedef struct rpsf_plan rpsf_plan;

int rpsf_abi_version(void);
const char* rpsf_last_error(void);
/* 1: patch size P runs natively (powers of two 16..512); 2: it runs embedded in the next power of two M >= 2 P - 1
 * (2 <= P <= 256: the transfer kernel is re-sampled on the M-grid so that the M-point circular convolution of the
 * zero-padded windowed patch reproduces the reference's P-point one, transform.py:151-165 — same results, (M / P)^2
 * the work, the generic gather and colour-phase overlap-add kernels); 0: no device path */
int rpsf_patch_size_supported(int patch_size);

/* ---- transform: replaces ArrayPSFTransform.__init__ (transform.py:28-37) -------------------
 * coords: n x 2 host int32 (row, col) upper-left patch corners in unpadded frame coordinates,
 * in IndexedCube order (util.py:56-82).  compute_dtype: RPSF_F32 or RPSF_F64 (validation mode). */
int rpsf_transform_create(rpsf_transform** out, const int32_t* coords, int n_patches, int patch_size,
                          int compute_dtype, int device);
/* The same for one rank of a patch-row slab split ("holds only those patches' kernels"): `keep`
 * (n_patches bytes, may be NULL = all) marks the patches whose transfer kernel this device will hold.  The whole
 * coordinate list is still given, so colour classes — and with them the per-pixel summation order — are those of the
 * complete transform and sharded results stay bit-identical.  rpsf_transform_set_kernel then takes the kept patches
 * only, in list order; a plan whose row band needs a patch that was not kept fails with RPSF_E_INVALID_ARGUMENT. */
int rpsf_transform_create_subset(rpsf_transform** out, const int32_t* coords, int n_patches, int patch_size,
                                 int compute_dtype, int device, const uint8_t* keep);
int rpsf_transform_destroy(rpsf_transform* t);

/* Load the transfer kernel: `kernel_full` is the (N,P,P) complex cube in the reference layout
 * (IndexedCube.values of ArrayPSFTransform._transfer_kernel, transform.py:37,82; full unshifted
 * spectrum) on the device, complex64 (RPSF_F32) or complex128 (RPSF_F64).  Builds the private
 * Hermitian-half, 1/P^2-scaled, register-ordered copy that rpsf_apply reads. */
int rpsf_transform_set_kernel(rpsf_transform* t, const void* kernel_full, int kernel_dtype, void* stream);

/* number of colour classes of the overlap graph (4 for calculate_covering inputs, util.py:10-53) */
int rpsf_transform_num_colours(const rpsf_transform* t);

/* ---- construct: replaces the arithmetic of ArrayPSFTransform.construct (transform.py:78-82) -
 * source_fft / target_fft / kernel_out: `count` complex elements on the device, all complex64
 * (dtype RPSF_F32) or all complex128 (RPSF_F64).  The coordinate check (transform.py:74-76)
 * stays on the host. */
int rpsf_construct_kernel(const void* source_fft, const void* target_fft, void* kernel_out, int64_t count,
                          int dtype, double alpha, double epsilon, int device, void* stream);

/* ---- PSF cube FFT: replaces scipy.fft.fft2 in ArrayPSF.__init__ (psf.py:216-219) -----------
 * values: (n, P, P) real on the device; out: (n, P, P) complex, full spectrum, same precision. */
int rpsf_psf_fft2(const void* values, void* out, int64_t n_patches, int patch_size, int dtype, int device,
                  void* stream);

/* ---- builder: per-cell averaging of star cutouts -------------------
 * Replaces regularizepsf/builder.py:45-125 (_average_patches and its mean / percentile helpers):
 * every cutout is divided by its centre pixel [P/2][P/2] (builder.py:63,90); cell c stacks the
 * cutouts cell_items[cell_offsets[c] .. cell_offsets[c+1]) in that order (the reference's dict
 * insertion order; the matching of builder.py:45-51 stays on the host) and reduces the stack per
 * pixel: NaN-aware mean (np.nansum / count of finite values), np.nanmedian, or np.nanpercentile
 * (linear).  NaN results (empty or all-NaN stacks, 0/0) become 0 (builder.py:118-122).  float64,
 * bit-identical to numpy.  cutouts: device (n_cutouts, P, P) float64; out: device (n_cells, P, P)
 * float64; the index arrays are HOST pointers.  percentile in [0, 100] (RPSF_AVG_PERCENTILE only;
 * 50 is computed as the median, as builder.py:79-82 does).  Stream-ordered: the index arrays are
 * copied before the call returns, temporaries are cudaMallocAsync / cudaFreeAsync on `stream`. */
#define RPSF_AVG_MEAN 0
#define RPSF_AVG_MEDIAN 1
#define RPSF_AVG_PERCENTILE 2
int rpsf_average_patches(const double* cutouts, int64_t n_cutouts, int patch_size, const int64_t* cell_offsets,
                         const int32_t* cell_items, int64_t n_cells, int method, double percentile, double* out,
                         int device, void* stream);

/* The per-patch stages that follow the averaging in ArrayPSFBuilder.build (builder.py:236-258), one CTA per patch,
 * float64, device pointers, stream-ordered (scratch is cudaMallocAsync / cudaFreeAsync on `stream`):
 *   rpsf_plane_background  image_processing.py:13-46 (calculate_background): the least-squares plane
 *                          c0 * col + c1 * row + c2 through the ring of pixels just inside the patch border that are
 *                          fainter than the centre; NaN where the patch is exactly 0.  out: (n, P, P).
 *   rpsf_isolate_cores     builder.py:239-258 in place on (n, P, P): subtract that plane, zeros -> NaN, drop the pixels
 *                          below 0.5 % of the centre (eroded, outside counted as faint), non-finite -> 0, keep the
 *                          4-connected island of the centre pixel dilated by one pixel, divide by the sum.
 * Every mask is an integer decision and follows the reference; the plane is solved from the 3 x 3 normal equations
 * (minimum-norm when the ring is degenerate, like LAPACK gelsd) and the sum is a tree, so values agree with the
 * reference to ~1e-13 of the patch maximum rather than bit for bit (tests/test_gpu_builder.py states 1e-11). */
int rpsf_plane_background(const double* patches, int64_t n_patches, int patch_size, double* out, int device, void* stream);
/* The per-star body of _find_patches (image_processing.py:78-121), one CTA per detected star: the width x width window
 * of the frame at the rounded corner through np.pad(mode="reflect"), scipy.ndimage.shift(order=3, mode="mirror") by
 * (-corner + round(corner) - 0.5) (cubic B-spline prefilter + 4 x 4 interpolation, rounded to the frame's dtype as
 * scipy does), minus its plane background (NaN where the shifted patch is 0); accepted[i] = every pixel below
 * saturation_threshold (a NaN fails) and star_minimum < centre < star_maximum.  frame: device (H, W) float32 / float64,
 * contiguous; corners: HOST (n, 2) float64 = the reference's dict keys (row - width/2, col - width/2); out: device
 * (n, width, width) float64; accepted: device n bytes.  Values agree with scipy to ~1e-15 of the patch maximum (float64
 * frames).  The pixel-mask patch (:102-106, :119) is left to the caller: the reference casts its spline-shifted values
 * to bool by truncation, which only scipy's own instruction order reproduces.  Synchronises `stream` before returning
 * (the corner list is read from pageable memory). */
int rpsf_star_cutouts(const void* frame, int frame_dtype, int height, int width, const double* corners, int64_t n_stars,
                      int cutout_size, double saturation_threshold, double star_minimum, double star_maximum, double* out,
                      unsigned char* accepted, int device, void* stream);
int rpsf_isolate_cores(double* patches, int64_t n_patches, int patch_size, int device, void* stream);

/* ---- plan: geometry of apply() for one frame shape ------------------------------------------
 * Replaces the padding / slicing bookkeeping of ArrayPSFTransform.apply (transform.py:119-123,
 * 141-149, 167-177).  [row_begin,row_end) is the band of output rows this plan owns (0,H for
 * the whole frame; a sub-range for patch-row slabs across GPUs).  max_batch frames share the
 * workspace. */
int rpsf_plan_create(rpsf_plan** out, rpsf_transform* t, int height, int width, int pad_mode, int row_begin,
                     int row_end, int max_batch);
int rpsf_plan_destroy(rpsf_plan* p);
/* info[0]=active patches, [1]=colours, [2]=workspace bytes, [3]=first frame row read,
 * [4]=one past the last frame row read, [5]=1 if colour 0 tiles the band exactly,
 * [6]=overlap-add kernel: 2 = streaming chains (registers only), 1 = single-launch row-pair gather
 *      through a shared-memory plane, 0 = one launch per colour class; + 8 when apply runs as the fused
 *      persistent pipeline (rpsf_plan_set_fused),
 * [7]=teams per CTA of the row-pair gather */
int rpsf_plan_info(const rpsf_plan* p, int64_t info[8]);
/* overlap-add kernel choice: 0 = automatic (streaming chains when every owned row pair is a chain of
 * half-overlapping patch columns — any calculate_covering grid —, else the row-pair gather when patch
 * corner rows share one parity, else colour phases); 1 = force the colour-phase kernel, 2 = force the
 * row-pair gather (test hooks) */
int rpsf_plan_set_overlap_mode(rpsf_plan* p, int mode);
/* gather + row-FFT kernel choice: 0 = automatic (the persistent bulk-copy kernel), 1 = force the
 * one-row-pair-per-team kernel with direct loads (test hook; same arithmetic up to rounding) */
int rpsf_plan_set_gather_mode(rpsf_plan* p, int mode);

/* Patches whose packed half-spectrum fits one SM's shared memory (P <= 128 in float32, P <= 64 in float64) and whose
 * corners lie on multiples of P/2 (every calculate_covering grid) take a two-launch path: one CTA carries a patch of a
 * frame through gather, window, both FFT axes, the transfer multiply, both inverse axes and the second window without
 * its spectrum ever leaving shared memory, writes the corrected P x P plane, and an elementwise kernel adds the planes
 * that cover each output tile in list order — the reference's own `+=` order (transform.py:167-169).  Measured at
 * 1024^2 / 128 px: 29.0 us per frame at 8 frames per call against 22.2 for the three kernels, 39.8 against 41.0 for one
 * frame: opt-in.  0 = automatic and 1 = never select the three kernels, 2 = this path or RPSF_E_UNSUPPORTED;
 * RPSF_SMALL=1 makes 2 the default of new plans that have the tables. */
int rpsf_plan_set_small_mode(rpsf_plan* p, int mode);

/* column pass choice.  2 = the PAIRED pass (k2_chain; coverings, else RPSF_E_UNSUPPORTED): a CTA walks the patches that
 * share a corner column top to bottom and writes, per band of P/2 output rows, the row-windowed sum of the two patches
 * that overlap there, so the column pass writes half as much and the overlap-add reads half as much.  Measured at 2048^2
 * / 256 px: 21 % less DRAM traffic, 2-5 % more throughput from 8 frames per call, slower below; the chain kernel is no
 * longer HBM-bound (0.54 of the peak), so 0 = automatic and 1 = classic both select the in-place column pass with the
 * pair sum in the overlap-add kernel.  The two agree to rounding (different association of the same sum).
 * RPSF_PAIRED=1 makes 2 the default of new plans that have the tables.
 * rpsf_plan_column_info: info[0] = 1 if the next apply of this plan takes the paired pass, info[1] = bytes of the
 * paired workspace per frame (0 if the plan has none). */
int rpsf_plan_set_column_mode(rpsf_plan* p, int mode);
int rpsf_plan_column_info(const rpsf_plan* p, int64_t info[2]);

/* pipeline choice.  2 = run apply as ONE persistent cooperative launch whose spectrum hand-overs stay in L2
 * (rpsf_fused.cuh; coverings with 256-px patches in float32 — RPSF_E_UNSUPPORTED otherwise): DRAM traffic falls from
 * 346 to 121 MB per 2048^2 frame, bit-identical results, but measured slower than the three stand-alone kernels
 * (DESIGN.md section 4), which 0 = automatic and 1 = never therefore both select today.  RPSF_FUSED=1 in the
 * environment makes 2 the default of new plans, RPSF_FUSED=0 leaves the pipeline's tables out of the plan. */
int rpsf_plan_set_fused(rpsf_plan* p, int mode);
/* pipeline statistics of the fused launch (diagnostics; a few atomics per wait when enabled).  Reads and resets the
 * counters into out[8] (may be NULL), then switches collection on or off: [0..2] SM cycles the roles K1 / K2 / K3 spent
 * waiting on a producer / consumer counter (one sample per warp or column group), [3..5] cycles from role start to
 * role end summed over the role's warps, [6] column units whose tile could not be prefetched, [7] column units. */
int rpsf_plan_fused_stats(rpsf_plan* p, int enable, uint64_t out[8]);
/* timeline of the fused launch (diagnostics): reads into out[3][max_batch * n_bands] the %globaltimer nanoseconds at
 * which every band (frame-major sequence) was completed by role K1, by K2 and freed by K3 during the LAST launch, then
 * clears the record and switches collection on or off.  out may be NULL. */
int rpsf_plan_fused_trace(rpsf_plan* p, int enable, uint64_t* out, int64_t out_len);

/* ---- saturation arguments of apply (transform.py:88-90,125-138,171-172) ---------------------
 * threshold = +inf (the default) disables the branch.  Otherwise every apply on this plan pads
 * the frame 2P per side, masks padded > threshold, dilates the mask `dilation` times with the
 * 3x3 cross (dilation < 1: until stable, as scipy.ndimage.binary_dilation), replaces masked
 * pixels IN RASTER ORDER by the nanmean of padded[i-w//2:i+w//2, j-w//2:j+w//2], corrects the
 * filled frame and writes the raw pixel back into every masked output pixel.  Needs the whole
 * frame resident (img_row0 = 0, img_rows = height). */
int rpsf_plan_set_saturation(rpsf_plan* p, double threshold, int dilation, int neighborhood_width);

/* ---- apply, device-resident: replaces ArrayPSFTransform.apply (transform.py:85-177) ---------
 * image: `batch` frames of compute-dtype pixels; row `img_row0 + i` of frame b is at
 *        image + b*img_frame_stride + i*img_pitch (elements); rows [info[3], info[4]) must be
 *        resident.  out: rows [row_begin,row_end) are written; row r of frame b is at
 *        out + b*out_frame_stride + (r - out_row0)*out_pitch. */
int rpsf_apply(rpsf_plan* p, const void* image, int64_t img_pitch, int64_t img_frame_stride, int img_row0,
               int img_rows, void* out, int64_t out_pitch, int64_t out_frame_stride, int out_row0, int batch,
               void* stream);

/* ---- apply, host buffers (the call a reference user makes): H2D + convert + apply + D2H ------
 * image: `batch` contiguous (H, W) frames of image_dtype on the host (pinned or pageable);
 * out: `batch` contiguous (row_end-row_begin, W) frames of out_dtype (RPSF_F32 / RPSF_F64).
 * Synchronous on return. */
int rpsf_apply_host(rpsf_plan* p, const void* image, int image_dtype, void* out, int out_dtype, int batch);

/* ---- helpers -------------------------------------------------------------------------------- */
/* 2-D dtype conversion on the device (the `.astype(float)` of transform.py:117) */
int rpsf_convert(const void* src, int src_dtype, int64_t src_pitch, void* dst, int dst_dtype, int64_t dst_pitch,
                 int rows, int cols, int device, void* stream);
/* stage outputs for tests: copy the plan's spectrum workspace pointer / size */
int rpsf_plan_workspace(const rpsf_plan* p, void** ptr, int64_t* bytes);
/* run only the first `stages` kernels of apply (1 = K1, 2 = K1+K2, 3 = all); test hook */
int rpsf_apply_stages(rpsf_plan* p, const void* image, int64_t img_pitch, int64_t img_frame_stride, int img_row0,
                      int img_rows, void* out, int64_t out_pitch, int64_t out_frame_stride, int out_row0,
                      int batch, int stages, void* stream);
/* ---- fused output gather for patch-row slabs across GPUs (SURVEY.md sections 5 and 8e) ----------
 * The reference has no counterpart (it is single-process).  With mirrors set, the overlap-add kernel of
 * rpsf_apply stores every owned pixel to `out` AND to the same position of each mirror buffer — the other
 * ranks' full frames, mapped into this process with the IPC calls below, written over NVLink while the
 * kernel runs — so no all-gather follows.  Mirrors must have the layout of `out` (pitch, frame stride,
 * first row); n = 0 clears them.  Only for coverings (the streaming overlap-add), without saturation.
 * The caller orders the ranks (e.g. a barrier after the call, before anyone reads its frame). */
int rpsf_plan_set_output_mirrors(rpsf_plan* p, int n, void* const* mirrors);
/* cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle (lazy peer access) / cudaIpcCloseMemHandle;
 * free an rpsf_ipc_alloc buffer with rpsf_device_free once every peer has closed it */
int rpsf_ipc_alloc(void** ptr, int64_t bytes, int device, unsigned char handle[64]);
int rpsf_ipc_open(void** ptr, const unsigned char handle[64], int device);
int rpsf_ipc_close(void* ptr, int device);

/* Per-stage device timing for bench.py's roofline: when enabled, rpsf_apply records CUDA events
 * on the caller's stream around K1, K2 and the K3 colour phases.  rpsf_plan_read_timing
 * synchronises those events, adds the elapsed milliseconds of every apply call since the last
 * read into ms[0..2] (K1 incl. the zero-fill when colour 0 does not tile the band, K2, K3),
 * stores the call count and resets. */
int rpsf_plan_enable_timing(rpsf_plan* p, int enabled);
int rpsf_plan_read_timing(rpsf_plan* p, double ms[3], int* calls);
/* the np.pad index map the gather kernel uses (host evaluation): source index for padded
 * position i of an axis of length n, or -1 for the zero fill of RPSF_PAD_CONSTANT */
int rpsf_pad_index(int i, int n, int pad_mode);
/* for host-only bindings: cudaMalloc + synchronous copy of a host buffer (e.g.
 * the kernel cube for rpsf_transform_set_kernel), and the matching free (synchronises first) */
int rpsf_upload(void** device_ptr, const void* src_host, int64_t bytes, int device);
int rpsf_device_free(void* device_ptr, int device);
/* test hook: synchronous device -> host copy of raw bytes (cudaMemcpy) */
int rpsf_copy_to_host(void* dst_host, const void* src_device, int64_t bytes, int device);
/* number of kernel launches issued by the library since load (bench.py's gpu_launches) */
int64_t rpsf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RPSF_B200_H */
