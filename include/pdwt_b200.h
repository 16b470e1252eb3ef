/*
 * pdwt_b200.h -- C ABI of the B200-native PDWT hot path (libpdwt_b200.so).
 *
 * Plain C: pointers, ints, floats; no C++ or torch types cross this boundary.  Two layers:
 *
 *  Layer A  "drivers"  -- one entry point per transform driver of the reference, same argument meaning
 *           (device image, host array of device sub-band pointers, device scratch, w_info by value), plus the
 *           things the reference kept in process-global state: a per-instance filter handle instead of the
 *           `__constant__ c_kern_*` symbols (common.h:28-36), an explicit stream instead of the legacy default
 *           stream, and a `batch` count (independent planes).
 *  Layer B  "object"   -- the reference's `Wavelets` class (wt.h:20-76) as an opaque handle, method for method;
 *           this is what a ctypes / cgo / JNI binding (or pypwt's Cython) binds.  include/wt.h is the C++ face
 *           of the same object with the reference's public data members.
 *
 * Every function returns 0 on success or a negative pdwt_status unless stated otherwise; CUDA failures are
 * reported (the reference checks none).  Nothing in this library falls back to the CPU: without a
 * usable CUDA device every compute entry point returns PDWT_ERR_CUDA.
 */
#ifndef PDWT_B200_H
#define PDWT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDWT_MAX_FILTER_WIDTH 40 /* MAX_FILTER_WIDTH, common.h:15 */

typedef enum pdwt_status {
    PDWT_OK = 0,
    PDWT_ERR_ARG = -1,       /* bad argument (NULL, size, level) */
    PDWT_ERR_WAVELET = -2,   /* unknown wavelet name: separable.cu:42-45 returns -2 as well */
    PDWT_ERR_FILTER_LEN = -3,/* custom filter longer than PDWT_MAX_FILTER_WIDTH (wt.cu:562-565 returns -1) */
    PDWT_ERR_STATE = -4,     /* call not allowed in the current w_state (wt.cu:237,274,311,476) */
    PDWT_ERR_ALLOC = -5,
    PDWT_ERR_CUDA = -6       /* a CUDA runtime call or kernel failed; see pdwt_last_cuda_error() */
} pdwt_status;

/* struct w_info, utils.h:9-19 -- identical layout, passed by value exactly like the reference drivers do */
typedef struct pdwt_w_info {
    int ndims;   /* 1 (batched rows) or 2 */
    int Nr, Nc;  /* rows, columns of one plane */
    int nlevels;
    int do_swt;  /* undecimated transform */
    int hlen;    /* filter length */
} pdwt_w_info;

/* enum w_state, wt.h:8-17 -- same values */
typedef enum pdwt_w_state {
    PDWT_W_INIT, PDWT_W_FORWARD, PDWT_W_INVERSE, PDWT_W_THRESHOLD, PDWT_W_CREATION_ERROR,
    PDWT_W_FORWARD_ERROR, PDWT_W_INVERSE_ERROR, PDWT_W_THRESHOLD_ERROR
} pdwt_w_state;

const char* pdwt_version(void);
int pdwt_last_cuda_error(void);              /* last cudaError_t seen by this library on the calling thread */
const char* pdwt_last_cuda_error_string(void);
int pdwt_device_count(void);                 /* 0 without a driver / device (never an error) */

/* ---------------------------------------------------------------------------------------------------------
 * Filters.  Replaces w_compute_filters_separable (separable.cu:19-54), w_compute_filters
 * (nonseparable.cu:32-83) and the cudaMemcpyToSymbol uploads: taps live in the handle and travel to the
 * kernels as launch parameters, so instances never share or overwrite each other's filters.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pdwt_filters pdwt_filters;

/* Returns hlen (>0) like the reference does -- 2 for "haar"/"db1"/"bior1.1"/"rbior1.1" when !do_swt
 * (separable.cu:24-28) -- or PDWT_ERR_WAVELET.  No device work. */
int pdwt_filters_create(pdwt_filters** out, const char* wname, int do_swt);
/* Custom 1-D bank (Wavelets::set_filters_forward/inverse, wt.cu:560-602; separable.cu:57-73). */
int pdwt_filters_create_custom(pdwt_filters** out, int hlen, const float* dec_lo, const float* dec_hi,
                               const float* rec_lo, const float* rec_hi);
void pdwt_filters_destroy(pdwt_filters* f);
/* Custom 2-D quadruple of the non-separable drivers (w_set_filters_forward_nonseparable / _inverse_nonseparable,
 * nonseparable.cu:86-106): four hlen x hlen filters, row-major, the reference's array layout; direction +1 forward,
 * -1 inverse.  Each direction keeps its own quadruple (the reference shares ONE set of __constant__ symbols between
 * both, so its last upload wins).  Without one the drivers use the outer products of the 1-D banks. */
int pdwt_filters_set_2d(pdwt_filters* f, int direction, const float* ll, const float* lh, const float* hl, const float* hh);
int pdwt_filters_has_2d(const pdwt_filters* f, int direction);
int pdwt_filters_hlen(const pdwt_filters* f);
/* copies the four 1-D banks (hlen floats each) to host arrays; any pointer may be NULL */
int pdwt_filters_get(const pdwt_filters* f, float* dec_lo, float* dec_hi, float* rec_lo, float* rec_hi);
/* number of entries and name of entry i of the built-in table (filters.cpp:5919-6002) */
int pdwt_wavelet_count(void);
const char* pdwt_wavelet_name(int i);

/* ---------------------------------------------------------------------------------------------------------
 * Layer A: transform drivers.  Signature of the reference: int f(DTYPE* d_image, DTYPE** d_coeffs,
 * DTYPE* d_tmp, w_info winfos) (separable.h:12-28, haar.h:9-16, nonseparable.h:15-21).
 *   d_image  : device, batch planes of Nr*Nc floats, contiguous.
 *   d_coeffs : HOST array of device pointers [A, H1,V1,D1, ...] (2-D) / [A, D1, ...] (1-D), common.cu:399,429;
 *              each sub-band holds `batch` contiguous planes of its own size; d_coeffs[0] planes have the
 *              reference's allocation size (level-1 size for the DWT, full size for the SWT, common.cu:402-422).
 *   d_tmp    : device scratch, batch * 2*Nr*Nc floats (wt.cu:128-130).
 *   stream   : a cudaStream_t (NULL = legacy default stream, the reference's behaviour).
* Forward leaves A_L in d_coeffs[0] (no D2D fix-up copies); inverse overwrites d_image and d_tmp like the reference
 * (wt.cu:273-307) and MAY overwrite d_coeffs[0] (the generic two-pass kernels do, the fused families keep it).  Results: see DESIGN.md "Arithmetic contract".
 * --------------------------------------------------------------------------------------------------------- */
#define PDWT_DRIVER_ARGS const pdwt_filters *f, float *d_image, float **d_coeffs, float *d_tmp, pdwt_w_info winfos, \
                         int batch, void *stream
int pdwt_forward_separable(PDWT_DRIVER_ARGS);         /* w_forward_separable          separable.cu:179 */
int pdwt_forward_separable_1d(PDWT_DRIVER_ARGS);      /* w_forward_separable_1d       separable.cu:214 */
int pdwt_inverse_separable(PDWT_DRIVER_ARGS);         /* w_inverse_separable          separable.cu:332 */
int pdwt_inverse_separable_1d(PDWT_DRIVER_ARGS);      /* w_inverse_separable_1d       separable.cu:368 */
int pdwt_forward_swt_separable(PDWT_DRIVER_ARGS);     /* w_forward_swt_separable      separable.cu:496 */
int pdwt_forward_swt_separable_1d(PDWT_DRIVER_ARGS);  /* w_forward_swt_separable_1d   separable.cu:519 */
int pdwt_inverse_swt_separable(PDWT_DRIVER_ARGS);     /* w_inverse_swt_separable      separable.cu:629 */
int pdwt_inverse_swt_separable_1d(PDWT_DRIVER_ARGS);  /* w_inverse_swt_separable_1d   separable.cu:653 */
int pdwt_haar_forward2d(PDWT_DRIVER_ARGS);            /* haar_forward2d               haar.cu:61  */
int pdwt_haar_inverse2d(PDWT_DRIVER_ARGS);            /* haar_inverse2d               haar.cu:87  */
int pdwt_haar_forward1d(PDWT_DRIVER_ARGS);            /* haar_forward1d               haar.cu:164 */
int pdwt_haar_inverse1d(PDWT_DRIVER_ARGS);            /* haar_inverse1d               haar.cu:193 */
int pdwt_forward_nonseparable(PDWT_DRIVER_ARGS);      /* w_forward                    nonseparable.cu:233 */
int pdwt_inverse_nonseparable(PDWT_DRIVER_ARGS);      /* w_inverse                    nonseparable.cu:261 */
int pdwt_forward_swt_nonseparable(PDWT_DRIVER_ARGS);  /* w_forward_swt                nonseparable.cu:408 */
int pdwt_inverse_swt_nonseparable(PDWT_DRIVER_ARGS);  /* w_inverse_swt                nonseparable.cu:430 */

/* Dispatch of Wavelets::forward / ::inverse (wt.cu:247-266, 283-303): picks one of the 16 drivers above from
 * (ndims, hlen==2 && !do_swt, do_swt, do_separable). */
int pdwt_forward(PDWT_DRIVER_ARGS, int do_separable);
int pdwt_inverse(PDWT_DRIVER_ARGS, int do_separable);

/* w_call_soft_thresh / w_call_hard_thresh, common.cu:219-282 (incl. per-level beta/sqrt(2) when normalize>0 and
 * the hard-threshold quirk of common.cu:270).  In place on the sub-bands. */
int pdwt_call_soft_thresh(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int normalize,
                          int batch, void* stream);
int pdwt_call_hard_thresh(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int normalize,
                          int batch, void* stream);
/* the other proximal operators and coefficient helpers of the class */
int pdwt_call_group_soft_thresh(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int normalize,
                                int batch, void* stream);                       /* w_call_group_soft_thresh common.cu:311 */
int pdwt_call_proj_linf(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int batch,
                        void* stream);                                          /* w_call_proj_linf         common.cu:285 */
int pdwt_shrink(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int batch,
                void* stream);                                                  /* w_shrink                 common.cu:343 */
int pdwt_add_coeffs(float** dst, float** src, pdwt_w_info winfos, float alpha, int batch,
                    void* stream);                                              /* w_add_coeffs(_1d)        common.cu:499 */
int pdwt_call_circshift(float* d_image, float* d_image2, pdwt_w_info winfos, int sr, int sc, int inplace, int batch,
                        void* stream);                                          /* w_call_circshift         common.cu:375 */

/* Wavelets::norm1 / ::norm2sq, wt.cu:398-418 / 370-395 (cuBLAS asum / nrm2 replaced by one warp-shuffle
 * reduction kernel over all sub-bands).  out: HOST array of `batch` floats, one norm per plane; synchronises
 * `stream`.  norm2sq returns the true sum of squares also in 1-D (the reference sums asum() there, wt.cu:389). */
int pdwt_norm1(float** d_coeffs, pdwt_w_info winfos, int batch, float* out, void* stream);
int pdwt_norm2sq(float** d_coeffs, pdwt_w_info winfos, int batch, float* out, void* stream);

/* size helpers: w_div2 (utils.cu:24-27), sub-band geometry of get_coeff (wt.cu:481-504), allocation sizes of
 * w_create_coeffs_buffer(_1d) (common.cu:400-445) */
int pdwt_div2(int n);
int pdwt_num_coeffs(pdwt_w_info winfos);                         /* 3*L+1 or L+1 */
int pdwt_coeff_dims(pdwt_w_info winfos, int num, int* nr, int* nc); /* logical size of sub-band `num` */
size_t pdwt_coeff_alloc_elems(pdwt_w_info winfos, int num);      /* floats per plane to allocate for `num` */
int pdwt_max_level(int Nr, int Nc, int ndims, int hlen);         /* w_ilog2(N/(hlen-1)), wt.cu:156-159 */

/* ---------------------------------------------------------------------------------------------------------
 * Layer B: the Wavelets object (wt.h:20-76), method for method.  `batch` (>=1) is the only extension: the
 * object then owns `batch` independent planes and every method acts on all of them.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pdwt_wavelets pdwt_wavelets;

/* Wavelets::Wavelets(img, Nr, Nc, wname, levels, memisonhost, do_separable, do_cycle_spinning, do_swt, ndim),
 * wt.cu:84-185.  img may be NULL (zero image).  Always returns an object unless allocation fails; creation
 * problems are reported the reference's way through state == PDWT_W_CREATION_ERROR (an unknown wavelet sets it
 * instead of hanging).  do_cycle_spinning: forward() applies a random circular shift (rand(), as the
 * reference, wt.cu:242-246) and inverse() undoes it. */
int pdwt_wavelets_create(pdwt_wavelets** out, const float* img, int Nr, int Nc, const char* wname, int levels,
                         int memisonhost, int do_separable, int do_cycle_spinning, int do_swt, int ndim, int batch);
int pdwt_wavelets_copy(pdwt_wavelets** out, const pdwt_wavelets* src); /* copy-ctor, wt.cu:191-222 */
void pdwt_wavelets_destroy(pdwt_wavelets* w);

int pdwt_wavelets_forward(pdwt_wavelets* w);                       /* wt.cu:236-271 */
int pdwt_wavelets_inverse(pdwt_wavelets* w);                       /* wt.cu:273-307 */
int pdwt_wavelets_soft_threshold(pdwt_wavelets* w, float beta, int do_thresh_appcoeffs, int normalize); /* :310 */
int pdwt_wavelets_hard_threshold(pdwt_wavelets* w, float beta, int do_thresh_appcoeffs, int normalize); /* :320 */
int pdwt_wavelets_group_soft_threshold(pdwt_wavelets* w, float beta, int do_thresh_appcoeffs, int normalize); /* wt.cu:330 */
int pdwt_wavelets_shrink(pdwt_wavelets* w, float beta, int do_thresh_appcoeffs);     /* wt.cu:341 */
int pdwt_wavelets_proj_linf(pdwt_wavelets* w, float beta, int do_thresh_appcoeffs);  /* wt.cu:350 */
int pdwt_wavelets_circshift(pdwt_wavelets* w, int sr, int sc, int inplace);          /* wt.cu:365 */
int pdwt_wavelets_add_wavelet(pdwt_wavelets* w, const pdwt_wavelets* other, float alpha); /* wt.cu:624; the reference's codes */
int pdwt_wavelets_current_shift(const pdwt_wavelets* w, int* sr, int* sc);            /* current_shift_r / _c, wt.h:27-28 */
int pdwt_wavelets_norm1(pdwt_wavelets* w, float* out);             /* wt.cu:398; out[batch] on the host */
int pdwt_wavelets_norm2sq(pdwt_wavelets* w, float* out);           /* wt.cu:370 */
int pdwt_wavelets_get_image(pdwt_wavelets* w, float* img);         /* wt.cu:421; returns the element count */
int pdwt_wavelets_get_coeff(pdwt_wavelets* w, float* coeff, int num); /* wt.cu:475; element count, 0 after inverse */
int pdwt_wavelets_set_image(pdwt_wavelets* w, const float* img, int mem_is_on_device);          /* wt.cu:427 */
int pdwt_wavelets_set_coeff(pdwt_wavelets* w, const float* coeff, int num, int mem_is_on_device); /* wt.cu:437 */
int pdwt_wavelets_set_filters_forward(pdwt_wavelets* w, const char* name, unsigned len, const float* lo,
                                      const float* hi);            /* wt.cu:560 (separable mode) */
int pdwt_wavelets_set_filters_inverse(pdwt_wavelets* w, const float* lo, const float* hi); /* wt.cu:585 */
/* the same in non-separable mode: four len x len filters (A, H, V, D order of the reference = LL, LH, HL, HH) */
int pdwt_wavelets_set_filters_forward_2d(pdwt_wavelets* w, const char* name, unsigned len, const float* f1, const float* f2,
                                         const float* f3, const float* f4);
int pdwt_wavelets_set_filters_inverse_2d(pdwt_wavelets* w, const float* f1, const float* f2, const float* f3,
                                         const float* f4);
int pdwt_wavelets_sync(pdwt_wavelets* w);                          /* cudaStreamSynchronize of the object's stream */
int pdwt_wavelets_set_stream(pdwt_wavelets* w, void* stream);
/* on != 0: get_image / set_image / get_coeff / set_coeff only enqueue their host<->device copies on the object's
 * stream (use pinned host memory and pdwt_wavelets_sync before touching it); several objects on different streams
 * then overlap H2D, kernels and D2H.  Default 0 = the reference's blocking copies (wt.cu:421-434). */
int pdwt_wavelets_set_async(pdwt_wavelets* w, int on);
/* norm1 / norm2sq right after a threshold return sums the threshold kernel produced on its way; call this after
 * writing coefficients through pdwt_wavelets_coeff_int_ptr() behind the object's back */
int pdwt_wavelets_invalidate_norm_cache(pdwt_wavelets* w);

/* public data members of the class, wt.h:24-33 */
int pdwt_wavelets_state(const pdwt_wavelets* w);
pdwt_w_info pdwt_wavelets_info(const pdwt_wavelets* w);
int pdwt_wavelets_batch(const pdwt_wavelets* w);
int pdwt_wavelets_do_separable(const pdwt_wavelets* w);
const char* pdwt_wavelets_wname(const pdwt_wavelets* w);
intptr_t pdwt_wavelets_image_int_ptr(const pdwt_wavelets* w);          /* wt.cu:660 */
intptr_t pdwt_wavelets_coeff_int_ptr(const pdwt_wavelets* w, int num); /* wt.cu:665 */
intptr_t pdwt_wavelets_tmp_int_ptr(const pdwt_wavelets* w);
/* number of kernels this object has launched so far (bench.py's gpu_launches) */
long long pdwt_wavelets_launch_count(const pdwt_wavelets* w);
long long pdwt_launch_count(void);     /* process-wide */


/* ---------------------------------------------------------------------------------------------------------
 * Layer C: a batch of independent planes spread over the GPUs of one node (the reference has no
 * multi-GPU support, TODO.txt:15).  One process (or thread) per GPU; rank r of G owns a contiguous block of planes
 * (pdwt_shard_block) as ONE batched pdwt_wavelets object, the transforms never communicate, and inputs / outputs
 * move device to device over NCCL (NVLink): grouped ncclSend / ncclRecv from / to the root's device buffer, one
 * ncclAllGather for per-plane norms.  NCCL is bound at run time (libnccl.so.2); without it these entry points
 * return PDWT_ERR_CUDA and pdwt_shard_last_error() says why.  All pointers are DEVICE pointers unless noted.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pdwt_shard pdwt_shard;
int pdwt_shard_nccl_version(void);                 /* 0: no usable NCCL */
const char* pdwt_shard_last_error(void);
int pdwt_shard_unique_id(unsigned char id[128]);   /* ncclGetUniqueId on one rank; hand the bytes to the others */
int pdwt_shard_create(pdwt_shard** out, const unsigned char id[128], int nranks, int rank); /* binds the current device */
void pdwt_shard_destroy(pdwt_shard* s);
int pdwt_shard_rank(const pdwt_shard* s);
int pdwt_shard_nranks(const pdwt_shard* s);
/* block of rank `rank`: contiguous, the first n % nranks ranks own one plane more */
void pdwt_shard_block(long long n, int nranks, int rank, long long* first, long long* count);
/* root's n planes of `plane` floats -> every rank's block, and back; enqueued on `stream` */
int pdwt_shard_scatter(pdwt_shard* s, const float* d_full, float* d_mine, long long n, size_t plane, int root, void* stream);
int pdwt_shard_gather(pdwt_shard* s, const float* d_mine, float* d_full, long long n, size_t plane, int root, void* stream);
/* the same with this rank's block held by a batched object (batch == its plane count), on the object's stream:
 * scatter into d_image (state becomes W_INIT like set_image), gather of d_image or of sub-band `num` (`plane` = the
 * logical size from pdwt_coeff_dims; refused after inverse() like get_coeff, wt.cu:476-479) */
int pdwt_shard_scatter_image(pdwt_shard* s, pdwt_wavelets* w, const float* d_full, long long n, size_t plane, int root);
int pdwt_shard_gather_image(pdwt_shard* s, pdwt_wavelets* w, float* d_full, long long n, size_t plane, int root);
int pdwt_shard_gather_coeff(pdwt_shard* s, pdwt_wavelets* w, int num, float* d_full, long long n, size_t plane, int root);
/* per-plane norms of the WHOLE batch on every rank: h_all = HOST array of n floats; which = 1 norm1, 2 norm2sq */
int pdwt_shard_norms(pdwt_shard* s, pdwt_wavelets* w, int which, float* h_all, long long n);

/* ---------------------------------------------------------------------------------------------------------
 * Per-kernel timing (measurement aid, no counterpart in the reference).  Between _begin and _end every kernel
 * launch of this library is bracketed by a CUDA-event pair on the stream it is launched on; _end synchronises,
 * aggregates by kernel tag (first-seen order) and returns the number of distinct tags (entries beyond
 * max_entries are dropped).  Process-wide; not meant to be left on in production.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct pdwt_profile_entry {
    char name[64];
    int launches;
    double ms_total, ms_min, ms_max;
} pdwt_profile_entry;
int pdwt_profile_begin(void);
int pdwt_profile_end(pdwt_profile_entry* out, int max_entries);

#ifdef __cplusplus
}
#endif
#endif /* PDWT_B200_H */
