// pdwt_wavelets.cu -- the `Wavelets` class of include/wt.h (drop-in for reference src/wt.h / wt.cu) and Layer B
// of the C ABI (the same object behind an opaque handle).  Pure host logic: state machine, buffer ownership,
// level clamping, get/set; every transform goes through the Layer A drivers of pdwt_capi.cu.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include <new>

#include "pdwt_common.cuh"
#include "pdwt_object.h"

namespace pdwt {
int norm_impl(float** c, pdwt_w_info w, int batch, int mode, float* out, cudaStream_t s, double* d_sums, double* h_sums);
int norm_finish(pdwt_w_info w, int batch, int mode, float* out, cudaStream_t s, const double* d_sums, double* h_sums);
int threshold_norms(float** c, float beta, pdwt_w_info w, int app, int normalize, int batch, cudaStream_t s, int hard,
                    double* d_sums2, const HostPublish* hp);
void norm_accumulate(pdwt_w_info w, int batch, int mode, float* out, const double* h_sums);
unsigned next_publish_tag();
int wait_published(const unsigned long long* h, int nsums, unsigned tag, double* vals, cudaStream_t s);
int norm_published(float** c, pdwt_w_info w, int batch, int mode, float* out, cudaStream_t s, double* d_sums,
                   const HostPublish& hp, const unsigned long long* h_words, double* vals);
}
using namespace pdwt;

#define W_TRY(expr, errstate)             \
    do {                                  \
        int rc__ = (expr);                \
        if (rc__ < 0) {                   \
            last_error = rc__;            \
            state = (errstate);           \
            return;                       \
        }                                 \
    } while (0)

static int cuda_rc(cudaError_t e) { return note_cuda(e); }
// PDWT_NORM_CACHE=0: norm1()/norm2sq() always re-read the coefficients (for callers that write into the sub-bands through
// coeff_int_ptr() between a threshold and a norm, which no method of the class can see)
static bool norm_cache_enabled()
{
    const char* e = getenv("PDWT_NORM_CACHE");
    return !e || atoi(e) != 0;
}

Wavelets::Wavelets()
    : d_image(NULL), d_coeffs(NULL), d_tmp(NULL), current_shift_r(0), current_shift_c(0), do_separable(1),
      do_cycle_spinning(0), state(W_INIT), batch(1), last_error(0), stream(NULL), async_copies(0), filters(NULL),
      d_sums(NULL), h_sums(NULL), launches(0), norm_cache(0), h_sums_dev(NULL), thr_tag(0)
{
    memset(wname, 0, sizeof wname);
    memset(&winfos, 0, sizeof winfos);
}

int Wavelets::alloc_buffers()
{
    const size_t plane = (size_t)winfos.Nr * winfos.Nc;
    int rc;
    if ((rc = cuda_rc(cudaMalloc(&d_image, sizeof(DTYPE) * plane * batch))) < 0) return rc;
    if ((rc = cuda_rc(cudaMalloc(&d_tmp, sizeof(DTYPE) * 2 * plane * batch))) < 0) return rc;  // wt.cu:128-130
    // Norm scratch (HostPublish, pdwt_common.cuh).  With cap = kMaxSeg * batch:
    //   device (doubles): [0, cap) reduction sums | [cap, 3 cap) L1 and L2 sums left by a threshold | two tickets; all
    //           zero between launches
    //   host (pinned, mapped; 8-byte words): [0, 2 cap) published reduction sums | [2 cap, 6 cap) published threshold
    //           sums | [6 cap, 8 cap) the decoded doubles (host only)
    const size_t cap = (size_t)kMaxSeg * batch;
    if ((rc = cuda_rc(cudaMalloc(&d_sums, sizeof(double) * (3 * cap + 2)))) < 0) return rc;
    if ((rc = cuda_rc(cudaMemset(d_sums, 0, sizeof(double) * (3 * cap + 2)))) < 0) return rc;
    if ((rc = cuda_rc(cudaDeviceSynchronize())) < 0) return rc;   // once per object: the zeros precede any stream's work
    if ((rc = cuda_rc(cudaHostAlloc(&h_sums, sizeof(double) * 8 * cap, cudaHostAllocMapped))) < 0) return rc;
    memset(h_sums, 0, sizeof(double) * 8 * cap);
    if ((rc = cuda_rc(cudaHostGetDevicePointer(&h_sums_dev, h_sums, 0))) < 0) return rc;
    return PDWT_OK;
}

void Wavelets::free_buffers()
{
    if (d_image) cudaFree(d_image);
    if (d_tmp) cudaFree(d_tmp);
    if (d_sums) cudaFree(d_sums);
    if (h_sums) cudaFreeHost(h_sums);
    if (d_coeffs) {
        // one allocation backs every sub-band (see the constructor); d_coeffs[1] is its base when there are details
        const int n = pdwt_num_coeffs(winfos);
        if (n > 1 && d_coeffs[1]) cudaFree(d_coeffs[1]);
        if (d_coeffs[0]) cudaFree(d_coeffs[0]);
        free(d_coeffs);
    }
    if (filters) pdwt_filters_destroy(filters);
    d_image = d_tmp = NULL;
    d_coeffs = NULL;
    d_sums = h_sums = h_sums_dev = NULL;
    filters = NULL;
}

// Constructor from an image, reference wt.cu:84-185.
Wavelets::Wavelets(DTYPE* img, int Nr, int Nc, const char* name, int levels, int memisonhost, int do_separable_,
                   int do_cycle_spinning_, int do_swt, int ndim, int batch_)
    : d_image(NULL), d_coeffs(NULL), d_tmp(NULL), current_shift_r(0), current_shift_c(0), do_separable(do_separable_),
      do_cycle_spinning(do_cycle_spinning_), state(W_INIT), batch(batch_ < 1 ? 1 : batch_), last_error(0), stream(NULL),
      async_copies(0), filters(NULL), d_sums(NULL), h_sums(NULL), launches(0), norm_cache(0), h_sums_dev(NULL), thr_tag(0)
{
    memset(wname, 0, sizeof wname);
    winfos.Nr = Nr;
    winfos.Nc = Nc;
    winfos.nlevels = levels;
    winfos.do_swt = do_swt;
    winfos.ndims = ndim;
    winfos.hlen = 0;
    if (levels < 1) {  // wt.cu:111-114
        puts("Warning: cannot initialize wavelet coefficients with nlevels < 1. Forcing nlevels = 1");
        winfos.nlevels = 1;
    }
    if (Nr < 1 || Nc < 1 || !name) {
        last_error = PDWT_ERR_ARG;
        state = W_CREATION_ERROR;
        winfos.Nr = winfos.Nc = 0;
        return;
    }
    if (Nr == 1) {  // 1-D data, wt.cu:133-136
        ndim = 1;
        winfos.ndims = 1;
    }
    if (ndim == 1 && do_separable == 0) {  // wt.cu:138-142 (the member is updated too)
        puts("Warning: 1D DWT was requestred, which is incompatible with non-separable transform.");
        puts("Ignoring the do_separable option.");
        do_separable = 1;
    }
    strncpy(wname, name, sizeof(wname) - 1);

    W_TRY(alloc_buffers(), W_CREATION_ERROR);
    const size_t bytes = sizeof(DTYPE) * (size_t)Nr * Nc * batch;
    if (!img)
        W_TRY(cuda_rc(cudaMemset(d_image, 0, bytes)), W_CREATION_ERROR);
    else
        W_TRY(cuda_rc(cudaMemcpy(d_image, img, bytes, memisonhost ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice)),
              W_CREATION_ERROR);

    // Filters (wt.cu:144-153).  Unknown names set the error state instead of looping forever.
    int hlen = pdwt_filters_create(&filters, wname, do_swt);
    if (hlen <= 0) {
        printf("ERROR: unknown wavelet name %s\n", wname);
        last_error = hlen;
        state = W_CREATION_ERROR;
        winfos.nlevels = 0;  // no sub-bands are allocated
        return;
    }
    winfos.hlen = hlen;
    if (ndim != 1 && ndim != 2) {
        printf("ERROR: ndim=%d is not implemented\n", ndim);
        last_error = PDWT_ERR_ARG;
        state = W_CREATION_ERROR;
        winfos.nlevels = 0;
        return;
    }
    // Level clamp, wt.cu:155-165
    const int wmaxlev = pdwt_max_level(Nr, Nc, ndim, hlen);
    if (winfos.nlevels > wmaxlev) {
        printf("Warning: required level (%d) is greater than the maximum possible level for %s (%d) on a %dx%d image.\n",
               winfos.nlevels, wname, wmaxlev, winfos.Nc, winfos.Nr);
        printf("Forcing nlevels = %d\n", wmaxlev);
        winfos.nlevels = wmaxlev;
    }
    if (winfos.nlevels < 1) {  // image smaller than the filter: nothing can be computed
        last_error = PDWT_ERR_ARG;
        state = W_CREATION_ERROR;
        winfos.nlevels = 0;
        return;
    }
    // Sub-band buffers (w_create_coeffs_buffer(_1d), common.cu:400-445): same pointer table, same sizes, zeroed;
    // the detail sub-bands share ONE allocation (each plane group 256-byte aligned) instead of 3L cudaMallocs.
    const int n = pdwt_num_coeffs(winfos);
    d_coeffs = (DTYPE**)calloc(n, sizeof(DTYPE*));
    if (!d_coeffs) {
        last_error = PDWT_ERR_ALLOC;
        state = W_CREATION_ERROR;
        return;
    }
    size_t total = 0;
    for (int k = 1; k < n; k++) total += (pdwt_coeff_alloc_elems(winfos, k) * batch + 63) / 64 * 64;
    DTYPE* pool = NULL;
    if (total) {
        W_TRY(cuda_rc(cudaMalloc(&pool, sizeof(DTYPE) * total)), W_CREATION_ERROR);
        W_TRY(cuda_rc(cudaMemset(pool, 0, sizeof(DTYPE) * total)), W_CREATION_ERROR);
    }
    size_t off = 0;
    for (int k = 1; k < n; k++) {
        d_coeffs[k] = pool + off;
        off += (pdwt_coeff_alloc_elems(winfos, k) * batch + 63) / 64 * 64;
    }
    const size_t a_elems = pdwt_coeff_alloc_elems(winfos, 0) * batch;
    W_TRY(cuda_rc(cudaMalloc(&d_coeffs[0], sizeof(DTYPE) * a_elems)), W_CREATION_ERROR);
    W_TRY(cuda_rc(cudaMemset(d_coeffs[0], 0, sizeof(DTYPE) * a_elems)), W_CREATION_ERROR);
    if (do_cycle_spinning && do_swt)   // wt.cu:173
        puts("Warning: makes little sense to use Cycle spinning with stationary Wavelet transform");
    if (do_cycle_spinning && ndim == 1) {   // wt.cu:175-179
        puts("ERROR: cycle spinning is not implemented for 1D. Use SWT instead.");
        last_error = PDWT_ERR_ARG;
        state = W_CREATION_ERROR;
    }
    // The fills and copies above ran on the legacy default stream, which does NOT order against the non-blocking
    // stream the object may be given next (set_stream): they must have landed before the constructor returns.
    W_TRY(cuda_rc(cudaDeviceSynchronize()), W_CREATION_ERROR);
}

// Copy constructor (deep), reference wt.cu:191-222
Wavelets::Wavelets(const Wavelets& W)
    : d_image(NULL), d_coeffs(NULL), d_tmp(NULL), current_shift_r(W.current_shift_r), current_shift_c(W.current_shift_c),
      do_separable(W.do_separable), do_cycle_spinning(W.do_cycle_spinning), winfos(W.winfos), state(W.state),
      batch(W.batch), last_error(0), stream(W.stream), async_copies(W.async_copies), filters(NULL), d_sums(NULL), h_sums(NULL), launches(0), norm_cache(0), h_sums_dev(NULL), thr_tag(0)
{
    memcpy(wname, W.wname, sizeof wname);
    if (winfos.Nr < 1 || winfos.Nc < 1) return;
    // the copies below run on the legacy default stream: whatever the source's own (possibly non-blocking) stream still
    // has in flight must be complete first
    W_TRY(cuda_rc(cudaStreamSynchronize((cudaStream_t)W.stream)), W_CREATION_ERROR);
    W_TRY(alloc_buffers(), W_CREATION_ERROR);
    const size_t plane = (size_t)winfos.Nr * winfos.Nc;
    W_TRY(cuda_rc(cudaMemcpy(d_image, W.d_image, sizeof(DTYPE) * plane * batch, cudaMemcpyDeviceToDevice)),
          W_CREATION_ERROR);
    if (W.filters) {
        float L[PDWT_MAX_FILTER_WIDTH], H[PDWT_MAX_FILTER_WIDTH], IL[PDWT_MAX_FILTER_WIDTH], IH[PDWT_MAX_FILTER_WIDTH];
        const int hlen = pdwt_filters_get(W.filters, L, H, IL, IH);
        pdwt_filters_create_custom(&filters, hlen, L, H, IL, IH);
    }
    if (!W.d_coeffs || winfos.nlevels < 1) return;
    const int n = pdwt_num_coeffs(winfos);
    d_coeffs = (DTYPE**)calloc(n, sizeof(DTYPE*));
    size_t total = 0;
    for (int k = 1; k < n; k++) total += (pdwt_coeff_alloc_elems(winfos, k) * batch + 63) / 64 * 64;
    DTYPE* pool = NULL;
    if (total) W_TRY(cuda_rc(cudaMalloc(&pool, sizeof(DTYPE) * total)), W_CREATION_ERROR);
    size_t off = 0;
    for (int k = 1; k < n; k++) {
        d_coeffs[k] = pool + off;
        const size_t e = pdwt_coeff_alloc_elems(winfos, k) * batch;
        W_TRY(cuda_rc(cudaMemcpy(d_coeffs[k], W.d_coeffs[k], sizeof(DTYPE) * e, cudaMemcpyDeviceToDevice)),
              W_CREATION_ERROR);
        off += (e + 63) / 64 * 64;
    }
    const size_t a_elems = pdwt_coeff_alloc_elems(winfos, 0) * batch;
    W_TRY(cuda_rc(cudaMalloc(&d_coeffs[0], sizeof(DTYPE) * a_elems)), W_CREATION_ERROR);
    W_TRY(cuda_rc(cudaMemcpy(d_coeffs[0], W.d_coeffs[0], sizeof(DTYPE) * a_elems, cudaMemcpyDeviceToDevice)),
          W_CREATION_ERROR);
    W_TRY(cuda_rc(cudaDeviceSynchronize()), W_CREATION_ERROR);   // as in the image constructor
}

Wavelets::~Wavelets() { free_buffers(); }

// reference wt.cu:236-271
void Wavelets::forward()
{
    if (state == W_CREATION_ERROR) {
        puts("Warning: forward transform not computed, as there was an error when creating the wavelets");
        return;
    }
    norm_cache = 0;
    const long long before = pdwt_launch_count();
    if (do_cycle_spinning) {   // wt.cu:242-246: a random circular shift of the image before the transform
        current_shift_r = rand() % winfos.Nr;
        current_shift_c = rand() % winfos.Nc;
        W_TRY(pdwt_call_circshift(d_image, d_tmp, winfos, current_shift_r, current_shift_c, 1, batch, stream),
              W_FORWARD_ERROR);
    }
    W_TRY(pdwt_forward(filters, d_image, d_coeffs, d_tmp, winfos, batch, stream, do_separable), W_FORWARD_ERROR);
    launches += pdwt_launch_count() - before;
    state = W_FORWARD;
}

// reference wt.cu:273-307
void Wavelets::inverse()
{
    if (state == W_INVERSE) {
        puts("Warning: W.inverse() has already been run. Inverse is available in W.get_image()");
        return;
    }
    if (state == W_FORWARD_ERROR || state == W_THRESHOLD_ERROR || state == W_CREATION_ERROR) {
        puts("Warning: inverse transform not computed, as there was an error in a previous stage");
        return;
    }
    norm_cache = 0;
    const long long before = pdwt_launch_count();
    W_TRY(pdwt_inverse(filters, d_image, d_coeffs, d_tmp, winfos, batch, stream, do_separable), W_INVERSE_ERROR);
    if (do_cycle_spinning)     // wt.cu:305: shift back
        W_TRY(pdwt_call_circshift(d_image, d_tmp, winfos, -current_shift_r, -current_shift_c, 1, batch, stream),
              W_INVERSE_ERROR);
    launches += pdwt_launch_count() - before;
    state = W_INVERSE;
}

// reference wt.cu:310-317
void Wavelets::soft_threshold(DTYPE beta, int do_thresh_appcoeffs, int normalize)
{
    if (state == W_INVERSE) {
        puts("Warning: Wavelets(): cannot threshold coefficients, as they were modified by W.inverse()");
        return;
    }
    if (state == W_CREATION_ERROR) return;
    const long long before = pdwt_launch_count();
    // the same launch leaves the L1 / L2 norms of the thresholded coefficients in d_sums (second and third block): a
    // norm1() / norm2sq() that follows needs no pass over the coefficients
    norm_cache = 0;
    HostPublish hp;
    publish_target(1, &hp);
    W_TRY(threshold_norms(d_coeffs, beta, winfos, do_thresh_appcoeffs, normalize, batch, (cudaStream_t)stream, 0,
                          d_sums + (size_t)kMaxSeg * batch, &hp),
          W_THRESHOLD_ERROR);
    thr_tag = hp.tag;
    norm_cache = (pdwt_num_coeffs(winfos) <= kMaxSeg && norm_cache_enabled()) ? 3 : 0;
    launches += pdwt_launch_count() - before;
}

// reference wt.cu:320-327
void Wavelets::hard_threshold(DTYPE beta, int do_thresh_appcoeffs, int normalize)
{
    if (state == W_INVERSE) {
        puts("Warning: Wavelets(): cannot threshold coefficients, as they were modified by W.inverse()");
        return;
    }
    if (state == W_CREATION_ERROR) return;
    const long long before = pdwt_launch_count();
    norm_cache = 0;
    HostPublish hp;
    publish_target(1, &hp);
    W_TRY(threshold_norms(d_coeffs, beta, winfos, do_thresh_appcoeffs, normalize, batch, (cudaStream_t)stream, 1,
                          d_sums + (size_t)kMaxSeg * batch, &hp),
          W_THRESHOLD_ERROR);
    thr_tag = hp.tag;
    norm_cache = (pdwt_num_coeffs(winfos) <= kMaxSeg && norm_cache_enabled()) ? 3 : 0;
    launches += pdwt_launch_count() - before;
}

// reference wt.cu:330-338
void Wavelets::group_soft_threshold(DTYPE beta, int do_thresh_appcoeffs, int normalize)
{
    if (state == W_INVERSE) {
        puts("Warning: Wavelets(): cannot threshold coefficients, as they were modified by W.inverse()");
        return;
    }
    if (state == W_CREATION_ERROR) return;
    norm_cache = 0;
    const long long before = pdwt_launch_count();
    W_TRY(pdwt_call_group_soft_thresh(d_coeffs, beta, winfos, do_thresh_appcoeffs, normalize, batch, stream),
          W_THRESHOLD_ERROR);
    launches += pdwt_launch_count() - before;
}

// reference wt.cu:341-348 (L2 proximal: every coefficient times 1/(1+beta))
void Wavelets::shrink(DTYPE beta, int do_thresh_appcoeffs)
{
    if (state == W_INVERSE) {
        puts("Warning: Wavelets(): cannot threshold coefficients, as they were modified by W.inverse()");
        return;
    }
    if (state == W_CREATION_ERROR) return;
    norm_cache = 0;
    const long long before = pdwt_launch_count();
    W_TRY(pdwt_shrink(d_coeffs, beta, winfos, do_thresh_appcoeffs, batch, stream), W_THRESHOLD_ERROR);
    launches += pdwt_launch_count() - before;
}

// reference wt.cu:350-358 (projection onto the L-infinity ball)
void Wavelets::proj_linf(DTYPE beta, int do_thresh_appcoeffs)
{
    if (state == W_INVERSE) {
        puts("Warning: Wavelets(): cannot threshold coefficients, as they were modified by W.inverse()");
        return;
    }
    if (state == W_CREATION_ERROR) return;
    norm_cache = 0;
    const long long before = pdwt_launch_count();
    W_TRY(pdwt_call_proj_linf(d_coeffs, beta, winfos, do_thresh_appcoeffs, batch, stream), W_THRESHOLD_ERROR);
    launches += pdwt_launch_count() - before;
}

// reference wt.cu:365-367.  inplace = 1: result in d_image; otherwise in d_tmp.
void Wavelets::circshift(int sr, int sc, int inplace)
{
    if (state == W_CREATION_ERROR || !d_image) return;
    const long long before = pdwt_launch_count();
    const int rc = pdwt_call_circshift(d_image, d_tmp, winfos, sr, sc, inplace, batch, stream);
    if (rc < 0) last_error = rc;
    launches += pdwt_launch_count() - before;
}

// reference wt.cu:624-657: this += alpha * W (coefficients only); same checks and return codes
int Wavelets::add_wavelet(Wavelets W, DTYPE alpha) { return add_wavelet_ref(W, alpha); }
int Wavelets::add_wavelet_ref(const Wavelets& W, DTYPE alpha)
{
    if ((winfos.nlevels != W.winfos.nlevels) || (strcasecmp(wname, W.wname))) {
        puts("ERROR: add_wavelet(): right operand is not the same transform (wname, level)");
        return -1;
    }
    if (state == W_INVERSE || W.state == W_INVERSE) {
        puts("WARNING: add_wavelet(): this operation makes no sense when wavelet has just been inverted");
        return 1;
    }
    if (winfos.Nr != W.winfos.Nr || winfos.Nc != W.winfos.Nc || winfos.ndims != W.winfos.ndims || batch != W.batch) {
        puts("ERROR: add_wavelet(): operands do not have the same geometry");
        return -2;
    }
    if ((winfos.do_swt != 0) ^ (W.winfos.do_swt != 0)) {
        puts("ERROR: add_wavelet(): operands should both use SWT or DWT");
        return -3;
    }
    if ((do_cycle_spinning != 0 && W.do_cycle_spinning != 0) &&
        ((current_shift_r != W.current_shift_r) || (current_shift_c != W.current_shift_c))) {
        puts("ERROR: add_wavelet(): operands do not have the same current shift");
        return -4;
    }
    if (!d_coeffs || !W.d_coeffs) return -2;
    if (W.stream != stream) cudaStreamSynchronize((cudaStream_t)W.stream);   // the operand's pending work
    const long long before = pdwt_launch_count();
    norm_cache = 0;
    const int rc = pdwt_add_coeffs(d_coeffs, W.d_coeffs, winfos, alpha, batch, stream);
    launches += pdwt_launch_count() - before;
    if (rc < 0) {
        last_error = rc;
        return rc;
    }
    return 0;
}

// where launch `which` (0: a reduction, 1: a threshold that leaves its norms) delivers its sums, see alloc_buffers()
void Wavelets::publish_target(int which, HostPublish* out) const
{
    const size_t cap = (size_t)kMaxSeg * batch;
    const int nseg = pdwt_num_coeffs(winfos);
    HostPublish& hp = *out;
    hp.h_out = reinterpret_cast<unsigned long long*>(h_sums_dev) + (which ? 2 * cap : 0);
    hp.ticket = reinterpret_cast<unsigned*>(d_sums + 3 * cap) + which;
    hp.tag = next_publish_tag();
    hp.nsums = (which ? 2 : 1) * batch * nseg;
}

int Wavelets::norms(int mode, DTYPE* out)
{
    if (state == W_CREATION_ERROR || !d_coeffs) return PDWT_ERR_STATE;
    const long long before = pdwt_launch_count();
    int rc;
    const size_t cap = (size_t)kMaxSeg * batch;
    const unsigned long long* words = reinterpret_cast<const unsigned long long*>(h_sums);
    double* vals = h_sums + 6 * cap;
    const int n1 = batch * pdwt_num_coeffs(winfos);
    if (norm_cache & (1 << mode)) {   // published by the last threshold: L1 sums, then L2 sums (batch * ncoeffs each)
        rc = wait_published(words + 2 * cap, 2 * n1, thr_tag, vals, (cudaStream_t)stream);
        if (rc == PDWT_OK) norm_accumulate(winfos, batch, mode, out, vals + (size_t)mode * n1);
    } else {
        HostPublish hp;
        publish_target(0, &hp);
        rc = norm_published(d_coeffs, winfos, batch, mode, out, (cudaStream_t)stream, d_sums, hp, words, vals);
    }
    launches += pdwt_launch_count() - before;
    if (rc < 0) last_error = rc;
    return rc;
}
int Wavelets::norm1_batched(DTYPE* out) { return norms(0, out); }
int Wavelets::norm2sq_batched(DTYPE* out) { return norms(1, out); }

// reference wt.cu:398-418 / 370-395; with batch > 1 the scalar form returns the sum over planes
DTYPE Wavelets::norm1()
{
    DTYPE* v = (DTYPE*)malloc(sizeof(DTYPE) * batch);
    DTYPE res = 0.0f;
    if (v && norms(0, v) == PDWT_OK)
        for (int p = 0; p < batch; p++) res += v[p];
    free(v);
    return res;
}
DTYPE Wavelets::norm2sq()
{
    DTYPE* v = (DTYPE*)malloc(sizeof(DTYPE) * batch);
    DTYPE res = 0.0f;
    if (v && norms(1, v) == PDWT_OK)
        for (int p = 0; p < batch; p++) res += v[p];
    free(v);
    return res;
}

// reference wt.cu:421-424 (blocking D2H; the copy is ordered after the object's stream).  With async_copies set the
// call only enqueues the copy: `img` (pinned) is valid after the caller synchronises the stream.
int Wavelets::get_image(DTYPE* img)
{
    if (!d_image || !img) return 0;
    const size_t n = (size_t)winfos.Nr * winfos.Nc * batch;
    int rc = cuda_rc(cudaMemcpyAsync(img, d_image, sizeof(DTYPE) * n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (rc == PDWT_OK && !async_copies) rc = cuda_rc(cudaStreamSynchronize((cudaStream_t)stream));
    if (rc < 0) {
        last_error = rc;
        return 0;
    }
    return n > 0x7fffffffull ? 0x7fffffff : (int)n;   // the reference's int count, saturated
}

// reference wt.cu:427-434
void Wavelets::set_image(DTYPE* img, int mem_is_on_device)
{
    if (!d_image || !img) return;
    const size_t n = (size_t)winfos.Nr * winfos.Nc * batch;
    int rc = cuda_rc(cudaMemcpyAsync(d_image, img, sizeof(DTYPE) * n,
                                     mem_is_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                     (cudaStream_t)stream));
    if (rc == PDWT_OK && !mem_is_on_device && !async_copies) rc = cuda_rc(cudaStreamSynchronize((cudaStream_t)stream));
    if (rc < 0) last_error = rc;
    if (state != W_CREATION_ERROR) state = W_INIT;
}

// Sub-band `num` holds `batch` planes.  For num == 0 the planes sit `alloc` floats apart but only the logical A_L part
// of each is exchanged (wt.cu:437-508).
static int copy_coeff(Wavelets* W, DTYPE* host_or_dev, int num, cudaMemcpyKind kind, bool to_device)
{
    int nr, nc;
    if (!W->d_coeffs || pdwt_coeff_dims(W->winfos, num, &nr, &nc) != PDWT_OK) return 0;
    const size_t n = (size_t)nr * nc, stride = pdwt_coeff_alloc_elems(W->winfos, num);
    cudaStream_t s = (cudaStream_t)W->stream;
    cudaError_t e;
    if (to_device)
        e = cudaMemcpy2DAsync(W->d_coeffs[num], stride * sizeof(DTYPE), host_or_dev, n * sizeof(DTYPE), n * sizeof(DTYPE),
                              W->batch, kind, s);
    else
        e = cudaMemcpy2DAsync(host_or_dev, n * sizeof(DTYPE), W->d_coeffs[num], stride * sizeof(DTYPE), n * sizeof(DTYPE),
                              W->batch, kind, s);
    if (e == cudaSuccess && !W->async_copies) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        W->last_error = note_cuda(e);
        return 0;
    }
    return n * W->batch > 0x7fffffffull ? 0x7fffffff : (int)(n * W->batch);
}

// reference wt.cu:475-508
int Wavelets::get_coeff(DTYPE* coeff, int num)
{
    if (state == W_INVERSE) {
        puts("Warning: get_coeff(): inverse() has been performed, the coefficients has been modified and do not make sense anymore.");
        return 0;
    }
    if (!coeff) return 0;
    return copy_coeff(this, coeff, num, cudaMemcpyDeviceToHost, false);
}

// reference wt.cu:437-468
void Wavelets::set_coeff(DTYPE* coeff, int num, int mem_is_on_device)
{
    if (!coeff) return;
    norm_cache = 0;
    copy_coeff(this, coeff, num, mem_is_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, true);
}

// reference wt.cu:560-583.  Separable mode: the 1-D analysis pair.  Non-separable mode: four len x len filters
// (nonseparable.cu:86-95); the object's 1-D banks are cleared, the quadruple of each direction lives in the handle.
int Wavelets::set_filters_forward(char* filtername, unsigned int len, DTYPE* filter1, DTYPE* filter2, DTYPE* filter3,
                                  DTYPE* filter4)
{
    if (len > PDWT_MAX_FILTER_WIDTH) {
        printf("ERROR: Wavelets.set_filters_forward(): filter length (%d) exceeds the maximum size (%d)\n", len,
               PDWT_MAX_FILTER_WIDTH);
        return -1;
    }
    if (!do_separable && (filter3 == NULL || filter4 == NULL)) {
        puts("ERROR: Wavelets.set_filters_forward(): expected argument 4 and 5 for non-separable filtering");
        return -2;
    }
    if (!filter1 || !filter2 || len < 1) return -3;
    float Z[PDWT_MAX_FILTER_WIDTH] = {0};
    pdwt_filters* nf = NULL;
    if (do_separable) {
        if (pdwt_filters_create_custom(&nf, (int)len, filter1, filter2, Z, Z) < 0) return -3;
    } else {
        if (pdwt_filters_create_custom(&nf, (int)len, Z, Z, Z, Z) < 0) return -3;
        if (pdwt_filters_set_2d(nf, 1, filter1, filter2, filter3, filter4) < 0) {
            pdwt_filters_destroy(nf);
            return -3;
        }
    }
    if (filters) {
        cudaStreamSynchronize((cudaStream_t)stream);   // kernels in flight may still read the old quadruple
        pdwt_filters_destroy(filters);
    }
    filters = nf;
    winfos.hlen = (int)len;
    if (filtername) {
        memset(wname, 0, sizeof wname);
        strncpy(wname, filtername, sizeof(wname) - 1);
    }
    return 0;
}

// reference wt.cu:585-602: same length as the forward filters
int Wavelets::set_filters_inverse(DTYPE* filter1, DTYPE* filter2, DTYPE* filter3, DTYPE* filter4)
{
    if (!do_separable) {
        if (filter3 == NULL || filter4 == NULL) {
            puts("ERROR: Wavelets.set_filters_inverse(): expected argument 4 and 5 for non-separable filtering");
            return -2;
        }
        if (!filters || !filter1 || !filter2) return -3;
        cudaStreamSynchronize((cudaStream_t)stream);
        return pdwt_filters_set_2d(filters, -1, filter1, filter2, filter3, filter4) < 0 ? -3 : 0;
    }
    if (!filters || !filter1 || !filter2) return -3;
    float L[PDWT_MAX_FILTER_WIDTH], H[PDWT_MAX_FILTER_WIDTH];
    const int hlen = pdwt_filters_get(filters, L, H, NULL, NULL);
    pdwt_filters* nf = NULL;
    if (pdwt_filters_create_custom(&nf, hlen, L, H, filter1, filter2) < 0) return -3;
    cudaStreamSynchronize((cudaStream_t)stream);
    pdwt_filters_destroy(filters);
    filters = nf;
    return 0;
}

__intptr_t Wavelets::image_int_ptr(void) { return (__intptr_t)d_image; }                 // wt.cu:660
__intptr_t Wavelets::coeff_int_ptr(int num) { return d_coeffs ? (__intptr_t)d_coeffs[num] : 0; }  // wt.cu:665

// reference wt.cu:513-552
void Wavelets::print_informations()
{
    const char* yn[2] = {"no", "yes"};
    puts("------------- Wavelet transform infos ------------");
    printf("Data dimensions : ");
    if (winfos.ndims == 2)
        printf("(%d, %d)\n", winfos.Nr, winfos.Nc);
    else if (winfos.Nr == 1)
        printf("%d\n", winfos.Nc);
    else
        printf("(%d, %d) [batched 1D transform]\n", winfos.Nr, winfos.Nc);
    if (batch > 1) printf("Batch : %d planes\n", batch);
    printf("Wavelet name : %s\n", wname);
    printf("Number of levels : %d\n", winfos.nlevels);
    printf("Stationary WT : %s\n", yn[winfos.do_swt != 0]);
    printf("Cycle spinning : %s\n", yn[do_cycle_spinning != 0]);
    printf("Separable transform : %s\n", yn[do_separable != 0]);
    size_t elems = 3 * (size_t)winfos.Nr * winfos.Nc;  // image + 2 tmp
    for (int k = 0; k < pdwt_num_coeffs(winfos); k++) elems += pdwt_coeff_alloc_elems(winfos, k);
    printf("Estimated memory footprint : %.2f MB\n", elems * batch * sizeof(DTYPE) / 1e6);
    int device = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&device) == cudaSuccess && cudaGetDeviceProperties(&prop, device) == cudaSuccess)
        printf("Running on device : %s\n", prop.name);
    puts("--------------------------------------------------");
}

// ================================================================================================ Layer B

extern "C" {

int pdwt_wavelets_create(pdwt_wavelets** out, const float* img, int Nr, int Nc, const char* wname, int levels,
                         int memisonhost, int do_separable, int do_cycle_spinning, int do_swt, int ndim, int batch)
{
    if (!out) return PDWT_ERR_ARG;
    *out = new (std::nothrow) pdwt_wavelets(img, Nr, Nc, wname, levels, memisonhost, do_separable, do_cycle_spinning,
                                            do_swt, ndim, batch);
    return *out ? PDWT_OK : PDWT_ERR_ALLOC;
}
int pdwt_wavelets_copy(pdwt_wavelets** out, const pdwt_wavelets* src)
{
    if (!out || !src) return PDWT_ERR_ARG;
    *out = new (std::nothrow) pdwt_wavelets(src->W);
    return *out ? PDWT_OK : PDWT_ERR_ALLOC;
}
void pdwt_wavelets_destroy(pdwt_wavelets* w) { delete w; }

#define CHECK_W if (!w) return PDWT_ERR_ARG
static int after(pdwt_wavelets* w, int errstate)
{
    return ((int)w->W.state == errstate) ? w->W.last_error : PDWT_OK;
}
int pdwt_wavelets_forward(pdwt_wavelets* w)
{
    CHECK_W;
    if (w->W.state == W_CREATION_ERROR) return PDWT_ERR_STATE;
    w->W.forward();
    return after(w, W_FORWARD_ERROR);
}
int pdwt_wavelets_inverse(pdwt_wavelets* w)
{
    CHECK_W;
    if (w->W.state == W_INVERSE || w->W.state == W_CREATION_ERROR) return PDWT_ERR_STATE;
    w->W.inverse();
    return after(w, W_INVERSE_ERROR);
}
int pdwt_wavelets_soft_threshold(pdwt_wavelets* w, float beta, int app, int normalize)
{
    CHECK_W;
    if (w->W.state == W_INVERSE || w->W.state == W_CREATION_ERROR) return PDWT_ERR_STATE;
    w->W.soft_threshold(beta, app, normalize);
    return after(w, W_THRESHOLD_ERROR);
}
int pdwt_wavelets_hard_threshold(pdwt_wavelets* w, float beta, int app, int normalize)
{
    CHECK_W;
    if (w->W.state == W_INVERSE || w->W.state == W_CREATION_ERROR) return PDWT_ERR_STATE;
    w->W.hard_threshold(beta, app, normalize);
    return after(w, W_THRESHOLD_ERROR);
}
int pdwt_wavelets_group_soft_threshold(pdwt_wavelets* w, float beta, int app, int normalize)
{
    CHECK_W;
    if (w->W.state == W_INVERSE || w->W.state == W_CREATION_ERROR) return PDWT_ERR_STATE;
    w->W.group_soft_threshold(beta, app, normalize);
    return after(w, W_THRESHOLD_ERROR);
}
int pdwt_wavelets_shrink(pdwt_wavelets* w, float beta, int app)
{
    CHECK_W;
    if (w->W.state == W_INVERSE || w->W.state == W_CREATION_ERROR) return PDWT_ERR_STATE;
    w->W.shrink(beta, app);
    return after(w, W_THRESHOLD_ERROR);
}
int pdwt_wavelets_proj_linf(pdwt_wavelets* w, float beta, int app)
{
    CHECK_W;
    if (w->W.state == W_INVERSE || w->W.state == W_CREATION_ERROR) return PDWT_ERR_STATE;
    w->W.proj_linf(beta, app);
    return after(w, W_THRESHOLD_ERROR);
}
int pdwt_wavelets_circshift(pdwt_wavelets* w, int sr, int sc, int inplace)
{
    CHECK_W;
    if (w->W.state == W_CREATION_ERROR) return PDWT_ERR_STATE;
    w->W.circshift(sr, sc, inplace);
    return PDWT_OK;
}
int pdwt_wavelets_add_wavelet(pdwt_wavelets* w, const pdwt_wavelets* other, float alpha)
{
    if (!w || !other) return PDWT_ERR_ARG;
    return w->W.add_wavelet_ref(other->W, alpha);   // no by-value copy across the C boundary
}
int pdwt_wavelets_current_shift(const pdwt_wavelets* w, int* sr, int* sc)
{
    if (!w) return PDWT_ERR_ARG;
    if (sr) *sr = w->W.current_shift_r;
    if (sc) *sc = w->W.current_shift_c;
    return PDWT_OK;
}
int pdwt_wavelets_norm1(pdwt_wavelets* w, float* out) { CHECK_W; return w->W.norm1_batched(out); }
int pdwt_wavelets_norm2sq(pdwt_wavelets* w, float* out) { CHECK_W; return w->W.norm2sq_batched(out); }
int pdwt_wavelets_get_image(pdwt_wavelets* w, float* img) { CHECK_W; return w->W.get_image(img); }
int pdwt_wavelets_get_coeff(pdwt_wavelets* w, float* coeff, int num) { CHECK_W; return w->W.get_coeff(coeff, num); }
int pdwt_wavelets_set_image(pdwt_wavelets* w, const float* img, int on_device)
{
    CHECK_W;
    w->W.set_image(const_cast<float*>(img), on_device);
    return PDWT_OK;
}
int pdwt_wavelets_set_coeff(pdwt_wavelets* w, const float* coeff, int num, int on_device)
{
    CHECK_W;
    w->W.set_coeff(const_cast<float*>(coeff), num, on_device);
    return PDWT_OK;
}
int pdwt_wavelets_set_filters_forward(pdwt_wavelets* w, const char* name, unsigned len, const float* lo, const float* hi)
{
    CHECK_W;
    return w->W.set_filters_forward(const_cast<char*>(name), len, const_cast<float*>(lo), const_cast<float*>(hi));
}
int pdwt_wavelets_set_filters_inverse(pdwt_wavelets* w, const float* lo, const float* hi)
{
    CHECK_W;
    return w->W.set_filters_inverse(const_cast<float*>(lo), const_cast<float*>(hi));
}
int pdwt_wavelets_set_filters_forward_2d(pdwt_wavelets* w, const char* name, unsigned len, const float* f1, const float* f2,
                                         const float* f3, const float* f4)
{
    CHECK_W;
    return w->W.set_filters_forward(const_cast<char*>(name), len, const_cast<float*>(f1), const_cast<float*>(f2),
                                    const_cast<float*>(f3), const_cast<float*>(f4));
}
int pdwt_wavelets_set_filters_inverse_2d(pdwt_wavelets* w, const float* f1, const float* f2, const float* f3, const float* f4)
{
    CHECK_W;
    return w->W.set_filters_inverse(const_cast<float*>(f1), const_cast<float*>(f2), const_cast<float*>(f3),
                                    const_cast<float*>(f4));
}
int pdwt_wavelets_sync(pdwt_wavelets* w)
{
    CHECK_W;
    return note_cuda(cudaStreamSynchronize((cudaStream_t)w->W.stream));
}
int pdwt_wavelets_set_stream(pdwt_wavelets* w, void* stream)
{
    CHECK_W;
    if (w->W.stream != stream) {
        // the norm cache is published in the order of the stream that ran the threshold: finish that stream's work
        // before the object moves on, and forget the cache
        const int rc = note_cuda(cudaStreamSynchronize((cudaStream_t)w->W.stream));
        w->W.invalidate_norm_cache();
        w->W.stream = stream;
        return rc;
    }
    return PDWT_OK;
}
int pdwt_wavelets_invalidate_norm_cache(pdwt_wavelets* w)
{
    CHECK_W;
    w->W.invalidate_norm_cache();
    return PDWT_OK;
}
int pdwt_wavelets_set_async(pdwt_wavelets* w, int on)
{
    CHECK_W;
    w->W.async_copies = on ? 1 : 0;
    return PDWT_OK;
}
int pdwt_wavelets_state(const pdwt_wavelets* w) { return w ? (int)w->W.state : (int)W_CREATION_ERROR; }
pdwt_w_info pdwt_wavelets_info(const pdwt_wavelets* w)
{
    pdwt_w_info z;
    memset(&z, 0, sizeof z);
    return w ? w->W.winfos : z;
}
int pdwt_wavelets_batch(const pdwt_wavelets* w) { return w ? w->W.batch : 0; }
int pdwt_wavelets_do_separable(const pdwt_wavelets* w) { return w ? w->W.do_separable : 0; }
const char* pdwt_wavelets_wname(const pdwt_wavelets* w) { return w ? w->W.wname : ""; }
intptr_t pdwt_wavelets_image_int_ptr(const pdwt_wavelets* w) { return w ? (intptr_t)w->W.d_image : 0; }
intptr_t pdwt_wavelets_coeff_int_ptr(const pdwt_wavelets* w, int num)
{
    if (!w || !w->W.d_coeffs || num < 0 || num >= pdwt_num_coeffs(w->W.winfos)) return 0;
    return (intptr_t)w->W.d_coeffs[num];
}
intptr_t pdwt_wavelets_tmp_int_ptr(const pdwt_wavelets* w) { return w ? (intptr_t)w->W.d_tmp : 0; }
long long pdwt_wavelets_launch_count(const pdwt_wavelets* w) { return w ? w->W.launch_count() : 0; }

}  // extern "C"
