/*
 * wt.h -- drop-in C++ face of the B200-native PDWT hot path.
 *
 * Same class name, same public data members in the same order, same constructor and method signatures as the
 * reference header (reference src/wt.h:20-76, enum w_state wt.h:8-17, struct w_info utils.h:9-19), so code written
 * against PDWT (demo.cpp, the pypwt Cython binding) recompiles against this header and links libpdwt_b200.so
 * instead of libpdwt.so.  The methods forward to the C ABI in pdwt_b200.h; the members added after `state` are
 * implementation details of this library.
 *
 * Differences in behaviour are listed in DESIGN.md ("Deliberate deviations"); the ones visible here: filters are
 * per instance (no process-global constant memory), an unknown wavelet name yields state == W_CREATION_ERROR
 * instead of hanging, norm2sq() returns the sum of squares also for 1-D transforms, CUDA errors are recorded in
 * `last_error` instead of being ignored, and inverse() leaves d_coeffs[0] intact on the fused kernel families.
 * get_image() / get_coeff() return the element count like the reference, saturated at INT_MAX for batches of 2^31
 * floats or more (0 still means failure).
 *
 * norm1() / norm2sq() right after soft_threshold() / hard_threshold() return sums that the threshold kernel produced
 * on its way (no second pass over the coefficients).  Writes through the raw pointers of coeff_int_ptr() are invisible
 * to that cache: call invalidate_norm_cache() after modifying coefficients behind the object's back.
 */
#ifndef WT_H
#define WT_H

#include "pdwt_b200.h"

namespace pdwt { struct HostPublish; }   /* internal (csrc/pdwt_common.cuh) */

#ifndef DTYPE
#define DTYPE float /* reference filters.h:16-23 (single-precision build) */
#endif

typedef pdwt_w_info w_info; /* utils.h:9-19 */

typedef enum w_state { /* wt.h:8-17 */
    W_INIT, W_FORWARD, W_INVERSE, W_THRESHOLD, W_CREATION_ERROR, W_FORWARD_ERROR, W_INVERSE_ERROR, W_THRESHOLD_ERROR
} w_state;

class Wavelets {
  public:
    // ---- reference members (wt.h:24-33), same order and meaning
    DTYPE* d_image;    // image (input, or result of inverse()), device, batch * Nr*Nc
    DTYPE** d_coeffs;  // host array of device sub-band pointers [A, H1,V1,D1, ...] / [A, D1, ...]
    DTYPE* d_tmp;      // device scratch, batch * 2*Nr*Nc
    int current_shift_r;
    int current_shift_c;
    char wname[128];
    int do_separable;
    int do_cycle_spinning;
    w_info winfos;
    w_state state;
    // ---- extensions
    int batch;         // independent planes handled by every method (1 = the reference's behaviour)
    int last_error;    // pdwt_status of the most recent failing call (0 if none)
    void* stream;      // cudaStream_t all work is enqueued on (NULL = legacy default stream)
    int async_copies;  // 1: get_image/set_image/get_coeff/set_coeff only ENQUEUE their host copies on `stream` (pinned
                       // host memory; the caller synchronises) so that objects on different streams pipeline
                       // H2D / kernels / D2H.  0 = the reference's blocking behaviour (wt.cu:421-434)

    Wavelets();
    Wavelets(DTYPE* img, int Nr, int Nc, const char* wname, int levels, int memisonhost = 1, int do_separable = 1,
             int do_cycle_spinning = 0, int do_swt = 0, int ndim = 2, int batch = 1);
    Wavelets(const Wavelets& W);
    ~Wavelets();

    void forward();
    void soft_threshold(DTYPE beta, int do_thresh_appcoeffs = 0, int normalize = 0);
    void hard_threshold(DTYPE beta, int do_thresh_appcoeffs = 0, int normalize = 0);
    void group_soft_threshold(DTYPE beta, int do_thresh_appcoeffs = 0, int normalize = 0);
    void shrink(DTYPE beta, int do_thresh_appcoeffs = 1);
    void proj_linf(DTYPE beta, int do_thresh_appcoeffs = 1);
    void circshift(int sr, int sc, int inplace = 1);
    void inverse();
    DTYPE norm2sq();
    DTYPE norm1();
    int get_image(DTYPE* img);
    void print_informations();
    int get_coeff(DTYPE* coeff, int num);
    void set_image(DTYPE* img, int mem_is_on_device = 0);
    void set_coeff(DTYPE* coeff, int num, int mem_is_on_device = 0);
    int set_filters_forward(char* filtername, unsigned int len, DTYPE* filter1, DTYPE* filter2, DTYPE* filter3 = 0,
                            DTYPE* filter4 = 0);
    int set_filters_inverse(DTYPE* filter1, DTYPE* filter2, DTYPE* filter3 = 0, DTYPE* filter4 = 0);
    int add_wavelet(Wavelets W, DTYPE alpha = 1.0f);              // by value, like the reference (wt.h:73)
    int add_wavelet_ref(const Wavelets& W, DTYPE alpha = 1.0f);   // the same without the deep copy
    __intptr_t image_int_ptr(void);
    __intptr_t coeff_int_ptr(int num);

    // batch-aware variants of the scalar queries: out[batch] on the host
    int norm1_batched(DTYPE* out);
    int norm2sq_batched(DTYPE* out);
    long long launch_count() const { return launches; }
    void invalidate_norm_cache() { norm_cache = 0; }   // after writes through coeff_int_ptr()

  private:
    Wavelets& operator=(const Wavelets&);  // "do not use" in the reference (wt.cu:36-73)
    pdwt_filters* filters;
    double* d_sums;   // device scratch of the norm reductions
    double* h_sums;   // pinned host mirror
    long long launches;
    int norm_cache;   // bit 0 / 1: the L1 / L2 sums of the current coefficients were published by the last threshold
    double* h_sums_dev;            // device view of h_sums (mapped pinned memory)
    unsigned thr_tag;              // tag that announces the last threshold's sums in h_sums
    void publish_target(int which, pdwt::HostPublish* hp) const;
    int alloc_buffers();
    void free_buffers();
    int norms(int mode, DTYPE* out);
};

#endif
