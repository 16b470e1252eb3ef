// ref_shim.cpp -- extern "C" handles onto the UNMODIFIED reference `Wavelets` class (our code, not the
// reference's): lets python (ctypes) drive /root/reference/src/wt.cu's own public API on the GPU box to
//   (a) dump golden vectors (tests/golden/make_golden.py), (b) pin the CPU restatement, (c) time the
//   reference's stock kernels (bench.py --impl reference).
// Built only by `make -C oracle ref` into oracle/_ref/libpdwt_ref.so together with the reference sources where
// they lie (nothing is copied into the repo).  TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h>
#include "wt.h"   // the reference's header, found through -I/root/reference/src

extern "C" {

void* ref_create(float* img, int Nr, int Nc, const char* wname, int levels, int memisonhost, int do_separable,
                 int do_cycle_spinning, int do_swt, int ndim)
{
    return new Wavelets(img, Nr, Nc, wname, levels, memisonhost, do_separable, do_cycle_spinning, do_swt, ndim);
}
void ref_destroy(void* w) { delete static_cast<Wavelets*>(w); }
void ref_forward(void* w) { static_cast<Wavelets*>(w)->forward(); }
void ref_inverse(void* w) { static_cast<Wavelets*>(w)->inverse(); }
void ref_soft_threshold(void* w, float beta, int app, int normalize)
{
    static_cast<Wavelets*>(w)->soft_threshold(beta, app, normalize);
}
void ref_hard_threshold(void* w, float beta, int app, int normalize)
{
    static_cast<Wavelets*>(w)->hard_threshold(beta, app, normalize);
}
void ref_group_soft_threshold(void* w, float beta, int app, int normalize)
{
    static_cast<Wavelets*>(w)->group_soft_threshold(beta, app, normalize);
}
void ref_shrink(void* w, float beta, int app) { static_cast<Wavelets*>(w)->shrink(beta, app); }
void ref_proj_linf(void* w, float beta, int app) { static_cast<Wavelets*>(w)->proj_linf(beta, app); }
void ref_circshift(void* w, int sr, int sc, int inplace) { static_cast<Wavelets*>(w)->circshift(sr, sc, inplace); }
int ref_add_wavelet(void* w, void* other, float alpha)
{
    return static_cast<Wavelets*>(w)->add_wavelet(*static_cast<Wavelets*>(other), alpha);   // by value: deep copy
}
float ref_norm1(void* w) { return static_cast<Wavelets*>(w)->norm1(); }
float ref_norm2sq(void* w) { return static_cast<Wavelets*>(w)->norm2sq(); }
int ref_get_image(void* w, float* out) { return static_cast<Wavelets*>(w)->get_image(out); }
int ref_get_coeff(void* w, float* out, int num) { return static_cast<Wavelets*>(w)->get_coeff(out, num); }
void ref_set_image(void* w, float* img, int on_device) { static_cast<Wavelets*>(w)->set_image(img, on_device); }
void ref_set_coeff(void* w, float* c, int num, int on_device) { static_cast<Wavelets*>(w)->set_coeff(c, num, on_device); }
// custom filter banks, wt.cu:560-602 (separable: f3 = f4 = NULL; non-separable: four len x len filters)
int ref_set_filters_forward(void* w, const char* name, unsigned len, float* f1, float* f2, float* f3, float* f4)
{
    return static_cast<Wavelets*>(w)->set_filters_forward(const_cast<char*>(name), len, f1, f2, f3, f4);
}
int ref_set_filters_inverse(void* w, float* f1, float* f2, float* f3, float* f4)
{
    return static_cast<Wavelets*>(w)->set_filters_inverse(f1, f2, f3, f4);
}
int ref_nlevels(void* w) { return static_cast<Wavelets*>(w)->winfos.nlevels; }
int ref_hlen(void* w) { return static_cast<Wavelets*>(w)->winfos.hlen; }
int ref_ndims(void* w) { return static_cast<Wavelets*>(w)->winfos.ndims; }
int ref_state(void* w) { return (int)static_cast<Wavelets*>(w)->state; }
int ref_sync(void) { return (int)cudaDeviceSynchronize(); }
int ref_last_error(void) { return (int)cudaGetLastError(); }

// fwd+inv timing loop with CUDA events on the legacy default stream the reference launches on.
// Returns milliseconds for `iters` forward()+inverse() pairs (state is reset by forward()).
float ref_time_fwd_inv(void* w, int iters)
{
    Wavelets* W = static_cast<Wavelets*>(w);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; i++) {
        W->forward();
        W->inverse();
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

// `iters` steps of forward()+inverse(), step i on object ws[(first+i) % nw] (rotating objects defeat L2 reuse
// between steps).  CUDA events on the legacy default stream.  Returns milliseconds for all steps.
float ref_time_rotating(void** ws, int nw, int first, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; i++) {
        Wavelets* W = static_cast<Wavelets*>(ws[(first + i) % nw]);
        W->forward();
        W->inverse();
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

// end to end through the reference's public API with host buffers: set_image (H2D) -> forward -> inverse ->
// get_image (D2H), every step.
float ref_time_e2e(void** ws, int nw, int first, int iters, float** host_in, float* host_out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; i++) {
        Wavelets* W = static_cast<Wavelets*>(ws[(first + i) % nw]);
        W->set_image(host_in[(first + i) % nw], 0);
        W->forward();
        W->inverse();
        W->get_image(host_out);
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}
}
