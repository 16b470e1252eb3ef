// pdwt_common.cuh -- shared declarations of libpdwt_b200 (internal; the public surface is include/pdwt_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "../../include/pdwt_b200.h"

namespace pdwt {

// Filter taps travel to every kernel as a launch parameter (constant bank 0) -- the per-instance replacement of the
// reference's process-global `__constant__ c_kern_L/H/IL/IH` (common.h:28-31).  644 bytes.
struct Taps {
    int hlen;
    float L[PDWT_MAX_FILTER_WIDTH];   // analysis low-pass   (c_kern_L)
    float H[PDWT_MAX_FILTER_WIDTH];   // analysis high-pass  (c_kern_H)
    float IL[PDWT_MAX_FILTER_WIDTH];  // synthesis low-pass  (c_kern_IL)
    float IH[PDWT_MAX_FILTER_WIDTH];  // synthesis high-pass (c_kern_IH)
    // Non-separable mode with a CUSTOM filter quadruple (Wavelets::set_filters_forward/_inverse with four 2-D filters,
    // wt.cu:560-602): 4 x hlen*hlen floats, (LL, LH, HL, HH) in the reference's array layout, for the direction of the
    // driver that was called (device copy for the kernels, host copy for launchers that build parameter tables).
    // NULL = the outer products of the 1-D banks (w_outer, nonseparable.cu:16-24).
    const float* k2d;
    const float* hk2d;
};
// filter f (0 LL, 1 LH, 2 HL, 3 HH) at reference index idx = row * hlen + col
__host__ __device__ inline float k2d_at(const float* k2d, int hlen, int f, int idx) { return k2d[f * hlen * hlen + idx]; }

// ---- index rules (reference lines cited at each use) ------------------------------------
__host__ __device__ inline int half_up(int n) { return (n + 1) >> 1; }  // w_div2, utils.cu:24-27

// decimating analysis fold incl. the odd-size "repeat last sample" rule, separable.cu:114-121
__host__ __device__ inline int fold_dec(int i, int N)
{
    const int odd = N & 1;
    if (i < 0) i += N + odd;
    if (i > N - 1) i = (i == N && odd) ? N - 1 : i - (N + odd);
    return i;
}
// analysis centre, separable.cu:98-107
__host__ __device__ inline int centre_fwd(int hlen) { return (hlen & 1) ? hlen / 2 : hlen / 2 - 1; }

// synthesis geometry, separable.cu:249-264: taps per polyphase branch, centre, virtual index shift
struct SynGeom {
    int taps, c, shift;
};
__host__ __device__ inline SynGeom syn_geometry(int hlen)
{
    SynGeom s;
    const int h2 = hlen / 2;
    s.c = h2 / 2;
    s.shift = (h2 & 1) ? 0 : 1;
    s.taps = (h2 & 1) ? 2 * s.c + 1 : 2 * s.c;
    return s;
}
// undecimated fold with a single +-N wrap, separable.cu:423-433 / 575-579
__host__ __device__ inline int fold_swt(int g, int jf, int c, int N)
{
    int i = g + jf - c;
    if (jf < c - g) i += N;
    if (jf > N - 1 - g + c) i -= N;
    return i;
}
__host__ __device__ inline int swt_inv_taps(int hlen) { return (hlen & 1) ? 2 * (hlen / 2) + 1 : 2 * (hlen / 2); }

// ---- error plumbing -----------------------------------------------------------------------------------------
int note_cuda(cudaError_t e);  // records e for pdwt_last_cuda_error(); returns PDWT_OK / PDWT_ERR_CUDA
void count_launch(int n = 1);

#define PDWT_CUDA(call)                                        \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return ::pdwt::note_cuda(e__); \
    } while (0)
#define PDWT_LAUNCH_CHECK()                                      \
    do {                                                         \
        ::pdwt::count_launch();                                  \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return ::pdwt::note_cuda(e__);   \
    } while (0)

// ---- optional per-kernel timing (pdwt_profile_begin/_end): a CUDA-event pair around every launch made inside the
// scope, on the stream the kernel is launched on.  Costs nothing when profiling is off.
bool profiling_on();
void prof_open(const char* tag, cudaStream_t s, void** ev0);
void prof_close(const char* tag, cudaStream_t s, void* ev0);
struct ProfScope {
    const char* tag;
    cudaStream_t s;
    void* ev0;
    ProfScope(const char* t, cudaStream_t st) : tag(t), s(st), ev0(nullptr)
    {
        if (profiling_on()) prof_open(tag, s, &ev0);
    }
    ~ProfScope()
    {
        if (ev0) prof_close(tag, s, ev0);
    }
};
const char* prof_tag(const char* base, int rows, int cols);  // "base[rows x cols]" (interned) while profiling, else base
#define PDWT_PROF(tag, stream) ::pdwt::ProfScope prof_scope__((tag), (stream))

inline int idiv_up(int a, int b) { return (a + b - 1) / b; }

// Function attributes (dynamic shared memory limits) are per DEVICE: PDWT_ONCE_PER_DEVICE(call) runs `call` the first
// time its call site is reached on each device of the process.  The body runs under the site's lock, and the device is
// marked done only when the call has succeeded, so concurrent first launches from several host threads (one object per
// thread) neither skip the attribute nor leave a failed call un-retried.
struct PerDeviceOnce {
    std::mutex mu;
    unsigned long long mask = 0;
    template <typename F>
    cudaError_t run(F body, int* dev_out = nullptr)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev_out) *dev_out = dev;
        std::lock_guard<std::mutex> g(mu);
        if (dev >= 0 && dev < 64 && ((mask >> dev) & 1ull)) return cudaSuccess;
        e = body();
        if (e == cudaSuccess && dev >= 0 && dev < 64) mask |= 1ull << dev;
        return e;
    }
};
#define PDWT_ONCE_PER_DEVICE(call)                                                    \
    do {                                                                              \
        static ::pdwt::PerDeviceOnce once__;                                          \
        cudaError_t eo__ = once__.run([&]() -> cudaError_t { return (call); });       \
        if (eo__ != cudaSuccess) return ::pdwt::note_cuda(eo__);                      \
    } while (0)

// level geometry of one plane
inline int level_size(int N, int l, int do_swt)
{
    if (do_swt) return N;
    for (int i = 0; i < l; i++) N = half_up(N);
    return N;
}

// ---- generic (any size, any hlen <= 40) per-pass launchers, pdwt_generic.cu ---------------------------------
// Each takes plane strides (floats) for the batch dimension (grid.z = plane).
struct Plane2 {  // a pointer + per-plane stride
    float* p;
    size_t stride;
};
int g_fwd_rows(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int batch, cudaStream_t s);
int g_fwd_cols(const Taps& t, Plane2 t1, Plane2 t2, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
               cudaStream_t s);
int g_inv_cols(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 t1, Plane2 t2, int n, int Nc, int M,
               int batch, cudaStream_t s);
int g_inv_rows(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int n, int M, int batch, cudaStream_t s);
int g_swt_fwd_rows(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int level, int batch,
                   cudaStream_t s);
int g_swt_fwd_cols(const Taps& t, Plane2 t1, Plane2 t2, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc,
                   int level, int batch, cudaStream_t s);
int g_swt_inv_cols(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 t1, Plane2 t2, int Nr, int Nc,
                   int level, int batch, cudaStream_t s);
int g_swt_inv_rows(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int Nc, int level, int batch,
                   cudaStream_t s);
int g_haar2d_fwd(Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch, cudaStream_t s);
int g_haar2d_inv(Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int Nr2, int Nc2, int batch,
                 cudaStream_t s);
int g_haar1d_fwd(Plane2 img, Plane2 A, Plane2 D, int Nr, int Nc, int batch, cudaStream_t s);
int g_haar1d_inv(Plane2 img, Plane2 A, Plane2 D, int Nr, int Nc, int Nc2, int batch, cudaStream_t s);
int g_nonsep_fwd(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                 cudaStream_t s);
int g_nonsep_inv(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int Nr2, int Nc2,
                 int batch, cudaStream_t s);
int g_nonsep_swt_fwd(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                     int batch, cudaStream_t s);
int g_nonsep_swt_inv(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                     int batch, cudaStream_t s);

// ---- fused fast paths, pdwt_fused.cu: return 1 if they handled the level, 0 if the shape is not covered ---
int f_dwt2_fwd_level(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                     cudaStream_t s);
int f_dwt2_inv_level(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int nr, int nc, int Mr, int Mc,
                     int batch, cudaStream_t s);

// ---- warp-streaming FFMA2 kernels, pdwt_stream.cu.  One call takes SEVERAL consecutive levels and serves them with one
// launch (cross-level queue, see the file); `plans` caches the queues and their device counters (NULL: one level per call).
struct StreamLevelIO {   // one level: image / approximation `img` (Nr x Nc)  <->  sub-bands A, H, V, D (Nr/2 x Nc/2)
    Plane2 img, A, H, V, D;
    int Nr, Nc;
    int a_reused;        // forward: A is read again by a later kernel (keep it in L2); inverse: same for img
};
struct StreamPlans;
StreamPlans* stream_plans_create();
void stream_plans_destroy(StreamPlans* p);
// forward: lv[0] is the finest level, lv[i+1].img == lv[i].A.  Returns how many LEADING levels were launched (0: the
// first one's shape is not covered), < 0 on error.
int s_dwt2_fwd_levels(const Taps& t, StreamPlans* plans, const StreamLevelIO* lv, int nlev, int batch, cudaStream_t s);
// inverse: lv[0] is the coarsest level of the group, lv[i+1].A == lv[i].img.  Same return convention.
int s_dwt2_inv_levels(const Taps& t, StreamPlans* plans, const StreamLevelIO* lv, int nlev, int batch, cudaStream_t s);

// ---- fused SWT level kernels, pdwt_swt.cu: same convention; w_swt2_supported tells whether EVERY level 1..nlevels of a
// transform can take them (the caller's buffer plan depends on it)
int w_swt2_supported(const Taps& t, int Nr, int Nc, int nlevels, int batch);
int w_swt2_fwd_level(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                     int batch, cudaStream_t s);
int w_swt2_inv_level(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int Nr, int Nc, int level,
                     int batch, cudaStream_t s);

// ---- staged row-pass kernels of the (batched) 1-D transforms, pdwt_rows.cu: same convention; the g_*_rows launchers try
// them first
int r_fwd_rows(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int batch, cudaStream_t s);
int r_inv_rows(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int n, int M, int batch, cudaStream_t s);
int r_swt_fwd_rows(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int level, int batch, cudaStream_t s);
int r_swt_inv_rows(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int Nc, int level, int batch, cudaStream_t s);

// ---- all levels of a batched 1-D DWT in one launch (a row lives in shared memory), pdwt_rows_all.cu: same convention.
// bands[0] = A_L, bands[l] = D_l (l = 1..L)
int r_dwt1_fwd_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s);
int r_dwt1_inv_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s);
int r_swt1_fwd_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s);
int r_swt1_inv_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s);

// ---- register-tiled non-separable DWT level kernels, pdwt_nonsep.cu: same convention
int n_nonsep_fwd_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                       cudaStream_t s);
int n_nonsep_inv_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int Nr2, int Nc2,
                       int batch, cudaStream_t s);

// ---- register-tiled non-separable SWT level kernels, pdwt_nonsep_swt.cu: same convention (level = 1, 2, ...)
int n_nonsep_swt_fwd_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                           int batch, cudaStream_t s);
int n_nonsep_swt_inv_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                           int batch, cudaStream_t s);

// ---- element-wise + reductions, pdwt_elementwise.cu ----------------------------------------------------------
constexpr int kMaxSeg = 64;
struct SegTable {  // list of (sub-band, length, parameter) handled by one launch
    int nseg;
    float* ptr[kMaxSeg];
    unsigned long long n[kMaxSeg];       // floats per plane to process
    unsigned long long stride[kMaxSeg];  // floats between planes
    float beta[kMaxSeg];
    unsigned char ro[kMaxSeg];           // 1: the element-wise kernels only READ this segment (it is there for its norm)
};
// Hand-over of the sums to the host without a memset in front of the launch, a copy behind it or a system-scope fence
// inside it: the LAST block to finish (device ticket) writes every double into mapped pinned memory as two 8-byte words
// {32 data bits, 32-bit tag of this launch} -- an aligned 8-byte store is the unit that reaches host memory whole (the
// scheme of NCCL's LL protocol) -- and leaves the device scratch and the ticket zeroed for the next launch.  The host
// spins until both words of every sum carry the tag (wait_published) instead of cudaMemcpyAsync + cudaStreamSynchronize.
struct HostPublish {
    unsigned long long* h_out;   // device view of the pinned words: 2 per sum
    unsigned* ticket;            // zero between launches
    unsigned tag;                // announces THIS launch's sums (never 0)
    int nsums;
};
// sums != NULL: the kernel also accumulates sum |out| into sums[plane*nseg + seg] and sum out^2 into
// sums[batch*nseg + plane*nseg + seg] (the norm of the thresholded coefficients comes for free)
int e_threshold(const SegTable& tab, int op /*0 soft, 1 hard, 2 proj_linf, 3 scale by beta*/, int batch, cudaStream_t s,
                double* sums = nullptr, const HostPublish* hp = nullptr);
struct GroupTable {  // group soft threshold: per level the detail triple (h, v NULL in 1-D) and, optionally, A
    int nlev;
    float *h[32], *v[32], *d[32], *a[32];
    unsigned long long n[32], stride_d[32], stride_a;
    float beta[32];
};
int e_group_soft(const GroupTable& tab, int batch, cudaStream_t s);
struct PairTable {  // dst += alpha * src, sub-band by sub-band
    int nseg;
    float* dst[kMaxSeg];
    const float* src[kMaxSeg];
    unsigned long long n[kMaxSeg], stride_dst[kMaxSeg], stride_src[kMaxSeg];
};
int e_axpy(const PairTable& tab, float alpha, int batch, cudaStream_t s);
int e_circshift(const float* in, float* out, size_t stride, int Nr, int Nc, int sr, int sc, int batch, cudaStream_t s);
// sums[plane*nseg + seg] (double, device) += sum |v| (mode 0) or sum v^2 (mode 1)
int e_reduce(const SegTable& tab, int mode, int batch, double* d_sums, cudaStream_t s, const HostPublish* hp = nullptr);

}  // namespace pdwt

namespace pdwt {
bool fused_supports_hlen(int hlen);

// Programmatic dependent launch: the level kernels of one transform are queued back to back on one stream; each lets
// its successor's CTAs start (launch latency, barrier init, address set-up) while it is still draining, and blocks
// at pdl_wait() until its predecessor has completed and flushed before touching global memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// PDWT_PDL (default 2, measured on B200, C2: 112.6 -> 98.4 us per fwd+inv): 0 = plain stream order; 1 = dependents may launch as soon as every CTA of this kernel has started (measured
// on B200, C2: ~6% slower, the early CTAs hold shared memory and registers while they wait); 2 = dependents may launch
// when every CTA is in its last row pair (only the launch latency and the prologue overlap the tail).
inline int pdl_mode()
{
    const char* e = getenv("PDWT_PDL");
    return e ? atoi(e) : 2;
}

template <typename Kern, typename... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, unsigned block, size_t smem, cudaStream_t s, const Args&... args)
{
    const bool no_pdl = pdl_mode() == 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

}  // namespace pdwt
