// pdwt_rows.cu -- row-pass kernels for the (batched) 1-D transforms: separable DWT and SWT, forward and inverse
// (SURVEY 8 a-1, a-5, a-7, a-8 in their 1-D role: separable.cu:213-236, 368-395, 519-537, 653-672).
//
// The reference gives one thread one output and lets it gather its taps from global memory with a fold per tap.  A 1-D
// level is pure streaming (14 FMA per sample for db7 against 8 bytes), so these kernels only organise the traffic:
// a CTA stages one segment of a row -- halo included, the reference's periodic / odd-size fold applied once per staged
// element -- in shared memory with asynchronous copies, then every thread produces 4 (DWT forward, SWT) or 8 (DWT
// inverse) consecutive outputs from 128-bit shared loads and stores them as 128-bit vectors.  Any row length.
// Arithmetic is the reference's chain per output (fmaf from 0 in ascending tap order; the SWT inverse's
// round(v*k), halve, add), so results are bit-identical to the generic kernels and to the reference.
#include <string.h>

#include "pdwt_common.cuh"

namespace pdwt {

constexpr int kRowThreads = 256;

__device__ __forceinline__ void rows_cp4(float* smem_dst, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void rows_cp_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ int rows_clamp(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }
__device__ __forceinline__ int rows_wrap1(int i, int N)   // single +-N wrap (separable.cu:265-273, 428-433), then clamp
{
    i += (i < 0) ? N : 0;
    i -= (i >= N) ? N : 0;
    return rows_clamp(i, N - 1);
}
__device__ __forceinline__ void rows_store4(float* p, const float (&v)[4], int g, int n, bool vec)
{
    if (vec && g + 4 <= n) {
        *reinterpret_cast<float4*>(p + g) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (g + i < n) p[g + i] = v[i];
    }
}

// ============================================================================================ DWT forward rows
// lo/hi[k] = sum_j x[fold(2k - C + j)] * L/H[hlen-1-j]   (w_kern_forward_pass1, separable.cu:91-131)
template <int HLEN>
struct RowFwdCfg {
    static constexpr int SEG = 4 * kRowThreads;               // outputs per CTA
    static constexpr int C = HLEN / 2 - 1;
    static constexpr int NV = (6 + HLEN + 3) / 4;             // 16-byte vectors per thread window
    static constexpr int NS = 8 * (kRowThreads - 1) + 4 * NV; // staged floats
};
template <int HLEN>
__global__ void __launch_bounds__(kRowThreads)
    k_rows_dwt_fwd(const __grid_constant__ Taps t, const float* __restrict__ img, size_t s_img, float* __restrict__ lo,
                   size_t s_lo, float* __restrict__ hi, size_t s_hi, int Nc, int n, int vec)
{
    using K = RowFwdCfg<HLEN>;
    __shared__ __align__(16) float S[K::NS];
    const int tid = threadIdx.x, k0 = blockIdx.x * K::SEG;
    const size_t row = blockIdx.y, pz = blockIdx.z;
    const float* x = img + pz * s_img + row * Nc;
    pdl_wait();
    {   // staged columns [u_lo, u_hi) lie inside the row: only the others pay for the fold
        const int xs = 2 * k0 - K::C;
        const int u_lo = max(0, -xs), u_hi = min(K::NS, Nc - xs);
#pragma unroll
        for (int k = 0; k < (K::NS + kRowThreads - 1) / kRowThreads; k++) {
            const int u = tid + kRowThreads * k;
            if (u < K::NS) {
                int xi = xs + u;
                if (u < u_lo || u >= u_hi) xi = rows_clamp(fold_dec(xi, Nc), Nc - 1);
                rows_cp4(S + u, x + xi);
            }
        }
    }
    rows_cp_wait();
    __syncthreads();
    pdl_launch_dependents();
    const int k = k0 + 4 * tid;
    if (k >= n) return;
    float w[4 * K::NV];
#pragma unroll
    for (int i = 0; i < K::NV; i++) {
        const float4 f = *reinterpret_cast<const float4*>(S + 8 * tid + 4 * i);
        w[4 * i] = f.x; w[4 * i + 1] = f.y; w[4 * i + 2] = f.z; w[4 * i + 3] = f.w;
    }
    float al[4] = {0.f, 0.f, 0.f, 0.f}, ah[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < HLEN; j++) {
        const float kl = t.L[HLEN - 1 - j], kh = t.H[HLEN - 1 - j];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            al[i] = fmaf(w[2 * i + j], kl, al[i]);
            ah[i] = fmaf(w[2 * i + j], kh, ah[i]);
        }
    }
    rows_store4(lo + pz * s_lo + row * n, al, k, n, vec);
    rows_store4(hi + pz * s_hi + row * n, ah, k, n, vec);
}

// ============================================================================================ DWT inverse rows
// img[2m+e] = sum_j t1[wrap(m - CC + e*SHIFT + j)] * IL[hlen-1-(2j+off_e)] + (same with t2, IH), off_e = e ? SHIFT : 1-SHIFT
// (w_kern_inverse_pass2, separable.cu:293-328)
template <int HLEN>
struct RowInvCfg {
    static constexpr int H2 = HLEN / 2, CC = H2 / 2, SHIFT = (H2 & 1) ? 0 : 1, WIN = H2 + SHIFT;
    static constexpr int SEGC = 4 * kRowThreads;              // coefficient positions per CTA (8 outputs per thread)
    static constexpr int NV = (3 + WIN + 3) / 4;
    static constexpr int NS = 4 * (kRowThreads - 1) + 4 * NV;
};
template <int HLEN>
__global__ void __launch_bounds__(kRowThreads)
    k_rows_dwt_inv(const __grid_constant__ Taps t, const float* __restrict__ t1, size_t s_1, const float* __restrict__ t2,
                   size_t s_2, float* __restrict__ img, size_t s_img, int n, int M, int vec)
{
    using K = RowInvCfg<HLEN>;
    __shared__ __align__(16) float S1[K::NS], S2[K::NS];
    const int tid = threadIdx.x, m0 = blockIdx.x * K::SEGC;
    const size_t row = blockIdx.y, pz = blockIdx.z;
    const float* p1 = t1 + pz * s_1 + row * n;
    const float* p2 = t2 + pz * s_2 + row * n;
    pdl_wait();
    {
        const int xs = m0 - K::CC;
        const int u_lo = max(0, -xs), u_hi = min(K::NS, n - xs);   // no wrap needed inside [u_lo, u_hi)
#pragma unroll
        for (int k = 0; k < (K::NS + kRowThreads - 1) / kRowThreads; k++) {
            const int u = tid + kRowThreads * k;
            if (u < K::NS) {
                int x = xs + u;
                if (u < u_lo || u >= u_hi) x = rows_wrap1(x, n);
                rows_cp4(S1 + u, p1 + x);
                rows_cp4(S2 + u, p2 + x);
            }
        }
    }
    rows_cp_wait();
    __syncthreads();
    pdl_launch_dependents();
    const int m = m0 + 4 * tid;
    if (2 * m >= M) return;
    float w1[4 * K::NV], w2[4 * K::NV];
#pragma unroll
    for (int i = 0; i < K::NV; i++) {
        const float4 f = *reinterpret_cast<const float4*>(S1 + 4 * tid + 4 * i);
        const float4 g = *reinterpret_cast<const float4*>(S2 + 4 * tid + 4 * i);
        w1[4 * i] = f.x; w1[4 * i + 1] = f.y; w1[4 * i + 2] = f.z; w1[4 * i + 3] = f.w;
        w2[4 * i] = g.x; w2[4 * i + 1] = g.y; w2[4 * i + 2] = g.z; w2[4 * i + 3] = g.w;
    }
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int p = q >> 1, e = q & 1, i0 = p + (e ? K::SHIFT : 0), off = e ? K::SHIFT : 1 - K::SHIFT;
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int j = 0; j < K::H2; j++) {
            a1 = fmaf(w1[i0 + j], t.IL[HLEN - 1 - (2 * j + off)], a1);
            a2 = fmaf(w2[i0 + j], t.IH[HLEN - 1 - (2 * j + off)], a2);
        }
        o[q] = __fadd_rn(a1, a2);
    }
    float* out = img + pz * s_img + row * M;
    const float oa[4] = {o[0], o[1], o[2], o[3]}, ob[4] = {o[4], o[5], o[6], o[7]};
    rows_store4(out, oa, 2 * m, M, vec);
    rows_store4(out, ob, 2 * m + 4, M, vec);
}

// ================================================================================================= SWT rows
// forward: lo/hi[g] = sum_j x[wrap(g + (j - C) f)] * L/H[hlen-1-j]            (separable.cu:409-448)
// inverse: img[g]   = sum_j t1[wrap(g + (j - C') f)] * IL[hlen-1-j] / 2 + ... (separable.cu:593-626), C' = hlen/2
// FMODE 1 / 2: f = 1 / 2, a thread's taps come from one register window; FMODE 4: f % 4 == 0, one aligned 128-bit shared
// load per tap.
template <int HLEN, int FMODE>
struct RowSwtCfg {
    static constexpr int SEG = 4 * kRowThreads;
    static constexpr int NV = (FMODE * (HLEN - 1) + 4 + 3) / 4;   // window vectors (FMODE 1, 2)
};
static inline size_t rows_swt_smem(int hlen, int f, int arrays)
{
    return sizeof(float) * arrays * (size_t)((4 * kRowThreads + (hlen - 1) * f + 8 + 3) & ~3);
}

template <int HLEN, int FMODE, bool INV>
__global__ void __launch_bounds__(kRowThreads)
    k_rows_swt(const __grid_constant__ Taps t, const float* __restrict__ in1, size_t s_1, const float* __restrict__ in2,
               size_t s_2, float* __restrict__ out1, size_t s_o1, float* __restrict__ out2, size_t s_o2, int Nc, int f, int vec)
{
    using K = RowSwtCfg<HLEN, FMODE>;
    extern __shared__ __align__(16) float smem[];
    const int ns = K::SEG + (HLEN - 1) * f + 8;   // staged floats per array (multiple of 4: f is 1, 2 or a multiple of 4)
    const int pitch = (ns + 3) & ~3;
    float* S1 = smem;
    float* S2 = smem + pitch;                      // INV only
    const int tid = threadIdx.x, g0 = blockIdx.x * K::SEG;
    const size_t row = blockIdx.y, pz = blockIdx.z;
    const int c = (INV ? HLEN / 2 : HLEN / 2 - 1) * f;
    const float* p1 = in1 + pz * s_1 + row * Nc;
    const float* p2 = INV ? in2 + pz * s_2 + row * Nc : nullptr;
    pdl_wait();
    {
        const int xs = g0 - c;
        const int u_lo = max(0, -xs), u_hi = min(ns, Nc - xs);   // no wrap needed inside [u_lo, u_hi)
        for (int u = tid; u < ns; u += kRowThreads) {
            int x = xs + u;
            if (u < u_lo || u >= u_hi) x = rows_wrap1(x, Nc);
            rows_cp4(S1 + u, p1 + x);
            if (INV) rows_cp4(S2 + u, p2 + x);
        }
    }
    rows_cp_wait();
    __syncthreads();
    pdl_launch_dependents();
    const int g = g0 + 4 * tid;
    if (g >= Nc) return;
    float r1[4] = {0.f, 0.f, 0.f, 0.f}, r2[4] = {0.f, 0.f, 0.f, 0.f};
    auto tap = [&](const int j, const float (&v1)[4], const float (&v2)[4]) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (INV) {   // res += v * k / 2: round(v*k), exact halving, add (separable.cu:621-622)
                r1[i] = __fadd_rn(r1[i], __fmul_rn(v1[i], t.IL[HLEN - 1 - j]) * 0.5f);
                r2[i] = __fadd_rn(r2[i], __fmul_rn(v2[i], t.IH[HLEN - 1 - j]) * 0.5f);
            } else {
                r1[i] = fmaf(v1[i], t.L[HLEN - 1 - j], r1[i]);
                r2[i] = fmaf(v1[i], t.H[HLEN - 1 - j], r2[i]);
            }
        }
    };
    if (FMODE == 4) {
#pragma unroll
        for (int j = 0; j < HLEN; j++) {
            const float4 a = *reinterpret_cast<const float4*>(S1 + 4 * tid + j * f);
            const float4 b = INV ? *reinterpret_cast<const float4*>(S2 + 4 * tid + j * f) : a;
            const float v1[4] = {a.x, a.y, a.z, a.w}, v2[4] = {b.x, b.y, b.z, b.w};
            tap(j, v1, v2);
        }
    } else {
        float w1[4 * K::NV], w2[4 * K::NV];
#pragma unroll
        for (int i = 0; i < K::NV; i++) {
            const float4 a = *reinterpret_cast<const float4*>(S1 + 4 * tid + 4 * i);
            const float4 b = INV ? *reinterpret_cast<const float4*>(S2 + 4 * tid + 4 * i) : a;
            w1[4 * i] = a.x; w1[4 * i + 1] = a.y; w1[4 * i + 2] = a.z; w1[4 * i + 3] = a.w;
            w2[4 * i] = b.x; w2[4 * i + 1] = b.y; w2[4 * i + 2] = b.z; w2[4 * i + 3] = b.w;
        }
#pragma unroll
        for (int j = 0; j < HLEN; j++) {
            const float v1[4] = {w1[FMODE * j], w1[FMODE * j + 1], w1[FMODE * j + 2], w1[FMODE * j + 3]};
            const float v2[4] = {w2[FMODE * j], w2[FMODE * j + 1], w2[FMODE * j + 2], w2[FMODE * j + 3]};
            tap(j, v1, v2);
        }
    }
    if (INV) {
        const float o[4] = {__fadd_rn(r1[0], r2[0]), __fadd_rn(r1[1], r2[1]), __fadd_rn(r1[2], r2[2]), __fadd_rn(r1[3], r2[3])};
        rows_store4(out1 + pz * s_o1 + row * Nc, o, g, Nc, vec);
    } else {
        rows_store4(out1 + pz * s_o1 + row * Nc, r1, g, Nc, vec);
        rows_store4(out2 + pz * s_o2 + row * Nc, r2, g, Nc, vec);
    }
}

// ================================================================================================ launchers
static bool rows_enabled()
{
    const char* e = getenv("PDWT_FORCE_GENERIC");
    if (e && *e && *e != '0') return false;
    e = getenv("PDWT_PATH");
    return !(e && !strcmp(e, "generic"));
}
static bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

template <int HLEN>
static int launch_rows_fwd(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int batch, cudaStream_t s)
{
    using K = RowFwdCfg<HLEN>;
    const int n = half_up(Nc);
    const int vec = !(n & 3) && !(lo.stride & 3) && !(hi.stride & 3) && al16(lo.p) && al16(hi.p);
    PDWT_PROF(prof_tag("k_rows_dwt_fwd", Nr, Nc), s);
    PDWT_CUDA(launch_pdl(k_rows_dwt_fwd<HLEN>, dim3(idiv_up(n, K::SEG), Nr, batch), kRowThreads, 0, s, t, (const float*)img.p,
                         img.stride, lo.p, lo.stride, hi.p, hi.stride, Nc, n, vec));
    PDWT_LAUNCH_CHECK();
    return 1;
}
template <int HLEN>
static int launch_rows_inv(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int n, int M, int batch, cudaStream_t s)
{
    using K = RowInvCfg<HLEN>;
    if (n < K::WIN) return 0;   // the single wrap must suffice
    const int vec = !(M & 3) && !(img.stride & 3) && al16(img.p);
    PDWT_PROF(prof_tag("k_rows_dwt_inv", Nr, M), s);
    PDWT_CUDA(launch_pdl(k_rows_dwt_inv<HLEN>, dim3(idiv_up(idiv_up(M, 2), K::SEGC), Nr, batch), kRowThreads, 0, s, t,
                         (const float*)t1.p, t1.stride, (const float*)t2.p, t2.stride, img.p, img.stride, n, M, vec));
    PDWT_LAUNCH_CHECK();
    return 1;
}
template <int HLEN, bool INV>
static int launch_rows_swt(const Taps& t, Plane2 in1, Plane2 in2, Plane2 out1, Plane2 out2, int Nr, int Nc, int level,
                           int batch, cudaStream_t s)
{
    const int f = 1 << (level - 1);
    const size_t smem = rows_swt_smem(HLEN, f, INV ? 2 : 1);
    if (smem > 160 * 1024 || (HLEN - 1) * f >= Nc) return 0;
    const int vec = !(Nc & 3) && !(out1.stride & 3) && al16(out1.p) && (INV || (!(out2.stride & 3) && al16(out2.p)));
    const dim3 grid(idiv_up(Nc, 4 * kRowThreads), Nr, batch);
    PDWT_PROF(prof_tag(INV ? "k_rows_swt_inv" : "k_rows_swt_fwd", Nr, f), s);
#define PDWT_ROWS_SWT_LAUNCH(FM)                                                                                          \
    do {                                                                                                                  \
        PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_rows_swt<HLEN, FM, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
        PDWT_CUDA(launch_pdl(k_rows_swt<HLEN, FM, INV>, grid, kRowThreads, smem, s, t, (const float*)in1.p, in1.stride,  \
                             (const float*)in2.p, in2.stride, out1.p, out1.stride, out2.p, out2.stride, Nc, f, vec));    \
    } while (0)
    if (f == 1)
        PDWT_ROWS_SWT_LAUNCH(1);
    else if (f == 2)
        PDWT_ROWS_SWT_LAUNCH(2);
    else
        PDWT_ROWS_SWT_LAUNCH(4);
#undef PDWT_ROWS_SWT_LAUNCH
    PDWT_LAUNCH_CHECK();
    return 1;
}

#define PDWT_ROWS_HLEN_SWITCH(fn, ...)              \
    switch (t.hlen) {                               \
        case 2: return fn<2>(__VA_ARGS__);          \
        case 4: return fn<4>(__VA_ARGS__);          \
        case 6: return fn<6>(__VA_ARGS__);          \
        case 8: return fn<8>(__VA_ARGS__);          \
        case 10: return fn<10>(__VA_ARGS__);        \
        case 12: return fn<12>(__VA_ARGS__);        \
        case 14: return fn<14>(__VA_ARGS__);        \
        case 16: return fn<16>(__VA_ARGS__);        \
        case 18: return fn<18>(__VA_ARGS__);        \
        case 20: return fn<20>(__VA_ARGS__);        \
        default: return 0;                          \
    }
template <int HLEN>
static int launch_rows_swt_fwd(const Taps& t, Plane2 a, Plane2 b, Plane2 c, Plane2 d, int Nr, int Nc, int level, int batch,
                               cudaStream_t s)
{
    return launch_rows_swt<HLEN, false>(t, a, b, c, d, Nr, Nc, level, batch, s);
}
template <int HLEN>
static int launch_rows_swt_inv(const Taps& t, Plane2 a, Plane2 b, Plane2 c, Plane2 d, int Nr, int Nc, int level, int batch,
                               cudaStream_t s)
{
    return launch_rows_swt<HLEN, true>(t, a, b, c, d, Nr, Nc, level, batch, s);
}

// 1 = handled, 0 = not covered (the caller launches the generic kernel), < 0 = error
int r_fwd_rows(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int batch, cudaStream_t s)
{
    if (!rows_enabled() || Nr > 65535 || batch > 65535 || Nc < t.hlen || t.hlen < 4) return 0;
    PDWT_ROWS_HLEN_SWITCH(launch_rows_fwd, t, img, lo, hi, Nr, Nc, batch, s)
}
int r_inv_rows(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int n, int M, int batch, cudaStream_t s)
{
    if (!rows_enabled() || Nr > 65535 || batch > 65535 || t.hlen < 4) return 0;
    PDWT_ROWS_HLEN_SWITCH(launch_rows_inv, t, t1, t2, img, Nr, n, M, batch, s)
}
int r_swt_fwd_rows(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int level, int batch, cudaStream_t s)
{
    if (!rows_enabled() || Nr > 65535 || batch > 65535 || level < 1 || level > 16) return 0;
    PDWT_ROWS_HLEN_SWITCH(launch_rows_swt_fwd, t, img, img, lo, hi, Nr, Nc, level, batch, s)
}
int r_swt_inv_rows(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int Nc, int level, int batch, cudaStream_t s)
{
    if (!rows_enabled() || Nr > 65535 || batch > 65535 || level < 1 || level > 16) return 0;
    PDWT_ROWS_HLEN_SWITCH(launch_rows_swt_inv, t, t1, t2, img, img, Nr, Nc, level, batch, s)
}

}  // namespace pdwt
