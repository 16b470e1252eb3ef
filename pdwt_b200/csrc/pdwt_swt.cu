// pdwt_swt.cu -- fused per-level kernels of the separable 2-D stationary (undecimated) wavelet transform for sm_100a
// (BASELINE config C3 = sym8, 4 levels, 2048^2).
//
// The reference runs two kernels per level through two full-size scratch planes (separable.cu:496-515, 629-649) and
// every thread gathers its hlen dilated taps from global memory; at level l the taps are f = 2^(l-1) samples apart, so
// neighbouring threads share nothing and the passes run at L2 speed.  Here one kernel per level does both passes on a
// shared-memory tile.  The trick is the tile's row set: a CTA takes rows of ONE residue class modulo f
// (g = ry + f*m), so along y the dilation disappears inside the tile -- TH outputs need TH + hlen - 1 staged rows
// instead of TH + (hlen-1)*f.  Along x the tile is a contiguous run of columns (coalesced global accesses) plus the
// dilated halo of (hlen-1)*f columns.
//
//   forward :  S_in[R][TW + (hlen-1) f]  --row pass-->  S_lo, S_hi [R][TW]  --column pass-->  A, H, V, D
//   inverse :  {A,H} then {V,D} tiles [R][TW + (hlen-1) f]  --column pass-->  S_t1, S_t2 [TH][TW + (hlen-1) f]
//              --row pass-->  image
//
// Arithmetic is the reference's, bit for bit: forward = one fmaf chain from 0 over ascending j
// (separable.cu:427-445, 470-489); inverse = round(v*k), exact halving, add, the two branch sums added last
// (separable.cu:575-588, 615-625).  Index folding is the reference's single +-N wrap (separable.cu:428-433).
// Shapes the tiles cannot serve (odd hlen, hlen > 20, dilations whose halo does not fit in shared memory) return 0 and
// the caller falls back to the generic two-pass kernels.
#include "pdwt_common.cuh"

namespace pdwt {

constexpr int kSwtThreads = 256;
constexpr size_t kSwtSmemCap = 200 * 1024;

// 4-byte asynchronous copy global -> shared: the whole tile is in flight at once, no register staging
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all0() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

typedef unsigned long long u64;
__device__ __forceinline__ u64 sw_pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void sw_unpack2(u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// two IEEE fp32 operations per instruction (FFMA2 / FMUL2 / FADD2): packing changes no bit
__device__ __forceinline__ u64 sw_ffma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 sw_fmul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sw_fadd2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ int wrap1(int i, int N)   // fold_swt as a function of the unwrapped index, then clamped
{
    i += (i < 0) ? N : 0;
    i -= (i >= N) ? N : 0;
    return min(max(i, 0), N - 1);   // only indices of masked outputs can still be outside
}

// ================================================================================================== forward
template <int HLEN>
struct SwtFwdCfg {
    static constexpr int TW = 64, TH = 32;
    static constexpr int C = HLEN / 2 - 1;          // centre_fwd for even hlen (separable.cu:417-421), in units of f
    static constexpr int R = TH + HLEN - 1;         // staged rows (one residue class)
    static constexpr int RPT = 4;                   // output rows per column-pass task
    static size_t smem(int f) { return sizeof(float) * ((size_t)R * (TW + (HLEN - 1) * f) + 2 * (size_t)R * TW); }
};

// F = compile-time dilation (tap offsets become immediates), 0 = run-time `f_rt`
template <int HLEN, int F>
__global__ void __launch_bounds__(kSwtThreads, 2)
    k_swt_fwd_fused(const __grid_constant__ Taps t, const float* __restrict__ src, size_t s_src, float* __restrict__ A,
                    size_t s_a, float* __restrict__ H, float* __restrict__ V, float* __restrict__ D, size_t s_d, int Nr,
                    int Nc, int f_rt)
{
    using K = SwtFwdCfg<HLEN>;
    const int f = F ? F : f_rt;
    extern __shared__ __align__(16) float smem[];
    const int pitch = K::TW + (HLEN - 1) * f;
    float* S_in = smem;
    float* S_lo = smem + K::R * pitch;
    float* S_hi = S_lo + K::R * K::TW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ry = blockIdx.y % f, mt = blockIdx.y / f;
    const int gx0 = blockIdx.x * K::TW;
    src += (size_t)blockIdx.z * s_src;
    pdl_wait();

    // ---- stage the tile: row t of the tile is image row ry + f*(mt*TH + t - C), column u is gx0 - C*f + u
    for (int r = warp; r < K::R; r += kSwtThreads / 32) {
        const int gy = wrap1(ry + f * (mt * K::TH + r - K::C), Nr);
        const float* row = src + (size_t)gy * Nc;
        for (int u = lane; u < pitch; u += 32) cp_async4(S_in + r * pitch + u, row + wrap1(gx0 - K::C * f + u, Nc));
    }
    cp_async_wait_all0();
    __syncthreads();

    // ---- row pass, w_kern_forward_swt_pass1 (separable.cu:409-448): lo/hi[x] = sum_j in[x + (j - C) f] * L/H[hlen-1-j]
    for (int r = warp; r < K::R; r += kSwtThreads / 32) {
#pragma unroll
        for (int xx = lane; xx < K::TW; xx += 32) {
            const float* p = S_in + r * pitch + xx;
            u64 lh = 0ull;   // (lo, hi) += v * (L, H)[hlen-1-j]
#pragma unroll
            for (int j = 0; j < HLEN; j++) {
                const float v = p[j * f];
                lh = sw_ffma2(sw_pack2(v, v), sw_pack2(t.L[HLEN - 1 - j], t.H[HLEN - 1 - j]), lh);
            }
            float lo, hi;
            sw_unpack2(lh, lo, hi);
            S_lo[r * K::TW + xx] = lo;
            S_hi[r * K::TW + xx] = hi;
        }
    }
    __syncthreads();
    pdl_launch_dependents();

    // ---- column pass, w_kern_forward_swt_pass2 (separable.cu:452-493): inside the residue class the taps are adjacent
    // tile rows.  A task = one column x RPT consecutive tile rows, 4 sub-bands.
    const int xx = tid % K::TW;
    const int gx = gx0 + xx;
    for (int m0 = (tid / K::TW) * K::RPT; m0 < K::TH; m0 += (kSwtThreads / K::TW) * K::RPT) {
        u64 ah[K::RPT], vd[K::RPT];   // (A, H) += lo * (L, H)[..],  (V, D) += hi * (L, H)[..]
#pragma unroll
        for (int o = 0; o < K::RPT; o++) ah[o] = vd[o] = 0ull;
#pragma unroll
        for (int i = 0; i < K::RPT + HLEN - 1; i++) {
            const float v1 = S_lo[(m0 + i) * K::TW + xx], v2 = S_hi[(m0 + i) * K::TW + xx];
            const u64 p1 = sw_pack2(v1, v1), p2 = sw_pack2(v2, v2);
#pragma unroll
            for (int o = 0; o < K::RPT; o++) {
                const int j = i - o;
                if (j >= 0 && j < HLEN) {
                    const u64 k = sw_pack2(t.L[HLEN - 1 - j], t.H[HLEN - 1 - j]);
                    ah[o] = sw_ffma2(p1, k, ah[o]);
                    vd[o] = sw_ffma2(p2, k, vd[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < K::RPT; o++) {
            const int gy = ry + f * (mt * K::TH + m0 + o);
            if (gy < Nr && gx < Nc) {
                const size_t off = (size_t)gy * Nc + gx;
                float a, h, v, d;
                sw_unpack2(ah[o], a, h);
                sw_unpack2(vd[o], v, d);
                A[(size_t)blockIdx.z * s_a + off] = a;
                H[(size_t)blockIdx.z * s_d + off] = h;
                V[(size_t)blockIdx.z * s_d + off] = v;
                D[(size_t)blockIdx.z * s_d + off] = d;
            }
        }
    }
}

// ================================================================================================== inverse
template <int HLEN>
struct SwtInvCfg {
    static constexpr int TW = 128, TH = 16;
    static constexpr int C = HLEN / 2;              // inverse centre (separable.cu:564-569), in units of f
    static constexpr int R = TH + HLEN - 1;
    static constexpr int RPT = 4;
    // two coefficient tiles + t1 + t2, all CI = TW + (hlen-1) f columns wide
    static size_t smem(int f) { return sizeof(float) * (2 * (size_t)R + 2 * (size_t)TH) * (TW + (HLEN - 1) * f); }
};

// One synthesis tap of the reference: res += v * k / 2 = round(v*k), exact halving, add (separable.cu:581-584).  Halving
// commutes with the rounding of the product (a power-of-two scaling, barring underflow into the denormals, i.e. for
// |v*k| >= 2^-125), so round(v * (k/2)) is the same number and the tap is one multiply and one add -- NOT an FMA, the
// product is rounded on its own.  Two arrays share one FMUL2, (pl, ph) = (v0, v1) * (IL/2, IH/2), followed by two scalar
// add.rn (ptxas contracts a packed mul.rn.f32x2 + add.rn.f32x2 pair into one FFMA2 even with --fmad=false, which would
// round once: checked in the SASS).
template <int HLEN>
struct SwtHalfTaps {
    float2 k[HLEN];   // (IL, IH)[hlen-1-j] / 2, exact
};
template <int HLEN, int F>
__global__ void __launch_bounds__(kSwtThreads, 2)
    k_swt_inv_fused(const __grid_constant__ SwtHalfTaps<HLEN> t, const float* __restrict__ A, size_t s_a, const float* __restrict__ H,
                    const float* __restrict__ V, const float* __restrict__ D, size_t s_d, float* __restrict__ dst,
                    size_t s_dst, int Nr, int Nc, int f_rt)
{
    using K = SwtInvCfg<HLEN>;
    const int f = F ? F : f_rt;
    auto khalf = [&](const int j) { return sw_pack2(t.k[j].x, t.k[j].y); };
    extern __shared__ __align__(16) float smem[];
    const int ci = K::TW + (HLEN - 1) * f;           // columns of t1/t2 the row pass of this tile reads
    float* S_c0 = smem;                              // A, then V
    float* S_c1 = S_c0 + K::R * ci;                  // H, then D
    float* S_t = S_c1 + K::R * ci;                   // [2][TH][ci]: t1, t2
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ry = blockIdx.y % f, mt = blockIdx.y / f;
    const int gx0 = blockIdx.x * K::TW;
    pdl_wait();

#pragma unroll 1
    for (int pair = 0; pair < 2; pair++) {
        const float* c0 = (pair ? V + (size_t)blockIdx.z * s_d : A + (size_t)blockIdx.z * s_a);
        const float* c1 = (pair ? D : H) + (size_t)blockIdx.z * s_d;
        if (pair) __syncthreads();   // the column pass of the first pair has finished reading S_c0 / S_c1
        for (int r = warp; r < K::R; r += kSwtThreads / 32) {
            const size_t rowo = (size_t)wrap1(ry + f * (mt * K::TH + r - K::C), Nr) * Nc;
            for (int u = lane; u < ci; u += 32) {
                const size_t o = rowo + wrap1(gx0 - K::C * f + u, Nc);
                cp_async4(S_c0 + r * ci + u, c0 + o);
                cp_async4(S_c1 + r * ci + u, c1 + o);
            }
        }
        cp_async_wait_all0();
        __syncthreads();
        // column pass, w_kern_inverse_swt_pass1 (separable.cu:553-589): t = IL_y(c0) + IH_y(c1), RPT rows per task
        float* S_out = S_t + pair * K::TH * ci;
        const int ncb = (ci + 31) / 32;
        for (int task = warp; task < ncb * (K::TH / K::RPT); task += kSwtThreads / 32) {
            const int u = (task % ncb) * 32 + lane, m0 = (task / ncb) * K::RPT;
            if (u < ci) {
                float rl[K::RPT], rh[K::RPT];
#pragma unroll
                for (int o = 0; o < K::RPT; o++) rl[o] = rh[o] = 0.f;
#pragma unroll
                for (int i = 0; i < K::RPT + HLEN - 1; i++) {
                    const u64 v01 = sw_pack2(S_c0[(m0 + i) * ci + u], S_c1[(m0 + i) * ci + u]);
#pragma unroll
                    for (int o = 0; o < K::RPT; o++) {
                        const int j = i - o;
                        if (j >= 0 && j < HLEN) {
                            float pl, ph;   // the two products rounded on their own (one FMUL2), then two plain adds
                            sw_unpack2(sw_fmul2(v01, khalf(j)), pl, ph);
                            rl[o] = __fadd_rn(rl[o], pl);
                            rh[o] = __fadd_rn(rh[o], ph);
                        }
                    }
                }
#pragma unroll
                for (int o = 0; o < K::RPT; o++) S_out[(m0 + o) * ci + u] = __fadd_rn(rl[o], rh[o]);
            }
        }
    }
    __syncthreads();
    pdl_launch_dependents();

    // ---- row pass, w_kern_inverse_swt_pass2 (separable.cu:593-626): img[x] = IL_x(t1) + IH_x(t2), taps f apart
    dst += (size_t)blockIdx.z * s_dst;
    const float* S_t1 = S_t;
    const float* S_t2 = S_t + K::TH * ci;
    const int xx = tid % K::TW;
    const int gx = gx0 + xx;
    for (int m = tid / K::TW; m < K::TH; m += kSwtThreads / K::TW) {
        const float* p1 = S_t1 + m * ci + xx;
        const float* p2 = S_t2 + m * ci + xx;
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int j = 0; j < HLEN; j++) {
            float q1, q2;
            sw_unpack2(sw_fmul2(sw_pack2(p1[j * f], p2[j * f]), khalf(j)), q1, q2);
            a1 = __fadd_rn(a1, q1);
            a2 = __fadd_rn(a2, q2);
        }
        const int gy = ry + f * (mt * K::TH + m);
        if (gy < Nr && gx < Nc) dst[(size_t)gy * Nc + gx] = __fadd_rn(a1, a2);
    }
}

// ================================================================================================ launchers
template <int HLEN>
static int launch_swt_fwd(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                          int batch, cudaStream_t s)
{
    using K = SwtFwdCfg<HLEN>;
    const int f = 1 << (level - 1);
    const size_t smem = K::smem(f);
    // the reference's single wrap must suffice (it does for every level the Wavelets class allows) and the tile must fit
    if (smem > kSwtSmemCap || (HLEN - 1) * f >= Nr || (HLEN - 1) * f >= Nc) return 0;
    const int rows_per_class = idiv_up(Nr, f);
    dim3 grid(idiv_up(Nc, K::TW), f * idiv_up(rows_per_class, K::TH), batch);
    if (grid.y > 65535u) return 0;
    PDWT_PROF(prof_tag("k_swt_fwd_fused", Nr, f), s);
#define PDWT_SWT_FWD(FF)                                                                                                 \
    do {                                                                                                                 \
        static PerDeviceOnce once;                                                                                       \
        if (once.first()) {                                                                                              \
            PDWT_CUDA(cudaFuncSetAttribute(k_swt_fwd_fused<HLEN, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                           (int)kSwtSmemCap));                                                           \
        }                                                                                                                \
        PDWT_CUDA(launch_pdl(k_swt_fwd_fused<HLEN, FF>, grid, kSwtThreads, smem, s, t, (const float*)src.p, src.stride,  \
                             A.p, A.stride, H.p, V.p, D.p, H.stride, Nr, Nc, f));                                        \
    } while (0)
    switch (f) {
        case 1: PDWT_SWT_FWD(1); break;
        case 2: PDWT_SWT_FWD(2); break;
        case 4: PDWT_SWT_FWD(4); break;
        case 8: PDWT_SWT_FWD(8); break;
        default: PDWT_SWT_FWD(0); break;
    }
#undef PDWT_SWT_FWD
    PDWT_LAUNCH_CHECK();
    return 1;
}

template <int HLEN>
static int launch_swt_inv(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int Nr, int Nc, int level,
                          int batch, cudaStream_t s)
{
    using K = SwtInvCfg<HLEN>;
    const int f = 1 << (level - 1);
    const size_t smem = K::smem(f);
    if (smem > kSwtSmemCap || (HLEN - 1) * f >= Nr || (HLEN - 1) * f >= Nc) return 0;
    const int rows_per_class = idiv_up(Nr, f);
    dim3 grid(idiv_up(Nc, K::TW), f * idiv_up(rows_per_class, K::TH), batch);
    if (grid.y > 65535u) return 0;
    PDWT_PROF(prof_tag("k_swt_inv_fused", Nr, f), s);
    SwtHalfTaps<HLEN> ht;
    for (int j = 0; j < HLEN; j++) ht.k[j] = make_float2(t.IL[HLEN - 1 - j] * 0.5f, t.IH[HLEN - 1 - j] * 0.5f);
#define PDWT_SWT_INV(FF)                                                                                                 \
    do {                                                                                                                 \
        static PerDeviceOnce once;                                                                                       \
        if (once.first()) {                                                                                              \
            PDWT_CUDA(cudaFuncSetAttribute(k_swt_inv_fused<HLEN, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                           (int)kSwtSmemCap));                                                           \
        }                                                                                                                \
        PDWT_CUDA(launch_pdl(k_swt_inv_fused<HLEN, FF>, grid, kSwtThreads, smem, s, ht, (const float*)A.p, A.stride,     \
                             (const float*)H.p, (const float*)V.p, (const float*)D.p, H.stride, dst.p, dst.stride, Nr,   \
                             Nc, f));                                                                                    \
    } while (0)
    switch (f) {
        case 1: PDWT_SWT_INV(1); break;
        case 2: PDWT_SWT_INV(2); break;
        case 4: PDWT_SWT_INV(4); break;
        case 8: PDWT_SWT_INV(8); break;
        default: PDWT_SWT_INV(0); break;
    }
#undef PDWT_SWT_INV
    PDWT_LAUNCH_CHECK();
    return 1;
}

#define PDWT_SWT_HLEN_SWITCH(fn, ...)               \
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
static int swt_levels_fit(int Nr, int Nc, int nlevels)
{
    const int f = 1 << (nlevels - 1);
    if ((HLEN - 1) * f >= Nr || (HLEN - 1) * f >= Nc) return 0;
    if (SwtFwdCfg<HLEN>::smem(f) > kSwtSmemCap || SwtInvCfg<HLEN>::smem(f) > kSwtSmemCap) return 0;
    if ((long long)f * idiv_up(idiv_up(Nr, f), SwtInvCfg<HLEN>::TH) > 65535) return 0;
    return 1;
}
int w_swt2_supported(const Taps& t, int Nr, int Nc, int nlevels, int batch)
{
    if (batch > 65535 || nlevels < 1 || nlevels > 16) return 0;
    PDWT_SWT_HLEN_SWITCH(swt_levels_fit, Nr, Nc, nlevels)
}

// 1 = handled, 0 = shape not covered (caller uses the generic two-pass kernels), < 0 = error.
// In-place note: the forward of level l reads A_{l-1} and writes A_l; the caller gives distinct planes.
int w_swt2_fwd_level(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                     int batch, cudaStream_t s)
{
    if (batch > 65535 || level < 1 || level > 16) return 0;
    PDWT_SWT_HLEN_SWITCH(launch_swt_fwd, t, src, A, H, V, D, Nr, Nc, level, batch, s)
}

int w_swt2_inv_level(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int Nr, int Nc, int level,
                     int batch, cudaStream_t s)
{
    if (batch > 65535 || level < 1 || level > 16) return 0;
    PDWT_SWT_HLEN_SWITCH(launch_swt_inv, t, A, H, V, D, dst, Nr, Nc, level, batch, s)
}

}  // namespace pdwt
