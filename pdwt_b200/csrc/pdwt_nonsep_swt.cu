// pdwt_nonsep_swt.cu -- tiled kernels of the NON-separable 2-D SWT (undecimated) for sm_100a.
//
// The reference gives every output pixel to one thread that gathers its hlen x hlen taps, 2^(level-1) samples apart,
// from global memory (nonseparable.cu:304-354 forward, 360-401 inverse): 4*hlen^2 multiply-adds per pixel and level
// with one load per tap and no reuse.  The transform is FP32-bound, so -- like the decimated non-separable kernels in
// pdwt_nonsep.cu -- the level kernels here are register-tiled direct convolutions out of shared memory:
//
//   * a CTA owns TH output rows of ONE residue class modulo f = 2^(level-1) (image rows ry + f*m) and TW contiguous
//     columns: inside the tile the dilation along y disappears (TH + hlen - 1 staged rows instead of TH + (hlen-1)*f),
//     along x the tile carries the dilated halo of (hlen-1)*f columns.  The periodic fold of the reference (a single
//     +-N wrap, nonseparable.cu:325-336) is applied while staging, with 4-byte cp.async copies;
//   * a thread owns 4 consecutive output columns.  Forward: per tap it multiplies its 4 samples with the tap's four
//     filter values, packed:  (A,H)[u] += x.F32 * (K_LL, K_LH),  (V,D)[u] += x.F32 * (K_HL, K_HH)  -- 8 FFMA2 per tap.
//     Inverse: the four sub-band tiles are staged interleaved as (A,H) and (V,D) pairs; a synthesis tap is the
//     reference's round(v*K), exact quartering, add (nonseparable.cu:390-393): ONE FMUL2 against the pre-quartered pair
//     (K/4 is exact and commutes with the product's rounding) and ONE FFMA2 with the multiplicand (1,1), which arrives
//     as a kernel parameter so that ptxas cannot contract the two roundings into one (see pdwt_swt.cu);
//   * the 2-D filters travel as a kernel parameter (constant bank -> uniform registers), in accumulation order.  They
//     are the reference's outer products rounded to fp32 (w_outer, nonseparable.cu:16-24) or a custom quadruple.
// Every output is the reference's chain: from 0, taps in (jy, jx) lexicographic order; the inverse adds
// ((ra + rh) + rv) + rd.  Shapes the tiles cannot serve (halo wider than the plane, dilation > 8, hlen > 20) fall back
// to the generic kernels.
#include "pdwt_common.cuh"

namespace pdwt {

namespace {

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
__device__ __forceinline__ void sw_cp4(void* smem_dst, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src)
                 : "memory");
}

constexpr int kThreads = 256;
constexpr int TW = 64, TH = 16, U = 4;   // output tile: 16 thread columns (4 outputs each) x 16 rows of one residue class

template <int HLEN>
struct SwtK {
    float4 k[HLEN][HLEN];   // [jy][jx] -> (LL, LH, HL, HH) in accumulation order (pre-quartered for the inverse)
    float2 one;             // (1, 1): the multiplicand of the inverse's rounding-preserving add
};

// staged columns of a tile, padded to a multiple of 4 plus 4 (rows of consecutive ty land in different banks)
__host__ __device__ constexpr int swt_pitch(int hlen, int f) { return ((TW + (hlen - 1) * f + 3) / 4) * 4 + 4; }

// INV = false: src -> A, H, V, D (w_kern_forward_swt).  INV = true: A, H, V, D -> dst (w_kern_inverse_swt).
template <int HLEN, int FF, bool INV>
__global__ void __launch_bounds__(kThreads, 2)
    k_nonsep_swt_tiled(const __grid_constant__ SwtK<HLEN> kt, const float* __restrict__ img, size_t s_img,
                       const float* __restrict__ A, size_t s_a, const float* __restrict__ H, const float* __restrict__ V,
                       const float* __restrict__ D, size_t s_d, float* __restrict__ o0, float* __restrict__ o1,
                       float* __restrict__ o2, float* __restrict__ o3, size_t s_o0, size_t s_o, int Nr, int Nc, int ntr)
{
    constexpr int PITCH = swt_pitch(HLEN, FF);
    constexpr int INR = TH + HLEN - 1, INW = TW + (HLEN - 1) * FF;
    // forward: analysis centre c = hlen/2 - 1 taps (nonseparable.cu:311-320); inverse: c = hlen/2 taps (:366-371)
    constexpr int CT = INV ? HLEN / 2 : HLEN / 2 - 1;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    // blockIdx.y = residue class ry (rows ry + FF*m) x tile of TH class rows
    const int ry = blockIdx.y / ntr, m0 = (blockIdx.y - ry * ntr) * TH;
    const int gx0 = blockIdx.x * TW;
    const int plane = blockIdx.z;
    pdl_wait();

    // ---- staging with the reference's single +-N wrap (fold_swt) per row and per column
    {
        const int lane = tid & 31;
        const int xs = gx0 - CT * FF;
        for (int r = tid >> 5; r < INR; r += kThreads / 32) {
            int y = ry + FF * (m0 + r - CT);
            y += (y < 0) ? Nr : 0;
            y -= (y >= Nr) ? Nr : 0;
            y = y < 0 ? 0 : (y >= Nr ? Nr - 1 : y);   // rows of class positions past the plane are staged but never used
            for (int u = lane; u < INW; u += 32) {
                int x = xs + u;
                x += (x < 0) ? Nc : 0;
                x -= (x >= Nc) ? Nc : 0;
                x = x < 0 ? 0 : (x >= Nc ? Nc - 1 : x);
                const size_t o = (size_t)y * Nc + x;
                if constexpr (!INV) {
                    sw_cp4(smem + r * PITCH + u, img + (size_t)plane * s_img + o);
                } else {
                    float* dah = smem + 2 * (r * PITCH + u);
                    float* dvd = smem + 2 * (INR * PITCH) + 2 * (r * PITCH + u);
                    sw_cp4(dah, A + (size_t)plane * s_a + o);
                    sw_cp4(dah + 1, H + (size_t)plane * s_d + o);
                    sw_cp4(dvd, V + (size_t)plane * s_d + o);
                    sw_cp4(dvd + 1, D + (size_t)plane * s_d + o);
                }
            }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();

    const int tx = tid & 15, ty = tid >> 4;
    const int gy = ry + FF * (m0 + ty), gx = gx0 + U * tx;
    if (gy >= Nr || gx >= Nc) return;
    u64 acc0[U], acc1[U];   // forward: (A,H), (V,D); inverse: (ra,rh), (rv,rd)
#pragma unroll
    for (int u = 0; u < U; u++) acc0[u] = acc1[u] = 0ull;
    const u64 one = sw_pack2(kt.one.x, kt.one.y);

#pragma unroll 1
    for (int jy = 0; jy < HLEN; jy++) {
        if constexpr (!INV) {
            const float* row = smem + (ty + jy) * PITCH + U * tx;
            if constexpr (FF <= 2) {
                // dense window: 4 + (hlen-1)*FF consecutive samples, every one of them used
                constexpr int NV = (U + (HLEN - 1) * FF + 3) / 4;
                float x[NV * 4];
#pragma unroll
                for (int i = 0; i < NV; i++) {
                    const float4 f = reinterpret_cast<const float4*>(row)[i];
                    x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w;
                }
#pragma unroll
                for (int jx = 0; jx < HLEN; jx++) {
                    const float4 k = kt.k[jy][jx];
                    const u64 k0 = sw_pack2(k.x, k.y), k1 = sw_pack2(k.z, k.w);
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const float v = x[u + jx * FF];
                        acc0[u] = sw_ffma2(sw_pack2(v, v), k0, acc0[u]);
                        acc1[u] = sw_ffma2(sw_pack2(v, v), k1, acc1[u]);
                    }
                }
            } else {
#pragma unroll
                for (int jx = 0; jx < HLEN; jx++) {   // FF % 4 == 0: the four samples of a tap are one aligned vector
                    const float4 f = *reinterpret_cast<const float4*>(row + jx * FF);
                    const float x[4] = {f.x, f.y, f.z, f.w};
                    const float4 k = kt.k[jy][jx];
                    const u64 k0 = sw_pack2(k.x, k.y), k1 = sw_pack2(k.z, k.w);
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        acc0[u] = sw_ffma2(sw_pack2(x[u], x[u]), k0, acc0[u]);
                        acc1[u] = sw_ffma2(sw_pack2(x[u], x[u]), k1, acc1[u]);
                    }
                }
            }
        } else {
            const float2* rah = reinterpret_cast<const float2*>(smem) + (ty + jy) * PITCH + U * tx;
            const float2* rvd = rah + INR * PITCH;
#pragma unroll
            for (int jx = 0; jx < HLEN; jx++) {
                u64 ah[U], vd[U];
                if constexpr ((FF & 1) == 0) {   // even offset: two aligned 16-byte loads per pair tile
#pragma unroll
                    for (int i = 0; i < U / 2; i++) {
                        const float4 f = *reinterpret_cast<const float4*>(rah + jx * FF + 2 * i);
                        const float4 g = *reinterpret_cast<const float4*>(rvd + jx * FF + 2 * i);
                        ah[2 * i] = sw_pack2(f.x, f.y); ah[2 * i + 1] = sw_pack2(f.z, f.w);
                        vd[2 * i] = sw_pack2(g.x, g.y); vd[2 * i + 1] = sw_pack2(g.z, g.w);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const float2 f = rah[jx * FF + u], g = rvd[jx * FF + u];
                        ah[u] = sw_pack2(f.x, f.y);
                        vd[u] = sw_pack2(g.x, g.y);
                    }
                }
                const float4 k = kt.k[jy][jx];   // pre-quartered
                const u64 k0 = sw_pack2(k.x, k.y), k1 = sw_pack2(k.z, k.w);
#pragma unroll
                for (int u = 0; u < U; u++) {
                    acc0[u] = sw_ffma2(sw_fmul2(ah[u], k0), one, acc0[u]);   // round(v*K)/4, then add: two roundings
                    acc1[u] = sw_ffma2(sw_fmul2(vd[u], k1), one, acc1[u]);
                }
            }
        }
    }

    const size_t off = (size_t)gy * Nc + gx;
    if constexpr (!INV) {
        float oa[U], oh[U], ov[U], od[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            sw_unpack2(acc0[u], oa[u], oh[u]);
            sw_unpack2(acc1[u], ov[u], od[u]);
        }
        float* pa = o0 + (size_t)plane * s_o0 + off;
        float* ph = o1 + (size_t)plane * s_o + off;
        float* pv = o2 + (size_t)plane * s_o + off;
        float* pd = o3 + (size_t)plane * s_o + off;
        const bool vec = ((Nc & 3) == 0) && ((s_o0 & 3) == 0) && ((s_o & 3) == 0) &&
                         (((uintptr_t)o0 | (uintptr_t)o1 | (uintptr_t)o2 | (uintptr_t)o3) & 15) == 0;
        if (vec) {
            *reinterpret_cast<float4*>(pa) = make_float4(oa[0], oa[1], oa[2], oa[3]);
            *reinterpret_cast<float4*>(ph) = make_float4(oh[0], oh[1], oh[2], oh[3]);
            *reinterpret_cast<float4*>(pv) = make_float4(ov[0], ov[1], ov[2], ov[3]);
            *reinterpret_cast<float4*>(pd) = make_float4(od[0], od[1], od[2], od[3]);
        } else {
#pragma unroll
            for (int u = 0; u < U; u++)
                if (gx + u < Nc) {
                    pa[u] = oa[u]; ph[u] = oh[u]; pv[u] = ov[u]; pd[u] = od[u];
                }
        }
    } else {
        float r[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            float ra, rh, rv, rd;
            sw_unpack2(acc0[u], ra, rh);
            sw_unpack2(acc1[u], rv, rd);
            r[u] = __fadd_rn(__fadd_rn(__fadd_rn(ra, rh), rv), rd);   // nonseparable.cu:396
        }
        float* po = o0 + (size_t)plane * s_o0 + off;
        if (((Nc & 3) == 0) && ((s_o0 & 3) == 0) && (((uintptr_t)o0) & 15) == 0) {
            *reinterpret_cast<float4*>(po) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
            for (int u = 0; u < U; u++)
                if (gx + u < Nc) po[u] = r[u];
        }
    }
}

template <int HLEN, int FF, bool INV>
int launch_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch, cudaStream_t s)
{
    constexpr int PITCH = swt_pitch(HLEN, FF);
    constexpr size_t smem = sizeof(float) * (INV ? 4 : 1) * (size_t)(TH + HLEN - 1) * PITCH;
    if (smem > 200 * 1024) return 0;
    if ((HLEN - 1) * FF >= Nc || (HLEN - 1) * FF >= Nr) return 0;   // the single wrap must suffice
    PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_nonsep_swt_tiled<HLEN, FF, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem));
    SwtK<HLEN> kt;
    kt.one = make_float2(1.0f, 1.0f);
    const float *f0 = INV ? t.IL : t.L, *f1 = INV ? t.IH : t.H;
    for (int jy = 0; jy < HLEN; jy++)
        for (int jx = 0; jx < HLEN; jx++) {
            const int ki = (HLEN - 1 - jy) * HLEN + (HLEN - 1 - jx);   // nonseparable.cu:339-342 / 390-393
            float4 k;
            if (t.hk2d) {
                k = make_float4(k2d_at(t.hk2d, HLEN, 0, ki), k2d_at(t.hk2d, HLEN, 1, ki), k2d_at(t.hk2d, HLEN, 2, ki),
                                k2d_at(t.hk2d, HLEN, 3, ki));
            } else {   // w_outer: one fp32 multiplication each, no contraction, no excess precision
                const volatile float ly = f0[HLEN - 1 - jy], hy = f1[HLEN - 1 - jy];
                const volatile float lx = f0[HLEN - 1 - jx], hx = f1[HLEN - 1 - jx];
                volatile float ll = ly * lx, lh = ly * hx, hl = hy * lx, hh = hy * hx;
                k = make_float4(ll, lh, hl, hh);
            }
            if (INV) k = make_float4(k.x * 0.25f, k.y * 0.25f, k.z * 0.25f, k.w * 0.25f);   // exact
            kt.k[jy][jx] = k;
        }
    const int nclass = idiv_up(Nr, FF);          // rows of the longest residue class
    const int ntr = idiv_up(nclass, TH);
    dim3 grid(idiv_up(Nc, TW), FF * ntr, batch);
    if (grid.y > 65535u) return 0;
    PDWT_PROF(prof_tag(INV ? "k_nonsep_swt_inv_tiled" : "k_nonsep_swt_fwd_tiled", Nr, FF), s);
    if (INV)
        PDWT_CUDA(launch_pdl(k_nonsep_swt_tiled<HLEN, FF, INV>, grid, kThreads, smem, s, kt, (const float*)nullptr, (size_t)0,
                             (const float*)A.p, A.stride, (const float*)H.p, (const float*)V.p, (const float*)D.p, H.stride,
                             img.p, (float*)nullptr, (float*)nullptr, (float*)nullptr, img.stride, (size_t)0, Nr, Nc, ntr));
    else
        PDWT_CUDA(launch_pdl(k_nonsep_swt_tiled<HLEN, FF, INV>, grid, kThreads, smem, s, kt, (const float*)img.p, img.stride,
                             (const float*)nullptr, (size_t)0, (const float*)nullptr, (const float*)nullptr,
                             (const float*)nullptr, (size_t)0, A.p, H.p, V.p, D.p, A.stride, H.stride, Nr, Nc, ntr));
    PDWT_LAUNCH_CHECK();
    return 1;
}

template <int HLEN, bool INV>
int by_dilation(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level, int batch,
                cudaStream_t s)
{
    switch (level) {
        case 1: return launch_level<HLEN, 1, INV>(t, img, A, H, V, D, Nr, Nc, batch, s);
        case 2: return launch_level<HLEN, 2, INV>(t, img, A, H, V, D, Nr, Nc, batch, s);
        case 3: return launch_level<HLEN, 4, INV>(t, img, A, H, V, D, Nr, Nc, batch, s);
        case 4: return launch_level<HLEN, 8, INV>(t, img, A, H, V, D, Nr, Nc, batch, s);
        default: return 0;
    }
}

}  // namespace

#define PDWT_NSSWT_HLEN_SWITCH(INV)                                                                     \
    switch (t.hlen) {                                                                                   \
        case 2: return by_dilation<2, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);                \
        case 4: return by_dilation<4, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);                \
        case 6: return by_dilation<6, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);                \
        case 8: return by_dilation<8, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);                \
        case 10: return by_dilation<10, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);              \
        case 12: return by_dilation<12, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);              \
        case 14: return by_dilation<14, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);              \
        case 16: return by_dilation<16, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);              \
        case 18: return by_dilation<18, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);              \
        case 20: return by_dilation<20, INV>(t, img, A, H, V, D, Nr, Nc, level, batch, s);              \
        default: return 0;                                                                              \
    }

// 1 = handled, 0 = shape / filter length / dilation not covered (the caller uses the generic kernel), < 0 = error
int n_nonsep_swt_fwd_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                           int batch, cudaStream_t s)
{
    PDWT_NSSWT_HLEN_SWITCH(false)
}
int n_nonsep_swt_inv_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                           int batch, cudaStream_t s)
{
    PDWT_NSSWT_HLEN_SWITCH(true)
}

}  // namespace pdwt
