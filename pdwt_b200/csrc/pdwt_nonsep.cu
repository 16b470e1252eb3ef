// pdwt_nonsep.cu -- tiled kernels of the NON-separable 2-D DWT for sm_100a (BASELINE config C4).
//
// The reference gives every output pixel to one thread that walks its hlen x hlen taps through global memory and four
// constant-memory filters (nonseparable.cu:114-225): 4*hlen^2 FMAs per output with one load per FMA quartet and no
// reuse.  The transform is FP32-bound (980 flop/pixel for db7, 2 levels), so the kernels here are register-tiled direct
// convolutions out of shared memory:
//
//   forward : a thread owns 4 consecutive output columns x 4 filters.  Per tap row it pulls its 2*3+hlen input samples
//             with 128-bit shared loads, per tap it reads the four filter values as ONE broadcast 128-bit load and
//             issues 8 FFMA2:  (A,H)[u] += x.F32 * (K_LL, K_LH),  (V,D)[u] += x.F32 * (K_HL, K_HH).
//   inverse : the four sub-band tiles are staged interleaved as (A,H) and (V,D) pairs; a thread owns 2 coefficient
//             positions = a 2 x 4 pixel block.  Per window position it loads the two pairs once and feeds every output
//             parity whose tap lands there:  (ra,rh) += (a,h) * (K_LL', K_LH'),  (rv,rd) += (v,d) * (K_HL', K_HH').
//
// The 2-D filters are the reference's: outer products of the 1-D banks ROUNDED to fp32 (w_outer, nonseparable.cu:16-24,
// 71-74; note the reference's H/V swap w.r.t. separable mode, SURVEY B2, which comes with them), formed once per CTA in
// shared memory with __fmul_rn.  Every output is the reference's chain: from 0, taps in (jy, jx) lexicographic order, one
// fmaf each (fma.rn.f32x2 = two IEEE fmaf); the inverse adds ((ra + rh) + rv) + rd (nonseparable.cu:222-223).
// Any plane size (the periodic / odd-size folds of nonseparable.cu:139-152 and 205-214 are applied while staging).
#include <type_traits>

#include "pdwt_common.cuh"

namespace pdwt {

typedef unsigned long long u64;
__device__ __forceinline__ u64 ns_pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void ns_unpack2(u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ns_ffma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ int ns_clamp(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

constexpr int kNsThreads = 256;

// 4-byte asynchronous copy global -> shared: no register staging, the whole tile in flight at once
__device__ __forceinline__ void ns_cp4(float* smem_dst, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src)
                 : "memory");
}

// ================================================================================================== forward
template <int HLEN>
struct NsFwdCfg {
    static constexpr int U = 4;                      // output columns per thread
    static constexpr int TW = 64, TH = 16;           // output tile: 16 thread columns x 16 rows
    static constexpr int C = HLEN / 2 - 1;           // analysis centre, even hlen (nonseparable.cu:121-130)
    static constexpr int INR = 2 * TH + HLEN - 2;
    static constexpr int INW = 2 * TW + HLEN - 2;
    static constexpr int NV = (2 * (U - 1) + HLEN + 3) / 4;   // 16-byte vectors a thread reads per tap row
    static constexpr int PITCH = ((2 * (TW - U) + 4 * NV + 7) / 8) * 8 + 8;   // floats; >= INW, multiple of 8
    // A thread's window starts at 16-byte vector 2 tx of its row: read in place, the lanes of a warp would be 32 bytes
    // apart (half the shared-memory bandwidth).  So a row is stored with its EVEN vectors first and its odd vectors
    // behind them (HALF vectors each): window vector i of lane tx is vector (i & 1) * HALF + tx + (i >> 1), consecutive
    // lanes read consecutive 16 bytes.
    static constexpr int HALF = PITCH / 8;
    static constexpr int MAXRG = 2;                  // a CTA stages once for up to MAXRG groups of TH output rows
    static constexpr size_t smem(int rg) { return sizeof(float) * ((size_t)(2 * TH * rg + HLEN - 2) * PITCH + 4 * HLEN * HLEN); }
    static_assert(PITCH >= INW && 2 * (TW / U - 1) + NV <= PITCH / 4, "tile pitch");
    __host__ __device__ static constexpr int slot(int u) { return ((((u >> 2) & 1) * HALF + (u >> 3)) << 2) + (u & 3); }
};

// KC (the default; PDWT_NS_FWD_CONST=0 switches it off): the filter table arrives as a
// kernel parameter and is read through the uniform datapath like the inverse's (see NsInvK) instead of from shared memory.
template <int HLEN>
struct NsFwdK {
    float4 k[HLEN][HLEN];   // [jy][jx] -> (LL, LH, HL, HH)
};
template <int HLEN, bool KC>
__global__ void __launch_bounds__(kNsThreads, 2)
    k_nonsep_fwd_tiled(const __grid_constant__ std::conditional_t<KC, NsFwdK<HLEN>, Taps> t, const float* __restrict__ img, size_t s_img, float* __restrict__ A,
                       size_t s_a, float* __restrict__ H, float* __restrict__ V, float* __restrict__ D, size_t s_d, int Nr,
                       int Nc, int rg)   // rg: groups of TH output rows per CTA
{
    using K = NsFwdCfg<HLEN>;
    extern __shared__ __align__(16) float smem[];
    const int inr = 2 * K::TH * rg + HLEN - 2;   // staged rows: the halo is paid once for rg row groups
    float* S_in = smem;
    float4* S_k = reinterpret_cast<float4*>(smem + inr * K::PITCH);   // [jy][jx] -> (LL, LH, HL, HH)
    const int tid = threadIdx.x;
    const int nr = half_up(Nr), nc = half_up(Nc);
    const int gx0 = blockIdx.x * K::TW, gy0 = blockIdx.y * (K::TH * rg);
    img += (size_t)blockIdx.z * s_img;

    // the four 2-D filters in accumulation order: tap (jy, jx) multiplies K[hlen-1-jy][hlen-1-jx] (nonseparable.cu:155-160)
    if constexpr (!KC) {
        for (int i = tid; i < HLEN * HLEN; i += kNsThreads) {
            const int jy = i / HLEN, jx = i - jy * HLEN;
            const float ly = t.L[HLEN - 1 - jy], hy = t.H[HLEN - 1 - jy], lx = t.L[HLEN - 1 - jx], hx = t.H[HLEN - 1 - jx];
            const int ki = (HLEN - 1 - jy) * HLEN + (HLEN - 1 - jx);
            S_k[i] = t.k2d ? make_float4(k2d_at(t.k2d, HLEN, 0, ki), k2d_at(t.k2d, HLEN, 1, ki), k2d_at(t.k2d, HLEN, 2, ki),
                                         k2d_at(t.k2d, HLEN, 3, ki))   // a custom filter quadruple (wt.cu:560-583)
                           : make_float4(__fmul_rn(ly, lx), __fmul_rn(ly, hx), __fmul_rn(hy, lx), __fmul_rn(hy, hx));
        }
    }
    pdl_wait();
    // input tile with the reference's fold (periodic; odd sizes repeat the last sample), nonseparable.cu:139-152
    // (asynchronous 4-byte copies: the whole tile is in flight at once; a warp takes a row, the row fold is applied once
    // per row and the column fold only by the tiles at the left and right edge)
    {
        const int xs = 2 * gx0 - K::C;
        const bool interior = xs >= 0 && xs + K::INW <= Nc;
        const int lane = tid & 31;
        for (int r = tid >> 5; r < inr; r += kNsThreads / 32) {
            const float* row = img + (size_t)ns_clamp(fold_dec(2 * gy0 - K::C + r, Nr), Nr - 1) * Nc;
            float* d = S_in + r * K::PITCH;
            if (interior) {
#pragma unroll
                for (int k = 0; k < (K::INW + 31) / 32; k++) {
                    const int u = 32 * k + lane;
                    if (32 * k + 31 < K::INW || u < K::INW) ns_cp4(d + K::slot(u), row + xs + u);
                }
            } else {
                for (int u = lane; u < K::INW; u += 32) ns_cp4(d + K::slot(u), row + ns_clamp(fold_dec(xs + u, Nc), Nc - 1));
            }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();

    const int tx = tid % (K::TW / K::U), ty0 = tid / (K::TW / K::U);
#pragma unroll 1
    for (int g = 0; g < rg; g++) {
    const int ty = ty0 + g * K::TH;
    const int gy = gy0 + ty, gx = gx0 + K::U * tx;
    if (gy >= nr || gx >= nc) continue;   // no barrier below
    u64 aAH[K::U], aVD[K::U];
#pragma unroll
    for (int u = 0; u < K::U; u++) aAH[u] = aVD[u] = 0ull;
    static_assert(K::U == 4, "a thread's window starts at vector 2 tx");
    const float* win = S_in + (2 * ty) * K::PITCH + 4 * tx;
#pragma unroll 1
    for (int jy = 0; jy < HLEN; jy++) {
        float x[K::NV * 4];
        const float4* rp = reinterpret_cast<const float4*>(win + jy * K::PITCH);
#pragma unroll
        for (int i = 0; i < K::NV; i++) {
            const float4 f = rp[(i & 1) * K::HALF + (i >> 1)];
            x[4 * i] = f.x; x[4 * i + 1] = f.y; x[4 * i + 2] = f.z; x[4 * i + 3] = f.w;
        }
        const float4* kp = S_k + jy * HLEN;
#pragma unroll
        for (int jx = 0; jx < HLEN; jx++) {
            float4 k;
            if constexpr (KC) k = t.k[jy][jx];   // jy uniform, jx compile-time
            else k = kp[jx];
            const u64 kah = ns_pack2(k.x, k.y), kvd = ns_pack2(k.z, k.w);
#pragma unroll
            for (int u = 0; u < K::U; u++) {
                const float v = x[2 * u + jx];
                aAH[u] = ns_ffma2(ns_pack2(v, v), kah, aAH[u]);
                aVD[u] = ns_ffma2(ns_pack2(v, v), kvd, aVD[u]);
            }
        }
    }
    float oa[K::U], oh[K::U], ov[K::U], od[K::U];
#pragma unroll
    for (int u = 0; u < K::U; u++) {
        ns_unpack2(aAH[u], oa[u], oh[u]);
        ns_unpack2(aVD[u], ov[u], od[u]);
    }
    const size_t off = (size_t)gy * nc + gx;
    float* pa = A + (size_t)blockIdx.z * s_a + off;
    float* ph = H + (size_t)blockIdx.z * s_d + off;
    float* pv = V + (size_t)blockIdx.z * s_d + off;
    float* pd = D + (size_t)blockIdx.z * s_d + off;
    const bool vec = ((nc & 3) == 0) && ((s_a & 3) == 0) && ((s_d & 3) == 0) &&
                     (((uintptr_t)A | (uintptr_t)H | (uintptr_t)V | (uintptr_t)D) & 15) == 0;
    if (vec) {   // nc % 4 == 0 and gx % 4 == 0: the whole vector is inside
        *reinterpret_cast<float4*>(pa) = make_float4(oa[0], oa[1], oa[2], oa[3]);
        *reinterpret_cast<float4*>(ph) = make_float4(oh[0], oh[1], oh[2], oh[3]);
        *reinterpret_cast<float4*>(pv) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        *reinterpret_cast<float4*>(pd) = make_float4(od[0], od[1], od[2], od[3]);
    } else {
#pragma unroll
        for (int u = 0; u < K::U; u++)
            if (gx + u < nc) {
                pa[u] = oa[u]; ph[u] = oh[u]; pv[u] = ov[u]; pd[u] = od[u];
            }
    }
    }   // row groups
}

// ================================================================================================== inverse
template <int HLEN>
struct NsInvCfg {
    static constexpr int H2 = HLEN / 2;
    static constexpr int CC = H2 / 2;                 // synthesis centre (nonseparable.cu:182-196)
    static constexpr int SHIFT = (H2 & 1) ? 0 : 1;    // the reference's "virtual id" shift for even hlen/2
    static constexpr int WIN = H2 + SHIFT;            // coefficient window behind one 2x2 output block, per axis
    static constexpr int P = 2;                       // coefficient positions (along x) per thread
    static constexpr int TWC = 64, THC = 8;           // coefficient tile: 32 thread columns x 8 rows
    static constexpr int INW = TWC + WIN - 1;
    static constexpr int PITCH = INW;                 // in float2 pairs; even (WIN is odd), so a thread's window is 16-byte aligned
    static constexpr int NW = WIN + P - 1;            // window columns of a thread (both positions); even
    static_assert(P == 2 && (PITCH % 2) == 0 && (NW % 2) == 0, "128-bit window loads");
    static constexpr int MAXRG = 4;                   // a CTA stages once for up to MAXRG groups of THC coefficient rows
    // resident CTAs per SM the register allocation aims for: 3 (80 registers) where ptxas gets there without spilling
    static constexpr int MINB = (HLEN == 12 || HLEN == 16 || HLEN == 20) ? 2 : 3;
    // (A,H) tile + (V,D) tile as float2, then K'[ey][ex][jy][jx] as float4
    static constexpr size_t smem(int rg) { return sizeof(float2) * 2 * (size_t)(rg * THC + WIN - 1) * PITCH; }
};
// The synthesis products K'[ey][ex][jy][jx] = (LL', LH', HL', HH') travel as a KERNEL PARAMETER (constant bank): the
// inner loop fetches them through the uniform datapath (LDCU) into uniform registers that FFMA2 reads directly, instead
// of 28 broadcast shared-memory loads per window row that competed with the window loads for the one shared-memory pipe
// of the SM (the kernel was co-limited by it: 60 wavefronts against 224 FP32-pipe cycles per warp and window row).
template <int HLEN>
struct NsInvK {
    float4 k[2][2][HLEN / 2][HLEN / 2];
};

template <int HLEN>
__global__ void __launch_bounds__(kNsThreads, NsInvCfg<HLEN>::MINB)
    k_nonsep_inv_tiled(const __grid_constant__ NsInvK<HLEN> kt, float* __restrict__ img, size_t s_img, const float* __restrict__ A,
                       size_t s_a, const float* __restrict__ H, const float* __restrict__ V, const float* __restrict__ D,
                       size_t s_d, int Nr, int Nc, int Nr2, int Nc2,   // Nr x Nc coefficients -> Nr2 x Nc2 pixels
                       int rg)                                          // groups of THC coefficient rows per CTA
{
    using K = NsInvCfg<HLEN>;
    constexpr int H2 = K::H2, SHIFT = K::SHIFT, WIN = K::WIN;
    extern __shared__ __align__(16) float smem[];
    const int inr = rg * K::THC + WIN - 1;          // staged rows: the halo is paid once for rg row groups
    float2* S_ah = reinterpret_cast<float2*>(smem);
    float2* S_vd = S_ah + inr * K::PITCH;
    const int tid = threadIdx.x;
    const int cx0 = blockIdx.x * K::TWC, cy0 = blockIdx.y * (K::THC * rg);

    pdl_wait();
    const float* pa = A + (size_t)blockIdx.z * s_a;
    const float* ph = H + (size_t)blockIdx.z * s_d;
    const float* pv = V + (size_t)blockIdx.z * s_d;
    const float* pd = D + (size_t)blockIdx.z * s_d;
    // coefficient tiles with the reference's single periodic wrap (nonseparable.cu:205-214), interleaved in pairs
    for (int i = tid; i < inr * K::INW; i += kNsThreads) {
        const int r = i / K::INW, u = i - r * K::INW;
        int y = cy0 - K::CC + r, x = cx0 - K::CC + u;
        y += (y < 0) ? Nr : 0;
        y -= (y >= Nr) ? Nr : 0;
        x += (x < 0) ? Nc : 0;
        x -= (x >= Nc) ? Nc : 0;
        const size_t o = (size_t)ns_clamp(y, Nr - 1) * Nc + ns_clamp(x, Nc - 1);
        float* dah = reinterpret_cast<float*>(S_ah + r * K::PITCH + u);
        float* dvd = reinterpret_cast<float*>(S_vd + r * K::PITCH + u);
        ns_cp4(dah, pa + o);
        ns_cp4(dah + 1, ph + o);
        ns_cp4(dvd, pv + o);
        ns_cp4(dvd + 1, pd + o);
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    pdl_launch_dependents();

    const int tx = tid % (K::TWC / K::P), ty0 = tid / (K::TWC / K::P);
    img += (size_t)blockIdx.z * s_img;
#pragma unroll 1
    for (int g = 0; g < rg; g++) {
    const int ty = ty0 + g * K::THC;
    if (cy0 + ty >= Nr) break;               // warp-uniform: a warp is one tile row
    u64 rah[K::P][2][2], rvd[K::P][2][2];   // [position][ey][ex] -> (ra, rh), (rv, rd)
#pragma unroll
    for (int p = 0; p < K::P; p++)
#pragma unroll
        for (int e = 0; e < 4; e++) rah[p][e >> 1][e & 1] = rvd[p][e >> 1][e & 1] = 0ull;
    const float2* wah = S_ah + ty * K::PITCH + K::P * tx;
    const float2* wvd = S_vd + ty * K::PITCH + K::P * tx;
    // Shared-memory traffic is what bounds this kernel (ncu: mio_throttle, 1 wavefront per clock and SM), so a window
    // row is fetched with 128-bit loads (two (A,H) or (V,D) pairs each: dense 512 bytes per warp) and every filter
    // quadruple K' is fetched ONCE and applied to both positions of the thread.
#pragma unroll 1
    for (int dy = 0; dy < WIN; dy++) {
        u64 dah[K::NW], dvd[K::NW];   // window column dx of position 0 = column dx-1 of position 1
#pragma unroll
        for (int q = 0; q < K::NW / 2; q++) {
            const float4 f = *reinterpret_cast<const float4*>(wah + dy * K::PITCH + 2 * q);
            const float4 g = *reinterpret_cast<const float4*>(wvd + dy * K::PITCH + 2 * q);
            dah[2 * q] = ns_pack2(f.x, f.y);
            dah[2 * q + 1] = ns_pack2(f.z, f.w);
            dvd[2 * q] = ns_pack2(g.x, g.y);
            dvd[2 * q + 1] = ns_pack2(g.z, g.w);
        }
#pragma unroll
        for (int ey = 0; ey < 2; ey++) {
            const int jy = dy - (ey ? SHIFT : 0);   // run-time (dy is), uniform
            if (jy < 0 || jy >= H2) continue;
#pragma unroll
            for (int ex = 0; ex < 2; ex++)
#pragma unroll
                for (int jx = 0; jx < H2; jx++) {   // ascending (jy, jx) for every output: the reference's order
                    const float4 k = kt.k[ey][ex][jy][jx];   // jy uniform, the rest compile-time
                    const u64 k01 = ns_pack2(k.x, k.y), k23 = ns_pack2(k.z, k.w);
#pragma unroll
                    for (int p = 0; p < K::P; p++) {
                        const int dx = jx + p + (ex ? SHIFT : 0);
                        rah[p][ey][ex] = ns_ffma2(dah[dx], k01, rah[p][ey][ex]);
                        rvd[p][ey][ex] = ns_ffma2(dvd[dx], k23, rvd[p][ey][ex]);
                    }
                }
        }
    }
#pragma unroll
    for (int ey = 0; ey < 2; ey++) {
        const int gy = 2 * (cy0 + ty) + ey;
        if (gy >= Nr2) continue;
#pragma unroll
        for (int p = 0; p < K::P; p++)
#pragma unroll
            for (int ex = 0; ex < 2; ex++) {
                const int gx = 2 * (cx0 + K::P * tx + p) + ex;
                if (gx >= Nc2) continue;
                float ra, rh, rv, rd;
                ns_unpack2(rah[p][ey][ex], ra, rh);
                ns_unpack2(rvd[p][ey][ex], rv, rd);
                img[(size_t)gy * Nc2 + gx] = __fadd_rn(__fadd_rn(__fadd_rn(ra, rh), rv), rd);
            }
    }
    }   // row groups
}

// ================================================================================================ launchers
template <int HLEN>
static int launch_ns_fwd(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                         cudaStream_t s)
{
    using K = NsFwdCfg<HLEN>;
    PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_nonsep_fwd_tiled<HLEN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)K::smem(K::MAXRG)));
    PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_nonsep_fwd_tiled<HLEN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)K::smem(K::MAXRG)));
    // row groups per CTA, as in launch_ns_inv: amortise the staged halo while the grid still fills the GPU several times
    int rg = K::MAXRG;
    if (const char* e = getenv("PDWT_NS_RG")) rg = atoi(e);
    else
        while (rg > 1 && (long long)idiv_up(half_up(Nc), K::TW) * idiv_up(half_up(Nr), K::TH * rg) * batch < 3LL * 2 * 148) rg >>= 1;
    if (rg < 1) rg = 1;
    if (rg > K::MAXRG) rg = K::MAXRG;
    dim3 grid(idiv_up(half_up(Nc), K::TW), idiv_up(half_up(Nr), K::TH * rg), batch);
    if (grid.y > 65535u) return 0;
    PDWT_PROF(prof_tag("k_nonsep_fwd_tiled", Nr, Nc), s);
    // the filter table as a kernel parameter (uniform-datapath operands) is the default: C4 db7 377 -> 348 us on B200
    // (PDWT_NS_FWD_CONST=0 forms the table in shared memory instead; a custom quadruple always does)
    static const bool kconst = []() { const char* e = getenv("PDWT_NS_FWD_CONST"); return !e || atoi(e) != 0; }();
    if (kconst && !t.k2d) {   // opt-in: the filter table as a kernel parameter (same products: one fp32 multiplication each)
        NsFwdK<HLEN> kt;
        for (int jy = 0; jy < HLEN; jy++)
            for (int jx = 0; jx < HLEN; jx++) {
                const volatile float ly = t.L[HLEN - 1 - jy], hy = t.H[HLEN - 1 - jy];
                const volatile float lx = t.L[HLEN - 1 - jx], hx = t.H[HLEN - 1 - jx];
                volatile float ll = ly * lx, lh = ly * hx, hl = hy * lx, hh = hy * hx;
                kt.k[jy][jx] = make_float4(ll, lh, hl, hh);
            }
        PDWT_CUDA(launch_pdl(k_nonsep_fwd_tiled<HLEN, true>, grid, kNsThreads, K::smem(rg), s, kt, (const float*)img.p,
                             img.stride, A.p, A.stride, H.p, V.p, D.p, H.stride, Nr, Nc, rg));
    } else {
        PDWT_CUDA(launch_pdl(k_nonsep_fwd_tiled<HLEN, false>, grid, kNsThreads, K::smem(rg), s, t, (const float*)img.p,
                             img.stride, A.p, A.stride, H.p, V.p, D.p, H.stride, Nr, Nc, rg));
    }
    PDWT_LAUNCH_CHECK();
    return 1;
}

template <int HLEN>
static int launch_ns_inv(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int Nr2, int Nc2,
                         int batch, cudaStream_t s)
{
    using K = NsInvCfg<HLEN>;
    if (Nr < K::WIN || Nc < K::WIN) return 0;   // the single wrap must suffice
    PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_nonsep_inv_tiled<HLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)K::smem(K::MAXRG)));
    // Row groups per CTA: more groups amortise the staged halo (WIN - 1 rows) and the staging latency over more
    // arithmetic, as long as the grid still fills the GPU several times over (2 CTAs per SM are resident).
    int rg = K::MAXRG;
    if (const char* e = getenv("PDWT_NS_RG")) rg = atoi(e);
    else
        while (rg > 1 && (long long)idiv_up(Nc, K::TWC) * idiv_up(Nr, K::THC * rg) * batch < 3LL * 2 * 148) rg >>= 1;
    if (rg < 1) rg = 1;
    if (rg > K::MAXRG) rg = K::MAXRG;
    dim3 grid(idiv_up(Nc, K::TWC), idiv_up(Nr, K::THC * rg), batch);
    if (grid.y > 65535u) return 0;
    // synthesis products per output parity e (0 = even output index): tap j multiplies I?[hlen-1-(2j+off_e)] with
    // off_e = e ? SHIFT : 1-SHIFT (nonseparable.cu:186-204); each product is ONE fp32 multiplication,
    // rounded to nearest, exactly what the reference's w_outer leaves in its filter arrays (nonseparable.cu:16-24)
    NsInvK<HLEN> kt;
    for (int ey = 0; ey < 2; ey++)
        for (int ex = 0; ex < 2; ex++)
            for (int jy = 0; jy < K::H2; jy++)
                for (int jx = 0; jx < K::H2; jx++) {
                    const int oy = ey ? K::SHIFT : 1 - K::SHIFT, ox = ex ? K::SHIFT : 1 - K::SHIFT;
                    const volatile float ly = t.IL[HLEN - 1 - (2 * jy + oy)], hy = t.IH[HLEN - 1 - (2 * jy + oy)];
                    const volatile float lx = t.IL[HLEN - 1 - (2 * jx + ox)], hx = t.IH[HLEN - 1 - (2 * jx + ox)];
                    volatile float ll = ly * lx, lh = ly * hx, hl = hy * lx, hh = hy * hx;   // no contraction, no excess precision
                    kt.k[ey][ex][jy][jx] = make_float4(ll, lh, hl, hh);
                    if (t.hk2d) {   // a custom filter quadruple (wt.cu:585-602), reference indexing nonseparable.cu:216-219
                        const int ki = (HLEN - 1 - (2 * jy + oy)) * HLEN + (HLEN - 1 - (2 * jx + ox));
                        kt.k[ey][ex][jy][jx] = make_float4(k2d_at(t.hk2d, HLEN, 0, ki), k2d_at(t.hk2d, HLEN, 1, ki),
                                                           k2d_at(t.hk2d, HLEN, 2, ki), k2d_at(t.hk2d, HLEN, 3, ki));
                    }
                }
    PDWT_PROF(prof_tag("k_nonsep_inv_tiled", Nr2, Nc2), s);
    PDWT_CUDA(launch_pdl(k_nonsep_inv_tiled<HLEN>, grid, kNsThreads, K::smem(rg), s, kt, img.p, img.stride, (const float*)A.p,
                         A.stride, (const float*)H.p, (const float*)V.p, (const float*)D.p, H.stride, Nr, Nc, Nr2, Nc2, rg));
    PDWT_LAUNCH_CHECK();
    return 1;
}

#define PDWT_NS_HLEN_SWITCH(fn, ...)                \
    switch (t.hlen) {                               \
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

// 1 = handled, 0 = shape / filter length not covered (the caller uses the generic kernel), < 0 = error
int n_nonsep_fwd_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                       cudaStream_t s)
{
    if (batch > 65535 || Nr < t.hlen || Nc < t.hlen) return 0;
    PDWT_NS_HLEN_SWITCH(launch_ns_fwd, t, img, A, H, V, D, Nr, Nc, batch, s)
}

int n_nonsep_inv_level(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int Nr2, int Nc2,
                       int batch, cudaStream_t s)
{
    if (batch > 65535) return 0;
    PDWT_NS_HLEN_SWITCH(launch_ns_inv, t, img, A, H, V, D, Nr, Nc, Nr2, Nc2, batch, s)
}

}  // namespace pdwt
