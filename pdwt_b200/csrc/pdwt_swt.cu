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
//   forward :  S_in[R][TW + (hlen-1) f]  --row pass-->  (lo, hi) pairs [R][TW]  --column pass-->  A, H, V, D
//   inverse :  (A,H) then (V,D) pair tiles [R][TW + (hlen-1) f]  --column pass-->  (t1, t2) pairs [TH][TW + (hlen-1) f]
//              --row pass-->  image
// Values that are multiplied by the same tap pair sit next to each other in shared memory, so one 64-bit load
// delivers the packed operand of an FFMA2 / FMUL2.
//
// Arithmetic is the reference's, bit for bit: forward = one fmaf chain from 0 over ascending j
// (separable.cu:427-445, 470-489); inverse = round(v*k), exact halving, add, the two branch sums added last
// (separable.cu:575-588, 615-625).  Index folding is the reference's single +-N wrap (separable.cu:428-433).
// Shapes the tiles cannot serve (odd hlen, hlen > 20, dilations whose halo does not fit in shared memory) return 0 and
// the caller falls back to the generic two-pass kernels.
// The inverse of the common shapes (hlen <= 16, dilation <= 8) runs k_swt_inv_stream further down instead: a register
// window per column walks down the rows, nothing is restaged or recomputed (C3: 272 -> 223 us for the four levels).
#include "pdwt_common.cuh"
#include <algorithm>
#include <mutex>
#include <vector>

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
// a + b on both halves, rounded once, as an FFMA2 whose multiplicand `ones` = (1, 1) arrives as a KERNEL PARAMETER:
// a * 1 is exact, so this IS add.rn.  ptxas contracts a packed product with the packed add behind it into one FFMA2 --
// for add.rn.f32x2 and also for an fma with the literal (1, 1), even with --fmad=false (seen in the SASS) -- which
// would round once where the reference rounds twice; a multiplicand it cannot see through keeps the two roundings.
__device__ __forceinline__ u64 sw_add2_exact(u64 a, u64 ones, u64 b)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(ones), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sw_fadd2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ void sts_u64(unsigned addr, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ u64 lds_u64(unsigned addr)
{
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ int wrap1(int i, int N)   // fold_swt as a function of the unwrapped index, then clamped
{
    i += (i < 0) ? N : 0;
    i -= (i >= N) ? N : 0;
    return min(max(i, 0), N - 1);   // only indices of masked outputs can still be outside
}

// Row passes with compile-time dilation F <= 8: a thread produces the 4 outputs x0, x0 + F, x0 + 2F, x0 + 3F of one tile
// row -- they share all but one of their taps' samples, so hlen + 3 shared-memory loads serve 4 hlen taps.  The lanes of
// a warp (4-byte samples; half-warp for 8-byte pairs) are spread over NB/4 or F columns x0 = r + 4F g (r < F) of several
// rows; with a row pitch congruent to F modulo the number of banks NB (32 x 4 bytes, 16 x 8 bytes) every load of the
// warp touches each bank once.
__host__ __device__ constexpr int swt_pitch(int width, int f, int nbanks)
{
    return (f >= 1 && f <= 8) ? width + (((f - width) % nbanks) + nbanks) % nbanks : width;
}

// ================================================================================================== forward
template <int HLEN>
struct SwtFwdCfg {
    static constexpr int TW = 64, TH = 32;
    static constexpr int C = HLEN / 2 - 1;          // centre_fwd for even hlen (separable.cu:417-421), in units of f
    static constexpr int R = TH + HLEN - 1;         // staged rows (one residue class)
    static constexpr int RPT = 4;                   // output rows per column-pass task
    static size_t smem(int f) { return sizeof(float) * ((size_t)R * swt_pitch(TW + (HLEN - 1) * f, f, 32) + 1 + 2 * (size_t)R * TW); }
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
    const int width = K::TW + (HLEN - 1) * f, pitch = swt_pitch(width, f, 32);
    float* S_in = smem;
    u64* S_lh = reinterpret_cast<u64*>(smem + ((K::R * pitch + 1) & ~1));   // [R][TW] (lo, hi) pairs, 8-byte aligned
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ry = blockIdx.y % f, mt = blockIdx.y / f;
    const int gx0 = blockIdx.x * K::TW;
    src += (size_t)blockIdx.z * s_src;
    pdl_wait();

    // ---- stage the tile: row t of the tile is image row ry + f*(mt*TH + t - C), column u is gx0 - C*f + u
    // Tiles whose columns need no fold (all but the first and last of a row of tiles) take the copies with immediate
    // offsets, one LDGSTS per 32 columns; the others fold every column index (separable.cu:428-433).
    const int xs = gx0 - K::C * f;
    const bool interior = F != 0 && xs >= 0 && xs + width <= Nc;
    for (int r = warp; r < K::R; r += kSwtThreads / 32) {
        const int gy = wrap1(ry + f * (mt * K::TH + r - K::C), Nr);
        const float* row = src + (size_t)gy * Nc;
        if (interior) {
            constexpr int WC = K::TW + (HLEN - 1) * F;   // = width
            const float* g = row + xs + lane;
            float* d = S_in + r * pitch + lane;
#pragma unroll
            for (int k = 0; k < (WC + 31) / 32; k++)
                if (32 * k + 31 < WC || 32 * k + lane < WC) cp_async4(d + 32 * k, g + 32 * k);
        } else {
            for (int u = lane; u < width; u += 32) cp_async4(S_in + r * pitch + u, row + wrap1(xs + u, Nc));
        }
    }
    cp_async_wait_all0();
    __syncthreads();

    // ---- row pass, w_kern_forward_swt_pass1 (separable.cu:409-448): lo/hi[x] = sum_j in[x + (j - C) f] * L/H[hlen-1-j]
    if (F >= 1 && F <= 8) {
        // 4 outputs F apart per thread (see swt_pitch): a warp takes 4 rows x 32 columns
        constexpr int FF = F ? F : 1;
        const int l8 = lane & 7, mrow = lane >> 3;
        const int xl = (l8 % FF) + 4 * FF * (l8 / FF);
        constexpr int NCH = K::TW / 32;
        for (int item = warp; item < ((K::R + 3) / 4) * NCH; item += kSwtThreads / 32) {
            const int r = (item / NCH) * 4 + mrow, x0 = (item % NCH) * 32 + xl;
            if (r < K::R) {
                const float* p = S_in + r * pitch + x0;
                u64 lh[4] = {0ull, 0ull, 0ull, 0ull};   // (lo, hi) += v * (L, H)[hlen-1-j], j ascending for every output
#pragma unroll
                for (int i = 0; i < HLEN + 3; i++) {
                    const float v = p[i * FF];
                    const u64 vv = sw_pack2(v, v);
#pragma unroll
                    for (int o = 0; o < 4; o++) {
                        const int j = i - o;
                        if (j >= 0 && j < HLEN) lh[o] = sw_ffma2(vv, sw_pack2(t.L[HLEN - 1 - j], t.H[HLEN - 1 - j]), lh[o]);
                    }
                }
#pragma unroll
                for (int o = 0; o < 4; o++) S_lh[r * K::TW + x0 + o * FF] = lh[o];
            }
        }
    } else {
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
                S_lh[r * K::TW + xx] = lh;
            }
        }
    }
    __syncthreads();
    pdl_launch_dependents();

    // ---- column pass, w_kern_forward_swt_pass2 (separable.cu:452-493): inside the residue class the taps are adjacent
    // tile rows.  A task = one column x RPT consecutive tile rows, 4 sub-bands.
    const int xx = tid % K::TW;
    const int gx = gx0 + xx;
    // output addressing hoisted out of the task loop: four plane pointers at this thread's column, one row step
    // (the stores were 116 of the 271 instructions of a task when every one recomputed plane, row and column)
    float* oA = A + (size_t)blockIdx.z * s_a + gx;
    float* oH = H + (size_t)blockIdx.z * s_d + gx;
    float* oV = V + (size_t)blockIdx.z * s_d + gx;
    float* oD = D + (size_t)blockIdx.z * s_d + gx;
    asm volatile("" : "+l"(oA), "+l"(oH), "+l"(oV), "+l"(oD));   // keep them: ptxas otherwise redoes blockIdx.z * stride per store
    const size_t row_step = (size_t)f * Nc;
    const bool col_in = gx < Nc;
    for (int m0 = (tid / K::TW) * K::RPT; m0 < K::TH; m0 += (kSwtThreads / K::TW) * K::RPT) {
        u64 ah[K::RPT], vd[K::RPT];   // (A, H) += lo * (L, H)[..],  (V, D) += hi * (L, H)[..]
#pragma unroll
        for (int o = 0; o < K::RPT; o++) ah[o] = vd[o] = 0ull;
#pragma unroll
        for (int i = 0; i < K::RPT + HLEN - 1; i++) {
            float v1, v2;
            sw_unpack2(S_lh[(m0 + i) * K::TW + xx], v1, v2);
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
        const int gy0 = ry + f * (mt * K::TH + m0);
        size_t off = (size_t)gy0 * Nc;
#pragma unroll
        for (int o = 0; o < K::RPT; o++) {
            if (col_in && gy0 + o * f < Nr) {
                float a, h, v, d;
                sw_unpack2(ah[o], a, h);
                sw_unpack2(vd[o], v, d);
                oA[off] = a;
                oH[off] = h;
                oV[off] = v;
                oD[off] = d;
            }
            off += row_step;
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
    static size_t smem(int f) { return sizeof(float) * (2 * (size_t)R + 2 * (size_t)TH) * swt_pitch(TW + (HLEN - 1) * f, f, 16); }
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
    float2 one;       // (1, 1), opaque to the compiler (see sw_add2_exact)
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
    const u64 ones = sw_pack2(t.one.x, t.one.y);
    extern __shared__ __align__(16) float smem[];
    const int ci = K::TW + (HLEN - 1) * f;           // columns of t1/t2 the row pass of this tile reads
    const int cip = swt_pitch(ci, f, 16);            // row pitch of the pair tiles (elements of 8 bytes)
    float* S_c = smem;                               // [R][cip] pairs (A, H), then (V, D)
    float* S_t = S_c + 2 * K::R * cip;               // [TH][cip] pairs (t1, t2)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ry = blockIdx.y % f, mt = blockIdx.y / f;
    const int gx0 = blockIdx.x * K::TW;
    pdl_wait();

#pragma unroll 1
    for (int pair = 0; pair < 2; pair++) {
        const float* c0 = (pair ? V + (size_t)blockIdx.z * s_d : A + (size_t)blockIdx.z * s_a);
        const float* c1 = (pair ? D : H) + (size_t)blockIdx.z * s_d;
        if (pair) __syncthreads();   // the column pass of the first pair has finished reading S_c
        const int xs = gx0 - K::C * f;
        const bool interior = F != 0 && xs >= 0 && xs + ci <= Nc;   // no column fold: copies with immediate offsets
        for (int r = warp; r < K::R; r += kSwtThreads / 32) {
            const size_t rowo = (size_t)wrap1(ry + f * (mt * K::TH + r - K::C), Nr) * Nc;
            if (interior) {
                constexpr int WC = K::TW + (HLEN - 1) * F;   // = ci
                const float* g0 = c0 + rowo + xs + lane;
                const float* g1 = c1 + rowo + xs + lane;
                float* d = S_c + 2 * (r * cip + lane);
#pragma unroll
                for (int k = 0; k < (WC + 31) / 32; k++)
                    if (32 * k + 31 < WC || 32 * k + lane < WC) {
                        cp_async4(d + 64 * k, g0 + 32 * k);
                        cp_async4(d + 64 * k + 1, g1 + 32 * k);
                    }
            } else {
                for (int u = lane; u < ci; u += 32) {
                    const size_t o = rowo + wrap1(xs + u, Nc);
                    cp_async4(S_c + 2 * (r * cip + u), c0 + o);
                    cp_async4(S_c + 2 * (r * cip + u) + 1, c1 + o);
                }
            }
        }
        cp_async_wait_all0();
        __syncthreads();
        // column pass, w_kern_inverse_swt_pass1 (separable.cu:553-589): t = IL_y(c0) + IH_y(c1), RPT rows per task
        float* S_out = S_t + pair;   // t1 in the even words, t2 in the odd ones
        const int ncb = (ci + 31) / 32;
        for (int task = warp; task < ncb * (K::TH / K::RPT); task += kSwtThreads / 32) {
            const int u = (task % ncb) * 32 + lane, m0 = (task / ncb) * K::RPT;
            if (u < ci) {
                u64 rlh[K::RPT];   // (rl, rh)
#pragma unroll
                for (int o = 0; o < K::RPT; o++) rlh[o] = 0ull;
#pragma unroll
                for (int i = 0; i < K::RPT + HLEN - 1; i++) {
                    const u64 v01 = reinterpret_cast<const u64*>(S_c)[(m0 + i) * cip + u];
#pragma unroll
                    for (int o = 0; o < K::RPT; o++) {
                        const int j = i - o;
                        // the two products rounded on their own (one FMUL2), then added (see sw_add2_exact)
                        if (j >= 0 && j < HLEN) rlh[o] = sw_add2_exact(sw_fmul2(v01, khalf(j)), ones, rlh[o]);
                    }
                }
#pragma unroll
                for (int o = 0; o < K::RPT; o++) {
                    float rl, rh;
                    sw_unpack2(rlh[o], rl, rh);
                    S_out[2 * ((m0 + o) * cip + u)] = __fadd_rn(rl, rh);
                }
            }
        }
    }
    __syncthreads();
    pdl_launch_dependents();

    // ---- row pass, w_kern_inverse_swt_pass2 (separable.cu:593-626): img[x] = IL_x(t1) + IH_x(t2), taps f apart
    dst += (size_t)blockIdx.z * s_dst;
    const u64* S_t12 = reinterpret_cast<const u64*>(S_t);
    if (F >= 1 && F <= 8) {
        // 4 outputs F apart per thread (see swt_pitch): a warp takes 8 rows x 16 columns (F <= 4) or 4 rows x 32 columns
        constexpr int FF = F ? F : 1;
        constexpr int LPR = FF < 4 ? 4 : FF, RPW = 32 / LPR, CW = 4 * LPR, NCH = K::TW / CW;
        const int ll = lane % LPR, mrow = lane / LPR;
        const int xl = (ll % FF) + 4 * FF * (ll / FF);
        const bool vec_ok = FF == 1 && (Nc & 3) == 0 && (s_dst & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
        for (int item = warp; item < (K::TH / RPW) * NCH; item += kSwtThreads / 32) {
            const int m = (item / NCH) * RPW + mrow, x0 = (item % NCH) * CW + xl;
            const u64* p12 = S_t12 + m * cip + x0;
            u64 acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int i = 0; i < HLEN + 3; i++) {
                const u64 v12 = p12[i * FF];
#pragma unroll
                for (int o = 0; o < 4; o++) {
                    const int j = i - o;
                    if (j >= 0 && j < HLEN) acc[o] = sw_add2_exact(sw_fmul2(v12, khalf(j)), ones, acc[o]);
                }
            }
            float res[4];
#pragma unroll
            for (int o = 0; o < 4; o++) {
                float a1, a2;
                sw_unpack2(acc[o], a1, a2);
                res[o] = __fadd_rn(a1, a2);
            }
            const int gy = ry + f * (mt * K::TH + m), gx = gx0 + x0;
            if (gy < Nr) {
                float* q = dst + (size_t)gy * Nc + gx;
                if (vec_ok && gx + 3 < Nc) {
                    *reinterpret_cast<float4*>(q) = make_float4(res[0], res[1], res[2], res[3]);
                } else {
#pragma unroll
                    for (int o = 0; o < 4; o++)
                        if (gx + o * FF < Nc) q[o * FF] = res[o];
                }
            }
        }
    } else {
        const int xx = tid % K::TW;
        const int gx = gx0 + xx;
        for (int m = tid / K::TW; m < K::TH; m += kSwtThreads / K::TW) {
            const u64* p12 = S_t12 + m * cip + xx;
            u64 a12 = 0ull;
#pragma unroll
            for (int j = 0; j < HLEN; j++) a12 = sw_add2_exact(sw_fmul2(p12[j * f], khalf(j)), ones, a12);
            float a1, a2;
            sw_unpack2(a12, a1, a2);
            const int gy = ry + f * (mt * K::TH + m);
            if (gy < Nr && gx < Nc) dst[(size_t)gy * Nc + gx] = __fadd_rn(a1, a2);
        }
    }
}

// ======================================================================================== inverse, streaming
// The tiled inverse above recomputes: its column pass covers the dilated halo of the row pass (TW + 15 f columns for
// TW outputs: 1.9x at f = 8) and every tile restages hlen - 1 rows for its TH = 16 (another 1.9x of staging).  This
// kernel removes the second factor and most of the first:
//   * a thread owns ONE column of t1/t2 and walks down the rows of one residue class modulo f.  The hlen rows its
//     column pass needs live in a REGISTER window of (A,H) and (V,D) pairs; the loop is unrolled hlen times so the
//     window rotates through static register names (no moves) -- a new row costs 4 coalesced loads, not a restage;
//   * the finished (t1, t2) pair goes to a double-buffered row in shared memory, one block barrier per row, and the row
//     pass (taps f apart) reads it from there.  The CTA is as wide as the register file allows (up to 640 columns), so
//     the halo of 15 f columns is a small share of it: 1.03x (f = 1) ... 1.23x (f = 8) of the column pass, none of the
//     row pass.
// Arithmetic and index folding are those of the tiled kernel (separable.cu:553-626), bit for bit.
constexpr int kSwtStreamMaxThreads = 512;
template <int HLEN, int F>
__global__ void __launch_bounds__(kSwtStreamMaxThreads, 1)
    k_swt_inv_stream(const __grid_constant__ SwtHalfTaps<HLEN> t, const float* __restrict__ A, size_t s_a,
                     const float* __restrict__ H, const float* __restrict__ V, const float* __restrict__ D, size_t s_d,
                     float* __restrict__ dst, size_t s_dst, int Nr, int Nc, int two, int ch)
{
    constexpr int C = HLEN / 2;
    __shared__ u64 S_t[2][kSwtStreamMaxThreads + (HLEN - 1) * F];   // (t1, t2) of the current / next row; the tail is
                                                                    // never written: threads that read it have no output
    auto khalf = [&](const int j) { return sw_pack2(t.k[j].x, t.k[j].y); };
    const u64 ones = sw_pack2(t.one.x, t.one.y);
    const int tid = threadIdx.x;
    const int ry = blockIdx.y % F, chunk = blockIdx.y / F;
    const int gx0 = blockIdx.x * two;
    const int m0 = chunk * ch;
    const int m1 = min(m0 + ch, (Nr - ry + F - 1) / F);             // class rows [m0, m1) of this chunk are inside the image
    const int xc = wrap1(gx0 - C * F + tid, Nc);                    // this thread's column of t1/t2
    // the four source pointers walk down the class rows m0 - C, m0 - C + 1, ... (image rows f apart, folded back by Nr:
    // the reference's single wrap for every row an output needs -- (hlen-1) f < Nr -- and in range for the two rows
    // that are fetched ahead past the chunk's end)
    int gy = ry + F * (m0 - C);
    gy += (gy < 0) ? Nr : 0;
    gy = min(max(gy, 0), Nr - 1);   // only chunks without rows (m0 >= m1) can still be outside
    // one 32-bit element offset serves the four planes (uniform base pointers): Nr * Nc < 2^31 is checked by the launcher
    const float* bA = A + (size_t)blockIdx.z * s_a;
    const float* bH = H + (size_t)blockIdx.z * s_d;
    const float* bV = V + (size_t)blockIdx.z * s_d;
    const float* bD = D + (size_t)blockIdx.z * s_d;
    asm volatile("" : "+l"(bA), "+l"(bH), "+l"(bV), "+l"(bD));   // keep them: ptxas otherwise redoes blockIdx.z * stride in every step
    const unsigned step_dn = (unsigned)F * (unsigned)Nc, plane_n = (unsigned)Nr * (unsigned)Nc;
    unsigned off = (unsigned)gy * (unsigned)Nc + (unsigned)xc;
    auto next_row = [&]() {
        gy += F;
        off += step_dn;
        const bool w = gy >= Nr;
        gy -= w ? Nr : 0;
        off -= w ? plane_n : 0u;
    };
    const bool has_out = tid < two && gx0 + tid < Nc;
    // row pass output of step m is class row m - 1: the offset starts one row early (never stored through)
    float* bO = dst + (size_t)blockIdx.z * s_dst;
    asm volatile("" : "+l"(bO));
    unsigned oo = (unsigned)(ry + F * m0) * (unsigned)Nc + (unsigned)(gx0 + tid) - step_dn;
    unsigned st_off = (unsigned)__cvta_generic_to_shared(&S_t[0][tid]);
    asm volatile("" : "+r"(st_off));   // opaque: otherwise ptxas re-derives it from SR_TID in every unrolled step
    constexpr unsigned kRowBytes = sizeof(S_t[0]);
    pdl_wait();

    // window slot s holds class row m - C + j of the current output row m, with s = (u + j) % HLEN at unrolled step u.
    // The slot of tap 0 is free as soon as that tap is done -- the first thing a step does -- and it is the slot of the
    // next step's LAST tap: the loads of the next row go straight into it and have almost two steps to arrive.
    // (Measured on C3: landing registers that keep 1, 2 or 4 rows in flight, an L2 prefetch 8 rows ahead and a
    // shared-memory ring fed by LDGSTS all run at the same speed or slower; so does the kernel without its barrier.
    // It is bound by FP32 issue -- 96 two-cycle packed instructions per row and thread plus ~55 others.)
    u64 wAH[HLEN], wVD[HLEN];
#pragma unroll
    for (int j = 0; j < HLEN; j++) {
        wAH[j] = sw_pack2(__ldg(bA + off), __ldg(bH + off));
        wVD[j] = sw_pack2(__ldg(bV + off), __ldg(bD + off));
        next_row();
    }

    // Step m: the column pass of class row m (into one shared row) and the row pass of class row m - 1 (out of the
    // other one) are independent, so they sit in ONE basic block and ptxas interleaves them -- the shared-memory loads
    // of the row pass are covered by the column pass's arithmetic.  One barrier per step.  The first step's row pass and
    // the last step's column pass compute values nobody keeps (the store is predicated; the loads stay in range).
    for (int mb = m0; mb <= m1; mb += HLEN) {
#pragma unroll
        for (int u = 0; u < HLEN; u++) {
            const int m = mb + u;
            if (m <= m1) {   // uniform over the CTA
                if (m >= m1) pdl_launch_dependents();
                // row pass of the previous step's row, w_kern_inverse_swt_pass2 (separable.cu:593-626)
                const unsigned rd = st_off + ((u + 1) & 1) * kRowBytes;
                u64 acc = 0ull;
#pragma unroll
                for (int j = 0; j < HLEN; j++) acc = sw_add2_exact(sw_fmul2(lds_u64(rd + j * F * 8), khalf(j)), ones, acc);
                // column pass, w_kern_inverse_swt_pass1 (separable.cu:553-589)
                u64 r1 = sw_add2_exact(sw_fmul2(wAH[u % HLEN], khalf(0)), ones, 0ull);
                u64 r2 = sw_add2_exact(sw_fmul2(wVD[u % HLEN], khalf(0)), ones, 0ull);
                wAH[u % HLEN] = sw_pack2(__ldg(bA + off), __ldg(bH + off));
                wVD[u % HLEN] = sw_pack2(__ldg(bV + off), __ldg(bD + off));
                next_row();
#pragma unroll
                for (int j = 1; j < HLEN; j++) {
                    r1 = sw_add2_exact(sw_fmul2(wAH[(u + j) % HLEN], khalf(j)), ones, r1);
                    r2 = sw_add2_exact(sw_fmul2(wVD[(u + j) % HLEN], khalf(j)), ones, r2);
                }
                float a, b, c, d, a1, a2;
                sw_unpack2(r1, a, b);
                sw_unpack2(r2, c, d);
                sts_u64(st_off + (u & 1) * kRowBytes, sw_pack2(__fadd_rn(a, b), __fadd_rn(c, d)));
                sw_unpack2(acc, a1, a2);
                if (has_out && m > m0) bO[oo] = __fadd_rn(a1, a2);
                oo += step_dn;
                __syncthreads();
            }
        }
    }
}

// Geometry of a streaming launch.  A CTA is `threads` columns wide (the last (hlen-1) f of them only feed the row pass
// of the others) and walks `ch` rows of one residue class.  Both are chosen by a cost model of the whole grid: CTAs
// that fit on an SM at 128 registers per thread, waves of CTAs, steps per CTA (plus a few for the window fill), cost
// of a step proportional to the warps resident on the SM -- the kernel is bound by FP32 issue, so what matters is that
// the LAST wave is as full as the others and that the halo share stays small.
struct SwtStreamPlan {
    int threads, two, ntiles, ch, nchunks;
};
static bool swt_stream_plan(int hlen, int Nr, int Nc, int f, int batch, SwtStreamPlan& sp)
{
    struct Key {
        int hlen, Nr, Nc, f, batch, dev;
        SwtStreamPlan sp;
        bool ok;
    };
    static std::mutex mu;
    static std::vector<Key> cache;
    static const int env_cw = getenv("PDWT_SWT_CW") ? atoi(getenv("PDWT_SWT_CW")) : 0;
    static const int env_ch = getenv("PDWT_SWT_CH") ? atoi(getenv("PDWT_SWT_CH")) : 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        dev = 0;
    }
    std::lock_guard<std::mutex> lock(mu);
    for (const Key& k : cache)
        if (k.hlen == hlen && k.Nr == Nr && k.Nc == Nc && k.f == f && k.batch == batch && k.dev == dev) {
            sp = k.sp;
            return k.ok;
        }
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) {
        cudaGetLastError();
        sms = 148;
    }
    const int halo = (hlen - 1) * f, rpc = idiv_up(Nr, f);
    double best = -1.0;
    sp.threads = 0;
    const int max_thr = kSwtStreamMaxThreads, thr_per_sm = 512;   // 65536 registers / 128 per thread
    const double extra = 6.0;                                      // window fill and launch, in steps
    for (int cw = 64; cw <= max_thr; cw += 32) {
        if (env_cw && cw != env_cw) continue;
        const int two = cw - halo;
        if (two < 32 && two < Nc) continue;
        if (two < 1) continue;
        const int nt = idiv_up(Nc, two);
        const int cps = std::min(4, thr_per_sm / cw);   // 65536 registers / (registers per thread x cw)
        const long long slots = (long long)sms * cps;
        for (int n = 1; n <= rpc && n <= 1024; n++) {
            int ch = idiv_up(rpc, n);
            if (env_ch) ch = env_ch;
            const int nch = idiv_up(rpc, ch);
            if ((long long)f * nch > 65535) continue;
            const long long ncta = (long long)nt * f * nch * batch;
            const long long waves = (ncta + slots - 1) / slots;
            // a wave that does not fill the machine leaves its CTAs the SM to themselves: they run faster
            const double per_sm = waves > 1 ? (double)cps : std::min<double>(cps, (double)ncta / sms < 1.0 ? 1.0 : (double)ncta / sms);
            const double cost = (double)waves * (ch + extra) * std::max(per_sm * cw / 32.0, 16.0);   // < 16 warps on an SM do not keep the FP32 pipe busy
            if (best < 0 || cost < best * 0.999 || (cost <= best * 1.001 && cw > sp.threads)) {
                best = cost;
                sp.threads = cw;
                sp.two = two;
                sp.ntiles = nt;
                sp.ch = ch;
                sp.nchunks = nch;
            }
            if (env_ch) break;
        }
    }
    const bool ok = sp.threads != 0;
    if (cache.size() < 64) cache.push_back(Key{hlen, Nr, Nc, f, batch, dev, sp, ok});
    return ok;
}

template <int HLEN, int F>
static int launch_swt_inv_stream_f(const SwtHalfTaps<HLEN>& ht, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int Nr,
                                   int Nc, int batch, cudaStream_t s)
{
    SwtStreamPlan sp;
    if ((long long)Nr * Nc >= (1ll << 31) || !swt_stream_plan(HLEN, Nr, Nc, F, batch, sp)) return 0;   // 32-bit offsets in the kernel
    dim3 grid(sp.ntiles, F * sp.nchunks, batch);
    PDWT_PROF(prof_tag("k_swt_inv_stream", Nr, F), s);
    PDWT_CUDA(launch_pdl(k_swt_inv_stream<HLEN, F>, grid, sp.threads, 0, s, ht, (const float*)A.p, A.stride,
                         (const float*)H.p, (const float*)V.p, (const float*)D.p, H.stride, dst.p, dst.stride, Nr, Nc,
                         sp.two, sp.ch));
    PDWT_LAUNCH_CHECK();
    return 1;
}
static bool swt_inv_stream_on()
{
    const char* e = getenv("PDWT_SWT_INV_STREAM");   // 0: the tiled kernels for every level
    return !(e && e[0] == '0');
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
        PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_swt_fwd_fused<HLEN, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                  (int)kSwtSmemCap));                                                    \
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
    SwtHalfTaps<HLEN> ht;
    for (int j = 0; j < HLEN; j++) ht.k[j] = make_float2(t.IL[HLEN - 1 - j] * 0.5f, t.IH[HLEN - 1 - j] * 0.5f);
    ht.one = make_float2(1.0f, 1.0f);
    if constexpr (HLEN <= 16) {
        if (f <= 8 && swt_inv_stream_on()) {
            int done = 0;
            switch (f) {
                case 1: done = launch_swt_inv_stream_f<HLEN, 1>(ht, A, H, V, D, dst, Nr, Nc, batch, s); break;
                case 2: done = launch_swt_inv_stream_f<HLEN, 2>(ht, A, H, V, D, dst, Nr, Nc, batch, s); break;
                case 4: done = launch_swt_inv_stream_f<HLEN, 4>(ht, A, H, V, D, dst, Nr, Nc, batch, s); break;
                default: done = launch_swt_inv_stream_f<HLEN, 8>(ht, A, H, V, D, dst, Nr, Nc, batch, s); break;
            }
            if (done) return done;
        }
    }
    PDWT_PROF(prof_tag("k_swt_inv_fused", Nr, f), s);
#define PDWT_SWT_INV(FF)                                                                                                 \
    do {                                                                                                                 \
        PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_swt_inv_fused<HLEN, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                  (int)kSwtSmemCap));                                                    \
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
