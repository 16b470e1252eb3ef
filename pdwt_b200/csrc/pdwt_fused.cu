// pdwt_fused.cu -- fused per-level kernels of the separable 2-D DWT (the headline path).
//
// The reference runs two kernels per level through two full-size scratch images (row pass -> tmp -> column pass;
// separable.cu:179-209, 332-364): 42 B/pixel of HBM traffic for db7 L3 against 16 B/pixel algorithmic.  Here one
// kernel per level does both passes on a tile held in shared memory:
//
//   forward :  global --(16-byte cp.async, periodic halo folded while staging)--> S_in
//              row pass   S_in  -> S_lo / S_hi   (4 outputs x 2 filters per thread, taps fully unrolled, tap values
//                                                 read as constant-bank operands of the FFMA)
//              column pass S_lo/S_hi -> A,H,V,D  (4 columns x RPT rows per thread, 128-bit coalesced stores)
//   inverse :  A,H,V,D tiles -> column synthesis -> S_t1/S_t2 -> row synthesis -> image tile
//
// Arithmetic is the reference's: every output is one fmaf chain from 0 over its taps in ascending j, the row-pass
// result is rounded to fp32 before the column pass (as the reference's d_tmp round trip does), and the inverse adds
// the two branch sums last -- so the results are bit-identical to the generic kernels and to the reference.
//
// Tile geometry is a compile-time function of the filter length; shapes that the fast staging path cannot serve
// (image edges, widths that are not a multiple of 4, odd sizes) take a per-element folded staging loop in the same
// kernel, so every 2-D shape is covered for even hlen in [4, kMaxFusedHlen].
#include "pdwt_common.cuh"

namespace pdwt {

constexpr int kFusedThreads = 256;
constexpr int kMaxFusedHlen = 20;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// round n up to a multiple of 4 whose quarter is odd (conflict-free 128-bit rows when lanes walk down rows)
__host__ __device__ constexpr int odd_quad_pitch(int n)
{
    int p = (n + 3) / 4 * 4;
    return ((p / 4) & 1) ? p : p + 4;
}

// ================================================================================================== forward
template <int HLEN, int TW, int TH>
struct FwdCfg {
    static constexpr int C = HLEN / 2 - 1;            // analysis centre for even hlen, separable.cu:103-107
    static constexpr int HALO = HLEN - 2;
    static constexpr int SH = (4 - (C & 3)) & 3;      // staging starts at the 16-byte boundary left of 2*gx0 - C
    static constexpr int INR = 2 * TH + HALO;         // staged rows
    static constexpr int INW = 2 * TW + HALO + SH;    // staged columns that are actually consumed
    static constexpr int INW4 = (INW + 3) / 4;        // ... in 16-byte vectors
    static constexpr int NV = (SH + 6 + HLEN + 3) / 4;  // vectors one row-pass task reads (4 outputs)
    static constexpr int INP = odd_quad_pitch((INW4 * 4 > 2 * TW - 8 + NV * 4) ? INW4 * 4 : 2 * TW - 8 + NV * 4);
    static constexpr int MP = odd_quad_pitch(TW);     // pitch of S_lo / S_hi
    static constexpr int RPT = TH >= 32 ? 4 : 2;      // output rows per column-pass task
    static constexpr size_t SMEM = sizeof(float) * ((size_t)INR * INP + 2 * (size_t)INR * MP);
    static_assert(TH % RPT == 0 && TW % 16 == 0, "tile shape");
};

template <int HLEN, int TW, int TH>
__global__ void __launch_bounds__(kFusedThreads, 2)
    k_fwd2d(const __grid_constant__ Taps t, const float* __restrict__ src, size_t s_src, float* __restrict__ A,
            size_t s_a, float* __restrict__ H, float* __restrict__ V, float* __restrict__ D, size_t s_d, int Nr, int Nc)
{
    using K = FwdCfg<HLEN, TW, TH>;
    extern __shared__ __align__(16) float smem[];
    float* S_in = smem;
    float* S_lo = smem + K::INR * K::INP;
    float* S_hi = S_lo + K::INR * K::MP;

    const int tid = threadIdx.x;
    const int nr = half_up(Nr), nc = half_up(Nc);
    const int gx0 = blockIdx.x * TW, gy0 = blockIdx.y * TH;
    src += (size_t)blockIdx.z * s_src;
    const int x_al = 2 * gx0 - K::C - K::SH;  // multiple of 4
    const int y_0 = 2 * gy0 - K::C;
    pdl_wait();   // the kernel that wrote `src` (previous level) has completed

    // ---- stage the input tile, folding the periodic / odd-size extension (separable.cu:114-121) while loading
    const bool fast = ((Nc & 3) == 0) && x_al >= 0 && (x_al + K::INW4 * 4 <= Nc) && ((((uintptr_t)src) & 15) == 0);
    if (fast) {
        for (int i = tid; i < K::INR * K::INW4; i += kFusedThreads) {
            const int r = i / K::INW4, v = i - r * K::INW4;
            const int y = clampi(fold_dec(y_0 + r, Nr), 0, Nr - 1);
            cp_async16(S_in + r * K::INP + 4 * v, src + (size_t)y * Nc + x_al + 4 * v);
        }
        cp_async_wait_all();
    } else if (x_al >= 0 && x_al + K::INW4 * 4 <= Nc) {
        // rows that are not 16-byte aligned (Nc % 4 != 0: odd widths and their descendants) but a tile inside the plane:
        // no column fold, 4-byte asynchronous copies, the whole tile in flight at once (the element-wise loop below
        // made 4095 x 4097 five times slower than 4096 x 4096)
        for (int i = tid; i < K::INR * (K::INW4 * 4); i += kFusedThreads) {
            const int r = i / (K::INW4 * 4), j = i - r * (K::INW4 * 4);
            const int y = clampi(fold_dec(y_0 + r, Nr), 0, Nr - 1);
            cp_async4(S_in + r * K::INP + j, src + (size_t)y * Nc + x_al + j);
        }
        cp_async_wait_all();
    } else {
        for (int i = tid; i < K::INR * (K::INW4 * 4); i += kFusedThreads) {
            const int r = i / (K::INW4 * 4), j = i - r * (K::INW4 * 4);
            const int y = clampi(fold_dec(y_0 + r, Nr), 0, Nr - 1);
            const int x = clampi(fold_dec(x_al + j, Nc), 0, Nc - 1);
            S_in[r * K::INP + j] = src[(size_t)y * Nc + x];
        }
    }
    __syncthreads();

    // ---- row pass (w_kern_forward_pass1, separable.cu:91-131).  A warp covers 8 rows x 4 groups of 4 outputs; the 8
    // lanes of a quarter-warp read 8 different rows (pitch/4 odd => conflict-free 128-bit reads).
    {
        const int lane = tid & 31, warp = tid >> 5;
        constexpr int RB = (K::INR + 7) / 8, QB = TW / 16;
        for (int wt = warp; wt < RB * QB; wt += kFusedThreads / 32) {
            const int rb = wt / QB, qb = wt - rb * QB;
            const int r = rb * 8 + (lane & 7), q = qb * 4 + (lane >> 3);
            if (r < K::INR) {
                float v[K::NV * 4];
                const float4* rp = reinterpret_cast<const float4*>(S_in + r * K::INP + 8 * q);
#pragma unroll
                for (int i = 0; i < K::NV; i++) {
                    const float4 f = rp[i];
                    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
                }
                float lo[4] = {0.f, 0.f, 0.f, 0.f}, hi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < HLEN; j++) {
                    const float kl = t.L[HLEN - 1 - j], kh = t.H[HLEN - 1 - j];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float x = v[K::SH + 2 * k + j];
                        lo[k] = fmaf(x, kl, lo[k]);
                        hi[k] = fmaf(x, kh, hi[k]);
                    }
                }
                *reinterpret_cast<float4*>(S_lo + r * K::MP + 4 * q) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(S_hi + r * K::MP + 4 * q) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            }
        }
    }
    __syncthreads();

    pdl_launch_dependents();   // only the column pass and the stores are left
    // ---- column pass (w_kern_forward_pass2, separable.cu:135-176): A = L_y(lo), H = H_y(lo), V = L_y(hi), D = H_y(hi)
    {
        constexpr int QN = TW / 4, RG = TH / K::RPT;
        const bool vec_ok = ((nc & 3) == 0) && ((s_a & 3) == 0) && ((s_d & 3) == 0) &&
                            (((uintptr_t)A | (uintptr_t)H | (uintptr_t)V | (uintptr_t)D) & 15) == 0;
        for (int task = tid; task < QN * 2 * RG; task += kFusedThreads) {
            const int q = task % QN, rest = task / QN, arr = rest & 1, rg = rest >> 1;
            const float* base = (arr ? S_hi : S_lo) + (2 * rg * K::RPT) * K::MP + 4 * q;
            float4 aL[K::RPT], aH[K::RPT];
#pragma unroll
            for (int o = 0; o < K::RPT; o++) aL[o] = aH[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 2 * K::RPT + K::HALO; i++) {
                const float4 x = *reinterpret_cast<const float4*>(base + i * K::MP);
#pragma unroll
                for (int o = 0; o < K::RPT; o++) {
                    const int j = i - 2 * o;
                    if (j >= 0 && j < HLEN) {
                        const float kl = t.L[HLEN - 1 - j], kh = t.H[HLEN - 1 - j];
                        aL[o].x = fmaf(x.x, kl, aL[o].x); aL[o].y = fmaf(x.y, kl, aL[o].y);
                        aL[o].z = fmaf(x.z, kl, aL[o].z); aL[o].w = fmaf(x.w, kl, aL[o].w);
                        aH[o].x = fmaf(x.x, kh, aH[o].x); aH[o].y = fmaf(x.y, kh, aH[o].y);
                        aH[o].z = fmaf(x.z, kh, aH[o].z); aH[o].w = fmaf(x.w, kh, aH[o].w);
                    }
                }
            }
            float* oL = arr ? V + (size_t)blockIdx.z * s_d : A + (size_t)blockIdx.z * s_a;
            float* oH = (arr ? D : H) + (size_t)blockIdx.z * s_d;
            const int gx = gx0 + 4 * q;
#pragma unroll
            for (int o = 0; o < K::RPT; o++) {
                const int gy = gy0 + rg * K::RPT + o;
                if (gy >= nr || gx >= nc) continue;
                const size_t off = (size_t)gy * nc + gx;
                if (vec_ok) {  // nc % 4 == 0 => the whole vector is in range
                    *reinterpret_cast<float4*>(oL + off) = aL[o];
                    *reinterpret_cast<float4*>(oH + off) = aH[o];
                } else {
                    const float l4[4] = {aL[o].x, aL[o].y, aL[o].z, aL[o].w};
                    const float h4[4] = {aH[o].x, aH[o].y, aH[o].z, aH[o].w};
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (gx + k < nc) {
                            oL[off + k] = l4[k];
                            oH[off + k] = h4[k];
                        }
                }
            }
        }
    }
}

// ================================================================================================== inverse
template <int HLEN, int TWC, int THC>
struct InvCfg {
    static constexpr int H2 = HLEN / 2;
    static constexpr int CC = H2 / 2;                  // synthesis centre, separable.cu:252-264
    static constexpr int SHIFT = (H2 & 1) ? 0 : 1;     // virtual index shift of the even-half-length branch
    static constexpr int TAPS = H2;                    // taps per polyphase branch (even hlen)
    static constexpr int SHC = (4 - (CC & 3)) & 3;     // staging starts at the 16-byte boundary left of cx0 - CC
    static constexpr int INR = THC + SHIFT + TAPS - 1; // staged coefficient rows
    static constexpr int INW = TWC + SHIFT + TAPS - 1 + SHC;
    static constexpr int INW4 = (INW + 3) / 4;
    static constexpr int NV = (SHC + SHIFT + TAPS + 3 + 3) / 4;  // vectors a row-pass task (8 outputs) reads per array
    static constexpr int PW = ((INW4 * 4 > TWC - 4 + NV * 4) ? INW4 * 4 : TWC - 4 + NV * 4);  // pitch (floats)
    static constexpr int RP = 2;                       // output row PAIRS per column-pass task
    static constexpr int CROWS = RP + SHIFT + TAPS - 1; // coefficient rows such a task reads
    static constexpr size_t SMEM = sizeof(float) * (4 * (size_t)INR * PW + 2 * (size_t)(2 * THC) * PW);
    static_assert(THC % RP == 0 && TWC % 4 == 0, "tile shape");
};

template <int HLEN, int TWC, int THC>
__global__ void __launch_bounds__(kFusedThreads, 2)
    k_inv2d(const __grid_constant__ Taps t, const float* __restrict__ A, size_t s_a, const float* __restrict__ H,
            const float* __restrict__ V, const float* __restrict__ D, size_t s_d, float* __restrict__ dst, size_t s_dst,
            int nr, int nc, int Mr, int Mc)
{
    using K = InvCfg<HLEN, TWC, THC>;
    extern __shared__ __align__(16) float smem[];
    float* S_c = smem;                          // 4 coefficient tiles [4][INR][PW]
    float* S_t = smem + 4 * K::INR * K::PW;     // column-synthesised tiles [2][2*THC][PW]
    const int tid = threadIdx.x;
    const int cx0 = blockIdx.x * TWC, cy0 = blockIdx.y * THC;
    const int x_al = cx0 - K::CC - K::SHC;      // multiple of 4
    const int y_0 = cy0 - K::CC;
    const float* srcs[4] = {A + (size_t)blockIdx.z * s_a, H + (size_t)blockIdx.z * s_d, V + (size_t)blockIdx.z * s_d,
                            D + (size_t)blockIdx.z * s_d};
    pdl_wait();   // the kernel that wrote the coefficients (previous level) has completed

    // ---- stage the four coefficient tiles with the single periodic wrap of separable.cu:265-273
    const bool fast = ((nc & 3) == 0) && x_al >= 0 && (x_al + K::INW4 * 4 <= nc) &&
                      (((uintptr_t)srcs[0] | (uintptr_t)srcs[1] | (uintptr_t)srcs[2] | (uintptr_t)srcs[3]) & 15) == 0;
    if (fast) {
        for (int i = tid; i < 4 * K::INR * K::INW4; i += kFusedThreads) {
            const int a = i / (K::INR * K::INW4), rem = i - a * (K::INR * K::INW4);
            const int r = rem / K::INW4, v = rem - r * K::INW4;
            int y = y_0 + r;
            y += (y < 0) ? nr : 0;
            y -= (y >= nr) ? nr : 0;
            y = clampi(y, 0, nr - 1);
            cp_async16(S_c + (a * K::INR + r) * K::PW + 4 * v, srcs[a] + (size_t)y * nc + x_al + 4 * v);
        }
        cp_async_wait_all();
    } else if (x_al >= 0 && x_al + K::INW4 * 4 <= nc) {   // unaligned rows, tile inside the plane: see the forward kernel
        for (int i = tid; i < 4 * K::INR * (K::INW4 * 4); i += kFusedThreads) {
            const int a = i / (K::INR * K::INW4 * 4), rem = i - a * (K::INR * K::INW4 * 4);
            const int r = rem / (K::INW4 * 4), j = rem - r * (K::INW4 * 4);
            int y = y_0 + r;
            y += (y < 0) ? nr : 0;
            y -= (y >= nr) ? nr : 0;
            y = clampi(y, 0, nr - 1);
            cp_async4(S_c + (a * K::INR + r) * K::PW + j, srcs[a] + (size_t)y * nc + x_al + j);
        }
        cp_async_wait_all();
    } else {
        for (int i = tid; i < 4 * K::INR * (K::INW4 * 4); i += kFusedThreads) {
            const int a = i / (K::INR * K::INW4 * 4), rem = i - a * (K::INR * K::INW4 * 4);
            const int r = rem / (K::INW4 * 4), j = rem - r * (K::INW4 * 4);
            int y = y_0 + r, x = x_al + j;
            y += (y < 0) ? nr : 0;
            y -= (y >= nr) ? nr : 0;
            x += (x < 0) ? nc : 0;
            x -= (x >= nc) ? nc : 0;
            y = clampi(y, 0, nr - 1);
            x = clampi(x, 0, nc - 1);
            S_c[(a * K::INR + r) * K::PW + j] = srcs[a][(size_t)y * nc + x];
        }
    }
    __syncthreads();

    // ---- column synthesis (w_kern_inverse_pass1, separable.cu:246-289): t1 = IL_y(A) + IH_y(H), t2 = IL_y(V) + IH_y(D).
    // A task produces 2*RP consecutive output rows x 4 columns of one of (t1, t2).  Output row k of the task (tile row
    // 2*RP*pg + k) has half-index hl = (k+SHIFT)/2 and tap phase off = 1 - ((k+SHIFT)&1), all compile-time.
    {
        constexpr int PG = THC / K::RP;
        for (int task = tid; task < K::INW4 * 2 * PG; task += kFusedThreads) {
            const int v = task % K::INW4, rest = task / K::INW4, pair = rest & 1, pg = rest >> 1;
            const float* lo_t = S_c + ((2 * pair) * K::INR + pg * K::RP) * K::PW + 4 * v;      // A or V
            const float* hi_t = S_c + ((2 * pair + 1) * K::INR + pg * K::RP) * K::PW + 4 * v;  // H or D
            float4 cl[K::CROWS], ch[K::CROWS];
#pragma unroll
            for (int i = 0; i < K::CROWS; i++) {
                cl[i] = *reinterpret_cast<const float4*>(lo_t + i * K::PW);
                ch[i] = *reinterpret_cast<const float4*>(hi_t + i * K::PW);
            }
            float* out = S_t + (pair * 2 * THC + 2 * K::RP * pg) * K::PW + 4 * v;
#pragma unroll
            for (int k = 0; k < 2 * K::RP; k++) {
                const int hl = (k + K::SHIFT) / 2, off = 1 - ((k + K::SHIFT) & 1);
                float4 al = make_float4(0.f, 0.f, 0.f, 0.f), ah = al;
#pragma unroll
                for (int j = 0; j < K::TAPS; j++) {
                    const float kl = t.IL[HLEN - 1 - (2 * j + off)], kh = t.IH[HLEN - 1 - (2 * j + off)];
                    const float4 a = cl[hl + j], h = ch[hl + j];
                    al.x = fmaf(a.x, kl, al.x); al.y = fmaf(a.y, kl, al.y);
                    al.z = fmaf(a.z, kl, al.z); al.w = fmaf(a.w, kl, al.w);
                    ah.x = fmaf(h.x, kh, ah.x); ah.y = fmaf(h.y, kh, ah.y);
                    ah.z = fmaf(h.z, kh, ah.z); ah.w = fmaf(h.w, kh, ah.w);
                }
                *reinterpret_cast<float4*>(out + k * K::PW) =
                    make_float4(__fadd_rn(al.x, ah.x), __fadd_rn(al.y, ah.y), __fadd_rn(al.z, ah.z), __fadd_rn(al.w, ah.w));
            }
        }
    }
    __syncthreads();

    pdl_launch_dependents();   // only the row synthesis and the stores are left
    // ---- row synthesis (w_kern_inverse_pass2, separable.cu:293-328): img = IL_x(t1) + IH_x(t2).  A task produces 8
    // consecutive outputs of one row; output k has hl = (k+SHIFT)/2, off = 1 - ((k+SHIFT)&1).
    {
        constexpr int OG = 2 * TWC / 8;  // groups of 8 outputs per row
        dst += (size_t)blockIdx.z * s_dst;
        const bool vec_ok = ((Mc & 3) == 0) && ((((uintptr_t)dst) & 15) == 0);
        for (int task = tid; task < 2 * THC * OG; task += kFusedThreads) {
            const int og = task % OG, lr = task / OG;
            const int gy = 2 * cy0 + lr, gx = 2 * cx0 + 8 * og;
            if (gy >= Mr || gx >= Mc) continue;
            float v1[K::NV * 4], v2[K::NV * 4];
            const float4* p1 = reinterpret_cast<const float4*>(S_t + lr * K::PW + 4 * og);
            const float4* p2 = reinterpret_cast<const float4*>(S_t + (2 * THC + lr) * K::PW + 4 * og);
#pragma unroll
            for (int i = 0; i < K::NV; i++) {
                const float4 f = p1[i], g = p2[i];
                v1[4 * i] = f.x; v1[4 * i + 1] = f.y; v1[4 * i + 2] = f.z; v1[4 * i + 3] = f.w;
                v2[4 * i] = g.x; v2[4 * i + 1] = g.y; v2[4 * i + 2] = g.z; v2[4 * i + 3] = g.w;
            }
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int hl = (k + K::SHIFT) / 2, off = 1 - ((k + K::SHIFT) & 1);
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int j = 0; j < K::TAPS; j++) {
                    a1 = fmaf(v1[K::SHC + hl + j], t.IL[HLEN - 1 - (2 * j + off)], a1);
                    a2 = fmaf(v2[K::SHC + hl + j], t.IH[HLEN - 1 - (2 * j + off)], a2);
                }
                o[k] = __fadd_rn(a1, a2);
            }
            float* op = dst + (size_t)gy * Mc + gx;
            if (vec_ok && gx + 8 <= Mc) {
                *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (gx + k < Mc) op[k] = o[k];
            }
        }
    }
}

// ================================================================================================ launchers
// Two tile shapes per direction.  The large tile (64 x 32 outputs, 2 CTAs/SM) has the smaller halo; the small tile
// (32 x 16 outputs, ~27 KB of shared memory, 4+ CTAs/SM) is for levels with so few pixels that a grid of large tiles
// would leave most SMs idle: there latency, not traffic, is the cost (PDWT_FUSED_TILE=0|1 forces one of them).
static int tile_choice(long long outputs)
{
    const char* e = getenv("PDWT_FUSED_TILE");
    if (e && *e) return atoi(e) != 0;
    return outputs < 2LL * 148 * 64 * 32 * 4 ? 1 : 0;   // fewer than ~4 waves of large tiles
}

template <int HLEN, int TW, int TH>
static int launch_fwd_t(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                        cudaStream_t s)
{
    using K = FwdCfg<HLEN, TW, TH>;
    PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_fwd2d<HLEN, TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
    dim3 grid(idiv_up(half_up(Nc), TW), idiv_up(half_up(Nr), TH), batch);
    PDWT_PROF(prof_tag("k_fwd2d", Nr, Nc), s);
    PDWT_CUDA(launch_pdl(k_fwd2d<HLEN, TW, TH>, grid, kFusedThreads, K::SMEM, s, t, (const float*)src.p, src.stride, A.p,
                         A.stride, H.p, V.p, D.p, H.stride, Nr, Nc));
    PDWT_LAUNCH_CHECK();
    return 1;
}
template <int HLEN>
static int launch_fwd(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                      cudaStream_t s)
{
    if (tile_choice((long long)half_up(Nr) * half_up(Nc) * batch))
        return launch_fwd_t<HLEN, 32, 16>(t, src, A, H, V, D, Nr, Nc, batch, s);
    return launch_fwd_t<HLEN, 64, 32>(t, src, A, H, V, D, Nr, Nc, batch, s);
}

template <int HLEN, int TWC, int THC>
static int launch_inv_t(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int nr, int nc, int Mr, int Mc,
                        int batch, cudaStream_t s)
{
    using K = InvCfg<HLEN, TWC, THC>;
    PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_inv2d<HLEN, TWC, THC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
    dim3 grid(idiv_up(nc, TWC), idiv_up(nr, THC), batch);
    PDWT_PROF(prof_tag("k_inv2d", Mr, Mc), s);
    PDWT_CUDA(launch_pdl(k_inv2d<HLEN, TWC, THC>, grid, kFusedThreads, K::SMEM, s, t, (const float*)A.p, A.stride,
                         (const float*)H.p, (const float*)V.p, (const float*)D.p, H.stride, dst.p, dst.stride, nr, nc, Mr,
                         Mc));
    PDWT_LAUNCH_CHECK();
    return 1;
}
template <int HLEN>
static int launch_inv(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int nr, int nc, int Mr, int Mc,
                      int batch, cudaStream_t s)
{
    if (tile_choice((long long)nr * nc * batch))
        return launch_inv_t<HLEN, 32, 16>(t, A, H, V, D, dst, nr, nc, Mr, Mc, batch, s);
    return launch_inv_t<HLEN, 64, 32>(t, A, H, V, D, dst, nr, nc, Mr, Mc, batch, s);
}

#define PDWT_HLEN_SWITCH(fn, ...)                   \
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

int f_dwt2_fwd_level(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                     cudaStream_t s)
{
    if (batch > 65535) return 0;
    PDWT_HLEN_SWITCH(launch_fwd, t, src, A, H, V, D, Nr, Nc, batch, s)
}

int f_dwt2_inv_level(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int nr, int nc, int Mr, int Mc,
                     int batch, cudaStream_t s)
{
    if (batch > 65535) return 0;
    PDWT_HLEN_SWITCH(launch_inv, t, A, H, V, D, dst, nr, nc, Mr, Mc, batch, s)
}

bool fused_supports_hlen(int hlen) { return hlen >= 4 && hlen <= kMaxFusedHlen && !(hlen & 1); }

}  // namespace pdwt
