// pdwt_rows_all.cu -- ALL levels of the (batched) 1-D separable DWT in one launch (SURVEY 8 a-3 / a-6 in their 1-D role:
// w_forward_separable_1d, separable.cu:213-236; w_inverse_separable_1d, separable.cu:368-395).
//
// The reference (and pdwt_rows.cu) runs one kernel per level: every intermediate approximation makes a round trip
// through HBM, 16 B per sample and direction for three levels where 8 are compulsory.  A row of a 1-D transform is
// independent of every other row and short enough to live in shared memory, so here ONE CTA takes ONE row through all
// levels: the row is staged once (with the periodic / odd-size extension of separable.cu:114-121 unrolled into a halo,
// so that the tap loops never fold an index), every level reads its input from shared memory and leaves its
// approximation there for the next one; only the detail bands and the last approximation go to global memory
// (forward), only the detail bands and the coarsest approximation are read (inverse).
// Arithmetic is the reference's chain per output (fmaf from 0 in ascending tap order, the inverse adds its two branch
// sums last), so results are bit-identical to the per-level kernels and to the reference.  Rows that do not fit
// (more than kRowsAllMax samples) and filter lengths other than even 4..20 keep the per-level kernels.  The 1-D SWT has
// its own pair of all-level kernels at the end of the file.
#include "pdwt_common.cuh"

namespace pdwt {

namespace {

constexpr int kThreads = 256;
constexpr int kRowsAllMax = 16384;   // samples per row: 3 buffers of half that, 100 KB of shared memory at most

__device__ __forceinline__ void ra_cp4(float* smem_dst, const float* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void ra_cp_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ int ra_wrap1(int i, int N)   // single +-N wrap (separable.cu:265-273), then clamp
{
    i += (i < 0) ? N : 0;
    i -= (i >= N) ? N : 0;
    return i < 0 ? 0 : (i > N - 1 ? N - 1 : i);
}
__device__ __forceinline__ int ra_fold(int i, int N)    // fold_dec, clamped (indices past one period are never used)
{
    i = fold_dec(i, N);
    return i < 0 ? 0 : (i > N - 1 ? N - 1 : i);
}
__device__ __forceinline__ void ra_store4(float* p, const float (&v)[4], int g, int n, bool vec)
{
    if (vec && g + 4 <= n) {
        *reinterpret_cast<float4*>(p + g) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (g + i < n) p[g + i] = v[i];
    }
}

typedef unsigned long long u64;
__device__ __forceinline__ u64 ra_pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void ra_unpack2(u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ra_ffma2(u64 a, u64 b, u64 c)   // two IEEE fp32 FMAs, one issue slot
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// tap pairs in accumulation order: forward (L, H)[hlen-1-j]; inverse (IL, IH)[hlen-1-(2j+off_e)] per output parity e
template <int HLEN>
struct RowTaps {
    float2 lh[HLEN];
    float2 ilh[2][HLEN / 2];
};

struct RowsAll {
    float* d[33];      // d[0] = A_L, d[l] = D_l (l = 1..L): plane bases
    size_t s[33];      // plane strides (floats)
    int n[33];         // n[l] = samples per row at level l (n[0] = Nc)
    unsigned vec;      // bit l: band l may be stored / the image (bit 0... see launcher) with 128-bit vectors
};

__host__ __device__ constexpr int round4(int v) { return (v + 3) & ~3; }

// ------------------------------------------------------------------------------------------------------- forward
// extended layout of a level's input x (N samples): E[e] = x[fold(e - C)], e = 0 .. 2 n + hlen - 3 (n = ceil(N/2)):
// output k reads E[2k .. 2k + hlen - 1]
template <int HLEN>
struct FwdAll {
    static constexpr int C = HLEN / 2 - 1;
    static constexpr int NV = (6 + HLEN + 3) / 4;   // 16-byte vectors of a thread's window (4 outputs)
    __host__ __device__ static int cap(int N) { return 2 * round4(half_up(N)) + 4 * NV; }
};

template <int HLEN>
__global__ void __launch_bounds__(kThreads)
    k_rows_dwt_fwd_all(const __grid_constant__ RowTaps<HLEN> t, const float* __restrict__ img, size_t s_img,
                       const __grid_constant__ RowsAll lv, int L)
{
    using K = FwdAll<HLEN>;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const size_t row = blockIdx.x, pz = blockIdx.y;
    const int N0 = lv.n[0];
    float* cur = smem;                       // extended input of the current level
    float* nxt = smem + K::cap(N0);          // extended input of the next one (capacity cap(n[1]))
    const float* x = img + pz * s_img + row * N0;
    pdl_wait();
    {
        const int E = 2 * half_up(N0) + HLEN - 2;
        for (int i = tid; i < N0; i += kThreads) ra_cp4(cur + K::C + i, x + i);   // the row itself: no fold
        if (tid < K::C) ra_cp4(cur + tid, x + ra_fold(tid - K::C, N0));            // left halo
        for (int e = N0 + K::C + tid; e < E; e += kThreads) ra_cp4(cur + e, x + ra_fold(e - K::C, N0));   // right halo
        ra_cp_wait();
    }
    __syncthreads();
    pdl_launch_dependents();
    for (int l = 1; l <= L; l++) {
        const int N = lv.n[l - 1], n = lv.n[l];
        const bool last = l == L;
        float* hi = lv.d[l] + pz * lv.s[l] + row * n;
        float* lo_g = lv.d[0] + pz * lv.s[0] + row * n;    // only the last level's approximation leaves the SM
        const bool vec_hi = (lv.vec >> l) & 1u, vec_lo = lv.vec & 1u;
        // w_kern_forward_pass1 (separable.cu:91-131): lo/hi[k] = sum_j x[fold(2k - C + j)] * L/H[hlen-1-j]
        for (int k = 4 * tid; k < n; k += 4 * kThreads) {
            float w[4 * K::NV];
#pragma unroll
            for (int i = 0; i < K::NV; i++) {
                const float4 f = *reinterpret_cast<const float4*>(cur + 2 * k + 4 * i);
                w[4 * i] = f.x; w[4 * i + 1] = f.y; w[4 * i + 2] = f.z; w[4 * i + 3] = f.w;
            }
            u64 acc[4] = {0ull, 0ull, 0ull, 0ull};   // (lo, hi) of the 4 outputs
#pragma unroll
            for (int j = 0; j < HLEN; j++) {
                const u64 kk = ra_pack2(t.lh[j].x, t.lh[j].y);
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] = ra_ffma2(ra_pack2(w[2 * i + j], w[2 * i + j]), kk, acc[i]);
            }
            float al[4], ah[4];
#pragma unroll
            for (int i = 0; i < 4; i++) ra_unpack2(acc[i], al[i], ah[i]);
            ra_store4(hi, ah, k, n, vec_hi);
            if (last) {
                ra_store4(lo_g, al, k, n, vec_lo);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (k + i < n) nxt[k + i + K::C] = al[i];   // the next level's sample k sits at E[k + C]
            }
        }
        if (last) break;
        __syncthreads();
        {   // the halo of the next level's extended input: E[e] = lo[fold(e - C)] for e - C outside [0, n)
            const int E = 2 * half_up(n) + HLEN - 2;
            for (int e = tid; e < E; e += kThreads) {
                const int i = e - K::C;
                if (i < 0 || i >= n) nxt[e] = nxt[ra_fold(i, n) + K::C];
            }
        }
        __syncthreads();
        float* tmp = cur;
        cur = nxt;
        nxt = tmp;
        (void)N;
    }
}

// ------------------------------------------------------------------------------------------------------- inverse
// extended layout of a level's coefficient row c (n samples): E[e] = c[wrap(e - CC)]: output pair m reads
// E[m + e*SHIFT + j], j < hlen/2
template <int HLEN>
struct InvAll {
    static constexpr int H2 = HLEN / 2, CC = H2 / 2, SHIFT = (H2 & 1) ? 0 : 1, WIN = H2 + SHIFT;
    static constexpr int NV = (3 + WIN + 3) / 4;
    __host__ __device__ static int cap(int n) { return round4(n) + 4 * NV + 4; }
};

template <int HLEN>
__global__ void __launch_bounds__(kThreads)
    k_rows_dwt_inv_all(const __grid_constant__ RowTaps<HLEN> t, float* __restrict__ img, size_t s_img,
                       const __grid_constant__ RowsAll lv, int L)
{
    using K = InvAll<HLEN>;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const size_t row = blockIdx.x, pz = blockIdx.y;
    const int capb = K::cap(lv.n[1]);
    float* Ea = smem;               // approximation of the current level, extended
    float* Ed = smem + capb;        // detail band of the current level, extended
    float* Eo = smem + 2 * capb;    // the level's output = the next level's approximation, extended
    // extended staging of a band (n samples): E[e] = c[wrap(e - CC)]; only the halo pays for the wrap
    auto stage = [&](float* dst, const float* c, int n, int E) {
        for (int i = tid; i < n && i + K::CC < E; i += kThreads) ra_cp4(dst + K::CC + i, c + i);
        if (tid < K::CC) ra_cp4(dst + tid, c + ra_wrap1(tid - K::CC, n));
        for (int e = n + K::CC + tid; e < E; e += kThreads) ra_cp4(dst + e, c + ra_wrap1(e - K::CC, n));
    };
    pdl_wait();
    stage(Ea, lv.d[0] + pz * lv.s[0] + row * lv.n[L], lv.n[L], min(half_up(lv.n[L - 1]) + K::WIN + 3, capb));
    for (int l = L; l >= 1; l--) {
        const int n = lv.n[l], M = lv.n[l - 1];
        stage(Ed, lv.d[l] + pz * lv.s[l] + row * n, n, min(half_up(M) + K::WIN + 3, capb));
        ra_cp_wait();
        __syncthreads();
        if (l == 1) pdl_launch_dependents();
        float* out_g = img + pz * s_img + row * M;
        const bool vec = lv.vec & 1u;
        // w_kern_inverse_pass2 (separable.cu:293-328): img[2m+e] = sum_j a[..]*IL[hlen-1-(2j+off_e)] + d[..]*IH[..]
        for (int m = 4 * tid; 2 * m < M; m += 4 * kThreads) {
            float w1[4 * K::NV], w2[4 * K::NV];
#pragma unroll
            for (int i = 0; i < K::NV; i++) {
                const float4 f = *reinterpret_cast<const float4*>(Ea + m + 4 * i);
                const float4 g = *reinterpret_cast<const float4*>(Ed + m + 4 * i);
                w1[4 * i] = f.x; w1[4 * i + 1] = f.y; w1[4 * i + 2] = f.z; w1[4 * i + 3] = f.w;
                w2[4 * i] = g.x; w2[4 * i + 1] = g.y; w2[4 * i + 2] = g.z; w2[4 * i + 3] = g.w;
            }
            float o[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int p = q >> 1, e = q & 1, i0 = p + (e ? K::SHIFT : 0);
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int j = 0; j < K::H2; j++) {
                    a1 = fmaf(w1[i0 + j], t.ilh[e][j].x, a1);
                    a2 = fmaf(w2[i0 + j], t.ilh[e][j].y, a2);
                }
                o[q] = __fadd_rn(a1, a2);
            }
            if (l == 1) {
                const float oa[4] = {o[0], o[1], o[2], o[3]}, ob[4] = {o[4], o[5], o[6], o[7]};
                ra_store4(out_g, oa, 2 * m, M, vec);
                ra_store4(out_g, ob, 2 * m + 4, M, vec);
            } else {
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (2 * m + q < M) Eo[2 * m + q + K::CC] = o[q];   // sample g of the next level sits at E[g + CC]
            }
        }
        if (l == 1) break;
        __syncthreads();
        {   // halo of the next level's extended approximation (M samples): E[e] = a[wrap(e - CC)]
            const int E = min(half_up(lv.n[l - 2]) + K::WIN + 3, capb);
            for (int e = tid; e < E; e += kThreads) {
                const int i = e - K::CC;
                if (i < 0 || i >= M) Eo[e] = Eo[ra_wrap1(i, M) + K::CC];
            }
        }
        __syncthreads();
        float* tmp = Ea;
        Ea = Eo;
        Eo = tmp;
    }
}

bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

template <int HLEN>
int launch_fwd_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    using K = FwdAll<HLEN>;
    RowsAll lv;
    memset(&lv, 0, sizeof lv);
    lv.n[0] = Nc;
    for (int l = 1; l <= L; l++) lv.n[l] = half_up(lv.n[l - 1]);
    if (lv.n[L - 1] < HLEN) return 0;   // the single fold must suffice at every level
    for (int l = 0; l <= L; l++) {
        lv.d[l] = bands[l].p;
        lv.s[l] = bands[l].stride;
        const int n = l ? lv.n[l] : lv.n[L];
        if (!(n & 3) && !(bands[l].stride & 3) && al16(bands[l].p)) lv.vec |= 1u << l;
    }
    const size_t smem = sizeof(float) * ((size_t)K::cap(Nc) + K::cap(lv.n[1]));
    if (smem > 160 * 1024) return 0;
    PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_rows_dwt_fwd_all<HLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    RowTaps<HLEN> rt;
    memset(&rt, 0, sizeof rt);
    for (int j = 0; j < HLEN; j++) rt.lh[j] = make_float2(t.L[HLEN - 1 - j], t.H[HLEN - 1 - j]);
    PDWT_PROF(prof_tag("k_rows_dwt_fwd_all", Nr, Nc), s);
    PDWT_CUDA(launch_pdl(k_rows_dwt_fwd_all<HLEN>, dim3(Nr, batch), kThreads, smem, s, rt, (const float*)img.p, img.stride, lv, L));
    PDWT_LAUNCH_CHECK();
    return 1;
}

template <int HLEN>
int launch_inv_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    using K = InvAll<HLEN>;
    RowsAll lv;
    memset(&lv, 0, sizeof lv);
    lv.n[0] = Nc;
    for (int l = 1; l <= L; l++) lv.n[l] = half_up(lv.n[l - 1]);
    if (lv.n[L] < K::WIN) return 0;     // the single wrap must suffice at every level
    for (int l = 0; l <= L; l++) {
        lv.d[l] = bands[l].p;
        lv.s[l] = bands[l].stride;
    }
    if (!(Nc & 3) && !(img.stride & 3) && al16(img.p)) lv.vec |= 1u;
    const size_t smem = sizeof(float) * 3 * (size_t)K::cap(lv.n[1]);
    if (smem > 160 * 1024) return 0;
    PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_rows_dwt_inv_all<HLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    RowTaps<HLEN> rt;
    memset(&rt, 0, sizeof rt);
    for (int e = 0; e < 2; e++) {
        const int off = e ? K::SHIFT : 1 - K::SHIFT;
        for (int j = 0; j < K::H2; j++) rt.ilh[e][j] = make_float2(t.IL[HLEN - 1 - (2 * j + off)], t.IH[HLEN - 1 - (2 * j + off)]);
    }
    PDWT_PROF(prof_tag("k_rows_dwt_inv_all", Nr, Nc), s);
    PDWT_CUDA(launch_pdl(k_rows_dwt_inv_all<HLEN>, dim3(Nr, batch), kThreads, smem, s, rt, img.p, img.stride, lv, L));
    PDWT_LAUNCH_CHECK();
    return 1;
}


// ================================================================================================ SWT, all levels
// The undecimated transform of a row keeps its length, so two row buffers suffice: level l (taps f = 2^(l-1) apart) reads
// buffer A and leaves its approximation in buffer B, the details go straight to global memory.  A buffer holds the row in
// EXTENDED form -- E[e] = x[wrap(e - c)], e = 0 .. N + (hlen-1) f - 1, with c the level's centre -- so the tap loops
// never fold an index: after a level has written the core of the next buffer, a short second loop copies the
// (hlen-1) f halo entries of the NEXT level around it.  Consecutive lanes take consecutive samples: every shared-memory
// access is conflict-free whatever the dilation.  Arithmetic: forward = one fmaf chain from 0 over ascending j
// (separable.cu:427-445); inverse = round(v*k), exact halving, add, the two branch sums added last (separable.cu:615-625),
// scalar multiply and add as in pdwt_rows.cu (the taps are halved beforehand, which is exact).  Bit-identical to the
// per-level kernels.
template <int HLEN>
struct SwtRowTaps {
    float2 lh[HLEN];    // (L, H)[hlen-1-j]
    float2 ilh[HLEN];   // (IL, IH)[hlen-1-j] / 2 (exact)
    float2 one;         // (1, 1), opaque to the compiler
};

__device__ __forceinline__ u64 ra_fmul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 ra_add2_exact(u64 a, u64 ones, u64 b)   // a * (1, 1) + b: add.rn on both halves
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(ones), "l"(b));
    return d;
}
// halo of the extended layout with centre c and dilation f around a core that sits at [c, c + N): entries e < c are
// x[e - c + N], entries e >= c + N are x[e - c - N]
template <typename T>
__device__ __forceinline__ void swt_fill_halo(T* E, int N, int c, int ext, int tid)
{
    for (int e = tid; e < ext - N; e += kThreads) {
        const int dst = e < c ? e : e + N;             // the halo positions in ascending order
        E[dst] = E[dst < c ? dst + N : dst - N];
    }
}

// One level for 4 consecutive outputs per thread.  FM = 1 / 2: taps f = 1 / 2 apart come out of one register window of
// 128-bit loads (hlen + 3 or 2 hlen + 2 samples serve 4 hlen taps); FM = 4: f is a multiple of 4 and every tap is one
// aligned 128-bit load (the four outputs share nothing).  E index of tap j of output g is g + j f (extended layout).
template <int HLEN, int FM>
__device__ __forceinline__ void swt_fwd_level(const SwtRowTaps<HLEN>& rt, const float* Ein, float* Eout, float* dD, float* dA,
                                              int N, int f, int cn, bool last, bool vecD, bool vecA, int tid)
{
    constexpr int NV = (FM * (HLEN - 1) + 4 + 3) / 4;
    for (int g = 4 * tid; g < N; g += 4 * kThreads) {
        u64 acc[4] = {0ull, 0ull, 0ull, 0ull};
        if (FM == 4) {
#pragma unroll
            for (int j = 0; j < HLEN; j++) {
                const float4 v = *reinterpret_cast<const float4*>(Ein + g + j * f);
                const u64 k = ra_pack2(rt.lh[j].x, rt.lh[j].y);
                acc[0] = ra_ffma2(ra_pack2(v.x, v.x), k, acc[0]);
                acc[1] = ra_ffma2(ra_pack2(v.y, v.y), k, acc[1]);
                acc[2] = ra_ffma2(ra_pack2(v.z, v.z), k, acc[2]);
                acc[3] = ra_ffma2(ra_pack2(v.w, v.w), k, acc[3]);
            }
        } else {
            float w[4 * NV];
#pragma unroll
            for (int i = 0; i < NV; i++) {
                const float4 v = *reinterpret_cast<const float4*>(Ein + g + 4 * i);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < HLEN; j++) {
                const u64 k = ra_pack2(rt.lh[j].x, rt.lh[j].y);
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] = ra_ffma2(ra_pack2(w[FM * j + i], w[FM * j + i]), k, acc[i]);
            }
        }
        float a[4], d[4];
#pragma unroll
        for (int i = 0; i < 4; i++) ra_unpack2(acc[i], a[i], d[i]);
        ra_store4(dD, d, g, N, vecD);
        if (last) ra_store4(dA, a, g, N, vecA);
        else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (g + i < N) Eout[cn + g + i] = a[i];
        }
    }
}

template <int HLEN>
__global__ void __launch_bounds__(kThreads, 3)
    k_rows_swt_fwd_all(const __grid_constant__ SwtRowTaps<HLEN> rt, const float* __restrict__ img, size_t s_img,
                       const __grid_constant__ RowsAll lv, int L, int pitch)
{
    extern __shared__ __align__(16) float smem[];
    const int N = lv.n[0], tid = threadIdx.x;
    const size_t row = blockIdx.x, pz = blockIdx.y;
    float* Ein = smem;
    float* Eout = smem + pitch;
    pdl_wait();
    {   // level 1: f = 1
        const int c = HLEN / 2 - 1;
        const float* p = img + pz * s_img + row * N;
        for (int x = tid; x < N; x += kThreads) ra_cp4(Ein + c + x, p + x);
        ra_cp_wait();
        __syncthreads();
        swt_fill_halo(Ein, N, c, N + (HLEN - 1), tid);
        __syncthreads();
    }
    for (int l = 1; l <= L; l++) {
        const int f = 1 << (l - 1);
        const int cn = (HLEN / 2 - 1) * 2 * f;         // centre of the next level
        float* dD = lv.d[l] + pz * lv.s[l] + row * N;
        float* dA = lv.d[0] + pz * lv.s[0] + row * N;
        const bool last = l == L, vecD = (lv.vec >> l) & 1, vecA = lv.vec & 1;
        if (last) pdl_launch_dependents();
        if (l == 1) swt_fwd_level<HLEN, 1>(rt, Ein, Eout, dD, dA, N, f, cn, last, vecD, vecA, tid);
        else if (l == 2) swt_fwd_level<HLEN, 2>(rt, Ein, Eout, dD, dA, N, f, cn, last, vecD, vecA, tid);
        else swt_fwd_level<HLEN, 4>(rt, Ein, Eout, dD, dA, N, f, cn, last, vecD, vecA, tid);
        if (last) break;
        __syncthreads();
        swt_fill_halo(Eout, N, cn, N + (HLEN - 1) * 2 * f, tid);
        __syncthreads();
        float* t = Ein; Ein = Eout; Eout = t;
    }
}

// inverse level: approximation and details in two separate extended arrays (128-bit loads of consecutive lanes stay
// contiguous: interleaved (A, D) pairs put the lanes 32 bytes apart and every load cost two wavefronts -- measured 189 us
// for the launch instead of 100); a tap is round(v * k/2) added to its branch sum, scalar, the two sums added last
template <int HLEN, int FM>
__device__ __forceinline__ void swt_inv_level(const SwtRowTaps<HLEN>& rt, const u64 ones, const float* EA, const float* ED, float* EoutA,
                                              float* dI, int N, int f, int cn, bool last, bool vecI, int tid)
{
    constexpr int NV = (FM * (HLEN - 1) + 4 + 3) / 4;
    for (int g = 4 * tid; g < N; g += 4 * kThreads) {
        float r1[4] = {0.f, 0.f, 0.f, 0.f}, r2[4] = {0.f, 0.f, 0.f, 0.f};
        if (FM == 4) {
            // packed over NEIGHBOURING outputs (the halves of a 128-bit load are register pairs already): FMUL2 by (k, k),
            // then the add as an FFMA2 by an opaque (1, 1) -- the product keeps its own rounding (see pdwt_swt.cu)
            u64 q1[2] = {0ull, 0ull}, q2[2] = {0ull, 0ull};
#pragma unroll
            for (int j = 0; j < HLEN; j++) {
                const float4 a = *reinterpret_cast<const float4*>(EA + g + j * f);
                const float4 d = *reinterpret_cast<const float4*>(ED + g + j * f);
                const u64 ka = ra_pack2(rt.ilh[j].x, rt.ilh[j].x), kd = ra_pack2(rt.ilh[j].y, rt.ilh[j].y);
                q1[0] = ra_add2_exact(ra_fmul2(ra_pack2(a.x, a.y), ka), ones, q1[0]);
                q1[1] = ra_add2_exact(ra_fmul2(ra_pack2(a.z, a.w), ka), ones, q1[1]);
                q2[0] = ra_add2_exact(ra_fmul2(ra_pack2(d.x, d.y), kd), ones, q2[0]);
                q2[1] = ra_add2_exact(ra_fmul2(ra_pack2(d.z, d.w), kd), ones, q2[1]);
            }
            ra_unpack2(q1[0], r1[0], r1[1]);
            ra_unpack2(q1[1], r1[2], r1[3]);
            ra_unpack2(q2[0], r2[0], r2[1]);
            ra_unpack2(q2[1], r2[2], r2[3]);
        } else {
            float wa[4 * NV], wd[4 * NV];
#pragma unroll
            for (int i = 0; i < NV; i++) {
                const float4 a = *reinterpret_cast<const float4*>(EA + g + 4 * i);
                const float4 d = *reinterpret_cast<const float4*>(ED + g + 4 * i);
                wa[4 * i] = a.x; wa[4 * i + 1] = a.y; wa[4 * i + 2] = a.z; wa[4 * i + 3] = a.w;
                wd[4 * i] = d.x; wd[4 * i + 1] = d.y; wd[4 * i + 2] = d.z; wd[4 * i + 3] = d.w;
            }
#pragma unroll
            for (int j = 0; j < HLEN; j++) {
                const float ka = rt.ilh[j].x, kd = rt.ilh[j].y;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    r1[i] = __fadd_rn(r1[i], __fmul_rn(wa[FM * j + i], ka));
                    r2[i] = __fadd_rn(r2[i], __fmul_rn(wd[FM * j + i], kd));
                }
            }
        }
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = __fadd_rn(r1[i], r2[i]);
        if (last) ra_store4(dI, o, g, N, vecI);
        else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (g + i < N) EoutA[cn + g + i] = o[i];
        }
    }
}

template <int HLEN>
__global__ void __launch_bounds__(kThreads, 3)
    k_rows_swt_inv_all(const __grid_constant__ SwtRowTaps<HLEN> rt, float* __restrict__ img, size_t s_img,
                       const __grid_constant__ RowsAll lv, int L, int pitch)
{
    extern __shared__ __align__(16) float smem[];
    const int N = lv.n[0], tid = threadIdx.x;
    const size_t row = blockIdx.x, pz = blockIdx.y;
    float *EA = smem, *ED = smem + pitch, *EoutA = smem + 2 * pitch, *EoutD = smem + 3 * pitch;   // extended rows
    const u64 ones = ra_pack2(rt.one.x, rt.one.y);
    pdl_wait();
    {   // the coarsest approximation and its details
        const int c = (HLEN / 2) << (L - 1);
        const float* pa = lv.d[0] + pz * lv.s[0] + row * N;
        const float* pd = lv.d[L] + pz * lv.s[L] + row * N;
        for (int x = tid; x < N; x += kThreads) {
            ra_cp4(EA + c + x, pa + x);
            ra_cp4(ED + c + x, pd + x);
        }
    }
    for (int l = L; l >= 1; l--) {
        const int f = 1 << (l - 1);
        const int c = (HLEN / 2) * f, cn = (HLEN / 2) * (f >> 1);
        ra_cp_wait();      // this level's details (and, the first time, the approximation) have landed
        __syncthreads();   // ... for every thread, and the previous level's stores into the approximation are complete
        swt_fill_halo(EA, N, c, N + (HLEN - 1) * f, tid);
        swt_fill_halo(ED, N, c, N + (HLEN - 1) * f, tid);
        if (l > 1) {       // the next level's details travel while this one computes
            const float* p = lv.d[l - 1] + pz * lv.s[l - 1] + row * N;
            for (int x = tid; x < N; x += kThreads) ra_cp4(EoutD + cn + x, p + x);
        }
        __syncthreads();
        const bool last = l == 1;
        if (last) pdl_launch_dependents();
        float* dI = img + pz * s_img + row * N;
        const bool vecI = lv.vec & 1;
        if (l == 1) swt_inv_level<HLEN, 1>(rt, ones, EA, ED, EoutA, dI, N, f, cn, last, vecI, tid);
        else if (l == 2) swt_inv_level<HLEN, 2>(rt, ones, EA, ED, EoutA, dI, N, f, cn, last, vecI, tid);
        else swt_inv_level<HLEN, 4>(rt, ones, EA, ED, EoutA, dI, N, f, cn, last, vecI, tid);
        if (last) break;
        float* t = EA; EA = EoutA; EoutA = t;
        t = ED; ED = EoutD; EoutD = t;
    }
}

template <int HLEN>
int launch_swt_all(bool inverse, const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    const int fmax = 1 << (L - 1);
    if ((long long)(HLEN - 1) * fmax >= Nc) return 0;   // the single wrap must suffice at every level
    const int pitch = round4(Nc + (HLEN - 1) * fmax) + 8;   // + slack: the window of a row's last, partial group of 4 reads past the end
    const size_t smem = sizeof(float) * 2 * (size_t)pitch * (inverse ? 2 : 1);
    if (smem > 160 * 1024) return 0;
    RowsAll lv;
    memset(&lv, 0, sizeof lv);
    lv.n[0] = Nc;
    for (int l = 0; l <= L; l++) {
        lv.d[l] = bands[l].p;
        lv.s[l] = bands[l].stride;
        if (!inverse && !(Nc & 3) && !(bands[l].stride & 3) && al16(bands[l].p)) lv.vec |= 1u << l;
    }
    if (inverse && !(Nc & 3) && !(img.stride & 3) && al16(img.p)) lv.vec |= 1u;
    SwtRowTaps<HLEN> rt;
    rt.one = make_float2(1.0f, 1.0f);
    for (int j = 0; j < HLEN; j++) {
        rt.lh[j] = make_float2(t.L[HLEN - 1 - j], t.H[HLEN - 1 - j]);
        rt.ilh[j] = make_float2(t.IL[HLEN - 1 - j] * 0.5f, t.IH[HLEN - 1 - j] * 0.5f);
    }
    if (inverse) {
        PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_rows_swt_inv_all<HLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        PDWT_PROF(prof_tag("k_rows_swt_inv_all", Nr, Nc), s);
        PDWT_CUDA(launch_pdl(k_rows_swt_inv_all<HLEN>, dim3(Nr, batch), kThreads, smem, s, rt, img.p, img.stride, lv, L, pitch));
    } else {
        PDWT_ONCE_PER_DEVICE(cudaFuncSetAttribute(k_rows_swt_fwd_all<HLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        PDWT_PROF(prof_tag("k_rows_swt_fwd_all", Nr, Nc), s);
        PDWT_CUDA(launch_pdl(k_rows_swt_fwd_all<HLEN>, dim3(Nr, batch), kThreads, smem, s, rt, (const float*)img.p, img.stride, lv, L, pitch));
    }
    PDWT_LAUNCH_CHECK();
    return 1;
}
template <int HLEN>
int launch_swt_fwd_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    return launch_swt_all<HLEN>(false, t, img, bands, Nr, Nc, L, batch, s);
}
template <int HLEN>
int launch_swt_inv_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    return launch_swt_all<HLEN>(true, t, img, bands, Nr, Nc, L, batch, s);
}

}  // namespace

#define PDWT_ROWSALL_SWITCH(fn)                                                   \
    switch (t.hlen) {                                                             \
        case 4: return fn<4>(t, img, bands, Nr, Nc, L, batch, s);                 \
        case 6: return fn<6>(t, img, bands, Nr, Nc, L, batch, s);                 \
        case 8: return fn<8>(t, img, bands, Nr, Nc, L, batch, s);                 \
        case 10: return fn<10>(t, img, bands, Nr, Nc, L, batch, s);               \
        case 12: return fn<12>(t, img, bands, Nr, Nc, L, batch, s);               \
        case 14: return fn<14>(t, img, bands, Nr, Nc, L, batch, s);               \
        case 16: return fn<16>(t, img, bands, Nr, Nc, L, batch, s);               \
        case 18: return fn<18>(t, img, bands, Nr, Nc, L, batch, s);               \
        case 20: return fn<20>(t, img, bands, Nr, Nc, L, batch, s);               \
        default: return 0;                                                        \
    }

// bands[0] = A_L, bands[l] = D_l (l = 1..L), each with its plane stride; rows of Nc samples, Nr rows per plane.
// 1 = handled, 0 = not covered (the caller runs the per-level kernels), < 0 = error
int r_dwt1_fwd_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    if (L < 2 || L > 32 || Nc > kRowsAllMax || Nr > 0x7fffffff || batch > 65535) return 0;
    PDWT_ROWSALL_SWITCH(launch_fwd_all)
}
int r_dwt1_inv_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    if (L < 2 || L > 32 || Nc > kRowsAllMax || batch > 65535) return 0;
    PDWT_ROWSALL_SWITCH(launch_inv_all)
}

// the same for the 1-D SWT (w_forward_swt_separable_1d / w_inverse_swt_separable_1d, separable.cu:519-537, 653-672): bands[l] are
// full-length rows; 0 when a level's dilated halo would wrap more than once or the row does not fit in shared memory
int r_swt1_fwd_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    if (L < 2 || L > 16 || Nc > kRowsAllMax || Nr > 0x7fffffff || batch > 65535) return 0;
    PDWT_ROWSALL_SWITCH(launch_swt_fwd_all)
}
int r_swt1_inv_all(const Taps& t, Plane2 img, const Plane2* bands, int Nr, int Nc, int L, int batch, cudaStream_t s)
{
    if (L < 2 || L > 16 || Nc > kRowsAllMax || Nr > 0x7fffffff || batch > 65535) return 0;
    PDWT_ROWSALL_SWITCH(launch_swt_inv_all)
}

}  // namespace pdwt
