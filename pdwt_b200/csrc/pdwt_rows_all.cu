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
// (more than kRowsAllMax samples), SWT and filter lengths other than even 4..20 keep the per-level kernels.
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

}  // namespace pdwt
