// pdwt_generic.cu -- shape-generic kernels: every transform pass of the reference for ANY plane size (odd sizes
// included), any filter length up to 40 taps and any batch.  One thread per output element, 32x8 thread blocks
// so that warps read and write whole 128-byte lines along x.  These are the safety net and the path for the
// configurations the fused kernels (pdwt_fused.cu) do not specialise; the arithmetic of BOTH is the same:
// one fmaf per tap, taps visited in ascending j, starting from 0 -- i.e. bit-identical to the reference kernels,
// whose `res += v * k` nvcc contracts to the same FFMA chain.
//
// Filter taps are read from the launch parameters (`__grid_constant__ Taps`), i.e. constant bank 0, with a
// run-time index -- the per-instance equivalent of the reference's c_kern_* constant arrays (common.h:28-36).
#include "pdwt_common.cuh"

namespace pdwt {

#define GX 32
#define GY 8
#define IDX2                                            \
    const int gx = blockIdx.x * GX + threadIdx.x;       \
    const int gy = blockIdx.y * GY + threadIdx.y;       \
    const size_t pz = blockIdx.z

static inline dim3 grid2(int nx, int ny, int batch) { return dim3(idiv_up(nx, GX), idiv_up(ny, GY), batch); }
static const dim3 kBlock(GX, GY, 1);

// ---------------------------------------------------------------------------------------------- separable DWT
// w_kern_forward_pass1, separable.cu:91-131
__global__ void __launch_bounds__(GX* GY) k_fwd_rows(const __grid_constant__ Taps t, const float* __restrict__ img,
                                                     size_t s_img, float* __restrict__ lo, size_t s_lo,
                                                     float* __restrict__ hi, size_t s_hi, int Nr, int Nc)
{
    IDX2;
    const int n = half_up(Nc);
    if (gy >= Nr || gx >= n) return;
    const int hlen = t.hlen, c = centre_fwd(hlen);
    const float* x = img + pz * s_img + (size_t)gy * Nc;
    float al = 0.f, ah = 0.f;
    for (int j = 0; j < hlen; j++) {
        const float v = x[fold_dec(2 * gx - c + j, Nc)];
        al = fmaf(v, t.L[hlen - 1 - j], al);
        ah = fmaf(v, t.H[hlen - 1 - j], ah);
    }
    lo[pz * s_lo + (size_t)gy * n + gx] = al;
    hi[pz * s_hi + (size_t)gy * n + gx] = ah;
}

// w_kern_forward_pass2, separable.cu:135-176.  Nc = width of t1/t2.
__global__ void __launch_bounds__(GX* GY)
    k_fwd_cols(const __grid_constant__ Taps t, const float* __restrict__ t1, const float* __restrict__ t2, size_t s_t,
               float* __restrict__ A, float* __restrict__ H, float* __restrict__ V, float* __restrict__ D, size_t s_a,
               size_t s_d, int Nr, int Nc)
{
    IDX2;
    const int n = half_up(Nr);
    if (gy >= n || gx >= Nc) return;
    const int hlen = t.hlen, c = centre_fwd(hlen);
    const float* p1 = t1 + pz * s_t + gx;
    const float* p2 = t2 + pz * s_t + gx;
    float ra = 0.f, rh = 0.f, rv = 0.f, rd = 0.f;
    for (int j = 0; j < hlen; j++) {
        const size_t y = (size_t)fold_dec(2 * gy - c + j, Nr) * Nc;
        const float v1 = p1[y], v2 = p2[y];
        const float kl = t.L[hlen - 1 - j], kh = t.H[hlen - 1 - j];
        ra = fmaf(v1, kl, ra);
        rh = fmaf(v1, kh, rh);
        rv = fmaf(v2, kl, rv);
        rd = fmaf(v2, kh, rd);
    }
    const size_t o = (size_t)gy * Nc + gx;
    A[pz * s_a + o] = ra;
    H[pz * s_d + o] = rh;
    V[pz * s_d + o] = rv;
    D[pz * s_d + o] = rd;
}

// w_kern_inverse_pass1, separable.cu:246-289.  n = coefficient rows, M = output rows, Nc = width.
__global__ void __launch_bounds__(GX* GY)
    k_inv_cols(const __grid_constant__ Taps t, const float* __restrict__ A, const float* __restrict__ H,
               const float* __restrict__ V, const float* __restrict__ D, size_t s_a, size_t s_d,
               float* __restrict__ t1, float* __restrict__ t2, size_t s_t, int n, int Nc, int M)
{
    IDX2;
    if (gy >= M || gx >= Nc) return;
    const int hlen = t.hlen;
    const SynGeom sg = syn_geometry(hlen);
    const int g = gy + sg.shift, half = g / 2, off = 1 - (g & 1);
    const int j_lo = sg.c - half, j_hi = n - 1 - half + sg.c;
    const float *pa = A + pz * s_a + gx, *ph = H + pz * s_d + gx, *pv = V + pz * s_d + gx, *pd = D + pz * s_d + gx;
    float ra = 0.f, rh = 0.f, rv = 0.f, rd = 0.f;
    for (int j = 0; j < sg.taps; j++) {
        int y = half - sg.c + j;
        if (j < j_lo) y += n;
        if (j > j_hi) y -= n;
        const size_t o = (size_t)y * Nc;
        const float kl = t.IL[hlen - 1 - (2 * j + off)], kh = t.IH[hlen - 1 - (2 * j + off)];
        ra = fmaf(pa[o], kl, ra);
        rh = fmaf(ph[o], kh, rh);
        rv = fmaf(pv[o], kl, rv);
        rd = fmaf(pd[o], kh, rd);
    }
    t1[pz * s_t + (size_t)gy * Nc + gx] = ra + rh;
    t2[pz * s_t + (size_t)gy * Nc + gx] = rv + rd;
}

// w_kern_inverse_pass2, separable.cu:293-328.  n = width of t1/t2, M = output width.
__global__ void __launch_bounds__(GX* GY)
    k_inv_rows(const __grid_constant__ Taps t, const float* __restrict__ t1, size_t s_1, const float* __restrict__ t2,
               size_t s_2, float* __restrict__ img, size_t s_img, int Nr, int n, int M)
{
    IDX2;
    if (gy >= Nr || gx >= M) return;
    const int hlen = t.hlen;
    const SynGeom sg = syn_geometry(hlen);
    const int g = gx + sg.shift, half = g / 2, off = 1 - (g & 1);
    const int j_lo = sg.c - half, j_hi = n - 1 - half + sg.c;
    const float* p1 = t1 + pz * s_1 + (size_t)gy * n;
    const float* p2 = t2 + pz * s_2 + (size_t)gy * n;
    float a1 = 0.f, a2 = 0.f;
    for (int j = 0; j < sg.taps; j++) {
        int x = half - sg.c + j;
        if (j < j_lo) x += n;
        if (j > j_hi) x -= n;
        a1 = fmaf(p1[x], t.IL[hlen - 1 - (2 * j + off)], a1);
        a2 = fmaf(p2[x], t.IH[hlen - 1 - (2 * j + off)], a2);
    }
    img[pz * s_img + (size_t)gy * M + gx] = a1 + a2;
}

// ---------------------------------------------------------------------------------------------- separable SWT
// w_kern_forward_swt_pass1, separable.cu:409-448
__global__ void __launch_bounds__(GX* GY)
    k_swt_fwd_rows(const __grid_constant__ Taps t, const float* __restrict__ img, size_t s_img, float* __restrict__ lo,
                   size_t s_lo, float* __restrict__ hi, size_t s_hi, int Nr, int Nc, int fac)
{
    IDX2;
    if (gy >= Nr || gx >= Nc) return;
    const int hlen = t.hlen, c = centre_fwd(hlen) * fac;
    const float* x = img + pz * s_img + (size_t)gy * Nc;
    float al = 0.f, ah = 0.f;
    for (int j = 0; j < hlen; j++) {
        const float v = x[fold_swt(gx, j * fac, c, Nc)];
        al = fmaf(v, t.L[hlen - 1 - j], al);
        ah = fmaf(v, t.H[hlen - 1 - j], ah);
    }
    lo[pz * s_lo + (size_t)gy * Nc + gx] = al;
    hi[pz * s_hi + (size_t)gy * Nc + gx] = ah;
}

// w_kern_forward_swt_pass2, separable.cu:452-493
__global__ void __launch_bounds__(GX* GY)
    k_swt_fwd_cols(const __grid_constant__ Taps t, const float* __restrict__ t1, const float* __restrict__ t2,
                   size_t s_t, float* __restrict__ A, float* __restrict__ H, float* __restrict__ V,
                   float* __restrict__ D, size_t s_a, size_t s_d, int Nr, int Nc, int fac)
{
    IDX2;
    if (gy >= Nr || gx >= Nc) return;
    const int hlen = t.hlen, c = centre_fwd(hlen) * fac;
    const float* p1 = t1 + pz * s_t + gx;
    const float* p2 = t2 + pz * s_t + gx;
    float ra = 0.f, rh = 0.f, rv = 0.f, rd = 0.f;
    for (int j = 0; j < hlen; j++) {
        const size_t y = (size_t)fold_swt(gy, j * fac, c, Nr) * Nc;
        const float v1 = p1[y], v2 = p2[y];
        const float kl = t.L[hlen - 1 - j], kh = t.H[hlen - 1 - j];
        ra = fmaf(v1, kl, ra);
        rh = fmaf(v1, kh, rh);
        rv = fmaf(v2, kl, rv);
        rd = fmaf(v2, kh, rd);
    }
    const size_t o = (size_t)gy * Nc + gx;
    A[pz * s_a + o] = ra;
    H[pz * s_d + o] = rh;
    V[pz * s_d + o] = rv;
    D[pz * s_d + o] = rd;
}

// w_kern_inverse_swt_pass1, separable.cu:553-589: `res += v * k / 2` = round(v*k), halve (exact), add.
__global__ void __launch_bounds__(GX* GY)
    k_swt_inv_cols(const __grid_constant__ Taps t, const float* __restrict__ A, const float* __restrict__ H,
                   const float* __restrict__ V, const float* __restrict__ D, size_t s_a, size_t s_d,
                   float* __restrict__ t1, float* __restrict__ t2, size_t s_t, int Nr, int Nc, int fac)
{
    IDX2;
    if (gy >= Nr || gx >= Nc) return;
    const int hlen = t.hlen, c = (hlen / 2) * fac, taps = swt_inv_taps(hlen);
    const float *pa = A + pz * s_a + gx, *ph = H + pz * s_d + gx, *pv = V + pz * s_d + gx, *pd = D + pz * s_d + gx;
    float ra = 0.f, rh = 0.f, rv = 0.f, rd = 0.f;
    for (int j = 0; j < taps; j++) {
        const size_t o = (size_t)fold_swt(gy, j * fac, c, Nr) * Nc;
        const float kl = t.IL[hlen - 1 - j], kh = t.IH[hlen - 1 - j];
        ra = __fadd_rn(ra, __fmul_rn(pa[o], kl) * 0.5f);
        rh = __fadd_rn(rh, __fmul_rn(ph[o], kh) * 0.5f);
        rv = __fadd_rn(rv, __fmul_rn(pv[o], kl) * 0.5f);
        rd = __fadd_rn(rd, __fmul_rn(pd[o], kh) * 0.5f);
    }
    t1[pz * s_t + (size_t)gy * Nc + gx] = ra + rh;
    t2[pz * s_t + (size_t)gy * Nc + gx] = rv + rd;
}

// w_kern_inverse_swt_pass2, separable.cu:593-626
__global__ void __launch_bounds__(GX* GY)
    k_swt_inv_rows(const __grid_constant__ Taps t, const float* __restrict__ t1, size_t s_1,
                   const float* __restrict__ t2, size_t s_2, float* __restrict__ img, size_t s_img, int Nr, int Nc,
                   int fac)
{
    IDX2;
    if (gy >= Nr || gx >= Nc) return;
    const int hlen = t.hlen, c = (hlen / 2) * fac, taps = swt_inv_taps(hlen);
    const float* p1 = t1 + pz * s_1 + (size_t)gy * Nc;
    const float* p2 = t2 + pz * s_2 + (size_t)gy * Nc;
    float a1 = 0.f, a2 = 0.f;
    for (int j = 0; j < taps; j++) {
        const int x = fold_swt(gx, j * fac, c, Nc);
        a1 = __fadd_rn(a1, __fmul_rn(p1[x], t.IL[hlen - 1 - j]) * 0.5f);
        a2 = __fadd_rn(a2, __fmul_rn(p2[x], t.IH[hlen - 1 - j]) * 0.5f);
    }
    img[pz * s_img + (size_t)gy * Nc + gx] = a1 + a2;
}

// ------------------------------------------------------------------------------------------------------- Haar
// kern_haar2d_fwd, haar.cu:10-37: 0.5*((a+c)+(b+d)) etc. in exactly this association (x0.5 is exact).
__global__ void __launch_bounds__(GX* GY)
    k_haar2d_fwd(const float* __restrict__ img, size_t s_img, float* __restrict__ A, size_t s_a, float* __restrict__ H,
                 float* __restrict__ V, float* __restrict__ D, size_t s_d, int Nr, int Nc)
{
    IDX2;
    const int nr = half_up(Nr), nc = half_up(Nc);
    if (gy >= nr || gx >= nc) return;
    const int x0 = 2 * gx, y0 = 2 * gy;
    int x1 = x0 + 1, y1 = y0 + 1;
    if ((Nc & 1) && x1 == Nc) x1--;
    if ((Nr & 1) && y1 == Nr) y1--;
    const float* p = img + pz * s_img;
    const float a = p[(size_t)y0 * Nc + x0], b = p[(size_t)y0 * Nc + x1];
    const float c = p[(size_t)y1 * Nc + x0], d = p[(size_t)y1 * Nc + x1];
    const float sac = __fadd_rn(a, c), sbd = __fadd_rn(b, d), dac = __fsub_rn(a, c), dbd = __fsub_rn(b, d);
    const size_t o = (size_t)gy * nc + gx;
    A[pz * s_a + o] = 0.5f * __fadd_rn(sac, sbd);
    V[pz * s_d + o] = 0.5f * __fsub_rn(sac, sbd);
    H[pz * s_d + o] = 0.5f * __fadd_rn(dac, dbd);
    D[pz * s_d + o] = 0.5f * __fsub_rn(dac, dbd);
}

// kern_haar2d_inv, haar.cu:41-58.  (a,b,c,d) = (A,V,H,D)[gy/2, gx/2]; output parity selects the butterfly.
__global__ void __launch_bounds__(GX* GY)
    k_haar2d_inv(float* __restrict__ img, size_t s_img, const float* __restrict__ A, size_t s_a,
                 const float* __restrict__ H, const float* __restrict__ V, const float* __restrict__ D, size_t s_d,
                 int Nc, int Nr2, int Nc2)
{
    IDX2;
    if (gy >= Nr2 || gx >= Nc2) return;
    const size_t i = (size_t)(gy >> 1) * Nc + (gx >> 1);
    const float a = A[pz * s_a + i], b = V[pz * s_d + i], c = H[pz * s_d + i], d = D[pz * s_d + i];
    const float u = (gy & 1) ? __fsub_rn(a, c) : __fadd_rn(a, c);
    const float w = (gy & 1) ? __fsub_rn(b, d) : __fadd_rn(b, d);
    img[pz * s_img + (size_t)gy * Nc2 + gx] = 0.5f * ((gx & 1) ? __fsub_rn(u, w) : __fadd_rn(u, w));
}

// Vectorised Haar level for even sizes with nc % 4 == 0 and 16-byte aligned planes: a thread turns 2 x 8 pixels into 4
// outputs of each sub-band (forward) or back (inverse) -- 128-bit loads and stores only, same expressions and
// association as above (haar.cu:27-35, 45-54).  Pure streaming: 8 B per pixel and level.
__device__ __forceinline__ void haar_bfly(float a, float b, float c, float d, float& A, float& V, float& H, float& D)
{
    const float sac = __fadd_rn(a, c), sbd = __fadd_rn(b, d), dac = __fsub_rn(a, c), dbd = __fsub_rn(b, d);
    A = 0.5f * __fadd_rn(sac, sbd);
    V = 0.5f * __fsub_rn(sac, sbd);
    H = 0.5f * __fadd_rn(dac, dbd);
    D = 0.5f * __fsub_rn(dac, dbd);
}
__global__ void __launch_bounds__(256)
    k_haar2d_fwd_v4(const float* __restrict__ img, size_t s_img, float* __restrict__ A, size_t s_a, float* __restrict__ H,
                    float* __restrict__ V, float* __restrict__ D, size_t s_d, int nr, int nc)
{
    const int q = blockIdx.x * 256 + threadIdx.x, gy = blockIdx.y;   // q: group of 4 output columns
    const size_t pz = blockIdx.z;
    pdl_wait();
    if (4 * q >= nc) return;
    const float4* r0 = reinterpret_cast<const float4*>(img + pz * s_img + (size_t)(2 * gy) * (2 * nc)) + 2 * q;
    const float4* r1 = r0 + nc / 2;
    const float4 t0 = __ldg(r0), t1 = __ldg(r0 + 1), b0 = __ldg(r1), b1 = __ldg(r1 + 1);
    pdl_launch_dependents();
    float4 oa, ov, oh, od;
    haar_bfly(t0.x, t0.y, b0.x, b0.y, oa.x, ov.x, oh.x, od.x);
    haar_bfly(t0.z, t0.w, b0.z, b0.w, oa.y, ov.y, oh.y, od.y);
    haar_bfly(t1.x, t1.y, b1.x, b1.y, oa.z, ov.z, oh.z, od.z);
    haar_bfly(t1.z, t1.w, b1.z, b1.w, oa.w, ov.w, oh.w, od.w);
    const size_t o = (size_t)gy * nc + 4 * q;
    *reinterpret_cast<float4*>(A + pz * s_a + o) = oa;
    *reinterpret_cast<float4*>(V + pz * s_d + o) = ov;
    *reinterpret_cast<float4*>(H + pz * s_d + o) = oh;
    *reinterpret_cast<float4*>(D + pz * s_d + o) = od;
}
// inverse: (a,b,c,d) = (A,V,H,D); out[2y][2x] = .5((a+c)+(b+d)), [2y][2x+1] = .5((a+c)-(b+d)), [2y+1][2x] = .5((a-c)+(b-d)),
// [2y+1][2x+1] = .5((a-c)-(b-d))  (haar.cu:45-54)
__global__ void __launch_bounds__(256)
    k_haar2d_inv_v4(float* __restrict__ img, size_t s_img, const float* __restrict__ A, size_t s_a,
                    const float* __restrict__ H, const float* __restrict__ V, const float* __restrict__ D, size_t s_d,
                    int nr, int nc)
{
    const int q = blockIdx.x * 256 + threadIdx.x, gy = blockIdx.y;
    const size_t pz = blockIdx.z;
    pdl_wait();
    if (4 * q >= nc) return;
    const size_t i = (size_t)gy * nc + 4 * q;
    const float4 a = __ldg(reinterpret_cast<const float4*>(A + pz * s_a + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(V + pz * s_d + i));
    const float4 c = __ldg(reinterpret_cast<const float4*>(H + pz * s_d + i));
    const float4 d = __ldg(reinterpret_cast<const float4*>(D + pz * s_d + i));
    pdl_launch_dependents();
    float4 t0, t1, b0, b1;
    // the butterfly is its own inverse pattern: (ee, eo, oe, oo) of one 2x2 block come out in the slots (A, V, H, D)
    haar_bfly(a.x, b.x, c.x, d.x, t0.x, t0.y, b0.x, b0.y);
    haar_bfly(a.y, b.y, c.y, d.y, t0.z, t0.w, b0.z, b0.w);
    haar_bfly(a.z, b.z, c.z, d.z, t1.x, t1.y, b1.x, b1.y);
    haar_bfly(a.w, b.w, c.w, d.w, t1.z, t1.w, b1.z, b1.w);
    float4* r0 = reinterpret_cast<float4*>(img + pz * s_img + (size_t)(2 * gy) * (2 * nc)) + 2 * q;
    float4* r1 = r0 + nc / 2;
    r0[0] = t0; r0[1] = t1; r1[0] = b0; r1[1] = b1;
}

// kern_haar1d_fwd, haar.cu:132-146: the factor is a DOUBLE literal (haar.cu:128): float add, double multiply.
__global__ void __launch_bounds__(GX* GY)
    k_haar1d_fwd(const float* __restrict__ img, size_t s_img, float* __restrict__ A, size_t s_a, float* __restrict__ D,
                 size_t s_d, int Nr, int Nc)
{
    IDX2;
    const int nc = half_up(Nc);
    if (gy >= Nr || gx >= nc) return;
    int x1 = 2 * gx + 1;
    if ((Nc & 1) && x1 == Nc) x1--;
    const float* p = img + pz * s_img + (size_t)gy * Nc;
    const float a = p[2 * gx], b = p[x1];
    A[pz * s_a + (size_t)gy * nc + gx] = (float)(0.70710678118654746 * (double)__fadd_rn(a, b));
    D[pz * s_d + (size_t)gy * nc + gx] = (float)(0.70710678118654746 * (double)__fsub_rn(a, b));
}

// kern_haar1d_inv, haar.cu:149-160.  Nc = coefficient width, Nc2 = output width.
__global__ void __launch_bounds__(GX* GY)
    k_haar1d_inv(float* __restrict__ img, size_t s_img, const float* __restrict__ A, size_t s_a,
                 const float* __restrict__ D, size_t s_d, int Nr, int Nc, int Nc2)
{
    IDX2;
    if (gy >= Nr || gx >= Nc2) return;
    const float a = A[pz * s_a + (size_t)gy * Nc + (gx >> 1)], b = D[pz * s_d + (size_t)gy * Nc + (gx >> 1)];
    const float r = (gx & 1) ? __fsub_rn(a, b) : __fadd_rn(a, b);
    img[pz * s_img + (size_t)gy * Nc2 + gx] = (float)(0.70710678118654746 * (double)r);
}

// Vectorised 1-D Haar level for even widths with nc % 4 == 0 and 16-byte aligned planes: 8 samples <-> 4 + 4 coefficients
// per thread, 128-bit accesses, the reference's float add / double multiply / round sequence (haar.cu:128,143-157).
__device__ __forceinline__ float haar_s(float v) { return (float)(0.70710678118654746 * (double)v); }
__global__ void __launch_bounds__(256)
    k_haar1d_fwd_v4(const float* __restrict__ img, size_t s_img, float* __restrict__ A, size_t s_a, float* __restrict__ D,
                    size_t s_d, int nc)
{
    const int q = blockIdx.x * 256 + threadIdx.x;
    const size_t row = blockIdx.y, pz = blockIdx.z;
    pdl_wait();
    if (4 * q >= nc) return;
    const float4* r = reinterpret_cast<const float4*>(img + pz * s_img + row * (2 * (size_t)nc)) + 2 * q;
    const float4 x0 = __ldg(r), x1 = __ldg(r + 1);
    pdl_launch_dependents();
    const size_t o = row * nc + 4 * q;
    *reinterpret_cast<float4*>(A + pz * s_a + o) = make_float4(haar_s(__fadd_rn(x0.x, x0.y)), haar_s(__fadd_rn(x0.z, x0.w)),
                                                              haar_s(__fadd_rn(x1.x, x1.y)), haar_s(__fadd_rn(x1.z, x1.w)));
    *reinterpret_cast<float4*>(D + pz * s_d + o) = make_float4(haar_s(__fsub_rn(x0.x, x0.y)), haar_s(__fsub_rn(x0.z, x0.w)),
                                                              haar_s(__fsub_rn(x1.x, x1.y)), haar_s(__fsub_rn(x1.z, x1.w)));
}
__global__ void __launch_bounds__(256)
    k_haar1d_inv_v4(float* __restrict__ img, size_t s_img, const float* __restrict__ A, size_t s_a,
                    const float* __restrict__ D, size_t s_d, int nc)
{
    const int q = blockIdx.x * 256 + threadIdx.x;
    const size_t row = blockIdx.y, pz = blockIdx.z;
    pdl_wait();
    if (4 * q >= nc) return;
    const size_t i = row * nc + 4 * q;
    const float4 a = __ldg(reinterpret_cast<const float4*>(A + pz * s_a + i));
    const float4 d = __ldg(reinterpret_cast<const float4*>(D + pz * s_d + i));
    pdl_launch_dependents();
    float4* r = reinterpret_cast<float4*>(img + pz * s_img + row * (2 * (size_t)nc)) + 2 * q;
    r[0] = make_float4(haar_s(__fadd_rn(a.x, d.x)), haar_s(__fsub_rn(a.x, d.x)), haar_s(__fadd_rn(a.y, d.y)),
                       haar_s(__fsub_rn(a.y, d.y)));
    r[1] = make_float4(haar_s(__fadd_rn(a.z, d.z)), haar_s(__fsub_rn(a.z, d.z)), haar_s(__fadd_rn(a.w, d.w)),
                       haar_s(__fsub_rn(a.w, d.w)));
}

// ---------------------------------------------------------------------------------------------- non-separable
// The four 2-D filters are outer products of the 1-D banks rounded to float on the host (w_outer,
// nonseparable.cu:16-24,71-74): K_LL[i][j] = L[i]*L[j], K_LH = L[i]*H[j], K_HL = H[i]*L[j], K_HH = H[i]*H[j] with i
// along y.  __fmul_rn reproduces those products bit for bit, so they are formed on the fly instead of being
// uploaded (and re-uploaded before every inverse, wt.cu:298).

// w_kern_forward, nonseparable.cu:114-170
__global__ void __launch_bounds__(GX* GY)
    k_nonsep_fwd(const __grid_constant__ Taps t, const float* __restrict__ img, size_t s_img, float* __restrict__ A,
                 size_t s_a, float* __restrict__ H, float* __restrict__ V, float* __restrict__ D, size_t s_d, int Nr,
                 int Nc)
{
    IDX2;
    const int nr = half_up(Nr), nc = half_up(Nc);
    if (gy >= nr || gx >= nc) return;
    const int hlen = t.hlen, c = centre_fwd(hlen);
    const float* p = img + pz * s_img;
    float ra = 0.f, rh = 0.f, rv = 0.f, rd = 0.f;
    for (int jy = 0; jy < hlen; jy++) {
        const float* row = p + (size_t)fold_dec(2 * gy - c + jy, Nr) * Nc;
        const float ly = t.L[hlen - 1 - jy], hy = t.H[hlen - 1 - jy];
        for (int jx = 0; jx < hlen; jx++) {
            const float v = row[fold_dec(2 * gx - c + jx, Nc)];
            const float lx = t.L[hlen - 1 - jx], hx = t.H[hlen - 1 - jx];
            const int ki = (hlen - 1 - jy) * hlen + (hlen - 1 - jx);   // nonseparable.cu:155-160
            ra = fmaf(v, t.k2d ? k2d_at(t.k2d, hlen, 0, ki) : __fmul_rn(ly, lx), ra);
            rh = fmaf(v, t.k2d ? k2d_at(t.k2d, hlen, 1, ki) : __fmul_rn(ly, hx), rh);
            rv = fmaf(v, t.k2d ? k2d_at(t.k2d, hlen, 2, ki) : __fmul_rn(hy, lx), rv);
            rd = fmaf(v, t.k2d ? k2d_at(t.k2d, hlen, 3, ki) : __fmul_rn(hy, hx), rd);
        }
    }
    const size_t o = (size_t)gy * nc + gx;
    A[pz * s_a + o] = ra;
    H[pz * s_d + o] = rh;
    V[pz * s_d + o] = rv;
    D[pz * s_d + o] = rd;
}

// w_kern_inverse, nonseparable.cu:176-225.  Nr,Nc = coefficient size, Nr2,Nc2 = output size.
__global__ void __launch_bounds__(GX* GY)
    k_nonsep_inv(const __grid_constant__ Taps t, float* __restrict__ img, size_t s_img, const float* __restrict__ A,
                 size_t s_a, const float* __restrict__ H, const float* __restrict__ V, const float* __restrict__ D,
                 size_t s_d, int Nr, int Nc, int Nr2, int Nc2)
{
    IDX2;
    if (gy >= Nr2 || gx >= Nc2) return;
    const int hlen = t.hlen;
    const SynGeom sg = syn_geometry(hlen);
    const int vy = gy + sg.shift, vx = gx + sg.shift;
    const int hy = vy / 2, hx = vx / 2, oy = 1 - (vy & 1), ox = 1 - (vx & 1);
    const float *pa = A + pz * s_a, *ph = H + pz * s_d, *pv = V + pz * s_d, *pd = D + pz * s_d;
    float ra = 0.f, rh = 0.f, rv = 0.f, rd = 0.f;
    for (int jy = 0; jy < sg.taps; jy++) {
        int y = hy - sg.c + jy;
        if (jy < sg.c - hy) y += Nr;
        if (jy > Nr - 1 - hy + sg.c) y -= Nr;
        const float ly = t.IL[hlen - 1 - (2 * jy + oy)], hyv = t.IH[hlen - 1 - (2 * jy + oy)];
        for (int jx = 0; jx < sg.taps; jx++) {
            int x = hx - sg.c + jx;
            if (jx < sg.c - hx) x += Nc;
            if (jx > Nc - 1 - hx + sg.c) x -= Nc;
            const float lx = t.IL[hlen - 1 - (2 * jx + ox)], hxv = t.IH[hlen - 1 - (2 * jx + ox)];
            const size_t i = (size_t)y * Nc + x;
            const int ki = (hlen - 1 - (2 * jy + oy)) * hlen + (hlen - 1 - (2 * jx + ox));   // nonseparable.cu:216-219
            ra = fmaf(pa[i], t.k2d ? k2d_at(t.k2d, hlen, 0, ki) : __fmul_rn(ly, lx), ra);
            rh = fmaf(ph[i], t.k2d ? k2d_at(t.k2d, hlen, 1, ki) : __fmul_rn(ly, hxv), rh);
            rv = fmaf(pv[i], t.k2d ? k2d_at(t.k2d, hlen, 2, ki) : __fmul_rn(hyv, lx), rv);
            rd = fmaf(pd[i], t.k2d ? k2d_at(t.k2d, hlen, 3, ki) : __fmul_rn(hyv, hxv), rd);
        }
    }
    img[pz * s_img + (size_t)gy * Nc2 + gx] = __fadd_rn(__fadd_rn(__fadd_rn(ra, rh), rv), rd);
}

// w_kern_forward_swt, nonseparable.cu:304-354
__global__ void __launch_bounds__(GX* GY)
    k_nonsep_swt_fwd(const __grid_constant__ Taps t, const float* __restrict__ img, size_t s_img,
                     float* __restrict__ A, size_t s_a, float* __restrict__ H, float* __restrict__ V,
                     float* __restrict__ D, size_t s_d, int Nr, int Nc, int fac)
{
    IDX2;
    if (gy >= Nr || gx >= Nc) return;
    const int hlen = t.hlen, c = centre_fwd(hlen) * fac;
    const float* p = img + pz * s_img;
    float ra = 0.f, rh = 0.f, rv = 0.f, rd = 0.f;
    for (int jy = 0; jy < hlen; jy++) {
        const float* row = p + (size_t)fold_swt(gy, jy * fac, c, Nr) * Nc;
        const float ly = t.L[hlen - 1 - jy], hy = t.H[hlen - 1 - jy];
        for (int jx = 0; jx < hlen; jx++) {
            const float v = row[fold_swt(gx, jx * fac, c, Nc)];
            const float lx = t.L[hlen - 1 - jx], hx = t.H[hlen - 1 - jx];
            const int ki = (hlen - 1 - jy) * hlen + (hlen - 1 - jx);   // nonseparable.cu:339-342
            ra = fmaf(v, t.k2d ? k2d_at(t.k2d, hlen, 0, ki) : __fmul_rn(ly, lx), ra);
            rh = fmaf(v, t.k2d ? k2d_at(t.k2d, hlen, 1, ki) : __fmul_rn(ly, hx), rh);
            rv = fmaf(v, t.k2d ? k2d_at(t.k2d, hlen, 2, ki) : __fmul_rn(hy, lx), rv);
            rd = fmaf(v, t.k2d ? k2d_at(t.k2d, hlen, 3, ki) : __fmul_rn(hy, hx), rd);
        }
    }
    const size_t o = (size_t)gy * Nc + gx;
    A[pz * s_a + o] = ra;
    H[pz * s_d + o] = rh;
    V[pz * s_d + o] = rv;
    D[pz * s_d + o] = rd;
}

// w_kern_inverse_swt, nonseparable.cu:360-401: round(v*K), quarter (exact), add.
__global__ void __launch_bounds__(GX* GY)
    k_nonsep_swt_inv(const __grid_constant__ Taps t, float* __restrict__ img, size_t s_img,
                     const float* __restrict__ A, size_t s_a, const float* __restrict__ H, const float* __restrict__ V,
                     const float* __restrict__ D, size_t s_d, int Nr, int Nc, int fac)
{
    IDX2;
    if (gy >= Nr || gx >= Nc) return;
    const int hlen = t.hlen, c = (hlen / 2) * fac, taps = swt_inv_taps(hlen);
    const float *pa = A + pz * s_a, *ph = H + pz * s_d, *pv = V + pz * s_d, *pd = D + pz * s_d;
    float ra = 0.f, rh = 0.f, rv = 0.f, rd = 0.f;
    for (int jy = 0; jy < taps; jy++) {
        const size_t yo = (size_t)fold_swt(gy, jy * fac, c, Nr) * Nc;
        const float ly = t.IL[hlen - 1 - jy], hy = t.IH[hlen - 1 - jy];
        for (int jx = 0; jx < taps; jx++) {
            const size_t i = yo + fold_swt(gx, jx * fac, c, Nc);
            const float lx = t.IL[hlen - 1 - jx], hx = t.IH[hlen - 1 - jx];
            const int ki = (hlen - 1 - jy) * hlen + (hlen - 1 - jx);   // nonseparable.cu:390-393
            ra = __fadd_rn(ra, __fmul_rn(pa[i], t.k2d ? k2d_at(t.k2d, hlen, 0, ki) : __fmul_rn(ly, lx)) * 0.25f);
            rh = __fadd_rn(rh, __fmul_rn(ph[i], t.k2d ? k2d_at(t.k2d, hlen, 1, ki) : __fmul_rn(ly, hx)) * 0.25f);
            rv = __fadd_rn(rv, __fmul_rn(pv[i], t.k2d ? k2d_at(t.k2d, hlen, 2, ki) : __fmul_rn(hy, lx)) * 0.25f);
            rd = __fadd_rn(rd, __fmul_rn(pd[i], t.k2d ? k2d_at(t.k2d, hlen, 3, ki) : __fmul_rn(hy, hx)) * 0.25f);
        }
    }
    img[pz * s_img + (size_t)gy * Nc + gx] = __fadd_rn(__fadd_rn(__fadd_rn(ra, rh), rv), rd);
}

// ---------------------------------------------------------------------------------------------------- launchers
int g_fwd_rows(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int batch, cudaStream_t s)
{
    if (const int d = r_fwd_rows(t, img, lo, hi, Nr, Nc, batch, s)) return d < 0 ? d : 0;   // staged row kernels
    PDWT_PROF(__func__, s);
    k_fwd_rows<<<grid2(half_up(Nc), Nr, batch), kBlock, 0, s>>>(t, img.p, img.stride, lo.p, lo.stride, hi.p, hi.stride,
                                                                Nr, Nc);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_fwd_cols(const Taps& t, Plane2 t1, Plane2 t2, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
               cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    k_fwd_cols<<<grid2(Nc, half_up(Nr), batch), kBlock, 0, s>>>(t, t1.p, t2.p, t1.stride, A.p, H.p, V.p, D.p, A.stride,
                                                                H.stride, Nr, Nc);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_inv_cols(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 t1, Plane2 t2, int n, int Nc, int M,
               int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    k_inv_cols<<<grid2(Nc, M, batch), kBlock, 0, s>>>(t, A.p, H.p, V.p, D.p, A.stride, H.stride, t1.p, t2.p, t1.stride,
                                                      n, Nc, M);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_inv_rows(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int n, int M, int batch, cudaStream_t s)
{
    if (const int d = r_inv_rows(t, t1, t2, img, Nr, n, M, batch, s)) return d < 0 ? d : 0;
    PDWT_PROF(__func__, s);
    k_inv_rows<<<grid2(M, Nr, batch), kBlock, 0, s>>>(t, t1.p, t1.stride, t2.p, t2.stride, img.p, img.stride, Nr, n, M);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_swt_fwd_rows(const Taps& t, Plane2 img, Plane2 lo, Plane2 hi, int Nr, int Nc, int level, int batch,
                   cudaStream_t s)
{
    if (const int d = r_swt_fwd_rows(t, img, lo, hi, Nr, Nc, level, batch, s)) return d < 0 ? d : 0;
    PDWT_PROF(__func__, s);
    k_swt_fwd_rows<<<grid2(Nc, Nr, batch), kBlock, 0, s>>>(t, img.p, img.stride, lo.p, lo.stride, hi.p, hi.stride, Nr,
                                                           Nc, 1 << (level - 1));
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_swt_fwd_cols(const Taps& t, Plane2 t1, Plane2 t2, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc,
                   int level, int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    k_swt_fwd_cols<<<grid2(Nc, Nr, batch), kBlock, 0, s>>>(t, t1.p, t2.p, t1.stride, A.p, H.p, V.p, D.p, A.stride,
                                                           H.stride, Nr, Nc, 1 << (level - 1));
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_swt_inv_cols(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 t1, Plane2 t2, int Nr, int Nc,
                   int level, int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    k_swt_inv_cols<<<grid2(Nc, Nr, batch), kBlock, 0, s>>>(t, A.p, H.p, V.p, D.p, A.stride, H.stride, t1.p, t2.p,
                                                           t1.stride, Nr, Nc, 1 << (level - 1));
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_swt_inv_rows(const Taps& t, Plane2 t1, Plane2 t2, Plane2 img, int Nr, int Nc, int level, int batch,
                   cudaStream_t s)
{
    if (const int d = r_swt_inv_rows(t, t1, t2, img, Nr, Nc, level, batch, s)) return d < 0 ? d : 0;
    PDWT_PROF(__func__, s);
    k_swt_inv_rows<<<grid2(Nc, Nr, batch), kBlock, 0, s>>>(t, t1.p, t1.stride, t2.p, t2.stride, img.p, img.stride, Nr,
                                                           Nc, 1 << (level - 1));
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_haar2d_fwd(Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    const int nr = Nr / 2, nc = Nc / 2;
    if (!(Nr & 1) && !(Nc & 1) && !(nc & 3) && nr <= 65535 && !(img.stride & 3) && !(A.stride & 3) && !(H.stride & 3) &&
        !(((uintptr_t)img.p | (uintptr_t)A.p | (uintptr_t)H.p | (uintptr_t)V.p | (uintptr_t)D.p) & 15)) {
        PDWT_CUDA(launch_pdl(k_haar2d_fwd_v4, dim3(idiv_up(nc / 4, 256), nr, batch), 256, 0, s, (const float*)img.p,
                             img.stride, A.p, A.stride, H.p, V.p, D.p, H.stride, nr, nc));
        PDWT_LAUNCH_CHECK();
        return 0;
    }
    k_haar2d_fwd<<<grid2(half_up(Nc), half_up(Nr), batch), kBlock, 0, s>>>(img.p, img.stride, A.p, A.stride, H.p, V.p,
                                                                          D.p, H.stride, Nr, Nc);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_haar2d_inv(Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int Nr2, int Nc2, int batch,
                 cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    // Nr x Nc = coefficient plane, Nr2 x Nc2 = output plane
    if (Nr2 == 2 * Nr && Nc2 == 2 * Nc && !(Nc & 3) && Nr <= 65535 && !(img.stride & 3) && !(A.stride & 3) &&
        !(H.stride & 3) && !(((uintptr_t)img.p | (uintptr_t)A.p | (uintptr_t)H.p | (uintptr_t)V.p | (uintptr_t)D.p) & 15)) {
        PDWT_CUDA(launch_pdl(k_haar2d_inv_v4, dim3(idiv_up(Nc / 4, 256), Nr, batch), 256, 0, s, img.p, img.stride,
                             (const float*)A.p, A.stride, (const float*)H.p, (const float*)V.p, (const float*)D.p,
                             H.stride, Nr, Nc));
        PDWT_LAUNCH_CHECK();
        return 0;
    }
    k_haar2d_inv<<<grid2(Nc2, Nr2, batch), kBlock, 0, s>>>(img.p, img.stride, A.p, A.stride, H.p, V.p, D.p, H.stride,
                                                          Nc, Nr2, Nc2);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_haar1d_fwd(Plane2 img, Plane2 A, Plane2 D, int Nr, int Nc, int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    if (!(Nc & 7) && Nr <= 65535 && !(img.stride & 3) && !(A.stride & 3) && !(D.stride & 3) &&
        !(((uintptr_t)img.p | (uintptr_t)A.p | (uintptr_t)D.p) & 15)) {
        PDWT_CUDA(launch_pdl(k_haar1d_fwd_v4, dim3(idiv_up(Nc / 8, 256), Nr, batch), 256, 0, s, (const float*)img.p,
                             img.stride, A.p, A.stride, D.p, D.stride, Nc / 2));
        PDWT_LAUNCH_CHECK();
        return 0;
    }
    k_haar1d_fwd<<<grid2(half_up(Nc), Nr, batch), kBlock, 0, s>>>(img.p, img.stride, A.p, A.stride, D.p, D.stride, Nr,
                                                                 Nc);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_haar1d_inv(Plane2 img, Plane2 A, Plane2 D, int Nr, int Nc, int Nc2, int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    if (Nc2 == 2 * Nc && !(Nc & 3) && Nr <= 65535 && !(img.stride & 3) && !(A.stride & 3) && !(D.stride & 3) &&
        !(((uintptr_t)img.p | (uintptr_t)A.p | (uintptr_t)D.p) & 15)) {
        PDWT_CUDA(launch_pdl(k_haar1d_inv_v4, dim3(idiv_up(Nc / 4, 256), Nr, batch), 256, 0, s, img.p, img.stride,
                             (const float*)A.p, A.stride, (const float*)D.p, D.stride, Nc));
        PDWT_LAUNCH_CHECK();
        return 0;
    }
    k_haar1d_inv<<<grid2(Nc2, Nr, batch), kBlock, 0, s>>>(img.p, img.stride, A.p, A.stride, D.p, D.stride, Nr, Nc, Nc2);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_nonsep_fwd(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                 cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    k_nonsep_fwd<<<grid2(half_up(Nc), half_up(Nr), batch), kBlock, 0, s>>>(t, img.p, img.stride, A.p, A.stride, H.p,
                                                                          V.p, D.p, H.stride, Nr, Nc);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_nonsep_inv(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int Nr2, int Nc2,
                 int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    k_nonsep_inv<<<grid2(Nc2, Nr2, batch), kBlock, 0, s>>>(t, img.p, img.stride, A.p, A.stride, H.p, V.p, D.p, H.stride,
                                                          Nr, Nc, Nr2, Nc2);
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_nonsep_swt_fwd(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                     int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    k_nonsep_swt_fwd<<<grid2(Nc, Nr, batch), kBlock, 0, s>>>(t, img.p, img.stride, A.p, A.stride, H.p, V.p, D.p,
                                                            H.stride, Nr, Nc, 1 << (level - 1));
    PDWT_LAUNCH_CHECK();
    return 0;
}
int g_nonsep_swt_inv(const Taps& t, Plane2 img, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int level,
                     int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    k_nonsep_swt_inv<<<grid2(Nc, Nr, batch), kBlock, 0, s>>>(t, img.p, img.stride, A.p, A.stride, H.p, V.p, D.p,
                                                            H.stride, Nr, Nc, 1 << (level - 1));
    PDWT_LAUNCH_CHECK();
    return 0;
}

}  // namespace pdwt
