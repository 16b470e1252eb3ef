// pdwt_elementwise.cu -- soft / hard threshold and L1 / L2 reductions over a table of sub-bands.
//
// The reference launches one 16x16-block kernel per level (common.cu:219-282) and, for the norms, one blocking
// cuBLAS call per sub-band (wt.cu:370-418).  Here ONE launch covers every sub-band of every plane: grid.y walks
// the segment table, grid.z the planes, and each thread streams 128-bit vectors (scalar head/tail for unaligned
// planes).  Pure HBM streaming: 8 B/coefficient for a threshold, 4 B/coefficient for a norm.
#include "pdwt_common.cuh"

namespace pdwt {

// w_kern_soft_thresh*, common.cu:13-52
__device__ __forceinline__ float soft1(float v, float beta) { return copysignf(fmaxf(fabsf(v) - beta, 0.0f), v); }
// w_kern_hard_thresh*, common.cu:57-97 with W_SIGN (common.cu:7): max(sign(|v|-beta), 0) * v  (keeps 0*v = +-0)
__device__ __forceinline__ float hard1(float v, float beta)
{
    const float s = (fabsf(v) - beta > 0.0f) ? 1.0f : -1.0f;
    return fmaxf(s, 0.0f) * v;
}

// w_kern_proj_linf*, common.cu:101-138: projection onto the L-infinity ball
__device__ __forceinline__ float linf1(float v, float beta) { return copysignf(fminf(fabsf(v), beta), v); }
// OP: 0 soft, 1 hard, 2 proj_linf, 3 scale by beta (w_shrink's cublas scal, common.cu:343-369)
template <int OP>
__device__ __forceinline__ float ew1(float v, float beta)
{
    return OP == 0 ? soft1(v, beta) : OP == 1 ? hard1(v, beta) : OP == 2 ? linf1(v, beta) : __fmul_rn(beta, v);
}

template <int OP, bool SUMS>
__global__ void __launch_bounds__(256) k_threshold(const __grid_constant__ SegTable tab, double* __restrict__ sums, int batch)
{
    const int seg = blockIdx.y;
    float* p = tab.ptr[seg] + (size_t)blockIdx.z * tab.stride[seg];
    const size_t n = tab.n[seg];
    const float beta = tab.beta[seg];
    // scalar head up to the first 16-byte boundary, vector body, scalar tail
    size_t head = ((16 - ((uintptr_t)p & 15)) & 15) >> 2;
    if (head > n) head = n;
    const size_t nvec = (n - head) >> 2;
    // the grid is sized for the largest segment: a smaller one uses only as many blocks as give a thread >= 4 vectors
    const size_t nb = min((size_t)gridDim.x, (nvec + 1023) / 1024 + 1);
    if (blockIdx.x >= nb) return;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = nb * blockDim.x;
    float4* pv = reinterpret_cast<float4*>(p + head);
    const bool ro = tab.ro[seg] != 0;   // read-only segment: only its sums are wanted
    double s1 = 0.0, s2 = 0.0;          // sum |out|, sum out^2 (SUMS); per vector in float like k_reduce, then double
    auto apply = [&](float4 v) {
        if (!ro) {
            v.x = ew1<OP>(v.x, beta); v.y = ew1<OP>(v.y, beta); v.z = ew1<OP>(v.z, beta); v.w = ew1<OP>(v.w, beta);
        }
        if (SUMS) {
            s1 += (double)((fabsf(v.x) + fabsf(v.y)) + (fabsf(v.z) + fabsf(v.w)));
            s2 += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
        }
        return v;
    };
    auto apply1 = [&](float v) {
        if (!ro) v = ew1<OP>(v, beta);
        if (SUMS) {
            s1 += (double)fabsf(v);
            s2 += (double)(v * v);
        }
        return v;
    };
    size_t i = tid;
    for (; i + 3 * nth < nvec; i += 4 * nth) {   // four independent 128-bit loads in flight per thread
        const float4 a = apply(pv[i]), b = apply(pv[i + nth]), c = apply(pv[i + 2 * nth]), d = apply(pv[i + 3 * nth]);
        if (!ro) {
            pv[i] = a;
            pv[i + nth] = b;
            pv[i + 2 * nth] = c;
            pv[i + 3 * nth] = d;
        }
    }
    for (; i < nvec; i += nth) {
        const float4 a = apply(pv[i]);
        if (!ro) pv[i] = a;
    }
    if (tid < head) {
        const float a = apply1(p[tid]);
        if (!ro) p[tid] = a;
    }
    const size_t tail0 = head + (nvec << 2);
    if (tail0 + tid < n) {
        const float a = apply1(p[tail0 + tid]);
        if (!ro) p[tail0 + tid] = a;
    }
    if (SUMS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        __shared__ double w1[8], w2[8];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) {
            w1[wid] = s1;
            w2[wid] = s2;
        }
        __syncthreads();
        if (wid == 0) {
            s1 = lane < 8 ? w1[lane] : 0.0;
            s2 = lane < 8 ? w2[lane] : 0.0;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (lane == 0) {
                const size_t k = (size_t)blockIdx.z * tab.nseg + seg;
                atomicAdd(&sums[k], s1);
                atomicAdd(&sums[(size_t)batch * tab.nseg + k], s2);
            }
        }
    }
}

// sum |v| (MODE 0) or sum v^2 (MODE 1): per-thread double accumulation of float4 loads, warp shuffle tree,
// one shared-memory pass across the 8 warps, one double atomicAdd per block.
// (256, 4): without the bound ptxas took 96 registers, two blocks per SM, and the loads' latency showed (ncu: 8 cycles of
// long_scoreboard per issue, 2.7 TB/s)
template <int MODE>
__global__ void __launch_bounds__(256, 4) k_reduce(const __grid_constant__ SegTable tab, double* __restrict__ sums)
{
    const int seg = blockIdx.y;
    const float* p = tab.ptr[seg] + (size_t)blockIdx.z * tab.stride[seg];
    const size_t n = tab.n[seg];
    size_t head = ((16 - ((uintptr_t)p & 15)) & 15) >> 2;
    if (head > n) head = n;
    const size_t nvec = (n - head) >> 2;
    const size_t nb = min((size_t)gridDim.x, (nvec + 1023) / 1024 + 1);   // blocks that take part in this segment
    if (blockIdx.x >= nb) return;                                        // (no atomic from the others)
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = nb * blockDim.x;
    const float4* pv = reinterpret_cast<const float4*>(p + head);
    double acc = 0.0;
    auto term = [](float v) -> float { return MODE ? v * v : fabsf(v); };
    // the four terms of one vector are added in float (exact enough: 4 terms), the running sum in double
    auto vsum = [&](const float4 v) -> double { return (double)((term(v.x) + term(v.y)) + (term(v.z) + term(v.w))); };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = tid; i < nvec; i += 4 * nth) {   // four independent (guarded) 128-bit loads in flight per thread
        const float4 a = __ldg(pv + i);
        const float4 b = i + nth < nvec ? __ldg(pv + i + nth) : z4;
        const float4 c = i + 2 * nth < nvec ? __ldg(pv + i + 2 * nth) : z4;
        const float4 d = i + 3 * nth < nvec ? __ldg(pv + i + 3 * nth) : z4;
        acc += (vsum(a) + vsum(b)) + (vsum(c) + vsum(d));
    }
    if (tid < head) acc += (double)term(p[tid]);
    const size_t tail0 = head + (nvec << 2);
    if (tail0 + tid < n) acc += (double)term(p[tail0 + tid]);

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double warp_sum[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_sum[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        acc = lane < 8 ? warp_sum[lane] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) atomicAdd(&sums[(size_t)blockIdx.z * tab.nseg + seg], acc);
    }
}

static int blocks_for(const SegTable& tab, int per_sm = 4)
{
    unsigned long long nmax = 0;
    for (int i = 0; i < tab.nseg; i++) nmax = tab.n[i] > nmax ? tab.n[i] : nmax;
    // 256 threads x 4 floats x 4 vectors in flight per thread per pass; at most 4 blocks per SM per segment (the segment
    // table puts nseg of these grids side by side)
    unsigned long long b = (nmax + 4095) / 4096;
    if (b < 1) b = 1;
    if (b > 148ull * per_sm) b = 148ull * per_sm;
    return (int)b;
}

int e_threshold(const SegTable& tab, int op, int batch, cudaStream_t s, double* sums)
{
    PDWT_PROF(__func__, s);
    if (tab.nseg == 0) return 0;
    dim3 grid(blocks_for(tab, sums ? 2 : 4), tab.nseg, batch);
    if (sums) {   // thresholds that also deliver the norms of their result
        switch (op) {
            case 0: k_threshold<0, true><<<grid, 256, 0, s>>>(tab, sums, batch); break;
            case 1: k_threshold<1, true><<<grid, 256, 0, s>>>(tab, sums, batch); break;
            case 2: k_threshold<2, true><<<grid, 256, 0, s>>>(tab, sums, batch); break;
            default: k_threshold<3, true><<<grid, 256, 0, s>>>(tab, sums, batch); break;
        }
    } else {
        switch (op) {
            case 0: k_threshold<0, false><<<grid, 256, 0, s>>>(tab, nullptr, batch); break;
            case 1: k_threshold<1, false><<<grid, 256, 0, s>>>(tab, nullptr, batch); break;
            case 2: k_threshold<2, false><<<grid, 256, 0, s>>>(tab, nullptr, batch); break;
            default: k_threshold<3, false><<<grid, 256, 0, s>>>(tab, nullptr, batch); break;
        }
    }
    PDWT_LAUNCH_CHECK();
    return 0;
}

// ---- group soft threshold, w_kern_group_soft_thresh(_1d), common.cu:141-196: one scaling factor per position from the
// Euclidean norm of (H, V, D[, A]) -- or (D[, A]) in 1-D -- at that position.  The expression is written exactly as in the
// reference so that nvcc contracts it the same way (the golden vectors pin the result).
__global__ void __launch_bounds__(256) k_group_soft(const __grid_constant__ GroupTable tab)
{
    const int lev = blockIdx.y;
    const size_t n = tab.n[lev], pz = blockIdx.z;
    const float beta = tab.beta[lev];
    float* c_h = tab.h[lev] ? tab.h[lev] + pz * tab.stride_d[lev] : nullptr;
    float* c_v = tab.v[lev] ? tab.v[lev] + pz * tab.stride_d[lev] : nullptr;
    float* c_d = tab.d[lev] + pz * tab.stride_d[lev];
    float* c_a = tab.a[lev] ? tab.a[lev] + pz * tab.stride_a : nullptr;
    for (size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += (size_t)gridDim.x * blockDim.x) {
        float val_h = 0.0f, val_v = 0.0f, val_d = 0.0f, val_a = 0.0f;
        float norm = 0, res = 0;
        val_d = c_d[tid];
        if (c_h) {   // 2-D
            val_h = c_h[tid];
            val_v = c_v[tid];
            norm = val_h * val_h + val_v * val_v + val_d * val_d;
        } else {
            norm = val_d * val_d;
        }
        if (c_a != nullptr) {
            val_a = c_a[tid];
            norm += val_a * val_a;
        }
        norm = sqrtf(norm);
        if (norm == 0)
            res = 0;
        else
            res = max(1 - beta / norm, 0.0);
        if (c_h) {
            c_h[tid] *= res;
            c_v[tid] *= res;
        }
        c_d[tid] *= res;
        if (c_a != nullptr) c_a[tid] *= res;
    }
}

int e_group_soft(const GroupTable& tab, int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    if (tab.nlev == 0) return 0;
    unsigned long long nmax = 0;
    for (int i = 0; i < tab.nlev; i++) nmax = tab.n[i] > nmax ? tab.n[i] : nmax;
    unsigned long long b = (nmax + 1023) / 1024;
    b = b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b);
    k_group_soft<<<dim3((unsigned)b, tab.nlev, batch), 256, 0, s>>>(tab);
    PDWT_LAUNCH_CHECK();
    return 0;
}

// ---- dst += alpha * src over a table of sub-band pairs (w_add_coeffs, common.cu:499-526: cublas axpy)
__global__ void __launch_bounds__(256) k_axpy(const __grid_constant__ PairTable tab, float alpha)
{
    const int seg = blockIdx.y;
    const size_t n = tab.n[seg], pz = blockIdx.z;
    float* dst = tab.dst[seg] + pz * tab.stride_dst[seg];
    const float* src = tab.src[seg] + pz * tab.stride_src[seg];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = fmaf(alpha, src[i], dst[i]);
}
int e_axpy(const PairTable& tab, float alpha, int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    if (tab.nseg == 0) return 0;
    unsigned long long nmax = 0;
    for (int i = 0; i < tab.nseg; i++) nmax = tab.n[i] > nmax ? tab.n[i] : nmax;
    unsigned long long b = (nmax + 1023) / 1024;
    b = b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b);
    k_axpy<<<dim3((unsigned)b, tab.nseg, batch), 256, 0, s>>>(tab, alpha);
    PDWT_LAUNCH_CHECK();
    return 0;
}

// ---- circular shift, w_kern_circshift (common.cu:200-211): out[y][x] = in[(y - sr) mod Nr][(x - sc) mod Nc]
__global__ void __launch_bounds__(256) k_circshift(const float* __restrict__ in, float* __restrict__ out, size_t stride,
                                                   int Nr, int Nc, int sr, int sc)
{
    const int gx = blockIdx.x * 256 + threadIdx.x, gy = blockIdx.y;
    if (gx >= Nc) return;
    int r = gy - sr, c = gx - sc;
    if (r < 0) r += Nr;
    if (c < 0) c += Nc;
    out[(size_t)blockIdx.z * stride + (size_t)gy * Nc + gx] = in[(size_t)blockIdx.z * stride + (size_t)r * Nc + c];
}
int e_circshift(const float* in, float* out, size_t stride, int Nr, int Nc, int sr, int sc, int batch, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    if (Nr > 65535 || batch > 65535) return PDWT_ERR_ARG;
    k_circshift<<<dim3(idiv_up(Nc, 256), Nr, batch), 256, 0, s>>>(in, out, stride, Nr, Nc, sr, sc);
    PDWT_LAUNCH_CHECK();
    return 0;
}

int e_reduce(const SegTable& tab, int mode, int batch, double* d_sums, cudaStream_t s)
{
    PDWT_PROF(__func__, s);
    if (tab.nseg == 0) return 0;
    dim3 grid(blocks_for(tab, 4), tab.nseg, batch);   // one double atomic per block: keep them few
    if (mode)
        k_reduce<1><<<grid, 256, 0, s>>>(tab, d_sums);
    else
        k_reduce<0><<<grid, 256, 0, s>>>(tab, d_sums);
    PDWT_LAUNCH_CHECK();
    return 0;
}

}  // namespace pdwt
