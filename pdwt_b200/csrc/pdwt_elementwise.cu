// pdwt_elementwise.cu -- soft / hard threshold and L1 / L2 reductions over a table of sub-bands.
//
// The reference launches one 16x16-block kernel per level (common.cu:219-282) and, for the norms, one blocking
// cuBLAS call per sub-band (wt.cu:370-418).  Here ONE launch covers every sub-band of every plane: a persistent grid
// walks the segment table as one flat list of 32 KB tiles (see walk_tiles), each thread streaming 8 independent
// 128-bit vectors per tile (scalar head/tail for unaligned planes).  Pure HBM streaming: 8 B/coefficient for a
// threshold, 4 B/coefficient for a norm.
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

// ---- flat persistent walk over (plane, segment) -----------------------------------------------------------------
// The sub-bands of a transform differ in size by 4^level, so a grid with one axis per segment ends up with many short
// blocks (two dependent rounds of loads each) and the HBM pipe never fills (ncu: 3.2 TB/s).  Instead the whole table is
// ONE list of tiles -- kTileV 128-bit vectors, 8 per thread -- laid out plane by plane, segment by segment; block b of
// the (at most 4 x SMs) resident blocks walks the CONTIGUOUS tile range [b*T/nb, (b+1)*T/nb) with 8 independent
// 128-bit loads in flight per thread and flushes its partial sums (block reduction, one double atomic) only where
// its range crosses into another sub-band: a handful of atomics per block.
constexpr int kEwThreads = 256;
constexpr int kVecPerThread = 8;
constexpr int kTileV = kEwThreads * kVecPerThread;   // 128-bit vectors per tile (32 KB)

__host__ __device__ inline unsigned long long seg_tiles(unsigned long long n)
{
    const unsigned long long nv = n >> 2;   // upper bound of the aligned vectors in the body
    return nv <= (unsigned long long)kTileV ? 1ull : (nv + kTileV - 1) / kTileV;   // >= 1: tile 0 owns head and tail
}

struct SegCursor {   // where a block stands in the tile list
    const float* p;  // first float of the (plane, segment)
    size_t n, head, nvec;
    unsigned long long lt, lt_end;   // local tile range of this block inside the segment
    int seg, plane;
};

// block-wide sum of one double per thread; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* scratch /*[8]*/)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < kEwThreads / 32 ? scratch[lane] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    __syncthreads();   // scratch may be reused right away
    return v;
}

// see HostPublish (pdwt_common.cuh).  Every thread of every block calls it once, after its last atomic on `sums`.
__device__ __forceinline__ void publish_sums(const HostPublish& hp, double* sums)
{
    if (!hp.h_out) return;
    __shared__ unsigned is_last;
    if (threadIdx.x == 0) {
        __threadfence();   // this block's atomics are performed before its ticket
        is_last = atomicAdd(hp.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const unsigned long long tag = (unsigned long long)hp.tag << 32;
    for (int i = threadIdx.x; i < hp.nsums; i += blockDim.x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(__ldcg(sums + i));   // the atomics live in L2
        sums[i] = 0.0;
        asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(hp.h_out + 2 * i), "l"(tag | (b & 0xffffffffull)),
                     "l"(tag | (b >> 32))
                     : "memory");
    }
    if (threadIdx.x == 0) *hp.ticket = 0;   // every block has drawn its ticket: nobody touches it before the next launch
}

// Drives `tile(cursor, first vector index of the tile, full)`, `edge(cursor, index)` for the scalar head / tail elements
// and `flush(cursor)` at the end of every (plane, segment) piece of the block's range.  All control flow is block-uniform.
template <class TileFn, class EdgeFn, class FlushFn>
__device__ __forceinline__ void walk_tiles(const SegTable& tab, int batch, unsigned long long* pre /*shared [kMaxSeg+1]*/,
                                           TileFn tile, EdgeFn edge, FlushFn flush)
{
    if (threadIdx.x == 0) {
        unsigned long long a = 0;
        for (int s = 0; s < tab.nseg; s++) {
            pre[s] = a;
            a += seg_tiles(tab.n[s]);
        }
        pre[tab.nseg] = a;
    }
    __syncthreads();
    const unsigned long long TP = pre[tab.nseg], total = TP * (unsigned long long)batch;
    unsigned long long t = total * blockIdx.x / gridDim.x;
    const unsigned long long t1 = total * (blockIdx.x + 1ull) / gridDim.x;
    if (t >= t1) return;
    SegCursor c;
    c.plane = (int)(t / TP);
    const unsigned long long r = t - (unsigned long long)c.plane * TP;
    c.seg = 0;
    while (pre[c.seg + 1] <= r) c.seg++;
    c.lt = r - pre[c.seg];
    for (;;) {
        c.p = tab.ptr[c.seg] + (size_t)c.plane * tab.stride[c.seg];
        c.n = tab.n[c.seg];
        c.head = ((16 - ((uintptr_t)c.p & 15)) & 15) >> 2;   // scalar head up to the first 16-byte boundary
        if (c.head > c.n) c.head = c.n;
        c.nvec = (c.n - c.head) >> 2;
        const unsigned long long ntile = pre[c.seg + 1] - pre[c.seg];
        c.lt_end = min(ntile, c.lt + (t1 - t));
        t += c.lt_end - c.lt;
        if (c.lt == 0) {   // tile 0 also owns the scalar head and tail (at most 3 + 3 elements)
            const size_t tail0 = c.head + (c.nvec << 2);
            if (threadIdx.x < c.head) edge(c, (size_t)threadIdx.x);
            else if (threadIdx.x >= 32 && tail0 + (threadIdx.x - 32) < c.n) edge(c, tail0 + (threadIdx.x - 32));
        }
        for (unsigned long long lt = c.lt; lt < c.lt_end; lt++) {
            const size_t v0 = (size_t)lt * kTileV;
            if (v0 >= c.nvec) break;
            tile(c, v0, v0 + kTileV <= c.nvec);
        }
        flush(c);
        if (t >= t1) return;
        c.lt = 0;
        if (++c.seg == tab.nseg) {
            c.seg = 0;
            c.plane++;
        }
    }
}

template <int OP, bool SUMS>
__global__ void __launch_bounds__(kEwThreads, SUMS ? 3 : 4)
    k_threshold(const __grid_constant__ SegTable tab, double* sums, int batch, const HostPublish hp)
{
    __shared__ unsigned long long pre[kMaxSeg + 1];
    __shared__ double scratch[kEwThreads / 32];
    double s1 = 0.0, s2 = 0.0;   // sum |out|, sum out^2 (SUMS); per vector in float like k_reduce, then double
    auto apply = [&](float4 v, const float beta, const bool ro) {
        if (!ro) {
            v.x = ew1<OP>(v.x, beta); v.y = ew1<OP>(v.y, beta); v.z = ew1<OP>(v.z, beta); v.w = ew1<OP>(v.w, beta);
        }
        if (SUMS) {
            s1 += (double)((fabsf(v.x) + fabsf(v.y)) + (fabsf(v.z) + fabsf(v.w)));
            s2 += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
        }
        return v;
    };
    walk_tiles(
        tab, batch, pre,
        [&](const SegCursor& c, const size_t v0, const bool full) {
            const float beta = tab.beta[c.seg];
            const bool ro = tab.ro[c.seg] != 0;   // read-only segment: only its sums are wanted
            float4* pv = reinterpret_cast<float4*>(const_cast<float*>(c.p) + c.head) + v0 + threadIdx.x;
            float4 v[kVecPerThread];
            if (full) {
#pragma unroll
                for (int k = 0; k < kVecPerThread; k++) v[k] = pv[k * kEwThreads];   // 8 independent 128-bit loads
#pragma unroll
                for (int k = 0; k < kVecPerThread; k++) v[k] = apply(v[k], beta, ro);
                if (!ro) {
#pragma unroll
                    for (int k = 0; k < kVecPerThread; k++) pv[k * kEwThreads] = v[k];
                }
            } else {
                const size_t left = c.nvec - v0;   // vectors in this (last) tile
#pragma unroll
                for (int k = 0; k < kVecPerThread; k++)
                    if (threadIdx.x + (size_t)k * kEwThreads < left) v[k] = pv[k * kEwThreads];
#pragma unroll
                for (int k = 0; k < kVecPerThread; k++)
                    if (threadIdx.x + (size_t)k * kEwThreads < left) {
                        v[k] = apply(v[k], beta, ro);
                        if (!ro) pv[k * kEwThreads] = v[k];
                    }
            }
        },
        [&](const SegCursor& c, const size_t i) {
            float* q = const_cast<float*>(c.p) + i;
            float v = *q;
            if (!tab.ro[c.seg]) {
                v = ew1<OP>(v, tab.beta[c.seg]);
                *q = v;
            }
            if (SUMS) {
                s1 += (double)fabsf(v);
                s2 += (double)(v * v);
            }
        },
        [&](const SegCursor& c) {
            if (SUMS) {
                const double a = block_sum(s1, scratch), b = block_sum(s2, scratch);
                if (threadIdx.x == 0) {
                    const size_t k = (size_t)c.plane * tab.nseg + c.seg;
                    atomicAdd(&sums[k], a);
                    atomicAdd(&sums[(size_t)batch * tab.nseg + k], b);
                }
                s1 = s2 = 0.0;
            }
        });
    if (SUMS) publish_sums(hp, sums);
}

// sum |v| (MODE 0) or sum v^2 (MODE 1): the four terms of one vector are added in float (exact enough: 4 terms), the
// running sum in double; warp shuffle tree, one shared-memory pass across the 8 warps, one double atomicAdd per
// (block, sub-band) piece.
template <int MODE>
__global__ void __launch_bounds__(kEwThreads, 4)
    k_reduce(const __grid_constant__ SegTable tab, double* sums, int batch, const HostPublish hp)
{
    __shared__ unsigned long long pre[kMaxSeg + 1];
    __shared__ double scratch[kEwThreads / 32];
    double acc = 0.0;
    auto term = [](float v) -> float { return MODE ? v * v : fabsf(v); };
    auto vsum = [&](const float4 v) -> double { return (double)((term(v.x) + term(v.y)) + (term(v.z) + term(v.w))); };
    walk_tiles(
        tab, batch, pre,
        [&](const SegCursor& c, const size_t v0, const bool full) {
            const float4* pv = reinterpret_cast<const float4*>(c.p + c.head) + v0 + threadIdx.x;
            float4 v[kVecPerThread];
            const size_t left = c.nvec - v0;
#pragma unroll
            for (int k = 0; k < kVecPerThread; k++)   // 8 independent 128-bit loads in flight per thread
                v[k] = (full || threadIdx.x + (size_t)k * kEwThreads < left) ? __ldg(pv + k * kEwThreads)
                                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
            acc += ((vsum(v[0]) + vsum(v[1])) + (vsum(v[2]) + vsum(v[3]))) +
                   ((vsum(v[4]) + vsum(v[5])) + (vsum(v[6]) + vsum(v[7])));
        },
        [&](const SegCursor& c, const size_t i) { acc += (double)term(c.p[i]); },
        [&](const SegCursor& c) {
            const double a = block_sum(acc, scratch);
            if (threadIdx.x == 0) atomicAdd(&sums[(size_t)c.plane * tab.nseg + c.seg], a);
            acc = 0.0;
        });
    publish_sums(hp, sums);
}

// resident blocks that walk the tile list: up to per_sm per SM, never more than there are tiles
static int blocks_for(const SegTable& tab, int batch, int per_sm)
{
    unsigned long long tiles = 0;
    for (int i = 0; i < tab.nseg; i++) tiles += seg_tiles(tab.n[i]);
    tiles *= (unsigned long long)batch;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) {
        cudaGetLastError();
        sms = 148;
    }
    unsigned long long b = (unsigned long long)sms * per_sm;
    if (b > tiles) b = tiles;
    if (b < 1) b = 1;
    return (int)b;
}

int e_threshold(const SegTable& tab, int op, int batch, cudaStream_t s, double* sums, const HostPublish* hpp)
{
    const HostPublish hp = hpp ? *hpp : HostPublish{nullptr, nullptr, 0, 0};
    PDWT_PROF(__func__, s);
    if (tab.nseg == 0) return 0;
    const int th = kEwThreads;
    dim3 grid(blocks_for(tab, batch, sums ? 3 : 4));
    if (sums) {   // thresholds that also deliver the norms of their result
        switch (op) {
            case 0: k_threshold<0, true><<<grid, th, 0, s>>>(tab, sums, batch, hp); break;
            case 1: k_threshold<1, true><<<grid, th, 0, s>>>(tab, sums, batch, hp); break;
            case 2: k_threshold<2, true><<<grid, th, 0, s>>>(tab, sums, batch, hp); break;
            default: k_threshold<3, true><<<grid, th, 0, s>>>(tab, sums, batch, hp); break;
        }
    } else {
        switch (op) {
            case 0: k_threshold<0, false><<<grid, th, 0, s>>>(tab, nullptr, batch, hp); break;
            case 1: k_threshold<1, false><<<grid, th, 0, s>>>(tab, nullptr, batch, hp); break;
            case 2: k_threshold<2, false><<<grid, th, 0, s>>>(tab, nullptr, batch, hp); break;
            default: k_threshold<3, false><<<grid, th, 0, s>>>(tab, nullptr, batch, hp); break;
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

int e_reduce(const SegTable& tab, int mode, int batch, double* d_sums, cudaStream_t s, const HostPublish* hpp)
{
    const HostPublish hp = hpp ? *hpp : HostPublish{nullptr, nullptr, 0, 0};
    PDWT_PROF(__func__, s);
    if (tab.nseg == 0) return 0;
    dim3 grid(blocks_for(tab, batch, 4));
    if (mode)
        k_reduce<1><<<grid, kEwThreads, 0, s>>>(tab, d_sums, batch, hp);
    else
        k_reduce<0><<<grid, kEwThreads, 0, s>>>(tab, d_sums, batch, hp);
    PDWT_LAUNCH_CHECK();
    return 0;
}

}  // namespace pdwt
