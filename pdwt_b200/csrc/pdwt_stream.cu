// pdwt_stream.cu -- warp-streaming kernels of the separable 2-D DWT for sm_100a (the headline path).
//
// One WARP is one worker.  It owns a strip of 64 output columns and marches down the rows of a chunk; nothing is
// shared between warps, so the kernels contain no block-level barrier at all.
//
//   forward (one level, reference separable.cu:91-176)
//     * input rows arrive in a per-warp shared-memory ring filled by the TMA engine (cp.async.bulk, one row piece per
//       copy, completion on an mbarrier); the periodic extension of separable.cu:114-121 is folded into the copies
//       (a wrapped row index, and a second copy for the columns that wrap), so the tap loops never test an index;
//     * row pass: each lane produces (lo, hi) of its 2 output columns with packed FFMA2 -- one multiplicand x
//       broadcast against the tap pair (L[j], H[j]) held in a uniform register pair, accumulator pair (lo, hi);
//     * column pass in scatter form: the fresh (lo, hi) pair is multiplied by the scalar taps L[j] / H[j] into the
//       hlen/2 pending output rows, accumulator pairs (A, V) and (H, D); a rotating register file of hlen/2 slots,
//       the loop body unrolled over hlen input rows so every register index is static;
//     * finished rows leave as 64-bit coalesced stores straight from registers.
//   inverse (one level, reference separable.cu:246-328)
//     * coefficient rows are read with coalesced 64-bit loads into a register window of hlen/2 (+1) rows, column
//       synthesis runs on (A, V) / (H, D) pairs against scalar taps, the two branch sums are added last;
//     * the two synthesised rows (t1, t2) go through a small per-warp shared-memory tile (swizzled, conflict-free)
//       so that each lane can do the row synthesis of 8 consecutive pixels of one row; 128-bit coalesced stores.
//
// Arithmetic contract: every output is the reference's own fmaf chain -- from 0, ascending tap index, row-pass
// result rounded to fp32 before the column pass, inverse branch sums added last -- so results are bit-identical to
// the reference CUDA build (fma.rn.f32x2 is two IEEE fp32 FMAs).  FFMA2 halves the issue slots the FP32 pipe needs,
// which is what lets shared-memory loads, address arithmetic and stores hide behind the FMAs (tools/ubench_fma.cu).
//
// Shapes these kernels take: even hlen in [4, 20], even Nr, Nc % 4 == 0, Nc large enough that a strip wraps at
// most once, 16-byte aligned planes.  Everything else goes to pdwt_fused.cu / pdwt_generic.cu.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "pdwt_common.cuh"

namespace pdwt {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// d = a * b + c on both halves, round-to-nearest (SASS FFMA2; broadcast forms are chosen by ptxas)
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---- mbarrier + bulk-copy (TMA engine) primitives ------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// non-blocking probe of the same condition: 1 if the phase with this parity has completed
__device__ __forceinline__ unsigned mbar_test(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// immediate (never suspending) probe, for polling several barriers in turn
__device__ __forceinline__ unsigned mbar_poll(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// global -> shared::cta, `bytes` multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only), and its hand-over to an mbarrier
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned bar)
{
    asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int round4(int n) { return (n + 3) & ~3; }

// dynamic shared memory that lets exactly two CTAs share an SM (228 KB per SM, 1 KB reserved per CTA)
constexpr size_t kTwoPerSmBytes = 112 * 1024;

static int sm_count()
{
    static const int n = []() {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) {
            cudaGetLastError();
            v = 148;
        }
        return v;
    }();
    return n;
}

// One lane of the (converged) warp, chosen by the hardware; ptxas keeps the guarded block uniform, so the bulk copies
// inside compile to a single UBLKCP each instead of a per-lane loop.
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ================================================================================================== forward
// Warp-specialised CTA: NCW consumer warps + ONE producer warp.
//   * producer: feeds every consumer's private ring through the TMA engine.  Issuing a TMA request stalls the issuing
//     warp for 100+ cycles (measured with clock64: two cp.async.bulk per row pair cost a consumer ~300 of its ~770
//     cycles), so no arithmetic warp ever issues one.  Interior super-slots take ONE tensor-map request
//     (cp.async.bulk.tensor.3d, box WW x SR x 1); super-slots that need the periodic extension of
//     separable.cu:114-121 (image edges) are staged row by row with cp.async.bulk pieces, wrapped columns included.
//   * consumer: owns a strip of 64 output columns, waits on its ring's "full" mbarrier, runs the FFMA2 row pass and
//     scatter column pass out of registers, and hands super-slots back through an "empty" mbarrier.
// Every loop bound derives from blockIdx (all consumers of a CTA share the row chunk) and the warp index is read through
// a shuffle, so ptxas can prove the control flow warp-uniform and keeps the filter taps in uniform registers
// (FFMA2 R, R, UR, R) and the loop control on the uniform datapath.
template <int HLEN>
struct FwdGeom {
    static constexpr int NC = 2;                        // output columns per lane
    static constexpr int H2 = HLEN / 2;
    static constexpr int C = H2 - 1;                    // analysis centre (even hlen), separable.cu:103-107
    static constexpr int AL = round4(C);                // the strip starts AL input columns left of 2*k0 (16-byte aligned)
    static constexpr int SH = AL - C;                   // first useful column inside the staged strip
    static constexpr int WO = 32 * NC;                  // output columns per consumer warp
    static constexpr int WW = round4(SH + 2 * WO + HLEN - 2);  // staged input columns per row
    static constexpr int NV = (SH + 2 * NC - 2 + HLEN + 3) / 4;  // 16-byte vectors a lane reads per row
    static constexpr int NCW = 4;                       // consumer warps per CTA
    static constexpr int SR = 8;                        // input rows per super-slot (one TMA request)
    static constexpr int NSS = 3;                       // super-slots per consumer ring
    static constexpr int SSB = SR * WW * 4;             // bytes per super-slot (multiple of 128)
    static constexpr int THREADS = (NCW + 1) * 32;
    static constexpr size_t SMEM = (size_t)NCW * NSS * SSB + NCW * NSS * 2 * sizeof(u64) + 128;
    static_assert(4 * (NV - 1) + 2 * NC * 31 + 4 <= WW, "lane window exceeds the staged strip");
    static_assert(SSB % 128 == 0 && WW <= 256 && SR % 2 == 0, "tensor-map box constraints");
};

template <int HLEN>
struct FwdParams {
    CUtensorMap tm;   // (Nc, Nr, batch) fp32 tensor over the source planes, box (WW, SR, 1); valid iff use_tm
    float2 lh[HLEN];  // (L[hlen-1-j], H[hlen-1-j]): row-pass tap pairs in the reference's accumulation order
    float ly[HLEN];   // L[hlen-1-j]
    float hy[HLEN];   // H[hlen-1-j]
    const float* src;
    float *A, *Hb, *V, *D;
    size_t s_src, s_a, s_d;  // plane strides (floats)
    int Nr, Nc, nr, nc;      // input and output plane sizes
    int TH;                  // output rows per chunk
    int ncg, nrc;            // column groups (NCW strips each), row chunks; grid = ncg * nrc * batch CTAs
    int use_tm;
    int pdl_early;           // PDWT_PDL=1: let the next kernel's CTAs in as soon as this one has started
    unsigned poll_ns;        // producer: sleep between two rounds of polling that found no free ring slot
};

__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* tm, int x, int y, int z, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

#ifdef PDWT_EXPERIMENTS
// timeline instrumentation (PDWT_EXPERIMENTS builds only): 8 globaltimer stamps for each of the first 1024 CTAs
__device__ unsigned long long g_timeline[4096 * 8];
__device__ __forceinline__ void tl_stamp(int cta, int slot)
{
    if (cta < 4096) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_timeline[cta * 8 + slot] = t;
    }
}
#define TL(slot, cond) do { if (cond) tl_stamp(blockIdx.x, slot); } while (0)
#else
#define TL(slot, cond) do { } while (0)
#endif

// LOWOCC: variant for grids that put at most 2 CTAs on an SM anyway (one image): it may use twice the registers, which
// buys an earlier prefetch of the next pair's first row (its shared-memory latency hides behind the second row's FFMA2s
// instead of sitting at the head of the next iteration; with the 4-CTA register budget that spills).
template <int HLEN, bool LOWOCC>
__global__ void __launch_bounds__(FwdGeom<HLEN>::THREADS, LOWOCC ? 2 : (HLEN <= 14 ? 4 : 3))
    k_fwd2d_stream(const __grid_constant__ FwdParams<HLEN> p)
{
    using G = FwdGeom<HLEN>;
    constexpr int H2 = G::H2, NC = G::NC, NSS = G::NSS, SR = G::SR, NCW = G::NCW;
    extern __shared__ unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
    const int item = blockIdx.x;

    const unsigned ring_s = (smem_u32(smem_raw) + 127u) & ~127u;   // tensor-map destinations are 128-byte aligned
    const unsigned bar_s = ring_s + NCW * NSS * G::SSB;            // full[w][s] at +16*(w*NSS+s), empty right behind it

    const int cg = item % p.ncg, rest = item / p.ncg, rc = rest % p.nrc, plane = rest / p.nrc;
    const int y0 = rc * p.TH;
    const int ny = min(p.TH, p.nr - y0);
    const int npairs = ny + H2 - 1;          // input row pairs this chunk consumes
    const int nss = (2 * npairs + SR - 1) / SR;   // super-slots per consumer
    const int vr0 = 2 * y0 - G::C;           // first input row (virtual: < 0 or >= Nr wraps around)
    const int nstrips = min(NCW, (p.nc - cg * NCW * G::WO + G::WO - 1) / G::WO);   // strips of this CTA inside the image
    TL(0, threadIdx.x == 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NCW * NSS * 2; i++) mbar_init(bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();   // the only block-wide barrier: after it the warps only meet through mbarriers
    if (p.pdl_early) pdl_launch_dependents();
    pdl_wait();        // the previous level's kernel (or whatever wrote `src`) has completed
    TL(1, threadIdx.x == 0);

    if (warp == NCW) {
        // ===================================================================================== producer warp
        const float* src = p.src + (size_t)plane * p.s_src;
        // super-slot k of strip w -> ring slot k % NSS
        auto issue = [&](const int k, const int w) {
            const int slot = k % NSS;
            int row0 = vr0 + SR * k;
            // rows needed from this super-slot all inside the image? (rows past the chunk's last pair may fall outside:
            // the tensor request zero-fills them and nobody reads them)
            const int last_needed = min(row0 + SR, vr0 + 2 * npairs) - 1;
            const bool rows_in = row0 >= 0 && last_needed < p.Nr;
            row0 += (row0 < 0) ? p.Nr : 0;
            row0 -= (row0 >= p.Nr) ? p.Nr : 0;
            const unsigned full = bar_s + 16 * (w * NSS + slot);
            const int xs = 2 * (cg * NCW + w) * G::WO - G::AL;
            const unsigned dst = ring_s + (w * NSS + slot) * G::SSB;
            if (p.use_tm && rows_in && xs >= 0 && xs + G::WW <= p.Nc) {
                if (elect_one()) {
                    mbar_expect_tx(full, G::SSB);
                    tma_load_3d(dst, &p.tm, xs, vr0 + SR * k, plane, full);
                }
            } else {
                // periodic extension (separable.cu:114-121, even sizes): every lane copies 16-byte pieces with a
                // wrapped row and column index (xs, Nc and WW are multiples of 4, so a piece never straddles the
                // wrap), then hands its copies to the barrier.  LDGSTS instead of per-row bulk copies: a bulk copy
                // costs the issuing warp ~120 cycles and an edge super-slot would need 16 of them.
                constexpr int CPR = G::WW / 4;   // pieces per row
                constexpr int NIT = (SR * CPR + 31) / 32;
                const float* g[NIT];
#pragma unroll
                for (int it = 0; it < NIT; it++) {   // all addresses first, then the copies back to back
                    const int i = lane + 32 * it;
                    const int r = i / CPR, c = i - r * CPR;
                    int row = row0 + r;
                    row -= (row >= p.Nr) ? p.Nr : 0;
                    int x = xs + 4 * c;
                    x += (x < 0) ? p.Nc : 0;
                    x -= (x >= p.Nc) ? p.Nc : 0;
                    g[it] = src + (size_t)row * p.Nc + x;
                }
#pragma unroll
                for (int it = 0; it < NIT; it++)
                    if (lane + 32 * it < SR * CPR) cp_async16(dst + (lane + 32 * it) * 16, g[it]);
                cp_async_mbar_arrive(full);   // +1 pending now, -1 when this lane's copies have landed
                __syncwarp();
                if (lane == 0) mbar_arrive(full);   // the barrier's own count of 1
            }
            __syncwarp();
        };
        // Round-robin over the strips: a strip is served as soon as ITS consumer has handed the ring slot back, so one
        // slow consumer does not hold up the refills of the others (the consumers of a CTA drift apart by whole pairs).
        int next[NCW];
#pragma unroll
        for (int w = 0; w < NCW; w++) next[w] = 0;
        int remaining = nstrips * nss;
        while (remaining > 0) {
            bool any = false;
#pragma unroll
            for (int w = 0; w < NCW; w++) {
                const int k = next[w];
                if (w < nstrips && k < nss) {
                    unsigned ok = 1;
                    if (k >= NSS)   // parity of the consumer's (k/NSS)-th release of this slot
                        ok = mbar_poll(bar_s + 16 * (w * NSS + k % NSS) + 8, ((k / NSS) + 1) & 1);
                    if (__shfl_sync(0xffffffffu, ok, 0)) {
                        issue(k, w);
                        next[w] = k + 1;
                        remaining--;
                        any = true;
                    }
                }
            }
            TL(2, lane == 0 && next[0] == 1 && remaining == nstrips * (nss - 1));
            if (!any) __nanosleep(p.poll_ns);
        }
#ifdef PDWT_EXPERIMENTS
        if (lane == 0 && blockIdx.x < 1024) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_timeline[blockIdx.x * 8 + 7] = smid;   // slot 7: the SM this CTA ran on
        }
#endif
        return;
    }

    // ========================================================================================= consumer warps
    if (warp >= nstrips) return;
    const int k0 = (cg * NCW + warp) * G::WO;
    const unsigned my_ring = ring_s + warp * NSS * G::SSB;
    const unsigned my_bar = bar_s + 16 * warp * NSS;

    // pending output rows: acc[pp] belongs to the output row that is pp row pairs back; pairs (A,V) and (H,D) per column.
    // The file SHIFTS by one slot per row pair, and the shift rides on the second row's FFMA2s (destination slot pp+1,
    // addend slot pp), so every register index is static with a loop body of ONE row pair: the code stays a few KB
    // instead of the 88 KB of a body unrolled over hlen rows (instruction-fetch stalls were 13 % of the warp time of
    // the batched level-1 kernel, profiles/r01m_ncu_full_summary.txt).
    u64 aLV[H2 + 1][NC], aHD[H2 + 1][NC];
#pragma unroll
    for (int s = 0; s <= H2; s++)
#pragma unroll
        for (int c = 0; c < NC; c++) aLV[s][c] = aHD[s][c] = 0ull;

    const int kcol = k0 + NC * lane;                 // first output column of this lane
    const bool col_ok = kcol < p.nc;                 // nc is even, so the pair is in or out as a whole
    const size_t o0 = (size_t)y0 * p.nc + kcol;
    float* oA = p.A + (size_t)plane * p.s_a + o0;
    float* oH = p.Hb + (size_t)plane * p.s_d + o0;
    float* oV = p.V + (size_t)plane * p.s_d + o0;
    float* oD = p.D + (size_t)plane * p.s_d + o0;
    // this lane's window inside a staged row, as a generic pointer (plain loads keep their order w.r.t. the barriers)
    const char* lane_ring = static_cast<const char*>(__cvta_shared_to_generic(my_ring)) + 2 * NC * lane * 4;

    // row pass, w_kern_forward_pass1 (separable.cu:91-131): (lo, hi)[c] = sum_j x[2k - C + j] * (L, H)[hlen-1-j]
    auto row_pass = [&](const float (&xv)[G::NV * 4], u64 (&lohi)[NC]) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
            u64 acc = 0ull;
#pragma unroll
            for (int j = 0; j < HLEN; j++) {
                const float x = xv[G::SH + 2 * c + j];
                acc = ffma2(pack2(x, x), pack2(p.lh[j].x, p.lh[j].y), acc);
            }
            lohi[c] = acc;
        }
    };
    auto load_row = [&](float (&xv)[G::NV * 4], const char* rowp) {
        const float4* rp = reinterpret_cast<const float4*>(rowp);
#pragma unroll
        for (int i = 0; i < G::NV; i++) {
            const float4 f = rp[i];
            xv[4 * i] = f.x; xv[4 * i + 1] = f.y; xv[4 * i + 2] = f.z; xv[4 * i + 3] = f.w;
        }
    };

    int q = 0;                                    // row pair index within the chunk
    unsigned soff = 0, bar = my_bar, parity = 0;  // ring position (byte offset, full barrier, phase) of the super-slot
    for (int k = 0; k < nss; k++) {
        mbar_wait(bar, parity);
        TL(3, threadIdx.x == 0 && k == 0);
        const char* ssp = lane_ring + soff;
#pragma unroll
        for (int pin = 0; pin < SR / 2; pin++) {
            if (q < npairs) {                     // warp-uniform; false only in the tail of the chunk's last super-slot
                float xa[G::NV * 4], xb[G::NV * 4];
                load_row(xa, ssp + pin * (2 * G::WW * 4));
                load_row(xb, ssp + pin * (2 * G::WW * 4) + G::WW * 4);
                if (q + 1 >= npairs) pdl_launch_dependents();   // last row pair of this warp
                u64 lohi0[NC], lohi1[NC];
                row_pass(xa, lohi0);
                // column pass, w_kern_forward_pass2 (separable.cu:135-176), scatter form: the first row of the pair is
                // tap j = 2*pp of the output row pp pairs back (in place) ...
#pragma unroll
                for (int pp = 0; pp < H2; pp++) {
                    const u64 kl = pack2(p.ly[2 * pp], p.ly[2 * pp]), kh = pack2(p.hy[2 * pp], p.hy[2 * pp]);
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        aLV[pp][c] = ffma2(lohi0[c], kl, pp == 0 ? 0ull : aLV[pp][c]);
                        aHD[pp][c] = ffma2(lohi0[c], kh, pp == 0 ? 0ull : aHD[pp][c]);
                    }
                }
                row_pass(xb, lohi1);
                // ... the second row is tap j = 2*pp + 1, and its result moves one slot up: next pair it is pp+1 back
#pragma unroll
                for (int pp = H2 - 1; pp >= 0; pp--) {
                    const u64 kl = pack2(p.ly[2 * pp + 1], p.ly[2 * pp + 1]), kh = pack2(p.hy[2 * pp + 1], p.hy[2 * pp + 1]);
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        aLV[pp + 1][c] = ffma2(lohi1[c], kl, aLV[pp][c]);
                        aHD[pp + 1][c] = ffma2(lohi1[c], kh, aHD[pp][c]);
                    }
                }
                TL(4, q == 0 && threadIdx.x == 0);
                TL(5, q == H2 - 1 && threadIdx.x == 0);
                // slot H2 received its last tap (j = hlen-1) in this pair
                if (q >= H2 - 1) {
                    if (col_ok) {
                        float a0, v0, a1, v1, h0, d0, h1, d1;
                        unpack2(aLV[H2][0], a0, v0);
                        unpack2(aLV[H2][1], a1, v1);
                        unpack2(aHD[H2][0], h0, d0);
                        unpack2(aHD[H2][1], h1, d1);
                        *reinterpret_cast<float2*>(oA) = make_float2(a0, a1);
                        *reinterpret_cast<float2*>(oH) = make_float2(h0, h1);
                        *reinterpret_cast<float2*>(oV) = make_float2(v0, v1);
                        *reinterpret_cast<float2*>(oD) = make_float2(d0, d1);
                    }
                    oA += p.nc; oH += p.nc; oV += p.nc; oD += p.nc;
                }
                q++;
            }
        }
        // every lane has consumed all rows of this super-slot: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + 8);
        soff += G::SSB;
        bar += 16;
        if (soff == NSS * G::SSB) {
            soff = 0;
            bar = my_bar;
            parity ^= 1;
        }
    }
    TL(6, threadIdx.x == 0);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (EncodeTiledFn)f;
    }();
    return fn;
}

template <int HLEN>
static int launch_fwd_stream(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                             cudaStream_t s)
{
    using G = FwdGeom<HLEN>;
    const int nr = Nr / 2, nc = Nc / 2;
    // shapes the TMA staging can serve (see the file header); 0 = "not handled"
    if ((Nr & 1) || (Nc & 3) || Nc < G::WW || Nr < HLEN) return 0;
    if ((((uintptr_t)src.p) & 15) || (src.stride & 3)) return 0;
    if ((((uintptr_t)A.p | (uintptr_t)H.p | (uintptr_t)V.p | (uintptr_t)D.p) & 7) || (A.stride & 1) || (H.stride & 1))
        return 0;
    static PerDeviceOnce once;
    static int per_sm_dev[64];   // resident CTAs per SM (standard variant), per device
    const bool first = once.first();
    int& per_sm = per_sm_dev[once.dev];
    if (first) {
        PDWT_CUDA(cudaFuncSetAttribute(k_fwd2d_stream<HLEN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        PDWT_CUDA(cudaFuncSetAttribute(k_fwd2d_stream<HLEN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(G::SMEM > kTwoPerSmBytes ? G::SMEM : kTwoPerSmBytes)));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fwd2d_stream<HLEN, false>, G::THREADS, G::SMEM) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 3;
        }
    }
    FwdParams<HLEN> p;
    memset(&p.tm, 0, sizeof p.tm);
    p.use_tm = 0;
    if (EncodeTiledFn enc = encode_tiled_fn()) {
        const cuuint64_t dims[3] = {(cuuint64_t)Nc, (cuuint64_t)Nr, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {(cuuint64_t)Nc * 4, (cuuint64_t)src.stride * 4};  // bytes, dims 1 and 2
        const cuuint32_t box[3] = {(cuuint32_t)G::WW, (cuuint32_t)G::SR, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)src.p, dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            p.use_tm = (getenv("PDWT_NO_TENSORMAP") == nullptr);
    }
    for (int j = 0; j < HLEN; j++) {
        p.lh[j] = make_float2(t.L[HLEN - 1 - j], t.H[HLEN - 1 - j]);
        p.ly[j] = t.L[HLEN - 1 - j];
        p.hy[j] = t.H[HLEN - 1 - j];
    }
    p.src = src.p; p.A = A.p; p.Hb = H.p; p.V = V.p; p.D = D.p;
    p.s_src = src.stride; p.s_a = A.stride; p.s_d = H.stride;
    p.Nr = Nr; p.Nc = Nc; p.nr = nr; p.nc = nc;
    p.ncg = idiv_up(nc, G::WO * G::NCW);
    // Chunk height.  A CTA (NCW strips x TH output rows) costs TH + hlen/2 - 1 row pairs per consumer warp (the vertical
    // halo is row-pass work only, but the scatter accumulators need the same warm-up), an SM holds up to 4 CTAs and
    // works through the pairs of its resident warps at a roughly constant rate, so the kernel ends when the busiest SM
    // does: minimise (CTAs on the busiest SM) x (pairs per CTA).  For one 4096^2 image that picks TH = 56 (296 CTAs = 2
    // per SM) instead of a power of two that leaves 40 SMs with half the work of the others.
    int TH = 64;
    {
        const int sms = sm_count();
        double best = 1e30;
        for (int th = 128; th >= 4; th -= 2) {
            const long long ctas = (long long)p.ncg * idiv_up(nr, th) * batch;
            const long long full = ctas / ((long long)sms * per_sm), rest = ctas % ((long long)sms * per_sm);
            const double units = (double)(full * per_sm + (rest + sms - 1) / sms) * (th + G::H2 - 1);
            // one CTA per SM = a lone warp per scheduler: measured ~650 cycles per row pair against ~435 per scheduler
            // with two or more warps sharing it
            const double cost = units * (ctas <= sms ? 1.5 : 1.0);
            if (cost < best * 0.999) {
                best = cost;
                TH = th;
            }
        }
    }
    if (const char* e = getenv("PDWT_TH")) TH = atoi(e) > 0 ? atoi(e) : TH;
    p.TH = TH;
    // a consumer needs > 1 us per super-slot and two more are staged behind it: the producer can afford to sleep
    static const unsigned poll_ns = []() { const char* e = getenv("PDWT_POLL_NS"); return e ? (unsigned)atoi(e) : 200u; }();
    p.poll_ns = poll_ns;
    p.pdl_early = pdl_mode() == 1;
    p.nrc = idiv_up(nr, TH);
    const long long nctas = (long long)p.ncg * p.nrc * batch;
    if (nctas > 0x7fffffff) return 0;
    PDWT_PROF(prof_tag("k_fwd2d_stream", Nr, Nc), s);
    // at most 2 CTAs per SM: the high-register variant loses no occupancy (PDWT_LOWOCC=0|1 forces the choice)
    bool lowocc = nctas <= 2LL * sm_count();
    if (const char* e = getenv("PDWT_LOWOCC")) lowocc = atoi(e) != 0;
    // ... and it asks for so much shared memory that NO SM can take a third CTA: the block scheduler does not spread a
    // grid of 2 x SMs CTAs evenly by itself, and the kernel ends with the busiest SM (PDWT_TWOPERSM=0 switches it off)
    static const bool cap2 = []() { const char* e = getenv("PDWT_TWOPERSM"); return !e || atoi(e) != 0; }();
    if (lowocc)
        PDWT_CUDA(launch_pdl(k_fwd2d_stream<HLEN, true>, dim3((unsigned)nctas), G::THREADS,
                             cap2 && G::SMEM < kTwoPerSmBytes ? kTwoPerSmBytes : G::SMEM, s, p));
    else
        PDWT_CUDA(launch_pdl(k_fwd2d_stream<HLEN, false>, dim3((unsigned)nctas), G::THREADS, G::SMEM, s, p));
    PDWT_LAUNCH_CHECK();
    return 1;
}

// ================================================================================================== inverse
// Index rules of the synthesis passes (separable.cu:246-328), per axis, for the output pair
// (2m, 2m+1):   even output: coefficients m-CC+j,        taps  IL/IH[hlen-1-(2j+1-SHIFT)]
//               odd  output: coefficients m-CC+SHIFT+j,  taps  IL/IH[hlen-1-(2j+SHIFT)]        j = 0 .. hlen/2-1
// with CC = (hlen/2)/2 and SHIFT = 1 when hlen/2 is even (the reference's "virtual id for shift").
template <int HLEN>
struct InvGeom {
    static constexpr int H2 = HLEN / 2;
    static constexpr int CC = H2 / 2;
    static constexpr int SHIFT = (H2 & 1) ? 0 : 1;
    static constexpr int WIN = H2 + SHIFT;              // coefficient rows (columns) behind one output pair
    static constexpr int NSLOT = WIN + 1;               // register window: WIN rows + the row being pulled from the ring
    static constexpr int UNR = NSLOT * (NSLOT >= 7 ? 1 : 2);  // row pairs per unrolled loop body (even, >= 8)
    static constexpr int RS = UNR;                      // ring slots = the unroll period, so every ring address is static
    static constexpr int DEPTH = 7;                     // coefficient rows in flight global -> shared (cp.async groups)
    static constexpr int ROWB = 4 * 64 * 4;             // ring bytes per coefficient row: A,H,V,D x 64 columns
    static constexpr int ALC = (CC + 1) & ~1;           // the strip starts ALC coefficient columns left of k0 (even)
    static constexpr int SHC = ALC - CC;
    static constexpr int NP = (SHC + WIN + 3 + 1) & ~1; // (t1,t2) pairs a row-synthesis lane reads (4 coefficient columns)
    static constexpr int WOUT = 4 * ((64 - NP) / 4) + 4; // coefficient columns a warp turns into pixels
    static constexpr int LPR = WOUT / 4;                // row-synthesis lanes per output row (<= 16)
    // tile row: 32 chunks of 16 bytes (2 (t1,t2) pairs each).  Even chunks sit at positions 0..15, odd chunks at 20..35:
    // the writers (lane -> chunk lane) and the readers (lane -> chunks 2*lane' + v, i.e. CONSECUTIVE positions
    // lane' + v/2 in one of the halves) are both bank-conflict free and every offset is a compile-time immediate
    static constexpr int ODD0 = 20;
    static constexpr int TROWB = (ODD0 + 16) * 16;      // bytes per tile row
    static constexpr size_t TILEB = 2 * 2 * TROWB;      // double-buffered tile: 2 output rows
    static constexpr size_t SMEM = TILEB + (size_t)RS * ROWB;
    static_assert(LPR <= 16 && WOUT - 4 + NP <= 64, "row-synthesis window exceeds the strip");
    static_assert(UNR % 2 == 0 && UNR % NSLOT == 0 && DEPTH < RS, "static ring / tile addressing");
};

template <int HLEN>
struct InvParams {
    float il[2][HLEN / 2], ih[2][HLEN / 2];  // [output parity][j]: IL / IH taps in accumulation order
    float2 lh[2][HLEN / 2];                  // the same as (IL, IH) pairs for the row synthesis
    const float *A, *Hb, *V, *D;
    float* dst;
    size_t s_a, s_d, s_dst;                  // plane strides (floats)
    int nr, nc, Mr, Mc;                      // coefficient and output plane sizes (Mr = 2 nr, Mc = 2 nc)
    int TM;                                  // output row PAIRS per chunk
    int ncb, nrc;
    int pdl_early;
};


// 8-byte asynchronous copy global -> shared (LDGSTS), tracked by cp.async groups
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int HLEN>
__global__ void __launch_bounds__(32, 12) k_inv2d_stream(const __grid_constant__ InvParams<HLEN> p)
{
    using G = InvGeom<HLEN>;
    constexpr int H2 = G::H2, NSLOT = G::NSLOT, WIN = G::WIN, SHIFT = G::SHIFT, UNR = G::UNR, DEPTH = G::DEPTH;
    extern __shared__ __align__(16) unsigned char smem_inv[];
    const int lane = threadIdx.x;
    const int item = blockIdx.x;
    const int cb = item % p.ncb, rest = item / p.ncb, rc = rest % p.nrc, plane = rest / p.nrc;
    const int k0 = cb * G::WOUT, m0 = rc * p.TM;
    const int nm = min(p.TM, p.nr - m0);

    // ---- column synthesis side: this lane owns coefficient columns (col, col+1) of the strip, wrapped periodically
    int col = k0 - G::ALC + 2 * lane;
    col += (col < 0) ? p.nc : 0;
    col -= (col >= p.nc) ? p.nc : 0;
    int lrow = m0 - G::CC;                    // next coefficient row to load (wrapped: separable.cu:265-273)
    lrow += (lrow < 0) ? p.nr : 0;
    int to_wrap = p.nr - lrow;
    // ONE running per-lane pointer (into A) and three uniform byte distances to the same element of H, V, D: per row
    // that is a 64-bit add for each address and one for the advance, nothing else
    const char* pa = reinterpret_cast<const char*>(p.A + (size_t)plane * p.s_a + (size_t)lrow * p.nc + col);
    const ptrdiff_t plane_d = (ptrdiff_t)((size_t)plane * p.s_d) - (ptrdiff_t)((size_t)plane * p.s_a);
    const ptrdiff_t dH = (reinterpret_cast<const char*>(p.Hb) - reinterpret_cast<const char*>(p.A)) + 4 * plane_d;
    const ptrdiff_t dV = (reinterpret_cast<const char*>(p.V) - reinterpret_cast<const char*>(p.A)) + 4 * plane_d;
    const ptrdiff_t dD = (reinterpret_cast<const char*>(p.D) - reinterpret_cast<const char*>(p.A)) + 4 * plane_d;
    const ptrdiff_t row_b = 4 * (ptrdiff_t)p.nc, wrap_b = 4 * (ptrdiff_t)p.nc * p.nr;

    // Coefficient rows travel global -> shared -> registers.  Each lane prefetches ITS OWN two columns of A, H, V, D
    // DEPTH rows ahead with 8-byte cp.async copies into a private 32-byte cell per ring row, and later reads the same
    // cell back: the ring is a latency buffer (DEPTH KB per warp in flight), nothing in it is shared between lanes, so
    // it needs no barrier -- only cp.async group accounting (one group per row, empty past the chunk's last row).
    // Row r of the chunk lives in ring slot r % RS; RS equals the unroll period of the main loop, so every slot index
    // below is a compile-time constant.
    const unsigned ring_s = smem_u32(smem_inv) + (unsigned)G::TILEB + 8 * lane;
    const float* ring_g = reinterpret_cast<const float*>(smem_inv + G::TILEB) + 2 * lane;
    int rows_left = nm + WIN - 1;             // coefficient rows of this chunk not yet requested
    auto issue_row = [&](const int slot) {
        if (rows_left > 0) {
            const unsigned dst = ring_s + slot * G::ROWB;
            cp_async8(dst, pa);
            cp_async8(dst + 256, pa + dH);
            cp_async8(dst + 512, pa + dV);
            cp_async8(dst + 768, pa + dD);
            pa += row_b;
            if (--to_wrap == 0) {
                pa -= wrap_b;
                to_wrap = p.nr;
            }
        }
        cp_async_commit();
        rows_left--;
    };
    u64 wA[NSLOT], wH[NSLOT], wV[NSLOT], wD[NSLOT];  // register window, slot = (row index within the chunk) % NSLOT
    // row r of the chunk: ring slot r % RS -> window slot r % NSLOT, then request row r + DEPTH
    auto load_row = [&](const int r) {
        cp_async_wait<DEPTH - 1>();           // all but the newest DEPTH-1 groups have landed: row r is there
        const float* c = ring_g + (r % G::RS) * (G::ROWB / 4);
        const float2 a = *reinterpret_cast<const float2*>(c);
        const float2 h = *reinterpret_cast<const float2*>(c + 64);
        const float2 v = *reinterpret_cast<const float2*>(c + 128);
        const float2 d = *reinterpret_cast<const float2*>(c + 192);
        wA[r % NSLOT] = pack2(a.x, a.y);
        wH[r % NSLOT] = pack2(h.x, h.y);
        wV[r % NSLOT] = pack2(v.x, v.y);
        wD[r % NSLOT] = pack2(d.x, d.y);
        issue_row((r + DEPTH) % G::RS);
    };
#pragma unroll
    for (int i = 0; i < NSLOT; i++) wA[i] = wH[i] = wV[i] = wD[i] = 0ull;
    TL(0, lane == 0);
    if (p.pdl_early) pdl_launch_dependents();
    pdl_wait();        // the previous level's kernel (or whatever wrote the coefficients) has completed
    TL(1, lane == 0);
#pragma unroll
    for (int i = 0; i < DEPTH; i++) issue_row(i);
#pragma unroll
    for (int i = 0; i < WIN; i++) load_row(i);

    // ---- row synthesis side: lanes 0..LPR-1 take the even output row, lanes 16..16+LPR-1 the odd one; each turns 4
    // coefficient columns into 8 pixels
    const int g = lane >> 4, lq = lane & 15;
    const int px0 = 2 * k0 + 8 * lq;
    const bool row_lane = lq < G::LPR;
    float* out = p.dst + (size_t)plane * p.s_dst + (size_t)(2 * m0 + g) * p.Mc + px0;
    const bool st0 = row_lane && px0 + 4 <= p.Mc, st1 = row_lane && px0 + 8 <= p.Mc;
    // tile addressing (see InvGeom): the writer stores chunk `lane`, the reader loads chunks 2*lq + v
    unsigned char* const tile_wr = smem_inv + 16 * ((lane >> 1) + (lane & 1) * G::ODD0);
    const unsigned char* const tile_rd = smem_inv + g * G::TROWB + 16 * lq;

    int s = 0;
    for (;;) {
#pragma unroll
        for (int u = 0; u < UNR; u++) {       // body: UNR output row pairs; register, ring and tile indices all static
#ifdef PDWT_EXPERIMENTS
            if (s == 1) TL(3, lane == 0);
            if (s >= nm && lane == 0 && blockIdx.x < 4096) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                g_timeline[blockIdx.x * 8 + 7] = smid;
                tl_stamp(blockIdx.x, 6);
            }
#endif
            if (s >= nm) return;
            if (s + 1 >= nm) pdl_launch_dependents();
            if (s + 1 < nm) load_row(u + WIN);   // the row that the NEXT pair adds to the window (one pair of slack)
            const int sb = u % NSLOT, buf = u & 1;
            // column synthesis, w_kern_inverse_pass1 (separable.cu:246-289): t1 = IL_y(A) + IH_y(H), t2 = IL_y(V) + IH_y(D)
#pragma unroll
            for (int par = 0; par < 2; par++) {
                u64 sa = 0ull, sh = 0ull, sv = 0ull, sd = 0ull;
#pragma unroll
                for (int j = 0; j < H2; j++) {
                    const int sl = (sb + (par ? SHIFT : 0) + j) % NSLOT;
                    const u64 kl = pack2(p.il[par][j], p.il[par][j]), kh = pack2(p.ih[par][j], p.ih[par][j]);
                    sa = ffma2(wA[sl], kl, sa);
                    sh = ffma2(wH[sl], kh, sh);
                    sv = ffma2(wV[sl], kl, sv);
                    sd = ffma2(wD[sl], kh, sd);
                }
                float t1a, t1b, t2a, t2b;
                unpack2(fadd2(sa, sh), t1a, t1b);
                unpack2(fadd2(sv, sd), t2a, t2b);
                *reinterpret_cast<float4*>(tile_wr + (buf * 2 + par) * G::TROWB) = make_float4(t1a, t2a, t1b, t2b);
            }
            __syncwarp();
            // row synthesis, w_kern_inverse_pass2 (separable.cu:293-328): img = IL_x(t1) + IH_x(t2)
            if (row_lane) {
                u64 tw[G::NP];
#pragma unroll
                for (int v = 0; v < G::NP / 2; v++) {
                    const float4 f = *reinterpret_cast<const float4*>(tile_rd + buf * 2 * G::TROWB +
                                                                      16 * ((v >> 1) + (v & 1) * G::ODD0));
                    tw[2 * v] = pack2(f.x, f.y);
                    tw[2 * v + 1] = pack2(f.z, f.w);
                }
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int e = k & 1;
                    const int i0 = (k >> 1) + G::SHC + (e ? SHIFT : 0);
                    u64 acc = 0ull;
#pragma unroll
                    for (int j = 0; j < H2; j++) acc = ffma2(tw[i0 + j], pack2(p.lh[e][j].x, p.lh[e][j].y), acc);
                    float r1, r2;
                    unpack2(acc, r1, r2);
                    o[k] = __fadd_rn(r1, r2);
                }
                if (st0) *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
                if (st1) *reinterpret_cast<float4*>(out + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
            out += 2 * (size_t)p.Mc;
            s++;
        }
    }
}

template <int HLEN>
static int launch_inv_stream(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int nr, int nc, int Mr,
                             int Mc, int batch, cudaStream_t s)
{
    using G = InvGeom<HLEN>;
    if (Mr != 2 * nr || Mc != 2 * nc || (nc & 1) || nc < 64 || nr < G::WIN) return 0;
    if ((((uintptr_t)A.p | (uintptr_t)H.p | (uintptr_t)V.p | (uintptr_t)D.p) & 7) || (A.stride & 1) || (H.stride & 1))
        return 0;
    if ((((uintptr_t)dst.p) & 15) || (dst.stride & 3)) return 0;
    InvParams<HLEN> p;
    for (int par = 0; par < 2; par++) {
        const int off = par ? G::SHIFT : 1 - G::SHIFT;
        for (int j = 0; j < HLEN / 2; j++) {
            p.il[par][j] = t.IL[HLEN - 1 - (2 * j + off)];
            p.ih[par][j] = t.IH[HLEN - 1 - (2 * j + off)];
            p.lh[par][j] = make_float2(p.il[par][j], p.ih[par][j]);
        }
    }
    p.A = A.p; p.Hb = H.p; p.V = V.p; p.D = D.p; p.dst = dst.p;
    p.s_a = A.stride; p.s_d = H.stride; p.s_dst = dst.stride;
    p.nr = nr; p.nc = nc; p.Mr = Mr; p.Mc = Mc;
    p.ncb = idiv_up(nc, G::WOUT);
    // Chunk height (output row PAIRS per one-warp CTA).  An item costs TM loop iterations plus a prologue worth ~2.5 of
    // them (WIN-1 extra coefficient rows and the pipeline fill); items are spread over sms x per_sm resident warps, so
    // the kernel takes about ceil(items / slots) x (TM + 2.5): pick the TM that minimises it (4096^2: TM = 32, exactly
    // one item per slot).
    static PerDeviceOnce once;
    static int per_sm_dev[64];
    const bool first = once.first();
    int& per_sm = per_sm_dev[once.dev];
    if (first) {
        PDWT_CUDA(cudaFuncSetAttribute(k_inv2d_stream<HLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_inv2d_stream<HLEN>, 32, G::SMEM) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 12;
        }
    }
    int TM = 32;
    {
        const long long slots = (long long)sm_count() * per_sm;
        double best = 1e30;
        for (int tm = 64; tm >= 2; tm--) {
            const long long items = (long long)p.ncb * idiv_up(nr, tm) * batch;
            const double cost = (double)((items + slots - 1) / slots) * (tm + 2.5);
            if (cost < best * 0.999) {
                best = cost;
                TM = tm;
            }
        }
    }
    if (const char* e = getenv("PDWT_TM")) TM = atoi(e) > 0 ? atoi(e) : TM;
    p.TM = TM;
    p.nrc = idiv_up(nr, TM);
    p.pdl_early = pdl_mode() == 1;
    const long long nitems = (long long)p.ncb * p.nrc * batch;
    if (nitems > 0x7fffffff) return 0;
    PDWT_PROF(prof_tag("k_inv2d_stream", Mr, Mc), s);
    PDWT_CUDA(launch_pdl(k_inv2d_stream<HLEN>, dim3((unsigned)nitems), 32, G::SMEM, s, p));
    PDWT_LAUNCH_CHECK();
    return 1;
}

#ifdef PDWT_EXPERIMENTS
}  // namespace pdwt
extern "C" int pdwt_debug_timeline(unsigned long long* out, int n)
{
    if (n > 4096 * 8) n = 4096 * 8;
    return (int)cudaMemcpyFromSymbol(out, pdwt::g_timeline, sizeof(unsigned long long) * n);
}
namespace pdwt {
#endif

#define PDWT_STREAM_HLEN_SWITCH(fn, ...)            \
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

int s_dwt2_fwd_level(const Taps& t, Plane2 src, Plane2 A, Plane2 H, Plane2 V, Plane2 D, int Nr, int Nc, int batch,
                     cudaStream_t s)
{
    PDWT_STREAM_HLEN_SWITCH(launch_fwd_stream, t, src, A, H, V, D, Nr, Nc, batch, s)
}

int s_dwt2_inv_level(const Taps& t, Plane2 A, Plane2 H, Plane2 V, Plane2 D, Plane2 dst, int nr, int nc, int Mr, int Mc,
                     int batch, cudaStream_t s)
{
    PDWT_STREAM_HLEN_SWITCH(launch_inv_stream, t, A, H, V, D, dst, nr, nc, Mr, Mc, batch, s)
}

}  // namespace pdwt
