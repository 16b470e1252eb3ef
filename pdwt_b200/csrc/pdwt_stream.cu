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

#include <algorithm>
#include <functional>
#include <mutex>
#include <queue>
#include <vector>

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

// ---- cross-level execution (one launch per direction for all levels of a transform) --------------------------------
// The reference queues one pair of kernels per level, each waiting for the previous one's last thread block
// (separable.cu:179-209, 332-364).  Here the work items of ALL levels (and all planes of a batch) are served by ONE
// launch with a small dataflow runtime on the device:
//   * an item = one CTA's work: (level, plane, row chunk, column block).  Level-1 items (forward: the finest level;
//     inverse: the coarsest) need nothing from the launch; they are handed out by a ticket counter in natural order.
//   * every other item reads an approximation produced inside the launch.  Each producing level keeps one completion
//     counter per block of 2^fr_shift output rows and plane; every warp that finishes a chunk bumps the counters of the
//     blocks it covers (acq_rel), the LAST arrival at a block bumps the counter of every dependent chunk (all column
//     blocks of one row chunk of the next level), and the arrival that completes a chunk pushes its items into the
//     ready queue.
//   * a CTA that starts takes a READY dependent item if there is one (its input was written microseconds ago and is
//     still in the 126 MB L2; the latency-bound small levels spread between the bandwidth-bound large items instead of
//     forming a tail of launches), else the next level-1 ticket, else it claims the next queue position and waits for it.
// Nothing ever waits while holding resources that its producers need: an item is only started when its input is
// complete (or when nothing else is left), so the scheme cannot deadlock whatever order the hardware starts CTAs in.
// Counters are cumulative over launches (`epoch`), queue entries carry the epoch, the two ticket counters alternate by
// launch parity (a launch's first ticket zeroes the other one): nothing has to be reset between launches.
constexpr int kMaxLv = 6;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_inc_acqrel(unsigned* p)
{
    unsigned o;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(o) : "l"(p) : "memory");
    return o;
}

struct QCtl {
    const int4* items;          // (level | plane << 4, first row, rows, column block); NULL: one level, item = blockIdx
    const int4* chunks;         // per dependent chunk: (blocks it waits for, first item, items, 0)
    const int* dep_off;         // per (producing level, plane, block): its range in dep_list
    const int* dep_list;        // dependent chunks
    unsigned* ctrl;             // [0],[1] ticket counters (launch parity), [2] queue head, [3] queue tail, counters behind
    unsigned long long* queue;  // ready items: epoch << 32 | item
    unsigned n0, basehp, epoch; // level-1 items; queue positions used by earlier launches; number of this launch (1, 2, ..)
    int cc_off;                 // chunk counters: ctrl[cc_off + chunk]
};
struct QLevelCtl {
    int flag_off;               // block counters of this level: ctrl[flag_off + plane * nb + block]
    int dep_base;               // dep_off index of (this level, plane 0, block 0)
    int fr_shift, nb;           // block = 2^fr_shift output rows, nb blocks per plane
    int arrivals;               // warps that bump a block's counter per launch
    int signals;                // this level's output is read by items of the launch
};

// The item this CTA works on.  Called by one CONVERGED warp; the result is warp-uniform (it comes out of a warp reduction,
// i.e. a uniform register).  Lane 0 does the atomics, but every loop condition is a vote: a data-dependent loop run by a
// single lane makes ptxas treat the rest of the kernel as possibly diverged (BSSY/BSYNC around every branch, WARPSYNC,
// the uniform datapath lost -- the level-1 inverse kernel went from 33 to 50 us that way).
__device__ __forceinline__ int q_wait_entry(const QCtl& q, unsigned pos)
{
    unsigned long long v;
    for (;;) {
        v = ld_acquire_u64(q.queue + pos);   // every lane polls the same word
        if (__all_sync(0xffffffffu, (unsigned)(v >> 32) == q.epoch)) break;
        __nanosleep(200);
    }
    return (int)__reduce_max_sync(0xffffffffu, (unsigned)v);
}
__device__ __forceinline__ int q_take(const QCtl& q)
{
    const bool lane0 = (threadIdx.x & 31) == 0;
    unsigned* c = q.ctrl;
    volatile unsigned* vc = c;
    {   // a dependent item that is ready goes before further level-1 work.  The position is claimed with a plain
        // atomicAdd (a CAS loop collapses under contention: thousands of one-warp CTAs start at once); CTAs racing for the
        // last ready items may claim positions that are filled a little later and wait for them.
        unsigned h = 0, t = 0;
        if (lane0) {
            h = vc[2];
            t = vc[3];
        }
        if (__any_sync(0xffffffffu, (int)(t - h) > 0)) {
            unsigned pos = 0;
            if (lane0) pos = atomicAdd(c + 2, 1u) - q.basehp;
            return q_wait_entry(q, __reduce_max_sync(0xffffffffu, pos));
        }
    }
    unsigned* t0 = c + (q.epoch & 1);
    unsigned tk = 0xffffffffu;
    if (lane0 && *(volatile unsigned*)t0 < q.n0) {
        tk = atomicAdd(t0, 1u);
        if (tk == 0) atomicExch(c + ((q.epoch & 1) ^ 1), 0u);   // for the next launch
    }
    tk = __reduce_min_sync(0xffffffffu, tk);
    if (tk < q.n0) return (int)tk;
    unsigned pos = 0;   // only dependent items are left: claim the next queue position and wait for it
    if (lane0) pos = atomicAdd(c + 2, 1u) - q.basehp;
    return q_wait_entry(q, __reduce_max_sync(0xffffffffu, pos));
}
// one warp: its share of output rows [r0, r0 + nrows) of (level, plane) is stored
__device__ __forceinline__ void q_signal(const QCtl& q, const QLevelCtl& L, int plane, int r0, int nrows)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    const int b0 = r0 >> L.fr_shift, b1 = (r0 + nrows - 1) >> L.fr_shift;
    unsigned* blk = q.ctrl + L.flag_off + plane * L.nb;
    const int* doff = q.dep_off + L.dep_base + plane * L.nb;
    for (int base = b0; base <= b1; base += 32) {
        const int i = base + lane;
        if (i > b1) continue;
        if (atom_inc_acqrel(blk + i) + 1 != q.epoch * (unsigned)L.arrivals) continue;
        // last arrival at this block: tell the chunks that wait for it
        for (int d = doff[i]; d < doff[i + 1]; d++) {
            const int c = q.dep_list[d];
            const int4 ch = q.chunks[c];
            if (atom_inc_acqrel(q.ctrl + q.cc_off + c) + 1 != q.epoch * (unsigned)ch.x) continue;
            const unsigned pos = atomicAdd(q.ctrl + 3, (unsigned)ch.z) - q.basehp;   // the chunk is complete: its items are ready
            for (int k = 0; k < ch.z; k++)
                st_release_u64(q.queue + pos + k, ((unsigned long long)q.epoch << 32) | (unsigned)(ch.y + k));
        }
    }
}

struct alignas(64) FwdLevel {
    CUtensorMap tm;   // (Nc, Nr, batch) fp32 tensor over the source planes, box (WW, SR, 1); valid iff use_tm
    const float* src;
    float *A, *Hb, *V, *D;
    size_t s_src, s_a, s_d;  // plane strides (floats)
    int Nr, Nc, nr, nc;      // input and output plane sizes
    int TH;                  // output rows per chunk (single-level launches; queue items carry their own rows)
    int ncg, nrc;            // column groups (NCW strips each), row chunks
    int use_tm;
    QLevelCtl q;             // completion counters of this level (cross-level launches)
};

template <int HLEN>
struct FwdParams {
    FwdLevel lev[kMaxLv];
    float2 lh[HLEN];  // (L[hlen-1-j], H[hlen-1-j]): row-pass tap pairs in the reference's accumulation order
    float ly[HLEN];   // L[hlen-1-j]
    float hy[HLEN];   // H[hlen-1-j]
    QCtl q;                  // the launch's work queue (q.items == NULL: one level, item = blockIdx)
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

// LOWOCC: variant for grids that put at most 2 CTAs on an SM anyway (one image): it may use more registers.
template <int HLEN, bool LOWOCC>
__global__ void __launch_bounds__(FwdGeom<HLEN>::THREADS, LOWOCC ? 2 : (HLEN <= 14 ? 4 : 3))
    k_fwd2d_stream(const __grid_constant__ FwdParams<HLEN> p)
{
    using G = FwdGeom<HLEN>;
    constexpr int H2 = G::H2, NC = G::NC, NSS = G::NSS, SR = G::SR, NCW = G::NCW;
    extern __shared__ unsigned char smem_raw[];
    __shared__ int4 s_item;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction

    const unsigned ring_s = (smem_u32(smem_raw) + 127u) & ~127u;   // tensor-map destinations are 128-byte aligned
    const unsigned bar_s = ring_s + NCW * NSS * G::SSB;            // full[w][s] at +16*(w*NSS+s), empty right behind it
    TL(0, threadIdx.x == 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NCW * NSS * 2; i++) mbar_init(bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.q.items && warp == 0) {   // one converged warp draws the item
        const int idx = q_take(p.q);
        if (lane == 0) s_item = p.q.items[idx];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();   // the only block-wide barrier: after it the warps only meet through mbarriers
    int level = 0, plane, y0, ny, cg;
    if (p.q.items) {   // through warp reductions: their results live in uniform registers (CREDUX), see the inverse kernel
        const int4 it = s_item;
        const int lp = __reduce_max_sync(0xffffffffu, (unsigned)it.x);
        level = lp & 15;
        plane = lp >> 4;
        y0 = __reduce_max_sync(0xffffffffu, (unsigned)it.y);
        ny = __reduce_max_sync(0xffffffffu, (unsigned)it.z);
        cg = __reduce_max_sync(0xffffffffu, (unsigned)it.w);
    } else {
        const int item = blockIdx.x, ncg0 = p.lev[0].ncg, nrc0 = p.lev[0].nrc;
        cg = item % ncg0;
        const int rest = item / ncg0;
        plane = rest / nrc0;
        y0 = (rest % nrc0) * p.lev[0].TH;
        ny = min(p.lev[0].TH, p.lev[0].nr - y0);
    }
    const FwdLevel& L = p.lev[level];
    const int Nr = L.Nr, Nc = L.Nc, nc_o = L.nc;
    const int npairs = ny + H2 - 1;          // input row pairs this chunk consumes
    const int nss = (2 * npairs + SR - 1) / SR;   // super-slots per consumer
    const int vr0 = 2 * y0 - G::C;           // first input row (virtual: < 0 or >= Nr wraps around)
    const int nstrips = min(NCW, (nc_o - cg * NCW * G::WO + G::WO - 1) / G::WO);   // strips of this CTA inside the image
    if (p.pdl_early) pdl_launch_dependents();
    pdl_wait();        // the previous kernel of the stream (whatever wrote the level-1 source) has completed
    TL(1, threadIdx.x == 0);
#ifdef PDWT_EXPERIMENTS
    if (threadIdx.x == 0 && blockIdx.x < 4096)   // slot 5: what this CTA worked on
        g_timeline[blockIdx.x * 8 + 5] = (unsigned long long)level | ((unsigned long long)ny << 4) | ((unsigned long long)plane << 16) | ((unsigned long long)y0 << 32);
#endif

    if (warp == NCW) {
        // ===================================================================================== producer warp
        const float* src = L.src + (size_t)plane * L.s_src;
        const CUtensorMap* tm = &L.tm;
        const bool use_tm = L.use_tm != 0;
        // level > 0: the source was written by other CTAs of this launch (generic proxy); the item was only handed out
        // once those rows were complete (acquire in q_take, then the block barrier): order the TMA engine's reads behind it
        // (level 0 as well: its source was written by the previous kernel of the stream behind griddepcontrol.wait)
        asm volatile("fence.proxy.async.global;" ::: "memory");
        TL(2, lane == 0);   // dependencies satisfied
        // super-slot k of strip w -> ring slot k % NSS
        auto issue = [&](const int k, const int w) {
            const int slot = k % NSS;
            int row0 = vr0 + SR * k;
            // rows needed from this super-slot all inside the image? (rows past the chunk's last pair may fall outside:
            // the tensor request zero-fills them and nobody reads them)
            const int last_needed = min(row0 + SR, vr0 + 2 * npairs) - 1;
            const bool rows_in = row0 >= 0 && last_needed < Nr;
            row0 += (row0 < 0) ? Nr : 0;
            row0 -= (row0 >= Nr) ? Nr : 0;
            const unsigned full = bar_s + 16 * (w * NSS + slot);
            const int xs = 2 * (cg * NCW + w) * G::WO - G::AL;
            const unsigned dst = ring_s + (w * NSS + slot) * G::SSB;
            if (use_tm && rows_in && xs >= 0 && xs + G::WW <= Nc) {
                if (elect_one()) {
                    mbar_expect_tx(full, G::SSB);
                    tma_load_3d(dst, tm, xs, vr0 + SR * k, plane, full);
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
                    row -= (row >= Nr) ? Nr : 0;
                    int x = xs + 4 * c;
                    x += (x < 0) ? Nc : 0;
                    x -= (x >= Nc) ? Nc : 0;
                    g[it] = src + (size_t)row * Nc + x;
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
            if (!any) __nanosleep(p.poll_ns);
        }
#ifdef PDWT_EXPERIMENTS
        if (lane == 0 && blockIdx.x < 4096) {
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
    const bool col_ok = kcol < nc_o;                 // nc is even, so the pair is in or out as a whole
    const size_t o0 = (size_t)y0 * nc_o + kcol;
    float* oA = L.A + (size_t)plane * L.s_a + o0;
    float* oH = L.Hb + (size_t)plane * L.s_d + o0;
    float* oV = L.V + (size_t)plane * L.s_d + o0;
    float* oD = L.D + (size_t)plane * L.s_d + o0;
    // this lane's window inside a staged row, as a generic pointer (plain loads keep their order w.r.t. the barriers)
    unsigned lane_off = 2 * NC * lane * 4;
    asm volatile("" : "+r"(lane_off));   // opaque: otherwise ptxas re-derives it from SR_TID (S2R, ~25 cycles) in front of every row pair
    const char* lane_ring = static_cast<const char*>(__cvta_shared_to_generic(my_ring)) + lane_off;

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

    // The output pointers run H2-1 rows AHEAD of the stores (a row leaves when its last tap has arrived, H2-1 pairs after
    // its first): they advance every pair and the stores are predicated, so a row pair is straight-line code -- with a
    // branch around the stores ptxas has to agree on ONE register assignment at every merge point and shifts the
    // accumulator file with 14 MOVs per pair.
    oA -= (size_t)(H2 - 1) * nc_o; oH -= (size_t)(H2 - 1) * nc_o; oV -= (size_t)(H2 - 1) * nc_o; oD -= (size_t)(H2 - 1) * nc_o;
    int q = 0;                                    // row pair index within the chunk
    unsigned touch = 0;                           // see the slot release below (short filters only)
    // one row pair whose two rows sit at `rowp` / `rowp + WW floats` of the ring
    auto pair_step = [&](const char* rowp) {
        float xa[G::NV * 4], xb[G::NV * 4];
        load_row(xa, rowp);
        load_row(xb, rowp + G::WW * 4);
        if (q + 1 >= npairs) pdl_launch_dependents();   // last row pair of this warp
        u64 lohi0[NC], lohi1[NC];
        row_pass(xa, lohi0);
        // column pass, w_kern_forward_pass2 (separable.cu:135-176), scatter form: the first row of the pair is tap
        // j = 2*pp of the output row pp pairs back (in place) ...
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
        if (H2 < SR / 2) touch ^= (unsigned)lohi0[0] ^ (unsigned)lohi0[NC - 1] ^ (unsigned)lohi1[0] ^ (unsigned)lohi1[NC - 1];
        TL(4, q == 0 && threadIdx.x == 0);
        // slot H2 received its last tap (j = hlen-1) in this pair
        if (col_ok && q >= H2 - 1) {
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
        oA += nc_o; oH += nc_o; oV += nc_o; oD += nc_o;
        q++;
    };
    unsigned soff = 0, bar = my_bar, parity = 0;  // ring position (byte offset, full barrier, phase) of the super-slot
    const int nfull = npairs / (SR / 2), rem = npairs - nfull * (SR / 2);
    for (int k = 0; k < nfull; k++) {             // whole super-slots: SR/2 row pairs of straight-line code
        mbar_wait(bar, parity);
        TL(3, threadIdx.x == 0 && k == 0);
        const char* ssp = lane_ring + soff;
#pragma unroll
        for (int pin = 0; pin < SR / 2; pin++) pair_step(ssp + pin * (2 * G::WW * 4));
        // Every lane has consumed all rows of this super-slot: hand it back to the producer.  The release must not be
        // ISSUED before the shared-memory loads of the slot have completed (the TMA refill is not ordered against generic
        // loads in flight, see k_inv2d_tma); here every load feeds the row pass and the arrive sits behind that
        // arithmetic in program order, and to keep it there whatever the scheduler does the barrier address is made to
        // depend on accumulators that all SR/2 row pairs of the slot went into ((x * x) & 2 is 0 for every x).
        __syncwarp();
        if (lane == 0) {
            // slot min(SR/2, H2) took in the last min(SR/2, H2) pairs; with short filters (H2 < SR/2) the older pairs of
            // the slot went into `touch` as they were computed
            unsigned x = (unsigned)aLV[SR / 2 < H2 ? SR / 2 : H2][0] ^ (unsigned)aLV[SR / 2 < H2 ? SR / 2 : H2][NC - 1];
            if (H2 < SR / 2) x ^= touch;
            asm volatile(
                "{\n"
                ".reg .u32 t;\n"
                "mul.lo.u32 t, %1, %1;\n"
                "and.b32 t, t, 2;\n"
                "add.u32 t, t, %0;\n"
                "mbarrier.arrive.shared::cta.b64 _, [t];\n"
                "}\n" ::"r"(bar + 8), "r"(x)
                : "memory");
        }
        soff += G::SSB;
        bar += 16;
        if (soff == NSS * G::SSB) {
            soff = 0;
            bar = my_bar;
            parity ^= 1;
        }
    }
    if (rem > 0) {                                // the chunk's last, partial super-slot
        mbar_wait(bar, parity);
        TL(3, threadIdx.x == 0 && nfull == 0);
        const char* ssp = lane_ring + soff;
        for (int pin = 0; pin < rem; pin++) pair_step(ssp + pin * (2 * G::WW * 4));
    }
    TL(6, threadIdx.x == 0);
    // this strip of the chunk is complete: release it to the next level's items
    if (p.q.items && L.q.signals) q_signal(p.q, L.q, plane, y0, ny);
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

// ---- plans of the cross-level launches (host side) ---------------------------------------------------------------------
// One plan = the item, chunk and dependency tables of one (direction, geometry, batch) on one stream, plus its control
// block and ready queue on the device.  Plans live in the filter handle that the transform was called with (one per
// Wavelets object), so two objects never share counters.
struct StreamPlan {
    int dir, dev, hlen, Nr, Nc, batch, nlev, variant;
    cudaStream_t stream;
    int4* d_items = nullptr;   // items, then chunks
    int* d_dep = nullptr;      // dep_off, then dep_list
    unsigned* d_ctrl = nullptr;
    unsigned long long* d_queue = nullptr;
    unsigned nitems = 0, n0 = 0, nhp = 0, nchunks = 0, ndep_off = 0, epoch = 0;
    QLevelCtl lq[kMaxLv];
    int cc_off = 0;
    size_t nctrl = 0;
    bool dirty = false;      // a launch failed: the counters are out of step with `epoch`
    void release()
    {
        if (d_items) cudaFree(d_items);
        if (d_dep) cudaFree(d_dep);
        if (d_ctrl) cudaFree(d_ctrl);
        if (d_queue) cudaFree(d_queue);
    }
};
struct StreamPlans {
    std::mutex mu;
    std::vector<StreamPlan*> v;
};
StreamPlans* stream_plans_create() { return new StreamPlans(); }
void stream_plans_destroy(StreamPlans* ps)
{
    if (!ps) return;
    for (StreamPlan* pl : ps->v) {
        pl->release();
        delete pl;
    }
    delete ps;
}

// one work item as the host sees it; items are generated level by level, plane by plane, row chunk by row chunk, column
// block fastest -- the items of one chunk are consecutive
struct QItem {
    int level, plane, r0, nrows, col;   // what the kernel gets
    int dep0, dep1;                     // (level > 0) virtual block range of level-1, same plane, that it reads (may wrap)
};
struct QLevel {
    int nb, arrivals;                   // counter blocks per plane; warps that complete one block
};

// find the plan or build it: make_items fills the items (level 0 first) and, per level, the block geometry
static int get_plan(StreamPlans* ps, int dir, int hlen, int Nr, int Nc, int batch, int nlev, int variant,
                    const std::function<void(std::vector<QItem>&, QLevel*, int*)>& make_items, cudaStream_t s, StreamPlan** out)
{
    int dev = 0;
    PDWT_CUDA(cudaGetDevice(&dev));
    for (StreamPlan* pl : ps->v) {
        if (pl->dir == dir && pl->dev == dev && pl->stream == s && pl->hlen == hlen && pl->Nr == Nr && pl->Nc == Nc &&
            pl->batch == batch && pl->nlev == nlev && pl->variant == variant) {
            *out = pl;
            return PDWT_OK;
        }
    }
    if (ps->v.size() >= 16) {   // a handle that keeps changing geometry: drop the oldest plan (its stream may still run
        StreamPlan* old = ps->v.front();   // it, so wait for that stream first)
        cudaStreamSynchronize(old->stream);
        old->release();
        delete old;
        ps->v.erase(ps->v.begin());
    }
    StreamPlan* pl = new StreamPlan();
    pl->dir = dir; pl->dev = dev; pl->hlen = hlen; pl->Nr = Nr; pl->Nc = Nc; pl->batch = batch; pl->nlev = nlev;
    pl->variant = variant;
    pl->stream = s;
    std::vector<QItem> items;
    QLevel lv[kMaxLv];
    int fr_shift[kMaxLv];
    make_items(items, lv, fr_shift);
    // control block: [0..3] tickets / queue head / tail, block counters of the producing levels, chunk counters
    size_t off = 4;
    int dep_base = 0;
    for (int l = 0; l < nlev; l++) {
        QLevelCtl& q = pl->lq[l];
        q.fr_shift = fr_shift[l];
        q.nb = lv[l].nb;
        q.arrivals = lv[l].arrivals;
        q.signals = l + 1 < nlev;
        q.flag_off = (int)off;
        q.dep_base = dep_base;
        if (q.signals) {
            off += (size_t)lv[l].nb * batch;
            dep_base += lv[l].nb * batch;
        }
    }
    pl->cc_off = (int)off;
    // chunks of the dependent levels and the blocks they wait for
    std::vector<int4> tab(items.size());
    std::vector<int4> chunks;
    std::vector<std::vector<int>> deps((size_t)dep_base);
    unsigned n0 = 0;
    for (size_t i = 0; i < items.size(); i++) {
        const QItem& it = items[i];
        tab[i] = make_int4(it.level | (it.plane << 4), it.r0, it.nrows, it.col);
        if (it.level == 0) {
            n0++;
            continue;
        }
        const bool same = !chunks.empty() && i > 0 && items[i - 1].level == it.level && items[i - 1].plane == it.plane &&
                          items[i - 1].r0 == it.r0;
        if (same) {
            chunks.back().z++;
            continue;
        }
        const int c = (int)chunks.size(), lp = it.level - 1, nb = lv[lp].nb;
        int need = 0;
        for (int v = it.dep0; v <= it.dep1 && v < it.dep0 + nb; v++) {   // at most every block of the plane once
            int j = v % nb;
            if (j < 0) j += nb;
            deps[(size_t)pl->lq[lp].dep_base + (size_t)it.plane * nb + j].push_back(c);
            need++;
        }
        chunks.push_back(make_int4(need, (int)i, 1, 0));
    }
    pl->n0 = n0;
    pl->nitems = (unsigned)items.size();
    pl->nhp = pl->nitems - n0;
    pl->nchunks = (unsigned)chunks.size();
    off += chunks.size();
    pl->nctrl = off;
    std::vector<int> dep_tab(deps.size() + 1);
    size_t ndl = 0;
    for (size_t i = 0; i < deps.size(); i++) {
        dep_tab[i] = (int)ndl;
        ndl += deps[i].size();
    }
    dep_tab[deps.size()] = (int)ndl;
    pl->ndep_off = (unsigned)dep_tab.size();
    for (auto& d : deps) dep_tab.insert(dep_tab.end(), d.begin(), d.end());
    tab.insert(tab.end(), chunks.begin(), chunks.end());
    cudaError_t e = cudaMalloc(&pl->d_items, sizeof(int4) * tab.size());
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_dep, sizeof(int) * dep_tab.size());
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_ctrl, sizeof(unsigned) * pl->nctrl);
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_queue, sizeof(unsigned long long) * (pl->nhp + 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(pl->d_items, tab.data(), sizeof(int4) * tab.size(), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(pl->d_dep, dep_tab.data(), sizeof(int) * dep_tab.size(), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(pl->d_ctrl, 0, sizeof(unsigned) * pl->nctrl, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(pl->d_queue, 0, sizeof(unsigned long long) * (pl->nhp + 1), s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // the host tables are pageable and about to go out of scope
    if (e != cudaSuccess) {
        pl->release();
        delete pl;
        return note_cuda(e);
    }
    ps->v.push_back(pl);
    *out = pl;
    return PDWT_OK;
}

// fill the queue part of a launch's parameters and advance the plan by one launch
static void plan_bind(StreamPlan* pl, QCtl& q)
{
    q.items = pl->d_items;
    q.chunks = pl->d_items + pl->nitems;
    q.dep_off = pl->d_dep;
    q.dep_list = pl->d_dep + pl->ndep_off;
    q.ctrl = pl->d_ctrl;
    q.queue = pl->d_queue;
    q.n0 = pl->n0;
    q.basehp = pl->epoch * pl->nhp;
    q.epoch = ++pl->epoch;
    q.cc_off = pl->cc_off;
}

// counters out of step after a failed launch: zero them and restart the epochs
static int plan_prepare(StreamPlan* pl, cudaStream_t s)
{
    if (pl->dirty || pl->epoch > 0x3fffffu) {   // (epoch * warps per block must stay far from 2^31)
        PDWT_CUDA(cudaMemsetAsync(pl->d_ctrl, 0, sizeof(unsigned) * pl->nctrl, s));
        PDWT_CUDA(cudaMemsetAsync(pl->d_queue, 0, sizeof(unsigned long long) * (pl->nhp + 1), s));
        pl->epoch = 0;
        pl->dirty = false;
    }
    return PDWT_OK;
}

// PDWT_MULTI=1 (read per call: the tests switch it inside one process) serves all levels of a transform with one launch
static bool multi_level_on()
{
    const char* e = getenv("PDWT_MULTI");
    return e && atoi(e) != 0;
}

// The inverse level kernel with a producer warp and a tensor-map ring (k_inv2d_tma) against the one whose arithmetic warps
// stage their own rows (k_inv2d_stream).  Measured on B200 (tools/time_inv_variants.py, db7, level 1): the TMA variant wins
// 5-7 % once the launch covers the machine several times over with large planes (8 x 4096^2: 203.9 against 214.6 us;
// 64 x 2048^2: 404.6 against 433.3; 8 x 2048^2: 64.2 against 67.7) and loses on one or two 4096^2 planes (37.2 against
// 35.3 us, 63.8 against 60.6) and on small planes whatever the batch (64 x 1024^2: 147.7 against 123.6 -- a larger
// share of edge super-slots, which the producer stages in 8-byte pieces).  PDWT_INV_TMA=0|1 (read per call) overrides.
static bool inv_tma_wanted(long long plane_px, int batch)
{
    if (const char* e = getenv("PDWT_INV_TMA")) return atoi(e) != 0;
    return plane_px >= (1ll << 22) && plane_px * batch >= (1ll << 25);
}

// largest power of two <= 16 that divides a and b
static int pow2_div(int a, int b)
{
    int f = 16;
    while (f > 1 && ((a % f) || (b % f))) f >>= 1;
    return f;
}
static int ilog2i(int v)
{
    int s = 0;
    while ((1 << s) < v) s++;
    return s;
}

template <int HLEN>
static bool fwd_stream_eligible(const StreamLevelIO& io)
{
    using G = FwdGeom<HLEN>;
    const int Nr = io.Nr, Nc = io.Nc;
    // shapes the TMA staging can serve (see the file header)
    if ((Nr & 1) || (Nc & 3) || Nc < G::WW || Nr < HLEN) return false;
    if ((((uintptr_t)io.img.p) & 15) || (io.img.stride & 3)) return false;
    if ((((uintptr_t)io.A.p | (uintptr_t)io.H.p | (uintptr_t)io.V.p | (uintptr_t)io.D.p) & 7) || (io.A.stride & 1) ||
        (io.H.stride & 1))
        return false;
    return true;
}

// Chunk height of a level launched on its own.  A CTA (NCW strips x TH output rows) costs TH + hlen/2 - 1 row pairs per
// consumer warp (the vertical halo is row-pass work only, but the accumulators need the same warm-up), an SM holds up to
// per_sm CTAs and works through the pairs of its resident warps at a roughly constant rate, so the kernel ends when the
// busiest SM does: minimise (CTAs on the busiest SM) x (pairs per CTA).  For one 4096^2 image that picks TH = 56 (296
// CTAs = 2 per SM) instead of a power of two that leaves 40 SMs with half the work of the others.
static int pick_th(int ncg, int nr, int batch, int per_sm, int H2)
{
    int TH = 64;
    const int sms = sm_count();
    double best = 1e30;
    for (int th = 128; th >= 4; th -= 2) {
        const long long ctas = (long long)ncg * idiv_up(nr, th) * batch;
        const long long full = ctas / ((long long)sms * per_sm), rest = ctas % ((long long)sms * per_sm);
        const double units = (double)(full * per_sm + (rest + sms - 1) / sms) * (th + H2 - 1);
        // one CTA per SM = a lone warp per scheduler: measured ~650 cycles per row pair against ~435 per scheduler
        // with two or more warps sharing it
        const double cost = units * (ctas <= sms ? 1.5 : 1.0);
        if (cost < best * 0.999) {
            best = cost;
            TH = th;
        }
    }
    return TH;
}

// Forward levels io[0..nlev): io[l].img (Nr x Nc) -> io[l].A/H/V/D; io[l+1].img must be io[l].A.  Returns the number of
// LEADING levels it has launched (0 = the first level's shape is not covered), < 0 on error.
template <int HLEN>
static int launch_fwd_stream(const Taps& t, StreamPlans* plans, const StreamLevelIO* io, int nlev, int batch, cudaStream_t s)
{
    using G = FwdGeom<HLEN>;
    int n = 0;
    while (n < nlev && n < kMaxLv && fwd_stream_eligible<HLEN>(io[n])) n++;
    if (n == 0) return 0;
    // One launch per level is the default: the cross-level launch (PDWT_MULTI=1) is bit-identical but measured slower on
    // B200: the level kernels are co-limited by FP32 issue and HBM, so overlapping levels frees nothing.
    if (!plans || !multi_level_on()) n = 1;
    static PerDeviceOnce once;
    static int per_sm_dev[64];   // resident CTAs per SM (standard variant), per device
    int dev = 0;
    {
        const cudaError_t eo = once.run([&]() -> cudaError_t {
            int d = 0, per_sm = 0;
            cudaError_t e = cudaGetDevice(&d);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(k_fwd2d_stream<HLEN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(k_fwd2d_stream<HLEN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(G::SMEM > kTwoPerSmBytes ? G::SMEM : kTwoPerSmBytes));
            if (e != cudaSuccess) return e;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fwd2d_stream<HLEN, false>, G::THREADS, G::SMEM) != cudaSuccess || per_sm < 1) {
                cudaGetLastError();
                per_sm = 3;
            }
            per_sm_dev[d & 63] = per_sm;
            return cudaSuccess;
        }, &dev);
        if (eo != cudaSuccess) return note_cuda(eo);
    }
    // Two CTAs (8 consumer warps) per SM with the high-register variant is the default for every grid: measured on B200 it
    // is as fast as or faster than four CTAs of the 96-register variant (8 x 4096^2 level 1: 189 vs 205 us; one image:
    // 33.1 vs 35.3 us) -- fewer, faster warps and fewer concurrent DRAM streams.  PDWT_LOWOCC=0 selects the other one.
    bool lowocc = true;
    if (const char* e = getenv("PDWT_LOWOCC")) lowocc = atoi(e) != 0;
    const int per_sm = lowocc ? 2 : per_sm_dev[dev & 63];
    FwdParams<HLEN> p;
    memset(&p, 0, sizeof p);
    EncodeTiledFn enc = encode_tiled_fn();
    const int env_th = []() { const char* e = getenv("PDWT_TH"); return e ? atoi(e) : 0; }();
    for (int l = 0; l < n; l++) {
        FwdLevel& L = p.lev[l];
        const int Nr = io[l].Nr, Nc = io[l].Nc, nr = Nr / 2, nc = Nc / 2;
        if (enc) {
            const cuuint64_t dims[3] = {(cuuint64_t)Nc, (cuuint64_t)Nr, (cuuint64_t)batch};
            const cuuint64_t strides[2] = {(cuuint64_t)Nc * 4, (cuuint64_t)io[l].img.stride * 4};  // bytes, dims 1 and 2
            const cuuint32_t box[3] = {(cuuint32_t)G::WW, (cuuint32_t)G::SR, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            if (enc(&L.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)io[l].img.p, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                L.use_tm = (getenv("PDWT_NO_TENSORMAP") == nullptr);
        }
        L.src = io[l].img.p; L.A = io[l].A.p; L.Hb = io[l].H.p; L.V = io[l].V.p; L.D = io[l].D.p;
        L.s_src = io[l].img.stride; L.s_a = io[l].A.stride; L.s_d = io[l].H.stride;
        L.Nr = Nr; L.Nc = Nc; L.nr = nr; L.nc = nc;
        L.ncg = idiv_up(nc, G::WO * G::NCW);
        L.TH = env_th > 0 ? env_th : pick_th(L.ncg, nr, batch, per_sm, G::H2);   // a level launched on its own
        L.nrc = idiv_up(nr, L.TH);
    }
    for (int j = 0; j < HLEN; j++) {
        p.lh[j] = make_float2(t.L[HLEN - 1 - j], t.H[HLEN - 1 - j]);
        p.ly[j] = t.L[HLEN - 1 - j];
        p.hy[j] = t.H[HLEN - 1 - j];
    }
    // a consumer needs > 1 us per super-slot and two more are staged behind it: the producer can afford to sleep
    static const unsigned poll_ns = []() { const char* e = getenv("PDWT_POLL_NS"); return e ? (unsigned)atoi(e) : 200u; }();
    p.poll_ns = poll_ns;
    p.pdl_early = pdl_mode() == 1;
    // ... and it asks for so much shared memory that NO SM can take a third CTA: the block scheduler does not spread a
    // grid of 2 x SMs CTAs evenly by itself, and the kernel ends with the busiest SM (PDWT_TWOPERSM=0 switches it off)
    static const bool cap2 = []() { const char* e = getenv("PDWT_TWOPERSM"); return !e || atoi(e) != 0; }();
    long long nctas = 0;
    StreamPlan* pl = nullptr;
    std::unique_lock<std::mutex> plan_lock;
    if (n > 1) {
        // Row chunks of the queue.  One image: each level keeps the height its own cost model picks.  A batch: 128 rows
        // (vertical halo 5 %) while plenty of work is left, tapering over the last two planes so that the launch does
        // not end with a few long items; a level below the first halves the height (its planes have half the rows).
        auto chunk_rows = [&](int l, int plane) -> int {
            if (env_th > 0) return env_th;
            if (batch == 1) return p.lev[l].TH;
            const int d = batch - 1 - plane;
            const int base = d >= 2 ? 128 : (d == 1 ? 64 : 32);
            return std::max(16, base >> l);
        };
        auto make_items = [&](std::vector<QItem>& items, QLevel* lv, int* fr_shift) {
            for (int l = 0; l < n; l++) {
                int fr = 16;
                for (int b = 0; b < batch; b++) fr = std::min(fr, pow2_div(p.lev[l].nr, chunk_rows(l, b)));
                fr_shift[l] = ilog2i(fr);
                lv[l].nb = p.lev[l].nr / fr;
                lv[l].arrivals = idiv_up(p.lev[l].nc, G::WO);   // every strip (consumer warp) of a row
            }
            for (int l = 0; l < n; l++)
                for (int b = 0; b < batch; b++) {
                    const int th = chunk_rows(l, b), nr = p.lev[l].nr, nch = idiv_up(nr, th);
                    for (int k = 0; k < nch; k++) {
                        QItem it;
                        it.level = l; it.plane = b; it.r0 = k * th; it.nrows = std::min(th, nr - it.r0);
                        it.dep0 = it.dep1 = 0;
                        if (l > 0) {   // input rows 2 y0 - C .. of the level above, virtual (they may wrap)
                            const int v0 = 2 * it.r0 - G::C, v1 = v0 + 2 * (it.nrows + G::H2 - 1) - 1;
                            it.dep0 = v0 >> fr_shift[l - 1];   // arithmetic shift: floor
                            it.dep1 = v1 >> fr_shift[l - 1];
                        }
                        for (int cg = 0; cg < p.lev[l].ncg; cg++) {
                            it.col = cg;
                            items.push_back(it);
                        }
                    }
                }
        };
        plan_lock = std::unique_lock<std::mutex>(plans->mu);
        int rc = get_plan(plans, 0, HLEN, io[0].Nr, io[0].Nc, batch, n, env_th, make_items, s, &pl);
        if (rc < 0) return rc;
        rc = plan_prepare(pl, s);
        if (rc < 0) return rc;
        for (int l = 0; l < n; l++) p.lev[l].q = pl->lq[l];
        plan_bind(pl, p.q);
        nctas = pl->nitems;
    } else {
        nctas = (long long)p.lev[0].ncg * p.lev[0].nrc * batch;
    }
    if (nctas > 0x7fffffff) return 0;
    PDWT_PROF(prof_tag(n > 1 ? "k_fwd2d_stream_levels" : "k_fwd2d_stream", io[0].Nr, io[0].Nc), s);
    cudaError_t e;
    if (lowocc)
        e = launch_pdl(k_fwd2d_stream<HLEN, true>, dim3((unsigned)nctas), G::THREADS,
                       cap2 && G::SMEM < kTwoPerSmBytes ? kTwoPerSmBytes : G::SMEM, s, p);
    else
        e = launch_pdl(k_fwd2d_stream<HLEN, false>, dim3((unsigned)nctas), G::THREADS, G::SMEM, s, p);
    if (e == cudaSuccess) {
        count_launch();
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        if (pl) pl->dirty = true;
        return note_cuda(e);
    }
    return n;
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

struct InvLevel {
    const float *A, *Hb, *V, *D;
    float* dst;
    size_t s_a, s_d, s_dst;                  // plane strides (floats)
    int nr, nc, Mr, Mc;                      // coefficient and output plane sizes (Mr = 2 nr, Mc = 2 nc)
    int TM;                                  // output row PAIRS per chunk (single-level launches)
    int ncb, nrc;
    QLevelCtl q;                             // completion counters of this level (cross-level launches): rows of dst
};

template <int HLEN>
struct InvParams {
    InvLevel lev[kMaxLv];                    // lev[0] is the coarsest level of the launch; lev[k].A == lev[k-1].dst
    float il[2][HLEN / 2], ih[2][HLEN / 2];  // [output parity][j]: IL / IH taps in accumulation order
    float2 lh[2][HLEN / 2];                  // the same as (IL, IH) pairs for the row synthesis
    QCtl q;                                  // the launch's work queue (q.items == NULL: one level, item = blockIdx)
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
__global__ void __launch_bounds__(32, HLEN <= 14 ? 16 : 12) k_inv2d_stream(const __grid_constant__ InvParams<HLEN> p)
{
    using G = InvGeom<HLEN>;
    constexpr int H2 = G::H2, NSLOT = G::NSLOT, WIN = G::WIN, SHIFT = G::SHIFT, UNR = G::UNR, DEPTH = G::DEPTH;
    extern __shared__ __align__(16) unsigned char smem_inv[];
    const int lane = threadIdx.x;
    int level = 0, plane, m0, nm, cb;
    if (p.q.items) {
        // lane 0 draws the ticket; the item reaches the warp through warp reductions, whose results live in UNIFORM
        // registers (CREDUX): everything derived from it stays on the uniform datapath, as with blockIdx
        const int idx = q_take(p.q);
        int4 it = make_int4(0, 0, 0, 0);
        if (lane == 0) it = p.q.items[idx];
        const int lp = __reduce_max_sync(0xffffffffu, (unsigned)it.x);
        level = lp & 15;
        plane = lp >> 4;
        m0 = __reduce_max_sync(0xffffffffu, (unsigned)it.y);
        nm = __reduce_max_sync(0xffffffffu, (unsigned)it.z);
        cb = __reduce_max_sync(0xffffffffu, (unsigned)it.w);
    } else {
        const int item = blockIdx.x, ncb0 = p.lev[0].ncb, nrc0 = p.lev[0].nrc;
        cb = item % ncb0;
        const int rest = item / ncb0;
        plane = rest / nrc0;
        m0 = (rest % nrc0) * p.lev[0].TM;
        nm = min(p.lev[0].TM, p.lev[0].nr - m0);
    }
    const InvLevel& L = p.lev[level];
    const int nr = L.nr, nc = L.nc, Mc = L.Mc;
    const int k0 = cb * G::WOUT;

    // ---- column synthesis side: this lane owns coefficient columns (col, col+1) of the strip, wrapped periodically
    int col = k0 - G::ALC + 2 * lane;
    col += (col < 0) ? nc : 0;
    col -= (col >= nc) ? nc : 0;
    int lrow = m0 - G::CC;                    // next coefficient row to load (wrapped: separable.cu:265-273)
    lrow += (lrow < 0) ? nr : 0;
    int to_wrap = nr - lrow;
    // ONE running per-lane pointer (into A) and three uniform byte distances to the same element of H, V, D: per row
    // that is a 64-bit add for each address and one for the advance, nothing else
    const char* pa = reinterpret_cast<const char*>(L.A + (size_t)plane * L.s_a + (size_t)lrow * nc + col);
    const ptrdiff_t plane_d = (ptrdiff_t)((size_t)plane * L.s_d) - (ptrdiff_t)((size_t)plane * L.s_a);
    const ptrdiff_t dH = (reinterpret_cast<const char*>(L.Hb) - reinterpret_cast<const char*>(L.A)) + 4 * plane_d;
    const ptrdiff_t dV = (reinterpret_cast<const char*>(L.V) - reinterpret_cast<const char*>(L.A)) + 4 * plane_d;
    const ptrdiff_t dD = (reinterpret_cast<const char*>(L.D) - reinterpret_cast<const char*>(L.A)) + 4 * plane_d;
    ptrdiff_t row_b = 4 * (ptrdiff_t)nc, wrap_b = 4 * (ptrdiff_t)nc * nr;
    int nr_w = nr;
    // opaque to ptxas: otherwise it re-derives these from the (dynamically indexed) level parameters in front of every
    // coefficient row -- an LDC, two 64-bit multiplies and three selects per row pair
    asm volatile("" : "+l"(row_b), "+l"(wrap_b), "+r"(nr_w));

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
                to_wrap = nr_w;
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
    pdl_wait();        // the previous kernel of the stream (whatever wrote the coefficients) has completed
    TL(1, lane == 0);
    // level > 0: A was written by other CTAs of this launch and the item was handed out once those rows were complete
    // (acquire by lane 0 in q_take); every lane's own L1-allocating loads come after its own fence
    if (level > 0) __threadfence();
#pragma unroll
    for (int i = 0; i < DEPTH; i++) issue_row(i);
#pragma unroll
    for (int i = 0; i < WIN; i++) load_row(i);

    // ---- row synthesis side: lanes 0..LPR-1 take the even output row, lanes 16..16+LPR-1 the odd one; each turns 4
    // coefficient columns into 8 pixels
    const int g = lane >> 4, lq = lane & 15;
    const int px0 = 2 * k0 + 8 * lq;
    const bool row_lane = lq < G::LPR;
    float* out = L.dst + (size_t)plane * L.s_dst + (size_t)(2 * m0 + g) * Mc + px0;
    const bool st0 = row_lane && px0 + 4 <= Mc, st1 = row_lane && px0 + 8 <= Mc;
    // tile addressing (see InvGeom): the writer stores chunk `lane`, the reader loads chunks 2*lq + v
    unsigned tile_wr_off = 16 * ((lane >> 1) + (lane & 1) * G::ODD0), tile_rd_off = g * G::TROWB + 16 * lq;
    asm volatile("" : "+r"(tile_wr_off), "+r"(tile_rd_off));   // opaque: no re-derivation from the lane index per row pair
    unsigned char* const tile_wr = smem_inv + tile_wr_off;
    const unsigned char* const tile_rd = smem_inv + tile_rd_off;

    int s = 0;
    bool more = true;
    while (more) {
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
            if (s >= nm) {
                more = false;
                break;
            }
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
                // the two branch sums are added last (separable.cu:283-288).  Four scalar adds that write the tile vector's
                // registers in place: a packed add would leave (t1a,t1b),(t2a,t2b) and need moves to interleave them
                float a0, a1, h0, h1, v0, v1, d0, d1;
                unpack2(sa, a0, a1);
                unpack2(sh, h0, h1);
                unpack2(sv, v0, v1);
                unpack2(sd, d0, d1);
                *reinterpret_cast<float4*>(tile_wr + (buf * 2 + par) * G::TROWB) =
                    make_float4(__fadd_rn(a0, h0), __fadd_rn(v0, d0), __fadd_rn(a1, h1), __fadd_rn(v1, d1));
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
            out += 2 * (size_t)Mc;
            s++;
        }
    }
    // this block of the chunk is complete: release it to the next level's items
    if (p.q.items && L.q.signals) q_signal(p.q, L.q, plane, 2 * m0, 2 * nm);
}

// ---- the inverse level kernel with a TMA-fed ring (one level per launch) ---------------------------------------------------
// Same arithmetic and the same consumer code as k_inv2d_stream, but the coefficient rows reach shared memory the way the
// forward's input rows do: CTA = NCW consumer warps + ONE producer warp; a consumer owns a strip of 64 coefficient columns
// and a private ring of NSS super-slots x SR rows x 4 planes (A, H, V, D); interior super-slots are FOUR tensor-map
// requests (box 64 x SR per plane), edge super-slots (periodic wrap in rows or columns, separable.cu:265-273) are staged by
// the producer's lanes with 8-byte LDGSTS pieces that arrive on the same "full" mbarrier.  The arithmetic warps issue no
// global loads at all: 2368 one-warp CTAs reading four planes in 256-byte pieces (9 500 concurrent DRAM streams) become
// 8-row boxes, and the per-lane LDGSTS + cp.async group accounting leaves the consumers' instruction stream.
template <int HLEN>
struct InvTmaGeom {
    using G = InvGeom<HLEN>;
    static constexpr int NCW = 4;                        // consumer warps per CTA
    static constexpr int SR = (G::UNR % 4 == 0) ? 4 : 2; // coefficient rows per super-slot (static row index in the body)
    static constexpr int NSS = 16 / SR;                  // super-slots per consumer ring (16 rows = 16 KB per consumer)
    static constexpr int PLB = SR * 64 * 4;              // bytes of one plane's box
    static constexpr int SSB = 4 * PLB;                  // bytes per super-slot (A, H, V, D)
    static constexpr int THREADS = (NCW + 1) * 32;
    static constexpr size_t SMEM = (size_t)NCW * NSS * SSB + (size_t)NCW * G::TILEB + NCW * NSS * 2 * sizeof(u64) + 128;
    static_assert(G::UNR % SR == 0, "the row within a super-slot must be static in the unrolled body");
};

template <int HLEN>
struct InvTmaParams {
    CUtensorMap tm[4];                       // A, H, V, D as (nc, nr, batch) tensors, box (64, SR, 1)
    float il[2][HLEN / 2], ih[2][HLEN / 2];  // [output parity][j]: IL / IH taps in accumulation order
    float2 lh[2][HLEN / 2];                  // the same as (IL, IH) pairs for the row synthesis
    const float* src[4];                     // A, H, V, D
    size_t s_src[4];                         // plane strides (floats)
    float* dst;
    size_t s_dst;
    int nr, nc, Mr, Mc;                      // coefficient and output plane sizes (Mr = 2 nr, Mc = 2 nc)
    int TM;                                  // output row PAIRS per chunk
    int ncb, ncg, nrc;                       // column blocks (strips), groups of NCW strips, row chunks
    int pdl_early;
    unsigned poll_ns;
};

template <int HLEN>
__global__ void __launch_bounds__(InvTmaGeom<HLEN>::THREADS, HLEN <= 14 ? 3 : 2)
    k_inv2d_tma(const __grid_constant__ InvTmaParams<HLEN> p)
{
    using T = InvTmaGeom<HLEN>;
    using G = InvGeom<HLEN>;
    constexpr int H2 = G::H2, NSLOT = G::NSLOT, WIN = G::WIN, SHIFT = G::SHIFT, UNR = G::UNR;
    constexpr int NCW = T::NCW, SR = T::SR, NSS = T::NSS;
    extern __shared__ unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
    const int item = blockIdx.x;
    const unsigned ring_s = (smem_u32(smem_raw) + 127u) & ~127u;      // tensor-map destinations are 128-byte aligned
    unsigned char* const ring_g = smem_raw + (ring_s - smem_u32(smem_raw));
    const unsigned tiles_off = NCW * NSS * T::SSB;
    const unsigned bar_s = ring_s + tiles_off + NCW * (unsigned)G::TILEB;   // full[w][s] at +16*(w*NSS+s), empty behind it

    const int cg = item % p.ncg, rest = item / p.ncg, rc = rest % p.nrc, plane = rest / p.nrc;
    const int m0 = rc * p.TM;
    const int nm = min(p.TM, p.nr - m0);
    const int nrows = nm + WIN - 1;                 // coefficient rows this chunk consumes
    const int nss = (nrows + SR - 1) / SR;
    const int vr0 = m0 - G::CC;                     // first coefficient row (virtual: may be < 0 or run past nr)
    const int nstrips = min(NCW, p.ncb - cg * NCW);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NCW * NSS * 2; i++) mbar_init(bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (p.pdl_early) pdl_launch_dependents();
    pdl_wait();

    if (warp == NCW) {
        // ===================================================================================== producer warp
        // the coefficients were written through the generic proxy (by the previous kernel of the stream): order the TMA
        // engine's reads (async proxy) behind griddepcontrol.wait explicitly
        asm volatile("fence.proxy.async.global;" ::: "memory");
        auto issue = [&](const int k, const int w) {
            const int slot = k % NSS;
            const int row0v = vr0 + SR * k;
            const int last_needed = min(row0v + SR, vr0 + nrows) - 1;
            const bool rows_in = row0v >= 0 && last_needed < p.nr;
            const unsigned full = bar_s + 16 * (w * NSS + slot);
            const int xs = (cg * NCW + w) * G::WOUT - G::ALC;     // first coefficient column of the strip
            const unsigned dst = ring_s + (w * NSS + slot) * T::SSB;
            if (rows_in && xs >= 0 && xs + 64 <= p.nc) {
                if (elect_one()) {
                    mbar_expect_tx(full, T::SSB);
#pragma unroll
                    for (int a = 0; a < 4; a++) tma_load_3d(dst + a * T::PLB, &p.tm[a], xs, row0v, plane, full);
                }
            } else {
                // periodic wrap in rows and / or columns: 8-byte pieces (xs and nc are even), wrapped index per piece
                constexpr int PPR = 32;                    // pieces per row
                constexpr int NIT = (4 * SR * PPR) / 32;   // pieces per lane
#pragma unroll
                for (int it = 0; it < NIT; it++) {
                    const int i = lane + 32 * it;
                    const int a = i / (SR * PPR), r = (i / PPR) % SR, c = i % PPR;
                    int row = row0v + r;
                    row += (row < 0) ? p.nr : 0;
                    row -= (row >= p.nr) ? p.nr : 0;
                    row = row < 0 ? 0 : (row >= p.nr ? p.nr - 1 : row);   // rows past the chunk's last are never used
                    int x = xs + 2 * c;
                    x += (x < 0) ? p.nc : 0;
                    x -= (x >= p.nc) ? p.nc : 0;
                    const float* g = p.src[a] + (size_t)plane * p.s_src[a] + (size_t)row * p.nc + x;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + a * T::PLB + r * 256 + c * 8), "l"(g) : "memory");
                }
                cp_async_mbar_arrive(full);   // +1 pending now, -1 when this lane's copies have landed
                __syncwarp();
                if (lane == 0) mbar_arrive(full);
            }
            __syncwarp();
        };
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
                    if (k >= NSS) ok = mbar_poll(bar_s + 16 * (w * NSS + k % NSS) + 8, ((k / NSS) + 1) & 1);
                    if (__shfl_sync(0xffffffffu, ok, 0)) {
                        issue(k, w);
                        next[w] = k + 1;
                        remaining--;
                        any = true;
                    }
                }
            }
            if (!any) __nanosleep(p.poll_ns);
        }
        return;
    }

    // ========================================================================================= consumer warps
    if (warp >= nstrips) return;
    const int cb = cg * NCW + warp;
    const int k0 = cb * G::WOUT;
    const unsigned my_bar = bar_s + 16 * warp * NSS;
    unsigned char* const my_tile = ring_g + tiles_off + warp * G::TILEB;
    // this lane's two columns inside a staged row of plane a: ring + slot * SSB + a * PLB + row * 256 + 8 * lane
    unsigned lane_off = 8 * lane;
    asm volatile("" : "+r"(lane_off));
    const unsigned char* const my_ring = ring_g + warp * NSS * T::SSB + lane_off;

    u64 wA[NSLOT], wH[NSLOT], wV[NSLOT], wD[NSLOT];  // register window, slot = (row index within the chunk) % NSLOT
#pragma unroll
    for (int i = 0; i < NSLOT; i++) wA[i] = wH[i] = wV[i] = wD[i] = 0ull;
    unsigned soff = 0, bar = my_bar, parity = 0;     // ring position of the super-slot that holds the NEXT row to load
    // row r of the chunk (r % SR == rin, static): ring -> window slot r % NSLOT
    // A super-slot may only go back to the producer once its rows ARE in registers.  An arrive issued right behind the
    // shared-memory loads does not wait for them (nothing has consumed their results yet), and the TMA refill -- async
    // proxy, unordered against this warp's generic loads still in flight -- can overtake them: with three CTAs per SM
    // about one super-slot in 10^4 then delivered rows of the NEXT refill (the forward kernel releases a slot after the
    // arithmetic that consumed it, so it never saw this).  Here the arrive is (1) deferred until the first row of the
    // next super-slot is loaded, one step later, and (2) made to DEPEND on every load of the slot it releases: one
    // register of each goes into an XOR chain x and the barrier address is offset by (x * x) & 2 -- zero for every
    // x, but not for ptxas (in PTX: LLVM knows that bit 1 of a square is clear and would drop the chain).
    unsigned pbar = 0;   // "empty" barrier of the super-slot read last
    auto release_prev = [&](const int r) {   // rows r - SR .. r - 1 (static window slots) made up that super-slot
        unsigned x = 0;
#pragma unroll
        for (int q = 1; q <= SR; q++) {
            const int sl = (r - q) % NSLOT;
            x ^= (unsigned)wA[sl] ^ (unsigned)wH[sl];
            x ^= (unsigned)wV[sl] ^ (unsigned)wD[sl];
        }
        __syncwarp();
        if (lane == 0)
            asm volatile(
                "{\n"
                ".reg .u32 t;\n"
                "mul.lo.u32 t, %1, %1;\n"
                "and.b32 t, t, 2;\n"
                "add.u32 t, t, %0;\n"
                "mbarrier.arrive.shared::cta.b64 _, [t];\n"
                "}\n" ::"r"(pbar), "r"(x)
                : "memory");
    };
    auto load_row = [&](const int r, const int rin) {
        if (rin == 0) {
            if (r >= SR) release_prev(r);
            mbar_wait(bar, parity);
        }
        const float* c = reinterpret_cast<const float*>(my_ring + soff + rin * 256);
        const float2 a = *reinterpret_cast<const float2*>(c);
        const float2 h = *reinterpret_cast<const float2*>(c + T::PLB / 4);
        const float2 v = *reinterpret_cast<const float2*>(c + 2 * (T::PLB / 4));
        const float2 d = *reinterpret_cast<const float2*>(c + 3 * (T::PLB / 4));
        wA[r % NSLOT] = pack2(a.x, a.y);
        wH[r % NSLOT] = pack2(h.x, h.y);
        wV[r % NSLOT] = pack2(v.x, v.y);
        wD[r % NSLOT] = pack2(d.x, d.y);
        if (rin == SR - 1) {   // on to the next super-slot; the last ones of a chunk are never refilled, nobody waits for them
            pbar = bar + 8;
            soff += T::SSB;
            bar += 16;
            if (soff == NSS * T::SSB) {
                soff = 0;
                bar = my_bar;
                parity ^= 1;
            }
        }
    };
#pragma unroll
    for (int i = 0; i < WIN; i++) load_row(i, i % SR);

    // ---- row synthesis side: lanes 0..LPR-1 take the even output row, lanes 16..16+LPR-1 the odd one
    const int g = lane >> 4, lq = lane & 15;
    const int px0 = 2 * k0 + 8 * lq;
    const bool row_lane = lq < G::LPR;
    float* out = p.dst + (size_t)plane * p.s_dst + (size_t)(2 * m0 + g) * p.Mc + px0;
    const bool st0 = row_lane && px0 + 4 <= p.Mc, st1 = row_lane && px0 + 8 <= p.Mc;
    unsigned tile_wr_off = 16 * ((lane >> 1) + (lane & 1) * G::ODD0), tile_rd_off = g * G::TROWB + 16 * lq;
    asm volatile("" : "+r"(tile_wr_off), "+r"(tile_rd_off));
    unsigned char* const tile_wr = my_tile + tile_wr_off;
    const unsigned char* const tile_rd = my_tile + tile_rd_off;
    ptrdiff_t out_step = 2 * (ptrdiff_t)p.Mc;
    asm volatile("" : "+l"(out_step));

    int s = 0;
    bool more = true;
    while (more) {
#pragma unroll
        for (int u = 0; u < UNR; u++) {       // body: UNR output row pairs; register, ring-row and tile indices all static
            if (s >= nm) {
                more = false;
                break;
            }
            if (s + 1 >= nm) pdl_launch_dependents();
            if (s + 1 < nm) load_row(u + WIN, (u + WIN) % SR);   // the row that the NEXT pair adds to the window
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
                float a0, a1, h0, h1, v0, v1, d0, d1;
                unpack2(sa, a0, a1);
                unpack2(sh, h0, h1);
                unpack2(sv, v0, v1);
                unpack2(sd, d0, d1);
                *reinterpret_cast<float4*>(tile_wr + (buf * 2 + par) * G::TROWB) =
                    make_float4(__fadd_rn(a0, h0), __fadd_rn(v0, d0), __fadd_rn(a1, h1), __fadd_rn(v1, d1));
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
            out += out_step;
            s++;
        }
    }
}

template <int HLEN>
static bool inv_stream_eligible(const StreamLevelIO& io)
{
    using G = InvGeom<HLEN>;
    const int Mr = io.Nr, Mc = io.Nc, nr = Mr / 2, nc = Mc / 2;
    if ((Mr & 1) || (Mc & 1) || (nc & 1) || nc < 64 || nr < G::WIN) return false;
    if ((((uintptr_t)io.A.p | (uintptr_t)io.H.p | (uintptr_t)io.V.p | (uintptr_t)io.D.p) & 7) || (io.A.stride & 1) ||
        (io.H.stride & 1))
        return false;
    if ((((uintptr_t)io.img.p) & 15) || (io.img.stride & 3)) return false;
    return true;
}

// Chunk height (output row PAIRS per one-warp CTA) of a level launched on its own.  An item costs TM loop iterations plus a
// prologue worth ~2.5 of them (WIN-1 extra coefficient rows and the pipeline fill); items are spread over sms x per_sm
// resident warps, so the kernel takes about ceil(items / slots) x (TM + 2.5): pick the TM that minimises it (4096^2:
// TM = 32, exactly one item per slot).
static int pick_tm(int ncb, int nr, int batch, int per_sm)
{
    int TM = 32;
    const long long slots = (long long)sm_count() * per_sm;
    double best = 1e30;
    for (int tm = 64; tm >= 2; tm--) {
        const long long items = (long long)ncb * idiv_up(nr, tm) * batch;
        const double cost = (double)((items + slots - 1) / slots) * (tm + 2.5);
        if (cost < best * 0.999) {
            best = cost;
            TM = tm;
        }
    }
    return TM;
}

// one level through the TMA-fed kernel; 1 = launched, 0 = shape not covered (16-byte-aligned rows needed), < 0 = error
template <int HLEN>
static int launch_inv_tma(const Taps& t, const StreamLevelIO& io, int batch, cudaStream_t s)
{
    using G = InvGeom<HLEN>;
    using T = InvTmaGeom<HLEN>;
    const int Mr = io.Nr, Mc = io.Nc, nr = Mr / 2, nc = Mc / 2;
    if ((G::ALC & 3) != 0) return 0;   // a tensor-map box must start on a 16-byte boundary: hlen 12, 14, 16, 18
    if (!inv_stream_eligible<HLEN>(io) || (nc & 3) || nr < T::SR) return 0;
    const Plane2 src[4] = {io.A, io.H, io.V, io.D};
    for (int a = 0; a < 4; a++)
        if ((((uintptr_t)src[a].p) & 15) || (src[a].stride & 3)) return 0;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return 0;
    static PerDeviceOnce once;
    static int per_sm_dev[64];
    int dev = 0;
    {
        const cudaError_t eo = once.run([&]() -> cudaError_t {
            int d = 0, per_sm = 0;
            cudaError_t e = cudaGetDevice(&d);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(k_inv2d_tma<HLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
            if (e != cudaSuccess) return e;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_inv2d_tma<HLEN>, T::THREADS, T::SMEM) != cudaSuccess || per_sm < 1) {
                cudaGetLastError();
                per_sm = 2;
            }
            per_sm_dev[d & 63] = per_sm;
            return cudaSuccess;
        }, &dev);
        if (eo != cudaSuccess) return note_cuda(eo);
    }
    const int per_sm = per_sm_dev[dev & 63];
    InvTmaParams<HLEN> p;
    memset(&p, 0, sizeof p);
    for (int a = 0; a < 4; a++) {
        const cuuint64_t dims[3] = {(cuuint64_t)nc, (cuuint64_t)nr, (cuuint64_t)batch};
        const cuuint64_t strides[2] = {(cuuint64_t)nc * 4, (cuuint64_t)src[a].stride * 4};
        const cuuint32_t box[3] = {64, (cuuint32_t)T::SR, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&p.tm[a], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)src[a].p, dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return 0;
        p.src[a] = src[a].p;
        p.s_src[a] = src[a].stride;
    }
    for (int par = 0; par < 2; par++) {
        const int off = par ? G::SHIFT : 1 - G::SHIFT;
        for (int j = 0; j < HLEN / 2; j++) {
            p.il[par][j] = t.IL[HLEN - 1 - (2 * j + off)];
            p.ih[par][j] = t.IH[HLEN - 1 - (2 * j + off)];
            p.lh[par][j] = make_float2(p.il[par][j], p.ih[par][j]);
        }
    }
    p.dst = io.img.p;
    p.s_dst = io.img.stride;
    p.nr = nr; p.nc = nc; p.Mr = Mr; p.Mc = Mc;
    p.ncb = idiv_up(nc, G::WOUT);
    p.ncg = idiv_up(p.ncb, T::NCW);
    // chunk height: CTAs (NCW strips x TM row pairs, cost TM + 2.5) over sms x per_sm slots, as pick_tm
    int TM = 32;
    {
        const long long slots = (long long)sm_count() * per_sm;
        double best = 1e30;
        for (int tm = 128; tm >= 4; tm -= 2) {
            const long long items = (long long)p.ncg * idiv_up(nr, tm) * batch;
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
    static const unsigned poll_ns = []() { const char* e = getenv("PDWT_POLL_NS"); return e ? (unsigned)atoi(e) : 200u; }();
    p.poll_ns = poll_ns;
    const long long nctas = (long long)p.ncg * p.nrc * batch;
    if (nctas > 0x7fffffff) return 0;
    PDWT_PROF(prof_tag("k_inv2d_tma", Mr, Mc), s);
    PDWT_CUDA(launch_pdl(k_inv2d_tma<HLEN>, dim3((unsigned)nctas), T::THREADS, T::SMEM, s, p));
    PDWT_LAUNCH_CHECK();
    return 1;
}

// Inverse levels io[0..nlev), coarsest first: io[l].A/H/V/D (Nr/2 x Nc/2) -> io[l].img (Nr x Nc); io[l+1].A must be
// io[l].img.  Returns the number of LEADING levels launched (0 = the first level's shape is not covered), < 0 on error.
template <int HLEN>
static int launch_inv_stream(const Taps& t, StreamPlans* plans, const StreamLevelIO* io, int nlev, int batch, cudaStream_t s)
{
    using G = InvGeom<HLEN>;
    int n = 0;
    while (n < nlev && n < kMaxLv && inv_stream_eligible<HLEN>(io[n])) n++;
    if (n == 0) return 0;
    if (!plans || !multi_level_on()) n = 1;   // see launch_fwd_stream
    if (n == 1 && inv_tma_wanted((long long)io[0].Nr * io[0].Nc, batch)) {   // the TMA-fed variant of the same level kernel
        const int done = launch_inv_tma<HLEN>(t, io[0], batch, s);
        if (done != 0) return done;
    }
    static PerDeviceOnce once;
    static int per_sm_dev[64];
    int dev = 0;
    {
        const cudaError_t eo = once.run([&]() -> cudaError_t {
            int d = 0, per_sm = 0;
            cudaError_t e = cudaGetDevice(&d);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(k_inv2d_stream<HLEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return e;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_inv2d_stream<HLEN>, 32, G::SMEM) != cudaSuccess || per_sm < 1) {
                cudaGetLastError();
                per_sm = 12;
            }
            per_sm_dev[d & 63] = per_sm;
            return cudaSuccess;
        }, &dev);
        if (eo != cudaSuccess) return note_cuda(eo);
    }
    const int per_sm = per_sm_dev[dev & 63];
    InvParams<HLEN> p;
    memset(&p, 0, sizeof p);
    for (int par = 0; par < 2; par++) {
        const int off = par ? G::SHIFT : 1 - G::SHIFT;
        for (int j = 0; j < HLEN / 2; j++) {
            p.il[par][j] = t.IL[HLEN - 1 - (2 * j + off)];
            p.ih[par][j] = t.IH[HLEN - 1 - (2 * j + off)];
            p.lh[par][j] = make_float2(p.il[par][j], p.ih[par][j]);
        }
    }
    const int env_tm = []() { const char* e = getenv("PDWT_TM"); return e ? atoi(e) : 0; }();
    for (int l = 0; l < n; l++) {
        InvLevel& L = p.lev[l];
        const int Mr = io[l].Nr, Mc = io[l].Nc, nr = Mr / 2, nc = Mc / 2;
        L.A = io[l].A.p; L.Hb = io[l].H.p; L.V = io[l].V.p; L.D = io[l].D.p; L.dst = io[l].img.p;
        L.s_a = io[l].A.stride; L.s_d = io[l].H.stride; L.s_dst = io[l].img.stride;
        L.nr = nr; L.nc = nc; L.Mr = Mr; L.Mc = Mc;
        L.ncb = idiv_up(nc, G::WOUT);
        L.TM = env_tm > 0 ? env_tm : pick_tm(L.ncb, nr, batch, per_sm);   // a level launched on its own
        L.nrc = idiv_up(nr, L.TM);
    }
    p.pdl_early = pdl_mode() == 1;
    long long nitems = 0;
    StreamPlan* pl = nullptr;
    std::unique_lock<std::mutex> plan_lock;
    if (n > 1) {
        // Row chunks (output row PAIRS) of the queue: one image keeps each level's own cost model; a batch uses 32 pairs
        // and halves them on the last plane, whose finest level is what the launch ends with.
        auto chunk_rows = [&](int l, int plane) -> int {
            if (env_tm > 0) return env_tm;
            if (batch == 1) return p.lev[l].TM;
            return (batch - 1 - plane >= 1) ? 32 : 16;
        };
        auto make_items = [&](std::vector<QItem>& items, QLevel* lv, int* fr_shift) {
            for (int l = 0; l < n; l++) {
                int fr = 16;
                for (int b = 0; b < batch; b++) fr = std::min(fr, pow2_div(p.lev[l].Mr, 2 * chunk_rows(l, b)));
                fr_shift[l] = ilog2i(fr);
                lv[l].nb = p.lev[l].Mr / fr;
                lv[l].arrivals = p.lev[l].ncb;   // every column block (one warp) of a row
            }
            for (int l = 0; l < n; l++)
                for (int b = 0; b < batch; b++) {
                    const int tm = chunk_rows(l, b), nr = p.lev[l].nr, nch = idiv_up(nr, tm);
                    for (int k = 0; k < nch; k++) {
                        QItem it;
                        it.level = l; it.plane = b; it.r0 = k * tm; it.nrows = std::min(tm, nr - it.r0);
                        it.dep0 = it.dep1 = 0;
                        if (l > 0) {   // coefficient rows m0 - CC .. = rows of the coarser level's output, virtual
                            const int v0 = it.r0 - G::CC, v1 = v0 + it.nrows + G::WIN - 1 - 1;
                            it.dep0 = v0 >> fr_shift[l - 1];   // arithmetic shift: floor
                            it.dep1 = v1 >> fr_shift[l - 1];
                        }
                        for (int cb = 0; cb < p.lev[l].ncb; cb++) {
                            it.col = cb;
                            items.push_back(it);
                        }
                    }
                }
        };
        plan_lock = std::unique_lock<std::mutex>(plans->mu);
        int rc = get_plan(plans, 1, HLEN, io[n - 1].Nr, io[n - 1].Nc, batch, n, env_tm, make_items, s, &pl);
        if (rc < 0) return rc;
        rc = plan_prepare(pl, s);
        if (rc < 0) return rc;
        for (int l = 0; l < n; l++) p.lev[l].q = pl->lq[l];
        plan_bind(pl, p.q);
        nitems = pl->nitems;
    } else {
        nitems = (long long)p.lev[0].ncb * p.lev[0].nrc * batch;
    }
    if (nitems > 0x7fffffff) return 0;
    PDWT_PROF(prof_tag(n > 1 ? "k_inv2d_stream_levels" : "k_inv2d_stream", io[n - 1].Nr, io[n - 1].Nc), s);
    // PDWT_INV_SMEM_KB (experiments): pad the dynamic shared memory so that fewer one-warp CTAs share an SM
    static const size_t smem_pad = []() { const char* e = getenv("PDWT_INV_SMEM_KB"); return e ? (size_t)atoi(e) * 1024 : (size_t)0; }();
    const size_t smem_bytes = smem_pad > G::SMEM && smem_pad <= 96 * 1024 ? smem_pad : G::SMEM;
    cudaError_t e = launch_pdl(k_inv2d_stream<HLEN>, dim3((unsigned)nitems), 32, smem_bytes, s, p);
    if (e == cudaSuccess) {
        count_launch();
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        if (pl) pl->dirty = true;
        return note_cuda(e);
    }
    return n;
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

int s_dwt2_fwd_levels(const Taps& t, StreamPlans* plans, const StreamLevelIO* lv, int nlev, int batch, cudaStream_t s)
{
    PDWT_STREAM_HLEN_SWITCH(launch_fwd_stream, t, plans, lv, nlev, batch, s)
}

int s_dwt2_inv_levels(const Taps& t, StreamPlans* plans, const StreamLevelIO* lv, int nlev, int batch, cudaStream_t s)
{
    PDWT_STREAM_HLEN_SWITCH(launch_inv_stream, t, plans, lv, nlev, batch, s)
}

}  // namespace pdwt
