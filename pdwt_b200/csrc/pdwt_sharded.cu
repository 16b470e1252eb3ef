// pdwt_sharded.cu -- Layer C of the C ABI: a batch of independent planes spread over the GPUs of one node.
//
// The reference has no multi-GPU support (TODO.txt:15 "device selection").  The path shards by independent units
// (SURVEY 8e): image i of a batch is transformed on its own, so rank r of G owns a contiguous block of planes as ONE
// batched Wavelets object and the transform kernels never communicate.  The only data movement is input distribution
// and output collection, done here DEVICE TO DEVICE over NCCL (NVLink / NVSwitch): grouped ncclSend / ncclRecv from
// and to the root's device buffer, ncclAllGather for the per-plane norms.  Nothing bounces through the host.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries an NCCL -- torch.distributed
// brings its own -- gets that instance (same SONAME), a plain C++ client gets the system library, and the rest of
// libpdwt_b200.so works on machines without any NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <new>

#include "pdwt_common.cuh"
#include "pdwt_object.h"

using namespace pdwt;

namespace {

struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

thread_local char tl_nccl_err[256] = "";

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !api.so; i++) api.so = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!api.so) return;
#define PDWT_NCCL_SYM(field, sym)                                      \
    *reinterpret_cast<void**>(&api.field) = dlsym(api.so, sym);       \
    if (!api.field) return;
        PDWT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        PDWT_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        PDWT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        PDWT_NCCL_SYM(Send, "ncclSend")
        PDWT_NCCL_SYM(Recv, "ncclRecv")
        PDWT_NCCL_SYM(AllGather, "ncclAllGather")
        PDWT_NCCL_SYM(GroupStart, "ncclGroupStart")
        PDWT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        PDWT_NCCL_SYM(GetVersion, "ncclGetVersion")
        PDWT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PDWT_NCCL_SYM
        api.ok = true;
    });
    return api;
}

int nccl_rc(ncclResult_t r)
{
    if (r == ncclSuccess) return PDWT_OK;
    snprintf(tl_nccl_err, sizeof tl_nccl_err, "NCCL: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "error");
    return PDWT_ERR_CUDA;
}
#define PDWT_NCCL(call)                 \
    do {                                \
        int rc__ = nccl_rc(call);       \
        if (rc__ < 0) return rc__;      \
    } while (0)

}  // namespace

struct pdwt_shard {
    ncclComm_t comm;
    int nranks, rank;
    float* d_scratch;      // all-gather staging (per-plane scalars)
    size_t scratch_elems;
};


extern "C" {

const char* pdwt_shard_last_error(void) { return tl_nccl_err; }

int pdwt_shard_nccl_version(void)
{
    int v = 0;
    if (!nccl().ok || nccl().GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int pdwt_shard_unique_id(unsigned char id[128])
{
    if (!id) return PDWT_ERR_ARG;
    if (!nccl().ok) {
        snprintf(tl_nccl_err, sizeof tl_nccl_err, "libnccl.so.2 could not be loaded");
        return PDWT_ERR_CUDA;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    PDWT_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, &u, 128);
    return PDWT_OK;
}

int pdwt_shard_create(pdwt_shard** out, const unsigned char id[128], int nranks, int rank)
{
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return PDWT_ERR_ARG;
    *out = nullptr;
    if (!nccl().ok) {
        snprintf(tl_nccl_err, sizeof tl_nccl_err, "libnccl.so.2 could not be loaded");
        return PDWT_ERR_CUDA;
    }
    pdwt_shard* s = new (std::nothrow) pdwt_shard();
    if (!s) return PDWT_ERR_ALLOC;
    s->nranks = nranks;
    s->rank = rank;
    s->d_scratch = nullptr;
    s->scratch_elems = 0;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    const int rc = nccl_rc(nccl().CommInitRank(&s->comm, nranks, u, rank));   // binds the CURRENT device
    if (rc < 0) {
        delete s;
        return rc;
    }
    *out = s;
    return PDWT_OK;
}

void pdwt_shard_destroy(pdwt_shard* s)
{
    if (!s) return;
    if (s->d_scratch) cudaFree(s->d_scratch);
    nccl().CommDestroy(s->comm);
    delete s;
}

int pdwt_shard_rank(const pdwt_shard* s) { return s ? s->rank : PDWT_ERR_ARG; }
int pdwt_shard_nranks(const pdwt_shard* s) { return s ? s->nranks : PDWT_ERR_ARG; }

// contiguous blocks, the first n % nranks ranks own one plane more (no NCCL needed: pure arithmetic)
void pdwt_shard_block(long long n, int nranks, int rank, long long* first, long long* count)
{
    const long long base = n / nranks, extra = n % nranks;
    if (first) *first = rank * base + (rank < extra ? rank : extra);
    if (count) *count = base + (rank < extra ? 1 : 0);
}

// root's n planes of `plane` floats (device) -> every rank's block (device).  Grouped point-to-point: the root's NVLink
// egress is the only cost; its own block is a device-to-device copy.
int pdwt_shard_scatter(pdwt_shard* s, const float* d_full, float* d_mine, long long n, size_t plane, int root, void* stream)
{
    if (!s || n < 0 || root < 0 || root >= s->nranks) return PDWT_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    long long f, c;
    pdwt_shard_block(n, s->nranks, s->rank, &f, &c);
    if (s->rank == root && !d_full) return PDWT_ERR_ARG;
    if (c > 0 && !d_mine) return PDWT_ERR_ARG;
    PDWT_NCCL(nccl().GroupStart());
    int rc = PDWT_OK;
    if (s->rank == root) {
        for (int r = 0; r < s->nranks && rc == PDWT_OK; r++) {
            long long rf, rcnt;
            pdwt_shard_block(n, s->nranks, r, &rf, &rcnt);
            if (rcnt == 0 || r == root) continue;
            rc = nccl_rc(nccl().Send(d_full + (size_t)rf * plane, (size_t)rcnt * plane, ncclFloat, r, s->comm, st));
        }
    } else if (c > 0) {
        rc = nccl_rc(nccl().Recv(d_mine, (size_t)c * plane, ncclFloat, root, s->comm, st));
    }
    const int rc2 = nccl_rc(nccl().GroupEnd());
    if (rc < 0) return rc;
    if (rc2 < 0) return rc2;
    if (s->rank == root && c > 0 && d_mine != d_full + (size_t)f * plane)
        PDWT_CUDA(cudaMemcpyAsync(d_mine, d_full + (size_t)f * plane, sizeof(float) * (size_t)c * plane,
                                  cudaMemcpyDeviceToDevice, st));
    return PDWT_OK;
}

// every rank's block (contiguous planes of `plane` floats, device) -> the root's n planes (device)
int pdwt_shard_gather(pdwt_shard* s, const float* d_mine, float* d_full, long long n, size_t plane, int root, void* stream)
{
    if (!s || n < 0 || root < 0 || root >= s->nranks) return PDWT_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    long long f, c;
    pdwt_shard_block(n, s->nranks, s->rank, &f, &c);
    if (s->rank == root && !d_full) return PDWT_ERR_ARG;
    if (c > 0 && !d_mine) return PDWT_ERR_ARG;
    PDWT_NCCL(nccl().GroupStart());
    int rc = PDWT_OK;
    if (s->rank == root) {
        for (int r = 0; r < s->nranks && rc == PDWT_OK; r++) {
            long long rf, rcnt;
            pdwt_shard_block(n, s->nranks, r, &rf, &rcnt);
            if (rcnt == 0 || r == root) continue;
            rc = nccl_rc(nccl().Recv(d_full + (size_t)rf * plane, (size_t)rcnt * plane, ncclFloat, r, s->comm, st));
        }
    } else if (c > 0) {
        rc = nccl_rc(nccl().Send(d_mine, (size_t)c * plane, ncclFloat, root, s->comm, st));
    }
    const int rc2 = nccl_rc(nccl().GroupEnd());
    if (rc < 0) return rc;
    if (rc2 < 0) return rc2;
    if (s->rank == root && c > 0 && d_mine != d_full + (size_t)f * plane)
        PDWT_CUDA(cudaMemcpyAsync(d_full + (size_t)f * plane, d_mine, sizeof(float) * (size_t)c * plane,
                                  cudaMemcpyDeviceToDevice, st));
    return PDWT_OK;
}

// ---- the same with the rank's block held by a batched Wavelets object (its batch = this rank's plane count) ----------
static int check_obj(pdwt_shard* s, pdwt_wavelets* w, long long n, long long* cnt)
{
    if (!s || n < 0) return PDWT_ERR_ARG;
    long long f;
    pdwt_shard_block(n, s->nranks, s->rank, &f, cnt);
    if (*cnt > 0 && (!w || w->W.batch != *cnt || !w->W.d_image)) return PDWT_ERR_ARG;
    return PDWT_OK;
}

int pdwt_shard_scatter_image(pdwt_shard* s, pdwt_wavelets* w, const float* d_full, long long n, size_t plane, int root)
{
    long long c;
    const int rc0 = check_obj(s, w, n, &c);
    if (rc0 < 0) return rc0;
    if (c > 0 && plane != (size_t)w->W.winfos.Nr * w->W.winfos.Nc) return PDWT_ERR_ARG;
    const int rc = pdwt_shard_scatter(s, d_full, c > 0 ? w->W.d_image : nullptr, n, plane, root, c > 0 ? w->W.stream : nullptr);
    if (rc == PDWT_OK && c > 0 && w->W.state != W_CREATION_ERROR) {
        w->W.state = W_INIT;   // as set_image does (wt.cu:427-434)
        w->W.invalidate_norm_cache();
    }
    return rc;
}

int pdwt_shard_gather_image(pdwt_shard* s, pdwt_wavelets* w, float* d_full, long long n, size_t plane, int root)
{
    long long c;
    const int rc = check_obj(s, w, n, &c);
    if (rc < 0) return rc;
    if (c > 0 && plane != (size_t)w->W.winfos.Nr * w->W.winfos.Nc) return PDWT_ERR_ARG;
    return pdwt_shard_gather(s, c > 0 ? w->W.d_image : nullptr, d_full, n, plane, root, c > 0 ? w->W.stream : nullptr);
}

// sub-band `num`: `plane` = its logical size (pdwt_coeff_dims), identical on every rank.  The approximation band's
// planes sit further apart than their logical size (it doubles as scratch, common.cu:402-406): they are packed into
// d_tmp first (free after forward()), so that one message per peer suffices.
int pdwt_shard_gather_coeff(pdwt_shard* s, pdwt_wavelets* w, int num, float* d_full, long long n, size_t plane, int root)
{
    long long c;
    const int rc = check_obj(s, w, n, &c);
    if (rc < 0) return rc;
    const float* src = nullptr;
    cudaStream_t st = nullptr;
    if (c > 0) {
        int nr, nc;
        if (!w->W.d_coeffs || pdwt_coeff_dims(w->W.winfos, num, &nr, &nc) != PDWT_OK || (size_t)nr * nc != plane)
            return PDWT_ERR_ARG;
        if (w->W.state == W_INVERSE) return PDWT_ERR_STATE;   // wt.cu:476-479
        st = (cudaStream_t)w->W.stream;
        const size_t stride = pdwt_coeff_alloc_elems(w->W.winfos, num);
        src = w->W.d_coeffs[num];
        if (stride != plane) {
            PDWT_CUDA(cudaMemcpy2DAsync(w->W.d_tmp, plane * sizeof(float), src, stride * sizeof(float), plane * sizeof(float),
                                        (size_t)c, cudaMemcpyDeviceToDevice, st));
            src = w->W.d_tmp;
        }
    }
    return pdwt_shard_gather(s, src, d_full, n, plane, root, st);
}

// per-plane norms of the whole batch on EVERY rank (host array of n floats): which = 1 norm1, 2 norm2sq (wt.cu:398-418,
// 370-395).  One ncclAllGather of the padded local vectors.
int pdwt_shard_norms(pdwt_shard* s, pdwt_wavelets* w, int which, float* h_all, long long n)
{
    long long c;
    int rc = check_obj(s, w, n, &c);
    if (rc < 0) return rc;
    if (!h_all || (which != 1 && which != 2)) return PDWT_ERR_ARG;
    const long long maxc = (n + s->nranks - 1) / s->nranks;
    if (maxc == 0) return PDWT_OK;
    const size_t need = (size_t)maxc * (s->nranks + 1);
    if (s->scratch_elems < need) {
        if (s->d_scratch) cudaFree(s->d_scratch);
        s->d_scratch = nullptr;
        s->scratch_elems = 0;
        PDWT_CUDA(cudaMalloc(&s->d_scratch, sizeof(float) * need));
        s->scratch_elems = need;
    }
    float* h = (float*)calloc(need, sizeof(float));
    if (!h) return PDWT_ERR_ALLOC;
    cudaStream_t st = c > 0 ? (cudaStream_t)w->W.stream : nullptr;
    if (c > 0) rc = (which == 1) ? w->W.norm1_batched(h) : w->W.norm2sq_batched(h);
    cudaError_t e = cudaSuccess;
    if (rc == PDWT_OK) e = cudaMemcpyAsync(s->d_scratch, h, sizeof(float) * maxc, cudaMemcpyHostToDevice, st);
    if (rc == PDWT_OK && e == cudaSuccess)
        rc = nccl_rc(nccl().AllGather(s->d_scratch, s->d_scratch + maxc, (size_t)maxc, ncclFloat, s->comm, st));
    if (rc == PDWT_OK && e == cudaSuccess)
        e = cudaMemcpyAsync(h + maxc, s->d_scratch + maxc, sizeof(float) * maxc * s->nranks, cudaMemcpyDeviceToHost, st);
    if (rc == PDWT_OK && e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (rc == PDWT_OK && e == cudaSuccess)
        for (int r = 0; r < s->nranks; r++) {
            long long rf, rcnt;
            pdwt_shard_block(n, s->nranks, r, &rf, &rcnt);
            memcpy(h_all + rf, h + maxc * (r + 1), sizeof(float) * rcnt);
        }
    free(h);
    if (e != cudaSuccess) return note_cuda(e);
    return rc;
}

}  // extern "C"
