// pdwt_capi.cu -- Layer A of the C ABI (include/pdwt_b200.h): filter handles, the 16 transform drivers,
// threshold and norm drivers, size helpers.  Host-side level loops only; all arithmetic is in the kernels.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pdwt_common.cuh"
#include "filter_bank.inc"

using namespace pdwt;

struct pdwt_filters {
    Taps taps;
    char name[128];
    pdwt::StreamPlans* plans;   // queues + device counters of the cross-level launches made with this handle (lazy)
    // custom 2-D filter quadruples of the non-separable mode, [0] forward / [1] inverse (wt.cu:560-602): host copy and
    // device copy of 4 x hlen*hlen floats (LL, LH, HL, HH); NULL = outer products of the 1-D banks
    float* h_k2d[2];
    float* d_k2d[2];
};

namespace pdwt {
static thread_local int tl_last_cuda = 0;
static std::atomic<long long> g_launches{0};
int note_cuda(cudaError_t e)
{
    if (e == cudaSuccess) return PDWT_OK;
    tl_last_cuda = (int)e;
    return PDWT_ERR_CUDA;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// ---- per-kernel profiler -------------------------------------------------------------------------------------
struct ProfRec {
    const char* tag;
    cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<bool> g_prof_on{false};
bool profiling_on() { return g_prof_on.load(std::memory_order_relaxed); }
void prof_open(const char*, cudaStream_t s, void** ev0)
{
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    *ev0 = (void*)e;
}
void prof_close(const char* tag, cudaStream_t s, void* ev0)
{
    cudaEvent_t e1;
    if (cudaEventCreate(&e1) != cudaSuccess) return;
    cudaEventRecord(e1, s);
    std::lock_guard<std::mutex> g(g_prof_mu);
    g_prof.push_back(ProfRec{tag, (cudaEvent_t)ev0, e1});
}
const char* prof_tag(const char* base, int rows, int cols)
{
    if (!profiling_on()) return base;
    static std::map<std::string, std::string> interned;  // node-based: c_str() stays valid
    char buf[96];
    snprintf(buf, sizeof buf, "%s[%dx%d]", base, rows, cols);
    std::lock_guard<std::mutex> g(g_prof_mu);
    return interned.emplace(buf, buf).first->second.c_str();
}
// Kernel family used for the separable 2-D DWT: 2 = warp-streaming FFMA2 kernels, falling back per level to
// 1 = tile-fused kernels, falling back to 0 = generic two-pass kernels.  PDWT_PATH=stream|fused|generic (or the older
// PDWT_FORCE_GENERIC=1) caps the choice; the tests run all three against the same golden vectors.
static int path_cap()
{
    const char* e = getenv("PDWT_FORCE_GENERIC");
    if (e && *e && *e != '0') return 0;
    e = getenv("PDWT_PATH");
    if (e && !strcmp(e, "generic")) return 0;
    if (e && !strcmp(e, "fused")) return 1;
    return 2;
}
// Levels with at most this many output pixels (x batch) go to the tile kernels even when the streaming family could
// take them: a streaming warp needs hlen/2-1 row pairs of warm-up plus its chunk, serially, which is all latency when
// the whole level is a few hundred thousand pixels (PDWT_SMALL_PX overrides; 0 = always stream).
static long long small_level_px()
{
    const char* e = getenv("PDWT_SMALL_PX");   // read per call: the tests switch it inside one process
    return e ? atoll(e) : 0;
}
}  // namespace pdwt

extern "C" {

const char* pdwt_version(void) { return "pdwt_b200 0.1 (sm_100a)"; }
int pdwt_last_cuda_error(void) { return tl_last_cuda; }
const char* pdwt_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)tl_last_cuda); }
long long pdwt_launch_count(void) { return g_launches.load(); }
int pdwt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------ profiler
int pdwt_profile_begin(void)
{
    std::lock_guard<std::mutex> g(g_prof_mu);
    for (auto& r : g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    g_prof.clear();
    g_prof_on.store(true);
    return PDWT_OK;
}
int pdwt_profile_end(pdwt_profile_entry* out, int max_entries)
{
    g_prof_on.store(false);
    std::lock_guard<std::mutex> g(g_prof_mu);
    std::map<std::string, pdwt_profile_entry> agg;
    std::vector<std::string> order;
    int rc = PDWT_OK;
    for (auto& r : g_prof) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(r.e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (e != cudaSuccess) rc = note_cuda(e);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        auto it = agg.find(r.tag);
        if (it == agg.end()) {
            pdwt_profile_entry z;
            memset(&z, 0, sizeof z);
            strncpy(z.name, r.tag, sizeof(z.name) - 1);
            z.ms_min = 1e30;
            it = agg.emplace(r.tag, z).first;
            order.push_back(r.tag);
        }
        it->second.launches++;
        it->second.ms_total += ms;
        if (ms < it->second.ms_min) it->second.ms_min = ms;
        if (ms > it->second.ms_max) it->second.ms_max = ms;
    }
    g_prof.clear();
    if (rc < 0) return rc;
    int n = 0;
    for (auto& k : order) {
        if (out && n < max_entries) out[n] = agg[k];
        n++;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------- filters
int pdwt_wavelet_count(void) { return PDWT_NUM_BANKS; }
const char* pdwt_wavelet_name(int i) { return (i >= 0 && i < PDWT_NUM_BANKS) ? pdwt_bank_index[i].name : NULL; }

int pdwt_filters_create(pdwt_filters** out, const char* wname, int do_swt)
{
    if (!out || !wname) return PDWT_ERR_ARG;
    *out = NULL;
    pdwt_filters* f = (pdwt_filters*)calloc(1, sizeof(pdwt_filters));
    if (!f) return PDWT_ERR_ALLOC;
    strncpy(f->name, wname, sizeof(f->name) - 1);
    // Haar aliases use the dedicated kernels when the transform is decimated (separable.cu:24-28)
    if (!do_swt && (!strcasecmp(wname, "haar") || !strcasecmp(wname, "db1") || !strcasecmp(wname, "bior1.1") ||
                    !strcasecmp(wname, "rbior1.1"))) {
        f->taps.hlen = 2;
        *out = f;
        return 2;
    }
    for (int b = 0; b < PDWT_NUM_BANKS; b++) {
        if (strcasecmp(wname, pdwt_bank_index[b].name)) continue;
        const int n = pdwt_bank_index[b].hlen;
        const float* p = pdwt_bank_pool + pdwt_bank_index[b].offset;
        f->taps.hlen = n;
        memcpy(f->taps.L, p, n * sizeof(float));
        memcpy(f->taps.H, p + n, n * sizeof(float));
        memcpy(f->taps.IL, p + 2 * n, n * sizeof(float));
        memcpy(f->taps.IH, p + 3 * n, n * sizeof(float));
        *out = f;
        return n;
    }
    free(f);
    return PDWT_ERR_WAVELET;
}

int pdwt_filters_create_custom(pdwt_filters** out, int hlen, const float* dec_lo, const float* dec_hi,
                               const float* rec_lo, const float* rec_hi)
{
    if (!out || hlen < 1) return PDWT_ERR_ARG;
    *out = NULL;
    if (hlen > PDWT_MAX_FILTER_WIDTH) return PDWT_ERR_FILTER_LEN;
    pdwt_filters* f = (pdwt_filters*)calloc(1, sizeof(pdwt_filters));
    if (!f) return PDWT_ERR_ALLOC;
    f->taps.hlen = hlen;
    if (dec_lo) memcpy(f->taps.L, dec_lo, hlen * sizeof(float));
    if (dec_hi) memcpy(f->taps.H, dec_hi, hlen * sizeof(float));
    if (rec_lo) memcpy(f->taps.IL, rec_lo, hlen * sizeof(float));
    if (rec_hi) memcpy(f->taps.IH, rec_hi, hlen * sizeof(float));
    strcpy(f->name, "custom");
    *out = f;
    return hlen;
}

void pdwt_filters_destroy(pdwt_filters* f)
{
    if (!f) return;
    if (f->plans) stream_plans_destroy(f->plans);
    for (int d = 0; d < 2; d++) {
        free(f->h_k2d[d]);
        if (f->d_k2d[d]) cudaFree(f->d_k2d[d]);
    }
    free(f);
}
int pdwt_filters_hlen(const pdwt_filters* f) { return f ? f->taps.hlen : PDWT_ERR_ARG; }

// w_set_filters_forward_nonseparable / w_set_filters_inverse_nonseparable, nonseparable.cu:86-106: four hlen x hlen
// filters (row-major, the reference's array layout) for one direction.  Unlike the reference, where both directions
// share ONE set of __constant__ symbols (the last upload wins), each direction keeps its own quadruple.
int pdwt_filters_set_2d(pdwt_filters* f, int direction, const float* ll, const float* lh, const float* hl, const float* hh)
{
    if (!f || (direction != 1 && direction != -1) || !ll || !lh || !hl || !hh) return PDWT_ERR_ARG;
    const int d = direction > 0 ? 0 : 1, n = f->taps.hlen * f->taps.hlen;
    if (n < 1) return PDWT_ERR_ARG;
    float* h = (float*)malloc(sizeof(float) * 4 * n);
    if (!h) return PDWT_ERR_ALLOC;
    memcpy(h, ll, sizeof(float) * n);
    memcpy(h + n, lh, sizeof(float) * n);
    memcpy(h + 2 * n, hl, sizeof(float) * n);
    memcpy(h + 3 * n, hh, sizeof(float) * n);
    float* dv = nullptr;
    cudaError_t e = cudaMalloc(&dv, sizeof(float) * 4 * n);
    if (e == cudaSuccess) e = cudaMemcpy(dv, h, sizeof(float) * 4 * n, cudaMemcpyHostToDevice);   // blocking: visible to any stream
    if (e != cudaSuccess) {
        free(h);
        if (dv) cudaFree(dv);
        return note_cuda(e);
    }
    free(f->h_k2d[d]);
    if (f->d_k2d[d]) cudaFree(f->d_k2d[d]);
    f->h_k2d[d] = h;
    f->d_k2d[d] = dv;
    return PDWT_OK;
}
int pdwt_filters_has_2d(const pdwt_filters* f, int direction)
{
    return (f && (direction == 1 || direction == -1)) ? (f->d_k2d[direction > 0 ? 0 : 1] != nullptr) : 0;
}
int pdwt_filters_get(const pdwt_filters* f, float* dec_lo, float* dec_hi, float* rec_lo, float* rec_hi)
{
    if (!f) return PDWT_ERR_ARG;
    const size_t n = f->taps.hlen * sizeof(float);
    if (dec_lo) memcpy(dec_lo, f->taps.L, n);
    if (dec_hi) memcpy(dec_hi, f->taps.H, n);
    if (rec_lo) memcpy(rec_lo, f->taps.IL, n);
    if (rec_hi) memcpy(rec_hi, f->taps.IH, n);
    return f->taps.hlen;
}

// ---------------------------------------------------------------------------------------------- size helpers
int pdwt_div2(int n) { return half_up(n); }
int pdwt_num_coeffs(pdwt_w_info w) { return (w.ndims == 2 ? 3 : 1) * w.nlevels + 1; }

int pdwt_coeff_dims(pdwt_w_info w, int num, int* nr, int* nc)  // wt.cu:481-504
{
    if (num < 0 || num >= pdwt_num_coeffs(w)) return PDWT_ERR_ARG;
    const int scale = (num == 0) ? w.nlevels : (w.ndims == 2 ? (num - 1) / 3 + 1 : num);
    int r = w.Nr, c = w.Nc;
    if (!w.do_swt)
        for (int i = 0; i < scale; i++) {
            if (w.ndims == 2) r = half_up(r);
            c = half_up(c);
        }
    if (nr) *nr = r;
    if (nc) *nc = c;
    return PDWT_OK;
}

size_t pdwt_coeff_alloc_elems(pdwt_w_info w, int num)  // common.cu:400-445
{
    int r, c;
    if (num == 0) {  // level-1 sized (full size for the SWT): it doubles as scratch for the inverse
        r = (w.do_swt || w.ndims == 1) ? w.Nr : half_up(w.Nr);
        c = w.do_swt ? w.Nc : half_up(w.Nc);
        return (size_t)r * c;
    }
    if (pdwt_coeff_dims(w, num, &r, &c) != PDWT_OK) return 0;
    return (size_t)r * c;
}

int pdwt_max_level(int Nr, int Nc, int ndims, int hlen)  // wt.cu:156-159, utils.cu:14-20
{
    if (hlen < 2) return 0;
    int N = (ndims == 2) ? (Nr < Nc ? Nr : Nc) : Nc;
    int i = N / (hlen - 1), l = 0;
    while (i >>= 1) ++l;
    return l;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- driver guts
namespace {

struct Ctx {
    const Taps& t;
    float* img;
    float** c;
    float* tmp;
    pdwt_w_info w;
    int batch;
    cudaStream_t s;
    size_t s_img, s_tmp;
    StreamPlans* plans;
    Plane2 image() const { return Plane2{img, s_img}; }
    Plane2 scratch(size_t off = 0) const { return Plane2{tmp + off, s_tmp}; }
    Plane2 coeff(int k) const { return Plane2{c[k], pdwt_coeff_alloc_elems(w, k)}; }
};

#define TRY(x)                 \
    do {                       \
        int rc__ = (x);        \
        if (rc__ < 0) return rc__; \
    } while (0)

int check_args(const pdwt_filters* f, float* d_image, float** d_coeffs, float* d_tmp, pdwt_w_info w, int batch,
               bool need_filters)
{
    if ((need_filters && !f) || !d_image || !d_coeffs || !d_tmp) return PDWT_ERR_ARG;
    if (w.Nr < 1 || w.Nc < 1 || w.nlevels < 1 || batch < 1 || batch > 65535) return PDWT_ERR_ARG;
    if (w.ndims != 1 && w.ndims != 2) return PDWT_ERR_ARG;
    if (need_filters && (f->taps.hlen < 1 || f->taps.hlen > PDWT_MAX_FILTER_WIDTH)) return PDWT_ERR_ARG;
    return PDWT_OK;
}

// Where the approximation of level l (0-based, l < L-1 intermediate) lives when no D2D fix-up copy is wanted:
// the LAST level must land in d_coeffs[0], earlier ones alternate with d_tmp.
inline bool lands_in_c0(int L, int l) { return ((L - 1 - l) & 1) == 0; }

// ---- separable 2-D DWT --------------------------------------------------------------------------------------
// Buffer plan of the fused families: the approximation of level l (1 <= l < L) lives in d_tmp at its own offset (the
// levels never share storage, so the items of different levels may run concurrently inside one launch), the last one
// in d_coeffs[0] as the API demands.  No D2D fix-up copies (separable.cu:234) and, unlike the reference, inverse() leaves
// d_coeffs[0] intact.  d_tmp holds 2 Nr Nc floats per plane; the approximations need less than Nr Nc / 2.
Plane2 approx_plane(const Ctx& x, int l)
{
    if (l == x.w.nlevels) return x.coeff(0);
    size_t off = 0;
    for (int i = 1; i < l; i++)
        off += ((size_t)level_size(x.w.Nr, i, 0) * level_size(x.w.Nc, i, 0) + 63) & ~(size_t)63;
    return x.scratch(off);
}

int fwd_sep_2d(const Ctx& x)
{
    const int L = x.w.nlevels;
    int Nr = x.w.Nr, Nc = x.w.Nc;
    const int cap = path_cap();
    if (fused_supports_hlen(x.t.hlen) && cap >= 1 && L <= 32) {
        StreamLevelIO io[32];
        for (int l = 0; l < L; l++) {
            io[l].img = l ? approx_plane(x, l) : x.image();
            io[l].A = approx_plane(x, l + 1);
            io[l].H = x.coeff(3 * l + 1);
            io[l].V = x.coeff(3 * l + 2);
            io[l].D = x.coeff(3 * l + 3);
            io[l].Nr = level_size(x.w.Nr, l, 0);
            io[l].Nc = level_size(x.w.Nc, l, 0);
            io[l].a_reused = l + 1 < L;
        }
        for (int l = 0; l < L;) {
            int done = 0;
            if (cap >= 2) {
                int n = 0;   // consecutive levels large enough for the streaming family
                while (l + n < L && (long long)half_up(io[l + n].Nr) * half_up(io[l + n].Nc) * x.batch > small_level_px()) n++;
                if (n > 0) TRY(done = s_dwt2_fwd_levels(x.t, x.plans, io + l, n, x.batch, x.s));
            }
            if (!done) {
                TRY(f_dwt2_fwd_level(x.t, io[l].img, io[l].A, io[l].H, io[l].V, io[l].D, io[l].Nr, io[l].Nc, x.batch, x.s));
                done = 1;
            }
            l += done;
        }
        return PDWT_OK;
    }
    // generic two-pass scheme with the reference's buffer plan (separable.cu:179-209)
    Plane2 t1 = x.scratch(), t2 = x.scratch((size_t)x.w.Nr * half_up(x.w.Nc));
    for (int l = 0; l < L; l++) {
        Plane2 src = l ? x.coeff(0) : x.image();
        TRY(g_fwd_rows(x.t, src, t1, t2, Nr, Nc, x.batch, x.s));
        TRY(g_fwd_cols(x.t, t1, t2, x.coeff(0), x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), Nr,
                       half_up(Nc), x.batch, x.s));
        Nr = half_up(Nr);
        Nc = half_up(Nc);
    }
    return PDWT_OK;
}

int inv_sep_2d(const Ctx& x)
{
    const int L = x.w.nlevels;
    const int cap = path_cap();
    if (fused_supports_hlen(x.t.hlen) && cap >= 1 && L <= 32) {
        StreamLevelIO io[32];   // coarsest level first
        for (int k = 0; k < L; k++) {
            const int l = L - 1 - k;
            io[k].A = approx_plane(x, l + 1);
            io[k].H = x.coeff(3 * l + 1);
            io[k].V = x.coeff(3 * l + 2);
            io[k].D = x.coeff(3 * l + 3);
            io[k].img = l ? approx_plane(x, l) : x.image();
            io[k].Nr = level_size(x.w.Nr, l, 0);
            io[k].Nc = level_size(x.w.Nc, l, 0);
            io[k].a_reused = l > 0;
        }
        for (int k = 0; k < L;) {
            int done = 0;
            if (cap >= 2) {
                int n = 0;
                while (k + n < L && (long long)half_up(io[k + n].Nr) * half_up(io[k + n].Nc) * x.batch > small_level_px()) n++;
                if (n > 0) TRY(done = s_dwt2_inv_levels(x.t, x.plans, io + k, n, x.batch, x.s));
            }
            if (!done) {
                TRY(f_dwt2_inv_level(x.t, io[k].A, io[k].H, io[k].V, io[k].D, io[k].img, half_up(io[k].Nr), half_up(io[k].Nc),
                                     io[k].Nr, io[k].Nc, x.batch, x.s));
                done = 1;
            }
            k += done;
        }
        return PDWT_OK;
    }
    Plane2 t1 = x.scratch(), t2 = x.scratch((size_t)x.w.Nr * half_up(x.w.Nc));  // separable.cu:343-344
    for (int l = L - 1; l >= 0; l--) {
        const int Mr = level_size(x.w.Nr, l, 0), Mc = level_size(x.w.Nc, l, 0);
        const int nr = half_up(Mr), nc = half_up(Mc);
        TRY(g_inv_cols(x.t, x.coeff(0), x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), t1, t2, nr, nc, Mr,
                       x.batch, x.s));
        TRY(g_inv_rows(x.t, t1, t2, l ? x.coeff(0) : x.image(), Mr, nc, Mc, x.batch, x.s));
    }
    return PDWT_OK;
}

// ---- (batched) 1-D: row passes only; A ping-pongs so that the last level lands in d_coeffs[0] ---------------
// every level of a 1-D DWT from one launch (pdwt_rows_all.cu); false: run the per-level kernels
bool rows_all(const Ctx& x, bool inverse, int* rc, bool swt = false)
{
    const int L = x.w.nlevels;
    *rc = PDWT_OK;
    if (path_cap() < 1 || L < 2 || L > 32) return false;
    Plane2 bands[33];
    for (int l = 0; l <= L; l++) bands[l] = x.coeff(l);
    const int done = swt ? (inverse ? r_swt1_inv_all(x.t, x.image(), bands, x.w.Nr, x.w.Nc, L, x.batch, x.s)
                                    : r_swt1_fwd_all(x.t, x.image(), bands, x.w.Nr, x.w.Nc, L, x.batch, x.s))
                   : inverse ? r_dwt1_inv_all(x.t, x.image(), bands, x.w.Nr, x.w.Nc, L, x.batch, x.s)
                             : r_dwt1_fwd_all(x.t, x.image(), bands, x.w.Nr, x.w.Nc, L, x.batch, x.s);
    if (done < 0) *rc = done;
    return done != 0;
}

int fwd_sep_1d(const Ctx& x, bool swt)
{
    const int L = x.w.nlevels;
    int Nc = x.w.Nc;
    int rc_all;
    if (rows_all(x, false, &rc_all, swt)) return rc_all;
    Plane2 cur = x.image();
    for (int l = 0; l < L; l++) {
        Plane2 dstA = lands_in_c0(L, l) ? x.coeff(0) : x.scratch();
        if (swt)
            TRY(g_swt_fwd_rows(x.t, cur, dstA, x.coeff(l + 1), x.w.Nr, Nc, l + 1, x.batch, x.s));
        else {
            TRY(g_fwd_rows(x.t, cur, dstA, x.coeff(l + 1), x.w.Nr, Nc, x.batch, x.s));
            Nc = half_up(Nc);
        }
        cur = dstA;
    }
    return PDWT_OK;
}

int inv_sep_1d(const Ctx& x, bool swt)
{
    const int L = x.w.nlevels;
    int rc_all;
    if (rows_all(x, true, &rc_all, swt)) return rc_all;
    Plane2 cur = x.coeff(0);
    bool cur_is_c0 = true;
    for (int l = L - 1; l >= 0; l--) {
        Plane2 dst = (l == 0) ? x.image() : (cur_is_c0 ? x.scratch() : x.coeff(0));
        if (swt)
            TRY(g_swt_inv_rows(x.t, cur, x.coeff(l + 1), dst, x.w.Nr, x.w.Nc, l + 1, x.batch, x.s));
        else
            TRY(g_inv_rows(x.t, cur, x.coeff(l + 1), dst, x.w.Nr, level_size(x.w.Nc, l + 1, 0),
                           level_size(x.w.Nc, l, 0), x.batch, x.s));
        cur = dst;
        cur_is_c0 = !cur_is_c0;
    }
    return PDWT_OK;
}

// ---- separable 2-D SWT (reference buffer plan, separable.cu:496-515 / 629-649) -------------------------------
// Fused SWT levels ping-pong the approximation between d_coeffs[0] and d_tmp (a level must not overwrite the plane its
// neighbours' halos still read), arranged so that the last level lands in d_coeffs[0].  The generic two-pass scheme
// needs all of d_tmp for its row-pass outputs, so the choice is made once per transform: fused only if EVERY level fits.
bool swt_all_levels_fused(const Ctx& x)
{
    if (path_cap() < 1) return false;
    return w_swt2_supported(x.t, x.w.Nr, x.w.Nc, x.w.nlevels, x.batch) != 0;
}

int fwd_swt_2d(const Ctx& x)
{
    const int Nr = x.w.Nr, Nc = x.w.Nc, L = x.w.nlevels;
    if (swt_all_levels_fused(x)) {
        Plane2 cur = x.image();
        for (int l = 0; l < L; l++) {
            Plane2 dstA = lands_in_c0(L, l) ? x.coeff(0) : x.scratch();
            int done = 0;
            TRY(done = w_swt2_fwd_level(x.t, cur, dstA, x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), Nr, Nc,
                                        l + 1, x.batch, x.s));
            if (!done) return PDWT_ERR_ARG;   // cannot happen: swt_all_levels_fused() said yes
            cur = dstA;
        }
        return PDWT_OK;
    }
    Plane2 t1 = x.scratch(), t2 = x.scratch((size_t)Nr * Nc);
    for (int l = 0; l < L; l++) {
        TRY(g_swt_fwd_rows(x.t, l ? x.coeff(0) : x.image(), t1, t2, Nr, Nc, l + 1, x.batch, x.s));
        TRY(g_swt_fwd_cols(x.t, t1, t2, x.coeff(0), x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), Nr, Nc,
                           l + 1, x.batch, x.s));
    }
    return PDWT_OK;
}

int inv_swt_2d(const Ctx& x)
{
    const int Nr = x.w.Nr, Nc = x.w.Nc, L = x.w.nlevels;
    if (swt_all_levels_fused(x)) {
        Plane2 cur = x.coeff(0);
        bool cur_is_c0 = true;
        for (int l = L - 1; l >= 0; l--) {
            Plane2 dst = (l == 0) ? x.image() : (cur_is_c0 ? x.scratch() : x.coeff(0));
            int done = 0;
            TRY(done = w_swt2_inv_level(x.t, cur, x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), dst, Nr, Nc,
                                        l + 1, x.batch, x.s));
            if (!done) return PDWT_ERR_ARG;
            cur = dst;
            cur_is_c0 = !cur_is_c0;
        }
        return PDWT_OK;
    }
    Plane2 t1 = x.scratch(), t2 = x.scratch((size_t)Nr * Nc);
    for (int l = L - 1; l >= 0; l--) {
        TRY(g_swt_inv_cols(x.t, x.coeff(0), x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), t1, t2, Nr, Nc,
                           l + 1, x.batch, x.s));
        TRY(g_swt_inv_rows(x.t, t1, t2, l ? x.coeff(0) : x.image(), Nr, Nc, l + 1, x.batch, x.s));
    }
    return PDWT_OK;
}

// ---- single-kernel-per-level families: Haar 2-D / 1-D, non-separable DWT / SWT -----------------------------------
enum Family { HAAR2, HAAR1, NONSEP, NONSEP_SWT };

int fwd_single(const Ctx& x, Family fam)
{
    const int L = x.w.nlevels;
    int Nr = x.w.Nr, Nc = x.w.Nc;
    Plane2 cur = x.image();
    for (int l = 0; l < L; l++) {
        Plane2 dstA = lands_in_c0(L, l) ? x.coeff(0) : x.scratch();
        switch (fam) {
            case HAAR2:
                TRY(g_haar2d_fwd(cur, dstA, x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), Nr, Nc, x.batch,
                                 x.s));
                Nr = half_up(Nr);
                Nc = half_up(Nc);
                break;
            case HAAR1:
                TRY(g_haar1d_fwd(cur, dstA, x.coeff(l + 1), x.w.Nr, Nc, x.batch, x.s));
                Nc = half_up(Nc);
                break;
            case NONSEP: {
                int done = 0;
                if (path_cap() >= 1)
                    TRY(done = n_nonsep_fwd_level(x.t, cur, dstA, x.coeff(3 * l + 1), x.coeff(3 * l + 2),
                                                  x.coeff(3 * l + 3), Nr, Nc, x.batch, x.s));
                if (!done)
                    TRY(g_nonsep_fwd(x.t, cur, dstA, x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), Nr, Nc,
                                     x.batch, x.s));
            }
                Nr = half_up(Nr);
                Nc = half_up(Nc);
                break;
            case NONSEP_SWT: {
                int done = 0;
                if (path_cap() >= 1)
                    TRY(done = n_nonsep_swt_fwd_level(x.t, cur, dstA, x.coeff(3 * l + 1), x.coeff(3 * l + 2),
                                                      x.coeff(3 * l + 3), Nr, Nc, l + 1, x.batch, x.s));
                if (!done)
                    TRY(g_nonsep_swt_fwd(x.t, cur, dstA, x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), Nr, Nc,
                                         l + 1, x.batch, x.s));
                break;
            }
        }
        cur = dstA;
    }
    return PDWT_OK;
}

int inv_single(const Ctx& x, Family fam)
{
    const int L = x.w.nlevels;
    Plane2 cur = x.coeff(0);
    bool cur_is_c0 = true;
    for (int l = L - 1; l >= 0; l--) {
        Plane2 dst = (l == 0) ? x.image() : (cur_is_c0 ? x.scratch() : x.coeff(0));
        const int Mr = level_size(x.w.Nr, l, 0), Mc = level_size(x.w.Nc, l, 0);
        switch (fam) {
            case HAAR2:
                TRY(g_haar2d_inv(dst, cur, x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), half_up(Mr),
                                 half_up(Mc), Mr, Mc, x.batch, x.s));
                break;
            case HAAR1:
                TRY(g_haar1d_inv(dst, cur, x.coeff(l + 1), x.w.Nr, half_up(Mc), Mc, x.batch, x.s));
                break;
            case NONSEP: {
                int done = 0;
                if (path_cap() >= 1)
                    TRY(done = n_nonsep_inv_level(x.t, dst, cur, x.coeff(3 * l + 1), x.coeff(3 * l + 2),
                                                  x.coeff(3 * l + 3), half_up(Mr), half_up(Mc), Mr, Mc, x.batch, x.s));
                if (!done)
                    TRY(g_nonsep_inv(x.t, dst, cur, x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3),
                                     half_up(Mr), half_up(Mc), Mr, Mc, x.batch, x.s));
                break;
            }
            case NONSEP_SWT: {
                int done = 0;
                if (path_cap() >= 1)
                    TRY(done = n_nonsep_swt_inv_level(x.t, dst, cur, x.coeff(3 * l + 1), x.coeff(3 * l + 2),
                                                      x.coeff(3 * l + 3), x.w.Nr, x.w.Nc, l + 1, x.batch, x.s));
                if (!done)
                    TRY(g_nonsep_swt_inv(x.t, dst, cur, x.coeff(3 * l + 1), x.coeff(3 * l + 2), x.coeff(3 * l + 3), x.w.Nr,
                                         x.w.Nc, l + 1, x.batch, x.s));
                break;
            }
        }
        cur = dst;
        cur_is_c0 = !cur_is_c0;
    }
    return PDWT_OK;
}

const Taps kNoTaps = {};

// the handle's plan cache, created on first use (the handle is `const` for the caller: the cache is an implementation detail)
StreamPlans* plans_of(const pdwt_filters* f)
{
    if (!f) return nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    pdwt_filters* m = const_cast<pdwt_filters*>(f);
    if (!m->plans) m->plans = stream_plans_create();
    return m->plans;
}

}  // namespace

// the taps a driver hands to its kernels: the non-separable drivers add the custom quadruple of their direction
static Taps taps_for(const pdwt_filters* f, int direction)
{
    Taps t = f->taps;
    t.k2d = t.hk2d = nullptr;
    if (direction) {
        t.k2d = f->d_k2d[direction > 0 ? 0 : 1];
        t.hk2d = f->h_k2d[direction > 0 ? 0 : 1];
    }
    return t;
}
#define MAKE_CTX_NS(direction)                                                                \
    TRY(check_args(f, d_image, d_coeffs, d_tmp, winfos, batch, true));                        \
    const Taps taps_ns = taps_for(f, direction);                                              \
    Ctx x{taps_ns, d_image, d_coeffs, d_tmp, winfos, batch, (cudaStream_t)stream,             \
          (size_t)winfos.Nr * winfos.Nc, 2 * (size_t)winfos.Nr * winfos.Nc, plans_of(f)}

#define MAKE_CTX(need_f)                                                                      \
    TRY(check_args(f, d_image, d_coeffs, d_tmp, winfos, batch, need_f));                      \
    Ctx x{(need_f) ? f->taps : kNoTaps, d_image, d_coeffs, d_tmp, winfos, batch, (cudaStream_t)stream, \
          (size_t)winfos.Nr * winfos.Nc, 2 * (size_t)winfos.Nr * winfos.Nc, plans_of(f)}

extern "C" {

int pdwt_forward_separable(PDWT_DRIVER_ARGS) { MAKE_CTX(true); return fwd_sep_2d(x); }
int pdwt_forward_separable_1d(PDWT_DRIVER_ARGS) { MAKE_CTX(true); return fwd_sep_1d(x, false); }
int pdwt_inverse_separable(PDWT_DRIVER_ARGS) { MAKE_CTX(true); return inv_sep_2d(x); }
int pdwt_inverse_separable_1d(PDWT_DRIVER_ARGS) { MAKE_CTX(true); return inv_sep_1d(x, false); }
int pdwt_forward_swt_separable(PDWT_DRIVER_ARGS) { MAKE_CTX(true); return fwd_swt_2d(x); }
int pdwt_forward_swt_separable_1d(PDWT_DRIVER_ARGS) { MAKE_CTX(true); return fwd_sep_1d(x, true); }
int pdwt_inverse_swt_separable(PDWT_DRIVER_ARGS) { MAKE_CTX(true); return inv_swt_2d(x); }
int pdwt_inverse_swt_separable_1d(PDWT_DRIVER_ARGS) { MAKE_CTX(true); return inv_sep_1d(x, true); }
int pdwt_haar_forward2d(PDWT_DRIVER_ARGS) { MAKE_CTX(false); return fwd_single(x, HAAR2); }
int pdwt_haar_inverse2d(PDWT_DRIVER_ARGS) { MAKE_CTX(false); return inv_single(x, HAAR2); }
int pdwt_haar_forward1d(PDWT_DRIVER_ARGS) { MAKE_CTX(false); return fwd_single(x, HAAR1); }
int pdwt_haar_inverse1d(PDWT_DRIVER_ARGS) { MAKE_CTX(false); return inv_single(x, HAAR1); }
int pdwt_forward_nonseparable(PDWT_DRIVER_ARGS) { MAKE_CTX_NS(1); return fwd_single(x, NONSEP); }
int pdwt_inverse_nonseparable(PDWT_DRIVER_ARGS) { MAKE_CTX_NS(-1); return inv_single(x, NONSEP); }
int pdwt_forward_swt_nonseparable(PDWT_DRIVER_ARGS) { MAKE_CTX_NS(1); return fwd_single(x, NONSEP_SWT); }
int pdwt_inverse_swt_nonseparable(PDWT_DRIVER_ARGS) { MAKE_CTX_NS(-1); return inv_single(x, NONSEP_SWT); }

// Wavelets::forward dispatch, wt.cu:247-266
int pdwt_forward(PDWT_DRIVER_ARGS, int do_separable)
{
    const bool haar = (winfos.hlen == 2) && !winfos.do_swt;
    if (winfos.ndims == 1) {
        if (haar) return pdwt_haar_forward1d(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
        return winfos.do_swt ? pdwt_forward_swt_separable_1d(f, d_image, d_coeffs, d_tmp, winfos, batch, stream)
                             : pdwt_forward_separable_1d(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
    }
    if (winfos.ndims != 2) return PDWT_ERR_ARG;
    if (haar) return pdwt_haar_forward2d(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
    if (do_separable)
        return winfos.do_swt ? pdwt_forward_swt_separable(f, d_image, d_coeffs, d_tmp, winfos, batch, stream)
                             : pdwt_forward_separable(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
    return winfos.do_swt ? pdwt_forward_swt_nonseparable(f, d_image, d_coeffs, d_tmp, winfos, batch, stream)
                         : pdwt_forward_nonseparable(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
}

// Wavelets::inverse dispatch, wt.cu:283-303
int pdwt_inverse(PDWT_DRIVER_ARGS, int do_separable)
{
    const bool haar = (winfos.hlen == 2) && !winfos.do_swt;
    if (winfos.ndims == 1) {
        if (haar) return pdwt_haar_inverse1d(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
        return winfos.do_swt ? pdwt_inverse_swt_separable_1d(f, d_image, d_coeffs, d_tmp, winfos, batch, stream)
                             : pdwt_inverse_separable_1d(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
    }
    if (winfos.ndims != 2) return PDWT_ERR_ARG;
    if (haar) return pdwt_haar_inverse2d(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
    if (do_separable)
        return winfos.do_swt ? pdwt_inverse_swt_separable(f, d_image, d_coeffs, d_tmp, winfos, batch, stream)
                             : pdwt_inverse_separable(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
    return winfos.do_swt ? pdwt_inverse_swt_nonseparable(f, d_image, d_coeffs, d_tmp, winfos, batch, stream)
                         : pdwt_inverse_nonseparable(f, d_image, d_coeffs, d_tmp, winfos, batch, stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------ thresholds and norms
namespace {

// segment table of every sub-band in the reference's visiting order: details of level 1..L first, A last
// (wt.cu:404-417).  beta_of(level) < 0 skips nothing; `with_app` adds A_L (logical size).
int launch_tables(float** c, pdwt_w_info w, int batch, cudaStream_t s, bool with_details, bool with_app,
                  const float* beta_levels, float beta_app, int op /*0 soft,1 hard,2 asum,3 sumsq,4 proj_linf,5 scale*/,
                  double* d_sums,
                  int* nseg_out, bool app_readonly = false, const HostPublish* hp = nullptr)
{
    SegTable tab;
    memset(&tab, 0, sizeof tab);
    int total = 0;
    auto flush = [&](void) -> int {
        if (tab.nseg == 0) return 0;
        int rc;
        if (op < 2)
            rc = e_threshold(tab, op, batch, s, d_sums, d_sums ? hp : nullptr);   // d_sums != NULL: the norms of the result come with it
        else if (op < 4)
            rc = e_reduce(tab, op - 2, batch, d_sums, s, hp);
        else
            rc = e_threshold(tab, op - 2, batch, s);   // element-wise ops 2 (proj_linf) and 3 (scale)
        tab.nseg = 0;
        return rc;
    };
    auto push = [&](int k, float beta) -> int {
        int r, cdim;
        if (pdwt_coeff_dims(w, k, &r, &cdim) != PDWT_OK) return PDWT_ERR_ARG;
        if (!c[k]) return PDWT_ERR_ARG;
        if ((op == 2 || op == 3 || d_sums) && tab.nseg == kMaxSeg) return PDWT_ERR_ARG;  // sums need ONE table (layout)
        if (tab.nseg == kMaxSeg) TRY(flush());
        tab.ptr[tab.nseg] = c[k];
        tab.n[tab.nseg] = (unsigned long long)r * cdim;
        tab.stride[tab.nseg] = pdwt_coeff_alloc_elems(w, k);
        tab.beta[tab.nseg] = beta;
        tab.ro[tab.nseg] = (k == 0 && app_readonly) ? 1 : 0;
        tab.nseg++;
        total++;
        return 0;
    };
    const int per = (w.ndims == 2) ? 3 : 1;
    if (with_details)
        for (int l = 0; l < w.nlevels; l++)
            for (int b = 1; b <= per; b++) TRY(push(per * l + b, beta_levels ? beta_levels[l] : 0.f));
    if (with_app) TRY(push(0, beta_app));
    if (nseg_out) *nseg_out = total;
    return flush();
}

// w_call_soft_thresh / w_call_hard_thresh, common.cu:219-282.  d_sums2 != NULL (2 * batch * ncoeffs doubles on the device):
// the same launch leaves sum |c| and sum c^2 of every sub-band AFTER the threshold there, A_L included (read-only when
// it is not thresholded) -- SURVEY 8f N1.
// hp != NULL: d_sums2 is zero on entry, the launch publishes the sums to the host and leaves it zero (HostPublish).
int threshold_impl(float** c, float beta, pdwt_w_info w, int app, int normalize, int batch, cudaStream_t s, int hard,
                   double* d_sums2 = nullptr, const HostPublish* hp = nullptr)
{
    if (!c || w.nlevels < 1 || w.nlevels > 32 || batch < 1) return PDWT_ERR_ARG;
    if (d_sums2) {
        if (pdwt_num_coeffs(w) > kMaxSeg) d_sums2 = nullptr;
        else if (!hp) PDWT_CUDA(cudaMemsetAsync(d_sums2, 0, sizeof(double) * 2 * batch * pdwt_num_coeffs(w), s));
    }
    float beta_app = beta;
    if (app && !hard && normalize > 0) {  // beta2 = beta / sqrt(2)^nlevels (soft only; hard passes beta, common.cu:270)
        const int half = w.nlevels / 2;
        beta_app /= (float)(1 << half);
        if (half * 2 != w.nlevels) beta_app = (float)((double)beta_app / 1.4142135623730951);
    }
    float bl[32];
    for (int l = 0; l < w.nlevels; l++) {
        if (normalize > 0) beta = (float)((double)beta / 1.4142135623730951);  // common.cu:244 (SQRT_2 is a double)
        bl[l] = beta;
    }
    if (d_sums2) return launch_tables(c, w, batch, s, true, true, bl, beta_app, hard ? 1 : 0, d_sums2, nullptr, app == 0, hp);
    return launch_tables(c, w, batch, s, true, app != 0, bl, beta_app, hard ? 1 : 0, nullptr, nullptr);
}

// w_call_proj_linf (common.cu:285-308) and w_shrink (common.cu:343-369): one element-wise launch over the sub-bands
int elementwise_impl(float** c, float beta, pdwt_w_info w, int app, int batch, cudaStream_t s, int op)
{
    if (!c || w.nlevels < 1 || w.nlevels > 32 || batch < 1) return PDWT_ERR_ARG;
    float bl[32];
    for (int l = 0; l < w.nlevels; l++) bl[l] = beta;
    return launch_tables(c, w, batch, s, true, app != 0, bl, beta, op, nullptr, nullptr);
}

// w_call_group_soft_thresh, common.cu:311-341
int group_soft_impl(float** c, float beta, pdwt_w_info w, int app, int normalize, int batch, cudaStream_t s)
{
    if (!c || w.nlevels < 1 || w.nlevels > 32 || batch < 1) return PDWT_ERR_ARG;
    GroupTable tab;
    memset(&tab, 0, sizeof tab);
    tab.nlev = w.nlevels;
    tab.stride_a = pdwt_coeff_alloc_elems(w, 0);
    for (int l = 0; l < w.nlevels; l++) {
        if (normalize > 0) beta = (float)((double)beta / 1.4142135623730951);   // common.cu:335 (SQRT_2 is a double)
        const int kd = (w.ndims == 2) ? 3 * l + 3 : l + 1;
        int r, cc;
        if (pdwt_coeff_dims(w, kd, &r, &cc) != PDWT_OK || !c[kd]) return PDWT_ERR_ARG;
        if (w.ndims == 2) {
            if (!c[3 * l + 1] || !c[3 * l + 2]) return PDWT_ERR_ARG;
            tab.h[l] = c[3 * l + 1];
            tab.v[l] = c[3 * l + 2];
        }
        tab.d[l] = c[kd];
        tab.a[l] = (app && l == w.nlevels - 1) ? c[0] : nullptr;
        tab.n[l] = (unsigned long long)r * cc;
        tab.stride_d[l] = pdwt_coeff_alloc_elems(w, kd);
        tab.beta[l] = beta;
    }
    return e_group_soft(tab, batch, s);
}

}  // namespace

namespace pdwt {
int norm_finish(pdwt_w_info w, int batch, int mode, float* out, cudaStream_t s, const double* d_sums, double* h_sums);
// thresholds of the Wavelets object: d_sums2 receives the norms of the result (see threshold_impl)
int threshold_norms(float** c, float beta, pdwt_w_info w, int app, int normalize, int batch, cudaStream_t s, int hard,
                    double* d_sums2, const HostPublish* hp)
{
    return threshold_impl(c, beta, w, app, normalize, batch, s, hard, d_sums2, hp);
}
// shared with the Wavelets object: reduction with caller-provided scratch (d_sums/h_sums: batch*kMaxSeg doubles)
int norm_impl(float** c, pdwt_w_info w, int batch, int mode, float* out, cudaStream_t s, double* d_sums, double* h_sums)
{
    if (!c || !out || w.nlevels < 1 || batch < 1) return PDWT_ERR_ARG;
    const int nseg = pdwt_num_coeffs(w);
    if (nseg > kMaxSeg) return PDWT_ERR_ARG;
    PDWT_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * batch * nseg, s));
    int pushed = 0;
    TRY(launch_tables(c, w, batch, s, true, true, nullptr, 0.f, 2 + mode, d_sums, &pushed));
    return norm_finish(w, batch, mode, out, s, d_sums, h_sums);
}

// one double per plane and sub-band (details first, A last) -> the reference's float accumulation, sub-band by sub-band
// in its order (wt.cu:373-394, 400-417)
void norm_accumulate(pdwt_w_info w, int batch, int mode, float* out, const double* h_sums)
{
    const int nseg = pdwt_num_coeffs(w);
    for (int p = 0; p < batch; p++) {
        float res = 0.0f;
        for (int k = 0; k < nseg; k++) {
            const double v = h_sums[(size_t)p * nseg + k];
            if (mode == 0)
                res += (float)v;
            else {
                const float nrm2 = (float)sqrt(v);
                res += nrm2 * nrm2;
            }
        }
        out[p] = res;
    }
}

// the sums are on the device: fetch them, then accumulate
int norm_finish(pdwt_w_info w, int batch, int mode, float* out, cudaStream_t s, const double* d_sums, double* h_sums)
{
    const int nseg = pdwt_num_coeffs(w);
    PDWT_CUDA(cudaMemcpyAsync(h_sums, d_sums, sizeof(double) * batch * nseg, cudaMemcpyDeviceToHost, s));
    PDWT_CUDA(cudaStreamSynchronize(s));
    norm_accumulate(w, batch, mode, out, h_sums);
    return PDWT_OK;
}

// ---- HostPublish, host side ----------------------------------------------------------------------------------
unsigned next_publish_tag()
{
    static unsigned tag = 0;   // process-wide; 0 is what fresh buffers hold
    unsigned t;
    do t = __atomic_add_fetch(&tag, 1u, __ATOMIC_RELAXED); while (t == 0);
    return t;
}
// Spin until the launch that carries `tag` has published its nsums doubles into the pinned words h; they are decoded
// into vals.  Every 1024 probes the stream is queried: if it has drained (or failed) and the words are still not there,
// the caller gets an error instead of a hang.
int wait_published(const unsigned long long* h, int nsums, unsigned tag, double* vals, cudaStream_t s)
{
    const volatile unsigned long long* w = h;
    unsigned spins = 0;
    bool drained = false;
    for (int i = 0; i < nsums; i++) {
        unsigned long long a, b;
        for (;;) {
            a = w[2 * i];
            b = w[2 * i + 1];
            if ((unsigned)(a >> 32) == tag && (unsigned)(b >> 32) == tag) break;
            if (drained) return PDWT_ERR_CUDA;   // the stream is idle and nothing arrived
            if ((++spins & 1023u) == 0) {
                const cudaError_t e = cudaStreamQuery(s);
                if (e == cudaErrorNotReady) continue;
                if (e != cudaSuccess) return note_cuda(e);
                drained = true;   // one more look at the words, then give up
            }
        }
        const unsigned long long bits = (a & 0xffffffffull) | (b << 32);
        memcpy(&vals[i], &bits, sizeof bits);
    }
    return PDWT_OK;
}
// norm of the coefficients with the result delivered through hp (d_sums zero on entry and on exit)
int norm_published(float** c, pdwt_w_info w, int batch, int mode, float* out, cudaStream_t s, double* d_sums,
                   const HostPublish& hp, const unsigned long long* h_words, double* vals)
{
    if (!c || !out || w.nlevels < 1 || batch < 1) return PDWT_ERR_ARG;
    if (pdwt_num_coeffs(w) > kMaxSeg) return PDWT_ERR_ARG;
    int pushed = 0;
    TRY(launch_tables(c, w, batch, s, true, true, nullptr, 0.f, 2 + mode, d_sums, &pushed, false, &hp));
    TRY(wait_published(h_words, hp.nsums, hp.tag, vals, s));
    norm_accumulate(w, batch, mode, out, vals);
    return PDWT_OK;
}
}  // namespace pdwt

extern "C" {

int pdwt_call_soft_thresh(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int normalize,
                          int batch, void* stream)
{
    return threshold_impl(d_coeffs, beta, winfos, do_thresh_appcoeffs, normalize, batch, (cudaStream_t)stream, 0);
}
int pdwt_call_hard_thresh(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int normalize,
                          int batch, void* stream)
{
    return threshold_impl(d_coeffs, beta, winfos, do_thresh_appcoeffs, normalize, batch, (cudaStream_t)stream, 1);
}

int pdwt_call_group_soft_thresh(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int normalize,
                                int batch, void* stream)
{
    return group_soft_impl(d_coeffs, beta, winfos, do_thresh_appcoeffs, normalize, batch, (cudaStream_t)stream);
}
int pdwt_call_proj_linf(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int batch, void* stream)
{
    return elementwise_impl(d_coeffs, beta, winfos, do_thresh_appcoeffs, batch, (cudaStream_t)stream, 4);
}
int pdwt_shrink(float** d_coeffs, float beta, pdwt_w_info winfos, int do_thresh_appcoeffs, int batch, void* stream)
{
    return elementwise_impl(d_coeffs, 1.0f / (1.0f + beta), winfos, do_thresh_appcoeffs, batch, (cudaStream_t)stream, 5);
}
/* w_add_coeffs / w_add_coeffs_1d, common.cu:499-526: dst += alpha * src for every sub-band (logical sizes; the
 * reference's 1-D variant halves with floor and so skips the tail of odd-sized levels -- fixed here) */
int pdwt_add_coeffs(float** dst, float** src, pdwt_w_info w, float alpha, int batch, void* stream)
{
    if (!dst || !src || w.nlevels < 1 || batch < 1) return PDWT_ERR_ARG;
    const int nseg = pdwt_num_coeffs(w);
    if (nseg > kMaxSeg) return PDWT_ERR_ARG;
    PairTable tab;
    memset(&tab, 0, sizeof tab);
    for (int k = 0; k < nseg; k++) {
        int r, c;
        if (pdwt_coeff_dims(w, k, &r, &c) != PDWT_OK || !dst[k] || !src[k]) return PDWT_ERR_ARG;
        tab.dst[k] = dst[k];
        tab.src[k] = src[k];
        tab.n[k] = (unsigned long long)r * c;
        tab.stride_dst[k] = tab.stride_src[k] = pdwt_coeff_alloc_elems(w, k);
    }
    tab.nseg = nseg;
    return e_axpy(tab, alpha, batch, (cudaStream_t)stream);
}
/* w_call_circshift, common.cu:375-395.  inplace: the result is in d_image (d_image2 is scratch), else in d_image2. */
int pdwt_call_circshift(float* d_image, float* d_image2, pdwt_w_info w, int sr, int sc, int inplace, int batch, void* stream)
{
    if (!d_image || !d_image2 || w.Nr < 1 || w.Nc < 1 || batch < 1) return PDWT_ERR_ARG;
    const int Nr = w.Nr, Nc = w.Nc;
    if (sr < 0) sr += Nr;
    if (sc < 0) sc += Nc;
    sr = ((sr % Nr) + Nr) % Nr;
    sc = ((sc % Nc) + Nc) % Nc;
    if (w.ndims == 1) sr = 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t plane = (size_t)Nr * Nc;
    if (inplace) {
        PDWT_CUDA(cudaMemcpyAsync(d_image2, d_image, sizeof(float) * plane * batch, cudaMemcpyDeviceToDevice, s));
        return e_circshift(d_image2, d_image, plane, Nr, Nc, sr, sc, batch, s);
    }
    return e_circshift(d_image, d_image2, plane, Nr, Nc, sr, sc, batch, s);
}

static int norm_entry(float** d_coeffs, pdwt_w_info w, int batch, float* out, void* stream, int mode)
{
    if (batch < 1) return PDWT_ERR_ARG;
    double *d = nullptr, *h = nullptr;
    const size_t bytes = sizeof(double) * (size_t)batch * kMaxSeg;
    PDWT_CUDA(cudaMalloc(&d, bytes));
    h = (double*)malloc(bytes);
    int rc = h ? norm_impl(d_coeffs, w, batch, mode, out, (cudaStream_t)stream, d, h) : PDWT_ERR_ALLOC;
    free(h);
    cudaFree(d);
    return rc;
}
int pdwt_norm1(float** d_coeffs, pdwt_w_info winfos, int batch, float* out, void* stream)
{
    return norm_entry(d_coeffs, winfos, batch, out, stream, 0);
}
int pdwt_norm2sq(float** d_coeffs, pdwt_w_info winfos, int batch, float* out, void* stream)
{
    return norm_entry(d_coeffs, winfos, batch, out, stream, 1);
}

}  // extern "C"
