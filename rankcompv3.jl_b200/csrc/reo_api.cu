// reo_api.cu -- the C ABI of libreo_cuda.so (include/reo.h) and the host orchestration of the
// identify_degs loop (reference: src/RankCompV3.jl:339-438).  Host code only decides launches and
// reads back three counters per evaluation; every number in the result is computed on the device.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dlfcn.h>
#include <sched.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <thread>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "reo_internal.cuh"

namespace {

std::mutex g_err_mu;
std::string g_create_err;

template <typename T>
struct DBuf {  // grow-only device buffer
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// NCCL is resolved at run time (dlopen) so that the single-GPU path has no link-time dependency and a
// process that already carries a libnccl.so.2 (e.g. PyTorch's) shares it.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define REO_NCCL_SYM(field, name)                                              \
        field = reinterpret_cast<decltype(field)>(dlsym(lib, name));           \
        if (!field) { err = std::string("libnccl lacks ") + name; lib = nullptr; return false; }
        REO_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        REO_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        REO_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        REO_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        REO_NCCL_SYM(AllGather, "ncclAllGather")
        REO_NCCL_SYM(AllReduce, "ncclAllReduce")
        REO_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef REO_NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

}  // namespace

// A few host threads that copy slices of a pageable input matrix into pinned bounce buffers (one memcpy thread
// cannot keep up with PCIe; see do_stage).
class CopyPool {
public:
    explicit CopyPool(int n) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return (int)th_.size(); }
    // runs f(part, nparts) on every thread and returns when all are done
    void run(const std::function<void(int, int)>& f) {
        std::unique_lock<std::mutex> g(mu_);
        job_ = &f; pending_ = (int)th_.size(); ++gen_;
        cv_.notify_all();
        done_.wait(g, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    void loop(int i) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)>* f;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_; f = job_;
            }
            (*f)(i, (int)th_.size());
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)>* job_ = nullptr;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

struct LevelLog {   // what the reference prints per level k (src:418-420, 432-435)
    int32_t iters_done = 0, converged = 0;
    std::vector<int32_t> n_deg, n_ref;
};

#define REO_NBOUNCE 4
struct ReoDev {
    int dev = 0;
    int num_sms = 148;
    cudaStream_t st = nullptr;
    cudaStream_t st_copy = nullptr;      // host->device copies of the matrix, overlapped with ranking on `st`
    std::vector<cudaEvent_t> copy_ev;    // one per in-flight chunk
    cudaEvent_t ev[8] = {};
    ReoStaged S;
    DBuf<uint8_t> raw, raw2, pb, sub;
    DBuf<uint16_t> ranks;
    DBuf<uint32_t> planes, panel, k1_send, k1_gather;
    DBuf<uint8_t> word_np;
    DBuf<int32_t> slot_of_sample, sample_of_slot, word_order, iota, col_gene, changed_gene, table, perm, perm2, counts,
        fblist, widelist, small_i, stage_lists, table_red, table_all, list_gene;
    DBuf<int8_t> changed_sign, updown, list_sign;
    DBuf<uint8_t> mask_a, mask_b;
    DBuf<double> result, sorted, sorted_p, se, small_d, std_ws;
    DBuf<unsigned int> counter;
    DBuf<int> flags;
    DBuf<unsigned long long> fb_keys;
    DBuf<uint32_t> fb_rank;
    DBuf<long long> small_ll;
    ReoSortWs sortws;
    int32_t* h_counts = nullptr;  // pinned
    uint8_t* h_out = nullptr;     // pinned staging for results (grow-only)
    size_t h_out_cap = 0;
    uint8_t* bounce[REO_NBOUNCE] = {};   // pinned bounce buffers for pageable input (lazily allocated)
    size_t bounce_cap = 0;
    cudaEvent_t bounce_ev[REO_NBOUNCE] = {};
    CopyPool* pool = nullptr;
    bool early_pending = false;   // a copy of result columns 2..14 is in flight on st_copy (ev[7] marks its end)
    int32_t* table_cur = nullptr; // tables the statistics read: `table` (one rank) or `table_red` (sum over ranks)
    int64_t list_cap = 0;         // entries of the gene lists (col_gene, changed_*, list_*, iota)
    std::vector<cudaEvent_t> pev; // event pairs bracketing every pair-kernel launch of the current call
    int n_pev = 0;
    int64_t iota_r = -1;
    ncclComm_t comm = nullptr;    // set by reo_comm_init_rank / multi-device reo_create
    // the per-evaluation statistics sequence (K3..K6) as a CUDA graph, one per mask ping-pong direction
    cudaGraphExec_t eval_exec[2] = {nullptr, nullptr};
    std::vector<const void*> eval_key;
    int64_t eval_r = -1;
    double eval_pd = 0.0, eval_qd = 0.0;
};

struct reo_handle_s {
    std::vector<ReoDev> devs;
    uint64_t seed = 0;
    std::string err;
    int rank = 0, world = 1;
    reo_allgather_fn ag_fn = nullptr;
    void* ag_ctx = nullptr;
    // launch accounting of the current call
    int pair_launches = 0, kernel_launches = 0;
    int64_t compares = 0;
    // single-process multi-GPU: one rank handle per device, driven by one host thread each
    std::vector<reo_handle_s*> subs;
    std::vector<LevelLog> logs;   // per level k of the last reo_identify_degs
    int64_t ordered_triples = 0;  // rows x columns x samples the evaluated pairs stand for (W_ord of SURVEY 8d)
    bool no_narrow = false;       // do_stage retry: copy a pageable matrix raw (it turned out to hold non-integral values)
};

// reo_host.cpp: narrowing copies of the pageable-input pipeline (1 = every value was an integer in 0..65535)
extern "C" int reo_host_narrow_i64(const int64_t* src, uint16_t* dst, size_t n);
extern "C" int reo_host_narrow_i32(const int32_t* src, uint16_t* dst, size_t n);
extern "C" int reo_host_narrow_f64(const double* src, uint16_t* dst, size_t n);
extern "C" int reo_host_narrow_f32(const float* src, uint16_t* dst, size_t n);

namespace {

int host_narrow(int dtype, const void* src, uint16_t* dst, size_t n) {
    switch (dtype) {
        case REO_I64: return reo_host_narrow_i64((const int64_t*)src, dst, n);
        case REO_I32: return reo_host_narrow_i32((const int32_t*)src, dst, n);
        case REO_F64: return reo_host_narrow_f64((const double*)src, dst, n);
        case REO_F32: return reo_host_narrow_f32((const float*)src, dst, n);
        default: return 0;
    }
}

int fail(reo_handle_t h, int code, const std::string& msg) {
    if (h) h->err = msg;
    else { std::lock_guard<std::mutex> g(g_err_mu); g_create_err = msg; }
    return code;
}
int fail_cuda(reo_handle_t h, cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return fail(h, e == cudaErrorMemoryAllocation ? REO_ERR_OOM : REO_ERR_CUDA, m);
}
#define CK(call)                                                     \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) return fail_cuda(h, _e, #call);       \
    } while (0)

// REO_DEBUG=1: synchronise after every kernel launch and name the kernel that faulted
bool debug_sync() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("REO_DEBUG"); v = (e && e[0] && e[0] != '0') ? 1 : 0; }
    return v == 1;
}
#define CKL(call)                                                                  \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e == cudaSuccess && debug_sync()) _e = cudaDeviceSynchronize();       \
        if (_e != cudaSuccess) return fail_cuda(h, _e, #call);                     \
    } while (0)

size_t dtype_size(int dtype) {
    switch (dtype) {
        case REO_I64: case REO_F64: return 8;
        case REO_I32: case REO_F32: return 4;
        default: return 0;
    }
}

// ---- get_major_reo_lower_count, src:81-92 (host; integer result) --------------------------------
long double log_pmf_half(int n, int k) {
    return lgammal((long double)n + 1) - lgammal((long double)k + 1) - lgammal((long double)(n - k) + 1)
           - (long double)n * logl(2.0L);
}
long double cdf_half(int n, int x) {  // P[X <= x], X ~ Binomial(n, 1/2); terms decay going down from x <= n/2
    if (x < 0) return 0.0L;
    long double term = 1.0L, sum = 0.0L;
    for (int k = x; k >= 0; --k) {
        sum += term;
        if (term < 1e-25L * sum) break;
        term *= (long double)k / (long double)(n - k + 1);
    }
    return expl(log_pmf_half(n, x)) * sum;
}
long double two_sided_binom_p(int n, int x) {  // HypothesisTests.pvalue(Binomial(n), x), tail = :both
    long double lo = cdf_half(n, x), hi = 1.0L - cdf_half(n, x - 1);
    long double p = 2.0L * std::min(lo, hi);
    return p > 1.0L ? 1.0L : p;
}

struct LevelPlan {  // the two-group view of level k
    int WA = 0, nA = 0, nB = 0, padA = 0, padB = 0, thrA = 0, thrB = 0;
    int mixed = 0;
    uint32_t maskA = 0, maskB = 0;
    int segA0 = 0, mixedW = 0, segB0 = 0, segB0len = 0, segB1 = 0;
    std::vector<int32_t> word_order;
};

LevelPlan make_plan(const ReoStaged& S, int k, const int32_t* thresholds, double pval_reo) {
    LevelPlan P;
    P.nA = S.lev_n[k];
    P.nB = (int)S.c - P.nA;
    for (int w = 0; w < S.lev_words[k]; ++w) P.word_order.push_back(S.lev_word0[k] + w);
    P.WA = S.lev_words[k];
    if (S.mixed_word >= 0) {
        // two levels whose tails share one word: [full words of A][mixed][full words of B], no pad slots counted
        P.word_order.push_back(S.mixed_word);
        P.mixed = 1;
        const uint32_t m0 = (S.mixed_rem[0] >= 32) ? 0xffffffffu : ((1u << S.mixed_rem[0]) - 1u);
        const uint32_t m1 = ((S.mixed_rem[1] >= 32) ? 0xffffffffu : ((1u << S.mixed_rem[1]) - 1u)) << S.mixed_rem[0];
        P.maskA = (k == 0) ? m0 : m1;
        P.maskB = (k == 0) ? m1 : m0;
    }
    for (int g = 0; g < S.gnum; ++g) {
        if (g == k) continue;
        for (int w = 0; w < S.lev_words[g]; ++w) P.word_order.push_back(S.lev_word0[g] + w);
        if (S.mixed_word < 0) P.padB += S.lev_words[g] * 32 - S.lev_n[g];
    }
    if (S.mixed_word < 0) P.padA = S.lev_words[k] * 32 - P.nA;
    // the same order as arithmetic segments (levels are staged in order, so "all other levels" is the run
    // before level k followed by the run after it)
    P.segA0 = S.lev_word0[k];
    P.mixedW = S.mixed_word < 0 ? 0 : S.mixed_word;
    {
        const int after = S.lev_word0[k] + S.lev_words[k] + (S.mixed_word >= 0 && k == 0 ? 1 : 0);
        if (S.mixed_word >= 0) {           // two levels: the other level is one run
            P.segB0 = S.lev_word0[1 - k]; P.segB0len = S.lev_words[1 - k]; P.segB1 = 0;
        } else {
            P.segB0 = 0; P.segB0len = S.lev_word0[k]; P.segB1 = after;
        }
    }
    if (thresholds) { P.thrA = thresholds[0 + 2 * k]; P.thrB = thresholds[1 + 2 * k]; }
    else { P.thrA = reo_threshold(P.nA, pval_reo); P.thrB = reo_threshold(P.nB, pval_reo); }
    return P;
}

int ensure_std_ws(reo_handle_t h, ReoDev& D) {
    if (D.std_ws.p) return REO_OK;
    CK(D.std_ws.ensure(2 * 256));   // leaf sums of the two passes (reo_launch_trimmed_std)
    return REO_OK;
}

int ensure_h_out(reo_handle_t h, ReoDev& D, size_t bytes) {
    if (bytes <= D.h_out_cap) return REO_OK;
    if (D.h_out) cudaFreeHost(D.h_out);
    D.h_out = nullptr; D.h_out_cap = 0;
    CK(cudaMallocHost((void**)&D.h_out, bytes));
    D.h_out_cap = bytes;
    return REO_OK;
}

int upload_plan(reo_handle_t h, ReoDev& D, const LevelPlan& P) {
    // src:85: findfirst(pval .> pval_reo) returns nothing when no count can reach significance -- the reference throws there
    if (P.thrA < 0 || P.thrB < 0)
        return fail(h, REO_ERR_ARG, "no stable-REO threshold for this pval_reo (src:85 finds no count): pval_reo must be below 1");
    CK(D.word_order.ensure(P.word_order.size()));
    CK(cudaMemcpyAsync(D.word_order.p, P.word_order.data(), P.word_order.size() * 4, cudaMemcpyHostToDevice, D.st));
    return REO_OK;
}

// ---- K1 -----------------------------------------------------------------------------------------
int do_stage(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld, const int32_t* group_id,
             int32_t gnum, uint32_t flags) {
    ReoDev& D = h->devs[0];
    ReoStaged& S = D.S;
    S.valid = false; S.flt = false; S.flt_f32 = false;
    const size_t es = dtype_size(dtype);
    if (!data || !group_id || es == 0 || r < 1 || c < 1 || ld < r) return fail(h, REO_ERR_ARG, "reo_stage: bad argument");
    if (gnum < 2) return fail(h, REO_ERR_DIM, "Only 1 level in 'group', at least 2 levels!");
    if (r > 131072) return fail(h, REO_ERR_UNSUPPORTED, "more than 131072 genes");
    const int rank_bytes = r > 65535 ? 4 : 2;   // dense ranks: u16 up to 65535 genes, u32 beyond
    if (c > (int64_t)1 << 24) return fail(h, REO_ERR_UNSUPPORTED, "more than 2^24 samples");
    std::vector<int> lev_n(gnum, 0);
    for (int64_t s = 0; s < c; ++s) {
        if (group_id[s] < 0 || group_id[s] >= gnum) return fail(h, REO_ERR_DIM, "group_id out of range");
        lev_n[group_id[s]]++;
    }
    for (int g = 0; g < gnum; ++g)
        if (lev_n[g] == 0) return fail(h, REO_ERR_DIM, "empty group level");
    CK(cudaSetDevice(D.dev));
    S.r = r; S.c = c; S.gnum = gnum;
    S.NT = (int)((r + REO_TILE - 1) / REO_TILE);
    S.lev_n = lev_n;
    S.lev_words.assign(gnum, 0); S.lev_word0.assign(gnum, 0);
    int W = 0;
    S.mixed_word = -1; S.mixed_rem[0] = S.mixed_rem[1] = 0;
    const int rem0 = lev_n[0] % 32, rem1 = gnum == 2 ? lev_n[1] % 32 : 0;
    if (gnum == 2 && rem0 > 0 && rem1 > 0 && rem0 + rem1 <= 32) {
        // [full words of level 0][one word with both tails][full words of level 1]
        S.lev_word0[0] = 0; S.lev_words[0] = lev_n[0] / 32;
        S.mixed_word = S.lev_words[0];
        S.lev_word0[1] = S.mixed_word + 1; S.lev_words[1] = lev_n[1] / 32;
        S.mixed_rem[0] = rem0; S.mixed_rem[1] = rem1;
        W = S.lev_words[0] + 1 + S.lev_words[1];
    } else {
        for (int g = 0; g < gnum; ++g) { S.lev_word0[g] = W; S.lev_words[g] = (lev_n[g] + 31) / 32; W += S.lev_words[g]; }
    }
    S.W = W;
    const int64_t nslots = (int64_t)W * 32;
    const int64_t rpad = (int64_t)S.NT * REO_TILE;
    std::vector<int32_t> slot_of_sample(c), sample_of_slot(nslots, -1);
    {
        std::vector<int> fill(gnum, 0);
        for (int64_t s = 0; s < c; ++s) {
            const int g = group_id[s];
            const int f = fill[g]++;
            int32_t slot = S.lev_word0[g] * 32 + f;
            if (S.mixed_word >= 0 && f >= S.lev_words[g] * 32)  // tail sample -> the shared word
                slot = S.mixed_word * 32 + (g == 0 ? 0 : S.mixed_rem[0]) + (f - S.lev_words[g] * 32);
            slot_of_sample[s] = slot; sample_of_slot[slot] = (int32_t)s;
        }
    }
    CK(D.slot_of_sample.ensure(c));
    CK(D.sample_of_slot.ensure(nslots));
    CK(cudaMemcpyAsync(D.slot_of_sample.p, slot_of_sample.data(), c * 4, cudaMemcpyHostToDevice, D.st));
    CK(cudaMemcpyAsync(D.sample_of_slot.p, sample_of_slot.data(), nslots * 4, cudaMemcpyHostToDevice, D.st));

    // K1 sharding (NCCL communicator present): this rank copies, ranks and bit-slices only the sample words
    // [w_lo, w_hi); the staged blocks are all-gathered over NVLink and re-ordered into planes[t][w].
    // (small device-resident matrices are staged redundantly: two extra collectives cost more than ranking 200
    // columns; for host input the share of the host->device copy saved pays much earlier)
    int64_t shard_min = (flags & REO_DATA_ON_DEVICE) ? (int64_t)1 << 24 : (int64_t)1 << 20;
    if (const char* e = getenv("REO_K1_SHARD_MIN")) shard_min = atoll(e);   // tests force the sharded path on small inputs
    const bool shard = (h->world > 1 && D.comm != nullptr && r * c >= shard_min);
    const int wq = shard ? (W + h->world - 1) / h->world : W;
    const int w_lo = shard ? std::min(W, h->rank * wq) : 0;
    const int w_hi = shard ? std::min(W, (h->rank + 1) * wq) : W;
    std::vector<int32_t> my_samples;   // original sample indices staged by this rank, ascending
    for (int64_t s = 0; s < c; ++s) {
        const int w = slot_of_sample[s] / 32;
        if (w >= w_lo && w < w_hi) my_samples.push_back((int32_t)s);
    }
    const int64_t nmy = (int64_t)my_samples.size();
    const bool on_dev = (flags & REO_DATA_ON_DEVICE) != 0;
    // list position j -> column of the device matrix (compact copy for host input, the caller's matrix otherwise)
    std::vector<int32_t> src_col(std::max<int64_t>(nmy, 1)), col_of_sample(c, -1);
    for (int64_t j = 0; j < nmy; ++j) { src_col[j] = on_dev ? my_samples[j] : (int32_t)j; col_of_sample[my_samples[j]] = src_col[j]; }
    CK(D.stage_lists.ensure((size_t)2 * std::max<int64_t>(nmy, 1) + c));
    int32_t* d_src_col = D.stage_lists.p; int32_t* d_sample_id = d_src_col + std::max<int64_t>(nmy, 1);
    int32_t* d_col_of_sample = d_sample_id + std::max<int64_t>(nmy, 1);
    if (nmy > 0) {
        CK(cudaMemcpyAsync(d_src_col, src_col.data(), nmy * 4, cudaMemcpyHostToDevice, D.st));
        CK(cudaMemcpyAsync(d_sample_id, my_samples.data(), nmy * 4, cudaMemcpyHostToDevice, D.st));
    }
    CK(cudaMemcpyAsync(d_col_of_sample, col_of_sample.data(), c * 4, cudaMemcpyHostToDevice, D.st));

    CK(D.ranks.ensure((size_t)nslots * rpad * (rank_bytes / 2)));
    // no clearing: bitplanes_kernel masks pad slots and genes >= r itself
    CK(D.flags.ensure(16));
    CK(cudaMemsetAsync(D.flags.p, 0, 16 * sizeof(int), D.st));
    CK(D.fblist.ensure(std::max<int64_t>(nmy, 1)));
    CK(D.widelist.ensure(std::max<int64_t>(nmy, 1)));

    // raw matrix: already on the device, or copied in runs of consecutive columns that overlap with ranking
    const uint8_t* dev_data = (const uint8_t*)data;
    int64_t dev_ld = ld;
    if (!on_dev) {
        CK(D.raw.ensure((size_t)r * std::max<int64_t>(nmy, 1) * es));
        dev_data = D.raw.p; dev_ld = r;
    }
    // columns per rank launch: one CTA per column and one CTA per SM, so whole waves of SMs; host input is copied
    // in chunks of >= 16 MB so that the copy of chunk n+1 overlaps the ranking of chunk n
    int64_t chunk = nmy;
    if (!on_dev) {
        chunk = std::max<int64_t>(1, (int64_t)(16u << 20) / (int64_t)(r * es));
        chunk = std::max<int64_t>(D.num_sms, chunk / D.num_sms * D.num_sms);
        if (const char* e = getenv("REO_CHUNK_COLS")) chunk = std::max<int64_t>(1, atoll(e));
    }
    chunk = std::max<int64_t>(chunk, 1);
    bool pageable = false;
    if (!on_dev) {
        cudaPointerAttributes pa;
        const cudaError_t pe = cudaPointerGetAttributes(&pa, data);
        if (pe != cudaSuccess) { cudaGetLastError(); pageable = true; }
        else pageable = (pa.type == cudaMemoryTypeUnregistered);
        static const bool no_bounce = getenv("REO_NO_BOUNCE") != nullptr;
        if (no_bounce) pageable = false;
    }
    // Host count matrices are narrowed to u16 while they pass through the bounce buffers (reo_host.cpp) -- pageable
    // ones, which need the bounce anyway, and page-locked ones too as long as their chunks do narrow (the copy threads
    // read host memory faster than PCIe carries 8-byte elements); a page-locked chunk that does not is sent by DMA
    // straight from the caller's memory, and so are all later ones.
    // the caller blocks in this call: all host cores copy (B200 host, 16 cores: 8 threads narrow 73 GB/s of input, 16
    // threads 86), shared evenly between the ranks of one box
    int copy_threads = 16;
    if (const char* e = getenv("REO_COPY_THREADS")) copy_threads = atoi(e);
    {
        int hc = (int)std::thread::hardware_concurrency();
        cpu_set_t cs;   // a cpuset may leave this process fewer cores than the machine has
        if (sched_getaffinity(0, sizeof(cs), &cs) == 0 && CPU_COUNT(&cs) > 0) hc = hc > 0 ? std::min(hc, CPU_COUNT(&cs)) : CPU_COUNT(&cs);
        if (hc > 0) copy_threads = std::min(copy_threads, std::max(1, hc / std::max(1, h->world)));
        copy_threads = std::max(1, copy_threads);
    }
    static const bool no_narrow_env = getenv("REO_NO_NARROW") != nullptr;
    bool try_narrow = !on_dev && !no_narrow_env && !h->no_narrow;
    {
        static const bool no_bounce = getenv("REO_NO_BOUNCE") != nullptr;
        if (no_bounce) try_narrow = false;
        // page-locked input: DMA needs no CPU at all, narrowing only wins with enough threads to outrun PCIe (measured:
        // 16 threads 270 ms against 287 ms by plain DMA for the 4.8 GB matrix, 8 threads 325 ms; two ranks with 8 threads
        // each, sharing their chunks between DMA and narrowing: 145.5 ms against 139.3 ms by plain DMA)
        if (!pageable && copy_threads < 12) try_narrow = false;
    }
    if (pageable || try_narrow) {
        const size_t need = (size_t)chunk * r * es;
        if (need > D.bounce_cap) {
            for (auto& bptr : D.bounce) { if (bptr) cudaFreeHost(bptr); bptr = nullptr; }
            D.bounce_cap = 0;
            for (auto& bptr : D.bounce) CK(cudaMallocHost((void**)&bptr, need));
            D.bounce_cap = need;
        }
        for (auto& e : D.bounce_ev) if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        if (!D.pool) D.pool = new CopyPool(copy_threads);
    }
    bool narrowed_any = false;
    if (!on_dev) {   // the copy stream must not overtake work still queued on `st` that reads/frees `raw`
        CK(cudaEventRecord(D.ev[5], D.st));
        CK(cudaStreamWaitEvent(D.st_copy, D.ev[5], 0));
    }
    size_t n_chunk = 0, n_narrow = 0;
    bool bounce_used[REO_NBOUNCE] = {};
    static const int raw_every = getenv("REO_RAW_EVERY") ? atoi(getenv("REO_RAW_EVERY")) : 3;
    const bool small_first = c >= 1024;   // heuristic only: either order of tiers gives the same ranks
    for (int64_t j0 = 0; j0 < nmy;) {
        int64_t n = 1;   // run of consecutive original columns, at most `chunk` long
        while (j0 + n < nmy && n < chunk && my_samples[j0 + n] == my_samples[j0] + n) ++n;
        bool chunk_narrow = false;
        if (!on_dev) {
            // copy on the copy stream, rank on the compute stream as soon as this chunk has landed
            const uint8_t* src = (const uint8_t*)data + (size_t)my_samples[j0] * ld * es;
            if (pageable || try_narrow) {
                // a few host threads fill a pinned bounce buffer while the DMA engine drains the previous ones (for a
                // pageable matrix the driver's own path is several times slower than PCIe)
                const int b = (int)(n_chunk % REO_NBOUNCE);
                if (bounce_used[b]) CK(cudaEventSynchronize(D.bounce_ev[b]));   // its previous chunk has left the buffer
                uint8_t* dstb = D.bounce[b];
                const size_t colb = (size_t)r * es, ldb = (size_t)ld * es;
                // page-locked input: every raw_every-th chunk goes by plain DMA while the copy threads narrow the
                // others -- PCIe and the host cores then work side by side (the DMA of a raw chunk takes about as long
                // as narrowing two)
                const bool dma_turn = !pageable && raw_every > 0 && (n_chunk % (size_t)raw_every) == (size_t)raw_every - 1;
                if (try_narrow && !dma_turn) {
                    std::atomic<int> all_small{1};
                    D.pool->run([&](int part, int nparts) {
                        if (ldb == colb) {   // contiguous chunk: equal element ranges, whatever the number of columns
                            const int64_t tot = n * r;
                            const int64_t e0 = tot * part / nparts / 16 * 16;
                            const int64_t e1 = part + 1 == nparts ? tot : tot * (part + 1) / nparts / 16 * 16;
                            if (e1 > e0 && !host_narrow(dtype, src + e0 * es, (uint16_t*)dstb + e0, (size_t)(e1 - e0))) all_small.store(0);
                            return;
                        }
                        const int64_t c0 = n * part / nparts, c1 = n * (part + 1) / nparts;
                        for (int64_t q = c0; q < c1 && all_small.load(std::memory_order_relaxed); ++q)
                            if (!host_narrow(dtype, src + q * ldb, (uint16_t*)dstb + q * r, (size_t)r)) all_small.store(0);
                    });
                    chunk_narrow = all_small.load() != 0;
                    if (!chunk_narrow && !pageable) try_narrow = false;   // page-locked: plain DMA from here on
                }
                if (!chunk_narrow && pageable)
                    D.pool->run([&](int part, int nparts) {
                        const int64_t c0 = n * part / nparts, c1 = n * (part + 1) / nparts;
                        if (ldb == colb) { if (c1 > c0) memcpy(dstb + c0 * colb, src + c0 * colb, (size_t)(c1 - c0) * colb); }
                        else for (int64_t q = c0; q < c1; ++q) memcpy(dstb + q * colb, src + q * ldb, colb);
                    });
                narrowed_any |= chunk_narrow;
                n_narrow += chunk_narrow ? 1 : 0;
                if (chunk_narrow || pageable) {
                    // a narrowed chunk lands at the start of the chunk's own place in `raw`
                    CK(cudaMemcpyAsync(D.raw.p + (size_t)j0 * r * es, dstb, (size_t)n * (chunk_narrow ? (size_t)r * 2 : colb),
                                       cudaMemcpyHostToDevice, D.st_copy));
                    CK(cudaEventRecord(D.bounce_ev[b], D.st_copy));
                    bounce_used[b] = true;
                } else {
                    CK(cudaMemcpy2DAsync(D.raw.p + (size_t)j0 * r * es, colb, src, ldb, colb, (size_t)n, cudaMemcpyHostToDevice, D.st_copy));
                }
            } else {
                CK(cudaMemcpy2DAsync(D.raw.p + (size_t)j0 * r * es, (size_t)r * es, src, (size_t)ld * es, (size_t)r * es,
                                     (size_t)n, cudaMemcpyHostToDevice, D.st_copy));
            }
            if (n_chunk >= D.copy_ev.size()) {
                cudaEvent_t ev;
                CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                D.copy_ev.push_back(ev);
            }
            CK(cudaEventRecord(D.copy_ev[n_chunk], D.st_copy));
            CK(cudaStreamWaitEvent(D.st, D.copy_ev[n_chunk], 0));
            ++n_chunk;
        }
        // tier 0: small presence bitmap (value range < 65536), 8 CTAs per SM; wider columns are listed for tier 1.
        // Bulk inputs (few samples, counts up to ~1e6) would all overflow it, so they start at tier 1 directly.
        // (a narrowed chunk: u16 elements; the base is shifted so that column j of the list is at base + j * r elements)
        const uint8_t* cdata = chunk_narrow ? D.raw.p + (size_t)j0 * r * (es - 2) : dev_data;
        const int cdtype = chunk_narrow ? REO_U16_STAGED : dtype;
        if (small_first)
            CKL(reo_launch_rank_columns(cdata, cdtype, r, dev_ld, j0, (int)n, nullptr, 0, d_src_col, d_sample_id,
                                        D.slot_of_sample.p, D.ranks.p, rank_bytes, rpad, D.flags.p + 2, D.flags.p,
                                        D.flags.p + 3, D.widelist.p, D.st));
        else
            CKL(reo_launch_rank_columns(cdata, cdtype, r, dev_ld, j0, (int)n, nullptr, 1, d_src_col, d_sample_id,
                                        D.slot_of_sample.p, D.ranks.p, rank_bytes, rpad, D.flags.p + 2, D.flags.p,
                                        D.flags.p + 1, D.fblist.p, D.st));
        h->kernel_launches++;
        j0 += n;
    }
    if (!on_dev && getenv("REO_TIMING"))
        fprintf(stderr, "[reo timing] rank %d %s host input: %zu copy chunks, %zu narrowed to u16\n", h->rank,
                pageable ? "pageable" : "page-locked", n_chunk, n_narrow);
    if (narrowed_any) {   // flags[4]: this rank narrowed a chunk (the retry below must be taken by every rank or none)
        D.h_counts[8] = 1;
        CK(cudaMemcpyAsync(D.flags.p + 4, D.h_counts + 8, sizeof(int), cudaMemcpyHostToDevice, D.st));
    }
    CK(cudaMemcpyAsync(D.h_counts, D.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    if (D.h_counts[3] > 0 && !D.h_counts[0]) {
        // tier 1: 1.3 M-value bitmap, one CTA per SM; still wider columns are listed for the sort-based fallback
        CKL(reo_launch_rank_columns(dev_data, dtype, r, dev_ld, 0, D.h_counts[3], D.widelist.p, 1, d_src_col, d_sample_id,
                                    D.slot_of_sample.p, D.ranks.p, rank_bytes, rpad, D.flags.p + 2, D.flags.p,
                                    D.flags.p + 1, D.fblist.p, D.st));
        h->kernel_launches++;
        CK(cudaMemcpyAsync(D.h_counts, D.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, D.st));
        CK(cudaStreamSynchronize(D.st));
    }
    const int nfb = D.h_counts[1];
    if (nfb > 0 && !D.h_counts[0]) {   // columns whose value range exceeds the bitmap: sort-based dense rank
        int64_t rp2 = 1;
        while (rp2 < r) rp2 <<= 1;
        const int batch = std::min(nfb, 2 * D.num_sms);
        CK(D.fb_keys.ensure((size_t)batch * rp2));
        CK(D.fb_rank.ensure((size_t)batch * rp2));
        for (int b0 = 0; b0 < nfb; b0 += batch) {
            const int nb = std::min(batch, nfb - b0);
            CKL(reo_launch_rank_fallback(dev_data, dtype, r, dev_ld, D.fblist.p + b0, nb, d_src_col, d_sample_id,
                                         D.slot_of_sample.p, D.ranks.p, rank_bytes, rpad, D.flags.p + 2, D.fb_keys.p,
                                         D.fb_rank.p, rp2, D.st));
            h->kernel_launches++;
        }
    }
    const bool more_ranked = (D.h_counts[3] > 0 || nfb > 0) && !D.h_counts[0];   // a later tier ran after the read-back
    if (shard) {   // non-integrality and the largest dense rank are properties of the whole matrix
        const ncclResult_t nr = g_nccl.AllReduce(D.flags.p, D.flags.p + 8, 5, ncclInt32, ncclMax, D.comm, D.st);
        if (nr != ncclSuccess) return fail(h, REO_ERR_COMM, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(nr));
        CK(cudaMemcpyAsync(D.h_counts, D.flags.p + 8, 5 * sizeof(int), cudaMemcpyDeviceToHost, D.st));
        CK(cudaStreamSynchronize(D.st));
        narrowed_any = D.h_counts[4] != 0;
    } else if (more_ranked) {
        CK(cudaMemcpyAsync(D.h_counts, D.flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, D.st));
        CK(cudaStreamSynchronize(D.st));
    }
    if (D.h_counts[0] && narrowed_any) {
        // some chunks hold non-integral values while others were narrowed to u16: the float path needs the raw matrix
        // on the device, so stage once more without narrowing (every rank takes this branch or none does)
        h->no_narrow = true;
        const int rc2 = do_stage(h, data, dtype, r, c, ld, group_id, gnum, flags);
        h->no_narrow = false;
        return rc2;
    }
    if (D.h_counts[0]) {
        // non-integral values: the 0.1 tie band of is_greater (src:72) is not transitive, so ranks cannot be
        // used -- stage the raw values as FP64 and let the pair kernel compare them directly
        if (dtype != REO_F64 && dtype != REO_F32) return fail(h, REO_ERR_ARG, "non-integral values in an integer matrix");
        S.flt = true; S.flt_f32 = (dtype == REO_F32); S.B = 0; S.NP = 1;
    } else {
        const int distinct = std::max(D.h_counts[2], 1);
        int B = 1;
        while ((1 << B) < distinct) ++B;
        if (B > REO_MAX_BITS) return fail(h, REO_ERR_UNSUPPORTED, "rank needs more than 20 bits");
        S.B = B; S.NP = B + 1;
    }
    CK(D.planes.ensure((size_t)S.NT * S.tile_stride()));
    S.planes = D.planes.p;
    const size_t wb = S.word_stride();   // words of one (tile, sample word) block
    uint32_t* stage_out = S.planes;
    if (shard) {
        CK(D.k1_send.ensure((size_t)S.NT * wq * wb));
        CK(D.k1_gather.ensure((size_t)h->world * S.NT * wq * wb));
        stage_out = D.k1_send.p;
    }
    if (S.flt)
        CKL(reo_launch_fstage(dev_data, dtype, r, dev_ld, D.sample_of_slot.p, d_col_of_sample, S.NT, w_lo, w_hi - w_lo,
                              shard ? wq : W, (uint32_t)h->seed, (uint32_t)(h->seed >> 32), stage_out, D.st));
    else
        CKL(reo_launch_bitplanes(D.ranks.p, rank_bytes, rpad, r, D.sample_of_slot.p, S.NT, w_lo, w_hi - w_lo,
                                 shard ? wq : W, S.NP, (uint32_t)h->seed, (uint32_t)(h->seed >> 32), stage_out, D.st));
    h->kernel_launches++;
    if (shard) {
        const size_t cnt = (size_t)S.NT * wq * wb;
        const ncclResult_t nr = g_nccl.AllGather(D.k1_send.p, D.k1_gather.p, cnt, ncclUint32, D.comm, D.st);
        if (nr != ncclSuccess) return fail(h, REO_ERR_COMM, std::string("ncclAllGather(planes): ") + g_nccl.GetErrorString(nr));
        CKL(reo_launch_unshard_planes(D.k1_gather.p, S.planes, S.NT, W, wq, (int)wb, D.st));
        h->kernel_launches += 2;
    }
    S.word_np = nullptr;
    static const bool no_skip = getenv("REO_NO_PLANE_SKIP") != nullptr;
    if (!S.flt && !no_skip) {   // which planes each sample word really uses (the pair kernel skips an empty top plane)
        CK(D.word_np.ensure((size_t)W));
        CK(cudaMemsetAsync(D.flags.p + 10, 0, sizeof(int), D.st));
        CKL(reo_launch_word_planes(S.planes, S.NT, W, S.NP, D.word_np.p, D.flags.p + 10, D.st));
        CK(cudaMemcpyAsync(D.h_counts + 12, D.flags.p + 10, sizeof(int), cudaMemcpyDeviceToHost, D.st));
        h->kernel_launches++;
        S.word_np = D.word_np.p;
        S.planes_per_word = -1.0;   // read from h_counts[12] once the stream has been synchronised (stats)
    } else {
        S.planes_per_word = (double)S.NP;
    }
    // gene lists are padded with -1 up to whole T-tile blocks plus a pair of tiles (the pair kernel copies the ids of
    // two column tiles per step); identity list for "all genes are references" (cached while r is unchanged)
    const int64_t list_cap = rpad + (int64_t)(8 + 8 + 2) * REO_TILE;
    D.list_cap = list_cap;
    if (D.iota_r != r) {
        std::vector<int32_t> iota(list_cap, -1);
        for (int64_t i = 0; i < r; ++i) iota[i] = (int32_t)i;
        CK(D.iota.ensure(list_cap));
        CK(cudaMemcpyAsync(D.iota.p, iota.data(), list_cap * 4, cudaMemcpyHostToDevice, D.st));
        CK(cudaStreamSynchronize(D.st));  // iota is a host temporary
        D.iota_r = r;
    }
    CK(D.col_gene.ensure(list_cap));
    CK(D.changed_gene.ensure(list_cap));
    CK(D.changed_sign.ensure(list_cap));
    CK(D.list_gene.ensure(list_cap));
    CK(D.list_sign.ensure(list_cap));
    CK(D.counts.ensure(8));
    CK(D.counter.ensure(1));
    CK(D.mask_a.ensure(r));
    CK(D.mask_b.ensure(r));
    // table: this rank's partial sums (r x 9); with several ranks the statistics read the sum over ranks
    CK(D.table.ensure((size_t)r * 9));
    D.table_cur = D.table.p;
    if (h->world > 1) {
        CK(D.table_red.ensure((size_t)r * 9));
        D.table_cur = D.table_red.p;
        if (!D.comm) CK(D.table_all.ensure((size_t)h->world * r * 9));
    }
    S.valid = true;
    return REO_OK;
}

// ---- K2 launches --------------------------------------------------------------------------------
bool pairs_v1() {   // REO_PAIRS_V1=1: first-generation pair kernel on the rank path too (A/B comparisons)
    static const bool v = [] { const char* e = getenv("REO_PAIRS_V1"); return e && e[0] && e[0] != '0'; }();
    return v;
}
// a column set of n genes is worth the symmetric treatment (permuted panel [C ; N]) from this size on
bool sym_worth(int64_t n, int64_t r) {
    static const int64_t div = getenv("REO_SYM_DIV") ? atoll(getenv("REO_SYM_DIV")) : 16;   // measured: 16 beats 8 by 0.9 % on 30k x 20k
    return n >= 1024 && n * div >= r;
}
// relative cost (pair evaluations) of accumulating n columns into the tables of r genes
double tables_cost(int64_t n, int64_t r, bool flt) {
    if (!flt && !pairs_v1() && (n == r || sym_worth(n, r))) return (double)n * ((double)r - 0.5 * (double)n);
    return (double)n * (double)r;
}

int record_pair_events(reo_handle_t h, ReoDev& D, bool begin) {
    if (begin && (size_t)(2 * D.n_pev + 2) > D.pev.size()) {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        D.pev.push_back(a); D.pev.push_back(b);
    }
    CK(cudaEventRecord(D.pev[2 * D.n_pev + (begin ? 0 : 1)], D.st));
    if (!begin) { D.n_pev++; h->pair_launches++; h->kernel_launches++; }
    return REO_OK;
}

// this rank's share of `total` executed comparisons (the shares of all ranks add up to total)
int64_t rank_share(reo_handle_t h, int64_t total) {
    return total / h->world + (h->rank < (int)(total % h->world) ? 1 : 0);
}

// First-generation kernel (raw-FP64 variant; REO_PAIRS_V1): every row gene of this rank's row-tile shard against the
// `ncols` panel columns listed (ascending, padded with -1) in col_gene_dev.
int launch_tables_v1(reo_handle_t h, ReoDev& D, const LevelPlan& P, const int32_t* col_gene_dev,
                     const int8_t* col_sign_dev, int ncols, bool all_genes) {
    const ReoStaged& S = D.S;
    if (ncols <= 0) return REO_OK;
    const int ntc = (ncols + REO_TILE - 1) / REO_TILE;
    const uint32_t* colp = S.planes;
    if (!all_genes) {
        CK(D.panel.ensure((size_t)ntc * S.tile_stride()));
        if (S.flt) CKL(reo_launch_gather_panel_flt(S.planes, S.W, col_gene_dev, ntc, D.panel.p, D.st));
        else CKL(reo_launch_gather_panel(S.planes, S.W, S.NP, col_gene_dev, ntc, D.panel.p, D.st));
        h->kernel_launches++;
        colp = D.panel.p;
    }
    const int tpr = (S.NT + h->world - 1) / h->world;
    ReoPairParams p;
    memset(&p, 0, sizeof(p));
    p.row_planes = S.planes; p.col_planes = colp; p.col_gene = col_gene_dev; p.col_sign = col_sign_dev;
    p.segA0 = P.segA0; p.mixedW = P.mixedW; p.segB0 = P.segB0; p.segB0len = P.segB0len; p.segB1 = P.segB1;
    p.table = D.table.p; p.counter = D.counter.p;
    p.W = S.W; p.WA = P.WA; p.NP = S.NP; p.r = (int)S.r;
    p.t0 = std::min(S.NT, h->rank * tpr); p.t1 = std::min(S.NT, (h->rank + 1) * tpr);
    p.ntc = ntc;
    const int ntr = p.t1 - p.t0;
    if (ntr <= 0) return REO_OK;
    // Work items = row tile x chunk of column tiles, handed out dynamically.  At least ~16 items per resident CTA;
    // beyond that an item only needs enough sample words (~64) to amortise its prologue and table flush -- with
    // thousands of samples one column tile per item keeps the tail of the launch below 1 %.
    const int want_items = 16 * 3 * D.num_sms;
    int njc = std::max(1, std::min(ntc, (want_items + ntr - 1) / ntr));
    p.jchunk = (ntc + njc - 1) / njc;
    p.jchunk = std::min(p.jchunk, std::max(1, (64 + S.W - 1) / S.W));
    const int cts = reo_pairs_col_tiles_per_step(S.flt);   // whole steps: no half-empty step at the end of a chunk
    p.jchunk = (p.jchunk + cts - 1) / cts * cts;
    p.njchunks = (ntc + p.jchunk - 1) / p.jchunk;
    p.nA = P.nA; p.nB = P.nB; p.padA = P.padA; p.padB = P.padB; p.thrA = P.thrA; p.thrB = P.thrB;
    p.mixed = P.mixed; p.maskA = P.maskA; p.maskB = P.maskB;
    p.flt = S.flt ? (S.flt_f32 ? 2 : 1) : 0;
    CK(cudaMemsetAsync(D.counter.p, 0, sizeof(unsigned int), D.st));
    int rc = record_pair_events(h, D, true);
    if (rc) return rc;
    CKL(reo_launch_pairs(p, D.num_sms, D.st));
    if ((rc = record_pair_events(h, D, false))) return rc;
    const int64_t rows = std::min<int64_t>(S.r, (int64_t)p.t1 * REO_TILE) - (int64_t)p.t0 * REO_TILE;
    h->compares += std::max<int64_t>(rows, 0) * (int64_t)ncols * S.c;
    h->ordered_triples += std::max<int64_t>(rows, 0) * (int64_t)ncols * S.c;
    return REO_OK;
}

// Accumulate sign(j) * e(category(i, j)) into the table of every gene i, for the column set C of `ncols` genes:
//   m_new == nullptr : C = { j : m_old[j] }, all signs +1 (full build for the reference set m_old);
//   m_new != nullptr : C = m_old xor m_new, sign +1 for genes that enter the set and -1 for genes that leave it
//                      (incremental update; mask_diff has left the list in changed_gene / changed_sign).
// Three shapes (v2 kernel): every gene is a column -> symmetric sweep over the staged planes themselves; a large C ->
// permuted panel [C ; N], symmetric C x C plus one-sided N x C; a small C -> one-sided, all genes x gathered C.
int launch_tables(reo_handle_t h, ReoDev& D, const LevelPlan& P, const uint8_t* m_old, const uint8_t* m_new, int ncols) {
    const ReoStaged& S = D.S;
    if (ncols <= 0) return REO_OK;
    const bool all = (m_new == nullptr && ncols == (int)S.r);
    if (S.flt || pairs_v1()) {
        if (all) return launch_tables_v1(h, D, P, D.iota.p, nullptr, ncols, true);
        if (m_new) return launch_tables_v1(h, D, P, D.changed_gene.p, D.changed_sign.p, ncols, false);
        CKL(reo_launch_mask_to_list(m_old, S.r, D.col_gene.p, D.counts.p + 4, D.st));
        h->kernel_launches++;
        return launch_tables_v1(h, D, P, D.col_gene.p, nullptr, ncols, false);
    }
    const int T = reo_pairs2_block_edge(S.W, S.NP);
    ReoPair2Params p;
    memset(&p, 0, sizeof(p));
    p.T = T; p.rank = h->rank; p.world = h->world;
    p.segA0 = P.segA0; p.mixedW = P.mixedW; p.segB0 = P.segB0; p.segB0len = P.segB0len; p.segB1 = P.segB1;
    p.table = D.table.p; p.counter = D.counter.p;
    p.word_np = S.word_np;
    p.W = S.W; p.WA = P.WA; p.NP = S.NP;
    p.nA = P.nA; p.nB = P.nB; p.padA = P.padA; p.padB = P.padB; p.thrA = P.thrA; p.thrB = P.thrB;
    p.mixed = P.mixed; p.maskA = P.maskA; p.maskB = P.maskB;
    const int64_t rr = S.r, nc = ncols;
    int64_t executed;   // pair evaluations x samples actually needed (self pairs and pads excluded)
    if (all) {
        p.row_planes = p.col_planes = S.planes;
        p.row_gene = p.col_gene = D.iota.p;
        p.ntr = p.ntc = p.nsym = S.NT;
        executed = rr * (rr - 1) / 2 * S.c;
    } else if (sym_worth(nc, rr)) {
        const int ntc = (ncols + REO_TILE - 1) / REO_TILE;
        const int nsymp = (ntc + T - 1) / T * T;
        const int ntn = (int)((rr - nc + REO_TILE - 1) / REO_TILE);
        CKL(reo_launch_sym_lists(S.r, m_old, m_new, T, D.list_gene.p, D.list_sign.p, D.counts.p + 5, D.list_cap, D.st));
        CK(D.panel.ensure((size_t)(nsymp + ntn + 1) * S.tile_stride()));
        CKL(reo_launch_gather_panel(S.planes, S.W, S.NP, D.list_gene.p, nsymp + ntn, D.panel.p, D.st));
        h->kernel_launches += 2;
        p.row_planes = p.col_planes = D.panel.p;
        p.row_gene = p.col_gene = D.list_gene.p;
        p.row_sign = p.col_sign = m_new ? D.list_sign.p : nullptr;
        p.ntr = nsymp + ntn; p.ntc = p.nsym = ntc;
        executed = (nc * (nc - 1) / 2 + (rr - nc) * nc) * S.c;
    } else {
        const int ntc = (ncols + REO_TILE - 1) / REO_TILE;
        const int32_t* cg = D.changed_gene.p;
        if (!m_new) {
            CKL(reo_launch_mask_to_list(m_old, S.r, D.col_gene.p, D.counts.p + 4, D.st));
            h->kernel_launches++;
            cg = D.col_gene.p;
        }
        CK(D.panel.ensure((size_t)(ntc + 1) * S.tile_stride()));
        CKL(reo_launch_gather_panel(S.planes, S.W, S.NP, cg, ntc, D.panel.p, D.st));
        h->kernel_launches++;
        p.row_planes = S.planes; p.col_planes = D.panel.p;
        p.row_gene = D.iota.p; p.col_gene = cg;
        p.col_sign = m_new ? D.changed_sign.p : nullptr;
        p.ntr = S.NT; p.ntc = ntc; p.nsym = 0;
        executed = (rr * nc - nc) * S.c;
    }
    CK(cudaMemsetAsync(D.counter.p, 0, sizeof(unsigned int), D.st));
    int rc = record_pair_events(h, D, true);
    if (rc) return rc;
    CKL(reo_launch_pairs2(p, D.num_sms, D.st));
    if ((rc = record_pair_events(h, D, false))) return rc;
    h->compares += rank_share(h, executed);
    h->ordered_triples += rank_share(h, rr * nc * S.c);
    return REO_OK;
}

// full build for the mask in mask_dev; ncols = its population count when the host already knows it (-1: read it back)
int build_tables_full(reo_handle_t h, ReoDev& D, const LevelPlan& P, const uint8_t* mask_dev, int ncols = -1) {
    const ReoStaged& S = D.S;
    CK(cudaMemsetAsync(D.table.p, 0, (size_t)S.r * 9 * sizeof(int32_t), D.st));
    if (ncols < 0) {
        CKL(reo_launch_mask_to_list(mask_dev, S.r, D.col_gene.p, D.counts.p + 4, D.st));
        h->kernel_launches++;
        CK(cudaMemcpyAsync(D.h_counts + 4, D.counts.p + 4, sizeof(int32_t), cudaMemcpyDeviceToHost, D.st));
        CK(cudaStreamSynchronize(D.st));
        ncols = D.h_counts[4];
    }
    return launch_tables(h, D, P, mask_dev, nullptr, ncols);
}

// Several ranks: every rank holds partial sums for ALL genes (its share of the pair tiles); the statistics read
// their sum.  NCCL all-reduce on the library's stream, or the user's all-gather callback followed by a local sum.
int reduce_tables(reo_handle_t h, ReoDev& D) {
    if (h->world <= 1) return REO_OK;
    const size_t n = (size_t)D.S.r * 9;
    if (D.comm) {
        ncclResult_t nr = g_nccl.AllReduce(D.table.p, D.table_red.p, n, ncclInt32, ncclSum, D.comm, D.st);
        if (nr != ncclSuccess) return fail(h, REO_ERR_COMM, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(nr));
        h->kernel_launches++;
        return REO_OK;
    }
    if (!h->ag_fn) return fail(h, REO_ERR_COMM, "world > 1 but neither an NCCL communicator nor an all-gather callback is set");
    CK(cudaMemcpyAsync(D.table_all.p + (size_t)h->rank * n, D.table.p, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, D.st));
    CK(cudaStreamSynchronize(D.st));
    if (h->ag_fn(h->ag_ctx, D.table_all.p, (uint64_t)n * sizeof(int32_t)) != 0) return fail(h, REO_ERR_COMM, "all-gather callback failed");
    CKL(reo_launch_sum_slices(D.table_all.p, h->world, (int64_t)n, D.table_red.p, D.st));
    h->kernel_launches++;
    return REO_OK;
}

// One evaluation of src:402-417 on the current tables = enqueue_mcc + enqueue_eval.
// src:402-406: tables -> result columns 2..14.  Those 13 columns are final for this evaluation, so their
// device->host copy starts on the copy stream at once and overlaps the rest of the sequence (early_dst: the host
// image of `result` for this level, pinned).
int enqueue_mcc(reo_handle_t h, ReoDev& D, int64_t r, double* early_dst) {
    if (D.early_pending) CK(cudaStreamWaitEvent(D.st, D.ev[7], 0));   // the previous copy still reads `result`
    CKL(reo_launch_mccullagh_tables(D.table_cur, r, D.result.p, D.st));
    if (early_dst) {
        CK(cudaEventRecord(D.ev[6], D.st));
        CK(cudaStreamWaitEvent(D.st_copy, D.ev[6], 0));
        CK(cudaMemcpyAsync(early_dst + 2 * r, D.result.p + 2 * r, (size_t)r * 13 * sizeof(double), cudaMemcpyDeviceToHost, D.st_copy));
        CK(cudaEventRecord(D.ev[7], D.st_copy));
        D.early_pending = true;
    }
    return REO_OK;
}

// src:409-417: sort + trimmed std, empirical-null p, BH, new mask, symmetric difference, and the three counters the
// host decides on (read back into pinned memory).
int enqueue_eval(reo_handle_t h, ReoDev& D, int64_t r, double pval_deg, double padj_deg, uint8_t* mask_cur,
                 uint8_t* mask_new) {
    CKL(reo_launch_sort_f64(D.result.p + (size_t)r * 11, r, D.sorted.p, D.perm.p, D.sortws, D.st));     // src:409
    CKL(reo_launch_trimmed_std(D.sorted.p, r, D.se.p, D.std_ws.p, D.st));                               // src:411
    // src:412-413: p-values from the sorted d1; their ascending order follows from it too (p decreases with |d1|):
    // no second sort.  BH also writes the new reference mask (src:417).
    CKL(reo_launch_p_order(D.sorted.p, D.perm.p, r, D.result.p, D.se.p, D.sorted_p.p, D.perm2.p, D.st));
    CKL(reo_launch_bh(D.sorted_p.p, D.perm2.p, r, D.result.p + r, D.sorted.p /* free after p_order: scratch */,
                      mask_new, pval_deg, padj_deg, D.st));
    CKL(reo_launch_mask_diff(r, mask_cur, mask_new, D.counts.p, D.changed_gene.p, D.changed_sign.p, D.st));
    CK(cudaMemcpyAsync(D.h_counts, D.counts.p, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, D.st));
    return REO_OK;
}

void drop_eval_graphs(ReoDev& D) {
    for (auto& e : D.eval_exec) { if (e) cudaGraphExecDestroy(e); e = nullptr; }
}

// Launch the evaluation sequence: the table statistics first, then the remaining 7 small kernels as one CUDA graph
// (their launch gaps otherwise dominate).
int run_eval(reo_handle_t h, ReoDev& D, int64_t r, double pval_deg, double padj_deg, uint8_t* mask_cur,
             uint8_t* mask_new, double* early_dst) {
    h->kernel_launches += 10;  // mccullagh, sort x2, std x3, p_order, bh x2, diff
    static const bool no_graph = getenv("REO_NO_GRAPH") != nullptr;
    CK(reo_sort_reserve(D.sortws, r, D.st));
    const int rc0 = enqueue_mcc(h, D, r, early_dst);
    if (rc0 != REO_OK) return rc0;
    if (debug_sync() || no_graph) return enqueue_eval(h, D, r, pval_deg, padj_deg, mask_cur, mask_new);
    const std::vector<const void*> key = {D.table_cur, D.result.p, D.sorted.p, D.perm.p, D.sorted_p.p, D.perm2.p, D.se.p,
                                          D.counts.p, D.changed_gene.p, D.changed_sign.p, D.sortws.keys, D.std_ws.p,
                                          D.mask_a.p, D.mask_b.p, D.h_counts};
    if (key != D.eval_key || r != D.eval_r || pval_deg != D.eval_pd || padj_deg != D.eval_qd) {
        drop_eval_graphs(D);
        D.eval_key = key; D.eval_r = r; D.eval_pd = pval_deg; D.eval_qd = padj_deg;
    }
    const int which = (mask_cur == D.mask_a.p) ? 0 : 1;
    if (!D.eval_exec[which]) {
        CK(cudaStreamBeginCapture(D.st, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_eval(h, D, r, pval_deg, padj_deg, mask_cur, mask_new);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(D.st, &g);
        if (rc != REO_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail_cuda(h, e, "cudaStreamEndCapture");
        const cudaError_t e2 = cudaGraphInstantiate(&D.eval_exec[which], g, 0);
        cudaGraphDestroy(g);
        if (e2 != cudaSuccess) { D.eval_exec[which] = nullptr; return fail_cuda(h, e2, "cudaGraphInstantiate"); }
    }
    CK(cudaGraphLaunch(D.eval_exec[which], D.st));
    return REO_OK;
}

// run f(rank handle, rank) on every device of a multi-device handle, one host thread per device
template <typename F>
int run_multi(reo_handle_t parent, F f) {
    const int n = (int)parent->subs.size();
    std::vector<int> rcs(n, REO_OK);
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i) th.emplace_back([&, i] { rcs[i] = f(parent->subs[i], i); });
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rcs[i] != REO_OK) { parent->err = parent->subs[i]->err; return rcs[i]; }
    return REO_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int reo_version(void) { return REO_VERSION; }

void* reo_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void reo_host_free(void* p) { if (p) cudaFreeHost(p); }

const char* reo_last_error(reo_handle_t h) {
    if (h) return h->err.c_str();
    std::lock_guard<std::mutex> g(g_err_mu);
    return g_create_err.c_str();
}

static int create_single(reo_handle_t* out, int dev, uint64_t seed) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1) {
        cudaGetLastError();
        return fail(nullptr, REO_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= count) return fail(nullptr, REO_ERR_ARG, "reo_create: bad device index");
    reo_handle_s* h = new (std::nothrow) reo_handle_s();
    if (!h) return fail(nullptr, REO_ERR_OOM, "host allocation failed");
    h->seed = seed;
    h->devs.resize(1);
    ReoDev& D = h->devs[0];
    D.dev = dev;
    if ((e = cudaSetDevice(D.dev)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&D.st, cudaStreamNonBlocking)) != cudaSuccess) {
        delete h; cudaGetLastError();
        return fail(nullptr, REO_ERR_CUDA, std::string("reo_create: ") + cudaGetErrorString(e));
    }
    cudaDeviceGetAttribute(&D.num_sms, cudaDevAttrMultiProcessorCount, D.dev);
    if ((e = cudaStreamCreateWithFlags(&D.st_copy, cudaStreamNonBlocking)) != cudaSuccess) {
        delete h; cudaGetLastError();
        return fail(nullptr, REO_ERR_CUDA, std::string("reo_create: ") + cudaGetErrorString(e));
    }
    for (auto& ev : D.ev) cudaEventCreate(&ev);
    if (cudaMallocHost((void**)&D.h_counts, 16 * sizeof(int32_t)) != cudaSuccess) {
        delete h; cudaGetLastError();
        return fail(nullptr, REO_ERR_OOM, "pinned allocation failed");
    }
    *out = h;
    return REO_OK;
}

int reo_create(reo_handle_t* out, int ndev, const int* devs, uint64_t seed, uint32_t /*flags*/) {
    if (!out || ndev < 1) return fail(nullptr, REO_ERR_ARG, "reo_create: bad argument");
    if (ndev == 1) return create_single(out, devs ? devs[0] : 0, seed);
    // single process, several GPUs (the Julia deployment): one rank handle per device, NCCL communicators
    // from ncclCommInitAll, one host thread per device inside every call
    {
        std::lock_guard<std::mutex> g(g_nccl_mu);
        if (!g_nccl.load()) return fail(nullptr, REO_ERR_COMM, g_nccl.err);
    }
    reo_handle_s* parent = new (std::nothrow) reo_handle_s();
    if (!parent) return fail(nullptr, REO_ERR_OOM, "host allocation failed");
    parent->seed = seed;
    std::vector<int> devlist(ndev);
    for (int i = 0; i < ndev; ++i) devlist[i] = devs ? devs[i] : i;
    for (int i = 0; i < ndev; ++i) {
        reo_handle_t sub = nullptr;
        const int rc = create_single(&sub, devlist[i], seed);
        if (rc != REO_OK) { for (auto* s : parent->subs) reo_destroy(s); delete parent; return rc; }
        sub->rank = i; sub->world = ndev;
        parent->subs.push_back(sub);
    }
    std::vector<ncclComm_t> comms(ndev);
    const ncclResult_t nr = g_nccl.CommInitAll(comms.data(), ndev, devlist.data());
    if (nr != ncclSuccess) {
        const std::string m = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(nr);
        for (auto* s : parent->subs) reo_destroy(s);
        delete parent;
        return fail(nullptr, REO_ERR_COMM, m);
    }
    for (int i = 0; i < ndev; ++i) parent->subs[i]->devs[0].comm = comms[i];
    *out = parent;
    return REO_OK;
}

/* 128-byte NCCL unique id for reo_comm_init_rank (rank 0 creates it, every rank receives a copy). */
int reo_comm_unique_id(void* out128) {
    if (!out128) return REO_ERR_ARG;
    std::lock_guard<std::mutex> g(g_nccl_mu);
    if (!g_nccl.load()) return fail(nullptr, REO_ERR_COMM, g_nccl.err);
    ncclUniqueId id;
    const ncclResult_t nr = g_nccl.GetUniqueId(&id);
    if (nr != ncclSuccess) return fail(nullptr, REO_ERR_COMM, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(nr));
    memcpy(out128, &id, sizeof(id));
    return REO_OK;
}

/* One process per GPU: join an NCCL communicator; gene-row tiles are then sharded by rank and the tables
 * all-gathered by NCCL on the library's own stream. */
int reo_comm_init_rank(reo_handle_t h, int rank, int world, const void* id128) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world || !h->subs.empty())
        return fail(h, REO_ERR_ARG, "reo_comm_init_rank: bad argument");
    {
        std::lock_guard<std::mutex> g(g_nccl_mu);
        if (!g_nccl.load()) return fail(h, REO_ERR_COMM, g_nccl.err);
    }
    ReoDev& D = h->devs[0];
    CK(cudaSetDevice(D.dev));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    if (D.comm) { g_nccl.CommDestroy(D.comm); D.comm = nullptr; }
    const ncclResult_t nr = g_nccl.CommInitRank(&D.comm, world, id, rank);
    if (nr != ncclSuccess) return fail(h, REO_ERR_COMM, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(nr));
    h->rank = rank; h->world = world;
    D.S.valid = false;
    return REO_OK;
}

int reo_destroy(reo_handle_t h) {
    if (!h) return REO_OK;
    for (auto* s : h->subs) reo_destroy(s);
    for (ReoDev& D : h->devs) {
        cudaSetDevice(D.dev);
        if (D.st) cudaStreamSynchronize(D.st);
        if (D.comm) { g_nccl.CommDestroy(D.comm); D.comm = nullptr; }
        drop_eval_graphs(D);
        D.table_red.release(); D.table_all.release(); D.list_gene.release(); D.list_sign.release();
        D.k1_send.release(); D.k1_gather.release(); D.stage_lists.release(); D.widelist.release();
        D.raw.release(); D.raw2.release(); D.pb.release(); D.sub.release(); D.ranks.release(); D.planes.release(); D.panel.release(); D.slot_of_sample.release();
        D.sample_of_slot.release(); D.word_order.release(); D.iota.release(); D.col_gene.release();
        D.changed_gene.release(); D.table.release(); D.perm.release(); D.perm2.release(); D.counts.release(); D.fblist.release();
        D.small_i.release(); D.changed_sign.release(); D.updown.release(); D.mask_a.release(); D.mask_b.release();
        D.result.release(); D.sorted.release(); D.sorted_p.release(); D.se.release(); D.small_d.release();
        D.counter.release(); D.flags.release(); D.fb_keys.release(); D.fb_rank.release(); D.small_ll.release();
        if (D.sortws.keys) cudaFree(D.sortws.keys);
        if (D.sortws.idx) cudaFree(D.sortws.idx);
        if (D.sortws.pos) cudaFree(D.sortws.pos);
        if (D.sortws.cnt) cudaFree(D.sortws.cnt);
        if (D.h_counts) cudaFreeHost(D.h_counts);
        if (D.h_out) cudaFreeHost(D.h_out);
        for (auto& bptr : D.bounce) if (bptr) cudaFreeHost(bptr);
        for (auto& e : D.bounce_ev) if (e) cudaEventDestroy(e);
        delete D.pool; D.pool = nullptr;
        D.std_ws.release();
        for (auto& ev : D.ev) if (ev) cudaEventDestroy(ev);
        for (auto& ev : D.pev) cudaEventDestroy(ev);
        for (auto& ev : D.copy_ev) cudaEventDestroy(ev);
        if (D.st_copy) cudaStreamDestroy(D.st_copy);
        if (D.st) cudaStreamDestroy(D.st);
    }
    delete h;
    return REO_OK;
}

int reo_set_collective(reo_handle_t h, int rank, int world, reo_allgather_fn fn, void* ctx) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) return fail(h, REO_ERR_ARG, "reo_set_collective: handle already drives several devices");
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) return fail(h, REO_ERR_ARG, "reo_set_collective: bad argument");
    h->rank = rank; h->world = world; h->ag_fn = fn; h->ag_ctx = ctx;
    h->devs[0].S.valid = false;  // table sizing depends on world
    return REO_OK;
}

int reo_threshold(int n, double pval) {
    if (n < 0) return -1;
    // pval_min = pvalue(Binomial(n), 0) = min(1, 2 * 2^-n)   (src:83)
    long double pmin = ldexpl(1.0L, 1 - n);
    if (pmin > 1.0L) pmin = 1.0L;
    if (!(pmin < (long double)pval)) return n;  // warn path, src:88-90
    // p(x) is non-decreasing on 0..floor(n/2): binary search the first x with p(x) > pval (src:85)
    int lo = 0, hi = n / 2;
    if (!(two_sided_binom_p(n, hi) > (long double)pval)) return -1;  // findfirst -> nothing
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (two_sided_binom_p(n, mid) > (long double)pval) hi = mid; else lo = mid + 1;
    }
    return n - lo + 1;
}

int reo_stage(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld, const int32_t* group_id,
              int32_t gnum, uint32_t flags) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) {
        if (flags & REO_DATA_ON_DEVICE) return fail(h, REO_ERR_UNSUPPORTED, "REO_DATA_ON_DEVICE needs a single-device handle");
        return run_multi(h, [&](reo_handle_t s, int) { return reo_stage(s, data, dtype, r, c, ld, group_id, gnum, flags); });
    }
    h->kernel_launches = 0; h->pair_launches = 0; h->compares = 0; h->ordered_triples = 0;
    h->devs[0].n_pev = 0;
    return do_stage(h, data, dtype, r, c, ld, group_id, gnum, flags);
}

int reo_stage_info(reo_handle_t h, int32_t* rank_bits, int32_t* sample_words, int32_t* gene_tiles) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) return reo_stage_info(h->subs[0], rank_bits, sample_words, gene_tiles);
    const ReoStaged& S = h->devs[0].S;
    if (!S.valid) return fail(h, REO_ERR_STATE, "no staged matrix");
    if (rank_bits) *rank_bits = S.B;
    if (sample_words) *sample_words = S.W;
    if (gene_tiles) *gene_tiles = S.NT;
    return REO_OK;
}

int reo_pair_counts(reo_handle_t h, int32_t k, const int32_t* rows, int32_t nrows, const int32_t* cols, int32_t ncols,
                    int32_t* nre, int32_t* rest) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) {
        const int rc = reo_pair_counts(h->subs[0], k, rows, nrows, cols, ncols, nre, rest);
        if (rc) h->err = h->subs[0]->err;
        return rc;
    }
    ReoDev& D = h->devs[0];
    const ReoStaged& S = D.S;
    if (!S.valid) return fail(h, REO_ERR_STATE, "no staged matrix");
    if (!rows || !cols || !nre || !rest || nrows < 1 || ncols < 1 || k < 0 || k >= S.gnum)
        return fail(h, REO_ERR_ARG, "reo_pair_counts: bad argument");
    for (int i = 0; i < nrows; ++i) if (rows[i] < 0 || rows[i] >= S.r) return fail(h, REO_ERR_ARG, "row index out of range");
    for (int i = 0; i < ncols; ++i) if (cols[i] < 0 || cols[i] >= S.r) return fail(h, REO_ERR_ARG, "col index out of range");
    CK(cudaSetDevice(D.dev));
    int32_t thr_dummy[128] = {0};
    LevelPlan P = make_plan(S, k, S.gnum <= 64 ? thr_dummy : nullptr, 0.01);
    int rc = upload_plan(h, D, P);
    if (rc) return rc;
    const size_t n = (size_t)nrows * ncols;
    CK(D.small_i.ensure(nrows + ncols + 2 * n));
    int32_t* d_rows = D.small_i.p; int32_t* d_cols = d_rows + nrows; int32_t* d_nre = d_cols + ncols; int32_t* d_rest = d_nre + n;
    CK(cudaMemcpyAsync(d_rows, rows, nrows * 4, cudaMemcpyHostToDevice, D.st));
    CK(cudaMemcpyAsync(d_cols, cols, ncols * 4, cudaMemcpyHostToDevice, D.st));
    CKL(reo_launch_pair_counts_small(S, D.word_order.p, P.WA, d_rows, nrows, d_cols, ncols, d_nre, d_rest, P.padA, P.padB, P.mixed, P.maskA,
                                     P.maskB, D.st));
    CK(cudaMemcpyAsync(nre, d_nre, n * 4, cudaMemcpyDeviceToHost, D.st));
    CK(cudaMemcpyAsync(rest, d_rest, n * 4, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    return REO_OK;
}

int reo_tables_delta(reo_handle_t h, int32_t k, const int32_t* thresholds, double pval_reo, const uint8_t* mask_from,
                     const uint8_t* mask_to, int32_t* table) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) {
        if (!table) return fail(h, REO_ERR_ARG, "reo_tables: bad argument");
        const int64_t rr = h->subs[0]->devs[0].S.r;
        std::vector<std::vector<int32_t>> scratch(h->subs.size());
        return run_multi(h, [&](reo_handle_t s, int i) {
            int32_t* out = table;
            if (i > 0) { scratch[i].resize((size_t)std::max<int64_t>(rr, 1) * 9); out = scratch[i].data(); }
            return reo_tables_delta(s, k, thresholds, pval_reo, mask_from, mask_to, out);
        });
    }
    ReoDev& D = h->devs[0];
    const ReoStaged& S = D.S;
    if (!S.valid) return fail(h, REO_ERR_STATE, "no staged matrix");
    if (!mask_from || !table || k < 0 || k >= S.gnum) return fail(h, REO_ERR_ARG, "reo_tables: bad argument");
    CK(cudaSetDevice(D.dev));
    LevelPlan P = make_plan(S, k, thresholds, pval_reo);
    int rc = upload_plan(h, D, P);
    if (rc) return rc;
    CK(cudaMemcpyAsync(D.mask_a.p, mask_from, S.r, cudaMemcpyHostToDevice, D.st));
    if ((rc = build_tables_full(h, D, P, D.mask_a.p))) return rc;
    if (mask_to) {
        CK(cudaMemcpyAsync(D.mask_b.p, mask_to, S.r, cudaMemcpyHostToDevice, D.st));
        CKL(reo_launch_mask_diff(S.r, D.mask_a.p, D.mask_b.p, D.counts.p, D.changed_gene.p, D.changed_sign.p, D.st));
        CK(cudaMemcpyAsync(D.h_counts, D.counts.p, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, D.st));
        CK(cudaStreamSynchronize(D.st));
        if ((rc = launch_tables(h, D, P, D.mask_a.p, D.mask_b.p, D.h_counts[2]))) return rc;
    }
    if ((rc = reduce_tables(h, D))) return rc;
    CK(cudaMemcpyAsync(table, D.table_cur, (size_t)S.r * 9 * sizeof(int32_t), cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    return REO_OK;
}

int reo_tables(reo_handle_t h, int32_t k, const int32_t* thresholds, double pval_reo, const uint8_t* mask,
               int32_t* table) {
    return reo_tables_delta(h, k, thresholds, pval_reo, mask, nullptr, table);
}

int reo_mccullagh(reo_handle_t h, const int64_t* tables, int64_t n, int32_t k, double* out) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) { const int rc = reo_mccullagh(h->subs[0], tables, n, k, out); if (rc) h->err = h->subs[0]->err; return rc; }
    ReoDev& D = h->devs[0];
    if (!tables || !out || n < 1 || k < 2 || k > 9) return fail(h, REO_ERR_ARG, "reo_mccullagh: bad argument");
    CK(cudaSetDevice(D.dev));
    CK(D.small_ll.ensure((size_t)n * k * k));
    CK(D.small_d.ensure((size_t)n * 5));
    CK(cudaMemcpyAsync(D.small_ll.p, tables, (size_t)n * k * k * 8, cudaMemcpyHostToDevice, D.st));
    CKL(reo_launch_mccullagh_kxk((const int64_t*)D.small_ll.p, n, k, D.small_d.p, D.st));
    CK(cudaMemcpyAsync(out, D.small_d.p, (size_t)n * 5 * 8, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    return REO_OK;
}

int reo_sort_f64(reo_handle_t h, const double* x, int64_t n, double* sorted, int32_t* perm) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) { const int rc = reo_sort_f64(h->subs[0], x, n, sorted, perm); if (rc) h->err = h->subs[0]->err; return rc; }
    ReoDev& D = h->devs[0];
    if (!x || !sorted || n < 1) return fail(h, REO_ERR_ARG, "reo_sort_f64: bad argument");
    CK(cudaSetDevice(D.dev));
    CK(D.small_d.ensure((size_t)2 * n));
    CK(D.perm.ensure(n));
    CK(cudaMemcpyAsync(D.small_d.p, x, n * 8, cudaMemcpyHostToDevice, D.st));
    CKL(reo_launch_sort_f64(D.small_d.p, n, D.small_d.p + n, D.perm.p, D.sortws, D.st));
    CK(cudaMemcpyAsync(sorted, D.small_d.p + n, n * 8, cudaMemcpyDeviceToHost, D.st));
    if (perm) CK(cudaMemcpyAsync(perm, D.perm.p, n * 4, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    return REO_OK;
}

int reo_empirical_null(reo_handle_t h, const double* delta1, int64_t n, double* pval, double* se) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) { const int rc = reo_empirical_null(h->subs[0], delta1, n, pval, se); if (rc) h->err = h->subs[0]->err; return rc; }
    ReoDev& D = h->devs[0];
    if (!delta1 || !pval || n < 1) return fail(h, REO_ERR_ARG, "reo_empirical_null: bad argument");
    if (n <= 10) return fail(h, REO_ERR_BOUNDS, "BoundsError: r <= 10 (src:411)");
    CK(cudaSetDevice(D.dev));
    CK(D.small_d.ensure((size_t)3 * n + 1));
    double* d_x = D.small_d.p; double* d_s = d_x + n; double* d_p = d_s + n; double* d_se = d_p + n;
    CK(cudaMemcpyAsync(d_x, delta1, n * 8, cudaMemcpyHostToDevice, D.st));
    CKL(reo_launch_sort_f64(d_x, n, d_s, nullptr, D.sortws, D.st));
    { int rc2 = ensure_std_ws(h, D); if (rc2) return rc2; }
    CKL(reo_launch_trimmed_std(d_s, n, d_se, D.std_ws.p, D.st));
    CKL(reo_launch_null_pvals(d_x, n, d_se, d_p, D.st));
    CK(cudaMemcpyAsync(pval, d_p, n * 8, cudaMemcpyDeviceToHost, D.st));
    if (se) CK(cudaMemcpyAsync(se, d_se, 8, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    return REO_OK;
}

int reo_bh(reo_handle_t h, const double* p, int64_t n, double* padj) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) { const int rc = reo_bh(h->subs[0], p, n, padj); if (rc) h->err = h->subs[0]->err; return rc; }
    ReoDev& D = h->devs[0];
    if (!p || !padj || n < 1) return fail(h, REO_ERR_ARG, "reo_bh: bad argument");
    CK(cudaSetDevice(D.dev));
    CK(D.small_d.ensure((size_t)4 * n + n / 1024 + 2));
    CK(D.perm.ensure(n));
    double* d_x = D.small_d.p; double* d_s = d_x + n; double* d_q = d_s + n; double* d_w = d_q + n;
    CK(cudaMemcpyAsync(d_x, p, n * 8, cudaMemcpyHostToDevice, D.st));
    if (n > 1) {
        CKL(reo_launch_sort_f64(d_x, n, d_s, D.perm.p, D.sortws, D.st));
        CKL(reo_launch_bh(d_s, D.perm.p, n, d_q, d_w, nullptr, 0.0, 0.0, D.st));
    } else {
        d_q = d_x;
    }
    CK(cudaMemcpyAsync(padj, d_q, n * 8, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    return REO_OK;
}

// matrix on the device for the pre-processing entry points: the caller's device pointer, or a copy in raw2
static int prep_input(reo_handle_t h, ReoDev& D, const void* data, int dtype, int64_t r, int64_t c, int64_t ld,
                      uint32_t flags, const uint8_t** dev, int64_t* dev_ld) {
    const size_t es = dtype_size(dtype);
    if (!data || es == 0 || r < 1 || c < 1 || ld < r) return fail(h, REO_ERR_ARG, "bad matrix argument");
    CK(cudaSetDevice(D.dev));
    if (flags & REO_DATA_ON_DEVICE) { *dev = (const uint8_t*)data; *dev_ld = ld; return REO_OK; }
    CK(D.raw2.ensure((size_t)r * c * es));
    CK(cudaMemcpy2DAsync(D.raw2.p, (size_t)r * es, data, (size_t)ld * es, (size_t)r * es, (size_t)c, cudaMemcpyHostToDevice, D.st));
    *dev = D.raw2.p; *dev_ld = r;
    return REO_OK;
}

/* pseudobulk_group, src:56-67 / 608-612: out[:, p] = sum over the cells cell_list[cell_ptr[p] .. cell_ptr[p+1]) (in that
 * order).  out is r x nprofiles column-major, Int64 for integer input and Float64 for float input; out_host and/or
 * out_dev (handle-owned, valid until the next reo_pseudobulk) may be NULL. */
int reo_pseudobulk(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld, const int32_t* cell_ptr,
                   const int32_t* cell_list, int32_t nprofiles, uint32_t flags, void* out_host, void** out_dev) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) { const int rc = reo_pseudobulk(h->subs[0], data, dtype, r, c, ld, cell_ptr, cell_list, nprofiles, flags, out_host, out_dev); if (rc) h->err = h->subs[0]->err; return rc; }
    ReoDev& D = h->devs[0];
    if (!cell_ptr || !cell_list || nprofiles < 1) return fail(h, REO_ERR_ARG, "reo_pseudobulk: bad argument");
    const int32_t ncells = cell_ptr[nprofiles];
    if (cell_ptr[0] != 0 || ncells < 0) return fail(h, REO_ERR_ARG, "reo_pseudobulk: bad cell_ptr");
    for (int p = 0; p < nprofiles; ++p) if (cell_ptr[p + 1] < cell_ptr[p]) return fail(h, REO_ERR_ARG, "reo_pseudobulk: bad cell_ptr");
    for (int32_t k = 0; k < ncells; ++k) if (cell_list[k] < 0 || cell_list[k] >= c) return fail(h, REO_ERR_DIM, "reo_pseudobulk: cell index out of range");
    const uint8_t* dev; int64_t dld;
    int rc = prep_input(h, D, data, dtype, r, c, ld, flags, &dev, &dld);
    if (rc) return rc;
    CK(D.small_i.ensure((size_t)nprofiles + 1 + (size_t)std::max(ncells, 1)));
    CK(cudaMemcpyAsync(D.small_i.p, cell_ptr, ((size_t)nprofiles + 1) * 4, cudaMemcpyHostToDevice, D.st));
    if (ncells > 0) CK(cudaMemcpyAsync(D.small_i.p + nprofiles + 1, cell_list, (size_t)ncells * 4, cudaMemcpyHostToDevice, D.st));
    CK(D.pb.ensure((size_t)r * nprofiles * 8));
    CKL(reo_launch_pseudobulk(dev, dtype, r, dld, D.small_i.p, D.small_i.p + nprofiles + 1, nprofiles, D.pb.p, D.st));
    if (out_host) CK(cudaMemcpyAsync(out_host, D.pb.p, (size_t)r * nprofiles * 8, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    if (out_dev) *out_dev = D.pb.p;
    return REO_OK;
}

/* src:618, 626: per_cell[s] = #{ i : data[i,s] > 0 } (c ints), per_gene[i] = #{ s : data[i,s] > 0 } (r ints). */
int reo_detect_counts(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld, uint32_t flags,
                      int32_t* per_cell, int32_t* per_gene) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) { const int rc = reo_detect_counts(h->subs[0], data, dtype, r, c, ld, flags, per_cell, per_gene); if (rc) h->err = h->subs[0]->err; return rc; }
    ReoDev& D = h->devs[0];
    if (!per_cell || !per_gene) return fail(h, REO_ERR_ARG, "reo_detect_counts: bad argument");
    const uint8_t* dev; int64_t dld;
    int rc = prep_input(h, D, data, dtype, r, c, ld, flags, &dev, &dld);
    if (rc) return rc;
    CK(D.small_i.ensure((size_t)r + c));
    CK(cudaMemsetAsync(D.small_i.p, 0, ((size_t)r + c) * 4, D.st));
    CKL(reo_launch_detect_counts(dev, dtype, r, c, dld, D.small_i.p, D.small_i.p + c, D.st));
    CK(cudaMemcpyAsync(per_cell, D.small_i.p, (size_t)c * 4, cudaMemcpyDeviceToHost, D.st));
    CK(cudaMemcpyAsync(per_gene, D.small_i.p + c, (size_t)r * 4, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    return REO_OK;
}

/* src:624-628: out = data[gene_list, cell_list], r2 x c2 column-major, same dtype.  out_host / out_dev as above. */
int reo_subset(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld, const int32_t* gene_list,
               int64_t r2, const int32_t* cell_list, int64_t c2, uint32_t flags, void* out_host, void** out_dev) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) { const int rc = reo_subset(h->subs[0], data, dtype, r, c, ld, gene_list, r2, cell_list, c2, flags, out_host, out_dev); if (rc) h->err = h->subs[0]->err; return rc; }
    ReoDev& D = h->devs[0];
    if (!gene_list || !cell_list || r2 < 1 || c2 < 1) return fail(h, REO_ERR_ARG, "reo_subset: bad argument");
    for (int64_t i = 0; i < r2; ++i) if (gene_list[i] < 0 || gene_list[i] >= r) return fail(h, REO_ERR_DIM, "reo_subset: gene index out of range");
    for (int64_t s = 0; s < c2; ++s) if (cell_list[s] < 0 || cell_list[s] >= c) return fail(h, REO_ERR_DIM, "reo_subset: cell index out of range");
    const uint8_t* dev; int64_t dld;
    int rc = prep_input(h, D, data, dtype, r, c, ld, flags, &dev, &dld);
    if (rc) return rc;
    const size_t es = dtype_size(dtype);
    CK(D.small_i.ensure((size_t)r2 + c2));
    CK(cudaMemcpyAsync(D.small_i.p, gene_list, (size_t)r2 * 4, cudaMemcpyHostToDevice, D.st));
    CK(cudaMemcpyAsync(D.small_i.p + r2, cell_list, (size_t)c2 * 4, cudaMemcpyHostToDevice, D.st));
    CK(D.sub.ensure((size_t)r2 * c2 * es));
    CKL(reo_launch_subset(dev, dtype, dld, D.small_i.p, r2, D.small_i.p + r2, c2, D.sub.p, D.st));
    if (out_host) CK(cudaMemcpyAsync(out_host, D.sub.p, (size_t)r2 * c2 * es, cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    if (out_dev) *out_dev = D.sub.p;
    return REO_OK;
}

/* Test hook (CPU, no GPU needed): the tile pairs rank `rank` of `world` evaluates for a table build over `ncols` column
 * genes out of `r` genes staged with `sample_words` words and `planes` bit planes -- the library's own partition code
 * (reo_pairs2.cu), replayed on the host.  mode: 0 = the shape the library would choose, 1 = force one-sided.
 * out: triples (row tile, column tile, 1 = the pairs update their row genes | 2 = ... their column genes) in the panel
 * coordinates of that shape; *nsym_tiles / *ntr / *ntc describe it.  Returns the number of triples (may exceed cap). */
long long reo_debug_pair_plan(int64_t r, int64_t ncols, int32_t sample_words, int32_t planes, int32_t rank, int32_t world,
                              int32_t mode, int32_t* out, int64_t cap, int32_t* nsym_tiles, int32_t* ntr, int32_t* ntc) {
    if (r < 1 || ncols < 1 || ncols > r || world < 1 || rank < 0 || rank >= world || sample_words < 1 || planes < 2) return -1;
    ReoPair2Params p;
    memset(&p, 0, sizeof(p));
    p.T = reo_pairs2_block_edge(sample_words, planes);
    p.rank = rank; p.world = world; p.W = sample_words; p.NP = planes;
    const int NT = (int)((r + REO_TILE - 1) / REO_TILE);
    const int tc = (int)((ncols + REO_TILE - 1) / REO_TILE);
    if (ncols == r && mode == 0) { p.ntr = p.ntc = p.nsym = NT; }
    else if (mode == 0 && sym_worth(ncols, r)) {
        const int nsymp = (tc + p.T - 1) / p.T * p.T;
        p.ntr = nsymp + (int)((r - ncols + REO_TILE - 1) / REO_TILE); p.ntc = p.nsym = tc;
    } else { p.ntr = NT; p.ntc = tc; p.nsym = 0; }
    if (nsym_tiles) *nsym_tiles = p.nsym;
    if (ntr) *ntr = p.ntr;
    if (ntc) *ntc = p.ntc;
    return reo_pairs2_plan(p, out, cap);
}

/* Per-level iteration log of the last reo_identify_degs (the reference's @info lines, src:418-420, 432-435). */
int reo_iter_log(reo_handle_t h, int32_t k, int32_t* iters_done, int32_t* converged, int32_t* n_deg, int32_t* n_ref,
                 int32_t cap) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) { const int rc = reo_iter_log(h->subs[0], k, iters_done, converged, n_deg, n_ref, cap); if (rc) h->err = h->subs[0]->err; return rc; }
    if (k < 0 || k >= (int)h->logs.size()) return fail(h, REO_ERR_ARG, "reo_iter_log: no such level in the last reo_identify_degs");
    const LevelLog& L = h->logs[k];
    if (iters_done) *iters_done = L.iters_done;
    if (converged) *converged = L.converged;
    const int n = std::min<int>(cap, (int)L.n_deg.size());
    for (int i = 0; i < n; ++i) { if (n_deg) n_deg[i] = L.n_deg[i]; if (n_ref) n_ref[i] = L.n_ref[i]; }
    return REO_OK;
}

int reo_identify_degs(reo_handle_t h, const void* data, int dtype, int64_t r, int64_t c, int64_t ld,
                      const int32_t* group_id, int32_t gnum, const int32_t* thresholds, double pval_reo,
                      double pval_deg, double padj_deg, const uint8_t* ref_mask, int32_t n_iter, int32_t n_conv,
                      uint32_t flags, double* result, int8_t* updown, uint8_t* final_ref, int32_t* iters_done,
                      reo_stats* stats) {
    if (!h) return REO_ERR_ARG;
    if (!h->subs.empty()) {
        if (flags & REO_DATA_ON_DEVICE) return fail(h, REO_ERR_UNSUPPORTED, "REO_DATA_ON_DEVICE needs a single-device handle");
        if (!result || !updown || r < 1) return fail(h, REO_ERR_ARG, "reo_identify_degs: NULL argument");
        const int K = gnum == 2 ? 1 : std::max(gnum, 1);
        const size_t n = h->subs.size();
        std::vector<std::vector<double>> res(n);
        std::vector<std::vector<int8_t>> ud(n);
        std::vector<reo_stats> sts(n);
        const int rc = run_multi(h, [&](reo_handle_t s, int i) {
            double* ro = result; int8_t* uo = updown;
            if (i > 0) { res[i].resize((size_t)K * r * 15); ud[i].resize((size_t)K * r); ro = res[i].data(); uo = ud[i].data(); }
            return reo_identify_degs(s, data, dtype, r, c, ld, group_id, gnum, thresholds, pval_reo, pval_deg, padj_deg,
                                     ref_mask, n_iter, n_conv, i == 0 ? flags : (flags & ~REO_OUT_PINNED), ro, uo,
                                     i == 0 ? final_ref : nullptr,
                                     i == 0 ? iters_done : nullptr, &sts[i]);
        });
        if (rc == REO_OK && stats) {
            *stats = sts[0];
            for (size_t i = 1; i < n; ++i) {  // whole-job work; times are the slowest rank's
                stats->compares += sts[i].compares; stats->ordered_triples += sts[i].ordered_triples;
                stats->kernel_launches += sts[i].kernel_launches; stats->pair_launches += sts[i].pair_launches;
                stats->ms_total = std::max(stats->ms_total, sts[i].ms_total);
                stats->ms_pairs = std::max(stats->ms_pairs, sts[i].ms_pairs);
                stats->ms_stage = std::max(stats->ms_stage, sts[i].ms_stage);
                stats->ms_wall = std::max(stats->ms_wall, sts[i].ms_wall);
            }
        }
        return rc;
    }
    if (!data || !group_id || !ref_mask || !result || !updown) return fail(h, REO_ERR_ARG, "reo_identify_degs: NULL argument");
    if (gnum < 2) return fail(h, REO_ERR_DIM, "Only 1 level in 'group', at least 2 levels!");
    if (r <= 10) return fail(h, REO_ERR_BOUNDS, "BoundsError: r <= 10 (src:411)");
    ReoDev& D = h->devs[0];
    const auto wall0 = std::chrono::steady_clock::now();
    CK(cudaSetDevice(D.dev));
    h->kernel_launches = 0; h->pair_launches = 0; h->compares = 0; h->ordered_triples = 0;
    cudaEvent_t e_start = D.ev[0], e_staged = D.ev[1], e_end = D.ev[2];
    D.n_pev = 0;
    CK(cudaEventRecord(e_start, D.st));
    int rc = do_stage(h, data, dtype, r, c, ld, group_id, gnum, flags);
    if (rc) return rc;
    CK(cudaEventRecord(e_staged, D.st));
    const auto wall1 = std::chrono::steady_clock::now();
    const ReoStaged& S = D.S;
    const int K = gnum == 2 ? 1 : gnum;
    CK(D.result.ensure((size_t)r * 15));
    CK(D.sorted.ensure(r + r / 1024 + 2)); CK(D.sorted_p.ensure(r)); CK(D.perm.ensure(r)); CK(D.perm2.ensure(r)); CK(D.se.ensure(1)); CK(D.updown.ensure(r));
    if ((rc = ensure_std_ws(h, D))) return rc;
    reo_stats st_local;
    memset(&st_local, 0, sizeof(st_local));
    double ms_pairs = 0.0;
    // results are assembled in a pinned staging buffer so that caller outputs stay untouched on failure; with
    // REO_OUT_PINNED the caller's own (page-locked) buffers take that role
    const bool direct = (flags & REO_OUT_PINNED) != 0;
    const size_t res_bytes = (size_t)K * r * 15 * sizeof(double);
    if ((rc = ensure_h_out(h, D, res_bytes + 2 * (size_t)K * r + 64))) return rc;
    double* res_host = direct ? result : reinterpret_cast<double*>(D.h_out);
    int8_t* ud_host = direct ? updown : reinterpret_cast<int8_t*>(D.h_out + res_bytes);
    uint8_t* fr_host = (direct && final_ref) ? final_ref : D.h_out + res_bytes + (size_t)K * r;
    std::vector<int32_t> it_host(K, 0);
    D.early_pending = false;
    struct CopyGuard {   // no copy into host memory may outlive the call, whatever the exit path
        ReoDev& D;
        ~CopyGuard() { if (D.early_pending) { cudaStreamSynchronize(D.st_copy); D.early_pending = false; } }
    } copy_guard{D};

    h->logs.assign(K, LevelLog());
    for (int k = 0; k < K; ++k) {
        LevelPlan P = make_plan(S, k, thresholds, pval_reo);
        if ((rc = upload_plan(h, D, P))) return rc;
        uint8_t* mask_cur = D.mask_a.p;
        uint8_t* mask_new = D.mask_b.p;
        CK(cudaMemcpyAsync(mask_cur, ref_mask, r, cudaMemcpyHostToDevice, D.st));
        CK(cudaMemsetAsync(D.result.p, 0, (size_t)r * 15 * sizeof(double), D.st));
        int i_iter = 0, n_eval = 0, converged = 0;
        bool have_tables = false;
        while (i_iter < n_iter) {
            if (!have_tables) {   // the initial mask is a host array: count it here, no read-back
                int n0 = 0;
                for (int64_t i = 0; i < r; ++i) n0 += ref_mask[i] != 0;
                if ((rc = build_tables_full(h, D, P, mask_cur, n0))) return rc;
                have_tables = true;
            }
            if ((rc = reduce_tables(h, D))) return rc;
            static const bool timing = getenv("REO_TIMING") != nullptr;
            if (timing) CK(cudaEventRecord(D.ev[3], D.st));
            if ((rc = run_eval(h, D, r, pval_deg, padj_deg, mask_cur, mask_new, res_host + (size_t)k * r * 15))) return rc;
            if (timing) {
                CK(cudaEventRecord(D.ev[4], D.st));
                CK(cudaEventSynchronize(D.ev[4]));
                float ms = 0; cudaEventElapsedTime(&ms, D.ev[3], D.ev[4]);
                fprintf(stderr, "[reo timing]   evaluation %d: statistics sequence %.3f ms\n", n_eval, ms);
            }
            CK(cudaStreamSynchronize(D.st));
            const int n_ref = D.h_counts[0], n_inds = D.h_counts[1], n_chg = D.h_counts[2];
            if (n_eval < REO_MAX_ITER_LOG) { st_local.n_deg[n_eval] = (int32_t)r - n_inds; st_local.n_ref[n_eval] = n_ref; }
            h->logs[k].n_deg.push_back((int32_t)r - n_inds); h->logs[k].n_ref.push_back(n_ref);
            n_eval++;
            if (abs(n_ref - n_inds) < n_conv) {  // src:419-422
                converged = 1;
                break;
            }
            i_iter++;
            if (i_iter < n_iter) {
                // tables for the new reference set: signed update over the symmetric difference when
                // that is cheaper than a rebuild (the G x G categories are never stored)
                if (tables_cost(n_chg, r, S.flt) < tables_cost(n_inds, r, S.flt)) {
                    if ((rc = launch_tables(h, D, P, mask_cur, mask_new, n_chg))) return rc;
                    std::swap(mask_cur, mask_new);
                } else {
                    std::swap(mask_cur, mask_new);
                    if ((rc = build_tables_full(h, D, P, mask_cur, n_inds))) return rc;
                }
            }
        }
        if (n_eval == 0) CK(cudaMemsetAsync(D.result.p, 0, (size_t)r * 15 * sizeof(double), D.st));
        CKL(reo_launch_updown(D.result.p, r, pval_deg, padj_deg, D.updown.p, D.st));
        h->kernel_launches++;
        // columns 2..14 of the last evaluation are already on their way (enqueue_mcc): only pval and padj are left
        const size_t tail_cols = D.early_pending ? 2 : 15;
        CK(cudaMemcpyAsync(res_host + (size_t)k * r * 15, D.result.p, (size_t)r * tail_cols * 8, cudaMemcpyDeviceToHost, D.st));
        CK(cudaMemcpyAsync(ud_host + (size_t)k * r, D.updown.p, r, cudaMemcpyDeviceToHost, D.st));
        CK(cudaMemcpyAsync(fr_host + (size_t)k * r, mask_cur, r, cudaMemcpyDeviceToHost, D.st));
        CK(cudaStreamSynchronize(D.st));
        if (D.early_pending) { CK(cudaStreamSynchronize(D.st_copy)); D.early_pending = false; }
        it_host[k] = n_eval;
        st_local.iters_done = n_eval; st_local.converged = converged;
        h->logs[k].iters_done = n_eval; h->logs[k].converged = converged;
    }
    const auto wall2 = std::chrono::steady_clock::now();
    CK(cudaEventRecord(e_end, D.st));
    CK(cudaEventSynchronize(e_end));
    const auto wall3 = std::chrono::steady_clock::now();
    float ms_stage = 0, ms_total = 0;
    cudaEventElapsedTime(&ms_stage, e_start, e_staged);
    cudaEventElapsedTime(&ms_total, e_start, e_end);
    for (int i = 0; i < D.n_pev; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, D.pev[2 * i], D.pev[2 * i + 1]);
        ms_pairs += ms;
        if (getenv("REO_TIMING")) fprintf(stderr, "[reo timing] rank %d pair launch %d: %.3f ms\n", h->rank, i, (double)ms);
    }
    if (!direct) {
        memcpy(result, res_host, res_bytes);
        memcpy(updown, ud_host, (size_t)K * r);
        if (final_ref) memcpy(final_ref, fr_host, (size_t)K * r);
    }
    if (iters_done) memcpy(iters_done, it_host.data(), K * sizeof(int32_t));
    if (stats) {
        st_local.rank_bits = S.B; st_local.sample_words = S.W; st_local.compares = h->compares;
        st_local.ordered_triples = h->ordered_triples;
        st_local.planes_per_word = S.planes_per_word < 0.0 ? (double)D.h_counts[12] / std::max(1, S.W) : S.planes_per_word;
        st_local.ms_stage = ms_stage; st_local.ms_pairs = ms_pairs; st_local.ms_total = ms_total;
        st_local.ms_stats = ms_total - ms_stage - ms_pairs;
        st_local.pair_launches = h->pair_launches; st_local.kernel_launches = h->kernel_launches;
        st_local.ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
        if (getenv("REO_TIMING")) {
            auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
                return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[reo timing] stage %.3f loop %.3f endsync %.3f copyout %.3f | dev: stage %.3f pairs %.3f total %.3f\n",
                    ms(wall0, wall1), ms(wall1, wall2), ms(wall2, wall3), ms(wall3, std::chrono::steady_clock::now()),
                    (double)ms_stage, ms_pairs, (double)ms_total);
        }
        *stats = st_local;
    }
    return REO_OK;
}

}  // extern "C"
