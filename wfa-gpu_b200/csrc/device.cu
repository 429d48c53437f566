/*
 * device.cu -- per-GPU context behind the C-ABI of include/wfagpu_b200.h.
 *
 * Replaces the launch glue of the reference (lib/sequence_alignment.cu:31-470,
 * lib/sequence_packing.cu:96-116 and the device half of lib/align.cu:42-481):
 * buffers are grow-only and cached per device, nothing is memset per batch, the
 * snapshot / decision arenas are sized by the step table, rings and launch shape
 * are provisioned from the scores of the previous batch, pairs that outgrow the
 * provision or the budget are re-dispatched on the GPU with a budget picked from
 * their score bounds, and pairs with an 'N' run through the byte-compare kernels.
 * There is no CPU path.
 */
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "wfa_kernels.cuh"

using namespace wfagpu;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "[wfagpu] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_),        \
                    __FILE__, __LINE__, cudaGetErrorString(e_));                                  \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)

namespace {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n, bool keep = false, cudaStream_t s = 0)
    {
        if (n <= cap) return 0;
        /* geometric growth for small buffers only: a 40 GB arena that has to hold 41 GB must not become 60 GB */
        size_t ncap = (n * sizeof(T) > ((size_t)256 << 20)) ? n : std::max(n, cap + cap / 2);
        T *np = nullptr;
        if (!keep && p) {                      /* nothing to carry over: give the old buffer back first */
            cudaFree(p);
            p = nullptr;
            cap = 0;
        }
        CK(cudaMalloc(&np, ncap * sizeof(T)));
        if (keep && p && cap) CK(cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s));
        if (keep) CK(cudaStreamSynchronize(s));
        if (p) cudaFree(p);
        p = np;
        cap = ncap;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n)
    {
        if (n <= cap) return 0;
        size_t ncap = std::max(n, cap + cap / 2);
        T *np = nullptr;
        CK(cudaMallocHost(&np, ncap * sizeof(T)));
        if (p) cudaFreeHost(p);
        p = np;
        cap = ncap;
        return 0;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

/* WFAGPU_TRACE=1: host-side time stamps of a pass (where does the launching thread spend its time) */
struct HostTrace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    char buf[512];
    int len = 0;
    HostTrace() : on(getenv("WFAGPU_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) { buf[0] = 0; }
    void mark(const char *what)
    {
        if (!on) return;
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        len += snprintf(buf + len, sizeof(buf) - (size_t)len > 0 ? sizeof(buf) - (size_t)len : 0, " %s=%.0fus", what, us);
        if (len > (int)sizeof(buf) - 1) len = (int)sizeof(buf) - 1;
    }
    ~HostTrace() { if (on) fprintf(stderr, "[wfagpu trace]%s\n", buf); }
};

enum { CTR_QUEUE = 0, CTR_POOL = 1, CTR_RETRY = 2, CTR_ASCII = 3, CTR_TBQ = 4, CTR_BQ = 5, CTR_WORDS = 8 };

struct LaunchCfg {
    int group_threads;   /* 32 = warp per pair */
    int groups_per_cta;
    int ctas;
    int stages;
    int n_cap;
    int row_stride, center;
    int seq_words;
    size_t smem;
    int A, E1, G;
    bool global_ring;     /* large tier: rings in global memory, int32 offsets */
    bool ckpt;            /* checkpointed traceback (ring snapshots) instead of decision bytes */
    bool quad;            /* wfa_quad_kernel: four diagonals per thread, packed int16 SIMD */
    bool quad_pairs;      /* ... and two scores per barrier interval: rings one row deeper */
    int ring_m, ring_g;   /* rows of the M ring / of the I and D rings */
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool wf_timed = false;             /* ev[6], ev[7] bracket the first pass's wavefront kernel */
    cudaEvent_t ev_tab = nullptr;      /* after the last upload of a host-built table (step table, snapshot offsets) */
    bool tab_inflight = false;
    std::vector<int32_t> bound_sorted; /* scratch of the bound-first provisioning */
    std::vector<uint32_t> bound_count;
    int first_steps = 0;               /* wavefront-step budget the first pass of this batch ran with (>= plan.max_steps) */
    DevBuf<char> ascii;
    DevBuf<uint32_t> packed;
    DevBuf<wfagpu_pair_t> pairs;
    DevBuf<uint32_t> order, retry[2], ascii_list;
    DevBuf<wfagpu_pair_out_t> out;
    DevBuf<uint32_t> pool;
    DevBuf<uint32_t> counters;
    DevBuf<unsigned long long> cells;
    DevBuf<uint4> arena;
    DevBuf<uint32_t> scratch;
    DevBuf<int32_t> band_lo;
    DevBuf<int32_t> gring;
    DevBuf<char> slots, text;
    DevBuf<wfagpu_cigar_ref_t> refs;
    DevBuf<unsigned long long> heads;
    PinBuf<char> h_ascii;              /* staging for callers whose sequence buffer is pageable */
    PinBuf<char> h_text;
    PinBuf<wfagpu_cigar_ref_t> h_refs;
    PinBuf<unsigned long long> h_heads;
    DevBuf<wfagpu_step_t> steps;
    DevBuf<int32_t> bound;             /* per pair: score upper bound for the pruning */
    DevBuf<uint32_t> ck_off;           /* arena offset of every ring snapshot (checkpointed traceback) */
    std::vector<uint64_t> h_ck_off;    /* [j] = units used by the snapshots of scores < j * period */
    PinBuf<uint32_t> h_ck32;
    PinBuf<int32_t> h_bound;           /* re-dispatch: bounds and the retry list on the host */
    PinBuf<uint32_t> h_retry;
    int ck_key[1] = {-1};              /* period the offsets were built for */
    PinBuf<wfagpu_pair_t> h_pairs;
    PinBuf<uint32_t> h_order;
    PinBuf<uint32_t> h_seed;           /* one word: the re-dispatch counter a first pass starts from */
    PinBuf<wfagpu_pair_out_t> h_out;
    PinBuf<uint32_t> h_pool;
    PinBuf<uint32_t> h_counters;
    PinBuf<unsigned long long> h_cells;
    PinBuf<wfagpu_step_t> h_steps;
    int tab_key[5] = {-1, -1, -1, -1, -1};
    int tab_d_end = 0;
    uint64_t tab_arena_units = 0;
    size_t n = 0;
    size_t ascii_bytes = 0;
    size_t packed_words = 0;
    uint32_t max_len = 0;
    uint32_t kt_max = 0;               /* largest |tlen - plen| of the batch */
    uint32_t minlen_max = 0;           /* largest min(plen, tlen) of the batch */
    size_t budget_cache = 0;           /* arena budget from the last cudaMemGetInfo (0 = ask again) */
    wfagpu_plan_t plan{};
    wfagpu_batch_stats_t stats{};
    bool have_events = false;
    bool text_queued = false; /* cigar_text kernels already follow the last alignment pass */
    int last_d_end = 0;
    bool capped = false;      /* the first pass provisioned fewer diagonals than the budget allows */
};

} // namespace

struct wfagpu_device {
    int dev = 0;
    cudaDeviceProp prop{};
    Slot slots[2];
    bool count_cells = false;
    int force_threads = 0, force_stages = 0, force_ctas_per_sm = 0, force_warp = -1;
    bool no_ckpt = false, no_bound = false, force_bound = false, no_quad = false, no_quad_pairs = false, no_band_tb = false, no_prebound = false, no_bound_order = false, no_skip_open = false;
    int force_period = 0;
    int arena_mb = 0;
    /* largest score seen in the last batch, per penalty set: sizes the rings of the next first pass */
    int hint_dist = 0;
    double hint_mean = 0;              /* mean score of the finished pairs of that batch */
    int hint_key[4] = {-1, -1, -1, -1};   /* x, o, e, length class (log2 of the longest sequence) */
    bool use_hint = true;
    bool force_large = false;
    bool device_text = true;   /* WFAGPU_HOST_CIGAR=1 leaves the text to the host */
    bool leased = false;       /* handed out by wfagpu_device_open and not released yet */
    bool independent = false;  /* wfagpu_device_rescore in progress: do not learn hints from it */
    int quad_min = 256;        /* smallest ring half width that runs the four-diagonals-per-thread kernel */
    int hint_margin_pm = 83, hint_min_pm = 31;   /* provisioning margins over the last batch's largest score, per mille */
    int hint_lift_pm = 250;    /* the first pass may exceed the -e budget by this much when the last batch needed it (0 = never) */
    int max_steps_cap = 60000; /* most wavefront steps a pair may take (WFAGPU_MAX_STEPS_CAP lowers it: tests) */
};

/* Contexts are leased: wfagpu_device_open hands out an idle context of that GPU (or makes a new one), so two
 * host threads -- or two workers of one call on the same GPU ("0,0") -- never share streams, slots or staging
 * buffers; wfagpu_device_release returns it to the pool with its grown buffers.  Only the provisioning hint
 * (largest / mean score of the last batch per penalty set) is shared per GPU, under g_mu. */
struct DevShared {
    int dev = -1;
    int hint_dist = 0;
    double hint_mean = 0;
    int hint_key[4] = {-1, -1, -1, -1};
};
static std::mutex g_mu;
static std::vector<wfagpu_device *> g_devices;
static std::vector<DevShared> g_shared;

/* a hint learnt on 150 bp reads says nothing about 10 kbp reads: hints are only used within a length class */
static int length_class(uint32_t max_len)
{
    int c = 0;
    while (max_len > 1) { max_len >>= 1; ++c; }
    return c;
}

static DevShared &shared_of(int dev)          /* g_mu held */
{
    for (auto &h : g_shared) if (h.dev == dev) return h;
    DevShared h;
    h.dev = dev;
    g_shared.push_back(h);
    return g_shared.back();
}
static void pull_hint(wfagpu_device *d)
{
    std::lock_guard<std::mutex> lk(g_mu);
    const DevShared &h = shared_of(d->dev);
    d->hint_dist = d->use_hint ? h.hint_dist : 0;
    d->hint_mean = h.hint_mean;
    for (int i = 0; i < 4; ++i) d->hint_key[i] = h.hint_key[i];
}
static void push_hint(wfagpu_device *d)
{
    std::lock_guard<std::mutex> lk(g_mu);
    DevShared &h = shared_of(d->dev);
    h.hint_dist = d->hint_dist;
    h.hint_mean = d->hint_mean;
    for (int i = 0; i < 4; ++i) h.hint_key[i] = d->hint_key[i];
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

extern "C" wfagpu_device_t *wfagpu_device_open(int dev)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto *d : g_devices)
        if (d->dev == dev && !d->leased) {
            cudaSetDevice(dev);
            d->leased = true;
            return d;
        }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || dev < 0 || dev >= count) {
        fprintf(stderr, "[wfagpu] CUDA device %d is not available (%s); this library has no CPU fallback\n", dev,
                e != cudaSuccess ? cudaGetErrorString(e) : "out of range");
        return nullptr;
    }
    if (cudaSetDevice(dev) != cudaSuccess) return nullptr;
    wfagpu_device *d = new wfagpu_device();
    d->dev = dev;
    if (cudaGetDeviceProperties(&d->prop, dev) != cudaSuccess) {
        delete d;
        return nullptr;
    }
    if (d->prop.major < 10) {
        fprintf(stderr, "[wfagpu] device %d is sm_%d%d; this build carries sm_100a code only\n", dev, d->prop.major,
                d->prop.minor);
        delete d;
        return nullptr;
    }
    for (auto &s : d->slots) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete d;
            return nullptr;
        }
        for (auto &ev : s.ev) cudaEventCreate(&ev);
        cudaEventCreateWithFlags(&s.ev_tab, cudaEventDisableTiming);
    }
    d->count_cells = env_int("WFAGPU_COUNT_CELLS", 0) != 0;
    d->force_threads = env_int("WFAGPU_THREADS", 0);
    d->force_stages = env_int("WFAGPU_STAGES", 0);
    d->force_ctas_per_sm = env_int("WFAGPU_CTAS_PER_SM", 0);
    d->force_warp = env_int("WFAGPU_WARP_KERNEL", -1);
    d->no_ckpt = env_int("WFAGPU_NO_CKPT", 0) != 0;
    d->no_bound = env_int("WFAGPU_NO_BOUND", 0) != 0;
    d->no_quad = env_int("WFAGPU_NO_QUAD", 0) != 0;
    d->no_band_tb = env_int("WFAGPU_NO_BAND_TB", 0) != 0;
    d->no_prebound = env_int("WFAGPU_NO_PREBOUND", 0) != 0;
    d->no_bound_order = env_int("WFAGPU_NO_BOUND_ORDER", 0) != 0;
    d->no_skip_open = env_int("WFAGPU_NO_SKIP_OPEN", 0) != 0;
    /* two scores per barrier interval: measured equal to one (22.9 vs 22.8 ms per 8192 x 10 kbp pairs) because the deeper
     * rings cost the fifth resident CTA; opt-in */
    d->no_quad_pairs = env_int("WFAGPU_QUAD_PAIRS", 0) == 0;
    d->force_bound = env_int("WFAGPU_FORCE_BOUND", 0) != 0;
    d->arena_mb = env_int("WFAGPU_ARENA_MB", 0);
    d->force_period = env_int("WFAGPU_CK_PERIOD", 0);
    if (d->force_period != 7 && d->force_period != 15 && d->force_period != 31) d->force_period = 0;
    d->use_hint = env_int("WFAGPU_NO_HINT", 0) == 0;
    d->force_large = env_int("WFAGPU_FORCE_LARGE", 0) != 0;
    d->device_text = env_int("WFAGPU_HOST_CIGAR", 0) == 0;
    d->quad_min = std::max(1, env_int("WFAGPU_QUAD_MIN", 256));
    d->hint_margin_pm = std::max(0, env_int("WFAGPU_HINT_MARGIN_PM", 83));
    d->hint_min_pm = std::min(d->hint_margin_pm, std::max(0, env_int("WFAGPU_HINT_MIN_PM", 31)));
    d->hint_lift_pm = std::max(0, env_int("WFAGPU_HINT_LIFT_PM", 250));
    d->max_steps_cap = std::min(60000, std::max(16, env_int("WFAGPU_MAX_STEPS_CAP", 60000)));
    d->leased = true;
    g_devices.push_back(d);
    return d;
}

extern "C" void wfagpu_device_release(wfagpu_device_t *d)
{
    if (!d) return;
    std::lock_guard<std::mutex> lk(g_mu);
    d->leased = false;
}

extern "C" void wfagpu_device_close_all(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto *d : g_devices) {
        cudaSetDevice(d->dev);
        for (auto &s : d->slots) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            s.ascii.release(); s.packed.release(); s.pairs.release(); s.order.release();
            s.retry[0].release(); s.retry[1].release(); s.ascii_list.release(); s.out.release();
            s.pool.release(); s.counters.release(); s.cells.release(); s.arena.release(); s.bound.release(); s.ck_off.release(); s.h_ck32.release(); s.h_bound.release(); s.h_retry.release();
            s.scratch.release(); s.steps.release(); s.band_lo.release(); s.gring.release(); s.slots.release(); s.text.release(); s.refs.release(); s.heads.release();
            s.h_text.release(); s.h_refs.release(); s.h_heads.release(); s.h_ascii.release();
            s.h_pairs.release(); s.h_order.release(); s.h_seed.release(); s.h_out.release(); s.h_pool.release();
            s.h_counters.release(); s.h_cells.release(); s.h_steps.release();
            for (auto &ev : s.ev) if (ev) cudaEventDestroy(ev);
            if (s.ev_tab) cudaEventDestroy(s.ev_tab);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
        delete d;
    }
    g_devices.clear();
    g_shared.clear();
}

extern "C" int wfagpu_device_sm_count(wfagpu_device_t *d) { return d ? d->prop.multiProcessorCount : 0; }

static inline uint32_t packed_words_for(uint32_t len) { return ((((len + 7u) >> 3) + 1u) + 3u) & ~3u; }

extern "C" int wfagpu_device_upload(wfagpu_device_t *d, int slot, const char *ascii, size_t ascii_bytes,
                                    const wfagpu_pair_t *pairs, size_t n)
{
    if (!d || slot < 0 || slot > 1) return -1;
    CK(cudaSetDevice(d->dev));
    Slot &s = d->slots[slot];
    s.n = n;
    s.ascii_bytes = ascii_bytes;
    memset(&s.stats, 0, sizeof(s.stats));
    if (n == 0) return 0;
    if (s.h_pairs.ensure(n) || s.h_order.ensure(n)) return -1;
    /* packed layout + longest-first schedule */
    size_t words = 0;
    uint32_t max_len = 0, kt_max = 0, minlen_max = 0;
    uint64_t min_sum = ~0ull, max_sum = 0;
    for (size_t i = 0; i < n; ++i) {
        wfagpu_pair_t p = pairs[i];
        const uint64_t sum = (uint64_t)p.plen + p.tlen;
        kt_max = std::max(kt_max, p.plen > p.tlen ? p.plen - p.tlen : p.tlen - p.plen);
        minlen_max = std::max(minlen_max, std::min(p.plen, p.tlen));
        min_sum = std::min(min_sum, sum);
        max_sum = std::max(max_sum, sum);
        p.p_word = (uint32_t)words;
        words += packed_words_for(p.plen);
        p.t_word = (uint32_t)words;
        words += packed_words_for(p.tlen);
        p.flags = 0;
        s.h_pairs.p[i] = p;
        s.h_order.p[i] = (uint32_t)i;
        max_len = std::max(max_len, std::max(p.plen, p.tlen));
    }
    if (words >= (1ull << 32)) {
        fprintf(stderr, "[wfagpu] batch too large for 32-bit packed offsets; use a smaller batch_size\n");
        return -1;
    }
    s.packed_words = words;
    s.max_len = max_len;
    s.kt_max = kt_max;
    s.minlen_max = minlen_max;
    if ((uint64_t)min_sum * 8 < (uint64_t)max_sum * 7) {
        /* longest first only pays when the lengths differ by more than ~12 %: for uniform
         * reads the queue order is irrelevant and the sort would dominate the host time */
        const wfagpu_pair_t *hp = s.h_pairs.p;
        std::stable_sort(s.h_order.p, s.h_order.p + n, [hp](uint32_t a, uint32_t b) {
            return (uint64_t)hp[a].plen + hp[a].tlen > (uint64_t)hp[b].plen + hp[b].tlen;
        });
    }
    if (s.ascii.ensure(ascii_bytes + 64) || s.packed.ensure(words + 16) || s.pairs.ensure(n) || s.order.ensure(n) ||
        s.retry[0].ensure(n) || s.retry[1].ensure(n) || s.ascii_list.ensure(n) || s.out.ensure(n) ||
        s.counters.ensure(CTR_WORDS) || s.h_counters.ensure(CTR_WORDS) || s.cells.ensure(1) || s.h_cells.ensure(1))
        return -1;
    CK(cudaEventRecord(s.ev[0], s.stream));
    CK(cudaMemcpyAsync(s.ascii.p, ascii, ascii_bytes, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemsetAsync(s.ascii.p + ascii_bytes, 0, 64, s.stream));
    CK(cudaMemcpyAsync(s.pairs.p, s.h_pairs.p, n * sizeof(wfagpu_pair_t), cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.order.p, s.h_order.p, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
    CK(cudaEventRecord(s.ev[1], s.stream));
    s.stats.h2d_bytes += ascii_bytes + n * (sizeof(wfagpu_pair_t) + sizeof(uint32_t));
    s.have_events = true;
    return 0;
}

/* Shared-memory / launch-shape policy (replaces available_shared_mem_per_block and
 * the <<<num_workers, tpb>>> choice of lib/sequence_alignment.cu:81-108,211-330).
 *
 * The wavefront rings live in shared memory, so their width decides how many CTAs
 * share an SM.  Measured on B200 (10 kbp / 5 %): 1 CTA x 1024 threads 62 k pairs/s,
 * 2 x 512 95 k, 3 x 384 109 k -- more, smaller CTAs hide the per-score barrier.
 * `n_want` is the half width the launch should be able to hold. */
static int choose_cfg(wfagpu_device *d, int x, int o, int e, int n_want, int n_min, uint32_t max_len, size_t n_items,
                      bool ascii, bool bt, LaunchCfg *c)
{
    const int A = std::max(o + e, x) + 1, E1 = e + 1, G = A;
    const size_t smem_max = d->prop.sharedMemPerBlockOptin;        /* 227 KB on B200 */
    const size_t smem_sm = d->prop.sharedMemPerMultiprocessor;     /* 228 KB */
    const int seq_words = (int)packed_words_for(max_len);
    c->A = A; c->E1 = E1; c->G = G; c->seq_words = seq_words;
    n_want = std::max(1, n_want);
    /* ring row geometry.  Checkpointed traceback copies rows in 16-byte units: diagonal 0 sits on
     * an 8-element boundary and a row has 8 spare cells on both sides. */
    bool ck = false;       /* aligned row geometry: snapshots (16-byte units) and the quad kernel (LDS.64) need it */
    auto ctr = [&](int ncap) { return ck ? ((ncap + 2 * G + 8 + 7) & ~7) : ncap + 2 * G + 1; };
    auto rs = [&](int ncap) { return ck ? 2 * ctr(ncap) : 2 * (ncap + 2 * G + 1) + 2; };

    c->global_ring = false;
    c->ckpt = false;
    c->quad = false;
    c->quad_pairs = false;
    c->ring_m = A; c->ring_g = E1;
    bool large = d->force_large || max_len >= (1u << 15);
    if (!large) {
        /* rings that do not fit one CTA's shared memory even single-buffered go to the large tier */
        if (exact_smem_bytes(A, E1, rs(n_want), seq_words, 1, 1, true) > smem_max) large = true;
    }
    if (large) {
        c->global_ring = true;
        c->groups_per_cta = 1;
        c->stages = 1;
        c->n_cap = n_want;
        /* packed pairs: four diagonals per thread with 128-bit loads (wfa_quadg_kernel): rows 16-byte aligned */
        c->quad = !ascii && !d->no_quad;
        ck = c->quad;
        c->row_stride = rs(n_want);
        c->center = c->quad ? ctr(n_want) : n_want + 2 * G + 1;
        c->smem = exact_smem_bytes(A, E1, 0, seq_words, 1, 1, c->quad);     /* sequences + control (+ schedule records) only */
        if (!ascii && c->smem > smem_max) {
            fprintf(stderr, "[wfagpu] sequences of %u bases do not fit in shared memory; not supported yet\n", max_len);
            return -2;
        }
        int occ;
        if (c->quad) {
            /* one CTA per SM -- 148 x 0.7 MB of rings stay resident in the 126 MB L2 (two per SM: 681 instead of 902 pairs/s
             * at 50 kbp / 15 %); measured on B200, 592 pairs: 512 threads 1030, 768 1077, 1024 (64 registers) 1098 pairs/s */
            c->group_threads = d->force_threads ? d->force_threads : (n_want >= 2048 ? 1024 : (n_want >= 512 ? 512 : 256));
            occ = quadg_max_ctas_per_sm(c->group_threads, c->smem, bt);
            if (occ < 1) return -1;
            occ = std::min(occ, d->force_ctas_per_sm ? d->force_ctas_per_sm : 1);
        } else {
            /* wide wavefronts (50 kbp / 15 %: ~10 k diagonals per score) keep 1024 threads busy: 581 vs 416
             * pairs/s against 512 threads on B200 */
            c->group_threads = d->force_threads ? d->force_threads : (n_want >= 2048 ? 1024 : 512);
            occ = large_max_ctas_per_sm(c->group_threads, c->smem, ascii, bt);
            if (occ < 1) return -1;
            occ = std::min(occ, d->force_ctas_per_sm ? d->force_ctas_per_sm : 2);
        }
        c->ctas = (int)std::max<size_t>(1, std::min<size_t>((size_t)occ * d->prop.multiProcessorCount, n_items));
        return 0;
    }
    bool warp = (rs(n_want) <= 400) && max_len <= 1024;
    if (d->force_warp >= 0) warp = d->force_warp != 0;

    if (warp) {
        c->group_threads = 32;
        c->groups_per_cta = 8;
        c->stages = 2;
        c->n_cap = n_want;
        c->row_stride = rs(c->n_cap);
        c->center = c->n_cap + 2 * G + 1;
        c->smem = exact_smem_bytes(A, E1, c->row_stride, seq_words, c->groups_per_cta, c->stages);
        while (c->smem > smem_max && c->groups_per_cta > 1) {
            c->groups_per_cta >>= 1;
            c->smem = exact_smem_bytes(A, E1, c->row_stride, seq_words, c->groups_per_cta, c->stages);
        }
        if (c->smem > smem_max) warp = false;
    }
    if (!warp) {
        c->groups_per_cta = 1;
        c->ckpt = bt && !d->no_ckpt;
        /* packed pairs on shared-memory rings run four diagonals per thread (with backtrace: snapshots only) */
        /* (measured on B200: 1 kbp / 10 %, ring half width 196: 9.5 ms per 50 k pairs with one diagonal per thread and 16 CTAs per
         * SM against 11.0 ms; from a half width of ~300 on the four-diagonal kernel wins) */
        c->quad = !ascii && !d->no_quad && (!bt || c->ckpt) && n_want >= d->quad_min;
        /* two scores per barrier: score d + 1 must not read the extended M of score d (x >= 2, o + e >= 2) and takes
         * its extend sources from registers (e == 1); costs one more row per ring */
        c->quad_pairs = c->quad && !d->no_quad_pairs && x >= 2 && o + e >= 2 && e == 1;
        ck = c->ckpt || c->quad;
        const int RA = A + (c->quad_pairs ? 1 : 0), RE = E1 + (c->quad_pairs ? 1 : 0);
        c->ring_m = RA; c->ring_g = RE;
        int best_k = 0, stages = 1, n_cap = n_want;
        size_t smem = 0;
        /* CTA size follows the average width of the pruned wavefront (~ n_want): a quarter of it, 64..192
         * threads, and as many CTAs per SM as give ~960 threads (measured on B200: 10 kbp / 5 % is flat from
         * 128 to 192 threads at 5 CTAs; 1 kbp / 10 % runs 17.2 ms with 6 x 160 and 13.1 ms with 15 x 64) */
        const int t_pref = std::min(192, std::max(64, (((n_want + 3) / 4) + 31) & ~31));
        const int kmax = d->force_ctas_per_sm ? d->force_ctas_per_sm : std::min(16, std::max(1, 960 / t_pref));
        /* n_min <= n_want: the rings may be provisioned below n_want (down to n_min) when that buys
         * another resident CTA -- the few pairs that then outgrow them are re-dispatched */
        n_min = std::max(1, std::min(n_min, n_want));
        /* do k CTAs of the smallest shape really fit with this much dynamic shared memory?  The arithmetic budget below is
         * a first filter; the runtime's occupancy calculator decides (allocation granularity: a ring 6 diagonals wider than
         * the arithmetic allowed cost rank 3 of an 8-GPU run its fifth CTA per SM, 33.4 instead of 29.2 ms per step) */
        auto fits = [&](int k, size_t bytes, size_t budget) {
            if (bytes > budget) return false;
            const int occ_k = c->quad ? quad_max_ctas_per_sm(64, bytes, bt) : exact_max_ctas_per_sm(64, 1, bytes, ascii, bt, c->ckpt);
            return occ_k >= k;
        };
        for (int k = kmax; k >= 1 && !best_k; --k) {
            /* every resident CTA also reserves 1 KB of system shared memory */
            const size_t budget = std::min(smem_max, smem_sm / k - 1024);
            for (int st = 2; st >= 1; --st) {
                if (d->force_stages && st != d->force_stages) continue;
                if (!fits(k, exact_smem_bytes(RA, RE, rs(n_min), seq_words, 1, st, true), budget)) continue;
                int lo = n_min, hi = n_want;             /* widest rings that still fit k CTAs */
                while (lo < hi) {
                    const int mid = (lo + hi + 1) / 2;
                    if (fits(k, exact_smem_bytes(RA, RE, rs(mid), seq_words, 1, st, true), budget)) lo = mid; else hi = mid - 1;
                }
                best_k = k; stages = st; n_cap = lo;
                smem = exact_smem_bytes(RA, RE, rs(n_cap), seq_words, 1, st, true);
                break;
            }
        }
        if (!best_k) {
            /* does not fit even alone: hold as many diagonals as one CTA can */
            stages = 1;
            const size_t fixed = exact_smem_bytes(RA, RE, 0, seq_words, 1, stages, true);
            const int rows_q = RA + 2 * RE;
            if (fixed + (size_t)rows_q * rs(1) * 2 > smem_max) return -2;   /* the sequences alone do not fit */
            n_cap = (int)((smem_max - fixed - 64) / ((size_t)rows_q * 4)) - 2 * G - (ck ? 16 : 2);
            if (n_cap < 1) return -2;
            smem = exact_smem_bytes(RA, RE, rs(n_cap), seq_words, 1, stages, true);
            best_k = 1;
        }
        c->stages = stages;
        c->n_cap = n_cap;
        c->row_stride = rs(n_cap);
        c->center = ctr(n_cap);
        c->smem = smem;
        /* about 1152 threads per SM in total, never more threads than half the widest wavefront */
        /* (measured on B200: 3 x 384, 5 x 192 -- the per-score overhead is paid per warp) */
        int t = best_k == 1 ? 1024 : (best_k == 2 ? 512 : (best_k == 3 ? 384 : ((960 / best_k) / 32) * 32));
        const int width = 2 * n_cap + 1;
        if (c->quad) {
            /* a thread covers four diagonals per trip: one to two trips over the average window (~ n_cap);
             * the kernel needs ~80 registers: never so many threads that the registers cost a resident CTA */
            t = std::min(std::min(t, 512), std::max(64, ((n_cap / 6) + 31) & ~31));
            while (t > 64 && !d->force_threads && quad_max_ctas_per_sm(t, smem, bt) < best_k) t -= 32;
        } else {
            while (t > 64 && t * 2 > width) t -= 32;
        }
        t = std::max(64, t);
        if (d->force_threads) t = d->force_threads;
        c->group_threads = t;
    }
    const int threads = warp ? 32 * c->groups_per_cta : c->group_threads;
    int occ = c->quad ? quad_max_ctas_per_sm(c->group_threads, c->smem, bt)
                      : exact_max_ctas_per_sm(c->group_threads, c->groups_per_cta, c->smem, ascii, bt, c->ckpt);
    if (occ < 1) {
        fprintf(stderr, "[wfagpu] kernel configuration does not fit (threads=%d smem=%zu)\n", threads, c->smem);
        return -1;
    }
    if (d->force_ctas_per_sm) occ = std::min(occ, d->force_ctas_per_sm);
    size_t ctas = (size_t)occ * d->prop.multiProcessorCount;
    const size_t need = (n_items + c->groups_per_cta - 1) / c->groups_per_cta;
    c->ctas = (int)std::max<size_t>(1, std::min(ctas, need));
    return 0;
}

/* Remember the largest score of the batch whose results sit in s.h_out: the next first
 * pass with the same penalties provisions its rings for that (plus a margin). */
static void learn_hint(wfagpu_device *d, Slot &s, size_t n)
{
    int dmax = 0;
    double sum = 0;
    size_t cnt = 0;
    for (size_t i = 0; i < n; ++i)
        if (s.h_out.p[i].status & WFAGPU_ST_FINISHED) {
            dmax = std::max(dmax, s.h_out.p[i].distance);
            sum += s.h_out.p[i].distance;
            ++cnt;
        }
    d->hint_dist = d->use_hint ? dmax : 0;
    d->hint_mean = cnt ? sum / (double)cnt : 0;
    d->hint_key[0] = s.plan.x; d->hint_key[1] = s.plan.o; d->hint_key[2] = s.plan.e; d->hint_key[3] = length_class(s.max_len);
    push_hint(d);
}

/* The host-built tables are uploaded from pinned memory that the next pass rewrites.  Waiting for the last such upload
 * (an event) is enough; waiting for the whole stream would also wait for the chunk's own sequence upload and pack, which
 * a first pass has just queued (0.5 ms per 12500 x 1 kbp chunk whenever first pass and re-dispatch budgets alternate). */
static int tables_idle(Slot &s)
{
    if (!s.tab_inflight) return 0;
    if (s.ev_tab) CK(cudaEventSynchronize(s.ev_tab)); else CK(cudaStreamSynchronize(s.stream));
    s.tab_inflight = false;
    return 0;
}
static int tables_sent(Slot &s)
{
    if (s.ev_tab) CK(cudaEventRecord(s.ev_tab, s.stream));
    s.tab_inflight = true;
    return 0;
}

/* Ring-snapshot layout of the checkpointed traceback for period P: snapshot j (score j * P) holds
 * (A-1) + 2e rows of the 16-byte units that cover [-n, n] at that score.  h_ck_off[j] = units used
 * by the snapshots before j. */
static int ensure_ck_table(Slot &s, const wfagpu_plan_t &plan, int period)
{
    if (s.ck_key[0] == period) return 0;
    const int de = s.tab_d_end;
    const int A = std::max(plan.o + plan.e, plan.x) + 1;
    const uint64_t rows_ck = (uint64_t)(A - 1 + 2 * plan.e);
    const int n_ck = (de - 1) / period + 2;
    s.h_ck_off.assign((size_t)n_ck + 1, 0);
    if (s.h_ck32.ensure((size_t)n_ck + 1)) return -1;
    if (tables_idle(s)) return -1;            /* an earlier pass may still be reading the pinned copy */
    uint64_t acc = 0;
    s.h_ck32.p[0] = 0;
    for (int j = 1; j <= n_ck; ++j) {
        s.h_ck_off[j] = acc;
        s.h_ck32.p[j] = (uint32_t)std::min<uint64_t>(acc, 0xffffffffu);
        const long long dj = (long long)j * period;
        if (dj < de) {
            const int n = s.h_steps.p[dj].n;
            acc += rows_ck * (uint64_t)((((n + 7) & ~7) + ((n + 8) & ~7)) >> 3);
        }
    }
    if (s.ck_off.ensure((size_t)n_ck + 1)) return -1;
    CK(cudaMemcpyAsync(s.ck_off.p, s.h_ck32.p, ((size_t)n_ck + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
    if (tables_sent(s)) return -1;
    s.ck_key[0] = period;
    return 0;
}

/* Step table for a budget of `max_steps` wavefront steps (cached per slot; host copy in s.h_steps). */
static int ensure_step_table(Slot &s, const wfagpu_plan_t &plan, int max_steps)
{
    const int max_dist = std::min<long long>((long long)max_steps * (std::max(plan.x, plan.o + plan.e) + 1) + 16, 1 << 30);
    const int tab_win = plan.band > 0 ? (plan.band_width > 0 ? plan.band_width : 512) : 0;
    if (s.tab_key[0] == plan.x && s.tab_key[1] == plan.o && s.tab_key[2] == plan.e && s.tab_key[3] == max_steps &&
        s.tab_key[4] == tab_win)
        return 0;
    if (s.h_steps.ensure((size_t)max_dist + 1)) return -1;
    if (tables_idle(s)) return -1;        /* a previous pass may still be reading the pinned table */
    uint64_t units = 0;
    const int de = wfagpu_build_step_table(plan.x, plan.o, plan.e, max_steps, max_dist, tab_win, s.h_steps.p, &units);
    if (de < 1) return -1;
    if (s.steps.ensure((size_t)de + 1)) return -1;
    CK(cudaMemcpyAsync(s.steps.p, s.h_steps.p, (size_t)de * sizeof(wfagpu_step_t), cudaMemcpyHostToDevice, s.stream));
    if (tables_sent(s)) return -1;
    s.tab_key[0] = plan.x; s.tab_key[1] = plan.o; s.tab_key[2] = plan.e; s.tab_key[3] = max_steps; s.tab_key[4] = tab_win;
    s.tab_d_end = de;
    s.tab_arena_units = units;
    s.ck_key[0] = -1;     /* snapshot offsets follow the table */
    return 0;
}

/* Score upper bounds only (wfa_bound_kernel) for `n_items` entries of `order_dev`, with a budget of
 * `max_steps`: the re-dispatch loop uses them to pick a budget that is enough in one go and to prune
 * that pass per pair.  Leaves the step table of `max_steps` cached. */
static int run_bound_only(wfagpu_device *d, Slot &s, const wfagpu_plan_t &plan, int max_steps, const uint32_t *order_dev,
                          size_t n_items)
{
    if (ensure_step_table(s, plan, max_steps)) return -1;
    if (s.bound.ensure(s.n + 1)) return -1;
    KernelParams kp{};
    kp.packed = s.packed.p;
    kp.pairs = s.pairs.p;
    kp.order = order_dev;
    kp.n_items = (uint32_t)n_items;
    kp.steps = s.steps.p;
    kp.d_end = s.tab_d_end;
    kp.x = plan.x; kp.o = plan.o; kp.e = plan.e;
    kp.A = std::max(plan.o + plan.e, plan.x) + 1;
    kp.E1 = plan.e + 1;
    kp.bound = s.bound.p;
    kp.bound_queue = s.counters.p + CTR_BQ;
    constexpr int kWarps = 8;
    const int occ = bound_max_ctas_per_sm(kp.A, kp.E1, kWarps);
    if (occ < 1) return -2;
    CK(cudaMemsetAsync(s.counters.p + CTR_BQ, 0, sizeof(uint32_t), s.stream));
    const int ctas = (int)std::min<size_t>((size_t)occ * d->prop.multiProcessorCount, (n_items + kWarps - 1) / kWarps);
    cudaError_t e = launch_bound(kp, std::max(1, ctas), kWarps, s.stream);
    if (e != cudaSuccess) {
        fprintf(stderr, "[wfagpu] bound kernel launch failed: %s\n", cudaGetErrorString(e));
        return -1;
    }
    s.stats.launches += 1;
    return 0;
}

/* Launch one pass over `n_items` entries of `order_dev`. */
static int launch_pass(wfagpu_device *d, Slot &s, const wfagpu_plan_t &plan, int max_steps, const uint32_t *order_dev,
                       size_t n_items, uint32_t *retry_dev, bool ascii, bool first_pass, bool use_hint,
                       bool *capped_out, bool have_bounds = false)
{
    HostTrace tr;
    if (first_pass) pull_hint(d);            /* what the last batch on this GPU needed (any context) */
    if (ensure_step_table(s, plan, max_steps)) return -1;
    const wfagpu_step_t *tab = s.h_steps.p;
    const int d_full = s.tab_d_end;
    int n_full = 0;
    for (int dd = d_full - 1; dd >= 0 && dd >= d_full - 4; --dd) n_full = std::max(n_full, (int)tab[dd].n);
    /* Scores to provision for: the full budget, or (first pass only) what recent batches with
     * these penalties needed plus a margin -- pairs that outgrow it are re-dispatched.  The kernel
     * prunes every cell that cannot reach the target diagonal within d_end - 1 (score-bound
     * pruning), so the rings only have to hold max_d min(n_d, kt_max + (d_end - 1 - d) / e). */
    int d_want = d_full;
    bool hinted = false;
    if (use_hint && plan.band <= 0 && d->hint_dist > 0 && d->hint_key[0] == plan.x && d->hint_key[1] == plan.o && d->hint_key[2] == plan.e &&
        d->hint_key[3] == length_class(s.max_len)) {
        d_want = (int)std::min<long long>((long long)d->hint_dist + (long long)d->hint_dist * d->hint_margin_pm / 1000 + 8, d_full - 1) + 1;
        hinted = d_want < d_full;
    }
    /* no pair of the batch can score more than substitutions all along the shorter sequence plus one
     * gap for the length difference: a generous -e (or a stale hint) does not widen the rings */
    int d_reach = d_full;                    /* scores beyond this cannot occur in this batch */
    if (plan.band <= 0) {
        const long long cap = (long long)plan.x * s.minlen_max + plan.o + (long long)plan.e * s.kt_max + 2;
        if (cap < d_reach) d_reach = (int)std::max<long long>(cap, 2);
        if (d_reach < d_want) d_want = d_reach;
    }
    /* Packed exact first pass where the per-pair score bounds pay (long reads): the bound kernel runs first and its
     * largest bound sizes the rings -- what this batch needs, not what the last one needed, so a first call, a call
     * without hints and a call after a batch of another kind all run at the speed of a well-hinted one.  Costs one
     * host round trip per pass (bounds back: 4 bytes per pair). */
    int d_p99 = 0;                           /* 99 % of the bounded pairs finish below this score (0 = unknown) */
    size_t n_seed = 0;                       /* pairs at the end of the order list that skip this pass (no bound within the budget) */
    if (first_pass && !have_bounds && plan.band <= 0 && !ascii && !d->no_bound && !d->no_prebound && n_items == s.n &&
        bound_max_ctas_per_sm(std::max(plan.o + plan.e, plan.x) + 1, plan.e + 1, 8) >= 1) {   /* penalties whose rings fit the bound kernel */
        bool pays = d->force_bound;
        if (!pays) {
            const bool h = d_want < d_full && d->hint_mean > 0 && d->hint_dist > 0;
            const double m = h ? d->hint_mean : 0.5 * (d_want - 1);
            const double r = std::min(2.0, std::max(1.0, (d_want - 1) / std::max(1.0, m)));
            const double cells = r * r / 4 + (1.5 * r - 1) * (1 - r / 2);
            pays = (cells - 0.5) * m > 130.0;
        }
        if (!pays && d_want < d_full && d->hint_dist > 0 && s.max_len >= 2000 && n_items >= 1024 && order_dev == s.order.p) {
            /* a hint tight enough that bounding every pair does not pay: is it still true for this batch?  Bound the first
             * 128 pairs of the queue; if they already outgrow the hint (every pair would be re-dispatched) or stay far
             * below it (rings wider than needed), provision from this batch's own bounds after all */
            constexpr size_t kSample = 128;
            int rcb = run_bound_only(d, s, plan, max_steps, order_dev, kSample);
            if (rcb) return rcb;
            if (s.h_bound.ensure(s.n + 1)) return -1;
            CK(cudaMemcpyAsync(s.h_bound.p, s.bound.p, s.n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            long long smax = 0;
            for (size_t i = 0; i < kSample; ++i) {
                const uint32_t idx = s.h_order.p[i];
                if (idx < s.n) smax = std::max<long long>(smax, std::min<int32_t>(s.h_bound.p[idx], d_full));
            }
            const long long covered = (long long)d->hint_dist + (long long)d->hint_dist * d->hint_min_pm / 1000;
            pays = smax > covered || smax * 13 < (long long)d->hint_dist * 10;
        }
        if (pays) {
            int rcb = run_bound_only(d, s, plan, max_steps, order_dev, n_items);
            if (rcb) return rcb;
            if (s.h_bound.ensure(s.n + 1)) return -1;
            CK(cudaMemcpyAsync(s.h_bound.p, s.bound.p, s.n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            std::vector<int32_t> &bv = s.bound_sorted;
            bv.clear();
            for (size_t i = 0; i < s.n; ++i)
                if (s.h_bound.p[i] < d_full - 1) bv.push_back(s.h_bound.p[i]);
            have_bounds = true;
            if (order_dev == s.order.p && !d->no_bound_order) {
                /* longest first by the pair's own bound (work ~ bound^2): the queue's tail is made of the cheapest pairs
                 * (counting sort, descending).  Pairs the bound pass could not bound within this budget go to the end of
                 * the list and straight to the re-dispatch tier: their first pass would run to the end of the budget for
                 * nothing (a pair flagged for the byte-compare kernel stays: the wavefront kernel routes it).
                 * The order list was uploaded before the sync. */
                std::vector<uint32_t> &cnt = s.bound_count;
                cnt.assign((size_t)d_full + 2, 0u);
                auto key_of = [&](int32_t b) -> size_t {
                    if (b == 0x7fffffff) return (size_t)d_full;              /* flagged: first */
                    if (b >= d_full - 1) return (size_t)d_full + 1;          /* open: last (descending order below) */
                    return (size_t)std::max(b, 0);
                };
                size_t n_open = 0;
                for (size_t i = 0; i < s.n; ++i) {
                    const size_t k = key_of(s.h_bound.p[i]);
                    cnt[k] += 1;
                    n_open += (k == (size_t)d_full + 1);
                }
                uint32_t acc = 0;
                for (int b = d_full; b >= 0; --b) { const uint32_t c0 = cnt[(size_t)b]; cnt[(size_t)b] = acc; acc += c0; }
                cnt[(size_t)d_full + 1] = acc;
                for (size_t i = 0; i < s.n; ++i) s.h_order.p[cnt[key_of(s.h_bound.p[i])]++] = (uint32_t)i;
                CK(cudaMemcpyAsync(s.order.p, s.h_order.p, s.n * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
                if (n_open > 0 && !d->no_skip_open) {
                    n_seed = n_open;
                    n_items = s.n - n_open;
                }
            }
            if (!bv.empty()) {
                const size_t k99 = (bv.size() - 1) - (bv.size() - 1) / 100;
                std::nth_element(bv.begin(), bv.begin() + (long)k99, bv.end());
                const int b99 = bv[k99];
                const int bmax = *std::max_element(bv.begin() + (long)k99, bv.end());
                d_want = std::min(d_full, bmax + 2);
                d_want = std::min(d_want, d_reach);
                d_p99 = std::min(d_want, b99 + 2);
                hinted = false;
            }
        }
    }
    const int pen_e = plan.e;
    const long long kt_max = s.kt_max;
    auto n_need = [&](int d_end) -> int {
        /* n_d grows, kt + (Dmax - d) / e shrinks: the largest min sits where they cross */
        const long long Dmax = d_end - 1;
        const long long ktm = std::min<long long>(kt_max, Dmax / pen_e);
        int lo = 0, hi = d_end;                      /* first d with n_d >= ktm + (Dmax - d) / e */
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if ((long long)tab[mid].n >= ktm + (Dmax - mid) / pen_e) hi = mid; else lo = mid + 1;
        }
        long long best = 1;
        for (int dd = std::max(0, lo - 2); dd < std::min(d_end, lo + 3); ++dd)
            best = std::max(best, std::min<long long>(tab[dd].n, ktm + (Dmax - dd) / pen_e));
        return (int)best;
    };
    int n_want = plan.band > 0 ? n_full : n_need(d_want);
    /* with a hint, rings for hint + 3 % are enough for (almost) every pair */
    int n_min = n_want;
    if (hinted && plan.band <= 0)
        n_min = std::min(n_want, n_need((int)std::min<long long>((long long)d->hint_dist + (long long)d->hint_dist * d->hint_min_pm / 1000 + 4, d_full - 1) + 1));
    if (d_p99 > 0) n_min = std::min(n_want, n_need(d_p99));      /* rings for 99 % of the pairs when that buys a resident CTA */
    const bool banded = plan.band > 0;
    LaunchCfg c{};
    int rc = 0;
    int win = 0;
    if (banded) {
        /* adaptive band: window = the reference's threads_per_block (lib/sequence_alignment.cu:270-277) */
        win = plan.band_width > 0 ? plan.band_width : 512;
        c.A = std::max(plan.o + plan.e, plan.x) + 1; c.E1 = plan.e + 1; c.G = c.A;
        c.seq_words = (int)packed_words_for(s.max_len);
        c.groups_per_cta = 1;
        /* a few diagonals per thread: the per-score overhead (window bookkeeping, barrier) is paid
         * per thread, and small CTAs let more pairs share an SM (measured on B200, W = 512:
         * 512 threads 114 k, 256 threads 165 k, 128 threads 189 k pairs/s at 10 kbp / 5 %) */
        c.group_threads = std::min(1024, std::max(64, ((win / 4) + 31) & ~31));
        /* packed pairs: four diagonals per thread (wfa_bandq_kernel), one quad per thread and score */
        c.quad = !ascii && !d->no_quad;
        /* a window spans win / 4 + 1 quads: two trips of half of them (measured on B200, W = 512, 10 kbp / 5 %:
         * 64 threads 32.5 ms, 96 threads 31.4 ms, 128 threads 33.0 ms per 8192 pairs end to end) */
        if (c.quad) c.group_threads = std::min(512, std::max(64, (((win / 4 + 2) / 2) + 31) & ~31));
        if (d->force_threads) c.group_threads = d->force_threads;
        auto band_smem = [&](int stages) {
            return c.quad ? bandq_smem_bytes(c.A, win, c.seq_words, stages) : banded_smem_bytes(c.A, win, c.seq_words, stages);
        };
        c.stages = 2;
        c.smem = band_smem(c.stages);
        if (c.smem > d->prop.sharedMemPerBlockOptin) {
            c.stages = 1;
            c.smem = band_smem(c.stages);
        }
        if (c.smem > d->prop.sharedMemPerBlockOptin) {
            fprintf(stderr, "[wfagpu] band of %d diagonals does not fit in shared memory\n", win);
            return -2;
        }
        c.n_cap = n_full; c.row_stride = win; c.center = 0;
        int occ = c.quad ? bandq_max_ctas_per_sm(c.group_threads, c.smem, plan.with_cigar != 0)
                         : banded_max_ctas_per_sm(c.group_threads, c.smem, ascii, plan.with_cigar != 0);
        if (occ < 1) return -1;
        c.ctas = (int)std::max<size_t>(1, std::min<size_t>((size_t)occ * d->prop.multiProcessorCount, n_items));
    } else {
        rc = choose_cfg(d, plan.x, plan.o, plan.e, n_want, n_min, s.max_len, n_items, ascii, plan.with_cigar != 0, &c);
    }
    if (rc) return rc;
    tr.mark("cfg");
    /* scores this launch can reach and the decision units they need */
    int d_end = banded ? d_full : d_want;
    uint64_t arena_units = s.tab_arena_units;
    if (!banded) {
        if (c.n_cap < n_want) {
            int lo = 1, hi = d_want;                 /* largest d_end whose pruned wavefronts fit the rings */
            while (lo < hi) {
                const int mid = (lo + hi + 1) / 2;
                if (n_need(mid) <= c.n_cap) lo = mid; else hi = mid - 1;
            }
            d_end = lo;
        }
        arena_units = d_end < d_full ? tab[d_end].row_off : s.tab_arena_units;
    }
    *capped_out = d_end < d_reach;          /* provisioned below what the budget (and the lengths) allow */

    /* memory budget of the arenas: a third of the free device memory */
    size_t arena_budget = 0;
    {
        /* cudaMemGetInfo takes 0.2 - 15 ms on a busy box and would leave the GPU idle between the pack
         * kernel and this pass: ask only when the answer can have changed (first use, arena grew) */
        if (d->arena_mb > 0) {
            s.budget_cache = (size_t)d->arena_mb << 20;          /* WFAGPU_ARENA_MB: fixed budget (tests) */
        } else if (s.budget_cache == 0) {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            s.budget_cache = std::max<size_t>((free_b + s.arena.cap * sizeof(uint4)) / 3, (size_t)64 << 20);
        }
        arena_budget = s.budget_cache;
    }
    tr.mark("meminfo");
    int period = 0;
    if (c.ckpt) {
        /* Snapshot period: the traceback recomputes ~ score * P cells per pair against ~ score^2 in
         * the forward pass, the snapshots take ~ 1/P of the cells: short periods for low scores,
         * longer ones when memory is tight. */
        const int d_expect = hinted ? std::min(d_end, d->hint_dist + 1) : d_end;
        period = d->force_period ? d->force_period : (d_expect >= 3000 ? 31 : 15);   /* measured on B200: shorter never wins */
        for (;;) {
            if (ensure_ck_table(s, plan, period)) return -1;
            arena_units = s.h_ck_off[(size_t)(d_end - 1) / period + 1];      /* snapshots of scores j * P < d_end */
            if (period >= 31 || d->force_period || (arena_units * sizeof(uint4) + 1) * n_items <= arena_budget) break;
            period = 2 * period + 1;
        }
        if (arena_units >= 0xffffffffull) { fprintf(stderr, "[wfagpu] snapshot arena exceeds 32-bit offsets\n"); return -1; }
    }

    if (!plan.with_cigar) arena_units = 0;
    /* adaptive band on packed pairs: decision rows per pair, walked by a traceback kernel after the launch */
    const bool band_tb = banded && c.quad && plan.with_cigar && !d->no_band_tb;
    /* keep the arenas within a third of the free device memory: fewer, not smaller, groups
     * (decision bytes: one arena per resident group; snapshots: one per pair of a sub-launch) */
    size_t items_per_launch = n_items;
    for (int attempt = 0;; ++attempt) {
        const size_t budget = arena_budget;
        if (c.ckpt || band_tb) {
            const size_t per_pair = (size_t)arena_units * sizeof(uint4) + (band_tb ? ((size_t)d_end + 1) * sizeof(int32_t) : 0) + 1;
            items_per_launch = std::max<size_t>(1, std::min<size_t>(n_items, budget / per_pair));
            c.ctas = (int)std::min<size_t>((size_t)c.ctas, items_per_launch);
        } else {
            const size_t per_group = (size_t)arena_units * sizeof(uint4) * (size_t)c.groups_per_cta + 1;
            const size_t max_ctas = std::max<size_t>(1, budget / per_group);
            if ((size_t)c.ctas > max_ctas) c.ctas = (int)max_ctas;
        }
        /* an arena that has to grow is checked against the memory that is free right now (the budget may date from an
         * emptier device: other slots, contexts and processes allocate too); the old arena is given back first */
        const size_t want = (((c.ckpt || band_tb) ? items_per_launch : (size_t)c.ctas * c.groups_per_cta) * arena_units + 1);
        if (want <= s.arena.cap || d->arena_mb > 0 || attempt >= 4) break;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); break; }
        const size_t avail = free_b + s.arena.cap * sizeof(uint4);
        const size_t spare = (size_t)2 << 30;                    /* text slots, op pool, staging of this pass */
        if (want * sizeof(uint4) + spare <= avail) break;
        arena_budget = std::max<size_t>((size_t)64 << 20, std::min(arena_budget / 2, avail > spare ? (avail - spare) / 2 : (size_t)0));
        s.budget_cache = arena_budget;
    }
    const size_t groups = (size_t)c.ctas * c.groups_per_cta;
    const uint64_t gring_elems = c.global_ring ? (uint64_t)(c.A + 2 * c.E1) * (uint64_t)c.row_stride : 0;
    if (s.gring.ensure(groups * gring_elems + 1)) return -1;
    const uint32_t scratch_words = plan.with_cigar ? (uint32_t)((2 * (size_t)d_end + 31) / 16 + 2) : 1;
    const size_t arenas = (c.ckpt || band_tb) ? items_per_launch : groups;
    if (arenas * arena_units + 1 > s.arena.cap) s.budget_cache = 0;      /* the arena grows: re-read the free memory next time */
    if (s.arena.ensure(arenas * arena_units + 1) || s.scratch.ensure(groups * scratch_words + 1)) return -1;
    const uint32_t band_lo_words = (banded && plan.with_cigar) ? (uint32_t)d_end + 1 : 0;
    if (s.band_lo.ensure((band_tb ? items_per_launch : groups) * (size_t)band_lo_words + 1)) return -1;
    /* op pool: worst case for this pass on top of what is already used */
    /* later passes run after read_counters(): the host copy of the pool head is current */
    const uint32_t pool_used = first_pass ? 0u : s.h_counters.p[CTR_POOL];
    const size_t pool_need = (size_t)pool_used + (plan.with_cigar ? n_items * (size_t)scratch_words : 0) + 16;
    if (pool_need >= (1ull << 32)) {
        fprintf(stderr, "[wfagpu] op pool exceeds 32-bit offsets; use a smaller batch_size\n");
        return -1;
    }
    if (s.pool.ensure(pool_need, true, s.stream)) return -1;

    tr.mark("buffers");
    if (n_seed > 0) {
        /* the re-dispatch list starts with the pairs that skip this pass */
        if (s.h_seed.ensure(1)) return -1;
        s.h_seed.p[0] = (uint32_t)n_seed;
        CK(cudaMemcpyAsync(s.counters.p + CTR_RETRY, s.h_seed.p, sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(retry_dev, s.order.p + n_items, n_seed * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s.stream));
    } else {
        CK(cudaMemsetAsync(s.counters.p + CTR_RETRY, 0, sizeof(uint32_t), s.stream));
    }

    KernelParams kp{};
    kp.packed = s.packed.p;
    kp.ascii = s.ascii.p;
    kp.pairs = s.pairs.p;
    kp.order = order_dev;
    kp.n_items = (uint32_t)n_items;
    kp.queue = s.counters.p + CTR_QUEUE;
    kp.tb_queue = s.counters.p + CTR_TBQ;
    kp.bound_queue = s.counters.p + CTR_BQ;
    /* per-pair bounds: first pass of the packed exact path only (a re-dispatched pair never
     * depends on them) */
    /* have_bounds: s.bound already holds this pass's bounds (run_bound_only) */
    bool use_bound = !banded && !ascii && (first_pass || have_bounds) && !d->no_bound;
    if (use_bound && !d->force_bound && !have_bounds) {
        /* Worth it?  With D = d_end - 1 and a typical score m, the launch bound leaves
         * cells(D/m) * m^2 cells per pair (cells(r) = r^2/4 + (1.5r - 1)(1 - r/2), 1 from r = 2 on), the
         * pair's own bound 0.5 * m^2; the bound pass costs ~32 cells per score at ~3x the cost per
         * cell (measured on B200: ~240 against ~2.75 warp instructions). */
        const double m = (hinted && d->hint_mean > 0) ? d->hint_mean : 0.5 * (d_end - 1);
        const double r = std::min(2.0, std::max(1.0, (d_end - 1) / std::max(1.0, m)));
        const double cells = r * r / 4 + (1.5 * r - 1) * (1 - r / 2);
        use_bound = (cells - 0.5) * m > 130.0;
    }
    if (use_bound && s.bound.ensure(s.n + 1)) return -1;
    kp.bound = use_bound ? s.bound.p : nullptr;
    kp.steps = s.steps.p;
    kp.d_end = d_end;
    kp.n_cap = c.n_cap;
    kp.x = plan.x; kp.o = plan.o; kp.e = plan.e;
    kp.A = c.A; kp.E1 = c.E1; kp.G = c.G;
    kp.row_stride = c.row_stride;
    kp.center = c.center;
    kp.seq_words = c.seq_words;
    kp.with_bt = plan.with_cigar;
    kp.stages = c.stages;
    kp.quad_pairs = c.quad_pairs ? 1 : 0;
    kp.ring_m = c.ring_m; kp.ring_g = c.ring_g;
    kp.band = plan.band;
    kp.win = win;
    kp.band_lo = s.band_lo.p;
    kp.band_lo_words = band_lo_words;
    kp.band_tb = band_tb ? 1 : 0;
    kp.gring = c.global_ring ? s.gring.p : nullptr;
    kp.gring_elems = gring_elems;
    kp.arena = s.arena.p;
    kp.arena_units = arena_units;
    kp.ck_off = c.ckpt ? s.ck_off.p : nullptr;
    kp.ck_period = period;
    kp.ops_scratch = s.scratch.p;
    kp.ops_scratch_words = scratch_words;
    kp.ops_pool = s.pool.p;
    kp.ops_pool_head = s.counters.p + CTR_POOL;
    kp.ops_pool_words = (uint32_t)std::min<size_t>(s.pool.cap, 0xffffffffu);
    kp.out = s.out.p;
    kp.retry_list = retry_dev;
    kp.retry_count = s.counters.p + CTR_RETRY;
    kp.ascii_list = s.ascii_list.p;
    kp.ascii_count = s.counters.p + CTR_ASCII;
    kp.cells = d->count_cells ? s.cells.p : nullptr;
    if (env_int("WFAGPU_VERBOSE", 0))
        fprintf(stderr, "[wfagpu] pass: items=%zu max_steps=%d n_cap=%d d_end=%d group=%d x%d ctas=%d stages=%d smem=%zu ascii=%d band=%d tier=%s period=%d arena=%.1f MB\n",
                n_items, max_steps, c.n_cap, d_end, c.group_threads, c.groups_per_cta, c.ctas, c.stages, c.smem,
                (int)ascii, plan.band > 0 ? win : 0, c.global_ring ? (c.quad ? "global-int32x4" : "global-int32") : (c.quad ? (c.quad_pairs ? (c.ckpt ? "smem-int16x4x2+ckpt" : "smem-int16x4x2") : (c.ckpt ? "smem-int16x4+ckpt" : "smem-int16x4")) : (c.ckpt ? "smem-int16+ckpt" : "smem-int16")), period, arenas * arena_units * 16.0 / 1e6);
    int tb_ctas = 0;
    constexpr int kTbWarps = 8;
    if (c.ckpt) {
        const int occ = traceback_max_ctas_per_sm(c.A, period, kTbWarps, ascii);
        if (occ < 1) { fprintf(stderr, "[wfagpu] traceback kernel does not fit\n"); return -1; }
        tb_ctas = occ * d->prop.multiProcessorCount;
    }
    int bound_ctas = 0;
    constexpr int kBoundWarps = 8;
    if (use_bound && !have_bounds) {
        const int occ = bound_max_ctas_per_sm(c.A, c.E1, kBoundWarps);
        if (occ < 1) kp.bound = nullptr; else bound_ctas = occ * d->prop.multiProcessorCount;
    }
    tr.mark("occ");
    for (size_t off = 0; off < n_items; off += items_per_launch) {
        const size_t cnt = std::min(items_per_launch, n_items - off);
        kp.order = order_dev + off;
        kp.n_items = (uint32_t)cnt;
        if (kp.bound && !have_bounds) {
            CK(cudaMemsetAsync(s.counters.p + CTR_BQ, 0, sizeof(uint32_t), s.stream));
            const int ctas = (int)std::min<size_t>((size_t)bound_ctas, (cnt + kBoundWarps - 1) / kBoundWarps);
            cudaError_t eb = launch_bound(kp, ctas, kBoundWarps, s.stream);
            if (eb != cudaSuccess) {
                fprintf(stderr, "[wfagpu] bound kernel launch failed: %s\n", cudaGetErrorString(eb));
                return -1;
            }
            s.stats.launches += 1;
        }
        CK(cudaMemsetAsync(s.counters.p + CTR_QUEUE, 0, sizeof(uint32_t), s.stream));
        const bool time_it = first_pass && items_per_launch >= n_items && s.ev[6] && s.ev[7];
        if (time_it) CK(cudaEventRecord(s.ev[6], s.stream));
        cudaError_t e = (banded && c.quad) ? launch_bandq(kp, c.group_threads, c.ctas, c.smem, s.stream)
                      : banded ? launch_banded(kp, c.group_threads, c.ctas, c.smem, ascii, s.stream)
                      : (c.quad && c.global_ring) ? launch_quadg(kp, c.group_threads, (int)std::min<size_t>(c.ctas, cnt), c.smem, s.stream)
                      : c.quad ? launch_quad(kp, c.group_threads, (int)std::min<size_t>(c.ctas, cnt), c.smem, s.stream)
                               : launch_exact(kp, c.group_threads, c.groups_per_cta, (int)std::min<size_t>(c.ctas, cnt), c.smem, ascii, s.stream);
        if (e != cudaSuccess) {
            fprintf(stderr, "[wfagpu] alignment kernel launch failed: %s\n", cudaGetErrorString(e));
            return -1;
        }
        if (time_it) { CK(cudaEventRecord(s.ev[7], s.stream)); s.wf_timed = true; }
        tr.mark("fwd");
        s.stats.launches += 1;
        if (band_tb) {
            e = launch_band_traceback(kp, s.stream);
            if (e != cudaSuccess) {
                fprintf(stderr, "[wfagpu] band traceback kernel launch failed: %s\n", cudaGetErrorString(e));
                return -1;
            }
            s.stats.launches += 1;
        }
        if (c.ckpt) {
            /* ring snapshots -> 2-bit ops, a warp per pair */
            CK(cudaMemsetAsync(s.counters.p + CTR_TBQ, 0, sizeof(uint32_t), s.stream));
            const int ctas = (int)std::min<size_t>((size_t)tb_ctas, (cnt + kTbWarps - 1) / kTbWarps);
            e = launch_traceback(kp, ctas, kTbWarps, ascii, s.stream);
            if (e != cudaSuccess) {
                fprintf(stderr, "[wfagpu] traceback kernel launch failed: %s\n", cudaGetErrorString(e));
                return -1;
            }
            s.stats.launches += 1;
        }
    }
    tr.mark("tb");
    if (first_pass) {
        s.stats.n_cap = (uint32_t)c.n_cap;
        s.stats.cta_threads = (uint32_t)(c.group_threads * c.groups_per_cta);
        s.stats.ctas = (uint32_t)c.ctas;
        s.stats.d_end = (uint32_t)d_end;
    }
    s.last_d_end = std::max(s.last_d_end, d_end);
    s.text_queued = false;
    return 0;
}

/* CIGAR text on the device: queued right behind the alignment kernel so that it overlaps with
 * the host's work on the previous chunk; redone by download() if pairs had to be re-dispatched. */
static int enqueue_text(wfagpu_device *d, Slot &s, size_t n, int d_end_bound)
{
    (void)d;
    /* slack slots: 10 characters per op, at most 2 ops per score */
    const size_t slot_bytes = n * ((size_t)20 * (size_t)d_end_bound + 64) + 64;
    if (s.slots.ensure(slot_bytes) || s.text.ensure(slot_bytes) || s.refs.ensure(n) || s.heads.ensure(4) ||
        s.h_refs.ensure(n) || s.h_heads.ensure(4))
        return -1;
    CK(cudaMemsetAsync(s.heads.p, 0, 4 * sizeof(unsigned long long), s.stream));
    CigarParams cp{};
    cp.ascii = s.ascii.p;
    cp.pairs = s.pairs.p;
    cp.out = s.out.p;
    cp.ops_pool = s.pool.p;
    cp.n_pairs = (uint32_t)n;
    cp.slots = s.slots.p;
    cp.slot_bytes = slot_bytes;
    cp.slot_head = s.heads.p;
    cp.text = s.text.p;
    cp.text_head = s.heads.p + 1;
    cp.refs = s.refs.p;
    cp.overflow = reinterpret_cast<uint32_t *>(s.heads.p + 2);
    launch_cigar_text(cp, s.stream);
    CK(cudaGetLastError());
    s.stats.launches += 2;
    CK(cudaMemcpyAsync(s.h_heads.p, s.heads.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(s.h_refs.p, s.refs.p, n * sizeof(wfagpu_cigar_ref_t), cudaMemcpyDeviceToHost, s.stream));
    s.text_queued = true;
    return 0;
}

extern "C" int wfagpu_device_align(wfagpu_device_t *d, int slot, size_t n, const wfagpu_plan_t *plan, int resident)
{
    (void)resident;
    if (!d || slot < 0 || slot > 1 || !plan) return -1;
    if (plan->x < 1 || plan->e < 1 || plan->o < 0) {
        /* x = 0 or e = 0 make a wavefront its own source (and divide by zero in the pruning quotient) */
        fprintf(stderr, "[wfagpu] penalties must satisfy x >= 1, e >= 1, o >= 0 (got %d,%d,%d)\n", plan->x, plan->o, plan->e);
        return -3;
    }
    CK(cudaSetDevice(d->dev));
    Slot &s = d->slots[slot];
    if (n != s.n) return -1;
    s.plan = *plan;
    if (n == 0) return 0;
    /* the batch may be re-aligned while it stays resident: start from clean counters */
    s.stats.launches = 0;
    s.stats.redispatched = 0;
    s.stats.ascii_pairs = 0;
    CK(cudaMemsetAsync(s.counters.p, 0, CTR_WORDS * sizeof(uint32_t), s.stream));
    CK(cudaMemsetAsync(s.cells.p, 0, sizeof(unsigned long long), s.stream));
    CK(cudaMemsetAsync(s.out.p, 0, n * sizeof(wfagpu_pair_out_t), s.stream));   /* a pair may skip the first pass: no stale record */
    CK(cudaEventRecord(s.ev[2], s.stream));
    PackParams pp{s.ascii.p, s.packed.p, s.pairs.p, (uint32_t)n};
    launch_pack(pp, s.stream);
    CK(cudaGetLastError());
    s.stats.launches += 1;
    CK(cudaEventRecord(s.ev[3], s.stream));
    s.last_d_end = 0;
    s.wf_timed = false;
    /* -e is where the reference hands a pair to its CPU aligner (lib/align.cu:237); here such pairs cost a second pass on the
     * GPU.  When the last batch of this kind needed scores a little beyond -e (at most a quarter more), the first pass is
     * given that budget straight away: 100 k x 1 kbp / 10 % with -e 300 re-dispatched 4.4 % of its pairs in every chunk. */
    int first_steps = plan->max_steps;
    pull_hint(d);
    if (plan->band <= 0 && d->hint_dist > 0 && d->hint_lift_pm > 0 && d->hint_key[0] == plan->x && d->hint_key[1] == plan->o &&
        d->hint_key[2] == plan->e && d->hint_key[3] == length_class(s.max_len)) {
        const long long want = (long long)d->hint_dist + (long long)d->hint_dist * d->hint_margin_pm / 1000 + 12;
        const long long most = (long long)plan->max_steps + (long long)plan->max_steps * d->hint_lift_pm / 1000;
        if (want > first_steps && d->hint_dist <= most) first_steps = (int)std::min<long long>(want, d->max_steps_cap);
    }
    s.first_steps = first_steps;
    int rc = launch_pass(d, s, *plan, first_steps, s.order.p, n, s.retry[0].p, false, true, true, &s.capped);
    if (rc) return rc;
    d->device_text = env_int("WFAGPU_HOST_CIGAR", 0) == 0;   /* re-read: tests flip it between runs */
    if (plan->with_cigar && d->device_text && enqueue_text(d, s, n, s.last_d_end)) return -1;
    CK(cudaEventRecord(s.ev[4], s.stream));                  /* align time = bound + wavefronts + traceback + CIGAR text */
    return 0;
}

extern "C" int wfagpu_device_download(wfagpu_device_t *d, int slot, size_t n, wfagpu_pair_out_t *out, uint32_t **ops,
                                      size_t *ops_used, uint32_t *pair_flags)
{
    if (!d || slot < 0 || slot > 1) return -1;
    CK(cudaSetDevice(d->dev));
    Slot &s = d->slots[slot];
    if (n != s.n) return -1;
    if (ops) *ops = nullptr;
    if (ops_used) *ops_used = 0;
    if (n == 0) return 0;
    const wfagpu_plan_t plan = s.plan;

    auto read_counters = [&]() -> int {
        CK(cudaMemcpyAsync(s.h_counters.p, s.counters.p, CTR_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        return 0;
    };
    if (read_counters()) return -1;

    /* Re-dispatched pairs always run the exact kernels: a band that lost the alignment does not find it
     * again with a larger budget (the reference hands such pairs to the CPU WFA, utils/wfa_cpu.c:30-86). */
    wfagpu_plan_t replan = plan;
    replan.band = 0;

    /* Budget oracle: score upper bounds of the pending pairs (bound kernel with the largest budget)
     * -> the number of wavefront steps that is enough for all of them, 0 if one has no bound. */
    auto bound_budget = [&](int cur, uint32_t pending, long long *need) -> int {
        *need = 0;
        const int kMaxSteps = d->max_steps_cap;
        int rc = run_bound_only(d, s, replan, kMaxSteps, s.retry[cur].p, pending);
        if (rc == -2) return 0;                     /* the bound kernel does not fit these penalties: keep doubling */
        if (rc) return rc;
        if (s.h_bound.ensure(s.n + 1) || s.h_retry.ensure(pending + 1)) return -1;
        CK(cudaMemcpyAsync(s.h_bound.p, s.bound.p, s.n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(s.h_retry.p, s.retry[cur].p, pending * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        const int d_big = s.tab_d_end;
        int maxb = 0;
        for (uint32_t i = 0; i < pending; ++i) {
            const uint32_t idx = s.h_retry.p[i];
            if (idx >= s.n) return -1;
            const int b = s.h_bound.p[idx];
            if (b >= d_big - 1) return 0;           /* not found within the largest budget: keep doubling */
            maxb = std::max(maxb, b);
        }
        long long steps_needed = 1;                 /* the reference counts the score-0 wavefront as step 1 */
        for (int dd = 1; dd <= maxb; ++dd) steps_needed += (s.h_steps.p[dd].kind == WFAGPU_STEP_MDI);
        *need = std::max<long long>(steps_needed + 3, 8);
        return 0;
    };

    /* ---- re-dispatch tier: pairs that outgrew the provisioned rings or the budget get, in one go,
     * the budget their score bounds call for (and are pruned by them); without bounds (banded,
     * byte-compare pairs) first the full budget, then the budget doubles until everything finishes ---- */
    /* pairs the GPU cannot finish (more than 60000 wavefront steps, or wavefronts wider than one CTA can hold):
     * their indices, taken from a device retry list */
    std::vector<uint32_t> failed;
    auto fail_pending = [&](int buf, uint32_t count) -> int {
        if (count == 0) return 0;
        if (s.h_retry.ensure((size_t)count + 1)) return -1;
        CK(cudaMemcpyAsync(s.h_retry.p, s.retry[buf].p, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        for (uint32_t i = 0; i < count; ++i)
            if (s.h_retry.p[i] < s.n) failed.push_back(s.h_retry.p[i]);
        return 0;
    };
    auto redispatch = [&](bool ascii, int start_steps, bool capped) -> int {
        int cur = 0;
        long long steps = start_steps;
        uint32_t pending = s.h_counters.p[CTR_RETRY];
        bool oracle_tried = false;
        while (pending > 0) {
            s.stats.redispatched += pending;
            long long next = capped ? steps : std::max<long long>(steps * 2, 64);
            bool have_bounds = false;
            if (!oracle_tried && !ascii && !d->no_bound) {
                oracle_tried = true;
                long long need = 0;
                int rcb = bound_budget(cur, pending, &need);
                if (rcb < 0) return rcb;
                if (rcb == 0 && need > 0) { next = need; have_bounds = true; }
            }
            if (next > d->max_steps_cap) {
                /* per pair, not per job: these pairs are reported as failed, everything else keeps its result */
                fprintf(stderr, "[wfagpu] %u pairs need more than %lld wavefront steps; reported as failed\n", pending, steps);
                return fail_pending(cur, pending);
            }
            steps = next;
            bool now_capped = false;
            int rc = launch_pass(d, s, replan, (int)steps, s.retry[cur].p, pending, s.retry[cur ^ 1].p, ascii, false, false,
                                 &now_capped, have_bounds);
            if (rc) return rc;
            if (read_counters()) return -1;
            if (now_capped && s.h_counters.p[CTR_RETRY] > 0) {
                fprintf(stderr, "[wfagpu] %u pairs exceed the wavefront capacity of one CTA; reported as failed\n",
                        s.h_counters.p[CTR_RETRY]);
                return fail_pending(cur ^ 1, s.h_counters.p[CTR_RETRY]);
            }
            capped = false;
            pending = s.h_counters.p[CTR_RETRY];
            cur ^= 1;
        }
        return 0;
    };
    int rc = redispatch(false, std::max(plan.max_steps, s.first_steps), s.capped);
    if (rc) return rc;

    /* ---- pairs with non-ACGT bytes: byte-compare kernel on the ASCII copy ---- */
    const uint32_t n_ascii = s.h_counters.p[CTR_ASCII];
    if (n_ascii > 0) {
        s.stats.ascii_pairs = n_ascii;
        bool capped = false;
        rc = launch_pass(d, s, plan, plan.max_steps, s.ascii_list.p, n_ascii, s.retry[0].p, true, false, false, &capped);
        if (rc) return rc;
        if (read_counters()) return -1;
        rc = redispatch(true, plan.max_steps, capped);
        if (rc) return rc;
    }
    CK(cudaEventRecord(s.ev[5], s.stream));

    /* ---- results back to the host ---- */
    if (s.h_out.ensure(n)) return -1;
    const uint32_t pool_used = s.h_counters.p[CTR_POOL];
    if (s.h_pool.ensure((size_t)pool_used + 1)) return -1;
    CK(cudaMemcpyAsync(s.h_out.p, s.out.p, n * sizeof(wfagpu_pair_out_t), cudaMemcpyDeviceToHost, s.stream));
    if (pool_used)
        CK(cudaMemcpyAsync(s.h_pool.p, s.pool.p, (size_t)pool_used * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
    if (pair_flags) CK(cudaMemcpyAsync(s.h_pairs.p, s.pairs.p, n * sizeof(wfagpu_pair_t), cudaMemcpyDeviceToHost, s.stream));
    if (d->count_cells) CK(cudaMemcpyAsync(s.h_cells.p, s.cells.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (uint32_t idx : failed) {
        s.h_out.p[idx].status = WFAGPU_ST_FAILED;
        s.h_out.p[idx].distance = -1;
        s.h_out.p[idx].n_ops = 0;
    }
    s.stats.failed_pairs = (uint32_t)failed.size();
    memcpy(out, s.h_out.p, n * sizeof(wfagpu_pair_out_t));
    if (!d->independent) learn_hint(d, s, n);
    if (pair_flags)
        for (size_t i = 0; i < n; ++i) pair_flags[i] = s.h_pairs.p[i].flags;
    if (ops) *ops = s.h_pool.p;
    if (ops_used) *ops_used = pool_used;
    s.stats.d2h_bytes += n * sizeof(wfagpu_pair_out_t) + (size_t)pool_used * 4;
    if (d->count_cells) s.stats.cells = s.h_cells.p[0];
    if (s.have_events) {
        cudaEventElapsedTime(&s.stats.ms_h2d, s.ev[0], s.ev[1]);
        cudaEventElapsedTime(&s.stats.ms_pack, s.ev[2], s.ev[3]);
        cudaEventElapsedTime(&s.stats.ms_align, s.ev[3], s.ev[5]);
        cudaEventElapsedTime(&s.stats.ms_total, s.ev[0], s.ev[5]);
    }
    return 0;
}

extern "C" int wfagpu_device_rescore(wfagpu_device_t *d, int slot, size_t n, const wfagpu_plan_t *plan, int32_t *scores)
{
    if (!d || !plan || !scores || slot < 0 || slot > 1) return -1;
    if (n != d->slots[slot].n) return -1;
    /* a different code path on purpose: what it shares with the production path is the packing, the step table
     * and the extend */
    const bool nb = d->no_bound, nq = d->no_quad, uh = d->use_hint;
    const int fw = d->force_warp;
    d->no_bound = true; d->no_quad = true; d->use_hint = false; d->force_warp = 0; d->independent = true;
    wfagpu_plan_t p2 = *plan;
    p2.with_cigar = 0;
    std::vector<wfagpu_pair_out_t> out(n);
    int rc = wfagpu_device_align(d, slot, n, &p2, 1);
    if (!rc) rc = wfagpu_device_download(d, slot, n, out.data(), nullptr, nullptr, nullptr);
    d->no_bound = nb; d->no_quad = nq; d->use_hint = uh; d->force_warp = fw; d->independent = false;
    if (rc) return rc;
    for (size_t i = 0; i < n; ++i) scores[i] = (out[i].status & WFAGPU_ST_FINISHED) ? out[i].distance : -1;
    return 0;
}

extern "C" int wfagpu_device_wait(wfagpu_device_t *d, int slot, float *ms_pack, float *ms_align)
{
    if (!d || slot < 0 || slot > 1) return -1;
    CK(cudaSetDevice(d->dev));
    Slot &s = d->slots[slot];
    CK(cudaStreamSynchronize(s.stream));
    if (ms_pack) cudaEventElapsedTime(ms_pack, s.ev[2], s.ev[3]);
    if (ms_align) cudaEventElapsedTime(ms_align, s.ev[3], s.ev[4]);
    s.stats.ms_wavefront = 0;
    if (s.wf_timed) cudaEventElapsedTime(&s.stats.ms_wavefront, s.ev[6], s.ev[7]);
    /* pairs the timed pass left for the re-dispatch tier / the byte-compare kernel (both run in download()) */
    if (s.n) {
        CK(cudaMemcpyAsync(s.h_counters.p, s.counters.p, CTR_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        s.stats.pending_pairs = s.h_counters.p[CTR_RETRY] + s.h_counters.p[CTR_ASCII];
    }
    if (s.n && d->use_hint) {
        /* read the result records back (16 B per pair) so that a re-run of the resident
         * batch is provisioned like the next batch of a stream would be */
        if (s.h_out.ensure(s.n)) return -1;
        CK(cudaMemcpyAsync(s.h_out.p, s.out.p, s.n * sizeof(wfagpu_pair_out_t), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        learn_hint(d, s, s.n);
    }
    return 0;
}

/* 1 if `ptr` is page-locked (cudaHostAlloc / cudaHostRegister), 0 if pageable or unknown */
extern "C" int wfagpu_host_is_pinned(const void *ptr)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    return at.type == cudaMemoryTypeHost ? 1 : 0;
}

/* Page-locked staging area of a slot (grow-only): a caller with a pageable sequence buffer copies the batch here
 * (with its own threads) and uploads from it, so that the H2D copy is asynchronous DMA like for pinned callers.
 * The previous upload from this slot must have been waited for (download). */
extern "C" char *wfagpu_device_staging(wfagpu_device_t *d, int slot, size_t bytes)
{
    if (!d || slot < 0 || slot > 1) return nullptr;
    if (cudaSetDevice(d->dev) != cudaSuccess) return nullptr;
    Slot &s = d->slots[slot];
    if (s.h_ascii.ensure(bytes + 64)) return nullptr;
    return s.h_ascii.p;
}

/* Page-locked host memory for the aligner's own buffers (the reference's TODO, utils/sequence_reader.c:73):
 * NULL when no CUDA device / driver is usable -- the caller then falls back to plain calloc. */
extern "C" void *wfagpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    memset(p, 0, bytes);
    return p;
}
extern "C" void wfagpu_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int wfagpu_host_register(void *ptr, size_t bytes)
{
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fprintf(stderr, "[wfagpu] cudaHostRegister failed: %s\n", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

extern "C" int wfagpu_host_unregister(void *ptr)
{
    return cudaHostUnregister(ptr) == cudaSuccess ? 0 : -1;
}

extern "C" int wfagpu_device_download_text(wfagpu_device_t *d, int slot, size_t n, const char **text, size_t *text_bytes,
                                           const wfagpu_cigar_ref_t **refs)
{
    if (!d || slot < 0 || slot > 1 || !text || !text_bytes || !refs) return -1;
    CK(cudaSetDevice(d->dev));
    Slot &s = d->slots[slot];
    if (n != s.n) return -1;
    *text = nullptr; *text_bytes = 0; *refs = nullptr;
    if (n == 0) return 0;
    if (!s.text_queued && enqueue_text(d, s, n, s.last_d_end)) return -1;
    CK(cudaStreamSynchronize(s.stream));
    if (s.h_heads.p[2] != 0) {
        fprintf(stderr, "[wfagpu] CIGAR text pool overflow\n");
        return -1;
    }
    const size_t used = (size_t)s.h_heads.p[1];
    if (s.h_text.ensure(used + 1)) return -1;
    if (used) CK(cudaMemcpyAsync(s.h_text.p, s.text.p, used, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    s.stats.d2h_bytes += used + n * sizeof(wfagpu_cigar_ref_t);
    *text = s.h_text.p;
    *text_bytes = used;
    *refs = s.h_refs.p;
    return 0;
}

extern "C" void wfagpu_device_last_stats(wfagpu_device_t *d, int slot, wfagpu_batch_stats_t *st)
{
    if (!d || !st || slot < 0 || slot > 1) return;
    *st = d->slots[slot].stats;
}

extern "C" int wfagpu_device_pack_only(wfagpu_device_t *d, const char *ascii, size_t ascii_bytes, wfagpu_pair_t *pairs,
                                       size_t n, uint32_t *packed_out, size_t packed_words)
{
    if (!d) return -1;
    if (wfagpu_device_upload(d, 0, ascii, ascii_bytes, pairs, n)) return -1;
    Slot &s = d->slots[0];
    PackParams pp{s.ascii.p, s.packed.p, s.pairs.p, (uint32_t)n};
    launch_pack(pp, s.stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s.stream));
    if (packed_words < s.packed_words) return -1;
    CK(cudaMemcpy(packed_out, s.packed.p, s.packed_words * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pairs, s.pairs.p, n * sizeof(wfagpu_pair_t), cudaMemcpyDeviceToHost));
    return 0;
}

/* ------------------------------------------------------------------------ */
/* utils/device_query.cu:27-54 equivalents                                   */
extern "C" void get_num_cuda_devices(int *n)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) c = 0;
    if (n) *n = c;
}

extern "C" char *get_cuda_dev_name(int dev)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return nullptr;
    return strdup(prop.name);
}

extern "C" int get_cuda_SM_count(int dev)
{
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return v;
}

extern "C" void get_cuda_capability(int dev, int *major, int *minor)
{
    int a = 0, b = 0;
    cudaDeviceGetAttribute(&a, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMinor, dev);
    if (major) *major = a;
    if (minor) *minor = b;
}
