// cf_api.cu -- the C ABI of include/cf_b200.h: context, resident plans, one-shot runs and the
// RNG parity kernels.  No CPU fallback: every entry point needs a CUDA device.
#include "../../include/cf_b200.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>

#include "cf_dlm.cuh"
#include "cf_multi.cuh"
#include "cf_dupire.cuh"
#include "cf_bs.cuh"
#include "cf_kernels.cuh"
#include "cf_tables.h"
#include "cf_pick.h"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
constexpr size_t kFastSmemLimit = 227 * 1024;   // opt-in shared memory per block on sm_100

struct CfError : std::runtime_error { using std::runtime_error::runtime_error; };

#define CF_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw CfError(std::string(#call) + ": " + cudaGetErrorString(e__));                    \
    } while (0)

template <class F>
int guarded(F&& f)
{
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
    catch (...) { g_err = "unknown error"; return 1; }
}

// the default memory pool keeps what it has been given instead of returning it at every synchronisation
void keep_pool_memory(int dev)
{
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
}

// ---- the devices of the context (cf_init).  Device 0 of the list is driven by the calling thread; the others, if
// any, by one worker thread each (launches on all devices are issued at the same time, not one device after the other).
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    std::atomic<int> state{0};            // 0 idle, 1 job posted, 2 job done, 3 quit
    std::exception_ptr err;

    template <class Pred> void spinThenWait(Pred done)
    {
        for (int i = 0; i < 20000 && !done(); ++i) { /* ~100 us of polling: a run is a fraction of a millisecond */ }
        if (done()) return;
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, done);
    }
    void post(std::function<void()> f)
    {
        { std::lock_guard<std::mutex> lk(m); job = std::move(f); err = nullptr; state.store(1, std::memory_order_release); }
        cv.notify_all();
    }
    void wait()
    {
        spinThenWait([&] { return state.load(std::memory_order_acquire) == 2; });
        state.store(0, std::memory_order_release);
        if (err) std::rethrow_exception(err);
    }
    void stop()
    {
        if (!th.joinable()) return;
        { std::lock_guard<std::mutex> lk(m); state.store(3, std::memory_order_release); }
        cv.notify_all();
        th.join();
    }
    ~Worker() { stop(); }
};

struct DeviceCtx {
    int id = -1, sms = 0, index = 0;      // CUDA ordinal, SM count, position in the context
    cudaStream_t stream = nullptr;        // owned: host-buffer runs of a multi-device context
    unsigned char* staging = nullptr;     // pinned block the table uploads of a plan are gathered in
    cudaEvent_t copied = nullptr;         // the last batch has left the staging block
    std::unique_ptr<Worker> worker;       // devices 1 .. n - 1
};
std::vector<std::unique_ptr<DeviceCtx>> g_devs;
thread_local DeviceCtx* t_dev = nullptr;  // the device this thread is bound to
int g_gen = 0;                            // bumped whenever the context is closed
thread_local int t_gen = 0;
#define g_sms (t_dev->sms)

void bind(DeviceCtx* d)
{
    if (t_dev != d) { CF_CUDA(cudaSetDevice(d->id)); t_dev = d; }
}
// bind the calling thread to another device of the context for the lifetime of the scope
struct DeviceScope {
    DeviceCtx* prev;
    explicit DeviceScope(DeviceCtx* d) : prev(t_dev) { bind(d); }
    ~DeviceScope() { if (prev && prev != t_dev) { cudaSetDevice(prev->id); t_dev = prev; } }
};

void worker_main(DeviceCtx* d)
{
    Worker& w = *d->worker;
    try { bind(d); } catch (...) { /* reported by the first job */ }
    for (;;) {
        w.spinThenWait([&] { const int s = w.state.load(std::memory_order_acquire); return s == 1 || s == 3; });
        if (w.state.load(std::memory_order_acquire) == 3) return;
        try { bind(d); w.job(); } catch (...) { w.err = std::current_exception(); }
        { std::lock_guard<std::mutex> lk(w.m); w.state.store(2, std::memory_order_release); }
        w.cv.notify_all();
    }
}

std::unique_ptr<DeviceCtx> open_device(int id, int index, bool withWorker)
{
    auto d = std::make_unique<DeviceCtx>();
    d->id = id; d->index = index;
    CF_CUDA(cudaSetDevice(id));
    CF_CUDA(cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, id));
    keep_pool_memory(id);
    if (withWorker) {
        CF_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        if (index > 0) {
            // the worker threads are stopped before the process tears the CUDA runtime down
            static const bool registered = [] { std::atexit([] { for (auto& dd : g_devs) if (dd->worker) dd->worker->stop(); }); return true; }();
            (void)registered;
            d->worker = std::make_unique<Worker>();
            d->worker->th = std::thread(worker_main, d.get());
        }
    }
    return d;
}

void retire_plans();     // the plans of the context go with it (defined after cf_plan)

// Pinned result blocks of host-buffer runs, kept across plans: cudaHostAlloc / cudaFreeHost cost about a millisecond
// each, which a one-shot cf_run_* (a plan per call) paid on every call.
struct PinnedCache {
    std::mutex m;
    std::vector<std::pair<double*, size_t>> blocks;
    double* take(size_t n, size_t& cap)
    {
        {
            std::lock_guard<std::mutex> lock(m);
            size_t best = blocks.size();
            for (size_t i = 0; i < blocks.size(); ++i)
                if (blocks[i].second >= n && (best == blocks.size() || blocks[i].second < blocks[best].second)) best = i;
            if (best < blocks.size()) {
                double* p = blocks[best].first; cap = blocks[best].second;
                blocks.erase(blocks.begin() + long(best));
                return p;
            }
        }
        double* p = nullptr;
        cap = std::max<size_t>(n, 2048);
        CF_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p), cap * sizeof(double), cudaHostAllocPortable));
        return p;
    }
    void give(double* p, size_t cap)
    {
        if (!p) return;
        std::lock_guard<std::mutex> lock(m);
        if (blocks.size() < 8) blocks.emplace_back(p, cap); else cudaFreeHost(p);
    }
    void clear()
    {
        std::lock_guard<std::mutex> lock(m);
        for (auto& b : blocks) cudaFreeHost(b.first);
        blocks.clear();
    }
};
PinnedCache g_pinned;

void close_devices()
{
    retire_plans();
    g_pinned.clear();
    for (auto& d : g_devs) {
        if (d->worker) d->worker->stop();
        cudaSetDevice(d->id);
        cudaDeviceSynchronize();
        if (d->stream) cudaStreamDestroy(d->stream);
        if (d->staging) cudaFreeHost(d->staging);
        if (d->copied) cudaEventDestroy(d->copied);
    }
    g_devs.clear();
    t_dev = nullptr;
    ++g_gen;
}

// Binds the calling thread to device 0 of the context; without cf_init the context is the current CUDA device.
void ensure_init()
{
    if (g_devs.empty()) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw CfError("cf_b200: no CUDA device available (this library has no CPU fallback)");
        int dev = 0;
        CF_CUDA(cudaGetDevice(&dev));
        g_devs.push_back(open_device(dev, 0, false));
        t_dev = nullptr;
        t_gen = g_gen;
    }
    if (t_gen != g_gen) { t_dev = nullptr; t_gen = g_gen; }          // the context has been re-created since this thread bound
    if (!t_dev) bind(g_devs[0].get());
    else CF_CUDA(cudaSetDevice(t_dev->id));                          // the caller may have changed the current device
}

// Table uploads of one plan are batched: while an UploadBatch is open, upload() copies the host data into a pinned
// staging block and hands out a slice of one device arena; closing the batch sends everything in a single
// cudaMemcpyAsync (a plan has a few dozen tables of a few KB: one copy each costs more than the kernels save).
struct UploadBatch {
    static constexpr size_t kCapacity = size_t(4) << 20;
    unsigned char* dev = nullptr;       // device arena of this batch (owned by the plan)
    size_t used = 0;
    static UploadBatch*& current() { static thread_local UploadBatch* b = nullptr; return b; }
    static unsigned char*& staging() { return t_dev->staging; }      // of the device this thread is bound to
    static cudaEvent_t& copied() { return t_dev->copied; }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owned = true;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    // stream-ordered allocations from the device's default pool (kept by the pool after the first call)
    ~DevBuf() { release(); }
    // allocation and release are ordered on stream s (the stream the buffer is used on; nullptr: the legacy default stream)
    void release(cudaStream_t s = nullptr) { if (p && owned) cudaFreeAsync(p, s); p = nullptr; n = 0; owned = true; }
    void alloc(size_t count, cudaStream_t s = nullptr)
    {
        release(s);
        n = count;
        if (count) CF_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), s));
    }
    void upload(const T* src, size_t count, cudaStream_t s = nullptr)
    {
        UploadBatch* b = UploadBatch::current();
        const size_t bytes = count * sizeof(T), off = b ? (b->used + 255) / 256 * 256 : 0;
        if (b && count && off + bytes <= UploadBatch::kCapacity) {
            release();
            std::memcpy(UploadBatch::staging() + off, src, bytes);
            p = reinterpret_cast<T*>(b->dev + off); n = count; owned = false;
            b->used = off + bytes;
            return;
        }
        alloc(count);
        if (count) CF_CUDA(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, s));
    }
};

// Opens a batch on construction, sends it on close().  arena: the plan's buffer that owns the device block.
struct UploadScope {
    UploadBatch batch;
    explicit UploadScope(DevBuf<unsigned char>& arena)
    {
        if (!UploadBatch::staging()) {
            CF_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&UploadBatch::staging()), UploadBatch::kCapacity, cudaHostAllocPortable));
            CF_CUDA(cudaEventCreateWithFlags(&UploadBatch::copied(), cudaEventDisableTiming));
        } else {
            CF_CUDA(cudaEventSynchronize(UploadBatch::copied()));     // the previous batch has left the staging block
        }
        arena.alloc(UploadBatch::kCapacity);
        batch.dev = arena.p;
        UploadBatch::current() = &batch;
    }
    void close()
    {
        if (UploadBatch::current() != &batch) return;
        UploadBatch::current() = nullptr;
        if (batch.used) CF_CUDA(cudaMemcpyAsync(batch.dev, UploadBatch::staging(), batch.used, cudaMemcpyHostToDevice, nullptr));
        CF_CUDA(cudaEventRecord(UploadBatch::copied(), nullptr));
    }
    ~UploadScope() { if (UploadBatch::current() == &batch) UploadBatch::current() = nullptr; }
};

// Scratch of one plan: path history, per-block partials, per-warp vol-adjoint tables.  Grown on demand on the launch
// stream (stream-ordered allocation: a buffer is only released after the work queued before it on that stream), never
// shrunk.  A plan is used on one stream at a time; two plans never share scratch.
struct Scratch {
    DevBuf<double> hist, state, partial, partialRev, wtab, btab, tmp, local, local2;
    DevBuf<uint32_t> live;
    template <class T> void need(DevBuf<T>& b, size_t n, cudaStream_t s) { if (b.n < n) b.alloc(n, s); }
};

using cf::KernelFn;

KernelFn pick(int mdl, int prd, bool aad, int rng)
{
    KernelFn fn = cf::pick_path_kernel(mdl, prd, aad, rng);
    if (!fn) throw CfError("cf_b200: model/product combination not implemented on the device");
    return fn;
}

size_t smem_for(int mdl, bool aad, int D, int m, int E, int dim, bool sobol, int lutN, int nPayRows, bool big = false)
{
    if (mdl == CF_MODEL_DUPIRE)
        return aad ? cf::smem_sizes<CF_MODEL_DUPIRE, true>(D, m, E, dim, sobol, lutN, nPayRows, big).total
                   : cf::smem_sizes<CF_MODEL_DUPIRE, false>(D, m, E, dim, sobol, lutN, nPayRows, big).total;
    return aad ? cf::smem_sizes<CF_MODEL_BS, true>(D, m, E, dim, sobol, lutN, nPayRows, big).total
               : cf::smem_sizes<CF_MODEL_BS, false>(D, m, E, dim, sobol, lutN, nPayRows, big).total;
}

size_t adj_size(const cf_model* mdl)
{
    if (mdl->kind == CF_MODEL_DUPIRE) {
        if (mdl->n_times > 0 && mdl->time_col1 && mdl->time_col2 && mdl->time_w1 && mdl->time_w2)
            return 1 + size_t(mdl->n_knots) * mdl->n_times;
        return 1 + size_t(mdl->n_steps) * mdl->n_knots;
    }
    if (mdl->kind == CF_MODEL_BS) return 1 + 2 * size_t(mdl->n_steps) + 4 * size_t(mdl->n_events);
    if (mdl->kind == CF_MODEL_DISPLACED) return size_t(cf::dlm_adj_size(mdl->n_assets, mdl->n_steps, mdl->n_events));
    throw CfError("cf_b200: model kind not implemented");
}

void validate(const cf_model* mdl, const cf_product* prd, const cf_rng* rng)
{
    if (!mdl || !prd || !rng) throw CfError("cf_b200: null descriptor");
    if (mdl->n_steps < 1 || mdl->n_events < 1) throw CfError("cf_b200: empty timeline");
    if (mdl->n_events != prd->n_events) throw CfError("cf_b200: model and product disagree on the number of event dates");
    if (!mdl->is_event) throw CfError("cf_b200: is_event missing");
    int ev = 0;
    for (int i = 0; i <= mdl->n_steps; ++i) ev += mdl->is_event[i] ? 1 : 0;
    if (ev != mdl->n_events) throw CfError("cf_b200: is_event does not mark n_events points");
    const bool multi = mdl->kind == CF_MODEL_DISPLACED;
    if (!multi && mdl->n_assets != 1) throw CfError("cf_b200: Black-Scholes and Dupire are single-asset models");
    if (mdl->kind == CF_MODEL_DUPIRE) {
        if (mdl->n_knots < 1 || !mdl->log_spots || !mdl->interp_vols) throw CfError("cf_b200: Dupire tables missing");
        for (int j = 0; j + 1 < mdl->n_knots; ++j)
            if (!(mdl->log_spots[j] < mdl->log_spots[j + 1])) throw CfError("cf_b200: Dupire spot knots must increase");
    } else if (mdl->kind == CF_MODEL_BS) {
        if (!mdl->bs_drifts || !mdl->bs_stds) throw CfError("cf_b200: Black-Scholes tables missing");
        for (int i = 1; i <= mdl->n_steps; ++i)
            if (!mdl->is_event[i]) throw CfError("cf_b200: every Black-Scholes step must end on an event date");
    } else if (multi) {
        if (mdl->n_assets < 1 || mdl->n_assets > 16) throw CfError("cf_b200: the displaced model supports 1 to 16 assets");
        if (!mdl->dlm_spots || !mdl->dlm_chol || !mdl->dlm_alphas || !mdl->dlm_dynamics || !mdl->dlm_dyn_fwd || !mdl->dlm_drifts
            || !mdl->dlm_stds || !mdl->dlm_fwd_factors) throw CfError("cf_b200: displaced model tables missing");
        for (int i = 1; i <= mdl->n_steps; ++i)
            if (!mdl->is_event[i]) throw CfError("cf_b200: every step of the displaced model must end on an event date");
        for (int k = 0; k < mdl->n_assets; ++k)
            if (mdl->dlm_dynamics[k] < 0 || mdl->dlm_dynamics[k] > 3) throw CfError("cf_b200: unknown dynamics");
    } else throw CfError("cf_b200: model kind not implemented");
    if (prd->kind == CF_PRODUCT_EUROPEAN) { if (prd->n_payoffs != 1 || prd->n_events != 1) throw CfError("cf_b200: European has one event and one payoff"); }
    else if (prd->kind == CF_PRODUCT_UOC) { if (prd->n_payoffs != 2) throw CfError("cf_b200: UOC has two payoffs"); if (!(prd->smooth > 0)) throw CfError("cf_b200: UOC smooth must be > 0"); }
    else if (prd->kind == CF_PRODUCT_EUROPEANS) {
        if (!prd->strikes || !prd->strike_offsets || prd->n_payoffs < 1) throw CfError("cf_b200: Europeans needs strikes and their offsets per event");
        if (prd->strike_offsets[0] != 0 || prd->strike_offsets[prd->n_events] != prd->n_payoffs) throw CfError("cf_b200: Europeans strike offsets do not cover the payoffs");
        for (int e = 0; e < prd->n_events; ++e)
            if (prd->strike_offsets[e + 1] < prd->strike_offsets[e]) throw CfError("cf_b200: Europeans strike offsets must not decrease");
    }
    else if (prd->kind == CF_PRODUCT_AUTOCALL) {
        if (prd->n_payoffs != 1 || !prd->weights || !prd->event_dt) throw CfError("cf_b200: Autocall needs references, the period length and has one payoff");
        if (!(prd->smooth > 0) || !(prd->strike > 0)) throw CfError("cf_b200: Autocall smooth and strike must be > 0");
        if (mdl->is_event[0]) throw CfError("cf_b200: an Autocall has no sample today");
    } else if (prd->kind == CF_PRODUCT_BASKETS) {
        if (prd->n_events != 1 || prd->n_payoffs < 1 || !prd->strikes || !prd->weights) throw CfError("cf_b200: Baskets needs weights, strikes and a single event");
    } else if (prd->kind == CF_PRODUCT_CONTINGENT) {
        if (mdl->kind != CF_MODEL_BS) throw CfError("cf_b200: the contingent bond runs under Black-Scholes only on the device");
        if (prd->n_payoffs != 1 || prd->n_events < 2 || !prd->event_dt) throw CfError("cf_b200: ContingentBond needs at least one period, the coverages and has one payoff");
        if (!(prd->smooth >= 0)) throw CfError("cf_b200: ContingentBond smooth must be >= 0");
    } else if (prd->kind == CF_PRODUCT_MULTISTATS) {
        const int A = mdl->n_assets, E = prd->n_events;
        if (prd->n_payoffs != (2 * E - 1) * (A + A * (A + 1) / 2)) throw CfError("cf_b200: MultiStats payoff count does not match assets and dates");
    } else throw CfError("cf_b200: product kind not implemented");
    const bool multiPrd = prd->kind == CF_PRODUCT_AUTOCALL || prd->kind == CF_PRODUCT_BASKETS || prd->kind == CF_PRODUCT_MULTISTATS;
    if (multi != multiPrd) throw CfError("cf_b200: model/product combination not implemented on the device");
    if (prd->n_payoffs > 2048) throw CfError("cf_b200: too many payoffs for one product");
    if (rng->kind == CF_RNG_SOBOL) {
        if (mdl->n_steps * mdl->n_assets > cf::sobol_max_dim()) throw CfError("cf_b200: Sobol dimension exceeds 1101");
    } else if (rng->kind != CF_RNG_MRG32K3A) throw CfError("cf_b200: unknown RNG kind");
}

// ---- the participants of the rank sum (cf_comm.cuh): the devices of this context, or one device per process.
// A receive block is 2 x world x capacity slots of 16 bytes (256 MB per device for 8 participants at the default
// capacity of 2^20 doubles, CF_COMM_CAPACITY)
struct CommState {
    int world = 0;                        // participants, 0: no communicator
    int rank0 = 0;                        // rank of local device 0 (local device k is participant rank0 + k)
    size_t cap = 0;                       // doubles per row
    bool enabled = false;                 // launches exchange (cf_comm_enable)
    bool ipc = false;                     // participants are processes (CUDA IPC) rather than devices of this process
    uint32_t epoch = 0;                   // exchanges so far, the same on every participant
    long long timeout = 20000000000ll;    // clocks (~10 s) a participant waits for its peers
    std::vector<unsigned char*> local;    // receive block of each local device (cudaMalloc)
    std::vector<void*> opened;            // remote blocks opened through IPC
    unsigned char* block[cf::kMaxPeers] = {};   // receive block of participant r as addressable from this process
    int* status = nullptr;                // mapped pinned host word: epoch of an exchange that timed out, 0: none
    int* statusDev = nullptr;

    size_t blockBytes() const { return size_t(2) * size_t(world) * cap * sizeof(cf::PeerSlot); }
    bool on() const { return world > 1 && enabled; }
    cf::DPeers peers(int localIndex, uint32_t ep) const
    {
        cf::DPeers p{};
        p.world = world; p.rank = rank0 + localIndex; p.epoch = ep; p.cap = cap; p.timeout = timeout;
        (void)localIndex;
        for (int r = 0; r < world; ++r) p.buf[r] = reinterpret_cast<cf::PeerSlot*>(block[r]);
        p.status = statusDev;
        return p;
    }
    void check() const
    {
        if (status && *reinterpret_cast<volatile int*>(status) != 0)
            throw CfError("cf_b200: a participant of the rank sum never arrived (exchange " + std::to_string(*status)
                          + " timed out): results of that run are NaN");
    }
};
CommState g_comm;

void comm_destroy()
{
    for (auto& d : g_devs) { cudaSetDevice(d->id); cudaDeviceSynchronize(); }
    for (void* q : g_comm.opened) cudaIpcCloseMemHandle(q);
    for (size_t k = 0; k < g_comm.local.size(); ++k) {
        if (k < g_devs.size()) cudaSetDevice(g_devs[k]->id);
        if (g_comm.local[k]) cudaFree(g_comm.local[k]);
    }
    if (g_comm.status) cudaFreeHost(g_comm.status);
    g_comm = CommState{};
    if (t_dev) cudaSetDevice(t_dev->id);
}

// receive blocks of the local devices (zeroed), tickets, the status word
void comm_allocate(int world, int rank0, size_t cap)
{
    if (world < 2 || world > cf::kMaxPeers) throw CfError("cf_comm: the number of participants must be 2 .. 16");
    if (cap == 0) throw CfError("cf_comm: zero capacity");
    comm_destroy();
    g_comm.world = world; g_comm.rank0 = rank0; g_comm.cap = cap;
    CF_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_comm.status), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    *g_comm.status = 0;
    CF_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_comm.statusDev), g_comm.status, 0));
    for (auto& d : g_devs) {
        DeviceScope sc(d.get());
        unsigned char* b = nullptr;
        CF_CUDA(cudaMalloc(reinterpret_cast<void**>(&b), g_comm.blockBytes()));
        CF_CUDA(cudaMemset(b, 0, g_comm.blockBytes()));
        CF_CUDA(cudaDeviceSynchronize());
        g_comm.local.push_back(b);
    }
}

// Shard of participant k of P over n paths: boundaries on multiples of 256 paths (one Sobol window) and even (an
// antithetic pair of mrg32k3a is never split; the reference guarantees the same with its 64-path batches,
// mcBase.h:312); the last participant takes the remainder.  Mirrors compfinance_b200/dist.py::shard_range.
void shard_range(uint64_t n, int k, int P, uint64_t& first, uint64_t& count)
{
    uint64_t step = 256;
    uint64_t per = (n / uint64_t(P)) / step * step;
    if (per == 0) { step = 2; per = (n / uint64_t(P)) / step * step; }
    first = uint64_t(k) * per;
    count = k < P - 1 ? per : n - per * uint64_t(P - 1);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// The plan of one device: tables resident in HBM, scratch, kernels.
struct DevPlan {
    DeviceCtx* dev = nullptr;             // the device the tables live on
    DevBuf<double> outBuf, localBuf, perBuf, aggBuf;   // host-buffer runs: results (and local sums before the rank sum)
    int mdlKind = 0, prdKind = 0, rngKind = 0;
    int D = 0, m = 0, E = 0, dim = 0, nPay = 0;
    size_t nAdj = 0;           // table adjoints including the spot leaf
    DevBuf<unsigned char> arena;      // one device block for all the tables uploaded by make_plan
    bool tablesInFlight = true;       // the first launch orders its stream after the table upload (null stream)
    Scratch scratch;
    DevBuf<unsigned long long> dbgTimes;    // CF_DEBUG_TIMES: phase stamps of the Dupire kernels (cf_plan_debug_times)
    cf::KArgs base{};
    DevBuf<uint8_t> isEvent;
    DevBuf<double> tabA, tabB, num, ff, disc, libors, eventDt;
    DevBuf<uint32_t> sobolDir;
    DevBuf<uint64_t> mrgJump;
    DevBuf<uint8_t> lut;
    int lutN = 0, storeG = 1;
    // Black-Scholes fast path (cf_bs.cuh): European / UOC
    bool bsFast = false;
    DevBuf<double2> bsDs;
    // Dupire fast kernel (cf_dupire.cuh)
    bool fast = false, hasTimeMap = false;
    int nTimes = 0;
    DevBuf<int32_t> tk1, tk2;
    DevBuf<double> tc1, tc2;
    // packed tables of the fast kernel: padded vol rows, buckets, cell records, step bits, time map
    DevBuf<double2> ab, bk, cells, c12;
    DevBuf<uint32_t> stepBits;
    DevBuf<int32_t> k12;
    DevBuf<uint8_t> flushOps;
    DevBuf<double2> spanW;            // span reverse kernel: per step (padded) time weights and column offsets of its two targets
    DevBuf<uint2> spanOff;
    int spanS = 0;                    // steps per lane; 0: the span kernel cannot run this plan
    int nCells = 0;
    cf::DArgs dbase{};
    // displaced multi-asset model (cf_dlm.cuh)
    cf::LArgs lbase{};
    DevBuf<double> eStrikes;
    DevBuf<int32_t> eOff;
    DevBuf<double> lSpots, lChol, lAlphas, lDynFwd, lDrifts, lStds, lStepPack, lFf, lNum, lStrikes, lPw, lW;
    DevBuf<int32_t> lDyn;
    int A = 1;
    int partialStride = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;   // recorded since last query
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;

    ~DevPlan()
    {
        for (auto& e : events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto& e : pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    }

    size_t outSize(bool aad) const { return aad ? size_t(nPay) + 1 + nAdj : size_t(nPay); }

    std::pair<cudaEvent_t, cudaEvent_t> takeEvents()
    {
        std::pair<cudaEvent_t, cudaEvent_t> ev;
        if (events.size() >= 256) {       // nobody polls cf_plan_kernel_ms: recycle the oldest pair instead of growing
            pool.push_back(events.front());
            events.erase(events.begin());
        }
        if (!pool.empty()) { ev = pool.back(); pool.pop_back(); }
        else { CF_CUDA(cudaEventCreate(&ev.first)); CF_CUDA(cudaEventCreate(&ev.second)); }
        return ev;
    }

    // The sum over participants of a result vector already reduced on this device (generic paths; the Dupire fast path
    // does it inside its own reduction kernel).
    void exchange(const double* dLocal, int nOut, double* dOut, const cf::DPeers& px, cudaStream_t s)
    {
        if (size_t(nOut) > px.cap) throw CfError("cf_b200: result vector longer than the communicator's capacity");
        const int block = 256, grid = std::max(1, std::min(4 * dev->sms, (nOut + block - 1) / block));
        cf::peer_exchange_kernel<<<grid, block, 0, s>>>(dLocal, nOut, dOut, px);
        CF_CUDA(cudaGetLastError());
        ++g_launches;
    }

    // px: the participants of the rank sum, or null (single GPU, or the caller reduces)
    void launch(bool aad, const double* w, uint64_t first, uint64_t n, double* dOut, double* dPerPath,
                double* dPerAgg, cudaStream_t s, const cf::DPeers* px = nullptr)
    {
        if (px && n == 0) {               // a participant without paths still takes part in the exchange
            const int nOut = int(outSize(aad));
            scratch.need(scratch.tmp, size_t(nOut), s);
            CF_CUDA(cudaMemsetAsync(scratch.tmp.p, 0, sizeof(double) * size_t(nOut), s));
            exchange(scratch.tmp.p, nOut, dOut, *px, s);
            return;
        }
        if (n == 0) throw CfError("cf_b200: n_paths must be > 0");
        if (rngKind == CF_RNG_SOBOL && first + n > 0xffffffffull) throw CfError("cf_b200: Sobol index exceeds 2^32 - 1");
        // the reference's skipTo takes an unsigned path index (mrg32k3a.h:192); the jump matrices cover 32 bits of pair index
        if (rngKind != CF_RNG_SOBOL && first + n > (1ull << 32)) throw CfError("cf_b200: mrg32k3a path index exceeds 2^32 - 1");
        if (tablesInFlight) {
            if (UploadBatch::copied()) CF_CUDA(cudaStreamWaitEvent(s, UploadBatch::copied(), 0));
            tablesInFlight = false;
        }
        const uint64_t nb64 = (n + cf::kBlock - 1) / cf::kBlock;
        if (nb64 > 0x7fffffffull) throw CfError("cf_b200: too many paths in one launch");
        const int nBatches = int(nb64);
        if (mdlKind == CF_MODEL_DISPLACED) { launchDlm(aad, w, first, n, dOut, dPerPath, dPerAgg, s, px); return; }
        if (bsFast) { launchFastBS(aad, w, first, n, dOut, dPerPath, dPerAgg, s, px); return; }
        if (fast && (!aad || hasTimeMap)) { launchFast(aad, w, first, n, dOut, dPerPath, dPerAgg, s, px); return; }

        const int grid = std::min(nBatches, 2 * g_sms);
        const size_t tabAdj = mdlKind == CF_MODEL_DUPIRE ? 1 + size_t(D) * m : nAdj;   // generic kernel: table adjoints
        const size_t stride = aad ? size_t(nPay) + 1 + tabAdj : size_t(nPay);
        partialStride = int(stride);
        scratch.need(scratch.partial, size_t(grid) * stride, s);
        if (aad) scratch.need(scratch.hist, size_t(2) * D * size_t(grid) * cf::kBlock, s);
        cf::KArgs a = base;
        a.first_path = first; a.n_paths = n; a.n_batches = nBatches;
        a.w[0] = a.w[1] = 0.0;
        if (aad) for (int k = 0; k < nPay && k < cf::kMaxPay; ++k) a.w[k] = w[k];
        if (aad && prdKind == CF_PRODUCT_EUROPEANS) {
            CF_CUDA(cudaMemcpyAsync(lW.p, w, sizeof(double) * size_t(nPay), cudaMemcpyHostToDevice, s));
            CF_CUDA(cudaStreamSynchronize(s));
        }
        a.wlong = lW.p;
        a.partial = scratch.partial.p; a.partial_stride = partialStride;
        a.per_path_payoffs = dPerPath; a.per_path_agg = dPerAgg;
        a.hist = scratch.hist.p;
        KernelFn fn = pick(mdlKind, prdKind, aad, rngKind);
        const int payRows = prdKind == CF_PRODUCT_EUROPEANS ? nPay : 0;
        size_t smem = smem_for(mdlKind, aad, D, m, E, dim, rngKind == CF_RNG_SOBOL, lutN, payRows);
        if (smem > kFastSmemLimit) {
            // long schedules: table A and the table adjoints leave shared memory (KArgs::big_tables)
            a.big_tables = 1;
            smem = smem_for(mdlKind, aad, D, m, E, dim, rngKind == CF_RNG_SOBOL, lutN, payRows, true);
        }
        if (smem > kFastSmemLimit) throw CfError("cf_b200: tables do not fit in shared memory (n_steps or n_payoffs too large)");
        CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        auto ev = takeEvents();
        CF_CUDA(cudaEventRecord(ev.first, s));
        fn<<<grid, cf::kBlock, smem, s>>>(a);
        CF_CUDA(cudaEventRecord(ev.second, s));
        events.push_back(ev);
        CF_CUDA(cudaGetLastError());
        const int nOutAll = int(outSize(aad));
        double* dLocal = dOut;
        if (px) { scratch.need(scratch.local, size_t(nOutAll), s); dLocal = scratch.local.p; }
        if (aad && mdlKind == CF_MODEL_DUPIRE && hasTimeMap) {
            // generic kernel produced interp_vols adjoints: reduce, then apply the time map
            scratch.need(scratch.tmp, stride, s);
            cf::reduce_partials_kernel<<<(int(stride) + 127) / 128, 128, 0, s>>>(scratch.partial.p, grid, partialStride, int(stride), scratch.tmp.p);
            cf::collapse_time_kernel<<<(nOutAll + 127) / 128, 128, 0, s>>>(scratch.tmp.p, nPay + 2, D, m, nTimes, tk1.p, tk2.p, tc1.p, tc2.p, dLocal);
            g_launches += 1;
        } else {
            cf::reduce_partials_kernel<<<(nOutAll + 127) / 128, 128, 0, s>>>(scratch.partial.p, grid, partialStride, nOutAll, dLocal);
        }
        CF_CUDA(cudaGetLastError());
        g_launches += 2;
        if (px) exchange(dLocal, nOutAll, dOut, *px, s);
    }

    using LKernel = cf::LKernel;


    void launchDlm(bool aad, const double* w, uint64_t first, uint64_t n, double* dOut, double* dPerPath,
                   double* dPerAgg, cudaStream_t s, const cf::DPeers* px)
    {
        // one block per SM of 8 warps (255 registers); CF_DLM_WARPS=12 tries 12 warps with mrg32k3a (168 registers: measured slower)
        const bool sobol = rngKind == CF_RNG_SOBOL;
        const int amax = cf::dlm_bucket(A);
        int warps = 0, stepsInSmem = 0;
        size_t smem = 0;
        static const int forcedWarps = [] { const char* e = std::getenv("CF_DLM_WARPS"); return e ? std::atoi(e) : 0; }();
        for (int tryWarps : {12, 8}) {
            if ((sobol || forcedWarps != 12) && tryWarps != 8) continue;        // 8 warps measured faster than 12 (registers)
            for (int trySteps : {1, 0}) {
                if (warps || (trySteps && !cf::dlm_steps_fit(A, D))) continue;
                smem = cf::dlm_smem(A, amax, D, E, nPay, dim, sobol, aad, tryWarps, lbase.has_alpha != 0, trySteps != 0).total;
                if (smem <= kFastSmemLimit) { warps = tryWarps; stepsInSmem = trySteps; }
            }
        }
        if (!warps) throw CfError("cf_b200: displaced model tables do not fit in shared memory (n_steps * n_assets or payoffs too large)");
        const int threads = warps * 32;
        const int nBatchesL = int((n + uint64_t(threads) - 1) / uint64_t(threads));
        const int grid = std::min(nBatchesL, g_sms);
        const size_t stride = aad ? size_t(nPay) + 1 + nAdj : size_t(nPay);
        partialStride = int(stride);
        scratch.need(scratch.partial, size_t(grid) * stride, s);
        cf::LArgs a = lbase;
        a.first_path = first; a.n_paths = n; a.n_batches = nBatchesL; a.steps_in_smem = stepsInSmem;
        if (aad) {
            scratch.need(scratch.hist, size_t(D) * (2 * A + 1) * size_t(grid) * threads, s);
            const size_t tabDoubles = size_t(grid) * warps * size_t(cf::dlm_warp_tab(A, D, E));
            scratch.need(scratch.tmp, tabDoubles, s);
            CF_CUDA(cudaMemsetAsync(scratch.tmp.p, 0, sizeof(double) * tabDoubles, s));
            a.warp_tab = scratch.tmp.p;
            if (lbase.has_alpha) {
                const size_t cols = size_t(A) * size_t(grid) * threads;
                scratch.need(scratch.local2, cols, s);
                CF_CUDA(cudaMemsetAsync(scratch.local2.p, 0, sizeof(double) * cols, s));
                a.alpha_cols = scratch.local2.p;
            }
            // the weights are read by the kernel from device memory: stage them on the launch stream
            CF_CUDA(cudaMemcpyAsync(lW.p, w, sizeof(double) * size_t(nPay), cudaMemcpyHostToDevice, s));
            CF_CUDA(cudaStreamSynchronize(s));
        }
        a.w = lW.p;
        a.partial = scratch.partial.p; a.partial_stride = partialStride;
        a.per_path_payoffs = dPerPath; a.per_path_agg = dPerAgg; a.hist = scratch.hist.p;
        LKernel fn = cf::pick_dlm_kernel(A, prdKind, aad, rngKind, warps);
        if (!fn) throw CfError("cf_b200: MultiStats is a value-only test instrument on the device (no AAD)");
        CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        auto ev = takeEvents();
        CF_CUDA(cudaEventRecord(ev.first, s));
        fn<<<grid, threads, smem, s>>>(a);
        CF_CUDA(cudaEventRecord(ev.second, s));
        events.push_back(ev);
        CF_CUDA(cudaGetLastError());
        const int nOut = int(outSize(aad));
        double* dLocal = dOut;
        if (px) { scratch.need(scratch.local, size_t(nOut), s); dLocal = scratch.local.p; }
        cf::reduce_partials_kernel<<<(nOut + 127) / 128, 128, 0, s>>>(scratch.partial.p, grid, partialStride, nOut, dLocal);
        CF_CUDA(cudaGetLastError());
        g_launches += 2;
        if (px) exchange(dLocal, nOut, dOut, *px, s);
    }

    using DKernel = cf::DKernel;
    // Paths per thread of the forward kernel: 2 (the Sobol low-bit lookups are shared by a thread's paths); 1 when the run
    // has fewer warp-units than the GPU has warp slots, so that a small run (one shard of eight) still fills the SMs.
    int forwardP(uint64_t n) const
    {
        static const int forced = [] { const char* e = std::getenv("CF_DUPIRE_FWD_P"); return e ? std::atoi(e) : 0; }();
        if (forced == 1 || forced == 2) return forced;
        const uint64_t units2 = (n + 511) / 512 * 8;
        return units2 * 8 <= uint64_t(g_sms) * cf::kFwdWarps * 5 ? 1 : 2;
    }

    // Runs are cut into launches of at most kFastChunk paths: the log-spot history of one launch is
    // n_steps * 8 bytes per path (1.3 GB for 2^20 paths x 156 steps).
    static constexpr uint64_t kFastChunk = 1ull << 21;

    // ---- itemised risk of Dupire x Europeans (cf_multi.cuh): strikes sorted per event, classes, result maps
    bool multiReady = false;
    int multiCmax = 1;
    double spot0 = 0.0;
    DevBuf<double> mK;
    DevBuf<int32_t> mEvent, mRank, mOrig;
    DevBuf<long long> mThi, mTlo;
    DevBuf<double> mT, mPartial, mSums;

    size_t multiOutSize() const { return size_t(nPay) + nAdj * size_t(nPay); }

    // dOut: [n_payoffs] payoff sums, then [nAdj][n_payoffs] sums over paths of d payoff / d table
    void launchMulti(uint64_t first, uint64_t n, double* dOut, cudaStream_t s, const cf::DPeers* px)
    {
        if (!multiReady) throw CfError("cf_b200: this plan has no strike-class kernel");
        const size_t nOutAll = multiOutSize();
        double* dLocal = dOut;
        if (px) { scratch.need(scratch.local, nOutAll, s); dLocal = scratch.local.p; }
        if (px && n == 0) {
            CF_CUDA(cudaMemsetAsync(dLocal, 0, sizeof(double) * nOutAll, s));
            exchange(dLocal, int(nOutAll), dOut, *px, s);
            return;
        }
        if (n == 0) throw CfError("cf_b200: n_paths must be > 0");
        if (rngKind == CF_RNG_SOBOL && first + n > 0xffffffffull) throw CfError("cf_b200: Sobol index exceeds 2^32 - 1");
        if (rngKind != CF_RNG_SOBOL && first + n > (1ull << 32)) throw CfError("cf_b200: mrg32k3a path index exceeds 2^32 - 1");
        if (tablesInFlight) {
            if (UploadBatch::copied()) CF_CUDA(cudaStreamWaitEvent(s, UploadBatch::copied(), 0));
            tablesInFlight = false;
        }
        const size_t tabLen = 1 + size_t(D) * m;
        const size_t tSize = size_t(E) * multiCmax * tabLen;
        if (mThi.n < tSize) { mThi.alloc(tSize, s); mTlo.alloc(tSize, s); mT.alloc(tSize, s); }
        CF_CUDA(cudaMemsetAsync(mThi.p, 0, sizeof(long long) * tSize, s));
        CF_CUDA(cudaMemsetAsync(mTlo.p, 0, sizeof(long long) * tSize, s));
        const uint64_t nb64 = (n + cf::kBlock - 1) / cf::kBlock;
        if (nb64 > 0x7fffffffull) throw CfError("cf_b200: too many paths in one launch");
        const int grid = int(std::min<uint64_t>(nb64, uint64_t(2) * dev->sms));
        if (mPartial.n < size_t(grid) * nPay) mPartial.alloc(size_t(grid) * nPay, s);
        if (mSums.n < size_t(nPay)) mSums.alloc(size_t(nPay), s);
        // fixed point of the remainders: |lo| <= 2^-11 per addend, n addends: 2^(lo_bits - 11) n < 2^62
        int loBits = 50;
        while (loBits > 20 && std::ldexp(double(n), loBits - 11) >= std::ldexp(1.0, 62)) --loBits;
        // the coarse part: |x| 2^10 n < 2^62 with |x| up to ~ 2^10 spot
        if (std::ldexp(double(n) * std::fabs(spot0), 20) >= std::ldexp(1.0, 62)) throw CfError("cf_b200: spot x paths too large for the fixed-point accumulators of the itemised risk");
        cf::MArgs a{};
        // batches dealt round-robin over the blocks: measured 8.2 ms against 9.5 ms with contiguous ranges and incremental
        // jumps (config 4, 2^22 paths) -- the opposite of the generic kernel; CF_MULTI_STRIDED=0 switches
        static const int strided = [] { const char* e = std::getenv("CF_MULTI_STRIDED"); return e ? std::atoi(e) : 1; }();
        a.strided = strided;
        a.first_path = first; a.n_paths = n; a.n_batches = int(nb64);
        a.seed1 = base.seed1; a.seed2 = base.seed2; a.dim = dim;
        a.sobol_dir = sobolDir.p; a.mrg_jump = mrgJump.p;
        a.D = D; a.m = m; a.E = E; a.is_event = isEvent.p; a.spot = spot0;
        a.interp_vols = tabA.p; a.log_spots = tabB.p;
        a.ksorted = mK.p; a.koff = eOff.p; a.n_payoffs = nPay; a.cmax = multiCmax;
        a.partial = mPartial.p; a.Thi = mThi.p; a.Tlo = mTlo.p; a.lo_scale = std::ldexp(1.0, loBits); a.per_path_payoffs = nullptr;
        a.lut = lut.p; a.lut_n = lutN; a.lut_x0 = base.lut_x0; a.lut_scale = base.lut_scale;
        const bool sob = rngKind == CF_RNG_SOBOL;
        const size_t smem = cf::multi_smem(D, m, nPay, dim, sob, lutN).total;
        if (smem > kFastSmemLimit / 2) throw CfError("cf_run_aad_multi: tables do not fit in shared memory");
        auto fn = cf::pick_multi_kernel(rngKind);
        CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        auto ev = takeEvents();
        CF_CUDA(cudaEventRecord(ev.first, s));
        fn<<<grid, cf::kBlock, smem, s>>>(a);
        CF_CUDA(cudaEventRecord(ev.second, s));
        events.push_back(ev);
        CF_CUDA(cudaGetLastError());
        cf::reduce_partials_kernel<<<(nPay + 127) / 128, 128, 0, s>>>(mPartial.p, grid, nPay, nPay, mSums.p);
        cf::multi_unsort_kernel<<<(nPay + 127) / 128, 128, 0, s>>>(mSums.p, mOrig.p, nPay, dLocal);
        const size_t nSuffix = size_t(E) * tabLen;
        cf::multi_suffix_kernel<<<unsigned((nSuffix + 255) / 256), 256, 0, s>>>(mThi.p, mTlo.p, std::ldexp(1.0, -loBits), mT.p, E, multiCmax, tabLen);
        const size_t nParam = 1 + size_t(m) * nTimes;
        cf::multi_collapse_kernel<<<unsigned((nParam * nPay + 255) / 256), 256, 0, s>>>(mT.p, E, multiCmax, D, m, nTimes, tk1.p, tk2.p, tc1.p, tc2.p,
                                                                                       mEvent.p, mRank.p, mOrig.p, nPay, dLocal + nPay);
        CF_CUDA(cudaGetLastError());
        g_launches += 5;
        if (px) exchange(dLocal, int(nOutAll), dOut, *px, s);
    }

    // The reverse sweep has two forms (cf_dupire.cuh): classic (one path per lane, 8 warps: many live paths per SM) and
    // span (one warp per live path, a lane per S consecutive steps, 16 warps: few live paths per SM, the shard of a
    // multi-GPU run).  Measured cross-over: about 1100 paths per SM.  CF_DUPIRE_REV = span | classic forces one.
    enum { kRevSpan = 0, kRevClassic = 2 };
    int reverseForm(uint64_t nPad) const
    {
        static const int forced = [] {
            const char* e = std::getenv("CF_DUPIRE_REV");
            return !e ? -1 : (std::strcmp(e, "classic") == 0 ? int(kRevClassic) : (std::strcmp(e, "span") == 0 ? int(kRevSpan) : -1));
        }();
        int form = forced >= 0 ? forced : (nPad <= uint64_t(g_sms) * 1100 ? int(kRevSpan) : int(kRevClassic));
        // the span kernel's blocks scan the live mask of the whole launch
        if (form == kRevSpan && (spanS == 0 || nPad / 32 > uint64_t(cf::kRevSMaxWords))) form = kRevClassic;
        return form;
    }

    // kernel launch on stream s, optionally with programmatic stream serialization (the kernel may start while its
    // predecessor runs and waits for it in griddepcontrol.wait)
    template <class Args>
    static void launchKernel(void (*fn)(const Args), int grid, int block, size_t smem, cudaStream_t s, const Args& a, bool pdl)
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(unsigned(grid)); cfg.blockDim = dim3(unsigned(block)); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        CF_CUDA(cudaLaunchKernelEx(&cfg, fn, a));
    }

    // Black-Scholes x {European, UOC}: the shared forward kernel with the log-normal step, the span reverse of cf_bs.cuh
    // over the live paths, one reduction (with the rank sum).  Launches of at most 2^21 paths (history: n_steps x 8 bytes
    // per path).
    void launchFastBS(bool aad, const double* w, uint64_t first, uint64_t n, double* dOut,
                      double* dPerPath, double* dPerAgg, cudaStream_t s, const cf::DPeers* px)
    {
        static const bool pdl = [] { const char* e = std::getenv("CF_PDL"); return !e || std::atoi(e) != 0; }();
        constexpr uint64_t kChunkBS = kFastChunk;
        const bool sob = rngKind == CF_RNG_SOBOL;
        const uint64_t maxChunk = std::min<uint64_t>(n, kChunkBS);
        const int fwdP = forwardP(maxChunk), fwdWarps = cf::kFwdWarps;
        const uint64_t quantum = 256ull * fwdP;
        const int histRow = (D + 3) / 4 * 4;
        const uint64_t maxPad = (maxChunk + quantum - 1) / quantum * quantum;
        const int gridF = std::min(int(maxPad / quantum) * 8, g_sms);
        const int minGridR = int((maxPad / 32 + cf::kBsMaxWords - 1) / cf::kBsMaxWords);
        const int gridR = std::max(int(std::min<uint64_t>(maxPad / 32, uint64_t(g_sms))), minGridR);
        const int nAdjBs = cf::bs_adj_size(D, E);
        scratch.need(scratch.partial, size_t(gridF) * (size_t(nPay) + 1), s);
        if (aad) {
            scratch.need(scratch.hist, size_t(histRow) * maxPad, s);
            scratch.need(scratch.state, 2 * maxPad, s);
            scratch.need(scratch.live, size_t(maxPad / 32), s);
            scratch.need(scratch.partialRev, size_t(gridR) * nAdjBs, s);
        }
        DKernel fwd = cf::pick_bs_forward(prdKind, aad, rngKind, fwdP);
        const size_t smemF = fwdP == 2 ? cf::dupire_smem_fwd4<2, cf::kFwdChunk>(D, 0, dim, sob, 0, fwdWarps, true).total
                                       : cf::dupire_smem_fwd4<1, cf::kFwdChunk1>(D, 0, dim, sob, 0, fwdWarps, true).total;
        const int S = cf::dupire_span_steps(D);
        DKernel rev = aad ? cf::pick_bs_reverse(prdKind, S) : nullptr;
        const size_t smemR = cf::bs_smem_rev(D, E).total;
        CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(fwd), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemF)));
        if (aad) CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(rev), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemR)));
        auto ev = takeEvents();
        CF_CUDA(cudaEventRecord(ev.first, s));
        for (uint64_t off = 0; off < n; off += kChunkBS) {
            const uint64_t cnt = std::min<uint64_t>(kChunkBS, n - off);
            cf::DArgs a = dbase;
            a.first_path = first + off; a.n_paths = cnt;
            a.n_pad = (cnt + quantum - 1) / quantum * quantum;
            a.accumulate = off ? 1 : 0;
            a.w[0] = a.w[1] = 0.0;
            if (aad) for (int k = 0; k < nPay && k < cf::kMaxPay; ++k) a.w[k] = w[k];
            a.partial = scratch.partial.p; a.partial_rev = scratch.partialRev.p;
            a.hist = scratch.hist.p; a.state = scratch.state.p; a.live = scratch.live.p;
            a.per_path_payoffs = dPerPath ? dPerPath + off * nPay : nullptr;
            a.per_path_agg = dPerAgg ? dPerAgg + off : nullptr;
            a.n_units = int(a.n_pad / quantum) * 8;
            launchKernel(fwd, gridF, fwdWarps * 32, smemF, s, a, false);
            ++g_launches;
            if (aad) {
                launchKernel(rev, gridR, cf::kBsRevBlock, smemR, s, a, pdl);
                ++g_launches;
            }
        }
        CF_CUDA(cudaEventRecord(ev.second, s));
        events.push_back(ev);
        const int nHead = aad ? nPay + 1 : nPay, nTail = aad ? nAdjBs : 0, nOut = nHead + nTail;
        cf::DPeers pr{};
        if (px) {
            if (size_t(nOut) > px->cap) throw CfError("cf_b200: result vector longer than the communicator's capacity");
            pr = *px;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(unsigned((nOut * 32 + 255) / 256)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = (pdl && aad) ? 1 : 0;
        const double* cHead = scratch.partial.p; const double* cTail = scratch.partialRev.p;
        CF_CUDA(cudaLaunchKernelEx(&cfg, cf::rows_reduce_kernel, cHead, gridF, nPay + 1, nHead, cTail, gridR, nAdjBs, nTail, dOut, pr));
        ++g_launches;
    }

    void launchFast(bool aad, const double* w, uint64_t first, uint64_t n, double* dOut,
                    double* dPerPath, double* dPerAgg, cudaStream_t s, const cf::DPeers* px)
    {
        static const bool pdl = [] { const char* e = std::getenv("CF_PDL"); return !e || std::atoi(e) != 0; }();
        const bool sob = rngKind == CF_RNG_SOBOL;
        const int fwdP = forwardP(std::min<uint64_t>(n, kFastChunk)), fwdWarps = cf::kFwdWarps;
        const uint64_t quantum = 256ull * fwdP;
        const int histRow = (D + 3) / 4 * 4;
        const uint64_t maxChunk = std::min<uint64_t>(n, kFastChunk);
        const uint64_t maxPad = (maxChunk + quantum - 1) / quantum * quantum;
        const int maxUnitsF = int(maxPad / quantum) * 8;
        const int gridF = std::min(maxUnitsF, g_sms);       // units are dealt round-robin: small runs still use every SM
        const int form = reverseForm(maxPad);
        const bool span = form == kRevSpan;
        const int revWarps = span ? cf::kRevSWarps : cf::kRevWarps;
        const int revMaxWords = span ? cf::kRevSMaxWords : cf::kRevMaxWords;
        // classic: reverse blocks own contiguous ranges of live-mask words (32 paths each), at most revMaxWords per block;
        // span: equal shares of the live paths of the launch, one block per SM
        const int minGridR = span ? 1 : int((maxPad / 32 + revMaxWords - 1) / revMaxWords);
        const int gridR = std::max(int(std::min<uint64_t>(maxPad / 32, uint64_t(g_sms))), minGridR);
        const size_t tabLen = size_t(nTimes) * m;
        scratch.need(scratch.partial, size_t(gridF) * (size_t(nPay) + 1), s);
        if (aad) {
            scratch.need(scratch.hist, size_t(histRow) * maxPad, s);
            scratch.need(scratch.state, 2 * maxPad, s);
            scratch.need(scratch.live, size_t(maxPad / 32), s);
            scratch.need(scratch.partialRev, size_t(gridR), s);
            if (!span) scratch.need(scratch.wtab, size_t(gridR) * revWarps * tabLen, s);
            scratch.need(scratch.btab, size_t(gridR) * tabLen, s);
        }
        DKernel fwd;
        size_t smemF;
        static const int forcedCh = [] { const char* e = std::getenv("CF_DUPIRE_FWD_CH"); return e ? std::atoi(e) : 0; }();
        const int fwdCh = fwdP == 2 ? cf::kFwdChunk : (forcedCh == cf::kFwdChunk ? cf::kFwdChunk : cf::kFwdChunk1);
        fwd = cf::pick_dupire_forward(prdKind, aad, rngKind, fwdP, fwdCh);
        smemF = fwdP == 2 ? cf::dupire_smem_fwd4<2, cf::kFwdChunk>(D, m, dim, sob, nCells, fwdWarps).total
              : fwdCh == cf::kFwdChunk ? cf::dupire_smem_fwd4<1, cf::kFwdChunk>(D, m, dim, sob, nCells, fwdWarps).total
                                       : cf::dupire_smem_fwd4<1, cf::kFwdChunk1>(D, m, dim, sob, nCells, fwdWarps).total;
        auto rev = span ? cf::pick_dupire_reverse_span(prdKind, spanS) : cf::pick_dupire_reverse(prdKind);
        const size_t smemR = span ? cf::dupire_smem_revs(spanS, m, nCells, nTimes).total : cf::dupire_smem_rev(D, m, nCells).total;
        CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(fwd), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemF)));
        if (aad) CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(rev), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemR)));
        static const bool dbgT = std::getenv("CF_DEBUG_TIMES") != nullptr;
        if (dbgT && !dbgTimes.p) { dbgTimes.alloc(3 * 1024 * 8, s); CF_CUDA(cudaMemsetAsync(dbgTimes.p, 0, 3 * 1024 * 8 * 8, s)); }
        auto ev = takeEvents();
        CF_CUDA(cudaEventRecord(ev.first, s));
        for (uint64_t off = 0; off < n; off += kFastChunk) {
            const uint64_t cnt = std::min<uint64_t>(kFastChunk, n - off);
            cf::DArgs a = dbase;
            a.first_path = first + off; a.n_paths = cnt;
            a.n_pad = (cnt + quantum - 1) / quantum * quantum;
            a.accumulate = off ? 1 : 0;
            a.w[0] = a.w[1] = 0.0;
            if (aad) for (int k = 0; k < nPay && k < cf::kMaxPay; ++k) a.w[k] = w[k];
            a.partial = scratch.partial.p; a.partial_rev = scratch.partialRev.p;
            a.wtab = scratch.wtab.p; a.btab = scratch.btab.p; a.hist = scratch.hist.p; a.state = scratch.state.p; a.live = scratch.live.p;
            a.per_path_payoffs = dPerPath ? dPerPath + off * nPay : nullptr;
            a.per_path_agg = dPerAgg ? dPerAgg + off : nullptr;
            a.n_units = int(a.n_pad / quantum) * 8;
            a.dbg = dbgTimes.p;
            launchKernel(fwd, gridF, fwdWarps * 32, smemF, s, a, false);
            ++g_launches;
            if (aad) {
                launchKernel(rev, gridR, revWarps * 32, smemR, s, a, pdl);
                ++g_launches;
            }
        }
        CF_CUDA(cudaEventRecord(ev.second, s));
        events.push_back(ev);
        const int nOut = int(outSize(aad));
        cf::DPeers pr{};
        if (px) {
            if (size_t(nOut) > px->cap) throw CfError("cf_b200: result vector longer than the communicator's capacity");
            pr = *px;
        }
        {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(unsigned((nOut * 32 + 255) / 256)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = (pdl && aad) ? 1 : 0;
            const double* cPartial = scratch.partial.p; const double* cRev = scratch.partialRev.p; const double* cBtab = scratch.btab.p;
            CF_CUDA(cudaLaunchKernelEx(&cfg, cf::dupire_reduce_kernel, cPartial, gridF, nPay, cRev, cBtab, gridR, m, nTimes,
                                       aad ? 1 : 0, dOut, pr));
        }
        ++g_launches;
    }
};

namespace {

std::unique_ptr<DevPlan> make_plan(const cf_model* mdl, const cf_product* prd, const cf_rng* rng)
{
    ensure_init();
    validate(mdl, prd, rng);
    auto p = std::make_unique<DevPlan>();
    p->dev = t_dev;
    p->spot0 = mdl->spot;
    UploadScope uploads(p->arena);
    p->mdlKind = mdl->kind; p->prdKind = prd->kind; p->rngKind = rng->kind;
    p->D = mdl->n_steps; p->E = mdl->n_events; p->dim = mdl->n_steps * mdl->n_assets;
    p->m = mdl->kind == CF_MODEL_DUPIRE ? mdl->n_knots : 0;
    p->nPay = prd->n_payoffs;
    p->nAdj = adj_size(mdl);
    p->A = mdl->n_assets;
    p->isEvent.upload(mdl->is_event, size_t(p->D) + 1);
    if (mdl->kind == CF_MODEL_DISPLACED) {
        const size_t A = size_t(p->A), D = size_t(p->D), E = size_t(p->E);
        p->lSpots.upload(mdl->dlm_spots, A); p->lChol.upload(mdl->dlm_chol, A * A); p->lAlphas.upload(mdl->dlm_alphas, A);
        p->lDyn.upload(mdl->dlm_dynamics, A);
        p->lDynFwd.upload(mdl->dlm_dyn_fwd, D * A); p->lDrifts.upload(mdl->dlm_drifts, D * A); p->lStds.upload(mdl->dlm_stds, D * A);
        p->lFf.upload(mdl->dlm_fwd_factors, E * A);
        {
            std::vector<double> pack(4 * (D * A + 4), 0.0);        // four spare entries: the kernel steps assets in groups of four
            for (size_t i = 0; i < D * A; ++i) { pack[4 * i] = mdl->dlm_dyn_fwd[i]; pack[4 * i + 1] = mdl->dlm_drifts[i]; pack[4 * i + 2] = mdl->dlm_stds[i]; }
            p->lStepPack.upload(pack.data(), pack.size());
        }
        if (mdl->numeraires) p->lNum.upload(mdl->numeraires, E);
        if (prd->strikes && prd->kind == CF_PRODUCT_BASKETS) p->lStrikes.upload(prd->strikes, size_t(prd->n_payoffs));
        if (prd->weights) p->lPw.upload(prd->weights, A);
        p->lW.alloc(size_t(prd->n_payoffs));
        CF_CUDA(cudaMemset(p->lW.p, 0, sizeof(double) * size_t(prd->n_payoffs)));
    } else if (mdl->kind == CF_MODEL_DUPIRE) {
        p->tabA.upload(mdl->interp_vols, size_t(p->D) * p->m);
        p->tabB.upload(mdl->log_spots, size_t(p->m));
    } else {
        p->tabA.upload(mdl->bs_drifts, size_t(p->D));
        p->tabB.upload(mdl->bs_stds, size_t(p->D));
    }
    if (mdl->kind == CF_MODEL_DUPIRE) {
        // Uniform-cell lookup for the spot bucket: cells no wider than half the smallest knot spacing,
        // entry = number of knots <= left edge of the cell (capped by table size -> binary search).
        const int m = p->m;
        double minDx = 1e300, minVol = 1e300;
        for (int j = 0; j + 1 < m; ++j) minDx = std::min(minDx, mdl->log_spots[j + 1] - mdl->log_spots[j]);
        for (size_t i = 0; i < size_t(p->D) * m; ++i) minVol = std::min(minVol, std::fabs(mdl->interp_vols[i]));
        p->storeG = minVol > 1.0e-6 ? 0 : 1;   // g - v is recovered from log-spot increments unless some vol ~ 0
        const double range = m > 1 ? mdl->log_spots[m - 1] - mdl->log_spots[0] : 0.0;
        if (m > 1 && m < 255 && range / minDx * 2.0 + 2.0 <= 8192.0) {
            const int n = int(std::ceil(range / minDx * 2.0)) + 1;
            const double scale = double(n) / range;   // cell width = range / n <= minDx / 2
            std::vector<uint8_t> lut(size_t(n) + 1);
            for (int c = 0; c <= n; ++c) {
                const double edge = mdl->log_spots[0] + double(c) / scale;
                int ub = 0;
                while (ub < m && mdl->log_spots[ub] <= edge) ++ub;
                lut[size_t(c)] = uint8_t(ub);
            }
            p->lutN = n + 1;
            p->lut.upload(lut.data(), lut.size());
            p->base.lut_x0 = mdl->log_spots[0];
            p->base.lut_scale = scale;
        }
    }
    if (mdl->kind == CF_MODEL_DISPLACED) {
        cf::LArgs& l = p->lbase;
        l.A = p->A; l.D = p->D; l.E = p->E; l.today = mdl->is_event[0] ? 1 : 0;
        l.spots = p->lSpots.p; l.chol = p->lChol.p; l.alphas = p->lAlphas.p; l.dyn = p->lDyn.p;
        for (int k = 0; k < p->A; ++k)
            for (int j = 0; j <= k; ++j) l.cholv[k * (k + 1) / 2 + j] = mdl->dlm_chol[size_t(k) * p->A + j];
        l.step_pack = reinterpret_cast<const double4*>(p->lStepPack.p);
        l.dynFwd = p->lDynFwd.p; l.drifts = p->lDrifts.p; l.stds = p->lStds.p; l.ff = p->lFf.p; l.num = p->lNum.p;
        l.n_payoffs = prd->n_payoffs; l.n_strikes = prd->kind == CF_PRODUCT_BASKETS ? prd->n_payoffs : 0;
        l.strike = prd->strike; l.ko = prd->barrier; l.smooth = prd->smooth; l.coupon = prd->coupon;
        l.cpn_dt = prd->event_dt ? prd->event_dt[0] : 0.0;
        l.has_alpha = 0;
        for (int k = 0; k < p->A; ++k) if (mdl->dlm_dynamics[k] >= 2) l.has_alpha = 1;

        l.strikes = p->lStrikes.p; l.pweights = p->lPw.p;
    }
    if (prd->kind == CF_PRODUCT_EUROPEANS) {
        p->eStrikes.upload(prd->strikes, size_t(prd->n_payoffs));
        p->eOff.upload(prd->strike_offsets, size_t(prd->n_events) + 1);
        p->lW.alloc(size_t(prd->n_payoffs));
        CF_CUDA(cudaMemset(p->lW.p, 0, sizeof(double) * size_t(prd->n_payoffs)));
    }
    if (mdl->numeraires && mdl->kind != CF_MODEL_DISPLACED) p->num.upload(mdl->numeraires, size_t(p->E));
    if (mdl->fwd_factors) p->ff.upload(mdl->fwd_factors, size_t(p->E));
    if (mdl->discounts) p->disc.upload(mdl->discounts, size_t(p->E));
    if (mdl->libors && mdl->kind == CF_MODEL_BS) p->libors.upload(mdl->libors, size_t(p->E));
    if (prd->kind == CF_PRODUCT_CONTINGENT) p->eventDt.upload(prd->event_dt, size_t(prd->n_events) - 1);
    if (rng->kind == CF_RNG_SOBOL) {
        const auto& full = cf::sobol_direction_table();
        const int nd = cf::sobol_max_dim();
        std::vector<uint32_t> sub(size_t(32) * p->dim);
        for (int b = 0; b < 32; ++b)
            for (int d = 0; d < p->dim; ++d) sub[size_t(b) * p->dim + d] = full[size_t(b) * nd + d];
        p->sobolDir.upload(sub.data(), sub.size());
    } else {
        const auto jump = cf::mrg_jump_matrices(uint64_t(p->dim));
        p->mrgJump.upload(jump.data(), jump.size());
    }
    cf::KArgs& a = p->base;
    a.seed1 = rng->seed1; a.seed2 = rng->seed2; a.dim = p->dim;
    a.sobol_dir = p->sobolDir.p; a.mrg_jump = p->mrgJump.p;
    p->lbase.seed1 = rng->seed1; p->lbase.seed2 = rng->seed2; p->lbase.dim = p->dim;
    p->lbase.sobol_dir = p->sobolDir.p; p->lbase.mrg_jump = p->mrgJump.p;
    a.n_steps = p->D; a.n_events = p->E; a.n_knots = p->m;
    a.is_event = p->isEvent.p; a.spot = mdl->spot;
    a.tabA = p->tabA.p; a.tabB = p->tabB.p;
    a.numeraires = p->num.p; a.fwd_factors = p->ff.p; a.discounts = p->disc.p; a.libors = p->libors.p;
    a.lut = p->lut.p; a.lut_n = p->lutN; a.store_g = p->storeG;
    if (mdl->kind == CF_MODEL_DUPIRE) {
        p->hasTimeMap = mdl->n_times > 0 && mdl->time_col1 && mdl->time_col2 && mdl->time_w1 && mdl->time_w2;
        if (p->hasTimeMap) {
            p->nTimes = mdl->n_times;
            for (int i = 0; i < p->D; ++i)
                if (mdl->time_col1[i] < 0 || mdl->time_col1[i] >= p->nTimes || mdl->time_col2[i] < 0 || mdl->time_col2[i] >= p->nTimes)
                    throw CfError("cf_b200: Dupire time map column out of range");
            p->tk1.upload(mdl->time_col1, size_t(p->D)); p->tk2.upload(mdl->time_col2, size_t(p->D));
            p->tc1.upload(mdl->time_w1, size_t(p->D)); p->tc2.upload(mdl->time_w2, size_t(p->D));
        }
        // The fast kernels (cf_dupire.cuh) need: 2..30 knots (32 accumulator rows per lane), vols bounded away
        // from 0 (g - v is recovered by a division), a timeline that ends on an event date, tables that fit.
        const int m = p->m, D = p->D;
        p->fast = m >= 2 && m <= 30 && p->storeG == 0 && mdl->is_event[D] != 0;
        if (p->fast) {
            const double* x = mdl->log_spots;
            double minDx = 1e300;
            for (int j = 0; j + 1 < m; ++j) minDx = std::min(minDx, x[j + 1] - x[j]);
            const double shift = 0.5 * (x[0] + x[m - 1]);      // log-spots are carried relative to the centre of the grid
            const double x0 = x[0] - shift, range = x[m - 1] - x[0];
            // uniform cells no wider than half the smallest knot spacing: at most one knot per cell (+ margin)
            const double width = 0.5 * minDx;
            const double nc = std::floor(range / width) + 3.0;
            if (nc > 1024.0) p->fast = false;
            else {
                const int nCells = int(nc);
                const double scale = 1.0 / width;
                const double delta = 1.0e-9 * width;      // >> rounding of the cell index, << width
                std::vector<double2> cells;
                cells.resize(size_t(nCells));
                for (int c = 0; c < nCells; ++c) {
                    const double edge = x0 + double(c) * width - delta;
                    int ub = 0;
                    while (ub < m && x[ub] - shift <= edge) ++ub;
                    double packed = 0.0;
                    const int64_t bits = int64_t(ub);     // low word = ub0
                    std::memcpy(&packed, &bits, sizeof(double));
                    cells[size_t(c)] = make_double2(ub < m ? x[ub] - shift : DBL_MAX, packed);
                }
                p->nCells = nCells;
                p->cells.upload(cells.data(), cells.size());
                // bucket u = #knots <= X: left knot and 1/width; the two flat buckets have 1/width = 0
                std::vector<double2> bk;
                bk.resize(size_t(m) + 1);
                bk[0] = make_double2(x[0] - shift, 0.0);
                for (int u = 1; u < m; ++u) bk[size_t(u)] = make_double2(x[u - 1] - shift, 1.0 / (x[u] - x[u - 1]));
                bk[size_t(m)] = make_double2(x[m - 1] - shift, 0.0);
                p->bk.upload(bk.data(), bk.size());
                // rows of interpVols as per-bucket lines: vol = A + B * X (interp.h:46-62 restated; flat outside)
                std::vector<double2> ab;
                ab.resize(size_t(D) * (m + 1));
                for (int i = 0; i < D; ++i) {
                    const double* y = mdl->interp_vols + size_t(i) * m;
                    double2* r = ab.data() + size_t(i) * (m + 1);
                    r[0] = make_double2(y[0], 0.0);
                    for (int u = 1; u < m; ++u) {
                        const double B = (y[u] - y[u - 1]) * bk[size_t(u)].y;
                        r[u] = make_double2(y[u - 1] - B * bk[size_t(u)].x, B);
                    }
                    r[m] = make_double2(y[m - 1], 0.0);
                }
                p->ab.upload(ab.data(), ab.size());
                // bit i: timeline point i + 1 is an event date (the last point is handled outside the loops)
                const int nWords = (D + 31) / 32;
                std::vector<uint32_t> bits(size_t(nWords), 0u);
                for (int i = 0; i + 1 < D; ++i)
                    if (mdl->is_event[i + 1]) bits[size_t(i >> 5)] |= 1u << (i & 31);
                p->stepBits.upload(bits.data(), bits.size());
                // reverse sweep: which accumulator component holds which time column (simulated in sweep order)
                std::vector<double2> wxy;
                wxy.resize(size_t(D));
                std::vector<int32_t> colxy(size_t(2) * D, 0);
                std::vector<uint8_t> ops(size_t(D), 0);
                if (p->hasTimeMap) {
                    int cx = -1, cy = -1;
                    for (int i = D - 1; i >= 0; --i) {
                        const int a1 = mdl->time_col1[i], b1 = mdl->time_col2[i];
                        const double c1 = mdl->time_w1[i], c2 = mdl->time_w2[i];
                        if (i == D - 1) { cx = a1; cy = b1; wxy[size_t(i)] = make_double2(c1, c2); }
                        else {
                            const int straight = (cx != a1) + (cy != b1), crossed = (cx != b1) + (cy != a1);
                            if (straight <= crossed) {
                                ops[size_t(i)] = uint8_t((cx != a1 ? 1 : 0) | (cy != b1 ? 2 : 0));
                                cx = a1; cy = b1; wxy[size_t(i)] = make_double2(c1, c2);
                            } else {
                                ops[size_t(i)] = uint8_t((cx != b1 ? 1 : 0) | (cy != a1 ? 2 : 0));
                                cx = b1; cy = a1; wxy[size_t(i)] = make_double2(c2, c1);
                            }
                        }
                        colxy[size_t(2 * i)] = cx; colxy[size_t(2 * i + 1)] = cy;
                    }
                } else {
                    for (int i = 0; i < D; ++i) wxy[size_t(i)] = make_double2(0.0, 0.0);
                }
                p->c12.upload(wxy.data(), wxy.size());
                p->k12.upload(colxy.data(), colxy.size());
                p->flushOps.upload(ops.data(), ops.size());
                // span reverse kernel: lane l sweeps steps [S l, S l + S); in round j the 32 lanes add to the time columns of
                // steps S l + j.  A step has up to two targets (column, weight); they are ordered (A, B) so that within a
                // round no column is the A target of two lanes, nor the B target of two lanes.
                {
                    const int S = cf::dupire_span_steps(D);
                    const int Dp = 32 * S;
                    std::vector<double2> sw(static_cast<size_t>(std::max(Dp, 1)), make_double2(0.0, 0.0));
                    const uint32_t rowBytes = uint32_t(sizeof(double) * cf::kRevSRow);
                    const uint32_t sink = uint32_t(p->nTimes) * rowBytes;       // absent targets (and padding steps) add 0 to a row nobody reads
                    std::vector<uint2> off(static_cast<size_t>(std::max(Dp, 1)), make_uint2(sink, sink));
                    bool ok = S > 0 && p->hasTimeMap && p->nTimes <= 255
                              && cf::dupire_smem_revs(S, m, nCells, p->nTimes).total <= kFastSmemLimit;
                    for (int j = 0; ok && j < S; ++j) {
                        std::vector<int> usedA, usedB;
                        auto free_ = [](const std::vector<int>& used, int col) { return std::find(used.begin(), used.end(), col) == used.end(); };
                        for (int l = 0; l < 32 && ok; ++l) {
                            const int i = S * l + j;
                            if (i >= D) break;
                            int c1 = mdl->time_col1[i], c2 = mdl->time_col2[i];
                            double v1 = mdl->time_w1[i], v2 = mdl->time_w2[i];
                            if (c2 == c1) { v1 += v2; v2 = 0.0; }
                            const bool two = v2 != 0.0;
                            int cA = -1, cB = -1;
                            double vA = 0.0, vB = 0.0;
                            if (two) {
                                if (free_(usedA, c1) && free_(usedB, c2)) { cA = c1; vA = v1; cB = c2; vB = v2; }
                                else if (free_(usedA, c2) && free_(usedB, c1)) { cA = c2; vA = v2; cB = c1; vB = v1; }
                                else ok = false;
                            } else {
                                if (free_(usedA, c1)) { cA = c1; vA = v1; }
                                else if (free_(usedB, c1)) { cB = c1; vB = v1; }
                                else ok = false;
                            }
                            if (!ok) break;
                            if (cA >= 0) usedA.push_back(cA);
                            if (cB >= 0) usedB.push_back(cB);
                            sw[size_t(i)] = make_double2(vA, vB);
                            off[size_t(i)] = make_uint2((cA >= 0 ? uint32_t(cA) * rowBytes : sink) | ((i + 1 < D && mdl->is_event[i + 1]) ? cf::kSpanEvent : 0u),
                                                        cB >= 0 ? uint32_t(cB) * rowBytes : sink);
                        }
                    }
                    if (ok) {
                        p->spanS = S;
                        p->spanW.upload(sw.data(), sw.size());
                        p->spanOff.upload(off.data(), off.size());
                    }
                }
                    const bool sob = rng->kind == CF_RNG_SOBOL;
                if (cf::dupire_smem_fwd4<2, cf::kFwdChunk>(D, m, p->dim, sob, nCells, cf::kFwdWarps).total > kFastSmemLimit
                    || cf::dupire_smem_fwd4<1, cf::kFwdChunk1>(D, m, p->dim, sob, nCells, cf::kFwdWarps).total > kFastSmemLimit
                    || cf::dupire_smem_rev(D, m, nCells).total > kFastSmemLimit) p->fast = false;
                // Moro's branch test |u - 1/2| < 0.42 (gaussians.h:54) as a range of the RNG integer z: u(z) is
                // monotone, so the central set is an interval [lo, hi]; searched with the device's own arithmetic
                // (u = c z for Sobol, z / (m1 + 1) for mrg32k3a; u - 1/2 is exact on both sides of 1/2)
                {
                    auto central = [sob](uint32_t zz) {
                        volatile double u = sob ? CF_ONEOVER2POW32 * double(zz) : double(zz) / 4294967088.0;
                        volatile double xx = u - 0.5;
                        return std::fabs(xx) < 0.42;
                    };
                    uint32_t lo = 0u, hi = 0x80000000u;            // !central(lo), central(hi)
                    while (hi - lo > 1u) { const uint32_t mid = lo + (hi - lo) / 2u; if (central(mid)) hi = mid; else lo = mid; }
                    const uint32_t first = hi;
                    lo = 0x80000000u; hi = 0xffffffffu;           // central(lo), !central(hi)
                    while (hi - lo > 1u) { const uint32_t mid = lo + (hi - lo) / 2u; if (central(mid)) lo = mid; else hi = mid; }
                    p->dbase.tail_lo = first; p->dbase.tail_span = lo - first;
                }
                cf::DArgs& d = p->dbase;
                d.seed1 = rng->seed1; d.seed2 = rng->seed2; d.dim = p->dim;
                d.sobol_dir = p->sobolDir.p; d.mrg_jump = p->mrgJump.p;
                d.n_steps = D; d.n_knots = m; d.n_slots = m + 2; d.n_times = p->nTimes;
                d.ev_bits = p->stepBits.p; d.ev0 = mdl->is_event[0] ? 1 : 0;
                d.spot = mdl->spot; d.shift = shift;
                d.ab = p->ab.p; d.yrows = p->tabA.p; d.bk = p->bk.p; d.cells = p->cells.p; d.n_cells = nCells;
                d.cell_scale = scale; d.cell_off = -x0 * scale;
                d.wxy = p->c12.p; d.colxy = p->k12.p; d.flush_ops = p->flushOps.p;
                d.span_w = p->spanW.p; d.span_off = p->spanOff.p; d.span_S = p->spanS;
                d.n_payoffs = prd->n_payoffs; d.is_put = prd->is_put;
                d.strike = prd->strike; d.barrier = prd->barrier; d.smooth = prd->smooth;
            }
        }
    }
    if (mdl->kind == CF_MODEL_BS && (prd->kind == CF_PRODUCT_EUROPEAN || prd->kind == CF_PRODUCT_UOC)) {
        // fast path (cf_bs.cuh): log-normal steps in log space.  The barrier is monitored on the forward of every sample;
        // the fast kernels monitor the spot, so every forward factor must be 1 (the UOC of mcPrd.h asks for the forward
        // to the sample date itself); the payoff date may carry a forward factor, a discount and a numeraire
        const int D = p->D, E = p->E;
        bool ok = cf::dupire_span_steps(D) > 0 && mdl->spot > 0.0
                  && cf::dupire_smem_fwd4<2, cf::kFwdChunk>(D, 0, p->dim, rng->kind == CF_RNG_SOBOL, 0, cf::kFwdWarps, true).total <= kFastSmemLimit
                  && cf::dupire_smem_fwd4<1, cf::kFwdChunk1>(D, 0, p->dim, rng->kind == CF_RNG_SOBOL, 0, cf::kFwdWarps, true).total <= kFastSmemLimit
                  && cf::bs_smem_rev(D, E).total <= kFastSmemLimit;
        if (prd->kind == CF_PRODUCT_UOC && mdl->fwd_factors)
            for (int e = 0; e < E; ++e) ok = ok && mdl->fwd_factors[e] == 1.0;
        for (int i = 0; i < D; ++i) ok = ok && mdl->bs_stds[i] > 1.0e-12;      // g_i is recovered from the log-spots by a division
        const double numT = mdl->numeraires ? mdl->numeraires[E - 1] : 1.0, ffT = mdl->fwd_factors ? mdl->fwd_factors[E - 1] : 1.0;
        const double discT = mdl->discounts ? mdl->discounts[E - 1] : 1.0;
        ok = ok && numT > 0.0 && ffT > 0.0 && discT > 0.0;
        static const bool off = [] { const char* e = std::getenv("CF_BS_FAST"); return e && std::atoi(e) == 0; }();
        if (ok && !off) {
            std::vector<double2> ds(static_cast<size_t>(D));
            for (int i = 0; i < D; ++i) ds[size_t(i)] = make_double2(mdl->bs_drifts[i], mdl->bs_stds[i]);
            p->bsDs.upload(ds.data(), ds.size());
            const int nWords = (D + 31) / 32;
            std::vector<uint32_t> bits(size_t(nWords), 0u);
            for (int i = 0; i + 1 < D; ++i)
                if (mdl->is_event[i + 1]) bits[size_t(i >> 5)] |= 1u << (i & 31);
            p->stepBits.upload(bits.data(), bits.size());
            cf::DArgs& d = p->dbase;
            d.seed1 = rng->seed1; d.seed2 = rng->seed2; d.dim = p->dim;
            d.sobol_dir = p->sobolDir.p; d.mrg_jump = p->mrgJump.p;
            d.n_steps = D; d.n_knots = 0; d.n_cells = 0; d.n_events = E;
            d.ev_bits = p->stepBits.p; d.ev0 = mdl->is_event[0] ? 1 : 0;
            d.spot = mdl->spot; d.shift = 0.0;
            d.bs_ds = p->bsDs.p; d.fwd_factor = ffT; d.bs_num = numT; d.bs_disc = discT;
            d.pay_scale = prd->kind == CF_PRODUCT_EUROPEAN ? discT / numT : 1.0 / numT;
            d.n_payoffs = prd->n_payoffs; d.is_put = prd->is_put;
            d.strike = prd->strike; d.barrier = prd->barrier; d.smooth = prd->smooth;
            const bool sob = rng->kind == CF_RNG_SOBOL;
            auto central = [sob](uint32_t zz) {
                volatile double u = sob ? CF_ONEOVER2POW32 * double(zz) : double(zz) / 4294967088.0;
                volatile double xx = u - 0.5;
                return std::fabs(xx) < 0.42;
            };
            uint32_t lo = 0u, hi = 0x80000000u;
            while (hi - lo > 1u) { const uint32_t mid = lo + (hi - lo) / 2u; if (central(mid)) hi = mid; else lo = mid; }
            const uint32_t firstC = hi;
            lo = 0x80000000u; hi = 0xffffffffu;
            while (hi - lo > 1u) { const uint32_t mid = lo + (hi - lo) / 2u; if (central(mid)) lo = mid; else hi = mid; }
            d.tail_lo = firstC; d.tail_span = lo - firstC;
            p->bsFast = true;
        }
    }
    a.n_payoffs = prd->n_payoffs; a.is_put = prd->is_put;
    a.strike = prd->strike; a.barrier = prd->barrier; a.smooth = prd->smooth; a.coupon = prd->coupon; a.event_dt = p->eventDt.p;
    a.strikes = p->eStrikes.p; a.strike_off = p->eOff.p;
    if (prd->kind == CF_PRODUCT_EUROPEANS) p->fast = false;      // many payoffs: generic kernel
    if (mdl->kind == CF_MODEL_DUPIRE && prd->kind == CF_PRODUCT_EUROPEANS && p->hasTimeMap && p->D <= cf::kMultiMaxSteps) {
        // itemised risk by strike class (cf_multi.cuh): strikes ascending within each event
        const int nPay = p->nPay, E = p->E;
        std::vector<double> ksorted(static_cast<size_t>(nPay));
        std::vector<int32_t> payEvent(static_cast<size_t>(nPay)), payRank(static_cast<size_t>(nPay)), payOrig(static_cast<size_t>(nPay));
        int cmax = 1;
        for (int e = 0; e < E; ++e) {
            const int k0 = prd->strike_offsets[e], k1 = prd->strike_offsets[e + 1];
            std::vector<int> order;
            for (int k = k0; k < k1; ++k) order.push_back(k);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return prd->strikes[x] < prd->strikes[y]; });
            for (int r = 0; r < k1 - k0; ++r) {
                ksorted[size_t(k0 + r)] = prd->strikes[order[size_t(r)]];
                payEvent[size_t(k0 + r)] = e; payOrig[size_t(k0 + r)] = order[size_t(r)];
            }
            // class of a path = #strikes strictly below S_e; a payoff is in the money for the classes above the position
            // of the LAST strike equal to its own
            for (int r = 0; r < k1 - k0; ++r) {
                int last = r;
                while (last + 1 < k1 - k0 && ksorted[size_t(k0 + last + 1)] == ksorted[size_t(k0 + r)]) ++last;
                payRank[size_t(k0 + r)] = last;
            }
            cmax = std::max(cmax, k1 - k0 + 1);
        }
        p->multiCmax = cmax;
        p->mK.upload(ksorted.data(), ksorted.size());
        p->mEvent.upload(payEvent.data(), payEvent.size()); p->mRank.upload(payRank.data(), payRank.size()); p->mOrig.upload(payOrig.data(), payOrig.size());
        p->multiReady = true;
    }
    uploads.close();
    return p;
}

// ---- RNG parity kernels -----------------------------------------------------------------------
__global__ void sobol_states_kernel(const uint32_t* __restrict__ dir, int dim, uint64_t first, uint64_t n,
                                    uint32_t* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char raw[];
    uint32_t* dirlow = reinterpret_cast<uint32_t*>(raw);
    uint32_t* base = dirlow + size_t(dim) * cf::kLowBits;
    cf::sobol_load_low(dirlow, dir, dim);
    const uint64_t nBatches = (n + cf::kBlock - 1) / cf::kBlock;
    for (uint64_t batch = blockIdx.x; batch < nBatches; batch += gridDim.x) {
        const uint32_t n0 = uint32_t(first + batch * cf::kBlock + 1);
        const uint32_t H0 = n0 >> cf::kLowBits;
        __syncthreads();
        cf::sobol_block_base(base, dir, dim, H0);
        __syncthreads();
        const uint64_t p = batch * cf::kBlock + threadIdx.x;
        cf::SobolThread st;
        st.init(uint32_t(first + p + 1), H0);
        if (p < n)
            for (int d = 0; d < dim; ++d) out[p * dim + d] = st.state(dirlow, base, dim, d);
    }
}

template <int RNGK>
__global__ void rng_draw_kernel(const uint32_t* __restrict__ dir, const uint64_t* __restrict__ jump,
                                uint32_t seed1, uint32_t seed2, int dim, uint64_t first, uint64_t n,
                                int gaussian, double* __restrict__ out, uint32_t* __restrict__ outInt)
{
    extern __shared__ __align__(16) unsigned char raw[];
    uint32_t* dirlow = reinterpret_cast<uint32_t*>(raw);
    uint32_t* base = dirlow + size_t(dim) * cf::kLowBits;
    if (RNGK == CF_RNG_SOBOL) cf::sobol_load_low(dirlow, dir, dim);
    const uint64_t nBatches = (n + cf::kBlock - 1) / cf::kBlock;
    for (uint64_t batch = blockIdx.x; batch < nBatches; batch += gridDim.x) {
        const uint64_t p = batch * cf::kBlock + threadIdx.x;
        const uint64_t pabs = first + p;
        cf::SobolThread st;
        cf::MrgThread mrg;
        if (RNGK == CF_RNG_SOBOL) {
            const uint32_t n0 = uint32_t(first + batch * cf::kBlock + 1);
            const uint32_t H0 = n0 >> cf::kLowBits;
            __syncthreads();
            cf::sobol_block_base(base, dir, dim, H0);
            __syncthreads();
            st.init(uint32_t(pabs + 1), H0);
        } else {
            mrg.init(seed1, seed2, pabs >> 1, jump);
        }
        if (p >= n) continue;
        const bool odd = (pabs & 1ull) != 0;
        for (int d = 0; d < dim; ++d) {
            double u;
            if (RNGK == CF_RNG_SOBOL) u = CF_ONEOVER2POW32 * double(st.state(dirlow, base, dim, d));
            else {
                const uint32_t z = mrg.next();
                if (outInt) { outInt[p * dim + d] = z; continue; }
                u = cf::mrg_uniform(z);
            }
            double r;
            if (gaussian) { r = cf::inv_normal_cdf(u); if (RNGK != CF_RNG_SOBOL && odd) r = -r; }
            else r = (RNGK != CF_RNG_SOBOL && odd) ? 1.0 - u : u;
            out[p * dim + d] = r;
        }
    }
}

__global__ void inv_normal_kernel(const double* __restrict__ p, double* __restrict__ out, uint64_t n)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cf::inv_normal_cdf(p[i]);
}


// every 32-bit numerator: the lean quotient of mrg_uniform against the IEEE division
__global__ void mrg_uniform_selftest_kernel(unsigned long long* mismatches)
{
    unsigned long long bad = 0;
    for (unsigned long long z = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; z < (1ull << 32); z += uint64_t(gridDim.x) * blockDim.x)
        if (cf::mrg_uniform(uint32_t(z)) != cf::mrg_uniform_ieee(uint32_t(z))) ++bad;
    if (bad) atomicAdd(mismatches, bad);
}

// FP64 peak microbenchmark: 8 independent DFMA chains per thread (roofline denominator of bench.py)
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;
}

void rng_run(const cf_rng* rng, int dim, uint64_t first, uint64_t n, int gaussian, double* out, uint32_t* outInt,
             bool statesOnly)
{
    ensure_init();
    if (!rng) throw CfError("cf_b200: null rng");
    if (dim < 1 || n == 0) throw CfError("cf_b200: empty request");
    const bool sobol = rng->kind == CF_RNG_SOBOL;
    if (sobol && dim > cf::sobol_max_dim()) throw CfError("cf_b200: Sobol dimension exceeds 1101");
    if (sobol && first + n > 0xffffffffull) throw CfError("cf_b200: Sobol index exceeds 2^32 - 1");
    if (!sobol && first + n > (1ull << 32)) throw CfError("cf_b200: mrg32k3a path index exceeds 2^32 - 1 (skipTo takes an unsigned, mrg32k3a.h:192)");
    DevBuf<uint32_t> dDir, dInt;
    DevBuf<uint64_t> dJump;
    DevBuf<double> dOut;
    std::vector<uint32_t> sub;
    std::vector<uint64_t> jump;
    if (sobol) {
        const auto& full = cf::sobol_direction_table();
        const int nd = cf::sobol_max_dim();
        sub.resize(size_t(32) * dim);
        for (int b = 0; b < 32; ++b)
            for (int d = 0; d < dim; ++d) sub[size_t(b) * dim + d] = full[size_t(b) * nd + d];
        dDir.upload(sub.data(), sub.size());
    } else {
        jump = cf::mrg_jump_matrices(uint64_t(dim));
        dJump.upload(jump.data(), jump.size());
    }
    if (out) dOut.alloc(n * dim);
    if (outInt) dInt.alloc(n * dim);
    const uint64_t nBatches = (n + cf::kBlock - 1) / cf::kBlock;
    const int grid = int(std::min<uint64_t>(nBatches, uint64_t(8) * g_sms));
    const size_t smem = sobol ? sizeof(uint32_t) * size_t(dim) * (cf::kLowBits + 2) : 0;
    if (smem > 200 * 1024) throw CfError("cf_b200: dimension too large");
    if (statesOnly) {
        CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(sobol_states_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        sobol_states_kernel<<<grid, cf::kBlock, smem>>>(dDir.p, dim, first, n, dInt.p);
    } else if (sobol) {
        CF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(rng_draw_kernel<CF_RNG_SOBOL>), cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        rng_draw_kernel<CF_RNG_SOBOL><<<grid, cf::kBlock, smem>>>(dDir.p, nullptr, 0, 0, dim, first, n, gaussian, dOut.p, nullptr);
    } else {
        rng_draw_kernel<CF_RNG_MRG32K3A><<<grid, cf::kBlock, 0>>>(nullptr, dJump.p, rng->seed1, rng->seed2, dim, first, n, gaussian, dOut.p, dInt.p);
    }
    ++g_launches;
    CF_CUDA(cudaGetLastError());
    if (out) CF_CUDA(cudaMemcpy(out, dOut.p, n * dim * sizeof(double), cudaMemcpyDeviceToHost));
    if (outInt) CF_CUDA(cudaMemcpy(outInt, dInt.p, n * dim * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CF_CUDA(cudaDeviceSynchronize());
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// The public plan: one DevPlan per device of the context.
struct cf_plan;
namespace { std::vector<cf_plan*>& live_plans() { static std::vector<cf_plan*> v; return v; } }

struct cf_plan {
    std::vector<std::unique_ptr<DevPlan>> dev;   // empty once the context the plan was created in has been closed
    double* hostOut = nullptr;            // pinned staging of the results of host-buffer runs
    size_t hostCap = 0;

    cf_plan() { live_plans().push_back(this); }
    DevPlan& d0()
    {
        if (dev.empty()) throw CfError("cf_b200: this plan belongs to a context that has been closed (cf_init / cf_shutdown)");
        return *dev[0];
    }
    const DevPlan& d0() const { return const_cast<cf_plan*>(this)->d0(); }
    // frees the device side (tables, scratch) on the devices it lives on
    void retire()
    {
        for (auto& p : dev) {
            if (!p) continue;
            cudaSetDevice(p->dev->id);
            cudaDeviceSynchronize();      // launches may be in flight on the caller's streams; the tables go back to the pool
            p.reset();
        }
        dev.clear();
    }
    double* pinned(size_t n)
    {
        if (hostCap < n) {
            g_pinned.give(hostOut, hostCap);
            hostOut = nullptr; hostCap = 0;
            hostOut = g_pinned.take(n, hostCap);
        }
        return hostOut;
    }
    ~cf_plan()
    {
        retire();
        g_pinned.give(hostOut, hostCap);
        if (t_dev) cudaSetDevice(t_dev->id);
        auto& v = live_plans();
        v.erase(std::remove(v.begin(), v.end(), this), v.end());
    }
};

namespace {

void retire_plans() { for (cf_plan* p : live_plans()) p->retire(); }

std::unique_ptr<cf_plan> make_multi_plan(const cf_model* mdl, const cf_product* prd, const cf_rng* rng)
{
    ensure_init();
    auto mp = std::make_unique<cf_plan>();
    for (auto& d : g_devs) {
        DeviceScope sc(d.get());
        mp->dev.push_back(make_plan(mdl, prd, rng));
    }
    return mp;
}

// Runs `job(k)` for the local devices k < nLocal at the same time: device 0 on the calling thread, the others on
// their worker threads.
template <class F>
void on_devices(int nLocal, F&& job)
{
    for (int k = 1; k < nLocal; ++k) g_devs[size_t(k)]->worker->post([&job, k] { job(k); });
    std::exception_ptr first;
    try { job(0); } catch (...) { first = std::current_exception(); }
    for (int k = 1; k < nLocal; ++k) {
        try { g_devs[size_t(k)]->worker->wait(); } catch (...) { if (!first) first = std::current_exception(); }
    }
    if (first) std::rethrow_exception(first);
}

enum class RunKind { Value, Aad, Multi };
double g_lastKernelMs = 0.0;

// One run of a plan over paths [first, first + n) with host results.  With a communicator the range is sharded over
// its participants -- the devices of this context, or this process's device among the processes of the job -- and the
// rank sum is part of the launch; every participant ends with the full sums.
//   hOut: outSize doubles ([n_payoffs] | [n_payoffs] agg adjoints | multi: [n_payoffs] then [nAdj][n_payoffs])
void run_plan(cf_plan& mp, RunKind kind, const double* w, uint64_t first, uint64_t n, double* hOut, double* hPerPath,
              double* hPerAgg)
{
    ensure_init();
    g_comm.check();
    DevPlan& p0 = mp.d0();
    if (mp.dev.size() != g_devs.size()) throw CfError("cf_b200: this plan was created in another context");
    const bool aad = kind != RunKind::Value;
    const size_t nOut = kind == RunKind::Multi ? p0.multiOutSize() : p0.outSize(aad);
    const int nPay = p0.nPay;
    const bool shard = g_comm.on();
    if (shard && g_comm.ipc && (hPerPath || hPerAgg))
        throw CfError("cf_b200: per-path outputs are not gathered across the processes of a communicator");
    if (shard && nOut > g_comm.cap) throw CfError("cf_b200: result vector longer than the communicator's capacity");
    const int nLocal = shard ? int(mp.dev.size()) : 1;
    const uint32_t epoch = shard ? ++g_comm.epoch : 0u;
    double* pinned = mp.pinned(nOut);
    auto job = [&](int k) {
        DevPlan& p = *mp.dev[size_t(k)];
        cudaStream_t s = p.dev->stream;                      // null in a single-device context: the legacy default stream
        uint64_t f = 0, c = n;
        cf::DPeers px{};
        if (shard) { shard_range(n, g_comm.rank0 + k, g_comm.world, f, c); px = g_comm.peers(k, epoch); }
        if (p.outBuf.n < nOut) p.outBuf.alloc(nOut, s);
        double* dPer = nullptr;
        double* dAgg = nullptr;
        if (hPerPath && c) { if (p.perBuf.n < c * size_t(nPay)) p.perBuf.alloc(c * size_t(nPay), s); dPer = p.perBuf.p; }
        if (hPerAgg && c) { if (p.aggBuf.n < c) p.aggBuf.alloc(c, s); dAgg = p.aggBuf.p; }
        if (kind == RunKind::Multi) p.launchMulti(first + f, c, p.outBuf.p, s, shard ? &px : nullptr);
        else p.launch(aad, w, first + f, c, p.outBuf.p, dPer, dAgg, s, shard ? &px : nullptr);
        if (k == 0) CF_CUDA(cudaMemcpyAsync(pinned, p.outBuf.p, sizeof(double) * nOut, cudaMemcpyDeviceToHost, s));
        if (dPer) CF_CUDA(cudaMemcpyAsync(hPerPath + f * size_t(nPay), dPer, sizeof(double) * c * size_t(nPay), cudaMemcpyDeviceToHost, s));
        if (dAgg) CF_CUDA(cudaMemcpyAsync(hPerAgg + f, dAgg, sizeof(double) * c, cudaMemcpyDeviceToHost, s));
        if (k == 0 || dPer || dAgg) CF_CUDA(cudaStreamSynchronize(s));
    };
    on_devices(nLocal, job);
    g_comm.check();
    std::memcpy(hOut, pinned, sizeof(double) * nOut);
    // device time of the path kernels of this run on the first local device (cf_last_run_kernel_ms)
    if (!p0.events.empty()) {
        float ms = 0.f;
        const auto& ev = p0.events.back();
        if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) g_lastKernelMs = ms;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

const char* cf_last_error(void) { return g_err.c_str(); }
uint64_t cf_launch_count(void) { return g_launches.load(); }

int cf_init(int n_devices, const int* device_ids)
{
    return guarded([&] {
        if (n_devices < 1 || n_devices > cf::kMaxPeers || !device_ids) throw CfError("cf_init: 1 .. 16 devices");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) throw CfError("cf_init: no CUDA device available (this library has no CPU fallback)");
        for (int k = 0; k < n_devices; ++k) {
            if (device_ids[k] < 0 || device_ids[k] >= n) throw CfError("cf_init: device id out of range");
            for (int j = 0; j < k; ++j) if (device_ids[j] == device_ids[k]) throw CfError("cf_init: a device is listed twice");
        }
        if (g_comm.world) comm_destroy();
        close_devices();
        for (int k = 0; k < n_devices; ++k) g_devs.push_back(open_device(device_ids[k], k, n_devices > 1));
        t_dev = nullptr; t_gen = g_gen;
        bind(g_devs[0].get());
        g_launches = 0;
        if (n_devices > 1) {
            // the devices of this process are the participants of the rank sum: peer access, receive blocks, flags
            for (int k = 0; k < n_devices; ++k) {
                DeviceScope sc(g_devs[size_t(k)].get());
                for (int j = 0; j < n_devices; ++j) {
                    if (j == k) continue;
                    int can = 0;
                    CF_CUDA(cudaDeviceCanAccessPeer(&can, device_ids[k], device_ids[j]));
                    if (!can) throw CfError("cf_init: the devices of a context must have peer access to each other (NVLink)");
                    const cudaError_t pe = cudaDeviceEnablePeerAccess(device_ids[j], 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CF_CUDA(pe);
                    cudaGetLastError();
                }
            }
            static const size_t cap = [] { const char* v = std::getenv("CF_COMM_CAPACITY"); return v ? size_t(std::atoll(v)) : (size_t(1) << 20); }();
            comm_allocate(n_devices, 0, cap);
            for (int k = 0; k < n_devices; ++k) g_comm.block[k] = g_comm.local[size_t(k)];
            g_comm.enabled = true;
        }
    });
}

int cf_shutdown(void)
{
    return guarded([&] {
        if (g_comm.world) comm_destroy();
        close_devices();
    });
}

int cf_device_count(void) { return int(g_devs.size()); }
int cf_context_generation(void) { return g_gen; }

/* ---- communicator over the processes of a job (one device per process) ---- */
int cf_comm_create(int world, int rank, size_t capacity_doubles, void* handle_out)
{
    return guarded([&] {
        ensure_init();
        if (!handle_out) throw CfError("cf_comm_create: null handle");
        if (g_devs.size() != 1) throw CfError("cf_comm_create: a process of a multi-process job drives one device");
        if (rank < 0 || rank >= world) throw CfError("cf_comm_create: bad rank");
        comm_allocate(world, rank, capacity_doubles);
        g_comm.ipc = true;
        cudaIpcMemHandle_t h;
        CF_CUDA(cudaIpcGetMemHandle(&h, g_comm.local[0]));
        static_assert(sizeof(cudaIpcMemHandle_t) == CF_COMM_HANDLE_BYTES, "handle size");
        std::memcpy(handle_out, &h, sizeof(h));
    });
}

int cf_comm_connect(const void* handles)
{
    return guarded([&] {
        ensure_init();
        if (!g_comm.world || !g_comm.ipc || !handles) throw CfError("cf_comm_connect: call cf_comm_create first");
        for (int r = 0; r < g_comm.world; ++r) {
            if (r == g_comm.rank0) { g_comm.block[r] = g_comm.local[0]; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, static_cast<const unsigned char*>(handles) + size_t(r) * CF_COMM_HANDLE_BYTES, sizeof(h));
            void* q = nullptr;
            CF_CUDA(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
            g_comm.opened.push_back(q);
            g_comm.block[r] = static_cast<unsigned char*>(q);
        }
        g_comm.enabled = true;
    });
}

int cf_comm_enable(int on)
{
    return guarded([&] {
        if (on && g_comm.world < 2) throw CfError("cf_comm_enable: no communicator");
        if (on) for (int r = 0; r < g_comm.world; ++r) if (!g_comm.block[r]) throw CfError("cf_comm_enable: the communicator is not connected");
        g_comm.enabled = on != 0;
    });
}

int cf_comm_destroy(void) { return guarded([&] { comm_destroy(); }); }

int cf_comm_info(int* world, int* rank, int* enabled)
{
    if (world) *world = g_comm.world;
    if (rank) *rank = g_comm.rank0;
    if (enabled) *enabled = g_comm.on() ? 1 : 0;
    return 0;
}

int cf_comm_status(void) { return g_comm.status ? *reinterpret_cast<volatile int*>(g_comm.status) : 0; }

int cf_shard_range(uint64_t n_paths, int rank, int world, uint64_t* first, uint64_t* count)
{
    return guarded([&] {
        if (world < 1 || rank < 0 || rank >= world || !first || !count) throw CfError("cf_shard_range: bad arguments");
        shard_range(n_paths, rank, world, *first, *count);
    });
}

size_t cf_table_adjoint_size(const cf_model* mdl, const cf_product* prd)
{
    (void)prd;
    try { return mdl ? adj_size(mdl) : 0; } catch (const std::exception& e) { g_err = e.what(); return 0; }
}

int cf_plan_create(const cf_model* mdl, const cf_product* prd, const cf_rng* rng, cf_plan** out)
{
    return guarded([&] {
        if (!out) throw CfError("cf_plan_create: null out");
        *out = make_multi_plan(mdl, prd, rng).release();
    });
}

void cf_plan_destroy(cf_plan* plan) { delete plan; }

size_t cf_plan_out_size(const cf_plan* plan, int aad) { return plan ? plan->d0().outSize(aad != 0) : 0; }

namespace {
// the participants of the rank sum for a launch on the caller's stream (one local device per process)
const cf::DPeers* launch_peers(cf::DPeers& px)
{
    if (!g_comm.on()) return nullptr;
    if (g_devs.size() != 1) throw CfError("cf_plan_launch: a multi-device context runs through cf_plan_run_* / cf_run_*");
    g_comm.check();
    px = g_comm.peers(0, ++g_comm.epoch);
    return &px;
}
}  // namespace

int cf_plan_launch_value(cf_plan* plan, uint64_t first_path, uint64_t n_paths, double* d_out, void* stream)
{
    return guarded([&] {
        if (!plan || !d_out) throw CfError("cf_plan_launch_value: null argument");
        ensure_init();
        cf::DPeers px{};
        plan->d0().launch(false, nullptr, first_path, n_paths, d_out, nullptr, nullptr, static_cast<cudaStream_t>(stream), launch_peers(px));
    });
}

int cf_plan_launch_aad(cf_plan* plan, const double* payoff_weights, uint64_t first_path, uint64_t n_paths,
                       double* d_out, void* stream)
{
    return guarded([&] {
        if (!plan || !d_out || !payoff_weights) throw CfError("cf_plan_launch_aad: null argument");
        ensure_init();
        cf::DPeers px{};
        plan->d0().launch(true, payoff_weights, first_path, n_paths, d_out, nullptr, nullptr, static_cast<cudaStream_t>(stream), launch_peers(px));
    });
}

int cf_plan_run_value(cf_plan* plan, uint64_t first_path, uint64_t n_paths, double* payoff_sums, double* per_path_payoffs)
{
    return guarded([&] {
        if (!plan || !payoff_sums) throw CfError("cf_plan_run_value: null argument");
        std::vector<double> h(plan->d0().outSize(false));
        run_plan(*plan, RunKind::Value, nullptr, first_path, n_paths, h.data(), per_path_payoffs, nullptr);
        std::memcpy(payoff_sums, h.data(), sizeof(double) * size_t(plan->d0().nPay));
    });
}

int cf_plan_run_aad(cf_plan* plan, const double* payoff_weights, uint64_t first_path, uint64_t n_paths,
                    double* payoff_sums, double* agg_sum, double* table_adjoints, double* per_path_payoffs, double* per_path_agg)
{
    return guarded([&] {
        if (!plan || !payoff_weights || !payoff_sums || !agg_sum || !table_adjoints) throw CfError("cf_plan_run_aad: null argument");
        const DevPlan& p = plan->d0();
        std::vector<double> h(p.outSize(true));
        run_plan(*plan, RunKind::Aad, payoff_weights, first_path, n_paths, h.data(), per_path_payoffs, per_path_agg);
        std::memcpy(payoff_sums, h.data(), sizeof(double) * size_t(p.nPay));
        *agg_sum = h[size_t(p.nPay)];
        std::memcpy(table_adjoints, h.data() + p.nPay + 1, sizeof(double) * p.nAdj);
    });
}

double cf_last_run_kernel_ms(void) { return g_lastKernelMs; }

int cf_plan_debug_times(cf_plan* plan, unsigned long long* out /* [3][1024][8] */)
{
    return guarded([&] {
        if (!plan || !out || !plan->d0().dbgTimes.p) throw CfError("cf_plan_debug_times: run with CF_DEBUG_TIMES=1 first");
        ensure_init();
        CF_CUDA(cudaDeviceSynchronize());
        CF_CUDA(cudaMemcpy(out, plan->d0().dbgTimes.p, 3 * 1024 * 8 * 8, cudaMemcpyDeviceToHost));
    });
}

int cf_plan_kernel_ms(cf_plan* plan, double* avg_ms, int* n_launches)
{
    return guarded([&] {
        if (!plan) throw CfError("cf_plan_kernel_ms: null plan");
        ensure_init();
        DevPlan& p = plan->d0();
        double tot = 0.0;
        int n = 0;
        for (auto& ev : p.events) {
            CF_CUDA(cudaEventSynchronize(ev.second));
            float ms = 0.f;
            CF_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
            tot += ms; ++n;
            p.pool.push_back(ev);
        }
        p.events.clear();
        if (avg_ms) *avg_ms = n ? tot / n : 0.0;
        if (n_launches) *n_launches = n;
    });
}

int cf_run_value(const cf_model* mdl, const cf_product* prd, const cf_rng* rng, uint64_t first_path,
                 uint64_t n_paths, double* payoff_sums, double* per_path_payoffs)
{
    return guarded([&] {
        if (!payoff_sums) throw CfError("cf_run_value: payoff_sums is null");
        auto plan = make_multi_plan(mdl, prd, rng);
        std::vector<double> h(plan->d0().outSize(false));
        run_plan(*plan, RunKind::Value, nullptr, first_path, n_paths, h.data(), per_path_payoffs, nullptr);
        std::memcpy(payoff_sums, h.data(), sizeof(double) * size_t(plan->d0().nPay));
    });
}

int cf_run_aad(const cf_model* mdl, const cf_product* prd, const cf_rng* rng, uint64_t first_path,
               uint64_t n_paths, const double* payoff_weights, double* payoff_sums, double* agg_sum,
               double* table_adjoints, double* per_path_payoffs, double* per_path_agg)
{
    return guarded([&] {
        if (!payoff_weights || !payoff_sums || !agg_sum || !table_adjoints) throw CfError("cf_run_aad: null output");
        static const bool timing = std::getenv("CF_TIMING") != nullptr;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        const auto t0 = now();
        auto plan = make_multi_plan(mdl, prd, rng);
        const auto t1 = now();
        const DevPlan& p = plan->d0();
        std::vector<double> h(p.outSize(true));
        run_plan(*plan, RunKind::Aad, payoff_weights, first_path, n_paths, h.data(), per_path_payoffs, per_path_agg);
        if (timing) std::fprintf(stderr, "cf_run_aad: make_plan %.0f us, launches + wait + copy back %.0f us\n", us(t0, t1), us(t1, now()));
        std::memcpy(payoff_sums, h.data(), sizeof(double) * size_t(p.nPay));
        *agg_sum = h[size_t(p.nPay)];
        std::memcpy(table_adjoints, h.data() + p.nPay + 1, sizeof(double) * p.nAdj);
    });
}

int cf_run_aad_multi(const cf_model* mdl, const cf_product* prd, const cf_rng* rng, uint64_t first_path,
                     uint64_t n_paths, double* payoff_sums, double* risk_tables)
{
    return guarded([&] {
        auto plan = make_multi_plan(mdl, prd, rng);
        if (cf_plan_run_aad_multi(plan.get(), first_path, n_paths, payoff_sums, risk_tables) != 0) throw CfError(g_err);
    });
}

int cf_plan_run_aad_multi(cf_plan* plan, uint64_t first_path, uint64_t n_paths, double* payoff_sums, double* risk_tables)
{
    return guarded([&] {
        if (!plan || !payoff_sums || !risk_tables) throw CfError("cf_run_aad_multi: null argument");
        const DevPlan& p = plan->d0();
        const int nPay = p.nPay;
        const size_t nAdj = p.nAdj;
        if (n_paths == 0) throw CfError("cf_b200: n_paths must be > 0");
        if (p.multiReady) {
            // Dupire x Europeans: one sweep per maturity, accumulated by strike class (cf_multi.cuh)
            std::vector<double> h(p.multiOutSize());
            run_plan(*plan, RunKind::Multi, nullptr, first_path, n_paths, h.data(), nullptr, nullptr);
            std::memcpy(payoff_sums, h.data(), sizeof(double) * size_t(nPay));
            std::memcpy(risk_tables, h.data() + nPay, sizeof(double) * nAdj * size_t(nPay));
            return;
        }
        // any other pair: one adjoint sweep per payoff, column k of the risk matrix is the aggregate risk with weights e_k
        std::vector<double> h(p.outSize(true)), w(size_t(nPay), 0.0);
        for (int k = 0; k < nPay; ++k) {
            std::fill(w.begin(), w.end(), 0.0);
            w[size_t(k)] = 1.0;
            run_plan(*plan, RunKind::Aad, w.data(), first_path, n_paths, h.data(), nullptr, nullptr);
            if (k == 0) std::memcpy(payoff_sums, h.data(), sizeof(double) * size_t(nPay));
            for (size_t q = 0; q < nAdj; ++q) risk_tables[q * size_t(nPay) + size_t(k)] = h[size_t(nPay) + 1 + q];
        }
    });
}

int cf_sobol_states(int dim, uint64_t first_path, uint64_t n_paths, uint32_t* out)
{
    return guarded([&] {
        if (!out) throw CfError("cf_sobol_states: null out");
        cf_rng r{CF_RNG_SOBOL, 0, 0};
        rng_run(&r, dim, first_path, n_paths, 0, nullptr, out, true);
    });
}

uint32_t cf_sobol_direction_number(int bit, int dim)
{
    if (bit < 0 || bit >= 32 || dim < 0 || dim >= cf::sobol_max_dim()) return 0;
    return cf::sobol_direction_table()[size_t(bit) * cf::sobol_max_dim() + dim];
}

int cf_sobol_max_dim(void) { return cf::sobol_max_dim(); }

int cf_rng_draw(const cf_rng* rng, int dim, uint64_t first_path, uint64_t n_paths, int gaussian, double* out)
{
    return guarded([&] {
        if (!out) throw CfError("cf_rng_draw: null out");
        rng_run(rng, dim, first_path, n_paths, gaussian, out, nullptr, false);
    });
}

int cf_mrg_numerators(const cf_rng* rng, int dim, uint64_t first_path, uint64_t n_paths, uint32_t* out)
{
    return guarded([&] {
        if (!out || !rng || rng->kind != CF_RNG_MRG32K3A) throw CfError("cf_mrg_numerators: needs an mrg32k3a rng and out");
        rng_run(rng, dim, first_path, n_paths, 0, nullptr, out, false);
    });
}


int cf_measure_fp64_peak(double* tflops, double* ms_out)
{
    return guarded([&] {
        ensure_init();
        DevBuf<double> d; d.alloc(1);
        const int iters = 1 << 14, grid = g_sms * 8, block = 256;
        cudaEvent_t e0, e1;
        CF_CUDA(cudaEventCreate(&e0)); CF_CUDA(cudaEventCreate(&e1));
        double best = 1e30;
        for (int rep = 0; rep < 5; ++rep) {
            CF_CUDA(cudaEventRecord(e0));
            fp64_peak_kernel<<<grid, block>>>(d.p, iters, 0.999999, 1e-7);
            CF_CUDA(cudaEventRecord(e1));
            CF_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            CF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
            ++g_launches;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        const double flops = double(grid) * block * 8.0 * iters * 2.0;
        if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
        if (ms_out) *ms_out = best;
    });
}

int cf_selftest_mrg_uniform(uint64_t* mismatches)
{
    return guarded([&] {
        ensure_init();
        if (!mismatches) throw CfError("cf_selftest_mrg_uniform: null argument");
        DevBuf<unsigned long long> d; d.alloc(1);
        CF_CUDA(cudaMemset(d.p, 0, sizeof(unsigned long long)));
        mrg_uniform_selftest_kernel<<<g_sms * 8, 256>>>(d.p);
        ++g_launches;
        CF_CUDA(cudaGetLastError());
        unsigned long long h = 0;
        CF_CUDA(cudaMemcpy(&h, d.p, sizeof(h), cudaMemcpyDeviceToHost));
        *mismatches = h;
    });
}

int cf_device_sm_count(void) { try { ensure_init(); return g_sms; } catch (const std::exception& e) { g_err = e.what(); return -1; } }

int cf_inv_normal(const double* p, double* out, uint64_t n)
{
    return guarded([&] {
        ensure_init();
        if (!p || !out) throw CfError("cf_inv_normal: null argument");
        if (n == 0) return;
        DevBuf<double> dIn, dOut;
        dIn.upload(p, n);
        dOut.alloc(n);
        inv_normal_kernel<<<unsigned((n + 255) / 256), 256>>>(dIn.p, dOut.p, n);
        ++g_launches;
        CF_CUDA(cudaGetLastError());
        CF_CUDA(cudaMemcpy(out, dOut.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
