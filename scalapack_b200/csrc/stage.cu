// stage.cu -- host <-> HBM transfer tickets (see stage.h).  Host-side plumbing of the drop-in boundary: the reference
// works on the caller's host array in place (SRC/pdgetrf.f:1); here the array is streamed through HBM while the
// factorisation runs, so the PCIe time hides under the FP64 update instead of being paid before and after it.
#include "stage.h"

#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>

namespace slb {

bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

namespace {

// ---- worker pool: one 2-D memcpy split over several host threads ------------------------------------------------
class CopyPool {
public:
    explicit CopyPool(int nthreads)
    {
        for (int i = 0; i < nthreads; ++i) th_.emplace_back([this] { worker(); });
    }
    // dst[j*dpitch .. +width) = src[j*spitch .. +width) for j in [0, n); the calling thread takes part
    void copy2d(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, int64_t n)
    {
        if (n <= 0 || width == 0) return;
        if (dpitch == width && spitch == width) { width *= (size_t)n; n = 1; dpitch = spitch = width; }
        // pieces of >= 256 KiB: a long column is cut along its length, short columns are grouped
        int64_t split = 1, group = 1;
        if (width >= (size_t)(512 << 10)) split = (int64_t)((width + (256 << 10) - 1) / (256 << 10));
        else group = (int64_t)std::max<size_t>(1, (256 << 10) / width);
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = { dst, dpitch, src, spitch, width, n, split, group };
            next_.store(0); total_ = split > 1 ? n * split : (n + group - 1) / group;
            left_.store((int)th_.size());
            ++gen_;
        }
        cv_.notify_all();
        run();
        std::unique_lock<std::mutex> lk(mu_);
        dcv_.wait(lk, [this] { return left_.load() == 0; });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
private:
    struct Job { char *dst; size_t dpitch; const char *src; size_t spitch; size_t width; int64_t n, split, group; };
    void run()
    {
        const Job j = job_;
        for (;;) {
            int64_t i = next_.fetch_add(1);
            if (i >= total_) break;
            if (j.split > 1) {
                const int64_t col = i / j.split, part = i % j.split;
                const size_t piece = (j.width + j.split - 1) / j.split, off = (size_t)part * piece;
                if (off < j.width) memcpy(j.dst + col * j.dpitch + off, j.src + col * j.spitch + off, std::min(piece, j.width - off));
            } else {
                const int64_t c0 = i * j.group, c1 = std::min(j.n, c0 + j.group);
                for (int64_t c = c0; c < c1; ++c) memcpy(j.dst + c * j.dpitch, j.src + c * j.spitch, j.width);
            }
        }
    }
    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            run();
            if (left_.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(mu_); dcv_.notify_all(); }
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_; std::condition_variable cv_, dcv_;
    Job job_{}; std::atomic<int64_t> next_{ 0 }; int64_t total_ = 0; std::atomic<int> left_{ 0 };
    uint64_t gen_ = 0; bool stop_ = false;
};

// ---- one I/O thread per direction: runs the queued pageable transfers in order ----------------------------------
struct Slot { char *buf = nullptr; cudaEvent_t ev = nullptr; };
constexpr int NSLOT = 4;

class IoThread {
public:
    IoThread(int nworkers, size_t slot_bytes) : pool_(nworkers), slot_bytes_(slot_bytes)
    {
        th_ = std::thread([this] { loop(); });
    }
    void submit(std::function<void(IoThread &)> f)
    {
        { std::lock_guard<std::mutex> lk(mu_); q_.push_back(std::move(f)); }
        cv_.notify_one();
    }
    CopyPool &pool() { return pool_; }
    Slot &slot(int i)
    {
        Slot &s = slots_[i % NSLOT];
        if (!s.buf) {
            SLB_CUDA(cudaHostAlloc((void **)&s.buf, slot_bytes_, cudaHostAllocDefault));
            SLB_CUDA(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        }
        return s;
    }
    size_t slot_bytes() const { return slot_bytes_; }
private:
    void loop()
    {
        SLB_CUDA(cudaSetDevice(rt().device));
        for (;;) {
            std::function<void(IoThread &)> f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !q_.empty(); });
                f = std::move(q_.front()); q_.pop_front();
            }
            f(*this);
        }
    }
    CopyPool pool_;
    size_t slot_bytes_;
    Slot slots_[NSLOT];
    std::thread th_;
    std::mutex mu_; std::condition_variable cv_;
    std::deque<std::function<void(IoThread &)>> q_;
};

int copy_threads()
{
    int64_t forced = opt("copy_threads", 0);
    if (forced > 0) return (int)forced;
    unsigned hw = std::thread::hardware_concurrency();
    int local = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) local = atoi(e);
    else if (const char *e2 = getenv("WORLD_SIZE")) local = atoi(e2);
    if (local < 1) local = 1;
    int per = (int)(hw ? hw : 8) / local / 2;       // two directions share the process's cores
    return per < 1 ? 1 : (per > 8 ? 8 : per);
}

IoThread &io_thread(int dir)
{
    // leaked on purpose: the threads live as long as the process (joining them from a static destructor would race
    // with the CUDA runtime's own teardown)
    static IoThread *t[2] = { nullptr, nullptr };
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!t[dir]) t[dir] = new IoThread(copy_threads() - 1, (size_t)opt("bounce_mb", 32) << 20);
    return *t[dir];
}

struct Ticket {
    cudaEvent_t ev = nullptr;
    std::atomic<int> state{ 0 };      // 0 queued, 1 device event recorded, 2 host side complete (pageable download)
    bool download = false;
};

}  // namespace

// test hook (pure host code): one 2-D copy through a worker pool of `threads` threads (the pageable-memory path of HostLink)
void test_copy2d(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, int64_t n, int threads)
{
    CopyPool pool(threads > 1 ? threads - 1 : 0);
    for (int rep = 0; rep < 3; ++rep) pool.copy2d(dst, dpitch, src, spitch, width, n);       // the pool is reused job after job
}

struct HostLink::Impl {
    std::deque<Ticket> tickets;       // stable addresses; only the issuing thread appends
};

HostLink::HostLink(const HostMat &h, void *dev, int64_t ldd) : im_(new Impl), h_(h), dev_(dev), ldd_(ldd) {}

HostLink::~HostLink()
{
    finish();
    for (auto &t : im_->tickets) if (t.ev) cudaEventDestroy(t.ev);
}

int HostLink::upload(int64_t r0, int64_t r1, int64_t c0, int64_t c1)
{
    Runtime &r = rt();
    im_->tickets.emplace_back();
    Ticket *t = &im_->tickets.back();
    const int id = (int)im_->tickets.size() - 1;
    SLB_CUDA(cudaEventCreateWithFlags(&t->ev, cudaEventDisableTiming));
    const int64_t nr = r1 - r0, nc = c1 - c0;
    if (nr <= 0 || nc <= 0) { SLB_CUDA(cudaEventRecord(t->ev, r.s_h2d)); t->state.store(1); return id; }
    const size_t es = h_.elem;
    char *hp = (char *)h_.p + (size_t)(r0 + c0 * h_.ld) * es;
    char *dp = (char *)dev_ + (size_t)(r0 + c0 * ldd_) * es;
    const size_t hpitch = (size_t)h_.ld * es, dpitch = (size_t)ldd_ * es, width = (size_t)nr * es;
    up_bytes_ += (int64_t)(width * (size_t)nc);
    counter_add("h2d_bytes", (int64_t)(width * (size_t)nc));
    if (h_.pinned) {
        SLB_CUDA(cudaMemcpy2DAsync(dp, dpitch, hp, hpitch, width, (size_t)nc, cudaMemcpyHostToDevice, r.s_h2d));
        SLB_CUDA(cudaEventRecord(t->ev, r.s_h2d));
        t->state.store(1);
        return id;
    }
    cudaStream_t s = r.s_h2d;
    io_thread(0).submit([=](IoThread &io) {
        // chunks of whole columns (a column longer than a slot is cut along its length)
        const size_t sb = io.slot_bytes();
        const size_t rows_per = width <= sb ? width : sb / es * es;
        int k = 0;
        for (size_t ro = 0; ro < width; ro += rows_per) {
            const size_t w = std::min(rows_per, width - ro);
            const int64_t cpc = std::max<int64_t>(1, (int64_t)(sb / w));
            for (int64_t cc = 0; cc < nc; cc += cpc, ++k) {
                const int64_t n = std::min(cpc, nc - cc);
                Slot &sl = io.slot(k);
                SLB_CUDA(cudaEventSynchronize(sl.ev));                       // the DMA that last read this slot is done
                io.pool().copy2d(sl.buf, w, hp + ro + (size_t)cc * hpitch, hpitch, w, n);
                SLB_CUDA(cudaMemcpy2DAsync(dp + ro + (size_t)cc * dpitch, dpitch, sl.buf, w, w, (size_t)n, cudaMemcpyHostToDevice, s));
                SLB_CUDA(cudaEventRecord(sl.ev, s));
            }
        }
        SLB_CUDA(cudaEventRecord(t->ev, s));
        t->state.store(1, std::memory_order_release);
    });
    return id;
}

int HostLink::download(int64_t r0, int64_t r1, int64_t c0, int64_t c1, cudaEvent_t after)
{
    Runtime &r = rt();
    im_->tickets.emplace_back();
    Ticket *t = &im_->tickets.back();
    t->download = true;
    const int id = (int)im_->tickets.size() - 1;
    SLB_CUDA(cudaEventCreateWithFlags(&t->ev, cudaEventDisableTiming));
    const int64_t nr = r1 - r0, nc = c1 - c0;
    cudaStream_t s = r.s_d2h;
    if (nr <= 0 || nc <= 0) { SLB_CUDA(cudaEventRecord(t->ev, s)); t->state.store(2); return id; }
    const size_t es = h_.elem;
    char *hp = (char *)h_.p + (size_t)(r0 + c0 * h_.ld) * es;
    const char *dp = (const char *)dev_ + (size_t)(r0 + c0 * ldd_) * es;
    const size_t hpitch = (size_t)h_.ld * es, dpitch = (size_t)ldd_ * es, width = (size_t)nr * es;
    down_bytes_ += (int64_t)(width * (size_t)nc);
    counter_add("d2h_bytes", (int64_t)(width * (size_t)nc));
    if (h_.pinned) {
        if (after) SLB_CUDA(cudaStreamWaitEvent(s, after, 0));
        SLB_CUDA(cudaMemcpy2DAsync(hp, hpitch, dp, dpitch, width, (size_t)nc, cudaMemcpyDeviceToHost, s));
        SLB_CUDA(cudaEventRecord(t->ev, s));
        t->state.store(2);
        return id;
    }
    io_thread(1).submit([=](IoThread &io) {
        if (after) SLB_CUDA(cudaStreamWaitEvent(s, after, 0));
        const size_t sb = io.slot_bytes();
        const size_t rows_per = width <= sb ? width : sb / es * es;
        struct Chunk { size_t ro, w; int64_t cc, n; };
        std::vector<Chunk> ch;
        for (size_t ro = 0; ro < width; ro += rows_per) {
            const size_t w = std::min(rows_per, width - ro);
            const int64_t cpc = std::max<int64_t>(1, (int64_t)(sb / w));
            for (int64_t cc = 0; cc < nc; cc += cpc) ch.push_back({ ro, w, cc, std::min(cpc, nc - cc) });
        }
        auto issue = [&](size_t k) {
            Slot &sl = io.slot((int)k);
            SLB_CUDA(cudaMemcpy2DAsync(sl.buf, ch[k].w, dp + ch[k].ro + (size_t)ch[k].cc * dpitch, dpitch, ch[k].w, (size_t)ch[k].n,
                                       cudaMemcpyDeviceToHost, s));
            SLB_CUDA(cudaEventRecord(sl.ev, s));
        };
        // the copy engine fills slot k+1 .. k+NSLOT-1 while the workers empty slot k
        size_t issued = 0;
        for (size_t k = 0; k < ch.size(); ++k) {
            while (issued < ch.size() && issued < k + NSLOT - 1) issue(issued++);
            Slot &sl = io.slot((int)k);
            SLB_CUDA(cudaEventSynchronize(sl.ev));
            io.pool().copy2d(hp + ch[k].ro + (size_t)ch[k].cc * hpitch, hpitch, sl.buf, ch[k].w, ch[k].w, ch[k].n);
        }
        SLB_CUDA(cudaEventRecord(t->ev, s));
        t->state.store(2, std::memory_order_release);
    });
    return id;
}

bool HostLink::done(int ticket)
{
    Ticket &t = im_->tickets[(size_t)ticket];
    const int st = t.state.load(std::memory_order_acquire);
    if (st == 0) return false;
    if (t.download && !h_.pinned) return st == 2;
    cudaError_t e = cudaEventQuery(t.ev);
    if (e == cudaSuccess) return true;
    if (e != cudaErrorNotReady) SLB_CUDA(e);
    cudaGetLastError();
    return false;
}

void HostLink::wait(int ticket)
{
    Ticket &t = im_->tickets[(size_t)ticket];
    const int need = (t.download && !h_.pinned) ? 2 : 1;
    while (t.state.load(std::memory_order_acquire) < need) std::this_thread::sleep_for(std::chrono::microseconds(50));
    SLB_CUDA(cudaEventSynchronize(t.ev));
}

void HostLink::stream_wait(int ticket, cudaStream_t s)
{
    Ticket &t = im_->tickets[(size_t)ticket];
    while (t.state.load(std::memory_order_acquire) < 1) std::this_thread::sleep_for(std::chrono::microseconds(20));
    SLB_CUDA(cudaStreamWaitEvent(s, t.ev, 0));
}

void HostLink::finish()
{
    for (size_t i = 0; i < im_->tickets.size(); ++i) wait((int)i);
}

}  // namespace slb

extern "C" void slb200_test_copy2d(char *dst, size_t dpitch, const char *src, size_t spitch, size_t width, int64_t n, int threads)
{ slb::test_copy2d(dst, dpitch, src, spitch, width, n, threads); }

