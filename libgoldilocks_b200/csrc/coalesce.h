// coalesce.h -- single-element calls from many host threads, gathered into one batch.
//
// The reference's own entry points (goldilocks_ed448_verify, goldilocks_ed448_sign, goldilocks_x448; ed448.h:157-165,
// 108-118, point_448.h x448) take ONE element.  On the GPU such a call is a batch of one: 1-3 ms of single-lane latency
// for a throughput of a few hundred per second and thread.  A server that verifies from a pool of threads calls them
// concurrently, though, and the elements are independent -- so the library can put concurrent calls into one `*_batch`
// launch without the caller changing a line: the first thread to arrive becomes the leader of a gathering, waits up to
// `window_us` (or until `max_batch` requests are in), runs the batch for everybody on its own thread and hands the
// results back; threads that arrive meanwhile start the next gathering, so gatherings overlap with running batches.
// At most `max_inflight` batches run at a time: a batch takes about as long whether it holds 10 or 10 000 elements (one lane's
// latency), so when the device is busy the next gathering keeps collecting until a running batch returns instead of adding a small
// batch to the queue -- the batch size adapts to the load, the window only bounds the wait on an idle device.
//
// Off by default (window 0): a lone caller would only pay the window as extra latency.  Host-only code, no CUDA types:
// the gate is exercised on the CPU tier (tests/coalesce/harness.cpp) with a stand-in for the batch call.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <vector>

namespace coalesce {

struct Settings {
    std::atomic<unsigned> window_us{0};      /* 0 = every call runs alone (the default) */
    std::atomic<unsigned> max_batch{4096};   /* a gathering this large stops waiting for its window */
    std::atomic<unsigned> max_inflight{3};   /* batches running at a time (one per context of a device, shard.h LANES) */
};
struct Stats {
    std::atomic<unsigned long long> calls{0}, batches{0}, largest{0};
};

// Req: any struct with a `bool done` member (set by the gate) that carries the call's arguments and result slots.
// run(batch, n) executes all n requests and fills their results; it is called on the leader's thread, outside the lock.
template <class Req>
class Gate {
public:
    template <class Run>
    void submit(Req *r, const Settings &cfg, Stats &st, Run run) {
        std::unique_lock<std::mutex> lk(mu_);
        r->done = false;
        pending_.push_back(r);
        st.calls.fetch_add(1, std::memory_order_relaxed);
        if (gathering_) {                                   /* somebody is collecting: join and sleep until served */
            if (pending_.size() >= cfg.max_batch.load()) cv_leader_.notify_one();
            cv_done_.wait(lk, [r] { return r->done; });
            return;
        }
        gathering_ = true;                                  /* leader of this gathering */
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(cfg.window_us.load());
        for (;;) {
            const bool ripe = pending_.size() >= cfg.max_batch.load() || std::chrono::steady_clock::now() >= deadline;
            if (ripe && inflight_ < cfg.max_inflight.load()) break;
            if (ripe) cv_leader_.wait(lk);                  /* the device is busy: keep collecting until a batch returns */
            else cv_leader_.wait_until(lk, deadline);
        }
        std::vector<Req *> batch;
        batch.swap(pending_);
        gathering_ = false;                                 /* the next arrival leads the next gathering while this batch runs */
        inflight_++;
        lk.unlock();
        st.batches.fetch_add(1, std::memory_order_relaxed);
        unsigned long long big = st.largest.load(std::memory_order_relaxed);
        while (batch.size() > big && !st.largest.compare_exchange_weak(big, batch.size())) {}
        run(batch.data(), batch.size());
        lk.lock();
        inflight_--;
        for (Req *q : batch) q->done = true;                /* under the lock: the requests live on their callers' stacks */
        cv_done_.notify_all();
        cv_leader_.notify_one();                            /* a leader that waits for a free slot */
    }
private:
    std::mutex mu_;
    std::condition_variable cv_leader_, cv_done_;
    std::vector<Req *> pending_;
    bool gathering_ = false;
    unsigned inflight_ = 0;
};

}  // namespace coalesce
