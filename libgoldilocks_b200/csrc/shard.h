// shard.h -- one host-pointer batch spread over several GPUs (north_star: "a batch shards trivially
// across the 8 GPUs of one box with plain per-device streams and no NCCL").
//
// Every element of a batch is independent, so a `*_batch` call over [0, n) is cut into contiguous ranges, one run of
// ranges per device of the configured set.  A range is executed by re-entering the same C-ABI function on a worker
// thread that is bound to its device (and to one of LANES independent per-device contexts: stream + arena), so the
// single-device code path is the only code path.  No collective, no peer traffic; the host thread that made the call
// blocks until every range is back.
//
// Two shapes of plan:
//   * heavy operations (verification, signing, ladders, scalar multiplications): one range per device -- a second cut
//     would pay the per-call passes (key grouping, table waves) twice;
//   * light, PCIe-bound operations (field, point and codec entry points): ranges of ~CHUNK_BYTES of traffic, dealt
//     round-robin to the LANES contexts of the range's device, so the copy-in of one chunk, the kernel of the next and
//     the copy-out of a third overlap on the device's two copy engines and its SMs.
#pragma once
#include <cuda_runtime.h>
#include <sched.h>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace shard {

constexpr int LANES = 3;                       /* contexts per device: copy-in | kernel | copy-out of three chunks overlap */
constexpr size_t CHUNK_BYTES = 24u << 20;      /* traffic per pipelined chunk */
constexpr size_t MIN_CHUNK = 4096;             /* elements; below this a cut costs more than it hides */

struct Piece { size_t lo, hi; int slot; };     /* slot = device index within the set x LANES + lane */

// The plan is pure arithmetic (tested on the CPU tier through goldilocks_b200_shard_plan).
// heavy: min(ndev, n / min_per_dev) equal ranges on lane 0.  light: the same device ranges, each cut into chunks.
inline std::vector<Piece> plan(size_t n, int ndev, size_t min_per_dev, size_t bytes_per_elem, bool pipelined) {
    std::vector<Piece> out;
    if (n == 0 || ndev < 1) return out;
    if (min_per_dev < 1) min_per_dev = 1;
    size_t parts = n / min_per_dev;
    if (parts > (size_t)ndev) parts = (size_t)ndev;
    if (parts < 1) parts = 1;
    for (size_t d = 0; d < parts; d++) {
        const size_t lo = n / parts * d + (d < n % parts ? d : n % parts);
        const size_t hi = lo + n / parts + (d < n % parts ? 1 : 0);
        if (!pipelined) { out.push_back({lo, hi, (int)d * LANES}); continue; }
        size_t chunk = bytes_per_elem ? CHUNK_BYTES / bytes_per_elem : (hi - lo);
        if (chunk < MIN_CHUNK) chunk = MIN_CHUNK;
        size_t k = (hi - lo + chunk - 1) / chunk;              /* chunks of this device, equalised */
        if (k < 1) k = 1;
        for (size_t c = 0; c < k; c++) {
            const size_t a = lo + (hi - lo) / k * c + (c < (hi - lo) % k ? c : (hi - lo) % k);
            const size_t b = a + (hi - lo) / k + (c < (hi - lo) % k ? 1 : 0);
            out.push_back({a, b, (int)d * LANES + (int)(c % LANES)});
        }
    }
    return out;
}

struct Job {
    std::function<int()> fn;     /* returns GOLDILOCKS_SUCCESS (-1) or FAILURE (0) */
    int result = 0;
    std::string err;
    struct Group *group = nullptr;
};
struct Group {
    std::mutex mu;
    std::condition_variable cv;
    size_t remaining = 0;
};

extern thread_local bool t_worker;   /* set on worker threads: their calls run on their own device, never re-shard */
extern thread_local int t_lane;      /* which of the device's LANES contexts this thread uses (0 on caller threads) */
const std::string &worker_error();   /* the calling thread's last error text (abi.cu's g_err) */

class Worker {
public:
    Worker(int dev, int lane) : dev_(dev), lane_(lane) {
        th_ = std::thread([this] { loop(); });
        th_.detach();            /* workers live until the process exits; they idle on their queue */
    }
    void submit(Job *j) {
        { std::lock_guard<std::mutex> g(mu_); q_.push_back(j); }
        cv_.notify_one();
    }
    int dev() const { return dev_; }
private:
    void loop() {
        t_worker = true;
        t_lane = lane_;
        cudaSetDevice(dev_);
        for (;;) {
            Job *j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !q_.empty(); });
                j = q_.front(); q_.pop_front();
            }
            j->result = j->fn();
            if (j->result != -1) j->err = worker_error();
            Group *g = j->group;
            { std::lock_guard<std::mutex> gl(g->mu); g->remaining--; if (g->remaining == 0) g->cv.notify_all(); }   /* notify under the lock: the group lives on the caller's stack */
        }
    }
    int dev_, lane_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Job *> q_;
    std::thread th_;
};

struct Pool {
    std::vector<int> devs;
    std::vector<Worker *> workers;   /* devs.size() x LANES, created once and kept (never destroyed: detached threads) */
};
std::shared_ptr<Pool> current();     /* null = no device set configured: every call runs on the caller's current device */

// Run fn(lo, count) over the given pieces on the pool's workers; SUCCESS iff every piece succeeded.
template <class Fn>
int run_pieces(const std::shared_ptr<Pool> &pool, const std::vector<Piece> &pieces, std::string *err, Fn fn) {
    std::vector<Job> jobs(pieces.size());
    Group g;
    g.remaining = pieces.size();
    for (size_t i = 0; i < pieces.size(); i++) {
        const Piece p = pieces[i];
        jobs[i].group = &g;
        jobs[i].fn = [fn, p] { return (int)fn(p.lo, p.hi - p.lo); };
    }
    for (size_t i = 0; i < pieces.size(); i++) pool->workers[(size_t)pieces[i].slot]->submit(&jobs[i]);
    { std::unique_lock<std::mutex> lk(g.mu); g.cv.wait(lk, [&] { return g.remaining == 0; }); }
    int r = -1;
    for (auto &j : jobs)
        if (j.result != -1) { if (r == -1 && err) *err = j.err; r = 0; }
    return r;
}
template <class Fn>
int run(const std::shared_ptr<Pool> &pool, size_t n, size_t min_per_dev, size_t bytes_per_elem, bool pipelined, std::string *err, Fn fn) {
    return run_pieces(pool, plan(n, (int)pool->devs.size(), min_per_dev, bytes_per_elem, pipelined), err, fn);
}
// fn(slot, 1) once on every worker (every lane of every device of the set)
template <class Fn>
int run_everywhere(const std::shared_ptr<Pool> &pool, std::string *err, Fn fn) {
    std::vector<Piece> pieces;
    for (size_t s = 0; s < pool->workers.size(); s++) pieces.push_back({s, s + 1, (int)s});
    return run_pieces(pool, pieces, err, fn);
}

}  // namespace shard
