// CPU-tier exercise of csrc/coalesce.h: T threads make K single-element "calls" each through a Gate whose batch runner is a
// stand-in (sleeps like a kernel launch, result = 3 x + 1).  Prints one JSON line; tests/test_coalesce.py checks it.
//   g++ -O2 -std=c++17 -pthread -o harness harness.cpp && ./harness <threads> <calls per thread> <window_us> <max_batch>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include "../../libgoldilocks_b200/csrc/coalesce.h"

struct Req { bool done; unsigned long long x, y; std::thread::id ran_on; };

int main(int argc, char **argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 16, calls = argc > 2 ? atoi(argv[2]) : 50;
    coalesce::Settings cfg;
    coalesce::Stats st;
    cfg.window_us.store(argc > 3 ? (unsigned)atoi(argv[3]) : 200);
    cfg.max_batch.store(argc > 4 ? (unsigned)atoi(argv[4]) : 4096);
    coalesce::Gate<Req> gate;
    std::atomic<unsigned long long> wrong{0}, over{0}, foreign{0};
    std::atomic<int> running{0}, most_running{0};
    const unsigned maxb = cfg.max_batch.load();
    auto run = [&](Req **q, size_t n) {
        if (n > maxb + (size_t)threads) over++;            /* arrivals between the wake-up and the swap may ride along, never more than one per thread */
        const int now = ++running;
        int seen = most_running.load();
        while (now > seen && !most_running.compare_exchange_weak(seen, now)) {}
        std::this_thread::sleep_for(std::chrono::microseconds(300));
        --running;
        for (size_t i = 0; i < n; i++) { q[i]->y = 3 * q[i]->x + 1; q[i]->ran_on = std::this_thread::get_id(); }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            for (int k = 0; k < calls; k++) {
                Req r = {false, (unsigned long long)t * 1000003ull + (unsigned long long)k, 0, {}};
                gate.submit(&r, cfg, st, run);
                if (r.y != 3 * r.x + 1 || !r.done) wrong++;
                if (r.ran_on != std::this_thread::get_id()) foreign++;
            }
        });
    for (auto &th : pool) th.join();
    printf("{\"threads\": %d, \"calls\": %llu, \"batches\": %llu, \"largest\": %llu, \"wrong\": %llu, \"oversized\": %llu, \"served_by_another_thread\": %llu, \"most_batches_at_a_time\": %d}\n",
           threads, st.calls.load(), st.batches.load(), st.largest.load(), wrong.load(), over.load(), foreign.load(), most_running.load());
    return wrong.load() ? 1 : 0;
}
