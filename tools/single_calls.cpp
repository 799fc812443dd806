// tools/single_calls.cpp -- the reference's own one-signature entry point (goldilocks_ed448_verify, ed448.h:157-165) called from T native
// threads, K calls each, against libgoldilocks_b200.so: batches of one (window 0) versus gathered calls (goldilocks_b200_coalesce).
// The corpus is signed by the library's own batch signer (keys and messages from a counter-mode SHAKE-free generator); every fourth
// signature is corrupted and every status is checked against the batch call's answer.  Prints one JSON object.
//   g++ -O2 -std=c++17 -pthread -Iinclude -o tools/single_calls tools/single_calls.cpp -Llibgoldilocks_b200 -lgoldilocks_b200 -Wl,-rpath,'$ORIGIN/../libgoldilocks_b200'
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <goldilocks_b200.h>

static uint64_t g_state = 0x9e3779b97f4a7c15ull;
static uint8_t next_byte() { g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17; return (uint8_t)(g_state >> 32); }

struct Result { double per_s; unsigned long long calls, batches, largest; int wrong; };

static Result run(int T, int K, unsigned window_us, const std::vector<uint8_t> &sig, const std::vector<uint8_t> &pk, const std::vector<uint8_t> &msg,
                  const std::vector<size_t> &off, const std::vector<goldilocks_error_t> &want) {
    unsigned long long c0, b0, big;
    goldilocks_b200_coalesce(window_us, 0);
    goldilocks_b200_coalesce_stats(&c0, &b0, &big);
    std::atomic<int> wrong{0};
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            for (int k = 0; k < K; k++) {
                const size_t i = (size_t)t * K + k;
                const goldilocks_error_t st = goldilocks_ed448_verify(&sig[114 * i], &pk[57 * i], &msg[off[i]], off[i + 1] - off[i], 0, NULL, 0);
                if (st != want[i]) wrong++;
            }
        });
    for (auto &x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    unsigned long long c1, b1;
    goldilocks_b200_coalesce_stats(&c1, &b1, &big);
    goldilocks_b200_coalesce(0, 0);
    return {T * (double)K / dt, c1 - c0, b1 - b0, big, wrong.load()};
}

int main(int argc, char **argv) {
    const int maxT = argc > 1 ? atoi(argv[1]) : 1024, K = argc > 2 ? atoi(argv[2]) : 16;
    const unsigned window = argc > 3 ? (unsigned)atoi(argv[3]) : 250;
    const size_t n = (size_t)maxT * K;
    std::vector<uint8_t> sk(57 * n), pk(57 * n), sig(114 * n), msg;
    std::vector<size_t> off(n + 1, 0);
    for (auto &b : sk) b = next_byte();
    for (size_t i = 0; i < n; i++) { const size_t len = 16 + next_byte() % 48; for (size_t j = 0; j < len; j++) msg.push_back(next_byte()); off[i + 1] = msg.size(); }
    if (goldilocks_ed448_derive_public_key_batch(pk.data(), sk.data(), n) != GOLDILOCKS_SUCCESS) { fprintf(stderr, "derive: %s\n", goldilocks_b200_last_error()); return 1; }
    if (goldilocks_ed448_sign_batch(sig.data(), sk.data(), pk.data(), msg.data(), off.data(), 0, NULL, 0, n) != GOLDILOCKS_SUCCESS) { fprintf(stderr, "sign: %s\n", goldilocks_b200_last_error()); return 1; }
    for (size_t i = 0; i < n; i += 4) sig[114 * i + 60] ^= 1;
    std::vector<goldilocks_error_t> want(n);
    if (goldilocks_ed448_verify_batch(want.data(), sig.data(), pk.data(), msg.data(), off.data(), 0, NULL, 0, n) != GOLDILOCKS_SUCCESS) { fprintf(stderr, "verify: %s\n", goldilocks_b200_last_error()); return 1; }
    size_t good = 0;
    for (auto s : want) good += s == GOLDILOCKS_SUCCESS;
    if (good != n - (n + 3) / 4) { fprintf(stderr, "unexpected accept count %zu of %zu\n", good, n); return 1; }
    run(4, 2, 0, sig, pk, msg, off, want);   /* warm-up of the one-element shapes */
    printf("{\"how\": \"tools/single_calls.cpp: goldilocks_ed448_verify from native threads, %d calls each, statuses checked against the batch call\", \"runs\": [", K);
    bool first = true;
    int bad = 0;
    auto emit = [&](const char *mode, int T, unsigned w) {
        const Result r = run(T, K, w, sig, pk, msg, off, want);
        bad += r.wrong;
        printf("%s{\"mode\": \"%s\", \"threads\": %d, \"window_us\": %u, \"verifies_per_s\": %.1f, \"batches\": %llu, \"largest_batch\": %llu, \"wrong\": %d}", first ? "" : ", ", mode, T, w, r.per_s,
               r.batches, w ? r.largest : 1ull, r.wrong);
        first = false;
    };
    emit("alone", 1, 0);
    emit("alone", 8, 0);
    for (int T = 16; T <= maxT; T *= 4) emit("gathered", T, window);
    if (maxT != 16 && maxT != 64 && maxT != 256 && maxT != 1024) emit("gathered", maxT, window);
    printf("]}\n");
    return bad ? 1 : 0;
}
