// tests/libm_pin.cpp -- TEST INFRASTRUCTURE: pins pt_expf / pt_logf (pbrt-v3-distributed_b200/csrc/pt_explog.cuh, compiled
// here as plain C++) against the host libm's std::exp / std::log -- what the reference calls -- for EVERY float bit pattern.
//   g++ -O2 -std=c++17 -ffp-contract=off -pthread tests/libm_pin.cpp -o /tmp/libm_pin && /tmp/libm_pin
#include <atomic>
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

#include "../pbrt-v3-distributed_b200/csrc/pt_explog.cuh"

using namespace b200pt;

int main(int argc, char **argv) {
    const unsigned nThreads = std::max(1u, std::thread::hardware_concurrency());
    const uint64_t stride = argc > 1 ? (uint64_t)atoll(argv[1]) : 1;  // 1 = exhaustive
    std::atomic<uint64_t> badExp{0}, badLog{0}, firstExp{~0ull}, firstLog{~0ull};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nThreads; ++t)
        pool.emplace_back([&, t]() {
            uint64_t be = 0, bl = 0;
            for (uint64_t u = t * stride; u < (1ull << 32); u += nThreads * stride) {
                const float x = uint_as_float((uint32_t)u);
                const float e0 = std::exp(x), e1 = pt_expf(x);
                const float l0 = std::log(x), l1 = pt_logf(x);
                const bool nanE = std::isnan(e0) && std::isnan(e1), nanL = std::isnan(l0) && std::isnan(l1);
                if (!nanE && float_as_uint(e0) != float_as_uint(e1)) {
                    ++be;
                    uint64_t cur = firstExp.load();
                    while (u < cur && !firstExp.compare_exchange_weak(cur, u)) {
                    }
                }
                if (!nanL && float_as_uint(l0) != float_as_uint(l1)) {
                    ++bl;
                    uint64_t cur = firstLog.load();
                    while (u < cur && !firstLog.compare_exchange_weak(cur, u)) {
                    }
                }
            }
            badExp += be;
            badLog += bl;
        });
    for (auto &th : pool) th.join();
    printf("expf: %llu mismatches", (unsigned long long)badExp.load());
    if (badExp) printf(" (first at bits 0x%08llx)", (unsigned long long)firstExp.load());
    printf("\nlogf: %llu mismatches", (unsigned long long)badLog.load());
    if (badLog) printf(" (first at bits 0x%08llx)", (unsigned long long)firstLog.load());
    printf("\n%s over %s float bit patterns\n", (badExp || badLog) ? "FAIL" : "OK", stride == 1 ? "all 2^32" : "a strided subset of the");
    return (badExp || badLog) ? 1 : 0;
}
