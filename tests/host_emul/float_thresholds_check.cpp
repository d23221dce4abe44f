// Exhaustive host check of manhattanslam_b200/csrc/float_thresholds.h: for every one of the 2^32 float bit patterns x and every
// double literal c the superpixel kernels compare against,
//     ((double)x <  c) == D_LT(x, c)      ((double)x >  c) == D_GT(x, c)
//     ((double)x >= c) == D_GE(x, c)      ((double)x <= c) == D_LE(x, c)
// (NaNs, infinities, zeros and subnormals included).  Prints the number of mismatches; exit code 1 if any.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <thread>
#include <vector>
#include <atomic>
#include "float_thresholds.h"

#define CHECK_ALL(x, c, bad)                                         \
    do {                                                             \
        const double xd_ = (double)(x);                              \
        bad += (xd_ < (c)) != D_LT(x, c);                            \
        bad += (xd_ > (c)) != D_GT(x, c);                            \
        bad += (xd_ >= (c)) != D_GE(x, c);                           \
        bad += (xd_ <= (c)) != D_LE(x, c);                           \
    } while (0)

int main(int argc, char **argv) {
    const unsigned nThreads = argc > 1 ? (unsigned)atoi(argv[1]) : std::max(1u, std::thread::hardware_concurrency());
    const uint64_t stride = argc > 2 ? (uint64_t)atoll(argv[2]) : 1;
    std::vector<uint32_t> centres;
    for (double c : {0.4, 0.01, 0.1, 0.05, 0.2}) {
        const float f = (float)c;
        uint32_t b;
        memcpy(&b, &f, 4);
        centres.push_back(b);
    }
    std::atomic<unsigned long long> total(0);
    std::vector<std::thread> pool;
    const uint64_t N = 1ull << 32, per = (N + nThreads - 1) / nThreads;
    for (unsigned t = 0; t < nThreads; t++)
        pool.emplace_back([&, t] {
            unsigned long long bad = 0;
            const uint64_t lo = t * per, hi = std::min(N, lo + per);
            for (uint64_t u = lo; u < hi; u++) {
                // stride > 1: every stride-th bit pattern, and EVERY pattern within 2^21 ulps of a literal's neighbouring floats
                if (stride > 1 && u % stride != 0) {
                    bool near = false;
                    for (uint32_t c : centres) near |= (uint32_t)(u & 0x7fffffffu) - (c - (1u << 21)) < (1u << 22);
                    if (!near) continue;
                }
                const uint32_t b = (uint32_t)u;
                float x;
                memcpy(&x, &b, 4);
                CHECK_ALL(x, 0.4, bad);    // HUBER_RANGE
                CHECK_ALL(x, -0.4, bad);
                CHECK_ALL(x, 0.01, bad);   // Newton step, valid depth of updatePixels
                CHECK_ALL(x, -0.01, bad);
                CHECK_ALL(x, 0.1, bad);    // valid depth of updateSeeds, MAX_ANGLE_COS
                CHECK_ALL(x, -0.1, bad);
                CHECK_ALL(x, 0.05, bad);   // valid depth of calculateSpDepthNorms
                CHECK_ALL(x, 0.2, bad);    // updateDiff
            }
            total += bad;
        });
    for (auto &th : pool) th.join();
    printf("float_thresholds: %llu mismatches (stride %llu: %s) x 8 literals x 4 predicates\n", total.load(), (unsigned long long)stride,
           stride > 1 ? "every stride-th float + all floats within 2^21 ulps of a literal" : "all 2^32 floats");
    return total.load() ? 1 : 0;
}
