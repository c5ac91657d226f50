// Host-side, bit-exact, multi-threaded re-implementation of numpy's LEGACY normal generator
// (np.random.normal / RandomState.normal: MT19937 + polar Box-Muller "legacy_gauss").
//
// Why it exists: the reference's FastDVDnet fine-tune perturbs its input with noise drawn on the host from
// the global numpy RNG (utils/utils_image.py:183-192 called at packages/fastdvdnet/test_fastdvdnet.py:359).
// Reproducing the reference bit for bit means consuming that exact stream; numpy's generator is single
// threaded (~25 ns/sample -> 0.16 s for one 8x3x512x512 draw, i.e. more than the GPU time of the ADMM
// iterations between two fine-tune calls).  The Mersenne-Twister word stream is inherently serial but cheap
// (~2 ns/word); everything after it (uniform doubles, the rejection test, log/sqrt) is embarrassingly
// parallel once acceptance flags are prefix-summed.  Same libm, same operation order => identical doubles
// and identical final generator state (verified against numpy in tests/test_host_rng.py).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "sci_b200.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfU, UPPER = 0x80000000U, LOWER = 0x7fffffffU;

inline uint32_t twist(uint32_t a, uint32_t b) {
    const uint32_t y = (a & UPPER) | (b & LOWER);
    return (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 1U)) & MATRIX_A);
}

// next state block from the previous one (same recurrence as numpy's mt19937_gen)
void next_block(const uint32_t* old, uint32_t* neu) {
    int k = 0;
    for (; k < MT_N - MT_M; ++k) neu[k] = old[k + MT_M] ^ twist(old[k], old[k + 1]);
    for (; k < MT_N - 1; ++k) neu[k] = neu[k - (MT_N - MT_M)] ^ twist(old[k], old[k + 1]);
    neu[MT_N - 1] = neu[MT_M - 1] ^ twist(old[MT_N - 1], neu[0]);
}

inline uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680U;
    y ^= (y << 15) & 0xefc60000U;
    y ^= (y >> 18);
    return y;
}

inline double word_pair_to_double(uint32_t w0, uint32_t w1) {      // mt19937_next_double
    const int32_t a = (int32_t)(temper(w0) >> 5), b = (int32_t)(temper(w1) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

}  // namespace

// key[624], *pos, *has_gauss, *gauss: numpy legacy state (RandomState.get_state()), updated in place.
// out[n] = loc + scale * legacy_gauss(), exactly as RandomState.normal(loc, scale, n).
extern "C" int sci_host_legacy_normal(uint32_t* key, int* pos, int* has_gauss, double* gauss, double loc, double scale,
                                      double* out, long n, int nthreads) {
    if (!key || !pos || !has_gauss || !gauss || (!out && n > 0) || n < 0 || *pos < 0 || *pos > MT_N) return SCI_EINVAL;
    if (n == 0) return SCI_OK;
    long start = 0;
    if (*has_gauss) {
        out[0] = loc + scale * (*gauss);
        *has_gauss = 0;
        *gauss = 0.0;
        start = 1;
        if (n == 1) return SCI_OK;
    }
    const long m = n - start;                 // normals still to produce
    const long pairs = (m + 1) / 2;           // accepted candidate pairs needed
    if (nthreads < 1) nthreads = 1;
    const int T = (int)std::min<long>(nthreads, std::max<long>(1, pairs / 4096));

    // word stream as a growing list of UNTEMPERED state blocks; block 0 is the current key, consumption starts at *pos
    std::vector<uint32_t> blocks(key, key + MT_N);
    const long pos0 = *pos;
    auto ensure_words = [&](long words_from_pos0) {
        const long need_blocks = (pos0 + words_from_pos0 + MT_N - 1) / MT_N;
        long have = (long)(blocks.size() / MT_N);
        if (need_blocks > have) {
            blocks.resize((size_t)need_blocks * MT_N);
            for (; have < need_blocks; ++have) next_block(&blocks[(size_t)(have - 1) * MT_N], &blocks[(size_t)have * MT_N]);
        }
    };

    long cand_done = 0, acc_done = 0;         // candidates examined / accepted so far
    long last_cand = -1;                      // index of the candidate that supplied the final pair
    double last_cached = 0.0;
    while (acc_done < pairs) {
        const long want = pairs - acc_done;
        const long batch = (long)(want * 1.2732395447351628 * 1.02) + 64;      // 4/pi acceptance, small slack
        ensure_words(4 * (cand_done + batch));
        const uint32_t* W = blocks.data() + pos0;
        std::vector<long> counts(T, 0);
        const long per = (batch + T - 1) / T;
        auto accept = [&](long c, double& x1, double& x2, double& r2) {
            const uint32_t* w = W + 4 * c;
            x1 = 2.0 * word_pair_to_double(w[0], w[1]) - 1.0;
            x2 = 2.0 * word_pair_to_double(w[2], w[3]) - 1.0;
            r2 = x1 * x1 + x2 * x2;
            return !(r2 >= 1.0 || r2 == 0.0);
        };
        {   // pass 1: acceptance counts per thread
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    const long c0 = cand_done + t * per, c1 = std::min(cand_done + batch, c0 + per);
                    long cnt = 0;
                    double x1, x2, r2;
                    for (long c = c0; c < c1; ++c) cnt += accept(c, x1, x2, r2);
                    counts[t] = cnt;
                });
            for (auto& t : th) t.join();
        }
        std::vector<long> offs(T + 1, 0);
        for (int t = 0; t < T; ++t) offs[t + 1] = offs[t] + counts[t];
        {   // pass 2: write the accepted pairs at their prefix-summed positions
            std::vector<long> lastc(T, -1);
            std::vector<double> lastx(T, 0.0);
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    const long c0 = cand_done + t * per, c1 = std::min(cand_done + batch, c0 + per);
                    long a = acc_done + offs[t];
                    double x1, x2, r2;
                    for (long c = c0; c < c1 && a < pairs; ++c) {
                        if (!accept(c, x1, x2, r2)) continue;
                        const double f = sqrt(-2.0 * log(r2) / r2);
                        const long o = start + 2 * a;
                        out[o] = loc + scale * (f * x2);                         // first call returns f*x2 ...
                        if (o + 1 < n) out[o + 1] = loc + scale * (f * x1);      // ... and caches f*x1 for the next
                        if (a == pairs - 1) { lastc[t] = c; lastx[t] = f * x1; }
                        ++a;
                    }
                });
            for (auto& t : th) t.join();
            for (int t = 0; t < T; ++t)
                if (lastc[t] >= 0) { last_cand = lastc[t]; last_cached = lastx[t]; }
        }
        acc_done = std::min(pairs, acc_done + offs[T]);
        cand_done += batch;
    }
    // generator state right after the words of candidate `last_cand`
    const long g = pos0 + 4 * (last_cand + 1);
    long blk = g / MT_N, p = g % MT_N;
    if (p == 0 && blk > 0) { blk -= 1; p = MT_N; }         // numpy leaves pos == 624 until the next draw
    memcpy(key, &blocks[(size_t)blk * MT_N], MT_N * sizeof(uint32_t));
    *pos = (int)p;
    if (m & 1) { *has_gauss = 1; *gauss = last_cached; }    // odd count: the second value of the last pair stays cached
    return SCI_OK;
}
