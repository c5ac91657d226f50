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
#include <stdlib.h>
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
//
// Streaming form (round 2): the word stream is produced CHUNK by chunk (SCI_RNG_CHUNK_BLOCKS state blocks, default 6400 =
// 16 MB) by one generator thread while the worker threads turn the previous chunk into normals (pass 1: acceptance counts per
// thread, prefix sum, pass 2: write at the prefix-summed positions).  The first version materialised the whole stream first:
// 3 GB for one 2048x2048x24 colour draw, value-initialised and page-faulted by the serial thread - 1.3 s of its 1.6 s per 1e8
// normals.  A chunk buffer holds the blocks [k*NB - 1, (k+1)*NB]: one block before (numpy leaves pos == 624 at a block end)
// and one after (a candidate's four words may straddle the chunk end).
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

    long NB = 6400;
    if (const char* e = getenv("SCI_RNG_CHUNK_BLOCKS")) NB = std::max(2L, atol(e));
    // small draws: no more blocks per chunk than the draw can need (4/pi candidates per pair, 4 words each, + slack)
    NB = std::min(NB, std::max(2L, (long)(pairs * 1.2732395447351628 * 1.05 * 4) / MT_N + 3));
    const size_t buf_words = (size_t)(NB + 2) * MT_N;
    std::vector<uint32_t> buf[2];
    buf[0].resize(buf_words);
    buf[1].resize(buf_words);
    const long pos0 = *pos;                   // consumption starts at word pos0 of block 0 (= the current key)

    // chunk k's buffer: slot 0 = block k*NB - 1, slots 1..NB = blocks k*NB .. (k+1)*NB - 1, slot NB + 1 = block (k+1)*NB
    auto fill_first = [&](std::vector<uint32_t>& b) {
        memset(b.data(), 0, MT_N * sizeof(uint32_t));
        memcpy(b.data() + MT_N, key, MT_N * sizeof(uint32_t));
        for (long j = 2; j <= NB + 1; ++j) next_block(b.data() + (size_t)(j - 1) * MT_N, b.data() + (size_t)j * MT_N);
    };
    auto fill_next = [&](const std::vector<uint32_t>& prev, std::vector<uint32_t>& b) {
        memcpy(b.data(), prev.data() + (size_t)NB * MT_N, 2 * MT_N * sizeof(uint32_t));      // blocks (k+1)*NB - 1 and (k+1)*NB
        for (long j = 2; j <= NB + 1; ++j) next_block(b.data() + (size_t)(j - 1) * MT_N, b.data() + (size_t)j * MT_N);
    };

    long acc_done = 0;                        // accepted pairs so far
    long last_cand = -1;                      // candidate that supplied the final pair
    double last_cached = 0.0;
    int last_buf = 0;
    long last_chunk = 0;
    fill_first(buf[0]);
    for (long k = 0; acc_done < pairs; ++k) {
        const int cur = (int)(k & 1);
        std::thread gen([&, k, cur] { fill_next(buf[cur], buf[cur ^ 1]); });        // chunk k + 1 while chunk k is consumed
        // candidates whose FIRST word lies in this chunk's own blocks: word index pos0 + 4c in [k*NB*624, (k+1)*NB*624)
        const long w_lo = k * NB * MT_N, w_hi = (k + 1) * NB * MT_N;
        const long c_lo = w_lo <= pos0 ? 0 : (w_lo - pos0 + 3) / 4;
        const long c_hi = (w_hi - pos0 + 3) / 4;                                     // exclusive
        const long ncand = std::max(0L, c_hi - c_lo);
        // word w of the stream sits at buf[(w - w_lo) + MT_N]
        const uint32_t* B0 = buf[cur].data();
        const long base = MT_N - w_lo + pos0;                                        // B0[base + 4c + j] = word j of candidate c
        const int T = (int)std::min<long>(nthreads, std::max<long>(1, ncand / 4096));
        const long per = (ncand + T - 1) / T;
        auto accept = [&](long c, double& x1, double& x2, double& r2) {
            const uint32_t* w = B0 + (base + 4 * c);
            x1 = 2.0 * word_pair_to_double(w[0], w[1]) - 1.0;
            x2 = 2.0 * word_pair_to_double(w[2], w[3]) - 1.0;
            r2 = x1 * x1 + x2 * x2;
            return !(r2 >= 1.0 || r2 == 0.0);
        };
        std::vector<long> counts(T, 0);
        auto count_range = [&](int t) {
            const long c0 = c_lo + t * per, c1 = std::min(c_hi, c0 + per);
            long cnt = 0;
            double x1, x2, r2;
            for (long c = c0; c < c1; ++c) cnt += accept(c, x1, x2, r2);
            counts[t] = cnt;
        };
        std::vector<long> lastc(T, -1);
        std::vector<double> lastx(T, 0.0);
        std::vector<long> offs(T + 1, 0);
        auto write_range = [&](int t) {
            const long c0 = c_lo + t * per, c1 = std::min(c_hi, c0 + per);
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
        };
        if (T == 1) {
            count_range(0);
            offs[1] = counts[0];
            write_range(0);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back(count_range, t);
            for (auto& t : th) t.join();
            for (int t = 0; t < T; ++t) offs[t + 1] = offs[t] + counts[t];
            th.clear();
            for (int t = 0; t < T; ++t) th.emplace_back(write_range, t);
            for (auto& t : th) t.join();
        }
        for (int t = 0; t < T; ++t)
            if (lastc[t] >= 0) { last_cand = lastc[t]; last_cached = lastx[t]; last_buf = cur; last_chunk = k; }
        acc_done = std::min(pairs, acc_done + offs[T]);
        gen.join();
    }
    // generator state right after the words of candidate `last_cand`
    const long g = pos0 + 4 * (last_cand + 1);
    long blk = g / MT_N, p = g % MT_N;
    if (p == 0 && blk > 0) { blk -= 1; p = MT_N; }         // numpy leaves pos == 624 until the next draw
    // block blk lives in the buffer of the chunk that held the last candidate, at slot blk - last_chunk*NB + 1 (0 .. NB + 1)
    const long slot = blk - last_chunk * NB + 1;
    if (slot < 0 || slot > NB + 1) return SCI_EINVAL;       // cannot happen: a candidate spans at most 4 words past its chunk
    if (!(slot == 0 && last_chunk == 0))
        memcpy(key, buf[last_buf].data() + (size_t)slot * MT_N, MT_N * sizeof(uint32_t));
    *pos = (int)p;
    if (m & 1) { *has_gauss = 1; *gauss = last_cached; }    // odd count: the second value of the last pair stays cached
    return SCI_OK;
}
