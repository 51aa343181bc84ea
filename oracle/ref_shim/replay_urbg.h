/*
 * oracle/ref_shim/replay_urbg.h -- TEST INFRASTRUCTURE (see oracle/tm_oracle.h).
 *
 * Force-included (-include) when oracle/Makefile compiles the UNMODIFIED reference sources.
 * It substitutes a replay engine for std::mt19937_64 so that the reference's own samplers
 * (std::poisson_distribution, std::gamma_distribution, std::normal_distribution) consume a
 * host-supplied sequence of uniforms instead of Mersenne-Twister output.  With
 * word = (uint64)(u * 2^53) << 11, std::generate_canonical<double,53> returns exactly u.
 *
 * When no buffer is armed (tm_replay_arm(nullptr, 0)) the engine falls back to a real
 * std::mt19937_64, which is what the CPU-baseline timing uses.
 */
#ifndef TM_REPLAY_URBG_H
#define TM_REPLAY_URBG_H
#include <random>
#include <cstdint>

struct tm_replay_state {
    const double *u = nullptr;
    uint64_t n = 0, pos = 0, overflow = 0;
};
inline thread_local tm_replay_state tm_replay;

inline void tm_replay_arm(const double *u, uint64_t n) { tm_replay = tm_replay_state{u, n, 0, 0}; }

namespace std {
struct tm_replay_urbg {
    using result_type = uint64_t;
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return ~result_type(0); }
    std::mersenne_twister_engine<uint_fast64_t, 64, 312, 156, 31, 0xb5026f5aa96619e9ULL, 29,
        0x5555555555555555ULL, 17, 0x71d67fffeda60000ULL, 37, 0xfff7eee000000000ULL, 43,
        6364136223846793005ULL> real;
    template <class T> void seed(T s) { real.seed(s); }
    result_type operator()()
    {
        if (!tm_replay.u) return real();
        double u = 0.75;
        if (tm_replay.pos < tm_replay.n) u = tm_replay.u[tm_replay.pos++];
        else tm_replay.overflow++;
        return (uint64_t)(u * 9007199254740992.0) << 11;
    }
};
}
#define mt19937_64 tm_replay_urbg
#endif
