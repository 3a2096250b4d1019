#include <cosma/auto_strategy.hpp>
#include <cosma/environment_variables.hpp>
#include <cosma/schedule.hpp>

#include <algorithm>
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <vector>

namespace cosma {

long long schedule_footprint_elements(const Strategy& strategy) {
    const int P = static_cast<int>(strategy.P);
    std::vector<int> ranks;
    if (P <= 32) {
        for (int r = 0; r < P; ++r) ranks.push_back(r);
    } else {
        ranks = {0, 1, P / 3, P / 2, P - 2, P - 1};
    }
    long long worst = 0;
    for (int r : ranks) {
        const Schedule s(strategy, r);
        worst = std::max<long long>(worst, s.arena_elements(0) + s.arena_elements(1) + s.arena_elements(2));
    }
    return worst;
}

Strategy fit_strategy_to_memory(int m, int n, int k, size_t P, const std::string& prefix, long long budget_elements) {
    Strategy s = parse_strategy(m, n, k, P, prefix);
    if (schedule_footprint_elements(s) <= budget_elements) return s;
    // tighten the Strategy's own limit (its model counts the reference's buffers, ours are leaner) until the real arenas fit
    long long limit = s.memory_used > 0 ? s.memory_used : std::numeric_limits<long long>::max() / 4;
    const Strategy::quiet_errors hush;  // infeasible limits are part of the search, not news
    for (int it = 0; it < 600; ++it) {
        limit = limit - std::max<long long>(limit / 48, 1);  // fine steps: every limit may select a different splitting
        if (limit <= 0) break;
        try {
            s = parse_strategy(m, n, k, P, prefix, limit);
        } catch (const std::exception&) {
            break;  // the Strategy cannot go lower
        }
        if (schedule_footprint_elements(s) <= budget_elements) return s;
    }
    throw std::runtime_error("cosma: the multiplication does not fit " + std::to_string(budget_elements) +
                             " elements of device memory per rank even with sequential steps");
}

Strategy automatic_strategy(int m, int n, int k, size_t P, const std::string& steps, size_t elem_bytes) {
    const long long reference_limit = get_max_memory_elements(elem_bytes);
    const char* dev = std::getenv("COSMA_B200_DEVICE_MEMORY_MB");
    if (dev && *dev && std::atoll(dev) > 0) {
        const long long budget = std::atoll(dev) * 1024LL * 1024LL / static_cast<long long>(elem_bytes);
        Strategy s = fit_strategy_to_memory(m, n, k, P, steps, budget);
        if (reference_limit == std::numeric_limits<long long>::max() || s.memory_used <= reference_limit) return s;
    }
    return parse_strategy(m, n, k, P, steps, reference_limit);
}

}  // namespace cosma
