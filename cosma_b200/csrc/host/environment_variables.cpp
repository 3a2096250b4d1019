#include <cosma/environment_variables.hpp>

#include <algorithm>
#include <cctype>
#include <cstdlib>

namespace cosma {
bool env_var_defined(const char* name) { return std::getenv(name) != nullptr; }

bool get_bool_env_var(const std::string& name, bool default_value) {
    const char* v = std::getenv(name.c_str());
    if (!v) return default_value;
    std::string s(v);
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::toupper(c); });
    if (s == "ON" || s == "TRUE" || s == "1") return true;
    if (s == "OFF" || s == "FALSE" || s == "0") return false;
    return default_value;
}

int get_int_env_var(const std::string& name, int default_value) {
    const char* v = std::getenv(name.c_str());
    return v ? std::atoi(v) : default_value;
}

int get_min_local_dimension() { return get_int_env_var("COSMA_MIN_LOCAL_DIMENSION", 200); }
int get_cosma_dim_threshold() { return get_int_env_var("COSMA_DIM_THRESHOLD", 0); }
bool get_adapt_strategy() { return get_bool_env_var("COSMA_ADAPT_STRATEGY", true); }
bool get_overlap_comm_and_comp() { return get_bool_env_var("COSMA_OVERLAP_COMM_AND_COMP", false); }

long long get_max_memory_elements(std::size_t elem_bytes) {
    const char* v = std::getenv("COSMA_CPU_MAX_MEMORY");
    if (!v) return std::numeric_limits<long long>::max();
    return std::atoll(v) * 1024LL * 1024LL / static_cast<long long>(elem_bytes);
}
}  // namespace cosma
