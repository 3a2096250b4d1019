// cosma::Strategy -- the communication-optimal schedule of parallel/sequential splits of (m, n, k) over P ranks.
// Public surface and results identical to the reference (src/cosma/strategy.hpp:16-183, strategy.cpp:82-1022):
// same fields, same constructors, same step lists for the same inputs (tests/test_planning.py pins this against the
// reference build in oracle/_ref and against the golden strategy of tests/mapper.cpp:363-387).
#pragma once
#include <cosma/math_utils.hpp>

#include <cstddef>
#include <iosfwd>
#include <limits>
#include <string>
#include <tuple>
#include <vector>

namespace cosma {

class Strategy {
  public:
    int m = 0, n = 0, k = 0;
    size_t P = 0;
    long long memory_limit = 0;

    int min_m = 0, min_n = 0, min_k = 0;   // base-case dimensions induced by the steps

    std::vector<int> divisors = {};        // divisor of each step
    std::string split_dimension = "";      // 'm' | 'n' | 'k' per step
    std::string step_type = "";            // 'p' (parallel) | 's' (sequential) per step
    bool topology = false;
    bool use_busy_waiting = true;
    long long memory_used = 0;
    int n_parallel_steps = 0;
    int n_sequential_steps = 0;
    int n_parallel_steps_before_gemm_a = 0;
    int n_parallel_steps_before_gemm_b = 0;
    int n_parallel_steps_before_gemm_c = 0;
    bool irregular = true;                 // some step does not divide its dimension evenly

    Strategy();
    Strategy(const Strategy& other);
    Strategy& operator=(const Strategy& other) = default;
    // (possibly incomplete) prefix of steps given by the caller, completed automatically
    Strategy(int mm, int nn, int kk, size_t PP, std::vector<int>& divs, std::string& dims, std::string& types,
             long long mem_limit = std::numeric_limits<long long>::max(), bool top = false, bool overlap = false,
             bool busy_waiting = true);
    Strategy(int mm, int nn, int kk, size_t PP, long long mem_limit = std::numeric_limits<long long>::max(),
             bool top = false, bool overlap = false, bool busy_waiting = true);

    static int get_min_dim_size();
    size_t n_steps() const { return divisors.size(); }
    bool empty() const { return divisors.empty(); }

    void square_strategy(bool& incomplete_strategy);
    bool add_step(long long& prev_m, long long& prev_n, long long& prev_k, int& prev_P, char step, char dim_label, int divisor);
    void throw_exception(const std::string& message);
    // while one of these lives on the calling thread, throw_exception does not print (used by searches over memory limits)
    struct quiet_errors {
        quiet_errors();
        ~quiet_errors();
    };

    bool split_m(size_t i) const { return split_dimension[i] == 'm'; }
    bool split_n(size_t i) const { return split_dimension[i] == 'n'; }
    bool split_k(size_t i) const { return split_dimension[i] == 'k'; }
    bool split_A(size_t i) const { return split_m(i) || split_k(i); }
    bool split_B(size_t i) const { return split_k(i) || split_n(i); }
    bool split_C(size_t i) const { return split_m(i) || split_n(i); }
    bool split(char label, size_t i) const;
    bool sequential_step(size_t i) const { return step_type[i] == 's'; }
    bool parallel_step(size_t i) const { return step_type[i] == 'p'; }
    int divisor(size_t i) const { return divisors[i]; }
    int divisor_m(size_t i) const { return split_m(i) ? divisors[i] : 1; }
    int divisor_n(size_t i) const { return split_n(i) ? divisors[i] : 1; }
    int divisor_k(size_t i) const { return split_k(i) ? divisors[i] : 1; }
    int divisor_row(char matrix, size_t i) const;
    int divisor_col(char matrix, size_t i) const;
    bool final_step(size_t i) const { return i == n_steps(); }
    int parallel_steps_before_gemm(char label) const;

    static std::tuple<long long, long long, long long> initial_memory(long long m, long long n, long long k, int P);

    void check_if_valid();
    void check_if_irregular();
    void compress_steps();
    void compute_min_sizes();
    bool should_overlap_comm_and_comp(int step) const;
    void enable_overlapping_comm_and_comp();
    bool overlap_enabled() const { return overlap_comm_and_comp; }

    int n_rows(char label) const;
    int n_cols(char label) const;

    // "pm2,sn4,pk2" -- the notation of the reference miniapp's -s flag (utils/parse_strategy.hpp:24-61)
    std::string to_string() const;

    bool operator==(const Strategy& other) const;
    bool operator!=(const Strategy& other) const { return !(*this == other); }
    friend std::ostream& operator<<(std::ostream& os, const Strategy& other);

  private:
    bool overlap_comm_and_comp = false;
    bool divide(std::vector<int>& div_factors, int& dim_i, long long& dim1, long long& dim2, long long& dim3, int& P,
                const char label);
};

// Parses "pm2,sn4,pk2" into a Strategy prefix; empty string -> automatic strategy (utils/parse_strategy.hpp:35-61).
Strategy parse_strategy(int m, int n, int k, size_t P, const std::string& steps,
                        long long memory_limit = std::numeric_limits<long long>::max(), bool overlap = false);

}  // namespace cosma
