// Integer helpers of the strategy search (reference src/cosma/math_utils.hpp, math_utils.cpp:4-226).
#pragma once
#include <cstddef>
#include <tuple>
#include <utility>
#include <vector>

namespace cosma {
namespace math_utils {
int gcd(int a, int b);
long long divide_and_round_up(long long x, long long y);
int next_multiple_of(int n_to_round, int multiple);
std::vector<int> find_divisors(int n);
// divisors (dm, dn, dk) with dm*dn*dk <= P making m/dm, n/dn, k/dk as cubic as possible (math_utils.cpp:47-118)
std::tuple<int, int, int> balanced_divisors(long long m, long long n, long long k, int P, int min_local_problem_size);
std::vector<int> decompose(int n);  // prime factors, ascending
int closest_divisor(int P, int dimension, double target);
int int_div_up(int numerator, int denominator);
double square_score(int rows, int cols);
double square_score(int m, int n, int k);
std::pair<int, int> invert_cantor_pairing(int z);
int cantor_pairing(int i, int j);
bool is_power_of_2(std::size_t n);
std::size_t next_greater_power_of_2(std::size_t n, std::size_t power_of_2 = 1);
std::size_t next_power_of_2(std::size_t n);
}  // namespace math_utils
}  // namespace cosma
