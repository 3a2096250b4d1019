#include <cosma/math_utils.hpp>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>

namespace cosma {
namespace math_utils {

int gcd(int a, int b) {
    while (b != 0) {
        const int r = a % b;
        a = b;
        b = r;
    }
    return a;
}

long long divide_and_round_up(long long x, long long y) { return 1 + (x - 1) / y; }

int next_multiple_of(int n_to_round, int multiple) {
    if (multiple == 0) return n_to_round;
    const int rem = n_to_round % multiple;
    return rem == 0 ? n_to_round : n_to_round + multiple - rem;
}

std::vector<int> find_divisors(int n) {
    std::vector<int> small, large;
    for (int d = 1; static_cast<long long>(d) * d <= n; ++d) {
        if (n % d) continue;
        small.push_back(d);
        if (d != n / d) large.push_back(n / d);
    }
    small.insert(small.end(), large.rbegin(), large.rend());
    return small;  // ascending, identical to the reference's linear scan
}

// The scoring deliberately keeps the reference's arithmetic (math_utils.cpp:86-116): tile sizes m/d are INTEGER
// quotients, the deviation from the cubic target is a double that is truncated to int, candidates that use more
// ranks always win, ties are broken by the smaller truncated error.
std::tuple<int, int, int> balanced_divisors(long long m, long long n, long long k, int P, int min_local_problem_size) {
    const long long cap_m = std::max(1LL, std::min(std::min(m, n), m / min_local_problem_size));
    const long long cap_n = std::max(1LL, std::min(std::min(k, n), n / min_local_problem_size));
    const long long cap_k = std::max(1LL, std::min(std::min(k, n), k / min_local_problem_size));

    if (cap_m < P && cap_n < P && cap_k < P && cap_m * cap_n < P && cap_m * cap_n * cap_k < P)
        P = static_cast<int>(cap_m * cap_n * cap_k);

    int d[3] = {static_cast<int>(m), static_cast<int>(n), static_cast<int>(k)};
    std::sort(d, d + 3);
    double target;
    if (d[2] >= P) target = std::cbrt(1.0 * d[2] / P * d[0] * d[1]);
    else if (d[1] * d[2] >= P) target = std::cbrt(1.0 * d[1] * d[2] / P * d[0]);
    else target = std::cbrt(1.0 * d[0] * d[1] * d[2] / P);

    int best_err = std::numeric_limits<int>::max();
    int bm = 1, bn = 1, bk = 1;
    for (int dm : find_divisors(P)) {
        if (dm > cap_m) break;
        const int lower_bound = static_cast<int>(std::abs(m / dm - target));
        if (lower_bound > best_err) continue;
        for (int dn : find_divisors(P / dm)) {
            if (dn > cap_n) break;
            const int dk = std::min((P / dm) / dn, static_cast<int>(cap_k));
            const int err = static_cast<int>(std::abs(m / dm - target) + std::abs(n / dn - target) + std::abs(k / dk - target));
            const int used = dm * dn * dk, best_used = bm * bn * bk;
            if (used > best_used || (used == best_used && err < best_err)) {
                bm = dm; bn = dn; bk = dk;
                best_err = err;
            }
        }
    }
    return std::make_tuple(bm, bn, bk);
}

std::vector<int> decompose(int n) {
    std::vector<int> factors;
    while (n % 2 == 0) { factors.push_back(2); n /= 2; }
    for (int f = 3; f <= std::sqrt(n); f += 2)
        while (n % f == 0) { factors.push_back(f); n /= f; }
    if (n > 2) factors.push_back(n);
    return factors;
}

int closest_divisor(int P, int dimension, double target) {
    int best_div = 1, best_err = std::numeric_limits<int>::max();
    for (int d : find_divisors(P)) {
        const int err = static_cast<int>(std::abs(1.0 * dimension / d - target));
        if (err <= best_err) { best_div = d; best_err = err; }
    }
    return best_div;
}

int int_div_up(int numerator, int denominator) {
    return numerator / denominator + (((numerator < 0) ^ (denominator > 0)) && (numerator % denominator));
}

double square_score(int rows, int cols) {
    const double r1 = 1.0 * rows / cols, r2 = 1.0 * cols / rows;
    return (r1 + r2) / (2.0 * std::max(r1, r2));
}
double square_score(int m, int n, int k) { return square_score(m, k) * square_score(k, n) * square_score(m, n); }

std::pair<int, int> invert_cantor_pairing(int z) {
    const int w = static_cast<int>(std::floor((std::sqrt(8.0 * z + 1) - 1) / 2));
    const int y = z - (w * w + w) / 2;
    return {w - y, y};
}
int cantor_pairing(int i, int j) { return (i + j) * (i + j + 1) / 2 + j; }

bool is_power_of_2(std::size_t n) { return !(n & (n - 1)); }
std::size_t next_greater_power_of_2(std::size_t n, std::size_t p) {
    while (n != 0) { n -= (n & p); p <<= 1; }
    return p;
}
std::size_t next_power_of_2(std::size_t n) { return is_power_of_2(n) ? n : next_greater_power_of_2(n); }

}  // namespace math_utils
}  // namespace cosma
