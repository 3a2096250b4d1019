// Restatement of the reference strategy search (src/cosma/strategy.cpp). Structure is ours; every decision --
// which dimension is split when, by which divisor, when ranks are dropped, when sequential steps are inserted --
// follows the reference so that step lists are identical for identical inputs.
#include <cosma/strategy.hpp>
#include <cosma/environment_variables.hpp>

#include <algorithm>
#include <array>
#include <iostream>
#include <sstream>
#include <stdexcept>

namespace cosma {

namespace {
// peak memory of the two largest communication rounds of each matrix (buffers ping-pong, strategy.cpp:293-326)
long long two_largest_sum(std::vector<long long> v) {
    std::sort(v.rbegin(), v.rend());
    long long s = 0;
    for (size_t i = 0; i < v.size() && i < 2; ++i) s += v[i];
    return s;
}
long long buffered_memory(const std::vector<long long>& a, const std::vector<long long>& b, const std::vector<long long>& c) {
    return two_largest_sum(a) + two_largest_sum(b) + two_largest_sum(c);
}

// Memory needed by the replicated matrix of each of the (up to three) parallel splits divm/divn/divk, taken in
// order of decreasing dimension (ties: smaller divisor first), shrinking P and the split dimension as we go
// (strategy.cpp:245-291). Splitting m replicates B, n replicates A, k "replicates" (reduces) C.
std::array<long long, 3> replication_memory(long long m, long long n, long long k, int divm, int divn, int divk, int P) {
    struct Dim { long long len; int div; int mat; };  // mat: 0 = A, 1 = B, 2 = C
    std::array<Dim, 3> dims = {{{m, divm, 1}, {n, divn, 0}, {k, divk, 2}}};
    std::sort(dims.begin(), dims.end(), [](const Dim& x, const Dim& y) {
        return x.len > y.len || (x.len == y.len && x.div < y.div);
    });
    std::array<long long, 3> mem = {0, 0, 0};
    for (int i = 0; i < 3; ++i) {
        if (dims[i].div <= 1) continue;
        const long long other = dims[(i + 1) % 3].len * dims[(i + 2) % 3].len;
        mem[dims[i].mat] = math_utils::divide_and_round_up(other * dims[i].div, P);
        P /= dims[i].div;
        dims[i].len /= dims[i].div;
    }
    return mem;
}
}  // namespace

int Strategy::get_min_dim_size() {
    static const int cached = get_min_local_dimension();
    return cached;
}

Strategy::Strategy() = default;
Strategy::Strategy(const Strategy& other) = default;

Strategy::Strategy(int mm, int nn, int kk, size_t PP, std::vector<int>& divs, std::string& dims, std::string& types,
                   long long mem_limit, bool top, bool overlap, bool busy_waiting)
    : m(mm), n(nn), k(kk), P(PP), memory_limit(mem_limit), divisors(divs), split_dimension(dims), step_type(types),
      topology(top), use_busy_waiting(busy_waiting), overlap_comm_and_comp(overlap) {
    bool incomplete = false;
    square_strategy(incomplete);
    check_if_valid();
    check_if_irregular();
    compute_min_sizes();
}

Strategy::Strategy(int mm, int nn, int kk, size_t PP, long long mem_limit, bool top, bool overlap, bool busy_waiting)
    : m(mm), n(nn), k(kk), P(PP), memory_limit(mem_limit), topology(top), use_busy_waiting(busy_waiting),
      overlap_comm_and_comp(overlap) {
    bool incomplete = false;
    square_strategy(incomplete);
    check_if_valid();
    check_if_irregular();
    compute_min_sizes();
}

bool Strategy::operator==(const Strategy& o) const {
    return m == o.m && n == o.n && k == o.k && P == o.P && memory_limit == o.memory_limit && divisors == o.divisors &&
           step_type == o.step_type && split_dimension == o.split_dimension &&
           overlap_comm_and_comp == o.overlap_comm_and_comp;
}

std::tuple<long long, long long, long long> Strategy::initial_memory(long long m, long long n, long long k, int P) {
    return std::make_tuple(math_utils::divide_and_round_up(m * k, P), math_utils::divide_and_round_up(k * n, P),
                           math_utils::divide_and_round_up(m * n, P));
}

// Appends one step that divides `dim_label` by `divisor`, unless that would push the local dimension under
// COSMA_MIN_LOCAL_DIMENSION: then the largest admissible smaller divisor is used, or the step is dropped; for
// parallel steps the ranks that can no longer be used are removed from P (they idle). strategy.cpp:119-183.
bool Strategy::add_step(long long& prev_m, long long& prev_n, long long& prev_k, int& prev_P, char step, char dim_label,
                        int divisor) {
    long long& dim = dim_label == 'm' ? prev_m : (dim_label == 'n' ? prev_n : prev_k);
    const int floor_dim = get_min_dim_size();
    int used = divisor;
    if (dim / divisor < floor_dim) {
        const int smaller = static_cast<int>(dim / floor_dim);
        used = (smaller > 1 && dim / smaller >= floor_dim) ? smaller : 0;
        if (step == 'p') {
            const int keep = used ? used : 1;
            P = P / divisor * keep;
            prev_P = prev_P / divisor * keep;
        }
        if (!used) return false;
    } else if (step == 'p') {
        prev_P /= divisor;
    }
    split_dimension += dim_label;
    step_type += step;
    divisors.push_back(used);
    dim /= used;
    return true;
}

// Consumes prime factors of one dimension's divisor while that dimension stays the largest (always at least one),
// and emits them as ONE parallel step. strategy.cpp:186-242.
bool Strategy::divide(std::vector<int>& factors, int& pos, long long& m_, long long& n_, long long& k_, int& P_,
                      const char label) {
    const long long dim1 = label == 'm' ? m_ : (label == 'n' ? n_ : k_);
    const long long other = label == 'm' ? std::max(n_, k_) : (label == 'n' ? std::max(m_, k_) : std::max(m_, n_));
    const int nf = static_cast<int>(factors.size());
    int next = pos < nf ? factors[pos] : 1;
    int taken = 1;
    bool any = false, first = true;
    bool largest = dim1 >= other;
    while (pos < nf && (largest || first)) {
        taken = next;
        any = true;
        ++pos;
        if (pos >= nf) break;
        next *= factors[pos];
        first = false;
        largest = dim1 / taken >= other;
    }
    return any ? add_step(m_, n_, k_, P_, 'p', label, taken) : false;
}

void Strategy::square_strategy(bool& incomplete_strategy) {
    long long m_ = m, n_ = n, k_ = k;
    int P_ = static_cast<int>(P);
    memory_used = 0;

    long long ia, ib, ic;
    std::tie(ia, ib, ic) = initial_memory(m_, n_, k_, P_);
    std::vector<long long> mem_a = {ia}, mem_b = {ib}, mem_c = {ic};

    // replay the steps given by the caller
    for (size_t i = 0; i < divisors.size(); ++i) {
        const int div = divisors[i];
        if (step_type[i] == 'p') {
            if (!split_A(i)) mem_a.push_back(math_utils::divide_and_round_up(m_ * k_ * div, P_));
            else if (!split_B(i)) mem_b.push_back(math_utils::divide_and_round_up(k_ * n_ * div, P_));
            else mem_c.push_back(math_utils::divide_and_round_up(m_ * n_ * div, P_));
            P_ /= div;
        }
        m_ /= divisor_m(i);
        n_ /= divisor_n(i);
        k_ /= divisor_k(i);
    }

    incomplete_strategy = P_ > 1;
    if (!incomplete_strategy) {
        memory_used = buffered_memory(mem_a, mem_b, mem_c);
        if (memory_limit < memory_used)
            throw_exception("This multiplication requires the memory for at least " + std::to_string(memory_used) +
                            " units, but only " + std::to_string(memory_limit) +
                            " units are allowed. Either increase the memory limit or change the strategy by using more "
                            "sequential steps.");
        return;
    }

    const std::string no_memory =
        "Not enough memory for this strategy. Either decrease the min_dim_size in the strategy to allow dimensions to be "
        "further split OR increase the memory limit in the strategy to allow COSMA to use more memory.";

    int divm, divn, divk;
    auto memory_with = [&](int dm, int dn, int dk) {
        const auto extra = replication_memory(m_, n_, k_, dm, dn, dk, P_);
        auto a = mem_a, b = mem_b, c = mem_c;
        a.push_back(extra[0]);
        b.push_back(extra[1]);
        c.push_back(extra[2]);
        return buffered_memory(a, b, c);
    };
    std::tie(divm, divn, divk) = math_utils::balanced_divisors(m_, n_, k_, P_, get_min_dim_size());
    long long used = memory_with(divm, divn, divk);
    // not enough memory: halve the largest dimension sequentially and search again on the smaller problem
    while (used > memory_limit) {
        const char dim = (m_ >= std::max(k_, n_)) ? 'm' : ((n_ >= std::max(m_, k_)) ? 'n' : 'k');
        if (!add_step(m_, n_, k_, P_, 's', dim, 2)) throw_exception(no_memory);
        std::tie(divm, divn, divk) = math_utils::balanced_divisors(m_, n_, k_, P_, get_min_dim_size());
        used = memory_with(divm, divn, divk);
    }
    memory_used = used;
    P_ = divm * divn * divk;

    // interleave the prime factors of divm, divn, divk: always split the currently largest dimension
    std::vector<int> fm = math_utils::decompose(divm), fn = math_utils::decompose(divn), fk = math_utils::decompose(divk);
    int mi = 0, ni = 0, ki = 0;
    const int total = static_cast<int>(fm.size() + fn.size() + fk.size());
    while (mi + ni + ki < total) {
        const long long mm = mi >= static_cast<int>(fm.size()) ? 1 : m_;
        const long long nn = ni >= static_cast<int>(fn.size()) ? 1 : n_;
        const long long kk = ki >= static_cast<int>(fk.size()) ? 1 : k_;
        if (mm >= std::max(nn, kk) && divide(fm, mi, m_, n_, k_, P_, 'm')) continue;
        if (nn >= std::max(mm, kk) && divide(fn, ni, m_, n_, k_, P_, 'n')) continue;
        if (kk >= std::max(mm, nn) && divide(fk, ki, m_, n_, k_, P_, 'k')) continue;
        throw_exception(no_memory);
    }

    // merge runs: consecutive parallel steps on the same dimension multiply; a run of sequential steps collapses
    // to at most one step per dimension, in the order m, n, k
    std::vector<int> new_div;
    std::string new_dim, new_type;
    P = 1;
    for (size_t i = 0; i < divisors.size();) {
        if (step_type[i] == 'p') {
            int div = divisors[i];
            size_t j = i + 1;
            while (j < divisors.size() && step_type[j] == 'p' && split_dimension[j] == split_dimension[i]) div *= divisors[j++];
            new_type += 'p';
            new_dim += split_dimension[i];
            new_div.push_back(div);
            P *= div;
            i = j;
        } else {
            int prod[3] = {1, 1, 1};
            size_t j = i;
            while (j < divisors.size() && step_type[j] == 's') {
                prod[split_dimension[j] == 'm' ? 0 : (split_dimension[j] == 'n' ? 1 : 2)] *= divisors[j];
                ++j;
            }
            for (int d = 0; d < 3; ++d)
                if (prod[d] > 1) {
                    new_dim += "mnk"[d];
                    new_type += 's';
                    new_div.push_back(prod[d]);
                }
            i = j;
        }
    }
    split_dimension = new_dim;
    step_type = new_type;
    divisors = new_div;
}

namespace {
thread_local int g_quiet_strategy_errors = 0;
}
Strategy::quiet_errors::quiet_errors() { ++g_quiet_strategy_errors; }
Strategy::quiet_errors::~quiet_errors() { --g_quiet_strategy_errors; }

void Strategy::throw_exception(const std::string& message) {
    // the reference prints the offending strategy before throwing (strategy.cpp:571-576); searches that EXPECT failures silence it
    if (g_quiet_strategy_errors == 0) std::cout << "Splitting strategy not well defined.\n" << message << std::endl << *this << std::endl;
    throw std::runtime_error(message);
}

bool Strategy::split(char label, size_t i) const {
    return label == 'A' ? split_A(i) : (label == 'B' ? split_B(i) : split_C(i));
}
int Strategy::divisor_row(char matrix, size_t i) const {
    if (matrix == 'A' || matrix == 'C') return divisor_m(i);
    if (matrix == 'B') return divisor_k(i);
    return 1;
}
int Strategy::divisor_col(char matrix, size_t i) const {
    if (matrix == 'A') return divisor_k(i);
    if (matrix == 'B' || matrix == 'C') return divisor_n(i);
    return 1;
}
int Strategy::parallel_steps_before_gemm(char label) const {
    if (label == 'A') return n_parallel_steps_before_gemm_a;
    if (label == 'B') return n_parallel_steps_before_gemm_b;
    if (label == 'C') return n_parallel_steps_before_gemm_c;
    return -1;
}

// strategy.cpp:644-791
void Strategy::check_if_valid() {
    if (empty() && P != 1) throw_exception("Strategy empty but number of ranks P != 1");
    int mi = m, ni = n, ki = k, Pi = static_cast<int>(P);
    n_parallel_steps = 0;
    n_parallel_steps_before_gemm_a = n_parallel_steps_before_gemm_b = n_parallel_steps_before_gemm_c = 0;
    int P_a = 1, P_b = 1, P_c = 1;  // ranks sharing one block of A / B / C

    for (size_t i = 0; i < n_steps(); ++i) {
        const int div = divisors[i];
        if (div <= 1)
            throw_exception("Divisors in each step must be larger than 1.Divisor in step " + std::to_string(i) + " = " +
                            std::to_string(div) + ".");
        const char dim = split_dimension[i], type = step_type[i];
        if (dim != 'm' && dim != 'n' && dim != 'k') throw_exception("Split dimension in each step must be m, n or k");
        if (type != 'p' && type != 's') throw_exception("Step type should be either p or s.");

        if (type == 'p') {
            ++n_parallel_steps;
            if (!split_A(i)) ++n_parallel_steps_before_gemm_a;
            if (!split_B(i)) ++n_parallel_steps_before_gemm_b;
            if (!split_C(i)) ++n_parallel_steps_before_gemm_c;
            if (Pi <= 1)
                throw_exception("Not enough processors for this division strategy.The product of all divisors in a parallel "
                                "step should be equal to the number of processors");
            if (Pi % div != 0)
                throw_exception("The number of processors left in each parallel step should be divisible by divisor.");
            Pi /= div;
            if (!split_A(i)) P_a *= div;
            else if (!split_B(i)) P_b *= div;
            else if (!split_C(i)) P_c *= div;
            else throw_exception("Invalid strategy: In each step, some matrix has to be split.");
        } else {
            ++n_sequential_steps;
            if (split_A(i)) n_parallel_steps_before_gemm_a = 0;
            if (split_B(i)) n_parallel_steps_before_gemm_b = 0;
            if (split_C(i)) n_parallel_steps_before_gemm_c = 0;
        }

        (dim == 'm' ? mi : (dim == 'n' ? ni : ki)) /= div;

        // column-major pieces: a block shared by q ranks needs at least q columns (only n and k count columns)
        if (i + 1 == n_steps()) {
            if (ki < P_a)
                throw_exception("Dimension k at step " + std::to_string(i) + " = " + std::to_string(ki) +
                                ", which is less than the number of processors left = " + std::to_string(P_a));
            if (ni < std::max(P_b, P_c))
                throw_exception("Dimension n at step " + std::to_string(i) + " = " + std::to_string(ni) +
                                ", which is less than the number of processors left = " + std::to_string(std::min(P_b, P_c)));
        }
    }
    if (Pi != 1)
        throw_exception("Too many processors. The number of processors should be equal to the product of divisors in all "
                        "parallel steps.");
}

void Strategy::compress_steps() {
    int prod[6] = {1, 1, 1, 1, 1, 1};  // p:m,n,k then s:m,n,k
    for (size_t i = 0; i < n_steps(); ++i) {
        const int base = parallel_step(i) ? 0 : 3;
        prod[base + 0] *= divisor_m(i);
        prod[base + 1] *= divisor_n(i);
        prod[base + 2] *= divisor_k(i);
    }
    divisors.clear();
    split_dimension.clear();
    step_type.clear();
    for (int i = 0; i < 6; ++i)
        if (prod[i] > 1) {
            divisors.push_back(prod[i]);
            step_type += i < 3 ? 'p' : 's';
            split_dimension += "mnk"[i % 3];
        }
}

void Strategy::compute_min_sizes() {
    min_m = m; min_n = n; min_k = k;
    for (size_t s = 0; s < n_steps(); ++s) {
        min_m /= divisor_m(s);
        min_n /= divisor_n(s);
        min_k /= divisor_k(s);
    }
}

// strategy.cpp:851-901
bool Strategy::should_overlap_comm_and_comp(int step) const {
    if (step != static_cast<int>(n_steps()) - 1) return false;
    const int div = divisor(step);
    const bool possible = (split_m(step) && min_n >= div) || (split_n(step) && min_k >= div) || (split_k(step) && min_n >= div);
    int newm = min_m, newn = min_n, newk = min_k;
    if (split_n(step)) newk /= div; else newn /= div;
    const double before = math_utils::square_score(min_m, min_n, min_k);
    const double after = math_utils::square_score(newm, newn, newk);
    return possible && overlap_comm_and_comp && (after - before) / before >= 0.5;
}

int Strategy::n_rows(char label) const { return label == 'A' || label == 'C' ? m : (label == 'B' ? k : -1); }
int Strategy::n_cols(char label) const { return label == 'A' ? k : (label == 'B' || label == 'C' ? n : -1); }

// strategy.cpp:929-955
void Strategy::enable_overlapping_comm_and_comp() {
    if (empty()) return;
    const int last = static_cast<int>(n_steps()) - 1;
    if (split_m(last) && min_n >= divisor_m(last)) {
        overlap_comm_and_comp = true;
        irregular = irregular || (min_n % divisor_m(last) != 0);
    } else if (split_n(last) && min_k >= divisor_n(last)) {
        overlap_comm_and_comp = true;
        irregular = irregular || (min_k % divisor_n(last) != 0);
    } else if (split_k(last) && min_n >= divisor_k(last)) {
        overlap_comm_and_comp = true;
        irregular = irregular || (min_n % divisor_k(last) != 0);
    }
}

void Strategy::check_if_irregular() {
    int mm = m, nn = n, kk = k;
    irregular = true;
    for (size_t i = 0; i < n_steps(); ++i) {
        if (mm % divisor_m(i) != 0 || nn % divisor_n(i) != 0 || kk % divisor_k(i) != 0) return;
        mm /= divisor_m(i);
        nn /= divisor_n(i);
        kk /= divisor_k(i);
    }
    irregular = false;
}

std::string Strategy::to_string() const {
    std::string s;
    for (size_t i = 0; i < n_steps(); ++i) {
        if (i) s += ',';
        s += step_type[i];
        s += split_dimension[i];
        s += std::to_string(divisors[i]);
    }
    return s;
}

std::ostream& operator<<(std::ostream& os, const Strategy& st) {
    os << "Matrix dimensions (m, n, k) = (" << st.m << ", " << st.n << ", " << st.k << ")\n";
    os << "Number of processors: " << st.P << "\n";
    if (st.topology) os << "Communication-aware topology turned on.\n";
    os << "Overlap of communication and computation: " << (st.overlap_comm_and_comp ? "ON" : "OFF") << ".\n";
    os << "Divisions strategy: \n";
    for (size_t i = 0; i < st.n_steps(); ++i)
        os << (st.step_type[i] == 'p' ? "parallel (" : "sequential (") << st.split_dimension[i] << " / " << st.divisors[i]
           << ")\n";
    os << "Required memory per rank (in #elements): " << st.memory_used << "\n";
    os << "Available memory per rank (in #elements): ";
    if (st.memory_limit < std::numeric_limits<long long>::max()) os << st.memory_limit;
    else os << "not specified (assumed: infinite)";
    os << "\n";
    return os;
}

Strategy parse_strategy(int m, int n, int k, size_t P, const std::string& steps, long long memory_limit, bool overlap) {
    std::vector<int> divs;
    std::string dims, types;
    std::stringstream ss(steps);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        // tolerate spaces and the reference's "-s" spellings such as "pm2"
        tok.erase(std::remove_if(tok.begin(), tok.end(), [](unsigned char c) { return std::isspace(c); }), tok.end());
        if (tok.empty()) continue;
        if (tok.size() < 3 || (tok[0] != 'p' && tok[0] != 's') || (tok[1] != 'm' && tok[1] != 'n' && tok[1] != 'k'))
            throw std::runtime_error("cannot parse strategy step '" + tok + "' (expected e.g. pm2, sn4, pk2)");
        types += tok[0];
        dims += tok[1];
        divs.push_back(std::stoi(tok.substr(2)));
    }
    if (divs.empty()) return Strategy(m, n, k, P, memory_limit, false, overlap);
    return Strategy(m, n, k, P, divs, dims, types, memory_limit, false, overlap);
}

}  // namespace cosma
