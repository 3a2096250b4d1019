// Plan-time lowering of [ALLGATHER A] [ALLGATHER B] GEMM [REDUCE C] into a two-stream micro-op program (see overlap.hpp).
#include <cosma/overlap.hpp>

#include <cosma/environment_variables.hpp>

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace cosma {

namespace {

struct Range {
    std::int64_t lo = 0, hi = 0;
    std::int64_t len() const { return hi > lo ? hi - lo : 0; }
    bool empty() const { return hi <= lo; }
};
Range intersect(Range a, Range b) { return Range{std::max(a.lo, b.lo), std::min(a.hi, b.hi)}; }

// Cost model of the persistent GEMM kernels: 128 x 128 (real) output tiles handed out round-robin to `ctas` CTAs, so a launch
// lasts ceil(tiles / ctas) tile times whatever the fill of its last wave.
struct Model {
    OverlapTuning t;
    std::int64_t m = 0;
    std::int64_t tiles_m() const { return ((t.complex_type ? 2 : 1) * m + 127) / 128; }
    double tile_ms(std::int64_t k) const {
        const double flops = (t.complex_type ? 8.0 * 64 : 2.0 * 128) * 128.0 * static_cast<double>(k);
        return flops / (t.sm_gflops * 1e9) * 1e3;
    }
    std::int64_t tiles(std::int64_t cols) const { return tiles_m() * ((cols + 127) / 128); }
    double gemm_ms(std::int64_t cols, std::int64_t k, int ctas) const {
        if (cols <= 0 || k <= 0) return 0.0;
        const std::int64_t T = tiles(cols);
        return static_cast<double>((T + ctas - 1) / ctas) * tile_ms(k);
    }
    double transfer_ms(std::int64_t elements, double gbps) const { return static_cast<double>(elements) * t.elem_bytes / (gbps * 1e6); }
    // Width (a multiple of the granule, at most `avail`) of a panel of depth k on `ctas` CTAs that lasts at least `cover_ms` and wastes
    // as little of its last wave as possible.
    std::int64_t choose_width(std::int64_t avail, std::int64_t k, int ctas, double cover_ms) const {
        const std::int64_t g = std::max(1, t.col_granule);
        const std::int64_t n_g = avail / g;
        if (n_g <= 1) return avail;
        std::int64_t c_min = 0;
        for (std::int64_t c = 1; c <= n_g; ++c)
            if (gemm_ms(c * g, k, ctas) >= cover_ms) { c_min = c; break; }
        if (c_min == 0) return avail;  // even the whole region is shorter than the transfer
        std::int64_t best = c_min;
        double best_waste = 2.0;
        for (std::int64_t c = c_min; c <= std::min(n_g, 2 * c_min + 8); ++c) {
            const std::int64_t T = tiles(c * g), waves = (T + ctas - 1) / ctas;
            const double waste = static_cast<double>(waves * ctas - T) / static_cast<double>(waves * ctas);
            if (waste < best_waste - 1e-12) { best_waste = waste; best = c; }
        }
        return best * g;
    }
};

}  // namespace

OverlapTuning overlap_tuning_from_env(char dtype, int sms) {
    OverlapTuning t;
    t.sms = sms > 0 ? sms : 148;
    t.complex_type = dtype == 'z' || dtype == 'c';
    t.elem_bytes = (dtype == 'd' || dtype == 'z' ? 8 : 4) * (t.complex_type ? 2 : 1);
    t.sm_gflops = (dtype == 'd' || dtype == 'z') ? 250.0 : 1080.0;
    if (const char* v = std::getenv("COSMA_OVERLAP_COMM_AND_COMP")) {
        std::string s(v);
        for (auto& c : s) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
        if (s == "OFF" || s == "0" || s == "FALSE") t.enabled = false;
        if (s == "FORCE") t.force = true;
    }
    auto num = [](const char* name, double dflt) {
        const char* v = std::getenv(name);
        return v && *v ? std::atof(v) : dflt;
    };
    t.reserved_sms = static_cast<int>(num("COSMA_B200_OVERLAP_SMS", t.reserved_sms));
    t.reserved_sms = std::max(1, std::min(t.reserved_sms, t.sms / 2));
    t.link_gbps = std::max(1.0, num("COSMA_B200_OVERLAP_GBPS", t.link_gbps));
    t.cover = std::max(0.0, num("COSMA_B200_OVERLAP_COVER", t.cover));
    t.col_granule = std::max(1, static_cast<int>(num("COSMA_B200_OVERLAP_GRANULE", t.col_granule)));
    return t;
}

OverlapProgram plan_overlap(const Schedule& sch, const OverlapTuning& t) {
    OverlapProgram out;
    auto no = [&](const std::string& why) {
        out.enabled = false;
        out.why = why;
        out.ops.clear();
        return out;
    };
    if (!t.enabled) return no("switched off (COSMA_OVERLAP_COMM_AND_COMP)");
    const auto& ops = sch.ops();
    int gi = -1;
    for (size_t i = 0; i < ops.size(); ++i)
        if (ops[i].kind == OpKind::GEMM) {
            if (gi >= 0) return no("more than one base-case GEMM (sequential steps)");
            gi = static_cast<int>(i);
        }
    if (gi < 0) return no("no GEMM");
    const ScheduleOp& g = ops[gi];
    const std::int64_t m = g.m, n = g.n, k = g.k;
    if (m <= 0 || n <= 0 || k <= 0) return no("empty GEMM");

    // the allgathers that feed the GEMM directly (at most one per operand), and the reduce that takes its result
    int ag[2] = {-1, -1};
    int first_async = gi;
    for (int i = gi - 1; i >= 0; --i) {
        const ScheduleOp& o = ops[i];
        if (o.kind != OpKind::ALLGATHER || !o.regular || o.ring.size() != 2 || o.matrix > 1 || ag[o.matrix] >= 0) break;
        if (o.dst_off != (o.matrix == 0 ? g.a_off : g.b_off)) break;
        ag[o.matrix] = i;
        first_async = i;
    }
    int red = -1;
    if (gi + 1 < static_cast<int>(ops.size())) {
        const ScheduleOp& o = ops[gi + 1];
        if (o.kind == OpKind::REDUCE && o.regular && o.ring.size() == 2 && o.src_off == g.c_off) red = gi + 1;
    }
    if (ag[0] < 0 && ag[1] < 0 && red < 0) return no("no ring-of-two collective next to the GEMM");

    // geometry: own k block of A, own column block of B, this rank's / the peer's column half of C
    Range Ko{0, k}, Kp{0, 0}, Bo{0, n}, mine{0, 0}, peer{0, n};
    std::int64_t ag_elems = 0, ex_elems = 0;
    if (ag[0] >= 0) {
        const ScheduleOp& o = ops[ag[0]];
        const std::int64_t cnt = o.piece[0][0];
        if (cnt % m != 0 || 2 * cnt != m * k) return no("A pieces are not k blocks of the GEMM operand");
        const std::int64_t kb = cnt / m;
        Ko = Range{o.my_pos * kb, (o.my_pos + 1) * kb};
        Kp = Range{(1 - o.my_pos) * kb, (2 - o.my_pos) * kb};
        ag_elems += cnt;
    }
    if (ag[1] >= 0) {
        const ScheduleOp& o = ops[ag[1]];
        const std::int64_t cnt = o.piece[0][0];
        if (cnt % k != 0 || 2 * cnt != k * n) return no("B pieces are not column blocks of the GEMM operand");
        const std::int64_t nb = cnt / k;
        Bo = Range{o.my_pos * nb, (o.my_pos + 1) * nb};
        ag_elems += cnt;
    }
    if (red >= 0) {
        const ScheduleOp& o = ops[red];
        const std::int64_t cnt = o.piece[0][0];
        if (cnt % m != 0 || 2 * cnt != m * n) return no("C pieces are not column blocks of the GEMM result");
        const std::int64_t nc = cnt / m;
        mine = Range{o.my_pos * nc, (o.my_pos + 1) * nc};
        peer = Range{(1 - o.my_pos) * nc, (2 - o.my_pos) * nc};
        ex_elems = cnt;
        if (o.beta != BetaMode::ZERO && o.tmp_off < 0) return no("reduce without a staging region");
    }

    Model model{t, m};
    const int full = t.sms, narrow = std::max(1, t.sms - t.reserved_sms);
    const double t_ag = model.transfer_ms(ag_elems, t.link_gbps), t_ex = model.transfer_ms(ex_elems, t.link_gbps);
    const bool have_ag = ag[0] >= 0 || ag[1] >= 0;

    std::vector<MicroOp>& P = out.ops;
    auto add = [&](const MicroOp& o) {
        P.push_back(o);
        return static_cast<int>(P.size()) - 1;
    };
    bool aligned = true;
    double est = 0.0;
    auto gemm = [&](Range cols, Range ks, BetaMode beta, bool is_narrow, bool from_pieces, const std::vector<int>& wait) {
        if (cols.empty() || ks.empty()) return -1;
        MicroOp o;
        o.kind = MicroKind::GEMM;
        o.stream = 0;
        o.wait = wait;
        o.m = static_cast<int>(m);
        o.n = static_cast<int>(cols.len());
        o.k = static_cast<int>(ks.len());
        o.beta = beta;
        o.narrow = is_narrow;
        o.lda = m; o.ldb = k; o.ldc = m;
        if (from_pieces) {  // the gathers are still in flight: the own pieces where the caller (or an outer step) left them
            o.a_off = ag[0] >= 0 ? ops[ag[0]].src_off + (ks.lo - Ko.lo) * m : g.a_off + ks.lo * m;
            o.b_off = ag[1] >= 0 ? ops[ag[1]].src_off + (cols.lo - Bo.lo) * k + ks.lo : g.b_off + cols.lo * k + ks.lo;
        } else {
            o.a_off = g.a_off + ks.lo * m;
            o.b_off = g.b_off + cols.lo * k + ks.lo;
        }
        o.c_off = g.c_off + cols.lo * m;
        for (std::int64_t off : {o.a_off, o.b_off, o.c_off})
            if ((off * t.elem_bytes) % 16 != 0) aligned = false;
        est += model.gemm_ms(o.n, o.k, is_narrow ? narrow : full);
        return add(o);
    };

    // ops before the overlapped tail run as they are, on the compute stream
    int last_serial = -1;
    for (int i = 0; i < first_async; ++i) {
        MicroOp o;
        o.kind = MicroKind::SERIAL;
        o.op = i;
        last_serial = add(o);
    }
    std::vector<int> all_ag;
    for (int i = first_async; i < gi; ++i) {
        MicroOp o;
        o.kind = MicroKind::ALLGATHER;
        o.stream = 1;
        o.op = i;
        if (last_serial >= 0) o.wait.push_back(last_serial);
        all_ag.push_back(add(o));
    }

    const Range all_k{0, k};
    Range S1{0, 0};
    bool s1_in_mine = false;
    double g1_ms = 0.0;
    if (have_ag) {
        Range R1 = intersect(Bo, peer);
        if (R1.empty()) {
            R1 = intersect(Bo, mine);
            s1_in_mine = true;
        }
        const std::int64_t w1 = model.choose_width(R1.len(), Ko.len(), narrow, t.cover * t_ag);
        // at the outer end of the region, so that what is left of it stays one contiguous range
        S1 = (R1.hi == n && R1.lo != 0) ? Range{R1.hi - w1, R1.hi} : Range{R1.lo, R1.lo + w1};
        g1_ms = model.gemm_ms(S1.len(), Ko.len(), narrow);
        gemm(S1, Ko, g.beta, true, true, {});
        if (!s1_in_mine) {
            gemm(S1, Kp, BetaMode::ONE, false, false, all_ag);
            const Range rest = S1.lo == peer.lo ? Range{S1.hi, peer.hi} : Range{peer.lo, S1.lo};
            gemm(rest, all_k, g.beta, false, false, all_ag);
        } else {
            gemm(peer, all_k, g.beta, false, false, all_ag);
        }
    } else {
        gemm(peer, all_k, g.beta, false, false, {});
    }
    double exposed = std::max(0.0, t_ag - g1_ms);

    int after = gi + 1;
    if (red >= 0) {
        const ScheduleOp& r = ops[red];
        MicroOp ex;
        ex.kind = MicroKind::EXCHANGE;
        ex.stream = 1;
        ex.wait.push_back(static_cast<int>(P.size()) - 1);  // the peer's half is complete
        ex.ring_index = r.ring_index;
        ex.peer = 1 - r.my_pos;
        ex.send_off = r.src_off + peer.lo * m;
        ex.count = ex_elems;
        ex.recv_off = r.beta == BetaMode::ZERO ? r.dst_off : r.tmp_off;
        ex.recv_off_zero = r.dst_off;
        ex.beta = r.beta;
        const int ex_idx = add(ex);
        // this rank's half, the first part beside the exchange kernels
        double narrow_ms = 0.0;
        Range rest = mine;
        if (s1_in_mine) {
            if (!Kp.empty()) {
                narrow_ms += model.gemm_ms(S1.len(), Kp.len(), narrow);
                gemm(S1, Kp, BetaMode::ONE, true, false, all_ag);
            }
            rest = S1.lo == mine.lo ? Range{S1.hi, mine.hi} : Range{mine.lo, S1.lo};
        }
        const double need = t.cover * t_ex - narrow_ms;
        if (need > 0.0 && !rest.empty()) {
            const std::int64_t w = model.choose_width(rest.len(), k, narrow, need);
            narrow_ms += model.gemm_ms(w, k, narrow);
            gemm(Range{rest.lo, rest.lo + w}, all_k, g.beta, true, false, all_ag);
            rest.lo += w;
        }
        gemm(rest, all_k, g.beta, false, false, all_ag);
        exposed += std::max(0.0, t_ex - narrow_ms);
        // own half of the sum: received half (+ beta * what C held) + own partial result
        MicroOp acc;
        acc.kind = MicroKind::ACCUMULATE;
        acc.stream = 0;
        acc.dst_off = r.dst_off;
        acc.count = ex_elems;
        acc.wait.push_back(ex_idx);
        if (r.beta != BetaMode::ZERO) {
            acc.add_off = r.tmp_off;
            acc.beta = r.beta;
            acc.beta_term = true;
            add(acc);
            acc.beta_term = false;  // the own-half term still waits for the exchange: it is the first to touch C when beta == 0
        }
        acc.add_off = r.src_off + mine.lo * m;
        acc.beta = BetaMode::ONE;
        add(acc);
        est += (r.beta != BetaMode::ZERO ? 2.0 : 1.0) * model.transfer_ms(3 * ex_elems, 6000.0);
        after = red + 1;
    }
    for (int i = after; i < static_cast<int>(ops.size()); ++i) {
        MicroOp o;
        o.kind = MicroKind::SERIAL;
        o.op = i;
        add(o);
    }

    const double serial_gbps = 3.0 * t.link_gbps;  // what the unconstrained NCCL kernels of the serial schedule reach
    out.est_comm_ms = t_ag + t_ex;
    out.est_serial_ms = model.gemm_ms(n, k, full) + model.transfer_ms(ag_elems + ex_elems, serial_gbps) +
                        (red >= 0 && ops[red].beta != BetaMode::ZERO ? model.transfer_ms(3 * ex_elems, 6000.0) : 0.0);
    out.est_overlap_ms = est + exposed;
    if (!t.force) {
        if (!aligned) return no("a panel would not start on a 16-byte boundary");
        if (out.est_overlap_ms >= out.est_serial_ms) return no("estimated no gain over the serial schedule");
    }
    out.enabled = true;
    out.why = std::string("overlapped:") + (ag[0] >= 0 ? " allgather A" : "") + (ag[1] >= 0 ? " allgather B" : "") + (red >= 0 ? " reduce C" : "");
    return out;
}

std::vector<std::int64_t> OverlapProgram::serialize() const {
    std::vector<std::int64_t> v;
    for (const auto& o : ops) {
        v.push_back(static_cast<int>(o.kind));
        v.push_back(o.stream);
        v.push_back(static_cast<std::int64_t>(o.wait.size()));
        for (int w : o.wait) v.push_back(w);
        switch (o.kind) {
            case MicroKind::GEMM:
                for (std::int64_t x : {o.a_off, o.b_off, o.c_off, o.lda, o.ldb, o.ldc, static_cast<std::int64_t>(o.m), static_cast<std::int64_t>(o.n),
                                       static_cast<std::int64_t>(o.k), static_cast<std::int64_t>(static_cast<int>(o.beta)),
                                       static_cast<std::int64_t>(o.narrow ? 1 : 0)})
                    v.push_back(x);
                break;
            case MicroKind::ALLGATHER:
            case MicroKind::SERIAL:
                v.push_back(o.op);
                break;
            case MicroKind::EXCHANGE:
                for (std::int64_t x : {static_cast<std::int64_t>(o.ring_index), static_cast<std::int64_t>(o.peer), o.send_off, o.recv_off, o.recv_off_zero,
                                       o.count, static_cast<std::int64_t>(static_cast<int>(o.beta))})
                    v.push_back(x);
                break;
            case MicroKind::ACCUMULATE:
                for (std::int64_t x : {o.dst_off, o.add_off, o.count, static_cast<std::int64_t>(static_cast<int>(o.beta)),
                                       static_cast<std::int64_t>(o.beta_term ? 1 : 0)})
                    v.push_back(x);
                break;
        }
    }
    return v;
}

}  // namespace cosma
