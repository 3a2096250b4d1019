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
    // NCCL send/recv and allgather between two NVSwitch peers move ~11 GB/s per CTA when limited to a few CTAs (measured on B200:
    // 89 GB/s with 8, 37 GB/s with 4; profiles/r2_bench_n2_*.json)
    t.link_gbps = std::max(1.0, num("COSMA_B200_OVERLAP_GBPS", 10.5 * t.reserved_sms));
    // planning as if the copy-engine transport were bound (tests of the panel programs without a GPU; cosma_b200_plan_bind_arenas sets it)
    if (const char* v = std::getenv("COSMA_B200_OVERLAP_ZERO_SM"))
        if (v[0] == 'O' && v[1] == 'N') {
            t.zero_sm = true;
            t.link_gbps = std::max(1.0, num("COSMA_B200_OVERLAP_GBPS", 550.0));
        }
    t.cover = std::max(0.0, num("COSMA_B200_OVERLAP_COVER", t.cover));
    if (t.force && !std::getenv("COSMA_B200_OVERLAP_COVER")) t.cover = 0.0;  // tests: the smallest panels that exercise every branch
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
    const int full = t.sms, narrow = t.zero_sm ? t.sms : std::max(1, t.sms - t.reserved_sms);
    const double t_ag = model.transfer_ms(ag_elems, t.link_gbps), t_ex = model.transfer_ms(ex_elems, t.link_gbps);
    const bool have_ag = ag[0] >= 0 || ag[1] >= 0;
    const Range all_k{0, k};
    const std::int64_t gr = std::max(1, t.col_granule);

    // Where the first panel comes from: the own column block of B inside the peer's half of C (it is needed first anyway), else inside
    // the own half. It sits at the OUTER end of its half, so that everything else stays contiguous around the boundary of the halves.
    Range R1{0, 0};
    bool s1_in_mine = false;
    if (have_ag) {
        R1 = intersect(Bo, peer);
        if (R1.empty()) {
            R1 = intersect(Bo, mine);
            s1_in_mine = true;
        }
    }
    const bool r1_low = !(R1.hi == n && R1.lo != 0);  // the region starts at column 0 (or covers everything): the first panel is its prefix

    // A layout = three widths. w1: the first panel (own pieces only, beside the allgathers). spill: columns of the own half computed
    // together with the end of the peer's half, to fill that launch's last wave. wm: the panel of the own half computed beside the
    // exchange; whatever is computed before the exchange has arrived goes to the partial buffer and is added afterwards ("early"
    // columns), the rest is computed onto the received half directly (beta = 1).
    struct Layout {
        std::int64_t w1 = 0, spill = 0, wm = 0;
    };
    std::vector<MicroOp>& P = out.ops;
    bool aligned = true;

    // emits the program of a layout (dry: only its estimated duration, penalties for exposed transfers included)
    auto build = [&](const Layout& L, bool dry) -> double {
        double est = 0.0;
        auto add = [&](const MicroOp& o) {
            P.push_back(o);
            return static_cast<int>(P.size()) - 1;
        };
        // c_override >= 0: the result goes there (element offset into the C arena) instead of into the GEMM's own result buffer
        auto gemm = [&](Range cols, Range ks, BetaMode beta, bool is_narrow, bool from_pieces, const std::vector<int>& wait, std::int64_t c_override) {
            if (cols.empty() || ks.empty()) return -1;
            est += model.gemm_ms(cols.len(), ks.len(), is_narrow ? narrow : full);
            if (dry) return -1;
            MicroOp o;
            o.kind = MicroKind::GEMM;
            o.stream = 0;
            o.wait = wait;
            o.m = static_cast<int>(m);
            o.n = static_cast<int>(cols.len());
            o.k = static_cast<int>(ks.len());
            o.beta = beta;
            o.narrow = is_narrow && !t.zero_sm;
            o.lda = m; o.ldb = k; o.ldc = m;
            if (from_pieces) {  // the gathers are still in flight: the own pieces where the caller (or an outer step) left them
                o.a_off = ag[0] >= 0 ? ops[ag[0]].src_off + (ks.lo - Ko.lo) * m : g.a_off + ks.lo * m;
                o.b_off = ag[1] >= 0 ? ops[ag[1]].src_off + (cols.lo - Bo.lo) * k + ks.lo : g.b_off + cols.lo * k + ks.lo;
            } else {
                o.a_off = g.a_off + ks.lo * m;
                o.b_off = g.b_off + cols.lo * k + ks.lo;
            }
            o.c_off = c_override >= 0 ? c_override : g.c_off + cols.lo * m;
            for (std::int64_t off : {o.a_off, o.b_off, o.c_off})
                if ((off * t.elem_bytes) % 16 != 0) aligned = false;
            return add(o);
        };

        std::vector<int> all_ag;
        if (!dry) {
            // ops before the overlapped tail run as they are, on the compute stream
            int last_serial = -1;
            for (int i = 0; i < first_async; ++i) {
                MicroOp o;
                o.kind = MicroKind::SERIAL;
                o.op = i;
                last_serial = add(o);
            }
            for (int i = first_async; i < gi; ++i) {
                MicroOp o;
                o.kind = MicroKind::ALLGATHER;
                o.stream = 1;
                o.op = i;
                if (last_serial >= 0) o.wait.push_back(last_serial);
                all_ag.push_back(add(o));
            }
        }
        // the spill: columns of the own half next to the boundary of the halves
        Range spill{0, 0};
        if (red >= 0 && L.spill > 0) spill = mine.lo == peer.hi ? Range{mine.lo, mine.lo + L.spill} : Range{mine.hi - L.spill, mine.hi};
        Range S1{0, 0};
        if (have_ag) {
            S1 = r1_low ? Range{R1.lo, R1.lo + L.w1} : Range{R1.hi - L.w1, R1.hi};
            const double g1 = model.gemm_ms(S1.len(), Ko.len(), narrow);
            est += std::max(0.0, t.cover * t_ag - g1);  // what the first panel does not hide
            gemm(S1, Ko, g.beta, true, true, {}, -1);
            if (!s1_in_mine) {
                gemm(S1, Kp, BetaMode::ONE, false, false, all_ag, -1);
                // the rest of the peer's half (without a reduce: of everything) and the spill: one range next to the boundary, plus --
                // when there is no reduce and the own block of B is the upper half -- the other half below it
                Range rest = S1.lo == peer.lo ? Range{S1.hi, peer.hi} : Range{peer.lo, S1.lo};
                if (!spill.empty()) rest = spill.lo == rest.hi ? Range{rest.lo, spill.hi} : Range{spill.lo, rest.hi};
                gemm(rest, all_k, g.beta, false, false, all_ag, -1);
            } else {
                Range pp = peer;
                if (!spill.empty()) pp = spill.lo == pp.hi ? Range{pp.lo, spill.hi} : Range{spill.lo, pp.hi};
                gemm(pp, all_k, g.beta, false, false, all_ag, -1);
            }
        } else {
            Range pp = peer;
            if (!spill.empty()) pp = spill.lo == pp.hi ? Range{pp.lo, spill.hi} : Range{spill.lo, pp.hi};
            gemm(pp, all_k, g.beta, false, false, {}, -1);
        }

        if (red >= 0) {
            const ScheduleOp& r = ops[red];
            int ex_idx = -1;
            if (!dry) {
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
                ex_idx = add(ex);
            }
            // this rank's half: [spill | ... | S1 (if it lies in this half)], the spill at the boundary, S1 at the outer end
            std::vector<Range> early;
            if (!spill.empty()) early.push_back(spill);
            double beside = 0.0;  // GEMM time beside the exchange
            Range rest = mine;
            if (!spill.empty()) rest = spill.lo == mine.lo ? Range{spill.hi, mine.hi} : Range{mine.lo, spill.lo};
            if (s1_in_mine) {
                rest = S1.lo == mine.lo ? Range{S1.hi, rest.hi} : Range{rest.lo, S1.lo};
                if (!Kp.empty()) {
                    beside += model.gemm_ms(S1.len(), Kp.len(), narrow);
                    gemm(S1, Kp, BetaMode::ONE, true, false, all_ag, -1);
                }
            }
            if (L.wm > 0 && !rest.empty()) {
                // next to the spill (or the boundary), so that the early columns there form one range
                const std::int64_t w = std::min(L.wm, rest.len());
                const bool at_low = mine.lo == peer.hi;  // the boundary is the low end of this half
                const Range M1 = at_low ? Range{rest.lo, rest.lo + w} : Range{rest.hi - w, rest.hi};
                beside += model.gemm_ms(M1.len(), k, narrow);
                gemm(M1, all_k, g.beta, true, false, all_ag, -1);
                if (!early.empty() && early.back().hi == M1.lo) early.back().hi = M1.hi;
                else if (!early.empty() && M1.hi == early.back().lo) early.back().lo = M1.lo;
                else early.push_back(M1);
                rest = at_low ? Range{M1.hi, rest.hi} : Range{rest.lo, M1.lo};
            }
            if (s1_in_mine) early.push_back(S1);
            est += std::max(0.0, t.cover * t_ex - beside);  // what the panels beside the exchange do not hide
            // the received half: C = beta * C + received (skipped at run time when beta == 0: the exchange then lands in C itself)
            MicroOp acc;
            acc.kind = MicroKind::ACCUMULATE;
            acc.stream = 0;
            acc.wait.push_back(ex_idx);
            if (r.beta != BetaMode::ZERO) {
                acc.dst_off = r.dst_off;
                acc.add_off = r.tmp_off;
                acc.count = ex_elems;
                acc.beta = r.beta;
                acc.beta_term = true;
                if (!dry) add(acc);
                est += model.transfer_ms(3 * ex_elems, 6000.0);
            }
            // ... += the early columns of the own partial result
            acc.beta = BetaMode::ONE;
            acc.beta_term = false;
            for (const Range& e : early) {
                acc.dst_off = r.dst_off + (e.lo - mine.lo) * m;
                acc.add_off = r.src_off + e.lo * m;
                acc.count = e.len() * m;
                if (!dry) add(acc);
                est += model.transfer_ms(3 * acc.count, 6000.0) + 0.005;
            }
            // ... and the remaining columns computed onto it
            if (!rest.empty()) {
                const int gi2 = gemm(rest, all_k, BetaMode::ONE, false, false, all_ag, r.dst_off + (rest.lo - mine.lo) * m);
                if (gi2 >= 0) P[gi2].wait.push_back(ex_idx);
            }
        }
        if (!dry)
            for (int i = (red >= 0 ? red + 1 : gi + 1); i < static_cast<int>(ops.size()); ++i) {
                MicroOp o;
                o.kind = MicroKind::SERIAL;
                o.op = i;
                add(o);
            }
        return est;
    };

    // choose the layout: whole waves everywhere, transfers hidden (the estimate carries a penalty for what is not)
    Layout best;
    {
        const std::int64_t n1 = have_ag ? std::max<std::int64_t>(R1.len() / gr, 1) : 0;
        const std::int64_t mine_free = red >= 0 ? mine.len() : 0;
        if (t.force && gr < 128) {
            // tests: small fixed panels that exercise every branch
            best.w1 = have_ag ? std::min<std::int64_t>(gr, R1.len()) : 0;
            const std::int64_t left = mine_free - (s1_in_mine ? best.w1 : 0);
            best.spill = std::min<std::int64_t>(gr, std::max<std::int64_t>(left - gr, 0));
            best.wm = std::min<std::int64_t>(gr, std::max<std::int64_t>(left - best.spill, 0));
        } else {
            // the round granule: the fewest column tiles that make whole waves of the full grid
            std::int64_t c0 = 1;
            while (c0 < 64 && (model.tiles_m() * c0) % full != 0) ++c0;
            const std::int64_t max_spill = std::min<std::int64_t>(c0 + 1, 64);
            double best_cost = -1.0;
            for (std::int64_t c1 = have_ag ? 1 : 0; c1 <= n1; ++c1) {
                Layout L;
                L.w1 = have_ag ? (c1 == n1 ? R1.len() : c1 * gr) : 0;
                const std::int64_t left = mine_free - (s1_in_mine ? L.w1 : 0);
                for (std::int64_t e = 0; e <= max_spill && e * gr <= left; ++e) {
                    L.spill = e * gr;
                    L.wm = 0;
                    const double c = build(L, true);
                    if (best_cost < 0.0 || c < best_cost - 1e-9) { best_cost = c; best = L; }
                    if (red < 0) break;
                }
                if (!have_ag) break;
            }
            if (red >= 0) {
                const std::int64_t left = mine_free - (s1_in_mine ? best.w1 : 0) - best.spill;
                Layout L = best;
                for (std::int64_t c = 0; c * gr <= left; ++c) {
                    L.wm = c * gr;
                    const double cst = build(L, true);
                    if (cst < best_cost - 1e-9) { best_cost = cst; best = L; }
                }
            }
        }
    }
    P.clear();
    aligned = true;
    const double est = build(best, false);

    out.est_comm_ms = t_ag + t_ex;
    out.est_serial_ms = model.gemm_ms(n, k, full) + model.transfer_ms(ag_elems + ex_elems, t.serial_gbps) +
                        (red >= 0 && ops[red].beta != BetaMode::ZERO ? model.transfer_ms(3 * ex_elems, 6000.0) : 0.0);
    out.est_overlap_ms = est;
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
