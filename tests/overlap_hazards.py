"""TEST INFRASTRUCTURE: static hazard analysis of the overlapped micro-op programs (include/cosma/overlap.hpp).

tests/schedule_sim.py::run_overlapped interprets a program in PROGRAM order -- one valid serialisation of its two streams. The executor
(csrc/multiply_exec.cu::plan_run_overlapped) runs the two streams concurrently and orders them only through the `wait` lists, and with
the copy-engine transport the ring mate writes into this rank's landing zones whenever IT is ready. This module checks, for the very
programs the library emits (any size: nothing is computed), that every pair of conflicting accesses is ordered:

  * happens-before = program order inside a stream + the `wait` edges (cudaStreamWaitEvent on the other stream's completion event),
    transitively;
  * every micro-op has a read set, a write set and -- for communication ops -- a landing zone written by the mate;
  * two micro-ops of one rank conflict when one writes (or lets the mate write) what the other reads or writes; a conflict must be
    ordered one way or the other;
  * NCCL transport: the mate's data lands while the op runs. Copy-engine transport: the mate pushes as soon as it has seen this rank's
    ENTERED flag, which is raised at the START of the call (peer_transport.cu) -- so a landing zone is in flight from call entry until the
    op's ARRIVED wait, and everything that touches it must come AFTER the op.

Reference for the semantics being protected: the one-sided windows of src/cosma/one_sided_communicator.cpp:417-656, 776-1016 (what may be
read or overwritten while a transfer is in flight)."""
import numpy as np


def _intervals(off, rows, cols, ld):
    """[start, stop) element intervals of a column-major rows x cols sub-matrix at `off` with leading dimension ld (merged when dense)."""
    if rows <= 0 or cols <= 0:
        return np.zeros((0, 2), dtype=np.int64)
    if rows == ld or cols == 1:
        return np.array([[off, off + rows + ld * (cols - 1)]], dtype=np.int64)
    starts = off + ld * np.arange(cols, dtype=np.int64)
    return np.stack([starts, starts + rows], axis=1)


def _flat(off, count):
    return np.array([[off, off + count]], dtype=np.int64) if count > 0 else np.zeros((0, 2), dtype=np.int64)


def _merge(iv):
    """Sorted, disjoint form of an interval list."""
    iv = iv[np.argsort(iv[:, 0], kind="stable")]
    reach = np.maximum.accumulate(iv[:, 1])
    first = np.ones(len(iv), dtype=bool)
    first[1:] = iv[1:, 0] > reach[:-1]          # a gap before this interval: a new run starts
    starts = iv[first, 0]
    stops = reach[np.append(np.nonzero(first)[0][1:] - 1, len(iv) - 1)]
    return np.stack([starts, stops], axis=1)


def _overlap(a, b):
    """Do two interval lists (each sorted by start, disjoint) share an element?"""
    if len(a) == 0 or len(b) == 0:
        return False
    if len(a) > len(b):
        a, b = b, a
    idx = np.searchsorted(b[:, 1], a[:, 0], side="right")  # first interval of b that ends after a starts
    ok = idx < len(b)
    return bool(np.any(b[idx[ok], 0] < a[ok, 1]))


class Access:
    """Footprints of one micro-op: reads / writes / zone, each {arena index: interval list}."""

    def __init__(self):
        self.reads, self.writes, self.zone = {}, {}, {}

    @staticmethod
    def _add(d, arena, iv):
        if len(iv):
            d[arena] = _merge(iv if arena not in d else np.concatenate([d[arena], iv]))

    def read(self, arena, iv):
        self._add(self.reads, arena, iv)

    def write(self, arena, iv):
        self._add(self.writes, arena, iv)

    def lands(self, arena, iv):
        self._add(self.zone, arena, iv)


def _hits(x, y):
    return any(a in y and _overlap(x[a], y[a]) for a in x)


def accesses(micro, sched, beta_zero):
    """Access sets of every micro-op of one rank's program. beta_zero: the user's beta is 0 at run time (BetaMode USER ops then neither
    read C nor stage the received half)."""
    def reads_c(mode):
        return mode == 1 or (mode == 2 and not beta_zero)
    out = []
    for o in micro:
        acc = Access()
        kind = o["kind"]
        if kind == "gemm":
            acc.read(0, _intervals(o["a_off"], o["m"], o["k"], o["lda"]))
            acc.read(1, _intervals(o["b_off"], o["k"], o["n"], o["ldb"]))
            c = _intervals(o["c_off"], o["m"], o["n"], o["ldc"])
            if reads_c(o["beta"]):
                acc.read(2, c)
            acc.write(2, c)
        elif kind == "exchange":
            acc.read(2, _flat(o["send_off"], o["count"]))
            acc.lands(2, _flat(o["recv_off"] if reads_c(o["beta"]) else o["recv_off_zero"], o["count"]))
        elif kind == "accumulate":
            if not (o["beta_term"] and not reads_c(o["beta"])):  # skipped at run time: the exchange landed in C itself
                acc.read(2, _flat(o["add_off"], o["count"]))
                if reads_c(o["beta"]):
                    acc.read(2, _flat(o["dst_off"], o["count"]))
                acc.write(2, _flat(o["dst_off"], o["count"]))
        else:  # allgather / serial: a schedule op
            s = sched[o["op"]]
            x, pos, div = s["matrix"], s["my_pos"], len(s["ring"])
            mine = sum(s["piece"][pos])
            total = sum(sum(p) for p in s["piece"])
            if s["kind"] == "allgather":
                acc.read(x, _flat(s["src_off"], mine))
                if kind == "allgather" and div == 2 and s["regular"]:
                    cnt = s["piece"][0][0]
                    acc.write(x, _flat(s["dst_off"] + pos * cnt, cnt))          # own piece into its slot
                    acc.lands(x, _flat(s["dst_off"] + (1 - pos) * cnt, cnt))     # the mate's piece
                else:
                    acc.lands(x, _flat(s["dst_off"], total))
            else:  # reduce
                acc.read(x, _flat(s["src_off"], total))
                if reads_c(s["beta"]):
                    acc.lands(x, _flat(s["tmp_off"], mine))
                    acc.read(x, _flat(s["dst_off"], mine))
                    acc.write(x, _flat(s["dst_off"], mine))
                else:
                    acc.lands(x, _flat(s["dst_off"], mine))
        out.append(acc)
    return out


def happens_before(micro):
    """hb[i, j] = micro-op i has completed before micro-op j starts, by stream order and wait edges (transitive closure)."""
    n = len(micro)
    hb = np.zeros((n, n), dtype=bool)
    last = {}
    for j, o in enumerate(micro):
        preds = list(o["wait"])
        if o["stream"] in last:
            preds.append(last[o["stream"]])
        for p in preds:
            assert 0 <= p < j, "a micro-op depends on a later one"
            hb[p, j] = True
            hb[:, j] |= hb[:, p]
        last[o["stream"]] = j
    return hb


def hazards(micro, sched, beta_zero, copy_engine):
    """-> list of human-readable hazards of one rank's program ([] = every conflicting pair is ordered)."""
    acc = accesses(micro, sched, beta_zero)
    hb = happens_before(micro)
    found = []

    def name(i):
        o = micro[i]
        return "#%d %s(stream %d)" % (i, o["kind"], o["stream"])
    n = len(micro)
    for i in range(n):
        for j in range(i + 1, n):
            a, b = acc[i], acc[j]
            plain = (_hits(a.writes, b.reads) or _hits(a.writes, b.writes) or _hits(a.reads, b.writes))
            zone_i = _hits(a.zone, b.reads) or _hits(a.zone, b.writes) or _hits(a.zone, b.zone)
            zone_j = _hits(b.zone, a.reads) or _hits(b.zone, a.writes)
            if plain and not (hb[i, j] or hb[j, i]):
                found.append("%s and %s touch the same memory and are not ordered" % (name(i), name(j)))
            # the copy-engine rule holds for the ops the executor hands to the peer transport; SERIAL ops (larger rings, outer reduces)
            # stay NCCL kernels, whose data lands while the op runs
            ce_i = copy_engine and micro[i]["stream"] == 1 and micro[i]["kind"] in ("allgather", "exchange")
            ce_j = copy_engine and micro[j]["stream"] == 1 and micro[j]["kind"] in ("allgather", "exchange")
            if zone_i:
                # j touches what the mate writes for i
                if ce_i:
                    if not hb[i, j]:
                        found.append("%s touches the landing zone of %s before its arrival has been waited for" % (name(j), name(i)))
                elif not (hb[i, j] or hb[j, i]):
                    found.append("%s touches the landing zone of %s while it may be in flight" % (name(j), name(i)))
            if zone_j:
                # i (the earlier op in program order) touches what the mate writes for j
                if ce_j:
                    found.append("%s touches the landing zone of %s, which is in flight from the start of the call" % (name(i), name(j)))
                elif not (hb[i, j] or hb[j, i]):
                    found.append("%s touches the landing zone of %s while it may be in flight" % (name(i), name(j)))
    return found


def final_order(micro):
    """The executor makes the caller's stream wait for the LAST communication op only (plan_run_overlapped): every other communication
    op must happen before it or before an op of the compute stream. -> list of micro-op indices nothing waits for."""
    hb = happens_before(micro)
    comm = [i for i, o in enumerate(micro) if o["stream"] == 1]
    compute = [i for i, o in enumerate(micro) if o["stream"] == 0]
    if not comm:
        return []
    tail = comm[-1]
    return [i for i in comm[:-1] if not (hb[i, tail] or any(hb[i, c] for c in compute))]
