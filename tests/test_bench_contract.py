"""bench.py's reference arm (the one leg that runs without a GPU): one JSON line with the contract's keys, at N = 1 (the reference's
cosma::multiply in-process) and at N = 2 (the unmodified reference miniapp on 2 minimpi ranks)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "cpu_baseline", "e2e", "gpu_launches"}


@pytest.mark.parametrize("gpus", [1, 2])
def test_reference_arm_line(ref, gpus):
    env = dict(os.environ, RANK="0", WORLD_SIZE=str(gpus))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus), "--steps", "1", "--warmup", "1",
                          "--mnk", "768,640,512"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert KEYS <= set(d), sorted(KEYS - set(d))
    assert d["impl"] == "reference" and d["n_gpus"] == gpus and d["unit"] == "TFLOP/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the reference arm runs on OUR arm's config: the very dict bench.workload_config gives our line (the sample it times is described
    # under cpu_baseline.sample); the strategy is the reference's own Strategy(m, n, k, P), which the planning tests pin equal to ours
    import bench
    assert d["config"] == bench.workload_config("dgemm m=768 n=640 k=512 (override)", 768, 640, 512, {1: "", 2: "pm2"}[gpus], "d"), d["config"]


def test_other_ranks_of_the_reference_arm_do_nothing(ref):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_host_affinity_helper_is_best_effort():
    """cosma_b200.affinity: cpulist parsing; without a GPU / sysfs entry nothing is bound and nothing raises."""
    import os
    from cosma_b200 import affinity
    assert affinity._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and affinity._cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = affinity.bind_to_gpu(0)
    assert set(info) == {"numa_node", "cpus", "bound"}
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


WATCHDOG_SCRIPT = r"""
import json, os, sys, time
sys.path.insert(0, %r)
import bench
dog = bench.Watchdog(int(os.environ["RANK"]))
mode = sys.argv[1]
if mode == "measured":      # the headline measurement is in, an extra never returns
    dog.line = {"metric": bench.METRIC, "value": 1.5, "unit": "TFLOP/s"}
    dog.stage = "also.pzgemm"
elif mode == "printed":     # the line is out, the teardown never returns
    print(json.dumps({"value": 2.5}), flush=True)
    dog.printed = True
    dog.stage = "teardown"
time.sleep(60)
"""


@pytest.mark.parametrize("mode,rank,rc,lines", [("measured", 0, 0, 1), ("nothing", 0, 3, 1), ("printed", 0, 0, 1), ("measured", 1, 0, 0)])
def test_bench_watchdog(mode, rank, rc, lines):
    """bench.Watchdog: a stage that never returns costs the extras, not the line -- rank 0 prints what was measured (exit 0), or an
    explicit value-less line when nothing was (exit 3); after the line is out it leaves quietly; other ranks never print."""
    env = dict(os.environ, RANK=str(rank), COSMA_B200_BENCH_DEADLINE_S="1")
    out = subprocess.run([sys.executable, "-c", WATCHDOG_SCRIPT % ROOT, mode], capture_output=True, text=True, timeout=50, env=env, cwd=ROOT)
    assert out.returncode == rc, (out.returncode, out.stderr[-1000:])
    got = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(got) == lines, out.stdout
    if mode == "measured" and rank == 0:
        assert got[0]["value"] == 1.5 and "also.pzgemm" in got[0]["incomplete"]
    if mode == "nothing":
        assert got[0]["value"] is None and "start-up" in got[0]["incomplete"]
    if mode == "printed":
        assert got[0] == {"value": 2.5}


def test_our_arm_assembles_the_line_stage_by_stage(monkeypatch, capsys):
    """bench.main() of our arm with the GPU work stubbed out: the line exists (for the watchdog) as soon as the device-resident
    measurement is in, gains e2e / parity / also as they complete, and is printed once with the contract's keys."""
    import bench
    from cosma_b200 import _lib
    _lib.load()
    seen = []

    class FakeEnv:
        world, rank, local_rank, dev, affinity, dist = 1, 0, 0, None, None, None

        def barrier(self):
            pass

        def max_over_ranks(self, x):
            return x

    def fake_run_multiply(env, m, n, k, dtype, steps, warmup, strategy="", with_e2e=True, with_parity=True, sample_clocks=False, comm=None,
                          on_measured=None, stage=None):
        out = {"strategy": "", "ms_per_step": 2.0, "value": 2.0 * m * n * k / 2e-3 * 1e-12, "launches": steps,
               "roofline": {"bound": "tensor", "achieved": 1.0, "peak": 2.0, "unit": "TFLOP/s", "frac": 0.5, "traffic": None}}
        if on_measured:
            on_measured(out)
            seen.append(dict(dog_line()))
        if with_e2e:
            out["e2e"] = {"value": 1.0, "unit": "TFLOP/s", "h2d_bytes_per_step": 8 * (m * k + k * n), "d2h_bytes_per_step": 8 * m * n}
        if with_parity:
            out["parity"] = {"ok": True, "exact": True}
        return out

    dogs = []
    real_dog = bench.Watchdog

    def make_dog(rank, reps=0):
        os.environ["COSMA_B200_BENCH_DEADLINE_S"] = "0"  # no timer in the test process
        d = real_dog(rank, reps)
        dogs.append(d)
        return d

    def dog_line():
        return dogs[0].line

    monkeypatch.setattr(bench, "Env", FakeEnv)
    monkeypatch.setattr(bench, "run_multiply", fake_run_multiply)
    monkeypatch.setattr(bench, "Watchdog", make_dog)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "1", "--steps", "4", "--warmup", "3", "--no-cpu-baseline"])
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    try:
        assert bench.main() == 0
    finally:
        os.environ.pop("COSMA_B200_BENCH_DEADLINE_S", None)
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    need = (KEYS - {"impl"}) | {"roofline", "parity", "clocks", "also"}
    assert need <= set(d), sorted(need - set(d))
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["config"]["m"] == 32768 and d["gpu_launches"] == 4
    assert d["config"] == bench.workload_config(bench.WORKLOADS["cfg3"][4] % 1, 32768, 32768, 32768, "", "d") and d["frac_of_peak"] > 0 and "peak_per_gpu_tflops" in d
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["parity"]["ok"] and d["also"]["cfg2"]["e2e"]["value"] == 1.0
    # what the watchdog would have printed had the e2e stage of the headline workload never returned
    assert seen[0]["value"] == d["value"] and seen[0]["e2e"] is None and "also" not in seen[0]
    assert dogs[0].printed and dogs[0].stage == "teardown"
