"""Starts N ranks of a C++ host program on one node, torchrun-style environment, for boxes without mpirun:

    python -m cosma_b200.launch -np 4 ./cosma_miniapp -m 8192 -n 8192 -k 8192

Each rank gets RANK, WORLD_SIZE, LOCAL_RANK (= the GPU it drives), MASTER_ADDR=127.0.0.1 and COSMA_B200_PG_PORT (a free
port for cosma::pg's rendezvous, include/cosma/process_group.hpp). With a real MPI the same programs start under mpirun."""
import argparse
import os
import signal
import socket
import subprocess
import sys
import time


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _child_setup():
    """Own session (so the whole rank can be killed as a group) and SIGKILL when the launcher itself dies (PR_SET_PDEATHSIG): a rank must
    never outlive the test run that started it and sit on a GPU."""
    os.setsid()
    try:
        import ctypes
        ctypes.CDLL(None).prctl(1, signal.SIGKILL)
    except Exception:
        pass


def launch(np_, argv, timeout=None, env_extra=None, capture=False):
    """Returns (exit code, [stdout of each rank] if capture). Any failing rank takes the others down."""
    port = free_port()
    procs = []
    for r in range(np_):
        env = dict(os.environ)
        env.update({"RANK": str(r), "WORLD_SIZE": str(np_), "LOCAL_RANK": str(r), "MASTER_ADDR": "127.0.0.1",
                    "COSMA_B200_PG_PORT": str(port)})
        if timeout:  # a rank that hears nothing for longer than the whole job may take gives up by itself (orphan protection)
            env.setdefault("COSMA_B200_PG_RECV_TIMEOUT", str(int(timeout) + 30))
        if env_extra:
            env.update(env_extra)
        procs.append(subprocess.Popen(argv, env=env, stdout=subprocess.PIPE if capture else None,
                                      stderr=subprocess.STDOUT if capture else None, preexec_fn=_child_setup))
    t0 = time.time()
    code = 0
    live = set(range(np_))
    try:
        while live:
            for r in list(live):
                rc = procs[r].poll()
                if rc is not None:
                    live.discard(r)
                    if rc != 0 and code == 0:
                        code = rc
            if code != 0 or (timeout and time.time() - t0 > timeout):
                if code == 0:
                    code = 124
                break
            time.sleep(0.02)
    finally:
        for r in live:
            try:
                os.killpg(procs[r].pid, signal.SIGKILL)
            except ProcessLookupError:
                pass
    outs = []
    for p in procs:
        if capture:
            try:
                outs.append(p.communicate(timeout=5)[0].decode(errors="replace"))
            except Exception:
                outs.append("")
        else:
            p.wait()
    return code, outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-np", type=int, default=1)
    ap.add_argument("--timeout", type=float, default=None)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    if not a.cmd:
        ap.error("no program given")
    code, _ = launch(a.np, a.cmd, timeout=a.timeout)
    return code


if __name__ == "__main__":
    sys.exit(main())
