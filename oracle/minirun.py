#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- launcher for the minimpi stand-in (oracle/stubs/minimpi.cpp).

    python oracle/minirun.py -np R [--threads T] prog args...

Starts R copies of `prog` as ranks 0..R-1 of one MPI_COMM_WORLD: one unix socketpair per pair of ranks, handed to the
children through MINIMPI_RANK / MINIMPI_SIZE / MINIMPI_FDS. The exit code is the first non-zero child exit code. Used
by tests/ and by bench.py's reference arm to run the UNMODIFIED reference COSMA (oracle/_ref) on several ranks.
"""
import os
import socket
import subprocess
import sys


def launch(nranks, argv, threads=None, env=None, stdout=None, timeout=None):
    """Runs argv on nranks ranks; returns (exit code, [stdout of each rank] or None)."""
    pairs = {}
    for i in range(nranks):
        for j in range(i + 1, nranks):
            pairs[(i, j)] = socket.socketpair(socket.AF_UNIX, socket.SOCK_STREAM)
    procs = []
    for r in range(nranks):
        fds = []
        for p in range(nranks):
            if p == r:
                fds.append(-1)
            elif r < p:
                fds.append(pairs[(r, p)][0].fileno())
            else:
                fds.append(pairs[(p, r)][1].fileno())
        e = dict(os.environ if env is None else env)
        e.update(MINIMPI_RANK=str(r), MINIMPI_SIZE=str(nranks), MINIMPI_FDS=",".join(map(str, fds)))
        if threads is not None:
            e.update(OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS=str(threads))
        procs.append(subprocess.Popen(argv, env=e, pass_fds=[f for f in fds if f >= 0], stdout=stdout))
    for a, b in pairs.values():
        a.close()
        b.close()
    code, outs = 0, []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
        if p.returncode != 0 and code == 0:
            code = p.returncode
    return code, (outs if stdout is not None else None)


def main():
    args = sys.argv[1:]
    nranks, threads = 1, None
    while args and args[0].startswith("-"):
        if args[0] == "-np":
            nranks = int(args[1])
        elif args[0] == "--threads":
            threads = int(args[1])
        else:
            break
        args = args[2:]
    if not args:
        print(__doc__)
        return 2
    code, _ = launch(nranks, args, threads)
    return code if code >= 0 else 128 - code


if __name__ == "__main__":
    sys.exit(main())
