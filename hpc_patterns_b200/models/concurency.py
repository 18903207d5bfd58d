"""Python mirror of the concurrency benchmark (``bin/concurency``).

``bench()`` has the contract of the reference's ``bench<T>()`` (concurency/bench.hpp:37-40):
it returns ``(total_us, per_command_us)`` with the per-command list filled only in serial
mode.  ``run_cli()`` runs the native driver in-process and returns its exit status and
output; ``sweep()`` is the Python form of ``run_sycl.sh`` / ``run_omp.sh``.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

from .. import native
from ..utils.parse import parse_log, render

REFERENCE_GROUPS: List[List[str]] = [["C", "C"], ["C", "M2D"], ["C", "D2M"], ["M2D", "D2M"], ["H2D", "D2H"]]
B200_GROUPS: List[List[str]] = [["C", "D2P"], ["D2P", "P2D"], ["A", "H2D"], ["A", "D2P"]]
CUDA_MODES = ("in_order", "out_of_order", "host_threads", "nowait", "fused")
CPU_MODES = ("host_threads", "nowait")


def sanitize_command(token: str) -> str:
    """``M2D`` -> ``MD`` (every '2' is dropped, concurency/main.cpp:14-19)."""
    return native().strip_twos(token)


def bench(mode: str, commands: Sequence[str], params: Dict[str, int], *, backend: str = "auto",
          enable_profiling: bool = False, n_queues: int = -1, n_repetitions: int = 10,
          verbose: bool = False) -> Tuple[int, List[int]]:
    # Parameter keys name commands too (globalsize_M2D); sanitise them like the command tokens.
    clean = {}
    for k, v in params.items():
        head, sep, cmd = k.partition("_")
        clean[head + sep + sanitize_command(cmd) if head == "globalsize" else k] = int(v)
    r = native().concurency_bench(backend, mode, [sanitize_command(c) for c in commands], clean, enable_profiling,
                                  n_queues,
                                  n_repetitions, verbose)
    return int(r["total_us"]), [int(x) for x in r["per_command_us"]]


def run_cli(argv: Sequence[str], backend: str = "auto") -> Tuple[int, str, str]:
    """Run ``concurency <argv...>`` in-process; returns (exit status, stdout, stderr)."""
    rc, out, err = native().concurency_main(list(argv), backend)
    return int(rc), out, err


def sweep(modes: Iterable[str], groups: Optional[Sequence[Sequence[str]]] = None, backend: str = "auto",
          envs: Sequence[Dict[str, str]] = ({},), extra_args: Sequence[str] = ()) -> str:
    """Run modes x groups under each environment; returns the rendered SUCCESS/FAILURE tables."""
    groups = REFERENCE_GROUPS if groups is None else groups
    log: List[str] = []
    for env in envs:
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        log.append("export " + (" ".join(f"{k}={v}" for k, v in env.items()) or "DEFAULT=1"))
        try:
            for mode in modes:
                argv = [mode, *extra_args]
                for g in groups:
                    argv += ["--commands", *g]
                _, out, err = run_cli(argv, backend)
                log.append(out)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return render(parse_log("\n".join(log)))


def main(argv: Optional[List[str]] = None) -> int:
    import sys

    argv = list(sys.argv[1:] if argv is None else argv)
    backend = "auto"
    if "--backend" in argv:
        i = argv.index("--backend")
        backend = argv[i + 1]
        del argv[i:i + 2]
    rc, out, err = run_cli(argv, backend)
    sys.stdout.write(out)
    sys.stderr.write(err)
    return rc


if __name__ == "__main__":
    raise SystemExit(main())
