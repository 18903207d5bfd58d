"""Turn a concurrency-benchmark log into SUCCESS/FAILURE tables.

Same job as the reference's ``concurency/parse.py:12-32``: the sweep scripts write one
log that interleaves ``export VAR=...`` lines (the environment of the following runs) and
the benchmark's verdict lines ``## <mode> | <cmd> <cmd> ... | SUCCESS|FAILURE: ...``;
this prints one table per environment with a row per command group and a column per
mode.  Usage: ``python -m hpc_patterns_b200.utils.parse <log> [tablefmt]``.
"""
from __future__ import annotations

import re
import sys
from collections import OrderedDict
from typing import Dict, List, Optional

Tables = "OrderedDict[str, OrderedDict[str, OrderedDict[str, str]]]"

_VERDICT = re.compile(r"^\s*##\s*(?P<mode>[^|]+)\|(?P<cmds>[^|]*)\|\s*(?P<res>SUCCESS|FAILURE)")
_EXPORT = re.compile(r"\bexport\s+(?P<env>.+?)\s*$")


def parse_log(text: str):
    """-> {environment: {command group: {mode: 'SUCCESS'|'FAILURE'}}} in order of appearance."""
    tables: "OrderedDict[Optional[str], OrderedDict[str, OrderedDict[str, str]]]" = OrderedDict()
    env: Optional[str] = None
    for line in text.splitlines():
        m = _VERDICT.match(line)
        if m:
            group = " ".join(m.group("cmds").split())
            tables.setdefault(env, OrderedDict()).setdefault(group, OrderedDict())[m.group("mode").strip()] = \
                m.group("res")
            continue
        e = _EXPORT.search(line)
        if e and "##" not in line:
            env = e.group("env")
    return tables


def _simple_table(rows: List[Dict[str, str]], headers: List[str]) -> str:
    widths = [max(len(h), *(len(str(r.get(h, ""))) for r in rows)) for h in headers]
    fmt = "  ".join("{:<%d}" % w for w in widths)
    lines = [fmt.format(*headers), fmt.format(*("-" * w for w in widths))]
    lines += [fmt.format(*(str(r.get(h, "")) for h in headers)) for r in rows]
    return "\n".join(lines)


def render(tables, tablefmt: str = "simple") -> str:
    out: List[str] = []
    for env, groups in tables.items():
        rows = [{"commands": g, **modes} for g, modes in groups.items()]
        headers: List[str] = ["commands"]
        for r in rows:
            for k in r:
                if k not in headers:
                    headers.append(k)
        out.append(str(env))
        try:
            from tabulate import tabulate

            out.append(tabulate(rows, headers="keys", tablefmt=tablefmt))
        except ImportError:  # pragma: no cover
            out.append(_simple_table(rows, headers))
        out.append("")
    return "\n".join(out)


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: parse <logfile> [tablefmt]", file=sys.stderr)
        return 2
    with open(argv[0]) as f:
        text = f.read()
    out = render(parse_log(text), argv[1] if len(argv) > 1 else "simple")
    if out:  # a log without verdict lines prints nothing at all (like the reference's parser)
        print(out)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
