"""Line coverage of the Python layer without third-party packages.

`pytest-cov` / `coverage` are not part of the offline image, so `run_test.sh` falls back to
this pytest plugin (`-p tools.pycov --pycov`): it records executed lines of the files under
`infinistore_b200/` with `sys.monitoring` (CPython >= 3.12, near-zero overhead: every line
event disables itself after the first hit) and prints a per-file table plus the total at the
end of the session.  Executable lines come from the compiled code objects (`co_lines`), the
same definition coverage.py uses.  Reference counterpart: run_test.sh `--cov=infinistore`.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "infinistore_b200") + os.sep
_hits: dict[str, set[int]] = {}
_TOOL = 3  # sys.monitoring tool id reserved for coverage tools


def _executable_lines(path: str) -> set[int]:
    with open(path, "rb") as f:
        src = f.read()
    out: set[int] = set()

    def walk(code: types.CodeType):
        for _, _, line in code.co_lines():
            if line:
                out.add(line)
        for c in code.co_consts:
            if isinstance(c, types.CodeType):
                walk(c)

    try:
        walk(compile(src, path, "exec"))
    except SyntaxError:
        pass
    # docstrings / bare string constants show up as lines of their code object: drop them
    return out


def start():
    mon = getattr(sys, "monitoring", None)
    if mon is None:
        return False
    try:
        mon.use_tool_id(_TOOL, "istore-pycov")
    except ValueError:
        return False

    pkg, hits, disable = PKG, _hits, mon.DISABLE  # locals: module globals die at shutdown

    def on_line(code: types.CodeType, line: int):
        try:
            fn = code.co_filename
            if fn.startswith(pkg):
                hits.setdefault(fn, set()).add(line)
        except Exception:  # noqa: BLE001 - never let the tracer break the traced program
            pass
        return disable  # one hit per line is all we need

    mon.register_callback(_TOOL, mon.events.LINE, on_line)
    mon.set_events(_TOOL, mon.events.LINE)
    return True


def report(out=sys.stdout, fail_under: float = 0.0) -> float:
    rows = []
    total_exec = total_hit = 0
    for dirpath, _, files in os.walk(PKG):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            path = os.path.join(dirpath, f)
            lines = _executable_lines(path)
            hit = _hits.get(path, set()) & lines
            total_exec += len(lines)
            total_hit += len(hit)
            rows.append((os.path.relpath(path, ROOT), len(lines), len(hit)))
    out.write("\n---------- coverage: infinistore_b200 (tools/pycov.py) ----------\n")
    for name, n, h in sorted(rows):
        pct = 100.0 * h / n if n else 100.0
        out.write(f"{name:48s} {n:5d} {n - h:5d} {pct:5.0f}%\n")
    pct = 100.0 * total_hit / total_exec if total_exec else 100.0
    out.write(f"{'TOTAL':48s} {total_exec:5d} {total_exec - total_hit:5d} {pct:5.0f}%\n")
    if fail_under and pct < fail_under:
        out.write(f"coverage {pct:.1f}% is below --pycov-fail-under={fail_under}\n")
    return pct


# ---- pytest plugin hooks
def pytest_addoption(parser):
    g = parser.getgroup("pycov")
    g.addoption("--pycov", action="store_true", help="line coverage of infinistore_b200/")
    g.addoption("--pycov-fail-under", type=float, default=0.0)


def pytest_configure(config):
    if config.getoption("--pycov"):
        config._pycov_on = start()


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    if getattr(config, "_pycov_on", False):
        import io

        buf = io.StringIO()
        pct = report(buf, config.getoption("--pycov-fail-under"))
        terminalreporter.write(buf.getvalue())
        config._pycov_pct = pct


def pytest_sessionfinish(session, exitstatus):
    cfg = session.config
    under = cfg.getoption("--pycov-fail-under")
    if getattr(cfg, "_pycov_on", False) and under:
        import io

        if report(io.StringIO()) < under and session.exitstatus == 0:
            session.exitstatus = 1
