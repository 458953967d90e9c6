#!/usr/bin/env python
"""Run the reference's OWN test-suite (staged by tools/make_ref.py under oracle/_ref/suite/) against one of
three arrangements of ``import lshrs`` and print one JSON line with the outcome.

    python tests/ref_suite_runner.py --variant reference            # the unmodified reference (harness check)
    python tests/ref_suite_runner.py --variant two_import           # reference LSHRS, B200 hasher + rerank
    python tests/ref_suite_runner.py --variant dropin               # lshrs_b200.compat: B200 everything on the path
    python tests/ref_suite_runner.py --variant dropin --cpu-double  # host logic only, oracle-backed C-ABI double

Used by tests/test_reference_suite.py (subprocess).  ``--cpu-double`` swaps liblshx for tests/fake_lshx.py
so the host layer can be checked without a GPU; without it the real library runs and a run that launched
no kernel is reported as such.
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
REF = REPO / "oracle" / "_ref"


class _Tally:
    def __init__(self):
        self.passed = self.failed = self.skipped = self.errors = 0
        self.failures: list[str] = []

    def pytest_runtest_logreport(self, report):
        if report.when == "call":
            if report.passed:
                self.passed += 1
            elif report.failed:
                self.failed += 1
                self.failures.append(report.nodeid)
            elif report.skipped:
                self.skipped += 1
        elif report.failed:
            self.errors += 1
            self.failures.append(f"{report.nodeid} ({report.when})")
        elif report.skipped and report.when == "setup":
            self.skipped += 1


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", choices=("reference", "two_import", "dropin"), required=True)
    ap.add_argument("--cpu-double", action="store_true")
    ap.add_argument("-k", default=None)
    args = ap.parse_args()
    if not (REF / "suite" / "tests").is_dir():
        print(json.dumps({"staged": False}))
        return 3
    # sys.path[0] is tests/ (the script's directory): drop it so that `tests` resolves to the staged suite
    sys.path[:] = [p for p in sys.path if Path(p or ".").resolve() != (REPO / "tests").resolve()]
    front = {
        "reference": [REF / "reference"],
        "two_import": [REF / "two_import", REPO],
        "dropin": [REPO / "lshrs_b200" / "compat", REF / "reference", REPO],
    }[args.variant]
    sys.path[:0] = [str(p) for p in [*front, REF / "stubs", REF / "suite"]]
    fake = None
    if args.cpu_double:
        sys.path.append(str(REPO / "tests"))
        import fake_lshx

        fake = fake_lshx.install()
    import pytest

    tally = _Tally()
    argv = [str(REF / "suite" / "tests"), "-q", "-p", "no:cacheprovider", "--rootdir", str(REF / "suite"),
            "-o", "addopts=", "-W", "ignore"]
    if args.k:
        argv += ["-k", args.k]
    rc = pytest.main(argv, plugins=[tally])
    import lshrs

    launches = None
    if args.variant != "reference":
        from lshrs_b200 import _native

        launches = fake.launches if fake is not None else (_native.launch_count() if _native._lib is not None else 0)
    print(json.dumps({"staged": True, "variant": args.variant, "cpu_double": bool(args.cpu_double), "rc": int(rc),
                      "passed": tally.passed, "failed": tally.failed, "errors": tally.errors,
                      "skipped": tally.skipped, "failures": tally.failures[:20], "lshrs": str(lshrs.__file__),
                      "lshrs_LSHRS_module": lshrs.LSHRS.__module__, "kernel_launches": launches}))
    return 0 if rc == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
