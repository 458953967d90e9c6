#!/usr/bin/env python
"""Run the reference's OWN test-suite (staged by tools/make_ref.py under oracle/_ref/suite/) against one of
three arrangements of ``import lshrs`` and print one JSON line with the outcome.

    python tests/ref_suite_runner.py --variant reference            # the unmodified reference (harness check)
    python tests/ref_suite_runner.py --variant two_import           # reference LSHRS, B200 hasher + rerank
    python tests/ref_suite_runner.py --variant dropin               # lshrs_b200.compat: B200 everything on the path
    python tests/ref_suite_runner.py --variant dropin --cpu-double  # host logic only, oracle-backed C-ABI double
    python tests/ref_suite_runner.py --variant dropin --device-store   # ... and the suite's MockStorage fixture is
                                                                       # the bucket store in HBM (DeviceBucketStorage)

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


class _DeviceStoreFixture:
    """Swaps the suite's ``MockStorage`` (reference tests/conftest.py:15-78, a dict of sets that records what it is
    sent) for the same recorder on top of ``lshrs_b200.DeviceBucketStorage``: every test that builds its LSHRS on
    the fixture then runs with the bucket store in HBM -- packed ``index()``, device joins for the queries."""

    def pytest_collection_finish(self, session):
        import numpy as np

        from lshrs_b200 import DeviceBucketStorage

        class MockStorage(DeviceBucketStorage):
            def __init__(self, *, fail_on_flush: bool = False) -> None:
                super().__init__()
                self.batches, self.all_operations, self.removed_indices = [], [], []
                self.batch_add_call_count = 0
                self.close_called = self.clear_called = False
                self._fail_on_flush = fail_on_flush

            def _record(self, operations) -> None:
                self.batch_add_call_count += 1
                self.batches.append(list(operations))
                self.all_operations.extend(operations)

            def batch_add(self, operations) -> None:
                if self._fail_on_flush:
                    raise ConnectionError("Simulated Redis failure")
                self._record(operations)
                super().batch_add(operations)

            def add_packed(self, signatures, ids) -> None:
                if self._fail_on_flush:
                    raise ConnectionError("Simulated Redis failure")
                sig = np.asarray(signatures)
                self._record([(b, sig[i, b].tobytes(), int(ids[i])) for i in range(sig.shape[0])
                              for b in range(sig.shape[1])])
                super().add_packed(signatures, ids)

            def remove_indices(self, indices) -> None:
                self.removed_indices.append(list(indices))
                super().remove_indices(indices)

            def clear(self) -> None:
                self.clear_called = True
                super().clear()

            def close(self) -> None:
                self.close_called = True

            @property
            def total_operations(self) -> int:
                return len(self.all_operations)

            @property
            def unique_indices(self) -> set:
                return {idx for _, _, idx in self.all_operations}

        patched = 0
        for name, mod in list(sys.modules.items()):
            if name.endswith("conftest") and hasattr(mod, "MockStorage") and hasattr(mod, "make_lsh"):
                mod.MockStorage = MockStorage
                patched += 1
        self.patched = patched


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", choices=("reference", "two_import", "dropin"), required=True)
    ap.add_argument("--cpu-double", action="store_true")
    ap.add_argument("--device-store", action="store_true")
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
    store = _DeviceStoreFixture() if args.device_store else None
    rc = pytest.main(argv, plugins=[tally] + ([store] if store else []))
    import lshrs

    launches = None
    if args.variant != "reference":
        from lshrs_b200 import _native

        launches = fake.launches if fake is not None else (_native.launch_count() if _native._lib is not None else 0)
    print(json.dumps({"staged": True, "variant": args.variant, "cpu_double": bool(args.cpu_double), "rc": int(rc),
                      "device_store": bool(store and getattr(store, "patched", 0)), "passed": tally.passed, "failed": tally.failed, "errors": tally.errors,
                      "skipped": tally.skipped, "failures": tally.failures[:20], "lshrs": str(lshrs.__file__),
                      "lshrs_LSHRS_module": lshrs.LSHRS.__module__, "kernel_launches": launches}))
    return 0 if rc == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
