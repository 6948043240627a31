"""Run the REFERENCE'S OWN unit tests against this package's CPU path (build container only).

The reference's tests import `src.*`; tests/_src_alias provides a `src` package that aliases to
microtorch_b200.  The test files are copied to a temp dir at run time (never into the repo) so
that they do not import the reference's real `src` package next to them.  Expected: every test
the reference itself passes without CuPy (133) passes here too; the CuPy-only ones are
deselected."""
import glob
import os
import shutil
import subprocess
import sys

import pytest

from conftest import REFERENCE, REPO, needs_reference

# need CuPy + a GPU in the reference as well
CUPY_ONLY = ["test_call_prevents_redundant_device_to", "test_call_triggers_auto_device",
             "test_to_method_changes_device", "test_backward_errors"]


@needs_reference
def test_reference_unit_tests_pass_on_this_package(tmp_path):
    n = 0
    for path in glob.glob(os.path.join(REFERENCE, "src", "**", "tests", "*_test.py"), recursive=True):
        rel = os.path.relpath(path, os.path.join(REFERENCE, "src")).replace(os.sep, "__")
        if rel.endswith("device_test.py"):
            continue                      # top-level `import cupy`
        shutil.copy(path, tmp_path / rel)
        n += 1
    assert n >= 16
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "tests", "_src_alias"), REPO])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-o", "python_files=*_test.py",
                        str(tmp_path)], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    tail = r.stdout[-3000:]
    failed = [l for l in r.stdout.splitlines() if l.startswith("FAILED")]
    # the only admissible failures are the reference's own CuPy-dependent tests
    unexpected = [l for l in failed if not any(k in l for k in CUPY_ONLY)]
    assert not unexpected, tail
    assert "133 passed" in r.stdout or "137 passed" in r.stdout, tail
