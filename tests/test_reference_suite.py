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


@needs_reference
def test_notebook_cells_run_unchanged_on_this_package(tmp_path):
    """Drop-in check at the user level: the code cells of the reference's demo.ipynb are executed as they are (read from
    /root/reference at run time, never copied into the repo) with `src` aliased to this package.  Only the environment is
    adapted: matplotlib is stubbed (not installed here), the `%matplotlib` magic is dropped, the device string becomes
    "cpu" (no GPU in the build container) and the loop is cut to 40 steps.  The printed loss stream must reproduce the
    values published in the notebook's own output."""
    import json
    nb = json.load(open(os.path.join(REFERENCE, "demo.ipynb")))
    cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
    published = {}
    for c in nb["cells"]:
        for out in c.get("outputs", []):
            for line in "".join(out.get("text", [])).splitlines():
                if line.startswith("Step ") and ", loss: " in line:
                    k, v = line[5:].split(", loss: ")
                    published[int(k)] = float(v)
    assert len(published) >= 1000
    code = "\n\n".join(cells)
    code = code.replace("%matplotlib inline", "").replace('device="cuda"', 'device="cpu"')
    code = code.replace("training_steps = 5000", "training_steps = 40")
    assert 'device="cpu"' in code and "training_steps = 40" in code
    (tmp_path / "matplotlib").mkdir()
    (tmp_path / "matplotlib" / "__init__.py").write_text("")
    (tmp_path / "matplotlib" / "pyplot.py").write_text(
        "class _Anything:\n    def __getattr__(self, name):\n        return _Anything()\n"
        "    def __call__(self, *a, **k):\n        return _Anything()\n"
        "import sys\nsys.modules[__name__] = _Anything()\n")
    script = tmp_path / "notebook_cells.py"
    script.write_text(code)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(tmp_path), os.path.join(REPO, "tests", "_src_alias"), REPO])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    got = {}
    for line in r.stdout.splitlines():
        if line.startswith("Step ") and ", loss: " in line:
            k, v = line[5:].split(", loss: ")
            got[int(k)] = float(v)
    assert sorted(got) == list(range(40)), r.stdout[-2000:]
    assert "number of parameters 2817" in r.stdout
    for k in range(40):
        assert abs(got[k] - published[k]) <= 1e-6 * abs(published[k]), (k, got[k], published[k])
