"""Drop-in proof (``-m gpu``): the reference's OWN unit tests for the hot-path layers, unmodified, against the ``compyute``
import shim (repo root ``compyute/`` -> compyute_b200) on the cuda device.

The reference tests (dakofler/Compyute ``tests/nn/modules/{convolutions,linear,poolings,normalizations,activations,
containers,regularizations,reshapes}_test.py``, ``tests/nn/test_optimizers.py``, ``tests/nn/test_losses.py``) build a layer,
load it with the parameters of a ``torch.nn`` twin and compare forward / backward results with ``allclose(1e-5)`` (db and
optimizer state 1e-4).  They are run in a subprocess whose ``tests`` package is the reference's (they import
``tests.utils``), with ``COMPYUTE_SHIM_DEVICE=cuda`` so that the ``device=cpu`` their utilities hard-code lands on the GPU.
Out-of-scope cases in the same files (Conv1D, transposed convolutions, LayerNorm, GELU, ...) are deselected by name.

The test files are never copied into the repository: ``oracle/fetch_reference_tests.sh`` places them under the git-ignored
``oracle/_ref/reference_tests`` (which travels to the GPU box like a built library); without them the test skips.
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SELECTION = [  # (file, -k expression, compute mode)
    ("tests/nn/modules/convolutions_test.py", "test_conv2d", "fp32"),
    ("tests/nn/modules/linear_test.py", "", "fp32"),
    ("tests/nn/modules/poolings_test.py", "maxpool2d or avgpool2d", "fp32"),
    ("tests/nn/modules/normalizations_test.py", "batchnorm1d or batchnorm2d", "fp32"),
    ("tests/nn/modules/activations_test.py", "test_relu", "fp32"),
    ("tests/nn/modules/containers_test.py", "sequential or residual", "fp32"),
    ("tests/nn/modules/regularizations_test.py", "", "fp32"),
    ("tests/nn/modules/reshapes_test.py", "", "fp32"),
    ("tests/nn/test_optimizers.py", "", "fp32"),
    ("tests/nn/test_losses.py", "cross_entropy", "fp32"),
    # "fp32" above is the tensor-core exact mode (3xTF32); the same contraction tests on the FFMA kernels: identical contract
    ("tests/nn/modules/convolutions_test.py", "test_conv2d", "fp32_simt"),
    ("tests/nn/modules/linear_test.py", "", "fp32_simt"),
]


def _reference_tests_dir():
    for cand in (os.environ.get("COMPYUTE_REFERENCE_TESTS"), os.path.join(ROOT, "oracle", "_ref", "reference_tests"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "tests", "nn", "modules")):
            return cand
    return None


@pytest.mark.parametrize("path,expr,mode", SELECTION, ids=[f"{os.path.basename(p)}[{m}]" for p, _, m in SELECTION])
def test_reference_unit_tests_on_cuda(path, expr, mode):
    ref = _reference_tests_dir()
    if ref is None:
        pytest.skip("reference tests not available (run oracle/fetch_reference_tests.sh where /root/reference exists)")
    env = dict(os.environ)
    stubs = os.path.join(ROOT, "compyute", "_teststubs")  # `import torchtune` in normalizations_test (out-of-scope RMSNorm case)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, stubs] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env["COMPYUTE_SHIM_DEVICE"] = "cuda"
    env["COMPYUTE_B200_MODE"] = mode
    cmd = [sys.executable, "-m", "pytest", path, "-q", "-x", "-p", "no:cacheprovider", "--rootdir", ref, "-c", os.devnull]
    if expr:
        cmd += ["-k", expr]
    r = subprocess.run(cmd, cwd=ref, env=env, capture_output=True, text=True, timeout=1200)
    tail = (r.stdout or "")[-3000:]
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "dropin_reference_tests.jsonl"), "a") as f:
            last = [ln for ln in (r.stdout or "").splitlines() if "passed" in ln or "failed" in ln or "error" in ln.lower()]
            f.write(json.dumps({"file": path, "k": expr, "mode": mode, "rc": r.returncode, "summary": last[-1] if last else ""}) + "\n")
    assert r.returncode == 0, f"reference tests failed:\n{tail}\n{(r.stderr or '')[-1500:]}"
    assert " passed" in tail and "no tests ran" not in tail, tail
