"""The reference's OWN demo scripts, unchanged, against this repo's build (BASELINE north_star: "1_test_solve.py and
2_test_creatematrix.py run unchanged").  A temporary tree is assembled from symlinks to the scripts and `utils/` where they lie
under /root/reference, a COPY of the two small asset folders (the scripts write into them), this repo's `XM/build` and a headless
`open3d` shim (2_test_creatematrix.py imports utils.visualization -> open3d, which opens windows).  Needs both the reference tree
and a GPU: skipped wherever one of them is missing (the reference tree does not travel to the GPU box; the calls the scripts make
are covered there by tests/test_gpu_solve.py::test_xm_module_runs_the_reference_demo_call and ::test_simple2_pipeline_end_to_end)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = "/root/reference"

OPEN3D_SHIM = '''
"""Headless stand-in for open3d: just enough surface for utils/visualization.py to import and run without a display."""
import types
import numpy as np
class _Any:
    def __init__(self, *a, **k): self.lines = []; self.colors = None; self.points = None
    def __getattr__(self, name): return lambda *a, **k: None
class _LineSet(_Any):
    @staticmethod
    def create_camera_visualization(*a, **k):
        ls = _LineSet(); ls.lines = list(range(8)); return ls
camera = types.SimpleNamespace(PinholeCameraIntrinsic=_Any)
visualization = types.SimpleNamespace(Visualizer=_Any, draw_geometries=lambda *a, **k: None)
geometry = types.SimpleNamespace(LineSet=_LineSet, PointCloud=_Any, TriangleMesh=_Any)
utility = types.SimpleNamespace(Vector3dVector=lambda x: np.asarray(x), Vector2iVector=lambda x: np.asarray(x))
io = types.SimpleNamespace(write_point_cloud=lambda *a, **k: None, read_point_cloud=lambda *a, **k: _Any())
'''


@pytest.fixture
def ref_tree(tmp_path):
    import torch
    if not os.path.isdir(os.path.join(REF, "utils")) or not torch.cuda.is_available():
        pytest.skip("needs the reference tree (/root/reference) AND a GPU on the same machine")
    t = tmp_path / "tree"
    t.mkdir()
    for name in ("1_test_solve.py", "2_test_creatematrix.py", "utils"):
        os.symlink(os.path.join(REF, name), t / name)
    (t / "assets").mkdir()
    for a in ("SIMPLE1", "SIMPLE2"):
        shutil.copytree(os.path.join(REF, "assets", a), t / "assets" / a)
    (t / "XM").mkdir()
    os.symlink(os.path.join(ROOT, "XM", "build"), t / "XM" / "build")
    (t / "shim").mkdir()
    (t / "shim" / "open3d.py").write_text(OPEN3D_SHIM)
    return t


def run(tree, script):
    env = dict(os.environ, PYTHONPATH=str(tree / "shim") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    return subprocess.run([sys.executable, script], cwd=str(tree), env=env, capture_output=True, text=True, timeout=900)


def load_bin(path):
    with open(path, "rb") as f:
        r = int.from_bytes(f.read(4), "little"); c = int.from_bytes(f.read(4), "little")
        return np.fromfile(f, dtype=np.float64, count=r * c).reshape((r, c), order="F")


def test_1_test_solve_runs_unchanged(ref_tree):
    out = run(ref_tree, "1_test_solve.py")
    assert out.returncode == 0, out.stderr[-3000:]
    assert "BM finished with rank 3" in out.stdout
    s = load_bin(ref_tree / "assets" / "SIMPLE1" / "s.bin")
    R = load_bin(ref_tree / "assets" / "SIMPLE1" / "R.bin")
    assert R.shape == (447, 3) and s.shape == (149, 1) and s[0, 0] == 1.0
    np.testing.assert_allclose(s[1:6, 0], [0.99764946, 1.0008994, 1.00081364, 1.00099953, 1.00061465], atol=1e-7)


def test_2_test_creatematrix_runs_unchanged(ref_tree):
    out = run(ref_tree, "2_test_creatematrix.py")
    assert out.returncode == 0, out.stderr[-3000:]
    d = ref_tree / "assets" / "SIMPLE2"
    R = load_bin(d / "R.bin"); s = load_bin(d / "s.bin")
    assert R.shape[0] == 279 and R.shape[1] in (3, 4) and s.shape == (93, 1)
    assert abs(float(np.mean(s)) - 1.0) < 0.01
