"""The C++ drop-in layer (include/clothgpu_compat/ + libclothcompat.so): a headless copy of the reference demo's
main() + StepTime() (tests/cpp/self_contact_headless.cpp) compiled against it.

CPU: it builds, links and runs its set-up side; computing without a GPU fails loudly (no CPU fallback).
GPU: 40 steps of the default self_contact_sparse scene through the reference-named API reproduce the golden
trajectory recorded from the reference itself (tests/golden/traj_config1.npz)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from cloth_sim_self_collision_b200 import clothgpu as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "cpp", "_build", "self_contact_headless")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def demo():
    if not os.path.exists(G.LIB_PATH):
        G.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cloth_sim_self_collision_b200", "compat")])
    return DEMO


def test_compat_builds_links_and_sets_up(demo, tmp_path):
    out = subprocess.run([demo, "4", "50", "0", str(tmp_path / "o.bin"), "--check-link"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "link ok: np=255 ntri=400 nquad=546 ncrs=2400 nbvh=799" in out.stdout
    hdr = open(os.path.join(ROOT, "include", "clothgpu_compat", "cloth_compat.h")).read()
    for name in ("StepTime_InternalDynamicsILU", "StepTime_InternalDynamics", "UpdateIntermidiateVelocity",
                 "GetIntermidiateVelocityContactResolved", "class CMatrixSquareSparse", "class CPreconditionerILU",
                 "Solve_PCG", "Solve_CG", "InnerProduct", "AXPY", "BuildBoundingBoxChild_Prx",
                 "BuildBoundingBoxChild_CCD", "GetContactElement_Proximity", "GetContactElement_CCD",
                 "class CContactElement", "class CNodeBVH", "class CAABB3D", "class CJaggedArray",
                 "MakeBVHTopology_TopDown"):
        assert name in hdr, name


def test_compat_refuses_to_compute_without_gpu(demo, tmp_path):
    if G.device_count() > 0:
        pytest.skip("a GPU is visible")
    out = subprocess.run([demo, "4", "50", "1", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert out.returncode == 1
    assert "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_compat_demo_reproduces_reference_trajectory(demo, tmp_path):
    gold = np.load(os.path.join(GOLD, "traj_config1.npz"))
    bvh = np.ascontiguousarray(gold["bvh"], dtype=np.int32)
    bvh_file = tmp_path / "bvh.bin"
    bvh.tofile(bvh_file)
    steps = [int(s) for s in gold["steps"]]
    step = 40 if 40 in steps else steps[0]
    outf = tmp_path / "o.bin"
    out = subprocess.run([demo, "4", "50", str(step), str(outf), "--bvh", str(bvh_file)], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    raw = open(outf, "rb").read()
    npnt, = struct.unpack("i", raw[:4])
    assert npnt == 255
    x = np.frombuffer(raw, dtype=np.float64, count=npnt * 3, offset=12)
    v = np.frombuffer(raw, dtype=np.float64, count=npnt * 3, offset=12 + npnt * 24)
    k = steps.index(step)
    ex = np.abs(x - gold["x"][k]).max() / np.abs(gold["x"][k]).max()
    ev = np.abs(v - gold["v"][k]).max() / np.abs(gold["v"][k]).max()
    assert np.array_equal(x, gold["x"][k]) and np.array_equal(v, gold["v"][k]), (ex, ev)


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own demo sources, UNCHANGED, compiled against include/clothgpu_compat (tests/cpp/Makefile: a shadow
# tree makes their "../x.h" includes resolve to the drop-in headers, tests/cpp/glut_stub stands in for GLUT).  The
# binaries are built where /root/reference exists and travel with the snapshot.
REF_BINS = {name: os.path.join(ROOT, "tests", "cpp", "_build", "ref_" + name)
            for name in ("self_contact_sparse", "internal_cloth_sparse")}


@pytest.fixture(scope="module")
def ref_demos(demo):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    missing = [b for b in REF_BINS.values() if not os.path.exists(b)]
    if missing:
        pytest.skip("reference demos not built (needs /root/reference at build time): %s" % missing)
    return REF_BINS


def _run_demo(binary, keys, steps, dump):
    env = dict(os.environ, CLOTHGPU_GLUT_KEYS=keys, CLOTHGPU_GLUT_STEPS=str(steps), CLOTHGPU_GLUT_DUMP=str(dump))
    return subprocess.run([binary], capture_output=True, text=True, env=env, timeout=900)


def _read_dump(path):
    raw = open(path, "rb").read()
    npnt, = struct.unpack("i", raw[:4])
    x = np.frombuffer(raw, dtype=np.float64, count=npnt * 3, offset=4)
    v = np.frombuffer(raw, dtype=np.float64, count=npnt * 3, offset=4 + npnt * 24)
    return npnt, x, v


def test_reference_demos_refuse_to_compute_without_gpu(ref_demos, tmp_path):
    if G.device_count() > 0:
        pytest.skip("a GPU is visible")
    for b in ref_demos.values():
        out = _run_demo(b, "a", 1, tmp_path / "o.bin")
        assert out.returncode != 0 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_unchanged_self_contact_sparse_main_reproduces_reference(ref_demos, tmp_path):
    """self_contact_sparse/main.cpp as shipped: its own main() (scene, MakeBVHTopology_TopDown, SetPattern,
    Initialize_ILU0) and StepTime() through the drop-in, 40 steps: bit-identical to the reference's trajectory."""
    gold = np.load(os.path.join(GOLD, "traj_config1.npz"))
    steps = [int(s) for s in gold["steps"]]
    for step in (10, 40):
        out = _run_demo(ref_demos["self_contact_sparse"], "a", step, tmp_path / "o.bin")
        assert out.returncode == 0, out.stderr + out.stdout
        npnt, x, v = _read_dump(tmp_path / "o.bin")
        k = steps.index(step)
        assert npnt == 255
        assert np.array_equal(x, gold["x"][k]) and np.array_equal(v, gold["v"][k]), step


@pytest.mark.gpu
def test_unchanged_internal_cloth_sparse_main_reproduces_reference(ref_demos, tmp_path):
    """internal_cloth_sparse/main.cpp as shipped (26 x 26, pinned row), sphere collider selected with two presses of the
    space key, then 'a': the host callback collider goes through the per-vertex table path."""
    gold = np.load(os.path.join(GOLD, "traj_internal26.npz"))
    for k, step in enumerate(int(s) for s in gold["steps"]):
        out = _run_demo(ref_demos["internal_cloth_sparse"], "  a", step, tmp_path / "o.bin")
        assert out.returncode == 0, out.stderr + out.stdout
        npnt, x, v = _read_dump(tmp_path / "o.bin")
        assert npnt == 26 * 26
        assert np.array_equal(x, gold["x"][k]) and np.array_equal(v, gold["v"][k]), step
