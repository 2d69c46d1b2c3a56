"""The headless command-line front end (cloth_sim_self_collision_b200/cli/clothsim.cpp -> clothsim): the two sparse
demos of the reference from the command line, state dump / restore, the reference's per-step stdout lines."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from cloth_sim_self_collision_b200 import clothgpu as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cloth_sim_self_collision_b200", "clothsim")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(G.LIB_PATH):
        G.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cloth_sim_self_collision_b200", "compat")])
    return CLI


def read_state(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"CLOTHSIM"
    ver, npnt, step = struct.unpack("iii", raw[8:20])
    assert ver == 1
    x = np.frombuffer(raw, dtype=np.float64, count=npnt * 3, offset=20)
    v = np.frombuffer(raw, dtype=np.float64, count=npnt * 3, offset=20 + npnt * 24)
    return step, x, v


def test_cli_usage_and_no_cpu_fallback(cli):
    out = subprocess.run([cli], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr
    if G.device_count() == 0:
        out = subprocess.run([cli, "self_contact", "--steps", "1"], capture_output=True, text=True)
        assert out.returncode == 1 and "clothsim:" in out.stderr


@pytest.mark.gpu
def test_cli_self_contact_matches_reference_and_restores(cli, tmp_path):
    gold = np.load(os.path.join(GOLD, "traj_config1.npz"))
    steps = [int(s) for s in gold["steps"]]
    a, b, c = (str(tmp_path / n) for n in ("a.bin", "b.bin", "c.bin"))
    out = subprocess.run([cli, "self_contact", "--steps", "40", "--dump", a, "--normals", str(tmp_path / "n.bin")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    step, x, v = read_state(a)
    k = steps.index(40)
    assert step == 40 and np.array_equal(x, gold["x"][k]) and np.array_equal(v, gold["v"][k])
    # the reference's stdout lines, from the statistics of each step
    its = [int(m) for m in re.findall(r"iteration:(\d+)", out.stdout)]
    prox = [int(m) for m in re.findall(r"Proximity      Contact Elem Size: (\d+)", out.stdout)]
    assert its == [int(s[0]) for s in gold["stats"][:40]]
    assert prox == [int(s[1]) for s in gold["stats"][:40]]
    assert out.stdout.count("CCD iter: 0") == 40
    assert out.stdout.count("update middle velocity") == int(sum(s[4] for s in gold["stats"][:40]))
    nrm = np.fromfile(str(tmp_path / "n.bin")).reshape(-1, 3)
    assert nrm.shape == (255, 3) and np.allclose(np.linalg.norm(nrm, axis=1), 1.0)
    # dump after 10 steps, restore, 30 more: the same state
    assert subprocess.run([cli, "self_contact", "--steps", "10", "--dump", b, "--quiet"], capture_output=True).returncode == 0
    out = subprocess.run([cli, "self_contact", "--steps", "30", "--restore", b, "--dump", c, "--quiet"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    step, x2, v2 = read_state(c)
    assert step == 40 and np.array_equal(x2, x) and np.array_equal(v2, v)


@pytest.mark.gpu
def test_cli_internal_demo_matches_reference(cli, tmp_path):
    gold = np.load(os.path.join(GOLD, "traj_internal26.npz"))
    a = str(tmp_path / "a.bin")
    out = subprocess.run([cli, "internal", "--collider", "sphere", "--steps", "40", "--dump", a], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    step, x, v = read_state(a)
    assert np.array_equal(x, gold["x"][1]) and np.array_equal(v, gold["v"][1])
    assert [int(m) for m in re.findall(r"iteration:(\d+)", out.stdout)] == [int(i) for i in gold["iters"]]
    # the unpreconditioned solver variant runs too (StepTime_InternalDynamics)
    out = subprocess.run([cli, "internal", "--solver", "cg", "--steps", "3", "--quiet"], capture_output=True, text=True)
    assert out.returncode == 0 and "checksum=" in out.stdout
