"""CPU-side checks of the product library: it loads, exports every symbol include/clothgpu.h declares, validates its
arguments on the host and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cloth_sim_self_collision_b200 import clothgpu as G
from cloth_sim_self_collision_b200 import scenes as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(G.LIB_PATH):
        G.build()
    return G.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "clothgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clothgpu_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), "libclothgpu.so does not export %s" % s


def test_version_and_struct_layout(lib):
    assert lib.clothgpu_version() == 200
    # clothgpu_params: 10 doubles, 2 ints, 4 doubles
    assert C.sizeof(G.Params) == 10 * 8 + 2 * 4 + 4 * 8


def test_argument_validation_and_no_cpu_fallback(lib):
    sc = S.square_cloth_xz(2, 3, 0.1)
    h = C.c_void_p()
    bad_tri = sc["tri"].copy()
    bad_tri[0] = sc["np"] + 5

    def create(tri=sc["tri"], colidx=sc["colidx"]):
        ip = lambda a: np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int))
        return lib.clothgpu_create(C.byref(h), -1, 1, sc["np"], sc["ntri"], sc["nquad"], ip(tri), ip(sc["quad"]),
                                   ip(sc["bc"]), sc["xyz0"].ctypes.data_as(C.POINTER(C.c_double)), ip(sc["rowptr"]),
                                   ip(colidx), 0, None, None, None)

    assert create(tri=bad_tri) == 1  # CLOTHGPU_ERR_ARG
    assert b"triangle" in lib.clothgpu_last_error()
    bad_col = sc["colidx"].copy()
    bad_col[:2] = bad_col[:2][::-1]
    assert create(colidx=bad_col) == 1
    if lib.clothgpu_device_count() == 0:
        assert create() == 2  # CLOTHGPU_ERR_CUDA: no silent CPU path
        assert b"no CPU fallback" in lib.clothgpu_last_error()
        with pytest.raises(G.ClothGPUError):
            G.ClothGPU(sc, with_bvh=False)


def test_median_split_bvh_is_a_valid_tree():
    sc = S.square_cloth_xz(7, 5, 0.1)
    nodes = S.median_split_bvh(sc)
    ntri = sc["ntri"]
    assert nodes.shape == (2 * ntri - 1, 3) and nodes[0, 0] == -1
    leaves = nodes[nodes[:, 2] == -1]
    assert sorted(leaves[:, 1].tolist()) == list(range(ntri))
    for i, (par, c0, c1) in enumerate(nodes):
        if c1 != -1:
            assert nodes[c0, 0] == i and nodes[c1, 0] == i and c0 > i and c1 > i


def _selftest(lib, sc_or_pattern, force_ns=0):
    if isinstance(sc_or_pattern, dict):
        n, rp, ci = sc_or_pattern["np"], sc_or_pattern["rowptr"], sc_or_pattern["colidx"]
    else:
        n, rp, ci = sc_or_pattern
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    st = np.zeros(16, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = lib.clothgpu_selftest_triplan(int(n), ip(rp), ip(ci), int(force_ns), ip(st))
    assert rc == 0, lib.clothgpu_last_error()
    return st.reshape(2, 8)


@pytest.mark.parametrize("nx,nz,force", [(2, 3, 0), (4, 50, 0), (4, 50, 3), (23, 40, 0), (23, 40, 7), (63, 63, 0),
                                         (63, 63, 16), (127, 127, 0)])
def test_trisolve_plan_invariants(lib, nx, nz, force):
    """The strip / micro-step plan of the triangular solves (csrc/trisolve.cu) is host logic: every row scheduled
    once, dependencies earlier or imported in time, packed-value map exact -- verified inside the library."""
    sc = S.square_cloth_xz(nx, nz, 0.1)
    st = _selftest(lib, sc, force)
    for d in range(2):
        ns, R, nms, longest, ghosts, makespan = st[d, :6]
        assert ns >= 1 and ns * R >= sc["np"] and nms >= longest >= 1
        if force:
            assert ns == -(-sc["np"] // -(-sc["np"] // force))
    if (nx, nz, force) == (127, 127, 0):
        # 128 x 128 grid: 509 dependency levels; the pipelined plan must stay close to that critical path
        assert st[0, 5] < 509 * 1.6 and st[1, 5] < 509 * 1.6


def test_trisolve_plan_irregular_pattern(lib):
    """A random symmetric pattern (not a grid): rows with many / no dependencies, long-range couplings."""
    rng = np.random.default_rng(5)
    n = 700
    pairs = set()
    for i in range(n):
        for j in rng.integers(0, n, size=rng.integers(0, 9)):
            if i != j:
                pairs.add((i, int(j)))
                pairs.add((int(j), i))
    rows = [[] for _ in range(n)]
    for i, j in pairs:
        rows[i].append(j)
    rp = np.zeros(n + 1, dtype=np.int32)
    ci = []
    for i in range(n):
        rows[i].sort()
        ci += rows[i]
        rp[i + 1] = len(ci)
    for force in (0, 1, 5, 33):
        _selftest(lib, (n, rp, np.array(ci, dtype=np.int32)), force)


def test_trisolve_plan_degenerate_patterns(lib):
    """Edge cases of the planner: a single row, a purely diagonal matrix (no dependencies at all: one level), a
    dense-ish arrow pattern (one row depending on all others: more operands than the unrolled kernels hold)."""
    st = _selftest(lib, (1, np.array([0, 0]), np.zeros(0, dtype=np.int32)))
    assert st[0, 0] == 1 and st[0, 2] >= 1
    n = 40
    st = _selftest(lib, (n, np.zeros(n + 1, dtype=np.int32), np.zeros(0, dtype=np.int32)))
    assert st[0, 2] % 4 == 0  # micro-steps are padded to whole TMA stages
    # arrow: last row coupled to every other row (15 lower operands > TS_MAXD = 8)
    n = 16
    rows = [[n - 1] for _ in range(n - 1)] + [list(range(n - 1))]
    rp = np.zeros(n + 1, dtype=np.int32)
    ci = []
    for i in range(n):
        ci += rows[i]
        rp[i + 1] = len(ci)
    for force in (0, 1, 3):
        _selftest(lib, (n, rp, np.array(ci, dtype=np.int32)), force)


def test_sqrt_comparison_thresholds(lib):
    """The impulse chain compares squared lengths against host-computed thresholds instead of taking square roots
    (csrc/imp_fast.cuh): `sqrt(x) < c` must be exactly `x < ge` and `sqrt(x) > c` exactly `x > le`, for the constants the
    reference uses (1e-10, bvh_aabb.cpp:350) and for typical clearances.  IEEE sqrt is correctly rounded and monotone,
    so checking the two doubles on either side of each threshold proves the equivalence."""
    lib.clothgpu_selftest_sqrt_thresholds.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(5)
    cs = [1.0e-10, 0.01, 0.02, 0.05, 1.0, 3.0] + list(10.0 ** rng.uniform(-12, 3, 200))
    for c in cs:
        ge, le = C.c_double(), C.c_double()
        assert lib.clothgpu_selftest_sqrt_thresholds(c, C.byref(ge), C.byref(le)) == 0
        ge, le = ge.value, le.value
        assert np.sqrt(ge) >= c and np.sqrt(np.nextafter(ge, -np.inf)) < c, c
        assert np.sqrt(le) <= c and np.sqrt(np.nextafter(le, np.inf)) > c, c
    assert lib.clothgpu_selftest_sqrt_thresholds(0.0, C.byref(C.c_double()), C.byref(C.c_double())) != 0
