"""ctypes binding of libclothgpu.so (include/clothgpu.h) -- the product's Python host side.

This module only marshals numpy arrays into the C ABI; every computation happens in the CUDA library.  There is
no CPU fallback: if the library is missing or no GPU is visible, construction fails loudly.

The method names follow the reference functions they stand for (cited in include/clothgpu.h), e.g.
``assemble`` = the left/right-hand side of ``StepTime_InternalDynamicsILU``, ``pcg`` = ``Solve_PCG``,
``contacts_ccd`` = ``GetContactElement_CCD``, ``resolve_self_contact`` = ``GetIntermidiateVelocityContactResolved``,
``step_self_contact`` = ``StepTime()`` of self_contact_sparse/main.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# CLOTHGPU_FAST=1 selects the non-parity fast build (libclothgpu_fast.so: FMA contraction on, tree reductions)
FAST = os.environ.get("CLOTHGPU_FAST", "0") not in ("", "0")
LIB_PATH = os.path.join(HERE, "libclothgpu_fast.so" if FAST else "libclothgpu.so")
# development switch: A/B a variant build of the same sources (tools/build_variant.sh), e.g. different -D tunables
if os.environ.get("CLOTHGPU_LIB_PATH"):
    LIB_PATH = os.path.abspath(os.environ["CLOTHGPU_LIB_PATH"])

NSTATS = 256
NPHASE = 16
PHASE_NAMES = ["assembly", "ilu_factor", "pcg", "bvh_refit", "traverse_narrow", "sort_unique", "impulse", "riz",
               "velocity_fixup"]

(BUF_XYZ, BUF_UVW, BUF_XYZ_PREV, BUF_VAL_DIA, BUF_VAL_CRS, BUF_ILU_DIA, BUF_ILU_CRS, BUF_VEC_R, BUF_VEC_X, BUF_VEC_P,
 BUF_VEC_Z, BUF_BOXES, BUF_ELEMS, BUF_LEAF_PAIRS, BUF_ZONE, BUF_UVW0, BUF_COLLIDER_TABLE) = range(17)
_INT_BUFS = {BUF_ELEMS, BUF_LEAF_PAIRS, BUF_ZONE}

COLLIDER_NONE, COLLIDER_PLANE_Z, COLLIDER_SPHERE, COLLIDER_SPHERE_INSIDE, COLLIDER_TABLE = range(5)


class ClothGPUError(RuntimeError):
    pass


class Params(C.Structure):
    """clothgpu_params (include/clothgpu.h)."""
    _fields_ = [("dt", C.c_double), ("lambda_", C.c_double), ("myu", C.c_double), ("stiff_bend", C.c_double),
                ("gravity", C.c_double * 3), ("mass_point", C.c_double), ("stiff_contact", C.c_double),
                ("contact_clearance", C.c_double), ("collider_kind", C.c_int), ("reserved", C.c_int),
                ("collider_par", C.c_double * 4)]


def make_params(dt, lam, myu, stiff_bend, gravity, mass_point, stiff_contact, clearance, col_kind=0,
                col_par=(0, 0, 0, 0)):
    p = Params()
    p.dt, p.lambda_, p.myu, p.stiff_bend = float(dt), float(lam), float(myu), float(stiff_bend)
    for k in range(3):
        p.gravity[k] = float(gravity[k])
    p.mass_point, p.stiff_contact, p.contact_clearance = float(mass_point), float(stiff_contact), float(clearance)
    p.collider_kind = int(col_kind)
    for k in range(4):
        p.collider_par[k] = float(col_par[k])
    return p


def params_from(obj):
    """Accepts a Params, or any object with the oracle's Params attributes (dt, lam, myu, ...)."""
    if isinstance(obj, Params):
        return obj
    return make_params(obj.dt, obj.lam, obj.myu, obj.stiff_bend, obj.gravity, obj.mass_point, obj.stiff_contact,
                       obj.clearance, obj.col_kind, obj.col_par)


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libclothgpu.so (in-tree)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j", str(min(8, os.cpu_count() or 1))]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ClothGPUError("libclothgpu.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.clothgpu_last_error.restype = C.c_char_p
    lib.clothgpu_build_mode.restype = C.c_char_p
    lib.clothgpu_buf_len.restype = C.c_longlong
    lib.clothgpu_launch_count.restype = C.c_longlong
    lib.clothgpu_buf_len.argtypes = [C.c_void_p, C.c_int]
    lib.clothgpu_alloc_host.argtypes = [C.POINTER(C.c_void_p), C.c_longlong]
    lib.clothgpu_free_host.argtypes = [C.c_void_p]
    lib.clothgpu_launch_count.argtypes = [C.c_void_p]
    lib.clothgpu_write.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
    lib.clothgpu_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
    _lib = lib
    return lib


def device_count():
    return load().clothgpu_device_count()


def bvh_build_topdown(tri, xyz, return_fallbacks=False):
    """MakeBVHTopology_TopDown (bvh_aabb.cpp:294-313) through the C ABI (host code, no GPU needed): the reference's
    own CNodeBVH array, (2*ntri-1) x {iroot, ichild0, ichild1}."""
    lib = load()
    tri = _ci(tri).ravel()
    xyz = _cd(xyz).ravel()
    ntri = tri.size // 3
    nodes = np.empty((max(2 * ntri - 1, 1), 3), dtype=np.int32)
    nfb = C.c_int(0)
    n = lib.clothgpu_bvh_build_topdown(C.c_int(ntri), _ip(tri), C.c_int(xyz.size // 3), _dp(xyz), _ip(nodes),
                                       C.c_int(nodes.shape[0]), C.byref(nfb))
    if n < 0:
        raise ClothGPUError("clothgpu error %d: %s" % (-n, lib.clothgpu_last_error().decode()))
    nodes = np.ascontiguousarray(nodes[:n])
    return (nodes, int(nfb.value)) if return_fallbacks else nodes


def _ci(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _cd(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def build_adjacency(elem, nnoel, nnode, with_diagonal=False):
    """CJaggedArray::SetEdgeOfElem + Sort on the GPU (clothgpu_build_adjacency): (index[nnode+1], array[nnz]) with
    ascending unique columns -- the block-CRS pattern for the hinge quads, aEdge for the triangles."""
    lib = load()
    elem = _ci(elem).ravel()
    nelem = elem.size // nnoel
    index = np.zeros(nnode + 1, dtype=np.int32)
    n = lib.clothgpu_build_adjacency(C.c_int(nelem), C.c_int(nnoel), _ip(elem), C.c_int(nnode), C.c_int(int(with_diagonal)),
                                     _ip(index), None, C.c_int(0))
    if n < 0:
        raise ClothGPUError("clothgpu error %d: %s" % (-n, lib.clothgpu_last_error().decode()))
    array = np.zeros(max(n, 1), dtype=np.int32)
    n2 = lib.clothgpu_build_adjacency(C.c_int(nelem), C.c_int(nnoel), _ip(elem), C.c_int(nnode), C.c_int(int(with_diagonal)),
                                      _ip(index), _ip(array), C.c_int(array.size))
    if n2 != n:
        raise ClothGPUError("clothgpu error: %s" % lib.clothgpu_last_error().decode())
    return index, array[:n]


def _elem_call(fn, n, ins, outs, scalars=()):
    lib = load()
    rc = getattr(lib, fn)(C.c_int(n), *[_dp(a) for a in ins], *[C.c_double(v) for v in scalars], *[_dp(a) for a in outs])
    if rc != 0:
        raise ClothGPUError("clothgpu error %d: %s" % (rc, lib.clothgpu_last_error().decode()))
    return outs


def elem_psd_sym22(s2):
    """MakePositiveDefinite_Sim22 (cloth_internal_physics.h:15-67) for n stresses {s00, s11, s01}."""
    s2 = _cd(s2).reshape(-1, 3)
    return _elem_call("clothgpu_elem_psd_sym22", len(s2), [s2], [np.zeros_like(s2)])[0]


def elem_cst(C0, c, lam, myu):
    """WdWddW_CST (cloth_internal_physics.h:70-176) for n triangles: (W[n], dW[n,3,3], ddW[n,3,3,3,3])."""
    C0 = _cd(C0).reshape(-1, 9)
    c = _cd(c).reshape(-1, 9)
    n = len(C0)
    return _elem_call("clothgpu_elem_cst", n, [C0, c], [np.zeros(n), np.zeros((n, 3, 3)), np.zeros((n, 3, 3, 3, 3))],
                      (lam, myu))


def elem_bend(C0, c, stiff):
    """WdWddW_Bend (cloth_internal_physics.h:180-238) for n hinges: (W[n], dW[n,4,3], ddW[n,4,4,3,3])."""
    C0 = _cd(C0).reshape(-1, 12)
    c = _cd(c).reshape(-1, 12)
    n = len(C0)
    return _elem_call("clothgpu_elem_bend", n, [C0, c], [np.zeros(n), np.zeros((n, 4, 3)), np.zeros((n, 4, 4, 3, 3))],
                      (stiff,))


def elem_contact(pd_normal, stiff, clearance):
    """WdWddW_Contact (cloth_internal_physics.h:243-280) for n vertices given the callback's {pd, n}."""
    pdn = _cd(pd_normal).reshape(-1, 4)
    n = len(pdn)
    return _elem_call("clothgpu_elem_contact", n, [pdn], [np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3, 3))],
                      (stiff, clearance))


def dist_face_vertex(p):
    """DistanceFaceVertex (bvh_aabb.cpp:321-340): p[n,4,3] -> (dist, w0, w1)."""
    p = _cd(p).reshape(-1, 12)
    n = len(p)
    return _elem_call("clothgpu_dist_face_vertex", n, [p], [np.zeros(n), np.zeros(n), np.zeros(n)])


def dist_edge_edge(p):
    """DistanceEdgeEdge (bvh_aabb.cpp:343-384): p[n,4,3] = {p0,p1,q0,q1} -> (dist, ratio_p, ratio_q)."""
    p = _cd(p).reshape(-1, 12)
    n = len(p)
    return _elem_call("clothgpu_dist_edge_edge", n, [p], [np.zeros(n), np.zeros(n), np.zeros(n)])


def coplanar_interp(s, e):
    """FindCoplanerInterp (bvh_aabb.cpp:598-662): s, e [n,4,3] -> t[n]."""
    s = _cd(s).reshape(-1, 12)
    e = _cd(e).reshape(-1, 12)
    return _elem_call("clothgpu_coplanar_interp", len(s), [s, e], [np.zeros(len(s))])[0]


class ClothGPU:
    """A batch of `nscene` cloth scenes sharing one topology, resident on one GPU."""

    def __init__(self, sc, nscene=1, device=-1, with_bvh=True):
        self.lib = load()
        self.sc = sc
        self.nscene = int(nscene)
        self.np = int(sc["np"])
        self.ndof = self.np * 3
        self.ncrs = int(sc["rowptr"][-1])
        nodes = sc.get("bvh") if with_bvh else None
        self.nnode = 0 if nodes is None else int(nodes.shape[0])
        self._keep = [_ci(sc["tri"]), _ci(sc["quad"]), _ci(sc["bc"]), _cd(sc["xyz0"]), _ci(sc["rowptr"]),
                      _ci(sc["colidx"])]
        tri, quad, bc, xyz0, rowptr, colidx = self._keep
        h = C.c_void_p()
        if self.nnode:
            nodes = _ci(nodes)
            ei, ea = _ci(sc["edge_index"]), _ci(sc["edge_array"])
            args = (_ip(nodes), _ip(ei), _ip(ea))
        else:
            args = (None, None, None)
        rc = self.lib.clothgpu_create(C.byref(h), C.c_int(device), C.c_int(self.nscene), C.c_int(self.np),
                                      C.c_int(sc["ntri"]), C.c_int(sc["nquad"]), _ip(tri), _ip(quad), _ip(bc),
                                      _dp(xyz0), _ip(rowptr), _ip(colidx), C.c_int(self.nnode), *args)
        self.h = h
        self._host_allocs = []
        self._check(rc)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise ClothGPUError("clothgpu error %d: %s" % (rc, self.lib.clothgpu_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            for ptr in getattr(self, "_host_allocs", []):
                self.lib.clothgpu_free_host(ptr)
            self._host_allocs = []
            self.lib.clothgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def buf_len(self, buf):
        return int(self.lib.clothgpu_buf_len(self.h, buf))

    def write(self, buf, arr, scene=0):
        arr = _ci(arr) if buf in _INT_BUFS else _cd(arr)
        self._check(self.lib.clothgpu_write(self.h, buf, scene, arr.ctypes.data_as(C.c_void_p), arr.size))

    def read(self, buf, count=None, scene=0):
        n = self.buf_len(buf) if count is None else int(count)
        out = np.zeros(n, dtype=np.int32 if buf in _INT_BUFS else np.float64)
        self._check(self.lib.clothgpu_read(self.h, buf, scene, out.ctypes.data_as(C.c_void_p), n))
        return out

    def upload_state(self, xyz, uvw, scene=0):
        xyz, uvw = _cd(xyz), _cd(uvw)
        assert xyz.size == self.ndof and uvw.size == self.ndof
        self._check(self.lib.clothgpu_upload_state(self.h, scene, _dp(xyz), _dp(uvw)))

    def download_state(self, scene=0, out=None):
        """X, V of one scene; `out=(x, v)` fills caller-owned buffers (e.g. from `alloc_host`) instead of new ones."""
        if out is None:
            x = np.zeros(self.ndof)
            v = np.zeros(self.ndof)
        else:
            x, v = out
            assert x.size == self.ndof and v.size == self.ndof and x.dtype == np.float64 and v.dtype == np.float64
            assert x.flags.c_contiguous and v.flags.c_contiguous
        self._check(self.lib.clothgpu_download_state(self.h, scene, _dp(x), _dp(v)))
        return x, v

    def alloc_host(self, n):
        """n doubles of page-locked host memory as a numpy array (freed with the context)."""
        ptr = C.c_void_p()
        self._check(self.lib.clothgpu_alloc_host(C.byref(ptr), int(n) * 8))
        self._host_allocs.append(ptr)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(int(n),))

    def set_scene_gravity(self, scene, g):
        g = _cd(g)
        self._check(self.lib.clothgpu_set_scene_gravity(self.h, scene, _dp(g)))

    def synchronize(self):
        self._check(self.lib.clothgpu_synchronize(self.h))

    def timer_start(self):
        self._check(self.lib.clothgpu_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._check(self.lib.clothgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.clothgpu_launch_count(self.h))

    def set_profiling(self, on):
        self._check(self.lib.clothgpu_set_profiling(self.h, 1 if on else 0))

    def phase_ms(self):
        out = np.zeros(NPHASE)
        self._check(self.lib.clothgpu_phase_ms(self.h, _dp(out)))
        return dict(zip(PHASE_NAMES, out[:len(PHASE_NAMES)]))

    def set_kernel_timing(self, on):
        self._check(self.lib.clothgpu_set_kernel_timing(self.h, 1 if on else 0))

    def kernel_times(self):
        """{kernel name: (launches, total ms)} accumulated since set_kernel_timing(True)."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.clothgpu_kernel_times(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split()
            out[name] = (int(n), float(ms))
        return out

    # ------------------------------------------------------------------ implicit step
    def assemble(self, p):
        p = params_from(p)
        self._check(self.lib.clothgpu_assemble(self.h, C.byref(p)))

    def matvec(self, alpha, buf_x, beta, buf_y):
        self._check(self.lib.clothgpu_matvec(self.h, C.c_double(alpha), buf_x, C.c_double(beta), buf_y))

    def dot(self, buf_a, buf_b):
        out = np.zeros(self.nscene)
        self._check(self.lib.clothgpu_dot(self.h, buf_a, buf_b, _dp(out)))
        return out

    def axpy(self, a, buf_x, buf_y):
        self._check(self.lib.clothgpu_axpy(self.h, C.c_double(a), buf_x, buf_y))

    def ilu_setvalue(self):
        self._check(self.lib.clothgpu_ilu_setvalue(self.h))

    def ilu_factor(self):
        self._check(self.lib.clothgpu_ilu_factor(self.h))

    def ilu_solve(self, buf):
        self._check(self.lib.clothgpu_ilu_solve(self.h, buf))

    def _solver(self, fn, conv_ratio, iteration):
        cr = np.full(self.nscene, float(conv_ratio))
        it = np.full(self.nscene, int(iteration), dtype=np.int32)
        self._check(fn(self.h, _dp(cr), _ip(it)))
        return cr, it

    def pcg(self, conv_ratio=1.0e-4, iteration=100):
        return self._solver(self.lib.clothgpu_pcg, conv_ratio, iteration)

    def cg(self, conv_ratio=1.0e-4, iteration=1000):
        return self._solver(self.lib.clothgpu_cg, conv_ratio, iteration)

    def step_internal_ilu(self, p):
        p = params_from(p)
        cr = np.zeros(self.nscene)
        it = np.zeros(self.nscene, dtype=np.int32)
        self._check(self.lib.clothgpu_step_internal_ilu(self.h, C.byref(p), _dp(cr), _ip(it)))
        return cr, it

    def step_internal_cg(self, p):
        p = params_from(p)
        cr = np.zeros(self.nscene)
        it = np.zeros(self.nscene, dtype=np.int32)
        self._check(self.lib.clothgpu_step_internal_cg(self.h, C.byref(p), _dp(cr), _ip(it)))
        return cr, it

    def update_intermediate_velocity(self, p):
        p = params_from(p)
        self._check(self.lib.clothgpu_update_intermediate_velocity(self.h, C.byref(p)))

    # ------------------------------------------------------------------ self collision
    def bvh_refit_prx(self, buf_xyz, delta):
        self._check(self.lib.clothgpu_bvh_refit_prx(self.h, buf_xyz, C.c_double(delta)))

    def bvh_refit_ccd(self, buf_xyz, buf_uvw, dt):
        self._check(self.lib.clothgpu_bvh_refit_ccd(self.h, buf_xyz, buf_uvw, C.c_double(dt)))

    def contacts_prx(self, buf_xyz, delta):
        n = np.zeros(self.nscene, dtype=np.int32)
        self._check(self.lib.clothgpu_contacts_prx(self.h, buf_xyz, C.c_double(delta), _ip(n)))
        return n

    def contacts_ccd(self, buf_xyz, buf_uvw, dt, delta):
        n = np.zeros(self.nscene, dtype=np.int32)
        self._check(self.lib.clothgpu_contacts_ccd(self.h, buf_xyz, buf_uvw, C.c_double(dt), C.c_double(delta), _ip(n)))
        return n

    def elems(self, n, scene=0):
        return self.read(BUF_ELEMS, int(n) * 5, scene).reshape(-1, 5)

    def leaf_pairs(self, scene=0):
        n = np.zeros(self.nscene, dtype=np.int32)
        self._check(self.lib.clothgpu_leaf_pair_counts(self.h, _ip(n)))
        return self.read(BUF_LEAF_PAIRS, int(n[scene]) * 2, scene).reshape(-1, 2)

    def impulse_prx(self, buf_xyz, buf_uvw, delta, stiffness, dt, mass):
        self._check(self.lib.clothgpu_impulse_prx(self.h, buf_xyz, buf_uvw, C.c_double(delta), C.c_double(stiffness),
                                                  C.c_double(dt), C.c_double(mass)))

    def impulse_ccd(self, buf_xyz, buf_uvw, delta, stiffness, dt, mass):
        self._check(self.lib.clothgpu_impulse_ccd(self.h, buf_xyz, buf_uvw, C.c_double(delta), C.c_double(stiffness),
                                                  C.c_double(dt), C.c_double(mass)))

    def riz_reset(self):
        self._check(self.lib.clothgpu_riz_reset(self.h))

    def riz_make(self):
        self._check(self.lib.clothgpu_riz_make(self.h))

    def riz_apply(self, buf_xyz, buf_uvw):
        self._check(self.lib.clothgpu_riz_apply(self.h, buf_xyz, buf_uvw))

    def resolve_self_contact(self, p):
        p = params_from(p)
        applied = np.zeros(self.nscene, dtype=np.int32)
        stats = np.zeros((self.nscene, NSTATS), dtype=np.int32)
        self._check(self.lib.clothgpu_resolve_self_contact(self.h, C.byref(p), _ip(applied), _ip(stats)))
        return applied, stats

    # ------------------------------------------------------------------ whole steps
    def make_normal(self, buf_xyz=0, scene=0):
        """MakeNormal (self_contact_sparse/main.cpp:183-211): unit vertex normals, np x 3."""
        out = np.zeros(self.ndof)
        self._check(self.lib.clothgpu_make_normal(self.h, C.c_int(buf_xyz), C.c_int(scene), _dp(out)))
        return out

    def step_self_contact(self, p, nsteps=1):
        p = params_from(p)
        stats = np.zeros((self.nscene, NSTATS), dtype=np.int32)
        self._check(self.lib.clothgpu_step_self_contact(self.h, C.byref(p), int(nsteps), _ip(stats)))
        return stats

    def step_internal(self, p, nsteps=1):
        p = params_from(p)
        it = np.zeros(self.nscene, dtype=np.int32)
        self._check(self.lib.clothgpu_step_internal(self.h, C.byref(p), int(nsteps), _ip(it)))
        return it
