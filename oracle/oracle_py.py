"""ctypes front-end for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``load_ref()``    -> oracle/_ref/libclothref.so   (the unmodified reference, symbols ``ref_*``)
* ``load_oracle()`` -> oracle/_build/liboracle.so   (the plain-C restatement, symbols ``orc_*``)

Both export the same flat interface (see oracle/ref_shim.cpp / oracle/cloth_oracle.h), so a test can run the
same call against either.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libclothref.so")
ORC_SO = os.path.join(HERE, "_build", "liboracle.so")

NSTATS = 128

_i = C.c_int
_d = C.c_double
_vp = C.c_void_p


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ci(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _cd(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(ref=True):
    """(Re)build the checkers with oracle/Makefile. `make ref` is a no-op when /root/reference is absent."""
    subprocess.check_call(["make", "-s", "-C", HERE, "_build/liboracle.so"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Params:
    """Scene constants as the demos define them (self_contact_sparse/main.cpp:138-156)."""

    def __init__(self, dt, lam, myu, stiff_bend, gravity, mass_point, stiff_contact, clearance,
                 col_kind=0, col_par=(0, 0, 0, 0)):
        self.dt = float(dt)
        self.lam = float(lam)
        self.myu = float(myu)
        self.stiff_bend = float(stiff_bend)
        self.gravity = _cd(gravity)
        self.mass_point = float(mass_point)
        self.stiff_contact = float(stiff_contact)
        self.clearance = float(clearance)
        self.col_kind = int(col_kind)
        self.col_par = _cd(col_par)


class OracleLib:
    def __init__(self, path, prefix):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        self.kind = "reference" if prefix == "ref_" else "port"

    def _f(self, name, restype=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    def has(self, name):
        return hasattr(self.lib, self.prefix + name)

    # ---------------------------------------------------------------- setup
    def bvh_build(self, tri, xyz):
        tri = _ci(tri)
        xyz = _cd(xyz)
        ntri = tri.size // 3
        nodes = np.zeros((2 * ntri - 1, 3), dtype=np.int32)
        n = self._f("bvh_build", _i)(_i(ntri), _ip(tri), _i(xyz.size // 3), _dp(xyz), _ip(nodes), _i(nodes.shape[0]))
        assert n == 2 * ntri - 1, (n, ntri)
        return nodes

    def edge_of_elem(self, elem, nnoel, nnode, sort):
        elem = _ci(elem)
        nelem = elem.size // nnoel
        cap = nelem * nnoel * nnoel
        index = np.zeros(nnode + 1, dtype=np.int32)
        array = np.zeros(cap, dtype=np.int32)
        nnz = self._f("edge_of_elem", _i)(_ip(elem), _i(nelem), _i(nnoel), _i(nnode), _i(1 if sort else 0),
                                         _ip(index), _ip(array), _i(cap))
        return index, array[:nnz].copy()

    def ilu_diaind(self, rowptr, colidx):
        np_ = rowptr.size - 1
        out = np.zeros(np_, dtype=np.int32)
        self._f("ilu_diaind")(_i(np_), _ip(_ci(rowptr)), _ip(_ci(colidx)), _ip(out))
        return out

    # ---------------------------------------------------------------- element energies
    def cst(self, C0, c, lam, myu):
        C0 = _cd(C0); c = _cd(c)
        W = _d(0)
        dW = np.zeros(9); ddW = np.zeros(81)
        self._f("cst")(_dp(C0), _dp(c), _d(lam), _d(myu), C.byref(W), _dp(dW), _dp(ddW))
        return W.value, dW.reshape(3, 3), ddW.reshape(3, 3, 3, 3)

    def bend(self, C0, c, stiff):
        C0 = _cd(C0); c = _cd(c)
        W = _d(0)
        dW = np.zeros(12); ddW = np.zeros(144)
        self._f("bend")(_dp(C0), _dp(c), _d(stiff), C.byref(W), _dp(dW), _dp(ddW))
        return W.value, dW.reshape(4, 3), ddW.reshape(4, 4, 3, 3)

    def contact(self, c, stiff, clearance, col_kind, col_par):
        c = _cd(c); col_par = _cd(col_par)
        W = _d(0)
        dW = np.zeros(3); ddW = np.zeros(9)
        self._f("contact")(_dp(c), _d(stiff), _d(clearance), _i(col_kind), _dp(col_par), C.byref(W), _dp(dW), _dp(ddW))
        return W.value, dW, ddW.reshape(3, 3)

    # ---------------------------------------------------------------- implicit step pieces
    def assemble(self, sc, xyz, uvw, p):
        """sc: scene dict (tri, quad, bc, xyz0, rowptr, colidx). Returns (valDia, valCrs, b)."""
        np_ = sc["np"]
        ncrs = int(sc["rowptr"][-1])
        dia = np.zeros(np_ * 9); crs = np.zeros(ncrs * 9); b = np.zeros(np_ * 3)
        xyz = _cd(xyz); uvw = _cd(uvw)
        self._f("assemble")(_i(np_), _i(sc["ntri"]), _i(sc["nquad"]), _ip(sc["tri"]), _ip(sc["quad"]), _ip(sc["bc"]),
                            _dp(sc["xyz0"]), _dp(xyz), _dp(uvw), _ip(sc["rowptr"]), _ip(sc["colidx"]),
                            _d(p.dt), _d(p.lam), _d(p.myu), _d(p.stiff_bend), _dp(p.gravity), _d(p.mass_point),
                            _d(p.stiff_contact), _d(p.clearance), _i(p.col_kind), _dp(p.col_par),
                            _dp(dia), _dp(crs), _dp(b))
        return dia, crs, b

    def matvec(self, sc, dia, crs, alpha, x, beta, y):
        y = _cd(y).copy()
        self._f("matvec")(_i(sc["np"]), _ip(sc["rowptr"]), _ip(sc["colidx"]), _dp(_cd(dia)), _dp(_cd(crs)),
                          _d(alpha), _dp(_cd(x)), _d(beta), _dp(y))
        return y

    def dot(self, a, b):
        a = _cd(a); b = _cd(b)
        return self._f("dot", _d)(_i(a.size), _dp(a), _dp(b))

    def ilu_factor(self, sc, dia, crs):
        dia = _cd(dia).copy(); crs = _cd(crs).copy()
        self._f("ilu_factor")(_i(sc["np"]), _ip(sc["rowptr"]), _ip(sc["colidx"]), _dp(dia), _dp(crs))
        return dia, crs

    def ilu_solve(self, sc, idia, icrs, vec):
        vec = _cd(vec).copy()
        self._f("ilu_solve")(_i(sc["np"]), _ip(sc["rowptr"]), _ip(sc["colidx"]), _dp(_cd(idia)), _dp(_cd(icrs)), _dp(vec))
        return vec

    def pcg(self, sc, dia, crs, idia, icrs, r, conv_ratio=1.0e-4, iteration=100):
        r = _cd(r).copy()
        x = np.zeros_like(r)
        cr = _d(conv_ratio); it = _i(iteration)
        self._f("pcg")(_i(sc["np"]), _ip(sc["rowptr"]), _ip(sc["colidx"]), _dp(_cd(dia)), _dp(_cd(crs)),
                       _dp(_cd(idia)), _dp(_cd(icrs)), _dp(r), _dp(x), C.byref(cr), C.byref(it))
        return x, r, cr.value, it.value

    def cg(self, sc, dia, crs, r, conv_ratio=1.0e-4, iteration=1000):
        r = _cd(r).copy()
        x = np.zeros_like(r)
        cr = _d(conv_ratio); it = _i(iteration)
        self._f("cg")(_i(sc["np"]), _ip(sc["rowptr"]), _ip(sc["colidx"]), _dp(_cd(dia)), _dp(_cd(crs)),
                      _dp(r), _dp(x), C.byref(cr), C.byref(it))
        return x, r, cr.value, it.value

    def make_normal(self, sc, xyz):
        """MakeNormal (self_contact_sparse/main.cpp:183-211); only in the C restatement (the reference keeps it in its
        GLUT demo source)."""
        out = np.zeros(sc["np"] * 3)
        self._f("make_normal")(_i(sc["np"]), _i(sc["ntri"]), _ip(_ci(sc["tri"])), _dp(_cd(xyz)), _dp(out))
        return out

    def update_intermediate_velocity(self, sc, xyz, uvw, p):
        uvw = _cd(uvw).copy()
        self._f("update_intermediate_velocity")(
            _i(sc["np"]), _i(sc["ntri"]), _i(sc["nquad"]), _ip(sc["tri"]), _ip(sc["quad"]), _ip(sc["bc"]),
            _dp(sc["xyz0"]), _dp(_cd(xyz)), _dp(uvw), _d(p.dt), _d(p.lam), _d(p.myu), _d(p.stiff_bend),
            _dp(p.gravity), _d(p.mass_point))
        return uvw

    # ---------------------------------------------------------------- BVH / detection
    def refit_prx(self, sc, xyz, delta):
        nodes = sc["bvh"]
        boxes = np.zeros((nodes.shape[0], 6))
        self._f("refit_prx")(_i(nodes.shape[0]), _ip(nodes), _i(sc["ntri"]), _ip(sc["tri"]), _i(sc["np"]),
                             _dp(_cd(xyz)), _d(delta), _dp(boxes))
        return boxes

    def refit_ccd(self, sc, xyz, uvw, dt):
        nodes = sc["bvh"]
        boxes = np.zeros((nodes.shape[0], 6))
        self._f("refit_ccd")(_i(nodes.shape[0]), _ip(nodes), _i(sc["ntri"]), _ip(sc["tri"]), _i(sc["np"]),
                             _dp(_cd(xyz)), _dp(_cd(uvw)), _d(dt), _dp(boxes))
        return boxes

    def leaf_pairs(self, sc, boxes, cap=None):
        nodes = sc["bvh"]
        cap = cap or 64 * sc["ntri"]
        out = np.zeros((cap, 2), dtype=np.int32)
        n = self._f("leaf_pairs", _i)(_i(nodes.shape[0]), _ip(nodes), _dp(_cd(boxes)), _ip(out), _i(cap))
        assert n <= cap
        return out[:n].copy()

    def contacts_prx(self, sc, boxes, xyz, delta, cap=1 << 20):
        nodes = sc["bvh"]
        out = np.zeros((cap, 5), dtype=np.int32)
        n = self._f("contacts_prx", _i)(_i(nodes.shape[0]), _ip(nodes), _dp(_cd(boxes)), _i(sc["ntri"]), _ip(sc["tri"]),
                                        _i(sc["np"]), _dp(_cd(xyz)), _d(delta), _ip(out), _i(cap))
        assert n <= cap
        return out[:n].copy()

    def contacts_ccd(self, sc, boxes, xyz, uvw, dt, delta, cap=1 << 20):
        nodes = sc["bvh"]
        out = np.zeros((cap, 5), dtype=np.int32)
        n = self._f("contacts_ccd", _i)(_i(nodes.shape[0]), _ip(nodes), _dp(_cd(boxes)), _i(sc["ntri"]), _ip(sc["tri"]),
                                        _i(sc["np"]), _dp(_cd(xyz)), _dp(_cd(uvw)), _d(dt), _d(delta), _ip(out), _i(cap))
        assert n <= cap
        return out[:n].copy()

    def dist_fv(self, p):
        w0 = _d(0); w1 = _d(0)
        d = self._f("dist_fv", _d)(_dp(_cd(p)), C.byref(w0), C.byref(w1))
        return d, w0.value, w1.value

    def dist_ee(self, p):
        w0 = _d(0); w1 = _d(0)
        d = self._f("dist_ee", _d)(_dp(_cd(p)), C.byref(w0), C.byref(w1))
        return d, w0.value, w1.value

    def coplanar(self, s, e):
        return self._f("coplanar", _d)(_dp(_cd(s)), _dp(_cd(e)))

    def is_fv_prx(self, ids, p, box, delta):
        return self._f("is_fv_prx", _i)(_ip(_ci(ids)), _dp(_cd(p)), _dp(_cd(box)), _d(delta))

    def is_ee_prx(self, ids, p, delta):
        return self._f("is_ee_prx", _i)(_ip(_ci(ids)), _dp(_cd(p)), _d(delta))

    def is_fv_ccd(self, ids, p, q, box):
        return self._f("is_fv_ccd", _i)(_ip(_ci(ids)), _dp(_cd(p)), _dp(_cd(q)), _dp(_cd(box)))

    def is_ee_ccd(self, ids, p, q):
        return self._f("is_ee_ccd", _i)(_ip(_ci(ids)), _dp(_cd(p)), _dp(_cd(q)))

    # ---------------------------------------------------------------- response
    def impulse_prx(self, uvw, delta, stiffness, dt, mass, xyz, elems):
        uvw = _cd(uvw).copy(); elems = _ci(elems)
        self._f("impulse_prx")(_i(uvw.size // 3), _dp(uvw), _d(delta), _d(stiffness), _d(dt), _d(mass), _dp(_cd(xyz)),
                               _i(elems.size // 5), _ip(elems))
        return uvw

    def impulse_ccd(self, uvw, delta, stiffness, dt, mass, xyz, elems):
        uvw = _cd(uvw).copy(); elems = _ci(elems)
        self._f("impulse_ccd")(_i(uvw.size // 3), _dp(uvw), _d(delta), _d(stiffness), _d(dt), _d(mass), _dp(_cd(xyz)),
                               _i(elems.size // 5), _ip(elems))
        return uvw

    def riz_new(self):
        return RIZ(self)

    # ---------------------------------------------------------------- whole-scene stepping
    def sim(self, sc):
        return Sim(self, sc)


class RIZ:
    def __init__(self, lib):
        self.lib = lib
        self.h = _vp(lib._f("riz_new", _vp)())

    def make(self, elems, sc):
        elems = _ci(elems)
        self.lib._f("riz_make")(self.h, _i(elems.size // 5), _ip(elems), _i(sc["np"]), _ip(sc["edge_index"]),
                                _ip(sc["edge_array"]))

    def zones(self, np_):
        out = np.zeros(np_, dtype=np.int32)
        n = self.lib._f("riz_get", _i)(self.h, _i(np_), _ip(out))
        return n, out

    def apply(self, uvw, xyz, uvw0):
        uvw = _cd(uvw).copy()
        self.lib._f("riz_apply")(self.h, _i(uvw.size // 3), _dp(uvw), _dp(_cd(xyz)), _dp(_cd(uvw0)))
        return uvw

    def __del__(self):
        try:
            self.lib._f("riz_free")(self.h)
        except Exception:
            pass


class Sim:
    """Persistent reference/oracle scene: the init sequence of self_contact_sparse/main.cpp:561-583."""

    def __init__(self, lib, sc):
        self.lib = lib
        self.sc = sc
        nodes = sc.get("bvh")
        nnode = 0 if nodes is None else nodes.shape[0]
        self.h = _vp(lib._f("sim_new", _vp)(
            _i(sc["np"]), _i(sc["ntri"]), _i(sc["nquad"]), _ip(sc["tri"]), _ip(sc["quad"]), _ip(sc["bc"]),
            _dp(sc["xyz0"]), _i(nnode), _ip(nodes) if nnode else None))

    def set_state(self, xyz, uvw):
        self.lib._f("sim_set_state")(self.h, _dp(_cd(xyz)), _dp(_cd(uvw)))

    def get_state(self):
        n = self.sc["np"] * 3
        x = np.zeros(n); v = np.zeros(n)
        self.lib._f("sim_get_state")(self.h, _dp(x), _dp(v))
        return x, v

    def step_internal(self, p):
        cr = _d(0); it = _i(0)
        self.lib._f("sim_step_internal")(self.h, _d(p.dt), _d(p.lam), _d(p.myu), _d(p.stiff_bend), _dp(p.gravity),
                                         _d(p.mass_point), _d(p.stiff_contact), _d(p.clearance), _i(p.col_kind),
                                         _dp(p.col_par), C.byref(cr), C.byref(it))
        return cr.value, it.value

    def step_internal_cg(self, p):
        """StepTime_InternalDynamics (solve_internal_sparse.h:185-253)"""
        cr = _d(0); it = _i(0)
        self.lib._f("sim_step_internal_cg")(self.h, _d(p.dt), _d(p.lam), _d(p.myu), _d(p.stiff_bend), _dp(p.gravity),
                                            _d(p.mass_point), _d(p.stiff_contact), _d(p.clearance), _i(p.col_kind),
                                            _dp(p.col_par), C.byref(cr), C.byref(it))
        return cr.value, it.value

    def resolve_contacts(self, uvw, xyz_prev, p):
        uvw = _cd(uvw).copy()
        stats = np.zeros(NSTATS, dtype=np.int32)
        applied = self.lib._f("sim_resolve_contacts", _i)(self.h, _dp(uvw), _dp(_cd(xyz_prev)), _d(p.dt),
                                                          _d(p.clearance), _d(p.mass_point), _d(p.stiff_contact),
                                                          _ip(stats))
        return uvw, bool(applied), stats

    def step_self_contact(self, p):
        stats = np.zeros(NSTATS, dtype=np.int32)
        self.lib._f("sim_step_self_contact")(self.h, _d(p.dt), _d(p.lam), _d(p.myu), _d(p.stiff_bend), _dp(p.gravity),
                                             _d(p.mass_point), _d(p.stiff_contact), _d(p.clearance), _i(p.col_kind),
                                             _dp(p.col_par), _ip(stats))
        return stats

    def __del__(self):
        try:
            self.lib._f("sim_free")(self.h)
        except Exception:
            pass


def load_ref():
    if not os.path.exists(REF_SO):
        return None
    return OracleLib(REF_SO, "ref_")


def load_oracle():
    if not os.path.exists(ORC_SO):
        build(ref=False)
    return OracleLib(ORC_SO, "orc_")
