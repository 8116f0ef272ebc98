"""oracle/solvers.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loaders for the two CPU max-flow checkers:

* ``solve_port``  -- oracle/bk_lattice.c, our C restatement of the reference BK solver
                     (lib/maxflow/src/maxflow.cpp:118-604) on an implicit lattice;
* ``solve_ref``   -- oracle/_ref/libbkref.so, the *unmodified* reference solver compiled from
                     /root/reference by oracle/Makefile (present only where that build ran; the
                     built .so travels to the GPU box).

Both take the dict produced by ``oracle.energy_terms.build_problem`` and return
``(flow, mask uint8[shape], times)``.  Only tests/, bench.py's cpu_baseline / ``--impl reference`` arm
and __graft_entry__.smoke() may import this module.
"""
import ctypes
import os
import subprocess

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "_build", "libbk_lattice.so")
_REF_SO = os.path.join(_HERE, "_ref", "libbkref.so")

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_u8_p = ctypes.POINTER(ctypes.c_uint8)
_c_i64_p = ctypes.POINTER(ctypes.c_int64)


def build(ref=True):
    """Compile the checkers (port always; the reference build only when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    if ref and os.path.isdir("/root/reference/lib/maxflow/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def have_ref():
    return os.path.exists(_REF_SO)


_port = None
_ref = None


def _load_port():
    global _port
    if _port is None:
        if not os.path.exists(_PORT_SO):
            build(ref=False)
        lib = ctypes.CDLL(_PORT_SO)
        lib.bk_lattice_solve.restype = ctypes.c_int
        lib.bk_lattice_solve.argtypes = [ctypes.c_int, _c_i64_p, ctypes.POINTER(_c_double_p),
                                         ctypes.POINTER(_c_double_p), _c_double_p, ctypes.c_double,
                                         _c_u8_p, _c_double_p, _c_double_p]
        _port = lib
    return _port


def _load_ref():
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(_REF_SO)
        lib.bkref_lattice_solve.restype = ctypes.c_int
        lib.bkref_lattice_solve.argtypes = [ctypes.c_int, _c_i64_p, ctypes.POINTER(_c_double_p),
                                            ctypes.POINTER(_c_double_p), _c_double_p, _c_double_p,
                                            _c_u8_p, _c_u8_p, ctypes.c_int, _c_u8_p, _c_double_p,
                                            _c_double_p]
        _ref = lib
    return _ref


def _dptr(a):
    return a.ctypes.data_as(_c_double_p) if a is not None else _c_double_p()


def _axis_ptrs(arrs, ndim):
    P = _c_double_p * ndim
    if arrs is None:
        return P(*[_c_double_p() for _ in range(ndim)]), []
    keep = [numpy.ascontiguousarray(a, dtype=numpy.float64) for a in arrs]
    return P(*[_dptr(a) for a in keep]), keep


def solve_port(prob):
    """BK restatement (oracle/bk_lattice.c) on a build_problem() dict."""
    lib = _load_port()
    shape = tuple(int(s) for s in prob["shape"])
    ndim = len(shape)
    n = int(numpy.prod(shape))
    shp = (ctypes.c_int64 * ndim)(*shape)
    wf, k1 = _axis_ptrs(prob["wf"], ndim)
    wb, k2 = _axis_ptrs(prob["wb"], ndim)
    tr = numpy.ascontiguousarray(prob["tr"], dtype=numpy.float64)
    mask = numpy.empty(n, dtype=numpy.uint8)
    flow = ctypes.c_double(0)
    times = (ctypes.c_double * 2)()
    rc = lib.bk_lattice_solve(ndim, shp, wf, wb, _dptr(tr), float(prob["flow_const"]),
                              mask.ctypes.data_as(_c_u8_p), ctypes.byref(flow), times)
    if rc != 0:
        raise RuntimeError("bk_lattice_solve failed (%d)" % rc)
    return flow.value, mask.reshape(shape), {"setup_s": times[0], "maxflow_s": times[1]}


def solve_ref(prob, use_sum_edge=False):
    """The real reference BK (oracle/_ref/libbkref.so), replaying the reference call sequence
    regional -> boundary -> fg -> bg (generate.py:159-172)."""
    lib = _load_ref()
    shape = tuple(int(s) for s in prob["shape"])
    ndim = len(shape)
    n = int(numpy.prod(shape))
    shp = (ctypes.c_int64 * ndim)(*shape)
    wf, k1 = _axis_ptrs(prob["wf"], ndim)
    wb, k2 = _axis_ptrs(prob["wb"], ndim)
    src = prob.get("src")
    snk = prob.get("snk")
    if src is not None:
        src = numpy.ascontiguousarray(src, dtype=numpy.float64)
        snk = numpy.ascontiguousarray(snk, dtype=numpy.float64)
    fg = numpy.ascontiguousarray(prob["fg"], dtype=numpy.uint8)
    bg = numpy.ascontiguousarray(prob["bg"], dtype=numpy.uint8)
    mask = numpy.empty(n, dtype=numpy.uint8)
    flow = ctypes.c_double(0)
    times = (ctypes.c_double * 3)()
    rc = lib.bkref_lattice_solve(ndim, shp, wf, wb, _dptr(src), _dptr(snk),
                                 fg.ctypes.data_as(_c_u8_p) if fg.any() else _c_u8_p(),
                                 bg.ctypes.data_as(_c_u8_p) if bg.any() else _c_u8_p(),
                                 1 if use_sum_edge else 0,
                                 mask.ctypes.data_as(_c_u8_p), ctypes.byref(flow), times)
    if rc != 0:
        raise RuntimeError("bkref_lattice_solve: instance exceeds the reference's int32 ids (graph.h:62,82)")
    return flow.value, mask.reshape(shape), {"fill_s": times[0], "maxflow_s": times[1], "readout_s": times[2]}


_SPARSE_PORT_SO = os.path.join(_HERE, "_build", "libbk_sparse.so")
_sparse_port = None


def solve_sparse_port(n, i, j, cap, rev, tw_ops):
    """BK restatement for general graphs (oracle/bk_sparse.c); same contract as solve_sparse_ref."""
    global _sparse_port
    if _sparse_port is None:
        if not os.path.exists(_SPARSE_PORT_SO) or os.path.getmtime(_SPARSE_PORT_SO) < os.path.getmtime(os.path.join(_HERE, "bk_sparse.c")):
            subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
        _sparse_port = ctypes.CDLL(_SPARSE_PORT_SO)
    return _solve_sparse(_sparse_port.bk_sparse_solve, n, i, j, cap, rev, tw_ops)


def solve_sparse(n, i, j, cap, rev, tw_ops):
    """The real reference solver where it is built, otherwise the restatement (they agree bit for bit)."""
    return (solve_sparse_ref if have_ref() else solve_sparse_port)(n, i, j, cap, rev, tw_ops)


def solve_sparse_ref(n, i, j, cap, rev, tw_ops):
    """The real reference BK on a general sparse graph: `tw_ops` = sequence of (nodes, src, snk) array triples replayed
    as add_tweights calls in order, then sum_edge(i[k], j[k], cap[k], rev[k]) in order (oracle/ref_driver.cpp).
    Returns (flow, mask uint8[n] with 1 = not SINK, maxflow seconds)."""
    return _solve_sparse(_load_ref().bkref_sparse_solve, n, i, j, cap, rev, tw_ops)


def _solve_sparse(fn, n, i, j, cap, rev, tw_ops):
    c_i32_p = ctypes.POINTER(ctypes.c_int32)
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int64, c_i32_p, c_i32_p, _c_double_p, _c_double_p,
                                       ctypes.c_int64, c_i32_p, _c_double_p, _c_double_p, _c_u8_p, _c_double_p,
                                       _c_double_p]
    ei = numpy.ascontiguousarray(i, dtype=numpy.int32)
    ej = numpy.ascontiguousarray(j, dtype=numpy.int32)
    ec = numpy.ascontiguousarray(cap, dtype=numpy.float64)
    er = numpy.ascontiguousarray(rev, dtype=numpy.float64)
    if tw_ops:
        tn = numpy.ascontiguousarray(numpy.concatenate([numpy.asarray(o[0]).ravel() for o in tw_ops]), dtype=numpy.int32)
        ts = numpy.ascontiguousarray(numpy.concatenate([numpy.broadcast_to(numpy.asarray(o[1], dtype=numpy.float64), numpy.asarray(o[0]).shape).ravel() for o in tw_ops]))
        tk = numpy.ascontiguousarray(numpy.concatenate([numpy.broadcast_to(numpy.asarray(o[2], dtype=numpy.float64), numpy.asarray(o[0]).shape).ravel() for o in tw_ops]))
    else:
        tn = numpy.zeros(0, numpy.int32)
        ts = tk = numpy.zeros(0)
    mask = numpy.empty(int(n), dtype=numpy.uint8)
    flow = ctypes.c_double(0)
    secs = ctypes.c_double(0)
    rc = fn(int(n), ei.size, ei.ctypes.data_as(c_i32_p), ej.ctypes.data_as(c_i32_p), _dptr(ec), _dptr(er),
                                tn.size, tn.ctypes.data_as(c_i32_p), _dptr(ts), _dptr(tk),
                                mask.ctypes.data_as(_c_u8_p), ctypes.byref(flow), ctypes.byref(secs))
    if rc != 0:
        raise RuntimeError("sparse BK solve failed (%d)" % rc)
    return flow.value, mask, secs.value
