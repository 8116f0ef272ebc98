// oracle/ref_pyshim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// pybind11 stand-in for the reference's Boost.Python binding (lib/maxflow/src/wrapper.cpp:59-89,
// 125-134) so that the *unmodified* reference Python package under /root/reference can be imported
// in this container to generate golden vectors (tests/golden/make_golden.py).  Boost.Python is not
// installed here and the shipped Pythongraph wrapper lacks `return` statements
// (pythongraph.h:20-21), so Graph<double,double,double> is bound directly.
// Built by oracle/Makefile into oracle/_ref/maxflow<EXT_SUFFIX>; never loaded by product code.
#include <pybind11/pybind11.h>
#include "graph.h"

namespace py = pybind11;
typedef Graph<double, double, double> GD;

PYBIND11_MODULE(maxflow, m)
{
    py::class_<GD> cls(m, "GraphDouble");
    cls.def(py::init([](int n, int e) { return new GD(n, e, NULL); }))
        .def("add_node", [](GD& g, int n) { return g.add_node(n); })
        .def("add_edge", &GD::add_edge)
        .def("sum_edge", &GD::sum_edge)
        .def("add_tweights", &GD::add_tweights)
        .def("maxflow", [](GD& g) { return g.maxflow(); })
        .def("what_segment", [](GD& g, int i) { return g.what_segment(i); })
        .def("reset", &GD::reset)
        .def("get_edge", &GD::get_edge)
        .def("get_node_num", &GD::get_node_num)
        .def("get_arc_num", &GD::get_arc_num)
        .def("get_trcap", &GD::get_trcap);
    py::enum_<GD::termtype>(cls, "termtype")
        .value("SOURCE", GD::SOURCE)
        .value("SINK", GD::SINK);
    // medpy/graphcut/__init__.py:206-208 imports all three names.
    m.attr("GraphFloat") = cls;
    m.attr("GraphInt") = cls;
}
