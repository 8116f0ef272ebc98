"""Host-side mirror of ``medpy.graphcut.write`` (reference: medpy/graphcut/write.py:27-76)."""

__all__ = ["graph_to_dimacs"]


def graph_to_dimacs(g, f):
    """Write the ``Graph`` description ``g`` to the file-like ``f`` as a DIMACS max-flow problem, line for line what
    the reference writes: node 1 is the source, node 2 the sink, graph node v becomes v + 2; zero weights are left
    out; the reverse arc of an edge follows it directly."""
    lines = ["c Created by medpy", "c Oskar Maier, oskar.maier@googlemail.com", "c",
             "c problem line", "p max {} {}".format(g.get_node_count() + 2, len(g.get_edges())),
             "c source descriptor", "n 1 s", "c sink descriptor", "n 2 t",
             "c terminal arcs (t-weights)"]
    for node, (to_source, to_sink) in list(g.get_tweights().items()):
        if not 0 == to_source:
            lines.append("a 1 {} {}".format(node + 2, to_source))
        if not 0 == to_sink:
            lines.append("a {} 2 {}".format(node + 2, to_sink))
    lines.append("c inter-node arcs (n-weights)")
    for (a, b), (there, back) in list(g.get_nweights().items()):
        if not 0 == there:
            lines.append("a {} {} {}".format(a + 2, b + 2, there))
        if not 0 == back:
            lines.append("a {} {} {}".format(b + 2, a + 2, back))
    f.write("\n".join(lines) + "\nc end-of-file")
