"""Shared helpers for the tests: load a golden fixture as (mesh, inputs, expected)."""
import os
from types import SimpleNamespace

import numpy as np

from tdgl_b200.mesh import EdgeMesh, Mesh
from tdgl_b200.synthetic import TerminalInfo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["film20_fixed", "film20_adaptive", "strip_transport"]
DYNAMIC_CASES = ["film20_ramp"]          # time-dependent vector potential


def load_case(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    em = EdgeMesh(g["centers"], g["edges"], g["boundary_edge_indices"], g["directions"],
                  g["edge_lengths"], g["dual_edge_lengths"])
    mesh = Mesh(g["sites"], g["elements"], g["boundary_indices"], areas=g["areas"],
                edge_mesh=em)
    terms = []
    currents = {}
    for k in range(int(g["n_terminals"])):
        name_k = str(g[f"term{k}_name"])
        terms.append(TerminalInfo(name_k, g[f"term{k}_sites"], g[f"term{k}_edges"],
                                  g[f"term{k}_bedges"], float(g[f"term{k}_length"])))
        currents[name_k] = float(g[f"term{k}_current"])
    opts = {k[4:]: g[k].item() for k in g.files if k.startswith("opt_")}
    max_steps = int(g["max_steps"])
    A_func = None
    if "ramp" in g.files:                  # (b_max, t_ramp), see oracle/make_golden.py
        from tdgl_b200.synthetic import uniform_field_vector_potential

        b_max, t_ramp = (float(v) for v in g["ramp"])
        A1 = uniform_field_vector_potential(g["centers"], 1.0)
        A_func = lambda t: min(max(t / t_ramp, 0.0), 1.0) * b_max * A1  # noqa: E731
    return SimpleNamespace(
        A_func=A_func,
        g=g, mesh=mesh, A=g["A_applied"], eps=g["epsilon"], u=float(g["u"]),
        gamma=float(g["gamma"]), terminals=tuple(terms), currents=currents, opts=opts,
        end_time=float(g["end_time"]), max_steps=None if max_steps < 0 else max_steps,
        probes=[int(i) for i in g["probe_points"]] if "probe_points" in g.files else None)
