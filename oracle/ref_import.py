"""ORACLE tooling — imports the UNMODIFIED reference modules.

Root: ``$FMM_REFERENCE_ROOT``, else ``/root/reference`` (the build container), else ``oracle/_ref`` — the byte-for-byte staged
copy ``oracle/build_ref.py`` makes of the hot-path files (git-ignored; it travels to the GPU box with the snapshot so the
reference arm of ``bench.py`` runs the reference's own code there). Used by ``oracle/make_golden.py`` to generate the
committed fixtures, by the cross-checks in ``tests/test_oracle.py`` and by ``bench.py --impl reference`` / its
``cpu_baseline`` leg. Import recipe: SURVEY.md section 8(c).
"""
from __future__ import annotations

import importlib
import importlib.util
import json
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF = os.environ.get("FMM_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/Fall_2_Spatial_Temporal_SR")
                                                else _STAGED)
F2 = os.path.join(REF, "Fall_2_Spatial_Temporal_SR")


def available() -> bool:
    return os.path.isdir(F2)


def load_gstcan():
    """Returns (stgcan, graph, bilstm, combination) reference modules."""
    if F2 not in sys.path:
        sys.path.insert(0, F2)
    stg = importlib.import_module("Model.stgcan")
    g = importlib.import_module("Model.graph")
    bl = importlib.import_module("Model.bilstm")
    shim = types.ModuleType("Model.st_gcn")
    shim.__path__ = []
    sys.modules.update({"Model.st_gcn": shim, "Model.st_gcn.stgcan": stg, "Model.st_gcn.graph": g})
    comb = importlib.import_module("Model.combination")
    return stg, g, bl, comb


def install_layout(stg, g, name, num_node, neighbor_link, center):
    """Teach the reference Graph an extra layout by subclassing it."""
    base = g.Graph

    class GraphX(base):
        def get_edge(self, layout):
            if layout == name:
                self.num_node = num_node
                self.edge = [(i, i) for i in range(num_node)] + list(neighbor_link)
                self.center = center
            else:
                base.get_edge(self, layout)

    stg.Graph = GraphX
    return GraphX


def load_notebook_sensor():
    """exec() the sensor-branch cell of the notebook: CNN1D, BiLSTM, CNN_BiLSTM classes."""
    nb = json.load(open(os.path.join(REF, "GSTCAN_HAR_conv_10kfold.ipynb")))
    ns: dict = {}
    exec("".join(nb["cells"][2]["source"]), ns)
    return ns


def load_tragcn(seq_len: int):
    """Root-level TRAGCN family loaded as a synthetic package; ``Transform`` /
    ``PositionalEncoding`` hard-code 30 frames as a default argument (D5), re-pointed at ``seq_len``."""
    if "TRAGCN" not in sys.modules:
        pkg = types.ModuleType("TRAGCN")
        pkg.__path__ = []
        sys.modules["TRAGCN"] = pkg
        for name in ("EmbGCN", "GRU", "TA", "TRAGCN"):
            spec = importlib.util.spec_from_file_location(f"TRAGCN.{name}", os.path.join(REF, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"TRAGCN.{name}"] = mod
            spec.loader.exec_module(mod)
    ta = sys.modules["TRAGCN.TA"]
    ta.Transform.__init__.__defaults__ = (seq_len,)
    ta.PositionalEncoding.__init__.__defaults__ = (seq_len,)
    return sys.modules["TRAGCN.TRAGCN"]


def load_musa():
    """Multimodal_Fall3/model/musa_model.py (imports cleanly: torch + numpy only)."""
    spec = importlib.util.spec_from_file_location("musa_model", os.path.join(REF, "Multimodal_Fall3", "model", "musa_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
