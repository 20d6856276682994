"""Make the reference's hot-path modules importable in THIS container (tooling for
golden-vector generation and the live oracle-vs-reference tests; never used by the
product, never available on the GPU box).

Recipe from SURVEY.md section 8(c): `timm` and `torchmetrics` are absent from the image;
the hot path only needs `timm.models.layers.{DropPath,to_2tuple,trunc_normal_}`
(DropPath is never instantiated) and an unused torchmetrics symbol.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GRAFP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "encoder", "graph_encoder.py"))


def install_stubs() -> None:
    import torch.nn as nn
    if "timm" not in sys.modules:
        layers = types.ModuleType("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, p=0.0):
                super().__init__()
                self.drop_prob = p

            def forward(self, x):
                return x

        layers.DropPath = DropPath
        layers.to_2tuple = lambda v: (v, v)
        layers.trunc_normal_ = nn.init.trunc_normal_
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        timm.models = models
        models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    if "torchmetrics" not in sys.modules:
        tmf = types.ModuleType("torchmetrics.functional")
        tmf.pairwise_cosine_similarity = lambda *a, **k: None
        tm = types.ModuleType("torchmetrics")
        tm.functional = tmf
        sys.modules.update({"torchmetrics": tm, "torchmetrics.functional": tmf})


def import_reference():
    """Returns a namespace with the reference classes/functions on the hot path."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import yaml
    from encoder.graph_encoder import GraphEncoder
    from encoder.gcn_lib.torch_vertex import Grapher, DyGraphConv2d, MRConv2d
    from encoder.gcn_lib.torch_edge import DenseDilatedKnnGraph, dense_knn_matrix
    from encoder.gcn_lib.torch_nn import batched_index_select, BasicConv
    from simclr.simclr import SimCLR
    from simclr.ntxent import ntxent_loss
    with open(os.path.join(REFERENCE_ROOT, "config", "grafp.yaml")) as f:
        cfg = yaml.safe_load(f)
    ns = types.SimpleNamespace(
        GraphEncoder=GraphEncoder, Grapher=Grapher, DyGraphConv2d=DyGraphConv2d,
        MRConv2d=MRConv2d, DenseDilatedKnnGraph=DenseDilatedKnnGraph,
        dense_knn_matrix=dense_knn_matrix, batched_index_select=batched_index_select,
        BasicConv=BasicConv, SimCLR=SimCLR, ntxent_loss=ntxent_loss, cfg=cfg)
    return ns


def import_reference_reranker():
    """CrossAttentionClassifier of the reference's downstream.py (lines 30-79).  The file itself cannot be imported
    here (tensorboard / DGL / dataset imports at module level), so only the class definition is compiled, from the
    reference source where it lies, into a namespace that provides torch / nn / F."""
    import ast
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    path = os.path.join(REFERENCE_ROOT, "downstream.py")
    tree = ast.parse(open(path).read(), path)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CrossAttentionClassifier"]
    if not cls:
        raise RuntimeError("CrossAttentionClassifier not found in %s" % path)
    ns = {"torch": torch, "nn": nn, "F": F}
    exec(compile(ast.Module(body=cls, type_ignores=[]), path, "exec"), ns)
    return ns["CrossAttentionClassifier"]
