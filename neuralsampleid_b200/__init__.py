"""neuralsampleid_b200 -- B200 (sm_100a) kernels behind NeuralSampleID's GraphEncoder hot path.

Public surface mirrors the reference's module API for this path (same class names,
constructor/forward signatures and state_dict keys):

    from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder
    from neuralsampleid_b200.encoder.gcn_lib.torch_vertex import Grapher, DyGraphConv2d, MRConv2d
    from neuralsampleid_b200.encoder.gcn_lib.torch_edge import DenseDilatedKnnGraph, dense_knn_matrix
    from neuralsampleid_b200.encoder.gcn_lib.torch_nn import batched_index_select, BasicConv
    from neuralsampleid_b200.simclr.simclr import SimCLR
    from neuralsampleid_b200.simclr.ntxent import ntxent_loss
"""
__version__ = "0.1.0"
