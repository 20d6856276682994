from .torch_nn import BasicConv, act_layer, batched_index_select, norm_layer
from .torch_edge import DenseDilated, DenseDilatedKnnGraph, dense_knn_matrix, pairwise_distance
from .torch_vertex import DyGraphConv2d, GraphConv2d, Grapher, MRConv2d
