"""fall_multimodal_b200 — B200-native (sm_100a) training/inference step of the ST-GCN fall/HAR
classifiers of musaru/Fall_Multimodal, behind the reference's PyTorch module API."""
from .graph import Graph, register_layout  # noqa: F401
from .stgcan import STGCAN, Channel_Attention, GraphConvolution, st_gcan  # noqa: F401
from .sensor import CNN1D, BiLSTM, ChannelAttention, CNN_BiLSTM  # noqa: F401
from .tragcn import TARGCN  # noqa: F401
from .fusion import ThreeStreamSTGCAN, TwoStreamSTGCAN, TwoStreamSTGCAN_CNN1D, TwoStreamSTGCAN_BiLSTM  # noqa: F401
from .notebook import StreamSpatialTemporalGraph, TwoStreamSpatialTemporalGraph  # noqa: F401
from .build_model import build_model  # noqa: F401
from .head import linear_cross_entropy  # noqa: F401
from .optim import FusedRMSprop  # noqa: F401

__all__ = ["Graph", "register_layout", "STGCAN", "st_gcan", "GraphConvolution", "Channel_Attention", "CNN1D",
           "TwoStreamSTGCAN", "TwoStreamSTGCAN_CNN1D", "TwoStreamSTGCAN_BiLSTM", "ThreeStreamSTGCAN", "BiLSTM", "ChannelAttention", "CNN_BiLSTM", "TARGCN",
           "StreamSpatialTemporalGraph", "TwoStreamSpatialTemporalGraph", "build_model", "linear_cross_entropy", "FusedRMSprop"]
