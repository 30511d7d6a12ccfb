"""``build_model(config)`` of the reference (``Fall_2_Spatial_Temporal_SR/Model/build_model.py:5-18``): model name -> class.

``config`` is the reference's yacs node (``config.MODEL.NAME``, ``config.DATA.IN_CHANNELS`` / ``NUM_CLASSES`` / ``SENSOR_DIM``,
``config.GRAPH.LAYOUT`` / ``STRATEGY``); any object with those attributes works, and so does a nested dict.
"""
from __future__ import annotations

from .fusion import TwoStreamSTGCAN, TwoStreamSTGCAN_BiLSTM
from .sensor import BiLSTM
from .stgcan import STGCAN


class _Node:
    def __init__(self, d):
        self._d = d

    def __getattr__(self, k):
        v = self._d[k]
        return _Node(v) if isinstance(v, dict) else v


def build_model(config):
    if isinstance(config, dict):
        config = _Node(config)
    name = config.MODEL.NAME
    graph_args = {"layout": config.GRAPH.LAYOUT, "strategy": config.GRAPH.STRATEGY}
    if name == "stgcn":
        return STGCAN(config.DATA.IN_CHANNELS, graph_args, num_class=config.DATA.NUM_CLASSES)
    if name == "bilstm":
        return BiLSTM(input_size=config.DATA.SENSOR_DIM, hidden_size=64, num_layers=1, dropout_prob=0.3,
                      num_classes=config.DATA.NUM_CLASSES, feature="mean")
    if name == "two_stgcan":
        return TwoStreamSTGCAN(config.DATA.IN_CHANNELS, graph_args, num_class=config.DATA.NUM_CLASSES)
    if name == "two_stgcan_bilstm":
        return TwoStreamSTGCAN_BiLSTM(config.DATA.IN_CHANNELS, graph_args, num_class=config.DATA.NUM_CLASSES,
                                      bilstm_input_size=config.DATA.SENSOR_DIM)
    raise RuntimeError(f"Model name [{name}] is not implemented.")
