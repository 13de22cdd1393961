"""egopack_b200 -- B200-native (sm_100a) implementation of EgoPack's temporal-graph hot path.

Drop-in module mirrors (same constructor / forward signatures and state_dict keys as the reference):

    egopack_b200.models.graph.Graph                       <- models/graph.py
    egopack_b200.models.temporal_pooling.trn_pooling.TRNPooling
    egopack_b200.models.tasks.{RecognitionTask, LTATask, OSCCTask, PNRTask}
    egopack_b200.models.graphONE.graphONE.GraphONE
    egopack_b200.models.transforms.{RadiusGraph, LTATemporalConnectivity}

All arithmetic runs in hand-written CUDA kernels behind the C ABI of ``include/egopack_b200.h``.
"""
from . import config
from .config import get_precision, precision, set_precision
from .data import Batch, Data

__all__ = ["config", "precision", "set_precision", "get_precision", "Data", "Batch"]
__version__ = "0.1.0"
