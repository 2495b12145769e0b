"""B200-native Deep Speaker hot path: drop-in for /root/reference/model.py's DeepSpeakerModel,
TripletMarginLoss and PairwiseDistance, backed by hand-written sm_100a CUDA behind a C ABI
(include/dsk.h, lib/libdsk.so)."""
from .model import (DeepSpeakerModel, PairwiseDistance, TripletMarginLoss, allpairs_topk,  # noqa: F401
                    select_hard_triplets)

from .pipeline import EmbeddingPipeline  # noqa: F401,E402
from .head import CrossEntropyLoss  # noqa: F401,E402
from .optim import FusedAdagrad  # noqa: F401,E402
from .steps import train_step  # noqa: F401,E402

__all__ = ["train_step", "CrossEntropyLoss", "FusedAdagrad", "EmbeddingPipeline", "DeepSpeakerModel", "PairwiseDistance", "TripletMarginLoss", "select_hard_triplets", "allpairs_topk"]
