"""ravvent_basecaller_b200 -- B200-native (sm_100a) implementation of the Ravvent
inference hot path behind the reference's own class surface:

    EventDetector(...).run(raw)                       (event_detection/event_detector.py)
    Basecaller(...).load_weights / _encode_input / greedy_search_prediction /
        beam_search_prediction / tokens_to_nuc_sequences            (basecaller.py)
    Merger(scores_id).merge(list of SeqLogitsPair)                 (merger.py)

Importing this package requires the in-tree CUDA library (libravvent_b200.so);
there is no CPU or framework fallback."""
from . import _lib                                     # noqa: F401  (fails loudly when the library is missing)
from ._lib import RavventError, device_count, launch_count
from .event_detector import Event, EventDetector
from .basecaller import Basecaller
from . import data_loader
from .merger import Merger, SeqLogitsPair
from .data_loader import nuc_tk
from .sharding import ShardedBasecaller, shard_range

__all__ = ["Basecaller", "EventDetector", "Event", "Merger", "SeqLogitsPair", "RavventError", "data_loader", "nuc_tk",
           "device_count", "launch_count", "ShardedBasecaller", "shard_range"]
