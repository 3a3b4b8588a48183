"""dedark_yolo_b200 -- B200-native (sm_100a) implementation of Dedark-YOLO's low-light hot path.

Public surface (mirrors the reference's names):
    lowlight_recovery          drop-in nn.Module         (ultralytics/nn/modules/llie.py)
    preprocess_batch           synthesis + recovery loss (models/yolo/detect/train.py:70-111)
    apply_lowlight             offline darkener, pre-encode array (utils/lowlight_process.py:57-74)
    darken_directory           the whole offline tool: nvJPEG decode -> darken -> nvJPEG encode (utils/lowlight_process.py:10-87)
    add_recovery_term          loss term                 (utils/loss.py:393-416)
    HostBatchPrefetcher        double-buffered H2D staging of the dataloader's pinned uint8 batches
    RecoveryPipeline           device-resident synth -> fwd -> bwd step used by bench.py

Importing this package loads ``lib/libdedark_b200.so`` and fails loudly if it has not been built.
"""
from ._lib import launch_count, lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from .llie import ConvBlock, ExtractParameters2, lowlight_recovery  # noqa: F401
from .lowlight import HostBatchPrefetcher, add_recovery_term, apply_lowlight, preprocess_batch  # noqa: F401
from .offline import darken_directory  # noqa: F401
from .pipeline import RecoveryPipeline  # noqa: F401

__version__ = "0.1.0"
