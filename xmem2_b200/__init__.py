"""xmem2_b200 — B200-native implementation of the XMem++ per-frame memory-attention inference path.

Module paths mirror the reference repository (`inference.inference_core`, `inference.memory_manager`,
`inference.kv_memory_store`, `inference.frame_selection.frame_selection`, `model.network`,
`model.aggregate`; `util.tensor_util` and
`util.configuration` exist for internal use).  `install()` registers them under those top-level names so the reference's own drivers
(`process_video.py`, `inference/run_on_video.py`) import this implementation unchanged — see INTEGRATION.md.
"""
import importlib
import sys

__version__ = '0.1.0'

# util.tensor_util / util.configuration are NOT aliased: the reference's own modules carry extra driver-side helpers
# (compute_array_iou, the argparse Configuration) and their pad/unpad/VIDEO_INFERENCE_CONFIG are semantically identical
# to the copies this package uses internally.
_DROP_IN = ['inference.inference_core', 'inference.memory_manager', 'inference.kv_memory_store', 'model.network',
            'model.aggregate', 'inference.frame_selection.frame_selection']


def install():
    """Alias this package's modules over the reference's module paths (call before importing the drivers).
    Modules this package does not provide (data readers, image saver, ...) keep resolving to the reference."""
    for name in _DROP_IN:
        sys.modules[name] = importlib.import_module(f'{__name__}.{name}')
