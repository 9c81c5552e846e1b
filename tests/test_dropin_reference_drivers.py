"""Drop-in check against the LIVE reference checkout (build container only; skipped where /root/reference is absent):
after `xmem2_b200.install()` the reference's own driver module imports and builds its main objects on top of THIS
package's InferenceCore / MemoryManager / KeyValueMemoryStore / XMem (inference/run_on_video.py:148-162) — same module
paths, constructor signatures and attribute surface.  No frame is processed here (no CUDA device)."""
import os
import sys
import types

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present on this box')


@pytest.fixture
def reference_on_path(monkeypatch):
    stub = types.ModuleType('progressbar')          # inference/data/video_reader.py:7 imports progressbar2 (absent here)
    stub.ProgressBar = type('ProgressBar', (), {'__init__': lambda s, *a, **k: None, 'update': lambda s, *a, **k: None,
                                                'finish': lambda s, *a, **k: None})
    stub.progressbar = lambda x, *a, **k: x
    monkeypatch.setitem(sys.modules, 'progressbar', stub)
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('inference', 'model', 'util', 'dataset')}
    for k in saved:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend(REF)
    import xmem2_b200
    xmem2_b200.install()
    yield
    for k in [k for k in sys.modules if k.split('.')[0] in ('inference', 'model', 'util', 'dataset')]:
        sys.modules.pop(k, None)
    sys.modules.update(saved)


def test_reference_driver_binds_to_this_package(reference_on_path):
    import inference.run_on_video as rov                     # the reference's driver, unmodified
    import xmem2_b200.inference.inference_core as mine_core
    import xmem2_b200.model.network as mine_net
    assert rov.InferenceCore is mine_core.InferenceCore and rov.XMem is mine_net.XMem
    assert rov.VIDEO_INFERENCE_CONFIG['top_k'] == 30 and rov.VIDEO_INFERENCE_CONFIG['mem_every'] == 10
    # what _load_main_objects does (run_on_video.py:148-162), minus the data loaders
    config = rov.VIDEO_INFERENCE_CONFIG.copy()
    config['model'] = None
    network = rov.XMem(config, None, pretrained_key_encoder=False, pretrained_value_encoder=False).eval()
    assert config['key_dim'] == 64 and config['value_dim'] == 512 and config['hidden_dim'] == 64     # network.py:178-180
    processor = rov.InferenceCore(network, config=config)
    processor.set_all_labels([1])
    # attribute surface the drivers / GUI read
    for attr in ('memory', 'network', 'config', 'mem_every', 'deep_update_every', 'enable_long_term', 'curr_ti', 'last_mem_ti',
                 'all_labels'):
        assert hasattr(processor, attr), attr
    m = processor.memory
    for attr in ('temporary_work_mem', 'permanent_work_mem', 'long_mem', 'frame_id_to_permanent_mem_idx', 'hidden', 'top_k',
                 'max_mt_frames', 'min_mt_frames', 'num_prototypes', 'max_long_elements'):
        assert hasattr(m, attr), attr
    assert m.temporary_work_mem.size == 0 and m.permanent_work_mem.size == 0 and processor.permanent_memory_frames == []
    # upstream checkpoints load by key (network.py:184-198): same names and shapes as the reference's own module
    from xmem2_b200.util.synth import xmem_param_spec
    sd = network.state_dict()
    assert set(sd) == set(xmem_param_spec())
    # no CUDA device here: the hot path must refuse loudly instead of falling back
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        processor.step(torch.zeros(3, 64, 96))
