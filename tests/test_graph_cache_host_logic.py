"""CPU test of InferenceCore's recorded-graph bookkeeping with test doubles for the CUDA pieces: first frame of a
signature runs eagerly (warm-up), the second records, later frames and later cores on the same network hit the cache,
and a core that takes over a shared graph detaches the previous user's hidden state from the shared buffer."""
import torch

from xmem2_b200 import lib
from xmem2_b200.inference import inference_core as ic


class _FakeLib:
    def xm_add_launch_count(self, n):
        pass

    def xm_launch_count(self):
        return 0


class _FakeMem:
    hidden_dim = 64

    def __init__(self):
        self.h = torch.zeros(1, 1, 64, 2, 2)
        self.plans = 0
        self._ws = object()          # every manager starts with its own K1 workspace
        self.adopted = []

    def layout_signature(self):
        return ('layout',)

    def get_hidden(self):
        return self.h

    def set_hidden(self, h):
        self.h = h

    def upload_plan(self, hw, dev):
        self.plans += 1

    def adopt_workspace(self, ws):
        self._ws = ws
        self.adopted.append(ws)


class _FakeGraph:
    replays = 0

    def replay(self):
        _FakeGraph.replays += 1


def test_warmup_record_replay_and_cross_core_reuse(monkeypatch):
    monkeypatch.setattr(lib, 'load', lambda: _FakeLib())
    captures = []

    def fake_capture(self, image, mem_frame):
        captures.append(mem_frame)
        return {'image': image.clone(), 'owner': [None], 'hidden': torch.zeros(1, 1, 64, 2, 2), 'graph': _FakeGraph(),
                'launches': 3, 'prob': torch.zeros(2, 4, 4), 'ws': self.memory._ws}

    monkeypatch.setattr(ic.InferenceCore, '_capture', fake_capture)
    net = torch.nn.Linear(1, 1)

    def make_core():
        c = ic.InferenceCore.__new__(ic.InferenceCore)
        c.network, c.memory, c.all_labels = net, _FakeMem(), [1]
        c._graphs, c._graph_warm, c._g_out = {}, set(), None
        return c

    img = torch.zeros(1, 3, 32, 32)
    a = make_core()
    assert a._graph_step(img, False) is None                       # warm-up frame runs eagerly
    assert a._graph_step(img, False) is not None and captures == [False]
    assert a._graph_step(img, False) is not None and captures == [False]          # replay, no new recording
    assert a._graph_step(img, True) is None and a._graph_step(img, True) is not None and captures == [False, True]
    assert a.memory.plans == 3                                       # the device plan is refreshed before every replay
    b = make_core()                                                  # next video on the same network
    assert b._graph_step(img, False) is not None and captures == [False, True]
    assert a.memory.get_hidden().data_ptr() != b.memory.get_hidden().data_ptr()   # previous user got its own copy
    # the recorded kernels point into the FIRST core's K1 workspace: the inheriting core must adopt it
    assert b.memory.adopted == [a.memory._ws] and b.memory._ws is a.memory._ws
    assert len(ic._graph_cache(net)) == 2
    other = torch.nn.Linear(1, 1)                                    # a different network never sees these graphs
    assert ic._graph_cache(other) == {}
