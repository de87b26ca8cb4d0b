"""Host-side bookkeeping of the pipelined delivery (rtgs.ray_tracer.PendingFrame) with a stub library: frames are
collected in submission order, whichever result() is asked for first."""
import collections

import pytest

from rtgs import _native, ray_tracer


class _StubLib:
    def __init__(self):
        self.collected = 0
        self.fail_at = None

    def rtgs_render_host_collect(self, handle):
        self.collected += 1
        return -2 if self.fail_at == self.collected else 0

    def rtgs_last_error(self):
        return b"stub failure"


class _Scene:
    def __init__(self):
        self._pending = collections.deque()
        self.handle = 1234


@pytest.fixture
def stub(monkeypatch):
    lib = _StubLib()
    monkeypatch.setattr(_native, "load", lambda: lib)
    return lib


def _submit(scene, tag):
    f = ray_tracer.PendingFrame(scene, tag)
    scene._pending.append(f)
    return f


def test_results_are_collected_in_submission_order(stub):
    scene = _Scene()
    a, b = _submit(scene, "A"), _submit(scene, "B")
    assert b.result() == "B" and stub.collected == 2 and a.done and b.done   # B's result() collected A first
    assert a.result() == "A" and stub.collected == 2                          # already done: no second collect
    assert not scene._pending
    c = _submit(scene, "C")
    assert c.result() == "C" and stub.collected == 3


def test_a_failed_collect_raises_and_leaves_the_queue_consistent(stub):
    scene = _Scene()
    a, b = _submit(scene, "A"), _submit(scene, "B")
    stub.fail_at = 1
    with pytest.raises(RuntimeError):
        b.result()
    assert a.done and not b.done and list(scene._pending) == [b]
    assert b.result() == "B" and stub.collected == 2
