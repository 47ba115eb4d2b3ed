"""CPU: the request queue in front of tts_batch (cosyvoice_b200/batcher.py) with a fake model: ragged admission
(batch closed by size or by the wait budget, arrival order kept), per-request results and failures, and the int16 PCM wire format of
the reference's servers (runtime/python/fastapi/server.py:42, grpc/server.py:64)."""
import threading
import time

import numpy as np
import pytest
import torch

from cosyvoice_b200.batcher import TtsBatcher, pcm16, pcm16_decode


class FakeModel:
    """tts_batch returns, for request i, a waveform whose samples are all `tag` and whose length is 10 * len(text)"""

    def __init__(self, delay=0.0, fail_on=None):
        self.calls, self.delay, self.fail_on = [], delay, fail_on
        self.gate = threading.Event()
        self.gate.set()

    def tts_batch(self, inputs):
        self.gate.wait()
        self.calls.append([i["tag"] for i in inputs])
        time.sleep(self.delay)
        if self.fail_on is not None and any(i["tag"] == self.fail_on for i in inputs):
            raise RuntimeError("boom")
        return [torch.full((1, 10 * i["text"].shape[1]), float(i["tag"]) / 100.0) for i in inputs]


def req(tag, n=3):
    return dict(text=torch.zeros(1, n, dtype=torch.int32), tag=tag)


def test_pcm16_is_the_reference_expression():
    g = torch.Generator().manual_seed(0)
    w = torch.rand(1, 1000, generator=g) * 1.98 - 0.99
    assert pcm16(w) == (w.numpy() * (2 ** 15)).astype(np.int16).tobytes()
    assert len(pcm16(w)) == 2000
    assert np.frombuffer(pcm16(torch.tensor([[0.5, -0.5, 0.0]])), dtype="<i2").tolist() == [16384, -16384, 0]
    back = pcm16_decode(pcm16(w))                    # grpc/server.py:45-46 reads prompts this way
    assert back.shape == w.shape and (back - w).abs().max().item() <= 2.0 ** -15


def test_batches_close_on_size_and_keep_arrival_order():
    m = FakeModel()
    m.gate.clear()                                   # hold the worker inside its first call until everything is queued
    with TtsBatcher(m, max_batch=4, max_wait_ms=10_000) as b:
        first = b.submit(**req(0))
        time.sleep(0.05)                             # the worker has taken request 0... or is still waiting for a full batch
        futs = [first] + [b.submit(**req(t, n=t + 1)) for t in range(1, 10)]
        m.gate.set()
        res = [f.result(timeout=10) for f in futs]
    assert [int(round(r[0, 0].item() * 100)) for r in res] == list(range(10))
    assert [r.shape[1] for r in res] == [30] + [10 * (t + 1) for t in range(1, 10)]
    flat = [t for c in m.calls for t in c]
    assert flat == list(range(10))                   # never reordered
    assert all(len(c) <= 4 for c in m.calls) and sum(b.batches) == 10
    assert max(len(c) for c in m.calls) == 4         # full batches formed while requests were waiting


def test_wait_budget_closes_a_partial_batch():
    m = FakeModel()
    with TtsBatcher(m, max_batch=32, max_wait_ms=30) as b:
        t0 = time.monotonic()
        f = b.submit(**req(7))
        f.result(timeout=5)
        dt = time.monotonic() - t0
    assert m.calls == [[7]]
    assert 0.02 <= dt < 2.0                          # waited for company, but not for long


def test_failure_reaches_the_requests_of_that_batch_only():
    m = FakeModel(fail_on=2)
    with TtsBatcher(m, max_batch=2, max_wait_ms=10_000) as b:
        m.gate.clear()
        futs = [b.submit(**req(t)) for t in range(4)]          # batches [0,1] and [2,3]
        m.gate.set()
        assert futs[0].result(timeout=5).shape == (1, 30) and futs[1].result(timeout=5).shape == (1, 30)
        for f in futs[2:]:
            with pytest.raises(RuntimeError):
                f.result(timeout=5)


def test_pcm_future_and_close():
    m = FakeModel()
    b = TtsBatcher(m, max_batch=8, max_wait_ms=5)
    f = b.submit_pcm(**req(50, n=2))
    assert f.result(timeout=5) == (np.full((1, 20), 0.5, dtype=np.float32) * 2 ** 15).astype(np.int16).tobytes()
    b.close()
    with pytest.raises(RuntimeError):
        b.submit(**req(1))
