"""Host dispatcher (dsocr/dispatch.py) without a GPU: group planning and the work-stealing pool with stub engines."""
import threading
import time

import pytest

from dsocr.dispatch import EnginePool, plan_groups


def test_plan_groups():
    assert plan_groups(0, 4, 512) == []
    g = plan_groups(1024, 8, 512)
    assert len(g) == 8 and all(len(r) == 128 for r in g) and g[0] == range(0, 128) and g[-1] == range(896, 1024)
    g = plan_groups(1024, 1, 512)
    assert [len(r) for r in g] == [512, 512]
    g = plan_groups(1030, 2, 256)              # 3 rounds of 2 groups, no straggler group
    assert len(g) == 6 and sum(len(r) for r in g) == 1030 and max(len(r) for r in g) - min(len(r) for r in g) <= 1
    assert [len(r) for r in plan_groups(3, 8, 512)] == [1, 1, 1]
    flat = [i for r in plan_groups(777, 3, 100) for i in r]
    assert flat == list(range(777))


class Stub:
    def __init__(self, delay):
        self.delay, self.calls, self.lock = delay, [], threading.Lock()

    def decode_pages(self, pages, vs, seg0, seg1, image_id, params):
        time.sleep(self.delay * len(pages))
        with self.lock:
            self.calls.append(len(pages))
        return [("out", p) for p in pages]

    def close(self):
        pass


def test_pool_orders_results_and_steals_work():
    fast, slow = Stub(0.0002), Stub(0.002)
    pool = EnginePool([fast, slow], max_group=8)
    pages = list(range(100))
    out = pool.decode_pages(pages, None, [], [], 0, None)
    assert out == [("out", p) for p in pages]                     # page order whatever engine ran them
    assert sum(fast.calls) + sum(slow.calls) == 100
    assert sum(fast.calls) > sum(slow.calls)                       # the faster engine pulled more groups
    assert pool.last_assignment == [fast.calls, slow.calls]


def test_pool_propagates_errors():
    class Bad(Stub):
        def decode_pages(self, *a, **k):
            raise RuntimeError("image embedding failed: boom")

    pool = EnginePool([Stub(0.0), Bad(0.0)], max_group=4)
    with pytest.raises(RuntimeError, match="image embedding failed"):
        pool.decode_pages(list(range(32)), None, [], [], 0, None)
