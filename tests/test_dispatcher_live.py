"""The live, threaded dispatcher (csrc/dispatcher.cpp: AppState + run_worker + executor bookkeeping) on CPU with
step-driven mock backends.  SURVEY.md 8 rows a4-a8 and BASELINE config 1 ("4 users x 8 requests vs 2 mock
backends, CPU-only: dispatch-order parity").  The completion order of the oracle's simulated clock is imposed on
the real threads, then the dispatch log must equal the oracle's trace exactly.
"""
import os
import re

import pytest

import ollamamq_b200 as mq
from oracle.dispatch_oracle import OracleC, simulate as oracle_simulate

USERS4 = ["alice", "bob", "charlie", "david"]
ARR4 = [(0, u) for u in USERS4 for _ in range(8)]


def _enqueue_all_before_first_pass(d, users):
    """All tasks must be queued before the scheduler's first pass (event model of SURVEY.md 3.2): take the
    backends offline (which, like in the reference, does not wake the scheduler), enqueue all but the last task,
    bring them back, enqueue the last one (its notify wakes the loop)."""
    for b in range(d.n_backends):
        d.set_online(b, False)
    streams = [d.submit(u, max_new_tokens=1) for u in users[:-1]]
    d.wait_parked()
    assert d.log() == []
    for b in range(d.n_backends):
        d.set_online(b, True)
    streams.append(d.submit(users[-1], max_new_tokens=1))
    d.wait_parked()
    return streams


@pytest.mark.parametrize("kw", [{}, {"vip": "charlie"}, {"boost": "david"}])
@pytest.mark.parametrize("svc_seed", [None, 0])
def test_config1_live_dispatcher_matches_oracle(kw, svc_seed):
    if svc_seed is None:
        svc = lambda u, s, b: 1
    else:
        import numpy as np
        rng = np.random.default_rng(svc_seed)
        tbl = {(u, s): int(rng.integers(1, 5)) for u in USERS4 for s in range(8)}
        svc = lambda u, s, b: tbl[(u, s)]
    completions = []
    ref = oracle_simulate(OracleC(2), ARR4, svc, on_complete=lambda t, b, u, s: completions.append(b), **kw)
    d = mq.Dispatcher(mock_backends=2, capacity=1)
    try:
        if "vip" in kw:
            d.set_vip(kw["vip"])
        if "boost" in kw:
            d.set_boost(kw["boost"])
        streams = _enqueue_all_before_first_pass(d, [u for _, u in ARR4])
        for b in completions:
            assert d.mock_complete(b)
            d.wait_parked()
        d.drain(5000)
        assert d.log() == ref
        assert all(s.rc == 0 and s.status == 200 and s.body for s in streams)
        for u in USERS4:
            st = d.user_stats(u)
            assert (st["processed"], st["dropped"], st["queued"], st["processing"]) == (8, 0, 0, 0)
        assert sum(d.backend_stats(b)["processed_count"] for b in range(2)) == 32
    finally:
        d.close()


def test_blocked_user_and_ip():
    d = mq.Dispatcher(mock_backends=1)
    try:
        d.block_user("mallory")
        with pytest.raises(mq.MQError) as e:                       # 403 "User blocked" (:375-378)
            d.submit("mallory", max_new_tokens=1)
        assert e.value.rc == -13
        d.block_ip("10.0.0.9")
        with pytest.raises(mq.MQError):                            # 403 "IP blocked" (:370-373)
            d.submit("alice", ip="10.0.0.9", max_new_tokens=1)
        # blocked after enqueue: popped, dropped at the executor pre-flight, no backend call (:271-280)
        d.set_online(0, False)
        s1 = d.submit("bob", ip="10.0.0.1", max_new_tokens=1)
        d.wait_parked()
        d.block_ip("10.0.0.1")                                     # bob's last IP
        d.set_online(0, True)
        s2 = d.submit("carol", max_new_tokens=1)
        d.wait_parked()
        s1.wait(5)
        assert s1.rc == -13 and s1.status is None
        assert d.user_stats("bob")["dropped"] == 1
        assert d.backend_stats(0)["processed_count"] == 1         # backend counter still moves (:339)
        assert d.mock_complete(0)
        s2.wait(5)
        assert s2.rc == 0
    finally:
        d.close()


def test_backend_error_is_a_value_not_a_panic():
    d = mq.Dispatcher(mock_backends=1)
    try:
        d.mock_fail_next(0, 1)
        s = d.submit("alice", max_new_tokens=1).wait(5)
        assert s.rc < 0 and s.status is None and s.err.startswith("Backend error:")   # HTTP 500 text (:423-425)
        d.drain(2000)
        assert d.user_stats("alice")["dropped"] == 1               # (:326-327)
        assert d.backend_stats(0)["active_requests"] == 0
        s = d.submit("alice", max_new_tokens=1)
        d.wait_parked()
        assert d.mock_complete(0, rc=-5)                           # error before the Status part
        s.wait(5)
        assert s.rc == -5 and s.status is None
        assert d.user_stats("alice")["dropped"] == 2
    finally:
        d.close()


def test_client_disconnect_paths():
    d = mq.Dispatcher(mock_backends=1)
    try:
        # (a) mid-stream: the chunk send fails -> relay stops, dropped++ (:305-308,:318-319)
        got = []
        s = mq.Stream(on_chunk=lambda b: (got.append(b), False)[1])
        d.submit("alice", sink=s, max_new_tokens=5)
        d.wait_parked()
        assert d.mock_complete(0)
        s.wait(5)
        assert len(got) == 1 and s.rc != 0
        d.drain(2000)
        assert d.user_stats("alice") == {"queued": 0, "processing": 0, "processed": 0, "dropped": 1}
        # (b) gone while still queued: popped, consumes a turn + counter, dropped, backend released (:278-280)
        d.set_online(0, False)
        s1 = d.submit("bob", max_new_tokens=1)
        d.wait_parked()
        d.client_gone(s1.task_id)
        d.set_online(0, True)
        s2 = d.submit("bob", max_new_tokens=1)
        d.wait_parked()
        assert d.log()[-2:] == [("bob", 0, 0), ("bob", 1, 0)]
        assert d.mock_complete(0)
        s2.wait(5)
        d.drain(2000)
        st = d.user_stats("bob")
        assert (st["processed"], st["dropped"]) == (1, 1)
    finally:
        d.close()


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports every function include/ollamamq_b200.h declares."""
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                            "ollamamq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(mq_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 50
    missing = [n for n in sorted(names) if not hasattr(mq.lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_for_the_forward_pass():
    """Without a B200 the worker refuses to open: there is no CPU path to fall back to."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oracle.llama_ref import TINY_LLAMA
    assert mq.lib.mq_worker_count() == 0
    with pytest.raises(mq.MQError) as e:
        mq.Worker(0, mq.model_cfg(TINY_LLAMA, max_batch=4, max_seq=64, max_prefill_tokens=64))
    assert e.value.rc == -19


def test_block_list_persistence_format(tmp_path):
    """blocked_items.json round trip; format pinned by BlockedConfig (dispatcher.rs:21-25,98-115)."""
    import json
    path = str(tmp_path / "blocked_items.json")
    d = mq.Dispatcher(mock_backends=1)
    try:
        d.set_block_file(path)                      # nothing to load yet
        d.block_user("mallory")
        d.block_ip("10.1.2.3")
        d.block_user('we"ird')
        cfg = json.load(open(path))
        assert sorted(cfg) == ["ips", "users"]
        assert cfg["ips"] == ["10.1.2.3"] and sorted(cfg["users"]) == sorted(["mallory", 'we"ird'])
        assert open(path).read().startswith('{\n  "ips": [')          # serde_json::to_string_pretty layout
        d.block_user("mallory", False)              # unblock rewrites the file too
        assert json.load(open(path))["users"] == ['we"ird']
    finally:
        d.close()
    d2 = mq.Dispatcher(mock_backends=1)            # restart: the list survives (checkpoint/resume)
    try:
        d2.set_block_file(path)
        with pytest.raises(mq.MQError) as e:
            d2.submit('we"ird', max_new_tokens=1)
        assert e.value.rc == -13
        with pytest.raises(mq.MQError):
            d2.submit("bob", ip="10.1.2.3", max_new_tokens=1)
        d2.submit("bob", ip="10.9.9.9", max_new_tokens=1)
        d2.wait_parked()
        assert d2.mock_complete(0)
        d2.drain(2000)
    finally:
        d2.close()


def test_health_prober_keeps_mock_backends_online():
    d = mq.Dispatcher(mock_backends=2)
    try:
        d.set_online(1, False)
        d.start_health(20)                          # reference period is 10 s; mocks always answer
        import time
        time.sleep(0.15)
        assert d.backend_stats(1)["is_online"]
    finally:
        d.close()


def test_health_prober_takes_a_backend_offline_and_recovery_does_not_wake_the_scheduler():
    """dispatcher.rs:171-193 with a backend that stops answering the probe: is_online flips, NEW work skips it (:201-209),
    what it has in flight still completes and is accounted (:314-341), and when it answers again the flag flips back but
    nobody notifies the scheduler (the prober never calls notify): queued work moves only at the next natural wake-up."""
    import time
    d = mq.Dispatcher(mock_backends=2, capacity=1)
    try:
        d.start_health(10)
        a = d.submit("alice", max_new_tokens=1)                # -> backend 1 (first dispatch, :248-254)
        b = d.submit("bob", max_new_tokens=1)                  # -> backend 0
        d.wait_parked()
        assert sorted(x[2] for x in d.log()) == [0, 1]
        d.mock_set_healthy(0, False)                           # backend 0's /api/tags stops answering
        for _ in range(200):
            if not d.backend_stats(0)["is_online"]:
                break
            time.sleep(0.005)
        assert not d.backend_stats(0)["is_online"] and d.backend_stats(1)["is_online"]
        assert d.mock_complete(0)                              # its in-flight request still completes ...
        d.wait_parked()
        assert d.user_stats("bob")["processed"] == 1           # ... and is counted
        assert d.backend_stats(0)["active_requests"] == 0
        c = d.submit("carol", max_new_tokens=1)                # backend 1 busy, backend 0 offline: stays queued
        d.wait_parked()
        assert d.user_stats("carol")["queued"] == 1 and len(d.log()) == 2
        assert d.mock_complete(1)                              # backend 1 frees -> carol goes THERE, not to 0
        d.wait_parked()
        assert d.log()[-1] == ("carol", 0, 1)
        e = d.submit("erin", max_new_tokens=1)                 # every eligible backend busy / offline
        d.wait_parked()
        assert d.user_stats("erin")["queued"] == 1
        d.mock_set_healthy(0, True)                            # backend 0 answers again
        for _ in range(200):
            if d.backend_stats(0)["is_online"]:
                break
            time.sleep(0.005)
        assert d.backend_stats(0)["is_online"]
        time.sleep(0.1)
        assert d.user_stats("erin")["queued"] == 1             # recovery did not wake run_worker (:181-191 has no notify)
        f = d.submit("frank", max_new_tokens=1)                # the next notify does; who goes first is the positional
        d.wait_parked()                                        # round-robin's business (:236-240), where is not: backend 0
        assert d.log()[-1][2] == 0 and d.log()[-1][0] in ("erin", "frank") and len(d.log()) == 4
        while d.mock_complete(0) or d.mock_complete(1):
            d.wait_parked()
        d.drain(5000)
        for s in (a, b, c, e, f):
            assert s.rc == 0
    finally:
        d.close()
