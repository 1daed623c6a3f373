"""Dispatch-order parity: the product scheduler (csrc/sched.cpp through the
C ABI) must make bit-identical decisions to the CPU oracle restating dispatcher.rs:195-262,314-341.

Integer work: the bar is exact equality of the ordered (user, user_seq, backend) list.
The reference pins no vectors; tests/golden/dispatch_seed.json holds the traces SURVEY.md 3.2 derives by hand.
"""
import json
import os
import random

import pytest
from hypothesis import given, settings, strategies as st

import ollamamq_b200 as mq
from ollamamq_b200.dispatcher import simulate as product_simulate, PROCESSED, DROPPED, UNCOUNTED
from oracle.dispatch_oracle import OracleC, OraclePy, simulate as oracle_simulate

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = json.load(open(os.path.join(HERE, "golden", "dispatch_seed.json")))
USERS4 = SEED["users"]
ARR4 = [(0, u) for u in USERS4 for _ in range(SEED["requests_per_user"])]  # user-major, all before the first pass


def _kw(name):
    if name.startswith("vip="):
        return {"vip": name[4:]}
    if name.startswith("boost="):
        return {"boost": name[6:]}
    return {}


# ----------------------------------------------------------------------------------- golden seed vectors
@pytest.mark.parametrize("name", list(SEED["traces"]))
@pytest.mark.parametrize("cls", [OracleC, OraclePy])
def test_oracle_matches_seed_vectors(name, cls):
    got = oracle_simulate(cls(2), ARR4, lambda u, s, b: 1, **_kw(name))
    assert [list(d) for d in got] == SEED["traces"][name]


@pytest.mark.parametrize("name", list(SEED["traces"]))
def test_product_matches_seed_vectors(name):
    got = product_simulate(mq.Scheduler(2), ARR4, lambda u, s, b: 1, **_kw(name))
    assert [[d.user, d.user_seq, d.backend] for d in got] == SEED["traces"][name]


def test_config1_variable_service_times():
    """BASELINE config 1 variant: per-task service times from numpy default_rng(0).integers(1,5)."""
    import numpy as np
    rng = np.random.default_rng(0)
    svc = {(u, s): int(rng.integers(1, 5)) for u in USERS4 for s in range(8)}
    f = lambda u, s, b: svc[(u, s)]
    for kw in ({}, {"vip": "charlie"}, {"boost": "david"}):
        ref = oracle_simulate(OracleC(2), ARR4, f, **kw)
        ref2 = oracle_simulate(OraclePy(2), ARR4, f, **kw)
        got = product_simulate(mq.Scheduler(2), ARR4, f, **kw)
        assert ref == ref2
        assert [d.key() for d in got] == ref
        assert len(got) == 32


# ----------------------------------------------------------------------------------- property tests
NAMES = ["alice", "bob", "charlie", "david", "eve", "Zed", "zed", "anonymous", "user10", "user2", "émile",
         "a", "aa", "ab", ""]


@st.composite
def traces(draw):
    n_users = draw(st.integers(1, 10))
    users = draw(st.lists(st.sampled_from(NAMES), min_size=n_users, max_size=n_users, unique=True))
    n_backends = draw(st.integers(1, 8))
    capacity = draw(st.sampled_from([1, 1, 1, 2, 4]))
    n_req = draw(st.integers(0, 60))
    arrivals = [(draw(st.integers(0, 12)), draw(st.sampled_from(users))) for _ in range(n_req)]
    seed = draw(st.integers(0, 2 ** 31))
    vip = draw(st.one_of(st.none(), st.sampled_from(users)))
    boost = draw(st.one_of(st.none(), st.sampled_from(users)))
    drop_p = draw(st.sampled_from([0.0, 0.0, 0.3]))
    return users, n_backends, capacity, arrivals, seed, vip, boost, drop_p


@settings(max_examples=150, deadline=None)
@given(traces())
def test_product_equals_oracles_on_random_traces(tr):
    users, n_backends, capacity, arrivals, seed, vip, boost, drop_p = tr
    rs = random.Random(seed)
    svc_tbl, out_tbl = {}, {}

    def svc(u, s, b):
        return svc_tbl.setdefault((u, s), rs.randint(1, 6))

    def outc(u, s):
        if (u, s) not in out_tbl:
            r = rs.random()
            out_tbl[(u, s)] = PROCESSED if r >= drop_p else (DROPPED if r < drop_p * 0.7 else UNCOUNTED)
        return out_tbl[(u, s)]

    # freeze the tables with one oracle pass so all three runs see the same numbers
    ref = oracle_simulate(OracleC(n_backends, capacity), arrivals, svc, vip=vip, boost=boost, outcomes=outc)
    ref2 = oracle_simulate(OraclePy(n_backends, capacity), arrivals, svc, vip=vip, boost=boost, outcomes=outc)
    got = product_simulate(mq.Scheduler(n_backends, capacity), arrivals, svc, vip=vip, boost=boost, outcomes=outc)
    assert ref == ref2
    assert [d.key() for d in got] == ref
    # every task is dispatched exactly once, FIFO inside a user
    assert len(ref) == len(arrivals)
    per_user = {}
    for u, s, b in ref:
        assert s == per_user.get(u, 0)
        per_user[u] = s + 1
        assert 0 <= b < n_backends


@settings(max_examples=60, deadline=None)
@given(st.integers(2, 6), st.integers(1, 4), st.integers(0, 2 ** 31))
def test_vip_has_absolute_priority(n_users, n_backends, seed):
    """While the VIP has queued work no other user is dispatched (dispatcher.rs:230)."""
    rs = random.Random(seed)
    users = ["u%02d" % i for i in range(n_users)]
    vip = users[rs.randrange(n_users)]
    s = mq.Scheduler(n_backends)
    s.set_vip(vip)
    for u in users:
        for _ in range(rs.randint(1, 5)):
            s.enqueue(u)
    inflight = []
    while True:
        ds = s.drain()
        for d in ds:
            if s.user_stats(vip)["queued"] > 0:
                assert d.user == vip
            inflight.append(d)
        if not inflight:
            break
        d = inflight.pop(rs.randrange(len(inflight)))
        s.complete(d.backend, d.user, PROCESSED)


def test_boost_every_second_dispatch_not_fifth():
    """Trap 1 of SURVEY.md: Boost fires when global_counter % 2 == 0 (dispatcher.rs:233), whatever the README says."""
    s = mq.Scheduler(1)
    s.set_boost("b")
    for u, n in (("a", 6), ("b", 30), ("c", 6)):
        for _ in range(n):
            s.enqueue(u)
    order = []
    for _ in range(16):
        d = s.next()
        order.append(d.user)
        s.complete(d.backend, d.user, PROCESSED)
    assert order[0::2] == ["b"] * 8               # counter 0,2,4,... -> boost (while it has queued work)
    assert set(order[1::2]) - {"b"}               # odd turns go through the ordinary positional RR


# ----------------------------------------------------------------------------------- edge cases
def test_empty_and_parking():
    s = mq.Scheduler(2)
    assert s.next() is None                        # no users: loop parks (:221-222)
    s.enqueue("x")
    s.set_online(0, False)
    s.set_online(1, False)
    assert s.next() is None                        # no eligible backend: parks without touching state (:208-209)
    assert s.counter == 0 and s.user_stats("x")["queued"] == 1
    s.set_online(1, True)
    d = s.next()
    assert d.key() == ("x", 0, 1)
    assert s.next() is None


def test_first_dispatch_goes_to_backend_1():
    """last_backend_idx starts at 0 and the tie-break is strictly-greater (dispatcher.rs:93,250)."""
    s = mq.Scheduler(3)
    for _ in range(4):
        s.enqueue("u")
    assert [d.backend for d in s.drain()] == [1, 2, 0]


def test_anonymous_default_and_bytewise_order():
    s = mq.Scheduler(1, capacity=8)
    s.enqueue(None)
    for u in ["b", "B", "é", "a"]:
        s.enqueue(u)
    got = [d.user for d in s.drain()]
    # sorted bytewise: B < a < anonymous < b < é; the positional index then skips as users leave the list
    # (trap 2 of SURVEY.md): idx0 of [B,a,anonymous,b,é] -> B, idx1 of [a,anonymous,b,é] -> anonymous, ...
    assert got == ["B", "anonymous", "é", "a", "b"]
    o = OraclePy(1, capacity=8)
    o.enqueue(None)
    for u in ["b", "B", "é", "a"]:
        o.enqueue(u)
    assert got == [o.next()[0] for _ in range(5)]


def test_least_connections_with_capacity():
    s = mq.Scheduler(2, capacity=2)
    for _ in range(4):
        s.enqueue("u")
    ds = s.drain()
    assert [d.backend for d in ds] == [1, 0, 1, 0]   # equal load -> RR; then least-connections keeps them even
    assert s.next() is None
    s.complete(1, "u", PROCESSED)
    s.enqueue("u")
    assert s.next().backend == 1                      # backend 1 now has fewer connections


def test_completion_counters_and_saturation():
    s = mq.Scheduler(1)
    s.enqueue("u")
    s.enqueue("u")
    s.enqueue("u")
    for oc in (PROCESSED, DROPPED, UNCOUNTED):
        d = s.next()
        s.processing("u", +1)
        s.complete(d.backend, "u", oc)
        s.processing("u", -1)
    st_ = s.user_stats("u")
    assert (st_["processed"], st_["dropped"], st_["processing"], st_["queued"]) == (1, 1, 0, 0)
    b = s.backend_stats(0)
    assert b["processed_count"] == 3 and b["active_requests"] == 0   # backend counter moves on every exit (:339)
    s.complete(0, "u", UNCOUNTED)                                     # saturating_sub (:338)
    assert s.backend_stats(0)["active_requests"] == 0
    s.processing("u", -5)
    assert s.user_stats("u")["processing"] == 0


def test_vip_boost_mutual_exclusion_and_tui_order():
    s = mq.Scheduler(1)
    s.set_vip("a")
    s.set_boost("a")      # tui.rs:169-175: boost on the VIP holder clears VIP
    s.enqueue("a")
    s.enqueue("b")
    s.enqueue("b")
    s.enqueue("c")
    # TUI order: (queued+processing desc, processed+dropped desc, name asc) (tui.rs:70-80)
    assert s.users_tui_order() == ["b", "a", "c"]
    d = s.next()          # counter 0 -> boost a
    assert d.user == "a"


def test_offline_backend_skipped_for_new_work_only():
    s = mq.Scheduler(2)
    for _ in range(3):
        s.enqueue("u")
    ds = s.drain()
    assert [d.backend for d in ds] == [1, 0]
    s.set_online(1, False)
    s.complete(1, "u", PROCESSED)     # in-flight request on the offline backend still completes
    assert s.next() is None           # backend 1 is free but offline; backend 0 busy
    s.complete(0, "u", PROCESSED)
    assert s.next().backend == 0


def test_scheduler_rejects_bad_arguments():
    with pytest.raises(mq.MQError):
        mq.Scheduler(0)
    s = mq.Scheduler(1)
    with pytest.raises(mq.MQError):
        s.enqueue("x" * 300)
    with pytest.raises(mq.MQError):
        s.complete(5, "u", PROCESSED)


# ----------------------------------------------------------------------------------- config 3 (extension)
def _zipf_counts(n_users=32, total=256, s=1.1, seed=0):
    """Request counts proportional to Zipf(s) over the users, totalling `total`."""
    import numpy as np
    w = 1.0 / np.arange(1, n_users + 1) ** s
    raw = w / w.sum() * total
    counts = np.maximum(1, np.floor(raw).astype(int))
    rng = np.random.default_rng(seed)
    while counts.sum() < total:
        counts[rng.integers(0, n_users)] += 1
    while counts.sum() > total:
        i = int(np.argmax(counts))
        counts[i] -= 1
    return counts.tolist()


def test_config3_multi_vip_boost_extension_priority_fairness():
    """BASELINE config 3: 32 users, Zipf-skewed arrivals, 2 VIP + 4 Boost.  The reference has ONE VIP and ONE Boost
    slot (dispatcher.rs:57-58); the set semantics are an extension (first member in the :224-228 sort order wins)
    that must (a) be identical in product and both oracles, (b) reduce to the reference with one member, and
    (c) order mean queue wait VIP < Boost < everyone else under load."""
    import numpy as np
    users = ["user%02d" % i for i in range(32)]
    counts = _zipf_counts()
    rng = np.random.default_rng(0)
    arrivals = []
    for u, c in zip(users, counts):
        arrivals += [(int(t), u) for t in rng.integers(0, 40, c)]
    vips, boosts = ["user09", "user20"], ["user03", "user12", "user25", "user30"]
    svc = lambda u, s, b: 2 + (hash((u, s)) % 3 == 0)        # deterministic within the process
    ref = oracle_simulate(OracleC(4), arrivals, svc, vip=vips, boost=boosts)
    ref2 = oracle_simulate(OraclePy(4), arrivals, svc, vip=vips, boost=boosts)
    got = product_simulate(mq.Scheduler(4), arrivals, svc, vip=vips, boost=boosts)
    assert ref == ref2 and [d.key() for d in got] == ref and len(ref) == 256
    # (b) singleton sets == the reference's single slots
    one = oracle_simulate(OracleC(4), arrivals, svc, vip=["user09"], boost=["user03"])
    assert one == oracle_simulate(OracleC(4), arrivals, svc, vip="user09", boost="user03")
    # (c) priority fairness: mean queue wait on the simulated clock (dispatch time - arrival time), per class
    t_disp = {}
    oracle_simulate(OracleC(4), arrivals, svc, vip=vips, boost=boosts,
                    on_dispatch=lambda t, u, s, b: t_disp.__setitem__((u, s), t))
    t_arr = {}
    for t, u in sorted(arrivals, key=lambda x: x[0]):       # FIFO inside a user: seq-th arrival = seq-th dispatch
        t_arr.setdefault(u, []).append(t)

    def mean_wait(group):
        return float(np.mean([t_disp[(u, s)] - t_arr[u][s] for u in group for s in range(len(t_arr[u]))]))

    others = [u for u in users if u not in vips + boosts]
    assert mean_wait(vips) < mean_wait(boosts) < mean_wait(others), (mean_wait(vips), mean_wait(boosts), mean_wait(others))


def test_a_user_holds_at_most_one_flag_in_the_extension():
    s = mq.Scheduler(1)
    s.add_vip("a")
    s.add_vip("b")
    s.add_boost("a")          # moves a from VIP to Boost (tui.rs:169-175 generalised)
    for u in ("a", "b", "c"):
        s.enqueue(u)
    assert [s.next().user for _ in range(1)] == ["b"]     # the only VIP left
    s.complete(0, "b", PROCESSED)
    assert s.next().user in ("a", "c")                     # counter is odd: ordinary round-robin turn


def test_decision_bench_matches_the_oracle_driver(capsys):
    """SURVEY.md 8(d) micro-benchmark: the product's bulk driver (mq_debug_sched_bench) and the oracle's (orc_bench) run
    the same trace - U users x R requests, unit service time, lowest backend completes first - and must make the same
    number of dispatches; the rates are printed for profiles/r01_dispatch_bench.txt (pytest -s)."""
    import ctypes as C
    import time
    import ollamamq_b200 as mq
    L = OracleC(1).L
    rows = []
    for users, backends, capacity in ((4, 2, 1), (64, 1, 1), (64, 8, 1), (256, 1, 1), (256, 8, 32)):
        n, sec = C.c_uint64(), C.c_double()
        mq.check(mq.lib.mq_debug_sched_bench(users, 32, backends, capacity, C.byref(n), C.byref(sec)))
        t0 = time.perf_counter()
        no = L.orc_bench(users, 32, backends, capacity)
        dt = time.perf_counter() - t0
        assert n.value == no == users * 32
        rows.append((users, backends, capacity, n.value / max(sec.value, 1e-9), no / max(dt, 1e-9)))
    with capsys.disabled():
        for r in rows:
            print("dispatch bench: %4d users %d backends capacity %2d | C++ scheduler %.2e decisions/s | oracle "
                  "(re-sorts per decision, incl. enqueue) %.2e decisions/s" % r)
