"""dispatch_oracle.py — TEST INFRASTRUCTURE ONLY.

Independent Python restatement of ollamaMQ's scheduling decision (second opinion for dispatch_oracle.c):
    /root/reference/src/dispatcher.rs:195-262  run_worker loop body
    /root/reference/src/dispatcher.rs:314-341  executor epilogue
    /root/reference/src/dispatcher.rs:364-405  enqueue

PARITY UNPINNED by the reference (no golden vectors, no tests, cannot be compiled here): pinned only against
the hand-derived traces of SURVEY.md 3.2 (tests/golden/dispatch_seed.json).

Also provides `OracleC`, a ctypes wrapper around the plain-C oracle, and `simulate`, the event model of
SURVEY.md 3.2 (independent of ollamamq_b200.dispatcher.simulate).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from collections import deque
from typing import Callable, Dict, List, Optional, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))


class OraclePy:
    def __init__(self, n_backends: int, capacity: int = 1, boost_mod: int = 2):
        self.queues: Dict[str, deque] = {}            # HashMap<String, VecDeque<Task>>
        self.processed: Dict[str, int] = {}
        self.dropped: Dict[str, int] = {}
        self.popped: Dict[str, int] = {}
        self.backends = [{"active": 0, "processed": 0, "online": True} for _ in range(n_backends)]
        self.vip: List[str] = []      # EXTENSION (config 3): sets; the reference has one Option<String> each
        self.boost: List[str] = []
        self.counter = 0
        self.current_idx = 0
        self.last_idx = 0
        self.capacity = capacity
        self.boost_mod = boost_mod

    def enqueue(self, user: Optional[str]):
        u = "anonymous" if user is None else user
        self.queues.setdefault(u, deque()).append(object())

    def add_vip(self, u):
        if u in self.boost:
            self.boost.remove(u)
        if u not in self.vip:
            self.vip.append(u)

    def add_boost(self, u):
        if u in self.vip:
            self.vip.remove(u)
        if u not in self.boost:
            self.boost.append(u)

    def set_vip(self, u):
        self.vip = []
        if u is not None:
            self.add_vip(u)

    def set_boost(self, u):
        self.boost = []
        if u is not None:
            self.add_boost(u)

    def set_online(self, b, online):
        self.backends[b]["online"] = bool(online)

    def next(self) -> Optional[Tuple[str, int, int]]:
        online = [i for i, b in enumerate(self.backends) if b["online"] and b["active"] < self.capacity]
        if not online:
            return None
        active = [u for u, q in self.queues.items() if len(q) > 0]
        if not active:
            return None
        # byte-wise string order, like Rust's String Ord
        active.sort(key=lambda u: (self.processed.get(u, 0), u.encode("utf-8")))
        target = None
        for u in active:                      # sorted: the first VIP met wins (one VIP: `active.contains(vip)`, :230)
            if u in self.vip:
                target = u
                break
        if target is None and self.boost and self.counter % self.boost_mod == 0:
            for u in active:
                if u in self.boost:
                    target = u
                    break
        if target is None:
            if self.current_idx >= len(active):
                self.current_idx = 0
            target = active[self.current_idx]
            self.current_idx += 1
        self.queues[target].popleft()
        seq = self.popped.get(target, 0)
        self.popped[target] = seq + 1
        self.counter += 1
        min_conns = min(self.backends[i]["active"] for i in online)
        cands = [i for i in online if self.backends[i]["active"] == min_conns]
        pos = next((k for k, i in enumerate(cands) if i > self.last_idx), 0)
        sel = cands[pos]
        self.last_idx = sel
        self.backends[sel]["active"] += 1
        return (target, seq, sel)

    def complete(self, backend: int, user: str, outcome: int = 0):
        if outcome == 0:
            self.processed[user] = self.processed.get(user, 0) + 1
        elif outcome == 1:
            self.dropped[user] = self.dropped.get(user, 0) + 1
        b = self.backends[backend]
        b["active"] = max(0, b["active"] - 1)
        b["processed"] += 1


def build_c_oracle() -> str:
    so = os.path.join(HERE, "_build", "libdispatch_oracle.so")
    src = os.path.join(HERE, "dispatch_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return so


class OracleC:
    _lib = None

    def __init__(self, n_backends: int, capacity: int = 1, boost_mod: int = 2):
        if OracleC._lib is None:
            L = C.CDLL(build_c_oracle())
            L.orc_new.restype = C.c_void_p
            L.orc_new.argtypes = [C.c_int, C.c_int, C.c_int]
            L.orc_free.argtypes = [C.c_void_p]
            L.orc_enqueue.argtypes = [C.c_void_p, C.c_char_p]
            L.orc_set_vip.argtypes = [C.c_void_p, C.c_char_p]
            L.orc_set_boost.argtypes = [C.c_void_p, C.c_char_p]
            L.orc_add_vip.argtypes = [C.c_void_p, C.c_char_p]
            L.orc_add_boost.argtypes = [C.c_void_p, C.c_char_p]
            L.orc_set_online.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.orc_next.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_int)]
            L.orc_complete.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
            for f in ("orc_user_processed", "orc_user_dropped", "orc_user_queued"):
                getattr(L, f).restype = C.c_long
                getattr(L, f).argtypes = [C.c_void_p, C.c_char_p]
            for f in ("orc_backend_active", "orc_backend_processed"):
                getattr(L, f).restype = C.c_long
                getattr(L, f).argtypes = [C.c_void_p, C.c_int]
            L.orc_bench.restype = C.c_long
            L.orc_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
            OracleC._lib = L
        self.L = OracleC._lib
        self.h = self.L.orc_new(n_backends, capacity, boost_mod)
        self._buf = C.create_string_buffer(256)

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    def enqueue(self, user):
        self.L.orc_enqueue(self.h, None if user is None else user.encode())

    def set_vip(self, u):
        self.L.orc_set_vip(self.h, None if u is None else u.encode())

    def set_boost(self, u):
        self.L.orc_set_boost(self.h, None if u is None else u.encode())

    def add_vip(self, u):
        self.L.orc_add_vip(self.h, u.encode())

    def add_boost(self, u):
        self.L.orc_add_boost(self.h, u.encode())

    def set_online(self, b, online):
        self.L.orc_set_online(self.h, b, 1 if online else 0)

    def next(self):
        seq, be = C.c_long(), C.c_int()
        if not self.L.orc_next(self.h, self._buf, 256, C.byref(seq), C.byref(be)):
            return None
        return (self._buf.value.decode(), seq.value, be.value)

    def complete(self, backend, user, outcome=0):
        self.L.orc_complete(self.h, backend, user.encode(), outcome)


def simulate(orc, arrivals: List[Tuple[int, Optional[str]]], service_time: Callable[[str, int, int], int],
             vip=None, boost=None, outcomes=None, events=None, on_complete=None,
             on_dispatch=None) -> List[Tuple[str, int, int]]:
    """Event model of SURVEY.md 3.2, restated independently of the product harness.

    A completion = {user counter++, backend freed} atomically (dispatcher.rs:314-341 has no .await between
    them).  Completions at equal time are separate events, ordered by backend index then dispatch order; the
    scheduler runs to quiescence after every completion and after every batch of same-time arrivals;
    completions at time t precede arrivals at time t.  `events`: optional {time: [(kind, arg...)]} control
    events applied before anything else at that time (("vip", u), ("boost", u), ("online", b, flag)).
    """
    for v in ([vip] if isinstance(vip, str) else (vip or [])):
        orc.add_vip(v)
    for b in ([boost] if isinstance(boost, str) else (boost or [])):
        orc.add_boost(b)
    pend = sorted(range(len(arrivals)), key=lambda i: (arrivals[i][0], i))
    inflight = []  # [finish, backend, order, user, seq]
    out = []
    order = 0
    ai = 0
    ev_times = sorted(events) if events else []
    ei = 0
    INF = 1 << 62
    t = 0

    def run():
        nonlocal order
        while True:
            d = orc.next()
            if d is None:
                return
            out.append(d)
            if on_dispatch is not None:
                on_dispatch(t, *d)
            inflight.append([t + int(service_time(*d)), d[2], order, d[0], d[1]])
            order += 1

    while ai < len(pend) or inflight or ei < len(ev_times):
        ta = arrivals[pend[ai]][0] if ai < len(pend) else INF
        tc = min(x[0] for x in inflight) if inflight else INF
        te = ev_times[ei] if ei < len(ev_times) else INF
        t = min(ta, tc, te)
        if te == t:
            for ev in events[t]:
                if ev[0] == "vip":
                    orc.set_vip(ev[1])
                elif ev[0] == "boost":
                    orc.set_boost(ev[1])
                elif ev[0] == "online":
                    orc.set_online(ev[1], ev[2])
            ei += 1
        while True:
            due = [x for x in inflight if x[0] == t]
            if not due:
                break
            x = min(due, key=lambda y: (y[1], y[2]))
            inflight.remove(x)
            orc.complete(x[1], x[3], 0 if outcomes is None else outcomes(x[3], x[4]))
            if on_complete is not None:
                on_complete(t, x[1], x[3], x[4])
            run()
        if ta == t:
            while ai < len(pend) and arrivals[pend[ai]][0] == t:
                orc.enqueue(arrivals[pend[ai]][1])
                ai += 1
            run()
        elif te == t:
            run()
    return out
