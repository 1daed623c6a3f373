"""Host-side mirror of the reference's scheduling interface (AppState + run_worker, dispatcher.rs:49-96,
164-262) over the C ABI.  The arithmetic lives in csrc/sched.cpp; this file only marshals."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

from . import _lib
from ._lib import lib, check

PROCESSED, DROPPED, UNCOUNTED = 0, 1, 2


@dataclass(frozen=True)
class Dispatch:
    user: str
    user_seq: int
    backend: int
    task_id: int = 0

    def key(self):
        return (self.user, self.user_seq, self.backend)


class Scheduler:
    """`AppState` + the loop body of `run_worker` as a clock-less state machine."""

    def __init__(self, n_backends: int, capacity: int = 1, boost_mod: int = 2):
        self._h = lib.mq_sched_new(n_backends, capacity)
        if not self._h:
            raise _lib.MQError(-22, _lib.last_error())
        if boost_mod != 2:
            check(lib.mq_sched_set_boost_mod(self._h, boost_mod))
        self.n_backends = n_backends

    def close(self):
        if self._h:
            lib.mq_sched_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _b(s: Optional[str]):
        return None if s is None else s.encode("utf-8")

    def enqueue(self, user: Optional[str]) -> int:
        tid = C.c_uint64()
        check(lib.mq_sched_enqueue(self._h, self._b(user), C.byref(tid)))
        return tid.value

    def next(self) -> Optional[Dispatch]:
        d = _lib.Dispatch()
        rc = check(lib.mq_sched_next(self._h, C.byref(d)))
        if rc == 0:
            return None
        return Dispatch(d.user.decode("utf-8"), d.user_seq, d.backend, d.task_id)

    def drain(self) -> List[Dispatch]:
        """Run the loop until it would park (dispatcher.rs:344-349)."""
        out = []
        while True:
            d = self.next()
            if d is None:
                return out
            out.append(d)

    def complete(self, backend: int, user: str, outcome: int = PROCESSED):
        check(lib.mq_sched_complete(self._h, backend, self._b(user), outcome))

    def processing(self, user: str, delta: int):
        check(lib.mq_sched_processing(self._h, self._b(user), delta))

    def set_vip(self, user: Optional[str]):
        check(lib.mq_sched_set_vip(self._h, self._b(user)))

    def set_boost(self, user: Optional[str]):
        check(lib.mq_sched_set_boost(self._h, self._b(user)))

    def add_vip(self, user: str):
        """Extension (config 3): several VIPs; first in the reference sort order wins."""
        check(lib.mq_sched_add_vip(self._h, self._b(user)))

    def add_boost(self, user: str):
        check(lib.mq_sched_add_boost(self._h, self._b(user)))

    def set_online(self, backend: int, online: bool):
        check(lib.mq_sched_set_online(self._h, backend, 1 if online else 0))

    def set_capacity(self, capacity: int):
        check(lib.mq_sched_set_capacity(self._h, capacity))

    def user_stats(self, user: str) -> dict:
        st = _lib.UserStats()
        check(lib.mq_sched_user_stats(self._h, self._b(user), C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def backend_stats(self, backend: int) -> dict:
        st = _lib.BackendStats()
        check(lib.mq_sched_backend_stats(self._h, backend, C.byref(st)))
        return {"active_requests": st.active_requests, "processed_count": st.processed_count,
                "is_online": bool(st.is_online)}

    def users_tui_order(self) -> List[str]:
        n = lib.mq_sched_user_count(self._h)
        buf = C.create_string_buffer(_lib.MQ_USER_MAX)
        out = []
        for i in range(n):
            check(lib.mq_sched_user_name(self._h, i, buf, len(buf)))
            out.append(buf.value.decode("utf-8"))
        return out

    @property
    def counter(self) -> int:
        return lib.mq_sched_counter(self._h)


def simulate(sched: Scheduler, arrivals, service_time, vip=None, boost=None, outcomes=None):
    """Drive a scheduler on a simulated clock (event model of SURVEY.md 3.2).

    arrivals: list of (time, user) — tasks enqueued at integer times, in list order within a time.
    service_time(user, user_seq, backend) -> int >= 1 ticks.
    Completions at equal time are separate events ordered by backend index (then dispatch order); after each
    completion, and after each batch of same-time arrivals, the scheduler runs to quiescence.  Completions
    at time t come before arrivals at time t.  Returns the dispatch list.
    """
    import heapq

    for v in ([vip] if isinstance(vip, str) else (vip or [])):
        sched.add_vip(v)
    for b in ([boost] if isinstance(boost, str) else (boost or [])):
        sched.add_boost(b)
    arr = sorted(enumerate(arrivals), key=lambda x: (x[1][0], x[0]))
    ai = 0
    heap = []  # (finish_time, backend, order, user, seq)
    order = 0
    out: List[Dispatch] = []

    def run():
        nonlocal order
        for d in sched.drain():
            out.append(d)
            heapq.heappush(heap, (t + int(service_time(d.user, d.user_seq, d.backend)), d.backend, order, d.user,
                                  d.user_seq))
            order += 1

    while ai < len(arr) or heap:
        t = min(arr[ai][1][0] if ai < len(arr) else 1 << 62, heap[0][0] if heap else 1 << 62)
        # each completion is its own event: {processed_counts[u]++, backend freed} atomically
        # (dispatcher.rs:314-341 has no .await in between), then the scheduler runs to quiescence
        while heap and heap[0][0] == t:
            _, b, _, u, seq = heapq.heappop(heap)
            oc = PROCESSED if outcomes is None else outcomes(u, seq)
            sched.complete(b, u, oc)
            run()
        if ai < len(arr) and arr[ai][1][0] == t:
            while ai < len(arr) and arr[ai][1][0] == t:
                sched.enqueue(arr[ai][1][1])
                ai += 1
            run()
    return out
