"""Dispatch-decision micro-benchmark: decisions/s of the C++ scheduler (csrc/sched.cpp: incremental
sorted active list) next to the C restatement of the reference's loop (oracle/dispatch_oracle.c: collect, sort,
pick - what dispatcher.rs:211-245 does on every pass), same driver loop, same trace; CPU only.

    python tools/dispatch_bench.py [reqs_per_user]
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ollamamq_b200 as mq  # noqa: E402
from oracle.dispatch_oracle import OracleC  # noqa: E402


def product(users, reqs, backends, capacity):
    n, sec = C.c_uint64(), C.c_double()
    mq.check(mq.lib.mq_debug_sched_bench(users, reqs, backends, capacity, C.byref(n), C.byref(sec)))
    return n.value, sec.value


def oracle(users, reqs, backends, capacity):
    L = OracleC(1).L
    t0 = time.perf_counter()
    n = L.orc_bench(users, reqs, backends, capacity)
    return n, time.perf_counter() - t0          # includes the enqueue phase


if __name__ == "__main__":
    reqs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    print("# users backends capacity | dispatches | C++ scheduler decisions/s | oracle (re-sort per decision, incl. enqueue) decisions/s")
    for users, backends, capacity in ((4, 2, 1), (64, 1, 1), (64, 8, 1), (64, 1, 64), (256, 1, 1), (256, 8, 32)):
        n, sec = product(users, reqs, backends, capacity)
        no, seco = oracle(users, reqs, backends, capacity)
        assert n == no == users * reqs, (n, no)
        print("%5d %8d %8d | %10d | %12.0f | %12.0f" % (users, backends, capacity, n, n / sec, no / seco))
