"""Dispatch-decision micro-benchmark: decisions/s of the C++ scheduler (csrc/sched.cpp: incremental
sorted active list); CPU only.  The comparison with the C restatement of the reference's loop (collect, sort, pick on
every pass - dispatcher.rs:211-245) lives in tests/test_dispatch.py::test_decision_bench_matches_the_oracle_driver,
because only tests may execute oracle/.

    python tools/dispatch_bench.py [reqs_per_user]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ollamamq_b200 as mq  # noqa: E402


def product(users, reqs, backends, capacity):
    n, sec = C.c_uint64(), C.c_double()
    mq.check(mq.lib.mq_debug_sched_bench(users, reqs, backends, capacity, C.byref(n), C.byref(sec)))
    return n.value, sec.value


if __name__ == "__main__":
    reqs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    print("# users backends capacity | dispatches | C++ scheduler decisions/s")
    for users, backends, capacity in ((4, 2, 1), (64, 1, 1), (64, 8, 1), (64, 1, 64), (256, 1, 1), (256, 8, 32)):
        n, sec = product(users, reqs, backends, capacity)
        assert n == users * reqs, n
        print("%5d %8d %8d | %10d | %12.0f" % (users, backends, capacity, n, n / sec))
