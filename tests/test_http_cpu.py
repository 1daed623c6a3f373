"""HTTP ingress (csrc/http_server.cpp) on CPU over the live dispatcher with mock backends — SURVEY.md 8(f) rank 1.
Checks the reference's observable HTTP behaviour: main.rs:89-121 (routes, /health, fallback) and
dispatcher.rs:354-428 (X-User-ID, 403 / 500 bodies, streamed relay, query string dropped)."""
import http.client
import json
import socket
import threading
import time

import pytest

import ollamamq_b200 as mq


class Served:
    def __init__(self, backends=2, allow_all=False, auto=True):
        self.d = mq.Dispatcher(mock_backends=backends, capacity=1)
        self.port = self.d.serve_http(0, "127.0.0.1", allow_all)
        self.stop = False
        self.auto = auto
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):  # the mock backends finish a request only when told to
        while not self.stop:
            if self.auto:
                for b in range(self.d.n_backends):
                    self.d.mock_complete(b)
            time.sleep(0.002)

    def close(self):
        self.stop = True
        self.t.join()
        self.d.close()

    def request(self, method, path, body=None, headers=None):
        c = http.client.HTTPConnection("127.0.0.1", self.port, timeout=10)
        c.request(method, path, body=body, headers=headers or {})
        r = c.getresponse()
        data = r.read()
        c.close()
        return r.status, dict(r.getheaders()), data


@pytest.fixture
def srv():
    s = Served()
    yield s
    s.close()


def test_health_is_not_queued(srv):
    st, h, body = srv.request("GET", "/health")
    assert (st, body) == (200, b"OK")
    assert srv.d.log() == []


def test_route_table_and_fallback(srv):
    assert srv.request("POST", "/nope")[0] == 404
    assert srv.request("GET", "/api/blobs/a/b")[0] == 404           # {digest} is one segment
    for p in ["/", "/api/tags", "/api/version", "/v1/models", "/v1/models/llama", "/api/blobs/sha256:abc", "/api/ps"]:
        assert srv.request("GET", p)[0] == 200, p
    s2 = Served(allow_all=True)
    try:
        assert s2.request("GET", "/anything/else")[0] == 200          # --allow-all-routes: proxied like the rest
    finally:
        s2.close()


def test_streamed_relay_user_header_and_query_dropped(srv):
    body = json.dumps({"model": "m", "messages": [{"role": "user", "content": "Req 1"}], "stream": True}).encode()
    st, h, data = srv.request("POST", "/api/chat?x=1", body, {"X-User-ID": "alice", "Content-Type": "application/json"})
    assert st == 200 and h.get("Transfer-Encoding") == "chunked"
    assert data.startswith(b'{"tok":0}')                              # http.client de-chunks the body
    st, h, data = srv.request("POST", "/api/generate", b'{"prompt":"hi"}')   # no header -> "anonymous" (:364-368)
    assert st == 200
    users = [u for u, _, _ in srv.d.log()]
    assert users == ["alice", "anonymous"]
    assert srv.d.user_stats("alice")["processed"] == 1


def test_keep_alive_two_requests_one_connection(srv):
    c = http.client.HTTPConnection("127.0.0.1", srv.port, timeout=10)
    for i in range(2):
        c.request("POST", "/v1/completions", body=b'{"prompt":"x"}', headers={"X-User-ID": "bob"})
        r = c.getresponse()
        assert r.status == 200 and r.read()
    c.close()
    assert srv.d.user_stats("bob")["processed"] == 2


def test_blocked_user_and_ip_get_403_with_reference_bodies(srv):
    srv.d.block_user("mallory")
    st, _, body = srv.request("POST", "/api/chat", b"{}", {"X-User-ID": "mallory"})
    assert (st, body) == (403, b"User blocked")
    srv.d.block_ip("127.0.0.1")
    st, _, body = srv.request("POST", "/api/chat", b"{}", {"X-User-ID": "alice"})
    assert (st, body) == (403, b"IP blocked")
    assert srv.request("GET", "/health")[0] == 200                   # /health never goes through proxy_handler


def test_backend_error_is_http_500(srv):
    srv.d.mock_fail_next(0, 1)
    srv.d.mock_fail_next(1, 1)
    st, _, body = srv.request("POST", "/api/chat", b"{}", {"X-User-ID": "carol"})
    assert st == 500 and body.startswith(b"Backend error:")
    assert srv.d.user_stats("carol")["dropped"] == 1


def test_client_gone_while_queued_is_dropped():
    s = Served(backends=1, auto=False)
    try:
        s.d.set_online(0, False)
        sk = socket.create_connection(("127.0.0.1", s.port))
        sk.sendall(b"POST /api/chat HTTP/1.1\r\nHost: x\r\nX-User-ID: dave\r\nContent-Length: 2\r\n\r\n{}")
        for _ in range(200):                                          # wait until it sits in dave's queue
            try:
                if s.d.user_stats("dave")["queued"] == 1:
                    break
            except mq.MQError:
                pass
            time.sleep(0.01)
        sk.close()
        time.sleep(0.2)                                               # connection thread notices POLLRDHUP
        s.d.set_online(0, True)
        s.d.submit("eve", max_new_tokens=1)                           # a notify wakes the scheduler
        s.d.wait_parked()
        while s.d.mock_complete(0):
            s.d.wait_parked()
        s.d.drain(5000)
        st = s.d.user_stats("dave")
        assert (st["processed"], st["dropped"]) == (0, 1)             # popped, dropped at pre-flight (:278-280)
    finally:
        s.close()


def test_malformed_requests_do_not_take_the_server_down(srv):
    """Garbage on the wire: every connection is answered or closed, and the server keeps serving."""
    import random
    rnd = random.Random(1)
    blobs = [b"\r\n\r\n", b"GET\r\n\r\n", b"GET /health\r\n\r\n", b"POST /api/chat HTTP/1.1\r\nContent-Length: 99999999999\r\n\r\n",
             b"POST /api/chat HTTP/1.1\r\nTransfer-Encoding: chunked\r\n\r\n5\r\nhello\r\n0\r\n\r\n",
             b"POST /api/chat HTTP/1.1\r\nContent-Length: -5\r\n\r\n", b"\x00" * 3000, b"A" * 70000,
             b"POST /api/chat HTTP/1.1\r\nX-User-ID: " + b"u" * 5000 + b"\r\nContent-Length: 2\r\n\r\n{}",
             b"POST /api/embed HTTP/1.1\r\nContent-Length: 7\r\n\r\n\xff\xfe{\"a\":"]
    blobs += [bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 600))) + b"\r\n\r\n" for _ in range(40)]
    for blob in blobs:
        s = socket.create_connection(("127.0.0.1", srv.port), timeout=5)
        try:
            s.sendall(blob)
            s.shutdown(socket.SHUT_WR)
            s.settimeout(5)
            while s.recv(65536):
                pass
        except (ConnectionError, socket.timeout, OSError):
            pass
        finally:
            s.close()
    st, _, body = srv.request("GET", "/health")
    assert st == 200 and body == b"OK"
    st, _, _ = srv.request("POST", "/api/chat", body=b'{"model":"m","messages":[]}', headers={"X-User-ID": "after-fuzz"})
    assert st == 200


def _admin(srv, path, obj=None, method="POST"):
    st, _, body = srv.request(method, path, json.dumps(obj).encode() if obj is not None else None,
                              {"Content-Type": "application/json"})
    return st, body


def test_admin_control_surface_has_the_dashboard_key_semantics(srv):
    """Headless control path: POST /admin/{vip,boost,block,unblock} = the p / b / x / X / u keys of
    tui.rs:126-237 - VIP and Boost toggle, setting one clears the other when it names the same user; GET /admin/state is
    the dashboard snapshot.  With one VIP / one Boost the scheduler is exactly the reference's."""
    for u in ("alice", "bob"):                                        # users appear once they have sent something
        assert srv.request("POST", "/api/chat", b"{}", {"X-User-ID": u})[0] == 200
    state = lambda: json.loads(_admin(srv, "/admin/state", method="GET")[1])
    assert state()["vip"] == [] and [u["id"] for u in state()["users"]] == ["alice", "bob"]
    assert _admin(srv, "/admin/vip", {"user": "alice"})[0] == 200
    assert state()["vip"] == ["alice"]
    assert _admin(srv, "/admin/boost", {"user": "alice"})[0] == 200   # 'b' on the VIP user: Boost set, VIP cleared
    assert state()["vip"] == [] and state()["boost"] == ["alice"]
    assert _admin(srv, "/admin/vip", {"user": "bob"})[0] == 200
    assert _admin(srv, "/admin/vip", {"user": "alice"})[0] == 200     # one VIP slot: replaces bob; clears alice's Boost
    assert state()["vip"] == ["alice"] and state()["boost"] == []
    assert _admin(srv, "/admin/vip", {"user": "alice"})[0] == 200     # 'p' again: toggled off
    assert state()["vip"] == []
    # extension (BASELINE config 3): sets
    assert _admin(srv, "/admin/vip", {"user": "alice", "mode": "add"})[0] == 200
    assert _admin(srv, "/admin/vip", {"user": "bob", "mode": "add"})[0] == 200
    assert sorted(state()["vip"]) == ["alice", "bob"]
    assert _admin(srv, "/admin/vip", {"mode": "clear"})[0] == 200 and state()["vip"] == []
    # block / unblock: 'x' blocks the user, 'X' its last address, 'u' lifts both
    assert _admin(srv, "/admin/block", {"user": "bob"})[0] == 200
    assert srv.request("POST", "/api/chat", b"{}", {"X-User-ID": "bob"})[:1] == (403,)
    assert state()["blocked_users"] == ["bob"]
    assert _admin(srv, "/admin/block", {"user": "alice", "mode": "ip"})[0] == 200
    assert state()["blocked_ips"] == ["127.0.0.1"]
    assert srv.request("POST", "/api/chat", b"{}", {"X-User-ID": "carol"})[2] == b"IP blocked"
    assert _admin(srv, "/admin/unblock", {"user": "alice"})[0] == 200 # admin routes are not behind the block list
    assert state()["blocked_ips"] == []
    assert _admin(srv, "/admin/unblock", {"user": "bob"})[0] == 200 and state()["blocked_users"] == []
    assert srv.request("POST", "/api/chat", b"{}", {"X-User-ID": "bob"})[0] == 200
    assert _admin(srv, "/admin/block", {"ip": "10.1.2.3"})[0] == 200 and state()["blocked_ips"] == ["10.1.2.3"]
    assert _admin(srv, "/admin/unblock", {"ip": "10.1.2.3"})[0] == 200 and state()["blocked_ips"] == []
    # errors
    assert _admin(srv, "/admin/vip", {})[0] == 400 and _admin(srv, "/admin/nope", {"user": "a"})[0] == 404
    assert srv.request("GET", "/admin/vip")[0] == 405


def test_vip_set_over_http_changes_dispatch_order_like_the_reference():
    """config 1 with vip = charlie set through the admin route: the dispatch log equals the oracle's."""
    from oracle.dispatch_oracle import OracleC, simulate
    s = Served(backends=2, auto=False)
    try:
        users = ["alice", "bob", "charlie", "david"]
        for b in range(2):
            s.d.set_online(b, False)                                  # queue everything first (the t=0 arrival trace)
        assert _admin(s, "/admin/vip", {"user": "charlie"})[0] == 200
        streams = [s.d.submit(u, max_new_tokens=1) for u in users for _ in range(8)]
        for b in range(2):
            s.d.set_online(b, True)
        s.d.submit("zz-wake", max_new_tokens=1)                       # a notify wakes the scheduler (recovery does not)
        s.d.wait_parked()
        for _ in range(200):
            progressed = False
            for b in range(2):                                        # completions ordered by backend index, one per event
                if s.d.mock_complete(b):
                    progressed = True
                    s.d.wait_parked()
            if not progressed:
                break
        s.d.drain(5000)
        got = [(u, seq, b) for u, seq, b in s.d.log() if u != "zz-wake"]
        arr = [(0, u) for u in users for _ in range(8)]
        ref = [r for r in simulate(OracleC(2), arr + [(0, "zz-wake")], lambda u, q, b: 1, vip="charlie") if r[0] != "zz-wake"]
        assert [g[0] for g in got][:8] == ["charlie"] * 8             # VIP absolute priority (:230)
        assert sorted(got) == sorted(ref)
    finally:
        s.close()


def test_a_thousand_connections_on_one_loop():
    """1 024 simultaneous keep-alive connections, every one with a request in flight, then answered: the epoll loop holds
    them all (round 1 ran a thread per connection), and stop() joins it with requests still open."""
    s = Served(backends=4, auto=False)
    N = 1024
    try:
        socks = []
        for i in range(N):
            sk = socket.create_connection(("127.0.0.1", s.port), timeout=20)
            body = b'{"prompt":"hello %d"}' % i
            sk.sendall(b"POST /api/generate HTTP/1.1\r\nHost: x\r\nX-User-ID: u%03d\r\nContent-Length: %d\r\n\r\n" % (i % 100, len(body)) + body)
            socks.append(sk)
        deadline = time.time() + 20
        while time.time() < deadline:
            snap = s.d.snapshot()
            if sum(u["queued"] + u["processing"] for u in snap["users"]) == N:
                break
            time.sleep(0.02)
        assert sum(u["queued"] + u["processing"] for u in s.d.snapshot()["users"]) == N
        s.auto = True                                                 # now let the mock backends answer
        got = 0
        for sk in socks:
            data = b""
            while b"0\r\n\r\n" not in data:
                chunk = sk.recv(65536)
                assert chunk, "connection closed before the response ended"
                data += chunk
            assert data.startswith(b"HTTP/1.1 200 OK") and b'{"tok":0}' in data
            got += 1
        assert got == N
        # second request on every 8th connection: still alive after the burst
        for sk in socks[::8]:
            sk.sendall(b"GET /health HTTP/1.1\r\nHost: x\r\n\r\n")
            assert sk.recv(4096).startswith(b"HTTP/1.1 200 OK")
        for sk in socks:
            sk.close()
    finally:
        s.close()


def test_malformed_json_body_is_a_400_and_dispatch_goes_on(srv):
    """ADVICE.md (round 1): `{"messages":[{"content":[}` froze the scheduler thread for every user and GPU.  The body is now
    parsed on the connection thread; the request still takes its fair-share turn and is answered like a backend would
    answer it (400 + an error object, counted as processed, :314-316), and the next request is served."""
    st, _, body = srv.request("POST", "/api/chat", b'{"messages":[{"content":[}', {"X-User-ID": "mallory"})
    assert st == 400 and json.loads(body)["error"]
    st, _, body = srv.request("POST", "/api/chat", b'{"x":' + b"[" * 200000, {"X-User-ID": "mallory"})
    assert st == 400
    assert srv.d.user_stats("mallory")["processed"] == 2
    st, _, data = srv.request("POST", "/api/chat", b'{"messages":[{"role":"user","content":"hi"}]}', {"X-User-ID": "alice"})
    assert st == 200 and data.startswith(b'{"tok":0}')
