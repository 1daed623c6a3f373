"""CPU: the dashboard - snapshot order, pane contents and key semantics of the reference's
tui.rs:55-267, against a dispatcher over step-driven mock backends (no GPU, no terminal)."""
import ollamamq_b200 as mq
from ollamamq_b200 import tui


def _dispatcher():
    d = mq.Dispatcher(mock_backends=2, capacity=1)
    # alice 3 tasks, bob 2, carol 1 (+ an IP each); two of them are dispatched at once (2 backends, capacity 1)
    for user, n, ip in (("alice", 3, "10.0.0.1"), ("bob", 2, "10.0.0.2"), ("carol", 1, "10.0.0.3")):
        for _ in range(n):
            d.submit(user, ip=ip, prompt_tokens=[1, 2, 3], max_new_tokens=1)
    d.wait_parked()
    return d


def test_snapshot_matches_the_reference_dashboard_order_and_totals():
    d = _dispatcher()
    try:
        snap = d.snapshot()
        ids = [u["id"] for u in snap["users"]]
        # tui.rs:70-80: queued + processing desc, then processed + dropped desc, then name asc
        assert ids == ["alice", "bob", "carol"]
        assert sum(u["queued"] + u["processing"] for u in snap["users"]) == 6
        assert sum(u["processing"] for u in snap["users"]) == 2
        assert [b["active"] for b in snap["backends"]] == [1, 1] and all(b["online"] for b in snap["backends"])
        assert snap["users"][0]["ip"] == "10.0.0.1" and snap["vip"] == [] and snap["boost"] == []
        # complete a few: processed counts move users within equal queue depth
        assert d.mock_complete(0) and d.mock_complete(1)
        d.wait_parked()
        snap = d.snapshot()
        by = {u["id"]: u for u in snap["users"]}
        assert sum(u["processed"] for u in snap["users"]) == 2
        assert tui.stats_line(snap, "users").endswith("Q: 4 | Done: 2 | Drop: 0")
        key = lambda u: (-(u["queued"] + u["processing"]), -(u["processed"] + u["dropped"]), u["id"].encode())
        assert [u["id"] for u in snap["users"]] == [u["id"] for u in sorted(by.values(), key=key)]
    finally:
        d.close()


def test_key_map_vip_boost_block_unblock_navigation():
    d = _dispatcher()
    try:
        dash = tui.Dashboard(d)
        snap = d.snapshot()
        dash.clamp(snap)
        assert dash.sel_user == 0 and dash.panel == "users"
        assert dash.on_key("p", snap)                       # alice becomes VIP
        snap = d.snapshot()
        assert snap["vip"] == ["alice"] and "[VIP]" in tui.rows_users(snap)[0][0]
        assert dash.on_key("b", snap)                       # Boost on the same user takes her VIP away (tui.rs:169-175)
        snap = d.snapshot()
        assert snap["boost"] == ["alice"] and snap["vip"] == []
        dash.on_key("j", snap)
        dash.on_key("p", snap)                              # bob VIP: alice keeps Boost (different user)
        snap = d.snapshot()
        assert snap["vip"] == ["bob"] and snap["boost"] == ["alice"]
        dash.on_key("p", snap)                              # toggles off
        assert d.snapshot()["vip"] == []
        for _ in range(5):
            dash.on_key("j", snap)                          # saturates at the last row
        assert dash.sel_user == 2
        dash.on_key("x", snap)                              # block carol (user)
        dash.on_key("X", snap)                              # ... and her IP
        snap = d.snapshot()
        assert snap["blocked_users"] == ["carol"] and snap["blocked_ips"] == ["10.0.0.3"]
        assert tui.rows_users(snap)[2][0].startswith("x carol") and "[BLOCKED]" in tui.rows_users(snap)[2][0]
        assert tui.rows_blocked(snap) == [("IP", "10.0.0.3"), ("USER", "carol")]
        dash.on_key("TAB", snap)
        dash.clamp(snap)
        assert dash.panel == "blocked" and dash.sel_blocked == 0
        dash.on_key("u", snap)                              # first blocked item is the IP
        snap = d.snapshot()
        assert snap["blocked_ips"] == [] and snap["blocked_users"] == ["carol"]
        dash.on_key("TAB", snap)
        dash.on_key("u", snap)                              # users panel: unblocks the selected user and her IP
        assert d.snapshot()["blocked_users"] == []
        for _ in range(9):
            dash.on_key("k", snap)
        assert dash.sel_user == 0
        assert dash.on_key("?", snap) and dash.show_help
        assert dash.on_key("q", snap) is False and dash.on_key("ESC", snap) is False
    finally:
        d.close()


def test_frame_has_the_four_panes_and_the_help_bar():
    d = _dispatcher()
    try:
        dash = tui.Dashboard(d)
        snap = d.snapshot()
        lines = dash.frame(snap, width=120, height=30)
        text = "\n".join(lines)
        for title in (" Ollama Instances ", " Active Users ", " Queue Status ", " Blocked Items "):
            assert title in text
        assert lines[0].startswith(" ollamaMQ  | Panel: USERS | VIP: None | Boost: None | Q: 6")
        assert tui.HELP_BAR in text and all(len(x) <= 120 for x in lines)
        assert any(line.startswith(" " * 0) and ">> " in line and "alice" in line for line in lines)
        q = tui.rows_queues(snap, 40)
        assert q[0][0] == "alice" and q[0][2] == "3 (50%)" and q[0][1].count("#") == int(min(3 / 20.0, 1.0) * 18)
        assert tui.rows_backends(snap)[0] == ("@ gpu0", "1", "0")
    finally:
        d.close()
