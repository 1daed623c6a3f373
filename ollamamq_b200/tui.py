"""Terminal dashboard over a Dispatcher - the same four panes, key map and orderings as the reference's ratatui
dashboard (/root/reference/src/tui.rs:55-464), drawn with `curses` from one `mq_dispatcher_snapshot_json` call per
tick (tui.rs:97-107: snapshot -> draw -> poll 100 ms).

Everything that decides WHAT is shown or WHICH state a key changes is in pure functions (`rows_*`, `stats_line`,
`Dashboard.on_key`), tested without a terminal; `run()` is only the curses loop.

    python -m ollamamq_b200.tui            # demo against two mock backends
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

HELP_BAR = " Tab: Switch | p: VIP | b: Boost | x: Block User | X: Block IP | u: Unblock | q: Quit"   # tui.rs:455
HELP_DETAIL = ["", "  VIP: 'p' | BOOST: 'b' | BLOCK: 'x' (User) / 'X' (IP) | UNBLOCK: 'u'",
               "  PANELS: 'Tab' | QUIT: 'q' or 'Esc'", "",
               "  * VIP | ! Boost | x Blocked | > Processing | o Queued"]                            # tui.rs:460


def _is_blocked(snap: dict, u: dict) -> bool:
    return u["id"] in snap["blocked_users"] or (u["ip"] != "" and u["ip"] in snap["blocked_ips"])    # tui.rs:396


def stats_line(snap: dict, panel: str) -> str:
    """Top bar (tui.rs:318-347): panel, VIP, Boost, Q = queued + processing, Done, Drop."""
    users = snap["users"]
    q = sum(u["queued"] + u["processing"] for u in users)
    return " ollamaMQ  | Panel: %s | VIP: %s | Boost: %s | Q: %d | Done: %d | Drop: %d" % (
        "USERS" if panel == "users" else "BLOCKED", ",".join(snap["vip"]) or "None", ",".join(snap["boost"]) or "None",
        q, sum(u["processed"] for u in users), sum(u["dropped"] for u in users))


def rows_backends(snap: dict) -> List[Tuple[str, str, str]]:
    """'Ollama Instances' pane (tui.rs:349-385): status dot + label, active, done."""
    return [(("@ " if b["online"] else "o ") + b["label"], str(b["active"]), str(b["processed"])) for b in snap["backends"]]


def rows_users(snap: dict) -> List[Tuple[str, str, str, str, str]]:
    """'Active Users' pane (tui.rs:387-419), users already in the order of tui.rs:70-80; the marker precedence is
    blocked > VIP > Boost > processing > queued > idle (tui.rs:400-405)."""
    out = []
    for u in snap["users"]:
        blocked, vip, boost = _is_blocked(snap, u), u["id"] in snap["vip"], u["id"] in snap["boost"]
        sym = "x " if blocked else "* " if vip else "! " if boost else "> " if u["processing"] > 0 else \
              "o " if u["queued"] > 0 else ". "
        name = sym + u["id"] + (" [VIP]" if vip else "") + (" [BST]" if boost else "") + (" [BLOCKED]" if blocked else "")
        out.append((name, u["ip"], str(u["queued"] + u["processing"]), str(u["processed"]), str(u["dropped"])))
    return out


def rows_queues(snap: dict, width: int) -> List[Tuple[str, str, str]]:
    """'Queue Status' pane (tui.rs:421-437): bar = min(q / 20, 1) of 45 % of the pane width, share of all queued."""
    total = sum(u["queued"] + u["processing"] for u in snap["users"])
    bar_max = int(width * 0.45)
    out = []
    for u in snap["users"]:
        q = u["queued"] + u["processing"]
        bar = int(min(q / 20.0, 1.0) * bar_max) if q > 0 else 0
        out.append((u["id"], ("#" * bar).ljust(bar_max), "%d (%.0f%%)" % (q, 100.0 * q / total if total else 0.0)))
    return out


def rows_blocked(snap: dict) -> List[Tuple[str, str]]:
    """'Blocked Items' pane (tui.rs:439-452): IPs and users together, sorted by value."""
    items = [("IP", ip) for ip in snap["blocked_ips"]] + [("USER", u) for u in snap["blocked_users"]]
    return sorted(items, key=lambda kv: kv[1].encode())


class Dashboard:
    """Selection state + key handling of tui.rs:97-267.  `d` needs set_vip / set_boost / block_user / block_ip /
    snapshot (ollamamq_b200.Dispatcher)."""

    def __init__(self, d):
        self.d = d
        self.panel = "users"
        self.sel_user: Optional[int] = None
        self.sel_blocked: Optional[int] = None
        self.show_help = False

    def clamp(self, snap: dict):
        """Selection defaults of render() (tui.rs:270-284)."""
        if self.panel == "users":
            if not snap["users"]:
                self.sel_user = None
            elif self.sel_user is None:
                self.sel_user = 0
        else:
            n = len(snap["blocked_ips"]) + len(snap["blocked_users"])
            if n == 0:
                self.sel_blocked = None
            elif self.sel_blocked is None:
                self.sel_blocked = 0

    def _user(self, snap: dict) -> Optional[dict]:
        if self.panel == "users" and self.sel_user is not None and self.sel_user < len(snap["users"]):
            return snap["users"][self.sel_user]
        return None

    def on_key(self, key: str, snap: dict) -> bool:
        """Apply one key press against the snapshot it was drawn from; returns False on quit."""
        if key in ("q", "ESC"):
            return False
        if key == "?":
            self.show_help = not self.show_help
        elif key in ("TAB", "l", "h"):
            self.panel = "blocked" if self.panel == "users" else "users"
        elif key == "p":                                   # toggle VIP; taking VIP drops the user's Boost (:126-150)
            u = self._user(snap)
            if u:
                if u["id"] in snap["vip"]:
                    self.d.set_vip(None)
                else:
                    self.d.set_vip(u["id"])
                    if u["id"] in snap["boost"]:
                        self.d.set_boost(None)
        elif key == "b":                                   # toggle Boost; taking Boost drops the user's VIP (:153-177)
            u = self._user(snap)
            if u:
                if u["id"] in snap["boost"]:
                    self.d.set_boost(None)
                else:
                    self.d.set_boost(u["id"])
                    if u["id"] in snap["vip"]:
                        self.d.set_vip(None)
        elif key == "x":                                   # block the selected user (:180-188)
            u = self._user(snap)
            if u:
                self.d.block_user(u["id"], True)
        elif key == "X":                                   # block the selected user's last IP (:190-200)
            u = self._user(snap)
            if u and u["ip"]:
                self.d.block_ip(u["ip"], True)
        elif key == "u":                                   # unblock (:202-237)
            if self.panel == "blocked":
                items = rows_blocked(snap)
                if self.sel_blocked is not None and self.sel_blocked < len(items):
                    kind, val = items[self.sel_blocked]
                    (self.d.block_ip if kind == "IP" else self.d.block_user)(val, False)
            else:
                u = self._user(snap)
                if u:
                    self.d.block_user(u["id"], False)
                    if u["ip"]:
                        self.d.block_ip(u["ip"], False)
        elif key in ("UP", "k"):                           # saturating moves (:239-262)
            if self.panel == "users":
                self.sel_user = max(0, (self.sel_user or 0) - 1)
            else:
                self.sel_blocked = max(0, (self.sel_blocked or 0) - 1)
        elif key in ("DOWN", "j"):
            if self.panel == "users":
                n = len(snap["users"])
                if n:
                    self.sel_user = 0 if self.sel_user is None else min(self.sel_user + 1, n - 1)
            else:
                n = len(snap["blocked_ips"]) + len(snap["blocked_users"])
                if n:
                    self.sel_blocked = 0 if self.sel_blocked is None else min(self.sel_blocked + 1, n - 1)
        return True

    # ---------------------------------------------------------------- text frame (what run() paints)
    def frame(self, snap: dict, width: int = 120, height: int = 30) -> List[str]:
        """The whole screen as text lines: stats bar, three columns 25 / 40 / 35 % (right one split 60 / 40),
        help bar (tui.rs:286-316).  Used by the tests and by `run()`."""
        self.clamp(snap)
        w1, w2 = width * 25 // 100, width * 40 // 100
        w3 = width - w1 - w2
        body_h = max(4, height - 2 - (len(HELP_DETAIL) + 1 if self.show_help else 0))
        top_h = body_h * 60 // 100

        def table(title, header, rows, w, h, sel=None):
            cols = len(header)
            cw = [max(4, (w - 4) // cols)] * cols
            cw[0] = max(4, w - 4 - sum(cw[1:]))
            fmt = lambda r: "".join(str(c)[:cw[i] - 1].ljust(cw[i]) for i, c in enumerate(r))
            lines = [("[" + title + "]").ljust(w)[:w], ("   " + fmt(header))[:w].ljust(w)]
            for i, r in enumerate(rows[: max(0, h - 2)]):
                lines.append(((">> " if sel == i else "   ") + fmt(r))[:w].ljust(w))
            return lines + [" " * w] * (h - len(lines))

        c1 = table(" Ollama Instances ", ("Backend", "Act", "Done"), rows_backends(snap), w1, body_h)
        c2 = table(" Active Users ", ("User ID", "Last IP", "Q", "Done", "Drop"), rows_users(snap), w2, body_h,
                   self.sel_user if self.panel == "users" else None)
        c3 = table(" Queue Status ", ("User ID", "Progress", "Num"), rows_queues(snap, w3), w3, top_h,
                   self.sel_user if self.panel == "users" else None)
        c3 += table(" Blocked Items ", ("Type", "Value"), rows_blocked(snap), w3, body_h - top_h,
                    self.sel_blocked if self.panel == "blocked" else None)
        lines = [stats_line(snap, self.panel)[:width]]
        lines += [a + b + c for a, b, c in zip(c1, c2, c3)]
        lines.append(HELP_BAR[:width])
        if self.show_help:
            lines += [" Help "] + [h[:width] for h in HELP_DETAIL]
        return lines

    def run(self, tick_ms: int = 100):  # pragma: no cover - needs a terminal
        import curses

        keymap = {27: "ESC", 9: "TAB", curses.KEY_UP: "UP", curses.KEY_DOWN: "DOWN"}

        def loop(scr):
            curses.curs_set(0)
            scr.timeout(tick_ms)                                           # event::poll(100 ms) (tui.rs:107)
            while True:
                snap = self.d.snapshot()
                h, w = scr.getmaxyx()
                scr.erase()
                for y, line in enumerate(self.frame(snap, w - 1, h - 1)[: h - 1]):
                    scr.addstr(y, 0, line[: w - 1], curses.A_BOLD if y == 0 else curses.A_NORMAL)
                scr.refresh()
                c = scr.getch()
                if c == -1:
                    continue
                key = keymap.get(c) or (chr(c) if 0 < c < 256 else None)
                if key and not self.on_key(key, snap):
                    return

        curses.wrapper(loop)


if __name__ == "__main__":  # pragma: no cover
    import ollamamq_b200 as mq

    d = mq.Dispatcher(mock_backends=2, capacity=1)
    for i in range(12):
        d.submit("user%d" % (i % 4), ip="10.0.0.%d" % (i % 4), prompt_tokens=[1, 2, 3], max_new_tokens=1)
    try:
        Dashboard(d).run()
    finally:
        d.close()
