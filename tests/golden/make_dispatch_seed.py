"""Extracts the three hand-derived dispatch traces of SURVEY.md section 3.2 into dispatch_seed.json.

These are the ONLY vectors available for dispatch order: the reference ships none and cannot be compiled in
this image (no Rust).  They were derived by reading /root/reference/src/dispatcher.rs:195-262,314-341, not by
running the reference, and are labelled as such in the fixture.
Run from the repo root:  python tests/golden/make_dispatch_seed.py
"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
txt = open(os.path.join(ROOT, "SURVEY.md")).read()
blk = txt[txt.index("plain (32 dispatches"):]
blk = blk[: blk.index("```")]
sections = re.split(r"^(plain|vip=\w+|boost=\w+)[^\n]*:\s*$", blk, flags=re.M)
out = {"provenance": "hand-derived from dispatcher.rs by reading (SURVEY.md 3.2); NOT reference output",
       "users": ["alice", "bob", "charlie", "david"], "requests_per_user": 8, "n_backends": 2,
       "capacity": 1, "service_time": 1, "event_model": "SURVEY.md 3.2", "traces": {}}
tok = re.compile(r"(\w+)#(\d+)>b(\d+)")
it = iter(sections[1:])
for name, body in zip(it, it):
    out["traces"][name.strip()] = [[u, int(s), int(b)] for u, s, b in tok.findall(body)]
for k, v in out["traces"].items():
    assert len(v) == 32, (k, len(v))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "dispatch_seed.json"), "w"), indent=1)
print({k: len(v) for k, v in out["traces"].items()})
