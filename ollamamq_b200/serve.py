"""`python -m ollamamq_b200.serve` - the reference binary's command line (/root/reference/src/main.rs:19-41, 48-153)
over the B200 library: where the reference takes `--ollama-urls`, this takes the GPUs to open workers on.

    python -m ollamamq_b200.serve --model llama3-8b --gpus 0,1 --port 11435 --timeout 300 [--no-tui] [--allow-all-routes]
                                  [--embed-model bge-small] [--weights DIR]

Weights: `--weights DIR` holds one `<tensor name>.pt` (torch bf16 tensor) per tensor of the C ABI naming
(include/ollamamq_b200.h); without it the worker is random-initialised (benchmarks, smoke tests).
Like the reference: the dashboard runs when stdout is a terminal and `--no-tui` is not given (main.rs:56), `q` in the
dashboard stops the process (main.rs:136-150), blocked items persist in ./blocked_items.json (dispatcher.rs:19).
"""
from __future__ import annotations

import argparse
import os
import signal
import sys
import threading

from . import Dispatcher, Encoder, Worker, encoder_cfg, model_cfg, models
from .tui import Dashboard

MODELS = {"llama3-8b": models.LLAMA3_8B, "qwen2.5-7b": models.QWEN25_7B, "phi3-mini": models.PHI3_MINI}
EMBED_MODELS = {"bge-small": models.BGE_SMALL}


def parse_args(argv=None):
    ap = argparse.ArgumentParser(prog="ollamamq_b200.serve", description=__doc__.split("\n")[0])
    ap.add_argument("-p", "--port", type=int, default=11435, help="port to listen on (reference default 11435)")
    ap.add_argument("--bind", default="0.0.0.0")
    ap.add_argument("-g", "--gpus", default="0", help="comma-separated GPU ids, one worker each (replaces --ollama-urls)")
    ap.add_argument("-t", "--timeout", type=int, default=300, help="whole-request timeout in seconds (reference default 300)")
    ap.add_argument("--no-tui", action="store_true", help="disable the dashboard")
    ap.add_argument("--allow-all-routes", action="store_true", help="answer routes outside the table instead of 404")
    ap.add_argument("--model", default="llama3-8b", choices=sorted(MODELS))
    ap.add_argument("--embed-model", default=None, choices=sorted(EMBED_MODELS), help="also serve the /api/embed routes")
    ap.add_argument("--weights", default=None, help="directory of <tensor>.pt files; default: random init")
    ap.add_argument("--capacity", type=int, default=0, help="requests in flight per GPU (0 = the worker's batch size; "
                    "1 = the reference's one-at-a-time behaviour)")
    ap.add_argument("--max-batch", type=int, default=64)
    ap.add_argument("--max-seq", type=int, default=4096)
    ap.add_argument("--block-file", default="blocked_items.json")
    return ap.parse_args(argv)


def _load(target, directory):
    import torch
    loaded = 0
    for fn in sorted(os.listdir(directory)):
        if fn.endswith(".pt"):
            try:
                target.load_weights({fn[:-3]: torch.load(os.path.join(directory, fn)).to(torch.bfloat16)})
                loaded += 1
            except Exception:  # a tensor of the other model in the same directory
                continue
    return loaded


def main(argv=None) -> int:
    a = parse_args(argv)
    gpus = [int(g) for g in a.gpus.split(",") if g != ""]
    workers, encoders = [], []
    for g in gpus:
        w = Worker(g, model_cfg(MODELS[a.model], max_batch=a.max_batch, max_seq=a.max_seq, max_prefill_tokens=4736,
                                use_graphs=1, use_pdl=1, model_name=a.model))
        if a.weights:
            _load(w, a.weights)
        else:
            w.init_random(seed=0)
        workers.append(w)
    d = Dispatcher(workers, capacity=a.capacity or a.max_batch)
    d.set_timeout(a.timeout)
    d.set_block_file(a.block_file)
    d.start_health(10000)
    if a.embed_model:
        for i, g in enumerate(gpus):
            e = Encoder(g, encoder_cfg(EMBED_MODELS[a.embed_model], model_name=a.embed_model))
            if a.weights:
                _load(e, a.weights)
            else:
                e.init_random(seed=0)
            d.attach_encoder(i, e)
            encoders.append(e)
    port = d.serve_http(port=a.port, bind=a.bind, allow_all_routes=a.allow_all_routes)
    print("ollamamq_b200: %d GPU worker(s), model %s%s, listening on %s:%d" %
          (len(workers), a.model, " + " + a.embed_model if a.embed_model else "", a.bind, port), file=sys.stderr)
    stop = threading.Event()
    signal.signal(signal.SIGINT, lambda *_: stop.set())
    signal.signal(signal.SIGTERM, lambda *_: stop.set())
    try:
        if not a.no_tui and sys.stdout.isatty():
            Dashboard(d).run()                      # returns on q / Esc, like the reference
        else:
            stop.wait()
    finally:
        d.close()
        for e in encoders:
            e.close()
        for w in workers:
            w.close()
    return 0


if __name__ == "__main__":  # pragma: no cover
    sys.exit(main())
