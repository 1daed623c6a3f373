"""Helper for tests/test_multirank_cpu.py: one rank of the N>1 path on CPU (gloo), mock backends instead of GPUs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import ollamamq_b200 as mq  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"], rank=rank,
                            world_size=world)
    mine = bench.shard_users(world)[rank]
    # each rank drives its own dispatcher (one backend per rank), like one worker per GPU
    d = mq.Dispatcher(mock_backends=1, capacity=bench.USERS)
    streams = [d.submit("user%02d" % u, max_new_tokens=3) for u in mine]
    d.wait_parked()
    while d.mock_complete(0):
        pass
    d.drain(5000)
    ntok = sum(len(s.chunks) for s in streams)
    local_time = 1.0 + rank  # pretend rank r took (1 + r) seconds: the job time is the max over ranks
    t = torch.tensor([local_time], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(ntok)], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        print(json.dumps({"shards": gathered, "max_time": t.item(), "tokens": n.item(),
                          "log_len": len(d.log())}))
    d.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
