"""2-rank data-parallel check (run under torchrun on the GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Both ranks train on the SAME batch, so the mean of their gradients equals each rank's own gradient and the
bucketed, overlapped all-reduce of training.TrainStep must reproduce a single-process run: losses and weights after
3 steps are compared with a world-size-1 TrainStep inside rank 0, eager and as a CUDA graph.  Also times 10 steps."""
import os
import random
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pools, patch = POOLS["btcv"], (32, 96, 96)
    data, targets = synthetic_batch(2, 1, 14, patch, pools, seed=1)
    x, tg = data.to(dev), [t.to(dev) for t in targets]

    def run(world_size, graph):
        random.seed(0)
        ts = TrainStep(1, 14, pools, patch, 0.2, 0.5, 1200, dev, world_size, seed=0, n_buckets=4)
        if graph:
            ts.enable_graph(x, tg, warmup=2)
        losses = [float(ts.step(x, tg)) for _ in range(3)]
        w = {k: v.detach().clone() for k, v in ts.network.state_dict().items()}
        return losses, w, ts

    ok = True
    for graph in (False, True):
        l_dp, w_dp, ts = run(world, graph)
        if rank == 0:
            l_1, w_1, _ = run(1, graph)
            worst = max(float((w_dp[k] - w_1[k]).norm() / w_1[k].norm().clamp_min(1e-12)) for k in w_1
                        if not k.endswith("conv.bias"))
            good = all(abs(a - b) < 5e-3 * abs(b) for a, b in zip(l_dp, l_1)) and worst < 0.2
            print("graph=%s dp losses %s single %s worst weight rel diff %.3e buckets %d -> %s"
                  % (graph, l_dp, l_1, worst, ts.arena.n_buckets, "OK" if good else "MISMATCH"), flush=True)
            ok = ok and good
        dist.barrier()
        if graph:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ts.step(x, tg)
            e1.record()
            torch.cuda.synchronize()
            if rank == 0:
                print("graphed dp step on the small patch: %.2f ms" % (e0.elapsed_time(e1) / 10), flush=True)
        ts._graph = None
        torch.cuda.synchronize()
        dist.barrier()
    if rank == 0:
        print("DP_CHECK", "PASS" if ok else "FAIL", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
