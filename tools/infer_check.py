"""2-rank sharded sliding-window check (run under torchrun on the GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/infer_check.py

predict_3D with set_tile_sharding(rank, world, None, "gather") must return on rank 0 the FULL (seg, softmax) of the
volume -- the reference's return contract -- equal to a single-process predict_3D of the same network; the other
rank returns (None, None) and uploads only its own x-planes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200.network_architecture.unetpp_d import softmax_helper  # noqa: E402
from e2enet_medical_b200.training import POOLS, build_network  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    patch = (64, 160, 160)
    torch.manual_seed(0)
    net = build_network(1, 16, POOLS["btcv"], patch, 48, deep_supervision=True).to(dev).eval()
    net.do_ds = False
    net.inference_apply_nonlin = softmax_helper
    vol = np.random.RandomState(0).randn(1, 150, 230, 239).astype(np.float32)      # 4 x 2 x 2 tiles, odd sizes
    args = (False, (0, 1, 2), True, 0.5, patch, None, True, "constant", None, True, False, True)
    ok = True
    net.set_tile_sharding(rank, world, None, "gather")
    seg, probs = net.predict_3D(vol, *args)
    h2d = net._last_h2d_bytes
    if rank == 0:
        net.set_tile_sharding(0, 1)
        seg1, probs1 = net.predict_3D(vol, *args)
        same = seg.shape == seg1.shape and probs.shape == probs1.shape and seg.dtype == np.int64
        err = float(np.abs(probs - probs1).max())
        agree = float((seg == seg1).mean())
        ok = same and err < 2e-5 and agree > 0.9999
        print("gather: shapes %s %s max|dprob| %.2e label agreement %.6f h2d bytes %d of %d -> %s"
              % (seg.shape, probs.shape, err, agree, h2d, vol.nbytes, "OK" if ok else "MISMATCH"), flush=True)
    else:
        ok = seg is None and probs is None and h2d < vol.nbytes
        print("rank %d: returned None, uploaded %d of %d bytes -> %s" % (rank, h2d, vol.nbytes, "OK" if ok else "BAD"), flush=True)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("INFER_CHECK", "PASS" if int(t) == 1 else "FAIL", flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
