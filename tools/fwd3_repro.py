"""bitwise run-to-run check of the kw-stacked forward (conv_tc3) incl. its fused statistics: python tools/fwd3_repro.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import ops  # noqa: E402
from e2enet_medical_b200.plans import build_shiftconv_plan  # noqa: E402

dev = torch.device("cuda:0")
for src, cout, sp, B in (([1], 48, (32, 96, 96), 2), ([48], 48, (32, 96, 96), 2), ([48, 48], 48, (32, 96, 96), 2)):
    plan = build_shiftconv_plan(src, cout, (1, 1, 1))
    assert plan.fwd3 is not None
    rs = np.random.RandomState(0)
    xs8 = [ops.nc_to_c8(torch.from_numpy(rs.standard_normal((B, c) + sp).astype(np.float32)).to(dev)) for c in src]
    w = torch.from_numpy((rs.standard_normal((cout, sum(src), 1, 3, 3)) / 3).astype(np.float32)).to(dev)
    outs = []
    for it in range(12):
        raw = torch.full((B, cout // 8) + sp + (8,), float("nan"), dtype=torch.bfloat16, device=dev)
        st = ops.run_gemm_chunks([plan.fwd3], w, None, xs8, sp, sp, B, [raw], sp, [cout // 8], 1, want_stats=True)
        torch.cuda.synchronize()
        outs.append((raw.clone(), st.clone()))
    bad_raw = sum(int(not torch.equal(outs[0][0].view(torch.int16), o[0].view(torch.int16))) for o in outs[1:])
    bad_st = sum(int(not torch.equal(outs[0][1], o[1])) for o in outs[1:])
    nan = int(torch.isnan(outs[0][0].float()).sum())
    d = max(float((outs[0][1] - o[1]).abs().max()) for o in outs[1:])
    print(src, "->", cout, "runs differing: raw %d / 11, stats %d / 11 (max |d stats| %.3e), NaN left in raw: %d" % (bad_raw, bad_st, d, nan), flush=True)
