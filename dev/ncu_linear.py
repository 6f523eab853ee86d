"""Launch the tcgen05 linear kernel on the step's dominant shapes (for `ncu -k regex:linear_tc_kernel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dynamo-depth_b200"))
from dd_b200 import functional as Fn  # noqa: E402

lib = Fn.L.load()
P = Fn.L.ptr
for (M, K, N) in [(245760, 64, 384), (15360, 1344, 224)]:
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda")
    b = torch.randn(N, device="cuda")
    gy = torch.randn(M, N, device="cuda")
    y = torch.empty(M, N, device="cuda")
    gx = torch.empty(M, K, device="cuda")
    gw = torch.empty(N, K, device="cuda")
    gb = torch.empty(N, device="cuda")
    st = Fn._stream()
    ws = torch.empty(lib.dd_linear_workspace_bytes(M, K, N), dtype=torch.uint8, device="cuda")
    lib.dd_linear_fwd(P(x), P(w), P(b), M, K, N, P(y), P(ws), ws.numel(), st)
    lib.dd_linear_bwd(P(x), P(w), P(gy), M, K, N, P(gx), P(gw), P(gb), P(ws), ws.numel(), st)
    torch.cuda.synchronize()
