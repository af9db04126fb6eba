import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fuser_b200 import ops, _lib as L
dt = torch.bfloat16
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
which = sys.argv[3] if len(sys.argv) > 3 else "bwd"
x = torch.randn(rows, n, device="cuda").to(dt)
dy = torch.randn(rows, n, device="cuda").to(dt)
sm = torch.softmax(x.float(), -1).to(dt)
print("plan", ops.plan_for(L.OP_SOFTMAX_BWD if which == "bwd" else L.OP_SOFTMAX_FWD, dt, rows, n), flush=True)
t0 = time.time()
if which == "bwd":
    dx = ops.softmax_bwd(dy, sm)
    torch.cuda.synchronize()
    ref = (sm.float() * dy.float() - sm.float() * (sm.float() * dy.float()).sum(-1, keepdim=True))
    print("ok bwd", time.time() - t0, (dx.float() - ref).abs().max().item(), flush=True)
else:
    y = ops.softmax_fwd(x)
    torch.cuda.synchronize()
    print("ok fwd", time.time() - t0, (y.float() - sm.float()).abs().max().item(), flush=True)
