import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from oai_analysis_2_b200 import ops
N, cin, dims = 2, 18, (80, 192, 192)
x = torch.randn(N, cin, *dims, device="cuda")
w = (torch.randn(cin, 27, 4, device="cuda") * 0.05).contiguous()
b = torch.randn(3, device="cuda")
out = torch.empty(N, 3, *dims, device="cuda")
for _ in range(2):
    ops.reg_conv3(x, cin, w, b, out, 3, 1, False, False, 0.1)
torch.cuda.synchronize()
