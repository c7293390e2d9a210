import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instantir_b200 import ops
dev = "cuda"
n, HW, C = 2, 128 * 128, 320
g_, b_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
xs = [torch.randn(n * HW, C, device=dev).to(torch.bfloat16) for _ in range(8)]
out = torch.empty(n * HW, C, device=dev, dtype=torch.bfloat16)
for x in xs:
    ops.groupnorm(x, g_, b_, out, n_img=n, HW=HW, C=C, silu=True)
torch.cuda.synchronize()
print("ok")
