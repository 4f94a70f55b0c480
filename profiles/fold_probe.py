"""Runs the ERB fold (forward + backward through the reference-shaped API) of one block a few times: the command the
ncu launch list of the fold kernels is taken on.  usage: python profiles/fold_probe.py [cin cout [reps]]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from orepnerv.model import NeRVBlock  # noqa: E402

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 96
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = NeRVBlock(ngf=cin, new_ngf=cout, stride=1, bias=True, norm='none', act='swish', deploy=False, conv_type='conv',
                branch_type='ERB').to(dev)
dK = torch.randn(cout, cin, 3, 3, device=dev)
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K, b = blk.get_equivalent_kernel_bias()
    (K * dK).sum().backward()
    torch.cuda.synchronize()
    print(f"rep {r}: fold fwd+bwd {1e3 * (time.perf_counter() - t0):.3f} ms (host clock, includes ATen glue)")
