"""Run one conv layer a few times (for ncu captures).  args: n H Cin Cout k [res] [pre]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from human_dynamics_b200.nets import PackedConv
n, H, Cin, Cout, k = [int(a) for a in sys.argv[1:6]]
res = len(sys.argv) > 6 and sys.argv[6] == '1'
pre = len(sys.argv) > 7 and sys.argv[7] == '1'
impl = os.environ.get('HD_IMPL', 'tc3h')
dev = torch.device('cuda')
rng = np.random.RandomState(0)
w = (rng.normal(0, 1, size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
pc = PackedConv(w, dev, np.ones(Cout, np.float32), np.zeros(Cout, np.float32), True, pad=(k // 2, k // 2), tc=impl)
x = torch.randn((n, H, H, Cin), device=dev)
out = torch.empty((n, H, H, Cout), device=dev)
r = torch.randn((n, H, H, Cout), device=dev) if res else None
pr = (torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev), 0, 1) if pre else None
if os.environ.get('HD_SPLIT', '0') == '1':
    xs = (x.half(), ((x - x.half().float()) * 2048).half())
    osp = (torch.empty((n, H, H, Cout), dtype=torch.float16, device=dev), torch.empty((n, H, H, Cout), dtype=torch.float16, device=dev))
    op = pc.bind(None, n, H, H, out if res else None, inp_split=xs, out_split=osp, res=r, impl='tc3h')
else:
    op = pc.bind(x, n, H, H, out, pre=pr, res=r, impl=impl)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    op.run(st)
torch.cuda.synchronize()
