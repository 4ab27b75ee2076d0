import os, sys
sys.path.insert(0, os.getcwd())
import torch
from parq_b200 import _lib, inputs as I
from parq_b200.raype import AddRayPEB200
dev = torch.device("cuda:0")
B, T, H, W = 16, 8, 60, 80
m = AddRayPEB200(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()
m.load_state_dict(I.make_raype_weights(0), strict=True)
m = m.to(dev)
feat = torch.randn(B, T, 1024, H, W, device=dev)
cam, Tcp, Twp, Twl = (t.to(dev) for t in I.make_geometry(B, T, H, W, seed=0))
for fn in (m.tokens, m.forward):
    fn(feat, cam, Tcp, Twp, Twl); torch.cuda.synchronize()
    _lib.profile_enable(list(_lib.PROFILE_TAGS), 64)
    fn(feat, cam, Tcp, Twp, Twl); torch.cuda.synchronize()
    print(fn.__name__, {k: v for k, v in _lib.profile_collect().items() if k in ("gemm", "rowwise")})
    _lib.profile_enable([], 0)
