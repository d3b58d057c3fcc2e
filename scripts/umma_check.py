import sys, numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from deep_reconstruction_with_epipolar_lines_mvster_b200 import ops
DEV='cuda'
for (cin,cout,kd,d,h,w) in [(32,32,1,1,4,128),(32,32,1,1,5,100),(64,64,1,1,9,33),(64,32,1,1,26,36),(32,32,3,4,10,12),(64,64,3,3,7,267),(64,64,3,8,2,2)]:
    rng = np.random.RandomState(7*cin+cout+kd+h)
    x = rng.normal(0,1,(2,cin,d,h,w)).astype(np.float32)
    wt = rng.normal(0,0.1,(kd,3,3,cin,cout)).astype(np.float32)
    bias = rng.normal(0,0.3,cout).astype(np.float32)
    ref = F.conv3d(torch.from_numpy(x).double(), torch.from_numpy(wt).double().permute(4,3,0,1,2), padding=(kd//2,1,1)) + torch.from_numpy(bias).double().view(1,-1,1,1,1)
    xd,wd,bd = torch.from_numpy(x).to(DEV), torch.from_numpy(wt).to(DEV), torch.from_numpy(bias).to(DEV)
    scale = max(1.0, ref.abs().max().item())
    out = {}
    for mode in (0,1,2):
        if mode == 0 and (h%2 or w%2): continue
        got = ops.conv3d_mid(xd,wd,bd,relu=False,tensor_cores=mode)
        torch.cuda.synchronize()
        out[mode] = (got.cpu().double()-ref).abs().max().item()/scale
    print((cin,cout,kd,d,h,w), {k: '%.2e'%v for k,v in out.items()}, flush=True)
