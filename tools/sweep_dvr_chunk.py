import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashionern_aaai2024_b200 as ern
from fashionern_aaai2024_b200 import synthetic as syn
dev = torch.device("cuda", 0)
dim, rows = 640, 4096
dvr = ern.DVR_module(dim); dvr.load_state_dict(syn.dvr_full_state(3, dim)); dvr = dvr.to(dev).eval()
pt, tk = torch.randn(rows, 13, dim, device=dev), torch.randn(rows, 77, dim, device=dev)
for mb in (256, 384, 512, 592, 768, 1024, 1184, 2048, 4096):
    dvr.max_batch = mb
    with torch.no_grad():
        for _ in range(2): dvr.encode(pt, tk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dvr.encode(pt, tk)
        e1.record(); torch.cuda.synchronize()
    print(mb, e0.elapsed_time(e1) / 5, flush=True)
