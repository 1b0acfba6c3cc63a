import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashionern_aaai2024_b200 as ern
from fashionern_aaai2024_b200 import synthetic as syn
dev = torch.device("cuda", 0)
dim, rows = 640, int(sys.argv[1]) if len(sys.argv) > 1 else 512
dvr = ern.DVR_module(dim); dvr.load_state_dict(syn.dvr_full_state(3, dim)); dvr = dvr.to(dev).eval()
pt, tk = torch.randn(rows, 13, dim, device=dev), torch.randn(rows, 77, dim, device=dev)
with torch.no_grad():
    for _ in range(2):
        dvr.encode(pt, tk)
torch.cuda.synchronize()
