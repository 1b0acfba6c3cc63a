import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashionern_aaai2024_b200 as ern
from fashionern_aaai2024_b200 import synthetic as syn
dev = torch.device("cuda", 0)
dim, rows = 640, 32768
sr = ern.VisualSR(dim); sr.load_state_dict(syn.visualsr_state(2, dim)); sr = sr.to(dev).eval()
x = torch.randn(rows, 13, dim, device=dev)
with torch.no_grad():
    for _ in range(2):
        sr(x)
torch.cuda.synchronize()
