import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashionern_aaai2024_b200 as ern
from fashionern_aaai2024_b200 import synthetic as syn
dev = torch.device("cuda", 0)
dim, rows = 640, int(sys.argv[1]) if len(sys.argv) > 1 else 32768
head = ern.CombinerSimple(dim, 4 * dim, 8 * dim); head.load_state_dict(syn.combiner_state(1, dim)); head = head.to(dev).eval()
a, b = torch.randn(rows, dim, device=dev), torch.randn(rows, dim, device=dev)
with torch.no_grad():
    for _ in range(3):
        head(a, b, want_bf16=True)
torch.cuda.synchronize()
