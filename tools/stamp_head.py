"""Phase timestamps of the fused small-batch fusion head (ERN_HEAD_STAMP_PTR profiling aid): CTA 0's %globaltimer at
start / after phase A / after barrier 1 / after phase B / after barrier 2 / end."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda", 0)
st = torch.zeros(8, dtype=torch.int64, device=dev)
os.environ["ERN_HEAD_STAMP_PTR"] = hex(st.data_ptr())
import fashionern_aaai2024_b200 as ern
from fashionern_aaai2024_b200 import synthetic as syn
dim = 640
head = ern.CombinerSimple(dim, 4 * dim, 8 * dim); head.load_state_dict(syn.combiner_state(1, dim)); head = head.to(dev).eval()
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
for rows in (1, 32, 64):
    a, b = torch.randn(rows, dim, device=dev), torch.randn(rows, dim, device=dev)
    acc = []
    with torch.no_grad():
        for it in range(6):
            flush.sum().item()
            head(a, b, want_bf16=True)
            torch.cuda.synchronize()
            t = st.cpu().tolist()
            acc.append([t[i + 1] - t[i] for i in range(5)])
    med = [sorted(x[i] for x in acc[1:])[len(acc[1:]) // 2] / 1e3 for i in range(5)]
    print(json.dumps({"rows": rows, "us_phaseA_barrier1_phaseB_barrier2_phaseC": med, "total_us": sum(med)}))
