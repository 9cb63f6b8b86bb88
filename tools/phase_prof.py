"""Per-phase cycle breakdown of the fused ct x ct kernel (DPFHE_KS_PROF build)."""
import os, sys, json
os.environ["DPFHE_KS_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
log_n, L, B = 13, 4, 1184
c = dp.Context(log_n, L)
N = 1 << log_n
a = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda"); b = torch.empty_like(a); out = torch.empty_like(a)
evk = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
c.fill_uniform(1, a, 2 * B); c.fill_uniform(2, b, 2 * B); c.fill_uniform(3, evk, 2 * L)
c.ct_mul_relin(a, b, evk, out, B); c.phase_cycles()
c.ct_mul_relin(a, b, evk, out, B)
cyc = c.phase_cycles()
names = ["tensor+own key terms", "INTT register passes", "INTT outer stage + publish", "wait for sibling digit", "digit fetch + lift + outer fwd stage",
         "NTT register passes", "MAC with key column", "canon + store"]
items = B * L
tot = float(cyc.sum())
print(json.dumps({"work_items": items, "cycles_per_item": tot / items,
                  "phases": {n: {"cycles_per_item": round(float(cyc[k]) / items), "share": round(float(cyc[k]) / tot, 3)} for k, n in enumerate(names)}}, indent=1))
