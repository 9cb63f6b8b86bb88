"""Pinned-memory PCIe rates on this host: H2D alone, D2H alone, both at once (the ceiling of the end-to-end number)."""
import json, time, torch
n = 1 << 30   # 1 GiB per transfer
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


run(True, True, 1)
print(json.dumps({"h2d_alone_GBps": run(True, False), "d2h_alone_GBps": run(False, True), "each_direction_when_concurrent_GBps": run(True, True)}))
