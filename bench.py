#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the FHE ciphertext-arithmetic hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port, all host threads)

One "step" = one pass of ct x ct multiply + relinearise (dpfhe_ct_mul_relin) over one batch of
synthetic ciphertexts at BASELINE.json config 2: N = 8192, L = 4, batch = 4096 per GPU.
  value      whole-job ct-mults/s, inputs resident in HBM, timed with CUDA events on the launch stream
  e2e        the same metric through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside)
  roofline   the fused key-switch kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  ntt        NTTs/s of the standalone forward transform on the same data, with its own roofline
  cpu_baseline  the CPU oracle (kind "port": the reference has no CPU evaluator) on a bounded sample
Multi-GPU: one process per GPU (torchrun), ciphertexts sharded by rank, no collective on the data path
(weak scaling); the final NCCL gather to rank 0 is timed separately and reported under "gather".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N, L, BATCH = 13, 4, 4096
N = 1 << LOG_N
P_WORDS = L * N
CT_BYTES = 2 * P_WORDS * 8
ALGO_BYTES_CT_MUL = 6 * P_WORDS * 8      # read 2 cts, write 1 ct   
ALGO_BYTES_NTT = 2 * N * 8               # read + write one limb
SEED = 0xD3390002                        # 0xD3390000 + config id 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dpfhe", choices=["dpfhe", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="ciphertexts per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational hybrid key-switching timing")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(kernel, units):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `units` work units, scaled from the committed
    ncu --set full capture (profiles/traffic.json holds bytes per unit at the profiled batch); None if absent"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f).get(kernel)
        return float(rec["dram_bytes_per_unit"]) * units if rec else None
    except Exception:
        return None


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def pick_threads(o):
    """thread count for the CPU arm: the fastest of {all processors, half of them (one per core on SMT hosts), OpenMP's own
    default} on a short probe — more threads are not always faster (measured on this pool's host: 128 threads 2.8 k, 64
    threads 6.6 k ct-mult/s), and launchers such as torchrun cap OMP_NUM_THREADS at 1"""
    procs = o.host_threads()
    cands = sorted({procs, max(1, procs // 2), max(1, o.max_threads())}, reverse=True)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    n = min(512, 8 * max(cands))     # several ciphertexts per thread: a probe of one each mostly times the thread start-up
    a = o.fill_uniform(SEED, 2 * n).reshape(n, 2, L, N)
    b = o.fill_uniform(SEED, 2 * n, first_poly=2 * n).reshape(n, 2, L, N)
    _, out = o.time_ct_mul_relin(a, b, evk, cands[0])
    best, best_t = cands[0], None
    for c in cands:
        t = min(o.time_ct_mul_relin(a, b, evk, c, out=out)[0] for _ in range(3))
        if best_t is None or t < best_t:
            best, best_t = c, t
    return best


def cpu_sample(oracle_ctx, seconds, threads):
    """times the oracle's ct_mul_relin on a bounded sample for about `seconds` of wall time;
    returns (ct-mults/s, ct-mults timed, seconds).  The sample is at most 1024 ciphertexts (1.5 GiB of host
    arrays) and is repeated until the time budget is used, so many-core hosts still get a 10-30 s measurement."""
    o = oracle_ctx
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    probe = max(threads, 1)
    a = o.fill_uniform(SEED, 2 * probe).reshape(probe, 2, L, N)
    b = o.fill_uniform(SEED, 2 * probe, first_poly=2 * probe).reshape(probe, 2, L, N)
    t, _ = o.time_ct_mul_relin(a, b, evk, threads)
    rate = probe / t
    n = int(max(probe, min(1024, rate * seconds)))
    n = max(threads, (n // max(threads, 1)) * max(threads, 1))
    a = o.fill_uniform(SEED, 2 * n).reshape(n, 2, L, N)
    b = o.fill_uniform(SEED, 2 * n, first_poly=2 * n).reshape(n, 2, L, N)
    _, out = o.time_ct_mul_relin(a, b, evk, threads)     # warm-up (page faults of the reused output buffer, thread pool)
    done, total = 0, 0.0
    while total < seconds and done < 64 * n:
        t, _ = o.time_ct_mul_relin(a, b, evk, threads, out=out)
        total += t
        done += n
    return done / total, done, total


def run_reference(args):
    """CPU arm: the oracle port on all host threads (the reference has no implementation to run)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    o = oracle.Oracle(LOG_N, L)
    threads = pick_threads(o)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    # bounded sample per step: about 2 s of CPU work, at most 1024 ciphertexts
    rate, _, _ = cpu_sample(o, 1.0, threads)
    n = int(max(threads, min(1024, args.batch, rate * 2.0)))
    a = o.fill_uniform(SEED, 2 * n).reshape(n, 2, L, N)
    b = o.fill_uniform(SEED, 2 * n, first_poly=2 * n).reshape(n, 2, L, N)
    _, out = o.time_ct_mul_relin(a, b, evk, threads)     # the output buffer is touched once and reused: no page faults in the timed steps
    for _ in range(args.warmup):
        o.time_ct_mul_relin(a, b, evk, threads, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.time_ct_mul_relin(a, b, evk, threads, out=out)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt
    line = {
        "impl": "reference", "metric": "ct_mult_relin_per_s", "value": value, "unit": "ct-mult/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "ct x ct multiply + relinearise, N=8192, L=4 (BASELINE.json config 2)",
                   "sample": "%d ciphertexts per step on the host CPU" % n},
        "cpu_baseline": {"value": value, "unit": "ct-mult/s", "cores": threads, "kind": "port",
                         "sample": "%d ct-mults per step x %d steps, oracle/dpfhe_oracle.c with OpenMP" % (n, args.steps)},
        "e2e": {"value": value, "unit": "ct-mult/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import deeppowers_b200 as dp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL's banner ("NCCL version ...") otherwise lands there
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch
    ctx = dp.Context(LOG_N, L, device=local_rank)
    shape = (B, 2, L, N)
    a = torch.empty(shape, dtype=torch.int64, device="cuda")
    b = torch.empty(shape, dtype=torch.int64, device="cuda")
    out = torch.empty(shape, dtype=torch.int64, device="cuda")
    evk = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
    first = rank * 2 * B          # each rank owns a disjoint slice of the global synthetic stream
    ctx.fill_uniform(SEED, a, 2 * B, first_poly=first)
    ctx.fill_uniform(SEED + 1, b, 2 * B, first_poly=first)
    ctx.fill_uniform(SEED + 2, evk, 2 * L)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ctx.launch_count()
    total_ms = timed(lambda: ctx.ct_mul_relin(a, b, evk, out, B), args.steps, args.warmup)
    # launches inside the timed region only: key_prepare_kernel + ks_fused_kernel per step
    launches = (ctx.launch_count() - n0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    peak, peak_src = peaks()
    kern_gbs = B * ALGO_BYTES_CT_MUL / (ms_per_step * 1e-3) / 1e9      # per GPU: one launch per step
    roofline = {"bound": "hbm", "kernel": "ks_fused_kernel<13,256,3,MUL_RELIN> (persistent cooperative, 3 CTAs/SM)", "achieved": kern_gbs, "peak": peak,
                "unit": "GB/s", "frac": kern_gbs / peak, "traffic": traffic_for("ks_fused_kernel_mul_relin", B),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": B * ALGO_BYTES_CT_MUL,
                "note": "64-bit modular integer work: integer issue (IMAD 2.0, IMAD.WIDE 2.55, IADD3 1.5 clk per warp-instruction per SM sub-partition, no ALU/IMAD overlap: profiles/r01/int_pipes*.txt) bounds this kernel below the HBM roofline; ncu: issue slots 56% busy (DESIGN.md section 6)"}

    # Secondary, explanatory roofline: issued warp-instructions per second against what the integer pipes sustain for
    # this instruction mix (profiles/r01/int_pipes_final.txt: IMAD.WIDE 2.55, IMAD 2.0, ALU ~1.4 clk per warp-instruction
    # per SM sub-partition, not overlapping; mix and instructions per ciphertext from profiles/r01/ncu_full_summary.json).
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            instr_per_ct = float(json.load(f)["ks_fused_kernel_mul_relin"]["warp_instructions_per_unit"])
        mix_clk = 0.28 * 2.55 + 0.24 * 2.0 + 0.40 * 1.4 + 0.08 * 1.0          # clocks per warp-instruction per sub-partition
        sm_clock = 1.965e9
        int_peak = torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * sm_clock / mix_clk
        achieved_int = (B / (ms_per_step * 1e-3)) * instr_per_ct
        roofline["int_issue"] = {"achieved_warp_instr_per_s": achieved_int, "peak_warp_instr_per_s": int_peak,
                                 "frac": achieved_int / int_peak, "warp_instr_per_ct_mult": instr_per_ct,
                                 "source": "ncu smsp__inst_executed.sum per launch / batch; pipe costs from the committed microbenchmark"}
    except Exception:
        pass

    # standalone NTT on the same data (NTTs/s half of the BASELINE metric)
    ntt_ms = timed(lambda: ctx.ntt_fwd(a, 2 * B), max(3, args.steps // 2), 2) / max(3, args.steps // 2)
    n_ntt = 2 * B * L
    ntt_gbs = n_ntt * ALGO_BYTES_NTT / (ntt_ms * 1e-3) / 1e9
    ntt = {"metric": "ntt_fwd_per_s", "value": world * n_ntt / (ntt_ms * 1e-3), "unit": "NTT/s", "ms_per_step": ntt_ms,
           "roofline": {"bound": "hbm", "kernel": "ntt_kernel<13,256,3,fwd> (one CTA per limb, 3 CTAs/SM)", "achieved": ntt_gbs, "peak": peak, "unit": "GB/s",
                        "frac": ntt_gbs / peak, "traffic": traffic_for("ntt_kernel_fwd", n_ntt)}}
    ctx.fill_uniform(SEED, a, 2 * B, first_poly=first)     # restore `a` (the NTT ran in place)

    # SURVEY.md section 8 row f-2 (informational, not the headline): the same ciphertexts through special-prime hybrid
    # key switching (context = the four ciphertext moduli + one special prime; 30 transforms per ct-mult instead of 16)
    extras = None
    if not args.no_extras:
        ctx5 = dp.Context(LOG_N, L + 1, device=local_rank)
        hkey = torch.empty((L, 2, L + 1, N), dtype=torch.int64, device="cuda")
        ctx5.fill_uniform(SEED + 3, hkey, 2 * L)
        k = max(3, args.steps // 3)
        hyb_ms = timed(lambda: ctx5.ct_mul_relin_hybrid(a, b, hkey, out, B, 65537), k, 2) / k
        extras = {"ct_mul_relin_hybrid": {"value": world * B / (hyb_ms * 1e-3), "unit": "ct-mult/s", "ms_per_step": hyb_ms,
                                          "kernel": "ks_hybrid_kernel<13,256,3,MUL_RELIN>", "GBps": B * ALGO_BYTES_CT_MUL / (hyb_ms * 1e-3) / 1e9,
                                          "note": "4 ciphertext limbs + 1 special prime, BGV rounding t=65537 (DESIGN.md 2.10)"}}
        ctx5.close()
        del hkey
        ctx.ct_mul_relin(a, b, evk, out, B)     # `out` is compared with the end-to-end result below

    # end to end through the host-buffer ABI: pinned host memory, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        ha = torch.empty(shape, dtype=torch.int64, pin_memory=True)
        hb = torch.empty(shape, dtype=torch.int64, pin_memory=True)
        ho = torch.empty(shape, dtype=torch.int64, pin_memory=True)
        hk = torch.empty((L, 2, L, N), dtype=torch.int64, pin_memory=True)
        ha.copy_(a); hb.copy_(b); hk.copy_(evk)
        torch.cuda.synchronize()
        na, nb, no, nk = (t.numpy().view(np.uint64) for t in (ha, hb, ho, hk))
        ctx.ct_mul_relin_host(na, nb, nk, no)       # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ctx.ct_mul_relin_host(na, nb, nk, no)
        dt = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_s = float(dt.item())
        ok = bool(torch.equal(ho.cuda(), out))     # the host path must reproduce the device path bit for bit
        e2e = {"value": world * B / e2e_s, "unit": "ct-mult/s", "h2d_bytes_per_step": 2 * B * CT_BYTES + 2 * L * P_WORDS * 8,
               "d2h_bytes_per_step": B * CT_BYTES, "ms_per_step": e2e_s * 1e3, "matches_device_path": ok,
               "pcie": {"h2d_GBps_per_gpu": (2 * B * CT_BYTES + 2 * L * P_WORDS * 8) / e2e_s / 1e9, "d2h_GBps_per_gpu": B * CT_BYTES / e2e_s / 1e9,
                        "note": "both directions run concurrently; the host-to-device stream (two operand batches per result batch) is the bound"},
               "api": "dpfhe_ct_mul_relin_host (pinned host buffers, 3-stage H2D/compute/D2H pipeline)"}
        del ha, hb, ho, hk

    # final result gather (the only collective): NCCL gather of every rank's output to rank 0
    gather = None
    if world > 1:
        bufs = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
        warm = [torch.empty_like(out[:8]) for _ in range(world)] if rank == 0 else None
        dist.gather(out[:8].contiguous(), warm, dst=0)      # NCCL channel set-up happens on the first call
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.gather(out, bufs, dst=0)
        g1.record()
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        gather = {"ms": float(gms.item()), "bytes_into_root": (world - 1) * B * CT_BYTES,
                  "GBps_into_root": (world - 1) * B * CT_BYTES / (float(gms.item()) * 1e-3) / 1e9,
                  "value_with_gather": world * B / ((ms_per_step + float(gms.item())) * 1e-3), "collective": "ncclGather via torch.distributed"}
        del bufs

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            import oracle
            oracle.build()
            o = oracle.Oracle(LOG_N, L)
            threads = pick_threads(o)
            rate, n, t = cpu_sample(o, args.cpu_seconds, threads)
            cpu = {"value": rate, "unit": "ct-mult/s", "cores": threads, "kind": "port",
                   "sample": "%d ct-mults in %.1f s, oracle/dpfhe_oracle.c with OpenMP (the reference has no CPU evaluator)" % (n, t)}
        line = {
            "metric": "ct_mult_relin_per_s", "value": value, "unit": "ct-mult/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "ct x ct multiply + relinearise, N=8192, L=4, batch=%d per GPU (BASELINE.json config 2)" % B,
                       "global_batch": world * B, "parallelism": "batch-sharded x%d, no data-path collective" % world,
                       "l2": "inputs+outputs are %.1f GiB per GPU (>> 126 MB L2), no flush needed" % (3 * B * CT_BYTES / 2**30),
                       "seed": hex(SEED)},
            "roofline": roofline, "ntt": ntt, "extras": extras, "cpu_baseline": cpu, "e2e": e2e, "gather": gather,
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
