#!/usr/bin/env python
"""bench.py -- unique-reads/s through the dada() core (BASELINE.json metric) on N B200s.

A "step" is one dada_uniques()-equivalent pass over one batch of synthetic dereplicated
reads (BASELINE.json configs[1]: 1e5 synthetic 250 nt uniques, Zipf abundances, 100 true
variants, Illumina-like qualities, tperr1 error matrix, selfConsist=FALSE, default options).

  value  : uniques/s with the packed uniques already resident in HBM (Resident.run)
  e2e    : uniques/s through the one-shot C-ABI call dada2b_run() on HOST buffers
           (pack + H2D + loop + D2H of every output inside the timed region)
  N > 1  : one process per GPU (torchrun); each rank denoises its own sample -- the
           reference's per-sample loop (R/dada.R:266) -- no data-path collective, "weak".
  --impl reference : the reference's own C++ (oracle/_ref, compiled unmodified from
           /root/reference/src in the build container) on the host cores, same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUNIQ_DEFAULT = 100000


def workload(n_uniques, seed):
    from tools import synth
    from tests import cases
    seqs, ab, q, truth = synth.illumina(n_uniques, seed=seed)
    return seqs, ab, q, cases.tperr1()


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed region through NVML (in-process: a polling
    `nvidia-smi -lms` child was observed to stall the stream synchronisations of this latency-bound loop
    for hundreds of ms at a time).  Falls back to one nvidia-smi query per second."""

    def __init__(self, index, period=0.5):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.rows = []          # (sm_mhz, max_mhz, reasons set)
        self.stop_flag = False

    def run(self):
        if os.environ.get("DADA2B_BENCH_NOCLOCKS"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                except Exception:
                    r = 0
                self.rows.append((float(sm), float(mx), {k for k, bit in names.items() if r & bit}))
                time.sleep(self.period)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                reasons = {n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6])
                           if v.strip().lower().startswith("active")}
                self.rows.append((float(out[0]), float(out[1]), reasons))
            except Exception:
                pass
            time.sleep(1.0)

    def finish(self):
        self.stop_flag = True
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for r in rows:
            reasons |= r[2]
        return {"sm_mhz": float(np.median([r[0] for r in rows])), "sm_max_mhz": float(max(r[1] for r in rows)),
                "reasons": sorted(reasons), "samples": len(rows)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


AB_VARIANTS = [("nwfwd2", {"DADA2B_NWFWD_V2": "1"}),
               ("small16x4", {"DADA2B_NWFWD_SMALL": "1"}),
               ("nwfwd2_small16x4", {"DADA2B_NWFWD_V2": "1", "DADA2B_NWFWD_SMALL": "1"}),
               ("fused_tail", {"DADA2B_FUSED_TAIL": "1"}),
               ("pivot", {"DADA2B_PIVOT": "1"}),
               ("twophase", {"DADA2B_TWOPHASE": "1"}),
               ("nwfwd2_twophase_bound16", {"DADA2B_NWFWD_V2": "1", "DADA2B_TWOPHASE": "1", "DADA2B_BOUND16": "1"}),
               ("all", {"DADA2B_NWFWD_V2": "1", "DADA2B_FUSED_TAIL": "1", "DADA2B_PIVOT": "1"}),
               ("everything", {"DADA2B_NWFWD_V2": "1", "DADA2B_FUSED_TAIL": "1", "DADA2B_PIVOT": "1", "DADA2B_TWOPHASE": "1", "DADA2B_BOUND16": "1"}),
               ("everything_small16x4", {"DADA2B_NWFWD_V2": "1", "DADA2B_FUSED_TAIL": "1", "DADA2B_PIVOT": "1", "DADA2B_TWOPHASE": "1", "DADA2B_BOUND16": "1",
                                         "DADA2B_NWFWD_SMALL": "1"})]


def _last_lines(txt, n=2, width=300):
    """The end of a failed leg's output: the exception line, not the middle of its traceback."""
    rows = [l.strip() for l in txt.strip().splitlines() if l.strip()]
    return " | ".join(rows[-n:])[-width:]


def experimental_ab(seqs, ab, q, err, last, budget_s, device, leg_cmd=None, steps=3, warmup=2, variants=None):
    """After the measured region (N=1 only): every off-by-default kernel variant runs the same workload
    in its own subprocess under a timeout, and its outputs are diffed against the default path's.  Reported under
    "experimental_ab"; never part of `value`/`e2e`.  Bounded by `budget_s` seconds in total."""
    import tempfile
    from tests import cases
    t_start = time.time()
    res = {}
    with tempfile.TemporaryDirectory() as td:
        wl, ref = os.path.join(td, "wl.npz"), os.path.join(td, "ref.npz")
        np.savez(wl, seqs=np.array(seqs), ab=np.asarray(ab), q=np.asarray(q), err=np.asarray(err))
        np.savez(ref, **cases.flatten(last))
        for tag, env in (variants or AB_VARIANTS):
            left = budget_s - (time.time() - t_start)
            if left < 20:
                res[tag] = {"skipped": "A/B time budget spent"}
                continue
            e = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(device)), **env)
            try:
                out = subprocess.run((leg_cmd or [sys.executable, os.path.join(ROOT, "tools", "ab_leg.py")]) + [wl, ref, str(steps), str(warmup)], env=e,
                                     capture_output=True, text=True, timeout=min(45, left))
                rows = [l for l in out.stdout.splitlines() if l.startswith("ABLEG ")]
                res[tag] = json.loads(rows[-1][6:]) if rows else {"failed": _last_lines(out.stderr or out.stdout)}
            except subprocess.TimeoutExpired:
                res[tag] = {"failed": "timeout"}
            except Exception as ex:                      # never let the A/B leg take the bench line down
                res[tag] = {"failed": repr(ex)[:200]}
            res[tag]["switches"] = sorted(env)
    return res


def bimera_leg(budget_s, device, cmd=None):
    """After the measured region (N=1 only): SURVEY.md 8(f3)'s bimera detection timed through its C-ABI on a synthetic
    sequence table, next to the reference's C_table_bimera2 on the host cores, in a subprocess under a timeout (the
    kernels are new: first run on hardware).  Reported under "bimera"; never part of `value`/`e2e`."""
    e = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(device)))
    try:
        out = subprocess.run(cmd or [sys.executable, os.path.join(ROOT, "tools", "bimera_leg.py")], env=e, capture_output=True, text=True,
                             timeout=budget_s)
        rows = [l for l in out.stdout.splitlines() if l.startswith("BIMLEG ")]
        return json.loads(rows[-1][7:]) if rows else {"failed": _last_lines(out.stderr or out.stdout)}
    except subprocess.TimeoutExpired:
        return {"failed": "timeout"}
    except Exception as ex:
        return {"failed": repr(ex)[:200]}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation, all host threads, same workload."""
    if rank != 0:
        return
    from oracle import ref
    ncores = os.cpu_count() or 1
    ref.set_threads(ncores)
    # same workload as the B200 arm (N x nuniques uniques in one sample when sharded), bounded to 2 x nuniques so that
    # warmup + steps passes of the CPU implementation end within a few minutes
    total = args.nuniques * (world if args.mode == "shard" else 1)
    n_s = min(total, 2 * args.nuniques)
    seqs, ab, q, err = workload(n_s, 12345)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ref.dada_uniques(seqs, ab, None, err, q, multithread=True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    tsum = sum(times)
    val = n_s * len(times) / tsum
    line = {"impl": "reference", "metric": "unique-reads/sec through dada()", "value": val, "unit": "uniques/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tsum / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32 + f64",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: %d synthetic 250 nt uniques per GPU (%d in total), 100 variants (Zipf), Illumina-like quals, "
                                   "tperr1, dada() selfConsist=FALSE, default options" % (args.nuniques, total)},
            "cpu_baseline": {"value": val, "unit": "uniques/s", "cores": ncores, "kind": "reference",
                             "sample": "%d-unique sample (%s) per step, multithread=TRUE on %d threads (parallelFor shim over std::thread)" % (n_s, "the full workload" if n_s == total else "bounded from %d" % total, ncores)},
            "e2e": {"value": val, "unit": "uniques/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nuniques", type=int, default=NUNIQ_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="shard", choices=["shard", "replicas"],
                    help="N>1: shard ONE sample of N x nuniques uniques over the ranks (NCCL all-gather per split round; weak scaling) "
                         "or run one independent sample per rank (no collective)")
    ap.add_argument("--ab-seconds", type=int, default=150,
                    help="N=1: total wall budget for the post-measurement A/B of the experimental kernel variants (0 = off)")
    ap.add_argument("--bimera-seconds", type=int, default=60,
                    help="N=1: timeout of the post-measurement bimera-detection leg (0 = off)")
    ap.add_argument("--watchdog", type=int, default=1500, help="dump stacks and exit after this many seconds")
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog, exit=True)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import dada2_b200
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shard = world > 1 and args.mode == "shard"
    if shard:
        seqs, ab, q, err = workload(args.nuniques * world, 12345)       # same sample on every rank
    else:
        seqs, ab, q, err = workload(args.nuniques, 12345 + rank)
    nraw = len(seqs)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # ---------------- value: inputs resident in HBM ----------------
    if shard:
        from dada2_b200 import multi
        res = multi.sharded_resident(seqs, ab, None, q, device=local_rank)
    else:
        res = dada2_b200.Resident(seqs, ab, None, q, device=local_rank)
    last = None
    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi's start-up takes ~1 s of driver calls: keep it out of the timed region
    for _ in range(args.warmup):
        last = res.run(err)
        flush.zero_()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < (0 if os.environ.get("DADA2B_BENCH_NOCLOCKS") else 10):
        time.sleep(0.05)
    sampler.rows.clear()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    step_ms, step_host = [], []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        ts = time.perf_counter()
        last = res.run(err)
        step_ms.append(round((time.perf_counter() - ts) * 1e3, 2))
        step_host.append([round(last["stats"][k], 2) for k in ("ms_setup", "ms_loop", "ms_final")])
        dev_ms += last["stats"]["ms_device"]
    barrier()
    t_val = time.perf_counter() - t0
    st = last["stats"]

    # ---------------- e2e: host buffers in, host buffers out, through the C-ABI ----------------
    e2e_ms = []
    est = None
    if shard:
        from dada2_b200 import _abi
        pin = _abi.PackedIn(seqs, ab, None, None, q)
        ecm = np.asfortranarray(np.asarray(err, dtype=np.float64))
        ostruct = _abi.make_opts(homo_gap=-8)
        for _ in range(max(1, args.warmup // 2)):
            res.reupload(pin); res.run_raw(ecm, ecm.shape[1], ostruct)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            ts = time.perf_counter()
            res.reupload(pin)                                   # dada2b_reupload: pack + H2D of this rank's copy
            est = res.run_raw(ecm, ecm.shape[1], ostruct)       # dada2b_run_resident: loop + D2H of every output
            e2e_ms.append(round((time.perf_counter() - ts) * 1e3, 2))
        barrier()
        t_e2e = time.perf_counter() - t0
        L0 = len(seqs[0])
        est = dict(est)
        est["h2d_bytes"] += nraw * ((((L0 + 15) // 16 + 3) & ~3) * 4 + ((L0 + 15) & ~15) + 7)   # packed upload of dada2b_reupload
    else:
        call = dada2_b200.PackedCall(seqs, ab, None, err, q)
        for _ in range(max(1, args.warmup // 2)):
            call.run(unpack=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            r, _ms = call.run(unpack=False)
            e2e_ms.append(round(_ms, 2))
            est = r["stats"]
        barrier()
        t_e2e = time.perf_counter() - t0
    clocks = sampler.finish()

    tt = torch.tensor([t_val, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_val, t_e2e = float(tt[0]), float(tt[1])
    units = nraw if shard else world * nraw          # uniques denoised per step by the whole job
    value = units * args.steps / t_val
    e2e = units * args.steps / t_e2e

    if rank == 0:
        L = len(seqs[0])
        bytes_per_pair = (L + 3) // 4 + 16 + L                      # S + O + Q for an aligned pair
        pairs = st["n_nw"]                                         # loop alignments (k_nwfwd + its fallback)
        k_ms = st["ms_k_align_nw"]
        n_launch = st["n_rounds"] + 1                              # one k_nwfwd launch per compare round
        peak, peak_src = measured_peak_gbs()
        achieved = (pairs * bytes_per_pair / 1e9) / (k_ms / 1e3) if k_ms > 0 else 0.0
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r1_final_k_nwfwd_metrics.json")))
        except Exception:
            pass
        traffic = prof.get("dram_bytes_per_pair")
        roofline = {"bound": "hbm", "kernel": "k_nwfwd (loop banded NW with forward-carried lambda; CUDA-event sum over its launches)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (traffic * pairs / n_launch) if traffic else None,
                    "traffic_note": "ncu dram read+write bytes per aligned pair x average pairs per launch (profiles/r1_final_k_nwfwd_full.txt)",
                    "peak_source": peak_src, "algorithmic_bytes_per_pair": bytes_per_pair,
                    "algorithmic_bytes_per_launch": bytes_per_pair * pairs / n_launch, "pairs_per_step": int(pairs),
                    "kernel_ms_per_step": k_ms, "launches_per_step": int(n_launch), "avg_launch_ms": k_ms / max(1, n_launch),
                    "nw_gcups": (pairs * (L * 33 - 16 * 17) / 1e9) / (k_ms / 1e3) if k_ms > 0 else 0.0,
                    "kernel_share_of_device_time": k_ms / st["ms_device"] if st["ms_device"] else None,
                    "issue_bound_evidence": prof.get("large_round"),
                    "note": "integer DP is ALU-issue bound, not HBM-bound (DESIGN.md 4.2): the HBM fraction is low by construction"}
        cpu = None
        parity = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref
            from tests import cases
            ncores = os.cpu_count() or 1
            kind = "reference" if ref.available() else "port"
            n_s = args.nuniques if ncores >= 8 else min(args.nuniques, 25000)
            if n_s == args.nuniques:
                s_seqs, s_ab, s_q = seqs, ab, q
            else:
                s_seqs, s_ab, s_q, _ = workload(n_s, 12345)
            if kind == "reference":
                ref.set_threads(ncores)
                t0 = time.perf_counter()
                cres = ref.dada_uniques(s_seqs, s_ab, None, err, s_q, multithread=True)
                dt = time.perf_counter() - t0
            else:
                from oracle import port
                ncores = 1
                t0 = time.perf_counter()
                cres = port.dada_uniques(s_seqs, s_ab, None, err, s_q)
                dt = time.perf_counter() - t0
            cpu = {"value": n_s / dt, "unit": "uniques/s", "cores": ncores, "kind": kind,
                   "sample": "%d-unique %s of the step workload, one pass, %.1f s" % (n_s, "= all" if n_s == nraw else "subsample", dt)}
            if n_s == nraw:
                try:
                    cases.assert_same(last, cres, rtol=1e-10, label="bench")
                    parity = "outputs identical to the CPU reference on this workload (ints exact, fp64 <= 1e-10)"
                except AssertionError as e:
                    parity = "MISMATCH: %s" % e
        ab_res = None
        clean_env = not any(k.startswith("DADA2B_") and k not in ("DADA2B_VERBOSE", "DADA2B_AB_TAG", "DADA2B_BENCH_NOCLOCKS") for k in os.environ)
        if world == 1 and args.ab_seconds > 0 and clean_env:
            try:
                ab_res = experimental_ab(seqs, ab, q, err, last, args.ab_seconds, local_rank)
            except Exception as ex:
                ab_res = {"failed": repr(ex)[:200]}
        bim_res = None
        if world == 1 and args.bimera_seconds > 0 and clean_env:
            bim_res = bimera_leg(args.bimera_seconds, local_rank)
        line = {"metric": "unique-reads/sec through dada()", "value": value, "unit": "uniques/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_val / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 DP + f64 lambda/p-value",
                "data": "synthetic",
                "config": {"workload": "BASELINE configs[1]: %d synthetic 250 nt uniques per GPU (%d in total), 100 variants (Zipf), Illumina-like quals, "
                                       "tperr1, dada() selfConsist=FALSE, default options" % (args.nuniques, units),
                           "per_gpu": ("ONE sample of %d uniques sharded over %d GPUs: raw r aligned on rank r %% N, one NCCL all-gather of "
                                       "new stored comparisons per split round, final tallies all-reduced" % (nraw, world)) if shard else
                                      "one sample per GPU (reference's per-sample loop); no data-path collective",
                           "l2": "256 MB buffer written between timed iterations (inputs 32 MB < 126 MB L2)",
                           "nclust": len(last["clustering"]["sequence"]), "rounds": st["n_rounds"], "shuffles": st["n_shuffles"],
                           "experimental": sorted(k for k in os.environ if k.startswith("DADA2B_") and k not in ("DADA2B_VERBOSE", "DADA2B_AB_TAG"))},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": "uniques/s", "h2d_bytes_per_step": int(est["h2d_bytes"]),
                        "d2h_bytes_per_step": int(est["d2h_bytes"]), "ms_per_step": 1e3 * t_e2e / args.steps},
                "gpu_launches": int(st["gpu_launches"]) * args.steps,
                "device_ms_per_step": dev_ms / args.steps,
                "step_ms": step_ms, "step_host_ms": step_host, "e2e_step_ms": e2e_ms,
                "kernel_ms": {k: st[k] for k in ("ms_k_classify", "ms_k_align_nw", "ms_k_align_gl", "ms_k_align_final")},
                "host_ms": {k: st[k] for k in ("ms_setup", "ms_loop", "ms_final", "ms_total")},
                "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
                "experimental_ab": ab_res, "bimera": bim_res}
        print(json.dumps(line))
    res.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
