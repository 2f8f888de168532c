#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric : GB/s of input scanned with bit-exact match ids (fsm_exec semantics)
workload (N=1): configs[1] -- one 256-state DFA (PCRE a[ -~]{7}\\z, built by the reference:
         re_comp -> fsm_determinise -> fsm_minimise; shipped as a golden fixture), 2^20 inputs
         x 1 KiB synthetic ASCII.  N>1: the batch is range-sharded, every rank scans its own
         2^20 x 1 KiB shard (weak scaling).  The per-shard result records reach every rank
         either FUSED with the scan (default: the scanning lanes store each 16 B record into
         every peer's gathered buffer over NVLink P2P; only a 4-byte NCCL handshake per step,
         on a side stream) or by ONE NCCL all-gather per step on a side stream (--gather nccl).

A "step" is one pass of the hot path over one batch.  `value` has inputs resident in HBM;
`e2e` goes through the host entry point of the C ABI (fsm_b200_exec_batch_host) with pinned
host buffers, H2D/D2H copies inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dist uniform|adversarial]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_INPUTS = 1 << 20
LENGTH = 1024
METRIC = "GB/s input scanned (bit-exact match ids)"


def load_cfg2_fsm():
    import goldenio
    cases = goldenio.load_exec_cases(os.path.join(goldenio.GOLDEN_DIR, "golden_exec.npz"))
    return next(c for c in cases if c["name"] == "cfg2:uniform")["fsm"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "MEASURED_PEAKS.json (measured)"
    return {"hbm_gbs": 6650.0}, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_leg(fsm, host_sample: np.ndarray, threads: int):
    """Times the reference's own fsm_exec (oracle/_ref, compiled from the reference sources) on
    the host cores; falls back to the oracle port when the compiled reference is absent.  The
    thread count is swept (all, 1/2, 1/4 of the host threads) and the BEST result of each mode is
    reported, so an oversubscribed or quota-limited box does not understate the CPU."""
    import reflib
    n = host_sample.shape[0]
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(host_sample.shape[1])
    flat = host_sample.reshape(-1)
    nbytes = flat.size
    sweep = sorted({max(1, threads), max(1, threads // 2), max(1, threads // 4)}, reverse=True)
    best = {"asis": (0.0, 0, None, 0.0), "amortised": (0.0, 0, None, 0.0)}
    if reflib.have_ref():
        R = reflib.Ref()
        h = R.from_flat(fsm)
        kind = "reference"
        run = lambda mode, t: R.exec_batch(h, flat, offsets, mode=mode, nthreads=t)
    else:
        O = reflib.Oracle()
        h = None
        kind = "port"
        run = lambda mode, t: O.exec_batch(fsm, flat, offsets, nthreads=t, validate_each=(mode == 0))
    for t in sweep:
        for name, mode in (("asis", 0), ("amortised", 1)):
            t0 = time.perf_counter(); rec = run(mode, t); dt = time.perf_counter() - t0
            if nbytes / dt / 1e9 > best[name][0]:
                best[name] = (nbytes / dt / 1e9, t, rec, dt)
    if h is not None:
        R.free(h)
    return {"kind": kind, "asis_gbs": best["asis"][0], "asis_threads": best["asis"][1], "asis_s": best["asis"][3],
            "amortised_gbs": best["amortised"][0], "amortised_threads": best["amortised"][1],
            "amortised_s": best["amortised"][3], "records": best["amortised"][2], "asis_records": best["asis"][2],
            "threads_swept": sweep}


def run_reference_arm(args):
    """--impl reference: the reference's CPU fsm_exec on this box's host cores, all threads,
    each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from libfsm_b200 import workloads
    fsm = load_cfg2_fsm()
    threads = host_threads()
    sample_n = int(os.environ.get("BENCH_REF_SAMPLE", 512 * threads))
    sample_n = max(1024, min(sample_n, 1 << 16))
    host = workloads.cfg2_host(sample_n, LENGTH, args.dist == "adversarial", seed=42)
    for _ in range(args.warmup):
        cpu_reference_leg(fsm, host[:max(256, sample_n // 8)], threads)
    t_total, last = 0.0, None
    for _ in range(args.steps):
        last = cpu_reference_leg(fsm, host, threads)
        t_total += last["asis_s"]        # the best thread count of the sweep
    nbytes = sample_n * LENGTH
    value = nbytes * args.steps / t_total / 1e9
    sample = f"{sample_n} x {LENGTH} B inputs of the same distribution per step (seed 42)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[1]: 256-state DFA a[ -~]{7}\\z, 2^20 x 1 KiB ASCII per GPU",
                   "distribution": args.dist, "reference_entry": "fsm_exec per input (as-is, per-call fsm_isdfa validation)"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": last["asis_threads"], "kind": last["kind"], "sample": sample,
                         "amortised_value": last["amortised_gbs"], "amortised_cores": last["amortised_threads"],
                         "threads_swept": last["threads_swept"]},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dist", default="uniform", choices=["uniform", "adversarial"])
    ap.add_argument("--variant", default="auto")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--gather-records", default="compact", choices=["compact", "full"],
                    help="fused gather payload: compact = 4-byte match ids ((ret==1)<<31 | end), full = 16-byte records")
    ap.add_argument("--handshake", default="flags", choices=["flags", "nccl", "none"],
                    help="fused gather completion signal: flags = the kernel's last CTA stores a step number into "
                         "every peer's memory; nccl = 4-byte NCCL all-reduce per step on a side stream")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N>1: fused = scanning lanes store records into every peer's buffer over NVLink P2P; "
                         "nccl = one NCCL all-gather per step on a side stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import libfsm_b200 as L
    from libfsm_b200 import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, f"WORLD_SIZE {world} != --gpus {args.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    fsm = load_cfg2_fsm()
    dfa = L.Dfa(fsm, device=local)
    L.set_exec_variant(args.variant)
    adversarial = args.dist == "adversarial"
    n = N_INPUTS
    # range shard `rank` of the global batch: its own seeded 2^20 x 1 KiB slice
    d_in = workloads.cfg2_device(n, LENGTH, adversarial, seed=42 + 1000 * rank, device=dev)
    # buffers in flight: the completion handshake / all-gather of step i is only waited for when its
    # buffer is reused at step i+nbuf, so a few steps of slack hide the collective's latency (its
    # kernel cannot co-reside with the persistent scan kernel and runs between scan launches)
    nbuf = int(os.environ.get("BENCH_NBUF", "4"))
    fused = world > 1 and args.gather == "fused"
    d_out = [torch.empty((n, 16), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    gathered = [torch.empty((world * n, 16), dtype=torch.uint8, device=dev) for _ in range(nbuf)] if world > 1 else None
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    main_stream = torch.cuda.current_stream()
    gather_done = [None] * nbuf
    ring = None
    compact = False
    token = torch.zeros(1, dtype=torch.int32, device=dev)
    if fused:
        from libfsm_b200.peer import GatherRing
        compact = args.gather_records == "compact"
        try:
            ring = GatherRing(n, world, rank, local, nbuf, elem_bytes=4 if compact else 16)
            ok = torch.ones(1, dtype=torch.int32, device=dev)
        except Exception as e:                      # no peer access on this box: NCCL all-gather instead
            print(f"[bench] rank {rank}: peer mapping failed ({e}); falling back to --gather nccl", file=sys.stderr)
            ring = None
            ok = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # every rank must agree on the path
        if int(ok.item()) == 0:
            if ring is not None:
                ring.close()
            ring, fused, compact = None, False, False
            args.gather = "nccl"
    if fused:
        peer_args = [ring.peer_slot_ptrs(b) for b in range(nbuf)]
        sig_args = [ring.signal_args(b) for b in range(nbuf)]
        own_out = [torch.empty((n, 16), dtype=torch.uint8, device=dev) for _ in range(nbuf)] if compact else None

    def local_out_ptr(b):
        # full records of this rank's own range: inside its gathered buffer (full) or beside it (compact)
        return own_out[b].data_ptr() if (fused and args.gather_records == "compact") else ring.local_slot_ptr(b)

    def step(i):
        b = i % nbuf
        if world > 1 and gather_done[b] is not None:
            main_stream.wait_event(gather_done[b])         # buffer free again
        if fused:
            # ONE kernel: scan + P2P stores of every record into every peer's gathered buffer
            # (+ with --handshake flags the completion flag, value = step number + 1)
            use_flags = args.handshake == "flags"
            dfa.exec_batch_gather(d_in, stride=LENGTH, length=LENGTH, n=n, out_ptr=local_out_ptr(b),
                                  peer_ptrs=peer_args[b][0], npeers=peer_args[b][1], compact=compact,
                                  sig_counter=sig_args[b][0] if use_flags else None,
                                  sig_flags=sig_args[b][1] if use_flags else None, sig_value=i + 1)
        else:
            dfa.exec_batch(d_in, stride=LENGTH, length=LENGTH, n=n, out=d_out[b])
        if world > 1 and (not fused or args.handshake == "nccl"):
            ev = torch.cuda.Event(); ev.record(main_stream)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                if fused:
                    dist.all_reduce(token)                  # 4-byte completion handshake, off the data path
                else:
                    dist.all_gather_into_tensor(gathered[b], d_out[b])
                gather_done[b] = torch.cuda.Event(); gather_done[b].record(side)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- parity gate: the results we are about to time are the reference's ---------------
    step(0); sync_all()
    import reflib
    oracle = reflib.Oracle()
    idx = torch.arange(0, n, 64, device=dev)
    sample_host = d_in[idx].cpu().numpy()
    off = np.arange(sample_host.shape[0] + 1, dtype=np.uint64) * np.uint64(LENGTH)
    want = oracle.exec_batch(fsm, sample_host.reshape(-1), off, nthreads=min(16, os.cpu_count() or 1))
    if fused and args.handshake == "flags":
        flags = ring.read_flags(0)                         # step 0 was launched with sig_value 1
        assert (flags == 1).all(), f"bench: completion flags {flags} != 1 after step 0"
    if fused and compact:
        mine = L.results_from_torch(own_out[0])
        assert (mine[::64] == want).all(), "bench: GPU results differ from the oracle"
        ids = ((mine["ret"] == 1).astype(np.uint32) << np.uint32(31)) | mine["end"]
        everything = ring.read(0)                          # peers' match ids (own slot is not written)
        ids_t = torch.from_numpy(ids.view(np.int32).copy()).to(dev)
        allids = torch.empty(world * n, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allids, ids_t)         # NCCL reference for the P2P-gathered ids
        torch.cuda.synchronize(dev)
        ref_ids = allids.cpu().numpy().view(np.uint32)
        for r in range(world):
            if r != rank:
                assert (everything[r * n:(r + 1) * n] == ref_ids[r * n:(r + 1) * n]).all(), "bench: fused gather != NCCL all-gather"
    elif fused:
        everything = ring.read(0)                          # this rank's gathered buffer: all ranks' records
        got = everything[rank * n:(rank + 1) * n][::64]
        assert (got == want).all(), "bench: GPU results differ from the oracle"
        # cross-check the P2P-gathered buffer against an NCCL all-gather of the same records
        mine_t = torch.from_numpy(everything[rank * n:(rank + 1) * n].view(np.uint8).reshape(n, 16).copy()).to(dev)
        dist.all_gather_into_tensor(gathered[0], mine_t)
        torch.cuda.synchronize(dev)
        assert (L.results_from_torch(gathered[0]) == everything).all(), "bench: fused gather != NCCL all-gather"
    else:
        got = L.results_from_torch(d_out[0][idx])
        assert (got == want).all(), "bench: GPU results differ from the oracle"
        if world > 1:
            mine = gathered[0][rank * n:(rank + 1) * n]
            assert torch.equal(mine, d_out[0]), "bench: all-gather slot mismatch"

    # ---- device-resident timing ----------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local); sampler.start()
    L.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_stream)
    for i in range(args.steps):
        step(i)
    if world > 1:
        main_stream.wait_stream(side)
    e1.record(main_stream)
    sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = L.launch_count()

    # kernel-only duration (CUDA events around each launch, on the launching stream)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    for a, b in kev:
        a.record(main_stream)
        if fused:
            dfa.exec_batch_gather(d_in, stride=LENGTH, length=LENGTH, n=n, out_ptr=local_out_ptr(0),
                                  peer_ptrs=peer_args[0][0], npeers=peer_args[0][1], compact=compact)
        else:
            dfa.exec_batch(d_in, stride=LENGTH, length=LENGTH, n=n, out=d_out[0])
        b.record(main_stream)
    torch.cuda.synchronize(dev)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    sampler.stop_flag = True; sampler.join()

    # ---- end to end through the host entry point, pinned host buffers --------------------
    h_in = torch.empty((n, LENGTH), dtype=torch.uint8, pin_memory=True)
    h_in.copy_(d_in)
    h_off = torch.arange(0, (n + 1) * LENGTH, LENGTH, dtype=torch.int64).pin_memory()
    h_out = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else max(3, min(args.steps, 10))
    for _ in range(2):
        dfa.exec_batch_hostptr(h_in.data_ptr(), h_off.data_ptr(), n, h_out.data_ptr())
    assert (L.results_from_torch(h_out)[::64] == want).all()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dfa.exec_batch_hostptr(h_in.data_ptr(), h_off.data_ptr(), n, h_out.data_ptr())
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    # ---- reduce over ranks (max time) ----------------------------------------------------
    t = torch.tensor([ms_total, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kernel_ms = (float(x) for x in t.cpu())
    bytes_step = n * LENGTH
    value = world * bytes_step * args.steps / (ms_total / 1e3) / 1e9
    e2e_value = world * bytes_step * e2e_steps / (e2e_ms / 1e3) / 1e9
    peaks, peak_src = measured_peaks()
    achieved = bytes_step / (kernel_ms / 1e3) / 1e9

    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r1_k1_traffic.json")
    if os.path.exists(tp) and args.dist == "uniform":
        tj = json.load(open(tp))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        traffic_src = f"{tj['source']}: dram__bytes_read.sum + dram__bytes_write.sum per launch of {tj['kernel']}"
    line = None
    if rank == 0:
        threads = host_threads()
        sample_n = max(1024, min(256 * threads, 1 << 15))
        cpu = cpu_reference_leg(fsm, workloads.cfg2_host(sample_n, LENGTH, adversarial, seed=42), threads)
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: 256-state DFA a[ -~]{7}\\z, 2^20 x 1 KiB ASCII per GPU",
                       "distribution": args.dist, "variant": args.variant, "table": dfa.info,
                       "l2": "1 GiB input per step > 126 MB L2: no flush needed",
                       "multi_gpu": ("single GPU" if world == 1 else
                                     f"range-sharded batch; scan fused with the gather: lanes store {'4 B match ids ((ret==1)<<31|end)' if args.gather_records == 'compact' else '16 B records'} into every peer's buffer over NVLink P2P; completion signal: {args.handshake}"
                                     if fused else "range-sharded batch, one NCCL all-gather of 16 B result records per step on a side stream")},
            "clocks": sampler.result(),
            "e2e": {"value": e2e_value, "unit": "GB/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": int(h_in.numel() + h_off.numel() * 8),
                    "d2h_bytes_per_step": int(h_out.numel()),
                    "entry": "fsm_b200_exec_batch_host, pinned host buffers"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "peak_source": peak_src,
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_step,
                         "traffic": traffic, "traffic_source": traffic_src},
            "cpu_baseline": {"value": cpu["asis_gbs"], "unit": "GB/s", "cores": cpu["asis_threads"], "kind": cpu["kind"],
                             "sample": f"{sample_n} x {LENGTH} B inputs, same distribution; reference fsm_exec per input (as-is); best of thread counts {cpu['threads_swept']}",
                             "amortised_value": cpu["amortised_gbs"], "amortised_cores": cpu["amortised_threads"]},
        }
        print(json.dumps(line))
    dfa.close()
    if world > 1:
        dist.barrier()
        if ring is not None:
            ring.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
