#!/usr/bin/env python
"""bench.py - dBG build throughput (G k-mers/s) of the B200 path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg2|cfg3|small] [--k 27]

One JSON line on stdout (rank 0).  A *step* is one full dBG build of the
workload: FASTA scan/pack (K1) + table reset + k-mer extraction into
hash-partitioned update records (K2a) + record insertion (K3).  Metric = k-mer insertions / s, an insertion being
one k-mer occurrence on one strand: 2 * sum max(n_r - k + 1, 1) over records
(BASELINE.md section 3).

 value   : device-resident input (the FASTA bytes already in HBM), CUDA events
           around exactly K steps, max over ranks.
 e2e     : the same build through the public host API with HOST buffers: pinned
           FASTA bytes -> H2D, build, D2H of the table statistics + checksum.
 roofline: the dominant kernel (k3_insert_records), timed live with CUDA events on
           the launching stream; algorithmic bytes = 16 B per insertion
           (SURVEY.md 8d).
 cpu_baseline: the reference's numba code (oracle/_ref, kind "reference") or,
           if that is unavailable, the C port (oracle/, kind "port"), one core,
           on a bounded sample of the same workload.
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

METRIC = "dbg_build_kmers_per_s"
UNIT = "G k-mers/s"


def workload(name, rank=0):
    from pangenome_b200 import synth
    if name == "cfg2":
        return synth.pangenome(10, 5_000_000), "synthetic 10 x 5 Mbp genomes, 1% SNP (BASELINE configs[1])"
    if name == "cfg3":
        return synth.pangenome(200, 5_000_000), "synthetic 200 x 5 Mbp pangenome (BASELINE configs[2])"
    if name == "small":
        return synth.pangenome(4, 1_000_000), "synthetic 4 x 1 Mbp genomes, 1% SNP"
    raise SystemExit("unknown workload %r" % name)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML (every ~2 ms) DURING the timed region."""

    def __init__(self, index=0):
        self.sm, self.reasons, self.max_mhz, self.stop_flag, self.th = [], set(), None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:
            self.nv = None
            self.err = str(e)

    def _loop(self):
        nv = self.nv
        names = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40),
                 ("sw_thermal_slowdown", 0x20), ("hw_power_brake_slowdown", 0x80))
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names:
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.th.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm)}


def cpu_reference_rate(data, k, sample_bases, prefer="reference"):
    """Time the CPU dBG stage on the first records of `data` totalling about
    `sample_bases` bases.  Returns (M inserts/s, kind, cores, sample text)."""
    import numpy as np
    # cut the sample at a record boundary
    pos, acc, cut = 0, 0, len(data)
    while True:
        nxt = data.find(b"\n>", pos + 1)
        if nxt < 0:
            break
        acc = nxt
        pos = nxt
        if acc >= sample_bases:
            cut = nxt + 1
            break
    sample = data[:cut]
    if len(sample) > sample_bases * 1.3:      # a single huge record: truncate its body instead
        sample = data[:int(sample_bases)]
        sample = sample[:sample.rfind(b"\n") + 1]
    kind = "port"
    if prefer == "reference":
        try:
            from oracle import refrun
            if refrun.available():
                refrun.run(b">w\n" + b"ACGTTGCATG" * 20 + b"\n", k, stages="dbg")      # JIT warm-up, untimed
                tm = {}
                res = refrun.run(sample, k, stages="dbg", timings=tm)
                import oracle
                n_ins = oracle.run(sample, k, stages=1)["n_inserts"]
                return n_ins / tm["dbg"] / 1e9, "reference", 1, "%d bytes / %d insertions of the workload, numba steady state (JIT excluded)" % (len(sample), n_ins)
        except Exception as e:  # numba missing on the box, etc.
            sys.stderr.write("bench: reference unavailable (%s), using the C port\n" % e)
    import oracle
    res = oracle.run(sample, k, stages=1)
    return res["n_inserts"] / res["times"]["dbg"] / 1e9, kind, 1, "%d bytes / %d insertions of the workload, C port of the reference" % (len(sample), res["n_inserts"])


def run_reference_arm(args):
    """--impl reference: the reference's own CPU code (numba, oracle/_ref; the C port if numba is missing) on a
    bounded sample of the same workload per step, sized so that the whole run stays within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data, wl = workload(args.workload if args.gpus == 1 or args.workload != "cfg2" else "cfg2")
    total_steps = max(1, args.warmup + args.steps)
    budget_s = 150.0
    # calibrate on a small sample (also warms the JIT), then size the per-step sample to the time budget
    r0, kind, cores, text = cpu_reference_rate(data, args.k, 100_000)
    per_step_s = budget_s / total_steps
    sample_bases = int(min(1_000_000, max(50_000, r0 * 1e9 * per_step_s / 2)))     # 2 insertions per base
    rates = []
    t0 = time.time()
    for i in range(total_steps):
        r, kind, cores, text = cpu_reference_rate(data, args.k, sample_bases)
        if i >= args.warmup:
            rates.append(r)
    val = sum(rates) / len(rates)
    ms = 1e3 * (time.time() - t0) / total_steps
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": {"workload": wl, "k": args.k, "rc": True},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": text,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--k", type=int, default=27)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from pangenome_b200 import engine, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1:
        from pangenome_b200 import multigpu
        return multigpu.bench(args, world, rank, local, ClockSampler)

    data, wl = workload(args.workload)
    k = args.k
    host = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    d_fasta = host.to("cuda", non_blocking=True)
    torch.cuda.synchronize()

    # one untimed build to know the unit count; buffers (table + record buckets) stay resident across steps
    packed = engine.PackedSeqs(d_fasta)
    n_ins = packed.n_insertions(k)
    n_rec = packed.record_prefix(2 ** 63, 2)
    builder = engine.TwoPhaseBuilder(k, _lib.PG_MODE_CANONICAL, packed.n_positions(k), estimate=False)
    table = builder.build(packed, n_rec)
    torch.cuda.synchronize()
    builder.verify()
    cap = table.capacity
    est_keys = builder.last_estimate
    used, entries = table.count()
    ref_table, _ = engine.build_dbg(packed, k)              # the fused single-launch path must agree
    ref_sum = ref_table.checksum()
    if used == 0 or table.checksum() != ref_sum:
        raise SystemExit("bench: two-phase build disagrees with the fused build")
    del ref_table
    stream = torch.cuda.current_stream()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kev = {}

    def step_device(record=False):
        # no host synchronisation inside a step: K1's record index stays on the device and K2a / K3 take
        # their bounds from it (pg_kmer_partition_dev), so the host runs ahead of the GPU
        builder.begin()                                     # empty the table: epoch bump, no HBM traffic
        p = engine.PackedSeqs(d_fasta, lazy=True)           # K1 (3 launches)
        return builder.build_async(p, ev=kev if record else None)    # K2a, count_short, K3

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record(stream)
    t = None
    for _ in range(args.steps):
        t = step_device(record=True)
    e1.record(stream)
    torch.cuda.synchronize()
    clk = clocks.stop()
    builder.verify()
    ms_total = e0.elapsed_time(e1)
    ms_step = ms_total / args.steps
    value = n_ins / (ms_step * 1e-3) / 1e9
    ins_ms = sum(a.elapsed_time(b) for a, b in kev["insert"]) / len(kev["insert"])
    part_ms = sum(a.elapsed_time(b) for a, b in kev["partition"]) / len(kev["partition"])

    # end to end through the public API, host buffers: every step uploads its own copy of the FASTA bytes from
    # pinned host memory (H2D inside the timed region), builds, and reads the table statistics back (D2H).
    # The upload of step i+1 is issued on a copy stream before step i's result is awaited (double-buffered
    # device input), the way a loader feeds a stream of files; each step still waits for ITS upload.
    copy_stream = torch.cuda.Stream()
    dev_in = [torch.empty_like(d_fasta) for _ in range(2)]

    k1_done = [None, None]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            if k1_done[i % 2] is not None:
                copy_stream.wait_event(k1_done[i % 2])       # K1 of the step that last read this buffer (device-side wait)
            dev_in[i % 2].copy_(host, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        return done

    def run_e2e(n):
        # Two steps in flight: while step i runs, the host reads step i-1's result (table statistics, D2H into
        # pinned memory) and the copy stream uploads step i+1's input.  Every step's input goes H2D and every
        # step's result comes D2H and is looked at by the host inside the timed region.
        nxt = upload(0)
        st_, pending = None, None
        for i in range(n):
            cur = nxt
            if i + 1 < n:
                nxt = upload(i + 1)
            builder.begin()                  # empty the table (epoch bump)
            stream.wait_event(cur)
            p = engine.PackedSeqs(dev_in[i % 2], lazy=True)
            k1_done[i % 2] = torch.cuda.Event()
            k1_done[i % 2].record(stream)
            tt = builder.build_async(p)
            fut = tt.stats_async(i)          # D2H of the statistics (distinct keys, overflow flag, short records, ...)
            if pending is not None:
                st_ = pending.wait()
                if int(st_[_lib.PG_STAT_OVERFLOW]):
                    raise SystemExit("bench: table overflow inside the e2e loop")
            pending = fut
        return pending.wait()
    run_e2e(2)
    torch.cuda.synchronize()
    g0, g1 = ev(), ev()
    g0.record(stream)
    st = run_e2e(args.steps)
    g1.record(stream)
    torch.cuda.synchronize()
    builder.verify()
    if int(st[_lib.PG_STAT_USED]) != used:
        raise SystemExit("bench: e2e build reports %d distinct keys, expected %d" % (int(st[_lib.PG_STAT_USED]), used))
    e2e_ms = g0.elapsed_time(g1) / args.steps
    e2e_val = n_ins / (e2e_ms * 1e-3) / 1e9
    cs = t.checksum()
    if cs != ref_sum:
        raise SystemExit("bench: table built inside the timed region has the wrong checksum")

    peak, peak_src = peaks()
    alg_bytes = 16.0 * n_ins
    part_bytes = 0.25 * packed.n_bases + 16.0 * packed.n_positions(k)
    achieved = alg_bytes / (ins_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "insert_traffic.json")
    if os.path.isfile(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(args.workload)
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic",
        "config": {"workload": wl, "k": k, "rc": True, "insertions_per_step": n_ins, "bases": packed.n_bases,
                   "table_slots": cap, "table_bytes": cap * 16, "distinct_canonical_keys": used,
                   "estimated_keys_from_1_in_256_sample": est_keys, "load_factor": used / cap,
                   "l2": "every step writes and re-reads %.1f GB of update records and randomly updates the %.1f GB table (both >> 126 MB L2), which evicts the input"
                         % (n_ins / 2 * 16 / 1e9, cap * 16 / 1e9)},
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(host.numel()),
                "d2h_bytes_per_step": int(8 * _lib.PG_STAT_WORDS),
                "pipelining": "input i+1 uploads and result i-1 is read while step i runs"},
        "gpu_launches": (3 + builder.launches_per_build) * args.steps,
        "clocks": clk,
        "roofline": {"kernel": "k3_insert_records", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "ms_per_launch": ins_ms, "algorithmic_bytes_per_launch": alg_bytes,
                     "convention": "insert: 16 B per insertion (SURVEY 8d), 2 insertions per record",
                     "other_kernels": {"k2a_partition": {"ms_per_launch": part_ms, "algorithmic_bytes_per_launch": part_bytes,
                                                         "achieved": part_bytes / (part_ms * 1e-3) / 1e9,
                                                         "frac": part_bytes / (part_ms * 1e-3) / 1e9 / peak,
                                                         "convention": "0.25 B/base read + 16 B/record written (materialised for the exchange)"}}},
        "checksum": list(cs),
    }
    if not args.no_cpu_baseline:
        r, kind, cores, text = cpu_reference_rate(data, k, 5_000_000)
        line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": kind, "sample": text,
                                "host_cpus": os.cpu_count()}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
