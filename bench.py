#!/usr/bin/env python
"""bench.py — pileup positions/s of the B200 engine on the BASELINE.json workload.

Workload (config C3): synthetic 10 Mb contig, 30x depth, 150 bp reads,
flags -q 20 -b 20, one region covering the contig, per GPU (weak scaling: every rank gets its
own 10 Mb shard with its own seed; sites are independent, so there is no data-path collective).

A "step" is one pass of the hot path over the whole batch: init + K0 read_precompute + K1 pileup.
  value : sites/s with the decoded batch already resident in HBM (kernels only)
  e2e   : sites/s through the C-ABI push path with HOST buffers (admission, H2D, kernels, D2H of results)
  --impl reference : the UNMODIFIED reference binary (oracle/_ref/bam-readcount) on a bounded sample of
                     the same synthetic workload, one process per host core.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONTIG_LEN = 10_000_000
DEPTH = 30
FLAGS = dict(min_mapq=20, min_bq=20)
WORKLOAD = "C3: synthetic 10 Mb contig, 30x, 150 bp reads, -q 20 -b 20, region chr1:1-10000000 (per GPU)"
METRIC = "pileup positions/sec"
UNIT = "positions/s"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def profiled_traffic():
    """DRAM bytes per K1 launch from the newest committed `ncu --set full` capture (profiles/traffic_*.json)."""
    import glob
    fs = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json")))
    if not fs:
        return None, None
    try:
        d = json.load(open(fs[-1]))
        return float(d["pileup_kernel"]["traffic"]), os.path.basename(fs[-1])
    except Exception:
        return None, None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference binary on a bounded sample
# ------------------------------------------------------------------------------------------------
def make_reference_sample(workdir: str, sample_bp: int, seed: int):
    """Write the first `sample_bp` bases of the synthetic workload as ref.fa + s.bam (+.bai)."""
    from bam_readcount_b200 import synth
    from oracle.oracle import REF_SAMTOOLS
    L = sample_bp + 400
    ref = synth.synth_reference(L, seed)
    batch = synth.synth_reads(ref, DEPTH, seed=seed)
    synth.write_fasta(os.path.join(workdir, "ref.fa"), "chr1", ref)
    synth.write_sam(os.path.join(workdir, "s.sam"), batch, [("chr1", L)])
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", os.path.join(workdir, "s.bam"), os.path.join(workdir, "s.sam")])
    subprocess.check_call([REF_SAMTOOLS, "index", os.path.join(workdir, "s.bam")])
    os.remove(os.path.join(workdir, "s.sam"))
    return batch.n_reads


def run_reference_step(workdir: str, sample_bp: int, procs: int):
    """One step: `procs` concurrent reference processes, each over the sample region. Returns (sites, seconds)."""
    from oracle.oracle import REF_BIN
    cmd = [REF_BIN, "-w", "0", "-q", "20", "-b", "20", "-f", os.path.join(workdir, "ref.fa"), os.path.join(workdir, "s.bam"),
           f"chr1:1-{sample_bp}"]
    t0 = time.perf_counter()
    ps = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL) for _ in range(procs)]
    lines = 0
    outs = []

    def drain(p):
        n = 0
        for chunk in iter(lambda: p.stdout.read(1 << 20), b""):
            n += chunk.count(b"\n")
        outs.append(n)
    ths = [threading.Thread(target=drain, args=(p,)) for p in ps]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for p in ps:
        p.wait()
    dt = time.perf_counter() - t0
    lines = sum(outs)
    return lines, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.oracle import have_reference_binary
    if not have_reference_binary():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bam-readcount not built (run oracle/build_ref.sh where /root/reference exists)"}))
        return 0
    cores = host_cores()
    procs = cores
    sample_bp = args.ref_sample_bp
    wd = tempfile.mkdtemp(prefix="brc_ref_")
    try:
        make_reference_sample(wd, sample_bp, 1234)
        for _ in range(args.warmup):
            run_reference_step(wd, min(sample_bp, 20000), procs)
        tot_sites, tot_t = 0, 0.0
        for _ in range(args.steps):
            s, dt = run_reference_step(wd, sample_bp, procs)
            tot_sites += s
            tot_t += dt
        one_sites, one_t = run_reference_step(wd, sample_bp, 1)
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    value = tot_sites / tot_t
    sample = (f"{procs} concurrent reference processes, each over the first {sample_bp} bp of the synthetic workload "
              f"(same generator/seed), stdout discarded; 1-process rate {one_sites / one_t:.0f} positions/s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32+f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_bp_per_process": sample_bp, "processes": procs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def our_arm(args):
    import torch
    import torch.distributed as dist
    from bam_readcount_b200 import synth
    from bam_readcount_b200.engine import CReadBatch, CRegion, CResults, Engine, Results

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    L = args.contig_len
    seed = 1234 + 1000 * rank
    t_gen = time.perf_counter()
    ref = synth.synth_reference(L, seed)
    batch = synth.synth_reads(ref, DEPTH, seed=seed)
    t_gen = time.perf_counter() - t_gen
    n = batch.n_reads

    eng = Engine(device=local, **FLAGS)
    eng.set_reference(0, "chr1", L, ref.tobytes(), 0)

    # ---- device-resident inputs (torch owns the memory; the engine only sees raw pointers) ----
    def dev(a, dt=None):
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint16:
            a = a.view(np.int16)
        elif a.dtype == np.uint32:
            a = a.view(np.int32)
        elif a.dtype == np.uint64:
            a = a.view(np.int64)
        return torch.from_numpy(a).cuda()
    # pools padded so 16-byte vector reads past the last record stay in bounds
    pad = np.zeros(64, dtype=np.uint8)
    d = dict(pos=dev(batch.pos), flag=dev(batch.flag), mapq=dev(batch.mapq), lib=dev(batch.lib), l_qseq=dev(batch.l_qseq),
             nm=dev(batch.nm), sm=dev(batch.sm), cigar_off=dev(batch.cigar_off), cigar=dev(np.concatenate([batch.cigar, np.zeros(16, np.uint32)])),
             seq_off=dev(batch.seq_off), seq=dev(np.concatenate([batch.seq, pad])), qual_off=dev(batch.qual_off),
             qual=dev(np.concatenate([batch.qual, pad])))
    cb = CReadBatch(n, None, d["pos"].data_ptr(), d["flag"].data_ptr(), d["mapq"].data_ptr(), d["lib"].data_ptr(),
                    d["l_qseq"].data_ptr(), d["nm"].data_ptr(), d["sm"].data_ptr(), d["cigar_off"].data_ptr(), d["cigar"].data_ptr(),
                    d["seq_off"].data_ptr(), d["seq"].data_ptr(), d["qual_off"].data_ptr(), d["qual"].data_ptr())
    reg = CRegion(0, 0, L, 0, 0, n, 0, 0, L)
    lib = eng.lib
    eng._check(lib.brc_plan_device(eng.h, C.byref(reg), 1, n, 0))
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)

    def step():
        eng._check(lib.brc_run_device(eng.h, C.byref(cb), None, sptr))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    k0_ms, k1_ms = [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        k0_ms.append(eng.stage_ms(0))   # CUDA events recorded by the engine on the launching stream
        k1_ms.append(eng.stage_ms(1))
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() * args.steps

    # ---- what was computed (outside the timed region) ----
    eng._check(lib.brc_fetch_device_results(eng.h, sptr))
    r = CResults()
    eng._check(lib.brc_get_results(eng.h, C.byref(r)))
    res = Results(r)
    n_sites = int((res.ncover[0] > 0).sum())
    n_events = int(res.ncover[0].sum())
    n_keys = int((res.pstats[0, 0] > 0).sum()) + int(res.n_sec)
    n_cig = int(batch.cigar.shape[0])
    # ALGORITHMIC bytes of one pass: reads + reference + 16 B/site + 52 B/key
    alg_bytes = 16 * n + 4 * n_cig + int(batch.seq.shape[0]) + int(batch.qual.shape[0]) + L + 16 * n_sites + 52 * n_keys

    # ---- e2e: host buffers through the push path (admission + H2D + kernels + D2H) ----
    e2e_ms = None
    h2d = 16 * 0
    if args.e2e_steps > 0:
        from bam_readcount_b200.engine import pin_batch
        eng2 = Engine(device=local, **FLAGS)
        eng2.set_reference(0, "chr1", L, ref.tobytes(), 0)
        hbatch = pin_batch(batch)      # the caller's host buffers, page-locked
        times = []
        for it in range(args.e2e_steps + 1):
            barrier()
            t0 = time.perf_counter()
            eng2.reset()
            eng2.begin_region(0, 0, L, False)
            eng2.push_reads(hbatch)
            eng2.end_region()
            eng2._check(eng2.lib.brc_compute(eng2.h))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it > 0:
                times.append(dt)
        e2e_ms = 1000.0 * sum(times) / len(times)
        h2d = (batch.pos.nbytes + batch.flag.nbytes + batch.mapq.nbytes + batch.lib.nbytes + batch.l_qseq.nbytes + batch.nm.nbytes +
               batch.sm.nbytes + batch.cigar_off.nbytes + batch.cigar.nbytes + batch.seq_off.nbytes + batch.seq.nbytes +
               batch.qual_off.nbytes + batch.qual.nbytes + 4 * n)
        d2h = res.n_slots * (4 + 4 + 1 + 1 + 4 + 52) + int(res.n_sec) * (4 + 1 + 4 + 8 + 4 + 52)
        eng2.close()

    # ---- reduce over ranks: max time, sum of units ----
    t = torch.tensor([elapsed_ms, e2e_ms or 0.0], device="cuda", dtype=torch.float64)
    u = torch.tensor([n_sites, n_events], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_max = float(t[0]), float(t[1])
    tot_sites, tot_events = float(u[0]), float(u[1])

    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = tot_sites / (ms_per_step / 1000.0)
        peak, peak_src = measured_peak_gbs()
        k1 = sum(k1_ms) / len(k1_ms)
        k0 = sum(k0_ms) / len(k0_ms)
        ach = alg_bytes / (k1 / 1000.0) / 1e9
        traffic, traffic_src = profiled_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32+f32 (f64 for one add)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "contig_len_per_gpu": L, "reads_per_gpu": n, "sites_per_gpu": n_sites,
                       "events_per_gpu": n_events, "keys_per_gpu": n_keys, "events_per_s": tot_events / (ms_per_step / 1000.0),
                       "l2": "inputs (%.0f MB/GPU) larger than the 126 MB L2; no flush" % (alg_bytes / 1e6),
                       "sharding": "one 10 Mb shard per GPU, no data-path collective", "gen_s": round(t_gen, 1)},
            "roofline": {"bound": "hbm", "kernel": "pileup_kernel (K1)", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "k1_ms": k1, "k0_ms": k0,
                         "step_frac": alg_bytes / (ms_per_step / 1000.0) / 1e9 / peak},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": tot_sites / (e2e_max / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_max}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle.oracle import have_reference_binary
                if have_reference_binary():
                    wd = tempfile.mkdtemp(prefix="brc_cpu_")
                    cores = host_cores()
                    make_reference_sample(wd, args.ref_sample_bp, 1234)
                    s, dt = run_reference_step(wd, args.ref_sample_bp, cores)
                    shutil.rmtree(wd, ignore_errors=True)
                    line["cpu_baseline"] = {"value": s / dt, "unit": UNIT, "cores": cores, "kind": "reference",
                                            "sample": f"{cores} concurrent reference processes x first {args.ref_sample_bp} bp of the same synthetic workload, stdout discarded"}
            except Exception as ex:  # the baseline is a reported extra, never fatal for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--contig-len", type=int, default=CONTIG_LEN)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-sample-bp", type=int, default=0, help="bases of the synthetic workload each reference process handles per step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.ref_sample_bp <= 0:   # keep the whole --steps K run within a few minutes (≈6 k positions/s per process when all cores are busy)
            args.ref_sample_bp = max(5_000, min(150_000, 720_000 // max(args.steps, 1)))
        return reference_arm(args)
    if args.ref_sample_bp <= 0:
        args.ref_sample_bp = 150_000
    return our_arm(args)


if __name__ == "__main__":
    sys.exit(main())
