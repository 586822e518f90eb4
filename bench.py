#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the k-mer hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the fused hot path (FASTQ chunk bytes -> 2-bit codes -> k=31 rolling hash ->
bincount) over one batch of synthetic 150 bp reads that is already resident in HBM (BASELINE.json
configs[1]: 10 M x 150 bp per GPU, 3.17 GB -- far larger than the 126 MB L2, so no L2 flush is
needed between steps).  N > 1: every rank counts its own shard (weak scaling) and the step ends
with ONE NCCL all-reduce of the int64 histogram.  Timing: CUDA events on the launching stream
around exactly K steps, barrier + synchronize on both sides, max over ranks.

The JSON line also carries
  roofline     achieved algorithmic GB/s of the dominant (tile) kernel vs the measured HBM peak
  e2e          the same metric through the host-buffer C-ABI call (pinned host chunk -> H2D in
               slices overlapped with the count -> D2H of the histogram), per step
  cpu_baseline the oracle's NumPy port of the reference path on a bounded sample (rank 0, N=1)
`--impl reference` times that CPU port on all host cores instead (the reference itself cannot be
imported: its npstructures dependency is absent).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
RECORD_BYTES = 317
METRIC = "Gbases/s k=31 hash+count on 150bp reads"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU and step")
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--buckets", type=int, default=1 << 14,
                    help="histogram bins: hash mod buckets (2^14 = the shared-memory-privatised table)")
    ap.add_argument("--window", type=int, default=0, help="minimizer window in bases (0 = plain k-mers)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configurations")
    ap.add_argument("--cpu-sample-reads", type=int, default=200_000)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU baseline = the oracle's NumPy port of the reference path (test infrastructure; timed, not shipped)
# ------------------------------------------------------------------------------------------------
_CPU_CHUNK = None


def _cpu_init(chunk_reads):
    """Every worker builds one ~4.8 MB synthetic chunk once (not timed)."""
    global _CPU_CHUNK
    from oracle import bnp_oracle as oracle
    _CPU_CHUNK = oracle.synthetic_fastq((os.getpid() % 1000) * chunk_reads, chunk_reads)


def _cpu_worker(job):
    k, buckets, window = job
    from oracle import bnp_oracle as oracle
    t0 = time.perf_counter()
    hist, size, n_bases = oracle.fastq_chunk_kmer_counts(_CPU_CHUNK, k, buckets, True, window_size=window)
    return time.perf_counter() - t0, n_bases, int(hist.sum())


def cpu_baseline(k, buckets, window, sample_reads, n_procs):
    """The reference op sequence (oracle port) on `n_procs` processes, each over ~4.8 MB chunks (bionumpy's default
    min_chunk_size, io/parser.py:96), input already in RAM.  Returns (Gbases/s, bases, wall seconds)."""
    chunk_reads = 15_000
    n_jobs = max(n_procs, sample_reads // chunk_reads)
    jobs = [(k, buckets, window)] * n_jobs
    if n_procs == 1:
        _cpu_init(chunk_reads)
        _cpu_worker(jobs[0])
        res = [_cpu_worker(j) for j in jobs]
        wall = sum(r[0] for r in res)
    else:
        with mp.get_context("fork").Pool(n_procs, initializer=_cpu_init, initargs=(chunk_reads,)) as pool:
            pool.map(_cpu_worker, jobs[:n_procs], chunksize=1)          # warm up (imports, page faults)
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs, chunksize=1)
            wall = time.perf_counter() - t0
    bases = sum(r[1] for r in res)
    return bases / wall / 1e9, bases, wall


# ------------------------------------------------------------------------------------------------
def sample_clocks(stop_evt, out):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    lines = []

    def reader():
        for line in p.stdout:
            lines.append(line)
    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop_evt.wait()
    p.terminate()
    t.join(timeout=2)
    out.extend(lines)


def summarize_clocks(lines, gpu_index):
    sm, mx, reasons = [], [], set()
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 9 or f[0] != str(gpu_index):
            continue
        try:
            sm.append(float(f[1]))
            mx.append(float(f[2]))
        except ValueError:
            continue
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return None
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(key):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port, all host cores) on the same config."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    sample = max(args.cpu_sample_reads, 15_000 * cores * 2)      # at least two ~4.8 MB chunks per core and step
    for i in range(args.warmup + args.steps):
        v, bases, wall = cpu_baseline(args.k, args.buckets, args.window, sample, cores)
        if i >= args.warmup:
            vals.append((v, wall))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(w for _, w in vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "Gbases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"synthetic {args.reads}x150bp FASTQ per GPU, k={args.k}, hash mod {args.buckets} bincount"
                               + (f", minimizer window {args.window}" if args.window else ""),
                   "sample": f"{sample} reads per step in ~4.8 MB chunks (the reference's default chunk size)"},
        "cpu_baseline": {"value": round(value, 4), "unit": "Gbases/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} reads x {args.steps} steps, one process per core"},
        "e2e": {"value": round(value, 4), "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from bionumpy_b200 import ops, _native as nv
    from bionumpy_b200.distributed import all_reduce_histogram

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    lib = nv.lib()
    n = args.reads
    first_record = rank * n                       # shard = disjoint record range (weak scaling)
    chunk = ops.synth_fastq(n, first_record=first_record, device=dev)
    n_bytes = chunk.numel()
    hist = torch.zeros(args.buckets, dtype=torch.int64, device=dev)
    status = nv.new_status(dev)

    def step():
        hist.zero_()
        nv.check(lib.bnpk_status_init(nv.ptr(status), nv.stream_ptr()))
        ops.chunk_kmer_count(chunk, args.k, args.buckets, hist=hist, window_size=args.window, status=status)
        all_reduce_histogram(hist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # correctness guard on the measured configuration: every k-mer landed in exactly one bin
    st = ops.read_status(status)
    per_read = (READ_LEN - (args.window or args.k) + 1)
    assert st.n_records == n and st.n_values == n * per_read, st.words
    assert int(hist.sum().item()) == world * n * per_read

    stop_evt, clock_lines = threading.Event(), []
    clk_thread = None
    if rank == 0:
        clk_thread = threading.Thread(target=sample_clocks, args=(stop_evt, clock_lines), daemon=True)
        clk_thread.start()
        time.sleep(0.3)

    lib.bnpk_profile_enable(1)
    launches0 = lib.bnpk_launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = lib.bnpk_launch_count() - launches0
    import ctypes
    tot_ms, n_l = ctypes.c_double(0), ctypes.c_uint64(0)
    lib.bnpk_profile_read(ctypes.byref(tot_ms), ctypes.byref(n_l))
    lib.bnpk_profile_enable(0)
    if rank == 0:
        time.sleep(0.2)
        stop_evt.set()
        clk_thread.join(timeout=3)
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * n * READ_LEN / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel (tile_tma_kernel): algorithmic bytes per launch / event-timed duration
    alg_bytes = n_bytes + 16 * n + 8 * args.buckets       # chunk once + row vector + histogram
    kern_ms = tot_ms.value / max(n_l.value, 1)
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic_from_profiles("tile_kernel_dram_bytes_per_10M_reads"),
                "kernel": "bnpk::tma::tile_tma_kernel (fused split+encode+hash+count, shared-memory staged)", "kernel_ms": round(kern_ms, 4),
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "kernel_share_of_step": round(kern_ms / ms_per_step, 3)}

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8->u32 codes, int64 counts", "data": "synthetic",
        "config": {"workload": f"synthetic {n}x150bp FASTQ per GPU ({n_bytes / 1e9:.2f} GB resident in HBM, > L2 so no "
                               f"flush between steps), 2-bit encode + k={args.k} rolling hash + bincount(hash mod "
                               f"{args.buckets})" + (f", minimizer window {args.window}" if args.window else ""),
                   "reads_per_gpu": n, "k": args.k, "buckets": args.buckets, "window": args.window,
                   "parallelism": f"reads sharded over {world} GPU(s), one NCCL all-reduce of the histogram per step"},
        "roofline": roofline, "gpu_launches": int(launches),
    }
    if rank == 0:
        clocks = summarize_clocks(clock_lines, local_rank)
        if clocks:
            line["clocks"] = clocks

    # ---- secondary configurations (same timing rules, fewer steps) ---------------------------------
    if not args.no_extra and world == 1:
        extra = {}
        for name, b, w in (("buckets_2^24_global_atomics", 1 << 24, 0), ("minimizers_w41_buckets_2^14", 1 << 14, 41),
                           ("k5_exact_4^5_bins", 4 ** 5, 0)):
            kk = 5 if name.startswith("k5") else args.k
            h2 = torch.zeros(b, dtype=torch.int64, device=dev)
            for _ in range(2):
                ops.chunk_kmer_count(chunk, kk, b, hist=h2, window_size=w, status=status)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                h2.zero_()
                ops.chunk_kmer_count(chunk, kk, b, hist=h2, window_size=w, status=status)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            extra[name] = {"ms_per_step": round(ms, 3), "Gbases/s": round(n * READ_LEN / ms / 1e6, 2),
                           "frac_of_hbm_roofline": round((n_bytes + 16 * n + 8 * b) / (ms * 1e-3) / 1e9 / peak, 4)}
            del h2
        line["extra"] = extra

    # ---- end to end through the host-buffer C-ABI call -----------------------------------------------
    e2e = None
    try:
        host = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
        host.copy_(chunk)
        torch.cuda.synchronize()
        del chunk
        pipe = ops.HostPipeline(n_bytes, slice_bytes=32 << 20)
        res_host = torch.empty(args.buckets, dtype=torch.int64).pin_memory()

        def e2e_step():
            hist.zero_()
            pipe.kmer_count(host, args.k, hist, window_size=args.window)     # H2D slices overlapped with the count
            all_reduce_histogram(hist)
            res_host.copy_(hist)                                             # the step's result back on the host
            torch.cuda.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        assert int(res_host.sum().item()) == world * n * per_read
        e2e = {"value": round(world * n * READ_LEN / dt / 1e9, 3), "unit": "Gbases/s",
               "h2d_bytes_per_step": n_bytes, "d2h_bytes_per_step": args.buckets * 8 + 128,
               "ms_per_step": round(dt * 1e3, 3), "steps": args.e2e_steps,
               "h2d_GB/s": round(n_bytes / dt / 1e9, 2),
               "api": "bnpk_pipeline_kmer_count_host (pinned host chunk -> sliced H2D || fused count -> D2H histogram)"}
        pipe.close()
    except Exception as exc:  # pragma: no cover
        e2e = {"error": repr(exc)}
    line["e2e"] = e2e

    # ---- CPU baseline: the reference path's NumPy port on this box's host cores (rank 0, N = 1) -------
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        v1, bases1, wall1 = cpu_baseline(args.k, args.buckets, args.window, 60_000, 1)
        n_sample = 15_000 * cores * 6                         # six ~4.8 MB chunks per core
        vN, basesN, wallN = cpu_baseline(args.k, args.buckets, args.window, n_sample, cores)
        line["cpu_baseline"] = {"value": round(vN, 4), "unit": "Gbases/s", "cores": cores, "kind": "port",
                                "sample": f"{n_sample} synthetic reads = {basesN / 1e9:.2f} Gbases in ~4.8 MB chunks, one process "
                                          f"per core, {wallN:.1f} s wall; single core: {v1:.4f} Gbases/s on 60000 reads",
                                "single_core_value": round(v1, 4)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
