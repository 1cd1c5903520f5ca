#!/usr/bin/env python3
"""bench.py -- headline benchmark of the sample-barcode assignment path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

One "step" = one pass of the assignment hot path (N gate + match + per-sample/hop counters) over one batch of
synthetic reads of the named config, per GPU.  The timed input is the format north_star (a) names: 2-bit base planes
+ invalid plane, one 64-bit compact record per read (include/sgdx.h), resident in HBM; the ASCII entry points are
measured beside it under the key "ascii".  Weak scaling: every rank owns its own shard (rank r holds reads
[r*n, (r+1)*n) of the seeded stream), no data-path collective; counters are merged on the host and, at N > 1, the
merged per-sample counters and hop map are compared with the single-GPU result of the same stream
(tests/golden/bench_digests.json).  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "barcode-assigned read pairs/sec (device-timed)"
UNIT = "read pairs/s"
ALG_BYTES = lambda L: (3 * L + 7) // 8 + 3  # 2-bit planes + N mask in, u16 index + u8 distance out
DIGESTS = os.path.join(ROOT, "tests", "golden", "bench_digests.json")


def workload_desc(w, n):
    il = f"{w.index1_len}+{w.barcode_len - w.index1_len}" if w.index1_len else f"{w.barcode_len} inline"
    return (f"{w.name}: {w.num_samples}-sample {il} bp, {n} synthetic read pairs per GPU per step, "
            f"--allowed-mismatches {w.max_mismatches} --min-delta {w.min_delta} --free-ns {w.free_ns}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nme, v in zip(names, r[5:9]):
                    if v.lower() == "active":
                        reasons.add(nme)
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference's matcher (oracle/ is test infrastructure; these two functions and
# nothing else in this file execute it).  Reads come from the oracle-side build of the generator, so the CPU legs
# never touch libsgdx.so.
# ------------------------------------------------------------------------------------------------
def cpu_port_run(orc, w, table, reads, threads):
    kind = orc.select_matcher(w.barcode_len)
    cfg = orc.Config(w.num_samples, w.barcode_len, w.max_mismatches, w.min_delta, w.free_ns,
                     -1 if w.max_no_calls is None else w.max_no_calls, kind, w.index1_len,
                     w.barcode_len - w.index1_len if w.index1_len else 0)
    o = orc.COracle(cfg, table, native=True)
    res = o.run(reads, threads=threads, use_cache=True, want_assignments=True)
    return res["seconds"]


def cpu_sample(orc, w, table, threads, target_s, cap=96_000_000):
    """A time-targeted sample of the workload's stream: about target_s seconds of CPU work on this box."""
    probe = 1_000_000
    secs = cpu_port_run(orc, w, table, orc.synth_reads(w, 0, probe, table, threads), threads)
    n = int(min(max(probe, probe * target_s / max(secs, 1e-3)), cap))
    return n, orc.synth_reads(w, 0, n, table, threads)


def cpu_baseline(w, target_s=12.0):
    """The oracle port (reference algorithm incl. its per-thread 100k LRU memo, matcher.rs:307-328), all host
    cores, on a bounded sample of the same seeded stream."""
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    table = orc.synth_table(w)
    n, reads = cpu_sample(orc, w, table, threads, target_s)
    secs = cpu_port_run(orc, w, table, reads, threads)
    return {"value": n / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} reads of the same seeded stream, {secs:.2f} s wall, oracle/sgdx_oracle.c "
                      f"-O3 -march=native, per-thread 100k-entry LRU memo as in the reference"}


def orc_kind_name(orc, w):
    return "pre-compute" if orc.select_matcher(w.barcode_len) == orc.MATCHER_PRECOMPUTE else "cached-hamming-distance"


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  It is Rust and cannot be built in this
    image (no cargo/rustc), so this times the oracle port with every host thread, on the same time-targeted sample
    of the workload's stream as the cpu_baseline leg of the GPU arm."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import sgdx_loader
    from oracle import oracle as orc
    sg = sgdx_loader.load()  # Python side only (the Workload parameters); libsgdx.so is never loaded here
    w = sg.synth.CONFIGS[args.config]
    table = orc.synth_table(w)
    threads = os.cpu_count() or 1
    n = args.reads or (w.n_reads // max(args.gpus, 1) if args.config == "C5" else min(w.n_reads, 100_000_000))
    if args.ref_sample:
        sample, reads = int(args.ref_sample), orc.synth_reads(w, 0, int(args.ref_sample), table, threads)
    else:
        sample, reads = cpu_sample(orc, w, table, threads, target_s=min(12.0, 120.0 / max(args.steps + args.warmup, 1)))
    for _ in range(args.warmup):
        cpu_port_run(orc, w, table, reads, threads)
    secs = [cpu_port_run(orc, w, table, reads, threads) for _ in range(args.steps)]
    total = sum(secs)
    value = sample * args.steps / total
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
          "sample": f"first {sample} reads of the workload's seeded stream per step (time-targeted bounded sample), "
                    f"oracle port with the reference's LRU memo"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_desc(w, n), "reads_per_gpu_per_step": n, "matcher": orc_kind_name(orc, w),
                   "sample_reads_per_step": sample},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Rust (no cargo/rustc in the image): CPU port of its algorithm, all host threads",
    }))
    return 0


# ------------------------------------------------------------------------------------------------
def digest_of(per_sample, hops):
    h = hashlib.sha256()
    h.update(per_sample.astype("<u8").tobytes())
    for k in sorted(hops):
        h.update(k)
        h.update(int(hops[k]).to_bytes(8, "little"))
    return h.hexdigest()


def bind_to_gpu_numa(torch, local):
    """Run this rank's host threads (and so the first touch of its pinned buffers) on the GPU's NUMA node."""
    try:
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{dev}"
        node = int(open(f"{base}/numa_node").read())
        cpus = open(f"{base}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"pci": dev, "numa_node": node, "cpus": cpus, "bound": bool(ids)}
    except Exception as ex:
        return {"bound": False, "why": repr(ex)[:120]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"],
                    help="BASELINE.json configs; C5 = C2's table, 2 G reads split over the ranks")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (default: the config's size, capped)")
    ap.add_argument("--impl", default="sgdx", choices=["sgdx", "reference"])
    ap.add_argument("--kernel", type=int, default=0, help="ASCII kernels: 0 auto, 1 direct, 2 seeded")
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-post", action="store_true", help="skip the side measurements (ascii, stress, post stages)")
    ap.add_argument("--e2e-chunk", type=int, default=4_000_000)
    ap.add_argument("--write-digest", action="store_true",
                    help="1 GPU: record the counters digest of this (config, total reads) in tests/golden/bench_digests.json")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    # rank 0 prints exactly one line on stdout; whatever libraries print (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    import sgdx_loader
    sg = sgdx_loader.load()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sgdx has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = sg.synth.CONFIGS[args.config]
    n = args.reads or (w.n_reads // world if args.config == "C5" else min(w.n_reads, 100_000_000))
    L, S = w.barcode_len, w.num_samples
    compact_ok = L <= 21
    table = w.table()
    kind = sg.select_matcher(L)
    cls = sg.PreComputeMatcher if kind == sg.MatcherKind.PreCompute else sg.CachedHammingDistanceMatcher
    samples = [sg.SampleMetadata(f"s{i}", bytes(table[i]), i) for i in range(S)]
    samples.append(sg.SampleMetadata(sg.UNDETERMINED_NAME, b"N" * L, S))
    stage = sg.BarcodeAssignmentStage(samples, cls, w.max_mismatches, w.min_delta, w.free_ns, w.max_no_calls,
                                      w.index_lengths, device=local, slots=4, kernel=args.kernel)
    m = stage.matcher
    side = not args.no_post and args.config != "C5"   # side measurements need the ASCII copy of the reads

    # ---- inputs resident in HBM: this rank's shard of the seeded stream, generated on the device (piecewise, so that
    # a 2 G-read shard never needs its 40 GB of ASCII at once) and packed into compact records by the device packer
    stream = torch.cuda.Stream()  # a real (non-default) stream: handle 0 would mean "the slot's own stream"
    torch.cuda.set_stream(stream)
    d_table = torch.from_numpy(table).cuda()
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    m.set_stream(0, stream.cuda_stream)  # launches go on torch's current stream so torch events see them
    d_rec = torch.empty(n, dtype=torch.int64, device="cuda") if compact_ok else torch.empty((n, 4), dtype=torch.int32, device="cuda")
    keep_ascii = side
    d_in = torch.empty((n, L), dtype=torch.uint8, device="cuda") if keep_ascii else None
    piece = 100_000_000
    scratch = None if keep_ascii else torch.empty((min(piece, n), L), dtype=torch.uint8, device="cuda")
    for off in range(0, n, piece):
        c = min(piece, n - off)
        buf = d_in[off:off + c] if keep_ascii else scratch[:c]
        w.reads_device(rank * n + off, c, d_table.data_ptr(), buf.data_ptr(), stream.cuda_stream)
        if compact_ok:
            m.pack_compact_device(buf.data_ptr(), c, d_rec.data_ptr() + 8 * off, slot=0)
        else:
            m.pack_device(buf.data_ptr(), c, d_rec.data_ptr() + 16 * off, slot=0)
    torch.cuda.synchronize()
    scratch = None

    if compact_ok:
        kname = m.compact_kernel_name()
        in_bytes, in_desc = 8, "compact 64-bit records (2-bit base planes + invalid plane, include/sgdx.h) resident in HBM"

        def step():
            m.match_compact_device(d_rec.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr(), slot=0)
    else:
        kname = m.packed_kernel_name()
        in_bytes, in_desc = 16, ("16-byte sgdx_read records (2-bit base planes + no-call and foreign-byte masks, include/sgdx.h) "
                                 "resident in HBM; 12 of the 16 bytes are algorithmic (SURVEY 8d)")

        def step():
            m.match_packed_device(d_rec.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr(), slot=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, clocks=False):
        """W warm-ups, K launches bracketed by barrier + synchronize, CUDA events on the launching stream."""
        for _ in range(max(warmup, 3)):
            fn()
        barrier()
        l0 = m.launch_count()
        sampler = ClockSampler(local) if clocks and rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        evs[0].record(stream)
        for k in range(steps):
            fn()
            evs[k + 1].record(stream)
        barrier()
        per = [evs[k].elapsed_time(evs[k + 1]) for k in range(steps)]
        total = evs[0].elapsed_time(evs[-1])
        return per, total, (sampler.stop() if sampler else None), m.launch_count() - l0

    m.reset_counters()
    per_launch_ms, total_ms, clocks, launches = timed(step, args.steps, args.warmup, clocks=True)
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n * args.steps / (total_ms_max * 1e-3)

    # ---- the counted work: one verification pass on fresh counters.  Every read must have been assigned (nothing
    # skipped or cached); per-GPU counters are summed on the host (no device collective on this path) and, where the
    # single-GPU result of the same stream is on record, the merged per-sample counters and hop map must equal it.
    m.reset_counters()
    step()
    torch.cuda.synchronize()
    per, totals = m.counters()
    assert totals["total_templates"] == n, (totals, n)
    assert int(per[:, 0].sum()) == n
    if n <= 500_000_000:
        bins = torch.bincount(d_idx.to(torch.int64) & 0xFFFF, minlength=S + 1)
        assert np.array_equal(bins.cpu().numpy(), per[:, 0].astype(np.int64)), "index array and counters disagree"
        del bins
    merged = sg.sharding.gather_metrics(stage.metrics(), dst=0)
    matched_frac, merged_check = None, None
    if rank == 0:
        assert merged.total_templates == world * n
        mp_ = merged.per_sample_array()
        matched_frac = float(mp_[:-1, 0].sum()) / max(merged.total_templates, 1)
        hops = merged.sample_barcode_hop_tracker.count_tracker if merged.sample_barcode_hop_tracker else {}
        dg = digest_of(mp_, hops)
        key = f"{w.table_seed:#x}:{w.read_seed:#x}:S{S}:L{L}:M{w.max_mismatches}:D{w.min_delta}:first0:reads{world * n}"
        known = {}
        try:
            known = json.load(open(DIGESTS))
        except Exception:
            pass
        merged_check = {"digest": dg, "key": key, "hop_keys": len(hops), "hop_reads": int(sum(hops.values())),
                        "single_gpu_digest": known.get(key),
                        "equals_single_gpu": (known[key] == dg) if key in known else None}
        if key in known and world > 1:
            assert known[key] == dg, f"merged counters / hop map of {world} shards differ from the single-GPU result ({key})"
        if args.write_digest and world == 1:
            known[key] = dg
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            for path in (DIGESTS, os.path.join(ROOT, "gpurun_out", "bench_digests.json")):
                json.dump(known, open(path, "w"), indent=1, sort_keys=True)

    # ---- roofline of the dominant (only) kernel: HBM.  achieved = SURVEY 8(d)'s algorithmic bytes per read x the
    # reads of one launch / the launch's CUDA-event duration; traffic = ncu dram bytes of the same launch (profiles/)
    peaks, peaks_src = {}, "measured (MEASURED_PEAKS.json)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    def roofline_of(kernel, avg_ms, reads, buffer_bytes):
        achieved = ALG_BYTES(L) * reads / (avg_ms * 1e-3) / 1e9
        buffers = buffer_bytes * reads / (avg_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel, {}).get(args.config if args.config != "C5" else "C2")
            if tr:
                traffic, traffic_src = tr["dram_bytes_per_read"] * reads, tr["source"]
        except Exception:
            pass
        return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks_src, "kernel": kernel,
                "algorithmic_bytes_per_read": ALG_BYTES(L), "buffer_bytes_per_read": buffer_bytes,
                "achieved_buffer_gbs": buffers, "frac_buffer": buffers / hbm_peak, "avg_launch_ms": avg_ms}

    roofline = roofline_of(kname, sum(per_launch_ms) / len(per_launch_ms), n, in_bytes + 3)
    if rank == 0 and kname == "sgdx_match_direct":
        popc, lop3 = C.c_double(0), C.c_double(0)
        sg.lib().sgdx_measure_int_peak(local, C.byref(popc), C.byref(lop3))
        pairs_per_s = n * S / (roofline["avg_launch_ms"] * 1e-3)
        roofline["integer"] = {
            "note": "the direct kernel scores every (read, sample) pair: one POPC each, so it is bound by the POPC "
                    "rate (measured on this box by sgdx_measure_int_peak), not by HBM",
            "pairs_per_s": pairs_per_s, "popc_peak_per_s": popc.value, "lop3_peak_per_s": lop3.value,
            "frac_of_popc_peak": pairs_per_s / popc.value if popc.value else None}

    # ---- end to end through the reference-facing C-ABI call: pinned host buffers, H2D + kernel + D2H per step, in
    # chunks over the handle's slots; and the bare copies of the same bytes as the ceiling of this host/box
    lib = sg.lib()
    nslots = m.slots
    chunk = min(args.e2e_chunk, n)

    def alloc_host(nbytes):
        p = C.c_void_p()
        sg._lib.check(lib.sgdx_alloc_host(max(nbytes, 1), C.byref(p)))
        return p

    def e2e_measure(submit, d_src, row_bytes, np_dtype, steps):
        """submit(in_ptr, count, idx_ptr, dist_ptr, slot).  Returns (reads/s over all ranks, h2d bytes, d2h bytes, ok)."""
        hin, hidx, hdist = alloc_host(n * row_bytes), alloc_host(n * 2), alloc_host(n)
        h_in = np.ctypeslib.as_array(C.cast(hin, C.POINTER(np.ctypeslib.as_ctypes_type(np_dtype))), d_src.shape)
        h_idx = np.ctypeslib.as_array(C.cast(hidx, C.POINTER(C.c_uint16)), (n,))
        torch.from_numpy(h_in).copy_(d_src)  # the same reads, now in pinned host memory (untimed)
        torch.cuda.synchronize()

        def one_pass():
            off, i = 0, 0
            while off < n:
                c = min(chunk, n - off)
                slot = i % nslots
                if i >= nslots:
                    m.sync(slot)
                submit(hin.value + off * row_bytes, c, hidx.value + off * 2, hdist.value + off, slot)
                off += c
                i += 1
            for s in range(nslots):
                m.sync(s)

        one_pass()  # warm-up (allocates the slots' device staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_pass()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        k = min(n, 1 << 20)
        ok = bool(np.array_equal(h_idx[:k], d_idx[:k].cpu().numpy().view(np.uint16)))
        for p in (hin, hidx, hdist):
            lib.sgdx_free_host(p)
        return world * n * steps / float(t2.item()), n * row_bytes, n * 3, ok

    def copy_ceiling(row_bytes, steps):
        """The same chunks, bytes and stream count as e2e_measure, copies only (no kernel): what this host can move.
        row_bytes may be fractional (dense records + their exception list): chunks copy round(c * row_bytes) bytes."""
        up_total = int(n * row_bytes) + 64
        hin, hout = alloc_host(up_total), alloc_host(n * 3)
        h_in = torch.from_numpy(np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_uint8)), (up_total,)))
        h_out = torch.from_numpy(np.ctypeslib.as_array(C.cast(hout, C.POINTER(C.c_uint8)), (n * 3,)))
        streams = [torch.cuda.Stream() for _ in range(nslots)]
        d_up = [torch.empty(int(chunk * row_bytes) + 64, dtype=torch.uint8, device="cuda") for _ in range(nslots)]
        d_dn = [torch.empty(chunk * 3, dtype=torch.uint8, device="cuda") for _ in range(nslots)]

        def one_pass():
            off, i = 0, 0
            while off < n:
                c = min(chunk, n - off)
                s = i % nslots
                with torch.cuda.stream(streams[s]):
                    a, b = int(off * row_bytes), int((off + c) * row_bytes)
                    d_up[s][: b - a].copy_(h_in[a:b], non_blocking=True)
                    h_out[off * 3:(off + c) * 3].copy_(d_dn[s][: c * 3], non_blocking=True)
                off += c
                i += 1
            torch.cuda.synchronize()

        one_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_pass()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        for p in (hin, hout):
            lib.sgdx_free_host(p)
        secs = float(t2.item()) / steps
        return {"reads_per_s": world * n / secs, "h2d_gbs": world * n * row_bytes / secs / 1e9,
                "d2h_gbs": world * n * 3 / secs / 1e9,
                "what": f"bare cudaMemcpyAsync of the same {row_bytes:.3g}+3 bytes per read, same chunks and {nslots} streams, "
                        f"all {world} rank(s) at once, no kernel"}

    def e2e_dense(steps):
        """Dense records (2 bits per base + exception list) in pinned host memory through sgdx_match_dense_batch; the
        chunks are packed beforehand (untimed), like the records of the other e2e legs."""
        B = sg.dense_bytes(L)
        host_ascii = d_in.cpu().numpy()
        hden, hidx, hdist = alloc_host(n * B + 64), alloc_host(n * 2), alloc_host(n)
        h_den = np.ctypeslib.as_array(C.cast(hden, C.POINTER(C.c_uint8)), (n * B + 64,))
        h_idx = np.ctypeslib.as_array(C.cast(hidx, C.POINTER(C.c_uint16)), (n,))
        chunks, eis, ers = [], [], []
        off = 0
        while off < n:
            c = min(chunk, n - off)
            dn, ei, er = sg.pack_dense_host(host_ascii[off:off + c], threads=os.cpu_count() or 1)
            h_den[off * B:(off + c) * B] = dn
            chunks.append((off, c, sum(len(x) for x in eis), len(ei)))
            eis.append(ei)
            ers.append(er)
            off += c
        n_exc = sum(len(x) for x in eis)
        esz = 8 if L <= 21 else 16  # compact records, or sgdx_read records above 21 bases
        hei, her = alloc_host(n_exc * 4 + 8), alloc_host(n_exc * esz + 8)
        if n_exc:
            np.ctypeslib.as_array(C.cast(hei, C.POINTER(C.c_uint32)), (n_exc,))[:] = np.concatenate(eis)
            np.ctypeslib.as_array(C.cast(her, C.POINTER(C.c_uint8)), (n_exc * esz,))[:] = np.concatenate(
                [np.ascontiguousarray(x).view(np.uint8).reshape(-1) for x in ers])
        del host_ascii, eis, ers

        def one_pass():
            for i, (o, c, e0, ne) in enumerate(chunks):
                slot = i % nslots
                if i >= nslots:
                    m.sync(slot)
                m.submit_dense_ptr(hden.value + o * B, c, hei.value + 4 * e0, her.value + esz * e0, ne, hidx.value + o * 2,
                                   hdist.value + o, slot)
            for s_ in range(nslots):
                m.sync(s_)

        one_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_pass()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        k = min(n, 1 << 20)
        ok = bool(np.array_equal(h_idx[:k], d_idx[:k].cpu().numpy().view(np.uint16)))
        for p_ in (hden, hidx, hdist, hei, her):
            lib.sgdx_free_host(p_)
        up = n * B + n_exc * (4 + esz)
        return world * n * steps / float(t2.item()), up, n * 3, ok, n_exc

    e2e, e2e_ascii = None, None
    if not args.no_e2e and args.config != "C5":
        m.set_stream(0, 0)
        e2e_steps = max(1, min(args.steps, 5))
        row = 8 if compact_ok else 16
        v, up, down, ok = e2e_measure(m.submit_compact_ptr if compact_ok else m.submit_packed_ptr, d_rec, row,
                                      np.int64 if compact_ok else np.int32, e2e_steps)
        ceil = copy_ceiling(row, e2e_steps)
        call = "sgdx_match_compact_batch" if compact_ok else "sgdx_match_packed_batch"
        e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": up, "d2h_bytes_per_step": down, "steps": e2e_steps,
               "chunk_reads": chunk, "slots": nslots, "call": call,
               "input": f"{row}-byte records in pinned host memory (written by the batching stage instead of its ASCII copy)",
               "timing": f"wall clock around {call}(host buffers)+sgdx_sync, max over ranks",
               "matches_device_resident_result": ok, "copy_ceiling": ceil,
               "frac_of_copy_ceiling": v / ceil["reads_per_s"], "numa": numa}
        if L <= 32 and keep_ascii and side:
            # the PCIe form of the same records: 2 bits per base + an exception list for the reads with a no-call
            vd, upd, downd, okd, n_exc = e2e_dense(e2e_steps)
            ceild = copy_ceiling(upd / n, e2e_steps)
            e2e = {"value": vd, "unit": UNIT, "h2d_bytes_per_step": upd, "d2h_bytes_per_step": downd, "steps": e2e_steps,
                   "chunk_reads": chunk, "slots": nslots, "call": "sgdx_match_dense_batch",
                   "input": f"dense records in pinned host memory: {sg.dense_bytes(L)} bytes per read (2 bits per base) + "
                            f"{n_exc} exceptions of {12 if L <= 21 else 20} bytes (reads with a no-call or a foreign byte), include/sgdx.h",
                   "timing": "wall clock around sgdx_match_dense_batch(host buffers)+sgdx_sync, max over ranks",
                   "matches_device_resident_result": okd, "copy_ceiling": ceild,
                   "frac_of_copy_ceiling": vd / ceild["reads_per_s"], "numa": numa,
                   "compact_records" if compact_ok else "packed_records": e2e}
        if keep_ascii and side:
            v, up, down, ok = e2e_measure(m.submit_ptr, d_in, L, np.uint8, e2e_steps)
            e2e_ascii = {"value": v, "unit": UNIT, "h2d_bytes_per_step": up, "d2h_bytes_per_step": down,
                         "steps": e2e_steps, "chunk_reads": chunk, "slots": nslots, "call": "sgdx_match_batch",
                         "matches_device_resident_result": ok}
        m.set_stream(0, stream.cuda_stream)

    # ---- beside the headline: the ASCII entry point on the same reads, and the same kernel on unfriendly streams
    ascii_side, stress = None, None
    if side:
        ref_idx, ref_dist = d_idx.clone(), d_dist.clone()
        per_a, _, _, _ = timed(lambda: m.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr(), slot=0),
                            min(args.steps, 10), args.warmup)
        same = bool(torch.equal(ref_idx, d_idx) and torch.equal(ref_dist, d_dist))
        ascii_side = {"device": roofline_of(m.kernel_name(), sum(per_a) / len(per_a), n, L + 3),
                      "reads_per_s": world * n / (sum(per_a) / len(per_a) * 1e-3),
                      "equals_compact_result": same, "e2e": e2e_ascii}
        del ref_idx, ref_dist
        stress = {}
        import dataclasses
        for name, ww in (("p_sub_5pct", dataclasses.replace(w, p_sub=0.05)),
                         ("one_mismatch_everywhere", dataclasses.replace(w, p_true=1.0, p_hop=0.0, p_sub=0.0, p_n=0.0))):
            ww.reads_device(rank * n, n, d_table.data_ptr(), d_in.data_ptr(), stream.cuda_stream)
            if name == "one_mismatch_everywhere":  # exactly one substituted base in every read
                g = torch.Generator(device="cuda").manual_seed(1234 + rank)
                pos = torch.randint(0, L, (n,), device="cuda", generator=g)
                rot = torch.randint(1, 4, (n,), device="cuda", generator=g)
                lut = torch.zeros(256, dtype=torch.int64, device="cuda")
                lut[list(b"ACGT")] = torch.arange(4, device="cuda")
                acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
                rows = torch.arange(n, device="cuda")
                cur = lut[d_in[rows, pos].to(torch.int64)]
                d_in[rows, pos] = acgt[(cur + rot) & 3]
                del pos, rot, rows, cur
            if compact_ok:
                m.pack_compact_device(d_in.data_ptr(), n, d_rec.data_ptr(), slot=0)
            else:
                m.pack_device(d_in.data_ptr(), n, d_rec.data_ptr(), slot=0)
            m.reset_counters()
            per_s, _, _, _ = timed(step, min(args.steps, 10), args.warmup)
            pc, tt = m.counters()
            avg = sum(per_s) / len(per_s)
            stress[name] = {"avg_launch_ms": avg, "reads_per_s": world * n / (avg * 1e-3),
                            "frac": ALG_BYTES(L) * n / (avg * 1e-3) / 1e9 / hbm_peak,
                            "matched_fraction": float(pc[:-1, 0].sum()) / max(tt["total_templates"], 1),
                            "one_mismatch_fraction": float(pc[:-1, 2].sum()) / max(tt["total_templates"], 1)}
        stress["note"] = ("same table and kernel as the headline; parity of both streams against the oracle: "
                          "tests/test_gpu_ph.py::test_stress_streams_equal_the_oracle")

    # ---- config 4's most_frequent_unmatched on the fast path: the unmatched table fed from the compact
    # records by the same launch sequence, top keys read back; parity of the counts: tests/test_gpu_post.py
    unmatched_fast = None
    if rank == 0 and world == 1 and side and compact_ok and args.config in ("C4", "C1"):
        st2 = sg.BarcodeAssignmentStage(samples, cls, w.max_mismatches, w.min_delta, w.free_ns, w.max_no_calls, w.index_lengths,
                                        device=local, slots=1, unmatched_on_device=True)
        m2 = st2.matcher
        m2.set_stream(0, stream.cuda_stream)
        w.reads_device(rank * n, n, d_table.data_ptr(), d_in.data_ptr(), stream.cuda_stream)
        m.pack_compact_device(d_in.data_ptr(), n, d_rec.data_ptr(), slot=0)
        torch.cuda.synchronize()

        def step_um():
            m2.match_compact_device(d_rec.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr(), slot=0)
        m_saved = m
        m = m2  # (timed() counts launches on m)
        per_u, _, _, _ = timed(step_um, min(args.steps, 5), args.warmup)
        m = m_saved
        info = m2.unmatched_info()
        top = m2.unmatched_top(5)
        runs = min(args.steps, 5) + max(args.warmup, 3)
        unmatched_fast = {"kernel": m2.compact_kernel_name() + " + sgdx_unmatched_count (keys from the compact records)",
                          "avg_ms": sum(per_u) / len(per_u), "reads_per_s": n / (sum(per_u) / len(per_u) * 1e-3),
                          "unmatched_reads_per_pass": info["unmatched_reads"] // runs, "distinct_keys": info["distinct_keys"],
                          "dropped_reads": info["dropped_reads"],
                          "top": [[k.decode(), c // runs] for k, c in top]}
        st2.close()

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_baseline(w)

    # ---- the stages either side of the assignment, device-timed on C2-shaped reads: not part of
    # `value`; their kernels are not counted in gpu_launches
    post = None
    if rank == 0 and world == 1 and side and args.config == "C2":
        try:
            d_in = d_rec = d_idx = d_dist = None
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_post
            post = bench_post.measure(reads=min(n, 50_000_000), steps=min(args.steps, 10), warmup=max(args.warmup, 3))
        except Exception as ex:  # never lose the bench line to a side measurement
            post = {"error": repr(ex)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_desc(w, n), "reads_per_gpu_per_step": n, "matcher": kind.value,
                       "kernel": kname, "input": in_desc,
                       "sharding": f"{world} shard(s) by read range, host-merged counters",
                       "l2": f"inputs {n * in_bytes / 1e6:.0f} MB + outputs {n * 3 / 1e6:.0f} MB per step, larger than the 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "matched_fraction": matched_frac, "merged_check": merged_check,
            "ascii": ascii_side, "stress": stress, "unmatched_fast_path": unmatched_fast, "post": post,
        }
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    stage.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
