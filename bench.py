#!/usr/bin/env python3
"""bench.py -- headline benchmark of the sample-barcode assignment path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

One "step" = one pass of the assignment hot path (N gate + match + per-sample/hop counters) over one
batch of synthetic reads of the named config, per GPU.  Weak scaling: every rank owns its own shard
(rank r holds reads [r*n, (r+1)*n) of the seeded stream), no data-path collective; counters are merged
on the host.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
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


def cpu_port_run(orc, w, table, reads, threads):
    kind = orc.select_matcher(w.barcode_len)
    cfg = orc.Config(w.num_samples, w.barcode_len, w.max_mismatches, w.min_delta, w.free_ns,
                     -1 if w.max_no_calls is None else w.max_no_calls, kind, w.index1_len,
                     w.barcode_len - w.index1_len if w.index1_len else 0)
    o = orc.COracle(cfg, table, native=True)
    res = o.run(reads, threads=threads, use_cache=True, want_assignments=True)
    return res["seconds"]


def cpu_baseline(sg, w, table, target_s=12.0):
    """The oracle port (reference algorithm incl. its per-thread 100k LRU memo, matcher.rs:307-328), all host
    cores, on a bounded sample of the same seeded stream."""
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    probe = 1_000_000
    secs = cpu_port_run(orc, w, table, w.reads_host(0, probe, table, threads), threads)
    n = int(min(max(probe, probe * target_s / max(secs, 1e-3)), 64_000_000))
    reads = w.reads_host(0, n, table, threads)
    secs = cpu_port_run(orc, w, table, reads, threads)
    return {"value": n / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} reads of the same seeded stream, {secs:.2f} s wall, oracle/sgdx_oracle.c "
                      f"-O3 -march=native, per-thread 100k-entry LRU memo as in the reference"}


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  It is Rust and cannot be built in
    this image (no cargo/rustc), so this times the oracle port with every host thread."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import sgdx_loader
    from oracle import oracle as orc
    sg = sgdx_loader.load()
    w = sg.synth.CONFIGS[args.config]
    table = w.table()
    threads = os.cpu_count() or 1
    n = args.reads or 100_000_000
    sample = int(args.ref_sample or min(n, 8_000_000))
    reads = w.reads_host(0, sample, table, threads)
    for _ in range(args.warmup):
        cpu_port_run(orc, w, table, reads[: max(sample // 8, 1)], threads)
    secs = [cpu_port_run(orc, w, table, reads, threads) for _ in range(args.steps)]
    total = sum(secs)
    value = sample * args.steps / total
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
          "sample": f"{sample} reads per step (bounded sample of the {n}-read workload), oracle port with LRU memo"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_desc(w, n), "sample_reads_per_step": sample},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Rust (no cargo/rustc in the image): CPU port of its algorithm, all host threads",
    }))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"],
                    help="BASELINE.json configs; C5 = C2's table, 2 G reads split over the ranks (40 GB on one GPU)")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (default: the config's size, capped)")
    ap.add_argument("--impl", default="sgdx", choices=["sgdx", "reference"])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 direct, 2 seeded")
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-post", action="store_true", help="skip the stages either side of the assignment")
    ap.add_argument("--e2e-chunk", type=int, default=4_000_000)
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w = sg.synth.CONFIGS[args.config]
    n = args.reads or (w.n_reads // world if args.config == "C5" else min(w.n_reads, 100_000_000))
    L, S = w.barcode_len, w.num_samples
    table = w.table()
    kind = sg.select_matcher(L)
    cls = sg.PreComputeMatcher if kind == sg.MatcherKind.PreCompute else sg.CachedHammingDistanceMatcher
    samples = [sg.SampleMetadata(f"s{i}", bytes(table[i]), i) for i in range(S)]
    samples.append(sg.SampleMetadata(sg.UNDETERMINED_NAME, b"N" * L, S))
    stage = sg.BarcodeAssignmentStage(samples, cls, w.max_mismatches, w.min_delta, w.free_ns, w.max_no_calls,
                                      w.index_lengths, device=local, slots=4, kernel=args.kernel)
    m = stage.matcher

    # ---- inputs resident in HBM: this rank's shard of the seeded stream, generated on the device
    stream = torch.cuda.Stream()  # a real (non-default) stream: handle 0 would mean "the slot's own stream"
    torch.cuda.set_stream(stream)
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, L), dtype=torch.uint8, device="cuda")
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    w.reads_device(rank * n, n, d_table.data_ptr(), d_in.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    m.set_stream(0, stream.cuda_stream)  # launches go on torch's current stream so torch events see them

    def step():
        m.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr(), slot=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    m.reset_counters()
    launches0 = m.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record(stream)
    for k in range(args.steps):
        step()
        evs[k + 1].record(stream)
    barrier()
    per_launch_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    total_ms = evs[0].elapsed_time(evs[-1])
    launches = m.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n * args.steps / (total_ms_max * 1e-3)

    # sanity on the counted work: every timed read must have been assigned (nothing skipped or cached);
    # per-GPU counters are summed on the host (no device collective on this path)
    per, totals = m.counters()
    assert totals["total_templates"] == n * args.steps, (totals, n, args.steps)
    assert int(per[:, 0].sum()) == n * args.steps
    merged = sg.sharding.gather_metrics(stage.metrics(), dst=0)
    matched_frac = None
    if rank == 0:
        assert merged.total_templates == world * n * args.steps
        mp_ = merged.per_sample_array()
        matched_frac = float(mp_[:-1, 0].sum()) / max(merged.total_templates, 1)

    # ---- roofline of the dominant (only) kernel: HBM.  achieved = SURVEY 8(d)'s algorithmic bytes per read x the
    # reads of one launch / the launch's CUDA-event duration; the ABI takes ASCII, so the buffers actually moved
    # are L+3 bytes per read (reported beside it); traffic = ncu dram bytes of the same launch (profiles/)
    peaks = {}
    peaks_src = "measured (MEASURED_PEAKS.json)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks_src = "fallback (B200_PROFILING.md 6.65 TB/s)"
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    avg_ms = sum(per_launch_ms) / len(per_launch_ms)
    kname = m.kernel_name()
    alg_bytes = ALG_BYTES(L) * n
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    buffers = (L + 3) * n / (avg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kname, {}).get(args.config)
        if t:
            traffic, traffic_src = t["dram_bytes_per_read"] * n, t["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks_src, "kernel": kname,
                "algorithmic_bytes_per_read": ALG_BYTES(L), "buffer_bytes_per_read": L + 3,
                "achieved_buffer_gbs": buffers, "frac_buffer": buffers / hbm_peak, "avg_launch_ms": avg_ms}
    if rank == 0 and kname == "sgdx_match_direct":
        popc, lop3 = C.c_double(0), C.c_double(0)
        sg.lib().sgdx_measure_int_peak(local, C.byref(popc), C.byref(lop3))
        pairs_per_s = n * S / (avg_ms * 1e-3)
        roofline["integer"] = {
            "note": "the direct kernel scores every (read, sample) pair: one POPC each, so it is bound by the POPC "
                    "rate (measured on this box by sgdx_measure_int_peak), not by HBM",
            "pairs_per_s": pairs_per_s, "popc_peak_per_s": popc.value, "lop3_peak_per_s": lop3.value,
            "frac_of_popc_peak": pairs_per_s / popc.value if popc.value else None}

    # ---- end to end through the reference-facing C-ABI call: pinned host buffers, H2D + kernel + D2H per step
    e2e = None
    if not args.no_e2e:
        m.set_stream(0, 0)
        lib = sg.lib()
        hin, hidx, hdist = C.c_void_p(), C.c_void_p(), C.c_void_p()
        sg._lib.check(lib.sgdx_alloc_host(n * L, C.byref(hin)))
        sg._lib.check(lib.sgdx_alloc_host(n * 2, C.byref(hidx)))
        sg._lib.check(lib.sgdx_alloc_host(n, C.byref(hdist)))
        h_in = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_uint8)), (n, L))
        h_idx = np.ctypeslib.as_array(C.cast(hidx, C.POINTER(C.c_uint16)), (n,))
        torch.from_numpy(h_in).copy_(d_in)  # the same reads, now in pinned host memory (untimed)
        torch.cuda.synchronize()
        chunk = min(args.e2e_chunk, n)
        nslots = m.slots

        def e2e_pass():
            off, i = 0, 0
            while off < n:
                c = min(chunk, n - off)
                slot = i % nslots
                if i >= nslots:
                    m.sync(slot)
                m.submit_ptr(hin.value + off * L, c, hidx.value + off * 2, hdist.value + off, slot)
                off += c
                i += 1
            for s in range(nslots):
                m.sync(s)

        e2e_pass()  # warm-up (allocates the slots' device staging)
        m.reset_counters()
        e2e_steps = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_pass()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ok = bool(np.array_equal(h_idx[: 1 << 20], d_idx[: 1 << 20].cpu().numpy().view(np.uint16)))
        e2e = {"value": world * n * e2e_steps / float(t2.item()), "unit": UNIT, "h2d_bytes_per_step": n * L,
               "d2h_bytes_per_step": n * 3, "steps": e2e_steps, "chunk_reads": chunk, "slots": nslots,
               "timing": "wall clock around sgdx_match_batch(host buffers)+sgdx_sync, max over ranks",
               "matches_device_resident_result": ok}
        for p in (hin, hidx, hdist):
            lib.sgdx_free_host(p)

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_baseline(sg, w, table)

    # ---- the stages either side of the assignment, device-timed on the same C2-shaped reads: not part
    # of `value`; their kernels are not counted in gpu_launches
    post = None
    if rank == 0 and world == 1 and not args.no_post and args.config == "C2":
        del d_in, d_idx, d_dist
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_post
        post = bench_post.measure(reads=min(n, 50_000_000), steps=min(args.steps, 10), warmup=max(args.warmup, 3))

    # ---- the compact-record input format next to the ASCII path, same workload: kernel time on
    # resident records and the end-to-end call from host records; reported beside `value` / `e2e`, not instead
    compact = None
    if rank == 0 and world == 1 and not args.no_post and L <= 21:
        try:
            d_in = d_idx = d_dist = None
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_compact
            compact = bench_compact.measure(config=args.config, reads=n, steps=min(args.steps, 10),
                                            warmup=max(args.warmup, 3), no_e2e=args.no_e2e)
        except Exception as ex:  # never lose the bench line to the side measurement
            import traceback
            traceback.print_exc()
            compact = {"error": repr(ex)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_desc(w, n), "reads_per_gpu_per_step": n, "matcher": kind.value,
                       "kernel": kname,
                       "input": "ASCII barcodes resident in HBM", "sharding": f"{world} shard(s) by read range, host-merged counters",
                       "l2": f"inputs {n * L / 1e6:.0f} MB + outputs {n * 3 / 1e6:.0f} MB per step, larger than the 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "matched_fraction": matched_frac, "post": post, "compact": compact,
        }
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    stage.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
