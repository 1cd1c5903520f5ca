#!/usr/bin/env python3
"""Device-timed measurements of the stages either side of the assignment on one B200:
base-quality counters, unmatched-barcode counting, segment extraction, read routing.

  python tools/bench_post.py [--reads N] [--steps K] [--warmup W] [--only qual,unmatched,extract,route]

CUDA events on the launching stream, >= 3 warm-ups, inputs larger than the 126 MB L2.  Prints one JSON object
(stage -> numbers); the roofline of every stage is HBM, `bytes` = the algorithmic bytes one launch must move.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(reads=50_000_000, steps=10, warmup=3, only="qual,qual_index,unmatched,extract,route,demultiplex", stride=150):
    """Returns {stage: numbers}; used by this script and by bench.py's "post" key."""
    import argparse as _a
    args = _a.Namespace(reads=reads, steps=steps, warmup=warmup, stride=stride)
    only = set(only.split(","))

    import numpy as np
    import torch
    import sgdx_loader
    sg = sgdx_loader.load()
    RS = sg.ReadStructure.from_str
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]

    w = sg.synth.C2
    n, L, S = args.reads, w.barcode_len, w.num_samples
    table = w.table()
    samples = [sg.SampleMetadata(f"s{i}", bytes(table[i]), i) for i in range(S)]
    samples.append(sg.SampleMetadata(sg.UNDETERMINED_NAME, b"N" * L, S))
    prev_stream = torch.cuda.current_stream()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, L), dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), stream.cuda_stream)
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    out = {"workload": f"C2 384x(10+10), {n} reads", "hbm_peak_gbs": peak, "steps": args.steps, "warmup": args.warmup}

    def timed(fn, before=None):
        for _ in range(max(args.warmup, 3)):
            if before:
                before()
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.steps):
            if before:
                before()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return sum(ms) / len(ms), min(ms)

    def line(name, ms, best, nbytes, kernel, extra=None):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"kernel": kernel, "avg_ms": ms, "best_ms": best, "reads_per_s": n / (ms * 1e-3),
                     "bytes": nbytes, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
        if extra:
            out[name].update(extra)

    stage = sg.BarcodeAssignmentStage(samples, sg.CachedHammingDistanceMatcher, w.max_mismatches, w.min_delta,
                                      w.free_ns, w.max_no_calls, w.index_lengths, slots=2)
    m = stage.matcher
    m.set_stream(0, stream.cuda_stream)
    m.match_device(d_in.data_ptr(), n, d_idx.data_ptr())
    torch.cuda.synchronize()
    match_ms, _ = timed(lambda: m.match_device(d_in.data_ptr(), n, d_idx.data_ptr()))
    out["match_only_ms"] = match_ms

    if "qual" in only:
        stride = args.stride
        g = torch.Generator(device="cuda")
        g.manual_seed(1)
        d_q = torch.randint(33, 75, (n, stride), dtype=torch.uint8, device="cuda", generator=g)
        rs = RS(f"{stride}T")
        ms, best = timed(lambda: m.qual_counts_device(d_q.data_ptr(), n, stride, rs, d_idx.data_ptr()))
        line("qual_counts_template", ms, best, n * (stride + 2), "sgdx_qual_counts<false, false>",
             {"read_structure": f"{stride}T", "bytes_per_read": stride + 2})
        del d_q
    if "qual_index" in only:
        g = torch.Generator(device="cuda")
        g.manual_seed(2)
        d_q = torch.randint(33, 75, (n, 10), dtype=torch.uint8, device="cuda", generator=g)
        rs = RS("10B")
        ms, best = timed(lambda: m.qual_counts_device(d_q.data_ptr(), n, 10, rs, d_idx.data_ptr()))
        line("qual_counts_index", ms, best, n * 12, "sgdx_qual_counts<true, true> (rows <= 32 bytes)", {"read_structure": "10B", "bytes_per_read": 12})
        del d_q
    if "extract" in only:
        d_i1 = d_in[:, :10].contiguous()
        d_i2 = d_in[:, 10:].contiguous()
        d_out = torch.empty((n, L), dtype=torch.uint8, device="cuda")
        rss = [RS("10B"), RS("10B")]
        ms, best = timed(lambda: m.extract_device([d_i1.data_ptr(), d_i2.data_ptr()], [10, 10], rss, n, d_out.data_ptr()))
        assert bool(torch.equal(d_out, d_in))
        line("extract", ms, best, n * 2 * L, "sgdx_extract", {"bytes_per_read": 2 * L})
        del d_i1, d_i2, d_out
    if "route" in only:
        d_order = torch.empty(n, dtype=torch.int32, device="cuda")
        d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
        ms, best = timed(lambda: m.route_device(d_idx.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr()))
        line("route", ms, best, n * (2 + 2 + 4), "sgdx_route_hist_cta + scans + sgdx_route_scatter_cta",
             {"bytes_per_read": 8, "note": "index array read twice (histogram, scatter), 4-byte read id written once"})
        del d_order
    if "demultiplex" in only:
        # Demultiplex::demultiplex for a whole paired-end dual-index batch resident on the device: I1 10B, I2 10B,
        # R1 150T, R2 150T with qualities -> extraction, assignment, four quality passes, routing (one call)
        nd = min(n, 20_000_000)
        g = torch.Generator(device="cuda")
        g.manual_seed(3)
        d_i1 = d_in[:nd, :10].contiguous()
        d_i2 = d_in[:nd, 10:].contiguous()
        d_qi = torch.randint(33, 75, (nd, 10), dtype=torch.uint8, device="cuda", generator=g)
        d_r = torch.randint(33, 75, (nd, 150), dtype=torch.uint8, device="cuda", generator=g)  # stands for bases and quals
        d_order = torch.empty(nd, dtype=torch.int32, device="cuda")
        d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
        d_dist = torch.empty(nd, dtype=torch.uint8, device="cuda")
        fastqs = [(d_i1.data_ptr(), d_qi.data_ptr(), 0, 10), (d_i2.data_ptr(), d_qi.data_ptr(), 0, 10),
                  (d_r.data_ptr(), d_r.data_ptr(), 0, 150), (d_r.data_ptr(), d_r.data_ptr(), 0, 150)]
        rss = [RS("10B"), RS("10B"), RS("150T"), RS("150T")]
        ms, best = timed(lambda: m.demultiplex_device(fastqs, rss, nd, d_idx.data_ptr(), d_dist.data_ptr(),
                                                      d_order.data_ptr(), d_off.data_ptr()))
        nbytes = nd * (20 + 20 + 300 + 20 + 3 + 8)
        out["demultiplex_device"] = {
            "kernel": "sgdx_extract (barcodes written as compact records) + " + m.compact_kernel_name() + " + 4 x sgdx_qual_counts + sgdx_route_*", "templates": nd,
            "avg_ms": ms, "best_ms": best, "templates_per_s": nd / (ms * 1e-3), "bytes": nbytes,
            "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
            "layout": "I1 10B + I2 10B + R1 150T + R2 150T, qualities of all four FASTQs counted, routing on"}
        del d_i1, d_i2, d_qi, d_r, d_order, d_dist
    stage.close()
    if "unmatched" in only:
        st2 = sg.BarcodeAssignmentStage(samples, sg.CachedHammingDistanceMatcher, w.max_mismatches, w.min_delta,
                                        w.free_ns, w.max_no_calls, w.index_lengths, slots=2, unmatched_on_device=True,
                                        most_unmatched_max_map_size=max(16_000_000, n // 4))
        m2 = st2.matcher
        m2.set_stream(0, stream.cuda_stream)
        ms, best = timed(lambda: m2.match_device(d_in.data_ptr(), n, d_idx.data_ptr()), before=m2.reset_counters)
        info = m2.unmatched_info()
        extra_ms = ms - match_ms
        um = info["unmatched_reads"]
        out["unmatched_count"] = {
            "kernel": "sgdx_unmatched_count (after the assignment kernel, same stream)", "match_plus_count_ms": ms,
            "extra_ms": extra_ms, "unmatched_reads": um, "distinct_keys": info["distinct_keys"],
            "dropped_reads": info["dropped_reads"], "unmatched_reads_per_s": um / (extra_ms * 1e-3) if extra_ms > 0 else None,
            "bytes": n * 2 + um * (L + 24), "note": "every step starts from an empty table (first-touch claims included)"}
        # the same on the fast input format: compact records through sgdx_match_ph, keys built from the planes
        d_rec = torch.empty(n, dtype=torch.int64, device="cuda")
        m2.pack_compact_device(d_in.data_ptr(), n, d_rec.data_ptr(), slot=0)
        ref_keys = m2.unmatched_info()["distinct_keys"]
        msc, _ = timed(lambda: m2.match_compact_device(d_rec.data_ptr(), n, d_idx.data_ptr(), slot=0), before=m2.reset_counters)
        ic = m2.unmatched_info()
        assert ic["distinct_keys"] == ref_keys and ic["unmatched_reads"] == um, (ic, info)
        out["unmatched_count_compact"] = {
            "kernel": m2.compact_kernel_name() + " + sgdx_unmatched_count<compact records>", "match_plus_count_ms": msc,
            "unmatched_reads": ic["unmatched_reads"], "distinct_keys": ic["distinct_keys"], "dropped_reads": ic["dropped_reads"],
            "note": "same table of keys as the ASCII run (asserted)"}
        del d_rec
        import time
        t0 = time.perf_counter()
        top = m2.unmatched_top(1000)
        out["unmatched_top1000_ms"] = 1e3 * (time.perf_counter() - t0)
        out["unmatched_top_first"] = [top[0][0].decode(), top[0][1]] if top else None
        st2.close()
    torch.cuda.synchronize()
    torch.cuda.set_stream(prev_stream)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=50_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="qual,qual_index,unmatched,extract,route,demultiplex")
    ap.add_argument("--stride", type=int, default=150)
    a = ap.parse_args()
    print(json.dumps(measure(a.reads, a.steps, a.warmup, a.only, a.stride)))


if __name__ == "__main__":
    sys.exit(main())
