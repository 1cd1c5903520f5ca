#!/usr/bin/env python3
"""Compact-record path (include/sgdx.h: one 64-bit word per read) on one B200, next to the ASCII path, on the
same seeded reads of a BASELINE config:

  python tools/bench_compact.py [--config C2] [--reads N] [--steps K] [--warmup W] [--pack-threads T]

Prints one JSON object:
  device.ascii / device.compact   kernel time (CUDA events around the launch, sgdx_sync's device_ms), inputs
                                  resident in HBM and larger than L2; compact = sgdx_match_compact on records
                                  packed beforehand by sgdx_pack_compact_device
  e2e.ascii                       bench.py's e2e: ASCII in pinned host memory -> sgdx_match_batch -> results on host
  e2e.compact_prepacked           records already in pinned host memory -> sgdx_match_compact_batch (8 B/read up)
  e2e.compact_pack_inside         ASCII in pageable host memory, sgdx_pack_compact_host into the pinned buffer
                                  chunk by chunk inside the timed region (what a batching stage does instead of
                                  its memcpy), then the same call
Every variant's index array is compared with the ASCII kernel's.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=4_000_000)
    ap.add_argument("--pack-threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args(argv)


def measure(**kw):
    """Returns the numbers as a dict; used by this script and by bench.py's "compact" key."""
    args = parse_args([])
    for k, v in kw.items():
        setattr(args, k, v)

    import numpy as np
    import torch
    import sgdx_loader
    sg = sgdx_loader.load()
    lib = sg.lib()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]

    w = sg.synth.CONFIGS[args.config]
    n, L, S = args.reads, w.barcode_len, w.num_samples
    table = w.table()
    kind = sg.select_matcher(L)
    cls = sg.PreComputeMatcher if kind == sg.MatcherKind.PreCompute else sg.CachedHammingDistanceMatcher
    m = cls([bytes(t) for t in table], w.max_mismatches, w.min_delta, w.free_ns, max_no_calls=w.max_no_calls,
            index_lengths=w.index_lengths, slots=4)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, L), dtype=torch.uint8, device="cuda")
    d_rec = torch.empty(n, dtype=torch.int64, device="cuda")
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_idx_c = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist_c = torch.empty(n, dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    m.pack_compact_device(d_in.data_ptr(), n, d_rec.data_ptr())
    m.sync()

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
            m.sync()
        ms = []
        for _ in range(args.steps):
            fn()
            ms.append(m.sync())
        return sum(ms) / len(ms), min(ms)

    out = {"workload": f"{args.config} {S}x{L}, {n} reads, 1 B200", "hbm_peak_gbs": peak, "steps": args.steps,
           "path_flags": m.path_flags(), "perfect_hash_kernel": bool(m.path_flags() & 64)}
    a_avg, a_best = timed(lambda: m.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr()))
    c_avg, c_best = timed(lambda: m.match_compact_device(d_rec.data_ptr(), n, d_idx_c.data_ptr(), d_dist_c.data_ptr()))
    torch.cuda.synchronize()
    same = bool(torch.equal(d_idx, d_idx_c) and torch.equal(d_dist, d_dist_c))
    out["device"] = {
        "ascii": {"kernel": m.kernel_name(), "avg_ms": a_avg, "best_ms": a_best, "reads_per_s": n / a_avg * 1e3,
                  "buffer_bytes_per_read": L + 3, "buffer_gbs": n * (L + 3) / a_avg / 1e6},
        "compact": {"kernel": "sgdx_match_ph" if m.path_flags() & 64 else "sgdx_match_compact", "avg_ms": c_avg, "best_ms": c_best, "reads_per_s": n / c_avg * 1e3,
                    "algorithmic_bytes_per_read": 11, "achieved_gbs": n * 11 / c_avg / 1e6,
                    "frac_of_hbm_peak": n * 11 / c_avg / 1e6 / peak, "equals_ascii_kernel": same},
    }

    if not args.no_e2e:
        hin, hrec, hidx, hdist = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        for p, b in ((hin, n * L), (hrec, n * 8), (hidx, n * 2), (hdist, n)):
            sg._lib.check(lib.sgdx_alloc_host(b, C.byref(p)))
        h_in = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_uint8)), (n, L))
        h_rec = np.ctypeslib.as_array(C.cast(hrec, C.POINTER(C.c_uint64)), (n,))
        h_idx = np.ctypeslib.as_array(C.cast(hidx, C.POINTER(C.c_uint16)), (n,))
        torch.from_numpy(h_in).copy_(d_in)
        torch.from_numpy(h_rec.view(np.int64)).copy_(d_rec)
        torch.cuda.synchronize()
        pageable = np.array(h_in)  # the extractor's output: ordinary host memory
        ref_idx = d_idx.cpu().numpy().view(np.uint16)
        chunk, nslots = min(args.chunk, n), m.slots
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(args.pack_threads)  # ctypes calls release the GIL

        def pack_async(off, c):  # a persistent pool of host threads, each packing its share of the chunk
            per = (c + args.pack_threads - 1) // args.pack_threads
            return [pool.submit(lib.sgdx_pack_compact_host, pageable.ctypes.data + (off + b) * L, min(per, c - b), L,
                                hrec.value + (off + b) * 8, 1) for b in range(0, c, per)]

        def run(mode):
            chunks = [(off, min(chunk, n - off)) for off in range(0, n, chunk)]
            futs = pack_async(*chunks[0]) if mode == "pack" else None
            for i, (off, c) in enumerate(chunks):
                slot = i % nslots
                if mode == "pack":  # chunk i + 1 is packed while chunk i is submitted and uploaded
                    for f in futs:
                        sg._lib.check(f.result())
                    futs = pack_async(*chunks[i + 1]) if i + 1 < len(chunks) else None
                if i >= nslots:
                    m.sync(slot)
                if mode == "ascii":
                    m.submit_ptr(hin.value + off * L, c, hidx.value + off * 2, hdist.value + off, slot)
                else:
                    m.submit_compact_ptr(hrec.value + off * 8, c, hidx.value + off * 2, hdist.value + off, slot)
            for s in range(nslots):
                m.sync(s)

        e2e = {}
        for mode, name, up in (("ascii", "ascii", L), ("prepacked", "compact_prepacked", 8), ("pack", "compact_pack_inside", 8)):
            h_idx[:] = 0xFFFF
            run(mode)
            steps = max(1, min(args.steps, 5))
            t0 = time.perf_counter()
            for _ in range(steps):
                run(mode)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / steps
            e2e[name] = {"reads_per_s": n / dt, "ms_per_step": dt * 1e3, "h2d_bytes_per_read": up, "d2h_bytes_per_read": 3,
                         "pcie_up_gbs": n * up / dt / 1e9, "equals_ascii_kernel": bool(np.array_equal(h_idx, ref_idx))}
        e2e["compact_pack_inside"]["pack_threads"] = args.pack_threads
        t0 = time.perf_counter()
        sg._lib.check(lib.sgdx_pack_compact_host(pageable.ctypes.data, n, L, hrec.value, args.pack_threads))
        dt = time.perf_counter() - t0
        e2e["host_pack_alone"] = {"reads_per_s": n / dt, "threads": args.pack_threads, "host_cores": os.cpu_count()}
        out["e2e"] = e2e
        for p in (hin, hrec, hidx, hdist):
            lib.sgdx_free_host(p)
    m.close()
    return out


def main():
    print(json.dumps(measure(**vars(parse_args()))))


if __name__ == "__main__":
    main()
