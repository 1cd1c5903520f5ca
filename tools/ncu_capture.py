#!/usr/bin/env python3
"""One launch of every assignment kernel of a BASELINE config inside a cudaProfilerStart/Stop window, for
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/<name> \
      python tools/ncu_capture.py --config C2
The window holds, in this order: the packed-record kernel of the config (compact 64-bit records, or 16-byte records
above 21 bases) and the ASCII kernel, each after three warm-up launches outside the window; with --post also the
routing kernels.  Reads: the config's seeded stream, resident in HBM (100 M reads, or the config's size if smaller)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--reads", type=int, default=0)
    ap.add_argument("--post", action="store_true")
    args = ap.parse_args()
    import torch
    import sgdx_loader
    sg = sgdx_loader.load()
    w = sg.synth.CONFIGS[args.config]
    n = args.reads or min(w.n_reads, 100_000_000)
    L, S = w.barcode_len, w.num_samples
    table = w.table()
    kind = sg.select_matcher(L)
    cls = sg.PreComputeMatcher if kind == sg.MatcherKind.PreCompute else sg.CachedHammingDistanceMatcher
    m = cls([bytes(t) for t in table], w.max_mismatches, w.min_delta, w.free_ns, max_no_calls=w.max_no_calls,
            index_lengths=w.index_lengths)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    m.set_stream(0, stream.cuda_stream)
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, L), dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), stream.cuda_stream)
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    if L <= 21:
        d_rec = torch.empty(n, dtype=torch.int64, device="cuda")
        m.pack_compact_device(d_in.data_ptr(), n, d_rec.data_ptr(), slot=0)
        packed = lambda: m.match_compact_device(d_rec.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr(), slot=0)
        pname = m.compact_kernel_name()
    else:
        d_rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
        m.pack_device(d_in.data_ptr(), n, d_rec.data_ptr(), slot=0)
        packed = lambda: m.match_packed_device(d_rec.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr(), slot=0)
        pname = m.packed_kernel_name()
    ascii_ = lambda: m.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr(), slot=0)
    d_order = torch.empty(n, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
    route = lambda: m.route_device(d_idx.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr())
    for fn in (packed, ascii_) + ((route,) if args.post else ()):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    packed()
    ascii_()
    if args.post:
        route()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"{args.config}: {n} reads, packed-record kernel {pname}, ASCII kernel {m.kernel_name()}")
    m.close()


if __name__ == "__main__":
    main()
