"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/*.h declares,
validates arguments, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sgdx_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sg = sgdx_loader.load()


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sgdx_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = sg.lib()
    declared = _declared("sgdx.h") + _declared("sgdx_synth.h")
    assert len(declared) >= 27
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported by libsgdx.so"
    from singular_demux_b200 import _lib
    assert sorted(_lib.SYMBOLS) == sorted(declared)


def test_struct_layouts_match_header():
    from singular_demux_b200 import _lib
    assert C.sizeof(_lib.Config) == 12 * 4
    assert C.sizeof(_lib.Totals) == 4 * 8
    assert C.sizeof(_lib.SynthParams) == 8 + 9 * 4 + 4  # padded to 8


def test_create_validates_before_touching_cuda():
    from singular_demux_b200 import _lib
    L = sg.lib()
    h = C.c_void_p()

    def create(S, Lb, table, **kw):
        cfg = _lib.Config(S, Lb, 1, 2, 1, -1, kw.get("matcher", 0), kw.get("i1", 0), kw.get("i2", 0), 0, 2, 0)
        return L.sgdx_create(C.byref(cfg), table, C.byref(h))

    assert create(0, 8, b"") == -1 and b"num_samples" in L.sgdx_last_error()
    assert create(1, 65, b"A" * 65) == -1 and b"barcode_len" in L.sgdx_last_error()
    assert create(1, 40, b"A" * 40, i1=36, i2=4) == -1 and b"hop checker" in L.sgdx_last_error()
    assert create(1, 8, b"ACGTNACG") == -1 and b"upper-case ACGT" in L.sgdx_last_error()
    assert create(1, 8, b"acgtacgt") == -1
    assert create(1, 8, b"ACGTACGT", i1=3, i2=4) == -1
    assert create(1, 8, b"ACGTACGT", matcher=7) == -1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(sg.SgdxError, match="no CPU fallback"):
        sg.CachedHammingDistanceMatcher([b"ACGTACGTACGTACGT"], 1, 2, 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "singular-demux_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "orc_", "sgdx_oracle"):
                    assert needle not in src, f"{f} references the oracle ({needle})"


def test_matcher_kind_and_selection_mirror_the_reference():
    assert sg.select_matcher(12) == sg.MatcherKind.PreCompute
    assert sg.select_matcher(13) == sg.MatcherKind.CachedHammingDistance
    assert sg.select_matcher(20, sg.MatcherKind.PreCompute) == sg.MatcherKind.PreCompute
    assert sg.select_matcher(8, sg.MatcherKind.CachedHammingDistance) == sg.MatcherKind.PreCompute
    assert sg.MatcherKind.from_str("pre-compute") == sg.MatcherKind.PreCompute
    assert sg.MatcherKind.from_str("Cached-Hamming-Distance") == sg.MatcherKind.CachedHammingDistance
    with pytest.raises(ValueError):
        sg.MatcherKind.from_str("nope")
    r = sg.MatchResult.Match(3, 1, b"ACGT")
    assert r.is_match() and not r.is_no_match() and r.barcode == b"ACGT"
    assert sg.MatchResult.NoMatch(b"AC").is_no_match()


def test_metrics_merge_is_elementwise_sum():
    a = sg.DemuxedGroupMetrics([sg.DemuxedGroupSampleMetrics(1, 2, 5), sg.DemuxedGroupSampleMetrics(0, 0, 3)], 8,
                               sample_barcode_hop_tracker=sg.SampleBarcodeHopTracker(3, 3, {b"TTTAAA": 1}))
    b = sg.DemuxedGroupMetrics([sg.DemuxedGroupSampleMetrics(4, 0, 4), sg.DemuxedGroupSampleMetrics(0, 0, 1)], 5,
                               sample_barcode_hop_tracker=sg.SampleBarcodeHopTracker(3, 3, {b"TTTAAA": 2, b"CCCGGG": 1}))
    a.update_with(b)
    assert a.total_templates == 13
    assert a.per_sample_array().tolist() == [[9, 5, 2], [4, 0, 0]]
    assert a.sample_barcode_hop_tracker.count_tracker == {b"TTTAAA": 3, b"CCCGGG": 1}
    rows = a.sample_barcode_hop_tracker.barcode_counts()
    assert rows[0] == sg.BarcodeCount("TTT+AAA", 3) and rows[1] == sg.BarcodeCount("CCC+GGG", 1)


def test_synthetic_generator_is_deterministic_and_shardable():
    w = sg.synth.Workload("t", 48, 16, 8, 1, 2, p_n=0.01, p_x=0.001)
    t = w.table()
    assert len({bytes(r) for r in t}) == 48 and set(np.unique(t)) <= set(b"ACGT")
    for half in (t[:, :8], t[:, 8:]):
        for i in range(48):
            d = (half != half[i]).sum(axis=1)
            d[i] = 99
            assert d.min() >= 3
    full = w.reads_host(0, 5000, t, threads=1)
    again = w.reads_host(0, 5000, t, threads=4)
    assert np.array_equal(full, again)
    assert np.array_equal(full[1234:2234], w.reads_host(1234, 1000, t))
    assert set(np.unique(full)) <= set(b"ACGTNR") and (full == ord("N")).any()


def test_unmatched_counter_downsizes_like_the_reference():
    c = sg.UnmatchedCounter(max_map_size=4, downsize_to=2)
    keys = np.frombuffer(b"AAACCCGGGTTTACG", dtype=np.uint8).reshape(5, 3)
    c.insert_counts(keys, np.array([5, 1, 4, 1, 1]))
    assert len(c.counts) <= 4 and c.counts[b"AAA"] == 5 and c.counts[b"GGG"] == 4
    assert c.most_frequent(1)[0] == sg.BarcodeCount("AAA", 5)


def test_cpp_host_mirror_builds_and_refuses_to_run_without_a_gpu():
    """The compiled-language host side (C++ mirror of the reference's interface) builds against the C ABI; on a
    box without a CUDA device it must fail loudly (there is no CPU fallback), not pretend to pass."""
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "tests", "cpp")], check=True, capture_output=True)
    exe = os.path.join(root, "tests", "cpp", "test_host_mirror")
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the gpu-marked test runs the program")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def test_headers_are_plain_c(tmp_path):
    """The boundary is a C ABI: include/*.h must compile as C99 (no C++ types, no torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "sgdx.h"\n#include "sgdx_synth.h"\n'
                   "int main(void) { sgdx_config c; sgdx_segment s = {0, 0, SGDX_SEG_TO_END, SGDX_SEG_TEMPLATE};\n"
                   "  sgdx_demux_out o = {0}; sgdx_fastq_view v = {0}; (void)c; (void)s; (void)o; (void)v; return 0; }\n")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_c_entry_point_runs_behind_the_exception_guard():
    """include/sgdx.h promises that no exception crosses the boundary: every `extern "C" int` function of the library
    wraps its body in sgdx::guarded (csrc/sgdx_guard.h)."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    total = 0
    for path in glob.glob(os.path.join(root, "singular-demux_b200", "csrc", "*.c*")):
        src = open(path).read()
        for m in re.finditer(r'^extern "C" int (\w+)\(', src, re.M):
            body_at = src.index(") {\n", m.start()) + 4
            assert src[body_at:body_at + 200].lstrip().startswith("return sgdx::guarded("), (os.path.basename(path), m.group(1))
            total += 1
    assert total >= 40
