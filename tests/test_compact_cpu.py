"""Host side of the compact read format (include/sgdx.h, sgdx_pack_compact_host): the AVX2/scalar packer in
libsgdx.so against a numpy restatement of the format, on every barcode length and on the bytes that matter
(N, lower case, control bytes, bytes with the high bit set), single- and multi-threaded.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()


records_numpy = orc.compact_records


ALPHABET = np.frombuffer(b"ACGTACGTACGTACGTNNacgtn\x00\x01\x41\x43\x47\x54\x4e\x80\xc1\xff\x0f\x11\x31\x51\x61", dtype=np.uint8)


@pytest.mark.parametrize("L", list(range(1, 22)))
def test_host_packer_equals_the_format_definition(L):
    rng = np.random.default_rng(1000 + L)
    n = 4099  # the last reads sit closer than 32 bytes to the end of the buffer: scalar tail
    arr = ALPHABET[rng.integers(0, ALPHABET.size, size=(n, L))]
    arr[::7] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(len(arr[::7]), L))]
    want = records_numpy(arr)
    assert np.array_equal(sg.pack_compact_host(arr), want)
    assert np.array_equal(sg.pack_compact_host(arr, threads=3), want)


def test_every_byte_value_at_every_position():
    L = 20
    arr = np.full((256 * L, L), ord("A"), dtype=np.uint8)
    for pos in range(L):
        arr[pos * 256:(pos + 1) * 256, pos] = np.arange(256, dtype=np.uint8)
    got = sg.pack_compact_host(arr)
    assert np.array_equal(got, records_numpy(arr))
    # only A/C/G/T leave the invalid plane clear, only 'N' is a no-call
    inv = (got >> np.uint64(42)) != 0
    col = np.tile(np.arange(256), L)
    assert set(col[~inv]) == {ord(c) for c in "ACGT"}
    lo = got & np.uint64((1 << 21) - 1)
    n_call = inv & ((lo & (got >> np.uint64(42))) == 0)
    assert set(col[n_call]) == {ord("N")}


def test_multithreaded_pack_of_a_large_batch_and_in_place_output():
    rng = np.random.default_rng(7)
    n, L = 1_000_003, 20
    arr = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=(n, L))]
    want = records_numpy(arr)
    out = np.zeros(n, dtype=np.uint64)
    assert sg.pack_compact_host(arr, threads=8, out=out) is out
    assert np.array_equal(out, want)
    assert np.array_equal(sg.pack_compact_host(arr[:0]), np.empty(0, dtype=np.uint64))


def test_packer_rejects_barcodes_longer_than_a_record():
    L = sg.lib()
    out = (C.c_uint64 * 4)()
    assert L.sgdx_pack_compact_host(b"A" * 88, 4, 22, out, 1) == -1
    assert b"compact records" in L.sgdx_last_error()
    assert L.sgdx_pack_compact_host(None, 4, 20, out, 1) == -1
    assert L.sgdx_pack_compact_host(None, 0, 20, None, 1) == 0


@pytest.mark.parametrize("cfg_name,M,D", [("C1", 1, 2), ("C2", 1, 2), ("C1", 2, 1), ("C2", 0, 5), ("C2", 2, 1)])
def test_short_seed_rule_holds_in_the_oracle(cfg_name, M, D):
    """The claim behind CompactParams.short_*: when the table's minimum pairwise distance exceeds 2M + delta, a read
    made of A/C/G/T only is Match(s, d) iff exactly one sample s is within d <= M of it (then the delta test cannot
    fail), else NoMatch -- in both rule sets.  Checked against the oracle (the restated reference rules)."""
    w = sg.synth.CONFIGS[cfg_name]
    table = w.table()
    S, L = table.shape
    pair = (table[:, None, :] != table[None, :, :]).sum(-1)
    np.fill_diagonal(pair, 99)
    assert pair.min() > 2 * M + D, "pick parameters the synthetic table satisfies"
    rng = np.random.default_rng(S + 10 * M + D)
    n = 6000
    reads = table[rng.integers(0, S, n)].copy()
    nsub = rng.integers(0, 5, n)  # 0..4 substitutions, the rest random reads
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    for r in range(n):
        pos = rng.choice(L, nsub[r], replace=False)
        reads[r, pos] = letters[rng.integers(0, 4, nsub[r])]
    reads[-500:] = letters[rng.integers(0, 4, (500, L))]
    dist = (reads[:, None, :] != table[None, :, :]).sum(-1)
    best, arg = dist.min(1), dist.argmin(1)
    want_idx = np.where(best <= M, arg, S).astype(np.uint16)
    want_dist = np.where(best <= M, best, 0xFF).astype(np.uint8)
    assert ((dist <= M).sum(1) <= 1).all()
    for kind in (orc.MATCHER_CACHED_HAMMING, orc.MATCHER_PRECOMPUTE):
        cfg = orc.Config(S, L, M, D, 1, -1, kind, 0, 0)
        got = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=2, use_cache=False)
        assert np.array_equal(got["idx"], want_idx)
        assert np.array_equal(got["dist"], want_dist)


@pytest.mark.parametrize("L", [1, 5, 8, 16, 20, 21, 24, 31, 32])
def test_host_packer_of_16_byte_records_equals_the_format_definition(L):
    """sgdx_pack_host (SURVEY Appendix B's `sgdx_pack`): {lo, hi, nmask, xmask} per read, any length up to 32."""
    rng = np.random.default_rng(2000 + L)
    n = 70_001
    arr = ALPHABET[rng.integers(0, ALPHABET.size, size=(n, L))]
    want = orc.packed_records(arr)
    assert np.array_equal(sg.pack_host(arr), want)
    assert np.array_equal(sg.pack_host(arr, threads=3), want)
    assert sg.pack_host(arr[:0]).shape == (0, 4)


def test_16_byte_packer_every_byte_value_and_argument_checks():
    L = 32
    arr = np.full((256 * L, L), ord("C"), dtype=np.uint8)
    for pos in range(L):
        arr[pos * 256:(pos + 1) * 256, pos] = np.arange(256, dtype=np.uint8)
    assert np.array_equal(sg.pack_host(arr), orc.packed_records(arr))
    lib = sg.lib()
    out = (C.c_uint32 * 16)()
    assert lib.sgdx_pack_host(b"A" * 132, 4, 33, out, 1) == -1 and b"barcode_len" in lib.sgdx_last_error()
    assert lib.sgdx_pack_host(None, 4, 20, out, 1) == -1


def test_host_packers_are_clean_under_asan(tmp_path):
    """tests/cpp/test_hostpack_asan.cpp: both packers, AVX2 and scalar, every length, exact-size buffers."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "hostpack_asan")
    build = subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                            os.path.join(root, "tests", "cpp", "test_hostpack_asan.cpp"),
                            os.path.join(root, "singular-demux_b200", "csrc", "sgdx_hostpack.cpp"), "-lpthread"],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr + build.stdout:
        pytest.skip("this g++ has no sanitizer runtime")
    assert build.returncode == 0, build.stderr
    for env in ({}, {"SGDX_PACK_NO_AVX2": "1"}):
        r = subprocess.run([exe], capture_output=True, text=True, env={**os.environ, **env}, timeout=300)
        assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr



@pytest.mark.parametrize("L", [1, 3, 4, 8, 12, 16, 20, 21, 22, 24, 25, 28, 29, 32])
def test_host_packer_of_dense_records_equals_the_format_definition(L):
    """sgdx_pack_dense_host: 2 bits per base for the plain reads, (index, compact record) for the others, in ascending
    read order, for any thread count; the exception room is checked."""
    rng = np.random.default_rng(3000 + L)
    n = 140_003
    arr = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L))]
    odd = rng.random((n, L)) < 0.01
    arr[odd] = ALPHABET[rng.integers(0, ALPHABET.size, int(odd.sum()))]
    arr[:3] = ord("N")
    want_dense, want_idx, want_rec, plain = orc.dense_records(arr)
    for threads in (1, 3):
        dense, eidx, erec = sg.pack_dense_host(arr, threads=threads)
        assert dense.shape == (n * sg.dense_bytes(L),)
        assert np.array_equal(dense.reshape(n, -1)[plain], want_dense[plain])
        assert np.array_equal(eidx, want_idx) and np.array_equal(erec, want_rec)
    with pytest.raises(sg.SgdxError):
        sg.pack_dense_host(arr, exc_cap=2)           # fewer slots than exceptions
    d0, i0, r0 = sg.pack_dense_host(arr[:0])
    assert d0.size == 0 and i0.size == 0
    with pytest.raises(sg.SgdxError):
        sg.pack_dense_host(np.full((4, 33), ord("A"), dtype=np.uint8))
