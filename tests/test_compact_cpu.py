"""Host side of the compact read format (include/sgdx.h, sgdx_pack_compact_host): the AVX2/scalar packer in
libsgdx.so against a numpy restatement of the format, on every barcode length and on the bytes that matter
(N, lower case, control bytes, bytes with the high bit set), single- and multi-threaded.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import sgdx_loader

sg = sgdx_loader.load()


def records_numpy(arr):
    """lo plane | hi plane << 21 | invalid plane << 42; code = (ascii >> 1) & 3; at an invalid position
    lo = 0 for 'N' and 1 for any other byte, hi = 0."""
    n, L = arr.shape
    a = arr.astype(np.uint64)
    ok = np.isin(arr, np.frombuffer(b"ACGT", dtype=np.uint8))
    isn = arr == ord("N")
    j = np.arange(L, dtype=np.uint64)
    lo = np.where(ok, (a >> np.uint64(1)) & np.uint64(1), np.where(isn, 0, 1).astype(np.uint64))
    hi = np.where(ok, (a >> np.uint64(2)) & np.uint64(1), np.uint64(0))
    inv = (~ok).astype(np.uint64)
    return ((lo << j).sum(axis=1) | ((hi << j).sum(axis=1) << np.uint64(21)) |
            ((inv << j).sum(axis=1) << np.uint64(42))).astype(np.uint64)


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
