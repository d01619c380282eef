"""Wire protocol: framing constants and the hand-written FlatBuffers codec.

No flatc exists on the image, so the native codec is cross-checked against an independent
pure-Python reader/writer (tests/fbpy.py) instead of flatc golden bytes — any
spec-conformant layout interoperates.
"""
import struct

import numpy as np
import pytest

import fbpy
from infinistore_b200 import _infinistore as m

T = m.testing


def test_struct_sizes_match_reference_framing():
    assert T.header_size() == 9       # magic u32 + op char + body_size u32, packed
    assert T.conn_info_size() == 30   # qpn, psn, gid[16], lid u16, mtu u32, packed


def test_remote_block_numpy_abi():
    blob = T.encode_allocate_response([(1, 7, 0x100000001000), (0, 0, 0)])
    arr = T.decode_allocate_response(blob)
    assert arr.dtype.itemsize == 16
    assert arr.dtype.fields["rkey"][1] == 0
    assert arr.dtype.fields["remote_addr"][1] == 8
    assert arr["rkey"].tolist() == [1, 0]
    assert arr["remote_addr"].tolist() == [0x100000001000, 0]
    assert arr["gen"].tolist() == [7, 0]


def test_remote_meta_roundtrip_and_python_reader():
    keys = ["key-%d" % i for i in range(5)] + [""]
    addrs = [1 << 44, (1 << 44) + 65536, 2 ** 63 + 5]
    blob = T.encode_remote_meta(keys, 4096, 77, addrs, "A", 3)
    d = T.decode_remote_meta(blob)
    assert [k.decode() for k in d["keys"]] == keys
    assert (d["block_size"], d["rkey"], d["op"], d["hint"]) == (4096, 77, "A", 3)
    assert d["remote_addrs"] == addrs
    # independent reader, reference field slots: keys=4 block_size=6 rkey=8 addrs=10 op=12
    t = fbpy.root(blob)
    assert [k.decode() for k in t.string_vector(4)] == keys
    assert t.scalar(6, "i") == 4096
    assert t.scalar(8, "I") == 77
    assert [v[0] for v in t.vector(10, "Q")] == addrs
    assert t.scalar(12, "b") == ord("A")


def test_defaults_are_omitted_like_flatc():
    # a COMMIT message has no keys, block_size or rkey (reference: libinfinistore.cpp:372)
    blob = T.encode_remote_meta([], 0, 0, [123, 456], "T")
    t = fbpy.root(blob)
    assert t.string_vector(4) is None
    assert t._field(6) == 0 and t._field(8) == 0
    d = T.decode_remote_meta(blob)
    assert d["keys"] == [] and d["block_size"] == 0 and d["remote_addrs"] == [123, 456]
    assert d["hint"] == -1  # absent extension field decodes as "any"


def test_u64_vector_is_8_byte_aligned():
    for nkeys in range(4):
        blob = T.encode_remote_meta(["k" * (i + 1) for i in range(nkeys)], 1, 0, [1, 2, 3], "A")
        t = fbpy.root(blob)
        p = t._indirect(10)
        assert (p + 4) % 8 == 0


def test_python_writer_to_native_reader():
    blob = fbpy.build_table([
        (4, "vec_str", [b"alpha", b"beta", b"x" * 40]),
        (6, "i32", 32768),
        (8, "u32", 9),
        (10, "vec_u64", [5, 6, 7, 8]),
        (12, "i8", ord("D")),
    ])
    d = T.decode_remote_meta(blob)
    assert d["keys"] == [b"alpha", b"beta", b"x" * 40]
    assert d["block_size"] == 32768 and d["rkey"] == 9 and d["op"] == "D"
    assert d["remote_addrs"] == [5, 6, 7, 8]

    blob = fbpy.build_table([(4, "vec_struct16", [(1, 2, 3), (4, 5, 6)])])
    arr = T.decode_allocate_response(blob)
    assert arr.tolist() == [(1, 2, 3), (4, 5, 6)]

    blob = fbpy.build_table([(4, "vec_str", [b"a", b"bb"])])
    assert T.decode_match_request(blob) == [b"a", b"bb"]


def test_local_meta_roundtrip():
    ipc = bytes(range(64))
    blocks = [("key%d" % i, i * 8192) for i in range(7)]
    blob = T.encode_local_meta(3, ipc, 16384, blocks)
    d = T.decode_local_meta(blob)
    assert d["device"] == 3 and d["ipc_handle"] == ipc and d["block_size"] == 16384
    assert [(k.decode(), o) for k, o in d["blocks"]] == blocks
    # reference slots: device=4 ipc_handle=6 block_size=8 blocks=10; Block: key=4 offset=6
    t = fbpy.root(blob)
    assert t.scalar(4, "i") == 3 and t.scalar(8, "i") == 16384
    assert bytes(v[0] for v in t.vector(6, "B")) == ipc
    tabs = t.table_vector(10)
    assert [(b.string(4).decode(), b.scalar(6, "Q")) for b in tabs] == blocks
    # device == 0 and offset == 0 are defaults: not stored
    blob0 = T.encode_local_meta(0, b"", 4096, [("k", 0)])
    t0 = fbpy.root(blob0)
    assert t0._field(4) == 0 and t0._field(6) == 0
    assert t0.table_vector(10)[0]._field(6) == 0
    assert T.decode_local_meta(blob0)["blocks"] == [(b"k", 0)]


def test_match_request_roundtrip():
    keys = ["A", "B", "key1", ""]
    blob = T.encode_match_request(keys)
    assert [k.decode() for k in T.decode_match_request(blob)] == keys
    assert [k.decode() for k in fbpy.root(blob).string_vector(4)] == keys


@pytest.mark.parametrize("decoder", ["decode_remote_meta", "decode_allocate_response",
                                     "decode_local_meta", "decode_match_request"])
def test_malformed_buffers_are_rejected_not_dereferenced(decoder):
    fn = getattr(T, decoder)
    good = T.encode_remote_meta(["abc", "def"], 4096, 1, [1, 2], "A")
    rng = np.random.default_rng(0)
    for blob in [b"", b"\x00", b"\xff" * 3, b"\xff" * 64, struct.pack("<I", 1 << 30) + b"\0" * 32]:
        with pytest.raises(Exception):
            fn(blob)
    # random corruption / truncation must never crash the process
    for _ in range(300):
        b = bytearray(good)
        for _ in range(rng.integers(1, 6)):
            b[rng.integers(0, len(b))] = rng.integers(0, 256)
        cut = rng.integers(0, len(b) + 1)
        try:
            fn(bytes(b[:cut]))
        except Exception:
            pass


def test_hash_is_stable_and_nonzero():
    h1, h2 = T.hash_key(b"hello")
    assert (h1, h2) == T.hash_key(b"hello")
    assert h1 != 0 and h1 != h2
    seen = {T.hash_key(b"key-%d" % i) for i in range(20000)}
    assert len(seen) == 20000
    for n in range(0, 40):  # every tail length of the 16-byte loop
        assert T.hash_key(b"a" * n) != T.hash_key(b"a" * n + b"\0")


def test_fd_passing_side_channel():
    """VMM / multicast handles travel as POSIX fds over an abstract unix socket (SCM_RIGHTS)."""
    assert T.fd_pass_selftest()
