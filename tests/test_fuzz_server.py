"""Property-based hardening tests (hypothesis): whatever bytes arrive on the service port,
the reactor keeps serving.  The reference dereferences unverified flatbuffers, trusts
body_size and spins forever on an unknown op byte."""
import socket
import struct

from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import make_conn

MAGIC = 0xDEADBEEF
OPS = b"RWSEDATCMPUH"


def _send_and_drain(port, payload: bytes):
    s = socket.create_connection(("127.0.0.1", port), timeout=2)
    s.settimeout(0.2)
    try:
        s.sendall(payload)
        try:
            while s.recv(65536):
                pass
        except (socket.timeout, ConnectionResetError, BrokenPipeError):
            pass
    finally:
        s.close()


@settings(max_examples=60, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(op=st.sampled_from(list(OPS) + [0, 0x7f, 0xff]),
       body=st.binary(min_size=0, max_size=512),
       claimed=st.one_of(st.none(), st.integers(min_value=0, max_value=(1 << 32) - 1)),
       magic_ok=st.booleans())
def test_random_requests_never_take_the_server_down(host_server, op, body, claimed, magic_ok):
    srv, port = host_server
    size = len(body) if claimed is None else claimed
    hdr = struct.pack("<IBI", MAGIC if magic_ok else MAGIC ^ 0x1, op, size)
    _send_and_drain(port, hdr + body)
    assert srv.running()
    # a well-formed client is still served
    conn = make_conn(port)
    try:
        assert conn.check_exist("never-written-key") is False
    finally:
        conn.close()


@settings(max_examples=30, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(chunks=st.lists(st.binary(min_size=1, max_size=64), min_size=1, max_size=12))
def test_arbitrary_byte_streams(host_server, chunks):
    srv, port = host_server
    _send_and_drain(port, b"".join(chunks))
    assert srv.running()
    before = srv.stats()["used_bytes"]
    conn = make_conn(port)
    try:
        blocks = conn.allocate_rdma(["fuzz-after"], 4096)
        assert len(blocks) == 1
    finally:
        conn.close()
    # the dead client's uncommitted reservation is released again
    import time
    deadline = time.time() + 5
    while srv.stats()["used_bytes"] != before and time.time() < deadline:
        time.sleep(0.01)
    assert srv.stats()["used_bytes"] == before


@settings(max_examples=40, deadline=None, derandomize=True)
@given(ops=st.lists(st.tuples(st.booleans(), st.integers(min_value=1, max_value=5 * 4096)),
                    min_size=1, max_size=60))
def test_mempool_model(ops):
    """The bitmap allocator against a set-of-granules model: no overlap, exact accounting,
    everything freed at the end."""
    from infinistore_b200 import _infinistore as m

    pool = m.testing.MemoryPool(64 * 4096, 4096, -1)
    live = {}
    owned = set()
    for is_alloc, size in ops:
        if is_alloc or not live:
            off = pool.allocate(size)
            k = (size + 4095) // 4096
            if off is None or off < 0:  # full, or too fragmented for a run of k granules
                continue
            g = set(range(off // 4096, off // 4096 + k))
            assert not (g & owned)
            owned |= g
            live[off] = size
        else:
            off, size = next(iter(live.items()))
            del live[off]
            k = (size + 4095) // 4096
            owned -= set(range(off // 4096, off // 4096 + k))
            assert pool.deallocate(off, size)
        assert pool.used_blocks() == len(owned)
    for off, size in list(live.items()):
        assert pool.deallocate(off, size)
    assert pool.used_blocks() == 0
