"""Randomised soak of the public API against a model (CPU, host-memory pool): three
connections write / read / touch / probe while the server evicts (or auto-extends), is purged,
checkpointed and restored, and clients reconnect.  Every successful read is checked against
the value the model expects.

    python tools/soak_cpu.py SEED SECONDS [MAX_OPS]
"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import infinistore_b200 as ist  # noqa: E402
from infinistore_b200 import _infinistore as m  # noqa: E402

PAGE = 1024  # elements (float32) = 4 KB


def make_server(rng):
    cfg = m.ServerConfig()
    cfg.service_port = 0
    cfg.host = "127.0.0.1"
    cfg.pool_backend = "host"
    cfg.prealloc_bytes = 256 * 16384
    cfg.minimal_allocate_size = 16
    force_auto = bool(os.environ.get("SOAK_AUTO"))
    cfg.evict = (rng.random() < 0.7) and not force_auto
    cfg.evict_ratio = 0.1
    cfg.auto_increase = (not cfg.evict) and (rng.random() < 0.5 or force_auto)
    cfg.extend_size = 1
    srv = m.Server(cfg)
    return cfg, srv, srv.start()


def connect(port):
    c = ist.InfinityConnection(ist.ClientConfig(
        host_addr="127.0.0.1", service_port=port, connection_type=ist.TYPE_RDMA, log_level="error"))
    c.connect()
    return c


def buffers(conn):
    src, dst = torch.zeros(64 * PAGE), torch.zeros(64 * PAGE)
    conn.register_mr(src)
    conn.register_mr(dst)
    return src, dst


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dur = float(sys.argv[2]) if len(sys.argv) > 2 else 60
    max_ops = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rng = random.Random(seed)
    cfg, srv, port = make_server(rng)
    conns = [connect(port) for _ in range(3)]
    bufs = [buffers(c) for c in conns]
    model = {}   # key -> (fill value, elements) of committed keys we believe may exist
    next_id = 0
    t_end = time.time() + dur
    ops = 0
    while time.time() < t_end and ops < max_ops:
        ci = rng.randrange(3)
        c = conns[ci]
        src, dst = bufs[ci]
        op = rng.random()
        ops += 1
        if op < 0.45:  # write batch
            n = rng.randint(1, 32)
            keys = []
            for _ in range(n):
                if model and rng.random() < 0.1:
                    keys.append(rng.choice(list(model)))
                else:
                    keys.append(f"k{next_id}")
                    next_id += 1
            vals = [float(rng.randint(1, 1 << 20)) for _ in keys]
            for i, v in enumerate(vals):
                src[i * PAGE:(i + 1) * PAGE] = v
            nbytes = rng.choice([4096, 4096, 8000, 16384, 20000])
            elems = min(PAGE, nbytes // 4)
            try:
                blocks = c.allocate_rdma(keys, nbytes)
            except Exception:  # pool full and nothing evictable
                continue
            c.rdma_write_cache(src, [i * PAGE for i in range(n)], elems, blocks)
            c.sync()
            for k, v, rkey in zip(keys, vals, blocks["rkey"]):
                if rkey != 0:  # not deduplicated: this write is the key's value
                    model[k] = (v, elems)
        elif op < 0.8 and model:  # read batch
            ks = [rng.choice(list(model)) for _ in range(rng.randint(1, 16))]
            elems = min(model[k][1] for k in ks)
            dst.zero_()
            try:
                c.read_cache(dst, [(k, i * PAGE) for i, k in enumerate(ks)], elems)
                c.sync()
            except Exception:  # something was evicted / purged: drop what is gone
                for k in set(ks):
                    if not c.check_exist(k):
                        model.pop(k, None)
                continue
            for i, k in enumerate(ks):
                got = dst[i * PAGE:i * PAGE + elems]
                assert bool((got == model[k][0]).all()), (k, model[k], float(got[0]))
        elif op < 0.85 and model:
            c.touch([rng.choice(list(model)) for _ in range(8)] + ["nope"])
        elif op < 0.9 and model:
            k = rng.choice(list(model))
            if not c.check_exist(k):
                model.pop(k)
        elif op < 0.93:
            ks = [f"k{i}" for i in range(rng.randint(0, max(1, next_id)), next_id)][:50]
            if ks:
                try:
                    c.get_match_last_index(ks)
                except Exception:
                    pass
        elif op < 0.95:
            path = f"/tmp/soak_{os.getpid()}.ckpt"
            dumped = srv.dump(path)
            if rng.random() < 0.5:
                srv.purge()
                loaded = srv.load(path)
                assert loaded == dumped, (loaded, dumped)
            os.unlink(path)
        elif op < 0.96:
            srv.purge()
            model.clear()
        elif op < 0.98:  # reconnect one client
            conns[ci].close()
            conns[ci] = connect(port)
            bufs[ci] = buffers(conns[ci])
        else:
            st = srv.stats()
            assert st["used_bytes"] <= st["pool_bytes"] and st["inflight"] == 0, st
    st = srv.stats()
    print(f"seed {seed}: {ops} ops, evict={cfg.evict} auto={cfg.auto_increase} keys={st['keys']} "
          f"evicted={st['evicted']} segs={st['segments']} model={len(model)} OK")
    for c in conns:
        c.close()
    srv.stop()


if __name__ == "__main__":
    main()
