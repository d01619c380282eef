"""One process per GPU (torchrun): every rank hosts a pool shard and talks to the others."""
from __future__ import annotations


from .. import _infinistore
from ..lib import ClientConfig, InfinityConnection, TYPE_RDMA
from .sharded import ShardedConnection


def shard_port(base_port: int, rank: int) -> int:
    return base_port + rank


def ring_peer(rank: int, world: int, hop: int = 1) -> int:
    """Shard used by `rank` under ring placement: with world >= 2 every byte crosses NVLink."""
    return (rank + hop) % world


def start_shard_server(device: int, port: int, pool_bytes: int, granule_kb: int = 64,
                       host: str = "127.0.0.1", auto_increase: bool = False,
                       log_level: str = "warning", evict: bool = False,
                       evict_ratio: float = 0.05):
    """Start a store server thread in this process with an HBM pool on `device`.
    Returns the native Server object (keep a reference; ``.stop()`` to shut down)."""
    cfg = _infinistore.ServerConfig()
    cfg.service_port = port
    cfg.host = host
    cfg.pool_backend = "hbm" if _infinistore.cuda_available() else "host"
    cfg.pool_devices = [device]
    cfg.prealloc_bytes = pool_bytes
    cfg.minimal_allocate_size = granule_kb
    cfg.auto_increase = auto_increase
    cfg.evict = evict
    cfg.evict_ratio = evict_ratio
    cfg.log_level = log_level
    srv = _infinistore.Server(cfg)
    srv.start()
    return srv


def connect_all(base_port: int, world: int, device: int, host: str = "127.0.0.1",
                **client_kwargs) -> ShardedConnection:
    """ShardedConnection over the shards of all ranks (call after a barrier that follows
    every rank's ``start_shard_server``)."""
    cfgs = [ClientConfig(host_addr=host, service_port=shard_port(base_port, r),
                         connection_type=TYPE_RDMA, device=device, **client_kwargs)
            for r in range(world)]
    conn = ShardedConnection(cfgs)
    conn.connect()
    return conn


def connect_peer(base_port: int, rank: int, world: int, device: int, host: str = "127.0.0.1",
                 hop: int = 1, **client_kwargs) -> InfinityConnection:
    conn = InfinityConnection(ClientConfig(
        host_addr=host, service_port=shard_port(base_port, ring_peer(rank, world, hop)),
        connection_type=TYPE_RDMA, device=device, **client_kwargs))
    conn.connect()
    return conn
