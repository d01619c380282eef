"""infinistore_b200 — a Blackwell-native KV-cache block store with infiniStore's API.

Same public surface as the reference package (infinistore/__init__.py:1-31); see
``lib.py`` for what changes underneath.
"""
from .lib import (
    InfinityConnection,
    DisableTorchCaching,
    ClientConfig,
    ServerConfig,
    TYPE_RDMA,
    TYPE_LOCAL_GPU,
    Logger,
    check_supported,
    LINK_ETHERNET,
    LINK_IB,
    register_server,
    stop_server,
    server_stats,
    purge_kv_map,
    get_kvmap_len,
    dump_kv_map,
    load_kv_map,
)

__version__ = "0.1.0"

__all__ = [
    "InfinityConnection",
    "DisableTorchCaching",
    "register_server",
    "stop_server",
    "server_stats",
    "ClientConfig",
    "ServerConfig",
    "TYPE_RDMA",
    "TYPE_LOCAL_GPU",
    "Logger",
    "check_supported",
    "LINK_ETHERNET",
    "LINK_IB",
    "purge_kv_map",
    "get_kvmap_len",
    "dump_kv_map",
    "load_kv_map",
]
