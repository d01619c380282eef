"""Scaling the store across the GPUs of an NVSwitch box.

The reference has no parallelism of its own: one server process, one pinned-DRAM pool
(SURVEY §2.2).  The first-class axes here are
  * pool sharding   - every GPU (rank) may host a pool shard; ``ShardedConnection`` routes
                      keys to shards by hash, so N control planes and N HBM pools work in
                      parallel and traffic spreads over the switch;
  * SPMD helpers    - one process per GPU under torchrun: ``start_shard_server``,
                      ``ring_peer``, ``connect_all``;
  * NVLS broadcast  - ``PrefixBroadcaster``: one writer, every GPU gets a replica through a
                      multicast mapping (kernels/kv_bcast_nvls.cu).
"""
from .sharded import ShardedConnection, shard_of
from .spmd import start_shard_server, ring_peer, connect_all, shard_port
from .nvls import PrefixBroadcaster, nvls_available

__all__ = ["ShardedConnection", "shard_of", "start_shard_server", "ring_peer", "connect_all",
           "shard_port", "PrefixBroadcaster", "nvls_available"]
