"""One-process-per-GPU plumbing around the engine's ghost exchange.

The device side (pack, NCCL all-to-all-v, unpack) lives in csrc/comm.cu.  What the host has to do
once at start-up is what the reference does per row, per message, forever
(ghostReceiverGCN, engine/ops/gcn_ops.cpp:310-318: ``globalToGhostVtcs[gvid] - localVtxCnt``):
translate the global ids a peer will send into this partition's ghost slots.  ``GhostPlan`` does that
with numpy; the id lists travel over whatever ``torch.distributed`` backend is up (NCCL on the GPU
box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .formats import PartitionGraph

FORWARD, BACKWARD = 0, 1


def recv_slots(ghost_gvids_sorted: np.ndarray, incoming_gvids: np.ndarray) -> np.ndarray:
    """Ghost slots (0-based inside the fg / bg block) for the rows a peer sends, in its send order.
    Ghost slots ascend by global id (graph/dataloader.cpp:311-322), so this is a binary search."""
    slots = np.searchsorted(ghost_gvids_sorted, incoming_gvids)
    if incoming_gvids.size and (slots.max() >= ghost_gvids_sorted.size or
                                not np.array_equal(ghost_gvids_sorted[slots], incoming_gvids)):
        raise ValueError("peer sends a vertex that is not a ghost of this partition")
    return slots.astype(np.uint32)


class GhostPlan:
    """Per (direction, peer): which local rows go out and which ghost slots come in."""

    def __init__(self, graph: PartitionGraph, rank: int, world: int):
        self.graph, self.rank, self.world = graph, rank, world
        self.send_ids = {FORWARD: graph.fwd_send, BACKWARD: graph.bwd_send}
        self.recv = {FORWARD: [np.zeros(0, np.uint32)] * world, BACKWARD: [np.zeros(0, np.uint32)] * world}

    def send_gvids(self, dir: int) -> List[np.ndarray]:
        """Global ids of the rows shipped to each peer (what verticesPushOut writes in front of every
        row, engine/utils.cpp:642-644)."""
        return [self.graph.local_to_global[ids] for ids in self.send_ids[dir]]

    def set_incoming(self, dir: int, peer: int, gvids: np.ndarray):
        ghosts = self.graph.src_ghost_gvid if dir == FORWARD else self.graph.dst_ghost_gvid
        self.recv[dir][peer] = recv_slots(ghosts, np.asarray(gvids, dtype=np.uint32))

    def complete(self) -> bool:
        """Every ghost slot is filled by exactly one peer."""
        for dir, n in ((FORWARD, self.graph.src_ghost_cnt), (BACKWARD, self.graph.dst_ghost_cnt)):
            got = np.concatenate(self.recv[dir]) if self.world else np.zeros(0, np.uint32)
            if got.size != n or (n and not np.array_equal(np.sort(got), np.arange(n, dtype=np.uint32))):
                return False
        return True

    def exchange_over(self, group=None):
        """Swap the id lists with every peer through torch.distributed (object collectives)."""
        import torch.distributed as dist

        for dir in (FORWARD, BACKWARD):
            mine = self.send_gvids(dir)
            everyone: List[Optional[list]] = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            for peer in range(self.world):
                if peer != self.rank:
                    self.set_incoming(dir, peer, everyone[peer][self.rank])
        if not self.complete():
            raise RuntimeError("ghost plan incomplete: partitions disagree about the edge cut")


def setup_engine_comm(engine, graph: PartitionGraph, rank: int, world: int, group=None,
                      peer_memory: bool = True) -> GhostPlan:
    """Create the engine's NCCL communicator and install the receive plan (GPU ranks only).
    With `peer_memory` the ghost exchange runs as one store-through-NVLink kernel (csrc/comm.cu:
    exchange_p2p); without it as pack -> NCCL all-to-all-v -> unpack."""
    import torch.distributed as dist

    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(box[0])
    plan = GhostPlan(graph, rank, world)
    plan.exchange_over(group)
    for dir in (FORWARD, BACKWARD):
        for peer in range(world):
            if peer != rank:
                engine.comm_set_recv_slots(dir, peer, plan.recv[dir][peer])
    if peer_memory:
        setup_peer_memory(engine, plan, rank, world, group)
    return plan


def setup_peer_memory(engine, plan: GhostPlan, rank: int, world: int, group=None):
    """Switch the exchanges to the peer-memory path: tell every sender where its rows land on the
    receiver (the receiver's slot list, mirrored) and map every peer's ghost blocks (CUDA IPC)."""
    import torch.distributed as dist

    for dir in (FORWARD, BACKWARD):
        everyone: List[Optional[list]] = [None] * world
        dist.all_gather_object(everyone, plan.recv[dir], group=group)
        for peer in range(world):
            if peer != rank:
                engine.comm_set_send_slots(dir, peer, everyone[peer][rank])  # where peer stores MY rows
    mine = {key: engine.comm_ipc_export(*key) for key in engine.ghost_tensors()}
    blobs: List[Optional[dict]] = [None] * world
    dist.all_gather_object(blobs, mine, group=group)
    for peer in range(world):
        if peer == rank:
            continue
        for (layer, name), blob in blobs[peer].items():
            engine.comm_ipc_import(layer, name, peer, blob)


def host_exchange_rows(plan: GhostPlan, dir: int, local_rows: np.ndarray, ghost_rows: np.ndarray, group=None):
    """CPU statement of one Scatter step (used by the gloo tests with the oracle as compute):
    rows listed in the send lists go to each peer and land in its ghost block."""
    import torch
    import torch.distributed as dist

    ops, bufs = [], []
    width = local_rows.shape[1]
    for peer in range(plan.world):
        if peer == plan.rank:
            continue
        ids = plan.send_ids[dir][peer]
        if ids.size:
            t = torch.from_numpy(np.ascontiguousarray(local_rows[ids]))
            ops.append(dist.P2POp(dist.isend, t, peer, group=group))
        n = plan.recv[dir][peer].size
        if n:
            r = torch.empty((n, width), dtype=torch.float32)
            bufs.append((peer, r))
            ops.append(dist.P2POp(dist.irecv, r, peer, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for peer, r in bufs:
        ghost_rows[plan.recv[dir][peer]] = r.numpy()
