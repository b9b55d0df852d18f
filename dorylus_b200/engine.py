"""Host-side mirror of the reference's Engine / ResourceComm interface over the C ABI.

Method names, argument meaning and the (layer, name) tensor addressing follow the reference
(src/graph-server/engine/engine.hpp:84-94, commmanager/resource_comm.hpp:13-28) so that the parity
tests read like calls into the reference; every method is a thin wrapper over one ``dory_*`` entry
point of include/dorylus_b200.h -- no arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (BACKWARD, FORWARD, GAT, GCN, DoryChunk, DoryConfig, DoryError, DoryStats)

_f32p = C.POINTER(C.c_float)


class Chunk:
    """== struct Chunk (src/common/utils.hpp:64-105)."""

    __slots__ = ("localId", "globalId", "lowBound", "upBound", "layer", "dir", "epoch", "vertex")

    def __init__(self, localId=0, globalId=0, lowBound=0, upBound=0, layer=0, dir=FORWARD, epoch=0, vertex=True):
        self.localId, self.globalId, self.lowBound, self.upBound = localId, globalId, lowBound, upBound
        self.layer, self.dir, self.epoch, self.vertex = layer, dir, epoch, vertex

    def c(self) -> DoryChunk:
        return DoryChunk(self.localId, self.globalId, self.lowBound, self.upBound, self.layer, self.dir,
                         self.epoch, 1 if self.vertex else 0)

    def copy(self) -> "Chunk":
        return Chunk(self.localId, self.globalId, self.lowBound, self.upBound, self.layer, self.dir,
                     self.epoch, self.vertex)

    def isFirstLayer(self) -> bool:
        return self.dir == FORWARD and self.layer == 0 and self.vertex

    def isLastLayer(self) -> bool:
        return self.dir == BACKWARD and self.layer == 0 and self.vertex

    def __repr__(self):
        return "%u:%s:%u:%u/%u: vtx %u" % (self.epoch, "F" if self.dir == FORWARD else "B", self.layer,
                                          self.localId, self.globalId, int(self.vertex))


def preprocess_edges(src: np.ndarray, dst: np.ndarray, parts: np.ndarray, num_vertices: int, part: int,
                     num_parts: int, undirected: bool = False) -> bytes:
    """== DataLoader::preprocess on an in-memory edge list; returns the graph.<part>.bin image."""
    lib = _lib.load()
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    parts = np.ascontiguousarray(parts, dtype=np.int32)
    if parts.size != num_vertices:
        raise ValueError("parts must have one entry per global vertex")
    img, n = C.c_void_p(), C.c_size_t()
    rc = lib.dory_preprocess_edges(src.ctypes.data_as(C.POINTER(C.c_uint32)), dst.ctypes.data_as(C.POINTER(C.c_uint32)),
                                   src.size, parts.ctypes.data_as(C.POINTER(C.c_int32)), num_vertices, part,
                                   num_parts, int(undirected), C.byref(img), C.byref(n))
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    return _take_image(lib, img, n.value)


def preprocess_incident_edges(src: np.ndarray, dst: np.ndarray, parts: np.ndarray, num_vertices: int, part: int,
                              num_parts: int, in_degree: np.ndarray, global_edges: int):
    """== dory_preprocess_incident_edges: the graph.<part>.bin image from the edge records incident to
    `part` alone, the whole graph's in-degrees and its record count."""
    lib = _lib.load()
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    parts = np.ascontiguousarray(parts, dtype=np.int32)
    in_degree = np.ascontiguousarray(in_degree, dtype=np.uint32)
    if parts.size != num_vertices or in_degree.size != num_vertices:
        raise ValueError("parts and in_degree must have one entry per global vertex")
    img, n = C.c_void_p(), C.c_size_t()
    u32p = C.POINTER(C.c_uint32)
    rc = lib.dory_preprocess_incident_edges(src.ctypes.data_as(u32p), dst.ctypes.data_as(u32p), src.size,
                                            parts.ctypes.data_as(C.POINTER(C.c_int32)), num_vertices, part, num_parts,
                                            in_degree.ctypes.data_as(u32p), int(global_edges), C.byref(img), C.byref(n))
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    return _take_image(lib, img, n.value)


_BYTES_LIMIT = 1 << 31  # larger images are handed out as arrays (tests lower it)


class _OwnedImage:
    """Keeps a dory_free()-able allocation alive for the numpy view that wraps it."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        if self.ptr:
            self.lib.dory_free(self.ptr)
            self.ptr = None


def _take_image(lib, img, n: int):
    """bytes for small images; beyond 2 GiB (one eighth of the Friendster shape is 4.1 GB) a uint8 array
    that wraps the library's allocation without copying it (load_partition / parse_graph_bin take both)."""
    if n < _BYTES_LIMIT:
        try:
            return C.string_at(img, n)
        finally:
            lib.dory_free(img)
    arr = np.ctypeslib.as_array((C.c_ubyte * n).from_address(img.value))
    owner = _OwnedImage(lib, C.c_void_p(img.value))
    view = arr.view()
    view.flags.writeable = False
    return _ImageArray(view, owner)


class _ImageArray(np.ndarray):
    def __new__(cls, arr, owner):
        obj = np.asarray(arr).view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


def preprocess_dir(dataset_dir: str, part: int, num_parts: int, undirected: bool = False) -> str:
    lib = _lib.load()
    d = dataset_dir if dataset_dir.endswith("/") else dataset_dir + "/"
    rc = lib.dory_preprocess_dir(d.encode(), part, num_parts, int(undirected))
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    return d + "graph.%d.bin" % part


def _image_buffer(image):
    if isinstance(image, np.ndarray):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        return image, image.ctypes.data_as(C.c_void_p), image.size
    buf = (C.c_char * len(image)).from_buffer_copy(image)
    return buf, C.cast(buf, C.c_void_p), len(image)


def read_features(dataset_dir: str, features_file: str, image, node_id: int, num_features: int):
    """== Engine::readFeaturesFile for the partition `image` (graph.<id>.bin bytes): (local rows,
    source-ghost rows); reads / writes the reference's feats<F0>.<id>.bin cache.  Host only."""
    from .formats import parse_graph_bin

    lib = _lib.load()
    g = parse_graph_bin(image)
    keep, ptr, n = _image_buffer(image)
    local = np.empty((g.local_vtx_cnt, num_features), dtype=np.float32)
    ghost = np.empty((g.src_ghost_cnt, num_features), dtype=np.float32)
    d = dataset_dir if dataset_dir.endswith("/") else dataset_dir + "/"
    rc = lib.dory_read_features(d.encode(), features_file.encode(), ptr, n, node_id, num_features,
                                local.ctypes.data_as(_f32p), ghost.ctypes.data_as(_f32p) if ghost.size else None)
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    return local, ghost


def read_labels(labels_file: str, image, kinds: int) -> np.ndarray:
    """== Engine::readLabelsFile: one-hot [V_p x kinds] for the partition `image`.  Host only."""
    from .formats import parse_graph_bin

    lib = _lib.load()
    g = parse_graph_bin(image)
    keep, ptr, n = _image_buffer(image)
    onehot = np.empty((g.local_vtx_cnt, kinds), dtype=np.float32)
    rc = lib.dory_read_labels(labels_file.encode(), ptr, n, kinds, onehot.ctypes.data_as(_f32p))
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    return onehot


def ghost_slots(my_image, my_id: int, peer_image, dir: int) -> np.ndarray:
    """== dory_ghost_slots: ghost slots (in partition `my_id`'s fg / bg block) of the rows the peer
    partition sends it in direction `dir`, in the peer's send order -- computed from the two
    graph.<id>.bin images alone (host only)."""
    lib = _lib.load()
    _mk, mp, ml = _image_buffer(my_image)
    _pk, pp, pl = _image_buffer(peer_image)
    n = C.c_uint32(0)
    rc = lib.dory_ghost_slots(mp, ml, my_id, pp, pl, dir, None, C.byref(n))
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    out = np.zeros(max(int(n.value), 1), np.uint32)
    rc = lib.dory_ghost_slots(mp, ml, my_id, pp, pl, dir, out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n))
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    return out[:int(n.value)]


def partition_edges(src: np.ndarray, dst: np.ndarray, num_vertices: int, num_parts: int, passes: int = 0):
    """== inputs/partitioner.cpp without METIS: owner per global vertex (int32) and the edge cut
    (records whose endpoints have different owners).  Host only."""
    lib = _lib.load()
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    parts = np.empty(num_vertices, dtype=np.int32)
    cut = C.c_uint64()
    rc = lib.dory_partition_edges(src.ctypes.data_as(C.POINTER(C.c_uint32)), dst.ctypes.data_as(C.POINTER(C.c_uint32)),
                                  src.size, num_vertices, num_parts, passes,
                                  parts.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(cut))
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    return parts, int(cut.value)


def partition_file(bsnap_path: str, num_parts: int, out_dir: str) -> str:
    """Reads graph.bsnap, writes <out_dir>/<name>.parts and .comm like the reference's partitioner."""
    lib = _lib.load()
    rc = lib.dory_partition_file(bsnap_path.encode(), num_parts, out_dir.encode())
    if rc != 0:
        raise DoryError(rc, lib.dory_last_error(None).decode())
    return os.path.join(out_dir, os.path.basename(bsnap_path) + ".parts")


class Engine:
    """One partition on one B200.  Mirrors ``Engine`` + the ``ResourceComm`` backend it owns."""

    def __init__(self, dims: Sequence[int], gnn_type: int = GCN, node_id: int = 0, num_nodes: int = 1,
                 device: int = 0, learning_rate: float = 0.01, flags: int = 0):
        self._lib = _lib.load()
        cfg = DoryConfig()
        cfg.abi_version = _lib.DORY_ABI_VERSION
        cfg.gnn_type = gnn_type
        cfg.n_layers = len(dims) - 1
        for i, d in enumerate(dims):
            cfg.dims[i] = int(d)
        cfg.node_id, cfg.num_nodes, cfg.device = node_id, num_nodes, device
        cfg.learning_rate, cfg.flags = learning_rate, flags
        self.layerConfig = list(dims)
        self.numLayers = len(dims) - 1
        self.gnn_type, self.nodeId, self.numNodes = gnn_type, node_id, num_nodes
        h = C.c_void_p()
        rc = self._lib.dory_create(C.byref(h), C.byref(cfg))
        if rc != 0:
            raise DoryError(rc, self._lib.dory_last_error(None).decode())
        self._h = h
        self._image = None

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != 0:
            raise DoryError(rc, self._lib.dory_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dory_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        self._check(self._lib.dory_sync(self._h))

    def set_option(self, key: str, value):
        self._check(self._lib.dory_set_option(self._h, key.encode(), str(value).encode()))

    # ------------------------------------------------------------------ graph + tensors
    def load_partition(self, image: bytes):
        """== Graph::init + preallocateGCN/GAT."""
        if isinstance(image, np.ndarray):
            ptr, n = C.c_void_p(image.ctypes.data), image.nbytes
        else:
            ptr, n = bytes(image) if not isinstance(image, bytes) else image, len(image)
        self._check(self._lib.dory_load_partition(self._h, ptr, n))
        cnt = (C.c_uint64 * 7)()
        self._check(self._lib.dory_graph_counts(self._h, cnt))
        (self.localVtxCnt, self.globalVtxCnt, self.srcGhostCnt, self.dstGhostCnt, self.localInEdgeCnt,
         self.localOutEdgeCnt, self.globalEdgeCnt) = (int(x) for x in cnt)

    def tensor_shape(self, layer: int, name: str):
        r, c = C.c_uint64(), C.c_uint32()
        self._check_name(self._lib.dory_tensor_shape(self._h, layer, name.encode(), C.byref(r), C.byref(c)), layer, name)
        return int(r.value), int(c.value)

    def _check_name(self, rc, layer, name):
        if rc != 0:
            raise DoryError(rc, "no tensor '%s' at layer %d" % (name, layer))

    def set_tensor(self, layer: int, name: str, host: np.ndarray):
        """savedNNTensors[layer][name] <- host."""
        host = np.ascontiguousarray(host, dtype=np.float32)
        if host.ndim == 1:
            host = host.reshape(-1, 1)
        self._check(self._lib.dory_set_tensor(self._h, layer, name.encode(), host.ctypes.data_as(_f32p),
                                              host.shape[0], host.shape[1]))

    def prefetch_tensor(self, layer: int, name: str, host: np.ndarray):
        """Start the host->device DMA of a tensor on the copy stream (input pipeline).  `host` must be
        a C-contiguous float32 array (pinned for a truly asynchronous copy) that the caller keeps alive
        and unchanged until commit_prefetch() + sync()."""
        if host.dtype != np.float32 or not host.flags.c_contiguous:
            raise ValueError("prefetch_tensor needs a C-contiguous float32 array (no implicit copies)")
        if host.ndim == 1:
            host = host.reshape(-1, 1)
        self._check(self._lib.dory_prefetch_tensor(self._h, layer, name.encode(), host.ctypes.data_as(_f32p),
                                                   host.shape[0], host.shape[1]))

    def commit_prefetch(self):
        self._check(self._lib.dory_commit_prefetch(self._h))

    def get_tensor(self, layer: int, name: str) -> np.ndarray:
        """savedNNTensors[layer][name] -> host (synchronises)."""
        r, c = self.tensor_shape(layer, name)
        out = np.empty((r, c), dtype=np.float32)
        if r:
            self._check(self._lib.dory_get_tensor(self._h, layer, name.encode(), out.ctypes.data_as(_f32p), r, c))
        return out

    def tensor_device(self, layer: int, name: str):
        p, r, c, ld = C.c_void_p(), C.c_uint64(), C.c_uint32(), C.c_uint32()
        self._check_name(self._lib.dory_tensor_device(self._h, layer, name.encode(), C.byref(p), C.byref(r),
                                                      C.byref(c), C.byref(ld)), layer, name)
        return p.value, int(r.value), int(c.value), int(ld.value)

    # ------------------------------------------------------------------ weights
    def init_weights(self):
        self._check(self._lib.dory_init_weights(self._h))

    def weight_shape(self, layer: int, name: str = "w"):
        return (self.layerConfig[layer], self.layerConfig[layer + 1]) if name == "w" else (self.layerConfig[layer + 1], 1)

    def set_weights(self, layer: int, host: np.ndarray, name: str = "w"):
        host = np.ascontiguousarray(host, dtype=np.float32)
        if host.ndim == 1:
            host = host.reshape(-1, 1)
        self._check(self._lib.dory_set_weights(self._h, layer, name.encode(), host.ctypes.data_as(_f32p),
                                               host.shape[0], host.shape[1]))

    def get_weights(self, layer: int, name: str = "w") -> np.ndarray:
        r, c = self.weight_shape(layer, name)
        out = np.empty((r, c), dtype=np.float32)
        self._check(self._lib.dory_get_weights(self._h, layer, name.encode(), out.ctypes.data_as(_f32p), r, c))
        return out

    def get_weight_grad(self, layer: int, name: str = "w") -> np.ndarray:
        r, c = self.weight_shape(layer, name)
        out = np.empty((r, c), dtype=np.float32)
        self._check(self._lib.dory_get_weight_grad(self._h, layer, name.encode(), out.ctypes.data_as(_f32p), r, c))
        return out

    def apply_update(self, layer: int):
        self._check(self._lib.dory_apply_update(self._h, layer))

    # ------------------------------------------------------------------ SAGA operators (reference names)
    def _op(self, fn, chunk: Chunk):
        c = chunk.c()
        self._check(fn(self._h, C.byref(c)))

    def aggregate(self, chunk: Chunk):
        self._op(self._lib.dory_aggregate, chunk)

    def applyVertex(self, chunk: Chunk):
        self._op(self._lib.dory_apply_vertex, chunk)

    def scatter(self, chunk: Chunk):
        self._op(self._lib.dory_scatter, chunk)

    def applyEdge(self, chunk: Chunk):
        self._op(self._lib.dory_apply_edge, chunk)

    def predictGAT(self, chunk: Chunk):
        self._op(self._lib.dory_predict, chunk)

    aggregateGCN = aggregateGAT = aggregate
    applyVertexGCN = applyVertexGAT = applyVertex
    scatterGCN = scatterGAT = scatter
    applyEdgeGCN = applyEdgeGAT = applyEdge

    def incLayer(self, chunk: Chunk) -> Chunk:
        """== incLayerGCN / incLayerGAT: returns the next chunk."""
        c = chunk.c()
        self._check(self._lib.dory_inc_layer(self._h, C.byref(c)))
        return Chunk(c.localId, c.globalId, c.lowBound, c.upBound, c.layer, c.dir, c.epoch, bool(c.vertex))

    incLayerGCN = incLayerGAT = incLayer

    def whole_chunk(self, layer: int = 0, dir: int = FORWARD, epoch: int = 1, vertex: bool = True) -> Chunk:
        """The single chunk CPU/GPU mode uses per partition (loadChunks, engine/utils.cpp:598-609)."""
        return Chunk(0, self.nodeId, 0, self.localVtxCnt, layer, dir, epoch, vertex)

    # ------------------------------------------------------------------ coarse entry points
    def forward(self, layer: int):
        self._check(self._lib.dory_forward(self._h, layer))

    def backward(self, layer: int):
        self._check(self._lib.dory_backward(self._h, layer))

    def epoch(self) -> dict:
        s = DoryStats()
        self._check(self._lib.dory_epoch(self._h, C.byref(s)))
        return self._stats(s)

    def epoch_async(self):
        """Enqueue one epoch without reading the statistics back (no synchronisation)."""
        self._check(self._lib.dory_epoch(self._h, None))

    def stats_enqueue(self, slot: int = 0):
        """Start the device->host copy of the statistics behind everything enqueued so far."""
        self._check(self._lib.dory_stats_enqueue(self._h, slot))

    def stats_collect(self, slot: int = 0) -> dict:
        """Wait for that copy (not for later work) and return it."""
        s = DoryStats()
        self._check(self._lib.dory_stats_collect(self._h, slot, C.byref(s)))
        return self._stats(s)

    def stats(self) -> dict:
        s = DoryStats()
        self._check(self._lib.dory_get_stats(self._h, C.byref(s)))
        return self._stats(s)

    @staticmethod
    def _stats(s: DoryStats) -> dict:
        return dict(acc_sum=s.acc_sum, loss_sum=s.loss_sum, val_rows=s.val_rows, epochs_done=s.epochs_done,
                    kernel_launches=int(s.kernel_launches), edges_aggregated=int(s.edges_aggregated))

    # ------------------------------------------------------------------ multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(_lib.DORY_UNIQUE_ID_BYTES)
        rc = lib.dory_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise DoryError(rc, lib.dory_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, _lib.DORY_UNIQUE_ID_BYTES)
        self._check(self._lib.dory_comm_init(self._h, C.cast(buf, C.c_void_p)))

    def comm_send_gvids(self, dir: int, peer: int) -> np.ndarray:
        n = C.c_uint32()
        self._check(self._lib.dory_comm_send_gvids(self._h, dir, peer, None, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint32)
        self._check(self._lib.dory_comm_send_gvids(self._h, dir, peer, out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n)))
        return out[: n.value]

    def comm_set_recv_slots(self, dir: int, peer: int, slots: np.ndarray):
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        self._check(self._lib.dory_comm_set_recv_slots(self._h, dir, peer, slots.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                       slots.size))

    def comm_set_send_slots(self, dir: int, peer: int, slots: np.ndarray):
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        self._check(self._lib.dory_comm_set_send_slots(self._h, dir, peer, slots.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                       slots.size))

    def tile_info(self, dir: int = FORWARD) -> dict:
        """== dory_tile_info: the shared-memory-staged aggregation's plan for one adjacency."""
        cov, w, r, n = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._lib.dory_tile_info(self._h, dir, C.byref(cov), C.byref(w), C.byref(r), C.byref(n)))
        return dict(coverage=float(cov.value), window_rows=int(w.value), tile_rows=int(r.value), n_tiles=int(n.value))

    def apply_first(self, layer: int) -> bool:
        """True when `layer` runs the apply-first schedule (DORY_FLAG_APPLY_FIRST, dory_layer_schedule)."""
        v = C.c_int(0)
        self._check(self._lib.dory_layer_schedule(self._h, layer, C.byref(v)))
        return bool(v.value)

    def ghost_tensors(self):
        """(layer, name) of every ghost block that takes part in an exchange."""
        L = self.numLayers
        if self.gnn_type == GCN:
            out = [(0, "fg")]  # the layer-0 input exchange (scatter of a layer-0 FORWARD chunk)
            for l in range(L):
                if self.apply_first(l):  # ghosts of t = in . W (forward) and of dL/dz (backward)
                    out += [(l, "fg_t"), (l, "bg_g")]
                elif l > 0:
                    out += [(l, "fg"), (l - 1, "bg")]
            return out
        return [(l, "fg_z") for l in range(L)] + [(l, "bg_d") for l in range(L)]

    def comm_ipc_export(self, layer: int, name: str) -> bytes:
        buf = C.create_string_buffer(_lib.DORY_IPC_BLOB_BYTES)
        self._check(self._lib.dory_comm_ipc_export(self._h, layer, name.encode(), C.cast(buf, C.c_void_p)))
        return buf.raw

    def comm_ipc_import(self, layer: int, name: str, peer: int, blob: bytes):
        buf = C.create_string_buffer(blob, _lib.DORY_IPC_BLOB_BYTES)
        self._check(self._lib.dory_comm_ipc_import(self._h, layer, name.encode(), peer, C.cast(buf, C.c_void_p)))

    # ------------------------------------------------------------------ timing helpers
    def event_record(self, slot: int):
        self._check(self._lib.dory_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._check(self._lib.dory_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def measure_fma_peak(self) -> float:
        """Non-tensor fp32 FMA throughput of this GPU in TFLOP/s (register-only FMA loop)."""
        t = C.c_float()
        self._check(self._lib.dory_measure_fma_peak(self._h, C.byref(t)))
        return float(t.value)

    def flush_l2(self, nbytes: int = 256 << 20):
        self._check(self._lib.dory_flush_l2(self._h, nbytes))
