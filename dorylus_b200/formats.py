"""On-disk formats of the reference's dataset directory (little-endian, numpy only).

Layouts follow SURVEY.md §10:

* ``graph.bsnap`` / ``graph.bsnap.edges`` -- ``BSHeaderType {int32 sizeOfVertexType; uint32 numVertices;
  uint64 numEdges}`` then ``numEdges x {uint32 src, uint32 dst}``
  (reference inputs/graphToBinary.cpp:15-20, graph/dataloader.hpp:11-15).
* ``graph.bsnap.parts`` -- text, one partition id per line (graph/dataloader.cpp:53-87).
* ``features.bsnap`` -- ``uint32 numFeatures`` then ``V x numFeatures`` fp32 (engine/utils.cpp:508-535).
* ``labels.bsnap`` -- ``uint32 labelKinds`` then ``V x uint32`` (engine/utils.cpp:567-595).
* ``graph.<id>.bin`` -- the preprocessed partition written by ``RawGraph::dump``
  (graph/graph.cpp:200-273) and read by ``Graph::init`` (graph/graph.cpp:7-115).
* ``feats<F0>.<id>.bin`` -- local rows then src-ghost rows (engine/utils.cpp:487-501).
* layer config -- text, one width per line (engine/utils.cpp:460-479).
* text inputs -- the converters of the reference's inputs/ directory (graphToBinary.cpp,
  featuresToBinary.cpp, labelsToBinary.cpp) as ``convert_*_text``.
"""
from __future__ import annotations

import dataclasses
import os
import struct
from typing import List

import numpy as np

BS_HEADER = struct.Struct("<iIQ")  # sizeOfVertexType, numVertices, numEdges


# --------------------------------------------------------------------------- raw dataset files
def write_bsnap_edges(path: str, num_vertices: int, src: np.ndarray, dst: np.ndarray) -> None:
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    assert src.shape == dst.shape and src.ndim == 1
    pairs = np.empty((src.size, 2), dtype=np.uint32)
    pairs[:, 0] = src
    pairs[:, 1] = dst
    with open(path, "wb") as f:
        f.write(BS_HEADER.pack(4, int(num_vertices), int(src.size)))
        pairs.tofile(f)


def read_bsnap_edges(path: str):
    with open(path, "rb") as f:
        sz, nv, ne = BS_HEADER.unpack(f.read(BS_HEADER.size))
        if sz != 4:
            raise ValueError("unsupported sizeOfVertexType %d" % sz)
        pairs = np.fromfile(f, dtype=np.uint32).reshape(-1, 2)
    return nv, pairs[:, 0].copy(), pairs[:, 1].copy()


def write_parts(path: str, parts: np.ndarray) -> None:
    with open(path, "w") as f:
        f.write("\n".join(str(int(p)) for p in parts))
        f.write("\n")


def read_parts(path: str) -> np.ndarray:
    out = []
    with open(path) as f:
        for line in f:
            if not line or not ("0" <= line[0] <= "9"):  # dataloader.cpp:67
                continue
            out.append(int(line.split()[0]))
    return np.asarray(out, dtype=np.int32)


def write_features(path: str, feats: np.ndarray) -> None:
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", feats.shape[1]))
        feats.tofile(f)


def read_features(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        (nf,) = struct.unpack("<I", f.read(4))
        return np.fromfile(f, dtype=np.float32).reshape(-1, nf)


def write_labels(path: str, labels: np.ndarray, kinds: int) -> None:
    labels = np.ascontiguousarray(labels, dtype=np.uint32)
    with open(path, "wb") as f:
        f.write(struct.pack("<I", int(kinds)))
        labels.tofile(f)


def read_labels(path: str):
    with open(path, "rb") as f:
        (kinds,) = struct.unpack("<I", f.read(4))
        return kinds, np.fromfile(f, dtype=np.uint32)


def write_layer_config(path: str, dims: List[int]) -> None:
    with open(path, "w") as f:
        for d in dims:
            f.write("%d\n" % d)


def read_layer_config(path: str) -> List[int]:
    with open(path) as f:
        return [int(line.strip()) for line in f if line.strip()]


# --------------------------------------------------------------------------- text -> binary converters
def convert_graph_text(snap_path: str, out_path: str, undirected: bool = False, with_header: bool = True):
    """== inputs/graphToBinary.cpp: a text edge list ("src dst" per line, lines starting with '#' or
    '%' skipped, reading stops at the first line that does not parse) -> graph.bsnap.  Self loops
    are dropped; with `undirected` every record is followed by its reverse and the header's edge
    count is doubled too (graphToBinary.cpp:36-53 counts the non-loop lines, main() :151 doubles).
    Byte-identical to the reference tool's output (tests/test_formats.py runs it).
    Returns (numVertices, numEdges as written in the header)."""
    src, dst = [], []
    with open(snap_path) as f:
        for line in f:
            if line[:1] in ("#", "%"):
                continue
            tok = line.split()
            try:
                a, b = int(tok[0]), int(tok[1])
            except (IndexError, ValueError):
                break  # `if (!(iss >> src >> dst)) break;`
            if a < 0 or b < 0:
                break
            if a == b:
                continue
            src.append(a)
            dst.append(b)
    s = np.asarray(src, dtype=np.uint32)
    d = np.asarray(dst, dtype=np.uint32)
    nv = int(max(s.max(initial=0), d.max(initial=0))) + 1
    ne = int(s.size) * (2 if undirected else 1)
    rec = np.empty((s.size, 4 if undirected else 2), dtype=np.uint32)
    rec[:, 0], rec[:, 1] = s, d
    if undirected:
        rec[:, 2], rec[:, 3] = d, s
    with open(out_path, "wb") as f:
        if with_header:
            f.write(BS_HEADER.pack(4, nv, ne))
        rec.tofile(f)
    return nv, ne


def _leading_digit(line: str) -> bool:
    return bool(line) and "0" <= line[0] <= "9"


def convert_features_text(path: str, num_features: int):
    """== inputs/featuresToBinary.cpp: one comma / space separated row per line -> <path>.bsnap.
    Faithful to the reference's filter: a (trimmed) line whose first character is not a digit is
    SKIPPED -- that includes rows whose first value is negative (featuresToBinary.cpp:51-52).
    Returns (rows written, lines skipped)."""
    import re

    rows, skipped = [], 0
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not _leading_digit(line):
                skipped += bool(line)
                continue
            vals = [v for v in re.split(r"[, ]+", line) if v != ""]
            if len(vals) != num_features:
                raise ValueError("row with %d values, header says %d" % (len(vals), num_features))
            rows.append(np.asarray([float(v) for v in vals], dtype=np.float32))
    feats = np.stack(rows) if rows else np.zeros((0, num_features), np.float32)
    write_features(path + ".bsnap", feats)
    return feats.shape[0], skipped


def convert_labels_text(path: str, label_kinds: int):
    """== inputs/labelsToBinary.cpp: one class id per line -> <path>.bsnap (same leading-digit filter)."""
    out, skipped = [], 0
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if not _leading_digit(line):
                skipped += 1
                continue
            m = 0
            while m < len(line) and line[m].isdigit():  # std::stoul stops at the first non-digit
                m += 1
            out.append(int(line[:m]))
    write_labels(path + ".bsnap", np.asarray(out, dtype=np.uint32), label_kinds)
    return len(out), skipped


def one_hot(labels: np.ndarray, kinds: int) -> np.ndarray:
    """readLabelsFile stores labels one-hot as fp32 (engine/utils.cpp:575-588)."""
    out = np.zeros((labels.size, kinds), dtype=np.float32)
    out[np.arange(labels.size), labels.astype(np.int64)] = 1.0
    return out


# --------------------------------------------------------------------------- graph.<id>.bin
@dataclasses.dataclass
class PartitionGraph:
    """Everything ``Graph::init`` reads from ``graph.<id>.bin`` (graph/graph.cpp:7-115)."""

    local_vtx_cnt: int
    global_vtx_cnt: int
    src_ghost_cnt: int
    dst_ghost_cnt: int
    local_in_edge_cnt: int
    local_out_edge_cnt: int
    global_edge_cnt: int
    local_to_global: np.ndarray  # u32[V_p]
    norms: np.ndarray  # f32[V_p]   vtxDataVec
    src_ghost_gvid: np.ndarray  # u32[Gs] ascending; slot k lives at local id V_p + k
    src_ghost_lvid: np.ndarray
    dst_ghost_gvid: np.ndarray
    dst_ghost_lvid: np.ndarray
    num_nodes: int
    fwd_send: List[np.ndarray]  # forwardLocalVtxDsts[peer]  (local ids, ascending)
    bwd_send: List[np.ndarray]  # backwardLocalVtxDsts[peer]
    col_ptrs: np.ndarray  # u64[V_p+1]   forwardAdj (CSC)
    row_idxs: np.ndarray  # u32[E_in]
    fwd_vals: np.ndarray  # f32[E_in]
    row_ptrs: np.ndarray  # u64[V_p+1]   backwardAdj (CSR)
    col_idxs: np.ndarray  # u32[E_out]
    bwd_vals: np.ndarray  # f32[E_out]


def parse_graph_bin(buf: bytes) -> PartitionGraph:
    mv = memoryview(buf)
    off = 0

    def take(dtype, n):
        nonlocal off
        dt = np.dtype(dtype)
        a = np.frombuffer(mv, dtype=dt, count=n, offset=off)
        off += dt.itemsize * n
        return a

    lv, gv, sg, dg = (int(x) for x in take("<u4", 4))
    ine, oute, ge = (int(x) for x in take("<u8", 3))
    l2g = take("<u4", lv).copy()
    norms = take("<f4", lv).copy()
    sgp = take("<u4", 2 * sg).reshape(-1, 2)
    dgp = take("<u4", 2 * dg).reshape(-1, 2)
    (nn,) = (int(x) for x in take("<u4", 1))
    fwd, bwd = [], []
    for lst in (fwd, bwd):
        for _ in range(nn):
            (sz,) = (int(x) for x in take("<u4", 1))
            lst.append(take("<u4", sz).copy())
    (ccnt,) = (int(x) for x in take("<u4", 1))
    (nnz_f,) = (int(x) for x in take("<u8", 1))
    fvals = take("<f4", nnz_f).copy()
    colp = take("<u8", lv + 1).copy()
    rowi = take("<u4", nnz_f).copy()
    (rcnt,) = (int(x) for x in take("<u4", 1))
    (nnz_b,) = (int(x) for x in take("<u8", 1))
    bvals = take("<f4", nnz_b).copy()
    rowp = take("<u8", lv + 1).copy()
    coli = take("<u4", nnz_b).copy()
    if off != len(buf):
        raise ValueError("graph.bin: %d trailing bytes" % (len(buf) - off))
    if ccnt != lv or rcnt != lv:
        raise ValueError("graph.bin: CSC/CSR vertex count mismatch")
    return PartitionGraph(lv, gv, sg, dg, ine, oute, ge, l2g, norms,
                          sgp[:, 0].copy(), sgp[:, 1].copy(), dgp[:, 0].copy(), dgp[:, 1].copy(),
                          nn, fwd, bwd, colp, rowi, fvals, rowp, coli, bvals)


def read_graph_bin(path: str) -> PartitionGraph:
    with open(path, "rb") as f:
        return parse_graph_bin(f.read())


def partition_rows(g: PartitionGraph, global_rows: np.ndarray):
    """Split a global per-vertex array into (local rows, src-ghost rows) the way
    ``readFeaturesFile`` does (engine/utils.cpp:518-533)."""
    return global_rows[g.local_to_global], global_rows[g.src_ghost_gvid]


def dataset_paths(dataset_dir: str) -> dict:
    d = dataset_dir if dataset_dir.endswith("/") else dataset_dir + "/"
    return {
        "dir": d,
        "edges": d + "graph.bsnap.edges",
        "parts": d + "graph.bsnap.parts",
        "features": os.path.join(os.path.dirname(d.rstrip("/")), "features.bsnap"),
        "labels": os.path.join(os.path.dirname(d.rstrip("/")), "labels.bsnap"),
    }
