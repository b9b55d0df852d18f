"""Single-partition graph arrays from an edge list in plain numpy (TEST INFRASTRUCTURE ONLY).

What DataLoader::preprocess + Graph::init leave in memory for numNodes == 1
(graph/dataloader.cpp:153-185,225-330, graph/graph.cpp:7-115), restated with sorts instead of the
reference's per-vertex vectors so that `bench.py --impl reference` can build its workload without
loading the product library: no ghosts, local id == global id, in-edges of a vertex in edge-file
order (stable sort by destination, quirk Q4), out-edges likewise by source, edge value
(deg_in(src)+1)^-1/2 * (deg_in(dst)+1)^-1/2 evaluated like the reference (pow in double, narrowed to
float, the two float factors multiplied in float), vtxDataVec = norm * norm.  Self loops are dropped
(dataloader.cpp:268-269).  tests/test_oracle.py holds it byte-identical to the compiled reference
loader and to the product's preprocessor.
"""
from __future__ import annotations

import numpy as np

from dorylus_b200.formats import PartitionGraph


def single_partition_graph(src: np.ndarray, dst: np.ndarray, num_vertices: int) -> PartitionGraph:
    V = int(num_vertices)
    keep = src != dst
    if not keep.all():
        src, dst = src[keep], dst[keep]
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    E = int(src.size)
    indeg = np.bincount(dst, minlength=V)
    outdeg = np.bincount(src, minlength=V)
    norm = np.power((indeg + 1).astype(np.float64), -0.5).astype(np.float32)  # float vtxNorm = std::pow(deg, -.5)
    col_ptrs = np.zeros(V + 1, np.uint64)
    col_ptrs[1:] = np.cumsum(indeg)
    row_ptrs = np.zeros(V + 1, np.uint64)
    row_ptrs[1:] = np.cumsum(outdeg)
    by_dst = np.argsort(dst, kind="stable")
    row_idxs = src[by_dst]
    fwd_vals = norm[row_idxs] * norm[dst[by_dst]]  # float * float, dataloader.cpp:164
    del by_dst
    by_src = np.argsort(src, kind="stable")
    col_idxs = dst[by_src]
    bwd_vals = norm[src[by_src]] * norm[col_idxs]  # dataloader.cpp:177
    del by_src
    empty = np.zeros(0, np.uint32)
    return PartitionGraph(
        local_vtx_cnt=V, global_vtx_cnt=V, src_ghost_cnt=0, dst_ghost_cnt=0, local_in_edge_cnt=E,
        local_out_edge_cnt=E, global_edge_cnt=E, local_to_global=np.arange(V, dtype=np.uint32),
        norms=(norm * norm).astype(np.float32), src_ghost_gvid=empty, src_ghost_lvid=empty, dst_ghost_gvid=empty,
        dst_ghost_lvid=empty, num_nodes=1, fwd_send=[empty], bwd_send=[empty], col_ptrs=col_ptrs,
        row_idxs=row_idxs, fwd_vals=fwd_vals.astype(np.float32), row_ptrs=row_ptrs, col_idxs=col_idxs,
        bwd_vals=bwd_vals.astype(np.float32))


def graph_bin_image(g: PartitionGraph) -> bytes:
    """graph.<id>.bin of a SINGLE-partition graph (RawGraph::dump, graph/graph.cpp:200-273): counts, localToGlobal,
    vtxDataVec, (no ghost pairs), numNodes, empty send lists, CSC {cols, nnz, values, columnPtrs, rowIdxs}, CSR
    likewise.  What Graph::init reads back -- the reference-engine library (oracle/ref_engine.cpp) loads it."""
    assert g.num_nodes == 1 and g.src_ghost_cnt == 0 and g.dst_ghost_cnt == 0
    V = g.local_vtx_cnt
    parts = [np.array([V, g.global_vtx_cnt, 0, 0], "<u4").tobytes(),
             np.array([g.local_in_edge_cnt, g.local_out_edge_cnt, g.global_edge_cnt], "<u8").tobytes(),
             g.local_to_global.astype("<u4").tobytes(), g.norms.astype("<f4").tobytes(),
             np.array([1], "<u4").tobytes(),            # numNodes
             np.array([0], "<u4").tobytes(),            # forwardLocalVtxDsts[0]: empty
             np.array([0], "<u4").tobytes()]            # backwardLocalVtxDsts[0]: empty
    for ptrs, idxs, vals in ((g.col_ptrs, g.row_idxs, g.fwd_vals), (g.row_ptrs, g.col_idxs, g.bwd_vals)):
        parts += [np.array([V], "<u4").tobytes(), np.array([idxs.size], "<u8").tobytes(), vals.astype("<f4").tobytes(),
                  ptrs.astype("<u8").tobytes(), idxs.astype("<u4").tobytes()]
    return b"".join(parts)
