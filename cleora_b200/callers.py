"""Callers of the embedding loop that hand it a starting matrix or tap it at several depths.

Reference: ``pycleora/__init__.py`` -- ``embed_with_node_features`` (:167-203), ``embed_multiscale`` (:279-309),
``update_graph`` (:515-523), ``embed_inductive`` (:540-580).  Same names, arguments, defaults and errors; every
iteration runs in the device-resident loop behind ``cleora_embed`` (no per-iteration host round trip).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .pycleora import SparseMatrix

_DEFAULT_DIM = 256
_DEFAULT_ITERS = 40


def _rows_of(graph: SparseMatrix, entity_ids: Sequence[str]) -> np.ndarray:
    """Row index per id, -1 where the graph does not hold it (the reference skips unknown ids silently)."""
    lookup = {eid: i for i, eid in enumerate(graph.entity_ids)}
    return np.fromiter((lookup.get(e, -1) for e in entity_ids), dtype=np.int64, count=len(entity_ids))


def embed_with_node_features(
    graph: SparseMatrix,
    node_features: Dict[str, np.ndarray],
    num_iterations: int = _DEFAULT_ITERS,
    propagation: str = "left",
    normalization: str = "l2",
    feature_weight: float = 0.5,
    num_workers: Optional[int] = None,
) -> np.ndarray:
    """Blend caller-supplied features into the deterministic start, then run the loop (:167-203).

    ``X0[i] = (1-w)*init[i] + w*feat`` for every id present in the graph, in f32 with the python-float weights
    (numpy's weak-scalar promotion), exactly the arithmetic of the reference's row assignment.
    """
    from . import embed
    if not node_features:
        raise ValueError("node_features must be a non-empty dict of entity_id -> feature_vector")
    ids = list(node_features.keys())
    width = len(node_features[ids[0]])
    feats = []
    for eid in ids:
        f = np.array(node_features[eid], dtype=np.float32)
        feats.append(f)
    start = graph.initialize_deterministically(width)
    rows = _rows_of(graph, ids)
    for eid, r, f in zip(ids, rows, feats):
        if r < 0:
            continue
        if len(f) != width:
            raise ValueError(f"Feature for '{eid}' has dimension {len(f)}, expected {width}")
        start[r] = (1 - feature_weight) * start[r] + feature_weight * f
    return embed(graph, feature_dim=width, num_iterations=num_iterations, propagation=propagation,
                 normalization=normalization, initial_embeddings=start, num_workers=num_workers)


def embed_multiscale(
    graph: SparseMatrix,
    feature_dim: int = _DEFAULT_DIM,
    scales: List[int] = None,
    propagation: str = "left",
    normalization: str = "l2",
    seed: int = 0,
    num_workers: Optional[int] = None,
    whiten: bool = True,
) -> np.ndarray:
    """Iterates tapped at each depth in ``scales`` and concatenated column-wise (:279-309).

    The loop has no state besides X, so the taps are the ends of consecutive device-resident segments:
    segment k runs ``scales[k] - scales[k-1]`` iterations starting from the previous tap.
    """
    from . import _DEVICE_NORMS, _validate_propagation
    _validate_propagation(propagation)
    if scales is None:
        scales = [10, 20, 30, 40]
    if not scales or not all(isinstance(s, int) and s > 0 for s in scales):
        raise ValueError("scales must be a non-empty list of positive integers")
    if normalization not in _DEVICE_NORMS:
        raise ValueError(f"Unknown normalization method: {normalization}. Use 'l2', 'l1', 'spectral', or 'none'.")
    depths = sorted(scales)
    n = graph.num_entities
    out = np.empty((n, feature_dim * len(depths)), np.float32)
    x: Optional[np.ndarray] = None
    done = 0
    for k, depth in enumerate(depths):
        steps = depth - done
        if steps > 0 or x is None:
            x, _ = graph.embed_device(feature_dim, steps, propagation, _DEVICE_NORMS[normalization], seed, x, 0.0,
                                      0.0, whiten)
        done = depth
        out[:, k * feature_dim:(k + 1) * feature_dim] = x
    return out


def update_graph(existing_edges: List[str], new_edges: List[str], columns: str, hyperedge_trim_n: int = 16,
                 num_workers: Optional[int] = None) -> SparseMatrix:
    """Rebuild over the union of both edge lists, old lines first (:515-523)."""
    lines = [*existing_edges, *new_edges]
    return SparseMatrix.from_iterator(iter(lines), columns, hyperedge_trim_n, num_workers)


def embed_inductive(
    trained_graph: SparseMatrix,
    trained_embeddings: np.ndarray,
    existing_edges: List[str],
    new_edges: List[str],
    columns: str,
    num_iterations: int = _DEFAULT_ITERS,
    propagation: str = "left",
    normalization: str = "l2",
    hyperedge_trim_n: int = 16,
    num_workers: Optional[int] = None,
) -> Tuple[SparseMatrix, np.ndarray]:
    """Warm-start the loop on a grown graph from embeddings trained on the old one (:540-580).

    New entities start from ``0.01 * N(0,1)`` drawn from numpy's global generator (one ``randn`` call of the full
    shape, like the reference, so a caller-seeded run is reproducible across the two implementations).
    """
    from . import embed
    if trained_embeddings.shape[0] != trained_graph.num_entities:
        raise ValueError(f"trained_embeddings has {trained_embeddings.shape[0]} rows but graph has "
                         f"{trained_graph.num_entities} entities")
    grown = update_graph(existing_edges, new_edges, columns, hyperedge_trim_n, num_workers)
    width = trained_embeddings.shape[1]
    start = np.random.randn(grown.num_entities, width).astype(np.float32) * 0.01
    old_rows = _rows_of(trained_graph, grown.entity_ids)
    known = old_rows >= 0
    start[known] = trained_embeddings[old_rows[known]]
    result = embed(grown, feature_dim=width, num_iterations=num_iterations, propagation=propagation,
                   normalization=normalization, initial_embeddings=start, num_workers=num_workers)
    return grown, result
