"""B200 distance-solver driver: same API and tree output as the reference's
``DistanceSolver`` (cassiopeia/solver/DistanceSolver.py:29-463), numeric work on the GPU.

``solve`` follows reference :127-231 step for step -- root handling and error cases of
``setup_dissimilarity_map`` (:351-395), the join loop (:185-205), ``root_tree``, dropping
the root row (:214-222), ``populate_tree`` / ``collapse_unifurcations`` /
``collapse_mutationless_edges`` (:224-231) -- except that the dissimilarity map is built
by the CUDA map kernel and the whole ``while N > 2`` loop runs in ``cas_nj_solve`` /
``cas_upgma_solve``.  The join list comes back in id space and is replayed through the
same graph-building calls the reference makes, in the same order, so child order and
rooting match.

Deviations (documented in DESIGN.md): the n x n map is stored on the tree as a DataFrame
only when n <= ``materialize_limit`` (the reference always stores it, CassiopeiaTree.py:1958);
asymmetric user-supplied maps and dissimilarity functions other than
``weighted_hamming_distance`` without a precomputed map raise ``DistanceSolverError``
(no CPU fallback).
"""
from __future__ import annotations

import abc
import gc
import itertools
import os
import threading
from typing import Callable, Dict, List, Optional, Tuple

import networkx as nx
import numpy as np
import pandas as pd

from .. import engine, host_prep
from ..data import utilities as data_utilities
from ..mixins import DistanceSolverError
from . import dissimilarity_functions, solver_utilities

#: "b200_exact" (default): the reference's join list from maintained row sums + sequential sums only where needed;
#: "b200_strict": the reference's arithmetic literally (every row sum re-accumulated every iteration; verification);
#: "b200_fast": float32 map, maintained sums -- the reference's tree only when no near-tie occurs.
B200_IMPLEMENTATIONS = ("b200_exact", "b200_strict", "b200_fast")


class DistanceSolver(abc.ABC):
    #: store the map on the tree (pandas) only up to this many samples
    materialize_limit = 8192
    _algo = None  # "nj" | "upgma"

    def __init__(self, dissimilarity_function: Optional[Callable] = dissimilarity_functions.weighted_hamming_distance,
                 add_root: bool = False, prior_transformation: str = "negative_log", threads: int = 1):
        if not hasattr(self, "_implementation"):
            self._implementation = "b200_exact"
        self.prior_transformation = prior_transformation
        self.dissimilarity_function = dissimilarity_function
        self.add_root = add_root
        self.threads = threads
        self._pending_device_map = None
        self.last_join_list = None  # (joins, last2, names) of the most recent solve

    # ----------------------------------------------------------- configuration
    @property
    def precision(self) -> str:
        return {"b200_fast": "fast", "b200_strict": "strict"}.get(self._implementation, "exact")

    @property
    def _map_dtype(self) -> str:
        return "f32" if self.precision == "fast" else "f64"

    #: "auto" | "exact" | "tensor".  auto:
    #: * unweighted maps -> the int8 tcgen05 kernel (bit-exact and ~4x faster than the CUDA-core kernel);
    #: * prior-weighted maps in the reference-exact modes -> the exact CUDA-core kernel (the reference accumulates
    #:   float64 weights in character order: reproducing those roundings bit for bit is not a GEMM);
    #: * prior-weighted maps in fast mode with at most 64 states per character -> the fixed-point int8 tcgen05
    #:   kernel (four 7-bit digit planes, |error| <= 2^-23 of the largest weight; 1.6x the CUDA-core kernel at
    #:   50k x 40 x 50), else the CUDA-core kernel (the one-hot operand grows with the state count).
    map_method = "auto"

    def _map_method_for(self, cassiopeia_tree, character_matrix) -> str:
        if self.map_method != "auto":
            return self.map_method
        if character_matrix.shape[1] > 512:
            return "exact"
        if dissimilarity_functions.gpu_kernel_name(self.dissimilarity_function) != "weighted_hamming_distance":
            return "exact"
        # the one-hot operands of the tensor path are m * (largest state label + 2) bytes per cell: sparse, large
        # labels (compact_states keeps labels <= 4094 as they are) would make it slower than the CUDA-core kernel
        # and its workspace enormous
        values = character_matrix.to_numpy()
        largest = int(values.max()) if values.size else 0
        if character_matrix.shape[1] * (max(largest, 1) + 2) > 16384:
            return "exact"
        if cassiopeia_tree.priors:
            return "tensor" if (self.precision == "fast" and largest <= 64) else "exact"
        return "tensor"

    # ------------------------------------------------------------ map handling
    def _compute_device_map(self, cassiopeia_tree, character_matrix: pd.DataFrame):
        """Map of `character_matrix` on the device, in the reference's row order; also stored
        on the tree as a float64 DataFrame when small enough."""
        n = character_matrix.shape[0]
        names, D = data_utilities.dissimilarity_map_frame_inputs(
            character_matrix, cassiopeia_tree.priors, cassiopeia_tree.missing_state_indicator,
            self.dissimilarity_function, self.prior_transformation,
            dtype=self._map_dtype, method=self._map_method_for(cassiopeia_tree, character_matrix), to_host=False)
        if n <= self.materialize_limit:
            if self._map_dtype == "f64":
                host = D[:, :n].cpu().numpy()
            else:
                # the stored frame is float64 like the reference's; recompute exactly
                _, host = data_utilities.dissimilarity_map_frame_inputs(
                    character_matrix, cassiopeia_tree.priors, cassiopeia_tree.missing_state_indicator,
                    self.dissimilarity_function, self.prior_transformation, dtype="f64", method="exact")
            cassiopeia_tree.set_dissimilarity_map(pd.DataFrame(host, index=names, columns=names))
        return names, D

    def _upload_map(self, dissimilarity_map: pd.DataFrame):
        import torch

        names = list(dissimilarity_map.index)
        arr = dissimilarity_map.to_numpy(dtype=np.float64)
        n = arr.shape[0]
        if arr.shape != (n, n):
            raise DistanceSolverError("dissimilarity map must be square")
        if not np.array_equal(arr, arr.T):
            raise DistanceSolverError("the B200 solvers need a symmetric dissimilarity map")
        ld = engine.padded_ld(n)
        host = np.zeros((n, ld), dtype=np.float64 if self._map_dtype == "f64" else np.float32)
        host[:, :n] = arr
        return names, torch.from_numpy(host).cuda()

    def setup_dissimilarity_map(self, cassiopeia_tree, layer: Optional[str] = None) -> None:
        """Reference :351-395: implicit root, error cases, map creation if absent."""
        self._pending_device_map = None
        if cassiopeia_tree.root_sample_name is None:
            if self.add_root:
                self.setup_root_finder(cassiopeia_tree)
            else:
                raise DistanceSolverError(
                    "Please specify an explicit root sample in the Cassiopeia Tree or specify the "
                    "solver to add an implicit root")
        if self._pending_device_map is None and cassiopeia_tree.get_dissimilarity_map() is None:
            if self.dissimilarity_function is None:
                raise DistanceSolverError(
                    "Please specify a dissimilarity function or populate the CassiopeiaTree object "
                    "with a dissimilarity map")
            character_matrix = (cassiopeia_tree.layers[layer] if layer is not None
                                else cassiopeia_tree.character_matrix)
            self._pending_device_map = self._compute_device_map(cassiopeia_tree, character_matrix)

    def get_dissimilarity_map(self, cassiopeia_tree, layer: Optional[str] = None) -> pd.DataFrame:
        """Reference :94-125 (host frame; `solve` itself keeps the map on the device)."""
        self.setup_dissimilarity_map(cassiopeia_tree, layer)
        stored = cassiopeia_tree.get_dissimilarity_map()
        if stored is None and self._pending_device_map is not None:
            names, D = self._pending_device_map
            n = len(names)
            stored = pd.DataFrame(D[:, :n].double().cpu().numpy(), index=names, columns=names)
        return stored

    def _device_map(self, cassiopeia_tree, layer: Optional[str]) -> Tuple[List[str], object]:
        self.setup_dissimilarity_map(cassiopeia_tree, layer)
        if self._pending_device_map is not None:
            pending, self._pending_device_map = self._pending_device_map, None
            return pending
        return self._upload_map(cassiopeia_tree.get_dissimilarity_map())

    # ------------------------------------------------------------------ solve
    def solve(self, cassiopeia_tree, layer: Optional[str] = None, collapse_mutationless_edges: bool = False,
              logfile: str = "stdout.log") -> None:
        names, D = self._device_map(cassiopeia_tree, layer)
        n = len(names)
        # the parts of the host tail that do not depend on the join list are computed by a worker thread while the
        # GPU loop runs (the C call that enqueues the loop and the wait for its result both release the GIL)
        pre: Dict[str, object] = {}
        worker = None
        if (self.fast_tail and self.overlap_tail and os.environ.get("CAS_TAIL_THREAD", "1") != "0"
                and self._fast_tail_applicable(cassiopeia_tree, names, layer)):
            worker = threading.Thread(target=self._tail_precompute, name="cas-b200-tail",
                                      args=(pre, cassiopeia_tree, names, layer, collapse_mutationless_edges))
            worker.start()
        try:
            joins, last2 = engine.agglomerate(D, n, self._algo, self.precision)
        finally:
            if worker is not None:
                worker.join()
        del D
        self.last_join_list = (joins, last2, names)
        self.finish_tree(cassiopeia_tree, names, joins, last2, layer, collapse_mutationless_edges,
                         precomputed=pre if pre.get("ok") else None)

    def _tail_precompute(self, out: Dict[str, object], cassiopeia_tree, names, layer, collapse_mutationless_edges) -> None:
        """Join-independent inputs of `_finish_tree_arrays` (no mutation of the tree object): internal node names, the
        character matrix without the root sample, the leaves' state rows in map order.  Any failure leaves `out`
        without the "ok" mark and the tail computes everything itself."""
        try:
            n_joins = max(len(names) - 2, 0)
            out["internal_names"] = list(itertools.islice(solver_utilities.node_name_generator(), n_joins))
            cm = cassiopeia_tree.character_matrix
            root = cassiopeia_tree.root_sample_name
            dropped = cm.drop(index=root) if (cm is not None and root in cm.index) else None
            out["dropped"] = dropped
            leaf_states = None
            if collapse_mutationless_edges:
                character_matrix = (cassiopeia_tree.layers[layer] if layer is not None
                                    else (dropped if dropped is not None else cm))
                values = character_matrix.to_numpy()
                pos = character_matrix.index.get_indexer(list(names))
                leaf_states = np.zeros((len(names), values.shape[1]), dtype=values.dtype)
                leaf_states[pos >= 0] = values[pos[pos >= 0]]
            out["leaf_states"] = leaf_states
            # the per-leaf state lists populate_tree will attach (our tree class takes them ready-made)
            if hasattr(cassiopeia_tree, "accept_prepared_leaf_rows"):
                frame = cassiopeia_tree.layers[layer] if layer is not None else (dropped if dropped is not None else cm)
                if frame is not None:
                    out["leaf_rows"] = (frame.index.tolist(), frame.to_numpy().tolist())
            out["ok"] = True
        except Exception:  # pragma: no cover - the inline path reports the real error
            out.clear()

    #: use the array-based tail (cassiopeia_b200/tree_tail.py) when the stock ``root_tree`` is in place
    fast_tail = True
    #: compute the join-independent inputs of that tail in a worker thread while the GPU loop runs
    overlap_tail = True

    def finish_tree(self, cassiopeia_tree, names: List[str], joins: np.ndarray, last2: np.ndarray,
                    layer: Optional[str] = None, collapse_mutationless_edges: bool = False,
                    precomputed: Optional[Dict[str, object]] = None) -> None:
        """Host tail of `solve`: turns an id-space join list (leaf at map row p has id p, the
        t-th join creates id n+t) into the rooted, populated tree of reference :181-231.

        Two equivalent routes: the array-based one (`_finish_tree_arrays`) and the
        generic one that replays the join list through the reference's graph-building calls and hooks
        (`_finish_tree_generic`; used when a subclass overrides ``root_tree``, for fewer than 3 samples and
        for non-integer state matrices)."""
        if self.fast_tail and self._fast_tail_applicable(cassiopeia_tree, names, layer):
            # the tail allocates ~10 container objects per node and frees none: the cyclic collector's
            # generation scans would be pure overhead (a third of the tail's time at 10^5 nodes)
            gc_was_enabled = gc.isenabled()
            gc.disable()
            try:
                self._finish_tree_arrays(cassiopeia_tree, names, joins, last2, layer, collapse_mutationless_edges,
                                         precomputed)
            finally:
                if gc_was_enabled:
                    gc.enable()
        else:
            self._finish_tree_generic(cassiopeia_tree, names, joins, last2, layer, collapse_mutationless_edges)

    _stock_root_tree = None  # set by the concrete solvers to their own root_tree

    def _fast_tail_applicable(self, cassiopeia_tree, names, layer) -> bool:
        if len(names) < 3 or type(self).root_tree is not type(self)._stock_root_tree:
            return False
        if self._algo != "upgma" and cassiopeia_tree.root_sample_name not in names:
            return False
        character_matrix = cassiopeia_tree.layers[layer] if layer is not None else cassiopeia_tree.character_matrix
        if character_matrix is None:
            return True
        return all(np.issubdtype(dt, np.integer) for dt in character_matrix.dtypes)

    def _finish_tree_arrays(self, cassiopeia_tree, names, joins, last2, layer, collapse_mutationless_edges,
                            precomputed: Optional[Dict[str, object]] = None) -> None:
        from .. import tree_tail

        pre = precomputed if (precomputed and len(precomputed.get("internal_names", ())) == len(joins)) else None
        if pre is not None:
            internal_names = pre["internal_names"]
        else:
            node_name_generator = solver_utilities.node_name_generator()
            internal_names = list(itertools.islice(node_name_generator, len(joins)))
        # remove root from character matrix before populating tree (:214-222)
        if pre is not None:
            if pre["dropped"] is not None:
                cassiopeia_tree.character_matrix = pre["dropped"]
        elif cassiopeia_tree.root_sample_name in cassiopeia_tree.character_matrix.index:
            cassiopeia_tree.character_matrix = cassiopeia_tree.character_matrix.drop(
                index=cassiopeia_tree.root_sample_name)
        leaf_states = None
        if pre is not None:
            leaf_states = pre["leaf_states"]
        elif collapse_mutationless_edges:
            character_matrix = (cassiopeia_tree.layers[layer] if layer is not None
                                else cassiopeia_tree.character_matrix)
            values = character_matrix.to_numpy()
            pos = character_matrix.index.get_indexer(list(names))
            leaf_states = np.zeros((len(names), values.shape[1]), dtype=values.dtype)
            leaf_states[pos >= 0] = values[pos[pos >= 0]]
        tree, internal_states = tree_tail.build_final_tree(
            names, internal_names, joins, last2, self._algo, cassiopeia_tree.root_sample_name, leaf_states,
            cassiopeia_tree.missing_state_indicator, collapse_mutationless_edges)
        if pre is not None and pre.get("leaf_rows") is not None:
            cassiopeia_tree.accept_prepared_leaf_rows(*pre["leaf_rows"])
        try:
            cassiopeia_tree.populate_tree(tree, layer=layer)
        finally:
            getattr(cassiopeia_tree, "__dict__", {}).pop("_prepared_leaf_rows", None)  # never outlives this call
        if internal_states:
            bulk = getattr(cassiopeia_tree, "set_internal_character_states", None)
            if bulk is not None:
                bulk(internal_states)
            else:  # the reference's tree object: one public call per internal node
                for node, states in internal_states.items():
                    cassiopeia_tree.set_character_states(node, states)

    def _finish_tree_generic(self, cassiopeia_tree, names, joins, last2, layer, collapse_mutationless_edges) -> None:
        """Replays the join list through the reference's graph-building sequence (:181-199), roots the tree
        with the (possibly overridden) ``root_tree`` hook and runs the tree object's own collapse passes."""
        node_name_generator = solver_utilities.node_name_generator()
        n = len(names)
        tree = nx.Graph()
        tree.add_nodes_from(names)
        label = list(names)
        for a, b in np.asarray(joins).tolist():
            new_node_name = next(node_name_generator)
            label.append(new_node_name)
            tree.add_node(new_node_name)
            tree.add_edges_from([(new_node_name, label[a]), (new_node_name, label[b])])
        if n >= 2:
            remaining = [label[int(last2[0])], label[int(last2[1])]]
        else:
            remaining = list(names)

        tree = self.root_tree(tree, cassiopeia_tree.root_sample_name, remaining)

        # remove root from character matrix before populating tree (:214-222)
        if cassiopeia_tree.root_sample_name in cassiopeia_tree.character_matrix.index:
            cassiopeia_tree.character_matrix = cassiopeia_tree.character_matrix.drop(
                index=cassiopeia_tree.root_sample_name)

        cassiopeia_tree.populate_tree(tree, layer=layer)
        cassiopeia_tree.collapse_unifurcations()
        if collapse_mutationless_edges:
            cassiopeia_tree.collapse_mutationless_edges(infer_ancestral_characters=True)

    # ------------------------------------------------------- reference hooks
    @abc.abstractmethod
    def root_tree(self, tree: nx.Graph, root_sample: str, remaining_samples: List[str]) -> nx.DiGraph:
        ...

    @abc.abstractmethod
    def setup_root_finder(self, cassiopeia_tree) -> None:
        ...

    def find_cherry(self, dissimilarity_map: np.ndarray) -> Tuple[int, int]:
        """Positions (i < j) of the pair the rule joins first (one scan on the GPU)."""
        import torch

        arr = np.asarray(dissimilarity_map, dtype=np.float64)
        n = arr.shape[0]
        ld = engine.padded_ld(n)
        host = np.zeros((n, ld), dtype=np.float64)
        host[:, :n] = arr
        joins, _ = engine.agglomerate(torch.from_numpy(host).cuda(), n, self._algo, "strict", max_iters=1)
        return int(joins[0, 0]), int(joins[0, 1])

    def update_dissimilarity_map(self, dissimilarity_map: pd.DataFrame, cherry: Tuple[str, str],
                                 new_node: str) -> pd.DataFrame:
        """Map after joining `cherry` into `new_node`: the two rows/columns are dropped and the
        new node appended last (reference NeighborJoiningSolver.py:173-216 / UPGMASolver.py:140-193)."""
        index = list(dissimilarity_map.index)
        i, j = index.index(cherry[0]), index.index(cherry[1])
        size_i, size_j = self._cluster_sizes(cherry, new_node)
        arr = engine.update_map(dissimilarity_map.to_numpy(dtype=np.float64), i, j, self._algo, size_i, size_j)
        names = [nm for p, nm in enumerate(index) if p != i and p != j] + [new_node]
        return pd.DataFrame(arr, index=names, columns=names)

    def _cluster_sizes(self, cherry, new_node) -> Tuple[int, int]:
        return 1, 1
