"""Decoder output -> scorer, on the device, without igraph.

The reference decoder (``PaceVaeV3.decode``, ``src/encoders/pace.py:1666-1749``) grows one igraph
object per candidate in nested Python loops: at step ``idx`` (= PACE vertex ``idx``; vertices 0 and
1 are the start sign and the input node) it

* samples the type of the new vertex from ``softmax(add_node(h))``      (``pace.py:1710-1713``),
* draws, for every earlier PACE vertex ``vi + 1`` (``vi = idx-2 .. 0``), the edge
  ``vi + 1 -> idx`` with probability ``sigmoid(add_edge(...))[:, vi]``   (``pace.py:1716-1741``),

and finally ``from_pace_graph_to_labeled_graph`` (``pace.py:1290-1305``) keeps the n real vertices
(PACE vertices ``2 .. n+1`` -> labeled vertices ``0 .. n-1``), subtracts the 3 reserved types from
the labels and drops the edges that leave the input node.  The labeled graph then goes, one at a
time, through ``BNLearnWrapper.score`` (``src/problem/bn/bnlearn.py:27-61``).

Here the same two draws stay CUDA tensors.  ``DecodeState.step`` is what a decoder loop calls in
place of the igraph bookkeeping; ``wire()`` turns the accumulated draws into the scorer's wire
arrays (``labels[B, n]``, ``ebits[B, n, EW]``: the reference's own ``l*/e*`` candidate format,
``src/toolkit/labeled.py:116-154``) with torch ops only, and ``BicScorer.score_wire`` consumes them
without a host round trip.  A candidate whose sampled types are not a permutation of the n
variables (a reserved type, a repeated variable, an early output node) is rejected by the scorer
with NaN — the reference fails on those as well (``IndexError`` in ``from_pace_graph_to_labeled_graph``
or the permutation assert of ``bnlearn.py:34-35``).

The VAE itself is out of scope: nothing here needs the model, only the
tensors its decoder samples from.
"""
from __future__ import annotations

from typing import Optional, Tuple

NUM_RESERVED_TYPES = 3    # input node (0), output node (1), start sign (2): labeled type = PACE type - 3


def edge_words(n: int) -> int:
    return (int(n) + 31) // 32


def decoded_to_wire(vertex_types, edge_draws, n: int, type_offset: int = NUM_RESERVED_TYPES):
    """Sampled decoder state -> wire arrays, on the tensors' device.

    vertex_types : int tensor ``[B, n]`` — PACE type sampled at steps ``idx = 2 .. n+1``
                   (``new_types`` of ``pace.py:1712``); BN variable = type - ``type_offset``.
    edge_draws   : bool / int tensor ``[B, n, n]`` — ``edge_draws[b, v, u]`` != 0 <=> the step that
                   created labeled vertex v drew the edge from labeled vertex u (PACE ``u+2 -> v+2``,
                   ``decisions`` of ``pace.py:1727-1741``); entries with u >= v are ignored, as are
                   the draws for the input node, which ``from_pace_graph_to_labeled_graph`` drops.

    Returns ``(labels, ebits)``: int32 ``[B, n]`` holding the uint16 labels (65535 = invalid) and
    int32 ``[B, n, EW]`` holding the bit patterns of the uint32 edge words.
    """
    import torch
    vt = vertex_types.reshape(-1, n).to(torch.int64)
    B = vt.shape[0]
    lab = vt - type_offset
    lab = torch.where((lab < 0) | (lab >= n), torch.full_like(lab, 65535), lab)
    EW = edge_words(n)
    d = (edge_draws.reshape(B, n, n) != 0)
    lower = torch.tril(torch.ones(n, n, dtype=torch.bool, device=d.device), diagonal=-1)   # u < v
    d = d & lower
    pad = EW * 32 - n
    if pad:
        d = torch.nn.functional.pad(d, (0, pad))
    weights = (torch.ones(32, dtype=torch.int64, device=d.device) << torch.arange(32, device=d.device))
    words = (d.reshape(B, n, EW, 32).to(torch.int64) * weights).sum(dim=3)        # 0 .. 2^32-1
    words = torch.where(words >= (1 << 31), words - (1 << 32), words).to(torch.int32)   # uint32 bit pattern
    return lab.to(torch.int32), words


class DecodeState:
    """Device-side stand-in for the igraph bookkeeping of ``PaceVaeV3.decode``.

    Usage inside a decoder loop (``pace.py:1692-1743``)::

        st = DecodeState(batch_size, n, device)
        for idx in range(2, n + 2):
            ...                                   # transformer step, unchanged
            st.step(idx, type_scores, edge_scores)   # replaces np.random.choice + g.add_vertex/add_edge
        scores = scorer.score_wire(*st.wire())    # CUDA in, CUDA out

    ``step`` samples exactly what the reference samples: one categorical draw per candidate from
    ``softmax(type_scores)`` and one Bernoulli draw per earlier vertex from ``edge_scores``.
    """

    def __init__(self, batch_size: int, n: int, device, generator=None):
        import torch
        self.B, self.n = int(batch_size), int(n)
        self.types = torch.zeros((self.B, self.n), dtype=torch.int64, device=device)
        self.draws = torch.zeros((self.B, self.n, self.n), dtype=torch.bool, device=device)
        self.generator = generator

    def step(self, idx: int, type_scores, edge_scores, force_type: Optional[int] = None):
        """idx: PACE vertex being created (2 .. n+1).  type_scores ``[B, T]`` (logits, ``add_node``
        output); edge_scores ``[B, idx-1]`` or ``[B, idx-1, 1]`` (probabilities, entry ``vi`` = edge
        from PACE vertex ``vi + 1``).  Returns the sampled types ``[B]``."""
        import torch
        v = idx - 2
        if not 0 <= v < self.n:
            raise ValueError(f"idx must be in 2..{self.n + 1}")
        if force_type is None:
            probs = torch.softmax(type_scores.reshape(self.B, -1).float(), dim=1)
            new_types = torch.multinomial(probs, 1, generator=self.generator)[:, 0]      # pace.py:1712
        else:
            new_types = torch.full((self.B,), int(force_type), dtype=torch.int64, device=self.types.device)
        self.types[:, v] = new_types
        es = edge_scores.reshape(self.B, -1)[:, :idx - 1]
        rnd = torch.rand(es.shape, device=es.device, generator=self.generator)
        decisions = rnd < es                                                           # pace.py:1727-1728
        # entry vi <-> PACE vertex vi + 1; labeled vertex u = vi - 1 (vi = 0 is the input node: dropped)
        if v > 0:
            self.draws[:, v, :v] = decisions[:, 1:v + 1]
        return new_types

    def wire(self) -> Tuple:
        return decoded_to_wire(self.types, self.draws, self.n)
