"""Which ranks share an NVLink domain (a host), and the process groups that follow from it.

The reference is run as one process per VM, each VM with one GPU, joined over TCP
(Readme.md:51-52, 122-127; distributedVggf.py:283-289) -- every rank is its own "node".  The fused
all-reduce of this framework (parallel.symm + csrc/allreduce.cu) works over peer-mapped memory, i.e.
inside one NVLink/NVSwitch domain.  This module decides, once at start-up, how a job is laid out:

* one node                      -> flat fused all-reduce over all ranks (the B200 x8 case);
* several nodes, >1 GPU each    -> hierarchical: fused all-reduce inside each node (scaled by
  1/world), then one NCCL all-reduce of the result between the ranks that hold the same local index;
* several nodes, 1 GPU each (the reference's deployment) or uneven nodes -> NCCL over all ranks.

``B200_NODE_ID`` overrides the host name (containers that share a host), ``B200_FAKE_NODE_SIZE=k``
pretends that ranks [0,k), [k,2k), ... live on different hosts (used by the tests).
"""
from __future__ import annotations

import os
import socket
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch.distributed as dist


@dataclass(frozen=True)
class NodeLayout:
    node_of: Tuple[int, ...]          # node index of every rank (nodes numbered by first appearance)
    rank: int

    @property
    def world(self) -> int:
        return len(self.node_of)

    @property
    def n_nodes(self) -> int:
        return max(self.node_of) + 1 if self.node_of else 1

    def members(self, node: int) -> List[int]:
        return [r for r, n in enumerate(self.node_of) if n == node]

    @property
    def node(self) -> int:
        return self.node_of[self.rank]

    @property
    def local_rank(self) -> int:
        return self.members(self.node).index(self.rank)

    @property
    def local_size(self) -> int:
        return len(self.members(self.node))

    @property
    def uniform(self) -> bool:
        return len({len(self.members(n)) for n in range(self.n_nodes)}) == 1

    def mode(self) -> str:
        """``flat`` | ``hierarchical`` | ``nccl`` -- see the module docstring."""
        if self.n_nodes == 1:
            return "flat"
        if self.uniform and self.local_size > 1:
            return "hierarchical"
        return "nccl"


def layout_from_ids(ids: Sequence[str], rank: int) -> NodeLayout:
    order: List[str] = []
    for i in ids:
        if i not in order:
            order.append(i)
    return NodeLayout(tuple(order.index(i) for i in ids), rank)


def node_identity(rank: int) -> str:
    fake = int(os.environ.get("B200_FAKE_NODE_SIZE", "0") or 0)
    if fake > 0:
        return f"fake-node-{rank // fake}"
    return os.environ.get("B200_NODE_ID") or socket.gethostname()


def detect_layout(group=None) -> NodeLayout:
    """Collective over ``group`` (default WORLD): every rank learns every rank's host."""
    if not (dist.is_available() and dist.is_initialized()):
        return NodeLayout((0,), 0)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ids: List[Optional[str]] = [None] * world
    dist.all_gather_object(ids, node_identity(rank), group=group)
    return layout_from_ids([str(i) for i in ids], rank)


def make_hierarchy_groups(layout: NodeLayout):
    """(node_group, cross_group) for this rank.  Every rank creates every group, in the same order
    (``new_group`` is collective over WORLD)."""
    if not layout.uniform:
        raise ValueError("hierarchical reduction needs the same number of ranks on every node")
    node_group = cross_group = None
    for n in range(layout.n_nodes):
        g = dist.new_group(layout.members(n))
        if n == layout.node:
            node_group = g
    for l in range(layout.local_size):
        g = dist.new_group([layout.members(n)[l] for n in range(layout.n_nodes)])
        if l == layout.local_rank:
            cross_group = g
    return node_group, cross_group
