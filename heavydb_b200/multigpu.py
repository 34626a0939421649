"""Multi-GPU host logic of the path: one process per GPU (torch.distributed), fragments sharded with the reference's
rule, partial aggregate tables merged with ONE all-reduce per dense accumulator array.

Reference behaviour replaced:
  * fragment -> device:  ``fragment_id % num_devices`` (Fragmenter/InsertOrderFragmenter.cpp:435-443, remapped per
    query in QueryEngine/Execute.cpp:1130-1175);
  * per-device result merge on the HOST: ``Executor::reduceMultiDeviceResults`` -> ``ResultSetStorage::reduce``
    (Execute.cpp:1696,1772-1792; ResultSetReduction.cpp:203-396).  Here every GPU leaves its partial table as dense,
    position-aligned arrays initialised to the identity of their reduction (COUNT/SUM -> 0 with ncclSum, MIN ->
    +inf with ncclMin, MAX -> -inf with ncclMax, "group touched" flags -> 0 with ncclMax), so the merge is exactly
    ``len(arrays)`` all-reduces over NVLink/NVSwitch and every rank ends up with the global table.

torch is plumbing here (process group + tensor views); the arrays themselves are produced by libb2q.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

from . import abi


def shard_fragments(fragment_ids: Iterable[int], rank: int, world: int) -> List[int]:
    """The reference's placement rule: fragment f lives on device f % num_devices."""
    return [f for f in fragment_ids if f % world == rank]


class CudaArray:
    """Zero-copy view of a raw device pointer for torch (``__cuda_array_interface__`` v2)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TYPESTR = {abi.DT_FLOAT64: "<f8", abi.DT_INT64: "<i8", abi.DT_UINT8: "|u1"}


def reduce_op(dist, redop: int):
    return {abi.RED_SUM: dist.ReduceOp.SUM, abi.RED_MIN: dist.ReduceOp.MIN, abi.RED_MAX: dist.ReduceOp.MAX}[redop]


def allreduce_tensors(tensors_and_ops: Sequence[tuple], dist) -> None:
    """[(tensor, redop)] -> in-place all-reduce of each (works for CUDA/NCCL and CPU/gloo tensors alike).  Bitwise OR
    (the estimator bitmap, reduce_estimator_results CardinalityEstimator.cpp:142-161) is not an NCCL reduction: the
    bitmaps are all-gathered and OR-ed locally."""
    for t, op in tensors_and_ops:
        if op == abi.RED_BOR:
            parts = [t.new_empty(t.shape) for _ in range(dist.get_world_size())]
            dist.all_gather(parts, t)
            t.zero_()
            for p in parts:
                t.bitwise_or_(p)
            continue
        dist.all_reduce(t, op=reduce_op(dist, op))


def allreduce_partial(partial, torch, dist) -> int:
    """All-reduce every dense array of a ``heavydb_b200.executor.Partial`` in place.  Returns the bytes reduced."""
    if not partial.is_mergeable():
        raise RuntimeError("baseline-hash partial tables are not position-aligned across devices; "
                           "they cannot be merged by an all-reduce (SURVEY.md §7 hard part 3)")
    items, nbytes = [], 0
    for ptr, n, dt, op in partial.arrays():
        t = torch.as_tensor(CudaArray(ptr, n, _TYPESTR[dt]), device="cuda")
        items.append((t, op))
        nbytes += t.numel() * t.element_size()
    allreduce_tensors(items, dist)
    return nbytes
