"""One-process-per-GPU data parallelism for the fused step (the reference has no distributed code at all).

Semantics: every rank holds the full fp32 weights + Adam state, takes its own shard of the global
batch, keeps PER-REPLICA BatchNorm statistics (the reference is single-device; there is no SyncBN to match), and the
gradients are summed over ranks and divided by the world size inside the Adam kernel (``hyper[5] = 1/world``).

The flat gradient buffer ``G`` is reduced in contiguous ranges (two buckets, each split once more by the step plan), ordered so that communication hides under compute:
the image encoder has 2/3 of the parameters but few FLOPs, so its backward runs FIRST and bucket 0 (image encoder +
head, ~85 MB) is all-reduced on a side stream while the audio encoder's backward (87 % of the FLOPs) is still running;
the audio encoder follows in two ranges: layer3..fc (94 % of its parameters, ~42 MB) as soon as layer3's backward is
done -- under the backward of layer2 / layer1 / the stem -- and the small remainder (~3 MB) at the end of the step.  The wgrad kernels write straight into ``G``, so there is no pack/copy step.
The data plane is the library's own NCCL communicator (``mml_comm_init`` / ``mml_allreduce_bucket``, csrc/comm.cu): NCCL over
NVLink / NVSwitch with a capped CTA budget (``MML_NCCL_MAX_CTAS``, default 32: the all-reduces run under the audio encoder's
backward and every SM NCCL takes is one its persistent kernels lose -- measured at N = 2 on the final round-2 build: 8 CTAs 2.89 ms/step
end to end, 16 2.65, 32 2.59).  ``torch.distributed`` remains the CONTROL plane (rendezvous,
hand-over of the communicator id, the initial broadcast, barriers); both the collectives and the cross-stream dependencies are
captured into the step's CUDA graph.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def bucket_ranges(offsets: dict, total: int) -> List[Tuple[int, int]]:
    """[(begin, end)] element ranges of G: bucket 0 = image encoder + head (ready first), bucket 1 = audio encoder."""
    img = [o for n, o in offsets.items() if n.startswith("image_encoder.")]
    if not img:
        return [(0, total)]
    split = min(img)
    return [(split, total), (0, split)]


class DataParallel:
    def __init__(self, process_group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (use mml_b200.dist.init_from_env())")
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self.buckets: List[Tuple[int, int]] = []
        self.engine = None
        self.max_ctas = int(os.environ.get("MML_NCCL_MAX_CTAS", "32"))

    def _ensure_comm(self, device: torch.device) -> None:
        """Create the library-owned communicator of this device once: rank 0 draws the id, torch.distributed hands it over."""
        from . import ops

        idx = device.index if device.index is not None else torch.cuda.current_device()
        if ops.comm_world(idx) == self.world_size:
            return
        payload = [ops.comm_unique_id(idx) if self.rank == 0 else None]
        dist.broadcast_object_list(payload, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        ops.comm_init(idx, payload[0], self.rank, self.world_size, self.max_ctas)

    def attach(self, engine) -> None:
        self.engine = engine
        self.buckets = bucket_ranges(engine.fs.offsets, engine.fs.total)
        self.comm_stream = torch.cuda.Stream(device=engine.device)
        if self.world_size > 1:
            self._ensure_comm(engine.device)
        engine.world = self.world_size
        engine.allreduce = self._allreduce if self.world_size > 1 else None
        if hasattr(engine, "allreduce_range"):
            engine.allreduce_range = self._allreduce_range if self.world_size > 1 else None

    def _allreduce(self, plan, idx: int, update=None) -> None:
        """All-reduce bucket ``idx`` of G on the communication stream, then run ``update`` (the Adam launch for that range)
        on the same stream; the last bucket joins the communication stream back into the producer stream."""
        eng = self.engine
        if idx >= len(self.buckets):
            return
        a, b = self.buckets[idx]
        producer = torch.cuda.current_stream(eng.device)  # the stream whose backward just finished this bucket
        self.comm_stream.wait_stream(producer)
        with torch.cuda.stream(self.comm_stream):
            self._sum(eng.fs.G[a:b])
            if update is not None:
                update()
        if idx == len(self.buckets) - 1:
            producer.wait_stream(self.comm_stream)  # everything of this step is ordered before what follows on the main stream

    def _allreduce_range(self, a: int, b: int, producers, update=None, join: bool = False) -> None:
        """All-reduce G[a:b) on the communication stream once every stream in ``producers`` has finished what it has queued,
        then run ``update`` (that range's Adam launch) there; ``join``: the current stream waits for the communication stream
        (last range of a step).  Ranges are issued in the same order on every rank (same captured schedule)."""
        eng = self.engine
        for st in producers:
            self.comm_stream.wait_stream(st)
        with torch.cuda.stream(self.comm_stream):
            self._sum(eng.fs.G[a:b])
            if update is not None:
                update()
        if join:
            torch.cuda.current_stream(eng.device).wait_stream(self.comm_stream)

    @staticmethod
    def _sum(buf: torch.Tensor) -> None:
        from . import ops

        ops.allreduce_bucket(buf)

    def broadcast_state(self, engine) -> None:
        """Make every rank start from rank 0's weights / Adam state / running statistics."""
        fs = engine.fs
        for t in (fs.P, fs.M, fs.V, fs.S, fs.NBT, fs.step):
            dist.broadcast(t, src=0, group=self.group)
        fs.refresh_shadows()


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the default process group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [begin, end) of the global batch owned by ``rank`` (contiguous, equal shards)."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by the world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per
