"""FedAvg weighted parameter aggregation on the GPU (BASELINE.json config 5).

The reference has NO federated implementation: ``MML_Suite/train_congruent_federated.py`` and friends are empty files,
only ``federated/federated_utils.py:7-41`` (base64 of ``torch.save``) exists.  This module provides what that script
would need on the aggregation path: theta <- sum_k (n_k / sum_j n_j) * theta_k over the clients' flat parameter buffers
(float parameters and BatchNorm running statistics; ``num_batches_tracked`` is taken from client 0), as ONE HBM-bound
kernel over K flat buffers (mml_fedavg), or -- when every client lives on its own GPU -- as an in-place pre-scale
followed by a sum all-reduce over NCCL.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


def weights_from_counts(num_samples: Sequence[float], device) -> torch.Tensor:
    n = torch.tensor([float(v) for v in num_samples], dtype=torch.float64)
    if (n < 0).any() or float(n.sum()) <= 0:
        raise ValueError("client sample counts must be non-negative with a positive sum")
    return (n / n.sum()).to(torch.float32).to(device)


def aggregate_flat(buffers: List[torch.Tensor], num_samples: Sequence[float], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = sum_k w_k * buffers[k][i]; buffers: K contiguous fp32 CUDA tensors of equal length."""
    K = len(buffers)
    if K == 0 or K != len(num_samples):
        raise ValueError("need one sample count per client")
    n = buffers[0].numel()
    dev = buffers[0].device
    for b in buffers:
        if b.numel() != n or b.dtype != torch.float32 or not b.is_cuda or not b.is_contiguous() or b.device != dev:
            raise ValueError("client buffers must be contiguous fp32 CUDA tensors of equal length on one device")
        if b.data_ptr() % 16 != 0:
            raise ValueError("client buffers must be 16-byte aligned (the kernel reads float4; pass the flat buffer, not an odd-offset slice)")
    w = weights_from_counts(num_samples, dev)
    ptrs = torch.tensor([b.data_ptr() for b in buffers], dtype=torch.int64, device=dev)
    if out is None:
        out = torch.empty(n, device=dev)
    ops.fedavg(ptrs, w, K, out)
    return out


def federated_round(models: Sequence[torch.nn.Module], num_samples: Sequence[float]) -> None:
    """Aggregate K client models that live on ONE GPU and push the average back into every client (in place)."""
    engines = []
    for m in models:
        p = next(m.parameters())
        engines.append(m._get_engine(p.device))
    for attr in ("P", "S"):
        bufs = [getattr(e.fs, attr) for e in engines]
        avg = aggregate_flat(bufs, num_samples)
        for b in bufs:
            b.copy_(avg)
    for e in engines[1:]:
        e.fs.NBT.copy_(engines[0].fs.NBT)
    for e in engines:
        e.fs.refresh_shadows()


def federated_allreduce(model: torch.nn.Module, my_num_samples: float, group=None) -> None:
    """One client per rank/GPU: pre-scale by n_k / sum n, then sum all-reduce (NCCL) of parameters and running stats."""
    import torch.distributed as dist

    eng = model._get_engine(next(model.parameters()).device)
    n = torch.tensor([float(my_num_samples)], device=eng.device, dtype=torch.float64)
    dist.all_reduce(n, group=group)
    w = torch.tensor([float(my_num_samples) / float(n.item())], device=eng.device, dtype=torch.float32)
    for buf in (eng.fs.P, eng.fs.S):
        ops.scale_inplace(buf, w, 0)
        dist.all_reduce(buf, group=group)
    dist.broadcast(eng.fs.NBT, src=0, group=group)
    eng.fs.refresh_shadows()


class FederatedSimulator:
    """K simulated clients on one GPU -- the driver ``MML_Suite/train_congruent_federated.py`` would be (that file is empty in
    the reference).  Every client is a full model with its own optimizer; a round = ``local_steps`` fused train steps per
    client on that client's batches, then the FedAvg kernel over the clients' flat parameter / running-statistics buffers
    (weights n_k / sum n), pushed back into every client.  Adam moments stay local (plain FedAvg)."""

    def __init__(self, model_factory, optimizer_factory, num_clients: int, device):
        self.device = torch.device(device)
        self.clients = [model_factory().to(self.device) for _ in range(num_clients)]
        for k, m in enumerate(self.clients):
            m._mml_client_id = k  # mixed into the dropout seed: clients with congruent weights still draw different masks
        self.optimizers = [optimizer_factory(m) for m in self.clients]
        for m in self.clients[1:]:
            m.load_state_dict(self.clients[0].state_dict())  # congruent start

    def round(self, client_batches, loss_functions, num_samples, local_steps: int = 1):
        """client_batches[k] = list of batch dicts for client k; returns the mean local loss per client."""
        losses = []
        for m, opt, batches in zip(self.clients, self.optimizers, client_batches):
            tot = 0.0
            for s in range(local_steps):
                tot += m.train_step(batches[s % len(batches)], opt, loss_functions, self.device, None)["loss"]
            losses.append(tot / local_steps)
        federated_round(self.clients, num_samples)
        return losses
