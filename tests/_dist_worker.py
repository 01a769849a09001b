"""One rank of the multi-GPU numerics check (launched by tests/test_dist_gpu.py through torch.distributed.run, one process per
GPU, NCCL).  Prints ``DIST_OK <name>`` per passed check on rank 0 and exits non-zero on the first failure.

Checks (oracle/late_fusion_oracle.py::data_parallel_grads / fedavg):
  1. the gradient buffer after the step's bucketed all-reduce, divided by the world size, equals the MEAN over replicas of
     the oracle's per-replica gradient (per-replica BatchNorm, no SyncBN).  The oracle's backward is teacher-forced on the
     activations each rank stored (tests/test_step_gpu.py docstring), which makes the comparison well conditioned;
  2. it also equals the sum of the two shards' gradients computed WITHOUT data parallelism on one GPU;
  3. parameters, Adam moments and the bf16 shadow are bit-identical across ranks after 3 (eager + graph-replayed) steps;
  4. ``federated_allreduce`` (one client per rank, pre-scale + NCCL sum) equals the oracle's FedAvg of the clients' states.
"""
import copy
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import late_fusion_oracle as O  # noqa: E402
from mml_b200 import dist as mdist  # noqa: E402
from mml_b200 import fedavg  # noqa: E402
from mml_b200.avmnist import AVMNIST  # noqa: E402
from mml_b200.resnet import ResNet18, ResNet34  # noqa: E402


class Term:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.CrossEntropyLoss(), 1.0


LOSS = {"cross_entropy": Term()}


def make_batch(d, lo, hi):
    return {"audio_original": d["audio"][lo:hi], "audio_missing_index": d["audio_mask"][lo:hi], "image_original": d["image"][lo:hi],
            "image_missing_index": d["image_mask"][lo:hi], "labels": d["labels"][lo:hi], "pattern_name": ["ai"] * (hi - lo)}


def nchw(t):
    return t.detach().float().cpu().permute(0, 3, 1, 2).contiguous()


def forced_from_plan(plan, state):
    forced = {}
    for pre, ep in (("audio_encoder.", plan.audio), ("image_encoder.", plan.image)):
        for name, t in ep.taps.items():
            forced[pre + name] = nchw(t)
        forced[pre + "avgpool"] = ep.pooled.detach().cpu().clone()
        act = torch.nn.functional.batch_norm(forced[pre + "conv1"], None, None, state[pre + "bn1.weight"], state[pre + "bn1.bias"], True, 0.1, 1e-5)
        forced[pre + "relu1"] = torch.relu(act)  # fp32: the fused stem tail pools the un-rounded BatchNorm outputs and rounds only the winner
    return forced


def build(dev, dp, seed=0, dropout=0.5, graphs=True):
    torch.manual_seed(seed)
    model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=dropout).to(dev)
    if dp is not None:
        model.enable_data_parallel(dp)
    eng = model._get_engine(dev)
    eng.use_graphs = graphs
    if dp is not None:
        dp.broadcast_state(eng)
    return model, eng


def ok(rank, name, extra=""):
    if rank == 0:
        print(f"DIST_OK {name} {extra}", flush=True)


def main():
    rank, local_rank, world = mdist.init_from_env("nccl")
    assert world >= 2, "launch with torchrun --nproc-per-node >= 2"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    per = 8
    B = per * world
    hw = (32, 94)
    d = O.synthetic_batch(B, 4242, hw)
    lo, hi = mdist.shard_batch(B, rank, world)

    # ---- 1 + 2: reduced gradients ------------------------------------------------------------------------------------
    dp = mdist.DataParallel()
    model, eng = build(dev, dp, graphs=False)
    torch.manual_seed(0)
    state = O.init_avmnist_state()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    model.train_step(make_batch(d, lo, hi), opt, LOSS, dev, None, dropout_mask=d["dropout_mask"][lo:hi])
    plan = next(iter(eng.plans.values()))
    G = eng.fs.G.detach().clone()  # SUM over ranks (the Adam kernel divides by the world size)
    A = O.apply_missing_mask(d["audio"], d["audio_mask"])[lo:hi]
    I = O.apply_missing_mask(d["image"], d["image_mask"])[lo:hi]
    ref = O.train_step(copy.deepcopy(state), {}, A, I, d["labels"][lo:hi], d["dropout_mask"][lo:hi], 0.5, apply_update=False,
                       forced=forced_from_plan(plan, state))
    names = [n for n, _ in model.named_parameters()]
    mine = torch.cat([ref["grads"][n].reshape(-1) for n in names]).to(dev)
    dist.all_reduce(mine)  # mean over replicas of the oracle's per-replica gradients == O.data_parallel_grads semantics
    mine /= world
    got = torch.cat([eng.fs._view(G, n, p).reshape(-1) for n, p in model.named_parameters()]) / world
    rel = float((got - mine).norm() / mine.norm())
    assert rel < 3e-2, f"DP-reduced gradient vs oracle data_parallel_grads (teacher forced): rel L2 {rel}"
    ok(rank, "reduced_gradients_equal_oracle_data_parallel", f"rel_l2={rel:.4f}")
    # every rank must hold the same reduced buffer
    g0 = G.clone()
    dist.broadcast(g0, src=0)
    assert torch.equal(g0, G), "reduced gradient buffers differ between ranks"
    ok(rank, "reduced_gradients_identical_on_all_ranks")
    # ... and it is the plain sum of the shards' gradients computed without data parallelism
    total = torch.zeros_like(G)
    for r in range(world):
        m1, e1 = build(dev, None, graphs=False)
        o1 = torch.optim.Adam(m1.parameters(), lr=5e-4, weight_decay=1e-4)
        a, b = mdist.shard_batch(B, r, world)
        m1.train_step(make_batch(d, a, b), o1, LOSS, dev, None, dropout_mask=d["dropout_mask"][a:b])
        total += e1.fs.G
        del m1, e1, o1
    rel2 = float((total - G).norm() / G.norm())
    assert rel2 < 1e-4, f"all-reduced G vs sum of single-GPU shard gradients: rel L2 {rel2}"
    ok(rank, "reduced_gradients_equal_sum_of_shards", f"rel_l2={rel2:.2e}")

    # ---- 3: replicas stay bit-identical over eager + graph steps ---------------------------------------------------------
    dp2 = mdist.DataParallel()
    model2, eng2 = build(dev, dp2, graphs=True)
    opt2 = torch.optim.Adam(model2.parameters(), lr=5e-4, weight_decay=1e-4)
    for step in range(4):  # eager, eager, capture + replay, replay
        dd = O.synthetic_batch(B, 100 + step, hw)
        model2.train_step(make_batch(dd, lo, hi), opt2, LOSS, dev, None)
    torch.cuda.synchronize(dev)
    for nm in ("P", "M", "V", "Wb"):
        t = getattr(eng2.fs, nm)
        ref_t = t.clone()
        dist.broadcast(ref_t, src=0)
        assert torch.equal(ref_t, t), f"fs.{nm} differs between ranks after 4 steps"
    assert int(eng2.fs.step.item()) == 4
    ok(rank, "replicas_bit_identical_after_4_steps")

    # ---- 4: FedAvg across GPUs -------------------------------------------------------------------------------------------
    n_k = [1000.0 * (r + 1) for r in range(world)]
    client, ceng = build(dev, None, seed=100 + rank, graphs=False)
    with torch.no_grad():
        client.audio_encoder.bn1.running_mean.add_(0.1 * (rank + 1))
    states = []
    for r in range(world):
        torch.manual_seed(100 + r)
        st = O.init_avmnist_state()
        st["audio_encoder.bn1.running_mean"] = st["audio_encoder.bn1.running_mean"] + 0.1 * (r + 1)
        states.append(st)
    want = O.fedavg(states, n_k)
    fedavg.federated_allreduce(client, n_k[rank])
    sd = client.state_dict()
    for k, v in want.items():
        gotv = sd[k].detach().cpu()
        if v.dtype.is_floating_point:
            assert torch.allclose(gotv, v, rtol=1e-5, atol=1e-7), f"federated_allreduce: {k}"
        else:
            assert int(gotv) == int(v), k
    ok(rank, "federated_allreduce_equals_oracle_fedavg")

    dist.barrier()
    torch.cuda.synchronize(dev)
    if rank == 0:
        print("DIST_ALL_OK", flush=True)
    sys.stdout.flush()
    os._exit(0)  # NCCL communicators captured in CUDA graphs: leave without tearing the process group down (see bench.py)


if __name__ == "__main__":
    main()
