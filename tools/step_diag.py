"""GPU diagnostic: one fused train step vs the CPU oracle on identical weights / inputs; prints per-tensor errors."""
import os
import sys
import time
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import late_fusion_oracle as O  # noqa: E402
from mml_b200.avmnist import AVMNIST  # noqa: E402
from mml_b200.resnet import ResNet18, ResNet34  # noqa: E402


class Term:
    def __init__(self):
        self.loss_fn, self.weight = torch.nn.CrossEntropyLoss(), 1.0


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    aH, aW = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (112, 112)
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = AVMNIST(ResNet18(1, 64), ResNet34(1, 128), 128, dropout=0.5)
    torch.manual_seed(0)
    state = O.init_avmnist_state()
    model.to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    d = O.synthetic_batch(B, 1234, (aH, aW))
    batch = {"audio_original": d["audio"], "audio_missing_index": d["audio_mask"], "image_original": d["image"], "image_missing_index": d["image_mask"],
             "labels": d["labels"], "pattern_name": ["ai"] * B}
    A = O.apply_missing_mask(d["audio"], d["audio_mask"])
    I = O.apply_missing_mask(d["image"], d["image_mask"])
    opt_state = {}
    emu = os.environ.get("EMU", "0") == "1"
    for step in range(steps):
        t0 = time.time()
        out = model.train_step(batch, opt, {"cross_entropy": Term()}, dev, None, dropout_mask=d["dropout_mask"])
        torch.cuda.synchronize()
        t1 = time.time()
        ref = O.train_step(state, opt_state, A, I, d["labels"], d["dropout_mask"], 0.5, emulate_bf16=emu)
        plan = next(iter(model._engine.plans.values()))
        lg = plan.logits.cpu()
        print(f"step {step}: loss gpu {out['loss']:.6f} ref {ref['loss']:.6f} | max|dlogit| {(lg - ref['logits']).abs().max():.4e} (max|logit| {ref['logits'].abs().max():.3f}) "
              f"| gpu {1e3 * (t1 - t0):.1f} ms")
        if step == 0:
            worst = []
            for name, p in model.named_parameters():
                g = p.grad.detach().cpu().float()
                r = ref["grads"][name]
                rel = float((g - r).norm() / (r.norm() + 1e-20))
                cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-20))
                worst.append((rel, cos, name, float(r.norm())))
            worst.sort(reverse=True)
            for enc in ("audio_encoder.", "image_encoder.", "net."):
                ga = torch.cat([p.grad.detach().cpu().float().reshape(-1) for n, p in model.named_parameters() if n.startswith(enc)])
                ra = torch.cat([ref["grads"][n].reshape(-1) for n, _ in model.named_parameters() if n.startswith(enc)])
                print(f"  {enc:16s} rel L2 {float((ga - ra).norm() / ra.norm()):.4f} cosine {float((ga * ra).sum() / (ga.norm() * ra.norm())):.5f}")
            print("worst gradient tensors (rel L2 err, cosine, name, |ref|):")
            for w in worst[:25]:
                print("   %.4f  %.5f  %-50s %.3e" % w)
            rels = torch.tensor([w[0] for w in worst])
            print(f"grad rel-L2: median {rels.median():.4f} mean {rels.mean():.4f} max {rels.max():.4f}; n={len(worst)}")
            gall = torch.cat([p.grad.detach().cpu().float().reshape(-1) for _, p in model.named_parameters()])
            rall = torch.cat([ref["grads"][n].reshape(-1) for n, _ in model.named_parameters()])
            print(f"global grad: rel L2 {float((gall - rall).norm() / rall.norm()):.4f} cosine {float((gall * rall).sum() / (gall.norm() * rall.norm())):.5f}")
            # running stats
            sd = model.state_dict()
            for k in ("audio_encoder.bn1.running_mean", "audio_encoder.bn1.running_var", "image_encoder.layer4.2.bn2.running_var", "audio_encoder.layer4.1.bn2.running_mean"):
                e = float((sd[k].cpu() - state[k]).abs().max() / (state[k].abs().max() + 1e-12))
                print(f"   {k}: rel max err {e:.4e}; nbt {int(sd[k.rsplit('.', 1)[0] + '.num_batches_tracked'])}")
    # eval forward
    model.eval()
    with torch.no_grad():
        ev = model.forward(A=A.to(dev), I=I.to(dev)).cpu()
    evr = O.validation_step(state, A, I, d["labels"])["logits"]
    print(f"eval logits: max|d| {(ev - evr).abs().max():.4e} (max|ref| {evr.abs().max():.3f})")
    print("launches/step:", plan.launches_per_step)


if __name__ == "__main__":
    main()
