"""Worker of tests/test_gpu_dp.py (one process per GPU, NCCL): three data-parallel training steps in fp32 mode.

Checks, on rank 0 after gathering:  (1) the replicas' parameters and Adam moments are BIT-identical after every rank applied
the all-reduced gradient;  (2) they equal the CPU oracle's step on the MEAN of the ranks' gradients (each replica with its
own batch-norm batch statistics) within 2e-6;  (3) each replica's batch-norm moving statistics equal the
oracle's for that replica's batches."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200 as tb  # noqa: E402
from oracle import tacotron_oracle as O  # noqa: E402


def batch(rank, N=4, Ti=21, To=30):
    g = torch.Generator().manual_seed(100 + rank)
    L = torch.tensor([Ti, 9 + rank, 14, 5][:N], dtype=torch.int32)
    inp = torch.randint(2, 80, (N, Ti), generator=g, dtype=torch.int32)
    for n in range(N):
        inp[n, L[n] - 1] = 1
        inp[n, L[n]:] = 0
    return dict(inputs=inp, input_lengths=L, mel_targets=torch.rand(N, To, 80, generator=g),
                linear_targets=torch.rand(N, To, 1025, generator=g), loss_coeff=torch.rand(N, generator=g) + 0.5)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    hp = tb.hparams.override(reduction_factor=5, model_type="deepvoice")
    S = 3
    named = tb.params.init_params(hp, S, seed=77, randomize_bn_state=True)
    eng = tb.Engine(hp, S, precision=os.environ.get("DP_PRECISION", "fp32"), device=rank, named_params=named)
    steps = 3
    spk = lambda r: torch.tensor([0, 2, 1, r % S], dtype=torch.int32)
    mine = dict(batch(rank), speaker_id=spk(rank))

    # the product's collective: two-bucket all-reduce, the early bucket beside the encoder's backward pass (dist.py)
    from importlib import import_module
    allreduce = import_module("multi-speaker-tacotron-tensorflow_b200.dist").OverlappedAllReduce(eng)
    assert allreduce.early[0] > 0 and allreduce.early[1] > 0 and sum(allreduce.early) == eng.layout.n_trainable

    first_grad = None
    for i in range(steps):
        eng.train_step(mine, allreduce=allreduce)
        if i == 0:
            first_grad = {k: (t * (1.0 / world)).cpu() for k, t in eng.named_gradients().items()}      # eng.grads holds the all-reduced SUM
    torch.cuda.synchronize()
    flat = torch.cat([eng.params, eng.adam_m, eng.adam_v])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    bn = [torch.empty_like(eng.bn_state) for _ in range(world)]
    dist.all_gather(bn, eng.bn_state)
    if rank == 0:
        for r in range(1, world):
            assert torch.equal(gathered[0], gathered[r]), "replica %d diverged from replica 0 (max %g)" % (r, (gathered[0] - gathered[r]).abs().max().item())
        # ---- oracle: mean of the per-replica gradients, clip, Adam; batch-norm state per replica
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        names = [k for k in named if not k.endswith(("moving_mean", "moving_var"))]
        P = {k: v.clone() for k, v in named.items()}
        bn_rep = [{k: v.clone() for k, v in named.items() if k not in names} for _ in range(world)]
        m = {k: torch.zeros_like(P[k]) for k in names}; v = {k: torch.zeros_like(P[k]) for k in names}
        for step in range(steps):
            gsum = {k: torch.zeros_like(P[k]) for k in names}
            for r in range(world):
                b = dict(batch(r), speaker_id=spk(r))
                Pr = dict(P); Pr.update(bn_rep[r])
                leaf = {k: (Pr[k].detach().clone().requires_grad_(True) if k in names else Pr[k]) for k in Pr}
                out = O.forward(leaf, hp, b["inputs"], b["input_lengths"], S, b["speaker_id"], b["mel_targets"], b["linear_targets"], speaker_mode="deepvoice")
                ls = O.losses(out, b["mel_targets"], b["linear_targets"], b["loss_coeff"], hp)
                gl = torch.autograd.grad(ls["loss"], [leaf[k] for k in names], allow_unused=True)
                for k, g in zip(names, gl):
                    if g is not None:
                        gsum[k] += g
                bn_rep[r] = {k: t.detach() for k, t in out["new_bn_state"].items()}
            mean = {k: gsum[k] / world for k in names}
            if step == 0:
                # the all-reduced gradient itself (Adam's sign-like first update hides gradient errors): whole-gradient cosine and norm
                a = torch.cat([first_grad[k].reshape(-1) for k in names]).double(); r_ = torch.cat([mean[k].reshape(-1) for k in names]).double()
                cos = float(a @ r_ / (a.norm() * r_.norm()))
                fp = eng.precision == "fp32"
                assert cos >= (0.99999 if fp else 0.9967), "all-reduced gradient: cosine %g to the oracle's mean gradient" % cos
                assert abs(float(a.norm() - r_.norm())) <= (1e-4 if fp else 9e-3) * float(r_.norm())
                print("DP_GRAD 1-cos=%.3g |g|=%.6f oracle=%.6f" % (1 - cos, float(a.norm()), float(r_.norm())), flush=True)
            clipped, gn = O.clip_by_global_norm(mean, 1.0)
            lr = O.learning_rate(hp, step, True)
            sub, m, v = O.adam_step({k: P[k].detach() for k in names}, clipped, m, v, step + 1, lr, hp.adam_beta1, hp.adam_beta2)
            P.update(sub)
        got = eng.named_parameters()
        worst = max((got[k].cpu() - P[k]).abs().max().item() for k in names)
        tol = 2e-6 if eng.precision == "fp32" else 2e-4
        assert worst <= tol, "parameters differ from the oracle's mean-gradient step by %g" % worst
        lay = eng.layout
        for r in range(world):
            views = tb.params.views(eng.params, bn[r], lay)
            wb = max((views[k].cpu() - bn_rep[r][k]).abs().max().item() for k in bn_rep[r])
            assert wb <= (5e-6 if eng.precision == "fp32" else 5e-3), "replica %d batch-norm moving statistics differ by %g" % (r, wb)
        print("DP_OK world=%d steps=%d worst_param_diff=%.3g grad_norm=%.6f" % (world, steps, worst, gn), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
