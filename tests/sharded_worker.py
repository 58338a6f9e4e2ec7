"""Worker of the multi-GPU parity test (tests/test_gpu_sharded.py spawns it under torchrun; bench.py
calls ``run_check`` once before it times a multi-GPU run).

Every rank builds the SAME seeded full model on the CPU (-> oracle parameters), a row-sharded copy
on its GPU, and draws its own local batch.  Checked per rank, against the CPU ORACLE (not against
this repo's own single-GPU path):

  A  sharded forward logits (rows of other ranks arrive through the NVLink exchange) vs the oracle;
  B  the owner-side COMBINED row gradients (``ShardedRuntime.combine_received``: the lists delivered by all
     ranks, duplicates summed) scattered into a dense shard vs the oracle's dense table gradients summed
     over all ranks' batches and restricted to the rows this rank owns; dense-parameter gradients vs the
     all-reduced oracle gradients;
  C  one fused optimizer step on the shards (sgd / adagrad / adam) vs the same optimizer applied on the CPU
     to the oracle gradients of the global batch (touched rows only);
for a uniform batch and for a SKEWED batch in which every id of the even columns is owned by rank 0
(worst case of the exchange buffers).
"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from deepctr_torch_b200 import sharded                      # noqa: E402
from deepctr_torch_b200.config import model_from_cfg       # noqa: E402
from oracle import ctr_oracle as O                          # noqa: E402

LOGIT_TOL, GRAD_TOL, STEP_TOL = 1e-5, 1e-4, 2e-4


def _rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _cpu_optimizer_step(kind, w, g, touched, state, step, l2x2=0.0):
    """The row-wise ("lazy") rule of csrc/rowopt.cu on the CPU, applied to the touched rows only."""
    idx = touched.nonzero().reshape(-1)
    gg = g[idx] + l2x2 * w[idx]
    if kind == "sgd":
        w[idx] -= 0.01 * gg
    elif kind == "adagrad":
        state.setdefault("s", torch.zeros_like(w))
        state["s"][idx] += gg * gg
        w[idx] -= 0.01 * gg / (state["s"][idx].sqrt() + 1e-10)
    elif kind == "adam":
        state.setdefault("m", torch.zeros_like(w))
        state.setdefault("v", torch.zeros_like(w))
        state["m"][idx] = 0.9 * state["m"][idx] + 0.1 * gg
        state["v"][idx] = 0.999 * state["v"][idx] + 0.001 * gg * gg
        bc1, bc2 = 1 - 0.9 ** step, 1 - 0.999 ** step
        w[idx] -= (0.001 / bc1) * state["m"][idx] / (state["v"][idx].sqrt() / bc2 ** 0.5 + 1e-8)


def run_check(dev, rank, world, B=2048, V=1003, D=16, optimizer="adagrad", verbose=False):
    """Returns {name: worst relative error}; all ranks must call it together."""
    torch.manual_seed(1234)
    cols = [O.sparse_col("C%d" % i, V + 7 * i, D) for i in range(26)] + [O.dense_col("I%d" % i) for i in range(13)]
    cfg = O.make_cfg("DeepFM", cols, cols, dnn_hidden_units=[64, 32], init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0)
    ref = model_from_cfg(cfg, "cpu", table_grad="dense")          # seeded: identical on every rank
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    full_state = {k: v.clone() for k, v in ref.state_dict().items()}
    table_vocab = {}
    for c in cols:
        if c["type"] == "sparse":
            table_vocab["embedding_dict.%s.weight" % c["name"]] = c["vocab"]
            table_vocab["linear_model.embedding_dict.%s.weight" % c["name"]] = c["vocab"]
    local_cfg = sharded.localize_cfg(cfg, world)
    sh = model_from_cfg(local_cfg, dev, table_grad="rowwise")
    sh.load_state_dict({k: v.to(dev) for k, v in sharded.scatter_full_state_dict(full_state, table_vocab, rank, world).items()})
    sharded.attach_shards(sh, cfg, rank, world, batch=B)
    sh.compile(optimizer, "binary_crossentropy")
    sh.train()
    bce = torch.nn.functional.binary_cross_entropy
    worst = {"logit": 0.0, "dense_grad": 0.0, "combined_rowgrad": 0.0, "optimizer_step": 0.0}
    opt_state = {}
    cur = {k: v.clone() for k, v in full_state.items()}     # CPU copy of the full model, stepped alongside
    sparse = [c for c in cols if c["type"] == "sparse"]
    for it, skew in enumerate((False, True, False)):
        X, y = O.synthetic_batch(cfg, B, seed=100 + 17 * it + rank, zipf_alpha=1.05 if rank % 2 else None)
        if skew:                                   # every id of the even columns lives on rank 0
            X[:, 0:26:2] = (X[:, 0:26:2] / world).floor() * world
        ref_logit, _, _, ref_grads = O.loss_and_grads(cfg, cur, X, y)
        # global-batch gradients = sum over ranks (loss reduction is 'sum')
        names = sorted(ref_grads)
        flat = torch.cat([ref_grads[k].reshape(-1) for k in names]).to(dev)
        dist.all_reduce(flat)
        flat = flat.cpu()
        tot, off = {}, 0
        for k in names:
            n = ref_grads[k].numel()
            tot[k] = flat[off:off + n].view_as(ref_grads[k])
            off += n
        # which rows did ANY rank touch (the lazy optimizer only moves those)
        touched = {}
        for f, c in enumerate(sparse):
            t = torch.zeros(c["vocab"], device=dev)
            t[X[:, f].long().to(dev)] = 1.0
            dist.all_reduce(t)
            touched[c["name"]] = (t > 0).cpu()

        sh.optim.zero_grad()
        from helpers import capture_logit
        y_pred, logit = capture_logit(sh, X.to(dev))
        bce(y_pred.squeeze(1), y.to(dev), reduction="sum").backward()
        done = sh.sharded.finish_step()
        sh.check_ids()
        worst["logit"] = max(worst["logit"], _rel(logit.cpu(), ref_logit))
        table_ids = set(id(p) for p in sh._plan.emb_params + sh._plan.lin_params)
        for k, p in sh.named_parameters():
            if id(p) not in table_ids:
                worst["dense_grad"] = max(worst["dense_grad"], _rel(p.grad.cpu(), tot[k]))
        ws = sh.sharded.combine_received(done)
        torch.cuda.synchronize()
        nf = len(sparse)
        for f, c in enumerate(sparse):
            # receive lists (and the owner-side plan) belong to the id columns of the plan, shared by the
            # embedding and the linear table of a feature
            for key, buf, pc in (("embedding_dict.%s.weight", ws["comb_emb"][f], sh._plan.emb_plan_col_host[f]),
                                 ("linear_model.embedding_dict.%s.weight", ws["comb_lin"][f].unsqueeze(1),
                                  sh._plan.lin_plan_col_host[f])):
                expect = sharded.shard_rows(tot[key % c["name"]], rank, world)
                n = int(ws["n_uniq"][pc])
                got = torch.zeros_like(expect)
                rows = ws["uniq"][pc][:n].long().cpu()
                assert rows.unique().numel() == n, "combined rows are not duplicate-free"
                got[rows] = buf[:n].cpu()
                worst["combined_rowgrad"] = max(worst["combined_rowgrad"], _rel(got, expect))
        # C: the fused optimizer step on the shards vs the CPU rule on the global gradients
        sh.optim.step()
        for c in sparse:
            for key in ("embedding_dict.%s.weight", "linear_model.embedding_dict.%s.weight"):
                k = key % c["name"]
                _cpu_optimizer_step(optimizer, cur[k], tot[k], touched[c["name"]], opt_state.setdefault(k, {}), it + 1)
        with torch.no_grad():                      # dense parameters: the torch optimizer of the same family
            for k, p in sh.named_parameters():
                if id(p) not in table_ids:
                    cur[k] = p.detach().cpu().clone()
        local = {k: v for k, v in sh.state_dict().items()}
        for c in sparse:
            for key in ("embedding_dict.%s.weight", "linear_model.embedding_dict.%s.weight"):
                k = key % c["name"]
                expect = sharded.shard_rows(cur[k], rank, world)
                got = local[k][:expect.shape[0]].cpu()
                worst["optimizer_step"] = max(worst["optimizer_step"], _rel(got, expect))
        if verbose and rank == 0:
            print("  batch %d (skew=%s): %s" % (it, skew, {k: "%.2e" % v for k, v in worst.items()}), flush=True)
    res = torch.tensor([worst[k] for k in sorted(worst)], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    return dict(zip(sorted(worst), [float(v) for v in res.cpu()]))


def passed(w):
    return (w["logit"] <= LOGIT_TOL and w["dense_grad"] <= GRAD_TOL and w["combined_rowgrad"] <= GRAD_TOL and
            w["optimizer_step"] <= STEP_TOL)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    sys.path.insert(0, os.path.join(REPO, "tests"))
    ok = True
    for opt in ("sgd", "adagrad", "adam"):
        w = run_check(dev, rank, world, optimizer=opt, verbose=True)
        if rank == 0:
            print("world %d optimizer %-8s worst relative errors vs oracle: %s" % (world, opt, w), flush=True)
        ok = ok and passed(w)
    if rank == 0:
        print("SHARDED CHECK", "PASSED" if ok else "FAILED", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
