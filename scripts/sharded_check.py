"""Correctness of the row-sharded multi-GPU path (run under torchrun with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/sharded_check.py

Every rank builds the SAME full model (seeded) as an unsharded reference on its own GPU and a sharded
copy (tables row-sharded over the ranks, peer-mapped).  Per rank-local batch it checks that
  * the sharded forward (remote rows read over NVLink) gives bit-identical logits,
  * the dense-parameter gradients match,
  * the row gradients delivered to each owner, scattered into a dense shard, equal the sum over
    ranks of the reference's dense table gradients restricted to the owner's rows.
"""
import copy
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from deepctr_torch_b200 import sharded
from deepctr_torch_b200.config import model_from_cfg
from oracle import ctr_oracle as O


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    B, V, D = 4096, 1003, 16
    cols = [O.sparse_col("C%d" % i, V + 7 * i, D) for i in range(26)] + [O.dense_col("I%d" % i) for i in range(13)]
    cfg = O.make_cfg("DeepFM", cols, cols, dnn_hidden_units=[64, 32], init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0)
    ref = model_from_cfg(cfg, "cpu", table_grad="dense")          # seeded: identical on every rank
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    full_state = {k: v.clone() for k, v in ref.state_dict().items()}
    ref = model_from_cfg(cfg, dev, table_grad="dense")
    ref.load_state_dict(full_state)
    ref.train()

    table_vocab = {}
    for c in cols:
        if c["type"] == "sparse":
            table_vocab["embedding_dict.%s.weight" % c["name"]] = c["vocab"]
            table_vocab["linear_model.embedding_dict.%s.weight" % c["name"]] = c["vocab"]
    local_cfg = sharded.localize_cfg(cfg, world)
    sh = model_from_cfg(local_cfg, dev, table_grad="rowwise")
    sh.load_state_dict({k: v.to(dev) for k, v in sharded.scatter_full_state_dict(full_state, table_vocab, rank, world).items()})
    sharded.attach_shards(sh, cfg, rank, world, batch=B)
    sh.train()

    X, y = O.synthetic_batch(cfg, B, seed=100 + rank, zipf_alpha=1.05 if rank % 2 else None)
    X, y = X.to(dev), y.to(dev)
    bce = torch.nn.functional.binary_cross_entropy

    yr = ref(X)
    bce(yr.squeeze(1), y, reduction="sum").backward()
    ys = sh(X)
    bce(ys.squeeze(1), y, reduction="sum").backward()
    sh.check_ids()
    ok_logit = torch.equal(yr.detach(), ys.detach())
    err_logit = float((yr - ys).abs().max())

    table_ids = set(id(p) for p in sh._plan.emb_params + sh._plan.lin_params)
    worst_dense = 0.0
    ref_named = dict(ref.named_parameters())
    for k, p in sh.named_parameters():
        if id(p) in table_ids:
            continue
        d = float((p.grad - ref_named[k].grad).abs().max() / (ref_named[k].grad.abs().max() + 1e-30))
        worst_dense = max(worst_dense, d)

    parity = sh.sharded.finish_step()             # all-reduce of dense grads = barrier for the pushes
    counts, ids, emb_rows, lin_rows = sh.sharded.received_row_grads(parity)
    torch.cuda.synchronize()
    worst_rows = 0.0
    sparse = [c for c in cols if c["type"] == "sparse"]
    for f, c in enumerate(sparse):
        for kind, key, rows, nf in (("emb", "embedding_dict.%s.weight", emb_rows, 0),
                                    ("lin", "linear_model.embedding_dict.%s.weight", lin_rows, len(sparse))):
            gfull = ref_named[key % c["name"]].grad.clone()
            dist.all_reduce(gfull)                                   # sum over the ranks' local batches
            expect = sharded.shard_rows(gfull, rank, world)
            n = int(counts[nf + f])
            got = torch.zeros_like(expect)
            src = rows[f][:n] if kind == "emb" else rows[f][:n].unsqueeze(1)
            got.index_add_(0, ids[nf + f][:n].long(), src)
            d = float((got - expect).abs().max() / (expect.abs().max() + 1e-30))
            worst_rows = max(worst_rows, d)
    res = torch.tensor([1.0 if ok_logit else 0.0, err_logit, worst_dense, worst_rows], device=dev)
    allres = [torch.empty_like(res) for _ in range(world)]
    dist.all_gather(allres, res)
    if rank == 0:
        for r, t in enumerate(allres):
            print("rank %d: logits bit-identical=%d (max abs diff %.2e), dense-grad rel err %.2e, "
                  "delivered row-grad rel err %.2e" % (r, int(t[0]), t[1], t[2], t[3]))
        good = all(t[0] == 1.0 and t[2] < 1e-5 and t[3] < 1e-4 for t in allres)
        print("SHARDED CHECK", "PASSED" if good else "FAILED")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
