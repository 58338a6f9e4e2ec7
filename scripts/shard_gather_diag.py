"""Why is the sharded forward gather slow?  (torchrun, 2+ ranks)  Times ctr_gather_fwd on the sharded plan
for ids that are all local / all remote / mixed, with and without the linear (4-byte) tables."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from deepctr_torch_b200 import _lib, ops, sharded
from oracle import ctr_oracle as O


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    dist.init_process_group("nccl", device_id=torch.device(dev))
    B, V = 65536, 1538462
    cols = [O.sparse_col("C%d" % (i + 1), V, 16) for i in range(26)] + [O.dense_col("I%d" % (i + 1)) for i in range(13)]
    cfg = O.make_cfg("DeepFM", cols, cols, init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0, dnn_hidden_units=[256, 128])
    model, _ = sharded.build_sharded(cfg, dev, rank, world, batch=B)
    plan = model._plan
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    ids = torch.randint(0, V - world, (B, 26), device=dev, generator=g)
    dense = torch.rand(B, 13, device=dev, generator=g)
    variants = {"mixed (uniform ids)": ids,
                "all local": ids - ids % world + rank,
                "all remote": ids - ids % world + (rank + 1) % world}
    emb_ptrs, lin_ptrs = plan.table_ptrs()
    blk = torch.empty(B, plan.ld, device=dev)
    lin = torch.empty(B, device=dev)
    fm = torch.empty(B, device=dev)

    def run(X, n_lin, want_fm=True):
        _lib.call("ctr_gather_fwd", ops._ptr(X), X.stride(0), B, plan.n_emb, plan.D, ops._ptr(emb_ptrs),
                  ops._ptr(plan.emb_cols), ops._ptr(plan.emb_vocab), n_lin, ops._ptr(lin_ptrs), ops._ptr(plan.lin_cols),
                  ops._ptr(plan.lin_vocab), plan.n_dense, ops._ptr(plan.dense_cols), 0, ops._ptr(plan.lin_dense_cols), None,
                  ops._ptr(blk), plan.ld, ops._ptr(lin), ops._ptr(fm) if want_fm else None, ops._ptr(plan.err_flag),
                  plan.n_shards, ops._stream())

    for name, idv in variants.items():
        X = torch.cat([idv.float(), dense], 1).contiguous()
        for n_lin, var in ((plan.n_lin, 0), (0, 0), (0, 1), (0, 2), (0, 3), (0, 4), (plan.n_lin, 3)):
            os.environ["CTR_GATHER_VARIANT"] = str(var)
            dist.barrier()
            for _ in range(3):
                run(X, n_lin)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                run(X, n_lin)
            e1.record()
            torch.cuda.synchronize()
            print("rank %d  %-20s n_lin=%2d variant %d : %.3f ms per gather" % (rank, name, n_lin, var, e0.elapsed_time(e1) / 5), flush=True)
    # the same rows through torch's own gather on the peer-mapped table of field 0 (reference point)
    peer = (rank + 1) % world
    rows = plan.layout.emb_rows[0]
    raw = sharded._RawCuda(plan.arena.peer_ptr[peer] + plan.layout.emb_off[0], rows * 16 * 4)
    remote_tab = torch.as_tensor(raw, device=torch.device(dev)).view(torch.float32).view(rows, 16)
    idx = torch.randint(0, rows, (B * 13,), device=dev, generator=g)
    out = torch.empty(idx.numel(), 16, device=dev)
    torch.index_select(remote_tab, 0, idx, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        torch.index_select(remote_tab, 0, idx, out=out)
    e1.record()
    torch.cuda.synchronize()
    print("rank %d  torch.index_select of %d remote rows: %.3f ms" % (rank, idx.numel(), e0.elapsed_time(e1) / 5), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
