"""Host logic of the row-sharded multi-GPU path, on CPU with gloo (world_size 2): shard layout
arithmetic, arena layout symmetry, gather-on-save / scatter-on-load of reference-compatible
state_dicts through torch.distributed."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepctr_torch_b200 import sharded


def test_shard_unshard_roundtrip():
    for vocab in (1, 2, 7, 8, 1000, 1538462):
        for world in (1, 2, 3, 8):
            full = torch.arange(vocab * 2, dtype=torch.float32).view(vocab, 2)
            shards = [sharded.shard_rows(full, r, world) for r in range(world)]
            assert [s.shape[0] for s in shards] == [sharded.local_rows(vocab, r, world) for r in range(world)]
            assert sum(s.shape[0] for s in shards) == vocab
            assert max(s.shape[0] for s in shards) == sharded.max_local_rows(vocab, world) or vocab < world
            for r, s in enumerate(shards):       # local index = id // world, owner = id % world
                for j in range(0, s.shape[0], max(1, s.shape[0] // 5)):
                    assert s[j, 0].item() == (j * world + r) * 2
            assert torch.equal(sharded.unshard_rows(shards, vocab), full)


def test_arena_layout_is_rank_independent_and_disjoint():
    L = sharded.ArenaLayout([1000, 77, 5], [1000, 77], dim=16, world=4, batch=64)
    spans = []
    for f in range(3):
        spans.append((L.emb_off[f], L.emb_rows[f] * 16 * 4))
    for f in range(2):
        spans.append((L.lin_off[f], L.lin_rows[f] * 4))
    for par in range(2):
        r = L.recv[par]
        spans += [(r["count"], 5 * 4), (r["ids"], 5 * L.cap * 4), (r["emb"], 3 * L.cap * 16 * 4), (r["lin"], 2 * L.cap * 4)]
    spans.sort()
    for (o1, n1), (o2, _) in zip(spans, spans[1:]):
        assert o1 + n1 <= o2
        assert o1 % 256 == 0
    assert spans[-1][0] + spans[-1][1] <= L.nbytes
    assert L.cap == 64 * 4


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        vocabs = {"embedding_dict.C1.weight": 11, "linear_model.embedding_dict.C1.weight": 11,
                  "embedding_dict.C2.weight": 4}
        full = {"embedding_dict.C1.weight": torch.randn(11, 8, generator=g),
                "linear_model.embedding_dict.C1.weight": torch.randn(11, 1, generator=g),
                "embedding_dict.C2.weight": torch.randn(4, 8, generator=g),
                "dnn.linears.0.weight": torch.randn(5, 3, generator=g)}
        local = sharded.scatter_full_state_dict(full, vocabs, rank, world)
        assert local["embedding_dict.C1.weight"].shape[0] == sharded.max_local_rows(11, world)
        assert torch.equal(local["embedding_dict.C1.weight"][:sharded.local_rows(11, rank, world)],
                           full["embedding_dict.C1.weight"][rank::world])
        back = sharded.gather_full_state_dict(local, vocabs)
        ok = all(torch.equal(back[k], full[k]) for k in full)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_state_dict_gather_scatter_gloo_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_localize_cfg_handles_columns_shared_between_linear_and_dnn():
    """make_cfg(model, cols, cols) lists the SAME column dicts twice; each must be localised once."""
    from oracle import ctr_oracle as O
    cols = [O.sparse_col("C%d" % i, 1003 + 7 * i, 16) for i in range(3)] + [O.dense_col("I0")]
    cfg = O.make_cfg("DeepFM", cols, cols, dnn_hidden_units=[8])
    for world in (2, 8):
        loc = sharded.localize_cfg(cfg, world)
        for c_full, c_loc in zip(cfg["dnn_columns"], loc["dnn_columns"]):
            if c_full["type"] == "sparse":
                assert c_loc["vocab"] == sharded.max_local_rows(c_full["vocab"], world)
        for c_full, c_loc in zip(cfg["linear_columns"], loc["linear_columns"]):
            if c_full["type"] == "sparse":
                assert c_loc["vocab"] == sharded.max_local_rows(c_full["vocab"], world)
        assert cfg["dnn_columns"][0]["vocab"] == 1003           # the logical cfg is untouched


@pytest.mark.parametrize("world", [2, 8])
def test_arena_layout_regions_are_disjoint_and_aligned(world):
    """Tables, the two receive lists and the forward-exchange buffers of the peer arena must not overlap
    (every rank derives peer addresses from these offsets alone) and must keep 16-byte alignment."""
    F, D, B = 26, 16, 65536
    vocabs = [1538462] * F
    L = sharded.ArenaLayout(vocabs, vocabs, D, world, B, n_id_cols=F)
    regions = []
    for f in range(F):
        regions.append((L.emb_off[f], L.emb_rows[f] * D * 4))
        regions.append((L.lin_off[f], L.lin_rows[f] * 4))
    nf = 2 * F
    for par in range(2):
        r = L.recv[par]
        regions += [(r["count"], nf * 4), (r["ids"], nf * L.cap * 4), (r["emb"], F * L.cap * D * 4), (r["lin"], F * L.cap * 4)]
    x = L.x
    regions += [(x["req_cnt"], world * 4), (x["req"], world * L.xcap * 8), (x["resp_emb"], world * L.xcap * D * 4),
                (x["resp_lin"], world * L.xcap * 4)]
    regions.sort()
    for (o0, n0), (o1, _) in zip(regions, regions[1:]):
        assert o0 % 16 == 0 and o0 + n0 <= o1, (o0, n0, o1)
    assert regions[-1][0] + regions[-1][1] <= L.nbytes
    # capacity: a peer can always hold twice the uniform share of one rank's ids (all of them for 2 ranks)
    full = B * F
    assert L.xcap >= min(full, 2 * ((full + world - 1) // world))
    assert L.xcap < (1 << 26) and L.cap < (1 << 26)          # (owner << 26) | slot encodings
    # rows per shard cover the vocabulary
    for v, rows in zip(vocabs, L.emb_rows):
        assert rows * world >= v
