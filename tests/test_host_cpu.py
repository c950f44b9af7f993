"""CPU: host-side logic of the drop-in -- module surface / state_dict parity with the reference layout, flat parameter
store, synthetic batch contract, GEMM tiling heuristics, data-parallel helpers under gloo (world size 2)."""
import argparse
import os

import pytest
import torch


def make_args(**kw):
    a = dict(bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3,
             FOCAL_LAMBDA=0.01)
    a.update(kw)
    return argparse.Namespace(**a)


@pytest.fixture(scope="module")
def model():
    from tris_b200.model_stage1 import TRIS
    return TRIS(make_args())


def test_state_dict_matches_reference_layout(model):
    from oracle import weights as W
    ref = W.make_tris_state_dict(0)        # key/shape inventory pinned against the reference in test_oracle.py
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys()) or set(sd.keys()) == set(ref.keys())
    assert len(sd) == 518
    for k, v in ref.items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
    model.load_state_dict(ref, strict=True)


def test_trainable_parameter_groups(model):
    backbone, new = model.trainable_parameters()
    nb, nn_ = sum(p.numel() for p in backbone), sum(p.numel() for p in new)
    assert abs(nb / 1e6 - 102.01) < 0.01 and abs(nn_ / 1e6 - 11.55) < 0.01
    ids = {id(p) for p in backbone} | {id(p) for p in new}
    assert id(model.logit_scale) not in ids
    first_new = dict(model.named_parameters())["vis_project.weight"]
    assert new[0] is first_new                                                     # order: vis_project, lan_project, attn_fusion


def test_forward_argument_checks(model):
    with pytest.raises(ValueError):
        model(torch.zeros(2, 3, 320, 320), torch.zeros(2, 19, dtype=torch.int32))
    with pytest.raises(ValueError):
        model(torch.zeros(2, 3, 300, 320), torch.zeros(2, 20, dtype=torch.int32))
    from tris_b200.model_stage1 import TRIS
    with pytest.raises(ValueError):
        TRIS(make_args(backbone="clip-ViT-B/16"))


def test_param_store_layout_cpu(model):
    from tris_b200.engine import _ADJ, _tris_group
    from tris_b200.store import ParamStore
    import copy
    m = copy.deepcopy(model)
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    st = ParamStore(m, "cpu", _tris_group, _ADJ)
    assert st.group_bounds[0] == 0 and st.group_bounds[1] < st.group_bounds[2] <= st.total
    assert abs(st.group_bounds[2] / 1e6 - 98.77) < 0.05                             # gradient payload of the all-reduce
    for k, p in m.named_parameters():                                               # parameters became views of the flat buffer
        assert torch.equal(p.detach(), before[k]) and p.data_ptr() == st.p(k).data_ptr()
        assert st.offsets[k] % 8 == 0
    w = st.cat("flat", [f"attn_fusion.v_proj{i}.0.weight" for i in (1, 2, 3)], (3072, 1024))
    assert torch.equal(w[1024:2048], before["attn_fusion.v_proj2.0.weight"].reshape(1024, 1024))
    assert all(_tris_group(k) == 2 for k in before if "attnpool" in k) and _tris_group("logit_scale") == 2
    st.publish_grads()
    assert dict(m.named_parameters())["vis_project.weight"].grad.data_ptr() == st.g("vis_project.weight").data_ptr()
    assert dict(m.named_parameters())["backbone.visual.attnpool.q_proj.weight"].grad is None


def test_synthetic_batch_contract():
    from tris_b200.synthetic import EOT, SOT, synthetic_batch
    img, ids, neg = synthetic_batch(6, 64, 20, 3, seed=5)
    assert img.shape == (6, 3, 64, 64) and img.dtype == torch.float32
    assert ids.shape == (6, 20) and ids.dtype == torch.int32 and neg.shape == (6, 3, 20)
    for row in torch.cat([ids, neg.reshape(-1, 20)]):
        assert row[0] == SOT and (row == EOT).sum() == 1
        e = int(row.argmax())
        assert 4 <= e <= 18 and torch.all(row[e + 1:] == 0) and torch.all(row[1:e] > 0)
    img2, ids2, _ = synthetic_batch(6, 64, 20, 3, seed=5)
    assert torch.equal(img, img2) and torch.equal(ids, ids2)


def test_gemm_tiling_heuristics():
    from tris_b200 import gemm as G
    assert G.conv_tile(10, 10) == (10, 10) and G.conv_tile(80, 80)[0] * G.conv_tile(80, 80)[1] <= 128
    th, tw = G.conv_tile(20, 20, max_rows=96, mult=16)
    assert th * tw % 16 == 0 and th * tw <= 96
    for n in (48, 64, 512, 3072):
        assert G._bn_for(n, 38) in (32, 64, 128, 256)
    assert G._bn_for(24, 8) == 32 and G._bn_for(24, 8, True) == 64
    assert G._split_for(400, 100) == 1 and G._split_for(8, 4800) > 1


def test_clip_load_aliases():
    from tris_b200 import clip_model
    assert clip_model._canonical("ViT-B-32") == clip_model._canonical("ViT-B/32") == "ViT-B/32"
    assert clip_model._canonical("RN50") == "RN50"
    with pytest.raises(RuntimeError):
        clip_model._canonical("nope")


# ------------------------------------------------------------------------------------------------ world size 2 (gloo)
def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from tris_b200 import dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert dp.env_rank() == (rank, rank, world) and dp.world_size() == world
        flat = torch.full((1000,), float(rank + 1))
        dp.broadcast_parameters(flat)
        assert torch.all(flat == 1.0)                                # rank 0's values everywhere
        grad = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        scale = dp.all_reduce_gradients(grad)
        want = torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world))
        assert torch.equal(grad, want) and scale == 1.0 / world       # SUM on the wire, mean folded into AdamW
        assert dp.max_over_ranks(float(rank), "cpu") == world - 1
        seeds = {dp.shard_seed(1234, r, i) for r in range(world) for i in range(3)}
        assert len(seeds) == 3 * world
        assert sorted(list(dp.shard_range(10, 0, world)) + list(dp.shard_range(10, 1, world))) == list(range(10))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_helpers_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_average_of_rank_gradients_equals_big_batch_mean():
    """The DP contract: mean over ranks of per-rank mean-loss gradients == gradient of the mean loss over the union
    batch when the per-sample terms are rank-local (true for every Stage-1 term: the in-batch contrast is rank-local by
    construction).  Checked on a toy rank-local loss with the same reduction algebra."""
    torch.manual_seed(0)
    w = torch.randn(16, requires_grad=True)
    xs = [torch.randn(8, 16) for _ in range(2)]
    per_rank = []
    for x in xs:
        (g,) = torch.autograd.grad(torch.tanh(x @ w).mean(), w)
        per_rank.append(g)
    (g_all,) = torch.autograd.grad(torch.tanh(torch.cat(xs) @ w).mean(), w)
    assert torch.allclose((per_rank[0] + per_rank[1]) * 0.5, g_all, atol=1e-6)


def test_cam_writer_npy_bytes_match_numpy_save(tmp_path):
    """validate.py --save_cam writes `{idx}_{img_id}.npy` with three raw system calls (header once per shape + payload): the file
    must be byte-identical to what the reference's np.save produces (validate.py:354-359), for the shapes / dtypes it dumps."""
    import os
    import numpy as np
    from tris_b200.infer import npy_header
    for shape, dt in (((480, 640), np.float32), ((1, 1, 320, 320), np.float32), ((7,), np.float64)):
        a = np.random.default_rng(0).random(shape).astype(dt)
        p1, p2 = str(tmp_path / "a.npy"), str(tmp_path / "b.npy")
        np.save(p1, a)
        fd = os.open(p2, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        os.write(fd, npy_header(a))
        os.write(fd, memoryview(a).cast("B"))
        os.close(fd)
        assert open(p1, "rb").read() == open(p2, "rb").read()
        assert np.array_equal(np.load(p2), a)
