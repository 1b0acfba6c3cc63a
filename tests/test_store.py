"""CPU (+ one GPU case): packed feature store round trip, sharded loader coverage, id gathers."""
import pytest
import torch

from fashionern_aaai2024_b200 import synthetic as syn
from fashionern_aaai2024_b200.sharded import shard_bounds
from fashionern_aaai2024_b200.store import FeatureStore


def make(tmp_path, n=1000, dim=64, patches=13):
    feats = syn.features(1, n, dim, unit=True)
    local = syn.patch_features(2, n, dim, patches)
    names = syn.caption_names(3, n, 50)
    return FeatureStore.save(str(tmp_path / "store"), feats, names, local, chunk_rows=300), feats, local, names


def test_round_trip_is_bit_exact_bf16(tmp_path):
    st, feats, local, names = make(tmp_path)
    assert (st.rows, st.dim, st.patches) == (1000, 64, 13) and st.names == names
    whole, off = st.load_shard(0, 1, device="cpu", chunk_rows=128)
    assert off == 0 and torch.equal(whole, feats.bfloat16())
    rows = [5, 999, 0, 5]
    assert torch.equal(st.gather(rows, device="cpu"), feats.bfloat16()[rows])
    assert torch.equal(st.load_local(rows, device="cpu"), local.bfloat16()[rows])
    m = st.name_to_row()
    assert all(names[m[nm]] == nm for nm in set(names)) and m[names[-1]] == 999   # repeated name -> last row


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shards_tile_the_gallery(tmp_path, world):
    st, feats, *_ = make(tmp_path, n=1001)
    parts, offs = zip(*[st.load_shard(r, world, device="cpu", chunk_rows=97) for r in range(world)])
    assert [o for o in offs] == [shard_bounds(1001, world, r)[0] for r in range(world)]
    assert torch.equal(torch.cat(parts), feats.bfloat16())


def test_empty_and_bad_store(tmp_path):
    st = FeatureStore.save(str(tmp_path / "e"), torch.zeros(0, 64))
    g, off = st.load_shard(0, 2, device="cpu")
    assert g.shape == (0, 64) and off == 0
    with pytest.raises(ValueError):
        st.load_local([0])
    with pytest.raises(ValueError):
        FeatureStore.save(str(tmp_path / "b"), torch.zeros(3, 8), names=["a"])


def test_import_of_the_reference_per_image_patch_files(tmp_path):
    # the reference keeps one torch.save()d [13, D] tensor per image (dataloader/fashioniq.py:69-70, cirr.py:55-56)
    n, dim = 37, 64
    feats = syn.features(4, n, dim)
    local = syn.patch_features(5, n, dim)
    names = syn.unique_names(n)
    src = tmp_path / "fashion_local13"
    src.mkdir()
    for i, nm in enumerate(names):
        torch.save(local[i] if i % 2 else local[i][None], str(src / f"{nm}.pth"))       # [13,D] and [1,13,D] both occur
    st = FeatureStore.import_patch_dir(str(tmp_path / "packed"), str(src), names, feats, workers=3, chunk_rows=10)
    assert (st.rows, st.dim, st.patches) == (n, dim, 13) and st.names == names
    rows = [36, 0, 7]
    assert torch.equal(st.load_local(rows, device="cpu"), local.bfloat16()[rows])
    assert torch.equal(st.gather(rows, device="cpu"), feats.bfloat16()[rows])
    with pytest.raises(FileNotFoundError):
        FeatureStore.import_patch_dir(str(tmp_path / "bad"), str(src), names + ["missing"], torch.zeros(n + 1, dim))
    torch.save(torch.zeros(5, dim), str(src / "short.pth"))
    with pytest.raises(ValueError):
        FeatureStore.import_patch_dir(str(tmp_path / "bad2"), str(src), names + ["short"], torch.zeros(n + 1, dim))


@pytest.mark.gpu
def test_shard_feeds_the_scoring_kernel(tmp_path, cuda_device):
    from fashionern_aaai2024_b200 import ops
    n, dim, k = 5000, 640, 20
    feats = syn.features(9, n, dim, unit=True)
    st = FeatureStore.save(str(tmp_path / "g"), feats)
    pred = syn.features(10, 64, dim, unit=True).bfloat16().to(cuda_device)
    full = ops.sim_topk(pred, feats.bfloat16().to(cuda_device), k, want_keys=True)[2]
    parts = []
    for r in range(3):
        shard, off = st.load_shard(r, 3, device=cuda_device, chunk_rows=512)
        parts.append(ops.sim_topk(pred, shard, k, id_offset=off, want_keys=True)[2])
    assert torch.equal(ops.topk_merge(torch.stack(parts), k)[2], full)


@pytest.mark.gpu
def test_stream_topk_equals_resident(tmp_path, cuda_device):
    from fashionern_aaai2024_b200 import ops
    from fashionern_aaai2024_b200.store import stream_topk
    n, dim, k = 9001, 640, 50
    feats = syn.features(19, n, dim, unit=True)
    st = FeatureStore.save(str(tmp_path / "s"), feats)
    pred = syn.features(20, 200, dim, unit=True).bfloat16().to(cuda_device)
    want = ops.sim_topk(pred, feats.bfloat16().to(cuda_device), k, want_keys=True)
    vals, ids, keys = stream_topk(st, pred, k, chunk_rows=2500)
    assert torch.equal(keys, want[2]) and torch.equal(ids, want[1]) and torch.equal(vals, want[0])
    # two "ranks" streaming their halves, merged: same answer
    parts = [stream_topk(st, pred, k, chunk_rows=1700, rank=r, world_size=2)[2] for r in range(2)]
    assert torch.equal(ops.topk_merge(torch.stack(parts), k)[2], want[2])
