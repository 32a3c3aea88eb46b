"""CPU: the reference's on-disk formats (dataset .npy pair, checkpoint dict) - SURVEY 8(f) N4."""
import os
import sys
import types

import numpy as np
import pytest
import torch


def test_npy_dataset_and_prefetcher(tmp_path, pkg):
    N = 37
    rng = np.random.RandomState(0)
    imgs = rng.uniform(-1, 1, (N, 3, 64, 64)).astype(np.float32)
    embs = rng.standard_normal((N, 512)).astype(np.float32)
    np.save(tmp_path / "images.npy", imgs)
    np.save(tmp_path / "emb.npy", embs)
    ds = pkg.data.NpyPairDataset(str(tmp_path / "images.npy"), str(tmp_path / "emb.npy"))
    assert len(ds) == N and ds[5][0].shape == (3, 64, 64) and torch.equal(ds[5][1], torch.from_numpy(embs[5]))
    seen = []
    for world, rank in ((1, 0), (2, 0), (2, 1)):
        pf = pkg.data.PinnedPrefetcher(ds, batch_size=4, device="cpu", shuffle=True, seed=3, rank=rank, world=world)
        got = list(pf)
        assert len(got) == len(pf) == (N // 4) // world
        for im, em in got:
            assert im.shape == (4, 3, 64, 64) and em.shape == (4, 512)
            # every row is a real (image, embedding) pair of the dataset
            for r in range(4):
                j = int(np.argmin(np.abs(embs - em[r].numpy()).sum(axis=1)))
                assert np.array_equal(embs[j], em[r].numpy()) and np.array_equal(imgs[j], im[r].numpy())
                seen.append((world, rank, j))
    w2 = [j for w, r, j in seen if w == 2]
    assert len(w2) == len(set(w2))             # the two ranks never see the same sample within an epoch
    with pytest.raises(ValueError):
        np.save(tmp_path / "bad.npy", embs[:5])
        pkg.data.NpyPairDataset(str(tmp_path / "images.npy"), str(tmp_path / "bad.npy"))


def test_checkpoint_roundtrip_reference_format(tmp_path, pkg):
    hm = pkg.host_model
    torch.manual_seed(0)
    G, D = hm.AuroraGenerator(), hm.AuroraDiscriminator()
    pkg.data.save_checkpoint(str(tmp_path / "ckpt.pth"), G, D)
    ck = torch.load(str(tmp_path / "ckpt.pth"))
    assert set(ck) == {"generator", "discriminator"}
    assert "gen_block_16.attn_block.moe.experts.3.net.2.bias" in ck["generator"]
    assert "gen_block_4.attn_block.moe.router.epsilon_c" in ck["generator"]
    assert "conv_layers.0.weight_g" in ck["discriminator"] and "conv_layers.0.weight_v" in ck["discriminator"]
    G2, D2 = hm.AuroraGenerator(), hm.AuroraDiscriminator()
    pkg.data.load_checkpoint(str(tmp_path / "ckpt.pth"), G2, D2)
    for (k, a), (_, b) in zip(G.state_dict().items(), G2.state_dict().items()):
        assert torch.equal(a, b), k
    for (k, a), (_, b) in zip(D.state_dict().items(), D2.state_dict().items()):
        assert torch.equal(a, b), k
    ref_dir = "/root/reference/moegan"
    if os.path.exists(ref_dir):               # ...and the unmodified reference loads it (build container only)
        sys.modules.setdefault("clip", types.ModuleType("clip"))
        sys.path.insert(0, ref_dir)
        import t2i_moe_gan as ref
        g_ref, d_ref = ref.AuroraGenerator(), ref.AuroraDiscriminator()
        g_ref.load_state_dict(ck["generator"])
        d_ref.load_state_dict(ck["discriminator"])
