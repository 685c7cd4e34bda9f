"""CPU checks of the rows around the hot path: the reference's CLI flags, its checkpoint file format
and pretrain transfer (utils.py:17-239), and the input pipeline (Data_Manager.py, train.py:174-357).  No kernel runs."""
import os

import pytest
import torch

from oracle import ref_port as rp


@pytest.fixture(scope="module")
def N(vcg):
    from vcg_b200 import Networks
    return Networks


def test_cli_accepts_every_reference_flag(vcg):
    """every option of the reference parser (train.py:588-656) parses, with the reference's defaults"""
    from vcg_b200 import train
    p = train.build_parser()
    a = p.parse_args([])
    ref_defaults = dict(architecture="autoencoder", paired=False, pretrained_doubleae=None, pretrained_doublevae=None,
                        data_dir="dataset", source_modality=None, target_modality=None, image_size=256, test_split=0.1,
                        dataset="hypersim", batch_size=5, epochs=100, lr=0.0002, lambda_kl=1e-5, lambda_gan=1.0,
                        lambda_identity=5.0, lambda_cycle=10.0, lambda_recon=1.0, output_dir="runs", save_freq=10,
                        log_image_freq=5, resume=None, num_workers=1, no_cuda=False)
    for k, v in ref_defaults.items():
        assert getattr(a, k) == v, (k, getattr(a, k), v)
    a = p.parse_args("--architecture cyclevaegan --unpaired --resume runs/x/best_model.pth --save_freq 2 --test_split 0.2 "
                     "--num_workers 4 --source_modality depth --target_modality normal --pretrained_doublevae a.pth "
                     "--log_image_freq 1 --dataset maps --no_cuda".split())
    assert a.resume.endswith("best_model.pth") and a.save_freq == 2 and a.num_workers == 4 and a.dataset == "maps"
    with pytest.raises(ValueError):
        a2 = p.parse_args(["--architecture", "vae", "--source_modality", "depth", "--target_modality", "normal"])
        train.main(a2)


def _ref_checkpoint(arch, path, step):
    """a checkpoint in the reference's dict format (utils.py:17-28) from the pinned oracle port + torch.optim.Adam"""
    torch.manual_seed(3)
    ora = rp.RefModel(arch, lr=2e-4)
    if step:
        ora.training_step(rp.synthetic_batch(1, same_xy=True, size=64))
    opt = {"optimizer": ora.opt_G.state_dict()} if ora.opt_D is None else \
        {"optimizer_G": ora.opt_G.state_dict(), "optimizer_D": ora.opt_D.state_dict()}
    torch.save({"epoch": 4, "model_state_dict": ora.state_dict(), "optimizer_states": opt, "loss": 1.25,
                "args": {"architecture": arch, "paired": False, "lr": 2e-4}}, path)
    return ora


def test_reference_format_checkpoint_round_trip(N, tmp_path):
    """reference-format file -> our model + FusedAdam -> our save -> torch.optim.Adam / oracle state: nothing lost"""
    from vcg_b200 import utils
    path = str(tmp_path / "ref.pth")
    ora = _ref_checkpoint("autoencoder", path, step=True)
    ours = N.Autoencoder()
    ours.configure_optimizers(lr=2e-4)
    epoch, loss = utils.load_checkpoint(ours, path, "cpu")
    assert (epoch, loss) == (4, 1.25)
    sd, ref_sd = ours.state_dict(), ora.state_dict()
    assert list(sd) == list(ref_sd)
    assert all(torch.equal(sd[k], ref_sd[k]) for k in sd)
    st = ours.optimizer.state_dict()["state"]
    ref_st = ora.opt_G.state_dict()["state"]
    assert set(st) == set(ref_st)
    for i in st:
        assert float(st[i]["step"]) == 1.0
        assert torch.equal(st[i]["exp_avg"], ref_st[i]["exp_avg"]) and torch.equal(st[i]["exp_avg_sq"], ref_st[i]["exp_avg_sq"])
    out = str(tmp_path / "ours.pth")
    utils.save_checkpoint(ours, 5, 0.75, {"architecture": "autoencoder", "paired": True}, out)
    ck = torch.load(out, map_location="cpu", weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_states", "loss", "args"} and ck["epoch"] == 5
    # the reference side: a plain torch Adam over reference-shaped parameters accepts the optimizer state
    params = [torch.nn.Parameter(v.clone()) for k, v in ck["model_state_dict"].items() if not rp.is_buffer(k)]
    ref_opt = torch.optim.Adam(params, lr=2e-4, betas=(0.5, 0.999))
    ref_opt.load_state_dict(ck["optimizer_states"]["optimizer"])
    assert torch.equal(ref_opt.state_dict()["state"][0]["exp_avg"], ref_st[0]["exp_avg"])
    # saved tensors own their storage (the live moments are views of one flat buffer)
    t = ck["optimizer_states"]["optimizer"]["state"][0]["exp_avg"]
    assert t.untyped_storage().nbytes() == t.numel() * 4
    with pytest.raises(FileNotFoundError):
        utils.load_checkpoint(ours, str(tmp_path / "missing.pth"), "cpu")


def test_two_optimizer_checkpoint_keys(N, tmp_path):
    from vcg_b200 import utils
    path = str(tmp_path / "gan.pth")
    ora = _ref_checkpoint("vaegan", path, step=False)
    ours = N.VAEGAN()
    utils.load_checkpoint(ours, path, "cpu")          # configures the optimizers itself, like the reference
    assert ours.optimizer_G is not None and ours.optimizer_D is not None
    assert set(ours.save_optimizer_states()) == {"optimizer_G", "optimizer_D"}
    assert torch.equal(ours.state_dict()["D.model.4.weight_v"], ora.state_dict()["D.model.4.weight_v"])


def test_pretrain_transfer_doublevae_to_cyclevae(N, tmp_path):
    """utils.py:124-239: G <- shared encoder + *_B parts, F <- shared encoder + *_A parts (swap asserts included)"""
    from vcg_b200 import utils
    torch.manual_seed(9)
    st = rp.init_state("doublevae")
    path = str(tmp_path / "dvae.pth")
    torch.save({"epoch": 0, "model_state_dict": st, "optimizer_states": {}, "loss": 0.0, "args": {}}, path)
    cyc = N.CycleVAE(paired=False)
    utils.load_pretrained_doublevae_to_cyclevae(cyc, path, "cpu")
    sd = cyc.state_dict()
    assert torch.equal(sd["G.decoder.model.5.conv.weight"], st["decoder_B.model.5.conv.weight"])
    assert torch.equal(sd["F.decoder.model.5.conv.weight"], st["decoder_A.model.5.conv.weight"])
    assert torch.equal(sd["G.variational_encoder_block.muConv.conv.weight"], st["vae_encoder_block_B.muConv.conv.weight"])
    assert torch.equal(sd["F.encoder.model.1.conv.weight"], st["encoder.model.1.conv.weight"])
    st2 = rp.init_state("doubleae")
    path2 = str(tmp_path / "dae.pth")
    torch.save({"epoch": 0, "model_state_dict": st2, "optimizer_states": {}, "loss": 0.0, "args": {}}, path2)
    cae = N.CycleAE(paired=True)
    utils.load_pretrained_doubleae_to_cycleae(cae, path2, "cpu")
    assert torch.equal(cae.state_dict()["G.decoder.model.0.conv1.weight"], st2["decoder_B.model.0.conv1.weight"])
    assert torch.equal(cae.state_dict()["F.decoder.model.0.conv1.weight"], st2["decoder_A.model.0.conv1.weight"])


# ------------------------------------------------------------------------------------------------ input pipeline
def _png(path, w, h, seed, mirror_halves=False):
    from PIL import Image
    g = torch.Generator().manual_seed(seed)
    img = (torch.rand(h, w, 3, generator=g) * 255).to(torch.uint8)
    if mirror_halves:
        img[:, w // 2:] = img[:, :w // 2]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    Image.fromarray(img.numpy()).save(path)


def test_datasets_follow_the_reference_layouts(vcg, tmp_path):
    from vcg_b200 import Data_Manager as DM
    root = tmp_path / "dataset"
    for i in range(5):
        _png(str(root / "maps" / "train" / f"{i}.png"), 80, 40, i, mirror_halves=True)
        _png(str(root / "maps" / "val" / f"{i}.png"), 80, 40, 10 + i)
        _png(str(root / "summer2winter" / "trainA" / f"a{i}.png"), 48, 48, 20 + i)
    for i in range(3):
        _png(str(root / "summer2winter" / "trainB" / f"b{i}.png"), 48, 48, 30 + i)
    for f in range(4):
        for m in ("depth", "normal"):
            _png(str(root / "hypersim" / "ai_001_001_kitchen" / "cam_00" / f"frame_{f:04d}_{m}.png"), 64, 48, 40 + f)
    # maps: both halves get the SAME random crop / flip (identical halves -> identical tensors)
    tf, _ = DM.build_transforms("maps", 32, True)
    maps = DM.SatelliteMapDataset(str(root / "maps"), "train", tf)
    s = maps[1]
    assert s["x"].shape == (3, 32, 32) and torch.equal(s["x"], s["y"]) and 0.0 <= float(s["x"].min()) and float(s["x"].max()) <= 1.0
    val = DM.SatelliteMapDataset(str(root / "maps"), "val", DM.build_transforms("maps", 32, False)[0])
    assert val[0]["x"].shape == (3, 32, 32) and not torch.equal(val[0]["x"], val[0]["y"])
    # summer2winter: length = the larger domain, x walks A cyclically
    s2w = DM.Summer2WinterDataset(str(root / "summer2winter"), "train", DM.build_transforms("summer2winter", 32, True)[0])
    assert len(s2w) == 5 and s2w[4]["y"].shape == (3, 32, 32)
    with pytest.raises(ValueError):
        DM.Summer2WinterDataset(str(root / "summer2winter"), "test")
    # hypersim: scene parsing, paired = same frame and same random state for both modalities
    tf, ctf = DM.build_transforms("hypersim", 32, True)
    hs = DM.HypersimDataset(str(root / "hypersim"), ["depth", "normal"], tf, ctf, paired_mode=True)
    assert len(hs) == 4 and hs.get_unique_scenes() == ["ai_001_001"] and hs.get_unique_scene_types() == ["kitchen"]
    it = hs[2]
    assert torch.equal(it["x"], it["y"]) and it["frame_id"] == "0002"        # the two modality files hold the same pixels
    one = DM.HypersimDataset(str(root / "hypersim"), ["depth"], None, None, paired_mode=True)
    assert torch.equal(one[0]["x"], one[0]["y"]) and one[0]["x"].shape == (3, 48, 64)
    assert len(hs.filter_by_scene_type(["office"]).samples) == 0
    with pytest.raises(ValueError):
        DM.HypersimDataset(str(root / "hypersim"), ["depth", "normal", "color"], paired_mode=True)


def test_data_parallel_sharding_and_loader(vcg, tmp_path):
    from vcg_b200 import Data_Manager as DM
    world, gb, n = 4, 8, 21
    per_rank = [DM.DistributedShard(n, gb, r, world, shuffle=True, seed=5).batches() for r in range(world)]
    assert all(len(b) == 3 for b in per_rank)                   # 2 full global batches + a ragged one of 5 -> 1 per rank
    ref = torch.randperm(n, generator=torch.Generator().manual_seed(5)).tolist()
    for step in range(2):                                       # rank r holds rows [r*B/W, (r+1)*B/W) of the global batch
        glob = [i for r in range(world) for i in per_rank[r][step]]
        assert glob == ref[step * gb:(step + 1) * gb]
    assert [len(b[2]) for b in per_rank] == [1, 1, 1, 1]
    assert len(DM.DistributedShard(n, gb, 0, world, drop_last=True).batches()) == 2
    with pytest.raises(ValueError):
        DM.DistributedShard(n, 6, 0, 4)
    root = tmp_path / "d"
    for i in range(6):
        _png(str(root / "maps" / "train" / f"{i}.png"), 40, 20, i)
    ds = DM.SatelliteMapDataset(str(root / "maps"), "train", DM.build_transforms("maps", 16, False)[0])
    loader = DM.DeviceLoader(ds, 4, device=None, rank=1, world=2, shuffle=False, num_workers=0, to_device=False)
    batches = list(loader)
    assert len(loader) == 2 and batches[0]["x"].shape == (2, 3, 16, 16) and batches[1]["x"].shape == (1, 3, 16, 16)
    assert torch.equal(batches[0]["x"][0], ds[2]["x"])


def test_bench_configurations_cover_baseline_json(N):
    """bench.py --config N: one entry per BASELINE.json configuration, each naming a constructible composite with the
    two-method training contract and the batch BASELINE.json quotes."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("vcg_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    base = json.load(open(os.path.join(root, "BASELINE.json")))
    assert sorted(bench.CONFIGS) == list(range(1, len(base["configs"]) + 1))
    for i, text in enumerate(base["configs"], start=1):
        cfg = bench.CONFIGS[i]
        assert f"batch {cfg['batch']}" in text, (i, text)
        model = getattr(N, cfg["cls"])(**cfg["kwargs"])
        model.configure_optimizers(lr=2e-4)
        model.configure_loss(**bench.LOSS_KW)
        assert callable(model.training_step) and callable(model.validation_step)
        assert len(bench.optimizers(model)) == (2 if hasattr(model, "optimizer_D") and model.optimizer_D is not None else 1)
    assert bench.CONFIGS[5]["metric"] == bench.METRIC and bench.CONFIGS[2]["kwargs"]["latent_dim"] == 1024


def test_inference_entry_finds_runs_and_loads_reference_checkpoints(vcg, N, tmp_path):
    """test.py of the reference (test.py:31-70, 110-142): run discovery by args.json + best_model.pth, model rebuilt from
    the checkpoint's own args in eval mode, comparison strip written (the forward pass itself needs the GPU)."""
    from vcg_b200 import test as T
    from vcg_b200 import utils as U
    import argparse
    import json
    runs = tmp_path / "runs"
    good, bad = runs / "vae_0101_0000_x_to_y_synthetic", runs / "no_checkpoint_here"
    good.mkdir(parents=True)
    bad.mkdir()
    args = argparse.Namespace(architecture="vae", paired=False, latent_dim=32, dataset="synthetic", image_size=256, lr=2e-4)
    with open(good / "args.json", "w") as f:
        json.dump(vars(args), f)
    with open(bad / "args.json", "w") as f:
        json.dump(vars(args), f)
    torch.manual_seed(5)
    src = N.VariationalAutoencoder(latent_dim=32)
    src.configure_optimizers(lr=2e-4)
    U.save_checkpoint(src, 7, 0.5, args, good / "best_model.pth")
    found = T.discover_runs(runs)
    assert [r["name"] for r in found] == [good.name] and found[0]["architecture"] == "vae"
    assert T.discover_runs(tmp_path / "missing") == []
    m = T.load_model_for_inference("vae", found[0]["checkpoint"], torch.device("cpu"))
    assert type(m).__name__ == "VariationalAutoencoder" and m.latent_dim == 32 and not m.training
    for (k, a), (_, b) in zip(sorted(src.state_dict().items()), sorted(m.state_dict().items())):
        assert torch.equal(a, b), k
    x = torch.rand(3, 3, 16, 16)
    T.save_strip(tmp_path / "strip.png", x, x * 0.5, 1.0 - x, max_samples=2)
    from PIL import Image
    assert Image.open(tmp_path / "strip.png").size == (48, 32)
    assert T.build_parser().parse_args([]).num_samples == 8
