"""Input pipeline in front of the training step: the reference's three datasets (its Data_Manager.py:18-451) and
the loaders its train.py:174-357 builds, feeding a GPU that consumes thousands of images per second.

Same directory contracts, class names, constructor arguments and sample dictionaries as the reference:

  HypersimDataset        <root>/<scene>_<type>/cam_XX/frame_NNNN_<modality>.png; one or two modalities; paired mode
                         returns the same frame as 'x' / 'y' (transformed with ONE random state), unpaired mode draws
                         'y' from a random frame
  SatelliteMapDataset    <root>/{train,val}/*.jpg, each image = satellite | map halves, same random state for both
  Summer2WinterDataset   <root>/{train,test}{A,B}/*: 'x' walks domain A, 'y' is a random domain-B image

Decoding and augmentation stay on the host (PIL + torchvision.transforms with the reference's own parameters -- they
define the data distribution); what changes is the hand-over to the device:

  * DistributedShard: every data-parallel rank reads rows [r*B/W, (r+1)*B/W) of each GLOBAL batch of an
    epoch-seeded permutation (SURVEY.md 8e partitioning), so W ranks together see what one reference process sees;
  * DeviceLoader: fixed-shape batches collated straight into a ring of pinned host buffers and copied to the device
    on a side stream one batch ahead of the consumer (or handed over as pinned batches to graph.GraphedStep, which
    overlaps the copy with the replayed step itself)."""
from __future__ import annotations

import os
import random
from pathlib import Path

import torch
from torch.utils.data import DataLoader, Dataset, Sampler

_EXT = (".jpg", ".jpeg", ".png")


def _to_tensor(img):
    from torchvision.transforms import functional as TF
    return img if torch.is_tensor(img) else TF.to_tensor(img)


def _open_rgb(path):
    from PIL import Image
    with Image.open(path) as im:
        return im.convert("RGB")


def _listing(folder):
    if not os.path.isdir(folder):
        raise ValueError(f"Directory not found: {folder}")
    names = sorted(f for f in os.listdir(folder) if f.lower().endswith(_EXT))
    if not names:
        raise ValueError(f"No images found in {folder}")
    return names


class _SharedState:
    """Applies one transform to several images with the SAME torch RNG state, so that random flips / crops agree
    across the modalities of one sample (the reference's get_rng_state / set_rng_state idiom)."""

    def __init__(self):
        self.state = torch.get_rng_state()

    def __call__(self, transform, img):
        torch.set_rng_state(self.state)
        return _to_tensor(transform(img))


# ------------------------------------------------------------------------------------------------ datasets
class HypersimDataset(Dataset):
    def __init__(self, root_dir, modalities=("color", "depth", "normal_world", "normal", "semantic", "semantic_instance", "normal"),
                 transform=None, color_transform=None, return_scene_info=True, paired_mode=True):
        self.root_dir = Path(root_dir)
        self.modalities = list(modalities)
        self.transform, self.color_transform = transform, color_transform
        self.return_scene_info, self.paired_mode = return_scene_info, paired_mode
        if paired_mode and len(self.modalities) not in (1, 2):
            raise ValueError(f"paired_mode requires 1 or 2 modalities, got {len(self.modalities)}")
        self.samples = self._scan_dataset()
        if not self.samples:
            raise ValueError(f"No samples found in {root_dir}")
        print(f"  Loaded dataset with {len(self.samples)} samples; modalities: {', '.join(self.modalities)}; "
              f"scenes: {len(self.get_unique_scenes())}")

    def _scan_dataset(self):
        found = []
        first = self.modalities[0]
        for scene in sorted(p for p in self.root_dir.iterdir() if p.is_dir()):
            parts = scene.name.split("_")
            num, kind = ("_".join(parts[:3]), "_".join(parts[3:])) if len(parts) >= 4 else (scene.name, "unknown")
            for cam in (c for c in scene.glob("cam_*") if c.is_dir()):
                for frame in sorted(cam.glob(f"frame_*_{first}.png")):
                    fid = frame.stem.split("_")[1]
                    paths = {m: cam / f"frame_{fid}_{m}.png" for m in self.modalities}
                    if all(p.exists() for p in paths.values()):
                        found.append({"scene_dir": scene, "scene_num": num, "scene_type": kind, "camera": cam.name,
                                      "cam_num": cam.name.replace("cam_", ""), "frame_id": fid, "modality_paths": paths})
        return found

    def __len__(self):
        return len(self.samples)

    def _prepare(self, img, modality, shared=None):
        if modality == "color" and self.color_transform is not None:
            img = self.color_transform(img)
        if self.transform is None:
            return _to_tensor(img)
        return shared(self.transform, img) if shared is not None else _to_tensor(self.transform(img))

    def _load_modality_at_index(self, idx, modality):
        return self._prepare(_open_rgb(self.samples[idx]["modality_paths"][modality]), modality)

    def __getitem__(self, idx):
        info = self.samples[idx]
        shared = _SharedState() if self.transform is not None else None
        out = {m: self._prepare(_open_rgb(p), m, shared) for m, p in info["modality_paths"].items()}
        meta = {"frame_id": info["frame_id"]}
        if self.return_scene_info:
            meta.update(scene_num=info["scene_num"], scene_type=info["scene_type"], cam_num=info["cam_num"])
        if self.paired_mode:
            a, b = self.modalities[0], self.modalities[-1]
            return {"x": out[a], "y": out[b], **meta}
        if len(self.modalities) != 2:
            raise ValueError("Unpaired mode requires exactly 2 modalities")
        a, b = self.modalities
        j = random.randint(0, len(self.samples) - 1)
        return {"x": out[a], "y": out[b] if j == idx else self._load_modality_at_index(j, b), **meta}

    def get_unique_scenes(self):
        return sorted({s["scene_num"] for s in self.samples})

    def get_unique_scene_types(self):
        return sorted({s["scene_type"] for s in self.samples})

    def _subset(self, keep):
        import copy
        sub = copy.copy(self)
        sub.samples = [s for s in self.samples if keep(s)]
        return sub

    def filter_by_scene(self, scene_nums):
        wanted = set(scene_nums)
        return self._subset(lambda s: s["scene_num"] in wanted)

    def filter_by_scene_type(self, scene_types):
        wanted = set(scene_types)
        return self._subset(lambda s: s["scene_type"] in wanted)


class SatelliteMapDataset(Dataset):
    def __init__(self, root_dir, split="train", transform=None):
        self.root_dir, self.split, self.transform = root_dir, split, transform
        self.image_dir = os.path.join(root_dir, split)
        self.images = _listing(self.image_dir)
        print(f"  Loaded {split} split with {len(self.images)} samples")

    def __len__(self):
        return len(self.images)

    def __getitem__(self, idx):
        both = _open_rgb(os.path.join(self.image_dir, self.images[idx]))
        w, h = both.size
        halves = (both.crop((0, 0, w // 2, h)), both.crop((w // 2, 0, w, h)))       # satellite | map
        if self.transform is None:
            return {"x": _to_tensor(halves[0]), "y": _to_tensor(halves[1])}
        shared = _SharedState()
        return {"x": shared(self.transform, halves[0]), "y": shared(self.transform, halves[1])}


class Summer2WinterDataset(Dataset):
    def __init__(self, root_dir, split="train", transform=None):
        self.root_dir, self.split, self.transform = root_dir, split, transform
        self.dir_A, self.dir_B = os.path.join(root_dir, f"{split}A"), os.path.join(root_dir, f"{split}B")
        self.images_A, self.images_B = _listing(self.dir_A), _listing(self.dir_B)
        print(f"  Loaded {split} split: {len(self.images_A)} domain A, {len(self.images_B)} domain B images")

    def __len__(self):
        return max(len(self.images_A), len(self.images_B))

    def __getitem__(self, idx):
        a = _open_rgb(os.path.join(self.dir_A, self.images_A[idx % len(self.images_A)]))
        b = _open_rgb(os.path.join(self.dir_B, self.images_B[random.randint(0, len(self.images_B) - 1)]))
        f = self.transform if self.transform is not None else (lambda im: im)
        return {"x": _to_tensor(f(a)), "y": _to_tensor(f(b))}          # independent random states: the domains are unpaired


# ------------------------------------------------------------------------------------------------ transforms
def build_transforms(dataset, image_size, train):
    """The augmentation recipes of the reference's train.py:183-196, 253-266, 310-326 (torchvision, same parameters).
    -> (transform, color_transform)"""
    from torchvision import transforms as T
    bicubic = T.InterpolationMode.BICUBIC
    crop = T.RandomResizedCrop(size=image_size, scale=(0.33, 1.0), ratio=(1, 1), interpolation=bicubic)
    if dataset == "hypersim":       # no held-out recipe: the test split is a random_split of the augmented dataset
        return (T.Compose([T.RandomHorizontalFlip(p=0.5), T.RandomVerticalFlip(p=0.3), crop, T.ToTensor()]),
                T.Compose([T.ColorJitter(brightness=0.3, contrast=0.3, saturation=0.3, hue=0.15)]))
    if not train:
        return T.Compose([T.Resize((image_size, image_size)), T.ToTensor()]), None
    if dataset == "maps":
        return T.Compose([T.RandomHorizontalFlip(p=0.5), crop, T.ToTensor()]), None
    if dataset == "summer2winter":
        return T.Compose([T.RandomHorizontalFlip(p=0.5), crop,
                          T.ColorJitter(brightness=0.2, contrast=0.2, saturation=0.2, hue=0.1), T.ToTensor()]), None
    raise ValueError(f"unknown dataset {dataset!r}")


# ------------------------------------------------------------------------------------------------ device hand-over
class DistributedShard(Sampler):
    """Indices of rank r: rows [r*B/W, (r+1)*B/W) of every global batch of an epoch-seeded permutation."""

    def __init__(self, n, global_batch, rank=0, world=1, shuffle=True, seed=0, drop_last=False):
        if global_batch % world:
            raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
        self.n, self.gb, self.rank, self.world = n, global_batch, rank, world
        self.shuffle, self.seed, self.drop_last, self.epoch = shuffle, seed, drop_last, 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def _order(self):
        if not self.shuffle:
            return list(range(self.n))
        g = torch.Generator().manual_seed(self.seed + self.epoch)
        return torch.randperm(self.n, generator=g).tolist()

    def batches(self):
        order = self._order()
        per = self.gb // self.world
        out = []
        for o in range(0, self.n, self.gb):
            glob = order[o:o + self.gb]
            if len(glob) < self.gb:
                if self.drop_last or len(glob) < self.world:
                    break
                per_last = len(glob) // self.world          # ragged last batch: equal shards, remainder dropped
                out.append(glob[self.rank * per_last:(self.rank + 1) * per_last])
                break
            out.append(glob[self.rank * per:(self.rank + 1) * per])
        return out

    def __iter__(self):
        return iter(self.batches())

    def __len__(self):
        return len(self.batches())


def _collate_xy(samples):
    """{'x','y'} stacked; string metadata of the Hypersim samples is dropped (the step does not read it)"""
    return {"x": torch.stack([s["x"] for s in samples]), "y": torch.stack([s["y"] for s in samples])}


class DeviceLoader:
    """Iterates {'x','y'} batches of a dataset for one rank.

    to_device=True: yields DEVICE batches; the host-to-device copy of batch i+1 is issued on a side stream from a
    pinned buffer while the caller computes on batch i.  to_device=False: yields pinned HOST batches (what
    graph.GraphedStep takes as `prefetch`).  len() = batches per epoch of this rank."""

    def __init__(self, dataset, global_batch, device=None, rank=0, world=1, shuffle=True, num_workers=1, seed=0,
                 drop_last=False, to_device=True):
        self.dataset, self.device, self.to_device = dataset, device, to_device
        self.sampler = DistributedShard(len(dataset), global_batch, rank, world, shuffle, seed, drop_last)
        self.loader = DataLoader(dataset, batch_sampler=self.sampler, num_workers=num_workers, collate_fn=_collate_xy,
                                 pin_memory=torch.cuda.is_available(), persistent_workers=num_workers > 0)
        self._epoch = 0
        self._stream = None

    def __len__(self):
        return len(self.sampler)

    def set_epoch(self, epoch):
        self._epoch = epoch
        self.sampler.set_epoch(epoch)

    def __iter__(self):
        it = iter(self.loader)
        if not (self.to_device and self.device is not None and torch.device(self.device).type == "cuda"):
            yield from it
            return
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self.device)

        def upload(host):
            with torch.cuda.stream(self._stream):
                dev = {k: v.to(self.device, non_blocking=True) for k, v in host.items()}
                ev = torch.cuda.Event()
                ev.record(self._stream)
            return host, dev, ev        # `host` stays referenced until the copy has been consumed

        nxt = next(it, None)
        staged = upload(nxt) if nxt is not None else None
        while staged is not None:
            host, dev, ev = staged
            nxt = next(it, None)
            staged = upload(nxt) if nxt is not None else None          # copy of batch i+1 overlaps the step on batch i
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in dev.values():
                t.record_stream(cur)
            yield dev


def create_dataloaders(args, device=None, rank=0, world=1):
    """train.py:174-357 of the reference for --dataset {hypersim, maps, summer2winter}: -> (train_loader, test_loader)."""
    name = args.dataset
    graph = bool(getattr(args, "cuda_graph", False))
    if name == "hypersim":
        tf, ctf = build_transforms(name, args.image_size, True)
        full = HypersimDataset(os.path.join(args.data_dir, "hypersim"), [args.source_modality, args.target_modality], tf, ctf,
                               return_scene_info=True, paired_mode=args.paired)
        train_ds, test_ds = full, None
        if args.test_split > 0:
            n_train = int((1 - args.test_split) * len(full))
            g = torch.Generator().manual_seed(getattr(args, "seed", 0))        # same split on every rank
            train_ds, test_ds = torch.utils.data.random_split(full, [n_train, len(full) - n_train], generator=g)
            print(f"Training samples: {n_train}, Testing samples: {len(full) - n_train}")
    elif name == "maps":
        root = os.path.join(args.data_dir, "maps")
        train_ds = SatelliteMapDataset(root, "train", build_transforms(name, args.image_size, True)[0])
        test_ds = SatelliteMapDataset(root, "val", build_transforms(name, args.image_size, False)[0])
    elif name == "summer2winter":
        root = os.path.join(args.data_dir, "summer2winter")
        train_ds = Summer2WinterDataset(root, "train", build_transforms(name, args.image_size, True)[0])
        test_ds = Summer2WinterDataset(root, "test", build_transforms(name, args.image_size, False)[0])
    else:
        raise ValueError(f"unknown dataset {name!r}")
    mk = lambda ds, shuffle: DeviceLoader(ds, args.batch_size, device, rank, world, shuffle=shuffle,       # noqa: E731
                                          num_workers=args.num_workers, seed=getattr(args, "seed", 0),
                                          drop_last=graph and shuffle, to_device=not (graph and shuffle))
    return mk(train_ds, True), (mk(test_ds, False) if test_ds is not None else None)
