"""Checkpoint files and pretrain transfer, in the reference's on-disk format (its utils.py:17-239).

A checkpoint is ``torch.save({'epoch', 'model_state_dict', 'optimizer_states', 'loss', 'args'})``:
``model_state_dict`` has the reference's keys / shapes / dtypes (fp32 OIHW masters, spectral-norm
``weight_orig`` / ``weight_u`` / ``weight_v``), ``optimizer_states`` is ``model.save_optimizer_states()``
(``{'optimizer'}`` or ``{'optimizer_G','optimizer_D'}`` with torch.optim.Adam state dicts).  Files
written by the reference load here and files written here load into the reference (the kernel-layout
filter copies, flat gradient / moment buffers and device step counters are derived state and never
reach the file).  Everything is written on rank 0 only when data-parallel."""
from __future__ import annotations

import os

import torch


def _cpu_state(obj):
    """detached CPU copies (own storage each: the Adam moments are views of one flat device buffer)"""
    if torch.is_tensor(obj):
        return obj.detach().to("cpu", copy=True).contiguous()
    if isinstance(obj, dict):
        return {k: _cpu_state(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu_state(v) for v in obj)
    return obj


def save_checkpoint(model, epoch, loss, args, filename):
    """utils.py:17-28 of the reference.  `args`: argparse.Namespace or dict."""
    ckpt = {
        "epoch": epoch,
        "model_state_dict": _cpu_state(model.state_dict()),
        "optimizer_states": _cpu_state(model.save_optimizer_states()),
        "loss": loss,
        "args": dict(args) if isinstance(args, dict) else vars(args),
    }
    tmp = f"{filename}.tmp"
    torch.save(ckpt, tmp)
    os.replace(tmp, filename)            # never leave a truncated checkpoint behind
    print(f"Checkpoint saved to {filename}")


def load_checkpoint(model, filename, device):
    """utils.py:31-54 of the reference: -> (epoch, loss).  The optimizers are configured with their defaults first if
    the caller has not done so."""
    if not os.path.exists(filename):
        raise FileNotFoundError(f"No checkpoint found at {filename}")
    ckpt = torch.load(filename, map_location=device, weights_only=False)
    model.load_state_dict(ckpt["model_state_dict"])
    unconfigured = all(getattr(model, n, None) is None for n in ("optimizer", "optimizer_G", "optimizer_D"))
    if unconfigured:
        try:
            model.configure_optimizers()
        except Exception:
            pass
    if "optimizer_states" in ckpt:
        model.load_optimizer_states(ckpt["optimizer_states"])
    epoch, loss = ckpt["epoch"], ckpt["loss"]
    print(f"Loaded checkpoint from {filename} (epoch {epoch}, loss {loss:.4f})")
    return epoch, loss


def _split_by_prefix(state, prefixes):
    out = {p: {} for p in prefixes}
    for key, value in state.items():
        for p in prefixes:
            if key.startswith(p + "."):
                out[p][key[len(p) + 1:]] = value
                break
    return out


def _check_same(module, source, what):
    for (name, got), (_, want) in zip(module.state_dict().items(), source.items()):
        if not torch.equal(got.cpu(), want.cpu()):
            raise AssertionError(f"{what} mismatch at {name} -- G and F may be swapped!")


def load_pretrained_doubleae_to_cycleae(cycleae_model, doubleae_checkpoint_path, device):
    """DoubleAutoencoder checkpoint -> Cycle model (utils.py:57-121): G (A->B) = shared encoder + decoder_B,
    F (B->A) = shared encoder + decoder_A."""
    if not os.path.exists(doubleae_checkpoint_path):
        raise FileNotFoundError(f"No DoubleAutoencoder checkpoint found at {doubleae_checkpoint_path}")
    print(f"Loading DoubleAutoencoder weights from {doubleae_checkpoint_path}")
    state = torch.load(doubleae_checkpoint_path, map_location=device, weights_only=False)["model_state_dict"]
    part = _split_by_prefix(state, ("encoder", "decoder_A", "decoder_B"))
    for gen, dec in ((cycleae_model.G, "decoder_B"), (cycleae_model.F, "decoder_A")):
        gen.encoder.load_state_dict(part["encoder"])
        gen.decoder.load_state_dict(part[dec])
    _check_same(cycleae_model.G.decoder, part["decoder_B"], "G.decoder")
    _check_same(cycleae_model.F.decoder, part["decoder_A"], "F.decoder")
    print("Successfully loaded DoubleAutoencoder weights: G = encoder + decoder_B, F = encoder + decoder_A")


def load_pretrained_doublevae_to_cyclevae(cycle_model, doublevae_checkpoint_path, device):
    """DoubleVariationalAutoencoder checkpoint -> CycleVAE / CycleVAEGAN (utils.py:124-239): G takes the shared
    encoder and the *_B bottleneck / decoder, F the *_A ones; the reference's swap asserts (:205-235) are kept."""
    if not os.path.exists(doublevae_checkpoint_path):
        raise FileNotFoundError(f"No DoubleVariationalAutoencoder checkpoint found at {doublevae_checkpoint_path}")
    print(f"Loading DoubleVariationalAutoencoder weights from {doublevae_checkpoint_path}")
    state = torch.load(doublevae_checkpoint_path, map_location=device, weights_only=False)["model_state_dict"]
    part = _split_by_prefix(state, ("encoder", "vae_encoder_block_A", "vae_encoder_block_B", "vae_decoder_block_A",
                                    "vae_decoder_block_B", "decoder_A", "decoder_B"))
    for gen, w in ((cycle_model.G, "B"), (cycle_model.F, "A")):
        gen.encoder.load_state_dict(part["encoder"])
        gen.variational_encoder_block.load_state_dict(part[f"vae_encoder_block_{w}"])
        gen.variational_decoder_block.load_state_dict(part[f"vae_decoder_block_{w}"])
        gen.decoder.load_state_dict(part[f"decoder_{w}"])
    print("Running sanity check on weight transfer...")
    _check_same(cycle_model.G.decoder, part["decoder_B"], "G.decoder")
    _check_same(cycle_model.F.decoder, part["decoder_A"], "F.decoder")
    _check_same(cycle_model.G.variational_decoder_block, part["vae_decoder_block_B"], "G.variational_decoder_block")
    _check_same(cycle_model.F.variational_decoder_block, part["vae_decoder_block_A"], "F.variational_decoder_block")
    print("Sanity check passed: G uses B components, F uses A components")


def truncate_tensorboard_events(tensorboard_dir, max_epoch):
    """The reference rewrites its TensorBoard event files on resume (utils.py:242-302).  Log plumbing is outside the
    accelerated path: events past `max_epoch` are left in place (TensorBoard shows the later write for a repeated
    step); returns the number of event files found."""
    if not os.path.isdir(tensorboard_dir):
        return 0
    return len([f for f in os.listdir(tensorboard_dir) if f.startswith("events.out.tfevents")])
