"""Command line of the harness: ``python -m constant_memory_waveglow_b200.cli {train,synth} ...``.

The reference's own ``train.py`` runs unchanged on the ``pytorch_lightning`` / ``model`` / ``datasets`` shims at the
repository root; its ``inference.py`` cannot in this image (``torchaudio.load`` / ``save`` need the absent TorchCodec),
so the two entry points exist here with the reference's flags (``train.py:81-93``, ``inference.py:62-70``) on top of the
RIFF reader/writer of ``datasets.py``.  Multi-GPU: launch ``train`` under ``torchrun --nproc-per-node N``.
"""
from __future__ import annotations

import argparse
import json
import math
import sys

import torch

from . import trainer as TR
from .datasets import wav_read, wav_write
from .utils import remove_weight_norms


def _timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    b.synchronize()
    return out, a.elapsed_time(b) / 1e3


def train(argv) -> None:
    p = argparse.ArgumentParser(prog="cli train", description="constant-memory WaveGlow training (train.py flags)")
    p = TR.Trainer.add_argparse_args(TR.LightModel.add_model_specific_args(p))
    p.add_argument("--config", type=str)
    p.add_argument("--ckpt-path", type=str)
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--lr", type=float, default=None, help="force learning rate")
    p.add_argument("--no-tf32", action="store_true", help="fp32 operands on the exact engine (train.py:92-97)")
    args = p.parse_args(argv)
    if args.no_tf32:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    config = json.load(open(args.config)) if args.config else None
    TR.seed_everything(args.seed)
    world = max(1, int(__import__("os").environ.get("WORLD_SIZE", "1")))
    if config is not None:
        config["data_loader"]["batch_size"] //= world          # train.py:51-53: the configured batch is global
    if args.ckpt_path:
        lit = TR.LightModel.load_from_checkpoint(args.ckpt_path, **({"config": config} if config is not None else {}))
    else:
        if config is None:
            p.error("--config or --ckpt-path is required")
        lit = TR.LightModel(config)

    class _ForceLR(TR.Callback):
        def on_train_start(self, trainer, pl_module):
            for o in trainer.optimizers:
                for g in o.param_groups:
                    g["lr"] = args.lr

    cbs = [TR.ModelSummary(max_depth=2), TR.LearningRateMonitor("epoch")] + ([_ForceLR()] if args.lr else [])
    if args.max_epochs is None and args.max_steps in (-1, None):
        args.max_epochs = 100                                   # train.py:76
    tr = TR.Trainer.from_argparse_args(args, callbacks=cbs, detect_anomaly=True)
    tr.fit(lit, ckpt_path=args.ckpt_path)


def synth(argv) -> None:
    p = argparse.ArgumentParser(prog="cli synth", description="analysis + synthesis of one file (inference.py flags)")
    p.add_argument("ckpt", type=str)
    p.add_argument("infile", type=str)
    p.add_argument("outfile", type=str)
    p.add_argument("-s", "--sigma", type=float, default=0.6)
    p.add_argument("-n", "--n-group", type=int, default=None)
    args = p.parse_args(argv)
    lit = TR.LightModel.load_from_checkpoint(args.ckpt, map_location="cpu")
    model, conditioner = lit.model, lit.conditioner
    model.apply(remove_weight_norms)
    dev = torch.device("cuda")
    model, conditioner = model.to(dev).eval(), conditioner.to(dev)
    from .datasets import wav_info
    sr = wav_info(args.infile).sample_rate
    y = wav_read(args.infile).mean(0, keepdim=True).to(dev)
    if args.n_group and y.shape[1] % args.n_group:
        y = y[:, :-(y.shape[1] % args.n_group)]
    cond = conditioner(y)
    with torch.no_grad():
        (z, logdet), cost = _timed(lambda: model(y.clone(), cond))
        z = z.squeeze()
        print(z.mean().item(), z.std().item())
        print("Forward LL:", logdet.mean().item() / z.size(0) - 0.5 * (
            z.pow(2).mean().item() / args.sigma ** 2 + math.log(2 * math.pi) + 2 * math.log(args.sigma)))
        print("Time cost: {:.4f}, Speed: {:.4f} kHz".format(cost, z.numel() / cost / 1000))
        x, cost = _timed(lambda: model.infer(cond, args.sigma))
    print("Time cost: {:.4f}, Speed: {:.4f} kHz".format(cost, x.numel() / cost / 1000))
    print(x.max().item(), x.min().item())
    wav_write(args.outfile, x.reshape(1, -1), sr)


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("train", "synth"):
        raise SystemExit("usage: python -m constant_memory_waveglow_b200.cli {train,synth} [flags]  (-h after the verb)")
    (train if argv[0] == "train" else synth)(argv[1:])


if __name__ == "__main__":
    main()
