"""Lightning-free training harness: what ``train.py`` / ``inference.py`` need from
``pytorch_lightning`` (absent from this image and its wheelhouse) and the reference's ``LightModel``.

``LightModel(config)`` mirrors ``model/lightning.py:16-68``: hyper-parameters from the config dict, the flow, the
conditioner and the criterion built by reflection (``get_instance``), ``configure_optimizers``, ``train_dataloader``,
``training_step`` (conditioner -> flow -> NLL, the four logged scalars) and ``forward = model.infer``.  Checkpoints
use Lightning's dictionary layout (``state_dict`` with ``model.`` / ``conditioner.`` prefixes, ``hyper_parameters``,
``optimizer_states``, ``epoch``, ``global_step``), so ``LightModel.load_from_checkpoint`` reads files written by either.

``Trainer`` is the part of ``pl.Trainer`` that ``train.py:73-78`` drives: ``add_argparse_args`` /
``from_argparse_args``, ``fit(model, ckpt_path=)``, callbacks (``on_train_start``, ``on_train_epoch_end``),
``log`` / ``log_dict`` with ``sync_dist``, per-epoch checkpoints.  Data parallelism is one process per GPU under
``torchrun`` (the reference lets Lightning's DDPPlugin spawn them, ``train.py:51-53,77``): the sampler shards the
dataset, ``parallel.FlowGradSync`` all-reduces each flow's gradients over NCCL while the reversible backward goes on,
and batches reach the device through ``datasets.DevicePrefetcher``.
"""
from __future__ import annotations

import argparse
import csv
import math
import os
import random
import time
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist
from torch import nn
from torch.utils.data import DataLoader
from torch.utils.data.distributed import DistributedSampler

from . import condition as module_condition
from . import datasets as module_data
from . import loss as module_loss
from .datasets import DevicePrefetcher
from .parallel import FlowGradSync, allreduce_scalars, flow_buckets
from .utils import get_instance

__all__ = ["LightModel", "LightningModule", "Trainer", "Callback", "seed_everything", "ModelSummary",
           "LearningRateMonitor", "DeviceStatsMonitor", "DDPPlugin", "AttributeDict"]


def seed_everything(seed: Optional[int] = None, workers: bool = False) -> int:
    """``pl.seed_everything`` (``train.py:49``): seeds python, numpy and torch; picks a seed when given None."""
    if seed is None:
        seed = int(os.environ.get("PL_GLOBAL_SEED", random.SystemRandom().randint(0, 2 ** 32 - 1)))
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    return seed


class AttributeDict(dict):
    """``self.hparams``: a dict whose keys read as attributes (``self.hparams.arch``, ``model/lightning.py:33``)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class Callback:
    def on_train_start(self, trainer, pl_module) -> None: ...
    def on_train_epoch_start(self, trainer, pl_module) -> None: ...
    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx) -> None: ...
    def on_train_epoch_end(self, trainer, pl_module) -> None: ...
    def on_train_end(self, trainer, pl_module) -> None: ...


class ModelSummary(Callback):
    """Prints the parameter count per sub-module down to ``max_depth`` at the start of training (rank 0)."""

    def __init__(self, max_depth: int = 1):
        self.max_depth = max_depth

    def on_train_start(self, trainer, pl_module) -> None:
        if not trainer.is_global_zero:
            return
        for name, m in pl_module.named_modules():
            if name and name.count(".") < self.max_depth:
                n = sum(p.numel() for p in m.parameters())
                print(f"  {name:<32s} {type(m).__name__:<24s} {n / 1e6:9.3f} M")
        print(f"  total trainable parameters: {sum(p.numel() for p in pl_module.parameters() if p.requires_grad) / 1e6:.3f} M")


class LearningRateMonitor(Callback):
    def __init__(self, logging_interval: Optional[str] = None):
        self.logging_interval = logging_interval

    def on_train_epoch_start(self, trainer, pl_module) -> None:
        for i, opt in enumerate(trainer.optimizers):
            trainer.logged_metrics[f"lr-{type(opt).__name__}" + (f"-{i}" if i else "")] = opt.param_groups[0]["lr"]


class DeviceStatsMonitor(Callback):
    def on_train_batch_end(self, trainer, pl_module, outputs, batch, batch_idx) -> None:
        if pl_module.device.type == "cuda":
            trainer.logged_metrics["max_memory_allocated_mb"] = torch.cuda.max_memory_allocated(pl_module.device) / 2 ** 20


class DDPPlugin:
    """Accepted for signature compatibility (``train.py:77``); the process layout comes from torchrun's environment."""

    def __init__(self, find_unused_parameters: bool = False, **kwargs):
        self.find_unused_parameters = find_unused_parameters


class LightningModule(nn.Module):
    """The slice of ``pl.LightningModule`` the reference touches: hparams, ``log``/``log_dict``, ``device``,
    ``load_from_checkpoint``."""

    def __init__(self):
        super().__init__()
        self._hparams = AttributeDict()
        self.trainer: Optional["Trainer"] = None

    @property
    def hparams(self) -> AttributeDict:
        return self._hparams

    def save_hyperparameters(self, *args) -> None:
        for a in args:
            if a is None:
                continue
            if isinstance(a, argparse.Namespace):
                a = vars(a)
            if not isinstance(a, dict):
                raise TypeError("save_hyperparameters expects dicts / namespaces")
            self._hparams.update(a)

    @property
    def device(self) -> torch.device:
        for t in list(self.parameters()) + list(self.buffers()):
            return t.device
        return torch.device("cpu")

    def log(self, name: str, value, prog_bar: bool = False, sync_dist: bool = False, **kw) -> None:
        if self.trainer is not None:
            self.trainer._log(name, value, sync_dist)

    def log_dict(self, values: Dict[str, Any], prog_bar: bool = False, sync_dist: bool = False, **kw) -> None:
        for k, v in values.items():
            self.log(k, v, prog_bar=prog_bar, sync_dist=sync_dist)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict: bool = True, **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        module = cls(**hp)
        module.load_state_dict(ckpt["state_dict"], strict=strict)
        return module

    def checkpoint(self, epoch: int = 0, global_step: int = 0, optimizers=()) -> Dict[str, Any]:
        return {
            "epoch": epoch,
            "global_step": global_step,
            "pytorch-lightning_version": "0+cmwg_b200",
            "state_dict": self.state_dict(),
            "optimizer_states": [o.state_dict() for o in optimizers],
            "lr_schedulers": [],
            "hyper_parameters": dict(self.hparams),
        }


class LightModel(LightningModule):
    """Reference ``model/lightning.py:16-68``."""

    @staticmethod
    def add_model_specific_args(parent_parser: argparse.ArgumentParser):
        parent_parser.add_argument_group("Lightning")
        return parent_parser

    def __init__(self, config: dict = None, **kwargs) -> None:
        super().__init__()
        self.save_hyperparameters(config)
        self.save_hyperparameters(kwargs)
        import constant_memory_waveglow_b200 as module_arch   # reference: `import model as module_arch`
        self.model = get_instance(module_arch, self.hparams.arch)
        self.conditioner = get_instance(module_condition, self.hparams.conditioner)
        self.criterion = get_instance(module_loss, self.hparams.loss)

    def configure_optimizers(self):
        cfg = self.hparams.optimizer
        extra = {}
        if cfg["type"] in ("Adam", "AdamW", "SGD") and self.device.type == "cuda" and "fused" not in cfg["args"]:
            extra["fused"] = True   # one multi-tensor launch instead of a python loop over 450 parameters
        return getattr(torch.optim, cfg["type"])(self.parameters(), **cfg["args"], **extra)

    def train_dataloader(self):
        train_data = get_instance(module_data, self.hparams.dataset)
        return DataLoader(train_data, **self.hparams.data_loader)

    def training_step(self, batch, batch_idx):
        x = batch
        cond = self.conditioner(x)
        z, logdet = self.model(x, cond)
        loss = self.criterion(z, logdet)
        with torch.no_grad():
            values = {"logdet": logdet.sum() / z.numel(), "z_mean": z.mean(), "z_std": z.std()}
        self.log_dict(values, prog_bar=True, sync_dist=True)
        self.log("loss", loss.detach(), prog_bar=False, sync_dist=True)
        return loss

    def forward(self, *args, **kwargs):
        return self.model.infer(*args, **kwargs)


class _Experiment:
    """Stand-in for the TensorBoard writer behind ``trainer.logger.experiment`` (``train.py:32-33``): audio goes to
    wave files in the log directory, scalars to ``metrics.csv``."""

    def __init__(self, log_dir: str):
        self.log_dir = log_dir

    def add_audio(self, tag: str, snd_tensor, global_step: int = 0, sample_rate: int = 22050) -> None:
        from .datasets import wav_write
        wav_write(os.path.join(self.log_dir, f"{tag}_step{global_step}.wav"), snd_tensor.reshape(1, -1), sample_rate)

    def add_scalar(self, tag: str, value, global_step: int = 0) -> None:
        with open(os.path.join(self.log_dir, "scalars.csv"), "a", newline="") as f:
            csv.writer(f).writerow([global_step, tag, float(value)])


class _Logger:
    def __init__(self, root: str):
        base = os.path.join(root, "lightning_logs")
        os.makedirs(base, exist_ok=True)
        v = 0
        while os.path.exists(os.path.join(base, f"version_{v}")):
            v += 1
        self.log_dir = os.path.join(base, f"version_{v}")
        os.makedirs(os.path.join(self.log_dir, "checkpoints"), exist_ok=True)
        self.experiment = _Experiment(self.log_dir)
        self._fields: Optional[List[str]] = None

    def log_metrics(self, metrics: Dict[str, float], step: int) -> None:
        row = {"step": step, **metrics}
        path = os.path.join(self.log_dir, "metrics.csv")
        if self._fields is None or any(k not in self._fields for k in row):
            self._fields = list(row) if self._fields is None else self._fields + [k for k in row if k not in self._fields]
            rows = []
            if os.path.exists(path):
                with open(path, newline="") as f:
                    rows = list(csv.DictReader(f))
            with open(path, "w", newline="") as f:
                w = csv.DictWriter(f, self._fields)
                w.writeheader()
                w.writerows(rows)
        with open(path, "a", newline="") as f:
            csv.DictWriter(f, self._fields).writerow(row)


_TRAINER_FLAGS = (
    ("max_epochs", int, None), ("max_steps", int, -1), ("limit_train_batches", float, 1.0),
    ("default_root_dir", str, None), ("gradient_clip_val", float, None), ("log_every_n_steps", int, 50),
    ("precision", str, "32"), ("gpus", int, None), ("devices", int, None), ("accelerator", str, None),
    ("num_nodes", int, 1), ("accumulate_grad_batches", int, 1), ("enable_checkpointing", int, 1),
    ("fast_dev_run", int, 0), ("resume_from_checkpoint", str, None),
)


class Trainer:
    """The part of ``pytorch_lightning.Trainer`` that ``train.py`` uses."""

    def __init__(self, callbacks: Optional[List[Callback]] = None, max_epochs: Optional[int] = None, max_steps: int = -1,
                 limit_train_batches: float = 1.0, default_root_dir: Optional[str] = None,
                 gradient_clip_val: Optional[float] = None, log_every_n_steps: int = 50, precision="32",
                 gpus: Optional[int] = None, devices: Optional[int] = None, accelerator: Optional[str] = None,
                 num_nodes: int = 1, accumulate_grad_batches: int = 1, enable_checkpointing=True, fast_dev_run=0,
                 resume_from_checkpoint: Optional[str] = None, benchmark: bool = False, detect_anomaly: bool = False,
                 strategy=None, logger=True, **unused):
        if accumulate_grad_batches != 1:
            raise NotImplementedError("accumulate_grad_batches != 1 (the per-flow gradient buckets are rewritten every step)")
        self.callbacks = list(callbacks or [])
        self.max_epochs = max_epochs if max_epochs is not None else (1000 if max_steps in (-1, None) else None)
        self.max_steps = -1 if max_steps is None else max_steps
        self.limit_train_batches = limit_train_batches
        self.default_root_dir = default_root_dir or os.getcwd()
        self.gradient_clip_val = gradient_clip_val
        self.log_every_n_steps = max(1, int(log_every_n_steps))
        self.precision = str(precision)
        self.enable_checkpointing = bool(enable_checkpointing)
        self.detect_anomaly = detect_anomaly   # honoured as a finiteness check of every logged loss
        self.resume_from_checkpoint = resume_from_checkpoint
        if fast_dev_run:
            self.max_steps, self.max_epochs, self.enable_checkpointing = int(fast_dev_run), 1, False
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.global_rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        want = max(int(g) if isinstance(g, int) or (isinstance(g, str) and g.isdigit()) else 0 for g in (gpus, devices))
        if want > 1 and self.world_size == 1:
            # train.py:51-53,73-78 asks Lightning to spawn one process per GPU; this harness is launched by torchrun
            # instead, and train.py has already divided the batch by the GPU count
            import warnings
            warnings.warn(f"Trainer(gpus/devices={want}) but WORLD_SIZE=1: this process trains on ONE GPU with 1/{want} of the "
                          f"configured global batch.  Launch with `torchrun --nproc-per-node {want} --master-addr 127.0.0.1 "
                          f"train.py ...` (or set CUDA_VISIBLE_DEVICES to a single GPU to silence this).", RuntimeWarning)
        self.current_epoch = 0
        self.global_step = 0
        self.optimizers: List[torch.optim.Optimizer] = []
        self.logged_metrics: Dict[str, Any] = {}
        self._sync_keys: set = set()
        self.logger = _Logger(self.default_root_dir) if (logger and self.global_rank == 0) else None
        self.last_checkpoint: Optional[str] = None

    # ---- argparse plumbing (train.py:82-83, 73) ----
    @staticmethod
    def add_argparse_args(parent_parser: argparse.ArgumentParser):
        g = parent_parser.add_argument_group("pl.Trainer")
        for name, typ, default in _TRAINER_FLAGS:
            g.add_argument(f"--{name}", type=typ, default=default)
        return parent_parser

    @classmethod
    def from_argparse_args(cls, args, **kwargs):
        params = {name: getattr(args, name) for name, _, _ in _TRAINER_FLAGS if hasattr(args, name)}
        params.update(kwargs)
        return cls(**params)

    @property
    def is_global_zero(self) -> bool:
        return self.global_rank == 0

    def _log(self, name: str, value, sync_dist: bool) -> None:
        self.logged_metrics[name] = value
        if sync_dist:
            self._sync_keys.add(name)

    def _flush_metrics(self, device) -> Dict[str, float]:
        keys = sorted(self.logged_metrics)
        vals = [float(self.logged_metrics[k]) for k in keys]          # the only host read-back of the step
        sync = [i for i, k in enumerate(keys) if k in self._sync_keys]
        if self.world_size > 1 and sync:
            red = allreduce_scalars([vals[i] for i in sync], device)     # metric sync of model/lightning.py:58-64
            for i, v in zip(sync, red):
                vals[i] = v
        return dict(zip(keys, vals))

    def save_checkpoint(self, path: str, module: LightningModule) -> None:
        if self.is_global_zero:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            torch.save(module.checkpoint(self.current_epoch, self.global_step, self.optimizers), path)
            self.last_checkpoint = path

    def fit(self, model: LightningModule, ckpt_path: Optional[str] = None) -> None:
        if torch.cuda.is_available():
            torch.cuda.set_device(self.local_rank)
            device = torch.device("cuda", self.local_rank)
        else:
            device = torch.device("cpu")   # the flow kernels raise on CPU tensors: there is no CPU fallback
        if self.world_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            dist.init_process_group("nccl" if device.type == "cuda" else "gloo",
                                    **({"device_id": device} if device.type == "cuda" else {}))
        if self.precision in ("16", "16-mixed", "bf16", "bf16-mixed"):
            from . import precision as _p
            _p.set_precision("fp16" if str(self.precision).startswith("16") else "bf16")
        model.trainer = self
        model.to(device).train()
        opt = model.configure_optimizers()
        self.optimizers = [opt] if isinstance(opt, torch.optim.Optimizer) else list(opt)
        ckpt_path = ckpt_path or self.resume_from_checkpoint
        if ckpt_path:
            ck = torch.load(ckpt_path, map_location=device, weights_only=False)
            model.load_state_dict(ck["state_dict"])
            for o, sd in zip(self.optimizers, ck.get("optimizer_states", [])):
                o.load_state_dict(sd)
            self.current_epoch = int(ck.get("epoch", -1)) + 1
            self.global_step = int(ck.get("global_step", 0))
        flow = getattr(model, "model", model)
        trainable = [p for p in model.parameters() if p.requires_grad]
        buckets = flow_buckets(flow)
        covered = {id(p) for b in buckets for p in b}
        rest = [p for p in trainable if id(p) not in covered]
        if rest:
            buckets.append(rest)
        sync = FlowGradSync(buckets)
        # CMWG_TRAIN_GRAPH=1: the whole step (forward, loss, reversible backward, all-reduces, optimizer) replayed as one CUDA
        # graph per batch shape (graphs.py); needs capturable optimizers and a training_step without host round trips
        graphed = None
        if os.environ.get("CMWG_TRAIN_GRAPH", "0") == "1" and device.type == "cuda" and len(self.optimizers) == 1:
            from .graphs import GraphedTrainStep
            o = self.optimizers[0]
            if not o.state:                                  # fresh optimizer: its step counters can live on the device
                for g_ in o.param_groups:
                    if "capturable" in g_:
                        g_["capturable"] = True
            if all(g_.get("capturable", False) for g_ in o.param_groups):
                # training_step's self.log(...) calls store device tensors in self.logged_metrics; captured once, those are the
                # graph's static outputs and every replay refreshes them in place
                graphed = GraphedTrainStep(model, lambda batch_: model.training_step(batch_, 0), o, sync,
                                           clip_grad_norm=self.gradient_clip_val or None)

        loader = model.train_dataloader()
        sampler = None
        if self.world_size > 1:
            # what Lightning does to the user's loader under DDP: same dataset, a DistributedSampler in place of shuffle
            sampler = DistributedSampler(loader.dataset, self.world_size, self.global_rank,
                                         shuffle=not isinstance(loader.sampler, torch.utils.data.SequentialSampler))
            loader = DataLoader(loader.dataset, batch_size=loader.batch_size, sampler=sampler,
                                num_workers=loader.num_workers, pin_memory=False, drop_last=loader.drop_last,
                                collate_fn=loader.collate_fn,
                                **({"prefetch_factor": loader.prefetch_factor} if loader.num_workers > 0 else {}))
        nb = len(loader)
        if isinstance(self.limit_train_batches, float) and self.limit_train_batches <= 1.0:
            nb = max(1, int(nb * self.limit_train_batches))
        else:
            nb = min(nb, int(self.limit_train_batches))

        for cb in self.callbacks:
            cb.on_train_start(self, model)
        t_last, s_last = time.perf_counter(), self.global_step
        done = False
        try:
            while not done and (self.max_epochs is None or self.current_epoch < self.max_epochs):
                if sampler is not None:
                    sampler.set_epoch(self.current_epoch)
                for cb in self.callbacks:
                    cb.on_train_epoch_start(self, model)
                for batch_idx, batch in enumerate(DevicePrefetcher(loader, device)):
                    if batch_idx >= nb:
                        break
                    if graphed is not None:
                        loss = graphed(batch)
                    else:
                        sync.zero_grad()
                        loss = model.training_step(batch, batch_idx)
                        loss.backward()
                        sync.finish()
                        if self.gradient_clip_val:
                            torch.nn.utils.clip_grad_norm_(trainable, self.gradient_clip_val)
                        for o in self.optimizers:
                            o.step()
                    self.global_step += 1
                    for cb in self.callbacks:
                        cb.on_train_batch_end(self, model, loss, batch, batch_idx)
                    if self.global_step % self.log_every_n_steps == 0:
                        m = self._flush_metrics(device)
                        if self.detect_anomaly and not math.isfinite(m.get("loss", 0.0)):
                            raise RuntimeError(f"non-finite loss {m.get('loss')} at step {self.global_step}")
                        if self.is_global_zero:
                            now = time.perf_counter()
                            rate = (self.global_step - s_last) * batch.shape[0] * self.world_size / max(now - t_last, 1e-9)
                            t_last, s_last = now, self.global_step
                            m["segments_per_s"] = rate
                            if self.logger:
                                self.logger.log_metrics({"epoch": self.current_epoch, **m}, self.global_step)
                            print(f"epoch {self.current_epoch} step {self.global_step} " +
                                  " ".join(f"{k}={v:.5g}" for k, v in m.items()), flush=True)
                    if 0 < self.max_steps <= self.global_step:
                        done = True
                        break
                for cb in self.callbacks:
                    cb.on_train_epoch_end(self, model)
                if self.enable_checkpointing and self.logger:
                    self.save_checkpoint(os.path.join(self.logger.log_dir, "checkpoints",
                                                      f"epoch={self.current_epoch}-step={self.global_step}.ckpt"), model)
                self.current_epoch += 1
        finally:
            sync.remove()
        for cb in self.callbacks:
            cb.on_train_end(self, model)
        if device.type == "cuda":
            torch.cuda.synchronize(device)
