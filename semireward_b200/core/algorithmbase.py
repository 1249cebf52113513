"""AlgorithmBase — the caller side of the hot path, with the reference's surface (semilearn/core/algorithmbase.py):
ctor `(args, net_builder, tb_log=None, logger=None)` (:64-138), `process_batch` (:282-306), `register_hook` / `call_hook`
/ `registered_hook` (:542-593), `sr_decay` (:177-183), `compute_prob` (:332-333), `train_step` (:335-345), and the
attributes hooks read (`it, epoch, model, ema_model, optimizer, scheduler, use_amp, clip_grad, out_dict, log_dict, ...`).

Out of scope here: datasets/loaders, evaluation, checkpoint/logging hooks.  `dataset`
resolves to None exactly like the reference's get_dataset for an unknown name (build.py:114-115), so algorithms are
driven with tensors the caller provides — the library mode of lighting/trainer.py:61-65."""
from __future__ import annotations

import contextlib
from collections import OrderedDict
from inspect import signature

import torch

from .hooks import EMAHook, Hook, ParamUpdateHook
from .optim import get_cosine_schedule_with_warmup, get_optimizer


class AlgorithmBase:
    def __init__(self, args, net_builder, tb_log=None, logger=None, **kwargs):
        self.args = args
        self.num_classes = args.num_classes
        self.ema_m = args.ema_m
        self.epochs = args.epoch
        self.num_train_iter = args.num_train_iter
        self.num_eval_iter = getattr(args, "num_eval_iter", 0)
        self.num_log_iter = getattr(args, "num_log_iter", 0)
        self.num_iter_per_epoch = int(self.num_train_iter // max(self.epochs, 1))
        self.lambda_u = args.ulb_loss_ratio
        self.use_cat = args.use_cat
        self.use_amp = args.amp
        self.clip_grad = args.clip_grad
        self.save_name = getattr(args, "save_name", None)
        self.save_dir = getattr(args, "save_dir", None)
        self.resume = getattr(args, "resume", False)
        self.algorithm = args.algorithm
        self.tb_log = tb_log
        self.print_fn = print if logger is None else logger.info
        self.ngpus_per_node = torch.cuda.device_count()
        self.amp_cm = contextlib.nullcontext
        self.loss_scaler = torch.amp.GradScaler("cuda")   # algorithmbase.py:94; unused while amp is off, but its state is a checkpoint key
        self.task_type = "cls"
        self.gpu = args.gpu
        self.rank = getattr(args, "rank", 0)
        self.distributed = getattr(args, "distributed", False)
        self.world_size = getattr(args, "world_size", 1)
        self.it = 0
        self.epoch = 0
        self.start_epoch = 0
        self.best_eval_metric, self.best_it = 0.0, 0
        self.net_builder = net_builder
        self.ema = None
        self.dataset_dict = None      # dataset construction is outside the hot path
        self.loader_dict = None
        self.model = self.set_model()
        self.ema_model = self.set_ema_model()
        self.optimizer, self.scheduler = self.set_optimizer()
        self.out_dict, self.log_dict = None, None
        self._hooks = []
        self.hooks_dict = OrderedDict()
        self.set_hooks()

    # -- construction -----------------------------------------------------------------------------
    def set_model(self):
        return self.net_builder(num_classes=self.num_classes, pretrained=getattr(self.args, "use_pretrain", False),
                                pretrained_path=getattr(self.args, "pretrain_path", None))

    def set_ema_model(self):
        ema_model = self.net_builder(num_classes=self.num_classes)
        ema_model.load_state_dict(self.model.state_dict())
        return ema_model

    def set_optimizer(self):
        optimizer = get_optimizer(self.model, self.args.optim, self.args.lr, getattr(self.args, "momentum", 0.9),
                                  self.args.weight_decay, self.args.layer_decay)
        scheduler = get_cosine_schedule_with_warmup(optimizer, self.num_train_iter, num_warmup_steps=self.args.num_warmup_iter)
        return optimizer, scheduler

    def set_hooks(self):
        self.register_hook(ParamUpdateHook(), None, "HIGHEST")
        self.register_hook(EMAHook(), None, "HIGH")   # algorithmbase.py:229-230: parameter update first, then the EMA

    # -- hook protocol ----------------------------------------------------------------------------
    _PRIORITY = dict(HIGHEST=0, VERY_HIGH=10, HIGH=30, ABOVE_NORMAL=40, NORMAL=50, BELOW_NORMAL=60, LOW=70, VERY_LOW=90, LOWEST=100)

    def register_hook(self, hook, name=None, priority="NORMAL"):
        assert isinstance(hook, Hook) or all(hasattr(hook, st) for st in Hook.stages), "a hook implements the six stage methods"
        if hasattr(hook, "priority"):
            raise ValueError('"priority" is a reserved attribute for hooks')
        hook.priority = self._PRIORITY[priority] if isinstance(priority, str) else int(priority)
        hook.name = name if name is not None else type(hook).__name__
        pos = len(self._hooks)
        for i in range(len(self._hooks) - 1, -1, -1):
            if hook.priority >= self._hooks[i].priority:
                pos = i + 1
                break
            pos = i
        self._hooks.insert(pos, hook)
        self.hooks_dict = OrderedDict((h.name, h) for h in self._hooks)

    def call_hook(self, fn_name, hook_name=None, *args, **kwargs):
        if hook_name is not None:
            return getattr(self.hooks_dict[hook_name], fn_name)(self, *args, **kwargs)
        for hook in self.hooks_dict.values():
            if hasattr(hook, fn_name):
                getattr(hook, fn_name)(self, *args, **kwargs)

    def registered_hook(self, hook_name):
        return hook_name in self.hooks_dict

    # -- per-step helpers -------------------------------------------------------------------------
    def process_batch(self, input_args=None, **kwargs):
        """Keeps only the batch keys that appear in train_step's signature and moves them to cuda:gpu
        (algorithmbase.py:282-306)."""
        if input_args is None:
            input_args = list(signature(self.train_step).parameters.keys())
        out = {}
        for arg, var in kwargs.items():
            if arg not in input_args or var is None:
                continue
            if isinstance(var, dict):
                var = {k: v.cuda(self.gpu, non_blocking=True) for k, v in var.items()}
            else:
                var = var.cuda(self.gpu, non_blocking=True)
            out[arg] = var
        return out

    def process_out_dict(self, out_dict=None, **kwargs):
        out_dict = {} if out_dict is None else out_dict
        out_dict.update(kwargs)
        return out_dict

    def process_log_dict(self, log_dict=None, prefix="train", **kwargs):
        log_dict = {} if log_dict is None else log_dict
        for k, v in kwargs.items():
            log_dict[f"{prefix}/{k}"] = v
        return log_dict

    def compute_prob(self, logits):
        return torch.softmax(logits, dim=-1)

    def sr_decay(self, max_sampling_time=8):
        return int(max(max_sampling_time, 1 + self.num_train_iter / self.it))

    def train_step(self, idx_lb, x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s):
        raise NotImplementedError

    def train_one_step(self, **batch):
        """before_train_step -> train_step -> after_train_step, as in AlgorithmBase.train (algorithmbase.py:362-371)."""
        self.call_hook("before_train_step")
        self.out_dict, self.log_dict = self.train_step(**self.process_batch(**batch))
        self.call_hook("after_train_step")
        self.it += 1
        return self.out_dict, self.log_dict

    # -- checkpoint (keys of algorithmbase.py:459-525) ----------------------------------------------
    def get_save_dict(self):
        """Same keys as the reference writes (algorithmbase.py:459-485), so either side loads the other's file.  `loss_scaler`
        is the state of a disabled-in-effect GradScaler: every SemiReward config runs `amp: False`, the reference still writes
        and reads the key unconditionally (:472, :505)."""
        d = dict(model=self.model.state_dict(), ema_model=self.ema_model.state_dict(), optimizer=self.optimizer.state_dict(),
                 loss_scaler=self._loss_scaler_state(), it=self.it + 1, epoch=self.epoch + 1, best_it=self.best_it,
                 best_eval_acc=self.best_eval_metric)
        if self.scheduler is not None:
            d["scheduler"] = self.scheduler.state_dict()
        return d

    def _loss_scaler_state(self):
        """State of the GradScaler as the reference would write it.  Without CUDA (CPU tests) torch disables the scaler and
        its state is {}, which an enabled scaler refuses to load (the reference's load_model, :505): write the initial state."""
        sd = self.loss_scaler.state_dict()
        return sd if sd else {"scale": 65536.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 0}

    def save_model(self, save_name, save_path):
        import os
        os.makedirs(save_path, exist_ok=True)
        fn = os.path.join(save_path, save_name)
        torch.save(self.get_save_dict(), fn)
        self.print_fn(f"model saved: {fn}")

    def load_model(self, load_path):
        """algorithmbase.py:498-525: model, EMA model, loss scaler, counters, optimizer (AdamW moments + step) and scheduler
        are all restored, so a resumed run continues the learning-rate schedule and the moments instead of restarting them."""
        ck = torch.load(load_path, map_location="cpu")
        self.model.load_state_dict(ck["model"])
        self.ema_model.load_state_dict(ck["ema_model"])
        if ck.get("loss_scaler") and self.loss_scaler.is_enabled():   # files written before the key existed here simply lack it
            self.loss_scaler.load_state_dict(ck["loss_scaler"])
        self.it, self.start_epoch, self.epoch = ck["it"], ck["epoch"], ck["epoch"]
        self.best_it, self.best_eval_metric = ck["best_it"], ck["best_eval_acc"]
        self.optimizer.load_state_dict(ck["optimizer"])
        if self.scheduler is not None and ck.get("scheduler") is not None:
            self.scheduler.load_state_dict(ck["scheduler"])
        if hasattr(self.model, "mark_weights_updated"):
            self.model.mark_weights_updated()
        self.print_fn("Model loaded")
        return ck

    @staticmethod
    def get_argument():
        return []
