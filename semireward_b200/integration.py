"""Running the native step under the reference's own loop.

`reference_algorithm(native_cls)` builds `class X(native_cls, semilearn.core.AlgorithmBase)`:
  * construction is the native one (model, optimizer, Rewarder/Generator on the native kernels; no datasets are built — the
    caller assigns `dataset_dict` / `loader_dict`, e.g. from the reference's own `set_dataset()` / `set_data_loader()`);
  * `train()`, `evaluate()`, `set_dataset()`, `set_data_loader()` and the imbalanced-algorithm mixin protocol resolve to the
    reference's methods (semilearn/core/algorithmbase.py:346-457, 146-222);
  * `set_hooks()` registers the native hooks of the step (ParamUpdateHook, EMAHook, PseudoLabelingHook, MaskingHook, ...) AND the
    reference's loop hooks with their priorities (algorithmbase.py:264-281): EvaluationHook, CheckpointHook,
    DistSamplerSeedHook, TimerHook, LoggingHook — so a `train()` run evaluates, checkpoints and logs like the reference;
  * the attributes those hooks and `evaluate()` read exist: `task_type`, `ce_loss`, `consistency_loss`, `loss_scaler`,
    `bn_controller`, `num_eval_iter`, `num_log_iter`, `save_dir`, `save_name`, `resume`, `results_dict`.

`register_into_reference()` puts those classes into `semilearn`'s ALGORITHMS registry under the reference's names and the
native builders into `semilearn.nets`, so `python train.py --c config/SemiReward/usb_cv/...yaml` picks them up unchanged.
Needs the `semilearn` package importable; nothing here is used by the native package itself."""
from __future__ import annotations

_SR_NAMES = ("srflexmatch", "srfixmatch", "srfreematch", "srsoftmatch", "srpseudolabel")


def reference_algorithm(native_cls, ref_base=None):
    if ref_base is None:
        from semilearn.core import AlgorithmBase as ref_base
    import semilearn.core.hooks as RH

    class Combined(native_cls, ref_base):
        __doc__ = f"{native_cls.__name__}: native train_step / hooks of the step, the reference's loop, evaluation and checkpoint hooks"

        def __init__(self, args, net_builder, tb_log=None, logger=None, **kwargs):
            native_cls.__init__(self, args, net_builder, tb_log, logger, **kwargs)   # does not chain into the reference ctor (datasets)
            from semilearn.core.criterions import CELoss, ConsistencyLoss
            from semilearn.core.utils import Bn_Controller
            self.ce_loss, self.consistency_loss = CELoss(), ConsistencyLoss()   # evaluate()'s regression branch / user hooks read them
            self.bn_controller = Bn_Controller()
            self.results_dict = {}

        def set_hooks(self):
            super().set_hooks()
            for hook, prio in ((RH.EvaluationHook(), "HIGH"), (RH.CheckpointHook(), "HIGH"), (RH.DistSamplerSeedHook(), "NORMAL"),
                               (RH.TimerHook(), "LOW"), (RH.LoggingHook(), "LOWEST")):
                self.register_hook(hook, None, prio)

    Combined.__name__ = native_cls.__name__
    Combined.__qualname__ = native_cls.__qualname__
    return Combined


def register_into_reference():
    """-> {name: combined class}.  Idempotent."""
    import semilearn.nets as ref_nets
    from semilearn.core.utils import ALGORITHMS as REF

    from . import nets
    from .core.registry import ALGORITHMS
    out = {}
    for name in _SR_NAMES:
        cur = REF[name] if name in REF else None
        if cur is not None and getattr(cur, "_srw_native", False):
            out[name] = cur
            continue
        cls = reference_algorithm(ALGORITHMS[name])
        cls._srw_native = True
        REF[name] = cls
        out[name] = cls
    for b in nets.__all__:
        if b[0].islower():
            setattr(ref_nets, b, getattr(nets, b))
    return out
