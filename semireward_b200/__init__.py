"""semireward_b200 — B200-native (sm_100a) implementation of SemiReward's per-step SSL train-step hot path behind the
reference's plugin surface:

    from semireward_b200 import get_config, get_net_builder, get_algorithm
    args = get_config({...reference YAML keys...})
    alg = get_algorithm(args, get_net_builder(args.net, False), None, None)
    out_dict, log_dict = alg.train_step(**alg.process_batch(**batch)); alg.call_hook("after_train_step")

Everything numeric runs in libsrw_b200.so (include/srw.h); importing the package does not need a GPU, using it does."""
from .core.registry import ALGORITHMS  # noqa: F401
from .algorithms import get_algorithm, name2alg  # noqa: F401
from .config import get_config  # noqa: F401
from . import nets  # noqa: F401


def get_net_builder(net_name, from_name: bool = False):
    """semilearn/core/utils/build.py:14-39: builders are attributes of the nets package."""
    if from_name:
        raise NotImplementedError("net_from_name (torchvision model zoo) is outside the hot path")
    if not hasattr(nets, net_name):
        raise KeyError(f"[!] Networks' Name is wrong, check net config, expected one of {sorted(nets.__all__)}, received: {net_name}")
    return getattr(nets, net_name)
