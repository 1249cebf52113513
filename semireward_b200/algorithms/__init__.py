"""Algorithm lookup with the reference's surface (semilearn/algorithms/__init__.py:8-18)."""
from ..core.registry import ALGORITHMS
from . import srfixmatch, srflexmatch, srfreematch, srpseudolabel, srsoftmatch  # noqa: F401  (register the five SR algorithms under the reference's names)

name2alg = ALGORITHMS


def get_algorithm(args, net_builder, tb_log, logger):
    if args.algorithm in ALGORITHMS:
        return ALGORITHMS[args.algorithm](args=args, net_builder=net_builder, tb_log=tb_log, logger=logger)
    raise KeyError(f"Unknown algorithm: {str(args.algorithm)}")
