"""Net builders with the reference's names (semilearn/nets/__init__.py), reachable through get_net_builder."""
from .bert import ClassificationBert, bert_base_cased, bert_base_uncased
from .hubert import ClassificationHubert, hubert_base
from .wrn import WideResNet, wrn_28_2, wrn_28_8
from .vit import (VisionTransformer, vit_base_patch16_96, vit_base_patch16_224, vit_small_patch2_32, vit_small_patch16_224,
                  vit_tiny_patch2_32)

__all__ = ["VisionTransformer", "vit_tiny_patch2_32", "vit_small_patch2_32", "vit_small_patch16_224", "vit_base_patch16_96",
           "vit_base_patch16_224", "ClassificationBert", "bert_base_uncased", "bert_base_cased", "ClassificationHubert", "hubert_base", "WideResNet", "wrn_28_2", "wrn_28_8"]
