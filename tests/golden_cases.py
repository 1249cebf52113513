"""Cases shared by tests/golden/make_golden.py (live reference -> fixtures) and tests/test_oracle_golden.py."""
STEPS = 7   # it 0 (no SR), 1-2 stage 1, 3 gap step (it == start_timing), 4 stage 2 + SR update (it % N_k == 0), 5 stage 2, 6 SR update

CASES = {
    "srflexmatch_d2": dict(cfg=dict(algorithm="srflexmatch"), depth=2, head_gain=4.0),
    "srflexmatch_d2_mixedmask": dict(cfg=dict(algorithm="srflexmatch", p_cutoff=0.3, ulb_dest_len=16), depth=2, head_gain=1.0),
    "srfreematch_d2": dict(cfg=dict(algorithm="srfreematch"), depth=2, head_gain=2.0),
    "srsoftmatch_d2": dict(cfg=dict(algorithm="srsoftmatch"), depth=2, head_gain=2.0),
    "srpseudolabel_d2": dict(cfg=dict(algorithm="srpseudolabel", p_cutoff=0.5, unsup_warm_up=0.05), depth=2, head_gain=2.0),
    "srfixmatch_d2": dict(cfg=dict(algorithm="srfixmatch", p_cutoff=0.5), depth=2, head_gain=2.0),   # mixed masks in every stage
}

# Text path (BASELINE configs[3]): ClassificationBert with a 2-layer random-init BertModel, use_cat False,
# dropout 0 (deterministic parity mode).  oracle/bert_oracle.py is pinned by tests/golden/bert_*.npz (make_golden_bert.py).
BERT_SMALL = dict(vocab_size=512, layers=2, max_position=64, max_length=32)
BERT_CASES = {
    # the BASELINE configs[3] algorithm: SoftMatch weights + uniform DistAlign, 2 classes (IMDb)
    "bert_srsoftmatch_l2": dict(cfg=dict(algorithm="srsoftmatch", num_classes=2), head_gain=4.0),
    # fixed threshold between the weak-view confidences (4 classes, ag_news) -> mixed masks in every stage
    "bert_srfixmatch_l2": dict(cfg=dict(algorithm="srfixmatch", num_classes=4, p_cutoff=0.85), head_gain=3.0),
}


def bert_small_cfg(**over):
    c = dict(algorithm="srsoftmatch", net="bert_base_uncased", optim="AdamW", lr=5e-5, layer_decay=0.75, weight_decay=5e-4,
             num_train_iter=64, num_warmup_iter=0, start_timing=3, N_k=2, batch_size=4, uratio=1, num_classes=2, ulb_dest_len=64,
             feature_dim=768, sr_lr=5e-4, sr_ema=False, use_cat=False, amp=False, ema_m=0.0, thresh_warmup=True, p_cutoff=0.95)
    c.update(over)
    return c

# Convolutional path (BASELINE configs[0]): WideResNet (depth 10 for the fixtures, widen 2), use_cat True,
# SGD + nesterov.  oracle/wrn_oracle.py is pinned by tests/golden/wrn_*.npz (make_golden_wrn.py).
WRN_CASES = {
    "wrn_srflexmatch_d10": dict(cfg=dict(algorithm="srflexmatch"), depth=10, head_gain=4.0),
    "wrn_srfixmatch_d10": dict(cfg=dict(algorithm="srfixmatch", p_cutoff=0.2), depth=10, head_gain=4.0),   # mixed masks in every stage
}


def wrn_small_cfg(**over):
    c = dict(algorithm="srflexmatch", net="wrn_28_2", optim="SGD", lr=0.03, momentum=0.9, layer_decay=1.0, weight_decay=1e-3,
             num_train_iter=64, num_warmup_iter=0, start_timing=3, N_k=2, batch_size=4, uratio=2, num_classes=100, ulb_dest_len=64,
             feature_dim=128, sr_lr=5e-4, sr_ema=False, use_cat=True, amp=False, ema_m=0.999, thresh_warmup=True, p_cutoff=0.95)
    c.update(over)
    return c

# Audio path (BASELINE configs[4]): ClassificationHubert with a 2-layer random-init HubertModel (full conv
# stem), use_cat False, 4000-sample clips (12 frames), every source of randomness off (dropout, LayerDrop, SpecAugment).
HUBERT_SMALL = dict(layers=2, samples=4000)
HUBERT_CASES = {
    "hubert_srflexmatch_l2": dict(cfg=dict(algorithm="srflexmatch"), head_gain=4.0),                   # the configs[4] algorithm
    "hubert_srfixmatch_l2": dict(cfg=dict(algorithm="srfixmatch", p_cutoff=0.5), head_gain=4.0),       # mixed masks
}


def hubert_small_cfg(**over):
    c = dict(algorithm="srflexmatch", net="hubert_base", optim="AdamW", lr=2e-5, layer_decay=0.75, weight_decay=5e-4, num_train_iter=64,
             num_warmup_iter=0, start_timing=3, N_k=2, batch_size=2, uratio=2, num_classes=10, ulb_dest_len=64, feature_dim=768, sr_lr=5e-4,
             sr_ema=False, use_cat=False, amp=False, ema_m=0.0, thresh_warmup=True, p_cutoff=0.95)
    c.update(over)
    return c


# Image input pipeline: the reference's transform_weak / transform_strong on 32 x 32 CIFAR-shaped images.
# tests/golden/augment_cifar.npz (make_golden_augment.py) stores the source images, the decisions the reference's generators
# produced, and the tensors the imported reference returned.
def augment_decisions_to_arrays(decs):
    import numpy as np
    n = len(decs)
    geo = np.zeros((n, 3), dtype=np.int64)
    op_id = np.full((n, 3), -1, dtype=np.int64)
    op_val = np.zeros((n, 3), dtype=np.float64)
    cut = np.full((n, 4), np.nan, dtype=np.float64)
    for i, d in enumerate(decs):
        geo[i] = (d.crop_top, d.crop_left, int(d.flip))
        for k, (o, v) in enumerate(d.ops):
            op_id[i, k], op_val[i, k] = o, v
        if d.cutout is not None:
            cut[i] = d.cutout
    return dict(geo=geo, op_id=op_id, op_val=op_val, cut=cut)


def augment_decisions_from_arrays(z):
    import numpy as np
    from oracle.augment_oracle import Decision
    out = []
    for i in range(len(z["geo"])):
        ops = [(int(o), float(v)) for o, v in zip(z["op_id"][i], z["op_val"][i]) if o >= 0]
        cut = None if np.isnan(z["cut"][i][0]) else tuple(float(v) for v in z["cut"][i])
        out.append(Decision(int(z["geo"][i][0]), int(z["geo"][i][1]), bool(z["geo"][i][2]), ops, cut))
    return out
