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
