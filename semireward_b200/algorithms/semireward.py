"""SemiReward modules behind the reference's class names (semilearn/algorithms/semireward/semireward.py):
Rewarder (:27-72), EMARewarder (:75-127), Generator (:6-24), cosine_similarity_n (:130-139), label_dim (:147-148).

The nn.Modules are parameter holders with the reference's state_dict keys and default initialisation; forward() and the
online training step run as single fused kernels (srw_rewarder_fwd / srw_generator_fwd / srw_rewarder_train)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _lib as L


def label_dim(x, default_dim=100):
    return int(max(default_dim, x))


def cosine_similarity_n(x1, x2):
    """(cos + 1) / 2 of two one-hot label matrices -> {1.0, 0.5}.  Kept for API parity; the fused training kernel
    evaluates the same target in closed form (equal ? 1 : 0.5)."""
    cos = torch.cosine_similarity(x1, x2, dim=-1, eps=1e-8)
    return ((cos + 1) / 2).view(x1.size(0), 1)


def _ws(B, dev, feature_dim):
    n = L.load().srw_rewarder_workspace_floats(B, feature_dim)
    return torch.empty(max(n, B * 449 + 16), dtype=torch.float32, device=dev)


class Generator(nn.Module):
    def __init__(self, feature_dim=384):
        super().__init__()
        self.feature_dim = feature_dim
        self.fc_layers = nn.Sequential(nn.Linear(feature_dim, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU(),
                                       nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 1))

    def _params(self):
        return [t for i in (0, 2, 4, 6) for t in (self.fc_layers[i].weight, self.fc_layers[i].bias)]

    @torch.no_grad()
    def generate_labels(self, feats):
        """relu(fc_layers(x)).long() as one kernel -> int64 [B] (srflexmatch.py:157-158).  `.long()` cuts the graph in
        the reference too: the Generator never receives a gradient."""
        feats = feats.detach()
        if feats.stride(-1) != 1:
            feats = feats.contiguous()
        B = feats.shape[0]
        labels = torch.empty(B, dtype=torch.long, device=feats.device)
        ws = _ws(B, feats.device, self.feature_dim)
        a = L.GeneratorFwdArgs(B=B, feature_dim=self.feature_dim, gp=L.ptr_array(self._params()), feats=feats.data_ptr(),
                               ld_feats=feats.stride(0), labels=labels.data_ptr(), workspace=ws.data_ptr())
        L.check(L.load().srw_generator_fwd(C.byref(a), L.stream_ptr()), "srw_generator_fwd")
        return labels

    def forward(self, x):
        raise NotImplementedError("use generate_labels(): the float output is only ever consumed through .long()")


class Rewarder(nn.Module):
    def __init__(self, label_dim, label_embedding_dim, feature_dim=384):
        super().__init__()
        if label_embedding_dim != 128:
            raise NotImplementedError("label_embedding_dim is 128 in every SemiReward algorithm (srflexmatch.py:49)")
        self.label_rows, self.feature_dim = int(label_dim), int(feature_dim)
        self.feature_fc = nn.Linear(feature_dim, 128)
        self.feature_norm = nn.LayerNorm(128)
        self.label_embedding = nn.Embedding(label_dim, label_embedding_dim)
        self.label_norm = nn.LayerNorm(label_embedding_dim)
        self.cross_attention_fc = nn.Linear(128, 1)
        self.mlp_fc1 = nn.Linear(128, 256)
        self.mlp_fc2 = nn.Linear(256, 128)
        self.ffn_fc1 = nn.Linear(128, 64)
        self.ffn_fc2 = nn.Linear(64, 1)
        self._adam = None

    def _params(self):
        return [self.feature_fc.weight, self.feature_fc.bias, self.feature_norm.weight, self.feature_norm.bias,
                self.label_embedding.weight, self.label_norm.weight, self.label_norm.bias, self.cross_attention_fc.weight,
                self.cross_attention_fc.bias, self.mlp_fc1.weight, self.mlp_fc1.bias, self.mlp_fc2.weight, self.mlp_fc2.bias,
                self.ffn_fc1.weight, self.ffn_fc1.bias, self.ffn_fc2.weight, self.ffn_fc2.bias]

    @torch.no_grad()
    def forward(self, features, label_indices):
        """reward [B,1] in (0,1).  Inference only: every consumer in the SR algorithms either detaches the inputs or
        thresholds the output (srflexmatch.py:99-101,165-169); the training path is train_step() below."""
        feats = features.detach()
        if feats.stride(-1) != 1:
            feats = feats.contiguous()
        B = feats.shape[0]
        labels = label_indices.to(torch.long).contiguous()
        reward = torch.empty(B, 1, dtype=torch.float32, device=feats.device)
        ws = _ws(B, feats.device, self.feature_dim)
        a = L.RewarderFwdArgs(B=B, feature_dim=self.feature_dim, label_rows=self.label_rows, rp=L.ptr_array(self._params()),
                              feats=feats.data_ptr(), ld_feats=feats.stride(0), labels=labels.data_ptr(), reward=reward.data_ptr(),
                              workspace=ws.data_ptr())
        L.check(L.load().srw_rewarder_fwd(C.byref(a), L.stream_ptr()), "srw_rewarder_fwd")
        return reward

    # -- checkpoint of the online-update optimizer (the reference's checkpoints drop the whole SemiReward state) --
    def optimizer_state(self):
        """Adam moments and step count of the fused online update on the CPU, None before the first update."""
        if self._adam is None:
            return None
        return dict(m=[t.detach().cpu() for t in self._adam["m"]], v=[t.detach().cpu() for t in self._adam["v"]], step=int(self._adam["step"]))

    def load_optimizer_state(self, state):
        if state is None:
            self._adam = None
            return
        ps = self._params()
        dev = ps[0].device
        gflat = torch.empty(sum(p.numel() for p in ps), dtype=torch.float32, device=dev)
        gs, off = [], 0
        for p in ps:
            gs.append(gflat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._adam = dict(m=[t.to(dev).clone() for t in state["m"]], v=[t.to(dev).clone() for t in state["v"]], g=gs, gflat=gflat, step=int(state["step"]))

    @torch.no_grad()
    def train_step(self, features, gen_labels, true_labels, lr, num_classes):
        """One online update (srflexmatch.py:173-208): reward = R(features, gen_labels); generator_loss = MSE(reward, 1);
        rewarder_loss = MSE(reward, cos-target(gen, true)); both backward passes and one torch.optim.Adam step, fused.
        Returns a [2] device tensor (generator_loss, rewarder_loss)."""
        ps = self._params()
        dev = ps[0].device
        if self._adam is None or self._adam["m"][0].device != dev:
            gflat = torch.empty(sum(p.numel() for p in ps), dtype=torch.float32, device=dev)
            gs, off = [], 0
            for p in ps:
                gs.append(gflat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self._adam = dict(m=[torch.zeros_like(p) for p in ps], v=[torch.zeros_like(p) for p in ps], g=gs, gflat=gflat, step=0)
        st = self._adam
        st["step"] += 1
        feats = features.detach()
        if feats.stride(-1) != 1:
            feats = feats.contiguous()
        B = feats.shape[0]
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        ws = _ws(B, dev, self.feature_dim)
        gl, tl = gen_labels.to(torch.long).contiguous(), true_labels.to(torch.long).contiguous()
        a = L.RewarderTrainArgs(B=B, feature_dim=self.feature_dim, label_rows=self.label_rows, num_classes=int(num_classes),
                                rp=L.ptr_array(ps), g=L.ptr_array(st["g"]), m=L.ptr_array(st["m"]), v=L.ptr_array(st["v"]),
                                feats=feats.data_ptr(), ld_feats=feats.stride(0), gen_labels=gl.data_ptr(), true_labels=tl.data_ptr(),
                                lr=float(lr), step=st["step"], phase=0, losses=losses.data_ptr(), workspace=ws.data_ptr())
        group = getattr(self, "_dp_group", None)
        if group is None:
            L.check(L.load().srw_rewarder_train(C.byref(a), L.stream_ptr()), "srw_rewarder_train")
        else:
            # Data parallel.  The reference wraps the Rewarder in torch DDP (srflexmatch.py:49-51) and calls backward twice after one
            # forward (:204-205); DDP's reducer only synchronises the first of them, so its optimizer steps with
            # mean_ranks(grad generator_loss) + LOCAL grad rewarder_loss and the ranks' Rewarders drift apart
            # (scripts/c3_ddp_probe.py runs the reference's module under gloo DDP).  Same here: the generator-loss gradient
            # is averaged, the rewarder-loss gradient stays local, one Adam step on their sum.
            from ..parallel import allreduce_mean_
            if "g2" not in st:
                g2flat = torch.empty_like(st["gflat"])
                st["g2"], off = [], 0
                for p in ps:
                    st["g2"].append(g2flat[off:off + p.numel()].view_as(p))
                    off += p.numel()
            losses2 = torch.empty(2, dtype=torch.float32, device=dev)
            a.phase, a.loss_select = 1, 1
            L.check(L.load().srw_rewarder_train(C.byref(a), L.stream_ptr()), "srw_rewarder_train")
            allreduce_mean_(st["gflat"], group)
            a.loss_select, a.g, a.losses = 2, L.ptr_array(st["g2"]), losses2.data_ptr()
            L.check(L.load().srw_rewarder_train(C.byref(a), L.stream_ptr()), "srw_rewarder_train")
            a.phase, a.loss_select, a.g, a.g_add = 2, 0, L.ptr_array(st["g"]), L.ptr_array(st["g2"])
            L.check(L.load().srw_rewarder_train(C.byref(a), L.stream_ptr()), "srw_rewarder_train")
        return losses


class EMARewarder(Rewarder):
    """semireward.py:75-127 (`sr_ema: True`, the argparse default; every shipped YAML sets it to False): the same forward as
    Rewarder on the LIVE parameters, followed by `ema = decay * ema + (1 - decay) * param` over all 17 tensors.  The averaged
    copies (`ema_params`, keyed by parameter name like the reference's dict) are never read by any algorithm there either;
    they are kept so that code written against the reference finds them.  One srw_ema_step launch per forward."""

    def __init__(self, label_dim, label_embedding_dim, feature_dim=384, ema_decay=0.9):
        super().__init__(label_dim, label_embedding_dim, feature_dim)
        self.ema_decay = float(ema_decay)
        self.ema_params = {}
        self.initialize_ema()
        self._ema_table = self._ema_key = None

    def initialize_ema(self):
        for name, param in self.named_parameters():
            if param.requires_grad:
                self.ema_params[name] = nn.Parameter(param.data.clone())

    @torch.no_grad()
    def update_ema(self):
        named = [(n, p) for n, p in self.named_parameters() if p.requires_grad]
        key = tuple(p.data_ptr() for _, p in named)
        if self._ema_key != key:
            rows = (L.EmaRow * len(named))()
            blk = 0
            for i, (n, p) in enumerate(named):
                e = self.ema_params[n]
                if e.device != p.device:   # semireward.py:100-101: first use after .cuda() restarts the average from the parameter
                    e.data = p.data.clone()
                rows[i].param, rows[i].shadow, rows[i].numel, rows[i].first_block = p.data_ptr(), e.data_ptr(), p.numel(), blk
                blk += (p.numel() + L.ADAMW_BLOCK_ELEMS - 1) // L.ADAMW_BLOCK_ELEMS
            host = torch.empty(C.sizeof(rows), dtype=torch.uint8)
            C.memmove(host.data_ptr(), C.addressof(rows), C.sizeof(rows))
            self._ema_table, self._ema_key, self._ema_n, self._ema_blocks = host.to(named[0][1].device), key, len(named), blk
            return
        a = L.EmaArgs(num_tensors=self._ema_n, total_blocks=self._ema_blocks, table=self._ema_table.data_ptr(), decay=self.ema_decay)
        L.check(L.load().srw_ema_step(C.byref(a), L.stream_ptr()), "srw_ema_step")

    @torch.no_grad()
    def forward(self, features, label_indices):
        reward = super().forward(features, label_indices)
        self.update_ema()
        return reward
