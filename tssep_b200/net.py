"""Drop-in ``MaskEstimator_v2`` (tssep/train/net.py:333-986) and its helper modules.

Constructor keywords, attribute names, module tree (hence ``repr`` and ``state_dict``
keys ``pre_net.net.0.weight_ih_l0`` ... ``post_net.linear2.bias``) and the
``forward(xs, aux) -> Output`` contract equal the reference's.  ``forward`` never
materialises the reference's intermediate tensors:

* conditioning (net.py:862-896) is folded into the first post_net input
  projection (``tssep_fold_embedding``): 'mul' scales the weight columns per
  speaker, 'cat' turns the embedding half of the weight into a per-speaker bias;
* the permutation-averaging trials (net.py:900-924) are de-duplicated: birnn0 /
  birnn1 are speaker-independent, so they run once for the K speakers and the R
  cyclic speaker orders only enter through R column-rotated copies of the
  speaker-concat layer's input weights (net.py:606-612);
* the trial mean, the inverse rotation, the speaker un-permutation and the
  sigmoid (net.py:928-986) live in the head GEMM: rotated head weights are
  concatenated along K, the mean is the fp32 accumulation, the un-permutation
  is the epilogue's plane map.
"""
from __future__ import annotations

import collections
import dataclasses

import numpy as np
import torch

from . import _lib, ops, torch_ops
from .configurable import Configurable
from .rnnp import RNNP_packed, param_key


@dataclasses.dataclass
class Output:
    mask: torch.Tensor
    logit: torch.Tensor
    embedding: torch.Tensor = None

    vad_mask: torch.Tensor = None
    vad_logit: torch.Tensor = None


class InstanceNorm(torch.nn.Module):
    """``(x - mean) / std`` over ``dim`` (tssep/train/net.py:250-285): population std unless ``unbiased``."""

    def __init__(self, dim=-1, unbiased=False):
        super().__init__()
        self.dim = dim
        self.unbiased = unbiased

    def extra_repr(self):
        return f"dim={self.dim!r}, unbiased={self.unbiased!r}"

    def forward(self, x):
        if not isinstance(self.dim, int):
            raise NotImplementedError("InstanceNorm over several axes at once is not implemented on the CUDA path")
        return ops.instance_norm(x, self.unbiased, dim=self.dim).to(x.dtype)


class InstanceNorm_v2(torch.nn.Module):
    """Mean removal over ``mean_dim``, then division by the root mean square over ``norm_dim``
    (tssep/train/net.py:288-330; equals ``InstanceNorm`` when both axes coincide)."""

    def __init__(self, mean_dim=-1, norm_dim=-1):
        super().__init__()
        self.mean_dim = mean_dim
        self.norm_dim = norm_dim

    def extra_repr(self):
        return f"mean_dim={self.mean_dim!r}, norm_dim={self.norm_dim!r}"

    def forward(self, x):
        if not isinstance(self.mean_dim, int) or not isinstance(self.norm_dim, int):
            # the reference itself cannot run this (x.shape[tuple] raises, net.py:326)
            raise NotImplementedError("InstanceNorm_v2 over several axes at once")
        if self.mean_dim % x.dim() == self.norm_dim % x.dim():
            return ops.instance_norm(x, False, dim=self.mean_dim).to(x.dtype)
        y = ops.instance_norm(x, dim=self.mean_dim, mode=1)
        return ops.instance_norm(y, dim=self.norm_dim, mode=2).to(x.dtype)


class _Einop(torch.nn.Module):
    """Placeholder keeping the reference's layer names; the rearrangement itself is fused."""

    def __init__(self, pattern, **axes):
        super().__init__()
        self.pattern, self.axes = pattern, axes

    def extra_repr(self):
        return ", ".join([repr(self.pattern)] + [f"{k}={v}" for k, v in self.axes.items()])


class Sequential(torch.nn.Sequential):
    """Container only (tssep/train/net.py:190-237); ``MaskEstimator_v2.forward`` drives the kernels."""


class _SequentialDict:
    """Names layers ``<key><idx>`` with a non-decreasing index (tssep/train/net.py:572-587)."""

    def __init__(self):
        self.data = collections.OrderedDict()
        self.idx = 0

    def add(self, key, value):
        for self.idx in range(self.idx, 100):
            k = f"{key}{self.idx}"
            if k not in self.data:
                self.data[k] = value
                return
        raise RuntimeError(key)


class _PinnedRing:
    """A few persistent pinned staging buffers for small host -> device uploads that must not block the host."""

    def __init__(self, depth: int = 4):
        self.bufs, self.events, self.i = [None] * depth, [None] * depth, 0

    # a derived cache: copies and pickles of the owning module start with an empty ring
    def __deepcopy__(self, memo):
        return _PinnedRing(len(self.bufs))

    def __getstate__(self):
        return {"depth": len(self.bufs)}

    def __setstate__(self, state):
        self.__init__(state["depth"])

    def upload(self, arr: np.ndarray, device) -> torch.Tensor:
        i, self.i = self.i, (self.i + 1) % len(self.bufs)
        n = int(arr.size)
        if self.bufs[i] is None or self.bufs[i].numel() < n:
            self.bufs[i] = torch.empty(max(n, 4096), dtype=torch.int32).pin_memory()  # first uses only
            self.events[i] = None
        if self.events[i] is not None:
            self.events[i].synchronize()  # the copy issued `depth` uploads ago; long done
        self.bufs[i][:n].copy_(torch.from_numpy(arr))
        out = self.bufs[i][:n].to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self.events[i] = ev
        return out


class MaskEstimator_v2(Configurable, torch.nn.Module):
    @classmethod
    def finalize_dogmatic_config(cls, config):
        config["aux_net"] = None
        if config["aux_net_output_size"] is None:
            config["aux_net_output_size"] = 100  # I-vectors by default (net.py:488-490)

    def __init__(self, *, idim=80, odim=None, layers=3, units=300, projs=320, dropout=0, nmask=1, pre_net="RNNP",
                 aux_net=None, aux_net_output_size=None, combination: str = "cat", ts_vad=False,
                 output_resolution: str = "tf", random_speaker_order=True, num_averaged_permutations=1,
                 input_normalizer=None, aux_normalizer=None, explicit_vad=False):
        super().__init__()
        if aux_net is not None:
            raise NotImplementedError("aux_net is forced to None by the reference (net.py:487)")
        if odim is None:
            odim = idim
        self.odim, self.nmask = odim, nmask
        self.output_resolution = output_resolution
        self.random_speaker_order = random_speaker_order
        self.num_averaged_permutations = num_averaged_permutations
        self.ts_vad = ts_vad
        self.input_normalizer, self.aux_normalizer = input_normalizer, aux_normalizer
        self.explicit_vad = explicit_vad
        self.layers, self.units, self.projs = layers, units, projs
        if not ts_vad:
            assert num_averaged_permutations == 1, (ts_vad, num_averaged_permutations)
        if pre_net == "RNNP":
            self.pre_net = RNNP_packed(idim=idim, elayers=1, cdim=units, hdim=odim, dropout=dropout, typ="blstm")
        elif pre_net in [None, False]:
            self.pre_net = torch.nn.Identity()
        else:
            raise ValueError(pre_net)
        self.aux_net = aux_net
        self.combination = combination
        if combination == "cat":
            assert aux_net_output_size is not None, (combination, aux_net_output_size)
            first = odim + aux_net_output_size
        elif combination == "mul":
            first = odim
        else:
            raise NotImplementedError(combination) if combination == "film" else ValueError(combination)
        self.aux_size = aux_net_output_size
        post = _SequentialDict()
        ts_factor = 1
        for l in range(layers):
            if l == layers - 1 and ts_vad is not False:
                assert 2 < ts_vad < 20, ts_vad
                post.add("rearrange", _Einop("... spk time feature -> ... 1 time (spk feature)", spk=ts_vad))
                ts_factor = ts_vad
            post.add("birnn", RNNP_packed(idim=(first if l == 0 else projs) * ts_factor, elayers=1, cdim=units,
                                          hdim=projs, dropout=dropout, typ="blstm"))
            if l < layers - 1:
                post.add("dropout", torch.nn.Dropout(p=dropout))
                post.add("activation", torch.nn.Tanh())
        if output_resolution == "tf":
            out_features = (odim + int(explicit_vad)) * nmask * ts_factor
            pattern = ("... spk time (mask freq) -> ... spk mask time freq" if ts_vad is False
                       else "... 1 time (spk mask freq) -> ... spk mask time freq")
        elif output_resolution == "t":
            assert explicit_vad is False, explicit_vad
            out_features = nmask * ts_factor
            pattern = ("... spk time mask -> ... spk mask time freq" if ts_vad is False
                       else "... 1 time (spk mask) -> ... spk mask time freq")
        else:
            raise ValueError(output_resolution)
        post.add("linear", torch.nn.Linear(in_features=projs, out_features=out_features))
        post.add("rearrange", _Einop(pattern, mask=nmask))
        self.post_net = Sequential(post.data)
        self.final_activation = torch.nn.Sigmoid()
        self._head_cache = None
        self._rot_cache = None
        self._staging = _PinnedRing()

    def extra_repr(self) -> str:
        return f"combination={self.combination!r},"

    def invalidate_caches(self):
        """Drops every packed weight copy.  The caches notice ``load_state_dict`` / ``.to()`` / optimizer steps on their
        own (``rnnp.param_key``); in-place edits through ``.data`` are the one case that needs this call."""
        self._head_cache = self._rot_cache = None
        for m in self.modules():
            if isinstance(m, RNNP_packed):
                m.invalidate_caches()

    # -- derived weight caches -------------------------------------------------
    def _birnns(self):
        return [m for n, m in self.post_net.named_children() if n.startswith("birnn")]

    def _head_linear(self):
        return [m for n, m in self.post_net.named_children() if n.startswith("linear")][0]

    def _head_pack(self, K, R, row_len, row_len_padded=None):
        """Head weights for the fused GEMM; ``row_len_padded`` > ``row_len`` inserts zero rows behind every block of
        ``row_len`` output columns (one block = the F bins of one (speaker, mask) plane), so that the columns of every
        plane start at a multiple of 8: the epilogue's 128-byte stores then cover whole 32-byte sectors of the padded
        output rows in EVERY plane, not only in plane 0."""
        lin = self._head_linear()
        row_len_padded = row_len if row_len_padded is None else row_len_padded
        key = (param_key(lin), K, R, row_len, row_len_padded)
        if self._head_cache is None or self._head_cache[0] != key:
            with torch.no_grad():
                P = lin.in_features
                W = lin.weight.detach().float()
                b = lin.bias.detach().float()
                if self.ts_vad is not False:
                    Wv = W.view(K, -1, P)
                    bv = b.view(K, -1)
                    q = torch.arange(K, device=W.device)
                    Wr = torch.cat([Wv[(q - r) % K] for r in range(R)], dim=-1)  # (K, blk, R*P)
                    W2 = Wr.reshape(-1, R * P)
                    b2 = torch.stack([bv[(q - r) % K] for r in range(R)], 0).mean(0).reshape(-1)
                else:
                    W2, b2 = W, b
                if row_len_padded != row_len:
                    pad = row_len_padded - row_len
                    W2 = torch.nn.functional.pad(W2.reshape(-1, row_len, W2.shape[1]), (0, 0, 0, pad)).reshape(-1, W2.shape[1])
                    b2 = torch.nn.functional.pad(b2.reshape(-1, row_len), (0, pad)).reshape(-1)
                pack = {"w": ops.cast_bf16(W2.contiguous()), "ld": ops.operand_ld(W2.shape[1]),
                        "b": b2.contiguous(), "kdim": W2.shape[1], "n": W2.shape[0]}
            self._head_cache = (key, pack)
        return self._head_cache[1]

    def _rotated_input_weights(self, pk, K, R):
        """R column-rotated copies of the speaker-concat layer's W_ih, stacked on rows."""
        key = (param_key(self._birnns()[-1]), K, R)
        if self._rot_cache is None or self._rot_cache[0] != key:
            with torch.no_grad():
                P = pk.I // K
                Wv = pk.w_ih_f32.view(8 * pk.Up, K, P)
                q = torch.arange(K, device=Wv.device)
                Wr = torch.stack([Wv[:, (q - r) % K] for r in range(R)], 0).reshape(R * 8 * pk.Up, K * P)
                ld = ops.operand_ld(K * P)
                self._rot_cache = (key, {"w": ops.cast_bf16(Wr.contiguous(), ld), "ld": ld})
        return self._rot_cache[1]

    # -- forward ---------------------------------------------------------------
    def forward(self, xs, aux=None, _features_bf16=None) -> Output:
        """xs (T, idim) or (B, T, idim) float32; aux list of K (A,) tensors / list of B (K, A).

        ``_features_bf16=(tensor, ld)`` lets ``Model.forward`` hand over the bf16 feature rows the
        feature kernel already produced.
        """
        if self.training and torch.is_grad_enabled():
            return self.forward_train(xs, aux)
        (_, _, out), = self.forward_waves(xs, aux, wave=None, _features_bf16=_features_bf16)
        return out

    def forward_train(self, xs, aux) -> Output:
        """Differentiable forward for the training step (BASELINE config 5; tssep/train/net.py:809-986 followed line by
        line, autograd through it).  The projected BLSTM layers run on this package's kernels forward AND backward
        (``tssep_b200.autograd.RNNPLayerFn``: tcgen05 GEMMs, tensor-memory recurrence, BPTT kernel); the speaker
        conditioning, the rearrangements, the head ``Linear`` and the sigmoid are torch ops (elementwise / plain
        library GEMM).  birnn0 / birnn1 treat every (item, speaker) row on its own, so the permutation-averaging
        trials are expanded AFTER them instead of before (identical values, 1/R of the work)."""
        _lib.require_cuda(xs)
        if xs.dim() == 2:
            batched = False
            K = len(aux)
            perm = np.random.permutation(K) if self.random_speaker_order else np.arange(K)
            aux_t = torch.stack([aux[i] for i in perm], dim=0)[None]
            perms = perm[None]
            xs = xs[None]
        elif xs.dim() == 3:
            batched = True
            K = len(aux[0])
            perms = np.stack([np.random.permutation(K) if self.random_speaker_order else np.arange(K)
                              for _ in range(len(aux))])
            aux_t = torch.stack([torch.stack([a[i] for i in p], dim=0) for a, p in zip(aux, perms)], dim=0)
            if self.aux_normalizer is not None:
                aux_t = self.aux_normalizer(aux_t)
        else:
            raise RuntimeError(xs.shape)
        B, T = xs.shape[:2]
        if self.input_normalizer is not None:
            xs = self.input_normalizer(xs)
        xs = self.pre_net(xs.float()) if isinstance(self.pre_net, RNNP_packed) else xs.float()   # (B, T, F)
        emb = aux_t.float().unsqueeze(-2)                                                         # (B, K, 1, A)
        if self.combination == "mul":
            h = xs[:, None] * emb                                                                 # (B, K, T, F)
        else:
            h = torch.cat([xs[:, None].expand(B, K, T, xs.shape[-1]), emb.expand(B, K, T, emb.shape[-1])], dim=-1)
        birnns = self._birnns()
        L, R = self.layers, self.num_averaged_permutations
        tsv = self.ts_vad is not False
        n_indep = L - 1 if tsv else L
        for l in range(n_indep):
            h = birnns[l](h)
            if l < L - 1:
                h = torch.tanh(h)
        if tsv:
            idx = ((np.arange(K)[:, None] + np.arange(K)[None, :]) % K)[:R].ravel()              # net.py:913-916
            he = h[:, torch.as_tensor(idx, device=h.device)].reshape(B * R, K, T, h.shape[-1])
            he = he.movedim(1, 2).reshape(B * R, T, K * h.shape[-1])                              # '(spk feature)'
            h = birnns[L - 1](he)                                                                 # (B*R, T, P)
        logit = self._head_linear()(h)
        nmask = self.nmask
        if self.output_resolution == "tf":
            fh = self.odim + int(self.explicit_vad)
            if tsv:
                logit = logit.reshape(B * R, T, K, nmask, fh).permute(0, 2, 3, 1, 4)              # spk mask time freq
            else:
                logit = logit.reshape(B, K, T, nmask, fh).permute(0, 1, 3, 2, 4)
        else:
            if tsv:
                logit = logit.reshape(B * R, T, K, nmask).permute(0, 2, 3, 1)
            else:
                logit = logit.reshape(B, K, T, nmask).permute(0, 1, 3, 2)
            logit = logit[..., None].expand(*logit.shape, self.odim)
        if tsv and R > 1:                                                                         # net.py:928-955
            revert = torch.as_tensor(np.argsort(idx), device=logit.device)
            logit = logit.reshape(B, R * K, *logit.shape[2:])[:, revert]
            logit = logit.reshape(B, K, R, *logit.shape[2:]).mean(dim=2)
        iperm = torch.as_tensor(np.argsort(perms, axis=-1), device=logit.device)
        logit = logit[torch.arange(B, device=logit.device)[:, None], iperm]                      # net.py:957-967
        embedding = emb
        if not batched:
            logit, embedding = logit[0], embedding[0]
        if self.explicit_vad:
            mask = torch.sigmoid(logit)
            vad = mask[..., 0]
            return Output(mask=mask[..., 1:] * vad[..., None], logit=None, vad_mask=vad, vad_logit=logit[..., 0],
                          embedding=embedding)
        return Output(mask=torch.sigmoid(logit), logit=logit, embedding=embedding)

    def forward_waves(self, xs, aux=None, wave=None, _features_bf16=None, out_wave=None):
        """Same computation as ``forward`` for a batch of B items, produced in waves of ``wave`` items:
        a generator of ``(lo, hi, Output)`` for items [lo, hi).

        The recurrences cost T dependent steps whatever their batch; a step advances up to
        ``ops.recurrence_ts_capacity`` rows at the same latency.  pre_net and the TS-VAD layer (1 and R rows per
        item) run ONCE for all B items, the K-rows-per-item layers per ``wave`` items (their input projections G are
        the largest buffer of the path: 0.36 GB per 10-min meeting, speaker and layer), and the head writes each output
        wave's logit / mask when the consumer asks for it, so the big outputs of one wave can be dropped before the
        next is produced.  ``out_wave`` (default ``wave``)
        is the number of items per yielded ``Output``: smaller output waves bound the memory of the
        GB-sized logit / mask tensors without touching the recurrence batching.  Results are independent
        of ``wave`` and ``out_wave``; speaker permutations are drawn for all items up front, in order.
        """
        _lib.require_cuda(xs)
        dev = xs.device
        if xs.dim() == 2:
            batched, B = False, 1
            K = len(aux)
            perms = [np.random.permutation(K) if self.random_speaker_order else np.arange(K)]
            aux_t = torch.stack([a for a in aux], dim=0)[None]
        elif xs.dim() == 3:
            batched, B = True, xs.shape[0]
            K = len(aux[0])
            perms = [np.random.permutation(K) if self.random_speaker_order else np.arange(K) for _ in range(len(aux))]
            aux_t = torch.stack([torch.stack([x for x in a], 0) if isinstance(a, (tuple, list)) else a for a in aux], 0)
        else:
            raise RuntimeError(xs.shape)
        _lib.require_cuda(aux_t)
        # Both index tables go to the device NOW, in one asynchronous copy from a persistent pinned buffer: a
        # pageable copy makes torch synchronise the stream (the host could not enqueue the next batch while
        # this one runs), and a fresh pin_memory() may stall in cudaHostAlloc.
        perm_np = np.stack(perms)  # (B, K)
        nmask = self.nmask
        if self.ts_vad is not False:
            planes = (np.arange(B)[:, None, None] * K + perm_np[:, :, None]) * nmask + np.arange(nmask)[None, None, :]
        else:
            planes = ((np.arange(B)[:, None] * K + perm_np)[:, :, None]) * nmask + np.arange(nmask)[None, None, :]
        packed = np.concatenate([perm_np.reshape(-1), planes.reshape(-1)]).astype(np.int32)
        packed_t = self._staging.upload(packed, dev)
        perm_t = packed_t[:B * K].view(B, K).long()
        plane_map = packed_t[B * K:]
        aux_p = torch.gather(aux_t.float(), 1, perm_t[:, :, None].expand(B, K, aux_t.shape[-1]))  # slot order
        if batched and self.aux_normalizer is not None:
            aux_p = self.aux_normalizer(aux_p)
        A = aux_p.shape[-1]
        T, Din = xs.shape[-2:]
        R = self.num_averaged_permutations
        L = self.layers
        if self.ts_vad is not False:
            assert self.ts_vad == K, (self.ts_vad, K)
            if L < 2:
                raise NotImplementedError("ts_vad with layers == 1 is not implemented")

        # features -> bf16 rows
        if self.input_normalizer is not None:
            xs = self.input_normalizer(xs)
            _features_bf16 = None
        if _features_bf16 is not None:
            xb, ld = _features_bf16
        else:
            xb = ops.cast_bf16(xs.reshape(B * T, Din).float())
            ld = ops.operand_ld(Din)

        del xs, _features_bf16  # only the bf16 rows are used from here on

        # pre_net on the mixture (net.py:860)
        if isinstance(self.pre_net, RNNP_packed):
            for pk in self.pre_net.layer_packs():
                G = pk.input_gemm(xb, ld, B * T)
                H = pk.recurrence(G, B, T)
                del G
                ld = ops.operand_ld(pk.hdim)
                xb = torch.empty((B * T, ld), dtype=torch.bfloat16, device=dev)
                pk.projection(H, B * T, xb, mode=ops.EPI_BF16, ldo=ld, act=0)
                del H
                F = pk.hdim
        else:
            F = Din

        birnns = self._birnns()
        packs = [m.layer_packs()[0] for m in birnns]
        pk0 = packs[0]
        Up = pk0.Up
        e = aux_p.reshape(B * K, A).contiguous()
        P = self.projs
        ldp = ops.operand_ld(P)
        if self.combination == "mul":
            assert A == F, ("combination='mul' needs aux_size == odim", A, F)
        else:
            assert pk0.I == F + A, (pk0.I, F, A)
        tsv = self.ts_vad is not False
        n_indep = L - 1 if tsv else L  # layers that treat every (item, speaker) row on its own
        # `wave`: items per pass of the speaker-independent layers -- one number, or the list of wave sizes
        if wave is None:
            wave_sizes = [B]
        elif isinstance(wave, (list, tuple)):
            wave_sizes = [int(w) for w in wave]
            assert sum(wave_sizes) == B and min(wave_sizes) >= 1, (wave_sizes, B)
        else:
            w = max(1, min(int(wave), B))
            wave_sizes = [min(w, B - lo) for lo in range(0, B, w)]
        gd = ops.g_dtype()
        gmode = ops.EPI_BF16 if gd == torch.bfloat16 else ops.EPI_F32
        # output of the speaker-independent layers for ALL items: speaker-concat rows (B, T, K*P) for the TS-VAD layer
        # (net.py:606-612), plain (B*K, T, P) rows otherwise
        y_ld = ops.operand_ld(K * P) if tsv else ldp
        y = torch.empty((B * T if tsv else B * K * T, y_ld), dtype=torch.bfloat16, device=dev)

        # ---- speaker-independent layers, `wave` items at a time (their G buffer is the largest of the path) ----
        lo = 0
        for Bw in wave_sizes:
            hi = lo + Bw
            rows = Bw * K
            # birnn0 with the conditioning folded into its input projection (net.py:862-896): 'mul' scales the weight
            # columns per speaker, 'cat' turns the embedding half of the weight into a per-speaker bias
            bias_k = torch.empty((rows, 8 * Up), dtype=torch.float32, device=dev)
            G = torch.empty((rows * T, 8 * Up), dtype=gd, device=dev)
            e_w = e[lo * K:hi * K]
            x_w = xb[lo * T:]
            if self.combination == "mul":
                ldk = ops.operand_ld(F)
                Wk = torch.empty((rows * 8 * Up, ldk), dtype=torch.bfloat16, device=dev)
                torch_ops.op.fold_embedding(0, pk0.w_ih_f32, pk0.I, pk0.bias, e_w, rows, 8 * Up, F, A, Wk, ldk, bias_k)
                ops.gemm(x_w, ld, Wk, ldk, T, 8 * Up, F, G, mode=gmode, ldo=8 * Up, batch=rows,
                         a_stride=T * ld, a_div=K, b_stride=8 * Up * ldk, bias=bias_k, bias_stride=8 * Up,
                         out_stride=T * 8 * Up)
                del Wk
            else:  # cat
                torch_ops.op.fold_embedding(1, pk0.w_ih_f32, pk0.I, pk0.bias, e_w, rows, 8 * Up, F, A, None, 0, bias_k)
                ops.gemm(x_w, ld, pk0.w_ih, pk0.ld_in, T, 8 * Up, F, G, mode=gmode, ldo=8 * Up, batch=rows,
                         a_stride=T * ld, a_div=K, b_stride=0, bias=bias_k, bias_stride=8 * Up,
                         out_stride=T * 8 * Up)
            yw, yw_ld = None, None
            for l in range(n_indep):
                pk = packs[l]
                if l > 0:
                    G = pk.input_gemm(yw, yw_ld, rows * T)
                    del yw
                H = pk.recurrence(G, rows, T)
                del G
                if l < n_indep - 1:
                    yw_ld = ldp
                    yw = torch.empty((rows * T, yw_ld), dtype=torch.bfloat16, device=dev)
                    pk.projection(H, rows * T, yw, mode=ops.EPI_BF16, ldo=yw_ld, act=1)
                elif tsv:
                    # tanh(proj) straight into the speaker-concat layout (B, T, K*P)   net.py:606-612
                    pk.projection(H, 0, y[lo * T:], mode=ops.EPI_BF16, ldo=y_ld, act=1, batch=rows,
                                  a_stride=T * 2 * pk.Up, M=T, out_stride=P, out_div=K, out_stride_hi=T * y_ld)
                else:
                    pk.projection(H, rows * T, y[lo * K * T:], mode=ops.EPI_BF16, ldo=y_ld, act=0)
                del H
            lo = hi
        del xb

        # ---- TS-VAD layer: all speakers of an item in one row, R cyclic speaker orders as R rotated weight copies ----
        if tsv:
            pk = packs[L - 1]
            rot = self._rotated_input_weights(pk, K, R)
            rows = B * R
            G = torch.empty((rows * T, 8 * pk.Up), dtype=gd, device=dev)
            ops.gemm(y, y_ld, rot["w"], rot["ld"], T, 8 * pk.Up, K * P, G, mode=gmode, ldo=8 * pk.Up,
                     batch=rows, a_stride=T * y_ld, a_div=R, b_stride=8 * pk.Up * rot["ld"], b_mod=R,
                     bias=pk.bias, bias_stride=0, out_stride=T * 8 * pk.Up)
            H = pk.recurrence(G, rows, T)
            del G
            # trial-concat layout (B, T, R*P) feeding the averaged head
            y_ld = ops.operand_ld(R * P)
            y = torch.empty((B * T, y_ld), dtype=torch.bfloat16, device=dev)
            pk.projection(H, 0, y, mode=ops.EPI_BF16, ldo=y_ld, act=0, batch=rows, a_stride=T * 2 * pk.Up, M=T,
                          out_stride=P, out_div=R, out_stride_hi=T * y_ld)
            del H

        # head: Linear + rearrange + trial mean + un-permute + sigmoid (net.py:629-668, :928-986), per wave
        nmask, odim = self.nmask, self.odim
        fh = odim + int(self.explicit_vad)
        tf = self.output_resolution == "tf"
        row_len = fh if tf else 1
        pitch = ops.round_up(fh, 8) if tf else row_len   # padded row length of the head outputs (tf resolution)
        hp = self._head_pack(K, R, row_len, pitch)
        per_item, nb = (1, K * nmask) if self.ts_vad is not False else (K, nmask)  # GEMM batch entries / planes per item
        embedding_all = aux_p.unsqueeze(-2)
        out_wave = max(wave_sizes) if out_wave is None else max(1, min(int(out_wave), B))
        for lo in range(0, B, out_wave):
            hi = min(B, lo + out_wave)
            Bw = hi - lo
            items = Bw * per_item
            y_w = y[lo * per_item * T:]
            pm = plane_map if Bw == B else plane_map[lo * K * nmask:hi * K * nmask] - lo * K * nmask
            shape = (Bw, K, nmask, T, fh if tf else odim)
            if tf:
                # rows of fh = 513 floats at a pitch of 520 (7 zero-weight columns per plane): every row of every plane
                # starts on a 32-byte sector, so the 128-byte stores of the head epilogue cover whole sectors
                # (1.9 -> 3.5 TB/s of output); the tensors handed out are the (..., :fh) views of the padded buffers,
                # which the mask consumers (iSTFT, activity) read in place
                logit = torch.empty((*shape[:-1], pitch), dtype=torch.float32, device=dev)[..., :fh]
                mask = torch.empty((*shape[:-1], pitch), dtype=torch.float32, device=dev)[..., :fh]
                ops.gemm(y_w, y_ld, hp["w"], hp["ld"], T, hp["n"], hp["kdim"], logit, mode=ops.EPI_HEAD, batch=items,
                         a_stride=T * y_ld, b_stride=0, b_mod=1, bias=hp["b"], alpha=1.0 / R, mask=mask,
                         plane_map=pm, n_blocks=nb, row_len=pitch)
            else:
                logit = torch.empty(shape, dtype=torch.float32, device=dev)
                mask = torch.empty(shape, dtype=torch.float32, device=dev)
                small = torch.empty((items * T, nb), dtype=torch.float32, device=dev)
                ops.gemm(y_w, y_ld, hp["w"], hp["ld"], items * T, hp["n"], hp["kdim"], small, mode=ops.EPI_F32, ldo=nb,
                         b_mod=1, bias=hp["b"], alpha=1.0 / R)
                torch_ops.op.head_expand_t(small, items, T, nb, odim, pm, logit, mask)
                del small
            embedding = embedding_all[lo:hi]
            if hi == B:
                del y, y_w
            if not batched:
                logit, mask, embedding = logit[0], mask[0], embedding[0]
            if self.explicit_vad:
                vad_mask = mask[..., 0]
                out = Output(mask=mask[..., 1:] * vad_mask[..., None], logit=None, vad_mask=vad_mask,
                             vad_logit=logit[..., 0], embedding=embedding)
            else:
                out = Output(mask=mask, logit=logit, embedding=embedding)
            del logit, mask
            yield lo, hi, out
            del out
