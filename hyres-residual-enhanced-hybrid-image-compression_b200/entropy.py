"""Entropy models of the product path: parameter / CDF-table holders with the reference's
names (``entropy_bottleneck.*``, ``gaussian_conditional.*`` state-dict keys) whose compute
runs in the CUDA kernels of csrc/entropy.cu and whose coding runs in csrc/rans.cpp.

Replaces, for this path, compressai 1.2.6's ``EntropyBottleneck`` / ``GaussianConditional``
as used at models/checkerboard.py:30-31,96-101,140-142,159-165,172-173,206,261-276.
Table building (``update``) is a one-off host step: the pmf is evaluated with torch on the
CPU exactly as the reference's ``update()`` does when run from src/updata.py, and quantised
by ``hyres_pmf_to_quantized_cdf``.
"""
import math

import numpy as np
import threading

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import coder, ops


class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (grad_output < 0)) * grad_output, None


class LowerBound(nn.Module):
    """max(x, bound) whose gradient also passes when it would raise x (compressai.ops)."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9, entropy_coder_precision=16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.likelihood_bound = float(likelihood_bound)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._tables_cache = None

    # -- CDF tables --
    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            c = coder.pmf_to_quantized_cdf(prob.detach().cpu().numpy(), self.entropy_coder_precision)
            cdf[i, : c.size] = torch.from_numpy(c)
        return cdf

    def _check_tables(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if self._quantized_cdf.dim() != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if self._offset.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if self._cdf_length.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def tables(self):
        """Host copy of the tables for the coder (cached until the buffers change)."""
        self._check_tables()
        key = (self._quantized_cdf.data_ptr(), self._quantized_cdf._version, tuple(self._quantized_cdf.shape))
        if self._tables_cache is None or self._tables_cache[0] != key:
            t = coder.CdfTables(self._quantized_cdf.cpu().numpy(), self._cdf_length.cpu().numpy(),
                                self._offset.cpu().numpy())
            self._tables_cache = (key, t)
        return self._tables_cache[1]

    @staticmethod
    def _device_key(device):
        d = torch.device(device)
        if d.type == "cuda" and d.index is None:
            d = torch.device("cuda", torch.cuda.current_device())
        return str(d)

    def coder_rows(self, device):
        """int32 [3, n_cdfs] on ``device``: the host coder's packed-table layout (``coder.table_layout``), what the
        device front-end needs to emit coder slots / codes (cached with the tables)."""
        t = self.tables()
        cache = getattr(self, "_rows_cache", None)
        # (the cache holds the table object itself: an id() alone could be reused by the tables of a later update())
        key = self._device_key(device)
        if cache is None or cache[0] is not t or cache[1] != key:
            rows = torch.from_numpy(coder.table_layout(t)).to(device)
            self._rows_cache = cache = (t, key, rows)
        return cache[2]

    def device_tables(self, device):
        """The coder's packed tables on ``device`` for the device-resident coder (``coder.DeviceTables``, cached with
        the host tables)."""
        t = self.tables()
        cache = getattr(self, "_dev_tables_cache", None)
        key = self._device_key(device)
        if cache is None or cache[0] is not t or cache[1] != key:
            self._dev_tables_cache = cache = (t, key, coder.DeviceTables(t, device))
        return cache[2]

    # -- coding of already-quantised symbols --
    _tls = threading.local()  # pinned staging buffers are per thread: CodecPipeline runs one batch per worker thread

    @classmethod
    def _host_i32(cls, t, slot):
        """int32 [B, n] numpy view of a tensor; CUDA tensors come back through a cached pinned buffer (a pageable
        ``.cpu()`` of the 35 MB symbol tensors of a 2048x1408 image runs at a fraction of the PCIe rate)."""
        B = t.size(0)
        t = t.reshape(B, -1)
        if t.dtype != torch.int32:
            t = t.to(torch.int32)
        if not t.is_cuda:
            return t.contiguous().numpy()
        key = (slot, t.numel())
        pinned = cls._tls.__dict__.setdefault("pinned", {})
        buf = pinned.get(key)
        if buf is None:
            if len(pinned) > 16:
                pinned.clear()
            buf = pinned[key] = torch.empty(t.numel(), dtype=torch.int32).pin_memory()
        view = buf.view(B, -1)
        view.copy_(t, non_blocking=True)
        ops.stream_wait_blocking(t.device)
        return view.numpy()

    def encode_symbols(self, symbols, indexes):
        """symbols / indexes: int32 tensors [B, ...] (any device) -> list of B byte strings."""
        return self.encode_symbol_groups([(symbols, indexes)])[0]

    def encode_symbol_groups(self, groups, slots=False):
        """groups: list of (symbols, indexes) pairs of equal shape -> list (per group) of lists of B byte strings.
        All strings of all groups are coded in ONE batch, so e.g. the anchor and non-anchor passes of a batch of
        8 tiles keep 16 host threads busy instead of 8 twice.  ``slots=True``: the second tensor of every pair holds
        coder slots from the device front-end (``ops.gc_symbols(..., rows=...)``) instead of CDF indexes; same bytes."""
        syms, idxs = [], []
        for g, (symbols, indexes) in enumerate(groups):
            if symbols.dim() < 2:
                raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
            if symbols.size() != indexes.size():
                raise ValueError("`inputs` and `indexes` should have the same size.")
            syms.append(self._host_i32(symbols, ("s", g)))
            idxs.append(self._host_i32(indexes, ("i", g)))
        out = coder.encode_batch(syms, idxs, self.tables(), slots=slots)
        res, k = [], 0
        for a in syms:
            res.append(out[k:k + a.shape[0]])
            k += a.shape[0]
        return res

    @classmethod
    def _pinned_i32(cls, slot, n):
        pinned = cls._tls.__dict__.setdefault("pinned", {})
        key = (slot, n)
        buf = pinned.get(key)
        if buf is None:
            if len(pinned) > 16:
                pinned.clear()
            buf = pinned[key] = torch.empty(n, dtype=torch.int32).pin_memory()
        return buf

    def decode_symbols(self, strings, indexes, slot=None, codes=False):
        """-> int32 CPU tensor shaped like ``indexes`` (pinned when the indexes came from the GPU).  ``slot``: decode
        into this thread's cached pinned buffer of that name instead of a fresh allocation (pinning 35 MB per pass
        costs milliseconds); the result is then only valid until the same thread decodes into the slot again.
        ``codes=True``: ``indexes`` holds decoder codes (``ops.gc_codes``); the entries of known symbols are then
        unspecified in the result (``ops.gc_dequant(..., pass_id)`` does not read them)."""
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if len(strings) != indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        ix = self._host_i32(indexes, ("d", 0))
        if slot is not None and indexes.is_cuda:
            buf = self._pinned_i32(("o", slot), ix.size)
            coder.decode_batch(list(strings), ix, self.tables(), out=buf.numpy().reshape(ix.shape), codes=codes)
            return buf.view(indexes.shape)
        out = coder.decode_batch(list(strings), ix, self.tables(), codes=codes,
                                 out=np.zeros(ix.shape, dtype=np.int32) if codes else None)
        res = torch.from_numpy(out).reshape(indexes.shape)
        return res.pin_memory() if indexes.is_cuda else res


class EntropyBottleneck(EntropyModel):
    """Factorised prior (filters (3,3,3,3)); parameters named as in compressai 1.2.x."""

    def __init__(self, channels, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(**kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise ValueError("the B200 EntropyBottleneck kernel is specialised for filters=(3,3,3,3)")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        self.matrices, self.biases, self.factors = nn.ParameterList(), nn.ParameterList(), nn.ParameterList()
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            self.matrices.append(nn.Parameter(torch.full((self.channels, filters[i + 1], filters[i]), float(init))))
            self.biases.append(nn.Parameter(torch.empty(self.channels, filters[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.factors.append(nn.Parameter(torch.zeros(self.channels, filters[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(self.channels, 1, 1))
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, inputs, stop_gradient):
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = self.matrices[i].detach() if stop_gradient else self.matrices[i]
            logits = torch.matmul(F.softplus(matrix), logits)
            bias = self.biases[i].detach() if stop_gradient else self.biases[i]
            logits = logits + bias
            if i < len(self.filters):
                factor = self.factors[i].detach() if stop_gradient else self.factors[i]
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def loss(self):
        """aux loss on the quantiles (tiny host-side torch graph; keeps autograd for the aux optimiser)."""
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def update(self, force=False, update_quantiles=False):
        if self._offset.numel() > 0 and not force:
            return False
        dev = self.quantiles.device
        with torch.no_grad():
            qt = self.quantiles.detach().cpu()
            medians = qt[:, 0, 1]
            minima = torch.clamp(torch.ceil(medians - qt[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(qt[:, 0, 2] - medians).int(), min=0)
            pmf_start = medians - minima
            pmf_length = maxima + minima + 1
            max_length = int(pmf_length.max().item())
            samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
            cpu = {k: [p.detach().cpu() for p in getattr(self, k)] for k in ("matrices", "biases", "factors")}

            def cum(v):
                for i in range(5):
                    v = torch.matmul(F.softplus(cpu["matrices"][i]), v) + cpu["biases"][i]
                    if i < 4:
                        v = v + torch.tanh(cpu["factors"][i]) * torch.tanh(v)
                return v

            lower, upper = cum(samples - 0.5), cum(samples + 0.5)
            pmf = (torch.sigmoid(upper) - torch.sigmoid(lower))[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
            self._offset = (-minima).to(dev)
            self._cdf_length = (pmf_length + 2).to(dev)
        self._tables_cache = None
        return True

    def kernel_params(self):
        """[C,58] fp32: softplus(matrices) | biases | tanh(factors), the layout of csrc/entropy.cu."""
        with torch.no_grad():
            parts = [F.softplus(m).reshape(self.channels, -1) for m in self.matrices]
            parts += [b.reshape(self.channels, -1) for b in self.biases]
            parts += [torch.tanh(f).reshape(self.channels, -1) for f in self.factors]
            return torch.cat(parts, dim=1).float().contiguous()

    def medians_vector(self):
        return self.quantiles.detach()[:, 0, 1].float().contiguous()

    def forward(self, x, training=None):
        """(B,C,H,W) fp32 CUDA -> (outputs, likelihood), both (B,C,H,W)."""
        if training is None:
            training = self.training
        if not x.is_cuda:
            raise RuntimeError("hyres_b200 runs on a CUDA sm_100 device only (no CPU path)")
        z = x.permute(0, 2, 3, 1).contiguous().float()
        r = ops.eb_forward(z, self.kernel_params(), self.medians_vector(), lik_noise=training, out_noise=training,
                           seed=int(torch.randint(0, 2 ** 62, (1,)).item()), lik_bound=self.likelihood_bound,
                           want_zhat_nchw=True)
        return r["zhat_nchw"], r["lik"]

    @staticmethod
    def _build_indexes(size, device="cpu"):
        N, Cc = size[0], size[1]
        view = [1] * len(size)
        view[1] = -1
        return torch.arange(Cc, dtype=torch.int32, device=device).view(*view).repeat(N, 1, *size[2:])

    def device_indexes(self, size, device):
        """``_build_indexes(size)`` as a cached CUDA tensor (the device coder reads the CDF row of every symbol)."""
        key = (tuple(int(v) for v in size), str(device))
        cache = self.__dict__.setdefault("_dev_index_cache", {})
        ix = cache.get(key)
        if ix is None:
            if len(cache) > 8:
                cache.clear()
            ix = cache[key] = self._build_indexes(size, device=device).contiguous()
        return ix

    def compress(self, x):
        """(B,C,H,W) -> list of strings (symbols = round(x - median))."""
        med = self.quantiles.detach()[:, 0, 1].reshape(1, -1, *([1] * (x.dim() - 2))).to(x.device)
        symbols = torch.round(x - med).int()
        return self.encode_symbols(symbols, self._build_indexes(x.size()))

    def decompress(self, strings, size):
        out_size = (len(strings), self._quantized_cdf.size(0), *size)
        sym = self.decode_symbols(strings, self._build_indexes(out_size))
        dev = self.quantiles.device
        med = self.quantiles.detach()[:, 0, 1].reshape(1, -1, *([1] * len(size)))
        return sym.to(dev).to(med.dtype) + med


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(**kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_cumulative(inputs):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        """Inverse normal CDF in float64 (the reference calls scipy.stats.norm.ppf): Newton
        iterations on Phi(x) = 0.5*erfc(-x/sqrt(2)) from an erfinv start."""
        qv = float(quantile)
        x = float(math.sqrt(2.0) * torch.erfinv(torch.tensor(2.0 * qv - 1.0, dtype=torch.float64)).item())
        if not math.isfinite(x):
            x = -6.0 if qv < 0.5 else 6.0
        for _ in range(8):
            cdf = 0.5 * math.erfc(-x / math.sqrt(2.0))
            pdf = math.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
            x -= (cdf - qv) / pdf
        return x

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        dev = self.scale_table.device
        with torch.no_grad():
            table = self.scale_table.detach().cpu()
            multiplier = -self._standardized_quantile(self.tail_mass / 2)
            pmf_center = torch.ceil(table * multiplier).int()
            pmf_length = 2 * pmf_center + 1
            max_length = int(torch.max(pmf_length).item())
            samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
            samples_scale = table.unsqueeze(1).float()
            upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
            lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
            pmf = upper - lower
            tail_mass = 2 * lower[:, :1]
            self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
            self._offset = (-pmf_center).to(dev)
            self._cdf_length = (pmf_length + 2).to(dev)
        self._tables_cache = None
