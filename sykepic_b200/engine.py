"""Per-GPU inference engine: the Python host of the C-ABI library.

One `Engine` owns one `spk_ctx` (one GPU, one CUDA stream).  PyTorch is used only
to read the checkpoint (`torch.load` of the reference's `best_state.pth`
state_dict) and to own pinned-host / device buffers; every computation is a
call into libsykepic_b200.so.

Reference behaviour mirrored here:
  * model directory layout and `config.ini` keys: `prepare_model`
    sykepic/compute/probability.py:118-130, `get_img_shape` / `get_transforms` /
    `get_network` sykepic/train/config.py:20-77;
  * network structure: `TorchVisionNet` sykepic/train/network.py:11-72 over
    torchvision ResNet / DenseNet, described to the library from the state_dict keys;
  * per-bin flow: `process_sample` + `net_pass` probability.py:133-197.
"""

import ctypes as C
import os
import re
from configparser import ConfigParser, NoOptionError
from pathlib import Path

import numpy as np

from . import _lib
from ._lib import ptr

SOFTMAX_EXP = 1.3  # sykepic/compute/probability.py:18
BN_EPS = 1e-5  # torch.nn.BatchNorm2d default (torchvision resnet / densenet)


class ModelSpec:
    """What the reference's `prepare_model` reads from a model directory."""

    def __init__(self, classes, img_shape, border, arch, state_dict, imagenet_normalization=False):
        self.classes = list(classes)
        self.img_shape = tuple(img_shape)
        self.border = border
        self.arch = arch
        self.state_dict = state_dict
        # sykepic/train/config.py:55-56 appends Normalize to the TRAIN transform only; the eval
        # transform of the prob path never normalises.  Kept for information.
        self.imagenet_normalization = imagenet_normalization

    @classmethod
    def from_dir(cls, model_dir):
        import torch

        model_dir = Path(model_dir)
        with open(model_dir / "class_names.txt") as fh:
            classes = fh.read().splitlines()
        config = ConfigParser()
        config.read(model_dir / "config.ini")
        img_shape = tuple(int(i) for i in config.get("image", "shape").split(","))
        border = config.get("image", "border")
        if border not in _lib.BORDER:
            # image.py:20-28 silently leaves self.border unset for unknown names and fails later
            raise ValueError(f"unknown border {border!r} in {model_dir / 'config.ini'}")
        try:
            norm = config.getboolean("image", "imagenet_normalization")
        except (NoOptionError, ValueError):
            norm = False
        arch = config.get("model", "network")
        # the `weights` key only selects what torchvision downloads before load_state_dict
        # overwrites it (config.py:65-70); nothing is ever downloaded here.
        sd = torch.load(model_dir / "best_state.pth", map_location="cpu", weights_only=True)
        return cls(classes, img_shape, border, arch, sd, norm)


def _np(sd, key):
    t = sd[key]
    if hasattr(t, "detach"):
        t = t.detach().cpu().float().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


class _IdPool:
    def __init__(self, first=1):
        self.next = first
        self.free = []

    def get(self):
        if self.free:
            return self.free.pop()
        i = self.next
        self.next += 1
        return i

    def put(self, i):
        if i is not None and i not in self.free:
            self.free.append(i)


class Engine:
    """B200 engine for one model.  `precision`: "fp32" (CUDA-core FFMA convolutions, probabilities within 1e-6 of the
    reference), "bf16" (tcgen05 tensor-core convolutions, within 2e-2) or "fp32_tc" (fp32-level accuracy on the 16-bit
    tensor cores: activations and weights carried as fp16 hi + lo pairs = 22 significant bits, K accumulated in chunks that
    are summed in fp32 registers; ~6x the fp32 throughput, within 2e-5 of the reference on ResNet-18 / -50 / DenseNet-121;
    activations beyond |65504| raise ArithmeticError)."""

    def __init__(self, spec, device=0, precision="bf16", max_batch=256, conv_impl="auto", stream=None, pre_chunk=None):
        import torch

        if not isinstance(spec, ModelSpec):
            spec = ModelSpec.from_dir(spec)
        self.spec = spec
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SpkError(_lib.SPK_ERR_CUDA, "no CUDA device: sykepic_b200 has no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.precision = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16, "fp32_tc": _lib.PRECISION_FP32_TC}[precision]
        self.precision_name = precision
        self.conv_impl = {"auto": _lib.CONV_AUTO, "simt": _lib.CONV_SIMT, "tcgen05": _lib.CONV_TCGEN05, "taps": _lib.CONV_TCGEN05_TAPS}[conv_impl]
        self.max_batch = int(max_batch)
        c, th, tw = spec.img_shape
        if c != 3:
            # data.py:220-223 builds a [1,T,T] tensor that no torchvision backbone accepts unmodified
            raise ValueError(f"image shape {spec.img_shape}: only 3-channel models are supported")
        self.th, self.tw = th, tw
        self.k = len(spec.classes)
        with torch.cuda.device(self.device):
            self.stream = stream or torch.cuda.Stream(device=self.device)
        self.ctx = C.c_void_p()
        _lib.check(self.lib.spk_create(device, C.c_void_p(self.stream.cuda_stream), C.byref(self.ctx)))
        self._keep = []  # numpy arrays referenced by the library during graph construction
        self._channels = {}  # buffer id -> channels of the tensor it currently holds (graph construction)
        self._slice_reads = False  # DenseNet: convolutions read the first c channels of a wider concat buffer
        self._build_graph()
        self.softmax_scale = float(np.float32(np.log(SOFTMAX_EXP)))  # probability.py:192-193
        # K1 decodes / resizes a whole chunk of a bin per launch (it only reaches HBM speed on thousands of ROIs);
        # K2 + K3 then walk the chunk in batches.  4096 ROIs = 205 MB of u8 planes at T = 224.
        self.pre_chunk = max(self.max_batch, int(pre_chunk or os.environ.get("SYKEPIC_PRE_CHUNK", "4096")))
        self._x = torch.empty((self.pre_chunk, th, tw), dtype=torch.uint8, device=self.device)
        self._thr_dev = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.spk_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(rc, self.ctx)

    @property
    def launches(self):
        return int(self.lib.spk_launch_count(self.ctx))

    def net_bytes(self):
        return int(self.lib.spk_net_bytes(self.ctx))

    # ------------------------------------------------------------------ graph construction
    def _conv(self, sd, wkey, bnkey, inb, outb, stride, pad, relu, res=-1, in_off=0, out_off=0):
        w = _np(sd, wkey)
        cout, cin, kh, kw = w.shape
        have = self._channels.get(inb)
        if in_off == 0 and have is not None and inb != 0 and cin != have and not self._slice_reads:
            # a grouped convolution (weight [Cout, Cin/groups, k, k]) would otherwise be run as a dense one over the
            # first Cin/groups channels of its input
            raise ValueError(f"{wkey}: weight expects {cin} input channels, the tensor it reads has {have} "
                             "(grouped convolutions are not supported)")
        self._channels[outb] = max(self._channels.get(outb, 0), out_off + cout) if out_off else cout
        bn = [None] * 4
        if bnkey is not None:
            bn = [_np(sd, f"{bnkey}.{n}") for n in ("weight", "bias", "running_mean", "running_var")]
        self._keep.extend([w, *bn])
        self._ck(self.lib.spk_net_conv(self.ctx, inb, in_off, outb, out_off, res, ptr(w), cout, cin, kh, kw, stride, pad,
                                       ptr(bn[0]), ptr(bn[1]), ptr(bn[2]), ptr(bn[3]), BN_EPS, None, int(relu),
                                       self.conv_impl))
        return cout

    def _bn_relu(self, sd, bnkey, inb, outb, channels, relu=True):
        bn = [_np(sd, f"{bnkey}.{n}") for n in ("weight", "bias", "running_mean", "running_var")]
        self._keep.extend(bn)
        self._ck(self.lib.spk_net_bn_relu(self.ctx, inb, outb, channels, ptr(bn[0]), ptr(bn[1]), ptr(bn[2]), ptr(bn[3]),
                                          BN_EPS, int(relu)))

    # torchvision architectures whose children()[:-1] this engine reproduces (TorchVisionNet, network.py:48-55).  Grouped
    # or dilated variants (resnext*, wide_resnet* keeps plain convolutions and is fine) share the ResNet key layout but
    # not its arithmetic, so the architecture NAME from config.ini is checked, not only the state_dict keys.
    SUPPORTED_ARCH = re.compile(r"(resnet(18|34|50|101|152)|wide_resnet(50|101)_2|densenet(121|161|169|201))$")

    def _build_graph(self):
        sd = self.spec.state_dict
        if not self.SUPPORTED_ARCH.match(str(self.spec.arch)):
            raise ValueError(f"unsupported network {self.spec.arch!r}: this engine runs torchvision resnet18/34/50/101/152, "
                             "wide_resnet50_2/101_2 and densenet121/161/169/201")
        self._ck(self.lib.spk_net_begin(self.ctx, self.th, self.tw, 1, self.precision, self.max_batch))
        if "base.0.conv0.weight" in sd:
            feat_buf = self._build_densenet(sd)
        elif "base.0.weight" in sd and "base.4.0.conv1.weight" in sd:
            feat_buf = self._build_resnet(sd)
        else:
            raise ValueError(f"unsupported network {self.spec.arch!r}: state_dict is neither a torchvision ResNet nor DenseNet")
        # head: every head.<i>.weight in index order (Dropout entries have no parameters; network.py:56-63)
        idx = sorted(int(m.group(1)) for k in sd if (m := re.fullmatch(r"head\.(\d+)\.weight", k)))
        ws = [_np(sd, f"head.{i}.weight") for i in idx]
        bs = [_np(sd, f"head.{i}.bias") for i in idx]
        dims = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        if dims[-1] != self.k:
            raise ValueError(f"checkpoint has {dims[-1]} outputs, class_names.txt has {self.k} classes")
        n = len(ws)
        wp = (C.c_void_p * n)(*[ptr(w) for w in ws])
        bp = (C.c_void_p * n)(*[ptr(b) for b in bs])
        dm = (C.c_int * (n + 1))(*dims)
        self._ck(self.lib.spk_net_head(self.ctx, feat_buf, n, wp, bp, dm))
        self._ck(self.lib.spk_net_end(self.ctx))
        self._keep.clear()
        slow = int(self.lib.spk_net_simt_layers(self.ctx))
        if slow:
            from .utils import logger

            logger.get_logger("prob").warning(
                f"{slow} convolution layer(s) of {self.spec.arch!r} have no tensor-core kernel in precision "
                f"{self.precision_name!r} and run on CUDA cores (see stderr): expect a large slowdown")

    def _build_resnet(self, sd):
        """torchvision ResNet children()[:-1] (conv1, bn1, relu, maxpool, layer1-4, avgpool) from the
        `base.<i>` keys TorchVisionNet produces (network.py:48-62)."""
        ids = _IdPool()
        a = ids.get()
        self._conv(sd, "base.0.weight", "base.1", 0, a, 2, 3, True)
        x = ids.get()
        self._ck(self.lib.spk_net_maxpool(self.ctx, a, x, 3, 2, 1))
        self._channels[x] = self._channels[a]
        ids.put(a)
        for stage in (4, 5, 6, 7):
            blocks = sorted({int(m.group(1)) for k in sd if (m := re.match(rf"base\.{stage}\.(\d+)\.", k))})
            for b in blocks:
                p = f"base.{stage}.{b}"
                stride = 2 if (stage > 4 and b == 0) else 1
                bottleneck = f"{p}.conv3.weight" in sd
                identity, ds = x, None
                if f"{p}.downsample.0.weight" in sd:
                    ds = ids.get()
                    self._conv(sd, f"{p}.downsample.0.weight", f"{p}.downsample.1", x, ds, stride, 0, False)
                    identity = ds
                t1 = ids.get()
                if bottleneck:  # 1x1 -> 3x3 (stride here: torchvision "v1.5") -> 1x1
                    self._conv(sd, f"{p}.conv1.weight", f"{p}.bn1", x, t1, 1, 0, True)
                    t2 = ids.get()
                    self._conv(sd, f"{p}.conv2.weight", f"{p}.bn2", t1, t2, stride, 1, True)
                    out = ids.get()
                    self._conv(sd, f"{p}.conv3.weight", f"{p}.bn3", t2, out, 1, 0, True, res=identity)
                    ids.put(t2)
                else:
                    self._conv(sd, f"{p}.conv1.weight", f"{p}.bn1", x, t1, stride, 1, True)
                    out = ids.get()
                    self._conv(sd, f"{p}.conv2.weight", f"{p}.bn2", t1, out, 1, 1, True, res=identity)
                ids.put(t1)
                ids.put(ds)
                ids.put(x)
                x = out
        return x

    def _build_densenet(self, sd):
        """torchvision DenseNet forward: features -> ReLU -> global average pool, then the syke-pic
        head.  The reference's TorchVisionNet raises for DenseNet at these sizes;
        this is the defined behaviour ("reference-undefined; parity vs torchvision")."""
        p = "base.0."
        ids = _IdPool()
        self._slice_reads = True
        h = (self.th + 6 - 7) // 2 + 1
        w = (self.tw + 6 - 7) // 2 + 1
        a = ids.get()
        c = self._conv(sd, p + "conv0.weight", p + "norm0", 0, a, 2, 3, True)
        h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        blk = 1
        src, src_kind = a, "maxpool"
        t1, t2 = ids.get(), ids.get()
        while (p + f"denseblock{blk}.denselayer1.norm1.weight") in sd:
            nl = 1
            while (p + f"denseblock{blk}.denselayer{nl + 1}.norm1.weight") in sd:
                nl += 1
            growth = sd[p + f"denseblock{blk}.denselayer1.conv2.weight"].shape[0]
            total = c + nl * growth
            cat = ids.get()
            self._ck(self.lib.spk_net_buffer(self.ctx, cat, h, w, total))
            if src_kind == "maxpool":
                self._ck(self.lib.spk_net_maxpool(self.ctx, src, cat, 3, 2, 1))
            else:
                self._ck(self.lib.spk_net_avgpool(self.ctx, src, cat, 2, 2))
            for layer in range(1, nl + 1):
                q = p + f"denseblock{blk}.denselayer{layer}"
                self._bn_relu(sd, q + ".norm1", cat, t1, c)
                self._conv(sd, q + ".conv1.weight", q + ".norm2", t1, t2, 1, 0, True)
                self._conv(sd, q + ".conv2.weight", None, t2, cat, 1, 1, False, out_off=c)
                c += growth
            t = p + f"transition{blk}"
            if (t + ".norm.weight") not in sd:
                self._bn_relu(sd, p + "norm5", cat, t1, c)
                return t1
            # transition: BN -> ReLU -> 1x1 conv -> 2x2 average pool (into the next block's buffer)
            self._bn_relu(sd, t + ".norm", cat, t1, c)
            src, src_kind = ids.get(), "avgpool"
            c = self._conv(sd, t + ".conv.weight", None, t1, src, 1, 0, False)
            h, w = h // 2, w // 2
            blk += 1
        raise ValueError("DenseNet state_dict without a final dense block")

    # ------------------------------------------------------------------ thresholds
    def set_thresholds(self, thresholds):
        """Per-class thresholds for the fused label rule (sykepic/compute/prediction.py:49-71).

        `thresholds`: dict name -> float (classes without an entry can never be "classified"),
        a number (scalar rule: idxmax, classified iff p > thr) or None."""
        torch = self.torch
        if thresholds is None:
            self._thr_dev = None
            return
        if isinstance(thresholds, (int, float)):
            # scalar rule (prediction.py:57-59): (idxmax, p > thr) == "best class strictly above thr"
            q = np.full(self.k, int(self.lib.spk_threshold_quantize(float(thresholds), 1)), np.int32)
        else:
            q = np.full(self.k, _lib.INT32_MAX, np.int32)
            for i, name in enumerate(self.spec.classes):
                if name in thresholds:
                    q[i] = self.lib.spk_threshold_quantize(float(thresholds[name]), 0)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self._thr_dev = torch.from_numpy(q).to(self.device)
        self.stream.synchronize()

    # ------------------------------------------------------------------ device steps
    def preprocess(self, roi_dev, roi_len, start_dev, w_dev, h_dev, n, out, out_dtype=_lib.DTYPE_U8, channels=1,
                   layout=_lib.LAYOUT_NCHW, offset=0):
        """K1 on `n` ROIs starting at descriptor `offset` (device tensors)."""
        self._ck(self.lib.spk_preprocess(
            self.ctx, ptr(roi_dev), roi_len, start_dev.data_ptr() + 8 * offset, w_dev.data_ptr() + 4 * offset,
            h_dev.data_ptr() + 4 * offset, n, self.th, self.tw, _lib.BORDER[self.spec.border], channels, out_dtype,
            layout, None, ptr(out)))

    def forward(self, x_u8, n, probs_out, label_out=None, classified_out=None):
        """K2 + K3 on a preprocessed u8 batch [n, T, T]."""
        self._ck(self.lib.spk_forward(self.ctx, ptr(x_u8), n, self.softmax_scale, ptr(self._thr_dev), ptr(probs_out),
                                      ptr(label_out), ptr(classified_out)))

    def last_logits(self, n):
        """Logits [n, K] of the last `forward` (host copy; test / debug tap)."""
        addr = self.lib.spk_last_logits(self.ctx)

        class _Raw:
            __cuda_array_interface__ = {"shape": (n, self.k), "typestr": "<f4", "data": (addr, False), "version": 2}

        self.synchronize()
        with self.torch.cuda.device(self.device):
            return self.torch.as_tensor(_Raw(), device=self.device).cpu().numpy().copy()

    def read_buffer(self, buf, n):
        """Activation buffer `buf` of the last forward as float32 [n, h, w, c] (test / debug tap)."""
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.spk_net_read_buffer(self.ctx, buf, n, None, 0, C.byref(h), C.byref(w), C.byref(c)))
        out = np.empty((n, h.value, w.value, c.value), np.float32)
        self._ck(self.lib.spk_net_read_buffer(self.ctx, buf, n, ptr(out), out.size, C.byref(h), C.byref(w), C.byref(c)))
        return out

    # ------------------------------------------------------------------ per-kernel timing
    def profile_begin(self, stamps=False):
        """Per-launch timing on: CUDA events around every launch, or (`stamps`) in-kernel global-timer stamps that leave the
        launches back to back (programmatic dependent launch keeps overlapping) and report IN-STEP times."""
        self._ck(self.lib.spk_profile_mode(self.ctx, 1 if stamps else 0))
        self._ck(self.lib.spk_profile_begin(self.ctx))

    def profile_read(self, detail=False):
        """-> {category: {"ms", "flops", "bytes", "launches"}} (+ "detail": list of per-launch tuples)."""
        n = len(_lib.PROF_CATEGORIES)
        ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        la = (C.c_int64 * n)()
        cap = 1 << 20 if detail else 0
        buf = C.create_string_buffer(cap) if detail else None
        self._ck(self.lib.spk_profile_read(self.ctx, ms, fl, by, la, buf, cap))
        out = {name: {"ms": ms[i], "flops": fl[i], "bytes": by[i], "launches": int(la[i])}
               for i, name in enumerate(_lib.PROF_CATEGORIES) if la[i]}
        if detail:
            rows = []
            for line in buf.value.decode().splitlines():
                c, t, f, b, what = line.split(" ", 4)
                rows.append((_lib.PROF_CATEGORIES[int(c)], float(t), float(f), float(b), what))
            out["detail"] = rows
        return out

    def profile_end(self):
        self._ck(self.lib.spk_profile_end(self.ctx))

    def fault_count(self):
        v = C.c_int64()
        self._ck(self.lib.spk_fault_count(self.ctx, C.byref(v)))
        return v.value

    def synchronize(self):
        self._ck(self.lib.spk_synchronize(self.ctx))

    # ------------------------------------------------------------------ one bin
    def run_bin_device(self, roi_dev, roi_len, start_dev, w_dev, h_dev, n, probs_dev, label_dev=None, cls_dev=None,
                       batch_size=None):
        """Decode + transform + network for the `n` ROIs of one bin, everything already on the device."""
        bs = min(batch_size or self.max_batch, self.max_batch)
        for c in range(0, n, self.pre_chunk):
            cm = min(self.pre_chunk, n - c)
            self.preprocess(roi_dev, roi_len, start_dev, w_dev, h_dev, cm, self._x, offset=c)
            for i in range(0, cm, bs):
                m = min(bs, cm - i)
                lo = c + i
                self.forward(self._x[i:i + m], m, probs_dev[lo:lo + m], None if label_dev is None else label_dev[lo:lo + m],
                             None if cls_dev is None else cls_dev[lo:lo + m])

    def run_bin(self, adc_text, roi_bytes, batch_size=None, want_labels=False):
        """.adc text + .roi bytes (host) -> (roi_id int32[N], probs float32[N,K][, label, classified]).

        Raises `FaultyBin` (ValueError) when a ROI runs past the .roi bytes and `EmptyResize`
        when a ROI would resize to a 0-pixel side -- the two cases in which the reference skips
        the whole bin (probability.py:111-114)."""
        roi_id, w, h, start = parse_adc(adc_text)
        roi_np = np.frombuffer(roi_bytes, dtype=np.uint8) if not isinstance(roi_bytes, np.ndarray) else roi_bytes
        out = self.run_rois(roi_id, w, h, start, roi_np, batch_size, want_labels)
        return (roi_id, *out) if want_labels else (roi_id, out)

    def run_rois(self, roi_id, w, h, start, roi_np, batch_size=None, want_labels=False):
        """ROI descriptors + the byte stream they index (host numpy) -> probs float32[N,K]
        (or (probs, label int32[N], classified bool[N]))."""
        return self.submit_rois(w, h, start, roi_np, None, batch_size, want_labels).result()

    def _pinned(self, nbytes):
        """A pinned host byte buffer of at least `nbytes` from the engine's pool."""
        pool = self.__dict__.setdefault("_pin_pool", [])
        for i, t in enumerate(pool):
            if t.numel() >= nbytes:
                return pool.pop(i)
        return self.torch.empty(max(4096, int(nbytes * 1.25)), dtype=self.torch.uint8, pin_memory=True)

    def _unpin(self, t):
        pool = self.__dict__.setdefault("_pin_pool", [])
        pool.append(t)
        pool.sort(key=lambda x: x.numel())
        del pool[12:]

    def submit_rois(self, w, h, start, roi, roi_len=None, batch_size=None, want_labels=False):
        """Asynchronous `run_rois`: enqueues H2D, K1, K2 + K3 and the D2H of the results on the engine's stream and
        returns a handle whose `.result()` waits for them.  `roi`: the .roi byte stream, a numpy array or a (pinned)
        uint8 torch tensor of which the first `roi_len` bytes are used.  Descriptors go up in one pinned staging copy,
        results come back into pinned buffers, so nothing on this path blocks the host until `.result()`."""
        torch = self.torch
        n = len(w)
        w = np.ascontiguousarray(w, np.int32)
        h = np.ascontiguousarray(h, np.int32)
        start = np.ascontiguousarray(start, np.int64)
        if torch.is_tensor(roi):
            roi_t = roi
        else:
            roi_t = torch.from_numpy(np.ascontiguousarray(roi, np.uint8))
        if roi_len is None:
            roi_len = roi_t.numel()
        validate_rois(w, h, start, roi_len, self.th, self.tw)
        if n == 0:
            probs_h = np.zeros((0, self.k), np.float32)
            return _Ready((probs_h, np.zeros(0, np.int32), np.zeros(0, bool)) if want_labels else probs_h)
        k = self.k
        desc = self._pinned(16 * n)
        dview = desc.numpy()
        dview[:8 * n].view(np.int64)[:] = start
        dview[8 * n:12 * n].view(np.int32)[:] = w
        dview[12 * n:16 * n].view(np.int32)[:] = h
        out_bytes = n * k * 4 + (n * 4 + n if want_labels else 0)
        out = self._pinned(out_bytes)
        # inputs go up on their own stream: the copy of bin i+1 (~25 MB) then overlaps the kernels of bin i
        cs = self.__dict__.get("_copy_stream")
        if cs is None:
            with torch.cuda.device(self.device):
                cs = self._copy_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.device(self.device), torch.cuda.stream(cs):
            roi_dev = roi_t[:max(int(roi_len), 1)].to(self.device, non_blocking=True)
            desc_dev = desc[:16 * n].to(self.device, non_blocking=True)
            up = torch.cuda.Event()
            up.record(cs)
        roi_dev.record_stream(self.stream)
        desc_dev.record_stream(self.stream)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.stream.wait_event(up)
            start_dev = desc_dev[:8 * n].view(torch.int64)
            w_dev = desc_dev[8 * n:12 * n].view(torch.int32)
            h_dev = desc_dev[12 * n:16 * n].view(torch.int32)
            res_dev = torch.empty(out_bytes, dtype=torch.uint8, device=self.device)
            probs = res_dev[:n * k * 4].view(torch.float32).view(n, k)
            label = res_dev[n * k * 4:n * k * 4 + n * 4].view(torch.int32) if want_labels else None
            cls = res_dev[n * k * 4 + n * 4:out_bytes] if want_labels else None
            self.run_bin_device(roi_dev, int(roi_len), start_dev, w_dev, h_dev, n, probs, label, cls, batch_size)
            out[:out_bytes].copy_(res_dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return _Pending(self, ev, n, k, want_labels, out, desc, (roi_t, roi_dev, desc_dev, res_dev))


class _Ready:
    def __init__(self, value):
        self.value = value

    def result(self):
        return self.value


class _Pending:
    """Results of one `Engine.submit_rois` call, in flight on the engine's stream."""

    def __init__(self, eng, event, n, k, want_labels, out, desc, keep):
        self.eng, self.event, self.n, self.k, self.want_labels = eng, event, n, k, want_labels
        self.out, self.desc, self.keep = out, desc, keep

    def result(self):
        self.event.synchronize()
        eng = self.eng
        if eng.precision == _lib.PRECISION_FP32_TC:
            # the split format of FP32_TC holds |x| <= 65504: the convolution epilogue counts every tile that left that range
            faults = eng.fault_count()
            if faults > eng.__dict__.get("_faults_seen", 0):
                eng._faults_seen = faults
                eng._unpin(self.out)
                eng._unpin(self.desc)
                self.keep = None
                raise ArithmeticError("fp32_tc: an activation left the fp16 range of the split format (|x| > 65504); "
                                      "these probabilities are not valid -- rerun with --precision fp32")
        n, k = self.n, self.k
        raw = self.out.numpy()
        probs = raw[:n * k * 4].view(np.float32).reshape(n, k).copy()
        ret = probs
        if self.want_labels:
            label = raw[n * k * 4:n * k * 4 + n * 4].view(np.int32).copy()
            cls = raw[n * k * 4 + n * 4:n * k * 4 + n * 5].astype(bool)
            ret = (probs, label, cls)
        self.eng._unpin(self.out)
        self.eng._unpin(self.desc)
        self.keep = None
        return ret


# ---------------------------------------------------------------------- host helpers over the C ABI
def parse_adc(adc_text):
    """.adc text (str or bytes) -> (roi_id int32[N], width int32[N], height int32[N], start int64[N]).

    sykepic/utils/ifcb.py:101-110: ROI id = 1-based line number, rows with width or height < 1 skipped."""
    lib = _lib.load()
    data = adc_text.encode() if isinstance(adc_text, str) else bytes(adc_text)
    cap = data.count(b"\n") + data.count(b"\r") + 1
    roi_id = np.empty(cap, np.int32)
    w = np.empty(cap, np.int32)
    h = np.empty(cap, np.int32)
    start = np.empty(cap, np.int64)
    n = C.c_int64()
    lines = C.c_int64()
    _lib.check(lib.spk_adc_parse(data, len(data), cap, ptr(roi_id), ptr(w), ptr(h), ptr(start), C.byref(n), C.byref(lines)))
    k = n.value
    return roi_id[:k].copy(), w[:k].copy(), h[:k].copy(), start[:k].copy()


def load_bin(sample_path, roi_buf, th, tw):
    """One bin from disk through the C ABI (`spk_bin_load`: both file reads, the .adc parse and the geometry checks in one
    call that holds no interpreter lock).  `roi_buf`: writable uint8 buffer (numpy array or pinned torch tensor) that
    receives the .roi bytes.  -> (roi_id, w, h, start, roi_len).  Raises FaultyBin / EmptyResize / AdcParseError / OSError
    as the per-bin error policy expects; `CapacityError` (with `.needed`) when `roi_buf` is too small."""
    lib = _lib.load()
    sample_path = Path(sample_path)
    adc_path, roi_path = sample_path.with_suffix(".adc"), sample_path.with_suffix(".roi")
    cap = os.path.getsize(adc_path) // 24 + 16  # a row has 24 comma-separated fields: at least 47 bytes with its newline
    roi_id = np.empty(cap, np.int32)
    w = np.empty(cap, np.int32)
    h = np.empty(cap, np.int32)
    start = np.empty(cap, np.int64)
    n, roi_len = C.c_int64(), C.c_int64()
    nbytes = roi_buf.numel() if hasattr(roi_buf, "numel") else roi_buf.size
    rc = lib.spk_bin_load(os.fsencode(adc_path), os.fsencode(roi_path), ptr(roi_buf), nbytes, cap, ptr(roi_id), ptr(w), ptr(h),
                          ptr(start), C.byref(n), C.byref(roi_len), th, tw)
    if rc == _lib.SPK_ERR_CAPACITY and roi_len.value > nbytes:
        raise CapacityError(roi_len.value)
    _lib.check(rc)
    k = n.value
    return roi_id[:k], w[:k], h[:k], start[:k], roi_len.value


class CapacityError(Exception):
    def __init__(self, needed):
        super().__init__(f"buffer too small: {needed} bytes needed")
        self.needed = needed


def write_prob_csv(csv_path, classes, roi_id, probs):
    """`.prob.csv` formatted and written by the library in one call (probability.py:200-206) -> bytes written."""
    lib = _lib.load()
    header = ("roi," + ",".join(classes) + "\n").encode()
    roi_id = np.ascontiguousarray(roi_id, dtype=np.int32)
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    k = probs.shape[1] if probs.ndim == 2 else len(classes)
    done = C.c_int64()
    _lib.check(lib.spk_prob_csv_write(os.fsencode(csv_path), header, ptr(roi_id), ptr(probs), len(roi_id), k, C.byref(done)))
    return done.value


def validate_rois(w, h, start, roi_len, th, tw):
    lib = _lib.load()
    bad = C.c_int64()
    _lib.check(lib.spk_rois_validate(ptr(w), ptr(h), ptr(start), len(w), int(roi_len), th, tw, C.byref(bad)))


def format_prob_csv(classes, roi_id, probs):
    """bytes of the .prob.csv (sykepic/compute/probability.py:200-206)."""
    lib = _lib.load()
    header = ("roi," + ",".join(classes) + "\n").encode()
    roi_id = np.ascontiguousarray(roi_id, dtype=np.int32)
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    n = len(roi_id)
    k = probs.shape[1] if probs.ndim == 2 else len(classes)
    cap = len(header) + n * (12 + 8 * k + 1) + 16
    buf = C.create_string_buffer(cap)
    ln = C.c_int64()
    _lib.check(lib.spk_format_prob_csv(header, ptr(roi_id), ptr(probs), n, k, buf, cap, C.byref(ln)))
    return buf.raw[:ln.value]
