"""Helpers shared by the `-m gpu` tests: a raw spk context and a K1 runner."""

import ctypes as C

import numpy as np

from sykepic_b200 import _lib, engine


class RawCtx:
    def __init__(self, device=0):
        import torch

        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.ctx = C.c_void_p()
        _lib.check(self.lib.spk_create(device, C.c_void_p(self.stream.cuda_stream), C.byref(self.ctx)))

    def close(self):
        if self.ctx.value:
            self.lib.spk_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def ck(self, rc):
        _lib.check(rc, self.ctx)

    def sync(self):
        self.ck(self.lib.spk_synchronize(self.ctx))

    def preprocess(self, roi_bytes, w, h, start, t, border="mode", out_dtype=_lib.DTYPE_U8, channels=1,
                   layout=_lib.LAYOUT_NCHW, lut=None):
        """-> numpy array [n, C, T, T] (NCHW) or [n, T, T, C] (NHWC); u8 -> [n, T, T]."""
        torch = self.torch
        n = len(w)
        tdt = {_lib.DTYPE_U8: torch.uint8, _lib.DTYPE_F32: torch.float32, _lib.DTYPE_BF16: torch.bfloat16}[out_dtype]
        shape = (n, t, t) if out_dtype == _lib.DTYPE_U8 else ((n, channels, t, t) if layout == _lib.LAYOUT_NCHW else (n, t, t, channels))
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            roi_dev = torch.from_numpy(np.ascontiguousarray(roi_bytes, np.uint8)).to(self.device)
            if roi_dev.numel() == 0:
                roi_dev = torch.zeros(16, dtype=torch.uint8, device=self.device)
            w_dev = torch.from_numpy(np.ascontiguousarray(w, np.int32)).to(self.device)
            h_dev = torch.from_numpy(np.ascontiguousarray(h, np.int32)).to(self.device)
            s_dev = torch.from_numpy(np.ascontiguousarray(start, np.int64)).to(self.device)
            lut_dev = None if lut is None else torch.from_numpy(np.ascontiguousarray(lut, np.float32)).to(self.device)
            out = torch.full(shape, 7, dtype=tdt, device=self.device)
            self.ck(self.lib.spk_preprocess(self.ctx, roi_dev.data_ptr(), len(roi_bytes), s_dev.data_ptr(), w_dev.data_ptr(),
                                            h_dev.data_ptr(), n, t, t, _lib.BORDER[border], channels, out_dtype, layout,
                                            None if lut_dev is None else lut_dev.data_ptr(), out.data_ptr()))
            self.sync()
            if out_dtype == _lib.DTYPE_BF16:
                return out.float().cpu().numpy()
            return out.cpu().numpy()

    def fault_count(self):
        v = C.c_int64()
        self.ck(self.lib.spk_fault_count(self.ctx, C.byref(v)))
        return v.value


def bin_arrays(b):
    rid, w, h, start = engine.parse_adc(b["adc_text"])
    return rid, w, h, start, np.asarray(b["roi_bytes"], np.uint8)
