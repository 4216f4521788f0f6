"""Tensor-level wrapper over the C ABI (include/cervix_b200.h).

Every method takes/returns torch tensors that live on a CUDA device, hands their
``data_ptr()`` to the library together with the current torch stream, and returns freshly
allocated outputs (torch's caching allocator owns all memory - the library allocates
nothing).  Activations are NHWC (``[n, h, w, c]`` contiguous), fp32 or bf16.

This is the ONLY compute backend of the product: it raises when the shared library is
missing or a tensor is not on a CUDA device.  (tests/emu_backend.py implements the same
method set with plain torch ops, as the per-kernel specification the CUDA results are
checked against, and lets the host-side graph logic be exercised on CPU-only CI.)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, check


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError("cervix_b200 activations must be float32 or bfloat16, got %s" % t.dtype)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class ConvGeom:
    """Geometry of one nn.Conv2d call on an NHWC input."""
    __slots__ = ("n", "h", "w", "cin", "cout", "kh", "kw", "stride", "pad", "dil", "ho", "wo")

    def __init__(self, n, h, w, cin, cout, kh, kw, stride, pad, dil):
        self.n, self.h, self.w, self.cin, self.cout = n, h, w, cin, cout
        self.kh, self.kw, self.stride, self.pad, self.dil = kh, kw, stride, pad, dil
        self.ho = (h + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        self.wo = (w + 2 * pad - dil * (kw - 1) - 1) // stride + 1

    def desc(self, dtype: int) -> ConvDesc:
        return ConvDesc(self.n, self.h, self.w, self.cin, self.cout, self.kh, self.kw, self.stride, self.pad,
                        self.dil, self.ho, self.wo, dtype)


class CudaBackend:
    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.CervixError("cervix_b200 needs a CUDA device; there is no CPU fallback")
        self._sm100 = None

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _chk(*tensors):
        for t in tensors:
            if t is None:
                continue
            if not t.is_cuda:
                raise _lib.CervixError("cervix_b200: tensor on %s - the CUDA path has no CPU fallback" % t.device)
            if not t.is_contiguous():
                raise _lib.CervixError("cervix_b200: non-contiguous tensor passed to the C ABI")

    def is_sm100(self) -> bool:
        if self._sm100 is None:
            self._sm100 = bool(self.lib.cvx_device_is_sm100())
        return self._sm100

    # ------------------------------------------------------------------ layout
    def to_nhwc(self, x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        self._chk(x)
        n, c, h, w = x.shape
        y = torch.empty((n, h, w, c), dtype=dtype, device=x.device)
        check(self.lib.cvx_nchw_to_nhwc(_p(x), _p(y), n, c, h, w, _dt(y), self._stream()), "cvx_nchw_to_nhwc")
        return y

    def split_patches(self, images: torch.Tensor, new_size: int, patch: int, mean, std, dtype: torch.dtype) -> torch.Tensor:
        """[n,3,h,w] fp32 in [0,1] -> [n*k*k, patch, patch, 3] NHWC `dtype`: bilinear resize to new_size^2, x-major
        patch split and (v - mean) / std in one launch (cvx_split_patches)."""
        self._chk(images)
        if images.dtype != torch.float32:
            raise TypeError("split_patches: fp32 NCHW images expected, got %s" % images.dtype)
        n, c, h, w = images.shape
        k = new_size // patch
        y = torch.empty((n * k * k, patch, patch, c), dtype=dtype, device=images.device)
        m = (C.c_float * c)(*[float(v) for v in mean])
        s = (C.c_float * c)(*[float(v) for v in std])
        check(self.lib.cvx_split_patches(_p(images), _p(y), n, c, h, w, new_size, patch, m, s, _dt(y), self._stream()),
              "cvx_split_patches")
        return y

    def split_patches_u8(self, images: torch.Tensor, new_size: int, patch: int, mean, std, dtype: torch.dtype) -> torch.Tensor:
        """[n,h,w,3] uint8 (decoded images) -> [n*k*k, patch, patch, 3] NHWC `dtype`, bit-exact with
        PIL.Image.resize((new_size, new_size), BILINEAR) -> crops -> ToTensor -> Normalize (cvx_split_patches_u8)."""
        from .multimodal.pil_resample import bilinear_tables
        self._chk(images)
        if images.dtype != torch.uint8 or images.dim() != 4:
            raise TypeError("split_patches_u8: uint8 [n,h,w,3] images expected, got %s %s" % (images.dtype, tuple(images.shape)))
        n, h, w, c = images.shape
        k = new_size // patch
        key = (h, w, new_size, str(images.device))
        tabs = self._pil_tables.get(key) if hasattr(self, "_pil_tables") else None
        if tabs is None:
            if not hasattr(self, "_pil_tables"):
                self._pil_tables = {}
            tabs = self._pil_tables[key] = tuple(torch.from_numpy(a).to(images.device).contiguous()
                                                 for a in bilinear_tables(w, new_size) + bilinear_tables(h, new_size))
        xmin, xcnt, xk, ymin, ycnt, yk = tabs
        y = torch.empty((n * k * k, patch, patch, c), dtype=dtype, device=images.device)
        m = (C.c_float * c)(*[float(v) for v in mean])
        s = (C.c_float * c)(*[float(v) for v in std])
        check(self.lib.cvx_split_patches_u8(_p(images), _p(y), n, h, w, c, new_size, patch, _p(xmin), _p(xcnt), _p(xk),
                                            int(xk.shape[1]), _p(ymin), _p(ycnt), _p(yk), int(yk.shape[1]), m, s, _dt(y),
                                            self._stream()), "cvx_split_patches_u8")
        return y

    def finish_batch_u8(self, images_u8: torch.Tensor, labels_u8: Optional[torch.Tensor], num_classes: int,
                        dtype: torch.dtype):
        """uint8 [n,h,w,3] pixels (+ uint8 [n,h,w] class map) -> (NHWC activation / 255 in `dtype`, int64 class map
        with values >= num_classes clamped to num_classes) - the loader tail of dataloader.py:40-42 on the device."""
        self._chk(images_u8, labels_u8)
        if any(t is not None and t.dtype != torch.uint8 for t in (images_u8, labels_u8)):
            raise TypeError("finish_batch_u8: uint8 tensors expected")
        x = None if images_u8 is None else torch.empty(images_u8.shape, dtype=dtype, device=images_u8.device)
        t = None if labels_u8 is None else torch.empty(labels_u8.shape, dtype=torch.int64, device=labels_u8.device)
        check(self.lib.cvx_finish_batch_u8(_p(images_u8), _p(x), 0 if images_u8 is None else images_u8.numel(), _p(labels_u8), _p(t),
                                           0 if labels_u8 is None else labels_u8.numel(), num_classes,
                                           _lib.BF16 if dtype == torch.bfloat16 else _lib.F32,
                                           self._stream()), "cvx_finish_batch_u8")
        return x, t

    def augment_batch(self, samples, src, tables, luts, cubic, batch: int, h: int, w: int, max_elems: int, tmp_bytes: int,
                      vec_cols: int, out=None):
        """The training augmentation of dataloader.py:55-154 for a packed batch (utils/dataloader.py pack_batch):
        descriptors uint8 [batch, 96], sources uint8, tables int32, colour tables uint8 [batch*768], OpenCV's bicubic
        weights int16 [32,32,16] -> (uint8 [batch,h,w,3], uint8 [batch,h,w]).  Four launches (cvx_aug_*)."""
        self._chk(samples, src, tables, luts, cubic)
        if samples.dtype != torch.uint8 or samples.numel() != batch * C.sizeof(_lib.AugSample) or src.dtype != torch.uint8 \
                or tables.dtype != torch.int32 or luts.dtype != torch.uint8 or luts.numel() != batch * 768 \
                or cubic.dtype != torch.int16 or cubic.numel() != 32 * 32 * 16:
            raise TypeError("augment_batch: packed blobs of utils.dataloader.pack_batch expected")
        dev = src.device
        u8 = lambda *shape: torch.empty(shape, dtype=torch.uint8, device=dev)   # noqa: E731
        canvas, blurred, lab0 = u8(batch, h, w, 3), u8(batch, h, w, 3), u8(batch, h, w)
        img, lab = out if out is not None else (u8(batch, h, w, 3), u8(batch, h, w))
        tmp = u8(max(int(tmp_bytes), 16))
        st = self._stream()
        check(self.lib.cvx_aug_resize_rows(_p(samples), batch, _p(src), _p(tables), _p(tmp), int(max_elems), st), "cvx_aug_resize_rows")
        check(self.lib.cvx_aug_compose(_p(samples), batch, _p(src), _p(tmp), _p(tables), _p(canvas), _p(lab0), h, w, st), "cvx_aug_compose")
        check(self.lib.cvx_aug_blur5(_p(samples), batch, _p(canvas), _p(blurred), h, w, st), "cvx_aug_blur5")
        check(self.lib.cvx_aug_rotate_jitter(_p(samples), batch, _p(canvas), _p(blurred), _p(lab0), _p(tables), _p(cubic), _p(luts),
                                             _p(img), _p(lab), h, w, int(vec_cols), st), "cvx_aug_rotate_jitter")
        return img, lab

    def to_nchw(self, x: torch.Tensor) -> torch.Tensor:
        self._chk(x)
        n, h, w, c = x.shape
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_nhwc_to_nchw(_p(x), _p(y), n, c, h, w, _dt(x), self._stream()), "cvx_nhwc_to_nchw")
        return y

    def pack_weight(self, w: torch.Tensor, dtype: torch.dtype, transpose_flip: bool) -> torch.Tensor:
        self._chk(w)
        cout, cin, kh, kw = w.shape
        shape = (kh * kw, cin, cout) if transpose_flip else (kh * kw, cout, cin)
        out = torch.empty(shape, dtype=dtype, device=w.device)
        check(self.lib.cvx_pack_weight(_p(w), _p(out), cout, cin, kh, kw, _dt(out), int(transpose_flip), self._stream()),
              "cvx_pack_weight")
        return out

    def unpack_wgrad(self, g: torch.Tensor, cout, cin, kh, kw) -> torch.Tensor:
        self._chk(g)
        if kh * kw == 1 and g.dtype == torch.float32:
            return g.view(cout, cin, 1, 1)          # the packed [1][cout][cin] layout of a 1x1 filter IS its OIHW layout
        out = torch.empty((cout, cin, kh, kw), dtype=torch.float32, device=g.device)
        check(self.lib.cvx_unpack_wgrad(_p(g), _p(out), cout, cin, kh, kw, self._stream()), "cvx_unpack_wgrad")
        return out

    def pack_dw_weight(self, w: torch.Tensor) -> torch.Tensor:
        self._chk(w)
        c = w.shape[0]
        out = torch.empty((9, c), dtype=torch.float32, device=w.device)
        check(self.lib.cvx_pack_dw_weight(_p(w), _p(out), c, self._stream()), "cvx_pack_dw_weight")
        return out

    def unpack_dw_wgrad(self, g: torch.Tensor) -> torch.Tensor:
        self._chk(g)
        c = g.shape[1]
        out = torch.empty((c, 1, 3, 3), dtype=torch.float32, device=g.device)
        check(self.lib.cvx_unpack_dw_wgrad(_p(g), _p(out), c, self._stream()), "cvx_unpack_dw_wgrad")
        return out

    def cat_channels(self, xs: Sequence[torch.Tensor]) -> torch.Tensor:
        self._chk(*xs)
        n, h, w, _ = xs[0].shape
        ctot = sum(int(x.shape[3]) for x in xs)
        y = torch.empty((n, h, w, ctot), dtype=xs[0].dtype, device=xs[0].device)
        off = 0
        for x in xs:
            c = int(x.shape[3])
            check(self.lib.cvx_copy_channels(_p(x), c, 0, _p(y), ctot, off, n * h * w, c, _dt(x), self._stream()),
                  "cvx_copy_channels")
            off += c
        return y

    def slice_channels(self, x: torch.Tensor, off: int, c: int) -> torch.Tensor:
        self._chk(x)
        n, h, w, ctot = x.shape
        y = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
        check(self.lib.cvx_copy_channels(_p(x), ctot, off, _p(y), c, 0, n * h * w, c, _dt(x), self._stream()),
              "cvx_copy_channels")
        return y

    # ------------------------------------------------------------------ dense conv
    def conv_fwd(self, x, wp, bias, g: ConvGeom, tc: bool) -> torch.Tensor:
        self._chk(x, wp, bias)
        y = torch.empty((g.n, g.ho, g.wo, g.cout), dtype=x.dtype, device=x.device)
        d = g.desc(_dt(x))
        fn = self.lib.cvx_conv_fwd_tc if tc else self.lib.cvx_conv_fwd
        check(fn(C.byref(d), _p(x), _p(wp), _p(bias), _p(y), self._stream()), "cvx_conv_fwd" + ("_tc" if tc else ""))
        return y

    def conv_dgrad(self, dy, wpt, g: ConvGeom, tc: bool) -> torch.Tensor:
        self._chk(dy, wpt)
        dx = torch.empty((g.n, g.h, g.w, g.cin), dtype=dy.dtype, device=dy.device)
        d = g.desc(_dt(dy))
        fn = self.lib.cvx_conv_dgrad_tc if tc else self.lib.cvx_conv_dgrad
        check(fn(C.byref(d), _p(dy), _p(wpt), _p(dx), self._stream()), "cvx_conv_dgrad" + ("_tc" if tc else ""))
        return dx

    def conv_wgrad(self, x, dy, g: ConvGeom, tc: bool) -> torch.Tensor:
        self._chk(x, dy)
        dwp = torch.zeros((g.kh * g.kw, g.cout, g.cin), dtype=torch.float32, device=x.device)
        d = g.desc(_dt(x))
        fn = self.lib.cvx_conv_wgrad_tc if tc else self.lib.cvx_conv_wgrad
        check(fn(C.byref(d), _p(x), _p(dy), _p(dwp), self._stream()), "cvx_conv_wgrad" + ("_tc" if tc else ""))
        return dwp

    def bias_grad(self, dy) -> torch.Tensor:
        self._chk(dy)
        c = int(dy.shape[-1])
        rows = dy.numel() // c
        out = torch.empty((c,), dtype=torch.float32, device=dy.device)
        ws = self._ws64((c,), dy.device)
        check(self.lib.cvx_bias_grad(_p(dy), _p(out), _p(ws), rows, c, _dt(dy), self._stream()), "cvx_bias_grad")
        return out

    def subsample(self, x, s: int) -> torch.Tensor:
        self._chk(x)
        n, h, w, c = x.shape
        y = torch.empty((n, (h - 1) // s + 1, (w - 1) // s + 1, c), dtype=x.dtype, device=x.device)
        check(self.lib.cvx_subsample(_p(x), _p(y), n, h, w, c, s, _dt(x), self._stream()), "cvx_subsample")
        return y

    def subsample_bwd(self, dy, h: int, w: int, s: int) -> torch.Tensor:
        self._chk(dy)
        n, _, _, c = dy.shape
        dx = torch.empty((n, h, w, c), dtype=dy.dtype, device=dy.device)
        check(self.lib.cvx_subsample_bwd(_p(dy), _p(dx), n, h, w, c, s, _dt(dy), self._stream()), "cvx_subsample_bwd")
        return dx

    def im2col_narrow(self, x, g: ConvGeom, kpad: int) -> torch.Tensor:
        self._chk(x)
        y = torch.empty((g.n, g.ho, g.wo, kpad), dtype=x.dtype, device=x.device)
        d = g.desc(_dt(x))
        check(self.lib.cvx_im2col_narrow(C.byref(d), _p(x), _p(y), kpad, self._stream()), "cvx_im2col_narrow")
        return y

    def maxpool_fwd(self, x) -> torch.Tensor:
        self._chk(x)
        n, h, w, c = x.shape
        y = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), dtype=x.dtype, device=x.device)
        check(self.lib.cvx_maxpool3x3s2_fwd(_p(x), _p(y), n, h, w, c, _dt(x), self._stream()), "cvx_maxpool3x3s2_fwd")
        return y

    def maxpool_bwd(self, x, y, dy) -> torch.Tensor:
        self._chk(x, y, dy)
        n, h, w, c = x.shape
        dx = torch.empty_like(x)
        check(self.lib.cvx_maxpool3x3s2_bwd(_p(x), _p(y), _p(dy), _p(dx), n, h, w, c, _dt(x), self._stream()),
              "cvx_maxpool3x3s2_bwd")
        return dx

    # ------------------------------------------------------------------ depthwise
    def dw_fwd(self, x, w9c, g: ConvGeom, relu_in: bool) -> torch.Tensor:
        self._chk(x, w9c)
        y = torch.empty((g.n, g.ho, g.wo, g.cin), dtype=x.dtype, device=x.device)
        d = g.desc(_dt(x))
        check(self.lib.cvx_dwconv_fwd(C.byref(d), _p(x), _p(w9c), _p(y), int(relu_in), self._stream()), "cvx_dwconv_fwd")
        return y

    def dw_bwd_data(self, dy, w9c, x, g: ConvGeom, relu_in: bool) -> torch.Tensor:
        self._chk(dy, w9c, x)
        dx = torch.empty((g.n, g.h, g.w, g.cin), dtype=dy.dtype, device=dy.device)
        d = g.desc(_dt(dy))
        check(self.lib.cvx_dwconv_bwd_data(C.byref(d), _p(dy), _p(w9c), _p(x), _p(dx), int(relu_in), self._stream()),
              "cvx_dwconv_bwd_data")
        return dx

    def dw_bwd_weight(self, x, dy, g: ConvGeom, relu_in: bool) -> torch.Tensor:
        self._chk(x, dy)
        out = torch.empty((9, g.cin), dtype=torch.float32, device=x.device)
        ws = self._ws64((9 * g.cin,), x.device)
        d = g.desc(_dt(x))
        check(self.lib.cvx_dwconv_bwd_weight(C.byref(d), _p(x), _p(dy), _p(out), _p(ws), int(relu_in), self._stream()),
              "cvx_dwconv_bwd_weight")
        return out

    # ------------------------------------------------------------------ batch norm
    def bn_forward(self, x, residual, gamma, beta, rmean, rvar, act: int, training: bool, momentum: float, eps: float):
        self._chk(x, residual, gamma, beta, rmean, rvar)
        c = int(x.shape[-1])
        rows = x.numel() // c
        y = torch.empty_like(x)
        mean = torch.empty((c,), dtype=torch.float32, device=x.device)
        invstd = torch.empty((c,), dtype=torch.float32, device=x.device)
        ws = self._ws64((2 * c + 2,), x.device)
        check(self.lib.cvx_bn_forward(_p(x), _p(residual), _p(y), _p(gamma), _p(beta), _p(rmean), _p(rvar), _p(mean),
                                      _p(invstd), _p(ws), rows, c, _dt(x), act, int(training), float(momentum),
                                      float(eps), self._stream()), "cvx_bn_forward")
        return y, mean, invstd

    def bn_backward(self, dy, x, y, gamma, mean, invstd, act: int, training: bool, want_dres: bool, beta=None):
        """beta given (forward without residual): the activation mask is recomputed from x and y is not read."""
        self._chk(dy, x, y, gamma, mean, invstd, beta)
        c = int(x.shape[-1])
        rows = x.numel() // c
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if want_dres else None
        dgamma = torch.empty((c,), dtype=torch.float32, device=x.device)
        dbeta = torch.empty((c,), dtype=torch.float32, device=x.device)
        ws = self._ws64((2 * c + 2,), x.device)
        check(self.lib.cvx_bn_backward(_p(dy), _p(x), _p(y), _p(gamma), _p(beta), _p(mean), _p(invstd), _p(dx), _p(dres),
                                       _p(dgamma), _p(dbeta), _p(ws), rows, c, _dt(x), act, int(training),
                                       self._stream()), "cvx_bn_backward")
        return dx, dres, dgamma, dbeta

    # ------------------------------------------------------------------ small ops
    def relu_fwd(self, x):
        self._chk(x)
        y = torch.empty_like(x)
        check(self.lib.cvx_relu_fwd(_p(x), _p(y), x.numel(), _dt(x), self._stream()), "cvx_relu_fwd")
        return y

    def relu_bwd(self, dy, y):
        self._chk(dy, y)
        dx = torch.empty_like(dy)
        check(self.lib.cvx_relu_bwd(_p(dy), _p(y), _p(dx), dy.numel(), _dt(dy), self._stream()), "cvx_relu_bwd")
        return dx

    def add(self, a, b):
        self._chk(a, b)
        out = torch.empty_like(a)
        check(self.lib.cvx_add(_p(a), _p(b), _p(out), a.numel(), _dt(a), self._stream()), "cvx_add")
        return out

    def spatial_reduce(self, x, scale: float):
        self._chk(x)
        n, h, w, c = x.shape
        y = torch.empty((n, 1, 1, c), dtype=x.dtype, device=x.device)
        check(self.lib.cvx_spatial_reduce(_p(x), _p(y), n, h * w, c, float(scale), _dt(x), self._stream()),
              "cvx_spatial_reduce")
        return y

    def spatial_broadcast(self, x, h: int, w: int, scale: float):
        self._chk(x)
        n, _, _, c = x.shape
        y = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
        check(self.lib.cvx_spatial_broadcast(_p(x), _p(y), n, h * w, c, float(scale), _dt(x), self._stream()),
              "cvx_spatial_broadcast")
        return y

    def upsample_fwd(self, x, ho: int, wo: int):
        self._chk(x)
        n, hi, wi, c = x.shape
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        check(self.lib.cvx_upsample_fwd(_p(x), _p(y), n, hi, wi, ho, wo, c, _dt(x), self._stream()), "cvx_upsample_fwd")
        return y

    def upsample_bwd(self, dy, hi: int, wi: int):
        self._chk(dy)
        n, ho, wo, c = dy.shape
        dx = torch.empty((n, hi, wi, c), dtype=dy.dtype, device=dy.device)
        check(self.lib.cvx_upsample_bwd(_p(dy), _p(dx), n, hi, wi, ho, wo, c, _dt(dy), self._stream()), "cvx_upsample_bwd")
        return dx

    def upsample_concat(self, x, tail):
        """cat([bilinear(x -> tail's size), tail], channels) without materialising the upsampled tensor: the upsample
        kernel writes its channel slice of the concat buffer (deeplabv3_plus.py:184-185)."""
        self._chk(x, tail)
        n, hi, wi, c = x.shape
        _, ho, wo, ct = tail.shape
        y = torch.empty((n, ho, wo, c + ct), dtype=x.dtype, device=x.device)
        st = self._stream()
        check(self.lib.cvx_upsample_into(_p(x), _p(y), n, hi, wi, ho, wo, c, c + ct, 0, _dt(x), st), "cvx_upsample_into")
        check(self.lib.cvx_copy_channels(_p(tail), ct, 0, _p(y), c + ct, c, n * ho * wo, ct, _dt(x), st), "cvx_copy_channels")
        return y

    def upsample_concat_bwd(self, dy, c: int, hi: int, wi: int):
        """-> (gradient of the low-resolution input, gradient of the concatenated tail)."""
        self._chk(dy)
        n, ho, wo, ctot = dy.shape
        dx = torch.empty((n, hi, wi, c), dtype=dy.dtype, device=dy.device)
        check(self.lib.cvx_upsample_from_bwd(_p(dy), _p(dx), n, hi, wi, ho, wo, c, ctot, 0, _dt(dy), self._stream()),
              "cvx_upsample_from_bwd")
        return dx, self.slice_channels(dy, c, ctot - c)

    def upsample_to_nchw_fwd(self, x, ho: int, wo: int):
        self._chk(x)
        n, hi, wi, c = x.shape
        y = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_upsample_to_nchw_fwd(_p(x), _p(y), n, hi, wi, ho, wo, c, _dt(x), self._stream()),
              "cvx_upsample_to_nchw_fwd")
        return y

    def upsample_to_nchw_bwd(self, dy, hi: int, wi: int, dtype: torch.dtype):
        self._chk(dy)
        n, c, ho, wo = dy.shape
        dx = torch.empty((n, hi, wi, c), dtype=dtype, device=dy.device)
        check(self.lib.cvx_upsample_to_nchw_bwd(_p(dy), _p(dx), n, hi, wi, ho, wo, c, _dt(dx), self._stream()),
              "cvx_upsample_to_nchw_bwd")
        return dx

    def dropout_fwd(self, x, p: float, seed: int, step_dev=None, want_mask: bool = True):
        self._chk(x, step_dev)
        y = torch.empty_like(x)
        mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_mask else None
        check(self.lib.cvx_dropout_fwd(_p(x), _p(y), _p(mask), x.numel(), float(p), int(seed) & (2 ** 64 - 1),
                                       _p(step_dev), _dt(x), self._stream()), "cvx_dropout_fwd")
        return y, mask

    def dropout_bwd_seeded(self, dy, p: float, seed: int, step_dev=None):
        """Gradient of dropout_fwd(..., seed, step_dev) with the keep decisions recomputed (no stored mask)."""
        self._chk(dy, step_dev)
        dx = torch.empty_like(dy)
        check(self.lib.cvx_dropout_bwd_seeded(_p(dy), _p(dx), dy.numel(), float(p), int(seed) & (2 ** 64 - 1), _p(step_dev),
                                              _dt(dy), self._stream()), "cvx_dropout_bwd_seeded")
        return dx

    def dropout_bwd(self, dy, mask, p: float):
        self._chk(dy, mask)
        dx = torch.empty_like(dy)
        check(self.lib.cvx_dropout_bwd(_p(dy), _p(mask), _p(dx), dy.numel(), float(p), _dt(dy), self._stream()),
              "cvx_dropout_bwd")
        return dx

    # ------------------------------------------------------------------ loss
    def seg_loss_stats(self, logits, target, onehot, cls_w, alpha: float, gamma: float, thr: float):
        self._chk(logits, target, onehot, cls_w)
        n, c, h, w = logits.shape
        stats = self._ws64((4 + 6 * c,), logits.device)
        check(self.lib.cvx_seg_loss_stats(_p(logits), _p(target), _p(onehot), _p(cls_w), _p(stats), n, c, h, w,
                                          float(alpha), float(gamma), float(thr), self._stream()), "cvx_seg_loss_stats")
        return stats

    def seg_loss_finalize(self, stats, c: int, beta: float, smooth: float):
        self._chk(stats)
        res = torch.empty((4,), dtype=torch.float32, device=stats.device)
        check(self.lib.cvx_seg_loss_finalize(_p(stats), _p(res), c, float(beta), float(smooth), self._stream()),
              "cvx_seg_loss_finalize")
        return res

    def seg_loss_grad(self, logits, target, onehot, cls_w, stats, g, alpha, gamma, beta, smooth):
        self._chk(logits, target, onehot, cls_w, stats, g)
        n, c, h, w = logits.shape
        d = torch.empty_like(logits)
        check(self.lib.cvx_seg_loss_grad(_p(logits), _p(target), _p(onehot), _p(cls_w), _p(stats), _p(g), _p(d), n, c,
                                         h, w, float(alpha), float(gamma), float(beta), float(smooth), self._stream()),
              "cvx_seg_loss_grad")
        return d

    # ------------------------------------------------------------------ fusion-head row operators (fp32)
    def seg_layernorm_fwd(self, x, w, b, groups: int, seg: int, eps: float, mode: int):
        self._chk(x, w, b)
        c = int(x.shape[-1])
        y = torch.empty_like(x)
        stats = torch.empty((groups, 3), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_seg_layernorm_fwd(_p(x), _p(w), _p(b), _p(y), _p(stats), groups, seg, c, float(eps), mode,
                                             self._stream()), "cvx_seg_layernorm_fwd")
        return y, stats

    def seg_layernorm_bwd(self, dy, x, w, stats, groups: int, seg: int, eps: float, mode: int):
        self._chk(dy, x, w, stats)
        c = int(x.shape[-1])
        dx = torch.empty_like(x)
        dw = torch.empty((c,), dtype=torch.float32, device=x.device)
        db = torch.empty((c,), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_seg_layernorm_bwd(_p(dy), _p(x), _p(w), _p(stats), _p(dx), _p(dw), _p(db), groups, seg, c,
                                             float(eps), mode, self._stream()), "cvx_seg_layernorm_bwd")
        return dx, dw, db

    def gelu_fwd(self, x):
        self._chk(x)
        y = torch.empty_like(x)
        check(self.lib.cvx_gelu_fwd(_p(x), _p(y), x.numel(), self._stream()), "cvx_gelu_fwd")
        return y

    def gelu_bwd(self, dy, x):
        self._chk(dy, x)
        dx = torch.empty_like(x)
        check(self.lib.cvx_gelu_bwd(_p(dy), _p(x), _p(dx), x.numel(), self._stream()), "cvx_gelu_bwd")
        return dx

    def graph_gather(self, x, groups: int, nodes: int, rowptr, col, w):
        self._chk(x, rowptr, col, w)
        c = int(x.shape[-1])
        out = torch.empty_like(x)
        check(self.lib.cvx_graph_gather(_p(x), _p(out), groups, nodes, c, _p(rowptr), _p(col), _p(w), self._stream()),
              "cvx_graph_gather")
        return out

    def gate_pool_fwd(self, x, gate, groups: int, seg: int):
        self._chk(x, gate)
        c = int(x.shape[-1])
        pooled = torch.empty((groups, c), dtype=torch.float32, device=x.device)
        att = torch.empty((groups * seg,), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_gate_pool_fwd(_p(x), _p(gate), _p(pooled), _p(att), groups, seg, c, self._stream()),
              "cvx_gate_pool_fwd")
        return pooled, att

    def gate_pool_bwd(self, dpooled, x, att, groups: int, seg: int):
        self._chk(dpooled, x, att)
        c = int(x.shape[-1])
        dx = torch.empty_like(x)
        dgate = torch.empty((groups * seg,), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_gate_pool_bwd(_p(dpooled), _p(x), _p(att), _p(dx), _p(dgate), groups, seg, c, self._stream()),
              "cvx_gate_pool_bwd")
        return dx, dgate

    def attn_small_fwd(self, qkv, b: int, n: int, h: int, d: int, scale: float, drop_p: float, seed: int, step_dev=None):
        self._chk(qkv, step_dev)
        out = torch.empty((b, n, h * d), dtype=torch.float32, device=qkv.device)
        probs = torch.empty((b, h, n, n), dtype=torch.float32, device=qkv.device)
        check(self.lib.cvx_attn_small_fwd(_p(qkv), _p(out), _p(probs), b, n, h, d, float(scale), float(drop_p),
                                          int(seed) & (2 ** 64 - 1), _p(step_dev), self._stream()), "cvx_attn_small_fwd")
        return out, probs

    def attn_small_bwd(self, dout, qkv, probs, b, n, h, d, scale, drop_p, seed, step_dev=None):
        self._chk(dout, qkv, probs, step_dev)
        dqkv = torch.empty_like(qkv)
        check(self.lib.cvx_attn_small_bwd(_p(dout), _p(qkv), _p(probs), _p(dqkv), b, n, h, d, float(scale), float(drop_p),
                                          int(seed) & (2 ** 64 - 1), _p(step_dev), self._stream()), "cvx_attn_small_bwd")
        return dqkv

    def l2norm_fwd(self, x):
        self._chk(x)
        rows, c = x.shape
        y = torch.empty_like(x)
        norms = torch.empty((rows,), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_l2norm_fwd(_p(x), _p(y), _p(norms), rows, c, self._stream()), "cvx_l2norm_fwd")
        return y, norms

    def l2norm_bwd(self, dy, y, norms):
        self._chk(dy, y, norms)
        rows, c = y.shape
        dx = torch.empty_like(y)
        check(self.lib.cvx_l2norm_bwd(_p(dy), _p(y), _p(norms), _p(dx), rows, c, self._stream()), "cvx_l2norm_bwd")
        return dx

    def rows_gather(self, x, idx, fill):
        self._chk(x, idx, fill)
        rows, c = int(idx.shape[0]), int(x.shape[-1])
        y = torch.empty((rows, c), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_rows_gather(_p(x), _p(idx), _p(fill), _p(y), rows, c, self._stream()), "cvx_rows_gather")
        return y

    def rows_scatter_add(self, dy, idx, src_rows: int, want_fill: bool):
        self._chk(dy, idx)
        rows, c = dy.shape
        dx = torch.empty((src_rows, c), dtype=torch.float32, device=dy.device)
        dfill = torch.empty((c,), dtype=torch.float32, device=dy.device) if want_fill else None
        check(self.lib.cvx_rows_scatter_add(_p(dy), _p(idx), _p(dx), _p(dfill), rows, src_rows, c, self._stream()),
              "cvx_rows_scatter_add")
        return dx, dfill

    # ------------------------------------------------------------------ segment-table row operators (all modalities at once)
    @staticmethod
    def _param_sets(ws, bs, row_start, dws=None, dbs=None):
        ps = _lib.ParamSets()
        n = len(row_start) - 1
        if n > _lib.MAX_PARAM_SETS:
            raise _lib.CervixError("cervix_b200: at most %d parameter sets per launch" % _lib.MAX_PARAM_SETS)
        ps.sets = n
        for i in range(n):
            ps.w[i] = None if ws is None else ws[i].data_ptr()
            ps.b[i] = None if bs is None else bs[i].data_ptr()
            ps.dw[i] = None if dws is None else dws[i].data_ptr()
            ps.db[i] = None if dbs is None else dbs[i].data_ptr()
        for i in range(n + 1):
            ps.row_start[i] = int(row_start[i])
        return ps

    def segtab_layernorm_fwd(self, x, ws, bs, tab, eps: float, mode: int):
        """tab: SegTable-like with device int32 ``seg_start, seg_len, seg_set, row_seg`` and host ``row_start, segments``."""
        self._chk(x, *ws, *bs)
        c = int(x.shape[-1])
        y = torch.empty_like(x)
        stats = torch.empty((tab.segments, 3), dtype=torch.float32, device=x.device)
        ps = self._param_sets(ws, bs, tab.row_start)
        check(self.lib.cvx_segtab_layernorm_fwd(_p(x), C.byref(ps), _p(y), _p(stats), _p(tab.seg_start), _p(tab.seg_len),
                                                _p(tab.seg_set), tab.segments, c, float(eps), mode, self._stream()),
              "cvx_segtab_layernorm_fwd")
        return y, stats

    def segtab_layernorm_bwd(self, dy, x, ws, stats, tab, eps: float, mode: int):
        self._chk(dy, x, stats, *ws)
        c = int(x.shape[-1])
        dx = torch.empty_like(x)
        dws = [torch.empty((c,), dtype=torch.float32, device=x.device) for _ in ws]
        dbs = [torch.empty((c,), dtype=torch.float32, device=x.device) for _ in ws]
        ps = self._param_sets(ws, None, tab.row_start, dws, dbs)
        check(self.lib.cvx_segtab_layernorm_bwd(_p(dy), _p(x), C.byref(ps), _p(stats), _p(dx), _p(tab.seg_start),
                                                _p(tab.seg_len), _p(tab.seg_set), _p(tab.row_seg), tab.segments, c, float(eps),
                                                mode, self._stream()), "cvx_segtab_layernorm_bwd")
        return dx, dws, dbs

    def segtab_gate_pool_fwd(self, x, gate, tab):
        self._chk(x, gate)
        c = int(x.shape[-1])
        pooled = torch.empty((tab.segments, c), dtype=torch.float32, device=x.device)
        att = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_segtab_gate_pool_fwd(_p(x), _p(gate), _p(pooled), _p(att), _p(tab.seg_start), _p(tab.seg_len),
                                                tab.segments, tab.max_len, c, self._stream()), "cvx_segtab_gate_pool_fwd")
        return pooled, att

    def segtab_gate_pool_bwd(self, dpooled, x, att, tab):
        self._chk(dpooled, x, att)
        c = int(x.shape[-1])
        dx = torch.empty_like(x)
        dgate = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
        check(self.lib.cvx_segtab_gate_pool_bwd(_p(dpooled), _p(x), _p(att), _p(dx), _p(dgate), _p(tab.seg_start),
                                                _p(tab.seg_len), tab.segments, tab.max_len, c, self._stream()),
              "cvx_segtab_gate_pool_bwd")
        return dx, dgate

    def segtab_bcast_add(self, x, t, tab, tok_of_seg):
        self._chk(x, t, tok_of_seg)
        y = torch.empty_like(x)
        check(self.lib.cvx_segtab_bcast_add(_p(x), _p(t), _p(y), _p(tab.row_seg), _p(tok_of_seg), int(x.shape[0]),
                                            int(x.shape[1]), self._stream()), "cvx_segtab_bcast_add")
        return y

    def segtab_bcast_add_bwd(self, dy, tab, seg_of_tok, tokens: int):
        self._chk(dy, seg_of_tok)
        dt = torch.empty((tokens, dy.shape[1]), dtype=torch.float32, device=dy.device)
        check(self.lib.cvx_segtab_bcast_add_bwd(_p(dy), _p(dt), _p(seg_of_tok), _p(tab.seg_start), _p(tab.seg_len), tokens,
                                                int(dy.shape[1]), self._stream()), "cvx_segtab_bcast_add_bwd")
        return dt

    def gemm_grouped(self, problems):
        """One launch for a list of independent fp32 GEMMs C = A B^T (+ bias) (cvx_gemm_grouped).  Each problem is a dict
        with tensors ``a``, ``b``, ``c`` (+ optional ``bias``, ``rowsum``), ints ``m, n, k`` and element strides
        ``lda_m, lda_k, ldb_n, ldb_k, ldc``.  Lists longer than the per-launch limit are split."""
        lim = _lib.MAX_GEMM_PROBLEMS
        for s0 in range(0, len(problems), lim):
            chunk = problems[s0:s0 + lim]
            arr = (_lib.GemmProblem * len(chunk))()
            for i, q in enumerate(chunk):
                self._chk_f32(q["a"], q["b"], q["c"], q.get("bias"), q.get("rowsum"))
                arr[i].a, arr[i].b, arr[i].c = q["a"].data_ptr(), q["b"].data_ptr(), q["c"].data_ptr()
                arr[i].bias = None if q.get("bias") is None else q["bias"].data_ptr()
                arr[i].rowsum = None if q.get("rowsum") is None else q["rowsum"].data_ptr()
                arr[i].m, arr[i].n, arr[i].k = int(q["m"]), int(q["n"]), int(q["k"])
                arr[i].lda_m, arr[i].lda_k = int(q["lda_m"]), int(q["lda_k"])
                arr[i].ldb_n, arr[i].ldb_k, arr[i].ldc = int(q["ldb_n"]), int(q["ldb_k"]), int(q["ldc"])
            check(self.lib.cvx_gemm_grouped(arr, len(chunk), self._stream()), "cvx_gemm_grouped")

    @staticmethod
    def _chk_f32(*tensors):
        for t in tensors:
            if t is None:
                continue
            if not t.is_cuda or t.dtype != torch.float32:
                raise _lib.CervixError("cervix_b200: gemm_grouped takes fp32 CUDA tensors (got %s on %s)" % (t.dtype, t.device))

    def softmax_ce(self, logits, labels, loss, weight: float, want_grad: bool, out=None):
        self._chk(logits, labels, loss, out)
        b, k = logits.shape
        d = (out if out is not None else torch.empty_like(logits)) if want_grad else None
        check(self.lib.cvx_softmax_ce(_p(logits), _p(labels), _p(loss), _p(d), b, k, float(weight), self._stream()),
              "cvx_softmax_ce")
        return d

    def masked_mse(self, a, b, sel, loss, weight: float, inv_count: float, want_grad: bool):
        self._chk(a, b, sel, loss)
        rows, c = a.shape
        da = torch.empty_like(a) if want_grad else None
        db = torch.empty_like(a) if want_grad else None
        check(self.lib.cvx_masked_mse(_p(a), _p(b), _p(sel), _p(loss), _p(da), _p(db), rows, c, float(weight),
                                      float(inv_count), self._stream()), "cvx_masked_mse")
        return da, db

    # ------------------------------------------------------------------ fused separable-conv chain (csrc/sepconv.cu)
    def sepconv_fused_ok(self, x: torch.Tensor, cin: int, cout: int, stride: int, dil: int, pad: int) -> bool:
        """The fused chain takes bf16 NHWC tensors, 3x3/stride 1/dilation 1/pad 1 depthwise, channel counts % 8."""
        return (x.dtype == torch.bfloat16 and self.is_sm100() and stride == 1 and dil == 1 and pad == 1
                and cin % 8 == 0 and cout % 8 == 0 and cout <= 2048)

    # ------------------------------------------------------------------ per-step arena of zeroed fp64 workspaces
    # Every statistics / reduction kernel accumulates into a small fp64 workspace that has to start at zero: ~360
    # cudaMemsetAsync nodes per training step (0.5 ms).  A trainer brackets its step with zero_arena_begin / _end: the
    # arena is cleared by ONE fill at the start of the step, workspaces are carved from it in call order (the same
    # addresses on every step, so a captured graph stays valid) and the library is told to skip its own memsets.
    _ARENA_DOUBLES = 4 * 1024 * 1024

    def zero_arena_begin(self, device):
        if getattr(self, "_arena", None) is None or self._arena.device != device:
            self._arena = torch.zeros((self._ARENA_DOUBLES,), dtype=torch.float64, device=device)
        else:
            self._arena.zero_()
        self._arena_pos = 0
        self._arena_on = True
        self.lib.cvx_set_ws_prezeroed(1)

    def zero_arena_end(self):
        self._arena_on = False
        self.lib.cvx_set_ws_prezeroed(0)

    def _ws64(self, shape, dev):
        n = 1
        for d in shape:
            n *= int(d)
        if getattr(self, "_arena_on", False):
            a = (self._arena_pos + 1) & ~1              # 16-byte aligned slices
            if a + n <= self._arena.numel() and self._arena.device == dev:
                self._arena_pos = a + n
                return self._arena[a:a + n].view(shape)
            return torch.zeros(shape, dtype=torch.float64, device=dev)     # arena full: the library will not clear it
        return torch.empty(shape, dtype=torch.float64, device=dev)

    @staticmethod
    def _f32(n, dev):
        return torch.empty((n,), dtype=torch.float32, device=dev)

    def dwf_fwd(self, x, w9c, in_scale, in_shift, relu_in: bool, g: ConvGeom, want_stats: bool):
        self._chk(x, w9c, in_scale, in_shift)
        y = torch.empty_like(x)
        stats = self._ws64((2, g.cin), x.device) if want_stats else None
        d = g.desc(_dt(x))
        check(self.lib.cvx_dwf_fwd(C.byref(d), _p(x), _p(w9c), _p(in_scale), _p(in_shift), int(relu_in), _p(y), _p(stats),
                                   self._stream()), "cvx_dwf_fwd")
        return y, stats

    def dwf_bwd(self, dd, dside, negk, kmean, x, w9c, in_scale, in_shift, relu_in: bool, addend, g: ConvGeom, want_sums: bool):
        self._chk(dd, dside, negk, kmean, x, w9c, in_scale, in_shift, addend)
        gout = torch.empty_like(x)
        dw9c = torch.empty((9, g.cin), dtype=torch.float32, device=x.device)
        ws = self._ws64((9, g.cin), x.device)
        sums = self._ws64((2, g.cin), x.device) if want_sums else None
        d = g.desc(_dt(x))
        check(self.lib.cvx_dwf_bwd(C.byref(d), _p(dd), _p(dside), _p(negk), _p(kmean), _p(x), _p(w9c), _p(in_scale),
                                   _p(in_shift), int(relu_in), _p(addend), _p(gout), _p(dw9c), _p(ws), _p(sums),
                                   self._stream()), "cvx_dwf_bwd")
        return gout, dw9c, sums

    def bn_stats(self, x):
        self._chk(x)
        c = int(x.shape[-1])
        stats = self._ws64((2, c), x.device)
        check(self.lib.cvx_bn_stats(_p(x), _p(stats), x.numel() // c, c, _dt(x), self._stream()), "cvx_bn_stats")
        return stats

    def bn_affine(self, stats, rows: int, gamma, beta, rmean, rvar, momentum: float, eps: float, mean_offset=None):
        self._chk(stats, gamma, beta, rmean, rvar, mean_offset)
        c = int(gamma.shape[0])
        out = torch.empty((4, c), dtype=torch.float32, device=gamma.device)      # mean, invstd, scale, shift
        check(self.lib.cvx_bn_affine(_p(stats), rows, _p(gamma), _p(beta), _p(mean_offset), _p(rmean), _p(rvar), out[0].data_ptr(),
                                     out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), c, float(momentum), float(eps),
                                     self._stream()), "cvx_bn_affine")
        return out[0], out[1], out[2], out[3]

    def pw_fold(self, weight, scale, shift, dtype):
        self._chk(weight, scale, shift)
        cout, cin = int(weight.shape[0]), int(weight.shape[1])
        wp = torch.empty((1, cout, cin), dtype=dtype, device=weight.device)
        wpt = torch.empty((1, cin, cout), dtype=dtype, device=weight.device)
        bias = self._f32(cout, weight.device)
        check(self.lib.cvx_pw_fold(_p(weight), _p(scale), _p(shift), _p(wp), _p(wpt), _p(bias), cout, cin, self._stream()),
              "cvx_pw_fold")
        return wp, wpt, bias

    def conv_fwd_ex(self, x, wp, bias, g: ConvGeom, side=None, side_scale=None, want_stats=False):
        self._chk(x, wp, bias, side, side_scale)
        y = torch.empty((g.n, g.ho, g.wo, g.cout), dtype=x.dtype, device=x.device)
        stats = self._ws64((2, g.cout), x.device) if want_stats else None
        d = g.desc(_dt(x))
        check(self.lib.cvx_conv_fwd_tc_ex(C.byref(d), _p(x), _p(wp), _p(bias), _p(side), _p(side_scale), _p(stats), _p(y),
                                          self._stream()), "cvx_conv_fwd_tc_ex")
        return y, stats

    def conv_fwd_act(self, x, wp, bias, g: ConvGeom, act: int = 0, side=None, side_scale=None):
        """y = act(conv(x) + bias + side_scale * side): the inference group conv -> folded bn -> (+identity) -> relu."""
        self._chk(x, wp, bias, side, side_scale)
        y = torch.empty((g.n, g.ho, g.wo, g.cout), dtype=x.dtype, device=x.device)
        d = g.desc(_dt(x))
        check(self.lib.cvx_conv_fwd_tc_act(C.byref(d), _p(x), _p(wp), _p(bias), _p(side), _p(side_scale), int(act), _p(y),
                                           self._stream()), "cvx_conv_fwd_tc_act")
        return y

    def conv_dgrad_ex(self, dy, wpt, g: ConvGeom, bias=None, side=None, side_scale=None):
        self._chk(dy, wpt, bias, side, side_scale)
        dx = torch.empty((g.n, g.h, g.w, g.cin), dtype=dy.dtype, device=dy.device)
        d = g.desc(_dt(dy))
        check(self.lib.cvx_conv_dgrad_tc_ex(C.byref(d), _p(dy), _p(wpt), _p(bias), _p(side), _p(side_scale), _p(dx),
                                            self._stream()), "cvx_conv_dgrad_tc_ex")
        return dx

    def affine_act(self, p, scale, shift, res, act: int):
        self._chk(p, scale, shift, res)
        c = int(p.shape[-1])
        y = torch.empty_like(p)
        check(self.lib.cvx_affine_act(_p(p), _p(res), _p(y), _p(scale), _p(shift), p.numel() // c, c, act, self._stream()),
              "cvx_affine_act")
        return y

    def bn_bwd_sums(self, dy, y, p, act: int):
        self._chk(dy, y, p)
        c = int(p.shape[-1])
        sums = self._ws64((2, c), p.device)
        check(self.lib.cvx_bn_bwd_sums(_p(dy), _p(y), _p(p), _p(sums), p.numel() // c, c, act, self._stream()), "cvx_bn_bwd_sums")
        return sums

    def bn_bwd_coef(self, sums, rows: int, mean, invstd, gamma):
        self._chk(sums, mean, invstd, gamma)
        c = int(gamma.shape[0])
        out = torch.empty((5, c), dtype=torch.float32, device=gamma.device)      # a, b, cc, dgamma, dbeta
        check(self.lib.cvx_bn_bwd_coef(_p(sums), rows, _p(mean), _p(invstd), _p(gamma), out[0].data_ptr(), out[1].data_ptr(),
                                       out[2].data_ptr(), out[3].data_ptr(), out[4].data_ptr(), c, self._stream()),
              "cvx_bn_bwd_coef")
        return out[0], out[1], out[2], out[3], out[4]

    def bn_bwd_affine(self, dy, y, p, a, b, cc, act: int, want_g: bool):
        self._chk(dy, y, p, a, b, cc)
        c = int(p.shape[-1])
        dp = torch.empty_like(p)
        gout = torch.empty_like(p) if want_g else None
        check(self.lib.cvx_bn_bwd_affine(_p(dy), _p(y), _p(p), _p(a), _p(b), _p(cc), _p(dp), _p(gout), p.numel() // c, c, act,
                                         self._stream()), "cvx_bn_bwd_affine")
        return dp, gout

    def pw_bwd_coef(self, gp, weight, scale, invstd, mean, rows: int):
        self._chk(gp, weight, scale, invstd, mean)
        cout, cin = int(weight.shape[0]), int(weight.shape[1])
        dw = torch.empty_like(weight)
        out = torch.empty((5, cin), dtype=torch.float32, device=weight.device)   # colsum, dgamma, dbeta, negk, kmean
        check(self.lib.cvx_pw_bwd_coef(_p(gp), _p(weight), _p(scale), _p(invstd), _p(mean), rows, _p(dw), out[0].data_ptr(),
                                       out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), out[4].data_ptr(), cout, cin,
                                       self._stream()), "cvx_pw_bwd_coef")
        return dw, out[1], out[2], out[3], out[4]

    # ------------------------------------------------------------------ optimizer
    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, wd, step_t, grad_scale=1.0):
        self._chk(p, g, m, v)
        check(self.lib.cvx_adam_step(_p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2),
                                     float(eps), float(wd), int(step_t), float(grad_scale), self._stream()),
              "cvx_adam_step")

    def adam_step_dev(self, p, g, m, v, hyper, step_dev):
        self._chk(p, g, m, v, hyper, step_dev)
        check(self.lib.cvx_adam_step_dev(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(hyper), _p(step_dev), self._stream()),
              "cvx_adam_step_dev")

    def sgd_step(self, p, g, buf, lr, momentum, wd, nesterov, first_step, grad_scale=1.0):
        self._chk(p, g, buf)
        check(self.lib.cvx_sgd_step(_p(p), _p(g), _p(buf), p.numel(), float(lr), float(momentum), float(wd),
                                    int(nesterov), int(first_step), float(grad_scale), self._stream()), "cvx_sgd_step")

    def sgd_step_dev(self, p, g, buf, hyper, nesterov: bool):
        self._chk(p, g, buf, hyper)
        check(self.lib.cvx_sgd_step_dev(_p(p), _p(g), _p(buf), p.numel(), _p(hyper), int(nesterov), self._stream()),
              "cvx_sgd_step_dev")

    def multi_gather_chunk(self) -> int:
        return int(self.lib.cvx_multi_gather_chunk())

    def multi_gather(self, src_ptrs, chunk_tensor, chunk_start, dst_offsets, sizes, dst):
        """dst[dst_offsets[i] : +sizes[i]] = tensor at address src_ptrs[i] (zeros when the address is 0)."""
        self._chk(src_ptrs, chunk_tensor, chunk_start, dst_offsets, sizes, dst)
        check(self.lib.cvx_multi_gather(_p(src_ptrs), _p(chunk_tensor), _p(chunk_start), _p(dst_offsets), _p(sizes),
                                        int(chunk_tensor.numel()), _p(dst), self._stream()), "cvx_multi_gather")

    # ------------------------------------------------------------------ inference post-processing
    def seg_postprocess(self, logits, crop, out_hw, want_probs: bool = False):
        """logits [C,H,W] fp32 (one image, NCHW) -> (uint8 class map [out_h,out_w], probabilities or None)."""
        self._chk(logits)
        c, h, w = logits.shape
        cy, cx, ch, cw = (int(v) for v in crop)
        oh, ow = int(out_hw[0]), int(out_hw[1])
        cls = torch.empty((oh, ow), dtype=torch.uint8, device=logits.device)
        probs = torch.empty((oh, ow, c), dtype=torch.float32, device=logits.device) if want_probs else None
        check(self.lib.cvx_seg_postprocess(_p(logits), c, h, w, cy, cx, ch, cw, oh, ow, _p(cls), _p(probs), self._stream()),
              "cvx_seg_postprocess")
        return cls, probs

    def confusion_matrix(self, pred, gt, classes: int, hist=None):
        """hist[gt][pred] += 1 over the pixels with gt < classes; pred / gt uint8 tensors of equal size."""
        self._chk(pred, gt)
        if hist is None:
            hist = torch.zeros((classes, classes), dtype=torch.int64, device=pred.device)
        check(self.lib.cvx_confusion_matrix(_p(pred), _p(gt), pred.numel(), int(classes), _p(hist), self._stream()),
              "cvx_confusion_matrix")
        return hist


_BACKEND = None


def get_backend():
    """The process-wide compute backend (the CUDA library).  Raises if it cannot be loaded."""
    global _BACKEND
    if _BACKEND is None:
        _BACKEND = CudaBackend()
    return _BACKEND


def set_backend(b):
    """Test hook: install a different implementation of the backend method set."""
    global _BACKEND
    prev = _BACKEND
    _BACKEND = b
    return prev
