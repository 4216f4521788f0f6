"""Plain-torch emulation of the backend method set (TEST INFRASTRUCTURE ONLY).

``EmuBackend`` restates, op by op, what each C-ABI entry point of include/cervix_b200.h must
compute, using stock torch ops in fp32.  It is used two ways:
  * on CPU-only CI it is installed in place of the CUDA backend so the host-side graph logic
    (module wiring, autograd shells, weight packing conventions) is checked against the oracle;
  * on the GPU box every CUDA kernel is compared with the corresponding method here.
The product never imports this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _nchw(x):  # NHWC -> NCHW fp32
    return x.permute(0, 3, 1, 2).float()


def _nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


class EmuBackend:
    name = "emu"

    def is_sm100(self):
        return False

    # ---- layout
    def to_nhwc(self, x, dtype):
        return _nhwc(x.float(), dtype)

    def split_patches(self, images, new_size, patch, mean, std, dtype):
        x = F.interpolate(images.float(), size=(new_size, new_size), mode="bilinear", align_corners=False)
        c = x.shape[1]
        x = (x - torch.tensor(list(mean)).view(1, c, 1, 1)) / torch.tensor(list(std)).view(1, c, 1, 1)
        k, b = new_size // patch, x.shape[0]
        x = x.reshape(b, c, k, patch, k, patch).permute(0, 4, 2, 3, 5, 1)      # [b, ix, iy, py, px, c]
        return x.reshape(b * k * k, patch, patch, c).contiguous().to(dtype)

    def finish_batch_u8(self, images_u8, labels_u8, num_classes, dtype):
        x = None if images_u8 is None else (images_u8.float() * (1.0 / 255.0)).to(dtype)
        t = None if labels_u8 is None else labels_u8.long().clamp(max=num_classes)
        return x, t

    def augment_batch(self, samples, src, tables, luts, cubic, batch, h, w, max_elems, tmp_bytes, vec_cols, out=None):
        """Interpreter of csrc/augment.cu on the PACKED blobs (descriptors, integer tables and colour tables exactly as
        the host module built them), in numpy: checks the host side of utils/dataloader.py against the goldens on CPU."""
        import numpy as np
        from cervix_b200.utils.dataloader import AUG_SAMPLE
        from oracle import augment_ref as A
        rec = np.frombuffer(samples.cpu().numpy().tobytes(), dtype=AUG_SAMPLE)
        srcb, tab = src.cpu().numpy(), tables.cpu().numpy().astype(np.int64)
        lutb, cub = luts.cpu().numpy(), cubic.cpu().numpy().astype(np.int64).reshape(32, 32, 4, 4)
        imgs = np.zeros((batch, h, w, 3), np.uint8)
        labs = np.zeros((batch, h, w), np.uint8)

        def resample(cur, off, n_out, taps, axis):
            lo, cnt = tab[off:off + n_out], tab[off + n_out:off + 2 * n_out]
            k = tab[off + 2 * n_out:off + 2 * n_out + n_out * taps].reshape(n_out, taps)
            outs = []
            for o in range(n_out):
                seg = np.take(cur, np.arange(lo[o], lo[o] + cnt[o]), axis=axis)
                wts = k[o, :cnt[o]].reshape([-1 if a == axis else 1 for a in range(3)])
                outs.append(np.clip(((seg * wts).sum(axis) + (1 << 21)) >> 22, 0, 255))
            return np.stack(outs, axis=axis)

        for i, s in enumerate(rec):
            ih, iw, nh, nw = int(s["ih"]), int(s["iw"]), int(s["nh"]), int(s["nw"])
            img = srcb[s["src_off"]:s["src_off"] + ih * iw * 3].reshape(ih, iw, 3).astype(np.int64)
            lab = srcb[s["lab_off"]:s["lab_off"] + ih * iw].reshape(ih, iw)
            if iw != nw:
                img = resample(img, int(s["xtab"]), nw, int(s["xtaps"]), 1)
            if ih != nh:
                img = resample(img, int(s["ytab"]), nh, int(s["ytaps"]), 0)
            lab = lab[tab[s["ynn"]:s["ynn"] + nh]][:, tab[s["xnn"]:s["xnn"] + nw]]
            img = img.astype(np.uint8)
            if s["flip"]:
                img, lab = img[:, ::-1], lab[:, ::-1]
            img = A.paste((h, w), 128, img, int(s["dx"]), int(s["dy"]))
            lab = A.paste((h, w), 0, lab, int(s["dx"]), int(s["dy"]))
            if s["blur"]:
                img = A.cv_gaussian5_u8(img)
            if s["rotate"]:
                r0 = int(s["rot"])
                ad, bd = tab[r0:r0 + w], tab[r0 + w:r0 + 2 * w]
                x0, y0 = tab[r0 + 2 * w:r0 + 2 * w + h], tab[r0 + 2 * w + h:r0 + 2 * w + 2 * h]
                X, Y = (x0[:, None] + 16 + ad[None, :]) >> 5, (y0[:, None] + 16 + bd[None, :]) >> 5
                wt = cub[Y & 31, X & 31]
                acc = np.full((h, w, 3), 1 << 14, np.int64)
                pix = img.astype(np.int64)
                for k1 in range(4):
                    sy = (Y >> 5) - 1 + k1
                    for k2 in range(4):
                        sx = (X >> 5) - 1 + k2
                        ok = (sx >= 0) & (sx < w) & (sy >= 0) & (sy < h)
                        px = np.where(ok[..., None], pix[np.clip(sy, 0, h - 1), np.clip(sx, 0, w - 1)], 128)
                        acc += px * wt[:, :, k1, k2][..., None]
                img = np.clip(acc >> 15, 0, 255).astype(np.uint8)
                nx, ny = (x0[:, None] + 512 + ad[None, :]) >> 10, (y0[:, None] + 512 + bd[None, :]) >> 10
                ok = (nx >= 0) & (nx < w) & (ny >= 0) & (ny < h)
                lab = np.where(ok, lab[np.clip(ny, 0, h - 1), np.clip(nx, 0, w - 1)], 0).astype(np.uint8)
            if s["lut"] >= 0:
                lt = lutb[s["lut"]:s["lut"] + 768]
                hsv = A.cv_rgb2hsv_u8(img)
                hsv = np.stack([lt[hsv[..., 0]], lt[256 + hsv[..., 1].astype(np.int64)], lt[512 + hsv[..., 2].astype(np.int64)]], -1)
                assert vec_cols == (w // A.CV_SIMD_PIXELS) * A.CV_SIMD_PIXELS
                img = A.cv_hsv2rgb_u8(hsv)
            imgs[i], labs[i] = img, lab
        ti, tl = torch.from_numpy(imgs).to(src.device), torch.from_numpy(labs).to(src.device)
        if out is not None:
            out[0].copy_(ti)
            out[1].copy_(tl)
            return out
        return ti, tl

    def upsample_concat(self, x, tail):
        return torch.cat([self.upsample_fwd(x, tail.shape[1], tail.shape[2]), tail], dim=3).contiguous()

    def upsample_concat_bwd(self, dy, c, hi, wi):
        return self.upsample_bwd(dy[..., :c].contiguous(), hi, wi), dy[..., c:].contiguous()

    def to_nchw(self, x):
        return _nchw(x).contiguous()

    def pack_weight(self, w, dtype, transpose_flip):
        cout, cin, kh, kw = w.shape
        if not transpose_flip:
            return w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).contiguous().to(dtype)
        wf = torch.flip(w, dims=(2, 3))
        return wf.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout).contiguous().to(dtype)

    def unpack_wgrad(self, g, cout, cin, kh, kw):
        return g.reshape(kh, kw, cout, cin).permute(2, 3, 0, 1).contiguous().float()

    def pack_dw_weight(self, w):
        return w.reshape(w.shape[0], 9).t().contiguous().float()

    def unpack_dw_wgrad(self, g):
        return g.t().reshape(g.shape[1], 1, 3, 3).contiguous().float()

    def cat_channels(self, xs):
        return torch.cat(list(xs), dim=3).contiguous()

    def slice_channels(self, x, off, c):
        return x[..., off:off + c].contiguous()

    # ---- dense conv (weights arrive PACKED, exactly as the C ABI receives them)
    @staticmethod
    def _unpack(wp, g):
        return wp.float().reshape(g.kh, g.kw, g.cout, g.cin).permute(2, 3, 0, 1).contiguous()

    @staticmethod
    def _unpack_t(wpt, g):
        w = wpt.float().reshape(g.kh, g.kw, g.cin, g.cout).permute(3, 2, 0, 1)
        return torch.flip(w, dims=(2, 3)).contiguous()

    def conv_fwd(self, x, wp, bias, g, tc):
        y = F.conv2d(_nchw(x), self._unpack(wp, g), None if bias is None else bias.float(), g.stride, g.pad, g.dil)
        assert y.shape[2] == g.ho and y.shape[3] == g.wo
        return _nhwc(y, x.dtype)

    def conv_dgrad(self, dy, wpt, g, tc):
        w = self._unpack_t(wpt, g)
        dx = torch.nn.grad.conv2d_input((g.n, g.cin, g.h, g.w), w, _nchw(dy), g.stride, g.pad, g.dil)
        return _nhwc(dx, dy.dtype)

    def conv_wgrad(self, x, dy, g, tc):
        dw = torch.nn.grad.conv2d_weight(_nchw(x), (g.cout, g.cin, g.kh, g.kw), _nchw(dy), g.stride, g.pad, g.dil)
        return dw.permute(2, 3, 0, 1).reshape(g.kh * g.kw, g.cout, g.cin).contiguous()

    def bias_grad(self, dy):
        return dy.float().reshape(-1, dy.shape[-1]).sum(0)

    def subsample(self, x, s):
        return x[:, ::s, ::s, :].contiguous()

    def subsample_bwd(self, dy, h, w, s):
        dx = torch.zeros((dy.shape[0], h, w, dy.shape[3]), dtype=dy.dtype, device=dy.device)
        dx[:, ::s, ::s, :] = dy
        return dx

    def im2col_narrow(self, x, g, kpad):
        cols = F.unfold(_nchw(x), (g.kh, g.kw), dilation=g.dil, padding=g.pad, stride=g.stride)  # [n, cin*taps, L]
        n = x.shape[0]
        cols = cols.reshape(n, g.cin, g.kh * g.kw, g.ho, g.wo).permute(0, 3, 4, 2, 1).reshape(n, g.ho, g.wo, -1)
        return F.pad(cols, (0, kpad - cols.shape[-1])).contiguous().to(x.dtype)

    def maxpool_fwd(self, x):
        return _nhwc(F.max_pool2d(_nchw(x), 3, 2, 1), x.dtype)

    def maxpool_bwd(self, x, y, dy):
        with torch.enable_grad():
            xi = _nchw(x).detach().requires_grad_(True)
            F.max_pool2d(xi, 3, 2, 1).backward(_nchw(dy))
        return _nhwc(xi.grad, x.dtype)

    # ---- depthwise
    @staticmethod
    def _dw_w(w9c):
        return w9c.t().reshape(w9c.shape[1], 1, 3, 3).contiguous()

    def dw_fwd(self, x, w9c, g, relu_in):
        xi = _nchw(x)
        if relu_in:
            xi = F.relu(xi)
        return _nhwc(F.conv2d(xi, self._dw_w(w9c), None, g.stride, g.pad, g.dil, groups=g.cin), x.dtype)

    def dw_bwd_data(self, dy, w9c, x, g, relu_in):
        dx = torch.nn.grad.conv2d_input((g.n, g.cin, g.h, g.w), self._dw_w(w9c), _nchw(dy), g.stride, g.pad, g.dil,
                                        groups=g.cin)
        if relu_in:
            dx = dx * (_nchw(x) > 0)
        return _nhwc(dx, dy.dtype)

    def dw_bwd_weight(self, x, dy, g, relu_in):
        xi = _nchw(x)
        if relu_in:
            xi = F.relu(xi)
        dw = torch.nn.grad.conv2d_weight(xi, (g.cin, 1, 3, 3), _nchw(dy), g.stride, g.pad, g.dil, groups=g.cin)
        return dw.reshape(g.cin, 9).t().contiguous()

    # ---- batch norm
    @staticmethod
    def _act(v, act):
        return F.relu(v) if act == 1 else (F.relu6(v) if act == 2 else v)

    @staticmethod
    def _mask(y, act):
        if act == 1:
            return (y > 0).float()
        if act == 2:
            return ((y > 0) & (y < 6)).float()
        return torch.ones_like(y, dtype=torch.float32)

    def bn_forward(self, x, residual, gamma, beta, rmean, rvar, act, training, momentum, eps):
        c = x.shape[-1]
        xf = x.float().reshape(-1, c)
        if training:
            mean = xf.double().mean(0)
            var = xf.double().var(0, unbiased=False)
            n = xf.shape[0]
            if rmean is not None:
                rmean.mul_(1 - momentum).add_(momentum * mean.float())
                rvar.mul_(1 - momentum).add_(momentum * (var * n / max(n - 1, 1)).float())
            mean, invstd = mean.float(), (1.0 / torch.sqrt(var + eps)).float()
        else:
            mean, invstd = rmean.clone().float(), 1.0 / torch.sqrt(rvar.float() + eps)
        y = (xf - mean) * (invstd * gamma.float()) + beta.float()
        if residual is not None:
            y = y + residual.float().reshape(-1, c)
        return self._act(y, act).reshape(x.shape).to(x.dtype), mean, invstd

    def bn_backward(self, dy, x, y, gamma, mean, invstd, act, training, want_dres, beta=None):
        c = x.shape[-1]
        dz = dy.float().reshape(-1, c)
        if act != 0:
            if beta is not None:      # mask recomputed from x (forward without residual): y is not needed
                sc = gamma.float() * invstd
                y = self._act(x.float().reshape(-1, c) * sc + (beta.float() - mean * sc), act)
            dz = dz * self._mask(y.float().reshape(-1, c), act)
        xhat = (x.float().reshape(-1, c) - mean) * invstd
        dbeta = dz.double().sum(0).float()
        dgamma = (dz * xhat).double().sum(0).float()
        k = gamma.float() * invstd
        if training:
            n = dz.shape[0]
            dx = k * (dz - dbeta / n - xhat * dgamma / n)
        else:
            dx = k * dz
        dres = dz.reshape(x.shape).to(x.dtype) if want_dres else None
        return dx.reshape(x.shape).to(x.dtype), dres, dgamma, dbeta

    # ---- small ops
    def relu_fwd(self, x):
        return F.relu(x)

    def relu_bwd(self, dy, y):
        return dy * (y > 0)

    def add(self, a, b):
        return (a.float() + b.float()).to(a.dtype)

    def spatial_reduce(self, x, scale):
        return (x.float().sum(dim=(1, 2), keepdim=True) * scale).to(x.dtype)

    def spatial_broadcast(self, x, h, w, scale):
        return (x.float() * scale).expand(x.shape[0], h, w, x.shape[3]).contiguous().to(x.dtype)

    def upsample_fwd(self, x, ho, wo):
        return _nhwc(F.interpolate(_nchw(x), size=(ho, wo), mode="bilinear", align_corners=True), x.dtype)

    def upsample_bwd(self, dy, hi, wi):
        n, ho, wo, c = dy.shape
        with torch.enable_grad():
            x = torch.zeros((n, c, hi, wi), dtype=torch.float32, device=dy.device, requires_grad=True)
            F.interpolate(x, size=(ho, wo), mode="bilinear", align_corners=True).backward(_nchw(dy))
        return _nhwc(x.grad, dy.dtype)

    def upsample_to_nchw_fwd(self, x, ho, wo):
        return F.interpolate(_nchw(x), size=(ho, wo), mode="bilinear", align_corners=True).contiguous()

    def upsample_to_nchw_bwd(self, dy, hi, wi, dtype):
        n, c, ho, wo = dy.shape
        with torch.enable_grad():
            x = torch.zeros((n, c, hi, wi), dtype=torch.float32, device=dy.device, requires_grad=True)
            F.interpolate(x, size=(ho, wo), mode="bilinear", align_corners=True).backward(dy.float())
        return _nhwc(x.grad, dtype)

    def dropout_fwd(self, x, p, seed, step_dev=None):
        if step_dev is not None:
            seed = seed + 7919 * int(step_dev)
        g = torch.Generator(device="cpu").manual_seed(seed % (2 ** 63))
        mask = (torch.rand(x.shape, generator=g) >= p).to(torch.uint8).to(x.device)
        return (x.float() * mask / (1 - p)).to(x.dtype), mask

    def dropout_bwd(self, dy, mask, p):
        return (dy.float() * mask / (1 - p)).to(dy.dtype)

    # ---- loss
    def seg_loss_stats(self, logits, target, onehot, cls_w, alpha, gamma, thr):
        n, c, h, w = logits.shape
        z = logits.double().permute(0, 2, 3, 1).reshape(-1, c)
        t = target.reshape(-1)
        valid = (t >= 0) & (t < c)
        ts = torch.where(valid, t, torch.zeros_like(t))
        logp = torch.log_softmax(z, -1)
        p = logp.exp()
        wt = (cls_w.double()[ts] if cls_w is not None else torch.ones_like(ts, dtype=torch.float64)) * valid
        u = wt * logp.gather(1, ts[:, None])[:, 0]
        pt = u.exp()
        focal = -((1 - pt) ** gamma) * alpha * u
        oh = onehot.double().reshape(-1, c + 1)[:, :c] if onehot is not None else \
            F.one_hot(torch.where(valid, t, torch.full_like(t, c)), c + 1)[:, :c].double()
        hard = (p.float() > thr).double()
        stats = torch.cat([torch.stack([-(u.sum()), wt.sum(), focal.sum(), torch.tensor(float(z.shape[0]), dtype=torch.float64, device=z.device)]),
                           (oh * p).sum(0), p.sum(0), oh.sum(0), (oh * hard).sum(0), hard.sum(0), oh.sum(0)])
        return stats

    @staticmethod
    def _dice(tp, sp, st, beta, smooth):
        b2 = beta * beta
        fp, fn = sp - tp, st - tp
        return (((1 + b2) * tp + smooth) / ((1 + b2) * tp + b2 * fn + fp + smooth)).mean()

    def seg_loss_finalize(self, stats, c, beta, smooth):
        s = stats
        ce = s[0] / s[1]
        focal = s[2] / s[3]
        dice = 1 - self._dice(s[4:4 + c], s[4 + c:4 + 2 * c], s[4 + 2 * c:4 + 3 * c], beta, smooth)
        fsc = self._dice(s[4 + 3 * c:4 + 4 * c], s[4 + 4 * c:4 + 5 * c], s[4 + 5 * c:4 + 6 * c], beta, smooth)
        return torch.stack([ce, focal, dice, fsc]).float()

    def seg_loss_grad(self, logits, target, onehot, cls_w, stats, g, alpha, gamma, beta, smooth):
        # autograd through an independent fp64 restatement of the three losses
        with torch.enable_grad():
            return self._seg_loss_grad(logits, target, onehot, cls_w, g, alpha, gamma, beta, smooth)

    def _seg_loss_grad(self, logits, target, onehot, cls_w, g, alpha, gamma, beta, smooth):
        n, c, h, w = logits.shape
        z = logits.double().detach().requires_grad_(True)
        flat = z.permute(0, 2, 3, 1).reshape(-1, c)
        t = target.reshape(-1)
        valid = (t >= 0) & (t < c)
        ts = torch.where(valid, t, torch.zeros_like(t))
        logp = torch.log_softmax(flat, -1)
        wt = (cls_w.double()[ts] if cls_w is not None else torch.ones_like(ts, dtype=torch.float64)) * valid
        u = wt * logp.gather(1, ts[:, None])[:, 0]
        ce = -(u.sum()) / wt.sum()
        focal = (-((1 - u.exp()) ** gamma) * alpha * u).mean()
        p = logp.exp()
        oh = onehot.double().reshape(-1, c + 1)[:, :c] if onehot is not None else \
            F.one_hot(torch.where(valid, t, torch.full_like(t, c)), c + 1)[:, :c].double()
        dice = 1 - self._dice((oh * p).sum(0), p.sum(0), oh.sum(0), beta, smooth)
        gd = g.double()
        (gd[0] * ce + gd[1] * focal + gd[2] * dice).backward()
        return z.grad.float()

    # ---- fused separable-conv chain (specification of csrc/sepconv.cu, dwconv_fused.cu, conv_tc.cu's TcEpi)
    def sepconv_fused_ok(self, x, cin, cout, stride, dil, pad):
        return stride == 1 and dil == 1 and pad == 1

    @staticmethod
    def _virt(x, in_scale, in_shift, relu_in):
        v = _nchw(x)
        if in_scale is not None:
            v = v * in_scale.view(1, -1, 1, 1) + in_shift.view(1, -1, 1, 1)
        return F.relu(v) if relu_in else v

    @staticmethod
    def _stats(y):
        v = y.float().reshape(-1, y.shape[-1]).double()
        return torch.stack([v.sum(0), (v * v).sum(0)])

    def dwf_fwd(self, x, w9c, in_scale, in_shift, relu_in, g, want_stats):
        v = self._virt(x, in_scale, in_shift, relu_in)
        y = _nhwc(F.conv2d(v, self._dw_w(w9c), None, 1, 1, 1, groups=g.cin), x.dtype)
        return y, (self._stats(y) if want_stats else None)

    def dwf_bwd(self, dd, dside, negk, kmean, x, w9c, in_scale, in_shift, relu_in, addend, g, want_sums):
        v = self._virt(x, in_scale, in_shift, relu_in)
        if dside is not None:     # bn1's backward applied on load (zero padding applies to the assembled gradient)
            dd = (dd.float() + negk * dside.float() + kmean).to(torch.float32)
        gx = torch.nn.grad.conv2d_input((g.n, g.cin, g.h, g.w), self._dw_w(w9c), _nchw(dd), 1, 1, 1, groups=g.cin)
        if relu_in:
            gx = gx * (v > 0)
        gq = _nhwc(gx, x.dtype)
        sums = None
        if want_sums:
            a = gq.float().reshape(-1, g.cin).double(); b = x.float().reshape(-1, g.cin).double()
            sums = torch.stack([a.sum(0), (a * b).sum(0)])
        dw = torch.nn.grad.conv2d_weight(v, (g.cin, 1, 3, 3), _nchw(dd), 1, 1, 1, groups=g.cin)
        out = gq if addend is None else (gx + _nchw(addend)).permute(0, 2, 3, 1).contiguous().to(x.dtype)
        return out, dw.reshape(g.cin, 9).t().contiguous(), sums

    def bn_stats(self, x):
        return self._stats(x)

    def bn_affine(self, stats, rows, gamma, beta, rmean, rvar, momentum, eps, mean_offset=None):
        mean = stats[0] / rows
        var = (stats[1] / rows - mean * mean).clamp_min(0)
        invstd = 1.0 / torch.sqrt(var + eps)
        if rmean is not None:
            unbiased = var * rows / (rows - 1) if rows > 1 else var
            true_mean = mean if mean_offset is None else mean + mean_offset.double()
            rmean.copy_(((1 - momentum) * rmean.double() + momentum * true_mean).float())
            rvar.copy_(((1 - momentum) * rvar.double() + momentum * unbiased).float())
        mean, invstd = mean.float(), invstd.float()
        scale = gamma.float() * invstd
        return mean, invstd, scale, beta.float() - mean * scale

    def pw_fold(self, weight, scale, shift, dtype):
        w = weight.float().reshape(weight.shape[0], weight.shape[1])
        wp = (w * scale.view(1, -1)).to(dtype)
        return wp.unsqueeze(0).contiguous(), wp.t().unsqueeze(0).contiguous(), w @ shift

    def conv_fwd_ex(self, x, wp, bias, g, side=None, side_scale=None, want_stats=False):
        y = F.conv2d(_nchw(x), self._unpack(wp, g), None if bias is None else bias.float(), g.stride, g.pad, g.dil)
        if side is not None:
            y = y + _nchw(side) * side_scale.view(1, -1, 1, 1)
        y = _nhwc(y, x.dtype)
        return y, (self._stats(y) if want_stats else None)

    def conv_fwd_act(self, x, wp, bias, g, act=0, side=None, side_scale=None):
        y = F.conv2d(_nchw(x), self._unpack(wp, g), None if bias is None else bias.float(), g.stride, g.pad, g.dil)
        y = y.permute(0, 2, 3, 1).float()
        if side is not None:
            y = y + (side.float() if side_scale is None else side.float() * side_scale.view(1, 1, 1, -1))
        if act == 1:
            y = torch.relu(y)
        return y.to(x.dtype)

    def conv_dgrad_ex(self, dy, wpt, g, bias=None, side=None, side_scale=None):
        w = self._unpack_t(wpt, g)
        dx = torch.nn.grad.conv2d_input((g.n, g.cin, g.h, g.w), w, _nchw(dy), g.stride, g.pad, g.dil)
        if bias is not None:
            dx = dx + bias.view(1, -1, 1, 1)
        if side is not None:
            dx = dx + _nchw(side) * side_scale.view(1, -1, 1, 1)
        return _nhwc(dx, dy.dtype)

    def affine_act(self, p, scale, shift, res, act):
        v = p.float() * scale + shift
        if res is not None:
            v = v + res.float()
        return self._act(v, act).to(p.dtype)

    def _g(self, dy, y, act):
        g = dy.float()
        return g if act == 0 else g * self._mask(y.float(), act)

    def bn_bwd_sums(self, dy, y, p, act):
        g = self._g(dy, y, act).reshape(-1, p.shape[-1]).double()
        return torch.stack([g.sum(0), (g * p.float().reshape(-1, p.shape[-1]).double()).sum(0)])

    def bn_bwd_coef(self, sums, rows, mean, invstd, gamma):
        sg, sgp = sums[0], sums[1]
        mu, is_ = mean.double(), invstd.double()
        dgam = is_ * (sgp - mu * sg)
        sc = gamma.double() * is_
        b = -sc * is_ * dgam / rows
        return sc.float(), b.float(), (-sc * sg / rows - b * mu).float(), dgam.float(), sg.float()

    def bn_bwd_affine(self, dy, y, p, a, b, cc, act, want_g):
        g = self._g(dy, y, act)
        dp = (a * g + b * p.float() + cc).to(p.dtype)
        return dp, (g.to(p.dtype) if want_g else None)

    def pw_bwd_coef(self, gp, weight, scale, invstd, mean, rows):
        cout, cin = weight.shape[0], weight.shape[1]
        G = gp.reshape(cout, cin).float()
        w = weight.float().reshape(cout, cin)
        colsum = (w * G).sum(0)
        dgam = invstd * colsum
        k = scale * invstd * dgam / rows
        return (G * scale.view(1, -1)).reshape(weight.shape), dgam, torch.zeros_like(dgam), -k, k * mean

    # ---- fusion-head row operators
    @staticmethod
    def _ln(x, w, b, groups, seg, eps, mode):
        c = x.shape[-1]
        xs = x.reshape(groups, seg * c)
        mean = xs.mean(1, keepdim=True)
        xc = xs - mean
        sd = xc.pow(2).mean(1, keepdim=True).sqrt()
        r = 1 / (sd + eps) if mode == 0 else 1 / torch.sqrt(sd * sd + eps)
        y = (xc * r).reshape(-1, c) * w + b
        return y.reshape(x.shape), torch.cat([mean, r, sd], 1)

    def seg_layernorm_fwd(self, x, w, b, groups, seg, eps, mode):
        return self._ln(x, w, b, groups, seg, eps, mode)

    def seg_layernorm_bwd(self, dy, x, w, stats, groups, seg, eps, mode):
        with torch.enable_grad():
            xi = x.detach().requires_grad_(True); wi = w.detach().requires_grad_(True)
            bi = torch.zeros_like(w).requires_grad_(True)
            y, _ = self._ln(xi, wi, bi, groups, seg, eps, mode)
            y.backward(dy)
        return xi.grad, wi.grad, bi.grad

    def gelu_fwd(self, x):
        return F.gelu(x)

    def gelu_bwd(self, dy, x):
        with torch.enable_grad():
            xi = x.detach().requires_grad_(True)
            F.gelu(xi).backward(dy)
        return xi.grad

    def graph_gather(self, x, groups, nodes, rowptr, col, w):
        c = x.shape[-1]
        xs = x.reshape(groups, nodes, c)
        out = torch.zeros_like(xs)
        rp = rowptr.tolist(); cl = col.tolist(); ww = w.tolist()
        for i in range(nodes):
            for e in range(rp[i], rp[i + 1]):
                out[:, i] += ww[e] * xs[:, cl[e]]
        return out.reshape(x.shape)

    @staticmethod
    def _pool(x, gate, groups, seg):
        c = x.shape[-1]
        g = gate.reshape(groups, seg)
        e = (g - g.max(1, keepdim=True).values).exp()
        att = e / (e.sum(1, keepdim=True) + 1e-16)
        pooled = (att.unsqueeze(-1) * x.reshape(groups, seg, c)).sum(1)
        return pooled, att.reshape(-1)

    def gate_pool_fwd(self, x, gate, groups, seg):
        return self._pool(x, gate, groups, seg)

    def gate_pool_bwd(self, dpooled, x, att, groups, seg):
        c = x.shape[-1]
        xs = x.reshape(groups, seg, c); a = att.reshape(groups, seg)
        dx = (a.unsqueeze(-1) * dpooled.unsqueeze(1)).reshape(x.shape)
        datt = (dpooled.unsqueeze(1) * xs).sum(-1)
        dgate = a * (datt - (a * datt).sum(1, keepdim=True))
        return dx, dgate.reshape(-1)

    @staticmethod
    def _attn(qkv, b, n, h, d, scale):
        q, k, v = qkv.reshape(b, n, 3, h, d).permute(2, 0, 3, 1, 4)
        p = ((q * scale) @ k.transpose(-2, -1)).softmax(-1)
        return (p @ v).transpose(1, 2).reshape(b, n, h * d), p

    def attn_small_fwd(self, qkv, b, n, h, d, scale, drop_p, seed, step_dev=None):
        assert drop_p == 0.0, "the emulation backend covers the deterministic (eval) attention only"
        return self._attn(qkv, b, n, h, d, scale)

    def attn_small_bwd(self, dout, qkv, probs, b, n, h, d, scale, drop_p, seed, step_dev=None):
        with torch.enable_grad():
            qi = qkv.detach().requires_grad_(True)
            self._attn(qi, b, n, h, d, scale)[0].backward(dout)
        return qi.grad

    def l2norm_fwd(self, x):
        n = x.norm(dim=1).clamp_min(1e-12)
        return x / n.unsqueeze(1), n

    def l2norm_bwd(self, dy, y, norms):
        return (dy - y * (dy * y).sum(1, keepdim=True)) / norms.unsqueeze(1)

    def rows_gather(self, x, idx, fill):
        y = x[idx.clamp_min(0).long()]
        if fill is not None:
            y = torch.where((idx < 0).unsqueeze(1), fill.unsqueeze(0).expand_as(y), y)
        else:
            y = y * (idx >= 0).unsqueeze(1)
        return y.contiguous()

    def rows_scatter_add(self, dy, idx, src_rows, want_fill):
        dx = torch.zeros((src_rows, dy.shape[1]), dtype=dy.dtype, device=dy.device)
        ok = idx >= 0
        dx.index_add_(0, idx[ok].long(), dy[ok])
        dfill = dy[~ok].sum(0) if want_fill else None
        return dx, dfill

    def softmax_ce(self, logits, labels, loss, weight, want_grad, out=None):
        b = logits.shape[0]
        lse = torch.logsumexp(logits, 1)
        loss += weight * (lse - logits.gather(1, labels[:, None])[:, 0]).sum() / b
        if not want_grad:
            return None
        d = weight * (torch.softmax(logits, 1) - F.one_hot(labels, logits.shape[1]).float()) / b
        if out is not None:
            out.copy_(d)
            return out
        return d

    def masked_mse(self, a, b, sel, loss, weight, inv_count, want_grad):
        d = (a - b) * sel.bool().unsqueeze(1)
        loss += weight * (d * d).sum() * inv_count
        if not want_grad:
            return None, None
        return 2 * d * weight * inv_count, -2 * d * weight * inv_count

    # ---- optimizer
    def adam_step(self, p, g, m, v, lr, beta1, beta2, eps, wd, step_t, grad_scale=1.0):
        gi = g * grad_scale
        if wd != 0:
            gi = gi + wd * p
        m.mul_(beta1).add_((1 - beta1) * gi)
        v.mul_(beta2).add_((1 - beta2) * gi * gi)
        bc1 = 1 - beta1 ** step_t
        bc2 = 1 - beta2 ** step_t
        p.sub_((lr / bc1) * m / (v.sqrt() / (bc2 ** 0.5) + eps))

    def adam_step_dev(self, p, g, m, v, hyper, step_dev):
        lr, b1, b2, eps, wd, gs = [float(t) for t in hyper]
        self.adam_step(p, g, m, v, lr, b1, b2, eps, wd, int(step_dev), gs)

    # ---- segment-table operators: plain loops over the segments with the fixed-size primitives above
    @staticmethod
    def _segs(tab):
        return list(zip(tab.seg_start.tolist(), tab.seg_len.tolist(), tab.seg_set.tolist()))

    def segtab_layernorm_fwd(self, x, ws, bs, tab, eps, mode):
        y = torch.empty_like(x)
        stats = torch.empty((tab.segments, 3), dtype=torch.float32)
        for s, (r0, n, k) in enumerate(self._segs(tab)):
            yy, st = self._ln(x[r0:r0 + n], ws[k].detach(), bs[k].detach(), 1, n, eps, mode)
            y[r0:r0 + n] = yy
            stats[s] = st[0]
        return y, stats

    def segtab_layernorm_bwd(self, dy, x, ws, stats, tab, eps, mode):
        dx = torch.empty_like(x)
        c = x.shape[-1]
        dws = [torch.zeros(c) for _ in ws]
        dbs = [torch.zeros(c) for _ in ws]
        for (r0, n, k) in self._segs(tab):
            a, b, d = self.seg_layernorm_bwd(dy[r0:r0 + n], x[r0:r0 + n], ws[k], None, 1, n, eps, mode)
            dx[r0:r0 + n] = a
            dws[k] += b
            dbs[k] += d
        return dx, dws, dbs

    def segtab_gate_pool_fwd(self, x, gate, tab):
        pooled = torch.empty((tab.segments, x.shape[-1]))
        att = torch.empty((x.shape[0],))
        for s, (r0, n, _) in enumerate(self._segs(tab)):
            p, a = self._pool(x[r0:r0 + n], gate[r0:r0 + n], 1, n)
            pooled[s] = p[0]
            att[r0:r0 + n] = a
        return pooled, att

    def segtab_gate_pool_bwd(self, dpooled, x, att, tab):
        dx = torch.empty_like(x)
        dgate = torch.empty((x.shape[0],))
        for s, (r0, n, _) in enumerate(self._segs(tab)):
            a, b = self.gate_pool_bwd(dpooled[s:s + 1], x[r0:r0 + n], att[r0:r0 + n], 1, n)
            dx[r0:r0 + n] = a
            dgate[r0:r0 + n] = b
        return dx, dgate

    def segtab_bcast_add(self, x, t, tab, tok_of_seg):
        tok = tok_of_seg.long()[tab.row_seg.long()]
        add = torch.where((tok >= 0)[:, None], t[tok.clamp_min(0)], torch.zeros(1))
        return x + add

    def segtab_bcast_add_bwd(self, dy, tab, seg_of_tok, tokens):
        dt = torch.zeros((tokens, dy.shape[1]))
        starts, lens = tab.seg_start.tolist(), tab.seg_len.tolist()
        for tok, s in enumerate(seg_of_tok.tolist()):
            if s >= 0:
                dt[tok] = dy[starts[s]:starts[s] + lens[s]].sum(0)
        return dt

    def gemm_grouped(self, problems):
        for q in problems:
            m, n, k = q["m"], q["n"], q["k"]
            A = torch.as_strided(q["a"], (m, k), (q["lda_m"], q["lda_k"]))
            Bm = torch.as_strided(q["b"], (n, k), (q["ldb_n"], q["ldb_k"]))
            out = A.double() @ Bm.double().t()
            if q.get("bias") is not None:
                out = out + q["bias"].double()[None, :]
            torch.as_strided(q["c"], (m, n), (q["ldc"], 1)).copy_(out.float())
            if q.get("rowsum") is not None:
                q["rowsum"].reshape(-1)[:m].copy_(A.double().sum(1).float())

    def sgd_step_dev(self, p, g, buf, hyper, nesterov):
        lr, mom, _, _, wd, gs = [float(t) for t in hyper]
        self.sgd_step(p, g, buf, lr, mom, wd, nesterov, False, gs)

    def sgd_step(self, p, g, buf, lr, momentum, wd, nesterov, first_step, grad_scale=1.0):
        gi = g * grad_scale
        if wd != 0:
            gi = gi + wd * p
        if momentum != 0:
            if first_step:
                buf.copy_(gi)
            else:
                buf.mul_(momentum).add_(gi)
            gi = gi + momentum * buf if nesterov else buf
        p.sub_(lr * gi)

    # ------------------------------------------------------------------ inference post-processing
    def seg_postprocess(self, logits, crop, out_hw, want_probs=False):
        """Plain-torch statement of csrc/postprocess.cu (cv2.resize INTER_LINEAR coordinates, fp32)."""
        c, h, w = logits.shape
        cy, cx, ch, cw = (int(v) for v in crop)
        oh, ow = int(out_hw[0]), int(out_hw[1])
        pr = torch.softmax(logits.float(), 0)[:, cy:cy + ch, cx:cx + cw]

        def coords(n_out, n_src):
            scale = torch.tensor(float(n_src) / float(n_out), dtype=torch.float64).float()
            f = (torch.arange(n_out, dtype=torch.float32) + 0.5) * scale - 0.5
            s = torch.floor(f)
            f = f - s
            s = s.long()
            lo = s < 0
            s = torch.where(lo, torch.zeros_like(s), s)
            f = torch.where(lo, torch.zeros_like(f), f)
            hi = s >= n_src - 1
            s = torch.where(hi, torch.full_like(s, n_src - 1), s)
            f = torch.where(hi, torch.zeros_like(f), f)
            return s, torch.clamp(s + 1, max=n_src - 1), f

        y0, y1, fy = coords(oh, ch)
        x0, x1, fx = coords(ow, cw)
        top = pr[:, y0][:, :, x0] * (1 - fx) + pr[:, y0][:, :, x1] * fx
        bot = pr[:, y1][:, :, x0] * (1 - fx) + pr[:, y1][:, :, x1] * fx
        out = top * (1 - fy)[None, :, None] + bot * fy[None, :, None]
        cls = out.argmax(0).to(torch.uint8)
        return cls, (out.permute(1, 2, 0).contiguous() if want_probs else None)

    def confusion_matrix(self, pred, gt, classes, hist=None):
        if hist is None:
            hist = torch.zeros((classes, classes), dtype=torch.int64)
        a, b = gt.reshape(-1).long(), pred.reshape(-1).long()
        k = (a < classes) & (b < classes)
        hist += torch.bincount(classes * a[k] + b[k], minlength=classes * classes).reshape(classes, classes)
        return hist
