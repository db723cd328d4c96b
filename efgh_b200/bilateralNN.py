"""BilateralConvFlex - drop-in for reference nets/bilateralNN.py:55-263 on hand-written CUDA.

Same constructor arguments, same forward signature, same parameter / buffer names
(`feat_indices`, `out_indices`, `blur_conv.{0,2,..}.{weight,bias}`, `bias`), so reference checkpoints
load with strict=True (reference main.py:136) and reference nets/enet.py:30-141 can use it unchanged.

What differs underneath: splat is one vector-atomic scatter into a vertex-major
matrix instead of two sparse COO coalesces of a (4N, C) temporary; the (1, C, F, H) gathered tensor
is never materialised - neighbour rows are copied straight into the tensor-core convolution's operand
pipeline; the density normalisation is one in-place pass over the splat matrix.  Forward and backward both
run through the C ABI in include/efgh_b200.h (the data gradient as a gather-form convolution on the same
tensor-core kernel); there is no PyTorch fallback.  (Whole scans / batches without per-call overheads:
efgh_b200.pipeline.ScanPipeline.)

Only batch size 1 is meaningful in the reference (bilateralNN.py:162-165: "batch size can only be 1
for now"; the splat indices carry no batch offset), so B != 1 raises here instead of silently mixing
the batch entries.
"""
import os
import weakref

import torch
import torch.nn as nn

from . import _capi
from .generate_data import blur_offsets

DELETE_TMP_VARIABLES = False
# Test hook: when set to a dict, every forward stores its convolutions' activations (post-ReLU, vertex-major) under "ys" -
# the backward parity tests need the ReLU active set the kernels actually used (tests/test_gpu_backward.py).
DEBUG_KEEP = None

_ACT = {"none": 0, "relu": 1, "leaky": 2}

# Arithmetic of the lattice convolution's contraction (forward):
#   "3xtf32" tcgen05 tensor cores with 3xTF32 operand splitting - fp32-equivalent, the default;
#   "tf32"   tcgen05, one TF32 pass - what cuDNN gives the reference under torch's default allow_tf32;
#   "fp32"   CUDA-core FFMA kernel (also the automatic choice for shapes the tensor-core kernel rejects).
CONV_PRECISION = os.environ.get("EFGH_CONV_PRECISION", "3xtf32")
_NSPLIT = {"3xtf32": 3, "tf32": 1}
# Data gradient of the convolutions as a gather-form convolution on the tensor-core kernel (else the scatter-form
# CUDA-core kernel).
DGRAD_ON_TENSOR_CORES = os.environ.get("EFGH_DGRAD", "tc") != "scatter"
# Weight gradient on the tensor cores (tcgen05, 3xTF32, MN-major operands; csrc/wgrad_tc.cu), else the fp32 CUDA-core kernel.
WGRAD_ON_TENSOR_CORES = os.environ.get("EFGH_WGRAD", "tc") != "ffma"


def init_weights(m):
    """reference nets/bilateralNN.py:42-53"""
    if isinstance(m, (nn.Conv2d, nn.Linear, nn.ConvTranspose2d)):
        m.weight.data.normal_(0, 1e-3)
        if m.bias is not None:
            m.bias.data.zero_()
        invalidate_weight_cache(m.weight)                # `.data` edits do not bump the parameter's version
    elif isinstance(m, nn.BatchNorm2d):
        m.weight.data.fill_(1)
        m.bias.data.zero_()


def _idx_bits(t):
    if t.dtype == torch.int64:
        return 64
    if t.dtype == torch.int32:
        return 32
    raise TypeError("lattice indices must be int64 or int32, got %s" % t.dtype)


def _rows(t):
    """(1, R, n) index / weight tensor -> (tensor, leading dimension) with unit stride along n."""
    t = t[0]
    if t.shape[-1] > 1 and t.stride(-1) != 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[-1], 1)
    return t, max(ld, 1)


# ---- thin wrappers over the C ABI -----------------------------------------------------------------

def scatter(feat_cn, w, off, row_shift, rows, want_wsum):
    """feat_cn (C,n) any strides; w (1,4,n); off (1,4,n).  Returns S (rows, C) [, wsum (rows,)]."""
    C, n = feat_cn.shape
    S = torch.zeros((rows, C), dtype=torch.float32, device=feat_cn.device)
    wsum = torch.zeros((rows,), dtype=torch.float32, device=feat_cn.device) if want_wsum else None
    w2, w_ld = _rows(w)
    o2, o_ld = _rows(off)
    _capi.check(_capi.lib().efgh_bcl_scatter(feat_cn.data_ptr(), feat_cn.stride(0), feat_cn.stride(1), C, None, 0, 0, 0, n, None,
                                             w2.data_ptr(), w_ld, o2.data_ptr(), _idx_bits(o2), o_ld, row_shift,
                                             S.data_ptr(), C, _capi.ptr(wsum), _capi.stream_ptr()),
                "efgh_bcl_scatter")
    return S, wsum


def normalize_(S, wsum):
    """S[r, :] *= 1/(wsum[r] + 1e-5) in place; returns the factors (needed by the backward pass)."""
    inv = torch.empty_like(wsum)
    _capi.check(_capi.lib().efgh_bcl_normalize(S.data_ptr(), S.stride(0), S.shape[1], wsum.data_ptr(), inv.data_ptr(),
                                               wsum.numel(), None, 0, _capi.stream_ptr()), "efgh_bcl_normalize")
    return inv


def gather(Z, row_scale, w, off, row_shift, bias, n):
    """Z (rows, C) vertex-major -> (C, n) channel-major (the reference's layout)."""
    C = Z.shape[1]
    out = torch.empty((C, n), dtype=torch.float32, device=Z.device)
    w2, w_ld = _rows(w)
    o2, o_ld = _rows(off)
    _capi.check(_capi.lib().efgh_bcl_gather(Z.data_ptr(), Z.stride(0), C, _capi.ptr(row_scale), n, None,
                                            w2.data_ptr(), w_ld, o2.data_ptr(), _idx_bits(o2), o_ld, row_shift,
                                            _capi.ptr(bias), out.data_ptr(), out.stride(0), 1, _capi.stream_ptr()),
                "efgh_bcl_gather")
    return out


def conv(X, nbr, Wt, bias, act, h, wkey=None):
    """X (rows, C) (already normalised); nbr (1,F,h) or None; Wt (F*C, M) -> Y (h, M).
    wkey: the parameter Wt was derived from (the packed tensor-core image is cached on it)."""
    C = X.shape[1]
    M = Wt.shape[1]
    Y = torch.empty((h, M), dtype=torch.float32, device=X.device)
    if nbr is not None:
        nb2, nb_ld = _rows(nbr)
        F, bits, nbp = nb2.shape[0], _idx_bits(nb2), nb2.data_ptr()
    else:
        F, bits, nbp, nb_ld = 1, 64, None, 0
    L = _capi.lib()
    nsplit = _NSPLIT.get(CONV_PRECISION, 0)
    if (nsplit and L.efgh_bcl_conv_tc_supported(C, F, M, nsplit) and X.stride(0) % 4 == 0 and X.data_ptr() % 16 == 0
            and (bias is None or bias.data_ptr() % 16 == 0)):
        K = Wt.shape[0]
        img = _cached("img", wkey if wkey is not None else Wt, nsplit, lambda: _tc_image(Wt, nsplit))
        split = L.efgh_bcl_conv_tc_groups(K, M) > 1     # wide output + long contraction: partial sums are added in L2
        if split:
            Y.zero_()
        _capi.check(L.efgh_bcl_conv_tc(X.data_ptr(), X.stride(0), C, None, 0, nbp, bits, nb_ld, F, h,
                                       None, img.data_ptr(), _capi.ptr(bias), M, act, Y.data_ptr(), M, nsplit,
                                       1 if split else 0, _capi.stream_ptr()), "efgh_bcl_conv_tc")
        if split:
            _capi.check(L.efgh_bcl_bias_act(Y.data_ptr(), M, M, h, None, _capi.ptr(bias), act, _capi.stream_ptr()),
                        "efgh_bcl_bias_act")
        return Y
    _capi.check(L.efgh_bcl_conv(X.data_ptr(), X.stride(0), C, None, nbp, bits, nb_ld, F, h,
                                None, Wt.data_ptr(), _capi.ptr(bias), M, act, Y.data_ptr(), M, 0,
                                _capi.stream_ptr()), "efgh_bcl_conv")
    return Y


def conv_dgrad(dY, act_out, act, nbr, Wt, C, rows):
    h, M = dY.shape
    if nbr is not None:
        nb2, nb_ld = _rows(nbr)
        F, bits, nbp = nb2.shape[0], _idx_bits(nb2), nb2.data_ptr()
        dX = torch.zeros((rows, C), dtype=torch.float32, device=dY.device)
    else:
        F, bits, nbp, nb_ld = 1, 64, None, 0
        dX = torch.empty((rows, C), dtype=torch.float32, device=dY.device)
    _capi.check(_capi.lib().efgh_bcl_conv_dgrad(dY.data_ptr(), dY.stride(0), _capi.ptr(act_out),
                                                act_out.stride(0) if act_out is not None else 0, act, M, nbp, bits,
                                                nb_ld, F, h, None, Wt.data_ptr(), C, dX.data_ptr(), C,
                                                _capi.stream_ptr()), "efgh_bcl_conv_dgrad")
    return dX


def _act_bwd(dY, act_out, act, out=None):
    """dY * act'(Y) (the activation's derivative is read off its output, as the kernels' loaders do)."""
    if act == _ACT["relu"]:
        return torch.mul(dY, act_out > 0, out=out)
    if act == _ACT["leaky"]:
        return torch.where(act_out > 0, dY, 0.1 * dY, out=out)
    if out is None:
        return dY
    out.copy_(dY)
    return out


def _tc_image(Wt, nsplit):
    K, M = Wt.shape
    L = _capi.lib()
    img = torch.empty(L.efgh_bcl_packed_weight_bytes(K, M, nsplit) // 4, dtype=torch.float32, device=Wt.device)
    _capi.check(L.efgh_bcl_pack_weights(Wt.data_ptr(), K, M, nsplit, img.data_ptr(), _capi.stream_ptr()),
                "efgh_bcl_pack_weights")
    return img


def neighbours_symmetric(nbr, mirror):
    """True when `g = nbr[f, h] >= 0` implies `nbr[mirror[f], g] == h` for every tap - the property that lets
    the data gradient of the lattice convolution be computed as a gather over the same neighbour table.  It
    holds whenever no neighbour key left the hash's key box (the reference's key2int aliasing,
    nets/transforms.py:124-131, is the only way to break it).  One device reduction + one host read."""
    nb = nbr[0]
    H = nb.shape[1]
    back = torch.gather(nb[list(mirror)], 1, nb.clamp(min=0).long())
    here = torch.arange(H, device=nb.device, dtype=back.dtype)[None]
    return bool(((back == here) | (nb < 0)).all())


def conv_dgrad_tc(dY, act_out, act, nbr, W, mirror, rows, nsplit, symmetric=None):
    """Data gradient of one convolution on the tensor-core kernel, or None when the shape does not fit it.

    The scatter form dX[nbr[f,h]+1, c] += sum_m dY[h,m] W[m,c,f] becomes, through the lattice's symmetry
    (tap f of h is g  <=>  tap mirror(f) of g is h), the forward-shaped gather
        dX[g+1, c] = sum_t sum_m dYm[nbr[t,g]+1, m] * W[m, c, mirror(t)],
    i.e. efgh_bcl_conv_tc over the same neighbour table with re-laid weights.  The kernel wants an output
    width that is a multiple of 32, so the first C % 32 channels (the four el_minus_gr channels of E-Net) stay
    on the FFMA scatter kernel."""
    L = _capi.lib()
    h, Mk = dY.shape
    C = W.shape[1]
    if nbr is None:
        if not L.efgh_bcl_conv_tc_supported(Mk, 1, C, nsplit):
            return None
        dYm = _act_bwd(dY, act_out, act)
        if not dYm.is_contiguous():
            dYm = dYm.contiguous()
        img = _tc_image(W[:, :, 0, 0].contiguous(), nsplit)          # (K = M_k, C)
        split = L.efgh_bcl_conv_tc_groups(Mk, C) > 1
        dX = (torch.zeros if split else torch.empty)((h, C), dtype=torch.float32, device=dY.device)
        _capi.check(L.efgh_bcl_conv_tc(dYm.data_ptr(), Mk, Mk, None, 0, None, 64, 0, 1, h, None, img.data_ptr(), None, C,
                                       0, dX.data_ptr(), C, nsplit, 1 if split else 0, _capi.stream_ptr()),
                    "efgh_bcl_conv_tc(dgrad)")
        return dX
    rem = C % 32
    Cg = C - rem
    nb2, nb_ld = _rows(nbr)
    F = nb2.shape[0]
    if rem % 4 or Cg == 0 or not L.efgh_bcl_conv_tc_supported(Mk, F, Cg, nsplit):
        return None
    if symmetric is None:
        symmetric = neighbours_symmetric(nbr, mirror)
    if not symmetric:
        return None
    Xp = torch.empty((h + 1, Mk), dtype=torch.float32, device=dY.device)
    Xp[0].zero_()                                                     # the sink row: absent neighbours
    _act_bwd(dY, act_out, act, out=Xp[1:])
    Wg = W[:, rem:, :, 0][:, :, list(mirror)].permute(2, 0, 1).reshape(F * Mk, Cg).contiguous()
    img = _tc_image(Wg, nsplit)
    dX = torch.zeros((rows, C), dtype=torch.float32, device=dY.device)
    split = L.efgh_bcl_conv_tc_groups(F * Mk, Cg) > 1
    out = dX[1:, rem:]
    _capi.check(L.efgh_bcl_conv_tc(Xp.data_ptr(), Mk, Mk, None, 0, nb2.data_ptr(), _idx_bits(nb2), nb_ld, F, h, None,
                                   img.data_ptr(), None, Cg, 0, out.data_ptr(), C, nsplit, 1 if split else 0,
                                   _capi.stream_ptr()), "efgh_bcl_conv_tc(dgrad)")
    if rem:
        Wr = W[:, :rem, :, 0].permute(2, 1, 0).reshape(F * rem, Mk).contiguous()       # (F*rem, M) for the scatter form
        _capi.check(L.efgh_bcl_conv_dgrad(dY.data_ptr(), dY.stride(0), _capi.ptr(act_out),
                                          act_out.stride(0) if act_out is not None else 0, act, Mk, nb2.data_ptr(),
                                          _idx_bits(nb2), nb_ld, F, h, None, Wr.data_ptr(), rem, dX.data_ptr(), C,
                                          _capi.stream_ptr()), "efgh_bcl_conv_dgrad")
    return dX


def conv_wgrad(X, row_scale, nbr, dY, act_out, act, want_bias):
    h, M = dY.shape
    C = X.shape[1]
    if nbr is not None:
        nb2, nb_ld = _rows(nbr)
        F, bits, nbp = nb2.shape[0], _idx_bits(nb2), nb2.data_ptr()
    else:
        F, bits, nbp, nb_ld = 1, 64, None, 0
    dWt = torch.zeros((F * C, M), dtype=torch.float32, device=dY.device)
    db = torch.zeros((M,), dtype=torch.float32, device=dY.device) if want_bias else None
    L = _capi.lib()
    if (WGRAD_ON_TENSOR_CORES and row_scale is None and L.efgh_bcl_conv_wgrad_tc_supported(C, F, M) and X.stride(0) % 4 == 0
            and X.data_ptr() % 16 == 0):
        dYm = _act_bwd(dY, act_out, act)                   # the tensor-core kernel takes the masked gradient
        if not dYm.is_contiguous() or dYm.data_ptr() % 16:
            dYm = dYm.contiguous()
        _capi.check(L.efgh_bcl_conv_wgrad_tc(X.data_ptr(), X.stride(0), C, nbp, bits, nb_ld, F, h, None, dYm.data_ptr(), dYm.stride(0), M,
                                             dWt.data_ptr(), _capi.ptr(db), _capi.stream_ptr()), "efgh_bcl_conv_wgrad_tc")
        return dWt, db
    _capi.check(_capi.lib().efgh_bcl_conv_wgrad(X.data_ptr(), X.stride(0), C, _capi.ptr(row_scale), nbp, bits, nb_ld, F,
                                                h, None, dY.data_ptr(), dY.stride(0), _capi.ptr(act_out),
                                                act_out.stride(0) if act_out is not None else 0, act, M,
                                                dWt.data_ptr(), _capi.ptr(db), _capi.stream_ptr()),
                "efgh_bcl_conv_wgrad")
    return dWt, db


# Re-laid weights ((K, M) matrices and the tensor-core kernel's packed images) are cached per parameter OBJECT in a
# module-level table (not on the parameter: attributes of a Parameter travel with torch.save(model)).  An entry is
# valid while (version, storage address, device, dtype) are unchanged: torch bumps `_version` on every in-place update
# through the tensor itself (optimizer step, load_state_dict, copy_), and `.to(device)` / `.half()` change the address.
# Edits through `.data` (init_weights below, the reference's initialisers, EMA / clipping code) change none of those,
# so (a) a forward that autograd records - training - drops its parameters' entries first (BilateralConvFlex.forward):
# re-laying the weights once per step is small next to the convolution; (b) init_weights invalidates explicitly;
# (c) callers that edit `.data` at inference time call invalidate_weight_cache() (BilateralConvFlex.invalidate_cache()).
_WEIGHT_CACHE = {}      # id(parameter) -> (weakref, state key, {(kind, extra): tensor})


def invalidate_weight_cache(*params):
    """Drop the cached re-laid / packed weights of `params` (all parameters when called without arguments)."""
    if not params:
        _WEIGHT_CACHE.clear()
        return
    for p in params:
        _WEIGHT_CACHE.pop(id(p), None)


def _cached(kind, W, extra, build):
    state = (W._version, W.data_ptr(), W.device, W.dtype)
    ent = _WEIGHT_CACHE.get(id(W))
    if ent is None or ent[0]() is not W or ent[1] != state:
        wid = id(W)
        try:
            ref = weakref.ref(W, lambda _r, wid=wid: _WEIGHT_CACHE.pop(wid, None))
        except TypeError:                                # not weak-referenceable: just rebuild every call
            return build()
        ent = (ref, state, {})
        _WEIGHT_CACHE[wid] = ent
    d = ent[2]
    key = (kind, extra)
    if key not in d:
        d[key] = build()
    return d[key]


def _wt_first(W):
    """Conv2d weight (M, C, F, 1) -> (F*C, M) row-major, k = f*C + c."""
    M, C, F, _ = W.shape
    return _cached("first", W, None, lambda: W.detach()[:, :, :, 0].permute(2, 1, 0).reshape(F * C, M).contiguous())


def _wt_point(W):
    """Conv2d 1x1 weight (M, C, 1, 1) -> (C, M)."""
    return _cached("point", W, None, lambda: W.detach()[:, :, 0, 0].t().contiguous())


class _BCLFunction(torch.autograd.Function):
    """splat -> [norm] -> gather-conv -> (ReLU -> 1x1 conv)* -> [act] -> [slice + bias]"""

    @staticmethod
    def forward(ctx, cfg, features, in_bary, in_off, nbr, out_bary, out_off, slice_bias, *wb):
        do_splat, do_slice, use_norm, final_act = cfg[:4]
        if features.shape[0] != 1:
            raise ValueError("BilateralConvFlex: batch size must be 1 (reference bilateralNN.py:162-165)")
        feat = features[0]
        if feat.dtype != torch.float32:
            feat = feat.float()
        H = nbr.shape[-1]
        with torch.cuda.device(feat.device):
            inv = None
            if do_splat:
                S, wsum = scatter(feat, in_bary, in_off, 1, H + 1, use_norm)
                if use_norm:
                    inv = normalize_(S, wsum)      # S now holds the normalised splat (bilateralNN.py:210-211)
            else:
                S = torch.zeros((H + 1, feat.shape[0]), dtype=torch.float32, device=feat.device)
                S[1:] = feat.t()
            nconv = len(wb) // 2
            xs, ys, wts, acts = [], [], [], []
            X, nb = S, nbr
            for k in range(nconv):
                W, b = wb[2 * k], wb[2 * k + 1]
                Wt = _wt_first(W) if k == 0 else _wt_point(W)
                act = _ACT["relu"] if k < nconv - 1 else final_act
                Y = conv(X, nb, Wt, b, act, H, wkey=W)
                xs.append(X); ys.append(Y); wts.append(Wt); acts.append(act)
                X, nb = Y, None
            if DEBUG_KEEP is not None:
                DEBUG_KEEP["ys"] = ys
            if do_slice:
                n_out = out_bary.shape[-1]
                out = gather(X, None, out_bary, out_off, 0, slice_bias, n_out)[None]
            else:
                out = X.t()[None]
        ctx.cfg = cfg
        ctx.n_in = feat.shape[-1]
        ctx.acts = acts
        ctx.has_slice_bias = slice_bias is not None
        ctx.save_for_backward(in_bary, in_off, nbr, out_bary, out_off, inv, *xs, *ys, *wts, *wb[0::2])
        ctx.nconv = nconv
        ctx.nbr_symmetric = getattr(nbr, "_efgh_symmetric", None)     # set by GenerateData when it already knows
        return out

    @staticmethod
    def backward(ctx, gout):
        do_splat, do_slice, use_norm, final_act, mirror = ctx.cfg
        saved = ctx.saved_tensors
        in_bary, in_off, nbr, out_bary, out_off, inv = saved[:6]
        nconv = ctx.nconv
        xs = saved[6:6 + nconv]
        ys = saved[6 + nconv:6 + 2 * nconv]
        wts = saved[6 + 2 * nconv:6 + 3 * nconv]
        Ws = saved[6 + 3 * nconv:6 + 4 * nconv]
        nsplit = _NSPLIT.get(CONV_PRECISION, 0)
        H = nbr.shape[-1]
        g = gout[0]
        if g.dtype != torch.float32:
            g = g.float()
        grads_wb = [None] * (2 * nconv)
        d_slice_bias = None
        with torch.cuda.device(g.device):
            if do_slice:
                dY, _ = scatter(g, out_bary, out_off, 0, H, False)         # adjoint of slice
                if ctx.has_slice_bias and ctx.needs_input_grad[7]:
                    d_slice_bias = g.sum(dim=1)
            else:
                dY = g.t().contiguous()                                    # (H, C_out)
            for k in range(nconv - 1, -1, -1):
                act = ctx.acts[k]
                act_out = ys[k] if act != 0 else None
                X = xs[k]
                first = k == 0
                need_w = ctx.needs_input_grad[8 + 2 * k]
                need_b = ctx.needs_input_grad[8 + 2 * k + 1]
                if need_w or need_b:
                    dWt, db = conv_wgrad(X, None, nbr if first else None, dY, act_out, act, need_b)
                    if need_w:
                        if first:
                            F = nbr.shape[1]
                            C = X.shape[1]
                            grads_wb[0] = dWt.view(F, C, -1).permute(2, 1, 0).unsqueeze(-1)
                        else:
                            grads_wb[2 * k] = dWt.t()[:, :, None, None]
                    if need_b:
                        grads_wb[2 * k + 1] = db
                if first and not ctx.needs_input_grad[1]:
                    dY = None
                    break
                dX = None
                if nsplit and DGRAD_ON_TENSOR_CORES:
                    dX = conv_dgrad_tc(dY, act_out, act, nbr if first else None, Ws[k], mirror, X.shape[0], nsplit,
                                       symmetric=ctx.nbr_symmetric)
                if dX is None:
                    dX = conv_dgrad(dY, act_out, act, nbr if first else None, wts[k], X.shape[1], X.shape[0])
                dY = dX
            dfeat = None
            if ctx.needs_input_grad[1] and dY is not None:
                if do_splat:
                    dfeat = gather(dY, inv, in_bary, in_off, 1, None, ctx.n_in)[None]   # adjoint of splat (+norm)
                else:
                    dfeat = dY[1:].t()[None]
        return (None, dfeat, None, None, None, None, None, d_slice_bias, *grads_wb)


class BilateralConvFlex(nn.Module):
    def __init__(self,
                 d, neighborhood_size,
                 num_input, num_output,
                 DEVICE,
                 use_bias,
                 use_leaky,
                 use_norm,
                 do_splat,
                 do_slice,
                 last_relu,
                 chunk_size=1024 * 1024 * 25):
        """Arguments as reference nets/bilateralNN.py:56-80.  `chunk_size` is accepted and ignored: the
        fused gather-convolution never materialises the (C, F, H) tensor that chunking bounded."""
        super(BilateralConvFlex, self).__init__()
        if d != 3:
            raise NotImplementedError("efgh_b200 implements the d=3 lattice only")
        self.d = d
        self.d1 = d + 1
        self.neighborhood_size = neighborhood_size
        self.filter_size = self.get_filter_size()
        offs = [tuple(o) for o in blur_offsets(neighborhood_size, d).tolist()]
        self._mirror = tuple(offs.index(tuple(-v for v in o)) for o in offs)     # tap f <-> the tap with the negated offset
        self.num_input = num_input
        self.num_output = num_output
        self.DEVICE = DEVICE
        self.use_bias = use_bias
        self.use_leaky = use_leaky
        self.do_splat = do_splat
        self.do_slice = do_slice
        self.last_relu = last_relu
        self.use_norm = use_norm
        self.MAX_SIZE = chunk_size

        num_final_output = num_output[-1]
        self.register_buffer('feat_indices', torch.arange(num_input, dtype=torch.long))
        if self.do_slice:
            self.register_buffer('out_indices', torch.arange(num_final_output, dtype=torch.long))

        # Same module tree as the reference (bilateralNN.py:103-135) so state_dict keys match; the Conv2d
        # modules only hold the parameters - forward() feeds them to the CUDA kernels.
        layers = []
        n_in = num_input
        for idx, n_out in enumerate(num_output[:-1]):
            layers.append(nn.Conv2d(n_in, n_out, kernel_size=(self.filter_size, 1) if idx == 0 else (1, 1),
                                    stride=1, padding=0, bias=True))
            layers.append(nn.ReLU(inplace=False))
            n_in = n_out
        layers.append(nn.Conv2d(n_in, num_final_output,
                                kernel_size=(self.filter_size, 1) if len(num_output) == 1 else (1, 1),
                                stride=1, padding=0, bias=True))
        if self.last_relu:
            layers.append(nn.LeakyReLU(0.1, inplace=False) if use_leaky else nn.ReLU(inplace=False))
        self.blur_conv = nn.Sequential(*layers)
        for m in self.blur_conv.modules():
            init_weights(m)

        if self.do_slice and self.use_bias:
            self.register_parameter('bias', nn.Parameter(data=torch.zeros((num_final_output,), dtype=torch.float32),
                                                         requires_grad=True))

    def get_filter_size(self):
        return (self.neighborhood_size + 1) ** self.d1 - self.neighborhood_size ** self.d1

    def invalidate_cache(self):
        """Call after editing weights through `.data` under no_grad (those edits are invisible to the version counter)."""
        invalidate_weight_cache(*[p for p in self.parameters()])

    def forward(self, features,
                in_barycentric, in_lattice_offset,
                blur_neighbors,
                out_barycentric, out_lattice_offset):
        """Shapes as reference nets/bilateralNN.py:152-160.  Returns (1, C_out, N_out) when slicing, else
        (1, C_out, H) - the latter as a transposed view of the vertex-major result."""
        if not features.is_cuda:
            raise _capi.EfghError("BilateralConvFlex runs on CUDA tensors only; there is no CPU path")
        final_act = 0
        if self.last_relu:
            final_act = _ACT["leaky"] if self.use_leaky else _ACT["relu"]
        cfg = (self.do_splat, self.do_slice, self.use_norm, final_act, self._mirror)
        wb = []
        for m in self.blur_conv:
            if isinstance(m, nn.Conv2d):
                wb += [m.weight, m.bias]
        slice_bias = self.bias if (self.do_slice and self.use_bias) else None
        if torch.is_grad_enabled() and any(p.requires_grad for p in wb):
            invalidate_weight_cache(*wb)                 # training: weights change between calls, possibly through `.data`
        return _BCLFunction.apply(cfg, features, in_barycentric, in_lattice_offset, blur_neighbors,
                                  out_barycentric if self.do_slice else None,
                                  out_lattice_offset if self.do_slice else None, slice_bias, *wb)
