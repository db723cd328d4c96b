"""Torch-CPU restatement of the bilateral convolution layer - TEST INFRASTRUCTURE ONLY.

Follows reference nets/bilateralNN.py:148-263 (BilateralConvFlex.forward) for batch size 1, which is the
only batch size the reference supports (bilateralNN.py:162-165).  Written as a pure function of tensors
so torch autograd yields the oracle gradients (reference row a17: autograd over splat / gather / conv /
slice).  dtype float64 gives a tight "truth" for tolerance tests; float32 mimics the reference's own
arithmetic.  Parity status: PINNED against the live reference (tests/golden/bcl_*.npz).
"""
import torch


def bcl_forward(features, in_bary, in_off, nbrs, convs, *, use_norm=True, do_splat=True,
                do_slice=False, out_bary=None, out_off=None, slice_bias=None,
                last_relu=False, use_leaky=True, dtype=None, relu_masks=None, pre_acts=None):
    """features (1,C_in,N) | (1,C_in,H) if not do_splat; in_bary (1,4,N); in_off (1,4,N) int64;
    nbrs (1,F,H) int64 with -1 = absent; convs = [(W0 (C1,C_in,F,1), b0), (W1 (C2,C1,1,1), b1), ...].
    Returns (1,C_out,H), or (1,C_out,N_out) when slicing.

    Test hooks (not part of the reference): `pre_acts`, a list that receives the pre-activation (H, C) of every
    inter-convolution ReLU; `relu_masks`, a list of (H, C) bool tensors used INSTEAD of `y > 0` - a ReLU's
    derivative is discontinuous at 0, so a backward comparison must run both sides with the same active set
    (the tests assert that the two active sets differ only where the pre-activation is within the forward tolerance
    of zero)."""
    dt = dtype or features.dtype
    feat = features[0].to(dt)                       # (C,N)
    nb = nbrs[0]                                    # (F,H)
    H = nb.shape[-1]
    C = feat.shape[0]
    if do_splat:                                    # bilateralNN.py:176-191
        w = in_bary[0].to(dt)                       # (4,N)
        rows = (in_off[0] + 1).reshape(-1)          # (4N,) r-major then n, +1: row 0 is the sink
        contrib = (w[:, None, :] * feat[None, :, :]).permute(0, 2, 1).reshape(-1, C)   # (4N,C)
        S = torch.zeros(H + 1, C, dtype=dt).index_add(0, rows, contrib)
        if use_norm:                                # bilateralNN.py:193-211
            W = torch.zeros(H + 1, dtype=dt).index_add(0, rows, w.reshape(-1))
            S = S * (1.0 / (W + 1e-5))[:, None]
    else:                                           # bilateralNN.py:215-221
        S = torch.cat([torch.zeros(1, C, dtype=dt), feat.t()], 0)
    X = S[nb + 1]                                   # (F,H,C)  bilateralNN.py:240-242
    W0, b0 = convs[0]
    y = torch.einsum("fhc,mcf->hm", X, W0[..., 0].to(dt)) + b0.to(dt)   # Conv2d (F,1) :244
    for li, (Wk, bk) in enumerate(convs[1:]):       # ReLU between convs, :111-115
        if pre_acts is not None:
            pre_acts.append(y.detach())
        y = torch.relu(y) if relu_masks is None else y * relu_masks[li].to(dt)
        y = y @ Wk[:, :, 0, 0].to(dt).t() + bk.to(dt)
    if last_relu:                                   # :121-134
        y = torch.nn.functional.leaky_relu(y, 0.1) if use_leaky else torch.relu(y)
    if not do_slice:
        return y.t()[None]                          # (1,C_out,H) :248-249
    ob = out_bary[0].to(dt)                         # (4,N_out)  :251-257 (no +1 here)
    out = (ob[:, :, None] * y[out_off[0]]).sum(0)   # (N_out,C_out)
    if slice_bias is not None:                      # :260-261
        out = out + slice_bias.to(dt)
    return out.t()[None]


def stem_forward(pc, layers, leaky=True, dtype=torch.float64):
    """E-Net stem, reference nets/enet.py:24-28,111 + nets/net_utils.py:35-43: three pointwise Conv1d(k=1), each
    followed by LeakyReLU(0.1) (use_leaky) or ReLU, on the UNSCALED cloud.  pc (3, N); layers [(W (out,in[,1]), b)]*3.
    Returns (c3, N)."""
    x = torch.as_tensor(pc).to(dtype)
    for W, b in layers:
        W = torch.as_tensor(W).to(dtype)
        W = W[:, :, 0] if W.dim() == 3 else W
        x = W @ x + torch.as_tensor(b).to(dtype)[:, None]
        x = torch.where(x > 0, x, 0.1 * x) if leaky else torch.clamp(x, min=0)
    return x
