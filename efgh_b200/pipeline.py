"""ScanPipeline - whole-scan forward (lattice build + all BCL layers) enqueued without a host round trip.

This is the path BASELINE.json's metric is measured on: for one LiDAR scan it does what
reference nets/enet.py:107-141 does between `generate_data(pc)` and `bcn5(...)` - the five-level lattice
build (reference nets/generate_data.py:117-193) followed by the five BilateralConvFlex layers, each fed
`cat(el_minus_gr_l, previous output)` - but

  * every buffer is pre-allocated at capacity and every count (points per level, vertices per level) stays
    in device memory, so the ~60 kernels of a scan are enqueued back to back on one stream;
  * the `torch.cat` of reference enet.py:113-137 is never materialised: el_minus_gr and the previous
    level's output are splatted into column ranges of the same vertex-major matrix;
  * level l's output (vertex-major) IS level l+1's point-major feature matrix, so nothing is transposed;
  * the reference-format int64 tensors (pc1_lattice_offset, pc1_blur_neighbors) are still produced - they
    are part of the lattice build's contract - next to int32 copies that the BCL kernels read.

Several pipelines on different CUDA streams run concurrently; scans are independent, which is
also how they shard across GPUs.

`batch=B` ("ragged batched lattices"): B scans go through ONE launch sequence.  Their point
streams are concatenated, every scan keeps its own hash table / key box / insertion order
(efgh_lattice_*_batch), vertex indices are global, and the splat / convolution kernels see one big lattice.
The ~55 launches and, more importantly, the latency-bound small kernels of the coarse levels (a few thousand
vertices each) are paid once per batch instead of once per scan.
"""
import functools

import numpy as np
import torch

from . import _capi
from .generate_data import GenerateData, STATE_WORDS, check_status, VertexCapExceeded

_ACT = {"none": 0, "relu": 1, "leaky": 2}


def _on_own_device(fn):
    """The C ABI launches on the CURRENT CUDA device; streams and buffers belong to self.dev."""
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        with torch.cuda.device(self.dev):
            return fn(self, *a, **k)
    return wrapped


class ScanPipeline(object):
    def __init__(self, n_points, scales_filter_map, bcl_plan, weights, device, stem_channels=32,
                 vertex_cap_factor=1.0, emit_int64=True, last_relu=False, use_leaky=True, use_norm=True,
                 precision="3xtf32", batch=1, gather_splat=True, stem=None):
        """bcl_plan: [(C_in, [C_mid, C_out]), ...] one entry per level (reference nets/enet.py:30-83);
        weights: per level [(W0 (C_mid,C_in,F,1), b0), (W1 (C_out,C_mid,1,1), b1)] torch tensors;
        vertex_cap_factor: capacity of every vertex-side buffer as a multiple of n_points;
        precision: "3xtf32" (tcgen05, fp32-equivalent), "tf32" (tcgen05, one pass) or "fp32" (CUDA cores);
        gather_splat: levels >= 1 (whose input features are the previous level's point-major output rows) splat
        through the vertex -> contributions lists of the lattice build - no atomics, zero-fill and normalisation
        fused.  Level 0 (channel-major (C, N) input, working set beyond the L2 in batch mode) keeps the atomic scatter;
        stem: None, or ([(W1, b1), (W2, b2), (W3, b3)], use_leaky) - E-Net's pointwise `conv_in` (reference
        nets/enet.py:24-28; W as Conv1d weights (out, in, 1)): the level-0 splat then COMPUTES the stem features from
        the cloud and enqueue() ignores feat0;
        batch: scans per launch sequence, each of n_points points (inputs are then (3, batch*n_points) /
        (C, batch*n_points), scan b in columns [b*n_points, (b+1)*n_points))."""
        self.dev = torch.device(device)
        self.L = _capi.lib()
        self.B = int(batch)
        assert 1 <= self.B <= 64
        self.gather_splat = bool(gather_splat)
        self.level0_splat = "vector-atomic scatter + normalise"
        self.stem = None
        if stem is not None:
            layers, leaky = stem
            assert len(layers) == 3
            dims = [int(layers[0][0].shape[1])] + [int(W.shape[0]) for W, _ in layers]
            assert dims[3] == stem_channels and all(int(W.shape[1]) == dims[i] for i, (W, _) in enumerate(layers))
            flat = torch.cat([t.detach().to(torch.float32).reshape(-1) for W, b in layers for t in (W, b)])
            assert flat.numel() == self.L.efgh_bcl_stem_weight_floats(*dims)
            self.stem = {"dims": dims, "w": flat.to(self.dev).contiguous(), "slope": 0.1 if leaky else 0.0}
        self.batch_api = self.B > 1 or self.gather_splat       # the batch entry points also serve a batch of one
        self.n_scan = int(n_points)
        self.n0 = int(n_points) * self.B
        self.smap = scales_filter_map
        self.plan = bcl_plan
        self.nlev = len(scales_filter_map)
        assert len(bcl_plan) == self.nlev and len(weights) == self.nlev
        self.emit_int64 = emit_int64
        self.final_act = 0 if not last_relu else (_ACT["leaky"] if use_leaky else _ACT["relu"])
        self.use_norm = use_norm
        self.precision = precision
        self.nsplit = {"3xtf32": 3, "tf32": 1, "fp32": 0}[precision]
        self._graphs = {}
        self.overlap_lattice = True
        self._lat_stream = None
        self.gd = GenerateData(3, scales_filter_map, "cuda")
        dev = self.dev
        f32, i32, i64 = torch.float32, torch.int32, torch.int64
        cap = max(int(vertex_cap_factor * self.n0), 1024)
        cap_scan = max(int(vertex_cap_factor * self.n_scan), 1024)     # one scan's share: sizes its hash table
        self.levels = []
        n_cap = self.n0
        n_cap_scan = self.n_scan
        prev_c = stem_channels
        with torch.cuda.device(dev):
            self.states = torch.zeros((self.nlev, STATE_WORDS), dtype=i32, device=dev)
            ws_bytes = 0
            for li, (scale, radius) in enumerate(scales_filter_map):
                cin, (cmid, cout) = bcl_plan[li]
                assert cin == prev_c + 4, "level %d: C_in must be 4 + previous C_out" % li
                assert radius != -1, "ScanPipeline needs a blur radius on every level"
                F = self.gd.get_filter_size(radius)
                h_cap = min(4 * n_cap, cap)
                lv = {
                    # one scan's hash table: 4 x its vertex capacity (load factor <= 0.25), never more than 8 per point.
                    # The tables of a whole batch should stay in the 126 MB L2 (8 scans x 8 MB); at 2 x capacity
                    # (load ~0.4) the longer probe sequences were measured to cost more than the footprint saves.
                    "table": int(self.L.efgh_lattice_table_entries(n_cap_scan, 2 * cap_scan)),
                    "info": torch.zeros(max(int(self.L.efgh_lattice_batch_info_ints(self.B)), 1), dtype=i32, device=dev),
                    "gs": self.gather_splat and li > 0,
                    "voff": torch.zeros(int(self.L.efgh_lattice_vertex_offsets_ints(h_cap)), dtype=i32, device=dev) if self.gather_splat and li > 0 else None,
                    "prow": torch.zeros((n_cap, 8), dtype=f32, device=dev) if self.gather_splat and li > 0 else None,
                    "contrib": torch.zeros(4 * n_cap, dtype=i32, device=dev) if self.gather_splat and li > 0 else None,
                    "n_cap": n_cap, "h_cap": h_cap, "F": F, "scale": float(scale), "cin": cin, "cmid": cmid, "cout": cout,
                    "divisor": float(np.float32(self.gd.expected_std * scale)),
                    "offs": torch.from_numpy(self.gd.radius2offset[radius].astype(np.int32)).to(dev),
                    "bary": torch.empty((4, n_cap), dtype=f32, device=dev),
                    "elmgr": torch.empty((4, n_cap), dtype=f32, device=dev),
                    "loff32": torch.empty((4, n_cap), dtype=i32, device=dev),
                    "nbr32": torch.empty((F, h_cap), dtype=i32, device=dev),
                    "loff64": torch.empty((4, n_cap), dtype=i64, device=dev) if emit_int64 else None,
                    "nbr64": torch.empty((F, h_cap), dtype=i64, device=dev) if emit_int64 else None,
                    "next": torch.empty((3, h_cap), dtype=f32, device=dev) if li != self.nlev - 1 else None,
                    "S": torch.empty((h_cap + 1, cin), dtype=f32, device=dev),
                    "wsum": torch.empty((h_cap + 1,), dtype=f32, device=dev),
                    "Y": torch.empty((h_cap, cmid), dtype=f32, device=dev),
                    "Z": torch.empty((h_cap, cout), dtype=f32, device=dev),
                }
                (W0, b0), (W1, b1) = weights[li]
                M0, C0, F0, _ = W0.shape
                assert (M0, C0, F0) == (cmid, cin, F)
                lv["Wt0"] = W0.detach().to(dev, f32)[:, :, :, 0].permute(2, 1, 0).reshape(F * cin, cmid).contiguous()
                lv["b0"] = b0.detach().to(dev, f32).contiguous()
                lv["Wt1"] = W1.detach().to(dev, f32)[:, :, 0, 0].t().contiguous()
                lv["b1"] = b1.detach().to(dev, f32).contiguous()
                lv["tc"] = bool(self.nsplit and self.L.efgh_bcl_conv_tc_supported(cin, F, cmid, self.nsplit)
                                and self.L.efgh_bcl_conv_tc_supported(cmid, 1, cout, self.nsplit)
                                and self.L.efgh_bcl_conv_tc_groups(cmid, cout) == 1)
                if lv["tc"]:
                    for nm, K, M in (("img0", F * cin, cmid), ("img1", cmid, cout)):
                        img = torch.empty(self.L.efgh_bcl_packed_weight_bytes(K, M, self.nsplit) // 4, dtype=f32, device=dev)
                        _capi.check(self.L.efgh_bcl_pack_weights(lv["Wt" + nm[-1]].data_ptr(), K, M, self.nsplit, img.data_ptr(),
                                                                 torch.cuda.current_stream(dev).cuda_stream), "efgh_bcl_pack_weights")
                        lv[nm] = img
                    lv["split0"] = self.L.efgh_bcl_conv_tc_groups(F * cin, cmid) > 1
                if self.batch_api:
                    ws_bytes = max(ws_bytes, self.L.efgh_lattice_batch_workspace_bytes(self.B, lv["table"], n_cap))
                else:
                    ws_bytes = max(ws_bytes, self.L.efgh_lattice_workspace_bytes(n_cap))
                self.levels.append(lv)
                n_cap = h_cap
                n_cap_scan = min(4 * n_cap_scan, cap_scan)
                prev_c = cout
            self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            self._starts0 = [b * self.n_scan for b in range(self.B + 1)]
            self.scan_start = torch.tensor(self._starts0, dtype=i32, device=dev)
            self._pc_dev = torch.empty((3, self.n0), dtype=f32, device=dev)
            self._feat_dev = torch.empty((stem_channels, self.n0), dtype=f32, device=dev)
        # per launch sequence: clear/points/assign, vertices, zero, splat (+ normalise | level-0 transpose), conv1, conv2
        self.launches_per_scan = self.nlev * (3 + 1 + 1 + 1 + 2) + sum(0 if lv["gs"] else 1 for lv in self.levels)

    @_on_own_device
    def set_scan_sizes(self, sizes):
        """Batched pipelines: ragged batch - scan b has sizes[b] (1 <= sizes[b] <= n_points) points, stored back to
        back in the (3, B*n_points) / (C, B*n_points) inputs.  Synchronising (rewrites a device array that enqueued
        work may still read); previously captured graphs stay valid (they read the same array)."""
        assert self.B > 1 and len(sizes) == self.B and all(1 <= int(v) <= self.n_scan for v in sizes)
        torch.cuda.synchronize(self.dev)
        self._starts0 = [0]
        for v in sizes:
            self._starts0.append(self._starts0[-1] + int(v))
        self.scan_start.copy_(torch.tensor(self._starts0, dtype=torch.int32))
        torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------------------------------
    @_on_own_device
    def enqueue(self, pc, feat0, stream=None, timers=None):
        """pc (3,N) f32, feat0 (C_stem,N) f32 device tensors.  Enqueues the whole scan on `stream` (default:
        current).  Returns the last level's output buffer Z (h_cap, C_out) - valid rows = states[-1, 1].
        timers: optional dict; stages whose name is a key (or every stage if "*" is a key) get a CUDA event
        pair appended to timers[name].

        The lattice build of level l+1 depends only on level l's vertices, not on level l's BCL, so (unless stages
        are being timed) the five lattice levels run on a private side stream and the BCL chain follows them on
        `stream` through one event per level: the ~0.2 ms lattice chain hides behind the BCL chain."""
        L, ck = self.L, _capi.check
        main = stream if stream is not None else torch.cuda.current_stream(self.dev)
        s = main.cuda_stream
        assert pc.shape[-1] == self.n0 and pc.stride(-1) == 1 and (self.stem is not None or feat0.stride(-1) == 1)
        ws, wsn = self.ws.data_ptr(), self.ws.numel()
        lat = None
        if self.overlap_lattice and timers is None:
            if self._lat_stream is None:
                self._lat_stream = torch.cuda.Stream(self.dev)
            lat = self._lat_stream
            ev0 = torch.cuda.Event()
            ev0.record(main)
            lat.wait_event(ev0)                       # inputs ready / previous scan of this pipeline finished
        s_lat = lat.cuda_stream if lat is not None else s

        def timed(name, fn):
            if timers is None or not (name in timers or "*" in timers):
                return fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st = stream if stream is not None else torch.cuda.current_stream(self.dev)
            a.record(st)
            fn()
            b.record(st)
            timers.setdefault(name, []).append((a, b))

        pts_ptr, pts_ld = pc.data_ptr(), pc.stride(0)
        if self.stem is not None:
            prev_ptr, prev_sc, prev_sn, prev_c = None, 0, 1, self.stem["dims"][3]
        else:
            prev_ptr, prev_sc, prev_sn, prev_c = feat0.data_ptr(), feat0.stride(0), 1, feat0.shape[0]
        n_dev = None
        seg = self.scan_start.data_ptr()                      # batched: point-stream boundaries of the level
        if self.batch_api:
            n_dev = seg + 4 * self.B                          # total points of a (possibly ragged) batch
        for li, lv in enumerate(self.levels):
            st = self.states[li].data_ptr()
            n_cap, h_cap, cin = lv["n_cap"], lv["h_cap"], lv["cin"]
            h_dev = st + 4           # &state.hash_cnt
            if self.batch_api:
                info = lv["info"].data_ptr()
                timed("L%d.points" % li, lambda: ck(L.efgh_lattice_points_batch(
                    pts_ptr, pts_ld, n_cap, seg, self.B, lv["table"], lv["scale"], lv["bary"].data_ptr(),
                    lv["elmgr"].data_ptr(), n_cap, h_cap, st, info, _capi.ptr(lv["voff"]), _capi.ptr(lv["prow"]), ws, wsn, s_lat),
                    "efgh_lattice_points_batch"))
                timed("L%d.vertices" % li, lambda: ck(L.efgh_lattice_vertices_batch(
                    n_cap, seg, self.B, lv["table"], _capi.ptr(lv["loff64"]), lv["loff32"].data_ptr(), n_cap,
                    lv["offs"].data_ptr(), lv["F"], h_cap, _capi.ptr(lv["nbr64"]), lv["nbr32"].data_ptr(), h_cap,
                    _capi.ptr(lv["next"]), h_cap, lv["divisor"], st, info, _capi.ptr(lv["contrib"]), ws, wsn, s_lat),
                    "efgh_lattice_vertices_batch"))
                seg = info                                    # vertex_start of this level = scan_start of the next
            else:
                timed("L%d.points" % li, lambda: ck(L.efgh_lattice_points(
                    pts_ptr, pts_ld, n_cap, n_dev, lv["scale"], lv["bary"].data_ptr(), lv["elmgr"].data_ptr(), n_cap, h_cap,
                    st, ws, wsn, s_lat), "efgh_lattice_points"))
                timed("L%d.vertices" % li, lambda: ck(L.efgh_lattice_vertices(
                    n_cap, _capi.ptr(lv["loff64"]), lv["loff32"].data_ptr(), n_cap, lv["offs"].data_ptr(), lv["F"], h_cap,
                    _capi.ptr(lv["nbr64"]), lv["nbr32"].data_ptr(), h_cap, _capi.ptr(lv["next"]), h_cap, lv["divisor"],
                    st, ws, wsn, s_lat), "efgh_lattice_vertices"))
            S = lv["S"].data_ptr()
            zero_y = lv["tc"] and lv["split0"]                # split-K accumulator of the tensor-core conv
            # zero-fill of this level's accumulators: on the lattice stream, i.e. off the BCL chain's critical path
            gs = lv["gs"]                                     # the gather-form splat writes every row of S itself
            timed("L%d.zero" % li, lambda: ck(L.efgh_bcl_zero(None if gs else S, cin, cin, None if gs else lv["wsum"].data_ptr(),
                                                              lv["Y"].data_ptr() if zero_y else None,
                                                              lv["cmid"], lv["cmid"], h_cap, h_dev, 1, s_lat), "efgh_bcl_zero"))
            if lat is not None:
                ev = torch.cuda.Event()
                ev.record(lat)
                main.wait_event(ev)                   # BCL of this level may start; the next level's lattice runs on

            def splat():
                if gs:
                    # [el_minus_gr (4 ch, channel-major) ; previous features (point-major rows)] summed per vertex in
                    # registers through the lattice's vertex -> contributions lists, normalised, written once
                    ck(L.efgh_bcl_splat_gather(lv["prow"].data_ptr(), prev_ptr, prev_sn, prev_c, lv["voff"].data_ptr(),
                                               lv["contrib"].data_ptr(), h_cap, h_dev, 1 if self.use_norm else 0, S, cin,
                                               None, s), "efgh_bcl_splat_gather")
                    return
                if li == 0 and self.stem is not None:
                    # [el_minus_gr ; conv_in(xyz)]: the stem's three pointwise layers run on the splat's shared-memory tile
                    d = self.stem["dims"]
                    ck(L.efgh_bcl_scatter_stem(lv["elmgr"].data_ptr(), n_cap, 1, 4, pc.data_ptr(), pc.stride(0), d[0], d[1], d[2], d[3],
                                               self.stem["w"].data_ptr(), self.stem["slope"], n_cap, n_dev, lv["bary"].data_ptr(),
                                               n_cap, lv["loff32"].data_ptr(), 32, n_cap, 1, S, cin,
                                               lv["wsum"].data_ptr() if self.use_norm else None, s), "efgh_bcl_scatter_stem")
                else:
                  # [el_minus_gr (4 ch, channel-major) ; previous features] -> one scatter, no torch.cat
                  ck(L.efgh_bcl_scatter(lv["elmgr"].data_ptr(), n_cap, 1, 4, prev_ptr, prev_sc, prev_sn, prev_c, n_cap, n_dev,
                                      lv["bary"].data_ptr(), n_cap, lv["loff32"].data_ptr(), 32, n_cap, 1, S, cin,
                                      lv["wsum"].data_ptr() if self.use_norm else None, s), "efgh_bcl_scatter")
                if self.use_norm:
                    ck(L.efgh_bcl_normalize(S, cin, cin, lv["wsum"].data_ptr(), None, h_cap + 1, h_dev, 1, s),
                       "efgh_bcl_normalize")
            timed("L%d.splat" % li, splat)
            if lv["tc"]:
                # conv1 on tensor cores: long contraction -> partial sums added in L2, bias + ReLU deferred to
                # conv2's loader (in_bias / in_act), so Y holds raw sums and is never re-written
                split = lv["split0"]

                def conv1():
                    ck(L.efgh_bcl_conv_tc(S, cin, cin, None, 0, lv["nbr32"].data_ptr(), 32, h_cap, lv["F"], h_cap, h_dev,
                                          lv["img0"].data_ptr(), lv["b0"].data_ptr(), lv["cmid"], _ACT["relu"], lv["Y"].data_ptr(),
                                          lv["cmid"], self.nsplit, 1 if split else 0, s), "efgh_bcl_conv_tc")
                timed("L%d.conv1" % li, conv1)
                timed("L%d.conv2" % li, lambda: ck(L.efgh_bcl_conv_tc(
                    lv["Y"].data_ptr(), lv["cmid"], lv["cmid"], lv["b0"].data_ptr() if split else None, _ACT["relu"], None, 32, 0,
                    1, h_cap, h_dev, lv["img1"].data_ptr(), lv["b1"].data_ptr(), lv["cout"], self.final_act, lv["Z"].data_ptr(),
                    lv["cout"], self.nsplit, 0, s), "efgh_bcl_conv_tc"))
            else:
                timed("L%d.conv1" % li, lambda: ck(L.efgh_bcl_conv(
                    S, cin, cin, None, lv["nbr32"].data_ptr(), 32, h_cap, lv["F"],
                    h_cap, h_dev, lv["Wt0"].data_ptr(), lv["b0"].data_ptr(), lv["cmid"], _ACT["relu"], lv["Y"].data_ptr(),
                    lv["cmid"], 0, s), "efgh_bcl_conv"))
                timed("L%d.conv2" % li, lambda: ck(L.efgh_bcl_conv(
                    lv["Y"].data_ptr(), lv["cmid"], lv["cmid"], None, None, 32, 0, 1, h_cap, h_dev, lv["Wt1"].data_ptr(),
                    lv["b1"].data_ptr(), lv["cout"], self.final_act, lv["Z"].data_ptr(), lv["cout"], 0, s), "efgh_bcl_conv"))
            if lv["next"] is not None:
                pts_ptr, pts_ld = lv["next"].data_ptr(), h_cap
            prev_ptr, prev_sc, prev_sn, prev_c = lv["Z"].data_ptr(), 1, lv["cout"], lv["cout"]
            n_dev = h_dev
        return self.levels[-1]["Z"]

    @_on_own_device
    def graph_for(self, pc, feat0, stream):
        """CUDA graph of enqueue(pc, feat0) (captured once per input-buffer pair, then replayed): the ~55 launches
        of a scan become one graph launch, which removes the per-launch host cost and most inter-kernel gaps."""
        key = (pc.data_ptr(), feat0.data_ptr() if feat0 is not None else 0)
        g = self._graphs.get(key)
        if g is None:
            self.enqueue(pc, feat0, stream=stream)          # warm-up outside capture (function attributes, lazy init)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                self.enqueue(pc, feat0, stream=stream)
            self._graphs[key] = g
        return g

    @_on_own_device
    def forward_host(self, pc_host, feat_host, out_host, state_host, stream=None, use_graph=False, starts_host=None,
                     compute_stream=None):
        """End-to-end call on HOST buffers (pinned for async copies): H2D of the cloud and stem features,
        the whole scan, D2H of the level records and of the first out_host.shape[0] rows of the last
        level's output.  Everything is enqueued on `stream`; synchronise it before reading the outputs.
        Batched pipelines take lists of B per-scan host tensors ((3,n) / (C,n) each) - or one (3, B*n) / (C, B*n)
        tensor already in batch layout - and, if given, fill starts_host (nlev, B+1) int32 with every level's
        per-scan vertex boundaries.
        compute_stream: run the kernels there while the copies stay on `stream` - two pipelines with their own copy
        streams and ONE shared compute stream overlap PCIe transfers with kernels without letting the kernels of
        different batches compete for the SMs and the L2."""
        st = stream if stream is not None else torch.cuda.current_stream(self.dev)
        cs = compute_stream if compute_stream is not None else st
        with torch.cuda.stream(st):
            if isinstance(pc_host, (list, tuple)):
                n = self.n_scan
                for b in range(self.B):                     # one strided DMA per matrix (no staging buffer)
                    for dst, src in ((self._pc_dev, pc_host[b]),) + (((self._feat_dev, feat_host[b]),) if self.stem is None else ()):
                        assert src.is_contiguous() and src.dtype == torch.float32 and src.shape[1] <= n
                        _capi.check(self.L.efgh_copy_matrix_async(dst.data_ptr() + 4 * self._starts0[b], dst.stride(0),
                                                                  src.data_ptr(), src.shape[1], src.shape[0], src.shape[1],
                                                                  1, st.cuda_stream), "efgh_copy_matrix_async")
            else:
                self._pc_dev.copy_(pc_host, non_blocking=True)
                if self.stem is None:
                    self._feat_dev.copy_(feat_host, non_blocking=True)
        if cs is not st:
            ev = torch.cuda.Event()
            ev.record(st)
            cs.wait_event(ev)
        fdev = self._feat_dev if self.stem is None else None      # fused stem: the cloud is the only input
        with torch.cuda.stream(cs):
            if use_graph:
                self.graph_for(self._pc_dev, fdev, cs).replay()
                Z = self.levels[-1]["Z"]
            else:
                Z = self.enqueue(self._pc_dev, fdev, stream=cs)
        if cs is not st:
            ev = torch.cuda.Event()
            ev.record(cs)
            st.wait_event(ev)
        with torch.cuda.stream(st):
            out_host.copy_(Z[:out_host.shape[0]], non_blocking=True)
            state_host.copy_(self.states, non_blocking=True)
            if starts_host is not None:
                for li, lv in enumerate(self.levels):
                    starts_host[li].copy_(lv["info"][:self.B + 1], non_blocking=True)
        return out_host, state_host

    def counts(self):
        """Synchronising read of the per-level records: returns [H_0..H_4]; raises on a status bit."""
        host = self.states.cpu()
        for li in range(self.nlev):
            if int(host[li, 2]) & 2:
                raise VertexCapExceeded("level %d: more vertices than vertex_cap_factor * N allows" % li)
            check_status(int(host[li, 2]), li)
        return [int(host[li, 1]) for li in range(self.nlev)]

    def vertex_starts(self):
        """Batched mode: per level, the (B+1) global vertex boundaries of the scans (synchronising read)."""
        return [lv["info"][:self.B + 1].cpu().tolist() if self.batch_api else [0, c]
                for lv, c in zip(self.levels, self.counts())]

    def level_dicts(self, scan=None):
        """The reference-format per-level dicts of the last scan (views of the pipeline's buffers).  Batched
        pipelines: pass `scan=b` to get scan b's dicts with LOCAL vertex indices, i.e. what a single-scan build
        of that cloud returns (copies, not views)."""
        cnt = self.counts()
        if self.B > 1:
            assert scan is not None, "batched pipeline: level_dicts(scan=b)"
            vs = self.vertex_starts()
            out, p0, p1 = [], self._starts0[scan], self._starts0[scan + 1]
            for li, lv in enumerate(self.levels):
                v0, v1 = vs[li][scan], vs[li][scan + 1]
                lo = (lv["loff64"] if self.emit_int64 else lv["loff32"])[:, p0:p1] - v0
                nb = (lv["nbr64"] if self.emit_int64 else lv["nbr32"])[:, v0:v1]
                nb = torch.where(nb >= 0, nb - v0, nb)
                out.append({"pc1_barycentric": lv["bary"][None, :, p0:p1], "pc1_el_minus_gr": lv["elmgr"][None, :, p0:p1],
                            "pc1_lattice_offset": lo[None], "pc1_blur_neighbors": nb[None], "pc1_hash_cnt": v1 - v0})
                p0, p1 = v0, v1
            return out
        out, n = [], self.n0
        for li, lv in enumerate(self.levels):
            H = cnt[li]
            out.append({"pc1_barycentric": lv["bary"][None, :, :n], "pc1_el_minus_gr": lv["elmgr"][None, :, :n],
                        "pc1_lattice_offset": (lv["loff64"] if self.emit_int64 else lv["loff32"])[None, :, :n],
                        "pc1_blur_neighbors": (lv["nbr64"] if self.emit_int64 else lv["nbr32"])[None, :, :H],
                        "pc1_hash_cnt": H})
            n = H
        return out

    def outputs(self, scan=None):
        """[(1, C_out, H_l) views] - what bcn1..bcn5 return in reference enet.py:113-141 (batched: of scan `scan`)."""
        cnt = self.counts()
        if self.B > 1:
            assert scan is not None, "batched pipeline: outputs(scan=b)"
            vs = self.vertex_starts()
            return [lv["Z"][vs[li][scan]:vs[li][scan + 1]].t()[None] for li, lv in enumerate(self.levels)]
        return [lv["Z"][:cnt[li]].t()[None] for li, lv in enumerate(self.levels)]

    # ------------------------------------------------------------------------------------------
    def algorithmic_bytes(self, counts):
        """SURVEY.md §8(d) per-level byte model, summed: lattice 76 N + 132 H;
        BCL fwd 4 C_in N + 48 N + 8 C_in (H+1) + 120 H + 4 C_out H."""
        total, per_level, n = 0, [], self.n0
        for li, lv in enumerate(self.levels):
            H = counts[li]
            lat = 76 * n + 132 * H
            bcl = 4 * lv["cin"] * n + 48 * n + 8 * lv["cin"] * (H + 1) + 120 * H + 4 * lv["cout"] * H
            per_level.append((lat, bcl))
            total += lat + bcl
            n = H
        return total, per_level

    def conv_flops(self, counts):
        """2 * H * (F*C_in*C_mid + C_mid*C_out) per level."""
        return [2 * counts[li] * (lv["F"] * lv["cin"] * lv["cmid"] + lv["cmid"] * lv["cout"])
                for li, lv in enumerate(self.levels)]


def make_enet_weights(bcl_plan, filter_size=15, std=0.1, seed=0):
    """Random-init BCL weights with the reference's shapes (nets/bilateralNN.py:103-135)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for cin, (cmid, cout) in bcl_plan:
        out.append([(torch.randn(cmid, cin, filter_size, 1, generator=g) * std, torch.randn(cmid, generator=g) * std),
                    (torch.randn(cout, cmid, 1, 1, generator=g) * std, torch.randn(cout, generator=g) * std)])
    return out
