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
import os

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
                 precision="3xtf32", batch=1, gather_splat=True, stem=None, train=False, level0_gather=None):
        """bcl_plan: [(C_in, [C_mid, C_out]), ...] one entry per level (reference nets/enet.py:30-83);
        weights: per level [(W0 (C_mid,C_in,F,1), b0), (W1 (C_out,C_mid,1,1), b1)] torch tensors;
        vertex_cap_factor: capacity of every vertex-side buffer as a multiple of n_points;
        precision: "3xtf32" (tcgen05, fp32-equivalent), "tf32" (tcgen05, one pass) or "fp32" (CUDA cores);
        gather_splat: levels >= 1 (whose input features are the previous level's point-major output rows) splat
        through the vertex -> contributions lists of the lattice build - no atomics, zero-fill and normalisation
        fused;
        level0_gather: level 0 the same way, from point-major (N, C_stem) feature rows that efgh_bcl_stem_rows computes
        from the cloud (stem given) or that are transposed from feat0.  Off by default = vector-atomic scatter of the
        channel-major (C, N) input (stem evaluated on the scatter's tile) + normalisation pass: measured on 16-scan
        batches (r2) the gather form's splat is faster (667 vs ~740 us) but materialising the 268 MB of feature rows
        and building level 0's contribution lists costs more than that saves (1 190 vs 930 us in total);
        stem: None, or ([(W1, b1), (W2, b2), (W3, b3)], use_leaky) - E-Net's pointwise `conv_in` (reference
        nets/enet.py:24-28; W as Conv1d weights (out, in, 1)): the level-0 splat then COMPUTES the stem features from
        the cloud and enqueue() ignores feat0;
        batch: scans per launch sequence, each of n_points points (inputs are then (3, batch*n_points) /
        (C, batch*n_points), scan b in columns [b*n_points, (b+1)*n_points));
        train: keep what backward() needs (normalisation factors, activated conv1 outputs) and allocate the gradient
        buffers; weights can then be replaced between steps with load_weights()."""
        self.dev = torch.device(device)
        self.L = _capi.lib()
        self.B = int(batch)
        assert 1 <= self.B <= 64
        self.train = bool(train)
        self.wgrad_tc = os.environ.get("EFGH_WGRAD", "tc") != "ffma"     # weight gradients on the tensor cores (else the fp32 CUDA-core kernel)
        assert not (train and stem is not None), "training: the stem runs in torch (autograd); pass its output as feat0"
        self.gather_splat = bool(gather_splat)
        self.stem = None
        if stem is not None:
            layers, leaky = stem
            assert len(layers) == 3
            dims = [int(layers[0][0].shape[1])] + [int(W.shape[0]) for W, _ in layers]
            assert dims[3] == stem_channels and all(int(W.shape[1]) == dims[i] for i, (W, _) in enumerate(layers))
            flat = torch.cat([t.detach().to(torch.float32).reshape(-1) for W, b in layers for t in (W, b)])
            assert flat.numel() == self.L.efgh_bcl_stem_weight_floats(*dims)
            self.stem = {"dims": dims, "w": flat.to(self.dev).contiguous(), "slope": 0.1 if leaky else 0.0}
        self.gs0 = self.gather_splat and bool(level0_gather)
        self.level0_splat = ("gather through vertex -> contributions lists from point-major stem rows" if self.gs0
                             else "vector-atomic scatter + normalise")
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
                    "table": int(self.L.efgh_lattice_table_entries(n_cap_scan, int(float(os.environ.get("EFGH_TABLE_FACTOR", "2")) * cap_scan))),
                    "info": torch.zeros(max(int(self.L.efgh_lattice_batch_info_ints(self.B)), 1), dtype=i32, device=dev),
                    "gs": self.gather_splat and (li > 0 or self.gs0),
                    "voff": torch.zeros(int(self.L.efgh_lattice_vertex_offsets_ints(h_cap)), dtype=i32, device=dev) if self.gather_splat and (li > 0 or self.gs0) else None,
                    "prow": torch.zeros((n_cap, 8), dtype=f32, device=dev) if self.gather_splat and (li > 0 or self.gs0) else None,
                    "contrib": torch.zeros(4 * n_cap, dtype=i32, device=dev) if self.gather_splat and (li > 0 or self.gs0) else None,
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
                lv["tc"] = bool(self.nsplit and self.L.efgh_bcl_conv_tc_supported(cin, F, cmid, self.nsplit)
                                and self.L.efgh_bcl_conv_tc_supported(cmid, 1, cout, self.nsplit)
                                and self.L.efgh_bcl_conv_tc_groups(cmid, cout) == 1)
                if lv["tc"]:
                    for nm, K, M in (("img0", F * cin, cmid), ("img1", cmid, cout)):
                        lv[nm] = torch.empty(self.L.efgh_bcl_packed_weight_bytes(K, M, self.nsplit) // 4, dtype=f32, device=dev)
                    lv["split0"] = self.L.efgh_bcl_conv_tc_groups(F * cin, cmid) > 1
                if self.train:
                    assert lv["tc"] and self.gather_splat, "training needs the tensor-core convolution and the gather-form splat"
                    cg = cin - 4                                       # channels that carry a gradient (all but el_minus_gr)
                    assert self.L.efgh_bcl_conv_tc_supported(cmid, F, cg, self.nsplit) and self.L.efgh_bcl_conv_tc_supported(cout, 1, cmid, self.nsplit)
                    lv["cg"] = cg
                    lv["inv"] = torch.zeros((h_cap + 1,), dtype=f32, device=dev)
                    lv["dA"] = torch.zeros((h_cap + 1, cmid), dtype=f32, device=dev)       # row 0: the sink row of the gather-form dgrad
                    lv["dS"] = torch.zeros((h_cap + 1, cg), dtype=f32, device=dev)
                    lv["dZ"] = torch.zeros((h_cap, cout), dtype=f32, device=dev)
                    lv["imgD0"] = torch.empty(self.L.efgh_bcl_packed_weight_bytes(F * cmid, cg, self.nsplit) // 4, dtype=f32, device=dev)
                    lv["imgD1"] = torch.empty(self.L.efgh_bcl_packed_weight_bytes(cout, cmid, self.nsplit) // 4, dtype=f32, device=dev)
                    lv["splitD0"] = self.L.efgh_bcl_conv_tc_groups(F * cmid, cg) > 1
                    offs = [tuple(o) for o in self.gd.radius2offset[radius].tolist()]
                    lv["mirror"] = [offs.index(tuple(-v for v in o)) for o in offs]     # tap f <-> the tap with the negated offset
                    lv["mirror_t"] = torch.tensor(lv["mirror"], dtype=torch.int64, device=dev)   # (device copy: load_weights may run under graph capture)
                if self.batch_api:
                    ws_bytes = max(ws_bytes, self.L.efgh_lattice_batch_workspace_bytes(self.B, lv["table"], n_cap))
                else:
                    ws_bytes = max(ws_bytes, self.L.efgh_lattice_workspace_bytes(n_cap))
                self.levels.append(lv)
                n_cap = h_cap
                n_cap_scan = min(4 * n_cap_scan, cap_scan)
                prev_c = cout
            self.ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            if self.train:        # weight gradients: one flat buffer (one memset per step), views per level
                sizes = []
                for lv in self.levels:
                    sizes += [lv["F"] * lv["cin"] * lv["cmid"], lv["cmid"], lv["cmid"] * lv["cout"], lv["cout"]]
                self._gflat = torch.zeros(sum(sizes), dtype=f32, device=dev)
                o = 0
                for lv in self.levels:
                    for nm, shape in (("gWt0", (lv["F"] * lv["cin"], lv["cmid"])), ("gb0", (lv["cmid"],)),
                                      ("gWt1", (lv["cmid"], lv["cout"])), ("gb1", (lv["cout"],))):
                        n = int(np.prod(shape))
                        lv[nm] = self._gflat[o:o + n].view(shape)
                        o += n
                self._loss = torch.zeros((), dtype=f32, device=dev)
                self._dfeat0 = torch.empty((stem_channels, self.n0), dtype=f32, device=dev)
                self._side = None
            self.load_weights(weights)
            self._starts0 = [b * self.n_scan for b in range(self.B + 1)]
            self.scan_start = torch.tensor(self._starts0, dtype=i32, device=dev)
            self._pc_dev = torch.empty((3, self.n0), dtype=f32, device=dev)
            self._feat_dev = torch.empty((stem_channels, self.n0), dtype=f32, device=dev)
            self._feat_rows = torch.empty((self.n0, stem_channels), dtype=f32, device=dev) if self.gs0 else None   # point-major level-0 features
        # per launch sequence: clear/points/assign, vertices, zero, splat (+ normalise | level-0 transpose), conv1, conv2
        self.launches_per_scan = self.nlev * (3 + 1 + 1 + 1 + 2) + sum(0 if lv["gs"] else 1 for lv in self.levels)

    @_on_own_device
    def load_weights(self, weights):
        """(Re-)lay the convolution weights for the kernels: (K, M) matrices, packed tensor-core images and - when
        training - the images of the two data-gradient convolutions.  Stream-ordered on the current stream, no host
        sync; previously captured graphs stay valid (same buffers)."""
        L, ck = self.L, _capi.check
        s = torch.cuda.current_stream(self.dev).cuda_stream
        f32 = torch.float32
        for lv, ((W0, b0), (W1, b1)) in zip(self.levels, weights):
            F, cin, cmid, cout = lv["F"], lv["cin"], lv["cmid"], lv["cout"]
            assert tuple(W0.shape) == (cmid, cin, F, 1) and tuple(W1.shape) == (cout, cmid, 1, 1)
            W0 = W0.detach().to(self.dev, f32)
            W1 = W1.detach().to(self.dev, f32)
            lv["Wt0"] = W0[:, :, :, 0].permute(2, 1, 0).reshape(F * cin, cmid).contiguous()
            lv["b0"] = b0.detach().to(self.dev, f32).contiguous()
            lv["Wt1"] = W1[:, :, 0, 0].t().contiguous()
            lv["b1"] = b1.detach().to(self.dev, f32).contiguous()
            if lv["tc"]:
                for nm, K, M in (("0", F * cin, cmid), ("1", cmid, cout)):
                    ck(L.efgh_bcl_pack_weights(lv["Wt" + nm].data_ptr(), K, M, self.nsplit, lv["img" + nm].data_ptr(), s), "efgh_bcl_pack_weights")
            if self.train:
                # conv1's data gradient as a gather over the same neighbour table: dS[g, c] = sum_t sum_m dY[nbr[t,g], m] W0[m, c, mirror(t)]
                cg = lv["cg"]
                Wg = W0[:, 4:, :, 0].index_select(2, lv["mirror_t"]).permute(2, 0, 1).reshape(F * cmid, cg).contiguous()
                ck(L.efgh_bcl_pack_weights(Wg.data_ptr(), F * cmid, cg, self.nsplit, lv["imgD0"].data_ptr(), s), "efgh_bcl_pack_weights")
                Wd = W1[:, :, 0, 0].contiguous()                        # (K = cout, N = cmid)
                ck(L.efgh_bcl_pack_weights(Wd.data_ptr(), cout, cmid, self.nsplit, lv["imgD1"].data_ptr(), s), "efgh_bcl_pack_weights")

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
        if self.gs0:
            # level-0 features as point-major rows (on the BCL stream: overlaps the lattice build of level 0)
            def rows0():
                if self.stem is not None:
                    d = self.stem["dims"]
                    ck(L.efgh_bcl_stem_rows(pc.data_ptr(), pc.stride(0), d[0], d[1], d[2], d[3], self.stem["w"].data_ptr(), self.stem["slope"],
                                            self.n0, None, self._feat_rows.data_ptr(), self._feat_rows.stride(0), s), "efgh_bcl_stem_rows")
                else:
                    with torch.cuda.stream(main):
                        self._feat_rows.copy_(feat0.t())
            timed("L0.stem", rows0)
            prev_ptr, prev_sc, prev_sn, prev_c = self._feat_rows.data_ptr(), 1, self._feat_rows.stride(0), self._feat_rows.shape[1]
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
                                               lv["inv"].data_ptr() if self.train else None, s), "efgh_bcl_splat_gather")
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
                    ck(L.efgh_bcl_normalize(S, cin, cin, lv["wsum"].data_ptr(), lv["inv"].data_ptr() if self.train else None,
                                            h_cap + 1, h_dev, 1, s), "efgh_bcl_normalize")
            timed("L%d.splat" % li, splat)
            if lv["tc"]:
                # conv1 on tensor cores: long contraction -> partial sums added in L2, bias + ReLU deferred to
                # conv2's loader (in_bias / in_act), so Y holds raw sums and is never re-written
                split = lv["split0"]

                def conv1():
                    ck(L.efgh_bcl_conv_tc(S, cin, cin, None, 0, lv["nbr32"].data_ptr(), 32, h_cap, lv["F"], h_cap, h_dev,
                                          lv["img0"].data_ptr(), lv["b0"].data_ptr(), lv["cmid"], _ACT["relu"], lv["Y"].data_ptr(),
                                          lv["cmid"], self.nsplit, 1 if split else 0, s), "efgh_bcl_conv_tc")
                    if split and self.train:    # backward needs the ACTIVATED conv1 output: apply the deferred bias + ReLU in place
                        ck(L.efgh_bcl_bias_act(lv["Y"].data_ptr(), lv["cmid"], lv["cmid"], h_cap, h_dev, lv["b0"].data_ptr(), _ACT["relu"], s),
                           "efgh_bcl_bias_act")
                timed("L%d.conv1" % li, conv1)
                timed("L%d.conv2" % li, lambda: ck(L.efgh_bcl_conv_tc(
                    lv["Y"].data_ptr(), lv["cmid"], lv["cmid"], lv["b0"].data_ptr() if split and not self.train else None, _ACT["relu"], None, 32, 0,
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

    # ------------------------------------------------------------------------------------------
    # training (SURVEY.md §8 row a17 for a whole batch; reference: autograd over nets/bilateralNN.py:148-263)
    @_on_own_device
    def loss_half_mean_square(self, stream=None):
        """loss = mean over the B scans of 0.5 * mean(Z_b^2) of the last level's output; fills the last level's dZ.
        Returns (loss 0-d device tensor, dZ buffer).  Stream-ordered, no host sync."""
        assert self.train
        lv = self.levels[-1]
        s = (stream if stream is not None else torch.cuda.current_stream(self.dev)).cuda_stream
        seg = lv["info"].data_ptr() if self.batch_api else None
        _capi.check(self.L.efgh_bcl_loss_half_mean_square(lv["Z"].data_ptr(), lv["cout"], lv["cout"], seg, self.B, lv["dZ"].data_ptr(),
                                                          lv["cout"], self._loss.data_ptr(), lv["h_cap"], s), "efgh_bcl_loss_half_mean_square")
        return self._loss, lv["dZ"]

    @_on_own_device
    def backward(self, dZ=None, stream=None):
        """Backward of the last enqueue() through all levels: the last level's dZ buffer (filled by the caller or by
        loss_half_mean_square; vertex-major (h_cap, C_out)) -> weight gradients (weight_grads()) and the gradient of
        feat0, returned as a (C_stem, N) tensor (a buffer of the pipeline).

        Per level, last to first (all counts stay on the device):
          conv2: dA = dZ W1 on the tensor cores, written below the sink row of dA; dW1, db1 (efgh_bcl_conv_wgrad)
          ReLU:  dA *= (Y > 0) in place
          conv1: dW0, db0 from the gathered splat matrix; dS = gather-form convolution of dA with mirrored taps over the
                 channels that carry a gradient (all but the 4 el_minus_gr channels) - valid because the neighbour
                 table is mirror-symmetric unless EFGH_ST_ALIASED is set (checked by counts())
          splat: d(previous output)[n, :] = sum_r bary[r, n] * inv[row] * dS[row, :] (efgh_bcl_gather) -> the previous level's dZ
        The weight-gradient kernels only feed the optimizer, so they run on a side stream next to the data-gradient chain."""
        assert self.train
        L, ck = self.L, _capi.check
        main = stream if stream is not None else torch.cuda.current_stream(self.dev)
        s = main.cuda_stream
        if self._side is None:
            self._side = torch.cuda.Stream(self.dev)
        side = self._side
        ss = side.cuda_stream
        with torch.cuda.stream(main):
            if dZ is not None and dZ.data_ptr() != self.levels[-1]["dZ"].data_ptr():
                self.levels[-1]["dZ"][:dZ.shape[0]].copy_(dZ)
            self._gflat.zero_()
        inv_of = (lambda lv: lv["inv"].data_ptr()) if self.use_norm else (lambda lv: None)
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        n_dev_top = (self.scan_start.data_ptr() + 4 * self.B) if self.batch_api else None
        for li in range(self.nlev - 1, -1, -1):
            lv = self.levels[li]
            st = self.states[li].data_ptr()
            h_dev, h_cap = st + 4, lv["h_cap"]
            cin, cmid, cout, cg, F = lv["cin"], lv["cmid"], lv["cout"], lv["cg"], lv["F"]
            dZl, dA, dS = lv["dZ"].data_ptr(), lv["dA"].data_ptr(), lv["dS"].data_ptr()
            dA1 = dA + 4 * cmid                                  # row 1: vertex 0
            # --- conv2
            if self.final_act:
                ck(L.efgh_bcl_act_bwd(dZl, cout, lv["Z"].data_ptr(), cout, cout, self.final_act, h_cap, h_dev, s), "efgh_bcl_act_bwd")
            ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)          # dZ of this level is complete
            if self.wgrad_tc:
                ck(L.efgh_bcl_conv_wgrad_tc(lv["Y"].data_ptr(), cmid, cmid, None, 32, 0, 1, h_cap, h_dev, dZl, cout, cout,
                                            lv["gWt1"].data_ptr(), lv["gb1"].data_ptr(), ss), "efgh_bcl_conv_wgrad_tc")
            else:
                ck(L.efgh_bcl_conv_wgrad(lv["Y"].data_ptr(), cmid, cmid, None, None, 32, 0, 1, h_cap, h_dev, dZl, cout, None, 0, 0, cout,
                                         lv["gWt1"].data_ptr(), lv["gb1"].data_ptr(), ss), "efgh_bcl_conv_wgrad")
            ck(L.efgh_bcl_conv_tc(dZl, cout, cout, None, 0, None, 32, 0, 1, h_cap, h_dev, lv["imgD1"].data_ptr(), None, cmid, 0,
                                  dA1, cmid, self.nsplit, 0, s), "efgh_bcl_conv_tc(dgrad 1x1)")
            ck(L.efgh_bcl_act_bwd(dA1, cmid, lv["Y"].data_ptr(), cmid, cmid, _ACT["relu"], h_cap, h_dev, s), "efgh_bcl_act_bwd")
            # --- conv1
            ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)          # masked dA is complete
            if self.wgrad_tc:
                ck(L.efgh_bcl_conv_wgrad_tc(lv["S"].data_ptr(), cin, cin, lv["nbr32"].data_ptr(), 32, h_cap, F, h_cap, h_dev, dA1, cmid, cmid,
                                            lv["gWt0"].data_ptr(), lv["gb0"].data_ptr(), ss), "efgh_bcl_conv_wgrad_tc")
            else:
                ck(L.efgh_bcl_conv_wgrad(lv["S"].data_ptr(), cin, cin, None, lv["nbr32"].data_ptr(), 32, h_cap, F, h_cap, h_dev, dA1, cmid,
                                         None, 0, 0, cmid, lv["gWt0"].data_ptr(), lv["gb0"].data_ptr(), ss), "efgh_bcl_conv_wgrad")
            split = lv["splitD0"]
            if split:
                ck(L.efgh_bcl_zero(None, cg, cg, None, dS + 4 * cg, cg, cg, h_cap, h_dev, 0, s), "efgh_bcl_zero")
            ck(L.efgh_bcl_conv_tc(dA, cmid, cmid, None, 0, lv["nbr32"].data_ptr(), 32, h_cap, F, h_cap, h_dev, lv["imgD0"].data_ptr(),
                                  None, cg, 0, dS + 4 * cg, cg, self.nsplit, 1 if split else 0, s), "efgh_bcl_conv_tc(dgrad)")
            # --- splat adjoint -> gradient of the previous level's output (or of feat0)
            n_cap = lv["n_cap"]
            if li > 0:
                prev = self.levels[li - 1]
                n_dev = self.states[li - 1].data_ptr() + 4
                ck(L.efgh_bcl_gather(dS, cg, cg, inv_of(lv), n_cap, n_dev, lv["bary"].data_ptr(), n_cap, lv["loff32"].data_ptr(), 32,
                                     n_cap, 1, None, prev["dZ"].data_ptr(), 1, cg, s), "efgh_bcl_gather")
            else:
                ck(L.efgh_bcl_gather(dS, cg, cg, inv_of(lv), n_cap, n_dev_top, lv["bary"].data_ptr(), n_cap, lv["loff32"].data_ptr(), 32,
                                     n_cap, 1, None, self._dfeat0.data_ptr(), self._dfeat0.stride(0), 1, s), "efgh_bcl_gather")
        ev = torch.cuda.Event()
        ev.record(side)
        main.wait_event(ev)
        return self._dfeat0

    def aliased_levels(self):
        """Levels whose neighbour table lost its mirror symmetry (EFGH_ST_ALIASED: a neighbour key left the key box and the
        reference's key2int aliasing found another vertex) - backward()'s gather-form data gradient is not valid there.
        Synchronising read."""
        host = self.states.cpu()
        return [li for li in range(self.nlev) if int(host[li, 2]) & 8]

    def weight_grads(self):
        """[[(dW0 (C_mid, C_in, F, 1), db0), (dW1 (C_out, C_mid, 1, 1), db1)] per level] in the reference's parameter layout
        (views of the pipeline's flat gradient buffer: consume them before the next backward())."""
        out = []
        for lv in self.levels:
            g0 = lv["gWt0"].view(lv["F"], lv["cin"], lv["cmid"]).permute(2, 1, 0).unsqueeze(-1)
            g1 = lv["gWt1"].t()[:, :, None, None]
            out.append([(g0, lv["gb0"]), (g1, lv["gb1"])])
        return out

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
