"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

from efgh_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from efgh_b200 import _capi
    _capi.lib()  # fail loudly if the extension is missing
    return torch.device("cuda:0")


def _to_np(levels):
    out = []
    for d in levels:
        out.append({k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in d.items()})
    return out


def _run(pc, smap, dev, exact):
    from efgh_b200.generate_data import GenerateData
    gd = GenerateData(3, smap, "cuda", exact=exact)
    pc1, data = gd(torch.from_numpy(pc).to(dev))
    assert pc1.is_cuda and pc1.dtype == torch.float32
    return _to_np(data)


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("name", H.lattice_golden_names())
def test_lattice_small_golden(name, exact, dev):
    pc, smap, _, levels = H.load_lattice_golden(name)
    got = _run(pc, smap, dev, exact)
    assert len(got) == len(levels)
    for li, (g, w) in enumerate(zip(got, levels)):
        H.assert_level_equal(g, w, "%s L%d exact=%s" % (name, li, exact))


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("case", H.digest_cases())
def test_lattice_fullsize_digests(case, exact, dev):
    """BASELINE.json configs 1, 2 and 5 clouds: bit-exact against digests of the live reference's output."""
    dig = H.load_digests()
    sensor, seed = case.split("/")
    pc = synth.synth_scan(int(seed[4:]), sensor)
    assert H.digest(pc) == dig[case + "/pc"]
    got = _run(pc, synth.SCALE_MAP, dev, exact)
    assert ",".join(str(g["pc1_hash_cnt"]) for g in got) == dig[case + "/cnt"]
    for li, g in enumerate(got):
        for k in ("pc1_barycentric", "pc1_el_minus_gr", "pc1_lattice_offset", "pc1_blur_neighbors"):
            assert H.digest(g[k]) == dig["%s/L%d/%s" % (case, li, k)], (case, li, k)


@pytest.mark.parametrize("n,spread,seed", [(1, 3.0, 0), (5, 0.01, 1), (33, 100.0, 2), (1000, 5.0, 3), (50000, 30.0, 4)])
def test_lattice_random_vs_oracle(n, spread, seed, dev):
    from oracle import lattice as ol
    rng = np.random.default_rng(seed)
    pc = (rng.standard_normal((3, n)) * spread).astype(np.float32)
    smap = [[1.0, 1], [0.5, 2], [0.25, 1]] if n <= 1000 else synth.SCALE_MAP
    want = ol.generate(pc, smap)
    for exact in (True, False):
        got = _run(pc, smap, dev, exact)
        for li, (g, w) in enumerate(zip(got, want)):
            H.assert_level_equal(g, w, "n=%d L%d exact=%s" % (n, li, exact))


def test_lattice_properties_fullsize(dev):
    """Size-independent properties at the full 131k size: offsets are first-occurrence ranks, barycentric
    weights sum to 1, neighbour 0 is the vertex itself and the neighbour relation is mirror-symmetric."""
    pc = synth.synth_scan(5, "os1-64")
    got = _run(pc, synth.SCALE_MAP, dev, False)
    for li, g in enumerate(got):
        off = g["pc1_lattice_offset"][0]
        Hn = g["pc1_hash_cnt"]
        stream = off.T.reshape(-1)                      # point-major, remainder-minor
        uniq, first = np.unique(stream, return_index=True)
        assert len(uniq) == Hn and uniq[0] == 0 and uniq[-1] == Hn - 1
        assert np.all(np.diff(first[np.argsort(uniq)]) > 0)   # index order == first-occurrence order
        assert np.abs(g["pc1_barycentric"][0].sum(0) - 1).max() < 1e-5
        nb = g["pc1_blur_neighbors"][0]
        assert np.array_equal(nb[0], np.arange(Hn))
        for f in range(1, 15):
            src = np.nonzero(nb[f] >= 0)[0]
            assert np.array_equal(nb[15 - f][nb[f][src]], src)


def test_lattice_empty_cloud(dev):
    got = _run(np.zeros((3, 0), np.float32), synth.SCALE_MAP[:2], dev, True)
    assert [g["pc1_hash_cnt"] for g in got] == [0, 0]
    assert got[0]["pc1_blur_neighbors"].shape == (1, 15, 0)


def _bcl_module_from_golden(z, dev):
    from efgh_b200.bilateralNN import BilateralConvFlex
    cfg = z["cfg"].tolist()
    num_input, do_splat, do_slice, use_norm, last_relu, use_leaky = cfg[:6]
    m = BilateralConvFlex(3, 1, num_input, cfg[6:], "cuda", use_bias=True, use_leaky=bool(use_leaky),
                          use_norm=bool(use_norm), do_splat=bool(do_splat), do_slice=bool(do_slice),
                          last_relu=bool(last_relu), chunk_size=-1)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p_")}
    sd["feat_indices"] = torch.arange(num_input)
    if do_slice:
        sd["out_indices"] = torch.arange(cfg[-1])
    m.load_state_dict(sd, strict=True)
    return m.to(dev), bool(do_slice)


@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("name", H.bcl_golden_names())
def test_bcl_forward_backward_vs_reference_golden(name, idx_dtype, dev, golden_dir):
    """Tolerance: 1e-5 relative to the tensor's max magnitude (north_star: fp32 features within 1e-5)."""
    z = np.load("%s/bcl_%s.npz" % (golden_dir, name))
    m, do_slice = _bcl_module_from_golden(z, dev)
    feat = torch.from_numpy(z["feat"]).to(dev).requires_grad_(True)
    bary = torch.from_numpy(z["bary"]).to(dev)
    off = torch.from_numpy(z["off"].astype(np.int64)).to(dev).to(idx_dtype)
    nbr = torch.from_numpy(z["nbr"].astype(np.int64)).to(dev).to(idx_dtype)
    out = m(feat, bary, off, nbr, bary if do_slice else None, off if do_slice else None)
    assert tuple(out.shape) == z["out"].shape
    assert H.rel_err(out.detach().cpu().numpy(), z["out"]) < 1e-5
    out.backward(torch.from_numpy(z["gout"]).to(dev))
    assert H.rel_err(feat.grad.cpu().numpy(), z["gfeat"]) < 1e-5
    for k, p in m.named_parameters():
        assert H.rel_err(p.grad.cpu().numpy(), z["g_" + k]) < 2e-5, k


PER_LAYER_TOL = 1e-5      # north_star: features within 1e-5 relative on the SAME inputs (fp32-equivalent paths)
CHAINED_TOL = 5e-5        # five chained layers against a float64 chain: per-layer errors compound


@pytest.mark.parametrize("precision", ["3xtf32", "fp32"])
def test_bcl_enet_chain_vs_oracle(precision, dev, monkeypatch):
    """Config 1 of BASELINE.json: 16k cloud, lattice build + the five E-Net BCLs chained as
    reference nets/enet.py:113-141 does.  Each layer is checked against the float64 oracle fed with the SAME
    input the CUDA layer saw (PER_LAYER_TOL), and the whole chain against a pure float64 chain (CHAINED_TOL)."""
    from efgh_b200.generate_data import GenerateData
    from efgh_b200 import bilateralNN
    from efgh_b200.bilateralNN import BilateralConvFlex
    from oracle import bcl as obcl
    monkeypatch.setattr(bilateralNN, "CONV_PRECISION", precision)
    torch.manual_seed(0)
    pc = synth.synth_scan(2, "os1-64-16k")
    gd = GenerateData(3, synth.SCALE_MAP, "cuda", exact=False)
    _, data = gd(torch.from_numpy(pc).to(dev))
    prev = torch.randn(1, 32, pc.shape[1], device=dev)
    chain_ref = prev.cpu().double()
    for li, (cin, nout) in enumerate(synth.ENET_BCL):
        d = data[li]
        m = BilateralConvFlex(3, 1, cin, nout, "cuda", True, True, True, True, False, False, chunk_size=-1).to(dev)
        for p in m.parameters():
            torch.nn.init.normal_(p, 0, 0.1)
        x = torch.cat((d["pc1_el_minus_gr"], prev), dim=1)
        with torch.no_grad():
            y = m(x, d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
        convs = [(m.blur_conv[0].weight.detach().cpu(), m.blur_conv[0].bias.detach().cpu()),
                 (m.blur_conv[2].weight.detach().cpu(), m.blur_conv[2].bias.detach().cpu())]
        args = (d["pc1_barycentric"].cpu(), d["pc1_lattice_offset"].cpu(), d["pc1_blur_neighbors"].cpu(), convs)
        same_in = obcl.bcl_forward(x.cpu().double(), *args, dtype=torch.float64)
        chain_ref = obcl.bcl_forward(torch.cat((d["pc1_el_minus_gr"].cpu().double(), chain_ref), 1), *args, dtype=torch.float64)
        assert tuple(y.shape) == tuple(same_in.shape) == (1, nout[-1], d["pc1_hash_cnt"])
        e1 = H.rel_err(y.cpu().numpy(), same_in.numpy())
        e2 = H.rel_err(y.cpu().numpy(), chain_ref.numpy())
        print("%s level %d: same-input rel err %.2e, chained %.2e" % (precision, li, e1, e2))
        assert e1 < PER_LAYER_TOL, "level %d same-input rel err %g" % (li, e1)
        assert e2 < CHAINED_TOL, "level %d chained rel err %g" % (li, e2)
        prev = y


@pytest.mark.parametrize("gather_splat", [True, False])
@pytest.mark.parametrize("sensor,factor", [("os1-64-16k", 4.0), ("os1-64", 1.0)])
def test_scan_pipeline_vs_oracle(sensor, factor, gather_splat, dev):
    """The sync-free whole-scan pipeline (what bench.py times): lattice dicts bit-exact against the C oracle,
    BCL outputs of all five levels within 1e-5 of the float64 oracle (config 1 and config 2 clouds)."""
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    from oracle import lattice as ol, bcl as obcl
    pc = synth.synth_scan(4, sensor)
    N = pc.shape[1]
    weights = make_enet_weights(synth.ENET_BCL, seed=3)
    pipe = ScanPipeline(N, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=factor, gather_splat=gather_splat)
    feat0 = torch.randn(32, N, generator=torch.Generator().manual_seed(1))
    for _ in range(2):  # twice: buffers are reused across scans
        pipe.enqueue(torch.from_numpy(pc).to(dev), feat0.to(dev))
    want = ol.generate(pc, synth.SCALE_MAP)
    got = _to_np(pipe.level_dicts())
    for li, (g, w) in enumerate(zip(got, want)):
        H.assert_level_equal(g, w, "pipeline L%d" % li)
    outs = pipe.outputs()
    chain = feat0[None].double()
    gpu_prev = feat0[None].double()
    for li, w in enumerate(want):
        args = (torch.from_numpy(w["pc1_barycentric"]), torch.from_numpy(w["pc1_lattice_offset"]),
                torch.from_numpy(w["pc1_blur_neighbors"]), weights[li])
        elm = torch.from_numpy(w["pc1_el_minus_gr"]).double()
        same_in = obcl.bcl_forward(torch.cat((elm, gpu_prev), 1), *args, dtype=torch.float64)
        chain = obcl.bcl_forward(torch.cat((elm, chain), 1), *args, dtype=torch.float64)
        got_l = outs[li].cpu()
        assert tuple(got_l.shape) == tuple(chain.shape)
        e1, e2 = H.rel_err(got_l.numpy(), same_in.numpy()), H.rel_err(got_l.numpy(), chain.numpy())
        print("pipeline level %d: same-input rel err %.2e, chained %.2e" % (li, e1, e2))
        assert e1 < PER_LAYER_TOL, "pipeline BCL level %d same-input rel err %g" % (li, e1)
        assert e2 < CHAINED_TOL, "pipeline BCL level %d chained rel err %g" % (li, e2)
        gpu_prev = got_l.double()


@pytest.mark.parametrize("gather_splat", [True, False])
@pytest.mark.parametrize("sizes", [None, (16384, 5000, 1, 12001)])
def test_batched_pipeline_matches_single_scans(sizes, gather_splat, dev):
    """Ragged batch: B scans through one launch sequence.  Per scan the lattice must equal the C oracle
    bit for bit (own hash table / key box / insertion order, local vertex numbering), and the BCL outputs must
    match the float64 oracle within the per-layer tolerance - i.e. batching changes nothing but the launch count."""
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    from oracle import lattice as ol, bcl as obcl
    B = 4 if sizes else 3
    n = 16384
    sizes = list(sizes) if sizes else [n] * B
    clouds = [synth.synth_scan(30 + b, "os1-64-16k")[:, :sizes[b]] for b in range(B)]
    weights = make_enet_weights(synth.ENET_BCL, seed=3)
    pipe = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=4.0, batch=B, gather_splat=gather_splat)
    if sizes != [n] * B:
        pipe.set_scan_sizes(sizes)
    g = torch.Generator().manual_seed(2)
    feats = [torch.randn(32, sizes[b], generator=g) for b in range(B)]
    pc_all = torch.zeros(3, B * n)
    ft_all = torch.zeros(32, B * n)
    o = 0
    for b in range(B):
        pc_all[:, o:o + sizes[b]] = torch.from_numpy(clouds[b])
        ft_all[:, o:o + sizes[b]] = feats[b]
        o += sizes[b]
    for _ in range(2):                                  # twice: buffers are reused
        pipe.enqueue(pc_all.to(dev), ft_all.to(dev))
    tot = pipe.counts()
    vs = pipe.vertex_starts()
    for b in range(B):
        want = ol.generate(clouds[b], synth.SCALE_MAP)
        got = _to_np(pipe.level_dicts(scan=b))
        for li, (gl, wl) in enumerate(zip(got, want)):
            H.assert_level_equal(gl, wl, "batch scan %d L%d" % (b, li))
        outs = pipe.outputs(scan=b)
        prev = feats[b][None].double()
        for li, w in enumerate(want):
            args = (torch.from_numpy(w["pc1_barycentric"]), torch.from_numpy(w["pc1_lattice_offset"]),
                    torch.from_numpy(w["pc1_blur_neighbors"]), weights[li])
            ref = obcl.bcl_forward(torch.cat((torch.from_numpy(w["pc1_el_minus_gr"]).double(), prev), 1), *args, dtype=torch.float64)
            got_l = outs[li].cpu()
            assert tuple(got_l.shape) == tuple(ref.shape)
            e = H.rel_err(got_l.numpy(), ref.numpy())
            assert e < PER_LAYER_TOL, "batch scan %d level %d rel err %g" % (b, li, e)
            prev = got_l.double()
    for li in range(len(tot)):
        assert vs[li][0] == 0 and vs[li][-1] == tot[li] and all(a < c for a, c in zip(vs[li], vs[li][1:]))


def test_batched_forward_host_matches_enqueue(dev):
    """The end-to-end entry point bench.py's e2e leg uses (pinned per-scan host buffers -> strided H2D -> batch ->
    D2H of result rows, level records and per-scan vertex boundaries) returns what the device-resident path does."""
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    B, n = 3, 16384
    clouds = [torch.from_numpy(synth.synth_scan(40 + b, "os1-64-16k")) for b in range(B)]
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(32, n, generator=g) for _ in range(B)]
    weights = make_enet_weights(synth.ENET_BCL, seed=3)
    pipe = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=4.0, batch=B)
    pipe.enqueue(torch.cat(clouds, 1).to(dev), torch.cat(feats, 1).to(dev))
    tot = pipe.counts()
    want_vs = pipe.vertex_starts()
    want_Z = pipe.levels[-1]["Z"][:tot[-1]].cpu()
    out = torch.empty((tot[-1], 256)).pin_memory()
    st = torch.empty((5, 24), dtype=torch.int32).pin_memory()
    vs = torch.empty((5, B + 1), dtype=torch.int32).pin_memory()
    stream = torch.cuda.Stream(dev)
    for use_graph in (False, True):
        out.zero_()
        pipe.forward_host([c.pin_memory() for c in clouds], [f.pin_memory() for f in feats], out, st, stream=stream,
                          use_graph=use_graph, starts_host=vs)
        stream.synchronize()
        assert [int(v) for v in st[:, 1]] == tot
        assert vs.tolist() == want_vs
        assert H.rel_err(out.numpy(), want_Z.numpy()) < 1e-6          # same kernels; only the atomics' order differs


@pytest.mark.parametrize("l0_gather", [False, True])
@pytest.mark.parametrize("name,batch", [("leaky", 1), ("relu", 1), ("leaky", 3)])
def test_fused_stem_pipeline(name, batch, l0_gather, dev, golden_dir):
    """E-Net's pointwise stem computed inside the level-0 splat.  Level-0 BCL output with the stem fused
    (only the cloud goes in) must match the float64 oracle fed with the reference-golden stem weights:
    oracle stem -> cat(el_minus_gr, stem) -> oracle BCL, within the per-layer tolerance."""
    import os
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    from oracle import lattice as ol, bcl as obcl
    g = np.load(os.path.join(golden_dir, "stem_%s.npz" % name))
    leaky = bool(g["leaky"])
    layers = [(torch.from_numpy(g["W%d" % i]), torch.from_numpy(g["b%d" % i])) for i in range(3)]
    # the oracle's stem itself is pinned to the live reference's output (tests/test_oracle_golden.py); here on a full cloud
    n = 16384
    clouds = [synth.synth_scan(50 + b, "os1-64-16k") for b in range(batch)]
    weights = make_enet_weights(synth.ENET_BCL, seed=4)
    # l0_gather: the stem as a kernel of its own writing point-major rows + gather-form splat at level 0 as well
    pipe = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=4.0, batch=batch, stem=(layers, leaky),
                        level0_gather=l0_gather)
    assert pipe.gs0 == l0_gather
    pc_all = torch.from_numpy(np.concatenate(clouds, axis=1)).to(dev)
    for _ in range(2):
        pipe.enqueue(pc_all, None)
    pipe.counts()
    for b in range(batch):
        want = ol.generate(clouds[b], synth.SCALE_MAP)
        stem64 = obcl.stem_forward(clouds[b], layers, leaky=leaky, dtype=torch.float64)
        prev = stem64[None]
        outs = pipe.outputs(scan=b) if batch > 1 else pipe.outputs()
        for li, w in enumerate(want[:2]):
            args = (torch.from_numpy(w["pc1_barycentric"]), torch.from_numpy(w["pc1_lattice_offset"]),
                    torch.from_numpy(w["pc1_blur_neighbors"]), weights[li])
            ref = obcl.bcl_forward(torch.cat((torch.from_numpy(w["pc1_el_minus_gr"]).double(), prev), 1), *args, dtype=torch.float64)
            got_l = outs[li].cpu()
            e = H.rel_err(got_l.numpy(), ref.numpy())
            assert e < PER_LAYER_TOL, "fused stem, scan %d level %d rel err %g" % (b, li, e)
            prev = got_l.double()


def test_neighbour_symmetry_flag(dev):
    """GenerateData(exact=False) tags blur_neighbors with `_efgh_symmetric` (status bit EFGH_ST_ALIASED clear): whenever
    the tag says symmetric, the brute-force device check must agree - BilateralConvFlex.backward relies on it to use
    the gather-form data gradient.  Tiny / degenerate clouds are the ones whose neighbour keys leave the key box."""
    from efgh_b200.generate_data import GenerateData, blur_offsets
    from efgh_b200.bilateralNN import neighbours_symmetric
    offs = [tuple(o) for o in blur_offsets(1, 3).tolist()]
    mirror = tuple(offs.index(tuple(-v for v in o)) for o in offs)
    rng = np.random.default_rng(0)
    seen = {True: 0, False: 0}
    clouds = [rng.standard_normal((3, n)).astype(np.float32) * s for n, s in ((2, 1.0), (5, 0.01), (7, 3.0), (33, 100.0), (300, 0.5), (4000, 20.0))]
    clouds.append(synth.synth_scan(3, "os1-64-16k"))
    for pc in clouds:
        gd = GenerateData(3, synth.SCALE_MAP, "cuda", exact=False)
        _, data = gd(torch.from_numpy(pc).to(dev))
        for d in data:
            nbr = d["pc1_blur_neighbors"]
            tag = getattr(nbr, "_efgh_symmetric", None)
            if tag is None:                 # sparse cloud: GenerateData fell back to exact mode, which does not tag
                continue
            seen[bool(tag)] += 1
            if tag:
                assert neighbours_symmetric(nbr, mirror)
    print("symmetric-tagged levels: %d, aliased levels: %d" % (seen[True], seen[False]))
    assert seen[True] > 0


@pytest.mark.parametrize("batch", [1, 2])
def test_gather_splat_heavy_vertices(batch, dev):
    """A much coarser second level gives its vertices hundreds of contributions each: the gather-form splat's
    CTA-per-vertex path (more than 128 contributions) and its per-warp path must both match the oracle."""
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    from oracle import lattice as ol, bcl as obcl
    smap = [[2.0, 1], [0.05, 1]]
    plan = [(36, [32, 32]), (36, [32, 32])]
    n = 8192
    clouds = [synth.synth_scan(60 + b, "os1-64-16k")[:, :n] for b in range(batch)]
    weights = make_enet_weights(plan, seed=6)
    pipe = ScanPipeline(n, smap, plan, weights, dev, vertex_cap_factor=4.0, batch=batch)
    g = torch.Generator().manual_seed(3)
    feats = [torch.randn(32, n, generator=g) for _ in range(batch)]
    pipe.enqueue(torch.from_numpy(np.concatenate(clouds, 1)).to(dev), torch.cat(feats, 1).to(dev))
    pipe.counts()
    heavy = int(pipe.levels[1]["voff"][pipe.levels[1]["h_cap"] + 1])
    assert heavy > 0, "test cloud produced no heavy vertex"
    for b in range(batch):
        want = ol.generate(clouds[b], smap)
        got = _to_np(pipe.level_dicts(scan=b) if batch > 1 else pipe.level_dicts())
        outs = pipe.outputs(scan=b) if batch > 1 else pipe.outputs()
        prev = feats[b][None].double()
        for li, w in enumerate(want):
            H.assert_level_equal(got[li], w, "heavy scan %d L%d" % (b, li))
            ref = obcl.bcl_forward(torch.cat((torch.from_numpy(w["pc1_el_minus_gr"]).double(), prev), 1),
                                   torch.from_numpy(w["pc1_barycentric"]), torch.from_numpy(w["pc1_lattice_offset"]),
                                   torch.from_numpy(w["pc1_blur_neighbors"]), weights[li], dtype=torch.float64)
            got_l = outs[li].cpu()
            e = H.rel_err(got_l.numpy(), ref.numpy())
            assert e < PER_LAYER_TOL, "heavy scan %d level %d rel err %g" % (b, li, e)
            prev = got_l.double()


def test_batched_pipeline_mixed_sensors_fullsize(dev):
    """BASELINE config 5 clouds in ONE ragged batch: nuScenes-32-like (34 816 pts), OS1-64 (65 536) and HDL-64-like
    (122 880) scans through one launch sequence; per scan the lattice equals the C oracle bit for bit and the
    first two BCL levels match the float64 oracle."""
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    from oracle import lattice as ol, bcl as obcl
    clouds = [synth.synth_scan(70, "nusc-32"), synth.synth_scan(71, "os1-64-64k"), synth.synth_scan(72, "hdl-64")]
    sizes = [c.shape[1] for c in clouds]
    n = max(sizes)
    B = len(clouds)
    weights = make_enet_weights(synth.ENET_BCL, seed=8)
    pipe = ScanPipeline(n, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=2.0, batch=B)
    pipe.set_scan_sizes(sizes)
    g = torch.Generator().manual_seed(9)
    feats = [torch.randn(32, sz, generator=g) for sz in sizes]
    pc_all = torch.zeros(3, B * n)
    ft_all = torch.zeros(32, B * n)
    o = 0
    for b in range(B):
        pc_all[:, o:o + sizes[b]] = torch.from_numpy(clouds[b])
        ft_all[:, o:o + sizes[b]] = feats[b]
        o += sizes[b]
    pipe.enqueue(pc_all.to(dev), ft_all.to(dev))
    pipe.counts()
    for b in range(B):
        want = ol.generate(clouds[b], synth.SCALE_MAP)
        got = _to_np(pipe.level_dicts(scan=b))
        for li, (gl, wl) in enumerate(zip(got, want)):
            H.assert_level_equal(gl, wl, "mixed batch scan %d L%d" % (b, li))
        outs = pipe.outputs(scan=b)
        prev = feats[b][None].double()
        for li, w in enumerate(want[:2]):
            ref = obcl.bcl_forward(torch.cat((torch.from_numpy(w["pc1_el_minus_gr"]).double(), prev), 1),
                                   torch.from_numpy(w["pc1_barycentric"]), torch.from_numpy(w["pc1_lattice_offset"]),
                                   torch.from_numpy(w["pc1_blur_neighbors"]), weights[li], dtype=torch.float64)
            got_l = outs[li].cpu()
            e = H.rel_err(got_l.numpy(), ref.numpy())
            assert e < PER_LAYER_TOL, "mixed batch scan %d level %d rel err %g" % (b, li, e)
            prev = got_l.double()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_batched_lattice_random_ragged_property(seed, dev):
    """Property: for random ragged batches (scan sizes 1 .. 3000, random clouds of random spread, up to 9 scans) the
    batched lattice build equals the C oracle scan by scan, bit for bit - whatever the mix of tiny and larger scans."""
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    from oracle import lattice as ol
    rng = np.random.default_rng(100 + seed)
    B = int(rng.integers(2, 10))
    n = 3000
    sizes = [int(rng.integers(1, n + 1)) for _ in range(B)]
    sizes[int(rng.integers(0, B))] = 1                                  # always one single-point scan
    clouds = [(rng.standard_normal((3, sz)) * float(rng.choice([0.05, 2.0, 30.0]))).astype(np.float32) for sz in sizes]
    smap = synth.SCALE_MAP[:3]
    plan = synth.ENET_BCL[:3]
    pipe = ScanPipeline(n, smap, plan, make_enet_weights(plan, seed=1), dev, vertex_cap_factor=64.0, batch=B)
    pipe.set_scan_sizes(sizes)
    pc_all = torch.zeros(3, B * n)
    o = 0
    for b in range(B):
        pc_all[:, o:o + sizes[b]] = torch.from_numpy(clouds[b])
        o += sizes[b]
    pipe.enqueue(pc_all.to(dev), torch.zeros(32, B * n, device=dev))
    pipe.counts()
    for b in range(B):
        want = ol.generate(clouds[b], smap)
        got = _to_np(pipe.level_dicts(scan=b))
        for li, (gl, wl) in enumerate(zip(got, want)):
            H.assert_level_equal(gl, wl, "random batch seed %d scan %d (n=%d) L%d" % (seed, b, sizes[b], li))


# ------------------------------------------------------------------------------------------------
# wider coverage: BASELINE.json config 5 sweep, slice path, backward at E-Net shapes, radius 2, determinism
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sensor", ["nusc-32", "os1-64-64k", "hdl-64"])
def test_lattice_scale_sweep_config5(sensor, dev):
    """Config 5: number of scales 1..5 (prefixes of the scale map) on the three sensor densities."""
    from oracle import lattice as ol
    pc = synth.synth_scan(9, sensor)
    for nscales in range(1, 6):
        smap = synth.SCALE_MAP[:nscales]
        want = ol.generate(pc, smap)
        got = _run(pc, smap, dev, exact=False)
        assert len(got) == nscales
        for li, (g, w) in enumerate(zip(got, want)):
            H.assert_level_equal(g, w, "%s scales=%d L%d" % (sensor, nscales, li))


def test_lattice_deterministic_and_stream_safe(dev):
    """Same cloud twice, and two clouds interleaved on two streams: identical results (no hidden global state)."""
    from efgh_b200.generate_data import GenerateData
    pcs = [torch.from_numpy(synth.synth_scan(s, "os1-64-16k")).to(dev) for s in (21, 22)]
    gds = [GenerateData(3, synth.SCALE_MAP, "cuda", exact=False) for _ in range(2)]
    base = [_to_np(gds[i](pcs[i])[1]) for i in range(2)]
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    outs = [None, None]
    for rep in range(3):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                outs[i] = gds[i](pcs[i])[1]
    torch.cuda.synchronize(dev)
    for i in range(2):
        for li, (g, w) in enumerate(zip(_to_np(outs[i]), base[i])):
            H.assert_level_equal(g, w, "stream %d L%d" % (i, li))


def _enet_level(dev, n_points=16384, cin=36, nout=(32, 32), do_slice=False, seed=5, radius=1, scale=1.0):
    from efgh_b200.generate_data import GenerateData
    from efgh_b200.bilateralNN import BilateralConvFlex
    pc = synth.synth_scan(seed, "os1-64-16k")[:, :n_points]
    gd = GenerateData(3, [[scale, radius]], "cuda", exact=True)
    _, data = gd(torch.from_numpy(pc).to(dev))
    d = data[0]
    torch.manual_seed(seed)
    m = BilateralConvFlex(3, radius, cin, list(nout), "cuda", True, True, True, True, do_slice, do_slice, chunk_size=-1).to(dev)
    for p in m.parameters():
        torch.nn.init.normal_(p, 0, 0.1)
    feat = torch.randn(1, cin, pc.shape[1], device=dev)
    return m, d, feat


def _oracle_out(m, d, feat, do_slice, dtype=torch.float64, requires_grad=False):
    from oracle import bcl as obcl
    convs = [(c.weight.detach().cpu().clone().requires_grad_(requires_grad), c.bias.detach().cpu().clone().requires_grad_(requires_grad))
             for c in m.blur_conv if isinstance(c, torch.nn.Conv2d)]
    f = feat.detach().cpu().double().requires_grad_(requires_grad)
    sb = m.bias.detach().cpu().clone().requires_grad_(requires_grad) if do_slice else None
    out = obcl.bcl_forward(f, d["pc1_barycentric"].cpu(), d["pc1_lattice_offset"].cpu(), d["pc1_blur_neighbors"].cpu(), convs,
                           do_slice=do_slice, out_bary=d["pc1_barycentric"].cpu() if do_slice else None,
                           out_off=d["pc1_lattice_offset"].cpu() if do_slice else None, slice_bias=sb,
                           last_relu=m.last_relu, use_leaky=m.use_leaky, dtype=dtype)
    return out, f, convs, sb


@pytest.mark.parametrize("do_slice", [False, True])
def test_bcl_enet_shape_forward_backward_vs_oracle(do_slice, dev):
    """E-Net level-0 shape (36 -> [32,32]) on a 16k cloud, forward AND backward (autograd of the float64 oracle),
    with and without the slice stage (reference bilateralNN.py:251-261).  Gradient tolerance 2e-5."""
    m, d, feat = _enet_level(dev, do_slice=do_slice)
    feat.requires_grad_(True)
    args = (d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"],
            d["pc1_barycentric"] if do_slice else None, d["pc1_lattice_offset"] if do_slice else None)
    out = m(feat, *args)
    ref, f_ref, convs, sb = _oracle_out(m, d, feat, do_slice, requires_grad=True)
    assert tuple(out.shape) == tuple(ref.shape)
    assert H.rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < PER_LAYER_TOL
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    out.backward(gout.to(dev))
    ref.backward(gout.double())
    assert H.rel_err(feat.grad.cpu().numpy(), f_ref.grad.numpy()) < 2e-5
    mods = [c for c in m.blur_conv if isinstance(c, torch.nn.Conv2d)]
    for c, (W, b) in zip(mods, convs):
        assert H.rel_err(c.weight.grad.cpu().numpy(), W.grad.numpy()) < 2e-5
        assert H.rel_err(c.bias.grad.cpu().numpy(), b.grad.numpy()) < 2e-5
    if do_slice:
        assert H.rel_err(m.bias.grad.cpu().numpy(), sb.grad.numpy()) < 2e-5


def test_bcl_radius2_filter65(dev):
    """neighborhood_size 2 -> 65 filter taps (reference get_filter_size), tensor-core path, vs the float64 oracle."""
    m, d, feat = _enet_level(dev, n_points=3000, cin=32, nout=(32, 64), radius=2, scale=0.5)
    assert d["pc1_blur_neighbors"].shape[1] == 65
    with torch.no_grad():
        out = m(feat, d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
    ref, _, _, _ = _oracle_out(m, d, feat, False)
    assert H.rel_err(out.cpu().numpy(), ref.numpy()) < PER_LAYER_TOL


@pytest.mark.parametrize("precision,tol", [("3xtf32", 1e-5), ("fp32", 1e-5), ("tf32", 5e-3)])
def test_bcl_precision_modes(precision, tol, dev, monkeypatch):
    """Stated tolerances of the three arithmetic modes of the lattice convolution."""
    from efgh_b200 import bilateralNN
    monkeypatch.setattr(bilateralNN, "CONV_PRECISION", precision)
    m, d, feat = _enet_level(dev, n_points=8192, cin=68, nout=(128, 128))
    with torch.no_grad():
        out = m(feat, d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
    ref, _, _, _ = _oracle_out(m, d, feat, False)
    err = H.rel_err(out.cpu().numpy(), ref.numpy())
    print("%s rel err %.2e" % (precision, err))
    assert err < tol


def test_bcl_partition_of_unity_fullsize(dev):
    """Size-independent property at the full 131k size: splatting a constant 1 with density normalisation, an
    identity filter (centre tap only) and slicing it back returns 1 at every point (barycentric weights sum to 1,
    the normalised splat of a constant is the constant)."""
    from efgh_b200.generate_data import GenerateData
    from efgh_b200.bilateralNN import BilateralConvFlex
    pc = synth.synth_scan(6, "os1-64")
    gd = GenerateData(3, [[1.0, 1]], "cuda", exact=False)
    _, data = gd(torch.from_numpy(pc).to(dev))
    d = data[0]
    C = 32
    m = BilateralConvFlex(3, 1, C, [C, C], "cuda", False, True, True, True, True, False, chunk_size=-1).to(dev)
    with torch.no_grad():
        m.blur_conv[0].weight.zero_(); m.blur_conv[0].bias.zero_()
        m.blur_conv[2].weight.zero_(); m.blur_conv[2].bias.zero_()
        for c in range(C):
            m.blur_conv[0].weight[c, c, 0, 0] = 1.0      # tap 0 is the vertex itself
            m.blur_conv[2].weight[c, c, 0, 0] = 1.0
        feat = torch.ones(1, C, pc.shape[1], device=dev)
        out = m(feat, d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"],
                d["pc1_barycentric"], d["pc1_lattice_offset"])
    assert tuple(out.shape) == (1, C, pc.shape[1])
    assert float((out - 1).abs().max()) < 2e-4       # 1e-5 in the normaliser's denominator, fp32 sums


@pytest.mark.gpu
def test_scatter_kernel_variants_in_subprocess():
    """The level-0 splat has two kernels (tile: CTA-wide tile + barriers; warp: warp-private tiles) of which the library
    picks one per case (warp with the fused stem, tile otherwise); EFGH_SCATTER forces either for every case.  The
    choice is latched at the first call, so each forced variant runs the oracle comparisons in a fresh process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for variant in ("warp", "tile"):
        env = dict(os.environ, EFGH_SCATTER=variant)
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_gpu_parity.py"),
                            "-k", "test_scan_pipeline_vs_oracle or test_fused_stem_pipeline or test_batched_pipeline_matches_single_scans"],
                           cwd=root, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, "EFGH_SCATTER=%s:\n%s\n%s" % (variant, r.stdout[-3000:], r.stderr[-2000:])
