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


@pytest.mark.parametrize("sensor,factor", [("os1-64-16k", 4.0), ("os1-64", 1.0)])
def test_scan_pipeline_vs_oracle(sensor, factor, dev):
    """The sync-free whole-scan pipeline (what bench.py times): lattice dicts bit-exact against the C oracle,
    BCL outputs of all five levels within 1e-5 of the float64 oracle (config 1 and config 2 clouds)."""
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights
    from oracle import lattice as ol, bcl as obcl
    pc = synth.synth_scan(4, sensor)
    N = pc.shape[1]
    weights = make_enet_weights(synth.ENET_BCL, seed=3)
    pipe = ScanPipeline(N, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=factor)
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
