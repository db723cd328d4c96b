"""GenerateData - drop-in for reference nets/generate_data.py:7-198, running on the GPU.

Same constructor, same `__call__(pc1) -> (pc1, [dict]*levels)` and the same dict keys / dtypes / shapes
(leading batch dimension 1, int64 indices, `pc1_hash_cnt` a Python int), so Enet.forward
(reference nets/enet.py:107-141) consumes it unchanged.  The cloud never leaves the device: the
reference's D2H at generate_data.py:122 and the four H2D copies per level (:181-184) are gone.

Two modes:
  exact=True  (default) one small D2H read of the vertex count per level, outputs are exact-size
              contiguous tensors - bit-for-bit what the reference returns;
  exact=False no host round trip inside the scan: buffers are allocated at capacity, the level records
              are read back once at the end and the outputs are narrowed views of those buffers.
"""
import itertools
import math

import numpy as np
import torch

from . import _capi

STATE_WORDS = 24


def blur_offsets(radius, d=3):
    """(F, d+1) neighbour offsets in the reference's traversal order (nets/transforms.py:95-122):
    step counts (i_0..i_d) in [0, radius] with at least one zero, last index fastest; a step in
    dimension k adds d+1 to coordinate k and subtracts 1 from all (transforms.py:81-87)."""
    d1 = d + 1
    rows = [[d1 * s - sum(steps) for s in steps]
            for steps in itertools.product(range(radius + 1), repeat=d1) if 0 in steps]
    return np.asarray(rows, dtype=np.int64)


class LatticeStatusError(RuntimeError):
    pass


class VertexCapExceeded(LatticeStatusError):
    pass


ST_ALIASED = 8      # informational bit of efgh_lattice_state.status (include/efgh_b200.h)


def check_status(status, level):
    status &= ~ST_ALIASED
    if status:
        why = []
        if status & 1:
            why.append("lattice coordinate outside +-2^20 (cloud far outside the supported range)")
        if status & 2:
            why.append("more lattice vertices than the capacity (raise vertex_cap_factor or use exact=True)")
        if status & 4:
            why.append("hash table full")
        raise LatticeStatusError("lattice level %d: %s" % (level, "; ".join(why)))


class GenerateData(object):
    def __init__(self, dim, scales_filter_map, device, exact=True, vertex_cap_factor=4.0):
        if dim != 3:
            raise NotImplementedError("efgh_b200 implements the d=3 lattice only (reference configs use dim: 3)")
        self.d0 = dim
        self.d1 = dim + 1
        self.scales_filter_map = scales_filter_map
        self.device = torch.device("cuda" if device in ("cuda", None) else device)
        if self.device.type != "cuda":
            raise _capi.EfghError("GenerateData runs on CUDA only; there is no CPU path (got device=%r)" % (device,))
        self.exact = exact
        self.vertex_cap_factor = vertex_cap_factor
        self.expected_std = (self.d0 + 1) * math.sqrt(2 / 3)          # generate_data.py:19
        # kept for parity with the reference's attributes (generate_data.py:20,30,52)
        left = torch.ones((self.d1, self.d0), dtype=torch.float32).triu()
        left[1:, ] += torch.diag(torch.arange(-1, -self.d0 - 1, -1, dtype=torch.float32))
        right = torch.diag(1. / (torch.arange(1, self.d0 + 1, dtype=torch.float32) *
                                 torch.arange(2, self.d0 + 2, dtype=torch.float32)).sqrt())
        self.elevate_mat = torch.mm(left, right)
        self.canonical = torch.tensor([[j if j <= self.d0 - i else j - self.d1 for j in range(self.d1)]
                                       for i in range(self.d1)], dtype=torch.long)
        self.radius2offset = {}
        self._offsets_dev = {}
        for radius in set(item for line in scales_filter_map for item in line[1:] if item != -1):
            self.radius2offset[radius] = blur_offsets(radius, self.d0)
        self._workspace = {}          # per (device, stream): instances may be shared by concurrent streams
        self.last_states = None

    def get_filter_size(self, radius):
        return (radius + 1) ** self.d1 - radius ** self.d1          # generate_data.py:114-115

    # -- internals ---------------------------------------------------------------------------
    def _offsets(self, radius, device):
        key = (radius, device)
        if key not in self._offsets_dev:
            self._offsets_dev[key] = torch.from_numpy(self.radius2offset[radius].astype(np.int32)).to(device)
        return self._offsets_dev[key]

    def _ws(self, n_cap, device):
        need = _capi.lib().efgh_lattice_workspace_bytes(int(n_cap))
        key = (device, _capi.stream_ptr())
        ws = self._workspace.get(key)
        if ws is None or ws.numel() < need:
            if len(self._workspace) >= 8:
                self._workspace.clear()
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._workspace[key] = ws
        return ws

    def __call__(self, pc1):
        L = _capi.lib()
        with torch.no_grad():
            dev = pc1.device if pc1.is_cuda else self.device
            pc1 = pc1.to(device=dev, dtype=torch.float32)              # generate_data.py:122 (stays on the GPU)
            with torch.cuda.device(dev):
                try:
                    return pc1, self._build(L, pc1, dev, self.exact)
                except VertexCapExceeded:
                    # sparse cloud: more vertices than vertex_cap_factor * N; redo with per-level sizing
                    return pc1, self._build(L, pc1, dev, True)

    def _build(self, L, pc1, dev, exact):
        stream = _capi.stream_ptr()
        nlev = len(self.scales_filter_map)
        states = torch.zeros((nlev, STATE_WORDS), dtype=torch.int32, device=dev)   # (reserved words are never written by the kernels)
        pts = pc1[:3]
        if pts.stride(-1) != 1:
            pts = pts.contiguous()
        n = pts.shape[-1]
        n0 = n
        n_dev = None                     # device-side count of the current level's points (None = n is exact)
        out = []
        filt = []
        for li, (scale, radius) in enumerate(self.scales_filter_map):
            has_next = li != nlev - 1
            F = self.get_filter_size(radius) if radius != -1 else 0
            st = states[li]
            ws = self._ws(n, dev)
            h_cap = 4 * n if exact else int(min(4 * n, max(self.vertex_cap_factor * n0, 1024)))
            bary = torch.empty((1, self.d1, n), dtype=torch.float32, device=dev)
            elmgr = torch.empty((1, self.d1, n), dtype=torch.float32, device=dev)
            _capi.check(L.efgh_lattice_points(pts.data_ptr(), max(pts.stride(0), n, 1), n, _capi.ptr(n_dev),
                                              float(scale), bary.data_ptr(), elmgr.data_ptr(), max(n, 1), h_cap,
                                              st.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                        "efgh_lattice_points")
            if exact:
                st_host = st.cpu()                                    # per-level sync: pc1_hash_cnt is a Python int
                check_status(int(st_host[2]), li)
                H = int(st_host[1])
                h_alloc = H
            else:
                H = None
                h_alloc = h_cap
            loff = torch.empty((1, self.d1, n), dtype=torch.int64, device=dev)
            nbr = torch.empty((1, F, h_alloc), dtype=torch.int64, device=dev) if F > 0 else None
            offs = self._offsets(radius, dev) if F > 0 else None
            nxt = torch.empty((3, h_alloc), dtype=torch.float32, device=dev) if has_next else None
            divisor = float(np.float32(self.expected_std * scale))   # generate_data.py:177
            _capi.check(L.efgh_lattice_vertices(n, loff.data_ptr(), None, max(n, 1), _capi.ptr(offs), F, h_alloc,
                                                _capi.ptr(nbr), None, max(h_alloc, 1), _capi.ptr(nxt), max(h_alloc, 1),
                                                divisor, st.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                        "efgh_lattice_vertices")
            if nbr is None:
                nbr = torch.zeros((1, 1), dtype=torch.int64, device=dev)  # generate_data.py:173,184
            out.append({"pc1_barycentric": bary, "pc1_el_minus_gr": elmgr, "pc1_lattice_offset": loff,
                        "pc1_blur_neighbors": nbr, "pc1_hash_cnt": H})
            filt.append(F)
            if has_next:
                pts = nxt
                n = h_alloc
                n_dev = None if exact else st[1:2]
        if not exact:
            host = states.cpu()                                       # the only sync of the scan
            n_true = n0
            if any(int(host[li, 2]) & 2 for li in range(nlev)):
                raise VertexCapExceeded()
            for li, d in enumerate(out):
                check_status(int(host[li, 2]), li)
                H = int(host[li, 1])
                d["pc1_hash_cnt"] = H
                for k in ("pc1_barycentric", "pc1_el_minus_gr", "pc1_lattice_offset"):
                    d[k] = d[k][:, :, :n_true]
                if filt[li] > 0:
                    d["pc1_blur_neighbors"] = d["pc1_blur_neighbors"][:, :, :H]
                    # no aliased lookup -> the table is mirror-symmetric: BilateralConvFlex.backward may use the
                    # gather-form data gradient without checking on the device
                    d["pc1_blur_neighbors"]._efgh_symmetric = not (int(host[li, 2]) & ST_ALIASED)
                n_true = H
        self.last_states = states
        return out

    def __repr__(self):
        return self.__class__.__name__ + '\n(scales_filter_map: {}\n)'.format(self.scales_filter_map)
