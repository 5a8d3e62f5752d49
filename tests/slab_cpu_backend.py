"""CPU stand-in for the slab compute backend (TESTS ONLY): the oracle does the arithmetic,
so watercube_b200.slab's decomposition / exchange logic can be checked without a GPU
(in-process virtual ranks and a real 2-process gloo run).  Same interface and the same
message formats as watercube_b200.slab.CudaSlabBackend."""
import numpy as np
import torch

from watercube_b200.slab import layer_of

f32 = np.float32
HEADER_WORDS = 8  # 32-byte migrant message header: [count, 7 x pad]


class OracleSlabBackend:
    def __init__(self, oracle, params, z_begin, z_end, migrant_capacity, nthreads=1):
        self.ob, self.p = oracle, params
        self.d = oracle.derive(params)
        self.G = int(params.grid_res)
        self.z_begin, self.z_end, self.M = z_begin, z_end, migrant_capacity
        self.nthreads = nthreads
        self.own = np.zeros((0, 8), f32)          # buffer 1, owned region
        self.sorted = np.zeros((0, 8), f32)       # buffer 2, owned region
        G2 = self.G * self.G
        self._mig = {k: [np.zeros(HEADER_WORDS + 8 * self.M, np.uint32) for _ in (0, 1)]
                     for k in ("out", "in")}
        self._lc = {k: [np.zeros(1 + G2, np.uint32) for _ in (0, 1)] for k in ("send", "recv")}
        self._halo = {}
        self.info = None

    # -- data
    def close(self):
        pass

    def upload(self, particles):
        self.own = np.ascontiguousarray(particles, f32).reshape(-1, 8).copy()

    def download(self, which=1):
        return (self.own if which == 1 else self.sorted).copy()

    @property
    def num_particles(self):
        return len(self.own)

    def sync(self, after_copy=False):
        pass

    @staticmethod
    def _t(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1))

    def clear_recv(self, d):
        self._mig["in"][d][:] = 0
        self._lc["recv"][d][:] = 0

    def lc_send(self, d):
        return self._t(self._lc["send"][d])

    def lc_recv(self, d):
        return self._t(self._lc["recv"][d])

    def mig_out(self, d):
        return self._t(self._mig["out"][d])

    def mig_in(self, d):
        return self._t(self._mig["in"][d])

    def _layer(self, P):
        return layer_of(P[:, 2], self.d.bin_size, self.G)

    # -- phases
    def _migrants(self, d):
        buf = self._mig["in"][d]
        n = int(buf[0])
        return buf[HEADER_WORDS:HEADER_WORDS + 8 * n].view(f32).reshape(n, 8)

    def sort_count(self):
        lay = self._layer(self.own)
        stay = self.own[(lay >= self.z_begin) & (lay < self.z_end)]
        below, above = self._migrants(0), self._migrants(1)
        self.errors = 0
        for m in (below, above):
            l = self._layer(m)
            self.errors += int(((l < self.z_begin) | (l >= self.z_end)).sum())
        self.virtual = np.concatenate([below, stay, above])      # order matters (stable sort)
        self.m_in = (len(below), len(above))
        s = self.ob.sort(self.virtual, self.d.bin_size, self.G)
        self.sort_out = s
        G2 = self.G * self.G
        self.sorted = self.ob.as_f32(s["sorted"]).copy()
        c = s["counts"].reshape(self.G, G2)
        n_first, n_last = int(c[self.z_begin].sum()), int(c[self.z_end - 1].sum())
        self._lc["send"][0][:] = np.concatenate([[n_first], c[self.z_begin]])
        self._lc["send"][1][:] = np.concatenate([[n_last], c[self.z_end - 1]])
        self._n = (len(self.sorted), n_first, n_last)

    def sync_info(self):
        n, n_first, n_last = self._n
        self.info = dict(n_owned=n, n_first=n_first, n_last=n_last,
                         n_ghost_below=int(self._lc["recv"][0][0]),
                         n_ghost_above=int(self._lc["recv"][1][0]), errors=self.errors,
                         migrants_in_below=self.m_in[0], migrants_in_above=self.m_in[1])
        return self.info

    def reorder(self):
        i = self.info
        self._halo = {("pos", 0): np.zeros((i["n_ghost_below"], 4), f32),
                      ("vel", 0): np.zeros((i["n_ghost_below"], 4), f32),
                      ("pos", 1): np.zeros((i["n_ghost_above"], 4), f32),
                      ("vel", 1): np.zeros((i["n_ghost_above"], 4), f32)}

    def halo_send(self, d, kind):
        i = self.info
        rows = slice(0, i["n_first"]) if d == 0 else slice(i["n_owned"] - i["n_last"], i["n_owned"])
        cols = slice(0, 4) if kind == "pos" else slice(4, 8)
        return self._t(np.ascontiguousarray(self.sorted[rows, cols]))

    def halo_recv(self, d, kind):
        return self._t(self._halo[(kind, d)])

    def _combined(self):
        """[ghost-low | owned | ghost-high] with a whole-grid table (only local layers filled)."""
        gl = np.concatenate([self._halo[("pos", 0)], self._halo[("vel", 0)]], axis=1)
        gh = np.concatenate([self._halo[("pos", 1)], self._halo[("vel", 1)]], axis=1)
        allp = np.ascontiguousarray(np.concatenate([gl, self.sorted, gh]))
        counts = self.sort_out["counts"].copy()
        G2 = self.G * self.G
        if self.z_begin > 0:
            counts[(self.z_begin - 1) * G2:self.z_begin * G2] = self._lc["recv"][0][1:]
        if self.z_end < self.G:
            counts[self.z_end * G2:(self.z_end + 1) * G2] = self._lc["recv"][1][1:]
        offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint32)
        return allp, counts, offsets, len(gl)

    def density(self):
        allp, counts, offsets, g0 = self._combined()
        P, _ = self.ob.density(allp, counts, offsets, self.p, nthreads=self.nthreads)
        self.sorted = self.ob.as_f32(P)[g0:g0 + len(self.sorted)].copy()

    def update(self, frame_dt):
        allp, counts, offsets, g0 = self._combined()
        dt = f32(frame_dt) * f32(self.p.time_scale)
        out, _ = self.ob.update(allp, counts, offsets, self.p, dt, nthreads=self.nthreads)
        self.own = self.ob.as_f32(out)[g0:g0 + len(self.sorted)].copy()
        lay = self._layer(self.own)
        for d, sel in ((0, lay < self.z_begin), (1, lay >= self.z_end)):
            m = self.own[sel]                                 # boolean mask keeps the order
            assert len(m) <= self.M, "migrant capacity"
            buf = self._mig["out"][d]
            buf[:] = 0
            buf[0] = len(m)
            buf[HEADER_WORDS:HEADER_WORDS + 8 * len(m)] = m.reshape(-1).view(np.uint32)
