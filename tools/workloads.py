"""The five BASELINE.json configurations as synthetic workloads, shared by bench.py and the
full-size GPU tests (measurement / test support, not product code).

Every generator is slab-addressable: `device_slab(z0, z1, dev)` yields exactly the sample
slices [z0, z1) of the global grid on any GPU, so a z-slab rank generates only its own
slices (34 GB of cfg4 are never staged through one place) and a single-context extraction
of the whole grid sees bit-identical samples.  Host copies (for the CPU reference, which
cannot generate anything itself) are downloads of device slabs.

  cfg1  201^3 float  cos x + cos y + cos z on [-4,4]^3, step .04 (README flow), iso 0
  cfg2  512^3 float  gyroid, 4 periods, 8 isovalues -1.2 .. 0.9
  cfg3  1024^3 u16   CT-like: Gaussian blobs 1000..3500 + texture + noise 0..15, iso 1500 and 1500.5
  cfg4  2048^3 float smooth noise: trilinear up-sampling (dyadic weights) of a hashed lattice, iso 0
  cfg5  768^3 float  white noise in [-1,1) (counter based), inclined grid (spnC), iso 0
"""
import math

import numpy as np

M64 = (1 << 64) - 1


def _i64(v):
    """python int -> the same 64 bits as a signed int64 constant (torch has no uint64 arithmetic)"""
    v &= M64
    return v - (1 << 64) if v >> 63 else v


def mix64(x):
    """splitmix64 finaliser on int64 tensors (two's complement wrap-around = arithmetic mod 2^64)"""
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * _i64(0x94D049BB133111EB)
    return x ^ ((x >> 31) & ((1 << 33) - 1))


class Workload:
    name = ""
    variant = "f32"          # element type of the grid (drop-in library variant)
    shape = (0, 0, 0)        # samples: NZ, NY, NX
    isos = (0.0,)
    geom_kw = None           # kwargs of dropin.Geometry (None: unit grid, spn0)
    sweep = False            # the isovalues share one classify pass (mc33cu_classify_sweep)
    describe = ""

    @property
    def npts(self):
        return self.shape[0] * self.shape[1] * self.shape[2]

    @property
    def sample_bytes(self):
        return {"f32": 4, "f64": 8, "u8": 1, "u16": 2, "u32": 4}[self.variant]

    def geometry(self):
        from mc33_c_library_b200.dropin import Geometry
        return Geometry(**self.geom_kw) if self.geom_kw else None

    def device_slab(self, z0, z1, dev):
        raise NotImplementedError

    def host_slab(self, z0, z1, dev):
        return self.device_slab(z0, z1, dev).cpu().numpy()


class Cfg1(Workload):
    name = "cfg1"
    shape = (201, 201, 201)
    isos = (0.0,)
    geom_kw = dict(r0=(-4.0, -4.0, -4.0), d=(0.04, 0.04, 0.04))
    describe = "201^3 float grid of cos x + cos y + cos z on [-4,4]^3 (generate_grid_from_fn flow), iso 0"

    def _table(self):
        # generate_grid_from_fn accumulates x += dx in double (reference MC33_util_grd.c:661-672)
        xs, x = [], -4.0
        for _ in range(201):
            xs.append(x)
            x += 0.04
        return np.array([math.cos(v) for v in xs], dtype=np.float64)

    def device_slab(self, z0, z1, dev):
        import torch
        c = torch.from_numpy(self._table()).to(dev)
        return ((c[None, None, :] + c[None, :, None]) + c[z0:z1, None, None]).to(torch.float32).contiguous()


class Cfg2(Workload):
    name = "cfg2"
    isos = (-1.2, -0.9, -0.6, -0.3, 0.0, 0.3, 0.6, 0.9)
    sweep = True

    def __init__(self, n=512, nz=None):
        self.n = n
        self.shape = (nz or n, n, n)
        self.describe = f"{n}x{n}x{self.shape[0]} float gyroid (4 periods per {n} samples), iso sweep of 8 values"

    def _tables(self, z0, z1):
        w = 2.0 * math.pi * 4.0
        t = (np.arange(self.n, dtype=np.float64) / (self.n - 1) - 0.5) * w
        tz = (np.arange(z0, z1, dtype=np.float64) / (self.n - 1) - 0.5) * w     # the same spacing continues along z
        return np.sin(t), np.cos(t), np.sin(tz), np.cos(tz)

    def device_slab(self, z0, z1, dev):
        import torch
        s, c, sz, cz = (torch.from_numpy(a).to(dev) for a in self._tables(z0, z1))
        out = torch.empty((z1 - z0, self.n, self.n), dtype=torch.float32, device=dev)
        for k in range(0, z1 - z0, 64):     # chunked: keeps the float64 temporaries small
            e = min(k + 64, z1 - z0)
            out[k:e] = (s[None, None, :] * c[None, :, None] + s[None, :, None] * cz[k:e, None, None]
                        + sz[k:e, None, None] * c[None, None, :]).to(torch.float32)
        return out

    def host_slab(self, z0, z1, dev=None):
        s, c, sz, cz = self._tables(z0, z1)
        return (s[None, None, :] * c[None, :, None] + s[None, :, None] * cz[:, None, None]
                + sz[:, None, None] * c[None, None, :]).astype(np.float32)


class Cfg3(Workload):
    name = "cfg3"
    variant = "u16"
    isos = (1500.0, 1500.5)
    CH = 64

    def __init__(self, n=1024):
        self.n = n
        self.shape = (n, n, n)
        self.describe = (f"{n}^3 uint16 CT-like volume (GRD_INTEGER, GRD_TYPE_SIZE 2): Gaussian blobs 1000..3500 + texture + "
                         "uniform noise 0..15; iso 1500 (integer: on-iso samples) and 1500.5")

    def _chunk(self, c, dev):
        """slices [64c, 64c+64) -- the noise generator is seeded per chunk, so any slab can be produced alone"""
        import torch
        n = self.n
        z0, z1 = c * self.CH, min((c + 1) * self.CH, n)
        g = torch.Generator(device=dev)
        g.manual_seed(1000 + c)
        ax = torch.linspace(-1, 1, n, device=dev)
        az = ax[z0:z1]
        v = torch.full((z1 - z0, n, n), 1000.0, device=dev)
        for c0, c1, c2, sg in ((0.1, -0.2, 0.05, 0.3), (-0.4, 0.3, -0.1, 0.25), (0.5, 0.5, 0.1, 0.2), (-0.3, -0.5, 0.6, 0.35)):
            v += 2500.0 / 3 * torch.exp(-((ax[None, None, :] - c0) ** 2 + (ax[None, :, None] - c1) ** 2 + (az[:, None, None] - c2) ** 2) / (2 * sg * sg))
        v += 20.0 * torch.sin(37 * ax[None, None, :]) * torch.sin(29 * ax[None, :, None]) * torch.sin(31 * az[:, None, None])
        v += torch.randint(0, 16, v.shape, device=dev, generator=g)
        return v.clamp(0, 65535).to(torch.int32).to(torch.uint16)

    def device_slab(self, z0, z1, dev):
        import torch
        out = torch.empty((z1 - z0, self.n, self.n), dtype=torch.uint16, device=dev)
        for c in range(z0 // self.CH, (z1 - 1) // self.CH + 1):
            a, b = max(z0, c * self.CH), min(z1, (c + 1) * self.CH)
            out[a - z0:b - z0] = self._chunk(c, dev)[a - c * self.CH:b - c * self.CH]
        return out


class Cfg4(Workload):
    name = "cfg4"
    isos = (0.0,)
    CELL = 32

    def __init__(self, n=2048):
        self.n = n
        self.shape = (n, n, n)
        self.describe = (f"{n}^3 float smooth noise: trilinear up-sampling (dyadic weights, exact in float) of a hashed lattice "
                         f"with {self.CELL}-sample cells, iso 0")

    def _lattice(self, dev):
        import torch
        m = (self.n + self.CELL - 1) // self.CELL
        i = torch.arange(m + 1, device=dev, dtype=torch.int64)
        h = (i[:, None, None] * 73856093) ^ (i[None, :, None] * 19349663) ^ (i[None, None, :] * 83492791)
        h = (h * 2654435761) & 0xFFFFFFFF
        h = ((h >> 13) ^ h) * 1274126177 & 0xFFFFFFFF
        return ((h >> 8).to(torch.float32) / float(1 << 23)) - 1.0          # [z][y][x] in [-1, 1)

    def device_slab(self, z0, z1, dev):
        import torch
        n, CELL = self.n, self.CELL
        L = self._lattice(dev)
        idx = torch.arange(n, device=dev)
        c0 = (idx // CELL).long()
        t = (idx % CELL).to(torch.float32) / CELL
        out = torch.empty((z1 - z0, n, n), dtype=torch.float32, device=dev)
        cache = {}

        def plane(k):            # lattice slice k up-sampled in y and x -> [n][n]
            if k not in cache:
                if len(cache) > 2:
                    cache.pop(min(cache))
                P = L[k]
                px = P[:, c0] * (1 - t)[None, :] + P[:, c0 + 1] * t[None, :]
                cache[k] = px[c0, :] * (1 - t)[:, None] + px[c0 + 1, :] * t[:, None]
            return cache[k]

        for z in range(z0, z1):
            k, tz = z // CELL, (z % CELL) / CELL
            out[z - z0] = plane(k) * (1 - tz) + plane(k + 1) * tz
        return out


INCLINED_A = ((1.0, 0.3, 0.2), (0.0, 0.95, 0.1), (0.0, 0.0, 0.9))


class Cfg5(Workload):
    name = "cfg5"
    isos = (0.0,)

    def __init__(self, n=768):
        self.n = n
        self.shape = (n, n, n)
        A = np.array(INCLINED_A)
        self.geom_kw = dict(nonortho=1, A=A, Ai=np.linalg.inv(A))
        self.describe = (f"{n}^3 float uniform white noise in [-1,1) (counter-based hash of the sample index), iso 0, inclined grid "
                         "(nonortho = 1, upper-triangular _A, spnC store): ambiguity-heavy, every MC33 sub-case")

    def device_slab(self, z0, z1, dev):
        import torch
        n = self.n
        out = torch.empty((z1 - z0, n, n), dtype=torch.float32, device=dev)
        plane = torch.arange(n * n, device=dev, dtype=torch.int64)
        for z in range(z0, z1):
            h = mix64((plane + z * n * n) * _i64(0x9E3779B97F4A7C15) + 12345)
            out[z - z0] = (((h >> 40) & 0xFFFFFF).to(torch.float32) / float(1 << 23) - 1.0).view(n, n)
        return out


def make(name, **kw):
    return {"cfg1": Cfg1, "cfg2": Cfg2, "cfg3": Cfg3, "cfg4": Cfg4, "cfg5": Cfg5}[name](**kw)


# --------------------------------------------------------------------------
# order-independent mesh digests (device side): a sharded extraction and a
# single-context extraction of the same grid must give the same three numbers
# --------------------------------------------------------------------------
def _bits32(t):
    import torch
    return t.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF


def vertex_digest(vkey, V, N):
    """sum over vertices of hash(canonical key, position bits, normal bits), mod 2^64.
    vkey [n] int64; V [n,3] float32/float64; N [n,3] float32"""
    import torch
    n = vkey.shape[0]
    acc = 0
    for a in range(0, n, 1 << 25):
        b = min(n, a + (1 << 25))
        h = mix64(vkey[a:b])
        if V.dtype == torch.float64:
            vb = V[a:b].contiguous().view(torch.int64)
            for j in range(3):
                h = mix64(h ^ vb[:, j])
        else:
            vb = _bits32(V[a:b]).view(-1, 3)
            h = mix64(h ^ vb[:, 0] ^ (vb[:, 1] << 32))
            h = mix64(h ^ vb[:, 2])
        nb = _bits32(N[a:b]).view(-1, 3)
        h = mix64(h ^ nb[:, 0] ^ (nb[:, 1] << 32))
        h = mix64(h ^ nb[:, 2])
        acc = (acc + int(h.sum().item())) & M64
    return acc


def triangle_digest(tcell, T, key_of):
    """sum over triangles of hash(cell, three vertex KEYS in winding order), mod 2^64.
    key_of(ids int64 [m]) -> canonical keys int64 [m] of those global vertex ids"""
    n = tcell.shape[0]
    acc = 0
    for a in range(0, n, 1 << 24):
        b = min(n, a + (1 << 24))
        ids = _bits32(T[a:b]).view(-1, 3)
        h = mix64(tcell[a:b])
        for j in range(3):
            h = mix64(h ^ key_of(ids[:, j]))
        acc = (acc + int(h.sum().item())) & M64
    return acc
