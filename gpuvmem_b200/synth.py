"""Synthetic ALMA-like uv-coverage generator (replaces MSFITSIO/casacore ingestion).

Fills the same in-memory layout ``readMS`` produces (reference
``src/MSFITSIO.cu:398-754``, ``include/MSFITSIO.cuh:82-121``): per (field,
channel, stokes) block, ``uvw`` as ``[Z][3]`` float64 **metres**, ``Vo`` as
``[Z][2]`` float32, ``weight`` as ``[Z]`` float32, channel frequencies stored as
float32, one field, one correlation (XX), plus the FITS-header values
(``headerValues``, ``include/MSFITSIO.cuh:140-150``) that ``MFS::configure`` reads.

Everything is deterministic in ``seed``. Host-side numpy only; no GPU needed.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

LIGHTSPEED = np.float32(2.99792458e8)  # include/MSFITSIO.cuh:54


@dataclass
class Problem:
    """One dataset: header + antenna model + per-channel visibility blocks."""
    M: int
    N: int
    DELTAX: float           # deg, negative (RA)
    DELTAY: float           # deg
    ra: float               # deg
    dec: float              # deg
    crpix1: float
    crpix2: float
    telescope: str
    antenna_diameter: float
    freqs: np.ndarray       # float32 [nchan]
    uvw: List[np.ndarray] = field(default_factory=list)   # [Z][3] float64 metres
    Vo: List[np.ndarray] = field(default_factory=list)    # [Z][2] float32
    w: List[np.ndarray] = field(default_factory=list)     # [Z] float32
    name: str = "synthetic"
    sources: list = field(default_factory=list)

    @property
    def nchan(self):
        return len(self.freqs)

    def total_vis(self):
        return int(sum(len(x) for x in self.w))

    def subset(self, nvis):
        """First ``nvis`` samples of every channel (bounded samples for CPU baselines)."""
        p = Problem(self.M, self.N, self.DELTAX, self.DELTAY, self.ra, self.dec, self.crpix1,
                    self.crpix2, self.telescope, self.antenna_diameter, self.freqs.copy(),
                    name=self.name + f"[:{nvis}]", sources=self.sources)
        for c in range(self.nchan):
            p.uvw.append(np.ascontiguousarray(self.uvw[c][:nvis]))
            p.Vo.append(np.ascontiguousarray(self.Vo[c][:nvis]))
            p.w.append(np.ascontiguousarray(self.w[c][:nvis]))
        return p


def _tracks(rng, nant, ntimes, dec_rad, bmin, bmax, ha_range=(-2.0, 2.0)):
    """Earth-rotation (u,v,w) tracks in metres, time-major like a Measurement Set."""
    r = np.exp(rng.uniform(np.log(bmin / 2.0), np.log(bmax / 2.0), nant))
    th = rng.uniform(0.0, 2.0 * np.pi, nant)
    pos = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(0.0, 0.02 * bmax, nant)], axis=1)
    ia, ib = np.triu_indices(nant, k=1)
    B = pos[ia] - pos[ib]                                   # equatorial-frame baselines (X, Y, Z)
    H = np.deg2rad(15.0 * np.linspace(ha_range[0], ha_range[1], ntimes))
    sH, cH = np.sin(H)[:, None], np.cos(H)[:, None]
    sd, cd = np.sin(dec_rad), np.cos(dec_rad)
    X, Y, Zc = B[None, :, 0], B[None, :, 1], B[None, :, 2]
    u = sH * X + cH * Y
    v = -sd * cH * X + sd * sH * Y + cd * Zc
    w = cd * cH * X - cd * sH * Y + sd * Zc
    return np.stack([u.ravel(), v.ravel(), w.ravel()], axis=1)


def _sky_vis(u, v, sources):
    """Exact visibilities of point/Gaussian components with the reference's sign
    convention V(u,v) = sum F exp(+2 pi i (u x + v y))."""
    out = np.zeros(len(u), dtype=np.complex128)
    for (flux, x, y, sigma) in sources:
        env = np.exp(-2.0 * np.pi ** 2 * sigma ** 2 * (u * u + v * v)) if sigma > 0 else 1.0
        out += flux * env * np.exp(2j * np.pi * (u * x + v * y))
    return out


def make_problem(N=512, nvis=1 << 20, nchan=1, freq0=2.3e11, bandwidth=0.0, seed=20261017,
                 nant=40, bmin=15.0, bmax=1000.0, dec_deg=-30.0, ra_deg=150.0, telescope="ALMA",
                 antenna_diameter=12.0, wterm=True, nsrc=6, weight_scale=1.0e3, grid_fill=0.9,
                 name=None) -> Problem:
    """ALMA-like problem: ``nvis`` samples per channel on a N x N image.

    The pixel size is chosen so that the longest projected baseline at the highest
    frequency reaches ``grid_fill`` of the grid half-width (otherwise
    ``vis_mod`` zeroes weights).
    """
    rng = np.random.default_rng(seed)
    nbl = nant * (nant - 1) // 2
    ntimes = -(-nvis // nbl)
    uvw_m = _tracks(rng, nant, ntimes, np.deg2rad(dec_deg), bmin, bmax)[:nvis]
    if not wterm:
        uvw_m[:, 2] = 0.0
    if nchan > 1:
        freqs = np.linspace(freq0 - bandwidth / 2, freq0 + bandwidth / 2, nchan).astype(np.float32)
    else:
        freqs = np.array([freq0], dtype=np.float32)
    lam_min = float(LIGHTSPEED / freqs.max())
    uvmax = np.abs(uvw_m[:, :2]).max() / lam_min
    deltau = uvmax / (grid_fill * (N / 2 - 2))
    dx_rad = 1.0 / (N * deltau)
    cdelt = np.rad2deg(dx_rad)

    # sky: a few components well inside the primary beam / field of view
    fov = N * dx_rad
    sources = []
    for s in range(nsrc):
        flux = float(rng.uniform(0.02, 0.2))
        x, y = (rng.uniform(-0.2, 0.2, 2) * fov).tolist()
        sigma = float(rng.choice([0.0, 1.0, 2.5]) * 4 * dx_rad)
        sources.append((flux, x, y, sigma))

    p = Problem(M=N, N=N, DELTAX=-cdelt, DELTAY=cdelt, ra=ra_deg, dec=dec_deg,
                crpix1=N / 2 + 1, crpix2=N / 2 + 1, telescope=telescope,
                antenna_diameter=antenna_diameter, freqs=freqs,
                name=name or f"alma-like-{N}px-{nvis}vis-{nchan}ch", sources=sources)
    for c in range(nchan):
        lam = float(LIGHTSPEED / freqs[c])
        u, v = uvw_m[:, 0] / lam, uvw_m[:, 1] / lam
        # weights ~ log-normal around 1/sigma^2
        w = (weight_scale * np.exp(rng.normal(0.0, 0.5, nvis))).astype(np.float32)
        vis = _sky_vis(u, v, sources)
        sig = 1.0 / np.sqrt(w.astype(np.float64))
        vis = vis + sig * (rng.normal(size=nvis) + 1j * rng.normal(size=nvis))
        Vo = np.stack([vis.real, vis.imag], axis=1).astype(np.float32)
        p.uvw.append(np.ascontiguousarray(uvw_m.copy()))
        p.Vo.append(np.ascontiguousarray(Vo))
        p.w.append(w)
    return p


def write_gvms(p: Problem, path, fields=None, corr_types=(9,)):
    """Write the GVMS container the C++ host layer reads (``csrc/host/msdata.hpp``): the values
    ``readMS`` + ``readFITSHeader`` would deliver. Default: one field at the image centre, one
    correlation (XX = 9). ``fields``: a list of Problems sharing ``p``'s header and frequencies, one per
    mosaic field, each optionally carrying ``field_centre = (ra, dec)`` in degrees. ``corr_types``:
    correlation codes (include/functions.cuh:25-59); every correlation after the first gets the
    first one's samples with a different amplitude — only LL/RR/XX/YY may be used by the engine."""
    import struct
    fields = [p] if fields is None else list(fields)
    with open(path, "wb") as f:
        f.write(b"GVMS0001")
        f.write(struct.pack("<qq", p.M, p.N))
        f.write(struct.pack("<6d", p.DELTAX, p.DELTAY, p.ra, p.dec, p.crpix1, p.crpix2))
        f.write(struct.pack("<ff", -1.0, p.antenna_diameter))
        f.write(p.telescope.encode()[:31].ljust(32, b"\0"))
        f.write(struct.pack("<iii", len(fields), p.nchan, len(corr_types)))
        f.write(struct.pack(f"<{len(corr_types)}i", *corr_types))
        for q in fields:
            fc = getattr(q, "field_centre", None) or (p.ra, p.dec)
            ra, dec = np.deg2rad(fc[0]), np.deg2rad(fc[1])
            f.write(struct.pack("<4d", ra, dec, ra, dec))
            f.write(np.asarray(p.freqs, dtype="<f4").tobytes())
            for c in range(p.nchan):
                for s in range(len(corr_types)):
                    f.write(struct.pack("<q", len(q.w[c])))
                    f.write(np.ascontiguousarray(q.uvw[c], dtype="<f8").tobytes())
                    f.write(np.ascontiguousarray(q.Vo[c] * (1.0 if s == 0 else -3.0 * s), dtype="<f4").tobytes())
                    f.write(np.ascontiguousarray(q.w[c], dtype="<f4").tobytes())


# BASELINE.json configs. Sizes can be scaled down for tests.
def config_c1(scale=1.0, **kw):
    return make_problem(N=512, nvis=int((1 << 20) * scale), nchan=1, freq0=6.9147e11,
                        name="C1-co65-shaped", **kw)


def config_c2(scale=1.0, **kw):
    # Baselines out to ~14 km (ALMA's most extended configurations): the 2048-pixel field is then
    # about 1.5x the 12 m primary-beam FWHM, so the default noise mask (noise < 10 min) leaves
    # nearly the whole image unmasked and "Mpix" means computed pixels. The reference's DChi2 and
    # this engine both skip masked pixels; a compact array (1 km) would leave ~2 % of them.
    kw.setdefault("bmax", 14000.0)
    kw.setdefault("bmin", 150.0)
    return make_problem(N=2048, nvis=int(10_000_000 * scale), nchan=1, freq0=2.3e11,
                        name="C2-alma-2048-10M", **kw)


def config_c3(scale=1.0, nchan=64, N=2048, **kw):
    # same reasoning as C2: at 100 GHz the 12 m primary beam is 58" FWHM; baselines out to ~14 km make
    # the 2048-pixel field ~0.7 FWHM, so the noise mask keeps (nearly) every pixel in the computation
    kw.setdefault("bmax", 14000.0)
    kw.setdefault("bmin", 150.0)
    return make_problem(N=N, nvis=int(1_000_000 * scale), nchan=nchan, freq0=1.0e11,
                        bandwidth=2.0e9, name=f"C3-mfs-{nchan}ch", **kw)


def config_c4(scale=1.0, **kw):
    # VLBI-like: 8 stations, very long baselines, default telescope -> Gaussian beam
    return make_problem(N=4096, nvis=int(50_000_000 * scale), nchan=1, freq0=2.3e11, nant=8,
                        bmin=1.0e6, bmax=1.0e7, telescope="EHT", antenna_diameter=12.0,
                        name="C4-m87-style-vlbi", **kw)


def config_c5(scale=1.0, **kw):
    # gridded-visibility mode: 8192 x 8192 uv grid, 200 M raw visibilities (Briggs R = 0 weighting and the
    # convolutional gridding happen inside MFS::configure, -g); extended array as in C2 so the field is unmasked
    kw.setdefault("bmax", 14000.0)
    kw.setdefault("bmin", 150.0)
    kw.setdefault("nant", 64)
    return make_problem(N=8192, nvis=int(200_000_000 * scale), nchan=1, freq0=2.3e11,
                        name="C5-gridded-8192-200M", **kw)
