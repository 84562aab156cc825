"""The host layer's dependency-free FITS reader/writer (csrc/host/fits.cpp; it stands in for what the
reference does through cfitsio, src/MSFITSIO.cu:65-360) cross-checked against an independent numpy
implementation (gpuvmem_b200/fits.py): both directions, every supported BITPIX, header semantics of
readFITSHeader and OCopyFITS."""
import numpy as np
import pytest

from gpuvmem_b200 import fits, host

HDR = {"BUNIT": "JY/BEAM", "CTYPE1": "RA---SIN", "CRVAL1": 201.365063, "CDELT1": -2.7777777777e-06, "CRPIX1": 65.0,
       "CUNIT1": "deg", "CTYPE2": "DEC--SIN", "CRVAL2": -43.019113, "CDELT2": 2.7777777777e-06, "CRPIX2": 65.0,
       "BMAJ": 1.2e-4, "BMIN": 9.0e-5, "BPA": 33.5, "RADESYS": "FK5", "EQUINOX": 2000.0, "TELESCOP": "ALMA",
       "NAXIS3": 1, "CTYPE3": "FREQ", "CRVAL3": 2.3e11, "CDELT3": 2.0e9, "CRPIX3": 1.0}


def test_header_values_as_readFITSHeader(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.standard_normal((96, 128)).astype(np.float32)      # NAXIS2 = 96 rows, NAXIS1 = 128
    p = str(tmp_path / "model.fits")
    fits.write_fits(p, img, HDR)
    assert host.load_host_library() is not None
    h, data = host.fits_read(p)
    assert (h["naxis1"], h["naxis2"], h["bitpix"], h["has_wcs"]) == (128, 96, -32, 1)
    assert h["cdelt1"] == HDR["CDELT1"] and h["cdelt2"] == HDR["CDELT2"]
    assert h["crval1"] == HDR["CRVAL1"] and h["crval2"] == HDR["CRVAL2"]
    assert h["crpix1"] == 65.0 and h["crpix2"] == 65.0
    assert (h["bmaj"], h["bmin"], h["bpa"]) == (HDR["BMAJ"], HDR["BMIN"], HDR["BPA"])
    assert h["noise"] == -1.0, "NOISE absent -> -1 (src/MSFITSIO.cu:289-299)"
    assert h["equinox"] == 2000.0
    assert np.array_equal(data, img)
    fits.write_fits(p, img, dict(HDR, NOISE=3.5e-4))
    assert abs(host.fits_read(p, want_data=False)[0]["noise"] - 3.5e-4) < 1e-10


@pytest.mark.parametrize("bitpix", [8, 16, 32, -32, -64])
def test_reads_every_bitpix_with_scaling(tmp_path, bitpix):
    rng = np.random.default_rng(bitpix + 100)
    n2, n1 = 7, 13
    raw = {8: rng.integers(0, 255, (n2, n1)), 16: rng.integers(-30000, 30000, (n2, n1)),
           32: rng.integers(-2 ** 30, 2 ** 30, (n2, n1)), -32: rng.standard_normal((n2, n1)),
           -64: rng.standard_normal((n2, n1))}[bitpix]
    dt = {8: ">u1", 16: ">i2", 32: ">i4", -32: ">f4", -64: ">f8"}[bitpix]
    cards = ["SIMPLE  =                    T", f"BITPIX  = {bitpix:>20d}", "NAXIS   =                    2",
             f"NAXIS1  = {n1:>20d}", f"NAXIS2  = {n2:>20d}", "BSCALE  =                  0.5", "BZERO   =                 10.0",
             "COMMENT  free text = not a value", "END"]
    head = "".join(c.ljust(80) for c in cards).encode()
    head += b" " * (-len(head) % 2880)
    body = raw.astype(dt).tobytes()
    body += b"\0" * (-len(body) % 2880)
    p = str(tmp_path / f"b{bitpix}.fits")
    open(p, "wb").write(head + body)
    h, data = host.fits_read(p)
    assert h["bitpix"] == bitpix and h["has_wcs"] == 0
    want = (0.5 * raw.astype(dt).astype(np.float64) + 10.0).astype(np.float32)
    assert np.array_equal(data, want)
    _, ref = fits.read_fits(p)
    assert np.allclose(data, ref, rtol=1e-7)


def test_writer_copies_the_template_header_like_OCopyFITS(tmp_path):
    rng = np.random.default_rng(5)
    tmpl = str(tmp_path / "model.fits")
    fits.write_fits(tmpl, np.zeros((64, 64), np.float32), HDR)
    img = rng.standard_normal((64, 64)).astype(np.float32)
    out = str(tmp_path / "out.fits")
    host.fits_write(out, img, template=tmpl, bunit="JY/PIXEL", niter=42, radesys="ICRS", equinox=2000.0,
                    crval1=10.5, crval2=-20.25)
    size = len(open(out, "rb").read())
    assert size % 2880 == 0
    h, data = fits.read_fits(out)                       # the independent reader
    assert np.array_equal(data.astype(np.float32), img)
    assert h["SIMPLE"] is True and h["BITPIX"] == -32 and h["NAXIS"] == 2 and h["NAXIS1"] == 64 and h["NAXIS2"] == 64
    assert h["BUNIT"] == "JY/PIXEL" and h["NITER"] == 42 and h["RADESYS"] == "ICRS"
    assert h["CRVAL1"] == 10.5 and h["CRVAL2"] == -20.25, "replaced"
    assert h["CDELT1"] == HDR["CDELT1"] and h["CRPIX2"] == 65.0 and h["CTYPE1"] == "RA---SIN" and h["TELESCOP"] == "ALMA", "copied"
    assert "NAXIS3" not in h and "CTYPE3" not in h and "CRVAL3" not in h, "degenerate axes of the template are dropped"
    # and our own reader reads it back
    h2, d2 = host.fits_read(out)
    assert np.array_equal(d2, img) and h2["crval1"] == 10.5 and h2["cdelt1"] == HDR["CDELT1"]
    # without a template
    host.fits_write(out, img[:10, :20], bunit="", niter=0)
    h, data = fits.read_fits(out)
    assert data.shape == (10, 20) and np.array_equal(data.astype(np.float32), img[:10, :20])


def test_rejects_what_is_not_an_image(tmp_path):
    p = str(tmp_path / "x.fits")
    open(p, "wb").write(b"GVMS0001" + b"\0" * 3000)
    with pytest.raises(RuntimeError):
        host.fits_read(p)
    cards = ["SIMPLE  =                    T", "BITPIX  =                    8", "NAXIS   =                    0", "END"]
    head = "".join(c.ljust(80) for c in cards).encode()
    open(p, "wb").write(head + b" " * (-len(head) % 2880))
    with pytest.raises(RuntimeError):
        host.fits_read(p)
