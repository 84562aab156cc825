"""A minimal FITS image reader/writer in numpy (primary HDU, 2-D, BITPIX -32/-64/16/32) — an
implementation independent of gpuvmem_b200/csrc/host/fits.cpp, used by the tests to cross-check
it and by callers to prepare the ``-m`` model header / ``-U`` mask the command line takes."""
import numpy as np

_BLOCK, _CARD = 2880, 80
_DTYPES = {8: ">u1", 16: ">i2", 32: ">i4", -32: ">f4", -64: ">f8"}


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "T" if value else "F"
        body = f"{key:<8}= {v:>20}"
    elif isinstance(value, str):
        body = f"{key:<8}= " + f"'{value:<8}'".ljust(20)
    elif isinstance(value, (int, np.integer)):
        body = f"{key:<8}= {int(value):>20d}"
    else:
        txt = f"{float(value):.17G}"
        if not any(ch in txt for ch in ".EN"):
            txt += "."
        body = f"{key:<8}= {txt:>20}"
    if comment:
        body += " / " + comment
    return body[:_CARD].ljust(_CARD)


def write_fits(path, data, header=None):
    """``data``: 2-D array [NAXIS2][NAXIS1] (stored as BITPIX -32); ``header``: dict of extra cards."""
    data = np.asarray(data, dtype=np.float32)
    assert data.ndim == 2
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", -32), _card("NAXIS", 2),
             _card("NAXIS1", data.shape[1]), _card("NAXIS2", data.shape[0])]
    for k, v in (header or {}).items():
        cards.append(_card(k, v))
    cards.append("END".ljust(_CARD))
    head = "".join(cards).encode("ascii")
    head += b" " * (-len(head) % _BLOCK)
    body = data.astype(">f4").tobytes()
    body += b"\0" * (-len(body) % _BLOCK)
    with open(path, "wb") as f:
        f.write(head + body)


def read_fits(path):
    """Returns (header dict, data [NAXIS2][NAXIS1] float64 with BSCALE/BZERO applied)."""
    raw = open(path, "rb").read()
    header, pos, done = {}, 0, False
    while not done:
        block = raw[pos:pos + _BLOCK].decode("ascii")
        assert len(block) == _BLOCK, "truncated FITS header"
        pos += _BLOCK
        for i in range(0, _BLOCK, _CARD):
            card = block[i:i + _CARD]
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:]
            if val.lstrip().startswith("'"):
                v = val.lstrip()[1:]
                v = v[:v.index("'")].rstrip()
            else:
                v = val.split("/")[0].strip()
                if v in ("T", "F"):
                    v = v == "T"
                else:
                    try:
                        v = int(v)
                    except ValueError:
                        v = float(v.replace("D", "E"))
            header[key] = v
    n1, n2, bitpix = header["NAXIS1"], header["NAXIS2"], header["BITPIX"]
    dt = np.dtype(_DTYPES[bitpix])
    data = np.frombuffer(raw, dtype=dt, count=n1 * n2, offset=pos).reshape(n2, n1).astype(np.float64)
    data = data * header.get("BSCALE", 1.0) + header.get("BZERO", 0.0)
    assert (len(raw) - pos) % _BLOCK == 0, "FITS data not padded to 2880-byte blocks"
    return header, data
