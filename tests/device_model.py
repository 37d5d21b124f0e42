"""numpy model of the arithmetic the tiled CUDA kernel performs (clownresampler_b200/csrc/crb_device.cu),
driven by the plan the C host code builds.  It exists so that the plan construction and the kernel's
integer identities (closed-form position, phase row formula, sign-split chains, one-instruction
truncating MAC, one-instruction normaliser) can be checked against the oracle on a machine without
a GPU.  It is test code, not a fallback: nothing in the package imports it."""
import numpy as np


def _mac_trunc(acc, S, k, bias):
    """hi32(S * k + ((acc << 32) | bias)) with int64 wraparound-free operands (all values fit)."""
    total = S.astype(object) * int(k) if np.isscalar(k) else S.astype(object) * k.astype(object)
    total = total + (acc.astype(object) * (1 << 32)) + bias.astype(object)
    return np.array([int(v) >> 32 for v in total], dtype=np.int64)


def resample(geo, rows, padded, q0, first_out, n_out, fmt=0):
    """frames [first_out, first_out + n_out) for a job whose frame 0 sits at 16.16 position q0 - delta."""
    ch = geo["channels"]
    padded = np.asarray(padded, dtype=np.int64).reshape(-1, ch)
    n = np.arange(first_out, first_out + n_out, dtype=object)
    q = n * geo["increment"] + int(q0)
    ws = np.array([(int(v) + 65535) >> 16 for v in q], dtype=np.int64)
    e = np.array([(int(w) << 16) - int(v) for w, v in zip(ws, q)], dtype=np.int64)
    assert e.min(initial=0) >= 0 and e.max(initial=0) <= 65535
    row = ((e + geo["delta"]) * geo["step"] >> 16) - geo["ks0"]
    for b in geo["breaks"]:
        row = row + (e >= b)
    out = np.zeros((n_out, ch + (fmt == 2)), dtype=np.int64)
    need = ws.max(initial=0) + geo["taps_max"] + 1
    if padded.shape[0] < need:   # columns past the end are zero-weight; give them something to read
        padded = np.vstack([padded, np.full((need - padded.shape[0], ch), 12345, dtype=np.int64)])
    for c in range(ch):
        accp = np.zeros(n_out, dtype=np.int64)
        accn = np.zeros(n_out, dtype=np.int64)
        for (col, length, off, neg) in geo["runs"]:
            for i in range(length):
                k = rows[row, col + i].astype(np.int64)
                s = padded[ws + off + i, c]
                S = s << 16
                bias = np.where(s < 0, 0xFFFFFFFF, 0).astype(np.int64)
                if neg:
                    accn = _mac_trunc(accn, S, k, bias)
                else:
                    accp = _mac_trunc(accp, S, k, bias)
        acc = accp - accn
        recip_row = rows[row, geo["n_cols"]].astype(np.int64)
        if fmt == 2:
            out[:, c] = acc
            out[:, ch] = recip_row >> geo["recip_shift"]
        elif geo["recip_shift"] == 15:
            out[:, c] = _mac_trunc(np.zeros(n_out, dtype=np.int64), acc << 2, recip_row, np.where(acc < 0, 0xFFFFFFFF, 0).astype(np.int64))
        else:
            qq = acc.astype(object) * recip_row.astype(object)
            out[:, c] = np.array([(int(v) + ((int(v) >> 63) & 32767)) >> 15 for v in qq], dtype=np.int64)
    if fmt == 1:
        out = np.clip(out, -0x7FFF, 0x7FFF)
    return out
