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


def unpack_rows(geo, rows):
    """Undo the 16-byte packing of the unstretched kernel's rows: returns generic rows
    [K0, K1, k2, k3, K4, kr] (small columns as |k| << 16, big columns as |k|, kr = (recip - 32768) << 17)."""
    if not geo["unstretched5"]:
        return rows
    r = rows.astype(np.int64) & 0xFFFFFFFF
    out = np.zeros((rows.shape[0], 8), dtype=np.int64)
    out[:, 0] = (r[:, 2] << 16) & 0xFFFFFFFF
    out[:, 1] = r[:, 2] & 0xFFFF0000
    out[:, 2] = r[:, 0]
    out[:, 3] = r[:, 1]
    out[:, 4] = r[:, 3] & 0xFFFF0000
    kr = (r[:, 3] << 16) & 0xFFFFFFFF
    out[:, 5] = np.where(kr >= 1 << 31, kr - (1 << 32), kr)
    return out


def _chain_accumulators(geo, rows, row, padded, ws, c):
    """Chain form of the general kernel (crb_kernels.cuh frame_chains): plain weights, 32-bit products.  A chain keeps its running
    sum in the upper 16 bits of a 32-bit register (mac_hi16), single columns are folded one by one (mac_t16).  Every intermediate
    is asserted to fit int32: the model proves the plan's range claims on the data it is fed."""
    fb = 2 * geo["channels"]
    n = ws.shape[0]
    accp = np.zeros(n, dtype=np.int64)
    accn = np.zeros(n, dtype=np.int64)
    for gi in range(geo["n_groups"]):
        first, count, _ = geo["groups"][gi]
        kind = geo["group_kinds"][gi]
        cls, single = kind >> 1, kind & 1
        chain = np.zeros(n, dtype=np.int64)
        folded = np.zeros(n, dtype=np.int64)
        for col in range(first, first + count):
            k = rows[row, col].astype(np.int64)
            assert geo["col_offsets"][col] % fb == 0
            m = padded[ws + geo["col_offsets"][col] // fb, c]
            if cls == 2:
                negative = np.where(k < 0, ~m, m) < 0       # sign bit of sample ^ (k >> 31)
            else:
                assert (k >= 0).all()
                negative = m < 0
            bias = np.where(negative, 0xFFFF, 0)
            if single:
                t = m * k + bias
                assert (np.abs(t) < 1 << 31).all() or ((t >= -(1 << 31)) & (t < 1 << 31)).all()
                folded += t >> 16
            else:
                chain = m * k + ((chain >> 16) << 16) + bias
                assert ((chain >= -(1 << 31)) & (chain < 1 << 31)).all()
        total = folded if single else chain >> 16
        if cls == 1:
            accn += total
        else:
            accp += total
    return accp, accn


def resample(geo, rows, padded, q0, first_out, n_out, fmt=0):
    """frames [first_out, first_out + n_out) for a job whose frame 0 sits at 16.16 position q0 - delta."""
    rows = unpack_rows(geo, rows)
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
        if geo.get("chain_mode"):
            accp, accn = _chain_accumulators(geo, rows, row, padded, ws, c)
        for (col, length, off, neg, big) in ([] if geo.get("chain_mode") else geo["runs"]):
            for i in range(length):
                k = rows[row, col + i].astype(np.int64)
                s = padded[ws + off + i, c]
                a = (s << 16) if big else s          # multiplicand: sample << 16 with |k|, or sample with |k| << 16
                bias = s & 0xFFFFFFFF                # the sign-extended sample as an unsigned 32-bit word
                if neg == 2:
                    # signed column: the bias is the sample with all bits flipped where the weight is negative
                    k = np.where(k >= 1 << 31, k - (1 << 32), k)
                    accp = _mac_trunc(accp, a, k, np.where(k < 0, ~s, s) & 0xFFFFFFFF)
                elif neg:
                    accn = _mac_trunc(accn, a, k, bias)
                else:
                    accp = _mac_trunc(accp, a, k, bias)
        acc = accp - accn
        word = rows[row, geo["n_cols"]].astype(np.int64)
        mode = geo["norm_mode"]
        recip = (word >> 17) + 32768 if mode >= 2 else (word >> 16) + 32768 if mode == 1 else word
        if fmt == 2:
            out[:, c] = acc
            out[:, ch] = recip
        elif mode == 3:
            out[:, c] = _mac_trunc(acc, acc, word, acc & 0xFFFFFFFF)
        elif mode == 2:
            out[:, c] = _mac_trunc(acc, acc, word, (acc >> 31) & 0xFFFFFFFF)
        elif mode == 1:
            out[:, c] = _mac_trunc(acc, acc << 1, word, (acc >> 31) & 0xFFFFFFFF)
        else:
            qq = acc.astype(object) * recip.astype(object)
            out[:, c] = np.array([(int(v) + ((int(v) >> 63) & 32767)) >> 15 for v in qq], dtype=np.int64)
    if fmt == 1:
        out = np.clip(out, -0x7FFF, 0x7FFF)
    return out
