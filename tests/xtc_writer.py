"""XTC writer for the round-trip test of the library's XTC reader (TEST INFRASTRUCTURE; the product only reads).

A transcription of the published xdr3dcoord compressor (GROMACS libxdrf / xdrfile, the codec MDAnalysis wraps): integer
coordinates at the file's precision, every atom either a "large" mixed-radix triple or, inside a run of up to eight
atoms, a "small" triple relative to its predecessor, adaptive small range, first two atoms of a run swapped."""
import struct

import numpy as np

MAGICINTS = [0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
             1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
             65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122,
             1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216]
FIRSTIDX = 9
LASTIDX = len(MAGICINTS)


class BitWriter(object):
    def __init__(self):
        self.acc = 0; self.nbits = 0

    def bits(self, nb, value):
        self.acc = (self.acc << nb) | (int(value) & ((1 << nb) - 1)); self.nbits += nb

    def ints(self, nb, sizes, nums):
        """three integers as one mixed-radix number, emitted low byte first (sendints)."""
        v = (int(nums[0]) * sizes[1] + int(nums[1])) * sizes[2] + int(nums[2])
        full, rest = divmod(nb, 8)
        if rest == 0 and full:
            full -= 1; rest = 8
        for _ in range(full):
            self.bits(8, v & 0xff); v >>= 8
        if rest:
            self.bits(rest, v)

    def bytes(self):
        pad = (-self.nbits) % 8
        n = (self.nbits + pad) // 8
        return (self.acc << pad).to_bytes(n, 'big') if n else b''


def sizeofint(size):
    num, nb = 1, 0
    while size >= num and nb < 32:
        nb += 1; num <<= 1
    return nb


def sizeofints(sizes):
    return max(1, (int(sizes[0]) * int(sizes[1]) * int(sizes[2]) - 1).bit_length()) if False else _sizeofints(sizes)


def _sizeofints(sizes):
    byts = [1]
    for s in sizes:
        tmp = 0
        for k in range(len(byts)):
            tmp = byts[k] * int(s) + tmp
            byts[k] = tmp & 0xff; tmp >>= 8
        while tmp:
            byts.append(tmp & 0xff); tmp >>= 8
    num, nb = 1, 0
    while byts[-1] >= num:
        nb += 1; num *= 2
    return nb + (len(byts) - 1) * 8


def compress_frame(xyz, precision=1000.0):
    """-> bytes of the coordinate block of one frame (after the header): natoms, precision, ..."""
    x = np.asarray(xyz, np.float64)
    n = len(x)
    out = struct.pack('>i', n)
    if n <= 9:
        return out + b''.join(struct.pack('>f', float(v)) for v in x.ravel())
    ip = np.where(x >= 0, np.floor(x * precision + 0.5), np.ceil(x * precision - 0.5)).astype(np.int64)
    minint = ip.min(0); maxint = ip.max(0)
    sizeint = (maxint - minint + 1).astype(np.int64)
    if (sizeint > 0xffffff).any():
        bitsizeint = [sizeofint(int(s)) for s in sizeint]; bitsize = 0
    else:
        bitsize = _sizeofints(sizeint)
    d = np.abs(np.diff(ip, axis=0)).sum(1)
    mindiff = int(d.min()) if len(d) else 0
    smallidx = FIRSTIDX
    while smallidx < LASTIDX - 1 and MAGICINTS[smallidx] < mindiff:
        smallidx += 1
    maxidx = min(LASTIDX - 1, smallidx + 8); minidx = maxidx - 8
    smaller = MAGICINTS[max(FIRSTIDX, smallidx - 1)] // 2
    smallnum = MAGICINTS[smallidx] // 2
    sizesmall = [MAGICINTS[smallidx]] * 3
    larger = MAGICINTS[maxidx] // 2
    head = struct.pack('>f', float(precision)) + struct.pack('>3i', *[int(v) for v in minint]) + struct.pack('>3i', *[int(v) for v in maxint]) \
        + struct.pack('>i', smallidx)
    bw = BitWriter()
    c = [list(map(int, r)) for r in ip]
    i = 0; prevrun = -1; prev = [0, 0, 0]
    while i < n:
        is_small = False
        this = c[i]
        if smallidx < maxidx and i >= 1 and all(abs(this[k] - prev[k]) < larger for k in range(3)):
            is_smaller = 1
        elif smallidx > minidx:
            is_smaller = -1
        else:
            is_smaller = 0
        if i + 1 < n and all(abs(this[k] - c[i + 1][k]) < smallnum for k in range(3)):
            c[i], c[i + 1] = c[i + 1], c[i]; this = c[i]; is_small = True
        tmp = [this[k] - int(minint[k]) for k in range(3)]
        if bitsize == 0:
            for k in range(3):
                bw.bits(bitsizeint[k], tmp[k])
        else:
            bw.ints(bitsize, [int(s) for s in sizeint], tmp)
        prev = list(this); i += 1
        run = []
        if not is_small and is_smaller == -1:
            is_smaller = 0
        while is_small and len(run) < 8 * 3:
            this = c[i]
            if is_smaller == -1 and sum((this[k] - prev[k]) ** 2 for k in range(3)) >= smaller * smaller:
                is_smaller = 0
            run += [this[k] - prev[k] + smallnum for k in range(3)]
            prev = list(this); i += 1
            is_small = i < n and all(abs(c[i][k] - prev[k]) < smallnum for k in range(3))
        if len(run) != prevrun or is_smaller != 0:
            prevrun = len(run)
            bw.bits(1, 1); bw.bits(5, len(run) + is_smaller + 1)
        else:
            bw.bits(1, 0)
        for k in range(0, len(run), 3):
            bw.ints(smallidx, sizesmall, run[k:k + 3])
        if is_smaller != 0:
            smallidx += is_smaller
            if is_smaller < 0:
                smallnum = smaller; smaller = MAGICINTS[smallidx - 1] // 2
            else:
                smaller = smallnum; smallnum = MAGICINTS[smallidx] // 2
            sizesmall = [MAGICINTS[smallidx]] * 3
    data = bw.bytes()
    pad = (-len(data)) % 4
    return out + head + struct.pack('>i', len(data)) + data + b'\0' * pad


def write_xtc(path, frames, precision=1000.0, dt=1.0, box=None):
    """frames [nframes, natoms, 3] in nm."""
    box = np.eye(3) * 10.0 if box is None else np.asarray(box, float)
    with open(path, 'wb') as fh:
        for k, x in enumerate(frames):
            fh.write(struct.pack('>3i', 1995, len(x), k) + struct.pack('>f', k * dt) + struct.pack('>9f', *box.ravel()))
            fh.write(compress_frame(x, precision))
