"""TEST INFRASTRUCTURE ONLY: a small Shorten ENCODER (format versions 1-3, 16-bit signed PCM), written independently from the
published format description, used to produce streams for the decoder in csrc/ld_shorten.cpp (no shorten binary or ICSI
file is available offline).  Block commands: DIFF0-3, QLPC, ZERO, BITSHIFT, BLOCKSIZE, VERBATIM, QUIT; Rice-coded residuals."""
import numpy as np

FN_DIFF0, FN_DIFF1, FN_DIFF2, FN_DIFF3, FN_QUIT, FN_BLOCKSIZE, FN_BITSHIFT, FN_QLPC, FN_ZERO, FN_VERBATIM = range(10)
TYPE_S16HL, TYPE_S16LH = 3, 5
LPCQUANT = 5


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value, n):
        for i in range(n - 1, -1, -1):
            self.bits.append((value >> i) & 1)

    def uvar(self, val, nbin):
        assert val >= 0
        self.bits += [0] * (val >> nbin) + [1]
        self.put(val & ((1 << nbin) - 1), nbin)

    def var(self, val, nbin):
        self.uvar(((~val) << 1) | 1 if val < 0 else val << 1, nbin + 1)

    def ulong(self, val):
        nbit = int(val).bit_length()
        self.uvar(nbit, 2)
        self.uvar(val, nbit)

    def bytes(self):
        bits = self.bits + [0] * (-len(self.bits) % 32)
        return np.packbits(np.array(bits, dtype=np.uint8)).tobytes()


def encode(channels, version=2, blocksize=256, nmean=4, maxnlpc=2, plan=None, ftype=TYPE_S16LH, verbatim=b""):
    """channels: list of equal-length int sequences.  plan(block_index, chan) -> (cmd, bitshift, qlpc coefficients) chooses
    the command of every block (default DIFF1, no bit shift)."""
    nchan, n = len(channels), len(channels[0])
    w = BitWriter()
    w.ulong(ftype); w.ulong(nchan); w.ulong(blocksize); w.ulong(maxnlpc); w.ulong(nmean); w.ulong(0)
    if verbatim:
        w.uvar(FN_VERBATIM, 2); w.uvar(len(verbatim), 5)
        for b in verbatim:
            w.uvar(b, 8)
    nwrap = max(3, maxnlpc)
    hist = [[0] * nwrap for _ in range(nchan)]            # last nwrap samples (stored, i.e. bit-shifted, domain)
    offs = [[0] * max(1, nmean) for _ in range(nchan)]
    bitshift, cur_bs = 0, blocksize
    lpcqoffset = (1 << (LPCQUANT - 1)) if version > 1 else 0
    pos, blk = 0, 0
    while pos < n:
        bs = min(cur_bs, n - pos)
        if bs != cur_bs:
            w.uvar(FN_BLOCKSIZE, 2); w.ulong(bs)
            cur_bs = bs
        for c in range(nchan):
            cmd, shift, q = plan(blk, c) if plan else (FN_DIFF1, 0, None)
            if shift != bitshift:
                w.uvar(FN_BITSHIFT, 2); w.uvar(shift, 2)
                bitshift = shift
            x = [int(v) >> bitshift for v in channels[c][pos:pos + bs]]
            assert all((int(v) >> bitshift) << bitshift == int(v) for v in channels[c][pos:pos + bs]), "bit shift loses bits"
            if nmean == 0:
                coffset = offs[c][0]
            else:
                s = (0 if version < 2 else nmean // 2) + sum(offs[c])
                m = int(s / nmean) if s < 0 and version >= 0 else s // nmean    # C division truncates toward zero
                coffset = m if (version < 2 or bitshift == 0) else (m + (1 << (bitshift - 1))) >> bitshift
            h = hist[c] + x
            o = nwrap
            if cmd == FN_ZERO:
                assert not any(x)
                res = None
            elif cmd == FN_DIFF0:
                res = [h[o + i] - coffset for i in range(bs)]
            elif cmd == FN_DIFF1:
                res = [h[o + i] - h[o + i - 1] for i in range(bs)]
            elif cmd == FN_DIFF2:
                res = [h[o + i] - (2 * h[o + i - 1] - h[o + i - 2]) for i in range(bs)]
            elif cmd == FN_DIFF3:
                res = [h[o + i] - (3 * (h[o + i - 1] - h[o + i - 2]) + h[o + i - 3]) for i in range(bs)]
            elif cmd == FN_QLPC:
                hh = [v - coffset for v in h]
                res = []
                for i in range(bs):
                    s = lpcqoffset + sum(q[j] * hh[o + i - j - 1] for j in range(len(q)))
                    res.append(hh[o + i] - (s >> LPCQUANT))
            w.uvar(cmd, 2)
            if res is not None:
                mean_abs = sum(abs(r) for r in res) / max(1, bs)
                resn = min(7, max(0, int(mean_abs).bit_length()))
                w.uvar(resn + (1 if version == 0 else 0), 3)
                if cmd == FN_QLPC:
                    w.uvar(len(q), 2)
                    for v in q:
                        w.var(v, LPCQUANT)
                for r in res:
                    w.var(r, resn)
            if nmean > 0:
                s = (0 if version < 2 else bs // 2) + sum(x)
                m = int(s / bs)                                                  # C division truncates toward zero
                offs[c] = offs[c][1:] + [m if version < 2 else m << bitshift]
            hist[c] = (hist[c] + x)[-nwrap:]
        pos += bs
        blk += 1
    w.uvar(FN_QUIT, 2)
    return b"ajkg" + bytes([version]) + w.bytes()
