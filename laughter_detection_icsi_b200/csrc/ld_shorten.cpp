// Host-side decoder for Shorten-compressed audio (T. Robinson's "shorten", format versions 1-3), the coding of the ICSI
// meeting corpus' NIST SPHERE files ("sample_coding pcm,embedded-shorten-v2.00").  The reference reaches such files through
// lhotse's Recording.from_file / load_audio (load_data.py:44-45; cluster_scripts/gen_eval_exp.py:7 passes chanN.sph directly).
// Algorithm restated from the published format (shorten 2.x/3.x tech report and man page): a bit stream of Rice-coded
// residuals of fixed polynomial predictors (DIFF0..3) or quantised LPC, per block and channel, with a running-mean offset.
// PARITY UNPINNED: no shorten encoder or ICSI file is available offline; the SPHERE loader verifies sample_count and the
// header's sample_checksum after decoding, and tests/ round-trip an independent Python encoder of the same format.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ld_b200.h"

namespace {

constexpr int kUlongSize = 2, kNskipSize = 1, kLpcqSize = 2, kLpcQuant = 5, kXbyteSize = 7, kTypeSize = 4, kChanSize = 0,
              kEnergySize = 3, kBitshiftSize = 2, kNwrap = 3, kFnSize = 2, kVerbatimCkSize = 5, kVerbatimByteSize = 8;
enum { FN_DIFF0 = 0, FN_DIFF1, FN_DIFF2, FN_DIFF3, FN_QUIT, FN_BLOCKSIZE, FN_BITSHIFT, FN_QLPC, FN_ZERO, FN_VERBATIM };
enum { TYPE_AU1 = 0, TYPE_S8, TYPE_U8, TYPE_S16HL, TYPE_U16HL, TYPE_S16LH, TYPE_U16LH, TYPE_ULAW, TYPE_AU2, TYPE_AU3, TYPE_ALAW };

struct BitReader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint32_t buf = 0;
    int nbit = 0;
    bool eof = false;
    void refill() {   // 32-bit big-endian words; a short tail is zero-filled
        uint32_t w = 0;
        for (int i = 0; i < 4; ++i) w = (w << 8) | (pos + i < n ? p[pos + i] : 0u);
        if (pos >= n) eof = true;
        pos += 4;
        buf = w; nbit = 32;
    }
    // Rice code: the number of 0 bits before the first 1 is the high part, nbin literal low bits follow
    long uvar(int nbin) {
        if (nbit == 0) refill();
        long result = 0;
        while (!(buf & (1u << --nbit))) {
            ++result;
            if (nbit == 0) { refill(); if (eof) return -1; }
        }
        while (nbin != 0) {
            if (nbit == 0) refill();
            if (nbit >= nbin) {
                result = (result << nbin) | static_cast<long>((buf >> (nbit - nbin)) & ((nbin == 32) ? 0xFFFFFFFFu : ((1u << nbin) - 1u)));
                nbit -= nbin; nbin = 0;
            } else {
                result = (result << nbit) | static_cast<long>(buf & ((1u << nbit) - 1u));
                nbin -= nbit; nbit = 0;
            }
        }
        return result;
    }
    long var(int nbin) {   // signed: the low bit of the unsigned code is the sign
        const long u = uvar(nbin + 1);
        return (u & 1) ? ~(u >> 1) : (u >> 1);
    }
    long ulong_get() {
        const int nb = static_cast<int>(uvar(kUlongSize));
        return uvar(nb);
    }
    long uint_get(int nbit_v0, int version) { return version == 0 ? uvar(nbit_v0) : ulong_get(); }
};

thread_local std::string g_shn_err;
int shn_fail(const std::string& m) { g_shn_err = m; return LD_ERR_INVALID; }

int ilog2(long v) { int n = 0; while ((1l << (n + 1)) <= v) ++n; return n; }

}  // namespace

extern "C" {

const char* ld_shorten_last_error(void) { return g_shn_err.c_str(); }

int ld_shorten_decode(const uint8_t* data, int64_t n_bytes, int16_t* out, int64_t cap, int32_t* n_chan_out, int64_t* n_out) {
    if (!data || n_bytes < 5 || !n_out) return shn_fail("bad arguments");
    if (std::memcmp(data, "ajkg", 4) != 0) return shn_fail("not a shorten stream (magic 'ajkg' missing)");
    const int version = data[4];
    if (version > 3) return shn_fail("unsupported shorten format version " + std::to_string(version));
    BitReader br{data + 5, static_cast<size_t>(n_bytes - 5)};
    const long ftype = br.uint_get(kTypeSize, version);
    const long nchan = br.uint_get(kChanSize, version);
    long blocksize = 256, maxnlpc = 0, nmean = 0;
    if (version > 0) {
        blocksize = br.uint_get(8, version);
        maxnlpc = br.uint_get(kLpcqSize, version);
        nmean = br.uint_get(0, version);
        const long nskip = br.uint_get(kNskipSize, version);
        for (long i = 0; i < nskip; ++i) br.uvar(kXbyteSize);
    }
    if (ftype != TYPE_S16HL && ftype != TYPE_S16LH) return shn_fail("shorten file type " + std::to_string(ftype) + " is not 16-bit signed PCM");
    if (nchan < 1 || nchan > 64 || blocksize < 1 || blocksize > (1 << 20) || maxnlpc < 0 || maxnlpc > 64 || nmean < 0 || nmean > 64)
        return shn_fail("implausible shorten header");
    const int nwrap = static_cast<int>(maxnlpc > kNwrap ? maxnlpc : kNwrap);
    const long lpcqoffset = version > 1 ? (1l << (kLpcQuant - 1)) : 0;
    std::vector<std::vector<long>> buffer(nchan, std::vector<long>(static_cast<size_t>(blocksize) + nwrap, 0));
    std::vector<std::vector<long>> offset(nchan, std::vector<long>(static_cast<size_t>(nmean > 1 ? nmean : 1), 0));   // signed types: mean 0
    std::vector<long> qlpc(static_cast<size_t>(maxnlpc > 0 ? maxnlpc : 1), 0);
    std::vector<std::vector<long>> shifted(nchan, std::vector<long>(static_cast<size_t>(blocksize), 0));   // block as it is written out
    int bitshift = 0;
    long chan = 0;
    int64_t written = 0;
    for (;;) {
        const long cmd = br.uvar(kFnSize);
        if (br.eof || cmd < 0) return shn_fail("shorten stream ends without a QUIT command");
        if (cmd == FN_QUIT) break;
        if (cmd == FN_BLOCKSIZE) {
            const long nb = br.uint_get(ilog2(blocksize), version);
            if (nb < 1 || nb > blocksize) return shn_fail("shorten BLOCKSIZE command out of range");
            blocksize = nb;
            continue;
        }
        if (cmd == FN_BITSHIFT) { bitshift = static_cast<int>(br.uvar(kBitshiftSize)); continue; }
        if (cmd == FN_VERBATIM) {
            const long ck = br.uvar(kVerbatimCkSize);
            for (long i = 0; i < ck; ++i) br.uvar(kVerbatimByteSize);
            continue;
        }
        if (cmd > FN_VERBATIM) return shn_fail("unknown shorten command " + std::to_string(cmd));
        long* cb = buffer[chan].data() + nwrap;   // cb[-nwrap .. -1] = tail of the previous block of this channel
        int resn = 0;
        if (cmd != FN_ZERO) {
            resn = static_cast<int>(br.uvar(kEnergySize));
            if (version == 0) --resn;
        }
        long coffset;
        if (nmean == 0) {
            coffset = offset[chan][0];
        } else {
            long sum = version < 2 ? 0 : nmean / 2;
            for (long i = 0; i < nmean; ++i) sum += offset[chan][i];
            coffset = version < 2 ? sum / nmean : (bitshift == 0 ? sum / nmean : ((sum / nmean + (1l << (bitshift - 1))) >> bitshift));
        }
        switch (cmd) {
            case FN_ZERO:
                for (long i = 0; i < blocksize; ++i) cb[i] = 0;
                break;
            case FN_DIFF0:
                for (long i = 0; i < blocksize; ++i) cb[i] = br.var(resn) + coffset;
                break;
            case FN_DIFF1:
                for (long i = 0; i < blocksize; ++i) cb[i] = br.var(resn) + cb[i - 1];
                break;
            case FN_DIFF2:
                for (long i = 0; i < blocksize; ++i) cb[i] = br.var(resn) + (2 * cb[i - 1] - cb[i - 2]);
                break;
            case FN_DIFF3:
                for (long i = 0; i < blocksize; ++i) cb[i] = br.var(resn) + 3 * (cb[i - 1] - cb[i - 2]) + cb[i - 3];
                break;
            case FN_QLPC: {
                const long nlpc = br.uvar(kLpcqSize);
                if (nlpc > maxnlpc) return shn_fail("shorten QLPC order exceeds the header's maximum");
                for (long i = 0; i < nlpc; ++i) qlpc[i] = br.var(kLpcQuant);
                for (long i = 0; i < nlpc; ++i) cb[i - nlpc] -= coffset;
                for (long i = 0; i < blocksize; ++i) {
                    long sum = lpcqoffset;
                    for (long j = 0; j < nlpc; ++j) sum += qlpc[j] * cb[i - j - 1];
                    cb[i] = br.var(resn) + (sum >> kLpcQuant);
                }
                if (coffset != 0)
                    for (long i = 0; i < blocksize; ++i) cb[i] += coffset;
                for (long i = 0; i < nlpc; ++i) cb[i - nlpc] += coffset;   // (history restored; it is overwritten by the wrap below)
                break;
            }
        }
        if (br.eof) return shn_fail("shorten stream truncated inside a block");
        if (nmean > 0) {   // running mean of the last nmean blocks
            long sum = version < 2 ? 0 : blocksize / 2;
            for (long i = 0; i < blocksize; ++i) sum += cb[i];
            for (long i = 1; i < nmean; ++i) offset[chan][i - 1] = offset[chan][i];
            offset[chan][nmean - 1] = version < 2 ? sum / blocksize : (sum / blocksize) * (1l << bitshift);
        }
        for (int i = -nwrap; i < 0; ++i) cb[i] = cb[i + blocksize];   // history for the next block (unshifted)
        for (long i = 0; i < blocksize; ++i) shifted[chan][i] = cb[i] * (1l << bitshift);   // the bit shift in force for THIS channel's block
        if (chan == nchan - 1) {   // all channels of this block are decoded: interleave them
            for (long i = 0; i < blocksize; ++i)
                for (long c = 0; c < nchan; ++c) {
                    long v = shifted[c][i];
                    v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
                    if (out && written < cap) out[written] = static_cast<int16_t>(v);
                    ++written;
                }
        }
        chan = (chan + 1) % nchan;
    }
    if (n_chan_out) *n_chan_out = static_cast<int32_t>(nchan);
    *n_out = written;
    if (out && written > cap) return shn_fail("output buffer too small for the decoded samples");
    return LD_OK;
}

}  // extern "C"
