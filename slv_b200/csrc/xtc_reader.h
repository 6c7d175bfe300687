// xtc_reader.h -- GROMACS XTC trajectory reader (host code behind the C ABI: slvgpu_xtc_scan / slvgpu_xtc_read).
//
// The reference reads XTC through MDAnalysis (util/slv_md-run:117-152, the `gromacs` branch -- the format the tutorial
// uses, doc/tutor/IV.Molecular-dynamics.md); MDAnalysis is not available here, so the published xdr3dcoord format is
// decoded directly: XDR (big-endian) frame header -- magic 1995, natoms, step, time, box[9] -- then the coordinate
// block: natoms; for natoms <= 9 raw floats; else precision, minint[3], maxint[3], smallidx, byte count and a bit
// stream in which every atom is either a "large" triple (mixed-radix number of sizeofints(maxint - minint + 1) bits) or,
// inside a run, a "small" triple of `smallidx` bits relative to the previous atom (radix magicints[smallidx]); the first
// two atoms of a run are stored swapped (water oxygen / hydrogen trick).  Coordinates come out as float32 nm records,
// exactly what the file holds.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

namespace slvxtc {

static const int magicints[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
                                645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
                                32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287,
                                660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042,
                                8388607, 10568983, 13316085, 16777216};
constexpr int FIRSTIDX = 9;
constexpr int LASTIDX = (int)(sizeof(magicints) / sizeof(magicints[0]));

struct BitReader {                      // MSB-first bit stream
  const unsigned char *p; size_t n, pos = 0; unsigned lastbits = 0; unsigned lastbyte = 0; bool bad = false;
  BitReader(const unsigned char *p_, size_t n_) : p(p_), n(n_) {}
  unsigned byte() { if (pos >= n) { bad = true; return 0; } return p[pos++]; }
  int bits(int nb) {
    const unsigned mask = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
    unsigned num = 0;
    while (nb >= 8) { lastbyte = (lastbyte << 8) | byte(); num |= (lastbyte >> lastbits) << (nb - 8); nb -= 8; }
    if (nb > 0) {
      if (lastbits < (unsigned)nb) { lastbits += 8; lastbyte = (lastbyte << 8) | byte(); }
      lastbits -= nb;
      num |= (lastbyte >> lastbits) & ((1u << nb) - 1u);
    }
    return (int)(num & mask);
  }
  // three integers packed as one mixed-radix number of nb bits (bytes least significant first)
  void ints(int nb, const unsigned sizes[3], int nums[3]) {
    int bytes[32]; int nbytes = 0;
    bytes[1] = bytes[2] = bytes[3] = 0;
    while (nb > 8) { bytes[nbytes++] = bits(8); nb -= 8; }
    if (nb > 0) bytes[nbytes++] = bits(nb);
    for (int i = 2; i > 0; --i) {
      unsigned num = 0;
      for (int j = nbytes - 1; j >= 0; --j) {
        num = (num << 8) | (unsigned)bytes[j];
        const unsigned q = num / sizes[i];
        bytes[j] = (int)q;
        num -= q * sizes[i];
      }
      nums[i] = (int)num;
    }
    nums[0] = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
  }
};

inline int sizeofint(unsigned size) {
  unsigned num = 1; int nb = 0;
  while (size >= num && nb < 32) { ++nb; num <<= 1; }
  return nb;
}
inline int sizeofints(const unsigned sizes[3]) {
  unsigned bytes[32]; unsigned nbytes = 1; bytes[0] = 1;
  for (int i = 0; i < 3; ++i) {
    unsigned tmp = 0, bc = 0;
    for (bc = 0; bc < nbytes; ++bc) { tmp = bytes[bc] * sizes[i] + tmp; bytes[bc] = tmp & 0xff; tmp >>= 8; }
    while (tmp != 0) { bytes[bc++] = tmp & 0xff; tmp >>= 8; }
    nbytes = bc;
  }
  unsigned num = 1; int nb = 0;
  --nbytes;
  while (bytes[nbytes] >= num) { ++nb; num *= 2; }
  return nb + (int)nbytes * 8;
}

struct File {
  FILE *fh = nullptr;
  ~File() { if (fh) fclose(fh); }
  bool rd(void *dst, size_t n) { return fread(dst, 1, n, fh) == n; }
  bool i32(int32_t &v) { unsigned char b[4]; if (!rd(b, 4)) return false; v = (int32_t)((uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3]); return true; }
  bool f32(float &v) { int32_t i; if (!i32(i)) return false; memcpy(&v, &i, 4); return true; }
};

// Reads one frame at the current position.  out == nullptr: skip the coordinates (scan).  Returns 1 on success, 0 at a
// clean end of file, -1 on a malformed frame.
inline int read_frame(File &f, int32_t &natoms, int32_t &step, float &time, float box[9], float *out, std::vector<unsigned char> &buf,
                      std::string &err) {
  int32_t magic;
  if (!f.i32(magic)) return 0;
  if (magic != 1995) { err = "not an XTC frame (magic number)"; return -1; }
  if (!f.i32(natoms) || !f.i32(step) || !f.f32(time)) { err = "truncated XTC header"; return -1; }
  for (int k = 0; k < 9; ++k) if (!f.f32(box[k])) { err = "truncated XTC header"; return -1; }
  int32_t size;
  if (!f.i32(size) || size != natoms || natoms < 0) { err = "inconsistent atom count in XTC frame"; return -1; }
  if (size <= 9) {
    for (int k = 0; k < 3 * size; ++k) { float v; if (!f.f32(v)) { err = "truncated XTC frame"; return -1; } if (out) out[k] = v; }
    return 1;
  }
  float precision; int32_t minint[3], maxint[3], smallidx, nbytes;
  if (!f.f32(precision)) { err = "truncated XTC frame"; return -1; }
  for (int k = 0; k < 3; ++k) if (!f.i32(minint[k])) { err = "truncated XTC frame"; return -1; }
  for (int k = 0; k < 3; ++k) if (!f.i32(maxint[k])) { err = "truncated XTC frame"; return -1; }
  if (!f.i32(smallidx) || !f.i32(nbytes) || nbytes < 0 || smallidx < FIRSTIDX || smallidx >= LASTIDX) { err = "bad XTC compression header"; return -1; }
  const size_t padded = ((size_t)nbytes + 3) & ~(size_t)3;
  if (!out) { if (fseek(f.fh, (long)padded, SEEK_CUR) != 0) { err = "seek failed"; return -1; } return 1; }
  buf.resize(padded + 8);
  if (!f.rd(buf.data(), padded)) { err = "truncated XTC frame"; return -1; }
  memset(buf.data() + padded, 0, 8);
  unsigned sizeint[3], bitsizeint[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) sizeint[k] = (unsigned)(maxint[k] - minint[k] + 1);
  int bitsize;
  if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffff) {
    for (int k = 0; k < 3; ++k) bitsizeint[k] = (unsigned)sizeofint(sizeint[k]);
    bitsize = 0;
  } else bitsize = sizeofints(sizeint);
  int tmpidx = smallidx - 1; if (tmpidx < FIRSTIDX) tmpidx = FIRSTIDX;
  int smaller = magicints[tmpidx] / 2, smallnum = magicints[smallidx] / 2;
  unsigned sizesmall[3] = {(unsigned)magicints[smallidx], (unsigned)magicints[smallidx], (unsigned)magicints[smallidx]};
  BitReader br(buf.data(), padded + 8);
  const float inv = 1.0f / precision;
  int run = 0, i = 0;
  float *lfp = out;
  while (i < size) {
    int thiscoord[3], prevcoord[3];
    if (bitsize == 0) { for (int k = 0; k < 3; ++k) thiscoord[k] = br.bits((int)bitsizeint[k]); }
    else br.ints(bitsize, sizeint, thiscoord);
    ++i;
    for (int k = 0; k < 3; ++k) { thiscoord[k] += minint[k]; prevcoord[k] = thiscoord[k]; }
    const int flag = br.bits(1);
    int is_smaller = 0;
    if (flag == 1) { run = br.bits(5); is_smaller = run % 3; run -= is_smaller; --is_smaller; }
    if (run > 0) {
      if (i + run / 3 > size) { err = "XTC run exceeds the atom count"; return -1; }
      for (int k = 0; k < run; k += 3) {
        br.ints(smallidx, sizesmall, thiscoord);
        ++i;
        for (int d = 0; d < 3; ++d) thiscoord[d] += prevcoord[d] - smallnum;
        if (k == 0) {            // the first two atoms of a run are stored in swapped order
          for (int d = 0; d < 3; ++d) { const int t = thiscoord[d]; thiscoord[d] = prevcoord[d]; prevcoord[d] = t; }
          for (int d = 0; d < 3; ++d) *lfp++ = prevcoord[d] * inv;
        } else {
          for (int d = 0; d < 3; ++d) prevcoord[d] = thiscoord[d];
        }
        for (int d = 0; d < 3; ++d) *lfp++ = thiscoord[d] * inv;
      }
    } else {
      for (int d = 0; d < 3; ++d) *lfp++ = thiscoord[d] * inv;
    }
    smallidx += is_smaller;
    if (smallidx < FIRSTIDX || smallidx >= LASTIDX) { err = "corrupt XTC stream (small index)"; return -1; }
    if (is_smaller < 0) { smallnum = smaller; smaller = smallidx > FIRSTIDX ? magicints[smallidx - 1] / 2 : 0; }
    else if (is_smaller > 0) { smaller = smallnum; smallnum = magicints[smallidx] / 2; }
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
    if (br.bad) { err = "XTC bit stream ends inside a frame"; return -1; }
  }
  return 1;
}

}  // namespace slvxtc
