// swr_jpeg.hpp — JPEG (JFIF) decoder for the glTF loader's textures: baseline / extended sequential and progressive
// Huffman DCT, 8-bit, 1 or 3 components, any sampling factors the common encoders write (4:4:4, 4:2:2, 4:2:0, 4:4:0,
// 4:1:1), restart intervals, 16-bit quantisation tables, Adobe APP14 colour transform flag. No arithmetic coding, no
// 12-bit, no CMYK, no lossless / hierarchical modes (rejected with a message).
//
// The reference decodes textures with the `image` crate (0.25, zune-jpeg underneath); that code is not vendored, so this
// is written from the JPEG specification (ITU-T T.81) and follows the de-facto standard reconstruction of libjpeg /
// libjpeg-turbo, which zune-jpeg also targets: the "islow" integer IDCT (13-bit constants, two passes), triangle-filter
// ("fancy") chroma upsampling for 2:1 horizontal and 2:1 x 2:1, replication otherwise, and the 16-bit fixed-point
// YCbCr -> RGB tables. tests/test_jpeg.py checks it bit for bit against libjpeg-turbo (through Pillow).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace swr {
namespace jpeg {

struct DecodedImage {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba;
};

class Decoder {
   public:
    Decoder(const uint8_t *data, size_t size, const std::string &name) : p_(data), n_(size), name_(name) {}

    DecodedImage decode() {
        if (n_ < 4 || p_[0] != 0xFF || p_[1] != 0xD8) fail("not a JPEG file");
        pos_ = 2;
        bool done = false;
        while (!done) {
            const int m = next_marker();
            switch (m) {
                case 0xC0: case 0xC1: case 0xC2: read_sof(m); break;
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    fail("unsupported JPEG process (lossless, hierarchical or arithmetic coding)");
                case 0xC4: read_dht(); break;
                case 0xDB: read_dqt(); break;
                case 0xDD: {
                    const size_t len = seg_len();
                    if (len != 4) fail("bad DRI segment");
                    restart_interval_ = be16(pos_ + 2);
                    pos_ += len;
                    break;
                }
                case 0xDA: read_scan(); break;
                case 0xD9: done = true; break;
                case 0xEE: read_app14(); break;
                default:
                    if (m >= 0xD0 && m <= 0xD7) break;  // stray restart marker
                    pos_ += seg_len();                  // APPn, COM, DNL, ...: skipped
                    break;
            }
        }
        if (!have_frame_) fail("no frame header");
        return reconstruct();
    }

   private:
    const uint8_t *p_;
    size_t n_, pos_ = 0;
    int scans_ = 0;
    std::string name_;

    struct Component {
        int id = 0, h = 1, v = 1, tq = 0;
        int td = 0, ta = 0;          // Huffman table selectors of the current scan
        int blocks_w = 0, blocks_h = 0;  // allocated blocks (whole MCUs)
        int width = 0, height = 0;   // downsampled size in samples: ceil(W * h / hmax)
        int pred = 0;
        std::vector<int16_t> coef;   // blocks_w * blocks_h * 64, natural order
    };
    struct Huff {
        bool set = false;
        uint8_t bits[17] = {0};
        uint8_t vals[256] = {0};
        int mincode[17] = {0}, maxcode[18] = {0}, valptr[17] = {0};
        uint16_t fast[512];  // 9-bit lookahead: (len << 8) | value, 0 = longer code
    };
    uint16_t qt_[4][64] = {{0}};
    bool qt_set_[4] = {false, false, false, false};
    Huff dc_[4], ac_[4];
    std::vector<Component> comp_;
    bool have_frame_ = false, progressive_ = false;
    int width_ = 0, height_ = 0, hmax_ = 1, vmax_ = 1, mcus_x_ = 0, mcus_y_ = 0;
    int restart_interval_ = 0;
    bool adobe_ = false;
    int adobe_transform_ = -1;
    // bit reader
    uint32_t bitbuf_ = 0;
    int bitcnt_ = 0;
    bool hit_marker_ = false;
    int eobrun_ = 0;

    [[noreturn]] void fail(const char *why) const { throw std::runtime_error("Missing data: Could not load texture '" + name_ + "': JPEG: " + why); }
    uint32_t be16(size_t o) const {
        if (o + 2 > n_) fail("truncated file");
        return ((uint32_t)p_[o] << 8) | p_[o + 1];
    }
    // DC predictors of a conforming stream stay inside 16 bits; a hostile one must not walk an int into overflow
    static int clamp_pred(int v) { return std::min(std::max(v, -32768), 32767); }
    size_t seg_len() const {
        const size_t len = be16(pos_);
        if (len < 2 || pos_ + len > n_) fail("truncated segment");
        return len;
    }
    int next_marker() {
        while (pos_ < n_ && p_[pos_] != 0xFF) pos_++;  // garbage between segments is tolerated
        while (pos_ < n_ && p_[pos_] == 0xFF) pos_++;  // fill bytes
        if (pos_ >= n_) fail("unexpected end of file (no EOI)");
        return p_[pos_++];
    }

    static const uint8_t *zigzag() {
        static const uint8_t z[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                      41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                      30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        return z;
    }

    void read_dqt() {
        const size_t len = seg_len(), end = pos_ + len;
        pos_ += 2;
        while (pos_ < end) {
            const int pq = p_[pos_] >> 4, tq = p_[pos_] & 15;
            pos_++;
            if (tq > 3 || pq > 1) fail("bad quantisation table header");
            if (pos_ + (size_t)64 * (pq + 1) > end) fail("truncated quantisation table");
            for (int i = 0; i < 64; i++) {
                const uint16_t v = pq ? (uint16_t)be16(pos_ + 2 * (size_t)i) : p_[pos_ + (size_t)i];
                qt_[tq][zigzag()[i]] = v;
            }
            pos_ += (size_t)64 * (pq + 1);
            qt_set_[tq] = true;
        }
    }

    void read_dht() {
        const size_t len = seg_len(), end = pos_ + len;
        pos_ += 2;
        while (pos_ < end) {
            const int tc = p_[pos_] >> 4, th = p_[pos_] & 15;
            pos_++;
            if (tc > 1 || th > 3) fail("bad Huffman table header");
            Huff &h = tc ? ac_[th] : dc_[th];
            if (pos_ + 16 > end) fail("truncated Huffman table");
            int total = 0;
            h.bits[0] = 0;
            for (int i = 1; i <= 16; i++) {
                h.bits[i] = p_[pos_ + (size_t)i - 1];
                total += h.bits[i];
            }
            pos_ += 16;
            if (total > 256 || pos_ + (size_t)total > end) fail("bad Huffman table");
            std::memcpy(h.vals, p_ + pos_, (size_t)total);
            pos_ += (size_t)total;
            // canonical codes (T.81 Annex C) + decoding tables (Annex F.2.2.3)
            int code = 0, k = 0;
            for (int l = 1; l <= 16; l++) {
                h.valptr[l] = k;
                h.mincode[l] = code;
                code += h.bits[l];
                k += h.bits[l];
                h.maxcode[l] = h.bits[l] ? code - 1 : -1;
                if (code > (1 << l)) fail("over-subscribed Huffman table");
                code <<= 1;
            }
            h.maxcode[17] = 0x7FFFFFFF;
            std::memset(h.fast, 0, sizeof(h.fast));
            code = 0, k = 0;
            for (int l = 1; l <= 9; l++) {
                for (int i = 0; i < h.bits[l]; i++, k++, code++) {
                    const int first = code << (9 - l), count = 1 << (9 - l);
                    for (int j = 0; j < count; j++) h.fast[first + j] = (uint16_t)((l << 8) | h.vals[k]);
                }
                code <<= 1;
            }
            h.set = true;
        }
    }

    void read_app14() {
        const size_t len = seg_len();
        if (len >= 14 && std::memcmp(p_ + pos_ + 2, "Adobe", 5) == 0) {
            adobe_ = true;
            adobe_transform_ = p_[pos_ + 13];
        }
        pos_ += len;
    }

    void read_sof(int marker) {
        if (have_frame_) fail("more than one frame header");
        const size_t len = seg_len();
        if (len < 8) fail("bad frame header");
        if (p_[pos_ + 2] != 8) fail("only 8-bit samples are supported");
        height_ = (int)be16(pos_ + 3);
        width_ = (int)be16(pos_ + 5);
        const int nc = p_[pos_ + 7];
        if (width_ == 0 || height_ == 0) fail("empty image (DNL-defined height is not supported)");
        if (nc != 1 && nc != 3) fail("only greyscale and 3-component images are supported");
        if (len != 8 + 3 * (size_t)nc) fail("bad frame header length");
        comp_.resize((size_t)nc);
        hmax_ = vmax_ = 1;
        for (int i = 0; i < nc; i++) {
            Component &c = comp_[(size_t)i];
            c.id = p_[pos_ + 8 + 3 * (size_t)i];
            c.h = p_[pos_ + 9 + 3 * (size_t)i] >> 4;
            c.v = p_[pos_ + 9 + 3 * (size_t)i] & 15;
            c.tq = p_[pos_ + 10 + 3 * (size_t)i];
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) fail("bad component parameters");
            hmax_ = std::max(hmax_, c.h);
            vmax_ = std::max(vmax_, c.v);
        }
        if (nc == 1) comp_[0].h = comp_[0].v = hmax_ = vmax_ = 1;  // a single component is never interleaved: factors are irrelevant
        mcus_x_ = (width_ + 8 * hmax_ - 1) / (8 * hmax_);
        mcus_y_ = (height_ + 8 * vmax_ - 1) / (8 * vmax_);
        for (Component &c : comp_) {
            c.blocks_w = mcus_x_ * c.h;
            c.blocks_h = mcus_y_ * c.v;
            c.width = (width_ * c.h + hmax_ - 1) / hmax_;
            c.height = (height_ * c.v + vmax_ - 1) / vmax_;
            if ((size_t)c.blocks_w * c.blocks_h > ((size_t)1 << 22)) fail("image too large (more than 16384 x 16384 samples in one component)");
            c.coef.assign((size_t)c.blocks_w * c.blocks_h * 64, 0);
        }
        progressive_ = marker == 0xC2;
        have_frame_ = true;
        pos_ += len;
    }

    // ---- entropy-coded segment ------------------------------------------------------------------------------------
    void fill_bits() {
        while (bitcnt_ <= 24) {
            uint32_t b = 0;
            if (!hit_marker_ && pos_ < n_) {
                b = p_[pos_];
                if (b == 0xFF) {
                    const uint8_t nx = pos_ + 1 < n_ ? p_[pos_ + 1] : 0xD9;
                    if (nx == 0) {
                        pos_ += 2;  // stuffed zero
                    } else {
                        hit_marker_ = true;  // leave the marker in place; feed zeros (T.81 F.2.2.5 note)
                        b = 0;
                    }
                } else {
                    pos_++;
                }
            } else {
                hit_marker_ = true;
            }
            bitbuf_ |= b << (24 - bitcnt_);
            bitcnt_ += 8;
        }
    }
    int get_bits(int n) {
        if (n == 0) return 0;
        if (bitcnt_ < n) fill_bits();
        const int v = (int)(bitbuf_ >> (32 - n));
        bitbuf_ <<= n;
        bitcnt_ -= n;
        return v;
    }
    int get_bit() { return get_bits(1); }
    int decode_huff(const Huff &h) {
        if (bitcnt_ < 16) fill_bits();
        const uint16_t f = h.fast[bitbuf_ >> 23];
        if (f) {
            const int l = f >> 8;
            bitbuf_ <<= l;
            bitcnt_ -= l;
            return f & 255;
        }
        int code = (int)(bitbuf_ >> 23), l = 9;  // 9 bits taken so far
        for (l = 10; l <= 16; l++) {
            code = (int)(bitbuf_ >> (32 - l));
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) break;
        }
        if (l > 16) fail("corrupt data: bad Huffman code");
        bitbuf_ <<= l;
        bitcnt_ -= l;
        return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    static int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }  // T.81 F.2.2.1
    int receive_extend(int t) { return t ? extend(get_bits(t), t) : 0; }

    void reset_entropy() {
        bitbuf_ = 0;
        bitcnt_ = 0;
        hit_marker_ = false;
        eobrun_ = 0;
        for (Component &c : comp_) c.pred = 0;
    }
    void process_restart(int expected) {
        // discard remaining bits, expect RSTn
        bitbuf_ = 0;
        bitcnt_ = 0;
        hit_marker_ = false;
        while (pos_ < n_ && p_[pos_] != 0xFF) pos_++;
        while (pos_ + 1 < n_ && p_[pos_] == 0xFF && p_[pos_ + 1] == 0xFF) pos_++;
        if (pos_ + 1 < n_ && p_[pos_] == 0xFF && p_[pos_ + 1] == 0xD0 + (expected & 7)) pos_ += 2;
        // a missing / wrong restart marker: carry on (libjpeg resynchronises too), predictors are reset either way
        eobrun_ = 0;
        for (Component &c : comp_) c.pred = 0;
    }

    // ---- block decoders ---------------------------------------------------------------------------------------------
    void decode_block_sequential(Component &c, int16_t *blk) {
        const Huff &hd = dc_[c.td], &ha = ac_[c.ta];
        const int t = decode_huff(hd);
        if (t > 11) fail("corrupt data: bad DC size");
        c.pred = clamp_pred(c.pred + receive_extend(t));
        blk[0] = (int16_t)c.pred;
        for (int k = 1; k < 64;) {
            const int rs = decode_huff(ha), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
            }
            k += r;
            if (k > 63) fail("corrupt data: coefficient index out of range");
            blk[zigzag()[k++]] = (int16_t)receive_extend(s);
        }
    }
    void decode_dc_first(Component &c, int16_t *blk, int al) {
        const int t = decode_huff(dc_[c.td]);
        if (t > 11) fail("corrupt data: bad DC size");
        c.pred = clamp_pred(c.pred + receive_extend(t));
        blk[0] = (int16_t)(c.pred * (1 << al));
    }
    void decode_dc_refine(int16_t *blk, int al) {
        if (get_bit()) blk[0] = (int16_t)(blk[0] | (1 << al));
    }
    void decode_ac_first(Component &c, int16_t *blk, int ss, int se, int al) {
        if (eobrun_ > 0) {
            eobrun_--;
            return;
        }
        const Huff &ha = ac_[c.ta];
        for (int k = ss; k <= se;) {
            const int rs = decode_huff(ha), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r < 15) {
                    eobrun_ = (1 << r) - 1;
                    if (r) eobrun_ += get_bits(r);
                    break;
                }
                k += 16;
                continue;
            }
            k += r;
            if (k > 63) fail("corrupt data: coefficient index out of range");
            blk[zigzag()[k++]] = (int16_t)(receive_extend(s) * (1 << al));
        }
    }
    void decode_ac_refine(Component &c, int16_t *blk, int ss, int se, int al) {  // T.81 G.1.2.3
        const int p1 = 1 << al, m1 = -(1 << al);
        const Huff &ha = ac_[c.ta];
        int k = ss;
        if (eobrun_ == 0) {
            for (; k <= se;) {
                const int rs = decode_huff(ha);
                int r = rs >> 4;
                const int s = rs & 15;
                int value = 0;
                if (s == 0) {
                    if (r < 15) {
                        eobrun_ = 1 << r;
                        if (r) eobrun_ += get_bits(r);
                        break;  // the rest of the band only gets correction bits
                    }
                } else {
                    if (s != 1) fail("corrupt data: bad refinement code");
                    value = get_bit() ? p1 : m1;
                }
                // advance over r zero-history coefficients, correcting the non-zero ones on the way
                while (k <= se) {
                    int16_t &co = blk[zigzag()[k]];
                    if (co != 0) {
                        if (get_bit() && (co & p1) == 0) co = (int16_t)(co >= 0 ? co + p1 : co + m1);
                    } else {
                        if (r == 0) {
                            if (value) co = (int16_t)value;
                            k++;
                            break;
                        }
                        r--;
                    }
                    k++;
                }
            }
        }
        if (eobrun_ > 0) {
            for (; k <= se; k++) {
                int16_t &co = blk[zigzag()[k]];
                if (co != 0 && get_bit() && (co & p1) == 0) co = (int16_t)(co >= 0 ? co + p1 : co + m1);
            }
            eobrun_--;
        }
    }

    void read_scan() {
        if (!have_frame_) fail("scan before frame header");
        const size_t len = seg_len();
        if (len < 6) fail("bad scan header");  // seg_len only guarantees the two length bytes
        if (++scans_ > 1000) fail("too many scans");  // a hostile file could make every 10-byte scan header cost a pass over all blocks
        const int ns = p_[pos_ + 2];
        if (ns < 1 || ns > (int)comp_.size() || len != 6 + 2 * (size_t)ns) fail("bad scan header");
        std::vector<int> sel((size_t)ns);
        for (int i = 0; i < ns; i++) {
            const int id = p_[pos_ + 3 + 2 * (size_t)i], tt = p_[pos_ + 4 + 2 * (size_t)i];
            int ci = -1;
            for (size_t k = 0; k < comp_.size(); k++)
                if (comp_[k].id == id) ci = (int)k;
            if (ci < 0) fail("scan names an unknown component");
            comp_[(size_t)ci].td = tt >> 4;
            comp_[(size_t)ci].ta = tt & 15;
            if (comp_[(size_t)ci].td > 3 || comp_[(size_t)ci].ta > 3) fail("bad Huffman table selector");
            sel[(size_t)i] = ci;
        }
        const int ss = p_[pos_ + 3 + 2 * (size_t)ns], se = p_[pos_ + 4 + 2 * (size_t)ns], ah = p_[pos_ + 5 + 2 * (size_t)ns] >> 4,
                  al = p_[pos_ + 5 + 2 * (size_t)ns] & 15;
        pos_ += len;
        if (progressive_) {
            if (ss > se || se > 63 || al > 13 || (ss == 0 && se != 0) || (ss > 0 && ns != 1)) fail("bad progressive scan parameters");
        } else if (ss != 0 || se != 63 || ah != 0 || al != 0) {
            fail("bad sequential scan parameters");
        }
        for (int ci : sel) {
            const Component &c = comp_[(size_t)ci];
            const bool need_dc = !progressive_ || (ss == 0 && ah == 0), need_ac = !progressive_ || ss > 0;
            if (need_dc && !dc_[c.td].set) fail("scan uses an undefined DC Huffman table");
            if (need_ac && !ac_[c.ta].set) fail("scan uses an undefined AC Huffman table");
        }
        reset_entropy();
        int rst_count = 0, rst_index = 0;
        auto mcu_done = [&](bool last) {
            if (restart_interval_ && !last && ++rst_count == restart_interval_) {
                process_restart(rst_index++);
                rst_count = 0;
            }
        };
        auto decode_one = [&](Component &c, int bx, int by) {
            int16_t *blk = &c.coef[((size_t)by * c.blocks_w + bx) * 64];
            if (!progressive_)
                decode_block_sequential(c, blk);
            else if (ss == 0)
                ah == 0 ? decode_dc_first(c, blk, al) : decode_dc_refine(blk, al);
            else
                ah == 0 ? decode_ac_first(c, blk, ss, se, al) : decode_ac_refine(c, blk, ss, se, al);
        };
        if (ns == 1) {  // non-interleaved: the component's own block raster, only blocks that hold image data
            Component &c = comp_[(size_t)sel[0]];
            const int bw = (c.width + 7) / 8, bh = (c.height + 7) / 8;
            for (int by = 0; by < bh; by++)
                for (int bx = 0; bx < bw; bx++) {
                    decode_one(c, bx, by);
                    mcu_done(by == bh - 1 && bx == bw - 1);
                }
        } else {
            for (int my = 0; my < mcus_y_; my++)
                for (int mx = 0; mx < mcus_x_; mx++) {
                    for (int ci : sel) {
                        Component &c = comp_[(size_t)ci];
                        for (int v = 0; v < c.v; v++)
                            for (int h = 0; h < c.h; h++) decode_one(c, mx * c.h + h, my * c.v + v);
                    }
                    mcu_done(my == mcus_y_ - 1 && mx == mcus_x_ - 1);
                }
        }
        // leave pos_ at the marker that ended the entropy-coded data
        if (!hit_marker_) {
            while (pos_ + 1 < n_ && !(p_[pos_] == 0xFF && p_[pos_ + 1] != 0 && !(p_[pos_ + 1] >= 0xD0 && p_[pos_ + 1] <= 0xD7))) pos_++;
        }
    }

    // ---- reconstruction ---------------------------------------------------------------------------------------------
    static uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

    // libjpeg jidctint.c ("islow"): CONST_BITS 13, PASS1_BITS 2
    static void idct_islow(const int16_t *in, const uint16_t *q, uint8_t *out, size_t out_stride) {
        const int C = 13, P = 2;
        const int F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299, F1_847 = 15137,
                  F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
        int ws[64];
        auto descale = [](long long x, int n) { return (int)((x + ((long long)1 << (n - 1))) >> n); };
        for (int col = 0; col < 8; col++) {
            long long z2 = in[16 + col] * (int)q[16 + col], z3 = in[48 + col] * (int)q[48 + col];
            long long z1 = (z2 + z3) * F0_541;
            long long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
            z2 = in[col] * (int)q[col];
            z3 = in[32 + col] * (int)q[32 + col];
            long long tmp0 = (z2 + z3) * ((long long)1 << C), tmp1 = (z2 - z3) * ((long long)1 << C);
            const long long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = in[56 + col] * (int)q[56 + col];
            tmp1 = in[40 + col] * (int)q[40 + col];
            tmp2 = in[24 + col] * (int)q[24 + col];
            tmp3 = in[8 + col] * (int)q[8 + col];
            z1 = tmp0 + tmp3;
            z2 = tmp1 + tmp2;
            z3 = tmp0 + tmp2;
            long long z4 = tmp1 + tmp3;
            const long long z5 = (z3 + z4) * F1_175;
            tmp0 *= F0_298, tmp1 *= F2_053, tmp2 *= F3_072, tmp3 *= F1_501;
            z1 *= -F0_899, z2 *= -F2_562, z3 *= -F1_961, z4 *= -F0_390;
            z3 += z5, z4 += z5;
            tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
            ws[col] = descale(tmp10 + tmp3, C - P);
            ws[56 + col] = descale(tmp10 - tmp3, C - P);
            ws[8 + col] = descale(tmp11 + tmp2, C - P);
            ws[48 + col] = descale(tmp11 - tmp2, C - P);
            ws[16 + col] = descale(tmp12 + tmp1, C - P);
            ws[40 + col] = descale(tmp12 - tmp1, C - P);
            ws[24 + col] = descale(tmp13 + tmp0, C - P);
            ws[32 + col] = descale(tmp13 - tmp0, C - P);
        }
        for (int row = 0; row < 8; row++) {
            const int *w = ws + 8 * row;
            long long z2 = w[2], z3 = w[6];
            long long z1 = (z2 + z3) * F0_541;
            long long tmp2 = z1 + z3 * (-F1_847), tmp3 = z1 + z2 * F0_765;
            long long tmp0 = ((long long)w[0] + w[4]) * ((long long)1 << C), tmp1 = ((long long)w[0] - w[4]) * ((long long)1 << C);
            const long long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = w[7], tmp1 = w[5], tmp2 = w[3], tmp3 = w[1];
            z1 = tmp0 + tmp3;
            z2 = tmp1 + tmp2;
            z3 = tmp0 + tmp2;
            long long z4 = tmp1 + tmp3;
            const long long z5 = (z3 + z4) * F1_175;
            tmp0 *= F0_298, tmp1 *= F2_053, tmp2 *= F3_072, tmp3 *= F1_501;
            z1 *= -F0_899, z2 *= -F2_562, z3 *= -F1_961, z4 *= -F0_390;
            z3 += z5, z4 += z5;
            tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
            uint8_t *o = out + (size_t)row * out_stride;
            const int S = C + P + 3;
            o[0] = clamp8(descale(tmp10 + tmp3, S) + 128);
            o[7] = clamp8(descale(tmp10 - tmp3, S) + 128);
            o[1] = clamp8(descale(tmp11 + tmp2, S) + 128);
            o[6] = clamp8(descale(tmp11 - tmp2, S) + 128);
            o[2] = clamp8(descale(tmp12 + tmp1, S) + 128);
            o[5] = clamp8(descale(tmp12 - tmp1, S) + 128);
            o[3] = clamp8(descale(tmp13 + tmp0, S) + 128);
            o[4] = clamp8(descale(tmp13 - tmp0, S) + 128);
        }
    }

    // component plane (downsampled) -> full resolution, libjpeg jdsample.c
    std::vector<uint8_t> upsample(const Component &c, const std::vector<uint8_t> &plane, int pstride) const {
        const int W = width_, H = height_;
        std::vector<uint8_t> out((size_t)W * H);
        const int hr = hmax_ / c.h, vr = vmax_ / c.v;
        const bool exact = hmax_ % c.h == 0 && vmax_ % c.v == 0;
        if (!exact) fail("fractional sampling ratios are not supported");
        const int cw = c.width, chh = c.height;
        auto at = [&](int x, int y) -> int { return plane[(size_t)std::min(std::max(y, 0), chh - 1) * pstride + std::min(std::max(x, 0), cw - 1)]; };
        if (hr == 1 && vr == 1) {
            for (int y = 0; y < H; y++) std::memcpy(&out[(size_t)y * W], &plane[(size_t)y * pstride], (size_t)W);
        } else if (hr == 2 && vr == 1 && cw > 2) {  // h2v1_fancy_upsample
            for (int y = 0; y < H; y++)
                for (int x = 0; x < cw; x++) {
                    const int cur = at(x, y);
                    int a, b;
                    if (x == 0) {
                        a = cur;
                        b = (cur * 3 + at(1, y) + 2) >> 2;
                    } else if (x == cw - 1) {
                        a = (cur * 3 + at(x - 1, y) + 1) >> 2;
                        b = cur;
                    } else {
                        a = (cur * 3 + at(x - 1, y) + 1) >> 2;
                        b = (cur * 3 + at(x + 1, y) + 2) >> 2;
                    }
                    if (2 * x < W) out[(size_t)y * W + 2 * x] = (uint8_t)a;
                    if (2 * x + 1 < W) out[(size_t)y * W + 2 * x + 1] = (uint8_t)b;
                }
        } else if (hr == 2 && vr == 2 && cw > 2) {  // h2v2_fancy_upsample: rows above / below are replicated at the image edges
            for (int y = 0; y < H; y++) {
                const int cy = y >> 1, oy = (y & 1) ? cy + 1 : cy - 1;
                for (int x = 0; x < cw; x++) {
                    const int cur = 3 * at(x, cy) + at(x, oy);
                    int a, b;
                    if (x == 0) {
                        const int next = 3 * at(1, cy) + at(1, oy);
                        a = (cur * 4 + 8) >> 4;
                        b = (cur * 3 + next + 7) >> 4;
                    } else if (x == cw - 1) {
                        const int last = 3 * at(x - 1, cy) + at(x - 1, oy);
                        a = (cur * 3 + last + 8) >> 4;
                        b = (cur * 4 + 7) >> 4;
                    } else {
                        const int last = 3 * at(x - 1, cy) + at(x - 1, oy), next = 3 * at(x + 1, cy) + at(x + 1, oy);
                        a = (cur * 3 + last + 8) >> 4;
                        b = (cur * 3 + next + 7) >> 4;
                    }
                    if (2 * x < W) out[(size_t)y * W + 2 * x] = (uint8_t)a;
                    if (2 * x + 1 < W) out[(size_t)y * W + 2 * x + 1] = (uint8_t)b;
                }
            }
        } else if (hr == 1 && vr == 2) {  // h1v2_fancy_upsample (4:4:0): vertical triangle filter, rows replicated at the edges
            for (int y = 0; y < H; y++) {
                const int cy = y >> 1, oy = (y & 1) ? cy + 1 : cy - 1, bias = (y & 1) ? 2 : 1;
                for (int x = 0; x < W; x++) out[(size_t)y * W + x] = (uint8_t)((3 * at(x, cy) + at(x, oy) + bias) >> 2);
            }
        } else {  // int_upsample / h2v1_upsample / h2v2_upsample: replication
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) out[(size_t)y * W + x] = (uint8_t)at(x / hr, y / vr);
        }
        return out;
    }

    DecodedImage reconstruct() {
        for (const Component &c : comp_)
            if (!qt_set_[c.tq]) fail("component uses an undefined quantisation table");
        std::vector<std::vector<uint8_t>> full;
        for (const Component &c : comp_) {
            const int pstride = c.blocks_w * 8;
            std::vector<uint8_t> plane((size_t)pstride * c.blocks_h * 8);
            for (int by = 0; by < c.blocks_h; by++)
                for (int bx = 0; bx < c.blocks_w; bx++)
                    idct_islow(&c.coef[((size_t)by * c.blocks_w + bx) * 64], qt_[c.tq], &plane[((size_t)by * 8) * pstride + (size_t)bx * 8], (size_t)pstride);
            full.push_back(upsample(c, plane, pstride));
        }
        DecodedImage img;
        img.width = (uint32_t)width_;
        img.height = (uint32_t)height_;
        img.rgba.resize((size_t)width_ * height_ * 4);
        const size_t np = (size_t)width_ * height_;
        if (comp_.size() == 1) {
            for (size_t i = 0; i < np; i++) {
                img.rgba[4 * i] = img.rgba[4 * i + 1] = img.rgba[4 * i + 2] = full[0][i];
                img.rgba[4 * i + 3] = 255;
            }
        } else {
            // JFIF / no marker: YCbCr. Adobe marker: transform 0 = RGB stored directly, 1 = YCbCr. Component ids 'R','G','B' = RGB.
            bool ycc = true;
            if (adobe_)
                ycc = adobe_transform_ != 0;
            else if (comp_[0].id == 'R' && comp_[1].id == 'G' && comp_[2].id == 'B')
                ycc = false;
            for (size_t i = 0; i < np; i++) {
                const int y = full[0][i], cb = full[1][i] - 128, cr = full[2][i] - 128;
                if (ycc) {  // jdcolor.c build_ycc_rgb_table: SCALEBITS 16, ONE_HALF folded into the Cb->G table
                    const int r = y + (int)((91881LL * cr + 32768) >> 16);
                    const int g = y + (int)(((-22554LL * cb + 32768) + (-46802LL * cr)) >> 16);
                    const int b = y + (int)((116130LL * cb + 32768) >> 16);
                    img.rgba[4 * i] = clamp8(r), img.rgba[4 * i + 1] = clamp8(g), img.rgba[4 * i + 2] = clamp8(b);
                } else {
                    img.rgba[4 * i] = full[0][i], img.rgba[4 * i + 1] = full[1][i], img.rgba[4 * i + 2] = full[2][i];
                }
                img.rgba[4 * i + 3] = 255;
            }
        }
        return img;
    }
};

inline DecodedImage decode(const uint8_t *data, size_t size, const std::string &name) { return Decoder(data, size, name).decode(); }

}  // namespace jpeg
}  // namespace swr
