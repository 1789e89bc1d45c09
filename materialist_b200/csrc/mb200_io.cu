// mb200_io.cu — on-disk image formats of the path (host code, no CUDA): what `mi.Bitmap(path)` / `mi.util.write_bitmap`
// do in the reference (myutils/misc.py:99-111 BestSaver.save_results, myutils/mi_plugin.py:701-739 load_estimated_brdf,
// render_final.py:182-202, inverse_img_w_mi.py:54 'envmaps/0.hdr').  SURVEY §8f-4.
//
//   Radiance RGBE .hdr : read (flat + new-style RLE scanlines), write (RLE)
//   OpenEXR            : read scanline files, compression NONE / ZIPS / ZIP / PIZ (what Mitsuba writes: PIZ, FLOAT), pixel
//                        types HALF / FLOAT / UINT, channels R,G,B(,A) or a single channel (Y / any name);
//                        write scanline ZIP, FLOAT, increasing Y
// PIZ is restated from the published OpenEXR algorithm (ImfPizCompressor / ImfHuf / ImfWav: bitmap + LUT, canonical
// Huffman with 6-bit packed code lengths and a run-length symbol, 2-D Haar wavelet wdec14 / wdec16).  Parity: every
// reader is checked bit-exactly against OpenCV's decoder on the reference's shipped files (tests/test_image_io.py).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/materialist_b200.h"

namespace {

bool read_file(const char* path, std::vector<uint8_t>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    buf.resize((size_t)n);
    const size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}
bool ends_with(const char* s, const char* suf) {
    const size_t a = strlen(s), b = strlen(suf);
    if (a < b) return false;
    for (size_t i = 0; i < b; ++i) { char c = s[a - b + i]; if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a'); if (c != suf[i]) return false; }
    return true;
}

// ------------------------------------------------------------------------------------------------ Radiance RGBE
struct HdrInfo { int W = 0, H = 0; size_t data = 0; };
bool hdr_header(const std::vector<uint8_t>& b, HdrInfo& hi) {
    if (b.size() < 11 || (memcmp(b.data(), "#?RADIANCE", 10) != 0 && memcmp(b.data(), "#?RGBE", 6) != 0)) return false;
    size_t p = 0; bool blank = false;
    while (p < b.size()) {                         // header lines until the empty one
        size_t e = p; while (e < b.size() && b[e] != '\n') ++e;
        if (e == p) { blank = true; p = e + 1; break; }
        p = e + 1;
    }
    if (!blank) return false;
    size_t e = p; while (e < b.size() && b[e] != '\n') ++e;
    std::string res((const char*)b.data() + p, e - p);
    int h = 0, w = 0;
    if (sscanf(res.c_str(), "-Y %d +X %d", &h, &w) != 2 || h <= 0 || w <= 0) return false;   // the only orientation Mitsuba / OpenCV write
    hi.H = h; hi.W = w; hi.data = e + 1;
    return true;
}
inline void rgbe_to_float(const uint8_t* q, float* o) {
    if (q[3] == 0) { o[0] = o[1] = o[2] = 0.f; return; }
    const float f = ldexpf(1.0f, (int)q[3] - (128 + 8));
    o[0] = q[0] * f; o[1] = q[1] * f; o[2] = q[2] * f;
}
int hdr_read(const std::vector<uint8_t>& b, const HdrInfo& hi, float* out) {
    size_t p = hi.data; const int W = hi.W;
    std::vector<uint8_t> line((size_t)W * 4);
    for (int y = 0; y < hi.H; ++y) {
        if (p + 4 > b.size()) return MB200_EINVAL;
        const bool rle = W >= 8 && W < 32768 && b[p] == 2 && b[p + 1] == 2 && (((int)b[p + 2] << 8) | b[p + 3]) == W;
        if (!rle) {                                // flat pixels
            if (p + (size_t)W * 4 > b.size()) return MB200_EINVAL;
            memcpy(line.data(), b.data() + p, (size_t)W * 4); p += (size_t)W * 4;
        } else {
            p += 4;
            for (int c = 0; c < 4; ++c) {
                int x = 0;
                while (x < W) {
                    if (p >= b.size()) return MB200_EINVAL;
                    int n = b[p++];
                    if (n > 128) {                 // run
                        n -= 128;
                        if (n == 0 || x + n > W || p >= b.size()) return MB200_EINVAL;
                        const uint8_t v = b[p++];
                        for (int i = 0; i < n; ++i) line[(size_t)(x++) * 4 + c] = v;
                    } else {                       // literal
                        if (n == 0 || x + n > W || p + (size_t)n > b.size()) return MB200_EINVAL;
                        for (int i = 0; i < n; ++i) line[(size_t)(x++) * 4 + c] = b[p++];
                    }
                }
            }
        }
        for (int x = 0; x < W; ++x) rgbe_to_float(&line[(size_t)x * 4], out + ((size_t)y * W + x) * 3);
    }
    return MB200_OK;
}
inline void float_to_rgbe(const float* c, uint8_t* q) {
    float v = c[0]; if (c[1] > v) v = c[1]; if (c[2] > v) v = c[2];
    if (!(v >= 1e-32f)) { q[0] = q[1] = q[2] = q[3] = 0; return; }
    int e; const float s = frexpf(v, &e) * 256.0f / v;
    q[0] = (uint8_t)(c[0] > 0.f ? c[0] * s : 0.f); q[1] = (uint8_t)(c[1] > 0.f ? c[1] * s : 0.f); q[2] = (uint8_t)(c[2] > 0.f ? c[2] * s : 0.f);
    q[3] = (uint8_t)(e + 128);
}
// fwrite / fclose with the error remembered: a full disk must not come back as MB200_OK
struct OutFile {
    FILE* f; bool ok;
    explicit OutFile(const char* path) : f(fopen(path, "wb")), ok(f != nullptr) {}
    void put(const void* p, size_t n) { if (ok && n && fwrite(p, 1, n, f) != n) ok = false; }
    int close() { if (f && fclose(f) != 0) ok = false; f = nullptr; return ok ? MB200_OK : MB200_EIO; }
    ~OutFile() { if (f) fclose(f); }
};
int hdr_write(const char* path, const float* img, int H, int W) {
    OutFile of(path);
    if (!of.f) return MB200_EINVAL;
    char head[128]; const int hl = snprintf(head, sizeof(head), "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", H, W);
    of.put(head, (size_t)hl);
    std::vector<uint8_t> px((size_t)W * 4), out;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) float_to_rgbe(img + ((size_t)y * W + x) * 3, &px[(size_t)x * 4]);
        if (W < 8 || W >= 32768) { of.put(px.data(), px.size()); continue; }
        out.clear();
        out.push_back(2); out.push_back(2); out.push_back((uint8_t)(W >> 8)); out.push_back((uint8_t)(W & 255));
        for (int c = 0; c < 4; ++c) {
            int x = 0;
            while (x < W) {
                int run = 1;                       // run length starting at x
                while (x + run < W && run < 127 && px[(size_t)(x + run) * 4 + c] == px[(size_t)x * 4 + c]) ++run;
                if (run >= 4) { out.push_back((uint8_t)(128 + run)); out.push_back(px[(size_t)x * 4 + c]); x += run; continue; }
                int lit = 0;                       // literal span until the next run of >= 4
                while (x + lit < W && lit < 128) {
                    int r = 1;
                    while (x + lit + r < W && r < 4 && px[(size_t)(x + lit + r) * 4 + c] == px[(size_t)(x + lit) * 4 + c]) ++r;
                    if (r >= 4) break;
                    ++lit;
                }
                if (lit == 0) lit = 1;
                out.push_back((uint8_t)lit);
                for (int i = 0; i < lit; ++i) out.push_back(px[(size_t)(x + i) * 4 + c]);
                x += lit;
            }
        }
        of.put(out.data(), out.size());
    }
    return of.close();
}

// ------------------------------------------------------------------------------------------------ OpenEXR
enum { EXR_UINT = 0, EXR_HALF = 1, EXR_FLOAT = 2 };
enum { EXR_NONE = 0, EXR_RLE = 1, EXR_ZIPS = 2, EXR_ZIP = 3, EXR_PIZ = 4 };
struct ExrChan { std::string name; int type; int xs, ys; };
struct ExrInfo {
    std::vector<ExrChan> ch; int comp = -1; int x0 = 0, y0 = 0, x1 = -1, y1 = -1; int line_order = 0; size_t table = 0;
    int W() const { return x1 - x0 + 1; } int H() const { return y1 - y0 + 1; }
    int lines_per_block() const { return comp == EXR_PIZ ? 32 : (comp == EXR_ZIP ? 16 : 1); }
};
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

bool exr_header(const std::vector<uint8_t>& b, ExrInfo& ei) {
    if (b.size() < 8 || rd32(b.data()) != 20000630u) return false;
    const uint32_t ver = rd32(b.data() + 4);
    if ((ver & 0xff) != 2 || (ver & 0x1a00)) return false;          // tiled / deep / multipart are not part of the path
    size_t p = 8;
    for (;;) {
        if (p >= b.size()) return false;
        if (b[p] == 0) { ++p; break; }
        size_t e = p; while (e < b.size() && b[e]) ++e;
        std::string name((const char*)b.data() + p, e - p); p = e + 1;
        e = p; while (e < b.size() && b[e]) ++e;
        std::string type((const char*)b.data() + p, e - p); p = e + 1;
        if (p + 4 > b.size()) return false;
        const uint32_t sz = rd32(b.data() + p); p += 4;
        if (p + sz > b.size()) return false;
        const uint8_t* v = b.data() + p;
        if (name == "channels" && type == "chlist") {
            size_t q = 0;
            while (q < sz && v[q]) {
                size_t e2 = q; while (e2 < sz && v[e2]) ++e2;
                ExrChan c; c.name.assign((const char*)v + q, e2 - q); q = e2 + 1;
                if (q + 16 > sz) return false;
                c.type = (int)rd32(v + q); c.xs = (int)rd32(v + q + 8); c.ys = (int)rd32(v + q + 12); q += 16;
                ei.ch.push_back(c);
            }
        } else if (name == "compression" && sz >= 1) ei.comp = v[0];
        else if (name == "dataWindow" && sz >= 16) { ei.x0 = (int)rd32(v); ei.y0 = (int)rd32(v + 4); ei.x1 = (int)rd32(v + 8); ei.y1 = (int)rd32(v + 12); }
        else if (name == "lineOrder" && sz >= 1) ei.line_order = v[0];
        p += sz;
    }
    ei.table = p;
    if (ei.ch.empty() || ei.comp < 0 || ei.W() <= 0 || ei.H() <= 0) return false;
    for (const ExrChan& c : ei.ch) if (c.xs != 1 || c.ys != 1 || c.type < 0 || c.type > 2) return false;
    return true;
}
inline float half_to_float(uint16_t h) {
    const uint32_t s = (uint32_t)(h >> 15) << 31; uint32_t e = (h >> 10) & 31, m = h & 1023; uint32_t u;
    if (e == 0) {
        if (m == 0) u = s;
        else { e = 127 - 15 + 1; while (!(m & 1024)) { m <<= 1; --e; } u = s | (e << 23) | ((m & 1023) << 13); }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 127 - 15) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}

// ---- ZIP: zlib + byte predictor + even/odd interleave
bool zip_undo(const uint8_t* src, size_t n_src, std::vector<uint8_t>& dst, size_t n_raw) {
    std::vector<uint8_t> tmp(n_raw);
    uLongf got = (uLongf)n_raw;
    if (uncompress(tmp.data(), &got, src, (uLong)n_src) != Z_OK || got != n_raw) return false;
    for (size_t i = 1; i < n_raw; ++i) tmp[i] = (uint8_t)(tmp[i - 1] + tmp[i] - 128);
    dst.resize(n_raw);
    const size_t half = (n_raw + 1) / 2;
    for (size_t i = 0, a = 0, c = half; i < n_raw;) { dst[i++] = tmp[a++]; if (i < n_raw) dst[i++] = tmp[c++]; }
    return true;
}
void zip_do(const uint8_t* raw, size_t n, std::vector<uint8_t>& out) {
    std::vector<uint8_t> tmp(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0, a = 0, c = half; i < n;) { tmp[a++] = raw[i++]; if (i < n) tmp[c++] = raw[i++]; }
    int prev = n ? tmp[0] : 0;
    for (size_t i = 1; i < n; ++i) { const int cur = tmp[i]; tmp[i] = (uint8_t)(cur - prev + (128 + 256)); prev = cur; }
    uLongf cap = compressBound((uLong)n); out.resize(cap);
    compress2(out.data(), &cap, tmp.data(), (uLong)n, Z_DEFAULT_COMPRESSION); out.resize(cap);
}

// ---- PIZ
struct BitReader {
    const uint8_t* p; const uint8_t* end; uint64_t c = 0; int lc = 0;
    BitReader(const uint8_t* a, const uint8_t* b) : p(a), end(b) {}
    inline uint32_t get(int n) { while (lc < n) { c = (c << 8) | (p < end ? *p : 0); ++p; lc += 8; } lc -= n; return (uint32_t)((c >> lc) & ((1ull << n) - 1)); }
};
constexpr int kHufEnc = (1 << 16) + 1;
bool huf_uncompress(const uint8_t* src, size_t n_src, uint16_t* out, size_t n_out) {
    if (n_src == 0) return n_out == 0;
    if (n_src < 20) return false;
    const uint32_t im = rd32(src), iM = rd32(src + 4), nBits = rd32(src + 12);
    if (im >= (uint32_t)kHufEnc || iM >= (uint32_t)kHufEnc || im > iM) return false;
    std::vector<uint8_t> len(kHufEnc, 0);
    const uint8_t* ptr = src + 20; const uint8_t* end = src + n_src;
    {   // hufUnpackEncTable: 6-bit code lengths, 59..62 = short zero runs (2..5), 63 = long zero run (8 more bits + 6)
        BitReader br(ptr, end);
        for (uint32_t s = im; s <= iM; ++s) {
            const uint32_t l = br.get(6);
            if (l == 63) { uint32_t z = br.get(8) + 6; if (s + z > iM + 1) return false; while (z--) len[s++] = 0; --s; }
            else if (l >= 59) { uint32_t z = l - 59 + 2; if (s + z > iM + 1) return false; while (z--) len[s++] = 0; --s; }
            else len[s] = (uint8_t)l;
        }
        ptr = br.p;                                 // the table ends on a byte boundary (unused bits dropped)
        if (ptr > end) return false;
    }
    // hufCanonicalCodeTable: shortest codes get the numerically LARGEST values; within a length codes follow symbol order
    uint64_t base[59]; uint32_t count[59] = {0}, first[59];
    for (int s = 0; s < kHufEnc; ++s) count[len[s]]++;
    { uint64_t c = 0; for (int l = 58; l > 0; --l) { const uint64_t nc = (c + count[l]) >> 1; base[l] = c; c = nc; } }
    std::vector<uint32_t> sorted; sorted.reserve(kHufEnc);
    { uint32_t off = 0; for (int l = 1; l <= 58; ++l) { first[l] = off; off += count[l]; } sorted.resize(off);
      uint32_t fill[59]; for (int l = 1; l <= 58; ++l) fill[l] = first[l];
      for (int s = 0; s < kHufEnc; ++s) if (len[s]) sorted[fill[len[s]]++] = (uint32_t)s; }
    // decode nBits bits; symbol iM is the run-length escape: next 8 bits = how many more copies of the previous value
    BitReader br(ptr, end);
    uint64_t bits_left = nBits; size_t o = 0;
    while (bits_left > 0 && o <= n_out) {
        uint64_t code = 0; int l = 0; int sym = -1;
        while (l < 58 && bits_left > 0) {
            code = (code << 1) | br.get(1); ++l; --bits_left;
            if (count[l] && code >= base[l] && code - base[l] < count[l]) { sym = (int)sorted[first[l] + (uint32_t)(code - base[l])]; break; }
        }
        if (sym < 0) break;                         // trailing pad bits
        if ((uint32_t)sym == iM) {
            if (bits_left < 8 || o == 0) return false;
            uint32_t n = br.get(8); bits_left -= 8;
            if (o + n > n_out) return false;
            const uint16_t v = out[o - 1];
            while (n--) out[o++] = v;
        } else {
            if (o >= n_out) return false;
            out[o++] = (uint16_t)sym;
        }
    }
    return o == n_out;
}
inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int16_t ls = (int16_t)l, hs = (int16_t)h;
    const int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
    a = (uint16_t)(int16_t)ai; b = (uint16_t)(int16_t)(ai - hi);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xffff, aa = (d + bb - 0x8000) & 0xffff;
    b = (uint16_t)bb; a = (uint16_t)aa;
}
void wav2_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const bool w14 = mx < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1; p2 = p; p >>= 1;
    while (p >= 1) {
        uint16_t* py = in; uint16_t* ey = in + (ptrdiff_t)oy * (ny - p2);
        const ptrdiff_t oy1 = (ptrdiff_t)oy * p, oy2 = (ptrdiff_t)oy * p2, ox1 = (ptrdiff_t)ox * p, ox2 = (ptrdiff_t)ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py; uint16_t* ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1; uint16_t* p10 = px + oy1; uint16_t* p11 = p10 + ox1;
                if (w14) { wdec14(*px, *p10, i00, i10); wdec14(*p01, *p11, i01, i11); wdec14(i00, i01, *px, *p01); wdec14(i10, i11, *p10, *p11); }
                else     { wdec16(*px, *p10, i00, i10); wdec16(*p01, *p11, i01, i11); wdec16(i00, i01, *px, *p01); wdec16(i10, i11, *p10, *p11); }
            }
            if (nx & p) {
                uint16_t* p10 = px + oy1;
                if (w14) wdec14(*px, *p10, i00, *p10); else wdec16(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {
            uint16_t* px = py; uint16_t* ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                if (w14) wdec14(*px, *p01, i00, *p01); else wdec16(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p; p >>= 1;
    }
}
bool piz_undo(const uint8_t* src, size_t n_src, const ExrInfo& ei, int lines, std::vector<uint8_t>& dst, size_t n_raw) {
    if (n_src == n_raw) { dst.assign(src, src + n_src); return true; }           // stored uncompressed
    if (n_src < 4) return false;
    const int W = ei.W();
    std::vector<uint8_t> bitmap(8192, 0);
    size_t p = 0;
    const uint16_t minNZ = (uint16_t)(src[0] | (src[1] << 8)), maxNZ = (uint16_t)(src[2] | (src[3] << 8)); p = 4;
    if (maxNZ >= 8192) return false;
    if (minNZ <= maxNZ) { const size_t n = (size_t)maxNZ - minNZ + 1; if (p + n > n_src) return false; memcpy(&bitmap[minNZ], src + p, n); p += n; }
    std::vector<uint16_t> lut(65536, 0);
    int k = 0;
    for (int i = 0; i < 65536; ++i) if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t)i;
    const uint16_t maxValue = (uint16_t)(k - 1);
    if (p + 4 > n_src) return false;
    const uint32_t length = rd32(src + p); p += 4;
    if (p + length > n_src) return false;
    const size_t n16 = n_raw / 2;
    std::vector<uint16_t> tmp(n16);
    if (!huf_uncompress(src + p, length, tmp.data(), n16)) return false;
    // channel-major planes -> wavelet decode each 16-bit plane of each channel
    std::vector<size_t> start(ei.ch.size()); size_t off = 0;
    for (size_t c = 0; c < ei.ch.size(); ++c) {
        const int size = ei.ch[c].type == EXR_HALF ? 1 : 2;
        start[c] = off;
        for (int j = 0; j < size; ++j) wav2_decode(tmp.data() + off + j, W, size, lines, W * size, maxValue);
        off += (size_t)W * lines * size;
    }
    for (size_t i = 0; i < n16; ++i) tmp[i] = lut[tmp[i]];
    // back to scanline order: for each line, each channel's W * size values
    dst.resize(n_raw);
    uint16_t* o = reinterpret_cast<uint16_t*>(dst.data());
    std::vector<size_t> cur = start;
    for (int y = 0; y < lines; ++y)
        for (size_t c = 0; c < ei.ch.size(); ++c) {
            const size_t n = (size_t)W * (ei.ch[c].type == EXR_HALF ? 1 : 2);
            memcpy(o, tmp.data() + cur[c], n * 2); o += n; cur[c] += n;
        }
    return true;
}

// which file channel feeds output channel k (R,G,B,A order; a single-channel file feeds channel 0 whatever its name)
int exr_out_channels(const ExrInfo& ei, int map[4]) {
    if (ei.ch.size() == 1) { map[0] = 0; return 1; }
    const char* want[4] = {"R", "G", "B", "A"}; int n = 0;
    for (int k = 0; k < 4; ++k) {
        map[k] = -1;
        for (size_t c = 0; c < ei.ch.size(); ++c) if (ei.ch[c].name == want[k]) map[k] = (int)c;
        if (map[k] < 0) break;
        n = k + 1;
    }
    return n >= 3 ? n : 0;
}
int exr_read(const std::vector<uint8_t>& b, const ExrInfo& ei, float* out, int C) {
    int map[4]; const int nc = exr_out_channels(ei, map);
    if (nc != C) return MB200_EINVAL;
    const int W = ei.W(), H = ei.H(), lpb = ei.lines_per_block(), nblocks = (H + lpb - 1) / lpb;
    if (ei.comp != EXR_NONE && ei.comp != EXR_ZIPS && ei.comp != EXR_ZIP && ei.comp != EXR_PIZ) return MB200_EUNSUPPORTED;
    size_t line_bytes = 0; std::vector<size_t> choff(ei.ch.size());
    for (size_t c = 0; c < ei.ch.size(); ++c) { choff[c] = line_bytes; line_bytes += (size_t)W * (ei.ch[c].type == EXR_HALF ? 2 : 4); }
    if (ei.table + (size_t)nblocks * 8 > b.size()) return MB200_EINVAL;
    std::vector<uint8_t> raw;
    for (int blk = 0; blk < nblocks; ++blk) {
        const uint64_t off = rd64(b.data() + ei.table + (size_t)blk * 8);
        if (off + 8 > b.size()) return MB200_EINVAL;
        const int y = (int)rd32(b.data() + off); const uint32_t sz = rd32(b.data() + off + 4);
        if (off + 8 + sz > b.size() || y < ei.y0 || y > ei.y1) return MB200_EINVAL;
        const int lines = std::min(lpb, ei.y1 - y + 1);
        const size_t n_raw = line_bytes * lines;
        const uint8_t* src = b.data() + off + 8;
        if (ei.comp == EXR_NONE || sz == n_raw) raw.assign(src, src + sz);
        else if (ei.comp == EXR_PIZ) { if (!piz_undo(src, sz, ei, lines, raw, n_raw)) return MB200_EINVAL; }
        else if (!zip_undo(src, sz, raw, n_raw)) return MB200_EINVAL;
        if (raw.size() != n_raw) return MB200_EINVAL;
        for (int l = 0; l < lines; ++l) {
            const uint8_t* L = raw.data() + line_bytes * l;
            float* o = out + (size_t)(y - ei.y0 + l) * W * C;
            for (int k = 0; k < C; ++k) {
                const ExrChan& ch = ei.ch[map[k]]; const uint8_t* q = L + choff[map[k]];
                for (int x = 0; x < W; ++x) {
                    float v;
                    if (ch.type == EXR_FLOAT) memcpy(&v, q + 4 * (size_t)x, 4);
                    else if (ch.type == EXR_HALF) v = half_to_float((uint16_t)(q[2 * x] | (q[2 * x + 1] << 8)));
                    else v = (float)rd32(q + 4 * (size_t)x);
                    o[(size_t)x * C + k] = v;
                }
            }
        }
    }
    return MB200_OK;
}
void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
void put_attr(std::vector<uint8_t>& v, const char* name, const char* type, const std::vector<uint8_t>& val) {
    v.insert(v.end(), name, name + strlen(name) + 1); v.insert(v.end(), type, type + strlen(type) + 1);
    put32(v, (uint32_t)val.size()); v.insert(v.end(), val.begin(), val.end());
}
int exr_write(const char* path, const float* img, int H, int W, int C) {
    if (C != 1 && C != 3 && C != 4) return MB200_EINVAL;
    // channels are stored in alphabetical order: A, B, G, R (or Y)
    std::vector<std::pair<std::string, int>> chans;
    if (C == 1) chans = {{"Y", 0}}; else { if (C == 4) chans.push_back({"A", 3}); chans.push_back({"B", 2}); chans.push_back({"G", 1}); chans.push_back({"R", 0}); }
    std::vector<uint8_t> h; put32(h, 20000630u); put32(h, 2u);
    std::vector<uint8_t> v;
    for (auto& c : chans) { v.insert(v.end(), c.first.begin(), c.first.end()); v.push_back(0); put32(v, EXR_FLOAT); put32(v, 0); put32(v, 1); put32(v, 1); }
    v.push_back(0); put_attr(h, "channels", "chlist", v);
    put_attr(h, "compression", "compression", std::vector<uint8_t>{EXR_ZIP});
    v.clear(); put32(v, 0); put32(v, 0); put32(v, (uint32_t)(W - 1)); put32(v, (uint32_t)(H - 1));
    put_attr(h, "dataWindow", "box2i", v); put_attr(h, "displayWindow", "box2i", v);
    put_attr(h, "lineOrder", "lineOrder", std::vector<uint8_t>{0});
    const float one = 1.f, zero = 0.f; v.clear(); v.resize(4); memcpy(v.data(), &one, 4); put_attr(h, "pixelAspectRatio", "float", v);
    v.clear(); v.resize(8); memcpy(v.data(), &zero, 4); memcpy(v.data() + 4, &zero, 4); put_attr(h, "screenWindowCenter", "v2f", v);
    v.clear(); v.resize(4); memcpy(v.data(), &one, 4); put_attr(h, "screenWindowWidth", "float", v);
    h.push_back(0);
    const int lpb = 16, nblocks = (H + lpb - 1) / lpb;
    std::vector<std::vector<uint8_t>> blocks(nblocks);
    std::vector<uint8_t> raw;
    for (int blk = 0; blk < nblocks; ++blk) {
        const int y0 = blk * lpb, lines = std::min(lpb, H - y0);
        raw.resize((size_t)lines * W * C * 4); size_t o = 0;
        for (int l = 0; l < lines; ++l)
            for (auto& c : chans)
                for (int x = 0; x < W; ++x) { memcpy(&raw[o], img + ((size_t)(y0 + l) * W + x) * C + c.second, 4); o += 4; }
        zip_do(raw.data(), raw.size(), blocks[blk]);
        if (blocks[blk].size() >= raw.size()) blocks[blk] = raw;     // stored raw when compression does not help (reader: size == raw size)
    }
    OutFile of(path);
    if (!of.f) return MB200_EINVAL;
    of.put(h.data(), h.size());
    uint64_t off = h.size() + (uint64_t)nblocks * 8;
    for (int blk = 0; blk < nblocks; ++blk) { of.put(&off, 8); off += 8 + blocks[blk].size(); }
    for (int blk = 0; blk < nblocks; ++blk) {
        const int32_t y = blk * lpb, sz = (int32_t)blocks[blk].size();
        of.put(&y, 4); of.put(&sz, 4); of.put(blocks[blk].data(), blocks[blk].size());
    }
    return of.close();
}


// ------------------------------------------------------------------------------------------------ PNG (bg.png / mask.png / previews)
// read: 8 / 16-bit gray, gray+alpha, RGB, RGBA and 8-bit palette, non-interlaced -> float in [0,1] (value / 255 or / 65535, what
// plt.imread returns for the reference's mi_plugin.py:717-731); write: 8-bit gray / RGB / RGBA from floats clamped to [0,1].
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
struct PngInfo { int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0; int channels() const { return ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 3 : ctype == 4 ? 2 : 4; } };
bool png_header(const std::vector<uint8_t>& b, PngInfo& pi) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (b.size() < 33 || memcmp(b.data(), sig, 8) != 0 || memcmp(b.data() + 12, "IHDR", 4) != 0) return false;
    pi.W = (int)be32(b.data() + 16); pi.H = (int)be32(b.data() + 20); pi.depth = b[24]; pi.ctype = b[25]; pi.interlace = b[28];
    if (pi.W <= 0 || pi.H <= 0 || pi.interlace != 0) return false;
    if (!((pi.depth == 8) || (pi.depth == 16 && pi.ctype != 3))) return false;
    return pi.ctype == 0 || pi.ctype == 2 || pi.ctype == 3 || pi.ctype == 4 || pi.ctype == 6;
}
int png_read(const std::vector<uint8_t>& b, const PngInfo& pi, float* out) {
    std::vector<uint8_t> idat, plte;
    size_t p = 8;
    while (p + 12 <= b.size()) {
        const uint32_t len = be32(b.data() + p);
        if (p + 12 + len > b.size()) return MB200_EINVAL;
        if (!memcmp(b.data() + p + 4, "IDAT", 4)) idat.insert(idat.end(), b.data() + p + 8, b.data() + p + 8 + len);
        else if (!memcmp(b.data() + p + 4, "PLTE", 4)) plte.assign(b.data() + p + 8, b.data() + p + 8 + len);
        else if (!memcmp(b.data() + p + 4, "IEND", 4)) break;
        p += 12 + len;
    }
    const int spp = pi.ctype == 3 ? 1 : pi.channels(), bps = pi.depth / 8, bpp = spp * bps;
    const size_t stride = (size_t)pi.W * bpp;
    std::vector<uint8_t> raw((stride + 1) * pi.H);
    uLongf got = (uLongf)raw.size();
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != raw.size()) return MB200_EINVAL;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    const int C = pi.channels();
    for (int y = 0; y < pi.H; ++y) {
        const uint8_t* src = raw.data() + (stride + 1) * y; const int ft = src[0]; ++src;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, bb = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
            int v = src[i];
            if (ft == 1) v += a; else if (ft == 2) v += bb; else if (ft == 3) v += (a + bb) >> 1;
            else if (ft == 4) { const int pp = a + bb - c, pa = abs(pp - a), pb = abs(pp - bb), pc = abs(pp - c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c); }
            else if (ft != 0) return MB200_EINVAL;
            cur[i] = (uint8_t)v;
        }
        float* o = out + (size_t)y * pi.W * C;
        for (int x = 0; x < pi.W; ++x) {
            if (pi.ctype == 3) {
                const size_t k = (size_t)cur[x] * 3;
                for (int c = 0; c < 3; ++c) o[x * 3 + c] = (k + c < plte.size() ? plte[k + c] : 0) / 255.0f;
            } else
                for (int c = 0; c < C; ++c) {
                    const uint8_t* q = &cur[(size_t)(x * spp + c) * bps];
                    o[x * C + c] = bps == 1 ? q[0] / 255.0f : (float)((q[0] << 8) | q[1]) / 65535.0f;
                }
        }
        prev.swap(cur);
    }
    return MB200_OK;
}
void png_chunk(OutFile& of, const char* type, const uint8_t* data, uint32_t len) {
    uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
    of.put(hdr, 8); of.put(data, len);
    uLong crc = crc32(0L, hdr + 4, 4); if (len) crc = crc32(crc, data, len);
    const uint8_t c4[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    of.put(c4, 4);
}
// srgb: colour channels through the sRGB transfer curve, as Mitsuba's Bitmap::convert(UInt8, srgb_gamma = true) does behind
// mi.util.write_bitmap for 8-bit files (trans_edit.py:47, the preview PNGs); alpha stays linear.  Raw (srgb = false) for data files
// such as bg.png / mask.png fixtures whose values must come back unchanged.
inline float srgb_oetf(float x) { return x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f; }
int png_write(const char* path, const float* img, int H, int W, int C, bool srgb) {
    if (C != 1 && C != 3 && C != 4) return MB200_EINVAL;
    std::vector<uint8_t> raw((size_t)H * ((size_t)W * C + 1));
    for (int y = 0; y < H; ++y) {
        uint8_t* row = raw.data() + (size_t)y * ((size_t)W * C + 1); row[0] = 0;
        for (size_t i = 0; i < (size_t)W * C; ++i) {
            float v = img[(size_t)y * W * C + i]; v = v != v ? 0.f : (v < 0.f ? 0.f : (v > 1.f ? 1.f : v));
            if (srgb && !(C == 4 && i % 4 == 3)) v = srgb_oetf(v);
            row[1 + i] = (uint8_t)(v * 255.0f + 0.5f);
        }
    }
    uLongf cap = compressBound((uLong)raw.size()); std::vector<uint8_t> z(cap);
    if (compress2(z.data(), &cap, raw.data(), (uLong)raw.size(), Z_DEFAULT_COMPRESSION) != Z_OK) return MB200_EINVAL;
    OutFile of(path);
    if (!of.f) return MB200_EINVAL;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    of.put(sig, 8);
    const uint8_t ihdr[13] = {(uint8_t)(W >> 24), (uint8_t)(W >> 16), (uint8_t)(W >> 8), (uint8_t)W, (uint8_t)(H >> 24), (uint8_t)(H >> 16), (uint8_t)(H >> 8), (uint8_t)H,
                              8, (uint8_t)(C == 1 ? 0 : C == 3 ? 2 : 6), 0, 0, 0};
    png_chunk(of, "IHDR", ihdr, 13); png_chunk(of, "IDAT", z.data(), (uint32_t)cap); png_chunk(of, "IEND", nullptr, 0);
    return of.close();
}

}  // namespace

extern "C" {

int mb200_image_info(const char* path, int* H, int* W, int* C) {
    if (!path || !H || !W || !C) return MB200_EINVAL;
    std::vector<uint8_t> b;
    if (!read_file(path, b)) return MB200_EINVAL;
    HdrInfo hi; ExrInfo ei; PngInfo pi;
    if (png_header(b, pi)) { *H = pi.H; *W = pi.W; *C = pi.channels(); return MB200_OK; }
    if (hdr_header(b, hi)) { *H = hi.H; *W = hi.W; *C = 3; return MB200_OK; }
    if (exr_header(b, ei)) { int map[4]; const int nc = exr_out_channels(ei, map); if (!nc) return MB200_EUNSUPPORTED; *H = ei.H(); *W = ei.W(); *C = nc; return MB200_OK; }
    return MB200_EUNSUPPORTED;
}

int mb200_image_read(const char* path, float* out, int H, int W, int C) {
    if (!path || !out) return MB200_EINVAL;
    std::vector<uint8_t> b;
    if (!read_file(path, b)) return MB200_EINVAL;
    HdrInfo hi; ExrInfo ei; PngInfo pi;
    if (png_header(b, pi)) { if (pi.H != H || pi.W != W || pi.channels() != C) return MB200_EINVAL; return png_read(b, pi, out); }
    if (hdr_header(b, hi)) { if (hi.H != H || hi.W != W || C != 3) return MB200_EINVAL; return hdr_read(b, hi, out); }
    if (exr_header(b, ei)) { if (ei.H() != H || ei.W() != W) return MB200_EINVAL; return exr_read(b, ei, out, C); }
    return MB200_EUNSUPPORTED;
}

int mb200_image_write_ex(const char* path, const float* img, int H, int W, int C, int flags) {
    if (!path || !img || H <= 0 || W <= 0) return MB200_EINVAL;
    if (ends_with(path, ".hdr")) return C == 3 ? hdr_write(path, img, H, W) : MB200_EINVAL;
    if (ends_with(path, ".exr")) return exr_write(path, img, H, W, C);
    if (ends_with(path, ".png")) return png_write(path, img, H, W, C, (flags & MB200_IMG_SRGB) != 0);
    return MB200_EUNSUPPORTED;
}
int mb200_image_write(const char* path, const float* img, int H, int W, int C) { return mb200_image_write_ex(path, img, H, W, C, 0); }

}  // extern "C"
