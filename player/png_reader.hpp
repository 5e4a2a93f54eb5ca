// Minimal image file reader for the headless player: 8/16-bit grayscale, gray+alpha, RGB and RGBA
// non-interlaced PNG (inflate through zlib, scan-line filters undone here) and binary PGM (P5).
// The reference's readers call cv::imread (xrslam-pc/player/src/IO/euroc_dataset_reader.cpp:60,
// tum_dataset_reader.cpp:60); colour images are reduced to gray with OpenCV's BGR2GRAY fixed-point
// weights like the reader's cv::cvtColor (euroc_dataset_reader.cpp:76-77).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>

namespace xplayer {

struct GrayImage {
    int width = 0, height = 0;
    std::vector<uint8_t> data;          // row-major, stride == width
    bool empty() const { return data.empty(); }
};

inline uint8_t rgb_to_gray(int r, int g, int b) {      // cv::cvtColor 8-bit (OpenCV 4): (R*9798 + G*19235 + B*3735 + 2^14) >> 15
    return (uint8_t)((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15);
}

inline bool read_file(const std::string &path, std::vector<uint8_t> &buf) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return false; }
    buf.resize((size_t)n);
    const size_t got = n ? std::fread(buf.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

inline bool decode_pgm(const std::vector<uint8_t> &buf, GrayImage &img, std::string &err) {
    size_t p = 2;
    int vals[3], got = 0;
    while (got < 3 && p < buf.size()) {
        if (buf[p] == '#') { while (p < buf.size() && buf[p] != '\n') ++p; continue; }
        if (isspace(buf[p])) { ++p; continue; }
        int v = 0; bool any = false;
        while (p < buf.size() && isdigit(buf[p])) { v = v * 10 + (buf[p] - '0'); ++p; any = true; }
        if (!any) { err = "bad PGM header"; return false; }
        vals[got++] = v;
    }
    if (got != 3 || p >= buf.size()) { err = "bad PGM header"; return false; }
    ++p;                                  // the single whitespace after maxval
    const int w = vals[0], h = vals[1], maxv = vals[2];
    const size_t bpp = maxv > 255 ? 2 : 1;
    if (w <= 0 || h <= 0 || buf.size() - p < (size_t)w * h * bpp) { err = "truncated PGM"; return false; }
    img.width = w; img.height = h; img.data.resize((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) img.data[i] = bpp == 1 ? buf[p + i] : buf[p + 2 * i];     // 16-bit: high byte
    return true;
}

inline bool decode_png(const std::vector<uint8_t> &buf, GrayImage &img, std::string &err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) { err = "not a PNG"; return false; }
    auto be32 = [&](size_t o) { return ((uint32_t)buf[o] << 24) | ((uint32_t)buf[o + 1] << 16) | ((uint32_t)buf[o + 2] << 8) | buf[o + 3]; };
    uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, palette;
    size_t p = 8;
    bool seen_ihdr = false, seen_end = false;
    while (p + 12 <= buf.size()) {
        const uint32_t len = be32(p);
        const char *type = (const char *)&buf[p + 4];
        if (p + 12 + (size_t)len > buf.size()) { err = "truncated PNG chunk"; return false; }
        const uint8_t *d = &buf[p + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) { err = "bad IHDR"; return false; }
            w = be32(p + 8); h = be32(p + 12); depth = d[8]; ctype = d[9]; interlace = d[12]; seen_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(d, d + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!std::memcmp(type, "IEND", 4)) { seen_end = true; break; }
        p += 12 + (size_t)len;
    }
    if (!seen_ihdr || !seen_end || w == 0 || h == 0) { err = "incomplete PNG"; return false; }
    if (interlace) { err = "interlaced PNG is not supported"; return false; }
    int channels;
    switch (ctype) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break; case 4: channels = 2; break; case 6: channels = 4; break;
                     default: err = "bad PNG colour type"; return false; }
    if (!((depth == 8) || (depth == 16 && ctype != 3))) { err = "PNG bit depth other than 8 / 16 is not supported"; return false; }
    const size_t bpp = (size_t)channels * depth / 8, stride = (size_t)w * bpp;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) { err = "PNG inflate failed"; return false; }
    // undo the scan-line filters in place (PNG specification, section 9)
    std::vector<uint8_t> prev(stride, 0);
    img.width = (int)w; img.height = (int)h; img.data.resize((size_t)w * h);
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t *line = &raw[(stride + 1) * y + 1];
        const int ft = raw[(stride + 1) * y];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? line[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred = 0;
            switch (ft) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = b; break;
            case 3: pred = (a + b) >> 1; break;
            case 4: { const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                      pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
            default: err = "bad PNG filter"; return false;
            }
            line[i] = (uint8_t)(line[i] + pred);
        }
        std::memcpy(prev.data(), line, stride);
        const size_t step = depth / 8;              // 16-bit samples: the high byte (imread(..., IMREAD_UNCHANGED) would keep 16 bits)
        for (uint32_t x = 0; x < w; ++x) {
            const uint8_t *px = line + (size_t)x * bpp;
            uint8_t g;
            if (ctype == 0 || ctype == 4) g = px[0];
            else if (ctype == 3) { const size_t k = (size_t)px[0] * 3; g = k + 2 < palette.size() ? rgb_to_gray(palette[k], palette[k + 1], palette[k + 2]) : 0; }
            else g = rgb_to_gray(px[0], px[step], px[2 * step]);
            img.data[(size_t)y * w + x] = g;
        }
    }
    return true;
}

inline bool read_gray_image(const std::string &path, GrayImage &img, std::string &err) {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf)) { err = "cannot read " + path; return false; }
    if (buf.size() >= 2 && buf[0] == 'P' && buf[1] == '5') return decode_pgm(buf, img, err);
    return decode_png(buf, img, err);
}

}  // namespace xplayer
