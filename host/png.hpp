// png.hpp — minimal PNG reader for the host driver (stands in for cv::imread, pose_functions.cpp:526,548,597).
// 8-bit non-interlaced PNGs: gray, gray+alpha, RGB, RGBA, palette.  Output layout follows cv::imread:
// IMREAD_GRAYSCALE -> 1 channel, default -> 3 channels BGR.  Needs zlib.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace host {

struct Image {
    int rows = 0, cols = 0, channels = 0;
    std::vector<uint8_t> data;   // row-major, rows x cols x channels, no padding
    bool empty() const { return data.empty(); }
    size_t step() const { return (size_t)cols * channels; }
};

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// returns an empty image on any failure (like cv::imread)
inline Image read_png(const std::string& path, bool grayscale) {
    Image out;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return out;
    std::vector<uint8_t> buf;
    {
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n <= 8) { fclose(f); return out; }
        buf.resize((size_t)n);
        if (fread(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); return out; }
        fclose(f);
    }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (memcmp(buf.data(), sig, 8) != 0) return out;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    for (size_t at = 8; at + 12 <= buf.size();) {
        const uint32_t len = be32(&buf[at]);
        const char* type = (const char*)&buf[at + 4];
        const uint8_t* d = &buf[at + 8];
        if (at + 12 + len > buf.size()) return out;
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(d); h = be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(d, d + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        at += 12 + len;
    }
    if (!w || !h || depth != 8 || interlace != 0) return out;
    int spp;  // samples per pixel in the file
    switch (ctype) {
        case 0: spp = 1; break;
        case 2: spp = 3; break;
        case 3: spp = 1; break;
        case 4: spp = 2; break;
        case 6: spp = 4; break;
        default: return out;
    }
    const size_t stride = (size_t)w * spp;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) return out;
    // undo the scanline filters in place
    std::vector<uint8_t> pix(stride * h);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t ft = raw[y * (stride + 1)];
        const uint8_t* s = &raw[y * (stride + 1) + 1];
        uint8_t* d = &pix[y * stride];
        const uint8_t* up = y ? &pix[(y - 1) * stride] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)spp ? d[x - spp] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)spp) ? up[x - spp] : 0;
            int v = s[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: return out;
            }
            d[x] = (uint8_t)v;
        }
    }
    out.rows = (int)h; out.cols = (int)w; out.channels = grayscale ? 1 : 3;
    out.data.resize((size_t)h * w * out.channels);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        uint8_t r, g, b;
        const uint8_t* p = &pix[i * spp];
        if (ctype == 0 || ctype == 4) r = g = b = p[0];
        else if (ctype == 3) {
            const size_t k = (size_t)p[0] * 3;
            if (k + 2 < plte.size()) { r = plte[k]; g = plte[k + 1]; b = plte[k + 2]; }
            else r = g = b = 0;
        }
        else { r = p[0]; g = p[1]; b = p[2]; }
        if (grayscale) {
            // cv::cvtColor BGR2GRAY fixed point: (R*4899 + G*9617 + B*1868 + 8192) >> 14
            out.data[i] = (ctype == 0 || ctype == 4) ? r : (uint8_t)((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14);
        } else {
            out.data[i * 3] = b; out.data[i * 3 + 1] = g; out.data[i * 3 + 2] = r;
        }
    }
    return out;
}

}  // namespace host
