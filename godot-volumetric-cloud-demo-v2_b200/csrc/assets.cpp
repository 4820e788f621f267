// Host-side asset pipeline: what Godot's importer does for the three input bitmaps before they
// reach clouds.glsl (cloud_sky.gd:311,321,331; *.import:24-27), minus the lossy BC7 step.
//
//   * TGA: image types 2 (raw true-colour) and 10 (RLE true-colour), 24 or 32 bpp, either origin.
//     cloud_sky/perlworlnoise.tga is type 10, 32 bpp, bottom-left origin, BGRA on disk.
//   * BMP: BITMAPINFOHEADER-family, BI_RGB, 24 or 32 bpp, bottom-up or top-down, rows padded to
//     4 bytes, BGR(A) on disk (cloud_sky/worlnoise.bmp, weather.bmp are 24 bpp bottom-up).
//   * strip -> volume slicing ("slices/horizontal = N, slices/vertical = 1").
//   * 2x2x2 box-filter mip chain re-quantised to 8 bits ("mipmaps/generate = true").
#include <cstdio>
#include <cstring>

#include "cs_internal.h"

namespace cs {

namespace {

bool read_file(const char* path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    out.resize((size_t)n);
    size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

inline uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t rd32(const uint8_t* p) { return rd16(p) | (rd16(p + 2) << 16); }

std::string decode_tga(const std::vector<uint8_t>& f, HostImage& out) {
    if (f.size() < 18) return "TGA: file too short";
    const uint8_t* h = f.data();
    int id_len = h[0], cmap_type = h[1], type = h[2];
    int cmap_len = (int)rd16(h + 5), cmap_bits = h[7];
    int w = (int)rd16(h + 12), hgt = (int)rd16(h + 14), bpp = h[16], desc = h[17];
    if (type != 2 && type != 10) return "TGA: only true-colour types 2 and 10 are supported";
    if (bpp != 24 && bpp != 32) return "TGA: only 24/32 bpp supported";
    if (w < 1 || hgt < 1) return "TGA: bad dimensions";
    size_t pos = 18 + (size_t)id_len + (cmap_type ? (size_t)cmap_len * ((cmap_bits + 7) / 8) : 0);
    int bytes = bpp / 8;
    size_t npx = (size_t)w * hgt;
    std::vector<uint8_t> raw(npx * bytes);  // file order
    if (type == 2) {
        if (pos + raw.size() > f.size()) return "TGA: truncated pixel data";
        memcpy(raw.data(), f.data() + pos, raw.size());
    } else {
        size_t i = 0;
        while (i < npx) {
            if (pos >= f.size()) return "TGA: truncated RLE stream";
            int c = f[pos++];
            size_t count = (size_t)(c & 0x7f) + 1;
            if (i + count > npx) return "TGA: RLE packet overruns the image";
            if (c & 0x80) {
                if (pos + bytes > f.size()) return "TGA: truncated RLE packet";
                for (size_t k = 0; k < count; k++) memcpy(&raw[(i + k) * bytes], &f[pos], bytes);
                pos += bytes;
            } else {
                if (pos + count * bytes > f.size()) return "TGA: truncated raw packet";
                memcpy(&raw[i * bytes], &f[pos], count * bytes);
                pos += count * bytes;
            }
            i += count;
        }
    }
    bool top_origin = (desc & 0x20) != 0, right_origin = (desc & 0x10) != 0;
    out.w = w; out.h = hgt; out.ch = bytes;
    out.px.resize(npx * bytes);
    for (int y = 0; y < hgt; y++) {
        int sy = top_origin ? y : hgt - 1 - y;
        for (int x = 0; x < w; x++) {
            int sx = right_origin ? w - 1 - x : x;
            const uint8_t* s = &raw[((size_t)sy * w + sx) * bytes];
            uint8_t* d = &out.px[((size_t)y * w + x) * bytes];
            d[0] = s[2]; d[1] = s[1]; d[2] = s[0];  // BGR(A) -> RGB(A)
            if (bytes == 4) d[3] = s[3];
        }
    }
    return "";
}

std::string decode_bmp(const std::vector<uint8_t>& f, HostImage& out) {
    if (f.size() < 54 || f[0] != 'B' || f[1] != 'M') return "BMP: bad signature";
    uint32_t data_off = rd32(&f[10]), hdr = rd32(&f[14]);
    if (hdr < 40) return "BMP: unsupported (pre-BITMAPINFOHEADER) header";
    int32_t w = (int32_t)rd32(&f[18]), h = (int32_t)rd32(&f[22]);
    uint32_t bpp = rd16(&f[28]), comp = rd32(&f[30]);
    if (comp != 0) return "BMP: only BI_RGB is supported";
    if (bpp != 24 && bpp != 32) return "BMP: only 24/32 bpp supported";
    bool top_down = h < 0;
    if (top_down) h = -h;
    if (w < 1 || h < 1) return "BMP: bad dimensions";
    int bytes = (int)bpp / 8;
    size_t stride = (((size_t)w * bytes) + 3) & ~(size_t)3;
    if ((size_t)data_off + stride * (size_t)h > f.size()) return "BMP: truncated pixel data";
    out.w = w; out.h = h; out.ch = 3;  // the 4th byte of 32-bpp BI_RGB is padding, not alpha
    out.px.resize((size_t)w * h * 3);
    for (int y = 0; y < h; y++) {
        const uint8_t* row = &f[data_off + stride * (size_t)(top_down ? y : h - 1 - y)];
        for (int x = 0; x < w; x++) {
            uint8_t* d = &out.px[((size_t)y * w + x) * 3];
            d[0] = row[x * bytes + 2]; d[1] = row[x * bytes + 1]; d[2] = row[x * bytes + 0];
        }
    }
    return "";
}

}  // namespace

std::string decode_image_file(const char* path, HostImage& out) {
    std::vector<uint8_t> f;
    if (!path || !read_file(path, f)) return std::string("cannot read ") + (path ? path : "(null)");
    if (f.size() >= 2 && f[0] == 'B' && f[1] == 'M') return decode_bmp(f, out);
    return decode_tga(f, out);
}

std::string strip_to_volume_rgba(const HostImage& s, int slices, std::vector<uint8_t>& out, int& n) {
    if (slices < 1 || s.w % slices != 0) return "strip width is not a multiple of the slice count";
    n = s.w / slices;
    if (s.h != n || slices != n) return "strip does not slice into a cube (need width = n*n, height = n, n slices)";
    out.resize((size_t)n * n * n * 4);
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                const uint8_t* src = &s.px[((size_t)y * s.w + (size_t)z * n + x) * s.ch];
                uint8_t* d = &out[(((size_t)z * n + y) * n + x) * 4];
                d[0] = src[0]; d[1] = src[1]; d[2] = src[2]; d[3] = s.ch == 4 ? src[3] : 255;
            }
    return "";
}

void expand_rgba(const uint8_t* src, size_t texels, int ch, std::vector<uint8_t>& dst) {
    dst.resize(texels * 4);
    if (ch == 4) { memcpy(dst.data(), src, texels * 4); return; }
    for (size_t i = 0; i < texels; i++) {
        dst[i * 4 + 0] = src[i * 3 + 0]; dst[i * 4 + 1] = src[i * 3 + 1]; dst[i * 4 + 2] = src[i * 3 + 2]; dst[i * 4 + 3] = 255;
    }
}

void build_volume_mips(std::vector<std::vector<uint8_t>>& levels, int n) {
    levels.resize(1);
    while (n > 1) {
        int m = n >> 1;
        const uint8_t* src = levels.back().data();
        std::vector<uint8_t> dst((size_t)m * m * m * 4);
        const size_t row = (size_t)n * 4, slab = (size_t)n * n * 4;
        for (int z = 0; z < m; z++)
            for (int y = 0; y < m; y++) {
                const uint8_t* a = src + (size_t)(2 * z) * slab + (size_t)(2 * y) * row;
                uint8_t* d = &dst[(((size_t)z * m + y) * m) * 4];
                for (int x = 0; x < m; x++, a += 8, d += 4)
                    for (int c = 0; c < 4; c++) {
                        unsigned s = a[c] + a[4 + c] + a[row + c] + a[row + 4 + c] + a[slab + c] + a[slab + 4 + c] +
                                     a[slab + row + c] + a[slab + row + 4 + c];
                        d[c] = (uint8_t)((s + 4u) >> 3);
                    }
            }
        levels.push_back(std::move(dst));
        n = m;
    }
}

}  // namespace cs
