// kc_png.cu — PNG decode/encode on the host side of the boundary (zlib only).
//
// The reference reaches its codecs through the `image` crate:
//   read_slot_image  src/shared.rs:218-261   image::open(path) -> as_flat_samples_u8 -> planes
//   Image node       src/node/image.rs:10-26 any read error => 1x1 (1,0,1,1)
//   Write node       src/node/write.rs:5-21  image::save_buffer(path, to_u8(), Rgba8)
// Here the 8-bit PNG subset those call sites can produce a SlotImage from is decoded in C++
// to the same interleaved u8 samples `as_flat_samples_u8` yields (gray 1/2/4/8 bit -> L8,
// gray+alpha -> LA8, RGB -> RGB8, RGBA -> RGBA8, palette -> RGB8 / RGBA8 with tRNS, tRNS on
// gray/RGB adds the alpha channel; Adam7 interlacing handled), and RGBA8 is encoded for the
// Write node.  16-bit files are rejected (the reference's `.unwrap()` on as_flat_samples_u8
// panics on them).  No device work happens in this file.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <new>
#include <stdexcept>
#include <vector>

#include "kc_internal.h"

namespace {

const uint8_t kSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undo the per-scanline filters of one (sub)image in place; `raw` holds h * (1 + stride) bytes
bool unfilter(uint8_t* raw, size_t h, size_t stride, size_t bpp) {
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t* prev = zero.data();
    for (size_t y = 0; y < h; ++y) {
        uint8_t* line = raw + y * (stride + 1);
        const uint8_t ft = line[0];
        uint8_t* cur = line + 1;
        switch (ft) {
            case 0: break;
            case 1: for (size_t i = bpp; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
            case 2: for (size_t i = 0; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + prev[i]); break;
            case 3:
                for (size_t i = 0; i < stride; ++i) {
                    const int a = i >= bpp ? cur[i - bpp] : 0;
                    cur[i] = (uint8_t)(cur[i] + ((a + prev[i]) >> 1));
                }
                break;
            case 4:
                for (size_t i = 0; i < stride; ++i) {
                    const int a = i >= bpp ? cur[i - bpp] : 0, c = i >= bpp ? prev[i - bpp] : 0;
                    cur[i] = (uint8_t)(cur[i] + paeth(a, prev[i], c));
                }
                break;
            default: return false;
        }
        prev = cur;
    }
    return true;
}

constexpr uint32_t kMaxPngSide = 1u << 16;
constexpr uint64_t kMaxPngBytes = 1ull << 30;

struct Header {
    uint32_t w = 0, h = 0;
    int depth = 0, color = 0, interlace = 0;
    int samples_per_px() const { return color == 0 ? 1 : color == 2 ? 3 : color == 3 ? 1 : color == 4 ? 2 : 4; }
};

int32_t decode_impl(const uint8_t* data, size_t n, std::vector<uint8_t>& out, uint32_t& W, uint32_t& H, uint32_t& CH) {
    if (n < 8 || memcmp(data, kSig, 8) != 0) KC_FAIL(KC_ERR_IMAGE, "not a PNG file");
    Header hd;
    bool have_hdr = false, have_trns = false;
    std::vector<uint8_t> idat, plte, trns;
    size_t p = 8;
    while (p + 12 <= n) {
        const uint32_t len = be32(data + p);
        const uint8_t* type = data + p + 4;
        if (p + 12 + (size_t)len > n) KC_FAIL(KC_ERR_IMAGE, "truncated PNG chunk");
        const uint8_t* body = data + p + 8;
        if (be32(body + len) != (uint32_t)crc32(crc32(0, type, 4), body, len)) KC_FAIL(KC_ERR_IMAGE, "PNG chunk CRC mismatch");
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) KC_FAIL(KC_ERR_IMAGE, "bad IHDR");
            hd.w = be32(body); hd.h = be32(body + 4);
            hd.depth = body[8]; hd.color = body[9]; hd.interlace = body[12];
            if (body[10] != 0 || body[11] != 0 || hd.interlace > 1) KC_FAIL(KC_ERR_IMAGE, "unsupported PNG method");
            have_hdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!memcmp(type, "tRNS", 4)) {
            trns.assign(body, body + len);
            have_trns = true;
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        p += 12 + (size_t)len;
    }
    if (!have_hdr || hd.w == 0 || hd.h == 0) KC_FAIL(KC_ERR_IMAGE, "PNG without a valid IHDR");
    // The header comes from outside: bound it BEFORE any arithmetic on it.  With both sides <= 2^16 and <= 8 bytes
    // per pixel every product below fits 64 bits; the byte cap (a 16384^2 RGBA8 image) also bounds what a
    // maximally compressed stream may ask us to allocate.
    if (hd.w > kMaxPngSide || hd.h > kMaxPngSide) KC_FAIL(KC_ERR_IMAGE, "PNG is %u x %u: sides above %u are refused", hd.w, hd.h, kMaxPngSide);
    if (hd.depth == 16) KC_FAIL(KC_ERR_IMAGE, "16-bit PNG: the reference only takes 8-bit samples (as_flat_samples_u8)");
    const bool depth_ok = (hd.color == 0 && (hd.depth == 1 || hd.depth == 2 || hd.depth == 4 || hd.depth == 8)) ||
                          (hd.color == 3 && (hd.depth == 1 || hd.depth == 2 || hd.depth == 4 || hd.depth == 8)) ||
                          ((hd.color == 2 || hd.color == 4 || hd.color == 6) && hd.depth == 8);
    if (!depth_ok) KC_FAIL(KC_ERR_IMAGE, "unsupported PNG colour type %d / bit depth %d", hd.color, hd.depth);
    if (hd.color == 3 && plte.size() < 3) KC_FAIL(KC_ERR_IMAGE, "palette PNG without PLTE");

    const size_t bits_px = (size_t)hd.samples_per_px() * hd.depth;
    const size_t bpp = bits_px >= 8 ? bits_px / 8 : 1;
    auto stride_of = [&](size_t w) { return (w * bits_px + 7) / 8; };
    // size of the filtered stream
    static const int ax[7] = {0, 4, 0, 2, 0, 1, 0}, ay[7] = {0, 0, 4, 0, 2, 0, 1}, adx[7] = {8, 8, 4, 4, 2, 2, 1}, ady[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t total = 0;
    if (!hd.interlace) total = (size_t)hd.h * (stride_of(hd.w) + 1);
    else
        for (int i = 0; i < 7; ++i) {
            const size_t pw = (hd.w > (uint32_t)ax[i]) ? (hd.w - ax[i] + adx[i] - 1) / adx[i] : 0;
            const size_t ph = (hd.h > (uint32_t)ay[i]) ? (hd.h - ay[i] + ady[i] - 1) / ady[i] : 0;
            if (pw && ph) total += ph * (stride_of(pw) + 1);
        }
    if (total > kMaxPngBytes || (uint64_t)hd.w * hd.h * 4 > kMaxPngBytes) KC_FAIL(KC_ERR_IMAGE, "PNG of %u x %u exceeds the decoder's %llu-byte limit", hd.w, hd.h, (unsigned long long)kMaxPngBytes);
    // deflate expands by at most 1032:1, so a header that promises more pixels than the IDAT
    // bytes could ever inflate to is rejected before anything of that size is allocated
    if (total > idat.size() * 1032 + 1024) KC_FAIL(KC_ERR_IMAGE, "PNG header promises more data than its IDAT chunks can hold");
    std::vector<uint8_t> raw(total);
    uLongf got = (uLongf)total;
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != total) KC_FAIL(KC_ERR_IMAGE, "PNG data does not inflate to the image size");

    // unfiltered samples at the file's bit depth, one packed row after another (no filter bytes)
    const size_t full_stride = stride_of(hd.w);
    std::vector<uint8_t> px((size_t)hd.h * full_stride, 0);
    if (!hd.interlace) {
        if (!unfilter(raw.data(), hd.h, full_stride, bpp)) KC_FAIL(KC_ERR_IMAGE, "bad PNG filter type");
        for (size_t y = 0; y < hd.h; ++y) memcpy(&px[y * full_stride], &raw[y * (full_stride + 1) + 1], full_stride);
    } else {
        size_t off = 0;
        for (int i = 0; i < 7; ++i) {
            const size_t pw = (hd.w > (uint32_t)ax[i]) ? (hd.w - ax[i] + adx[i] - 1) / adx[i] : 0;
            const size_t ph = (hd.h > (uint32_t)ay[i]) ? (hd.h - ay[i] + ady[i] - 1) / ady[i] : 0;
            if (!pw || !ph) continue;
            const size_t st = stride_of(pw);
            if (!unfilter(&raw[off], ph, st, bpp)) KC_FAIL(KC_ERR_IMAGE, "bad PNG filter type");
            for (size_t y = 0; y < ph; ++y) {
                const uint8_t* line = &raw[off + y * (st + 1) + 1];
                uint8_t* dst = &px[(ay[i] + y * ady[i]) * full_stride];
                for (size_t x = 0; x < pw; ++x) {
                    const size_t dx = ax[i] + x * adx[i];
                    if (bits_px >= 8) memcpy(dst + dx * bpp, line + x * bpp, bpp);
                    else {
                        const int d = hd.depth, per = 8 / d;
                        const int v = (line[x / per] >> (8 - d - (int)(x % per) * d)) & ((1 << d) - 1);
                        dst[dx / per] |= (uint8_t)(v << (8 - d - (int)(dx % per) * d));
                    }
                }
            }
            off += ph * (st + 1);
        }
    }

    // expand to 8-bit interleaved samples the way the image crate's decoder does (EXPAND)
    const size_t npx = (size_t)hd.w * hd.h;
    auto sample = [&](size_t y, size_t x) -> int {  // sub-byte sample of a 1-sample-per-pixel image
        const int d = hd.depth, per = 8 / d;
        return (px[y * full_stride + x / per] >> (8 - d - (int)(x % per) * d)) & ((1 << d) - 1);
    };
    if (hd.color == 3) {
        const bool alpha = have_trns;
        CH = alpha ? 4 : 3;
        out.resize(npx * CH);
        const size_t ncol = plte.size() / 3;
        for (size_t y = 0; y < hd.h; ++y)
            for (size_t x = 0; x < hd.w; ++x) {
                const size_t idx = hd.depth == 8 ? px[y * full_stride + x] : (size_t)sample(y, x);
                if (idx >= ncol) KC_FAIL(KC_ERR_IMAGE, "palette index out of range");
                uint8_t* o = &out[(y * hd.w + x) * CH];
                o[0] = plte[idx * 3]; o[1] = plte[idx * 3 + 1]; o[2] = plte[idx * 3 + 2];
                if (alpha) o[3] = idx < trns.size() ? trns[idx] : 255;
            }
    } else if (hd.color == 0) {
        const bool alpha = have_trns && trns.size() >= 2;
        const int key = alpha ? ((trns[0] << 8) | trns[1]) : -1;
        CH = alpha ? 2 : 1;
        out.resize(npx * CH);
        const int scale = hd.depth == 8 ? 1 : 255 / ((1 << hd.depth) - 1);
        for (size_t y = 0; y < hd.h; ++y)
            for (size_t x = 0; x < hd.w; ++x) {
                const int v = hd.depth == 8 ? px[y * full_stride + x] : sample(y, x);
                uint8_t* o = &out[(y * hd.w + x) * CH];
                o[0] = (uint8_t)(v * scale);
                if (alpha) o[1] = v == key ? 0 : 255;
            }
    } else if (hd.color == 2) {
        const bool alpha = have_trns && trns.size() >= 6;
        CH = alpha ? 4 : 3;
        if (!alpha) out = std::move(px);
        else {
            out.resize(npx * 4);
            const uint8_t kr = trns[1], kg = trns[3], kb = trns[5];
            for (size_t i = 0; i < npx; ++i) {
                const uint8_t* s = &px[i * 3];
                uint8_t* o = &out[i * 4];
                o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
                o[3] = (s[0] == kr && s[1] == kg && s[2] == kb) ? 0 : 255;
            }
        }
    } else {  // 4: gray + alpha, 6: RGBA
        CH = hd.color == 4 ? 2 : 4;
        out = std::move(px);
    }
    W = hd.w;
    H = hd.h;
    return KC_OK;
}

void chunk(std::vector<uint8_t>& f, const char* type, const uint8_t* body, size_t len) {
    put32(f, (uint32_t)len);
    const size_t at = f.size();
    f.insert(f.end(), type, type + 4);
    if (len) f.insert(f.end(), body, body + len);
    put32(f, (uint32_t)crc32(0, &f[at], (uInt)(4 + len)));
}

// files come from outside: running out of memory on one is an error code, not an exception through the C ABI
int32_t decode(const uint8_t* data, size_t n, std::vector<uint8_t>& out, uint32_t& W, uint32_t& H, uint32_t& CH) {
    try {
        return decode_impl(data, n, out, W, H, CH);
    } catch (const std::bad_alloc&) {
        KC_FAIL(KC_ERR_IMAGE, "out of memory while decoding a PNG");
    } catch (const std::length_error&) {
        KC_FAIL(KC_ERR_IMAGE, "PNG too large to decode");
    }
}

// RGBA8 (or any `ch` bytes per pixel) -> PNG; filter per scanline by the minimum-sum-of-
// absolute-differences heuristic over the five filter types, deflate level 6
int32_t encode(const uint8_t* px, uint32_t w, uint32_t h, int ch, std::vector<uint8_t>& f) {
    if (!px || w == 0 || h == 0) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "empty image");
    const int color = ch == 1 ? 0 : ch == 2 ? 4 : ch == 3 ? 2 : 6;
    const size_t stride = (size_t)w * ch, bpp = (size_t)ch;
    std::vector<uint8_t> raw((size_t)h * (stride + 1));
    std::vector<uint8_t> zero(stride, 0), cand(stride);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* cur = px + (size_t)y * stride;
        const uint8_t* prev = y ? px + (size_t)(y - 1) * stride : zero.data();
        uint8_t* dst = &raw[(size_t)y * (stride + 1)];
        unsigned long best = ~0ul;
        for (int ft = 0; ft < 5; ++ft) {
            unsigned long sum = 0;
            for (size_t i = 0; i < stride; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int v = cur[i];
                switch (ft) {
                    case 1: v -= a; break;
                    case 2: v -= b; break;
                    case 3: v -= (a + b) >> 1; break;
                    case 4: v -= paeth(a, b, c); break;
                    default: break;
                }
                cand[i] = (uint8_t)v;
                sum += (unsigned long)abs((int)(int8_t)cand[i]);
            }
            if (sum < best) {
                best = sum;
                dst[0] = (uint8_t)ft;
                memcpy(dst + 1, cand.data(), stride);
            }
        }
    }
    uLongf zn = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zn);
    if (compress2(z.data(), &zn, raw.data(), (uLong)raw.size(), 6) != Z_OK) KC_FAIL(KC_ERR_GENERIC, "deflate failed");
    f.clear();
    f.insert(f.end(), kSig, kSig + 8);
    uint8_t ihdr[13];
    ihdr[0] = (uint8_t)(w >> 24); ihdr[1] = (uint8_t)(w >> 16); ihdr[2] = (uint8_t)(w >> 8); ihdr[3] = (uint8_t)w;
    ihdr[4] = (uint8_t)(h >> 24); ihdr[5] = (uint8_t)(h >> 16); ihdr[6] = (uint8_t)(h >> 8); ihdr[7] = (uint8_t)h;
    ihdr[8] = 8; ihdr[9] = (uint8_t)color; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk(f, "IHDR", ihdr, 13);
    chunk(f, "IDAT", z.data(), zn);
    chunk(f, "IEND", nullptr, 0);
    return KC_OK;
}

int32_t read_file(const char* path, std::vector<uint8_t>& data) {
    FILE* fp = fopen(path, "rb");
    if (!fp) KC_FAIL(KC_ERR_IO, "cannot open %s", path);
    fseek(fp, 0, SEEK_END);
    const long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    data.resize(n > 0 ? (size_t)n : 0);
    const size_t got = data.empty() ? 0 : fread(data.data(), 1, data.size(), fp);
    fclose(fp);
    if (got != data.size()) KC_FAIL(KC_ERR_IO, "short read on %s", path);
    return KC_OK;
}

}  // namespace

int32_t kc_png_decode_vec(const uint8_t* data, size_t n, std::vector<uint8_t>& samples, uint32_t& w, uint32_t& h, uint32_t& ch) try {
    return decode(data, n, samples, w, h, ch);
} KC_ABI_CATCH
int32_t kc_png_decode_file_vec(const char* path, std::vector<uint8_t>& samples, uint32_t& w, uint32_t& h, uint32_t& ch) try {
    std::vector<uint8_t> data;
    KC_TRY(read_file(path, data));
    return decode(data.data(), data.size(), samples, w, h, ch);
} KC_ABI_CATCH
int32_t kc_png_write_file(const char* path, const uint8_t* px, uint32_t w, uint32_t h, int ch) try {
    std::vector<uint8_t> f;
    KC_TRY(encode(px, w, h, ch, f));
    FILE* fp = fopen(path, "wb");
    if (!fp) KC_FAIL(KC_ERR_IO, "cannot create %s", path);
    const size_t put = fwrite(f.data(), 1, f.size(), fp);
    if (fclose(fp) != 0 || put != f.size()) KC_FAIL(KC_ERR_IO, "short write on %s", path);
    return KC_OK;
} KC_ABI_CATCH

extern "C" {

int32_t kc_png_decode(const uint8_t* data, size_t n, uint8_t** samples, uint32_t* w, uint32_t* h, uint32_t* channels) try {
    if (!data || !samples || !w || !h || !channels) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    std::vector<uint8_t> v;
    KC_TRY(decode(data, n, v, *w, *h, *channels));
    *samples = (uint8_t*)malloc(v.size());
    if (!*samples) KC_FAIL(KC_ERR_GENERIC, "out of memory");
    memcpy(*samples, v.data(), v.size());
    return KC_OK;
} KC_ABI_CATCH

int32_t kc_png_decode_file(const char* path, uint8_t** samples, uint32_t* w, uint32_t* h, uint32_t* channels) try {
    if (!path) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "path is NULL");
    std::vector<uint8_t> data;
    KC_TRY(read_file(path, data));
    return kc_png_decode(data.data(), data.size(), samples, w, h, channels);
} KC_ABI_CATCH

int32_t kc_png_encode(const uint8_t* samples, uint32_t w, uint32_t h, uint32_t channels, uint8_t** png, size_t* n) try {
    if (!samples || !png || !n || channels < 1 || channels > 4) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "bad argument");
    std::vector<uint8_t> f;
    KC_TRY(encode(samples, w, h, (int)channels, f));
    *png = (uint8_t*)malloc(f.size());
    if (!*png) KC_FAIL(KC_ERR_GENERIC, "out of memory");
    memcpy(*png, f.data(), f.size());
    *n = f.size();
    return KC_OK;
} KC_ABI_CATCH

int32_t kc_png_encode_file(const char* path, const uint8_t* samples, uint32_t w, uint32_t h, uint32_t channels) try {
    if (!path || !samples || channels < 1 || channels > 4) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "bad argument");
    return kc_png_write_file(path, samples, w, h, (int)channels);
} KC_ABI_CATCH

}  // extern "C"
