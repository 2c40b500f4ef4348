// pm_mapio.cpp -- depth / confidence map files either side of the hot path: PFM and COLMAP .bin,
// read straight into / written straight from caller-owned (pinned) host buffers, so that a map goes
// file -> pinned buffer -> cudaMemcpyAsync -> kernel (and back) with no intermediate copy and no Python-level packing.
//
// Replaces, byte for byte on disk and value for value in memory, reference datasets/data_io.py:
//     read_pfm  :257-288    save_pfm :291-302 (+ the rest of the function)     read_bin :165-191     save_bin :194-223
// The reference builds a Python list of every float and struct.pack()s it (save_bin), walks the header byte by byte in the
// interpreter (read_bin) and flips / transposes through temporaries; here the row flip of PFM is folded into the
// scatter/gather I/O vectors (preadv / writev) and the planar <-> interleaved transposition of multi-channel .bin maps
// runs through one bounded staging block.
//
// Host code only (no kernels): it is compiled into libpmb200.so by the same nvcc invocation as the .cu files.
// In-memory layout is always [H, W, C] row-major, top row first -- what read_pfm / read_bin return.
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/patchmatch_b200.h"

extern "C" int pmb200_internal_fail(int code, const char *msg);

namespace {

int failf(int code, const char *fmt, const char *a, const char *b = "") {
    char buf[480];
    snprintf(buf, sizeof buf, fmt, a, b);
    return pmb200_internal_fail(code, buf);
}

struct Fd {
    int fd = -1;
    ~Fd() { if (fd >= 0) close(fd); }
};

bool read_fully(int fd, void *dst, size_t n, off_t off) {
    char *p = static_cast<char *>(dst);
    while (n) {
        const ssize_t r = pread(fd, p, n, off);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return false;
        p += r; off += r; n -= (size_t)r;
    }
    return true;
}

bool write_fully(int fd, const void *src, size_t n) {
    const char *p = static_cast<const char *>(src);
    while (n) {
        const ssize_t r = write(fd, p, n);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return false;
        p += r; n -= (size_t)r;
    }
    return true;
}

// gather / scatter a list of equally sized rows; handles partial transfers and IOV_MAX
template <bool kWrite>
bool rows_io(int fd, char *const *rows, size_t n_rows, size_t row_bytes, off_t off) {
    const size_t kBatch = 512;
    struct iovec iov[kBatch];
    size_t done = 0;
    while (done < n_rows) {
        const size_t cnt = std::min(kBatch, n_rows - done);
        for (size_t i = 0; i < cnt; ++i) { iov[i].iov_base = rows[done + i]; iov[i].iov_len = row_bytes; }
        size_t first = 0, total = cnt * row_bytes;
        while (total) {
            const ssize_t r = kWrite ? writev(fd, iov + first, (int)(cnt - first)) : preadv(fd, iov + first, (int)(cnt - first), off);
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) return false;
            off += r; total -= (size_t)r;
            size_t left = (size_t)r;
            while (left && first < cnt) {
                if (left >= iov[first].iov_len) { left -= iov[first].iov_len; ++first; }
                else { iov[first].iov_base = static_cast<char *>(iov[first].iov_base) + left; iov[first].iov_len -= left; left = 0; }
            }
        }
        done += cnt;
    }
    return true;
}

// one header line (up to and including '\n', or to the end of the header block)
bool next_line(const char *buf, size_t len, size_t &pos, std::string &line) {
    if (pos >= len) return false;
    const void *nl = memchr(buf + pos, '\n', len - pos);
    const size_t end = nl ? (size_t)(static_cast<const char *>(nl) - buf) + 1 : len;
    line.assign(buf + pos, end - pos);
    pos = end;
    return true;
}

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

std::string rstrip(std::string s) {
    while (!s.empty() && is_space(s.back())) s.pop_back();
    return s;
}

// re.match(r"^(\d+)\s(\d+)\s$", line): digits, ONE whitespace, digits, ONE whitespace, end (or end before a final '\n')
bool parse_dims(const std::string &line, long long &w, long long &h) {
    size_t i = 0;
    auto digits = [&](long long &v) {
        const size_t s = i;
        v = 0;
        while (i < line.size() && isdigit((unsigned char)line[i])) {
            if (v > (1ll << 40)) return false;
            v = v * 10 + (line[i++] - '0');
        }
        return i > s;
    };
    if (!digits(w) || i >= line.size() || !is_space(line[i++])) return false;
    if (!digits(h) || i >= line.size() || !is_space(line[i++])) return false;
    return i == line.size() || (i + 1 == line.size() && line[i] == '\n');
}

inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

int pfm_probe(int fd, const char *path, pmb200_map_info *info, off_t file_size) {
    char head[256];
    const size_t got = (size_t)std::max<ssize_t>(0, pread(fd, head, sizeof head, 0));
    size_t pos = 0;
    std::string line;
    if (!next_line(head, got, pos, line)) line.clear();
    const std::string magic = rstrip(line);
    if (magic == "PF") info->channels = 3;
    else if (magic == "Pf") info->channels = 1;
    else return failf(PMB200_EFORMAT, "Not a PFM file.%s%s", "", "");
    long long w = 0, h = 0;
    if (!next_line(head, got, pos, line) || !parse_dims(line, w, h) || w > INT_MAX || h > INT_MAX)
        return failf(PMB200_EFORMAT, "Malformed PFM header.%s%s", "", "");
    if (!next_line(head, got, pos, line)) line.clear();
    const std::string tok = rstrip(line);
    char *endp = nullptr;
    errno = 0;
    double scale = tok.empty() ? 0.0 : strtod(tok.c_str(), &endp);
    if (tok.empty() || !endp || *endp != '\0')
        return failf(PMB200_EFORMAT, "could not convert string to float: '%s'%s", tok.c_str());
    info->big_endian = !(scale < 0);  // data_io.py:275-279 (a NaN or zero scale reads as big-endian, as there)
    info->scale = scale < 0 ? -scale : scale;
    info->width = (int)w; info->height = (int)h;
    info->data_offset = (int64_t)pos;
    info->payload_floats = ((int64_t)file_size - (int64_t)pos) / 4;  // np.fromfile drops a trailing partial item
    if (info->payload_floats < 0) info->payload_floats = 0;
    (void)path;
    return 0;
}

int bin_probe(int fd, const char *path, pmb200_map_info *info, off_t file_size) {
    char head[128];
    const size_t got = (size_t)std::max<ssize_t>(0, pread(fd, head, sizeof head, 0));
    long long v[3] = {0, 0, 0};
    size_t i = 0;
    for (int k = 0; k < 3; ++k) {  // "width&height&channels&" (data_io.py:175-186)
        while (i < got && (head[i] == ' ' || head[i] == '\t')) ++i;
        const size_t s = i;
        while (i < got && isdigit((unsigned char)head[i]) && v[k] < (1ll << 40)) v[k] = v[k] * 10 + (head[i++] - '0');
        if (i == s) return failf(PMB200_EFORMAT, "Malformed COLMAP .bin header in %s%s", path);
        while (i < got && (head[i] == ' ' || head[i] == '\t')) ++i;
        if (i >= got || head[i] != '&') return failf(PMB200_EFORMAT, "Malformed COLMAP .bin header in %s%s", path);
        ++i;
    }
    if (v[0] > INT_MAX || v[1] > INT_MAX || v[2] > INT_MAX) return failf(PMB200_EFORMAT, "Malformed COLMAP .bin header in %s%s", path);
    info->width = (int)v[0]; info->height = (int)v[1]; info->channels = (int)v[2];
    info->scale = 1.0; info->big_endian = 0;
    info->data_offset = (int64_t)i;
    info->payload_floats = ((int64_t)file_size - (int64_t)i) / 4;
    if (info->payload_floats < 0) info->payload_floats = 0;
    return 0;
}

int probe_open(const char *path, int format, pmb200_map_info *info, Fd &f) {
    if (!path || !info) return pmb200_internal_fail(PMB200_EINVAL, "map: null pointer");
    if (format != PMB200_MAP_PFM && format != PMB200_MAP_COLMAP_BIN)
        return pmb200_internal_fail(PMB200_EINVAL, "Invalid input format; only pfm and bin are supported");
    f.fd = open(path, O_RDONLY | O_CLOEXEC);
    if (f.fd < 0) return failf(PMB200_EIO, "%s: '%s'", strerror(errno), path);
    struct stat st;
    if (fstat(f.fd, &st) != 0) return failf(PMB200_EIO, "%s: '%s'", strerror(errno), path);
    memset(info, 0, sizeof *info);
    info->format = format;
    return format == PMB200_MAP_PFM ? pfm_probe(f.fd, path, info, st.st_size) : bin_probe(f.fd, path, info, st.st_size);
}

int shape_mismatch(const pmb200_map_info &info) {
    char buf[200];
    if (info.format == PMB200_MAP_PFM)
        snprintf(buf, sizeof buf, "cannot reshape array of size %lld into shape (%d,%d,%d)", (long long)info.payload_floats,
                 info.height, info.width, info.channels);
    else
        snprintf(buf, sizeof buf, "cannot reshape array of size %lld into shape (%d,%d,%d)", (long long)info.payload_floats,
                 info.width, info.height, info.channels);
    return pmb200_internal_fail(PMB200_EFORMAT, buf);
}

}  // namespace

extern "C" int pmb200_map_probe(const char *path, int format, pmb200_map_info *info) {
    Fd f;
    return probe_open(path, format, info, f);
}

extern "C" int pmb200_map_read(const char *path, int format, float *out_host, int64_t capacity_floats, pmb200_map_info *info_out) {
    float *out = out_host;
    Fd f;
    pmb200_map_info info;
    const int rc = probe_open(path, format, &info, f);
    if (info_out && rc == 0) *info_out = info;
    if (rc != 0) return rc;
    const int64_t need = (int64_t)info.width * info.height * info.channels;
    if (info.payload_floats != need) return shape_mismatch(info);  // np.reshape raises ValueError
    if (need == 0) return 0;
    if (!out || capacity_floats < need) return pmb200_internal_fail(PMB200_EINVAL, "map_read: output buffer too small");
    const int H = info.height, W = info.width, C = info.channels;
    if (format == PMB200_MAP_PFM) {
        // file rows run bottom-up (np.flipud, data_io.py:285): scatter them straight into place
        const size_t row_bytes = (size_t)W * C * 4;
        std::vector<char *> rows((size_t)H);
        for (int r = 0; r < H; ++r) rows[(size_t)r] = reinterpret_cast<char *>(out) + (size_t)(H - 1 - r) * row_bytes;
        if (!rows_io<false>(f.fd, rows.data(), rows.size(), row_bytes, (off_t)info.data_offset))
            return failf(PMB200_EIO, "short read: '%s'%s", path);
        if (info.big_endian) {
            uint32_t *p = reinterpret_cast<uint32_t *>(out);
            for (int64_t i = 0; i < need; ++i) p[i] = bswap32(p[i]);
        }
        return 0;
    }
    // COLMAP .bin: the payload is [C][H][W] (reshape((w,h,c), order='F') + transpose, data_io.py:188-190)
    if (C == 1) {
        if (!read_fully(f.fd, out, (size_t)need * 4, (off_t)info.data_offset)) return failf(PMB200_EIO, "short read: '%s'%s", path);
        return 0;
    }
    const int64_t plane = (int64_t)H * W;
    const int64_t block = std::max<int64_t>(1, std::min<int64_t>(plane, (1 << 18) / C));  // <= 1 MiB of staging
    std::vector<float> stage((size_t)(block * C));
    for (int64_t p0 = 0; p0 < plane; p0 += block) {
        const int64_t n = std::min(block, plane - p0);
        for (int c = 0; c < C; ++c)
            if (!read_fully(f.fd, stage.data() + (size_t)c * block, (size_t)n * 4, (off_t)(info.data_offset + ((int64_t)c * plane + p0) * 4)))
                return failf(PMB200_EIO, "short read: '%s'%s", path);
        for (int64_t i = 0; i < n; ++i)
            for (int c = 0; c < C; ++c) out[(p0 + i) * C + c] = stage[(size_t)(c * block + i)];
    }
    return 0;
}

extern "C" int pmb200_map_write(const char *path, int format, const float *data_host, int height, int width, int channels, double scale) {
    const float *data = data_host;
    if (!path) return pmb200_internal_fail(PMB200_EINVAL, "map: null pointer");
    if (format != PMB200_MAP_PFM && format != PMB200_MAP_COLMAP_BIN)
        return pmb200_internal_fail(PMB200_EINVAL, "Invalid input format; only pfm and bin are supported");
    if (height < 0 || width < 0 || (channels != 1 && channels != 3))
        return pmb200_internal_fail(PMB200_EINVAL, "Image must have H x W x 3, H x W x 1 or H x W dimensions.");
    const int64_t count = (int64_t)height * width * channels;
    if (count > 0 && !data) return pmb200_internal_fail(PMB200_EINVAL, "map: null pointer");
    Fd f;
    f.fd = open(path, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
    if (f.fd < 0) return failf(PMB200_EIO, "%s: '%s'", strerror(errno), path);
    char head[128];
    int n;
    if (format == PMB200_MAP_PFM)  // data_io.py:291-302 tail: "Pf\n" / "PF\n", "W H\n", "%f\n" of -scale on a little-endian host
        n = snprintf(head, sizeof head, "%s\n%d %d\n%f\n", channels == 3 ? "PF" : "Pf", width, height, -scale);
    else                           // data_io.py:211-212
        n = snprintf(head, sizeof head, "%d&%d&%d&", width, height, channels);
    if (!write_fully(f.fd, head, (size_t)n)) return failf(PMB200_EIO, "%s: '%s'", strerror(errno), path);
    bool ok = true;
    if (count == 0) {
    } else if (format == PMB200_MAP_PFM) {
        const size_t row_bytes = (size_t)width * channels * 4;
        std::vector<char *> rows((size_t)height);
        for (int r = 0; r < height; ++r)
            rows[(size_t)r] = const_cast<char *>(reinterpret_cast<const char *>(data)) + (size_t)(height - 1 - r) * row_bytes;
        ok = rows_io<true>(f.fd, rows.data(), rows.size(), row_bytes, 0);
    } else if (channels == 1) {
        ok = write_fully(f.fd, data, (size_t)count * 4);
    } else {
        const int64_t plane = (int64_t)height * width;
        const int64_t block = std::min<int64_t>(plane, 1 << 18);
        std::vector<float> stage((size_t)block);
        for (int c = 0; c < channels && ok; ++c)
            for (int64_t p0 = 0; p0 < plane && ok; p0 += block) {
                const int64_t m = std::min(block, plane - p0);
                for (int64_t i = 0; i < m; ++i) stage[(size_t)i] = data[(p0 + i) * channels + c];
                ok = write_fully(f.fd, stage.data(), (size_t)m * 4);
            }
    }
    if (!ok) return failf(PMB200_EIO, "%s: '%s'", strerror(errno), path);
    const int fd = f.fd;
    f.fd = -1;
    if (close(fd) != 0) return failf(PMB200_EIO, "%s: '%s'", strerror(errno), path);
    return 0;
}
